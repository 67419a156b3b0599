#!/usr/bin/env python
"""Turn one gpurun_out/<tag>/ directory (written by tools/gpu_round.sh) into the tracked summaries under profiles/:
   profiles/<tag>_launches_<what>.csv   per-launch device time (ncu --metrics gpu__time_duration.sum), aggregated per kernel
   profiles/<tag>_ncu_<what>.csv        selected `ncu --set full` counters per profiled launch
   profiles/<tag>_bench.json            the bench lines of that run
Usage: python tools/summarise_profiles.py r01a
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]


def launches(path, out):
    rows = list(csv.reader(open(path, errors="replace")))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"]
    if not h:
        return
    hdr = rows[h[0]]
    kn, mv, gs, bs = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg, order, total = {}, [], 0.0
    for r in rows[h[0] + 1:]:
        if len(r) != len(hdr):
            continue
        name = r[kn].split("(")[0].replace("void ", "")
        ns = float(r[mv].replace(",", ""))
        if name not in agg:
            agg[name] = [0, 0.0, 1e30, 0.0]
            order.append(name)
        a = agg[name]
        a[0] += 1; a[1] += ns; a[2] = min(a[2], ns); a[3] = max(a[3], ns)
        total += ns
    with open(out, "w") as f:
        f.write("# source: %s  (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised launches: compare SHARES)\n" % os.path.relpath(path, ROOT))
        f.write("kernel,launches,total_us,share_pct,min_us,max_us\n")
        for name in sorted(order, key=lambda n: -agg[n][1]):
            a = agg[name]
            f.write("%s,%d,%.2f,%.2f,%.2f,%.2f\n" % (name, a[0], a[1] / 1e3, 100 * a[1] / total, a[2] / 1e3, a[3] / 1e3))


def full(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        return
    hdr, units = rows[0], rows[1]
    cols = [hdr.index("Kernel Name")] + [hdr.index(w) for w in WANT if w in hdr]
    with open(out, "w") as f:
        f.write("# source: %s  (ncu --set full --clock-control none --import-source on)\n" % os.path.relpath(rep, ROOT))
        w = csv.writer(f)
        w.writerow([hdr[c] + (" [%s]" % units[c] if units[c] else "") for c in cols])
        for r in rows[2:]:
            w.writerow([r[c].split("(")[0].replace("void ", "") if c == cols[0] else r[c] for c in cols])


LIB_NAMES = {"bp_fwd_kernel": "bp_fwd", "bp_fwd8_kernel": "bp_fwd", "bp_stats_kernel": "bp_fwd_stats",
             "bp_fwd_finish_kernel": "bp_fwd_finish", "bp_prep_kernel": "bp_prep", "bp_prep4_kernel": "bp_prep",
             "bp_bwd_fill_ghat_kernel": "bp_bwd_fill_ghat", "bp_bwd_order_kernel": "bp_bwd_order",
             "bp_bwd_gather_tile_kernel": "bp_bwd_gather", "transpose_maps_kernel": "relayout_transpose"}


def traffic(rep, out, tag):
    """profiles/roofline_traffic.json: per library kernel name, DRAM read+write bytes of its launch at level 0, 1, 2 of
    the headline fragment step (launch order inside one captured step = level order)."""
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        return
    hdr, units = rows[0], rows[1]
    kn, rd, wr = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per = {}
    for r in rows[2:]:
        name = r[kn].split("(")[0].replace("void ", "").replace("d3m::", "").split("<")[0]
        if name == "bp_bwd_sample_kernel":
            name = "bp_bwd_fill" if ", 1>" in r[kn].replace("(bool)", "").replace("(int)", "") or "true" in r[kn] else "bp_bwd_hist"
        else:
            name = LIB_NAMES.get(name, name)
        b = float(r[rd]) * scale.get(units[rd], 1.0) + float(r[wr]) * scale.get(units[wr], 1.0)
        per.setdefault(name, []).append(int(b))
    json.dump({"source": "profiles/%s_ncu_bp.csv (ncu --set full of `bench.py --profile-step bp`, one flushed step)" % tag,
               "unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)", "kernels": per},
              open(out, "w"), indent=1)


def main():
    tag = sys.argv[1]
    src = os.path.join(ROOT, "gpurun_out", tag)
    dst = os.path.join(ROOT, "profiles")
    os.makedirs(dst, exist_ok=True)
    for fn in sorted(os.listdir(src)):
        p = os.path.join(src, fn)
        if fn.startswith("launches_") and fn.endswith(".csv"):
            launches(p, os.path.join(dst, "%s_%s" % (tag, fn)))
        elif fn.endswith(".ncu-rep"):
            full(p, os.path.join(dst, "%s_ncu_%s.csv" % (tag, fn[:-8].replace("prof_", ""))))
            if fn == "prof_bp.ncu-rep":
                traffic(p, os.path.join(dst, "roofline_traffic.json"), tag)
        elif fn.startswith("bench") and fn.endswith(".json") and os.path.getsize(p):
            lines = [json.loads(l) for l in open(p) if l.strip().startswith("{")]
            json.dump(lines if len(lines) != 1 else lines[0], open(os.path.join(dst, "%s_%s" % (tag, fn)), "w"), indent=1)
        elif fn in ("gpu.txt", "pytest_gpu.log", "smoke.log", "tsdf_host_profile.txt"):
            open(os.path.join(dst, "%s_%s" % (tag, fn)), "w").write(open(p).read())
    print(sorted(os.listdir(dst)))


if __name__ == "__main__":
    main()

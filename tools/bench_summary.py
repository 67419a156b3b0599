"""Prints the figures of a bench.py JSON line that the round-2 work tracks.  python tools/bench_summary.py <bench.json>"""
import json
import sys

_text = open(sys.argv[1]).read().strip()
try:
    p = json.loads(_text)                       # pretty-printed copy under profiles/
except ValueError:
    p = json.loads(_text.splitlines()[-1])      # raw bench output: the JSON line is the last one
print("value %.3f G samples/s   ms/step graph %.4f eager %.4f host-issue %.4f   e2e %.3f ms   roofline.frac %.3f" % (
    p["value"] / 1e9, p["ms_per_step"], p["ms_per_step_eager"], p["host_issue_ms_per_step"], p["e2e"]["ms_per_step"],
    p["roofline"]["frac"]))
for lv, k in enumerate(p["kernel_us_per_level"]):
    print("  L%d" % lv, " ".join("%s=%.1f" % (n.replace("bp_", ""), v) for n, v in k.items()), " sum=%.0f" % sum(k.values()))
d = p.get("dense_level2")
if d:
    print("dense L2 %.4f ms  frac %.3f " % (d["ms_fwd_bwd"], d["frac_of_measured_hbm_peak"]),
          " ".join("%s=%.0f" % (n.replace("bp_", ""), 1e3 * v) for n, v in d["kernel_ms"].items()))
s = p.get("large_scene")
if s:
    print("large scene %.3f ms " % s["ms_per_step"], " ".join("%s=%.2f" % (n.replace("bp_", ""), v) for n, v in s["kernel_ms_rank0"].items()))
b = p.get("batched_fragments")
if b:
    print("batched 64 fragments %.3f ms  frac %.3f" % (b["ms_per_step"], b["frac_of_measured_hbm_peak_per_gpu"]))
t = p.get("tsdf")
if t:
    print("tsdf batch300 %.4f ms  frac %.3f  per-call resident %.0f f/s  e2e %.0f f/s  datagen3 %.0f f/s" % (
        t["ms_batch_300"], t["roofline"]["frac"], t["frames_per_s_per_call_resident"], t["e2e_frames_per_s"],
        t["e2e_datagen_3level_fps"]))
r = p.get("reference_gpu")
if r and "back_project" in r and "fragment" in r["back_project"]:
    print("reference on this GPU: fragment %.2f ms (x%.1f), dense L2 %.2f ms (x%.1f), tsdf %.0f f/s" % (
        r["back_project"]["fragment"]["ms_per_step"], r["back_project"]["fragment"]["speedup_ours"],
        r["back_project"]["dense_level2"]["ms_fwd_bwd"], r["back_project"]["dense_level2"]["speedup_ours"],
        r["tsdf"].get("frames_per_s", 0)))

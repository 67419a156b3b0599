#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares of one profiled kernel.
The ncu CLI prints metrics only per SASS instruction; this joins them with the line table of the cubin
(`nvdisasm -g`) by instruction order.   python tools/ncu_lines.py <file.ncu-rep> <kernel-id e.g. :::1> <so> <mangled-substr>"""
import csv, io, re, subprocess, sys, os, tempfile, collections

rep, kid, so, sub = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
k = int(kid.strip(":") or 1) - 1      # N-th profiled launch in the report (1-based)
hi = his[k]
end = his[k + 1] - 1 if k + 1 < len(his) else len(rows)
hdr = rows[hi]
ie, ss, src = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source")
sass = [(r[src].strip(), int(r[ie] or 0), int(r[ss] or 0)) for r in rows[hi + 1:end] if len(r) == len(hdr)]
print(rows[hi - 1][:2] if hi else "", len(sass), "SASS instructions")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, capture_output=True)
lines = None
for f in os.listdir(tmp):
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    # split per function
    parts = re.split(r"\n\s*\.text\.", out)
    for p in parts:
        name = p.split(":", 1)[0].strip()
        if sub in name:
            cur, lst = None, []
            for ln in p.splitlines():
                m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
                if m:
                    cur = (os.path.basename(m.group(1)), int(m.group(2)))
                    continue
                if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", ln):
                    lst.append(cur)
            if lines is None or abs(len(lst) - len(sass)) < abs(len(lines) - len(sass)):
                lines = lst
print("line table:", len(lines) if lines else None)
n = min(len(lines), len(sass))
agg = collections.defaultdict(lambda: [0, 0])
for (s, a, b), l in zip(sass[:n], lines[:n]):
    agg[l][0] += a; agg[l][1] += b
ti, ts = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
srcs = {}
for (fn, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top] if agg else []:
    path = os.path.join(os.path.dirname(os.path.abspath(so)), "csrc", fn) if fn else None
    if path and path not in srcs and os.path.exists(path):
        srcs[path] = open(path).read().splitlines()
    text = srcs.get(path, [""] * 100000)[ln - 1].strip()[:100] if path in srcs else ""
    print("%5.1f%% inst %5.1f%% stall  %s:%d  %s" % (100 * v[0] / max(ti, 1), 100 * v[1] / max(ts, 1), fn, ln, text))

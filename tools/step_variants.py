"""Headline fragment step (3 levels, sparse coords, fwd+bwd) under a list of environment settings (run on the GPU box).

    python tools/step_variants.py "" "D3M_PDL=0" "D3M_FWD_TVMIN=4;D3M_FWD_KU=4" ...

Each variant runs in its own process (the library reads its tuning switches once).  Per variant one line:
graph-replayed ms per step (L2 flushed between steps, like bench.py), eager ms, and per-level per-kernel device
times (us, CUDA events inside the library, serialised).
"""
import os, subprocess, sys

code = r'''
import numpy as np, torch, sys, os, json
sys.path.insert(0, os.getcwd())
import bench
from deep3dmap_b200 import back_project, _lib
dev = torch.device("cuda:0")
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
def cnt_fn(inp):
    return back_project(t(inp["coords"]), t(inp["origin"]), inp["voxel_size"], t(inp["feats"]), t(inp["KRcam"]))[1].cpu().numpy()
levels = bench.build_fragment_levels(cnt_fn)
dl = [dict(coords=t(i["coords"]), origin=t(i["origin"]), vs=i["voxel_size"], feats=t(i["feats"]).requires_grad_(True),
           KR=t(i["KRcam"]), go=t(i["grad_out"])) for i in levels]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def step_serial():
    for d in dl:
        d["feats"].grad = None
        vol, cnt = back_project(d["coords"], d["origin"], d["vs"], d["feats"], d["KR"])
        vol.backward(d["go"])
# STEP_PRIO=1: the finest level's stream (the critical path of the three-branch graph) gets high priority
streams = [torch.cuda.Stream(priority=(-1 if (os.environ.get("STEP_PRIO") == "1" and i == len(dl) - 1) else 0)) for i in range(len(dl))]
def step_streams():      # STEP_STREAMS=1: one stream per level, largest level first (bench.py step_level_streams)
    cur = torch.cuda.current_stream()
    for d, st in sorted(zip(dl, streams), key=lambda x: -x[0]["coords"].shape[0]):
        st.wait_stream(cur)
        with torch.cuda.stream(st):
            d["feats"].grad = None
            vol, cnt = back_project(d["coords"], d["origin"], d["vs"], d["feats"], d["KR"])
            vol.backward(d["go"])
    for st in streams: cur.wait_stream(st)
def step_one_backward():  # STEP_ONEBWD=1: three forwards, one autograd pass (bench.py step_one_backward)
    for d in dl: d["feats"].grad = None
    vols = [back_project(d["coords"], d["origin"], d["vs"], d["feats"], d["KR"])[0] for d in dl]
    torch.autograd.backward(vols, [d["go"] for d in dl])
step = step_streams if os.environ.get("STEP_STREAMS") == "1" else step_one_backward if os.environ.get("STEP_ONEBWD") == "1" else step_serial
def timed(fn, reps=30):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    torch.cuda.synchronize()
    for a, b in ev:
        flush.fill_(1); a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return round(float(np.mean(ts)), 4), round(ts[len(ts) // 2], 4), round(ts[0], 4)
for _ in range(3): step()
eager = timed(step)
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side): step()
torch.cuda.current_stream().wait_stream(side)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g): step()
for _ in range(3): g.replay()
graph = timed(g.replay)
per = []
for d in dl:
    acc = {}
    for _ in range(10):
        flush.fill_(1); torch.cuda.synchronize(); _lib.profile_begin()
        d["feats"].grad = None
        vol, cnt = back_project(d["coords"], d["origin"], d["vs"], d["feats"], d["KR"]); vol.backward(d["go"])
        for k, v in _lib.profile_end().items(): acc[k] = acc.get(k, 0.0) + v["ms"] * 100
    per.append({k.replace("bp_", ""): round(v, 1) for k, v in sorted(acc.items())})
print(json.dumps({"variant": os.environ.get("D3M_VARIANT", ""), "graph_ms(mean,med,min)": graph, "eager_ms": eager,
                  "kernel_us": per}), flush=True)
'''
for var in sys.argv[1:] or [""]:
    env = dict(os.environ, D3M_VARIANT=var)
    for kv in filter(None, var.split(";")):
        k, v = kv.split("=", 1)
        env[k] = v
    subprocess.run([sys.executable, "-c", code], env=env)

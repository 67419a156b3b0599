"""cProfile of the eager python path of one fragment step (3 levels fwd+bwd) -- where do the host microseconds go?"""
import cProfile, pstats, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from deep3dmap_b200 import back_project, synth
dev = torch.device("cuda:0")
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
def cnt_fn(inp):
    return back_project(t(inp["coords"]), t(inp["origin"]), inp["voxel_size"], t(inp["feats"]), t(inp["KRcam"]))[1].cpu().numpy()
levels = bench.build_fragment_levels(cnt_fn)
dl = [dict(coords=t(i["coords"]), origin=t(i["origin"]), vs=i["voxel_size"], feats=t(i["feats"]).requires_grad_(True), KR=t(i["KRcam"]), go=t(i["grad_out"])) for i in levels]
def step():
    for d in dl:
        d["feats"].grad = None
        vol, cnt = back_project(d["coords"], d["origin"], d["vs"], d["feats"], d["KR"])
        vol.backward(d["go"])
for _ in range(20): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): step()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("host issue %.1f us/step, with drain %.1f us/step" % ((t1 - t0) / 200 * 1e6, (t2 - t0) / 200 * 1e6))
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)

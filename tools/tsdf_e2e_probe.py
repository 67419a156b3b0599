"""Why does the per-call TSDF path vary 10x inside bench.py?  Times the (c) loop of bench_tsdf fresh, after a torch CPU
parallel region, and after the OpenMP C oracle ran (spinning OpenMP workers compete with the library's copy threads)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deep3dmap_b200 import TSDFVolume, synth
print(torch.__config__.parallel_info().replace("\n", " | ")[:400])
print("OMP_WAIT_POLICY", os.environ.get("OMP_WAIT_POLICY"), "KMP_BLOCKTIME", os.environ.get("KMP_BLOCKTIME"))
F = 300
K = synth.tsdf_intrinsics()
depths = np.stack([synth.tsdf_depth(f) for f in range(F)])
poses = np.stack([synth.tsdf_pose(f) for f in range(F)])
vol = TSDFVolume(np.array([[0.0, 20.48]] * 3), 0.04, margin=3)
def loop(tag):
    vol.reset(); torch.cuda.synchronize()
    ts = []
    t0 = time.perf_counter()
    for f in range(F):
        t1 = time.perf_counter()
        vol.integrate(None, depths[f], K, poses[f], 1.0)
        ts.append(time.perf_counter() - t1)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ts = np.array(ts) * 1e6
    print("%-34s %8.0f frames/s   per call us: median %.0f  p90 %.0f  max %.0f  first10 %s" % (
        tag, F / dt, np.median(ts), np.percentile(ts, 90), ts.max(), np.round(ts[:10]).astype(int).tolist()))
loop("fresh (incl. lazy init)")
loop("fresh again")
x = torch.randn(2000, 2000); y = x @ x
loop("right after torch CPU matmul")
loop("again")
import oracle
lv = synth.fragment_level_inputs(0)
oracle.back_project_fwd(lv["coords"], lv["origin"], lv["voxel_size"], lv["feats"], lv["KRcam"]) if hasattr(oracle, "back_project_fwd") else None
loop("right after the OpenMP C oracle")
loop("again")

"""A/B of the TSDF kernel tuning switches (run on the GPU box): one process per variant."""
import os, subprocess, sys
code = r'''
import numpy as np, torch, sys, os
sys.path.insert(0, os.getcwd())
from deep3dmap_b200 import TSDFVolume, synth
F=300
K=synth.tsdf_intrinsics()
d=torch.from_numpy(np.stack([synth.tsdf_depth(f) for f in range(F)])).cuda()
P=np.stack([synth.tsdf_pose(f) for f in range(F)])
v=TSDFVolume(np.array([[0.0,20.48]]*3),0.04,margin=3)
v.integrate_batch(d,K,P); torch.cuda.synchronize()
ts=[]
for _ in range(5):
    v.reset(); torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record(); v.integrate_batch(d,K,P); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
v.reset(); torch.cuda.synchronize(); a.record()
for f in range(F): v.integrate_batch(d[f:f+1],K,P[f:f+1])
b.record(); torch.cuda.synchronize()
w=torch.as_tensor(v.device_volumes()[1],device="cuda")
print("variant", os.environ.get("D3M_TSDF_VARIANT"), "group", os.environ.get("D3M_TSDF_GROUP"), "batch300 ms", min(ts), "per-frame us", a.elapsed_time(b)/F*1e3, "checksum", float(w.double().sum()))
'''
for var in sys.argv[1:] or ["0", "1", "2", "3"]:      # "<variant>" or "<variant>:<frames per group>"
    v, _, g = var.partition(":")
    subprocess.run([sys.executable, "-c", code], env=dict(os.environ, D3M_TSDF_VARIANT=v, D3M_TSDF_GROUP=g or "0"))

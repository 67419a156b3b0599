"""Where does the per-frame host-API time of TSDFVolume.integrate go?  (run on the GPU box)"""
import ctypes, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep3dmap_b200 import TSDFVolume, synth, _lib

K = synth.tsdf_intrinsics()
F = 200
depths = [synth.tsdf_depth(f) for f in range(F)]
poses = [synth.tsdf_pose(f) for f in range(F)]
vol = TSDFVolume(np.array([[0.0, 20.48]] * 3), 0.04, margin=3)
for f in range(20):
    vol.integrate(None, depths[f], K, poses[f], 1.0)
torch.cuda.synchronize()

def timeit(fn, n=F):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for f in range(n): fn(f)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    return (t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6

print("integrate(): host issue %.1f us/frame, incl. drain %.1f us/frame" % timeit(lambda f: vol.integrate(None, depths[f], K, poses[f], 1.0)))
d_dev = torch.from_numpy(np.stack(depths)).cuda()
P = np.stack(poses)
print("integrate_batch(1 resident frame): %.1f / %.1f us" % timeit(lambda f: vol.integrate_batch(d_dev[f:f+1], K, P[f:f+1])))
pin = torch.empty((480, 640), dtype=torch.float32).pin_memory()
pin_np = pin.numpy()
print("np.copyto pageable->pinned 1.2MB: %.1f / %.1f us" % timeit(lambda f: np.copyto(pin_np, depths[f])))
dd = torch.empty((480, 640), dtype=torch.float32, device="cuda")
print("H2D from pinned 1.2MB (torch copy_ non_blocking): %.1f / %.1f us" % timeit(lambda f: dd.copy_(pin, non_blocking=True)))
pg = [torch.from_numpy(d) for d in depths]
print("H2D from pageable 1.2MB (torch copy_): %.1f / %.1f us" % timeit(lambda f: dd.copy_(pg[f])))
L = _lib.lib()
Kf = np.ascontiguousarray(K.reshape(-1).astype(np.float32)); 
def raw(f):
    T = np.ascontiguousarray(poses[f].reshape(-1).astype(np.float32))
    L.d3m_tsdf_integrate_host(vol._h.ptr, depths[f].ctypes.data_as(ctypes.c_void_p), None, 480, 640, Kf.ctypes.data_as(ctypes.c_void_p), T.ctypes.data_as(ctypes.c_void_p), 1.0, 0, None)
print("raw C call d3m_tsdf_integrate_host: %.1f / %.1f us" % timeit(raw))

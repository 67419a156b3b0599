"""Where the host microseconds of one eager back_project call go: the bare C calls (ctypes, same arguments as the Python
mirror builds) vs the mirror (tensor allocation, autograd).  Level 2 of the bench fragment; GPU kept busy-free (host cost only:
the loop runs ahead of the device, a sync at the end)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from deep3dmap_b200 import back_project, voxel, _lib

dev = torch.device("cuda:0")
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
def cnt_fn(inp):
    return back_project(t(inp["coords"]), t(inp["origin"]), inp["voxel_size"], t(inp["feats"]), t(inp["KRcam"]))[1].cpu().numpy()
levels = bench.build_fragment_levels(cnt_fn)
L = _lib.lib()
REPS = 300

def loop(fn):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(REPS): fn()
    dt = (time.perf_counter() - t0) / REPS * 1e6
    torch.cuda.synchronize()
    return dt

for li, inp in enumerate(levels):
    coords, origin, feats, KR, go = t(inp["coords"]), t(inp["origin"]), t(inp["feats"]), t(inp["KRcam"]), t(inp["grad_out"])
    vs = float(inp["voxel_size"])
    V, B, C, H, W = feats.shape
    N = coords.shape[0]
    kind = voxel._COORD_KIND[coords.dtype]
    out = torch.empty((N, C + 1), device=dev); count = torch.empty((N,), device=dev)
    scratch = torch.empty((V, B, H, W, C), device=dev)
    ws_bytes = (L.d3m_back_project_fwd_workspace(N, B, V, C) + 255) // 256 * 256
    hist_elems = L.d3m_back_project_cell_hist_elems(N, B, V, H, W)
    buf = torch.empty((ws_bytes + 4 * hist_elems,), dtype=torch.uint8, device=dev)
    bws_bytes = L.d3m_back_project_bwd_workspace(N, B, V, C, H, W)
    bws = torch.empty((bws_bytes,), dtype=torch.uint8, device=dev)
    grad = torch.empty((V, B, C, H, W), device=dev)
    stream = voxel._stream(dev)
    args_f = (coords.data_ptr(), kind, N, origin.data_ptr(), B, vs, feats.data_ptr(), _lib.FEATS_NCHW, scratch.data_ptr(), V, C,
              H, W, KR.data_ptr(), out.data_ptr(), count.data_ptr(), buf.data_ptr() + ws_bytes, buf.data_ptr(), ws_bytes, stream)
    args_b = (coords.data_ptr(), kind, N, origin.data_ptr(), B, vs, V, C, H, W, KR.data_ptr(), go.data_ptr(), count.data_ptr(),
              buf.data_ptr() + ws_bytes, grad.data_ptr(), 1, bws.data_ptr(), bws_bytes, stream)
    c_f = loop(lambda: L.d3m_back_project_fwd(*args_f))
    c_fb = loop(lambda: (L.d3m_back_project_fwd(*args_f), L.d3m_back_project_bwd(*args_b)))
    f_ng = feats
    with torch.no_grad():
        py_f_nograd = loop(lambda: back_project(coords, origin, vs, f_ng, KR))
    fr = feats.clone().requires_grad_(True)
    py_f = loop(lambda: back_project(coords, origin, vs, fr, KR))
    def fb():
        fr.grad = None
        vol, _ = back_project(coords, origin, vs, fr, KR)
        vol.backward(go)
    py_fb = loop(fb)
    e = loop(lambda: torch.empty((N, C + 1), device=dev))
    print("level %d  N=%6d | C calls: fwd %.1f us, fwd+bwd %.1f us | mirror: fwd(no grad) %.1f, fwd(autograd) %.1f, fwd+backward() %.1f | "
          "torch.empty %.1f us" % (li, N, c_f, c_fb, py_f_nograd, py_f, py_fb, e), flush=True)

import time, numpy as np, torch, sys, os
sys.path.insert(0, os.getcwd())
from deep3dmap_b200.voxel import upload
dev = torch.device("cuda:0")
for mb in (1.2, 11, 28, 369):
    n = int(mb * (1 << 20) / 4)
    a = torch.from_numpy(np.random.default_rng(0).standard_normal(n).astype(np.float32))
    for f, name in ((lambda: a.to(dev), "torch .to()"), (lambda: upload(a, dev), "d3m_upload")):
        f(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5): o = f()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
        ok = bool((o.cpu() == a).all())
        print("%7.1f MB  %-12s %8.3f ms  %6.1f GB/s  equal=%s" % (mb, name, dt * 1e3, n * 4 / dt / 1e9, ok))

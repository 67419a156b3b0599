import time, numpy as np, torch, sys, os
sys.path.insert(0, os.getcwd())
from deep3dmap_b200.voxel import upload
dev = torch.device("cuda:0")
a = torch.from_numpy(np.random.default_rng(0).standard_normal(7 << 20).astype(np.float32))   # 28 MB
x = torch.randn(512, 512)
for mode in ("busy-loop gap", "sleep gap", "torch cpu work gap"):
    for f, name in ((lambda: a.to(dev), "torch .to()"), (lambda: upload(a, dev), "d3m_upload")):
        f(); torch.cuda.synchronize()
        ts = []
        for _ in range(8):
            if mode == "sleep gap": time.sleep(0.003)
            elif mode == "busy-loop gap":
                t1 = time.perf_counter()
                while time.perf_counter() - t1 < 0.003: pass
            else:
                for _ in range(20): torch.inverse(torch.eye(4) * 2)
                y = x @ x
            t0 = time.perf_counter(); o = f(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
        print("%-20s %-12s ms: %s" % (mode, name, " ".join("%.2f" % t for t in ts)))
print("torch threads", torch.get_num_threads(), "cpus", os.cpu_count())

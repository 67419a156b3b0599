"""cProfile of one dataloader sample through SeqRandomTransformSpace (where do the host milliseconds go?)"""
import cProfile, pstats, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deep3dmap_b200.transforms import SeqRandomTransformSpace
V, H, W = 9, 480, 640
rng = np.random.default_rng(5)
full = [np.clip(rng.standard_normal((300 >> l, 260 >> l, 90 >> l)).astype(np.float32), -1, 1) for l in range(3)]
K = np.array([[577.87, 0, 319.5], [0, 577.87, 239.5], [0, 0, 1]], dtype=np.float32)
pose = np.eye(4, dtype=np.float32); pose[:3, 3] = [4.0, 3.0, 1.5]
depth = np.full((V, H, W), 2.0, dtype=np.float32)
def data():
    return {"vol_origin": np.array([0.0, 0.0, -0.2], dtype=np.float32), "epoch": [3],
            "tsdf_list_full": [torch.from_numpy(t) for t in full], "extrinsics": torch.from_numpy(np.stack([pose] * V)).clone(),
            "intrinsics": torch.from_numpy(np.stack([K] * V)), "imgs": torch.zeros((V, 3, H, W)), "depth": torch.from_numpy(depth)}
torch.set_num_threads(int(os.environ.get('TORCH_THREADS', '1')))   # DataLoader workers run with 1
tr = SeqRandomTransformSpace([96, 96, 96], 0.04, max_epoch=8)
for _ in range(3): tr(data())
ds = [data() for _ in range(10)]
torch.cuda.synchronize(); t0 = time.perf_counter()
for d in ds: tr(d)
torch.cuda.synchronize(); print("ms per sample", (time.perf_counter() - t0) / 10 * 1e3)
ds = [data() for _ in range(10)]
pr = cProfile.Profile(); pr.enable()
for d in ds: tr(d)
pr.disable()
pstats.Stats(pr).sort_stats("cumtime").print_stats(30)

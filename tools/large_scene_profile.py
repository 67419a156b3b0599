"""Per-kernel times of the large-scene leg (BASELINE configs[4]) on one GPU + cell-occupancy statistics."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
import bench
from deep3dmap_b200 import synth, voxel
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
print(json.dumps(bench.bench_large_scene(torch, dist, dev, flush, 0, 1)))
# occupancy of the bilinear cells
V, L = 64, synth.LEVELS[2]
coords = torch.from_numpy(synth.large_scene_coords(dtype=np.int32)).to(dev)
R, c = synth.large_scene_cameras(V)
KR = torch.from_numpy(synth.krcam_from(R, c, synth.scaled_K(L["scale"]))[:, None].copy()).to(dev)
feats = torch.zeros((V, 1, L["H"], L["W"], L["C"]), device=dev)
_, cnt, hist = voxel.back_project_forward(coords, torch.zeros((1, 3), device=dev), synth.VOXEL_SIZE, feats, KR, cell_hist=True)
h = hist.flatten().cpu().numpy()
print("cells", h.size, "entries", int(h.sum()), "max", int(h.max()), "p50/p90/p99/p99.9", [int(np.percentile(h, q)) for q in (50, 90, 99, 99.9)],
      "sum k^2", float((h.astype(np.float64) ** 2).sum()), "cells>256", int((h > 256).sum()), "cells>2048", int((h > 2048).sum()))

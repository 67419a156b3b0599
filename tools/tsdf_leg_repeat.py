import os, sys, time, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from deep3dmap_b200 import _lib, TSDFVolume
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for i in range(3):
    r = bench.bench_tsdf(torch, dev, _lib, TSDFVolume, 6451.8, flush, with_cpu=(i == 1))
    print(i, {k: round(r[k], 1) for k in ("frames_per_s", "frames_per_s_per_call_resident", "e2e_frames_per_s", "e2e_datagen_3level_fps")}, flush=True)

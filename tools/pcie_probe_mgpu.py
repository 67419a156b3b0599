"""Host <-> device copy bandwidth with N ranks copying at the same time (torchrun): what the box gives the e2e leg of
bench.py, independent of any kernel.  Every rank moves the e2e step's volumes (51 MB in, 47 MB out) between pinned host
memory and its GPU: H2D alone, D2H alone, both directions at once; per-rank and aggregate GB/s (max time over ranks).
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_probe_mgpu.py"""
import json
import os

import torch
import torch.distributed as dist

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
H2D, D2H = 51_404_868, 47_417_792
hin = torch.empty(H2D, dtype=torch.uint8).pin_memory()
hout = torch.empty(D2H, dtype=torch.uint8).pin_memory()
din = torch.empty(H2D, dtype=torch.uint8, device=dev)
dout = torch.empty(D2H, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(kind, reps=20):
    def once():
        if kind in ("h2d", "both"):
            with torch.cuda.stream(s1):
                din.copy_(hin, non_blocking=True)
        if kind in ("d2h", "both"):
            with torch.cuda.stream(s2):
                hout.copy_(dout, non_blocking=True)
    for _ in range(3):
        once()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    for _ in range(reps):
        once()
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


res = {"ranks": world}
for kind in ("h2d", "d2h", "both"):
    ms = run(kind)
    nbytes = (H2D if kind != "d2h" else 0) + (D2H if kind != "h2d" else 0)
    res[kind] = {"ms": round(ms, 4), "GBs_per_rank": round(nbytes / ms / 1e6, 2), "GBs_aggregate": round(world * nbytes / ms / 1e6, 2)}
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()

"""Only the large-scene leg of bench.py under torchrun (scaling checks without paying for the whole bench)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = bench.bench_large_scene(torch, dist, dev, flush, rank, world)
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()

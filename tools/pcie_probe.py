"""Device → host bandwidth of the box, one process per GPU, all links at once (the ceiling of a sharded export's
sink). Under torchrun every rank copies 4K rgb24 frames (24.9 MB) from its HBM into its own pinned host buffer;
the line printed by rank 0 gives each rank's rate and the aggregate (bytes of all ranks / slowest rank's time).
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe.py [frames]
Run alone it reports the single-link figure."""
import json
import os
import sys

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 80
f = 3840*2160*3
slots = 8
dev = torch.empty((slots, f), dtype=torch.uint8, device="cuda")
host = torch.empty((slots, f), dtype=torch.uint8).pin_memory()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()


def run() -> float:
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(frames):
        host[k % slots].copy_(dev[k % slots], non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b)


run()
ms = min(run() for _ in range(3))
rates = [None]*world
if world > 1:
    dist.all_gather_object(rates, frames*f/ms/1e6)
else:
    rates = [frames*f/ms/1e6]
if rank == 0:
    slowest = frames*f/min(rates)/1e6
    print(json.dumps(dict(ranks=world, frames_per_rank=frames, frame_bytes=f, per_rank_gbs=[round(r, 1) for r in rates],
                          aggregate_gbs=round(world*frames*f/slowest/1e6, 1),
                          frames_per_s_ceiling=round(world*frames/(slowest/1e3)))), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()

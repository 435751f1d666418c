"""Frame-sharded export across the GPUs of one box.

The reference has no multi-GPU path. Here an export shards over TIME: rank r owns the contiguous frame
range `shard_range(n_frames, r, world)`, computes the (cheap) audio track for the whole clip so every
recurrence has exactly the reference's state, shades only its own frames into HBM, and rank 0
reassembles the stream in time order for the sink. The only exchange step is that reassembly, so the
only collective is point-to-point sends of finished frames to rank 0 (NCCL over NVLink on GPUs; the same
code runs over gloo on CPU tensors for the host-logic tests)."""
from __future__ import annotations

import os
from typing import Callable, Iterator, Optional

import torch


def shard_range(n_frames: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced: the first `n_frames % world` ranks get one frame more"""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n_frames, world)
    start = rank*base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def owner_of(frame: int, n_frames: int, world: int) -> int:
    base, extra = divmod(n_frames, world)
    edge = extra*(base + 1)
    return frame//(base + 1) if frame < edge else extra + (frame - edge)//max(base, 1)


def env_rank_world() -> tuple[int, int, int]:
    """(rank, world, local_rank) from the torchrun environment, (0, 1, 0) when absent"""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def init_process_group(backend: Optional[str] = None) -> tuple[int, int]:
    """Joins the torchrun rendezvous (NCCL when CUDA is there, else gloo). Idempotent."""
    import torch.distributed as dist
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29531")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world


class FrameGather:
    """Moves every rank's finished frames to rank 0 in time order.

        root:    for frames in gather.stream(local_frames): sink(frames)   # yields [m, ...] uint8 blocks
        others:  gather.stream(local_frames) sends and yields nothing

    `local_frames` is this rank's [n_local, ...] uint8 tensor (frames of its shard, in order). Blocks
    of at most `chunk` frames travel per message so rank 0 needs only `chunk` frames of staging."""

    def __init__(self, n_frames: int, rank: int, world: int, chunk: int = 32, group=None):
        self.n_frames, self.rank, self.world, self.chunk, self.group = n_frames, rank, world, max(1, chunk), group

    def stream(self, local_frames: torch.Tensor) -> Iterator[torch.Tensor]:
        import torch.distributed as dist
        start, stop = shard_range(self.n_frames, self.rank, self.world)
        assert local_frames.shape[0] == stop - start, (local_frames.shape, start, stop)
        if self.world == 1:
            for a in range(0, stop - start, self.chunk):
                yield local_frames[a:a + self.chunk]
            return
        if self.rank != 0:
            for a in range(0, stop - start, self.chunk):
                dist.send(local_frames[a:a + self.chunk].contiguous(), dst=0, group=self.group)
            return
        for a in range(0, stop - start, self.chunk):
            yield local_frames[a:a + self.chunk]
        staging = torch.empty((self.chunk, *local_frames.shape[1:]), dtype=local_frames.dtype, device=local_frames.device)
        for peer in range(1, self.world):
            p0, p1 = shard_range(self.n_frames, peer, self.world)
            for a in range(0, p1 - p0, self.chunk):
                block = staging[:min(self.chunk, p1 - p0 - a)]
                dist.recv(block, src=peer, group=self.group)
                yield block


def max_over_ranks(value: float, device=None) -> float:
    """Timing rule: a multi-GPU duration is the max over ranks"""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

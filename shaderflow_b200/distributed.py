"""Frame-sharded export across the GPUs of one box.

The reference has no multi-GPU path. Here an export shards over TIME: every rank computes the (cheap) audio
track for the whole clip so every recurrence has exactly the reference's state, shades only the frames it owns,
and the stream is reassembled in time order. The only exchange step of the path is that reassembly:

  * frames bound for a SINK (ffmpeg, a file, the null sink) are reassembled in HOST memory: ownership is
    block-cyclic (`block_owner`), every rank copies its frames device → host over its own PCIe link into its ring
    of one shared-memory segment, rank 0's writer thread streams them in time order (`negotiate_shared_sink`,
    csrc/sink.cu). No collective touches the data path; barriers bracket the export.
  * frames that STAY IN HBM for a consumer on rank 0 (`on_frame`, bench.py's `value` leg) use contiguous ranges
    (`shard_range`) and are reassembled in rank 0's HBM: peer stores over NVLink from the shading kernel itself
    (`PeerFrames`, CUDA IPC), or point-to-point NCCL sends when IPC is refused (`FrameGather`; the same code
    runs over gloo on CPU tensors for the host-logic tests)."""
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


def block_owner(frame: int, block: int, world: int) -> int:
    """Block-cyclic ownership of the sink-bound sharded export (csrc/sink.cu): frame g → rank (g / block) % world"""
    return (frame//block) % world


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
            # the frame exchange overlaps with shading: keep NCCL's point-to-point kernels small (every
            # resident NCCL CTA displaces shading CTAs on its SM)
            os.environ.setdefault("NCCL_MAX_P2P_NCHANNELS", "8")
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

    def __init__(self, n_frames: int, rank: int, world: int, chunk: int = 32, group=None, ahead: int = 8):
        self.n_frames, self.rank, self.world, self.chunk, self.group = n_frames, rank, world, max(1, chunk), group
        self.ahead = max(1, ahead)
        self._works: list = []          # sender: (work, block) pairs in flight
        self._plan: list = []           # root: (peer, frames) of every remote block, in time order
        self._slots: list = []          # root: ring of receive buffers
        self._posted: list = []         # root: (work, slot, frames) of the irecvs in flight, in time order
        self._next = 0                  # root: next plan entry to post
        self._full = None               # root: {peer: [frames, ...]} when every remote frame is staged
        self._rounds: list = []         # root: works of each batched round of receives

    # -- overlapped form: the exchange runs while the ranks are still shading ----------------------------
    # sender:  send_block(frames) as soon as a block of its shard is complete, finish() at the end
    # root:    post(shape, dtype, device) before shading, then `for block in drain()` after its own frames
    def send_block(self, frames: torch.Tensor) -> None:
        """Non-blocking send of a finished block of this rank's shard (in order) to rank 0"""
        import torch.distributed as dist
        assert self.rank != 0
        for a in range(0, frames.shape[0], self.chunk):
            block = frames[a:a + self.chunk].contiguous()
            self._works.append((dist.isend(block, dst=0, group=self.group), block))

    def finish(self) -> None:
        for work, _ in self._works:
            work.wait()
        self._works.clear()

    def post(self, frame_shape, dtype, device, budget_bytes: Optional[int] = None) -> None:
        """Root: posts the receives before this rank starts shading, so remote frames arrive over NVLink
        while every rank is still busy.

        When all remote frames fit `budget_bytes` of staging (default 24 GiB, SFB_GATHER_BUDGET_GB) the
        receives are posted round by round — block b of EVERY peer in one batched group, the order in which
        the senders finish them — and only the consumption is in time order. Larger exports fall back to a
        ring of `ahead` blocks received strictly in time order (the later peers then wait for their turn)."""
        import torch.distributed as dist
        assert self.rank == 0
        if budget_bytes is None:
            budget_bytes = int(float(os.environ.get("SFB_GATHER_BUDGET_GB", "24"))*(1 << 30))
        self._plan, self._rounds, self._full = [], [], None
        shards = {peer: shard_range(self.n_frames, peer, self.world) for peer in range(1, self.world)}
        for peer, (p0, p1) in shards.items():
            self._plan += [(peer, min(self.chunk, p1 - p0 - a)) for a in range(0, p1 - p0, self.chunk)]
        frame_bytes = torch.empty((), dtype=dtype).element_size()
        for d in frame_shape:
            frame_bytes *= int(d)
        remote = sum(p1 - p0 for p0, p1 in shards.values())
        if remote*frame_bytes <= budget_bytes:
            self._full = {peer: torch.empty((p1 - p0, *frame_shape), dtype=dtype, device=device) for peer, (p0, p1) in shards.items()}
            most = max((p1 - p0 for p0, p1 in shards.values()), default=0)
            for a in range(0, most, self.chunk):
                ops = [dist.P2POp(dist.irecv, buf[a:a + self.chunk], peer, self.group)
                       for peer, buf in self._full.items() if a < buf.shape[0]]
                self._rounds.append(dist.batch_isend_irecv(ops) if ops else [])
            return
        n = min(self.ahead, len(self._plan))
        self._slots = [torch.empty((self.chunk, *frame_shape), dtype=dtype, device=device) for _ in range(n)]
        self._posted, self._next = [], 0
        for slot in range(n):
            self._post_next(slot)

    def _post_next(self, slot: int) -> None:
        import torch.distributed as dist
        if self._next >= len(self._plan):
            return
        peer, frames = self._plan[self._next]
        self._next += 1
        block = self._slots[slot][:frames]
        self._posted.append((dist.irecv(block, src=peer, group=self.group), slot, frames))

    def drain(self) -> Iterator[torch.Tensor]:
        """Root: yields every remote block in time order. In ring mode the block's memory is reused for a
        later receive once the consumer has enqueued its use of it and asks for the next block."""
        assert self.rank == 0
        if self._full is not None:
            waited = 0
            for peer, buf in self._full.items():
                for b, a in enumerate(range(0, buf.shape[0], self.chunk)):
                    while waited <= b:                       # round b carries block b of every peer
                        for work in self._rounds[waited]:
                            work.wait()
                        waited += 1
                    yield buf[a:a + self.chunk]
            self._full, self._rounds = None, []
            return
        while self._posted:
            work, slot, frames = self._posted.pop(0)
            work.wait()
            yield self._slots[slot][:frames]
            self._post_next(slot)
        self._slots = []

    # -- plain form: everything after the fact (kept for small jobs and the host tests) ------------------
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


class PeerFrames:
    """Rank 0's staging for every remote frame, mapped into the other ranks over CUDA IPC so that they
    shade STRAIGHT INTO rank 0's HBM: the frame kernel's rgb24 stores travel over NVLink as peer writes, the
    exchange costs no kernel, no copy and no SM on either side, and it is complete when the shading is. What
    is left of the collective is one barrier (NCCL) telling rank 0 that every shard has been written.

    `negotiate` is collective over the group; it returns None on every rank when the mapping is not possible
    (not CUDA/NCCL, staging over budget, IPC refused) and the export then uses FrameGather instead.
    frames[k] is global frame `first_remote + k`."""

    def __init__(self, frames: torch.Tensor, first_remote: int, keepalive=None):
        self.frames, self.first_remote, self._keepalive = frames, first_remote, keepalive

    def slot(self, frame: int) -> torch.Tensor:
        return self.frames[frame - self.first_remote]

    @staticmethod
    def negotiate(n_frames: int, frame_shape, rank: int, world: int, device: int, group=None,
                  budget_bytes: Optional[int] = None, enable_peer=None) -> Optional["PeerFrames"]:
        import torch.distributed as dist
        if world < 2 or not torch.cuda.is_available() or dist.get_backend(group) != "nccl" or os.environ.get("SFB_NO_PEER_FRAMES"):
            return None
        if budget_bytes is None:
            budget_bytes = int(float(os.environ.get("SFB_PEER_BUDGET_GB", "96"))*(1 << 30))
        _, first_remote = shard_range(n_frames, 0, world)
        remote = n_frames - first_remote
        frame_bytes = 1
        for d in frame_shape:
            frame_bytes *= int(d)
        payload, frames, ok = [None], None, 1
        if rank == 0:
            try:
                if remote*frame_bytes > budget_bytes:
                    raise MemoryError("remote frames exceed the staging budget")
                from torch.multiprocessing.reductions import reduce_tensor
                frames = torch.empty((remote, *frame_shape), dtype=torch.uint8, device=f"cuda:{device}")
                payload = [reduce_tensor(frames)]
            except Exception:
                frames, payload = None, [None]
        dist.broadcast_object_list(payload, src=0, group=group)
        if payload[0] is None:
            ok = 0
        elif rank != 0:
            try:
                import inspect
                rebuild, args = payload[0]
                args = list(args)
                where = list(inspect.signature(rebuild).parameters).index("storage_device")
                owner = args[where]                        # device index of the exporting rank
                if owner != device and enable_peer is not None:
                    enable_peer(owner)                     # kernels on `device` may store into `owner`'s memory
                if not os.environ.get("SFB_PEER_OPEN_ON_OWNER"):
                    # map the allocation into THIS rank's device address space (cudaIpcOpenMemHandle maps into
                    # the current device and enables peer access lazily); the tensor only serves as a pointer
                    args[where] = device
                frames = rebuild(*args)                    # a view of rank 0's memory in this process
            except Exception:
                frames, ok = None, 0
        flag = torch.tensor([ok], dtype=torch.int32, device=f"cuda:{device}")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            return None
        return PeerFrames(frames, first_remote)


def negotiate_shared_sink(ctx, frame_bytes: int, rank: int, world: int, slots: int, device=None, group=None):
    """Collective: rank 0 creates the shared-memory segment of the sharded export's host rings, every other rank
    attaches to it (csrc/sink.cu). Returns this rank's `SharedSink`, or None on EVERY rank when any rank failed
    (no memfd/shm, pinning refused) — the export then falls back to staging in rank 0's HBM.
    ctx None = host-only sinks (CPU tests of the protocol)."""
    import torch.distributed as dist
    from shaderflow_b200 import _native as N
    sink, payload = None, [None]
    if rank == 0:
        try:
            sink = N.SharedSink(ctx, None, 0, world, slots, frame_bytes)
            payload = [sink.path]
        except Exception as error:                              # noqa: BLE001 — reported below, collectively
            payload = [None]
            _last_error[0] = str(error)
    dist.broadcast_object_list(payload, src=0, group=group)
    ok = 1
    if payload[0] is None:
        ok = 0
    elif rank != 0:
        try:
            sink = N.SharedSink(ctx, payload[0], rank, world, slots, frame_bytes)
        except Exception as error:                              # noqa: BLE001
            ok, _last_error[0] = 0, str(error)
    on = device if device is not None else ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    flag = torch.tensor([ok], dtype=torch.int32, device=on)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if int(flag.item()) == 0:
        if sink is not None:
            sink.close()
        return None
    return sink


_last_error = [None]


def max_over_ranks(value: float, device=None) -> float:
    """Timing rule: a multi-GPU duration is the max over ranks"""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

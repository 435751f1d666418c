"""Frame sharding and the reassembly exchange on CPU: world_size-2 (and 3) gloo process groups."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from shaderflow_b200.distributed import FrameGather, max_over_ranks, owner_of, shard_range


def test_shard_ranges_partition_the_export():
    for n in (0, 1, 7, 60, 3600, 3601):
        for world in (1, 2, 3, 8):
            ranges = [shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
            for f in range(0, n, max(1, n//17)):
                r = owner_of(f, n, world)
                assert ranges[r][0] <= f < ranges[r][1]
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank: int, world: int, port: int, n_frames: int, chunk: int, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = shard_range(n_frames, rank, world)
    # frame k is filled with (k mod 251) and stamped with its index
    local = torch.stack([torch.full((4, 6, 3), k % 251, dtype=torch.uint8) for k in range(a, b)]) if b > a \
        else torch.empty((0, 4, 6, 3), dtype=torch.uint8)
    got = [blk.clone() for blk in FrameGather(n_frames, rank, world, chunk=chunk).stream(local)]
    slowest = max_over_ranks(float(rank + 1))
    if rank == 0:
        frames = torch.cat(got) if got else torch.empty((0, 4, 6, 3), dtype=torch.uint8)
        out.put((frames.numpy(), slowest))
    else:
        assert got == [] and slowest == float(world)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_frames,chunk", [(2, 21, 4), (2, 8, 16), (3, 10, 3)])
def test_gather_reassembles_frames_in_time_order(world, n_frames, chunk):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, chunk, out)) for r in range(world)]
    for p in procs: p.start()
    frames, slowest = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert frames.shape == (n_frames, 4, 6, 3) and slowest == float(world)
    assert np.array_equal(frames[:, 0, 0, 0], np.arange(n_frames) % 251)


def _overlap_worker(rank: int, world: int, port: int, n_frames: int, chunk: int, ahead: int, budget, out):
    """The overlapped form scene.main uses: receives posted up front, blocks sent as they are finished"""
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = shard_range(n_frames, rank, world)
    gather = FrameGather(n_frames, rank, world, chunk=chunk, ahead=ahead)
    if rank == 0:
        gather.post((4, 6, 3), torch.uint8, "cpu", budget_bytes=budget)
        got = [torch.full((1, 4, 6, 3), k % 251, dtype=torch.uint8) for k in range(a, b)]      # its own frames
        got += [blk.clone() for blk in gather.drain()]
        out.put(torch.cat(got).numpy() if got else np.empty((0, 4, 6, 3), np.uint8))
    else:
        local = torch.stack([torch.full((4, 6, 3), k % 251, dtype=torch.uint8) for k in range(a, b)]) if b > a \
            else torch.empty((0, 4, 6, 3), dtype=torch.uint8)
        sent = 0
        for done in range(1, b - a + 1):                       # "shade" one frame, send every finished block
            if done - sent >= chunk or done == b - a:
                gather.send_block(local[sent:done]); sent = done
        gather.finish()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("budget", [None, 0], ids=["staged-rounds", "ring"])
@pytest.mark.parametrize("world,n_frames,chunk,ahead", [(2, 21, 4, 2), (3, 50, 4, 3), (2, 5, 16, 8), (3, 2, 4, 1)])
def test_overlapped_gather_keeps_time_order(world, n_frames, chunk, ahead, budget):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_overlap_worker, args=(r, world, port, n_frames, chunk, ahead, budget, out)) for r in range(world)]
    for p in procs: p.start()
    frames = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert frames.shape == (n_frames, 4, 6, 3)
    assert np.array_equal(frames[:, 0, 0, 0], np.arange(n_frames) % 251)


# ---------------------------------------------------------------------------------------------- #
# The sink-bound sharded export: block-cyclic ownership, one shared host ring per rank, rank 0's writer streams
# the frames in time order (csrc/sink.cu). Host-only sinks (ctx None) run the same protocol without a GPU.

def test_block_owner_matches_the_sink():
    from shaderflow_b200.distributed import block_owner
    assert [block_owner(g, 4, 3) for g in range(14)] == [0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 0, 0]
    assert [block_owner(g, 1, 2) for g in range(5)] == [0, 1, 0, 1, 0]


def _sink_worker(rank: int, world: int, port: int, n_frames: int, block: int, slots: int, frame_bytes: int, path, out):
    import torch.distributed as dist
    from shaderflow_b200.distributed import block_owner, negotiate_shared_sink
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sink = negotiate_shared_sink(None, frame_bytes, rank, world, slots)
    assert sink is not None
    for export in range(2):                                  # the ring serves consecutive exports
        fd = os.open(path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC) if rank == 0 else -1
        sink.begin(n_frames, block, fd)
        dist.barrier()
        rng = np.random.default_rng(rank)
        for g in range(n_frames):
            if block_owner(g, block, world) != rank:
                continue
            frame = np.full(frame_bytes, (g + export) % 251, dtype=np.uint8)
            frame[:8] = np.frombuffer(np.int64(g).tobytes(), np.uint8)
            if rng.random() < 0.3:
                import time; time.sleep(0.002)               # ranks drift apart: the writer must wait, owners must block
            sink.acquire()
            sink.submit_host(frame)
        frames, nbytes = sink.finish()
        dist.barrier()
        if rank == 0:
            os.close(fd)
            assert (frames, nbytes) == (n_frames, n_frames*frame_bytes)
    sink.close()
    if rank == 0:
        out.put("done")
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_frames,block,slots", [(2, 37, 4, 8), (3, 50, 3, 6), (2, 5, 1, 2), (3, 2, 4, 8)])
def test_shared_sink_streams_frames_in_time_order(tmp_path, world, n_frames, block, slots):
    frame_bytes = 5000                                       # not a multiple of the page size on purpose
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    path = str(tmp_path/"stream.rgb")
    procs = [ctx.Process(target=_sink_worker, args=(r, world, port, n_frames, block, slots, frame_bytes, path, out)) for r in range(world)]
    for p in procs: p.start()
    assert out.get(timeout=120) == "done"
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    data = np.fromfile(path, dtype=np.uint8).reshape(n_frames, frame_bytes)
    assert np.array_equal(data[:, :8].copy().view(np.int64)[:, 0], np.arange(n_frames))
    assert np.array_equal(data[:, 100], (np.arange(n_frames) + 1) % 251)      # the second export's payload


def test_shared_sink_reports_a_dead_sink_to_every_owner(tmp_path):
    """A sink whose descriptor fails (ffmpeg died): the writer flags it and acquire() raises instead of blocking"""
    from shaderflow_b200 import _native as N
    sink = N.SharedSink(None, None, 0, 1, 4, 4096)
    r, w = os.pipe()
    os.close(r)                                              # writes to `w` now fail with EPIPE
    import signal
    old = signal.signal(signal.SIGPIPE, signal.SIG_IGN)
    try:
        sink.begin(16, 2, w)
        with pytest.raises(RuntimeError, match="sink write failed"):
            for _ in range(16):
                sink.acquire()
                sink.submit_host(np.zeros(4096, np.uint8))
            sink.finish()
        sink.abort()
    finally:
        signal.signal(signal.SIGPIPE, old)
        os.close(w)
        sink.close()

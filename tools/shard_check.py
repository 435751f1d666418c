"""Under torchrun (N >= 2): a frame-sharded export must produce exactly the bytes of a single-GPU export, whichever
way the frames are reassembled:
  sink   output=bytes → block-cyclic ownership, every rank drains into the shared host ring (csrc/sink.cu)
  hbm    output=None, on_frame → contiguous ranges reassembled in rank 0's HBM (peer stores, or NCCL with SFB_NO_PEER_FRAMES=1)
usage: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/shard_check.py"""
import ctypes
import sys
import zlib
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from shaderflow_b200 import distributed as D, synthetic
from examples.demo import Visualizer, synthetic_background

rank, world, local = D.env_rank_world()
torch.cuda.set_device(local)
D.init_process_group("nccl")
Visualizer.background = synthetic_background(960, 540)
scene = Visualizer(device=local)
scene.initialize()
frames = 37
seconds = frames/60
scene.audio.load(synthetic.chirp(seconds), 44100)
W, H = 1280, 720
flags = dict(width=W, height=H, ssaa=2, subsample=2, fps=60.0, time=seconds)
frame_bytes = W*H*3
cudart = ctypes.CDLL("libcudart.so.12")


def crc_of_device(pointer: int) -> int:
    host = np.empty(frame_bytes, np.uint8)
    scene.cuda.sync()
    assert cudart.cudaMemcpy(ctypes.c_void_p(host.ctypes.data), ctypes.c_void_p(pointer), ctypes.c_size_t(frame_bytes), 2) == 0
    return zlib.crc32(host.tobytes())


ok = True
for run in range(2):                                   # twice: the rings / staging serve consecutive exports
    sharded = scene.main(output=bytes, **flags)
    seen = {}
    scene.main(output=None, on_frame=lambda index, pointer: seen.__setitem__(index, crc_of_device(pointer)), **flags)
    if rank == 0:
        single = scene.main(output=bytes, distributed=False, **flags)
        want = [zlib.crc32(single[k*frame_bytes:(k + 1)*frame_bytes]) for k in range(frames)]
        sink = [zlib.crc32(sharded[k*frame_bytes:(k + 1)*frame_bytes]) for k in range(len(sharded)//frame_bytes)]
        hbm = [seen.get(k) for k in range(frames)]
        good = (sink == want) and (hbm == want) and len(sharded) == len(single)
        ok = ok and good
        print(f"world {world} run {run}: sink {sum(a == b for a, b in zip(sink, want))}/{frames} frames identical, "
              f"hbm {sum(a == b for a, b in zip(hbm, want))}/{frames} -> {'OK' if good else 'MISMATCH'}", flush=True)

# A scene with GPU feedback (MotionBlur averages its own last frames: temporal texture) must NOT shard: every rank calls
# main() under torchrun, rank 0 exports alone, and the bytes are those of a plain single-process export.
from examples.demo import MotionBlur
MotionBlur.background = synthetic_background(320, 180)
blur = MotionBlur(device=local)
blur_flags = dict(width=320, height=180, ssaa=1, subsample=1, fps=60.0, time=24/60)
under_torchrun = blur.main(output=bytes, **blur_flags)
if rank == 0:
    alone = blur.main(output=bytes, distributed=False, **blur_flags)
    good = under_torchrun == alone and len(alone) == 320*180*3*24
    ok = ok and good
    print(f"world {world}: feedback scene (MotionBlur) exported by rank 0 alone -> {'OK' if good else 'MISMATCH'}", flush=True)
else:
    ok = ok and under_torchrun is None
torch.distributed.barrier()
torch.distributed.destroy_process_group()
sys.exit(0 if ok else 1)

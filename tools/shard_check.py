"""Under torchrun (N >= 2): a frame-sharded export must produce exactly the bytes of a single-GPU export.
usage: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/shard_check.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from shaderflow_b200 import distributed as D, synthetic
from examples.demo import Visualizer, synthetic_background

rank, world, local = D.env_rank_world()
torch.cuda.set_device(local)
D.init_process_group("nccl")
Visualizer.background = synthetic_background(960, 540)
scene = Visualizer(device=local)
scene.initialize()
seconds = 37/60
scene.audio.load(synthetic.chirp(seconds), 44100)
flags = dict(width=1280, height=720, ssaa=2, subsample=2, fps=60.0, time=seconds)
sharded = scene.main(output=bytes, **flags)
ok = True
if rank == 0:
    single = scene.main(output=bytes, distributed=False, **flags)
    frame = 1280*720*3
    same = [sharded[k*frame:(k + 1)*frame] == single[k*frame:(k + 1)*frame] for k in range(len(single)//frame)]
    ok = len(sharded) == len(single) and all(same)
    print(f"world {world}: sharded {len(sharded)} bytes, single {len(single)} bytes, identical frames {sum(same)}/{len(same)} -> {'OK' if ok else 'MISMATCH'}")
torch.distributed.barrier()
torch.distributed.destroy_process_group()
sys.exit(0 if ok else 1)

import sys, time, cProfile, pstats
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from shaderflow_b200 import _native as N, synthetic
from examples.demo import Visualizer
W, H = 3840, 2160
Visualizer.background = synthetic.background(1920, 1080)
scene = Visualizer(device=0); scene.initialize()
clip = synthetic.chirp(1.0)
pinned = torch.from_numpy(clip).pin_memory()
scene.audio.load(clip, 44100)
flags = dict(width=W, height=H, ssaa=2, subsample=2, fps=60.0, time=1.0)
def value(): scene.main(output=None, **flags)
def e2e():
    scene.audio.load(pinned.numpy(), 44100)
    scene.main(output="null", buffers=4, **flags)
def timed(fn):
    torch.cuda.synchronize(); t = time.perf_counter(); fn(); torch.cuda.synchronize(); return (time.perf_counter()-t)*1e3
for _ in range(3): timed(value)
print("value", [round(timed(value), 1) for _ in range(3)])
print("e2e  ", [round(timed(e2e), 1) for _ in range(4)])
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable(); e2e(); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(12)

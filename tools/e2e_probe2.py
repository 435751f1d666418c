import sys, time, gc, subprocess
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
def e2e():
    scene.audio.load(pinned.numpy(), 44100)
    scene.main(output="null", buffers=4, **flags)
def timed(fn):
    torch.cuda.synchronize(); t = time.perf_counter(); fn(); torch.cuda.synchronize(); return round((time.perf_counter()-t)*1e3, 1)
for _ in range(3): timed(e2e)
print("gc on, no smi :", [timed(e2e) for _ in range(12)])
gc.collect(); gc.disable()
print("gc off, no smi:", [timed(e2e) for _ in range(12)])
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader", "-lms", "200"], stdout=subprocess.DEVNULL)
time.sleep(0.5)
print("gc off, smi   :", [timed(e2e) for _ in range(12)])
p.terminate()

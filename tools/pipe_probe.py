"""Where does the e2e time go? Times the sink ring alone and an export's phases."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from shaderflow_b200 import _native as N, synthetic
ctx = N.Context(0)
W, H = 3840, 2160
frame = torch.zeros((H, W, 3), dtype=torch.uint8, device="cuda")
host = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory()
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(20): host.copy_(frame, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t
print(f"plain D2H pinned: {20*frame.numel()/dt/1e9:.1f} GB/s ({dt/20*1e3:.2f} ms/frame)")
for buffers in (2, 4, 8):
    t = time.perf_counter(); p = N.Pipe(ctx, -1, buffers, frame.numel()); t_open = time.perf_counter() - t
    t = time.perf_counter()
    for _ in range(60):
        ptr = p.acquire(); p.submit(None)
    p.sync(); t_run = time.perf_counter() - t
    t = time.perf_counter(); p.close(); t_close = time.perf_counter() - t
    print(f"pipe buffers={buffers}: open {t_open*1e3:.1f} ms, 60 frames {t_run*1e3:.1f} ms ({60*frame.numel()/t_run/1e9:.1f} GB/s), close {t_close*1e3:.1f} ms")
from examples.demo import Visualizer
Visualizer.background = synthetic.background(1920, 1080)
scene = Visualizer(device=0); scene.initialize()
clip = synthetic.chirp(1.0); scene.audio.load(clip, 44100)
for output in (None, "null", None, "null"):
    torch.cuda.synchronize(); t = time.perf_counter()
    scene.main(width=W, height=H, ssaa=2, subsample=2, time=1.0, output=output, buffers=4)
    torch.cuda.synchronize(); print(f"main(output={output}): {(time.perf_counter()-t)*1e3:.1f} ms")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
scene.main(width=W, height=H, ssaa=2, subsample=2, time=1.0, output="null", buffers=4)
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(18)

"""Where the host time of one exported frame goes (run on the GPU box):
    python tools/host_probe.py [width height ssaa]
Times scene.main at the given geometry with the kernels launched asynchronously, then profiles the same loop."""
import cProfile
import io
import pstats
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from shaderflow_b200 import synthetic
from examples.demo import Visualizer, synthetic_background

W, H, S = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080, 1)
frames = 240
Visualizer.background = synthetic_background(1920, 1080)
scene = Visualizer(device=0)
scene.initialize()
scene.audio.load(synthetic.noise(frames/60.0), 44100)
flags = dict(width=W, height=H, ssaa=S, subsample=2, fps=60.0, time=frames/60.0)
for output in (None, "null"):
    for _ in range(2):
        scene.main(output=output, **flags)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    scene.main(output=output, **flags)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{W}x{H} ssaa {S} output={output}: host loop {1e6*(t1 - t0)/frames:.1f} us/frame, "
          f"with GPU drain {1e6*(t2 - t0)/frames:.1f} us/frame = {frames/(t2 - t0):.0f} fps")
prof = cProfile.Profile()
prof.enable()
scene.main(output=None, **flags)
prof.disable()
torch.cuda.synchronize()
buf = io.StringIO()
pstats.Stats(prof, stream=buf).sort_stats("cumulative").print_stats(28)
print(buf.getvalue()[:6000])

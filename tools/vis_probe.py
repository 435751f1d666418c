"""Device time of the visualizer kernels (development aid): the fused 4K 2xSSAA frame (BASELINE configs[2]) and the
1080p ssaa-1 iScreen pass of the reference's default export (configs[1]).
usage: python tools/vis_probe.py [flags] [volume]      (SFB_ROWS_DEBUG=4 selects the per-tap phase 2 of the rows kernel)"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from shaderflow_b200 import _native as N, synthetic
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
ctx = N.Context(0)
W, H = 3840, 2160
tb = N.Texture(ctx, 1920, 1080, 3, N.DTYPE_U8); tb.write(np.flipud(synthetic.background()).copy())
rng = np.random.default_rng(0)
ts = N.Texture(ctx, 1, 115, 2, N.DTYPE_F32, linear=False, repeat_x=True, repeat_y=False); ts.write((rng.random((115, 1, 2))*500).astype(np.float32))
tw = N.Texture(ctx, 180, 1, 2, N.DTYPE_F32, linear=True, repeat_x=False, repeat_y=False); tw.write(rng.random((1, 180, 2)).astype(np.float32))
volume = float(sys.argv[2]) if len(sys.argv) > 2 else 0.8
u = N.Uniforms.defaults(W, H); u.iTime = 1.0; u.iSSAA = 2; u.extra[0][0] = volume; u.extra[1][0] = 0.2
out = torch.zeros((H, W, 3), dtype=torch.uint8, device="cuda")
sid = N.scene_lookup("visualizer")
def run(): ctx.render_frame(sid, u, [tb, ts, tw], W, H, 2, 2, 3, out, flags)
for _ in range(3): run()
torch.cuda.synchronize()
ts_ = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(); b.record(); torch.cuda.synchronize(); ts_.append(a.elapsed_time(b))
import os
print(f"4K ssaa2 flags={flags} volume={volume} debug={os.environ.get('SFB_ROWS_DEBUG', '0')}: best {min(ts_):.3f} ms avg {sum(ts_)/len(ts_):.3f} ms")
u1 = N.Uniforms.defaults(1920, 1080); u1.iTime = 1.0; u1.iSSAA = 1; u1.extra[0][0] = volume; u1.extra[1][0] = 0.2
screen = torch.zeros((1080, 1920, 4), dtype=torch.uint8, device="cuda")
def run1(): ctx.render_screen(sid, u1, [tb, ts, tw], 1920, 1080, screen, None, flags)
for _ in range(3): run1()
torch.cuda.synchronize()
ts_ = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run1(); b.record(); torch.cuda.synchronize(); ts_.append(a.elapsed_time(b))
print(f"1080p ssaa1 screen pass: best {min(ts_)*1e3:.1f} us avg {sum(ts_)/len(ts_)*1e3:.1f} us")
final = torch.zeros((1080, 1920, 3), dtype=torch.uint8, device="cuda")
def run2(): ctx.render_final(screen, 1920, 1080, 1920, 1080, 2, 3, final)
for _ in range(3): run2()
torch.cuda.synchronize()
ts_ = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): run2()
    b.record(); torch.cuda.synchronize(); ts_.append(a.elapsed_time(b)/10)
print(f"1080p final pass (subsample 2): best {min(ts_)*1e3:.1f} us")

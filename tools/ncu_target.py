"""Tiny single-GPU workloads for ncu captures: a few launches of the kernels the bench times.
usage: python tools/ncu_target.py [visualizer|c2|tiled1|stft|mandelbrot|tetration|raymarch|literal-<fractal>|piano|all] [n]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from shaderflow_b200 import _native as N, synthetic

what = sys.argv[1] if len(sys.argv) > 1 else "all"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = N.Context(0)
def dev(a): return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def visualizer_inputs(W, H, ssaa):
    tb = N.Texture(ctx, 1920, 1080, 3, N.DTYPE_U8); tb.write(np.flipud(synthetic.background()).copy())
    rng = np.random.default_rng(0)
    ts = N.Texture(ctx, 1, 115, 2, N.DTYPE_F32, linear=False, repeat_x=True, repeat_y=False); ts.write((rng.random((115, 1, 2))*500).astype(np.float32))
    tw = N.Texture(ctx, 180, 1, 2, N.DTYPE_F32, linear=True, repeat_x=False, repeat_y=False); tw.write(rng.random((1, 180, 2)).astype(np.float32))
    u = N.Uniforms.defaults(W, H); u.iTime = 1.0; u.iSSAA = ssaa; u.extra[0][0] = 0.8; u.extra[1][0] = 0.2
    return u, [tb, ts, tw]


if what in ("visualizer", "all"):
    W, H = 3840, 2160
    u, tex = visualizer_inputs(W, H, 2)
    out = torch.zeros((H, W, 3), dtype=torch.uint8, device="cuda")
    for _ in range(n):
        ctx.render_frame(N.scene_lookup("visualizer"), u, tex, W, H, 2, 2, 3, out)
    ctx.sync()
if what in ("c2", "tiled1", "all"):
    # BASELINE configs[1]: the reference's default export (ssaa 1, subsample 2): iScreen pass + final pass
    W, H = 1920, 1080
    u, tex = visualizer_inputs(W, H, 1)
    screen = N.Texture(ctx, W, H, 4, N.DTYPE_U8, linear=True, repeat_x=False, repeat_y=False)
    out = torch.zeros((H, W, 3), dtype=torch.uint8, device="cuda")
    pointer, _ = screen.storage()
    for _ in range(n):
        ctx.render_target(N.scene_lookup("visualizer"), u, tex, screen, N.RENDER_TILED if what == "tiled1" else 0)
        ctx.render_final(pointer, W, H, W, H, 2, 3, out)
    ctx.sync()
if what in ("stft", "all"):
    from shaderflow_b200.audio.spectrogram import BrokenSpectrogram
    from shaderflow_b200.audio.module import BrokenAudio
    sp = BrokenSpectrogram(audio=BrokenAudio()); sp.from_notes(15, 129, piano=True)
    seconds = 3600                                              # bench.py's STFT leg: 216 000 frames, 1.27 GB of PCM
    frames = seconds*60
    pcm = torch.rand((2, seconds*44100), device="cuda")*2 - 1
    _, dt, tell = N.frame_clock(frames, 60.0, 1.0, 44100, 2, seconds*44100)
    spec = torch.zeros((frames, 115, 2), device="cuda")
    for _ in range(n):
        ctx.stft_mel(pcm, dev(tell), 12, sp.device_bank("cuda:0"), spec_out=spec)
    ctx.sync()
for name in ("mandelbrot", "tetration", "raymarch"):
    if what in (name, "literal-" + name, "fractals", "all"):
        W, H = 7680, 4320                                       # BASELINE configs[3]: 4x SSAA = 530.8 M fragments
        u = N.Uniforms.defaults(W, H); u.iSSAA = 4.0
        out = torch.zeros((H, W, 3), dtype=torch.uint8, device="cuda")
        for _ in range(n):
            ctx.render_frame(N.scene_lookup(name), u, [], W, H, 4, 4, 3, out, N.RENDER_LITERAL if what.startswith("literal") else 0)
        ctx.sync()
print("done", ctx.launches)

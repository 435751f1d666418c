"""Tiny single-GPU workload for ncu captures: a few launches of the kernels the bench times.
usage: python tools/ncu_target.py [visualizer|stft|all] [n]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from shaderflow_b200 import _native as N, synthetic

what = sys.argv[1] if len(sys.argv) > 1 else "all"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = N.Context(0)
def dev(a): return torch.from_numpy(np.ascontiguousarray(a)).cuda()

if what in ("visualizer", "all"):
    W, H = 3840, 2160
    tb = N.Texture(ctx, 1920, 1080, 3, N.DTYPE_U8); tb.write(np.flipud(synthetic.background()).copy())
    rng = np.random.default_rng(0)
    ts = N.Texture(ctx, 1, 115, 2, N.DTYPE_F32, linear=False, repeat_x=True, repeat_y=False); ts.write((rng.random((115, 1, 2))*500).astype(np.float32))
    tw = N.Texture(ctx, 180, 1, 2, N.DTYPE_F32, linear=True, repeat_x=False, repeat_y=False); tw.write(rng.random((1, 180, 2)).astype(np.float32))
    u = N.Uniforms.defaults(W, H); u.iTime = 1.0; u.iSSAA = 2; u.extra[0][0] = 0.8; u.extra[1][0] = 0.2
    out = torch.zeros((H, W, 3), dtype=torch.uint8, device="cuda")
    for _ in range(n):
        ctx.render_frame(N.scene_lookup("visualizer"), u, [tb, ts, tw], W, H, 2, 2, 3, out)
    ctx.sync()
if what in ("stft", "all"):
    import scipy.sparse
    from shaderflow_b200.audio.spectrogram import BrokenSpectrogram
    from shaderflow_b200.audio.module import BrokenAudio
    sp = BrokenSpectrogram(audio=BrokenAudio()); sp.from_notes(15, 129, piano=True)
    seconds = 600
    frames = seconds*60
    pcm = torch.rand((2, seconds*44100), device="cuda")*2 - 1
    _, dt, tell = N.frame_clock(frames, 60.0, 1.0, 44100, 2, seconds*44100)
    spec = torch.zeros((frames, 115, 2), device="cuda")
    for _ in range(n):
        ctx.stft_mel(pcm, dev(tell), 12, sp.device_bank("cuda:0"), spec_out=spec)
    ctx.sync()
print("done", ctx.launches)

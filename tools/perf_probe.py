"""Quick device-time probe of the kernels (development aid; bench.py is the contract)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from oracle import audio_np as A, glsl_np as G
from shaderflow_b200 import _native as N

ctx = N.Context(0)

def timed(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sum(ts)/len(ts)

def dev(a): return torch.from_numpy(np.ascontiguousarray(a)).cuda()

cfg = A.TrackConfig(bank=A.BankConfig.from_notes(15, 129, piano=True))
indptr, idx, val = A.filterbank_csr(A.filterbank_matrix(cfg.bank))
csr = (dev(indptr), dev(idx), dev(val), 115)
for seconds in (60, 3600):
    frames = seconds*60
    n = seconds*44100
    pcm = torch.rand((2, n), device="cuda")*2 - 1
    _, dt, tell = N.frame_clock(frames, 60.0, 1.0, 44100, 2, n)
    tell_d, dt_d = dev(tell), dev(dt)
    spec = torch.zeros((frames, 115, 2), device="cuda")
    best, avg = timed(lambda: ctx.stft_mel(pcm, tell_d, 12, csr, spec_out=spec))
    print(f"stft_mel {seconds}s clip: {frames} frames best {best:.3f} ms avg {avg:.3f} ms -> {frames*735*2/best/1e3:.1f} Msamples/s, "
          f"algorithmic {(frames*6800)/best/1e6:.1f} GB/s")
    scal = torch.zeros((frames, 5), dtype=torch.float64, device="cuda")
    wave = torch.zeros((frames, 180, 2), device="cuda")
    if seconds == 60:
        best, avg = timed(lambda: ctx.audio_track(pcm, 44100, tell_d, dt_d, spec=spec, bins=115), reps=3, warm=1)
        print(f"  spec scan: {best:.3f} ms")
        best, avg = timed(lambda: ctx.audio_track(pcm, 44100, tell_d, dt_d, scalars=scal), reps=3, warm=1)
        print(f"  scalars: {best:.3f} ms")
        best, avg = timed(lambda: ctx.audio_track(pcm, 44100, tell_d, dt_d, wave=wave), reps=3, warm=1)
        print(f"  waveform: {best:.3f} ms")
    del pcm

# render
bg = G.synthetic_background(1920, 1080)
tb = N.Texture(ctx, 1920, 1080, 3, N.DTYPE_U8); tb.write(np.flipud(bg).copy())
ts = N.Texture(ctx, 1, 115, 2, N.DTYPE_F32, linear=False, repeat_x=True, repeat_y=False); ts.write(np.random.rand(115, 1, 2).astype(np.float32)*500)
tw = N.Texture(ctx, 180, 1, 2, N.DTYPE_F32, linear=True, repeat_x=False, repeat_y=False); tw.write(np.random.rand(1, 180, 2).astype(np.float32))
def run(scene, W, H, ssaa, sub, flags=0, vol=0.8):
    u = N.Uniforms.defaults(W, H); u.iTime = 1.0; u.iSSAA = ssaa
    u.extra[0][0] = vol; u.extra[1][0] = 0.2
    sid = N.scene_lookup(scene)
    tex = [tb, ts, tw] if scene == "visualizer" else ([ts] if scene == "bars" else ([tw] if scene == "waveform" else []))
    out = torch.zeros((H, W, 3), dtype=torch.uint8, device="cuda")
    best, avg = timed(lambda: ctx.render_frame(sid, u, tex, W, H, ssaa, sub, 3, out), reps=3, warm=1)
    print(f"{scene:11s} {W}x{H} ssaa{ssaa} flags{flags}: best {best:.3f} ms avg {avg:.3f} -> {1e3/best:.1f} fps; {W*H*ssaa*ssaa/best/1e6:.2f} Gsamples/s")
for flags in (0, 1):
    run("visualizer", 3840, 2160, 2, 2, flags)
run("visualizer", 1920, 1080, 1, 1)
for scene in ("shadertoy", "default", "bars", "waveform"):
    run(scene, 3840, 2160, 2, 2)
for scene in ("mandelbrot", "tetration", "raymarch"):
    run(scene, 7680, 4320, 4, 4)
print("launches", ctx.launches)

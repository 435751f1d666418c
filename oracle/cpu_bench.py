"""
TEST / BENCH INFRASTRUCTURE ONLY (oracle, kind "port"). Times the CPU restatement of the hot path on the
host cores for bench.py's `cpu_baseline` object and its `--impl reference` arm. The reference itself
(llvmpipe + moderngl) cannot run here or on the GPU box (no GL stack), so the baseline is
this numpy port: label "port", not "reference".

A sample is a set of row bands of one 4K 2×SSAA visualizer frame (the render target is 7680×4320
fragments); bands are spread over worker processes (numpy shading is single-threaded per process), and
frames/s is extrapolated from fragments shaded per second. The audio track of the frame (STFT, filterbank,
dynamics, volume/std, waveform: ~0.4 ms/frame) is timed separately and added in.
"""
from __future__ import annotations

import os
import time

import numpy as np

_STATE = {}


def _setup(width: int, height: int, ssaa: int, bg_size: tuple[int, int], seed: int):
    from oracle import audio_np as A
    from oracle import glsl_np as G
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    cfg = A.TrackConfig(bank=A.BankConfig.from_notes(15, 129, piano=True))
    clip = A.synth_chirp(1.0)
    k = 45
    t0 = time.perf_counter()
    track = A.audio_track(clip, k + 1, cfg)
    audio_per_frame = (time.perf_counter() - t0)/(k + 1)
    image = G.synthetic_background(*bg_size, seed=seed)
    tex = dict(
        background=G.Texture(np.flipud(image).copy(), linear=True),
        iSpectrogram=G.Texture(track["column"][k].reshape(-1, 1, 2).copy(), linear=False, repeat_x=True, repeat_y=False),
        iWaveform=G.Texture(track["wave"][k].reshape(1, -1, 2).copy(), linear=True, repeat_x=False, repeat_y=False),
    )
    u = G.Uniforms(iTime=float(track["time"][k]), iResolution=(width, height), iWantAspect=width/height, iSSAA=float(ssaa),
                   extra=dict(iAudioVolume=float(track["volume"][k]), iAudioSTD=float(track["std"][k])))
    _STATE.update(G=G, tex=tex, u=u, Wr=width*ssaa, Hr=height*ssaa, audio_per_frame=audio_per_frame)


def _init(width, height, ssaa, bg_size, seed):
    _setup(width, height, ssaa, bg_size, seed)


def _band(args) -> int:
    row0, rows = args
    G = _STATE["G"]
    f = G.varyings(_STATE["u"], _STATE["Wr"], _STATE["Hr"], rows=slice(row0, row0 + rows))
    out = G.SCENES[_STATE.get("scene", "visualizer")](_STATE["u"], f, _STATE["tex"])
    q = G.to_unorm8(out)
    s = int(_STATE["u"].iSSAA)
    if s >= 1 and rows % s == 0 and q.shape[1] % s == 0:
        # final.glsl where it is a box filter: mean of the s x s quantised sub-samples, 8-bit store
        box = q.reshape(rows//s, s, q.shape[1]//s, s, 4).astype(np.float32).mean(axis=(1, 3))/np.float32(255)
        G.to_unorm8(box[..., :3])
    return out.shape[0]*out.shape[1]


def _init_scene(scene, width, height, ssaa):
    from oracle import glsl_np as G
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    u = G.Uniforms(iResolution=(width, height), iWantAspect=width/height, iSSAA=float(ssaa))
    _STATE.update(G=G, tex={}, u=u, Wr=width*ssaa, Hr=height*ssaa, scene=scene, audio_per_frame=0.0)


def scene_sample(scene: str, width: int, height: int, ssaa: int, rows_per_band: int = 1, bands_per_worker: int = 1,
                 workers: int | None = None) -> dict:
    """The numpy port of a texture-less scene (mandelbrot / tetration / raymarch) on row bands spread over the
    target, one band at a time per worker process → frames/s extrapolated from fragments per second"""
    import multiprocessing as mp
    workers = workers or usable_cores()
    Hr = height*ssaa
    n_bands = workers*bands_per_worker
    stride = max(rows_per_band, (Hr - rows_per_band)//max(1, n_bands - 1))
    tasks = [(min(i*stride, Hr - rows_per_band), rows_per_band) for i in range(n_bands)]
    ctx = mp.get_context("spawn")
    with ctx.Pool(workers, initializer=_init_scene, initargs=(scene, width, height, ssaa)) as pool:
        pool.map(_band, [(0, 1)]*workers)                   # warm-up: imports, first touch
        t0 = time.perf_counter()
        counts = pool.map(_band, tasks, chunksize=1)
        seconds = time.perf_counter() - t0
    fragments = int(sum(counts))
    per_frame = seconds*(width*ssaa*height*ssaa)/fragments
    return dict(frames_per_s=1.0/per_frame, fragments=fragments, seconds=seconds, cores=workers,
                sample=f"{n_bands} bands x {rows_per_band} rows of the {width*ssaa}x{height*ssaa} target ({fragments} fragments, "
                       f"{seconds:.1f} s on {workers} processes), numpy port of the {scene} fragment; 8-bit stores and the box final pass included")


def usable_cores() -> int:
    """Cores this process may really use: affinity mask capped by the cgroup CPU quota"""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            text = open(path).read().split()
            if path.endswith("cpu.max"):
                if text[0] != "max":
                    n = min(n, max(1, int(int(text[0])/int(text[1]))))
            else:
                quota = int(text[0])
                period = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                if quota > 0:
                    n = min(n, max(1, quota//period))
        except Exception:
            pass
    return max(1, n)


def visualizer_sample(width: int = 3840, height: int = 2160, ssaa: int = 2, rows_per_band: int = 4,
                      bands_per_worker: int = 3, workers: int | None = None, bg_size=(1920, 1080), seed: int = 1) -> dict:
    """Shades `workers*bands_per_worker` bands of `rows_per_band` fragment rows on `workers` processes.
    → dict(frames_per_s, fragments, seconds, cores, sample)"""
    import multiprocessing as mp
    workers = workers or usable_cores()
    Hr = height*ssaa
    n_bands = workers*bands_per_worker
    stride = max(rows_per_band, (Hr - rows_per_band)//max(1, n_bands - 1))
    tasks = [(min(i*stride, Hr - rows_per_band), rows_per_band) for i in range(n_bands)]
    ctx = mp.get_context("spawn")
    with ctx.Pool(workers, initializer=_init, initargs=(width, height, ssaa, bg_size, seed)) as pool:
        pool.map(_band, tasks[:workers])                    # warm-up: imports, first-touch
        t0 = time.perf_counter()
        counts = pool.map(_band, tasks, chunksize=1)
        seconds = time.perf_counter() - t0
    _setup(width, height, ssaa, (64, 36), seed)             # audio side, measured in this process
    fragments = int(sum(counts))
    shade_per_frame = seconds*(width*ssaa*height*ssaa)/fragments
    per_frame = shade_per_frame + _STATE["audio_per_frame"]
    return dict(frames_per_s=1.0/per_frame, fragments=fragments, seconds=seconds, cores=workers,
                audio_ms_per_frame=_STATE["audio_per_frame"]*1e3,
                sample=f"{n_bands} bands x {rows_per_band} rows of the {width*ssaa}x{height*ssaa} iScreen target "
                       f"({fragments} fragments, {seconds:.1f} s on {workers} processes), numpy port of visualizer.frag; "
                       f"audio track {_STATE['audio_per_frame']*1e3:.2f} ms/frame added; 8-bit stores and the box final pass included")


def stft_sample(seconds: float = 20.0) -> dict:
    """STFT→filterbank of the numpy port, one thread: Msamples/s consumed (hop·channels per frame)"""
    from oracle import audio_np as A
    cfg = A.TrackConfig(bank=A.BankConfig.from_notes(15, 129, piano=True))
    frames = int(seconds*60)
    clip = A.synth_chirp(seconds)
    t0 = time.perf_counter()
    A.audio_track(clip, frames, cfg, waveform=False, scalars=False)
    dt = time.perf_counter() - t0
    return dict(msamples_per_s=frames*735*2/dt/1e6, seconds=dt, frames=frames)

"""
Generates tests/golden/audio_*.npz by running the REFERENCE's own numpy code (imported from
/root/reference through oracle/ref_loader.py, GUI deps stubbed) on the synthetic inputs of
BASELINE.json's configs. Only runs in the build container; the outputs are committed.

    python tests/golden/make_golden.py

Per frame the script replays, with the reference's primitives, exactly what the module `update()`s do:
    ShaderAudio.update          audio/module.py:447-458   (reader chunk → add_data → vol/std targets)
    ShaderDynamics.update       dynamics.py:276-278
    ShaderWaveform.update       audio/waveform.py:80-87
    ShaderSpectrogram.update    audio/spectrogram.py:298-311
driven by the reference's own freewheel SchedulerTask (scheduler.py) and BrokenAudioReader.stream
(ffmpeg.py:1276-1333, its ffmpeg child replaced by an in-memory PCM pipe).
"""
from __future__ import annotations

import importlib
import io
import sys
import zlib
from pathlib import Path
from types import SimpleNamespace
from unittest.mock import patch

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import audio_np, ref_loader  # noqa: E402

R = ref_loader.load()
ref_scheduler = importlib.import_module("shaderflow.scheduler")
ref_ffmpeg = importlib.import_module("shaderflow.ffmpeg")
HERE = Path(__file__).parent


def reference_track(x: np.ndarray, n_frames: int, fps: float, notes: tuple[int, int] | None,
                    keep=(0, 1, 2, 5, 6, 7, 30, 59)) -> dict:
    sr, ch = 44100, x.shape[0]

    # -- the reference's file reader over an in-memory "ffmpeg" ---------------------------------
    pcm = io.BytesIO(np.ascontiguousarray(x.T).tobytes())
    fake_popen = SimpleNamespace(stdout=pcm)
    class FakeFFmpeg:
        get_audio_channels   = staticmethod(lambda path: ch)
        get_audio_samplerate = staticmethod(lambda path: sr)
        def __getattr__(self, name): return lambda *a, **k: self
        def popen(self, **k): return fake_popen
    with patch.object(ref_ffmpeg, "FFmpeg", FakeFFmpeg):
        reader = ref_ffmpeg.BrokenAudioReader(path=Path("/dev/null"))
        stream = reader.stream

        audio = R.BrokenAudio()
        spec = R.BrokenSpectrogram(audio=audio)
        if notes is not None:
            spec.from_notes(start=notes[0], end=notes[1], piano=True)
        bins = spec.spectrogram_bins
        sdyn = R.DynamicNumber(frequency=4, zeta=1, response=0, dtype=np.float32)
        vol = R.DynamicNumber(frequency=2, zeta=1, response=0, value=0, integrate=True)
        std = R.DynamicNumber(frequency=10, zeta=1, response=0, value=0)
        points, chunk = 180, max(1, int(3*sr/180))

        out = dict(
            tell=np.zeros(n_frames, np.int64), time=np.zeros(n_frames), dt=np.zeros(n_frames),
            spec=np.zeros((n_frames, ch, bins), np.float32), column=np.zeros((n_frames, bins, ch), np.float32),
            vol_target=np.zeros(n_frames, np.float32), std_target=np.zeros(n_frames, np.float32),
            volume=np.zeros(n_frames), volume_integral=np.zeros(n_frames), std=np.zeros(n_frames),
            wave=np.zeros((n_frames, points, ch), np.float32),
            mag_frames=np.array([k for k in keep if k < n_frames]),
        )
        out["mag"] = np.zeros((len(out["mag_frames"]), ch, spec.fft_bins), np.float32)
        scene = SimpleNamespace(time=0.0, dt=0.0, rdt=0.0, frame=0)

        def frame(dt: float = 0.0):
            k = scene.frame
            out["tell"][k], out["time"][k], out["dt"][k] = 0, scene.time, scene.dt
            # ShaderAudio.update
            try:
                reader.chunk = scene.rdt
                audio.add_data(next(stream).T)
            except StopIteration:
                pass
            out["tell"][k] = audio.tell
            vol.target = 2 * R.root_mean_square(audio.get_last_n_seconds(0.1)) * (2**0.5)
            std.target = np.std(audio.get_last_n_seconds(0.1))
            out["vol_target"][k], out["std_target"][k] = vol.target, std.target
            # ShaderDynamics.update ×2
            vol.next(dt=abs(scene.dt)); std.next(dt=abs(scene.dt))
            out["volume"][k], out["volume_integral"][k], out["std"][k] = vol.value, vol.integral, std.value
            # ShaderWaveform.update
            off = audio.tell % chunk
            c = audio.data[:, -int(chunk*points + off + 1):-int(off + 1)].reshape(ch, -1, chunk)
            out["wave"][k] = np.ascontiguousarray(R.WaveformReducer.Average(c).T)
            # ShaderSpectrogram.update
            if sdyn.value.shape != (ch, bins):
                sdyn.set(np.zeros((ch, bins), np.float32))
            if k in out["mag_frames"]:
                out["mag"][list(out["mag_frames"]).index(k)] = spec.fft()
            nxt = spec.next()
            out["spec"][k] = nxt
            sdyn.target = nxt.T.reshape(2, -1)
            sdyn.next(dt=abs(scene.dt))
            out["column"][k] = np.frombuffer(sdyn.value.astype(np.float32).tobytes(), np.float32).reshape(bins, ch)
            # ShaderScene.next tail (scene.py:476-479)
            scene.dt, scene.rdt = dt*1.0, dt
            scene.time += scene.dt
            scene.frame += 1

        task = ref_scheduler.SchedulerTask(task=frame, frequency=fps, freewheel=True, precise=True)
        for _ in range(n_frames):
            task.next()

    M = spec.spectrogram_matrix()
    out["bank_indptr"], out["bank_indices"], out["bank_data"] = M.indptr, M.indices, M.data
    out["bank_frequencies"] = spec.spectrogram_frequencies
    out["input_crc32"] = np.array(zlib.crc32(x.tobytes()), np.uint32)
    return out


def main():
    cases = dict(
        # BASELINE.json configs[0]: 1 s 440 Hz sine, piano bins of examples/basic Visualizer
        audio_c1_sine=(audio_np.synth_sine(1.0), 60, 60.0, (15, 129)),
        # configs[1]-style broadband input (1 s of the seeded white noise)
        audio_noise=(audio_np.synth_noise(1.0, seed=0), 60, 60.0, (15, 129)),
        # configs[2]-style chirp, shortened; default 1000-bin bank; non-integer hop (fps 24 → 1837.5)
        audio_chirp_1000=(audio_np.synth_chirp(1.5), 30, 24.0, None),
        # clip shorter than the export: reader hits EOF mid-way
        audio_short=(audio_np.synth_noise(0.25, seed=3), 30, 60.0, (15, 129)),
    )
    for name, (x, frames, fps, notes) in cases.items():
        out = reference_track(x, frames, fps, notes)
        np.savez_compressed(HERE/f"{name}.npz", **out)
        print(name, {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim}, "tell", out["tell"][:4], out["tell"][-1])

    # Resolution.fit known answers (the only test the reference ships: resolution.py:90-116)
    fit = R.Resolution.fit
    table = [
        (dict(old=(1920, 1080)), None),
        (dict(old=(1920, 1080), new=(1280, None)), None),
        (dict(old=(1920, 1080), new=(None, 720)), None),
        (dict(old=(1920, 1080), new=(1280, None), ar=16/9), None),
        (dict(old=(1920, 1080), new=(None, 720), ar=16/9), None),
        (dict(old=(1920, 1080), new=(1000, 720), ar=2), None),
        (dict(old=(3840, 2160), new=(3800, 2100), max=(1920, 1080)), None),
        (dict(old=(3000, 3000), new=(2000, 2000), max=(6000, 720), ar=16/9), None),
        (dict(old=(1920, 1080), new=(3840, 2160), scale=0.5), None),
        (dict(old=(1921, 1081), new=(None, None)), None),
    ]
    import json
    rows = [dict(kwargs={k: list(v) if isinstance(v, tuple) else v for k, v in kw.items()},
                 result=list(fit(**kw))) for kw, _ in table]
    (HERE/"resolution_fit.json").write_text(json.dumps(rows, indent=1))
    print("resolution_fit", len(rows))


if __name__ == "__main__":
    main()

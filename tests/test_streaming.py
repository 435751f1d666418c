"""Streaming `BrokenAudio.add_data()` (audio/module.py:113-129): a scene that feeds its audio frame by frame.

No GPU here: the host logic — the growing clip, the recorded clock, which frames are transformed when, the scan over
all frames so far — runs against a stand-in context whose two calls are the oracle's functions on CPU tensors. The
kernels themselves are held to the same oracle in test_gpu_audio.py; tests/test_gpu_zstream.py runs the real thing."""
import numpy as np
import pytest
import torch

from oracle import audio_np as A


def reference_ring(chunks, channels: int, size: int):
    """The reference's ring, literally (audio/module.py:113-129): roll left, write at the end"""
    ring, tell = np.zeros((channels, size), np.float32), 0
    for chunk in chunks:
        n = chunk.shape[1]
        if n:                                   # (the reference itself cannot take an empty chunk: data[:, -0:] is the whole ring)
            ring = np.roll(ring, -n, axis=1)
            ring[:, -n:] = chunk[:, -size:]
        tell += n
        yield ring, tell


def test_add_data_keeps_the_reference_ring():
    from shaderflow_b200.audio.module import BrokenAudio
    rng = np.random.default_rng(3)
    audio = BrokenAudio()
    audio.samplerate, audio.buffer_seconds = 8000, 0.5          # a 4 000-sample ring, overrun several times
    sizes = [1, 133, 133, 0, 134, 4000, 2500, 7, 70000, 1]
    chunks = [rng.uniform(-1, 1, (2, n)).astype(np.float32) for n in sizes]
    for chunk, (ring, tell) in zip(chunks, reference_ring(chunks, 2, 4000)):
        assert audio.add_data(chunk) is not None
        assert audio.tell == tell and audio.streaming and audio.total_samples == tell
        assert np.array_equal(audio.data, ring)
        for n, off in ((256, 0), (4096 - 1, 0), (100, 35)):
            if n + off + 1 <= 4000:
                assert np.array_equal(audio.get_last_n_samples(n, offset=off), ring[:, -(n + off + 1):-(off + 1)])
    assert np.array_equal(audio.get_last_n_seconds(0.1), ring[:, -801:-1])
    with pytest.raises(ValueError):
        audio.add_data(np.zeros((1, 4), np.float32))
    loaded = BrokenAudio().load(np.zeros((2, 100), np.float32), 8000)
    with pytest.raises(RuntimeError):
        loaded.add_data(np.zeros((2, 4), np.float32))
    # load() ends a stream
    audio.load(np.ones((2, 10), np.float32), 8000)
    assert not audio.streaming and audio.tell == 0 and audio.total_samples == 10


def test_stream_device_clip_uploads_only_new_samples():
    from shaderflow_b200.audio.module import BrokenAudio
    audio = BrokenAudio()
    rng = np.random.default_rng(4)
    total = np.zeros((2, 0), np.float32)
    seen = []
    for n in (5, 1 << 20, 3, (1 << 21) + 11):                  # crosses both the host and the device capacity
        chunk = rng.uniform(-1, 1, (2, n)).astype(np.float32)
        total = np.concatenate([total, chunk], axis=1)
        before = audio._stream_uploaded if audio.streaming else 0
        audio.add_data(chunk)
        buf = audio._stream_device_clip("cpu")
        seen.append(audio._stream_uploaded - before)
        assert buf.shape[0] == 2 and buf.shape[1] >= audio.tell and buf.is_contiguous()
        assert np.array_equal(buf[:, :audio.tell].numpy(), total)
        assert not buf[:, audio.tell:].any()                  # the kernels rely on zeros past `tell` never being read, not on this
    assert seen == [5, 1 << 20, 3, (1 << 21) + 11]


class OracleContext:
    """Stand-in for _native.Context on CPU tensors: sfb_stft_mel / sfb_audio_track restated with oracle/audio_np"""
    torch_device = "cpu"

    def __init__(self):
        self.calls = []

    def sync(self):
        pass

    def stft_mel(self, pcm, tell, fft_n, csr=None, *, window=0, magnitude=0, volume=0, mag_out=None, spec_out=None):
        x = pcm.numpy()
        indptr, indices, data, bins = csr
        bank = (indptr.numpy(), indices.numpy(), data.numpy())
        self.calls.append(("stft", int(tell.shape[0])))
        for k, t in enumerate(tell.tolist()):
            mag = A.fft_magnitude(A.last_n(x, t, 1 << fft_n), window, magnitude)
            spec = A.volume(volume, A.filterbank_apply(bank, mag)).astype(np.float32)
            spec_out[k] = torch.from_numpy(np.ascontiguousarray(spec.T))

    def audio_track(self, pcm, samplerate, tell, dt, *, spec=None, bins=0, dynamics=(4.0, 1.0, 0.0, 1e-6), scalars=None,
                    wave=None, wave_points=180, wave_chunk=735, wave_reducer=0):
        x, frames = pcm.numpy(), int(tell.shape[0])
        assert dt.shape[0] == frames
        self.calls.append(("track", frames, spec is not None, scalars is not None, wave is not None))
        steps = [abs(v) for v in dt.tolist()]
        if spec is not None:
            ch = x.shape[0]
            dyn = A.Dynamics(frequency=dynamics[0], zeta=dynamics[1], response=dynamics[2], value=np.zeros((ch, bins), np.float32))
            for k in range(frames):
                dyn.target = spec[k].numpy().reshape(ch, -1).copy()
                dyn.next(dt=steps[k])
                spec[k] = torch.from_numpy(dyn.value.astype(np.float32).reshape(bins, ch))
        if scalars is not None:
            vol = A.Dynamics(frequency=2, zeta=1, response=0, integrate=True)
            std = A.Dynamics(frequency=10, zeta=1, response=0)
            for k, t in enumerate(tell.tolist()):
                vt, st = A.volume_std_targets(x, t, samplerate)
                vol.target, std.target = np.array(vt, dtype=np.float32), np.array(st, dtype=np.float32)
                vol.next(dt=steps[k]); std.next(dt=steps[k])
                scalars[k] = torch.tensor([float(vol.value), float(vol.integral), float(std.value), float(vt), float(st)], dtype=torch.float64)
        if wave is not None:
            for k, t in enumerate(tell.tolist()):
                wave[k] = torch.from_numpy(A.waveform_row(x, t, samplerate, wave_points*wave_chunk/samplerate,
                                                          wave_points/(wave_points*wave_chunk/samplerate), wave_reducer))


class NoTexture:
    """Stand-in for _native.Texture: remembers what it was given, owns nothing"""
    def __init__(self, ctx, width, height, components, dtype, linear=True, repeat_x=True, repeat_y=True):
        self.width, self.height, self.components, self.dtype, self.bound = width, height, components, dtype, None

    def bind_external(self, pointer): self.bound = pointer
    def write(self, *a, **k): pass
    def set_sampling(self, *a, **k): pass
    def destroy(self): pass


@pytest.fixture
def no_device(monkeypatch):
    from shaderflow_b200 import _native as N
    monkeypatch.setattr(N, "Texture", NoTexture)


def streamed_visualizer(clip: np.ndarray, tell: np.ndarray, seen: list):
    import examples.demo as demo

    class Streamed(demo.Visualizer):
        """Feeds frame k the samples the reference's file reader would have delivered (ffmpeg.py:1306-1330)"""
        fed = 0

        def update(self):
            upto = int(tell[min(self.frame_index, len(tell) - 1)])
            self.audio.add_data(clip[:, self.fed:upto])
            self.fed = upto

        def render(self):
            seen.append(dict(volume=float(self.audio.volume.value), integral=float(self.audio.volume.integral),
                             std=float(self.audio.std.value), tell=self.audio.tell,
                             column=self.spectrogram.columns[-1].numpy().copy(),
                             bound=self.spectrogram.texture.external[1] == self.spectrogram.columns[-1].data_ptr(),
                             wave=self.waveform.rows[0].numpy().copy()))
    return Streamed(backend="dry")


def test_streamed_export_publishes_what_the_whole_clip_export_would(no_device):
    """20 frames of a noise burst fed by add_data() from the scene's update(): every frame's volume / integral / std,
    spectrogram column and waveform row equal the oracle's whole-clip track — i.e. the reference fed the same way"""
    frames = 20
    clip = (A.synth_noise(0.5, seed=11)*np.linspace(0, 1, 22050, dtype=np.float32)).astype(np.float32)
    cfg = A.TrackConfig(bank=A.BankConfig.from_notes(15, 129, piano=True))
    want = A.audio_track(clip, frames, cfg)
    seen = []
    scene = streamed_visualizer(clip, want["tell"], seen)
    scene.initialize()
    fake = scene.cuda = OracleContext()
    scene.main(width=64, height=36, time=frames/60, fps=60.0, output=None, distributed=False)
    assert len(seen) == frames and [s["tell"] for s in seen] == want["tell"].tolist()
    for k, s in enumerate(seen):
        assert s["bound"]
        assert s["volume"] == want["volume"][k] and s["integral"] == want["volume_integral"][k] and s["std"] == want["std"][k]
        assert np.array_equal(s["column"], want["column"][k])
        assert np.array_equal(s["wave"], want["wave"][k])
    # each frame is transformed once (K1 sees one new frame per update); only the recurrences rerun from frame 0
    assert [c[1] for c in fake.calls if c[0] == "stft"] == [1]*frames
    scans = [c[1] for c in fake.calls if c[0] == "track" and c[2]]
    assert scans == list(range(1, frames + 1))
    assert scene.audio.stream["frames"] == frames


def test_frames_of_other_ranks_are_recorded_but_not_computed(no_device):
    """Sharded exports step every frame on every rank (scene.py: render_enabled False for foreign frames): the clock
    still records them, the STFT of skipped frames is caught up in one call, and the published state is unchanged"""
    frames = 12
    clip = A.synth_noise(0.3, seed=5).astype(np.float32)
    cfg = A.TrackConfig(bank=A.BankConfig.from_notes(15, 129, piano=True))
    want = A.audio_track(clip, frames, cfg)
    seen = []
    scene = streamed_visualizer(clip, want["tell"], seen)
    scene.initialize()
    fake = scene.cuda = OracleContext()
    scene.main(width=64, height=36, time=frames/60, fps=60.0, output=None, distributed=False, frames=(8, 12))
    assert len(seen) == 4
    for k, s in zip(range(8, 12), seen):
        assert s["volume"] == want["volume"][k] and np.array_equal(s["column"], want["column"][k])
        assert np.array_equal(s["wave"], want["wave"][k])
    assert [c[1] for c in fake.calls if c[0] == "stft"] == [1]*frames       # length 1 texture: transformed, not scanned
    assert [c[1] for c in fake.calls if c[0] == "track" and c[2]] == [9, 10, 11, 12]


@pytest.mark.parametrize("fps,seconds_of_audio,seconds", [(24.0, 0.3, 0.5), (60.0, 0.5, 0.4)])
def test_streamed_export_at_the_device_tests_settings(no_device, fps, seconds_of_audio, seconds):
    """The two set-ups tests/test_gpu_zstream.py exports on the device — a non-integer hop (44100/24) with the audio
    ending before the export does, and 60 fps — here against the stand-in context: the chunks come from the library's own
    frame clock (sfb_frame_clock, host code), and every frame publishes the oracle's whole-clip state"""
    from shaderflow_b200 import _native as N, synthetic
    clip = synthetic.noise(seconds_of_audio)
    frames = round(seconds*fps)
    _, dt, tell = N.frame_clock(frames, fps, 1.0, 44100, 2, clip.shape[1])
    assert tell[-1] == clip.shape[1] if seconds_of_audio < seconds else tell[-1] < clip.shape[1]
    cfg = A.TrackConfig(fps=fps, bank=A.BankConfig.from_notes(15, 129, piano=True))
    want = A.audio_track(clip, frames, cfg)
    assert np.array_equal(want["tell"], tell) and np.array_equal(want["dt"], dt)
    seen = []
    scene = streamed_visualizer(clip, tell, seen)
    scene.initialize()
    scene.cuda = OracleContext()
    scene.main(width=64, height=36, time=seconds, fps=fps, output=None, distributed=False)
    assert len(seen) == frames and [s["tell"] for s in seen] == tell.tolist()
    for k, s in enumerate(seen):
        assert s["volume"] == want["volume"][k] and s["std"] == want["std"][k] and s["integral"] == want["volume_integral"][k]
        assert np.array_equal(s["column"], want["column"][k]) and np.array_equal(s["wave"], want["wave"][k])

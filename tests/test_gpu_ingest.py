"""GPU side of the audio file ingestion: sfb_pcm_ingest against the numpy conversion for every
sample format, and an export whose audio comes from a WAV on disk against the same export fed the clip as a numpy
array — byte for byte, including a non-integer hop (24 fps at 44.1 kHz) and a file that ends before the export."""
import numpy as np
import pytest

from tests.test_audio_reader import CASES, write_wav

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def ctx():
    from shaderflow_b200 import _native as N
    c = N.Context(0)
    yield c
    c.destroy()


@pytest.mark.parametrize("name,tag,bits", CASES)
def test_ingest_kernel_equals_the_host_conversion(ctx, tmp_path, name, tag, bits):
    from shaderflow_b200.audio import reader as R
    rng = np.random.default_rng(bits)
    frames, channels = 100_003, 2                              # several staging buffers, an odd tail
    if tag == 3:
        data = rng.uniform(-1, 1, (frames, channels))
    elif bits == 8:
        data = rng.integers(0, 256, (frames, channels))
    else:
        data = rng.integers(-(1 << (bits - 1)), 1 << (bits - 1), (frames, channels))
    path = tmp_path/f"{name}.wav"
    write_wav(path, data, 44100, tag, bits)
    info = R.parse_wav(path)
    clip = R.upload_wav(ctx, info, 0, staging_bytes=64 << 10)
    ctx.sync()
    want, _ = R.read_wav(path)
    assert np.array_equal(clip.cpu().numpy(), want)


@pytest.mark.parametrize("fps,seconds_of_audio,seconds", [(60.0, 0.5, 0.5), (24.0, 0.5, 0.5), (60.0, 0.2, 0.5)])
def test_export_from_a_wav_file_equals_the_numpy_clip_export(tmp_path, fps, seconds_of_audio, seconds):
    from shaderflow_b200 import synthetic
    from examples.demo import Visualizer, synthetic_background
    pcm = synthetic.noise(seconds_of_audio)
    quantised = np.round(pcm*32767).astype(np.int64)           # what a 16-bit file can hold
    write_wav(tmp_path/"clip.wav", quantised.T, 44100, 1, 16)
    Visualizer.background = synthetic_background(240, 135)
    try:
        flags = dict(width=320, height=180, ssaa=2, subsample=2, fps=fps, time=seconds, output=bytes)
        from_file = Visualizer(device=0); from_file.initialize()
        from_file.audio.file = tmp_path/"clip.wav"
        assert from_file.audio._wav is not None
        a = from_file.main(**flags)
        from_array = Visualizer(device=0); from_array.initialize()
        from_array.audio.load((quantised/32768).astype(np.float32), 44100)
        b = from_array.main(**flags)
    finally:
        Visualizer.background = None
    assert len(a) == len(b) == 320*180*3*round(seconds*fps)
    assert a == b
    assert torch.equal(from_file.audio.clip_device, from_array.audio.clip_device)


def test_export_from_a_flac_file_equals_the_numpy_clip_export(tmp_path):
    """The same export from a FLAC stream (host decode, csrc/flac.cu; written by the independent test encoder): the
    frames are those of the clip loaded as an array"""
    from shaderflow_b200 import synthetic
    from examples.demo import Visualizer, synthetic_background
    from tests.flac_writer import write_flac
    quantised = np.round(synthetic.noise(0.3)*32767).astype(np.int64)
    (tmp_path/"clip.flac").write_bytes(write_flac(quantised.T, blocksize=4096, subframe="fixed2", stereo="mid_side"))
    Visualizer.background = synthetic_background(240, 135)
    try:
        flags = dict(width=320, height=180, ssaa=2, subsample=2, fps=60.0, time=0.3, output=bytes)
        from_file = Visualizer(device=0); from_file.initialize()
        from_file.audio.file = tmp_path/"clip.flac"
        a = from_file.main(**flags)
        from_array = Visualizer(device=0); from_array.initialize()
        from_array.audio.load((quantised/32768).astype(np.float32), 44100)
        b = from_array.main(**flags)
    finally:
        Visualizer.background = None
    assert len(a) == 320*180*3*18 and a == b

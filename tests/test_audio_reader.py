"""Audio file ingestion: RIFF/WAVE parsing and the sample-format conversions, on the host.
scipy's own WAV fixtures (part of the image's scipy installation) and files written here in every supported
sample format are read with shaderflow_b200.audio.reader and compared with scipy.io.wavfile + the conversion
ffmpeg's `-f f32le` performs (the reference's BrokenAudioReader, ffmpeg.py:1279-1287)."""
import struct
from pathlib import Path

import numpy as np
import pytest

from shaderflow_b200 import _native as N
from shaderflow_b200.audio import reader as R


def write_wav(path, data: np.ndarray, rate: int, tag: int, bits: int, extensible=False, junk=False):
    """data (frames, channels) already in the file's sample type (uint8 / int16 / int32 (for 24 and 32) / float)"""
    frames, channels = data.shape
    width = bits//8
    if bits == 24:
        raw = (data.astype("<i4").view(np.uint8).reshape(-1, 4)[:, :3]).tobytes()
    else:
        raw = data.astype({8: "u1", 16: "<i2", 32: "<i4" if tag == 1 else "<f4", 64: "<f8"}[bits]).tobytes()
    fmt = struct.pack("<HHIIHH", 0xFFFE if extensible else tag, channels, rate, rate*width*channels, width*channels, bits)
    if extensible:
        fmt += struct.pack("<HHI", 22, bits, 3) + struct.pack("<H", tag) + b"\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71"
    chunks = b""
    if junk:
        chunks += b"LIST" + struct.pack("<I", 5) + b"abcde" + b"\x00"          # odd-sized chunk + pad byte
    chunks += b"fmt " + struct.pack("<I", len(fmt)) + fmt
    chunks += b"data" + struct.pack("<I", len(raw)) + raw
    Path(path).write_bytes(b"RIFF" + struct.pack("<I", 4 + len(chunks)) + b"WAVE" + chunks)


CASES = [("u8", 1, 8), ("s16", 1, 16), ("s24", 1, 24), ("s32", 1, 32), ("f32", 3, 32), ("f64", 3, 64)]


@pytest.mark.parametrize("name,tag,bits", CASES)
@pytest.mark.parametrize("channels", [1, 2, 3])
def test_wav_formats_decode_like_ffmpeg_f32le(tmp_path, name, tag, bits, channels):
    rng = np.random.default_rng(bits + channels)
    frames = 1000
    if tag == 3:
        data = rng.uniform(-1, 1, (frames, channels))
        want = data.astype(np.float32 if bits == 32 else np.float64).astype(np.float32)
    elif bits == 8:
        data = rng.integers(0, 256, (frames, channels))
        want = ((data.astype(np.float64) - 128)/128).astype(np.float32)
    else:
        lo, hi = -(1 << (bits - 1)), (1 << (bits - 1)) - 1
        data = rng.integers(lo, hi + 1, (frames, channels))
        data[0, 0], data[1, 0] = lo, hi
        # libswresample: the integer widened to 32 bits, converted to float (round to nearest), scaled by 2^-31
        want = ((data.astype(np.int64) << (32 - bits)).astype(np.int32).astype(np.float32)*np.float32(2.0**-31)).astype(np.float32)
    path = tmp_path/f"{name}.wav"
    write_wav(path, data, 48000, tag, bits, extensible=(channels == 3), junk=(channels == 2))
    info = R.parse_wav(path)
    assert (info.samplerate, info.channels, info.frames, info.sample_bytes) == (48000, channels, frames, bits//8)
    pcm, rate = R.read_wav(path)
    assert rate == 48000 and pcm.shape == (channels, frames) and pcm.dtype == np.float32
    assert np.array_equal(pcm, want.T)


def test_scipy_fixture_files_parse_and_match_scipy():
    from scipy.io import wavfile
    import scipy.io
    data_dir = Path(scipy.io.__file__).parent/"tests"/"data"
    seen = 0
    for path in sorted(data_dir.glob("*.wav")):
        try:
            info = R.parse_wav(path)
        except ValueError:
            continue                                           # big-endian (RIFX), u-law, odd bit depths: not PCM we take
        try:
            rate, ref = wavfile.read(path)
        except Exception:
            continue
        ref = ref.reshape(len(ref), -1)
        pcm, got_rate = R.read_wav(path)
        assert got_rate == rate and pcm.shape == (ref.shape[1], ref.shape[0]), path.name
        if ref.dtype.kind == "f":
            want = ref.astype(np.float32)
        elif ref.dtype == np.uint8:
            want = ((ref.astype(np.float64) - 128)/128).astype(np.float32)
        else:
            bits = info.sample_bytes*8                          # scipy left-justifies 24-bit samples in int32
            want = (ref.astype(np.float64)/float(1 << (ref.dtype.itemsize*8 - 1))).astype(np.float32)
        assert np.allclose(pcm.T, want, atol=2.0**-24), path.name
        seen += 1
    assert seen >= 5


def test_truncated_and_foreign_files(tmp_path):
    path = tmp_path/"cut.wav"
    write_wav(path, np.arange(-50, 50, dtype=np.int64).reshape(-1, 2), 44100, 1, 16)
    blob = path.read_bytes()
    path.write_bytes(blob[:-30])                               # the header promises more than the file holds
    info = R.parse_wav(path)
    assert info.frames == (len(blob) - 30 - info.data_offset)//4
    assert R.read_wav(path)[0].shape == (2, info.frames)
    (tmp_path/"not.wav").write_bytes(b"OggS" + bytes(64))
    with pytest.raises(ValueError, match="not a RIFF"):
        R.parse_wav(tmp_path/"not.wav")


def test_audio_module_takes_a_wav_file(tmp_path):
    """`ShaderAudio(file=...)` / `audio.file = ...` with a WAV on disk: same clip as load(pcm) of the same samples"""
    from shaderflow_b200.audio.module import BrokenAudio
    rng = np.random.default_rng(0)
    data = rng.integers(-32768, 32768, (4410, 2))
    write_wav(tmp_path/"clip.wav", data, 44100, 1, 16)
    audio = BrokenAudio()
    audio.file = tmp_path/"clip.wav"
    assert audio.samplerate == 44100 and audio.channels == 2 and audio.total_samples == 4410
    assert np.array_equal(audio.clip, (data.T/32768).astype(np.float32))
    assert audio._wav is not None and audio._wav.format == N.PCM_S16
    audio.load(audio.clip, 44100)                              # an explicit clip replaces the file association
    assert audio._wav is None


@pytest.mark.timeout(120)
@pytest.mark.parametrize("name,tag,bits", CASES)
@pytest.mark.parametrize("channels", [1, 2])
def test_wav_formats_decode_like_the_real_libswresample(tmp_path, name, tag, bits, channels):
    """The same files through FFmpeg itself — WAV demuxer, pcm_* decoder, libswresample to packed float32: what the
    reference's `ffmpeg -f f32le` child writes into its pipe (ffmpeg.py:1279-1287). From the OpenCV wheel's libraries
    (tests/avcodec_bridge.py); the rules restated in the test above are thereby pinned to the library they restate"""
    from tests import avcodec_bridge as B
    if not B.available():
        pytest.skip("no FFmpeg libraries (OpenCV wheel) on this machine")
    rng = np.random.default_rng(100 + bits + channels)
    frames = 5000
    if tag == 3:
        data = rng.uniform(-1.5, 1.5, (frames, channels))
        data[:4, 0] = (0.0, -0.0, 1e-40, 3.0000001)                    # a denormal after the cast, a value past full scale
    elif bits == 8:
        data = rng.integers(0, 256, (frames, channels)); data[:3, 0] = (0, 128, 255)
    else:
        lo, hi = -(1 << (bits - 1)), (1 << (bits - 1)) - 1
        data = rng.integers(lo, hi + 1, (frames, channels)); data[:5, 0] = (lo, hi, 0, -1, 16777217 if bits == 32 else 1)
    path = tmp_path/f"{name}.wav"
    write_wav(path, data, 44100, tag, bits, junk=(channels == 2))
    want = B.wav_as_f32le(path, channels, 44100)
    pcm, rate = R.read_wav(path)
    assert want.shape == (frames, channels) and rate == 44100
    assert np.array_equal(pcm.T.view(np.uint32), want.view(np.uint32))         # bit patterns: -0.0 and denormals included


@pytest.mark.timeout(120)
def test_other_audio_files_decode_through_ffmpegs_libraries_in_process(tmp_path, golden_dir):
    """audio/avcodec.decode_file — what `ShaderAudio(file='song.opus')` falls to without an ffmpeg binary: demuxer →
    decoder → libswresample → planar float32. Held to the native readers on the formats both can read (packed s16 / s24 /
    f32 from WAV, FLAC as packed and — through the decoder's request_sample_fmt — as PLANAR samples, the layout lossy
    codecs deliver), and refused cleanly on files without audio"""
    from shaderflow_b200.audio import avcodec
    from shaderflow_b200.audio.module import read_audio_file
    if not avcodec.available():
        pytest.skip("no FFmpeg libraries (OpenCV wheel) on this machine")
    rng = np.random.default_rng(8)
    for name, tag, bits in (("s16", 1, 16), ("s24", 1, 24), ("f32", 3, 32)):
        data = rng.uniform(-1, 1, (3000, 2)) if tag == 3 else rng.integers(-(1 << (bits - 1)), 1 << (bits - 1), (3000, 2))
        write_wav(tmp_path/f"{name}.wav", data, 22050, tag, bits)
        pcm, rate = avcodec.decode_file(tmp_path/f"{name}.wav")
        want, _ = R.read_wav(tmp_path/f"{name}.wav")
        assert rate == 22050 and pcm.dtype == np.float32 and pcm.flags.c_contiguous and np.array_equal(pcm, want)
    gold = np.load(golden_dir/"flac_ffmpeg.npz")
    for name, planar in (("s16_stereo_level8", None), ("s16_stereo_level8", "s16p"), ("s24_stereo_level8", "s32p"), ("s16_mono_level8", None), ("s16_48k", "s16p")):
        (tmp_path/"clip.flac").write_bytes(gold[f"{name}.stream"].tobytes())
        want, want_rate = R.read_flac(tmp_path/"clip.flac")
        # an unknown suffix: the dispatcher cannot use the native reader
        (tmp_path/"clip.oga").write_bytes(gold[f"{name}.stream"].tobytes())
        pcm, rate = avcodec.decode_file(tmp_path/"clip.oga", dict(request_sample_fmt=planar) if planar else None)
        assert rate == want_rate and np.array_equal(pcm, want), (name, planar)
    import shutil
    if not (shutil.which("ffmpeg") and shutil.which("ffprobe")):
        pcm, rate = read_audio_file(tmp_path/"clip.oga")
        assert rate == 48000 and np.array_equal(pcm, want)
    cv2 = pytest.importorskip("cv2")
    writer = cv2.VideoWriter(str(tmp_path/"silent.avi"), cv2.VideoWriter_fourcc(*"MJPG"), 10.0, (16, 16))
    writer.write(np.zeros((16, 16, 3), np.uint8)); writer.release()
    with pytest.raises(RuntimeError, match="no audio stream"):
        avcodec.decode_file(tmp_path/"silent.avi")
    with pytest.raises(RuntimeError, match="cannot open"):
        avcodec.decode_file(tmp_path/"missing.opus")


@pytest.mark.timeout(120)
def test_a_lossy_file_decodes_end_to_end(tmp_path):
    """An MPEG-1 Layer II file, written here by FFmpeg's own encoder (tests/avcodec_bridge.transcode; its packets are
    self-framing, so their concatenation is the file), read through `ShaderAudio(file=...)`'s dispatcher: rate, channels
    and length are the stream's, and the samples are the original tones behind the codec's 481-sample delay"""
    import shutil
    from tests import avcodec_bridge as B
    from shaderflow_b200.audio import avcodec
    from shaderflow_b200.audio.module import BrokenAudio
    if not (B.available() and avcodec.available()):
        pytest.skip("no FFmpeg libraries (OpenCV wheel) on this machine")
    if shutil.which("ffmpeg") and shutil.which("ffprobe"):
        pytest.skip("an ffmpeg binary takes precedence")
    rate, n = 44100, 1152*20
    t = np.arange(n)/rate
    tones = np.stack([0.5*np.sin(2*np.pi*440*t) + 0.2*np.sin(2*np.pi*1330*t), 0.4*np.sin(2*np.pi*660*t)], 1)
    pcm = np.rint(tones*32767).astype(np.int64)
    packets = B.transcode(pcm, "mp2", 16, 1152, rate, dict(time_base=f"1/{rate}", b=192000),
                          sync=lambda p: p[0] == 0xFF and (p[1] & 0xE0) == 0xE0)
    assert len(packets) == 20
    (tmp_path/"tones.mp2").write_bytes(b"".join(packets))
    audio = BrokenAudio()
    audio.file = tmp_path/"tones.mp2"
    assert audio.samplerate == rate and audio.channels == 2 and audio.clip.dtype == np.float32
    assert audio.total_samples == n and audio.duration == pytest.approx(n/rate)
    lag = 481                                                          # MPEG audio Layer II: 480 + 1 samples
    want = (pcm/32768).T[:, :n - 2000]
    error = audio.clip[:, lag:lag + n - 2000] - want
    assert 10*np.log10((want**2).mean()/(error**2).mean()) > 30        # 192 kb/s: ≈ 36 dB on these tones

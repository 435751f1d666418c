"""FLAC ingestion (csrc/flac.cu through sfb_flac_decode; audio/reader.read_flac) against streams written by the
independent encoder tests/flac_writer.py: every subframe type, residual coding variant, stereo mode, sample size and
header form round-trips bit for bit; damaged streams are refused. Host-side code: no GPU needed."""
import numpy as np
import pytest

from shaderflow_b200 import _native as N
from shaderflow_b200.audio import reader
from tests.flac_writer import crc8, crc16, utf8_number, write_flac


def signal(frames, channels, bits, seed=0, wasted=0):
    rng = np.random.default_rng(seed)
    t = np.arange(frames)[:, None]
    amp = (1 << (bits - 1)) - 1
    x = 0.6*amp*np.sin(2*np.pi*(0.01 + 0.003*np.arange(channels))*t) + 0.02*amp*rng.standard_normal((frames, channels))
    x = np.clip(np.rint(x), -amp - 1, amp).astype(np.int64)
    return (x >> wasted) << wasted


def decode(stream):
    info, pcm = N.flac_decode(stream)
    return info, pcm


def test_checksums_and_coded_numbers():
    assert crc8(b"123456789") == 0xF4 and crc16(b"123456789") == 0xFEE8          # CRC-8/SMBUS, CRC-16/UMTS check values
    assert utf8_number(0x7F) == b"\x7f" and utf8_number(0x80) == b"\xc2\x80" and utf8_number(0x7FF) == b"\xdf\xbf"
    assert utf8_number(0x800) == b"\xe0\xa0\x80" and utf8_number(0xFFFF) == "￿".encode() and len(utf8_number(1 << 30)) == 6


@pytest.mark.parametrize("kind", ["constant", "verbatim", "fixed0", "fixed1", "fixed2", "fixed3", "fixed4", "lpc1", "lpc2", "lpc8", "lpc32"])
@pytest.mark.parametrize("bits", [8, 16, 24])
def test_every_subframe_type_round_trips(kind, bits):
    x = signal(700, 2, bits, seed=bits)
    if kind == "constant":
        x[:] = x[0]
    stream = write_flac(x, bits=bits, blocksize=256, subframe=kind)
    info, pcm = decode(stream)
    assert (info.samplerate, info.channels, info.bits_per_sample, info.total_samples, info.has_md5) == (44100, 2, bits, 700, 1)
    assert np.array_equal(pcm, x)


@pytest.mark.parametrize("stereo", ["left_side", "side_right", "mid_side"])
@pytest.mark.parametrize("bits", [16, 24, 32])
def test_stereo_decorrelation(stereo, bits):
    x = signal(1000, 2, bits, seed=3)
    x[:, 1] = x[:, 0]//3 + signal(1000, 1, bits - 4, seed=4)[:, 0]                  # correlated channels, odd sums
    kinds = "verbatim" if bits == 32 else "fixed2"                                  # side channel of a 32-bit stream: 33-bit samples
    _, pcm = decode(write_flac(x, bits=bits, blocksize=576, subframe=kinds, stereo=stereo))
    assert np.array_equal(pcm, x)


@pytest.mark.parametrize("porder,rice2,escape", [(0, False, ()), (1, False, ()), (3, False, (1,)), (4, True, ()), (2, True, (0, 3)), (5, False, (0,))])
def test_residual_coding_variants(porder, rice2, escape):
    x = signal(4096, 1, 16, seed=porder)
    x[100:140] = 0                                                                  # a silent stretch: zero-width escape partitions
    _, pcm = decode(write_flac(x, blocksize=4096, subframe="fixed3", porder=porder, rice2=rice2, escape=escape))
    assert np.array_equal(pcm, x)
    loud = np.random.default_rng(1).integers(-2**23, 2**23, (2304, 1))              # residuals needing Rice parameters > 14
    _, pcm = decode(write_flac(loud, bits=24, blocksize=2304, subframe="fixed1", porder=2, rice2=True))
    assert np.array_equal(pcm, loud)


@pytest.mark.parametrize("frames,blocksize,rate,channels", [(5000, 1152, 44100, 1), (1000, 192, 96000, 3), (777, 100, 12345, 2), (70000, 16384, 48000, 2),
                                                            (1300, 1000, 352800, 1), (300, 4608, 8000, 8), (513, 512, 22000, 2)])
def test_header_forms(frames, blocksize, rate, channels):
    """table and explicit block sizes (8 / 16 bit), a short last block, table and explicit sample rates, 1-8 channels"""
    x = signal(frames, channels, 16, seed=frames)
    info, pcm = decode(write_flac(x, rate=rate, blocksize=blocksize, subframe=["fixed2", "lpc4", "verbatim", "fixed0"][:channels]*2 if channels > 4 else
                                  ["fixed2", "lpc4", "verbatim", "fixed0"][:channels]))
    assert info.samplerate == rate and info.channels == channels and np.array_equal(pcm, x)


def test_wasted_bits_unknown_total_id3_and_large_frame_numbers():
    x = signal(3000, 2, 16, seed=9, wasted=3)
    stream = write_flac(x, blocksize=576, subframe="fixed2", wasted=3, id3=True, total_known=False, md5=False, size_from_streaminfo=True,
                        first_number=2**31 - 2, trailing=b"TAG" + bytes(125))
    info, pcm = decode(stream)
    assert info.total_samples == 0 and info.has_md5 == 0 and np.array_equal(pcm, x)


def test_damaged_streams_are_refused():
    x = signal(2000, 2, 16, seed=5)
    stream = bytearray(write_flac(x, blocksize=576, subframe="fixed2"))
    with pytest.raises(RuntimeError, match="no fLaC marker"):
        N.flac_decode(b"RIFF" + bytes(100))
    audio = stream.index(b"sfbtest") + 11
    hurt = bytearray(stream); hurt[audio + 40] ^= 0x10
    with pytest.raises(RuntimeError, match="FLAC"):
        N.flac_decode(bytes(hurt))
    hurt = bytearray(stream); hurt[audio + 2] ^= 0x01                                # header field: CRC-8
    with pytest.raises(RuntimeError, match="CRC-8|reserved|FLAC"):
        N.flac_decode(bytes(hurt))
    with pytest.raises(RuntimeError, match="FLAC"):
        N.flac_decode(bytes(stream[:len(stream) - 7]))                               # cut inside the last frame


def test_flac_files_load_like_wav_files(tmp_path):
    from scipy.io import wavfile
    from shaderflow_b200.audio.module import read_audio_file
    x = signal(9000, 2, 16, seed=6)
    (tmp_path/"clip.flac").write_bytes(write_flac(x, subframe="lpc6", stereo="mid_side"))
    wavfile.write(tmp_path/"clip.wav", 44100, x.astype(np.int16))
    a, ra = read_audio_file(tmp_path/"clip.flac")
    b, rb = read_audio_file(tmp_path/"clip.wav")
    assert ra == rb == 44100 and a.dtype == np.float32 and a.shape == (2, 9000) and np.array_equal(a, b)
    x24 = signal(4000, 1, 24, seed=7)
    (tmp_path/"deep.flac").write_bytes(write_flac(x24, bits=24, rate=48000, subframe="fixed4"))
    c, rc = reader.read_flac(tmp_path/"deep.flac")
    assert rc == 48000 and np.array_equal(c[0], (x24[:, 0].astype(np.float32)/np.float32(2**23)))
    wrong = bytearray(write_flac(x, subframe="fixed1"))
    at = wrong.index(b"fLaC") + 8 + 18
    wrong[at] ^= 0xFF                                                                # the MD5 of STREAMINFO
    (tmp_path/"wrong.flac").write_bytes(bytes(wrong))
    with pytest.raises(ValueError, match="MD5"):
        reader.read_flac(tmp_path/"wrong.flac")
    assert reader.read_flac(tmp_path/"wrong.flac", verify=False)[0].shape == (2, 9000)


# ---------------------------------------------------------------------------------------------------------------------
# Pinned to FFmpeg: streams written by libavcodec's FLAC encoder (tests/golden/make_golden_flac.py; the goldens travel),
# and — where the OpenCV wheel's FFmpeg libraries exist — FFmpeg's decoder beside sfb_flac_decode on fresh material

def test_streams_written_by_ffmpegs_encoder_decode_bit_for_bit(golden_dir):
    import hashlib
    gold = np.load(golden_dir/"flac_ffmpeg.npz")
    names = sorted({key.rsplit(".", 1)[0] for key in gold.files})
    assert len(names) == 16
    for name in names:
        rate, channels, bits, blocksize, frames = (int(v) for v in gold[f"{name}.format"])
        info, pcm = decode(gold[f"{name}.stream"].tobytes())
        assert (info.samplerate, info.channels, info.bits_per_sample, info.total_samples, info.has_md5) == (rate, channels, bits, frames, 1), name
        assert pcm.shape == (frames, channels)
        digest = hashlib.sha256(np.ascontiguousarray(pcm.astype(np.int32)).tobytes()).digest()
        assert digest == gold[f"{name}.sha256"].tobytes(), name


def test_read_flac_on_an_ffmpeg_stream(golden_dir, tmp_path):
    """audio/reader.read_flac (what `ShaderAudio(file=...)` calls): float32 planar, scaled by 2^(bits-1)"""
    gold = np.load(golden_dir/"flac_ffmpeg.npz")
    (tmp_path/"clip.flac").write_bytes(gold["s24_stereo_level8.stream"].tobytes())
    _, pcm = decode(gold["s24_stereo_level8.stream"].tobytes())
    clip, rate = reader.read_flac(tmp_path/"clip.flac")
    assert rate == 44100 and clip.dtype == np.float32 and np.array_equal(clip, (pcm.T/2.0**23).astype(np.float32))


def _bridge():
    from tests import avcodec_bridge as B
    if not B.available():
        pytest.skip("no FFmpeg libraries (OpenCV wheel) on this machine")
    return B


@pytest.mark.timeout(120)
@pytest.mark.parametrize("bits,channels,blocksize,level,options", [
    (16, 2, 4096, 3, {}), (16, 2, 1152, 9, dict(lpc_type="levinson", lpc_coeff_precision=12)), (16, 1, 4096, 8, dict(lpc_passes=2, lpc_type="cholesky")),
    (24, 2, 4096, 10, {}), (24, 2, 576, 5, dict(ch_mode="indep")), (8, 1, 256, 5, {}), (32, 2, 2048, 8, {}), (12, 2, 4096, 5, {})])
def test_fresh_material_through_ffmpegs_encoder(bits, channels, blocksize, level, options):
    """Not the committed streams: new samples, encoded by FFmpeg now, decoded by both decoders"""
    B = _bridge()
    x = signal(2*blocksize + 333, channels, bits, seed=blocksize + level)
    stream, held, held_bits = B.encode(x, bits=bits, blocksize=blocksize, level=level, **options)
    info, pcm = decode(stream)
    assert info.bits_per_sample == held_bits and info.has_md5 == 1
    assert np.array_equal(pcm, held)
    assert np.array_equal(B.decode(stream, channels, held_bits), held)


@pytest.mark.timeout(120)
@pytest.mark.parametrize("kind", ["constant", "verbatim", "fixed0", "fixed1", "fixed2", "fixed3", "fixed4", "lpc1", "lpc2", "lpc8", "lpc32"])
def test_ffmpegs_decoder_reads_the_test_encoders_streams_alike(kind):
    """The other direction: what tests/flac_writer.py writes is FLAC to FFmpeg too, and means the same samples — so the
    round trips above test sfb_flac_decode on valid streams, not on a private dialect"""
    B = _bridge()
    for bits, stereo, wasted in ((16, "independent", 0), (24, "mid_side", 0), (8, "left_side", 0), (16, "side_right", 3)):
        x = signal(1500, 2, bits, seed=bits, wasted=wasted)
        if kind == "constant":
            x[:] = x[0]
        stream = write_flac(x, bits=bits, blocksize=512, subframe=kind, stereo=stereo)
        assert np.array_equal(B.decode(stream, 2, bits), x), (kind, bits, stereo)
        assert np.array_equal(decode(stream)[1], x)

"""FLAC streams written by FFmpeg's encoder → tests/golden/flac_ffmpeg.npz (build container only).

sfb_flac_decode (csrc/flac.cu) was developed against this repository's own test encoder (tests/flac_writer.py). These
goldens pin it to a production encoder instead: libavcodec's `flac` (FFmpeg 8, from the OpenCV wheel — tests/avcodec_bridge.py
says how it is driven), at several compression levels, sample sizes, channel counts, block sizes, LPC methods and
stereo modes. For every case the file holds the stream's bytes and the SHA-256 of the samples it was made from (int32,
frame-major) — the samples themselves would be most of the file; FFmpeg's own DECODER returned them from every stream
before it was written here, and each stream's STREAMINFO carries their MD5 in FLAC's own convention.

    python tests/golden/make_golden_flac.py
"""
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from tests import avcodec_bridge as B      # noqa: E402

CASES = [  # name, bits, channels, blocksize, compression level, encoder options
    ("s16_stereo_level0", 16, 2, 1152, 0, {}),
    ("s16_stereo_level5", 16, 2, 4096, 5, {}),
    ("s16_stereo_level8", 16, 2, 4096, 8, {}),
    ("s16_stereo_level12", 16, 2, 4608, 12, {}),
    ("s16_mono_level8", 16, 1, 576, 8, {}),
    ("s16_mid_side", 16, 2, 2304, 8, dict(ch_mode="mid_side")),
    ("s16_left_side_cholesky", 16, 2, 2304, 8, dict(ch_mode="left_side", lpc_type="cholesky", lpc_passes=3)),
    ("s16_right_side", 16, 2, 1024, 5, dict(ch_mode="right_side")),
    ("s16_fixed_predictors", 16, 2, 4096, 5, dict(lpc_type="fixed")),
    ("s16_exact_rice", 16, 2, 4096, 8, dict(exact_rice_parameters=1, min_partition_order=0, max_partition_order=8)),
    ("u8_as_wasted_bits", 8, 2, 1024, 5, {}),
    ("s20_stereo", 20, 2, 4096, 5, {}),
    ("s24_stereo_level8", 24, 2, 4096, 8, {}),
    ("s24_mono_level5", 24, 1, 2304, 5, {}),
    ("s32_stereo", 32, 2, 4096, 5, {}),
    ("s16_48k", 16, 2, 4096, 5, dict()),
]


def material(frames: int, channels: int, bits: int, seed: int) -> np.ndarray:
    """Tones, a sweep, a noise floor, a silent stretch and a full-scale burst — what exercises predictors of every order,
    constant subframes, escapes and the extremes of the sample range"""
    rng = np.random.default_rng(seed)
    t = np.arange(frames)[:, None]
    amp = (1 << (bits - 1)) - 1
    x = 0.5*amp*np.sin(2*np.pi*(0.011 + 0.004*np.arange(channels))*t + 3e-6*t*t) + 0.2*amp*np.sin(2*np.pi*0.13*t)
    x = x + 0.01*amp*rng.standard_normal((frames, channels))
    x[frames//3:frames//3 + 300] = 0
    x[frames//2:frames//2 + 40] = rng.choice([-amp - 1, amp], (40, channels))
    return np.clip(np.rint(x), -amp - 1, amp).astype(np.int64)


def main() -> None:
    assert B.available(), "needs the OpenCV wheel's FFmpeg libraries"
    out = {}
    for k, (name, bits, channels, blocksize, level, options) in enumerate(CASES):
        rate = 48000 if name.endswith("48k") else 44100
        x = material(blocksize + 517, channels, bits, seed=k)
        stream, held, held_bits = B.encode(x, bits=bits, blocksize=blocksize, rate=rate, level=level, **options)
        assert np.array_equal(B.decode(stream, channels, held_bits), held), name          # FFmpeg reads its own stream back
        out[f"{name}.stream"] = np.frombuffer(stream, np.uint8)
        out[f"{name}.sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(held.astype(np.int32)).tobytes()).digest(), np.uint8)
        out[f"{name}.format"] = np.array([rate, channels, held_bits, blocksize, held.shape[0]])
        print(f"{name:28s} {len(stream):7d} B  {len(stream)/(held.size*held_bits/8):.2f} of PCM")
    np.savez_compressed(ROOT/"tests"/"golden"/"flac_ffmpeg.npz", **out)


if __name__ == "__main__":
    main()

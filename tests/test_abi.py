"""CPU-side checks of the drop-in boundary: the library loads, exports exactly what include/sfb200.h
declares, fails loudly without a GPU, and its host-only entry points agree with the oracle."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

from oracle import audio_np as A
from shaderflow_b200 import _native as N

HEADER = Path(__file__).resolve().parents[1]/"include"/"sfb200.h"


def declared_functions() -> list[str]:
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(sfb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = N.lib()
    names = declared_functions()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} declared in sfb200.h but not exported by libsfb200.so"
        assert name in N.exported_symbols(), f"{name} has no ctypes prototype in _native.py"
    assert sorted(N.exported_symbols()) == names
    assert lib.sfb_version() == 100


def test_uniform_block_layout_matches_header():
    # 4+2+4 floats, 4 ints, 2 floats, 2+2 ints, 15 floats, 6 floats, 16*4 floats
    assert ctypes.sizeof(N.Uniforms) == 4*(4 + 2 + 4 + 4 + 2 + 2 + 2 + 15 + 6 + 64)
    assert N.Uniforms.extra.offset == ctypes.sizeof(N.Uniforms) - 4*64
    u = N.Uniforms.defaults(3840, 2160)
    assert tuple(u.iResolution) == (3840.0, 2160.0) and u.iCameraMode == 1


def test_scene_registry():
    names = [N.scene_info(i)["name"] for i in range(8)]
    assert names == ["default", "shadertoy", "visualizer", "bars", "waveform", "mandelbrot", "tetration", "raymarch"]
    vis = N.scene_info(N.scene_lookup("visualizer"))
    assert vis["extra"] == ["iAudioVolume", "iAudioSTD"]
    assert vis["samplers"] == ["background", "iSpectrogram", "iWaveform"]
    assert vis["reference"].endswith("visualizer.frag")
    with pytest.raises(RuntimeError, match="no built-in scene"):
        N.scene_lookup("nope")


@pytest.mark.parametrize("fps,sr,frames,total", [(60.0, 44100, 400, -1), (24.0, 44100, 100, 70000),
                                                  (59.94, 48000, 300, -1), (60.0, 44100, 30, 11025), (30.0, 22050, 50, 0)])
def test_frame_clock_matches_oracle(fps, sr, frames, total):
    time, dt, tell = N.frame_clock(frames, fps, 1.0, sr, 2, total)
    t2, d2, r2 = A.frame_clock(frames, fps, 1.0)
    assert np.array_equal(time, t2) and np.array_equal(dt, d2)
    assert np.array_equal(tell, A.reader_tell(r2, sr, 2, total=None if total < 0 else total))


def test_frame_clock_matches_reference_golden(golden_dir):
    for name, fps in (("audio_c1_sine", 60.0), ("audio_chirp_1000", 24.0), ("audio_short", 60.0)):
        g = np.load(golden_dir/f"{name}.npz")
        total = {"audio_c1_sine": 44100, "audio_chirp_1000": 66150, "audio_short": 11025}[name]
        time, dt, tell = N.frame_clock(len(g["tell"]), fps, 1.0, 44100, 2, total)
        assert np.array_equal(tell, g["tell"]) and np.array_equal(time, g["time"]) and np.array_equal(dt, g["dt"])


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        N.Context(0)
    handle = ctypes.c_void_p()
    assert N.lib().sfb_ctx_create(0, None, ctypes.byref(handle)) == N.ECUDA
    assert b"CUDA" in N.lib().sfb_last_error() or b"device" in N.lib().sfb_last_error()


def test_argument_validation_without_gpu():
    lib = N.lib()
    assert lib.sfb_render_final(None, None, 0, 0, 0, 0, 0, 0, None) == N.EINVAL
    assert lib.sfb_frame_clock(-1, 60.0, 1.0, 44100, 2, -1, None, None, None) == N.EINVAL
    assert lib.sfb_scene_info_get(99, ctypes.byref(N.SceneInfo())) == N.EINVAL


def test_visualizer_launch_plan_is_host_only():
    """sfb_visualizer_plan needs no GPU: which kernel sfb_render_frame picks for the headline scene"""
    from shaderflow_b200 import _native as N
    u = N.Uniforms.defaults(3840, 2160); u.iTime = 1.0; u.extra[0][0] = 0.8
    rows, window = N.visualizer_plan(u, (1920, 1080), 3840, 2160, 2)           # BASELINE configs[2]
    assert rows == 8 and 20 <= window <= 40
    assert N.visualizer_plan(u, (3840, 2160), 3840, 2160, 2)[0] == 4           # coarser texel step: 4 rows per thread
    assert N.visualizer_plan(u, (1920, 1080), 3840, 2160, 4)[0] == 8
    assert N.visualizer_plan(u, (1920, 1080), 3840, 2160, 3) == (0, 0)         # ssaa 3: tiled kernel
    small = N.Uniforms.defaults(1920, 1080); small.extra[0][0] = 0.8
    rows, window = N.visualizer_plan(small, (1920, 1080), 1920, 1080, 1)       # BASELINE configs[1], the reference's default
    assert rows == 3 and window <= 32                                          # export: ~0.8 texel per fragment, 96-texel windows
    assert N.visualizer_plan(small, (3840, 2160), 1920, 1080, 1) == (0, 0)     # 1.6 texels per fragment: tiled kernel
    u.iCameraProjection = 2
    assert N.visualizer_plan(u, (1920, 1080), 3840, 2160, 2) == (0, 0)         # equirectangular camera is not separable
    u.iCameraProjection = 0; u.iCameraRight[1] = 0.1
    assert N.visualizer_plan(u, (1920, 1080), 3840, 2160, 2) == (0, 0)         # rotated basis

"""Shared setup for the GPU parity tests: the same inputs go to the oracle and to libsfb200."""
from __future__ import annotations

import numpy as np

from oracle import audio_np as A
from oracle import glsl_np as G


def piano_config(fps: float = 60.0, notes=(15, 129)) -> A.TrackConfig:
    bank = A.BankConfig.from_notes(*notes, piano=True) if notes else A.BankConfig()
    return A.TrackConfig(fps=fps, bank=bank)


def visualizer_inputs(frame: int = 45, bg_size=(240, 135), seconds: float = 1.0):
    """Oracle-side textures + uniforms of examples/basic Visualizer at one frame of a noise clip"""
    cfg = piano_config()
    track = A.audio_track(A.synth_noise(seconds, seed=0), frame + 1, cfg)
    image = G.synthetic_background(*bg_size)
    tex = dict(
        background=G.Texture(np.flipud(image).copy(), linear=True, repeat_x=True, repeat_y=True),
        iSpectrogram=G.Texture(track["column"][frame].reshape(-1, 1, 2).copy(), linear=False, repeat_x=True, repeat_y=False),
        iWaveform=G.Texture(track["wave"][frame].reshape(1, -1, 2).copy(), linear=True, repeat_x=False, repeat_y=False),
    )
    extra = dict(iAudioVolume=float(track["volume"][frame]), iAudioSTD=float(track["std"][frame]))
    return tex, extra, float(track["time"][frame])


def native_textures(ctx, tex: dict):
    """oracle Texture objects → libsfb200 textures with the same contents and sampling state"""
    from shaderflow_b200 import _native as N
    out = {}
    for name, t in tex.items():
        h, w, c = t.data.shape
        dtype = N.DTYPE_U8 if t.data.dtype == np.uint8 else N.DTYPE_F32
        nt = N.Texture(ctx, w, h, c, dtype, linear=t.linear, repeat_x=t.repeat_x, repeat_y=t.repeat_y)
        nt.write(np.ascontiguousarray(t.data))
        out[name] = nt
    return out


def native_uniforms(u: G.Uniforms, scene_info: dict):
    from shaderflow_b200 import _native as N
    W, H = u.iResolution
    n = N.Uniforms.defaults(W, H)
    for key in ("iTime", "iTau", "iDuration", "iWantAspect", "iQuality", "iSSAA", "iFramerate",
                "iCameraZoom", "iCameraIsometric", "iCameraFocalLength", "iCameraOrbital", "iCameraDolly", "iCameraSeparation"):
        setattr(n, key, float(getattr(u, key)))
    n.iFrame, n.iCameraMode, n.iCameraProjection = u.iFrame, u.iCameraMode, u.iCameraProjection
    n.iLayer = int(getattr(u, "iLayer", 0))
    for key in ("iCameraPosition", "iCameraRight", "iCameraUpward", "iCameraForward", "iCameraZenith"):
        getattr(n, key)[:] = tuple(float(v) for v in getattr(u, key))
    from shaderflow_b200.shader import pack_uniforms
    return pack_uniforms(n, {name: u.extra[name] for name in scene_info["extra"]}, scene_info["extra"], scene_info.get("extra_types"))

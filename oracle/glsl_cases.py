"""
TEST INFRASTRUCTURE ONLY (oracle). Never imported by the product path.

The render-path parity cases: deterministic inputs (uniform values, textures, target geometry) shared by
  * `tests/golden/make_golden_glsl.py` — feeds them to the REFERENCE's shader text through `oracle/glsl_exec.py`
    and commits the results as `tests/golden/glsl_*.npz`;
  * `tests/test_oracle_glsl.py` — pins `oracle/glsl_np.py` (the travelling restatement) to those goldens;
  * `tests/test_gpu_golden.py` — compares libsfb200's kernels with them on the B200.

A case names the reference example scene (class in examples/basic/demo.py or examples/fractals/fractals.py),
the `ShaderProgram` of that scene whose fragment runs, and the equivalent `glsl_np.SCENES` / libsfb200 key.
"""
from __future__ import annotations

import hashlib
import json
import re
from dataclasses import dataclass, field, fields

import numpy as np

from oracle import glsl_np as G

F = np.float32


@dataclass
class Case:
    name: str
    ref_scene: str
    program: str
    scene: str
    uniforms: G.Uniforms
    tex: dict = field(default_factory=dict)
    W: int = 48                 # final resolution (iResolution)
    H: int = 27
    ssaa: float = 2.0
    target: tuple | None = None  # explicit render-target size in fragments (programs that own a fixed-size texture)
    rows: list | None = None     # fragment rows evaluated (None = all)
    cols: slice | None = None    # columns stored in the float golden (all columns are always evaluated)
    final: tuple = ()            # subsample values for which final.glsl goldens are made from this case's screen
    ref_attrs: dict = field(default_factory=dict)

    @property
    def Wr(self):
        return self.target[0] if self.target else int(self.W*self.ssaa)

    @property
    def Hr(self):
        return self.target[1] if self.target else int(self.H*self.ssaa)


def sampler_name(key: str) -> str:
    """glsl_np texture key → sampler uniform of the assembled text (texture.py:354-368: `X{t}x{l}`; the bare
    name is `#define`d to layer L−1 of the newest frame — all bare-named textures here have one layer)"""
    return key if re.search(r"\dx\d+$", key) else key + "0x0"


def exec_uniforms(u: G.Uniforms) -> dict:
    out = {f.name: getattr(u, f.name) for f in fields(u) if f.name != "extra"}
    out.update(u.extra)
    return out


def exec_samplers(tex: dict) -> dict:
    return {sampler_name(k): t for k, t in tex.items()}


def digest(case: Case) -> str:
    h = hashlib.sha1()
    h.update(json.dumps({k: np.asarray(v, np.float64).round(9).tolist() for k, v in sorted(exec_uniforms(case.uniforms).items())}).encode())
    for key in sorted(case.tex):
        t = case.tex[key]
        h.update(key.encode()); h.update(np.ascontiguousarray(t.data).tobytes())
        h.update(bytes([t.linear, t.repeat_x, t.repeat_y]))
    h.update(repr((case.W, case.H, case.ssaa, case.target, case.rows)).encode())
    return h.hexdigest()

# ---------------------------------------------------------------------------------------------- #

def background(w=192, h=108, seed=1, **kw) -> G.Texture:
    return G.Texture(np.flipud(G.synthetic_background(w, h, seed)).copy(), linear=True, **kw)


def audio_rows(seed=3, bins=115, points=180):
    rng = np.random.default_rng(seed)
    spec = (rng.uniform(0, 1, (bins, 1, 2))**4*4000).astype(F)
    wave = rng.uniform(0, 0.6, (1, points, 2)).astype(F)
    return spec, wave


def visualizer_tex(bg: G.Texture, seed=3):
    spec, wave = audio_rows(seed)
    return dict(background=bg,
                iSpectrogram=G.Texture(spec, linear=False, repeat_x=True, repeat_y=False),
                iWaveform=G.Texture(wave, linear=True, repeat_x=False, repeat_y=False))


def U(W, H, ssaa=2.0, **kw):
    extra = kw.pop("extra", {})
    kw.setdefault("iWantAspect", W/H)
    return G.Uniforms(iResolution=(W, H), iSSAA=ssaa, extra=dict(extra), **kw)


def rotated_camera():
    """A free camera rolled 30° about its forward axis, displaced, zoomed, partly isometric"""
    c, s = np.cos(np.radians(30.0)), np.sin(np.radians(30.0))
    return dict(iCameraRight=(c, s, 0.0), iCameraUpward=(-s, c, 0.0), iCameraForward=(0.0, 0.0, 1.0),
                iCameraPosition=(0.1, -0.2, 0.0), iCameraZoom=1.3, iCameraIsometric=0.2, iCameraMode=0)


BANDS_4K = [0, 1436, 2160, 4312]      # first fragment row of the four 8-row bands of the 7680×4320 target


def small_cases() -> list[Case]:
    W, H = 48, 27
    bg = background()
    cases = [
        Case("default", "Basic", "iScreen", "default", U(W, H, iTau=0.37)),
        Case("default_stereo", "Basic", "iScreen", "default", U(W, H, iTau=0.11, iCameraProjection=1)),
        Case("default_equirect", "Basic", "iScreen", "default", U(W, H, iTau=0.62, iCameraProjection=2, iCameraZoom=0.7)),
        Case("default_rotated", "Basic", "iScreen", "default", U(W, H, iTau=0.8, **rotated_camera())),
        Case("shadertoy", "ShaderToy", "iScreen", "shadertoy", U(W, H, iTime=2.5)),
        Case("visualizer", "Visualizer", "iScreen", "visualizer",
             U(W, H, iTime=1.2345, iTau=0.12345, extra=dict(iAudioVolume=0.83, iAudioSTD=0.21)),
             visualizer_tex(bg), final=(2, 1)),
        Case("visualizer_quiet", "Visualizer", "iScreen", "visualizer",
             U(W, H, ssaa=1.0, iTime=7.7, iTau=0.77, extra=dict(iAudioVolume=0.05, iAudioSTD=0.0)),
             visualizer_tex(bg, seed=4), ssaa=1.0, final=(2,)),
        # the reference's default export: ssaa 1, subsample 2, target at the background's own resolution (~0.8 texel
        # per fragment: the separable kernel's 96-texel-window variant, then the unfused final pass)
        Case("visualizer_native", "Visualizer", "iScreen", "visualizer",
             U(96, 54, ssaa=1.0, iTime=2.2, iTau=0.22, extra=dict(iAudioVolume=0.9, iAudioSTD=0.3)),
             visualizer_tex(background(96, 54), seed=9), W=96, H=54, ssaa=1.0, final=(2,)),
        Case("visualizer_pillarbox", "Visualizer", "iScreen", "visualizer",
             U(W, H, iTime=3.3, iTau=0.33, iWantAspect=1.0, extra=dict(iAudioVolume=0.4, iAudioSTD=0.6)),
             visualizer_tex(bg, seed=5)),
        Case("visualizer_rotated", "Visualizer", "iScreen", "visualizer",
             U(W, H, iTime=5.0, iTau=0.5, extra=dict(iAudioVolume=0.6, iAudioSTD=0.1), **rotated_camera()),
             visualizer_tex(bg, seed=6)),
        Case("mandelbrot", "Mandelbrot", "iScreen", "mandelbrot", U(W, H)),
        Case("tetration", "Tetration", "iScreen", "tetration", U(W, H)),
        Case("raymarch", "RayMarch", "iScreen", "raymarch", U(W, H)),
        Case("raymarch_rotated", "RayMarch", "iScreen", "raymarch", U(W, H, **rotated_camera())),
    ]
    spec, wave = audio_rows(7)
    cases += [
        Case("bars", "MusicBars", "iScreen", "bars", U(W, H),
             dict(iSpectrogram=G.Texture(spec, linear=False, repeat_x=True, repeat_y=False))),
        Case("waveform", "Waveform", "iScreen", "waveform", U(W, H),
             dict(iWaveform=G.Texture(wave, linear=False, repeat_x=False, repeat_y=False))),
        Case("multishader_child", "MultiShader", "child", "multishader_child", U(W, H)),
        Case("dynamics", "Dynamics", "iScreen", "dynamics", U(W, H, extra=dict(iShaderDynamics=0.42)), dict(background=bg)),
        Case("audio", "Audio", "iScreen", "audio", U(W, H, extra=dict(iAudioVolume=0.37))),
    ]
    # programs that sample other programs' textures: the sampled textures are seeded 8-bit / float images
    rng = np.random.default_rng(11)
    Wr, Hr = 2*W, 2*H
    rgba = lambda seed: np.concatenate([np.flipud(G.synthetic_background(Wr, Hr, seed)),
                                        np.full((Hr, Wr, 1), 255, np.uint8)], -1)
    clamp = dict(repeat_x=False, repeat_y=False)
    cases.append(Case("multishader", "MultiShader", "iScreen", "multishader", U(W, H),
                      dict(child=G.Texture(rgba(21), linear=True))))      # only iScreen is repeat(False), scene.py:193))
    cases.append(Case("multipass_layer0", "Multipass", "iScreen", "multipass", U(W, H, iLayer=0), dict(background=bg)))
    cases.append(Case("multipass_layer1", "Multipass", "iScreen", "multipass", U(W, H, iLayer=1),
                      dict(background=bg, iScreen0x0=G.Texture(rgba(22), linear=True, **clamp))))
    cases.append(Case("motionblur_layer0", "MotionBlur", "iScreen", "motionblur",
                      U(W, H, iLayer=0, extra=dict(iScreenTemporal=10)), dict(background=bg)))
    history = {f"iScreen{t}x0": G.Texture(rgba(30 + t), linear=True, **clamp) for t in range(10)}
    cases.append(Case("motionblur_layer1", "MotionBlur", "iScreen", "motionblur",
                      U(W, H, iLayer=1, extra=dict(iScreenTemporal=10)), dict(background=bg, **history)))
    lw, lh = 32, 18
    cells = lambda: (rng.uniform(0, 1, (lh, lw, 1)) > 0.6).astype(F)
    life = dict(linear=False, repeat_x=True, repeat_y=True)
    for frame in (6, 7):
        cases.append(Case(f"life_simulation_f{frame}", "Life", "iLife", "life_simulation",
                          U(W, H, iFrame=frame, extra=dict(iLifePeriod=6, iLifeSize=(lw, lh))),
                          dict(iLife1x0=G.Texture(cells(), **life)), target=(lw, lh)))
    cases.append(Case("life_visuals", "Life", "iScreen", "life_visuals", U(W, H),
                      {f"iLife{t}x0": G.Texture(cells(), **life) for t in range(5)}))
    return cases


def band_case() -> Case:
    """BASELINE configs[2]'s geometry: 3840×2160, ssaa 2 → 7680×4320 fragments, 1920×1080 background; four
    8-row bands of the target (bottom edge, lower third, centre, top edge)"""
    W, H = 3840, 2160
    rows = [r + k for r in BANDS_4K for k in range(8)]
    return Case("visualizer_4k_bands", "Visualizer", "iScreen", "visualizer",
                U(W, H, iTime=16.6833, iTau=0.27805, iDuration=60.0, extra=dict(iAudioVolume=0.71, iAudioSTD=0.33)),
                visualizer_tex(background(1920, 1080), seed=8), W=W, H=H, rows=rows, cols=slice(3, None, 8), final=(2,))


FINAL_GEOMETRIES = [(1.0, 2), (2.0, 2), (3.0, 2), (4.0, 2), (4.0, 4), (1.0, 1), (2.0, 1), (1.5, 2), (2.0, 3)]


def final_screen(W: int, H: int, ssaa: float, seed: int = 40) -> np.ndarray:
    """A seeded RGBA8 iScreen of the render resolution for the final.glsl-only cases"""
    Wr, Hr = int(W*ssaa), int(H*ssaa)
    img = np.flipud(G.synthetic_background(Wr, Hr, seed + int(ssaa*2)))
    rng = np.random.default_rng(seed)
    alpha = rng.integers(0, 256, (Hr, Wr, 1), dtype=np.uint8)
    return np.ascontiguousarray(np.concatenate([img, alpha], -1))

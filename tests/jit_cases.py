"""Inputs shared by the CPU and GPU tests of the run-time GLSL → CUDA path (shaderflow_b200/glsl, csrc/jit).

    corpus      tests/shaders/{plasma,sdf,bits,textured,idioms}.frag — self-contained GLSL (no std-lib call), so the mechanical
                evaluator oracle/glsl_exec.py can execute the very text the translator compiles
    stdlib      tests/shaders/stdlib.frag — calls ShaderFlow's std-lib API group by group; its golden
                (tests/golden/jit_stdlib.npz) is that text evaluated behind the REFERENCE's own header and include
                files (tests/golden/make_golden_jit.py, build container only)
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

from oracle import glsl_exec as X
from oracle import glsl_np as G

SHADERS = Path(__file__).parent/"shaders"
CORPUS = ("plasma", "sdf", "bits", "textured", "idioms")
LATE = ("offsets", "toy", "materials", "syntax", "integers", "conversions", "semantics")
"""Corpus shaders added after the round's GPU budget was spent: held to the evaluator on the host (test_glsl_host.py) and
compiled by NVRTC (test_glsl_jit.py) like the others; their device run is in tests/test_gpu_zstream.py"""
W, H = 64, 36
VARYINGS = ("stxy", "glxy", "stuv", "astuv", "gluv", "agluv")
# what shader.py:190-239 declares in front of every fragment (the names the corpus reads)
HEADER = ("out vec4 fragColor; in vec2 fragCoord; in vec2 stxy; in vec2 glxy; in vec2 stuv; in vec2 astuv; in vec2 gluv; "
          "in vec2 agluv;\nuniform float iTime; uniform int iFrame; uniform vec2 iResolution;\n")
USER_UNIFORMS = dict(iTint=(0.2, 0.5, 0.9), iGain=1.5)
STDLIB_PROBES = 7
STDLIB_HEADER = "uniform sampler2D background0x0;\n#define background background0x0\n"
STDLIB_CAMERAS = (dict(), dict(iCameraProjection=1, iCameraSeparation=0.07),
                  dict(iCameraProjection=2, iCameraZoom=0.8),
                  dict(iCameraPosition=(0.1, -0.05, 0.0), iCameraZoom=1.3, iCameraIsometric=0.4, iCameraDolly=0.5, iCameraOrbital=0.2))


def uniforms(**kw) -> G.Uniforms:
    return G.Uniforms(iTime=0.7, iTau=0.3, iFrame=3, iResolution=(W, H), iWantAspect=W/H, **kw)


def corpus_textures() -> dict:
    rng = np.random.default_rng(5)
    return dict(picture=G.Texture(G.synthetic_background(24, 16, seed=2), linear=True, repeat_x=True, repeat_y=True),
                table=G.Texture(rng.random((1, 8, 4), dtype=np.float32), linear=False, repeat_x=False, repeat_y=False))


def stdlib_textures() -> dict:
    return dict(background=G.Texture(np.flipud(G.synthetic_background(40, 24, seed=4)).copy(), linear=True, repeat_x=False, repeat_y=False))


def evaluate(name: str, Wr: int = W, Hr: int = H):
    """The corpus shader's own text through the mechanical evaluator → (fragColor (Hr, Wr, 4), discarded (Hr, Wr))"""
    # gl_FragCoord is a builtin of the fragment stage: declared for the evaluator like any other input
    machine = X.Machine(HEADER + "in vec4 gl_FragCoord;\n" + (SHADERS/f"{name}.frag").read_text())
    u = uniforms()
    f = G.varyings(u, Wr, Hr)
    n = Wr*Hr
    inputs = dict(iTime=np.float32(u.iTime), iFrame=u.iFrame, iResolution=u.iResolution, **USER_UNIFORMS)
    for key in VARYINGS:
        inputs[key] = getattr(f, key).reshape(n, 2)
    inputs["fragCoord"] = inputs["stxy"]
    jj, ii = np.mgrid[0:Hr, 0:Wr]
    inputs["gl_FragCoord"] = np.stack([ii + 0.5, jj + 0.5, np.full(ii.shape, 0.5), np.ones(ii.shape)], -1).reshape(n, 4).astype(np.float32)
    out = machine.run(n, {k: v for k, v in inputs.items() if k in machine.inputs}, corpus_textures())
    color = np.broadcast_to(out["fragColor"].a, (n, 4)).reshape(Hr, Wr, 4).astype(np.float32)
    gone = machine.discarded
    gone = np.zeros(n, bool) if gone is None else np.broadcast_to(gone, (n,))
    return color, gone.reshape(Hr, Wr)

"""The GLSL -> CUDA translator checked WITHOUT a GPU: the C++ it emits, together with the run-time headers
(csrc/jit/glsl_rt.cuh, shaderflow_rt.cuh), is compiled for the host with g++ over a shim of the few CUDA built-ins
(tests/host_shim.h), executed per fragment on the CPU, and compared with the mechanical evaluator oracle/glsl_exec.py
running the same GLSL text. Same templates, same generated `Shader`, float32 with one rounding per operation — what
differs from the GPU build is the compiler back end and libm. tests/test_gpu_jit.py repeats the comparison on the B200."""
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import glsl_exec as X
from oracle import glsl_np as G
from shaderflow_b200 import glsl
from shaderflow_b200.shader import pack_uniforms
from shaderflow_b200 import _native as N
from tests import jit_cases as J

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")

MAIN = r'''
int main(int argc, char** argv) {
    static RenderParams P;
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(&P.u, sizeof(P.u), 1, f) != 1) return 2;
    fclose(f);
    P.Wr = atoi(argv[2]); P.Hr = atoi(argv[3]); P.W = P.Wr; P.H = P.Hr;
    P.inv_Wr = 1.0/double(P.Wr); P.inv_Hr = 1.0/double(P.Hr);
    for (int k = 0; 5 + 9*k < argc; k++) {                 // texture k: file w h padded comps dtype filter rx ry
        char** a = argv + 5 + 9*k;
        DevSampler& s = P.tex[k];
        s.hw = 0; s.w = atoi(a[1]); s.h = atoi(a[2]); s.padded = atoi(a[3]); s.comps = atoi(a[4]); s.dtype = atoi(a[5]);
        s.filter = atoi(a[6]); s.rx = atoi(a[7]); s.ry = atoi(a[8]);
        const size_t bytes = size_t(s.w)*s.h*s.padded*(s.dtype == SFB_DTYPE_U8 ? 1 : 4);
        void* data = malloc(bytes);
        FILE* t = fopen(a[0], "rb");
        if (!t || fread(data, 1, bytes, t) != bytes) return 3;
        fclose(t);
        s.lin = data;
    }
    FILE* out = fopen(argv[4], "wb");
    for (int j = 0; j < P.Hr; j++)
        for (int i = 0; i < P.Wr; i++) {
            g::Shader s(P, i, j);
            s.main();
            const float c[5] = {s.fragColor.x, s.fragColor.y, s.fragColor.z, s.fragColor.w, s.sfb_discarded ? 1.0f : 0.0f};
            fwrite(c, sizeof(float), 5, out);
        }
    fclose(out);
    return 0;
}
'''


def run_on_host(tmp_path, fragment: str, header: str, uniforms: G.Uniforms, extra: dict, textures: dict, Wr: int, Hr: int, block=None):
    """`block`: callable(translation) → a packed N.Uniforms, for callers that set more than time / frame / extras"""
    translation = glsl.translate(fragment, header)
    source = ('#include "host_shim.h"\n#include "sfb200.h"\n#include "render_params.h"\n#include "glsl_rt.cuh"\n#include "shaderflow_rt.cuh"\n'
              + translation.source + MAIN)
    if not ((tmp_path/"program").exists() and (tmp_path/"program.cpp").exists() and (tmp_path/"program.cpp").read_text() == source):
        (tmp_path/"program.cpp").write_text(source)                # (same text in the same directory: the binary is reused)
        build = subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-w", "-I", str(ROOT/"tests"), "-I", str(ROOT/"include"),
                                "-I", str(ROOT/"shaderflow_b200"/"csrc"), "-I", str(ROOT/"shaderflow_b200"/"csrc"/"jit"),
                                str(tmp_path/"program.cpp"), "-o", str(tmp_path/"program")], capture_output=True, text=True)
        assert build.returncode == 0, build.stderr[-3000:]
    if block is not None:
        block = block(translation)
    else:
        W, H = uniforms.iResolution
        block = N.Uniforms.defaults(W, H)
        block.iTime, block.iTau, block.iFrame = float(uniforms.iTime), float(uniforms.iTau), int(uniforms.iFrame)
        pack_uniforms(block, {name: extra[name] for name in translation.extra}, translation.extra, translation.extra_types)
    (tmp_path/"uniforms.bin").write_bytes(bytes(block))
    args = []
    for k, name in enumerate(translation.samplers):
        t = textures[name]
        data = t.data
        if data.shape[-1] == 3:                                # stored padded to 4 components, alpha reads 1
            data = np.concatenate([data, np.full(data.shape[:-1] + (1,), 255 if data.dtype == np.uint8 else 1.0, data.dtype)], -1)
        (tmp_path/f"texture{k}.bin").write_bytes(np.ascontiguousarray(data).tobytes())
        args += [str(tmp_path/f"texture{k}.bin"), data.shape[1], data.shape[0], data.shape[2], t.data.shape[2],
                 N.DTYPE_U8 if data.dtype == np.uint8 else N.DTYPE_F32, int(t.linear), int(t.repeat_x), int(t.repeat_y)]
    run = subprocess.run([str(tmp_path/"program"), str(tmp_path/"uniforms.bin"), str(Wr), str(Hr), str(tmp_path/"out.bin"), *map(str, args)],
                         capture_output=True, text=True)
    assert run.returncode == 0, (run.returncode, run.stderr[-1000:])
    out = np.fromfile(tmp_path/"out.bin", np.float32).reshape(Hr, Wr, 5)
    return out[..., :4], out[..., 4] > 0.5


@pytest.mark.parametrize("name", J.CORPUS + J.LATE)
def test_emitted_code_equals_the_evaluated_text_on_the_host(tmp_path, name):
    want, gone = J.evaluate(name)
    got, discarded = run_on_host(tmp_path, (J.SHADERS/f"{name}.frag").read_text(), J.HEADER, J.uniforms(), J.USER_UNIFORMS,
                                 J.corpus_textures(), J.W, J.H)
    assert np.array_equal(discarded, gone)
    err = np.abs(got - want)[~gone]
    assert err.max() <= 1e-3, (name, err.max())
    assert np.median(err) <= 1e-6 and (err <= 1e-5).mean() >= 0.99, (name, np.median(err), (err <= 1e-5).mean())


def test_swizzles_as_out_arguments_on_the_host(tmp_path):
    text = """
        void pR(inout vec2 p, float a) { p = cos(a)*p + sin(a)*vec2(p.y, -p.x); }
        float fold(inout vec3 p, out float side) { side = sign(p.x); p.x = abs(p.x); return length(p.xy); }
        void main() {
            vec3 p = vec3(gluv, 1.0);
            pR(p.xz, iTime);
            float s;
            float d = fold(p.zyx, s);
            pR(p.yx, d);
            vec4 q = vec4(p, d);
            pR(q.wy, 0.25);
            fragColor = vec4(q.xyz*0.25 + 0.5, 0.5 + 0.25*s + 0.01*q.w);
        }"""
    machine = X.Machine(J.HEADER + text)
    u = J.uniforms()
    f = G.varyings(u, J.W, J.H)
    n = J.W*J.H
    inputs = dict(iTime=np.float32(u.iTime), iFrame=u.iFrame, iResolution=u.iResolution)
    for key in J.VARYINGS:
        inputs[key] = getattr(f, key).reshape(n, 2)
    inputs["fragCoord"] = inputs["stxy"]
    out = machine.run(n, {k: v for k, v in inputs.items() if k in machine.inputs}, {})
    want = np.broadcast_to(out["fragColor"].a, (n, 4)).reshape(J.H, J.W, 4)
    got, _ = run_on_host(tmp_path, text, J.HEADER, u, {}, {}, J.W, J.H)
    assert np.abs(got - want).max() <= 2e-6


def test_std_lib_on_the_host_equals_the_references_glsl(tmp_path, golden_dir):
    """csrc/jit/shaderflow_rt.cuh against tests/golden/jit_stdlib.npz (stdlib.frag evaluated behind the reference's own
    header and include files), on the CPU"""
    gold = np.load(golden_dir/"jit_stdlib.npz")
    text = (J.SHADERS/"stdlib.frag").read_text()
    tex = {"background0x0": J.stdlib_textures()["background"]}
    for c, camera in enumerate(J.STDLIB_CAMERAS):
        for probe in range(J.STDLIB_PROBES if c == 0 else 1):
            u = J.uniforms(**camera)
            got, _ = run_with_camera(tmp_path, text, u, dict(iProbe=probe), tex)
            want = gold[f"camera{c}_probe{probe}"]
            err = np.abs(got - want)/np.maximum(1.0, np.abs(want))
            assert err.max() <= 1e-3 and (err <= 2e-5).mean() >= 0.99, (c, probe, err.max(), (err <= 2e-5).mean())


def run_with_camera(tmp_path, text, u, extra, tex):
    """run_on_host with the camera fields of `u` in the uniform block"""
    original = N.Uniforms.defaults

    def defaults(W, H):
        block = original(W, H)
        for key in ("iCameraZoom", "iCameraIsometric", "iCameraFocalLength", "iCameraOrbital", "iCameraDolly", "iCameraSeparation", "iWantAspect"):
            setattr(block, key, float(getattr(u, key)))
        block.iCameraMode, block.iCameraProjection = u.iCameraMode, u.iCameraProjection
        for key in ("iCameraPosition", "iCameraRight", "iCameraUpward", "iCameraForward", "iCameraZenith"):
            getattr(block, key)[:] = tuple(float(v) for v in getattr(u, key))
        return block
    N.Uniforms.defaults = staticmethod(defaults)
    try:
        return run_on_host(tmp_path, text, J.STDLIB_HEADER, u, extra, tex, J.W, J.H)
    finally:
        N.Uniforms.defaults = original


def test_matrix_uniforms_on_the_host(tmp_path):
    """`Uniform("mat3", name, value)` (variable.py's GlslType): packed one slot per column by pack_uniforms, read back as
    a matrix by the emitted code; the value travels as GL takes it — n*n numbers, column after column"""
    text = """
        uniform mat2 iTwist; uniform mat3 iBasis; uniform mat4 iProj; uniform float iGain;
        void main() {
            vec2 p = iTwist*gluv;
            vec3 q = iBasis*vec3(p, 1.0) + iBasis[2];
            vec4 h = iProj*vec4(q, 1.0);
            fragColor = vec4(h.xyz/h.w, iBasis[1][2] + iTwist[0].y + iProj[3][1])*iGain;
        }"""
    rng = np.random.default_rng(9)
    extra = dict(iTwist=rng.uniform(-1, 1, 4), iBasis=rng.uniform(-1, 1, (3, 3)), iProj=np.eye(4).ravel() + rng.uniform(-0.1, 0.1, 16), iGain=0.75)
    machine = X.Machine(J.HEADER + text)
    u = J.uniforms()
    f = G.varyings(u, J.W, J.H)
    n = J.W*J.H
    inputs = dict(iTime=np.float32(u.iTime), iFrame=u.iFrame, iResolution=u.iResolution,
                  **{k: np.asarray(v, np.float32).ravel() if k != "iGain" else np.float32(v) for k, v in extra.items()})
    for key in J.VARYINGS:
        inputs[key] = getattr(f, key).reshape(n, 2)
    out = machine.run(n, {k: v for k, v in inputs.items() if k in machine.inputs}, {})
    want = np.broadcast_to(out["fragColor"].a, (n, 4)).reshape(J.H, J.W, 4)
    got, _ = run_on_host(tmp_path, text, J.HEADER, u, extra, {}, J.W, J.H)
    assert np.abs(got - want).max() <= 2e-6
    with pytest.raises(RuntimeError, match="9 numbers"):
        run_on_host(tmp_path, text, J.HEADER, u, dict(extra, iBasis=np.zeros(4)), {}, J.W, J.H)

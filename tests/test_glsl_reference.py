"""The REFERENCE's own shader text through this backend's run-time path, without a GPU (build container only).

For the cases of oracle/glsl_cases.py the text the reference assembles and hands to the GL driver (its header, its whole
GLSL std-lib, camera.glsl, the example's fragment — captured by oracle/ref_scene.py from the reference's unmodified
Python) is translated by shaderflow_b200/glsl, the emitted C++ is compiled for the host over tests/host_shim.h
(tests/test_glsl_host.py's harness), run on the case's inputs and held to the committed golden of that case —
tests/golden/glsl_<case>.npz, the same text executed by the mechanical evaluator. So: reference text in, the
reference's pixels out, through the translator and the run-time headers the GPU build uses."""
import json
import shutil
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from oracle import glsl_cases as C
from tests.helpers import native_uniforms
from tests.test_glsl_host import run_on_host

ROOT = Path(__file__).resolve().parents[1]
REFERENCE = Path("/root/reference")
pytestmark = [pytest.mark.reference, pytest.mark.timeout(900),
              pytest.mark.skipif(not REFERENCE.exists() or shutil.which("g++") is None, reason="needs /root/reference and g++")]

# continuous shaders are held everywhere; the escape-time fractals flip whole fragments on an ulp
CASES = ["default", "default_stereo", "default_equirect", "default_rotated", "shadertoy", "visualizer", "visualizer_rotated",
         "mandelbrot", "tetration", "raymarch", "bars", "waveform", "dynamics", "audio", "multishader_child", "multishader",
         "multipass_layer1", "life_visuals", "visualizer_quiet", "visualizer_native", "visualizer_pillarbox", "raymarch_rotated",
         "multipass_layer0", "life_simulation_f6", "life_simulation_f7"]
# not here: motionblur reads 60 samplers (10 frames x 2 layers x 3 names), past the 20 slots of a run-time compiled program —
# its ahead-of-time kernel serves it (tests/test_gpu_golden.py)
DISCONTINUOUS = {"mandelbrot", "tetration", "raymarch", "raymarch_rotated", "life_visuals", "bars", "waveform", "visualizer",
                 "visualizer_rotated", "visualizer_quiet", "visualizer_native", "visualizer_pillarbox", "life_simulation_f6", "life_simulation_f7"}

_CAPTURE = """
import json, sys
sys.path.insert(0, sys.argv[1])
import numpy as np
from oracle import glsl_np as G, ref_scene
out = {}
for scene in sys.argv[2:]:
    cap = ref_scene.capture(scene, background=G.synthetic_background(16, 9))
    out[scene] = {name: program["fragment"] for name, program in cap["programs"].items()}
print("RESULT " + json.dumps(out))
"""


@pytest.fixture(scope="module")
def reference_text():
    """scene → program → assembled fragment text, captured in a child process (importing the reference rebinds `shaderflow`)"""
    cases = {c.name: c for c in C.small_cases()}
    scenes = sorted({cases[name].ref_scene for name in CASES})
    run = subprocess.run([sys.executable, "-c", _CAPTURE, str(ROOT), *scenes], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stderr[-3000:]
    return json.loads(next(line for line in run.stdout.splitlines() if line.startswith("RESULT "))[7:])


@pytest.mark.parametrize("name", CASES)
def test_reference_text_through_the_translator_reproduces_its_golden(tmp_path, golden_dir, reference_text, name):
    case = {c.name: c for c in C.small_cases()}[name]
    gold = np.load(golden_dir/f"glsl_{name}.npz")
    text = reference_text[case.ref_scene][case.program]
    from oracle import glsl_exec as X
    assert X.text_digest(text) == str(gold["fragment_sha1"]), "the reference's text is not the one the golden was made from"
    textures = C.exec_samplers(case.tex)

    class WithBlanks(dict):
        """A sampler the text names but this layer never reads (multipass layer 0 and its own previous output)"""
        def __missing__(self, key):
            from oracle import glsl_np as G
            return G.Texture(np.zeros((2, 2, 4), np.uint8))
    textures = WithBlanks(textures)

    def block(translation):
        info = dict(extra=translation.extra, extra_types=translation.extra_types, samplers=translation.samplers)
        return native_uniforms(case.uniforms, info) if not set(translation.extra) - set(case.uniforms.extra) else \
            native_uniforms(_with_defaults(case.uniforms, translation.extra), info)
    got, _ = run_on_host(tmp_path, text, "", case.uniforms, {}, textures, case.Wr, case.Hr, block=block)
    want = gold["screen_f32"]
    if case.rows is not None:
        got = got[case.rows]
    err = np.abs(got - want)/np.maximum(1.0, np.abs(want))          # colours before the store are not bounded by 1
    if name in DISCONTINUOUS:
        assert (err <= 1e-4).mean() >= 0.995, (name, (err <= 1e-4).mean(), err.max())
    else:
        assert err.max() <= 1e-4, (name, err.max())
    assert np.median(err) <= 1e-6


def _with_defaults(u, names):
    """Uniforms the assembled text reads but the case leaves at the scene's defaults"""
    import dataclasses
    known = C.exec_uniforms(u)
    extra = dict(u.extra)
    for name in names:
        if name not in extra:
            if name not in known:
                raise KeyError(f"the case has no value for uniform {name}")
            extra[name] = known[name]
    return dataclasses.replace(u, extra=extra)


def test_final_glsl_through_the_translator(tmp_path, golden_dir, reference_text):
    """fragment/final.glsl (the SSAA resolve every export ends with) as the reference assembles it, over the (ssaa,
    subsample) geometries of the golden — incl. the reference's default (1, 2)"""
    from oracle import glsl_np as G
    gold = np.load(golden_dir/"glsl_final.npz")
    text = reference_text["Basic"]["iFinal"]
    W, H = 40, 24
    for k, (ssaa, subsample) in enumerate(C.FINAL_GEOMETRIES):
        screen = G.Texture(C.final_screen(W, H, ssaa), linear=True, repeat_x=False, repeat_y=False)
        u = G.Uniforms(iResolution=(W, H), iWantAspect=W/H, iSSAA=float(ssaa), extra=dict(iSubsample=subsample))

        def block(translation):
            info = dict(extra=translation.extra, extra_types=translation.extra_types, samplers=translation.samplers)
            packed = native_uniforms(u, info)
            if hasattr(packed, "iSubsample"):
                packed.iSubsample = subsample
            return packed
        work = tmp_path/f"g{k}"
        work.mkdir()
        got, _ = run_on_host(work, text, "", u, {}, {"iScreen0x0": screen}, W, H, block=block)
        want = gold[f"s{ssaa}_k{subsample}_f32"]
        assert np.abs(got[..., :3] - want).max() <= 2e-6, (ssaa, subsample, np.abs(got[..., :3] - want).max())
        assert np.all(got[..., 3] == 1.0)


# ---------------------------------------------------------------------------------------------------------------------
# The CUDA std-lib (csrc/jit/shaderflow_rt.cuh) against the reference's GLSL std-lib, function by function

_ARGUMENT = {
    "float": "(gluv.x*1.3 + 0.37*float({k} + 1))", "int": "(int(stxy.x) % 5 + {k})", "bool": "(gluv.x > 0.1*float({k}))",
    "vec2": "(gluv*1.1 + vec2(0.2, -0.1)*float({k} + 1))", "vec3": "(vec3(gluv, astuv.x)*1.2 + 0.1*float({k}))",
    "vec4": "(vec4(gluv, astuv)*0.9 + 0.05*float({k}))", "mat2": "mat2(gluv, astuv + 0.1*float({k}))", "sampler2D": "background",
}
_RESULT = {"float": "vec4({r})", "int": "vec4(float({r}))", "bool": "vec4(float({r}))", "vec2": "vec4({r}, 0.0, 1.0)",
           "vec3": "vec4({r}, 1.0)", "vec4": "{r}", "mat2": "vec4(({r})[0], ({r})[1])"}


def _probe_shader(functions) -> str:
    cases = []
    for index, (rtype, name, params) in enumerate(functions):
        before, args = [], []
        for k, (direction, ptype) in enumerate(params):
            value = _ARGUMENT[ptype].format(k=k)
            if direction == "in":
                args.append(value)
            else:
                before.append(f"{ptype} held{k} = {value};")
                args.append(f"held{k}")
        call = f"{name}({', '.join(args)})"
        cases.append(f"        case {index}: {{ {' '.join(before)} fragColor = {_RESULT[rtype].format(r=call)}; break; }}")
    return "uniform int iProbe;\nvoid main() {\n    fragColor = vec4(0.0);\n    switch (iProbe) {\n" + "\n".join(cases) + "\n    }\n}\n"


def test_cuda_std_lib_equals_the_references_glsl_function_by_function(tmp_path, reference_text):
    """Every function of resources/shaders/include/shaderflow.glsl (86 with overloads) called with the same arguments
    twice: once with the reference's GLSL definition in front of the call (a user definition hides the CUDA std-lib
    entry of the same name), once without (the hand-written CUDA std-lib of csrc/jit/shaderflow_rt.cuh answers). Both go
    through the translator and run on the host; the two images must agree"""
    from oracle import glsl_np as G
    from shaderflow_b200.glsl import frontend
    from tests import jit_cases as J
    library = (REFERENCE/"shaderflow"/"resources"/"shaders"/"include"/"shaderflow.glsl").read_text()
    items, _ = frontend.parse(library, None, types=("Camera",))
    functions = [(i[1], i[2], [(d, t) for d, t, _ in i[3]]) for i in items if i[0] == "function"]
    assert len(functions) >= 80 and all(r in _RESULT and all(t in _ARGUMENT for _, t in p) for r, _, p in functions)
    probe = _probe_shader(functions)
    assembled = reference_text["Dynamics"]["iScreen"]                       # header + the whole std-lib + camera + the example
    theirs = assembled[:assembled.rindex("void main()")] + probe
    ours = probe
    background = {"background0x0": J.stdlib_textures()["background"]}
    # the comparison means something only if the two sides take their definitions from different places
    import re
    from shaderflow_b200 import glsl
    emitted = lambda text, header: set(re.findall(r"G_DEV [\w<>, :]+ (\w+)\(", glsl.translate(text, header).source))
    assert len(emitted(theirs, "")) >= 60 and emitted(ours, J.STDLIB_HEADER) == {"main"}
    worst = {}
    for index, (rtype, name, params) in enumerate(functions):
        u = J.uniforms(extra=dict(iProbe=index, iShaderDynamics=0.42))

        def block(translation):
            info = dict(extra=translation.extra, extra_types=translation.extra_types, samplers=translation.samplers)
            return native_uniforms(u, info)
        images = []
        for label, text, header in (("theirs", theirs, ""), ("ours", ours, J.STDLIB_HEADER)):
            work = tmp_path/label
            work.mkdir(exist_ok=True)
            image, _ = run_on_host(work, text, header, u, {}, background, 32, 18, block=block)
            images.append(image)
        a, b = images
        assert np.array_equal(np.isnan(a), np.isnan(b)), (index, name)
        error = np.nan_to_num(np.abs(a - b)/np.maximum(1.0, np.abs(a)), nan=0.0, posinf=0.0)
        worst[f"{name}/{len(params)}"] = max(worst.get(f"{name}/{len(params)}", 0.0), float(error.max()))
    bad = {k: v for k, v in worst.items() if v > 1e-5}
    assert not bad, bad


def test_cuda_camera_equals_the_references_camera_glsl(tmp_path, reference_text):
    """GetCamera(iCamera) under eight camera set-ups — the three projections, 2D / 3D modes, displaced, zoomed, isometric,
    rotated — once through the reference's camera.glsl, once through sfb_get_camera of csrc/jit/shaderflow_rt.cuh: every
    field of the Camera a fragment can read agrees"""
    from tests import jit_cases as J
    probe = """uniform int iProbe;
void main() {
    GetCamera(iCamera);       // the reference's macro pastes name##Mode, name##Position, …: the camera is named after its uniforms
    if (iProbe == 0) fragColor = vec4(iCamera.gluv, iCamera.stuv);
    if (iProbe == 1) fragColor = vec4(iCamera.agluv, iCamera.astuv);
    if (iProbe == 2) fragColor = vec4(iCamera.glxy, iCamera.stxy);
    if (iProbe == 3) fragColor = vec4(iCamera.origin, float(iCamera.out_of_bounds));
    if (iProbe == 4) fragColor = vec4(iCamera.target, float(iCamera.mode) + 0.1*float(iCamera.projection));
    if (iProbe == 5) fragColor = vec4(iCamera.position, iCamera.zoom) + vec4(iCamera.forward, iCamera.isometric) + 2.0*vec4(iCamera.right, iCamera.focal_length);
    if (iProbe == 6) fragColor = vec4(iCamera.up + 3.0*iCamera.zenith, iCamera.separation) + vec4(iCamera.backward + iCamera.left - iCamera.down, iCamera.orbital + 2.0*iCamera.dolly);
    if (iProbe == 7) fragColor = vec4(iCamera.plane_point, 1.0) + vec4(iCamera.plane_normal, 0.0);
}
"""
    assembled = reference_text["Dynamics"]["iScreen"]
    theirs = assembled[:assembled.rindex("void main()")] + probe
    rotated = C.rotated_camera() if hasattr(C, "rotated_camera") else {}
    cameras = list(J.STDLIB_CAMERAS) + [dict(iCameraMode=0), dict(iCameraMode=2, iCameraProjection=0, iCameraPosition=(0.3, 0.2, -1.0)),
                                        dict(iCameraProjection=1, iCameraSeparation=0.2, iCameraZoom=0.6), rotated]
    for c, camera in enumerate(cameras):
        for index in range(8):
            u = J.uniforms(extra=dict(iProbe=index, iShaderDynamics=0.42), **camera)

            def block(translation):
                info = dict(extra=translation.extra, extra_types=translation.extra_types, samplers=translation.samplers)
                return native_uniforms(u, info)
            images = []
            for label, text, header in (("theirs", theirs, ""), ("ours", probe, "")):
                work = tmp_path/label
                work.mkdir(exist_ok=True)
                images.append(run_on_host(work, text, header, u, {}, {}, 32, 18, block=block)[0])
            a, b = images
            assert np.array_equal(np.isnan(a), np.isnan(b)), (c, index)
            error = np.nan_to_num(np.abs(a - b)/np.maximum(1.0, np.abs(a)), nan=0.0, posinf=0.0)
            assert error.max() <= 1e-5, (c, camera, index, float(error.max()))

"""Run-time GLSL → CUDA path, the part that needs no GPU: the translator (shaderflow_b200/glsl) and NVRTC through the
C ABI (sfb_jit_compile). The compiled programs run in tests/test_gpu_jit.py."""
import ctypes.util
from pathlib import Path

import pytest

from shaderflow_b200 import _native as N
from shaderflow_b200 import glsl
from shaderflow_b200.glsl import cuda as emit
from tests import jit_cases as J

REFERENCE = Path("/root/reference")
needs_nvrtc = pytest.mark.skipif(not (ctypes.util.find_library("nvrtc") or Path("/usr/local/cuda/lib64/libnvrtc.so.12").exists()),
                                 reason="libnvrtc is not installed")


def body(source: str) -> str:
    return source[source.index("// functions"):]


def test_swizzles_literals_and_out_parameters():
    t = glsl.translate("""
        float f(inout vec3 v, out float w, vec2 p) { v.zx = p.yx*0.1; v.y += 1; w = v.xyz.z; return v[1]; }
        void main() { vec3 a = vec3(1); float w; fragColor = vec4(a.bgr, f(a, w, gluv)); fragColor.a = 1e-3; }""")
    code = body(t.source)
    assert "G_DEV float f(vec3& v, float& w, vec2 p)" in code
    assert "swz_set<2, 0>(v, (swz<1, 0>(p) * 0.100000001f))" in code           # the float32 value of 0.1
    assert "(v.y += 1)" in code and "swz<0, 1, 2>(v).z" in code and "v[1]" in code
    assert "swz<2, 1, 0>(a)" in code and "(fragColor.a = 0.00100000005f)" in code
    assert t.extra == [] and t.samplers == []


def test_struct_fields_win_over_swizzle_names_and_structs_get_constructors():
    t = glsl.translate("struct Pair { vec2 xy; float st; }; void main() { Pair p = Pair(gluv, 1); fragColor = vec4(p.xy, p.st, p.xy.yx.x); }")
    assert "G_DEV Pair(vec2 xy_, float st_) : xy(xy_), st(st_) {}" in t.source
    assert "vec4(p.xy, p.st, swz<1, 0>(p.xy).x)" in t.source


def test_only_what_main_reaches_takes_a_slot():
    header = "\n".join(f"uniform float iUnused{i};" for i in range(40)) + """
        uniform sampler2D a0x0; uniform sampler2D b0x0; uniform vec3 iTint; uniform int iMode; uniform float iTime;
        #define a a0x0
        vec4 bTexture(int t, int l, vec2 uv) { return texture(b0x0, uv); }
        float helper(float x) { return x*iUnused3; }"""
    t = glsl.translate("void main() { fragColor = texture(a, astuv)*vec4(iTint, iTime) + float(iMode); }", header)
    assert t.extra == ["iTint", "iMode"] and t.extra_types == ["vec3", "int"] and t.samplers == ["a0x0"]
    assert "iTime" not in t.extra                                               # a fixed field of sfb_uniforms
    assert "bTexture" not in t.source and "helper" not in t.source and "iUnused" not in t.source
    assert "vec3 iTint = sfb_extra<vec3>(0);" in t.source and "sampler2D a0x0 = sfb_sampler(0);" in t.source


def test_arrays_constants_loops_and_discard():
    t = glsl.translate("""
        const float K[3] = float[](1., 2., 3.);
        float pick(int i) { if (i > 2) discard; return K[i]; }
        void main() { float s = 0; for (int i = 0; i < K.length(); ++i) { s += pick(i); } fragColor = vec4(s); }""")
    assert "const arr<float, 3> K = arr<float, 3>{float(1.0f), float(2.0f), float(3.0f)};" in t.source     # arrays are values
    assert "{ sfb_discarded = true; return float(); }" in t.source
    assert "for (; (i < length_of(K)); (++i))" in t.source


def test_arrays_are_values():
    """Returned, assigned, passed by copy, members of structs (GLSL 3.30 §4.1.9): g::arr<T, N>, not C arrays"""
    t = glsl.translate("""
        struct Wave { float amp[3]; vec2 dir; };
        float[3] weights(float t) { return float[3](t, 1. - t, t*(1. - t)); }
        float total(float xs[3]) { xs[0] = 0.; return xs[0] + xs[1] + xs[2]; }
        void main() { float w[3] = weights(astuv.x); float copy[3]; copy = w; Wave wave = Wave(w, gluv);
                      fragColor = vec4(total(w), w[0], copy[1], wave.amp[2]); }""")
    assert "G_DEV arr<float, 3> weights(float t)" in t.source and "return arr<float, 3>{float(t)," in t.source
    assert "G_DEV float total(arr<float, 3> xs)" in t.source                   # by copy: the callee's store stays there
    assert "arr<float, 3> w = weights(swz<0>(astuv))" in t.source or "arr<float, 3> w = weights(astuv.x)" in t.source
    assert "arr<float, 3> amp;" in t.source and "G_DEV Wave(arr<float, 3> amp_, vec2 dir_)" in t.source


def test_small_print_of_the_grammar():
    """#version / #extension / #pragma / #line, layout() and invariant qualifiers, octal and suffixed literals, C++ keywords
    as GLSL names, functions used before their definition, non-square matrices"""
    t = glsl.translate("""#version 330
        #extension GL_ARB_gpu_shader5 : enable
        #pragma optimize(on)
        invariant gl_Position;
        layout(location = 0) out vec4 color;
        float f(float this, float new) { float class = this + new; return later(class); }
        void main() {
        #line 20 1
            uint a = 0xFFu, b = 017u; int d = 010;
            mat2x3 m = mat2x3(1, 2, 3, 4, 5, 6);
            color = vec4(m*gluv, f(float(a + b), float(d)));
        }
        float later(float x) { return x*x; }""", J.HEADER.replace("out vec4 fragColor; ", ""))
    assert "uint a = 255u;" in t.source and "b = 15u;" in t.source and "int d = 8;" in t.source
    assert "vec4& color = fragColor;" in t.source and "mat2x3 m = mat2x3(1, 2, 3, 4, 5, 6);" in t.source


def test_a_name_is_not_in_scope_in_its_own_initialiser():
    """camera.glsl:101 `vec2 gluv = gluv - ...;` inside a function reads the GLOBAL gluv (GLSL 3.30 §4.2.2); C++ would
    read the new variable — found by running the reference's stereoscopic camera through the translator"""
    t = glsl.translate("void main() { float x = 2.0; { float x = x*3.0; vec2 gluv = gluv - x; fragColor = vec4(gluv, x, 1.0); } }")
    assert "const auto sfb_outer1 = (x * 3.0f);" in t.source and "float x = sfb_outer1;" in t.source
    assert "vec2 gluv = sfb_outer2;" in t.source


def test_swizzles_as_out_arguments_copy_in_and_back():
    """`rotate(p.xz, a)` with an inout parameter: a swizzle is not an lvalue in C++, so the call copies in, calls, copies back"""
    text = """
        void pR(inout vec2 p, float a) { p = cos(a)*p + sin(a)*vec2(p.y, -p.x); }
        float fold(inout vec3 p, out float side) { side = sign(p.x); p.x = abs(p.x); return length(p.xy); }
        void main() { vec3 p = vec3(gluv, 1.0); pR(p.xz, iTime); float s; float d = fold(p.zyx, s); pR(p.xy, d); fragColor = vec4(p, s); }"""
    t = glsl.translate(text)
    assert "[&]() { auto sfb_arg0 = swz<0, 2>(p); pR(sfb_arg0, iTime); swz_set<0, 2>(p, sfb_arg0); }();" in t.source
    assert "auto sfb_result = fold(sfb_arg0, s); swz_set<2, 1, 0>(p, sfb_arg0); return sfb_result; }()" in t.source
    assert "pR(sfb_arg0, d)" in t.source                      # .xy is a swizzle too
    if ctypes.util.find_library("nvrtc") or Path("/usr/local/cuda/lib64/libnvrtc.so.12").exists():
        image, _ = N.jit_compile(glsl.program(t), glsl.headers())
        assert image[:4] == b"\x7fELF"


def test_translation_errors_are_reported():
    with pytest.raises(glsl.TranslationError, match="no main"):
        glsl.translate("float f() { return 1.0; }")
    with pytest.raises(glsl.TranslationError, match="noise3"):
        glsl.translate("void main() { fragColor = vec4(noise3(astuv.x), 1.0); }")
    assert "textureGrad(t, astuv" in glsl.translate("uniform sampler2D t; void main() { fragColor = textureGrad(t, astuv, vec2(0), vec2(0)); }").source
    # modf / frexp write through their second argument: a swizzle there is copied in and back like a user function's
    assert "swz_set<2, 3>(v, sfb_arg1)" in glsl.translate("void main() { vec4 v = vec4(0.0); fragColor.xy = modf(gluv, v.zw); fragColor.zw = v.zw; }").source
    with pytest.raises(glsl.TranslationError, match="expected"):
        glsl.translate("void main() { fragColor = vec4(1.0) }")
    many = "".join(f"uniform float u{i};" for i in range(17))
    with pytest.raises(glsl.TranslationError, match="more than 16"):
        glsl.translate("void main() { fragColor = vec4(" + "+".join(f"u{i}" for i in range(17)) + "); }", many)
    with pytest.raises(glsl.TranslationError, match="mat3x2"):
        glsl.translate("uniform mat3x2 m; void main() { fragColor = vec4(m[0], 1, 1); }")
    # square matrices take one slot per column (variable.py's GlslType: mat2, mat3, mat4)
    t = glsl.translate("uniform float a; uniform mat3 m; uniform mat2 k; void main() { fragColor = vec4(m[0]*a, k[1].x); }")
    assert t.extra == ["a", "m", "m", "m", "k", "k"] and t.extra_types == ["float", "mat3:0", "mat3:1", "mat3:2", "mat2:0", "mat2:1"]
    assert "mat3 m = sfb_extra_matrix<3>(1);" in t.source and "mat2 k = sfb_extra_matrix<2>(4);" in t.source
    five = "".join(f"uniform mat4 m{i};" for i in range(5))
    with pytest.raises(glsl.TranslationError, match="more than 16"):
        glsl.translate("void main() { fragColor = " + "+".join(f"m{i}[0]" for i in range(5)) + "; }", five)


def test_float_literals_carry_the_float32_value():
    assert emit.float_literal(0.1) == "0.100000001f" and emit.float_literal(1.0) == "1.0f" and emit.float_literal(1e-3) == "0.00100000005f"
    assert emit.float_literal(1e39) == "__int_as_float(0x7f800000)" and emit.float_literal(16777217.0) == "16777216.0f"


@needs_nvrtc
@pytest.mark.parametrize("name", J.CORPUS + J.LATE + ("stdlib",))
def test_corpus_compiles_to_sass(name):
    header = J.STDLIB_HEADER if name == "stdlib" else J.HEADER
    image, translation, log = glsl.build((J.SHADERS/f"{name}.frag").read_text(), header)
    assert image[:4] == b"\x7fELF" and len(image) > 10_000
    assert b"sfb_jit_screen" in image and b"sfb_jit_frame" in image
    assert "error" not in log.lower()
    if name == "textured":
        assert translation.samplers == ["picture", "table"] and translation.extra == ["iGain"]


@needs_nvrtc
def test_compile_errors_carry_the_compiler_log():
    with pytest.raises(N.CompileError) as info:
        glsl.build("void main() { fragColor = undeclared_function(gluv); }")
    assert "undeclared_function" in str(info.value) and info.value.log


REFERENCE_SCENES = ["Basic", "ShaderToy", "Visualizer", "MusicBars", "Waveform", "Mandelbrot", "Tetration", "RayMarch",
                    "MultiShader", "Multipass", "Dynamics", "Audio", "Life"]
_REFERENCE_SCRIPT = """
import json, sys
sys.path.insert(0, sys.argv[1])
from oracle import ref_scene
from shaderflow_b200 import _native as N, glsl
done = {}
for scene in sys.argv[2:]:
    for name, program in ref_scene.capture(scene)["programs"].items():
        translation = glsl.translate(program["fragment"])
        image, _ = N.jit_compile(glsl.program(translation), glsl.headers())
        done[scene + "." + name] = [image[:4] == b"\\x7fELF", translation.extra, translation.samplers]
print("RESULT " + json.dumps(done))
"""


@needs_nvrtc
@pytest.mark.reference
@pytest.mark.skipif(not REFERENCE.exists(), reason="needs /root/reference (build container)")
def test_the_references_own_assembled_programs_translate_and_compile():
    """The text the reference hands to the GL driver — its header, its whole std-lib and camera include, the example's
    fragment — goes through the translator and NVRTC unchanged (the reference's GLSL definitions then replace the CUDA
    std-lib entries of the same name). In a child process: importing the reference rebinds the `shaderflow` name."""
    import json, subprocess, sys
    root = str(Path(__file__).resolve().parents[1])
    run = subprocess.run([sys.executable, "-c", _REFERENCE_SCRIPT, root, *REFERENCE_SCENES], capture_output=True, text=True, timeout=900)
    assert run.returncode == 0, run.stderr[-3000:]
    done = json.loads(next(line for line in run.stdout.splitlines() if line.startswith("RESULT "))[7:])
    assert len(done) == 2*len(REFERENCE_SCENES) + 2 and all(ok for ok, _, _ in done.values()), done     # + MultiShader.child, Life.iLife
    assert done["Visualizer.iScreen"][1:] == [["iAudioVolume", "iAudioSTD"], ["iWaveform0x0", "iSpectrogram0x0", "background0x0"]]
    assert done["Life.iLife"][1:] == [["iLifePeriod", "iLifeSize"], ["iLife1x0"]]


@needs_nvrtc
def test_matrix_uniforms_compile_to_sass():
    image, translation, log = glsl.build("""uniform mat2 iTwist; uniform mat3 iBasis; uniform mat4 iProj;
        void main() { vec4 h = iProj*vec4(iBasis*vec3(iTwist*gluv, 1.0), 1.0); fragColor = h/h.w; }""", J.HEADER)
    assert image[:4] == b"\x7fELF" and translation.extra == ["iTwist"]*2 + ["iBasis"]*3 + ["iProj"]*4 and "error" not in log.lower()

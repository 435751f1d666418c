"""
Pins the render-path oracle to the REFERENCE'S SHADER TEXT.

`tests/golden/glsl_*.npz` were produced by executing the GLSL the reference hands to OpenGL (captured from the
reference's own Python by `oracle/ref_scene.py`) with the mechanical evaluator `oracle/glsl_exec.py`
(`tests/golden/make_golden_glsl.py`). Here:
  * `oracle/glsl_np.py` — the numpy restatement that travels to the GPU box and that the other parity tests
    use — must reproduce every golden to ≤ 1e-6 per float channel (bit-exact for most scenes);
  * where /root/reference exists, the text is captured again live and must hash to what the goldens record,
    and one case is re-executed from the live text;
  * the evaluator itself is unit-tested on the language features the in-scope shaders rely on.
"""
import hashlib

import numpy as np
import pytest

from oracle import glsl_cases as C
from oracle import glsl_exec as X
from oracle import glsl_np as G
from oracle import ref_loader

CASES = {c.name: c for c in C.small_cases()}
TOL = 1e-6
# scenes whose pixels are discontinuous / chaotic functions of their inputs: a last-ulp difference between
# numpy's float32 `power` on arrays and on scalars flips isolated pixels; fraction allowed to differ
LOOSE = {"tetration": 0.01}


def load(golden_dir, name):
    return np.load(golden_dir/f"glsl_{name}.npz")


def restated(case: C.Case, rows=None):
    u = case.uniforms
    f = G.varyings(u, case.Wr, case.Hr, rows=rows)
    return G.SCENES[case.scene](u, f, case.tex)


@pytest.mark.parametrize("name", list(CASES))
def test_restatement_equals_reference_text(golden_dir, name):
    case = CASES[name]
    gold = load(golden_dir, name)
    assert str(gold["inputs_sha1"]) == C.digest(case), "case inputs drifted from the committed golden"
    want = gold["screen_f32"]
    got = restated(case)
    assert got.shape == want.shape
    both_nan = np.isnan(got) & np.isnan(want)
    err = np.where(both_nan, 0.0, np.abs(got.astype(np.float64) - want.astype(np.float64)))
    err = np.nan_to_num(err, nan=np.inf)
    if name in LOOSE:
        assert (err > TOL).mean() <= LOOSE[name], f"{name}: {(err > TOL).mean():.4f} of channels differ"
    else:
        assert err.max() <= TOL, f"{name}: max |Δ| {err.max():.3e}"


@pytest.mark.parametrize("name", [n for n, c in CASES.items() if c.final])
def test_final_pass_of_scene_goldens(golden_dir, name):
    case = CASES[name]
    gold = load(golden_dir, name)
    screen_u8 = G.to_unorm8(restated(case))
    assert np.array_equal(screen_u8, gold["screen_u8"])
    for k in case.final:
        fin = G.final_pass(screen_u8, case.W, case.H, k)
        assert np.abs(fin - gold[f"final{k}_f32"]).max() <= TOL
        assert np.array_equal(G.to_unorm8(fin), gold[f"final{k}_u8"])


def test_final_glsl_geometries(golden_dir):
    """fragment/final.glsl:3-33 over (ssaa, subsample) incl. the reference default (1, 2) and non-integer ssaa"""
    gold = np.load(golden_dir/"glsl_final.npz")
    W, H = 40, 24
    for ssaa, k in C.FINAL_GEOMETRIES:
        fin = G.final_pass(C.final_screen(W, H, ssaa), W, H, k)
        assert np.abs(fin - gold[f"s{ssaa}_k{k}_f32"]).max() <= TOL, (ssaa, k)


def test_4k_bands_of_the_benchmarked_geometry(golden_dir):
    """3840×2160, ssaa 2, 1920×1080 background (BASELINE configs[2]): four 8-row bands of the 7680×4320 target"""
    case = C.band_case()
    gold = load(golden_dir, case.name)
    assert str(gold["inputs_sha1"]) == C.digest(case)
    got = restated(case, rows=case.rows)
    assert np.abs(got[:, case.cols] - gold["screen_f32"]).max() <= TOL
    full = np.zeros((case.Hr, case.Wr, 4), np.uint8)
    full[case.rows] = G.to_unorm8(got)
    out_rows = sorted({r//2 for r in case.rows})
    # The fused kernels replace final.glsl by an exact 2×2 box of texel centres. In float32 the
    # tap coordinates of final.glsl:17-28 miss the centres by ~1e-4 texel at this size, so the text's result is
    # within 3e-4 (0.08 LSB) of the box, not equal to it: fused 8-bit output may differ by 1 LSB on rounding ties
    box = full[case.rows].reshape(len(out_rows), 2, case.W, 2, 4).astype(np.float32)
    fin = G.final_pass(full, case.W, case.H, 2)[out_rows]
    assert np.array_equal(G.to_unorm8(fin), gold["final2_u8"])
    assert np.abs(fin[:, case.cols] - gold["final2_f32"]).max() <= TOL
    assert np.abs(box.mean(axis=(1, 3))[..., :3]/255 - fin).max() < 3e-4


# ---------------------------------------------------------------------------------------------- #
# live capture (build container only)

needs_reference = pytest.mark.skipif(not ref_loader.available(), reason="needs /root/reference")


@pytest.fixture(scope="module")
def live():
    import subprocess, sys, json, tempfile
    from pathlib import Path
    # a fresh interpreter: the reference's `shaderflow` must not meet the alias other tests install
    code = ("import sys, json, hashlib; sys.path.insert(0, %r)\n"
            "from oracle import ref_scene\n"
            "out = {}\n"
            "for s in %r:\n"
            "    cap = ref_scene.capture(s)\n"
            "    out[s] = {p: dict(fragment=v['fragment'], vertex=v['vertex']) for p, v in cap['programs'].items()}\n"
            "json.dump(out, open(sys.argv[1], 'w'))\n") % (str(Path(__file__).resolve().parents[1]),
                                                           sorted({c.ref_scene for c in CASES.values()}))
    with tempfile.TemporaryDirectory() as tmp:
        path = Path(tmp)/"cap.json"
        subprocess.run([sys.executable, "-c", code, str(path)], check=True, capture_output=True)
        return json.loads(path.read_text())


@needs_reference
def test_goldens_record_the_live_reference_text(golden_dir, live):
    for name, case in CASES.items():
        gold = load(golden_dir, name)
        src = live[case.ref_scene][case.program]
        assert X.text_digest(src["fragment"]) == str(gold["fragment_sha1"]), name
        assert X.text_digest(src["vertex"]) == str(gold["vertex_sha1"]), name


@needs_reference
def test_live_text_reexecutes_to_the_golden(golden_dir, live):
    for name in ("visualizer", "default_equirect", "life_simulation_f6"):
        case = CASES[name]
        src = live[case.ref_scene][case.program]
        prog = X.Program(src["vertex"], src["fragment"])
        got = prog.render(C.exec_uniforms(case.uniforms), C.exec_samplers(case.tex), case.Wr, case.Hr)
        assert np.array_equal(got, load(golden_dir, name)["screen_f32"], equal_nan=True), name


@needs_reference
def test_the_reference_text_contains_the_files_it_should(live):
    """The captured text really is header + shaderflow.glsl + camera.glsl + the example's fragment"""
    ref = ref_loader.REFERENCE
    frag = live["Visualizer"]["iScreen"]["fragment"]
    for rel in ("shaderflow/resources/shaders/include/shaderflow.glsl", "shaderflow/resources/shaders/include/camera.glsl",
                "examples/basic/shaders/visualizer.frag"):
        assert (ref/rel).read_text() in frag, rel
    assert frag.startswith("#version 330\n#define FRAGMENT")
    assert (ref/"shaderflow/resources/shaders/vertex/default.glsl").read_text() in live["Visualizer"]["iScreen"]["vertex"]
    assert (ref/"shaderflow/resources/shaders/fragment/final.glsl").read_text() in live["Visualizer"]["iFinal"]["fragment"]

# ---------------------------------------------------------------------------------------------- #
# the evaluator on its own

def run(body: str, lanes=4, inputs=None, header=""):
    src = f"#version 330\nout vec4 fragColor;\nin float x;\n{header}\nvoid main() {{\n{body}\n}}\n"
    m = X.Machine(src)
    inp = dict(x=np.arange(lanes, dtype=np.float32))
    inp.update(inputs or {})
    return np.broadcast_to(m.run(lanes, inp)["fragColor"].a, (lanes, 4))


def test_exec_float_loop_counter_is_strict_float32():
    # visualizer.frag:26 — 8 steps of TAU/8 do not reach TAU in float32: 9 iterations
    out = run("const float TAU = 6.2831853071795864; float n = 0; float directions = 8;"
              "for (float a=0; a<TAU; a+=TAU/directions) n += 1; fragColor = vec4(n);")
    assert out[0, 0] == 9


def test_exec_divergent_control_flow():
    out = run("""
        float acc = 0; int i;
        for (i = 0; i < 10; i++) {
            if (float(i) > x) break;
            if (i == 1) continue;
            acc += 1;
        }
        fragColor = vec4(acc, i, x < 2 ? 1 : 2, 0);
        if (x > 2.5) { fragColor.a = 7; return; }
        fragColor.a = 3;""")
    assert out[:, 0].tolist() == [1, 1, 2, 3]
    assert out[:, 1].tolist() == [1, 2, 3, 4]
    assert out[:, 2].tolist() == [1, 1, 2, 2]
    assert out[:, 3].tolist() == [3, 3, 3, 7]


def test_exec_functions_overloads_out_params_structs():
    header = """
        struct P { vec2 a; float b; };
        float f(float v) { return v*2; }
        float f(vec2 v) { return v.x + v.y; }
        void g(float v, out float twice, inout float acc) { twice = 2*v; acc += v; }
        P make(float v) { P p; p.a = vec2(v, 1); p.b = v > 1.5 ? 10 : 20; return p; }
        float early(float v) { if (v < 1.5) return -1; return 1; }
    """
    out = run("float t; float acc = 100; g(x, t, acc); P p = make(x);"
              "fragColor = vec4(f(x) + f(vec2(x, 1)), t + acc, p.b + p.a.y, early(x));", header=header)
    x = np.arange(4)
    assert np.array_equal(out[:, 0], 2*x + x + 1)
    assert np.array_equal(out[:, 1], 2*x + 100 + x)
    assert out[:, 2].tolist() == [21, 21, 11, 11]
    assert out[:, 3].tolist() == [-1, -1, 1, 1]


def test_exec_int_semantics_and_implicit_conversions():
    # tetration.frag:50 integer division; visualizer.frag:9 vec3(ints)/int; mandelbrot.frag:28 float(int)/int
    out = run("int it = int(x); int M = 3; float k = it / M; vec3 s = vec3(1, 11, 26) / 255;"
              "fragColor = vec4(k, s.y, float(it)/M, (7 % 4) + int(-1.7));")
    assert out[:, 0].tolist() == [0, 0, 0, 1]
    assert out[0, 1] == np.float32(11)/np.float32(255)
    assert np.allclose(out[:, 2], np.arange(4)/3)
    assert out[0, 3] == 2


def test_exec_matrix_swizzle_switch_arrays():
    header = "const int table[4] = int[4](5, 6, 7, 8);"
    out = run("""
        mat2 m = mat2(1, 2, 3, 4);            // columns (1,2) and (3,4)
        vec2 v = m * vec2(1, 10);             // (1+30, 2+40)
        vec4 c = vec4(0); c.zx = v; c.w = table[int(x)];
        switch (int(x)) { case 0: c.y = 1; break; case 1: c.y = 2; case 2: c.y += 5; break; default: c.y = 9; }
        fragColor = c;""", header=header)
    assert out[0].tolist() == [42, 1, 31, 5]
    assert out[:, 1].tolist() == [1, 7, 5, 9]
    assert out[:, 3].tolist() == [5, 6, 7, 8]


def test_exec_preprocessor():
    header = """
        #define TWICE(a) ((a)*2)
        #define GLUE(name) name##Value
        #ifndef NOPE
        #define PICK 3
        #else
        #define PICK 4
        #endif
        uniform float iDelta;
        #define iDelta (5.0)
        float fooValue = 11;
    """
    out = run("fragColor = vec4(TWICE(x + 1), GLUE(foo), PICK, iDelta);", header=header, inputs=dict(iDelta=1.0))
    assert out[:, 0].tolist() == [2, 4, 6, 8]
    assert out[0, 1:].tolist() == [11, 3, 5]


def test_exec_declaration_initialiser_sees_the_outer_name():
    # camera.glsl:104 `vec2 gluv = gluv - ...;` inside a block reads the varying
    out = run("float y = x; { float y = y + 1; fragColor = vec4(y); }")
    assert out[:, 0].tolist() == [1, 2, 3, 4]

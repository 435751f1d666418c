"""The PRODUCTION per-fragment code on the CPU: the scene functions of csrc/scenes.cuh (with glsl.cuh and sampler.cuh —
what screen_kernel / frame_kernel call for every fragment) compiled for the host with g++ over tests/aot_host_shim.h and
run on the cases of oracle/glsl_cases.py against tests/golden/glsl_*.npz — the reference's shader text, executed.

tests/test_gpu_golden.py makes the same comparison through the C ABI on the B200 (1e-3 per channel, north_star's gate);
here, without a GPU and with libm instead of libdevice, the transliteration itself is held to the reference text much
tighter: 1e-5 on the continuous scenes. -ffp-contract=off: one rounding per operation, as nvcc is told for these files
(no FMA contraction differences are tolerated by the goldens' 1e-6 pin either).

Covered here: the generic path of all 16 reference scenes (P.fast = 0), the default fast variants of the two fractals, and
the generic visualizer on four bands of the benchmarked 4K geometry. Also the table-driven visualizer
(scene_visualizer_fast; its constant-memory tap table is filled by the harness with the launcher's float loops). Not
here: the separable visualizer kernels (visualizer_rows.cu, visualizer_tiled.cuh) — those are held to this generic
path and to the goldens on the device
(tests/test_gpu_render.py, tests/test_gpu_golden.py)."""
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import glsl_cases as C
from shaderflow_b200 import _native as N
from tests.helpers import native_uniforms

ROOT = Path(__file__).resolve().parents[1]
CUDA_INCLUDE = Path("/usr/local/cuda/include")
pytestmark = pytest.mark.skipif(shutil.which("g++") is None or not (CUDA_INCLUDE/"cuda_runtime.h").exists(), reason="needs g++ and the CUDA headers")

from oracle.cpu_compiled import SOURCE, SHIM_DIR      # the harness also times the compiled port for bench.py


CASES = {c.name: c for c in C.small_cases()}
# continuous scenes are held everywhere; the others branch on thresholds and may flip a fragment on an ulp
CONTINUOUS = {"default", "default_stereo", "default_equirect", "default_rotated", "shadertoy", "dynamics", "audio",
              "multishader_child", "multishader", "multipass_layer0", "multipass_layer1", "motionblur_layer0", "motionblur_layer1"}


@pytest.fixture(scope="module")
def binary(tmp_path_factory):
    work = tmp_path_factory.mktemp("aot")
    (work/"scenes_host.cpp").write_text(SOURCE)
    build = subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-w", "-I", str(SHIM_DIR), "-I", str(ROOT/"include"),
                            "-I", str(CUDA_INCLUDE), "-I", str(ROOT/"shaderflow_b200"/"csrc"), str(work/"scenes_host.cpp"),
                            "-o", str(work/"scenes_host")], capture_output=True, text=True)
    assert build.returncode == 0, build.stderr[-3000:]
    return work/"scenes_host"


def run_scene(binary, tmp_path, case, fast: int = 0, rows=None) -> np.ndarray:
    sid = N.scene_lookup(case.scene)
    info = N.scene_info(sid)
    (tmp_path/"uniforms.bin").write_bytes(bytes(native_uniforms(case.uniforms, info)))
    textures = case.tex                                                        # keyed like the scene's sampler table
    args = []
    for k, name in enumerate(info["samplers"]):
        t = textures.get(name)
        data = np.zeros((1, 1, 4), np.uint8) if t is None else t.data          # declared but never read by this pass
        comps = data.shape[2]
        if comps == 3:                                                         # stored padded to 4 components, alpha reads 1
            data = np.concatenate([data, np.full(data.shape[:-1] + (1,), 255 if data.dtype == np.uint8 else 1.0, data.dtype)], -1)
        (tmp_path/f"texture{k}.bin").write_bytes(np.ascontiguousarray(data).tobytes())
        linear, rx, ry = (True, False, False) if t is None else (t.linear, t.repeat_x, t.repeat_y)
        args += [str(tmp_path/f"texture{k}.bin"), data.shape[1], data.shape[0], data.shape[2], comps,
                 N.DTYPE_U8 if data.dtype == np.uint8 else N.DTYPE_F32, int(linear), int(rx), int(ry)]
    import os
    env = dict(os.environ)
    if rows is not None:
        env["SFB_ROWS"] = ",".join(str(int(r)) for r in rows)
    done = subprocess.run([str(binary), str(tmp_path/"uniforms.bin"), str(sid), str(case.Wr), str(case.Hr), str(tmp_path/"out.bin"), str(fast),
                           *map(str, args)], capture_output=True, text=True, env=env)
    assert done.returncode == 0, (done.returncode, done.stderr[-500:])
    return np.fromfile(tmp_path/"out.bin", np.float32).reshape(case.Hr if rows is None else len(rows), case.Wr, 4)


@pytest.mark.parametrize("name", list(CASES))
def test_production_scene_functions_reproduce_the_reference_text(binary, tmp_path, golden_dir, name):
    case = CASES[name]
    want = np.load(golden_dir/f"glsl_{name}.npz")["screen_f32"]
    got = run_scene(binary, tmp_path, case)
    if case.rows is not None:
        got = got[case.rows]
    if case.scene == "life_simulation":
        got, want = got[..., :1], want[..., :1]                                # the program writes .r (a one-component target)
    error = np.abs(got - want)/np.maximum(1.0, np.abs(want))
    error = np.nan_to_num(error, nan=0.0)
    if name in CONTINUOUS:
        assert error.max() <= 5e-5, (name, float(error.max()))                 # (asin / atan of the equirectangular camera: 1.7e-5)
    else:
        least = 0.98 if case.scene == "tetration" else 0.995                   # 67 complex powers amplify an ulp of pow / exp
        assert (error <= 1e-5).mean() >= least, (name, float((error <= 1e-5).mean()), float(error.max()))
    assert np.median(error) <= 1e-6


def test_generic_visualizer_at_the_benchmarked_geometry(binary, tmp_path, golden_dir):
    """BASELINE configs[2]: 3840 x 2160, ssaa 2 → a 7680 x 4320 target over the 1920 x 1080 background. Four 8-row bands
    of it (the golden's), every fragment of those rows, through the generic visualizer scene function on the host"""
    case = C.band_case()
    gold = np.load(golden_dir/"glsl_visualizer_4k_bands.npz")
    want = gold["screen_f32"]
    got = run_scene(binary, tmp_path, case, rows=case.rows)
    got = got if case.cols is None else got[:, case.cols]
    error = np.abs(got - want)/np.maximum(1.0, np.abs(want))
    assert (error <= 1e-5).mean() >= 0.999, (float((error <= 1e-5).mean()), float(error.max()))
    assert np.median(error) <= 1e-6


@pytest.mark.parametrize("name", ["mandelbrot", "tetration"])
def test_default_fast_variants_of_the_fractals(binary, tmp_path, golden_dir, name):
    """What the launcher runs by default for the two fractals (render.cu: P.fast = 1 unless SFB_RENDER_LITERAL — the
    closed-form interior tests of the Mandelbrot set, one exp of the combined exponent per tetration step): the
    variants BASELINE configs[3] is measured with, against the reference text's goldens"""
    case = CASES[name]
    want = np.load(golden_dir/f"glsl_{name}.npz")["screen_f32"]
    literal, fast = run_scene(binary, tmp_path, case, fast=0), run_scene(binary, tmp_path, case, fast=1)
    for got, least in ((fast, 0.97 if name == "tetration" else 0.995),):
        error = np.abs(got - want)/np.maximum(1.0, np.abs(want))
        assert (error <= 1e-4).mean() >= least, (name, float((error <= 1e-4).mean()), float(error.max()))
        assert np.median(error) <= 1e-6
    assert (np.abs(fast - literal) <= 1e-4).mean() >= (0.97 if name == "tetration" else 0.999)


@pytest.mark.parametrize("name", [n for n, c in CASES.items() if c.scene == "visualizer"])
def test_visualizer_table_driven_path_reproduces_the_reference_text(binary, tmp_path, golden_dir, name):
    """scene_visualizer_fast — the hoisted tap table, texel arithmetic in byte units, the quad cache: what the
    one-thread-per-fragment kernels run by default for the Visualizer — against the same goldens and the literal path"""
    case = CASES[name]
    want = np.load(golden_dir/f"glsl_{name}.npz")["screen_f32"]
    literal, fast = run_scene(binary, tmp_path, case, fast=0), run_scene(binary, tmp_path, case, fast=1)
    error = np.abs(fast - want)/np.maximum(1.0, np.abs(want))
    assert (error <= 1e-5).mean() >= 0.995, (name, float((error <= 1e-5).mean()), float(error.max()))
    assert np.abs(fast - literal).max() <= 1e-5

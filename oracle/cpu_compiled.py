"""The ahead-of-time scene functions (csrc/scenes.cuh with glsl.cuh / sampler.cuh) compiled for the HOST — test and bench
infrastructure only, never on the product path.

    SOURCE / SHIM_DIR      the g++ harness tests/test_aot_host.py checks against the reference-text goldens
    visualizer_sample()    the same binary, -O2, timed on every host core over row bands of one 4K 2xSSAA frame: a
                           COMPILED CPU figure beside the numpy port of oracle/cpu_bench.py (bench.py --impl reference).
                           llvmpipe, which the reference would run on here, compiles GLSL to native code too — a numpy
                           baseline alone would flatter the GPU/CPU ratio
"""
from __future__ import annotations

import os
import shutil
import subprocess
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
SHIM_DIR = Path(__file__).resolve().parent
CUDA_INCLUDE = Path("/usr/local/cuda/include")

SOURCE = r'''
#include "aot_host_shim.h"
#include "sfb200.h"
#include "render_params.h"
#include "scenes.cuh"

template <int SCENE> static void run(const RenderParams& P, FILE* out) {
    const char* rows = getenv("SFB_ROWS");                  // "j0,j1,…": only these fragment rows (the 4K bands)
    for (int j = 0; j < P.Hr; j++) {
        if (rows) {
            bool wanted = false;
            for (const char* p = rows; *p; ) { if (atoi(p) == j) wanted = true; while (*p && *p != ',') p++; if (*p) p++; }
            if (!wanted) continue;
        }
        for (int i = 0; i < P.Wr; i++) {
            const glsl::vec4 c = glsl::shade<SCENE, false>(P, glsl::make_frag(P, i, j));
            const float v[4] = {c.x, c.y, c.z, c.w};
            fwrite(v, sizeof(float), 4, out);
        }
    }
}

int main(int argc, char** argv) {
    static RenderParams P;
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(&P.u, sizeof(P.u), 1, f) != 1) return 2;
    fclose(f);
    const int scene = atoi(argv[2]);
    P.Wr = atoi(argv[3]); P.Hr = atoi(argv[4]); P.W = int(P.u.iResolution[0]); P.H = int(P.u.iResolution[1]);
    P.inv_Wr = 1.0/double(P.Wr); P.inv_Hr = 1.0/double(P.Hr);
    P.fast = atoi(argv[6]);
    for (int k = 0; 7 + 9*k < argc; k++) {                 // texture k: file w h padded comps dtype filter rx ry
        char** a = argv + 7 + 9*k;
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
    {   // the tap table the launcher keeps in constant memory (render_kernels.cuh build_blur_table), same float loops
        const volatile float TAU_F = 6.2831853071795864f, directions = 8.0f, quality = 10.0f;
        int n = 0;
        for (volatile float angle = 0.0f; angle < TAU_F; angle = angle + TAU_F/directions) {
            const float c = cosf(angle), s = sinf(angle);
            for (volatile float walk = 1.0f/quality; walk <= 1.001f; walk = walk + 1.0f/quality)
                if (n < 90) glsl::c_blur.tap[n++] = make_float2(c*walk, s*walk);
        }
        if (n != 90) return 5;
        glsl::c_blur.tap[90] = glsl::c_blur.tap[91] = make_float2(0.0f, 0.0f);
    }
    FILE* out = fopen(argv[5], "wb");
    switch (scene) {
#define CASE(ID) case ID: run<ID>(P, out); break;
        CASE(SFB_SCENE_DEFAULT) CASE(SFB_SCENE_SHADERTOY) CASE(SFB_SCENE_VISUALIZER) CASE(SFB_SCENE_BARS) CASE(SFB_SCENE_WAVEFORM)
        CASE(SFB_SCENE_MANDELBROT) CASE(SFB_SCENE_TETRATION) CASE(SFB_SCENE_RAYMARCH) CASE(SFB_SCENE_MULTISHADER_CHILD)
        CASE(SFB_SCENE_MULTISHADER) CASE(SFB_SCENE_MULTIPASS) CASE(SFB_SCENE_MOTIONBLUR) CASE(SFB_SCENE_DYNAMICS) CASE(SFB_SCENE_AUDIO)
        CASE(SFB_SCENE_LIFE_SIMULATION) CASE(SFB_SCENE_LIFE_VISUALS) CASE(SFB_SCENE_PIANO)
        default: return 4;
    }
    fclose(out);
    return 0;
}
'''


def visualizer_sample(width: int = 3840, height: int = 2160, ssaa: int = 2, rows_per_band: int = 8, bands_per_worker: int = 4) -> dict:
    """Shades row bands of the (width*ssaa) x (height*ssaa) Visualizer target with the host-compiled generic scene
    function, one process per host core, and extrapolates to a frame → dict(frames_per_s, cores, sample)"""
    from oracle import cpu_bench, glsl_cases as C
    from shaderflow_b200 import _native as N
    from tests.helpers import native_uniforms
    if shutil.which("g++") is None or not (CUDA_INCLUDE/"cuda_runtime.h").exists():
        raise RuntimeError("needs g++ and the CUDA headers")
    cores = cpu_bench.usable_cores()
    Wr, Hr = int(width*ssaa), int(height*ssaa)
    case = C.band_case()
    with tempfile.TemporaryDirectory() as tmp:
        tmp = Path(tmp)
        (tmp/"scenes_host.cpp").write_text(SOURCE)
        build = subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-w", "-I", str(SHIM_DIR), "-I", str(ROOT/"include"),
                                "-I", str(CUDA_INCLUDE), "-I", str(ROOT/"shaderflow_b200"/"csrc"), str(tmp/"scenes_host.cpp"),
                                "-o", str(tmp/"scenes_host")], capture_output=True, text=True)
        if build.returncode != 0:
            raise RuntimeError(build.stderr[-500:])
        sid = N.scene_lookup("visualizer")
        info = N.scene_info(sid)
        u = case.uniforms
        u = type(u)(**{**{f: getattr(u, f) for f in u.__dataclass_fields__}, "iResolution": (width, height), "iWantAspect": width/height, "iSSAA": float(ssaa)})
        (tmp/"uniforms.bin").write_bytes(bytes(native_uniforms(u, info)))
        args = []
        for k, name in enumerate(info["samplers"]):
            t = case.tex[name]
            data = t.data
            comps = data.shape[2]
            if comps == 3:
                data = np.concatenate([data, np.full(data.shape[:-1] + (1,), 255, data.dtype)], -1)
            (tmp/f"texture{k}.bin").write_bytes(np.ascontiguousarray(data).tobytes())
            args += [str(tmp/f"texture{k}.bin"), data.shape[1], data.shape[0], data.shape[2], comps,
                     N.DTYPE_U8 if data.dtype == np.uint8 else N.DTYPE_F32, int(t.linear), int(t.repeat_x), int(t.repeat_y)]
        bands = cores*bands_per_worker
        starts = np.linspace(0, Hr - rows_per_band, bands).astype(int)

        def work(worker: int) -> int:
            rows = [int(s) + r for s in starts[worker::cores] for r in range(rows_per_band)]
            env = dict(os.environ, SFB_ROWS=",".join(map(str, rows)))
            done = subprocess.run([str(tmp/"scenes_host"), str(tmp/"uniforms.bin"), str(sid), str(Wr), str(Hr), str(tmp/f"out{worker}.bin"), "0",
                                   *map(str, args)], capture_output=True, text=True, env=env)
            if done.returncode != 0:
                raise RuntimeError(f"host scene binary failed ({done.returncode})")
            return len(rows)
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=cores) as pool:
            rows_done = sum(pool.map(work, range(cores)))
        took = time.perf_counter() - t0
    fragments = rows_done*Wr
    per_frame = took*(Wr*Hr)/fragments
    return dict(frames_per_s=1.0/per_frame, cores=cores,
                sample=f"{bands} bands x {rows_per_band} rows of the {Wr}x{Hr} iScreen target ({fragments} fragments, {took:.1f} s on {cores} "
                       "processes), csrc/scenes.cuh's generic visualizer function compiled for the host (g++ -O2, scalar); shading only")

"""
Generates tests/golden/glsl_*.npz: the REFERENCE's shader text, executed.

    python tests/golden/make_golden_glsl.py [case ...]        (build container only: needs /root/reference)

For every case of `oracle/glsl_cases.py` the reference's own Python builds the example scene headlessly
(`oracle/ref_scene.py`: scene graph, module order and `_build_shader` metaprogramming are the reference's
unmodified code; only the GL objects are recording stubs), the assembled vertex + fragment GLSL that it hands
to `opengl.program()` is executed by the mechanical evaluator `oracle/glsl_exec.py` on the case's inputs, and
the fragment colours before the colour store are committed. `final` entries run the reference's
`fragment/final.glsl` text on the RGBA8-quantised screen. The script also checks, against the captured
reference scene, what the cases assume about it: sampler names, texture sampling state, default uniform values.

Each file holds: screen_f32 (rows, cols, 4) float32 [column-strided for the 4K bands], screen_u8 where a final
pass follows, final{k}_f32 / final{k}_u8, the sha1 of the fragment text and of the case inputs.
"""
from __future__ import annotations

import hashlib
import sys
import time
from dataclasses import fields
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import glsl_cases as C          # noqa: E402
from oracle import glsl_exec as X           # noqa: E402
from oracle import glsl_np as G             # noqa: E402
from oracle import ref_scene                # noqa: E402

OUT = Path(__file__).parent
_captures: dict = {}


def captured(ref_scene_name: str) -> dict:
    if ref_scene_name not in _captures:
        _captures[ref_scene_name] = ref_scene.capture(ref_scene_name, background=G.synthetic_background(16, 9))
    return _captures[ref_scene_name]


def check_against_reference(case: C.Case, cap: dict) -> None:
    """The case's assumptions vs what the reference scene declares"""
    fragment = cap["programs"][case.program]["fragment"]
    declared = X.Machine(fragment).inputs
    for key, tex in case.tex.items():
        name = C.sampler_name(key)
        assert declared.get(name) == ("uniform", "sampler2D"), f"{case.name}: sampler {name} not declared by the reference"
        module = key if key in cap["textures"] else key.rstrip("0123456789x").rstrip("x")
        module = next(m for m in cap["textures"] if name.startswith(m))
        state = cap["textures"][module]
        if module != "iScreen" or case.ref_scene in ("Multipass", "MotionBlur"):
            assert state["filter"] == ("linear" if tex.linear else "nearest"), (case.name, key, state)
            assert (state["repeat_x"], state["repeat_y"]) == (tex.repeat_x, tex.repeat_y), (case.name, key, state)
    u = C.exec_uniforms(case.uniforms)
    for name in u:
        assert name in declared, f"{case.name}: uniform {name} not declared by the reference"
    defaults = G.Uniforms()
    for f in fields(defaults):
        if f.name.startswith("iCamera") and f.name in cap["uniforms"]:
            ref = np.asarray(cap["uniforms"][f.name][1], np.float64)
            assert np.allclose(ref, np.asarray(getattr(defaults, f.name), np.float64)), (f.name, ref)


def run_case(case: C.Case) -> dict:
    cap = captured(case.ref_scene)
    check_against_reference(case, cap)
    src = cap["programs"][case.program]
    prog = X.Program(src["vertex"], src["fragment"])
    t0 = time.time()
    screen = prog.render(C.exec_uniforms(case.uniforms), C.exec_samplers(case.tex), case.Wr, case.Hr, rows=case.rows)
    out = dict(fragment_sha1=X.text_digest(src["fragment"]),
               vertex_sha1=X.text_digest(src["vertex"]), inputs_sha1=C.digest(case))
    out["screen_f32"] = screen if case.cols is None else np.ascontiguousarray(screen[:, case.cols])
    if case.final:
        fin = cap["programs"]["iFinal"]
        fprog = X.Program(fin["vertex"], fin["fragment"])
        u8 = G.to_unorm8(screen)
        if case.rows is None:
            full = u8
        else:
            full = np.zeros((case.Hr, case.Wr, 4), np.uint8)
            full[case.rows] = u8
        for k in case.final:
            # output rows whose taps lie inside the evaluated rows (ssaa·k ∈ {2·2}: rows 2i, 2i+1)
            out_rows = None if case.rows is None else sorted({r//2 for r in case.rows})
            scr = G.Texture(full, linear=True, repeat_x=False, repeat_y=False)
            f32 = fprog.render(dict(iResolution=(case.W, case.H), iSubsample=k), {"iScreen0x0": scr},
                               case.W, case.H, rows=out_rows)
            out[f"final{k}_f32"] = f32[..., :3] if case.cols is None else np.ascontiguousarray(f32[:, case.cols, :3])
            out[f"final{k}_u8"] = G.to_unorm8(f32[..., :3])
        if case.rows is None:
            out["screen_u8"] = u8
    print(f"{case.name:24s} {screen.shape} in {time.time() - t0:.1f} s")
    return out


def run_final_cases() -> dict:
    """final.glsl alone over (ssaa, subsample) geometries incl. the reference's default (1, 2)"""
    cap = captured("Basic")
    fin = cap["programs"]["iFinal"]
    fprog = X.Program(fin["vertex"], fin["fragment"])
    W, H = 40, 24
    out = dict(fragment_sha1=X.text_digest(fin["fragment"]))
    for ssaa, k in C.FINAL_GEOMETRIES:
        scr = G.Texture(C.final_screen(W, H, ssaa), linear=True, repeat_x=False, repeat_y=False)
        f32 = fprog.render(dict(iResolution=(W, H), iSubsample=k), {"iScreen0x0": scr}, W, H)
        out[f"s{ssaa}_k{k}_f32"] = f32[..., :3]
        assert np.all(f32[..., 3] == 1.0)
    print("final.glsl geometries done")
    return out


def main(argv):
    cases = C.small_cases() + [C.band_case()]
    wanted = set(argv)
    for case in cases:
        if wanted and case.name not in wanted:
            continue
        np.savez_compressed(OUT/f"glsl_{case.name}.npz", **run_case(case))
    if not wanted or "final" in wanted:
        np.savez_compressed(OUT/"glsl_final.npz", **run_final_cases())


if __name__ == "__main__":
    main(sys.argv[1:])

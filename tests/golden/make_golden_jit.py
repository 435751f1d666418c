"""Golden for the CUDA implementation of ShaderFlow's GLSL std-lib (csrc/jit/shaderflow_rt.cuh), build container only:

    python tests/golden/make_golden_jit.py

tests/shaders/stdlib.frag (this repository's text: one probe function per group of the API) is placed where a user's
fragment goes in the text the REFERENCE assembles for the GL driver — its header, `include/shaderflow.glsl` and
`include/camera.glsl` come from /root/reference through oracle/ref_scene.py's capture of the Visualizer scene
(shaderflow/shader.py:190-239) — and that text is executed by the mechanical evaluator oracle/glsl_exec.py, vertex stage
included. The results (floats before any store) go to tests/golden/jit_stdlib.npz; tests/test_gpu_jit.py compiles the
same stdlib.frag against the CUDA std-lib at run time and compares."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import glsl_cases as C          # noqa: E402
from oracle import glsl_exec as X           # noqa: E402
from oracle import ref_scene                # noqa: E402
from tests import jit_cases as J            # noqa: E402

MARKER = "// Metaprogramming (Content)"


def main():
    cap = ref_scene.capture("Visualizer")
    src = cap["programs"]["iScreen"]
    head = src["fragment"][:src["fragment"].index(MARKER)]
    fragment = head + MARKER + "\n\n" + (J.SHADERS/"stdlib.frag").read_text()
    program = X.Program(src["vertex"], fragment)
    out = dict(reference_header_sha1=X.text_digest(head), fragment_sha1=X.text_digest((J.SHADERS/"stdlib.frag").read_text()))
    samplers = C.exec_samplers(J.stdlib_textures())
    for c, camera in enumerate(J.STDLIB_CAMERAS):
        for probe in range(J.STDLIB_PROBES if c == 0 else 1):
            u = C.exec_uniforms(J.uniforms(**camera))
            u["iProbe"] = probe
            declared = program.fragment.inputs
            u = {k: v for k, v in u.items() if k in declared or k in program.vertex.inputs}
            color = program.render(u, samplers, J.W, J.H)
            out[f"camera{c}_probe{probe}"] = color.astype(np.float32)
            print(c, probe, color.reshape(-1, 4).mean(0))
    np.savez_compressed(Path(__file__).parent/"jit_stdlib.npz", **out)


if __name__ == "__main__":
    main()

"""Run-time GLSL → CUDA translation for fragment shaders the backend has no ahead-of-time kernel for.

    frontend.py   GLSL 3.30 tokenizer / preprocessor / parser
    cuda.py       AST → `struct Shader : g::ShaderBase` (CUDA C++)
    program()     the whole NVRTC translation unit + its in-memory headers (csrc/jit/*.cuh, csrc/render_params.h,
                  include/sfb200.h)

`ShaderProgram.compile` (shader.py) calls `build()` when registry.py does not recognise the fragment; the reference does
the equivalent by handing the assembled text to the GL driver (shaderflow/shader.py:313-349)."""
from __future__ import annotations

from pathlib import Path

from shaderflow_b200.glsl.cuda import Translation, TranslationError, translate

_CSRC = Path(__file__).resolve().parent.parent/"csrc"
_INCLUDE = Path(__file__).resolve().parent.parent.parent/"include"

# what NVRTC has no host headers for: the two the C ABI header includes
_STDINT = """#pragma once
typedef signed char int8_t; typedef unsigned char uint8_t; typedef short int16_t; typedef unsigned short uint16_t;
typedef int int32_t; typedef unsigned int uint32_t; typedef long long int64_t; typedef unsigned long long uint64_t;
"""
_STDDEF = "#pragma once\n"


def headers() -> dict[str, str]:
    return {
        "stdint.h": _STDINT, "stddef.h": _STDDEF,
        "sfb200.h": (_INCLUDE/"sfb200.h").read_text(),
        "render_params.h": (_CSRC/"render_params.h").read_text(),
        "glsl_rt.cuh": (_CSRC/"jit"/"glsl_rt.cuh").read_text(),
        "shaderflow_rt.cuh": (_CSRC/"jit"/"shaderflow_rt.cuh").read_text(),
        "jit_kernels.cuh": (_CSRC/"jit"/"jit_kernels.cuh").read_text(),
    }


def program(translation: Translation) -> str:
    """The NVRTC translation unit of one fragment program"""
    return ('#include "sfb200.h"\n#include "render_params.h"\n#include "glsl_rt.cuh"\n#include "shaderflow_rt.cuh"\n'
            + translation.source + '#include "jit_kernels.cuh"\n')


def build(fragment: str, header: str = "", defines: dict[str, str] | None = None, fmad: bool = False):
    """GLSL → (SASS image, Translation, compile log); raises TranslationError / _native.CompileError"""
    from shaderflow_b200 import _native as N
    translation = translate(fragment, header, defines)
    image, log = N.jit_compile(program(translation), headers(), N.JIT_FMAD if fmad else 0)
    return image, translation, log

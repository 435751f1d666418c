"""ShaderProgram — one fullscreen fragment pass rendering into its own ShaderTexture.
API mirror of shaderflow/shader.py:98-425. Where the reference assembles GLSL and asks the GL driver to
compile it (shader.py:190-239,313-349), `compile()` recognises the fragment (registry.py) and selects
the ahead-of-time CUDA kernel; `render()` packs the module pipeline into one `sfb_uniforms` block and
launches it (sfb_render_screen / sfb_render_final)."""
from __future__ import annotations

from pathlib import Path
from typing import Any, Iterable, Optional, Union

import ctypes as C

import numpy as np
from attrs import Factory, define

from shaderflow_b200 import _native as N
from shaderflow_b200 import logger, registry
from shaderflow_b200.message import ShaderMessage
from shaderflow_b200.module import ShaderModule
from shaderflow_b200.texture import ShaderTexture, TextureBox
from shaderflow_b200.variable import FlatVariable, InVariable, OutVariable, ShaderVariable

_FIXED_FIELDS = {name for name, _ in N.Uniforms._fields_} - {"extra"}
_IMAGES: dict = {}           # SASS images of run-time compiled programs by digest of their CUDA source (per process)
DEFAULT_FRAGMENT = "// sfb200: scene=default"
FINAL_FRAGMENT = "// sfb200: scene=final"


def _field_kinds() -> dict[str, int]:
    """fixed field → 0 float, -1 int, n > 0 float vector of n"""
    kinds = {}
    for name, ctype in N.Uniforms._fields_:
        if name == "extra":
            continue
        length = getattr(ctype, "_length_", 0)
        kinds[name] = length if length else (-1 if issubclass(ctype, N.c_int32) else 0)
    return kinds


_FIELD_KINDS = _field_kinds()


def _scalars(value) -> list:
    """numpy scalar / array / tuple / number → flat list of Python numbers"""
    if isinstance(value, (int, float, bool)):
        return [value]
    if isinstance(value, np.ndarray):
        return value.ravel().tolist()
    if isinstance(value, (tuple, list)):
        return list(value)
    return np.asarray(value, dtype=np.float64).ravel().tolist()


def pack_uniforms(block: N.Uniforms, values: dict[str, Any], extra_names: list[str], extra_types: Optional[list[str]] = None) -> N.Uniforms:
    """name → value table of a pipeline → the POD block; numpy scalars/vectors cast to float32 like
    GL uniform uploads do (ctypes stores the float32 rounding of the Python float). `extra_types` (run-time compiled
    programs): GLSL type per slot — int / uint / bool components travel as their 32 bits, not as float values"""
    kinds = _FIELD_KINDS
    for name, value in values.items():
        kind = kinds.get(name)
        if kind is None or value is None:
            continue
        if kind > 0:
            flat = _scalars(value)
            field = getattr(block, name)
            for i in range(kind):
                field[i] = flat[i]
        elif kind < 0:
            setattr(block, name, int(value))
        else:
            setattr(block, name, float(value))
    for slot, name in enumerate(extra_names):
        value = values.get(name)
        if value is None:
            raise RuntimeError(f"Uniform '{name}' required by the shader is not in the scene's pipeline")
        row = block.extra[slot]
        if extra_types is not None and extra_types[slot].startswith("mat"):
            # column `c` of an n x n matrix given as GL takes it (n*n numbers, column after column; nested rows of
            # columns flatten to the same order)
            kind, _, column = extra_types[slot].partition(":")
            n, c = int(kind[3]), int(column or 0)
            flat = np.asarray(value, dtype=np.float64).ravel()
            if flat.size != n*n:
                raise RuntimeError(f"Uniform '{name}' is a {kind}: {n*n} numbers expected, got {flat.size}")
            for i in range(n):
                row[i] = float(flat[c*n + i])
            continue
        flat = _scalars(value)
        if extra_types is not None and not extra_types[slot].startswith(("float", "vec")):
            bits = np.asarray([int(v) & 0xFFFFFFFF for v in flat[:4]], dtype=np.uint32).view(np.float32)
            C.memmove(row, bits.ctypes.data, bits.nbytes)         # bit patterns: no float conversion (NaN payloads survive)
            continue
        for i in range(min(4, len(flat))):
            row[i] = flat[i]
    return block


@define
class ShaderProgram(ShaderModule):
    version: int = 330
    clear: bool = True
    instances: int = 1
    texture: ShaderTexture = None

    vertex_variables: list = Factory(list)
    fragment_variables: list = Factory(list)
    vertices: list = Factory(list)

    _vertex: Union[Path, str] = ""
    _fragment: Union[Path, str] = ""

    scene_id: Optional[int] = None
    """Index of the ahead-of-time kernel selected by compile() (SFB_SCENE_*)"""
    scene_info: Optional[dict] = None
    filter_flags: int = N.FILTER_EXACT
    """SFB_FILTER_EXACT (parity default) or SFB_FILTER_HARDWARE"""

    def build(self):
        self.texture = ShaderTexture(scene=self.scene, name=self.name, track=True)
        # Declarations kept for introspection; the CUDA kernels get varyings from the rasteriser rule
        self.fragment_variable(OutVariable("vec4", "fragColor"))
        self.vertex_variable(InVariable("vec2", "vertex_position"))
        self.vertex_variable(InVariable("vec2", "vertex_gluv"))
        for name in ("fragCoord", "stxy", "glxy", "stuv", "astuv", "gluv", "agluv"):
            self.traverse_variable(ShaderVariable("vec2", name))
        self.traverse_variable(FlatVariable("int", "instance"))
        for x in (-1, 1):
            for y in (-1, 1):
                self.add_vertice(x=x, y=y, u=x, v=y)
        self._fragment = DEFAULT_FRAGMENT

    # -- declarations (API compatibility) ------------------------------------------------------
    def vertex_variable(self, variable: ShaderVariable) -> None:
        if variable not in self.vertex_variables: self.vertex_variables.append(variable)

    def fragment_variable(self, variable: ShaderVariable) -> None:
        if variable not in self.fragment_variables: self.fragment_variables.append(variable)

    def common_variable(self, variable: ShaderVariable) -> None:
        self.fragment_variable(variable); self.vertex_variable(variable)

    def traverse_variable(self, variable: ShaderVariable) -> None:
        self.fragment_variable(variable.copy(direction="in"))
        self.vertex_variable(variable.copy(direction="out"))

    def add_vertice(self, x: float = 0, y: float = 0, u: float = 0, v: float = 0) -> None:
        self.vertices.extend((x, y, u, v))

    # -- sources -------------------------------------------------------------------------------
    @property
    def vertex(self) -> str:
        return registry.read(self._vertex)

    @vertex.setter
    def vertex(self, value: Union[Path, str]):
        self._vertex = value

    @property
    def fragment(self) -> str:
        return registry.read(self._fragment)

    @fragment.setter
    def fragment(self, value: Union[Path, str]):
        self._fragment = value
        self.scene_id = None

    # -- "compilation" -------------------------------------------------------------------------
    def compile(self) -> "ShaderProgram":
        if self.texture.final:
            self.scene_id, self.scene_info = -1, dict(name="final", extra=[], samplers=[])
            return self
        name = registry.resolve(self._fragment)
        if name is None:
            return self._compile_runtime()
        self.release_runtime()
        self.scene_id = N.scene_lookup(name)
        self.scene_info = N.scene_info(self.scene_id)
        return self

    def _compile_runtime(self) -> "ShaderProgram":
        """A fragment no ahead-of-time kernel exists for: translate the GLSL to CUDA and compile it now, as the
        reference hands any text to the GL driver (shader.py:313-349). The header in front of the user's text is
        what `_build_shader` assembles (shader.py:190-239): every pipeline variable's declaration and the modules'
        defines (texture aliases and accessors); the std-lib (shaderflow.glsl, camera.glsl) comes from
        csrc/jit/shaderflow_rt.cuh. Raises RuntimeError with the translator's / compiler's diagnostics."""
        import hashlib
        import os
        from shaderflow_b200 import glsl
        fragment = self.fragment
        # `#include "file"` lines resolve against the fragment's own directory and include_directories (shader.py:186,231-235)
        roots = [Path(self._fragment).parent] if isinstance(self._fragment, Path) else []
        roots += [Path(d) for d in self.include_directories]
        def include(match, depth=[0]):
            for root in roots:
                if (root/match.group(1)).is_file():
                    depth[0] += 1
                    if depth[0] > 64:
                        raise RuntimeError(f"ShaderProgram '{self.name}': #include nesting too deep ({match.group(1)})")
                    return self._include_regex.sub(include, (root/match.group(1)).read_text())
            raise RuntimeError(f"ShaderProgram '{self.name}': #include \"{match.group(1)}\" not found in {[str(r) for r in roots]}")
        fragment = self._include_regex.sub(include, fragment)
        header, seen = [], set()
        for variable in self.full_pipeline():
            if variable.name not in seen:
                seen.add(variable.name)
                header.append(variable.declaration)
        for module in self.scene.modules:
            header.extend(module.defines() or ())
        # contraction of a*b+c into fma like the ahead-of-time kernels (nvcc's default) and GL drivers; SFB_JIT_FMAD=0
        # compiles with one rounding per operation, the arithmetic the parity tests hold the translator to
        fmad = os.environ.get("SFB_JIT_FMAD", "1") != "0"
        try:
            translation = glsl.translate(fragment, "\n".join(header))
            source = glsl.program(translation)
            key = hashlib.sha1((source + f"|fmad={int(fmad)}|{N.LIBRARY.stat().st_mtime_ns}").encode()).hexdigest()
            if key == self._runtime_key and self._runtime_scene is not None:
                return self                                    # main() compiles every export: same text, same program
            image = _IMAGES.get(key)
            cache = os.environ.get("SFB_JIT_CACHE")            # optional on-disk cache of SASS images (a directory)
            if image is None and cache and (Path(cache)/f"{key}.cubin").is_file():
                image = (Path(cache)/f"{key}.cubin").read_bytes()
            if image is None:
                image, _ = N.jit_compile(source, glsl.headers(), N.JIT_FMAD if fmad else 0)
                if cache:
                    Path(cache).mkdir(parents=True, exist_ok=True)
                    (Path(cache)/f"{key}.cubin").write_bytes(image)
            _IMAGES[key] = image
        except (glsl.TranslationError, N.CompileError) as error:
            raise RuntimeError(logger.error(
                f"ShaderProgram '{self.name}': the fragment shader could not be compiled for the CUDA backend: {error}")) from None
        self.release_runtime()
        if self.scene.cuda is None:
            # a dry scene (no device): the text has been translated and compiled — a syntax / type check — but nothing is loaded
            self.scene_id = N.SCENE_PROGRAM_BASE
        else:
            self.scene_id = self._runtime_scene = self.scene.cuda.program_load(image, len(translation.samplers))
            self._runtime_key = key
        self.scene_info = dict(name=f"runtime:{registry.digest(self.fragment)}", extra=list(translation.extra),
                               extra_types=list(translation.extra_types), samplers=list(translation.samplers),
                               required=len(translation.samplers))
        return self

    _runtime_key: Optional[str] = None
    _runtime_scene: Optional[int] = None
    include_directories: list = Factory(list)
    """Directories `#include "file"` lines of a run-time compiled fragment are looked up in"""
    _include_regex = __import__("re").compile(r'^[ \t]*#include[ \t]+"(.+)"[ \t]*$', __import__("re").MULTILINE)

    def destroy(self) -> None:
        try:
            self.release_runtime()
        except Exception:
            pass                                   # the context may already be gone at interpreter exit

    def release_runtime(self) -> None:
        if self._runtime_scene is not None and self.scene.cuda is not None:
            self.scene.cuda.program_unload(self._runtime_scene)
        self._runtime_scene = self._runtime_key = None

    # -- uniforms ------------------------------------------------------------------------------
    _plan: Any = None
    """(modules seen, scene id, modules whose pipeline yields a name this program's kernel reads, textures whose
    boxes it samples): the reference declares a program's uniforms once, at compile() (shader.py:316-317), so
    WHICH module provides which name is fixed per compilation — only the values change per frame"""

    def gather_planned(self) -> tuple[dict, dict]:
        """gather(full_pipeline()) restricted to the modules that matter to this program's kernel"""
        modules = self.scene.modules
        plan = self._plan
        if plan is None or plan[0] != len(modules) or plan[1] != self.scene_id:
            needed = set(_FIELD_KINDS) | set(self.scene_info["extra"])
            wanted = set(self.scene_info["samplers"])
            providers, textures = [], []
            for module in modules:
                names = {v.name for v in (module.pipeline() or ()) if v.type != "sampler2D"}
                if names & needed:
                    providers.append(module)
                if isinstance(module, ShaderTexture) and (set(module.sampler_names()) & wanted):
                    textures.append(module)
            plan = self._plan = (len(modules), self.scene_id, providers, textures)
        values, samplers = {}, {}
        for module in plan[2]:
            for variable in module.pipeline():
                if variable.value is not None and variable.type != "sampler2D":
                    values[variable.name] = variable.value
        for module in plan[3]:
            samplers.update(module.sampler_names())
        return values, samplers

    def gather(self, pipeline: Iterable[ShaderVariable]) -> tuple[dict, dict]:
        """Splits a pipeline into {uniform name: value} and {sampler name: TextureBox}"""
        values, samplers = {}, {}
        for variable in pipeline:
            if variable.type == "sampler2D":
                samplers[variable.name] = variable.value
            elif variable.value is not None:
                values[variable.name] = variable.value
        for module in self.scene.modules:
            if isinstance(module, ShaderTexture):
                for alias, box in module.sampler_names().items():
                    samplers.setdefault(alias, box)
        return values, samplers

    def resolve_samplers(self, samplers: dict[str, TextureBox]) -> list:
        out = []
        required = self.scene_info.get("required", len(self.scene_info["samplers"]))
        for index, name in enumerate(self.scene_info["samplers"]):
            box = samplers.get(name)
            if box is None or box.texture is None:
                if index >= required:
                    break                      # the history of a temporal texture ends where the scene made it end
                raise RuntimeError(f"Shader '{self.name}' samples '{name}' but the scene has no such texture")
            out.append(box.texture)
        return out

    def uniform_block(self, values: dict) -> N.Uniforms:
        w, h = self.scene.resolution
        block = N.Uniforms.defaults(w, h)
        return pack_uniforms(block, values, self.scene_info["extra"], self.scene_info.get("extra_types"))

    def set_uniform(self, name: str, value: Any = None) -> None:
        if self.scene_id is None:
            raise RuntimeError("Shader hasn't been compiled yet")

    # -- rendering -----------------------------------------------------------------------------
    def render(self, target: Optional[int] = None) -> None:
        """target: device pointer for the final pass output (defaults to the scene's frame buffer)"""
        if self.scene_id is None:
            raise RuntimeError("Shader hasn't been compiled yet")
        cuda, scene = self.scene.cuda, self.scene
        if self.texture.final:
            # shader.py:391-396: only iScreen's texture, iResolution and iSubsample
            source = scene.shader.texture
            sw, sh = source.size
            pointer, _ = source.get_box().texture.storage()
            cuda.render_final(pointer, sw, sh, scene.width, scene.height, scene.subsample, 3,
                              target if target is not None else scene.frame_pointer)
            return
        values, samplers = self.gather_planned()
        textures = self.resolve_samplers(samplers)
        # one pass per layer into the newest temporal slot, then the history rolls (shader.py:398-405)
        for layer, box in enumerate(self.texture.row(0)):
            values["iLayer"] = layer
            block = self.uniform_block(values)
            cuda.render_target(self.scene_id, block, textures, box.texture, self.filter_flags)
            box.empty = False
        self.texture.roll()

    def render_fused(self, target: int, ssaa: int) -> None:
        """K3+K4 in one launch straight into `target` (rgb24): iScreen never exists in HBM"""
        values, samplers = self.gather_planned()
        values["iLayer"] = 0
        scene = self.scene
        block, textures = self.uniform_block(values), self.resolve_samplers(samplers)
        events = scene.kernel_events
        if events is not None:
            import torch
            start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            start.record()
        scene.cuda.render_frame(self.scene_id, block, textures, scene.width, scene.height, ssaa,
                                scene.subsample, 3, target, self.filter_flags)
        if events is not None:
            end.record()
            events.append((start, end))

    def update(self) -> None:
        self.render()

    def handle(self, message) -> None:
        if isinstance(message, ShaderMessage.Shader.Compile):
            self.compile()
        elif isinstance(message, ShaderMessage.Shader.Render):
            self.render()

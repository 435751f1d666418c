"""
TEST INFRASTRUCTURE ONLY (oracle). Never imported by the product path.

Builds the reference's OWN `ShaderScene` objects headlessly in the build container and captures what they
hand to OpenGL: the assembled vertex / fragment GLSL text of every `ShaderProgram`
(`ShaderProgram.compile` → `scene.opengl.program(vertex, fragment)`, shaderflow/shader.py:313-324, with the
header `_build_shader` assembles, :190-239), the uniform names / types / values of `full_pipeline()`
(module.py:71-78) and every texture's sampling state (texture.py:104-137).

No OpenGL exists here, so the GL *objects* are recording stubs (`FakeContext`); everything above them — scene
graph, module order, metaprogramming, the example scenes of `examples/basic/demo.py` and
`examples/fractals/fractals.py` — is the reference's unmodified Python. GUI / device modules that are absent
are stubbed like in `oracle/ref_loader.py`; three stubs are functional because the scene graph computes with
them: `ordered_set.OrderedSet`, `quaternion` (Hamilton product only) and `watchdog.events`.

Only works where `/root/reference` exists: used by `tests/golden/make_golden_glsl.py` (which commits what it
captures as fixtures) and by the `-m "not gpu"` tests that re-capture the text live.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import tempfile
import types
from pathlib import Path
from unittest.mock import MagicMock

import numpy as np

from oracle import ref_loader

REFERENCE = ref_loader.REFERENCE


class _OrderedSet(dict):
    def __init__(self, items=()):
        super().__init__()
        for x in items:
            self[x] = None

    def add(self, x):
        self[x] = None

    def __iter__(self):
        return iter(list(self.keys()))


class _Quaternion:
    """What the camera needs of numpy-quaternion: construction, Hamilton product, scaling, conjugate"""
    dtype = "quaternion"
    shape = ()

    def __init__(self, w=1.0, x=0.0, y=0.0, z=0.0):
        self.q = np.array([w, x, y, z], float)

    w = property(lambda s: s.q[0]); x = property(lambda s: s.q[1])
    y = property(lambda s: s.q[2]); z = property(lambda s: s.q[3])

    def __mul__(a, b):
        if not isinstance(b, _Quaternion):
            return _Quaternion(*(a.q*b))
        w1, x1, y1, z1 = a.q
        w2, x2, y2, z2 = b.q
        return _Quaternion(w1*w2 - x1*x2 - y1*y2 - z1*z2, w1*x2 + x1*w2 + y1*z2 - z1*y2,
                           w1*y2 - x1*z2 + y1*w2 + z1*x2, w1*z2 + x1*y2 - y1*x2 + z1*w2)

    def __rmul__(a, b):
        return _Quaternion(*(a.q*b))

    def __truediv__(a, b):
        return _Quaternion(*(a.q/b))

    def __add__(a, b):
        return _Quaternion(*(a.q + (b.q if isinstance(b, _Quaternion) else b)))

    def __sub__(a, b):
        return _Quaternion(*(a.q - (b.q if isinstance(b, _Quaternion) else b)))

    def conjugate(self):
        return _Quaternion(self.q[0], *(-self.q[1:]))

    def __bool__(self):
        return bool(self.q.any())


class FakeContext(MagicMock):
    """Stands in for `moderngl.Context`: records program() sources"""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.info = {"GL_MAX_VIEWPORT_DIMS": (32768, 32768), "GL_RENDERER": "oracle/ref_scene.py recording stub"}


class _FakeWindow:
    def __init__(self, **kw):
        self.ctx = FakeContext()
        self.keys = MagicMock()
        self.size = kw.get("size")


_installed = False


def install() -> None:
    """Puts the stubs and the reference on sys.path / sys.modules (idempotent)"""
    global _installed
    if _installed:
        return
    if not ref_loader.available():
        raise RuntimeError(f"Reference tree not found at {REFERENCE}")
    if getattr(sys.modules.get("shaderflow"), "__sfb200_alias__", False):
        raise RuntimeError("The 'shaderflow' alias of shaderflow_b200 is installed in this process; "
                           "load the reference in a fresh interpreter")
    mod = types.ModuleType("ordered_set"); mod.OrderedSet = _OrderedSet
    sys.modules["ordered_set"] = mod
    quat = types.ModuleType("quaternion"); quat.quaternion = _Quaternion
    quat.as_vector_part = lambda q: q.q[1:].copy()
    quat.as_float_array = lambda q: q.q.copy()
    sys.modules["quaternion"] = quat
    events = types.ModuleType("watchdog.events")
    events.FileSystemEventHandler = type("FileSystemEventHandler", (), {})
    sys.modules["watchdog.events"] = events
    headless = types.ModuleType("moderngl_window.context.headless"); headless.Window = _FakeWindow
    sys.modules["moderngl_window.context.headless"] = headless
    for name in ref_loader._STUBS + ("imgui_bundle.python_backends", "moderngl_window.context",
                                     "moderngl_window.context.base", "moderngl_window.integrations",
                                     "moderngl_window.integrations.imgui_bundle", "shaderflow.temp.imgui_window"):
        sys.modules.setdefault(name, MagicMock())
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ["WINDOW_BACKEND"] = "headless"
    if str(REFERENCE) not in sys.path:
        sys.path.insert(0, str(REFERENCE))
    _installed = True


def _load_examples():
    install()
    out = {}
    for key, rel in (("basic", "examples/basic/demo.py"), ("fractals", "examples/fractals/fractals.py")):
        spec = importlib.util.spec_from_file_location(f"_sfb_ref_examples_{key}", REFERENCE/rel)
        module = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(module)
        out[key] = module
    return out


def capture(scene_name: str, background: np.ndarray | None = None, **attrs) -> dict:
    """Instantiates the reference example scene `scene_name` (class name in demo.py / fractals.py) and returns

        programs   {program name: {"vertex": text, "fragment": text}}   as handed to `opengl.program()`
        uniforms   {name: (glsl type, value)} from `scene.shader.full_pipeline()` (sampler entries → texture name)
        textures   {texture module name: dict(filter, repeat_x, repeat_y, components, dtype, temporal, layers)}
        modules    [(class name, module name)] in creation order

    `background` (RGB8, top row first) replaces the network-downloaded images (demo.py:19-37)."""
    examples = _load_examples()
    basic = examples["basic"]
    if background is None:
        background = np.zeros((4, 4, 3), np.uint8)
    from PIL import Image
    with tempfile.TemporaryDirectory() as tmp:
        png = Path(tmp)/"background.png"
        Image.fromarray(background).save(png)
        for asset in ("street", "ethereal"):
            setattr(basic.Assets, asset, staticmethod(lambda p=png: p))
        cls = getattr(basic, scene_name, None) or getattr(examples["fractals"], scene_name)
        scene = cls()
        scene.initialize()
        for key, value in attrs.items():
            setattr(scene, key, value)
        programs = {}
        from shaderflow.shader import ShaderProgram
        from shaderflow.texture import ShaderTexture
        for module in scene.modules:
            if isinstance(module, ShaderProgram):
                scene.opengl.program.reset_mock()
                module.compile()
                vertex, fragment = scene.opengl.program.call_args.args
                programs[module.name] = dict(vertex=vertex, fragment=fragment)
        uniforms = {}
        for var in scene.shader.full_pipeline():
            value = var.value
            if var.type == "sampler2D":
                value = None
            elif isinstance(value, np.ndarray):
                value = value.tolist()
            uniforms[var.name] = (var.type, value)
        textures = {}
        for module in scene.modules:
            if isinstance(module, ShaderTexture) and module.name:
                textures[module.name] = dict(
                    filter=module.filter.value, repeat_x=module.repeat_x, repeat_y=module.repeat_y,
                    components=module.components, dtype=np.dtype(module.dtype).name,
                    temporal=module.temporal, layers=module.layers, mipmaps=module.mipmaps)
        # ShaderSpectrogram sets its texture's filter per frame in update() (audio/spectrogram.py:300), which is not
        # run here: report the state update() would set
        for module in scene.modules:
            if type(module).__name__ == "ShaderSpectrogram":
                textures[module.name]["filter"] = "linear" if module.smooth else "nearest"
        modules = [(type(m).__name__, getattr(m, "name", None)) for m in scene.modules]
    return dict(programs=programs, uniforms=uniforms, textures=textures, modules=modules)


if __name__ == "__main__":
    import json
    name = sys.argv[1] if len(sys.argv) > 1 else "Visualizer"
    cap = capture(name)
    print(json.dumps({k: v for k, v in cap.items() if k != "programs"}, indent=1, default=str))
    for prog, src in cap["programs"].items():
        print(f"--- {prog}: vertex {len(src['vertex'].splitlines())} lines, fragment {len(src['fragment'].splitlines())} lines")

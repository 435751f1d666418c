"""ShaderTexture — a `temporal × layers` matrix of 2D textures a scene samples from and/or renders to.
API mirror of shaderflow/texture.py:73-382; what was a moderngl Texture+FBO per box is a libsfb200
texture (cudaArray + texture object + pitch-linear mirror, sfb200.h) whose linear storage doubles as
the render target."""
from __future__ import annotations

import itertools
from collections import deque
from enum import Enum
from typing import Any, Iterable, Optional, Union

import numpy as np
from attrs import Factory, define, field

from shaderflow_b200 import _native as N
from shaderflow_b200.message import ShaderMessage
from shaderflow_b200.module import ShaderModule
from shaderflow_b200.variable import ShaderVariable, Uniform


def numpy2mgltype(type: Union[np.dtype, str]) -> Optional[str]:
    """Kept for API compatibility (texture.py:28-38)"""
    if isinstance(type, str):
        return type
    if isinstance(type, np.dtype):
        type = type.type
    return {np.uint8: "f1", np.uint16: "u2", np.float16: "f2", np.float32: "f4"}.get(type)


def _native_dtype(dtype: np.dtype) -> int:
    table = {np.dtype(np.uint8): N.DTYPE_U8, np.dtype(np.float32): N.DTYPE_F32, np.dtype(np.float16): N.DTYPE_F16}
    if np.dtype(dtype) not in table:
        raise TypeError(f"ShaderTexture dtype {dtype} is not supported by the CUDA backend (uint8, float16, float32)")
    return table[np.dtype(dtype)]


class TextureFilter(Enum):
    Nearest = "nearest"
    Linear = "linear"


class Anisotropy(Enum):
    x1, x2, x4, x8, x16 = 1, 2, 4, 8, 16


@define
class TextureBox:
    _texture: Optional[N.Texture] = None     # libsfb200 texture (sampler + render target), created on first use
    data: Optional[bytes] = field(default=None, repr=False)
    clear: bool = False
    empty: bool = True
    factory: Any = field(default=None, repr=False)
    """() -> N.Texture; render targets that are never sampled or stored (iScreen on the fused path, iFinal)
    never allocate their 100+ MB of cudaArray + mirror"""

    @property
    def texture(self) -> Optional[N.Texture]:
        if self._texture is None and self.factory is not None:
            self._texture = self.factory()
            if self.data:
                self._texture.write(self.data)
        return self._texture

    @property
    def fbo(self) -> Optional[N.Texture]:
        """The reference pairs each texture with an FBO; here the texture's storage is the target"""
        return self.texture

    def release(self) -> None:
        if self._texture is not None:
            self._texture.destroy()
            self._texture = None

    def __del__(self):
        try: self.release()
        except Exception: pass


def _remake(self, attribute, value):
    value = attribute.converter(value) if attribute.converter else value
    if getattr(self, attribute.name) != value:
        object.__setattr__(self, attribute.name, value)   # bypass hooks
        self.make()
    return value


def _reapply(self, attribute, value):
    value = attribute.converter(value) if attribute.converter else value
    if getattr(self, attribute.name) != value:
        object.__setattr__(self, attribute.name, value)
        self.apply()
    return value


@define
class ShaderTexture(ShaderModule):
    name: str = None

    final: bool = field(default=False, converter=bool)
    track: float = field(default=0.0, converter=float, on_setattr=_remake)
    filter: TextureFilter = field(default=TextureFilter.Linear, converter=TextureFilter, on_setattr=_reapply)
    anisotropy: Anisotropy = field(default=Anisotropy.x16, converter=Anisotropy)
    mipmaps: bool = field(default=False, converter=bool)
    repeat_x: bool = field(default=True, converter=bool, on_setattr=_reapply)
    repeat_y: bool = field(default=True, converter=bool, on_setattr=_reapply)
    _width: int = field(default=1, converter=int)
    _height: int = field(default=1, converter=int)
    components: int = field(default=4, converter=int, on_setattr=_remake)
    dtype: np.dtype = field(default=np.uint8, converter=np.dtype, on_setattr=_remake)
    matrix: deque = Factory(deque)
    temporal: int = field(default=1, converter=int, on_setattr=_remake)
    layers: int = field(default=1, converter=int, on_setattr=_remake)
    external: Any = None
    """(device pointer holder, pointer) when the texture samples straight from a GPU track buffer"""

    def build(self):
        self.make()

    def repeat(self, value: bool) -> "ShaderTexture":
        object.__setattr__(self, "repeat_x", bool(value))
        object.__setattr__(self, "repeat_y", bool(value))
        return self.apply()

    # -- size ----------------------------------------------------------------------------------
    @property
    def width(self) -> int:
        return self.resolution[0] if self.track else self._width

    @width.setter
    def width(self, value: int):
        if self._width != int(value):
            self._width = int(value)
            self.make()

    @property
    def height(self) -> int:
        return self.resolution[1] if self.track else self._height

    @height.setter
    def height(self, value: int):
        if self._height != int(value):
            self._height = int(value)
            self.make()

    @property
    def resolution(self) -> tuple[int, int]:
        if not self.track:
            return (self._width, self._height)
        base = self.scene.resolution if self.final else self.scene.render_resolution
        return tuple(max(1, int(x*self.track)) for x in base)

    @resolution.setter
    def resolution(self, value: tuple[int, int]):
        if not self.track:
            w, h = value
            if (int(w), int(h)) != (self._width, self._height):
                self._width, self._height = int(w), int(h)
                self.make()

    size = resolution

    @property
    def aspect_ratio(self) -> float:
        return self.width/(self.height or 1)

    @property
    def zeros(self) -> np.ndarray:
        return np.zeros((*self.size, self.components), dtype=self.dtype)

    @property
    def bytes_per_pixel(self) -> int:
        return self.dtype.itemsize*self.components

    @property
    def size_t(self) -> int:
        return self.width*self.height*self.bytes_per_pixel

    # -- matrix --------------------------------------------------------------------------------
    @property
    def boxes(self) -> Iterable[tuple[int, int, TextureBox]]:
        for it, row in enumerate(self.matrix):
            for ib, box in enumerate(row):
                yield (it, ib, box)

    def row(self, n: int = 0) -> Iterable[TextureBox]:
        yield from self.matrix[n]

    def make(self) -> "ShaderTexture":
        """(Re)allocates every box at the current size; full-frame contents written earlier survive
        a same-size re-make (texture.py:268-270)"""
        while len(self.matrix) > self.temporal: self.matrix.pop()
        while len(self.matrix) < self.temporal: self.matrix.append(deque())
        for row in self.matrix:
            while len(row) > self.layers: row.pop()
            while len(row) < self.layers: row.append(TextureBox())
        cuda = getattr(self.scene, "cuda", None)
        if cuda is None:
            return self
        w, h = self.size
        signature = (w, h, self.components, str(self.dtype), self.temporal, self.layers)
        if signature == getattr(self, "_made", None):
            return self                               # nothing changed: keep the GPU objects (and contents)
        self._made = signature
        for (_, _, box) in self.boxes:
            box.release()
            if box.data and (self.size_t != len(box.data)):
                box.data = None
            box.factory = (lambda w=w, h=h: N.Texture(cuda, w, h, self.components, _native_dtype(self.dtype),
                linear=(self.filter is TextureFilter.Linear), repeat_x=self.repeat_x, repeat_y=self.repeat_y))
        self.external = None
        return self

    def apply(self) -> "ShaderTexture":
        for (_, _, box) in self.boxes:
            if box._texture is not None:                      # lazily created ones pick the state up at creation
                box._texture.set_sampling(self.filter is TextureFilter.Linear, self.repeat_x, self.repeat_y)
        return self

    def destroy(self) -> None:
        for (_, _, box) in self.boxes:
            box.release()

    def get_box(self, temporal: int = 0, layer: int = -1) -> Optional[TextureBox]:
        return self.matrix[temporal][layer]

    @property
    def fbo(self):
        return self.get_box().fbo

    @property
    def texture(self):
        return self.get_box().texture

    def roll(self, n: int = 1) -> "ShaderTexture":
        self.matrix.rotate(n)
        return self

    # -- input ---------------------------------------------------------------------------------
    def write(self, data=None, *, temporal: int = 0, layer: int = -1, viewport: tuple = None) -> "ShaderTexture":
        """data: bytes / numpy array (host) or a CUDA tensor (device, no host round trip)"""
        box = self.get_box(temporal, layer)
        if box.texture is not None:
            if isinstance(data, np.ndarray):
                data = np.ascontiguousarray(data)
            if self.external is not None:
                box.texture.bind_external(None)
                self.external = None
            box.texture.write(data, viewport=viewport)
        if (not viewport) and not getattr(data, "is_cuda", False):
            box.data = bytes(data) if not isinstance(data, np.ndarray) else data.tobytes()
        box.empty = False
        return self

    def bind(self, holder: Any, pointer: int, *, temporal: int = 0, layer: int = -1) -> "ShaderTexture":
        """Zero-copy: sample the box from a device buffer laid out [height][width][components] (e.g. one
        frame's row of the GPU audio track). `holder` keeps the buffer alive."""
        box = self.get_box(temporal, layer)
        if box.texture is not None:
            box.texture.bind_external(pointer)
        self.external = (holder, pointer)
        box.empty = False
        return self

    def from_numpy(self, data: np.ndarray) -> "ShaderTexture":
        shape = list(data.shape)
        if len(shape) == 2:
            shape.append(1)
        self._height, self._width = int(shape[0]), int(shape[1])
        object.__setattr__(self, "components", int(shape[2]))
        object.__setattr__(self, "dtype", np.dtype(data.dtype))
        self.make()
        self.write(np.flipud(data).tobytes())      # GL's v=0 is the image's bottom row (texture.py:334)
        return self

    def from_image(self, image) -> "ShaderTexture":
        from PIL import Image
        if isinstance(image, np.ndarray):
            return self.from_numpy(image)
        if not isinstance(image, Image.Image):
            image = Image.open(image)
        return self.from_numpy(np.array(image))

    def clear(self, temporal: int = 0, layer: int = -1) -> "ShaderTexture":
        return self.write(self.zeros, temporal=temporal, layer=layer)

    def is_empty(self, temporal: int = 0, layer: int = -1) -> bool:
        return self.get_box(temporal, layer).empty

    # -- module --------------------------------------------------------------------------------
    def _coord2name(self, temporal: int, layer: int) -> str:
        return f"{self.name}{temporal}x{layer}"

    def defines(self) -> Iterable[str]:
        """GLSL the reference injects (texture.py:354-368). The CUDA scenes bind samplers by name, so
        this text is informational; the alias rule it encodes is applied in `sampler_names`"""
        if not self.name:
            return
        for temporal in range(self.temporal):
            yield f"#define {self.name}{temporal or ''} {self.name}{temporal}x{self.layers-1}"
        yield f"vec4 {self.name}Texture(int temporal, int layer, vec2 astuv) {{"
        for (t, l) in itertools.product(range(self.temporal), range(self.layers)):
            yield f"    if (temporal == {t} && layer == {l})"
            yield f"        return texture({self._coord2name(t, l)}, astuv);"
        yield "    return vec4(0.0);"
        yield "}"

    def sampler_names(self) -> dict[str, TextureBox]:
        """Every GLSL name that resolves to a box of this texture: `X{t}x{l}`, and the aliases
        `X` → `X0x{L-1}`, `X{t}` → `X{t}x{L-1}`"""
        names = {}
        if not self.name:
            return names
        for (t, l, box) in self.boxes:
            names[self._coord2name(t, l)] = box
        for t in range(len(self.matrix)):
            names[f"{self.name}{t or ''}"] = self.matrix[t][self.layers - 1]
        return names

    def handle(self, message):
        if self.track and isinstance(message, ShaderMessage.Shader.RecreateTextures):
            # The reference recreates every tracked texture here (texture.py:250-270): what was RENDERED into it
            # is gone, only data written with write() comes back. A single-box target is overwritten before it is
            # read, so its GPU objects are kept; a texture with history (temporal / layers > 1) feeds earlier
            # renders back into later ones and must start every export empty (zero-filled on creation).
            if self.temporal > 1 or self.layers > 1:
                self._made = None
            self.make()

    def pipeline(self) -> Iterable[ShaderVariable]:
        if not self.name:
            return
        yield Uniform("vec2", f"{self.name}Size", self.size)
        yield Uniform("int", f"{self.name}Layers", self.layers)
        yield Uniform("int", f"{self.name}Temporal", self.temporal)
        for (t, l, box) in self.boxes:
            yield Uniform("sampler2D", self._coord2name(t, l), box)

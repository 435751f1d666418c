"""Shader variables: the name/type/value carriers every module's `pipeline()` yields.
API mirror of shaderflow/variable.py:46-99. With the CUDA backend `declaration` is never compiled;
it is kept because user code and tools print it."""
from __future__ import annotations

import copy
from typing import Any, Optional

from attrs import define

_ORDER = ("interpolation", "direction", "qualifier", "type", "name")
_SIZES = dict(float="f", int="i", bool="i", vec2="2f", vec3="3f", vec4="4f")


@define(eq=False, slots=True)
class ShaderVariable:
    type: str
    name: str
    value: Optional[Any] = None
    qualifier: Optional[str] = None
    direction: Optional[str] = None
    interpolation: Optional[str] = None

    def __hash__(self) -> int:
        return hash(self.name)

    def __eq__(self, other) -> bool:
        return self.name == getattr(other, "name", None)

    def copy(self, **update) -> "ShaderVariable":
        clone = copy.deepcopy(self)
        for key, value in update.items():
            setattr(clone, key, value)
        return clone

    @property
    def size_string(self) -> Optional[str]:
        return _SIZES.get(self.type)

    @property
    def declaration(self) -> str:
        return " ".join(filter(None, (getattr(self, k, None) for k in _ORDER))).strip() + ";"


@define(eq=False, slots=True)
class Uniform(ShaderVariable):
    qualifier: Optional[str] = "uniform"


@define(eq=False, slots=True)
class InVariable(ShaderVariable):
    direction: Optional[str] = "in"


@define(eq=False, slots=True)
class OutVariable(ShaderVariable):
    direction: Optional[str] = "out"


@define(eq=False, slots=True)
class FlatVariable(ShaderVariable):
    interpolation: Optional[str] = "flat"

"""Second-order dynamics on the host (API mirror of shaderflow/dynamics.py): DynamicNumber and the
ShaderDynamics module that exposes value / integral / derivative as uniforms.

Host numpy is the right place for these: they are O(1) scalars or 3-vectors per frame (camera, user
knobs). The heavy instances — the per-bin spectrogram smoother and the audio volume/std pair — run on
the GPU as a batched scan (csrc/audio.cu) and only *publish* through this class."""
from __future__ import annotations

import math
from copy import deepcopy
from numbers import Number
from typing import Any, Iterable, Optional

import numpy as np
from attrs import define, field

from shaderflow_b200.module import ShaderModule
from shaderflow_b200.variable import ShaderVariable, Uniform


class _Arithmetic(Number):
    """Lets a DynamicNumber stand in for its value in expressions"""
    def __float__(self): return float(self.value)
    def __int__(self): return int(self.value)
    def __str__(self): return str(self.value)
    def __mul__(self, o): return self.value*o
    def __rmul__(self, o): return self*o
    def __add__(self, o): return self.value + o
    def __radd__(self, o): return self + o
    def __sub__(self, o): return self.value - o
    def __rsub__(self, o): return self - o
    def __truediv__(self, o): return self.value/o
    def __rtruediv__(self, o): return self/o
    def __floordiv__(self, o): return self.value//o
    def __rfloordiv__(self, o): return self//o
    def __mod__(self, o): return self.value % o
    def __rmod__(self, o): return self % o
    def __pow__(self, o): return self.value**o
    def __rpow__(self, o): return self**o


def _as_array(self, attribute, value):
    return self._ensure_numpy(value)


def _max_abs_difference(a: np.ndarray, b: np.ndarray) -> float:
    """max(abs(a - b)); scalars and short vectors (every camera system) skip numpy's dispatch — this runs
    for every system on every frame, including the frames a sharded rank only steps through"""
    if a.size == 1 and b.size == 1:
        return abs(a.item() - b.item())
    if a.size <= 4 and a.shape == b.shape:
        return max(abs(x - y) for x, y in zip(a.ravel().tolist(), b.ravel().tolist()))
    return float(np.abs(a - b).max())


@define(slots=False)
class DynamicNumber(_Arithmetic):
    """y'' k2 + y' k1 + y = x + k3 x' integrated with semi-implicit Euler (dynamics.py:197-250)"""

    def _ensure_numpy(self, value):
        if isinstance(value, np.ndarray):
            return value
        dtype = getattr(value, "dtype", self.dtype)
        if str(dtype) == "quaternion":
            return value
        return np.array(value, dtype=dtype)

    value: np.ndarray = field(default=0, on_setattr=_as_array)
    target: np.ndarray = field(default=0, on_setattr=_as_array)
    dtype: np.dtype = field(default=np.float64)
    initial: np.ndarray = field(default=None)
    frequency: float = 1.0
    zeta: float = 1.0
    response: float = 0.0
    precision: float = 1e-6
    integral: np.ndarray = 0.0
    integrate: bool = False
    derivative: np.ndarray = 0.0
    acceleration: np.ndarray = 0.0
    previous: np.ndarray = 0.0
    _at_rest: Any = None

    def __attrs_post_init__(self):
        self.set(self.target or self.value)

    def set(self, value, *, instant: bool = True) -> None:
        value = self._ensure_numpy(value)
        self.value = deepcopy(value) if instant else self.value
        self.target = deepcopy(value)
        self.initial = deepcopy(value)
        self.previous = deepcopy(value) if instant else self.previous
        zeros = np.zeros_like(value)
        self.integral, self.derivative, self.acceleration = deepcopy(zeros), deepcopy(zeros), deepcopy(zeros)

    def reset(self, instant: bool = False):
        self.set(self.initial, instant=instant)

    @property
    def radians(self) -> float:
        return math.tau*self.frequency

    @property
    def k1(self) -> float:
        return self.zeta/(math.pi*self.frequency)

    @property
    def k2(self) -> float:
        return 1.0/(self.radians*self.radians)

    @property
    def k3(self) -> float:
        return (self.response*self.zeta)/(math.tau*self.frequency)

    @property
    def damping(self) -> float:
        return self.radians*(abs(self.zeta*self.zeta - 1.0))**0.5

    def next(self, target=None, dt: float = 1.0):
        if not dt:
            return self.value
        if target is not None:
            self.target = self._ensure_numpy(target)
            if self.target.shape != self.value.shape:
                self.set(target)
        # A system that was at rest last time and whose target and value still hold the same bytes (in-place edits
        # of either are seen) is at rest again: every export steps ~10 such systems per frame, and a sharded rank
        # steps them through all the frames before its range
        key = None
        if not self.integrate:
            key = (self.target.tobytes(), self.value.tobytes())
            if key == self._at_rest:
                return self.value
        if _max_abs_difference(self.target, self.value) < self.precision:
            if self.integrate:
                self.integral += self.value*dt
            self._at_rest = key
            return self.value
        self._at_rest = None
        velocity = (self.target - self.previous)/dt
        self.previous = self.target
        if self.radians*dt < self.zeta:        # clamp k2 to stable values
            k1 = self.k1
            k2 = max(k1*dt, self.k2, 0.5*(k1 + dt)*dt)
        else:                                  # pole matching for fast systems
            t1 = math.exp(-1*self.zeta*self.radians*dt)
            a1 = 2*t1*(math.cos if self.zeta <= 1 else math.cosh)(self.damping*dt)
            t2 = 1/(1 + t1*t1 - a1)*dt
            k1 = t2*(1 - t1*t1)
            k2 = t2*dt
        self.value += self.derivative*dt
        self.acceleration = (self.target + self.k3*velocity - self.value - k1*self.derivative)/k2
        self.derivative += self.acceleration*dt
        if self.integrate:
            self.integral += self.value*dt
        return self.value

    @staticmethod
    def extract(*objects):
        return tuple(o.value if isinstance(o, DynamicNumber) else o for o in objects)


@define
class ShaderDynamics(ShaderModule, DynamicNumber):
    name: str = "iShaderDynamics"
    real: bool = False
    primary: bool = True
    differentiate: bool = False
    published: bool = False
    """True while a GPU track owns this system: update() then leaves value/integral alone"""

    def build(self) -> None:
        DynamicNumber.__attrs_post_init__(self)

    def setup(self) -> None:
        self.reset(instant=self.scene.freewheel)

    def update(self) -> None:
        if self.published:
            return
        self.next(dt=abs(self.scene.rdt if self.real else self.scene.dt))

    @property
    def type(self) -> Optional[str]:
        shape = np.shape(self.value)
        if not shape or shape[0] == 1:
            return "float"
        return {2: "vec2", 3: "vec3", 4: "vec4"}.get(shape[0])

    def pipeline(self) -> Iterable[ShaderVariable]:
        if not self.type:
            return
        if self.primary:
            yield Uniform(self.type, f"{self.name}", self.value)
        if self.integrate:
            yield Uniform(self.type, f"{self.name}Integral", self.integral)
        if self.differentiate:
            yield Uniform(self.type, f"{self.name}Derivative", self.derivative)

"""ShaderCamera — uniform producer for camera.glsl's `GetCamera` (API mirror of shaderflow/camera.py).

During an export nobody presses keys, so the camera is whatever the scene script set; the module keeps
the nine ShaderDynamics (camera.py:146-185), the basis vectors from the rotation quaternion and the
move/rotate/align/look helpers so scripted cameras work. numpy-quaternion is replaced by the small
`Quaternion` class below (the reference only uses construction, product, conjugate, norm and
vector part: camera.py:96-102,216-217)."""
from __future__ import annotations

import math
from enum import Enum
from typing import Iterable

import numpy as np
from attrs import define, field

from shaderflow_b200.dynamics import DynamicNumber, ShaderDynamics
from shaderflow_b200.keyboard import ShaderKeyboard
from shaderflow_b200.message import ShaderMessage
from shaderflow_b200.module import ShaderModule
from shaderflow_b200.variable import ShaderVariable, Uniform


class Quaternion(np.ndarray):
    """(w, x, y, z) as a 4-vector with the Hamilton product on `*`; element-wise maths otherwise, which
    is what the dynamics integrator needs"""
    def __new__(cls, w=1.0, x=0.0, y=0.0, z=0.0):
        return np.asarray((w, x, y, z), dtype=np.float64).view(cls)

    def __mul__(self, other):
        if isinstance(other, Quaternion):
            a, b = np.asarray(self), np.asarray(other)
            w = a[0]*b[0] - a[1]*b[1] - a[2]*b[2] - a[3]*b[3]
            x = a[0]*b[1] + a[1]*b[0] + a[2]*b[3] - a[3]*b[2]
            y = a[0]*b[2] - a[1]*b[3] + a[2]*b[0] + a[3]*b[1]
            z = a[0]*b[3] + a[1]*b[2] - a[2]*b[1] + a[3]*b[0]
            return Quaternion(w, x, y, z)
        return np.ndarray.__mul__(self, other)

    def conjugate(self):
        return Quaternion(self[0], -self[1], -self[2], -self[3])

    @property
    def vector(self) -> np.ndarray:
        return np.asarray(self[1:], dtype=np.float64)


class GlobalBasis:
    Origin   = np.array(( 0,  0,  0), dtype=np.float64)
    Null     = np.array(( 0,  0,  0), dtype=np.float64)
    Up       = np.array(( 0,  1,  0), dtype=np.float64)
    Down     = np.array(( 0, -1,  0), dtype=np.float64)
    Left     = np.array((-1,  0,  0), dtype=np.float64)
    Right    = np.array(( 1,  0,  0), dtype=np.float64)
    Forward  = np.array(( 0,  0,  1), dtype=np.float64)
    Backward = np.array(( 0,  0, -1), dtype=np.float64)


class CameraProjection(Enum):
    Perspective = 0
    Stereoscopic = 1
    Equirectangular = 2

    def __next__(self):
        return CameraProjection((self.value + 1) % 3)

    @classmethod
    def _missing_(cls, value):
        table = dict(perspective=0, default=0, stereoscopic=1, stereo=1, vr=1, sbs=1, spherical=2, equirectangular=2)
        table["360"] = 2
        if value in table:
            return cls(table[value])
        raise ValueError(f"{value} is not a valid {cls.__name__}")


class CameraMode(Enum):
    FreeCamera = 0
    Camera2D = 1
    Spherical = 2

    @classmethod
    def _missing_(cls, value):
        table = {"free": 0, "freecamera": 0, "2d": 1, "plane": 1, "flat": 1, "spherical": 2, "aligned": 2}
        if value in table:
            return cls(table[value])
        raise ValueError(f"{value} is not a valid {cls.__name__}")


class Algebra:
    @staticmethod
    def quaternion(axis, degrees: float) -> Quaternion:
        theta = math.radians(degrees/2)
        return Quaternion(math.cos(theta), *(math.sin(theta)*np.asarray(axis, dtype=np.float64)))

    @staticmethod
    def rotate_vector(vector, R) -> np.ndarray:
        R = R if isinstance(R, Quaternion) else Quaternion(*np.asarray(R))
        return (R * Quaternion(0, *vector) * R.conjugate()).vector

    @staticmethod
    def angle(A, B) -> float:
        A, B = DynamicNumber.extract(A, B)
        la, lb = np.linalg.norm(A), np.linalg.norm(B)
        if not la or not lb:
            return 0.0
        return float(np.degrees(np.arccos(np.clip(np.dot(A, B)/(la*lb), -1, 1))))

    @staticmethod
    def unit_vector(vector):
        magnitude = np.linalg.norm(vector)
        return (vector/magnitude) if magnitude else vector


@define(slots=False)
class ShaderCamera(ShaderModule):
    name: str = "iCamera"
    mode: CameraMode = field(default=CameraMode.Camera2D, converter=CameraMode)
    projection: CameraProjection = field(default=CameraProjection.Perspective, converter=CameraProjection)
    separation: ShaderDynamics = None
    rotation: ShaderDynamics = None
    position: ShaderDynamics = None
    zenith: ShaderDynamics = None
    zoom: ShaderDynamics = None
    isometric: ShaderDynamics = None
    focus: ShaderDynamics = None
    orbital: ShaderDynamics = None
    dolly: ShaderDynamics = None

    def build(self):
        def dyn(suffix, frequency, value, **kw):
            return ShaderDynamics(scene=self.scene, name=f"{self.name}{suffix}", real=True,
                                  frequency=frequency, zeta=1, response=0, value=value, **kw)
        # creation order = uniform order of the reference (camera.py:146-185)
        self.position   = dyn("Position", 4, np.copy(GlobalBasis.Origin))
        self.separation = dyn("Separation", 0.5, 0.05)
        self.rotation   = dyn("Rotation", 5, Quaternion(1, 0, 0, 0), primary=False)
        self.zenith     = dyn("Zenith", 1, np.copy(GlobalBasis.Up))
        self.zoom       = dyn("Zoom", 3, 1)
        self.isometric  = dyn("Isometric", 1, 0)
        self.focus      = dyn("FocalLength", 1, 1)
        self.orbital    = dyn("Orbital", 1, 0)
        self.dolly      = dyn("Dolly", 1, 0)

    @property
    def fov(self) -> float:
        return 2.0*math.degrees(math.atan(self.zoom.value - self.isometric.value))

    @fov.setter
    def fov(self, value: float):
        self.zoom.target = math.tan(math.radians(value)/2.0) + self.isometric.value

    def pipeline(self) -> Iterable[ShaderVariable]:
        yield Uniform("int",  f"{self.name}Mode",       value=self.mode.value)
        yield Uniform("int",  f"{self.name}Projection", value=self.projection.value)
        yield Uniform("vec3", f"{self.name}Right",      value=self.right)
        yield Uniform("vec3", f"{self.name}Upward",     value=self.up)
        yield Uniform("vec3", f"{self.name}Forward",    value=self.forward)

    # -- scripted motion -----------------------------------------------------------------------
    def move(self, direction, absolute: bool = False):
        self.position.target = self.position.target + (direction - (self.position.target*absolute))
        return self

    def rotate(self, direction, degrees: float = 0.0):
        q = Algebra.quaternion(direction, degrees) * Quaternion(*np.asarray(self.rotation.target))
        self.rotation.target = Quaternion(*(np.asarray(q)/np.linalg.norm(np.asarray(q))))
        return self

    def rotate2d(self, degrees: float = 0.0):
        target = Algebra.rotate_vector(self.zenith.value, Algebra.quaternion(self.forward_target, degrees))
        return self.align(self.up_target, target)

    def align(self, A, B, degrees: float = 0.0):
        A, B = DynamicNumber.extract(A, B)
        return self.rotate(Algebra.unit_vector(np.cross(A, B)), Algebra.angle(A, B) - degrees)

    def look(self, target):
        return self.align(self.forward_target, target - self.position.target)

    def apply_zoom(self, value: float) -> None:
        if value > 0: self.zoom.target = self.zoom.target*(1 + value)
        else:         self.zoom.target = self.zoom.target/(1 - value)

    def update(self):
        # camera.py:240-278 — every branch is keyed on a pressed key or the Spherical mode
        if self.mode == CameraMode.Spherical:
            self.align(self.right_target, self.zenith.target, 90)
        keys = self.scene.keyboard
        if keys is not None and any(keys._pressed.values()):
            dt = abs(self.scene.dt or self.scene.rdt)
            K = ShaderKeyboard.Keys
            move = np.copy(GlobalBasis.Null)
            flat = (self.mode == CameraMode.Camera2D)
            for key, vec in ((K.W, GlobalBasis.Up if flat else GlobalBasis.Forward), (K.A, GlobalBasis.Left),
                             (K.S, GlobalBasis.Down if flat else GlobalBasis.Backward), (K.D, GlobalBasis.Right)):
                if keys(key): move += vec
            if move.any():
                move = Algebra.rotate_vector(move, self.rotation.target)
                self.move(2*Algebra.unit_vector(move)*self.zoom.value*dt)

    def handle(self, message):
        if isinstance(message, ShaderMessage.Mouse.Scroll):
            self.apply_zoom(-0.05*message.dy)

    # -- basis ---------------------------------------------------------------------------------
    _basis_cache: dict = None

    def _basis(self, vector, which: str) -> np.ndarray:
        """Rotated basis vector; cached on the rotation's value (an export with a still camera asks for the
        same three vectors every frame)"""
        rotation = getattr(self.rotation, which)
        key = (which, vector.tobytes(), np.asarray(rotation, dtype=np.float64).tobytes())
        cache = self._basis_cache
        if cache is None or len(cache) > 64:
            cache = self._basis_cache = {}
        hit = cache.get(key)
        if hit is None:
            hit = cache[key] = Algebra.rotate_vector(vector, rotation)
        return hit.copy()

    right           = property(lambda s: s._basis(GlobalBasis.Right, "value"))
    right_target    = property(lambda s: s._basis(GlobalBasis.Right, "target"))
    left            = property(lambda s: -1*s.right)
    left_target     = property(lambda s: -1*s.right_target)
    up              = property(lambda s: s._basis(GlobalBasis.Up, "value"))
    up_target       = property(lambda s: s._basis(GlobalBasis.Up, "target"))
    down            = property(lambda s: -1*s.up)
    down_target     = property(lambda s: -1*s.up_target)
    forward         = property(lambda s: s._basis(GlobalBasis.Forward, "value"))
    forward_target  = property(lambda s: s._basis(GlobalBasis.Forward, "target"))
    backward        = property(lambda s: -1*s.forward)
    backward_target = property(lambda s: -1*s.forward_target)

    def _axis(i):  # noqa: N805
        def get(self): return self.position.value[i]
        def put(self, value): self.position.target[i] = value
        return property(get, put)
    x, y, z = _axis(0), _axis(1), _axis(2)
    del _axis

"""ShaderModule — the plug-in contract of a scene (API mirror of shaderflow/module.py:19-178):
self-registration into `scene.modules`, the hook set (build/setup/update/pipeline/handle/defines/
includes/ffhook/duration/destroy) and the relay/find/full_pipeline helpers. UI hooks are accepted and
never called (no window in the CUDA backend)."""
from __future__ import annotations

import itertools
import weakref
from typing import TYPE_CHECKING, Any, Iterable
from weakref import CallableProxyType, ProxyType

from attrs import Factory, define, field

from shaderflow_b200 import logger
from shaderflow_b200.variable import ShaderVariable

if TYPE_CHECKING:
    from shaderflow_b200.scene import ShaderScene

_uuid = itertools.count(1)


@define(slots=False)
class ShaderModule:
    scene: "ShaderScene" = field(default=None, repr=False)
    uuid: int = Factory(lambda: next(_uuid))
    name: str = None

    def __attrs_post_init__(self):
        from shaderflow_b200.scene import ShaderScene
        # the first module created is the scene itself; everyone holds a weak proxy to it
        if not isinstance(self.scene or self, (CallableProxyType, ProxyType)):
            self.scene = weakref.proxy(self.scene or self)
        if not isinstance(self.scene, ShaderScene):
            raise RuntimeError(logger.error(
                f"Module of type '{type(self).__name__}' must be added to a 'ShaderScene' instance: "
                f"initialize it with {type(self).__name__}(scene=<ShaderScene>, ...)"))
        self.scene.modules.append(self)
        self.commands()
        if not isinstance(self, ShaderScene):
            self.build()

    # -- hooks (override any) ------------------------------------------------------------------
    def build(self) -> None: ...
    def setup(self) -> None: ...
    def update(self) -> None: ...
    def handle(self, message: Any) -> None: ...
    def commands(self) -> None: ...
    def destroy(self) -> None: ...
    def ffhook(self, ffmpeg: Any) -> None: ...
    def ui(self) -> None: ...
    def __ui__(self) -> None: ...
    def __shaderflow_ui__(self) -> None: ...

    def pipeline(self) -> Iterable[ShaderVariable]:
        return []

    def includes(self) -> Iterable[Any]:
        return ()

    def defines(self) -> Iterable[str]:
        return ()

    @property
    def duration(self) -> float:
        return 0.0

    # -- helpers -------------------------------------------------------------------------------
    def full_pipeline(self) -> Iterable[ShaderVariable]:
        for module in self.scene.modules:
            yield from (module.pipeline() or ())

    def relay(self, message: Any) -> "ShaderModule":
        if isinstance(message, type):
            message = message()
        for module in list(self.scene.modules):
            module.handle(message)
        return self

    def find(self, type: type) -> Iterable["ShaderModule"]:
        for module in self.scene.modules:
            if isinstance(module, type):
                yield module

    def __del__(self) -> None:
        try:
            self.destroy()
        except Exception:
            pass

    # -- logging -------------------------------------------------------------------------------
    @property
    def who(self) -> str:
        return f"(Module {self.uuid:>2} • {type(self).__name__[:12].ljust(12)})"

    def log_info(self, *a, **k):  return logger.info(self.who, *a)
    def log_warn(self, *a, **k):  return logger.warn(self.who, *a)
    def log_error(self, *a, **k): return logger.error(self.who, *a)
    def log_debug(self, *a, **k): return logger.debug(self.who, *a)
    def log_minor(self, *a, **k): return logger.minor(self.who, *a)

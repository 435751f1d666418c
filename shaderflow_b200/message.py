"""Messages relayed between modules — the names and fields of shaderflow/message.py, because user `handle()` code
isinstance-checks them and reads their fields. An offline export only ever relays Shader.* and Window.FileDrop; the
input messages exist so that such code keeps importing and running.

The schema is one table (message → fields with their defaults); the classes are made from it."""
from __future__ import annotations

from typing import Any

import attrs


def _message(name: str, bases: tuple = (), /, **fields):
    made = attrs.make_class(name, {key: attrs.field(default=default) for key, default in fields.items()}, bases=bases or (object,))
    made.__qualname__ = name
    return made


def _group(name: str, **members) -> type:
    for key, member in members.items():
        member.__qualname__ = f"ShaderMessage.{name}.{key}"
    return type(name, (), members)


class _FileAccess:
    """Window.FileDrop: the dropped paths by position"""
    def get(self, index: int):
        files = self.files or ()
        return files[index] if index < len(files) else None

    first, second, third = (property(lambda self, k=k: self.get(k)) for k in range(3))


class _Extent:
    """Window.Resize"""
    @property
    def size(self) -> tuple:
        return (self.width, self.height)


_POINTER = dict(x=0, y=0, dx=0, dy=0, u=0.0, v=0.0, du=0.0, dv=0.0)
_BUTTON = dict(button=0, x=0, y=0, u=0.0, v=0.0)


class ShaderMessage:
    class Custom:
        data: Any

    Mouse = _group("Mouse",
        Position=_message("Position", **_POINTER), Drag=_message("Drag", **_POINTER),
        Press=_message("Press", **_BUTTON), Release=_message("Release", **_BUTTON),
        Scroll=_message("Scroll", dx=0, dy=0, du=0.0, dv=0.0), Enter=_message("Enter", state=False))

    Window = _group("Window",
        Resize=_message("Resize", (_Extent,), width=None, height=None), Iconify=_message("Iconify", state=None),
        FileDrop=_message("FileDrop", (_FileAccess,), files=None), Close=_message("Close"))

    Shader = _group("Shader",
        RecreateTextures=_message("RecreateTextures"), Compile=_message("Compile"), Render=_message("Render"))

    Keyboard = _group("Keyboard",
        Press=_message("Press", key=None, action=None, modifiers=None), KeyDown=_message("KeyDown", key=None, modifiers=None),
        KeyUp=_message("KeyUp", key=None, modifiers=None), Unicode=_message("Unicode", char=None))

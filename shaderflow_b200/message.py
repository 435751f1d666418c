"""Messages relayed between modules (API mirror of shaderflow/message.py). Only the Shader.* and
Window.FileDrop messages occur in an offline export; the input ones exist so user `handle()` code
that isinstance-checks them keeps working."""
from __future__ import annotations

from typing import Any, Optional

from attrs import define


def _motion(name: str):
    return define(type(name, (), dict(__annotations__=dict(x=int, y=int, dx=int, dy=int, u=float, v=float, du=float, dv=float),
                                      x=0, y=0, dx=0, dy=0, u=0.0, v=0.0, du=0.0, dv=0.0)))


def _click(name: str):
    return define(type(name, (), dict(__annotations__=dict(button=int, x=int, y=int, u=float, v=float),
                                      button=0, x=0, y=0, u=0.0, v=0.0)))


class ShaderMessage:

    class Custom:
        data: Any

    class Mouse:
        Position = _motion("Position")
        Drag = _motion("Drag")
        Press = _click("Press")
        Release = _click("Release")

        @define
        class Scroll:
            dx: int = 0
            dy: int = 0
            du: float = 0.0
            dv: float = 0.0

        @define
        class Enter:
            state: bool = False

    class Window:
        @define
        class Resize:
            width: Optional[int] = None
            height: Optional[int] = None

            @property
            def size(self) -> tuple:
                return self.width, self.height

        @define
        class Iconify:
            state: Optional[bool] = None

        @define
        class FileDrop:
            files: Optional[list] = None

            def get(self, index: int) -> Optional[str]:
                if self.files and index < len(self.files):
                    return self.files[index]

            first = property(lambda self: self.get(0))
            second = property(lambda self: self.get(1))
            third = property(lambda self: self.get(2))

        @define
        class Close:
            ...

    class Shader:
        @define
        class RecreateTextures:
            ...

        @define
        class Compile:
            ...

        @define
        class Render:
            ...

    class Keyboard:
        @define
        class Press:
            key: Optional[int] = None
            action: Optional[int] = None
            modifiers: Optional[int] = None

        @define
        class KeyDown:
            key: Optional[int] = None
            modifiers: Optional[int] = None

        @define
        class KeyUp:
            key: Optional[int] = None
            modifiers: Optional[int] = None

        @define
        class Unicode:
            char: Optional[str] = None

"""Keyboard state module (API mirror of shaderflow/keyboard.py). There is no window, so no key is
ever pressed; the key constants are a static table because ShaderCamera.update reads
`ShaderKeyboard.Keys.W` etc. every frame (camera.py:247-258)."""
from __future__ import annotations

from types import SimpleNamespace
from typing import Iterable

from attrs import Factory, define

from shaderflow_b200.message import ShaderMessage
from shaderflow_b200.module import ShaderModule
from shaderflow_b200.variable import ShaderVariable

_NAMES = ("A B C D E F G H I J K L M N O P Q R S T U V W X Y Z SPACE TAB ESCAPE ENTER UP DOWN LEFT RIGHT "
          "LEFT_SHIFT LEFT_CTRL LEFT_ALT F1 F2 F3 F4 F5 F6 F7 F8 F9 F10 F11 F12 "
          "NUMBER_0 NUMBER_1 NUMBER_2 NUMBER_3 NUMBER_4 NUMBER_5 NUMBER_6 NUMBER_7 NUMBER_8 NUMBER_9").split()


@define
class ShaderKeyboard(ShaderModule):
    Keys = SimpleNamespace(ACTION_PRESS=1, ACTION_RELEASE=0, **{name: 1000 + i for i, name in enumerate(_NAMES)})
    DirKeys = {name: 1000 + i for i, name in enumerate(_NAMES)}

    _pressed: dict = Factory(dict)

    @staticmethod
    def set_keymap(keymap) -> None:
        ShaderKeyboard.DirKeys = {k: getattr(keymap, k) for k in dir(keymap) if not k.startswith("_")}
        ShaderKeyboard.Keys = keymap

    def pressed(self, key=None) -> bool:
        return self._pressed.setdefault(key, False)

    def __call__(self, *a, **k) -> bool:
        return self.pressed(*a, **k)

    def pipeline(self) -> Iterable[ShaderVariable]:
        return ()

    def handle(self, message):
        if isinstance(message, ShaderMessage.Keyboard.Press):
            self._pressed[message.key] = (message.action != ShaderKeyboard.Keys.ACTION_RELEASE)

"""Frame time statistics (API mirror of shaderflow/frametimer.py, minus the imgui plot)."""
from __future__ import annotations

from collections import deque

from attrs import Factory, define

from shaderflow_b200.module import ShaderModule


@define
class ShaderFrametimer(ShaderModule):
    frametimes: deque = Factory(lambda: deque(maxlen=600))
    history: float = 2.0

    def update(self):
        self.frametimes.append(self.scene.rdt)

    def percent(self, percent: float = 1) -> list:
        cut = max(1, int(len(self.frametimes)*(percent/100)))
        return sorted(self.frametimes)[-cut:]

    @property
    def average_fps(self) -> float:
        total = sum(self.frametimes)
        return len(self.frametimes)/total if total else 0.0

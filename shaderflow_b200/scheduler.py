"""Task scheduler. API mirror of shaderflow/scheduler.py; the export path only uses the freewheel
mode, whose arithmetic (scheduler.py:86-89,152-168) decides every frame's dt and is reproduced exactly
(it is also what sfb_frame_clock computes in bulk)."""
from __future__ import annotations

import contextlib
import inspect
import time
from collections import deque
from typing import Any, Callable, Iterable, Optional

from attrs import Factory, define, field


def precise_sleep(seconds: float, *, error: float = 0.001) -> None:
    start = time.monotonic()
    if seconds - error > 0:
        time.sleep(seconds - error)
    while (time.monotonic() - start) < seconds:
        pass


@define(eq=False)
class SchedulerTask:
    task: Callable
    args: list = field(factory=list, repr=False)
    kwargs: dict = field(factory=dict, repr=False)
    output: Any = field(default=None, repr=False)
    context: Any = Factory(contextlib.nullcontext)
    enabled: bool = True
    once: bool = False
    frequency: float = 60.0
    frameskip: bool = True
    freewheel: bool = False
    precise: bool = False
    started: float = Factory(time.monotonic)
    next_call: Optional[float] = None
    last_call: Optional[float] = None
    _dt: bool = False

    def __attrs_post_init__(self):
        self._dt = "dt" in inspect.signature(self.task).parameters
        if self.freewheel:
            self.started = 0
        self.last_call = (self.last_call or self.started) - self.period
        self.next_call = (self.next_call or self.started)

    def __hash__(self) -> int:
        return id(self)

    @property
    def fps(self) -> float:
        return self.frequency

    @fps.setter
    def fps(self, value: float):
        self.frequency = value

    @property
    def period(self) -> float:
        return 1.0/self.frequency

    @period.setter
    def period(self, value: float):
        self.frequency = 1/value

    @property
    def should_delete(self) -> bool:
        return self.once and not self.enabled

    @property
    def should_live(self) -> bool:
        return not self.should_delete

    def __lt__(self, other) -> bool:
        return True if (self.once and not other.once) else (self.next_call < other.next_call)

    def __gt__(self, other) -> bool:
        return True if (not self.once and other.once) else (self.next_call > other.next_call)

    def next(self, block: bool = True) -> "SchedulerTask":
        if not self.freewheel:
            wait = max(0, self.next_call - time.monotonic())
            if (not block) and wait > 0:
                return self
            (precise_sleep if self.precise else time.sleep)(wait)
        now = self.next_call if self.freewheel else time.monotonic()
        if self._dt:
            self.kwargs["dt"] = now - self.last_call
            if not self.frameskip:
                self.kwargs["dt"] = min(self.kwargs["dt"], self.period)
        self.last_call = now
        with self.context:
            self.output = self.task(*self.args, **self.kwargs)
        while self.next_call <= now:
            self.next_call += self.period
        self.enabled = not self.once
        return self


@define
class Scheduler:
    Task = SchedulerTask
    tasks: deque = Factory(deque)

    def add(self, task: SchedulerTask) -> SchedulerTask:
        self.tasks.append(task)
        return task

    def new(self, task: Callable, **options) -> SchedulerTask:
        return self.add(SchedulerTask(task=task, **options))

    def once(self, task: Callable, **options) -> SchedulerTask:
        return self.add(SchedulerTask(task=task, **options, once=True))

    def delete(self, task: SchedulerTask) -> None:
        self.tasks.remove(task)

    def clear(self) -> None:
        self.tasks.clear()

    @property
    def enabled_tasks(self) -> Iterable[SchedulerTask]:
        return (t for t in self.tasks if t.enabled)

    @property
    def next_task(self) -> Optional[SchedulerTask]:
        return min(self.enabled_tasks, default=None)

    def _sanitize(self) -> None:
        alive = [t for t in self.tasks if t.should_live]
        self.tasks.clear()
        self.tasks.extend(alive)

    def next(self, block: bool = True) -> Optional[SchedulerTask]:
        task = self.next_task
        if task is None:
            return None
        try:
            return task.next(block=block)
        finally:
            if task.should_delete:
                self._sanitize()

    def all_once(self) -> None:
        for task in list(self.tasks):
            if task.once:
                task.next()
        self._sanitize()

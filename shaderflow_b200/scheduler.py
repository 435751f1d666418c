"""The export clock. The reference schedules its frame callback on a wall-clock task queue
(shaderflow/scheduler.py); an offline export only ever uses that queue in FREEWHEEL mode, where time is virtual:
the k-th call happens "at" the k-th due time and is handed the distance to the previous one. This module is
that virtual clock and nothing else — there is no sleeping, no wall clock, no frame skipping here, because this
backend has no realtime mode (scene.py: `realtime` is always False).

What has to match the reference bit for bit is the sequence of `dt` values, since scene time is their running
sum (scene.py:476-479) and everything downstream (audio read position, dynamics) keys off it. The reference's
arithmetic (scheduler.py:86-89,152-168), restated:

    due_0 = 0, previous_0 = -1/frequency
    call k:  dt_k = due_k - previous_k ;  previous_{k+1} = due_k ;  due_{k+1} = due_k (+ period until > due_k)

`sfb_frame_clock` (csrc/core.cu) computes the same sequence in bulk for the audio tracks; tests/test_host.py holds
both to the goldens produced by the reference's own SchedulerTask.

`scene.scheduler.new(...)`, `.once(...)`, `.next()` and `.clear()` keep their names because example scenes and
`ShaderScene.main` call them."""
from __future__ import annotations

import inspect
from typing import Any, Callable, Optional


class ClockedTask:
    """A callback on the virtual clock. `frequency` may be changed between calls (scene.next stores the
    scene's fps into its own task every frame, scene.py:472-473)."""

    __slots__ = ("task", "args", "kwargs", "frequency", "once", "enabled", "frameskip", "output",
                 "due", "previous", "_takes_dt", "_order")

    def __init__(self, task: Callable, *, frequency: float = 60.0, once: bool = False, frameskip: bool = True,
                 args: tuple = (), kwargs: Optional[dict] = None, order: int = 0, **_ignored: Any):
        # `freewheel`, `precise`, `context`, ... of the reference's constructor are accepted and have no meaning here
        self.task, self.args, self.kwargs = task, tuple(args), dict(kwargs or {})
        self.frequency, self.once, self.frameskip = float(frequency), bool(once), bool(frameskip)
        self.enabled, self.output = True, None
        self.due = 0.0
        self.previous = 0.0 - self.period
        self._takes_dt = "dt" in inspect.signature(task).parameters
        self._order = order

    @property
    def period(self) -> float:
        return 1.0/self.frequency

    fps = property(lambda self: self.frequency, lambda self, value: setattr(self, "frequency", float(value)))

    def sort_key(self) -> tuple:
        # one-shot tasks run before periodic ones; then earliest due; then creation order
        return (0 if self.once else 1, self.due, self._order)

    def fire(self) -> "ClockedTask":
        now = self.due
        if self._takes_dt:
            dt = now - self.previous
            self.kwargs["dt"] = dt if self.frameskip else min(dt, self.period)
        self.previous = now
        self.output = self.task(*self.args, **self.kwargs)
        if self.once:
            self.enabled = False
        else:
            due = self.due
            while due <= now:
                due += self.period
            self.due = due
        return self


class Scheduler:
    """Runs whichever enabled task is due first on the virtual clock"""
    Task = ClockedTask

    def __init__(self):
        self.tasks: list[ClockedTask] = []
        self._created = 0

    def _add(self, task: Callable, once: bool, options: dict) -> ClockedTask:
        self._created += 1
        entry = ClockedTask(task, once=once, order=self._created, **options)
        self.tasks.append(entry)
        return entry

    def new(self, task: Callable, **options) -> ClockedTask:
        return self._add(task, False, options)

    def once(self, task: Callable, **options) -> ClockedTask:
        return self._add(task, True, options)

    def delete(self, task: ClockedTask) -> None:
        self.tasks.remove(task)

    def clear(self) -> None:
        self.tasks.clear()

    def next(self, block: bool = True) -> Optional[ClockedTask]:
        ready = [t for t in self.tasks if t.enabled]
        if not ready:
            return None
        task = min(ready, key=ClockedTask.sort_key).fire()
        if task.once:
            self.tasks.remove(task)
        return task

    def all_once(self) -> None:
        for task in [t for t in self.tasks if t.once]:
            task.fire()
            self.tasks.remove(task)


SchedulerTask = ClockedTask

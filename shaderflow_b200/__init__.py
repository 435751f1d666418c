"""
shaderflow_b200 — B200-native backend of ShaderFlow's offline render hot loop.

The package mirrors the reference's Scene / Module / Texture / Variable Python API (same class names,
attributes and hooks) on top of `libsfb200.so` (hand-written sm_100a CUDA behind the C ABI of
include/sfb200.h). Call `shaderflow_b200.install_alias()` to make `import shaderflow` resolve here, so
scene files written against the reference (`from shaderflow.scene import ShaderScene`) run unchanged.

There is no OpenGL, no Triton and no CPU fallback in this package.
"""
from __future__ import annotations

import importlib
import logging
import os
import sys
from pathlib import Path

__version__ = "0.1.0"
__reference__ = "BrokenSource/ShaderFlow 0.11.3"

package = Path(__file__).parent
"""Path to the package directory"""

resources = package/"resources"
"""Kept for API compatibility (`shaderflow.resources`); built-in shaders live in csrc/scenes.cuh"""


class _Logger(logging.Logger):
    """stdlib stand-in for the reference's dearlog logger (shaderflow/__init__.py:1): same verbs,
    and like dearlog each call returns the message so `raise Error(logger.error(...))` works"""

    def _say(self, level: int, *parts) -> str:
        message = " ".join(str(p) for p in parts)
        if self.isEnabledFor(level):
            self._log(level, message, ())
        return message

    def info(self, *a, **k):  return self._say(logging.INFO, *a)
    def warn(self, *a, **k):  return self._say(logging.WARNING, *a)
    def error(self, *a, **k): return self._say(logging.ERROR, *a)
    def debug(self, *a, **k): return self._say(logging.DEBUG, *a)
    def minor(self, *a, **k): return self._say(logging.DEBUG, *a)
    def tip(self, *a, **k):   return self._say(logging.INFO, *a)
    def crit(self, *a, **k):  return self._say(logging.CRITICAL, *a)
    ok = note = info


logging.setLoggerClass(_Logger)
logger: _Logger = logging.getLogger("shaderflow_b200")  # type: ignore[assignment]
logging.setLoggerClass(logging.Logger)
if not logger.handlers:
    _handler = logging.StreamHandler()
    _handler.setFormatter(logging.Formatter("│ShaderFlow·B200│ %(levelname)-7s %(message)s"))
    logger.addHandler(_handler)
    logger.setLevel(os.environ.get("SHADERFLOW_LOGLEVEL", "WARNING").upper())


class _Directories:
    """The slice of platformdirs.PlatformDirs the examples touch (demo.py:25)"""
    @property
    def user_data_path(self) -> Path:
        try:
            from platformdirs import PlatformDirs
            return Path(PlatformDirs(appname="shaderflow", ensure_exists=True, opinion=True).user_data_path)
        except Exception:
            path = Path.home()/".local"/"share"/"shaderflow"
            path.mkdir(parents=True, exist_ok=True)
            return path


directories = _Directories()

# The reference pins numpy's BLAS to one thread (shaderflow/__init__.py:34); host-side numpy here is
# bookkeeping only, but keep the behaviour for code that relies on it
os.environ.setdefault("OMP_NUM_THREADS", "1")

_SUBMODULES = (
    "variable", "message", "resolution", "scheduler", "module", "dynamics", "keyboard", "frametimer",
    "camera", "texture", "shader", "exporting", "scene", "registry", "distributed",
    "audio", "audio.module", "audio.spectrogram", "audio.waveform", "piano", "piano.notes", "piano.module", "video",
)


def install_alias(name: str = "shaderflow") -> None:
    """Registers this package and its submodules under `name` in sys.modules, so user scenes written
    for the reference import the CUDA backend without edits"""
    me = sys.modules[__name__]
    existing = sys.modules.get(name)
    if existing is not None and existing is not me and not getattr(existing, "__sfb200_alias__", False):
        raise RuntimeError(f"a different '{name}' package is already imported")
    me.__sfb200_alias__ = True
    sys.modules[name] = me
    for sub in _SUBMODULES:
        sys.modules[f"{name}.{sub}"] = importlib.import_module(f"{__name__}.{sub}")

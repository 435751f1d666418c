"""
TEST INFRASTRUCTURE ONLY (oracle). Never imported by the product path.

Imports the *reference's own* numpy audio code (`/root/reference/shaderflow/audio`,
`dynamics.py`, `piano/notes.py`) in the build container, with the GUI / audio-device
third-party modules it drags in replaced by `MagicMock` stubs.

This only works where `/root/reference` exists (the build container). It is used by
`tests/golden/make_golden.py` to generate the committed golden vectors, and by the
`-m "not gpu"` tests that pin `oracle/audio_np.py` (the travelling numpy restatement)
against the reference itself.  The GPU box has no `/root/reference`: nothing run there
may call `load()`.
"""
from __future__ import annotations

import importlib
import os
import sys
from pathlib import Path
from types import SimpleNamespace
from unittest.mock import MagicMock

REFERENCE = Path(os.environ.get("SFB_REFERENCE", "/root/reference"))

# Modules the reference imports at module scope that are absent here
_STUBS = (
    "dearlog", "soundcard", "thefuzz", "thefuzz.process", "moderngl", "_moderngl",
    "imgui_bundle", "ordered_set", "watchdog", "watchdog.observers", "watchdog.events",
    "cyclopts", "parsenaut", "parsenaut._cyclopts", "turbopipe", "quaternion", "glfw",
    "moderngl_window", "pooch",
)


def available() -> bool:
    return (REFERENCE/"shaderflow"/"audio"/"spectrogram.py").exists()


def load() -> SimpleNamespace:
    """Returns a namespace with the reference classes of the audio path"""
    if not available():
        raise RuntimeError(f"Reference tree not found at {REFERENCE}")
    if "shaderflow_b200" in sys.modules and "shaderflow" in sys.modules:
        if getattr(sys.modules["shaderflow"], "__sfb200_alias__", False):
            raise RuntimeError("The 'shaderflow' alias of shaderflow_b200 is installed in this "
                "process; load the reference in a fresh interpreter")
    for name in _STUBS:
        sys.modules.setdefault(name, MagicMock())
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    if str(REFERENCE) not in sys.path:
        sys.path.insert(0, str(REFERENCE))
    mod   = importlib.import_module("shaderflow.audio.module")
    spec  = importlib.import_module("shaderflow.audio.spectrogram")
    wave  = importlib.import_module("shaderflow.audio.waveform")
    dyn   = importlib.import_module("shaderflow.dynamics")
    notes = importlib.import_module("shaderflow.piano.notes")
    res   = importlib.import_module("shaderflow.resolution")
    return SimpleNamespace(
        BrokenAudio=mod.BrokenAudio,
        root_mean_square=mod.root_mean_square,
        BrokenSpectrogram=spec.BrokenSpectrogram,
        SpectrogramWindow=spec.SpectrogramWindow,
        SpectrogramScale=spec.SpectrogramScale,
        SpectrogramInterpolation=spec.SpectrogramInterpolation,
        FourierMagnitude=spec.FourierMagnitude,
        FourierVolume=spec.FourierVolume,
        WaveformReducer=wave.WaveformReducer,
        DynamicNumber=dyn.DynamicNumber,
        PianoNote=notes.PianoNote,
        Resolution=res.Resolution,
    )

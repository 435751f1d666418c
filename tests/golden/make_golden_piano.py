"""
Generates tests/golden/piano_roll.npz by running the REFERENCE's own ShaderPiano.update
(/root/reference/shaderflow/piano/module.py:202-277, imported through oracle/ref_loader.py with the GUI
dependencies stubbed) on a seeded synthetic note list, frame by frame with the freewheel export clock
(frame 0: time 0, dt 0; then dt = 1/fps). Build container only:   python tests/golden/make_golden_piano.py

The module is driven without a ShaderScene: `update` is called on a plain object that carries the fields it
reads (scene.time / dt / realtime, tree, dynamics, globals) and texture stand-ins that record what is written.
"""
from __future__ import annotations

import importlib
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import ref_loader  # noqa: E402

R = ref_loader.load()
ref_piano = importlib.import_module("shaderflow.piano.module")
ShaderPiano, PianoNote = ref_piano.ShaderPiano, R.PianoNote


def synthetic_notes(seconds: float, seed: int = 5, channels: int = 4, rate: float = 16.0):
    """`channels` channels x `rate` notes/s in total, pitches 36-96, durations 0.1-1 s; channel by
    channel (the order a MIDI file's instruments are added in), plus glued / zero-length / overlapping cases"""
    rng = np.random.default_rng(seed)
    notes = []
    for channel in range(channels):
        n = int(seconds*rate/channels)
        starts = np.sort(rng.uniform(-0.5, seconds, n))
        for s in starts:
            notes.append((int(rng.integers(36, 97)), float(s), float(s + rng.uniform(0.1, 1.0)), channel, int(rng.integers(20, 128))))
    notes += [(60, 1.0, 1.5, 0, 90), (60, 1.5, 2.0, 1, 70), (60, 1.25, 1.26, 2, 50), (61, 2.0, 2.0, 0, 99), (62, 0.98, 1.0, 3, 64)]
    return notes


class Recorder:
    def __init__(self): self.data = None
    def write(self, data=None, **kw): self.data = np.array(data, copy=True)


def run(notes, n_frames: int, fps: float = 60.0):
    fields = {a.name: (a.default.factory() if hasattr(a.default, "factory") else a.default)
              for a in ShaderPiano.__attrs_attrs__ if a.name not in ("scene", "uuid")}
    scene = SimpleNamespace(time=0.0, dt=0.0, realtime=False)
    piano = SimpleNamespace(**fields)
    piano.scene = scene
    piano.keys_texture, piano.roll_texture, piano.channel_texture = Recorder(), Recorder(), Recorder()
    for name in ("_ranges", "_empty_keys", "_empty_roll", "notes_between", "update_global_ranges", "add_note",
                 "fluid_key_down", "fluid_key_up"):
        fn = getattr(ShaderPiano, name)
        setattr(piano, name, fn if isinstance(ShaderPiano.__dict__[name], staticmethod) else fn.__get__(piano))
    piano.lookup_time = piano.roll_time + piano.lookahead
    for (pitch, start, end, channel, velocity) in notes:
        piano.add_note(PianoNote(note=pitch, start=start, end=end, channel=channel, velocity=velocity))
    out = dict(roll=[], keys=[], chan=[], range=[], time=[], dt=[])
    for k in range(n_frames):
        ShaderPiano.update(piano)
        out["roll"].append(piano.roll_texture.data.astype(np.float32))
        out["keys"].append(np.array(piano.keys_texture.data, np.float32).reshape(-1))
        out["chan"].append(np.array(piano.channel_texture.data, np.float32).reshape(-1))
        out["range"].append(np.array(piano.note_range_dynamics.value, np.float32).copy())
        out["time"].append(scene.time); out["dt"].append(scene.dt)
        scene.dt = 1.0/fps                         # scene.py:476-479: time integrates after the modules ran
        scene.time += scene.dt
    return {k: np.array(v) for k, v in out.items()}, (piano.global_minimum_note, piano.global_maximum_note)


def main():
    notes = synthetic_notes(4.0)
    frames = list(range(0, 200, 1))
    res, (gmin, gmax) = run(notes, len(frames))
    keep = np.array([0, 1, 2, 30, 59, 60, 61, 75, 90, 119, 120, 150, 199])          # roll frames kept (each is 512 KB raw)
    np.savez_compressed(ROOT/"tests"/"golden"/"piano_roll.npz",
                        notes=np.array(notes, dtype=np.float64), keep=keep, roll=res["roll"][keep],
                        keys=res["keys"], chan=res["chan"], range=res["range"], time=res["time"], dt=res["dt"],
                        gmin=gmin, gmax=gmax)
    print("notes", len(notes), "frames", len(frames), "range", gmin, gmax,
          "max simultaneous", int((res["roll"][..., 1] > res["roll"][..., 0]).sum(axis=-1).max()),
          "file", (ROOT/"tests"/"golden"/"piano_roll.npz").stat().st_size)


if __name__ == "__main__":
    main()

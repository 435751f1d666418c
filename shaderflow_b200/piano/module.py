"""ShaderPiano — piano-roll producer. API mirror of the reference's shaderflow/piano/module.py:24-287
(same fields, `add_note` / `notes_between` / `normalize_velocities` / `pipeline`, same four textures), with
`update()` — a Python walk of a dict-of-dict-of-deque tree for every pitch on every frame plus three
texture uploads (module.py:202-277) — replaced by the GPU producer of csrc/piano.cu:

  once per export   sfb_piano_track     key-press targets, channel row and upcoming-pitch range of EVERY frame
                    sfb_dynamics_scan   the 128 key-press DynamicNumbers over frames (float32, numpy's op order)
                    the 2-lane note-range DynamicNumber runs on the host over the gathered targets
  per frame         sfb_piano_roll      the (128 x 256) roll texture straight into the texture's storage;
                    the keys / channel textures sample row k of the tracks (zero-copy bind)

Out of scope here: FluidSynth playback (module.py:289-328; realtime only). `load_midi` (module.py:166-196) reads
Standard MIDI Files natively (piano/midi.py) instead of through the third-party pretty_midi.
"""
from __future__ import annotations

import ctypes
from collections import deque
from collections.abc import Iterable
from pathlib import Path
from typing import Any, Optional

import numpy as np
from attrs import Factory, define

from shaderflow_b200 import _native as N
from shaderflow_b200 import logger
from shaderflow_b200.dynamics import DynamicNumber
from shaderflow_b200.module import ShaderModule
from shaderflow_b200.piano.notes import PianoNote
from shaderflow_b200.texture import ShaderTexture
from shaderflow_b200.variable import ShaderVariable, Uniform

MAX_CHANNELS = 32
MAX_ROLLING = 256
MAX_NOTE = 128


@define
class ShaderPiano(ShaderModule):
    name: str = "iPiano"
    tempo: deque = Factory(deque)
    keys_texture: ShaderTexture = None
    channel_texture: ShaderTexture = None
    roll_texture: ShaderTexture = None
    tempo_texture: ShaderTexture = None
    time_offset: float = 0
    roll_time: float = 2
    height: float = 0.275
    black_ratio: float = 0.6
    global_minimum_note: int = MAX_NOTE
    global_maximum_note: int = 0
    extra_keys: int = 6
    lookahead: float = 2
    release_before_end: float = 0.03

    key_press_dynamics: DynamicNumber = Factory(lambda: DynamicNumber(
        value=np.zeros(MAX_NOTE, dtype=np.float32), frequency=4, zeta=0.4, response=0, precision=0))
    note_range_dynamics: DynamicNumber = Factory(lambda: DynamicNumber(
        value=np.zeros(2, dtype=np.float32), frequency=0.05, zeta=1/(2**0.5), response=0))

    _order: list = Factory(list)
    """Every note in add_note order. The insertion index is part of the reference's iteration order (its
    dict-of-dict-of-deque tree, module.py:103-116, yields a pitch's notes by integer-second bucket, then by
    insertion): the GPU producer sorts on (pitch, bucket, insertion) and so does `notes_between` below —
    one flat list is the whole container."""
    _by_pitch: dict = Factory(dict)
    """pitch → indices into `_order`, built lazily for the host-side queries"""
    _device: Any = None
    _tracks: Any = None

    @property
    def lookup_time(self) -> float:
        return (self.roll_time + self.lookahead)

    # -- module --------------------------------------------------------------------------------
    def build(self):
        self.keys_texture    = ShaderTexture(scene=self.scene, name=f"{self.name}Keys").from_numpy(self._empty_keys())
        self.channel_texture = ShaderTexture(scene=self.scene, name=f"{self.name}Chan").from_numpy(self._empty_keys())
        self.roll_texture    = ShaderTexture(scene=self.scene, name=f"{self.name}Roll").from_numpy(self._empty_roll())
        self.tempo_texture   = ShaderTexture(scene=self.scene, name=f"{self.name}Tempo").from_numpy(np.zeros((100, 1, 2), np.float32))

    def setup(self):
        # a new export: fps, speed or the notes themselves may have changed since the tracks were computed
        self._tracks = None

    def _empty_keys(self) -> np.ndarray:
        return np.zeros((1, MAX_NOTE), dtype=np.float32)

    def _empty_roll(self) -> np.ndarray:
        return np.zeros((MAX_NOTE, MAX_ROLLING, 4), dtype=np.float32)

    # -- the score ----------------------------------------------------------------------------------
    def _invalidate(self) -> None:
        self._by_pitch.clear()
        self._device = self._tracks = None

    def clear(self):
        self._order.clear()
        self.global_minimum_note, self.global_maximum_note = MAX_NOTE, 0
        self._invalidate()

    def add_note(self, note: Optional[PianoNote]) -> None:
        if note is None:
            return
        self._order.append(note)
        self.global_minimum_note = min(self.global_minimum_note, note.note)
        self.global_maximum_note = max(self.global_maximum_note, note.note)
        self._invalidate()

    @property
    def notes(self) -> Iterable[PianoNote]:
        """Pitch by pitch in first-appearance order, bucket by bucket — a note spanning several integer seconds
        appears once per second it touches, as in the reference's tree walk (module.py:118-123)"""
        if not self._by_pitch:
            for index, note in enumerate(self._order):
                self._by_pitch.setdefault(note.note, []).append(index)
        for indices in self._by_pitch.values():
            spans = [(int(self._order[i].start), int(self._order[i].end), i) for i in indices]
            seconds: dict = {}                        # integer seconds of this pitch in first-touched order
            for lo, hi, _ in spans:
                for second in range(lo, hi + 1):
                    seconds.setdefault(second)
            for second in seconds:
                yield from (self._order[i] for lo, hi, i in spans if lo <= second <= hi)

    def __iter__(self) -> Iterable[PianoNote]:
        return self.notes

    @property
    def duration(self) -> float:
        return max((note.end for note in self._order), default=0)

    def notes_between(self, index: int, start: float, end: float) -> Iterable[PianoNote]:
        """Notes of pitch `index` that touch the integer seconds of [start, end] and have begun by `end`, each
        once, ordered by the first such second and then by insertion (module.py:131-141)"""
        if not self._by_pitch:
            list(self.notes)
        first, last = int(start), int(end)
        hits = []
        for i in self._by_pitch.get(index, ()):
            note = self._order[i]
            lo, hi = max(int(note.start), first), min(int(note.end), last)
            if lo <= hi and not (note.start > end):
                hits.append((lo, i))
        for _, i in sorted(hits):
            yield self._order[i]

    @property
    def maximum_velocity(self) -> Optional[int]:
        return max((note.velocity for note in self._order), default=None)

    @property
    def minimum_velocity(self) -> Optional[int]:
        return min((note.velocity for note in self._order), default=None)

    def normalize_velocities(self, minimum: int = 100, maximum: int = 100) -> None:
        """The reference computes an interpolated velocity and drops it (module.py:158-161):
        every note ends up at the midpoint. Kept, since exports depend on it."""
        midpoint = int((maximum + minimum)/2)
        for note in self._order:
            note.velocity = midpoint
        self._invalidate()

    def load_midi(self, path: Path):
        """module.py:166-196 without pretty_midi: piano/midi.py restates what it (and mido underneath) do. Notes are
        added instrument by instrument, the channel being the instrument's index; tempo changes fill the tempo
        texture, row k = (seconds, BPM)"""
        if not (path := Path(path)).exists():
            logger.warn(f"Input Midi file not found ({path})")
            return
        from shaderflow_b200.piano.midi import read_midi
        instruments, tempo = read_midi(path)
        for channel, instrument in enumerate(instruments):
            for (pitch, start, end, velocity) in instrument.notes:
                self.add_note(PianoNote(note=pitch, start=start, end=end, channel=channel, velocity=velocity))
        for when, bpm in tempo:
            self.tempo.append((when, bpm))
        rows = np.zeros((100, 1, 2), np.float32)
        for offset, (when, bpm) in enumerate(list(self.tempo)[:100]):
            rows[offset, 0] = (when, bpm)
        self.tempo_texture.write(rows)

    # -- GPU producer --------------------------------------------------------------------------------
    def _upload(self) -> None:
        """Notes → HBM, sorted by pitch (stable: insertion order inside a pitch) + the 129 pitch offsets"""
        import torch
        notes = sorted(enumerate(self._order), key=lambda item: (item[1].note, item[0]))
        packed = (N.PianoNote*max(1, len(notes)))()
        counts = np.zeros(MAX_NOTE + 1, dtype=np.int64)
        for slot, (order, note) in enumerate(notes):
            if not (0 <= int(note.note) < MAX_NOTE):
                raise RuntimeError(f"PianoNote pitch {note.note} outside 0..{MAX_NOTE - 1}")
            packed[slot] = N.PianoNote(float(note.start), float(note.end), int(note.note), int(note.channel), int(note.velocity), order)
            counts[int(note.note) + 1] += 1
        dev = f"cuda:{self.scene.device}"
        raw = np.frombuffer(bytes(packed), dtype=np.uint8).copy()
        self._device = dict(notes=torch.from_numpy(raw).to(dev),
                            offsets=torch.from_numpy(np.cumsum(counts).astype(np.int32)).to(dev),
                            overflow=torch.zeros(1, dtype=torch.int32, device=dev), count=len(notes))

    def prepare(self) -> None:
        """Everything update() derives from the tree besides the roll, for every frame of the export"""
        import torch
        scene = self.scene
        if self._device is None:
            self._upload()
        frames = scene.total_frames
        time, dt, _ = N.frame_clock(frames, scene.fps, scene.speed, 44100, 2, -1)
        dev = f"cuda:{scene.device}"
        time_d, dt_d = torch.from_numpy(time).to(dev), torch.from_numpy(dt).to(dev)
        keys = torch.zeros((frames, MAX_NOTE), dtype=torch.float32, device=dev)
        chan = torch.zeros((frames, MAX_NOTE), dtype=torch.float32, device=dev)
        upcoming = torch.zeros((frames, 2), dtype=torch.int32, device=dev)
        scene.cuda.piano_track(self._device["notes"], self._device["offsets"], time_d, frames, self.time_offset, self.roll_time,
                               self.lookup_time, self.release_before_end, self.global_minimum_note, self.global_maximum_note,
                               keys, chan, upcoming)
        d = self.key_press_dynamics
        scene.cuda.dynamics_scan(keys, MAX_NOTE, dt_d, frames, (d.frequency, d.zeta, d.response, d.precision))
        # the note range: 2 lanes, its value is a uniform → host, once per export (module.py:255-268)
        upcoming_h = upcoming.cpu().numpy()
        rng = self.note_range_dynamics
        rng.frequency = 0.5/self.lookup_time
        rng.set(np.zeros(2, dtype=np.float32))
        ranges = np.zeros((frames, 2), dtype=np.float32)
        for k in range(frames):
            if float(rng.value.sum()) == 0:
                rng.value[:] = (self.global_minimum_note, self.global_maximum_note)
            lo, hi = upcoming_h[k]
            rng.target[:] = ((lo, hi) if hi >= 0 else (self.global_minimum_note, self.global_maximum_note))
            rng.next(dt=abs(float(dt[k])))
            ranges[k] = rng.value
        self._tracks = dict(frames=frames, time=time, keys=keys, chan=chan, ranges=ranges,
                            key=self._tracks_key())

    def _tracks_key(self) -> tuple:
        scene = self.scene
        return (len(self._order), self.time_offset, self.roll_time, self.lookahead, self.release_before_end,
                scene.fps, scene.speed, scene.total_frames)

    def update(self):
        scene = self.scene
        if scene.cuda is None:
            return
        key = self._tracks_key()
        if self._tracks is None or self._tracks["frames"] != scene.total_frames or self._tracks["key"] != key:
            self.prepare()
        k = min(scene.frame_index, self._tracks["frames"] - 1)
        self.note_range_dynamics.value = self._tracks["ranges"][k].copy()
        if not scene.render_enabled:
            return                                   # a frame another rank shades
        time = float(self._tracks["time"][k]) + self.time_offset
        pointer, _ = self.roll_texture.get_box().texture.storage()
        scene.cuda.piano_roll(self._device["notes"], self._device["offsets"], time, self.roll_time, self.lookup_time,
                              self.global_minimum_note, self.global_maximum_note, pointer, self._device["overflow"])
        self.roll_texture.get_box().empty = False
        self.keys_texture.bind(self._tracks["keys"], self._tracks["keys"][k].data_ptr())
        self.channel_texture.bind(self._tracks["chan"], self._tracks["chan"][k].data_ptr())

    def pipeline(self) -> Iterable[ShaderVariable]:
        yield Uniform("int",   f"{self.name}GlobalMin",  self.global_minimum_note)
        yield Uniform("int",   f"{self.name}GlobalMax",  self.global_maximum_note)
        yield Uniform("vec2",  f"{self.name}Dynamic",    self.note_range_dynamics.value)
        yield Uniform("float", f"{self.name}RollTime",   self.roll_time)
        yield Uniform("float", f"{self.name}Extra",      self.extra_keys)
        yield Uniform("float", f"{self.name}Height",     self.height)
        yield Uniform("int",   f"{self.name}Limit",      MAX_ROLLING)
        yield Uniform("float", f"{self.name}BlackRatio", self.black_ratio)

    # -- FluidSynth (realtime playback only; module.py:289-328): not part of the offline path ----------
    fluidsynth: Any = None
    soundfont: Any = None

    def fluid_key_down(self, note: int, velocity: int = 127, channel: int = 0) -> None: ...
    def fluid_key_up(self, note: int, channel: int = 0) -> None: ...
    def fluid_all_notes_off(self) -> None: ...

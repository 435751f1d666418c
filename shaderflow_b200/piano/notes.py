"""PianoNote — midi index / name / frequency conversions with the API of the reference's shaderflow/piano/notes.py
(same constructors, static converters and properties). ShaderSpectrogram.from_notes places its bins on piano keys
with it; ShaderPiano stores its score as PianoNotes.

The conversions are plain module functions (equal temperament around A4 = midi 69); the class binds them."""
from __future__ import annotations

import functools
import math
from typing import Any

from attrs import define

PIANO_NOTES = "C C# D D# E F F# G G# A A# B".split()
_WHITE_CLASSES = frozenset((0, 2, 4, 5, 7, 9, 11))      # pitch classes of the white keys
_A4 = 69


@functools.lru_cache(maxsize=None)
def _name_of(index: int) -> str:
    octave, pitch_class = divmod(index, 12)
    return f"{PIANO_NOTES[pitch_class]}{octave - 1}"


def _index_of_name(name: str) -> int:
    letters, octave = name[:-1].upper(), int(name[-1])
    return PIANO_NOTES.index(letters) + 12*(octave + 1)


def _frequency_of(index: int, *, tuning: float = 440) -> float:
    return tuning * 2**((index - _A4)/12)


def _index_of_frequency(frequency: float, *, tuning: float = 440) -> int:
    return round(12*math.log2(frequency/tuning) + _A4)


@define(eq=False)
class PianoNote:
    note: int = 60
    start: float = 0.0
    end: float = 0.0
    channel: int = 0
    velocity: int = 100
    tuning: float = 440

    # -- static converters (same names as the reference) ---------------------------------------
    index_to_name = staticmethod(_name_of)
    name_to_index = staticmethod(_index_of_name)
    index_to_frequency = staticmethod(_frequency_of)
    frequency_to_index = staticmethod(_index_of_frequency)

    @staticmethod
    def name_to_frequency(name: str, *, tuning: float = 440) -> float:
        return _frequency_of(_index_of_name(name), tuning=tuning)

    @staticmethod
    def frequency_to_name(frequency: float, *, tuning: float = 440) -> str:
        return _name_of(_index_of_frequency(frequency, tuning=tuning))

    @staticmethod
    def is_white(note: int) -> bool:
        return (note % 12) in _WHITE_CLASSES

    @staticmethod
    def is_black(note: int) -> bool:
        return (note % 12) not in _WHITE_CLASSES

    # -- constructors ----------------------------------------------------------------------------
    @classmethod
    def from_index(cls, note: int, **kw) -> "PianoNote":
        return cls(note=note, **kw)

    @classmethod
    def from_name(cls, name: str, **kw) -> "PianoNote":
        return cls(note=_index_of_name(name), **kw)

    @classmethod
    def from_frequency(cls, frequency: float, **kw) -> "PianoNote":
        return cls(note=_index_of_frequency(frequency), **kw)

    @classmethod
    def get(cls, object: Any, **kw) -> "PianoNote":
        """PianoNote (updated in place with **kw) | midi index | name | frequency → PianoNote"""
        if isinstance(object, PianoNote):
            for key, value in kw.items():
                setattr(object, key, value)
            return object
        for kind, builder in ((int, cls.from_index), (str, cls.from_name), (float, cls.from_frequency)):
            if isinstance(object, kind):
                return builder(object, **kw)
        return cls(**kw)

    # -- views of the pitch / the time span ------------------------------------------------------
    @property
    def frequency(self) -> float:
        return _frequency_of(self.note, tuning=self.tuning)

    @frequency.setter
    def frequency(self, value: float):
        self.note = _index_of_frequency(value, tuning=self.tuning)

    @property
    def name(self) -> str:
        return _name_of(self.note)

    @name.setter
    def name(self, value: str):
        self.note = _index_of_name(value)

    @property
    def white(self) -> bool:
        return PianoNote.is_white(self.note)

    @property
    def black(self) -> bool:
        return PianoNote.is_black(self.note)

    @property
    def duration(self) -> float:
        return self.end - self.start

    @duration.setter
    def duration(self, value: float):
        self.end = self.start + value

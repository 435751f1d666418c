"""Standard MIDI Files → the notes and tempo changes `ShaderPiano.load_midi` adds (piano/module.py:166-196).

The reference hands the file to the third-party `pretty_midi` (which reads it with `mido`); neither is a dependency of
this backend nor installed in the build image, so this module restates what those two do on the way to
`midi.instruments[*].notes` and `midi.get_tempo_changes()` — pretty_midi 0.2.10's published algorithm:

  * delta times → absolute ticks per track; division = ticks per quarter note (SMPTE divisions are refused, as mido's
    `ticks_per_beat` has no meaning for them);
  * tempo map from the set_tempo events of track 0 ONLY (120 BPM until the first; a set_tempo at tick 0 replaces the
    default, a repeated tempo is ignored); tick → seconds piecewise linear, accumulated interval by interval in float64
    in pretty_midi's own order of operations, so the times are the same doubles;
  * note pairing per track and (channel, pitch): a note-off (or note-on with velocity 0) closes EVERY open note-on of
    that key that started on an earlier tick, and keeps the ones started on the same tick;
  * instruments are keyed (program at the time of the note-off, channel, track) and listed in order of first
    appearance; a note belongs to the instrument current when it ENDS; inside an instrument notes are in closing
    order. ShaderPiano numbers its channels by that instrument list (module.py:176).

PARITY UNPINNED: there is no pretty_midi / mido here to generate goldens from; tests/test_midi.py holds the parser to
files written by an independent SMF writer (running status, meta / sysex events, every rule above)."""
from __future__ import annotations

from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

MAX_TICK = 1e7          # pretty_midi's guard against corrupt files


@dataclass
class Event:
    tick: int
    kind: str                     # note_on, note_off, program_change, set_tempo, other
    channel: int = 0
    a: int = 0                    # note / program / tempo (µs per quarter)
    b: int = 0                    # velocity


@dataclass
class Instrument:
    program: int
    is_drum: bool
    notes: list = field(default_factory=list)       # (pitch, start_seconds, end_seconds, velocity), closing order


def _varlen(data: bytes, at: int) -> tuple[int, int]:
    value = 0
    for _ in range(4):
        byte = data[at]
        at += 1
        value = (value << 7) | (byte & 0x7F)
        if byte < 0x80:
            return value, at
    raise ValueError("variable-length quantity longer than four bytes")


def read_events(data: bytes) -> tuple[int, list[list[Event]]]:
    """→ (ticks per quarter note, per track: events with ABSOLUTE ticks)"""
    if data[:4] != b"MThd" or int.from_bytes(data[4:8], "big") < 6:
        raise ValueError("not a Standard MIDI File (no MThd header)")
    header = int.from_bytes(data[4:8], "big")
    tracks_declared = int.from_bytes(data[10:12], "big")
    division = int.from_bytes(data[12:14], "big")
    if division & 0x8000:
        raise ValueError("SMPTE time division is not supported (ticks per quarter note are)")
    at, tracks = 8 + header, []
    while at + 8 <= len(data) and len(tracks) < tracks_declared:
        name, size = data[at:at + 4], int.from_bytes(data[at + 4:at + 8], "big")
        body, at = data[at + 8:at + 8 + size], at + 8 + size
        if name != b"MTrk":
            continue                                         # alien chunks are skipped (SMF 1.0: "be prepared to encounter")
        events, tick, pos, status = [], 0, 0, 0
        while pos < len(body):
            delta, pos = _varlen(body, pos)
            tick += delta
            byte = body[pos]
            if byte == 0xFF:                                 # meta event: type, length, data
                kind = body[pos + 1]
                size, pos = _varlen(body, pos + 2)
                payload, pos = body[pos:pos + size], pos + size
                if kind == 0x51 and size == 3:
                    events.append(Event(tick, "set_tempo", a=int.from_bytes(payload, "big")))
                else:
                    events.append(Event(tick, "other"))
                continue                                     # (mido reads on past an end-of-track meta to the chunk's end)
            if byte in (0xF0, 0xF7):                         # sysex: length, data
                size, pos = _varlen(body, pos + 1)
                pos += size
                events.append(Event(tick, "other"))
                continue
            if byte & 0x80:
                status, pos = byte, pos + 1
            elif not status:
                raise ValueError("data byte without a running status")
            high, channel = status & 0xF0, status & 0x0F
            if high in (0xC0, 0xD0):                         # program change, channel pressure: one data byte
                first, pos = body[pos], pos + 1
                events.append(Event(tick, "program_change", channel, first) if high == 0xC0 else Event(tick, "other", channel))
            elif high in (0x80, 0x90, 0xA0, 0xB0, 0xE0):
                first, second, pos = body[pos], body[pos + 1], pos + 2
                kind = {0x80: "note_off", 0x90: "note_on"}.get(high, "other")
                events.append(Event(tick, kind, channel, first, second))
            else:                                            # system common / realtime inside a track: no data handled
                events.append(Event(tick, "other"))
        tracks.append(events)
    if not tracks:
        raise ValueError("the file has no MTrk chunk")
    return division, tracks


def tick_scales(tracks: list[list[Event]], resolution: int) -> list[tuple[int, float]]:
    """pretty_midi._load_tempo_changes: (tick, seconds per tick) from track 0's set_tempo events"""
    scales = [(0, 60.0/(120.0*resolution))]
    for event in tracks[0]:
        if event.kind != "set_tempo":
            continue
        if event.tick == 0:
            bpm = 6e7/event.a
            scales = [(0, 60.0/(bpm*resolution))]
        else:
            scale = 60.0/((6e7/event.a)*resolution)
            if scale != scales[-1][1]:
                scales.append((event.tick, scale))
    return scales


def tick_to_time(scales: list[tuple[int, float]], max_tick: int) -> np.ndarray:
    """pretty_midi._update_tick_to_time: seconds of every tick 0 … max_tick, float64"""
    max_tick = max(max_tick, max(tick for tick, _ in scales))
    table = np.zeros(max_tick + 1)
    last_end = 0.0
    for (start, scale), (end, _) in zip(scales[:-1], scales[1:]):
        table[start:end + 1] = last_end + scale*np.arange(end - start + 1)
        last_end = table[end]
    start, scale = scales[-1]
    table[start:] = last_end + scale*np.arange(max_tick + 1 - start)
    return table


def read_midi(path) -> tuple[list[Instrument], list[tuple[float, float]]]:
    """→ (instruments in pretty_midi's order, tempo changes as (seconds, BPM))"""
    try:
        resolution, tracks = read_events(Path(path).read_bytes())
    except IndexError:
        raise ValueError(f"'{path}': the MIDI file ends inside an event") from None
    scales = tick_scales(tracks, resolution)
    max_tick = max((event.tick for track in tracks for event in track), default=0) + 1
    if max_tick > MAX_TICK:
        raise ValueError(f"MIDI file has a largest tick of {max_tick}, it is likely corrupt")
    seconds = tick_to_time(scales, max_tick)
    instruments: dict[tuple[int, int, int], Instrument] = {}
    for index, track in enumerate(tracks):
        open_notes: dict[tuple[int, int], list[tuple[int, int]]] = {}
        program = [0]*16
        for event in track:
            if event.kind == "program_change":
                program[event.channel] = event.a
            elif event.kind == "note_on" and event.b > 0:
                open_notes.setdefault((event.channel, event.a), []).append((event.tick, event.b))
            elif event.kind == "note_off" or (event.kind == "note_on" and event.b == 0):
                key = (event.channel, event.a)
                if key not in open_notes:
                    continue                                   # a spurious note-off
                closing = [(start, velocity) for start, velocity in open_notes[key] if start != event.tick]
                keeping = [(start, velocity) for start, velocity in open_notes[key] if start == event.tick]
                for start, velocity in closing:
                    slot = (program[event.channel], event.channel, index)
                    if slot not in instruments:
                        instruments[slot] = Instrument(program[event.channel], event.channel == 9)
                    instruments[slot].notes.append((event.a, float(seconds[start]), float(seconds[event.tick]), velocity))
                if closing and keeping:
                    open_notes[key] = keeping
                else:
                    del open_notes[key]
    tempo = [(float(seconds[tick]), 60.0/(scale*resolution)) for tick, scale in scales]
    return list(instruments.values()), tempo

"""ShaderPiano.load_midi without pretty_midi (shaderflow_b200/piano/midi.py): files written by the small SMF writer
below — an independent statement of the format — come back as the notes, instruments and tempo map pretty_midi's
published algorithm gives. PARITY UNPINNED (no pretty_midi / mido in the image): the expectations are worked out here."""
import numpy as np
import pytest

from shaderflow_b200.piano import midi as M


def vlq(n: int) -> bytes:
    out = [n & 0x7F]
    while n > 0x7F:
        n >>= 7
        out.append((n & 0x7F) | 0x80)
    return bytes(reversed(out))


def track(events, running=False) -> bytes:
    """events: (delta, bytes of the message); running=True drops repeated status bytes of channel messages"""
    body, last = b"", None
    for delta, message in events:
        body += vlq(delta)
        if running and message[0] < 0xF0 and message[0] == last:
            body += message[1:]
        else:
            body += message
        last = message[0] if message[0] < 0xF0 else None
    body += vlq(0) + b"\xff\x2f\x00"
    return b"MTrk" + len(body).to_bytes(4, "big") + body


def smf(division: int, tracks, fmt: int = 1) -> bytes:
    return b"MThd" + (6).to_bytes(4, "big") + fmt.to_bytes(2, "big") + len(tracks).to_bytes(2, "big") + division.to_bytes(2, "big") + b"".join(tracks)


def on(ch, note, vel): return bytes([0x90 | ch, note, vel])
def off(ch, note, vel=64): return bytes([0x80 | ch, note, vel])
def program(ch, p): return bytes([0xC0 | ch, p])
def tempo(us): return b"\xff\x51\x03" + us.to_bytes(3, "big")
def name(text): return b"\xff\x03" + vlq(len(text)) + text


def test_variable_length_quantities_and_header_checks():
    for n in (0, 0x7F, 0x80, 0x3FFF, 0x4000, 0x0FFFFFFF):
        assert M._varlen(vlq(n) + b"\x00", 0) == (n, len(vlq(n)))
    with pytest.raises(ValueError, match="MThd"):
        M.read_events(b"RIFF" + bytes(20))
    with pytest.raises(ValueError, match="SMPTE"):
        M.read_events(smf(0xE728, [track([])]))
    with pytest.raises(ValueError, match="no MTrk"):
        M.read_events(smf(96, []))


def test_tempo_map_is_track_zeros_and_accumulates_like_pretty_midi(tmp_path):
    """120 BPM until told otherwise; a set_tempo at tick 0 replaces the default; a repeated tempo is no change; tempo
    events on other tracks are ignored; seconds accumulate interval by interval"""
    conductor = track([(0, name(b"conductor")), (0, tempo(500000)), (480, tempo(250000)), (240, tempo(250000)), (240, tempo(1000000))])
    other = track([(100, tempo(100000)), (0, on(0, 60, 90)), (1820, off(0, 60))])
    (tmp_path/"t.mid").write_bytes(smf(480, [conductor, other]))
    instruments, changes = M.read_midi(tmp_path/"t.mid")
    assert [round(b, 9) for _, b in changes] == [120.0, 240.0, 60.0]
    assert [t for t, _ in changes] == [0.0, 0.5, 0.5 + 480*(60.0/(240.0*480))]
    (pitch, start, end, velocity), = instruments[0].notes
    scale = [60.0/(bpm*480) for bpm in (120.0, 240.0, 60.0)]
    assert (pitch, velocity) == (60, 90) and start == 100*scale[0]
    assert end == (480*scale[0] + 480*scale[1]) + scale[2]*(1920 - 960)
    (tmp_path/"cut.mid").write_bytes(smf(480, [conductor, other])[:-6])
    with pytest.raises(ValueError, match="ends inside"):
        M.read_midi(tmp_path/"cut.mid")
    # no set_tempo at all: 120 BPM
    (tmp_path/"plain.mid").write_bytes(smf(96, [track([(0, on(3, 40, 1)), (96, off(3, 40))])], fmt=0))
    instruments, changes = M.read_midi(tmp_path/"plain.mid")
    assert changes == [(0.0, 120.0)] and instruments[0].notes == [(40, 0.0, 0.5, 1)]


def test_note_pairing_instruments_and_running_status(tmp_path):
    """One note-off closes every earlier note-on of its key and keeps the one from its own tick; velocity-0 note-ons are
    note-offs; spurious note-offs are ignored; a note belongs to the program current when it ENDS; instruments appear in
    order of their first closed note; channel 9 is drums; running status and sysex / meta events in between"""
    events = [
        (0, program(0, 5)), (0, on(0, 60, 100)), (10, on(0, 60, 80)),       # two open C4s
        (0, b"\xf0" + vlq(3) + b"\x7e\x7f\xf7"), (0, name(b"lead")),
        (10, off(0, 61)),                                                    # spurious (tick 20)
        (0, on(9, 36, 127)),
        (20, on(0, 60, 0)),                                                  # tick 40: closes both C4s, under program 5
        (0, on(0, 62, 90)),                                                  # a D4 from tick 40
        (10, on(0, 62, 70)),                                                 # and one from tick 50
        (0, program(0, 7)),
        (0, off(0, 62)),                                                     # tick 50: closes the first under program 7, keeps the second
        (10, off(9, 36)),                                                    # tick 60
        (10, off(0, 62)),                                                    # tick 70: the second D4
        (0, on(1, 72, 50)), (5, on(1, 72, 0)),                               # tick 70 .. 75
        (0, on(2, 50, 60)), (0, off(2, 50)),                                 # opened and closed on one tick: dropped, no instrument
    ]
    for running in (False, True):
        (tmp_path/"n.mid").write_bytes(smf(100, [track([(0, tempo(600000))]), track(events, running=running)]))
        instruments, _ = M.read_midi(tmp_path/"n.mid")
        s = 60.0/((6e7/600000)*100)
        assert [(i.program, i.is_drum) for i in instruments] == [(5, False), (7, False), (0, True), (0, False)]
        assert instruments[0].notes == [(60, 0.0, 40*s, 100), (60, 10*s, 40*s, 80)]
        assert instruments[1].notes == [(62, 40*s, 50*s, 90), (62, 50*s, 70*s, 70)]
        assert instruments[2].notes == [(36, 20*s, 60*s, 127)]
        assert instruments[3].notes == [(72, 70*s, 75*s, 50)]


def test_shader_piano_loads_a_midi_file(tmp_path):
    import examples.demo as demo
    from shaderflow.piano import ShaderPiano
    right = track([(0, program(0, 0)), (0, on(0, 64, 90)), (240, off(0, 64)), (0, on(0, 67, 70)), (240, off(0, 67))])
    left = track([(0, program(1, 32)), (0, on(1, 40, 60)), (480, off(1, 40))])
    (tmp_path/"song.mid").write_bytes(smf(480, [track([(0, tempo(500000)), (480, tempo(400000))]), right, left]))
    scene = demo.Basic(backend="dry"); scene.initialize()
    piano = ShaderPiano(scene=scene)
    piano.load_midi(tmp_path/"missing.mid")                                  # warns, like the reference
    piano.load_midi(tmp_path/"song.mid")
    notes = [(n.note, n.start, n.end, n.channel, n.velocity) for n in piano.notes]
    assert notes == [(64, 0.0, 0.25, 0, 90), (67, 0.25, 0.5, 0, 70), (40, 0.0, 0.5, 1, 60)]
    assert list(piano.tempo) == [(0.0, 120.0), (0.5, 150.0)]
    rows = np.frombuffer(piano.tempo_texture.get_box().data, np.float32).reshape(100, 2)
    assert rows[:3].tolist() == [[0.0, 120.0], [0.5, 150.0], [0.0, 0.0]]

"""oracle/piano_np.py against golden vectors produced by the reference's own ShaderPiano.update
(tests/golden/make_golden_piano.py), plus the host-side API of shaderflow_b200.piano.ShaderPiano."""
import numpy as np
import pytest

from oracle import piano_np as P


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(golden_dir/"piano_roll.npz")


def test_piano_oracle_matches_reference_update(gold):
    notes = [tuple(row) for row in gold["notes"]]
    assert notes == [tuple(float(x) for x in n) for n in P.synthetic_notes(4.0)]
    piano = P.Piano(notes)
    assert (piano.gmin, piano.gmax) == (int(gold["gmin"]), int(gold["gmax"]))
    keep = {int(k): i for i, k in enumerate(gold["keep"])}
    for k in range(len(gold["time"])):
        out = piano.frame(float(gold["time"][k]), float(gold["dt"][k]))
        assert np.array_equal(out["keys"], gold["keys"][k]), k             # bit-exact float32 recurrences
        assert np.array_equal(out["chan"], gold["chan"][k]), k
        assert np.array_equal(out["range"], gold["range"][k]), k
        if k in keep:
            assert np.array_equal(out["roll"], gold["roll"][keep[k]]), k
    assert gold["keys"].max() > 50 and (gold["chan"] >= 0).any()


def test_piano_module_api_without_gpu():
    """add_note / notes_between / normalize_velocities / pipeline mirror piano/module.py"""
    import examples.demo as demo
    from shaderflow_b200.piano import PianoNote, ShaderPiano
    scene = demo.Basic(backend="dry"); scene.initialize()
    piano = ShaderPiano(scene=scene)
    for (pitch, start, end, channel, velocity) in P.synthetic_notes(4.0):
        piano.add_note(PianoNote(note=pitch, start=start, end=end, channel=channel, velocity=velocity))
    ref = P.Piano(P.synthetic_notes(4.0))
    assert (piano.global_minimum_note, piano.global_maximum_note) == (ref.gmin, ref.gmax)
    assert abs(piano.duration - max(n[2] for n in P.synthetic_notes(4.0))) < 1e-12
    for midi in range(ref.gmin, ref.gmax + 1):
        for t0 in (0.0, 0.95, 1.2, 2.999, 3.5):
            got = [(n.note, n.start, n.end, n.channel, n.velocity) for n in piano.notes_between(midi, t0, t0 + piano.lookup_time)]
            want = [n[:5] for n in ref.notes_between(midi, t0, t0 + ref.lookup_time)]
            assert got == want, (midi, t0)
    # iteration order of the whole score = the reference's tree walk (pitch, then integer second, then insertion)
    walk = [n[:5] for block in ref.tree.values() for notes in block.values() for n in notes]
    assert [(n.note, n.start, n.end, n.channel, n.velocity) for n in piano] == walk
    names = {v.name: v.value for v in piano.pipeline()}
    assert names["iPianoLimit"] == 256 and names["iPianoRollTime"] == 2 and names["iPianoGlobalMin"] == ref.gmin
    assert [t.name for t in (piano.keys_texture, piano.channel_texture, piano.roll_texture, piano.tempo_texture)] == \
        ["iPianoKeys", "iPianoChan", "iPianoRoll", "iPianoTempo"]
    assert piano.roll_texture.size == (256, 128) and piano.keys_texture.size == (128, 1)
    piano.normalize_velocities(minimum=40, maximum=100)
    assert {n.velocity for n in piano.notes} == {70}                     # the reference's dropped interpolation (module.py:163-166)
    piano.load_midi("no_such_song.mid")                                   # warns and returns, like the reference (tests/test_midi.py reads real files)

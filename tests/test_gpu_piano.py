"""GPU parity of the piano-roll producer (csrc/piano.cu) against oracle/piano_np.py — itself pinned
to the reference's own ShaderPiano.update by tests/test_oracle_piano.py — and against the golden vectors directly;
then piano.frag and the PianoRoll scene end to end."""
import numpy as np
import pytest

from oracle import glsl_np as G
from oracle import piano_np as P
from tests.helpers import native_textures, native_uniforms

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def ctx():
    from shaderflow_b200 import _native as N
    c = N.Context(0)
    yield c
    c.destroy()


def upload(notes):
    from shaderflow_b200 import _native as N
    order = sorted(range(len(notes)), key=lambda i: (int(notes[i][0]), i))
    packed = (N.PianoNote*max(1, len(notes)))()
    counts = np.zeros(129, np.int64)
    for slot, i in enumerate(order):
        pitch, start, end, channel, velocity = notes[i]
        packed[slot] = N.PianoNote(float(start), float(end), int(pitch), int(channel), int(velocity), i)
        counts[int(pitch) + 1] += 1
    raw = torch.from_numpy(np.frombuffer(bytes(packed), np.uint8).copy()).cuda()
    return raw, torch.from_numpy(np.cumsum(counts).astype(np.int32)).cuda()


def gpu_tracks(ctx, notes, times, dts, ref: P.Piano):
    F = len(times)
    dev_notes, offsets = upload(notes)
    time_d, dt_d = torch.from_numpy(np.asarray(times, np.float64)).cuda(), torch.from_numpy(np.asarray(dts, np.float64)).cuda()
    target = torch.zeros((F, 128), device="cuda"); chan = torch.zeros((F, 128), device="cuda")
    upcoming = torch.zeros((F, 2), dtype=torch.int32, device="cuda")
    ctx.piano_track(dev_notes, offsets, time_d, F, ref.time_offset, ref.roll_time, ref.lookup_time, ref.release, ref.gmin, ref.gmax,
                    target, chan, upcoming)
    keys = target.clone()
    ctx.dynamics_scan(keys, 128, dt_d, F, (4.0, 0.4, 0.0, 0.0))
    ctx.sync()
    return dict(target=target.cpu().numpy(), chan=chan.cpu().numpy(), upcoming=upcoming.cpu().numpy(), keys=keys.cpu().numpy(),
                notes=dev_notes, offsets=offsets)


def gpu_roll(ctx, tracks, time, ref: P.Piano):
    roll = torch.full((128, 256, 4), 7.0, device="cuda")                 # stale contents must be cleared
    overflow = torch.zeros(1, dtype=torch.int32, device="cuda")
    ctx.piano_roll(tracks["notes"], tracks["offsets"], time + ref.time_offset, ref.roll_time, ref.lookup_time, ref.gmin, ref.gmax, roll, overflow)
    ctx.sync()
    assert int(overflow.item()) == 0
    return roll.cpu().numpy()


def test_piano_kernels_match_reference_golden(ctx, golden_dir):
    gold = np.load(golden_dir/"piano_roll.npz")
    notes = [tuple(row) for row in gold["notes"]]
    ref = P.Piano(notes)
    tr = gpu_tracks(ctx, notes, gold["time"], gold["dt"], ref)
    assert np.array_equal(tr["keys"], gold["keys"])                        # float32 recurrence, bit-exact
    assert np.array_equal(tr["chan"], gold["chan"])
    for slot, k in enumerate(gold["keep"]):
        assert np.array_equal(gpu_roll(ctx, tr, float(gold["time"][k]), ref), gold["roll"][slot]), k


@pytest.mark.parametrize("seed,seconds,rate,offset", [(1, 6.0, 40.0, 0.0), (2, 3.0, 400.0, 0.25), (3, 2.0, 6000.0, -0.4)])
def test_piano_kernels_match_oracle_on_dense_scores(ctx, seed, seconds, rate, offset):
    """denser scores: many simultaneous notes per pitch (slot order = the reference's bucket / insertion order),
    glued and zero-length notes, a time offset, more than 256 visited notes on one pitch"""
    notes = P.synthetic_notes(seconds, seed=seed, channels=6, rate=rate)
    if seed == 3:
        notes += [(64, 0.001*i, 0.001*i + 0.5, i % 5, 30 + i % 90) for i in range(400)]     # > MAX_ROLLING on one key
    ref = P.Piano(notes, time_offset=offset)
    fps = 50.0
    times = [0.0] + [k/fps for k in range(1, 40)]
    dts = [0.0] + [1/fps]*39
    tr = gpu_tracks(ctx, notes, times, dts, ref)
    for k, (t, dt) in enumerate(zip(times, dts)):
        out = ref.frame(t, dt)
        assert np.array_equal(tr["target"][k], out["key_target"]), k
        assert np.array_equal(tr["chan"][k], out["chan"]), k
        assert tuple(tr["upcoming"][k]) == tuple(out["upcoming"]), k
        assert np.array_equal(tr["keys"][k], out["keys"]), k
        if k % 6 == 0:
            assert np.array_equal(gpu_roll(ctx, tr, t, ref), out["roll"]), k


def test_piano_fragment_matches_oracle(ctx):
    from shaderflow_b200 import _native as N
    W, H = 320, 180
    notes = P.synthetic_notes(4.0)
    ref = P.Piano(notes)
    out = None
    for k in range(70):
        out = ref.frame(k/60 if k else 0.0, 1/60 if k else 0.0)
    time = 69/60
    tex = dict(iPianoKeys=G.Texture(out["keys"].reshape(1, 128, 1).copy()), iPianoChan=G.Texture(out["chan"].reshape(1, 128, 1).copy()),
               iPianoRoll=G.Texture(out["roll"].copy()))
    u = G.Uniforms(iTime=time, iResolution=(W, H), iWantAspect=W/H,
                   extra=dict(iPianoDynamic=tuple(out["range"]), iPianoExtra=6.0, iPianoHeight=0.275, iPianoBlackRatio=0.6,
                              iPianoRollTime=2.0, iPianoLimit=256))
    want = P.frag_piano(u, G.varyings(u, W, H), tex)
    sid = N.scene_lookup("piano"); info = N.scene_info(sid)
    nt = native_textures(ctx, tex)
    target = N.Texture(ctx, W, H, 4, N.DTYPE_U8)
    ctx.render_target(sid, native_uniforms(u, info), [nt[name] for name in info["samplers"]], target)
    ctx.sync()
    got = target.read()
    d = np.abs(got[..., :3].astype(int) - G.to_unorm8(want)[..., :3].astype(int))
    assert (d <= 1).mean() > 0.999 and (d == 0).mean() > 0.97, ((d <= 1).mean(), (d == 0).mean())
    assert len(np.unique(got[..., :3].reshape(-1, 3), axis=0)) > 20          # keyboard, roll background and coloured notes


def test_piano_scene_end_to_end():
    """PianoRoll through the public API: the frames equal piano.frag evaluated on the oracle's textures"""
    from examples.demo import PianoRoll
    PianoRoll.notes = P.synthetic_notes(4.0)
    try:
        scene = PianoRoll(device=0); scene.initialize()
        W, H, n = 160, 90, 40
        got = {}
        def grab(index, pointer):
            scene.cuda.sync(); got[index] = scene.frame_tensor.cpu().numpy().copy()
        scene.main(width=W, height=H, ssaa=1, subsample=1, time=n/60, fps=60.0, on_frame=grab)
    finally:
        PianoRoll.notes = None
    ref = P.Piano(P.synthetic_notes(4.0))
    for k in range(n):
        t = k/60 if k else 0.0
        out = ref.frame(t, 1/60 if k else 0.0)
        if k not in (0, 17, 39):
            continue
        tex = dict(iPianoKeys=G.Texture(out["keys"].reshape(1, 128, 1).copy()), iPianoChan=G.Texture(out["chan"].reshape(1, 128, 1).copy()),
                   iPianoRoll=G.Texture(out["roll"].copy()))
        u = G.Uniforms(iTime=t, iResolution=(W, H), iWantAspect=W/H,
                       extra=dict(iPianoDynamic=tuple(out["range"]), iPianoExtra=6.0, iPianoHeight=0.275, iPianoBlackRatio=0.6,
                                  iPianoRollTime=2.0, iPianoLimit=256))
        screen = G.to_unorm8(P.frag_piano(u, G.varyings(u, W, H), tex))
        want = G.to_unorm8(G.final_pass(screen, W, H, 1))
        d = np.abs(got[k].astype(int) - want.astype(int))
        assert (d <= 1).mean() > 0.998, (k, (d <= 1).mean())

"""GPU parity of the multi-program / multi-layer / temporal scenes (MultiShader, Multipass,
MotionBlur, Dynamics, Audio, Life) against oracle/glsl_np.py — pass by pass through sfb_render_target, then
whole scenes through the public API against an oracle replay of the reference's texture matrix semantics
(render into row 0 layer by layer, roll, final.glsl reads `iScreen` = matrix[0][last layer])."""
from collections import deque

import numpy as np
import pytest

from oracle import glsl_np as G
from tests.helpers import native_textures, native_uniforms

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

W, H = 160, 90


@pytest.fixture(scope="module")
def ctx():
    from shaderflow_b200 import _native as N
    c = N.Context(0)
    yield c
    c.destroy()


def background(size=(120, 68)):
    return G.Texture(np.flipud(G.synthetic_background(*size)).copy(), linear=True, repeat_x=True, repeat_y=True)


def gpu_pass(ctx, scene, u, tex, w, h, comps=4, dtype=np.uint8, linear=True):
    """One pass of `scene` into a w x h target texture of the given format → (h, w, comps) array"""
    from shaderflow_b200 import _native as N
    sid = N.scene_lookup(scene); info = N.scene_info(sid)
    nt = native_textures(ctx, tex)
    samplers = [nt[name] for name in info["samplers"] if name in nt]
    target = N.Texture(ctx, w, h, comps, N.DTYPE_U8 if dtype == np.uint8 else N.DTYPE_F32, linear=linear)
    ctx.render_target(sid, native_uniforms(u, info), samplers, target)
    ctx.sync()
    out = target.read()[..., :comps]
    for t in nt.values(): t.destroy()
    target.destroy()
    return out


def uniforms(**kw):
    extra = kw.pop("extra", {})
    return G.Uniforms(iTime=0.7, iTau=0.2, iResolution=(W, H), iWantAspect=W/H, extra=dict(extra), **kw)


def check_u8(got, ref_f32, frac=0.999):
    d = np.abs(got[..., :3].astype(int) - G.to_unorm8(ref_f32)[..., :3].astype(int))
    assert (d <= 1).mean() >= frac and (d == 0).mean() >= frac - 0.03, ((d <= 1).mean(), (d == 0).mean(), d.max())


def test_multishader_passes(ctx):
    u = uniforms(); f = G.varyings(u, W, H)
    child = G.frag_multishader_child(u, f, {})
    got = gpu_pass(ctx, "multishader_child", u, {}, W, H)
    check_u8(got, child, 1.0)
    assert (got[..., 3] == 255).all()
    tex = dict(child=G.Texture(G.to_unorm8(child), linear=True))
    check_u8(gpu_pass(ctx, "multishader", u, tex, W, H), G.frag_multishader(u, f, tex))


def test_multipass_layers(ctx):
    u = uniforms(); f = G.varyings(u, W, H)
    tex = dict(background=background())
    layer0 = G.frag_multipass(u, f, tex)
    check_u8(gpu_pass(ctx, "multipass", u, dict(tex, iScreen0x0=G.Texture(np.zeros((H, W, 4), np.uint8))), W, H), layer0)
    u.iLayer = 1
    tex["iScreen0x0"] = G.Texture(G.to_unorm8(layer0), linear=True)
    got = gpu_pass(ctx, "multipass", u, tex, W, H)
    ref = G.frag_multipass(u, f, tex)
    check_u8(got, ref, 0.998)
    # left half inverts red, right half is the 63-tap blur: both differ from layer 0
    assert np.abs(got[:, :W//2, 0].astype(int) - (255 - G.to_unorm8(layer0)[:, :W//2, 0].astype(int))).max() <= 1


@pytest.mark.parametrize("temporal", [1, 4, 10])
def test_motionblur_layers(ctx, temporal):
    rng = np.random.default_rng(temporal)
    u = uniforms(extra=dict(iScreenTemporal=temporal), iCameraZoom=0.9); f = G.varyings(u, W, H)
    tex = dict(background=background())
    history = {f"iScreen{t}x0": G.Texture(rng.integers(0, 256, (H, W, 4), dtype=np.uint8), linear=True) for t in range(temporal)}
    check_u8(gpu_pass(ctx, "motionblur", u, dict(tex, **history), W, H), G.frag_motionblur(u, f, tex))
    u.iLayer = 1
    check_u8(gpu_pass(ctx, "motionblur", u, dict(tex, **history), W, H), G.frag_motionblur(u, f, dict(tex, **history)))


def test_dynamics_and_audio_passes(ctx):
    for value in (0.0, 0.37, 1.0):
        u = uniforms(extra=dict(iShaderDynamics=value, iAudioVolume=value)); f = G.varyings(u, W, H)
        tex = dict(background=background())
        check_u8(gpu_pass(ctx, "dynamics", u, tex, W, H), G.frag_dynamics(u, f, tex))
        check_u8(gpu_pass(ctx, "audio", u, {}, W, H), G.frag_audio(u, f, {}), 1.0)


def test_life_passes(ctx):
    """simulation.glsl into a 1-component float32 NEAREST texture: exact integer states; visuals.glsl on 5 states"""
    rng = np.random.default_rng(3)
    lw, lh = 48, 27
    state = G.Texture(rng.integers(0, 2, (lh, lw, 1)).astype(np.float32), linear=False)
    for frame in (0, 6, 7):                                  # rule frames and a copy frame (period 6)
        u = uniforms(iFrame=frame, extra=dict(iLifePeriod=6, iLifeSize=(lw, lh)))
        ref = G.frag_life_simulation(u, G.varyings(u, lw, lh), dict(iLife1x0=state))
        got = gpu_pass(ctx, "life_simulation", u, dict(iLife1x0=state), lw, lh, comps=1, dtype=np.float32, linear=False)
        assert np.array_equal(got[..., 0], ref[..., 0]), frame
    states = {f"iLife{k}x0": G.Texture(rng.integers(0, 2, (lh, lw, 1)).astype(np.float32), linear=False) for k in range(5)}
    u = uniforms(); f = G.varyings(u, W, H)
    check_u8(gpu_pass(ctx, "life_visuals", u, states, W, H), G.frag_life_visuals(u, f, states))


# ---------------------------------------------------------------------------------------------- #
# whole scenes through the public API against an oracle replay of the texture matrix

def replay(passes, n_frames, w, h, ssaa, subsample, fps=60.0):
    """passes: list of programs in render order, each dict(name, scene, layers, temporal, size, fmt, extra(frame)→dict,
    static textures). Returns the final RGB8 frames the reference semantics produce."""
    mats = {}
    for p in passes:
        pw, ph = p.get("size") or (int(w*ssaa), int(h*ssaa))
        zero = (lambda: np.zeros((ph, pw, 4), np.uint8)) if p.get("fmt", "u8") == "u8" else (lambda: np.zeros((ph, pw, 1), np.float32))
        mats[p["name"]] = deque(deque(zero() for _ in range(p["layers"])) for _ in range(p["temporal"]))
        if "init" in p:
            p["init"](mats[p["name"]])
    frames = []
    for k in range(n_frames):
        time = k/fps if k else 0.0
        for p in passes:
            mat = mats[p["name"]]
            pw, ph = p.get("size") or (int(w*ssaa), int(h*ssaa))
            for layer in range(p["layers"]):
                tex = dict(p.get("static", {}))
                for name, m in mats.items():
                    lin = next(q for q in passes if q["name"] == name).get("linear", True)
                    for t, row in enumerate(m):
                        for l, data in enumerate(row):
                            tex[f"{name}{t}x{l}"] = G.Texture(data, linear=lin)
                    tex[name] = G.Texture(m[0][len(m[0]) - 1], linear=lin)
                u = G.Uniforms(iTime=time, iResolution=(w, h), iWantAspect=w/h, iFrame=k, iLayer=layer, iSSAA=ssaa,
                               extra=p["extra"](k, time))
                out = G.SCENES[p["scene"]](u, G.varyings(u, pw, ph), tex)
                mat[0][layer] = G.to_unorm8(out) if p.get("fmt", "u8") == "u8" else out[..., :1].astype(np.float32).copy()
            mat.rotate(1)
        screen = mats["iScreen"][0][len(mats["iScreen"][0]) - 1]
        frames.append(G.to_unorm8(G.final_pass(screen, w, h, subsample)))
    return frames


def run_scene(scene, n_frames, w, h, **flags):
    got = {}
    def grab(index, pointer):
        scene.cuda.sync()
        got[index] = scene.frame_tensor.cpu().numpy().copy() if scene._frame_target is None else None
    scene.main(width=w, height=h, time=n_frames/60, fps=60.0, on_frame=grab, **flags)
    return [got[k] for k in range(n_frames)]


def test_life_scene_end_to_end():
    """Life: two programs, a 10-deep float32 history, the reference's roll-then-sample order (visuals read the
    OLDEST slot as iLife0x0 — texture.py:301-304 rolls before the parent renders)"""
    from examples.demo import Life
    Life.life_seed = 11
    try:
        scene = Life(device=0); scene.initialize()
        # 100 x 56: NEAREST taps of the 192 x 108 state stay clear of texel edges (at 96 x 54 every tap sits exactly on
        # one, where a 1-ulp difference between FMA-contracted and numpy arithmetic picks the other cell)
        w, h, n = 100, 56, 14
        got = run_scene(scene, n, w, h, ssaa=1, subsample=1)
    finally:
        Life.life_seed = None
    rng = np.random.default_rng(11)
    first = rng.integers(0, 2, (192, 108)).astype(bool).astype(np.float32)
    def init(mat):
        mat[1][0] = first.reshape(-1).reshape(108, 192, 1).copy()           # raw bytes of the (192, 108) array as 108 rows of 192
    passes = [
        dict(name="iLife", scene="life_simulation", layers=1, temporal=10, size=(192, 108), fmt="f32", linear=False, init=init,
             extra=lambda k, t: dict(iLifePeriod=6, iLifeSize=(192, 108))),
        dict(name="iScreen", scene="life_visuals", layers=1, temporal=1, extra=lambda k, t: {}),
    ]
    ref = replay(passes, n, w, h, 1.0, 1)
    for k in range(n):
        d = np.abs(got[k].astype(int) - ref[k].astype(int))
        assert (d <= 1).mean() > 0.999, (k, (d <= 1).mean())
    assert any(np.abs(got[k].astype(int) - got[0].astype(int)).max() > 0 for k in range(1, n))   # the picture evolves


def test_motionblur_and_multishader_scenes_end_to_end():
    from examples.demo import MotionBlur, MultiShader
    bg = G.synthetic_background(96, 54)
    w, h, n = 96, 54, 12
    MotionBlur.background = bg
    try:
        scene = MotionBlur(device=0); scene.initialize()
        got = run_scene(scene, n, w, h, ssaa=1, subsample=1)
    finally:
        MotionBlur.background = None
    static = dict(background=G.Texture(np.flipud(bg).copy(), linear=True))
    passes = [dict(name="iScreen", scene="motionblur", layers=2, temporal=10, static=static,
                   extra=lambda k, t: dict(iScreenTemporal=10))]
    ref = replay(passes, n, w, h, 1.0, 1)
    for k in range(n):
        d = np.abs(got[k].astype(int) - ref[k].astype(int))
        assert (d <= 1).mean() > 0.999, ("motionblur", k, (d <= 1).mean())
    # the reference's quirk: iScreen = matrix[0][1] AFTER the roll is the oldest slot, so the first 9 frames are empty
    assert not got[0].any() and got[n - 1].any()

    scene = MultiShader(device=0); scene.initialize()
    got = run_scene(scene, 2, w, h, ssaa=1, subsample=1)
    passes = [dict(name="child", scene="multishader_child", layers=1, temporal=1, extra=lambda k, t: {}),
              dict(name="iScreen", scene="multishader", layers=1, temporal=1, extra=lambda k, t: {})]
    ref = replay(passes, 2, w, h, 1.0, 1)
    for k in range(2):
        d = np.abs(got[k].astype(int) - ref[k].astype(int))
        assert d.max() <= 1, ("multishader", k, d.max())

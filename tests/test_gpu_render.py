"""GPU parity of kernels K3/K4 and the fused frame kernel (through the C ABI) against oracle/glsl_np.py.

Tolerances. north_star: pixels within 1e-3 per float channel. Colours before the 8-bit store are compared
at 1e-3; discontinuous shaders (escape counts, hsv sectors, `if (y < bar)`) may flip a pixel on a 1-ulp
difference, so the gate is the FRACTION of channels within 1e-3, with a max-error gate
on the smooth scenes. 8-bit results may differ by 1 LSB where the pre-store value sits on a rounding tie.
"""
import numpy as np
import pytest

from oracle import glsl_np as G
from tests.helpers import native_textures, native_uniforms, visualizer_inputs

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

W, H = 256, 144
# scene → (minimum fraction of channels within 1e-3, max abs error allowed or None)
GATES = dict(default=(0.999, None), shadertoy=(1.0, 2e-5), visualizer=(0.999, None), bars=(0.999, None),
             waveform=(0.999, None), mandelbrot=(0.99, None), tetration=(0.97, None), raymarch=(0.995, None))


@pytest.fixture(scope="module")
def ctx():
    from shaderflow_b200 import _native as N
    c = N.Context(0)
    yield c
    c.destroy()


@pytest.fixture(scope="module")
def scene_inputs():
    tex, extra, time = visualizer_inputs()
    return tex, extra, time


def uniforms_for(scene, extra, time, W=W, H=H, **kw):
    u = G.Uniforms(iTime=time, iTau=0.3, iResolution=(W, H), iWantAspect=W/H, extra=dict(extra), **kw)
    return u


def gpu_screen(ctx, scene, u, tex, Wr, Hr, flags=0):
    from shaderflow_b200 import _native as N
    sid = N.scene_lookup(scene)
    info = N.scene_info(sid)
    nt = native_textures(ctx, tex)
    samplers = [nt[name] for name in info["samplers"]]
    rgba = torch.zeros((Hr, Wr, 4), dtype=torch.uint8, device="cuda")
    f32 = torch.zeros((Hr, Wr, 4), dtype=torch.float32, device="cuda")
    ctx.render_screen(sid, native_uniforms(u, info), samplers, Wr, Hr, rgba, f32, flags)
    ctx.sync()
    return rgba.cpu().numpy(), f32.cpu().numpy(), (sid, info, samplers, nt)


@pytest.mark.parametrize("scene", list(GATES))
def test_screen_pass_matches_oracle(ctx, scene, scene_inputs):
    tex, extra, time = scene_inputs
    u = uniforms_for(scene, extra, time)
    u.iSSAA = 2.0
    Wr, Hr = 2*W, 2*H
    ref = G.SCENES[scene](u, G.varyings(u, Wr, Hr), tex)
    rgba, f32, _ = gpu_screen(ctx, scene, u, tex, Wr, Hr)
    got, want = np.nan_to_num(f32[..., :3], nan=0.0, posinf=1e9, neginf=-1e9), np.nan_to_num(ref[..., :3], nan=0.0, posinf=1e9, neginf=-1e9)
    err = np.abs(np.clip(got, 0, 1) - np.clip(want, 0, 1))
    frac, cap = GATES[scene]
    assert (err <= 1e-3).mean() >= frac, (scene, (err <= 1e-3).mean(), err.max())
    if cap is not None:
        assert err.max() <= cap
    # the 8-bit store of the same colours
    d = np.abs(rgba[..., :3].astype(int) - G.to_unorm8(ref)[..., :3].astype(int))
    assert (d <= 1).mean() >= frac and (d == 0).mean() >= frac - 0.03


def test_visualizer_smooth_region_max_error(ctx, scene_inputs):
    """Away from the bar edges / waveform steps the visualizer is continuous: gate the max error there"""
    tex, extra, time = scene_inputs
    u = uniforms_for("visualizer", extra, time)
    ref = G.frag_visualizer(u, G.varyings(u, W, H), tex)
    _, f32, _ = gpu_screen(ctx, "visualizer", u, tex, W, H)
    err = np.abs(f32[..., :3] - ref[..., :3]).max(axis=-1)
    assert np.quantile(err, 0.995) < 2e-5 and (err > 1e-3).mean() < 1e-3


@pytest.mark.parametrize("scene", ["visualizer", "mandelbrot", "shadertoy", "default"])
@pytest.mark.parametrize("ssaa,subsample", [(2, 2), (2, 1), (4, 2), (1, 1), (3, 3)])
def test_fused_frame_matches_oracle_and_unfused(ctx, scene, ssaa, subsample, scene_inputs):
    from shaderflow_b200 import _native as N
    tex, extra, time = scene_inputs
    u = uniforms_for(scene, extra, time)
    ref = G.render(scene, u, tex, W, H, ssaa=float(ssaa), subsample=subsample)
    rgba, _, (sid, info, samplers, _) = gpu_screen(ctx, scene, u, tex, W*ssaa, H*ssaa)
    un = native_uniforms(u, info)
    fused = torch.zeros((H, W, 3), dtype=torch.uint8, device="cuda")
    ctx.render_frame(sid, un, samplers, W, H, ssaa, subsample, 3, fused)
    unfused = torch.zeros((H, W, 3), dtype=torch.uint8, device="cuda")
    ctx.render_final(torch.from_numpy(rgba).cuda(), W*ssaa, H*ssaa, W, H, subsample, 3, unfused)
    fused4 = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
    ctx.render_frame(sid, un, samplers, W, H, ssaa, subsample, 4, fused4)
    ctx.sync()
    fused, unfused, fused4 = fused.cpu().numpy(), unfused.cpu().numpy(), fused4.cpu().numpy()
    assert np.array_equal(fused4[..., :3], fused) and (fused4[..., 3] == 255).all()
    # fused vs unfused: same quantised sub-samples, box filter either way → at most a rounding tie apart
    d = np.abs(fused.astype(int) - unfused.astype(int))
    assert d.max() <= 1 and (d == 0).mean() > 0.7
    frac, _ = GATES[scene]
    for got in (fused, unfused):
        d = np.abs(got.astype(int) - ref["final_u8"].astype(int))
        assert (d <= 1).mean() >= frac, (scene, ssaa, subsample, (d <= 1).mean())
    # where the oracle's pre-store value is clear of a rounding tie the bytes must agree exactly
    v = ref["final_f32"]*255.0
    clear = np.abs(v - np.floor(v) - 0.5) > 0.26
    d = np.abs(fused.astype(int) - ref["final_u8"].astype(int))
    assert (d[clear] == 0).mean() >= frac - 0.005


@pytest.mark.parametrize("ssaa,subsample", [(1.0, 2), (1.5, 2), (0.5, 2), (2.0, 3), (1.0, 4)])
def test_unfused_final_pass_general_filters(ctx, ssaa, subsample, scene_inputs):
    """final.glsl when it is NOT a box filter (the reference's default ssaa=1, subsample=2 blends
    neighbours): screen + final must match the oracle's literal evaluation"""
    tex, extra, time = scene_inputs
    u = uniforms_for("visualizer", extra, time)
    ref = G.render("visualizer", u, tex, W, H, ssaa=ssaa, subsample=subsample)
    Wr, Hr = int(W*ssaa), int(H*ssaa)
    rgba, _, _ = gpu_screen(ctx, "visualizer", u, tex, Wr, Hr)
    out = torch.zeros((H, W, 3), dtype=torch.uint8, device="cuda")
    # feed the ORACLE's iScreen so this isolates the final pass
    ctx.render_final(torch.from_numpy(ref["screen_u8"]).cuda(), Wr, Hr, W, H, subsample, 3, out)
    ctx.sync()
    d = np.abs(out.cpu().numpy().astype(int) - ref["final_u8"].astype(int))
    assert d.max() <= 1 and (d == 0).mean() > 0.97
    from shaderflow_b200 import _native as N
    with pytest.raises(RuntimeError, match="box filter"):
        ctx.render_frame(2, N.Uniforms.defaults(W, H), [], W, H, 1, 2, 3, out)


@pytest.mark.parametrize("volume", [0.0, 0.4, 1.6, 3.0])
def test_fast_visualizer_path_equals_literal_transliteration(ctx, volume, scene_inputs):
    """The production kernel (table-driven taps, quad cache, PRMT widening) against the literal
    transliteration of visualizer.frag (SFB_RENDER_LITERAL): float32 re-association only"""
    from shaderflow_b200 import _native as N
    tex, extra, time = scene_inputs
    extra = dict(extra, iAudioVolume=volume)
    for (w, h) in ((W, H), (2*W, 2*H)):
        u = uniforms_for("visualizer", extra, time)
        _, fast, _ = gpu_screen(ctx, "visualizer", u, tex, w, h, N.FILTER_EXACT)
        _, literal, _ = gpu_screen(ctx, "visualizer", u, tex, w, h, N.RENDER_LITERAL)
        err = np.abs(fast[..., :3] - literal[..., :3])
        assert err.max() < 2e-5, (volume, err.max())
    # a background zoomed past the texture edge exercises the wrap branch of the quad fetch
    u = uniforms_for("visualizer", extra, time, iCameraZoom=1.6)
    _, fast, _ = gpu_screen(ctx, "visualizer", u, tex, W, H, N.FILTER_EXACT)
    _, literal, _ = gpu_screen(ctx, "visualizer", u, tex, W, H, N.RENDER_LITERAL)
    assert np.abs(fast[..., :3] - literal[..., :3]).max() < 2e-5


def test_hardware_filter_mode_error_is_reported(ctx, scene_inputs):
    """SFB_FILTER_HARDWARE (cudaTextureObject, 9-bit weights) is opt-in; its deviation from the exact
    path stays within what 1.8 fixed-point weights allow"""
    from shaderflow_b200 import _native as N
    tex, extra, time = scene_inputs
    u = uniforms_for("visualizer", extra, time)
    _, exact, _ = gpu_screen(ctx, "visualizer", u, tex, W, H, N.FILTER_EXACT)
    _, hw, _ = gpu_screen(ctx, "visualizer", u, tex, W, H, N.FILTER_HARDWARE)
    err = np.abs(exact[..., :3] - hw[..., :3])
    assert np.quantile(err, 0.99) < 4e-3 and err.mean() < 1e-3


def test_camera_projections_and_motion(ctx, scene_inputs):
    """Non-default camera uniforms: moved / zoomed / isometric perspective, stereoscopic, equirectangular"""
    tex, extra, time = scene_inputs
    variants = [
        dict(iCameraPosition=(0.3, -0.2, 0.0), iCameraZoom=0.7, iCameraIsometric=0.25),
        dict(iCameraProjection=1, iCameraSeparation=0.1),
        dict(iCameraProjection=2, iCameraZoom=0.8, iCameraForward=(0.0, 0.6, 0.8), iCameraUpward=(0.0, 0.8, -0.6)),
        dict(iCameraDolly=0.5, iCameraOrbital=0.2, iCameraFocalLength=1.5),
    ]
    for kw in variants:
        for scene in ("default", "raymarch"):
            u = uniforms_for(scene, extra, time, **kw)
            ref = G.SCENES[scene](u, G.varyings(u, W, H), tex)
            _, f32, _ = gpu_screen(ctx, scene, u, tex, W, H)
            err = np.abs(np.clip(np.nan_to_num(f32[..., :3]), 0, 1) - np.clip(np.nan_to_num(ref[..., :3]), 0, 1))
            assert (err <= 1e-3).mean() >= 0.99, (scene, kw, (err <= 1e-3).mean())


def test_texture_sampling_modes(ctx):
    """texture(): nearest/linear × repeat/clamp × u8/f32 × 1..4 components against the oracle sampler"""
    from shaderflow_b200 import _native as N
    rng = np.random.default_rng(7)
    uv = rng.uniform(-1.5, 2.5, (4096, 2)).astype(np.float32)
    uv[:64] = rng.integers(-2, 3, (64, 2)) + rng.choice([0.0, 0.5/13, 1/13, 0.5], (64, 2))   # texel edges / centres
    uv_d, out_d = torch.from_numpy(uv).cuda(), torch.zeros((4096, 4), dtype=torch.float32, device="cuda")
    for comps in (1, 2, 3, 4):
        for dtype in (np.uint8, np.float32):
            data = (rng.integers(0, 256, (9, 13, comps)).astype(np.uint8) if dtype == np.uint8
                    else rng.normal(size=(9, 13, comps)).astype(np.float32))
            nt = N.Texture(ctx, 13, 9, comps, N.DTYPE_U8 if dtype == np.uint8 else N.DTYPE_F32)
            nt.write(data)
            assert np.array_equal(nt.read()[..., :comps], data)
            for linear in (False, True):
                for rx, ry in ((True, True), (False, False), (True, False)):
                    nt.set_sampling(linear, rx, ry)
                    want = G.Texture(data, linear=linear, repeat_x=rx, repeat_y=ry).sample(uv)
                    nt.sample(uv_d, out_d, N.FILTER_EXACT); ctx.sync()
                    got = out_d.cpu().numpy()
                    if linear:
                        assert np.abs(got - want).max() < 2e-5*max(1.0, np.abs(want).max())
                    else:   # nearest: identical texels except where float rounding of u*W lands on an edge
                        assert (np.abs(got - want).max(axis=1) < 1e-6).mean() > 0.999
                    nt.sample(uv_d, out_d, N.FILTER_HARDWARE); ctx.sync()
                    hw = out_d.cpu().numpy()
                    span = float(np.ptp(data.astype(np.float32)/(255 if dtype == np.uint8 else 1)))
                    assert np.quantile(np.abs(hw - want).max(axis=1), 0.99) < span/128 + 1e-6
            nt.destroy()
    # sub-rectangle writes and the size limit (texture.py:251-252)
    nt = N.Texture(ctx, 8, 4, 2, N.DTYPE_F32)
    col = np.arange(8, dtype=np.float32).reshape(4, 1, 2)
    nt.write(col, viewport=(5, 0, 1, 4))
    back = nt.read()
    assert np.array_equal(back[:, 5], col[:, 0]) and not back[:, :5].any()
    with pytest.raises(RuntimeError, match="too large"):
        N.Texture(ctx, 1 << 20, 4, 4, N.DTYPE_U8)


def test_large_frame_properties(ctx, scene_inputs):
    """BASELINE size (3840x2160, 2xSSAA) through the fused kernel: the picture is symmetric where the
    shader is (shadertoy is constant along y), deterministic, and rgb24 rows are tightly packed"""
    from shaderflow_b200 import _native as N
    W4, H4 = 3840, 2160
    u = N.Uniforms.defaults(W4, H4); u.iTime = 1.25; u.iSSAA = 2.0
    a = torch.zeros((H4, W4, 3), dtype=torch.uint8, device="cuda")
    b = torch.zeros((H4, W4, 3), dtype=torch.uint8, device="cuda")
    sid = N.scene_lookup("shadertoy")
    ctx.render_frame(sid, u, [], W4, H4, 2, 2, 3, a)
    ctx.render_frame(sid, u, [], W4, H4, 2, 2, 3, b)
    ctx.sync()
    a, b = a.cpu().numpy(), b.cpu().numpy()
    assert np.array_equal(a, b)
    # green channel depends on y only; red/blue on x only
    assert (a[..., 1] == a[:, :1, 1]).all() and (a[..., 0] == a[:1, :, 0]).all()
    uo = G.Uniforms(iTime=1.25, iResolution=(W4, H4), iWantAspect=W4/H4)
    # oracle on the two bottom sub-sample rows only (the x profile): shade, 8-bit store, 2x2 box, store
    sub = G.to_unorm8(G.frag_shadertoy(uo, G.varyings(uo, 2*W4, 2*H4, rows=slice(0, 2)), {}))[..., :3].astype(np.float32)
    box = G.to_unorm8(sub.reshape(1, 2, W4, 2, 3).mean(axis=(1, 3))/255.0)
    assert np.abs(a[0].astype(int) - box[0].astype(int)).max() <= 1
    assert (a[0] == box[0]).mean() > 0.9


@pytest.mark.parametrize("bg_size,out,ssaa,volume,camera", [
    ((960, 540), (1920, 1080), 2, 1.6, {}),                      # J=8, interior windows by TMA bulk rows + wrapped edges
    ((960, 540), (1920, 1080), 2, 0.0, {}),                      # no blur: every tap on the centre position
    ((512, 288), (512, 288), 2, 3.0, {}),                        # coarser step: J=4, blur radius saturated
    ((512, 288), (256, 144), 4, 0.9, {}),                        # 4x4 sub-samples, J=4
    ((640, 360), (1280, 720), 1, 1.2, {}),                       # ssaa 1 (pixels = fragments)
    ((240, 135), (258, 146), 2, 1.0, {}),                        # partial tiles right/top, width not a multiple of 4
    ((960, 540), (960, 540), 2, 1.3, dict(iCameraZoom=1.7, iCameraPosition=(0.4, 0.1, 0.0))),   # moved camera: window crosses the wrap
    ((130, 74), (1920, 1080), 2, 2.0, {}),                       # background narrower than the window row: all wrapped loads
    ((640, 360), (640, 360), 1, 1.2, {}),                        # the reference's default export geometry: ~0.8 texel per
                                                                 # fragment → 3 rows per thread, 96-texel windows
    ((640, 360), (320, 180), 2, 2.0, {}),                        # the same step with 2x2 sub-samples: J=2, 96-texel windows
    ((640, 360), (640, 360), 2, 0.7, {}),                        # 0.4 texel per fragment: J=4
    ((800, 360), (640, 360), 1, 1.0, {}),                        # background of another aspect ratio
    ((1920, 1080), (1920, 1080), 1, 1.6, {}),                    # BASELINE configs[1] itself
])
def test_separable_visualizer_kernel_equals_tiled_and_oracle(ctx, bg_size, out, ssaa, volume, camera):
    """visualizer_rows.cu (one thread per fragment column, hinge-weight table) against the per-pixel tiled
    kernel (SFB_RENDER_TILED) on identical inputs — same mathematics, float32 re-association and ulp-level
    differences of the back-end math only, so the rgb24 bytes agree except on rounding ties — and, at a size the oracle shades in seconds, against the oracle."""
    from shaderflow_b200 import _native as N
    Wo, Ho = out
    tex, extra, time = visualizer_inputs(bg_size=bg_size)
    extra = dict(extra, iAudioVolume=volume)
    u = uniforms_for("visualizer", extra, time, W=Wo, H=Ho, **camera)
    sid = N.scene_lookup("visualizer"); info = N.scene_info(sid)
    nt = native_textures(ctx, tex)
    samplers = [nt[name] for name in info["samplers"]]
    un = native_uniforms(u, info)
    rows = torch.zeros((Ho, Wo, 3), dtype=torch.uint8, device="cuda")
    tiled = torch.zeros((Ho, Wo, 3), dtype=torch.uint8, device="cuda")
    rows4 = torch.zeros((Ho, Wo, 4), dtype=torch.uint8, device="cuda")
    ctx.render_frame(sid, un, samplers, Wo, Ho, ssaa, ssaa, 3, rows)
    ctx.render_frame(sid, un, samplers, Wo, Ho, ssaa, ssaa, 3, tiled, N.RENDER_TILED)
    ctx.render_frame(sid, un, samplers, Wo, Ho, ssaa, ssaa, 4, rows4)
    ctx.sync()
    rows, tiled, rows4 = rows.cpu().numpy(), tiled.cpu().numpy(), rows4.cpu().numpy()
    assert np.array_equal(rows4[..., :3], rows) and (rows4[..., 3] == 255).all()
    d = np.abs(rows.astype(int) - tiled.astype(int))
    # the two back ends round differently by ulps (pow through exp2/log2, reciprocal multiplies): a pixel
    # sitting exactly on a bar edge / waveform step may take the other branch — those are counted, not banned
    assert (d <= 1).mean() > 0.9999 and (d == 0).mean() > 0.995, (d.max(), (d > 1).sum(), (d == 0).mean())
    if Wo*Ho*ssaa*ssaa <= 600_000:
        ref = G.render("visualizer", u, tex, Wo, Ho, ssaa=float(ssaa), subsample=ssaa)
        d = np.abs(rows.astype(int) - ref["final_u8"].astype(int))
        assert (d <= 1).mean() >= 0.999, (d <= 1).mean()


@pytest.mark.parametrize("want_aspect", [None, 1.25])
def test_visualizer_screen_pass_into_rgba8_target(ctx, scene_inputs, want_aspect):
    """The iScreen pass of an unfused export (sfb_render_target into an RGBA8 texture) runs the separable kernel with
    one fragment per "pixel" (the tiled one when the geometry does not qualify): same bytes as the generic
    per-fragment pass, alpha included (0 where the fragment is out of
    bounds: a wanted aspect narrower than the target's leaves bars on both sides, camera.glsl:86)"""
    from shaderflow_b200 import _native as N
    tex, extra, time = scene_inputs
    u = uniforms_for("visualizer", extra, time)
    camera = want_aspect is not None
    if camera:
        u.iWantAspect = want_aspect
    rgba, _, (sid, info, samplers, _) = gpu_screen(ctx, "visualizer", u, tex, W, H)
    target = N.Texture(ctx, W, H, 4, N.DTYPE_U8)
    ctx.render_target(sid, native_uniforms(u, info), samplers, target)
    ctx.sync()
    got = target.read()
    d = np.abs(got.astype(int) - rgba.astype(int))
    assert (d <= 1).mean() > 0.9999 and (d == 0).mean() > 0.995, ((d <= 1).mean(), (d == 0).mean())
    assert np.array_equal(got[..., 3], rgba[..., 3])
    if camera:
        assert (got[..., 3] == 0).any() and (got[..., 3] == 255).any()


@pytest.mark.parametrize("scene", ["mandelbrot", "tetration", "raymarch"])
@pytest.mark.parametrize("ssaa,comps", [(2, 3), (4, 3), (4, 4)])
def test_lane_per_subsample_kernel_equals_the_generic_one(ctx, scene, ssaa, comps):
    """frame_lanes_kernel (one lane per sub-sample, shuffle box sum; Mandelbrot with the closed-form interior
    tests) against frame_kernel (one thread per pixel, the literal loop: SFB_RENDER_LITERAL): identical bytes,
    including partial tiles and a width that is not a multiple of 4"""
    from shaderflow_b200 import _native as N
    for Wo, Ho in ((250, 141), (256, 144)):
        u = N.Uniforms.defaults(Wo, Ho); u.iSSAA = float(ssaa)
        sid = N.scene_lookup(scene)
        a = torch.zeros((Ho, Wo, comps), dtype=torch.uint8, device="cuda")
        b = torch.zeros((Ho, Wo, comps), dtype=torch.uint8, device="cuda")
        ctx.render_frame(sid, u, [], Wo, Ho, ssaa, ssaa, comps, a)
        ctx.render_frame(sid, u, [], Wo, Ho, ssaa, ssaa, comps, b, N.RENDER_LITERAL)
        ctx.sync()
        if scene == "tetration":
            # the default launch evaluates pow(C.r, Z.x)·exp(−Z.y·C.t) as one exp of the combined exponent
            # (scenes.cuh): ulp-level differences that the chaotic map amplifies for isolated sub-samples (99.7 % of
            # sub-samples agree; a pixel averages ssaa² of them) — the gate is this scene's parity gate, 97 %
            d = (a.int() - b.int()).abs()
            assert (d <= 1).float().mean().item() >= 0.97, (ssaa, (d <= 1).float().mean().item())
        else:
            assert torch.equal(a, b), (scene, ssaa, (a != b).sum().item())


def test_mandelbrot_interior_tests_are_exact_at_8k(ctx):
    """BASELINE configs[3] geometry (7680x4320, ssaa 4 = 530.8 M fragments): skipping the iteration for points of
    the main cardioid / period-2 disc must not change a single byte"""
    from shaderflow_b200 import _native as N
    Wo, Ho = 7680, 4320
    u = N.Uniforms.defaults(Wo, Ho); u.iSSAA = 4.0
    sid = N.scene_lookup("mandelbrot")
    a = torch.zeros((Ho, Wo, 3), dtype=torch.uint8, device="cuda")
    b = torch.zeros((Ho, Wo, 3), dtype=torch.uint8, device="cuda")
    ctx.render_frame(sid, u, [], Wo, Ho, 4, 4, 3, a)
    ctx.render_frame(sid, u, [], Wo, Ho, 4, 4, 3, b, N.RENDER_LITERAL)
    ctx.sync()
    assert torch.equal(a, b), (a != b).sum().item()


def test_float16_textures_sample_and_serve_as_render_targets(ctx):
    """SFB_DTYPE_F16 (texture.py:28-38 "f2"): upload / read-back identity, exact sampling of half texels, and a pass
    rendered into a half RGBA target = the float32 pass rounded to half"""
    from shaderflow_b200 import _native as N
    rng = np.random.default_rng(4)
    uv = rng.uniform(-0.3, 1.3, (2000, 2)).astype(np.float32)
    uv_d = torch.from_numpy(uv).cuda()
    out_d = torch.zeros((2000, 4), dtype=torch.float32, device="cuda")
    for comps in (1, 2, 4):
        data = rng.normal(size=(9, 13, comps)).astype(np.float16)
        nt = N.Texture(ctx, 13, 9, comps, N.DTYPE_F16)
        nt.write(data)
        assert np.array_equal(nt.read()[..., :comps], data)
        for linear in (False, True):
            nt.set_sampling(linear, True, False)
            want = G.Texture(data.astype(np.float32), linear=linear, repeat_x=True, repeat_y=False).sample(uv)
            nt.sample(uv_d, out_d, N.FILTER_EXACT); ctx.sync()
            got = out_d.cpu().numpy()
            if linear:
                assert np.abs(got - want).max() < 2e-5*max(1.0, np.abs(want).max())
            else:
                assert (np.abs(got - want).max(axis=1) < 1e-6).mean() > 0.999
        nt.destroy()
    Wt, Ht = 96, 54
    u = N.Uniforms.defaults(Wt, Ht); u.iTime = 0.8
    sid = N.scene_lookup("shadertoy")
    half = N.Texture(ctx, Wt, Ht, 4, N.DTYPE_F16)
    full = N.Texture(ctx, Wt, Ht, 4, N.DTYPE_F32)
    ctx.render_target(sid, u, [], half)
    ctx.render_target(sid, u, [], full)
    ctx.sync()
    assert np.array_equal(half.read(), full.read().astype(np.float16))
    half.destroy(); full.destroy()

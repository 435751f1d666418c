"""ShaderVideo on the GPU: sfb_video_frame (flip + colour conversion into the texture's storage) against numpy, and the
module's update rule (shaderflow/video.py:57-66: one new frame whenever scene.time > frames_read/fps) through a scene
whose fragment is compiled at run time."""
import numpy as np
import pytest

from shaderflow_b200 import synthetic

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def ctx():
    from shaderflow_b200 import _native as N
    c = N.Context(0)
    yield c
    c.destroy()


def yuv_to_rgb(y, u, v, full):
    y, u, v = y.astype(np.float32), u.astype(np.float32) - np.float32(128), v.astype(np.float32) - np.float32(128)
    f = np.float32
    if full:
        r, g, b = y + f(1.402)*v, y - f(0.344136)*u - f(0.714136)*v, y + f(1.772)*u
    else:
        l = f(1.164383)*(y - f(16))
        r, g, b = l + f(1.596027)*v, l - f(0.391762)*u - f(0.812968)*v, l + f(2.017232)*u
    return np.clip(np.rint(np.stack([r, g, b], -1)), 0, 255).astype(np.uint8)


@pytest.mark.parametrize("fmt_name,sub", [("YUV420P", (1, 1)), ("YUV422P", (1, 0)), ("YUV444P", (0, 0))])
@pytest.mark.parametrize("full", [False, True])
def test_planar_yuv_frames(ctx, fmt_name, sub, full):
    from shaderflow_b200 import _native as N
    W, H = 50, 34
    rng = np.random.default_rng(11)
    cw, ch = (W + sub[0]) >> sub[0], (H + sub[1]) >> sub[1]
    y, u, v = rng.integers(0, 256, (H, W), np.uint8), rng.integers(0, 256, (ch, cw), np.uint8), rng.integers(0, 256, (ch, cw), np.uint8)
    fmt = getattr(N, "VIDEO_" + fmt_name) | (N.VIDEO_FULL_RANGE if full else 0)
    raw = np.concatenate([y.ravel(), u.ravel(), v.ravel()])
    assert raw.size == N.video_frame_bytes(fmt, W, H)
    tex = N.Texture(ctx, W, H, 3, N.DTYPE_U8)
    ctx.video_frame(torch.from_numpy(raw).cuda(), fmt, W, H, True, tex)
    ctx.sync()
    up = lambda p: np.repeat(np.repeat(p, 1 << sub[1], axis=0), 1 << sub[0], axis=1)[:H, :W]
    want = np.flipud(yuv_to_rgb(y, up(u), up(v), full))                 # texture row 0 = the image's bottom row
    got = tex.read()[..., :3]
    d = np.abs(got.astype(int) - want.astype(int))
    assert d.max() <= 1 and (d == 0).mean() > 0.999, (d.max(), (d == 0).mean())     # fma contraction at rounding ties


def test_rgb_frames_both_orientations(ctx):
    from shaderflow_b200 import _native as N
    W, H = 33, 21
    rng = np.random.default_rng(12)
    rgb, rgba = rng.integers(0, 256, (H, W, 3), np.uint8), rng.integers(0, 256, (H, W, 4), np.uint8)
    tex = N.Texture(ctx, W, H, 3, N.DTYPE_U8)
    ctx.video_frame(torch.from_numpy(rgb).cuda(), N.VIDEO_RGB24, W, H, True, tex); ctx.sync()
    assert np.array_equal(tex.read()[..., :3], np.flipud(rgb))
    ctx.video_frame(torch.from_numpy(rgb).cuda(), N.VIDEO_RGB24, W, H, False, tex); ctx.sync()
    assert np.array_equal(tex.read()[..., :3], rgb)
    tex4 = N.Texture(ctx, W, H, 4, N.DTYPE_U8)
    ctx.video_frame(torch.from_numpy(rgba).cuda(), N.VIDEO_RGBA32, W, H, True, tex4); ctx.sync()
    assert np.array_equal(tex4.read(), np.flipud(rgba))
    with pytest.raises(RuntimeError, match="frame 8x8"):
        ctx.video_frame(torch.zeros(192, dtype=torch.uint8, device="cuda"), N.VIDEO_RGB24, 8, 8, True, tex)


@pytest.mark.parametrize("video_fps,scene_fps", [(30, 60), (24, 60), (60, 30)])
def test_video_scene_follows_the_reference_update_rule(tmp_path, video_fps, scene_fps):
    """Raw rgb24 clip → ShaderVideo → a fragment showing the texture 1:1 → the exported frames ARE the clip's frames, the
    one video.py:57-66 would have uploaded by that time (at most one new frame per scene frame)"""
    from examples.demo import ShaderScene
    from shaderflow.video import ShaderVideo
    W, H, n = 64, 36, 12
    clip = synthetic.video_frames(W, H, n)
    path = tmp_path/"clip.rgb"
    path.write_bytes(clip.tobytes())

    times = {}

    class Player(ShaderScene):
        def update(self):
            times[self.frame_index] = self.time            # what ShaderVideo.update compares (accumulated float time)

        def build(self):
            self.video = ShaderVideo(scene=self, path=path, width=W, height=H, fps=video_fps)
            self.video.texture.filter = "nearest"
            self.shader.fragment = "void main() { fragColor = vec4(astexture(iVideo, astuv).rgb, 1.0); }"
    scene = Player()
    shown = {}
    def grab(index, pointer):
        scene.cuda.sync(); shown[index] = scene.frame_tensor.cpu().numpy().copy()
    total = 20
    scene.main(width=W, height=H, ssaa=1, subsample=1, fps=scene_fps, time=total/scene_fps, on_frame=grab)
    assert scene.shader.scene_id >= 1000
    read = 0
    for k in range(total):
        t = times[k]
        assert abs(t - k/scene_fps) < 1e-9
        if t > read/video_fps:
            read += 1
        if read == 0:
            continue                                   # nothing uploaded yet at t = 0 (strict >)
        current = min(read, n) - 1
        assert np.array_equal(shown[k], np.flipud(clip[current])), (k, current)     # frames are bottom row first
    assert read > 3


def test_video_example_scene_with_a_y4m_clip(tmp_path):
    from examples import demo
    clip = synthetic.video_frames(96, 54, 8)
    path = tmp_path/"clip.y4m"
    synthetic.write_y4m(path, clip, fps=30, colorspace="444")
    demo.Video.path = path
    try:
        scene = demo.Video()
        shown = {}
        def grab(index, pointer):
            scene.cuda.sync(); shown[index] = scene.frame_tensor.cpu().numpy().copy()
        scene.main(width=96, height=54, ssaa=1, subsample=1, fps=30.0, time=6/30, on_frame=grab)
    finally:
        demo.Video.path = None
    # 4:4:4 in BT.601 limited range loses at most a few codes per channel; frame k shows clip frame k-1
    d = np.abs(shown[5].astype(int) - np.flipud(clip[4]).astype(int))
    assert d.max() <= 4 and d.mean() < 1.0, (d.max(), d.mean())

"""ShaderVideo's file side without a GPU: Y4M / raw layouts and frame access (shaderflow_b200/video.py)"""
import numpy as np
import pytest

from shaderflow_b200 import _native as N
from shaderflow_b200 import synthetic, video


def test_y4m_header_layout_and_frames(tmp_path):
    clip = synthetic.video_frames(32, 18, 5)
    for colorspace, fmt, size in (("420jpeg", N.VIDEO_YUV420P, 32*18 + 2*16*9), ("422", N.VIDEO_YUV422P, 32*18 + 2*16*18),
                                  ("444", N.VIDEO_YUV444P, 3*32*18)):
        path = tmp_path/f"clip_{colorspace}.y4m"
        synthetic.write_y4m(path, clip, fps=24, colorspace=colorspace)
        info = video.parse_y4m(path)
        assert (info.width, info.height, info.fps, info.format, info.frame_bytes, info.frames) == (32, 18, 24.0, fmt, size, 5)
        assert info.top_down and info.stride == size + len(b"FRAME\n")
        frames = video.FileFrames(path, info)
        raw = path.read_bytes()
        for k in (0, 4):
            at = raw.index(b"\n") + 1 + k*info.stride + 6
            assert bytes(frames.frame(k)) == raw[at:at + size]
        # luma of a white / black pixel in limited range
        assert frames.frame(0)[:32*18].min() >= 16 and frames.frame(0)[:32*18].max() <= 235
    assert N.video_frame_bytes(N.VIDEO_YUV420P | N.VIDEO_FULL_RANGE, 31, 17) == 31*17 + 2*16*9 and N.video_frame_bytes(99, 4, 4) == 0


def test_y4m_rejects_what_it_cannot_read(tmp_path):
    (tmp_path/"not.y4m").write_bytes(b"RIFF....")
    with pytest.raises(ValueError, match="not a YUV4MPEG2"):
        video.parse_y4m(tmp_path/"not.y4m")
    (tmp_path/"deep.y4m").write_bytes(b"YUV4MPEG2 W4 H4 F30:1 C420p10\nFRAME\n" + bytes(48))
    with pytest.raises(ValueError, match="not supported"):
        video.parse_y4m(tmp_path/"deep.y4m")
    (tmp_path/"full.y4m").write_bytes(b"YUV4MPEG2 W4 H2 F30000:1001 C444 XCOLORRANGE=FULL\nFRAME\n" + bytes(24))
    info = video.parse_y4m(tmp_path/"full.y4m")
    assert info.format == N.VIDEO_YUV444P | N.VIDEO_FULL_RANGE and abs(info.fps - 29.97) < 0.01 and info.frames == 1


def test_headerless_frames(tmp_path):
    clip = synthetic.video_frames(16, 8, 3)
    path = tmp_path/"clip.rgb"
    path.write_bytes(clip.tobytes() + b"tail")                       # a partial frame at the end is not a frame
    info = video.raw_info(path, 16, 8, 30.0, N.VIDEO_RGB24, bottom_up=True)
    assert info.frames == 3 and not info.top_down and info.frame_bytes == 16*8*3
    assert np.array_equal(np.asarray(video.FileFrames(path, info).frame(2)).reshape(8, 16, 3), clip[2])


def test_dry_scene_checks_its_glsl_and_steps_the_video_rule(tmp_path):
    """Without a device (`backend="dry"`) a scene still compiles its own GLSL — translator + NVRTC, so a typo is found
    before any GPU is involved — and ShaderVideo still applies video.py:57-66's rule: one more frame is consumed whenever
    scene.time > frames_read/fps"""
    import ctypes.util
    from pathlib import Path
    if not (ctypes.util.find_library("nvrtc") or Path("/usr/local/cuda/lib64/libnvrtc.so.12").exists()):
        pytest.skip("libnvrtc is not installed")
    import examples.demo as demo
    from shaderflow.video import ShaderVideo
    clip = synthetic.video_frames(32, 18, 9)
    synthetic.write_y4m(tmp_path/"clip.y4m", clip, fps=24, colorspace="420jpeg")
    consumed = []

    class Player(demo.ShaderScene):
        def build(self):
            self.video = ShaderVideo(scene=self, path=tmp_path/"clip.y4m")
            self.shader.fragment = "void main() { fragColor = vec4(astexture(iVideo, astuv).rgb*iTau, 1.0); }"
        def update(self):
            consumed.append((self.time, self.video._frames))
    scene = Player(backend="dry")
    scene.main(width=32, height=18, fps=60.0, time=0.5)
    assert scene.shader.scene_id >= 1000 and scene.shader.scene_info["samplers"] == ["iVideo0x0"]
    assert (scene.video.width, scene.video.height, scene.video.fps) == (32, 18, 24.0) and len(consumed) == 30
    read = 0
    for time, before in consumed:                  # the scene's update runs before the video's in the same frame
        assert before == read
        if time > read/24.0:
            read += 1
    assert 9 < read <= 13                          # the clip has 9 frames: past its end the counter runs on, the last frame stays

    class Typo(demo.ShaderScene):
        def build(self):
            self.shader.fragment = "void main() { fragColor = vec4(astuv, iTimee, 1.0); }"
    with pytest.raises(RuntimeError, match="iTimee"):
        Typo(backend="dry").main(width=32, height=18, time=0.1)


def write_with_opencv(path, clip: np.ndarray, fps: float, fourcc: str) -> None:
    cv2 = pytest.importorskip("cv2")
    writer = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*fourcc), fps, (clip.shape[2], clip.shape[1]))
    if not writer.isOpened():
        pytest.skip(f"this OpenCV build cannot encode {fourcc}")
    for frame in clip:
        writer.write(np.ascontiguousarray(frame[..., ::-1]))           # OpenCV's frames are BGR
    writer.release()


def test_compressed_files_decode_through_opencv_when_there_is_no_ffmpeg(tmp_path, monkeypatch):
    """No ffmpeg binary (this image): a compressed file goes through the libavcodec inside OpenCV. A lossless codec
    returns the encoder's input exactly, frame by frame; a lossy one (MPEG-4) stays close; the end of the stream keeps
    the last frame; a dry scene steps the same update rule over it and a new export starts the decoder over"""
    import shutil
    monkeypatch.setattr(shutil, "which", lambda name, *a, **k: None)
    clip = synthetic.video_frames(64, 36, 7)
    write_with_opencv(tmp_path/"clip.mkv", clip, 24.0, "FFV1")
    assert video.CodecFrames.probe(tmp_path/"clip.mkv") == (64, 36, 24.0)
    info = video.VideoInfo(64, 36, 24.0, N.VIDEO_RGB24, 64*36*3)
    frames = video.CodecFrames(tmp_path/"clip.mkv", info)
    for k in (0, 1, 4, 6):                                             # sequential access, skipping allowed
        got = frames.frame(k)
        assert got.dtype == np.uint8 and got.shape == (64*36*3,) and np.array_equal(got.reshape(36, 64, 3), clip[k])
    assert np.array_equal(frames.frame(9).reshape(36, 64, 3), clip[6]) and frames.position == 7
    write_with_opencv(tmp_path/"clip.mp4", clip, 30.0, "mp4v")
    lossy = video.CodecFrames(tmp_path/"clip.mp4", info)
    error = np.abs(lossy.frame(3).reshape(36, 64, 3).astype(int) - clip[3].astype(int))
    assert error.mean() < 12 and video.CodecFrames.probe(tmp_path/"clip.mp4")[2] == 30.0
    with pytest.raises(RuntimeError, match="cannot"):
        video.CodecFrames(tmp_path/"missing.mp4", info)

    import examples.demo as demo
    from shaderflow.video import ShaderVideo

    class Player(demo.ShaderScene):
        def build(self):
            self.video = ShaderVideo(scene=self, path=tmp_path/"clip.mkv")
            self.shader.fragment = demo.shaders/"video.frag"
    scene = Player(backend="dry"); scene.initialize()
    assert isinstance(scene.video._reader, video.CodecFrames) and scene.video.info.format == N.VIDEO_RGB24
    assert (scene.video.width, scene.video.height, scene.video.fps) == (64, 36, 24.0)
    scene.main(width=64, height=36, fps=60.0, time=0.25)
    assert scene.video._frames == 6                                   # time 14/60 > 5/24
    scene.video._reader.frame(2)
    scene.main(width=64, height=36, fps=60.0, time=0.1)
    assert scene.video._reader.position == 0 and scene.video._frames == 3


def test_in_process_encoder_stands_in_for_the_ffmpeg_child(tmp_path):
    """exporting.InProcessEncoder: rawvideo rgb24 frames, bottom row first, written to its pipe — what the sink's writer
    thread sends an ffmpeg child — come out as a video file whose frames are the export's, top row first. Lossless
    container → byte for byte; MPEG-4 / Motion-JPEG → close. A partial last frame is dropped"""
    import os
    cv2 = pytest.importorskip("cv2")
    from shaderflow_b200.exporting import InProcessEncoder
    clip = synthetic.video_frames(64, 36, 6)                               # top row first
    for name, exact in (("out.mkv", True), ("out.mp4", False), ("out.avi", False)):
        encoder = InProcessEncoder(tmp_path/name, 64, 36, 30.0)
        for frame in clip:
            wire = np.ascontiguousarray(frame[::-1]).tobytes()             # the export's row order
            at = 0
            while at < len(wire):                                          # in pieces, like a pipe delivers them
                at += os.write(encoder.write_fd, wire[at:at + 5000])
        os.write(encoder.write_fd, b"\x00"*100)
        assert encoder.close() == 6
        capture = cv2.VideoCapture(str(tmp_path/name))
        assert capture.get(cv2.CAP_PROP_FPS) == 30.0 and int(capture.get(cv2.CAP_PROP_FRAME_WIDTH)) == 64
        for k, frame in enumerate(clip):
            ok, bgr = capture.read()
            assert ok, (name, k)
            error = np.abs(bgr[..., ::-1].astype(int) - frame.astype(int))
            assert (error.max() == 0) if exact else (error.mean() < 12), (name, k, error.max(), error.mean())
        assert not capture.read()[0]
    with pytest.raises(RuntimeError, match="cannot encode"):
        InProcessEncoder(tmp_path/"no_such_dir"/"out.mp4", 64, 36, 30.0)


def test_yuv_rule_against_ffmpegs_swscale(tmp_path):
    """What a Y4M clip looks like through the reference is libswscale's yuv420p → rgb24 (its ffmpeg child, default flags:
    integer tables, chroma replicated 2 x 2, no accurate rounding). sfb_video_frame computes BT.601 in float32 instead
    (tests/test_gpu_video.py holds the kernel to this formula within 1 LSB). The distance between the two, measured on
    the real library: at most 3 codes, 1.1 on average — swscale's fast path carries a bias of about one code; the float
    rule is the one closer to the RGB the clip was made from."""
    import ctypes, glob, importlib.util
    from pathlib import Path
    cv2 = pytest.importorskip("cv2")
    from tests.test_gpu_video import yuv_to_rgb
    site = Path(list(importlib.util.find_spec("cv2").submodule_search_locations)[0]).parent
    hits = [h for d in ("opencv_python_headless.libs", "opencv_python.libs") for h in glob.glob(str(site/d/"libswscale-*.so*"))]
    if not hits:
        pytest.skip("no libswscale in the OpenCV wheel")
    sws, P = ctypes.CDLL(hits[0]), ctypes.c_void_p
    sws.sws_getContext.restype, sws.sws_getContext.argtypes = P, [ctypes.c_int]*7 + [P, P, P]
    sws.sws_scale.argtypes = [P, P, P, ctypes.c_int, ctypes.c_int, P, P]
    sws.sws_freeContext.argtypes = [P]
    W, H = 320, 180
    clip = synthetic.video_frames(W, H, 2)
    synthetic.write_y4m(tmp_path/"clip.y4m", clip, fps=30, colorspace="420jpeg")
    info = video.parse_y4m(tmp_path/"clip.y4m")
    planes = np.asarray(video.FileFrames(tmp_path/"clip.y4m", info).frame(1))
    y = planes[:W*H].reshape(H, W).copy()
    u = planes[W*H:W*H + W*H//4].reshape(H//2, W//2).copy()
    v = planes[W*H + W*H//4:].reshape(H//2, W//2).copy()
    ours = yuv_to_rgb(y, np.repeat(np.repeat(u, 2, 0), 2, 1), np.repeat(np.repeat(v, 2, 0), 2, 1), False)
    context = sws.sws_getContext(W, H, 0, W, H, 2, 4, None, None, None)        # AV_PIX_FMT_YUV420P → RGB24, SWS_BICUBIC
    theirs = np.zeros((H, W, 3), np.uint8)
    src, src_stride = (P*4)(y.ctypes.data, u.ctypes.data, v.ctypes.data, None), (ctypes.c_int*4)(W, W//2, W//2, 0)
    dst, dst_stride = (P*4)(theirs.ctypes.data, None, None, None), (ctypes.c_int*4)(W*3, 0, 0, 0)
    assert sws.sws_scale(context, src, src_stride, 0, H, dst, dst_stride) == H
    sws.sws_freeContext(context)
    distance = np.abs(ours.astype(int) - theirs.astype(int))
    assert distance.max() <= 3 and distance.mean() < 1.3
    source = clip[1].astype(int)
    assert np.abs(ours - source).mean() <= np.abs(theirs - source).mean()

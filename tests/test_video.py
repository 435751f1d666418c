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

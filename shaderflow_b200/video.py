"""ShaderVideo — a video file as a texture (API mirror of shaderflow/video.py:12-66).

The reference asks ffmpeg for rgb24 frames, flips each with numpy, copies it to C order and uploads it when
`scene.time > frames_read/fps` (video.py:57-66). Here the frame source hands over the FILE's own bytes and the GPU does
the rest (`sfb_video_frame`: vertical flip, planar YUV → RGB, padding, straight into the texture's storage):

    .y4m                      YUV4MPEG2, 8-bit C420* / C422 / C444 — parsed here, frames memory-mapped
    .rgb / .raw / .rgba       headerless rgb24 / rgba frames (what this backend's own raw exports are; give width,
                              height and fps; `bottom_up=True` for an export of this backend, which is bottom row first)
    anything else             through an `ffmpeg` child (`-f rawvideo -pix_fmt rgb24`), like the reference, when the
                              binary exists; otherwise through the libavcodec inside OpenCV (`cv2.VideoCapture`, when
                              the package is installed): compressed files (mp4, mkv, avi, webm …) decode on the host
                              to rgb24 and take the same upload path

Frames travel file → pinned staging (two buffers) → H2D → kernel; a frame another rank shades (sharded exports) is
skipped without being read. When the file ends the last frame stays (the reference's iterator would raise)."""
from __future__ import annotations

import shutil
import subprocess
from pathlib import Path
from typing import Any, Optional

import numpy as np
from attrs import define, field

from shaderflow_b200 import _native as N
from shaderflow_b200.module import ShaderModule
from shaderflow_b200.texture import ShaderTexture

_Y4M_FORMATS = {"420": N.VIDEO_YUV420P, "422": N.VIDEO_YUV422P, "444": N.VIDEO_YUV444P}


@define
class VideoInfo:
    width: int
    height: int
    fps: float
    format: int                  # N.VIDEO_*
    frame_bytes: int
    first: int = 0               # byte offset of the first frame's pixels
    stride: int = 0              # bytes from one frame's pixels to the next
    frames: int = 0
    top_down: bool = True


def parse_y4m(path) -> VideoInfo:
    """YUV4MPEG2 stream header → layout. Every frame is 'FRAME' + parameters + newline + planar Y, U, V"""
    path = Path(path)
    with open(path, "rb") as f:
        header = f.readline(4096)
        if not header.startswith(b"YUV4MPEG2 ") or not header.endswith(b"\n"):
            raise ValueError(f"'{path}' is not a YUV4MPEG2 file")
        width = height = 0
        fps, colorspace, full = 25.0, "420", False
        for token in header.split()[1:]:
            key, value = chr(token[0]), token[1:].decode("ascii", "replace")
            if key == "W":
                width = int(value)
            elif key == "H":
                height = int(value)
            elif key == "F":
                num, den = value.split(":")
                fps = int(num)/int(den) if int(den) else 25.0
            elif key == "C":
                colorspace = value
            elif key == "X" and value.upper() == "COLORRANGE=FULL":
                full = True
        base = next((k for k in _Y4M_FORMATS if colorspace.startswith(k)), None)
        if base is None or any(bits in colorspace for bits in ("p10", "p12", "p14", "p16")) or colorspace.startswith("mono"):
            raise ValueError(f"'{path}': Y4M colourspace C{colorspace} is not supported (8-bit 420 / 422 / 444 are)")
        if width <= 0 or height <= 0:
            raise ValueError(f"'{path}': missing W / H in the Y4M header")
        fmt = _Y4M_FORMATS[base] | (N.VIDEO_FULL_RANGE if full else 0)
        frame_bytes = N.video_frame_bytes(fmt, width, height)
        start = f.tell()
        marker = f.readline(256)
        if not marker.startswith(b"FRAME"):
            return VideoInfo(width, height, fps, fmt, frame_bytes, start, 0, 0)
        size = path.stat().st_size
        stride = len(marker) + frame_bytes                       # frame headers without parameters are all alike
        return VideoInfo(width, height, fps, fmt, frame_bytes, start + len(marker), stride, (size - start)//stride)


def raw_info(path, width: int, height: int, fps: float, fmt: int, bottom_up: bool) -> VideoInfo:
    frame_bytes = N.video_frame_bytes(fmt, width, height)
    return VideoInfo(width, height, fps, fmt, frame_bytes, 0, frame_bytes, Path(path).stat().st_size//frame_bytes, top_down=not bottom_up)


class FileFrames:
    """Random access to the frames of an uncompressed file (memory-mapped)"""
    def __init__(self, path, info: VideoInfo):
        self.info = info
        self.data = np.memmap(path, dtype=np.uint8, mode="r") if info.frames else np.zeros(0, np.uint8)

    def __len__(self) -> int:
        return self.info.frames

    def frame(self, index: int) -> np.ndarray:
        at = self.info.first + index*self.info.stride
        return self.data[at:at + self.info.frame_bytes]


class PipeFrames:
    """Sequential rgb24 frames from an ffmpeg child, as shaderflow/ffmpeg.py:1116-1147 reads them"""
    def __init__(self, path, info: VideoInfo):
        self.info, self.position, self.last = info, 0, None
        self.child = subprocess.Popen([shutil.which("ffmpeg"), "-v", "error", "-i", str(path), "-f", "rawvideo", "-pix_fmt", "rgb24", "-"],
                                      stdout=subprocess.PIPE)

    def __len__(self) -> int:
        return 1 << 62

    def frame(self, index: int) -> Optional[np.ndarray]:
        while self.position <= index:
            raw = self.child.stdout.read(self.info.frame_bytes)
            if len(raw) < self.info.frame_bytes:
                return self.last
            self.last, self.position = np.frombuffer(raw, np.uint8), self.position + 1
        return self.last


class CodecFrames:
    """Sequential rgb24 frames of a compressed file through OpenCV's bundled FFmpeg — the same libavcodec / libswscale
    the reference's ffmpeg child runs, in process. Used when there is no `ffmpeg` binary to pipe from"""
    def __init__(self, path, info: VideoInfo):
        import cv2
        self.info, self.position, self.last, self.cv2 = info, 0, None, cv2
        self.capture = cv2.VideoCapture(str(path))
        if not self.capture.isOpened():
            raise RuntimeError(f"OpenCV cannot decode '{path}'")

    def __len__(self) -> int:
        return 1 << 62

    def frame(self, index: int) -> Optional[np.ndarray]:
        while self.position <= index:
            ok, bgr = self.capture.read()
            if not ok:
                return self.last                             # the stream ended: the last frame stays
            self.last, self.position = self.cv2.cvtColor(bgr, self.cv2.COLOR_BGR2RGB).reshape(-1), self.position + 1
        return self.last

    @staticmethod
    def available() -> bool:
        import importlib.util
        return importlib.util.find_spec("cv2") is not None

    @staticmethod
    def probe(path) -> tuple[int, int, float]:
        import cv2
        capture = cv2.VideoCapture(str(path))
        if not capture.isOpened():
            raise RuntimeError(f"OpenCV cannot open '{path}'")
        width, height = int(capture.get(cv2.CAP_PROP_FRAME_WIDTH)), int(capture.get(cv2.CAP_PROP_FRAME_HEIGHT))
        fps = float(capture.get(cv2.CAP_PROP_FPS)) or 25.0
        capture.release()
        return width, height, fps


def probe(path) -> tuple[int, int, float]:
    """(width, height, fps) through ffprobe (ffmpeg.py:1104-1113,1175-1190)"""
    ffprobe = shutil.which("ffprobe")
    if not ffprobe:
        raise RuntimeError(f"Reading '{path}' needs ffmpeg / ffprobe on PATH or the opencv package; uncompressed .y4m / .rgb files are read natively")
    out = subprocess.run([ffprobe, "-v", "error", "-select_streams", "v:0", "-show_entries", "stream=width,height,r_frame_rate",
                          "-of", "csv=p=0", str(path)], capture_output=True, text=True, check=True).stdout.strip().split(",")
    num, den = out[2].split("/")
    return int(out[0]), int(out[1]), int(num)/int(den)


@define
class ShaderVideo(ShaderModule):
    name: str = "iVideo"
    path: Path = None
    texture: ShaderTexture = None
    width: int = None
    height: int = None
    fps: float = None
    pixel_format: str = "rgb24"
    """Layout of a headerless file: 'rgb24' or 'rgba'"""
    bottom_up: bool = False
    """Headerless files: rows run bottom to top (raw exports of this backend)"""

    info: Any = field(default=None, repr=False)
    _reader: Any = field(default=None, repr=False)
    _frames: int = 0
    _staging: Any = field(default=None, repr=False)

    def __attrs_post_init__(self):
        ShaderModule.__attrs_post_init__(self)
        path = Path(self.path)
        suffix = path.suffix.lower()
        if suffix == ".y4m":
            self.info = parse_y4m(path)
            self._reader = FileFrames(path, self.info)
        elif suffix in (".rgb", ".raw", ".rgba"):
            if not (self.width and self.height and self.fps):
                raise RuntimeError("A headerless video file needs width=, height= and fps=")
            fmt = N.VIDEO_RGBA32 if (suffix == ".rgba" or self.pixel_format == "rgba") else N.VIDEO_RGB24
            self.info = raw_info(path, self.width, self.height, self.fps, fmt, self.bottom_up)
            self._reader = FileFrames(path, self.info)
        elif shutil.which("ffmpeg") and shutil.which("ffprobe"):
            width, height, fps = probe(path)
            self.info = VideoInfo(width, height, fps, N.VIDEO_RGB24, width*height*3)
            self._reader = PipeFrames(path, self.info)
        elif CodecFrames.available():
            width, height, fps = CodecFrames.probe(path)
            self.info = VideoInfo(width, height, fps, N.VIDEO_RGB24, width*height*3)
            self._reader = CodecFrames(path, self.info)
        else:
            probe(path)                                       # raises: names what is missing
        self.width, self.height = self.info.width, self.info.height
        self.fps = self.fps or self.info.fps
        # Note: you can set .temporal (video.py:46)
        self.texture = ShaderTexture(scene=self.scene, name=self.name, width=self.width, height=self.height,
                                     dtype=np.uint8, components=3)

    def setup(self):
        self._frames = 0
        if isinstance(self._reader, (PipeFrames, CodecFrames)) and self._reader.position:
            self._reader = type(self._reader)(self.path, self.info)     # sequential decoders start over

    def update(self) -> None:
        # video.py:57-66 — only write a new frame when due; at most one per scene frame
        if not (self.scene.time > (self._frames/self.fps)):
            return
        index = self._frames
        self._frames += 1
        if index >= len(self._reader):
            return                                        # past the end: the last frame stays
        self.texture.roll()
        if self.scene.cuda is None or not self.scene.render_enabled:
            return                                        # dry run, or a frame another rank shades
        frame = self._reader.frame(index)
        if frame is None:
            return
        self._upload(frame)

    def _upload(self, frame: np.ndarray) -> None:
        import torch
        device = f"cuda:{self.scene.device}"
        if self._staging is None:
            self._staging = dict(k=0, pinned=[torch.empty(self.info.frame_bytes, dtype=torch.uint8).pin_memory() for _ in range(2)],
                                 raw=[torch.empty(self.info.frame_bytes, dtype=torch.uint8, device=device) for _ in range(2)],
                                 done=[None, None])
        s = self._staging
        k = s["k"] = s["k"] ^ 1
        if s["done"][k] is not None:
            s["done"][k].synchronize()                     # the H2D out of this pinned buffer has finished
        s["pinned"][k].numpy()[:] = frame                  # page cache → pinned memory
        s["raw"][k].copy_(s["pinned"][k], non_blocking=True)
        s["done"][k] = torch.cuda.Event()
        s["done"][k].record(torch.cuda.current_stream(self.scene.device))
        box = self.texture.get_box()
        self.scene.cuda.video_frame(s["raw"][k], self.info.format, self.width, self.height, self.info.top_down, box.texture)
        box.empty = False

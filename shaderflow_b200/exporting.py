"""ExportingHelper — where finished frames go (API mirror of shaderflow/exporting.py).

The reference reads the GL framebuffer into a ring of mapped buffers and lets turbopipe write them to an
ffmpeg child (exporting.py:140-174). Here the ring lives in libsfb200 (sfb_pipe_*: device frames +
pinned host mirrors + a writer thread); the kernels render straight into the ring's device frame.
Sinks: an ffmpeg child when the binary exists (same rawvideo rgb24 + vflip contract, exporting.py:94-103),
a raw .rgb file, bytes, or a null sink that still pays the device→host copy. Without an ffmpeg binary a video file
is still written when the opencv package is installed: an in-process encoder thread (the libavcodec inside the
OpenCV wheel) takes the child's place at the other end of a pipe — same bytes on the wire, same vflip."""
from __future__ import annotations

import os
import shutil
import subprocess
import tempfile
import time
from enum import Enum
from pathlib import Path
from typing import TYPE_CHECKING, Any, Callable, Optional

from attrs import Factory, define

from shaderflow_b200 import _native as N
from shaderflow_b200 import logger

if TYPE_CHECKING:
    from shaderflow_b200.scene import ShaderScene


class OutputType(str, Enum):
    PATH = "file"
    PIPE = "pipe"
    RAW = "raw"
    NULL = "null"
    TCP = "tcp"


class InProcessEncoder:
    """Stands where the ffmpeg child stands when there is no ffmpeg binary: reads rawvideo rgb24 frames (bottom row
    first) from a pipe, flips them (the child's `-vf vflip`) and encodes them with `cv2.VideoWriter` — FFmpeg's
    libavcodec / libavformat as bundled by the OpenCV wheel. Codec by container: lossless FFV1 for .mkv, Motion-JPEG
    for .avi, MPEG-4 for everything else (the wheel carries no H.264 encoder). Video only: audio is not muxed."""
    CODECS = {".mkv": "FFV1", ".avi": "MJPG"}

    def __init__(self, path: Path, width: int, height: int, fps: float):
        import threading
        import cv2
        self.path, self.width, self.height, self.frames, self.error = Path(path), width, height, 0, None
        fourcc = self.CODECS.get(self.path.suffix.lower(), "mp4v")
        self.writer = cv2.VideoWriter(str(self.path), cv2.VideoWriter_fourcc(*fourcc), float(fps), (width, height))
        if not self.writer.isOpened():
            raise RuntimeError(f"OpenCV cannot encode '{self.path}' ({fourcc})")
        self.read_fd, self.write_fd = os.pipe()
        try:
            import fcntl
            fcntl.fcntl(self.write_fd, 1031, 1 << 20)          # F_SETPIPE_SZ: fewer wake-ups per 25 MB frame
        except Exception:
            pass
        self.thread = threading.Thread(target=self._run, name="sfb-encoder", daemon=True)
        self.thread.start()

    @staticmethod
    def available() -> bool:
        import importlib.util
        return importlib.util.find_spec("cv2") is not None

    def _run(self) -> None:
        import numpy as np
        size = self.width*self.height*3
        frame = np.empty(size, np.uint8)
        view = memoryview(frame)
        try:
            with os.fdopen(self.read_fd, "rb", buffering=0) as source:
                while True:
                    got = 0
                    while got < size:
                        n = source.readinto(view[got:])
                        if not n:
                            break
                        got += n
                    if got < size:
                        break                                   # the writer closed its end (a partial frame is dropped)
                    image = frame.reshape(self.height, self.width, 3)[::-1, :, ::-1]     # vflip, rgb → bgr
                    self.writer.write(np.ascontiguousarray(image))
                    self.frames += 1
        except Exception as error:                              # surfaced by ExportingHelper.check_process
            self.error = error
        finally:
            self.writer.release()

    def poll(self) -> None:
        if self.error is not None:
            raise RuntimeError(f"The in-process encoder stopped: {self.error}")

    def close(self) -> int:
        """The last frame has been written to the pipe: close it, wait for the encoder, → frames encoded"""
        if self.write_fd is not None:
            os.close(self.write_fd)
            self.write_fd = None
        self.thread.join()
        self.poll()
        return self.frames


@define
class ExportingHelper:
    scene: "ShaderScene"
    type: Optional[OutputType] = None
    frame: int = 0
    start: float = Factory(time.monotonic)
    relay: Optional[Callable[[int, int], None]] = None
    bar: Any = None
    process: Optional[subprocess.Popen] = None
    stdout: Any = None
    stderr: Any = None
    fileno: Optional[int] = None
    output: Any = None
    pipe_handle: Optional[N.Pipe] = None
    buffers: int = 5
    took: Optional[float] = None
    _file: Any = None
    _target: Optional[int] = None
    encoder: Optional[InProcessEncoder] = None

    pipe_output = property(lambda self: self.type is OutputType.PIPE)
    path_output = property(lambda self: self.type is OutputType.PATH)
    tcp_output = property(lambda self: self.type is OutputType.TCP)

    @property
    def total_frames(self) -> int:
        return max(1, round(self.scene.runtime*self.scene.fps))

    @property
    def frame_bytes(self) -> int:
        return self.scene.width*self.scene.height*3

    def open_bar(self) -> None:
        try:
            import tqdm
            quiet = bool(self.relay) or self.scene.realtime or os.environ.get("SHADERFLOW_PROGRESS", "1") == "0"
            self.bar = tqdm.tqdm(total=self.total_frames, disable=(True if quiet else None), desc=f"Scene ({self.scene.name}) → Video",
                                 unit=" frames", dynamic_ncols=True, mininterval=1/30, maxinterval=0.5, smoothing=0.1, leave=False)
        except Exception:
            self.bar = None

    def update(self) -> None:
        if self.relay: self.relay(self.frame, self.total_frames)
        if self.bar: self.bar.update(1)
        self.frame += 1

    @property
    def finished(self) -> bool:
        return self.frame >= self.total_frames

    # -- sink configuration ----------------------------------------------------------------------
    def ffmpeg_output(self, output) -> None:
        """Decides the sink from `output` (exporting.py:105-117 plus the sinks that need no ffmpeg)"""
        if output in ("pipe", "-", bytes):
            self.type = OutputType.PIPE
        elif str(output) in ("null", os.devnull):
            self.type = OutputType.NULL
        elif "tcp://" in str(output):
            raise NotImplementedError
        elif Path(str(output)).suffix.lower() in (".rgb", ".raw", ".rgb24"):
            self.type = OutputType.RAW
            self.output = Path(output).expanduser().absolute()
            self.output.parent.mkdir(parents=True, exist_ok=True)
        else:
            self.type = OutputType.PATH
            self.output = Path(output).expanduser().absolute()
            self.output.parent.mkdir(parents=True, exist_ok=True)

    def ffmpeg_command(self) -> list[str]:
        """rawvideo rgb24 frames, bottom row first, on stdin; vflip restores top-down
        (exporting.py:94-103). Codec selection is ffmpeg's default for the container."""
        s = self.scene
        command = [shutil.which("ffmpeg") or "ffmpeg", "-hide_banner", "-loglevel", "error", "-y",
                   "-f", "rawvideo", "-pix_fmt", "rgb24", "-s", f"{s.width}x{s.height}", "-r", f"{s.fps}", "-i", "-"]
        for module in s.modules:
            extra = module.ffhook(self)
            if extra:
                command += list(extra)
        command += ["-vf", "vflip", "-t", f"{s.runtime}"]
        command += ["-f", "matroska", "-"] if self.pipe_output else [str(self.output)]
        return command

    def open_fd(self) -> int:
        """Opens what the frames are written to (ffmpeg child / file / temporary file) → its descriptor, -1 for
        the null sink. The sharded export's writer (csrc/sink.cu) takes the descriptor as is."""
        fd = -1
        if self.type is OutputType.RAW:
            self._file = open(self.output, "wb")
            fd = self._file.fileno()
        elif self.type in (OutputType.PATH, OutputType.PIPE):
            if not shutil.which("ffmpeg"):
                if self.type is OutputType.PIPE:     # raw rgb24 bytes instead of an encoded stream
                    self._file = tempfile.TemporaryFile(mode="w+b")
                    fd = self._file.fileno()
                elif InProcessEncoder.available():
                    s = self.scene
                    if any(module.ffhook(self) for module in s.modules):
                        logger.warn("No ffmpeg binary: the in-process encoder writes video only, the audio is not muxed")
                    self.encoder = InProcessEncoder(self.output, s.width, s.height, s.fps)
                    fd = self.encoder.write_fd
                else:
                    raise RuntimeError(logger.error(
                        f"No ffmpeg binary on PATH (nor the opencv package) to encode '{self.output}'. Export raw frames "
                        "with output='frames.rgb', bytes with output='pipe', or discard them with output='null'"))
            else:
                self.stderr, self.stdout = tempfile.TemporaryFile(mode="r+b"), tempfile.TemporaryFile(mode="r+b")
                self.process = subprocess.Popen(self.ffmpeg_command(), stdin=subprocess.PIPE,
                                                stdout=self.stdout, stderr=self.stderr)
                fd = self.process.stdin.fileno()
        self.fileno = fd
        return fd

    def check_process(self) -> None:
        if self.encoder is not None:
            self.encoder.poll()
        if self.process is not None and self.process.poll() is not None:
            self.stderr.seek(0)
            raise RuntimeError("FFmpeg process closed unexpectedly with traceback:\n" + self.stderr.read().decode("utf-8"))

    def abort(self) -> None:
        """An export that failed half way: no child process, no acquired ring slot and no open file may survive"""
        if self.process is not None:
            try:
                self.process.stdin.close()
            except Exception:
                pass
            self.process.kill()
            self.process.wait()
            self.process = None
        if self.encoder is not None:
            try: self.encoder.close()
            except Exception: pass
            self.encoder = None
        if self._file is not None:
            try: self._file.close()
            except Exception: pass
            self._file = None
        ring = getattr(self.scene, "_sink_ring", None)
        if ring is not None:                         # its `acquired` flag / in-flight frames are unknown: drop it
            self.scene._sink_ring = None
            try: ring.close()
            except Exception: pass
        self.pipe_handle, self._target = None, None
        if self.bar is not None:
            self.bar.close()

    def popen(self) -> None:
        fd = self.open_fd()
        # one ring per scene serves consecutive exports: creating pinned buffers, a stream and a thread per
        # main() call is measurable on short exports
        buffers = max(2, int(self.buffers))
        ring = getattr(self.scene, "_sink_ring", None)
        if ring is not None and ring.handle and ring.frame_bytes == self.frame_bytes and ring.buffers == buffers:
            ring.set_fd(fd)
        else:
            if ring is not None:
                ring.close()
            ring = N.Pipe(self.scene.cuda, fd, buffers, self.frame_bytes)
            self.scene._sink_ring = ring
        self.pipe_handle = ring

    def make_buffers(self, n: int = 2) -> None:
        self.buffers = n

    # -- per frame -------------------------------------------------------------------------------
    def target(self) -> Optional[int]:
        """Device pointer the next frame is rendered into (a frame of the sink's ring)"""
        if self.pipe_handle is None:
            return None
        if self._target is None:
            self._target = self.pipe_handle.acquire()
        return self._target

    def pipe(self, turbo: bool = True) -> None:
        """Hands the frame just rendered to the sink (exporting.py:151-174)"""
        if self.pipe_handle is None:
            return
        self.check_process()
        self.pipe_handle.submit(self._target)
        self._target = None
        if not turbo:
            self.pipe_handle.sync()

    def finish(self) -> None:
        if self.pipe_handle is not None:
            self.pipe_handle.sync()                 # every frame written; the ring itself stays with the scene
            self.pipe_handle.set_fd(-1)
            self.pipe_handle = None
        if self.process is not None:
            logger.info("Waiting for FFmpeg process to finish encoding")
            self.process.stdin.close()
            self.process.wait()
            self.stdout.seek(0)
        if self.encoder is not None:
            self.encoder.close()
            self.encoder = None
        if self._file is not None and self.type is OutputType.RAW:
            self._file.close()
        if self.bar is not None:
            self.bar.close()
        self.took = time.monotonic() - self.start

    def result(self):
        if self.type is OutputType.PIPE:
            if self.process is not None:
                return self.stdout.read()
            self._file.seek(0)
            data = self._file.read()
            self._file.close()
            return data
        if self.type in (OutputType.PATH, OutputType.RAW):
            return self.output
        return None

    def log_stats(self, output) -> None:
        if self.scene.exporting:
            logger.info(f"Finished rendering ({output})")
        logger.info(f"• Stats: (Took {self.took:.2f}s) at ({self.frame/self.took:.2f}fps | "
                    f"{self.scene.runtime/self.took:.2f}x Realtime) with ({self.frame} Total Frames)")

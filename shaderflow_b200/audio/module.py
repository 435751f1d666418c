"""BrokenAudio / ShaderAudio (API mirror of shaderflow/audio/module.py).

B200-first change: an export knows its whole clip, so instead of a 30 s host ring that is np.roll'ed
every frame (audio/module.py:113-129 — the largest CPU cost of the reference's audio side) the clip is
uploaded once and stays resident in HBM as planar float32; every per-frame quantity is computed for ALL
frames at once by csrc/audio.cu and published frame by frame. `tell`, `get_last_n_samples` & co. keep
their meaning (they read the clip at the current frame's position).

Audio enters through `load(pcm, samplerate)` (numpy, shape (channels, samples)), a WAV / FLAC file, or — when
an ffmpeg binary exists — any file ffmpeg decodes. A scene that only knows its audio frame by frame (a synthesiser in
`update()`) feeds `add_data(chunk)` like the reference: the clip then GROWS in HBM (only the new samples are uploaded)
and the per-frame tracks are recomputed by the same batch kernels over the frames seen so far, so a streamed export
equals the batch export of the same samples. Soundcard capture / playback is out of scope."""
from __future__ import annotations

import math
import shutil
import subprocess
from enum import Enum
from pathlib import Path
from typing import Any, Iterable, Optional

import numpy as np
from attrs import define, field

from shaderflow_b200 import _native as N
from shaderflow_b200 import logger
from shaderflow_b200.dynamics import ShaderDynamics
from shaderflow_b200.module import ShaderModule


def root_mean_square(data) -> float:
    return np.sqrt(np.mean(np.square(data)))


class AudioMode(Enum):
    Realtime = "realtime"
    File = "file"


def read_audio_file(path: Path) -> tuple[np.ndarray, int]:
    """→ (pcm float32 (channels, samples), samplerate). RIFF/WAVE and FLAC natively (audio/reader.py), anything else
    through ffmpeg when a binary exists, else through FFmpeg's libraries in process (audio/avcodec.py)"""
    path = Path(path)
    if path.suffix.lower() in (".wav", ".wave", ".rf64"):
        from shaderflow_b200.audio.reader import read_wav
        return read_wav(path)
    if path.suffix.lower() in (".flac", ".fla"):
        from shaderflow_b200.audio.reader import read_flac
        return read_flac(path)
    ffmpeg, ffprobe = shutil.which("ffmpeg"), shutil.which("ffprobe")
    if not (ffmpeg and ffprobe):
        from shaderflow_b200.audio import avcodec
        if avcodec.available():                  # the same libavcodec / libswresample, in process (OpenCV wheel)
            return avcodec.decode_file(path)
        raise RuntimeError(f"Decoding '{path}' needs ffmpeg/ffprobe on PATH or the opencv package; load WAV / FLAC files "
                           "or call load(pcm, samplerate)")
    probe = subprocess.run([ffprobe, "-v", "error", "-select_streams", "a:0", "-show_entries",
        "stream=channels,sample_rate", "-of", "csv=p=0", str(path)], capture_output=True, text=True, check=True)
    rate, channels = [int(x) for x in probe.stdout.strip().split(",")][:2]
    raw = subprocess.run([ffmpeg, "-v", "error", "-i", str(path), "-f", "f32le", "-vn", "-"], capture_output=True, check=True).stdout
    return np.ascontiguousarray(np.frombuffer(raw, np.float32).reshape(-1, channels).T), rate


@define(slots=False)
class BrokenAudio:
    mode: AudioMode = field(default=AudioMode.Realtime, converter=AudioMode)
    dtype: Any = np.float32
    tell: int = 0
    """Samples consumed so far (at the current frame)"""

    clip: Optional[np.ndarray] = field(default=None, repr=False)
    """The whole clip on the host, (channels, samples) float32"""
    clip_device: Any = field(default=None, repr=False)
    """The same clip resident in HBM (torch tensor), uploaded on first use"""

    _samplerate: float = 44100
    _channels: int = 2
    _buffer_seconds: float = 30.0
    _file: Optional[Path] = None

    def __attrs_post_init__(self):
        pass

    # -- the clip --------------------------------------------------------------------------------
    def load(self, pcm: np.ndarray, samplerate: Optional[float] = None) -> "BrokenAudio":
        pcm = np.asarray(pcm, dtype=np.float32)
        if pcm.ndim == 1:
            pcm = pcm[None, :]
        self.clip = np.ascontiguousarray(pcm)
        self.clip_device = None
        self._wav = None
        self.streaming, self._stream_host, self._stream_uploaded = False, None, 0
        self._channels = int(pcm.shape[0])
        if samplerate:
            self._samplerate = samplerate
        self.mode = AudioMode.File
        self.tell = 0
        return self

    _wav: Any = field(default=None, repr=False)
    """audio/reader.WavInfo of the file the clip came from: the device copy is then made from the file's own sample
    bytes (pinned staging → H2D → sfb_pcm_ingest) instead of from the host float32 array"""

    def device_clip(self, device: int):
        if self.clip is None:
            raise RuntimeError("No audio loaded: set `file=` to an existing file or call load(pcm, samplerate)")
        if self.streaming:
            ctx = getattr(getattr(self, "scene", None), "cuda", None)
            return self._stream_device_clip(getattr(ctx, "torch_device", f"cuda:{device}"))
        if self.clip_device is None:
            import torch
            ctx = getattr(getattr(self, "scene", None), "cuda", None)
            if self._wav is not None and ctx is not None:
                from shaderflow_b200.audio.reader import upload_wav
                self.clip_device = upload_wav(ctx, self._wav, device)
            else:
                self.clip_device = torch.from_numpy(self.clip).to(f"cuda:{device}", non_blocking=False)
        return self.clip_device

    @property
    def total_samples(self) -> int:
        return 0 if self.clip is None else int(self.clip.shape[1])

    @property
    def data(self) -> np.ndarray:
        """What the reference's ring would hold now: the last `buffer_size` samples up to `tell`"""
        out = np.zeros(self.shape, dtype=self.dtype)
        if self.clip is not None and self.tell > 0:
            n = min(self.tell, self.buffer_size)
            out[:, self.buffer_size - n:] = self.clip[:, self.tell - n:self.tell]
        return out

    # -- streaming ---------------------------------------------------------------------------------
    streaming: bool = False
    """True once add_data() has been called: the clip is what was added so far and `tell` its length"""
    _stream_host: Any = field(default=None, repr=False)
    """(channels, capacity) float32; `clip` is the view of its first `tell` samples"""
    _stream_uploaded: int = 0
    """Samples of the stream that already are in `clip_device` (channels, device capacity)"""

    def add_data(self, data: np.ndarray) -> Optional[np.ndarray]:
        """audio/module.py:113-129: append (channels, length) samples; `tell` advances by length. Instead of rolling
        a 30 s ring the samples extend the clip (amortised doubling), so `data`, `get_last_n_samples` & co. read
        what the reference's ring would hold and the kernels index the stream by absolute sample"""
        data = np.array(data, dtype=np.float32)
        if data.ndim == 1:
            data = data[None, :]
        if not self.streaming:
            if self.clip is not None:
                raise RuntimeError("add_data() after load() / file=: a clip is either loaded whole or streamed")
            self._channels = int(data.shape[0])
            self._stream_host = np.zeros((self._channels, max(1 << 16, 2*data.shape[1])), dtype=np.float32)
            self._stream_uploaded, self.clip_device, self._wav = 0, None, None
            self.streaming, self.tell = True, 0
        if data.shape[0] != self._stream_host.shape[0]:
            raise ValueError(f"add_data(): {data.shape[0]} channels into a stream of {self._stream_host.shape[0]}")
        length, have = int(data.shape[1]), int(self.tell)
        if have + length > self._stream_host.shape[1]:
            grown = np.zeros((self._stream_host.shape[0], max(2*self._stream_host.shape[1], have + length)), dtype=np.float32)
            grown[:, :have] = self._stream_host[:, :have]
            self._stream_host = grown
        self._stream_host[:, have:have + length] = data
        self.tell = have + length
        self.clip = self._stream_host[:, :self.tell]
        return data

    def _stream_device_clip(self, device: str):
        """The stream in HBM: (channels, capacity) planar float32, zero beyond `tell`; only samples added since the
        last call cross PCIe. The kernels take the capacity as the plane stride and never read at or past `tell`"""
        import torch
        have, buf = int(self.tell), self.clip_device
        if buf is None or buf.shape[1] < have or str(buf.device) != str(torch.device(device)):
            grown = torch.zeros((self._stream_host.shape[0], max(1 << 20, 2*have)), dtype=torch.float32, device=device)
            if buf is not None and str(buf.device) == str(grown.device):
                grown[:, :self._stream_uploaded] = buf[:, :self._stream_uploaded]
            else:
                self._stream_uploaded = 0
            buf = self.clip_device = grown
        if have > self._stream_uploaded:
            fresh = np.ascontiguousarray(self._stream_host[:, self._stream_uploaded:have])
            buf[:, self._stream_uploaded:have] = torch.from_numpy(fresh).to(device)
            self._stream_uploaded = have
        return buf

    def get_last_n_samples(self, n: int, *, offset: int = 0) -> np.ndarray:
        """audio/module.py:137-138: the newest sample is excluded"""
        n, offset = int(n), int(offset)
        hi = int(self.tell) - offset - 1
        lo = hi - n
        out = np.zeros((self.channels, n), dtype=self.dtype)
        if self.clip is not None:
            a, b = max(lo, 0), max(hi, 0)
            if b > a:
                out[:, a - lo:b - lo] = self.clip[:, a:b]
        return out

    def get_last_n_seconds(self, n: float) -> np.ndarray:
        return self.get_last_n_samples(n*self.samplerate)

    def get_data_between_samples(self, start: int, end: int) -> np.ndarray:
        return self.data[:, int(start):int(end)]

    def get_data_between_seconds(self, start: float, end: float) -> np.ndarray:
        return self.get_data_between_samples(start*self.samplerate, end*self.samplerate)

    # -- properties ------------------------------------------------------------------------------
    @property
    def buffer_size(self) -> int:
        return int(self.samplerate*self.buffer_seconds)

    @property
    def shape(self) -> tuple[int, int]:
        return (self.channels, self.buffer_size)

    @property
    def samplerate(self) -> float:
        return self._samplerate or 44100

    @samplerate.setter
    def samplerate(self, value: float):
        self._samplerate = value

    @property
    def channels(self) -> int:
        return self._channels or 2

    @channels.setter
    def channels(self, value: int):
        self._channels = value

    @property
    def buffer_seconds(self) -> float:
        return self._buffer_seconds

    @buffer_seconds.setter
    def buffer_seconds(self, value: float):
        self._buffer_seconds = value

    @property
    def file(self) -> Optional[Path]:
        return self._file

    @file.setter
    def file(self, value):
        if value is None:
            return
        self._file = Path(value)
        if not self._file.exists():
            logger.warn(f"Audio File doesn't exist ({value})")
            return
        pcm, rate = read_audio_file(self._file)
        self.load(pcm, rate)
        if self._file.suffix.lower() in (".wav", ".wave", ".rf64"):
            from shaderflow_b200.audio.reader import parse_wav
            self._wav = parse_wav(self._file)

    stereo = property(lambda self: self.channels == 2)
    mono = property(lambda self: self.channels == 1)

    @property
    def duration(self) -> float:
        if self.clip is None:
            return math.inf if self.mode == AudioMode.Realtime else 0.0
        return self.total_samples/self.samplerate

    # realtime devices: out of scope
    def open_recorder(self, *a, **k):
        raise NotImplementedError("Realtime capture is out of scope of the CUDA offline backend")
    open_speaker = open_recorder

    def close_recorder(self): return self
    close_speaker = close_recorder

    def play(self, data) -> None: ...


@define
class ShaderAudio(BrokenAudio, ShaderModule):
    volume: ShaderDynamics = None
    std: ShaderDynamics = None
    final: bool = True

    clock: Any = field(default=None, repr=False)
    """(time, dt, tell) host arrays of the current export + their device copies"""
    scalars: Optional[np.ndarray] = field(default=None, repr=False)
    """[frames][SFB_SCALARS] float64 from the GPU scan (volume, integral, std, targets)"""

    def __attrs_post_init__(self):
        requested = self._file
        self._file = None
        BrokenAudio.__attrs_post_init__(self)
        ShaderModule.__attrs_post_init__(self)
        self.volume = ShaderDynamics(scene=self.scene, name=f"{self.name}Volume",
            frequency=2, zeta=1, response=0, value=0, integrate=True)
        self.std = ShaderDynamics(scene=self.scene, name=f"{self.name}STD",
            frequency=10, zeta=1, response=0, value=0)
        if requested is not None:
            self.file = requested

    @property
    def duration(self) -> float:
        if self.clip is None or self.streaming:      # a stream has no length of its own: the scene's runtime decides
            return 0.0
        return self.total_samples/self.samplerate

    stream: Any = field(default=None, repr=False)
    """Streaming only: dict(frames, tell, dt, tell_device, dt_device) — the clock of the frames seen so far this
    export, recorded as they happen (there is no clip to derive it from); the other audio modules read it"""

    def setup(self):
        self.clock, self.scalars, self.stream = None, None, None
        if not self.streaming:
            self.tell = 0                            # a stream keeps its position: the samples were fed by the scene

    def ffhook(self, ffmpeg) -> Optional[list]:
        """Mux the audio file into the export (audio/module.py:441-444)"""
        if self._file is not None and Path(self._file).exists():
            return ["-i", str(self._file), "-shortest"]

    # -- the batch -------------------------------------------------------------------------------
    def prepare(self) -> None:
        """Frame clock for the whole export + the volume/std track, once per main()"""
        import torch
        scene = self.scene
        frames = scene.total_frames
        time, dt, tell = N.frame_clock(frames, scene.fps, scene.speed, int(self.samplerate), self.channels, self.total_samples)
        dev = f"cuda:{scene.device}"
        self.clock = dict(frames=frames, time=time, dt=dt, tell=tell,
                          tell_device=torch.from_numpy(tell).to(dev), dt_device=torch.from_numpy(dt).to(dev))
        scalars = torch.zeros((frames, N.SCALARS), dtype=torch.float64, device=dev)
        scene.cuda.audio_track(self.device_clip(scene.device), int(self.samplerate), self.clock["tell_device"],
                               self.clock["dt_device"], scalars=scalars)
        self.scalars = scalars.cpu().numpy()      # one small D2H for the whole export

    def _record_stream_frame(self) -> dict:
        """One more frame of a streamed export: `tell` is what add_data() left, dt is the scene's (dynamics.py:197)"""
        import torch
        st = self.stream
        if st is None:
            st = self.stream = dict(frames=0, tell=np.zeros(256, np.int64), dt=np.zeros(256, np.float64))
        k = st["frames"]
        if k == st["tell"].shape[0]:
            st["tell"], st["dt"] = (np.concatenate([a, np.zeros_like(a)]) for a in (st["tell"], st["dt"]))
        st["tell"][k], st["dt"][k] = int(self.tell), float(self.scene.dt)
        st["frames"] = k + 1
        dev = getattr(self.scene.cuda, "torch_device", f"cuda:{self.scene.device}")
        # the whole prefix again: 16 B per frame, and the tracks below are recomputed over it anyway
        st["tell_device"] = torch.from_numpy(st["tell"][:k + 1].copy()).to(dev)
        st["dt_device"] = torch.from_numpy(st["dt"][:k + 1].copy()).to(dev)
        return st

    def _update_stream(self) -> None:
        """add_data() mode: volume / std of this frame from the batch kernel run over every frame so far (the two
        ShaderDynamics are recurrences from frame 0; the scan is the one a whole-clip export runs, so both agree)"""
        import torch
        scene = self.scene
        st = self._record_stream_frame()
        if not scene.render_enabled:
            return
        frames = st["frames"]
        scalars = torch.zeros((frames, N.SCALARS), dtype=torch.float64, device=st["tell_device"].device)
        scene.cuda.audio_track(self.device_clip(scene.device), int(self.samplerate), st["tell_device"], st["dt_device"],
                               scalars=scalars)
        self._publish(scalars[frames - 1].cpu().numpy())

    def update(self):
        if self.clip is None or self.scene.cuda is None:
            return
        if self.streaming:
            return self._update_stream()
        if self.clock is None or self.clock["frames"] != self.scene.total_frames:
            self.prepare()
        k = min(self.scene.frame_index, self.clock["frames"] - 1)
        self.tell = int(self.clock["tell"][k])
        if not self.scene.render_enabled:
            return                                   # a frame another rank shades: the tracks hold every frame's state
        self._publish(self.scalars[k])

    def _publish(self, row) -> None:
        # publish the GPU scan's state for this frame; the ShaderDynamics then only emit uniforms
        for dyn, value in ((self.volume, row[N.SCALAR_VOLUME]), (self.std, row[N.SCALAR_STD])):
            dyn.published = True
            dyn.value = np.array(value, dtype=np.float64)
        self.volume.integral = np.array(row[N.SCALAR_VOLUME_INTEGRAL], dtype=np.float64)
        self.volume.target = np.array(row[N.SCALAR_VOLUME_TARGET], dtype=np.float32)
        self.std.target = np.array(row[N.SCALAR_STD_TARGET], dtype=np.float32)

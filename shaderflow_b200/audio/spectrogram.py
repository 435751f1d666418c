"""BrokenSpectrogram / ShaderSpectrogram (API mirror of shaderflow/audio/spectrogram.py).

Configuration (fft_n, scale, interpolation, magnitude, window, volume, bins, frequency range,
from_notes) is the reference's; the filterbank matrix is built on the host once per configuration
(it is setup work: spectrogram.py:194-224) and shipped to the GPU as CSR. The per-frame work —
window·rfft·magnitude, filterbank product, per-bin second-order smoothing — runs for ALL frames of the
export in two launches of csrc/audio.cu (sfb_stft_mel, sfb_audio_track) on first use, and each frame's
update() just points the RG32F texture at its row of that device-resident track (zero copy)."""
from __future__ import annotations

import functools
import math
from typing import Any, Callable, Iterable, Optional

import numpy as np
from attrs import Factory, define, field

from shaderflow_b200 import _native as N
from shaderflow_b200 import logger
from shaderflow_b200.audio.module import BrokenAudio
from shaderflow_b200.dynamics import DynamicNumber
from shaderflow_b200.module import ShaderModule
from shaderflow_b200.piano.notes import PianoNote
from shaderflow_b200.texture import ShaderTexture
from shaderflow_b200.variable import ShaderVariable, Uniform


class FourierMagnitude:
    """Interpretation of the complex FFT bins; `kind` selects the kernel's epilogue"""
    def Amplitude(x): return np.abs(x)
    def Power(x): return (x*x.conjugate()).real
    Amplitude.kind, Power.kind = N.MAGNITUDE["amplitude"], N.MAGNITUDE["power"]


class FourierVolume:
    """Final mapping of a spectrogram bin. The reference's next() never applies it (dead code after the
    early return at spectrogram.py:176), so ShaderSpectrogram defaults to Linear; the others are the
    kernel's opt-in epilogue (`apply_volume=True`)"""
    def dBFS(x): return 10*np.log10(x)
    def Sqrt(x): return np.sqrt(x)
    def Linear(x): return x
    def dBFsTremx(x): return 10*(np.log10(x + 0.1) + 1)/1.0414
    dBFS.kind, Sqrt.kind, Linear.kind, dBFsTremx.kind = (N.VOLUME[k] for k in ("dbfs", "sqrt", "linear", "dbfs_tremx"))


class SpectrogramInterpolation:
    """Discrete → continuous interpolation kernels for the filterbank rows"""
    def make_euler(end: float = 1.54) -> Callable:
        return (lambda x: np.exp(-(2*x/end)**2)/(end*(math.pi**0.5)))

    def Dirac(x):
        out = np.zeros(x.shape)
        out[np.round(x) == 0] = 1
        return out

    Euler = make_euler(end=1.2)

    def Sinc(x): return np.abs(np.sinc(x))


class SpectrogramScale:
    """(forward, inverse) maps of the frequency axis"""
    Octave = ((lambda x: np.log(x)/np.log(2)), (lambda x: 2**x))
    MEL = ((lambda x: 2595*np.log10(1 + x/700)), (lambda x: 700*(10**(x/2595) - 1)))


class SpectrogramWindow:
    @functools.lru_cache
    def hann_poisson_window(N_: int, alpha: float = 2.0) -> np.ndarray:
        n = np.arange(N_)
        return (0.5*(1 - np.cos(2*np.pi*n/N_)))*np.exp(-alpha*np.abs(N_ - 2*n)/N_)

    @functools.lru_cache
    def hanning(size: int) -> np.ndarray:
        return np.hanning(size)

    @functools.lru_cache
    def none(size: int) -> np.ndarray:
        return np.ones(size)

    hanning.__wrapped__.kind = N.WINDOW["hanning"]
    hann_poisson_window.__wrapped__.kind = N.WINDOW["hann_poisson"]
    none.__wrapped__.kind = N.WINDOW["none"]


def _kind(fn, default: int = 0) -> int:
    return getattr(getattr(fn, "__wrapped__", fn), "kind", default)


@define(eq=False, slots=False)
class BrokenSpectrogram:
    audio: BrokenAudio = Factory(BrokenAudio)
    fft_n: int = field(default=12, converter=int)
    sample_rateio: int = field(default=1, converter=int)
    scale: tuple = SpectrogramScale.Octave
    interpolation: Callable = SpectrogramInterpolation.Euler
    magnitude: Callable = FourierMagnitude.Power
    window: Callable = SpectrogramWindow.hanning
    volume: Callable = FourierVolume.Sqrt
    minimum_frequency: float = 20.0
    maximum_frequency: float = 20000.0
    spectrogram_bins: int = 1000

    def config_key(self) -> tuple:
        return (self.fft_n, self.minimum_frequency, self.maximum_frequency, self.spectrogram_bins,
                self.sample_rateio, self.magnitude, self.interpolation, self.scale, self.volume,
                self.window, self.audio.samplerate)

    @property
    def fft_size(self) -> int:
        return int(2**(self.fft_n)*self.sample_rateio)

    @property
    def fft_bins(self) -> int:
        return int(self.fft_size/2 + 1)

    @property
    def fft_frequencies(self) -> np.ndarray:
        return np.fft.rfftfreq(self.fft_size, 1/(self.audio.samplerate*self.sample_rateio))

    @property
    def spectrogram_frequencies(self) -> np.ndarray:
        return self.scale[1](np.linspace(self.scale[0](self.minimum_frequency),
                                         self.scale[0](self.maximum_frequency), self.spectrogram_bins))

    def spectrogram_matrix(self):
        """(bins × fft_bins) float32 CSR: row b interpolates the FFT around centre frequency b
        (Whittaker–Shannon with the chosen kernel), |m| < 1e-5 dropped. Cached per configuration"""
        key = self.config_key()
        cached = getattr(self, "_matrix_cache", None)
        if cached is None or cached[0] != key:
            import scipy.sparse
            step = self.fft_frequencies[1]
            columns = np.arange(self.fft_bins)
            dense = np.array([self.interpolation(centre - columns) for centre in (self.spectrogram_frequencies/step)],
                             dtype=self.audio.dtype)
            dense[np.abs(dense) < 1e-5] = 0
            object.__setattr__(self, "_matrix_cache", (key, scipy.sparse.csr_matrix(dense)))
        return self._matrix_cache[1]

    def from_notes(self, start, end, bins: int = 1000, piano: bool = False, tuning: float = 440):
        start, end = PianoNote.get(start, tuning=tuning), PianoNote.get(end, tuning=tuning)
        logger.info(f"Making Spectrogram Piano Matrix from notes ({start.name} - {end.name})")
        self.minimum_frequency, self.maximum_frequency = start.frequency, end.frequency
        if not piano:
            self.spectrogram_bins = bins
        else:   # one bin per key, range widened by half a semitone each side
            half = 2**(0.5/12)
            self.spectrogram_bins = (end.note - start.note) + 1
            self.minimum_frequency /= half
            self.maximum_frequency *= half

    # -- GPU --------------------------------------------------------------------------------------
    def device_bank(self, device: str):
        import torch
        key = (self.config_key(), device)
        cached = getattr(self, "_bank_cache", None)
        if cached is None or cached[0] != key:
            m = self.spectrogram_matrix()
            bank = tuple(torch.from_numpy(np.ascontiguousarray(a)).to(device)
                         for a in (m.indptr.astype(np.int32), m.indices.astype(np.int32), m.data.astype(np.float32)))
            object.__setattr__(self, "_bank_cache", (key, (*bank, self.spectrogram_bins)))
        return self._bank_cache[1]

    def track(self, cuda: N.Context, pcm, tell, *, apply_volume: bool = False, want_magnitude: bool = False):
        """STFT → filterbank for every frame in `tell` at once → (spec [F, bins, ch], magnitude or None)"""
        import torch
        if self.sample_rateio != 1:
            raise NotImplementedError("sample_rateio != 1 (libsamplerate resampling) is out of scope")
        frames, ch = int(tell.shape[0]), int(pcm.shape[0])
        spec = torch.empty((frames, self.spectrogram_bins, ch), dtype=torch.float32, device=pcm.device)
        mag = torch.empty((frames, ch, self.fft_bins), dtype=torch.float32, device=pcm.device) if want_magnitude else None
        cuda.stft_mel(pcm, tell, self.fft_n, self.device_bank(str(pcm.device)), window=_kind(self.window),
                      magnitude=_kind(self.magnitude), volume=_kind(self.volume) if apply_volume else 0,
                      mag_out=mag, spec_out=spec)
        return spec, mag


@define(eq=False)
class ShaderSpectrogram(BrokenSpectrogram, ShaderModule):
    name: str = "iSpectrogram"
    length: float = 5
    offset: int = 0
    smooth: bool = False
    scrolling: bool = False
    dynamics: DynamicNumber = None
    texture: ShaderTexture = None
    apply_volume: bool = False
    """Opt-in: run `volume` as the kernel's epilogue (the reference never does, see FourierVolume)"""
    columns: Any = field(default=None, repr=False)
    """[frames][bins][channels] float32 device track (smoothed texture columns) of the current export"""

    @property
    def length_samples(self) -> int:
        return int(max(1, self.length*self.scene.fps))

    def __attrs_post_init__(self):
        ShaderModule.__attrs_post_init__(self)
        self.dynamics = DynamicNumber(frequency=4, zeta=1, response=0, dtype=np.float32)
        self.texture = ShaderTexture(scene=self.scene, name=self.name, dtype=np.float32, repeat_y=False)

    def setup(self):
        self.columns, self.offset, self._stream_raw = None, 0, None

    _stream_raw: Any = field(default=None, repr=False)
    """Streaming only: (frames done, [capacity][bins][channels] unsmoothed columns) — STFT → filterbank of a frame
    depends on the samples alone, so each frame is transformed once; the smoothing is a recurrence from frame 0"""

    def _stream_columns(self):
        """audio.add_data() mode: this frame's column from K1 on the new frame(s) + the dynamics scan over all so far"""
        import torch
        audio, scene = self.audio, self.scene
        st = audio.stream
        frames = st["frames"]
        pcm = audio.device_clip(scene.device)
        key = self.config_key()
        done, raw, have = self._stream_raw if self._stream_raw is not None else (0, None, None)
        if raw is None or have != key or done > frames - 1 or raw.shape[2] != pcm.shape[0]:
            done, raw = 0, torch.zeros((256, self.spectrogram_bins, pcm.shape[0]), dtype=torch.float32, device=pcm.device)
        if frames > raw.shape[0]:
            grown = torch.zeros((max(2*raw.shape[0], frames),) + tuple(raw.shape[1:]), dtype=torch.float32, device=pcm.device)
            grown[:done] = raw[:done]
            raw = grown
        if frames > done:
            fresh, _ = self.track(scene.cuda, pcm, st["tell_device"][done:frames].contiguous(), apply_volume=self.apply_volume)
            raw[done:frames] = fresh
        self._stream_raw = (frames, raw, key)
        if not scene.render_enabled and self.length_samples == 1:
            return None
        columns = raw[:frames].clone()
        d = self.dynamics
        scene.cuda.audio_track(pcm, int(audio.samplerate), st["tell_device"], st["dt_device"],
                               spec=columns, bins=self.spectrogram_bins, dynamics=(d.frequency, d.zeta, d.response, d.precision))
        return columns

    def prepare(self) -> None:
        audio, scene = self.audio, self.scene
        if getattr(audio, "clock", None) is None:
            audio.prepare()
        pcm = audio.device_clip(scene.device)
        spec, _ = self.track(scene.cuda, pcm, audio.clock["tell_device"], apply_volume=self.apply_volume)
        d = self.dynamics
        scene.cuda.audio_track(pcm, int(audio.samplerate), audio.clock["tell_device"], audio.clock["dt_device"],
                               spec=spec, bins=self.spectrogram_bins, dynamics=(d.frequency, d.zeta, d.response, d.precision))
        self.columns = spec
        self._prepared_for = (self.config_key(), audio.clock["frames"])

    def update(self):
        tex = self.texture
        want = (self.audio.channels, "linear" if self.smooth else "nearest", self.spectrogram_bins, self.length_samples)
        if (tex.components, tex.filter.value, tex.height, tex.width) != want:
            tex.components, tex.filter, tex.height, tex.width = want
        self.offset = (self.offset + 1) % self.length_samples
        if self.audio.clip is None or self.scene.cuda is None:
            return
        if getattr(self.audio, "streaming", False):
            if getattr(self.audio, "stream", None) is None:
                return                               # the audio module has not seen this frame (not a ShaderAudio)
            columns = self._stream_columns()
            if columns is None:
                return
            self.columns, k = columns, columns.shape[0] - 1
        else:
            if self.columns is None or getattr(self, "_prepared_for", None) != (self.config_key(), self.scene.total_frames):
                self.prepare()
            if not self.scene.render_enabled and self.length_samples == 1:
                return                               # a frame another rank shades (a scrolling texture still needs its column)
            k = min(self.scene.frame_index, self.columns.shape[0] - 1)
        row = self.columns[k]
        if self.length_samples == 1:
            self.texture.bind(self.columns, row.data_ptr())          # the whole texture IS this row
        else:
            self.texture.write(row, viewport=(self.offset, 0, 1, self.spectrogram_bins))   # device → device

    def pipeline(self) -> Iterable[ShaderVariable]:
        yield Uniform("int",   f"{self.name}Length", self.length_samples)
        yield Uniform("int",   f"{self.name}Bins",   self.spectrogram_bins)
        yield Uniform("float", f"{self.name}Offset", self.offset/self.length_samples)
        yield Uniform("int",   f"{self.name}Smooth", self.smooth)
        yield Uniform("float", f"{self.name}Min",    self.spectrogram_frequencies[0])
        yield Uniform("float", f"{self.name}Max",    self.spectrogram_frequencies[-1])
        yield Uniform("bool",  f"{self.name}Scroll", self.scrolling)

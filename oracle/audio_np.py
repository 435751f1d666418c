"""
TEST INFRASTRUCTURE ONLY (oracle, kind "port"). Never imported by the product path; only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs use it.

numpy restatement of the reference's CPU audio path, written against the *whole clip* instead of the
30 s ring so it can be evaluated at any frame index:

    scheduler + scene time        shaderflow/scheduler.py:134-173, shaderflow/scene.py:456-479
    file reader chunk rule        shaderflow/ffmpeg.py:1306-1330, shaderflow/audio/module.py:447-455
    ring / last-n window          shaderflow/audio/module.py:113-141
    windows                       shaderflow/audio/spectrogram.py:90-108
    fft + magnitude               shaderflow/audio/spectrogram.py:20-26,155-171
    scales, interpolators         shaderflow/audio/spectrogram.py:44-87
    filterbank matrix / product   shaderflow/audio/spectrogram.py:175-224
    from_notes                    shaderflow/audio/spectrogram.py:226-245, shaderflow/piano/notes.py
    second order dynamics         shaderflow/dynamics.py:172-250
    volume / std targets          shaderflow/audio/module.py:74-75,457-458
    waveform reducers             shaderflow/audio/waveform.py:14-22,64-87

Parity status: PINNED against the reference's own numpy code, imported in the build container through
`oracle/ref_loader.py` (tests/test_oracle_audio.py) and against the golden vectors that import
produced (tests/golden/audio_*.npz, generator tests/golden/make_golden.py). The reference itself ships
no known-answer test for this path, so there is nothing else to pin to; the north_star's
external pin, scipy.signal.stft within 1e-5, is checked in tests/test_oracle_audio.py too.

dtype rules follow numpy 2 (NEP 50) exactly as the reference experiences them: float32 arrays combined
with Python floats stay float32, 0-d float64 arrays combined with float32 give float64.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

F32 = np.float32

# ---------------------------------------------------------------------------------------------- #
# Time base: what every module sees on frame k of a freewheel export

def frame_clock(n_frames: int, fps: float = 60.0, speed: float = 1.0):
    """
    Returns (time[k], dt[k], rdt[k]) as seen by modules *during* frame k.

    scheduler.py:86-89 (freewheel: started=0, last_call=-period, next_call=0), :152-168
    (now=next_call, dt=now-last_call, next_call+=period) and scene.py:476-479 (time integrates at
    the END of next(), so frame 0 sees time=0, dt=0, rdt=0).
    """
    period = 1.0/fps
    last_call, next_call = (0.0 - period), 0.0
    time = np.zeros(n_frames, np.float64)
    dt   = np.zeros(n_frames, np.float64)
    rdt  = np.zeros(n_frames, np.float64)
    t = 0.0; sdt = 0.0; srdt = 0.0
    for k in range(n_frames):
        time[k], dt[k], rdt[k] = t, sdt, srdt
        now = next_call
        passed = now - last_call
        last_call = now
        while next_call <= now:
            next_call += period
        sdt  = passed*speed
        srdt = passed
        t   += sdt
    return time, dt, rdt


def reader_tell(rdt: np.ndarray, samplerate: int, channels: int, total: int | None = None,
                bytes_per_sample: int = 4) -> np.ndarray:
    """
    Samples consumed from the file AFTER frame k's ShaderAudio.update (== BrokenAudio.tell).
    ffmpeg.py:1306-1330: target += chunk; length = block*round((target - read/bps)*bps/block),
    at least one block. Frame 0 (rdt=0) therefore reads exactly one sample.
    """
    block = bytes_per_sample*channels
    bps   = block*samplerate
    target, read = 0.0, 0
    limit = None if total is None else total*block
    tell = np.zeros(len(rdt), np.int64)
    for k, chunk in enumerate(rdt):
        if limit is None or read < limit:
            target += float(chunk)
            length = (target - (read/bps))*bps
            length = int(block*round(length/block))
            length = max(length, block)
            if limit is not None:
                length = min(length, limit - read)
            read += length
        tell[k] = read//block
    return tell

# ---------------------------------------------------------------------------------------------- #
# Ring semantics on the whole clip

def last_n(x: np.ndarray, tell: int, n: int, offset: int = 0) -> np.ndarray:
    """
    audio/module.py:137-138 on a ring that holds the first `tell` samples of `x` right-aligned over
    zeros: data[:, -(n+offset+1) : -(offset+1)] == clip samples [tell-n-offset-1, tell-offset-1),
    zero where the index is negative. The newest sample is excluded (reference quirk, kept).
    """
    n, offset = int(n), int(offset)
    hi = int(tell) - offset - 1
    lo = hi - n
    out = np.zeros((x.shape[0], n), x.dtype)
    a, b = max(lo, 0), max(hi, 0)
    if b > a:
        out[:, a - lo:b - lo] = x[:, a:b]
    return out

# ---------------------------------------------------------------------------------------------- #
# Spectrogram

WINDOW_HANNING, WINDOW_HANN_POISSON, WINDOW_NONE = 0, 1, 2
MAGNITUDE_POWER, MAGNITUDE_AMPLITUDE = 0, 1
VOLUME_LINEAR, VOLUME_SQRT, VOLUME_DBFS, VOLUME_DBFS_TREMX = 0, 1, 2, 3
SCALE_OCTAVE, SCALE_MEL = 0, 1
INTERP_EULER, INTERP_DIRAC, INTERP_SINC = 0, 1, 2


def window(kind: int, size: int) -> np.ndarray:
    """spectrogram.py:90-108, float64"""
    if kind == WINDOW_HANNING:
        return np.hanning(size)
    if kind == WINDOW_HANN_POISSON:
        n = np.arange(size)
        return (0.5*(1 - np.cos(2*np.pi*n/size))) * np.exp(-2.0*np.abs(size - 2*n)/size)
    return np.ones(size)


def fft_magnitude(frame: np.ndarray, window_kind: int = WINDOW_HANNING,
                  magnitude: int = MAGNITUDE_POWER) -> np.ndarray:
    """spectrogram.py:169-171: float64 rfft of window*data, Power=re²+im², cast to float32"""
    z = np.fft.rfft(window(window_kind, frame.shape[-1]) * frame)
    if magnitude == MAGNITUDE_POWER:
        m = (z*z.conjugate()).real
    else:
        m = np.abs(z)
    return m.astype(F32)


def volume(kind: int, x: np.ndarray) -> np.ndarray:
    """spectrogram.py:28-41 — dead code in the reference's next() (:176 returns first), opt-in here"""
    if kind == VOLUME_SQRT:       return np.sqrt(x)
    if kind == VOLUME_DBFS:       return 10*np.log10(x)
    if kind == VOLUME_DBFS_TREMX: return 10*(np.log10(x + 0.1) + 1)/1.0414
    return x


def note_frequency(index: int, tuning: float = 440.0) -> float:
    """piano/notes.py:56-59"""
    return tuning * 2**((index - 69)/12)


def note_from_frequency(frequency: float, tuning: float = 440.0) -> int:
    """piano/notes.py:72-75"""
    return round(12*math.log2(frequency/tuning) + 69)


@dataclass
class BankConfig:
    """The attributes of BrokenSpectrogram that shape the filterbank (spectrogram.py:112-184)"""
    fft_n: int = 12
    samplerate: float = 44100
    minimum_frequency: float = 20.0
    maximum_frequency: float = 20000.0
    bins: int = 1000
    scale: int = SCALE_OCTAVE
    interpolation: int = INTERP_EULER
    euler_end: float = 1.2

    @classmethod
    def from_notes(cls, start: int, end: int, bins: int = 1000, piano: bool = False,
                   tuning: float = 440.0, **kw) -> "BankConfig":
        """spectrogram.py:226-245 (start/end are midi note indices)"""
        fmin, fmax = note_frequency(start, tuning), note_frequency(end, tuning)
        if piano:
            half = 2**(0.5/12)
            bins = (end - start) + 1
            fmin /= half
            fmax *= half
        return cls(minimum_frequency=fmin, maximum_frequency=fmax, bins=bins, **kw)

    @property
    def fft_size(self) -> int:
        return int(2**self.fft_n)

    @property
    def fft_bins(self) -> int:
        return int(self.fft_size/2 + 1)


def spectrogram_frequencies(cfg: BankConfig) -> np.ndarray:
    """spectrogram.py:73-87,186-192"""
    if cfg.scale == SCALE_OCTAVE:
        fwd = lambda x: np.log(x)/np.log(2)
        inv = lambda x: 2**x
    else:
        fwd = lambda x: 2595*np.log10(1 + x/700)
        inv = lambda x: 700*(10**(x/2595) - 1)
    return inv(np.linspace(fwd(cfg.minimum_frequency), fwd(cfg.maximum_frequency), cfg.bins))


def interpolate(cfg: BankConfig, x: np.ndarray) -> np.ndarray:
    """spectrogram.py:44-70"""
    if cfg.interpolation == INTERP_EULER:
        end = cfg.euler_end
        return np.exp(-(2*x/end)**2) / (end*(math.pi**0.5))
    if cfg.interpolation == INTERP_DIRAC:
        d = np.zeros(x.shape); d[np.round(x) == 0] = 1
        return d
    return np.abs(np.sinc(x))


def filterbank_matrix(cfg: BankConfig) -> np.ndarray:
    """spectrogram.py:194-216: dense (bins, fft_bins) float32, |m|<1e-5 zeroed"""
    df = np.fft.rfftfreq(cfg.fft_size, 1/cfg.samplerate)[1]
    cols = np.arange(cfg.fft_bins)
    m = np.array([interpolate(cfg, idx - cols) for idx in (spectrogram_frequencies(cfg)/df)], dtype=F32)
    m[np.abs(m) < 1e-5] = 0
    return m


def filterbank_csr(m: np.ndarray):
    """CSR (indptr, indices, data) of the dense matrix, row-major ascending columns — the layout
    scipy.sparse.csr_matrix(m) has (spectrogram.py:218-220) and the one the C-ABI takes"""
    rows, cols = np.nonzero(m)
    indptr = np.zeros(m.shape[0] + 1, np.int32)
    np.add.at(indptr, rows + 1, 1)
    return np.cumsum(indptr).astype(np.int32), cols.astype(np.int32), m[rows, cols].astype(F32)


def filterbank_apply(csr, mag: np.ndarray) -> np.ndarray:
    """spectrogram.py:176: (M · magᵀ)ᵀ, float32 accumulation in ascending column order per row
    (what scipy's csr_matvecs does). mag (ch, fft_bins) → (ch, bins)"""
    indptr, idx, val = csr
    bins = len(indptr) - 1
    out = np.zeros((mag.shape[0], bins), F32)
    width = int(np.max(np.diff(indptr))) if bins else 0
    for t in range(width):
        pos = indptr[:-1] + t
        ok = pos < indptr[1:]
        p = np.where(ok, pos, 0)
        term = (val[p][None, :] * mag[:, idx[p]]).astype(F32)
        out = np.where(ok[None, :], (out + term).astype(F32), out)
    return out

# ---------------------------------------------------------------------------------------------- #
# Dynamics

@dataclass
class Dynamics:
    """dynamics.py:90-250 (DynamicNumber). `value` keeps whatever dtype it was created with"""
    frequency: float = 1.0
    zeta: float = 1.0
    response: float = 0.0
    precision: float = 1e-6
    integrate: bool = False
    value: np.ndarray = field(default_factory=lambda: np.array(0, np.float64))
    target: np.ndarray = None
    previous: np.ndarray = None
    derivative: np.ndarray = None
    integral: np.ndarray = None
    acceleration: np.ndarray = None

    def __post_init__(self):
        self.set(self.value)

    def set(self, value):
        """dynamics.py:126-136 (instant=True)"""
        value = np.array(value)
        self.value, self.target, self.previous = value.copy(), value.copy(), value.copy()
        self.integral, self.derivative, self.acceleration = (np.zeros_like(value) for _ in range(3))

    # dynamics.py:172-195
    @property
    def radians(self): return math.tau*self.frequency
    @property
    def k1(self): return self.zeta/(math.pi*self.frequency)
    @property
    def k2(self): return 1.0/(self.radians*self.radians)
    @property
    def k3(self): return (self.response*self.zeta)/(math.tau*self.frequency)
    @property
    def damping(self): return self.radians*(abs(self.zeta*self.zeta - 1.0))**0.5

    def coefficients(self, dt: float) -> tuple[float, float]:
        """dynamics.py:231-242: clamp branch or pole matching; returns (k1, k2) for this dt"""
        if self.radians*dt < self.zeta:
            k1 = self.k1
            k2 = max(k1*dt, self.k2, 0.5*(k1 + dt)*dt)
        else:
            t1 = math.exp(-1*self.zeta*self.radians*dt)
            a1 = 2*t1*(math.cos if self.zeta <= 1 else math.cosh)(self.damping*dt)
            t2 = 1/(1 + t1*t1 - a1)*dt
            k1 = t2*(1 - t1*t1)
            k2 = t2*dt
        return k1, k2

    def next(self, target=None, dt: float = 1.0):
        """dynamics.py:197-250"""
        if not dt:
            return self.value
        if target is not None:
            self.target = target if isinstance(target, np.ndarray) else \
                np.array(target, dtype=getattr(target, "dtype", np.float64))
            if self.target.shape != self.value.shape:
                self.set(target)
        if np.abs(self.target - self.value).max() < self.precision:
            if self.integrate:
                self.integral = self.integral + self.value*dt
            return self.value
        velocity = (self.target - self.previous)/dt
        self.previous = self.target
        k1, k2 = self.coefficients(dt)
        self.value        = self.value + self.derivative*dt
        self.acceleration = (self.target + self.k3*velocity - self.value - k1*self.derivative)/k2
        self.derivative   = self.derivative + self.acceleration*dt
        if self.integrate:
            self.integral = self.integral + self.value*dt
        return self.value

# ---------------------------------------------------------------------------------------------- #
# Volume, std, waveform

def volume_std_targets(x: np.ndarray, tell: int, samplerate: float):
    """audio/module.py:457-458 with get_last_n_seconds(0.1) (:140-141)"""
    data = last_n(x, tell, 0.1*samplerate)
    vol = 2 * np.sqrt(np.mean(np.square(data))) * (2**0.5)
    std = np.std(data)
    return vol, std


REDUCER_AVERAGE, REDUCER_RMS, REDUCER_STD = 0, 1, 2

def waveform_row(x: np.ndarray, tell: int, samplerate: float, length: float = 3, rate: float = 60,
                 reducer: int = REDUCER_AVERAGE) -> np.ndarray:
    """waveform.py:64-87 → (points, channels) float32"""
    points = int(length*rate)
    chunk = max(1, int(length*samplerate/points))
    offset = int(tell) % chunk
    data = last_n(x, tell, chunk*points, offset=offset)
    c = data.reshape(x.shape[0], -1, chunk)
    if reducer == REDUCER_AVERAGE:
        r = np.sqrt(np.mean(np.abs(c), axis=2))
    elif reducer == REDUCER_RMS:
        r = np.sqrt(np.sqrt(np.mean(c**2, axis=2))*(2**0.5))
    else:
        r = np.sqrt(np.std(c, axis=2))
    return np.ascontiguousarray(r.T).astype(F32)

# ---------------------------------------------------------------------------------------------- #
# Whole track: what ShaderAudio + ShaderSpectrogram + ShaderWaveform produce for every frame

@dataclass
class TrackConfig:
    samplerate: int = 44100
    fps: float = 60.0
    speed: float = 1.0
    window: int = WINDOW_HANNING
    magnitude: int = MAGNITUDE_POWER
    volume: int = VOLUME_LINEAR
    bank: BankConfig = field(default_factory=BankConfig)
    spec_frequency: float = 4.0       # spectrogram.py:287-290
    spec_zeta: float = 1.0
    spec_response: float = 0.0
    wave_length: float = 3
    wave_rate: float = 60
    wave_reducer: int = REDUCER_AVERAGE


def audio_track(x: np.ndarray, n_frames: int, cfg: TrackConfig, *, keep_magnitude: bool = False,
                waveform: bool = True, scalars: bool = True, spectrogram: bool = True) -> dict:
    """
    Runs the reference's per-frame audio pipeline for frames 0..n_frames-1 of a freewheel export of
    clip `x` (channels, samples), float32. Returns arrays indexed by frame:
        tell      (F,)            samples consumed after the frame's update
        time, dt  (F,)            scene.time / scene.dt seen by modules
        mag       (F, ch, fft_bins) f32   [keep_magnitude]
        spec      (F, ch, bins)   f32     BrokenSpectrogram.next()
        column    (F, bins, ch)   f32     smoothed texture column, texel b = (L_b, R_b)
        vol_target, std_target (F,) f32 ; volume, volume_integral, std (F,) f64
        wave      (F, points, ch) f32
    """
    x = np.ascontiguousarray(x, dtype=F32)
    ch = x.shape[0]
    time, dt, rdt = frame_clock(n_frames, cfg.fps, cfg.speed)
    tell = reader_tell(rdt, cfg.samplerate, ch, total=x.shape[1])
    out = dict(tell=tell, time=time, dt=dt)

    bank = cfg.bank
    if spectrogram:
        csr = filterbank_csr(filterbank_matrix(bank))
        dyn = Dynamics(frequency=cfg.spec_frequency, zeta=cfg.spec_zeta, response=cfg.spec_response,
                       value=np.zeros((ch, bank.bins), F32))
        out["spec"]   = np.zeros((n_frames, ch, bank.bins), F32)
        out["column"] = np.zeros((n_frames, bank.bins, ch), F32)
        if keep_magnitude:
            out["mag"] = np.zeros((n_frames, ch, bank.fft_bins), F32)
    if scalars:
        vol = Dynamics(frequency=2, zeta=1, response=0, integrate=True)   # audio/module.py:413-417
        std = Dynamics(frequency=10, zeta=1, response=0)                  # audio/module.py:418-421
        for key in ("vol_target", "std_target"):
            out[key] = np.zeros(n_frames, F32)
        for key in ("volume", "volume_integral", "std"):
            out[key] = np.zeros(n_frames, np.float64)
    if waveform:
        points = int(cfg.wave_length*cfg.wave_rate)
        out["wave"] = np.zeros((n_frames, points, ch), F32)

    for k in range(n_frames):
        step = abs(float(dt[k]))
        if scalars:
            vt, st = volume_std_targets(x, tell[k], cfg.samplerate)
            out["vol_target"][k], out["std_target"][k] = vt, st
            vol.target = np.array(vt, dtype=F32)   # _ensure_numpy keeps the float32 of np scalars
            std.target = np.array(st, dtype=F32)
            vol.next(dt=step); std.next(dt=step)
            out["volume"][k], out["volume_integral"][k], out["std"][k] = vol.value, vol.integral, std.value
        if waveform:
            out["wave"][k] = waveform_row(x, tell[k], cfg.samplerate, cfg.wave_length,
                                          cfg.wave_rate, cfg.wave_reducer)
        if spectrogram:
            mag = fft_magnitude(last_n(x, tell[k], bank.fft_size), cfg.window, cfg.magnitude)
            spec = volume(cfg.volume, filterbank_apply(csr, mag)).astype(F32)
            # spectrogram.py:306: (ch,bins).T.reshape(2,-1) — memory reinterpretation of (bins,ch)
            dyn.target = np.ascontiguousarray(spec.T).reshape(ch, -1)
            dyn.next(dt=step)
            if keep_magnitude:
                out["mag"][k] = mag
            out["spec"][k] = spec
            out["column"][k] = dyn.value.astype(F32).reshape(bank.bins, ch)
    return out

# ---------------------------------------------------------------------------------------------- #
# Synthetic inputs of BASELINE.json's configs

def synth_sine(seconds: float = 1.0, hz: float = 440.0, sr: int = 44100) -> np.ndarray:
    t = np.arange(int(seconds*sr))/sr
    s = np.sin(2*np.pi*hz*t).astype(F32)
    return np.stack([s, s])


def synth_noise(seconds: float = 10.0, sr: int = 44100, seed: int = 0) -> np.ndarray:
    return np.random.default_rng(seed).uniform(-1, 1, (2, int(sr*seconds))).astype(F32)


def synth_chirp(seconds: float = 60.0, sr: int = 44100, f0: float = 20.0, f1: float = 20000.0) -> np.ndarray:
    """L = logarithmic chirp f0→f1 (scipy.signal.chirp's formula), R = time-reversed L"""
    t = np.arange(int(seconds*sr), dtype=np.float64)/sr
    beta = seconds/np.log(f1/f0)
    left = np.cos(2*np.pi*beta*f0*(np.power(f1/f0, t/seconds) - 1.0)).astype(F32)
    return np.stack([left, left[::-1].copy()])

"""GPU parity of kernels K1/K2 (through the C ABI) against the oracle and the reference's goldens."""
import numpy as np
import pytest

from oracle import audio_np as A
from tests.helpers import piano_config

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def ctx():
    from shaderflow_b200 import _native as N
    c = N.Context(0)
    yield c
    c.destroy()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def run_track(ctx, x, frames, cfg: A.TrackConfig, total=None, want_mag=True):
    from shaderflow_b200 import _native as N
    ch, n = x.shape
    time, dt, tell = N.frame_clock(frames, cfg.fps, cfg.speed, cfg.samplerate, ch, n if total is None else total)
    indptr, idx, val = A.filterbank_csr(A.filterbank_matrix(cfg.bank))
    pcm, tell_d, dt_d = dev(x), dev(tell), dev(dt)
    csr = (dev(indptr), dev(idx), dev(val), cfg.bank.bins)
    mag = torch.zeros((frames, ch, cfg.bank.fft_bins), dtype=torch.float32, device="cuda") if want_mag else None
    spec = torch.zeros((frames, cfg.bank.bins, ch), dtype=torch.float32, device="cuda")
    ctx.stft_mel(pcm, tell_d, cfg.bank.fft_n, csr, window=cfg.window, magnitude=cfg.magnitude,
                 volume=cfg.volume, mag_out=mag, spec_out=spec)
    raw = spec.clone()
    scalars = torch.zeros((frames, N.SCALARS), dtype=torch.float64, device="cuda")
    points = int(cfg.wave_length*cfg.wave_rate)
    chunk = max(1, int(cfg.wave_length*cfg.samplerate/points))
    wave = torch.zeros((frames, points, ch), dtype=torch.float32, device="cuda")
    ctx.audio_track(pcm, cfg.samplerate, tell_d, dt_d, spec=spec, bins=cfg.bank.bins,
                    dynamics=(cfg.spec_frequency, cfg.spec_zeta, cfg.spec_response, 1e-6),
                    scalars=scalars, wave=wave, wave_points=points, wave_chunk=chunk, wave_reducer=cfg.wave_reducer)
    ctx.sync()
    return dict(tell=tell, mag=None if mag is None else mag.cpu().numpy(), spec=raw.cpu().numpy().transpose(0, 2, 1),
                column=spec.cpu().numpy(), scalars=scalars.cpu().numpy(), wave=wave.cpu().numpy())


def rel_peak(a, b):
    peak = max(np.abs(b).max(), 1e-30)
    return np.abs(a.astype(np.float64) - b.astype(np.float64)).max()/peak


CASES = {
    "audio_c1_sine":    (lambda: A.synth_sine(1.0),           60, 60.0, (15, 129)),
    "audio_noise":      (lambda: A.synth_noise(1.0, seed=0),  60, 60.0, (15, 129)),
    "audio_chirp_1000": (lambda: A.synth_chirp(1.5),          30, 24.0, None),
    "audio_short":      (lambda: A.synth_noise(0.25, seed=3), 30, 60.0, (15, 129)),
}


@pytest.mark.parametrize("name", list(CASES))
def test_track_against_reference_golden(ctx, name, golden_dir):
    """Fixtures were produced by the reference's own numpy code (tests/golden/make_golden.py)"""
    from shaderflow_b200 import _native as N
    make, frames, fps, notes = CASES[name]
    gold = np.load(golden_dir/f"{name}.npz")
    got = run_track(ctx, make(), frames, piano_config(fps, notes))
    assert np.array_equal(got["tell"], gold["tell"])
    # float32 FFT vs the reference's float64 pocketfft: 1e-5 of the peak (north_star tolerance)
    for row, k in enumerate(gold["mag_frames"]):
        assert rel_peak(got["mag"][k], gold["mag"][row]) < 1e-5, k
    assert rel_peak(got["spec"], gold["spec"]) < 1e-5
    assert rel_peak(got["column"], gold["column"]) < 1e-5
    s = got["scalars"]
    assert np.allclose(s[:, N.SCALAR_VOLUME_TARGET], gold["vol_target"], rtol=2e-6, atol=1e-7)
    assert np.allclose(s[:, N.SCALAR_STD_TARGET], gold["std_target"], rtol=2e-6, atol=1e-7)
    assert np.allclose(s[:, N.SCALAR_VOLUME], gold["volume"], rtol=1e-5, atol=1e-7)
    assert np.allclose(s[:, N.SCALAR_VOLUME_INTEGRAL], gold["volume_integral"], rtol=1e-5, atol=1e-7)
    assert np.allclose(s[:, N.SCALAR_STD], gold["std"], rtol=1e-5, atol=1e-7)
    assert np.allclose(got["wave"], gold["wave"], rtol=2e-6, atol=1e-7)


def test_stft_against_scipy(ctx):
    """north_star: the spectrogram matches scipy.signal.stft within 1e-5 (SURVEY App. A.2 identity)"""
    import scipy.signal
    x = A.synth_chirp(2.0)
    n, hop, frames = 4096, 735, 100
    got = run_track(ctx, x, frames, piano_config())
    w = np.hanning(n)
    padded = np.concatenate([np.zeros((2, n + 1)), x.astype(np.float64)], axis=1)
    _, _, Z = scipy.signal.stft(padded, window=w, nperseg=n, noverlap=n - hop, boundary=None,
                                padded=False, detrend=False, scaling="spectrum")
    ref = (np.abs(Z)*w.sum())**2
    for k in (1, 2, 3, 17, 50, 99):
        assert rel_peak(got["mag"][k], ref[:, :, k]) < 1e-5, k


@pytest.mark.parametrize("fft_n", [8, 9, 10, 11, 13])
def test_other_fft_sizes(ctx, fft_n):
    x = A.synth_noise(0.5, seed=fft_n)
    cfg = A.TrackConfig(bank=A.BankConfig(fft_n=fft_n, bins=64, minimum_frequency=200, maximum_frequency=8000))
    got = run_track(ctx, x, 20, cfg)
    ref = A.audio_track(x, 20, cfg, keep_magnitude=True, waveform=False, scalars=False)
    assert rel_peak(got["mag"], ref["mag"]) < 1e-5
    assert rel_peak(got["spec"], ref["spec"]) < 1e-5


@pytest.mark.parametrize("window,magnitude,volume", [(1, 0, 0), (2, 1, 0), (0, 1, 1), (0, 0, 3)])
def test_window_magnitude_volume_variants(ctx, window, magnitude, volume):
    x = A.synth_noise(0.5, seed=5)
    cfg = piano_config(); cfg.window, cfg.magnitude, cfg.volume = window, magnitude, volume
    got = run_track(ctx, x, 12, cfg)
    ref = A.audio_track(x, 12, cfg, keep_magnitude=True, waveform=False, scalars=False)
    assert rel_peak(got["mag"], ref["mag"]) < 1e-5
    assert rel_peak(got["spec"][1:], ref["spec"][1:]) < 2e-5


def test_mono_and_unaligned_clip(ctx):
    """channels=1, and a stereo clip whose second channel is not 16-byte aligned (odd length)"""
    x = A.synth_noise(0.4, seed=9)[:, :17639]
    cfg = piano_config()
    ref = A.audio_track(x, 16, cfg, keep_magnitude=True, waveform=False, scalars=False)
    got = run_track(ctx, x, 16, cfg)
    assert rel_peak(got["mag"], ref["mag"]) < 1e-5
    mono = x[:1]
    got = run_track(ctx, mono, 16, cfg)
    m = A.fft_magnitude(A.last_n(mono, got["tell"][9], 4096))
    assert rel_peak(got["mag"][9], m) < 1e-5


def test_dynamics_scan_is_bit_exact(ctx):
    """Feeding the oracle's own spectrogram rows through the scan must reproduce numpy's float32
    recurrence bit for bit (no FMA contraction, same operation order)"""
    x = A.synth_noise(1.0, seed=2)
    cfg = piano_config()
    ref = A.audio_track(x, 60, cfg, waveform=False, scalars=False)
    from shaderflow_b200 import _native as N
    _, dt, tell = N.frame_clock(60, 60.0, 1.0, 44100, 2, x.shape[1])
    spec = dev(np.ascontiguousarray(ref["spec"].transpose(0, 2, 1)))
    ctx.audio_track(dev(x), 44100, dev(tell), dev(dt), spec=spec, bins=115)
    ctx.sync()
    assert np.array_equal(spec.cpu().numpy(), ref["column"])


def test_dynamics_early_out_and_silence(ctx):
    """All-zero audio never leaves the precision band: the scan must hold state exactly (early-out
    branch, dynamics.py:222-225), and an empty batch is a no-op"""
    x = np.zeros((2, 30000), np.float32)
    got = run_track(ctx, x, 30, piano_config())
    assert not got["column"].any() and not got["scalars"][:, :3].any() and not got["wave"].any()
    from shaderflow_b200 import _native as N
    empty = torch.zeros((0,), dtype=torch.int64, device="cuda")
    ctx.stft_mel(dev(x), empty, 12, None, mag_out=torch.zeros(1, device="cuda"))


def test_linearity_at_full_size(ctx):
    """BASELINE configs[2] size (60 s, 3600 frames): amplitude spectra are linear in the input and the
    GPU track equals the oracle on sampled frames"""
    x = A.synth_chirp(60.0)
    cfg = piano_config(); cfg.magnitude = A.MAGNITUDE_AMPLITUDE
    a = run_track(ctx, x, 3600, cfg)
    b = run_track(ctx, (2*x).astype(np.float32), 3600, cfg)
    assert rel_peak(b["mag"], 2*a["mag"]) < 1e-6
    cfg.magnitude = A.MAGNITUDE_POWER
    got = run_track(ctx, x, 3600, cfg, want_mag=True)
    for k in (0, 1, 599, 1800, 3599):
        m = A.fft_magnitude(A.last_n(x, got["tell"][k], 4096))
        assert rel_peak(got["mag"][k], m) < 1e-5, k

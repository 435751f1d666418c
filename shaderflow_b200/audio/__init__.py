from .module import BrokenAudio, ShaderAudio  # noqa: F401
from .spectrogram import BrokenSpectrogram, ShaderSpectrogram  # noqa: F401
from .waveform import ShaderWaveform, WaveformReducer  # noqa: F401

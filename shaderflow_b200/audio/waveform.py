"""ShaderWaveform (API mirror of shaderflow/audio/waveform.py): the last `length` seconds of audio
reduced to `length*samplerate` bars per channel, as an RG32F row texture. The reducer runs on the GPU
once per absolute chunk for the whole clip (csrc/audio.cu, waveform_chunks_kernel) — consecutive frames
share all but one chunk — and each frame's texture is a zero-copy view of its row."""
from __future__ import annotations

import math
from enum import Enum
from typing import Any, Iterable

import numpy as np
from attrs import define, field

from shaderflow_b200 import _native as N
from shaderflow_b200.audio.module import BrokenAudio
from shaderflow_b200.module import ShaderModule
from shaderflow_b200.texture import ShaderTexture
from shaderflow_b200.variable import ShaderVariable, Uniform


class WaveformReducer(Enum):
    """(channels, chunks, samples) → (channels, chunks); `.kind` selects the kernel's reducer"""
    def Average(x): return np.sqrt(np.mean(np.abs(x), axis=2))
    def RMS(x): return np.sqrt(np.sqrt(np.mean(x**2, axis=2))*(2**0.5))
    def STD(x): return np.sqrt(np.std(x, axis=2))
    Average.kind, RMS.kind, STD.kind = N.REDUCER["average"], N.REDUCER["rms"], N.REDUCER["std"]


@define
class ShaderWaveform(ShaderModule):
    name: str = "iWaveform"
    audio: BrokenAudio = None
    length: float = 3
    samplerate: float = 60
    reducer: Any = WaveformReducer.Average
    smooth: bool = True
    texture: ShaderTexture = None
    rows: Any = field(default=None, repr=False)
    """[frames][points][channels] float32 device track of the current export"""

    @property
    def length_samples(self) -> int:
        return int(max(1, self.length*self.scene.fps))

    def build(self):
        self.texture = ShaderTexture(scene=self.scene, filter=("linear" if self.smooth else "nearest"),
            components=self.audio.channels, name=self.name, width=self._points, height=1,
            mipmaps=False, dtype=np.float32).repeat(False)

    @property
    def chunk_size(self) -> int:
        return max(1, int(self.length*self.audio.samplerate/self._points))

    @property
    def _points(self) -> int:
        return int(self.length*self.samplerate)

    @property
    def _offset(self) -> int:
        return self.audio.tell % self.chunk_size

    @property
    def _cutoff(self) -> int:
        return int(self.chunk_size*math.floor(self.audio.buffer_size/self.chunk_size))

    def setup(self):
        self.rows = None

    def prepare(self) -> None:
        import torch
        audio, scene = self.audio, self.scene
        if getattr(audio, "clock", None) is None:
            audio.prepare()
        pcm = audio.device_clip(scene.device)
        frames = audio.clock["frames"]
        rows = torch.zeros((frames, self._points, audio.channels), dtype=torch.float32, device=pcm.device)
        scene.cuda.audio_track(pcm, int(audio.samplerate), audio.clock["tell_device"], audio.clock["dt_device"],
                               wave=rows, wave_points=self._points, wave_chunk=self.chunk_size,
                               wave_reducer=getattr(self.reducer, "kind", 0))
        self.rows = rows

    def _stream_row(self):
        """audio.add_data() mode: the reducer row of this frame alone (it depends on the samples, not on other frames)"""
        import torch
        audio, scene = self.audio, self.scene
        st = audio.stream
        pcm = audio.device_clip(scene.device)
        last = st["frames"] - 1
        row = torch.zeros((1, self._points, audio.channels), dtype=torch.float32, device=pcm.device)
        scene.cuda.audio_track(pcm, int(audio.samplerate), st["tell_device"][last:].contiguous(), st["dt_device"][last:].contiguous(),
                               wave=row, wave_points=self._points, wave_chunk=self.chunk_size,
                               wave_reducer=getattr(self.reducer, "kind", 0))
        return row

    def update(self):
        if self.audio.clip is None or self.scene.cuda is None:
            return
        if getattr(self.audio, "streaming", False):
            if self.texture.components != self.audio.channels:
                self.texture.components = self.audio.channels
            if self.scene.render_enabled and getattr(self.audio, "stream", None) is not None:
                self.rows = self._stream_row()
                self.texture.bind(self.rows, self.rows.data_ptr())
            return
        if self.rows is None or self.rows.shape[0] != self.scene.total_frames or self.rows.shape[1] != self._points:
            self.prepare()
        if self.texture.components != self.audio.channels:
            self.texture.components = self.audio.channels
        if not self.scene.render_enabled:
            return                                   # a frame another rank shades
        k = min(self.scene.frame_index, self.rows.shape[0] - 1)
        self.texture.bind(self.rows, self.rows[k].data_ptr())

    def pipeline(self) -> Iterable[ShaderVariable]:
        yield Uniform("int", f"{self.name}Length", self.length_samples)

"""Audio file ingestion without ffmpeg.

The reference decodes every file through an ffmpeg child asked for interleaved float32 PCM and consumes it in
chunks sized by `BrokenAudioReader.stream`'s drift-free rule (ffmpeg.py:1240-1333). In this backend the chunk
rule is `sfb_frame_clock` (how many samples each frame has consumed — including non-integer hops and a file that
ends mid-export), and the clip is resident in HBM; what is left of "reading the file" is getting its samples
there. RIFF/WAVE (and RF64) files are parsed here and need no decoder at all:

    host     the 'data' chunk is memory-mapped; nothing is converted on the CPU for the GPU path
    H2D      the file's own sample bytes, through two pinned staging buffers (a 16-bit file moves half the bytes of
             the float32 stream the reference pipes)
    device   `sfb_pcm_ingest` converts with libswresample's rules and transposes into the planar float32 clip

`read_wav` is the host-side equivalent (same arithmetic in numpy) for the API surface that hands out numpy
arrays (`BrokenAudio.clip`, `get_last_n_samples`). Compressed formats still go through ffmpeg when a binary
exists (audio/module.py: read_audio_file). FLAC streams are decoded natively on the host (`read_flac`, csrc/flac.cu —
a serial bit stream is no work for a GPU) with every frame CRC and the stream's MD5 checked."""
from __future__ import annotations

import struct
from dataclasses import dataclass
from pathlib import Path

import numpy as np

from shaderflow_b200 import _native as N

WAVE_FORMAT_PCM, WAVE_FORMAT_IEEE_FLOAT, WAVE_FORMAT_EXTENSIBLE = 0x0001, 0x0003, 0xFFFE


@dataclass
class WavInfo:
    path: Path
    samplerate: int
    channels: int
    format: int            # N.PCM_*
    sample_bytes: int
    data_offset: int       # byte offset of the first sample frame in the file
    frames: int            # sample frames in the 'data' chunk

    @property
    def block(self) -> int:
        return self.sample_bytes*self.channels

    @property
    def duration(self) -> float:
        return self.frames/self.samplerate


def parse_wav(path) -> WavInfo:
    """RIFF / RF64 chunk walk → where the samples are and what they are. Raises ValueError on anything that is
    not little-endian integer or IEEE-float PCM."""
    path = Path(path)
    size = path.stat().st_size
    with open(path, "rb") as f:
        head = f.read(12)
        if len(head) < 12 or head[:4] not in (b"RIFF", b"RF64") or head[8:12] != b"WAVE":
            raise ValueError(f"'{path}' is not a RIFF/WAVE file")
        rf64 = head[:4] == b"RF64"
        fmt, data, data_size64 = None, None, None
        while True:
            header = f.read(8)
            if len(header) < 8:
                break
            tag, length = header[:4], struct.unpack("<I", header[4:])[0]
            start = f.tell()
            if tag == b"ds64":
                body = f.read(min(length, 28))
                data_size64 = struct.unpack("<Q", body[8:16])[0]
            elif tag == b"fmt ":
                fmt = f.read(min(length, 40))
            elif tag == b"data":
                if rf64 and length == 0xFFFFFFFF and data_size64 is not None:
                    length = data_size64
                data = (start, min(length, size - start))       # a truncated file plays what is there
                break
            f.seek(start + length + (length & 1))               # chunks are word aligned
    if fmt is None or data is None or len(fmt) < 16:
        raise ValueError(f"'{path}': missing 'fmt ' or 'data' chunk")
    tag, channels, rate, _, block, bits = struct.unpack("<HHIIHH", fmt[:16])
    if tag == WAVE_FORMAT_EXTENSIBLE and len(fmt) >= 26:
        tag = struct.unpack("<H", fmt[24:26])[0]                # first two bytes of the SubFormat GUID
    width = (bits + 7)//8
    if channels < 1 or block != width*channels:
        raise ValueError(f"'{path}': inconsistent block alignment")
    if tag == WAVE_FORMAT_PCM and width in (1, 2, 3, 4):
        kind = {1: N.PCM_U8, 2: N.PCM_S16, 3: N.PCM_S24, 4: N.PCM_S32}[width]
    elif tag == WAVE_FORMAT_IEEE_FLOAT and width in (4, 8):
        kind = N.PCM_F32 if width == 4 else N.PCM_F64
    else:
        raise ValueError(f"'{path}': unsupported WAVE format tag {tag:#x} with {bits} bits per sample")
    return WavInfo(path=path, samplerate=int(rate), channels=int(channels), format=kind, sample_bytes=width,
                   data_offset=data[0], frames=data[1]//block)


def decode_pcm(raw: np.ndarray, fmt: int, channels: int) -> np.ndarray:
    """Interleaved little-endian sample bytes → (frames, channels) float32 with libswresample's conversions
    (the arithmetic of csrc/ingest.cu, in numpy)"""
    raw = np.ascontiguousarray(raw, dtype=np.uint8)
    if fmt == N.PCM_U8:
        x = (raw.astype(np.float32) - np.float32(128))*np.float32(1/128)
    elif fmt == N.PCM_S16:
        x = raw.view("<i2").astype(np.float32)*np.float32(1/32768)
    elif fmt == N.PCM_S24:
        b = raw.reshape(-1, 3).astype(np.uint32)
        v = ((b[:, 0] << 8) | (b[:, 1] << 16) | (b[:, 2] << 24)).astype(np.uint32).view(np.int32)
        x = v.astype(np.float32)*np.float32(1/2147483648)
    elif fmt == N.PCM_S32:
        x = raw.view("<i4").astype(np.float32)*np.float32(1/2147483648)
    elif fmt == N.PCM_F32:
        x = raw.view("<f4").astype(np.float32)
    elif fmt == N.PCM_F64:
        x = raw.view("<f8").astype(np.float32)
    else:
        raise ValueError(f"unknown sample format {fmt}")
    return x.reshape(-1, channels)


def read_wav(path) -> tuple[np.ndarray, int]:
    """→ (pcm float32 (channels, samples), samplerate) on the host"""
    info = parse_wav(path)
    raw = np.fromfile(info.path, dtype=np.uint8, offset=info.data_offset, count=info.frames*info.block)
    return np.ascontiguousarray(decode_pcm(raw, info.format, info.channels).T), info.samplerate


def read_flac(path, verify: bool = True) -> tuple[np.ndarray, int]:
    """→ (pcm float32 (channels, samples), samplerate). The stream is decoded on the host by csrc/flac.cu (every frame's
    CRC checked there); the MD5 of STREAMINFO, when the encoder recorded one, is checked here against the decoded
    samples. Conversion to float as libswresample does for s16 / s32 sources: x / 2^(bits−1)."""
    import hashlib
    data = np.fromfile(path, dtype=np.uint8)
    info, pcm = N.flac_decode(data)
    bits = int(info.bits_per_sample)
    if verify and info.has_md5:
        width = (bits + 7)//8
        packed = pcm.astype("<i4").view(np.uint8).reshape(-1, 4)[:, :width]          # little-endian, sign-extended to whole bytes
        if hashlib.md5(np.ascontiguousarray(packed).tobytes()).digest() != bytes(info.md5):
            raise ValueError(f"'{path}': the decoded samples do not match the stream's MD5 signature")
    scale = np.float32(1.0/float(1 << (bits - 1)))
    return np.ascontiguousarray((pcm.astype(np.float32)*scale).T), int(info.samplerate)


def upload_wav(ctx: "N.Context", info: WavInfo, device: int, staging_bytes: int = 32 << 20):
    """The file's samples → a resident planar float32 clip [channels][frames] on `device`: memory-mapped file →
    pinned staging (double-buffered) → H2D → sfb_pcm_ingest. Returns the torch tensor."""
    import torch
    dev = f"cuda:{device}"
    clip = torch.empty((info.channels, info.frames), dtype=torch.float32, device=dev)
    if info.frames == 0:
        return clip
    per = max(1, staging_bytes//info.block)                    # sample frames per staging buffer
    mapped = np.memmap(info.path, dtype=np.uint8, mode="r", offset=info.data_offset, shape=(info.frames*info.block,))
    pinned = [torch.empty(per*info.block, dtype=torch.uint8).pin_memory() for _ in range(2)]
    device_raw = [torch.empty(per*info.block, dtype=torch.uint8, device=dev) for _ in range(2)]
    consumed = [None, None]                                    # event: the H2D out of staging buffer k has finished
    stream = torch.cuda.current_stream(device)
    for index, first in enumerate(range(0, info.frames, per)):
        k = index & 1
        count = min(per, info.frames - first)
        if consumed[k] is not None:
            consumed[k].synchronize()                          # the pinned buffer is free again
        view = pinned[k][:count*info.block]
        view.numpy()[:] = mapped[first*info.block:(first + count)*info.block]      # page cache → pinned memory
        device_raw[k][:count*info.block].copy_(view, non_blocking=True)
        consumed[k] = torch.cuda.Event()
        consumed[k].record(stream)
        ctx.pcm_ingest(device_raw[k], count, info.channels, info.format, clip, info.frames, first)
    return clip

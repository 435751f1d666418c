"""FFmpeg's FLAC codec as an independent reference for sfb_flac_decode — test infrastructure only.

This image has no ffmpeg binary, but the OpenCV wheel bundles libavformat / libavcodec / libavutil (FFmpeg 8: avcodec 62),
with the `flac` encoder and decoder compiled in. They are driven here through ctypes:

    decode(stream)                        FLAC bytes → int samples (frames, channels), by FFmpeg's demuxer + decoder
    encode(samples, bits, ...)            int samples → FLAC bytes written by FFmpeg's ENCODER (its LPC analysis, Rice
                                          partitioning, stereo decorrelation and wasted-bits detection — what real files
                                          contain), framed by a STREAMINFO block built here (with the MD5 of the samples)

There are no FFmpeg headers in the image, so (with one checked exception in transcode(), for encoders other than FLAC) no
struct is ever WRITTEN: codec contexts are configured through
avcodec_parameters_to_context + AVOptions, and the encoder is fed the very AVFrames FFmpeg's decoder produced from a
verbatim-coded stream of the samples (tests/flac_writer.py). The few fields that are READ sit at the start of their
structs and have not moved since FFmpeg 5 — AVFormatContext.nb_streams / .streams, AVStream.codecpar, AVFrame.data[0] /
.nb_samples / .format, AVPacket.data / .size — and every read is sanity-checked before a pointer is followed.
"""
from __future__ import annotations

import ctypes
import glob
import hashlib
import importlib.util
import os
import struct
import tempfile
from pathlib import Path

import numpy as np

P = ctypes.c_void_p
_LIBS = None
_LIBS_EXTRA = {}
AV_SAMPLE_FMT_S16, AV_SAMPLE_FMT_S32 = 1, 2
AV_OPT_SEARCH_CHILDREN = 1


def _load():
    global _LIBS
    if _LIBS is not None:
        return _LIBS or None
    _LIBS = False
    spec = importlib.util.find_spec("cv2")
    if spec is None or not spec.submodule_search_locations:
        return None
    try:
        import cv2  # noqa: F401 — loads the bundled FFmpeg libraries and what they depend on (they carry no RPATH of their own)
    except Exception:
        return None
    site = Path(list(spec.submodule_search_locations)[0]).parent
    found = {}
    for stem in ("avutil", "swresample", "avcodec", "avformat"):
        hits = [h for d in ("opencv_python_headless.libs", "opencv_python.libs", "opencv_contrib_python.libs")
                for h in glob.glob(str(site/d/f"lib{stem}-*.so*"))]
        if not hits:
            return None
        try:
            found[stem] = ctypes.CDLL(hits[0], mode=ctypes.RTLD_GLOBAL)
        except OSError:
            return None
    u, c, f, r = found["avutil"], found["avcodec"], found["avformat"], found["swresample"]
    r.swr_alloc.restype = P
    r.swr_init.argtypes = [P]
    r.swr_convert.argtypes = [P, P, ctypes.c_int, P, ctypes.c_int]
    r.swr_free.argtypes = [ctypes.POINTER(P)]
    u.av_opt_set_sample_fmt.argtypes = [P, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
    c.avcodec_version.restype = ctypes.c_uint
    if not (59 <= (c.avcodec_version() >> 16) <= 62):              # the layouts read below were checked for FFmpeg 5 … 8
        return None
    f.avformat_open_input.argtypes = [ctypes.POINTER(P), ctypes.c_char_p, P, P]
    f.avformat_find_stream_info.argtypes = [P, P]
    f.av_read_frame.argtypes = [P, P]
    f.avformat_close_input.argtypes = [ctypes.POINTER(P)]
    for name in ("avcodec_find_decoder_by_name", "avcodec_find_encoder_by_name"):
        getattr(c, name).restype, getattr(c, name).argtypes = P, [ctypes.c_char_p]
    c.avcodec_find_decoder.restype, c.avcodec_find_decoder.argtypes = P, [ctypes.c_int]
    c.avcodec_alloc_context3.restype, c.avcodec_alloc_context3.argtypes = P, [P]
    c.avcodec_parameters_to_context.argtypes = [P, P]
    c.avcodec_open2.argtypes = [P, P, P]
    c.avcodec_free_context.argtypes = [ctypes.POINTER(P)]
    c.av_packet_alloc.restype = P
    c.av_packet_unref.argtypes = [P]
    c.av_packet_free.argtypes = [ctypes.POINTER(P)]
    for name in ("avcodec_send_packet", "avcodec_receive_frame", "avcodec_send_frame", "avcodec_receive_packet"):
        getattr(c, name).argtypes = [P, P]
    u.av_frame_alloc.restype = P
    u.av_frame_unref.argtypes = [P]
    u.av_frame_free.argtypes = [ctypes.POINTER(P)]
    u.av_opt_set.argtypes = [P, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
    u.av_opt_set_int.argtypes = [P, ctypes.c_char_p, ctypes.c_int64, ctypes.c_int]
    u.av_log_set_level.argtypes = [ctypes.c_int]
    u.av_log_set_level(8)                                            # AV_LOG_FATAL: the tests report, not stderr
    if not (c.avcodec_find_decoder_by_name(b"flac") and c.avcodec_find_encoder_by_name(b"flac")):
        return None
    _LIBS = (u, c, f)
    _LIBS_EXTRA["swresample"] = r
    return _LIBS


def available() -> bool:
    return _load() is not None


def _u32(raw: bytes, at: int) -> int:
    return int.from_bytes(raw[at:at + 4], "little")


def _u64(raw: bytes, at: int) -> int:
    return int.from_bytes(raw[at:at + 8], "little")


class _Demuxed:
    """An opened audio file: its only stream's codec parameters and a decoder on them (`codec`: the decoder expected,
    by name; None = whatever the demuxer found)"""
    def __init__(self, path: str, codec: str | None = "flac"):
        u, c, f = _load()
        self.fmt = P()
        if f.avformat_open_input(ctypes.byref(self.fmt), path.encode(), None, None) != 0:
            raise RuntimeError(f"FFmpeg cannot open {path}")
        if f.avformat_find_stream_info(self.fmt, None) < 0:
            raise RuntimeError("FFmpeg found no stream information")
        head = ctypes.string_at(self.fmt.value, 56)
        streams = _u64(head, 48)
        if _u32(head, 44) != 1 or not streams:                      # AVFormatContext.nb_streams / .streams
            raise RuntimeError("unexpected AVFormatContext layout")
        stream = ctypes.string_at(_u64(ctypes.string_at(streams, 8), 0), 24)
        if _u32(stream, 8) != 0:                                     # AVStream.index of the first stream
            raise RuntimeError("unexpected AVStream layout")
        self.codecpar = _u64(stream, 16)
        par = ctypes.string_at(self.codecpar, 8)
        decoder = c.avcodec_find_decoder_by_name(codec.encode()) if codec else c.avcodec_find_decoder(_u32(par, 4))
        if _u32(par, 0) != 1 or not decoder or (codec and _u32(par, 4) != _u32(ctypes.string_at(decoder, 24), 20)):
            raise RuntimeError("unexpected AVCodecParameters layout")        # AVMEDIA_TYPE_AUDIO, AVCodec.id of the named codec
        self.decoder = P(c.avcodec_alloc_context3(decoder))
        if c.avcodec_parameters_to_context(self.decoder, self.codecpar) != 0 or c.avcodec_open2(self.decoder, decoder, None) != 0:
            raise RuntimeError("FFmpeg's decoder did not open")
        self.packet, self.frame = P(c.av_packet_alloc()), P(u.av_frame_alloc())

    def frames(self):
        """Yields the decoder's AVFrame after every decoded block (valid until the next one)"""
        u, c, f = _load()
        def drain():
            while c.avcodec_receive_frame(self.decoder, self.frame) == 0:
                yield self.frame
                u.av_frame_unref(self.frame)
        while f.av_read_frame(self.fmt, self.packet) == 0:
            sent = c.avcodec_send_packet(self.decoder, self.packet)
            c.av_packet_unref(self.packet)
            if sent != 0:
                raise RuntimeError(f"FFmpeg's FLAC decoder refused a packet ({sent})")
            yield from drain()
        c.avcodec_send_packet(self.decoder, None)
        yield from drain()

    def close(self):
        u, c, f = _load()
        c.av_packet_free(ctypes.byref(self.packet)); u.av_frame_free(ctypes.byref(self.frame))
        c.avcodec_free_context(ctypes.byref(self.decoder)); f.avformat_close_input(ctypes.byref(self.fmt))


def _frame_samples(frame: P, channels: int) -> np.ndarray:
    head = ctypes.string_at(frame.value, 120)
    data, count, fmt = _u64(head, 0), _u32(head, 112), _u32(head, 116)       # AVFrame.data[0] / .nb_samples / .format
    if not data or not (0 < count <= 65535) or fmt not in (AV_SAMPLE_FMT_S16, AV_SAMPLE_FMT_S32):
        raise RuntimeError("unexpected AVFrame layout")
    dtype = np.int16 if fmt == AV_SAMPLE_FMT_S16 else np.int32
    return np.frombuffer(ctypes.string_at(data, count*channels*np.dtype(dtype).itemsize), dtype=dtype).reshape(count, channels).astype(np.int64)


def decode(stream: bytes, channels: int, bits: int) -> np.ndarray:
    """FLAC bytes → (frames, channels) int64 by FFmpeg. Its decoder left-justifies samples in 16 / 32 bits"""
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "stream.flac")
        Path(path).write_bytes(stream)
        source = _Demuxed(path)
        try:
            blocks = [_frame_samples(frame, channels) for frame in source.frames()]
        finally:
            source.close()
    container = 16 if bits <= 16 else 32
    return np.concatenate(blocks) >> (container - bits)


def streaminfo(samples: np.ndarray, bits: int, blocksize: int, rate: int, frames: list[bytes]) -> bytes:
    """'fLaC' + the STREAMINFO block (last metadata block) of a stream holding `samples` at `bits` per sample"""
    count, channels = samples.shape
    width = (bits + 7)//8
    md5 = hashlib.md5(b"".join(int(v).to_bytes(width, "little", signed=True) for v in samples.reshape(-1))).digest()
    packed = (rate << 44) | ((channels - 1) << 41) | ((bits - 1) << 36) | count
    body = (struct.pack(">HH", blocksize, blocksize) + min(map(len, frames)).to_bytes(3, "big") + max(map(len, frames)).to_bytes(3, "big")
            + packed.to_bytes(8, "big") + md5)
    return b"fLaC" + bytes([0x80]) + len(body).to_bytes(3, "big") + body


def transcode(samples: np.ndarray, codec: str, bits: int, blocksize: int, rate: int, settings: dict, sync=None) -> list[bytes]:
    """The packets FFmpeg's encoder `codec` makes of `samples`: a verbatim FLAC stream of them (tests/flac_writer.py) is
    decoded by FFmpeg and the decoder's AVFrames go straight into the encoder, whose context is configured from the
    demuxed stream's parameters + `settings` (AVOptions). `sync(packet)` sanity-checks the packet layout read here"""
    from tests.flac_writer import write_flac
    u, c, f = _load()
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "verbatim.flac")
        Path(path).write_bytes(write_flac(samples, rate=rate, bits=bits, blocksize=blocksize, subframe="verbatim"))
        source = _Demuxed(path)
        encoder_kind = c.avcodec_find_encoder_by_name(codec.encode())
        encoder = P(c.avcodec_alloc_context3(encoder_kind))
        packet = P(c.av_packet_alloc())
        frames: list[bytes] = []
        try:
            if c.avcodec_parameters_to_context(encoder, source.codecpar) != 0:
                raise RuntimeError("encoder parameters")
            if codec != "flac":
                # the parameters carried FLAC's codec id along. The one write into an FFmpeg struct in this file:
                # AVCodecContext.codec_id (after av_class, log_level_offset, codec_type, codec — unchanged since FFmpeg 1),
                # made only where the field demonstrably is: it must hold FLAC's id, next to codec_type = audio
                head = ctypes.string_at(encoder.value, 28)
                flac_id = _u32(ctypes.string_at(c.avcodec_find_decoder_by_name(b"flac"), 24), 20)
                if _u32(head, 12) != 1 or _u32(head, 24) != flac_id:
                    raise RuntimeError("unexpected AVCodecContext layout")
                ctypes.memmove(encoder.value + 24, ctypes.string_at(encoder_kind, 24)[20:24], 4)
            for key, value in settings.items():
                if u.av_opt_set(encoder, key.encode(), str(value).encode(), AV_OPT_SEARCH_CHILDREN) != 0:
                    raise RuntimeError(f"FFmpeg's {codec} encoder has no option {key}={value}")
            if c.avcodec_open2(encoder, encoder_kind, None) != 0:
                raise RuntimeError(f"FFmpeg's {codec} encoder did not open")

            def pull():
                while c.avcodec_receive_packet(encoder, packet) == 0:
                    head = ctypes.string_at(packet.value, 36)
                    data, size = _u64(head, 24), _u32(head, 32)              # AVPacket.data / .size
                    if size:
                        block = ctypes.string_at(data, size)
                        if sync is not None and not sync(block):
                            raise RuntimeError("unexpected AVPacket layout")
                        frames.append(block)
                    c.av_packet_unref(packet)
            for frame in source.frames():
                if c.avcodec_send_frame(encoder, frame) != 0:
                    raise RuntimeError(f"FFmpeg's {codec} encoder refused a frame")
                pull()
            c.avcodec_send_frame(encoder, None)
            pull()
        finally:
            c.av_packet_free(ctypes.byref(packet)); c.avcodec_free_context(ctypes.byref(encoder)); source.close()
    return frames


def encode(samples: np.ndarray, bits: int = 16, blocksize: int = 4096, rate: int = 44100, level: int = 5, **private) -> tuple[bytes, np.ndarray, int]:
    """(frames, channels) ints → (FLAC stream by FFmpeg's encoder, the samples it holds, their bit depth).

    The encoder's input is what FFmpeg's decoder makes of a verbatim-coded stream of `samples`, so 8-bit material
    arrives left-justified in 16 bits (the stream is 16-bit with 8 wasted bits) and 20-bit material in 24 — hence the
    returned samples / depth.
    `private` are the encoder's own AVOptions (lpc_type, lpc_passes, ch_mode, min_partition_order, …)."""
    samples = np.asarray(samples, np.int64)
    held_bits = 16 if bits <= 16 else (24 if bits <= 24 else 32)       # flacenc.c: S16 input is 16-bit, S32 is 24-bit or 32-bit
    held = samples << (held_bits - bits)
    settings = dict(time_base=f"1/{rate}", frame_size=blocksize, compression_level=level, strict=-2, **private)
    frames = transcode(samples, "flac", bits, blocksize, rate, settings, sync=lambda block: block[0] == 0xFF and (block[1] & 0xFC) == 0xF8)
    return streaminfo(held, held_bits, blocksize, rate, frames) + b"".join(frames), held, held_bits


def wav_as_f32le(path, channels: int, rate: int) -> np.ndarray:
    """What `ffmpeg -i file.wav -f f32le -` delivers (the reference's BrokenAudioReader, ffmpeg.py:1279-1330): FFmpeg's
    WAV demuxer + PCM decoder, then libswresample to packed float32. → (frames, channels) float32"""
    u, c, f = _load()
    r = _LIBS_EXTRA["swresample"]
    source = _Demuxed(str(path), codec=None)
    layout = {1: b"mono", 2: b"stereo"}[channels]
    converters, blocks = {}, []
    try:
        for frame in source.frames():
            head = ctypes.string_at(frame.value, 120)
            count, fmt = _u32(head, 112), _u32(head, 116)
            if not (0 < count <= (1 << 20)) or fmt > 4:              # u8, s16, s32, flt, dbl (packed): what PCM decoders emit
                raise RuntimeError("unexpected AVFrame layout")
            if fmt not in converters:
                swr = P(r.swr_alloc())
                for key in (b"in_chlayout", b"out_chlayout"):
                    assert u.av_opt_set(swr, key, layout, 0) == 0
                for key in (b"in_sample_rate", b"out_sample_rate"):
                    assert u.av_opt_set_int(swr, key, rate, 0) == 0
                assert u.av_opt_set_sample_fmt(swr, b"in_sample_fmt", fmt, 0) == 0
                assert u.av_opt_set_sample_fmt(swr, b"out_sample_fmt", 3, 0) == 0           # AV_SAMPLE_FMT_FLT
                assert r.swr_init(swr) == 0
                converters[fmt] = swr
            out = np.empty((count, channels), np.float32)
            planes = (P*8)(out.ctypes.data)
            done = r.swr_convert(converters[fmt], ctypes.cast(planes, P), count, frame, count)   # AVFrame starts with data[8]
            if done != count:
                raise RuntimeError(f"swr_convert returned {done} of {count}")
            blocks.append(out)
    finally:
        for swr in converters.values():
            r.swr_free(ctypes.byref(swr))
        source.close()
    return np.concatenate(blocks)

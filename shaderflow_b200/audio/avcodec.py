"""Compressed audio files (opus, mp3, ogg, m4a, …) without an ffmpeg binary.

The reference decodes every audio file through an ffmpeg child asked for float32 (BrokenAudioReader, ffmpeg.py:1240-1333).
WAV and FLAC are read natively (audio/reader.py); for everything else this module runs the SAME libraries in process when
the OpenCV wheel is installed — it bundles libavformat / libavcodec / libswresample — through ctypes:

    demuxer → the stream's decoder → libswresample to packed float32 (the conversion `-f f32le` performs) → planar clip

No FFmpeg headers ship with the wheel, so nothing is written into FFmpeg's structs: contexts are set up with
avcodec_parameters_to_context and AVOptions. Five fields are read, all at the head of their structs and unchanged since
FFmpeg 5 (AVFormatContext.nb_streams / .streams, AVStream.index / .codecpar, AVCodecParameters.codec_type, AVFrame
.nb_samples / .format, AVPacket.stream_index); each is range-checked before it is used, the library version is checked
first, and anything unexpected raises instead of guessing. Host-side, once per export; the clip then goes to HBM like
any other."""
from __future__ import annotations

import ctypes
import glob
import importlib.util
from pathlib import Path

import numpy as np

P = ctypes.c_void_p
_libs = None
AVMEDIA_TYPE_AUDIO = 1
AV_SAMPLE_FMT_FLT = 3
_SAMPLE_FORMATS = 12                       # u8 … s64p
_LAYOUT_CHANNELS = {"mono": 1, "stereo": 2, "2.1": 3, "3.0": 3, "3.0(back)": 3, "4.0": 4, "quad": 4, "quad(side)": 4, "3.1": 4,
                    "5.0": 5, "5.0(side)": 5, "4.1": 5, "5.1": 6, "5.1(side)": 6, "6.0": 6, "6.1": 7, "7.0": 7, "7.1": 8}


def _load():
    global _libs
    if _libs is not None:
        return _libs or None
    _libs = False
    spec = importlib.util.find_spec("cv2")
    if spec is None or not spec.submodule_search_locations:
        return None
    try:
        import cv2  # noqa: F401 — brings the bundled FFmpeg libraries and their dependencies into the process
    except Exception:
        return None
    site = Path(list(spec.submodule_search_locations)[0]).parent
    found = {}
    for stem in ("avutil", "swresample", "avcodec", "avformat"):
        hits = [h for d in ("opencv_python_headless.libs", "opencv_python.libs", "opencv_contrib_python.libs",
                            "opencv_contrib_python_headless.libs") for h in glob.glob(str(site/d/f"lib{stem}-*.so*"))]
        if not hits:
            return None
        try:
            found[stem] = ctypes.CDLL(hits[0], mode=ctypes.RTLD_GLOBAL)
        except OSError:
            return None
    u, r, c, f = found["avutil"], found["swresample"], found["avcodec"], found["avformat"]
    c.avcodec_version.restype = ctypes.c_uint
    if not (59 <= (c.avcodec_version() >> 16) <= 62):              # FFmpeg 5 … 8: the layouts this module reads
        return None
    f.avformat_open_input.argtypes = [ctypes.POINTER(P), ctypes.c_char_p, P, P]
    f.avformat_find_stream_info.argtypes = [P, P]
    f.av_find_best_stream.argtypes = [P, ctypes.c_int, ctypes.c_int, ctypes.c_int, P, ctypes.c_int]
    f.av_read_frame.argtypes = [P, P]
    f.avformat_close_input.argtypes = [ctypes.POINTER(P)]
    c.avcodec_find_decoder.restype, c.avcodec_find_decoder.argtypes = P, [ctypes.c_int]
    c.avcodec_alloc_context3.restype, c.avcodec_alloc_context3.argtypes = P, [P]
    c.avcodec_parameters_to_context.argtypes = [P, P]
    c.avcodec_open2.argtypes = [P, P, P]
    c.avcodec_free_context.argtypes = [ctypes.POINTER(P)]
    c.av_packet_alloc.restype = P
    c.av_packet_unref.argtypes = [P]
    c.av_packet_free.argtypes = [ctypes.POINTER(P)]
    c.avcodec_send_packet.argtypes = [P, P]
    c.avcodec_receive_frame.argtypes = [P, P]
    u.av_frame_alloc.restype = P
    u.av_frame_unref.argtypes = [P]
    u.av_frame_free.argtypes = [ctypes.POINTER(P)]
    u.av_opt_set.argtypes = [P, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
    u.av_opt_set_int.argtypes = [P, ctypes.c_char_p, ctypes.c_int64, ctypes.c_int]
    u.av_opt_set_sample_fmt.argtypes = [P, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
    u.av_opt_get_int.argtypes = [P, ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int64)]
    u.av_opt_get.argtypes = [P, ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(P)]
    u.av_free.argtypes = [P]
    u.av_log_set_level.argtypes = [ctypes.c_int]
    r.swr_alloc.restype = P
    r.swr_init.argtypes = [P]
    r.swr_convert.argtypes = [P, P, ctypes.c_int, P, ctypes.c_int]
    r.swr_free.argtypes = [ctypes.POINTER(P)]
    _libs = (u, r, c, f)
    return _libs


def available() -> bool:
    return _load() is not None


def _u32(raw: bytes, at: int) -> int:
    return int.from_bytes(raw[at:at + 4], "little")


def _u64(raw: bytes, at: int) -> int:
    return int.from_bytes(raw[at:at + 8], "little")


def _layout_channels(name: str) -> int:
    if name in _LAYOUT_CHANNELS:
        return _LAYOUT_CHANNELS[name]
    if name.endswith(" channels") and name.split()[0].isdigit():
        return int(name.split()[0])
    raise RuntimeError(f"unknown channel layout '{name}'")


def decode_file(path, options: dict | None = None) -> tuple[np.ndarray, int]:
    """→ (pcm float32 (channels, samples), samplerate): the first audio stream of `path`, decoded by FFmpeg's libraries
    and converted like `ffmpeg -f f32le`. `options`: AVOptions for the decoder (tests use request_sample_fmt)"""
    libs = _load()
    if libs is None:
        raise RuntimeError("the FFmpeg libraries of the opencv package are not available")
    u, r, c, f = libs
    u.av_log_set_level(16)                                          # AV_LOG_ERROR
    fmt, decoder, packet, frame, swr = P(), P(), P(), P(), P()
    blocks: list[np.ndarray] = []
    try:
        if f.avformat_open_input(ctypes.byref(fmt), str(path).encode(), None, None) != 0:
            raise RuntimeError(f"FFmpeg cannot open '{path}'")
        if f.avformat_find_stream_info(fmt, None) < 0:
            raise RuntimeError(f"'{path}': no stream information")
        wanted = f.av_find_best_stream(fmt, AVMEDIA_TYPE_AUDIO, -1, -1, None, 0)
        head = ctypes.string_at(fmt.value, 56)
        count, streams = _u32(head, 44), _u64(head, 48)             # AVFormatContext.nb_streams / .streams
        if wanted < 0:
            raise RuntimeError(f"'{path}' has no audio stream")
        if not (0 < count <= 64) or wanted >= count or not streams:
            raise RuntimeError("unexpected AVFormatContext layout")
        stream = ctypes.string_at(_u64(ctypes.string_at(streams + 8*wanted, 8), 0), 24)
        codecpar = _u64(stream, 16)                                 # AVStream.index / .codecpar
        if _u32(stream, 8) != wanted or not codecpar:
            raise RuntimeError("unexpected AVStream layout")
        par = ctypes.string_at(codecpar, 8)
        if _u32(par, 0) != AVMEDIA_TYPE_AUDIO:                      # AVCodecParameters.codec_type / .codec_id
            raise RuntimeError("unexpected AVCodecParameters layout")
        kind = c.avcodec_find_decoder(_u32(par, 4))
        if not kind:
            raise RuntimeError(f"'{path}': this FFmpeg build has no decoder for the stream")
        decoder = P(c.avcodec_alloc_context3(kind))
        if c.avcodec_parameters_to_context(decoder, codecpar) != 0:
            raise RuntimeError("avcodec_parameters_to_context failed")
        for key, value in (options or {}).items():
            if u.av_opt_set(decoder, key.encode(), str(value).encode(), 1) != 0:
                raise RuntimeError(f"decoder option {key}={value} was refused")
        if c.avcodec_open2(decoder, kind, None) != 0:
            raise RuntimeError(f"'{path}': the decoder did not open")
        rate = ctypes.c_int64()
        text = P()
        if u.av_opt_get_int(decoder, b"ar", 0, ctypes.byref(rate)) != 0 or u.av_opt_get(decoder, b"ch_layout", 0, ctypes.byref(text)) != 0:
            raise RuntimeError("the decoder does not tell its sample rate / channel layout")
        layout = ctypes.string_at(text.value)
        u.av_free(text)
        channels = _layout_channels(layout.decode())
        if not (1000 <= rate.value <= 768000):
            raise RuntimeError(f"implausible sample rate {rate.value}")
        packet, frame = P(c.av_packet_alloc()), P(u.av_frame_alloc())
        swr_format = None

        def drain():
            nonlocal swr, swr_format
            while c.avcodec_receive_frame(decoder, frame) == 0:
                head = ctypes.string_at(frame.value, 120)
                samples, sample_format = _u32(head, 112), _u32(head, 116)        # AVFrame.nb_samples / .format
                if not (0 < samples <= (1 << 22)) or sample_format >= _SAMPLE_FORMATS:
                    raise RuntimeError("unexpected AVFrame layout")
                if swr_format != sample_format:                     # (a decoder keeps one format; be exact anyway)
                    if swr:
                        r.swr_free(ctypes.byref(swr))
                    swr = P(r.swr_alloc())
                    ok = all(u.av_opt_set(swr, key, layout, 0) == 0 for key in (b"in_chlayout", b"out_chlayout"))
                    ok = ok and all(u.av_opt_set_int(swr, key, rate.value, 0) == 0 for key in (b"in_sample_rate", b"out_sample_rate"))
                    ok = ok and u.av_opt_set_sample_fmt(swr, b"in_sample_fmt", sample_format, 0) == 0
                    ok = ok and u.av_opt_set_sample_fmt(swr, b"out_sample_fmt", AV_SAMPLE_FMT_FLT, 0) == 0
                    if not ok or r.swr_init(swr) != 0:
                        raise RuntimeError("libswresample did not accept the stream's format")
                    swr_format = sample_format
                out = np.empty((samples, channels), np.float32)
                planes = (P*8)(out.ctypes.data)
                # AVFrame starts with data[8]: the frame pointer IS the array of plane pointers swr_convert takes
                if r.swr_convert(swr, ctypes.cast(planes, P), samples, frame, samples) != samples:
                    raise RuntimeError("swr_convert came up short")
                blocks.append(out)
                u.av_frame_unref(frame)

        while f.av_read_frame(fmt, packet) == 0:
            mine = _u32(ctypes.string_at(packet.value, 40), 36) == wanted        # AVPacket.stream_index
            if mine and c.avcodec_send_packet(decoder, packet) == 0:
                drain()
            c.av_packet_unref(packet)
        c.avcodec_send_packet(decoder, None)
        drain()
    finally:
        if swr:
            r.swr_free(ctypes.byref(swr))
        if packet:
            c.av_packet_free(ctypes.byref(packet))
        if frame:
            u.av_frame_free(ctypes.byref(frame))
        if decoder:
            c.avcodec_free_context(ctypes.byref(decoder))
        if fmt:
            f.avformat_close_input(ctypes.byref(fmt))
    if not blocks:
        raise RuntimeError(f"'{path}': no audio was decoded")
    return np.ascontiguousarray(np.concatenate(blocks).T), int(rate.value)

"""
ctypes binding of libsfb200.so (include/sfb200.h) — the only door between the Python host and the
CUDA backend. PyTorch tensors are used purely as device buffers: every wrapper takes `tensor.data_ptr()`.

There is no CPU fallback: `lib()` raises if the library is missing, `Context()` raises without a GPU.
"""
from __future__ import annotations

import ctypes as C
from ctypes import POINTER, byref, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p
from pathlib import Path

LIBRARY = Path(__file__).resolve().parent/"libsfb200.so"

MAX_EXTRA, MAX_SAMPLERS = 16, 20
OK, EINVAL, ECUDA, ENOTFOUND, ENOMEM, EIO, ESTATE = 0, -1, -2, -3, -4, -5, -6
WINDOW = dict(hanning=0, hann_poisson=1, none=2)
MAGNITUDE = dict(power=0, amplitude=1)
VOLUME = dict(linear=0, sqrt=1, dbfs=2, dbfs_tremx=3)
REDUCER = dict(average=0, rms=1, std=2)
DTYPE_U8, DTYPE_F32, DTYPE_F16 = 0, 1, 2
FILTER_NEAREST, FILTER_LINEAR = 0, 1
FILTER_EXACT, FILTER_HARDWARE, RENDER_LITERAL, RENDER_TILED = 0, 1, 2, 4
PCM_U8, PCM_S16, PCM_S24, PCM_S32, PCM_F32, PCM_F64 = range(6)
SCALARS = 5
SCENE_COUNT = 17
SCENE_PROGRAM_BASE = 1000
JIT_FMAD = 1
VIDEO_RGB24, VIDEO_RGBA32, VIDEO_YUV420P, VIDEO_YUV422P, VIDEO_YUV444P, VIDEO_FULL_RANGE = 0, 1, 2, 3, 4, 0x100
SCALAR_VOLUME, SCALAR_VOLUME_INTEGRAL, SCALAR_STD, SCALAR_VOLUME_TARGET, SCALAR_STD_TARGET = range(5)


class Uniforms(C.Structure):
    """struct sfb_uniforms"""
    _fields_ = [
        ("iTime", c_float), ("iTau", c_float), ("iDuration", c_float), ("iDeltatime", c_float),
        ("iResolution", c_float*2),
        ("iWantAspect", c_float), ("iQuality", c_float), ("iSSAA", c_float), ("iFramerate", c_float),
        ("iFrame", c_int32), ("iRealtime", c_int32), ("iLayer", c_int32), ("iMouseInside", c_int32),
        ("iMouse", c_float*2),
        ("iMouse1", c_int32), ("iMouse2", c_int32),
        ("iCameraMode", c_int32), ("iCameraProjection", c_int32),
        ("iCameraPosition", c_float*3), ("iCameraRight", c_float*3), ("iCameraUpward", c_float*3),
        ("iCameraForward", c_float*3), ("iCameraZenith", c_float*3),
        ("iCameraZoom", c_float), ("iCameraIsometric", c_float), ("iCameraFocalLength", c_float),
        ("iCameraOrbital", c_float), ("iCameraDolly", c_float), ("iCameraSeparation", c_float),
        ("extra", (c_float*4)*MAX_EXTRA),
    ]

    @classmethod
    def defaults(cls, width: int = 1920, height: int = 1080) -> "Uniforms":
        """A freshly built scene in an export (scene.py:687-703, camera.py:146-201)"""
        u = cls()
        u.iDuration, u.iResolution[:] = 10.0, (width, height)
        u.iWantAspect, u.iQuality, u.iSSAA, u.iFramerate = width/height, 0.5, 1.0, 60.0
        u.iCameraMode = 1
        u.iCameraRight[:], u.iCameraUpward[:], u.iCameraForward[:] = (1, 0, 0), (0, 1, 0), (0, 0, 1)
        u.iCameraZenith[:] = (0, 1, 0)
        u.iCameraZoom, u.iCameraFocalLength, u.iCameraSeparation = 1.0, 1.0, 0.05
        return u


class FlacInfo(C.Structure):
    """struct sfb_flac_info"""
    _fields_ = [("samplerate", c_int32), ("channels", c_int32), ("bits_per_sample", c_int32), ("has_md5", c_int32),
                ("total_samples", c_int64), ("min_block", c_int32), ("max_block", c_int32), ("md5", C.c_uint8*16)]


class DynamicsParams(C.Structure):
    _fields_ = [("frequency", c_double), ("zeta", c_double), ("response", c_double), ("precision", c_double)]


class PianoNote(C.Structure):
    """struct sfb_piano_note"""
    _fields_ = [("start", c_double), ("end", c_double), ("note", c_int32), ("channel", c_int32),
                ("velocity", c_int32), ("order", c_int32)]


class SceneInfo(C.Structure):
    _fields_ = [
        ("name", c_char_p), ("reference", c_char_p),
        ("n_extra", c_int), ("extra", c_char_p*MAX_EXTRA),
        ("n_samplers", c_int), ("samplers", c_char_p*MAX_SAMPLERS),
        ("n_required", c_int),
    ]


_PROTOTYPES = dict(
    sfb_version=(c_int, []),
    sfb_last_error=(c_char_p, []),
    sfb_device_count=(c_int, [POINTER(c_int)]),
    sfb_ctx_create=(c_int, [c_int, c_void_p, POINTER(c_void_p)]),
    sfb_ctx_destroy=(c_int, [c_void_p]),
    sfb_ctx_set_stream=(c_int, [c_void_p, c_void_p]),
    sfb_sync=(c_int, [c_void_p]),
    sfb_launch_count=(c_int, [c_void_p, POINTER(c_uint64)]),
    sfb_ctx_enable_peer=(c_int, [c_void_p, c_int]),
    sfb_frame_clock=(c_int, [c_int, c_double, c_double, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p]),
    sfb_stft_mel=(c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_int, c_int, c_int,
                          c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    sfb_audio_track=(c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_int,
                             c_void_p, c_int, POINTER(DynamicsParams), c_void_p, c_void_p, c_int, c_int, c_int]),
    sfb_tex_create=(c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_void_p)]),
    sfb_tex_destroy=(c_int, [c_void_p]),
    sfb_tex_set_sampling=(c_int, [c_void_p, c_int, c_int, c_int]),
    sfb_tex_write=(c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int]),
    sfb_tex_bind_external=(c_int, [c_void_p, c_void_p]),
    sfb_tex_read=(c_int, [c_void_p, c_void_p]),
    sfb_tex_storage=(c_int, [c_void_p, POINTER(c_void_p), POINTER(c_size_t)]),
    sfb_tex_sample=(c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    sfb_scene_lookup=(c_int, [c_char_p, POINTER(c_int)]),
    sfb_scene_info_get=(c_int, [c_int, POINTER(SceneInfo)]),
    sfb_render_screen=(c_int, [c_void_p, c_int, POINTER(Uniforms), POINTER(c_void_p), c_int, c_int,
                               c_int, c_int, c_void_p, c_void_p]),
    sfb_pcm_ingest=(c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_int64, c_int64]),
    sfb_dynamics_scan=(c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, POINTER(DynamicsParams)]),
    sfb_piano_track=(c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_double, c_double, c_double, c_double,
                             c_int, c_int, c_void_p, c_void_p, c_void_p]),
    sfb_piano_roll=(c_int, [c_void_p, c_void_p, c_void_p, c_double, c_double, c_double, c_int, c_int, c_void_p, c_void_p]),
    sfb_visualizer_plan=(c_int, [POINTER(Uniforms), c_int, c_int, c_int, c_int, c_int, POINTER(c_int), POINTER(c_int)]),
    sfb_render_target=(c_int, [c_void_p, c_int, POINTER(Uniforms), POINTER(c_void_p), c_int, c_int, c_void_p]),
    sfb_render_final=(c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    sfb_render_frame=(c_int, [c_void_p, c_int, POINTER(Uniforms), POINTER(c_void_p), c_int, c_int,
                              c_int, c_int, c_int, c_int, c_int, c_void_p]),
    sfb_render_frame_probe=(c_int, [c_void_p, c_int, POINTER(Uniforms), POINTER(c_void_p), c_int, c_int,
                                    c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    sfb_pipe_open=(c_int, [c_void_p, c_int, c_int, c_size_t, POINTER(c_void_p)]),
    sfb_pipe_acquire=(c_int, [c_void_p, POINTER(c_void_p)]),
    sfb_pipe_submit=(c_int, [c_void_p, c_void_p]),
    sfb_pipe_set_fd=(c_int, [c_void_p, c_int]),
    sfb_pipe_sync=(c_int, [c_void_p]),
    sfb_pipe_stats=(c_int, [c_void_p, POINTER(c_uint64), POINTER(c_uint64)]),
    sfb_pipe_close=(c_int, [c_void_p]),
    sfb_sink_open=(c_int, [c_void_p, c_char_p, c_int, c_int, c_int, c_size_t, POINTER(c_void_p)]),
    sfb_sink_path=(c_char_p, [c_void_p]),
    sfb_sink_begin=(c_int, [c_void_p, c_int64, c_int, c_int]),
    sfb_sink_owner=(c_int, [c_void_p, c_int64, POINTER(c_int)]),
    sfb_sink_acquire=(c_int, [c_void_p, POINTER(c_void_p)]),
    sfb_sink_submit=(c_int, [c_void_p]),
    sfb_sink_submit_host=(c_int, [c_void_p, c_void_p]),
    sfb_sink_finish=(c_int, [c_void_p, POINTER(c_uint64), POINTER(c_uint64)]),
    sfb_sink_abort=(c_int, [c_void_p]),
    sfb_sink_close=(c_int, [c_void_p]),
    sfb_flac_info_get=(c_int, [c_void_p, c_size_t, POINTER(FlacInfo)]),
    sfb_flac_decode=(c_int, [c_void_p, c_size_t, c_void_p, c_int64, POINTER(c_int64)]),
    sfb_video_frame_bytes=(c_size_t, [c_int, c_int, c_int]),
    sfb_video_frame=(c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    sfb_jit_compile=(c_int, [c_char_p, POINTER(c_char_p), POINTER(c_char_p), c_int, c_int, POINTER(c_void_p), POINTER(c_size_t),
                             POINTER(c_void_p)]),
    sfb_jit_free=(None, [c_void_p]),
    sfb_program_load=(c_int, [c_void_p, c_void_p, c_size_t, c_int, POINTER(c_int)]),
    sfb_program_unload=(c_int, [c_void_p, c_int]),
)

_lib = None


def lib() -> C.CDLL:
    """Loads libsfb200.so once and types every export of include/sfb200.h"""
    global _lib
    if _lib is None:
        if not LIBRARY.exists():
            raise RuntimeError(
                f"{LIBRARY} is missing. Build it with `python -m shaderflow_b200.build` "
                "(needs nvcc); shaderflow_b200 has no CPU fallback.")
        handle = C.CDLL(str(LIBRARY))
        for name, (restype, argtypes) in _PROTOTYPES.items():
            fn = getattr(handle, name)      # AttributeError here = header and library disagree
            fn.restype, fn.argtypes = restype, argtypes
        _lib = handle
    return _lib


def exported_symbols() -> list[str]:
    return list(_PROTOTYPES)


def check(code: int) -> None:
    if code != OK:
        message = lib().sfb_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libsfb200: {message} (code {code})")


class CompileError(RuntimeError):
    """sfb_jit_compile rejected the source; `.log` holds the compiler's diagnostics"""
    def __init__(self, message: str, log: str):
        super().__init__(message + ("\n" + log if log else ""))
        self.log = log


def flac_decode(data) -> tuple[FlacInfo, "np.ndarray"]:
    """A FLAC stream in memory (bytes / uint8 array) → (STREAMINFO, samples int32 [frames][channels], right-justified)"""
    import numpy as np
    raw = np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray, memoryview)) else np.ascontiguousarray(data, dtype=np.uint8)
    info, frames = FlacInfo(), c_int64()
    check(lib().sfb_flac_info_get(raw.ctypes.data, raw.size, byref(info)))
    total = int(info.total_samples)
    if total == 0:                                             # not recorded: walk the frames once to count
        check(lib().sfb_flac_decode(raw.ctypes.data, raw.size, None, 0, byref(frames)))
        total = frames.value
    pcm = np.empty((total, info.channels), dtype=np.int32)
    check(lib().sfb_flac_decode(raw.ctypes.data, raw.size, pcm.ctypes.data, total, byref(frames)))
    return info, pcm[:frames.value]


def video_frame_bytes(fmt: int, width: int, height: int) -> int:
    return int(lib().sfb_video_frame_bytes(fmt, width, height))


def jit_compile(source: str, headers: dict[str, str], flags: int = 0) -> tuple[bytes, str]:
    """CUDA source + in-memory headers → (SASS image for sm_100a, compile log). Needs NVRTC, not a GPU."""
    names = (c_char_p*len(headers))(*[k.encode() for k in headers])
    texts = (c_char_p*len(headers))(*[v.encode() for v in headers.values()])
    image, size, log = c_void_p(), c_size_t(), c_void_p()
    code = lib().sfb_jit_compile(source.encode(), names, texts, len(headers), flags, byref(image), byref(size), byref(log))
    text = C.string_at(log).decode("utf-8", "replace") if log else ""
    if log:
        lib().sfb_jit_free(log)
    if code != OK:
        message = lib().sfb_last_error().decode("utf-8", "replace")
        if code == EINVAL:
            raise CompileError(f"libsfb200: {message}", text)
        raise RuntimeError(f"libsfb200: {message} (code {code})")
    data = C.string_at(image, size.value)
    lib().sfb_jit_free(image)
    return data, text


def _ptr(obj) -> c_void_p | None:
    """tensor / ndarray / int / None → void*"""
    if obj is None:
        return None
    if isinstance(obj, int):
        return c_void_p(obj)
    if hasattr(obj, "data_ptr"):
        return c_void_p(obj.data_ptr())
    if hasattr(obj, "ctypes"):
        return c_void_p(obj.ctypes.data)
    raise TypeError(f"cannot take the address of {type(obj).__name__}")


def device_count() -> int:
    n = c_int(0)
    code = lib().sfb_device_count(byref(n))
    return n.value if code == OK else 0


def frame_clock(n_frames: int, fps: float = 60.0, speed: float = 1.0, samplerate: int = 44100,
                channels: int = 2, total_samples: int = -1):
    """→ (time[f64], dt[f64], tell[i64]) numpy arrays; host only (sfb_frame_clock)"""
    import numpy as np
    time, dt, tell = np.zeros(n_frames), np.zeros(n_frames), np.zeros(n_frames, np.int64)
    check(lib().sfb_frame_clock(n_frames, fps, speed, samplerate, channels, total_samples,
                                _ptr(time), _ptr(dt), _ptr(tell)))
    return time, dt, tell


def scene_lookup(name: str) -> int:
    scene = c_int(-1)
    check(lib().sfb_scene_lookup(name.encode(), byref(scene)))
    return scene.value


def visualizer_plan(uniforms: Uniforms, background: tuple[int, int], width: int, height: int, ssaa: int) -> tuple[int, int]:
    """(fragment rows per thread of the separable kernel or 0 = tiled kernel, window rows); host-only"""
    rows, window = c_int(), c_int()
    check(lib().sfb_visualizer_plan(byref(uniforms), background[0], background[1], width, height, ssaa, byref(rows), byref(window)))
    return rows.value, window.value


def scene_info(scene: int) -> dict:
    info = SceneInfo()
    check(lib().sfb_scene_info_get(scene, byref(info)))
    return dict(
        name=info.name.decode(), reference=info.reference.decode(),
        extra=[info.extra[i].decode() for i in range(info.n_extra)],
        samplers=[info.samplers[i].decode() for i in range(info.n_samplers)],
        required=info.n_required,
    )


class Texture:
    """sfb_tex handle"""
    def __init__(self, ctx: "Context", width: int, height: int, components: int, dtype: int,
                 linear: bool = True, repeat_x: bool = True, repeat_y: bool = True):
        self.ctx, self.handle = ctx, c_void_p()
        self.width, self.height, self.components, self.dtype = width, height, components, dtype
        check(lib().sfb_tex_create(ctx.handle, width, height, components, dtype,
                                   FILTER_LINEAR if linear else FILTER_NEAREST, int(repeat_x), int(repeat_y), byref(self.handle)))

    def set_sampling(self, linear: bool, repeat_x: bool, repeat_y: bool) -> None:
        check(lib().sfb_tex_set_sampling(self.handle, FILTER_LINEAR if linear else FILTER_NEAREST, int(repeat_x), int(repeat_y)))

    def write(self, data, viewport: tuple[int, int, int, int] | None = None) -> None:
        """data: numpy array / bytes (host) or a CUDA tensor (device), tightly packed"""
        x, y, w, h = viewport or (0, 0, self.width, self.height)
        on_device = bool(getattr(data, "is_cuda", False))
        if isinstance(data, (bytes, bytearray, memoryview)):
            keep = (C.c_char*len(data)).from_buffer_copy(data)
            check(lib().sfb_tex_write(self.handle, C.cast(keep, c_void_p), 0, x, y, w, h))
            return
        check(lib().sfb_tex_write(self.handle, _ptr(data), int(on_device), x, y, w, h))

    def bind_external(self, pointer) -> None:
        check(lib().sfb_tex_bind_external(self.handle, _ptr(pointer)))

    def read(self):
        import numpy as np
        padded = 4 if self.components == 3 else self.components
        dt = {DTYPE_U8: np.uint8, DTYPE_F32: np.float32, DTYPE_F16: np.float16}[self.dtype]
        out = np.empty((self.height, self.width, padded), dt)
        check(lib().sfb_tex_read(self.handle, _ptr(out)))
        return out

    def storage(self) -> tuple[int, int]:
        """(device pointer, bytes) of the linear storage, for use as a render target"""
        p, n = c_void_p(), c_size_t()
        check(lib().sfb_tex_storage(self.handle, byref(p), byref(n)))
        return p.value, n.value

    def sample(self, uv, out, flags: int = FILTER_EXACT) -> None:
        """uv (n, 2) f32 cuda → out (n, 4) f32 cuda"""
        check(lib().sfb_tex_sample(self.handle, _ptr(uv), uv.shape[0], flags, _ptr(out)))

    def destroy(self) -> None:
        if self.handle:
            lib().sfb_tex_destroy(self.handle)
            self.handle = c_void_p()

    def __del__(self):
        try: self.destroy()
        except Exception: pass


class Pipe:
    """sfb_pipe handle: device ring + pinned ring + writer thread"""
    def __init__(self, ctx: "Context", fd: int, buffers: int, frame_bytes: int):
        self.ctx, self.handle, self.frame_bytes, self.buffers = ctx, c_void_p(), frame_bytes, buffers
        check(lib().sfb_pipe_open(ctx.handle, fd, buffers, frame_bytes, byref(self.handle)))

    def acquire(self) -> int:
        p = c_void_p()
        check(lib().sfb_pipe_acquire(self.handle, byref(p)))
        return p.value

    def submit(self, frame=None) -> None:
        check(lib().sfb_pipe_submit(self.handle, _ptr(frame)))

    def sync(self) -> None:
        check(lib().sfb_pipe_sync(self.handle))

    def set_fd(self, fd: int) -> None:
        check(lib().sfb_pipe_set_fd(self.handle, fd))

    def stats(self) -> tuple[int, int]:
        f, b = c_uint64(), c_uint64()
        check(lib().sfb_pipe_stats(self.handle, byref(f), byref(b)))
        return f.value, b.value

    def close(self) -> None:
        if self.handle:
            handle, self.handle = self.handle, c_void_p()
            check(lib().sfb_pipe_close(handle))

    def __del__(self):
        try: self.close()
        except Exception: pass


class SharedSink:
    """sfb_sink handle: this rank's share of the sharded export's host ring (csrc/sink.cu). `ctx` None = host-only"""
    def __init__(self, ctx: "Context | None", path: str | None, rank: int, world: int, slots: int, frame_bytes: int):
        self.ctx, self.handle = ctx, c_void_p()
        self.rank, self.world, self.slots, self.frame_bytes = rank, world, slots, frame_bytes
        check(lib().sfb_sink_open(ctx.handle if ctx is not None else None, path.encode() if path else None,
                                  rank, world, slots, frame_bytes, byref(self.handle)))

    @property
    def path(self) -> str:
        return lib().sfb_sink_path(self.handle).decode()

    def begin(self, total_frames: int, block: int, fd: int = -1) -> None:
        check(lib().sfb_sink_begin(self.handle, total_frames, block, fd))

    def acquire(self) -> int:
        p = c_void_p()
        check(lib().sfb_sink_acquire(self.handle, byref(p)))
        return p.value

    def submit(self) -> None:
        check(lib().sfb_sink_submit(self.handle))

    def submit_host(self, frame=None) -> None:
        check(lib().sfb_sink_submit_host(self.handle, _ptr(frame)))

    def finish(self) -> tuple[int, int]:
        f, b = c_uint64(), c_uint64()
        check(lib().sfb_sink_finish(self.handle, byref(f), byref(b)))
        return f.value, b.value

    def abort(self) -> None:
        if self.handle:
            lib().sfb_sink_abort(self.handle)

    def close(self) -> None:
        if self.handle:
            handle, self.handle = self.handle, c_void_p()
            check(lib().sfb_sink_close(handle))

    def __del__(self):
        try: self.close()
        except Exception: pass


class Context:
    """sfb_ctx handle. `stream` defaults to torch's current stream on `device`"""
    def __init__(self, device: int = 0, stream: int | None = None):
        self.handle = c_void_p()
        self.device = device
        if stream is None:
            import torch
            if not torch.cuda.is_available():
                raise RuntimeError("shaderflow_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
            torch.cuda.set_device(device)
            stream = torch.cuda.current_stream(device).cuda_stream
        check(lib().sfb_ctx_create(device, c_void_p(stream), byref(self.handle)))

    @property
    def torch_device(self) -> str:
        """Where tensors handed to this context live"""
        return f"cuda:{self.device}"

    def set_stream(self, stream: int) -> None:
        check(lib().sfb_ctx_set_stream(self.handle, c_void_p(stream)))

    def sync(self) -> None:
        check(lib().sfb_sync(self.handle))

    def enable_peer(self, peer_device: int) -> None:
        check(lib().sfb_ctx_enable_peer(self.handle, peer_device))

    @property
    def launches(self) -> int:
        n = c_uint64()
        check(lib().sfb_launch_count(self.handle, byref(n)))
        return n.value

    # -- audio ------------------------------------------------------------------------------ #

    def stft_mel(self, pcm, tell, fft_n: int, csr=None, *, window: int = 0, magnitude: int = 0,
                 volume: int = 0, mag_out=None, spec_out=None) -> None:
        """pcm (ch, n) f32 cuda; tell (F,) i64 cuda; csr = (indptr, indices, data, bins) cuda tensors"""
        channels, n = pcm.shape
        indptr, indices, data, bins = csr if csr is not None else (None, None, None, 0)
        check(lib().sfb_stft_mel(self.handle, _ptr(pcm), n, channels, fft_n, _ptr(tell), tell.shape[0],
            window, magnitude, _ptr(indptr), _ptr(indices), _ptr(data), bins, volume, _ptr(mag_out), _ptr(spec_out)))

    def audio_track(self, pcm, samplerate: int, tell, dt, *, spec=None, bins: int = 0,
                    dynamics: tuple[float, float, float, float] = (4.0, 1.0, 0.0, 1e-6),
                    scalars=None, wave=None, wave_points: int = 180, wave_chunk: int = 735, wave_reducer: int = 0) -> None:
        channels, n = pcm.shape
        prm = DynamicsParams(*dynamics)
        check(lib().sfb_audio_track(self.handle, _ptr(pcm), n, channels, samplerate, _ptr(tell), _ptr(dt), tell.shape[0],
            _ptr(spec), bins, byref(prm), _ptr(scalars), _ptr(wave), wave_points, wave_chunk, wave_reducer))

    def dynamics_scan(self, values, lanes: int, dt, n_frames: int, dynamics: tuple[float, float, float, float]) -> None:
        """DynamicNumber.next over frames, in place on values [n_frames, lanes] float32 cuda"""
        prm = DynamicsParams(*dynamics)
        check(lib().sfb_dynamics_scan(self.handle, _ptr(values), lanes, _ptr(dt), n_frames, byref(prm)))

    def piano_track(self, notes, offsets, time, n_frames: int, time_offset: float, roll_time: float, lookup_time: float,
                    release: float, gmin: int, gmax: int, keys, channel, upcoming) -> None:
        check(lib().sfb_piano_track(self.handle, _ptr(notes), _ptr(offsets), _ptr(time), n_frames, time_offset, roll_time,
                                    lookup_time, release, gmin, gmax, _ptr(keys), _ptr(channel), _ptr(upcoming)))

    def piano_roll(self, notes, offsets, time: float, roll_time: float, lookup_time: float, gmin: int, gmax: int,
                   roll, overflow) -> None:
        check(lib().sfb_piano_roll(self.handle, _ptr(notes), _ptr(offsets), time, roll_time, lookup_time, gmin, gmax,
                                   _ptr(roll), _ptr(overflow)))

    # -- render ----------------------------------------------------------------------------- #

    @staticmethod
    def _samplers(textures):
        arr = (c_void_p*max(1, len(textures)))(*[t.handle for t in textures])
        return arr, len(textures)

    def render_screen(self, scene: int, uniforms: Uniforms, textures, width: int, height: int, dst,
                      dst_f32=None, flags: int = FILTER_EXACT) -> None:
        arr, n = self._samplers(textures)
        check(lib().sfb_render_screen(self.handle, scene, byref(uniforms), arr, n, flags, width, height, _ptr(dst), _ptr(dst_f32)))

    def render_target(self, scene: int, uniforms: Uniforms, textures, target: "Texture", flags: int = FILTER_EXACT) -> None:
        """The scene's pass into `target`'s own storage, in its format (a child program / layer / temporal slot)"""
        arr, n = self._samplers(textures)
        check(lib().sfb_render_target(self.handle, scene, byref(uniforms), arr, n, flags, target.handle))

    def render_final(self, screen, screen_w: int, screen_h: int, width: int, height: int,
                     subsample: int, components: int, dst) -> None:
        check(lib().sfb_render_final(self.handle, _ptr(screen), screen_w, screen_h, width, height, subsample, components, _ptr(dst)))

    def render_frame(self, scene: int, uniforms: Uniforms, textures, width: int, height: int,
                     ssaa: int, subsample: int, components: int, dst, flags: int = FILTER_EXACT) -> None:
        arr, n = self._samplers(textures)
        check(lib().sfb_render_frame(self.handle, scene, byref(uniforms), arr, n, flags, width, height, ssaa, subsample, components, _ptr(dst)))

    def render_frame_probe(self, scene: int, uniforms: Uniforms, textures, width: int, height: int,
                           ssaa: int, subsample: int, components: int, dst, screen_f32, flags: int = FILTER_EXACT) -> None:
        """render_frame that also stores every sub-sample's pre-store fragColor as float4 (parity tests)"""
        arr, n = self._samplers(textures)
        check(lib().sfb_render_frame_probe(self.handle, scene, byref(uniforms), arr, n, flags, width, height, ssaa, subsample,
                                           components, _ptr(dst), _ptr(screen_f32)))

    def pcm_ingest(self, raw, n_frames: int, channels: int, fmt: int, planar, clip_samples: int, offset: int) -> None:
        """Interleaved file samples (device bytes) → rows [offset, offset+n_frames) of the planar float32 clip"""
        check(lib().sfb_pcm_ingest(self.handle, _ptr(raw), n_frames, channels, fmt, _ptr(planar), clip_samples, offset))

    def video_frame(self, frame, fmt: int, width: int, height: int, top_down: bool, texture: "Texture") -> None:
        """One decoded frame (device bytes in the file's layout) → the texture's storage, flipped / converted on the GPU"""
        check(lib().sfb_video_frame(self.handle, _ptr(frame), fmt, width, height, int(top_down), texture.handle))

    def program_load(self, image: bytes, n_samplers: int) -> int:
        """SASS image of a run-time compiled program → scene id (>= SCENE_PROGRAM_BASE) for the render_* calls"""
        scene = c_int()
        check(lib().sfb_program_load(self.handle, image, len(image), n_samplers, byref(scene)))
        return scene.value

    def program_unload(self, scene: int) -> None:
        check(lib().sfb_program_unload(self.handle, scene))

    def destroy(self) -> None:
        if self.handle:
            lib().sfb_ctx_destroy(self.handle)
            self.handle = c_void_p()

    def __del__(self):
        try: self.destroy()
        except Exception: pass

"""ShaderScene — owns the modules, the time base and the per-frame driver.
API mirror of shaderflow/scene.py for the offline path: `initialize()`, `main()` (export / freewheel
branch, scene.py:493-639), `next()` (scene.py:456-479), `pipeline()` (scene.py:687-703), time and
resolution properties. Window creation, input events and imgui are out of scope: the CUDA backend has no
window, so `main()` always runs freewheel and `realtime` is always False.

Differences a user can see, all additive:
  * `ShaderScene(device=i)` picks the GPU; `backend="dry"` builds the module graph without a GPU
    (host-logic tests) and refuses to render;
  * `main(output=...)` also accepts "null" (render + device→host copy, frames dropped) and "*.rgb";
  * `main(frames=(a, b))` renders only frames [a, b) while stepping every frame's host state — the
    building block of the frame-sharded multi-GPU export (distributed.py)."""
from __future__ import annotations

import math
from pathlib import Path
from types import SimpleNamespace
from typing import Any, Iterable, Optional, Union

import numpy as np
from attrs import Factory, define, field

from shaderflow_b200 import _native as N
from shaderflow_b200 import logger
from shaderflow_b200.camera import ShaderCamera
from shaderflow_b200.exporting import ExportingHelper
from shaderflow_b200.frametimer import ShaderFrametimer
from shaderflow_b200.keyboard import ShaderKeyboard
from shaderflow_b200.message import ShaderMessage
from shaderflow_b200.module import ShaderModule
from shaderflow_b200.resolution import Resolution
from shaderflow_b200.scheduler import Scheduler
from shaderflow_b200.shader import ShaderProgram
from shaderflow_b200.variable import ShaderVariable, Uniform


class _Cli:
    """Stands in for the cyclopts App modules register commands on (scene.py:200-215)"""
    help = None
    def command(self, *a, **k):
        return (lambda fn: fn) if not a else a[0]
    def __call__(self, *a, **k):
        raise NotImplementedError("use ShaderScene.main(...) directly; the CLI launcher is out of scope")


@define
class ShaderScene(ShaderModule):
    backend: str = "cuda"
    """"cuda" or "dry" (no GPU objects; for host-logic tests)"""
    device: Optional[int] = None
    """CUDA device index; None = LOCAL_RANK or 0"""
    cuda: Optional[N.Context] = None
    """libsfb200 context (what `opengl` was to the reference)"""
    window: Any = None
    quality: float = field(default=50.0, converter=float)
    modules: list = Factory(list)
    ffmpeg: Any = None
    frametimer: ShaderFrametimer = None
    keyboard: ShaderKeyboard = None
    camera: ShaderCamera = None
    shader: ShaderProgram = None
    _final: ShaderProgram = None
    subsample: int = field(default=2, converter=lambda x: int(max(1, x)))
    cli: Any = Factory(_Cli)
    fuse: bool = True
    """Use the fused K3+K4 kernel whenever final.glsl degenerates to a box filter"""
    _sink_ring: Any = None
    _peer_frames: Any = None
    """Sharded export: rank 0's staging of the remote frames, mapped into every rank (distributed.PeerFrames)"""
    _peer_frames_key: Any = None
    _shared_sink: Any = None
    """Sharded export with a sink: this rank's ring of the shared host segment (distributed.negotiate_shared_sink)"""
    _shared_sink_key: Any = None
    _updaters: Any = None
    """(len(modules), modules whose update() runs every frame): rebuilt when a module is added"""
    """The frame sink's ring (pinned buffers, copy stream, writer thread), kept across main() calls"""
    kernel_events: Any = None
    """When a list: (start, end) torch CUDA events are appended around every shading launch (bench.py)"""

    # -- lifecycle -------------------------------------------------------------------------------
    def __attrs_post_init__(self) -> None:
        ShaderModule.__attrs_post_init__(self)
        self.name = (self.name or type(self).__name__)

    def __del__(self):
        for module in list(getattr(self, "modules", ())):
            if module is not self:
                try: module.destroy()
                except Exception: pass
        self._frame_buffer = None
        for attr in ("_sink_ring", "_shared_sink"):
            ring = getattr(self, attr, None)
            if ring is not None:
                try: setattr(self, attr, None); ring.close()
                except Exception: pass
        self._peer_frames = None

    def initialize(self) -> None:
        if self.shader is not None:
            return
        logger.info(f"Initializing scene {self.name} with backend {self.backend}")
        if self.backend == "cuda":
            import os
            device = self.device if self.device is not None else int(os.environ.get("LOCAL_RANK", 0))
            self.cuda = N.Context(device)          # raises without a GPU: no CPU fallback
            self.device = device
        self.frametimer = ShaderFrametimer(scene=self)
        self.keyboard = ShaderKeyboard(scene=self)
        self.camera = ShaderCamera(scene=self)
        # SSAA downsampler + main program (scene.py:185-194)
        self._final = ShaderProgram(scene=self, name="iFinal")
        self._final.texture.components = 3
        self._final.texture.dtype = np.uint8
        self._final.texture.final = True
        self._final.texture.track = 1.0
        self.shader = ShaderProgram(scene=self, name="iScreen")
        self.shader.texture.repeat(False)
        self.shader.texture.track = 1.0
        self.build()

    # -- time ------------------------------------------------------------------------------------
    time: float = field(default=0.0, converter=float)
    speed: float = field(default=1.0, converter=float)
    runtime: float = field(default=10.0, converter=float)
    fps: float = field(default=60.0, converter=float)
    dt: float = field(default=0.0, converter=float)
    rdt: float = field(default=0.0, converter=float)
    frame_index: int = 0
    """Frames stepped since main() started: the row GPU-track modules publish this frame"""

    @property
    def tau(self) -> float:
        return (self.time/self.runtime) % 1.0

    @property
    def cycle(self) -> float:
        return self.tau*math.tau

    @property
    def frametime(self) -> float:
        return 1.0/self.fps

    @frametime.setter
    def frametime(self, value: float):
        self.fps = 1.0/value

    @property
    def frame(self) -> int:
        return round(self.time*self.fps)

    @frame.setter
    def frame(self, value: int):
        self.time = value/self.fps

    @property
    def duration(self) -> float:
        return self.runtime

    @property
    def max_duration(self) -> float:
        return max((module.duration or 0.0) for module in self.modules)

    def set_duration(self, override: Optional[float] = None) -> float:
        self.runtime = (override or self.max_duration or self.runtime)
        self.runtime /= self.speed
        return self.runtime

    @property
    def total_frames(self) -> int:
        return max(1, round(self.runtime*self.fps))

    # -- window-ish attributes (accepted, inert) -----------------------------------------------------
    title: str = "ShaderFlow"
    resizable: bool = True
    fullscreen: bool = False
    exclusive: bool = False
    visible: bool = False
    mouse_gluv: tuple = (0.0, 0.0)
    mouse_inside: bool = False
    mouse_buttons: dict = Factory(lambda: {1: False, 2: False, 3: False})

    # -- resolution ------------------------------------------------------------------------------
    _width: int = field(default=1920)
    _height: int = field(default=1080)
    _ssaa: float = field(default=1.0, converter=lambda x: max(0.01, float(x)))
    _aspect_ratio: Optional[float] = None

    @property
    def width(self) -> int:
        return self._width

    @width.setter
    def width(self, value: int):
        self.resize(width=value)

    @property
    def height(self) -> int:
        return self._height

    @height.setter
    def height(self, value: int):
        self.resize(height=value)

    @property
    def ssaa(self) -> float:
        """Render at this multiple of the resolution, then downsample (O(N²) shading cost)"""
        return self._ssaa

    @ssaa.setter
    def ssaa(self, value: float):
        self._ssaa = max(0.01, float(value))
        self.relay(ShaderMessage.Shader.RecreateTextures)

    @property
    def resolution(self) -> tuple[int, int]:
        return (self.width, self.height)

    @resolution.setter
    def resolution(self, value: tuple[int, int]):
        self.resize(*value)

    @property
    def render_resolution(self) -> tuple[int, int]:
        return (int(self.width*self.ssaa), int(self.height*self.ssaa))

    @property
    def aspect_ratio(self) -> float:
        return self._aspect_ratio or (self.width/self.height)

    @aspect_ratio.setter
    def aspect_ratio(self, value: Optional[Union[float, str]]):
        if isinstance(value, str):
            value = eval(value.replace(":", "/").capitalize())
        self._aspect_ratio = value

    def resize(self, width: Optional[int] = None, height: Optional[int] = None, ratio=None,
               bounds: Optional[tuple] = None, ssaa: Optional[float] = None, scale: float = 1.0) -> tuple[int, int]:
        self.aspect_ratio = (ratio or self._aspect_ratio)
        self._ssaa = (ssaa or self._ssaa)
        resolution = Resolution.fit(old=(self._width, self._height), new=(width, height), max=bounds,
                                    ar=self._aspect_ratio, scale=scale)
        if resolution != (self.width, self.height):
            self._width, self._height = resolution
            self.relay(ShaderMessage.Shader.RecreateTextures)
            logger.info(f"Resized to {self.resolution}")
        return self.resolution

    # -- frame target ----------------------------------------------------------------------------
    _frame_buffer: Any = None
    _frame_target: Optional[int] = None

    @property
    def frame_tensor(self):
        """The scene's own [H, W, 3] uint8 device frame (bottom row first), used when no sink ring is"""
        import torch
        shape = (self.height, self.width, 3)
        if self._frame_buffer is None or tuple(self._frame_buffer.shape) != shape:
            self._frame_buffer = torch.zeros(shape, dtype=torch.uint8, device=f"cuda:{self.device}")
        return self._frame_buffer

    @property
    def frame_pointer(self) -> int:
        return self._frame_target if self._frame_target is not None else self.frame_tensor.data_ptr()

    def screenshot(self) -> np.ndarray:
        """Last rendered frame, top row first (scene.py:439-443)"""
        if self._frame_target is not None:
            raise RuntimeError("screenshot() is not available while a sink ring owns the frame")
        self.cuda.sync()
        return np.flipud(self.frame_tensor.cpu().numpy())

    # -- main loop -------------------------------------------------------------------------------
    scheduler: Scheduler = Factory(Scheduler)
    vsync: Any = None
    quit: bool = False
    realtime: bool = False
    exporting: bool = False
    freewheel: bool = False
    headless: bool = True
    render_enabled: bool = True

    @property
    def fusable(self) -> Optional[int]:
        """Integer ssaa when one fused launch can replace iScreen + iFinal (sfb_render_frame's rule)"""
        programs = [m for m in self.modules if isinstance(m, ShaderProgram)]
        if not self.fuse or len(programs) != 2:
            return None
        tex = self.shader.texture
        if tex.temporal != 1 or tex.layers != 1 or float(self.ssaa) != int(self.ssaa) or int(self.ssaa) < 1:
            return None
        s = int(self.ssaa)
        return s if (self.subsample == s or 2*self.subsample == s) else None

    def render(self) -> None:
        """Every program, children first, iFinal last (scene.py:469-471)"""
        s = self.fusable
        if s is not None:
            self.shader.render_fused(self.frame_pointer, s)
            return
        for module in reversed(self.modules):
            if isinstance(module, ShaderProgram):
                module.update()

    def next(self, dt: float = 0.0) -> None:
        """Update all modules, render, then integrate time (so frame 0 sees time=0, dt=0)"""
        updaters = self._updaters
        if updaters is None or updaters[0] != len(self.modules):
            updaters = self._updaters = (len(self.modules), [m for m in self.modules if not isinstance(m, ShaderProgram)])
        for module in updaters[1]:
            module.update()
        if self.render_enabled and self.cuda is not None:
            self.render()
        if self.vsync is not None:
            self.vsync.fps = self.fps
        # plain stores: these four are written once per frame (the attrs converters are for user input)
        store = object.__setattr__
        store(self, "dt", float(dt*self.speed))
        store(self, "rdt", float(dt))
        store(self, "time", self.time + self.dt)
        store(self, "frame_index", self.frame_index + 1)

    def main(self, *, width: Optional[int] = 1920, height: Optional[int] = 1080, scale: float = 1.0,
             ratio=None, fps: float = 60.0, frameskip: bool = True, fullscreen: bool = False,
             quality: float = 50.0, ssaa: float = 1.0, subsample: int = 2,
             output: Optional[Union[Path, str]] = None, time: Optional[float] = None, speed: float = 1.0,
             freewheel: bool = False, raw: bool = False, turbo: bool = True, buffers: int = 5,
             frames: Optional[tuple[int, int]] = None, on_frame=None, distributed: Optional[bool] = None):
        """Renders the scene offline. Same flags as the reference's main() (scene.py:493-561); returns
        the output path / bytes, or None without `output`. `frames=(a, b)` renders only that range;
        `on_frame(index, device_pointer)` is called after each rendered frame when no sink is open.
        Under torchrun (`torch.distributed` initialised, world > 1; or distributed=True) the export is
        frame-sharded (distributed.py): with a sink, block-cyclic frames drained by every rank over its own
        PCIe link into one shared host ring that rank 0's writer streams in time order; without one, contiguous
        ranges reassembled in rank 0's HBM for `on_frame`. Ranks other than 0 return None. Scenes with GPU
        feedback (a program texture with temporal > 1: MotionBlur, Life) do not shard — rank 0 exports alone."""
        from shaderflow_b200 import distributed as D
        rank, world = 0, 1
        if distributed is not False:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                rank, world = dist.get_rank(), dist.get_world_size()
            elif distributed:
                rank, world = D.init_process_group()
        self.initialize()
        self.exporting = bool(output)
        self.freewheel = True
        self.headless, self.realtime = True, False
        self.subsample, self.quality, self.speed, self.fps = subsample, quality, speed, fps
        self.time, self.dt, self.rdt, self.frame_index, self.quit = 0, 0.0, 0.0, 0, False
        self.relay(ShaderMessage.Shader.Compile)
        self.scheduler.clear()
        _width, _height = self.resize(width=width, height=height, ratio=ratio, scale=scale)
        for module in self.modules:
            module.setup()
        self.set_duration(eval(time) if isinstance(time, str) else time)
        if raw or ssaa < 1:                      # scene.py:591-592: pipe native frames
            self.resize(*[int(x*ssaa) for x in self.resolution], scale=1, ssaa=1)
        else:
            self.ssaa = ssaa
        for module in self.modules:              # programs picked their kernels; sizes are final
            if isinstance(module, ShaderProgram) and module.scene_id is None:
                module.compile()

        # frame k of a feedback scene reads frame k-1's pixels: every frame before it must have been shaded
        feedback = any(isinstance(m, ShaderProgram) and m.texture.temporal > 1 for m in self.modules)
        if world > 1 and feedback:
            logger.info(f"Scene {self.name} has GPU feedback (temporal textures): it does not frame-shard, rank 0 exports alone")
            if rank != 0:
                return None
            world = 1
        export = ExportingHelper(self)
        export.make_buffers(buffers)
        run = SimpleNamespace(export=export, rank=rank, world=world, on_frame=on_frame, turbo=turbo, frameskip=frameskip,
                              feedback=feedback, frames=frames, output=output, buffers=buffers)
        try:
            if world == 1:
                self._export_single(run)
            elif self.exporting:
                self._export_sharded_sink(run, D)
            else:
                self._export_sharded_hbm(run, D)
        except BaseException:
            # leave nothing half open: the scene must be able to export again (and ffmpeg must not linger)
            export.abort()
            raise
        finally:
            self._frame_target, self.render_enabled = None, True
        result = export.result()
        export.log_stats(output=result)
        return result

    def _frame_loop(self, export, last: int, prepare, shaded, frameskip: bool = True) -> None:
        """Steps frames 0 .. last-1 on the freewheel clock; `prepare(index)` decides before a frame whether it is
        shaded and where to, `shaded(index)` hands a shaded frame on"""
        self.vsync = self.scheduler.new(task=self.next, frequency=self.fps, freewheel=True,
                                        frameskip=frameskip, precise=True)
        prepare(0)
        while (task := self.scheduler.next()):           # → self.next(dt): one frame
            if task is not self.vsync:
                continue
            if self.render_enabled:
                shaded(export.frame)
            export.update()
            if self.quit or export.finished or export.frame >= last:
                break
            prepare(export.frame)

    def _export_single(self, run) -> None:
        export = run.export
        if self.exporting:
            export.ffmpeg_output(run.output)
            export.popen()
        export.open_bar()
        first, last = run.frames if run.frames is not None else (0, export.total_frames)
        shade_from = 0 if run.feedback else first            # feedback scenes need their history

        def prepare(index: int) -> None:
            self.render_enabled = (shade_from <= index < last)
            sunk = self.exporting and first <= index < last
            self._frame_target = export.target() if sunk else None

        def shaded(index: int) -> None:
            if not (first <= index < last):
                return
            if self.exporting:
                export.pipe(turbo=run.turbo)
            elif run.on_frame is not None:
                run.on_frame(index, self.frame_pointer)

        self._frame_loop(export, last, prepare, shaded, run.frameskip)
        export.finish()

    def _export_sharded_sink(self, run, D) -> None:
        """Sink-bound sharded export: block-cyclic ownership, every rank drains its own frames (csrc/sink.cu)"""
        import os
        import torch.distributed as dist
        export, rank, world = run.export, run.rank, run.world
        block = max(1, int(os.environ.get("SFB_SHARD_BLOCK", "4")))
        slots = max(2*block, int(run.buffers), 8)
        key = (export.frame_bytes, world, slots)
        if self._shared_sink_key != key:
            if self._shared_sink is not None:
                self._shared_sink.close()
            self._shared_sink = D.negotiate_shared_sink(self.cuda, export.frame_bytes, rank, world, slots, device=f"cuda:{self.device}")
            self._shared_sink_key = key
        sink = self._shared_sink
        if sink is None:                                      # no shared segment on this box: stage in rank 0's HBM
            return self._export_sharded_hbm(run, D)
        fd = -1
        if rank == 0:
            export.ffmpeg_output(run.output)
            fd = export.open_fd()
            export.open_bar()
        total = export.total_frames
        try:
            sink.begin(total, block, fd)
            dist.barrier()                                    # rank 0 has reset the counters: frames may be produced

            def prepare(index: int) -> None:
                self.render_enabled = D.block_owner(index, block, world) == rank
                self._frame_target = sink.acquire() if self.render_enabled else None

            def shaded(index: int) -> None:
                if rank == 0:
                    export.check_process()
                sink.submit()

            self._frame_loop(export, total, prepare, shaded)
            export.frame = total
            sink.finish()                                     # this rank's frames have left (rank 0: all written)
        except BaseException:
            sink.abort()                                      # unblocks the other ranks with an error
            self._shared_sink, self._shared_sink_key = None, None
            sink.close()
            raise
        self.cuda.sync()
        dist.barrier()                                        # the counters may be reset for the next export only now
        export.finish()

    def _export_sharded_hbm(self, run, D) -> None:
        """Sharded export whose frames stay in HBM (or whose box has no shared segment): contiguous ranges,
        reassembled in rank 0's HBM — peer stores over NVLink from the shading kernel (distributed.PeerFrames),
        else NCCL point-to-point (distributed.FrameGather) — and consumed there in time order"""
        import torch
        import torch.distributed as dist
        export, rank, world, on_frame = run.export, run.rank, run.world, run.on_frame
        if self.exporting and rank == 0:
            export.ffmpeg_output(run.output)
            export.popen()
        if rank == 0:
            export.open_bar()
        first, last = D.shard_range(export.total_frames, rank, world)
        staging, gather, sent = None, None, 0
        key = (export.total_frames, self.height, self.width, world)
        if self._peer_frames_key != key:
            self._peer_frames = None                          # frees the previous staging before the next is made
            self._peer_frames = D.PeerFrames.negotiate(export.total_frames, (self.height, self.width, 3), rank, world, self.device,
                                                       enable_peer=self.cuda.enable_peer)
            self._peer_frames_key = key
        peer = self._peer_frames
        if peer is None:
            gather = D.FrameGather(export.total_frames, rank, world, chunk=4)
            if rank == 0:
                gather.post((self.height, self.width, 3), torch.uint8, f"cuda:{self.device}")
            else:
                staging = torch.empty((last - first, self.height, self.width, 3), dtype=torch.uint8, device=f"cuda:{self.device}")

        def prepare(index: int) -> None:
            self.render_enabled = (first <= index < last)
            if peer is not None and rank != 0:
                self._frame_target = peer.slot(index).data_ptr() if self.render_enabled else None
            elif staging is not None:
                self._frame_target = staging[index - first].data_ptr() if self.render_enabled else None
            else:
                self._frame_target = export.target() if (self.exporting and self.render_enabled) else None

        def shaded(index: int) -> None:
            nonlocal sent
            if peer is not None and rank != 0:
                return                                        # the kernel stored the frame into rank 0's HBM
            if staging is not None:
                done = index - first + 1                      # frames of the shard shaded so far
                if done - sent >= gather.chunk or index == last - 1:
                    gather.send_block(staging[sent:done])
                    sent = done
            elif self.exporting:
                export.pipe(turbo=run.turbo)
            elif on_frame is not None:
                on_frame(index, self.frame_pointer)

        self._frame_loop(export, last, prepare, shaded)
        # the one exchange step of the path: finished frames → rank 0, in time order
        if peer is not None:
            self.cuda.sync()                                  # this rank's peer writes have landed
            dist.barrier()                                    # ... and so have everybody's
        if rank != 0:
            if gather is not None:
                gather.finish()
        else:
            index = last
            blocks = gather.drain() if gather is not None else (peer.frames[a:a + 16] for a in range(0, peer.frames.shape[0], 16))
            for block in blocks:
                for frame in block:
                    if export.pipe_handle is not None:
                        export.pipe_handle.submit(frame.data_ptr())     # D2D into the ring, then D2H + write
                    elif on_frame is not None:
                        on_frame(index, frame.data_ptr())
                    index += 1
        export.frame = export.total_frames
        self.cuda.sync()
        if peer is not None:
            dist.barrier()                                    # the staging may be written again only now
        export.finish()

    # -- module hooks ------------------------------------------------------------------------------
    def handle(self, message) -> None:
        if isinstance(message, ShaderMessage.Window.Close):
            self.quit = True
        elif isinstance(message, (ShaderMessage.Mouse.Drag, ShaderMessage.Mouse.Position)):
            self.mouse_gluv = (message.u, message.v)

    def pipeline(self) -> Iterable[ShaderVariable]:
        yield Uniform("int",   "iLayer",       None)
        yield Uniform("float", "iTime",        self.time)
        yield Uniform("float", "iTau",         self.tau)
        yield Uniform("float", "iDuration",    self.duration)
        yield Uniform("float", "iDeltatime",   self.dt)
        yield Uniform("vec2",  "iResolution",  self.resolution)
        yield Uniform("float", "iWantAspect",  self.aspect_ratio)
        yield Uniform("float", "iQuality",     self.quality/100)
        yield Uniform("float", "iSSAA",        self.ssaa)
        yield Uniform("float", "iFramerate",   self.fps)
        yield Uniform("int",   "iFrame",       self.frame)
        yield Uniform("bool",  "iRealtime",    self.realtime)
        yield Uniform("vec2",  "iMouse",       self.mouse_gluv)
        yield Uniform("bool",  "iMouseInside", self.mouse_inside)
        for i in range(1, 3):
            yield Uniform("bool", f"iMouse{i}", self.mouse_buttons[i])

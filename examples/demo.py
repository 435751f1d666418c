"""Example scenes for the CUDA backend — the user-side code of the reference's examples/basic/demo.py
and examples/fractals/fractals.py (same class names, same `build()` logic against the same API), with
two practical differences: assets the reference downloads are replaced by synthetic ones unless a path
is given, and audio is injected with `scene.audio.load(pcm, samplerate)` (the reference's examples ship
placeholder paths, demo.py:164,177,196).

    import shaderflow_b200; shaderflow_b200.install_alias()
    from examples.demo import Visualizer
    scene = Visualizer(); scene.initialize(); scene.audio.load(pcm, 44100)
    scene.main(width=3840, height=2160, ssaa=2, output="out.mp4")
"""
from pathlib import Path

import numpy as np

import shaderflow_b200
shaderflow_b200.install_alias()

from shaderflow.scene import ShaderScene          # noqa: E402
from shaderflow.shader import ShaderProgram       # noqa: E402
from shaderflow.texture import ShaderTexture      # noqa: E402
from shaderflow.variable import Uniform           # noqa: E402

shaders: Path = (Path(__file__).parent/"shaders")


def synthetic_background(width: int = 1920, height: int = 1080, seed: int = 1) -> np.ndarray:
    """Deterministic stand-in for the downloaded wallpaper (RGB8, top row first)"""
    from shaderflow_b200 import synthetic
    return synthetic.background(width, height, seed)


class Basic(ShaderScene):
    """Simplest ShaderScene"""
    ...


class ShaderToy(ShaderScene):
    """ShaderToy Default Shader"""
    def build(self):
        self.shader.fragment = (shaders/"shadertoy.frag")


class MultiShader(ShaderScene):
    """Basic scene with two shaders acting together"""
    def build(self):
        self.child = ShaderProgram(scene=self, name="child")
        # Left screen is green, right screen is black (inline GLSL of the reference, selected by directive here)
        self.child.fragment = ("// sfb200: scene=multishader_child")
        # Left screen is black, right screen is red; adds the content of the child shader
        self.shader.fragment = ("// sfb200: scene=multishader")


class Multipass(ShaderScene):
    """Multi layers done on a single shader"""
    background = None

    def build(self):
        image = self.background if self.background is not None else synthetic_background()
        ShaderTexture(scene=self, name="background").from_image(image)
        self.shader.texture.layers = 2
        self.shader.fragment = (shaders/"multipass.frag")


class MotionBlur(ShaderScene):
    """Poor's man Motion Blur"""
    background = None

    def build(self):
        image = self.background if self.background is not None else synthetic_background()
        ShaderTexture(scene=self, name="background").from_image(image)
        self.shader.texture.temporal = 10
        self.shader.texture.layers = 2
        self.shader.fragment = (shaders/"motionblur.frag")


class Dynamics(ShaderScene):
    """Second order system"""
    background = None

    def build(self):
        from shaderflow.dynamics import ShaderDynamics
        image = self.background if self.background is not None else synthetic_background()
        ShaderTexture(scene=self, name="background").from_image(image)
        self.dynamics = ShaderDynamics(scene=self, name="iShaderDynamics", frequency=4)
        self.shader.fragment = ("// sfb200: scene=dynamics")

    def update(self):
        import math
        # This is how square waves are born in the digital world
        self.dynamics.target = 0.5*(1 + np.sign(np.sin(2*math.pi*self.time*0.5)))


class Video(ShaderScene):
    """Video as a Texture demo. The reference plays a downloaded clip (examples/basic/demo.py:133-139); here `path` is any
    .y4m / raw rgb file (or what ffmpeg decodes), by default a synthetic Y4M clip written to the user data directory."""
    path = None

    def build(self):
        from shaderflow.video import ShaderVideo
        path = self.path
        if path is None:
            from shaderflow_b200 import synthetic
            path = shaderflow_b200.directories.user_data_path/"synthetic.y4m"
            if not path.exists():
                path.parent.mkdir(parents=True, exist_ok=True)
                synthetic.write_y4m(path, synthetic.video_frames(640, 360, 60), fps=30)
        self.video = ShaderVideo(scene=self, path=path)
        self.shader.fragment = (shaders/"video.frag")


class Audio(ShaderScene):
    """Basic audio processing (the reference opens a soundcard recorder; here a clip is loaded)"""
    def build(self):
        from shaderflow.audio import ShaderAudio
        self.audio = ShaderAudio(scene=self, name="iAudio")
        self.shader.fragment = ("// sfb200: scene=audio")


class Life(ShaderScene):
    """Conway's Game of Life in GLSL"""
    life_period: int = 6
    """Number of frames between each life update"""
    life_seed = None
    """Seed of the initial state (the reference draws it from numpy's global generator)"""

    def setup(self):
        width, height = 192, 108
        rng = np.random.default_rng(self.life_seed) if self.life_seed is not None else np.random
        random = (rng.integers(0, 2, (width, height)) if self.life_seed is not None
                  else rng.randint(0, 2, (width, height))).astype(bool)
        self.simulation.texture.size = (width, height)
        self.simulation.texture.write(random.astype(np.float32), temporal=1)

    def build(self):
        self.simulation = ShaderProgram(scene=self, name="iLife")
        self.simulation.texture.temporal = 10
        self.simulation.texture.filter = "nearest"
        self.simulation.texture.dtype = "f4"
        self.simulation.texture.components = 1
        self.simulation.texture.track = False
        self.simulation.fragment = (shaders/"life"/"simulation.glsl")
        self.shader.fragment = (shaders/"life"/"visuals.glsl")

    def pipeline(self):
        yield from ShaderScene.pipeline(self)
        yield Uniform("int", "iLifePeriod", self.life_period)


class Waveform(ShaderScene):
    """Audio Waveform Oscilloscope demo"""
    def build(self):
        from shaderflow.audio import ShaderAudio
        from shaderflow.audio.waveform import ShaderWaveform
        self.audio = ShaderAudio(scene=self, name="iAudio", file="/path/to/audio.ogg")
        self.waveform = ShaderWaveform(scene=self, audio=self.audio, smooth=False)
        self.shader.fragment = (shaders/"waveform.frag")


class MusicBars(ShaderScene):
    """Basic music bars"""
    def build(self):
        from shaderflow.audio import ShaderAudio
        from shaderflow.audio.spectrogram import ShaderSpectrogram
        from shaderflow.piano import PianoNote
        self.audio = ShaderAudio(scene=self, name="iAudio", file="/path/to/audio.ogg")
        self.spectrogram = ShaderSpectrogram(scene=self, audio=self.audio, length=0)
        self.spectrogram.from_notes(start=PianoNote.from_frequency(20), end=PianoNote.from_frequency(18000), piano=True)
        self.shader.fragment = (shaders/"bars.frag")


class Visualizer(ShaderScene):
    """Radial Bars Music Visualizer Scene"""
    background = None   # path / numpy image; synthetic when None

    def build(self):
        from shaderflow.audio import ShaderAudio
        from shaderflow.audio.spectrogram import ShaderSpectrogram
        from shaderflow.audio.waveform import ShaderWaveform
        from shaderflow.piano import PianoNote
        self.audio = ShaderAudio(scene=self, name="iAudio", file="/path/to/audio.opus")
        self.waveform = ShaderWaveform(scene=self, audio=self.audio)
        self.spectrogram = ShaderSpectrogram(scene=self, length=0, audio=self.audio, smooth=False)
        self.spectrogram.from_notes(start=PianoNote.from_frequency(20), end=PianoNote.from_frequency(14000), piano=True)
        image = self.background if self.background is not None else synthetic_background()
        self.back = ShaderTexture(scene=self, name="background").from_image(image)
        self.shader.fragment = (shaders/"visualizer.frag")

    def handle(self, message) -> None:
        from shaderflow.message import ShaderMessage
        ShaderScene.handle(self, message)
        if isinstance(message, ShaderMessage.Window.FileDrop):
            self.back.from_image(message.first)


class PianoRoll(ShaderScene):
    """Piano roll: ShaderPiano + examples/shaders/piano.frag (BASELINE config 5's scene; the reference ships the
    module but neither a scene nor a fragment for it). Notes come from `notes` rows (pitch, start, end, channel,
    velocity) or a seeded synthetic list: 4 channels, 16 notes/s, pitches 36-96, 0.1-1 s."""
    notes = None
    midi = None         # path of a Standard MIDI File: its notes instead of `notes` / the synthetic list
    seconds: float = 10.0

    def build(self):
        from shaderflow.piano import PianoNote, ShaderPiano
        self.piano = ShaderPiano(scene=self)
        if self.midi is not None:
            self.piano.load_midi(self.midi)
        for (pitch, start, end, channel, velocity) in (() if self.midi is not None else self.notes if self.notes is not None else synthetic_notes(self.seconds)):
            self.piano.add_note(PianoNote(note=int(pitch), start=float(start), end=float(end), channel=int(channel), velocity=int(velocity)))
        self.shader.fragment = (shaders/"piano.frag")


def synthetic_notes(seconds: float, seed: int = 5, channels: int = 4, rate: float = 16.0) -> list:
    rng = np.random.default_rng(seed)
    notes = []
    for channel in range(channels):
        for s in np.sort(rng.uniform(-0.5, seconds, int(seconds*rate/channels))):
            notes.append((int(rng.integers(36, 97)), float(s), float(s + rng.uniform(0.1, 1.0)), channel, int(rng.integers(20, 128))))
    return notes


class RayMarch(ShaderScene):
    """Ray Marching demo"""
    def build(self):
        self.shader.fragment = (shaders/"raymarch.frag")


class Mandelbrot(ShaderScene):
    """Mandelbrot fractal"""
    def build(self):
        self.shader.fragment = (shaders/"mandelbrot.frag")


class Tetration(ShaderScene):
    """Complex tetration fractal"""
    def build(self):
        self.shader.fragment = (shaders/"tetration.frag")

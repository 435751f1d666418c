"""Host-side logic of the drop-in layer (no GPU): the reference's own Resolution.fit tests ported
(resolution.py:90-116), scheduler/scene clock against the goldens, module graph order, uniform packing,
the GLSL registry, host DynamicNumber against the reference goldens."""
import json

import numpy as np
import pytest

from oracle import audio_np as A
from shaderflow_b200 import _native as N
from shaderflow_b200 import registry
from shaderflow_b200.dynamics import DynamicNumber
from shaderflow_b200.resolution import Resolution
from shaderflow_b200.scheduler import Scheduler
from shaderflow_b200.shader import pack_uniforms


class TestResolutionFit:
    """Ports of the reference's inline `class __pytest__` (resolution.py:90-116)"""
    def test_keep_nothing(self):
        assert Resolution.fit(old=(1920, 1080)) == (1920, 1080)

    def test_override_components(self):
        assert Resolution.fit(old=(1920, 1080), new=(1280, None)) == (1280, 1080)
        assert Resolution.fit(old=(1920, 1080), new=(None, 720)) == (1920, 720)

    def test_missing_components(self):
        with pytest.raises(ValueError):
            Resolution.fit(old=(1920, None), new=(1280, None))
        with pytest.raises(ValueError):
            Resolution.fit(old=(None, 1080), new=(None, None))

    def test_aspect_ratio(self):
        assert Resolution.fit(old=(1920, 1080), new=(1280, None), ar=16/9) == (1280, 720)
        assert Resolution.fit(old=(1920, 1080), new=(None, 720), ar=16/9) == (1280, 720)
        assert Resolution.fit(old=(1920, 1080), new=(1000, None), ar=2.0) == (1000, 500)
        assert Resolution.fit(old=(1920, 1080), new=(None, 500), ar=2.0) == (1000, 500)

    def test_aspect_ratio_prioritize_width(self):
        assert Resolution.fit(old=(1920, 1080), new=(1000, 720), ar=2) == (1000, 500)

    def test_limit_maximum_resolution(self):
        assert Resolution.fit(old=(3840, 2160), new=(3800, 2100), max=(1920, 1080)) == (1920, 1080)
        assert Resolution.fit(old=(3000, 3000), new=(2000, 2000), max=(6000, 720), ar=16/9) == (1280, 720)

    def test_against_reference_outputs(self, golden_dir):
        for row in json.loads((golden_dir/"resolution_fit.json").read_text()):
            kw = {k: (tuple(v) if isinstance(v, list) else v) for k, v in row["kwargs"].items()}
            assert list(Resolution.fit(**kw)) == row["result"]


def test_scheduler_freewheel_clock_matches_golden(golden_dir):
    """Driving the mirrored Scheduler the way ShaderScene.main does reproduces the reference's
    per-frame time/dt (goldens were produced by the reference's own SchedulerTask)"""
    gold = np.load(golden_dir/"audio_chirp_1000.npz")
    state = dict(time=0.0, dt=0.0, seen=[])
    def step(dt=0.0):
        state["seen"].append((state["time"], state["dt"]))
        state["dt"] = dt*1.0
        state["time"] += state["dt"]
    sched = Scheduler()
    task = sched.new(task=step, frequency=24.0, freewheel=True, precise=True)
    for _ in range(len(gold["time"])):
        assert sched.next() is task
    assert np.array_equal([t for t, _ in state["seen"]], gold["time"])
    assert np.array_equal([d for _, d in state["seen"]], gold["dt"])


def test_dry_scene_graph_and_pipeline():
    """Module creation order and uniform names are what the reference's metaprogramming would
    declare (scene.py:128-195, camera.py:146-201, audio/module.py:413-421)"""
    import examples.demo as demo
    scene = demo.Visualizer(backend="dry")
    scene.initialize()
    kinds = [type(m).__name__ for m in scene.modules]
    assert kinds[:4] == ["Visualizer", "ShaderFrametimer", "ShaderKeyboard", "ShaderCamera"]
    assert kinds[4:13] == ["ShaderDynamics"]*9
    assert [m.name for m in scene.modules[4:13]] == ["iCameraPosition", "iCameraSeparation", "iCameraRotation",
        "iCameraZenith", "iCameraZoom", "iCameraIsometric", "iCameraFocalLength", "iCameraOrbital", "iCameraDolly"]
    assert kinds[13:17] == ["ShaderProgram", "ShaderTexture", "ShaderProgram", "ShaderTexture"]
    assert kinds[17:] == ["ShaderAudio", "ShaderDynamics", "ShaderDynamics", "ShaderWaveform", "ShaderTexture",
                          "ShaderSpectrogram", "ShaderTexture", "ShaderTexture"]
    names = [v.name for v in scene.full_pipeline()]
    for required in ("iTime", "iResolution", "iCameraPosition", "iCameraZoom", "iAudioVolume", "iAudioVolumeIntegral",
                     "iAudioSTD", "iSpectrogramBins", "iWaveformLength", "backgroundSize", "background0x0", "iScreen0x0"):
        assert required in names
    assert "iCameraRotation" not in names                      # primary=False (camera.py:155-159)
    assert scene.spectrogram.spectrogram_bins == 115
    assert scene.spectrogram.texture.repeat_y is False and scene.waveform.texture.repeat_x is False
    assert scene.waveform.chunk_size == 735 and scene.waveform._points == 180
    # texture aliases of texture.py:354-368
    assert set(scene.back.sampler_names()) == {"background0x0", "background"}
    scene.shader.compile()
    assert scene.shader.scene_info["name"] == "visualizer"
    values, samplers = scene.shader.gather(scene.full_pipeline())
    block = scene.shader.uniform_block(values)
    assert tuple(block.iResolution) == (1920.0, 1080.0) and block.iCameraMode == 1 and block.iCameraZoom == 1.0
    assert tuple(block.iCameraForward) == (0.0, 0.0, 1.0) and abs(block.iQuality - 0.5) < 1e-7
    assert {"background", "iSpectrogram", "iWaveform"} <= set(samplers)


def test_dry_main_steps_time_like_the_reference():
    import examples.demo as demo
    scene = demo.ShaderToy(backend="dry")
    seen = []
    orig = scene.next
    scene.main(width=64, height=36, fps=24.0, time=1.25)
    time, dt, _ = A.frame_clock(30, 24.0)
    assert scene.frame_index == 30 and scene.total_frames == 30
    assert abs(scene.time - (time[-1] + 1/24)) < 1e-12
    assert scene.resolution == (64, 36) and scene.render_resolution == (64, 36)
    scene.ssaa = 2
    assert scene.render_resolution == (128, 72) and scene.fusable == 2
    scene.subsample = 3
    assert scene.fusable is None


def test_pack_uniforms_casts_and_reports_missing():
    block = pack_uniforms(N.Uniforms.defaults(64, 32), dict(iTime=np.float64(1/3), iResolution=(64, 32), iFrame=7,
        iCameraPosition=np.array([1, 2, 3.0]), iAudioVolume=np.array(0.25)), ["iAudioVolume"])
    assert block.iTime == np.float32(1/3) and block.iFrame == 7 and tuple(block.iCameraPosition) == (1.0, 2.0, 3.0)
    assert block.extra[0][0] == 0.25
    with pytest.raises(RuntimeError, match="iAudioSTD"):
        pack_uniforms(N.Uniforms.defaults(64, 32), {}, ["iAudioSTD"])


def test_registry_directive_and_normalisation():
    assert registry.resolve("// sfb200: scene=mandelbrot\nvoid main(){}") == "mandelbrot"
    a = "void main() {\n  fragColor = vec4(1.0); // white\n}"
    b = "/* header */ void main(){fragColor=vec4(1.0);}"
    assert registry.digest(a) == registry.digest(b)
    assert registry.resolve(a) is None
    assert len(registry.KNOWN_HASHES) == 16 and len(set(registry.KNOWN_HASHES.values())) == 16


@pytest.mark.reference
def test_registry_hashes_match_reference_tree():
    from pathlib import Path
    ref = Path("/root/reference")
    if not ref.exists():
        pytest.skip("no /root/reference here")
    for rel, name in (("examples/basic/shaders/visualizer.frag", "visualizer"), ("examples/basic/shaders/bars.frag", "bars"),
                      ("examples/fractals/shaders/mandelbrot.frag", "mandelbrot"),
                      ("shaderflow/resources/shaders/fragment/default.glsl", "default"),
                      ("examples/basic/shaders/multipass.frag", "multipass"), ("examples/basic/shaders/motionblur.frag", "motionblur"),
                      ("examples/basic/shaders/life/simulation.glsl", "life_simulation"),
                      ("examples/basic/shaders/life/visuals.glsl", "life_visuals")):
        assert registry.resolve(ref/rel) == name
    # the GLSL written inline in the reference's demo.py is recognised as the text the user scene passes
    import ast
    inline = {}
    for cls in (n for n in ast.parse((ref/"examples/basic/demo.py").read_text()).body if isinstance(n, ast.ClassDef)):
        for node in ast.walk(cls):
            if isinstance(node, ast.Assign) and isinstance(node.value, ast.Constant) and isinstance(node.value.value, str) \
                    and isinstance(node.targets[0], ast.Attribute) and node.targets[0].attr == "fragment":
                inline[(cls.name, node.targets[0].value.attr)] = registry.resolve(node.value.value)
    assert inline == {("MultiShader", "child"): "multishader_child", ("MultiShader", "shader"): "multishader",
                      ("Dynamics", "shader"): "dynamics", ("Audio", "shader"): "audio"}


def test_host_dynamics_matches_reference_golden(golden_dir):
    """The host DynamicNumber (camera, user knobs) against the reference's volume/std recurrences"""
    gold = np.load(golden_dir/"audio_noise.npz")
    vol = DynamicNumber(frequency=2, zeta=1, response=0, value=0, integrate=True)
    std = DynamicNumber(frequency=10, zeta=1, response=0, value=0)
    for k in range(len(gold["dt"])):
        vol.target = np.float32(gold["vol_target"][k]); std.target = np.float32(gold["std_target"][k])
        vol.next(dt=abs(gold["dt"][k])); std.next(dt=abs(gold["dt"][k]))
        assert vol.value == gold["volume"][k] and vol.integral == gold["volume_integral"][k] and std.value == gold["std"][k]


def test_camera_basis_and_scripted_rotation():
    import examples.demo as demo
    scene = demo.Basic(backend="dry"); scene.initialize()
    cam = scene.camera
    assert np.allclose(cam.right, (1, 0, 0)) and np.allclose(cam.up, (0, 1, 0)) and np.allclose(cam.forward, (0, 0, 1))
    cam.rotate((0, 1, 0), 90)
    cam.rotation.value = cam.rotation.target
    assert np.allclose(cam.forward, (1, 0, 0), atol=1e-12) and np.allclose(cam.right, (0, 0, -1), atol=1e-12)
    assert abs(cam.fov - 90.0) < 1e-9


def test_synthetic_inputs_agree_with_the_oracle_copies():
    from oracle import glsl_np as G
    from shaderflow_b200 import synthetic
    assert np.array_equal(synthetic.chirp(0.5), A.synth_chirp(0.5))
    assert np.array_equal(synthetic.noise(0.2, seed=3), A.synth_noise(0.2, seed=3))
    assert np.array_equal(synthetic.sine(0.1), A.synth_sine(0.1))
    assert np.array_equal(synthetic.background(96, 54), G.synthetic_background(96, 54))


def test_standalone_spectrogram_matrix_matches_reference_golden(golden_dir):
    """BrokenSpectrogram outside a scene: the CSR filterbank equals the reference's (piano and default banks)"""
    from shaderflow_b200.audio import BrokenAudio, BrokenSpectrogram
    for name, notes in (("audio_c1_sine", (15, 129)), ("audio_chirp_1000", None)):
        gold = np.load(golden_dir/f"{name}.npz")
        spec = BrokenSpectrogram(audio=BrokenAudio())
        if notes:
            spec.from_notes(*notes, piano=True)
        m = spec.spectrogram_matrix()
        assert np.array_equal(m.indptr, gold["bank_indptr"]) and np.array_equal(m.indices, gold["bank_indices"])
        assert np.array_equal(m.data, gold["bank_data"]) and np.array_equal(spec.spectrogram_frequencies, gold["bank_frequencies"])
        assert spec.spectrogram_matrix() is m      # cached per configuration


@pytest.mark.reference
def test_reference_example_files_run_unchanged_on_the_host_layer():
    """The reference's own examples/basic/demo.py and examples/fractals/fractals.py, imported as they are
    with `shaderflow` aliased to this package: scenes build, their GLSL files resolve by content hash"""
    import importlib.util, sys
    from pathlib import Path
    ref = Path("/root/reference/examples")
    if not ref.exists():
        pytest.skip("no /root/reference here")
    import shaderflow_b200
    shaderflow_b200.install_alias()
    found = {}
    for rel in ("basic/demo.py", "fractals/fractals.py"):
        spec = importlib.util.spec_from_file_location(f"ref_examples_{Path(rel).stem}", ref/rel)
        module = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(module)
        found.update(vars(module))
    expected = dict(Basic="default", ShaderToy="shadertoy", RayMarch="raymarch", Mandelbrot="mandelbrot",
                    Tetration="tetration", MusicBars="bars", Waveform="waveform")
    for cls, kernel in expected.items():
        scene = found[cls](backend="dry")
        scene.initialize()
        scene.shader.compile()
        assert scene.shader.scene_info["name"] == kernel, cls
        scene.main(width=64, height=36, time=0.05)           # host loop only (dry)
        assert scene.frame_index == 3


def test_multi_program_scene_graphs_compile_without_a_gpu():
    """The scenes of SURVEY §8f-2/3 on the dry backend: program / texture graph, kernel selection by directive,
    sampler names and aliases of layered and temporal textures (texture.py:354-368), required uniforms"""
    import examples.demo as demo
    from shaderflow_b200.shader import ShaderProgram
    expect = {
        demo.MultiShader: {"iScreen": "multishader", "child": "multishader_child"},
        demo.Multipass: {"iScreen": "multipass"}, demo.MotionBlur: {"iScreen": "motionblur"},
        demo.Dynamics: {"iScreen": "dynamics"}, demo.Audio: {"iScreen": "audio"},
        demo.Life: {"iScreen": "life_visuals", "iLife": "life_simulation"}, demo.PianoRoll: {"iScreen": "piano"},
    }
    for cls, programs in expect.items():
        scene = cls(backend="dry"); scene.initialize()
        found = {}
        for module in scene.modules:
            if isinstance(module, ShaderProgram) and not module.texture.final:
                module.compile()
                found[module.name] = module.scene_info["name"]
                values, samplers = module.gather(scene.full_pipeline())
                for name in module.scene_info["extra"]:
                    assert name in values, (cls.__name__, name)
                for name in module.scene_info["samplers"][:module.scene_info["required"]]:
                    assert name in samplers, (cls.__name__, name)
        assert found == programs, cls.__name__
    blur = demo.MotionBlur(backend="dry"); blur.initialize()
    names = blur.shader.texture.sampler_names()
    assert {"iScreen0x0", "iScreen9x1", "iScreen", "iScreen3"} <= set(names) and len(blur.shader.texture.matrix) == 10
    assert names["iScreen"] is blur.shader.texture.matrix[0][1] and names["iScreen3"] is blur.shader.texture.matrix[3][1]
    life = demo.Life(backend="dry"); life.initialize()
    assert life.simulation.texture.components == 1 and str(life.simulation.texture.dtype) == "float32"
    assert {v.name: v.value for v in life.pipeline()}["iLifePeriod"] == 6


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours) prints exactly one JSON line with the
    contract's keys; stdout carries nothing else"""
    import json, subprocess, sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    done = subprocess.run([sys.executable, str(root/"bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                           "--width", "640", "--height", "360"], capture_output=True, text=True, timeout=600)
    assert done.returncode == 0, done.stderr[-2000:]
    lines = [l for l in done.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["higher_is_better"] is True
    assert line["metric"] == "4K@2xSSAA music-visualizer frames/sec" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == dict(value=line["value"], unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)

"""End to end through the reference-facing Python API on the GPU: example scenes built against
`shaderflow.*` (alias of shaderflow_b200) → ShaderScene.main() → frames, checked against the oracle
run with the same inputs (audio track by oracle/audio_np.py, pixels by oracle/glsl_np.py)."""
import numpy as np
import pytest

from oracle import audio_np as A
from oracle import glsl_np as G

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

W, H = 256, 144


def collect(scene, **flags):
    """Runs main() with no sink and returns {frame index: (H, W, 3) uint8, bottom row first}"""
    frames = {}
    def grab(index, pointer):
        scene.cuda.sync()
        frames[index] = scene.frame_tensor.cpu().numpy().copy()
    scene.main(width=W, height=H, on_frame=grab, **flags)
    return frames


def oracle_visualizer_frame(k, clip, bg, ssaa, subsample, fps=60.0, runtime=1.0):
    cfg = A.TrackConfig(fps=fps, bank=A.BankConfig.from_notes(15, 129, piano=True))
    tr = A.audio_track(clip, k + 1, cfg)
    tex = dict(
        background=G.Texture(np.flipud(bg).copy(), linear=True),
        iSpectrogram=G.Texture(tr["column"][k].reshape(-1, 1, 2).copy(), linear=False, repeat_x=True, repeat_y=False),
        iWaveform=G.Texture(tr["wave"][k].reshape(1, -1, 2).copy(), linear=True, repeat_x=False, repeat_y=False),
    )
    u = G.Uniforms(iTime=tr["time"][k], iTau=(tr["time"][k]/runtime) % 1.0, iDuration=runtime, iResolution=(W, H),
                   iWantAspect=W/H, iFrame=round(tr["time"][k]*fps),
                   extra=dict(iAudioVolume=tr["volume"][k], iAudioSTD=tr["std"][k]))
    return G.render("visualizer", u, tex, W, H, ssaa=float(ssaa), subsample=subsample), tr


@pytest.mark.parametrize("ssaa,subsample", [(2, 2), (1, 2)])
def test_visualizer_scene_end_to_end(ssaa, subsample):
    from examples.demo import Visualizer, synthetic_background
    clip = A.synth_noise(1.0, seed=0)
    bg = synthetic_background(240, 135)
    Visualizer.background = bg
    try:
        scene = Visualizer()
        scene.initialize()
        scene.audio.load(clip, 44100)
        frames = collect(scene, ssaa=ssaa, subsample=subsample, time=1.0)
    finally:
        Visualizer.background = None
    assert sorted(frames) == list(range(60))
    assert (scene.fusable is not None) == (ssaa == 2)
    for k in (0, 1, 30, 59):
        ref, tr = oracle_visualizer_frame(k, clip, bg, ssaa, subsample)
        d = np.abs(frames[k].astype(int) - ref["final_u8"].astype(int))
        assert (d <= 1).mean() > 0.998, (k, (d <= 1).mean(), d.max())
        assert (d == 0).mean() > 0.9
    # module state published from the GPU track equals the oracle's host recurrences
    assert scene.audio.tell == tr["tell"][59]
    assert abs(float(scene.audio.volume.value) - tr["volume"][59]) < 1e-5
    assert abs(float(scene.audio.std.value) - tr["std"][59]) < 1e-5
    got = scene.spectrogram.columns.cpu().numpy()
    assert np.abs(got[59] - tr["column"][59]).max() <= 1e-5*max(tr["column"].max(), 1.0)


@pytest.mark.parametrize("name,scene_key", [("Basic", "default"), ("ShaderToy", "shadertoy"), ("Mandelbrot", "mandelbrot"),
                                            ("Tetration", "tetration"), ("RayMarch", "raymarch")])
def test_textureless_scenes_end_to_end(name, scene_key):
    import examples.demo as demo
    scene = getattr(demo, name)()
    frames = collect(scene, ssaa=2, subsample=2, time=0.1, fps=30.0, quality=20.0)
    assert sorted(frames) == [0, 1, 2]
    time, _, _ = A.frame_clock(3, 30.0)
    for k in (0, 2):
        u = G.Uniforms(iTime=time[k], iTau=(time[k]/0.1) % 1.0, iDuration=0.1, iQuality=0.2, iFramerate=30.0,
                       iResolution=(W, H), iWantAspect=W/H)
        ref = G.render(scene_key, u, {}, W, H, ssaa=2.0, subsample=2)
        d = np.abs(frames[k].astype(int) - ref["final_u8"].astype(int))
        assert (d <= 1).mean() > (0.97 if scene_key == "tetration" else 0.99), (name, k, (d <= 1).mean())


def test_bars_and_waveform_scenes():
    import examples.demo as demo
    clip = A.synth_chirp(0.5)
    for cls, key in ((demo.MusicBars, "bars"), (demo.Waveform, "waveform")):
        scene = cls()
        scene.initialize()
        scene.audio.load(clip, 44100)
        frames = collect(scene, ssaa=1, subsample=1, time=0.5)
        k = 29
        notes = (15, 133) if key == "bars" else (15, 129)
        cfg = A.TrackConfig(bank=A.BankConfig.from_notes(*notes, piano=True))
        tr = A.audio_track(clip, k + 1, cfg)
        tex = dict(iSpectrogram=G.Texture(tr["column"][k].reshape(-1, 1, 2).copy(), linear=False, repeat_x=True, repeat_y=False),
                   iWaveform=G.Texture(tr["wave"][k].reshape(1, -1, 2).copy(), linear=False, repeat_x=False, repeat_y=False))
        u = G.Uniforms(iTime=tr["time"][k], iResolution=(W, H), iWantAspect=W/H)
        ref = G.render(key, u, tex, W, H, ssaa=1.0, subsample=1)
        d = np.abs(frames[k].astype(int) - ref["final_u8"].astype(int))
        assert (d <= 1).mean() > 0.995, (key, (d <= 1).mean())


def test_export_sinks_and_frame_ranges(tmp_path):
    """output='*.rgb' writes every frame, bottom row first, identical to the no-sink render;
    'pipe' returns the same bytes; 'null' drops them; frames=(a,b) renders only that range"""
    import examples.demo as demo
    scene = demo.ShaderToy()
    plain = collect(scene, ssaa=1, subsample=1, time=0.2)
    assert len(plain) == 12
    path = scene.main(width=W, height=H, ssaa=1, subsample=1, time=0.2, output=tmp_path/"frames.rgb", buffers=3)
    raw = np.fromfile(path, np.uint8).reshape(12, H, W, 3)
    for k in range(12):
        assert np.array_equal(raw[k], plain[k])
    import shutil
    if not shutil.which("ffmpeg"):
        blob = scene.main(width=W, height=H, ssaa=1, subsample=1, time=0.2, output="pipe")
        assert isinstance(blob, bytes) and np.array_equal(np.frombuffer(blob, np.uint8).reshape(12, H, W, 3), raw)
    assert scene.main(width=W, height=H, ssaa=1, subsample=1, time=0.2, output="null") is None
    part = collect(scene, ssaa=1, subsample=1, time=0.2, frames=(5, 9))
    assert sorted(part) == [5, 6, 7, 8] and all(np.array_equal(part[k], plain[k]) for k in part)


def test_frame_ranges_of_feedback_scenes_keep_their_history():
    """MotionBlur averages the last 10 frames of its own output (temporal texture): exporting frames=(a, b) must
    shade every frame before `a` too, or the range starts from an empty history and differs from the full export.
    The same rule keeps such scenes from frame-sharding: under torchrun rank 0 exports them alone."""
    import examples.demo as demo
    demo.MotionBlur.background = demo.synthetic_background(160, 90)
    try:
        scene = demo.MotionBlur()
        full = collect(scene, ssaa=1, subsample=1, time=24/60)
        part = collect(scene, ssaa=1, subsample=1, time=24/60, frames=(14, 20))
    finally:
        demo.MotionBlur.background = None
    assert sorted(part) == list(range(14, 20))
    assert all(np.array_equal(part[k], full[k]) for k in part)
    assert not np.array_equal(full[14], full[3])             # the history really changes the picture


def test_a_failed_export_leaves_the_scene_usable(tmp_path):
    """An exception in the frame loop (here: a user hook) must not leave the sink ring acquired or the render state
    half-set: the next main() on the same scene works"""
    import examples.demo as demo
    scene = demo.ShaderToy()
    def boom(index, pointer):
        if index == 3:
            raise KeyError("user hook failed")
    with pytest.raises(KeyError):
        scene.main(width=W, height=H, ssaa=1, subsample=1, time=0.2, on_frame=boom)
    class Dies(demo.ShaderToy):
        fail = True
        def update(self):
            if type(self).fail and self.frame_index == 4:
                raise ValueError("module failed mid-export")
    bad = Dies()
    with pytest.raises(ValueError):
        bad.main(width=W, height=H, ssaa=1, subsample=1, time=0.2, output=tmp_path/"bad.rgb")
    assert bad.render_enabled and bad._frame_target is None
    Dies.fail = False
    path = bad.main(width=W, height=H, ssaa=1, subsample=1, time=0.2, output=tmp_path/"good.rgb")
    assert np.fromfile(path, np.uint8).size == 12*H*W*3


def test_unknown_glsl_is_compiled_at_run_time_and_hash_lookup_works():
    """A fragment the registry does not recognise goes through the GLSL → CUDA translator (tests/test_gpu_jit.py holds
    that path to the evaluator); text that is not GLSL is an error naming what the compiler did not accept"""
    from examples.demo import ShaderScene
    from shaderflow_b200 import registry
    class Custom(ShaderScene):
        def build(self):
            self.shader.fragment = "void main() { fragColor = vec4(1.0); }"
    scene = Custom()
    seen = []
    def grab(index, pointer):
        scene.cuda.sync(); seen.append(scene.frame_tensor.cpu().numpy().copy())
    scene.main(width=64, height=36, time=0.1, on_frame=grab)
    assert scene.shader.scene_id >= 1000 and len(seen) == 6 and all((f == 255).all() for f in seen)
    class Broken(ShaderScene):
        def build(self):
            self.shader.fragment = "void main() { fragColor = vec4(1.0) }"
    with pytest.raises(RuntimeError, match="could not be compiled for the CUDA backend"):
        Broken().main(width=64, height=36, time=0.1)
    assert {"default", "shadertoy", "visualizer", "bars", "waveform", "mandelbrot", "tetration", "raymarch",
            "multipass", "motionblur", "life_simulation", "life_visuals"} <= set(registry.KNOWN_HASHES.values())


def test_sharded_export_equals_single_gpu_export():
    """Frame-sharded export over 2 GPUs (peer writes into rank 0's HBM, or the NCCL fallback) must produce exactly the
    bytes of the single-GPU export — tools/shard_check.py under torchrun. Needs two visible GPUs."""
    import subprocess, sys
    from pathlib import Path
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = Path(__file__).resolve().parents[1]
    for env_extra in ({}, {"SFB_NO_PEER_FRAMES": "1"}):
        import os
        done = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                               "--master-port", "29541", str(root/"tools"/"shard_check.py")], capture_output=True, text=True, timeout=600,
                              env={**os.environ, **env_extra})
        assert done.returncode == 0 and "-> OK" in done.stdout, (env_extra, done.stdout[-2000:], done.stderr[-2000:])

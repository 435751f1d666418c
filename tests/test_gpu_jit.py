"""Run-time compiled fragment programs on the GPU (shaderflow_b200/glsl → sfb_jit_compile → sfb_program_load →
sfb_render_*), held to:
  * the mechanical evaluator oracle/glsl_exec.py executing the SAME GLSL text (tests/shaders/*.frag), float by float;
  * for ShaderFlow's std-lib API, tests/golden/jit_stdlib.npz = stdlib.frag evaluated behind the reference's own
    header and include files (tests/golden/make_golden_jit.py);
  * the ahead-of-time kernel of the same shader (examples/shaders/piano.frag has both).
Tolerance: north_star's 1e-3 per float channel is the gate on every fragment; transcendental functions differ by
an ulp or two between numpy and libdevice, everything else is the same float32 arithmetic, so the typical error is
also held below 1e-5."""
import numpy as np
import pytest

from oracle import glsl_np as G
from tests import jit_cases as J
from tests.helpers import native_textures, native_uniforms

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def ctx():
    from shaderflow_b200 import _native as N
    c = N.Context(0)
    yield c
    c.destroy()


def load(ctx, name, header):
    from shaderflow_b200 import glsl
    image, translation, _ = glsl.build((J.SHADERS/f"{name}.frag").read_text(), header)
    scene = ctx.program_load(image, len(translation.samplers))
    return scene, dict(extra=translation.extra, extra_types=translation.extra_types, samplers=translation.samplers)


def screen(ctx, scene, info, u, tex, Wr, Hr):
    nt = native_textures(ctx, tex)
    rgba = torch.zeros((Hr, Wr, 4), dtype=torch.uint8, device="cuda")
    f32 = torch.zeros((Hr, Wr, 4), dtype=torch.float32, device="cuda")
    ctx.render_screen(scene, native_uniforms(u, info), [nt[n] for n in info["samplers"]], Wr, Hr, rgba, f32, 0)
    ctx.sync()
    return rgba.cpu().numpy(), f32.cpu().numpy(), nt


@pytest.mark.parametrize("name", J.CORPUS)
def test_compiled_program_equals_the_evaluated_text(ctx, name):
    scene, info = load(ctx, name, J.HEADER)
    u = J.uniforms(extra=dict(J.USER_UNIFORMS))
    want, gone = J.evaluate(name)
    rgba, got, _ = screen(ctx, scene, info, u, J.corpus_textures(), J.W, J.H)
    if name == "textured":
        assert gone.any() and not gone.all()
    assert np.all(got[gone] == 0.0)                                   # discarded fragments keep the cleared target
    err = np.abs(got - want)[~gone]
    assert err.max() <= 1e-3, (name, err.max())
    assert np.median(err) <= 1e-6 and (err <= 1e-5).mean() >= 0.99, (name, np.median(err), (err <= 1e-5).mean())
    d = np.abs(rgba[~gone].astype(int) - G.to_unorm8(want)[~gone].astype(int))
    assert d.max() <= 1 and (d == 0).mean() >= 0.99
    ctx.program_unload(scene)


def test_std_lib_equals_the_references_glsl(ctx, golden_dir):
    gold = np.load(golden_dir/"jit_stdlib.npz")
    scene, info = load(ctx, "stdlib", J.STDLIB_HEADER)
    assert info["samplers"] == ["background0x0"] and info["extra"] == ["iProbe"]
    tex = {"background0x0": J.stdlib_textures()["background"]}
    for c, camera in enumerate(J.STDLIB_CAMERAS):
        for probe in range(J.STDLIB_PROBES if c == 0 else 1):
            u = J.uniforms(extra=dict(iProbe=probe), **camera)
            _, got, _ = screen(ctx, scene, info, u, tex, J.W, J.H)
            want = gold[f"camera{c}_probe{probe}"]
            err = np.abs(got - want)/np.maximum(1.0, np.abs(want))
            assert err.max() <= 1e-3, (c, probe, err.max())
            assert (err <= 2e-5).mean() >= 0.99, (c, probe, (err <= 2e-5).mean())
    ctx.program_unload(scene)


def test_fused_frame_equals_screen_then_final_and_targets_of_other_formats(ctx):
    from shaderflow_b200 import _native as N
    scene, info = load(ctx, "plasma", J.HEADER)
    W, H, S = 48, 28, 2
    u = G.Uniforms(iTime=0.7, iResolution=(W, H), iWantAspect=W/H, iSSAA=float(S), extra=dict(J.USER_UNIFORMS))
    nu = native_uniforms(u, info)
    rgba = torch.zeros((H*S, W*S, 4), dtype=torch.uint8, device="cuda")
    ctx.render_screen(scene, nu, [], W*S, H*S, rgba, None, 0)
    unfused = torch.zeros((H, W, 3), dtype=torch.uint8, device="cuda")
    ctx.render_final(rgba, W*S, H*S, W, H, S, 3, unfused)
    fused = torch.zeros((H, W, 3), dtype=torch.uint8, device="cuda")
    ctx.render_frame(scene, nu, [], W, H, S, S, 3, fused, 0)
    probe = torch.zeros((H*S, W*S, 4), dtype=torch.float32, device="cuda")
    again = torch.zeros((H, W, 3), dtype=torch.uint8, device="cuda")
    ctx.render_frame_probe(scene, nu, [], W, H, S, S, 3, again, probe, 0)
    ctx.sync()
    assert torch.equal(fused, again)
    d = (fused.int() - unfused.int()).abs()                    # final.glsl's float blend vs the integer box mean: rounding ties
    assert int(d.max()) <= 1 and float((d == 0).float().mean()) > 0.7
    assert np.array_equal(G.to_unorm8(probe.cpu().numpy()), rgba.cpu().numpy())
    # the same pass into a float texture (a child program's / layer's target): the colours before any store
    target = N.Texture(ctx, W*S, H*S, 4, N.DTYPE_F32)
    ctx.render_target(scene, nu, [], target)
    ctx.sync()
    assert np.array_equal(target.read(), probe.cpu().numpy())
    ctx.program_unload(scene)
    with pytest.raises(RuntimeError, match="not a loaded program"):
        ctx.render_frame(scene, nu, [], W, H, S, S, 3, fused, 0)


def test_user_scene_with_its_own_glsl_runs_through_the_public_api():
    """`shader.fragment = <any GLSL>` (shader.py:303-306): a fragment the registry does not know, reading a pipeline
    uniform, a user Uniform, a texture by its alias and the std-lib, rendered by main() — against the evaluator-free
    closed form of what it computes"""
    from examples.demo import ShaderScene
    from shaderflow_b200.texture import ShaderTexture
    from shaderflow_b200.variable import Uniform
    W, H = 64, 36
    image = G.synthetic_background(32, 18, seed=7)

    class Custom(ShaderScene):
        def build(self):
            ShaderTexture(scene=self, name="picture").from_numpy(image)
            self.shader.fragment = """
                uniform float iLevel;
                void main() {
                    vec3 c = texture(picture, astuv).rgb;
                    fragColor = vec4(mix(c, vec3(iLevel), step(0.5, astuv.x)), 1.0);
                    if (astuv.y > 0.75) fragColor.rgb = palette_magma(astuv.x);
                }"""
        def pipeline(self):
            yield from ShaderScene.pipeline(self)
            yield Uniform("float", "iLevel", 0.25)

    scene = Custom()
    frames = {}
    def grab(index, pointer):
        scene.cuda.sync(); frames[index] = scene.frame_tensor.cpu().numpy().copy()
    scene.main(width=W, height=H, ssaa=1, subsample=1, time=0.05, on_frame=grab)
    assert scene.shader.scene_id >= 1000 and scene.shader.scene_info["extra"] == ["iLevel"]
    frame = frames[0]
    right = frame[:H*3//4 - 1, W//2 + 1:]
    assert np.all(right == 64)                                               # round(0.25*255)
    tex = G.Texture(np.flipud(image).copy(), linear=True, repeat_x=True, repeat_y=True)
    u = G.Uniforms(iResolution=(W, H), iWantAspect=W/H)
    f = G.varyings(u, W, H)
    left = G.to_unorm8(tex.sample(f.astuv)[..., :3])[:H*3//4 - 1, :W//2 - 1]
    assert np.abs(frame[:H*3//4 - 1, :W//2 - 1].astype(int) - left.astype(int)).max() <= 1
    top = G.to_unorm8(G.palette_magma(f.astuv[..., 0]))[H*3//4 + 1:]
    assert np.abs(frame[H*3//4 + 1:].astype(int) - top.astype(int)).max() <= 1

    class Broken(ShaderScene):
        def build(self):
            self.shader.fragment = "void main() { fragColor = vec4(undefined_thing); }"
    with pytest.raises(RuntimeError, match="undefined_thing"):
        Broken().main(width=W, height=H, time=0.05)


def test_fragment_files_resolve_their_includes(tmp_path):
    """`shader.fragment = Path(...)` with `#include "file"` lines (shader.py:186,231-235): looked up next to the fragment
    and in include_directories, nested"""
    from examples.demo import ShaderScene
    (tmp_path/"lib").mkdir()
    (tmp_path/"lib"/"tone.glsl").write_text("float tone(float x) { return 0.25 + 0.5*step(0.5, x); }\n")
    (tmp_path/"colors.glsl").write_text('#include "tone.glsl"\nvec3 shade(vec2 p) { return vec3(tone(p.x), tone(p.y), 1.0); }\n')
    (tmp_path/"main.frag").write_text('#include "colors.glsl"\nvoid main() { fragColor = vec4(shade(astuv), 1.0); }\n')

    class FromFile(ShaderScene):
        def build(self):
            self.shader.include_directories.append(tmp_path/"lib")
            self.shader.fragment = tmp_path/"main.frag"
    scene = FromFile()
    frames = []
    def grab(index, pointer):
        scene.cuda.sync(); frames.append(scene.frame_tensor.cpu().numpy().copy())
    scene.main(width=32, height=16, ssaa=1, subsample=1, time=0.02, on_frame=grab)
    frame = frames[0]
    assert (frame[:8, :16] == (64, 64, 255)).all() and (frame[8:, 16:] == (191, 191, 255)).all()
    assert (frame[:8, 16:] == (191, 64, 255)).all() and (frame[8:, :16] == (64, 191, 255)).all()

    class Missing(ShaderScene):
        def build(self):
            self.shader.fragment = '#include "nowhere.glsl"\nvoid main() { fragColor = vec4(1.0); }'
    with pytest.raises(RuntimeError, match="nowhere.glsl"):
        Missing().main(width=32, height=16, time=0.02)


def test_piano_fragment_compiled_at_run_time_equals_its_ahead_of_time_kernel():
    """examples/shaders/piano.frag exists as GLSL and as a transliterated CUDA scene: without its `sfb200: scene=`
    directive the same file goes through the translator, and the export must not change"""
    from examples import demo
    from oracle import piano_np as P
    demo.PianoRoll.notes = P.synthetic_notes(4.0)
    text = (demo.shaders/"piano.frag").read_text().replace("// sfb200: scene=piano", "//")

    class Translated(demo.PianoRoll):
        def build(self):
            demo.PianoRoll.build(self)
            self.shader.fragment = text
    try:
        out = {}
        for kind in (demo.PianoRoll, Translated):
            scene = kind(device=0)
            frames = {}
            def grab(index, pointer, scene=scene, frames=frames):
                scene.cuda.sync(); frames[index] = scene.frame_tensor.cpu().numpy().copy()
            scene.main(width=160, height=90, ssaa=2, subsample=2, time=0.5, fps=60.0, on_frame=grab)
            out[kind] = (frames, scene.shader.scene_id)
    finally:
        demo.PianoRoll.notes = None
    assert out[demo.PianoRoll][1] < 1000 <= out[Translated][1]
    for k in (0, 13, 29):
        d = np.abs(out[demo.PianoRoll][0][k].astype(int) - out[Translated][0][k].astype(int))
        assert d.max() <= 1 and (d == 0).mean() > 0.999, (k, d.max(), (d == 0).mean())

"""Computes the registry hashes of the reference's shaders (build container only) and rewrites the
KNOWN_HASHES table of shaderflow_b200/registry.py. Only hashes are stored — no GLSL source."""
import re, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from shaderflow_b200 import registry

REF = Path("/root/reference")
FILES = {
    "default":    "shaderflow/resources/shaders/fragment/default.glsl",
    "shadertoy":  "examples/basic/shaders/shadertoy.frag",
    "visualizer": "examples/basic/shaders/visualizer.frag",
    "bars":       "examples/basic/shaders/bars.frag",
    "waveform":   "examples/basic/shaders/waveform.frag",
    "mandelbrot": "examples/fractals/shaders/mandelbrot.frag",
    "tetration":  "examples/fractals/shaders/tetration.frag",
    "raymarch":   "examples/basic/shaders/raymarch.frag",
    "multipass":  "examples/basic/shaders/multipass.frag",
    "motionblur": "examples/basic/shaders/motionblur.frag",
    "life_simulation": "examples/basic/shaders/life/simulation.glsl",
    "life_visuals":    "examples/basic/shaders/life/visuals.glsl",
}
# GLSL written inline in examples/basic/demo.py: (class, attribute the string is assigned to) → scene
INLINE = {
    ("MultiShader", "child"): "multishader_child", ("MultiShader", "shader"): "multishader",
    ("Dynamics", "shader"): "dynamics", ("Audio", "shader"): "audio",
}
lines = []
for name, rel in FILES.items():
    lines.append(f'    "{registry.digest((REF/rel).read_text())}": "{name}",  # {rel}')
import ast
demo = REF/"examples/basic/demo.py"
tree = ast.parse(demo.read_text())
for cls in (n for n in tree.body if isinstance(n, ast.ClassDef)):
    for node in ast.walk(cls):
        if isinstance(node, ast.Assign) and isinstance(node.value, ast.Constant) and isinstance(node.value.value, str):
            t = node.targets[0]
            if isinstance(t, ast.Attribute) and t.attr == "fragment" and isinstance(t.value, ast.Attribute):
                key = (cls.name, t.value.attr)
                if key in INLINE:
                    lines.append(f'    "{registry.digest(node.value.value)}": "{INLINE[key]}",  # examples/basic/demo.py:{node.lineno} ({cls.name}.{t.value.attr}, inline)')
path = ROOT/"shaderflow_b200"/"registry.py"
text = path.read_text()
text = re.sub(r"KNOWN_HASHES: dict\[str, str\] = \{.*?\n\}", "KNOWN_HASHES: dict[str, str] = {\n" + "\n".join(lines) + "\n}", text, flags=re.S)
path.write_text(text)
print("\n".join(lines))

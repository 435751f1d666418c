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
}
lines = []
for name, rel in FILES.items():
    lines.append(f'    "{registry.digest((REF/rel).read_text())}": "{name}",  # {rel}')
path = ROOT/"shaderflow_b200"/"registry.py"
text = path.read_text()
text = re.sub(r"KNOWN_HASHES: dict\[str, str\] = \{.*?\n\}", "KNOWN_HASHES: dict[str, str] = {\n" + "\n".join(lines) + "\n}", text, flags=re.S)
path.write_text(text)
print("\n".join(lines))

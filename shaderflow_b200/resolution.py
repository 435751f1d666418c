"""Resolution.fit — same contract as shaderflow/resolution.py:9-86 (its inline tests are ported in
tests/test_host.py, known answers in tests/golden/resolution_fit.json)."""
from __future__ import annotations

import builtins
import math
from typing import Optional

Pair = Optional[tuple]


class Resolution:

    @classmethod
    def fit(cls, old: Pair = None, new: Pair = None, max: Pair = None, ar: Optional[float] = None,
            scale: float = 1.0, multiple: int = 2) -> tuple[int, int]:
        """A target at `old` is asked to become `new` (either part may be None), optionally locked to
        aspect ratio `ar` and bounded by `max`; the answer is scaled and snapped to `multiple`."""
        ow, oh = old or (None, None)
        nw, nh = new or (None, None)
        mw, mh = max or (None, None)
        width, height = (nw or ow), (nh or oh)
        if not (width and height):
            raise ValueError(f"Can't get a resolution missing component(s): ({width=}, {height=})")

        if ar is None:
            width, height = min(width, mw or math.inf), min(height, mh or math.inf)
        else:
            by_width, by_height = (width, width/ar), (height*ar, height)
            if nh is None:        width, height = by_width
            elif nw is None:      width, height = by_height
            elif nw != ow:        width, height = by_width
            elif nh != oh:        width, height = by_height
            else:                 width, height = by_width
            # shrink both sides by the larger overshoot so the ratio survives the bounding box
            shrink = builtins.max(width/(min(width, mw or math.inf) or 1), height/(min(height, mh or math.inf) or 1)) or 1
            width, height = width/shrink, height/shrink

        return (multiple*round((width*scale)/multiple), multiple*round((height*scale)/multiple))


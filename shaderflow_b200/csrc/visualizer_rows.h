// visualizer_rows.h — entry of the separable visualizer kernel (visualizer_rows.cu), called by
// sfb_render_frame (render.cu). Not part of the ABI.
#pragma once
#include "scenes.cuh"

// Plans and launches the separable kernel when the frame qualifies (axis-aligned 2D camera, ssaa 1/2/4,
// vertical texel step small enough, window fits); *launched = kernels launched (0 when it did not). P must be fully filled
// (fill_params + geometry) with P.fast set.
// screen_alpha: with P.comps == 4, store fragColor.a (the RGBA8 iScreen pass of an unfused export) instead of 255.
int sfb_visualizer_rows_launch(const glsl::RenderParams& P, cudaStream_t stream, int* launched, int screen_alpha = 0);

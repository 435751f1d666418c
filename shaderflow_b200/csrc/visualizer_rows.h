// visualizer_rows.h — entry of the separable visualizer kernel (visualizer_rows.cu), called by
// sfb_render_frame (render.cu). Not part of the ABI.
#pragma once
#include "scenes.cuh"

// Plans and launches the separable kernel when the frame qualifies (axis-aligned 2D camera, ssaa 1/2/4,
// vertical texel step small enough, window fits); *launched = kernels launched (0 when it did not). P must be fully filled
// (fill_params + geometry) with P.fast set.
// screen_alpha: with P.comps == 4, store fragColor.a (the RGBA8 iScreen pass of an unfused export) instead of 255.
// background: the texture behind P.tex[0] (its tensor map stages the window with one 2D TMA load), may be NULL.
int sfb_visualizer_rows_launch(const glsl::RenderParams& P, cudaStream_t stream, int* launched, int screen_alpha = 0,
                               sfb_tex* background = nullptr);
// render.cu: cached CUtensorMap (device copy) over a texture's linear mirror for a box of box_w x box_h texels, or NULL
const void* sfb_background_tensor_map(sfb_tex* t, int box_w, int box_h);

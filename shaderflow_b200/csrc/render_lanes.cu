// render_lanes.cu — instantiates frame_lanes_kernel<scene, filter, ssaa> (fused K3+K4, one lane per sub-sample) for the
// ALU-bound scenes: Mandelbrot, Tetration, RayMarch at ssaa 2 and 4
#include "render_kernels.cuh"
using namespace sfb_render;

bool sfb_launch_frame_lanes(int scene, bool hw, const RenderParams& P, cudaStream_t st) {
    if (!(P.ssaa == 2 || P.ssaa == 4) || P.H > 65535) return false;
    if (scene == SFB_SCENE_MANDELBROT)     { hw ? launch_frame_lanes<SFB_SCENE_MANDELBROT, true>(P, st) : launch_frame_lanes<SFB_SCENE_MANDELBROT, false>(P, st); }
    else if (scene == SFB_SCENE_TETRATION) { hw ? launch_frame_lanes<SFB_SCENE_TETRATION, true>(P, st) : launch_frame_lanes<SFB_SCENE_TETRATION, false>(P, st); }
    else if (scene == SFB_SCENE_RAYMARCH)  { hw ? launch_frame_lanes<SFB_SCENE_RAYMARCH, true>(P, st) : launch_frame_lanes<SFB_SCENE_RAYMARCH, false>(P, st); }
    else return false;
    return true;
}

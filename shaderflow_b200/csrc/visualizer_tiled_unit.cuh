// visualizer_tiled_unit.cuh — one instantiation of the per-pixel tiled visualizer kernel per translation unit
// (visualizer_tiled_{1,2,3,4}.cu define VT_UNIT_S and include this): the ssaa² x 91 unrolled footprints make each
// instantiation a 40-second compile, so they build in parallel. c_blur is per translation unit: filled here.
#include "render_kernels.cuh"
#include "visualizer_tiled.cuh"

using namespace glsl;

#define SFB_VT_NAME2(s) sfb_launch_visualizer_tiled_##s
#define SFB_VT_NAME(s) SFB_VT_NAME2(s)

int SFB_VT_NAME(VT_UNIT_S)(const VisualizerParams& VP, cudaStream_t st) {
    constexpr int S = VT_UNIT_S;
    if (int e = sfb_render::build_blur_table()) return e;
    static bool configured[64] = {};                      // the attribute is per device
    int device = 0;
    SFB_CUDA(cudaGetDevice(&device));
    if (device >= 64 || !configured[device]) {
        SFB_CUDA(cudaFuncSetAttribute(visualizer_tiled_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(VT_SMEM)));
        if (device < 64) configured[device] = true;
    }
    dim3 block(VT_TILE_X, VT_TILE_Y), grid((VP.R.W + VT_TILE_X - 1)/VT_TILE_X, (VP.R.H + VT_TILE_Y - 1)/VT_TILE_Y);
    visualizer_tiled_kernel<S><<<grid, block, VT_SMEM, st>>>(VP);
    return SFB_OK;
}

// render_frame.cu — instantiates frame_kernel<scene, filter> (fused K3+K4, one thread per output pixel) for every scene
#include "render_kernels.cuh"
using namespace sfb_render;

int sfb_launch_frame(int scene, bool hardware_filter, const RenderParams& P, cudaStream_t stream) {
    if (scene == SFB_SCENE_VISUALIZER && P.fast)
        if (int e = build_blur_table()) return e;
    dispatch<LaunchFrame>(scene, hardware_filter, P, stream);
    return SFB_OK;
}

// render_screen.cu — instantiates screen_kernel<scene, filter> (K3: one thread per fragment) for every scene
#include "render_kernels.cuh"
using namespace sfb_render;

int sfb_launch_screen(int scene, bool hardware_filter, const RenderParams& P, cudaStream_t stream) {
    if (scene == SFB_SCENE_VISUALIZER && P.fast)
        if (int e = build_blur_table()) return e;
    dispatch<LaunchScreen>(scene, hardware_filter, P, stream);
    return SFB_OK;
}

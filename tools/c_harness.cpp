// Pure C harness over the C ABI (no Python / torch): renders one 4K 2xSSAA visualizer frame
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../include/sfb200.h"
#define CK(x) do { int rc = (x); if (rc) { printf("%s -> %d: %s\n", #x, rc, sfb_last_error()); return 1; } } while (0)
int main() {
    sfb_ctx* ctx; CK(sfb_ctx_create(0, nullptr, &ctx));
    sfb_tex *bg, *sp, *wv;
    CK(sfb_tex_create(ctx, 1920, 1080, 3, SFB_DTYPE_U8, SFB_FILTER_LINEAR, 1, 1, &bg));
    std::vector<unsigned char> img(1920*1080*3); for (size_t i = 0; i < img.size(); i++) img[i] = (unsigned char)(i*2654435761u >> 24);
    CK(sfb_tex_write(bg, img.data(), 0, 0, 0, 1920, 1080));
    CK(sfb_tex_create(ctx, 1, 115, 2, SFB_DTYPE_F32, SFB_FILTER_NEAREST, 1, 0, &sp));
    std::vector<float> col(230, 300.0f); CK(sfb_tex_write(sp, col.data(), 0, 0, 0, 1, 115));
    CK(sfb_tex_create(ctx, 180, 1, 2, SFB_DTYPE_F32, SFB_FILTER_LINEAR, 0, 0, &wv));
    std::vector<float> row(360, 0.3f); CK(sfb_tex_write(wv, row.data(), 0, 0, 0, 180, 1));
    sfb_uniforms u = {}; u.iTime = 1.0f; u.iDuration = 10; u.iResolution[0] = 3840; u.iResolution[1] = 2160; u.iWantAspect = 16.0f/9;
    u.iQuality = 0.5f; u.iSSAA = 2; u.iFramerate = 60; u.iCameraMode = 1; u.iCameraRight[0] = 1; u.iCameraUpward[1] = 1; u.iCameraForward[2] = 1;
    u.iCameraZoom = 1; u.iCameraFocalLength = 1; u.extra[0][0] = 0.8f; u.extra[1][0] = 0.2f;
    void* out; cudaMalloc(&out, 3840*2160*3);
    sfb_tex* samplers[3] = {bg, sp, wv};
    CK(sfb_render_frame(ctx, SFB_SCENE_VISUALIZER, &u, samplers, 3, 0, 3840, 2160, 2, 2, 3, out));
    CK(sfb_sync(ctx));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a);
    for (int i = 0; i < 5; i++) CK(sfb_render_frame(ctx, SFB_SCENE_VISUALIZER, &u, samplers, 3, 0, 3840, 2160, 2, 2, 3, out));
    cudaEventRecord(b); CK(sfb_sync(ctx)); float ms; cudaEventElapsedTime(&ms, a, b);
    printf("C harness ok: %.3f ms/frame\n", ms/5);
    return 0;
}

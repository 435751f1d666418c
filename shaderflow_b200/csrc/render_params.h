// render_params.h — the kernel parameter block of every per-pixel kernel, shared between the translation units of
// libsfb200.so and the run-time compiled programs (jit/: NVRTC sees this file as an in-memory header, so it includes
// nothing; whoever includes it has included include/sfb200.h before).
#pragma once

// What a kernel sees of a texture (passed by value inside the kernel parameter block)
struct DevSampler {
    unsigned long long hw;   // cudaTextureObject_t, 0 when only the linear mirror is valid
    const void* lin;         // [h][w][padded] texels, tightly packed
    int w, h;
    int padded;              // components as stored: 1, 2 or 4
    int comps;               // components the user declared (3 → alpha reads 1)
    int dtype;               // SFB_DTYPE_*
    int filter;              // SFB_FILTER_NEAREST / LINEAR
    int rx, ry;              // repeat (1) or clamp-to-edge (0)
};

struct RenderParams {
    sfb_uniforms u;
    DevSampler tex[SFB_MAX_SAMPLERS];
    int Wr, Hr;              // fragments of the iScreen target (render resolution)
    int W, H;                // final resolution
    int ssaa, subsample, comps;
    double inv_Wr, inv_Hr;   // 1/Wr, 1/Hr
    unsigned char* dst;
    int dst_dtype, dst_padded;   // format of `dst` for sfb_render_target (SFB_DTYPE_*, stored components)
    float* dst_f32;
    int fast;                // scene-specific fast path allowed (set by the launcher after checking formats)
};

// render_kernels.cuh — the templated per-scene kernels (K3 screen pass, fused K3+K4 per pixel, fused K3+K4 with one lane
// per sub-sample) and their dispatch over the scene registry. Each family is instantiated in its own translation unit
// (render_screen.cu / render_frame.cu / render_lanes.cu: 17 scenes x 2 filters each), so the library builds in parallel.
// c_blur (the visualizer's tap table, __constant__ memory) exists once per translation unit: every unit that can run
// the visualizer's table-driven path fills its own copy through build_blur_table().
#pragma once
#include "scenes.cuh"

namespace sfb_render {
using namespace glsl;

// Colour store to an 8-bit attachment: clamp to [0,1] (NaN → 0), round half to even
SFB_DEV unsigned int to_unorm8(float c) { return (unsigned int)__float2int_rn(__saturatef(c)*255.0f); }

// ------------------------------------------------------------------------------------------------
// K3: one thread per fragment of the Wr×Hr RGBA8 target

template <int SCENE, bool HW>
__global__ void __launch_bounds__(256) screen_kernel(const __grid_constant__ RenderParams P) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    const int j = blockIdx.y*blockDim.y + threadIdx.y;
    if (i >= P.Wr || j >= P.Hr) return;
    const Frag f = make_frag(P, i, j);
    const vec4 c = shade<SCENE, HW>(P, f);
    const size_t idx = size_t(j)*size_t(P.Wr) + size_t(i);
    if (P.dst_dtype == SFB_DTYPE_U8 && P.dst_padded == 4) {
        reinterpret_cast<uchar4*>(P.dst)[idx] =
            make_uchar4(to_unorm8(c.x), to_unorm8(c.y), to_unorm8(c.z), to_unorm8(c.w));
    } else {
        // a texture of another format as the colour attachment (sfb_render_target): its stored components
        const float comp[4] = {c.x, c.y, c.z, c.w};
        for (int k = 0; k < P.dst_padded; k++) {
            if (P.dst_dtype == SFB_DTYPE_U8)       P.dst[idx*size_t(P.dst_padded) + k] = (unsigned char)to_unorm8(comp[k]);
            else if (P.dst_dtype == SFB_DTYPE_F32) reinterpret_cast<float*>(P.dst)[idx*size_t(P.dst_padded) + k] = comp[k];
            else                                   reinterpret_cast<__half*>(P.dst)[idx*size_t(P.dst_padded) + k] = __float2half_rn(comp[k]);
        }
    }
    if (P.dst_f32) reinterpret_cast<float4*>(P.dst_f32)[idx] = make_float4(c.x, c.y, c.z, c.w);
}

// ------------------------------------------------------------------------------------------------
// Fused K3+K4: one thread per OUTPUT pixel, ssaa×ssaa shaded sub-samples, each quantised to 8 bit
// as the RGBA8 iScreen store would, integer box sum, one more 8-bit store. Tile = 32×8 pixels;
// rgb24 rows are staged in shared memory and leave as 32-bit words.

constexpr int TILE_X = 32, TILE_Y = 8;

template <int SCENE, bool HW>
__global__ void __launch_bounds__(TILE_X*TILE_Y) frame_kernel(const __grid_constant__ RenderParams P) {
    __shared__ unsigned int stage[TILE_Y][TILE_X];     // rgb24: first 24 words of each row are used
    const int x = blockIdx.x*TILE_X + threadIdx.x;
    const int y = blockIdx.y*TILE_Y + threadIdx.y;
    const bool inside = (x < P.W) && (y < P.H);
    const int S = P.ssaa;
    unsigned int r = 0, g = 0, b = 0;
    if (inside) {
        for (int sy = 0; sy < S; sy++) {
            for (int sx = 0; sx < S; sx++) {
                const Frag f = make_frag(P, x*S + sx, y*S + sy);
                const vec4 c = shade<SCENE, HW>(P, f);
                if (P.dst_f32)                                  // parity probe: fragColor before the 8-bit store
                    reinterpret_cast<float4*>(P.dst_f32)[size_t(y*S + sy)*size_t(P.Wr) + size_t(x*S + sx)] = make_float4(c.x, c.y, c.z, c.w);
                r += to_unorm8(c.x); g += to_unorm8(c.y); b += to_unorm8(c.z);
            }
        }
        // mean of S² bytes, then the RGB8 store of iFinal: round(mean) half-to-even
        const float inv = 1.0f/float(S*S);
        r = (unsigned int)__float2int_rn(float(r)*inv);
        g = (unsigned int)__float2int_rn(float(g)*inv);
        b = (unsigned int)__float2int_rn(float(b)*inv);
    }
    if (P.comps == 4) {
        if (inside) reinterpret_cast<uchar4*>(P.dst)[size_t(y)*size_t(P.W) + size_t(x)] = make_uchar4(r, g, b, 255);
        return;
    }
    const bool words = (P.W % 4 == 0) && (blockIdx.x*TILE_X + TILE_X <= P.W);
    if (!words) {
        if (inside) {
            unsigned char* p = P.dst + (size_t(y)*size_t(P.W) + size_t(x))*3;
            p[0] = r; p[1] = g; p[2] = b;
        }
        return;
    }
    unsigned char* row = reinterpret_cast<unsigned char*>(stage[threadIdx.y]);
    row[threadIdx.x*3 + 0] = r; row[threadIdx.x*3 + 1] = g; row[threadIdx.x*3 + 2] = b;
    __syncwarp();
    if (threadIdx.x < (TILE_X*3)/4 && y < P.H) {
        unsigned int* out = reinterpret_cast<unsigned int*>(P.dst + (size_t(y)*size_t(P.W) + size_t(blockIdx.x*TILE_X))*3);
        out[threadIdx.x] = stage[threadIdx.y][threadIdx.x];
    }
}

// ------------------------------------------------------------------------------------------------
// Fused K3+K4 for the ALU-bound scenes (fractals, ray marching): one LANE per shaded sub-sample instead of one
// thread looping over the S² sub-samples of its pixel. A warp shades a compact block of sub-samples — 32/S²
// neighbouring output pixels — so lanes that iterate for long (escape-time loops, ray marches) sit next to each
// other and the warp no longer waits S² times for its slowest lane; the 8-bit box sum is a shuffle reduction.
// CTA = 8 warps side by side on one output row; the rgb24 segment leaves through shared memory as words.

template <int SCENE, bool HW, int S>
__global__ void __launch_bounds__(256) frame_lanes_kernel(const __grid_constant__ RenderParams P) {
    constexpr int SS = S*S, PPW = 32/SS, NPIX = 8*PPW;             // lanes per pixel, pixels per warp, pixels per CTA
    static_assert(S == 2 || S == 4, "a warp holds whole pixels");
    __shared__ unsigned int stage[(NPIX*3 + 3)/4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % SS, sx = sub % S, sy = sub / S;
    const int x = blockIdx.x*NPIX + warp*PPW + lane/SS, y = blockIdx.y;
    const bool inside = x < P.W;
    unsigned int r = 0, g = 0, b = 0;
    if (inside) {
        const Frag f = make_frag(P, x*S + sx, y*S + sy);
        const vec4 c = shade<SCENE, HW>(P, f);
        if (P.dst_f32)
            reinterpret_cast<float4*>(P.dst_f32)[size_t(y*S + sy)*size_t(P.Wr) + size_t(x*S + sx)] = make_float4(c.x, c.y, c.z, c.w);
        r = to_unorm8(c.x); g = to_unorm8(c.y); b = to_unorm8(c.z);
    }
    #pragma unroll
    for (int o = 1; o < SS; o <<= 1) {
        r += __shfl_xor_sync(0xffffffffu, r, o); g += __shfl_xor_sync(0xffffffffu, g, o); b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    const float inv = 1.0f/float(SS);
    r = (unsigned int)__float2int_rn(float(r)*inv);
    g = (unsigned int)__float2int_rn(float(g)*inv);
    b = (unsigned int)__float2int_rn(float(b)*inv);
    const bool writer = inside && sub == 0;
    if (P.comps == 4) {
        if (writer) reinterpret_cast<uchar4*>(P.dst)[size_t(y)*size_t(P.W) + size_t(x)] = make_uchar4(r, g, b, 255);
        return;
    }
    const bool words = (P.W % 4 == 0) && (int(blockIdx.x)*NPIX + NPIX <= P.W);
    if (!words) {
        if (writer) { unsigned char* p = P.dst + (size_t(y)*size_t(P.W) + size_t(x))*3; p[0] = r; p[1] = g; p[2] = b; }
        return;
    }
    if (writer) {
        unsigned char* p = reinterpret_cast<unsigned char*>(stage) + (warp*PPW + lane/SS)*3;
        p[0] = r; p[1] = g; p[2] = b;
    }
    __syncthreads();
    if (threadIdx.x < (NPIX*3)/4) {
        unsigned int* out = reinterpret_cast<unsigned int*>(P.dst + (size_t(y)*size_t(P.W) + size_t(blockIdx.x)*NPIX)*3);
        out[threadIdx.x] = stage[threadIdx.x];
    }
}

// visualizer.frag:26-27 evaluated in strict float32 on the host: 9 angles x 10 walks, stored as
// dir*walk (the float counters make it 9 directions, not 8)
static inline int build_blur_table() {
    static bool done[64] = {};
    int device = 0;
    SFB_CUDA(cudaGetDevice(&device));
    if (device < 64 && done[device]) return SFB_OK;
    BlurTable table;
    const volatile float TAU_F = 6.2831853071795864f, directions = 8.0f, quality = 10.0f;
    int n = 0;
    for (volatile float angle = 0.0f; angle < TAU_F; angle = angle + TAU_F/directions) {
        const float c = cosf(angle), s = sinf(angle);
        for (volatile float walk = 1.0f/quality; walk <= 1.001f; walk = walk + 1.0f/quality) {
            if (n >= 90) SFB_FAIL(SFB_ESTATE, "blur table overflow: float loop semantics changed");
            table.tap[n++] = make_float2(c*walk, s*walk);
        }
    }
    if (n != 90) SFB_FAIL(SFB_ESTATE, "blur table has %d taps, expected 90", n);
    table.tap[90] = make_float2(0.0f, 0.0f); table.tap[91] = make_float2(0.0f, 0.0f);
    SFB_CUDA(cudaMemcpyToSymbol(c_blur, &table, sizeof(table)));
    if (device < 64) done[device] = true;
    return SFB_OK;
}

template <template <int, bool> class Launch, typename... A>
static inline void dispatch(int scene, bool hw, A... a) {
    #define SFB_CASE(S) case S: hw ? Launch<S, true>::run(a...) : Launch<S, false>::run(a...); break;
    switch (scene) {
        SFB_CASE(SFB_SCENE_DEFAULT) SFB_CASE(SFB_SCENE_SHADERTOY) SFB_CASE(SFB_SCENE_VISUALIZER)
        SFB_CASE(SFB_SCENE_BARS) SFB_CASE(SFB_SCENE_WAVEFORM) SFB_CASE(SFB_SCENE_MANDELBROT)
        SFB_CASE(SFB_SCENE_TETRATION) SFB_CASE(SFB_SCENE_RAYMARCH)
        SFB_CASE(SFB_SCENE_MULTISHADER_CHILD) SFB_CASE(SFB_SCENE_MULTISHADER) SFB_CASE(SFB_SCENE_MULTIPASS)
        SFB_CASE(SFB_SCENE_MOTIONBLUR) SFB_CASE(SFB_SCENE_DYNAMICS) SFB_CASE(SFB_SCENE_AUDIO)
        SFB_CASE(SFB_SCENE_LIFE_SIMULATION) SFB_CASE(SFB_SCENE_LIFE_VISUALS) SFB_CASE(SFB_SCENE_PIANO)
    }
    #undef SFB_CASE
}

template <int S, bool HW> struct LaunchScreen {
    static void run(const RenderParams& P, cudaStream_t st) {
        dim3 block(32, 8), grid((P.Wr + 31)/32, (P.Hr + 7)/8);
        screen_kernel<S, HW><<<grid, block, 0, st>>>(P);
    }
};
template <int S, bool HW> struct LaunchFrame {
    static void run(const RenderParams& P, cudaStream_t st) {
        dim3 block(TILE_X, TILE_Y), grid((P.W + TILE_X - 1)/TILE_X, (P.H + TILE_Y - 1)/TILE_Y);
        frame_kernel<S, HW><<<grid, block, 0, st>>>(P);
    }
};
// the ALU-bound scenes at ssaa 2 / 4: one lane per sub-sample
template <int SCENE, bool HW> static inline void launch_frame_lanes(const RenderParams& P, cudaStream_t st) {
    if (P.ssaa == 4) { dim3 grid((P.W + 15)/16, P.H); frame_lanes_kernel<SCENE, HW, 4><<<grid, 256, 0, st>>>(P); }
    else             { dim3 grid((P.W + 63)/64, P.H); frame_lanes_kernel<SCENE, HW, 2><<<grid, 256, 0, st>>>(P); }
}


}  // namespace sfb_render

// Launch entry points of the three units (called by render.cu)
int sfb_launch_screen(int scene, bool hardware_filter, const glsl::RenderParams& P, cudaStream_t stream);
int sfb_launch_frame(int scene, bool hardware_filter, const glsl::RenderParams& P, cudaStream_t stream);
// false when the scene / ssaa has no lane-per-sub-sample variant
bool sfb_launch_frame_lanes(int scene, bool hardware_filter, const glsl::RenderParams& P, cudaStream_t stream);

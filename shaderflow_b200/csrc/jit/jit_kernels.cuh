// jit_kernels.cuh — the two kernels of a run-time compiled program, around the `g::Shader` the translator emitted.
// Same launch shapes, stores and 8-bit rules as screen_kernel / frame_kernel of render_kernels.cuh (which run the
// ahead-of-time scenes), so sfb_render_screen / sfb_render_target / sfb_render_frame treat both kinds alike.
#pragma once

namespace g {

G_DEV unsigned to_unorm8(float c) { return (unsigned)__float2int_rn(__saturatef(c)*255.0f); }
G_DEV unsigned short float_to_half_bits(float f) { unsigned short h; asm("cvt.rn.f16.f32 %0, %1;" : "=h"(h) : "f"(f)); return h; }

// thread -> pixel inside the 32 x 8 tile of a CTA: a warp shades 16 x 2 pixels as eight 2 x 2 quads
// (lane = 4*quad + 2*dy + dx), so that dFdx / dFdy / fwidth are differences between lanes (xor 1 / xor 2), as between the
// fragments of a rasteriser's quad
G_DEV void tile_pixel(int& px, int& py) {
    const int lane = threadIdx.x, warp = threadIdx.y, quad = lane >> 2;
    px = (warp & 1)*16 + quad*2 + (lane & 1);
    py = (warp >> 1)*2 + ((lane >> 1) & 1);
}

G_DEV vec4 shade(const RenderParams& P, int i, int j, float derivative_scale) {
    Shader s(P, i, j);
    s.sfb_dscale = derivative_scale;
    s.main();
    return s.sfb_discarded ? vec4(0.0f) : s.fragColor;          // a discarded fragment keeps the cleared target
}

}  // namespace g

extern "C" __global__ void __launch_bounds__(256) sfb_jit_screen(const __grid_constant__ RenderParams P) {
    int px, py;
    g::tile_pixel(px, py);
    const int i = blockIdx.x*32 + px, j = blockIdx.y*8 + py;
    if (i >= P.Wr || j >= P.Hr) return;
    const g::vec4 c = g::shade(P, i, j, 1.0f);
    const size_t idx = size_t(j)*size_t(P.Wr) + size_t(i);
    for (int k = 0; k < P.dst_padded; k++) {
        const size_t at = idx*size_t(P.dst_padded) + k;
        if (P.dst_dtype == SFB_DTYPE_U8)       P.dst[at] = (unsigned char)g::to_unorm8(c.v[k]);
        else if (P.dst_dtype == SFB_DTYPE_F32) reinterpret_cast<float*>(P.dst)[at] = c.v[k];
        else                                   reinterpret_cast<unsigned short*>(P.dst)[at] = g::float_to_half_bits(c.v[k]);
    }
    if (P.dst_f32) reinterpret_cast<float4*>(P.dst_f32)[idx] = make_float4(c.x, c.y, c.z, c.w);
}

// fused iScreen + final pass: ssaa x ssaa sub-samples per output pixel, each quantised like the RGBA8 iScreen store,
// box mean, one more 8-bit store; rgb24 rows leave through shared memory as 32-bit words
extern "C" __global__ void __launch_bounds__(256) sfb_jit_frame(const __grid_constant__ RenderParams P) {
    constexpr int TILE_X = 32, TILE_Y = 8;
    __shared__ unsigned int stage[TILE_Y][TILE_X];
    int px, py;
    g::tile_pixel(px, py);
    const int x = blockIdx.x*TILE_X + px, y = blockIdx.y*TILE_Y + py;
    const bool inside = (x < P.W) && (y < P.H);
    const int S = P.ssaa;
    unsigned r = 0, gr = 0, b = 0;
    if (inside) {
        for (int sy = 0; sy < S; sy++)
            for (int sx = 0; sx < S; sx++) {
                // the lane next door shades the same sub-sample of the next output pixel, S fragments away
                const g::vec4 c = g::shade(P, x*S + sx, y*S + sy, 1.0f/float(S));
                if (P.dst_f32)
                    reinterpret_cast<float4*>(P.dst_f32)[size_t(y*S + sy)*size_t(P.Wr) + size_t(x*S + sx)] = make_float4(c.x, c.y, c.z, c.w);
                r += g::to_unorm8(c.x); gr += g::to_unorm8(c.y); b += g::to_unorm8(c.z);
            }
        const float inv = 1.0f/float(S*S);
        r = (unsigned)__float2int_rn(float(r)*inv); gr = (unsigned)__float2int_rn(float(gr)*inv); b = (unsigned)__float2int_rn(float(b)*inv);
    }
    if (P.comps == 4) {
        if (inside) reinterpret_cast<uchar4*>(P.dst)[size_t(y)*size_t(P.W) + size_t(x)] = make_uchar4(r, gr, b, 255);
        return;
    }
    const bool words = (P.W % 4 == 0) && (blockIdx.x*TILE_X + TILE_X <= P.W);
    if (!words) {
        if (inside) { unsigned char* p = P.dst + (size_t(y)*size_t(P.W) + size_t(x))*3; p[0] = r; p[1] = gr; p[2] = b; }
        return;
    }
    unsigned char* row = reinterpret_cast<unsigned char*>(stage[py]);
    row[px*3 + 0] = r; row[px*3 + 1] = gr; row[px*3 + 2] = b;
    __syncthreads();                                           // a tile row is shaded by two warps
    const int oy = blockIdx.y*TILE_Y + threadIdx.y;
    if (threadIdx.x < (TILE_X*3)/4 && oy < P.H) {
        unsigned int* out = reinterpret_cast<unsigned int*>(P.dst + (size_t(oy)*size_t(P.W) + size_t(blockIdx.x*TILE_X))*3);
        out[threadIdx.x] = stage[threadIdx.y][threadIdx.x];
    }
}

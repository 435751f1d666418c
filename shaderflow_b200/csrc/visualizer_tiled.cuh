// visualizer_tiled.cuh — per-pixel kernel of the headline scene (examples/basic/shaders/visualizer.frag),
// fused with the SSAA downsample (fragment/final.glsl). One CTA shades a 32x8 tile of OUTPUT pixels
// (S x S fragments each). It works for any camera; exports with the axis-aligned 2D camera and a fine enough
// texel step go through the separable kernel instead (visualizer_rows.cu, 4.4x faster), this one takes the
// rest: rotated / stereo / equirectangular cameras, ssaa 3, coarse steps, and the iScreen pass of unfused
// exports (one fragment per thread, RGBA8 store with fragColor.a).
//
// What dominates visualizer.frag is the blur loop (:19-33): 1 + 90 bilinear taps of the background per
// fragment, all inside a disc of `intensity` (<= 0.003 of the image height, i.e. <= 3.3 texels) around
// the fragment's own texture position. So the whole CTA reads one small window of the background:
//   A. every thread computes the texel-space position of its fragments' centre tap; a block-wide
//      min/max gives the window (origin x0,y0; <= 64 x 32 texels);
//   B. the window is staged ONCE into shared memory, already wrapped (REPEAT / CLAMP_TO_EDGE resolved
//      per texel) and already widened from RGBA8 to float — by a TMA 2D tile load of the raw texels
//      when the window lies inside the texture (cp.async.bulk.tensor, UTMALDG in SASS; the box must start
//      on a 16-byte boundary, so x0 is a multiple of 4 texels) followed by an
//      in-place widen, or by plain loads when it crosses an edge;
//   C. the taps per fragment then cost: 2 FFMA for the position (tap table in the constant bank), a
//      magic-constant floor (no F2I/FRND on the quarter-rate XU pipe), 2 LDS.128 + 2 LDS.64, 5 ops for
//      the four weights and 12 FFMA — no integer modulo, no byte unpacking, no global loads.
//      The kernel is bound by shared-memory wavefronts (ncu: 87 % of peak with one float4 per texel), so
//      the window is stored as horizontal PAIR records — rg[y][x] = (r,g) of texels x and x+1 (float4),
//      bb[y][x] = b of texels x and x+1 (float2) — which serves a bilinear footprint in 12 wavefronts
//      instead of 16. The 9th direction of the float loop (angle 8*(TAU/8) = 6.2831850 < TAU) coincides
//      with the first to 1e-6 texel, below the float32 resolution of the tap position itself, so ray 0 is
//      evaluated once and weighted twice: 81 footprints stand for the 91 taps;
//   D. the rest of main() per fragment, the RGBA8 store rule per sub-sample, the integer box sum and
//      the rgb24 store (staged through shared memory so rows leave as 32-bit words).
// Arithmetic differs from the literal transliteration (scenes.cuh: scene_visualizer) only by float32
// re-association: positions are advanced in texel space, and the 1/255 of the unorm decode is applied
// to the tap sum. tests/test_gpu_render.py gates this kernel against the literal one at 2e-5 and
// against the oracle at the north_star tolerance.
#pragma once
#include <cuda.h>
#include "scenes.cuh"

namespace glsl {

constexpr int VT_TILE_X = 32, VT_TILE_Y = 8;           // output pixels per CTA
constexpr int VT_WIN_W = 64, VT_WIN_H = 32;            // background window capacity, texels (row stride 64)
constexpr int VT_TMA_W = 64, VT_TMA_H = 32;            // TMA box (texels)

struct VisualizerParams {
    RenderParams R;
    const CUtensorMap* tmap;                           // device copy of the 2D uint32 view of the background's linear mirror
    int use_tma;
    int screen_alpha;                                  // iScreen pass (sfb_render_screen): store fragColor.a, not 255
};

// --- mbarrier / TMA helpers (inline PTX, sm_90+) ---------------------------------------------------
SFB_DEV unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
SFB_DEV void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
SFB_DEV void mbar_expect_tx(unsigned long long* bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
SFB_DEV void mbar_wait(unsigned long long* bar, unsigned int phase) {
    unsigned int done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    } while (!done);
}
SFB_DEV void tma_load_2d(void* dst, const void* tmap, unsigned long long* bar, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :: "r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}

// Per-fragment front end of main(): camera + background_uv → centre tap position in texel space
// (u*W - 0.5, v*H - 0.5), shaderflow.glsl:165-169,202-204 applied to visualizer.frag:16-18
struct VisFrag { Frag f; vec2 uv; vec2 tap; bool oob; };

SFB_HD VisFrag vis_front(const RenderParams& P, int i, int j, float fw, float fh, float hw, vec2 wobble, float zf) {
    VisFrag v;
    v.f = make_frag(P, i, j);
    Camera cam = get_camera(P.u, v.f);
    v.uv = cam.gluv; v.oob = cam.out_of_bounds;
    vec2 st = zoom(gluv2stuv(v.uv), zf, mk2(0.5f)) + wobble;
    const float ux = ((st.x*2.0f - 1.0f)*hw + 1.0f)/2.0f;
    const float uy = ((st.y*2.0f - 1.0f) + 1.0f)/2.0f;
    v.tap = mk2(ux*fw - 0.5f, uy*fh - 0.5f);
    return v;
}

// Everything of main() after the blur loop (visualizer.frag:35-73); `rgb` = blurred background
SFB_DEV vec4 vis_back(const RenderParams& P, const VisFrag& v, vec3 rgb) {
    const float iAudioVolume = P.u.extra[0][0], iAudioSTD = P.u.extra[1][0];
    const Frag& f = v.f; const vec2 uv = v.uv;
    const vec3 space = mk3(1.0f, 11.0f, 26.0f)/255.0f;
    rgb = rgb*(1.0f + 5.0f*iAudioSTD*powf(clamp(length(f.agluv) - 0.3f, 0.0f, 1.0f), 6.0f));
    const float c = -4.37113883e-08f, sn = -1.0f;                 // cos, sin of float32(-PI/2)
    vec2 music_uv = mk2(c*uv.x + sn*uv.y, (-sn)*uv.x + c*uv.y);
    music_uv = music_uv*(1.0f - 0.4f*powf(fabsf(iAudioVolume), 0.5f));
    const float radius = 0.17f;
    const float circle = fabsf(atan1n(music_uv));
    vec4 s = texture<false>(P.tex[1], mk2(0.0f, circle));
    vec2 freq = mk2(sqrtf(s.x/1000.0f), sqrtf(s.y/1000.0f));
    freq = freq*(0.05f + 3.0f*smoothstep(0.0f, 2.0f, circle));
    const float lm = length(music_uv);
    if (lm < radius) {
        rgb = rgb*0.5f;
    } else {
        const float bar = (music_uv.y < 0.0f) ? freq.x : freq.y;
        const float r = radius + 0.5f*bar;
        if (lm < r) rgb = mix(rgb, mk3(1.0f), smoothstep(0.0f, 1.0f, 0.5f + bar));
        else        rgb = rgb*powf((lm - r)*0.5f, 0.05f);
    }
    rgb = mix(rgb, space, smoothstep(0.0f, 1.0f, length(uv)/20.0f));
    vec2 vig = f.astuv*(1.0f - yx(f.astuv));
    rgb = rgb*powf(vig.x*vig.y*20.0f, 0.1f + 0.15f*iAudioVolume);
    vec4 fragColor = mk4(rgb, 1.0f);
    vec4 w = texture<false>(P.tex[2], mk2(f.astuv.x, 0.0f));
    if (1.0f - f.gluv.y < 0.2f*w.x) fragColor = fragColor*0.8f;
    if (1.0f + f.gluv.y < 0.2f*w.y) fragColor = fragColor*0.8f;
    return fragColor;
}

constexpr int VT_WIN = VT_WIN_W*VT_WIN_H;
constexpr size_t VT_SMEM = sizeof(float4)*VT_WIN + sizeof(float2)*VT_WIN;     // rg pairs + b pairs = 48 KB (dynamic)

SFB_DEV vec3 widen3(unsigned int word) {
    return mk3(__uint_as_float(__byte_perm(word, 0x4B000000u, 0x7540)) - 8388608.0f,
               __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7541)) - 8388608.0f,
               __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7542)) - 8388608.0f);
}

template <int S>
__global__ void __launch_bounds__(VT_TILE_X*VT_TILE_Y, 4)
visualizer_tiled_kernel(const __grid_constant__ VisualizerParams VP) {
    const RenderParams& P = VP.R;
    extern __shared__ __align__(128) unsigned char vt_smem[];
    float4* rg = reinterpret_cast<float4*>(vt_smem);                        // [VT_WIN_H][VT_WIN_W] (r,g) of texels x, x+1
    float2* bb = reinterpret_cast<float2*>(vt_smem + sizeof(float4)*VT_WIN);// [VT_WIN_H][VT_WIN_W] b of texels x, x+1
    __shared__ unsigned int stage[VT_TILE_Y][VT_TILE_X];
    __shared__ float red[4][VT_TILE_Y];
    __shared__ int win[4];                                                  // x0, y0, fits, tma
    __shared__ __align__(8) unsigned long long bar;

    const int tid = threadIdx.y*VT_TILE_X + threadIdx.x;
    const int x = blockIdx.x*VT_TILE_X + threadIdx.x, y = blockIdx.y*VT_TILE_Y + threadIdx.y;
    const bool inside = (x < P.W) && (y < P.H);
    const DevSampler& bg = P.tex[0];
    const float fw = float(bg.w), fh = float(bg.h), hw = float(bg.h)/float(bg.w);
    const float iTime = P.u.iTime, iAudioVolume = P.u.extra[0][0];
    const float zf = 0.95f + 0.01f*sinf(iTime) - 0.02f*iAudioVolume - 0.03f;
    const vec2 wobble = 0.005f*mk2(cosf(iTime*3.25135f), sinf(iTime*1.153469f));
    const float intensity = 0.01f*clamp(powf(iAudioVolume, 2.5f), 0.0f, 0.3f);
    const float scale = intensity*fh;                         // st displacement → texels (both axes: hw*fw == fh)

    // ---- A. window of the background this CTA touches ------------------------------------------
    vec2 tap[S*S];
    float lo_x = 3.0e38f, lo_y = 3.0e38f, hi_x = -3.0e38f, hi_y = -3.0e38f;
    #pragma unroll
    for (int s = 0; s < S*S; s++) {
        // clamp the coordinates of out-of-range threads onto the target so they do not widen the window
        const int px = min(x, P.W - 1), py = min(y, P.H - 1);
        tap[s] = vis_front(P, px*S + (s % S), py*S + (s / S), fw, fh, hw, wobble, zf).tap;
        lo_x = fminf(lo_x, tap[s].x); hi_x = fmaxf(hi_x, tap[s].x);
        lo_y = fminf(lo_y, tap[s].y); hi_y = fmaxf(hi_y, tap[s].y);
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo_x = fminf(lo_x, __shfl_xor_sync(0xffffffffu, lo_x, o)); hi_x = fmaxf(hi_x, __shfl_xor_sync(0xffffffffu, hi_x, o));
        lo_y = fminf(lo_y, __shfl_xor_sync(0xffffffffu, lo_y, o)); hi_y = fmaxf(hi_y, __shfl_xor_sync(0xffffffffu, hi_y, o));
    }
    if (threadIdx.x == 0) { red[0][threadIdx.y] = lo_x; red[1][threadIdx.y] = lo_y; red[2][threadIdx.y] = hi_x; red[3][threadIdx.y] = hi_y; }
    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (tid == 0) {
        float ax = red[0][0], ay = red[1][0], bx = red[2][0], by = red[3][0];
        for (int r = 1; r < VT_TILE_Y; r++) { ax = fminf(ax, red[0][r]); ay = fminf(ay, red[1][r]); bx = fmaxf(bx, red[2][r]); by = fmaxf(by, red[3][r]); }
        const float reach = scale*1.0001f + 1.0f;             // |dir*walk| <= 1.0000001; +1 keeps local coords >= 1
        // TMA needs the box to start on a 16-byte boundary of the innermost dimension: x0 % 4 == 0
        const int x0 = (int(floorf(ax - reach)) - 1) & ~3, y0 = int(floorf(ay - reach)) - 1;
        const int x1 = int(floorf(bx + reach)) + 2, y1 = int(floorf(by + reach)) + 2;   // last texel touched + 1
        const bool finite = (ax == ax) && (bx == bx) && (ay == ay) && (by == by) && fabsf(ax) < 1.0e9f && fabsf(bx) < 1.0e9f
                         && fabsf(ay) < 1.0e9f && fabsf(by) < 1.0e9f;
        win[0] = x0; win[1] = y0;
        // pair records need texel x+1 too: the last column of the window only ever serves as a neighbour
        win[2] = (finite && (x1 - x0) <= VT_WIN_W - 1 && (y1 - y0) <= VT_WIN_H) ? 1 : 0;
        win[3] = (win[2] && VP.use_tma && x0 >= 0 && y0 >= 0 && x0 + VT_TMA_W <= bg.w && y0 + VT_TMA_H <= bg.h) ? 1 : 0;
        if (win[3]) {
            // ---- B1. TMA: the raw RGBA8 box lands in the FIRST 8 KB of the window buffer -------------
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&bar, VT_TMA_W*VT_TMA_H*4);
            tma_load_2d(vt_smem, VP.tmap, &bar, x0, y0);
        }
    }
    __syncthreads();
    const int x0 = win[0], y0 = win[1];
    const bool fits = win[2] != 0;

    if (fits) {
        constexpr int PER = VT_WIN/(VT_TILE_X*VT_TILE_Y);                  // window texels per thread (8)
        unsigned int w0[PER], w1[PER];                                      // texel t and its right neighbour
        if (win[3]) {
            mbar_wait(&bar, 0);
            const unsigned int* raw = reinterpret_cast<const unsigned int*>(vt_smem);
            #pragma unroll
            for (int k = 0; k < PER; k++) {
                const int t = tid + k*VT_TILE_X*VT_TILE_Y;
                w0[k] = raw[t];
                w1[k] = raw[((t % VT_WIN_W) == VT_WIN_W - 1) ? t : t + 1];
            }
            __syncthreads();                                                 // raw box fully read before it is overwritten
        } else {
            // ---- B2. edge windows: per-texel wrap, plain loads -----------------------------------
            const unsigned int* texels = reinterpret_cast<const unsigned int*>(bg.lin);
            #pragma unroll
            for (int k = 0; k < PER; k++) {
                const int t = tid + k*VT_TILE_X*VT_TILE_Y;
                const int gy = wrap_index(y0 + (t / VT_WIN_W), bg.h, bg.ry);
                const unsigned int* row = texels + size_t(gy)*size_t(bg.w);
                w0[k] = __ldg(row + wrap_index(x0 + (t % VT_WIN_W), bg.w, bg.rx));
                w1[k] = __ldg(row + wrap_index(x0 + (t % VT_WIN_W) + 1, bg.w, bg.rx));
            }
        }
        #pragma unroll
        for (int k = 0; k < PER; k++) {
            const int t = tid + k*VT_TILE_X*VT_TILE_Y;
            const vec3 a = widen3(w0[k]), b = widen3(w1[k]);
            rg[t] = make_float4(a.x, a.y, b.x, b.y);
            bb[t] = make_float2(a.z, b.z);
        }
    }
    __syncthreads();

    // ---- C + D. shade ----------------------------------------------------------------------------
    unsigned int r8 = 0, g8 = 0, b8 = 0, a8 = 0;
    if (inside) {
        #pragma unroll
        for (int s = 0; s < S*S; s++) {
            vec4 c;
            // the per-fragment state is cheap next to the taps: recompute it instead of keeping it live
            const VisFrag v = vis_front(P, x*S + (s % S), y*S + (s / S), fw, fh, hw, wobble, zf);
            if (v.oob) {
                c = mk4(mk3(1.0f, 11.0f, 26.0f)/255.0f, 0.0f);
            } else if (!fits) {
                c = scene_visualizer_fast(P, v.f);                     // window too large: per-tap global path
            } else {
                // tile-local centre, >= 1 by construction of x0/y0
                const float cx = tap[s].x - float(x0), cy = tap[s].y - float(y0);
                const char* rg_base = reinterpret_cast<const char*>(rg);
                const char* bb_base = reinterpret_cast<const char*>(bb);
                float ar = 0.0f, ag = 0.0f, ab = 0.0f;
                auto footprint = [&](float dx, float dy) {
                    const float px = fmaf(dx, scale, cx), py = fmaf(dy, scale, cy);
                    // floor for 0.5 <= p < 2^22: p + (2^23 - 0.5) rounds to the integer floor(p) + 2^23
                    // (ties land on either neighbour; bilinear interpolation is continuous there)
                    const float tx = px + 8388607.5f, ty = py + 8388607.5f;
                    const float a = px - (tx - 8388608.0f), b = py - (ty - 8388608.0f);
                    // 8*(iy*64 + ix), modulo 2^32: the 2^23 offsets of the magic constant fold into one constant
                    const unsigned int off8 = ((unsigned int)__float_as_int(ty) << 9) + ((unsigned int)__float_as_int(tx) << 3)
                                            - ((0x4B000000u << 9) + (0x4B000000u << 3));
                    const float4 r0 = *reinterpret_cast<const float4*>(rg_base + 2u*off8);
                    const float4 r1 = *reinterpret_cast<const float4*>(rg_base + 2u*off8 + VT_WIN_W*16);
                    const float2 b0 = *reinterpret_cast<const float2*>(bb_base + off8);
                    const float2 b1 = *reinterpret_cast<const float2*>(bb_base + off8 + VT_WIN_W*8);
                    const float w11 = a*b, w10 = a - w11, w01 = b - w11, w00 = (1.0f - a) - w01;
                    ar = fmaf(w00, r0.x, ar); ag = fmaf(w00, r0.y, ag); ab = fmaf(w00, b0.x, ab);
                    ar = fmaf(w10, r0.z, ar); ag = fmaf(w10, r0.w, ag); ab = fmaf(w10, b0.y, ab);
                    ar = fmaf(w01, r1.x, ar); ag = fmaf(w01, r1.y, ag); ab = fmaf(w01, b1.x, ab);
                    ar = fmaf(w11, r1.z, ar); ag = fmaf(w11, r1.w, ag); ab = fmaf(w11, b1.y, ab);
                };
                // ray 0 (angle 0) stands for itself and for the 9th float-loop direction: weight 2
                #pragma unroll
                for (int t = 0; t < 10; t++) footprint(c_blur.tap[t].x, c_blur.tap[t].y);
                ar += ar; ag += ag; ab += ab;
                #pragma unroll 1
                for (int ray = 1; ray < 8; ray++) {
                    #pragma unroll
                    for (int t = 0; t < 10; t++) footprint(c_blur.tap[ray*10 + t].x, c_blur.tap[ray*10 + t].y);
                }
                footprint(0.0f, 0.0f);                                  // the undisplaced first tap (visualizer.frag:18)
                c = vis_back(P, v, mk3(ar, ag, ab)*((1.0f/255.0f)/(10.0f*8.0f)));
            }
            if (P.dst_f32)                                              // parity probe: fragColor before the 8-bit store
                reinterpret_cast<float4*>(P.dst_f32)[size_t(y*S + (s / S))*size_t(P.Wr) + size_t(x*S + (s % S))] = make_float4(c.x, c.y, c.z, c.w);
            r8 += (unsigned int)__float2int_rn(__saturatef(c.x)*255.0f);
            g8 += (unsigned int)__float2int_rn(__saturatef(c.y)*255.0f);
            b8 += (unsigned int)__float2int_rn(__saturatef(c.z)*255.0f);
            a8 += (unsigned int)__float2int_rn(__saturatef(c.w)*255.0f);
        }
        const float inv = 1.0f/float(S*S);
        a8 = VP.screen_alpha ? (unsigned int)__float2int_rn(float(a8)*inv) : 255u;
        r8 = (unsigned int)__float2int_rn(float(r8)*inv);
        g8 = (unsigned int)__float2int_rn(float(g8)*inv);
        b8 = (unsigned int)__float2int_rn(float(b8)*inv);
    }
    if (P.comps == 4) {
        if (inside) reinterpret_cast<uchar4*>(P.dst)[size_t(y)*size_t(P.W) + size_t(x)] = make_uchar4(r8, g8, b8, a8);
        return;
    }
    const bool words = (P.W % 4 == 0) && (blockIdx.x*VT_TILE_X + VT_TILE_X <= P.W);
    if (!words) {
        if (inside) { unsigned char* p = P.dst + (size_t(y)*size_t(P.W) + size_t(x))*3; p[0] = r8; p[1] = g8; p[2] = b8; }
        return;
    }
    unsigned char* row = reinterpret_cast<unsigned char*>(stage[threadIdx.y]);
    row[threadIdx.x*3 + 0] = r8; row[threadIdx.x*3 + 1] = g8; row[threadIdx.x*3 + 2] = b8;
    __syncwarp();
    if (threadIdx.x < (VT_TILE_X*3)/4 && y < P.H) {
        unsigned int* out = reinterpret_cast<unsigned int*>(P.dst + (size_t(y)*size_t(P.W) + size_t(blockIdx.x*VT_TILE_X))*3);
        out[threadIdx.x] = stage[threadIdx.y][threadIdx.x];
    }
}

} // namespace glsl

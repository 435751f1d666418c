// visualizer_rows.cu — separable production kernel of the headline scene (examples/basic/shaders/
// visualizer.frag fused with fragment/final.glsl), used when the camera is the axis-aligned 2D one every
// export runs with (camera.py:196-201 defaults: projection 0, right = x, up = y).
//
// Under that camera the texel-space position of a fragment's centre tap is SEPARABLE: x depends only on
// the fragment column, y only on the fragment row (camera.glsl:55-91 + visualizer.frag:16-18 are affine
// per axis). Every one of the 91 blur taps (visualizer.frag:19-33) adds the same offset to all fragments,
// so for one tap
//     all fragments of a column share (ix, fx)     → the horizontal lerp is done once per column and row
//     all fragments of a row    share (iy, fy)     → the vertical weights are warp-uniform table entries
// One thread owns one fragment COLUMN and J consecutive fragment rows (a "row group"). A bilinear tap is
// V(H) — the horizontal lerp H_r = T[r][ix] + fx*(T[r][ix+1]-T[r][ix]) of a few texel rows r (the window
// holds pre-differenced pair records: one FFMA per channel), followed by a vertical interpolation whose
// weights do not depend on the column. V is linear in H, which is what the kernel exploits:
//   phase 1  the 21 taps with dx = 0 (rays 90°, 270° and the undisplaced tap) share ONE horizontal lerp
//            per texel row; their vertical weights are summed per (row group, fragment row, texel row)
//            into a table when the CTA starts: 21 taps cost one pass over <= 4*nq texel rows;
//   phase 2  the 20 taps with dy = 0 (rays 0°, 180°; ray 0 counts twice, it also stands for the 9th
//            direction of the float loop) sum their H first, V is applied once;
//   phase 3  rays 45°/135° and 225°/315° pair up tap by tap (same dy): H is summed per pair, V once.
// V over the 4 texel rows a row group can touch is written with hinge functions,
//   V = H_0 + sum_r (H_{r+1}-H_r)*clamp(py - r0 - r, 0, 1),
// whose three weights per (row group, tap group, fragment row) come from a per-CTA shared table: one
// broadcast LDS.128 + 9 FFMA per fragment; H_0 is common to the J rows and is accumulated once.
// 91 taps cost 60 horizontal gathers + 22 vertical applications per column: ≈ 4.4 k instructions and
// ≈ 1.9 k shared-memory wavefronts per thread (8 fragments), against ≈ 26 k / 7.8 k for the
// one-thread-per-pixel bilinear footprints of visualizer_tiled.cuh. Mathematically identical to the
// literal loop; float32 re-association, and tap offsets that differ by < 1e-7 texel (sin 45° vs sin 135°
// in float32, cos 90° = -4.4e-8) are treated as equal (tests gate it against the literal transliteration).
//
// CTA = 64 fragment columns x 4 row groups (256 threads), J = 8 or 4 rows per group chosen by the host
// from the vertical texel step per fragment (4 texel rows must cover J fragment rows). The background
// window (64 x win_h texels) is staged by TMA bulk copies (cp.async.bulk, one 256-byte row each, mbarrier
// complete_tx) when it lies inside the texture, by wrapped loads at the edges. CTAs whose geometry does
// not fit (window too large, NaN) shade through the per-tap global path, like the tiled kernel.
#include "scenes.cuh"
#include "visualizer_tiled.cuh"
#include "visualizer_rows.h"

#include <math.h>
#include <atomic>

namespace glsl {

constexpr int VR_COLS = 64;            // fragment columns per CTA (threadIdx.x)
constexpr int VR_GROUPS = 4;           // row groups per CTA (threadIdx.y)
constexpr int VR_THREADS = VR_COLS*VR_GROUPS;
constexpr int VR_WIN_MAX = 96;         // widest window row stride (texels) a variant uses: 64 or 96
constexpr int VR_MAX_H = 40;           // window rows the launcher may ask for
constexpr int VR_ROWS = 4;             // texel rows interpolated per tap: covers a vertical span < 2 texels
constexpr int VR_HG = 21;              // vertical (hinge) groups: 0 = rays 0°/180°, 1-10 = rays 45°/135°, 11-20 = rays 225°/315°
constexpr int VR_MAXQ = 6;             // phase 1 walks at most 4*VR_MAXQ texel rows
constexpr int VR_VTAPS = 21;           // taps with dx = 0: rays 90°, 270°, the undisplaced tap
constexpr int VR_NI_MAX = 12;          // phase 2: texel columns the merged horizontal weights of a fragment column may span

struct RowsTaps {
    float hdx[20];                     // phase 2: dx of ray 0 (10 walks), then ray 180°
    float pdx[40];                     // phase 3: pair p = taps 2p, 2p+1 (rays 45°,135° for p < 10; rays 225°,315° after)
    float gdy[VR_HG];                  // dy of each hinge group
    float vdy[VR_VTAPS];               // phase 1: dy of rays 90°, 270°, then 0
};
__constant__ RowsTaps c_rows;

// Per-frame constants of visualizer.frag (functions of the uniforms only), once per frame. Two routes, chosen per
// kernel variant by measurement: the 96-texel-window variants (the 1080p default export: 150 us per frame) take them
// from the kernel parameter block, filled by the launcher on the host — a 1-thread kernel in front of every frame
// costs ~3 us, 2 % of such a frame; the 64-texel-window variants (4K 2xSSAA: 1.05 ms) read them from global memory
// where that 1-thread kernel put them — as parameter-block operands they change ptxas' allocation of the blur loop
// (56 / 112 B of spills instead of 16 / 32 B) and the frame takes 2 % longer, seven times what the launch costs.
struct FrameConsts {
    float zf, wobx, woby, scale;       // background zoom factor, wobble, blur radius in texels (:16-17,21)
    float std5, mscale, vexp, pad;     // 5*iAudioSTD, 1 - 0.4*pow(|vol|, 0.5), 0.1 + 0.15*vol (:35,39,71)
};
constexpr int VR_SLOTS = 64;
__device__ FrameConsts g_frame[VR_SLOTS];
SFB_HD FrameConsts frame_consts(float iTime, float volume, float stddev, float fh) {
    FrameConsts F;
    F.zf = 0.95f + 0.01f*sinf(iTime) - 0.02f*volume - 0.03f;
    F.wobx = 0.005f*cosf(iTime*3.25135f); F.woby = 0.005f*sinf(iTime*1.153469f);
    F.scale = (0.01f*fminf(fmaxf(powf(volume, 2.5f), 0.0f), 0.3f))*fh;     // st displacement → texels (hw*fw == fh)
    F.std5 = 5.0f*stddev;
    F.mscale = 1.0f - 0.4f*powf(fabsf(volume), 0.5f);
    F.vexp = 0.1f + 0.15f*volume;
    F.pad = 0.0f;
    return F;
}
__global__ void visualizer_frame_consts_kernel(float iTime, float volume, float stddev, float fh, int slot) {
    if (threadIdx.x == 0) g_frame[slot] = frame_consts(iTime, volume, stddev, fh);
}

struct VisRowsParams {
    RenderParams R;
    int win_h;                         // window rows staged (dynamic shared memory is sized for it)
    int debug;                         // SFB_ROWS_DEBUG bits (profiling only): 1 skip the taps, 2 skip the back end
    int screen_alpha;                  // comps == 4: store fragColor.a (iScreen pass of an unfused export) instead of 255
    const void* tmap;                  // device CUtensorMap of the background with a WW x win_h box, or NULL (bulk row copies)
    int slot;                          // WW == 64: g_frame entry written by visualizer_frame_consts_kernel for this frame
    FrameConsts K;                     // WW == 96: frame_consts() of this frame's uniforms, evaluated by the launcher
};

SFB_DEV void bulk_load_row(void* dst, const void* src, unsigned int bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// MUFU-direct square root / log2 / exp2 (PTX .approx: relative error <= 2^-22, against <= 2 ulp for the
// libdevice functions the literal transliteration calls — 3 to 6 times fewer instructions, no slow-path
// branches). The back end is evaluated once per fragment and feeds an 8-bit store: 1e-6 is invisible.
SFB_DEV float fast_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
SFB_DEV float fast_lg2(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
SFB_DEV float fast_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// pow(x, y) for x >= 0 and the small |y| used here (<= 0.5, which shrinks the error of lg2); pow(0, y > 0) = 0
SFB_DEV float pow_pos(float x, float y) { return fast_ex2(y*fast_lg2(x)); }

// What the back end reads of iSpectrogram when it is the usual RG32F column (spectrogram.py:272-282)
struct SpecColumn { const float2* texels; int h, ry; };

// Everything of main() after the blur loop (visualizer.frag:35-73) for the separable camera: the column
// gives (agluv.x, astuv.x, uv.x, waveform thresholds), the row gives (agluv.y, astuv.y, uv.y). Same
// mathematics and operation order as vis_back (visualizer_tiled.cuh), written branch-free so that two
// fragments interleave: pow with the constant exponent 6 is three multiplies, the fractional powers go
// through pow_pos, divisions by constants are multiplications, square roots are MUFU-direct.
SFB_DEV vec3 vis_back_sep(const RenderParams& P, const FrameConsts& K, const SpecColumn& sc, vec3 rgb, float agx, float asx,
                          float uvx, float wavx, float wavy, float agy, float asy, float uvy) {
    const vec3 space = mk3(1.0f, 11.0f, 26.0f)/255.0f;
    const float t = clamp(fast_sqrt(agx*agx + agy*agy) - 0.3f, 0.0f, 1.0f), t2 = t*t;
    rgb = rgb*(1.0f + K.std5*(t2*t2*t2));
    const float c = -4.37113883e-08f, sn = -1.0f;                 // cos, sin of float32(-PI/2)
    const float mx = (c*uvx + sn*uvy)*K.mscale, my = ((-sn)*uvx + c*uvy)*K.mscale;
    const float radius = 0.17f;
    const float circle = fabsf(atan2f(my, mx)*(1.0f/PI));
    float2 sv;                                                    // texture(iSpectrogram, (0, circle)), NEAREST
    if (sc.texels) {
        sv = __ldg(sc.texels + wrap_index(int(floorf(circle*float(sc.h))), sc.h, sc.ry));
    } else {
        const vec4 q = texture<false>(P.tex[1], mk2(0.0f, circle)); sv = make_float2(q.x, q.y);
    }
    const float h = clamp(circle*0.5f, 0.0f, 1.0f);
    const float fscale = 0.05f + 3.0f*(h*h*(3.0f - 2.0f*h));
    const float lm = fast_sqrt(mx*mx + my*my);
    const float bar = fast_sqrt(((my < 0.0f) ? sv.x : sv.y)*0.001f)*fscale;
    const float r = radius + 0.5f*bar;
    // inside the disc: rgb/2; on the bar: mix(rgb, 1, smoothstep(0.5 + bar)); outside: rgb*pow(.., 0.05)
    const float g = clamp(0.5f + bar, 0.0f, 1.0f), sm = g*g*(3.0f - 2.0f*g);
    const float dim = pow_pos((lm - r)*0.5f, 0.05f);
    const bool inside = lm < radius, ring = lm < r;
    const float mul = inside ? 0.5f : (ring ? (1.0f - sm) : dim), add = (!inside && ring) ? sm : 0.0f;
    rgb = rgb*mul + add;
    { const float e = clamp(fast_sqrt(uvx*uvx + uvy*uvy)*0.05f, 0.0f, 1.0f); rgb = mix(rgb, space, e*e*(3.0f - 2.0f*e)); }
    rgb = rgb*pow_pos((asx*(1.0f - asy))*(asy*(1.0f - asx))*20.0f, K.vexp);
    if (1.0f - agy < wavx) rgb = rgb*0.8f;
    if (1.0f + agy < wavy) rgb = rgb*0.8f;
    return rgb;
}

// The per-tap global path for CTAs whose geometry does not fit; kept out of line (it is cold and large)
__device__ __noinline__ void visualizer_unfitted(const RenderParams& P, int i, int j, float* rgb) {
    const vec4 c = scene_visualizer_fast(P, make_frag(P, i, j));
    rgb[0] = c.x; rgb[1] = c.y; rgb[2] = c.z;
}

#ifndef VR_MIN_CTAS
#define VR_MIN_CTAS 3
#endif
typedef unsigned long long vr_u64;
// two float32 lanes in one 64-bit register pair: fma.rn.f32x2 does two FMAs in one issue slot (the kernel is issue-limited)
SFB_DEV vr_u64 pack2(float lo, float hi) { vr_u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
SFB_DEV void unpack2(vr_u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
SFB_DEV vr_u64 fma2(vr_u64 a, vr_u64 b, vr_u64 c) { vr_u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
constexpr int VR_TEXEL_BYTES = 24;                             // window bytes per texel over both planes
constexpr int vr_max_h(int ww) { return ww == 64 ? VR_MAX_H : 32; }     // window rows a variant may stage

// S = ssaa (sub-samples per output pixel and axis), J = fragment rows per thread, WW = window row stride in texels
// (64 when the horizontal texel step per fragment is small, 96 up to ~1 texel per fragment: the reference's default
// export, ssaa = 1 at the background's own resolution)
template <int S, int J, int WW>
__global__ void __launch_bounds__(VR_THREADS, VR_MIN_CTAS)
visualizer_rows_kernel(const __grid_constant__ VisRowsParams VP) {
    static_assert(J % S == 0 && VR_COLS % S == 0, "a CTA shades whole output pixels");
    constexpr int VR_WIN_W = WW;
    const RenderParams& P = VP.R;
    const int win_h = VP.win_h;
    extern __shared__ __align__(128) unsigned char vr_smem[];
    float4* rg = reinterpret_cast<float4*>(vr_smem);                                     // [win_h][64] (r, g, r'-r, g'-g)
    float2* bb = reinterpret_cast<float2*>(vr_smem + sizeof(float4)*VR_WIN_W*win_h);     // [win_h][64] (b, b'-b)
    float4* tblH = reinterpret_cast<float4*>(vr_smem + (sizeof(float4) + sizeof(float2))*VR_WIN_W*win_h);  // [4][VR_HG][J] hinge weights + row offset
    float4* tblM = tblH + VR_GROUPS*VR_HG*J;                                             // [4][VR_MAXQ][J] merged weights of 4 texel rows
    float* wx = reinterpret_cast<float*>(tblM + VR_GROUPS*VR_MAXQ*J);                    // [VR_NI_MAX][64] phase 2: merged horizontal weights
    // even J: the vertical weights are laid out per PAIR of fragment rows, (w_j, w_j+1) adjacent, so that one f32x2 FMA
    // applies a texel-row difference to two fragment rows (same region, different arrangement):
    constexpr bool PAIRS = (J % 2) == 0;
    float4* tblQ = tblH;                                                                 // [4][VR_HG][J/2] (h0_j, h0_j+1, h1_j, h1_j+1)
    float2* tblW = reinterpret_cast<float2*>(tblQ + VR_GROUPS*VR_HG*(J/2));              // [4][VR_HG][J/2] (h2_j, h2_j+1)
    __shared__ unsigned int rowoffS[VR_GROUPS][VR_HG];                                   // window row offset of each (row group, hinge group)
    __shared__ float red[2][VR_THREADS/32];
    __shared__ float cyS[VR_GROUPS][J];
    __shared__ float4 rowS[VR_GROUPS][J];                                                // (agluv.y, astuv.y, uv.y, -) per fragment row
    __shared__ int win[5];                                                               // x0, y0, fits, tma, table overflow
    __shared__ unsigned int mhdr[VR_GROUPS][2];                                          // phase 1: first row offset, row quads
    __shared__ __align__(8) unsigned long long bar;

    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty*VR_COLS + tx;
    const int i = blockIdx.x*VR_COLS + tx;                        // fragment column
    const int jb = blockIdx.y*(VR_GROUPS*J) + ty*J;               // first fragment row of this thread
    const DevSampler& bg = P.tex[0];
    const float fw = float(bg.w), fh = float(bg.h), hw = float(bg.h)/float(bg.w);
    const FrameConsts K = (WW == 64) ? g_frame[VP.slot] : VP.K;
    const float zf = K.zf, scale = K.scale;
    const vec2 wobble = mk2(K.wobx, K.woby);

    // ---- A. centre-tap positions: x per column, y per row (coordinates clamped onto the target) ----
    const int ic = min(i, P.Wr - 1), jc = min(jb, P.Hr - 1);
    const VisFrag vc = vis_front(P, ic, jc, fw, fh, hw, wobble, zf);
    const float tapx = vc.tap.x;
    if (tx < J) {
        const VisFrag vr = vis_front(P, min(int(blockIdx.x)*VR_COLS, P.Wr - 1), min(jb + tx, P.Hr - 1), fw, fh, hw, wobble, zf);
        cyS[ty][tx] = vr.tap.y;
        rowS[ty][tx] = make_float4(vr.f.agluv.y, vr.f.astuv.y, vr.uv.y, 0.0f);
    }
    {
        float lo = tapx, hi = tapx;
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if ((tid & 31) == 0) { red[0][tid >> 5] = lo; red[1][tid >> 5] = hi; }
    }
    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (tid == 0) {
        float ax = red[0][0], bx = red[1][0], ay = cyS[0][0], by = cyS[0][0];
        for (int w = 1; w < VR_THREADS/32; w++) { ax = fminf(ax, red[0][w]); bx = fmaxf(bx, red[1][w]); }
        {   // rows are affine in j: the extremes are the first and the last row of the tile
            const float first = cyS[0][0], last = cyS[VR_GROUPS - 1][J - 1];
            ay = fminf(first, last); by = fmaxf(first, last);
        }
        const float reach = scale*1.0001f + 1.0f;                 // |dir*walk| <= 1.0000001; +1 keeps local coords >= 1
        // bulk copies need 16-byte aligned rows: x0 % 4 == 0
        const int x0 = (int(floorf(ax - reach)) - 1) & ~3, y0 = int(floorf(ay - reach)) - 1;
        const int x1 = int(floorf(bx + reach)) + 2, y1 = int(floorf(by + reach)) + 2;    // last texel touched + 1
        const bool finite = (ax == ax) && (bx == bx) && (ay == ay) && (by == by) && fabsf(ax) < 1.0e9f && fabsf(bx) < 1.0e9f
                         && fabsf(ay) < 1.0e9f && fabsf(by) < 1.0e9f;
        win[0] = x0; win[1] = y0; win[4] = 0;
        // pair records need texel x+1: the last window column only ever serves as a neighbour
        win[2] = (finite && (x1 - x0) <= VR_WIN_W - 1 && (y1 - y0) <= win_h) ? 1 : 0;
        win[3] = (win[2] && (bg.w % 4) == 0 && x0 >= 0 && y0 >= 0 && x0 + VR_WIN_W <= bg.w && y0 + win_h <= bg.h) ? 1 : 0;
        if (win[3]) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&bar, VR_WIN_W*4*win_h);
            // ---- B1. TMA: the raw RGBA8 window lands at the start of the window buffer (WW*4 bytes per row) —
            //          ONE tensor load (cp.async.bulk.tensor.2d, box WW x win_h) when the launcher made a map
            if (VP.tmap) tma_load_2d(vr_smem, VP.tmap, &bar, x0, y0);
        }
    }
    __syncthreads();
    const int x0 = win[0], y0 = win[1];
    const bool window_ok = win[2] != 0, tma = win[3] != 0;

    // ---- B1'. no tensor map (odd texture widths, SFB_NO_TMA): one bulk copy per window row -----------
    if (tma && !VP.tmap && tid < win_h) {
        const unsigned int* src = reinterpret_cast<const unsigned int*>(bg.lin) + size_t(y0 + tid)*size_t(bg.w) + size_t(x0);
        bulk_load_row(vr_smem + tid*(VR_WIN_W*4), src, VR_WIN_W*4, &bar);
    }

    // ---- C. vertical tables (warp-uniform in the main loop) ------------------------------------------
    if (window_ok) {
        const float y0f = float(y0);
        int bad = 0;
        // hinge weights of every (row group, tap group, fragment row)
        for (int e = tid; e < VR_GROUPS*VR_HG*J; e += VR_THREADS) {
            const int g = e/(VR_HG*J), rem = e - g*(VR_HG*J), k = rem/J, r = rem - k*J;
            const float dy = c_rows.gdy[k];
            const float p0 = fmaf(dy, scale, cyS[g][0] - y0f), pl = fmaf(dy, scale, cyS[g][J - 1] - y0f);
            const float py = fmaf(dy, scale, cyS[g][r] - y0f);
            const float r0f = floorf(fminf(p0, pl));
            const float t = py - r0f;
            const int r0 = int(r0f);
            if (!(t >= 0.0f && t <= float(VR_ROWS - 1)) || r0 < 0 || r0 + VR_ROWS > win_h) bad = 1;
            // byte offset of window row r0 in the (b, b'-b) plane, with the 2^23 exponent bits of the magic
            // floor of x folded in (modulo 2^32); the (r, g) plane uses twice this offset
            const unsigned int rowoff = (unsigned int)r0*(VR_WIN_W*8u) - (0x4B000000u << 3);
            if constexpr (PAIRS) {
                const int slot = (g*VR_HG + k)*(J/2) + (r >> 1), odd = r & 1;
                reinterpret_cast<float*>(tblQ + slot)[odd] = __saturatef(t);
                reinterpret_cast<float*>(tblQ + slot)[2 + odd] = __saturatef(t - 1.0f);
                reinterpret_cast<float*>(tblW + slot)[odd] = __saturatef(t - 2.0f);
                if (r == 0) rowoffS[g][k] = rowoff;
            } else {
                tblH[e] = make_float4(__saturatef(t), __saturatef(t - 1.0f), __saturatef(t - 2.0f), __uint_as_float(rowoff));
            }
        }
        // merged weights of the dx = 0 taps: W[g][row][r] = sum over the 21 taps of hat(py - row), the
        // bilinear weight of texel row `row` (hat(t) = max(0, 1 - |t|)); one entry per loop trip
        for (int e = tid; e < VR_GROUPS*VR_MAXQ*4*J; e += VR_THREADS) {
            const int g = e/(VR_MAXQ*4*J), rem = e - g*(VR_MAXQ*4*J), row = rem/J, r = rem - row*J;
            const float lo = fminf(cyS[g][0], cyS[g][J - 1]) - y0f, hi = fmaxf(cyS[g][0], cyS[g][J - 1]) - y0f;
            const int ry0 = int(floorf(lo - scale*1.0001f));
            const int ry1 = int(floorf(hi + scale*1.0001f)) + 1;                 // last texel row touched
            const int nq = (ry1 - ry0 + 4) >> 2;
            if (ry0 < 0 || nq < 1 || nq > VR_MAXQ || ry0 + 4*nq > win_h) { bad = 1; continue; }
            if (row >= 4*nq) continue;                                           // phase 1 never reads these
            const float cy = cyS[g][r] - y0f - float(ry0 + row);
            float w = 0.0f;
            #pragma unroll 7
            for (int k = 0; k < VR_VTAPS; k++) w += fmaxf(1.0f - fabsf(fmaf(c_rows.vdy[k], scale, cy)), 0.0f);
            if constexpr (PAIRS) {   // texel rows (4q, 4q+1) of fragment rows (j, j+1) in one float4, rows (4q+2, 4q+3) in the next
                const int c = row & 3;
                float* quad = reinterpret_cast<float*>(tblM + (((g*VR_MAXQ + (row >> 2))*(J/2) + (r >> 1))*2 + (c >> 1)));
                quad[(c & 1)*2 + (r & 1)] = w;
            } else {
                reinterpret_cast<float*>(tblM + (g*VR_MAXQ + (row >> 2))*J + r)[row & 3] = w;
            }
            if (rem == 0) { mhdr[g][0] = (unsigned int)ry0*(VR_WIN_W*8u) - (0x4B000000u << 3); mhdr[g][1] = (unsigned int)nq; }
        }
        if (bad) win[4] = 1;                                       // benign race: every writer stores 1
    }
    // merged horizontal weights of the dy = 0 taps (phase 2): the 20 taps of rays 0° (counted twice) and 180° read the
    // same 4 texel rows, so their horizontal lerps add up to ONE weighted sum over the ni texel columns they can touch,
    // wx[k][column] = sum over the taps of weight * hat(px_tap - texel k). The weights depend on the fragment column
    // only: each column's 4 threads (one per row group) compute them once per CTA.
    const int ni = int(2.0f*scale*1.0001f) + 3;                    // warp-uniform: texel columns floor(cx - s) .. floor(cx + s) + 1
    const bool merged_x = window_ok && ni <= VR_NI_MAX && !(VP.debug & 4);
    const float cxl = tapx - float(x0);                            // tile-local column position, >= 2 by construction of x0
    const int ix0 = int(floorf(cxl - scale*1.0001f));
    if (merged_x) {
        for (int k = ty; k < ni; k += VR_GROUPS) {
            const float at = float(ix0 + k);
            float w = 0.0f;
            #pragma unroll 5
            for (int t = 0; t < 10; t++) {
                w += 2.0f*fmaxf(1.0f - fabsf(fmaf(c_rows.hdx[t], scale, cxl) - at), 0.0f);
                w += fmaxf(1.0f - fabsf(fmaf(c_rows.hdx[10 + t], scale, cxl) - at), 0.0f);
            }
            wx[k*VR_COLS + tx] = w;
        }
    }

    // ---- B2. widen the window into pre-differenced pair records -------------------------------------
    if (window_ok) {
        constexpr int PER = (VR_WIN_W*vr_max_h(WW) + VR_THREADS - 1)/VR_THREADS;
        unsigned int w0[PER], w1[PER];
        const int n = VR_WIN_W*win_h;
        if (tma) {
            mbar_wait(&bar, 0);
            const unsigned int* raw = reinterpret_cast<const unsigned int*>(vr_smem);
            #pragma unroll
            for (int k = 0; k < PER; k++) {
                const int t = tid + k*VR_THREADS;
                if (t < n) { w0[k] = raw[t]; w1[k] = raw[((t % VR_WIN_W) == VR_WIN_W - 1) ? t : t + 1]; }
            }
        } else {
            const unsigned int* texels = reinterpret_cast<const unsigned int*>(bg.lin);
            #pragma unroll
            for (int k = 0; k < PER; k++) {
                const int t = tid + k*VR_THREADS;
                if (t < n) {
                    const int gy = wrap_index(y0 + (t / VR_WIN_W), bg.h, bg.ry);
                    const unsigned int* row = texels + size_t(gy)*size_t(bg.w);
                    w0[k] = __ldg(row + wrap_index(x0 + (t % VR_WIN_W), bg.w, bg.rx));
                    w1[k] = __ldg(row + wrap_index(x0 + (t % VR_WIN_W) + 1, bg.w, bg.rx));
                }
            }
        }
        __syncthreads();                                           // raw rows fully read before they are overwritten
        #pragma unroll
        for (int k = 0; k < PER; k++) {
            const int t = tid + k*VR_THREADS;
            if (t < n) {
                const vec3 a = widen3(w0[k]), b = widen3(w1[k]);
                rg[t] = make_float4(a.x, a.y, b.x - a.x, b.y - a.y);
                bb[t] = make_float2(a.z, b.z - a.z);
            }
        }
    }
    __syncthreads();
    const bool fits = window_ok && win[4] == 0;

    // ---- D. blur ------------------------------------------------------------------------------------
    float base0 = 0.0f, base1 = 0.0f, base2 = 0.0f;
    float acc[J][3];
    #pragma unroll
    for (int r = 0; r < J; r++) { acc[r][0] = 0.0f; acc[r][1] = 0.0f; acc[r][2] = 0.0f; }
    if (fits && !(VP.debug & 1)) {
        vr_u64 acc2[PAIRS ? J/2 : 1][3];                           // even J: channel c of fragment rows (2jp, 2jp + 1)
        #pragma unroll
        for (int jp = 0; jp < (PAIRS ? J/2 : 1); jp++) { acc2[jp][0] = 0ull; acc2[jp][1] = 0ull; acc2[jp][2] = 0ull; }
        const float cx = tapx - float(x0);                         // tile-local, >= 1 by construction of x0
        const char* rgB = reinterpret_cast<const char*>(rg);
        const char* bbB = reinterpret_cast<const char*>(bb);
        // Horizontal lerp of the 4 texel rows starting at byte offset `rowoff` (b plane units), at column
        // position cx + dx*scale: H[r] (+)= T + a*(T' - T)
        auto gather = [&](float dx, unsigned int rowoff, float (&H)[VR_ROWS][3], const bool first) {
            const float px = fmaf(dx, scale, cx);
            // floor for 0.5 <= p < 2^22: p + (2^23 - 0.5) rounds to floor(p) + 2^23 (ties land on either
            // neighbour; the interpolant is continuous there)
            const float tx_ = px + 8388607.5f;
            const float a = px - (tx_ - 8388608.0f);
            const unsigned int off8 = rowoff + (__float_as_uint(tx_) << 3);
            const char* pr = rgB + 2u*off8;
            const char* pb = bbB + off8;
            #pragma unroll
            for (int r = 0; r < VR_ROWS; r++) {
                const float4 q = *reinterpret_cast<const float4*>(pr + r*(VR_WIN_W*16));
                const float2 s = *reinterpret_cast<const float2*>(pb + r*(VR_WIN_W*8));
                if (first) { H[r][0] = fmaf(a, q.z, q.x); H[r][1] = fmaf(a, q.w, q.y); H[r][2] = fmaf(a, s.y, s.x); }
                else { H[r][0] = fmaf(a, q.z, H[r][0] + q.x); H[r][1] = fmaf(a, q.w, H[r][1] + q.y); H[r][2] = fmaf(a, s.y, H[r][2] + s.x); }
            }
        };
        // Vertical interpolation of H for the J fragment rows with the hinge weights T[0..J)
        auto apply = [&](const float4* T, const float4 e0, const float (&H)[VR_ROWS][3]) {
            base0 += H[0][0]; base1 += H[0][1]; base2 += H[0][2];
            float D[VR_ROWS - 1][3];
            #pragma unroll
            for (int r = 0; r < VR_ROWS - 1; r++) {
                D[r][0] = H[r + 1][0] - H[r][0]; D[r][1] = H[r + 1][1] - H[r][1]; D[r][2] = H[r + 1][2] - H[r][2];
            }
            #pragma unroll
            for (int r = 0; r < J; r++) {
                const float4 e = (r == 0) ? e0 : T[r];
                #pragma unroll
                for (int c = 0; c < 3; c++)
                    acc[r][c] = fmaf(e.x, D[0][c], fmaf(e.y, D[1][c], fmaf(e.z, D[2][c], acc[r][c])));
            }
        };

        // the same for even J, two fragment rows per f32x2 FMA: the differences are duplicated into both lanes once per
        // application (9 moves for J rows), the weights of a row pair are adjacent in the table (LDS.128 + LDS.64)
        auto apply2 = [&](int k, const float (&H)[VR_ROWS][3]) {
            base0 += H[0][0]; base1 += H[0][1]; base2 += H[0][2];
            vr_u64 D2[VR_ROWS - 1][3];
            #pragma unroll
            for (int r = 0; r < VR_ROWS - 1; r++) {
                #pragma unroll
                for (int c = 0; c < 3; c++) { const float d = H[r + 1][c] - H[r][c]; D2[r][c] = pack2(d, d); }
            }
            const ulonglong2* Q = reinterpret_cast<const ulonglong2*>(tblQ + (ty*VR_HG + k)*(J/2));
            const vr_u64* Wt = reinterpret_cast<const vr_u64*>(tblW + (ty*VR_HG + k)*(J/2));
            #pragma unroll
            for (int jp = 0; jp < J/2; jp++) {
                const ulonglong2 w = Q[jp];
                const vr_u64 w2 = Wt[jp];
                #pragma unroll
                for (int c = 0; c < 3; c++)
                    acc2[jp][c] = fma2(w.x, D2[0][c], fma2(w.y, D2[1][c], fma2(w2, D2[2][c], acc2[jp][c])));
            }
        };

        // phase 1: the dx = 0 taps through the merged weights, 4 texel rows at a time
        if constexpr (PAIRS) {
            const ulonglong2* M = reinterpret_cast<const ulonglong2*>(tblM + ty*(VR_MAXQ*J));
            unsigned int rowoff = mhdr[ty][0];
            const int nq = int(mhdr[ty][1]);
            #pragma unroll 1
            for (int q = 0; q < nq; q++, rowoff += 4u*(VR_WIN_W*8u)) {
                float H[VR_ROWS][3];
                gather(0.0f, rowoff, H, true);
                vr_u64 H2[VR_ROWS][3];
                #pragma unroll
                for (int r = 0; r < VR_ROWS; r++) {
                    #pragma unroll
                    for (int c = 0; c < 3; c++) H2[r][c] = pack2(H[r][c], H[r][c]);
                }
                #pragma unroll
                for (int jp = 0; jp < J/2; jp++) {
                    const ulonglong2 wa = M[(q*(J/2) + jp)*2], wb = M[(q*(J/2) + jp)*2 + 1];
                    #pragma unroll
                    for (int c = 0; c < 3; c++)
                        acc2[jp][c] = fma2(wa.x, H2[0][c], fma2(wa.y, H2[1][c], fma2(wb.x, H2[2][c], fma2(wb.y, H2[3][c], acc2[jp][c]))));
                }
            }
        } else {
            const float4* M = tblM + ty*(VR_MAXQ*J);
            unsigned int rowoff = mhdr[ty][0];
            const int nq = int(mhdr[ty][1]);
            #pragma unroll 1
            for (int q = 0; q < nq; q++, rowoff += 4u*(VR_WIN_W*8u)) {
                float H[VR_ROWS][3];
                gather(0.0f, rowoff, H, true);
                #pragma unroll
                for (int r = 0; r < J; r++) {
                    const float4 w = M[q*J + r];
                    #pragma unroll
                    for (int c = 0; c < 3; c++)
                        acc[r][c] = fmaf(w.x, H[0][c], fmaf(w.y, H[1][c], fmaf(w.z, H[2][c], fmaf(w.w, H[3][c], acc[r][c]))));
                }
            }
        }
        const float4* T = tblH + ty*(VR_HG*J);
        // phase 2: the dy = 0 taps; ray 0 is weighted twice
        if (merged_x) {
            // one pass over the ni texel columns with the merged weights: per column and texel row one 8-byte and one
            // 4-byte load (r, g, b of the record; the differences are not needed) instead of 20 two-texel gathers
            float4 e0 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            unsigned int rowoff0;
            if constexpr (PAIRS) rowoff0 = rowoffS[ty][0]; else { e0 = T[0]; rowoff0 = __float_as_uint(e0.w); }
            const unsigned int base8 = rowoff0 + (0x4B000000u << 3) + ((unsigned int)ix0 << 3);
            const char* pr = rgB + 2u*base8;
            const char* pb = bbB + base8;
            const float* wk = wx + tx;
            float H[VR_ROWS][3];
            #pragma unroll
            for (int r = 0; r < VR_ROWS; r++) { H[r][0] = 0.0f; H[r][1] = 0.0f; H[r][2] = 0.0f; }
            #pragma unroll 3
            for (int k = 0; k < ni; k++, pr += 16, pb += 8, wk += VR_COLS) {
                const float w = *wk;
                #pragma unroll
                for (int r = 0; r < VR_ROWS; r++) {
                    const float2 q = *reinterpret_cast<const float2*>(pr + r*(VR_WIN_W*16));
                    const float b = *reinterpret_cast<const float*>(pb + r*(VR_WIN_W*8));
                    H[r][0] = fmaf(w, q.x, H[r][0]); H[r][1] = fmaf(w, q.y, H[r][1]); H[r][2] = fmaf(w, b, H[r][2]);
                }
            }
            if constexpr (PAIRS) apply2(0, H); else apply(T, e0, H);
        } else {
            float4 e0 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            unsigned int rowoff;
            if constexpr (PAIRS) rowoff = rowoffS[ty][0]; else { e0 = T[0]; rowoff = __float_as_uint(e0.w); }
            float H[VR_ROWS][3];
            gather(c_rows.hdx[0], rowoff, H, true);
            #pragma unroll 3
            for (int t = 1; t < 10; t++) gather(c_rows.hdx[t], rowoff, H, false);
            #pragma unroll
            for (int r = 0; r < VR_ROWS; r++) { H[r][0] += H[r][0]; H[r][1] += H[r][1]; H[r][2] += H[r][2]; }
            #pragma unroll 2
            for (int t = 10; t < 20; t++) gather(c_rows.hdx[t], rowoff, H, false);
            if constexpr (PAIRS) apply2(0, H); else apply(T, e0, H);
        }
        // phase 3: the diagonal rays, two taps per vertical group
        #pragma unroll 2
        for (int p = 0; p < 20; p++) {
            const float4* Tp = T + (p + 1)*J;
            float4 e0 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            unsigned int rowoff;
            if constexpr (PAIRS) rowoff = rowoffS[ty][p + 1]; else { e0 = Tp[0]; rowoff = __float_as_uint(e0.w); }
            float H[VR_ROWS][3];
            gather(c_rows.pdx[2*p], rowoff, H, true);
            gather(c_rows.pdx[2*p + 1], rowoff, H, false);
            if constexpr (PAIRS) apply2(p + 1, H); else apply(Tp, e0, H);
        }
        if constexpr (PAIRS) {
            #pragma unroll
            for (int jp = 0; jp < J/2; jp++) {
                #pragma unroll
                for (int c = 0; c < 3; c++) unpack2(acc2[jp][c], acc[2*jp][c], acc[2*jp + 1][c]);
            }
        }
    }

    // ---- E. rest of main() per fragment, 8-bit store rule per sub-sample, box sum --------------------
    // The blurred colours go through shared memory so that the back end is ONE rolled loop (unrolled J
    // times it overflows the instruction cache: ncu showed 2.3 no-instruction stalls per issue).
    __syncthreads();                                               // table and window are dead: reuse them
    float4* stash = reinterpret_cast<float4*>(vr_smem);            // [J][256] blurred rgb, each thread reads its own
    unsigned char* stage = vr_smem + sizeof(float4)*J*VR_THREADS;  // rgb24 rows of the CTA
    {
        const float norm = (1.0f/255.0f)/(10.0f*8.0f);
        #pragma unroll
        for (int r = 0; r < J; r++)
            stash[r*VR_THREADS + tid] = make_float4((base0 + acc[r][0])*norm, (base1 + acc[r][1])*norm, (base2 + acc[r][2])*norm, 0.0f);
    }
    constexpr int PX = VR_COLS/S;                                  // output pixels per CTA row
    constexpr int RB = PX*3;                                       // staged bytes per output row
    const bool col_in = i < P.Wr;
    const bool words = (P.comps == 3) && (P.W % 4 == 0) && (int(blockIdx.x)*PX + PX <= P.W);
    SpecColumn sc;
    {
        const DevSampler& sp = P.tex[1];
        const bool column = sp.dtype == SFB_DTYPE_F32 && sp.padded == 2 && sp.w == 1 && sp.filter == SFB_FILTER_NEAREST;
        sc.texels = column ? reinterpret_cast<const float2*>(sp.lin) : nullptr; sc.h = sp.h; sc.ry = sp.ry;
    }
    const float agx = vc.f.agluv.x, asx = vc.f.astuv.x, uvx = vc.uv.x;
    float wavx, wavy;                                              // 0.2*texture(iWaveform, (astuv.x, 0)): per column
    {
        const DevSampler& wv = P.tex[2];
        if (wv.dtype == SFB_DTYPE_F32 && wv.padded == 2 && wv.h == 1 && wv.filter == SFB_FILTER_LINEAR && !wv.rx) {
            // the usual RG32F row, LINEAR + CLAMP_TO_EDGE (waveform.py:64-87): same arithmetic as texture<false>
            const float ub = asx*float(wv.w) - 0.5f, fx = floorf(ub), a = ub - fx;
            const int i0 = min(max(int(fx), 0), wv.w - 1), i1 = min(max(int(fx) + 1, 0), wv.w - 1);
            const float2 t0 = __ldg(reinterpret_cast<const float2*>(wv.lin) + i0), t1 = __ldg(reinterpret_cast<const float2*>(wv.lin) + i1);
            // rows j0 and j0 + 1 are the same texel row (h == 1): top*(1 - b) + bot*b with top == bot
            const float b = (0.0f*1.0f - 0.5f) - floorf(0.0f*1.0f - 0.5f);
            const float tx_ = t0.x*(1.0f - a) + t1.x*a, ty_ = t0.y*(1.0f - a) + t1.y*a;
            wavx = 0.2f*(tx_*(1.0f - b) + tx_*b); wavy = 0.2f*(ty_*(1.0f - b) + ty_*b);
        } else {
            const vec4 w = texture<false>(wv, mk2(asx, 0.0f)); wavx = 0.2f*w.x; wavy = 0.2f*w.y;
        }
    }
    // two fragments in flight per iteration (the back end is a long dependent chain)
    #pragma unroll (S == 1 ? 2 : 1)
    for (int pr = 0; pr < J/S; pr++) {
        unsigned int r8 = 0, g8 = 0, b8 = 0;
        const unsigned int a8 = (VP.screen_alpha && vc.oob) ? 0u : 255u;      // visualizer.frag:11-14 leaves alpha unwritten
        #pragma unroll (S >= 2 ? 2 : 1)
        for (int s = 0; s < S; s++) {
            const int r = pr*S + s;
            vec3 c;
            if (vc.oob) {
                c = mk3(1.0f, 11.0f, 26.0f)/255.0f;
            } else if (!fits) {                                    // geometry did not fit: per-tap global path
                float q[3]; visualizer_unfitted(P, ic, min(jb + r, P.Hr - 1), q); c = mk3(q[0], q[1], q[2]);
            } else {
                const float4 q = stash[r*VR_THREADS + tid];
                const float4 row = rowS[ty][r];
                c = (VP.debug & 2) ? mk3(q.x, q.y, q.z)
                                   : vis_back_sep(P, K, sc, mk3(q.x, q.y, q.z), agx, asx, uvx, wavx, wavy, row.x, row.y, row.z);
            }
            if (P.dst_f32 && col_in && jb + r < P.Hr)              // parity probe: fragColor before the 8-bit store
                reinterpret_cast<float4*>(P.dst_f32)[size_t(jb + r)*size_t(P.Wr) + size_t(i)] = make_float4(c.x, c.y, c.z, vc.oob ? 0.0f : 1.0f);
            r8 += (unsigned int)__float2int_rn(__saturatef(c.x)*255.0f);
            g8 += (unsigned int)__float2int_rn(__saturatef(c.y)*255.0f);
            b8 += (unsigned int)__float2int_rn(__saturatef(c.z)*255.0f);
        }
        #pragma unroll
        for (int o = 1; o < S; o <<= 1) {
            r8 += __shfl_xor_sync(0xffffffffu, r8, o); g8 += __shfl_xor_sync(0xffffffffu, g8, o); b8 += __shfl_xor_sync(0xffffffffu, b8, o);
        }
        const float inv = 1.0f/float(S*S);
        r8 = (unsigned int)__float2int_rn(float(r8)*inv);
        g8 = (unsigned int)__float2int_rn(float(g8)*inv);
        b8 = (unsigned int)__float2int_rn(float(b8)*inv);
        const int y = (jb + pr*S)/S, x = i/S;                      // output pixel
        if ((tx % S) == 0 && col_in && y < P.H) {
            if (P.comps == 4) {
                reinterpret_cast<uchar4*>(P.dst)[size_t(y)*size_t(P.W) + size_t(x)] = make_uchar4(r8, g8, b8, a8);
            } else if (!words) {
                unsigned char* p = P.dst + (size_t(y)*size_t(P.W) + size_t(x))*3; p[0] = r8; p[1] = g8; p[2] = b8;
            } else {
                unsigned char* p = stage + ((ty*J)/S + pr)*RB + (tx/S)*3; p[0] = r8; p[1] = g8; p[2] = b8;
            }
        }
    }
    if (!words) return;
    __syncthreads();
    constexpr int ROWS_OUT = VR_GROUPS*J/S, WPR = RB/4;            // staged rows, 32-bit words per row
    for (int e = tid; e < ROWS_OUT*WPR; e += VR_THREADS) {
        const int row = e/WPR, wc = e - row*WPR;
        const int y = blockIdx.y*ROWS_OUT + row;
        if (y < P.H) {
            unsigned int* out = reinterpret_cast<unsigned int*>(P.dst + (size_t(y)*size_t(P.W) + size_t(blockIdx.x)*PX)*3);
            out[wc] = reinterpret_cast<const unsigned int*>(stage)[e];
        }
    }
}

} // namespace glsl

using namespace glsl;

// Tap offsets dir*walk of visualizer.frag:26-27, evaluated in strict float32 on the host exactly like
// render.cu's build_blur_table (9 angles x 10 walks; the 9th direction coincides with the first and is
// folded into it). Order: ray 0 (10), rays 1-7 (70), the undisplaced first tap. Both constant tables of
// this translation unit are filled: the per-tap global path reads c_blur.
static int build_rows_tables() {
    static bool done[64] = {};
    int device = 0;
    SFB_CUDA(cudaGetDevice(&device));
    if (device < 64 && done[device]) return SFB_OK;
    BlurTable blur;
    RowsTaps rows;
    memset(&rows, 0, sizeof(rows));
    const volatile float TAU_F = 6.2831853071795864f, directions = 8.0f, quality = 10.0f;
    int n = 0;
    for (volatile float angle = 0.0f; angle < TAU_F; angle = angle + TAU_F/directions) {
        const float c = cosf(angle), s = sinf(angle);
        for (volatile float walk = 1.0f/quality; walk <= 1.001f; walk = walk + 1.0f/quality) {
            if (n >= 90) SFB_FAIL(SFB_ESTATE, "blur table overflow: float loop semantics changed");
            blur.tap[n++] = make_float2(c*walk, s*walk);
        }
    }
    if (n != 90) SFB_FAIL(SFB_ESTATE, "blur table has %d taps, expected 90", n);
    blur.tap[90] = make_float2(0.0f, 0.0f); blur.tap[91] = make_float2(0.0f, 0.0f);
    // ray r, walk t = blur.tap[10*r + t]; rays at 0°, 45°, ..., 315° (the 9th, at 360°, is folded into ray 0)
    auto ray = [&](int r, int t) { return blur.tap[10*r + t]; };
    const float tol = 1.0e-6f;
    for (int t = 0; t < 10; t++) {
        rows.hdx[t] = ray(0, t).x; rows.hdx[10 + t] = ray(4, t).x;
        rows.pdx[2*t] = ray(1, t).x; rows.pdx[2*t + 1] = ray(3, t).x;  rows.gdy[1 + t] = ray(1, t).y;
        rows.pdx[20 + 2*t] = ray(5, t).x; rows.pdx[20 + 2*t + 1] = ray(7, t).x; rows.gdy[11 + t] = ray(5, t).y;
        rows.vdy[t] = ray(2, t).y; rows.vdy[10 + t] = ray(6, t).y;
        // the groupings rest on these coincidences of the float32 table
        if (fabsf(ray(0, t).y) > tol || fabsf(ray(4, t).y) > tol || fabsf(ray(2, t).x) > tol || fabsf(ray(6, t).x) > tol
            || fabsf(ray(1, t).y - ray(3, t).y) > tol || fabsf(ray(5, t).y - ray(7, t).y) > tol
            || fabsf(ray(8, t).x - ray(0, t).x) > tol || fabsf(ray(8, t).y - ray(0, t).y) > tol)
            SFB_FAIL(SFB_ESTATE, "blur table symmetry broken at walk %d", t);
    }
    rows.gdy[0] = 0.0f; rows.vdy[20] = 0.0f;
    SFB_CUDA(cudaMemcpyToSymbol(c_blur, &blur, sizeof(blur)));
    SFB_CUDA(cudaMemcpyToSymbol(c_rows, &rows, sizeof(rows)));
    if (device < 64) done[device] = true;
    return SFB_OK;
}


static std::atomic<unsigned int> g_next_slot{0};

template <int S, int J, int WW> static cudaError_t launch_rows(VisRowsParams& VP, cudaStream_t st) {
    static bool configured[64] = {};                      // the attribute is per device
    int device = 0;
    if (cudaError_t e = cudaGetDevice(&device); e != cudaSuccess) return e;
    const size_t table = sizeof(float4)*VR_GROUPS*(VR_HG + VR_MAXQ)*J + sizeof(float)*VR_NI_MAX*VR_COLS;
    const size_t epilogue = sizeof(float4)*J*VR_THREADS + size_t(VR_GROUPS*J)*VR_COLS*3;   // stash + rgb24 staging
    if (device >= 64 || !configured[device]) {
        size_t most = size_t(VR_TEXEL_BYTES)*WW*vr_max_h(WW) + table;
        if (most < epilogue) most = epilogue;
        cudaError_t e = cudaFuncSetAttribute(visualizer_rows_kernel<S, J, WW>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(most));
        if (e != cudaSuccess) return e;
        if (device < 64) configured[device] = true;
    }
    size_t smem = size_t(VR_TEXEL_BYTES)*WW*size_t(VP.win_h) + table;
    if (smem < epilogue) smem = epilogue;
    dim3 block(VR_COLS, VR_GROUPS), grid((VP.R.Wr + VR_COLS - 1)/VR_COLS, (VP.R.Hr + VR_GROUPS*J - 1)/(VR_GROUPS*J));
    const sfb_uniforms& u = VP.R.u;
    VP.K = frame_consts(u.iTime, u.extra[0][0], u.extra[1][0], float(VP.R.tex[0].h));
    VP.slot = 0;
    if (WW == 64) {                     // same stream, in front of the frame: consecutive launches rotate through the slots
        VP.slot = int(g_next_slot.fetch_add(1u) % VR_SLOTS);
        visualizer_frame_consts_kernel<<<1, 32, 0, st>>>(u.iTime, u.extra[0][0], u.extra[1][0], float(VP.R.tex[0].h), VP.slot);
    }
    visualizer_rows_kernel<S, J, WW><<<grid, block, smem, st>>>(VP);
    return cudaSuccess;
}

// Launch planning on the host: the same vis_front the kernel evaluates gives the texel step per fragment.
// Launch planning, pure host code: which variant (J fragment rows per thread) shades this frame and how many
// window rows it stages; J = 0 when the frame does not qualify (the tiled kernel takes it). The same vis_front
// the kernel evaluates gives the texel step per fragment.
static void plan_rows(const RenderParams& P, int* J_out, int* win_h_out, int* win_w_out = nullptr) {
    *J_out = 0; *win_h_out = 0;
    if (win_w_out) *win_w_out = 0;
    const sfb_uniforms& u = P.u;
    const int S = P.ssaa;
    if (!(S == 1 || S == 2 || S == 4)) return;
    // separable camera: 2D ray through the z = 1 plane with the canonical basis (camera.glsl:55-91)
    const bool canonical = u.iCameraProjection == 0
        && u.iCameraRight[0] == 1.0f && u.iCameraRight[1] == 0.0f && u.iCameraRight[2] == 0.0f
        && u.iCameraUpward[0] == 0.0f && u.iCameraUpward[1] == 1.0f && u.iCameraUpward[2] == 0.0f;
    if (!canonical || P.Wr < 2 || P.Hr < 2) return;
    const DevSampler& bg = P.tex[0];
    const float fw = float(bg.w), fh = float(bg.h), hw = float(bg.h)/float(bg.w);
    const float iTime = u.iTime, vol = u.extra[0][0];
    const FrameConsts F = frame_consts(iTime, vol, 0.0f, fh);          // exactly what the kernel will get
    const float zf = F.zf, scale = F.scale;
    const vec2 wobble = mk2(F.wobx, F.woby);
    const vec2 t00 = vis_front(P, 0, 0, fw, fh, hw, wobble, zf).tap;
    const vec2 t10 = vis_front(P, P.Wr - 1, 0, fw, fh, hw, wobble, zf).tap;
    const vec2 t01 = vis_front(P, 0, P.Hr - 1, fw, fh, hw, wobble, zf).tap;
    const double sx = fabs(double(t10.x) - double(t00.x))/double(P.Wr - 1);
    const double sy = fabs(double(t01.y) - double(t00.y))/double(P.Hr - 1);
    if (!(sx == sx && sy == sy && scale == scale) || !(sx < 1e6 && sy < 1e6 && scale < 1e6)) return;
    if (t10.y != t00.y || t01.x != t00.x) return;                      // not separable after all
    // J fragment rows share the 4 texel rows of a tap: (J - 1) steps + the bilinear footprint must fit in them
    int J = 0;
    for (int candidate : {8, 4, 3, 2})
        if (candidate % S == 0 && double(candidate - 1)*sy <= 1.9) { J = candidate; break; }
    if (J == 0) return;
    const double reach = double(scale)*1.0001 + 1.0;
    const double span_x = (VR_COLS - 1)*sx + 2.0*reach + 9.0;
    const int WW = span_x <= 63.0 ? 64 : 96;
    if (span_x > double(WW - 1)) return;
    const int win_h = int(ceil((VR_GROUPS*J - 1)*sy + 2.0*reach)) + 11;   // + alignment, neighbours, the last row quad
    if (win_h > vr_max_h(WW)) return;
    *J_out = J; *win_h_out = win_h;
    if (win_w_out) *win_w_out = WW;
}

extern "C" int sfb_visualizer_plan(const sfb_uniforms* uniforms, int background_w, int background_h,
                                   int width, int height, int ssaa, int* rows_per_thread, int* window_rows) {
    SFB_REQUIRE(uniforms && rows_per_thread && window_rows, "sfb_visualizer_plan: null argument");
    SFB_REQUIRE(background_w > 0 && background_h > 0 && width > 0 && height > 0 && ssaa >= 1, "sfb_visualizer_plan: bad size");
    RenderParams P{};
    P.u = *uniforms;
    P.tex[0].w = background_w; P.tex[0].h = background_h;
    P.W = width; P.H = height; P.ssaa = ssaa; P.Wr = width*ssaa; P.Hr = height*ssaa;
    P.inv_Wr = 1.0/double(P.Wr); P.inv_Hr = 1.0/double(P.Hr);
    plan_rows(P, rows_per_thread, window_rows);
    return SFB_OK;
}

int sfb_visualizer_rows_launch(const RenderParams& P, cudaStream_t stream, int* launched, int screen_alpha, sfb_tex* background) {
    *launched = 0;
    static const bool disabled = getenv("SFB_NO_ROWS") != nullptr;     // debugging knob: tiled kernel only
    if (disabled) return SFB_OK;
    int J = 0, win_h = 0, WW = 0;
    plan_rows(P, &J, &win_h, &WW);
    if (J == 0) return SFB_OK;
    const int S = P.ssaa;
    if (int e = build_rows_tables()) return e;
    VisRowsParams VP;
    VP.R = P; VP.win_h = win_h; VP.screen_alpha = screen_alpha;
    VP.tmap = background ? sfb_background_tensor_map(background, WW, win_h) : nullptr;
    static const int debug = getenv("SFB_ROWS_DEBUG") ? atoi(getenv("SFB_ROWS_DEBUG")) : 0;
    VP.debug = debug;
    cudaError_t e = cudaErrorInvalidValue;
    #define SFB_ROWS_CASE(s_, j_, w_) if (S == s_ && J == j_ && WW == w_) e = launch_rows<s_, j_, w_>(VP, stream);
    // small texel step per fragment (supersampled exports): 64-texel windows
    SFB_ROWS_CASE(1, 8, 64) SFB_ROWS_CASE(2, 8, 64) SFB_ROWS_CASE(4, 8, 64)
    SFB_ROWS_CASE(1, 4, 64) SFB_ROWS_CASE(2, 4, 64) SFB_ROWS_CASE(4, 4, 64)
    // up to ~1 texel per fragment (ssaa 1 or 2 at the background's own resolution): 96-texel windows, fewer rows per thread
    SFB_ROWS_CASE(1, 4, 96) SFB_ROWS_CASE(1, 3, 96) SFB_ROWS_CASE(1, 2, 96)
    SFB_ROWS_CASE(2, 4, 96) SFB_ROWS_CASE(2, 2, 96)
    #undef SFB_ROWS_CASE
    if (e == cudaErrorInvalidValue) return SFB_OK;                 // no variant for this (S, J, WW): the tiled kernel takes it
    SFB_CUDA(e);
    *launched = (WW == 64) ? 2 : 1;                               // (frame constants +) the frame
    return SFB_OK;
}

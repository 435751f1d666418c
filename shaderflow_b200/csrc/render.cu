// render.cu — the C ABI of the render path, K4 (final.glsl), the sampler probe and the visualizer's tiled kernel;
// the per-scene kernels K3 / fused K3+K4 live in render_kernels.cuh and are instantiated in render_{screen,frame,lanes}.cu.
// Reference call sites replaced: shader.py:367-405 (render / render_to_fbo), scene.py:185-194
// (iFinal / iScreen programs), resources/shaders/fragment/final.glsl.
#include "render_kernels.cuh"
#include "visualizer_tiled.cuh"
#include "visualizer_rows.h"

#include <cudaTypedefs.h>
#include <stdlib.h>

using namespace glsl;

using sfb_render::to_unorm8;
using sfb_render::build_blur_table;

// ------------------------------------------------------------------------------------------------
// K4: fragment/final.glsl:3-33 over an RGBA8 iScreen, LINEAR + CLAMP_TO_EDGE (scene.py:192-193)

struct FinalParams {
    const unsigned char* screen; int Ws, Hs;
    int W, H, subsample, comps;
    unsigned char* dst;
    double inv_W, inv_H;
};

// texture(iScreen, uv).rgb scaled by 255: the weights follow the text (frac of u·W − 0.5, clamped texel indices);
// the unorm8 decode c/255 of the four texels is deferred to ONE multiplication after the taps are summed (a linear
// map commutes with the filter: the result differs from the literal order of operations by float32 rounding
// only, ≈ 1e-7, far inside the 8-bit store) — the literal version spent ~60 % of its instructions in 48 IEEE divisions
SFB_DEV vec3 screen_bilinear255(const FinalParams& P, vec2 uv) {
    const float ub = uv.x*float(P.Ws) - 0.5f, vb = uv.y*float(P.Hs) - 0.5f;
    const float fx = floorf(ub), fy = floorf(vb);
    const float a = ub - fx, b = vb - fy;
    const int i0 = int(fx), j0 = int(fy);
    const int ia = min(max(i0, 0), P.Ws - 1), ib = min(max(i0 + 1, 0), P.Ws - 1);
    const int ja = min(max(j0, 0), P.Hs - 1), jb = min(max(j0 + 1, 0), P.Hs - 1);
    const unsigned int* rowa = reinterpret_cast<const unsigned int*>(P.screen) + size_t(ja)*size_t(P.Ws);
    const unsigned int* rowb = reinterpret_cast<const unsigned int*>(P.screen) + size_t(jb)*size_t(P.Ws);
    // bytes → floats with PRMT + one FADD per channel (widen3): integer → float conversions issue at a quarter rate
    const vec3 c00 = widen3(__ldg(rowa + ia)), c10 = widen3(__ldg(rowa + ib));
    const vec3 c01 = widen3(__ldg(rowb + ia)), c11 = widen3(__ldg(rowb + ib));
    const float w11 = a*b, w10 = a - w11, w01 = b - w11, w00 = (1.0f - a) - w01;
    return c00*w00 + c10*w10 + c01*w01 + c11*w11;
}

// Tile = 32 x 8 output pixels; rgb24 rows are staged in shared memory and leave as 32-bit words
__global__ void __launch_bounds__(256) final_kernel(const __grid_constant__ FinalParams P) {
    __shared__ unsigned int stage[8][32];
    const int x = blockIdx.x*32 + threadIdx.x;
    const int y = blockIdx.y*8 + threadIdx.y;
    const bool inside = (x < P.W) && (y < P.H);
    unsigned int r8 = 0, g8 = 0, b8 = 0;
    if (inside) {
        // astuv of the W×H final target (same rasteriser rule as make_frag: float64 product, rounded once)
        const vec2 astuv = mk2(float((double(x) + 0.5)*P.inv_W), float((double(y) + 0.5)*P.inv_H));
        vec3 rgb;
        if (P.subsample == 1) {
            rgb = screen_bilinear255(P, astuv)*(1.0f/255.0f);
        } else {
            const int kernel = P.subsample;
            vec3 acc = mk3(0.0f);
            const vec2 pixel_size = mk2(1.0f/float(P.W), 1.0f/float(P.H));
            const vec2 corner = astuv - (pixel_size/2.0f);
            const vec2 step = pixel_size/float(kernel);          // loop-invariant in the text too: same value
            const vec2 origin = corner + step/2.0f;
            for (int sx = 0; sx < kernel; sx++)
                for (int sy = 0; sy < kernel; sy++) {
                    const vec2 offset = step*mk2(float(sx), float(sy));
                    acc = acc + screen_bilinear255(P, origin + offset);
                }
            rgb = acc*((1.0f/255.0f)/float(kernel*kernel));
        }
        r8 = to_unorm8(rgb.x); g8 = to_unorm8(rgb.y); b8 = to_unorm8(rgb.z);
    }
    if (P.comps == 4) {
        if (inside) reinterpret_cast<uchar4*>(P.dst)[size_t(y)*size_t(P.W) + size_t(x)] = make_uchar4(r8, g8, b8, 255);
        return;
    }
    const bool words = (P.W % 4 == 0) && (int(blockIdx.x)*32 + 32 <= P.W);
    if (!words) {
        if (inside) { unsigned char* p = P.dst + (size_t(y)*size_t(P.W) + size_t(x))*3; p[0] = r8; p[1] = g8; p[2] = b8; }
        return;
    }
    unsigned char* row = reinterpret_cast<unsigned char*>(stage[threadIdx.y]);
    row[threadIdx.x*3 + 0] = r8; row[threadIdx.x*3 + 1] = g8; row[threadIdx.x*3 + 2] = b8;
    __syncwarp();
    if (threadIdx.x < 24 && y < P.H) {
        unsigned int* out = reinterpret_cast<unsigned int*>(P.dst + (size_t(y)*size_t(P.W) + size_t(blockIdx.x)*32)*3);
        out[threadIdx.x] = stage[threadIdx.y][threadIdx.x];
    }
}

// The reference's DEFAULT export geometry — ssaa 1 (iScreen at the final resolution) with subsample 2 (scene.py:526-528)
// — in its own kernel: the four taps of a pixel sit a quarter texel around its centre, so they read the 3 x 3 texels
// around it and their sum is a separable 3 x 3 filter with weights ((1-a0), a0 + (1-a1), a1) per axis, a0 / a1 being
// the fractional tap positions exactly as final.glsl's float32 arithmetic produces them. A CTA of 32 x 32 pixels (four
// pixel rows per thread) widens its 34 x 34 texels once into shared memory — each texel serves 9 pixels, one tile load
// and one barrier per 1024 pixels — instead of loading and widening 16 texels per pixel. A thread whose taps do not
// land on the expected texels (never, short of NaN) takes the generic path.
constexpr int FINAL_TILE_H = 32;             // pixel rows per CTA: 4 adjacent rows per thread, one tile load and one barrier for 1024 pixels
__global__ void __launch_bounds__(256) final_half_step_kernel(const __grid_constant__ FinalParams P) {
    __shared__ float4 tile[FINAL_TILE_H + 2][34];
    __shared__ unsigned int stage[8][4][24];
    const int tid = threadIdx.y*32 + threadIdx.x;
    const int x0 = blockIdx.x*32, y0 = blockIdx.y*FINAL_TILE_H;
    for (int t = tid; t < (FINAL_TILE_H + 2)*34; t += 256) {
        const int r = t/34, c = t - r*34;
        const int gx = min(max(x0 - 1 + c, 0), P.Ws - 1), gy = min(max(y0 - 1 + r, 0), P.Hs - 1);
        const vec3 v = widen3(__ldg(reinterpret_cast<const unsigned int*>(P.screen) + size_t(gy)*size_t(P.Ws) + size_t(gx)));
        tile[r][c] = make_float4(v.x, v.y, v.z, 0.0f);
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, ly0 = threadIdx.y*4;
    const vec2 pixel_size = mk2(1.0f/float(P.W), 1.0f/float(P.H));
    const vec2 step = pixel_size/2.0f;
    // the column's part of final.glsl:17-28: tap coordinates for x = 0, 1, then texture()'s texel position u*W - 0.5
    const float astu = float((double(x) + 0.5)*P.inv_W);
    const float originx = (astu - pixel_size.x/2.0f) + step.x/2.0f;
    const float ub0 = (originx + step.x*0.0f)*float(P.Ws) - 0.5f, ub1 = (originx + step.x*1.0f)*float(P.Ws) - 0.5f;
    const float fx0 = floorf(ub0), fx1 = floorf(ub1);
    const bool column_ok = int(fx0) == x - 1 && int(fx1) == x;
    const float a0 = ub0 - fx0, a1 = ub1 - fx1;
    const float wx[3] = {1.0f - a0, a0 + (1.0f - a1), a1};
    // horizontal 3-tap sums of the 6 texel rows the thread's 4 pixel rows read (each serves up to 3 of them)
    vec3 h[6];
    #pragma unroll
    for (int r = 0; r < 6; r++) {
        const float4 t0 = tile[ly0 + r][threadIdx.x], t1 = tile[ly0 + r][threadIdx.x + 1], t2 = tile[ly0 + r][threadIdx.x + 2];
        h[r] = mk3(wx[0]*t0.x + wx[1]*t1.x + wx[2]*t2.x, wx[0]*t0.y + wx[1]*t1.y + wx[2]*t2.y, wx[0]*t0.z + wx[1]*t1.z + wx[2]*t2.z);
    }
    const bool words = (P.comps == 3) && (P.W % 4 == 0) && (x0 + 32 <= P.W);
    #pragma unroll
    for (int j = 0; j < 4; j++) {
        const int y = y0 + ly0 + j;
        const bool inside = (x < P.W) && (y < P.H);
        unsigned int r8 = 0, g8 = 0, b8 = 0;
        if (inside) {
            const float astv = float((double(y) + 0.5)*P.inv_H);
            const float originy = (astv - pixel_size.y/2.0f) + step.y/2.0f;
            const float vb0 = (originy + step.y*0.0f)*float(P.Hs) - 0.5f, vb1 = (originy + step.y*1.0f)*float(P.Hs) - 0.5f;
            const float fy0 = floorf(vb0), fy1 = floorf(vb1);
            vec3 rgb;
            if (column_ok && int(fy0) == y - 1 && int(fy1) == y) {
                const float b0 = vb0 - fy0, b1 = vb1 - fy1;
                rgb = ((mk3(0.0f) + h[j]*(1.0f - b0)) + h[j + 1]*(b0 + (1.0f - b1))) + h[j + 2]*b1;
                rgb = rgb*((1.0f/255.0f)/4.0f);
            } else {
                const vec2 origin = mk2(originx, originy);
                vec3 acc = mk3(0.0f);
                for (int sx = 0; sx < 2; sx++)
                    for (int sy = 0; sy < 2; sy++)
                        acc = acc + screen_bilinear255(P, origin + step*mk2(float(sx), float(sy)));
                rgb = acc*((1.0f/255.0f)/4.0f);
            }
            r8 = to_unorm8(rgb.x); g8 = to_unorm8(rgb.y); b8 = to_unorm8(rgb.z);
        }
        if (P.comps == 4) {
            if (inside) reinterpret_cast<uchar4*>(P.dst)[size_t(y)*size_t(P.W) + size_t(x)] = make_uchar4(r8, g8, b8, 255);
        } else if (!words) {
            if (inside) { unsigned char* p = P.dst + (size_t(y)*size_t(P.W) + size_t(x))*3; p[0] = r8; p[1] = g8; p[2] = b8; }
        } else {
            unsigned char* row = reinterpret_cast<unsigned char*>(stage[threadIdx.y][j]);
            row[threadIdx.x*3 + 0] = r8; row[threadIdx.x*3 + 1] = g8; row[threadIdx.x*3 + 2] = b8;
        }
    }
    if (words) {
        // a warp owns 4 pixel rows: their 4 x 96 bytes leave as 96 words through the warp's staging rows
        __syncwarp();
        for (int w = threadIdx.x; w < 96; w += 32) {
            const int j = w/24, c = w - j*24, y = y0 + ly0 + j;
            if (y < P.H)
                reinterpret_cast<unsigned int*>(P.dst + (size_t(y)*size_t(P.W) + size_t(x0))*3)[c] = stage[threadIdx.y][j][c];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// texture() probe (tests of the sampler rules)

template <bool HW>
__global__ void sample_kernel(DevSampler s, const float2* uv, int n, float4* out) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    const vec4 c = texture<HW>(s, mk2(uv[i].x, uv[i].y));
    out[i] = make_float4(c.x, c.y, c.z, c.w);
}

extern "C" int sfb_tex_sample(sfb_tex* tex, const float* uv_dev, int n, int flags, float* out_dev) {
    SFB_REQUIRE(tex && uv_dev && out_dev && n >= 0, "sfb_tex_sample: bad argument");
    if (n == 0) return SFB_OK;
    const DevSampler s = tex->dev();
    if (flags & SFB_FILTER_HARDWARE)
        sample_kernel<true><<<(n + 255)/256, 256, 0, tex->ctx->stream>>>(s, (const float2*)uv_dev, n, (float4*)out_dev);
    else
        sample_kernel<false><<<(n + 255)/256, 256, 0, tex->ctx->stream>>>(s, (const float2*)uv_dev, n, (float4*)out_dev);
    SFB_LAUNCH_CHECK(tex->ctx);
    return SFB_OK;
}

// ------------------------------------------------------------------------------------------------
// Scene registry

static const sfb_scene_info SCENES[SFB_SCENE_COUNT] = {
    {"default",    "shaderflow/resources/shaders/fragment/default.glsl", 0, {}, 0, {}, 0},
    {"shadertoy",  "examples/basic/shaders/shadertoy.frag",               0, {}, 0, {}, 0},
    {"visualizer", "examples/basic/shaders/visualizer.frag", 2, {"iAudioVolume", "iAudioSTD"},
                   3, {"background", "iSpectrogram", "iWaveform"}, 3},
    {"bars",       "examples/basic/shaders/bars.frag",      0, {}, 1, {"iSpectrogram"}, 1},
    {"waveform",   "examples/basic/shaders/waveform.frag",  0, {}, 1, {"iWaveform"}, 1},
    {"mandelbrot", "examples/fractals/shaders/mandelbrot.frag", 0, {}, 0, {}, 0},
    {"tetration",  "examples/fractals/shaders/tetration.frag",  0, {}, 0, {}, 0},
    {"raymarch",   "examples/basic/shaders/raymarch.frag",      0, {}, 0, {}, 0},
    {"multishader_child", "examples/basic/demo.py:74-79",       0, {}, 0, {}, 0},
    {"multishader", "examples/basic/demo.py:83-89",             0, {}, 1, {"child"}, 1},
    {"multipass",  "examples/basic/shaders/multipass.frag",     0, {}, 2, {"background", "iScreen0x0"}, 2},
    {"motionblur", "examples/basic/shaders/motionblur.frag",    1, {"iScreenTemporal"},
                   17, {"background", "iScreen0x0", "iScreen1x0", "iScreen2x0", "iScreen3x0", "iScreen4x0", "iScreen5x0",
                        "iScreen6x0", "iScreen7x0", "iScreen8x0", "iScreen9x0", "iScreen10x0", "iScreen11x0",
                        "iScreen12x0", "iScreen13x0", "iScreen14x0", "iScreen15x0"}, 2},
    {"dynamics",   "examples/basic/demo.py:120-125",            1, {"iShaderDynamics"}, 1, {"background"}, 1},
    {"audio",      "examples/basic/demo.py:150-154",            1, {"iAudioVolume"}, 0, {}, 0},
    {"life_simulation", "examples/basic/shaders/life/simulation.glsl", 2, {"iLifePeriod", "iLifeSize"}, 1, {"iLife1x0"}, 1},
    {"life_visuals", "examples/basic/shaders/life/visuals.glsl", 0, {},
                   5, {"iLife0x0", "iLife1x0", "iLife2x0", "iLife3x0", "iLife4x0"}, 5},
    {"piano",      "examples/shaders/piano.frag (this repository)", 6,
                   {"iPianoDynamic", "iPianoExtra", "iPianoHeight", "iPianoBlackRatio", "iPianoRollTime", "iPianoLimit"},
                   3, {"iPianoKeys", "iPianoChan", "iPianoRoll"}, 3},
};

extern "C" int sfb_scene_lookup(const char* name, int* scene) {
    SFB_REQUIRE(name && scene, "sfb_scene_lookup: null argument");
    for (int i = 0; i < SFB_SCENE_COUNT; i++)
        if (std::string(SCENES[i].name) == name) { *scene = i; return SFB_OK; }
    SFB_FAIL(SFB_ENOTFOUND, "sfb_scene_lookup: no built-in scene named '%s'", name);
}

extern "C" int sfb_scene_info_get(int scene, sfb_scene_info* info) {
    SFB_REQUIRE(info && scene >= 0 && scene < SFB_SCENE_COUNT, "sfb_scene_info_get: bad scene %d", scene);
    *info = SCENES[scene];
    return SFB_OK;
}

// ------------------------------------------------------------------------------------------------
// Launchers

static int fill_params(RenderParams& P, const char* who, int scene, const sfb_uniforms* uniforms,
                       sfb_tex* const* samplers, int n_samplers, int flags) {
    SFB_REQUIRE(uniforms, "%s: null uniforms", who);
    if (scene >= SFB_SCENE_PROGRAM_BASE) {                        // a run-time compiled program (jit.cu)
        int device = 0;
        SFB_CUDA(cudaGetDevice(&device));
        const int needed = sfb_program_samplers(scene, device);
        SFB_REQUIRE(needed >= 0, "%s: %d is not a loaded program of device %d", who, scene, device);
        SFB_REQUIRE(n_samplers >= needed && n_samplers <= SFB_MAX_SAMPLERS, "%s: the program reads %d samplers, got %d", who, needed, n_samplers);
        P.u = *uniforms;
        for (int i = 0; i < n_samplers; i++) {
            SFB_REQUIRE(samplers && samplers[i], "%s: sampler %d is null", who, i);
            P.tex[i] = samplers[i]->dev();
            SFB_REQUIRE(P.tex[i].lin, "%s: sampler %d has no storage", who, i);
        }
        P.fast = 0;
        return SFB_OK;
    }
    SFB_REQUIRE(scene >= 0 && scene < SFB_SCENE_COUNT, "%s: bad scene %d", who, scene);
    SFB_REQUIRE(n_samplers >= SCENES[scene].n_required && n_samplers <= SFB_MAX_SAMPLERS,
        "%s: scene '%s' needs %d samplers, got %d", who, SCENES[scene].name, SCENES[scene].n_required, n_samplers);
    if (scene == SFB_SCENE_MOTIONBLUR)
        SFB_REQUIRE(n_samplers >= 1 + int(uniforms->extra[0][0]),
            "%s: motionblur averages iScreenTemporal = %d frames but only %d history samplers are bound",
            who, int(uniforms->extra[0][0]), n_samplers - 1);
    P.u = *uniforms;
    for (int i = 0; i < n_samplers; i++) {
        SFB_REQUIRE(samplers && samplers[i], "%s: sampler %d is null", who, i);
        P.tex[i] = samplers[i]->dev();
        SFB_REQUIRE(P.tex[i].lin, "%s: sampler %d has no storage", who, i);
    }
    P.fast = 0;
    if ((scene == SFB_SCENE_MANDELBROT || scene == SFB_SCENE_TETRATION) && !(flags & SFB_RENDER_LITERAL))
        P.fast = 1;               // closed-form interior tests / one exp of the combined exponent (scenes.cuh)
    if (scene == SFB_SCENE_VISUALIZER && !(flags & SFB_RENDER_LITERAL)) {
        const DevSampler& bg = P.tex[0];
        if (bg.dtype == SFB_DTYPE_U8 && bg.padded == 4 && bg.filter == SFB_FILTER_LINEAR && bg.w >= 2 && bg.h >= 2
            && bg.w < (1 << 21) && bg.h < (1 << 21)) {
            if (int e = build_blur_table()) return e;
            P.fast = 1;
        }
    }
    return SFB_OK;
}

// 2D tensor map over the background's linear mirror, one RGBA8 texel = one uint32 element, box box_w x box_h
// (cached per texture and box shape; the descriptor is copied to device memory once)
const void* sfb_background_tensor_map(sfb_tex* t, int box_w, int box_h) {
    static const bool disabled = getenv("SFB_NO_TMA") != nullptr;      // debugging knob
    if (disabled) return nullptr;
    if (t->external || t->dtype != SFB_DTYPE_U8 || t->padded != 4 || (t->w % 4) != 0 || t->w < box_w || t->h < box_h
        || box_w > 256 || box_h > 256)
        return nullptr;
    for (const auto& m : t->tmaps)
        if (m.box_w == box_w && m.box_h == box_h) return m.dev;
    static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
    if (!encode) {
        void* fn = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return nullptr;
        encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    }
    alignas(64) CUtensorMap map;
    cuuint64_t dims[2] = {cuuint64_t(t->w), cuuint64_t(t->h)};
    cuuint64_t strides[1] = {cuuint64_t(t->w)*4};
    cuuint32_t box[2] = {cuuint32_t(box_w), cuuint32_t(box_h)}, estr[2] = {1, 1};
    void* dev = nullptr;
    const CUresult rc = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, t->lin, dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc == CUDA_SUCCESS && (cudaMalloc(&dev, sizeof(CUtensorMap)) != cudaSuccess
                               || cudaMemcpy(dev, &map, sizeof(CUtensorMap), cudaMemcpyHostToDevice) != cudaSuccess)) {
        if (dev) cudaFree(dev);
        dev = nullptr;
    }
    t->tmaps.push_back({box_w, box_h, dev});       // a failed encoding is remembered too (dev == NULL)
    return dev;
}
static const CUtensorMap* background_tensor_map(sfb_tex* t) {
    return static_cast<const CUtensorMap*>(sfb_background_tensor_map(t, VT_TMA_W, VT_TMA_H));
}

// visualizer_tiled_kernel<ssaa> is instantiated in visualizer_tiled_{1,2,3,4}.cu
int sfb_launch_visualizer_tiled_1(const VisualizerParams& VP, cudaStream_t st);
int sfb_launch_visualizer_tiled_2(const VisualizerParams& VP, cudaStream_t st);
int sfb_launch_visualizer_tiled_3(const VisualizerParams& VP, cudaStream_t st);
int sfb_launch_visualizer_tiled_4(const VisualizerParams& VP, cudaStream_t st);

// The iScreen pass of an unfused visualizer export (the reference's default ssaa=1, subsample=2 blends
// neighbours in final.glsl, so K3 and K4 stay separate): the tiled kernel with one fragment per thread and an
// RGBA8 store that keeps fragColor.a
static int visualizer_screen_pass(sfb_ctx* ctx, const RenderParams& P, sfb_tex* background, int flags) {
    VisualizerParams VP;
    VP.R = P;
    VP.R.W = P.Wr; VP.R.H = P.Hr; VP.R.ssaa = 1; VP.R.subsample = 1; VP.R.comps = 4;
    if (!(flags & SFB_RENDER_TILED)) {            // the separable kernel, one fragment per "output pixel", alpha kept
        int launched = 0;
        if (int e = sfb_visualizer_rows_launch(VP.R, ctx->stream, &launched, 1, background)) return e;
        if (launched) { ctx->launches += launched - 1; SFB_LAUNCH_CHECK(ctx); return SFB_OK; }
    }
    VP.tmap = background_tensor_map(background);
    VP.use_tma = VP.tmap ? 1 : 0;
    VP.screen_alpha = 1;
    if (int e = sfb_launch_visualizer_tiled_1(VP, ctx->stream)) return e;
    SFB_LAUNCH_CHECK(ctx);
    return SFB_OK;
}

extern "C" int sfb_render_screen(sfb_ctx* ctx, int scene, const sfb_uniforms* uniforms,
                                 sfb_tex* const* samplers, int n_samplers, int flags,
                                 int target_w, int target_h, void* dst_rgba8_dev, float* dst_f32_dev) {
    SFB_REQUIRE(ctx && dst_rgba8_dev, "sfb_render_screen: null ctx or destination");
    SFB_REQUIRE(target_w > 0 && target_h > 0, "sfb_render_screen: bad target %dx%d", target_w, target_h);
    RenderParams P{};
    if (int e = fill_params(P, "sfb_render_screen", scene, uniforms, samplers, n_samplers, flags)) return e;
    P.Wr = target_w; P.Hr = target_h;
    P.W = int(uniforms->iResolution[0]); P.H = int(uniforms->iResolution[1]);
    P.inv_Wr = 1.0/double(target_w); P.inv_Hr = 1.0/double(target_h);
    P.dst = static_cast<unsigned char*>(dst_rgba8_dev); P.dst_f32 = dst_f32_dev;
    P.dst_dtype = SFB_DTYPE_U8; P.dst_padded = 4;
    if (scene == SFB_SCENE_VISUALIZER && P.fast && !(flags & SFB_FILTER_HARDWARE) && !dst_f32_dev)
        return visualizer_screen_pass(ctx, P, samplers[0], flags);
    if (int e = (scene >= SFB_SCENE_PROGRAM_BASE) ? sfb_program_launch(scene, 0, P, ctx->stream)
                                                  : sfb_launch_screen(scene, (flags & SFB_FILTER_HARDWARE) != 0, P, ctx->stream)) return e;
    SFB_LAUNCH_CHECK(ctx);
    return SFB_OK;
}

extern "C" int sfb_render_target(sfb_ctx* ctx, int scene, const sfb_uniforms* uniforms,
                                 sfb_tex* const* samplers, int n_samplers, int flags, sfb_tex* target) {
    SFB_REQUIRE(ctx && target, "sfb_render_target: null ctx or target");
    SFB_REQUIRE(!target->external && target->lin, "sfb_render_target: the target has no storage of its own");
    // The target may also be in `samplers` (a layer-0 pass of multipass.frag binds iScreen0x0 without reading
    // it): like GL, that is only undefined if the pass actually samples it.
    RenderParams P{};
    if (int e = fill_params(P, "sfb_render_target", scene, uniforms, samplers, n_samplers, flags)) return e;
    P.Wr = target->w; P.Hr = target->h;
    P.W = int(uniforms->iResolution[0]); P.H = int(uniforms->iResolution[1]);
    P.inv_Wr = 1.0/double(P.Wr); P.inv_Hr = 1.0/double(P.Hr);
    P.dst = static_cast<unsigned char*>(target->lin); P.dst_f32 = nullptr;
    P.dst_dtype = target->dtype; P.dst_padded = target->padded;
    target->array_stale = true;                   // the cudaArray copy is old: sample through the linear mirror
    if (scene == SFB_SCENE_VISUALIZER && P.fast && !(flags & SFB_FILTER_HARDWARE) && target->dtype == SFB_DTYPE_U8 && target->padded == 4)
        return visualizer_screen_pass(ctx, P, samplers[0], flags);
    if (int e = (scene >= SFB_SCENE_PROGRAM_BASE) ? sfb_program_launch(scene, 0, P, ctx->stream)
                                                  : sfb_launch_screen(scene, (flags & SFB_FILTER_HARDWARE) != 0, P, ctx->stream)) return e;
    SFB_LAUNCH_CHECK(ctx);
    return SFB_OK;
}

extern "C" int sfb_render_final(sfb_ctx* ctx, const void* screen_rgba8_dev, int screen_w, int screen_h,
                                int width, int height, int subsample, int components, void* dst_dev) {
    SFB_REQUIRE(ctx && screen_rgba8_dev && dst_dev, "sfb_render_final: null argument");
    SFB_REQUIRE(screen_w > 0 && screen_h > 0 && width > 0 && height > 0, "sfb_render_final: bad size");
    SFB_REQUIRE(subsample >= 1 && subsample <= 16, "sfb_render_final: subsample %d out of range", subsample);
    SFB_REQUIRE(components == 3 || components == 4, "sfb_render_final: components must be 3 or 4");
    FinalParams P{static_cast<const unsigned char*>(screen_rgba8_dev), screen_w, screen_h,
                  width, height, subsample, components, static_cast<unsigned char*>(dst_dev), 1.0/double(width), 1.0/double(height)};
    dim3 block(32, 8), grid((width + 31)/32, (height + 7)/8);
    static const bool generic_only = getenv("SFB_FINAL_GENERIC") != nullptr;       // debugging knob
    if (subsample == 2 && screen_w == width && screen_h == height && !generic_only) {
        dim3 tall((width + 31)/32, (height + FINAL_TILE_H - 1)/FINAL_TILE_H);
        final_half_step_kernel<<<tall, block, 0, ctx->stream>>>(P);               // the reference's default export
    }
    else
        final_kernel<<<grid, block, 0, ctx->stream>>>(P);
    SFB_LAUNCH_CHECK(ctx);
    return SFB_OK;
}

static int render_frame(sfb_ctx* ctx, int scene, const sfb_uniforms* uniforms,
                        sfb_tex* const* samplers, int n_samplers, int flags,
                        int width, int height, int ssaa, int subsample, int components, void* dst_dev, float* screen_f32_dev) {
    SFB_REQUIRE(ctx && dst_dev, "sfb_render_frame: null ctx or destination");
    SFB_REQUIRE(width > 0 && height > 0, "sfb_render_frame: bad size %dx%d", width, height);
    SFB_REQUIRE(ssaa >= 1 && ssaa <= 16, "sfb_render_frame: ssaa %d out of range", ssaa);
    SFB_REQUIRE(subsample == ssaa || 2*subsample == ssaa,
        "sfb_render_frame: final.glsl is a box filter only for subsample == ssaa or ssaa/2 "
        "(got ssaa=%d subsample=%d); use sfb_render_screen + sfb_render_final", ssaa, subsample);
    SFB_REQUIRE(components == 3 || components == 4, "sfb_render_frame: components must be 3 or 4");
    RenderParams P{};
    if (int e = fill_params(P, "sfb_render_frame", scene, uniforms, samplers, n_samplers, flags)) return e;
    P.W = width; P.H = height; P.ssaa = ssaa; P.subsample = subsample; P.comps = components;
    P.Wr = width*ssaa; P.Hr = height*ssaa;
    P.inv_Wr = 1.0/double(P.Wr); P.inv_Hr = 1.0/double(P.Hr);
    P.dst = static_cast<unsigned char*>(dst_dev); P.dst_f32 = screen_f32_dev;
    if (scene == SFB_SCENE_VISUALIZER && P.fast && !(flags & SFB_FILTER_HARDWARE) && ssaa <= 4) {
        // production path of the headline scene: the separable kernel when the camera allows it, ...
        if (!(flags & SFB_RENDER_TILED)) {
            int launched = 0;
            if (int e = sfb_visualizer_rows_launch(P, ctx->stream, &launched, 0, samplers[0])) return e;
            if (launched) { ctx->launches += launched - 1; SFB_LAUNCH_CHECK(ctx); return SFB_OK; }
        }
        // ... else one thread per output pixel over a shared-memory window of the background
        VisualizerParams VP;
        VP.R = P;
        VP.tmap = background_tensor_map(samplers[0]);
        VP.use_tma = VP.tmap ? 1 : 0;
        VP.screen_alpha = 0;
        int e;
        switch (ssaa) {
            case 1: e = sfb_launch_visualizer_tiled_1(VP, ctx->stream); break;
            case 2: e = sfb_launch_visualizer_tiled_2(VP, ctx->stream); break;
            case 3: e = sfb_launch_visualizer_tiled_3(VP, ctx->stream); break;
            default: e = sfb_launch_visualizer_tiled_4(VP, ctx->stream); break;
        }
        if (e) return e;
        SFB_LAUNCH_CHECK(ctx);
        return SFB_OK;
    }
    const bool hw = (flags & SFB_FILTER_HARDWARE) != 0;
    if (scene >= SFB_SCENE_PROGRAM_BASE) {
        if (int e = sfb_program_launch(scene, 1, P, ctx->stream)) return e;
        SFB_LAUNCH_CHECK(ctx);
        return SFB_OK;
    }
    // the ALU-bound scenes at ssaa 2 / 4: one lane per sub-sample (render_lanes.cu); everything else one thread per pixel
    if ((flags & SFB_RENDER_LITERAL) || !sfb_launch_frame_lanes(scene, hw, P, ctx->stream))
        if (int e = sfb_launch_frame(scene, hw, P, ctx->stream)) return e;
    SFB_LAUNCH_CHECK(ctx);
    return SFB_OK;
}

extern "C" int sfb_render_frame(sfb_ctx* ctx, int scene, const sfb_uniforms* uniforms,
                                sfb_tex* const* samplers, int n_samplers, int flags,
                                int width, int height, int ssaa, int subsample, int components, void* dst_dev) {
    return render_frame(ctx, scene, uniforms, samplers, n_samplers, flags, width, height, ssaa, subsample, components, dst_dev, nullptr);
}

extern "C" int sfb_render_frame_probe(sfb_ctx* ctx, int scene, const sfb_uniforms* uniforms,
                                      sfb_tex* const* samplers, int n_samplers, int flags,
                                      int width, int height, int ssaa, int subsample, int components, void* dst_dev,
                                      float* screen_f32_dev) {
    SFB_REQUIRE(screen_f32_dev, "sfb_render_frame_probe: null float target");
    return render_frame(ctx, scene, uniforms, samplers, n_samplers, flags, width, height, ssaa, subsample, components, dst_dev, screen_f32_dev);
}

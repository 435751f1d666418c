// sfb_internal.h — shared between the translation units of libsfb200.so (not part of the ABI)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <string.h>

#include "../../include/sfb200.h"

// Thread-local error message (sfb_last_error)
void sfb_set_error(const char* fmt, ...);

#define SFB_FAIL(code, ...) do { sfb_set_error(__VA_ARGS__); return (code); } while (0)
#define SFB_REQUIRE(cond, ...) do { if (!(cond)) SFB_FAIL(SFB_EINVAL, __VA_ARGS__); } while (0)
#define SFB_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { \
    sfb_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    return SFB_ECUDA; } } while (0)
#define SFB_LAUNCH_CHECK(ctx) do { (ctx)->launches++; SFB_CUDA(cudaGetLastError()); } while (0)

struct sfb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    uint64_t launches = 0;
    int sm_count = 148;
    // Lazily built device tables of the STFT (twiddles per fft_n, windows per (kind, fft_n))
    float2* twiddle[16] = {};
    float*  window[3][16] = {};
    // Scratch for sfb_audio_track
    void*  scratch = nullptr;
    size_t scratch_bytes = 0;
    // Sink ring buffers kept across sfb_pipe_open/close: pinned host + device frame pairs (cudaHostAlloc
    // of a 4K frame ring costs tens of ms, which a short export would pay every time)
    struct RingBuffer { void* host; void* dev; size_t bytes; };
    std::vector<RingBuffer> ring_pool;
};

int sfb_ctx_scratch(sfb_ctx* ctx, size_t bytes, void** out);

#include "render_params.h"

struct sfb_tex {
    sfb_ctx* ctx = nullptr;
    cudaArray_t array = nullptr;
    cudaTextureObject_t obj = 0;
    void* lin = nullptr;
    const void* external = nullptr;
    // CUtensorMaps over `lin` (2D uint32 views, one per box shape the visualizer kernels stage their windows with);
    // dev = device copy of the descriptor (the kernels read it from global memory), NULL when encoding failed
    struct TensorMap { int box_w, box_h; void* dev; };
    std::vector<TensorMap> tmaps;
    bool array_stale = false;          // rendered into through sfb_tex_storage: cudaArray copy is old
    int w = 0, h = 0, comps = 0, padded = 0, dtype = 0, filter = 0, rx = 1, ry = 1;
    size_t texel_bytes() const { return size_t(padded)*(dtype == SFB_DTYPE_U8 ? 1 : (dtype == SFB_DTYPE_F16 ? 2 : 4)); }
    DevSampler dev() const {
        DevSampler s;
        s.hw = (external || array_stale) ? 0ull : (unsigned long long)obj;
        s.lin = external ? external : lin;
        s.w = w; s.h = h; s.padded = padded; s.comps = comps; s.dtype = dtype; s.filter = filter; s.rx = rx; s.ry = ry;
        return s;
    }
};

// Run-time compiled programs (jit.cu): scene ids >= SFB_SCENE_PROGRAM_BASE
int sfb_program_samplers(int scene, int device);          // samplers the program reads; -1 = not a program of this device
int sfb_program_launch(int scene, int kind, const RenderParams& P, cudaStream_t stream);   // kind 0 screen, 1 fused frame

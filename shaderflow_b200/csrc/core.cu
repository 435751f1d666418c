// core.cu — context, errors, host-side time base, textures.
#include "sfb_internal.h"

#include <cfenv>
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <vector>

// ------------------------------------------------------------------------------------------------
// Errors

static thread_local char g_error[1024] = "";

void sfb_set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

extern "C" const char* sfb_last_error(void) { return g_error; }
extern "C" int sfb_version(void) { return SFB_VERSION; }

extern "C" int sfb_device_count(int* count) {
    SFB_REQUIRE(count, "sfb_device_count: null argument");
    *count = 0;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) { *count = 0; SFB_FAIL(SFB_ECUDA, "no CUDA device: %s", cudaGetErrorString(e)); }
    return SFB_OK;
}

// ------------------------------------------------------------------------------------------------
// Context

extern "C" int sfb_ctx_create(int device, void* stream, sfb_ctx** out) {
    SFB_REQUIRE(out, "sfb_ctx_create: null out");
    int n = 0;
    if (int e = sfb_device_count(&n)) return e;
    if (device < 0 || device >= n)
        SFB_FAIL(SFB_ECUDA, "sfb_ctx_create: device %d not available (%d visible); there is no CPU fallback", device, n);
    SFB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    SFB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        SFB_FAIL(SFB_ECUDA, "sfb_ctx_create: device %d is sm_%d%d; libsfb200 is built for sm_100a only", device, prop.major, prop.minor);
    sfb_ctx* ctx = new sfb_ctx();
    ctx->device = device;
    ctx->stream = static_cast<cudaStream_t>(stream);
    ctx->sm_count = prop.multiProcessorCount;
    *out = ctx;
    return SFB_OK;
}

extern "C" int sfb_ctx_destroy(sfb_ctx* ctx) {
    if (!ctx) return SFB_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& t : ctx->twiddle) if (t) cudaFree(t);
    for (auto& k : ctx->window) for (auto& w : k) if (w) cudaFree(w);
    if (ctx->scratch) cudaFree(ctx->scratch);
    for (auto& b : ctx->ring_pool) { cudaFreeHost(b.host); cudaFree(b.dev); }
    delete ctx;
    return SFB_OK;
}

extern "C" int sfb_ctx_set_stream(sfb_ctx* ctx, void* stream) {
    SFB_REQUIRE(ctx, "sfb_ctx_set_stream: null ctx");
    ctx->stream = static_cast<cudaStream_t>(stream);
    return SFB_OK;
}

extern "C" int sfb_sync(sfb_ctx* ctx) {
    SFB_REQUIRE(ctx, "sfb_sync: null ctx");
    SFB_CUDA(cudaStreamSynchronize(ctx->stream));
    return SFB_OK;
}

extern "C" int sfb_launch_count(sfb_ctx* ctx, uint64_t* count) {
    SFB_REQUIRE(ctx && count, "sfb_launch_count: null argument");
    *count = ctx->launches;
    return SFB_OK;
}

extern "C" int sfb_ctx_enable_peer(sfb_ctx* ctx, int peer_device) {
    SFB_REQUIRE(ctx, "sfb_ctx_enable_peer: null ctx");
    if (peer_device == ctx->device) return SFB_OK;
    SFB_CUDA(cudaSetDevice(ctx->device));
    int can = 0;
    SFB_CUDA(cudaDeviceCanAccessPeer(&can, ctx->device, peer_device));
    if (!can) SFB_FAIL(SFB_ECUDA, "sfb_ctx_enable_peer: device %d cannot access device %d", ctx->device, peer_device);
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); e = cudaSuccess; }
    SFB_CUDA(e);
    return SFB_OK;
}

int sfb_ctx_scratch(sfb_ctx* ctx, size_t bytes, void** out) {
    if (bytes > ctx->scratch_bytes) {
        SFB_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->scratch) SFB_CUDA(cudaFree(ctx->scratch));
        ctx->scratch = nullptr; ctx->scratch_bytes = 0;
        SFB_CUDA(cudaMalloc(&ctx->scratch, bytes));
        ctx->scratch_bytes = bytes;
    }
    *out = ctx->scratch;
    return SFB_OK;
}

// ------------------------------------------------------------------------------------------------
// Time base (host). scheduler.py:86-89,134-173 + scene.py:476-479 + ffmpeg.py:1306-1330.
// Python's round() is round-half-even on doubles == nearbyint under FE_TONEAREST.

extern "C" int sfb_frame_clock(int n_frames, double fps, double speed, int samplerate, int channels,
                               int64_t total_samples, double* time_host, double* dt_host, int64_t* tell_host) {
    SFB_REQUIRE(n_frames >= 0 && fps > 0.0, "sfb_frame_clock: bad n_frames/fps");
    SFB_REQUIRE(!tell_host || (samplerate > 0 && channels > 0), "sfb_frame_clock: bad samplerate/channels");
    const int old_round = fegetround();
    fesetround(FE_TONEAREST);
    const double period = 1.0/fps;
    double last_call = 0.0 - period, next_call = 0.0;
    double t = 0.0, sdt = 0.0, srdt = 0.0;
    // reader state
    const int64_t block = 4*int64_t(channels);
    const double bps = double(block)*double(samplerate);
    const int64_t limit = (total_samples < 0) ? -1 : total_samples*block;
    double target = 0.0; int64_t read = 0;
    for (int k = 0; k < n_frames; k++) {
        if (time_host) time_host[k] = t;
        if (dt_host)   dt_host[k] = sdt;
        if (tell_host) {
            if (limit < 0 || read < limit) {
                target += srdt;                                   // reader.chunk = scene.rdt
                double length = (target - (double(read)/bps))*bps;
                int64_t bytes = block*int64_t(std::nearbyint(length/double(block)));
                if (bytes < block) bytes = block;
                if (limit >= 0 && bytes > limit - read) bytes = limit - read;
                read += bytes;
            }
            tell_host[k] = read/block;
        }
        const double now = next_call;
        const double passed = now - last_call;
        last_call = now;
        while (next_call <= now) next_call += period;
        sdt = passed*speed; srdt = passed;
        t += sdt;
    }
    fesetround(old_round);
    return SFB_OK;
}

// ------------------------------------------------------------------------------------------------
// Textures (texture.py:250-283,313-338)

static int tex_build_object(sfb_tex* t) {
    if (t->obj) { cudaDestroyTextureObject(t->obj); t->obj = 0; }
    cudaResourceDesc res{}; res.resType = cudaResourceTypeArray; res.res.array.array = t->array;
    cudaTextureDesc td{};
    td.addressMode[0] = t->rx ? cudaAddressModeWrap : cudaAddressModeClamp;
    td.addressMode[1] = t->ry ? cudaAddressModeWrap : cudaAddressModeClamp;
    td.filterMode = (t->filter == SFB_FILTER_LINEAR) ? cudaFilterModeLinear : cudaFilterModePoint;
    td.readMode = (t->dtype == SFB_DTYPE_U8) ? cudaReadModeNormalizedFloat : cudaReadModeElementType;
    td.normalizedCoords = 1;
    SFB_CUDA(cudaCreateTextureObject(&t->obj, &res, &td, nullptr));
    return SFB_OK;
}

extern "C" int sfb_tex_create(sfb_ctx* ctx, int width, int height, int components, int dtype,
                              int filter, int repeat_x, int repeat_y, sfb_tex** out) {
    SFB_REQUIRE(ctx && out, "sfb_tex_create: null argument");
    SFB_REQUIRE(width > 0 && height > 0, "sfb_tex_create: bad size %dx%d", width, height);
    SFB_REQUIRE(components >= 1 && components <= 4, "sfb_tex_create: components %d not in 1..4", components);
    SFB_REQUIRE(dtype == SFB_DTYPE_U8 || dtype == SFB_DTYPE_F32 || dtype == SFB_DTYPE_F16, "sfb_tex_create: bad dtype %d", dtype);
    // The reference raises when a texture exceeds GL_MAX_VIEWPORT_DIMS (texture.py:251-252); the
    // equivalent limit here is the 2D cudaArray extent
    cudaDeviceProp prop; SFB_CUDA(cudaGetDeviceProperties(&prop, ctx->device));
    if (width > prop.maxTexture2D[0] || height > prop.maxTexture2D[1])
        SFB_FAIL(SFB_EINVAL, "Texture size too large for this CUDA device: (%d, %d) > (%d, %d)",
                 width, height, prop.maxTexture2D[0], prop.maxTexture2D[1]);
    sfb_tex* t = new sfb_tex();
    t->ctx = ctx; t->w = width; t->h = height; t->comps = components;
    t->padded = (components == 3) ? 4 : components;
    t->dtype = dtype; t->filter = filter; t->rx = repeat_x ? 1 : 0; t->ry = repeat_y ? 1 : 0;
    const int bits = (dtype == SFB_DTYPE_U8) ? 8 : (dtype == SFB_DTYPE_F16 ? 16 : 32);
    const cudaChannelFormatKind kind = (dtype == SFB_DTYPE_U8) ? cudaChannelFormatKindUnsigned : cudaChannelFormatKindFloat;
    cudaChannelFormatDesc fmt = cudaCreateChannelDesc(bits, t->padded >= 2 ? bits : 0, t->padded == 4 ? bits : 0,
                                                      t->padded == 4 ? bits : 0, kind);
    cudaError_t e = cudaMallocArray(&t->array, &fmt, size_t(width), size_t(height));
    if (e != cudaSuccess) { delete t; SFB_FAIL(SFB_ECUDA, "cudaMallocArray(%dx%d): %s", width, height, cudaGetErrorString(e)); }
    const size_t bytes = size_t(width)*size_t(height)*t->texel_bytes();
    e = cudaMalloc(&t->lin, bytes);
    if (e != cudaSuccess) { cudaFreeArray(t->array); delete t; SFB_FAIL(SFB_ENOMEM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)); }
    SFB_CUDA(cudaMemsetAsync(t->lin, 0, bytes, ctx->stream));
    SFB_CUDA(cudaMemcpy2DToArrayAsync(t->array, 0, 0, t->lin, size_t(width)*t->texel_bytes(),
                                      size_t(width)*t->texel_bytes(), size_t(height), cudaMemcpyDeviceToDevice, ctx->stream));
    if (int err = tex_build_object(t)) { sfb_tex_destroy(t); return err; }
    *out = t;
    return SFB_OK;
}

extern "C" int sfb_tex_destroy(sfb_tex* t) {
    if (!t) return SFB_OK;
    cudaStreamSynchronize(t->ctx->stream);
    if (t->obj) cudaDestroyTextureObject(t->obj);
    if (t->array) cudaFreeArray(t->array);
    if (t->lin) cudaFree(t->lin);
    for (auto& m : t->tmaps) if (m.dev) cudaFree(m.dev);
    delete t;
    return SFB_OK;
}

extern "C" int sfb_tex_set_sampling(sfb_tex* t, int filter, int repeat_x, int repeat_y) {
    SFB_REQUIRE(t, "sfb_tex_set_sampling: null texture");
    t->filter = filter; t->rx = repeat_x ? 1 : 0; t->ry = repeat_y ? 1 : 0;
    return tex_build_object(t);
}

// Expands w*h*comps tightly packed host/device elements into the padded texel layout on the device
template <typename T>
__global__ void pad_rgb_kernel(const T* src, T* dst, size_t n, T alpha) {
    size_t i = blockIdx.x*size_t(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    dst[4*i + 0] = src[3*i + 0]; dst[4*i + 1] = src[3*i + 1]; dst[4*i + 2] = src[3*i + 2]; dst[4*i + 3] = alpha;
}

extern "C" int sfb_tex_write(sfb_tex* t, const void* data, int on_device, int x, int y, int w, int h) {
    SFB_REQUIRE(t && data, "sfb_tex_write: null argument");
    SFB_REQUIRE(x >= 0 && y >= 0 && w > 0 && h > 0 && x + w <= t->w && y + h <= t->h,
        "sfb_tex_write: viewport (%d,%d,%d,%d) outside %dx%d", x, y, w, h, t->w, t->h);
    sfb_ctx* ctx = t->ctx;
    const size_t esz = (t->dtype == SFB_DTYPE_U8) ? 1 : (t->dtype == SFB_DTYPE_F16 ? 2 : 4);
    const size_t texel = t->texel_bytes();
    const size_t n = size_t(w)*size_t(h);
    const void* packed = data;           // device pointer to w*h padded texels once staged
    void* staged = nullptr;
    const size_t src_bytes = n*size_t(t->comps)*esz;
    if (!on_device || t->comps == 3) {
        // stage: [src copy][padded copy]
        if (int e = sfb_ctx_scratch(ctx, src_bytes + n*texel + 256, &staged)) return e;
        char* src_dev = static_cast<char*>(staged);
        char* pad_dev = src_dev + ((src_bytes + 255)/256)*256;
        const void* src = data;
        if (!on_device) {
            // pageable source: cudaMemcpyAsync returns once the source has been consumed
            SFB_CUDA(cudaMemcpyAsync(src_dev, data, src_bytes, cudaMemcpyHostToDevice, ctx->stream));
            src = src_dev;
        }
        if (t->comps == 3) {
            const int threads = 256; const unsigned blocks = unsigned((n + threads - 1)/threads);
            if (esz == 1) pad_rgb_kernel<unsigned char><<<blocks, threads, 0, ctx->stream>>>((const unsigned char*)src, (unsigned char*)pad_dev, n, 255);
            else if (esz == 2) pad_rgb_kernel<unsigned short><<<blocks, threads, 0, ctx->stream>>>((const unsigned short*)src, (unsigned short*)pad_dev, n, 0x3C00);
            else pad_rgb_kernel<unsigned int><<<blocks, threads, 0, ctx->stream>>>((const unsigned int*)src, (unsigned int*)pad_dev, n, 0x3F800000u);
            SFB_LAUNCH_CHECK(ctx);
            packed = pad_dev;
        } else {
            packed = src;
        }
    }
    // Linear mirror, then the cudaArray
    char* lin = static_cast<char*>(t->lin) + (size_t(y)*size_t(t->w) + size_t(x))*texel;
    SFB_CUDA(cudaMemcpy2DAsync(lin, size_t(t->w)*texel, packed, size_t(w)*texel, size_t(w)*texel, size_t(h),
                               cudaMemcpyDeviceToDevice, ctx->stream));
    SFB_CUDA(cudaMemcpy2DToArrayAsync(t->array, size_t(x)*texel, size_t(y), packed, size_t(w)*texel,
                                      size_t(w)*texel, size_t(h), cudaMemcpyDeviceToDevice, ctx->stream));
    if (t->array_stale && w == t->w && h == t->h) t->array_stale = false;
    return SFB_OK;
}

extern "C" int sfb_tex_bind_external(sfb_tex* t, const void* data_dev) {
    SFB_REQUIRE(t, "sfb_tex_bind_external: null texture");
    t->external = data_dev;
    return SFB_OK;
}

extern "C" int sfb_tex_storage(sfb_tex* t, void** data_dev, size_t* bytes) {
    SFB_REQUIRE(t && data_dev, "sfb_tex_storage: null argument");
    t->array_stale = true;
    *data_dev = t->lin;
    if (bytes) *bytes = size_t(t->w)*size_t(t->h)*t->texel_bytes();
    return SFB_OK;
}

extern "C" int sfb_tex_read(sfb_tex* t, void* data_host) {
    SFB_REQUIRE(t && data_host, "sfb_tex_read: null argument");
    const void* src = t->external ? t->external : t->lin;
    SFB_CUDA(cudaMemcpyAsync(data_host, src, size_t(t->w)*size_t(t->h)*t->texel_bytes(), cudaMemcpyDeviceToHost, t->ctx->stream));
    SFB_CUDA(cudaStreamSynchronize(t->ctx->stream));
    return SFB_OK;
}

// ingest.cu — audio file ingestion without ffmpeg.
// The reference decodes every audio file through an ffmpeg child asked for interleaved float32 PCM
// (BrokenAudioReader.stream, ffmpeg.py:1279-1330: `-f f32le`), transposes each chunk and appends it to a host
// ring (audio/module.py:449-453). Here the clip lives in HBM as planar float32, so the file's own sample
// bytes are uploaded as they are on disk (pinned staging → H2D: 16-bit files move half the bytes) and ONE
// HBM-bound kernel does both the sample-format conversion ffmpeg would do and the transpose.
// Conversions are libswresample's: u8 → (x-128)/128, s16 → x/2^15, s24 → x/2^23 (ffmpeg widens it to s32 first),
// s32 → float(x)/2^31 (round-to-nearest int→float, then an exact scaling), f64 → (float)x.
#include "sfb_internal.h"

namespace {

template <int FORMAT> __device__ __forceinline__ float decode(const unsigned char* p) {
    if (FORMAT == SFB_PCM_U8)  return (float(int(p[0]) - 128))*(1.0f/128.0f);
    if (FORMAT == SFB_PCM_S16) return float(short(p[0] | (p[1] << 8)))*(1.0f/32768.0f);
    if (FORMAT == SFB_PCM_S24) {
        const int v = int((unsigned int)p[0] << 8 | (unsigned int)p[1] << 16 | (unsigned int)p[2] << 24);   // s24 → s32
        return __int2float_rn(v)*(1.0f/2147483648.0f);
    }
    if (FORMAT == SFB_PCM_S32) {
        const int v = int((unsigned int)p[0] | (unsigned int)p[1] << 8 | (unsigned int)p[2] << 16 | (unsigned int)p[3] << 24);
        return __int2float_rn(v)*(1.0f/2147483648.0f);
    }
    if (FORMAT == SFB_PCM_F32) {
        return __uint_as_float((unsigned int)p[0] | (unsigned int)p[1] << 8 | (unsigned int)p[2] << 16 | (unsigned int)p[3] << 24);
    }
    unsigned long long bits = 0;
    #pragma unroll
    for (int k = 0; k < 8; k++) bits |= (unsigned long long)p[k] << (8*k);
    return __double2float_rn(__longlong_as_double((long long)bits));
}

constexpr int bytes_of(int format) {
    return format == SFB_PCM_U8 ? 1 : format == SFB_PCM_S16 ? 2 : format == SFB_PCM_S24 ? 3 : format == SFB_PCM_F64 ? 8 : 4;
}

// One thread per sample frame (all channels of one instant): reads are contiguous across the warp, each
// channel's stores are contiguous floats. Grid-stride, sized to the SM count.
template <int FORMAT>
__global__ void __launch_bounds__(256) pcm_ingest_kernel(const unsigned char* __restrict__ raw, long long n_frames, int channels,
                                                         float* __restrict__ planar, long long clip_samples, long long offset) {
    constexpr int B = bytes_of(FORMAT);
    const long long stride = (long long)gridDim.x*blockDim.x;
    for (long long f = (long long)blockIdx.x*blockDim.x + threadIdx.x; f < n_frames; f += stride) {
        const unsigned char* p = raw + f*(long long)(channels*B);
        for (int c = 0; c < channels; c++)
            planar[(long long)c*clip_samples + offset + f] = decode<FORMAT>(p + c*B);
    }
}

}  // namespace

extern "C" int sfb_pcm_ingest(sfb_ctx* ctx, const void* raw_dev, int64_t n_frames, int channels, int format,
                              float* planar_dev, int64_t clip_samples, int64_t offset) {
    SFB_REQUIRE(ctx && raw_dev && planar_dev, "sfb_pcm_ingest: null argument");
    SFB_REQUIRE(channels >= 1 && channels <= 64, "sfb_pcm_ingest: %d channels", channels);
    SFB_REQUIRE(n_frames >= 0 && offset >= 0 && offset + n_frames <= clip_samples,
        "sfb_pcm_ingest: frames [%lld, %lld) outside the clip of %lld samples", (long long)offset, (long long)(offset + n_frames), (long long)clip_samples);
    if (n_frames == 0) return SFB_OK;
    const long long want = (n_frames + 255)/256;
    const int blocks = int(want < (long long)ctx->sm_count*8 ? want : (long long)ctx->sm_count*8);
    const unsigned char* raw = static_cast<const unsigned char*>(raw_dev);
    #define SFB_INGEST(F) case F: pcm_ingest_kernel<F><<<blocks, 256, 0, ctx->stream>>>(raw, n_frames, channels, planar_dev, clip_samples, offset); break;
    switch (format) {
        SFB_INGEST(SFB_PCM_U8) SFB_INGEST(SFB_PCM_S16) SFB_INGEST(SFB_PCM_S24) SFB_INGEST(SFB_PCM_S32)
        SFB_INGEST(SFB_PCM_F32) SFB_INGEST(SFB_PCM_F64)
        default: SFB_FAIL(SFB_EINVAL, "sfb_pcm_ingest: unknown sample format %d", format);
    }
    #undef SFB_INGEST
    SFB_LAUNCH_CHECK(ctx);
    return SFB_OK;
}

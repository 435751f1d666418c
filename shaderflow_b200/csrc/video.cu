// video.cu — the per-frame texture write of ShaderVideo (shaderflow/video.py:57-66: `np.flip(frame, axis=0)`, a C-order
// copy, `texture.write`) on the device: the decoded frame arrives as the file's own bytes (rgb24 / rgba / planar 8-bit
// YUV) and one kernel flips, converts and pads it straight into the texture's storage — no host-side flip or copy.
// YUV → RGB: BT.601 (or full-range JPEG) coefficients in float32, chroma taken from the co-sited / containing sample
// (no chroma interpolation), round half to even. HBM-bound: reads the frame once, writes 4 bytes per texel.
#include "sfb_internal.h"

namespace {

struct VideoParams {
    const unsigned char* src;
    unsigned char* dst;
    int w, h, format, top_down, full_range;
};

__device__ __forceinline__ unsigned char to_byte(float v) { return (unsigned char)__float2int_rn(fminf(fmaxf(v, 0.0f), 255.0f)); }

__global__ void __launch_bounds__(256) video_frame_kernel(const VideoParams P) {
    const int x = blockIdx.x*blockDim.x + threadIdx.x, y = blockIdx.y*blockDim.y + threadIdx.y;       // texture texel, row 0 = bottom
    if (x >= P.w || y >= P.h) return;
    const int sy = P.top_down ? (P.h - 1 - y) : y;
    uchar4 out;
    if (P.format == SFB_VIDEO_RGB24) {
        const unsigned char* p = P.src + (size_t(sy)*size_t(P.w) + size_t(x))*3;
        out = make_uchar4(p[0], p[1], p[2], 255);
    } else if (P.format == SFB_VIDEO_RGBA32) {
        out = reinterpret_cast<const uchar4*>(P.src)[size_t(sy)*size_t(P.w) + size_t(x)];
    } else {
        const int sub_x = (P.format == SFB_VIDEO_YUV444P) ? 0 : 1, sub_y = (P.format == SFB_VIDEO_YUV420P) ? 1 : 0;
        const int cw = (P.w + sub_x) >> sub_x, ch = (P.h + sub_y) >> sub_y;
        const unsigned char* Y = P.src;
        const unsigned char* U = Y + size_t(P.w)*size_t(P.h);
        const unsigned char* V = U + size_t(cw)*size_t(ch);
        const size_t c = size_t(sy >> sub_y)*size_t(cw) + size_t(x >> sub_x);
        const float yv = float(Y[size_t(sy)*size_t(P.w) + size_t(x)]), u = float(U[c]) - 128.0f, v = float(V[c]) - 128.0f;
        float r, g, b;
        if (P.full_range) { r = yv + 1.402f*v; g = yv - 0.344136f*u - 0.714136f*v; b = yv + 1.772f*u; }
        else { const float l = 1.164383f*(yv - 16.0f); r = l + 1.596027f*v; g = l - 0.391762f*u - 0.812968f*v; b = l + 2.017232f*u; }
        out = make_uchar4(to_byte(r), to_byte(g), to_byte(b), 255);
    }
    reinterpret_cast<uchar4*>(P.dst)[size_t(y)*size_t(P.w) + size_t(x)] = out;
}

}  // namespace

extern "C" size_t sfb_video_frame_bytes(int format, int width, int height) {
    const size_t n = size_t(width)*size_t(height);
    switch (format & 0xff) {
        case SFB_VIDEO_RGB24: return n*3;
        case SFB_VIDEO_RGBA32: return n*4;
        case SFB_VIDEO_YUV444P: return n*3;
        case SFB_VIDEO_YUV422P: return n + 2*size_t((width + 1)/2)*size_t(height);
        case SFB_VIDEO_YUV420P: return n + 2*size_t((width + 1)/2)*size_t((height + 1)/2);
    }
    return 0;
}

extern "C" int sfb_video_frame(sfb_ctx* ctx, const void* frame_dev, int format, int width, int height, int top_down, sfb_tex* texture) {
    SFB_REQUIRE(ctx && frame_dev && texture, "sfb_video_frame: null argument");
    SFB_REQUIRE(sfb_video_frame_bytes(format, width, height) > 0, "sfb_video_frame: unknown pixel format %d", format);
    SFB_REQUIRE(texture->w == width && texture->h == height, "sfb_video_frame: frame %dx%d, texture %dx%d", width, height, texture->w, texture->h);
    SFB_REQUIRE(texture->dtype == SFB_DTYPE_U8 && texture->padded == 4 && !texture->external && texture->lin,
        "sfb_video_frame: the texture must be 8-bit with 3 or 4 components and own its storage");
    VideoParams P{static_cast<const unsigned char*>(frame_dev), static_cast<unsigned char*>(texture->lin), width, height,
                  format & 0xff, top_down ? 1 : 0, (format & SFB_VIDEO_FULL_RANGE) ? 1 : 0};
    dim3 block(32, 8), grid((width + 31)/32, (height + 7)/8);
    video_frame_kernel<<<grid, block, 0, ctx->stream>>>(P);
    SFB_LAUNCH_CHECK(ctx);
    texture->array_stale = true;                      // sampled through the linear mirror until the next sfb_tex_write
    return SFB_OK;
}

// sampler.cuh — texture() for the transliterated scenes.
// Two implementations of the OpenGL 3.3 sampling rules the reference gets from the GL driver
// (texture.py:104-137,281-282 pick filter / wrap; oracle/glsl_np.py `Texture.sample` states the rules):
//   exact:    point-fetch texels from the pitch-linear mirror, float32 bilinear weights  (parity default)
//   hardware: cudaTextureObject_t on the cudaArray (9-bit weights, as GL hardware filters) (opt-in)
#pragma once
#include "glsl.cuh"
#include "sfb_internal.h"
#include <cuda_fp16.h>

namespace glsl {

SFB_DEV int wrap_index(int i, int n, int repeat) {
    if (repeat) { i %= n; return (i < 0) ? i + n : i; }
    return min(max(i, 0), n - 1);
}

// One texel as GL would return it: unorm8 → c/255, missing components → (0, 0, 1)
SFB_DEV vec4 texel_fetch(const DevSampler& s, int ix, int iy) {
    ix = wrap_index(ix, s.w, s.rx);
    iy = wrap_index(iy, s.h, s.ry);
    size_t idx = size_t(iy)*size_t(s.w) + size_t(ix);
    vec4 t = mk4(0.0f, 0.0f, 0.0f, 1.0f);
    if (s.dtype == SFB_DTYPE_U8) {
        if (s.padded == 4) {
            uchar4 c = __ldg(reinterpret_cast<const uchar4*>(s.lin) + idx);
            t = mk4(c.x/255.0f, c.y/255.0f, c.z/255.0f, c.w/255.0f);
        } else if (s.padded == 2) {
            uchar2 c = __ldg(reinterpret_cast<const uchar2*>(s.lin) + idx);
            t.x = c.x/255.0f; t.y = c.y/255.0f;
        } else {
            t.x = __ldg(reinterpret_cast<const unsigned char*>(s.lin) + idx)/255.0f;
        }
    } else if (s.dtype == SFB_DTYPE_F32) {
        if (s.padded == 4) {
            float4 c = __ldg(reinterpret_cast<const float4*>(s.lin) + idx);
            t = mk4(c.x, c.y, c.z, c.w);
        } else if (s.padded == 2) {
            float2 c = __ldg(reinterpret_cast<const float2*>(s.lin) + idx);
            t.x = c.x; t.y = c.y;
        } else {
            t.x = __ldg(reinterpret_cast<const float*>(s.lin) + idx);
        }
    } else {
        const __half* p = reinterpret_cast<const __half*>(s.lin) + idx*size_t(s.padded);
        t.x = __half2float(p[0]);
        if (s.padded >= 2) t.y = __half2float(p[1]);
        if (s.padded == 4) { t.z = __half2float(p[2]); t.w = __half2float(p[3]); }
    }
    if (s.comps == 3) t.w = 1.0f;
    return t;
}

template <bool HW>
SFB_DEV vec4 texture(const DevSampler& s, vec2 uv) {
    if (HW && s.hw != 0ull) {
        float4 c = tex2D<float4>((cudaTextureObject_t)s.hw, uv.x, uv.y);
        vec4 t = mk4(c.x, c.y, c.z, c.w);
        if (s.comps < 4) t.w = 1.0f;         // GL: missing alpha reads 1 (CUDA returns 0)
        return t;
    }
    float u = uv.x*float(s.w), v = uv.y*float(s.h);
    if (s.filter == SFB_FILTER_NEAREST)
        return texel_fetch(s, int(floorf(u)), int(floorf(v)));
    float ub = u - 0.5f, vb = v - 0.5f;
    float fx = floorf(ub), fy = floorf(vb);
    float a = ub - fx, b = vb - fy;
    int i0 = int(fx), j0 = int(fy);
    vec4 t00 = texel_fetch(s, i0, j0),     t10 = texel_fetch(s, i0 + 1, j0);
    vec4 t01 = texel_fetch(s, i0, j0 + 1), t11 = texel_fetch(s, i0 + 1, j0 + 1);
    vec4 top = t00*(1.0f - a) + t10*a;
    vec4 bot = t01*(1.0f - a) + t11*a;
    return top*(1.0f - b) + bot*b;
}

// shaderflow.glsl:165-169, 202-204
template <bool HW>
SFB_DEV vec4 gtexture(const DevSampler& s, vec2 gluv) {
    vec2 scale = mk2(float(s.h)/float(s.w), 1.0f);
    return texture<HW>(s, gluv2stuv(gluv*scale));
}
template <bool HW>
SFB_DEV vec4 stexture(const DevSampler& s, vec2 stuv) { return gtexture<HW>(s, stuv2gluv(stuv)); }

} // namespace glsl

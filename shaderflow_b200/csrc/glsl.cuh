// glsl.cuh — the GLSL vocabulary the transliterated scenes are written in: vec2/3/4 value types, the
// builtins they use, and ShaderFlow's std-lib (reference: shaderflow/resources/shaders/include/
// shaderflow.glsl; line numbers of that file are cited per helper). float32, accurate
// libdevice math (no --use_fast_math): parity against oracle/glsl_np.py is the first gate.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define SFB_DEV __device__ __forceinline__
// pure arithmetic helpers are also callable on the host (launch planning evaluates the same formulas)
#define SFB_HD __host__ __device__ __forceinline__

namespace glsl {

struct vec2 { float x, y; };
struct vec3 { float x, y, z; };
struct vec4 { float x, y, z, w; };

SFB_HD vec2 mk2(float x, float y) { return vec2{x, y}; }
SFB_HD vec2 mk2(float s) { return vec2{s, s}; }
SFB_HD vec3 mk3(float x, float y, float z) { return vec3{x, y, z}; }
SFB_HD vec3 mk3(float s) { return vec3{s, s, s}; }
SFB_HD vec4 mk4(float x, float y, float z, float w) { return vec4{x, y, z, w}; }
SFB_HD vec4 mk4(vec3 v, float w) { return vec4{v.x, v.y, v.z, w}; }
SFB_HD vec4 mk4(float s) { return vec4{s, s, s, s}; }
SFB_HD vec3 xyz(vec4 v) { return vec3{v.x, v.y, v.z}; }
SFB_HD vec2 xy(vec3 v) { return vec2{v.x, v.y}; }
SFB_HD vec2 yx(vec2 v) { return vec2{v.y, v.x}; }

#define SFB_VEC_OPS(T, ...) \
    SFB_HD T operator+(T a, T b) { return T{__VA_ARGS__(+)}; } \
    SFB_HD T operator-(T a, T b) { return T{__VA_ARGS__(-)}; } \
    SFB_HD T operator*(T a, T b) { return T{__VA_ARGS__(*)}; } \
    SFB_HD T operator/(T a, T b) { return T{__VA_ARGS__(/)}; }
#define SFB_C2(op) a.x op b.x, a.y op b.y
#define SFB_C3(op) a.x op b.x, a.y op b.y, a.z op b.z
#define SFB_C4(op) a.x op b.x, a.y op b.y, a.z op b.z, a.w op b.w
SFB_VEC_OPS(vec2, SFB_C2)
SFB_VEC_OPS(vec3, SFB_C3)
SFB_VEC_OPS(vec4, SFB_C4)

SFB_HD vec2 operator*(vec2 a, float s) { return vec2{a.x*s, a.y*s}; }
SFB_HD vec2 operator*(float s, vec2 a) { return vec2{s*a.x, s*a.y}; }
SFB_HD vec2 operator/(vec2 a, float s) { return vec2{a.x/s, a.y/s}; }
SFB_HD vec2 operator+(vec2 a, float s) { return vec2{a.x + s, a.y + s}; }
SFB_HD vec2 operator-(vec2 a, float s) { return vec2{a.x - s, a.y - s}; }
SFB_HD vec2 operator-(float s, vec2 a) { return vec2{s - a.x, s - a.y}; }
SFB_HD vec3 operator*(vec3 a, float s) { return vec3{a.x*s, a.y*s, a.z*s}; }
SFB_HD vec3 operator*(float s, vec3 a) { return vec3{s*a.x, s*a.y, s*a.z}; }
SFB_HD vec3 operator/(vec3 a, float s) { return vec3{a.x/s, a.y/s, a.z/s}; }
SFB_HD vec3 operator+(vec3 a, float s) { return vec3{a.x + s, a.y + s, a.z + s}; }
SFB_HD vec3 operator+(float s, vec3 a) { return vec3{s + a.x, s + a.y, s + a.z}; }
SFB_HD vec3 operator-(vec3 a, float s) { return vec3{a.x - s, a.y - s, a.z - s}; }
SFB_HD vec3 operator-(float s, vec3 a) { return vec3{s - a.x, s - a.y, s - a.z}; }
SFB_HD vec4 operator*(vec4 a, float s) { return vec4{a.x*s, a.y*s, a.z*s, a.w*s}; }
SFB_HD vec4 operator*(float s, vec4 a) { return vec4{s*a.x, s*a.y, s*a.z, s*a.w}; }
SFB_HD vec4 operator/(vec4 a, float s) { return vec4{a.x/s, a.y/s, a.z/s, a.w/s}; }

// Builtins. clamp/min/max drop NaNs like fminf/fmaxf (the oracle fixes the same rule).
SFB_HD float clamp(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
SFB_HD float mix(float a, float b, float t) { return a*(1.0f - t) + b*t; }
SFB_HD vec3 mix(vec3 a, vec3 b, float t) { return a*(1.0f - t) + b*t; }
SFB_HD float smoothstep(float e0, float e1, float x) {
    float t = clamp((x - e0)/(e1 - e0), 0.0f, 1.0f);
    return t*t*(3.0f - 2.0f*t);
}
SFB_HD float mod(float x, float y) { return x - y*floorf(x/y); }
SFB_HD float dot(vec2 a, vec2 b) { return a.x*b.x + a.y*b.y; }
SFB_HD float dot(vec3 a, vec3 b) { return a.x*b.x + a.y*b.y + a.z*b.z; }
SFB_HD float length(vec2 v) { return sqrtf(dot(v, v)); }
SFB_HD float length(vec3 v) { return sqrtf(dot(v, v)); }
SFB_HD vec3 normalize(vec3 v) { return v/length(v); }
SFB_HD vec3 cross(vec3 a, vec3 b) { return vec3{a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x}; }
SFB_HD vec3 abs3(vec3 v) { return vec3{fabsf(v.x), fabsf(v.y), fabsf(v.z)}; }
SFB_HD vec3 max3(vec3 v, float s) { return vec3{fmaxf(v.x, s), fmaxf(v.y, s), fmaxf(v.z, s)}; }
SFB_HD float sign(float x) { return (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f); }

// shaderflow.glsl constants (:7-11)
constexpr float PI    = 3.1415926535897932f;
constexpr float TAU   = 6.2831853071795864f;
constexpr float SQRT3 = 1.7320508075688772f;

SFB_HD vec2 stuv2gluv(vec2 s) { return (s*2.0f) - 1.0f; }                        // :96
SFB_HD vec2 gluv2stuv(vec2 g) { return (g + 1.0f)/2.0f; }                        // :100
SFB_HD vec2 zoom(vec2 uv, float z, vec2 anchor) { return (uv - anchor)*(z*z) + anchor; }  // :361-363

// mat2(c,-s,s,c)*v with GLSL's column-major constructor (:75-77): columns (c,-s), (s,c)
SFB_HD vec2 rotate2d_mul(float angle, vec2 v) {
    float c = cosf(angle), s = sinf(angle);
    return vec2{c*v.x + s*v.y, (-s)*v.x + c*v.y};
}
// rotate3d (:82-84)
SFB_HD vec3 rotate3d(vec3 v, vec3 axis, float angle) {
    return mix(dot(axis, v)*axis, v, cosf(angle)) + cross(axis, v)*sinf(angle);
}
SFB_HD float atan1n(vec2 p) { return atan2f(p.y, p.x)/PI; }                      // :378-380
SFB_HD float atan2_pos(float y, float x) {                                        // :382-388
    return (y < 0.0f) ? (TAU - atan2f(-y, x)) : atan2f(y, x);
}
SFB_HD float atan2n(float y, float x) { return atan2_pos(y, x)/TAU; }             // :394-396

SFB_HD vec3 palette(float t, vec3 A, vec3 B, vec3 C, vec3 D) {                    // :212-220
    if (t < 0.25f) return mix(A, B, t*4.0f);
    else if (t < 0.5f) return mix(B, C, (t - 0.25f)*4.0f);
    return mix(C, D, (t - 0.5f)*4.0f);
}
SFB_HD vec3 palette_magma(float t) {                                              // :222-226
    return palette(t, mk3(0.01060815f, 0.01808215f, 0.10018654f), mk3(0.38092887f, 0.12061482f, 0.32506528f),
                      mk3(0.79650140f, 0.10506637f, 0.31063031f), mk3(0.95922872f, 0.53307513f, 0.37488950f));
}

SFB_HD vec3 hsv2rgb(float h, float s, float v) {                                  // :406-425
    h = mod(h, TAU);
    float c = v*s;
    float x = c*(1.0f - fabsf(mod(h/(PI/3.0f), 2.0f) - 1.0f));
    float m = v - c;
    float sector = floorf(6.0f*(h/(2.0f*PI)));
    vec3 rgb = mk3(0.0f);                       // `default:` — also where a NaN hue lands
    if      (sector == 0.0f) rgb = mk3(c, x, 0.0f);
    else if (sector == 1.0f) rgb = mk3(x, c, 0.0f);
    else if (sector == 2.0f) rgb = mk3(0.0f, c, x);
    else if (sector == 3.0f) rgb = mk3(0.0f, x, c);
    else if (sector == 4.0f) rgb = mk3(x, 0.0f, c);
    else if (sector == 5.0f) rgb = mk3(c, 0.0f, x);
    return rgb + m;
}

SFB_HD float sdBox(vec3 origin, vec3 point, vec3 size) {                          // :285-288
    vec3 d = abs3(origin - point) - size/2.0f;
    return fminf(fmaxf(d.x, fmaxf(d.y, d.z)), 0.0f) + length(max3(d, 0.0f));
}

} // namespace glsl

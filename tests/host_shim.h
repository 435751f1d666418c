// Host stand-ins for the few CUDA built-ins csrc/jit/glsl_rt.cuh and shaderflow_rt.cuh use, so that the code the GLSL ->
// CUDA translator emits can be compiled with g++ and executed on the CPU (tests/test_glsl_host.py): the same templates,
// the same generated `Shader`, float32 with -ffp-contract=off — one rounding per operation, like the NVRTC build.
// Test infrastructure only. Screen-space derivatives need neighbouring lanes and read 0 here.
#pragma once
#define SFB_HOST_SHIM 1
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define __device__
#define __forceinline__ inline
struct sfb_host_dim3 { unsigned x = 0, y = 0, z = 0; };
static sfb_host_dim3 threadIdx, blockIdx;
struct uchar2 { unsigned char x, y; };
struct uchar4 { unsigned char x, y, z, w; };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
template <class T> inline T __ldg(const T* p) { return *p; }
inline unsigned __activemask() { return 1u; }
inline float __shfl_xor_sync(unsigned, float v, int) { return v; }
inline float sfb_host_half_to_float(unsigned short h) {
    const unsigned sign = (h >> 15) & 1u, exponent = (h >> 10) & 31u, mantissa = h & 1023u;
    float value;
    if (exponent == 0) value = std::ldexp(float(mantissa), -24);
    else if (exponent == 31) value = mantissa ? NAN : INFINITY;
    else value = std::ldexp(float(mantissa | 1024u), int(exponent) - 25);
    return sign ? -value : value;
}

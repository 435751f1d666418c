// Host stand-ins for the few CUDA built-ins the ahead-of-time scene functions use (csrc/glsl.cuh, sampler.cuh, scenes.cuh),
// so that the PRODUCTION per-fragment code — the functions screen_kernel / frame_kernel call — can be compiled with g++
// and run on the CPU against the goldens (tests/test_aot_host.py) and timed as a compiled CPU baseline (oracle/cpu_compiled.py). Test infrastructure only.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
template <class T> inline T __ldg(const T* p) { return *p; }
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
    const unsigned long long v = (unsigned long long)x | ((unsigned long long)y << 32);
    unsigned r = 0;
    for (int i = 0; i < 4; i++) r |= (unsigned)((v >> (8*((s >> (4*i)) & 7u))) & 0xFFu) << (8*i);
    return r;
}
inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
template <class T> inline T tex2D(unsigned long long, float, float) { return T{}; }      // the hardware-filter path is not taken (hw = 0)

// Microbenchmark: issue rate of FFMA vs fma.rn.f32x2 (FFMA2) vs FADD on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
__global__ void k_ffma(float* out, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0+4, x5=x0+5, x6=x0+6, x7=x0+7;
    for (int i = 0; i < ITERS; i++) {
        x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
        x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
    out[blockIdx.x*blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__global__ void k_ffma2(float* out, float a, float b) {
    unsigned long long x0, x1, x2, x3, x4, x5, x6, x7, A, B;
    float f = threadIdx.x;
    asm("mov.b64 %0, {%1, %2};" : "=l"(x0) : "f"(f), "f"(f + 1)); x1 = x0; x2 = x0; x3 = x0; x4=x0;x5=x0;x6=x0;x7=x0;
    asm("mov.b64 %0, {%1, %1};" : "=l"(A) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(B) : "f"(b));
    for (int i = 0; i < ITERS; i++) {
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x0) : "l"(A), "l"(B));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x1) : "l"(A), "l"(B));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x2) : "l"(A), "l"(B));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x3) : "l"(A), "l"(B));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x4) : "l"(A), "l"(B));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x5) : "l"(A), "l"(B));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x6) : "l"(A), "l"(B));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x7) : "l"(A), "l"(B));
    }
    float lo, hi; unsigned long long s = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(s));
    out[blockIdx.x*blockDim.x + threadIdx.x] = lo + hi;
}
__global__ void k_fadd(float* out, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0+4, x5=x0+5, x6=x0+6, x7=x0+7;
    for (int i = 0; i < ITERS; i++) {
        x0 = __fadd_rn(x0, a); x1 = __fadd_rn(x1, a); x2 = __fadd_rn(x2, a); x3 = __fadd_rn(x3, a);
        x4 = __fadd_rn(x4, a); x5 = __fadd_rn(x5, a); x6 = __fadd_rn(x6, a); x7 = __fadd_rn(x7, a);
    }
    out[blockIdx.x*blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__global__ void k_mix(float* out, float a, float b) {   // FFMA + integer LOP3 interleaved
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3; unsigned y0 = threadIdx.x, y1 = y0*3, y2 = y0*5, y3 = y0*7;
    for (int i = 0; i < ITERS; i++) {
        x0 = fmaf(x0, a, b); y0 = (y0 ^ 0x9e3779b9u) + (y0 << 3); x1 = fmaf(x1, a, b); y1 = (y1 ^ 0x9e3779b9u) + (y1 << 3);
        x2 = fmaf(x2, a, b); y2 = (y2 ^ 0x9e3779b9u) + (y2 << 3); x3 = fmaf(x3, a, b); y3 = (y3 ^ 0x9e3779b9u) + (y3 << 3);
    }
    out[blockIdx.x*blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + float(y0 ^ y1 ^ y2 ^ y3);
}
template <typename K> float run(K k, float* out) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<<<148*8, 256>>>(out, 1.0001f, 0.5f); cudaDeviceSynchronize();
    cudaEventRecord(a); k<<<148*8, 256>>>(out, 1.0001f, 0.5f); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
    float* out; cudaMalloc(&out, 148*8*256*4);
    double n = 148.0*8*256*ITERS*8;
    float t1 = run(k_ffma, out), t2 = run(k_ffma2, out), t3 = run(k_fadd, out), t4 = run(k_mix, out);
    printf("FFMA : %.3f ms  %.1f Tinstr-lanes/s (%.1f TFLOP/s)\n", t1, n/t1/1e9, 2*n/t1/1e9);
    printf("FFMA2: %.3f ms  %.1f Tinstr-lanes/s (%.1f TFLOP/s)\n", t2, n/t2/1e9, 4*n/t2/1e9);
    printf("FADD : %.3f ms  %.1f Tinstr-lanes/s\n", t3, n/t3/1e9);
    printf("MIX  : %.3f ms  (4 FFMA + 4x3 int per iter)\n", t4);
    return 0;
}

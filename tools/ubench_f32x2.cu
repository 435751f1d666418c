// Microbenchmark: issue rate of FFMA vs fma.rn.f32x2 (FFMA2, vector and scalar-broadcast operands) on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
__global__ void k_ffma(float* out, float a, float b) {
    float x[16];
    for (int k = 0; k < 16; k++) x[k] = threadIdx.x + k;
    for (int i = 0; i < ITERS; i++) {
        #pragma unroll
        for (int k = 0; k < 16; k++) x[k] = fmaf(x[k], a, b);
    }
    float s = 0; for (int k = 0; k < 16; k++) s += x[k];
    out[blockIdx.x*blockDim.x + threadIdx.x] = s;
}
__device__ __forceinline__ unsigned long long pack(float lo, float hi) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float2 unpack(unsigned long long v) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
// 8 independent pair chains = the same 16 float chains as k_ffma
__global__ void k_ffma2(float* out, float a, float a2, float b, float b2) {
    unsigned long long x[8];
    for (int k = 0; k < 8; k++) x[k] = pack(threadIdx.x + 2*k, threadIdx.x + 2*k + 1);
    const unsigned long long A = pack(a, a2), B = pack(b, b2);
    for (int i = 0; i < ITERS; i++) {
        #pragma unroll
        for (int k = 0; k < 8; k++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[k]) : "l"(A), "l"(B));
    }
    float s = 0; for (int k = 0; k < 8; k++) { float2 v = unpack(x[k]); s += v.x + v.y; }
    out[blockIdx.x*blockDim.x + threadIdx.x] = s;
}
// scalar-broadcast multiplier (both halves from one register)
__global__ void k_ffma2s(float* out, float a, float b, float b2) {
    unsigned long long x[8];
    for (int k = 0; k < 8; k++) x[k] = pack(threadIdx.x + 2*k, threadIdx.x + 2*k + 1);
    const unsigned long long A = pack(a, a), B = pack(b, b2);
    for (int i = 0; i < ITERS; i++) {
        #pragma unroll
        for (int k = 0; k < 8; k++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[k]) : "l"(A), "l"(B));
    }
    float s = 0; for (int k = 0; k < 8; k++) { float2 v = unpack(x[k]); s += v.x + v.y; }
    out[blockIdx.x*blockDim.x + threadIdx.x] = s;
}
// 8 FFMA2 + 8 broadcast LDS.128 per iteration (the visualizer's table loads riding along)
__global__ void k_ffma2_lds(float* out, float a, float a2, float b, float b2) {
    __shared__ float4 tab[256];
    tab[threadIdx.x] = make_float4(a, a2, b, b2);
    __syncthreads();
    unsigned long long x[8];
    for (int k = 0; k < 8; k++) x[k] = pack(threadIdx.x + 2*k, threadIdx.x + 2*k + 1);
    for (int i = 0; i < ITERS; i++) {
        #pragma unroll
        for (int k = 0; k < 8; k++) {
            const float4 t = tab[(i*8 + k) & 255];
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[k]) : "l"(pack(t.x, t.y)), "l"(pack(t.z, t.w)));
        }
    }
    float s = 0; for (int k = 0; k < 8; k++) { float2 v = unpack(x[k]); s += v.x + v.y; }
    out[blockIdx.x*blockDim.x + threadIdx.x] = s;
}
template <typename F> float run(F launch) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
    float* out; cudaMalloc(&out, 148*8*256*4);
    const double fma = 148.0*8*256*ITERS*16;      // scalar FMAs per launch, every kernel
    float t1 = run([&]{ k_ffma<<<148*8, 256>>>(out, 1.0001f, 0.5f); });
    float t2 = run([&]{ k_ffma2<<<148*8, 256>>>(out, 1.0001f, 1.0002f, 0.5f, 0.25f); });
    float t3 = run([&]{ k_ffma2s<<<148*8, 256>>>(out, 1.0001f, 0.5f, 0.25f); });
    float t4 = run([&]{ k_ffma2_lds<<<148*8, 256>>>(out, 1.0001f, 1.0002f, 0.5f, 0.25f); });
    printf("FFMA      : %.3f ms  %.1f TFLOP/s\n", t1, 2*fma/t1/1e9);
    printf("FFMA2     : %.3f ms  %.1f TFLOP/s\n", t2, 2*fma/t2/1e9);
    printf("FFMA2 bcst: %.3f ms  %.1f TFLOP/s\n", t3, 2*fma/t3/1e9);
    printf("FFMA2+LDS : %.3f ms  %.1f TFLOP/s (8 FFMA2 + 8 broadcast LDS.128 per iteration)\n", t4, 2*fma/t4/1e9);
    return 0;
}

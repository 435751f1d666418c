// Isolates the TMA 2D tile load: descriptor as struct member param / top-level param / global memory
#include <cstdio>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#define BW 64
#define BH 32
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void load_box(void* dst, const void* tmap, unsigned long long* bar, int x, int y) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(bar)));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(BW*BH*4) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :: "r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void wait(unsigned long long* bar) {
    unsigned done; do { asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)) : "memory"); } while (!done);
}
struct Wrapped { CUtensorMap tmap; int x, y; unsigned* out; };
__global__ void k_member(const __grid_constant__ Wrapped W) {
    __shared__ __align__(128) unsigned box[BW*BH]; __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) load_box(box, &W.tmap, &bar, W.x, W.y);
    __syncthreads(); wait(&bar);
    for (int i = threadIdx.x; i < BW*BH; i += blockDim.x) W.out[i] = box[i];
}
__global__ void k_top(const __grid_constant__ CUtensorMap tmap, int x, int y, unsigned* out) {
    __shared__ __align__(128) unsigned box[BW*BH]; __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) load_box(box, &tmap, &bar, x, y);
    __syncthreads(); wait(&bar);
    for (int i = threadIdx.x; i < BW*BH; i += blockDim.x) out[i] = box[i];
}
__global__ void k_global(const CUtensorMap* tmap, int x, int y, unsigned* out) {
    __shared__ __align__(128) unsigned box[BW*BH]; __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) load_box(box, tmap, &bar, x, y);
    __syncthreads(); wait(&bar);
    for (int i = threadIdx.x; i < BW*BH; i += blockDim.x) out[i] = box[i];
}
// D: the library kernel's shape — 32 KB float4 window, other shared arrays after it, barrier last, init before a
// __syncthreads, TMA inside nested ifs keyed on shared flags, coordinates derived from floats
__global__ void __launch_bounds__(256) k_shape(const CUtensorMap* tmap, float fx, float fy, unsigned* out, int mode) {
    __shared__ __align__(128) float4 window[32*64];
    __shared__ unsigned int stage[8][32];
    __shared__ float red[4][8];
    __shared__ int win[4];
    __shared__ __align__(8) unsigned long long bar;
    const int tid = threadIdx.y*32 + threadIdx.x;
    float lo = fx + threadIdx.x*0.01f;
    for (int o = 16; o > 0; o >>= 1) lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    if (threadIdx.x == 0) red[0][threadIdx.y] = lo;
    if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
    __syncthreads();
    if (tid == 0) {
        float ax = red[0][0]; for (int r = 1; r < 8; r++) ax = fminf(ax, red[0][r]);
        const int x0 = int(floorf(ax)) - 1, y0 = int(floorf(fy)) - 1;
        win[0] = x0; win[1] = y0; win[2] = 1; win[3] = (mode && x0 >= 0 && y0 >= 0) ? 1 : 0;
        if (win[3]) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar)), "r"(BW*BH*4) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                :: "r"(smem_u32(window)), "l"(tmap), "r"(smem_u32(&bar)), "r"(x0), "r"(y0) : "memory");
        }
    }
    __syncthreads();
    if (win[3]) {
        wait(&bar);
        const unsigned* raw = reinterpret_cast<const unsigned*>(window);
        for (int i = tid; i < BW*BH; i += 256) out[i] = raw[i];
    }
    stage[threadIdx.y][threadIdx.x] = tid;
}
int check(const char* name, unsigned* out_d, int W, int x, int y) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return 1; }
    static unsigned h[BW*BH]; cudaMemcpy(h, out_d, sizeof(h), cudaMemcpyDeviceToHost);
    int bad = 0; for (int j = 0; j < BH; j++) for (int i = 0; i < BW; i++) if (h[j*BW + i] != unsigned((y + j)*W + x + i)) bad++;
    printf("%s: %d mismatches\n", name, bad); return bad;
}
int main() {
    const int W = 1920, H = 1080;
    unsigned* img; cudaMalloc(&img, W*H*4);
    unsigned* host = new unsigned[W*H]; for (int i = 0; i < W*H; i++) host[i] = i;
    cudaMemcpy(img, host, W*H*4, cudaMemcpyHostToDevice);
    unsigned* out; cudaMalloc(&out, BW*BH*4);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    CUtensorMap tmap; cuuint64_t dims[2] = {W, H}, strides[1] = {W*4}; cuuint32_t box[2] = {BW, BH}, es[2] = {1, 1};
    CUresult rc = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, img, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d query=%d img=%p:", int(rc), int(q), (void*)img);
    for (int i = 0; i < 16; i++) printf(" %016llx", ((unsigned long long*)&tmap)[i]); printf("\n");
    int x = 300, y = 200;
    k_top<<<1, 256>>>(tmap, x, y, out); check("top-level param", out, W, x, y);
    Wrapped w; w.tmap = tmap; w.x = x; w.y = y; w.out = out;
    k_member<<<1, 256>>>(w); check("struct member param", out, W, x, y);
    CUtensorMap* td; cudaMalloc(&td, sizeof(tmap)); cudaMemcpy(td, &tmap, sizeof(tmap), cudaMemcpyHostToDevice);
    k_global<<<1, 256>>>(td, x, y, out); check("global memory", out, W, x, y);
    dim3 blk(32, 8);
    k_shape<<<1, blk>>>(td, 301.5f, 201.5f, out, 1); check("library-shaped kernel", out, W, 300, 200);
    return 0;
}

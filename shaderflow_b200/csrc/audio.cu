// audio.cu — kernels K1 (STFT → filterbank spectrogram) and K2 (dynamics scan, volume/std, waveform)
// and their C ABI. Reference call sites replaced: audio/spectrogram.py:155-176,298-311 (numpy.fft.rfft,
// scipy.sparse csr dot, DynamicNumber.next), audio/module.py:137-138,447-458, audio/waveform.py:64-87.
#include "sfb_internal.h"

#include <cmath>
#include <vector>

// =================================================================================================
// K1: one CTA per frame (persistent over frames), N/16 threads.
//   1. coalesced float4 loads of both channels' window [tell-N-1, tell-1) from the planar clip,
//      window multiply, packed as z = L + iR into shared memory                  (module.py:137-138)
//   2. in-place Stockham FFT, radix-16 passes (+ one radix-2/4/8 pass), twiddles from a smem table
//   3. split the two real spectra out of Z, magnitude                            (spectrogram.py:20-26,170)
//   4. CSR filterbank row per thread, optional volume epilogue, float2 store     (spectrogram.py:176)
// =================================================================================================

struct cpx { float x, y; };
__device__ __forceinline__ cpx operator+(cpx a, cpx b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ cpx operator-(cpx a, cpx b) { return {a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ cpx cmul(cpx a, cpx b) { return {a.x*b.x - a.y*b.y, a.x*b.y + a.y*b.x}; }
__device__ __forceinline__ cpx mul_mi(cpx a) { return {a.y, -a.x}; }     // a * (-i)

// DFT4 (forward, e^{-2πi nk/4}) on 4 values, natural order out
__device__ __forceinline__ void dft4(cpx& a, cpx& b, cpx& c, cpx& d) {
    cpx s0 = a + c, s1 = a - c, s2 = b + d, s3 = mul_mi(b - d);
    a = s0 + s2; b = s1 + s3; c = s0 - s2; d = s1 - s3;
}
__device__ __forceinline__ void dft2(cpx& a, cpx& b) { cpx t = a; a = t + b; b = t - b; }

#define SFB_C8  0.70710678118654752440f
#define SFB_C16 0.92387953251128675613f
#define SFB_S16 0.38268343236508977173f

// DFT8: n = b + 2a, k = c + 4d: two DFT4 over a, twiddle W8^{bc}, DFT2 over b
__device__ __forceinline__ void dft8(cpx* v) {
    dft4(v[0], v[2], v[4], v[6]);
    dft4(v[1], v[3], v[5], v[7]);
    const cpx w1 = {SFB_C8, -SFB_C8}, w3 = {-SFB_C8, -SFB_C8};
    cpx o0 = v[1], o1 = cmul(v[3], w1), o2 = mul_mi(v[5]), o3 = cmul(v[7], w3);
    cpx e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
    v[0] = e0 + o0; v[4] = e0 - o0;
    v[1] = e1 + o1; v[5] = e1 - o1;
    v[2] = e2 + o2; v[6] = e2 - o2;
    v[3] = e3 + o3; v[7] = e3 - o3;
}

// DFT16: n = b + 4a, k = c + 4d: X[c+4d] = Σ_b W4^{bd} W16^{bc} Σ_a x[b+4a] W4^{ac}
__device__ __forceinline__ void dft16(cpx* v) {
    #pragma unroll
    for (int b = 0; b < 4; b++) dft4(v[b], v[b + 4], v[b + 8], v[b + 12]);   // y_b[c] lands in v[b + 4c]
    // W16^{bc}
    const cpx W1 = {SFB_C16, -SFB_S16}, W2 = {SFB_C8, -SFB_C8}, W3 = {SFB_S16, -SFB_C16};
    const cpx W6 = {-SFB_C8, -SFB_C8}, W9 = {-SFB_C16, SFB_S16};
    v[1 + 4]  = cmul(v[1 + 4], W1);  v[2 + 4]  = cmul(v[2 + 4], W2);  v[3 + 4]  = cmul(v[3 + 4], W3);
    v[1 + 8]  = cmul(v[1 + 8], W2);  v[2 + 8]  = mul_mi(v[2 + 8]);    v[3 + 8]  = cmul(v[3 + 8], W6);
    v[1 + 12] = cmul(v[1 + 12], W3); v[2 + 12] = cmul(v[2 + 12], W6); v[3 + 12] = cmul(v[3 + 12], W9);
    #pragma unroll
    for (int c = 0; c < 4; c++) dft4(v[4*c], v[4*c + 1], v[4*c + 2], v[4*c + 3]);  // over b → d: X[c+4d] in v[4c+d]
    // reorder v[4c + d] → natural index c + 4d
    cpx t[16];
    #pragma unroll
    for (int c = 0; c < 4; c++)
        #pragma unroll
        for (int d = 0; d < 4; d++) t[c + 4*d] = v[4*c + d];
    #pragma unroll
    for (int i = 0; i < 16; i++) v[i] = t[i];
}

template <int R> __device__ __forceinline__ void dftR(cpx* v) {
    if constexpr (R == 16) dft16(v);
    else if constexpr (R == 8) dft8(v);
    else if constexpr (R == 4) dft4(v[0], v[1], v[2], v[3]);
    else if constexpr (R == 2) dft2(v[0], v[1]);
}

// Padded index into the data buffer: one extra slot every 16 keeps the radix-16 scatter conflict-free
__device__ __forceinline__ int pad(int i) { return i + (i >> 4); }

// Twiddle tables. For the pass with sub-transform size Ns and radix R a butterfly at k = j mod Ns needs
// w^r, w = exp(-2*pi*i*k/(Ns*R)), r = 1..R-1. Only the power-of-two members w^1, w^2, w^4, w^8 are tabled —
// laid out [m][k] so a warp reads consecutive entries (no bank conflicts) — and the rest are products of
// at most three table values (<= 3-4 ulp). Pass 1 (Ns = 1) needs none.
template <int LOG2N> struct TwiddleLayout {
    static constexpr int N = 1 << LOG2N, FULL = LOG2N/4, REM = LOG2N%4;
    // offset of pass p (p >= 1) in float2 entries; pass p has Ns = 16^p and log2(R) rows
    static constexpr int rows(int p) { return (p < FULL) ? 4 : REM; }
    static constexpr int ns(int p) { int v = 1; for (int i = 0; i < p; i++) v *= 16; return v; }
    static constexpr int offset(int p) { int o = 0; for (int q = 1; q < p; q++) o += rows(q)*ns(q); return o; }
    static constexpr int passes = FULL + (REM ? 1 : 0);
    static constexpr int total = offset(passes);
};

// One in-place Stockham pass of radix R over N points held in `buf`; each thread owns 16/R butterflies.
// Ns = product of the radices of the previous passes; `tw` = this pass's table ([log2 R][Ns]).
template <int LOG2N, int R>
__device__ __forceinline__ void stockham_pass(cpx* buf, const float2* tw, int Ns, int tid) {
    constexpr int N = 1 << LOG2N, T = N/16, M = 16/R;
    cpx v[M][R];
    if (tid < T) {
        #pragma unroll
        for (int m = 0; m < M; m++) {
            const int j = tid + T*m;                       // butterfly index in [0, N/R)
            #pragma unroll
            for (int r = 0; r < R; r++) v[m][r] = buf[pad(j + r*(N/R))];
            if (Ns > 1) {
                const int k = j & (Ns - 1);
                cpx w[R];
                #pragma unroll
                for (int b = 0, r = 1; r < R; r <<= 1, b++) { const float2 t = tw[b*Ns + k]; w[r] = cpx{t.x, t.y}; }
                #pragma unroll
                for (int r = 3; r < R; r++) if (r & (r - 1)) w[r] = cmul(w[r & (r - 1)], w[r & -r]);   // clear lowest bit x lowest bit
                #pragma unroll
                for (int r = 1; r < R; r++) v[m][r] = cmul(v[m][r], w[r]);
            }
            dftR<R>(v[m]);
        }
    }
    __syncthreads();
    if (tid < T) {
        #pragma unroll
        for (int m = 0; m < M; m++) {
            const int j = tid + T*m;
            const int k = j & (Ns - 1);
            const int j0 = ((j - k)*R) + k;
            #pragma unroll
            for (int r = 0; r < R; r++) buf[pad(j0 + r*Ns)] = v[m][r];
        }
    }
    __syncthreads();
}

// The first radix-16 pass (Ns = 1: no twiddles) with its 16 inputs already in registers: thread `tid` owns points
// tid + r*N/16 — the windowed samples it loaded itself — and scatters the butterfly's outputs into `buf`.
template <int LOG2N>
__device__ __forceinline__ void stockham_first_pass(cpx* buf, cpx (&v)[16], int tid) {
    constexpr int N = 1 << LOG2N, T = N/16;
    if (tid < T) {
        dft16(v);
        #pragma unroll
        for (int r = 0; r < 16; r++) buf[pad(tid*16 + r)] = v[r];
    }
    __syncthreads();
}

// Passes 2.. of the transform (the first one is stockham_first_pass)
template <int LOG2N>
__device__ __forceinline__ void fft_inplace(cpx* buf, const float2* tw, int tid) {
    using L = TwiddleLayout<LOG2N>;
    constexpr int FULL = LOG2N/4, REM = LOG2N%4;
    int Ns = 16;
    #pragma unroll
    for (int p = 1; p < FULL; p++) { stockham_pass<LOG2N, 16>(buf, tw + L::offset(p), Ns, tid); Ns *= 16; }
    if constexpr (REM == 1) stockham_pass<LOG2N, 2>(buf, tw + L::offset(FULL), Ns, tid);
    if constexpr (REM == 2) stockham_pass<LOG2N, 4>(buf, tw + L::offset(FULL), Ns, tid);
    if constexpr (REM == 3) stockham_pass<LOG2N, 8>(buf, tw + L::offset(FULL), Ns, tid);
}

struct StftParams {
    const float* pcm; long long n_samples; int channels;
    const long long* tell; int n_frames;
    const float2* twiddle; const float* window;
    int magnitude, volume;
    const int* indptr; const int* indices; const float* data; int bins;
    float* mag_out; float* spec_out;
};

__device__ __forceinline__ float apply_volume(int kind, float x) {
    switch (kind) {
        case SFB_VOLUME_SQRT:       return sqrtf(x);
        case SFB_VOLUME_DBFS:       return 10.0f*log10f(x);
        case SFB_VOLUME_DBFS_TREMX: return 10.0f*(log10f(x + 0.1f) + 1.0f)/1.0414f;
        default:                    return x;
    }
}

template <int LOG2N>
__global__ void __launch_bounds__((1 << LOG2N)/16 < 32 ? 32 : (1 << LOG2N)/16, ((1 << LOG2N) <= 4096) ? 3 : 1)
stft_mel_kernel(const __grid_constant__ StftParams P) {
    constexpr int N = 1 << LOG2N, BINS = N/2 + 1, TW = TwiddleLayout<LOG2N>::total;
    extern __shared__ __align__(16) unsigned char smem[];
    cpx*    buf = reinterpret_cast<cpx*>(smem);                               // pad(N) complex
    float2* tw  = reinterpret_cast<float2*>(smem + sizeof(cpx)*(N + N/16));   // per-pass twiddle tables
    float*  mag = reinterpret_cast<float*>(smem);                             // [2][BINS], aliases buf after the FFT
    const int tid = threadIdx.x, nthreads = blockDim.x;

    for (int i = tid; i < TW; i += nthreads) tw[i] = P.twiddle[i];
    // The window coefficients this thread needs are the same for every frame: thread `tid` always owns the
    // window positions tid + r*N/16 (the inputs of its first butterfly) — 16 registers, loaded once
    constexpr int T = N/16;
    float wreg[16];
    #pragma unroll
    for (int r = 0; r < 16; r++) wreg[r] = (tid < T) ? __ldg(P.window + tid + r*T) : 0.0f;

    for (int frame = blockIdx.x; frame < P.n_frames; frame += gridDim.x) {
        // ---- 1. window [tell-N-1, tell-1) of both channels straight from the clip into the first butterfly:
        //         lane-consecutive (coalesced) loads at stride N/16, window multiply, z = L + iR -------------
        const long long lo = P.tell[frame] - N - 1;
        const int before = (lo < 0) ? int(-lo > N ? N : -lo) : 0;                 // window samples before the clip start
        const long long room = P.n_samples - lo;
        const int avail = room > N ? N : (room < 0 ? 0 : int(room));             // window samples inside the clip
        cpx v[16];
        if (tid < T) {
            const float* xl = P.pcm + lo;                                         // may point before the clip: guarded
            const float* xr = P.pcm + P.n_samples + lo;
            const bool whole = (before == 0) && (avail == N);
            #pragma unroll
            for (int r = 0; r < 16; r++) {
                const int n = tid + r*T;
                const bool in = whole || (n >= before && n < avail);
                const float l = in ? __ldg(xl + n) : 0.0f;
                const float rr = (in && P.channels == 2) ? __ldg(xr + n) : 0.0f;
                v[r] = cpx{l*wreg[r], rr*wreg[r]};
            }
        }
        __syncthreads();                                                          // the previous frame is done with buf

        // ---- 2. FFT ------------------------------------------------------------------------
        stockham_first_pass<LOG2N>(buf, v, tid);
        fft_inplace<LOG2N>(buf, tw, tid);

        // ---- 3. split L/R spectra, magnitude (registers), then park in smem -----------------
        constexpr int PER = (BINS + (N/16 < 32 ? 32 : N/16) - 1)/(N/16 < 32 ? 32 : N/16);
        float ml[PER], mr[PER];
        #pragma unroll
        for (int s = 0; s < PER; s++) {
            const int k = tid + s*nthreads;
            ml[s] = mr[s] = 0.0f;
            if (k < BINS) {
                const cpx a = buf[pad(k)], b = buf[pad((N - k) & (N - 1))];
                // L̂ = (Z[k] + conj Z[N-k])/2 ; R̂ = (Z[k] - conj Z[N-k])/(2i)
                const cpx L = {0.5f*(a.x + b.x), 0.5f*(a.y - b.y)};
                const cpx R = {0.5f*(a.y + b.y), 0.5f*(b.x - a.x)};
                float pl = L.x*L.x + L.y*L.y, pr = R.x*R.x + R.y*R.y;
                if (P.magnitude == SFB_MAGNITUDE_AMPLITUDE) { pl = sqrtf(pl); pr = sqrtf(pr); }
                ml[s] = pl; mr[s] = pr;
            }
        }
        __syncthreads();
        #pragma unroll
        for (int s = 0; s < PER; s++) {
            const int k = tid + s*nthreads;
            if (k < BINS) {
                mag[k] = ml[s]; mag[BINS + k] = mr[s];
                if (P.mag_out) {
                    float* o = P.mag_out + (size_t)frame*P.channels*BINS;
                    o[k] = ml[s];
                    if (P.channels == 2) o[BINS + k] = mr[s];
                }
            }
        }
        __syncthreads();

        // ---- 4. filterbank rows (ascending column order, like scipy's csr_matvecs) ----------
        if (P.spec_out) {
            float* o = P.spec_out + (size_t)frame*P.bins*P.channels;
            for (int b = tid; b < P.bins; b += nthreads) {
                const int beg = __ldg(P.indptr + b), end = __ldg(P.indptr + b + 1);
                float accl = 0.0f, accr = 0.0f;
                for (int t = beg; t < end; t++) {
                    const int col = __ldg(P.indices + t);
                    const float val = __ldg(P.data + t);
                    accl = __fadd_rn(accl, __fmul_rn(val, mag[col]));
                    accr = __fadd_rn(accr, __fmul_rn(val, mag[BINS + col]));
                }
                accl = apply_volume(P.volume, accl); accr = apply_volume(P.volume, accr);
                if (P.channels == 2) reinterpret_cast<float2*>(o)[b] = make_float2(accl, accr);
                else o[b] = accl;
            }
        }
    }
}

// -------------------------------------------------------------------------------------------------

static int stft_tables(sfb_ctx* ctx, int fft_n, int window_kind) {
    const int N = 1 << fft_n;
    if (!ctx->twiddle[fft_n]) {
        // per-pass tables [log2 R][Ns] of exp(-2*pi*i * m*k/(Ns*R)), m = 1, 2, 4, 8 (see TwiddleLayout)
        std::vector<float2> tw;
        const int full = fft_n/4, rem = fft_n%4;
        int Ns = 1;
        for (int p = 0; p < full + (rem ? 1 : 0); p++) {
            const int R = (p < full) ? 16 : (1 << rem);
            if (p >= 1)
                for (int m = 1; m < R; m <<= 1)
                    for (int k = 0; k < Ns; k++) {
                        const double a = -2.0*M_PI*double(m)*double(k)/(double(Ns)*double(R));
                        tw.push_back(make_float2(float(cos(a)), float(sin(a))));
                    }
            Ns *= R;
        }
        if (tw.empty()) tw.push_back(make_float2(1.0f, 0.0f));
        SFB_CUDA(cudaMalloc(&ctx->twiddle[fft_n], sizeof(float2)*tw.size()));
        SFB_CUDA(cudaMemcpy(ctx->twiddle[fft_n], tw.data(), sizeof(float2)*tw.size(), cudaMemcpyHostToDevice));
    }
    if (!ctx->window[window_kind][fft_n]) {
        std::vector<float> w(N);
        for (int n = 0; n < N; n++) {
            double v = 1.0;
            if (window_kind == SFB_WINDOW_HANNING)            // np.hanning: symmetric
                v = 0.5 - 0.5*cos(2.0*M_PI*double(n)/double(N - 1));
            else if (window_kind == SFB_WINDOW_HANN_POISSON)  // spectrogram.py:93-98
                v = 0.5*(1.0 - cos(2.0*M_PI*double(n)/double(N)))*exp(-2.0*fabs(double(N - 2*n))/double(N));
            w[n] = float(v);
        }
        SFB_CUDA(cudaMalloc(&ctx->window[window_kind][fft_n], sizeof(float)*N));
        SFB_CUDA(cudaMemcpy(ctx->window[window_kind][fft_n], w.data(), sizeof(float)*N, cudaMemcpyHostToDevice));
    }
    return SFB_OK;
}

template <int LOG2N>
static int stft_launch(sfb_ctx* ctx, const StftParams& P) {
    constexpr int N = 1 << LOG2N;
    const int threads = (N/16 < 32) ? 32 : N/16;
    const size_t smem = sizeof(cpx)*(N + N/16) + sizeof(float2)*(TwiddleLayout<LOG2N>::total > 0 ? TwiddleLayout<LOG2N>::total : 1);
    static bool configured[16] = {};
    auto kernel = stft_mel_kernel<LOG2N>;
    if (!configured[LOG2N]) {
        SFB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        configured[LOG2N] = true;
    }
    int per_sm = 1;
    SFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    if (per_sm < 1) per_sm = 1;
    int grid = ctx->sm_count*per_sm;
    if (grid > P.n_frames) grid = P.n_frames;
    kernel<<<grid, threads, smem, ctx->stream>>>(P);
    SFB_LAUNCH_CHECK(ctx);
    return SFB_OK;
}

extern "C" int sfb_stft_mel(sfb_ctx* ctx, const float* pcm_dev, int64_t n_samples, int channels, int fft_n,
                            const int64_t* tell_dev, int n_frames, int window_kind, int magnitude_kind,
                            const int32_t* csr_indptr_dev, const int32_t* csr_indices_dev, const float* csr_data_dev,
                            int bins, int volume_kind, float* mag_out_dev, float* spec_out_dev) {
    SFB_REQUIRE(ctx, "sfb_stft_mel: null ctx");
    if (n_frames == 0) return SFB_OK;
    SFB_REQUIRE(pcm_dev && tell_dev, "sfb_stft_mel: null pcm / tell");
    SFB_REQUIRE(channels == 1 || channels == 2, "sfb_stft_mel: channels must be 1 or 2 (got %d)", channels);
    SFB_REQUIRE(fft_n >= 8 && fft_n <= 13, "sfb_stft_mel: fft_n %d outside 8..13", fft_n);
    SFB_REQUIRE(n_samples >= 0 && n_frames >= 0, "sfb_stft_mel: negative size");
    SFB_REQUIRE(window_kind >= 0 && window_kind <= 2, "sfb_stft_mel: bad window kind %d", window_kind);
    SFB_REQUIRE(magnitude_kind == 0 || magnitude_kind == 1, "sfb_stft_mel: bad magnitude kind %d", magnitude_kind);
    SFB_REQUIRE(volume_kind >= 0 && volume_kind <= 3, "sfb_stft_mel: bad volume kind %d", volume_kind);
    SFB_REQUIRE(!spec_out_dev || (csr_indptr_dev && csr_indices_dev && csr_data_dev && bins > 0),
        "sfb_stft_mel: spec output needs a filterbank");
    SFB_REQUIRE(mag_out_dev || spec_out_dev, "sfb_stft_mel: nothing to compute");
    if (int e = stft_tables(ctx, fft_n, window_kind)) return e;
    StftParams P{pcm_dev, (long long)n_samples, channels, (const long long*)tell_dev, n_frames,
                 ctx->twiddle[fft_n], ctx->window[window_kind][fft_n], magnitude_kind, volume_kind,
                 csr_indptr_dev, csr_indices_dev, csr_data_dev, bins, mag_out_dev, spec_out_dev};
    switch (fft_n) {
        case 8:  return stft_launch<8>(ctx, P);
        case 9:  return stft_launch<9>(ctx, P);
        case 10: return stft_launch<10>(ctx, P);
        case 11: return stft_launch<11>(ctx, P);
        case 12: return stft_launch<12>(ctx, P);
        default: return stft_launch<13>(ctx, P);
    }
}

// =================================================================================================
// K2
// =================================================================================================

__device__ __forceinline__ double warp_sum(double v) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0;
    if (warp == 0) { t = warp_sum(t); if (lane == 0) red[0] = t; }
    __syncthreads();
    return red[0];
}

// (2a) volume / std targets: one CTA per frame over clip samples [tell-n-1, tell-1) of every channel
// audio/module.py:457-458: vol = 2*sqrt(mean(x²))*sqrt(2); std = np.std(x). float32 results.
__global__ void __launch_bounds__(256) scalar_targets_kernel(const float* pcm, long long n_samples, int channels,
        const long long* tell, int n_frames, int n_last, float* vol_target, float* std_target) {
    __shared__ double red[32];
    const int frame = blockIdx.x;
    const long long hi = tell[frame] - 1, lo = hi - n_last;
    double s1 = 0.0, s2 = 0.0;
    for (int c = 0; c < channels; c++) {
        const float* x = pcm + (long long)c*n_samples;
        for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
            const float v = (i >= 0) ? __ldg(x + i) : 0.0f;
            s1 += double(v); s2 += double(v)*double(v);
        }
    }
    const double count = double(n_last)*double(channels);
    const double sum1 = block_sum(s1, red);
    const double sum2 = block_sum(s2, red);
    const double mean = sum1/count;
    // second pass for the variance, like np.std (mean of squared deviations)
    double d2 = 0.0;
    for (int c = 0; c < channels; c++) {
        const float* x = pcm + (long long)c*n_samples;
        for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
            const double v = ((i >= 0) ? double(__ldg(x + i)) : 0.0) - mean;
            d2 += v*v;
        }
    }
    const double var = block_sum(d2, red)/count;
    if (threadIdx.x == 0) {
        const float rms = sqrtf(float(sum2/count));
        vol_target[frame] = __fmul_rn(__fmul_rn(2.0f, rms), 1.41421356237309504880f);
        std_target[frame] = sqrtf(float(var));
    }
}

// (2b) scalar dynamics: one thread, sequential over frames, float64 state with float32 targets
// (dynamics.py:197-250 under numpy's promotion rules, see oracle/audio_np.py)
struct DynCoeff { double k1, k2; };

__device__ __forceinline__ DynCoeff dyn_coefficients(double frequency, double zeta, double dt) {
    const double pi = 3.141592653589793, tau = 6.283185307179586;
    const double radians = tau*frequency;
    DynCoeff c;
    if (radians*dt < zeta) {
        c.k1 = zeta/(pi*frequency);
        const double k2 = 1.0/(radians*radians);
        c.k2 = fmax(fmax(c.k1*dt, k2), 0.5*(c.k1 + dt)*dt);
    } else {
        const double damping = radians*sqrt(fabs(zeta*zeta - 1.0));
        const double t1 = exp(-1.0*zeta*radians*dt);
        const double a1 = 2.0*t1*((zeta <= 1.0) ? cos(damping*dt) : cosh(damping*dt));
        const double t2 = 1.0/(1.0 + t1*t1 - a1)*dt;
        c.k1 = t2*(1.0 - t1*t1);
        c.k2 = t2*dt;
    }
    return c;
}

__global__ void scalar_scan_kernel(const float* vol_target, const float* std_target, const double* dt,
                                   int n_frames, double* scalars) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    // audio/module.py:413-421: volume f=2 ζ=1 r=0 integrate; std f=10 ζ=1 r=0
    double vv = 0.0, vd = 0.0, vi = 0.0, sv = 0.0, sd = 0.0;
    for (int k = 0; k < n_frames; k++) {
        const double step = fabs(dt[k]);
        const double vt = double(vol_target[k]), st = double(std_target[k]);
        if (step != 0.0) {
            if (fabs(vt - vv) < 1e-6) {
                vi += vv*step;
            } else {
                const DynCoeff c = dyn_coefficients(2.0, 1.0, step);
                vv += vd*step;
                const double acc = (vt + 0.0 - vv - c.k1*vd)/c.k2;
                vd += acc*step;
                vi += vv*step;
            }
            if (!(fabs(st - sv) < 1e-6)) {
                const DynCoeff c = dyn_coefficients(10.0, 1.0, step);
                sv += sd*step;
                const double acc = (st + 0.0 - sv - c.k1*sd)/c.k2;
                sd += acc*step;
            }
        }
        double* o = scalars + (size_t)k*SFB_SCALARS;
        o[SFB_SCALAR_VOLUME] = vv; o[SFB_SCALAR_VOLUME_INTEGRAL] = vi; o[SFB_SCALAR_STD] = sv;
        o[SFB_SCALAR_VOLUME_TARGET] = vt; o[SFB_SCALAR_STD_TARGET] = st;
    }
}

// (1) spectrogram dynamics: one CTA, lanes = bins*channels state elements in registers, sequential
// over frames; float32 arithmetic without FMA contraction so it is bit-exact against numpy.
constexpr int SCAN_THREADS = 1024, SCAN_ELEMS = 4;

__global__ void __launch_bounds__(SCAN_THREADS) spec_scan_kernel(float* spec, int lanes, const double* dt,
        int n_frames, sfb_dynamics_params prm) {
    float value[SCAN_ELEMS], deriv[SCAN_ELEMS], prev[SCAN_ELEMS], target[SCAN_ELEMS], nxt[SCAN_ELEMS];
    #pragma unroll
    for (int e = 0; e < SCAN_ELEMS; e++) { value[e] = deriv[e] = prev[e] = 0.0f; nxt[e] = 0.0f; }
    const float precision = float(prm.precision);
    const float k3 = float((prm.response*prm.zeta)/(6.283185307179586*prm.frequency));
    auto load = [&](int k, float* dstv) {
        #pragma unroll
        for (int e = 0; e < SCAN_ELEMS; e++) {
            const int lane = threadIdx.x + e*SCAN_THREADS;
            dstv[e] = (lane < lanes && k < n_frames) ? spec[(size_t)k*lanes + lane] : 0.0f;
        }
    };
    load(0, nxt);
    for (int k = 0; k < n_frames; k++) {
        #pragma unroll
        for (int e = 0; e < SCAN_ELEMS; e++) target[e] = nxt[e];
        load(k + 1, nxt);                                  // prefetch: targets do not depend on state
        const double step = fabs(dt[k]);
        if (step != 0.0) {
            int moving = 0;
            #pragma unroll
            for (int e = 0; e < SCAN_ELEMS; e++) {
                const int lane = threadIdx.x + e*SCAN_THREADS;
                if (lane < lanes && !(fabsf(__fsub_rn(target[e], value[e])) < precision)) moving = 1;
            }
            if (__syncthreads_or(moving)) {               // np.abs(target - value).max() < precision
                const DynCoeff c = dyn_coefficients(prm.frequency, prm.zeta, step);
                const float dtf = float(step), k1 = float(c.k1), k2 = float(c.k2);
                #pragma unroll
                for (int e = 0; e < SCAN_ELEMS; e++) {
                    const float velocity = __fdiv_rn(__fsub_rn(target[e], prev[e]), dtf);
                    prev[e] = target[e];
                    value[e] = __fadd_rn(value[e], __fmul_rn(deriv[e], dtf));
                    const float num = __fsub_rn(__fsub_rn(__fadd_rn(target[e], __fmul_rn(k3, velocity)), value[e]),
                                                __fmul_rn(k1, deriv[e]));
                    const float acc = __fdiv_rn(num, k2);
                    deriv[e] = __fadd_rn(deriv[e], __fmul_rn(acc, dtf));
                }
            }
        }
        #pragma unroll
        for (int e = 0; e < SCAN_ELEMS; e++) {
            const int lane = threadIdx.x + e*SCAN_THREADS;
            if (lane < lanes) spec[(size_t)k*lanes + lane] = value[e];
        }
    }
}

// (3) waveform: reducer per absolute chunk m = samples [chunk*m-1, chunk*(m+1)-1) (waveform.py:64-83:
// the slice is aligned to multiples of the chunk size minus the excluded newest sample), one warp per
// (m, channel); then frame k gathers rows q_k-points .. q_k-1 with q_k = tell_k // chunk.
__global__ void __launch_bounds__(256) waveform_chunks_kernel(const float* pcm, long long n_samples, int channels,
        long long m0, int n_chunks, int chunk, int reducer, float* table /*[n_chunks][channels]*/) {
    const int warp = (blockIdx.x*blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_chunks*channels) return;
    const int c = warp % channels; const long long m = m0 + warp/channels;
    const float* x = pcm + (long long)c*n_samples;
    const long long lo = m*chunk - 1, hi = lo + chunk;
    double s1 = 0.0, s2 = 0.0;
    for (long long i = lo + lane; i < hi; i += 32) {
        const float v = (i >= 0 && i < n_samples) ? __ldg(x + i) : 0.0f;
        if (reducer == SFB_REDUCER_AVERAGE) s1 += double(fabsf(v));
        else { s1 += double(v); s2 += double(v)*double(v); }
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (lane == 0) {
        float r;
        if (reducer == SFB_REDUCER_AVERAGE)  r = sqrtf(float(s1/chunk));                       // sqrt(mean|x|)
        else if (reducer == SFB_REDUCER_RMS) r = sqrtf(__fmul_rn(sqrtf(float(s2/chunk)), 1.41421356237309504880f));
        else { const double mean = s1/chunk; r = sqrtf(sqrtf(float(fmax(s2/chunk - mean*mean, 0.0)))); }
        table[(size_t)(warp/channels)*channels + c] = r;
    }
}

__global__ void __launch_bounds__(256) waveform_gather_kernel(const float* table, long long m0, int channels,
        const long long* tell, int n_frames, int points, int chunk, float* wave) {
    const int frame = blockIdx.x;
    const long long q = tell[frame]/chunk;
    for (int i = threadIdx.x; i < points*channels; i += blockDim.x) {
        const long long m = q - points + i/channels;
        wave[(size_t)frame*points*channels + i] = table[(size_t)(m - m0)*channels + (i % channels)];
    }
}

extern "C" int sfb_audio_track(sfb_ctx* ctx, const float* pcm_dev, int64_t n_samples, int channels, int samplerate,
                               const int64_t* tell_dev, const double* dt_dev, int n_frames,
                               float* spec_inout_dev, int bins, const sfb_dynamics_params* spec_dynamics,
                               double* scalars_out_dev,
                               float* wave_out_dev, int wave_points, int wave_chunk, int wave_reducer) {
    SFB_REQUIRE(ctx, "sfb_audio_track: null ctx");
    if (n_frames == 0) return SFB_OK;
    SFB_REQUIRE(pcm_dev && tell_dev && dt_dev, "sfb_audio_track: null pcm / tell / dt");
    SFB_REQUIRE(channels >= 1 && channels <= 8 && samplerate > 0 && n_frames > 0, "sfb_audio_track: bad shape");
    if (spec_inout_dev) {
        SFB_REQUIRE(spec_dynamics && bins > 0, "sfb_audio_track: spectrogram scan needs bins and dynamics parameters");
        const int lanes = bins*channels;
        SFB_REQUIRE(lanes <= SCAN_THREADS*SCAN_ELEMS, "sfb_audio_track: bins*channels = %d exceeds %d", lanes, SCAN_THREADS*SCAN_ELEMS);
        spec_scan_kernel<<<1, SCAN_THREADS, 0, ctx->stream>>>(spec_inout_dev, lanes, dt_dev, n_frames, *spec_dynamics);
        SFB_LAUNCH_CHECK(ctx);
    }
    // One scratch request for everything below (per-frame targets + the waveform chunk table): the table
    // size depends on tell[], which lives on the device, so its two ends are fetched first
    long long m0 = 0; int n_chunks = 0;
    if (wave_out_dev) {
        SFB_REQUIRE(wave_points > 0 && wave_chunk > 0, "sfb_audio_track: bad waveform shape");
        SFB_REQUIRE(wave_reducer >= 0 && wave_reducer <= 2, "sfb_audio_track: bad reducer %d", wave_reducer);
        // tell is non-decreasing: chunks needed span [q_first - points, q_last)
        int64_t ends[2];
        SFB_CUDA(cudaMemcpyAsync(&ends[0], tell_dev, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        SFB_CUDA(cudaMemcpyAsync(&ends[1], tell_dev + (n_frames - 1), sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        SFB_CUDA(cudaStreamSynchronize(ctx->stream));
        m0 = ends[0]/wave_chunk - wave_points;
        n_chunks = int(ends[1]/wave_chunk - m0);
    }
    const size_t target_floats = scalars_out_dev ? 2*size_t(n_frames) : 0;
    const size_t table_floats = size_t(n_chunks > 0 ? n_chunks : 0)*size_t(channels);
    float* scratch = nullptr;
    if (target_floats + table_floats) {
        void* raw = nullptr;
        if (int e = sfb_ctx_scratch(ctx, sizeof(float)*(target_floats + table_floats), &raw)) return e;
        scratch = static_cast<float*>(raw);
    }
    if (scalars_out_dev) {
        float* vt = scratch; float* st = vt + n_frames;
        const int n_last = int(0.1*double(samplerate));           // get_last_n_seconds(0.1): int(n + offset)
        scalar_targets_kernel<<<n_frames, 256, 0, ctx->stream>>>(pcm_dev, (long long)n_samples, channels,
            (const long long*)tell_dev, n_frames, n_last, vt, st);
        SFB_LAUNCH_CHECK(ctx);
        scalar_scan_kernel<<<1, 32, 0, ctx->stream>>>(vt, st, dt_dev, n_frames, scalars_out_dev);
        SFB_LAUNCH_CHECK(ctx);
    }
    if (wave_out_dev && n_chunks > 0) {
        float* table = scratch + target_floats;
        const int warps = n_chunks*channels;
        waveform_chunks_kernel<<<(warps*32 + 255)/256, 256, 0, ctx->stream>>>(pcm_dev, (long long)n_samples, channels,
            m0, n_chunks, wave_chunk, wave_reducer, table);
        SFB_LAUNCH_CHECK(ctx);
        waveform_gather_kernel<<<n_frames, 256, 0, ctx->stream>>>(table, m0, channels, (const long long*)tell_dev,
            n_frames, wave_points, wave_chunk, wave_out_dev);
        SFB_LAUNCH_CHECK(ctx);
    }
    return SFB_OK;
}

// DynamicNumber.next over frames for any float32 vector (ShaderPiano's key presses, piano/module.py:63-67,268):
// the same sequential scan as the spectrogram columns, in place on values_inout_dev [n_frames][lanes]
extern "C" int sfb_dynamics_scan(sfb_ctx* ctx, float* values_inout_dev, int lanes, const double* dt_dev, int n_frames,
                                 const sfb_dynamics_params* params) {
    SFB_REQUIRE(ctx && values_inout_dev && dt_dev && params, "sfb_dynamics_scan: null argument");
    SFB_REQUIRE(n_frames > 0 && lanes > 0 && lanes <= SCAN_THREADS*SCAN_ELEMS,
        "sfb_dynamics_scan: lanes = %d outside 1..%d", lanes, SCAN_THREADS*SCAN_ELEMS);
    spec_scan_kernel<<<1, SCAN_THREADS, 0, ctx->stream>>>(values_inout_dev, lanes, dt_dev, n_frames, *params);
    SFB_LAUNCH_CHECK(ctx);
    return SFB_OK;
}

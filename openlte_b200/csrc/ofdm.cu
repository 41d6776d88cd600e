// OFDM demodulation for batches of subframes: CP removal + N-point forward DFT + DC drop / half-spectrum swap.
// Twin of samples_to_symbols_dl, /root/reference/liblte/src/liblte_phy.cc:8593-8644 (the reference calls FFTW3f
// there; "any correct unnormalised forward DFT with f32 I/O" is the contract).
//
// One CTA per OFDM symbol.  The symbol's N samples (float2, or int8 pairs straight from a capture file) are
// loaded coalesced into shared memory, transformed by a Stockham autosort FFT (radix-4 passes + one radix-2
// pass when log2 N is odd) that ping-pongs between two shared buffers, and the 12*N_rb used bins are written
// coalesced to the resource grid.  Twiddles come from a table computed in double precision on the host.
// HBM-bound by design: per symbol it reads N*8 B (2 B/sample for int8 input) and writes 12*N_rb*8 B.
#include "common.cuh"
#include "fft_dev.cuh"
#include <math.h>

namespace {

struct OfdmParams {
    int N, log2N, cp0, cpe, slot;   // FFT size, CP lengths, samples per slot
    int h;                          // 6*N_rb_dl
    int nsc;                        // 12*N_rb_dl (grid row length)
    // front-end frequency shift fused into the sample load (LTE_fdd_dl_fs_samp_buf::freq_shift, fs.cc:696-713); 0 = off
    float cfo_hz;
    double cfo_fs;                  // sample rate in Hz
    long long cfo_base;             // stream index of d_iq[0] in the caller's buffer numbering
};

// one sample of freq_shift(): phase = (i+1)*freq_offset (float product) * 2*pi / fs (double), cosf/sinf of the float-rounded phase
__device__ __forceinline__ float2 cfo_rotate(float2 v, long long i, float f_hz, double fs)
{
    const float t = __fmul_rn((float)(unsigned)(i + 1), f_hz);
    const float ph = (float)((double)t * 2.0 * 3.14159265358979323846 / fs);
    float sn, cs;
    sincosf(ph, &sn, &cs);
    return make_float2(__fadd_rn(__fmul_rn(v.x, cs), __fmul_rn(v.y, sn)), __fsub_rn(__fmul_rn(v.y, cs), __fmul_rn(v.x, sn)));
}

template <typename T> __device__ __forceinline__ float2 load_iq(const T *p, long long i);
template <> __device__ __forceinline__ float2 load_iq<float2>(const float2 *p, long long i) { return p[i]; }
template <> __device__ __forceinline__ float2 load_iq<char2>(const char2 *p, long long i)
{
    const char2 v = p[i];
    return make_float2((float)v.x, (float)v.y);
}


// grid: [n_sf][16][nsc] float2.  blockIdx.x = subframe * 16 + symbol.
// WINDOWS == false: blockIdx.x = subframe*16 + symbol, start_idx[subframe] is the subframe start.
// WINDOWS == true : blockIdx.x = window, start_idx[window] is the first sample of the N-sample DFT window; output row = window.
template <typename T, bool WINDOWS, bool CFO = false>
__global__ void __launch_bounds__(512)
ofdm_demod_kernel(const T *__restrict__ iq, long long n_samps, const long long *__restrict__ start_idx, OfdmParams P,
                  const float2 *__restrict__ tw /*[N] e^{-2 pi i k/N}*/, float2 *__restrict__ grid)
{
    extern __shared__ float2 sm[];
    float2 *bufA = sm, *bufB = sm + P.N;
    const int sf = blockIdx.x >> 4, sym = blockIdx.x & 15;
    const int slot = sym / 7, l = sym % 7;
    // window start: slot start + (N+CP_else)*l (+CP_0-CP_else for l>0) + CP_l - 1   (one sample early, :8607-8622)
    long long idx = WINDOWS ? start_idx[blockIdx.x]
                            : start_idx[sf] + (long long)slot * P.slot + (long long)(P.N + P.cpe) * l + (l > 0 ? P.cp0 - P.cpe : 0) +
                                  (l == 0 ? P.cp0 : P.cpe) - 1;
    for (int i = threadIdx.x; i < P.N; i += blockDim.x) {
        const long long s = idx + i;
        float2 v = (s >= 0 && s < n_samps) ? load_iq<T>(iq, s) : make_float2(0.f, 0.f);
        if (CFO) v = cfo_rotate(v, P.cfo_base + s, P.cfo_hz, P.cfo_fs);
        bufA[i] = v;
    }
    __syncthreads();
    float2 *src = stockham_forward(bufA, bufB, P.N, tw);
    // bins -> subcarriers (:8625-8634): symb[h+i] = X[i+1], symb[h-1-i] = X[N-1-i]
    float2 *out = grid + (WINDOWS ? (size_t)blockIdx.x : ((size_t)sf * 16 + sym)) * P.nsc;
    for (int k = threadIdx.x; k < 2 * P.h; k += blockDim.x) {
        const int bin = (k >= P.h) ? (k - P.h + 1) : (P.N - P.h + k);
        out[k] = src[bin];
    }
}


// ------------------------------------------------------------------------------------------------------------------
// Register-resident mixed-radix variant (the batched subframe path).  N = 16*16*8 (2048), 16*16*4, 16*8*4, 16*16, 16*8:
// N/16 threads per symbol, every thread carries 16 complex values through each Stockham pass as one radix-16 butterfly
// (or 16/R radix-R butterflies) held entirely in registers, so a 2048-point transform crosses shared memory twice
// instead of six times and needs three barriers.  The first pass reads the samples straight from global memory
// (coalesced: thread b reads x[b + (N/16) k]), the last one writes the used bins straight to the grid.
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mulmj(float2 a) { return make_float2(a.y, -a.x); }   // a * (-j)
// complex product with fused multiply-adds (this kernel is held to a tolerance, not to the bit pattern of the oracle's FFT: the fused form is the
// more accurate one and is four instructions instead of six)
__device__ __forceinline__ float2 cmulf(float2 a, float2 w)
{
    return make_float2(__fmaf_rn(a.x, w.x, -__fmul_rn(a.y, w.y)), __fmaf_rn(a.x, w.y, __fmul_rn(a.y, w.x)));
}
// a * (1 - j) / sqrt(2) and a * (-1 - j) / sqrt(2): two additions and two products
__device__ __forceinline__ float2 mul_w8_1(float2 a) { const float r = 0.70710678118654752440f; return make_float2(__fmul_rn(r, a.x + a.y), __fmul_rn(r, a.y - a.x)); }
__device__ __forceinline__ float2 mul_w8_3(float2 a) { const float r = 0.70710678118654752440f; return make_float2(__fmul_rn(r, a.y - a.x), -__fmul_rn(r, a.x + a.y)); }

template <int R> struct Dft;
template <> struct Dft<4> {
    __device__ __forceinline__ static void run(float2 *x)
    {
        const float2 t0 = cadd(x[0], x[2]), t1 = csub(x[0], x[2]), t2 = cadd(x[1], x[3]), t3 = mulmj(csub(x[1], x[3]));
        x[0] = cadd(t0, t2); x[1] = cadd(t1, t3); x[2] = csub(t0, t2); x[3] = csub(t1, t3);
    }
};
template <> struct Dft<8> {
    __device__ __forceinline__ static void run(float2 *x)
    {
        float2 e[4] = {x[0], x[2], x[4], x[6]}, o[4] = {x[1], x[3], x[5], x[7]};
        Dft<4>::run(e); Dft<4>::run(o);
        const float2 o1 = mul_w8_1(o[1]), o2 = mulmj(o[2]), o3 = mul_w8_3(o[3]);
        x[0] = cadd(e[0], o[0]); x[4] = csub(e[0], o[0]);
        x[1] = cadd(e[1], o1);   x[5] = csub(e[1], o1);
        x[2] = cadd(e[2], o2);   x[6] = csub(e[2], o2);
        x[3] = cadd(e[3], o3);   x[7] = csub(e[3], o3);
    }
};
template <> struct Dft<16> {
    __device__ __forceinline__ static void run(float2 *x)
    {
        float2 e[8], o[8];
#pragma unroll
        for (int i = 0; i < 8; i++) { e[i] = x[2 * i]; o[i] = x[2 * i + 1]; }
        Dft<8>::run(e); Dft<8>::run(o);
        // W16^k = exp(-j pi k / 8)
        const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, r = 0.70710678118654752440f;
        const float2 w[8] = {make_float2(1.f, 0.f), make_float2(c1, -s1), make_float2(r, -r), make_float2(s1, -c1),
                             make_float2(0.f, -1.f), make_float2(-s1, -c1), make_float2(-r, -r), make_float2(-c1, -s1)};
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const float2 t = (k == 0) ? o[0] : ((k == 4) ? mulmj(o[4]) : ((k == 2) ? mul_w8_1(o[2]) : ((k == 6) ? mul_w8_3(o[6]) : cmulf(o[k], w[k]))));
            x[k] = cadd(e[k], t); x[k + 8] = csub(e[k], t);
        }
    }
};

// One Stockham pass of radix R over a transform of current length n and stride s (n * s == N); thread `tid` handles
// butterflies tid, tid + NTHR, ...  SRC_GLOBAL: inputs come from the sample stream (first pass); DST_GRID: outputs go
// to the resource grid through the bin -> subcarrier map (last pass, where p == 0 and no twiddle applies).
// PAD_SRC / PAD_DST: the shared buffer is indexed through pad16() (one spare element per 16): the first pass writes 16 consecutive
// elements per thread, which without the padding puts every lane of a warp on the same banks.
__device__ __forceinline__ int pad16(int i) { return i + (i >> 4); }
template <int N, int R, int n, typename T, bool SRC_GLOBAL, bool DST_GRID, bool CFO = false, bool PAD_SRC = false, bool PAD_DST = false, bool CHECK = true>
__device__ __forceinline__ void stockham_pass(const T *__restrict__ iq, long long w0, long long n_samps, const float2 *__restrict__ src,
                                              float2 *__restrict__ dst, const float2 *__restrict__ tw, float2 *__restrict__ out, int h, int tid,
                                              float cfo_hz = 0.f, double cfo_fs = 1.0, long long cfo_base = 0)
{
    constexpr int NTHR = N / 16, s = N / n, n1 = n / R;
#pragma unroll
    for (int rep = 0; rep < 16 / R; rep++) {
        const int b = tid + rep * NTHR;            // butterfly index in [0, N/R)
        const int p = b / s, q = b - p * s;
        float2 x[R];
#pragma unroll
        for (int k = 0; k < R; k++) {
            const int i = q + s * (p + n1 * k);
            if (SRC_GLOBAL) {
                const long long sidx = w0 + i;
                x[k] = (!CHECK || (sidx >= 0 && sidx < n_samps)) ? load_iq<T>(iq, sidx) : make_float2(0.f, 0.f);
                if (CFO) x[k] = cfo_rotate(x[k], cfo_base + sidx, cfo_hz, cfo_fs);
            } else x[k] = src[PAD_SRC ? pad16(i) : i];
        }
        Dft<R>::run(x);
#pragma unroll
        for (int k = 0; k < R; k++) {
            if (DST_GRID) {
                const int bin = q + s * k;         // p == 0 in the last pass
                if (bin >= 1 && bin <= h) out[h + bin - 1] = x[k];
                else if (bin >= N - h) out[bin - (N - h)] = x[k];
            } else {
                const float2 v = (k == 0 || p == 0) ? x[k] : cmulf(x[k], __ldg(&tw[p * k * s]));
                dst[PAD_DST ? pad16(q + s * (R * p + k)) : q + s * (R * p + k)] = v;
            }
        }
    }
}

// STAGE (A/B variant, LTEB200_OFDM_TMA=1; int8 input, N = 2048): the symbol's samples come in by ONE bulk copy (cp.async.bulk + mbarrier, the
// TMA engine) into shared memory and the first pass reads them from there, instead of sixteen 2-byte global loads per thread.
template <int N, typename T, bool CFO, bool STAGE = false>
__global__ void __launch_bounds__(N / 16 < 32 ? 32 : N / 16)
ofdm_demod_r16_kernel(const T *__restrict__ iq, long long n_samps, const long long *__restrict__ start_idx, OfdmParams P,
                      const float2 *__restrict__ tw, float2 *__restrict__ grid)
{
    __shared__ float2 bufA[N + N / 16], bufB[N];
    __shared__ __align__(16) unsigned char stage[STAGE ? (int)sizeof(T) * N + 32 : 16];
    __shared__ __align__(8) unsigned long long stage_bar;
    const int tid = threadIdx.x;
    const int sf = blockIdx.x >> 4, sym = blockIdx.x & 15;
    const int slot = sym / 7, l = sym % 7;
    const long long w0 = start_idx[sf] + (long long)slot * P.slot + (long long)(P.N + P.cpe) * l + (l > 0 ? P.cp0 - P.cpe : 0) +
                         (l == 0 ? P.cp0 : P.cpe) - 1;
    float2 *out = grid + ((size_t)sf * 16 + sym) * P.nsc;
    const bool act = tid < N / 16;
    const bool inb = w0 >= 0 && w0 + N <= n_samps;          // whole window inside the stream (all but the edges of a capture): no per-sample test
    if constexpr (STAGE) {
        // 16-byte aligned superset of the window's bytes; windows too close to the ends of the buffer take the direct path
        const unsigned long long g0 = (unsigned long long)(iq + w0), a0 = g0 & ~15ull;
        const unsigned off = (unsigned)(g0 - a0), bytes = (off + (unsigned)sizeof(T) * N + 15u) & ~15u;
        const bool staged = w0 >= 16 && w0 + N + 16 <= n_samps;              // (uniform over the block)
        if (staged) {
            const unsigned bar = (unsigned)__cvta_generic_to_shared(&stage_bar), dst = (unsigned)__cvta_generic_to_shared(stage);
            if (tid == 0) {
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(a0), "r"(bytes), "r"(bar) : "memory");
            }
            __syncthreads();
            unsigned done;
            do {
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar) : "memory");
            } while (!done);
            const T *sp = reinterpret_cast<const T *>(stage + off);
            if (act) stockham_pass<N, 16, N, T, true, false, false, false, true, false>(sp, 0, N, nullptr, bufA, tw, out, P.h, tid);
            __syncthreads();
            if (act) stockham_pass<N, 16, 128, T, false, false, false, true, false>(iq, w0, n_samps, bufA, bufB, tw, out, P.h, tid);
            __syncthreads();
            if (act) stockham_pass<N, 8, 8, T, false, true>(iq, w0, n_samps, bufB, nullptr, tw, out, P.h, tid);
            return;
        }
    }
    if constexpr (N == 2048) {          // 16 x 16 x 8
        if (act) { if (inb) stockham_pass<N, 16, 2048, T, true, false, CFO, false, true, false>(iq, w0, n_samps, nullptr, bufA, tw, out, P.h, tid, P.cfo_hz, P.cfo_fs, P.cfo_base); else stockham_pass<N, 16, 2048, T, true, false, CFO, false, true>(iq, w0, n_samps, nullptr, bufA, tw, out, P.h, tid, P.cfo_hz, P.cfo_fs, P.cfo_base); }
        __syncthreads();
        if (act) stockham_pass<N, 16, 128, T, false, false, false, true, false>(iq, w0, n_samps, bufA, bufB, tw, out, P.h, tid);
        __syncthreads();
        if (act) stockham_pass<N, 8, 8, T, false, true>(iq, w0, n_samps, bufB, nullptr, tw, out, P.h, tid);
    } else if constexpr (N == 1024) {   // 16 x 16 x 4
        if (act) { if (inb) stockham_pass<N, 16, 1024, T, true, false, CFO, false, true, false>(iq, w0, n_samps, nullptr, bufA, tw, out, P.h, tid, P.cfo_hz, P.cfo_fs, P.cfo_base); else stockham_pass<N, 16, 1024, T, true, false, CFO, false, true>(iq, w0, n_samps, nullptr, bufA, tw, out, P.h, tid, P.cfo_hz, P.cfo_fs, P.cfo_base); }
        __syncthreads();
        if (act) stockham_pass<N, 16, 64, T, false, false, false, true, false>(iq, w0, n_samps, bufA, bufB, tw, out, P.h, tid);
        __syncthreads();
        if (act) stockham_pass<N, 4, 4, T, false, true>(iq, w0, n_samps, bufB, nullptr, tw, out, P.h, tid);
    } else if constexpr (N == 512) {    // 16 x 8 x 4
        if (act) { if (inb) stockham_pass<N, 16, 512, T, true, false, CFO, false, true, false>(iq, w0, n_samps, nullptr, bufA, tw, out, P.h, tid, P.cfo_hz, P.cfo_fs, P.cfo_base); else stockham_pass<N, 16, 512, T, true, false, CFO, false, true>(iq, w0, n_samps, nullptr, bufA, tw, out, P.h, tid, P.cfo_hz, P.cfo_fs, P.cfo_base); }
        __syncthreads();
        if (act) stockham_pass<N, 8, 32, T, false, false, false, true, false>(iq, w0, n_samps, bufA, bufB, tw, out, P.h, tid);
        __syncthreads();
        if (act) stockham_pass<N, 4, 4, T, false, true>(iq, w0, n_samps, bufB, nullptr, tw, out, P.h, tid);
    } else if constexpr (N == 256) {    // 16 x 16
        if (act) { if (inb) stockham_pass<N, 16, 256, T, true, false, CFO, false, true, false>(iq, w0, n_samps, nullptr, bufA, tw, out, P.h, tid, P.cfo_hz, P.cfo_fs, P.cfo_base); else stockham_pass<N, 16, 256, T, true, false, CFO, false, true>(iq, w0, n_samps, nullptr, bufA, tw, out, P.h, tid, P.cfo_hz, P.cfo_fs, P.cfo_base); }
        __syncthreads();
        if (act) stockham_pass<N, 16, 16, T, false, true, false, true, false>(iq, w0, n_samps, bufA, nullptr, tw, out, P.h, tid);
    } else {                  // 128 = 16 x 8
        if (act) { if (inb) stockham_pass<N, 16, 128, T, true, false, CFO, false, true, false>(iq, w0, n_samps, nullptr, bufA, tw, out, P.h, tid, P.cfo_hz, P.cfo_fs, P.cfo_base); else stockham_pass<N, 16, 128, T, true, false, CFO, false, true>(iq, w0, n_samps, nullptr, bufA, tw, out, P.h, tid, P.cfo_hz, P.cfo_fs, P.cfo_base); }
        __syncthreads();
        if (act) stockham_pass<N, 8, 8, T, false, true, false, true, false>(iq, w0, n_samps, bufA, nullptr, tw, out, P.h, tid);
    }
}

template <typename T>
static void launch_r16(const T *iq, long long n_samps, const long long *d_start, const OfdmParams &P, const float2 *tw, float2 *grid, int n_sf,
                       cudaStream_t stream)
{
    const int blocks = n_sf * 16;
    switch (P.N) {
    case 2048: { if constexpr (sizeof(T) == 2) {
                     if (P.cfo_hz == 0.f && getenv("LTEB200_OFDM_TMA")) { ofdm_demod_r16_kernel<2048, T, false, true><<<blocks, 128, 0, stream>>>(iq, n_samps, d_start, P, tw, grid); break; }
                 }
                 if (P.cfo_hz != 0.f) ofdm_demod_r16_kernel<2048, T, true><<<blocks, 128, 0, stream>>>(iq, n_samps, d_start, P, tw, grid);
                 else ofdm_demod_r16_kernel<2048, T, false><<<blocks, 128, 0, stream>>>(iq, n_samps, d_start, P, tw, grid); } break;
    case 1024: { if (P.cfo_hz != 0.f) ofdm_demod_r16_kernel<1024, T, true><<<blocks, 64, 0, stream>>>(iq, n_samps, d_start, P, tw, grid);
                 else ofdm_demod_r16_kernel<1024, T, false><<<blocks, 64, 0, stream>>>(iq, n_samps, d_start, P, tw, grid); } break;
    case 512:  { if (P.cfo_hz != 0.f) ofdm_demod_r16_kernel<512, T, true><<<blocks, 32, 0, stream>>>(iq, n_samps, d_start, P, tw, grid);
                 else ofdm_demod_r16_kernel<512, T, false><<<blocks, 32, 0, stream>>>(iq, n_samps, d_start, P, tw, grid); } break;
    case 256:  { if (P.cfo_hz != 0.f) ofdm_demod_r16_kernel<256, T, true><<<blocks, 32, 0, stream>>>(iq, n_samps, d_start, P, tw, grid);
                 else ofdm_demod_r16_kernel<256, T, false><<<blocks, 32, 0, stream>>>(iq, n_samps, d_start, P, tw, grid); } break;
    default:   { if (P.cfo_hz != 0.f) ofdm_demod_r16_kernel<128, T, true><<<blocks, 32, 0, stream>>>(iq, n_samps, d_start, P, tw, grid);
                 else ofdm_demod_r16_kernel<128, T, false><<<blocks, 32, 0, stream>>>(iq, n_samps, d_start, P, tw, grid); } break;
    }
}

}  // namespace

struct FftTwiddles { float2 *tw; };
static std::map<std::pair<int, int>, float2 *> g_tw;  // (device, N) -> table
static std::mutex g_tw_mu;

int get_twiddles(lteb200_ctx *ctx, int N, float2 **out)
{
    std::lock_guard<std::mutex> lk(g_tw_mu);
    auto key = std::make_pair(ctx->device, N);
    auto it = g_tw.find(key);
    if (it != g_tw.end()) { *out = it->second; return LTEB200_OK; }
    std::vector<float2> h(N);
    for (int k = 0; k < N; k++) {
        const double a = -2.0 * M_PI * (double)k / (double)N;
        h[k] = make_float2((float)cos(a), (float)sin(a));
    }
    float2 *d = nullptr;
    LTE_CUDA_OK(ctx, cudaMalloc(&d, sizeof(float2) * N));
    LTE_CUDA_OK(ctx, cudaMemcpy(d, h.data(), sizeof(float2) * N, cudaMemcpyHostToDevice));
    g_tw[key] = d;
    *out = d;
    return LTEB200_OK;
}

int fs_legal(int fs_idx, int N_rb_dl)
{
    // liblte_phy_update_n_rb_dl, liblte_phy.cc:2601-2641
    if (fs_idx < 0 || fs_idx > 4) return 0;
    switch (fs_idx) {
    // the reference accepts any N_rb_dl here, but its own subframe struct has rows of 1200 subcarriers (liblte_phy.h:226-239): 100 is the
    // widest carrier that exists (and what the grid / estimate planes and the estimator's shared memory are sized for)
    case 4: return N_rb_dl >= 1 && N_rb_dl <= 100;
    case 3: return N_rb_dl == 50 || N_rb_dl == 25 || N_rb_dl == 15 || N_rb_dl == 6;
    case 2: return N_rb_dl == 25 || N_rb_dl == 15 || N_rb_dl == 6;
    case 1: return N_rb_dl == 15 || N_rb_dl == 6;
    default: return N_rb_dl == 6;
    }
}

int ofdm_demod_launch(lteb200_ctx *ctx, const void *d_iq, int iq_fmt, long long n_samps, int fs_idx, int N_rb_dl,
                      const long long *d_start_idx, int n_sf, float *d_grid, cudaStream_t stream)
{
    if (!fs_legal(fs_idx, N_rb_dl)) { lteb200_set_error(ctx, "illegal (fs, N_rb_dl) = (%d, %d)", fs_idx, N_rb_dl); return LTEB200_ERR_ARGS; }
    OfdmParams P;
    P.N = 128 << fs_idx; P.log2N = 7 + fs_idx; P.cp0 = 10 << fs_idx; P.cpe = 9 << fs_idx; P.slot = 960 << fs_idx;
    P.h = 6 * N_rb_dl; P.nsc = 12 * N_rb_dl;
    P.cfo_hz = ctx->cfo_hz; P.cfo_fs = 1920000.0 * (double)(1 << fs_idx); P.cfo_base = ctx->cfo_base;
    float2 *tw = nullptr;
    int rc = get_twiddles(ctx, P.N, &tw);
    if (rc) return rc;
    const bool radix4 = getenv("LTEB200_FFT_RADIX4") != nullptr;   // the shared-memory radix-4 kernel (bit-identical to the oracle's FFT shim)
    const int threads = P.N / 4 < 64 ? 64 : (P.N / 4 > 512 ? 512 : P.N / 4);
    const size_t smem = sizeof(float2) * 2 * P.N;
    if (iq_fmt == LTEB200_IQ_F32) {
        if (radix4 && P.cfo_hz != 0.f) ofdm_demod_kernel<float2, false, true><<<n_sf * 16, threads, smem, stream>>>((const float2 *)d_iq, n_samps, d_start_idx, P, tw, (float2 *)d_grid);
        else if (radix4) ofdm_demod_kernel<float2, false><<<n_sf * 16, threads, smem, stream>>>((const float2 *)d_iq, n_samps, d_start_idx, P, tw, (float2 *)d_grid);
        else launch_r16<float2>((const float2 *)d_iq, n_samps, d_start_idx, P, tw, (float2 *)d_grid, n_sf, stream);
    } else if (iq_fmt == LTEB200_IQ_I8) {
        if (radix4 && P.cfo_hz != 0.f) ofdm_demod_kernel<char2, false, true><<<n_sf * 16, threads, smem, stream>>>((const char2 *)d_iq, n_samps, d_start_idx, P, tw, (float2 *)d_grid);
        else if (radix4) ofdm_demod_kernel<char2, false><<<n_sf * 16, threads, smem, stream>>>((const char2 *)d_iq, n_samps, d_start_idx, P, tw, (float2 *)d_grid);
        else launch_r16<char2>((const char2 *)d_iq, n_samps, d_start_idx, P, tw, (float2 *)d_grid, n_sf, stream);
    } else { lteb200_set_error(ctx, "unknown I/Q format %d", iq_fmt); return LTEB200_ERR_ARGS; }
    ctx->launches += 1;
    LTE_CUDA_OK(ctx, cudaGetLastError());
    return LTEB200_OK;
}

// N-point forward DFT of arbitrary sample windows (float2 stream): row w of d_out = the 12*N_rb_dl used bins of window w.
int fft_windows_launch(lteb200_ctx *ctx, const float *d_iq, long long n_samps, int fs_idx, int N_rb_dl, const long long *d_win_start, int n_win,
                       float *d_out, cudaStream_t stream)
{
    OfdmParams P;
    P.N = 128 << fs_idx; P.log2N = 7 + fs_idx; P.cp0 = 10 << fs_idx; P.cpe = 9 << fs_idx; P.slot = 960 << fs_idx;
    P.h = 6 * N_rb_dl; P.nsc = 12 * N_rb_dl;
    P.cfo_hz = 0.f; P.cfo_fs = 1.0; P.cfo_base = 0;   // cell search works on samples the caller has already shifted, as in the reference
    float2 *tw = nullptr;
    int rc = get_twiddles(ctx, P.N, &tw);
    if (rc) return rc;
    const int threads = P.N / 4 < 64 ? 64 : (P.N / 4 > 512 ? 512 : P.N / 4);
    ofdm_demod_kernel<float2, true><<<n_win, threads, sizeof(float2) * 2 * P.N, stream>>>((const float2 *)d_iq, n_samps, d_win_start, P, tw, (float2 *)d_out);
    ctx->launches += 1;
    LTE_CUDA_OK(ctx, cudaGetLastError());
    return LTEB200_OK;
}

// ---- standalone front end: int8 -> float de-interleave is implicit (char2 loads); frequency shift of a whole buffer
namespace {
template <typename T>
__global__ void freq_shift_kernel(const T *__restrict__ in, long long n, long long index_base, float f_hz, double fs, float2 *__restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float2 v = load_iq<T>(in, i);
    if (f_hz != 0.f) v = cfo_rotate(v, index_base + i, f_hz, fs);
    out[i] = v;
}
__global__ void freq_shift_split_kernel(float *__restrict__ is, float *__restrict__ qs, long long n, long long index_base, float f_hz, double fs)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float2 v = cfo_rotate(make_float2(is[i], qs[i]), index_base + i, f_hz, fs);
    is[i] = v.x; qs[i] = v.y;
}
}  // namespace

extern "C" int lteb200_freq_shift_batch(lteb200_ctx *ctx, const void *d_iq_in, int iq_fmt, long long n_samps, long long index_base, float freq_offset_hz,
                                        int fs_idx, float *d_iq_out, void *stream)
{
    if (!ctx || !d_iq_in || !d_iq_out || n_samps < 0 || fs_idx < 0 || fs_idx > 4) return LTEB200_ERR_ARGS;
    if (n_samps == 0) return LTEB200_OK;
    LTE_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    const double fs = 1920000.0 * (double)(1 << fs_idx);
    const unsigned blocks = (unsigned)((n_samps + 255) / 256);
    if (iq_fmt == LTEB200_IQ_F32) freq_shift_kernel<float2><<<blocks, 256, 0, (cudaStream_t)stream>>>((const float2 *)d_iq_in, n_samps, index_base, freq_offset_hz, fs, (float2 *)d_iq_out);
    else if (iq_fmt == LTEB200_IQ_I8) freq_shift_kernel<char2><<<blocks, 256, 0, (cudaStream_t)stream>>>((const char2 *)d_iq_in, n_samps, index_base, freq_offset_hz, fs, (float2 *)d_iq_out);
    else return LTEB200_ERR_ARGS;
    ctx->launches += 1;
    LTE_CUDA_OK(ctx, cudaGetLastError());
    return LTEB200_OK;
}

// Host twin of LTE_fdd_dl_fs_samp_buf::freq_shift(start_idx, num_samps, freq_offset) (fs.cc:696-713): in place on the split buffers.
extern "C" int lteb200_freq_shift(lteb200_ctx *ctx, float *i_buf, float *q_buf, uint32_t start_idx, uint32_t num_samps, float freq_offset, int fs_idx)
{
    if (!ctx || !i_buf || !q_buf || fs_idx < 0 || fs_idx > 4) return LTEB200_ERR_ARGS;
    if (num_samps == 0) return LTEB200_OK;
    LTE_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    float *di = nullptr, *dq = nullptr;
    LTE_CUDA_OK(ctx, cudaMalloc(&di, sizeof(float) * num_samps));
    LTE_CUDA_OK(ctx, cudaMalloc(&dq, sizeof(float) * num_samps));
    cudaMemcpy(di, i_buf + start_idx, sizeof(float) * num_samps, cudaMemcpyHostToDevice);
    cudaMemcpy(dq, q_buf + start_idx, sizeof(float) * num_samps, cudaMemcpyHostToDevice);
    freq_shift_split_kernel<<<(num_samps + 255) / 256, 256>>>(di, dq, num_samps, start_idx, freq_offset, 1920000.0 * (double)(1 << fs_idx));
    ctx->launches += 1;
    cudaMemcpy(i_buf + start_idx, di, sizeof(float) * num_samps, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaMemcpy(q_buf + start_idx, dq, sizeof(float) * num_samps, cudaMemcpyDeviceToHost);
    cudaFree(di); cudaFree(dq);
    if (e != cudaSuccess) { lteb200_set_error(ctx, "freq_shift: %s", cudaGetErrorString(e)); return LTEB200_ERR_CUDA; }
    return LTEB200_OK;
}

// Frequency offset applied by the OFDM demodulator while it loads samples (the fused form of the above): every later
// lteb200_get_dl_subframe_and_ce* / lteb200_dl_receive* call on this context rotates sample i of its stream by
// exp(-j 2 pi f (index_base + i + 1) / fs).  0 Hz switches it off.
extern "C" int lteb200_set_freq_offset(lteb200_ctx *ctx, float freq_offset_hz, long long index_base)
{
    if (!ctx) return LTEB200_ERR_ARGS;
    ctx->cfo_hz = freq_offset_hz;
    ctx->cfo_base = index_base;
    return LTEB200_OK;
}

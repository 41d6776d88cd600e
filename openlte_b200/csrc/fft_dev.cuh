// Shared-memory Stockham FFT shared by the OFDM demodulator (cell search path) and the OFDM modulator.
#pragma once
#include <cuda_runtime.h>

__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
    return make_float2(__fsub_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fadd_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}

// Forward DFT of bufA[0..N) (unnormalised, e^{-2 pi i jk/N}), radix 4 with a final radix-2 pass where needed; all threads of the CTA
// take part; returns the buffer that holds the result.  Same pass structure and operation order as oracle/fftw3_shim/fftw3.h, so the
// result is bit-identical to the oracle's FFT.
__device__ __forceinline__ float2 *stockham_forward(float2 *bufA, float2 *bufB, int N, const float2 *__restrict__ tw)
{
    // Stockham autosort, decimation in frequency.  n = remaining transform length, s = stride.
    float2 *src = bufA, *dst = bufB;
    int n = N, s = 1;
    while (n > 1) {
        if ((n & 3) == 0) {
            const int n1 = n >> 2;
            for (int t = threadIdx.x; t < (N >> 2); t += blockDim.x) {
                const int p = t / s, q = t - p * s;   // p in [0,n1), q in [0,s)
                const float2 a = src[q + s * p], b = src[q + s * (p + n1)], c = src[q + s * (p + 2 * n1)], d = src[q + s * (p + 3 * n1)];
                const float2 apc = make_float2(a.x + c.x, a.y + c.y), amc = make_float2(a.x - c.x, a.y - c.y);
                const float2 bpd = make_float2(b.x + d.x, b.y + d.y), bmd = make_float2(b.x - d.x, b.y - d.y);
                const float2 jbmd = make_float2(bmd.y, -bmd.x);  // -j*(b-d) for the forward transform
                const float2 w1 = tw[p * s], w2 = tw[2 * p * s], w3 = tw[3 * p * s];
                dst[q + s * (4 * p + 0)] = make_float2(apc.x + bpd.x, apc.y + bpd.y);
                dst[q + s * (4 * p + 1)] = cmul(make_float2(amc.x + jbmd.x, amc.y + jbmd.y), w1);
                dst[q + s * (4 * p + 2)] = cmul(make_float2(apc.x - bpd.x, apc.y - bpd.y), w2);
                dst[q + s * (4 * p + 3)] = cmul(make_float2(amc.x - jbmd.x, amc.y - jbmd.y), w3);
            }
            n >>= 2; s <<= 2;
        } else {
            const int n1 = n >> 1;
            for (int t = threadIdx.x; t < (N >> 1); t += blockDim.x) {
                const int p = t / s, q = t - p * s;
                const float2 a = src[q + s * p], b = src[q + s * (p + n1)];
                dst[q + s * (2 * p)] = make_float2(a.x + b.x, a.y + b.y);
                dst[q + s * (2 * p + 1)] = cmul(make_float2(a.x - b.x, a.y - b.y), tw[p * s]);
            }
            n >>= 1; s <<= 1;
        }
        __syncthreads();
        float2 *t_ = src; src = dst; dst = t_;
    }
    return src;
}

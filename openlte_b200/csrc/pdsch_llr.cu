// PDSCH symbol level for the max-log-MAP decoder: resource-element extraction, equalisation, demapping, descrambling, rate
// unmatching and quantisation of ONE transport block per CTA, written as the int8 image the pair-lane decoder (turbo_mlm2.cu, MODE_Q8)
// stages with a single bulk copy.  Reference functions restated: the RE walk of liblte_phy_pdsch_channel_decode
// (liblte_phy.cc:3745-3802, index lists built on the host), pre_decoder_and_matched_filter_dl (:7645-7796), layer_demapper_dl (:7473),
// modulation_demapper (:9502-9660), descrambling (:3831-3836), liblte_phy_rate_unmatch_turbo (:11246-11490) and the quantisation at the head
// of turbo_decode (:10636-10664).
//
// Why its own kernel: this phase is a latency-bound gather (16 bytes per resource element from the grid and the estimate, then short
// shared-memory walks) that wants many resident warps, while the decoder is ALU-bound with 168 registers per thread and two CTAs per SM.
// Fused into the decoder (round 1, and the first version of the pair kernel) it left the ALU pipe idle half of the time
// (5.77 ms fused against 4.4 ms split per 16,000 blocks when the split was made); on its own it runs at 4 CTAs per SM, and the decoder's input
// shrinks to 18 KB per block.
//
// The soft bits of a block stay in shared memory as int8 (the reference's demapper emits int8, descrambling multiplies by +-1), the sums of
// repeated bits are exact in integers, and q(v) = trunc(v * 127 / max) -- IEEE multiply and divide as in the reference -- is tabulated once per
// block for 0..max instead of being divided out 18,444 times.
#include "internal.h"
#include "pdsch_dev.cuh"

namespace {

__device__ __forceinline__ int qf(float v, float maxv) { return __float2int_rz(__fdiv_rn(__fmul_rn(v, 127.0f), maxv)); }  // NaN -> 0

// Equalised symbol -> demapped, descrambled soft bits of one resource element.  Not inlined: the demapper (atan2f for QPSK) would otherwise be
// replicated at every call site.
__device__ __noinline__ void emit_soft(float xr, float xi, int mod, int Qm, const uint8_t *__restrict__ cs_r, int8_t *soft_r)
{
    int bb[6];
    demap_bits(xr, xi, mod, bb);
#pragma unroll
    for (int q = 0; q < 6; q++)
        if (q < Qm) soft_r[q] = (int8_t)(cs_r[q] ? -bb[q] : bb[q]);
}

// soft path: equalised symbol + post-equalisation SNR -> max-log soft bits, descrambled
__device__ __noinline__ void emit_maxlog(float xr, float xi, float gain, int mod, int Qm, const uint8_t *__restrict__ cs_r, int8_t *soft_r)
{
    int bb[6];
    demap_maxlog(xr, xi, gain, mod, bb);
#pragma unroll
    for (int q = 0; q < 6; q++)
        if (q < Qm) soft_r[q] = (int8_t)(cs_r[q] ? -bb[q] : bb[q]);
}

// The reference's 16QAM / 64QAM demapper makes hard decisions (+-127, liblte_phy.cc:9576, :9607): the soft bits of a resource element are a
// function of Qm decision bits XOR Qm scrambling bits.  hard_bits() evaluates the very comparisons of demap_bits() (NaN behaves the same) and
// returns bit q = 1 where demap_bits() gives -127.
__device__ __forceinline__ uint32_t hard_bits(float xr, float xi, int mod)
{
    uint32_t h = (xr > 0.f ? 0u : 1u) | (xi > 0.f ? 0u : 2u);
    const float ar = fabsf(xr), ai = fabsf(xi);
    if (mod == LTEB200_MOD_16QAM) {
        const float t = 0.63245553203367586640f;
        h |= (ar < t ? 0u : 4u) | (ai < t ? 0u : 8u);
    } else {
        const float t2 = 0.30860669992418382490f, t4 = 0.61721339984836764980f, t6 = 0.92582009977255147470f;
        if (ar < t4) h |= (ar > t2 ? 0u : 16u); else h |= 4u | (ar < t6 ? 0u : 16u);
        if (ai < t4) h |= (ai > t2 ? 0u : 32u); else h |= 8u | (ai < t6 ? 0u : 32u);
    }
    return h;
}

constexpr int LUT_MAX = 4096;     // table of q over -max..max (127 x 16 repeats each way); beyond it the division is done per value

// first_kT: gather map in OUTPUT order: [x][pos] for the three streams (pos = (k mod 16) * NT + k / 16, -1 on the padding) followed by the 12
// tail entries d[3K + j].  q8 image per block: Ls[NP] | Lp1[NP] | Lp2[NP] | tail[16].
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 4)
pdsch_llr_q8_kernel(const __grid_constant__ FusedParams fp, int K, int NT, int soft_bytes, int8_t *__restrict__ q8)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_red[32];
    __shared__ int8_t lut[LUT_MAX];
    const int cb = blockIdx.x, t = threadIdx.x;
    const int a_loc = cb / fp.n_sf, sf = cb - a_loc * fp.n_sf;
    const FusedAlloc &A = fp.al[a_loc];
    const int sn = fp.subfr_num[sf];
    const int cls = sn == 0 ? 0 : (sn == 5 ? 1 : 2);
    const int M = A.M[cls], Qm = A.Qm, mod = A.mod, E = M * Qm;
    int8_t *soft = reinterpret_cast<int8_t *>(smem_raw);
    // the image is assembled in shared memory (scattered byte stores) and leaves with 16-byte stores; positions nobody feeds stay 0 = q(0)
    const int NP = 16 * NT, n_map = 3 * NP + 12, img_bytes = 3 * NP + 16;
    int8_t *stage = soft + soft_bytes;
    for (int p = t; p < img_bytes / 16; p += THREADS) reinterpret_cast<uint4 *>(stage)[p] = make_uint4(0, 0, 0, 0);
    const float2 *gr = fp.grid + (size_t)sf * 16 * fp.nsc, *ch = fp.ce + (size_t)sf * fp.N_ant * 14 * fp.nsc;
    const int *ix = A.re_idx + cls * A.idx_stride;
    const uint8_t *cs = A.cseq + (size_t)sn * A.c_stride;
    const bool maxlog = fp.demap == LTEB200_DEMAP_MAXLOG;
    __shared__ float s_s2;
    if (maxlog) {
        // noise variance from the pilot tones of symbol 0, port 0 (see pdsch_dev.cuh); one thread, fixed summation order (the CPU twin adds in
        // the same order, oracle/port/softdemap.cc)
        if (t == 0) {
            const int n = fp.nsc / 6;
            float acc = 0.f, pw = 0.f;
            for (int j = 1; j + 1 < n; j++) {
                const float2 c = ch[6 * j + fp.vshift], a = ch[6 * (j - 1) + fp.vshift], b = ch[6 * (j + 1) + fp.vshift];
                const float dr = __fsub_rn(c.x, __fmul_rn(0.5f, __fadd_rn(a.x, b.x))), di = __fsub_rn(c.y, __fmul_rn(0.5f, __fadd_rn(a.y, b.y)));
                acc = __fadd_rn(acc, __fadd_rn(__fmul_rn(dr, dr), __fmul_rn(di, di)));
                pw = __fadd_rn(pw, __fadd_rn(__fmul_rn(c.x, c.x), __fmul_rn(c.y, c.y)));
            }
            const float cnt = (float)(n - 2);
            float s2 = __fdiv_rn(__fmul_rn(0.6666666865348816f, acc), cnt);
            const float floor_ = __fmul_rn(__fdiv_rn(pw, cnt), 1e-6f);      // at most 60 dB
            if (!(s2 > floor_)) s2 = floor_;
            if (!(s2 > 0.f)) s2 = 1e-12f;
            s_s2 = s2;
        }
        __syncthreads();
    }
    const float s2 = maxlog ? s_s2 : 1.0f;
    // one port, reference demapper, 16QAM / 64QAM: decision bits XOR packed scrambling bits -> one table read gives the element's soft bytes
    const bool hard_fast = fp.N_ant == 1 && !maxlog && (mod == LTEB200_MOD_16QAM || mod == LTEB200_MOD_64QAM);
    __shared__ uint2 s_hard[64];
    if (hard_fast) {
        if (t < 64) {
            uint32_t lo = 0, hi = 0;
            for (int q = 0; q < 4; q++) lo |= (((t >> q) & 1) ? 0x81u : 0x7fu) << (8 * q);
            for (int q = 4; q < 6; q++) hi |= (((t >> q) & 1) ? 0x81u : 0x7fu) << (8 * (q - 4));
            s_hard[t] = make_uint2(lo, hi);
        }
        __syncthreads();
        const uint8_t *cp = A.cseqP + (size_t)sn * A.idx_stride;
        constexpr int RE_IT = 4;
        for (int r0 = t; r0 < M; r0 += RE_IT * THREADS) {
            int idx[RE_IT];
            float2 yv[RE_IT], hv[RE_IT];
            uint32_t cv[RE_IT];
#pragma unroll
            for (int u = 0; u < RE_IT; u++) idx[u] = (r0 + u * THREADS < M) ? ix[r0 + u * THREADS] : ix[r0];
#pragma unroll
            for (int u = 0; u < RE_IT; u++) { yv[u] = gr[idx[u]]; hv[u] = ch[idx[u]]; cv[u] = (r0 + u * THREADS < M) ? cp[r0 + u * THREADS] : 0u; }
#pragma unroll
            for (int u = 0; u < RE_IT; u++) {
                const int r = r0 + u * THREADS;
                if (r < M) {
                    float xr, xi;
                    zf_1port(yv[u], hv[u], xr, xi);
                    const uint2 w = s_hard[hard_bits(xr, xi, mod) ^ cv[u]];
                    if (Qm == 4) *reinterpret_cast<uint32_t *>(soft + 4 * r) = w.x;
                    else {
                        uint16_t *o = reinterpret_cast<uint16_t *>(soft + 6 * r);
                        o[0] = (uint16_t)w.x; o[1] = (uint16_t)(w.x >> 16); o[2] = (uint16_t)w.y;
                    }
                }
            }
        }
    } else if (fp.N_ant != 1) {
        // transmit diversity: one thread per group of N_ant consecutive resource elements (M is a multiple of N_ant here, the launcher checks)
        const size_t ps = (size_t)14 * fp.nsc;
        auto pw2 = [](float2 h) { return __fadd_rn(__fmul_rn(h.x, h.x), __fmul_rn(h.y, h.y)); };
        for (int r = t * fp.N_ant; r < M; r += THREADS * fp.N_ant) {
            if (fp.N_ant == 2) {
                float2 x0, x1;
                const float2 h0 = ch[ix[r]], h1 = ch[ps + ix[r]];
                sfbc2(gr[ix[r]], gr[ix[r + 1]], h0, h1, x0, x1);
                if (maxlog) {
                    // the reference divides by sqrt(|h0|^4 + |h1|^4) (sic, liblte_phy.cc:7700); maximum-ratio combining divides by |h0|^2 + |h1|^2
                    const float a0 = pw2(h0), a1 = pw2(h1), den = __fadd_rn(a0, a1);
                    const float rs = __fdiv_rn(__fsqrt_rn(__fadd_rn(__fmul_rn(a0, a0), __fmul_rn(a1, a1))), den), g = __fdiv_rn(den, s2);
                    emit_maxlog(__fmul_rn(x0.x, rs), __fmul_rn(x0.y, rs), g, mod, Qm, cs + r * Qm, soft + r * Qm);
                    emit_maxlog(__fmul_rn(x1.x, rs), __fmul_rn(x1.y, rs), g, mod, Qm, cs + (r + 1) * Qm, soft + (r + 1) * Qm);
                } else {
                    emit_soft(x0.x, x0.y, mod, Qm, cs + r * Qm, soft + r * Qm);
                    emit_soft(x1.x, x1.y, mod, Qm, cs + (r + 1) * Qm, soft + (r + 1) * Qm);
                }
            } else {
                const float2 y[4] = {gr[ix[r]], gr[ix[r + 1]], gr[ix[r + 2]], gr[ix[r + 3]]};
                const float2 h0 = ch[ix[r]], h1 = ch[ps + ix[r + 2]], h2 = ch[2 * ps + ix[r]], h3 = ch[3 * ps + ix[r + 2]];
                float2 x[4];
                sfbc4(y, h0, h1, h2, h3, x);
                if (maxlog) {
                    const float a0 = pw2(h0), a1 = pw2(h1), a2 = pw2(h2), a3 = pw2(h3), d02 = __fadd_rn(a0, a2), d13 = __fadd_rn(a1, a3);
                    const float r02 = __fdiv_rn(__fsqrt_rn(__fadd_rn(__fmul_rn(a0, a0), __fmul_rn(a2, a2))), d02), g02 = __fdiv_rn(d02, s2);
                    const float r13 = __fdiv_rn(__fsqrt_rn(__fadd_rn(__fmul_rn(a1, a1), __fmul_rn(a3, a3))), d13), g13 = __fdiv_rn(d13, s2);
                    for (int q = 0; q < 4; q++) {
                        const float rs = q < 2 ? r02 : r13, g = q < 2 ? g02 : g13;
                        emit_maxlog(__fmul_rn(x[q].x, rs), __fmul_rn(x[q].y, rs), g, mod, Qm, cs + (r + q) * Qm, soft + (r + q) * Qm);
                    }
                } else
                    for (int q = 0; q < 4; q++) emit_soft(x[q].x, x[q].y, mod, Qm, cs + (r + q) * Qm, soft + (r + q) * Qm);
            }
        }
    } else {
        // indices, grid and estimate values of RE_IT elements are fetched before the arithmetic starts: the gather latency is paid once per round
        constexpr int RE_IT = 4;
        for (int r0 = t; r0 < M; r0 += RE_IT * THREADS) {
            int idx[RE_IT];
            float2 yv[RE_IT], hv[RE_IT];
#pragma unroll
            for (int u = 0; u < RE_IT; u++) idx[u] = (r0 + u * THREADS < M) ? ix[r0 + u * THREADS] : ix[r0];
#pragma unroll
            for (int u = 0; u < RE_IT; u++) { yv[u] = gr[idx[u]]; hv[u] = ch[idx[u]]; }
#pragma unroll
            for (int u = 0; u < RE_IT; u++) {
                const int r = r0 + u * THREADS;
                if (r < M) {
                    float xr, xi;
                    zf_1port(yv[u], hv[u], xr, xi);
                    if (maxlog) {
                        const float hh = __fadd_rn(__fmul_rn(hv[u].x, hv[u].x), __fmul_rn(hv[u].y, hv[u].y));
                        emit_maxlog(xr, xi, __fdiv_rn(hh, s2), mod, Qm, cs + r * Qm, soft + r * Qm);
                    } else
                        emit_soft(xr, xi, mod, Qm, cs + r * Qm, soft + r * Qm);
                }
            }
        }
    }
    __syncthreads();
    // rate unmatching (liblte_phy.cc:11403-11415): the first visit assigns, repeats are added in visiting order.  The reference works in float
    // with the sentinel 10000 for "nothing received"; the soft bits are small integers, so integer sums are exact and the sentinel
    // comparisons are restated on the integers (a running sum that hits 10000 exactly is replaced, as there).
    constexpr int NUL = 10000;
    const int period = A.period;
    const bool two_visits = E <= 2 * period;          // every output is visited at most twice: v = x0 + x1, no sentinel can arise
    auto unmatch = [&](int fk) -> int {
        if (two_visits) {
            const int k1 = fk + period;
            const int x0 = (fk >= 0 && fk < E) ? soft[fk] : 0, x1 = (fk >= 0 && k1 < E) ? soft[k1] : 0;
            return x0 + x1;
        }
        int w = NUL;
        if (fk >= 0)
#pragma unroll 1
            for (int k = fk; k < E; k += period) {
                const int x = soft[k];
                w = (w == NUL) ? x : w + x;
            }
        return w == NUL ? 0 : w;                       // turbo_decode zeroes the sentinels (liblte_phy.cc:10636-10642)
    };
    // ---- pass 1: max |v| over the 3 K + 12 outputs (padding entries are -1 -> 0)
    const int *fkT = A.first_kT;
    const int n2 = E > period ? E - period : 0;        // e-bits [0, n2) are visited a second time at [period, period + n2)
    // sums v[k .. k+3] of four consecutive e-bits of the first lap with their second-lap partners, as int16 lanes (v0 = elements 0, 1; v1 = 2, 3):
    // bytes -> int16 by sign-replicating byte permutes, packed adds.  Needs 4 | k, 4 | period and k + 3 < min(E, period).
    const bool vec_ok = two_visits && (period & 3) == 0;
    const int n1v = (E < period ? E : period) & ~3;
    auto sum4 = [&](int k, uint32_t &v0, uint32_t &v1) {
        const uint32_t a = *reinterpret_cast<const uint32_t *>(soft + k);
        uint32_t b = 0;
        if (k < n2) {
            b = *reinterpret_cast<const uint32_t *>(soft + k + period);
            if (k + 4 > n2) b &= (1u << (8 * (n2 - k))) - 1u;                       // the second lap ends inside this word
        }
        uint32_t a0, a1, b0, b1;
        asm("prmt.b32 %0, %1, %1, 0x9180;" : "=r"(a0) : "r"(a));
        asm("prmt.b32 %0, %1, %1, 0xb3a2;" : "=r"(a1) : "r"(a));
        asm("prmt.b32 %0, %1, %1, 0x9180;" : "=r"(b0) : "r"(b));
        asm("prmt.b32 %0, %1, %1, 0xb3a2;" : "=r"(b1) : "r"(b));
        v0 = __vadd2(a0, b0); v1 = __vadd2(a1, b1);
    };
    // The reference-exact decoder quantises against the maximum over the 3 K code-block values and never looks at the 12 tail values
    // (liblte_phy.cc:10645-10664): for it the e-bits that feed the tail outputs are cleared (twelve threads, one tail output each), so the
    // maximum is the reference's and the image's tail bytes are 0.
    if (fp.max_no_tail && two_visits) {
        if (t < 12) {
            const int tail_k = fkT[3 * NP + t];
            if (tail_k >= 0) {
                if (tail_k < E) soft[tail_k] = 0;
                if (tail_k + period < E) soft[tail_k + period] = 0;
            }
        }
        __syncthreads();
    }
    int mi = 0;
    if (two_visits) {
        // the first-lap e-bits [0, period) are in one-to-one correspondence with the outputs that receive anything, so the maximum over
        // the outputs is a maximum over e-bits: no index loads
        const int n1 = E < period ? E : period;
        if (vec_ok) {
            uint32_t m2 = 0;                                                        // two int16 lanes of running max |v|
            for (int k = 4 * t; k < n1v; k += 4 * THREADS) {
                uint32_t v0, v1;
                sum4(k, v0, v1);
                m2 = __vmaxs2(m2, __vmaxs2(__vmaxs2(v0, __vsub2(0u, v0)), __vmaxs2(v1, __vsub2(0u, v1))));
            }
            mi = max((int)(m2 & 0xffffu), (int)(m2 >> 16));
            for (int k = n1v + t; k < n1; k += THREADS) mi = max(mi, abs((int)soft[k] + (k < n2 ? (int)soft[k + period] : 0)));
        } else
            for (int k = t; k < n1; k += THREADS) mi = max(mi, abs((int)soft[k] + (k < n2 ? (int)soft[k + period] : 0)));
    } else {
        for (int j = t; j < (fp.max_no_tail ? 3 * NP : n_map); j += THREADS) mi = max(mi, abs(unmatch(fkT[j])));
    }
    for (int o = 16; o > 0; o >>= 1) mi = max(mi, __shfl_xor_sync(0xffffffffu, mi, o));
    if ((t & 31) == 0) s_red[t >> 5] = mi;
    __syncthreads();
    if (t < 32) {
        int v = t < THREADS / 32 ? s_red[t] : 0;
        for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (t == 0) s_red[31] = v;                           // (its own slot: lanes of this warp have just read s_red[0..7])
    }
    __syncthreads();
    const int maxi = s_red[31];
    const float maxv = (float)maxi;
    // q(v) tabulated over -maxi..maxi (index v + maxi)
    const bool use_lut = 2 * maxi < LUT_MAX;
    if (use_lut) for (int p = t; p <= 2 * maxi; p += THREADS) lut[p] = (int8_t)qf((float)(p - maxi), maxv);
    __syncthreads();
    // ---- pass 2: quantised image, four consecutive output bytes per thread and round
    int8_t *img = q8 + (size_t)cb * (3 * NP + 16);
    const int4 *fk4 = reinterpret_cast<const int4 *>(fkT);
    if (vec_ok && use_lut) {
        // e-bit driven: four consecutive first-lap e-bits per step -> sums -> table -> the four image positions they feed (inverse gather map)
        const int8_t *lutc = lut + maxi;
        const uint16_t *inv = A.inv_kT;
        const int n1 = E < period ? E : period;
        for (int k = 4 * t; k < n1v; k += 4 * THREADS) {
            uint32_t v0, v1;
            sum4(k, v0, v1);
            const uint2 iv = *reinterpret_cast<const uint2 *>(inv + k);
            stage[iv.x & 0xffffu] = lutc[(int)(int16_t)(v0 & 0xffffu)];
            stage[iv.x >> 16] = lutc[(int)v0 >> 16];
            stage[iv.y & 0xffffu] = lutc[(int)(int16_t)(v1 & 0xffffu)];
            stage[iv.y >> 16] = lutc[(int)v1 >> 16];
        }
        for (int k = n1v + t; k < n1; k += THREADS) stage[inv[k]] = lutc[(int)soft[k] + (k < n2 ? (int)soft[k + period] : 0)];
        __syncthreads();
        for (int p = t; p < img_bytes / 16; p += THREADS) reinterpret_cast<uint4 *>(img)[p] = reinterpret_cast<const uint4 *>(stage)[p];
        return;
    }
    if (two_visits && use_lut) {
        const int8_t *lutc = lut + maxi;
        auto q2 = [&](int fk) -> uint32_t {
            const int x0 = (unsigned)fk < (unsigned)E ? (int)soft[fk] : 0, x1 = (unsigned)fk < (unsigned)n2 ? (int)soft[fk + period] : 0;
            return (uint32_t)(uint8_t)lutc[x0 + x1];
        };
        for (int j = t; j < 3 * NP / 4; j += THREADS) {
            const int4 f = fk4[j];
            reinterpret_cast<uint32_t *>(img)[j] = q2(f.x) | (q2(f.y) << 8) | (q2(f.z) << 16) | (q2(f.w) << 24);
        }
        if (t < 16) img[3 * NP + t] = (t < 12 && !fp.max_no_tail) ? (int8_t)q2(fkT[3 * NP + t]) : (int8_t)0;
        return;
    }
    auto q_int = [&](int v) -> uint32_t {
        const int q = use_lut ? (int)lut[v + maxi] : qf((float)v, maxv);
        return (uint32_t)q & 0xffu;
    };
    for (int j = t; j < 3 * NP / 4; j += THREADS) {
        const int4 f = fk4[j];
        const uint32_t w = q_int(unmatch(f.x)) | (q_int(unmatch(f.y)) << 8) | (q_int(unmatch(f.z)) << 16) | (q_int(unmatch(f.w)) << 24);
        reinterpret_cast<uint32_t *>(img)[j] = w;
    }
    if (t < 16) img[3 * NP + t] = (t < 12 && !fp.max_no_tail) ? (int8_t)q_int(unmatch(fkT[3 * NP + t])) : (int8_t)0;
}

}  // namespace

// One CTA per transport block of the group described by fp (fp.n_al allocations of the same K, fp.n_sf subframes); q8 receives n_al * n_sf images.
int pdsch_llr_q8_launch(lteb200_ctx *ctx, const FusedParams &fp, int K, int E_max, int8_t *d_q8, cudaStream_t stream)
{
    constexpr int THREADS = 256;
    const int NT = mlm_size_class(K);
    const size_t soft_bytes = ((size_t)E_max + 15) & ~(size_t)15;
    const size_t smem = soft_bytes + 3 * 16 * (size_t)NT + 16;
    if (smem > 200 * 1024) { lteb200_set_error(ctx, "PDSCH: %d soft bits per transport block exceed the shared-memory staging", E_max); return LTEB200_ERR_UNSUPPORTED; }
    LTE_CUDA_OK(ctx, cudaFuncSetAttribute(pdsch_llr_q8_kernel<THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pdsch_llr_q8_kernel<THREADS><<<fp.n_al * fp.n_sf, THREADS, smem, stream>>>(fp, K, NT, (int)soft_bytes, d_q8);
    ctx->launches += 1;
    LTE_CUDA_OK(ctx, cudaGetLastError());
    return LTEB200_OK;
}

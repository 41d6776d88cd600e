// Device functions shared by the staged PDSCH kernels (pdsch.cu) and the fused receive kernel (turbo_mlm.cu).
// Twins of pre_decoder_and_matched_filter_dl (1 port, liblte_phy.cc:7680-7690), modulation_demapper (:9502-9660) and
// get_soft_decision (:13880-13900), operation by operation (explicit IEEE intrinsics: no contraction).
#pragma once
#include "../../include/lte_b200.h"

__device__ __forceinline__ float soft_dist(float re, float im, float ere, float eim)
{
    // get_soft_decision(rx, exp, max_dist = 1): 1 - min(dist, 1 - 1/120)
    const float dr = __fsub_rn(re, ere), di = __fsub_rn(im, eim);
    float dist = __fsqrt_rn(__fadd_rn(__fmul_rn(dr, dr), __fmul_rn(di, di)));
    const float lim = __fsub_rn(1.0f, __fdiv_rn(1.0f, 120.0f));
    if (dist >= lim) dist = lim;
    return __fsub_rn(1.0f, dist);
}

// soft bits (int8 range) of one equalised symbol; positive = bit 0
__device__ __forceinline__ void demap_bits(float xr, float xi, int mod, int b[6])
{
    const float r2 = 0.70710678118654752440f;
    if (mod == LTEB200_MOD_QPSK) {
        const float ang = atan2f(xi, xr);
        const double a = (double)ang, hp = 1.57079632679489661923, pi = 3.14159265358979323846;
        int s0, s1;
        if (a >= 0 && a < hp) { s0 = 1; s1 = 1; }
        else if (a >= -hp && a < 0) { s0 = 1; s1 = -1; }
        else if (a >= hp && a < pi) { s0 = -1; s1 = 1; }
        else { s0 = -1; s1 = -1; }
        const int m = __float2int_rz(__fmul_rn(127.0f, soft_dist(xr, xi, s0 > 0 ? r2 : -r2, s1 > 0 ? r2 : -r2)));
        b[0] = s0 * m; b[1] = s1 * m;
    } else if (mod == LTEB200_MOD_16QAM) {
        const float t = 0.63245553203367586640f;  // 2/sqrt(10)
        b[0] = xr > 0 ? 127 : -127; b[1] = xi > 0 ? 127 : -127;
        b[2] = fabsf(xr) < t ? 127 : -127; b[3] = fabsf(xi) < t ? 127 : -127;
    } else if (mod == LTEB200_MOD_64QAM) {
        const float t2 = 0.30860669992418382490f, t4 = 0.61721339984836764980f, t6 = 0.92582009977255147470f;  // {2,4,6}/sqrt(42)
        const float ar = fabsf(xr), ai = fabsf(xi);
        b[0] = xr > 0 ? 127 : -127; b[1] = xi > 0 ? 127 : -127;
        if (ar < t4) { b[2] = 127; b[4] = ar > t2 ? 127 : -127; } else { b[2] = -127; b[4] = ar < t6 ? 127 : -127; }
        if (ai < t4) { b[3] = 127; b[5] = ai > t2 ? 127 : -127; } else { b[3] = -127; b[5] = ai < t6 ? 127 : -127; }
    } else {  // BPSK
        const double a = (double)atan2f(xi, xr), pi = 3.14159265358979323846;
        if (a > -pi / 4 && a < 3 * pi / 4) b[0] = __float2int_rz(__fmul_rn(127.0f, soft_dist(xr, xi, r2, r2)));
        else b[0] = -__float2int_rz(__fmul_rn(127.0f, soft_dist(xr, xi, -r2, -r2)));
    }
}

// zero forcing x = y * conj(h) / |h|^2 in the operation order of liblte_phy.cc:7684-7689
__device__ __forceinline__ void zf_1port(float2 y, float2 h, float &xr, float &xi)
{
    const float nrm = __fadd_rn(__fmul_rn(h.x, h.x), __fmul_rn(h.y, h.y));
    xr = __fdiv_rn(__fadd_rn(__fmul_rn(y.x, h.x), __fmul_rn(y.y, h.y)), nrm);
    xi = __fdiv_rn(__fsub_rn(__fmul_rn(y.y, h.x), __fmul_rn(y.x, h.y)), nrm);
}

// ---- soft path (BASELINE north star: "per-RE MMSE equalisation, soft QPSK/16QAM/64QAM demapping"; the reference equalises by zero forcing and
// makes hard 16/64QAM decisions with a FIXME, liblte_phy.cc:7680-7690, :9576, :9607).  Off by default -- the reference-exact path is the parity
// path -- and selected per context with lteb200_set_demapper(LTEB200_DEMAP_MAXLOG); it feeds the max-log-MAP decoder only.
//
// One antenna port: x_mmse = conj(h) y / (|h|^2 + s2) is biased by |h|^2 / (|h|^2 + s2); unbiased it is the zero-forcing symbol with noise
// variance s2 / |h|^2, so the exact max-log LLR of a Gray-mapped bit is  g * (min_{s: b=1} |x - s|^2 - min_{s: b=0} |x - s|^2), g = |h|^2 / s2
// (positive = bit 0, the reference's sign convention).  In units of the constellation's half spacing A (x = A u) the two minima are the
// piecewise-linear functions below.  The int8 soft bit is round(2 * LLR) clamped to +-127.
// s2 comes from the CRS: on symbol 0 the estimate at a pilot tone IS the least-squares estimate, so d_j = h_j - (h_{j-1} + h_{j+1}) / 2 over
// neighbouring pilot tones has variance 1.5 s2 on a channel that is smooth over 12 subcarriers.
__device__ __forceinline__ void maxlog_dim(float u, int Qm, float w[3])
{
    // u: one dimension (I or Q) in half-spacing units; w[0] = sign bit, w[1], w[2] = amplitude bits (as far as the constellation has them)
    const float v = fabsf(u);
    if (Qm == 2) { w[0] = __fmul_rn(4.0f, u); return; }
    if (Qm == 4) {
        w[0] = v <= 2.0f ? __fmul_rn(4.0f, u) : (u > 0.f ? __fsub_rn(__fmul_rn(8.0f, u), 8.0f) : __fadd_rn(__fmul_rn(8.0f, u), 8.0f));
        w[1] = __fmul_rn(4.0f, __fsub_rn(2.0f, v));
        return;
    }
    float a, c;        // sign bit: a u -+ c
    if (v <= 2.0f) { a = 4.0f; c = 0.0f; } else if (v <= 4.0f) { a = 8.0f; c = 8.0f; } else if (v <= 6.0f) { a = 12.0f; c = 24.0f; } else { a = 16.0f; c = 48.0f; }
    w[0] = u > 0.f ? __fsub_rn(__fmul_rn(a, u), c) : __fadd_rn(__fmul_rn(a, u), c);
    w[1] = v <= 2.0f ? __fsub_rn(24.0f, __fmul_rn(8.0f, v)) : (v <= 6.0f ? __fsub_rn(16.0f, __fmul_rn(4.0f, v)) : __fsub_rn(40.0f, __fmul_rn(8.0f, v)));
    w[2] = v <= 4.0f ? __fsub_rn(__fmul_rn(4.0f, v), 8.0f) : __fsub_rn(24.0f, __fmul_rn(4.0f, v));
}
// (xr, xi): unbiased equalised symbol; gain = post-equalisation SNR |h|^2 / s2; b[Qm] int8-range soft bits in the reference's bit order
__device__ __forceinline__ void demap_maxlog(float xr, float xi, float gain, int mod, int b[6])
{
    const int Qm = mod == LTEB200_MOD_QPSK ? 2 : (mod == LTEB200_MOD_16QAM ? 4 : (mod == LTEB200_MOD_64QAM ? 6 : 1));
    if (Qm == 1) {      // BPSK: points +-(1 + j)/sqrt(2): project on the diagonal
        const float t = __fmul_rn(__fmul_rn(2.0f, gain), __fmul_rn(2.8284271247461903f, __fmul_rn(0.70710678118654752440f, __fadd_rn(xr, xi))));
        b[0] = max(-127, min(127, __float2int_rn(t)));
        return;
    }
    const float inv_a = Qm == 2 ? 1.4142135623730951f : (Qm == 4 ? 3.1622776601683795f : 6.4807406984078604f);     // 1 / A
    const float a2 = Qm == 2 ? 0.5f : (Qm == 4 ? 0.1f : 0.023809523809523808f);                                      // A^2
    const float k = __fmul_rn(__fmul_rn(2.0f, gain), a2);
    float wi[3], wq[3];
    maxlog_dim(__fmul_rn(xr, inv_a), Qm, wi);
    maxlog_dim(__fmul_rn(xi, inv_a), Qm, wq);
#pragma unroll
    for (int t = 0; t < 3; t++)
        if (2 * t < Qm) {
            b[2 * t] = max(-127, min(127, __float2int_rn(__fmul_rn(k, wi[t]))));
            b[2 * t + 1] = max(-127, min(127, __float2int_rn(__fmul_rn(k, wq[t]))));
        }
}

// ---- transmit-diversity combining shared by PDSCH, PBCH and PDCCH (pre_decoder_and_matched_filter_dl + layer_demapper_dl,
// liblte_phy.cc:7691-7765, 7505-7512), in the reference's operation order (IEEE adds/multiplies, no contraction)
__device__ __forceinline__ float td_sum4(float a, float b, float c, float d) { return __fadd_rn(__fadd_rn(__fadd_rn(a, b), c), d); }
#define TD_MUL(a, b) __fmul_rn((a), (b))
// two ports: symbols y0,y1 of one pair, both estimates taken at the EVEN element (:7696-7699)
__device__ __forceinline__ void sfbc2(float2 y0, float2 y1, float2 h0, float2 h1, float2 &x0, float2 &x1)
{
    const float a0 = __fadd_rn(TD_MUL(h0.x, h0.x), TD_MUL(h0.y, h0.y)), a1 = __fadd_rn(TD_MUL(h1.x, h1.x), TD_MUL(h1.y, h1.y));
    const float n = __fsqrt_rn(__fadd_rn(TD_MUL(a0, a0), TD_MUL(a1, a1)));                        // sic: sqrt(|h0|^4 + |h1|^4)
    x0.x = __fdiv_rn(td_sum4(TD_MUL(h0.x, y0.x), TD_MUL(h0.y, y0.y), TD_MUL(h1.x, y1.x), TD_MUL(h1.y, y1.y)), n);
    x0.y = __fdiv_rn(td_sum4(TD_MUL(h0.x, y0.y), -TD_MUL(h0.y, y0.x), -TD_MUL(h1.x, y1.y), TD_MUL(h1.y, y1.x)), n);
    x1.x = __fdiv_rn(td_sum4(-TD_MUL(h1.x, y0.x), -TD_MUL(h1.y, y0.y), TD_MUL(h0.x, y1.x), TD_MUL(h0.y, y1.y)), n);
    x1.y = __fdiv_rn(td_sum4(TD_MUL(h1.x, y0.y), -TD_MUL(h1.y, y0.x), TD_MUL(h0.x, y1.y), -TD_MUL(h0.y, y1.x)), n);
}
// four ports: h0,h2 at element 0 of the group, h1,h3 at element 2 (:7723-7730)
__device__ __forceinline__ void sfbc4(const float2 y[4], float2 h0, float2 h1, float2 h2, float2 h3, float2 x[4])
{
    const float2 y0 = y[0], y1 = y[1], y2 = y[2], y3 = y[3];
    const float a0 = __fadd_rn(TD_MUL(h0.x, h0.x), TD_MUL(h0.y, h0.y)), a1 = __fadd_rn(TD_MUL(h1.x, h1.x), TD_MUL(h1.y, h1.y));
    const float a2 = __fadd_rn(TD_MUL(h2.x, h2.x), TD_MUL(h2.y, h2.y)), a3 = __fadd_rn(TD_MUL(h3.x, h3.x), TD_MUL(h3.y, h3.y));
    const float n02 = __fsqrt_rn(__fadd_rn(TD_MUL(a0, a0), TD_MUL(a2, a2))), n13 = __fsqrt_rn(__fadd_rn(TD_MUL(a1, a1), TD_MUL(a3, a3)));
    x[0].x = __fdiv_rn(td_sum4(TD_MUL(h0.x, y0.x), TD_MUL(h0.y, y0.y), TD_MUL(h2.x, y1.x), TD_MUL(h2.y, y1.y)), n02);
    x[0].y = __fdiv_rn(td_sum4(TD_MUL(h0.x, y0.y), -TD_MUL(h0.y, y0.x), -TD_MUL(h2.x, y1.y), TD_MUL(h2.y, y1.x)), n02);
    x[1].x = __fdiv_rn(td_sum4(-TD_MUL(h2.x, y0.x), -TD_MUL(h2.y, y0.y), TD_MUL(h0.x, y1.x), TD_MUL(h0.y, y1.y)), n02);
    x[1].y = __fdiv_rn(-td_sum4(-TD_MUL(h2.x, y0.y), TD_MUL(h2.y, y0.x), -TD_MUL(h0.x, y1.y), TD_MUL(h0.y, y1.x)), n02);
    x[2].x = __fdiv_rn(td_sum4(TD_MUL(h1.x, y2.x), TD_MUL(h1.y, y2.y), TD_MUL(h3.x, y3.x), TD_MUL(h3.y, y3.y)), n13);
    x[2].y = __fdiv_rn(td_sum4(TD_MUL(h1.x, y2.y), -TD_MUL(h1.y, y2.x), -TD_MUL(h3.x, y3.y), TD_MUL(h3.y, y3.x)), n13);
    x[3].x = __fdiv_rn(td_sum4(-TD_MUL(h3.x, y2.x), -TD_MUL(h3.y, y2.y), TD_MUL(h1.x, y3.x), TD_MUL(h1.y, y3.y)), n13);
    x[3].y = __fdiv_rn(-td_sum4(-TD_MUL(h3.x, y2.y), TD_MUL(h3.y, y2.x), -TD_MUL(h1.x, y3.y), TD_MUL(h1.y, y3.x)), n13);
}


// What the fused receive kernel needs to know about one allocation (device pointers into the context's cached tables).
struct FusedAlloc {
    const int *re_idx;        // [3][idx_stride]  PDSCH resource elements per subframe class (0, 5, other)
    const uint8_t *cseq;      // [10][c_stride]   scrambling bits per subframe number
    const uint8_t *cseqP;     // [10][idx_stride] the same, the Qm bits of a resource element packed into one byte
    const int32_t *first_k;   // [3(K+4)]         rate-unmatch gather map
    const int32_t *first_kT;  // the same in the decoder's window-transposed output order + 12 tail entries (pair kernel only)
    const uint16_t *inv_kT;   // [period] image position fed by e-bit k of the first lap (inverse of first_kT)
    int idx_stride, c_stride;
    int M[3];
    int Qm, mod, period, F, tbs, a_idx;
};
struct FusedParams {
    const float2 *grid, *ce;
    const int *subfr_num;
    const uint32_t *crc_tab;  // x^j mod g(CRC24A), j < 6168
    uint8_t *tb_out;
    int32_t *crc_ok;
    int nsc, n_sf, n_alloc_total, tb_stride, n_al;
    int N_ant;                // 1: zero forcing; 2 / 4: transmit-diversity combining of groups of N_ant resource elements
    int demap;                // LTEB200_DEMAP_*
    int tb_packed;            // transport-block output: 0 one bit per byte, 1 eight bits per byte (MSB first)
    int max_no_tail;          // quantisation maximum over the 3 K code-block values only (the reference's turbo_decode, liblte_phy.cc:10645-10660)
    int vshift;               // N_id_cell mod 6: offset of the port-0 pilot tones in symbol 0 (noise estimate of the soft path)
    FusedAlloc al[8];
};

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
    const int32_t *first_k;   // [3(K+4)]         rate-unmatch gather map
    const int32_t *first_kT;  // the same in the decoder's window-transposed output order + 12 tail entries (pair kernel only)
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
    FusedAlloc al[8];
};

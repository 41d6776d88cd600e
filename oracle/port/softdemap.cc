/* TEST INFRASTRUCTURE ONLY.  CPU twin of the soft ("north-star") symbol level of openlte_b200/csrc/pdsch_llr.cu + pdsch_dev.cuh:
 * noise variance from the CRS, per-RE MMSE weighting and exact max-log LLRs per bit.  The reference has no counterpart (it equalises by zero
 * forcing and makes hard 16/64QAM decisions, liblte_phy.cc:7680-7690, :9576 "FIXME: Need to make soft decisions"); this restates the
 * kernel's float definition operation by operation (IEEE single precision, no contraction: the build uses -ffp-contract=off) so the test can
 * demand identical int8 soft bits.
 *   bit order and sign: the reference's (positive = bit 0; I bits even, Q bits odd; 36.211 Gray mapping as modulation_mapper, liblte_phy.cc:8704)
 *   soft bit = round_to_nearest_even(2 * LLR) clamped to +-127
 */
#include "oracle_port.h"
#include <math.h>

namespace {
void maxlog_dim(float u, int Qm, float w[3])
{
    const float v = fabsf(u);
    if (Qm == 2) { w[0] = 4.0f * u; return; }
    if (Qm == 4) {
        w[0] = v <= 2.0f ? 4.0f * u : (u > 0.f ? 8.0f * u - 8.0f : 8.0f * u + 8.0f);
        w[1] = 4.0f * (2.0f - v);
        return;
    }
    float a, c;
    if (v <= 2.0f) { a = 4.0f; c = 0.0f; } else if (v <= 4.0f) { a = 8.0f; c = 8.0f; } else if (v <= 6.0f) { a = 12.0f; c = 24.0f; } else { a = 16.0f; c = 48.0f; }
    w[0] = u > 0.f ? a * u - c : a * u + c;
    w[1] = v <= 2.0f ? 24.0f - 8.0f * v : (v <= 6.0f ? 16.0f - 4.0f * v : 40.0f - 8.0f * v);
    w[2] = v <= 4.0f ? 4.0f * v - 8.0f : 24.0f - 4.0f * v;
}
int clamp127(float t)
{
    const long r = lrintf(t);       /* round to nearest even in the default rounding mode, like cvt.rni */
    return (int)(r < -127 ? -127 : (r > 127 ? 127 : r));
}
void demap_one(float xr, float xi, float gain, int mod, int b[6])
{
    const int Qm = mod == 1 ? 2 : (mod == 2 ? 4 : (mod == 3 ? 6 : 1));
    if (Qm == 1) {
        const float s = xr + xi;
        const float t = (2.0f * gain) * (2.8284271247461903f * (0.70710678118654752440f * s));
        b[0] = clamp127(t);
        return;
    }
    const float inv_a = Qm == 2 ? 1.4142135623730951f : (Qm == 4 ? 3.1622776601683795f : 6.4807406984078604f);
    const float a2 = Qm == 2 ? 0.5f : (Qm == 4 ? 0.1f : 0.023809523809523808f);
    const float k = (2.0f * gain) * a2;
    float wi[3], wq[3];
    maxlog_dim(xr * inv_a, Qm, wi);
    maxlog_dim(xi * inv_a, Qm, wq);
    for (int t = 0; t < 3; t++)
        if (2 * t < Qm) { b[2 * t] = clamp127(k * wi[t]); b[2 * t + 1] = clamp127(k * wq[t]); }
}
}  // namespace

/* h: channel estimate of port 0, symbol 0 ([nsc]); pilot tones at 6 j + vshift */
extern "C" float op_noise_var_crs(const float *h_re, const float *h_im, int nsc, int vshift)
{
    const int n = nsc / 6;
    float acc = 0.f, pw = 0.f;
    for (int j = 1; j + 1 < n; j++) {
        const int k = 6 * j + vshift;
        const float dr = h_re[k] - 0.5f * (h_re[k - 6] + h_re[k + 6]), di = h_im[k] - 0.5f * (h_im[k - 6] + h_im[k + 6]);
        acc = acc + (dr * dr + di * di);
        pw = pw + (h_re[k] * h_re[k] + h_im[k] * h_im[k]);
    }
    const float cnt = (float)(n - 2);
    float s2 = (0.6666666865348816f * acc) / cnt;
    const float floor_ = (pw / cnt) * 1e-6f;
    if (!(s2 > floor_)) s2 = floor_;
    if (!(s2 > 0.f)) s2 = 1e-12f;
    return s2;
}

/* One antenna port: y, h per resource element (M of them); cseq: scrambling bits [M * Qm]; soft: [M * Qm] int8, descrambled. */
extern "C" void op_soft_bits_1ant(const float *y_re, const float *y_im, const float *h_re, const float *h_im, int M, int mod, float s2,
                                  const uint8_t *cseq, int8_t *soft)
{
    const int Qm = mod == 1 ? 2 : (mod == 2 ? 4 : (mod == 3 ? 6 : 1));
    for (int r = 0; r < M; r++) {
        /* zero forcing in the reference's operation order (liblte_phy.cc:7684-7689) */
        const float nrm = h_re[r] * h_re[r] + h_im[r] * h_im[r];
        const float xr = (y_re[r] * h_re[r] + y_im[r] * h_im[r]) / nrm, xi = (y_im[r] * h_re[r] - y_re[r] * h_im[r]) / nrm;
        int b[6];
        demap_one(xr, xi, nrm / s2, mod, b);
        for (int q = 0; q < Qm; q++) soft[r * Qm + q] = (int8_t)(cseq[r * Qm + q] ? -b[q] : b[q]);
    }
}

/* Two ports (transmit diversity): d = the reference's combiner output (op_pre_decoder_dl, divides by sqrt(|h0|^4 + |h1|^4), liblte_phy.cc:7700),
 * h0 / h1 = the two ports' estimates at every resource element (the pair uses the EVEN element's, :7696-7699). */
extern "C" void op_soft_bits_2ant(const float *d_re, const float *d_im, const float *h0_re, const float *h0_im, const float *h1_re, const float *h1_im,
                                  int M, int mod, float s2, const uint8_t *cseq, int8_t *soft)
{
    const int Qm = mod == 1 ? 2 : (mod == 2 ? 4 : (mod == 3 ? 6 : 1));
    for (int r = 0; r + 1 < M; r += 2) {
        const float a0 = h0_re[r] * h0_re[r] + h0_im[r] * h0_im[r], a1 = h1_re[r] * h1_re[r] + h1_im[r] * h1_im[r], den = a0 + a1;
        const float rs = sqrtf(a0 * a0 + a1 * a1) / den, g = den / s2;
        for (int e = 0; e < 2; e++) {
            int b[6];
            demap_one(d_re[r + e] * rs, d_im[r + e] * rs, g, mod, b);
            for (int q = 0; q < Qm; q++) soft[(r + e) * Qm + q] = (int8_t)(cseq[(r + e) * Qm + q] ? -b[q] : b[q]);
        }
    }
}

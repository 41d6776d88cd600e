/* TEST INFRASTRUCTURE ONLY.  CPU restatement of the OFDM demodulator, CRS channel estimator and the PDSCH
 * symbol-level receive chain of the reference.
 *   samples_to_symbols_dl               /root/reference/liblte/src/liblte_phy.cc:8593-8644
 *   liblte_phy_get_dl_subframe_and_ce   liblte_phy.cc:5905-6200   (wrap_phase :14105-14116, generate_crs :8300-8333)
 *   generate_prs_c                      liblte_phy.cc:9669-9704
 *   calc_crc                            liblte_phy.cc:9713-9743
 *   RE extraction / ZF / demap / descramble / dlsch decode   liblte_phy.cc:3690-3853, 7645-7690, 9502-9660, 12762-12872
 * Float discipline as in the reference build: plain float +,-,*,/ (no contraction), sqrtf, glibc atan2f/cosf/sinf.
 */
#include "oracle_port.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <utility>

/* ------------------------------------------------------------------------------------------- sequences */
extern "C" void op_prs_c(uint32_t c_init, int len, uint8_t *c)
{
    /* Gold sequence of 36.211 7.2 in the reference's register convention: 31-bit registers, bit 0 is the oldest
     * element, new element enters at bit 30.  x1 after 1600 steps from 0x1 is the constant below (liblte_phy.cc:9691);
     * x2 starts at c_init and is advanced 1600-31 times (9682-9688). */
    uint32_t x1 = 0x54D21B24u, x2 = c_init;
    for (int i = 0; i < 1600 - 31; i++) {
        uint32_t n2 = ((x2 >> 3) ^ (x2 >> 2) ^ (x2 >> 1) ^ x2) & 1u;
        x2 = (x2 >> 1) | (n2 << 30);
    }
    for (int i = 0; i < len; i++) {
        uint32_t n1 = ((x1 >> 3) ^ x1) & 1u;
        uint32_t n2 = ((x2 >> 3) ^ (x2 >> 2) ^ (x2 >> 1) ^ x2) & 1u;
        x1 = (x1 >> 1) | (n1 << 30);
        x2 = (x2 >> 1) | (n2 << 30);
        c[i] = (uint8_t)(n1 ^ n2);
    }
}

extern "C" uint32_t op_crc(const uint8_t *bits, int n, uint32_t poly, int n_p)
{
    /* MSB-first long division; `poly` includes the leading term (CRC24A 0x1864CFB, CRC24B 0x1800063, CRC16 0x11021) */
    uint32_t rem = 0;
    const uint32_t top = 1u << n_p;
    for (int i = 0; i < n + n_p; i++) {
        rem = (rem << 1) | (i < n ? (bits[i] & 1u) : 0u);
        if (rem & top) rem ^= poly;
    }
    return rem & (top - 1);
}

static const int KTAB_STEP[4][3] = {{40, 512, 8}, {528, 1024, 16}, {1056, 2048, 32}, {2112, 6144, 64}};
extern "C" void op_segmentation_params(int tbs, int *K, int *F, int *C)
{
    /* liblte_phy_code_block_segmentation, C == 1 branch (liblte_phy.cc:9781-9795): smallest K >= tbs+24 */
    const int B = tbs + 24;
    *C = (B <= 6144) ? 1 : 0;   /* multi-code-block is broken in the reference: not restated */
    *K = 0;
    for (int r = 0; r < 4 && !*K; r++)
        for (int k = KTAB_STEP[r][0]; k <= KTAB_STEP[r][1]; k += KTAB_STEP[r][2])
            if (k >= B) { *K = k; break; }
    *F = *K - B;
}

/* --------------------------------------------------------------------------------------------- demapping */
extern "C" void op_zf_1ant(const float *y_re, const float *y_im, const float *h_re, const float *h_im, int n, float *x_re, float *x_im)
{
    for (int i = 0; i < n; i++) {   /* liblte_phy.cc:7680-7690 */
        float nrm = h_re[i] * h_re[i] + h_im[i] * h_im[i];
        x_re[i] = (y_re[i] * h_re[i] + y_im[i] * h_im[i]) / nrm;
        x_im[i] = (y_im[i] * h_re[i] - y_re[i] * h_im[i]) / nrm;
    }
}

/* pre_decoder_and_matched_filter_dl + layer_demapper_dl for transmit diversity (liblte_phy.cc:7645-7796, 7473-7514).
 * h_*: [N_ant][h_len]; returns M_symb and writes the layer-demapped symbols d[]. */
extern "C" int op_pre_decoder_dl(const float *yr, const float *yi, const float *h_re, const float *h_im, int h_len, int M, int N_ant,
                                 float *d_re, float *d_im)
{
    if (N_ant == 1) { op_zf_1ant(yr, yi, h_re, h_im, M, d_re, d_im); return M; }
    const float *hr[4], *hi[4];
    for (int p = 0; p < N_ant; p++) { hr[p] = h_re + (size_t)p * h_len; hi[p] = h_im + (size_t)p * h_len; }
    if (N_ant == 2) {
        for (int i = 0; i < M / 2; i++) {
            const float h0r = hr[0][2 * i], h0i = hi[0][2 * i], h1r = hr[1][2 * i], h1i = hi[1][2 * i];
            const float y0r = yr[2 * i], y0i = yi[2 * i], y1r = yr[2 * i + 1], y1i = yi[2 * i + 1];
            const float a0 = h0r * h0r + h0i * h0i, a1 = h1r * h1r + h1i * h1i;
            const float n = sqrtf(a0 * a0 + a1 * a1);                                  /* sic: sqrt(|h0|^4 + |h1|^4), :7700 */
            d_re[2 * i]     = (h0r * y0r + h0i * y0i + h1r * y1r + h1i * y1i) / n;
            d_im[2 * i]     = (h0r * y0i - h0i * y0r - h1r * y1i + h1i * y1r) / n;
            d_re[2 * i + 1] = (-h1r * y0r - h1i * y0i + h0r * y1r + h0i * y1i) / n;
            d_im[2 * i + 1] = (h1r * y0i - h1i * y0r + h0r * y1i - h0i * y1r) / n;
        }
        return (M / 2) * 2;
    }
    int i;
    for (i = 0; i < M / 4; i++) {
        const float h0r = hr[0][4 * i], h0i = hi[0][4 * i], h2r = hr[2][4 * i], h2i = hi[2][4 * i];
        const float h1r = hr[1][4 * i + 2], h1i = hi[1][4 * i + 2], h3r = hr[3][4 * i + 2], h3i = hi[3][4 * i + 2];
        const float y0r = yr[4 * i], y0i = yi[4 * i], y1r = yr[4 * i + 1], y1i = yi[4 * i + 1];
        const float y2r = yr[4 * i + 2], y2i = yi[4 * i + 2], y3r = yr[4 * i + 3], y3i = yi[4 * i + 3];
        const float a0 = h0r * h0r + h0i * h0i, a1 = h1r * h1r + h1i * h1i, a2 = h2r * h2r + h2i * h2i, a3 = h3r * h3r + h3i * h3i;
        const float n02 = sqrtf(a0 * a0 + a2 * a2), n13 = sqrtf(a1 * a1 + a3 * a3);
        d_re[4 * i]     = (h0r * y0r + h0i * y0i + h2r * y1r + h2i * y1i) / n02;
        d_im[4 * i]     = (h0r * y0i - h0i * y0r - h2r * y1i + h2i * y1r) / n02;
        d_re[4 * i + 1] = (-h2r * y0r - h2i * y0i + h0r * y1r + h0i * y1i) / n02;
        d_im[4 * i + 1] = -(-h2r * y0i + h2i * y0r - h0r * y1i + h0i * y1r) / n02;
        d_re[4 * i + 2] = (h1r * y2r + h1i * y2i + h3r * y3r + h3i * y3i) / n13;
        d_im[4 * i + 2] = (h1r * y2i - h1i * y2r - h3r * y3i + h3i * y3r) / n13;
        d_re[4 * i + 3] = (-h3r * y2r - h3i * y2i + h1r * y3r + h1i * y3i) / n13;
        d_im[4 * i + 3] = -(-h3r * y2i + h3i * y2r - h1r * y3i + h1i * y3r) / n13;
    }
    if (M % 4 != 0) {           /* odd tail: two symbols on ports 0/2 only (:7766-7795, with its different x1_im) */
        const float h0r = hr[0][4 * i], h0i = hi[0][4 * i], h2r = hr[2][4 * i], h2i = hi[2][4 * i];
        const float y0r = yr[4 * i], y0i = yi[4 * i], y1r = yr[4 * i + 1], y1i = yi[4 * i + 1];
        const float a0 = h0r * h0r + h0i * h0i, a2 = h2r * h2r + h2i * h2i;
        const float n02 = sqrtf(a0 * a0 + a2 * a2);
        d_re[4 * i]     = (h0r * y0r + h0i * y0i + h2r * y1r + h2i * y1i) / n02;
        d_im[4 * i]     = (h0r * y0i - h0i * y0r - h2r * y1i + h2i * y1r) / n02;
        d_re[4 * i + 1] = (-h2r * y0r - h2i * y0i + h0r * y1r + h0i * y1i) / n02;
        d_im[4 * i + 1] = (-h2r * y0i + h2i * y0r - h0r * y1r + h0i * y1i) / n02;
        return 4 * i + 2;
    }
    return 4 * i;
}

static inline float soft_dist(float re, float im, float ere, float eim)
{
    /* get_soft_decision(rx, exp, max_dist = 1), liblte_phy.cc:13880-13900 */
    float dr = re - ere, di = im - eim;
    float dist = sqrtf(dr * dr + di * di);
    const float lim = 1.0f - (1.0f / 120);
    if (dist >= lim) dist = lim;
    return 1.0f - dist;
}

extern "C" int op_demap(const float *d_re, const float *d_im, int M, int mod, int8_t *bits)
{
    const float r2 = (float)(1 / sqrt(2.0)), t10 = (float)(2 / sqrt(10.0));
    const float t42_2 = (float)(2 / sqrt(42.0)), t42_4 = (float)(4 / sqrt(42.0)), t42_6 = (float)(6 / sqrt(42.0));
    if (mod == 0) {            /* BPSK, liblte_phy.cc:9521-9539 */
        for (int i = 0; i < M; i++) {
            float ang = atan2f(d_im[i], d_re[i]);
            if (ang > -M_PI / 4 && ang < 3 * M_PI / 4) bits[i] = (int8_t)(127 * soft_dist(d_re[i], d_im[i], r2, r2));
            else bits[i] = (int8_t)(-(int8_t)(127 * soft_dist(d_re[i], d_im[i], -r2, -r2)));
        }
        return M;
    }
    if (mod == 1) {            /* QPSK, liblte_phy.cc:9540-9572: one magnitude for both bits, signs by quadrant */
        for (int i = 0; i < M; i++) {
            float ang = atan2f(d_im[i], d_re[i]);
            int s0, s1;
            if (ang >= 0 && ang < M_PI / 2) { s0 = 1; s1 = 1; }
            else if (ang >= -M_PI / 2 && ang < 0) { s0 = 1; s1 = -1; }
            else if (ang >= M_PI / 2 && ang < M_PI) { s0 = -1; s1 = 1; }
            else { s0 = -1; s1 = -1; }
            int8_t m = (int8_t)(127 * soft_dist(d_re[i], d_im[i], s0 * r2, s1 * r2));
            bits[2 * i] = (int8_t)(s0 * m); bits[2 * i + 1] = (int8_t)(s1 * m);
        }
        return 2 * M;
    }
    if (mod == 2) {            /* 16QAM hard decisions, liblte_phy.cc:9573-9605 */
        for (int i = 0; i < M; i++) {
            bits[4 * i + 0] = d_re[i] > 0 ? 127 : -127;
            bits[4 * i + 1] = d_im[i] > 0 ? 127 : -127;
            bits[4 * i + 2] = fabsf(d_re[i]) < t10 ? 127 : -127;
            bits[4 * i + 3] = fabsf(d_im[i]) < t10 ? 127 : -127;
        }
        return 4 * M;
    }
    for (int i = 0; i < M; i++) {   /* 64QAM hard decisions, liblte_phy.cc:9606-9658 */
        const float ar = fabsf(d_re[i]), ai = fabsf(d_im[i]);
        bits[6 * i + 0] = d_re[i] > 0 ? 127 : -127;
        bits[6 * i + 1] = d_im[i] > 0 ? 127 : -127;
        if (ar < t42_4) { bits[6 * i + 2] = 127; bits[6 * i + 4] = ar > t42_2 ? 127 : -127; }
        else            { bits[6 * i + 2] = -127; bits[6 * i + 4] = ar < t42_6 ? 127 : -127; }
        if (ai < t42_4) { bits[6 * i + 3] = 127; bits[6 * i + 5] = ai > t42_2 ? 127 : -127; }
        else            { bits[6 * i + 3] = -127; bits[6 * i + 5] = ai < t42_6 ? 127 : -127; }
    }
    return 6 * M;
}

static void centre_band(int N_rb_dl, int *first_sc, int *last_sc)
{
    /* PBCH/PSS/SSS subcarriers, liblte_phy.cc:3722-3741 */
    switch (N_rb_dl) {
    case 6:  *first_sc = 0;   *last_sc = 71;  break;
    case 15: *first_sc = 54;  *last_sc = 125; break;
    case 25: *first_sc = 114; *last_sc = 185; break;
    case 50: *first_sc = 264; *last_sc = 335; break;
    case 75: *first_sc = 414; *last_sc = 485; break;
    default: *first_sc = 564; *last_sc = 635; break;
    }
}

extern "C" int op_pdsch_re_indices(int N_rb_dl, int N_id_cell, int N_ant, int subfr, int N_pdcch_symbs, int N_prb,
                                   const int *prb0, const int *prb1, int *idx)
{
    int first_sc, last_sc, n = 0;
    centre_band(N_rb_dl, &first_sc, &last_sc);
    for (int L = N_pdcch_symbs; L < 14; L++)
        for (int pi = 0; pi < N_prb; pi++) {
            const int i = (L / 7 == 0) ? prb0[pi] : prb1[pi];
            for (int j = 0; j < 12; j++) {
                const int sc = 12 * i + j, l7 = L % 7;
                bool skip = false;
                if (N_ant == 1) skip = (l7 == 0 && (N_id_cell % 6) == (j % 6)) || (l7 == 4 && ((N_id_cell + 3) % 6) == (j % 6));
                else {
                    skip = ((l7 == 0 || l7 == 4) && (N_id_cell % 3) == (j % 3));
                    if (N_ant == 4 && l7 == 1 && (N_id_cell % 3) == (j % 3)) skip = true;
                }
                const bool centre = sc >= first_sc && sc <= last_sc;
                if (subfr == 0 && centre && L >= 7 && L <= 10) skip = true;            /* PBCH */
                if ((subfr == 0 || subfr == 5) && centre && (L == 6 || L == 5)) skip = true;  /* PSS, SSS */
                if (!skip) idx[n++] = L * 1200 + sc;
            }
        }
    return n;
}

/* ---------------------------------------------------------------------------------------------- OFDM */
struct FsParams { int N, cp0, cpe, slot; };
static FsParams fs_params(int fs_idx)
{
    FsParams p; p.N = 128 << fs_idx; p.cp0 = 10 << fs_idx; p.cpe = 9 << fs_idx; p.slot = 960 << fs_idx; return p;
}

/* Unnormalised forward DFT.  The reference delegates this to FFTW3f (unpinned, liblte_phy.cc:8624); the oracle's
 * DFT is computed in double precision (recursive radix-2) and rounded once to float, i.e. the correctly rounded
 * answer any single-precision FFT approximates.  Tolerance lives in the tests, not here. */
static void fft_rec(int n, int stride, const double *in_re, const double *in_im, double *out_re, double *out_im, int sign)
{
    if (n == 1) { out_re[0] = in_re[0]; out_im[0] = in_im[0]; return; }
    const int h = n / 2;
    fft_rec(h, 2 * stride, in_re, in_im, out_re, out_im, sign);
    fft_rec(h, 2 * stride, in_re + stride, in_im + stride, out_re + h, out_im + h, sign);
    for (int k = 0; k < h; k++) {
        const double a = sign * 2.0 * M_PI * k / n, c = cos(a), s = sin(a);
        const double tr = out_re[h + k] * c - out_im[h + k] * s, ti = out_re[h + k] * s + out_im[h + k] * c;
        const double er = out_re[k], ei = out_im[k];
        out_re[k] = er + tr; out_im[k] = ei + ti; out_re[h + k] = er - tr; out_im[h + k] = ei - ti;
    }
}
extern "C" void op_fft_forward(int n, const float *in_re, const float *in_im, float *out_re, float *out_im)
{
    std::vector<double> a(n), b(n), c(n), d(n);
    for (int i = 0; i < n; i++) { a[i] = in_re[i]; b[i] = in_im[i]; }
    fft_rec(n, 1, a.data(), b.data(), c.data(), d.data(), -1);
    for (int i = 0; i < n; i++) { out_re[i] = (float)c[i]; out_im[i] = (float)d[i]; }
}

extern "C" void op_samples_to_symbols_dl(const float *i_samps, const float *q_samps, int fs_idx, int N_rb_dl, uint32_t slot_start,
                                         int symbol_offset, float *symb_re, float *symb_im)
{
    const FsParams p = fs_params(fs_idx);
    const int cp = (symbol_offset % 7 == 0) ? p.cp0 : p.cpe;
    uint32_t index = slot_start + (uint32_t)(p.N + p.cpe) * symbol_offset;
    if (symbol_offset > 0) index += p.cp0 - p.cpe;
    std::vector<float> or_(p.N), oi(p.N);
    op_fft_forward(p.N, i_samps + index + cp - 1, q_samps + index + cp - 1, or_.data(), oi.data());   /* one sample early, :8621 */
    const int h = 6 * N_rb_dl;
    for (int i = 0; i < h; i++) {
        symb_re[h + i] = or_[i + 1];          symb_im[h + i] = oi[i + 1];             /* DC bin dropped */
        symb_re[h - 1 - i] = or_[p.N - 1 - i]; symb_im[h - 1 - i] = oi[p.N - 1 - i];
    }
}

extern "C" void op_crs(int N_s, int L, int N_id_cell, float *re, float *im)
{
    const float r2 = (float)(1 / sqrt(2.0));
    const uint32_t c_init = 1024u * (7 * (N_s + 1) + L + 1) * (2 * N_id_cell + 1) + 2 * N_id_cell + 1;   /* normal CP */
    uint8_t c[440];
    op_prs_c(c_init, 440, c);
    for (int i = 0; i < 220; i++) { re[i] = r2 * (1 - 2 * (float)c[2 * i]); im[i] = r2 * (1 - 2 * (float)c[2 * i + 1]); }
}

static inline void wrap_phase(float *p1, float p2)
{
    while ((*p1 - p2) >= M_PI) *p1 = *p1 - 2 * M_PI;      /* float difference compared and corrected in double, :14105-14116 */
    while ((*p1 - p2) <= -M_PI) *p1 = *p1 + 2 * M_PI;
}

extern "C" void op_chan_est(const float *rx_re, const float *rx_im, int N_rb_dl, int subfr, int N_id_cell, int N_ant,
                            float *ce_re, float *ce_im)
{
    const int nsc = 12 * N_rb_dl, vsh = N_id_cell % 6;
    std::vector<float> mag(5 * 1200), ang(5 * 1200), crs_re(16 * 220), crs_im(16 * 220);
    /* CRS of the 8 (slot, symbol) pairs used, indexed by grid symbol (liblte_phy.cc:5960-5967) */
    const int gsym[8] = {0, 1, 4, 7, 8, 11, 14, 15}, gslot[8] = {0, 0, 0, 1, 1, 1, 2, 2}, gl[8] = {0, 1, 4, 0, 1, 4, 0, 1};
    for (int g = 0; g < 8; g++) op_crs((subfr * 2 + gslot[g]) % 20, gl[g], N_id_cell, &crs_re[gsym[g] * 220], &crs_im[gsym[g] * 220]);
    for (int p = 0; p < N_ant; p++) {
        int v[5], sym[5], N_sym;
        if (p < 2) { N_sym = 5; const int s5[5] = {0, 4, 7, 11, 14}; for (int i = 0; i < 5; i++) { sym[i] = s5[i]; v[i] = ((i & 1) ^ p) ? 3 : 0; } }
        else { N_sym = 3; const int s3[3] = {1, 8, 15}; for (int i = 0; i < 3; i++) { sym[i] = s3[i]; v[i] = (p == 2) ? ((i & 1) ? 3 : 0) : ((i & 1) ? 6 : 3); } }
        for (int i = 0; i < N_sym; i++) {
            float *M = &mag[i * 1200], *A = &ang[i * 1200];
            const float *sr = rx_re + sym[i] * 1200, *si = rx_im + sym[i] * 1200;
            const float *rr = &crs_re[sym[i] * 220], *ri = &crs_im[sym[i] * 220];
            const int off = (v[i] + vsh) % 6;
            float fm = 0, fa = 0; int k = 0;
            for (int j = 0; j < 2 * N_rb_dl; j++) {
                k = 6 * j + off;
                const int mp = j + 110 - N_rb_dl;
                const float tr = sr[k] * rr[mp] + si[k] * ri[mp], ti = si[k] * rr[mp] - sr[k] * ri[mp];
                M[k] = sqrtf(tr * tr + ti * ti);
                A[k] = atan2f(ti, tr);
                if (j > 0) {
                    wrap_phase(&A[k], A[k - 6]);
                    fm = (M[k] - M[k - 6]) / 6; fa = (A[k] - A[k - 6]) / 6;
                    for (int z = 1; z < 6; z++) { M[k - z] = M[k - (z - 1)] - fm; A[k - z] = A[k - (z - 1)] - fa; }
                }
                if (j == 1) for (int z = 1; z < off + 1; z++) { M[k - 6 - z] = M[k - 6 - (z - 1)] - fm; A[k - 6 - z] = A[k - 6 - (z - 1)] - fa; }
            }
            for (int z = 1; z < (5 - off) + 1; z++) { M[k + z] = M[k + (z - 1)] - fm; A[k + z] = A[k + (z - 1)] - fa; }   /* sic: minus */
        }
        float *cr = ce_re + p * 16 * 1200, *ci = ce_im + p * 16 * 1200;
        auto put = [&](int z, int j, float m, float a) { cr[z * 1200 + j] = m * cosf(a); ci[z * 1200 + j] = m * sinf(a); };
        /* interpolate from CRS symbol index hi down over symbols (z_from .. z_to], polar, liblte_phy.cc:6067-6193 */
        auto seg = [&](int j, int lo, int hi, float div, int z_hi, int z_lo, bool start_below) {
            float *Alo = &ang[lo * 1200], *Ahi = &ang[hi * 1200], *Mlo = &mag[lo * 1200], *Mhi = &mag[hi * 1200];
            float fm = (Mhi[j] - Mlo[j]) / div;
            wrap_phase(&Ahi[j], Alo[j]);
            float fa = Ahi[j] - Alo[j];
            wrap_phase(&fa, 0);
            fa /= div;
            float m = Mhi[j], a = Ahi[j];
            if (start_below) { m -= fm; a -= fa; }
            for (int z = z_hi; z > z_lo; z--) { m -= fm; a -= fa; put(z, j, m, a); }
            return std::pair<float, float>(fm, fa);
        };
        for (int j = 0; j < nsc; j++) {
            if (N_sym == 3) {
                put(1, j, mag[j], ang[j]); put(8, j, mag[1200 + j], ang[1200 + j]);
                auto f = seg(j, 0, 1, 7.0f, 7, 1, false);
                put(0, j, mag[j] - f.first, ang[j] - f.second);          /* symbol 0 extrapolated, FIXME in the reference too */
                seg(j, 1, 2, 7.0f, 13, 8, true);
            } else {
                put(0, j, mag[j], ang[j]); put(4, j, mag[1200 + j], ang[1200 + j]);
                put(7, j, mag[2400 + j], ang[2400 + j]); put(11, j, mag[3600 + j], ang[3600 + j]);
                seg(j, 0, 1, 4.0f, 3, 0, false);
                seg(j, 1, 2, 3.0f, 6, 4, false);
                seg(j, 2, 3, 4.0f, 10, 7, false);
                seg(j, 3, 4, 3.0f, 13, 11, false);
            }
        }
    }
}

extern "C" void op_get_dl_subframe_and_ce(const float *i_samps, const float *q_samps, int fs_idx, int N_rb_dl, uint32_t frame_start_idx,
                                          int subfr, int N_id_cell, int N_ant, float *rx_re, float *rx_im, float *ce_re, float *ce_im)
{
    const FsParams p = fs_params(fs_idx);
    const uint32_t sf_start = frame_start_idx + (uint32_t)subfr * 2 * p.slot;
    for (int i = 0; i < 16; i++)
        op_samples_to_symbols_dl(i_samps, q_samps, fs_idx, N_rb_dl, sf_start + (i / 7) * p.slot, i % 7, rx_re + i * 1200, rx_im + i * 1200);
    op_chan_est(rx_re, rx_im, N_rb_dl, subfr, N_id_cell, N_ant, ce_re, ce_im);
}

/* ------------------------------------------------------------------------------------- PDSCH, one allocation */
/* ce_*: [N_ant][16][1200] (port-major, as LIBLTE_PHY_SUBFRAME_STRUCT); soft_out (optional): the descrambled soft bits */
extern "C" int op_pdsch_decode_nant(const float *rx_re, const float *rx_im, const float *ce_re, const float *ce_im, int N_rb_dl, int N_id_cell,
                                    int N_ant, int subfr, int N_pdcch_symbs, int mod, int N_prb, const int *prb, int tbs, int rv, int rnti,
                                    int turbo_mode, int n_iter, uint8_t *out_bits, float *soft_out, int *n_soft)
{
    std::vector<int> idx(14 * 12 * N_prb);
    const int Mre = op_pdsch_re_indices(N_rb_dl, N_id_cell, N_ant, subfr, N_pdcch_symbs, N_prb, prb, prb, idx.data());
    std::vector<float> yr(Mre), yi(Mre), hr((size_t)N_ant * Mre), hi((size_t)N_ant * Mre), xr(Mre + 4), xi(Mre + 4);
    for (int i = 0; i < Mre; i++) {
        yr[i] = rx_re[idx[i]]; yi[i] = rx_im[idx[i]];
        for (int p = 0; p < N_ant; p++) { hr[(size_t)p * Mre + i] = ce_re[(size_t)p * 16 * 1200 + idx[i]]; hi[(size_t)p * Mre + i] = ce_im[(size_t)p * 16 * 1200 + idx[i]]; }
    }
    const int M = op_pre_decoder_dl(yr.data(), yi.data(), hr.data(), hi.data(), Mre, Mre, N_ant, xr.data(), xi.data());
    std::vector<int8_t> soft(6 * (size_t)M + 8);
    const int G = op_demap(xr.data(), xi.data(), M, mod, soft.data());
    if (n_soft) *n_soft = G;
    std::vector<uint8_t> c(G);
    op_prs_c(((uint32_t)rnti << 14) | ((uint32_t)subfr << 9) | (uint32_t)N_id_cell, G, c.data());
    std::vector<float> llr(G);
    for (int i = 0; i < G; i++) llr[i] = (float)soft[i] * (1 - 2 * (float)c[i]);
    if (soft_out) memcpy(soft_out, llr.data(), sizeof(float) * G);
    int K, F, C;
    op_segmentation_params(tbs, &K, &F, &C);
    if (C != 1) return -4;
    std::vector<float> d(3 * (K + 4));
    op_rate_unmatch_turbo(llr.data(), G, K + 4, 1, 1, 250368, 8, 0, rv, d.data());
    std::vector<uint8_t> cb(K);
    if (turbo_mode == 0) op_turbo_decode_ref(d.data(), K, cb.data(), NULL);
    else op_turbo_decode_mlm(d.data(), K, n_iter, cb.data());
    const uint8_t *a = cb.data() + F;
    const uint32_t crc = op_crc(a, tbs, 0x1864CFB, 24);
    uint32_t rx = 0;
    for (int i = 0; i < 24; i++) rx = (rx << 1) | a[tbs + i];
    if (crc != rx) return -3;
    memcpy(out_bits, a, tbs);
    return 0;
}

extern "C" int op_pdsch_decode(const float *rx_re, const float *rx_im, const float *ce_re, const float *ce_im, int N_rb_dl, int N_id_cell,
                               int subfr, int N_pdcch_symbs, int mod, int N_prb, const int *prb, int tbs, int rv, int rnti,
                               int turbo_mode, int n_iter, uint8_t *out_bits)
{
    return op_pdsch_decode_nant(rx_re, rx_im, ce_re, ce_im, N_rb_dl, N_id_cell, 1, subfr, N_pdcch_symbs, mod, N_prb, prb, tbs, rv, rnti, turbo_mode,
                                n_iter, out_bits, NULL, NULL);
}

/* ------------------------------------------------------------------------------------------------ front end
 * LTE_fdd_dl_fs_samp_buf::freq_shift, LTE_fdd_dl_file_scan/src/LTE_fdd_dl_fs_samp_buf.cc:696-713: the phase is
 * (i+1)*freq_offset as a FLOAT product, scaled by 2*pi/fs in double, rounded to float by cosf/sinf.  (The member
 * function lives in the GNU Radio block, which does not build here: this row's parity is pinned on the formula alone.) */
extern "C" void op_freq_shift(float *i_buf, float *q_buf, uint32_t start_idx, uint32_t num_samps, float freq_offset, uint32_t fs)
{
    for (uint32_t i = start_idx; i < start_idx + num_samps; i++) {
        float f_samp_re = cosf((i + 1) * (freq_offset) * 2 * M_PI / fs);
        float f_samp_im = sinf((i + 1) * (freq_offset) * 2 * M_PI / fs);
        float tmp_i = i_buf[i], tmp_q = q_buf[i];
        i_buf[i] = tmp_i * f_samp_re + tmp_q * f_samp_im;
        q_buf[i] = tmp_q * f_samp_re - tmp_i * f_samp_im;
    }
}

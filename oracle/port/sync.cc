/* TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's cell-search functions:
 *   liblte_phy_dl_find_coarse_timing_and_freq_offset   /root/reference/liblte/src/liblte_phy.cc:5697-5852
 *   liblte_phy_find_pss_and_fine_timing                liblte_phy.cc:5306-5510   (generate_pss :8342-8368)
 *   liblte_phy_find_sss                                liblte_phy.cc:5578-5686   (generate_sss :8377-8474)
 * The DFT inside uses this port's double-precision transform (the reference uses FFTW3f), so correlation VALUES
 * agree with the reference to float tolerance while every integer decision is expected to be identical.
 */
#include "oracle_port.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

struct FsP { int N, cp0, cpe, slot, frame; };
static FsP fsp(int fs_idx) { FsP p; p.N = 128 << fs_idx; p.cp0 = 10 << fs_idx; p.cpe = 9 << fs_idx; p.slot = 960 << fs_idx; p.frame = p.slot * 20; return p; }

extern "C" int op_find_coarse_timing(const float *is, const float *qs, int fs_idx, int N_slots, float *freq_offset, uint32_t *symb_starts,
                                     int *n_corr_peaks)
{
    const FsP p = fsp(fs_idx);
    const int per_else = p.N + p.cpe, blank = per_else / 10;
    std::vector<float> ac(2 * 15360 * 2 + 16, 0.0f);
    for (int slot = 0; slot < N_slots; slot++)
        for (int i = 0; i < p.slot; i++) {
            float cr = 0, ci = 0;
            for (int j = 0; j < p.cpe; j++) {
                const size_t idx = (size_t)slot * p.slot + i + j;
                cr += is[idx] * is[idx + p.N] + qs[idx] * qs[idx + p.N];
                ci += is[idx] * qs[idx + p.N] - qs[idx] * is[idx + p.N];
            }
            ac[i] += cr * cr + ci * ci;
        }
    float mean = 0;
    for (int i = 0; i < p.slot; i++) { mean += ac[i]; ac[i + p.slot] = ac[i]; }
    mean /= p.slot;
    for (int i = 0; i < p.slot; i++) if (ac[i] <= mean) { ac[i] = 0; ac[i + p.slot] = 0; }
    const int lag = p.N + p.cp0 + per_else * 3;          /* first symbol x fourth symbol after it */
    for (int i = 0; i < p.slot; i++) ac[i] *= ac[lag + i];
    int peaks[5], n = 5;
    for (int k = 0; k < 5; k++) {
        float mx = 0; peaks[k] = 0;
        for (int j = 0; j < p.slot; j++) if (ac[j] > mx) { mx = ac[j]; peaks[k] = j; }
        if (mx == 0) { n = k; break; }
        int t = peaks[k];
        while (t > 0) t -= per_else;
        for (int j = 0; j < 7; j++) {
            t += per_else;
            for (int z = 0; z < blank; z++) {
                const uint32_t idx = (uint32_t)(t - (blank / 2) + z);      /* uint32 in the reference: negative wraps, fails the <= test */
                if (idx <= 15360u * 2) ac[idx] = 0;
            }
        }
    }
    *n_corr_peaks = n;
    float ferr[5] = {0, 0, 0, 0, 0};
    for (int slot = 0; slot < N_slots; slot++)
        for (int k = 0; k < n; k++) {
            float cr = 0, ci = 0;
            for (int j = 0; j < p.cpe; j++) {
                const size_t idx = (size_t)slot * p.slot + peaks[k] + j;
                cr += is[idx] * is[idx + p.N] + qs[idx] * qs[idx + p.N];
                ci += is[idx] * qs[idx + p.N] - qs[idx] * is[idx + p.N];
            }
            ferr[k] += atan2f(ci, cr) / (p.N * 2 * M_PI * (0.0005 / p.slot));
        }
    for (int k = 0; k < n; k++) {
        freq_offset[k] = ferr[k] / N_slots;
        int t = peaks[k];
        while (t > 0) t -= per_else;
        for (int j = 0; j < 7; j++) symb_starts[k * 7 + j] = (uint32_t)(t + (j + 1) * per_else);
    }
    return 0;
}

static void gen_pss(int N_id_2, float *re, float *im)
{
    const float root = N_id_2 == 0 ? 25 : (N_id_2 == 1 ? 29 : 34);
    for (uint32_t i = 0; i < 31; i++) { re[i] = cosf(-M_PI * root * i * (i + 1) / 63); im[i] = sinf(-M_PI * root * i * (i + 1) / 63); }
    for (uint32_t i = 31; i < 62; i++) { re[i] = cosf(-M_PI * root * (i + 1) * (i + 2) / 63); im[i] = sinf(-M_PI * root * (i + 1) * (i + 2) / 63); }
}

/* sum over the 62 taps placed at subcarrier (centre - 31 + j + shift); everything else is zero in the reference's arrays */
static void corr62(const float *rr, const float *ri, const float *sr, const float *si, int first, float *cr, float *ci)
{
    float a = 0, b = 0;
    for (int j = 0; j < 62; j++) {
        const int z = first + j;
        a += rr[z] * sr[j] + ri[z] * si[j];
        b += rr[z] * si[j] - ri[z] * sr[j];
    }
    *cr = a; *ci = b;
}

extern "C" int op_find_pss(const float *is, const float *qs, int fs_idx, int N_rb_dl, uint32_t *symb_starts, uint32_t *N_id_2,
                           uint32_t *pss_symb, float *pss_thresh, float *freq_offset)
{
    const FsP p = fsp(fs_idx);
    const int nsc = 12 * N_rb_dl, c0 = nsc / 2 - 31;
    float pr[3][62], pi[3][62];
    for (int k = 0; k < 3; k++) gen_pss(k, pr[k], pi[k]);
    std::vector<float> rr(1200), ri(1200);
    float corr_max = 0; int which = 0;
    for (int i = 0; i < 12; i++)
        for (int j = 0; j < 7; j++) {
            op_samples_to_symbols_dl(is, qs, fs_idx, N_rb_dl, symb_starts[j] + (uint32_t)p.slot * i, 0, rr.data(), ri.data());
            for (int k = 0; k < 3; k++)
                for (int sh = -1; sh <= 1; sh++) {
                    float cr, ci;
                    corr62(rr.data(), ri.data(), pr[k], pi[k], c0 + sh, &cr, &ci);
                    const float a = sqrtf(cr * cr + ci * ci);
                    if (a > corr_max) { which = sh; corr_max = a; *pss_symb = i * 7 + j; *N_id_2 = k; }
                }
        }
    *freq_offset = which * 15000.0f;
    const uint32_t N_s = *pss_symb / 7, N_symb = *pss_symb % 7;
    corr_max = 0; int timing = 0;
    for (int i = -40; i < 40; i++) {
        int idx = (int)(symb_starts[N_symb] + (uint32_t)p.slot * N_s);
        if (i < 0) { if (idx >= -i) idx += i; } else idx += i;
        op_samples_to_symbols_dl(is, qs, fs_idx, N_rb_dl, (uint32_t)idx, 0, rr.data(), ri.data());
        float cr, ci;
        corr62(rr.data(), ri.data(), pr[*N_id_2], pi[*N_id_2], c0 + which, &cr, &ci);
        const float a = sqrtf(cr * cr + ci * ci);
        if (a > corr_max) { corr_max = a; timing = i; }
    }
    *pss_thresh = corr_max;
    uint32_t t = symb_starts[N_symb] + (uint32_t)p.slot * N_s + (uint32_t)(int8_t)timing;
    while (t + p.N + p.cpe < (uint32_t)p.slot) t += p.frame;
    symb_starts[0] = t + (p.N + p.cpe) * 1 - p.slot;
    for (int i = 1; i < 7; i++) symb_starts[i] = t + (p.N + p.cpe) * i + p.N + p.cp0 - p.slot;
    return 0;
}

static void gen_sss(int N_id_1, int N_id_2, int8_t *s0, int8_t *s5)
{
    int q_prime = N_id_1 / 30, q = (N_id_1 + (q_prime * (q_prime + 1) / 2)) / 30, m_prime = N_id_1 + (q * (q + 1) / 2);
    int m0 = m_prime % 31, m1 = (m0 + (m_prime / 31) + 1) % 31;
    int xs[31] = {0}, xc[31] = {0}, xz[31] = {0};
    xs[4] = xc[4] = xz[4] = 1;
    for (int i = 0; i < 26; i++) {
        xs[i + 5] = (xs[i + 2] + xs[i]) % 2;
        xc[i + 5] = (xc[i + 3] + xc[i]) % 2;
        xz[i + 5] = (xz[i + 4] + xz[i + 2] + xz[i + 1] + xz[i]) % 2;
    }
    for (int i = 0; i < 31; i++) {
        const int s0m0 = 1 - 2 * xs[(i + m0) % 31], s1m1 = 1 - 2 * xs[(i + m1) % 31];
        const int c0 = 1 - 2 * xc[(i + N_id_2) % 31], c1 = 1 - 2 * xc[(i + N_id_2 + 3) % 31];
        const int z1m0 = 1 - 2 * xz[(i + (m0 % 8)) % 31], z1m1 = 1 - 2 * xz[(i + (m1 % 8)) % 31];
        s0[2 * i] = (int8_t)(s0m0 * c0); s0[2 * i + 1] = (int8_t)(s1m1 * c1 * z1m0);
        s5[2 * i] = (int8_t)(s1m1 * c0); s5[2 * i + 1] = (int8_t)(s0m0 * c1 * z1m1);
    }
}

extern "C" int op_find_sss(const float *is, const float *qs, int fs_idx, int N_rb_dl, uint32_t N_id_2, uint32_t *symb_starts, float pss_thresh,
                           uint32_t *N_id_1, uint32_t *frame_start_idx)
{
    const FsP p = fsp(fs_idx);
    const int nsc = 12 * N_rb_dl, c0 = nsc / 2 - 31;
    const float thresh = pss_thresh * 0.9;
    std::vector<float> rr(1200), ri(1200);
    op_samples_to_symbols_dl(is, qs, fs_idx, N_rb_dl, symb_starts[5], 0, rr.data(), ri.data());
    const uint32_t back = (p.N + p.cpe) * 4 + p.N + p.cp0;
    for (int i = 0; i < 168; i++) {
        int8_t s0[62], s5[62];
        gen_sss(i, N_id_2, s0, s5);
        for (int v = 0; v < 2; v++) {
            const int8_t *s = v ? s5 : s0;
            float cr = 0, ci = 0;
            for (int j = 0; j < 62; j++) { cr += rr[c0 + j] * (float)s[j] + ri[c0 + j] * 0.0f; ci += rr[c0 + j] * 0.0f - ri[c0 + j] * (float)s[j]; }
            if (sqrtf(cr * cr + ci * ci) > thresh) {
                const uint32_t need = back + (v ? (uint32_t)p.slot * 10 : 0);
                while (symb_starts[5] < need) symb_starts[5] += p.frame;
                *N_id_1 = i;
                *frame_start_idx = symb_starts[5] - need;
                return 0;
            }
        }
    }
    return -1;
}

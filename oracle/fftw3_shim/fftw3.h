/* TEST INFRASTRUCTURE ONLY -- not part of the product path.
 *
 * Minimal stand-in for the FFTW3 single-precision API so that the reference's
 * own liblte_phy.cc (which #includes "fftw3.h", /root/reference/liblte/hdr/liblte_phy.h:92)
 * can be compiled in a container that has no FFTW.  Only the five entry points
 * liblte_phy.cc uses are provided (plan creation liblte_phy.cc:2311-2330,
 * execution liblte_phy.cc:8521/8624).  The transform is a single-precision
 * Stockham autosort radix-4 (+ one radix-2 pass when log2(n) is odd) with
 * twiddles tabulated from double-precision sin/cos -- i.e. "any correct
 * unnormalised DFT with f32 I/O", which is all the reference assumes of FFTW
 * (the FFTW version is unpinned there).  Non power-of-two sizes (only planned
 * by liblte_phy_ul_init, never on the DL receive path) fall back to a plain
 * O(n^2) DFT in double.
 */
#ifndef ORACLE_FFTW3_SHIM_H
#define ORACLE_FFTW3_SHIM_H

#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef float fftwf_complex[2];

#define FFTW_FORWARD  (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE  (0U)
#define FFTW_ESTIMATE (1U << 6)

struct oracle_fft_plan_s {
    int            n;
    int            sign;
    int            pow2;
    fftwf_complex *in;
    fftwf_complex *out;
    float         *tw;      /* n complex twiddles e^{sign*2*pi*i*k/n} */
    float         *scratch; /* n complex */
};
typedef struct oracle_fft_plan_s *fftwf_plan;

static inline void *fftwf_malloc(size_t n) { void *p = NULL; if (posix_memalign(&p, 64, n ? n : 64)) return NULL; return p; }
static inline void  fftwf_free(void *p) { free(p); }

static inline fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags)
{
    (void)flags;
    fftwf_plan p = (fftwf_plan)malloc(sizeof(*p));
    p->n = n; p->sign = sign; p->in = in; p->out = out;
    p->pow2 = (n > 0) && ((n & (n - 1)) == 0);
    p->tw      = (float *)fftwf_malloc(sizeof(float) * 2 * (size_t)n);
    p->scratch = (float *)fftwf_malloc(sizeof(float) * 2 * (size_t)n);
    for (int k = 0; k < n; k++) {
        double a = (double)sign * 2.0 * M_PI * (double)k / (double)n;
        p->tw[2 * k + 0] = (float)cos(a);
        p->tw[2 * k + 1] = (float)sin(a);
    }
    return p;
}

static inline void fftwf_destroy_plan(fftwf_plan p)
{
    if (p) { fftwf_free(p->tw); fftwf_free(p->scratch); free(p); }
}

/* One Stockham radix-4 pass: n = 4*l*m ; x[(4j+q)*m + k]... (decimation in frequency) */
static inline void oracle_fft_pass4(int n, int s, int sign, const float *tw, const float *x, float *y)
{
    /* n: remaining transform length at this pass, s: stride (number of interleaved transforms) */
    const int n1 = n / 4;
    const int tstep = s; /* twiddle index step relative to the full-size table: full N = n*s */
    for (int p = 0; p < n1; p++) {
        const float w1r = tw[2 * (1 * p * tstep) + 0], w1i = tw[2 * (1 * p * tstep) + 1];
        const float w2r = tw[2 * (2 * p * tstep) + 0], w2i = tw[2 * (2 * p * tstep) + 1];
        const float w3r = tw[2 * (3 * p * tstep) + 0], w3i = tw[2 * (3 * p * tstep) + 1];
        const float *x0 = x + 2 * (size_t)s * (p);
        const float *x1 = x + 2 * (size_t)s * (p + n1);
        const float *x2 = x + 2 * (size_t)s * (p + 2 * n1);
        const float *x3 = x + 2 * (size_t)s * (p + 3 * n1);
        float *y0 = y + 2 * (size_t)s * (4 * p + 0);
        float *y1 = y + 2 * (size_t)s * (4 * p + 1);
        float *y2 = y + 2 * (size_t)s * (4 * p + 2);
        float *y3 = y + 2 * (size_t)s * (4 * p + 3);
        for (int q = 0; q < s; q++) {
            const float ar = x0[2 * q], ai = x0[2 * q + 1];
            const float br = x1[2 * q], bi = x1[2 * q + 1];
            const float cr = x2[2 * q], ci = x2[2 * q + 1];
            const float dr = x3[2 * q], di = x3[2 * q + 1];
            const float apcr = ar + cr, apci = ai + ci;
            const float amcr = ar - cr, amci = ai - ci;
            const float bpdr = br + dr, bpdi = bi + di;
            /* sign*j*(b-d) */
            float jr, ji;
            if (sign < 0) { jr =  (bi - di); ji = -(br - dr); }
            else          { jr = -(bi - di); ji =  (br - dr); }
            y0[2 * q] = apcr + bpdr;           y0[2 * q + 1] = apci + bpdi;
            float t1r = amcr + jr, t1i = amci + ji;
            float t2r = apcr - bpdr, t2i = apci - bpdi;
            float t3r = amcr - jr, t3i = amci - ji;
            y1[2 * q] = t1r * w1r - t1i * w1i; y1[2 * q + 1] = t1r * w1i + t1i * w1r;
            y2[2 * q] = t2r * w2r - t2i * w2i; y2[2 * q + 1] = t2r * w2i + t2i * w2r;
            y3[2 * q] = t3r * w3r - t3i * w3i; y3[2 * q + 1] = t3r * w3i + t3i * w3r;
        }
    }
}

static inline void oracle_fft_pass2(int n, int s, const float *tw, const float *x, float *y)
{
    const int n1 = n / 2;
    for (int p = 0; p < n1; p++) {
        const float wr = tw[2 * (p * s) + 0], wi = tw[2 * (p * s) + 1];
        const float *x0 = x + 2 * (size_t)s * p;
        const float *x1 = x + 2 * (size_t)s * (p + n1);
        float *y0 = y + 2 * (size_t)s * (2 * p + 0);
        float *y1 = y + 2 * (size_t)s * (2 * p + 1);
        for (int q = 0; q < s; q++) {
            const float ar = x0[2 * q], ai = x0[2 * q + 1];
            const float br = x1[2 * q], bi = x1[2 * q + 1];
            y0[2 * q] = ar + br; y0[2 * q + 1] = ai + bi;
            const float tr = ar - br, ti = ai - bi;
            y1[2 * q] = tr * wr - ti * wi; y1[2 * q + 1] = tr * wi + ti * wr;
        }
    }
}

static inline void fftwf_execute(const fftwf_plan p)
{
    const int n = p->n;
    if (!p->pow2) {
        /* direct DFT in double; never hit on the DL receive path */
        float *tmp = p->scratch;
        for (int k = 0; k < n; k++) {
            double sr = 0, si = 0;
            for (int t = 0; t < n; t++) {
                double a = (double)p->sign * 2.0 * M_PI * (double)(((long long)k * t) % n) / (double)n;
                double c = cos(a), s = sin(a);
                sr += p->in[t][0] * c - p->in[t][1] * s;
                si += p->in[t][0] * s + p->in[t][1] * c;
            }
            tmp[2 * k] = (float)sr; tmp[2 * k + 1] = (float)si;
        }
        memcpy(p->out, tmp, sizeof(float) * 2 * (size_t)n);
        return;
    }
    if (n == 1) { p->out[0][0] = p->in[0][0]; p->out[0][1] = p->in[0][1]; return; }
    /* ping-pong between out and scratch so the last pass lands in out */
    int npass = 0; { int m = n; while (m > 1) { if (m % 4 == 0) m /= 4; else m /= 2; npass++; } }
    const float *src = (const float *)p->in;
    float *bufs[2] = { (float *)p->out, p->scratch };
    int which = (npass % 2 == 1) ? 0 : 1; /* first destination */
    int m = n, s = 1;
    while (m > 1) {
        float *dst = bufs[which];
        if (m % 4 == 0) { oracle_fft_pass4(m, s, p->sign, p->tw, src, dst); m /= 4; s *= 4; }
        else            { oracle_fft_pass2(m, s, p->tw, src, dst);          m /= 2; s *= 2; }
        src = dst; which ^= 1;
    }
}

#endif

/* TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's turbo codec.
 *   turbo_decode            /root/reference/liblte/src/liblte_phy.cc:10620-10845
 *   viterbi_decode_siso     liblte_phy.cc:10341-10529
 *   conv_encode_soft        liblte_phy.cc:10070-10151
 *   interleaver (int8)      liblte_phy.cc:10960-10985, deinterleaver :11047-11072
 *   turbo_encode            liblte_phy.cc:10541-10586, constituent encoder :10855-10925
 * Written from the algorithm description, not transcribed: the SISO keeps 4 predecessor-compare bits per
 * step instead of the reference's full metric history, uses the state-independent step weight, and the
 * deinterleaver scatter is expressed through a "last writer" inverse map.
 */
#include "oracle_port.h"
#include "../../openlte_b200/csrc/qpp_table.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

extern "C" int op_qpp_params(int K, uint32_t *f1, uint32_t *f2)
{
    for (int r = 0; r < LTE_QPP_TABLE_ROWS; r++)
        if (LTE_QPP_TABLE[r][0] == K) { *f1 = LTE_QPP_TABLE[r][1]; *f2 = LTE_QPP_TABLE[r][2]; return 0; }
    *f1 = 0; *f2 = 0;           /* the reference leaves f1 = f2 = 0 for an unknown size: pi(i) = 0 */
    return -1;
}

/* pi(i) = ((f1*i + f2*i*i) mod 2^32) mod K  -- the uint32 wrap is the reference's behaviour.
 * op_set_qpp_spec(1) switches every function of this port that interleaves to the permutation 3GPP TS 36.212 5.1.3.2.3 defines,
 * (f1*i + f2*i*i) mod K without the wrap: the twin of the library's LTEB200_QPP_SPEC option (no reference counterpart). */
static int g_qpp_spec = 0;
extern "C" void op_set_qpp_spec(int on) { g_qpp_spec = on ? 1 : 0; }
extern "C" void op_qpp_pi(int K, uint32_t *pi)
{
    uint32_t f1, f2;
    op_qpp_params(K, &f1, &f2);
    for (uint32_t i = 0; i < (uint32_t)K; i++)
        pi[i] = g_qpp_spec ? (uint32_t)(((uint64_t)f1 * i + (uint64_t)f2 * i * i) % (uint64_t)K) : (uint32_t)(f1 * i + f2 * i * i) % (uint32_t)K;
}

/* One RSC constituent encoder run: feedback 1+D^2+D^3, feed-forward 1+D+D^3, then 4 termination steps in
 * which the register input is forced to 0.  z[] gets K+4 parity bits, fb[] the K+4 feedback bits. */
static void rsc(const uint8_t *in, int K, uint8_t *z, uint8_t *fb)
{
    int s1 = 0, s2 = 0, s3 = 0;                   /* D, D^2, D^3 */
    for (int i = 0; i < K + 4; i++) {
        int f = s2 ^ s3;
        int a = (i < K) ? ((in[i] + f) & 1) : 0;  /* parity of (fb + in); a filler value of 100 counts as 0 */
        fb[i] = (uint8_t)f;
        z[i]  = (uint8_t)(a ^ s1 ^ s3);
        s3 = s2; s2 = s1; s1 = a;
    }
}

extern "C" void op_turbo_encode(const uint8_t *c, int K, uint8_t *d)
{
    const int D = K + 4;
    std::vector<uint8_t> z(D), fb(D), zp(D), xp(D), cp(K);
    std::vector<uint32_t> pi(K);
    op_qpp_pi(K, pi.data());
    rsc(c, K, z.data(), fb.data());
    for (int i = 0; i < K; i++) cp[i] = c[pi[i]];
    rsc(cp.data(), K, zp.data(), xp.data());
    for (int i = 0; i < K; i++) { d[i] = c[i]; d[D + i] = z[i]; d[2 * D + i] = zp[i]; }
    /* tail multiplexing, liblte_phy.cc:10575-10586 */
    d[K + 0] = fb[K];         d[K + 1] = z[K + 1];          d[K + 2] = xp[K];             d[K + 3] = zp[K + 1];
    d[D + K + 0] = z[K];      d[D + K + 1] = fb[K + 2];     d[D + K + 2] = zp[K];         d[D + K + 3] = xp[K + 2];
    d[2 * D + K + 0] = fb[K + 1]; d[2 * D + K + 1] = z[K + 2]; d[2 * D + K + 2] = xp[K + 1]; d[2 * D + K + 3] = zp[K + 2];
}

/* ------------------------------------------------------------------------------------------------------
 * SISO "Viterbi" of the reference.  State j in 0..7, predecessors 2j mod 8 and 2j+1 mod 8.  Grouped as four
 * butterflies m = 0..3: predecessors (2m, 2m+1) feed successors (m, m+4).  With sa = +1 if a >= 0 else -1
 * (same for b), A = sa+sb, B = sb-sa, the per-butterfly branch value is M_m = {-A, B, -B, A}[m]:
 *     successor m   : from 2m   with  +M,  from 2m+1 with  -M
 *     successor m+4 : from 2m   with  -M,  from 2m+1 with  +M
 * (from the register/generator definition at liblte_phy.cc:10379-10408).  The second predecessor wins iff
 * its UNWEIGHTED sum is strictly smaller; the WEIGHTED sum (W = |a|+|b|) is what is stored (10452-10463).
 * All metrics are integers of magnitude < 2^24 so int32 arithmetic equals the reference's float arithmetic.
 */
extern "C" void op_siso(const int8_t *a, const int8_t *b, int K, int8_t *act)
{
    int32_t P[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    std::vector<uint8_t> cmp(K);     /* bit m: metric[2m] > metric[2m+1] at time i (what the traceback consults) */
    int Wmax = 0;
    for (int i = 0; i < K; i++) {
        const int sa = a[i] >= 0 ? 1 : -1, sb = b[i] >= 0 ? 1 : -1;
        const int W = abs((int)a[i]) + abs((int)b[i]);
        if (W > Wmax) Wmax = W;
        const int A = sa + sb, B = sb - sa;
        const int M[4] = {-A, B, -B, A};
        int32_t N[8]; uint8_t cb = 0;
        for (int m = 0; m < 4; m++) {
            const int32_t p0 = P[2 * m], p1 = P[2 * m + 1];
            if (p0 > p1) cb |= (uint8_t)(1u << m);
            N[m]     = (p0 + M[m] > p1 - M[m]) ? p1 - W * M[m] : p0 + W * M[m];
            N[m + 4] = (p0 - M[m] > p1 + M[m]) ? p1 + W * M[m] : p0 - W * M[m];
        }
        cmp[i] = cb;
        memcpy(P, N, sizeof(P));
    }
    /* end state: first strict minimum (liblte_phy.cc:10467-10481) */
    int s = 0; { float best = 1000000.0f; for (int j = 0; j < 8; j++) if ((float)P[j] < best) { best = (float)P[j]; s = j; } }
    const float wmax = (float)Wmax;
    for (int i = K - 1; i >= 0; i--) {
        /* transition into state s at time i+1: bit = MSB of s, soft = trunc(127 * (W_i / Wmax)) */
        const float w = (float)(abs((int)a[i]) + abs((int)b[i]));
        const int8_t mag = (int8_t)(127 * (w / wmax));      /* NaN (0/0) -> 0 as on x86 */
        act[i] = (s & 4) ? (int8_t)(-mag) : mag;
        const int m = s & 3;
        s = 2 * m + ((cmp[i] >> m) & 1);
    }
}

/* soft XOR: magnitude (|x|+|y|)>>1, negative iff exactly one operand is negative (liblte_phy.cc:10123-10147) */
static inline int8_t sxor(int x, int y)
{
    int mag = (abs(x) + abs(y)) >> 1;
    return (int8_t)(((x < 0) != (y < 0)) ? -mag : mag);
}
/* fb[0] = 127, fb[i+1] = sxor(act[i-1], act[i-2]) with act[<0] = +127 (conv_encode_soft g=03, liblte_phy.cc:10676-10685) */
static void feedback(const int8_t *act, int K, int8_t *fb)
{
    fb[0] = 127;
    for (int i = 0; i + 1 < K; i++) fb[i + 1] = sxor(i >= 1 ? act[i - 1] : 127, i >= 2 ? act[i - 2] : 127);
}
/* the three "merge" variants (liblte_phy.cc:10688-10707, 10777-10797, 10800-10820); `alt` is act1 for step 10 */
static inline int8_t merge(int act, int fb, int variant, int alt)
{
    if (act >= 0 && fb >= 0) return (int8_t)((act + fb) >> 1);
    if (act <  0 && fb <  0) return (int8_t)((-act - fb) >> 1);
    if (act >= 0 && fb <  0) return (int8_t)(-(((variant == 1 ? alt : act) - fb) >> 1));
    if (variant == 1) return (int8_t)(-((-alt + fb) >> 1));
    if (variant == 2) return (int8_t)(-((-act - fb) >> 1));
    return (int8_t)(-((-act + fb) >> 1));
}

extern "C" void op_turbo_decode_ref(float *d, int K, uint8_t *out, int8_t *dbg)
{
    const int D = K + 4;
    for (int i = 0; i < 3 * D; i++) if (d[i] == OP_RX_NULL) d[i] = 0;
    float maxv = 0;
    for (int i = 0; i < K; i++) for (int x = 0; x < 3; x++) { float v = fabsf(d[3 * i + x]); if (v > maxv) maxv = v; }
    std::vector<int8_t> q0(K), q1(K), q2(K), act1(K), fb(K), calc1(K), sysI(K), c1I(K), actA(K), actB(K), calcA(K), calcB(K),
        calc2(K, 0), calc3(K, 0);
    for (int i = 0; i < K; i++) {
        q0[i] = (int8_t)(d[3 * i + 0] * 127 / maxv);
        q1[i] = (int8_t)(d[3 * i + 1] * 127 / maxv);
        q2[i] = (int8_t)(d[3 * i + 2] * 127 / maxv);
    }
    std::vector<uint32_t> pi(K);
    op_qpp_pi(K, pi.data());
    op_siso(q1.data(), q0.data(), K, act1.data());
    feedback(act1.data(), K, fb.data());
    for (int i = 0; i < K; i++) calc1[i] = merge(act1[i], fb[i], 0, 0);
    for (int i = 0; i < K; i++) { sysI[i] = q0[pi[i]]; c1I[i] = calc1[pi[i]]; }
    op_siso(q2.data(), sysI.data(), K, actA.data());
    op_siso(q2.data(), c1I.data(), K, actB.data());
    feedback(actA.data(), K, fb.data());
    for (int i = 0; i < K; i++) calcA[i] = merge(actA[i], fb[i], 1, act1[i]);
    feedback(actB.data(), K, fb.data());
    for (int i = 0; i < K; i++) calcB[i] = merge(actB[i], fb[i], 2, 0);
    for (int i = 0; i < K; i++) { calc2[pi[i]] = calcA[i]; calc3[pi[i]] = calcB[i]; }   /* holes stay 0 (H12) */
    for (int i = 0; i < K; i++) out[i] = ((int)q0[i] + calc1[i] + calc2[i] + calc3[i] >= 0) ? 0 : 1;
    if (dbg) {
        memcpy(dbg + 0 * K, act1.data(), K);  memcpy(dbg + 1 * K, calc1.data(), K); memcpy(dbg + 2 * K, actA.data(), K);
        memcpy(dbg + 3 * K, actB.data(), K);  memcpy(dbg + 4 * K, calc2.data(), K); memcpy(dbg + 5 * K, calc3.data(), K);
    }
}

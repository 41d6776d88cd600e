/* TEST INFRASTRUCTURE ONLY.  CPU twin of the "mode B" turbo decoder (openlte_b200/csrc/turbo_mlm.cu).
 *
 * The reference has no max-log-MAP decoder: this is the north-star decoder named in
 * BASELINE.json, defined here in integer arithmetic so that the CUDA kernel can be checked bit for bit.
 * It decodes the reference's own code: constituent RSC g0 = 1+D^2+D^3 (feedback), g1 = 1+D+D^3
 * (liblte_phy.cc:10863-10865), the reference's tail multiplexing (liblte_phy.cc:10575-10586) and the
 * reference's interleaver INCLUDING its uint32 wrap-around (liblte_phy.cc:10982), so that on
 * inputs both decode correctly its hard bits equal turbo_decode()'s.
 *
 * Definition (16-bit fixed point: every intermediate provably fits int16, which is what lets the kernel use
 * packed 16x2 arithmetic; this twin computes in int and ASSERTS the range; ">>" arithmetic):
 *   quantisation   q(v) = trunc(v*127/maxv), maxv = max |d| over all 3(K+4) inputs, sentinel 10000 -> 0
 *   branch metric  gamma(u,p) = (u==0 ? Lsa : 0) + (p==0 ? Lp : 0),  Lsa = Lsys + La   (positive LLR = bit 0)
 *   recursions     alpha'(s') = max_s(alpha(s) - r + gamma),  beta(s) = max_s'(gamma - r' + beta'(s'))
 *                  r = alpha(0) on the steps i with i mod 4 == 0 (else 0), r' = beta'(0) on the steps with i mod 4 == 3 (else 0),
 *                  i = index inside the window: the state vector is re-referenced to its state 0 every 4th step, which bounds
 *                  every value by 3831 + 4*1277 (state 0 is always reachable)
 *   LLR            L = max_{u=0}(alpha+gamma+beta') - max_{u=1}(alpha+gamma+beta')   (common offsets cancel)
 *   extrinsic      Le = clamp(((L - Lsa)*3) >> 2, +-1023)          (0.75-scaled max-log-MAP)
 *   windows        length WIN = 16; a window's start alpha / end beta are the values the neighbouring window produced
 *                  in the PREVIOUS iteration (state-metric propagation); first iteration: all-zero boundaries.
 *                  Block start: state 0 known (others NEG = -8000).  Block end: beta from the three tail steps,
 *                  end state 0 (others NEG), floored at NEG.
 *   exchange       La2[i] = Le1[pi(i)];  La1[j] = clamp(sum_{i: pi(i)=j} Le2[i], +-1023) (0 if no such i)
 *   decision       bit = (Ls + La1 + Le1 >= 0) ? 0 : 1 after n_iter iterations
 */
#include "oracle_port.h"
#include <assert.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <algorithm>

namespace {
#ifndef MLM_WIN
#define MLM_WIN 16
#endif
const int WIN = MLM_WIN;
const int NEG = -8000;
const int LE_LIM = 1023, LA_LIM = 1023;

inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
inline int chk16(int v) { assert(v >= -32768 && v <= 32767); return v; }

/* trellis: state S = s1*4+s2*2+s3.  input u: a = u^s2^s3, parity p = a^s1^s3, next = (a<<2)|(S>>1) */
struct Trellis {
    int nxt[8][2], par[8][2];
    Trellis()
    {
        for (int S = 0; S < 8; S++)
            for (int u = 0; u < 2; u++) {
                int s1 = (S >> 2) & 1, s2 = (S >> 1) & 1, s3 = S & 1;
                int a = u ^ s2 ^ s3;
                par[S][u] = a ^ s1 ^ s3;
                nxt[S][u] = (a << 2) | (S >> 1);
            }
    }
};
const Trellis T;

inline int gam(int u, int p, int Lsa, int Lp) { return (u == 0 ? Lsa : 0) + (p == 0 ? Lp : 0); }

/* one constituent decoder pass over the whole block, window by window */
void dec_pass(int K, const int *Lsa, const int8_t *Lp, const int xt[3], const int zt[3], int *bndA /*[nw][8]*/,
              int *bndB /*[nw][8]*/, int *Le)
{
    const int nw = (K + WIN - 1) / WIN;
    std::vector<int> newA((size_t)nw * 8), newB((size_t)nw * 8);
    for (int w = 0; w < nw; w++) {
        const int k0 = w * WIN, k1 = std::min(K, k0 + WIN), n = k1 - k0;
        int a[8], b[8];
        int st[64][8];
        if (w == 0) { a[0] = 0; for (int s = 1; s < 8; s++) a[s] = NEG; }
        else for (int s = 0; s < 8; s++) a[s] = bndA[w * 8 + s];
        if (w == nw - 1) {
            /* beta at time K from the tail: three forced steps (u = feedback) ending in state 0 */
            int bt[8]; bt[0] = 0; for (int s = 1; s < 8; s++) bt[s] = NEG;
            for (int t = 2; t >= 0; t--) {
                int nb[8];
                for (int S = 0; S < 8; S++) {
                    int u = ((S >> 1) & 1) ^ (S & 1);      /* termination input makes a = 0 */
                    int v = gam(u, T.par[S][u], xt[t], zt[t]) - bt[0] + bt[T.nxt[S][u]];
                    nb[S] = chk16(v < NEG ? NEG : v);
                }
                memcpy(bt, nb, sizeof(bt));
            }
            for (int s = 0; s < 8; s++) b[s] = bt[s];
        } else for (int s = 0; s < 8; s++) b[s] = bndB[w * 8 + s];

        /* forward */
        for (int i = 0; i < n; i++) {
            for (int s = 0; s < 8; s++) st[i][s] = a[s];
            int na[8]; for (int s = 0; s < 8; s++) na[s] = INT32_MIN;
            const int ls = Lsa[k0 + i], lp = Lp[k0 + i], a0 = ((i & 3) == 0) ? a[0] : 0;   /* re-reference on every 4th step */
            for (int S = 0; S < 8; S++)
                for (int u = 0; u < 2; u++) {
                    int v = chk16(a[S] + chk16(gam(u, T.par[S][u], ls, lp) - a0));
                    int d = T.nxt[S][u];
                    if (v > na[d]) na[d] = v;
                }
            memcpy(a, na, sizeof(a));
        }
        if (w + 1 < nw) for (int s = 0; s < 8; s++) newA[(w + 1) * 8 + s] = a[s];   /* seeds window w+1 in the next iteration */
        /* backward + LLR */
        for (int i = n - 1; i >= 0; i--) {
            const int ls = Lsa[k0 + i], lp = Lp[k0 + i], b0 = ((i & 3) == 3) ? b[0] : 0;
            int m0 = INT32_MIN, m1 = INT32_MIN, nb[8];
            for (int S = 0; S < 8; S++) {
                int best = INT32_MIN;
                for (int u = 0; u < 2; u++) {
                    int x = chk16(chk16(gam(u, T.par[S][u], ls, lp) - b0) + b[T.nxt[S][u]]);
                    if (x > best) best = x;
                    int tot = chk16(st[i][S] + x);
                    if (u == 0) { if (tot > m0) m0 = tot; } else { if (tot > m1) m1 = tot; }
                }
                nb[S] = best;
            }
            memcpy(b, nb, sizeof(b));
            const int L = m0 - m1;
            Le[k0 + i] = clampi(((L - ls) * 3) >> 2, -LE_LIM, LE_LIM);
        }
        if (w > 0) for (int s = 0; s < 8; s++) newB[(w - 1) * 8 + s] = b[s];
    }
    for (int w = 1; w < nw; w++) memcpy(&bndA[w * 8], &newA[w * 8], 8 * sizeof(int));
    for (int w = 0; w + 1 < nw; w++) memcpy(&bndB[w * 8], &newB[w * 8], 8 * sizeof(int));
}
}  // namespace

extern "C" void op_turbo_decode_mlm(const float *d, int K, int n_iter, uint8_t *out)
{
    const int D = K + 4;
    float maxv = 0;
    std::vector<float> dd(3 * D);
    for (int i = 0; i < 3 * D; i++) { float v = d[i]; if (v == OP_RX_NULL) v = 0; dd[i] = v; maxv = fmaxf(maxv, fabsf(v)); }
    auto q = [&](float v) -> int { float r = v * 127 / maxv; return (r != r) ? 0 : (int)r; };
    std::vector<int8_t> Ls(K), Lp1(K), Lp2(K), LsI(K);
    for (int i = 0; i < K; i++) { Ls[i] = (int8_t)q(dd[3 * i]); Lp1[i] = (int8_t)q(dd[3 * i + 1]); Lp2[i] = (int8_t)q(dd[3 * i + 2]); }
    auto D0 = [&](int i) { return q(dd[3 * i + 0]); };
    auto D1 = [&](int i) { return q(dd[3 * i + 1]); };
    auto D2 = [&](int i) { return q(dd[3 * i + 2]); };
    /* tail multiplexing, liblte_phy.cc:10575-10586 */
    const int xt1[3] = {D0(K), D2(K), D1(K + 1)}, zt1[3] = {D1(K), D0(K + 1), D2(K + 1)};
    const int xt2[3] = {D0(K + 2), D2(K + 2), D1(K + 3)}, zt2[3] = {D1(K + 2), D0(K + 3), D2(K + 3)};
    std::vector<uint32_t> pi(K);
    op_qpp_pi(K, pi.data());
    for (int i = 0; i < K; i++) LsI[i] = Ls[pi[i]];
    const int nw = (K + WIN - 1) / WIN;
    std::vector<int> bA1((size_t)nw * 8, 0), bB1((size_t)nw * 8, 0), bA2((size_t)nw * 8, 0), bB2((size_t)nw * 8, 0);
    std::vector<int> La1(K, 0), Le1(K, 0), Le2(K, 0), Lsa(K);
    for (int it = 0; it < n_iter; it++) {
        for (int i = 0; i < K; i++) Lsa[i] = Ls[i] + La1[i];
        dec_pass(K, Lsa.data(), Lp1.data(), xt1, zt1, bA1.data(), bB1.data(), Le1.data());
        for (int i = 0; i < K; i++) Lsa[i] = LsI[i] + Le1[pi[i]];
        dec_pass(K, Lsa.data(), Lp2.data(), xt2, zt2, bA2.data(), bB2.data(), Le2.data());
        std::vector<int> acc(K, 0);
        for (int i = 0; i < K; i++) acc[pi[i]] += Le2[i];
        for (int j = 0; j < K; j++) La1[j] = clampi(acc[j], -LA_LIM, LA_LIM);
    }
    for (int j = 0; j < K; j++) out[j] = (Ls[j] + La1[j] + Le1[j] >= 0) ? 0 : 1;
}

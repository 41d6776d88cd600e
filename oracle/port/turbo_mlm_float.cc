/* TEST INFRASTRUCTURE ONLY.  Independent floating-point max-log-MAP turbo decoder: FULL-BLOCK forward/backward
 * recursions (no windows, no state-metric propagation, no fixed-point quantisation, no clamps), written from the code
 * definition alone -- constituent RSC g0 = 1+D^2+D^3 (feedback), g1 = 1+D+D^3 (liblte_phy.cc:10863-10865), trellis
 * termination and tail multiplexing of liblte_phy.cc:10575-10586, the reference's interleaver including its uint32
 * wrap (liblte_phy.cc:10982).  It shares no code and no numeric design with oracle/port/turbo_mlm.cc / the CUDA kernels:
 * the tests use it to bound what the 16-step-window int16 decoder loses (block error rate vs noise level).
 *
 * LLR convention of the reference: positive = bit 0.  Extrinsic scaling 0.75 (the usual max-log correction).
 */
#include "oracle_port.h"
#include <math.h>
#include <string.h>
#include <vector>

namespace {
struct Tr { int nxt[8][2], par[8][2]; };
Tr make_trellis()
{
    Tr t;
    for (int S = 0; S < 8; S++)
        for (int u = 0; u < 2; u++) {
            const int s1 = (S >> 2) & 1, s2 = (S >> 1) & 1, s3 = S & 1;
            const int a = u ^ s2 ^ s3;            /* feedback 1 + D^2 + D^3 */
            t.par[S][u] = a ^ s1 ^ s3;            /* feed-forward 1 + D + D^3 */
            t.nxt[S][u] = (a << 2) | (S >> 1);
        }
    return t;
}
const Tr T = make_trellis();
const double NEGINF = -1e30;

/* one constituent decoder over the whole block: returns the extrinsic (already scaled) per step */
void siso_full(int K, const double *Lsa, const double *Lp, const double xt[3], const double zt[3], double *Le)
{
    std::vector<double> alpha((size_t)(K + 1) * 8, NEGINF), beta((size_t)(K + 4) * 8, NEGINF);
    alpha[0] = 0.0;
    for (int k = 0; k < K; k++) {
        const double *a = &alpha[(size_t)k * 8];
        double *n = &alpha[(size_t)(k + 1) * 8];
        for (int S = 0; S < 8; S++) {
            if (a[S] <= NEGINF / 2) continue;
            for (int u = 0; u < 2; u++) {
                const double g = (u == 0 ? Lsa[k] : 0.0) + (T.par[S][u] == 0 ? Lp[k] : 0.0);
                const int d = T.nxt[S][u];
                if (a[S] + g > n[d]) n[d] = a[S] + g;
            }
        }
        double m = n[0];
        for (int S = 1; S < 8; S++) if (n[S] > m) m = n[S];
        for (int S = 0; S < 8; S++) if (n[S] > NEGINF / 2) n[S] -= m;
    }
    /* termination: three forced steps (input = feedback, register input 0) ending in state 0 */
    beta[(size_t)(K + 3) * 8 + 0] = 0.0;
    for (int k = K + 2; k >= K; k--) {
        const double *bn = &beta[(size_t)(k + 1) * 8];
        double *b = &beta[(size_t)k * 8];
        for (int S = 0; S < 8; S++) {
            const int u = ((S >> 1) & 1) ^ (S & 1);
            const int d = T.nxt[S][u];
            if (bn[d] <= NEGINF / 2) continue;
            b[S] = (u == 0 ? xt[k - K] : 0.0) + (T.par[S][u] == 0 ? zt[k - K] : 0.0) + bn[d];
        }
    }
    for (int k = K - 1; k >= 0; k--) {
        const double *bn = &beta[(size_t)(k + 1) * 8], *a = &alpha[(size_t)k * 8];
        double *b = &beta[(size_t)k * 8];
        double m0 = NEGINF, m1 = NEGINF;
        for (int S = 0; S < 8; S++) {
            double best = NEGINF;
            for (int u = 0; u < 2; u++) {
                const int d = T.nxt[S][u];
                if (bn[d] <= NEGINF / 2) continue;
                const double x = (u == 0 ? Lsa[k] : 0.0) + (T.par[S][u] == 0 ? Lp[k] : 0.0) + bn[d];
                if (x > best) best = x;
                if (a[S] > NEGINF / 2) { const double tot = a[S] + x; if (u == 0) { if (tot > m0) m0 = tot; } else { if (tot > m1) m1 = tot; } }
            }
            b[S] = best;
        }
        double m = b[0];
        for (int S = 1; S < 8; S++) if (b[S] > m) m = b[S];
        for (int S = 0; S < 8; S++) if (b[S] > NEGINF / 2) b[S] -= m;
        Le[k] = 0.75 * ((m0 - m1) - Lsa[k]);
    }
}
}  // namespace

/* d: interleaved d[3i+x] (sentinel 10000 = not received -> 0), out: K hard bits */
extern "C" void op_turbo_decode_mlm_float(const float *d, int K, int n_iter, uint8_t *out)
{
    const int D = K + 4;
    std::vector<double> v(3 * D);
    for (int i = 0; i < 3 * D; i++) v[i] = d[i] == OP_RX_NULL ? 0.0 : (double)d[i];
    std::vector<double> Ls(K), Lp1(K), Lp2(K), LsI(K), Lsa(K), Le1(K, 0.0), Le2(K, 0.0), La1(K, 0.0);
    for (int i = 0; i < K; i++) { Ls[i] = v[3 * i]; Lp1[i] = v[3 * i + 1]; Lp2[i] = v[3 * i + 2]; }
    auto D0 = [&](int i) { return v[3 * i]; };
    auto D1 = [&](int i) { return v[3 * i + 1]; };
    auto D2 = [&](int i) { return v[3 * i + 2]; };
    const double xt1[3] = {D0(K), D2(K), D1(K + 1)}, zt1[3] = {D1(K), D0(K + 1), D2(K + 1)};
    const double xt2[3] = {D0(K + 2), D2(K + 2), D1(K + 3)}, zt2[3] = {D1(K + 2), D0(K + 3), D2(K + 3)};
    std::vector<uint32_t> pi(K);
    op_qpp_pi(K, pi.data());
    for (int i = 0; i < K; i++) LsI[i] = Ls[pi[i]];
    for (int it = 0; it < n_iter; it++) {
        for (int i = 0; i < K; i++) Lsa[i] = Ls[i] + La1[i];
        siso_full(K, Lsa.data(), Lp1.data(), xt1, zt1, Le1.data());
        for (int i = 0; i < K; i++) Lsa[i] = LsI[i] + Le1[pi[i]];
        siso_full(K, Lsa.data(), Lp2.data(), xt2, zt2, Le2.data());
        std::fill(La1.begin(), La1.end(), 0.0);
        for (int i = 0; i < K; i++) La1[pi[i]] += Le2[i];       /* the wrapped interleaver maps 0..3 positions onto one */
    }
    for (int j = 0; j < K; j++) out[j] = (Ls[j] + La1[j] + Le1[j] >= 0.0) ? 0 : 1;
}

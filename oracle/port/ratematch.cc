/* TEST INFRASTRUCTURE ONLY.  CPU restatement of turbo rate matching / unmatching.
 *   liblte_phy_rate_unmatch_turbo   /root/reference/liblte/src/liblte_phy.cc:11246-11490
 *   liblte_phy_rate_match_turbo     liblte_phy.cc:11081-11236
 * 3GPP TS 36.212 5.1.4.1: three sub-block interleavers (32 columns, inter-column permutation; the third
 * stream uses the shifted permutation), circular buffer w = [v0 | v1,v2 interlaced], read from k0.
 */
#include "oracle_port.h"
#include <math.h>
#include <vector>

static const int P32[32] = {0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30,
                            1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31};

namespace {
struct Geometry {
    int D, R, Kpi, ND, Kw, Ncb, k0;
    std::vector<int> src;   /* w position -> index into the (padded) stream-major input: x*Kpi + t, t in [0,Kpi) */
    Geometry(int D_, int C, int tx_mode, int N_soft, int M_dl_harq, int chan_type, int rv) : D(D_)
    {
        R = (D + 31) / 32; Kpi = 32 * R; ND = Kpi - D; Kw = 3 * Kpi;
        src.assign(Kw, 0);
        for (int n = 0; n < Kpi; n++) {
            const int col = n / R, row = n % R;
            const int t01 = 32 * row + P32[col];                 /* streams 0 and 1: column-permuted, read column-wise */
            const int t2  = (P32[col] + 32 * row + 1) % Kpi;     /* stream 2: liblte_phy.cc:11347 */
            src[n] = 0 * Kpi + t01;
            src[Kpi + 2 * n] = 1 * Kpi + t01;
            src[Kpi + 2 * n + 1] = 2 * Kpi + t2;
        }
        const int K_mimo = (tx_mode == 3 || tx_mode == 4 || tx_mode == 8 || tx_mode == 9) ? 2 : 1;
        const int N_ir = N_soft / (K_mimo * (M_dl_harq < 8 ? M_dl_harq : 8));
        if (chan_type == 0 || chan_type == 1) Ncb = (N_ir / C < Kw) ? N_ir / C : Kw; else Ncb = Kw;
        k0 = (int)(R * (2 * ceilf((float)Ncb / (float)(8 * R)) * rv + 2));
    }
    bool is_null(int pos) const { return (src[pos] % Kpi) < ND; }   /* padded head of each stream is <NULL> */
};
}

extern "C" int op_rate_unmatch_turbo(const float *e, int E, int D, int C, int tx_mode, int N_soft, int M_dl_harq,
                                     int chan_type, int rv, float *d_out)
{
    Geometry g(D, C, tx_mode, N_soft, M_dl_harq, chan_type, rv);
    std::vector<float> w(g.Kw, OP_RX_NULL);
    int k = 0;
    for (long j = 0; k < E; j++) {
        const int pos = (int)((g.k0 + j) % g.Ncb);
        if (g.is_null(pos)) continue;
        if (w[pos] == OP_RX_NULL) w[pos] = e[k];              /* first visit */
        else if (e[k] != OP_RX_NULL) w[pos] += e[k];          /* repeats are soft-combined by addition */
        k++;
    }
    for (int pos = 0; pos < g.Kw; pos++) {
        const int x = g.src[pos] / g.Kpi, t = g.src[pos] % g.Kpi;
        if (t >= g.ND) d_out[3 * (t - g.ND) + x] = w[pos];
    }
    return 3 * D;
}

extern "C" void op_rate_match_turbo(const uint8_t *d, int D, int C, int tx_mode, int N_soft, int M_dl_harq,
                                    int chan_type, int rv, int E, uint8_t *e)
{
    Geometry g(D, C, tx_mode, N_soft, M_dl_harq, chan_type, rv);
    int k = 0;
    for (long j = 0; k < E; j++) {
        const int pos = (int)((g.k0 + j) % g.Ncb);
        if (g.is_null(pos)) continue;
        const int x = g.src[pos] / g.Kpi, t = g.src[pos] % g.Kpi;
        e[k++] = d[x * D + (t - g.ND)];
    }
}

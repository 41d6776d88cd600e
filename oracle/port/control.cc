/* TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's control-channel receive path
 * (/root/reference/liblte/src/liblte_phy.cc), SURVEY 8f ranks 1-2:
 *   rate_unmatch_conv                 :11597-11752      viterbi_decode (K=7, rate 1/3)   :10161-10331
 *   bch_channel_decode                :12581-12655      liblte_phy_bch_channel_decode    :3968-4100
 *   pcfich_channel_demap              :7887-7962        cfi_channel_decode               :13648-13708
 *   phich_channel_demap (positions)   :8224-8287        liblte_phy_pdcch_channel_decode  :4519-5134
 *   dci_channel_decode                :12952-13035      dci_1a_unpack :13273-13372, dci_1c_unpack :13400-13605
 * Written from the behaviour of those functions, not copied.  Grids are [16][1200] floats per component and channel
 * estimates [N_ant][16][1200], as LIBLTE_PHY_SUBFRAME_STRUCT lays them out.
 *
 * Reference defect kept out of scope (SURVEY-style hazard "H14"): for N_ant >= 2 the PCFICH/PDCCH receive code hands
 * pre_decoder_and_matched_filter_dl a port stride of 576 for arrays whose rows are 288 long (liblte_phy.h:405 vs
 * liblte_phy.cc:4893, 7935), so ports 1-3 read other ports' rows or memory past the array; the transmit side has the same
 * mismatch (:4262 vs liblte_phy.h:407).  The control-channel restatement therefore covers N_ant == 1 for PCFICH/PDCCH
 * and N_ant in {1,2,4} for PBCH (whose arrays are consistent, liblte_phy.h:368).
 */
#include "oracle_port.h"
#include <math.h>
#include <string.h>
#include <vector>

namespace {
const uint8_t COL_PERM_CC[32] = {1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31,
                                 0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30};      /* 36.212 table 5.1.4-2 */
const uint32_t G_CC[3] = {0133, 0171, 0165};
/* 36.212 table 5.3.4-1 */
const uint8_t CFI_CW[4][32] = {
    {0,1,1,0,1,1,0,1,1,0,1,1,0,1,1,0,1,1,0,1,1,0,1,1,0,1,1,0,1,1,0,1},
    {1,0,1,1,0,1,1,0,1,1,0,1,1,0,1,1,0,1,1,0,1,1,0,1,1,0,1,1,0,1,1,0},
    {1,1,0,1,1,0,1,1,0,1,1,0,1,1,0,1,1,0,1,1,0,1,1,0,1,1,0,1,1,0,1,1},
    {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0}};
/* 36.213 table 7.1.7.2.1-1, columns N_PRB = 2 and 3 (all DCI 1A for SI/P/RA-RNTI can address), I_TBS 0..26 */
const int TBS_NPRB2[27] = {32, 56, 72, 104, 120, 144, 176, 224, 256, 296, 328, 376, 440, 488, 552, 600, 632, 696, 776, 840, 904, 1000, 1064, 1128, 1192, 1256, 1480};
const int TBS_NPRB3[27] = {56, 88, 144, 176, 208, 224, 256, 328, 392, 456, 504, 584, 680, 744, 840, 904, 968, 1064, 1160, 1288, 1384, 1480, 1608, 1736, 1800, 1864, 2216};
/* 36.213 table 7.1.7.2.3-1 (DCI 1C) */
const int TBS_1C[32] = {40, 56, 72, 120, 136, 144, 176, 208, 224, 256, 280, 296, 328, 336, 392, 488,
                        552, 600, 632, 696, 776, 840, 904, 1000, 1064, 1128, 1224, 1288, 1384, 1480, 1608, 1736};

/* order in which the circular buffer's non-dummy positions are visited: target index into d[3*N_c] */
std::vector<int> cc_buffer_order(int N_c)
{
    const int R = (N_c + 31) / 32, K_pi = 32 * R, n_dummy = K_pi - N_c;
    std::vector<int> tgt;
    for (int x = 0; x < 3; x++)
        for (int j = 0; j < 32; j++)
            for (int i = 0; i < R; i++) {
                const int src = i * 32 + COL_PERM_CC[j];     /* position in the row-major padded matrix */
                if (src >= n_dummy) tgt.push_back(3 * (src - n_dummy) + x);
            }
    return tgt;
}
}  // namespace

extern "C" int op_tbs_1a(int i_tbs, int n_prb_1a) { return i_tbs < 0 || i_tbs > 26 ? -1 : (n_prb_1a == 2 ? TBS_NPRB2[i_tbs] : TBS_NPRB3[i_tbs]); }
extern "C" int op_tbs_1c(int i) { return i < 0 || i > 31 ? -1 : TBS_1C[i]; }

/* rate_unmatch_conv: soft-combine the E received bits into the 3*N_c mother-code positions; positions never received keep RX_NULL */
extern "C" int op_rate_unmatch_conv(const float *e, int E, int N_c, float *d)
{
    const std::vector<int> tgt = cc_buffer_order(N_c);
    const int n = (int)tgt.size();                   /* == 3*N_c */
    for (int i = 0; i < n; i++) d[i] = OP_RX_NULL;
    for (int k = 0; k < E; k++) {
        float &w = d[tgt[k % n]];
        if (w == OP_RX_NULL) w = e[k];
        else if (e[k] != OP_RX_NULL) w += e[k];
    }
    return n;
}

/* viterbi_decode for the LTE tail-biting code (K = 7, g = 133/171/165 octal) exactly as the reference runs it: all states start at
 * metric 0, hard-decision Hamming branch metrics decide the survivor while the *weighted* metric (sum|d| * mismatches) is what gets
 * accumulated, the traceback re-decides from the stored path metrics, and the list of trace states starts with every successive
 * minimum of the final column -- so more than N_d/3 bits come out; callers use the first N_d/3. */
extern "C" int op_viterbi_decode(const float *d, int N_d, uint8_t *c)
{
    const int n = N_d / 3;
    uint8_t out[64][2][3];
    for (int j = 0; j < 64; j++)
        for (int k = 0; k < 2; k++) {
            const uint32_t reg = ((uint32_t)(j >= 32) << 6) | (uint32_t)((2 * j + k) % 64);
            for (int o = 0; o < 3; o++) out[j][k][o] = (uint8_t)(__builtin_popcount(reg & G_CC[o]) & 1);
        }
    std::vector<float> pm((size_t)64 * (n + 1), 0.0f);     /* pm[t*64 + state] */
    for (int t = 0; t < n; t++) {
        float w = 0;
        int hard[3];
        for (int o = 0; o < 3; o++) { hard[o] = d[3 * t + o] >= 0 ? 0 : 1; w += fabsf(d[3 * t + o]); }
        for (int j = 0; j < 64; j++) {
            float br[2], p[2];
            for (int k = 0; k < 2; k++) {
                p[k] = pm[(size_t)t * 64 + (2 * j + k) % 64];
                int mism = 0;
                for (int o = 0; o < 3; o++) mism += (out[j][k][o] + hard[o]) % 2;
                br[k] = (float)mism;
            }
            pm[(size_t)(t + 1) * 64 + j] = (br[0] + p[0] > br[1] + p[1]) ? p[1] + w * br[1] : p[0] + w * br[0];
        }
    }
    std::vector<int> tb;
    float best = 1000000;
    for (int j = 0; j < 64; j++)
        if (pm[(size_t)n * 64 + j] < best) { best = pm[(size_t)n * 64 + j]; tb.push_back(j); }
    if (tb.empty()) tb.push_back(1000000 & 0xff);           /* unreachable for finite inputs */
    for (int t = n - 1; t >= 0; t--) {
        const int s0 = ((uint8_t)tb.back() << 1) % 64, s1 = s0 + 1;
        tb.push_back(pm[(size_t)t * 64 + s0] > pm[(size_t)t * 64 + s1] ? s1 : s0);
    }
    int idx = 0;
    for (int i = (int)tb.size() - 2; i >= 0; i--) {
        if (tb[i] < tb[i + 1]) c[idx++] = 0;
        else if (tb[i] > tb[i + 1]) c[idx++] = 1;
        else c[idx++] = tb[i] == 0 ? 0 : 1;
    }
    return idx;
}

static uint32_t crc16_of(const uint8_t *bits, int n) { return op_crc(bits, n, 0x11021, 16); }

/* bch_channel_decode: 1920 soft bits -> 24 MIB bits if the CRC16 matches under the antenna mask */
extern "C" int op_bch_decode_bits(const float *in1920, int N_ant, uint8_t *out24)
{
    float d[120];
    uint8_t c[200];
    op_rate_unmatch_conv(in1920, 1920, 40, d);
    op_viterbi_decode(d, 120, c);
    uint32_t rx = 0;
    for (int i = 0; i < 16; i++) rx = (rx << 1) | c[24 + i];
    const uint32_t mask = N_ant == 1 ? 0x0000 : (N_ant == 2 ? 0xFFFF : 0x5555);
    if ((crc16_of(c, 24) ^ mask) != rx) return 3;             /* LIBLTE_ERROR_INVALID_CRC */
    memcpy(out24, c, 24);
    return 0;
}

/* liblte_phy_bch_channel_decode: tries 1, 2, 4 ports x 4 positions inside the 40 ms period; first hit wins */
extern "C" int op_bch_channel_decode(const float *rx_re, const float *rx_im, const float *ce_re, const float *ce_im, int N_rb_dl, int N_id_cell,
                                     int *N_ant, uint8_t *out24, int *offset)
{
    float yr[240], yi[240], hr[4][240], hi[4][240];
    int idx = 0;
    for (int i = 0; i < 72; i++) {
        const int k = 6 * N_rb_dl - 36 + i;
        auto take = [&](int dst, int sym) {
            yr[dst] = rx_re[sym * 1200 + k]; yi[dst] = rx_im[sym * 1200 + k];
            for (int p = 0; p < 4; p++) { hr[p][dst] = ce_re[((size_t)p * 16 + sym) * 1200 + k]; hi[p][dst] = ce_im[((size_t)p * 16 + sym) * 1200 + k]; }
        };
        if (N_id_cell % 3 != i % 3) { take(idx, 7); take(idx + 48, 8); idx++; }    /* symbols 0-1 of slot 1 carry CRS for ports 0-3 */
        take(i + 96, 9);
        take(i + 168, 10);
    }
    std::vector<uint8_t> c(1920);
    op_prs_c((uint32_t)N_id_cell, 1920, c.data());
    *N_ant = 0;
    for (int p = 1; p <= 4; p++) {
        if (p == 3) continue;
        float dr[244], di[244];
        int8_t soft[488];
        const int M = op_pre_decoder_dl(yr, yi, &hr[0][0], &hi[0][0], 240, 240, p, dr, di);
        op_demap(dr, di, M, 1, soft);
        for (int off = 0; off < 4; off++) {
            std::vector<float> in(1920, OP_RX_NULL);
            for (int j = 0; j < 480; j++) in[off * 480 + j] = (float)soft[j] * (1 - 2 * (float)c[off * 480 + j]);
            if (op_bch_decode_bits(in.data(), p, out24) == 0) { *offset = off; *N_ant = p; return 0; }
        }
    }
    return 2;                                                  /* LIBLTE_ERROR_DECODE_FAIL */
}

/* ------------------------------------------------------------------------------------------------ PCFICH / PDCCH */
extern "C" int op_dci_sizes(int N_rb_dl, int *size_1a, int *size_1c)
{
    static const int rb[6] = {6, 15, 25, 50, 75, 100}, a[6] = {21, 22, 25, 27, 27, 28}, cc[6] = {9, 11, 13, 13, 14, 15};
    for (int i = 0; i < 6; i++)
        if (rb[i] == N_rb_dl) { *size_1a = a[i]; *size_1c = cc[i]; return 0; }
    *size_1a = 28; *size_1c = 15;                              /* the reference's fall-through */
    return 0;
}

/* PCFICH REG positions (36.211 6.7.4) and the PHICH REG positions the reference derives from them (:8224-8287) */
extern "C" void op_control_reg_positions(int N_rb_dl, int N_id_cell, float phich_res, int *pcfich_k /*[4]*/, int *phich_k /*[75]*/, int *N_reg_phich)
{
    float pc_n[4];
    const int k_hat = 6 * (N_id_cell % (2 * N_rb_dl));
    for (int i = 0; i < 4; i++) {
        pcfich_k[i] = (k_hat + (i * N_rb_dl / 2) * 6) % (N_rb_dl * 12);
        pc_n[i] = (pcfich_k[i] / 6) - 0.5f;
    }
    const int n_group = (int)ceilf(phich_res * ((float)N_rb_dl / 8.0f));
    *N_reg_phich = 3 * n_group;
    const uint32_t n_l = (uint32_t)(2 * N_rb_dl - 4);
    for (int m = 0; m < n_group; m++) {
        uint32_t n_hat[3];
        for (uint32_t i = 0; i < 3; i++) n_hat[i] = ((uint32_t)N_id_cell + (uint32_t)m + i * n_l / 3) % n_l;
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 3; j++)
                if ((float)n_hat[j] > pc_n[i]) n_hat[j]++;
        for (int i = 0; i < 3; i++) phich_k[3 * m + i] = (int)(n_hat[i] * 6);
    }
}

/* The REG walk of 36.211 6.8.5 as the reference does it (:4591-4701), then un-shift, un-interleave and CCE assembly (:4702-4838),
 * folded into one table: re_of[c*36 + j] = symbol*1200 + subcarrier of RE j of CCE c.  Returns N_cce. */
extern "C" int op_pdcch_cce_map(int N_rb_dl, int N_id_cell, int N_ant, int N_symbs, float phich_res, int *re_of /*[N_cce*36]*/, int *N_reg_out)
{
    int pk[4], hk[75], n_ph;
    op_control_reg_positions(N_rb_dl, N_id_cell, phich_res, pk, hk, &n_ph);
    int N_reg = N_symbs * (N_rb_dl * 3) - N_rb_dl - 4 - n_ph;
    if (N_ant == 4) N_reg -= N_rb_dl;
    if (N_reg_out) *N_reg_out = N_reg;
    if (N_reg <= 0) return 0;
    const int N_cce = N_reg / 9;
    std::vector<int> reg_re((size_t)N_reg * 4, -1);            /* REG m, element j -> RE */
    int m = 0;
    for (int k = 0; k < 12 * N_rb_dl; k++)
        for (int l = 0; l < N_symbs; l++) {
            const bool crs_sym = (l == 0) || (l == 1 && N_ant == 4);
            if (crs_sym) {
                bool ok = k % 6 == 0;
                if (l == 0) {
                    for (int i = 0; i < 4; i++) if (k == pk[i]) ok = false;
                    for (int i = 0; i < n_ph; i++) if (k == hk[i]) ok = false;
                }
                if (ok && m < N_reg) {
                    int j = 0;
                    for (int i = 0; i < 6; i++) if (N_id_cell % 3 != i % 3) reg_re[(size_t)m * 4 + j++] = l * 1200 + k + i;
                    m++;
                }
            } else if (k % 4 == 0 && m < N_reg) {
                for (int i = 0; i < 4; i++) reg_re[(size_t)m * 4 + i] = l * 1200 + k + i;
                m++;
            }
        }
    /* cyclic shift: shifted[(i + N_id_cell) % N_reg] = reg[i] */
    std::vector<int> shifted((size_t)N_reg * 4);
    for (int i = 0; i < N_reg; i++) memcpy(&shifted[(size_t)((i + N_id_cell) % N_reg) * 4], &reg_re[(size_t)i * 4], 4 * sizeof(int));
    /* de-interleaver: the k-th non-dummy position of the column-permuted read-out receives k; reading the matrix back row-major
     * (dummies first) gives, for interleaver input position i, the index of the shifted REG that carries it */
    const int R = (N_reg + 31) / 32, K_pi = 32 * R, n_dummy = K_pi - N_reg;
    std::vector<int> mat(K_pi, -1);
    int kk = 0;
    for (int j = 0; j < 32; j++)
        for (int i = 0; i < R; i++) {
            const int src = i * 32 + COL_PERM_CC[j];
            if (src >= n_dummy) mat[src] = kk++;
        }
    for (int c = 0; c < N_cce; c++)
        for (int r = 0; r < 9; r++) {
            const int from = mat[n_dummy + c * 9 + r];
            for (int j = 0; j < 4; j++) re_of[c * 36 + r * 4 + j] = shifted[(size_t)from * 4 + j];
        }
    return N_cce;
}

/* dci_channel_decode on descrambled soft bits: returns 1 and the matching RNTI if the CRC16 residue is SI-, P- or an RA-RNTI
 * (the order the reference tries them in cannot matter: the residue equals at most one value) */
extern "C" int op_dci_decode_bits(const float *soft, int E, int n_dci, uint8_t *dci_bits, int *rnti)
{
    std::vector<float> d(3 * (n_dci + 16));
    std::vector<uint8_t> c(3 * (n_dci + 16) + 80);
    op_rate_unmatch_conv(soft, E, n_dci + 16, d.data());
    op_viterbi_decode(d.data(), 3 * (n_dci + 16), c.data());
    uint32_t rx = 0;
    for (int i = 0; i < 16; i++) rx = (rx << 1) | c[n_dci + i];
    const uint32_t x = rx ^ crc16_of(c.data(), n_dci);
    if (!(x == 0xFFFF || x == 0xFFFE || (x >= 1 && x <= 0x3C))) return 0;
    memcpy(dci_bits, c.data(), n_dci);
    *rnti = (int)x;
    return 1;
}

namespace {
uint32_t take_bits(const uint8_t *&p, int n) { uint32_t v = 0; while (n--) v = (v << 1) | *p++; return v; }
int riv_bits(float count) { return (int)ceilf(logf(count) / logf(2)); }
void fill_common(int *o, int rnti, int tbs, int mcs, int n_prb, int rv, int N_ant)
{
    o[0] = rnti; o[1] = tbs; o[2] = mcs; o[3] = n_prb; o[4] = rv; o[5] = 1 /* QPSK */; o[6] = N_ant == 1 ? 1 : 2; o[7] = 0;
}
}  // namespace

/* alloc layout: {rnti, tbs, mcs, N_prb, rv_idx, mod_type, tx_mode, chan_type, prb slot0 [110], prb slot1 [110]} */
extern "C" int op_dci_1a_unpack(const uint8_t *bits, int rnti, int N_rb_dl, int N_ant, int *o)
{
    const uint8_t *p = bits;
    if (take_bits(p, 1) == 0) return 4;                       /* flagged as format 0: LIBLTE_ERROR_INVALID_CONTENTS */
    const uint32_t distributed = take_bits(p, 1);
    const uint32_t riv = take_bits(p, riv_bits((float)(N_rb_dl * (N_rb_dl + 1) / 2)));
    const int n_prb = (int)(riv / N_rb_dl) + 1, rb0 = (int)(riv % N_rb_dl);
    const int mcs = (int)take_bits(p, 5);
    take_bits(p, 3); take_bits(p, 1);                          /* HARQ process, NDI */
    const int rv = (int)take_bits(p, 2);
    const int tpc = (int)take_bits(p, 2);
    fill_common(o, rnti, 0, mcs, n_prb, rv, N_ant);
    if (!distributed)
        for (int i = 0; i < n_prb && i < 110; i++) o[8 + i] = o[118 + i] = rb0 + i;
    if (mcs >= 27) return 4;
    o[1] = op_tbs_1a(mcs, tpc % 2 == 0 ? 2 : 3);
    return 0;
}

extern "C" int op_dci_1c_unpack(const uint8_t *bits, int rnti, int N_rb_dl, int N_ant, int *o)
{
    const uint8_t *p = bits;
    const int gap_ind = N_rb_dl < 50 ? 0 : (int)take_bits(p, 1);
    int N_gap;
    if (N_rb_dl <= 10) N_gap = (int)ceilf(N_rb_dl / 2.0);
    else if (N_rb_dl == 11) N_gap = 4;
    else if (N_rb_dl <= 19) N_gap = 8;
    else if (N_rb_dl <= 26) N_gap = 12;
    else if (N_rb_dl <= 44) N_gap = 18;
    else if (N_rb_dl <= 49) N_gap = 27;
    else if (N_rb_dl <= 63) N_gap = gap_ind == 0 ? 27 : 9;
    else if (N_rb_dl <= 70) N_gap = gap_ind == 0 ? 32 : 16;
    else N_gap = gap_ind == 0 ? 48 : 16;
    const int vrb_gap1 = 2 * (N_gap < N_rb_dl - N_gap ? N_gap : N_rb_dl - N_gap);
    const int vrb_gap2 = (N_rb_dl / (2 * N_gap)) * 2 * N_gap;
    const int N_vrb = gap_ind == 0 ? vrb_gap1 : vrb_gap2;
    const int step = N_rb_dl <= 49 ? 2 : 4;
    const int Nt_vrb = gap_ind == 0 ? N_vrb : 2 * N_gap;
    const int g1 = vrb_gap1 / step;
    const uint32_t riv = take_bits(p, (int)ceilf(logf(g1 * (g1 + 1) / 2.0) / logf(2)));
    /* invert the RIV by enumeration, last match wins, exactly like the reference's double loop (:13496-13522) */
    int n_prb = 0, rb0 = 0;
    const int nd = N_vrb / step;
    for (int i = step; i <= nd * step; i += step)
        for (int j = 0; j <= (nd - 1) * step; j += step) {
            const int id = i / step, jd = j / step;
            const uint32_t v = (uint32_t)(id - 1) <= (uint32_t)(nd / 2) ? (uint32_t)(nd * (id - 1) + jd) : (uint32_t)(nd * (nd - id + 1) + (nd - 1 - jd));
            if (riv == v) { n_prb = i; rb0 = j; }
        }
    const int mcs = (int)take_bits(p, 5);
    const int P = N_rb_dl <= 10 ? 1 : (N_rb_dl <= 26 ? 2 : (N_rb_dl <= 63 ? 3 : 4));
    const uint32_t N_row = (uint32_t)(ceilf(Nt_vrb / (4.0 * P)) * P), N_null = 4 * N_row - Nt_vrb, Nt = (uint32_t)Nt_vrb;
    fill_common(o, rnti, 0, mcs, n_prb, 0, N_ant);
    uint32_t n_vrb = (uint32_t)rb0;
    for (int i = 0; i < n_prb && i < 110; i++, n_vrb++) {      /* distributed VRB -> PRB, 36.211 6.2.3.2, in the reference's unsigned arithmetic */
        const uint32_t nt = n_vrb % Nt, blk = Nt * (n_vrb / Nt);
        const uint32_t d1 = 2 * N_row * (nt % 2) + nt / 2 + blk, d2 = N_row * (nt % 4) + nt / 4 + blk;
        uint32_t even;
        if (N_null != 0 && nt >= Nt - N_null && nt % 2 == 1) even = d1 - N_row;
        else if (N_null != 0 && nt >= Nt - N_null && nt % 2 == 0) even = d1 - N_row + N_null / 2;
        else if (N_null != 0 && nt < Nt - N_null && nt % 4 >= 2) even = d2 - N_null / 2;
        else even = d2;
        const uint32_t odd = (even + Nt / 2) % Nt + blk;
        o[8 + i] = (int)(even < Nt / 2 ? even : even + N_gap - Nt / 2);
        o[118 + i] = (int)(odd < Nt / 2 ? odd : odd + N_gap - Nt / 2);
    }
    if (mcs >= 32) return 4;
    o[1] = op_tbs_1c(mcs);
    return 0;
}

/* liblte_phy_pdcch_channel_decode, N_ant == 1.  Outputs: cfi (PCFICH value), N_symbs, allocations (up to 6), and the raw result of
 * the 12 blind decodes in the reference's order {L=4 cand 0..3, L=8 cand 0..1} x {1A, 1C}: found flag, RNTI, DCI bits. */
extern "C" int op_pdcch_channel_decode(const float *rx_re, const float *rx_im, const float *ce_re, const float *ce_im, int N_rb_dl, int N_id_cell,
                                       int N_ant, int subfr, float phich_res, int *cfi_out, int *N_symbs_out, int *N_alloc, int *allocs /*[6][228]*/,
                                       int *raw_found /*[12]*/, int *raw_rnti /*[12]*/, uint8_t *raw_bits /*[12][32]*/)
{
    *N_alloc = 0; *cfi_out = 0; *N_symbs_out = 0;
    if (raw_found) memset(raw_found, 0, 12 * sizeof(int));
    if (N_ant != 1) return -100;
    /* PCFICH */
    int pk[4], hk[75], n_ph;
    op_control_reg_positions(N_rb_dl, N_id_cell, phich_res, pk, hk, &n_ph);
    float yr[16], yi[16], hr[16], hi[16], xr[16], xi[16];
    for (int i = 0; i < 4; i++) {
        int idx = 0;
        for (int j = 0; j < 6; j++)
            if (N_id_cell % 3 != j % 3) {
                const int k = pk[i] + j;
                yr[4 * i + idx] = rx_re[k]; yi[4 * i + idx] = rx_im[k]; hr[4 * i + idx] = ce_re[k]; hi[4 * i + idx] = ce_im[k];
                idx++;
            }
    }
    uint8_t c32[32];
    int8_t s32[32];
    op_prs_c((((uint32_t)(subfr + 1) * (2 * (uint32_t)N_id_cell + 1)) << 9) + (uint32_t)N_id_cell, 32, c32);
    op_zf_1ant(yr, yi, hr, hi, 16, xr, xi);
    op_demap(xr, xi, 16, 1, s32);
    int ber[4] = {0, 0, 0, 0};
    for (int i = 0; i < 32; i++) {
        const float v = (float)s32[i] * (1 - 2 * (float)c32[i]);
        const int hard = v >= 0 ? 0 : 1;
        for (int w = 0; w < 4; w++) ber[w] += CFI_CW[w][i] != hard;
    }
    int min_ber = 32, cfi = 0;
    for (int w = 0; w < 4; w++) if (ber[w] < min_ber) { min_ber = ber[w]; cfi = w + 1; }
    if (min_ber >= 4) return 3;                                /* LIBLTE_ERROR_INVALID_CRC */
    *cfi_out = cfi;
    const int N_symbs = cfi + (N_rb_dl <= 10 ? 1 : 0);
    *N_symbs_out = N_symbs;
    /* PDCCH */
    int N_reg = 0;
    std::vector<int> re_of((size_t)(N_symbs * N_rb_dl * 3 / 9 + 1) * 36);
    const int N_cce = N_symbs <= 4 ? op_pdcch_cce_map(N_rb_dl, N_id_cell, N_ant, N_symbs, phich_res, re_of.data(), &N_reg) : 0;
    std::vector<uint8_t> c(1152);
    op_prs_c(((uint32_t)subfr << 9) + (uint32_t)N_id_cell, 1152, c.data());
    int sz[2];
    op_dci_sizes(N_rb_dl, &sz[0], &sz[1]);
    int err = 1;                                               /* LIBLTE_ERROR_INVALID_INPUTS: what the reference returns when nothing decodes */
    for (int cand = 0; cand < 6; cand++) {
        const int L = cand < 4 ? 4 : 8, first = cand < 4 ? 4 * cand : 8 * (cand - 4), n_re = 36 * L;
        /* CCEs past the last one the control region holds: the reference reads its scratch arrays there -- zeros in a freshly
         * initialised struct, i.e. 0/0 = NaN through the equaliser, which its demapper turns into soft bit 0 (x86 float->int8 of
         * NaN).  Restated as exactly that: those resource elements contribute soft bits of 0.  (With N_cce >= 16, i.e. every
         * bandwidth >= 5 MHz at CFI >= 2, the case does not arise.) */
        std::vector<float> Yr(n_re, 0.f), Yi(n_re, 0.f), Hr(n_re, 1.f), Hi(n_re, 0.f), Xr(n_re), Xi(n_re), soft(2 * n_re);
        std::vector<int8_t> sb(2 * n_re);
        std::vector<char> missing(n_re, 0);
        for (int i = 0; i < n_re; i++) {
            if (first + i / 36 >= N_cce) { missing[i] = 1; continue; }
            const int re = re_of[(size_t)first * 36 + i];
            Yr[i] = rx_re[re]; Yi[i] = rx_im[re]; Hr[i] = ce_re[re]; Hi[i] = ce_im[re];
        }
        op_zf_1ant(Yr.data(), Yi.data(), Hr.data(), Hi.data(), n_re, Xr.data(), Xi.data());
        op_demap(Xr.data(), Xi.data(), n_re, 1, sb.data());
        for (int i = 0; i < n_re; i++) if (missing[i]) sb[2 * i] = sb[2 * i + 1] = 0;
        const int c_off = cand < 4 ? cand * 288 : (cand - 4) * 576;
        for (int j = 0; j < 2 * n_re; j++) soft[j] = (float)sb[j] * (1 - 2 * (float)c[c_off + j]);
        for (int fmt = 0; fmt < 2; fmt++) {
            uint8_t bits[32];
            int rnti = 0;
            const int slot = cand * 2 + fmt;
            const int found = op_dci_decode_bits(soft.data(), 2 * n_re, sz[fmt], bits, &rnti);
            if (raw_found) { raw_found[slot] = found; raw_rnti[slot] = found ? rnti : 0; if (found) memcpy(raw_bits + 32 * slot, bits, sz[fmt]); }
            if (!found || *N_alloc >= 6) continue;
            int *o = allocs + 228 * (*N_alloc);
            memset(o, 0, 228 * sizeof(int));
            err = fmt == 0 ? op_dci_1a_unpack(bits, rnti, N_rb_dl, N_ant, o) : op_dci_1c_unpack(bits, rnti, N_rb_dl, N_ant, o);
            if (err == 0) (*N_alloc)++;
        }
    }
    return err;
}

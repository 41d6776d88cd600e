// liblte_phy-compatible C++ surface (include/liblte_phy_b200.h) over the C ABI.  Host-side glue only: argument
// checks and error codes follow the reference functions cited beside each entry point.
#include "../../include/liblte_phy_b200.h"
#include "../../include/lte_b200.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static lteb200_ctx *C_(LIBLTE_PHY_STRUCT *p) { return (lteb200_ctx *)p->b200_ctx; }

// liblte_phy.cc:2210-2336
LIBLTE_ERROR_ENUM liblte_phy_init(LIBLTE_PHY_STRUCT **phy_struct, LIBLTE_PHY_FS_ENUM fs, uint16 N_id_cell, uint8 N_ant, uint32 N_rb_dl,
                                  uint32 N_sc_rb_dl, float phich_res)
{
    (void)N_id_cell; (void)phich_res;   // CRS tables are built per cell on first use; PHICH/PDCCH are not on this path
    if (!phy_struct || (int)fs < 0 || fs >= LIBLTE_PHY_FS_N_ITEMS) return LIBLTE_ERROR_INVALID_INPUTS;
    LIBLTE_PHY_STRUCT *p = (LIBLTE_PHY_STRUCT *)calloc(1, sizeof(LIBLTE_PHY_STRUCT));
    if (!p) return LIBLTE_ERROR_INVALID_INPUTS;
    const int k = (int)fs;
    p->fs = 1920000u << k;
    p->N_samps_per_symb = 128u << k; p->N_samps_cp_l_0 = 10u << k; p->N_samps_cp_l_else = 9u << k;
    p->N_samps_per_slot = 960u << k; p->N_samps_per_subfr = 1920u << k; p->N_samps_per_frame = 19200u << k;
    p->N_sc_rb_dl = N_sc_rb_dl; p->N_sc_rb_ul = 12;
    p->b200_fs_idx = k; p->b200_turbo_mode = LTEB200_TURBO_REF_EXACT; p->b200_turbo_iters = 8;
    liblte_phy_update_n_rb_dl(p, N_rb_dl);
    p->N_ant = N_ant; p->ul_init = false;
    lteb200_ctx *c = NULL;
    if (lteb200_create(&c, -1) != LTEB200_OK) { free(p); return LIBLTE_ERROR_INVALID_INPUTS; }   // no GPU: fail loudly, never fall back
    p->b200_ctx = c;
    *phy_struct = p;
    return LIBLTE_SUCCESS;
}

LIBLTE_ERROR_ENUM liblte_phy_cleanup(LIBLTE_PHY_STRUCT *p)
{
    if (!p) return LIBLTE_ERROR_INVALID_INPUTS;
    lteb200_destroy(C_(p));
    free(p);
    return LIBLTE_SUCCESS;
}

// liblte_phy.cc:2592-2652: legal (sample rate, bandwidth) pairs
LIBLTE_ERROR_ENUM liblte_phy_update_n_rb_dl(LIBLTE_PHY_STRUCT *p, uint32 N_rb_dl)
{
    if (!p) return LIBLTE_ERROR_INVALID_INPUTS;
    bool ok = false;
    switch (p->N_samps_per_symb) {
    case 2048: ok = true; break;
    case 1024: ok = N_rb_dl == 50 || N_rb_dl == 25 || N_rb_dl == 15 || N_rb_dl == 6; break;
    case 512:  ok = N_rb_dl == 25 || N_rb_dl == 15 || N_rb_dl == 6; break;
    case 256:  ok = N_rb_dl == 15 || N_rb_dl == 6; break;
    case 128:  ok = N_rb_dl == 6; break;
    }
    if (!ok) return LIBLTE_ERROR_INVALID_INPUTS;
    p->FFT_size = p->N_samps_per_symb;
    p->N_rb_dl = N_rb_dl; p->N_rb_ul = N_rb_dl;
    p->FFT_pad_size = (p->FFT_size - N_rb_dl * p->N_sc_rb_dl) / 2;
    return LIBLTE_SUCCESS;
}

// liblte_phy.cc:5697-5852.  The reference reads N_slots*slot + slot + CP + N samples; the caller's buffer must hold them (as there).
LIBLTE_ERROR_ENUM liblte_phy_dl_find_coarse_timing_and_freq_offset(LIBLTE_PHY_STRUCT *p, float *i_samps, float *q_samps, uint32 N_slots,
                                                                   LIBLTE_PHY_COARSE_TIMING_STRUCT *t)
{
    if (!p || !i_samps || !q_samps || !t) return LIBLTE_ERROR_INVALID_INPUTS;
    const long long need = (long long)N_slots * p->N_samps_per_slot + p->N_samps_per_slot + p->N_samps_cp_l_else + p->N_samps_per_symb;
    int rc = lteb200_find_coarse_timing_and_freq_offset(C_(p), i_samps, q_samps, need, p->b200_fs_idx, N_slots, t->freq_offset, &t->symb_starts[0][0],
                                                        &t->n_corr_peaks);
    return rc == LTEB200_OK ? LIBLTE_SUCCESS : LIBLTE_ERROR_INVALID_INPUTS;
}

// liblte_phy.cc:5306-5510
LIBLTE_ERROR_ENUM liblte_phy_find_pss_and_fine_timing(LIBLTE_PHY_STRUCT *p, float *i_samps, float *q_samps, uint32 *symb_starts, uint32 *N_id_2,
                                                      uint32 *pss_symb, float *pss_thresh, float *freq_offset)
{
    if (!p || !i_samps || !q_samps || !symb_starts || !N_id_2 || !pss_symb || !pss_thresh) return LIBLTE_ERROR_INVALID_INPUTS;
    uint32 mx = 0;
    for (int j = 0; j < 7; j++) mx = symb_starts[j] > mx ? symb_starts[j] : mx;
    const long long need = (long long)mx + 12LL * p->N_samps_per_slot + p->N_samps_per_symb + p->N_samps_cp_l_0 + 64;
    float fo = 0;
    int rc = lteb200_find_pss_and_fine_timing(C_(p), i_samps, q_samps, need, p->b200_fs_idx, (int)p->N_rb_dl, symb_starts, N_id_2, pss_symb, pss_thresh, &fo);
    if (freq_offset) *freq_offset = fo;
    return rc == LTEB200_OK ? LIBLTE_SUCCESS : LIBLTE_ERROR_INVALID_INPUTS;
}

// liblte_phy.cc:5578-5686: the entry default (INVALID_INPUTS) is what the reference returns when no SSS matches
LIBLTE_ERROR_ENUM liblte_phy_find_sss(LIBLTE_PHY_STRUCT *p, float *i_samps, float *q_samps, uint32 N_id_2, uint32 *symb_starts, float pss_thresh,
                                      uint32 *N_id_1, uint32 *frame_start_idx)
{
    if (!p || !i_samps || !q_samps || !symb_starts || !N_id_1 || !frame_start_idx) return LIBLTE_ERROR_INVALID_INPUTS;
    const long long need = (long long)symb_starts[5] + p->N_samps_cp_l_0 + p->N_samps_per_symb + 16;
    int rc = lteb200_find_sss(C_(p), i_samps, q_samps, need, p->b200_fs_idx, (int)p->N_rb_dl, N_id_2, symb_starts, pss_thresh, N_id_1, frame_start_idx);
    return rc == LTEB200_OK ? LIBLTE_SUCCESS : LIBLTE_ERROR_INVALID_INPUTS;
}

// liblte_phy.cc:5905-6200
LIBLTE_ERROR_ENUM liblte_phy_get_dl_subframe_and_ce(LIBLTE_PHY_STRUCT *p, float *i_samps, float *q_samps, uint32 frame_start_idx, uint8 subfr_num,
                                                    uint32 N_id_cell, uint8 N_ant, LIBLTE_PHY_SUBFRAME_STRUCT *sf)
{
    if (!p || !i_samps || !q_samps || !(N_ant == 1 || N_ant == 2 || N_ant == 4) || !sf) return LIBLTE_ERROR_INVALID_INPUTS;
    const uint32 start = frame_start_idx + (uint32)subfr_num * p->N_samps_per_subfr;
    // the last sample the reference reads is the end of symbol 1 of the NEXT subframe (liblte_phy.cc:5946-5957, 8621): nothing beyond it is touched here
    const long long need = (long long)start + p->N_samps_per_subfr + 2LL * p->N_samps_per_symb + p->N_samps_cp_l_0 + p->N_samps_cp_l_else;
    sf->num = subfr_num;
    int rc = lteb200_get_dl_subframe_and_ce(C_(p), i_samps, q_samps, need, p->b200_fs_idx, (int)p->N_rb_dl, frame_start_idx, subfr_num, N_id_cell, N_ant,
                                            &sf->rx_symb_re[0][0], &sf->rx_symb_im[0][0], &sf->rx_ce_re[0][0][0], &sf->rx_ce_im[0][0][0]);
    return rc == LTEB200_OK ? LIBLTE_SUCCESS : LIBLTE_ERROR_INVALID_INPUTS;
}

// liblte_phy.cc:3690-3853: INVALID_INPUTS for bad arguments, DECODE_FAIL when the CRC does not check (outputs untouched)
LIBLTE_ERROR_ENUM liblte_phy_pdsch_channel_decode(LIBLTE_PHY_STRUCT *p, LIBLTE_PHY_SUBFRAME_STRUCT *sf, LIBLTE_PHY_ALLOCATION_STRUCT *alloc,
                                                  uint32 N_pdcch_symbs, uint32 N_id_cell, uint8 N_ant, uint8 *out_bits, uint32 *N_out_bits)
{
    if (!p || !sf || !alloc || N_id_cell > 503 || !out_bits || !N_out_bits) return LIBLTE_ERROR_INVALID_INPUTS;
    lteb200_alloc a;
    memset(&a, 0, sizeof(a));
    a.mod_type = (int)alloc->mod_type; a.N_prb = (int)alloc->N_prb; a.tbs = (int)alloc->tbs; a.rv_idx = (int)alloc->rv_idx;
    a.tx_mode = (int)alloc->tx_mode; a.rnti = alloc->rnti;
    for (uint32 s = 0; s < 2; s++) for (uint32 j = 0; j < alloc->N_prb && j < 110; j++) a.prb[s][j] = (int)alloc->prb[s][j];
    int rc = lteb200_pdsch_channel_decode(C_(p), &sf->rx_symb_re[0][0], &sf->rx_symb_im[0][0], &sf->rx_ce_re[0][0][0], &sf->rx_ce_im[0][0][0],
                                          (int)p->N_rb_dl, sf->num, &a, N_pdcch_symbs, N_id_cell, N_ant, p->b200_turbo_mode, p->b200_turbo_iters,
                                          out_bits, N_out_bits);
    if (rc == LTEB200_OK) return LIBLTE_SUCCESS;
    return rc == LTEB200_ERR_ARGS ? LIBLTE_ERROR_INVALID_INPUTS : LIBLTE_ERROR_DECODE_FAIL;
}

// liblte_phy.cc:3968-4100: INVALID_INPUTS for bad arguments, DECODE_FAIL when none of the 12 hypotheses passes its CRC
LIBLTE_ERROR_ENUM liblte_phy_bch_channel_decode(LIBLTE_PHY_STRUCT *p, LIBLTE_PHY_SUBFRAME_STRUCT *sf, uint32 N_id_cell, uint8 *N_ant, uint8 *out_bits,
                                                uint32 *N_out_bits, uint8 *offset)
{
    if (!p || !sf || N_id_cell > 503 || !N_ant || !out_bits || !N_out_bits || !offset) return LIBLTE_ERROR_INVALID_INPUTS;
    const int rc = lteb200_bch_channel_decode(C_(p), &sf->rx_symb_re[0][0], &sf->rx_symb_im[0][0], &sf->rx_ce_re[0][0][0], &sf->rx_ce_im[0][0][0],
                                              (int)p->N_rb_dl, N_id_cell, N_ant, out_bits, N_out_bits, offset);
    if (rc == LTEB200_OK) return LIBLTE_SUCCESS;
    return rc == LTEB200_ERR_ARGS ? LIBLTE_ERROR_INVALID_INPUTS : LIBLTE_ERROR_DECODE_FAIL;
}

// liblte_phy.cc:4519-5134.  Return value as the reference's: INVALID_CRC when the PCFICH does not decode, INVALID_INPUTS when no DCI
// passes its CRC (sic), otherwise the result of unpacking the last DCI that did.
LIBLTE_ERROR_ENUM liblte_phy_pdcch_channel_decode(LIBLTE_PHY_STRUCT *p, LIBLTE_PHY_SUBFRAME_STRUCT *sf, uint32 N_id_cell, uint8 N_ant, float phich_res,
                                                  LIBLTE_RRC_PHICH_DURATION_ENUM phich_dur, LIBLTE_PHY_PCFICH_STRUCT *pcfich,
                                                  LIBLTE_PHY_PHICH_STRUCT *phich, LIBLTE_PHY_PDCCH_STRUCT *pdcch)
{
    if (!p || !sf || !pcfich || !phich || !pdcch) return LIBLTE_ERROR_INVALID_INPUTS;
    lteb200_dci dci[6];            // per call: distinct phy structs may be used from distinct threads
    uint32 cfi = 0, n_symbs = 0, n_alloc = 0;
    const int rc = lteb200_pdcch_channel_decode(C_(p), &sf->rx_symb_re[0][0], &sf->rx_symb_im[0][0], &sf->rx_ce_re[0][0][0], &sf->rx_ce_im[0][0][0],
                                                (int)p->N_rb_dl, sf->num, N_id_cell, N_ant, phich_res, (int)phich_dur, &cfi, &n_symbs, dci, &n_alloc);
    if (rc == LTEB200_ERR_CUDA || rc == LTEB200_ERR_UNSUPPORTED) return LIBLTE_ERROR_DECODE_FAIL;
    pcfich->N_reg = 4;
    phich->N_reg = 3 * (uint32)ceilf(phich_res * ((float)p->N_rb_dl / 8.0f));
    if (rc == LTEB200_ERR_DECODE) return LIBLTE_ERROR_INVALID_CRC;
    pcfich->cfi = cfi;
    pdcch->N_symbs = n_symbs;
    pdcch->N_alloc = n_alloc;
    for (uint32 a = 0; a < n_alloc; a++) {
        LIBLTE_PHY_ALLOCATION_STRUCT *o = &pdcch->alloc[a];
        const lteb200_alloc *in = &dci[a].alloc;
        o->pre_coder_type = LIBLTE_PHY_PRE_CODER_TYPE_TX_DIVERSITY;
        o->mod_type = (LIBLTE_PHY_MODULATION_TYPE_ENUM)in->mod_type;
        o->tbs = (uint32)in->tbs; o->rv_idx = (uint32)in->rv_idx; o->N_prb = (uint32)in->N_prb; o->tx_mode = (uint32)in->tx_mode;
        o->N_codewords = 1; o->rnti = (uint16)in->rnti; o->mcs = (uint8)dci[a].mcs;
        for (int s_ = 0; s_ < 2; s_++) for (int j = 0; j < in->N_prb && j < 110; j++) o->prb[s_][j] = (uint32)in->prb[s_][j];
    }
    if (rc == LTEB200_OK) return LIBLTE_SUCCESS;
    return rc == LTEB200_ERR_CONTENTS ? LIBLTE_ERROR_INVALID_CONTENTS : LIBLTE_ERROR_INVALID_INPUTS;
}

// liblte_phy.cc:11246-11490 (void: cannot fail; on a device error d_bits is left untouched and *N_d_bits = 0)
void liblte_phy_rate_unmatch_turbo(LIBLTE_PHY_STRUCT *p, float *e_bits, uint32 N_e_bits, uint8 *dummy_bits, uint32 N_dummy_bits, uint32 N_codeblocks,
                                   uint32 tx_mode, uint32 N_soft, uint32 M_dl_harq, LIBLTE_PHY_CHAN_TYPE_ENUM chan_type, uint32 rv_idx, float *d_bits,
                                   uint32 *N_d_bits)
{
    uint32 n = 0;
    if (p && lteb200_rate_unmatch_turbo(C_(p), e_bits, N_e_bits, dummy_bits, N_dummy_bits, N_codeblocks, tx_mode, N_soft, M_dl_harq, (int)chan_type, rv_idx,
                                        d_bits, &n) != LTEB200_OK) n = 0;
    if (N_d_bits) *N_d_bits = n;
}

// liblte_phy.cc:10620-10845: d_bits modified in place, N_fill_bits ignored, *N_c_bits never written
void turbo_decode(LIBLTE_PHY_STRUCT *p, float *d_bits, uint32 N_d_bits, uint32 N_fill_bits, uint8 *c_bits, uint32 *N_c_bits)
{
    (void)N_c_bits;
    if (p) lteb200_turbo_decode(C_(p), d_bits, N_d_bits, N_fill_bits, c_bits, p->b200_turbo_mode, p->b200_turbo_iters);
}
void liblte_phy_turbo_decode(LIBLTE_PHY_STRUCT *p, float *d_bits, uint32 N_d_bits, uint32 N_fill_bits, uint8 *c_bits, uint32 *N_c_bits)
{
    turbo_decode(p, d_bits, N_d_bits, N_fill_bits, c_bits, N_c_bits);
}

// ---- code block (de)segmentation: bit shuffling on the host, 36.212 5.1.2 as the reference implements it
static const uint32 K_STEPS[4][3] = {{40, 512, 8}, {528, 1024, 16}, {1056, 2048, 32}, {2112, 6144, 64}};
static uint32 k_at_least(uint32 b)
{
    for (int r = 0; r < 4; r++) for (uint32 k = K_STEPS[r][0]; k <= K_STEPS[r][1]; k += K_STEPS[r][2]) if (k >= b) return k;
    return 0;
}
static uint32 k_below(uint32 kp)
{
    uint32 best = 0;
    for (int r = 0; r < 4; r++) for (uint32 k = K_STEPS[r][0]; k <= K_STEPS[r][1]; k += K_STEPS[r][2]) if (k < kp) best = k;
    return best;
}
static uint32 crc24b(const uint8 *bits, uint32 n)
{
    uint32 rem = 0;
    // calc_crc (liblte_phy.cc:9713-9743) ORs the whole byte into the register: the filler bits in front of the first code block are
    // TX_NULL_SYMB = 100, not 0, and the reference's parity bits of that block depend on it -- kept, bit for bit
    for (uint32 i = 0; i < n + 24; i++) { rem = (rem << 1) | (i < n ? (uint32)bits[i] : 0u); if (rem & 0x1000000u) rem ^= 0x1800063u; }
    return rem;
}
struct SegGeom { uint32 L, C, F, Kp, Km, Cp, Cm; };
static SegGeom seg_geom(uint32 B)
{
    SegGeom g;
    if (B <= 6144) { g.L = 0; g.C = 1; g.Kp = k_at_least(B); g.Km = 0; g.Cp = 1; g.Cm = 0; g.F = g.Kp - B; return g; }
    g.L = 24; g.C = (uint32)ceilf((float)B / (float)(6144 - 24));
    const uint32 Bp = B + g.C * g.L;
    g.Kp = 0;
    for (int r = 0; r < 4 && !g.Kp; r++) for (uint32 k = K_STEPS[r][0]; k <= K_STEPS[r][1]; k += K_STEPS[r][2]) if (g.C * k >= Bp) { g.Kp = k; break; }
    g.Km = k_below(g.Kp);
    g.Cm = (g.C * g.Kp - Bp) / (g.Kp - g.Km); g.Cp = g.C - g.Cm;
    g.F = g.Cp * g.Kp + g.Cm * g.Km - Bp;
    return g;
}
// liblte_phy.cc:9753-9865 (the "+24 to N_c_bits" quirk of the multi-block branch, SURVEY H7, is kept)
void liblte_phy_code_block_segmentation(uint8 *b_bits, uint32 N_b_bits, uint32 *N_codeblocks, uint32 *N_filler_bits, uint8 *c_bits, uint32 N_c_bits_max,
                                        uint32 *N_c_bits)
{
    const SegGeom g = seg_geom(N_b_bits);
    *N_codeblocks = g.C; *N_filler_bits = g.F;
    for (uint32 i = 0; i < g.F; i++) c_bits[i] = 100;    // TX_NULL_SYMB
    uint32 k = g.F, s = 0;
    for (uint32 r = 0; r < g.C; r++) {
        const uint32 Kr = r < g.Cm ? g.Km : g.Kp;
        N_c_bits[r] = Kr;
        while (k < Kr - g.L) c_bits[r * N_c_bits_max + k++] = b_bits[s++];
        if (g.C > 1) {
            const uint32 p = crc24b(&c_bits[r * N_c_bits_max], N_c_bits[r]);
            N_c_bits[r] += g.L;
            while (k < Kr) { c_bits[r * N_c_bits_max + k] = (uint8)((p >> (23 - (k + g.L - Kr))) & 1u); k++; }
        }
        k = 0;
    }
}
// liblte_phy.cc:9875-9987
void liblte_phy_code_block_desegmentation(uint8 *c_bits, uint32 *N_c_bits, uint32 N_c_bits_max, uint32 tbs, uint8 *b_bits, uint32 N_b_bits)
{
    (void)N_c_bits; (void)N_b_bits;
    const SegGeom g = seg_geom(tbs + 24);
    uint32 k = g.F, s = 0;
    for (uint32 r = 0; r < g.C; r++) {
        const uint32 Kr = r < g.Cm ? g.Km : g.Kp;
        while (k < Kr - g.L) b_bits[s++] = c_bits[r * N_c_bits_max + k++];
        k = 0;
    }
}

/* TEST INFRASTRUCTURE ONLY -- not part of the product path.
 *
 * extern "C" facade over the UNMODIFIED reference sources
 * (/root/reference/liblte/src/liblte_phy.cc + liblte_common.cc), compiled where
 * they lie by oracle/Makefile into oracle/_ref/libref_oracle.so.  Nothing from
 * the reference is copied into this repository: this file only *calls* it.
 *
 * Conventions:
 *   H12  the reference malloc()s its 46 MB scratch struct and reads parts of it
 *        before writing; oracle/Makefile compiles liblte_phy.cc with
 *        -Wl,--wrap=malloc so the struct starts zeroed, and
 *        ref_turbo_decode() clears td_in_calc_2/3 before every call.  With
 *        that the reference is a pure function of its inputs.
 *   H13  built with -fsigned-char (int8 is plain char there).
 *   H8   file-local functions of liblte_phy.cc have external linkage; their
 *        prototypes are re-declared below exactly as at the cited lines.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <atomic>
#include <chrono>
#include <thread>
#include <vector>
#include "liblte_phy.h"

extern "C" void *__wrap_malloc(size_t n) { return calloc(1, n ? n : 1); } /* -Wl,--wrap=malloc */

/* ---- prototypes of file-local (but externally linked) reference functions ---- */
void turbo_encode(LIBLTE_PHY_STRUCT *, uint8 *, uint32, uint32, uint8 *, uint32 *);            /* liblte_phy.cc:1490 */
void turbo_decode(LIBLTE_PHY_STRUCT *, float *, uint32, uint32, uint8 *, uint32 *);            /* liblte_phy.cc:1530 */
void samples_to_symbols_dl(LIBLTE_PHY_STRUCT *, float *, float *, uint32, uint32, uint8, float *, float *); /* :1282 */
void generate_prs_c(uint32, uint32, uint32 *);                                                /* liblte_phy.cc:9669 */
void generate_crs(uint32, uint32, uint32, uint32, float *, float *);                          /* liblte_phy.cc:8300 */
void calc_crc(uint8 *, uint32, uint32, uint8 *, uint32);                                      /* liblte_phy.cc:9713 */
void modulation_mapper(uint8 *, uint32, LIBLTE_PHY_MODULATION_TYPE_ENUM, float *, float *, uint32 *);   /* :8704 */
void modulation_demapper(float *, float *, uint32, LIBLTE_PHY_MODULATION_TYPE_ENUM, int8 *, uint32 *);  /* :9502 */
void pre_decoder_and_matched_filter_dl(float *, float *, float *, float *, uint32, uint32, uint8,
                                       LIBLTE_PHY_PRE_CODER_TYPE_ENUM, float *, float *, uint32 *);      /* :7645 */
void layer_demapper_dl(float *, float *, uint32, uint8, uint32, LIBLTE_PHY_PRE_CODER_TYPE_ENUM,
                       float *, float *, uint32 *);                                                      /* :7473 */
LIBLTE_ERROR_ENUM dlsch_channel_decode(LIBLTE_PHY_STRUCT *, float *, uint32, uint32, uint32, uint32, uint32,
                                       uint32, uint8 *, uint32 *);                                       /* :12762 */
void dlsch_channel_encode(LIBLTE_PHY_STRUCT *, uint8 *, uint32, uint32, uint32, uint32, uint32, uint32,
                          uint32, uint32, uint32, uint8 *, uint32 *);                                    /* :12664 */
void viterbi_decode_siso(LIBLTE_PHY_STRUCT *, int8 *, uint32, uint32, uint32, uint32 *, int8 *, uint32 *); /* :10341 */

void rate_unmatch_conv(LIBLTE_PHY_STRUCT *, float *, uint32, uint32, float *, uint32 *);                  /* :1625 */
void viterbi_decode(LIBLTE_PHY_STRUCT *, float *, uint32, uint32, uint32, uint32 *, uint8 *, uint32 *);     /* :1445 */
LIBLTE_ERROR_ENUM dci_1a_unpack(uint8 *, uint32, LIBLTE_PHY_DCI_CA_PRESENCE_ENUM, uint16, uint32, uint8,
                                LIBLTE_PHY_ALLOCATION_STRUCT *);                                           /* :2021 */
LIBLTE_ERROR_ENUM dci_1c_unpack(uint8 *, uint32, uint16, uint32, uint8, LIBLTE_PHY_ALLOCATION_STRUCT *);    /* :2061 */
void dci_1a_pack(LIBLTE_PHY_ALLOCATION_STRUCT *, LIBLTE_PHY_DCI_CA_PRESENCE_ENUM, uint32, uint8, uint8 *, uint32 *);   /* :1997 */
void dci_channel_encode(LIBLTE_PHY_STRUCT *, uint8 *, uint32, uint16, uint8, uint32, uint8 *);                          /* :1905 */
void bch_channel_encode(LIBLTE_PHY_STRUCT *, uint8 *, uint32, uint8, uint8 *, uint32 *);                                /* :1821 */
extern uint32 TBS_71721[27][110];                                                                          /* :477 */
extern uint32 TBS_71723[32];                                                                               /* :749 */

extern "C" {

/* ------------------------------------------------------------------ context */
void *ref_phy_init(int fs_enum, int N_id_cell, int N_ant, int N_rb_dl, float phich_res)
{
    LIBLTE_PHY_STRUCT *p = NULL;
    if (LIBLTE_SUCCESS != liblte_phy_init(&p, (LIBLTE_PHY_FS_ENUM)fs_enum, (uint16)N_id_cell, (uint8)N_ant,
                                          (uint32)N_rb_dl, LIBLTE_PHY_N_SC_RB_DL_NORMAL_CP, phich_res))
        return NULL;
    return p;
}
void ref_phy_cleanup(void *h) { if (h) liblte_phy_cleanup((LIBLTE_PHY_STRUCT *)h); }
int  ref_phy_update_n_rb_dl(void *h, int N_rb_dl) { return (int)liblte_phy_update_n_rb_dl((LIBLTE_PHY_STRUCT *)h, N_rb_dl); }
unsigned long ref_sizeof_phy_struct(void) { return sizeof(LIBLTE_PHY_STRUCT); }
unsigned long ref_sizeof_subframe_struct(void) { return sizeof(LIBLTE_PHY_SUBFRAME_STRUCT); }

/* ---------------------------------------------------------- transmit-side channel coding */
/* dci_1a_pack (liblte_phy.cc:13153-13250) + dci_channel_encode (:12882-12942) at aggregation level 4 (288 bits), as
 * liblte_phy_pdcch_channel_encode calls them (:4301-4316).  Returns the payload size; *tbs = what dci_1a_pack stored in the allocation. */
int ref_dci_1a_encode(void *h, int rnti, int mcs, int N_prb, int prb0, int rv_idx, int N_rb_dl, uint8 *out288, int *tbs)
{
    LIBLTE_PHY_ALLOCATION_STRUCT a;
    memset(&a, 0, sizeof(a));
    a.rnti = (uint16)rnti; a.mcs = (uint8)mcs; a.N_prb = (uint32)N_prb; a.rv_idx = (uint32)rv_idx;
    for (int j = 0; j < N_prb && j < 110; j++) { a.prb[0][j] = prb0 + j; a.prb[1][j] = prb0 + j; }
    uint8 bits[128];
    uint32 n = 0;
    dci_1a_pack(&a, LIBLTE_PHY_DCI_CA_NOT_PRESENT, (uint32)N_rb_dl, 1, bits, &n);
    dci_channel_encode((LIBLTE_PHY_STRUCT *)h, bits, n, (uint16)rnti, 0, 288, out288);
    *tbs = (int)a.tbs;
    return (int)n;
}
/* bch_channel_encode (liblte_phy.cc:12510-12571): 24 MIB bits -> 1920 coded bits, BEFORE the scrambling of liblte_phy_bch_channel_encode (:3894-3899) */
int ref_bch_encode_bits(void *h, const uint8 *mib24, int N_ant, uint8 *out1920)
{
    uint32 n = 0;
    bch_channel_encode((LIBLTE_PHY_STRUCT *)h, (uint8 *)mib24, 24, (uint8)N_ant, out1920, &n);
    return (int)n;
}

/* liblte_phy_update_n_rb_dl (liblte_phy.cc:2592-2652) on a blank struct with the given samples per symbol: out = {error code, N_rb_dl, FFT_size, FFT_pad_size} */
void ref_update_n_rb_dl_probe(unsigned N_samps_per_symb, unsigned N_rb_dl, unsigned *out)
{
    LIBLTE_PHY_STRUCT *p = (LIBLTE_PHY_STRUCT *)calloc(1, sizeof(LIBLTE_PHY_STRUCT));
    p->N_samps_per_symb = N_samps_per_symb; p->N_sc_rb_dl = 12;
    out[0] = (unsigned)liblte_phy_update_n_rb_dl(p, N_rb_dl);
    out[1] = p->N_rb_dl; out[2] = p->FFT_size; out[3] = p->FFT_pad_size;
    free(p);
}

/* ------------------------------------------------- code block (de)segmentation */
/* liblte_phy.cc:9753-9865 / 9875-9987, called exactly as declared (liblte_phy.h:1337-1362) */
void ref_code_block_segmentation(const uint8 *b_bits, unsigned N_b_bits, unsigned *N_codeblocks, unsigned *N_filler_bits, uint8 *c_bits,
                                 unsigned N_c_bits_max, unsigned *N_c_bits)
{
    liblte_phy_code_block_segmentation((uint8 *)b_bits, N_b_bits, N_codeblocks, N_filler_bits, c_bits, N_c_bits_max, N_c_bits);
}
void ref_code_block_desegmentation(const uint8 *c_bits, unsigned *N_c_bits, unsigned N_c_bits_max, unsigned tbs, uint8 *b_bits, unsigned N_b_bits)
{
    liblte_phy_code_block_desegmentation((uint8 *)c_bits, N_c_bits, N_c_bits_max, tbs, b_bits, N_b_bits);
}

/* ------------------------------------------------------------------- turbo */
/* d_out: branch-major d[x*(K+4)+i], 3*(K+4) bytes (liblte_phy.cc:10571-10586) */
int ref_turbo_encode(void *h, const uint8 *c_bits, int K, uint8 *d_out)
{
    uint32 N_d = 0;
    turbo_encode((LIBLTE_PHY_STRUCT *)h, (uint8 *)c_bits, (uint32)K, 0, d_out, &N_d);
    return (int)N_d;
}
/* d_bits: interleaved d[3i+x], N_d_bits = 3*(K+4) floats, MODIFIED IN PLACE (sentinels zeroed,
 * liblte_phy.cc:10636-10642).  Intermediates are exported for stage-level parity checks. */
void ref_turbo_decode(void *h, float *d_bits, int N_d_bits, uint8 *c_bits)
{
    LIBLTE_PHY_STRUCT *p = (LIBLTE_PHY_STRUCT *)h;
    uint32 N_c = 0;
    memset(p->td_in_calc_2, 0, sizeof(p->td_in_calc_2));  /* H12 convention */
    memset(p->td_in_calc_3, 0, sizeof(p->td_in_calc_3));
    turbo_decode(p, d_bits, (uint32)N_d_bits, 0, c_bits, &N_c);
}
/* which: 0 in_act_1, 1 in_calc_1, 2 int_act_1, 3 int_act_2, 4 in_calc_2, 5 in_calc_3, 6 fb_1, 7 in_int, 8 in_int_1 */
void ref_turbo_get_intermediate(void *h, int which, int K, signed char *out)
{
    LIBLTE_PHY_STRUCT *p = (LIBLTE_PHY_STRUCT *)h;
    int8 *src = NULL;
    switch (which) {
    case 0: src = p->td_in_act_1; break;  case 1: src = p->td_in_calc_1; break;
    case 2: src = p->td_int_act_1; break; case 3: src = p->td_int_act_2; break;
    case 4: src = p->td_in_calc_2; break; case 5: src = p->td_in_calc_3; break;
    case 6: src = p->td_fb_1; break;      case 7: src = p->td_in_int; break;
    case 8: src = p->td_in_int_1; break;
    }
    if (src) memcpy(out, src, (size_t)K);
}
void ref_viterbi_decode_siso(void *h, signed char *d_bits, int N_d_bits, signed char *out, int *N_out)
{
    uint32 g[2] = {015, 013}; uint32 n = 0;
    viterbi_decode_siso((LIBLTE_PHY_STRUCT *)h, (int8 *)d_bits, (uint32)N_d_bits, 4, 2, g, (int8 *)out, &n);
    *N_out = (int)n;
}

/* -------------------------------------------------------- rate (un)matching */
void ref_rate_match_turbo(void *h, const uint8 *d_bits, int N_d_bits, int N_codeblocks, int tx_mode, int N_soft,
                          int M_dl_harq, int chan_type, int rv_idx, int N_e_bits, uint8 *e_bits)
{
    liblte_phy_rate_match_turbo((LIBLTE_PHY_STRUCT *)h, (uint8 *)d_bits, N_d_bits, N_codeblocks, tx_mode, N_soft,
                                M_dl_harq, (LIBLTE_PHY_CHAN_TYPE_ENUM)chan_type, rv_idx, N_e_bits, e_bits);
}
int ref_rate_unmatch_turbo(void *h, const float *e_bits, int N_e_bits, int N_dummy_bits, int N_codeblocks, int tx_mode,
                           int N_soft, int M_dl_harq, int chan_type, int rv_idx, float *d_bits)
{
    /* dummy bit VALUES are irrelevant (uint8 never equals 10000); zeros as in dlsch_channel_decode */
    static uint8 dummy[3 * LIBLTE_PHY_MAX_CODE_BLOCK_SIZE];
    uint32 N_d = 0;
    liblte_phy_rate_unmatch_turbo((LIBLTE_PHY_STRUCT *)h, (float *)e_bits, N_e_bits, dummy, N_dummy_bits, N_codeblocks,
                                  tx_mode, N_soft, M_dl_harq, (LIBLTE_PHY_CHAN_TYPE_ENUM)chan_type, rv_idx, d_bits, &N_d);
    return (int)N_d;
}

/* ------------------------------------------------------ sequences and CRCs */
void ref_generate_prs_c(unsigned c_init, int len, unsigned *c) { generate_prs_c(c_init, (uint32)len, (uint32 *)c); }
void ref_generate_crs(int N_s, int L, int N_id_cell, float *crs_re, float *crs_im)
{
    generate_crs(N_s, L, N_id_cell, LIBLTE_PHY_N_SC_RB_DL_NORMAL_CP, crs_re, crs_im);
}
/* which: 0 CRC24A, 1 CRC24B, 2 CRC16 */
void ref_calc_crc(const uint8 *a, int N_a, int which, uint8 *p_bits)
{
    uint32 poly = which == 0 ? 0x01864CFB : which == 1 ? 0x01800063 : 0x00011021;
    calc_crc((uint8 *)a, (uint32)N_a, poly, p_bits, which == 2 ? 16 : 24);
}

/* ------------------------------------------------------------ (de)mapping */
int ref_modulation_mapper(const uint8 *bits, int N_bits, int mod_type, float *d_re, float *d_im)
{
    uint32 M = 0;
    modulation_mapper((uint8 *)bits, N_bits, (LIBLTE_PHY_MODULATION_TYPE_ENUM)mod_type, d_re, d_im, &M);
    return (int)M;
}
int ref_modulation_demapper(const float *d_re, const float *d_im, int M_symb, int mod_type, signed char *bits)
{
    uint32 N = 0;
    modulation_demapper((float *)d_re, (float *)d_im, M_symb, (LIBLTE_PHY_MODULATION_TYPE_ENUM)mod_type, (int8 *)bits, &N);
    return (int)N;
}
/* h_re/h_im: [N_ant][h_len]; x out: layer-major (liblte_phy.cc:7672-7677) */
int ref_pre_decoder_dl(const float *y_re, const float *y_im, const float *h_re, const float *h_im, int h_len,
                       int M_ap_symb, int N_ant, float *x_re, float *x_im)
{
    uint32 M = 0;
    pre_decoder_and_matched_filter_dl((float *)y_re, (float *)y_im, (float *)h_re, (float *)h_im, h_len, M_ap_symb,
                                      (uint8)N_ant, LIBLTE_PHY_PRE_CODER_TYPE_TX_DIVERSITY, x_re, x_im, &M);
    return (int)M;
}
int ref_layer_demapper_dl(const float *x_re, const float *x_im, int M_layer_symb, int N_ant, float *d_re, float *d_im)
{
    uint32 M = 0;
    layer_demapper_dl((float *)x_re, (float *)x_im, M_layer_symb, (uint8)N_ant, 1, LIBLTE_PHY_PRE_CODER_TYPE_TX_DIVERSITY,
                      d_re, d_im, &M);
    return (int)M;
}

/* ------------------------------------------------------- OFDM demod and CE */
void ref_samples_to_symbols_dl(void *h, const float *i_samps, const float *q_samps, unsigned slot_start_idx,
                               int symbol_offset, float *symb_re, float *symb_im)
{
    samples_to_symbols_dl((LIBLTE_PHY_STRUCT *)h, (float *)i_samps, (float *)q_samps, slot_start_idx, symbol_offset, 0,
                          symb_re, symb_im);
}
void *ref_subframe_new(void) { return calloc(1, sizeof(LIBLTE_PHY_SUBFRAME_STRUCT)); }
void  ref_subframe_free(void *s) { free(s); }
void  ref_subframe_clear(void *s) { memset(s, 0, sizeof(LIBLTE_PHY_SUBFRAME_STRUCT)); }
void  ref_subframe_set_num(void *s, int num) { ((LIBLTE_PHY_SUBFRAME_STRUCT *)s)->num = num; }
/* field: 0 rx_symb_re [16][1200], 1 rx_symb_im, 2 rx_ce_re [4][16][1200], 3 rx_ce_im, 4 tx_symb_re [4][16][1200], 5 tx_symb_im */
float *ref_subframe_field(void *s, int field)
{
    LIBLTE_PHY_SUBFRAME_STRUCT *sf = (LIBLTE_PHY_SUBFRAME_STRUCT *)s;
    switch (field) {
    case 0: return &sf->rx_symb_re[0][0];   case 1: return &sf->rx_symb_im[0][0];
    case 2: return &sf->rx_ce_re[0][0][0];  case 3: return &sf->rx_ce_im[0][0][0];
    case 4: return &sf->tx_symb_re[0][0][0]; case 5: return &sf->tx_symb_im[0][0][0];
    }
    return NULL;
}
int ref_get_dl_subframe_and_ce(void *h, const float *i_samps, const float *q_samps, unsigned frame_start_idx,
                               int subfr_num, int N_id_cell, int N_ant, void *subframe)
{
    return (int)liblte_phy_get_dl_subframe_and_ce((LIBLTE_PHY_STRUCT *)h, (float *)i_samps, (float *)q_samps,
                                                  frame_start_idx, (uint8)subfr_num, N_id_cell, (uint8)N_ant,
                                                  (LIBLTE_PHY_SUBFRAME_STRUCT *)subframe);
}

/* ------------------------------------------------------------------- PDSCH */
static void fill_alloc(LIBLTE_PHY_ALLOCATION_STRUCT *a, int mod_type, int N_prb, const int *prb, int tbs, int rv_idx,
                       int rnti, int tx_mode)
{
    memset(a, 0, sizeof(*a));
    a->pre_coder_type = LIBLTE_PHY_PRE_CODER_TYPE_TX_DIVERSITY;
    a->mod_type       = (LIBLTE_PHY_MODULATION_TYPE_ENUM)mod_type;
    a->chan_type      = LIBLTE_PHY_CHAN_TYPE_DLSCH;
    a->tbs = tbs; a->rv_idx = rv_idx; a->N_prb = N_prb;
    for (int j = 0; j < N_prb; j++) { a->prb[0][j] = prb[j]; a->prb[1][j] = prb[j]; }
    a->N_codewords = 1; a->N_layers = 1; a->tx_mode = tx_mode; a->rnti = (uint16)rnti;
}
/* Encodes ONE allocation into subframe->tx_symb (liblte_phy.cc:3489).  Caller keeps G <= 10000 (H6). */
int ref_pdsch_channel_encode(void *h, void *subframe, int N_pdcch_symbs, int N_id_cell, int N_ant, int mod_type,
                             int N_prb, const int *prb, int tbs, int rv_idx, int rnti, int tx_mode, const uint8 *tb_bits)
{
    static LIBLTE_PHY_PDCCH_STRUCT pdcch;
    memset(&pdcch, 0, sizeof(pdcch));
    pdcch.N_symbs = N_pdcch_symbs; pdcch.N_alloc = 1;
    fill_alloc(&pdcch.alloc[0], mod_type, N_prb, prb, tbs, rv_idx, rnti, tx_mode);
    pdcch.alloc[0].msg[0].N_bits = tbs;
    memcpy(pdcch.alloc[0].msg[0].msg, tb_bits, (size_t)tbs);
    return (int)liblte_phy_pdsch_channel_encode((LIBLTE_PHY_STRUCT *)h, &pdcch, N_id_cell, (uint8)N_ant,
                                                (LIBLTE_PHY_SUBFRAME_STRUCT *)subframe);
}
int ref_pdsch_channel_decode(void *h, void *subframe, int N_pdcch_symbs, int N_id_cell, int N_ant, int mod_type,
                             int N_prb, const int *prb, int tbs, int rv_idx, int rnti, int tx_mode, uint8 *out_bits,
                             int *N_out_bits)
{
    static LIBLTE_PHY_ALLOCATION_STRUCT a;
    uint32 n = 0;
    fill_alloc(&a, mod_type, N_prb, prb, tbs, rv_idx, rnti, tx_mode);
    int err = (int)liblte_phy_pdsch_channel_decode((LIBLTE_PHY_STRUCT *)h, (LIBLTE_PHY_SUBFRAME_STRUCT *)subframe, &a,
                                                   N_pdcch_symbs, N_id_cell, (uint8)N_ant, out_bits, &n);
    *N_out_bits = (int)n;
    return err;
}
/* scratch views after a ref_pdsch_channel_decode call: 0 pdsch_soft_bits(int8) 1 pdsch_descramb_bits(f32)
 * 2 pdsch_d_re 3 pdsch_d_im 4 pdsch_y_est_re 5 pdsch_y_est_im 6 pdsch_c_est_re[0] 7 pdsch_c_est_im[0] 8 dlsch_rx_d_bits */
void *ref_pdsch_scratch(void *h, int which)
{
    LIBLTE_PHY_STRUCT *p = (LIBLTE_PHY_STRUCT *)h;
    switch (which) {
    case 0: return p->pdsch_soft_bits; case 1: return p->pdsch_descramb_bits;
    case 2: return p->pdsch_d_re; case 3: return p->pdsch_d_im;
    case 4: return p->pdsch_y_est_re; case 5: return p->pdsch_y_est_im;
    case 6: return p->pdsch_c_est_re[0]; case 7: return p->pdsch_c_est_im[0];
    case 8: return p->dlsch_rx_d_bits;
    }
    return NULL;
}
int ref_dlsch_channel_decode(void *h, const float *in_bits, int N_in_bits, int tbs, int tx_mode, int rv_idx,
                             uint8 *out_bits, int *N_out_bits)
{
    LIBLTE_PHY_STRUCT *p = (LIBLTE_PHY_STRUCT *)h;
    uint32 n = 0;
    memset(p->td_in_calc_2, 0, sizeof(p->td_in_calc_2));
    memset(p->td_in_calc_3, 0, sizeof(p->td_in_calc_3));
    int err = (int)dlsch_channel_decode(p, (float *)in_bits, N_in_bits, tbs, tx_mode, rv_idx, 8, 250368, out_bits, &n);
    *N_out_bits = (int)n;
    return err;
}
/* G coded bits for one transport block, through the reference's own DL-SCH encoder (C=1 only, H7). */
int ref_dlsch_channel_encode(void *h, const uint8 *tb_bits, int tbs, int tx_mode, int rv_idx, int G, int Q_m, uint8 *out_bits)
{
    uint32 n = 0;
    dlsch_channel_encode((LIBLTE_PHY_STRUCT *)h, (uint8 *)tb_bits, tbs, tbs, tx_mode, rv_idx, G, 2, Q_m, 8, 250368, out_bits, &n);
    return (int)n;
}

/* "Big allocation" PDSCH receive: liblte_phy_pdsch_channel_decode's own orchestration (liblte_phy.cc:3722-3849)
 * driven through the reference's STAGE functions with caller-sized buffers, because its fixed scratch
 * (5000 REs / 10000 bits, liblte_phy.h:349-363) overruns for G > 10000 (H6).  Single antenna port.
 * The RE walk below is the loop at liblte_phy.cc:3745-3802 for N_ant == 1 in a non-0/5 subframe or any
 * subframe when the PRBs avoid the centre band; it is the only logic restated here.
 * The scratch is caller-owned (BigScratch) so that a timing loop performs no allocation per transport block. */
struct BigScratch {
    int     maxre;
    float  *f;      /* 8 planes of maxre floats */
    int8   *soft;
    uint32 *c;
    float  *desc;
};
static void big_scratch_init(BigScratch *s, int N_prb)
{
    s->maxre = 12 * 14 * N_prb;
    s->f     = (float *)malloc(sizeof(float) * (size_t)s->maxre * 8);
    s->soft  = (int8 *)malloc((size_t)s->maxre * 6);
    s->c     = (uint32 *)malloc(sizeof(uint32) * (size_t)s->maxre * 6);
    s->desc  = (float *)malloc(sizeof(float) * (size_t)s->maxre * 6);
}
static void big_scratch_free(BigScratch *s) { free(s->f); free(s->soft); free(s->c); free(s->desc); }

static int pdsch_decode_big_core(LIBLTE_PHY_STRUCT *p, LIBLTE_PHY_SUBFRAME_STRUCT *sf, BigScratch *s, int N_pdcch_symbs, int N_id_cell,
                                 int mod_type, int N_prb, const int *prb, int tbs, int rv_idx, int rnti, int tx_mode, uint8 *out_bits,
                                 int *N_out_bits, signed char *soft_out)
{
    const int maxre = s->maxre;
    float *y_re = s->f, *y_im = y_re + maxre, *c_re = y_im + maxre, *c_im = c_re + maxre;
    float *x_re = c_im + maxre, *x_im = x_re + maxre, *d_re = x_im + maxre, *d_im = d_re + maxre;
    int8 *soft = s->soft; uint32 *c = s->c; float *desc = s->desc;
    uint32 idx = 0, M_layer = 0, M_symb = 0, N_bits = 0, n_out = 0;
    uint32 first_sc, last_sc;
    if (p->N_rb_dl == 6) { first_sc = 0; last_sc = 71; }
    else if (p->N_rb_dl == 15) { first_sc = 54; last_sc = 125; }
    else if (p->N_rb_dl == 25) { first_sc = 114; last_sc = 185; }
    else if (p->N_rb_dl == 50) { first_sc = 264; last_sc = 335; }
    else if (p->N_rb_dl == 75) { first_sc = 414; last_sc = 485; }
    else { first_sc = 564; last_sc = 635; }
    for (uint32 L = N_pdcch_symbs; L < 14; L++)
        for (int pi = 0; pi < N_prb; pi++) {
            uint32 i = prb[pi];
            for (uint32 j = 0; j < 12; j++) {
                uint32 sc = i * 12 + j;
                if ((L % 7) == 0 && (uint32)(N_id_cell % 6) == (j % 6)) continue;
                if ((L % 7) == 4 && (uint32)((N_id_cell + 3) % 6) == (j % 6)) continue;
                if (sf->num == 0 && sc >= first_sc && sc <= last_sc && L >= 7 && L <= 10) continue;
                if ((sf->num == 0 || sf->num == 5) && sc >= first_sc && sc <= last_sc && (L == 6 || L == 5)) continue;
                y_re[idx] = sf->rx_symb_re[L][sc]; y_im[idx] = sf->rx_symb_im[L][sc];
                c_re[idx] = sf->rx_ce_re[0][L][sc]; c_im[idx] = sf->rx_ce_im[0][L][sc];
                idx++;
            }
        }
    pre_decoder_and_matched_filter_dl(y_re, y_im, c_re, c_im, maxre, idx, 1, LIBLTE_PHY_PRE_CODER_TYPE_TX_DIVERSITY,
                                      x_re, x_im, &M_layer);
    layer_demapper_dl(x_re, x_im, M_layer, 1, 1, LIBLTE_PHY_PRE_CODER_TYPE_TX_DIVERSITY, d_re, d_im, &M_symb);
    modulation_demapper(d_re, d_im, M_symb, (LIBLTE_PHY_MODULATION_TYPE_ENUM)mod_type, soft, &N_bits);
    uint32 c_init = ((uint32)rnti << 14) | (0 << 13) | (sf->num << 9) | (uint32)N_id_cell;
    generate_prs_c(c_init, N_bits, c);
    for (uint32 k = 0; k < N_bits; k++) desc[k] = (float)soft[k] * (1 - 2 * (float)c[k]);
    if (soft_out) memcpy(soft_out, soft, N_bits);
    memset(p->td_in_calc_2, 0, sizeof(p->td_in_calc_2));
    memset(p->td_in_calc_3, 0, sizeof(p->td_in_calc_3));
    int err = (int)dlsch_channel_decode(p, desc, N_bits, tbs, tx_mode, rv_idx, 8, 250368, out_bits, &n_out);
    *N_out_bits = (int)n_out;
    return err == LIBLTE_SUCCESS ? 0 : (int)LIBLTE_ERROR_DECODE_FAIL;
}
int ref_pdsch_decode_big(void *h, void *subframe, int N_pdcch_symbs, int N_id_cell, int mod_type, int N_prb,
                         const int *prb, int tbs, int rv_idx, int rnti, int tx_mode, uint8 *out_bits, int *N_out_bits,
                         signed char *soft_out /* may be NULL */)
{
    BigScratch s;
    big_scratch_init(&s, N_prb);
    int err = pdsch_decode_big_core((LIBLTE_PHY_STRUCT *)h, (LIBLTE_PHY_SUBFRAME_STRUCT *)subframe, &s, N_pdcch_symbs, N_id_cell, mod_type,
                                    N_prb, prb, tbs, rv_idx, rnti, tx_mode, out_bits, N_out_bits, soft_out);
    big_scratch_free(&s);
    return err;
}

/* Native thread pool over subframes for the CPU arm of bench.py: n_threads std::threads, ONE LIBLTE_PHY_STRUCT,
 * subframe struct and scratch per thread, all created SERIALLY before the clock starts (FFT planning is not thread safe); subframes are
 * handed out through an atomic counter.  Per subframe: liblte_phy_get_dl_subframe_and_ce, then the PDSCH receive of every allocation.
 *   alloc_desc: n_alloc records of {mod_type, tbs, rnti, N_prb, prb[0..110)} (114 ints each)
 *   want_tb   : expected transport blocks [seg][n_alloc][tb_stride] (NULL: only the CRC verdict counts); subframe b is compared with
 *               want_tb[b % seg]
 * Returns the number of transport blocks that failed (CRC or bits); *seconds = wall time of the parallel region,
 * stage_s[0..1] = summed per-thread seconds in the front end and in the PDSCH receive. */
int ref_chain_pool(int n_threads, int fs_enum, int N_rb_dl, int N_id_cell, int cfi, const float *i_samps, const float *q_samps,
                   int n_sf, const long long *start_idx, const int *subfr_num, int n_alloc, const int *alloc_desc,
                   const uint8 *want_tb, int seg, int tb_stride, double *seconds, double *stage_s)
{
    struct Worker { LIBLTE_PHY_STRUCT *phy; LIBLTE_PHY_SUBFRAME_STRUCT *sf; BigScratch scr; uint8 *bits; int fails; double t_front, t_pdsch; };
    if (n_threads < 1) n_threads = 1;
    std::vector<Worker> w(n_threads);
    int max_prb = 1;
    for (int a = 0; a < n_alloc; a++) if (alloc_desc[114 * a + 3] > max_prb) max_prb = alloc_desc[114 * a + 3];
    for (int t = 0; t < n_threads; t++) {
        w[t].phy = NULL;
        if (LIBLTE_SUCCESS != liblte_phy_init(&w[t].phy, (LIBLTE_PHY_FS_ENUM)fs_enum, (uint16)N_id_cell, 1, (uint32)N_rb_dl,
                                              LIBLTE_PHY_N_SC_RB_DL_NORMAL_CP, 1.0f))
            return -1;
        w[t].sf = (LIBLTE_PHY_SUBFRAME_STRUCT *)calloc(1, sizeof(LIBLTE_PHY_SUBFRAME_STRUCT));
        big_scratch_init(&w[t].scr, max_prb);
        w[t].bits = (uint8 *)malloc(LIBLTE_PHY_MAX_CODE_BLOCK_SIZE + 64);
        w[t].fails = 0; w[t].t_front = w[t].t_pdsch = 0.0;
    }
    const uint32 n_sf_samps = w[0].phy->N_samps_per_subfr;
    std::atomic<int> next(0), ready(0);
    std::atomic<bool> go(false);
    auto body = [&](int t) {
        Worker &me = w[t];
        if (n_sf > 0) {
            /* untimed warm-up on this thread: one subframe through both stages, so that the first-touch page faults of the 44 MB scratch
             * struct (zero pages until written) and the cold instruction/data caches are not charged to the reference */
            const int sn = subfr_num[0];
            const uint32 fstart = (uint32)((unsigned long long)start_idx[0] - (unsigned long long)sn * n_sf_samps);
            liblte_phy_get_dl_subframe_and_ce(me.phy, (float *)i_samps, (float *)q_samps, fstart, (uint8)sn, N_id_cell, 1, me.sf);
            const int *d = alloc_desc;
            int n_out = 0;
            pdsch_decode_big_core(me.phy, me.sf, &me.scr, cfi, N_id_cell, d[0], d[3], d + 4, d[1], 0, d[2], 1, me.bits, &n_out, NULL);
        }
        ready.fetch_add(1);
        while (!go.load(std::memory_order_acquire)) std::this_thread::yield();
        for (;;) {
            const int b = next.fetch_add(1);
            if (b >= n_sf) break;
            const int sn = subfr_num[b];
            /* liblte_phy_get_dl_subframe_and_ce addresses a subframe as frame_start_idx + subfr_num*N (uint32 arithmetic) */
            const uint32 fstart = (uint32)((unsigned long long)start_idx[b] - (unsigned long long)sn * n_sf_samps);
            auto t0 = std::chrono::steady_clock::now();
            liblte_phy_get_dl_subframe_and_ce(me.phy, (float *)i_samps, (float *)q_samps, fstart, (uint8)sn, N_id_cell, 1, me.sf);
            auto t1 = std::chrono::steady_clock::now();
            for (int a = 0; a < n_alloc; a++) {
                const int *d = alloc_desc + 114 * a;
                int n_out = 0;
                int err = pdsch_decode_big_core(me.phy, me.sf, &me.scr, cfi, N_id_cell, d[0], d[3], d + 4, d[1], 0, d[2], 1, me.bits, &n_out, NULL);
                if (err != 0) me.fails++;
                else if (want_tb && memcmp(me.bits, want_tb + ((size_t)(b % seg) * n_alloc + a) * tb_stride, (size_t)d[1]) != 0) me.fails++;
            }
            auto t2 = std::chrono::steady_clock::now();
            me.t_front += std::chrono::duration<double>(t1 - t0).count();
            me.t_pdsch += std::chrono::duration<double>(t2 - t1).count();
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++) th.emplace_back(body, t);
    while (ready.load() < n_threads) std::this_thread::yield();
    auto T0 = std::chrono::steady_clock::now();
    go.store(true, std::memory_order_release);
    for (auto &x : th) x.join();
    auto T1 = std::chrono::steady_clock::now();
    *seconds = std::chrono::duration<double>(T1 - T0).count();
    int fails = 0;
    double tf = 0, tp = 0;
    for (int t = 0; t < n_threads; t++) {
        fails += w[t].fails; tf += w[t].t_front; tp += w[t].t_pdsch;
        liblte_phy_cleanup(w[t].phy); free(w[t].sf); big_scratch_free(&w[t].scr); free(w[t].bits);
    }
    if (stage_s) { stage_s[0] = tf; stage_s[1] = tp; }
    return fails;
}

/* ----------------------------------------------------------------- TX side */
int ref_map_crs(void *h, void *subframe, int N_id_cell, int N_ant)
{
    return (int)liblte_phy_map_crs((LIBLTE_PHY_STRUCT *)h, (LIBLTE_PHY_SUBFRAME_STRUCT *)subframe, N_id_cell, (uint8)N_ant);
}
int ref_map_pss(void *h, void *subframe, int N_id_2, int N_ant)
{
    return (int)liblte_phy_map_pss((LIBLTE_PHY_STRUCT *)h, (LIBLTE_PHY_SUBFRAME_STRUCT *)subframe, N_id_2, (uint8)N_ant);
}
int ref_map_sss(void *h, void *subframe, int N_id_1, int N_id_2, int N_ant)
{
    return (int)liblte_phy_map_sss((LIBLTE_PHY_STRUCT *)h, (LIBLTE_PHY_SUBFRAME_STRUCT *)subframe, N_id_1, N_id_2, (uint8)N_ant);
}
int ref_create_dl_subframe(void *h, void *subframe, int ant, float *i_samps, float *q_samps)
{
    return (int)liblte_phy_create_dl_subframe((LIBLTE_PHY_STRUCT *)h, (LIBLTE_PHY_SUBFRAME_STRUCT *)subframe, (uint8)ant,
                                              i_samps, q_samps);
}

/* -------------------------------------------------------------------- sync */
int ref_find_coarse_timing(void *h, const float *i_samps, const float *q_samps, int N_slots, float *freq_offset /*[5]*/,
                           unsigned *symb_starts /*[5][7]*/, int *n_corr_peaks)
{
    LIBLTE_PHY_COARSE_TIMING_STRUCT t;
    memset(&t, 0, sizeof(t));
    int err = (int)liblte_phy_dl_find_coarse_timing_and_freq_offset((LIBLTE_PHY_STRUCT *)h, (float *)i_samps,
                                                                    (float *)q_samps, N_slots, &t);
    memcpy(freq_offset, t.freq_offset, sizeof(t.freq_offset));
    memcpy(symb_starts, t.symb_starts, sizeof(t.symb_starts));
    *n_corr_peaks = (int)t.n_corr_peaks;
    return err;
}
float *ref_timing_abs_corr(void *h) { return ((LIBLTE_PHY_STRUCT *)h)->dl_timing_abs_corr; }
int ref_find_pss(void *h, const float *i_samps, const float *q_samps, unsigned *symb_starts /*[7] in/out*/, unsigned *N_id_2,
                 unsigned *pss_symb, float *pss_thresh, float *freq_offset)
{
    return (int)liblte_phy_find_pss_and_fine_timing((LIBLTE_PHY_STRUCT *)h, (float *)i_samps, (float *)q_samps,
                                                    (uint32 *)symb_starts, (uint32 *)N_id_2, (uint32 *)pss_symb, pss_thresh,
                                                    freq_offset);
}
int ref_find_sss(void *h, const float *i_samps, const float *q_samps, unsigned N_id_2, unsigned *symb_starts, float pss_thresh,
                 unsigned *N_id_1, unsigned *frame_start_idx)
{
    return (int)liblte_phy_find_sss((LIBLTE_PHY_STRUCT *)h, (float *)i_samps, (float *)q_samps, N_id_2, (uint32 *)symb_starts,
                                    pss_thresh, (uint32 *)N_id_1, (uint32 *)frame_start_idx);
}

/* ------------------------------------------------ control channels (callers) */
int ref_bch_channel_decode(void *h, void *subframe, int N_id_cell, int *N_ant, uint8 *out_bits, int *N_out_bits, int *offset)
{
    uint8 na = 0, off = 0; uint32 n = 0;
    int err = (int)liblte_phy_bch_channel_decode((LIBLTE_PHY_STRUCT *)h, (LIBLTE_PHY_SUBFRAME_STRUCT *)subframe, N_id_cell,
                                                 &na, out_bits, &n, &off);
    *N_ant = na; *N_out_bits = (int)n; *offset = off;
    return err;
}
/* Returns the decoded allocations as flat ints: per alloc {rnti, tbs, mcs, N_prb, rv_idx, mod_type, tx_mode, chan_type, prb[0..N_prb)} */
int ref_pdcch_channel_decode(void *h, void *subframe, int N_id_cell, int N_ant, float phich_res, int phich_dur, int *cfi,
                             int *N_alloc, int *alloc_out /* [6][8+110] */)
{
    static LIBLTE_PHY_PCFICH_STRUCT pcfich;
    static LIBLTE_PHY_PHICH_STRUCT  phich;
    static LIBLTE_PHY_PDCCH_STRUCT  pdcch;
    memset(&pdcch, 0, sizeof(pdcch));
    int err = (int)liblte_phy_pdcch_channel_decode((LIBLTE_PHY_STRUCT *)h, (LIBLTE_PHY_SUBFRAME_STRUCT *)subframe, N_id_cell,
                                                   (uint8)N_ant, phich_res, (LIBLTE_RRC_PHICH_DURATION_ENUM)phich_dur,
                                                   &pcfich, &phich, &pdcch);
    *cfi = (int)pdcch.N_symbs; *N_alloc = (int)pdcch.N_alloc;
    for (uint32 a = 0; a < pdcch.N_alloc && a < 6; a++) {
        int *o = alloc_out + a * (8 + 110);
        o[0] = pdcch.alloc[a].rnti; o[1] = pdcch.alloc[a].tbs; o[2] = pdcch.alloc[a].mcs; o[3] = pdcch.alloc[a].N_prb;
        o[4] = pdcch.alloc[a].rv_idx; o[5] = pdcch.alloc[a].mod_type; o[6] = pdcch.alloc[a].tx_mode; o[7] = pdcch.alloc[a].chan_type;
        for (uint32 j = 0; j < pdcch.alloc[a].N_prb; j++) o[8 + j] = pdcch.alloc[a].prb[0][j];
    }
    return err;
}

/* transmit side of the control channels, used to build fixtures */
int ref_bch_channel_encode(void *h, void *subframe, const uint8 *bits, int N_bits, int N_id_cell, int N_ant, unsigned sfn)
{
    return (int)liblte_phy_bch_channel_encode((LIBLTE_PHY_STRUCT *)h, (uint8 *)bits, N_bits, N_id_cell, (uint8)N_ant,
                                              (LIBLTE_PHY_SUBFRAME_STRUCT *)subframe, sfn);
}
/* alloc_in: per allocation {rnti, mcs, rv_idx, N_prb, first prb}; DCI 1A, localized, aggregation level 4 (all the reference can send) */
int ref_pdcch_channel_encode(void *h, void *subframe, int N_id_cell, int N_ant, float phich_res, int phich_dur, int cfi, int n_alloc,
                             const int *alloc_in)
{
    static LIBLTE_PHY_PCFICH_STRUCT pcfich;
    static LIBLTE_PHY_PHICH_STRUCT  phich;
    static LIBLTE_PHY_PDCCH_STRUCT  pdcch;
    memset(&pcfich, 0, sizeof(pcfich)); memset(&phich, 0, sizeof(phich)); memset(&pdcch, 0, sizeof(pdcch));
    pcfich.cfi = cfi;
    pdcch.N_alloc = n_alloc;
    for (int a = 0; a < n_alloc; a++) {
        const int *in = alloc_in + 5 * a;
        LIBLTE_PHY_ALLOCATION_STRUCT *al = &pdcch.alloc[a];
        al->chan_type = LIBLTE_PHY_CHAN_TYPE_DLSCH; al->rnti = (uint16)in[0]; al->mcs = (uint8)in[1]; al->rv_idx = in[2]; al->N_prb = in[3];
        for (int j = 0; j < in[3]; j++) { al->prb[0][j] = in[4] + j; al->prb[1][j] = in[4] + j; }
        al->pre_coder_type = LIBLTE_PHY_PRE_CODER_TYPE_TX_DIVERSITY; al->mod_type = LIBLTE_PHY_MODULATION_TYPE_QPSK;
        al->N_codewords = 1; al->tx_mode = N_ant == 1 ? 1 : 2;
    }
    return (int)liblte_phy_pdcch_channel_encode((LIBLTE_PHY_STRUCT *)h, &pcfich, &phich, &pdcch, N_id_cell, (uint8)N_ant, phich_res,
                                                (LIBLTE_RRC_PHICH_DURATION_ENUM)phich_dur, (LIBLTE_PHY_SUBFRAME_STRUCT *)subframe);
}
/* decode, full detail: per alloc {rnti, tbs, mcs, N_prb, rv_idx, mod_type, tx_mode, chan_type, prb[0][0..110), prb[1][0..110)}; also the PCFICH CFI */
int ref_pdcch_channel_decode2(void *h, void *subframe, int N_id_cell, int N_ant, float phich_res, int phich_dur, int *cfi, int *N_symbs,
                              int *N_alloc, int *alloc_out /* [6][8+220] */)
{
    static LIBLTE_PHY_PCFICH_STRUCT pcfich;
    static LIBLTE_PHY_PHICH_STRUCT  phich;
    static LIBLTE_PHY_PDCCH_STRUCT  pdcch;
    memset(&pcfich, 0, sizeof(pcfich)); memset(&phich, 0, sizeof(phich)); memset(&pdcch, 0, sizeof(pdcch));
    int err = (int)liblte_phy_pdcch_channel_decode((LIBLTE_PHY_STRUCT *)h, (LIBLTE_PHY_SUBFRAME_STRUCT *)subframe, N_id_cell,
                                                   (uint8)N_ant, phich_res, (LIBLTE_RRC_PHICH_DURATION_ENUM)phich_dur,
                                                   &pcfich, &phich, &pdcch);
    *cfi = (int)pcfich.cfi; *N_symbs = (int)pdcch.N_symbs; *N_alloc = (int)pdcch.N_alloc;
    for (uint32 a = 0; a < pdcch.N_alloc && a < 6; a++) {
        int *o = alloc_out + a * (8 + 220);
        o[0] = pdcch.alloc[a].rnti; o[1] = pdcch.alloc[a].tbs; o[2] = pdcch.alloc[a].mcs; o[3] = pdcch.alloc[a].N_prb;
        o[4] = pdcch.alloc[a].rv_idx; o[5] = pdcch.alloc[a].mod_type; o[6] = pdcch.alloc[a].tx_mode; o[7] = pdcch.alloc[a].chan_type;
        for (uint32 j = 0; j < pdcch.alloc[a].N_prb && j < 110; j++) { o[8 + j] = pdcch.alloc[a].prb[0][j]; o[8 + 110 + j] = pdcch.alloc[a].prb[1][j]; }
    }
    return err;
}
/* pieces, for checking the restatement function by function */
void ref_rate_unmatch_conv(void *h, const float *e_bits, int N_e, int N_c, float *d_bits, int *N_d)
{
    uint32 n = 0;
    rate_unmatch_conv((LIBLTE_PHY_STRUCT *)h, (float *)e_bits, N_e, N_c, d_bits, &n);
    *N_d = (int)n;
}
void ref_viterbi_decode(void *h, const float *d_bits, int N_d, uint8 *c_bits, int *N_c)
{
    uint32 g[3] = {0133, 0171, 0165}, n = 0;
    viterbi_decode((LIBLTE_PHY_STRUCT *)h, (float *)d_bits, N_d, 7, 3, g, c_bits, &n);
    *N_c = (int)n;
}
int ref_dci_1a_unpack(const uint8 *bits, int N_bits, int rnti, int N_rb_dl, int N_ant, int *o /*[8+220]*/)
{
    static LIBLTE_PHY_ALLOCATION_STRUCT al; memset(&al, 0, sizeof(al));
    int err = (int)dci_1a_unpack((uint8 *)bits, N_bits, LIBLTE_PHY_DCI_CA_NOT_PRESENT, (uint16)rnti, N_rb_dl, (uint8)N_ant, &al);
    o[0] = al.rnti; o[1] = al.tbs; o[2] = al.mcs; o[3] = al.N_prb; o[4] = al.rv_idx; o[5] = al.mod_type; o[6] = al.tx_mode; o[7] = al.chan_type;
    for (uint32 j = 0; j < al.N_prb && j < 110; j++) { o[8 + j] = al.prb[0][j]; o[8 + 110 + j] = al.prb[1][j]; }
    return err;
}
int ref_dci_1c_unpack(const uint8 *bits, int N_bits, int rnti, int N_rb_dl, int N_ant, int *o /*[8+220]*/)
{
    static LIBLTE_PHY_ALLOCATION_STRUCT al; memset(&al, 0, sizeof(al));
    int err = (int)dci_1c_unpack((uint8 *)bits, N_bits, (uint16)rnti, N_rb_dl, (uint8)N_ant, &al);
    o[0] = al.rnti; o[1] = al.tbs; o[2] = al.mcs; o[3] = al.N_prb; o[4] = al.rv_idx; o[5] = al.mod_type; o[6] = al.tx_mode; o[7] = al.chan_type;
    for (uint32 j = 0; j < al.N_prb && j < 110; j++) { o[8 + j] = al.prb[0][j]; o[8 + 110 + j] = al.prb[1][j]; }
    return err;
}
int ref_tbs_71721(int i_tbs, int n_prb) { return (int)TBS_71721[i_tbs][n_prb - 1]; }
int ref_tbs_71723(int i) { return (int)TBS_71723[i]; }

} /* extern "C" */

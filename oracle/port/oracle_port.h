/* TEST INFRASTRUCTURE ONLY -- never linked, imported or executed by the product path.
 *
 * CPU restatement ("port") of the reference's DL receive hot path, one function per row of
 * SURVEY.md section 8(a).  Each function cites the reference lines it follows.  The port is pinned
 * against the reference itself: tests/test_oracle_port.py compares every function here with
 * oracle/_ref/libref_oracle.so (the reference's own liblte_phy.cc compiled unmodified) wherever that
 * library is available, and against fixtures under tests/golden/ generated from it
 * (tests/golden/make_golden.py) everywhere else.
 *
 * Float discipline: built with -ffp-contract=off, no fast-math, -fsigned-char.
 */
#ifndef ORACLE_PORT_H
#define ORACLE_PORT_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define OP_RX_NULL 10000.0f

/* ---- turbo */
int  op_qpp_params(int K, uint32_t *f1, uint32_t *f2);                 /* 0 if K is a legal block size */
void op_set_qpp_spec(int on);                                        /* 1: 36.212 permutation without the uint32 wrap (library option LTEB200_QPP_SPEC) */
void op_qpp_pi(int K, uint32_t *pi);                                   /* 32-bit-overflowing pi(i), liblte_phy.cc:10982 */
void op_turbo_encode(const uint8_t *c, int K, uint8_t *d_branch_major);/* liblte_phy.cc:10541-10586 */
/* d: interleaved d[3i+x], 3(K+4) floats, modified in place like the reference; out: K bytes.
 * dbg (optional): 6*K int8 = act1, calc1, actA, actB, calc2, calc3 */
void op_turbo_decode_ref(float *d, int K, uint8_t *out, int8_t *dbg);
void op_siso(const int8_t *a, const int8_t *b, int K, int8_t *act);    /* viterbi_decode_siso, rate 2, g={015,013} */

/* ---- max-log-MAP ("mode B"; no reference counterpart): restated so that the CUDA kernel has a
 *      bit-exact CPU twin.  See port/turbo_mlm.cc for the numeric definition. */
void op_turbo_decode_mlm(const float *d, int K, int n_iter, uint8_t *out);
/* independent float, full-block (non-windowed) max-log-MAP: bounds the loss of the windowed int16 definition (port/turbo_mlm_float.cc) */
void op_turbo_decode_mlm_float(const float *d, int K, int n_iter, uint8_t *out);

/* ---- rate unmatching (a13) */
int  op_rate_unmatch_turbo(const float *e, int E, int D /*K+4*/, int C, int tx_mode, int N_soft, int M_dl_harq,
                           int chan_type, int rv, float *d_out /* 3*D interleaved */);
void op_rate_match_turbo(const uint8_t *d_branch_major, int D, int C, int tx_mode, int N_soft, int M_dl_harq,
                         int chan_type, int rv, int E, uint8_t *e);

/* ---- sequences, CRC (a11, a12) */
void     op_prs_c(uint32_t c_init, int len, uint8_t *c);               /* liblte_phy.cc:9669-9704 */
uint32_t op_crc(const uint8_t *bits, int n, uint32_t poly, int n_p);   /* liblte_phy.cc:9713-9743 */
void     op_segmentation_params(int tbs, int *K, int *F, int *C);      /* liblte_phy.cc:9753-9865, C==1 branch */

/* ---- demapping chain (a8-a11) */
void op_zf_1ant(const float *y_re, const float *y_im, const float *h_re, const float *h_im, int n, float *x_re, float *x_im);
int  op_pre_decoder_dl(const float *y_re, const float *y_im, const float *h_re, const float *h_im, int h_len, int M, int N_ant,
                       float *d_re, float *d_im);
int  op_demap(const float *d_re, const float *d_im, int M, int mod_type, int8_t *bits);
/* ---- soft symbol level of the max-log-MAP path (no reference counterpart: twin of pdsch_llr.cu's LTEB200_DEMAP_MAXLOG branch, port/softdemap.cc) */
float op_noise_var_crs(const float *h_re, const float *h_im, int nsc, int vshift);
void  op_soft_bits_1ant(const float *y_re, const float *y_im, const float *h_re, const float *h_im, int M, int mod, float s2, const uint8_t *cseq, int8_t *soft);
void  op_soft_bits_2ant(const float *d_re, const float *d_im, const float *h0_re, const float *h0_im, const float *h1_re, const float *h1_im, int M, int mod, float s2,
                        const uint8_t *cseq, int8_t *soft);
/* ---- RE extraction (a7): writes indices (symbol*1200+sc) of the PDSCH REs of one allocation */
int  op_pdsch_re_indices(int N_rb_dl, int N_id_cell, int N_ant, int subfr_num, int N_pdcch_symbs, int N_prb,
                         const int *prb_slot0, const int *prb_slot1, int *idx);

/* ---- OFDM demod + CE (a5, a6) */
void op_fft_forward(int n, const float *in_re, const float *in_im, float *out_re, float *out_im);
void op_samples_to_symbols_dl(const float *i_samps, const float *q_samps, int fs_idx, int N_rb_dl, uint32_t slot_start,
                              int symbol_offset, float *symb_re, float *symb_im);
void op_freq_shift(float *i_buf, float *q_buf, uint32_t start_idx, uint32_t num_samps, float freq_offset, uint32_t fs);  /* fs.cc:696-713 */
void op_crs(int N_s, int L, int N_id_cell, float *re220, float *im220);
void op_get_dl_subframe_and_ce(const float *i_samps, const float *q_samps, int fs_idx, int N_rb_dl, uint32_t frame_start_idx,
                               int subfr_num, int N_id_cell, int N_ant, float *rx_symb_re /*[16][1200]*/,
                               float *rx_symb_im, float *rx_ce_re /*[4][16][1200]*/, float *rx_ce_im);
/* CE only, from a given grid */
void op_chan_est(const float *rx_symb_re, const float *rx_symb_im, int N_rb_dl, int subfr_num, int N_id_cell, int N_ant,
                 float *rx_ce_re, float *rx_ce_im);

/* ---- whole PDSCH receive of one allocation, C == 1 (a7-a14 composed as liblte_phy.cc:3690-3853) */
int op_pdsch_decode(const float *rx_symb_re, const float *rx_symb_im, const float *rx_ce_re, const float *rx_ce_im,
                    int N_rb_dl, int N_id_cell, int subfr_num, int N_pdcch_symbs, int mod_type, int N_prb, const int *prb,
                    int tbs, int rv, int rnti, int turbo_mode, int n_iter, uint8_t *out_bits);

int op_pdsch_decode_nant(const float *rx_symb_re, const float *rx_symb_im, const float *rx_ce_re, const float *rx_ce_im, int N_rb_dl,
                         int N_id_cell, int N_ant, int subfr_num, int N_pdcch_symbs, int mod_type, int N_prb, const int *prb, int tbs, int rv,
                         int rnti, int turbo_mode, int n_iter, uint8_t *out_bits, float *soft_out, int *n_soft);

/* ---- sync (a2-a4) */
int op_find_coarse_timing(const float *i_samps, const float *q_samps, int fs_idx, int N_slots, float *freq_offset /*[5]*/,
                          uint32_t *symb_starts /*[5][7]*/, int *n_corr_peaks);
int op_find_pss(const float *i_samps, const float *q_samps, int fs_idx, int N_rb_dl, uint32_t *symb_starts /*[7] io*/,
                uint32_t *N_id_2, uint32_t *pss_symb, float *pss_thresh, float *freq_offset);
int op_find_sss(const float *i_samps, const float *q_samps, int fs_idx, int N_rb_dl, uint32_t N_id_2, uint32_t *symb_starts,
                float pss_thresh, uint32_t *N_id_1, uint32_t *frame_start_idx);

/* control channels (control.cc) */
int  op_tbs_1a(int i_tbs, int n_prb_1a);
int  op_tbs_1c(int i);
int  op_rate_unmatch_conv(const float *e, int E, int N_c, float *d);
int  op_viterbi_decode(const float *d, int N_d, uint8_t *c);
int  op_bch_decode_bits(const float *in1920, int N_ant, uint8_t *out24);
int  op_bch_channel_decode(const float *rx_re, const float *rx_im, const float *ce_re, const float *ce_im, int N_rb_dl, int N_id_cell,
                           int *N_ant, uint8_t *out24, int *offset);
int  op_dci_sizes(int N_rb_dl, int *size_1a, int *size_1c);
void op_control_reg_positions(int N_rb_dl, int N_id_cell, float phich_res, int *pcfich_k, int *phich_k, int *N_reg_phich);
int  op_pdcch_cce_map(int N_rb_dl, int N_id_cell, int N_ant, int N_symbs, float phich_res, int *re_of, int *N_reg_out);
int  op_dci_decode_bits(const float *soft, int E, int n_dci, uint8_t *dci_bits, int *rnti);
int  op_dci_1a_unpack(const uint8_t *bits, int rnti, int N_rb_dl, int N_ant, int *alloc228);
int  op_dci_1c_unpack(const uint8_t *bits, int rnti, int N_rb_dl, int N_ant, int *alloc228);
int  op_pdcch_channel_decode(const float *rx_re, const float *rx_im, const float *ce_re, const float *ce_im, int N_rb_dl, int N_id_cell,
                             int N_ant, int subfr, float phich_res, int *cfi, int *N_symbs, int *N_alloc, int *allocs, int *raw_found,
                             int *raw_rnti, uint8_t *raw_bits);
#ifdef __cplusplus
}
#endif
#endif

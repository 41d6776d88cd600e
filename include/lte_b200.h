/* libopenlte_b200.so -- C ABI of the B200-native LTE downlink receive hot path.
 *
 * Drop-in boundary for the functions OpenLTE's LTE_fdd_dl_file_scan exercises in liblte_phy
 * (reference: /root/reference/liblte/hdr/liblte_phy.h, /root/reference/liblte/src/liblte_phy.cc).
 * Plain pointers and sizes only -- no CUDA, torch or C++ types -- so that any FFI (C++, ctypes, cgo, JNI)
 * can bind it.  Two families of entry points:
 *
 *   lteb200_<stage>_batch(...)   DEVICE pointers, caller's CUDA stream, many subframes / code blocks per call.
 *                                This is the throughput path (SURVEY.md 8b "what the new library adds").
 *   lteb200_<stage>(...)         HOST pointers, one unit per call, same argument meaning as the liblte_phy
 *                                function named beside each declaration.  include/liblte_phy_b200.h wraps
 *                                these under the reference's own C++ names.
 *
 * Conventions shared with the reference: soft bits are LLR-like floats with POSITIVE = bit 0 and the
 * sentinel 10000.0f = "not received" (liblte_phy.cc:1620); hard bits are one bit per byte; I/Q is two
 * separate float arrays (re, im) at the host boundary (liblte_phy.h:1170-1177) and interleaved float2
 * (or int8 pairs, the capture file format) in the batched device API.
 *
 * All functions return LTEB200_OK (0) or a negative LTEB200_ERR_*; lteb200_last_error() gives the text.
 * There is NO CPU fallback: without a CUDA device every compute entry point fails with LTEB200_ERR_CUDA.
 * Functions marked HOST below are table builders that need neither a device nor a context.
 *
 * Threads and streams.  A context is used by ONE thread at a time, like a LIBLTE_PHY_STRUCT (its table caches, scratch arena and
 * last-error text are not locked); different contexts may be used from different threads concurrently (the one process-wide cache,
 * the per-allocation index tables, is locked).  *_batch calls enqueue their kernels on the caller's stream and return without
 * waiting for them; the first call with a new block size, cell or allocation also builds that table (a synchronous upload).  The
 * host-pointer twins and lteb200_dl_receive_host return when their results are in the caller's buffers.
 */
#ifndef LTE_B200_H
#define LTE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LTEB200_OK            0
#define LTEB200_ERR_ARGS     (-1)   /* maps to LIBLTE_ERROR_INVALID_INPUTS (liblte_common.h:60-66) */
#define LTEB200_ERR_CUDA     (-2)
#define LTEB200_ERR_DECODE   (-3)   /* maps to LIBLTE_ERROR_DECODE_FAIL / INVALID_CRC */
#define LTEB200_ERR_UNSUPPORTED (-4)
#define LTEB200_ERR_CONTENTS (-5)   /* maps to LIBLTE_ERROR_INVALID_CONTENTS */

#define LTEB200_RX_NULL_BIT  10000.0f

/* sample-rate index == LIBLTE_PHY_FS_ENUM (liblte_phy.h:196-203) */
enum { LTEB200_FS_1_92 = 0, LTEB200_FS_3_84, LTEB200_FS_7_68, LTEB200_FS_15_36, LTEB200_FS_30_72 };
/* == LIBLTE_PHY_MODULATION_TYPE_ENUM (liblte_phy.h:211-216) */
enum { LTEB200_MOD_BPSK = 0, LTEB200_MOD_QPSK, LTEB200_MOD_16QAM, LTEB200_MOD_64QAM };
/* turbo decoder selection */
enum {
    LTEB200_TURBO_REF_EXACT = 0, /* bit-exact twin of turbo_decode(), liblte_phy.cc:10620-10845 (3 SISO Viterbi passes) */
    LTEB200_TURBO_MAXLOGMAP = 1  /* n_iter-iteration max-log-MAP (BASELINE.json north_star); no reference counterpart */
};
/* I/Q sample formats of the batched API */
enum { LTEB200_IQ_F32 = 0 /* interleaved float re,im */, LTEB200_IQ_I8 = 1 /* interleaved int8 re,im (capture file) */ };

typedef struct lteb200_ctx lteb200_ctx;

/* ---------------------------------------------------------------------------------------------- context */
/* Replaces liblte_phy_init / liblte_phy_cleanup (liblte_phy.h:606-612, :638) as far as device state goes:
 * tables and scratch live in HBM behind this handle.  device < 0 selects the current CUDA device. */
int         lteb200_create(lteb200_ctx **ctx, int device);
void        lteb200_destroy(lteb200_ctx *ctx);
const char *lteb200_last_error(lteb200_ctx *ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
unsigned long long lteb200_launch_count(lteb200_ctx *ctx);
const char *lteb200_version(void);

/* ------------------------------------------------------------------------------------------------ turbo */
/* Batched turbo decode.  d_llr: DEVICE [n_cb][3*(K+4)] floats, interleaved d[3i+x] exactly as
 * liblte_phy_rate_unmatch_turbo emits and turbo_decode() consumes (liblte_phy.cc:11485, 10662); NOT modified.
 * d_bits: DEVICE [n_cb][K] bytes, one bit each.  mode: LTEB200_TURBO_*.  n_iter: mode 1 only.
 * Replaces: turbo_decode(), liblte_phy.cc:10620 (prototype :1530-1535). */
int lteb200_turbo_decode_batch(lteb200_ctx *ctx, const float *d_llr, int K, int n_cb, int mode, int n_iter,
                               uint8_t *d_bits, void *stream);
/* Host twin of turbo_decode(phy_struct, d_bits, N_d_bits, N_fill_bits, c_bits, N_c_bits): d_bits is modified in
 * place (sentinels zeroed) like the reference; N_c_bits is left untouched like the reference (H8). */
int lteb200_turbo_decode(lteb200_ctx *ctx, float *d_bits, uint32_t N_d_bits, uint32_t N_fill_bits, uint8_t *c_bits,
                         int mode, int n_iter);

/* ---------------------------------------------------------------------------------------- rate unmatching */
/* Batched liblte_phy_rate_unmatch_turbo (liblte_phy.h:1311-1323, liblte_phy.cc:11246-11490).
 * d_e: DEVICE [n_cb][e_stride] floats of which the first E are used; d_out: DEVICE [n_cb][3*(K+4)] floats. */
int lteb200_rate_unmatch_turbo_batch(lteb200_ctx *ctx, const float *d_e, int e_stride, int E, int K, int N_codeblocks,
                                     int tx_mode, int N_soft, int M_dl_harq, int chan_type, int rv_idx, int n_cb,
                                     float *d_out, void *stream);
/* HOST, no device needed: the gather map the kernels above work through, as the library builds it from the reference's geometry
 * (liblte_phy.cc:11283-11400).  first_k_out[3*i + x], 3*(K+4) entries: index of the first e-bit that lands on output d[3i+x] (-1: none
 * within one lap of the circular buffer); every further visit comes *period e-bits later.  For inspection and CPU tests. */
int lteb200_rate_unmatch_map(int K, int N_codeblocks, int tx_mode, int N_soft, int M_dl_harq, int chan_type, int rv_idx,
                             int32_t *first_k_out, int *period);
/* Host twin (same argument order as the reference; dummy_bits content is ignored, only N_dummy_bits = K+4 matters). */
int lteb200_rate_unmatch_turbo(lteb200_ctx *ctx, const float *e_bits, uint32_t N_e_bits, const uint8_t *dummy_bits,
                               uint32_t N_dummy_bits, uint32_t N_codeblocks, uint32_t tx_mode, uint32_t N_soft,
                               uint32_t M_dl_harq, int chan_type, uint32_t rv_idx, float *d_bits, uint32_t *N_d_bits);

/* --------------------------------------------------------------------------------- OFDM demod + CRS estimate */
/* Batched liblte_phy_get_dl_subframe_and_ce (liblte_phy.h:1170-1177, liblte_phy.cc:5905-6200).
 * d_iq: DEVICE sample stream (format iq_fmt), n_samps complex samples.  Subframe b starts at sample
 * start_idx[b] (HOST int64 array, n_sf entries) and has number subfr_num[b] (HOST int32 array, 0..9).  The stream
 * must extend 2 OFDM symbols past the last subframe (the estimator reads symbols 14-15, liblte_phy.cc:5946).
 * d_grid: DEVICE [n_sf][16][12*N_rb_dl] float2 (rx_symb);  d_ce: DEVICE [n_sf][N_ant][14][12*N_rb_dl] float2. */
int lteb200_get_dl_subframe_and_ce_batch(lteb200_ctx *ctx, const void *d_iq, int iq_fmt, long long n_samps, int fs_idx,
                                         int N_rb_dl, int N_id_cell, int N_ant, const long long *start_idx /* HOST */,
                                         const int *subfr_num /* HOST */, int n_sf, float *d_grid, float *d_ce, void *stream);
/* Host twin: arrays laid out as LIBLTE_PHY_SUBFRAME_STRUCT's rx_symb_re/im[16][1200], rx_ce_re/im[4][16][1200]
 * (liblte_phy.h:226-239).  i_samps/q_samps must hold n_samps floats each. */
int lteb200_get_dl_subframe_and_ce(lteb200_ctx *ctx, const float *i_samps, const float *q_samps, long long n_samps,
                                   int fs_idx, int N_rb_dl, uint32_t frame_start_idx, uint8_t subfr_num,
                                   uint32_t N_id_cell, uint8_t N_ant, float *rx_symb_re, float *rx_symb_im,
                                   float *rx_ce_re, float *rx_ce_im);

/* --------------------------------------------------------------------------------------------- PDSCH receive */
/* One allocation shape shared by every subframe of a batch (what LIBLTE_PHY_ALLOCATION_STRUCT carries,
 * liblte_phy.h:684-702, for the fields the receive path reads). */
typedef struct {
    int32_t mod_type;        /* LTEB200_MOD_* */
    int32_t N_prb;
    int32_t prb[2][110];     /* PRB indices per slot */
    int32_t tbs;
    int32_t rv_idx;
    int32_t tx_mode;
    int32_t rnti;
} lteb200_alloc;

/* HOST, no device needed: the resource elements of one allocation in the order the reference visits them (RE walk of
 * liblte_phy.cc:3745-3802 incl. the CRS, PBCH and PSS/SSS exclusions), idx_out[r] = symbol*(12*N_rb_dl) + subcarrier; *n_out = their number
 * (LTEB200_ERR_ARGS with *n_out set if capacity is too small).  This is the index list the symbol-level kernels gather through. */
int lteb200_pdsch_re_indices(int N_rb_dl, int N_id_cell, int N_ant, int subfr_num, int N_pdcch_symbs, const lteb200_alloc *alloc,
                             int32_t *idx_out, int capacity, int *n_out);

/* Batched liblte_phy_pdsch_channel_decode (liblte_phy.h:906-913, liblte_phy.cc:3690-3853) for N_ant == 1 and
 * single-code-block transport blocks (the envelope in which the reference itself is correct).
 * For every subframe b and allocation a: RE extraction, ZF, demap, descramble, rate unmatch, turbo decode,
 * CRC24A.  d_tb_bits: DEVICE [n_sf][n_alloc][tb_stride] bytes (first tbs valid);  d_crc_ok: DEVICE [n_sf][n_alloc]
 * int32 (1 = CRC24A pass, i.e. the reference would return LIBLTE_SUCCESS). */
int lteb200_pdsch_decode_batch(lteb200_ctx *ctx, const float *d_grid, const float *d_ce, int N_rb_dl, int N_id_cell,
                               int N_ant, const int *subfr_num /* HOST */, int n_sf, int N_pdcch_symbs,
                               const lteb200_alloc *allocs /* HOST */, int n_alloc, int turbo_mode, int n_iter,
                               uint8_t *d_tb_bits, int tb_stride, int32_t *d_crc_ok, void *stream);
/* The two halves of the max-log-MAP receive path as separate calls (what lteb200_pdsch_decode_batch runs back to back when every block size is a
 * multiple of 32).  All allocations of one call share one code block size K; at most 8 allocations.
 *   lteb200_pdsch_llr_q8_batch : RE extraction (liblte_phy.cc:3745-3802), ZF / transmit-diversity combining (:7645-7796), demapping (:9502-9660),
 *       descrambling (:3831-3836), rate unmatching (:11246-11490) and the quantisation at the head of turbo_decode (:10636-10664), one CTA per transport
 *       block.  d_q8: DEVICE, one int8 image per (allocation, subframe) at index a * n_sf + subframe, lteb200_q8_stride(K) bytes each:
 *       Lsys | Lpar1 | Lpar2 in the decoder's window-transposed order, then the 12 tail values.
 *   lteb200_turbo_decode_q8_batch : n_iter iterations of max-log-MAP on those images, then CRC24A and transport-block extraction
 *       (dlsch_channel_decode's tail, liblte_phy.cc:12841-12869); outputs as lteb200_pdsch_decode_batch. */
size_t lteb200_q8_stride(int K);   /* 0 if K is not handled by this path (not a multiple of 32) */
int lteb200_pdsch_llr_q8_batch(lteb200_ctx *ctx, const float *d_grid, const float *d_ce, int N_rb_dl, int N_id_cell, int N_ant,
                               const int *subfr_num /* HOST */, int n_sf, int N_pdcch_symbs, const lteb200_alloc *allocs /* HOST */, int n_alloc,
                               int8_t *d_q8, void *stream);
int lteb200_turbo_decode_q8_batch(lteb200_ctx *ctx, const int8_t *d_q8, int K, int n_cb /* = n_alloc * n_sf */, int n_iter,
                                  const lteb200_alloc *allocs /* HOST */, int n_alloc, int n_sf, uint8_t *d_tb_bits, int tb_stride,
                                  int32_t *d_crc_ok, void *stream);
/* Host twin of liblte_phy_pdsch_channel_decode for one allocation; returns LTEB200_ERR_DECODE on CRC failure and
 * leaves out_bits / N_out_bits untouched in that case, like the reference. */
int lteb200_pdsch_channel_decode(lteb200_ctx *ctx, const float *rx_symb_re, const float *rx_symb_im,
                                 const float *rx_ce_re, const float *rx_ce_im, int N_rb_dl, uint32_t subfr_num,
                                 const lteb200_alloc *alloc, uint32_t N_pdcch_symbs, uint32_t N_id_cell, uint8_t N_ant,
                                 int turbo_mode, int n_iter, uint8_t *out_bits, uint32_t *N_out_bits);

/* Stage-level device entry points (used by the parity tests and for profiling single kernels). */
/* samples_to_symbols_dl for 16 symbols per subframe (liblte_phy.cc:8593-8644, called at :5946-5957) */
int lteb200_ofdm_demod_batch(lteb200_ctx *ctx, const void *d_iq, int iq_fmt, long long n_samps, int fs_idx, int N_rb_dl,
                             const long long *start_idx /* HOST */, int n_sf, float *d_grid, void *stream);
/* the CRS estimator alone (liblte_phy.cc:5959-6193) from a resident grid */
int lteb200_chan_est_batch(lteb200_ctx *ctx, const float *d_grid, const int *subfr_num /* HOST */, int N_rb_dl, int N_id_cell,
                           int N_ant, int n_sf, float *d_ce, void *stream);
/* soft bits after RE extraction + ZF + demap + descramble: d_soft DEVICE [n_sf][soft_stride] floats (= pdsch_descramb_bits);
 * n_bits_out: HOST [n_sf] number of valid soft bits per subframe (subframes 0 and 5 lose REs to PBCH/PSS/SSS) */
int lteb200_pdsch_demod_batch(lteb200_ctx *ctx, const float *d_grid, const float *d_ce, int N_rb_dl, int N_id_cell,
                              int N_ant, const int *subfr_num /* HOST */, int n_sf, int N_pdcch_symbs, const lteb200_alloc *alloc,
                              float *d_soft, int soft_stride, int *n_bits_out /* HOST */, void *stream);

/* --------------------------------------------------------------------------------- whole chain, one call */
/* I/Q stream -> transport-block bits + CRC verdicts for n_sf subframes (OFDM demod, CE, and PDSCH receive of
 * every allocation), everything resident in HBM between stages. */
int lteb200_dl_receive_batch(lteb200_ctx *ctx, const void *d_iq, int iq_fmt, long long n_samps, int fs_idx, int N_rb_dl,
                             int N_id_cell, int N_ant, const long long *start_idx /* HOST */, const int *subfr_num /* HOST */, int n_sf,
                             int N_pdcch_symbs, const lteb200_alloc *allocs, int n_alloc, int turbo_mode, int n_iter,
                             uint8_t *d_tb_bits, int tb_stride, int32_t *d_crc_ok, void *stream);
/* Same through HOST buffers (pinned or pageable): copies I/Q in, results out, inside the call. */
int lteb200_dl_receive_host(lteb200_ctx *ctx, const void *h_iq, int iq_fmt, long long n_samps, int fs_idx, int N_rb_dl,
                            int N_id_cell, int N_ant, const long long *h_start_idx, const int *h_subfr_num, int n_sf,
                            int N_pdcch_symbs, const lteb200_alloc *allocs, int n_alloc, int turbo_mode, int n_iter,
                            uint8_t *h_tb_bits, int tb_stride, int32_t *h_crc_ok);
/* HOST, no device needed: the plan lteb200_dl_receive_host follows for a batch.  Chunk c = subframes [bounds_out[c], bounds_out[c+1]);
 * copy_end_out[c] = how many samples of the stream (cumulative, from sample 0) must have reached the device before chunk c's kernels
 * start: everything its last subframe reads, i.e. start + subframe + the two look-ahead symbols of the channel estimator -- with
 * gapped, duplicate or overlapping (non-decreasing) start indices as well.  K_first: code block size of the first allocation,
 * n_sm: multiprocessors of the device (148 on B200), capacity: entries available in bounds_out (n_chunks + 1 needed; *n_chunks_out is
 * set even when it does not fit).  For inspection and CPU tests. */
int lteb200_e2e_plan(int n_sf, int n_alloc, int K_first, int turbo_mode, int n_sm, int fs_idx, long long n_samps, const long long *h_start_idx,
                     int *bounds_out, long long *copy_end_out, int capacity, int *n_chunks_out);

/* ------------------------------------------------------------------------------------------------- sync */
/* Host twins of liblte_phy_dl_find_coarse_timing_and_freq_offset (liblte_phy.h:1134-1138, .cc:5697-5852),
 * liblte_phy_find_pss_and_fine_timing (liblte_phy.h:1068-1075, .cc:5306-5510) and liblte_phy_find_sss
 * (liblte_phy.h:1106-1113, .cc:5578-5686).  Outputs have the reference's meaning. */
int lteb200_find_coarse_timing_and_freq_offset(lteb200_ctx *ctx, const float *i_samps, const float *q_samps,
                                               long long n_samps, int fs_idx, uint32_t N_slots, float *freq_offset /*[5]*/,
                                               uint32_t *symb_starts /*[5][7]*/, uint32_t *n_corr_peaks);
int lteb200_find_pss_and_fine_timing(lteb200_ctx *ctx, const float *i_samps, const float *q_samps, long long n_samps,
                                     int fs_idx, int N_rb_dl, uint32_t *symb_starts /*[7] in/out*/, uint32_t *N_id_2,
                                     uint32_t *pss_symb, float *pss_thresh, float *freq_offset);
int lteb200_find_sss(lteb200_ctx *ctx, const float *i_samps, const float *q_samps, long long n_samps, int fs_idx,
                     int N_rb_dl, uint32_t N_id_2, uint32_t *symb_starts /*[7] in/out*/, float pss_thresh,
                     uint32_t *N_id_1, uint32_t *frame_start_idx);

/* ------------------------------------------------------------------------------------- control channels */
/* PBCH: batched liblte_phy_bch_channel_decode (liblte_phy.h:941-947, liblte_phy.cc:3968-4100).  d_grid: DEVICE [n][16][nsc]
 * float2 (subframe 0 of n frames), d_ce: DEVICE [n][4][14][nsc] float2 (four ports, as the scan estimates before it knows
 * N_ant).  Per subframe the 3 x 4 hypotheses (1/2/4 ports x position in the 40 ms period) are decoded in parallel and the
 * first hit in the reference's search order is reported. */
typedef struct {
    int32_t ok;              /* 1 if a hypothesis passed the CRC */
    int32_t N_ant;           /* 1, 2 or 4 */
    int32_t offset;          /* SFN mod 4 */
    uint8_t bits[24];        /* MIB, one bit per byte */
} lteb200_bch_result;
int lteb200_bch_decode_batch(lteb200_ctx *ctx, const float *d_grid, const float *d_ce, int N_rb_dl, int N_id_cell, int n,
                             lteb200_bch_result *d_out, void *stream);
/* Host twin: HOST planes laid out as LIBLTE_PHY_SUBFRAME_STRUCT (rx_symb_* [16][1200], rx_ce_* [4][16][1200]). */
int lteb200_bch_channel_decode(lteb200_ctx *ctx, const float *rx_symb_re, const float *rx_symb_im, const float *rx_ce_re,
                               const float *rx_ce_im, int N_rb_dl, uint32_t N_id_cell, uint8_t *N_ant, uint8_t *out_bits,
                               uint32_t *N_out_bits, uint8_t *offset);

/* PCFICH + PDCCH common search space: batched liblte_phy_pdcch_channel_decode (liblte_phy.h:1005-1013, liblte_phy.cc:4519-5134)
 * for N_ant == 1 (for 2 / 4 ports the reference indexes its channel-estimate scratch with the wrong stride and is undefined:
 * LTEB200_ERR_UNSUPPORTED).  Per subframe: CFI, then the reference's 12 blind decodes -- aggregation level 4 candidates 0-3,
 * level 8 candidates 0-1, each as DCI 1A then 1C -- with CRC16 checked against SI-, P- and RA-RNTIs.  Slot = 2*candidate + format.
 * d_ce: DEVICE [n_sf][1][14][nsc] float2; subfr_num: HOST array. */
typedef struct {
    int32_t cfi;             /* 0 = PCFICH did not decode (the reference's INVALID_CRC) */
    int32_t N_symbs;
    int32_t found[12];
    int32_t rnti[12];
    uint8_t bits[12][32];    /* DCI payload, one bit per byte */
} lteb200_pdcch_raw;
typedef struct {
    lteb200_alloc alloc;     /* what lteb200_pdsch_decode_batch consumes */
    int32_t mcs;
    int32_t format;          /* 0 = DCI 1A, 1 = DCI 1C */
} lteb200_dci;
int lteb200_pdcch_decode_batch(lteb200_ctx *ctx, const float *d_grid, const float *d_ce, int N_rb_dl, int N_id_cell, int N_ant,
                               float phich_res, const int *subfr_num, int n_sf, lteb200_pdcch_raw *d_out, void *stream);
/* HOST: where the common search space sits in the grid: out576[36*c + j] = symbol*(12*N_rb_dl) + subcarrier of element j of CCE c < 16
 * (REG walk, cyclic shift, REG de-interleaver and CCE assembly of liblte_phy.cc:4591-4838 folded into one table), -1 = no such CCE. */
int lteb200_pdcch_cce_map(int N_rb_dl, int N_id_cell, int N_ant, int N_symbs, float phich_res, int32_t *out576);
/* HOST: dci_1a_unpack / dci_1c_unpack (liblte_phy.cc:13273, 13400) over one subframe's raw results, in the reference's order,
 * at most 6 allocations; return value mirrors the reference's (see control.cu). */
int lteb200_pdcch_unpack(const lteb200_pdcch_raw *raw, int N_rb_dl, int N_ant, lteb200_dci *allocs, int *N_alloc);
/* Host twin of liblte_phy_pdcch_channel_decode: HOST planes as LIBLTE_PHY_SUBFRAME_STRUCT. */
int lteb200_pdcch_channel_decode(lteb200_ctx *ctx, const float *rx_symb_re, const float *rx_symb_im, const float *rx_ce_re,
                                 const float *rx_ce_im, int N_rb_dl, uint32_t subfr_num, uint32_t N_id_cell, uint8_t N_ant,
                                 float phich_res, int phich_dur, uint32_t *cfi, uint32_t *N_symbs, lteb200_dci *allocs,
                                 uint32_t *N_alloc);

/* The scan loop (LTE_fdd_dl_fs_samp_buf.cc:491-515) for n_sf subframes in one call: front end, PDCCH decode, and PDSCH decode of
 * every allocation a subframe's PDCCH announced (N_ant == 1).  d_iq: DEVICE samples; start_idx, subfr_num, results, tb_bits: HOST.
 * Transport-block bits (one per byte) are appended to tb_bits; results[b].tb_offset[a] is where allocation a of subframe b starts
 * (-1: CRC failed or nothing decodable).  Synchronises `stream` (the DCI contents decide what is launched next). */
typedef struct {
    int32_t     cfi, N_symbs, N_alloc;
    int32_t     pdcch_err;       /* return value of lteb200_pdcch_unpack for this subframe */
    lteb200_dci dci[6];
    int32_t     crc_ok[6];
    long long   tb_offset[6];
} lteb200_scan_result;
int lteb200_dl_scan_batch(lteb200_ctx *ctx, const void *d_iq, int iq_fmt, long long n_samps, int fs_idx, int N_rb_dl, int N_id_cell, int N_ant,
                          float phich_res, const long long *start_idx, const int *subfr_num, int n_sf, int turbo_mode, int n_iter,
                          lteb200_scan_result *results, uint8_t *tb_bits, long long tb_capacity, long long *tb_used, void *stream);

/* ------------------------------------------------------------------------------------------ transmit side */
/* Batched downlink transmitter for test-signal generation: liblte_phy_map_crs (liblte_phy.h:1031) +
 * liblte_phy_pdsch_channel_encode (liblte_phy.h:716-724, liblte_phy.cc:3489-3680) for every allocation + liblte_phy_create_dl_subframe
 * (liblte_phy.h:1109, liblte_phy.cc:5862) for n_sf subframes.  d_tb_bits: DEVICE [n_sf][n_alloc][tb_stride], one bit per byte.
 * d_grid (optional, DEVICE [n_sf][14][12*N_rb_dl] float2) receives the resource grids, d_iq (optional, DEVICE
 * [n_sf][1920 << fs_idx] float2) the time-domain subframes back to back.  One antenna port; transport blocks with TBS + 24 in the
 * turbo interleaver table (the reference's encoder produces no filler bits); the reference's wrapped interleaver. */
int lteb200_dl_transmit_batch(lteb200_ctx *ctx, const uint8_t *d_tb_bits, int tb_stride, const lteb200_alloc *allocs, int n_alloc, int fs_idx,
                              int N_rb_dl, int N_id_cell, int N_ant, const int *subfr_num, int n_sf, int N_pdcch_symbs, float *d_grid,
                              float *d_iq, void *stream);

/* The same with everything LTE_fdd_dl_file_gen puts on the air besides the PDSCH (LTE_fdd_dl_fg_samp_buf.cc:321-575): synchronisation signals
 * (liblte_phy_map_pss :5265, liblte_phy_map_sss :5520), PBCH (liblte_phy_bch_channel_encode :3863), PCFICH + PDCCH with one DCI 1A per allocation
 * (liblte_phy_pdcch_channel_encode :4113), and transmit diversity over 2 / 4 antenna ports (layer_mapper_dl :7293, pre_coder_dl :7526, CRS of every
 * port liblte_phy_map_crs :5144).  The grids are bit-identical to the reference transmitter's.
 *   d_grid (optional): DEVICE [n_sf][N_ant][14][12 N_rb_dl] float2;  d_iq (optional): DEVICE [n_sf][1920 << fs_idx] float2, the ports' signals
 *   summed with the flat per-port channel gains of `ctl`.
 * PCFICH/PDCCH: one antenna port only (like the receive side); at most four allocations (common search space, aggregation level 4), each QPSK with
 * contiguous PRBs, an SI/P/RA-RNTI and tbs = lteb200_dci_1a_tbs(mcs) -- what the reference's dci_1a_unpack understands. */
typedef struct {
    int32_t N_id_cell;
    int32_t N_ant;                       /* 1, 2 or 4 */
    float   ant_gain_re[4], ant_gain_im[4];
    int32_t sync;                        /* 1: PSS + SSS in subframes 0 and 5 */
    int32_t pbch;                        /* 1: PBCH in every subframe 0 */
    int32_t sfn_first;                   /* system frame number of the frame the first subframe of the batch belongs to */
    const uint8_t *mib_bits;             /* HOST [n_frames][24], one bit per byte; frame f of the batch begins at the f-th subframe number 0 (f = 0: the first, possibly partial one) */
    int32_t pdcch;                       /* 1: PCFICH + PDCCH */
    float   phich_res;                   /* Ng: 1/6, 1/2, 1 or 2 (PHICH groups are reserved, nothing is sent on them) */
    int32_t mcs[8];                      /* MCS field of each allocation's DCI */
} lteb200_tx_control;
int lteb200_dci_1a_tbs(int mcs);         /* 36.213 table 7.1.7.2.1-1, column N_PRB = 3 (what DCI 1A with an SI/P/RA-RNTI implies), -1 if mcs > 26 */
/* HOST, no device needed: the channel-coded control messages as the transmitter places them, for inspection and CPU tests.
 *   lteb200_dci_1a_encode: DCI format 1A of one allocation (contiguous PRBs from alloc->prb[0][0], alloc->rnti, alloc->rv_idx, MCS field mcs; HARQ process 0,
 *       localized VRBs -- dci_1a_pack, liblte_phy.cc:13153-13250) + CRC16 masked with the RNTI, tail-biting convolutional code and rate matching to the
 *       288 bits of aggregation level 4 (dci_channel_encode, :12882-12942).  Returns the transport block size the receiver will derive for an
 *       SI/P/RA-RNTI, -1 for other RNTIs (coded all the same) and for arguments the format cannot carry.
 *   lteb200_bch_encode: 24 MIB bits -> CRC16 with the antenna mask, convolutional code, rate matching to 1920 bits (bch_channel_encode, :12510-12571)
 *       and scrambling with the cell's Gold sequence (liblte_phy_bch_channel_encode, :3894-3899). */
int lteb200_dci_1a_encode(const lteb200_alloc *alloc, int mcs, int N_rb_dl, uint8_t *e288);
int lteb200_bch_encode(const uint8_t *mib24, int N_ant, int N_id_cell, uint8_t *out1920);
int lteb200_dl_transmit_frames_batch(lteb200_ctx *ctx, const uint8_t *d_tb_bits, int tb_stride, const lteb200_alloc *allocs, int n_alloc, int fs_idx,
                                     int N_rb_dl, const lteb200_tx_control *ctl, const int *subfr_num, int n_sf, int N_pdcch_symbs, float *d_grid,
                                     float *d_iq, void *stream);

/* -------------------------------------------------------------------------------------------- front end */
/* LTE_fdd_dl_fs_samp_buf::freq_shift (LTE_fdd_dl_file_scan/src/LTE_fdd_dl_fs_samp_buf.cc:696-713) and the int8 -> float
 * conversion of copy_input_to_samp_buf (:657-694).  d_iq_in: DEVICE samples in iq_fmt; d_iq_out: DEVICE float2 [n_samps].
 * Sample i is multiplied by exp(-j*2*pi*freq_offset_hz*(index_base+i+1)/fs) exactly as the reference computes it. */
int lteb200_freq_shift_batch(lteb200_ctx *ctx, const void *d_iq_in, int iq_fmt, long long n_samps, long long index_base,
                             float freq_offset_hz, int fs_idx, float *d_iq_out, void *stream);
/* Host twin: in place on the split float buffers, samples [start_idx, start_idx + num_samps). */
int lteb200_freq_shift(lteb200_ctx *ctx, float *i_buf, float *q_buf, uint32_t start_idx, uint32_t num_samps, float freq_offset, int fs_idx);
/* Fused form: the OFDM demodulator applies the shift while loading samples (no extra pass over the data). */
int lteb200_set_freq_offset(lteb200_ctx *ctx, float freq_offset_hz, long long index_base);

/* Demapper of the max-log-MAP receive path (lteb200_pdsch_decode_batch / dl_receive_* / dl_scan_batch with LTEB200_TURBO_MAXLOGMAP, and
 * lteb200_pdsch_llr_q8_batch).  REFERENCE (default): the reference's own equaliser and demapper, bit for bit -- zero forcing, QPSK soft by distance,
 * hard +-127 decisions for 16/64QAM (liblte_phy.cc:7680-7690, :9502-9660).  MAXLOG: the north-star soft path -- noise variance from the CRS,
 * per-RE MMSE weighting (|h|^2 / sigma^2) and exact max-log LLRs per bit (int8: round(2 LLR), +-127); one antenna port or transmit diversity
 * with proper maximum-ratio normalisation.  Block sizes that are not a multiple of 32 and the reference-exact decoder keep the reference demapper. */
enum { LTEB200_DEMAP_REFERENCE = 0, LTEB200_DEMAP_MAXLOG = 1 };
/* Transport-block output format of lteb200_pdsch_decode_batch / dl_receive_batch / dl_receive_host / turbo_decode_q8_batch.  BYTES (default): one
 * bit per byte like the reference's out_bits (liblte_phy.h:906-913).  PACKED: eight bits per byte, first bit in the MSB -- an eighth of the
 * device-to-host traffic; tb_stride is then in bytes of packed data (>= tbs / 8).  PACKED needs transport blocks without filler bits (tbs + 24 a
 * turbo block size), with the max-log-MAP decoder also a block size that is a multiple of 32, and an allocation the symbol-level kernel takes
 * whole (no ragged multi-port tail, soft bits within its shared-memory staging); it is refused otherwise.  lteb200_dl_scan_batch and the
 * host twin lteb200_pdsch_channel_decode always return one bit per byte (their outputs are defined that way), whatever is set here. */
enum { LTEB200_TB_BYTES = 0, LTEB200_TB_PACKED = 1 };
int lteb200_set_tb_format(lteb200_ctx *ctx, int format);
int lteb200_set_demapper(lteb200_ctx *ctx, int demap);
/* Internal interleaver of the turbo code, for every entry point that encodes or decodes (turbo_decode*, pdsch_decode*, dl_receive*, dl_scan*,
 * dl_transmit*).  REFERENCE (default): pi(i) = ((f1 i + f2 i^2) mod 2^32) mod K -- the reference computes the quadratic in uint32
 * (liblte_phy.cc:10982), which for the larger block sizes is NOT a permutation and not what an LTE eNodeB transmits; encoder and decoder
 * of the reference agree with each other, and this mode reproduces both bit for bit.  SPEC: pi(i) = (f1 i + f2 i^2) mod K as 3GPP TS 36.212
 * 5.1.3.2.3 defines it -- the mode for signals that did not come from the reference's own transmitter.  No reference counterpart. */
enum { LTEB200_QPP_REFERENCE = 0, LTEB200_QPP_SPEC = 1 };
int lteb200_set_interleaver(lteb200_ctx *ctx, int qpp);

/* --------------------------------------------------------------------------------------------- utilities */
/* Pseudo-random sequence c(n), generate_prs_c (liblte_phy.cc:9669-9704); d_c DEVICE [len] bytes. */
int lteb200_prs_c_batch(lteb200_ctx *ctx, uint32_t c_init, int len, uint8_t *d_c, void *stream);
/* CRC24A check of n blocks of (tbs+24) bits each (calc_crc, liblte_phy.cc:9713-9743). */
int lteb200_crc24a_check_batch(lteb200_ctx *ctx, const uint8_t *d_bits, int stride, int tbs, int n, int32_t *d_ok, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* LTE_B200_H */

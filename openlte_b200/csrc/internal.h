// Internal launchers shared between translation units.
#pragma once
#include "common.cuh"

int turbo_ref_decode_batch(lteb200_ctx *ctx, const float *d_llr, int K, int n_cb, uint8_t *d_bits, cudaStream_t stream);
int turbo_ref_decode_q8_batch(lteb200_ctx *ctx, const int8_t *d_q8, int K, int n_cb, uint8_t *d_bits, cudaStream_t stream);   // images of pdsch_llr_q8_launch with max_no_tail
int turbo_mlm_decode_batch(lteb200_ctx *ctx, const float *d_llr, int K, int n_cb, int n_iter, uint8_t *d_bits, cudaStream_t stream);

// Per-(cell, allocation, CFI) tables of the PDSCH symbol level, shared by the receive and the transmit side (pdsch.cu)
struct AllocTables {
    int *re_idx;      // [3][idx_stride]
    int idx_stride;
    int M[3];
    uint8_t *cseq;    // [10][c_stride]
    uint8_t *cseqP;   // [10][idx_stride]: the Qm scrambling bits of a resource element in one byte (bit q = bit q of the element)
    int c_stride;
    int Qm, K, F;
};
int get_alloc_tables(lteb200_ctx *ctx, int N_rb_dl, int N_id_cell, int N_ant, int N_pdcch_symbs, const lteb200_alloc &al, AllocTables *out);
int get_twiddles(lteb200_ctx *ctx, int N, float2 **out);                 // e^{-2 pi i k/N}, k < N (ofdm.cu)
int get_crs_table(lteb200_ctx *ctx, int N_id_cell, float2 **out);        // [20][3][220] (chanest.cu)
int rate_unmatch_get_map(lteb200_ctx *ctx, int D, int E, int C, int tx_mode, int N_soft, int M_dl_harq, int chan_type, int rv, RateUnmatchMap *out);

// Size class of the max-log-MAP kernel for a code block of K bits: row stride (= threads per CTA) of its shared-memory arrays.  Shared memory
// scales with it (about 528 bytes per unit), so 64 / 128 / 192 leave room for 6 / 3 / 2 CTAs per SM and 384 for one.
static inline int mlm_size_class(int K)
{
    const int nw = (K + 15) / 16;
    return nw <= 64 ? 64 : (nw <= 128 ? 128 : (nw <= 192 ? 192 : 384));
}

// Pair-lane formulation of the max-log-MAP kernel (turbo_mlm2.cu): block sizes that are multiples of 32 (every K >= 1024)
struct FusedParams;
bool mlm2_handles(int K);
size_t mlm2_q8_stride(int K);                    // bytes of one quantised input image
int turbo_mlm2_launch(lteb200_ctx *ctx, const float *d_llr, const int8_t *q8, int K, int n_cb, int n_iter, uint8_t *d_bits, const FusedParams *fused,
                      cudaStream_t stream);
// PDSCH symbol level + rate unmatching + quantisation into that image, one CTA per transport block (pdsch_llr.cu)
int pdsch_llr_q8_launch(lteb200_ctx *ctx, const FusedParams &fp, int K, int E_max, int8_t *d_q8, cudaStream_t stream);
int pdsch_decode_batch_dev(lteb200_ctx *ctx, const float *d_grid, const float *d_ce, int N_rb_dl, int N_id_cell, int N_ant, const int *subfr_num,
                           const int *d_subfr_dev, int n_sf, int N_pdcch_symbs, const lteb200_alloc *allocs, int n_alloc, int turbo_mode, int n_iter,
                           uint8_t *d_tb_bits, int tb_stride, int32_t *d_crc_ok, cudaStream_t stream);
// control region, synchronisation signals and PBCH of the transmit side (control.cu)
int control_tx_launch(lteb200_ctx *ctx, float2 *grid, size_t sf_stride, size_t port_stride, const int *d_sn, const int *h_sn, int n_sf, int N_rb_dl,
                      int cfi, const lteb200_tx_control *ctl, const lteb200_alloc *allocs, int n_alloc, cudaStream_t stream);

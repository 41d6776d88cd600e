// C++ drop-in surface: the liblte_phy entry points that LTE_fdd_dl_file_scan calls on the downlink receive path,
// under the reference's own names, argument order and error behaviour, implemented over libopenlte_b200's C ABI
// (include/lte_b200.h).  A caller such as LTE_fdd_dl_fs_samp_buf.cc recompiles against this header instead of
// liblte/hdr/liblte_phy.h and links libopenlte_b200.so -- see INTEGRATION.md.
//
// Replaces (reference: /root/reference/liblte/hdr/liblte_phy.h unless noted):
//   liblte_phy_init :606-612 / liblte_phy_cleanup :638 / liblte_phy_update_n_rb_dl :652-653
//   liblte_phy_dl_find_coarse_timing_and_freq_offset :1134-1138     liblte_phy_find_pss_and_fine_timing :1068-1075
//   liblte_phy_find_sss :1106-1113                                  liblte_phy_get_dl_subframe_and_ce :1170-1177
//   liblte_phy_pdsch_channel_decode :906-913                        liblte_phy_rate_unmatch_turbo :1311-1323
//   liblte_phy_code_block_segmentation / _desegmentation :1337-1362
//   turbo_decode -- liblte/src/liblte_phy.cc:1530-1535 (file-local prototype, external linkage) and the new
//   public alias liblte_phy_turbo_decode BASELINE.json names.
//   liblte_phy_bch_channel_decode :947-953 / liblte_phy_pdcch_channel_decode :1012-1020 (N_ant == 1 for the PDCCH, see lte_b200.h)
//
// Types keep the reference's names and field names so callers compile unchanged; LIBLTE_PHY_STRUCT keeps only the
// public "Generic" fields callers read (phy_struct->N_samps_per_subfr, ->fs, LTE_fdd_dl_fs_samp_buf.cc:194,707) plus
// an opaque handle to the device context -- the reference's 46 MB of host scratch has no counterpart here.
//
// Use next to the reference's other headers.  Every real caller also includes liblte_rrc.h / liblte_common.h / typedefs.h
// (LTE_fdd_dl_fs_samp_buf.h:55-56).  When those headers are on the include path this file takes the fixed-width typedefs,
// LIBLTE_ERROR_ENUM, LIBLTE_BIT_MSG_STRUCT and the RRC enums FROM THEM and adds only what liblte_phy.h would have declared;
// include/dropin/liblte_phy.h is a one-line forwarder, so that `-Iinclude/dropin` in front of the reference's liblte/hdr makes
// the caller's own `#include "liblte_phy.h"` pick this surface (tests/test_fs_scan_gpu.py builds the reference's
// LTE_fdd_dl_fs_samp_buf.cc that way and runs its work() on the shipped capture).  Without them it is self-contained.
#ifndef LIBLTE_PHY_B200_H
#define LIBLTE_PHY_B200_H

#include <stdint.h>

#if defined(__has_include)
#if __has_include("liblte_common.h") && __has_include("liblte_rrc.h") && !defined(LIBLTE_PHY_B200_STANDALONE)
#define LIBLTE_PHY_B200_REFERENCE_TYPES 1
#endif
#endif

#ifdef LIBLTE_PHY_B200_REFERENCE_TYPES
#include "liblte_common.h"   // typedefs.h (int8 is plain char there: callers must keep -fsigned-char semantics), LIBLTE_ERROR_ENUM, LIBLTE_BIT_MSG_STRUCT
#include "liblte_rrc.h"      // LIBLTE_RRC_PHICH_DURATION_ENUM, liblte_rrc_phich_resource_num, ...
#else
typedef signed char        int8;      // the reference's plain `char` is signed on x86 only: fixed here
typedef unsigned char      uint8;
typedef short              int16;
typedef unsigned short     uint16;
typedef int                int32;
typedef unsigned int       uint32;
typedef long long          int64;
typedef unsigned long long uint64;

#define LIBLTE_MAX_MSG_SIZE 5512
typedef enum {
    LIBLTE_SUCCESS = 0,
    LIBLTE_ERROR_INVALID_INPUTS,
    LIBLTE_ERROR_DECODE_FAIL,
    LIBLTE_ERROR_INVALID_CRC,
    LIBLTE_ERROR_INVALID_CONTENTS,
} LIBLTE_ERROR_ENUM;
typedef struct { uint32 N_bits; uint8 msg[LIBLTE_MAX_MSG_SIZE]; } LIBLTE_BIT_MSG_STRUCT;
typedef enum { LIBLTE_RRC_PHICH_DURATION_NORMAL = 0, LIBLTE_RRC_PHICH_DURATION_EXTENDED } LIBLTE_RRC_PHICH_DURATION_ENUM;
#endif

// ---- what liblte_phy.h declares on this path (liblte_phy.h:100-224)
#define LIBLTE_PHY_N_RB_DL_1_4MHZ 6
#define LIBLTE_PHY_N_RB_DL_3MHZ   15
#define LIBLTE_PHY_N_RB_DL_5MHZ   25
#define LIBLTE_PHY_N_RB_DL_10MHZ  50
#define LIBLTE_PHY_N_RB_DL_15MHZ  75
#define LIBLTE_PHY_N_RB_DL_20MHZ  100
#define LIBLTE_PHY_N_RB_DL_MAX    110
#define LIBLTE_PHY_N_SC_RB_DL_NORMAL_CP 12
#define LIBLTE_PHY_N_ANT_MAX 4
#define LIBLTE_PHY_N_SLOTS_PER_SUBFR 2
#define LIBLTE_PHY_MAX_CODE_BLOCK_SIZE 6176
#define LIBLTE_PHY_INIT_N_ID_CELL_UNKNOWN 0xFFFF
#define LIBLTE_PHY_N_MAX_ROUGH_CORR_SEARCH_PEAKS 5
// samples per frame / subframe / slot at each sample rate (liblte_phy.h:141-175): callers size their buffers with these
#define LIBLTE_PHY_N_SAMPS_PER_SLOT_30_72MHZ  15360
#define LIBLTE_PHY_N_SAMPS_PER_SUBFR_30_72MHZ (LIBLTE_PHY_N_SAMPS_PER_SLOT_30_72MHZ*2)
#define LIBLTE_PHY_N_SAMPS_PER_FRAME_30_72MHZ (LIBLTE_PHY_N_SAMPS_PER_SUBFR_30_72MHZ*10)
#define LIBLTE_PHY_N_SAMPS_PER_SLOT_15_36MHZ  7680
#define LIBLTE_PHY_N_SAMPS_PER_SUBFR_15_36MHZ (LIBLTE_PHY_N_SAMPS_PER_SLOT_15_36MHZ*2)
#define LIBLTE_PHY_N_SAMPS_PER_FRAME_15_36MHZ (LIBLTE_PHY_N_SAMPS_PER_SUBFR_15_36MHZ*10)
#define LIBLTE_PHY_N_SAMPS_PER_SLOT_7_68MHZ   3840
#define LIBLTE_PHY_N_SAMPS_PER_SUBFR_7_68MHZ  (LIBLTE_PHY_N_SAMPS_PER_SLOT_7_68MHZ*2)
#define LIBLTE_PHY_N_SAMPS_PER_FRAME_7_68MHZ  (LIBLTE_PHY_N_SAMPS_PER_SUBFR_7_68MHZ*10)
#define LIBLTE_PHY_N_SAMPS_PER_SLOT_3_84MHZ   1920
#define LIBLTE_PHY_N_SAMPS_PER_SUBFR_3_84MHZ  (LIBLTE_PHY_N_SAMPS_PER_SLOT_3_84MHZ*2)
#define LIBLTE_PHY_N_SAMPS_PER_FRAME_3_84MHZ  (LIBLTE_PHY_N_SAMPS_PER_SUBFR_3_84MHZ*10)
#define LIBLTE_PHY_N_SAMPS_PER_SLOT_1_92MHZ   960
#define LIBLTE_PHY_N_SAMPS_PER_SUBFR_1_92MHZ  (LIBLTE_PHY_N_SAMPS_PER_SLOT_1_92MHZ*2)
#define LIBLTE_PHY_N_SAMPS_PER_FRAME_1_92MHZ  (LIBLTE_PHY_N_SAMPS_PER_SUBFR_1_92MHZ*10)

typedef enum { LIBLTE_PHY_FS_1_92MHZ = 0, LIBLTE_PHY_FS_3_84MHZ, LIBLTE_PHY_FS_7_68MHZ, LIBLTE_PHY_FS_15_36MHZ, LIBLTE_PHY_FS_30_72MHZ, LIBLTE_PHY_FS_N_ITEMS } LIBLTE_PHY_FS_ENUM;
static const char liblte_phy_fs_text[LIBLTE_PHY_FS_N_ITEMS][20] = {"1.92", "3.84", "7.68", "15.36", "30.72"};   // liblte_phy.h:206-210
typedef enum { LIBLTE_PHY_PRE_CODER_TYPE_TX_DIVERSITY = 0, LIBLTE_PHY_PRE_CODER_TYPE_SPATIAL_MULTIPLEXING } LIBLTE_PHY_PRE_CODER_TYPE_ENUM;
typedef enum { LIBLTE_PHY_MODULATION_TYPE_BPSK = 0, LIBLTE_PHY_MODULATION_TYPE_QPSK, LIBLTE_PHY_MODULATION_TYPE_16QAM, LIBLTE_PHY_MODULATION_TYPE_64QAM } LIBLTE_PHY_MODULATION_TYPE_ENUM;
typedef enum { LIBLTE_PHY_CHAN_TYPE_DLSCH = 0, LIBLTE_PHY_CHAN_TYPE_PCH, LIBLTE_PHY_CHAN_TYPE_ULSCH, LIBLTE_PHY_CHAN_TYPE_ULCCH } LIBLTE_PHY_CHAN_TYPE_ENUM;

typedef struct {
    float rx_symb_re[16][LIBLTE_PHY_N_RB_DL_20MHZ * LIBLTE_PHY_N_SC_RB_DL_NORMAL_CP];
    float rx_symb_im[16][LIBLTE_PHY_N_RB_DL_20MHZ * LIBLTE_PHY_N_SC_RB_DL_NORMAL_CP];
    float rx_ce_re[LIBLTE_PHY_N_ANT_MAX][16][LIBLTE_PHY_N_RB_DL_20MHZ * LIBLTE_PHY_N_SC_RB_DL_NORMAL_CP];
    float rx_ce_im[LIBLTE_PHY_N_ANT_MAX][16][LIBLTE_PHY_N_RB_DL_20MHZ * LIBLTE_PHY_N_SC_RB_DL_NORMAL_CP];
    float tx_symb_re[LIBLTE_PHY_N_ANT_MAX][16][LIBLTE_PHY_N_RB_DL_20MHZ * LIBLTE_PHY_N_SC_RB_DL_NORMAL_CP];
    float tx_symb_im[LIBLTE_PHY_N_ANT_MAX][16][LIBLTE_PHY_N_RB_DL_20MHZ * LIBLTE_PHY_N_SC_RB_DL_NORMAL_CP];
    uint32 num;
} LIBLTE_PHY_SUBFRAME_STRUCT;

typedef struct {
    LIBLTE_BIT_MSG_STRUCT           msg[2];
    LIBLTE_PHY_PRE_CODER_TYPE_ENUM  pre_coder_type;
    LIBLTE_PHY_MODULATION_TYPE_ENUM mod_type;
    LIBLTE_PHY_CHAN_TYPE_ENUM       chan_type;
    uint32                          tbs;
    uint32                          rv_idx;
    uint32                          N_prb;
    uint32                          prb[LIBLTE_PHY_N_SLOTS_PER_SUBFR][LIBLTE_PHY_N_RB_DL_MAX];
    uint32                          N_codewords;
    uint32                          N_layers;
    uint32                          tx_mode;
    uint32                          harq_retx_count;
    uint16                          rnti;
    uint8                           mcs;
    uint8                           tpc;
    bool                            ndi;
    bool                            dl_alloc;
} LIBLTE_PHY_ALLOCATION_STRUCT;

typedef struct {
    float  freq_offset[LIBLTE_PHY_N_MAX_ROUGH_CORR_SEARCH_PEAKS];
    uint32 symb_starts[LIBLTE_PHY_N_MAX_ROUGH_CORR_SEARCH_PEAKS][7];
    uint32 n_corr_peaks;
} LIBLTE_PHY_COARSE_TIMING_STRUCT;

typedef struct {
    // "Generic" block of the reference struct (liblte_phy.h:586-603), same names
    uint32 fs;
    uint32 N_samps_per_symb;
    uint32 N_samps_cp_l_0;
    uint32 N_samps_cp_l_else;
    uint32 N_samps_per_slot;
    uint32 N_samps_per_subfr;
    uint32 N_samps_per_frame;
    uint32 N_rb_dl;
    uint32 N_rb_ul;
    uint32 N_sc_rb_dl;
    uint32 N_sc_rb_ul;
    uint32 FFT_pad_size;
    uint32 FFT_size;
    uint8  N_ant;
    bool   ul_init;
    // B200 additions
    void  *b200_ctx;        // lteb200_ctx*
    int    b200_fs_idx;
    int    b200_turbo_mode; // LTEB200_TURBO_REF_EXACT (default: bit-exact with the reference) or LTEB200_TURBO_MAXLOGMAP
    int    b200_turbo_iters;
} LIBLTE_PHY_STRUCT;

LIBLTE_ERROR_ENUM liblte_phy_init(LIBLTE_PHY_STRUCT **phy_struct, LIBLTE_PHY_FS_ENUM fs, uint16 N_id_cell, uint8 N_ant, uint32 N_rb_dl,
                                  uint32 N_sc_rb_dl, float phich_res);
LIBLTE_ERROR_ENUM liblte_phy_cleanup(LIBLTE_PHY_STRUCT *phy_struct);
LIBLTE_ERROR_ENUM liblte_phy_update_n_rb_dl(LIBLTE_PHY_STRUCT *phy_struct, uint32 N_rb_dl);

LIBLTE_ERROR_ENUM liblte_phy_dl_find_coarse_timing_and_freq_offset(LIBLTE_PHY_STRUCT *phy_struct, float *i_samps, float *q_samps, uint32 N_slots,
                                                                   LIBLTE_PHY_COARSE_TIMING_STRUCT *timing_struct);
LIBLTE_ERROR_ENUM liblte_phy_find_pss_and_fine_timing(LIBLTE_PHY_STRUCT *phy_struct, float *i_samps, float *q_samps, uint32 *symb_starts,
                                                      uint32 *N_id_2, uint32 *pss_symb, float *pss_thresh, float *freq_offset);
LIBLTE_ERROR_ENUM liblte_phy_find_sss(LIBLTE_PHY_STRUCT *phy_struct, float *i_samps, float *q_samps, uint32 N_id_2, uint32 *symb_starts,
                                      float pss_thresh, uint32 *N_id_1, uint32 *frame_start_idx);
LIBLTE_ERROR_ENUM liblte_phy_get_dl_subframe_and_ce(LIBLTE_PHY_STRUCT *phy_struct, float *i_samps, float *q_samps, uint32 frame_start_idx,
                                                    uint8 subfr_num, uint32 N_id_cell, uint8 N_ant, LIBLTE_PHY_SUBFRAME_STRUCT *subframe);
LIBLTE_ERROR_ENUM liblte_phy_pdsch_channel_decode(LIBLTE_PHY_STRUCT *phy_struct, LIBLTE_PHY_SUBFRAME_STRUCT *subframe,
                                                  LIBLTE_PHY_ALLOCATION_STRUCT *alloc, uint32 N_pdcch_symbs, uint32 N_id_cell, uint8 N_ant,
                                                  uint8 *out_bits, uint32 *N_out_bits);
// ---- control channels (liblte_phy.h:941-947, 967-1013)
#define LIBLTE_PHY_PDCCH_MAX_ALLOC 6
typedef struct { LIBLTE_PHY_ALLOCATION_STRUCT alloc[LIBLTE_PHY_PDCCH_MAX_ALLOC]; uint32 N_symbs; uint32 N_alloc; } LIBLTE_PHY_PDCCH_STRUCT;
typedef struct { float n[4]; uint32 k[4]; uint32 cfi; uint32 N_reg; } LIBLTE_PHY_PCFICH_STRUCT;
typedef struct { float z_re[3]; float z_im[3]; uint32 k[75]; uint32 N_reg; uint8 b[25][8]; bool present[25][8]; } LIBLTE_PHY_PHICH_STRUCT;
LIBLTE_ERROR_ENUM liblte_phy_bch_channel_decode(LIBLTE_PHY_STRUCT *phy_struct, LIBLTE_PHY_SUBFRAME_STRUCT *subframe, uint32 N_id_cell, uint8 *N_ant,
                                                uint8 *out_bits, uint32 *N_out_bits, uint8 *offset);
// N_ant == 1 only (see lte_b200.h); pcfich->cfi, pcfich->N_reg, phich->N_reg, pdcch->N_symbs and the allocation list are filled in
LIBLTE_ERROR_ENUM liblte_phy_pdcch_channel_decode(LIBLTE_PHY_STRUCT *phy_struct, LIBLTE_PHY_SUBFRAME_STRUCT *subframe, uint32 N_id_cell, uint8 N_ant,
                                                  float phich_res, LIBLTE_RRC_PHICH_DURATION_ENUM phich_dur, LIBLTE_PHY_PCFICH_STRUCT *pcfich,
                                                  LIBLTE_PHY_PHICH_STRUCT *phich, LIBLTE_PHY_PDCCH_STRUCT *pdcch);

void liblte_phy_rate_unmatch_turbo(LIBLTE_PHY_STRUCT *phy_struct, float *e_bits, uint32 N_e_bits, uint8 *dummy_bits, uint32 N_dummy_bits,
                                   uint32 N_codeblocks, uint32 tx_mode, uint32 N_soft, uint32 M_dl_harq, LIBLTE_PHY_CHAN_TYPE_ENUM chan_type,
                                   uint32 rv_idx, float *d_bits, uint32 *N_d_bits);
void liblte_phy_code_block_segmentation(uint8 *b_bits, uint32 N_b_bits, uint32 *N_codeblocks, uint32 *N_filler_bits, uint8 *c_bits,
                                        uint32 N_c_bits_max, uint32 *N_c_bits);
void liblte_phy_code_block_desegmentation(uint8 *c_bits, uint32 *N_c_bits, uint32 N_c_bits_max, uint32 tbs, uint8 *b_bits, uint32 N_b_bits);
// same mangled name and argument list as the reference's externally linked turbo_decode (liblte_phy.cc:1530-1535)
void turbo_decode(LIBLTE_PHY_STRUCT *phy_struct, float *d_bits, uint32 N_d_bits, uint32 N_fill_bits, uint8 *c_bits, uint32 *N_c_bits);
void liblte_phy_turbo_decode(LIBLTE_PHY_STRUCT *phy_struct, float *d_bits, uint32 N_d_bits, uint32 N_fill_bits, uint8 *c_bits, uint32 *N_c_bits);

#endif

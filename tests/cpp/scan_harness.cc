// GNU-Radio-free restatement of LTE_fdd_dl_file_scan's acquisition loop (LTE_fdd_dl_fs_samp_buf.cc:243-515), written
// against the drop-in header only: coarse timing -> PSS -> SSS -> subframe + channel estimate -> PDSCH decode of a given
// allocation.  (PBCH/PDCCH decode stay with the reference library -- SURVEY 8f -- so the allocation comes from argv.)
// usage: scan_harness <int8 iq file> <n_samples> <N_slots> <sib1_subframe> <rnti> <tbs> <n_prb> <mod> <rv> <cfi>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "liblte_phy_b200.h"

int main(int argc, char **argv)
{
    if (argc < 11) { fprintf(stderr, "usage\n"); return 2; }
    const long n = atol(argv[2]);
    const unsigned N_slots = atoi(argv[3]);
    std::vector<signed char> raw(2 * n);
    FILE *f = fopen(argv[1], "rb");
    if (!f || fread(raw.data(), 1, raw.size(), f) != raw.size()) { fprintf(stderr, "read failed\n"); return 2; }
    fclose(f);
    std::vector<float> i_buf(n), q_buf(n);
    for (long k = 0; k < n; k++) { i_buf[k] = (float)raw[2 * k]; q_buf[k] = (float)raw[2 * k + 1]; }   // copy_input_to_samp_buf, fs.cc:657-694

    LIBLTE_PHY_STRUCT *phy = NULL;
    if (liblte_phy_init(&phy, LIBLTE_PHY_FS_30_72MHZ, LIBLTE_PHY_INIT_N_ID_CELL_UNKNOWN, 4, LIBLTE_PHY_N_RB_DL_1_4MHZ,
                        LIBLTE_PHY_N_SC_RB_DL_NORMAL_CP, 1.0f) != LIBLTE_SUCCESS) { printf("INIT_FAILED\n"); return 1; }
    LIBLTE_PHY_COARSE_TIMING_STRUCT timing;
    if (liblte_phy_dl_find_coarse_timing_and_freq_offset(phy, i_buf.data(), q_buf.data(), N_slots, &timing) != LIBLTE_SUCCESS) { printf("COARSE_FAILED\n"); return 1; }
    printf("coarse n_corr_peaks=%u symb_start0=%u freq_offset0=%.2f\n", timing.n_corr_peaks, timing.symb_starts[0][0], timing.freq_offset[0]);
    uint32 N_id_2 = 0, pss_symb = 0, N_id_1 = 0, frame_start_idx = 0;
    float pss_thresh = 0, freq_offset = 0;
    if (liblte_phy_find_pss_and_fine_timing(phy, i_buf.data(), q_buf.data(), timing.symb_starts[0], &N_id_2, &pss_symb, &pss_thresh, &freq_offset) != LIBLTE_SUCCESS) { printf("PSS_FAILED\n"); return 1; }
    printf("pss N_id_2=%u pss_symb=%u freq_offset=%.0f\n", N_id_2, pss_symb, freq_offset);
    if (liblte_phy_find_sss(phy, i_buf.data(), q_buf.data(), N_id_2, timing.symb_starts[0], pss_thresh, &N_id_1, &frame_start_idx) != LIBLTE_SUCCESS) { printf("SSS_FAILED\n"); return 1; }
    const uint32 N_id_cell = 3 * N_id_1 + N_id_2;
    printf("sss N_id_1=%u frame_start_idx=%u N_id_cell=%u\n", N_id_1, frame_start_idx, N_id_cell);
    // MIB: subframe 0 of the detected frame, four-port estimate, bandwidth still unknown (fs.cc:378-386)
    static LIBLTE_PHY_SUBFRAME_STRUCT sf0;
    if ((long long)frame_start_idx + phy->N_samps_per_subfr + 2 * phy->N_samps_per_symb + 2 * phy->N_samps_cp_l_0 < (long long)n &&
        liblte_phy_get_dl_subframe_and_ce(phy, i_buf.data(), q_buf.data(), frame_start_idx, 0, N_id_cell, 4, &sf0) == LIBLTE_SUCCESS) {
        uint8 mib_ant = 0, mib_off = 0, mib[64];
        uint32 n_mib = 0;
        LIBLTE_ERROR_ENUM be = liblte_phy_bch_channel_decode(phy, &sf0, N_id_cell, &mib_ant, mib, &n_mib, &mib_off);
        printf("bch err=%d N_ant=%u offset=%u bits=", (int)be, mib_ant, mib_off);
        for (uint32 k = 0; k < n_mib; k++) putchar('0' + mib[k]);
        putchar('\n');
    }
    liblte_phy_update_n_rb_dl(phy, LIBLTE_PHY_N_RB_DL_20MHZ);                                   // what the MIB says (fs.cc:417)
    // the frame that starts one frame earlier (uint32 wrap, as the reference computes it) holds an even SFN in this capture
    const uint32 fs0 = frame_start_idx - phy->N_samps_per_frame;
    static LIBLTE_PHY_SUBFRAME_STRUCT subframe;
    if (liblte_phy_get_dl_subframe_and_ce(phy, i_buf.data(), q_buf.data(), fs0, (uint8)atoi(argv[4]), N_id_cell, 1, &subframe) != LIBLTE_SUCCESS) { printf("CE_FAILED\n"); return 1; }
    // PCFICH + PDCCH: the scan takes the allocation from here (LTE_fdd_dl_fs_samp_buf.cc:453, 499)
    static LIBLTE_PHY_PCFICH_STRUCT pcfich; static LIBLTE_PHY_PHICH_STRUCT phich; static LIBLTE_PHY_PDCCH_STRUCT pdcch;
    LIBLTE_ERROR_ENUM pe = liblte_phy_pdcch_channel_decode(phy, &subframe, N_id_cell, 1, 1.0f, LIBLTE_RRC_PHICH_DURATION_NORMAL, &pcfich, &phich, &pdcch);
    printf("pdcch err=%d cfi=%u N_symbs=%u N_alloc=%u", (int)pe, pcfich.cfi, pdcch.N_symbs, pdcch.N_alloc);
    if (pdcch.N_alloc > 0) printf(" rnti=0x%04X tbs=%u N_prb=%u prb0=%u mod=%d rv=%u", pdcch.alloc[0].rnti, pdcch.alloc[0].tbs, pdcch.alloc[0].N_prb,
                                  pdcch.alloc[0].prb[0][0], (int)pdcch.alloc[0].mod_type, pdcch.alloc[0].rv_idx);
    putchar('\n');
    if (pe == LIBLTE_SUCCESS && pdcch.N_alloc > 0) {
        static uint8 sib_bits[LIBLTE_MAX_MSG_SIZE];
        uint32 n_sib = 0;
        LIBLTE_ERROR_ENUM se = liblte_phy_pdsch_channel_decode(phy, &subframe, &pdcch.alloc[0], pdcch.N_symbs, N_id_cell, 1, sib_bits, &n_sib);
        printf("pdsch_from_dci err=%d N_bits=%u bits=", (int)se, n_sib);
        for (uint32 k = 0; k < n_sib; k++) putchar('0' + sib_bits[k]);
        putchar('\n');
    }
    static LIBLTE_PHY_ALLOCATION_STRUCT alloc;
    memset(&alloc, 0, sizeof(alloc));
    alloc.rnti = (uint16)strtol(argv[5], NULL, 0); alloc.tbs = atoi(argv[6]); alloc.N_prb = atoi(argv[7]);
    alloc.mod_type = (LIBLTE_PHY_MODULATION_TYPE_ENUM)atoi(argv[8]); alloc.rv_idx = atoi(argv[9]);
    alloc.pre_coder_type = LIBLTE_PHY_PRE_CODER_TYPE_TX_DIVERSITY; alloc.chan_type = LIBLTE_PHY_CHAN_TYPE_DLSCH;
    alloc.N_codewords = 1; alloc.N_layers = 1; alloc.tx_mode = 1;
    for (uint32 j = 0; j < alloc.N_prb; j++) { alloc.prb[0][j] = j; alloc.prb[1][j] = j; }
    static uint8 bits[LIBLTE_MAX_MSG_SIZE];
    uint32 N_bits = 0;
    LIBLTE_ERROR_ENUM e = liblte_phy_pdsch_channel_decode(phy, &subframe, &alloc, atoi(argv[10]), N_id_cell, 1, bits, &N_bits);
    printf("pdsch err=%d N_bits=%u bits=", (int)e, N_bits);
    for (uint32 k = 0; k < N_bits; k++) putchar('0' + bits[k]);
    putchar('\n');
    // error path: NULL output -> INVALID_INPUTS, as the reference's entry check (liblte_phy.cc:3699-3705)
    printf("null_check=%d\n", (int)liblte_phy_pdsch_channel_decode(phy, &subframe, &alloc, 2, N_id_cell, 1, NULL, &N_bits));
    liblte_phy_cleanup(phy);
    return 0;
}

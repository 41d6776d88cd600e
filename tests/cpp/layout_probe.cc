// Prints enum values, struct sizes and field offsets of the drop-in surface; built once against the reference's own liblte_phy.h (-DUSE_REF) and once
// against include/liblte_phy_b200.h in its self-contained mode by tests/test_library_abi.py, which compares the two outputs.
#include <stdio.h>
#include <stddef.h>
#ifdef USE_REF
#include "liblte_phy.h"
#else
#include "liblte_phy_b200.h"
#endif
int main(){
 printf("ERR %d %d %d %d %d %d\n", (int)LIBLTE_SUCCESS,(int)LIBLTE_ERROR_INVALID_INPUTS,(int)LIBLTE_ERROR_DECODE_FAIL,(int)LIBLTE_ERROR_INVALID_CRC,(int)LIBLTE_ERROR_INVALID_CONTENTS,0);
 printf("FS %d %d %d %d %d %d\n",(int)LIBLTE_PHY_FS_1_92MHZ,(int)LIBLTE_PHY_FS_3_84MHZ,(int)LIBLTE_PHY_FS_7_68MHZ,(int)LIBLTE_PHY_FS_15_36MHZ,(int)LIBLTE_PHY_FS_30_72MHZ,(int)LIBLTE_PHY_FS_N_ITEMS);
 printf("MOD %d %d %d %d\n",(int)LIBLTE_PHY_MODULATION_TYPE_BPSK,(int)LIBLTE_PHY_MODULATION_TYPE_QPSK,(int)LIBLTE_PHY_MODULATION_TYPE_16QAM,(int)LIBLTE_PHY_MODULATION_TYPE_64QAM);
 printf("CHAN %d %d %d %d\n",(int)LIBLTE_PHY_CHAN_TYPE_DLSCH,(int)LIBLTE_PHY_CHAN_TYPE_PCH,(int)LIBLTE_PHY_CHAN_TYPE_ULSCH,(int)LIBLTE_PHY_CHAN_TYPE_ULCCH);
 printf("PRE %d %d\n",(int)LIBLTE_PHY_PRE_CODER_TYPE_TX_DIVERSITY,(int)LIBLTE_PHY_PRE_CODER_TYPE_SPATIAL_MULTIPLEXING);
 printf("SUBFRAME %zu rx_symb_re %zu rx_ce_re %zu num %zu\n", sizeof(LIBLTE_PHY_SUBFRAME_STRUCT), offsetof(LIBLTE_PHY_SUBFRAME_STRUCT, rx_symb_re), offsetof(LIBLTE_PHY_SUBFRAME_STRUCT, rx_ce_re), offsetof(LIBLTE_PHY_SUBFRAME_STRUCT, num));
 printf("ALLOC %zu prb %zu tbs %zu rnti %zu mcs %zu mod %zu rv %zu N_prb %zu tx_mode %zu\n", sizeof(LIBLTE_PHY_ALLOCATION_STRUCT), offsetof(LIBLTE_PHY_ALLOCATION_STRUCT, prb), offsetof(LIBLTE_PHY_ALLOCATION_STRUCT, tbs), offsetof(LIBLTE_PHY_ALLOCATION_STRUCT, rnti), offsetof(LIBLTE_PHY_ALLOCATION_STRUCT, mcs), offsetof(LIBLTE_PHY_ALLOCATION_STRUCT, mod_type), offsetof(LIBLTE_PHY_ALLOCATION_STRUCT, rv_idx), offsetof(LIBLTE_PHY_ALLOCATION_STRUCT, N_prb), offsetof(LIBLTE_PHY_ALLOCATION_STRUCT, tx_mode));
 printf("COARSE %zu PDCCH %zu PCFICH %zu PHICH %zu\n", sizeof(LIBLTE_PHY_COARSE_TIMING_STRUCT), sizeof(LIBLTE_PHY_PDCCH_STRUCT), sizeof(LIBLTE_PHY_PCFICH_STRUCT), sizeof(LIBLTE_PHY_PHICH_STRUCT));
 printf("BITMSG %zu\n", sizeof(LIBLTE_BIT_MSG_STRUCT));
}

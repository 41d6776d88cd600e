// extern "C" handles on the drop-in's C++ entry points that are pure host code (include/liblte_phy_b200.h: code block (de)segmentation,
// liblte_phy_update_n_rb_dl), so that a CPU test can call them through ctypes; built by tests/test_host_tables.py with g++ against
// libopenlte_b200.so.
#include <stdlib.h>
#include "liblte_phy_b200.h"
extern "C" void shim_update_n_rb_dl_probe(unsigned N_samps_per_symb, unsigned N_rb_dl, unsigned *out)
{
    LIBLTE_PHY_STRUCT *p = (LIBLTE_PHY_STRUCT *)calloc(1, sizeof(LIBLTE_PHY_STRUCT));
    p->N_samps_per_symb = N_samps_per_symb; p->N_sc_rb_dl = 12;
    out[0] = (unsigned)liblte_phy_update_n_rb_dl(p, N_rb_dl);
    out[1] = p->N_rb_dl; out[2] = p->FFT_size; out[3] = p->FFT_pad_size;
    free(p);
}
extern "C" void shim_code_block_segmentation(uint8 *b_bits, uint32 N_b_bits, uint32 *N_codeblocks, uint32 *N_filler_bits, uint8 *c_bits, uint32 N_c_bits_max,
                                             uint32 *N_c_bits)
{
    liblte_phy_code_block_segmentation(b_bits, N_b_bits, N_codeblocks, N_filler_bits, c_bits, N_c_bits_max, N_c_bits);
}
extern "C" void shim_code_block_desegmentation(uint8 *c_bits, uint32 *N_c_bits, uint32 N_c_bits_max, uint32 tbs, uint8 *b_bits, uint32 N_b_bits)
{
    liblte_phy_code_block_desegmentation(c_bits, N_c_bits, N_c_bits_max, tbs, b_bits, N_b_bits);
}

// extern "C" handles on the drop-in's C++ entry points for code block (de)segmentation (include/liblte_phy_b200.h), so that a CPU test can
// call them through ctypes; built by tests/test_host_tables.py with g++ against libopenlte_b200.so.  These two functions are host code.
#include "liblte_phy_b200.h"
extern "C" void shim_code_block_segmentation(uint8 *b_bits, uint32 N_b_bits, uint32 *N_codeblocks, uint32 *N_filler_bits, uint8 *c_bits, uint32 N_c_bits_max,
                                             uint32 *N_c_bits)
{
    liblte_phy_code_block_segmentation(b_bits, N_b_bits, N_codeblocks, N_filler_bits, c_bits, N_c_bits_max, N_c_bits);
}
extern "C" void shim_code_block_desegmentation(uint8 *c_bits, uint32 *N_c_bits, uint32 N_c_bits_max, uint32 tbs, uint8 *b_bits, uint32 N_b_bits)
{
    liblte_phy_code_block_desegmentation(c_bits, N_c_bits, N_c_bits_max, tbs, b_bits, N_b_bits);
}

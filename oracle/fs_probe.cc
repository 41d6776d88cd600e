/* TEST INFRASTRUCTURE ONLY.  extern "C" handles on two private helpers of the reference's OWN file-scan block
 * (LTE_fdd_dl_file_scan/src/LTE_fdd_dl_fs_samp_buf.cc, compiled where it lies, unmodified, by oracle/Makefile into
 * oracle/_ref/libfs_probe.so together with the reference PHY):
 *   freq_shift              fs.cc:696-713   -- pins oracle/port op_freq_shift (and through it lteb200_freq_shift) bit for bit
 *   copy_input_to_samp_buf  fs.cc:657-694   -- the int8 capture format (I, Q, I, Q ... also across odd-sized chunks)
 * Only the access specifier is lifted (the class is declared with its private section readable in THIS translation unit; the block's
 * own object code is compiled from the untouched source). */
#include <stdlib.h>
/* everything the block's header pulls in is included first (guarded headers), so that lifting the access specifier touches that class only */
#include "LTE_fdd_dl_fs_api.h"
#include "liblte_phy.h"
#include "liblte_rrc.h"
#include <gnuradio/sync_block.h>
#define private public
#include "LTE_fdd_dl_fs_samp_buf.h"
#undef private

static LTE_fdd_dl_fs_samp_buf_sptr g_blk;      /* never destroyed: the destructor would clean up a PHY struct the block only creates in work() */
static LTE_fdd_dl_fs_samp_buf *blk()
{
    if (!g_blk) g_blk = LTE_fdd_dl_fs_make_samp_buf(sizeof(int8));
    return g_blk.get();
}

extern "C" {
/* in place on the caller's buffers, sample numbering from the buffer start as in the block */
void fs_probe_freq_shift(float *i_buf, float *q_buf, unsigned start_idx, unsigned num_samps, float freq_offset, unsigned fs_hz)
{
    LTE_fdd_dl_fs_samp_buf *b = blk();
    LIBLTE_PHY_STRUCT *fake = (LIBLTE_PHY_STRUCT *)calloc(1, sizeof(LIBLTE_PHY_STRUCT));
    fake->fs = fs_hz;
    float *si = b->i_buf, *sq = b->q_buf;
    LIBLTE_PHY_STRUCT *sp = b->phy_struct;
    b->i_buf = i_buf; b->q_buf = q_buf; b->phy_struct = fake;
    b->freq_shift(start_idx, num_samps, freq_offset);
    b->i_buf = si; b->q_buf = sq; b->phy_struct = sp;
    free(fake);
}
/* feeds the chunks one after the other (sizes in items = bytes) and returns how many complex samples the block holds; out_i/out_q receive them */
int fs_probe_copy_input_int8(const signed char *items, const int *chunk_sizes, int n_chunks, float *out_i, float *out_q, int capacity)
{
    LTE_fdd_dl_fs_samp_buf *b = blk();
    b->samp_buf_w_idx = 0; b->samp_buf_r_idx = 0; b->last_samp_was_i = false;
    gr_vector_const_void_star in(1);
    const signed char *p = items;
    for (int c = 0; c < n_chunks; c++) {
        in[0] = p;
        b->copy_input_to_samp_buf(in, chunk_sizes[c]);
        p += chunk_sizes[c];
    }
    const int n = (int)b->samp_buf_w_idx;
    for (int k = 0; k < n && k < capacity; k++) { out_i[k] = b->i_buf[k]; out_q[k] = b->q_buf[k]; }
    b->samp_buf_w_idx = 0; b->last_samp_was_i = false;
    return n;
}
}

/* TEST INFRASTRUCTURE ONLY.  Drives the reference's OWN LTE_fdd_dl_fs_samp_buf block (LTE_fdd_dl_file_scan/src/LTE_fdd_dl_fs_samp_buf.cc,
 * compiled where it lies) the way its flow graph does: a file source feeding int8 items into work() chunk by chunk
 * (LTE_fdd_dl_file_scan/python/LTE_fdd_dl_file_scan.py:35-38).  Built twice by oracle/Makefile:
 *   _ref/fs_scan_ref   against the reference's liblte_phy.cc (CPU)
 *   _ref/fs_scan_b200  against include/dropin/liblte_phy.h + libopenlte_b200.so (the drop-in surface, GPU)
 * and the test compares what the two print (MIB, SIB1, ... of the shipped capture).  The block asks for its configuration on stdin:
 * an empty line starts the scan (fs = 30.72 MHz, the default).
 * usage: fs_scan_xxx <capture.bin> [chunk_bytes] */
#include "LTE_fdd_dl_fs_samp_buf.h"
#include <stdio.h>
#include <stdlib.h>
#include <vector>

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s <int8 capture> [chunk_bytes]\n", argv[0]); return 2; }
    const int chunk = argc > 2 ? atoi(argv[2]) : 32768;
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    LTE_fdd_dl_fs_samp_buf_sptr blk = LTE_fdd_dl_fs_make_samp_buf(sizeof(int8));
    std::vector<signed char> buf(chunk);
    gr_vector_const_void_star in(1);
    gr_vector_void_star out;
    long total = 0;
    for (;;) {
        const int n = (int)fread(buf.data(), 1, (size_t)chunk, f);
        if (n <= 0) break;
        in[0] = buf.data();
        total += n;
        if (blk->work(n, in, out) == -1) break;   /* WORK_DONE */
    }
    fclose(f);
    fflush(stdout);
    fprintf(stderr, "fs_scan: %ld bytes fed\n", total);
    return 0;
}

/* TEST INFRASTRUCTURE ONLY: the few GNU Radio declarations LTE_fdd_dl_file_scan's block uses (gr::sync_block, gr::io_signature, gr_complex,
 * the item vectors, boost::shared_ptr), as inert stand-ins.  oracle/fs_scan_main.cc plays the scheduler: it calls work() with consecutive
 * chunks of the capture file, as blocks.file_source -> samp_buf does in LTE_fdd_dl_file_scan.py. */
#ifndef GR_STUB_SYNC_BLOCK_H
#define GR_STUB_SYNC_BLOCK_H
#include <complex>
#include <limits.h>
#include <math.h>
#include <memory>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <strings.h>
#include <sys/types.h>
#include <vector>
#ifndef LINE_MAX
#define LINE_MAX 2048
#endif
typedef std::complex<float> gr_complex;
typedef std::vector<const void *> gr_vector_const_void_star;
typedef std::vector<void *> gr_vector_void_star;
namespace boost { template <class T> using shared_ptr = std::shared_ptr<T>; }
namespace gr {
class io_signature {
public:
    typedef std::shared_ptr<io_signature> sptr;
    static sptr make(int, int, int) { return sptr(new io_signature()); }
};
class sync_block {
public:
    sync_block(const std::string &, io_signature::sptr, io_signature::sptr) : d_consumed(0) {}
    virtual ~sync_block() {}
    void consume_each(int n) { d_consumed += n; }
    long d_consumed;
};
}  // namespace gr
#endif

/* TEST INFRASTRUCTURE ONLY: stand-in for <gnuradio/attributes.h> so that the reference's LTE_fdd_dl_fs_samp_buf.cc compiles without GNU Radio. */
#ifndef GR_STUB_ATTRIBUTES_H
#define GR_STUB_ATTRIBUTES_H
#define __GR_ATTR_EXPORT
#define __GR_ATTR_IMPORT
#endif

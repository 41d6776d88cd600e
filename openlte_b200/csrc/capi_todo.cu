// Entry points declared in include/lte_b200.h whose kernels have not landed yet: they fail loudly.
#include "internal.h"
#define TODO(ctx, name) do { lteb200_set_error(ctx, name " is not implemented yet"); return LTEB200_ERR_UNSUPPORTED; } while (0)
extern "C" {
int lteb200_find_coarse_timing_and_freq_offset(lteb200_ctx *ctx, const float *, const float *, long long, int, uint32_t, float *, uint32_t *, uint32_t *) { TODO(ctx, "find_coarse_timing"); }
int lteb200_find_pss_and_fine_timing(lteb200_ctx *ctx, const float *, const float *, long long, int, int, uint32_t *, uint32_t *, uint32_t *, float *, float *) { TODO(ctx, "find_pss"); }
int lteb200_find_sss(lteb200_ctx *ctx, const float *, const float *, long long, int, int, uint32_t, uint32_t *, float, uint32_t *, uint32_t *) { TODO(ctx, "find_sss"); }
int lteb200_prs_c_batch(lteb200_ctx *ctx, uint32_t, int, uint8_t *, void *) { TODO(ctx, "prs_c_batch"); }
int lteb200_crc24a_check_batch(lteb200_ctx *ctx, const uint8_t *, int, int, int, int32_t *, void *) { TODO(ctx, "crc24a_check_batch"); }
}

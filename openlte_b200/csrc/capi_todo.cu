// Entry points declared in include/lte_b200.h whose kernels have not landed yet: they fail loudly.
#include "internal.h"
#define TODO(ctx, name) do { lteb200_set_error(ctx, name " is not implemented yet"); return LTEB200_ERR_UNSUPPORTED; } while (0)
extern "C" {
int lteb200_prs_c_batch(lteb200_ctx *ctx, uint32_t, int, uint8_t *, void *) { TODO(ctx, "prs_c_batch"); }
int lteb200_crc24a_check_batch(lteb200_ctx *ctx, const uint8_t *, int, int, int, int32_t *, void *) { TODO(ctx, "crc24a_check_batch"); }
}

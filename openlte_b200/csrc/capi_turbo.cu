// C-ABI entry points: turbo decoding.
#include "internal.h"

extern "C" int lteb200_turbo_decode_batch(lteb200_ctx *ctx, const float *d_llr, int K, int n_cb, int mode, int n_iter,
                                          uint8_t *d_bits, void *stream)
{
    if (!ctx || !d_llr || !d_bits || n_cb < 0) return LTEB200_ERR_ARGS;
    if (n_cb == 0) return LTEB200_OK;
    LTE_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    if (mode == LTEB200_TURBO_REF_EXACT) return turbo_ref_decode_batch(ctx, d_llr, K, n_cb, d_bits, (cudaStream_t)stream);
    if (mode == LTEB200_TURBO_MAXLOGMAP) return turbo_mlm_decode_batch(ctx, d_llr, K, n_cb, n_iter, d_bits, (cudaStream_t)stream);
    lteb200_set_error(ctx, "unknown turbo mode %d", mode);
    return LTEB200_ERR_ARGS;
}

// Host twin of turbo_decode(), liblte_phy.cc:10620: d_bits is modified in place (sentinels -> 0, :10636-10642),
// N_fill_bits is ignored and *N_c_bits is never written there either, so it is not a parameter here.
extern "C" int lteb200_turbo_decode(lteb200_ctx *ctx, float *d_bits, uint32_t N_d_bits, uint32_t N_fill_bits,
                                    uint8_t *c_bits, int mode, int n_iter)
{
    (void)N_fill_bits;
    if (!ctx || !d_bits || !c_bits || N_d_bits < 3 * 44 || N_d_bits % 3) return LTEB200_ERR_ARGS;
    const int K = (int)(N_d_bits / 3) - 4;
    LTE_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    float *d_llr = nullptr;
    uint8_t *d_out = nullptr;
    LTE_CUDA_OK(ctx, cudaMalloc(&d_llr, sizeof(float) * N_d_bits));
    LTE_CUDA_OK(ctx, cudaMalloc(&d_out, K));
    LTE_CUDA_OK(ctx, cudaMemcpy(d_llr, d_bits, sizeof(float) * N_d_bits, cudaMemcpyHostToDevice));
    int rc = lteb200_turbo_decode_batch(ctx, d_llr, K, 1, mode, n_iter, d_out, nullptr);
    if (rc == LTEB200_OK) {
        cudaError_t e = cudaMemcpy(c_bits, d_out, K, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { lteb200_set_error(ctx, "D2H: %s", cudaGetErrorString(e)); rc = LTEB200_ERR_CUDA; }
    }
    cudaFree(d_llr);
    cudaFree(d_out);
    if (rc == LTEB200_OK)
        for (uint32_t i = 0; i < N_d_bits; i++)
            if (d_bits[i] == LTEB200_RX_NULL_BIT) d_bits[i] = 0.f;
    return rc;
}

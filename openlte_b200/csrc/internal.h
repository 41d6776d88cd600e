// Internal launchers shared between translation units.
#pragma once
#include "common.cuh"

int turbo_ref_decode_batch(lteb200_ctx *ctx, const float *d_llr, int K, int n_cb, uint8_t *d_bits, cudaStream_t stream);
int turbo_mlm_decode_batch(lteb200_ctx *ctx, const float *d_llr, int K, int n_cb, int n_iter, uint8_t *d_bits, cudaStream_t stream);

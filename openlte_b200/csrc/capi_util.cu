// Small utilities of the C ABI: Gold sequence and batched CRC24A check.
#include "internal.h"

void prs_c_host(uint32_t c_init, int len, uint8_t *c);

namespace {
// One warp per block of n = tbs + 24 bits: the remainder of (message || parity) must be zero.  Each lane reduces a contiguous chunk
// and multiplies by x^(bits after the chunk) mod g; the XOR of the lanes is the remainder (GF(2) linearity of calc_crc).
__global__ void crc24a_check_kernel(const uint8_t *__restrict__ bits, int stride, int n, int n_blocks, int32_t *__restrict__ ok)
{
    const int blk = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (blk >= n_blocks) return;
    const uint8_t *a = bits + (size_t)blk * stride;
    const int per = (n + 31) / 32, beg = lane * per, end = min(n, beg + per);
    uint32_t rem = 0;
    for (int i = beg; i < end; i++) {
        rem = (rem << 1) | (a[i] & 1u);
        if (rem & 0x1000000u) rem ^= 0x1864CFBu;
    }
    for (int i = end; i < n; i++) {
        rem <<= 1;
        if (rem & 0x1000000u) rem ^= 0x1864CFBu;
    }
    for (int o = 16; o > 0; o >>= 1) rem ^= __shfl_xor_sync(0xffffffffu, rem, o);
    if (lane == 0) ok[blk] = (rem & 0xffffffu) == 0 ? 1 : 0;
}
}  // namespace

extern "C" int lteb200_prs_c_batch(lteb200_ctx *ctx, uint32_t c_init, int len, uint8_t *d_c, void *stream)
{
    if (!ctx || !d_c || len < 0) return LTEB200_ERR_ARGS;
    if (len == 0) return LTEB200_OK;
    LTE_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    std::vector<uint8_t> c(len);
    prs_c_host(c_init, len, c.data());          // a 31-stage LFSR pair: serial by nature, microseconds on the host
    // pageable source: the runtime stages the bytes before the call returns, so the vector may go out of scope
    LTE_CUDA_OK(ctx, cudaMemcpyAsync(d_c, c.data(), len, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    LTE_CUDA_OK(ctx, cudaStreamSynchronize((cudaStream_t)stream));
    return LTEB200_OK;
}

extern "C" int lteb200_crc24a_check_batch(lteb200_ctx *ctx, const uint8_t *d_bits, int stride, int tbs, int n, int32_t *d_ok, void *stream)
{
    if (!ctx || !d_bits || !d_ok || tbs < 1 || stride < tbs + 24 || n < 0) return LTEB200_ERR_ARGS;
    if (n == 0) return LTEB200_OK;
    LTE_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    crc24a_check_kernel<<<ceil_div(n, 8), 256, 0, (cudaStream_t)stream>>>(d_bits, stride, tbs + 24, n, d_ok);
    ctx->launches += 1;
    LTE_CUDA_OK(ctx, cudaGetLastError());
    return LTEB200_OK;
}

// Context, scratch arena and per-block-size tables.
#include "common.cuh"
#include "qpp_table.h"
#include <stdarg.h>

void lteb200_set_error(lteb200_ctx *ctx, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->last_error = buf;
    else fprintf(stderr, "lteb200: %s\n", buf);
}

extern "C" const char *lteb200_version(void) { return "openlte_b200 0.1 (sm_100a)"; }

extern "C" int lteb200_create(lteb200_ctx **out, int device)
{
    if (!out) return LTEB200_ERR_ARGS;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        lteb200_set_error(nullptr, "no CUDA device (%s): libopenlte_b200 has no CPU fallback",
                          e == cudaSuccess ? "count 0" : cudaGetErrorString(e));
        return LTEB200_ERR_CUDA;
    }
    lteb200_ctx *c = new lteb200_ctx();
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
    c->device = device;
    c->arena = nullptr;
    c->arena_bytes = 0;
    c->launches = 0;
    c->crc24a_tab = nullptr;
    c->ring_h = c->ring_d = nullptr; c->ring_bytes = c->ring_off = 0;
    c->cfo_hz = 0.f; c->cfo_base = 0;
    c->demap = LTEB200_DEMAP_REFERENCE;
    c->tb_packed = LTEB200_TB_BYTES;
    c->qpp_spec = LTEB200_QPP_REFERENCE;
    c->s_h2d = c->s_comp = c->s_d2h = c->s_front = nullptr;
    if (cudaSetDevice(device) != cudaSuccess) { delete c; return LTEB200_ERR_CUDA; }
    // Stage buffers come from the stream-ordered allocator; keep freed blocks in the pool across synchronisations,
    // otherwise every host-pointer call would re-acquire gigabytes from the driver.
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    *out = c;
    return LTEB200_OK;
}

extern "C" void lteb200_destroy(lteb200_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    if (c->arena) cudaFree(c->arena);
    if (c->crc24a_tab) cudaFree(c->crc24a_tab);
    if (c->ring_h) cudaFreeHost(c->ring_h);
    if (c->s_h2d) { cudaStreamDestroy(c->s_h2d); cudaStreamDestroy(c->s_comp); cudaStreamDestroy(c->s_d2h); cudaStreamDestroy(c->s_front); }
    for (auto &kv : c->qpp) { cudaFree(kv.second.pi); cudaFree(kv.second.inv_last); }
    for (auto &kv : c->mlm) cudaFree(kv.second.piT);   // one blob: piT | csr_start | csr_pos
    for (auto &kv : c->mlm2) cudaFree(kv.second);
    for (auto &kv : c->rum) { cudaFree(kv.second.first_k); cudaFree(kv.second.first_kT); cudaFree(kv.second.inv_kT); }
    delete c;
}

extern "C" int lteb200_set_tb_format(lteb200_ctx *c, int format)
{
    if (!c || (format != LTEB200_TB_BYTES && format != LTEB200_TB_PACKED)) return LTEB200_ERR_ARGS;
    c->tb_packed = format;
    return LTEB200_OK;
}

extern "C" int lteb200_set_interleaver(lteb200_ctx *c, int qpp)
{
    if (!c || (qpp != LTEB200_QPP_REFERENCE && qpp != LTEB200_QPP_SPEC)) return LTEB200_ERR_ARGS;
    c->qpp_spec = qpp;              // (tables are cached per (K, mode): nothing to invalidate)
    return LTEB200_OK;
}

extern "C" int lteb200_set_demapper(lteb200_ctx *c, int demap)
{
    if (!c || (demap != LTEB200_DEMAP_REFERENCE && demap != LTEB200_DEMAP_MAXLOG)) return LTEB200_ERR_ARGS;
    c->demap = demap;
    return LTEB200_OK;
}

extern "C" const char *lteb200_last_error(lteb200_ctx *c) { return c ? c->last_error.c_str() : "null context"; }
extern "C" unsigned long long lteb200_launch_count(lteb200_ctx *c) { return c ? c->launches : 0; }

int lteb200_arena(lteb200_ctx *ctx, size_t bytes, void **ptr)
{
    if (bytes > ctx->arena_bytes) {
        // Growing invalidates the old arena: drain outstanding work first.
        LTE_CUDA_OK(ctx, cudaDeviceSynchronize());
        if (ctx->arena) cudaFree(ctx->arena);
        ctx->arena = nullptr;
        ctx->arena_bytes = 0;
        size_t want = bytes + (bytes >> 3) + (1u << 20);
        LTE_CUDA_OK(ctx, cudaMalloc(&ctx->arena, want));
        ctx->arena_bytes = want;
    }
    *ptr = ctx->arena;
    return LTEB200_OK;
}

int lteb200_get_qpp(lteb200_ctx *ctx, int K, QppTables *out)
{
    const int key = K | (ctx->qpp_spec << 16);
    auto it = ctx->qpp.find(key);
    if (it != ctx->qpp.end()) { *out = it->second; return LTEB200_OK; }
    uint32_t f1 = 0, f2 = 0;
    bool found = false;
    for (int r = 0; r < LTE_QPP_TABLE_ROWS; r++)
        if (LTE_QPP_TABLE[r][0] == K) { f1 = LTE_QPP_TABLE[r][1]; f2 = LTE_QPP_TABLE[r][2]; found = true; break; }
    if (!found) { lteb200_set_error(ctx, "K=%d is not a turbo block size (36.212 table 5.1.3-3)", K); return LTEB200_ERR_ARGS; }
    std::vector<uint16_t> pi(K);
    std::vector<int16_t> inv(K, -1);
    for (uint32_t i = 0; i < (uint32_t)K; i++) {
        // uint32 wrap-around is the reference's behaviour (liblte_phy.cc:10982); LTEB200_QPP_SPEC: the permutation of 36.212 5.1.3.2.3
        uint32_t idx = ctx->qpp_spec ? (uint32_t)(((uint64_t)f1 * i + (uint64_t)f2 * i * i) % (uint64_t)K) : (uint32_t)(f1 * i + f2 * i * i) % (uint32_t)K;
        pi[i] = (uint16_t)idx;
        inv[idx] = (int16_t)i;  // increasing i: the last writer wins, as in the sequential scatter
    }
    QppTables t;
    LTE_CUDA_OK(ctx, cudaMalloc(&t.pi, K * sizeof(uint16_t)));
    LTE_CUDA_OK(ctx, cudaMalloc(&t.inv_last, K * sizeof(int16_t)));
    LTE_CUDA_OK(ctx, cudaMemcpy(t.pi, pi.data(), K * sizeof(uint16_t), cudaMemcpyHostToDevice));
    LTE_CUDA_OK(ctx, cudaMemcpy(t.inv_last, inv.data(), K * sizeof(int16_t), cudaMemcpyHostToDevice));
    ctx->qpp[key] = t;
    *out = t;
    return LTEB200_OK;
}

int lteb200_get_crc24a_tab(lteb200_ctx *ctx, const uint32_t **out)
{
    if (!ctx->crc24a_tab) {
        std::vector<uint32_t> t(6168);
        uint32_t v = 1;
        for (int j = 0; j < 6168; j++) {        // x^j mod g, g = 0x1864CFB (calc_crc's long division, liblte_phy.cc:9713-9743)
            t[j] = v;
            v <<= 1;
            if (v & 0x1000000u) v ^= 0x1864CFBu;
        }
        LTE_CUDA_OK(ctx, cudaMalloc(&ctx->crc24a_tab, sizeof(uint32_t) * t.size()));
        LTE_CUDA_OK(ctx, cudaMemcpy(ctx->crc24a_tab, t.data(), sizeof(uint32_t) * t.size(), cudaMemcpyHostToDevice));
    }
    *out = ctx->crc24a_tab;
    return LTEB200_OK;
}

namespace {
__global__ void ring_copy_kernel(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, int n_words)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += gridDim.x * blockDim.x) dst[i] = src[i];
}
}  // namespace

int lteb200_upload_small(lteb200_ctx *ctx, const void *h, size_t bytes, void *d, cudaStream_t stream)
{
    if (bytes == 0) return LTEB200_OK;
    const size_t need = (bytes + 255) & ~(size_t)255;
    if (!ctx->ring_h) {
        ctx->ring_bytes = 8u << 20;
        LTE_CUDA_OK(ctx, cudaHostAlloc((void **)&ctx->ring_h, ctx->ring_bytes, cudaHostAllocMapped));
        LTE_CUDA_OK(ctx, cudaHostGetDevicePointer((void **)&ctx->ring_d, ctx->ring_h, 0));
        ctx->ring_off = 0;
    }
    if (need > ctx->ring_bytes / 4 || ((uintptr_t)d & 3) != 0) {            // large or unaligned: ordinary copy
        LTE_CUDA_OK(ctx, cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, stream));
        return LTEB200_OK;
    }
    if (ctx->ring_off + need > ctx->ring_bytes) {
        // wrap: everything that still reads the old contents was enqueued before this point on streams of this context
        LTE_CUDA_OK(ctx, cudaDeviceSynchronize());
        ctx->ring_off = 0;
    }
    memcpy(ctx->ring_h + ctx->ring_off, h, bytes);
    const int n_words = (int)((bytes + 3) / 4);
    ring_copy_kernel<<<n_words > 4096 ? 8 : 1, 256, 0, stream>>>((const uint32_t *)(ctx->ring_d + ctx->ring_off), (uint32_t *)d, n_words);
    LTE_CUDA_OK(ctx, cudaGetLastError());
    ctx->launches += 1;
    ctx->ring_off += need;
    return LTEB200_OK;
}

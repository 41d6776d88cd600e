// Shared host/device plumbing for libopenlte_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/lte_b200.h"

#define LTE_CUDA_OK(ctx, call)                                                                      \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            lteb200_set_error((ctx), "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return LTEB200_ERR_CUDA;                                                                \
        }                                                                                           \
    } while (0)

void lteb200_set_error(lteb200_ctx *ctx, const char *fmt, ...);

// Per-block-size tables of the turbo internal interleaver, resident in HBM.
struct QppTables {
    uint16_t *pi;        // [K]  pi(i) with the reference's uint32 wrap (liblte_phy.cc:10982)
    int16_t  *inv_last;  // [K]  last i with pi(i) == j, or -1 (scatter "last writer wins", liblte_phy.cc:11069)
};

// Rate-(un)matching gather map for one (K, E, rv, N_cb) configuration.
struct RateUnmatchMap {
    int32_t *first_k;    // [3*(K+4)] first e index feeding output d[3i+x], or -1
    int32_t *first_kT;   // [3][16 NT] the same in window-transposed order (position (k mod 16) NT + k/16; -1 on padding) + [12] tail + [4] pad
    uint16_t *inv_kT;    // [period] inverse of first_kT: image position fed by e-bit k of the first lap
    int      n_nonnull;  // period of the circular buffer in e indices
};

// Per-block-size tables of the max-log-MAP kernel (window-transposed interleaver and its CSR inverse).
struct MlmTables { uint16_t *piT, *csr_start, *csr_pos; };

struct lteb200_ctx {
    int          device;
    std::string  last_error;
    std::mutex   mu;
    // growable scratch arena (one arena per context; calls on a context are stream-ordered by the caller)
    void        *arena;
    size_t       arena_bytes;
    std::map<int, QppTables> qpp;
    std::map<int, MlmTables> mlm;
    std::map<int, uint16_t *> mlm2;   // pair kernel (turbo_mlm2.cu): 2 * position of pi(k) per position of k
    std::map<std::vector<int>, RateUnmatchMap> rum;
    cudaStream_t s_h2d, s_comp, s_d2h, s_front;   // pipeline streams of the host-pointer entry point (lazily created)
    float        cfo_hz;       // front-end frequency shift fused into the OFDM sample load (0 = off)
    long long    cfo_base;
    int          demap;        // LTEB200_DEMAP_* for the max-log-MAP receive path
    int          tb_packed;    // LTEB200_TB_*
    int          qpp_spec;     // LTEB200_QPP_*: interleaver of the turbo code (tables are cached per (K, this))
    // pinned, device-mapped ring for small per-call parameter arrays (see lteb200_upload_small)
    char        *ring_h, *ring_d;
    size_t       ring_bytes, ring_off;
    uint32_t    *crc24a_tab;   // x^j mod g(CRC24A) for j < 6168 (lazily built)
    // statistics
    unsigned long long launches;
};

// Entry points whose results are DEFINED as one bit per byte (the scan's tb_bits, the host twins' out_bits -- liblte_phy.h:906-913) run
// with the byte format whatever lteb200_set_tb_format says; the caller's choice is restored on every exit path.
struct TbBytesScope {
    lteb200_ctx *c;
    int saved;
    explicit TbBytesScope(lteb200_ctx *ctx) : c(ctx), saved(ctx->tb_packed) { c->tb_packed = 0; }
    ~TbBytesScope() { c->tb_packed = saved; }
};

// Returns a device pointer to `bytes` of scratch; grows (synchronously) when needed.
int lteb200_arena(lteb200_ctx *ctx, size_t bytes, void **ptr);

// Simple bump allocator over the arena.
struct Bump {
    char  *base;
    size_t off;
    explicit Bump(void *p) : base((char *)p), off(0) {}
    template <typename T> T *take(size_t n)
    {
        off = (off + 255) & ~(size_t)255;
        T *r = (T *)(base + off);
        off += n * sizeof(T);
        return r;
    }
    static size_t pad(size_t bytes) { return (bytes + 255) & ~(size_t)255; }
};

// Small host arrays (per-call subframe numbers, start indices, parameter blocks) -> device memory in stream order WITHOUT using a
// copy engine: the bytes go into a pinned, device-mapped ring and a one-block kernel copies them.  A cudaMemcpyAsync here would queue
// behind whatever bulk transfers the caller has in flight on the same DMA engine and stall the compute stream (host pipeline).
int lteb200_upload_small(lteb200_ctx *ctx, const void *h, size_t bytes, void *d, cudaStream_t stream);
int lteb200_get_qpp(lteb200_ctx *ctx, int K, QppTables *out);
int lteb200_get_crc24a_tab(lteb200_ctx *ctx, const uint32_t **out);

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Stream-ordered temporaries and events of one call, released on EVERY exit path (the error macro returns from the middle of a function).
struct StreamScope {
    std::vector<std::pair<void *, cudaStream_t>> mem;
    std::vector<cudaEvent_t> evs;
    template <typename T> cudaError_t alloc(T **p, size_t bytes, cudaStream_t s)
    {
        const cudaError_t e = cudaMallocAsync((void **)p, bytes ? bytes : 1, s);
        if (e == cudaSuccess) mem.emplace_back((void *)*p, s);
        return e;
    }
    cudaError_t event(cudaEvent_t *e, unsigned flags = cudaEventDisableTiming)
    {
        const cudaError_t r = cudaEventCreateWithFlags(e, flags);
        if (r == cudaSuccess) evs.push_back(*e);
        return r;
    }
    // free now, in the order of stream s (the usual case inside a pipeline loop)
    void release(void *p, cudaStream_t s)
    {
        for (size_t i = 0; i < mem.size(); i++)
            if (mem[i].first == p) { mem.erase(mem.begin() + i); break; }
        cudaFreeAsync(p, s);
    }
    ~StreamScope()
    {
        for (auto &m : mem) cudaFreeAsync(m.first, m.second);
        for (auto &e : evs) cudaEventDestroy(e);
    }
};

// Turbo rate unmatching over code-block batches: twin of liblte_phy_rate_unmatch_turbo,
// /root/reference/liblte/src/liblte_phy.cc:11246-11490 (3GPP TS 36.212 5.1.4.1 undone).
//
// The reference rebuilds the sub-block interleaver on a dummy block, walks the circular buffer bit by bit and
// de-interleaves three times -- about ten passes over K_w elements of index arithmetic per call.  All of that
// depends only on (K, E, rv, N_cb), so here it is folded ONCE on the host into a gather map: for every output
// d[3i+x] the index of the first e-bit that lands on it and the period after which the circular buffer wraps
// onto it again.  The kernel is then one coalesced gather per output; repeats are soft-combined by addition in
// the reference's visiting order (liblte_phy.cc:11407-11412), unfilled positions carry the 10000 sentinel.
// HBM-bound: reads E floats, writes 3(K+4) floats per code block.
#include "internal.h"
#include <algorithm>

namespace {

const int P32[32] = {0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30,
                     1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31};

__global__ void __launch_bounds__(256)
rate_unmatch_kernel(const float *__restrict__ e, int e_stride, int E, const int32_t *__restrict__ first_k, int period,
                    int n_out, float *__restrict__ out)
{
    const int o = blockIdx.x * 256 + threadIdx.x;
    if (o >= n_out) return;
    const float *eb = e + (size_t)blockIdx.y * e_stride;
    int k = first_k[o];
    float w = LTEB200_RX_NULL_BIT;
    if (k >= 0) {
        for (; k < E; k += period) {
            const float v = eb[k];
            if (w == LTEB200_RX_NULL_BIT) w = v;                 // first visit (or still at the sentinel)
            else if (v != LTEB200_RX_NULL_BIT) w = __fadd_rn(w, v);
        }
    }
    out[(size_t)blockIdx.y * n_out + o] = w;
}

// Same with a per-code-block number of received bits (subframes 0/5 lose REs to PBCH/PSS/SSS).
__global__ void __launch_bounds__(256)
rate_unmatch_var_kernel(const float *__restrict__ e, int e_stride, const int *__restrict__ E_cb, const int32_t *__restrict__ first_k,
                        int period, int n_out, float *__restrict__ out)
{
    const int o = blockIdx.x * 256 + threadIdx.x;
    if (o >= n_out) return;
    const float *eb = e + (size_t)blockIdx.y * e_stride;
    const int E = E_cb[blockIdx.y];
    int k = first_k[o];
    float w = LTEB200_RX_NULL_BIT;
    if (k >= 0) {
        for (; k < E; k += period) {
            const float v = eb[k];
            if (w == LTEB200_RX_NULL_BIT) w = v;
            else if (v != LTEB200_RX_NULL_BIT) w = __fadd_rn(w, v);
        }
    }
    out[(size_t)blockIdx.y * n_out + o] = w;
}

}  // namespace

int rate_unmatch_launch_var(lteb200_ctx *ctx, const float *d_e, int e_stride, const int *d_E, int D, const RateUnmatchMap &m, int n_cb,
                            float *d_out, cudaStream_t stream)
{
    for (int base = 0; base < n_cb; base += 32768) {
        const int n = (n_cb - base) < 32768 ? (n_cb - base) : 32768;
        dim3 grid(ceil_div(3 * D, 256), n);
        rate_unmatch_var_kernel<<<grid, 256, 0, stream>>>(d_e + (size_t)base * e_stride, e_stride, d_E + base, m.first_k, m.n_nonnull, 3 * D,
                                                          d_out + (size_t)base * 3 * D);
        ctx->launches += 1;
    }
    LTE_CUDA_OK(ctx, cudaGetLastError());
    return LTEB200_OK;
}

// Soft-buffer size N_cb of one code block (liblte_phy.cc:11283-11310); 0 or less: bad arguments
static int soft_buffer_size(int D, int C, int tx_mode, int N_soft, int M_dl_harq, int chan_type)
{
    const int R = (D + 31) / 32, Kw = 3 * 32 * R;
    const int K_mimo = (tx_mode == 3 || tx_mode == 4 || tx_mode == 8 || tx_mode == 9) ? 2 : 1;
    const int N_ir = N_soft / (K_mimo * (M_dl_harq < 8 ? M_dl_harq : 8));
    int Ncb = Kw;
    if (chan_type == 0 || chan_type == 1) Ncb = (N_ir / C < Kw) ? N_ir / C : Kw;
    return Ncb;
}

// Host part of the gather map: first[3i+x] = rank (among the non-<NULL> circular-buffer positions, starting at k0 of redundancy version rv)
// of the first e-bit that lands on output d[3i+x], -1 if none does within one lap.  Returns the number of ranks = the period after which
// the circular buffer wraps onto the same outputs again.  Geometry exactly as the reference computes it (liblte_phy.cc:11283-11400).
static int build_first_map(int D, int Ncb, int rv, std::vector<int32_t> &first)
{
    const int R = (D + 31) / 32, Kpi = 32 * R, ND = Kpi - D, Kw = 3 * Kpi;
    const int k0 = (int)(R * (2 * ceilf((float)Ncb / (float)(8 * R)) * rv + 2));
    // circular-buffer position -> (stream x, index t in the padded stream); t < ND are the <NULL> pads
    std::vector<int> sx(Kw), st(Kw);
    for (int n = 0; n < Kpi; n++) {
        const int col = n / R, row = n % R;
        const int t01 = 32 * row + P32[col];
        const int t2 = (P32[col] + 32 * row + 1) % Kpi;
        sx[n] = 0; st[n] = t01;
        sx[Kpi + 2 * n] = 1; st[Kpi + 2 * n] = t01;
        sx[Kpi + 2 * n + 1] = 2; st[Kpi + 2 * n + 1] = t2;
    }
    first.assign(3 * (size_t)D, -1);
    int rank = 0;
    for (int j = 0; j < Ncb; j++) {
        const int pos = (k0 + j) % Ncb;
        if (st[pos] < ND) continue;
        first[3 * (st[pos] - ND) + sx[pos]] = rank++;
    }
    return rank;
}

// HOST, no device needed: the table above for inspection and CPU tests
extern "C" int lteb200_rate_unmatch_map(int K, int N_codeblocks, int tx_mode, int N_soft, int M_dl_harq, int chan_type, int rv_idx,
                                        int32_t *first_k_out, int *period)
{
    const int D = K + 4;
    if (!first_k_out || !period || D < 44 || D > 6148 || N_codeblocks < 1 || rv_idx < 0 || rv_idx > 3 || M_dl_harq < 1) return LTEB200_ERR_ARGS;
    const int Ncb = soft_buffer_size(D, N_codeblocks, tx_mode, N_soft, M_dl_harq, chan_type);
    if (Ncb <= 0) return LTEB200_ERR_ARGS;
    std::vector<int32_t> first;
    const int rank = build_first_map(D, Ncb, rv_idx, first);
    if (rank == 0) return LTEB200_ERR_ARGS;
    std::copy(first.begin(), first.end(), first_k_out);
    *period = rank;
    return LTEB200_OK;
}

// Builds (or fetches) the gather map on the device.
int rate_unmatch_get_map(lteb200_ctx *ctx, int D, int E, int C, int tx_mode, int N_soft, int M_dl_harq, int chan_type,
                         int rv, RateUnmatchMap *out)
{
    if (D < 44 || D > 6148 || E < 0 || C < 1 || rv < 0 || rv > 3 || M_dl_harq < 1) {
        lteb200_set_error(ctx, "rate_unmatch: bad arguments");
        return LTEB200_ERR_ARGS;
    }
    const int Ncb = soft_buffer_size(D, C, tx_mode, N_soft, M_dl_harq, chan_type);
    if (Ncb <= 0) { lteb200_set_error(ctx, "rate_unmatch: empty soft buffer"); return LTEB200_ERR_ARGS; }
    std::vector<int> key = {D, Ncb, rv};
    auto it = ctx->rum.find(key);
    if (it != ctx->rum.end()) { *out = it->second; return LTEB200_OK; }
    std::vector<int32_t> first;
    const int rank = build_first_map(D, Ncb, rv, first);
    if (rank == 0) { lteb200_set_error(ctx, "rate_unmatch: no transmittable position"); return LTEB200_ERR_ARGS; }
    RateUnmatchMap m;
    m.n_nonnull = rank;
    LTE_CUDA_OK(ctx, cudaMalloc(&m.first_k, sizeof(int32_t) * 3 * D));
    LTE_CUDA_OK(ctx, cudaMemcpy(m.first_k, first.data(), sizeof(int32_t) * 3 * D, cudaMemcpyHostToDevice));
    {
        // output order of the pair-lane decoder's input image (pdsch_llr.cu): three planes of 16 NT positions, then the 12 tail values
        const int K = D - 4, NT = mlm_size_class(K), NP = 16 * NT;
        std::vector<int32_t> T((size_t)3 * NP + 16, -1);
        for (int k = 0; k < K; k++)
            for (int x = 0; x < 3; x++) T[(size_t)x * NP + (k % 16) * NT + k / 16] = first[3 * k + x];
        for (int j = 0; j < 12; j++) T[(size_t)3 * NP + j] = first[3 * K + j];
        LTE_CUDA_OK(ctx, cudaMalloc(&m.first_kT, sizeof(int32_t) * T.size()));
        LTE_CUDA_OK(ctx, cudaMemcpy(m.first_kT, T.data(), sizeof(int32_t) * T.size(), cudaMemcpyHostToDevice));
        // the inverse: image position of the output that e-bit k (first lap, k < period) feeds; every e-bit of a lap feeds exactly one output
        std::vector<uint16_t> inv(((size_t)rank + 7) & ~(size_t)7, 0xffffu);
        for (size_t j = 0; j < T.size(); j++) if (T[j] >= 0) inv[T[j]] = (uint16_t)j;
        LTE_CUDA_OK(ctx, cudaMalloc(&m.inv_kT, sizeof(uint16_t) * inv.size()));
        LTE_CUDA_OK(ctx, cudaMemcpy(m.inv_kT, inv.data(), sizeof(uint16_t) * inv.size(), cudaMemcpyHostToDevice));
    }
    ctx->rum[key] = m;
    *out = m;
    return LTEB200_OK;
}

int rate_unmatch_launch(lteb200_ctx *ctx, const float *d_e, int e_stride, int E, int D, const RateUnmatchMap &m, int n_cb,
                        float *d_out, cudaStream_t stream)
{
    dim3 grid(ceil_div(3 * D, 256), n_cb);
    rate_unmatch_kernel<<<grid, 256, 0, stream>>>(d_e, e_stride, E, m.first_k, m.n_nonnull, 3 * D, d_out);
    ctx->launches += 1;
    LTE_CUDA_OK(ctx, cudaGetLastError());
    return LTEB200_OK;
}

extern "C" int lteb200_rate_unmatch_turbo_batch(lteb200_ctx *ctx, const float *d_e, int e_stride, int E, int K, int N_codeblocks,
                                                int tx_mode, int N_soft, int M_dl_harq, int chan_type, int rv_idx, int n_cb,
                                                float *d_out, void *stream)
{
    if (!ctx || !d_e || !d_out || n_cb < 0 || e_stride < E) return LTEB200_ERR_ARGS;
    if (n_cb > 65535) { lteb200_set_error(ctx, "at most 65535 code blocks per call (split the batch)"); return LTEB200_ERR_ARGS; }
    if (n_cb == 0) return LTEB200_OK;
    LTE_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    RateUnmatchMap m;
    int rc = rate_unmatch_get_map(ctx, K + 4, E, N_codeblocks, tx_mode, N_soft, M_dl_harq, chan_type, rv_idx, &m);
    if (rc) return rc;
    // grid.y is limited to 65535
    for (int base = 0; base < n_cb; base += 32768) {
        const int n = (n_cb - base) < 32768 ? (n_cb - base) : 32768;
        rc = rate_unmatch_launch(ctx, d_e + (size_t)base * e_stride, e_stride, E, K + 4, m, n,
                                 d_out + (size_t)base * 3 * (K + 4), (cudaStream_t)stream);
        if (rc) return rc;
    }
    return LTEB200_OK;
}

// Host twin of liblte_phy_rate_unmatch_turbo (liblte_phy.h:1311-1323).  dummy_bits' content is ignored: a uint8 can
// never equal the 10000 sentinel, so only N_dummy_bits matters in the reference as well.
extern "C" int lteb200_rate_unmatch_turbo(lteb200_ctx *ctx, const float *e_bits, uint32_t N_e_bits, const uint8_t *dummy_bits,
                                          uint32_t N_dummy_bits, uint32_t N_codeblocks, uint32_t tx_mode, uint32_t N_soft,
                                          uint32_t M_dl_harq, int chan_type, uint32_t rv_idx, float *d_bits, uint32_t *N_d_bits)
{
    (void)dummy_bits;
    if (!ctx || !e_bits || !d_bits || !N_d_bits || N_dummy_bits < 44) return LTEB200_ERR_ARGS;
    LTE_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    const int D = (int)N_dummy_bits;
    float *de = nullptr, *dd = nullptr;
    LTE_CUDA_OK(ctx, cudaMalloc(&de, sizeof(float) * (N_e_bits ? N_e_bits : 1)));
    LTE_CUDA_OK(ctx, cudaMalloc(&dd, sizeof(float) * 3 * D));
    LTE_CUDA_OK(ctx, cudaMemcpy(de, e_bits, sizeof(float) * N_e_bits, cudaMemcpyHostToDevice));
    int rc = lteb200_rate_unmatch_turbo_batch(ctx, de, (int)N_e_bits, (int)N_e_bits, D - 4, (int)N_codeblocks, (int)tx_mode,
                                              (int)N_soft, (int)M_dl_harq, chan_type, (int)rv_idx, 1, dd, nullptr);
    if (rc == LTEB200_OK) {
        cudaError_t e = cudaMemcpy(d_bits, dd, sizeof(float) * 3 * D, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { lteb200_set_error(ctx, "D2H: %s", cudaGetErrorString(e)); rc = LTEB200_ERR_CUDA; }
        else *N_d_bits = 3 * D;
    }
    cudaFree(de);
    cudaFree(dd);
    return rc;
}

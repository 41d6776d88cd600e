// Downlink transmit chain for batches of subframes: transport blocks -> I/Q samples, the inverse of the receive
// path, used to synthesise inputs on the device.  Twins of (all in /root/reference/liblte/src/liblte_phy.cc):
//   liblte_phy_pdsch_channel_encode :3489-3680 (RE walk shared with the receiver)     dlsch_channel_encode :12664-12753
//   calc_crc (CRC24A) :9713-9743      turbo_encode :10541-10586 (with the reference's wrapped interleaver, :10982)
//   liblte_phy_rate_match_turbo :11081-11240      generate_prs_c :9669      modulation_mapper :8704-8790
//   liblte_phy_map_crs :5144-5240 (port 0)        liblte_phy_create_dl_subframe :5862 / symbols_to_samples_dl :8484-8530
// One antenna port, single-code-block transport blocks with TBS + 24 in the interleaver table (no filler bits: the reference's
// encoder does not handle them).  The resource grid is bit-identical to the reference transmitter's; samples agree to float
// rounding of the IFFT (tests/test_tx_gpu.py).
#include "internal.h"
#include "fft_dev.cuh"
#include <math.h>

namespace {

struct TxAlloc {
    const int *re_idx;        // [3][idx_stride]
    const uint8_t *cseq;      // [10][c_stride]
    const int32_t *order;     // [period] circular-buffer visiting order -> index into the interleaved code word d[3i+x]
    int idx_stride, c_stride, M[3], Qm, mod, period, tbs, a_idx;
};
struct TxParams {
    TxAlloc al[8];
    int n_al, n_sf, n_alloc_total, tb_stride, nsc, K;
    const int *subfr_num;
    const uint8_t *tb;        // [n_sf][n_alloc_total][tb_stride]
    const uint16_t *pi;       // [K]
    const uint32_t *crc_tab;  // x^j mod g
    float2 *grid;             // [n_sf][14][nsc]
};

// rank in the circular buffer -> code-word index, from the receiver's first_k (rank of every code-word bit, -1 if never sent)
__global__ void invert_order_kernel(const int32_t *__restrict__ first_k, int n, int32_t *__restrict__ order)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && first_k[i] >= 0) order[first_k[i]] = i;
}

__device__ __forceinline__ float2 map_symbol(const uint8_t *b, int mod)
{
    // modulation_mapper: 36.211 7.1.1-7.1.4 with the reference's float constants
    const float s0 = 1.f - 2.f * (float)b[0];
    if (mod == LTEB200_MOD_BPSK) { const float v = __fdiv_rn(s0, (float)sqrt(2.0)); return make_float2(v, v); }
    const float s1 = 1.f - 2.f * (float)b[1];
    if (mod == LTEB200_MOD_QPSK) { const float r = (float)(1.0 / sqrt(2.0)); return make_float2(__fmul_rn(s0, r), __fmul_rn(s1, r)); }
    const float s2 = 1.f - 2.f * (float)b[2], s3 = 1.f - 2.f * (float)b[3];
    if (mod == LTEB200_MOD_16QAM) {
        const float r = (float)(1.0 / sqrt(10.0));
        return make_float2(__fmul_rn(__fmul_rn(s0, __fsub_rn(2.f, s2)), r), __fmul_rn(__fmul_rn(s1, __fsub_rn(2.f, s3)), r));
    }
    const float s4 = 1.f - 2.f * (float)b[4], s5 = 1.f - 2.f * (float)b[5];
    const float r = (float)(1.0 / sqrt(42.0));
    const float i = __fmul_rn(s0, __fsub_rn(4.f, __fmul_rn(s2, __fsub_rn(2.f, s4))));
    const float q = __fmul_rn(s1, __fsub_rn(4.f, __fmul_rn(s3, __fsub_rn(2.f, s5))));
    return make_float2(__fmul_rn(i, r), __fmul_rn(q, r));
}

// One CTA per transport block: CRC24A, turbo encoding, rate matching, scrambling, modulation, resource mapping.
__global__ void __launch_bounds__(256)
dlsch_encode_kernel(const __grid_constant__ TxParams P)
{
    extern __shared__ uint8_t sm_u8[];
    uint8_t *c = sm_u8;                       // [K] code block = transport block + CRC24A
    uint8_t *d = c + P.K;                     // [3(K+4)] code word, interleaved d[3i+x] (the receiver's layout)
    __shared__ uint32_t s_crc[8];
    __shared__ uint8_t s_tail[12];            // x1,z1 (3 each) then x2,z2
    const int a_loc = blockIdx.x / P.n_sf, sf = blockIdx.x - a_loc * P.n_sf;
    const TxAlloc &A = P.al[a_loc];
    const int K = P.K, tbs = A.tbs, tid = threadIdx.x;
    const uint8_t *tb = P.tb + ((size_t)sf * P.n_alloc_total + A.a_idx) * P.tb_stride;
    // CRC24A by linearity: parity = sum over set bits i of x^(tbs-1-i+24) mod g
    uint32_t rem = 0;
    for (int i = tid; i < tbs; i += blockDim.x) {
        const uint8_t b = tb[i] & 1;
        c[i] = b;
        if (b) rem ^= P.crc_tab[tbs - 1 - i + 24];
    }
    for (int o = 16; o > 0; o >>= 1) rem ^= __shfl_xor_sync(0xffffffffu, rem, o);
    if ((tid & 31) == 0) s_crc[tid >> 5] = rem;
    __syncthreads();
    if (tid < 24) {
        uint32_t r = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) r ^= s_crc[w];
        c[tbs + tid] = (uint8_t)((r >> (23 - tid)) & 1u);
    }
    __syncthreads();
    // the two constituent encoders run in two threads (feedback 1+D^2+D^3, feed-forward 1+D+D^3), three termination steps each
    if (tid == 0 || tid == 32) {
        const bool second = tid == 32;
        int s1 = 0, s2 = 0, s3 = 0;
        for (int i = 0; i < K; i++) {
            const int u = second ? c[P.pi[i]] : c[i];
            const int a = u ^ s2 ^ s3, z = a ^ s1 ^ s3;
            if (!second) { d[3 * i] = (uint8_t)u; d[3 * i + 1] = (uint8_t)z; } else d[3 * i + 2] = (uint8_t)z;
            s3 = s2; s2 = s1; s1 = a;
        }
        for (int t = 0; t < 3; t++) {           // termination: the input equals the feedback, so the register input is 0
            const int x = s2 ^ s3, z = s1 ^ s3;
            s_tail[(second ? 6 : 0) + t] = (uint8_t)x; s_tail[(second ? 6 : 0) + 3 + t] = (uint8_t)z;
            s3 = s2; s2 = s1; s1 = 0;
        }
    }
    __syncthreads();
    if (tid == 0) {                             // tail multiplexing, liblte_phy.cc:10575-10586
        const uint8_t *x1 = s_tail, *z1 = s_tail + 3, *x2 = s_tail + 6, *z2 = s_tail + 9;
        uint8_t *t = d + 3 * K;
        t[0] = x1[0]; t[1] = z1[0]; t[2] = x1[1];
        t[3] = z1[1]; t[4] = x1[2]; t[5] = z1[2];
        t[6] = x2[0]; t[7] = z2[0]; t[8] = x2[1];
        t[9] = z2[1]; t[10] = x2[2]; t[11] = z2[2];
    }
    __syncthreads();
    // rate matching (circular buffer), scrambling, modulation, mapping: one thread per resource element
    const int sn = P.subfr_num[sf];
    const int cls = sn == 0 ? 0 : (sn == 5 ? 1 : 2);
    const int M = A.M[cls];
    const int *ix = A.re_idx + cls * A.idx_stride;
    const uint8_t *cs = A.cseq + (size_t)sn * A.c_stride;
    float2 *g = P.grid + (size_t)sf * 14 * P.nsc;
    for (int r = tid; r < M; r += blockDim.x) {
        uint8_t b[6];
        for (int t = 0; t < A.Qm; t++) {
            const int k = r * A.Qm + t;
            b[t] = d[A.order[k % A.period]] ^ cs[k];
        }
        g[ix[r]] = map_symbol(b, A.mod);
    }
}

// CRS of antenna port 0 into an otherwise zeroed grid: symbols 0, 4, 7, 11 (liblte_phy_map_crs, liblte_phy.cc:5144-5240)
__global__ void crs_map_kernel(const float2 *__restrict__ crs /*[20][3][220]*/, const int *__restrict__ subfr_num, int N_rb_dl, int N_id_cell, int nsc,
                               float2 *__restrict__ grid)
{
    const int sf = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 4 * 2 * N_rb_dl) return;
    const int which = t / (2 * N_rb_dl), j = t - which * 2 * N_rb_dl;
    const int sym = which == 0 ? 0 : (which == 1 ? 4 : (which == 2 ? 7 : 11));
    const int slot = which >> 1, li = (which & 1) ? 2 : 0, v = (which & 1) ? 3 : 0;     // l = 0 -> table row 0, l = 4 -> row 2
    const int ns = 2 * subfr_num[sf] + slot;
    grid[((size_t)sf * 14 + sym) * nsc + 6 * j + (v + N_id_cell % 6) % 6] = crs[((size_t)ns * 3 + li) * 220 + j + 110 - N_rb_dl];
}

// symbols_to_samples_dl: bins 1..h <- subcarriers h..2h-1, bins N-h..N-1 <- subcarriers 0..h-1, unnormalised inverse DFT, cyclic prefix.
// The inverse transform is conj(FFT(conj(X))) on the shared forward FFT.  out: [n_sf][samples per subframe] float2.
__global__ void __launch_bounds__(512)
ofdm_mod_kernel(const float2 *__restrict__ grid, int N, int cp0, int cpe, int nsc, const float2 *__restrict__ tw, float2 *__restrict__ out)
{
    extern __shared__ float2 sm_f2[];
    float2 *bufA = sm_f2, *bufB = sm_f2 + N;
    const int sf = blockIdx.x / 14, sym = blockIdx.x - sf * 14, h = nsc / 2;
    const float2 *g = grid + ((size_t)sf * 14 + sym) * nsc;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        float2 v = make_float2(0.f, 0.f);
        if (i >= 1 && i <= h) v = g[h + i - 1];
        else if (i >= N - h) v = g[i - (N - h)];
        bufA[i] = make_float2(v.x, -v.y);
    }
    __syncthreads();
    const float2 *x = stockham_forward(bufA, bufB, N, tw);
    const int slot = sym / 7, l = sym % 7;
    const int cp = l == 0 ? cp0 : cpe;
    const long long base = (long long)sf * (14LL * N + 2LL * cp0 + 12LL * cpe) + (long long)slot * (7LL * N + cp0 + 6LL * cpe) +
                           (long long)l * (N + cpe) + (l > 0 ? cp0 - cpe : 0);
    for (int i = threadIdx.x; i < N + cp; i += blockDim.x) {
        const float2 v = x[i < cp ? N - cp + i : i - cp];
        out[base + i] = make_float2(v.x, -v.y);
    }
}

std::map<std::vector<int>, int32_t *> g_tx_order;     // (device, D, Ncb-key..., rv) -> inverted circular-buffer order
std::mutex g_tx_mu;

}  // namespace

extern "C" int lteb200_dl_transmit_batch(lteb200_ctx *ctx, const uint8_t *d_tb_bits, int tb_stride, const lteb200_alloc *allocs, int n_alloc, int fs_idx,
                                         int N_rb_dl, int N_id_cell, int N_ant, const int *subfr_num, int n_sf, int N_pdcch_symbs, float *d_grid,
                                         float *d_iq, void *stream_)
{
    if (!ctx || !d_tb_bits || !allocs || !subfr_num || n_sf < 1 || n_alloc < 1 || fs_idx < 0 || fs_idx > 4 || N_id_cell < 0 || N_id_cell > 503 ||
        (!d_grid && !d_iq))
        return LTEB200_ERR_ARGS;
    if (N_ant != 1) { lteb200_set_error(ctx, "dl_transmit: one antenna port only"); return LTEB200_ERR_UNSUPPORTED; }
    if (n_sf > 65535) { lteb200_set_error(ctx, "at most 65535 subframes per call (split the batch)"); return LTEB200_ERR_ARGS; }
    const int N = 128 << fs_idx, nsc = 12 * N_rb_dl;
    if (nsc >= N) { lteb200_set_error(ctx, "dl_transmit: N_rb_dl does not fit the sample rate"); return LTEB200_ERR_ARGS; }
    for (int i = 0; i < n_sf; i++) if (subfr_num[i] < 0 || subfr_num[i] > 9) return LTEB200_ERR_ARGS;
    cudaStream_t stream = (cudaStream_t)stream_;
    LTE_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> lk(g_tx_mu);
    float2 *grid = (float2 *)d_grid;
    if (!grid) LTE_CUDA_OK(ctx, cudaMallocAsync(&grid, sizeof(float2) * 14 * nsc * (size_t)n_sf, stream));
    int *d_sn = nullptr;
    LTE_CUDA_OK(ctx, cudaMallocAsync(&d_sn, sizeof(int) * n_sf, stream));
    int rc = lteb200_upload_small(ctx, subfr_num, sizeof(int) * n_sf, d_sn, stream);
    auto release = [&]() { cudaFreeAsync(d_sn, stream); if (!d_grid) cudaFreeAsync(grid, stream); };
    if (rc) { release(); return rc; }
    LTE_CUDA_OK(ctx, cudaMemsetAsync(grid, 0, sizeof(float2) * 14 * nsc * (size_t)n_sf, stream));
    float2 *crs = nullptr;
    rc = get_crs_table(ctx, N_id_cell, &crs);
    if (rc) { release(); return rc; }
    crs_map_kernel<<<dim3(ceil_div(8 * N_rb_dl, 128), n_sf), 128, 0, stream>>>(crs, d_sn, N_rb_dl, N_id_cell, nsc, grid);
    ctx->launches += 1;
    const uint32_t *crc_tab = nullptr;
    rc = lteb200_get_crc24a_tab(ctx, &crc_tab);
    // transport blocks: groups of up to 8 allocations with the same code block size share a launch
    std::vector<AllocTables> T(n_alloc);
    for (int a = 0; a < n_alloc && rc == LTEB200_OK; a++) {
        rc = get_alloc_tables(ctx, N_rb_dl, N_id_cell, N_ant, N_pdcch_symbs, allocs[a], &T[a]);
        if (rc == LTEB200_OK && (T[a].F != 0 || T[a].K != allocs[a].tbs + 24)) {
            lteb200_set_error(ctx, "dl_transmit: TBS %d needs filler bits (K = %d), which the reference's encoder does not produce", allocs[a].tbs, T[a].K);
            rc = LTEB200_ERR_UNSUPPORTED;
        }
        if (rc == LTEB200_OK && tb_stride < allocs[a].tbs) rc = LTEB200_ERR_ARGS;
    }
    for (int a = 0; a < n_alloc && rc == LTEB200_OK;) {
        TxParams P;
        memset(&P, 0, sizeof(P));
        const int K = T[a].K;
        QppTables qt;
        rc = lteb200_get_qpp(ctx, K, &qt);
        if (rc) break;
        int e = a;
        while (e < n_alloc && e - a < 8 && T[e].K == K) {
            RateUnmatchMap m;
            rc = rate_unmatch_get_map(ctx, K + 4, 0, 1, allocs[e].tx_mode, 250368, 8, 0, allocs[e].rv_idx, &m);
            if (rc) break;
            const std::vector<int> key = {ctx->device, K, allocs[e].tx_mode, allocs[e].rv_idx};
            int32_t *order = nullptr;
            auto it = g_tx_order.find(key);
            if (it != g_tx_order.end()) order = it->second;
            else {
                LTE_CUDA_OK(ctx, cudaMalloc(&order, sizeof(int32_t) * m.n_nonnull));
                invert_order_kernel<<<ceil_div(3 * (K + 4), 256), 256, 0, stream>>>(m.first_k, 3 * (K + 4), order);
                ctx->launches += 1;
                g_tx_order[key] = order;
            }
            TxAlloc &A = P.al[e - a];
            A.re_idx = T[e].re_idx; A.cseq = T[e].cseq; A.order = order; A.idx_stride = T[e].idx_stride; A.c_stride = T[e].c_stride;
            for (int c = 0; c < 3; c++) A.M[c] = T[e].M[c];
            A.Qm = T[e].Qm; A.mod = allocs[e].mod_type; A.period = m.n_nonnull; A.tbs = allocs[e].tbs; A.a_idx = e;
            e++;
        }
        if (rc) break;
        P.n_al = e - a; P.n_sf = n_sf; P.n_alloc_total = n_alloc; P.tb_stride = tb_stride; P.nsc = nsc; P.K = K;
        P.subfr_num = d_sn; P.tb = d_tb_bits; P.pi = qt.pi; P.crc_tab = crc_tab; P.grid = grid;
        const size_t smem = (size_t)K + 3 * (K + 4);
        dlsch_encode_kernel<<<P.n_al * n_sf, 256, smem, stream>>>(P);
        ctx->launches += 1;
        a = e;
    }
    if (rc == LTEB200_OK && d_iq) {
        float2 *tw = nullptr;
        rc = get_twiddles(ctx, N, &tw);
        if (rc == LTEB200_OK) {
            const int threads = N / 4 < 64 ? 64 : (N / 4 > 512 ? 512 : N / 4);
            ofdm_mod_kernel<<<n_sf * 14, threads, sizeof(float2) * 2 * N, stream>>>(grid, N, 10 << fs_idx, 9 << fs_idx, nsc, tw, (float2 *)d_iq);
            ctx->launches += 1;
        }
    }
    if (rc == LTEB200_OK) LTE_CUDA_OK(ctx, cudaGetLastError());
    release();
    return rc;
}

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
    float2 *grid;             // [n_sf][N_ant][14][nsc]
    int N_ant;
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
    const size_t ps = (size_t)14 * P.nsc;
    float2 *g = P.grid + (size_t)sf * P.N_ant * ps;
    auto symbol = [&](int r) -> float2 {
        uint8_t b[6];
        for (int t = 0; t < A.Qm; t++) {
            const int k = r * A.Qm + t;
            b[t] = d[A.order[k % A.period]] ^ cs[k];
        }
        return map_symbol(b, A.mod);
    };
    if (P.N_ant == 1) {
        for (int r = tid; r < M; r += blockDim.x) g[ix[r]] = symbol(r);
        return;
    }
    // transmit diversity (layer_mapper_dl + pre_coder_dl, liblte_phy.cc:7293-7345, :7526-7640): groups of N_ant symbols, M a multiple of N_ant
    const float r2 = (float)(1 / sqrt(2.0));
    auto sc = [&](float2 v, float sr, float si) { return make_float2(__fmul_rn(sr, v.x), __fmul_rn(si, v.y)); };
    for (int r = tid * P.N_ant; r < M; r += blockDim.x * P.N_ant) {
        if (P.N_ant == 2) {
            const float2 d0 = symbol(r), d1 = symbol(r + 1);
            g[ix[r]] = sc(d0, r2, r2);          g[ix[r + 1]] = sc(d1, r2, r2);
            g[ps + ix[r]] = sc(d1, -r2, r2);    g[ps + ix[r + 1]] = sc(d0, r2, -r2);
        } else {
            const float2 d0 = symbol(r), d1 = symbol(r + 1), d2 = symbol(r + 2), d3 = symbol(r + 3), z = make_float2(0.f, 0.f);
            g[ix[r]] = sc(d0, r2, r2);              g[ix[r + 1]] = sc(d1, r2, r2);              g[ix[r + 2]] = z;                           g[ix[r + 3]] = z;
            g[ps + ix[r]] = z;                      g[ps + ix[r + 1]] = z;                      g[ps + ix[r + 2]] = sc(d2, r2, r2);         g[ps + ix[r + 3]] = sc(d3, r2, r2);
            g[2 * ps + ix[r]] = sc(d1, -r2, r2);    g[2 * ps + ix[r + 1]] = sc(d0, r2, -r2);    g[2 * ps + ix[r + 2]] = z;                  g[2 * ps + ix[r + 3]] = z;
            g[3 * ps + ix[r]] = z;                  g[3 * ps + ix[r + 1]] = z;                  g[3 * ps + ix[r + 2]] = sc(d3, -r2, r2);    g[3 * ps + ix[r + 3]] = sc(d2, r2, -r2);
        }
    }
}

// CRS of every antenna port into an otherwise zeroed grid (liblte_phy_map_crs, liblte_phy.cc:5144-5240): ports 0 / 1 in symbols 0, 4, 7, 11
// with frequency offsets (0,3,0,3) / (3,0,3,0), ports 2 / 3 in symbols 1, 8 with offsets (0,3) / (3,6); sequence of (slot, l) from the cell's table.
__global__ void crs_map_kernel(const float2 *__restrict__ crs /*[20][3][220]*/, const int *__restrict__ subfr_num, int N_rb_dl, int N_id_cell, int nsc,
                               int N_ant, float2 *__restrict__ grid /*[n_sf][N_ant][14][nsc]*/)
{
    const int sf = blockIdx.y, p = blockIdx.z, t = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_sym = p < 2 ? 4 : 2;
    if (t >= n_sym * 2 * N_rb_dl) return;
    const int which = t / (2 * N_rb_dl), j = t - which * 2 * N_rb_dl;
    int sym, v;
    if (p < 2) { sym = which == 0 ? 0 : (which == 1 ? 4 : (which == 2 ? 7 : 11)); v = ((which & 1) ^ p) ? 3 : 0; }
    else { sym = which == 0 ? 1 : 8; v = (which == 0 ? 0 : 3) + (p == 3 ? 3 : 0); }
    const int slot = sym / 7, l = sym % 7, li = l == 0 ? 0 : (l == 1 ? 1 : 2);
    const int ns = 2 * subfr_num[sf] + slot;
    grid[(((size_t)sf * N_ant + p) * 14 + sym) * nsc + 6 * j + (v + N_id_cell % 6) % 6] = crs[((size_t)ns * 3 + li) * 220 + j + 110 - N_rb_dl];
}

// symbols_to_samples_dl: bins 1..h <- subcarriers h..2h-1, bins N-h..N-1 <- subcarriers 0..h-1, unnormalised inverse DFT, cyclic prefix.
// The inverse transform is conj(FFT(conj(X))) on the shared forward FFT.  out: [n_sf][samples per subframe] float2.
__global__ void __launch_bounds__(512)
ofdm_mod_kernel(const float2 *__restrict__ grid, int N, int cp0, int cpe, int nsc, const float2 *__restrict__ tw, float2 *__restrict__ out, int N_ant,
                float4 gain_re, float4 gain_im)
{
    extern __shared__ float2 sm_f2[];
    float2 *bufA = sm_f2, *bufB = sm_f2 + N;
    const int sf = blockIdx.x / 14, sym = blockIdx.x - sf * 14, h = nsc / 2;
    const int slot = sym / 7, l = sym % 7;
    const int cp = l == 0 ? cp0 : cpe;
    const long long base = (long long)sf * (14LL * N + 2LL * cp0 + 12LL * cpe) + (long long)slot * (7LL * N + cp0 + 6LL * cpe) +
                           (long long)l * (N + cpe) + (l > 0 ? cp0 - cpe : 0);
    const float gr[4] = {gain_re.x, gain_re.y, gain_re.z, gain_re.w}, gi[4] = {gain_im.x, gain_im.y, gain_im.z, gain_im.w};
    constexpr int MAX_PER_THREAD = 36;                       // (N + cp) / blockDim.x <= (2048 + 160) / 64
    float2 acc[MAX_PER_THREAD];
    for (int p = 0; p < N_ant; p++) {
        const float2 *g = grid + (((size_t)sf * N_ant + p) * 14 + sym) * nsc;
        __syncthreads();
        for (int i = threadIdx.x; i < N; i += blockDim.x) {
            float2 v = make_float2(0.f, 0.f);
            if (i >= 1 && i <= h) v = g[h + i - 1];
            else if (i >= N - h) v = g[i - (N - h)];
            bufA[i] = make_float2(v.x, -v.y);
        }
        __syncthreads();
        const float2 *x = stockham_forward(bufA, bufB, N, tw);
        int q = 0;
        for (int i = threadIdx.x; i < N + cp; i += blockDim.x, q++) {
            const float2 c = x[i < cp ? N - cp + i : i - cp];
            const float2 v = make_float2(c.x, -c.y);
            if (N_ant == 1 && gr[0] == 1.0f && gi[0] == 0.0f) { out[base + i] = v; continue; }     // one port, unit channel: the signal itself
            // flat channel of port p, the ports add at the receiver
            const float2 w = make_float2(__fsub_rn(__fmul_rn(gr[p], v.x), __fmul_rn(gi[p], v.y)), __fadd_rn(__fmul_rn(gr[p], v.y), __fmul_rn(gi[p], v.x)));
            acc[q] = p == 0 ? w : make_float2(__fadd_rn(acc[q].x, w.x), __fadd_rn(acc[q].y, w.y));
            if (p == N_ant - 1) out[base + i] = acc[q];
        }
    }
}

std::map<std::vector<int>, int32_t *> g_tx_order;     // (device, D, Ncb-key..., rv) -> inverted circular-buffer order
std::mutex g_tx_mu;

}  // namespace

extern "C" int lteb200_dl_transmit_batch(lteb200_ctx *ctx, const uint8_t *d_tb_bits, int tb_stride, const lteb200_alloc *allocs, int n_alloc, int fs_idx,
                                         int N_rb_dl, int N_id_cell, int N_ant, const int *subfr_num, int n_sf, int N_pdcch_symbs, float *d_grid,
                                         float *d_iq, void *stream_)
{
    if (N_ant != 1) { if (ctx) lteb200_set_error(ctx, "dl_transmit_batch: one antenna port (lteb200_dl_transmit_frames_batch takes 2 and 4)"); return ctx ? LTEB200_ERR_UNSUPPORTED : LTEB200_ERR_ARGS; }
    lteb200_tx_control ctl;
    memset(&ctl, 0, sizeof(ctl));
    ctl.N_id_cell = N_id_cell; ctl.N_ant = 1; ctl.ant_gain_re[0] = 1.0f;
    return lteb200_dl_transmit_frames_batch(ctx, d_tb_bits, tb_stride, allocs, n_alloc, fs_idx, N_rb_dl, &ctl, subfr_num, n_sf, N_pdcch_symbs, d_grid, d_iq, stream_);
}

extern "C" int lteb200_dl_transmit_frames_batch(lteb200_ctx *ctx, const uint8_t *d_tb_bits, int tb_stride, const lteb200_alloc *allocs, int n_alloc, int fs_idx,
                                                int N_rb_dl, const lteb200_tx_control *ctl, const int *subfr_num, int n_sf, int N_pdcch_symbs, float *d_grid,
                                                float *d_iq, void *stream_)
{
    if (!ctx || !ctl || !subfr_num || n_sf < 1 || n_alloc < 0 || (n_alloc > 0 && (!d_tb_bits || !allocs)) || fs_idx < 0 || fs_idx > 4 || (!d_grid && !d_iq))
        return LTEB200_ERR_ARGS;
    const int N_id_cell = ctl->N_id_cell, N_ant = ctl->N_ant;
    if (N_id_cell < 0 || N_id_cell > 503 || !(N_ant == 1 || N_ant == 2 || N_ant == 4)) return LTEB200_ERR_ARGS;
    if (n_sf > 65535) { lteb200_set_error(ctx, "at most 65535 subframes per call (split the batch)"); return LTEB200_ERR_ARGS; }
    const int N = 128 << fs_idx, nsc = 12 * N_rb_dl;
    if (nsc >= N) { lteb200_set_error(ctx, "dl_transmit: N_rb_dl does not fit the sample rate"); return LTEB200_ERR_ARGS; }
    for (int i = 0; i < n_sf; i++) if (subfr_num[i] < 0 || subfr_num[i] > 9) return LTEB200_ERR_ARGS;
    cudaStream_t stream = (cudaStream_t)stream_;
    LTE_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> lk(g_tx_mu);
    StreamScope scope;
    float2 *grid = (float2 *)d_grid;
    const size_t ps = (size_t)14 * nsc, sfs = ps * N_ant;
    if (!grid) LTE_CUDA_OK(ctx, scope.alloc(&grid, sizeof(float2) * sfs * (size_t)n_sf, stream));
    int *d_sn = nullptr;
    LTE_CUDA_OK(ctx, scope.alloc(&d_sn, sizeof(int) * n_sf, stream));
    int rc = lteb200_upload_small(ctx, subfr_num, sizeof(int) * n_sf, d_sn, stream);
    if (rc) return rc;
    LTE_CUDA_OK(ctx, cudaMemsetAsync(grid, 0, sizeof(float2) * sfs * (size_t)n_sf, stream));
    float2 *crs = nullptr;
    rc = get_crs_table(ctx, N_id_cell, &crs);
    if (rc) return rc;
    crs_map_kernel<<<dim3(ceil_div(8 * N_rb_dl, 128), n_sf, N_ant), 128, 0, stream>>>(crs, d_sn, N_rb_dl, N_id_cell, nsc, N_ant, grid);
    ctx->launches += 1;
    if (ctl->sync || ctl->pbch || ctl->pdcch) {
        rc = control_tx_launch(ctx, grid, sfs, ps, d_sn, subfr_num, n_sf, N_rb_dl, N_pdcch_symbs - (N_rb_dl <= 10 ? 1 : 0), ctl, allocs, n_alloc, stream);
        if (rc) return rc;
    }
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
        for (int c = 0; c < 3 && rc == LTEB200_OK; c++)
            if (T[a].M[c] % N_ant) { lteb200_set_error(ctx, "dl_transmit: allocation %d is not a whole number of transmit-diversity groups", a); rc = LTEB200_ERR_UNSUPPORTED; }
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
        P.n_al = e - a; P.n_sf = n_sf; P.n_alloc_total = n_alloc; P.tb_stride = tb_stride; P.nsc = nsc; P.K = K; P.N_ant = N_ant;
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
            const float4 g_re = make_float4(ctl->ant_gain_re[0], ctl->ant_gain_re[1], ctl->ant_gain_re[2], ctl->ant_gain_re[3]);
            const float4 g_im = make_float4(ctl->ant_gain_im[0], ctl->ant_gain_im[1], ctl->ant_gain_im[2], ctl->ant_gain_im[3]);
            ofdm_mod_kernel<<<n_sf * 14, threads, sizeof(float2) * 2 * N, stream>>>(grid, N, 10 << fs_idx, 9 << fs_idx, nsc, tw, (float2 *)d_iq, N_ant, g_re, g_im);
            ctx->launches += 1;
        }
    }
    if (rc == LTEB200_OK) LTE_CUDA_OK(ctx, cudaGetLastError());
    return rc;
}

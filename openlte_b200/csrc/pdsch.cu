// PDSCH symbol-level receive for batches of subframes: RE extraction, zero-forcing equaliser, demapping,
// descrambling, then CRC24A / transport-block extraction after turbo decoding.
// Twins of (all in /root/reference/liblte/src/liblte_phy.cc):
//   RE extraction loop of liblte_phy_pdsch_channel_decode   :3745-3802
//   pre_decoder_and_matched_filter_dl, 1 antenna port (ZF)  :7680-7690
//   modulation_demapper                                     :9502-9660   (get_soft_decision :13880-13900)
//   descrambling with generate_prs_c                        :3831-3836, 9669-9704
//   dlsch_channel_decode tail: desegmentation + CRC24A      :12841-12869, 9713-9743
//
// The RE walk depends only on (allocation, cell, subframe class 0 / 5 / other), so it is folded on the host into
// an index list per class; the kernel is then a fused gather -> ZF -> demap -> descramble, one thread per resource
// element, reading grid and channel estimate once (8 B + 8 B) and writing Q_m floats: HBM-bound.
#include "internal.h"
#include <algorithm>
#include "pdsch_dev.cuh"
#include <math.h>

namespace {

struct DemodParams {
    int nsc;          // 12*N_rb_dl
    int Qm, mod;
    int soft_stride;  // floats per subframe in the output
};

__device__ __forceinline__ void demap_store(float xr, float xi, int mod, int Qm, const uint8_t *__restrict__ c, float *__restrict__ o)
{
    int b[6];
    demap_bits(xr, xi, mod, b);
    for (int t = 0; t < Qm; t++) o[t] = __fmul_rn((float)b[t], (float)(1 - 2 * (int)c[t]));  // soft * (1 - 2c), :3835
}

#define MUL(a, b) __fmul_rn((a), (b))

// re_idx: [3][idx_stride] (symbol*nsc + subcarrier) per subframe class; M_cls[3]; cseq: [10][c_stride] scrambling bits per subframe number.
// One thread per group of N_ant consecutive PDSCH resource elements (1 port: zero forcing, liblte_phy.cc:7680-7690;
// 2 / 4 ports: transmit-diversity combining, :7691-7765, followed by layer_demapper_dl's interleave, :7505-7512).
__global__ void __launch_bounds__(128)
pdsch_demod_kernel(const float2 *__restrict__ grid /*[n_sf][16][nsc]*/, const float2 *__restrict__ ce /*[n_sf][N_ant][14][nsc]*/, int N_ant,
                   const int *__restrict__ subfr_num, const int *__restrict__ re_idx, int idx_stride, int3 M_cls,
                   const uint8_t *__restrict__ cseq, int c_stride, DemodParams P, float *__restrict__ soft)
{
    const int sf = blockIdx.y;
    const int sn = subfr_num[sf];
    const int cls = sn == 0 ? 0 : (sn == 5 ? 1 : 2);
    const int M = cls == 0 ? M_cls.x : (cls == 1 ? M_cls.y : M_cls.z);
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = g * N_ant;
    if (r + N_ant > M) return;
    const int *ix = re_idx + cls * idx_stride + r;
    const float2 *gr = grid + (size_t)sf * 16 * P.nsc;
    const float2 *cp = ce + (size_t)sf * N_ant * 14 * P.nsc;
    const size_t pstride = (size_t)14 * P.nsc;
    const uint8_t *c = cseq + (size_t)sn * c_stride + (size_t)r * P.Qm;
    float *o = soft + (size_t)sf * P.soft_stride + (size_t)r * P.Qm;
    if (N_ant == 1) {
        const float2 y = gr[ix[0]], h = cp[ix[0]];
        // zero forcing: x = y * conj(h) / |h|^2, operation order of liblte_phy.cc:7684-7689
        const float nrm = __fadd_rn(MUL(h.x, h.x), MUL(h.y, h.y));
        const float xr = __fdiv_rn(__fadd_rn(MUL(y.x, h.x), MUL(y.y, h.y)), nrm);
        const float xi = __fdiv_rn(__fsub_rn(MUL(y.y, h.x), MUL(y.x, h.y)), nrm);
        demap_store(xr, xi, P.mod, P.Qm, c, o);
    } else if (N_ant == 2) {
        float2 x0, x1;
        sfbc2(gr[ix[0]], gr[ix[1]], cp[ix[0]], cp[pstride + ix[0]], x0, x1);
        demap_store(x0.x, x0.y, P.mod, P.Qm, c, o);
        demap_store(x1.x, x1.y, P.mod, P.Qm, c + P.Qm, o + P.Qm);
    } else {
        const float2 y[4] = {gr[ix[0]], gr[ix[1]], gr[ix[2]], gr[ix[3]]};
        float2 x[4];
        sfbc4(y, cp[ix[0]], cp[pstride + ix[2]], cp[2 * pstride + ix[0]], cp[3 * pstride + ix[2]], x);
        for (int t = 0; t < 4; t++) demap_store(x[t].x, x[t].y, P.mod, P.Qm, c + t * P.Qm, o + t * P.Qm);
    }
}

// One CTA per code block: CRC24A over the transport block inside the decoded block, bit extraction.
// cb_bits: [n_cb][K]; F filler bits precede the TB (zero for every legal TBS).  CRC24A by GF(2) linearity: the remainder of the tbs + 24 bit
// message [a | p] is the XOR of x^(distance from the end) mod g over its set bits (crc_tab[j] = x^j mod g, lteb200_get_crc24a_tab); every thread
// takes a strided share of the bits, so the kernel is a few hundred cycles instead of the 6000-step dependent shift loop it used to be.
__global__ void __launch_bounds__(128)
crc_extract_kernel(const uint8_t *__restrict__ cb_bits, int K, int F, int tbs, int n_sf, int a_idx, int n_alloc,
                   const uint32_t *__restrict__ crc_tab, uint8_t *__restrict__ tb_out, int tb_stride, int32_t *__restrict__ crc_ok, int packed = 0)
{
    __shared__ uint32_t s_rem[4];
    const int sf = blockIdx.x;
    const uint8_t *a = cb_bits + (size_t)sf * K + F;
    uint8_t *o = tb_out + ((size_t)sf * n_alloc + a_idx) * tb_stride;
    const int n = tbs + 24;
    uint32_t rem = 0;
    if (packed) {
        // eight bits per output byte, first bit in the MSB (the launcher admits this format only without filler bits: F = 0, n = K, 8 | K)
        for (int b = threadIdx.x; b < n / 8; b += blockDim.x) {
            const uint2 v = *reinterpret_cast<const uint2 *>(a + 8 * b);
            uint32_t byte = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t bit = ((j < 4 ? v.x : v.y) >> (8 * (j & 3))) & 1u;
                byte |= bit << (7 - j);
                if (bit) rem ^= crc_tab[n - 1 - (8 * b + j)];
            }
            if (8 * b < tbs) o[b] = (uint8_t)byte;
        }
    } else
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t bit = a[i] & 1u;
        if (i < tbs) o[i] = a[i];
        if (bit) rem ^= crc_tab[n - 1 - i];
    }
    for (int o2 = 16; o2 > 0; o2 >>= 1) rem ^= __shfl_xor_sync(0xffffffffu, rem, o2);
    if ((threadIdx.x & 31) == 0) s_rem[threadIdx.x >> 5] = rem;
    __syncthreads();
    if (threadIdx.x == 0) crc_ok[(size_t)sf * n_alloc + a_idx] = ((s_rem[0] ^ s_rem[1] ^ s_rem[2] ^ s_rem[3]) & 0xffffffu) == 0 ? 1 : 0;
}

__global__ void soa_to_grid_kernel(const float *__restrict__ re, const float *__restrict__ im, int rows, int nsc, int src_stride,
                                   float2 *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * nsc) return;
    const int r = i / nsc, k = i - r * nsc;
    out[i] = make_float2(re[r * src_stride + k], im[r * src_stride + k]);
}

}  // namespace

void centre_band(int N_rb_dl, int *first_sc, int *last_sc)
{
    // liblte_phy.cc:3722-3741
    switch (N_rb_dl) {
    case 6: *first_sc = 0; *last_sc = 71; break;
    case 15: *first_sc = 54; *last_sc = 125; break;
    case 25: *first_sc = 114; *last_sc = 185; break;
    case 50: *first_sc = 264; *last_sc = 335; break;
    case 75: *first_sc = 414; *last_sc = 485; break;
    default: *first_sc = 564; *last_sc = 635; break;
    }
}

// RE walk of liblte_phy.cc:3745-3802 for subframe class cls (0: subframe 0, 1: subframe 5, 2: any other)
static void re_walk(int N_rb_dl, int N_id_cell, int N_ant, int cls, int N_pdcch_symbs, const lteb200_alloc &al, std::vector<int> &idx)
{
    int first_sc, last_sc;
    centre_band(N_rb_dl, &first_sc, &last_sc);
    const int nsc = 12 * N_rb_dl;
    idx.clear();
    for (int L = N_pdcch_symbs; L < 14; L++)
        for (int pi = 0; pi < al.N_prb; pi++) {
            const int i = al.prb[L / 7][pi];
            for (int j = 0; j < 12; j++) {
                const int sc = 12 * i + j, l7 = L % 7;
                bool skip;
                if (N_ant == 1) skip = (l7 == 0 && (N_id_cell % 6) == (j % 6)) || (l7 == 4 && ((N_id_cell + 3) % 6) == (j % 6));
                else skip = ((l7 == 0 || l7 == 4) && (N_id_cell % 3) == (j % 3)) || (N_ant == 4 && l7 == 1 && (N_id_cell % 3) == (j % 3));
                const bool centre = sc >= first_sc && sc <= last_sc;
                if (cls == 0 && centre && L >= 7 && L <= 10) skip = true;
                if (cls != 2 && centre && (L == 6 || L == 5)) skip = true;
                if (!skip) idx.push_back(L * nsc + sc);
            }
        }
}

// HOST, no device needed: the index list the symbol-level kernels gather through, for inspection and CPU tests
extern "C" int lteb200_pdsch_re_indices(int N_rb_dl, int N_id_cell, int N_ant, int subfr_num, int N_pdcch_symbs, const lteb200_alloc *alloc,
                                        int32_t *idx_out, int capacity, int *n_out)
{
    if (!alloc || !idx_out || !n_out || N_rb_dl < 6 || N_rb_dl > 110 || N_id_cell < 0 || N_id_cell > 503 || !(N_ant == 1 || N_ant == 2 || N_ant == 4) ||
        subfr_num < 0 || subfr_num > 9 || N_pdcch_symbs < 0 || N_pdcch_symbs > 4 || alloc->N_prb < 1 || alloc->N_prb > N_rb_dl)
        return LTEB200_ERR_ARGS;
    for (int s = 0; s < 2; s++)
        for (int j = 0; j < alloc->N_prb; j++)
            if (alloc->prb[s][j] < 0 || alloc->prb[s][j] >= N_rb_dl) return LTEB200_ERR_ARGS;
    std::vector<int> idx;
    re_walk(N_rb_dl, N_id_cell, N_ant, subfr_num == 0 ? 0 : (subfr_num == 5 ? 1 : 2), N_pdcch_symbs, *alloc, idx);
    *n_out = (int)idx.size();
    if ((int)idx.size() > capacity) return LTEB200_ERR_ARGS;
    std::copy(idx.begin(), idx.end(), idx_out);
    return LTEB200_OK;
}

static int seg_params(int tbs, int *K, int *F)
{
    // liblte_phy_code_block_segmentation, single code block (liblte_phy.cc:9781-9795)
    const int B = tbs + 24;
    if (tbs < 1 || B > 6144) return -1;
    const int steps[4][3] = {{40, 512, 8}, {528, 1024, 16}, {1056, 2048, 32}, {2112, 6144, 64}};
    for (int r = 0; r < 4; r++)
        for (int k = steps[r][0]; k <= steps[r][1]; k += steps[r][2])
            if (k >= B) { *K = k; *F = k - B; return 0; }
    return -1;
}

void prs_c_host(uint32_t c_init, int len, uint8_t *c);
int turbo_mlm_fused_batch(lteb200_ctx *ctx, int K, int n_iter, const FusedParams &fp, cudaStream_t stream);
int rate_unmatch_get_map(lteb200_ctx *ctx, int D, int E, int C, int tx_mode, int N_soft, int M_dl_harq, int chan_type, int rv, RateUnmatchMap *out);
int rate_unmatch_launch_var(lteb200_ctx *ctx, const float *d_e, int e_stride, const int *d_E, int D, const RateUnmatchMap &m, int n_cb,
                            float *d_out, cudaStream_t stream);

// Per-allocation device tables, cached by content (process-wide: the tables depend on the device and the allocation only).  A scan over a
// long capture sees many distinct DCI-derived allocations, false positives included, and each pins up to ~1 MB: the cache is bounded, and
// when it is full the oldest quarter is dropped -- after a device synchronisation, because kernels in flight may still read those tables.
struct CachedAlloc { AllocTables t; unsigned long long stamp; };   // stamp: last use (least recently used entries go first)
static std::map<std::vector<int>, CachedAlloc> g_alloc_tabs;
static unsigned long long g_alloc_clock = 0;
static std::mutex g_alloc_mu;
constexpr size_t ALLOC_CACHE_MAX = 1024;
static void alloc_cache_make_room()
{
    if (g_alloc_tabs.size() < ALLOC_CACHE_MAX) return;
    cudaDeviceSynchronize();
    std::vector<unsigned long long> stamps;
    for (auto &kv : g_alloc_tabs) stamps.push_back(kv.second.stamp);
    std::nth_element(stamps.begin(), stamps.begin() + stamps.size() / 4, stamps.end());
    const unsigned long long cut = stamps[stamps.size() / 4];
    for (auto it = g_alloc_tabs.begin(); it != g_alloc_tabs.end();) {
        if (it->second.stamp < cut) {
            cudaFree(it->second.t.re_idx); cudaFree(it->second.t.cseq); cudaFree(it->second.t.cseqP);
            it = g_alloc_tabs.erase(it);
        } else ++it;
    }
}

int get_alloc_tables(lteb200_ctx *ctx, int N_rb_dl, int N_id_cell, int N_ant, int N_pdcch_symbs, const lteb200_alloc &al, AllocTables *out)
{
    if (al.N_prb < 1 || al.N_prb > N_rb_dl || al.mod_type < 0 || al.mod_type > 3) { lteb200_set_error(ctx, "bad allocation"); return LTEB200_ERR_ARGS; }
    for (int s = 0; s < 2; s++)
        for (int j = 0; j < al.N_prb; j++)
            if (al.prb[s][j] < 0 || al.prb[s][j] >= N_rb_dl) { lteb200_set_error(ctx, "PRB index out of range"); return LTEB200_ERR_ARGS; }
    std::vector<int> key = {ctx->device, N_rb_dl, N_id_cell, N_ant, N_pdcch_symbs, al.mod_type, al.N_prb, al.rnti, al.tbs};
    for (int s = 0; s < 2; s++) for (int j = 0; j < al.N_prb; j++) key.push_back(al.prb[s][j]);
    std::lock_guard<std::mutex> lk(g_alloc_mu);
    auto it = g_alloc_tabs.find(key);
    if (it != g_alloc_tabs.end()) { it->second.stamp = ++g_alloc_clock; *out = it->second.t; return LTEB200_OK; }
    alloc_cache_make_room();
    AllocTables t;
    const int Qms[4] = {1, 2, 4, 6};
    t.Qm = Qms[al.mod_type];
    if (seg_params(al.tbs, &t.K, &t.F)) { lteb200_set_error(ctx, "tbs=%d needs more than one code block: outside the reference's correct envelope (SURVEY H7)", al.tbs); return LTEB200_ERR_UNSUPPORTED; }
    std::vector<int> idx[3];
    for (int c = 0; c < 3; c++) { re_walk(N_rb_dl, N_id_cell, N_ant, c, N_pdcch_symbs, al, idx[c]); t.M[c] = (int)idx[c].size(); }
    t.idx_stride = t.M[2] > 0 ? t.M[2] : 1;   // class 2 (no exclusions) is the largest
    std::vector<int> flat(3 * (size_t)t.idx_stride, 0);
    for (int c = 0; c < 3; c++) std::copy(idx[c].begin(), idx[c].end(), flat.begin() + (size_t)c * t.idx_stride);
    t.c_stride = t.idx_stride * t.Qm;
    std::vector<uint8_t> cs(10 * (size_t)t.c_stride);
    for (int sn = 0; sn < 10; sn++)   // c_init = rnti<<14 | q<<13 | subframe<<9 | N_id_cell, liblte_phy.cc:3831
        prs_c_host(((uint32_t)al.rnti << 14) | ((uint32_t)sn << 9) | (uint32_t)N_id_cell, t.c_stride, cs.data() + (size_t)sn * t.c_stride);
    std::vector<uint8_t> csp(10 * (size_t)t.idx_stride, 0);
    for (int sn = 0; sn < 10; sn++)
        for (int r = 0; r < t.idx_stride; r++)
            for (int q = 0; q < t.Qm; q++) csp[(size_t)sn * t.idx_stride + r] |= (uint8_t)(cs[(size_t)sn * t.c_stride + (size_t)r * t.Qm + q] << q);
    LTE_CUDA_OK(ctx, cudaMalloc(&t.re_idx, sizeof(int) * flat.size()));
    LTE_CUDA_OK(ctx, cudaMalloc(&t.cseq, cs.size()));
    LTE_CUDA_OK(ctx, cudaMalloc(&t.cseqP, csp.size()));
    LTE_CUDA_OK(ctx, cudaMemcpy(t.cseqP, csp.data(), csp.size(), cudaMemcpyHostToDevice));
    LTE_CUDA_OK(ctx, cudaMemcpy(t.re_idx, flat.data(), sizeof(int) * flat.size(), cudaMemcpyHostToDevice));
    LTE_CUDA_OK(ctx, cudaMemcpy(t.cseq, cs.data(), cs.size(), cudaMemcpyHostToDevice));
    g_alloc_tabs[key] = CachedAlloc{t, ++g_alloc_clock};
    *out = t;
    return LTEB200_OK;
}

static int upload_ints(lteb200_ctx *ctx, const int *h, int n, int **d_out, cudaStream_t stream)
{
    // small per-call parameter arrays: dedicated allocations cached by size class would be overkill here
    int *d = nullptr;
    LTE_CUDA_OK(ctx, cudaMallocAsync(&d, sizeof(int) * n, stream));
    if (int rc_ = lteb200_upload_small(ctx, h, sizeof(int) * n, d, stream)) return rc_;
    *d_out = d;
    return LTEB200_OK;
}

int pdsch_demod_launch(lteb200_ctx *ctx, const float *d_grid, const float *d_ce, int N_rb_dl, int N_id_cell, int N_ant,
                       const int *d_subfr_num, int n_sf, int N_pdcch_symbs, const lteb200_alloc &al, float *d_soft, int soft_stride,
                       AllocTables *tabs_out, cudaStream_t stream)
{
    if (!(N_ant == 1 || N_ant == 2 || N_ant == 4)) { lteb200_set_error(ctx, "PDSCH receive: N_ant must be 1, 2 or 4"); return LTEB200_ERR_ARGS; }
    AllocTables t;
    int rc = get_alloc_tables(ctx, N_rb_dl, N_id_cell, N_ant, N_pdcch_symbs, al, &t);
    if (rc) return rc;
    for (int c = 0; c < 3; c++)
        if (t.M[c] % N_ant) {   // the reference's own odd-tail handling for 4 ports is inconsistent (layer strides differ between :7676 and :7486): refuse
            lteb200_set_error(ctx, "allocation has %d resource elements, not a multiple of N_ant=%d", t.M[c], N_ant);
            return LTEB200_ERR_UNSUPPORTED;
        }
    if (soft_stride < t.M[2] * t.Qm) { lteb200_set_error(ctx, "soft_stride too small: need %d", t.M[2] * t.Qm); return LTEB200_ERR_ARGS; }
    DemodParams P;
    P.nsc = 12 * N_rb_dl; P.Qm = t.Qm; P.mod = al.mod_type; P.soft_stride = soft_stride;
    dim3 grid(ceil_div(t.M[2] / N_ant, 128), n_sf);
    pdsch_demod_kernel<<<grid, 128, 0, stream>>>((const float2 *)d_grid, (const float2 *)d_ce, N_ant, d_subfr_num, t.re_idx, t.idx_stride,
                                                 make_int3(t.M[0], t.M[1], t.M[2]), t.cseq, t.c_stride, P, d_soft);
    ctx->launches += 1;
    LTE_CUDA_OK(ctx, cudaGetLastError());
    if (tabs_out) *tabs_out = t;
    return LTEB200_OK;
}

extern "C" int lteb200_pdsch_demod_batch(lteb200_ctx *ctx, const float *d_grid, const float *d_ce, int N_rb_dl, int N_id_cell, int N_ant,
                                         const int *subfr_num, int n_sf, int N_pdcch_symbs, const lteb200_alloc *alloc, float *d_soft,
                                         int soft_stride, int *n_bits_out, void *stream)
{
    if (!ctx || !d_grid || !d_ce || !subfr_num || !alloc || !d_soft || n_sf < 1) return LTEB200_ERR_ARGS;
    if (n_sf > 65535) { lteb200_set_error(ctx, "at most 65535 subframes per call (split the batch)"); return LTEB200_ERR_ARGS; }
    LTE_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    for (int i = 0; i < n_sf; i++) if (subfr_num[i] < 0 || subfr_num[i] > 9) return LTEB200_ERR_ARGS;
    int *d_sn = nullptr;
    int rc = upload_ints(ctx, subfr_num, n_sf, &d_sn, (cudaStream_t)stream);
    if (rc) return rc;
    AllocTables t;
    rc = pdsch_demod_launch(ctx, d_grid, d_ce, N_rb_dl, N_id_cell, N_ant, d_sn, n_sf, N_pdcch_symbs, *alloc, d_soft, soft_stride, &t, (cudaStream_t)stream);
    cudaFreeAsync(d_sn, (cudaStream_t)stream);
    if (rc) return rc;
    if (n_bits_out) for (int i = 0; i < n_sf; i++) n_bits_out[i] = t.M[subfr_num[i] == 0 ? 0 : (subfr_num[i] == 5 ? 1 : 2)] * t.Qm;
    return LTEB200_OK;
}

// Full PDSCH receive of a batch, all allocations.  Scratch layout per allocation a: soft [n_sf][G_stride], d [n_sf][3(K+4)], cb [n_sf][K].
// Per-allocation part of the receive parameters for allocations [a, e) of equal block size (at most 8 per group)
static int fill_group(lteb200_ctx *ctx, const std::vector<AllocTables> &T, const lteb200_alloc *allocs, int n_alloc, int a, int *e_out, FusedParams &fp)
{
    int e = a;
    while (e < n_alloc && e - a < 8 && T[e].K == T[a].K) {
        RateUnmatchMap m;
        int rc = rate_unmatch_get_map(ctx, T[e].K + 4, 0, 1, allocs[e].tx_mode, 250368, 8, 0, allocs[e].rv_idx, &m);   // M_dl_harq 8, cat-1 N_soft: liblte_phy.cc:3843-3844
        if (rc) return rc;
        FusedAlloc &A = fp.al[e - a];
        A.re_idx = T[e].re_idx; A.cseq = T[e].cseq; A.cseqP = T[e].cseqP; A.first_k = m.first_k; A.first_kT = m.first_kT; A.inv_kT = m.inv_kT; A.idx_stride = T[e].idx_stride; A.c_stride = T[e].c_stride;
        for (int c = 0; c < 3; c++) A.M[c] = T[e].M[c];
        A.Qm = T[e].Qm; A.mod = allocs[e].mod_type; A.period = m.n_nonnull; A.F = T[e].F; A.tbs = allocs[e].tbs; A.a_idx = e;
        e++;
    }
    fp.n_al = e - a;
    *e_out = e;
    return LTEB200_OK;
}

// The two halves of the max-log-MAP receive path as separate entry points (all allocations of one call must share the block size K, K mod 32 == 0,
// at most 8 allocations): the symbol level writes one quantised image per (allocation, subframe), index a * n_sf + subframe, lteb200_q8_stride(K) bytes
// each; the decoder turns images into transport blocks + CRC verdicts.
extern "C" size_t lteb200_q8_stride(int K) { return mlm2_handles(K) ? mlm2_q8_stride(K) : 0; }

static int q8_common(lteb200_ctx *ctx, int N_rb_dl, int N_id_cell, int N_ant, int N_pdcch_symbs, const lteb200_alloc *allocs, int n_alloc, std::vector<AllocTables> &T,
                     FusedParams &fp)
{
    if (n_alloc < 1 || n_alloc > 8) { lteb200_set_error(ctx, "q8 path: 1..8 allocations per call"); return LTEB200_ERR_ARGS; }
    T.resize(n_alloc);
    for (int a = 0; a < n_alloc; a++) {
        int rc = get_alloc_tables(ctx, N_rb_dl, N_id_cell, N_ant, N_pdcch_symbs, allocs[a], &T[a]);
        if (rc) return rc;
        if (T[a].K != T[0].K || !mlm2_handles(T[a].K)) { lteb200_set_error(ctx, "q8 path: allocations must share one block size that is a multiple of 32"); return LTEB200_ERR_UNSUPPORTED; }
        for (int c = 0; c < 3; c++) if (T[a].M[c] % N_ant) { lteb200_set_error(ctx, "allocation is not a whole number of transmit-diversity groups"); return LTEB200_ERR_UNSUPPORTED; }
    }
    memset(&fp, 0, sizeof(fp));
    int e = 0;
    int rc = fill_group(ctx, T, allocs, n_alloc, 0, &e, fp);
    if (rc) return rc;
    if (e != n_alloc) return LTEB200_ERR_UNSUPPORTED;
    fp.N_ant = N_ant; fp.nsc = 12 * N_rb_dl; fp.n_alloc_total = n_alloc;
    fp.demap = ctx->demap; fp.vshift = N_id_cell % 6; fp.tb_packed = ctx->tb_packed;
    return LTEB200_OK;
}

extern "C" int lteb200_pdsch_llr_q8_batch(lteb200_ctx *ctx, const float *d_grid, const float *d_ce, int N_rb_dl, int N_id_cell, int N_ant, const int *subfr_num,
                                          int n_sf, int N_pdcch_symbs, const lteb200_alloc *allocs, int n_alloc, int8_t *d_q8, void *stream_)
{
    if (!ctx || !d_grid || !d_ce || !subfr_num || !allocs || !d_q8 || n_sf < 1) return LTEB200_ERR_ARGS;
    cudaStream_t stream = (cudaStream_t)stream_;
    LTE_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    for (int i = 0; i < n_sf; i++) if (subfr_num[i] < 0 || subfr_num[i] > 9) { lteb200_set_error(ctx, "subframe number out of range"); return LTEB200_ERR_ARGS; }
    std::vector<AllocTables> T;
    FusedParams fp;
    int rc = q8_common(ctx, N_rb_dl, N_id_cell, N_ant, N_pdcch_symbs, allocs, n_alloc, T, fp);
    if (rc) return rc;
    StreamScope scope;
    int *d_sn = nullptr;
    LTE_CUDA_OK(ctx, scope.alloc(&d_sn, sizeof(int) * n_sf, stream));
    if ((rc = lteb200_upload_small(ctx, subfr_num, sizeof(int) * n_sf, d_sn, stream))) return rc;
    fp.grid = (const float2 *)d_grid; fp.ce = (const float2 *)d_ce; fp.subfr_num = d_sn; fp.n_sf = n_sf;
    int E_max = 0;
    for (int a = 0; a < n_alloc; a++) E_max = std::max(E_max, T[a].M[2] * T[a].Qm);
    return pdsch_llr_q8_launch(ctx, fp, T[0].K, E_max, d_q8, stream);
}

extern "C" int lteb200_turbo_decode_q8_batch(lteb200_ctx *ctx, const int8_t *d_q8, int K, int n_cb, int n_iter, const lteb200_alloc *allocs, int n_alloc, int n_sf,
                                             uint8_t *d_tb_bits, int tb_stride, int32_t *d_crc_ok, void *stream_)
{
    if (!ctx || !d_q8 || !allocs || !d_tb_bits || !d_crc_ok || n_sf < 1 || n_alloc < 1 || n_alloc > 8 || n_cb != n_alloc * n_sf) return LTEB200_ERR_ARGS;
    LTE_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    if (!mlm2_handles(K)) { lteb200_set_error(ctx, "q8 decoder: K=%d is not a multiple of 32", K); return LTEB200_ERR_UNSUPPORTED; }
    FusedParams fp;
    memset(&fp, 0, sizeof(fp));
    for (int a = 0; a < n_alloc; a++) {
        int Ka = 0, Fa = 0;
        if (seg_params(allocs[a].tbs, &Ka, &Fa) || Ka != K) { lteb200_set_error(ctx, "q8 decoder: transport block size %d does not map onto K=%d", allocs[a].tbs, K); return LTEB200_ERR_ARGS; }
        if (tb_stride < (ctx->tb_packed ? (allocs[a].tbs + 7) / 8 : allocs[a].tbs)) { lteb200_set_error(ctx, "tb_stride < tbs"); return LTEB200_ERR_ARGS; }
        if (ctx->tb_packed && Fa != 0) { lteb200_set_error(ctx, "packed transport-block output needs tbs + 24 = K"); return LTEB200_ERR_UNSUPPORTED; }
        fp.al[a].F = Fa; fp.al[a].tbs = allocs[a].tbs; fp.al[a].a_idx = a;
    }
    const uint32_t *crc_tab = nullptr;
    int rc = lteb200_get_crc24a_tab(ctx, &crc_tab);
    if (rc) return rc;
    fp.crc_tab = crc_tab; fp.tb_out = d_tb_bits; fp.crc_ok = d_crc_ok; fp.n_sf = n_sf; fp.n_alloc_total = n_alloc; fp.tb_stride = tb_stride; fp.n_al = n_alloc;
    fp.tb_packed = ctx->tb_packed;
    return turbo_mlm2_launch(ctx, nullptr, d_q8, K, n_cb, n_iter, nullptr, &fp, (cudaStream_t)stream_);
}

extern "C" int lteb200_pdsch_decode_batch(lteb200_ctx *ctx, const float *d_grid, const float *d_ce, int N_rb_dl, int N_id_cell, int N_ant,
                                          const int *subfr_num, int n_sf, int N_pdcch_symbs, const lteb200_alloc *allocs, int n_alloc,
                                          int turbo_mode, int n_iter, uint8_t *d_tb_bits, int tb_stride, int32_t *d_crc_ok, void *stream_)
{
    return pdsch_decode_batch_dev(ctx, d_grid, d_ce, N_rb_dl, N_id_cell, N_ant, subfr_num, nullptr, n_sf, N_pdcch_symbs, allocs, n_alloc, turbo_mode, n_iter,
                                  d_tb_bits, tb_stride, d_crc_ok, (cudaStream_t)stream_);
}

// d_subfr_dev: the same subframe numbers already resident on the device (the host pipeline uploads them once for all its chunks), or null.
int pdsch_decode_batch_dev(lteb200_ctx *ctx, const float *d_grid, const float *d_ce, int N_rb_dl, int N_id_cell, int N_ant, const int *subfr_num,
                           const int *d_subfr_dev, int n_sf, int N_pdcch_symbs, const lteb200_alloc *allocs, int n_alloc, int turbo_mode, int n_iter,
                           uint8_t *d_tb_bits, int tb_stride, int32_t *d_crc_ok, cudaStream_t stream_)
{
    if (!ctx || !d_grid || !d_ce || !subfr_num || !allocs || !d_tb_bits || !d_crc_ok || n_sf < 1 || n_alloc < 1) return LTEB200_ERR_ARGS;
    if (n_sf > 65535) { lteb200_set_error(ctx, "at most 65535 subframes per call (split the batch)"); return LTEB200_ERR_ARGS; }
    cudaStream_t stream = (cudaStream_t)stream_;
    LTE_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    for (int i = 0; i < n_sf; i++) if (subfr_num[i] < 0 || subfr_num[i] > 9) { lteb200_set_error(ctx, "subframe number out of range"); return LTEB200_ERR_ARGS; }
    std::vector<AllocTables> T(n_alloc);
    int rc;
    for (int a = 0; a < n_alloc; a++) {
        rc = get_alloc_tables(ctx, N_rb_dl, N_id_cell, N_ant, N_pdcch_symbs, allocs[a], &T[a]);
        if (rc) return rc;
        if (tb_stride < (ctx->tb_packed ? (allocs[a].tbs + 7) / 8 : allocs[a].tbs)) { lteb200_set_error(ctx, "tb_stride < tbs"); return LTEB200_ERR_ARGS; }
        if (ctx->tb_packed && (T[a].F != 0 || (turbo_mode == LTEB200_TURBO_MAXLOGMAP && !mlm2_handles(T[a].K)))) {
            lteb200_set_error(ctx, "packed transport-block output: no filler bits (tbs + 24 = K), and block sizes that are multiples of 32 with the max-log-MAP decoder");
            return LTEB200_ERR_UNSUPPORTED;
        }
    }
    // ---- fused path: one kernel per group of allocations with equal K does demod + unmatch + decode + CRC per transport block
    bool fusable = turbo_mode == LTEB200_TURBO_MAXLOGMAP && !getenv("LTEB200_NO_FUSE");
    for (int a = 0; a < n_alloc && fusable; a++) {
        // transmit diversity works on whole groups of N_ant resource elements: ragged tails stay on the staged path (which drops the odd
        // element of a 2-port allocation like the reference and refuses the 4-port case)
        for (int c = 0; c < 3; c++) if (T[a].M[c] % N_ant != 0) fusable = false;
        // the fused kernel stages an allocation's soft bits in the shared memory of its forward-metric store: 64 floats per window column
        const int NT = mlm_size_class(T[a].K);
        if (!mlm2_handles(T[a].K) && T[a].M[2] * T[a].Qm > 64 * NT) fusable = false;    // (the pair kernel's symbol-level kernel stages any single-block allocation)
    }
    if (fusable) {
        int *d_sn = const_cast<int *>(d_subfr_dev);
        if (!d_sn) {
            rc = upload_ints(ctx, subfr_num, n_sf, &d_sn, stream);
            if (rc) return rc;
        }
        const uint32_t *crc_tab = nullptr;
        rc = lteb200_get_crc24a_tab(ctx, &crc_tab);
        for (int a = 0; a < n_alloc && rc == LTEB200_OK;) {
            FusedParams fp;
            memset(&fp, 0, sizeof(fp));
            fp.grid = (const float2 *)d_grid; fp.ce = (const float2 *)d_ce; fp.subfr_num = d_sn; fp.crc_tab = crc_tab;
            fp.N_ant = N_ant;
            fp.demap = ctx->demap; fp.vshift = N_id_cell % 6; fp.tb_packed = ctx->tb_packed;
            fp.tb_out = d_tb_bits; fp.crc_ok = d_crc_ok; fp.nsc = 12 * N_rb_dl; fp.n_sf = n_sf; fp.n_alloc_total = n_alloc; fp.tb_stride = tb_stride;
            int e = a;
            rc = fill_group(ctx, T, allocs, n_alloc, a, &e, fp);
            if (rc) break;
            if (mlm2_handles(T[a].K)) {
                // symbol level -> quantised LLR images (18 KB per K = 6144 block), then the pair-lane decoder with CRC + transport-block output
                int E_max = 0;
                for (int q = a; q < e; q++) E_max = std::max(E_max, T[q].M[2] * T[q].Qm);
                int8_t *d_q8 = nullptr;
                LTE_CUDA_OK(ctx, cudaMallocAsync(&d_q8, mlm2_q8_stride(T[a].K) * (size_t)fp.n_al * n_sf, stream));
                rc = pdsch_llr_q8_launch(ctx, fp, T[a].K, E_max, d_q8, stream);
                if (rc == LTEB200_OK) rc = turbo_mlm2_launch(ctx, nullptr, d_q8, T[a].K, fp.n_al * n_sf, n_iter, nullptr, &fp, stream);
                cudaFreeAsync(d_q8, stream);
            } else
                rc = turbo_mlm_fused_batch(ctx, T[a].K, n_iter, fp, stream);
            a = e;
        }
        if (!d_subfr_dev) cudaFreeAsync(d_sn, stream);
        return rc;
    }
    // ---- reference-exact decoder on quantised images: the symbol-level kernel of the max-log-MAP path (reference demapper, quantisation maximum
    // over the 3 K code-block values like turbo_decode's own) feeds turbo_ref's front end directly -- no float soft bits, no separate rate-unmatch
    // pass.  Bit-identical to the staged kernels below, which remain for allocations that do not fit (and for LTEB200_NO_FUSE=1).
    bool q8_ref = turbo_mode == LTEB200_TURBO_REF_EXACT && !getenv("LTEB200_NO_FUSE");
    for (int a = 0; a < n_alloc && q8_ref; a++) {
        for (int c = 0; c < 3; c++) if (T[a].M[c] % N_ant != 0) q8_ref = false;
        if ((size_t)T[a].M[2] * T[a].Qm + mlm2_q8_stride(T[a].K) > 190 * 1024) q8_ref = false;
    }
    if (q8_ref) {
        int *d_sn = const_cast<int *>(d_subfr_dev);
        if (!d_sn) {
            rc = upload_ints(ctx, subfr_num, n_sf, &d_sn, stream);
            if (rc) return rc;
        }
        const uint32_t *crc_tab = nullptr;
        rc = lteb200_get_crc24a_tab(ctx, &crc_tab);
        for (int a = 0; a < n_alloc && rc == LTEB200_OK;) {
            FusedParams fp;
            memset(&fp, 0, sizeof(fp));
            fp.grid = (const float2 *)d_grid; fp.ce = (const float2 *)d_ce; fp.subfr_num = d_sn; fp.crc_tab = crc_tab;
            fp.N_ant = N_ant;
            fp.demap = LTEB200_DEMAP_REFERENCE; fp.vshift = N_id_cell % 6; fp.tb_packed = 0; fp.max_no_tail = 1;
            fp.tb_out = d_tb_bits; fp.crc_ok = d_crc_ok; fp.nsc = 12 * N_rb_dl; fp.n_sf = n_sf; fp.n_alloc_total = n_alloc; fp.tb_stride = tb_stride;
            int e = a;
            rc = fill_group(ctx, T, allocs, n_alloc, a, &e, fp);
            if (rc) break;
            const int K = T[a].K;
            int E_max = 0;
            for (int q = a; q < e; q++) E_max = std::max(E_max, T[q].M[2] * T[q].Qm);
            int8_t *d_q8 = nullptr;
            uint8_t *d_cb = nullptr;
            LTE_CUDA_OK(ctx, cudaMallocAsync(&d_q8, mlm2_q8_stride(K) * (size_t)fp.n_al * n_sf, stream));
            LTE_CUDA_OK(ctx, cudaMallocAsync(&d_cb, (size_t)K * fp.n_al * n_sf, stream));
            rc = pdsch_llr_q8_launch(ctx, fp, K, E_max, d_q8, stream);
            if (rc == LTEB200_OK) rc = turbo_ref_decode_q8_batch(ctx, d_q8, K, fp.n_al * n_sf, d_cb, stream);
            for (int q = a; q < e && rc == LTEB200_OK; q++) {
                crc_extract_kernel<<<n_sf, 128, 0, stream>>>(d_cb + (size_t)(q - a) * n_sf * K, K, T[q].F, allocs[q].tbs, n_sf, q, n_alloc, crc_tab, d_tb_bits, tb_stride, d_crc_ok,
                                                             ctx->tb_packed);
                ctx->launches += 1;
            }
            cudaFreeAsync(d_q8, stream);
            cudaFreeAsync(d_cb, stream);
            a = e;
        }
        if (!d_subfr_dev) cudaFreeAsync(d_sn, stream);
        if (rc) return rc;
        LTE_CUDA_OK(ctx, cudaGetLastError());
        return LTEB200_OK;
    }
    if (ctx->tb_packed) { lteb200_set_error(ctx, "packed transport-block output is not available on the staged path (allocation too wide for the symbol-level kernel, ragged multi-port tail, or LTEB200_NO_FUSE)"); return LTEB200_ERR_UNSUPPORTED; }
    // per-call device parameter arrays: subframe numbers and per-(alloc, subframe) E
    std::vector<int> hpar((size_t)n_sf * (1 + n_alloc));
    for (int i = 0; i < n_sf; i++) hpar[i] = subfr_num[i];
    for (int a = 0; a < n_alloc; a++)
        for (int i = 0; i < n_sf; i++)
            hpar[(size_t)n_sf * (1 + a) + i] = T[a].M[subfr_num[i] == 0 ? 0 : (subfr_num[i] == 5 ? 1 : 2)] * T[a].Qm;
    int *d_par = nullptr;
    rc = upload_ints(ctx, hpar.data(), (int)hpar.size(), &d_par, stream);
    if (rc) return rc;
    // scratch for the symbol-level stages lives in a stream-ordered allocation (the arena belongs to the turbo decoder)
    size_t soft_floats = 0, d_floats = 0, cb_bytes = 0;
    std::vector<size_t> soft_off(n_alloc), d_off(n_alloc), cb_off(n_alloc);
    for (int a = 0; a < n_alloc; a++) {
        soft_off[a] = soft_floats; soft_floats += (size_t)n_sf * T[a].c_stride;
        d_off[a] = d_floats;       d_floats += (size_t)n_sf * 3 * (T[a].K + 4);
        cb_off[a] = cb_bytes;      cb_bytes += (size_t)n_sf * T[a].K;
    }
    float *d_soft = nullptr, *d_d = nullptr;
    uint8_t *d_cb = nullptr;
    LTE_CUDA_OK(ctx, cudaMallocAsync(&d_soft, sizeof(float) * soft_floats, stream));
    LTE_CUDA_OK(ctx, cudaMallocAsync(&d_d, sizeof(float) * d_floats, stream));
    LTE_CUDA_OK(ctx, cudaMallocAsync(&d_cb, cb_bytes, stream));
    for (int a = 0; a < n_alloc && rc == LTEB200_OK; a++) {
        rc = pdsch_demod_launch(ctx, d_grid, d_ce, N_rb_dl, N_id_cell, N_ant, d_par, n_sf, N_pdcch_symbs, allocs[a], d_soft + soft_off[a],
                                T[a].c_stride, nullptr, stream);
        if (rc) break;
        RateUnmatchMap m;
        rc = rate_unmatch_get_map(ctx, T[a].K + 4, 0, 1, allocs[a].tx_mode, 250368, 8, 0, allocs[a].rv_idx, &m);  // M_dl_harq 8, cat-1 N_soft: liblte_phy.cc:3843-3844
        if (rc) break;
        rc = rate_unmatch_launch_var(ctx, d_soft + soft_off[a], T[a].c_stride, d_par + (size_t)n_sf * (1 + a), T[a].K + 4, m, n_sf,
                                     d_d + d_off[a], stream);
    }
    // turbo decode: consecutive allocations with the same K form one batch
    for (int a = 0; a < n_alloc && rc == LTEB200_OK;) {
        int e = a + 1;
        while (e < n_alloc && T[e].K == T[a].K) e++;
        rc = lteb200_turbo_decode_batch(ctx, d_d + d_off[a], T[a].K, (e - a) * n_sf, turbo_mode, n_iter, d_cb + cb_off[a], stream);
        a = e;
    }
    const uint32_t *crc_tab_s = nullptr;
    if (rc == LTEB200_OK) rc = lteb200_get_crc24a_tab(ctx, &crc_tab_s);
    for (int a = 0; a < n_alloc && rc == LTEB200_OK; a++) {
        crc_extract_kernel<<<n_sf, 128, 0, stream>>>(d_cb + cb_off[a], T[a].K, T[a].F, allocs[a].tbs, n_sf, a, n_alloc, crc_tab_s, d_tb_bits, tb_stride, d_crc_ok);
        ctx->launches += 1;
    }
    cudaFreeAsync(d_soft, stream);
    cudaFreeAsync(d_d, stream);
    cudaFreeAsync(d_cb, stream);
    cudaFreeAsync(d_par, stream);
    if (rc) return rc;
    LTE_CUDA_OK(ctx, cudaGetLastError());
    return LTEB200_OK;
}

int soa_to_grid(lteb200_ctx *ctx, const float *d_re, const float *d_im, int rows, int nsc, int src_stride, float *d_out, cudaStream_t stream)
{
    soa_to_grid_kernel<<<ceil_div(rows * nsc, 256), 256, 0, stream>>>(d_re, d_im, rows, nsc, src_stride, (float2 *)d_out);
    ctx->launches += 1;
    LTE_CUDA_OK(ctx, cudaGetLastError());
    return LTEB200_OK;
}

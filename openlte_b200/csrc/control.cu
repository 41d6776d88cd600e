// Control-channel receive for batches of subframes: PBCH (MIB), PCFICH (CFI) and the PDCCH common search space (DCI 1A / 1C).
// Twins of (all in /root/reference/liblte/src/liblte_phy.cc):
//   liblte_phy_bch_channel_decode :3968-4100        bch_channel_decode   :12581-12655
//   liblte_phy_pdcch_channel_decode :4519-5134      pcfich_channel_demap :7887-7962     cfi_channel_decode :13648-13708
//   phich_channel_demap (REG positions only) :8224-8287                  dci_channel_decode :12952-13035
//   rate_unmatch_conv :11597-11752                  viterbi_decode (K = 7, rate 1/3) :10161-10331
//   dci_1a_unpack :13273-13372, dci_1c_unpack :13400-13605 (host side, lteb200_pdcch_unpack)
//
// Mapping to the GPU: one CTA per subframe, one warp per blind-decode hypothesis (PBCH: 3 antenna configurations x 4 positions
// in the 40 ms period; PDCCH: 6 candidates x 2 DCI sizes).  Everything that depends only on the cell configuration -- the REG
// walk, cyclic shift, REG de-interleaver, CCE assembly, scrambling sequences, the circular-buffer order of the convolutional
// rate matcher -- is folded into small tables built once on the host and cached in HBM.
//
// The reference's Viterbi decoder works on floats, but every value it ever holds is an integer below 2^24 (soft bits are
// int8-valued, sums of a few of them, times small counts), so the int32 arithmetic below reproduces it bit for bit,
// including its quirks: all states start at metric 0, the survivor is chosen on (Hamming count + accumulated metric) while the
// accumulated value is weighted by sum|d|, and the traceback re-decides from the stored path metrics.
//
// PCFICH/PDCCH are restricted to N_ant == 1: for 2 and 4 ports the reference passes a port stride of 576 for arrays whose rows
// are 288 long (liblte_phy.h:405 against liblte_phy.cc:4893 and :7935), so it combines ports 1-3 from other ports' rows or from
// memory past the array; there is no defined behaviour to reproduce.  PBCH supports 1, 2 and 4 ports.
#include "internal.h"
#include "pdsch_dev.cuh"
#include <math.h>

void prs_c_host(uint32_t c_init, int len, uint8_t *c);
int soa_to_grid(lteb200_ctx *ctx, const float *d_re, const float *d_im, int rows, int nsc, int src_stride, float *d_out, cudaStream_t stream);

namespace {

constexpr int N_HYP = 12;            // hypotheses (warps) per subframe in both kernels
constexpr int CC_MAX_N = 44;         // longest convolutional block on this path: DCI 1A at 20 MHz, 28 + 16 bits
constexpr int PM_STRIDE = (CC_MAX_N + 1) * 64;

const uint8_t kColPermCC[32] = {1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31,
                                0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30};   // 36.212 table 5.1.4-2

// ------------------------------------------------------------------------------------------------ device: convolutional code
__device__ __forceinline__ unsigned cc_out3(int in, int prev)
{
    const unsigned reg = ((unsigned)in << 6) | (unsigned)prev;
    return (__popc(reg & 0133u) & 1u) | ((__popc(reg & 0171u) & 1u) << 1) | ((__popc(reg & 0165u) & 1u) << 2);
}

// One warp decodes one block.  d: 3*N integer soft bits (positive = bit 0); pm: (N+1)*64 ints of scratch; bits: N decoded bits.
__device__ void cc_viterbi_warp(const int *__restrict__ d, int N, int *__restrict__ pm, uint8_t *__restrict__ bits, int lane)
{
    const unsigned o00 = cc_out3(0, 2 * lane), o01 = cc_out3(0, 2 * lane + 1), o10 = cc_out3(1, 2 * lane), o11 = cc_out3(1, 2 * lane + 1);
    pm[lane] = 0; pm[lane + 32] = 0;
    __syncwarp();
    for (int t = 0; t < N; t++) {
        const int d0 = d[3 * t], d1 = d[3 * t + 1], d2 = d[3 * t + 2];
        const unsigned hard = (unsigned)(d0 < 0) | ((unsigned)(d1 < 0) << 1) | ((unsigned)(d2 < 0) << 2);
        const int w = abs(d0) + abs(d1) + abs(d2);
        const int p0 = pm[t * 64 + 2 * lane], p1 = pm[t * 64 + 2 * lane + 1];
        int m0 = __popc(o00 ^ hard), m1 = __popc(o01 ^ hard);
        pm[(t + 1) * 64 + lane] = (m0 + p0 > m1 + p1) ? p1 + w * m1 : p0 + w * m0;       // state `lane`: input bit 0
        m0 = __popc(o10 ^ hard); m1 = __popc(o11 ^ hard);
        pm[(t + 1) * 64 + lane + 32] = (m0 + p0 > m1 + p1) ? p1 + w * m1 : p0 + w * m0;  // state `lane + 32`: input bit 1
        __syncwarp();
    }
    // final state: first index holding the strict minimum
    int best = pm[N * 64 + lane], idx = lane;
    { const int v = pm[N * 64 + lane + 32]; if (v < best) { best = v; idx = lane + 32; } }
    for (int o = 16; o > 0; o >>= 1) {
        const int ob = __shfl_xor_sync(0xffffffffu, best, o), oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ob < best || (ob == best && oi < idx)) { best = ob; idx = oi; }
    }
    if (lane == 0) {
        int later = idx;
        for (int t = N - 1; t >= 0; t--) {
            const int s0 = (later << 1) & 63, s1 = s0 + 1;
            const int earlier = pm[t * 64 + s0] > pm[t * 64 + s1] ? s1 : s0;
            bits[t] = later < earlier ? 0 : (later > earlier ? 1 : (later == 0 ? 0 : 1));
            later = earlier;
        }
    }
    __syncwarp();
}

// CRC16 (x^16 + x^12 + x^5 + 1) of n bits, serial; n <= 28 here
__device__ __forceinline__ unsigned crc16_bits(const uint8_t *b, int n)
{
    unsigned rem = 0;
    for (int i = 0; i < n + 16; i++) {
        rem = (rem << 1) | (i < n ? (unsigned)(b[i] & 1) : 0u);
        if (rem & 0x10000u) rem ^= 0x11021u;
    }
    return rem & 0xffffu;
}

__device__ __forceinline__ void qpsk_soft(float xr, float xi, int &b0, int &b1)
{
    int b[6];
    demap_bits(xr, xi, LTEB200_MOD_QPSK, b);
    b0 = b[0]; b1 = b[1];
}

// ------------------------------------------------------------------------------------------------ PBCH
struct BchParams { int nsc, k0, cell_mod3; };

// grid: [n][16][nsc] float2, ce: [n][4][14][nsc] float2, c: [1920] scrambling bits, order: [120] circular-buffer order for N_c = 40
__global__ void __launch_bounds__(32 * N_HYP)
bch_decode_kernel(const float2 *__restrict__ grid, const float2 *__restrict__ ce, const uint8_t *__restrict__ c, const int16_t *__restrict__ order,
                  BchParams P, lteb200_bch_result *__restrict__ out)
{
    extern __shared__ int smem_i[];
    int *pm = smem_i;                                   // [N_HYP][41*64]
    int *dd = pm + N_HYP * 41 * 64;                     // [N_HYP][120]
    float2 *y = (float2 *)(dd + N_HYP * 120);           // [240]
    float2 *h = y + 240;                                // [4][240]
    int16_t *soft = (int16_t *)(h + 4 * 240);           // [3][480]
    uint8_t *bits = (uint8_t *)(soft + 3 * 480);        // [N_HYP][40]
    __shared__ int ok[N_HYP];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float2 *g = grid + (size_t)blockIdx.x * 16 * P.nsc;
    const float2 *cp = ce + (size_t)blockIdx.x * 4 * 14 * P.nsc;
    // PBCH resource elements, liblte_phy.cc:4001-4031: symbols 7, 8 (CRS positions of all four ports skipped) then 9, 10
    for (int t = tid; t < 4 * 72; t += blockDim.x) {
        const int s = t / 72, i = t % 72, sym = 7 + s;
        int dst;
        if (s < 2) {
            if (P.cell_mod3 == i % 3) continue;
            dst = s * 48 + (i - i / 3 - (i % 3 > P.cell_mod3 ? 1 : 0));
        } else dst = 96 + (s - 2) * 72 + i;
        y[dst] = g[sym * P.nsc + P.k0 + i];
        for (int p = 0; p < 4; p++) h[p * 240 + dst] = cp[((size_t)p * 14 + sym) * P.nsc + P.k0 + i];
    }
    __syncthreads();
    // equalise + demap under the three antenna hypotheses: 240 + 120 + 60 independent groups
    for (int t = tid; t < 420; t += blockDim.x) {
        int b0, b1;
        if (t < 240) {
            float xr, xi;
            zf_1port(y[t], h[t], xr, xi);
            qpsk_soft(xr, xi, b0, b1);
            soft[2 * t] = (int16_t)b0; soft[2 * t + 1] = (int16_t)b1;
        } else if (t < 360) {
            const int i = 2 * (t - 240);
            float2 x0, x1;
            sfbc2(y[i], y[i + 1], h[i], h[240 + i], x0, x1);
            qpsk_soft(x0.x, x0.y, b0, b1); soft[480 + 2 * i] = (int16_t)b0; soft[480 + 2 * i + 1] = (int16_t)b1;
            qpsk_soft(x1.x, x1.y, b0, b1); soft[480 + 2 * i + 2] = (int16_t)b0; soft[480 + 2 * i + 3] = (int16_t)b1;
        } else {
            const int i = 4 * (t - 360);
            float2 x[4];
            sfbc4(&y[i], h[i], h[240 + i + 2], h[480 + i], h[720 + i + 2], x);
            for (int q = 0; q < 4; q++) { qpsk_soft(x[q].x, x[q].y, b0, b1); soft[960 + 2 * (i + q)] = (int16_t)b0; soft[960 + 2 * (i + q) + 1] = (int16_t)b1; }
        }
    }
    __syncthreads();
    // hypothesis (antenna configuration, quarter of the 40 ms period): descramble, rate unmatch, Viterbi, CRC under the antenna mask
    {
        const int hyp = warp >> 2, off = warp & 3;
        int *d = dd + warp * 120;
        const int16_t *sf = soft + hyp * 480;
        // 1920 received positions, 480 of them filled: every mother-code bit collects its four copies inside the filled quarter
        // (rate_unmatch_conv with RX_NULL elsewhere, liblte_phy.cc:4073-4080, 11687-11703)
        for (int m = lane; m < 120; m += 32) {
            int acc = 0;
            for (int r = 0; r < 4; r++) {
                const int j = r * 120 + m;
                acc += (int)sf[j] * (1 - 2 * (int)c[off * 480 + j]);
            }
            d[order[m]] = acc;
        }
        __syncwarp();
        cc_viterbi_warp(d, 40, pm + warp * 41 * 64, bits + warp * 40, lane);
        if (lane == 0) {
            const uint8_t *b = bits + warp * 40;
            unsigned rx = 0;
            for (int i = 0; i < 16; i++) rx = (rx << 1) | b[24 + i];
            const unsigned mask = hyp == 0 ? 0x0000u : (hyp == 1 ? 0xFFFFu : 0x5555u);   // liblte_phy.cc:12596-12598
            ok[warp] = (crc16_bits(b, 24) ^ mask) == rx;
        }
    }
    __syncthreads();
    if (tid == 0) {
        lteb200_bch_result r;
        r.ok = 0; r.N_ant = 0; r.offset = 0;
        for (int i = 0; i < 24; i++) r.bits[i] = 0;
        for (int w = 0; w < N_HYP; w++)                 // the reference's search order: 1, 2, 4 ports, offsets 0..3, first hit wins
            if (ok[w]) {
                r.ok = 1; r.N_ant = (w >> 2) == 0 ? 1 : ((w >> 2) == 1 ? 2 : 4); r.offset = w & 3;
                for (int i = 0; i < 24; i++) r.bits[i] = bits[w * 40 + i];
                break;
            }
        out[blockIdx.x] = r;
    }
}

// ------------------------------------------------------------------------------------------------ PCFICH + PDCCH
struct PdcchTables {                  // device pointers into one cached blob
    const int32_t *cce_re;            // [4][576]  RE (symbol*nsc + subcarrier) of element j of CCE j/36, per N_symbs 1..4; -1 = no such CCE
    const int32_t *pcfich_re;         // [16]
    const uint8_t *c_pdcch;           // [10][1152]
    const uint8_t *c_pcfich;          // [10][32]
    const int16_t *order[2];          // circular-buffer order for the two DCI sizes (+16 CRC bits)
    int n_dci[2];
};
struct PdcchParams { int nsc, N_rb_dl; };

__constant__ unsigned kCfiWords[4] = {0xB6DB6DB6u, 0x6DB6DB6Du, 0xDB6DB6DBu, 0x00000000u};   // 36.212 table 5.3.4-1, bit i of the code word at bit i

__global__ void __launch_bounds__(32 * N_HYP)
pdcch_decode_kernel(const float2 *__restrict__ grid, const float2 *__restrict__ ce /*[n][1][14][nsc]*/, const int *__restrict__ subfr_num,
                    PdcchTables T, PdcchParams P, lteb200_pdcch_raw *__restrict__ out)
{
    extern __shared__ int smem_i[];
    int *pm = smem_i;                                   // [N_HYP][PM_STRIDE]
    int *dd = pm + N_HYP * PM_STRIDE;                   // [N_HYP][3*CC_MAX_N]
    int16_t *soft = (int16_t *)(dd + N_HYP * 3 * CC_MAX_N);   // [6][576]
    uint8_t *bits = (uint8_t *)(soft + 6 * 576);        // [N_HYP][CC_MAX_N]
    __shared__ int s_cfi;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sn = subfr_num[blockIdx.x];
    const float2 *g = grid + (size_t)blockIdx.x * 16 * P.nsc;
    const float2 *cp = ce + (size_t)blockIdx.x * 14 * P.nsc;
    lteb200_pdcch_raw *o = out + blockIdx.x;
    // PCFICH: 16 QPSK symbols -> 32 bits -> nearest of the four code words (cfi_channel_decode: fewer than 4 bit errors)
    if (warp == 0) {
        unsigned hard = 0;
        if (lane < 16) {
            const int re = T.pcfich_re[lane];
            float xr, xi;
            zf_1port(g[re], cp[re], xr, xi);
            int b0, b1;
            qpsk_soft(xr, xi, b0, b1);
            const uint8_t *c = T.c_pcfich + sn * 32 + 2 * lane;
            const int v0 = b0 * (1 - 2 * (int)c[0]), v1 = b1 * (1 - 2 * (int)c[1]);
            hard = ((unsigned)(v0 < 0) << (2 * lane)) | ((unsigned)(v1 < 0) << (2 * lane + 1));
        }
        for (int s = 16; s > 0; s >>= 1) hard |= __shfl_xor_sync(0xffffffffu, hard, s);
        if (lane == 0) {
            int min_ber = 32, cfi = 0;
            for (int w = 0; w < 4; w++) { const int ber = __popc(hard ^ kCfiWords[w]); if (ber < min_ber) { min_ber = ber; cfi = w + 1; } }
            s_cfi = min_ber < 4 ? cfi : 0;
        }
    }
    __syncthreads();
    const int cfi = s_cfi;
    if (cfi == 0) {
        if (tid == 0) { o->cfi = 0; o->N_symbs = 0; }
        if (tid < N_HYP) { o->found[tid] = 0; o->rnti[tid] = 0; }
        o->bits[warp][lane] = 0;                        // deterministic output for subframes without a control region
        return;
    }
    const int N_symbs = cfi + (P.N_rb_dl <= 10 ? 1 : 0);
    // the 16 CCEs of the common search space: gather, zero forcing, QPSK demap, descramble with c(subframe, cell)
    // reserved CFI 4 on a <= 10 RB carrier would mean five control symbols: no CCE map exists, every element counts as missing (as in the oracle)
    const bool no_map = N_symbs > 4;
    const int32_t *re_of = T.cce_re + (no_map ? 0 : (N_symbs - 1) * 576);
    const uint8_t *c = T.c_pdcch + sn * 1152;
    for (int t = tid; t < 1152; t += blockDim.x) {
        // t < 576: aggregation level 4, candidate t/144, element t%144; else level 8, candidate (t-576)/288
        const int cand = t < 576 ? t / 144 : 4 + (t - 576) / 288;
        const int j = t < 576 ? t % 144 : (t - 576) % 288;
        const int e = t < 576 ? t : t - 576;            // element of the 16-CCE space
        const int re = no_map ? -1 : re_of[e];
        int b0 = 0, b1 = 0;                             // CCE beyond the control region: soft bits 0 (see oracle/port/control.cc)
        if (re >= 0) {
            float xr, xi;
            zf_1port(g[re], cp[re], xr, xi);
            qpsk_soft(xr, xi, b0, b1);
        }
        const int c_off = (t < 576 ? cand * 288 : (cand - 4) * 576) + 2 * j;
        soft[cand * 576 + 2 * j] = (int16_t)(b0 * (1 - 2 * (int)c[c_off]));
        soft[cand * 576 + 2 * j + 1] = (int16_t)(b1 * (1 - 2 * (int)c[c_off + 1]));
    }
    __syncthreads();
    {
        const int cand = warp >> 1, fmt = warp & 1;
        const int n_dci = T.n_dci[fmt], N = n_dci + 16, E = cand < 4 ? 288 : 576;
        int *d = dd + warp * 3 * CC_MAX_N;
        const int16_t *sf = soft + cand * 576;
        const int16_t *order = T.order[fmt];
        for (int m = lane; m < 3 * N; m += 32) {
            int acc = 0;
            for (int k = m; k < E; k += 3 * N) acc += sf[k];
            d[order[m]] = acc;
        }
        __syncwarp();
        uint8_t *b = bits + warp * CC_MAX_N;
        cc_viterbi_warp(d, N, pm + warp * PM_STRIDE, b, lane);
        unsigned rx = 0, x = 0;
        if (lane == 0) {
            for (int i = 0; i < 16; i++) rx = (rx << 1) | b[n_dci + i];
            x = rx ^ crc16_bits(b, n_dci);
        }
        x = __shfl_sync(0xffffffffu, x, 0);
        const int found = (x == 0xFFFFu || x == 0xFFFEu || (x >= 1u && x <= 0x3Cu)) ? 1 : 0;      // SI-, P-, RA-RNTI (liblte_phy.cc:4908-4935)
        if (lane == 0) { o->found[warp] = found; o->rnti[warp] = found ? (int)x : 0; }
        o->bits[warp][lane] = (found && lane < n_dci) ? b[lane] : 0;
    }
    if (tid == 0) { o->cfi = cfi; o->N_symbs = N_symbs; }
}

// ------------------------------------------------------------------------------------------------ host: tables
std::map<std::vector<int>, void *> g_ctl_tables;     // (device, kind, params...) -> device blob
std::mutex g_ctl_mu;

int upload_blob(lteb200_ctx *ctx, const std::vector<int> &key, const void *host, size_t bytes, void **out)
{
    void *d = nullptr;
    LTE_CUDA_OK(ctx, cudaMalloc(&d, bytes));
    LTE_CUDA_OK(ctx, cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice));
    g_ctl_tables[key] = d;
    *out = d;
    return LTEB200_OK;
}

// order in which rate_unmatch_conv's circular buffer visits the 3*N_c mother-code bits (sub-block interleaver, 36.212 5.1.4.2.1)
std::vector<int16_t> cc_order_host(int N_c)
{
    const int R = (N_c + 31) / 32, n_dummy = 32 * R - N_c;
    std::vector<int16_t> t;
    for (int x = 0; x < 3; x++)
        for (int j = 0; j < 32; j++)
            for (int i = 0; i < R; i++) {
                const int src = i * 32 + kColPermCC[j];
                if (src >= n_dummy) t.push_back((int16_t)(3 * (src - n_dummy) + x));
            }
    return t;
}

int get_cc_order(lteb200_ctx *ctx, int N_c, const int16_t **out)
{
    const std::vector<int> key = {ctx->device, 0, N_c};
    auto it = g_ctl_tables.find(key);
    if (it != g_ctl_tables.end()) { *out = (const int16_t *)it->second; return LTEB200_OK; }
    const std::vector<int16_t> t = cc_order_host(N_c);
    void *d = nullptr;
    const int rc = upload_blob(ctx, key, t.data(), t.size() * sizeof(int16_t), &d);
    *out = (const int16_t *)d;
    return rc;
}

int get_bch_scramble(lteb200_ctx *ctx, int N_id_cell, const uint8_t **out)
{
    const std::vector<int> key = {ctx->device, 1, N_id_cell};
    auto it = g_ctl_tables.find(key);
    if (it != g_ctl_tables.end()) { *out = (const uint8_t *)it->second; return LTEB200_OK; }
    std::vector<uint8_t> c(1920);
    prs_c_host((uint32_t)N_id_cell, 1920, c.data());      // liblte_phy.cc:4034
    void *d = nullptr;
    const int rc = upload_blob(ctx, key, c.data(), c.size(), &d);
    *out = (const uint8_t *)d;
    return rc;
}

void dci_sizes(int N_rb_dl, int *a, int *c)                // liblte_phy.cc:4844-4864
{
    switch (N_rb_dl) {
    case 6:  *a = 21; *c = 9;  break;
    case 15: *a = 22; *c = 11; break;
    case 25: *a = 25; *c = 13; break;
    case 50: *a = 27; *c = 13; break;
    case 75: *a = 27; *c = 14; break;
    default: *a = 28; *c = 15; break;
    }
}

// PCFICH and PHICH REG positions (first subcarrier), 36.211 6.7.4 / 6.9 as the reference evaluates them (:7903-7909, 8241-8282)
void control_reg_positions(int N_rb_dl, int N_id_cell, float phich_res, int pk[4], std::vector<int> &hk)
{
    float pn[4];
    const int k_hat = 6 * (N_id_cell % (2 * N_rb_dl));
    for (int i = 0; i < 4; i++) {
        pk[i] = (k_hat + (i * N_rb_dl / 2) * 6) % (12 * N_rb_dl);
        pn[i] = (float)(pk[i] / 6) - 0.5f;
    }
    const int n_group = (int)ceilf(phich_res * ((float)N_rb_dl / 8.0f));
    const uint32_t n_l = (uint32_t)(2 * N_rb_dl - 4);
    hk.clear();
    for (int m = 0; m < n_group; m++) {
        uint32_t n_hat[3];
        for (uint32_t i = 0; i < 3; i++) n_hat[i] = ((uint32_t)N_id_cell + (uint32_t)m + i * n_l / 3) % n_l;
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 3; j++)
                if ((float)n_hat[j] > pn[i]) n_hat[j]++;
        for (int i = 0; i < 3; i++) hk.push_back((int)(n_hat[i] * 6));
    }
}

// cce_re[j] for the first 16 CCEs with N_symbs control symbols; grid row length nsc.  REG walk :4591-4701, un-shift :4702-4716,
// REG de-interleaver :4717-4818, CCE assembly :4819-4838.
void build_cce_map(int N_rb_dl, int N_id_cell, int N_ant, int N_symbs, float phich_res, int nsc, int32_t *cce_re /*[576]*/)
{
    for (int i = 0; i < 576; i++) cce_re[i] = -1;
    int pk[4];
    std::vector<int> hk;
    control_reg_positions(N_rb_dl, N_id_cell, phich_res, pk, hk);
    int N_reg = N_symbs * N_rb_dl * 3 - N_rb_dl - 4 - (int)hk.size();
    if (N_ant == 4) N_reg -= N_rb_dl;
    if (N_reg < 9) return;
    std::vector<int32_t> reg((size_t)N_reg * 4, -1);
    int m = 0;
    for (int k = 0; k < 12 * N_rb_dl; k++)
        for (int l = 0; l < N_symbs; l++) {
            if (l == 0 || (l == 1 && N_ant == 4)) {      // symbols with CRS: REGs of six subcarriers, four usable
                bool ok = k % 6 == 0;
                if (l == 0) {
                    for (int i = 0; i < 4; i++) ok = ok && k != pk[i];
                    for (int q : hk) ok = ok && k != q;
                }
                if (!ok || m >= N_reg) continue;
                int j = 0;
                for (int i = 0; i < 6; i++)
                    if (N_id_cell % 3 != i % 3) reg[(size_t)m * 4 + j++] = l * nsc + k + i;
                m++;
            } else if (k % 4 == 0 && m < N_reg) {
                for (int i = 0; i < 4; i++) reg[(size_t)m * 4 + i] = l * nsc + k + i;
                m++;
            }
        }
    const int R = (N_reg + 31) / 32, n_dummy = 32 * R - N_reg;
    std::vector<int> deint(32 * R, -1);                  // interleaver input position -> index of the (shifted) REG carrying it
    int kk = 0;
    for (int j = 0; j < 32; j++)
        for (int i = 0; i < R; i++) {
            const int src = i * 32 + kColPermCC[j];
            if (src >= n_dummy) deint[src] = kk++;
        }
    const int N_cce = N_reg / 9;
    for (int cce = 0; cce < N_cce && cce < 16; cce++)
        for (int r = 0; r < 9; r++) {
            const int shifted = deint[n_dummy + cce * 9 + r];
            const int orig = ((shifted - N_id_cell) % N_reg + N_reg) % N_reg;      // shifted[(i + N_id_cell) % N_reg] = reg[i]
            for (int j = 0; j < 4; j++) cce_re[cce * 36 + r * 4 + j] = reg[(size_t)orig * 4 + j];
        }
}

int get_pdcch_tables(lteb200_ctx *ctx, int N_rb_dl, int N_id_cell, int N_ant, float phich_res, PdcchTables *T)
{
    const int nsc = 12 * N_rb_dl;
    const std::vector<int> key = {ctx->device, 2, N_rb_dl, N_id_cell, N_ant, (int)lrintf(phich_res * 6.0f)};
    void *blob = nullptr;
    auto it = g_ctl_tables.find(key);
    const size_t off_pc = 4 * 576 * sizeof(int32_t), off_c = off_pc + 16 * sizeof(int32_t), off_cp = off_c + 10 * 1152, total = off_cp + 10 * 32;
    if (it != g_ctl_tables.end()) blob = it->second;
    else {
        std::vector<uint8_t> h(total);
        int32_t *cce = (int32_t *)h.data();
        for (int ns = 1; ns <= 4; ns++) build_cce_map(N_rb_dl, N_id_cell, N_ant, ns, phich_res, nsc, cce + (ns - 1) * 576);
        int32_t *pc = (int32_t *)(h.data() + off_pc);
        int pk[4];
        std::vector<int> hk;
        control_reg_positions(N_rb_dl, N_id_cell, phich_res, pk, hk);
        for (int i = 0; i < 4; i++) {
            int j = 0;
            for (int q = 0; q < 6; q++)
                if (N_id_cell % 3 != q % 3) pc[4 * i + j++] = pk[i] + q;                          // symbol 0
        }
        for (int sn = 0; sn < 10; sn++) {
            prs_c_host(((uint32_t)sn << 9) + (uint32_t)N_id_cell, 1152, h.data() + off_c + sn * 1152);                              // :4841
            prs_c_host((((uint32_t)(sn + 1) * (2 * (uint32_t)N_id_cell + 1)) << 9) + (uint32_t)N_id_cell, 32, h.data() + off_cp + sn * 32);  // :7930
        }
        const int rc = upload_blob(ctx, key, h.data(), total, &blob);
        if (rc) return rc;
    }
    const uint8_t *b = (const uint8_t *)blob;
    T->cce_re = (const int32_t *)b; T->pcfich_re = (const int32_t *)(b + off_pc); T->c_pdcch = b + off_c; T->c_pcfich = b + off_cp;
    dci_sizes(N_rb_dl, &T->n_dci[0], &T->n_dci[1]);
    for (int f = 0; f < 2; f++) {
        const int rc = get_cc_order(ctx, T->n_dci[f] + 16, &T->order[f]);
        if (rc) return rc;
    }
    return LTEB200_OK;
}

// 36.213 table 7.1.7.2.1-1, columns N_PRB = 2 and 3 (what DCI 1A addressed to SI-/P-/RA-RNTI can select), and table 7.1.7.2.3-1 (DCI 1C)
const int kTbsNprb2[27] = {32, 56, 72, 104, 120, 144, 176, 224, 256, 296, 328, 376, 440, 488, 552, 600, 632, 696, 776, 840, 904, 1000, 1064, 1128, 1192, 1256, 1480};
const int kTbsNprb3[27] = {56, 88, 144, 176, 208, 224, 256, 328, 392, 456, 504, 584, 680, 744, 840, 904, 968, 1064, 1160, 1288, 1384, 1480, 1608, 1736, 1800, 1864, 2216};
const int kTbs1C[32] = {40, 56, 72, 120, 136, 144, 176, 208, 224, 256, 280, 296, 328, 336, 392, 488,
                        552, 600, 632, 696, 776, 840, 904, 1000, 1064, 1128, 1224, 1288, 1384, 1480, 1608, 1736};

struct BitReader {
    const uint8_t *p;
    uint32_t get(int n) { uint32_t v = 0; while (n-- > 0) v = (v << 1) | (*p++ & 1u); return v; }
};

void alloc_defaults(lteb200_dci *o, int rnti, int N_ant, int fmt)
{
    memset(o, 0, sizeof(*o));
    o->alloc.mod_type = LTEB200_MOD_QPSK;
    o->alloc.tx_mode = N_ant == 1 ? 1 : 2;
    o->alloc.rnti = rnti;
    o->format = fmt;
}

// dci_1a_unpack for SI-/P-/RA-RNTI (liblte_phy.cc:13273-13372).  0, or LTEB200_ERR_CONTENTS as the reference's INVALID_CONTENTS.
int unpack_1a(const uint8_t *bits, int rnti, int N_rb_dl, int N_ant, lteb200_dci *o)
{
    BitReader r{bits};
    alloc_defaults(o, rnti, N_ant, 0);
    if (r.get(1) == 0) return LTEB200_ERR_CONTENTS;                                        // format 0
    const uint32_t distributed = r.get(1);
    const int riv_len = (int)ceilf(logf((float)(N_rb_dl * (N_rb_dl + 1) / 2)) / logf(2));
    const uint32_t riv = r.get(riv_len);
    o->alloc.N_prb = (int)(riv / N_rb_dl) + 1;
    const int rb0 = (int)(riv % N_rb_dl);
    o->mcs = (int)r.get(5);
    r.get(3); r.get(1);                                                                     // HARQ process, NDI: unused for these RNTIs
    o->alloc.rv_idx = (int)r.get(2);
    const int tpc = (int)r.get(2);
    if (!distributed)
        for (int i = 0; i < o->alloc.N_prb && i < 110; i++) o->alloc.prb[0][i] = o->alloc.prb[1][i] = rb0 + i;
    if (o->mcs >= 27) return LTEB200_ERR_CONTENTS;
    o->alloc.tbs = (tpc % 2 == 0) ? kTbsNprb2[o->mcs] : kTbsNprb3[o->mcs];                 // N_PRB^1A = 2 or 3, 36.213 7.1.7
    return LTEB200_OK;
}

// dci_1c_unpack (liblte_phy.cc:13400-13605): RIV over distributed VRBs in steps of N_RB^step, then the VRB -> PRB interleaver of 36.211 6.2.3.2
int unpack_1c(const uint8_t *bits, int rnti, int N_rb_dl, int N_ant, lteb200_dci *o)
{
    BitReader r{bits};
    alloc_defaults(o, rnti, N_ant, 1);
    const uint32_t N = (uint32_t)N_rb_dl;
    const uint32_t gap2 = N < 50 ? 0u : r.get(1);
    uint32_t N_gap;
    if (N <= 10) N_gap = (uint32_t)ceilf(N / 2.0);
    else if (N == 11) N_gap = 4;
    else if (N <= 19) N_gap = 8;
    else if (N <= 26) N_gap = 12;
    else if (N <= 44) N_gap = 18;
    else if (N <= 49) N_gap = 27;
    else if (N <= 63) N_gap = gap2 ? 9 : 27;
    else if (N <= 70) N_gap = gap2 ? 16 : 32;
    else N_gap = gap2 ? 16 : 48;
    const uint32_t vrb_gap1 = 2 * (N_gap < N - N_gap ? N_gap : N - N_gap), vrb_gap2 = (N / (2 * N_gap)) * 2 * N_gap;
    const uint32_t N_vrb = gap2 ? vrb_gap2 : vrb_gap1, step = N <= 49 ? 2 : 4, Nt = gap2 ? 2 * N_gap : N_vrb;
    const uint32_t g1 = vrb_gap1 / step;
    const uint32_t riv = r.get((int)ceilf(logf(g1 * (g1 + 1) / 2.0) / logf(2)));
    const uint32_t nd = N_vrb / step;
    uint32_t L = 0, start = 0;
    for (uint32_t len = 1; len <= nd; len++)             // the reference enumerates (length, start) and keeps the last match (:13496-13522)
        for (uint32_t s0 = 0; s0 < nd; s0++) {
            const uint32_t v = (len - 1) <= nd / 2 ? nd * (len - 1) + s0 : nd * (nd - len + 1) + (nd - 1 - s0);
            if (v == riv) { L = len * step; start = s0 * step; }
        }
    o->alloc.N_prb = (int)L;
    o->mcs = (int)r.get(5);
    const uint32_t P = N <= 10 ? 1 : (N <= 26 ? 2 : (N <= 63 ? 3 : 4));
    const uint32_t N_row = (uint32_t)(ceilf(Nt / (4.0 * P)) * P), N_null = 4 * N_row - Nt;
    uint32_t n_vrb = start;
    for (uint32_t i = 0; i < L && i < 110; i++, n_vrb++) {
        const uint32_t nt = n_vrb % Nt, blk = Nt * (n_vrb / Nt);
        const uint32_t two_col = 2 * N_row * (nt % 2) + nt / 2 + blk, four_col = N_row * (nt % 4) + nt / 4 + blk;
        uint32_t even;
        if (N_null != 0 && nt >= Nt - N_null) even = (nt % 2 == 1) ? two_col - N_row : two_col - N_row + N_null / 2;
        else if (N_null != 0 && nt % 4 >= 2) even = four_col - N_null / 2;
        else even = four_col;
        const uint32_t odd = (even + Nt / 2) % Nt + blk;
        o->alloc.prb[0][i] = (int32_t)(even < Nt / 2 ? even : even + N_gap - Nt / 2);
        o->alloc.prb[1][i] = (int32_t)(odd < Nt / 2 ? odd : odd + N_gap - Nt / 2);
    }
    if (o->mcs >= 32) return LTEB200_ERR_CONTENTS;
    o->alloc.tbs = kTbs1C[o->mcs];
    return LTEB200_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ C ABI
// HOST utility: resource elements of the first 16 CCEs (the common search space), element j of CCE c at out[36*c + j] as
// symbol*(12*N_rb_dl) + subcarrier, -1 where the control region has no such CCE.  What the decode kernel gathers from; exposed so that
// the mapping can be checked without a device (tests/test_control_oracle.py) and reused by signal generators.
extern "C" int lteb200_pdcch_cce_map(int N_rb_dl, int N_id_cell, int N_ant, int N_symbs, float phich_res, int32_t *out576)
{
    if (!out576 || N_rb_dl < 6 || N_rb_dl > 110 || N_id_cell < 0 || N_id_cell > 503 || N_symbs < 1 || N_symbs > 4 || !(N_ant == 1 || N_ant == 2 || N_ant == 4))
        return LTEB200_ERR_ARGS;
    build_cce_map(N_rb_dl, N_id_cell, N_ant, N_symbs, phich_res, 12 * N_rb_dl, out576);
    return LTEB200_OK;
}

extern "C" int lteb200_bch_decode_batch(lteb200_ctx *ctx, const float *d_grid, const float *d_ce, int N_rb_dl, int N_id_cell, int n,
                                        lteb200_bch_result *d_out, void *stream)
{
    if (!ctx || !d_grid || !d_ce || !d_out || n < 0 || N_rb_dl < 6 || N_rb_dl > 110 || N_id_cell < 0 || N_id_cell > 503) return LTEB200_ERR_ARGS;
    if (n == 0) return LTEB200_OK;
    LTE_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> lk(g_ctl_mu);
    const uint8_t *c = nullptr;
    const int16_t *order = nullptr;
    int rc = get_bch_scramble(ctx, N_id_cell, &c);
    if (rc == LTEB200_OK) rc = get_cc_order(ctx, 40, &order);
    if (rc) return rc;
    BchParams P{12 * N_rb_dl, 6 * N_rb_dl - 36, N_id_cell % 3};
    const size_t smem = sizeof(int) * (N_HYP * 41 * 64 + N_HYP * 120) + sizeof(float2) * 5 * 240 + sizeof(int16_t) * 3 * 480 + N_HYP * 40;
    LTE_CUDA_OK(ctx, cudaFuncSetAttribute(bch_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per device: set on every call
    bch_decode_kernel<<<n, 32 * N_HYP, smem, (cudaStream_t)stream>>>((const float2 *)d_grid, (const float2 *)d_ce, c, order, P, d_out);
    ctx->launches += 1;
    LTE_CUDA_OK(ctx, cudaGetLastError());
    return LTEB200_OK;
}

extern "C" int lteb200_pdcch_decode_batch(lteb200_ctx *ctx, const float *d_grid, const float *d_ce, int N_rb_dl, int N_id_cell, int N_ant,
                                          float phich_res, const int *subfr_num, int n_sf, lteb200_pdcch_raw *d_out, void *stream_)
{
    if (!ctx || !d_grid || !d_ce || !d_out || !subfr_num || n_sf < 0 || N_rb_dl < 6 || N_rb_dl > 110 || N_id_cell < 0 || N_id_cell > 503)
        return LTEB200_ERR_ARGS;
    if (N_ant != 1) {
        lteb200_set_error(ctx, "pdcch_decode: N_ant = %d: the reference's PCFICH/PDCCH receive path is undefined for more than one port "
                               "(port stride 576 over rows of 288, liblte_phy.cc:4893 / liblte_phy.h:405)", N_ant);
        return LTEB200_ERR_UNSUPPORTED;
    }
    if (n_sf == 0) return LTEB200_OK;
    for (int i = 0; i < n_sf; i++) if (subfr_num[i] < 0 || subfr_num[i] > 9) return LTEB200_ERR_ARGS;
    cudaStream_t stream = (cudaStream_t)stream_;
    LTE_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> lk(g_ctl_mu);
    PdcchTables T;
    int rc = get_pdcch_tables(ctx, N_rb_dl, N_id_cell, N_ant, phich_res, &T);
    if (rc) return rc;
    int *d_sn = nullptr;
    LTE_CUDA_OK(ctx, cudaMallocAsync(&d_sn, sizeof(int) * n_sf, stream));
    if (int rc_ = lteb200_upload_small(ctx, subfr_num, sizeof(int) * n_sf, d_sn, stream)) return rc_;
    PdcchParams P{12 * N_rb_dl, N_rb_dl};
    const size_t smem = sizeof(int) * (N_HYP * PM_STRIDE + N_HYP * 3 * CC_MAX_N) + sizeof(int16_t) * 6 * 576 + N_HYP * CC_MAX_N;
    LTE_CUDA_OK(ctx, cudaFuncSetAttribute(pdcch_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pdcch_decode_kernel<<<n_sf, 32 * N_HYP, smem, stream>>>((const float2 *)d_grid, (const float2 *)d_ce, d_sn, T, P, d_out);
    ctx->launches += 1;
    LTE_CUDA_OK(ctx, cudaGetLastError());
    LTE_CUDA_OK(ctx, cudaFreeAsync(d_sn, stream));
    return LTEB200_OK;
}

// The allocation list of liblte_phy_pdcch_channel_decode from the raw blind-decode results, in the reference's order:
// aggregation level 4 candidates 0..3 then level 8 candidates 0..1, DCI 1A before 1C, at most 6 (liblte_phy.cc:4866-5131).
// Return value as the reference's: 0 if the last DCI whose CRC matched unpacked cleanly, LTEB200_ERR_CONTENTS if it did not,
// LTEB200_ERR_ARGS (its INVALID_INPUTS) if no CRC matched at all, LTEB200_ERR_DECODE if the PCFICH did not decode.
extern "C" int lteb200_pdcch_unpack(const lteb200_pdcch_raw *raw, int N_rb_dl, int N_ant, lteb200_dci *allocs, int *N_alloc)
{
    if (!raw || !allocs || !N_alloc) return LTEB200_ERR_ARGS;
    *N_alloc = 0;
    if (raw->cfi == 0) return LTEB200_ERR_DECODE;
    int err = LTEB200_ERR_ARGS;
    for (int slot = 0; slot < 12; slot++) {
        if (!raw->found[slot] || *N_alloc >= 6) continue;
        lteb200_dci *o = &allocs[*N_alloc];
        err = (slot & 1) == 0 ? unpack_1a(raw->bits[slot], raw->rnti[slot], N_rb_dl, N_ant, o) : unpack_1c(raw->bits[slot], raw->rnti[slot], N_rb_dl, N_ant, o);
        if (err == LTEB200_OK) (*N_alloc)++;
    }
    return err;
}

namespace {
// subframe-struct planes (host, [rows][1200] floats) -> device grid / channel-estimate layout of the batched kernels
int stage_planes(lteb200_ctx *ctx, const float *rx_re, const float *rx_im, const float *ce_re, const float *ce_im, int n_ports, int nsc,
                 float **d_in, float **d_grid, float **d_ce)
{
    const size_t plane = 16 * 1200;
    LTE_CUDA_OK(ctx, cudaMalloc(d_in, sizeof(float) * plane * (2 + 2 * n_ports)));
    LTE_CUDA_OK(ctx, cudaMalloc(d_grid, sizeof(float) * 2 * 16 * nsc));
    LTE_CUDA_OK(ctx, cudaMalloc(d_ce, sizeof(float) * 2 * n_ports * 14 * nsc));
    LTE_CUDA_OK(ctx, cudaMemcpy(*d_in, rx_re, sizeof(float) * plane, cudaMemcpyHostToDevice));
    LTE_CUDA_OK(ctx, cudaMemcpy(*d_in + plane, rx_im, sizeof(float) * plane, cudaMemcpyHostToDevice));
    for (int p = 0; p < n_ports; p++) {
        LTE_CUDA_OK(ctx, cudaMemcpy(*d_in + plane * (2 + 2 * p), ce_re + p * plane, sizeof(float) * plane, cudaMemcpyHostToDevice));
        LTE_CUDA_OK(ctx, cudaMemcpy(*d_in + plane * (3 + 2 * p), ce_im + p * plane, sizeof(float) * plane, cudaMemcpyHostToDevice));
    }
    int rc = soa_to_grid(ctx, *d_in, *d_in + plane, 16, nsc, 1200, *d_grid, nullptr);
    for (int p = 0; p < n_ports && rc == LTEB200_OK; p++)
        rc = soa_to_grid(ctx, *d_in + plane * (2 + 2 * p), *d_in + plane * (3 + 2 * p), 14, nsc, 1200, *d_ce + (size_t)p * 2 * 14 * nsc, nullptr);
    return rc;
}
}  // namespace

// Host twin of liblte_phy_bch_channel_decode (liblte_phy.cc:3968): rx_ce_* hold four ports, [4][16][1200].
extern "C" int lteb200_bch_channel_decode(lteb200_ctx *ctx, const float *rx_symb_re, const float *rx_symb_im, const float *rx_ce_re,
                                          const float *rx_ce_im, int N_rb_dl, uint32_t N_id_cell, uint8_t *N_ant, uint8_t *out_bits,
                                          uint32_t *N_out_bits, uint8_t *offset)
{
    if (!ctx || !rx_symb_re || !rx_symb_im || !rx_ce_re || !rx_ce_im || !N_ant || !out_bits || !N_out_bits || !offset || N_id_cell > 503)
        return LTEB200_ERR_ARGS;
    LTE_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    float *d_in = nullptr, *d_grid = nullptr, *d_ce = nullptr;
    lteb200_bch_result *d_out = nullptr, r;
    int rc = stage_planes(ctx, rx_symb_re, rx_symb_im, rx_ce_re, rx_ce_im, 4, 12 * N_rb_dl, &d_in, &d_grid, &d_ce);
    if (rc == LTEB200_OK && cudaMalloc(&d_out, sizeof(r)) != cudaSuccess) rc = LTEB200_ERR_CUDA;
    if (rc == LTEB200_OK) rc = lteb200_bch_decode_batch(ctx, d_grid, d_ce, N_rb_dl, (int)N_id_cell, 1, d_out, nullptr);
    if (rc == LTEB200_OK && cudaMemcpy(&r, d_out, sizeof(r), cudaMemcpyDeviceToHost) != cudaSuccess) {
        lteb200_set_error(ctx, "bch_channel_decode: %s", cudaGetErrorString(cudaGetLastError()));
        rc = LTEB200_ERR_CUDA;
    }
    cudaFree(d_in); cudaFree(d_grid); cudaFree(d_ce); cudaFree(d_out);
    if (rc != LTEB200_OK) return rc;
    *N_ant = 0;                                          // the reference clears it before searching (:4037)
    if (!r.ok) return LTEB200_ERR_DECODE;
    *N_ant = (uint8_t)r.N_ant; *offset = (uint8_t)r.offset; *N_out_bits = 24;
    memcpy(out_bits, r.bits, 24);
    return LTEB200_OK;
}

// Host twin of liblte_phy_pdcch_channel_decode (liblte_phy.cc:4519) for N_ant == 1.
extern "C" int lteb200_pdcch_channel_decode(lteb200_ctx *ctx, const float *rx_symb_re, const float *rx_symb_im, const float *rx_ce_re,
                                            const float *rx_ce_im, int N_rb_dl, uint32_t subfr_num, uint32_t N_id_cell, uint8_t N_ant,
                                            float phich_res, int phich_dur, uint32_t *cfi, uint32_t *N_symbs, lteb200_dci *allocs,
                                            uint32_t *N_alloc)
{
    if (!ctx || !rx_symb_re || !rx_symb_im || !rx_ce_re || !rx_ce_im || !cfi || !N_symbs || !allocs || !N_alloc || N_id_cell > 503 || subfr_num > 9)
        return LTEB200_ERR_ARGS;
    if (phich_dur != 0) { lteb200_set_error(ctx, "pdcch_channel_decode: extended PHICH duration is not handled (nor by the reference, liblte_phy.cc:8283)"); return LTEB200_ERR_UNSUPPORTED; }
    if (N_ant != 1) {
        lteb200_set_error(ctx, "pdcch_channel_decode: N_ant = %d is outside the envelope in which the reference is defined (see lte_b200.h)", (int)N_ant);
        return LTEB200_ERR_UNSUPPORTED;
    }
    LTE_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    float *d_in = nullptr, *d_grid = nullptr, *d_ce = nullptr;
    lteb200_pdcch_raw *d_out = nullptr, raw;
    int rc = stage_planes(ctx, rx_symb_re, rx_symb_im, rx_ce_re, rx_ce_im, 1, 12 * N_rb_dl, &d_in, &d_grid, &d_ce);
    if (rc == LTEB200_OK && cudaMalloc(&d_out, sizeof(raw)) != cudaSuccess) rc = LTEB200_ERR_CUDA;
    const int sn = (int)subfr_num;
    if (rc == LTEB200_OK) rc = lteb200_pdcch_decode_batch(ctx, d_grid, d_ce, N_rb_dl, (int)N_id_cell, 1, phich_res, &sn, 1, d_out, nullptr);
    if (rc == LTEB200_OK && cudaMemcpy(&raw, d_out, sizeof(raw), cudaMemcpyDeviceToHost) != cudaSuccess) {
        lteb200_set_error(ctx, "pdcch_channel_decode: %s", cudaGetErrorString(cudaGetLastError()));
        rc = LTEB200_ERR_CUDA;
    }
    cudaFree(d_in); cudaFree(d_grid); cudaFree(d_ce); cudaFree(d_out);
    if (rc != LTEB200_OK) return rc;
    *cfi = (uint32_t)raw.cfi; *N_symbs = (uint32_t)raw.N_symbs;
    int n = 0;
    rc = lteb200_pdcch_unpack(&raw, N_rb_dl, N_ant, allocs, &n);
    *N_alloc = (uint32_t)n;
    return rc;
}

// =====================================================================================================================================
// Transmit side of the control region and of the synchronisation / broadcast signals, for test-signal generation on
// the device.  Twins of liblte_phy_map_pss :5265, liblte_phy_map_sss :5520, liblte_phy_bch_channel_encode :3863 (+ bch_channel_encode
// :12510), pcfich_channel_map :7805 (+ cfi_channel_encode :13616), liblte_phy_pdcch_channel_encode :4113 (+ dci_1a_pack :13153,
// dci_channel_encode :12882, conv_encode :9997, rate_match_conv :11499).  The bit-level encoding of the few distinct messages (one MIB per
// 40 ms, one DCI per allocation) is host work; everything per subframe and per resource element -- scrambling with the subframe's Gold
// sequence, QPSK mapping, CCE placement, REG interleaving / cyclic shift / resource mapping (through the receiver's own CCE map, inverted)
// -- is one kernel with one CTA per subframe.  One antenna port for PCFICH/PDCCH (the reference is undefined beyond, see the receive side).
namespace {

unsigned crc16_host(const uint8_t *b, int n)        // calc_crc with CRC16 = 0x11021, liblte_phy.cc:9713-9743
{
    unsigned rem = 0;
    for (int i = 0; i < n + 16; i++) {
        rem = (rem << 1) | (i < n ? (b[i] & 1u) : 0u);
        if (rem & 0x10000u) rem ^= 0x11021u;
    }
    return rem & 0xffffu;
}
// tail-biting convolutional code, constraint length 7, g = {0133, 0171, 0165} (conv_encode with tail_bit = true, liblte_phy.cc:9997-10057)
void conv_encode_tb(const uint8_t *c, int n, uint8_t *d /*[3 n], d[3 i + j]*/)
{
    const unsigned g[3] = {0133, 0171, 0165};
    uint8_t s[7];
    for (int i = 0; i < 7; i++) s[i] = c[n - i - 1];
    for (int i = 0; i < n; i++) {
        for (int j = 6; j > 0; j--) s[j] = s[j - 1];
        s[0] = c[i];
        for (int j = 0; j < 3; j++) {
            unsigned acc = 0;
            for (int k = 0; k < 7; k++) acc += s[k] * ((g[j] >> (6 - k)) & 1u);
            d[3 * i + j] = (uint8_t)(acc & 1u);
        }
    }
}
// rate_match_conv: the circular buffer read in the order of the sub-block interleaver, NULLs skipped
void rate_match_cc(const uint8_t *d, int n_c, int n_e, uint8_t *e)
{
    const std::vector<int16_t> order = cc_order_host(n_c);
    for (int k = 0; k < n_e; k++) e[k] = d[order[(size_t)k % order.size()]];
}

struct CtrlTxParams {
    const int32_t *cce_re;      // [576] for this N_symbs
    const int32_t *pcfich_re;   // [16]
    const uint8_t *c_pdcch;     // [10][1152]
    const uint8_t *c_pcfich;    // [10][32]
    const uint8_t *dci_e;       // [n_dci][288] encoded + rate matched DCIs
    const uint8_t *bch_e;       // [n_bch_rows][480] encoded, rate matched and scrambled PBCH bits of one radio frame
    const int32_t *bch_row;     // [n_sf] row of bch_e for subframes that carry the PBCH, else -1
    const float2 *sync;         // [3][62] PSS, SSS of subframe 0, SSS of subframe 5
    const int *subfr_num;
    float2 *grid;               // port-0 plane of subframe sf at grid + sf * sf_stride
    size_t sf_stride, port_stride;
    int nsc, n_dci, cfi, N_id_cell, N_ant, do_sync, do_pdcch;
};

__device__ __forceinline__ float2 qpsk(int b0, int b1)
{
    const float r = (float)(1.0 / sqrt(2.0));        // modulation_mapper, liblte_phy.cc:8740-8747
    return make_float2(__fmul_rn(1.f - 2.f * (float)b0, r), __fmul_rn(1.f - 2.f * (float)b1, r));
}

__global__ void __launch_bounds__(128)
ctrl_tx_kernel(const __grid_constant__ CtrlTxParams P)
{
    const int sf = blockIdx.x, t = threadIdx.x;
    const int sn = P.subfr_num[sf];
    float2 *g = P.grid + (size_t)sf * P.sf_stride;
    if (P.do_pdcch) {
        // PCFICH: 32-bit CFI code word, scrambled, 16 QPSK symbols on four REGs of symbol 0
        if (t < 16) {
            const unsigned w = kCfiWords[P.cfi - 1];
            const uint8_t *c = P.c_pcfich + sn * 32;
            g[P.pcfich_re[t]] = qpsk(((w >> (2 * t)) & 1u) ^ c[2 * t], ((w >> (2 * t + 1)) & 1u) ^ c[2 * t + 1]);
        }
        // PDCCH: DCI a occupies CCEs 4a .. 4a+3 of the common search space; 144 QPSK symbols each; the REs of every other CCE stay zero
        const uint8_t *c = P.c_pdcch + sn * 1152;
        for (int m = t; m < 144 * P.n_dci; m += blockDim.x) {
            const int a = m / 144, j = m - a * 144;
            const uint8_t *e = P.dci_e + a * 288;
            const int re = P.cce_re[a * 144 + j];
            if (re >= 0) g[re] = qpsk(e[2 * j] ^ c[a * 288 + 2 * j], e[2 * j + 1] ^ c[a * 288 + 2 * j + 1]);
        }
    }
    if (P.do_sync && (sn == 0 || sn == 5)) {
        // PSS in symbol 6, SSS in symbol 5, 62 subcarriers around DC, on every port (liblte_phy.cc:5265, :5520)
        if (t < 62) {
            const int k = t - 31 + P.nsc / 2;
            for (int p = 0; p < P.N_ant; p++) {
                g[p * P.port_stride + 6 * P.nsc + k] = P.sync[t];
                g[p * P.port_stride + 5 * P.nsc + k] = P.sync[(sn == 0 ? 62 : 124) + t];
            }
        }
    }
    const int row = P.bch_row[sf];
    if (row >= 0) {
        // PBCH: 240 QPSK symbols (transmit diversity for 2 / 4 ports), symbols 7-10 of subframe 0, CRS positions of four ports left out in
        // symbols 7 and 8 (liblte_phy.cc:3934-3953).  The reference advances its symbol index over the ports without resetting it (idx, :3936):
        // port p's symbols 7/8 carry y_p[48 p + ...]; restated as is.
        const uint8_t *e = P.bch_e + (size_t)row * 480;
        const float r2 = (float)(1.0 / sqrt(2.0));
        auto y = [&](int p, int i) -> float2 {                  // pre_coder_dl output of port p, symbol i (liblte_phy.cc:7526-7640)
            if (P.N_ant == 1) return qpsk(e[2 * i], e[2 * i + 1]);
            if (P.N_ant == 2) {
                const int i0 = i & ~1;
                const float2 d0 = qpsk(e[2 * i0], e[2 * i0 + 1]), d1 = qpsk(e[2 * i0 + 2], e[2 * i0 + 3]);
                if (p == 0) return (i & 1) ? make_float2(__fmul_rn(r2, d1.x), __fmul_rn(r2, d1.y)) : make_float2(__fmul_rn(r2, d0.x), __fmul_rn(r2, d0.y));
                return (i & 1) ? make_float2(__fmul_rn(r2, d0.x), __fmul_rn(-r2, d0.y)) : make_float2(__fmul_rn(-r2, d1.x), __fmul_rn(r2, d1.y));
            }
            const int i0 = i & ~3, q = i & 3;
            const float2 d0 = qpsk(e[2 * i0], e[2 * i0 + 1]), d1 = qpsk(e[2 * i0 + 2], e[2 * i0 + 3]);
            const float2 d2 = qpsk(e[2 * i0 + 4], e[2 * i0 + 5]), d3 = qpsk(e[2 * i0 + 6], e[2 * i0 + 7]);
            const float2 zero = make_float2(0.f, 0.f);
            if (p == 0) return q == 0 ? make_float2(__fmul_rn(r2, d0.x), __fmul_rn(r2, d0.y)) : (q == 1 ? make_float2(__fmul_rn(r2, d1.x), __fmul_rn(r2, d1.y)) : zero);
            if (p == 1) return q == 2 ? make_float2(__fmul_rn(r2, d2.x), __fmul_rn(r2, d2.y)) : (q == 3 ? make_float2(__fmul_rn(r2, d3.x), __fmul_rn(r2, d3.y)) : zero);
            if (p == 2) return q == 0 ? make_float2(__fmul_rn(-r2, d1.x), __fmul_rn(r2, d1.y)) : (q == 1 ? make_float2(__fmul_rn(r2, d0.x), __fmul_rn(-r2, d0.y)) : zero);
            return q == 2 ? make_float2(__fmul_rn(-r2, d3.x), __fmul_rn(r2, d3.y)) : (q == 3 ? make_float2(__fmul_rn(r2, d2.x), __fmul_rn(-r2, d2.y)) : zero);
        };
        if (t < 72) {
            const int k = P.nsc / 2 - 36 + t;
            // rank of this subcarrier among the non-CRS ones of symbols 7 and 8
            int idx = 0;
            for (int i = 0; i < t; i++) idx += (P.N_id_cell % 3) != (i % 3);
            for (int p = 0; p < P.N_ant; p++) {
                float2 *gp = g + p * P.port_stride;
                if ((P.N_id_cell % 3) != (t % 3)) {
                    gp[7 * P.nsc + k] = y(p, 48 * p + idx);
                    gp[8 * P.nsc + k] = y(p, 48 * p + idx + 48);
                }
                gp[9 * P.nsc + k] = y(p, t + 96);
                gp[10 * P.nsc + k] = y(p, t + 168);
            }
        }
    }
}

}  // namespace

// DCI format 1A of one allocation (dci_1a_pack, liblte_phy.cc:13153-13250: localized VRBs, HARQ process 0) and its channel coding to 288 bits
// (dci_channel_encode :12882, aggregation level 4).  Returns the transport block size the receiver will derive (36.213 7.1.7), or -1.
static int dci_1a_encode_host(const lteb200_alloc &al, int mcs, int N_rb_dl, uint8_t *e288)
{
    if (al.N_prb < 1 || al.N_prb > N_rb_dl || mcs < 0 || mcs > 26) return -1;
    for (int j = 1; j < al.N_prb; j++) if (al.prb[0][j] != al.prb[0][0] + j) return -1;     // resource allocation type 2: contiguous
    const bool common = al.rnti == 0xFFFF || al.rnti == 0xFFFE || (al.rnti >= 0x0001 && al.rnti <= 0x003C);
    uint8_t dci[64];
    int n = 0;
    auto put = [&](unsigned v, int bits) { for (int i = bits - 1; i >= 0; i--) dci[n++] = (uint8_t)((v >> i) & 1u); };
    put(1, 1);                                           // format 0/1A flag: 1A
    put(0, 1);                                           // localized VRBs
    const int riv_len = (int)ceilf(logf((float)(N_rb_dl * (N_rb_dl + 1) / 2)) / logf(2.0f));
    unsigned riv;
    if ((al.N_prb - 1) <= N_rb_dl / 2) riv = (unsigned)(N_rb_dl * (al.N_prb - 1) + al.prb[0][0]);
    else riv = (unsigned)(N_rb_dl * (N_rb_dl - al.N_prb + 1) + (N_rb_dl - 1 - al.prb[0][0]));
    put(riv, riv_len);
    put((unsigned)mcs, 5);
    put(0, 3);                                           // HARQ process
    put(0, 1);                                           // new data indicator
    put((unsigned)al.rv_idx, 2);
    put(common ? 1u : 0u, 2);                            // TPC; for SI/P/RA-RNTI its LSB selects N_PRB^1A = 3
    if (n == 12 || n == 14 || n == 16 || n == 20 || n == 24 || n == 26 || n == 32 || n == 40 || n == 44 || n == 56) put(0, 1);
    // CRC16 masked with the RNTI
    const unsigned crc = crc16_host(dci, n) ^ (unsigned)(al.rnti & 0xffff);
    uint8_t c[80], d[240];
    memcpy(c, dci, n);
    for (int i = 0; i < 16; i++) c[n + i] = (uint8_t)((crc >> (15 - i)) & 1u);
    conv_encode_tb(c, n + 16, d);
    rate_match_cc(d, n + 16, 288, e288);
    return common ? kTbsNprb3[mcs] : -1;                 // the reference's receiver handles SI/P/RA-RNTI only (dci_1a_unpack, :13273)
}
extern "C" int lteb200_dci_1a_tbs(int mcs) { return (mcs >= 0 && mcs <= 26) ? kTbsNprb3[mcs] : -1; }
// HOST, no device needed: the coded DCI as the transmitter places it (for inspection and CPU tests)
extern "C" int lteb200_dci_1a_encode(const lteb200_alloc *alloc, int mcs, int N_rb_dl, uint8_t *e288)
{
    if (!alloc || !e288 || N_rb_dl < 6 || N_rb_dl > 110) return -1;
    return dci_1a_encode_host(*alloc, mcs, N_rb_dl, e288);
}

// 24 MIB bits -> 1920 coded, rate matched and scrambled bits (4 x 480: one row per radio frame of the 40 ms period)
static void bch_encode_host(const uint8_t *mib24, int N_ant, int N_id_cell, uint8_t *out1920)
{
    uint8_t c[40], d[120];
    const unsigned mask = N_ant == 1 ? 0x0000u : (N_ant == 2 ? 0xFFFFu : 0x5555u);          // bch_channel_encode, liblte_phy.cc:12521-12546
    const unsigned crc = crc16_host(mib24, 24) ^ mask;
    memcpy(c, mib24, 24);
    for (int i = 0; i < 16; i++) c[24 + i] = (uint8_t)((crc >> (15 - i)) & 1u);
    conv_encode_tb(c, 40, d);
    rate_match_cc(d, 40, 1920, out1920);
    std::vector<uint8_t> s(1920);
    prs_c_host((uint32_t)N_id_cell, 1920, s.data());
    for (int i = 0; i < 1920; i++) out1920[i] ^= s[i];
}

// HOST, no device needed
extern "C" int lteb200_bch_encode(const uint8_t *mib24, int N_ant, int N_id_cell, uint8_t *out1920)
{
    if (!mib24 || !out1920 || !(N_ant == 1 || N_ant == 2 || N_ant == 4) || N_id_cell < 0 || N_id_cell > 503) return LTEB200_ERR_ARGS;
    bch_encode_host(mib24, N_ant, N_id_cell, out1920);
    return LTEB200_OK;
}

void gen_pss(int N_id_2, float2 *out);                                   // sync.cu (generate_pss / generate_sss twins)
void gen_sss(int N_id_1, int N_id_2, float2 *s0, float2 *s5);

// Control region, synchronisation signals and PBCH into the grids of n_sf subframes.  grid: port-0 plane of subframe sf at grid + sf * sf_stride
// floats2, port p at + p * port_stride.  Called by lteb200_dl_transmit_frames_batch (tx.cu).
int control_tx_launch(lteb200_ctx *ctx, float2 *grid, size_t sf_stride, size_t port_stride, const int *d_sn, const int *h_sn, int n_sf, int N_rb_dl,
                      int cfi, const lteb200_tx_control *ctl, const lteb200_alloc *allocs, int n_alloc, cudaStream_t stream)
{
    const int N_id_cell = ctl->N_id_cell, N_ant = ctl->N_ant;
    if (ctl->pdcch && N_ant != 1) { lteb200_set_error(ctx, "dl_transmit: PCFICH/PDCCH for one antenna port only (as the receive side)"); return LTEB200_ERR_UNSUPPORTED; }
    if (ctl->pdcch && (n_alloc > 4 || cfi < 1 || cfi > 3)) { lteb200_set_error(ctx, "dl_transmit: at most 4 DCIs (common search space, aggregation level 4), CFI 1..3"); return LTEB200_ERR_ARGS; }
    std::lock_guard<std::mutex> lk(g_ctl_mu);
    CtrlTxParams P;
    memset(&P, 0, sizeof(P));
    P.grid = grid; P.sf_stride = sf_stride; P.port_stride = port_stride; P.nsc = 12 * N_rb_dl; P.cfi = cfi; P.N_id_cell = N_id_cell; P.N_ant = N_ant;
    P.subfr_num = d_sn; P.do_sync = ctl->sync; P.do_pdcch = ctl->pdcch;
    StreamScope scope;
    std::vector<uint8_t> host_blob;
    // ---- DCIs
    std::vector<uint8_t> dci_e((size_t)288 * (n_alloc > 0 ? n_alloc : 1), 0);
    if (ctl->pdcch) {
        PdcchTables T;
        int rc = get_pdcch_tables(ctx, N_rb_dl, N_id_cell, 1, ctl->phich_res, &T);
        if (rc) return rc;
        const int n_symbs = cfi + (N_rb_dl <= 10 ? 1 : 0);
        P.cce_re = T.cce_re + (n_symbs - 1) * 576; P.pcfich_re = T.pcfich_re; P.c_pdcch = T.c_pdcch; P.c_pcfich = T.c_pcfich;
        for (int a = 0; a < n_alloc; a++) {
            const int tbs = dci_1a_encode_host(allocs[a], ctl->mcs[a], N_rb_dl, dci_e.data() + 288 * a);
            if (tbs < 0 || tbs != allocs[a].tbs || allocs[a].mod_type != LTEB200_MOD_QPSK) {
                lteb200_set_error(ctx, "dl_transmit: allocation %d cannot be announced by a DCI 1A the reference's receiver understands (SI/P/RA-RNTI, contiguous PRBs, QPSK, "
                                       "tbs = table 7.1.7.2.1-1[mcs][N_PRB = 3] = %d)", a, tbs);
                return LTEB200_ERR_ARGS;
            }
        }
        P.n_dci = n_alloc;
    }
    // ---- PBCH rows: frame f of the batch starts at the f-th subframe number 0 (f = 0: the possibly partial first frame)
    std::vector<int32_t> bch_row(n_sf, -1);
    std::vector<uint8_t> bch_e;
    if (ctl->pbch) {
        if (!ctl->mib_bits) { lteb200_set_error(ctx, "dl_transmit: pbch without mib_bits"); return LTEB200_ERR_ARGS; }
        int frame = 0, group_mib = 0;
        bool have = false;
        uint8_t cur[1920];
        for (int i = 0; i < n_sf; i++) {
            if (i > 0 && h_sn[i] == 0 && h_sn[i - 1] != 0) frame++;
            if (h_sn[i] != 0) continue;
            const int sfn = ctl->sfn_first + frame;
            // the reference encodes when its buffer is empty (first call, or after the frame with sfn mod 4 == 3) from the MIB it is given then
            if (!have) { group_mib = frame; bch_encode_host(ctl->mib_bits + 24 * (size_t)group_mib, N_ant, N_id_cell, cur); have = true; }
            bch_row[i] = (int32_t)(bch_e.size() / 480);
            bch_e.insert(bch_e.end(), cur + (sfn % 4) * 480, cur + (sfn % 4) * 480 + 480);
            if (sfn % 4 == 3) have = false;
        }
    }
    // ---- sync sequences
    std::vector<float2> sync(3 * 62);
    if (ctl->sync) { gen_pss(N_id_cell % 3, sync.data()); gen_sss(N_id_cell / 3, N_id_cell % 3, sync.data() + 62, sync.data() + 124); }
    uint8_t *d_dci = nullptr, *d_bch = nullptr;
    int32_t *d_row = nullptr;
    float2 *d_sync = nullptr;
    LTE_CUDA_OK(ctx, scope.alloc(&d_dci, dci_e.size(), stream));
    LTE_CUDA_OK(ctx, scope.alloc(&d_bch, bch_e.size() ? bch_e.size() : 16, stream));
    LTE_CUDA_OK(ctx, scope.alloc(&d_row, sizeof(int32_t) * n_sf, stream));
    LTE_CUDA_OK(ctx, scope.alloc(&d_sync, sizeof(float2) * sync.size(), stream));
    LTE_CUDA_OK(ctx, cudaMemcpyAsync(d_dci, dci_e.data(), dci_e.size(), cudaMemcpyHostToDevice, stream));
    if (bch_e.size()) LTE_CUDA_OK(ctx, cudaMemcpyAsync(d_bch, bch_e.data(), bch_e.size(), cudaMemcpyHostToDevice, stream));
    LTE_CUDA_OK(ctx, cudaMemcpyAsync(d_row, bch_row.data(), sizeof(int32_t) * n_sf, cudaMemcpyHostToDevice, stream));
    LTE_CUDA_OK(ctx, cudaMemcpyAsync(d_sync, sync.data(), sizeof(float2) * sync.size(), cudaMemcpyHostToDevice, stream));
    P.dci_e = d_dci; P.bch_e = d_bch; P.bch_row = d_row; P.sync = d_sync;
    ctrl_tx_kernel<<<n_sf, 128, 0, stream>>>(P);
    ctx->launches += 1;
    LTE_CUDA_OK(ctx, cudaGetLastError());
    LTE_CUDA_OK(ctx, cudaStreamSynchronize(stream));      // the pageable host vectors above must outlive their copies
    return LTEB200_OK;
}

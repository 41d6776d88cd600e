// "Mode B" turbo decoder: n_iter-iteration scaled max-log-MAP over code-block batches (BASELINE.json north_star).
//
// The reference has no such decoder (its turbo_decode is three Viterbi passes and a vote); this is
// the north-star kernel.  It decodes the reference's own code -- RSC g0 = 1+D^2+D^3, g1 = 1+D+D^3
// (liblte_phy.cc:10863-10865), tail multiplexing of liblte_phy.cc:10575-10586, and the reference's QPP
// interleaver including its uint32 wrap (liblte_phy.cc:10982) -- so its hard bits equal
// turbo_decode()'s on every input both decode correctly.  Its numeric definition is integer-only and is
// restated on the CPU in oracle/port/turbo_mlm.cc (the bit-exact twin the tests compare against).
//
// B200 mapping: one CTA per code block, everything of the block resident in shared memory for all iterations
// (LLRs int8, extrinsics int16, forward metrics of every step int16, interleaver tables: 198 KB at K = 6144, so one CTA of
// 384 threads per SM); one thread per 16-step window holds the 8 forward/backward state metrics in
// registers (int32, DPX add-max), windows are coupled by state-metric propagation from the previous
// iteration, so all K/32 windows of a block advance in parallel without warm-up recursions.  Shared-memory
// arrays use a window-transposed layout pos(k) = (k mod 32)*nw + k/32 so that at every step the threads of a
// warp touch consecutive addresses (conflict free).  No tensor cores (no contraction on this path).
#include "internal.h"
#include "pdsch_dev.cuh"

#ifndef MLM_WIN
#define MLM_WIN 16   // window length (steps per thread); oracle/port/turbo_mlm.cc must use the same value
#endif

namespace {

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) of the per-block-size tables into shared memory, completion on an mbarrier: the copy
// engine stages the 37 KB of interleaver tables while the CTA's threads are busy with the input phase.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!done);
}

// the mbarrier exists only in the kernels that use the bulk copy
template <bool FUSED> __device__ __forceinline__ uint64_t *table_barrier() { __shared__ __align__(8) uint64_t bar; return &bar; }
template <> __device__ __forceinline__ uint64_t *table_barrier<true>() { return nullptr; }

constexpr int NEG = -8000;
constexpr int LE_LIM = 1023, LA_LIM = 1023;

__host__ __device__ constexpr int tr_a(int S, int u) { return u ^ ((S >> 1) & 1) ^ (S & 1); }
__host__ __device__ constexpr int tr_par(int S, int u) { return tr_a(S, u) ^ ((S >> 2) & 1) ^ (S & 1); }
__host__ __device__ constexpr int tr_nxt(int S, int u) { return (tr_a(S, u) << 2) | (S >> 1); }

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return max(lo, min(v, hi)); }
__device__ __forceinline__ int qf(float v, float maxv) { return __float2int_rz(__fdiv_rn(__fmul_rn(v, 127.0f), maxv)); }  // NaN -> 0

// two int16 lanes per register: lo = bits 0..15, hi = bits 16..31
__device__ __forceinline__ uint32_t pk(int lo, int hi) { return __byte_perm((uint32_t)lo, (uint32_t)hi, 0x5410); }
__device__ __forceinline__ uint32_t lolo(uint32_t x, uint32_t y) { return __byte_perm(x, y, 0x5410); }
__device__ __forceinline__ uint32_t hihi(uint32_t x, uint32_t y) { return __byte_perm(x, y, 0x7632); }
__device__ __forceinline__ int lo16(uint32_t x) { return (int)(short)(x & 0xffffu); }
__device__ __forceinline__ int hi16(uint32_t x) { return (int)x >> 16; }
__device__ __forceinline__ uint32_t add2(uint32_t a, uint32_t b) { return __vadd2(a, b); }                       // VIADD.16x2
__device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }                      // VIMNMX.S16x2
__device__ __forceinline__ uint32_t addmax2(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2(a, b, c); }  // VIADDMNMX.S16x2

// State metrics of the 8-state trellis, two states per register.
//   forward  (alpha): e01 = (a0,a2)  e23 = (a4,a6)  o01 = (a1,a3)  o23 = (a5,a7)
//   backward (beta) : ba  = (b0,b5)  bb  = (b2,b7)  bc  = (b4,b1)  bd  = (b6,b3)
// Predecessors of state d are 2(d&3) and 2(d&3)+1, so the even/odd split makes every add-compare-select a vertical
// 16x2 operation; one byte-permute per register restores the layout for the next step.
struct Pass {
    const int8_t *Ls;     // systematic LLRs (window-transposed, natural order)
    const int16_t *LeSrc; // extrinsic of the OTHER constituent decoder (its own order)
    const uint16_t *piT;  // second decoder: position of pi(k) for the position of k
    const uint16_t *cs, *cp;  // first decoder: CSR lists of the positions that de-interleave onto a position
    const int8_t *Lp;
    int16_t *LeOut;
    uint4 *st;            // [WIN][nw] forward metrics of every step (raw, 8 x int16)
    uint4 *bA, *bB;       // [nw] window-boundary metrics carried to the next iteration
    int xt[3], zt[3];
};

// NT: compile-time row stride of every window-transposed array (the largest window count, K = 6144), so that after full
// unrolling every shared-memory access of step i is `base + constant`.
template <int WIN, int NT, bool FWD_NORM>
__device__ __forceinline__ void fwd_step(uint32_t &e01, uint32_t &e23, uint32_t &o01, uint32_t &o23, int ls, int lp)
{
    // the state vector is re-referenced to alpha(0) on every 4th step only: four steps add at most 4*1277 to a spread of 3831
    const int dl = FWD_NORM ? lo16(e01) : 0;
    const int g11 = -dl, g01 = ls - dl, g10 = lp - dl, g00 = g01 + lp;
    const uint32_t Ga = pk(g00, g10), Gb = pk(g11, g01), Gc = pk(g01, g11), Gd = pk(g10, g00);
    const uint32_t n01 = addmax2(e01, Ga, add2(o01, Gb));
    const uint32_t n23 = addmax2(e23, Gc, add2(o23, Gd));
    const uint32_t n45 = addmax2(e01, Gb, add2(o01, Ga));
    const uint32_t n67 = addmax2(e23, Gd, add2(o23, Gc));
    e01 = lolo(n01, n23); o01 = hihi(n01, n23); e23 = lolo(n45, n67); o23 = hihi(n45, n67);
}

template <bool BWD_NORM>
__device__ __forceinline__ int bwd_step(uint32_t &ba, uint32_t &bb, uint32_t &bc, uint32_t &bd, const uint4 al, int ls, int lp)
{
    const int dl = BWD_NORM ? lo16(ba) : 0;                 // beta'(0)
    const int g11 = -dl, g01 = ls - dl, g10 = lp - dl, g00 = g01 + lp;
    const uint32_t G0a = pk(g00, g01), G0b = pk(g01, g00), G1a = pk(g11, g10), G1b = pk(g10, g11);
    // gamma + beta' of the 16 transitions, grouped by the layout of their source state
    const uint32_t x0e01 = add2(G0a, ba), x0e23 = add2(G0b, bb), x0o01 = add2(G0a, bc), x0o23 = add2(G0b, bd);
    const uint32_t x1e01 = add2(G1a, bc), x1e23 = add2(G1b, bd), x1o01 = add2(G1a, ba), x1o23 = add2(G1b, bb);
    const uint32_t t0 = addmax2(al.w, x0o23, addmax2(al.z, x0o01, addmax2(al.y, x0e23, add2(al.x, x0e01))));
    const uint32_t t1 = addmax2(al.w, x1o23, addmax2(al.z, x1o01, addmax2(al.y, x1e23, add2(al.x, x1e01))));
    const uint32_t mm = max2(lolo(t0, t1), hihi(t0, t1));   // (max over u=0 transitions, max over u=1 transitions)
    const uint32_t ne01 = max2(x0e01, x1e01), ne23 = max2(x0e23, x1e23), no01 = max2(x0o01, x1o01), no23 = max2(x0o23, x1o23);
    ba = lolo(ne01, no23); bb = hihi(ne01, no23); bc = lolo(ne23, no01); bd = hihi(ne23, no01);
    const int L = lo16(mm) - hi16(mm);
    return clampi(((L - ls) * 3) >> 2, -LE_LIM, LE_LIM);
}

// systematic + a-priori LLR of position p, computed where it is consumed (no intermediate array, no extra barrier):
//   second decoder: interleave          Lsa2(p) = Ls[q] + Le1[q],  q = position of pi(k)
//   first decoder : de-interleave+combine Lsa1(p) = Ls[p] + clamp(sum of Le2 over the positions that map onto p)
// 0..3 contributions per position for the reference's wrapped interleaver (exactly 1 for every spec-correct block size):
// the first three are fetched unconditionally (independent loads), the rare rest in a loop.
__device__ __forceinline__ int la_first(const Pass &P, int p)
{
    const int t0 = P.cs[p], cnt = (int)P.cs[p + 1] - t0;
    const int i0 = P.cp[t0], i1 = P.cp[t0 + 1], i2 = P.cp[t0 + 2];
    const int v0 = P.LeSrc[i0], v1 = P.LeSrc[i1], v2 = P.LeSrc[i2];
    int acc = (cnt > 0 ? v0 : 0) + (cnt > 1 ? v1 : 0) + (cnt > 2 ? v2 : 0);
#pragma unroll 1
    for (int t = t0 + 3; t < t0 + cnt; t++) acc += (int)P.LeSrc[P.cp[t]];
    return clampi(acc, -LA_LIM, LA_LIM);
}
template <bool SECOND>
__device__ __forceinline__ int lsa_of(const Pass &P, int p)
{
    if (SECOND) { const int q = P.piT[p]; return (int)P.Ls[q] + (int)P.LeSrc[q]; }
    return (int)P.Ls[p] + la_first(P, p);
}

template <int WIN, int NT, bool SECOND>
__device__ __forceinline__ void dec_pass(const Pass &P, int w, int nw, int K)
{
    const bool active = w < nw;
    const int n = active ? min(WIN, K - WIN * w) : 0;
    uint32_t e01 = 0, e23 = 0, o01 = 0, o23 = 0, ba = 0, bb = 0, bc = 0, bd = 0;
    if (active) {
        if (w == 0) { e01 = pk(0, NEG); e23 = pk(NEG, NEG); o01 = e23; o23 = e23; }
        else { const uint4 v = P.bA[w]; e01 = v.x; e23 = v.y; o01 = v.z; o23 = v.w; }
        if (w == nw - 1) {
            // beta at time K: three forced termination steps (u = feedback, so the register input is 0) ending in state 0
            int b[8];
            b[0] = 0;
#pragma unroll
            for (int s = 1; s < 8; s++) b[s] = NEG;
#pragma unroll
            for (int t = 2; t >= 0; t--) {
                int nb[8];
                const int ls = P.xt[t], lp = P.zt[t], b0 = b[0];
#pragma unroll
                for (int S = 0; S < 8; S++) {
                    const int u = ((S >> 1) & 1) ^ (S & 1);
                    const int g = (u == 0 ? ls : 0) + (tr_par(S, u) == 0 ? lp : 0);
                    nb[S] = max(g - b0 + b[tr_nxt(S, u)], NEG);
                }
#pragma unroll
                for (int S = 0; S < 8; S++) b[S] = nb[S];
            }
            ba = pk(b[0], b[5]); bb = pk(b[2], b[7]); bc = pk(b[4], b[1]); bd = pk(b[6], b[3]);
        } else { const uint4 v = P.bB[w]; ba = v.x; bb = v.y; bc = v.z; bd = v.w; }
    }
    __syncthreads();  // every boundary of the previous iteration is read before any is overwritten

    const int8_t *Lp = P.Lp + w;
    uint4 *st = P.st + w;
    int16_t *Le = P.LeOut + w;
    if (n == WIN) {
        // ---- full window: straight-line code, all addresses are base + constant.  The window's inputs are fetched
        // up front (they do not depend on the recursion) so that no shared-memory latency sits on the dependent chain.
        uint32_t in[WIN];
#pragma unroll
        for (int i = 0; i < WIN; i++) in[i] = pk(lsa_of<SECOND>(P, i * NT + w), (int)Lp[i * NT]);   // (Lsa, Lp)
#pragma unroll
        for (int i = 0; i < WIN; i++) {
            st[i * NT] = make_uint4(e01, e23, o01, o23);
            if ((i & 3) == 0) fwd_step<WIN, NT, true>(e01, e23, o01, o23, lo16(in[i]), hi16(in[i]));
            else fwd_step<WIN, NT, false>(e01, e23, o01, o23, lo16(in[i]), hi16(in[i]));
        }
        if (w + 1 < nw) P.bA[w + 1] = make_uint4(e01, e23, o01, o23);
        uint4 a0 = st[(WIN - 1) * NT], a1 = st[(WIN - 2) * NT];                          // forward metrics, two steps ahead
#pragma unroll
        for (int i = WIN - 1; i >= 0; i--) {
            const uint4 al = a0;
            a0 = a1;
            if (i >= 2) a1 = st[(i - 2) * NT];
            if ((i & 3) == 3) Le[i * NT] = (int16_t)bwd_step<true>(ba, bb, bc, bd, al, lo16(in[i]), hi16(in[i]));
            else Le[i * NT] = (int16_t)bwd_step<false>(ba, bb, bc, bd, al, lo16(in[i]), hi16(in[i]));
        }
        if (w > 0) P.bB[w - 1] = make_uint4(ba, bb, bc, bd);
    } else if (n > 0) {
        // ---- the short last window of block sizes that are not a multiple of WIN
        for (int i = 0; i < n; i++) {
            st[i * NT] = make_uint4(e01, e23, o01, o23);
            const int ls = lsa_of<SECOND>(P, i * NT + w);
            if ((i & 3) == 0) fwd_step<WIN, NT, true>(e01, e23, o01, o23, ls, Lp[i * NT]);
            else fwd_step<WIN, NT, false>(e01, e23, o01, o23, ls, Lp[i * NT]);
        }
        for (int i = n - 1; i >= 0; i--) {
            const int ls = lsa_of<SECOND>(P, i * NT + w);
            if ((i & 3) == 3) Le[i * NT] = (int16_t)bwd_step<true>(ba, bb, bc, bd, st[i * NT], ls, Lp[i * NT]);
            else Le[i * NT] = (int16_t)bwd_step<false>(ba, bb, bc, bd, st[i * NT], ls, Lp[i * NT]);
        }
        if (w > 0) P.bB[w - 1] = make_uint4(ba, bb, bc, bd);
    }
    __syncthreads();
}

// piT  [WIN*NT] : pos(pi(k)) stored at pos(k), pos(k) = (k mod WIN)*NT + k/WIN, NT = 6144/WIN    (global, per K)
// csr_start [NP+1], csr_pos [K] : for POSITION p = pos(j), the positions pos(i) of all i with pi(i) == j   (global, per K)
// FUSED == false: input = rate-unmatched LLRs llr[n_cb][3(K+4)], output = decoded bits out[n_cb][K].
// FUSED == true : the whole PDSCH receive of one transport block in one kernel -- RE extraction, ZF, demapping and
//                 descrambling into shared memory (aliasing the forward-metric store, which is idle until the first pass),
//                 rate unmatching through the gather map straight into each thread's window, decoding, then CRC24A and
//                 transport-block extraction from the decisions.  Block index = local_alloc * n_sf + subframe.
template <int WIN, int NT, bool FUSED>
__global__ void __launch_bounds__(NT)
turbo_mlm_kernel(const float *__restrict__ llr, int K, int n_iter, const uint16_t *__restrict__ piT,
                 const uint16_t *__restrict__ csr_start, const uint16_t *__restrict__ csr_pos, uint8_t *__restrict__ out,
                 const __grid_constant__ FusedParams fp)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // NT: row stride of the window-transposed arrays = the largest window count of this size class (compile-time)
    const int nw = (K + WIN - 1) / WIN;
    constexpr int NP = WIN * NT;     // padded length
    uint4 *st = (uint4 *)smem_raw;                                  // 16*NP bytes
    uint4 *bA1 = (uint4 *)(smem_raw + (size_t)16 * NP);             // 4 x nw uint4
    uint4 *bB1 = bA1 + NT, *bA2 = bB1 + NT, *bB2 = bA2 + NT;
    int16_t *Le1 = (int16_t *)(bB2 + NT);                           // extrinsics of the two constituent decoders
    int16_t *Le2 = Le1 + NP;
    uint16_t *s_piT = (uint16_t *)(Le2 + NP);                       // interleaver and its CSR inverse, staged once per block
    // standalone decoder: one contiguous image of the global table blob (bulk copy); fused receive kernel: the tables are staged by the
    // threads after the input phase (measured faster there: the copy engine's transfer competes with the gather and its completion is
    // waited for at a barrier, profiles/), with the tighter NP+2 spacing
    constexpr int CSP = FUSED ? 2 : 8;
    uint16_t *s_cs = s_piT + NP, *s_cp = s_cs + (NP + CSP);
    int8_t *Ls = (int8_t *)(s_cp + NP + CSP);
    int8_t *Lp1 = Ls + NP, *Lp2 = Lp1 + NP;
    __shared__ float s_red[32];
    __shared__ int s_tail[12];

    const int cb = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    constexpr uint32_t TAB_BYTES = (3 * NP + 16) * 2;
    uint64_t *s_bar = table_barrier<FUSED>();
    if constexpr (!FUSED) {
        if (tid == 0) {
            mbar_init(s_bar, 1);
            mbar_expect_tx(s_bar, TAB_BYTES);
            bulk_copy_g2s(s_piT, piT, TAB_BYTES, s_bar);         // piT | csr_start (NP+8) | csr_pos (NP+8), laid out like this in HBM
        }
    }
    const float *d = llr + (size_t)cb * 3 * (K + 4);
    // ---- input: thread w owns steps 16w..16w+15 = 48 consecutive floats (192 B, 16-byte aligned), read once as float4.
    // max |d| over all 3(K+4) inputs (sentinel -> 0) first, then quantisation straight from registers.
    float v[3 * WIN];
    float m = 0.f;
    const bool own = tid < nw;
    const int nown = own ? min(WIN, K - WIN * tid) : 0;
    float tail_v = 0.f;
    if (FUSED) {
        const int a_loc = cb / fp.n_sf, sf = cb - a_loc * fp.n_sf;
        const FusedAlloc &A = fp.al[a_loc];
        const int sn = fp.subfr_num[sf];
        const int cls = sn == 0 ? 0 : (sn == 5 ? 1 : 2);
        const int M = A.M[cls], E = M * A.Qm;
        float *soft = reinterpret_cast<float *>(st);                       // <= 20,700 floats, the metric store is idle now
        const float2 *gr = fp.grid + (size_t)sf * 16 * fp.nsc, *ch = fp.ce + (size_t)sf * fp.N_ant * 14 * fp.nsc;
        const int *ix = A.re_idx + cls * A.idx_stride;
        const uint8_t *cs = A.cseq + (size_t)sn * A.c_stride;
        if (fp.N_ant != 1) {
            // transmit diversity: one thread per group of N_ant consecutive resource elements (M is a multiple of N_ant here, the launcher checks)
            const size_t ps = (size_t)14 * fp.nsc;
            auto emit = [&](int r, float2 x) {
                int b[6];
                demap_bits(x.x, x.y, A.mod, b);
                for (int t = 0; t < A.Qm; t++) soft[r * A.Qm + t] = __fmul_rn((float)b[t], (float)(1 - 2 * (int)cs[r * A.Qm + t]));
            };
            for (int r = tid * fp.N_ant; r < M; r += nt * fp.N_ant) {
                if (fp.N_ant == 2) {
                    float2 x0, x1;
                    sfbc2(gr[ix[r]], gr[ix[r + 1]], ch[ix[r]], ch[ps + ix[r]], x0, x1);
                    emit(r, x0); emit(r + 1, x1);
                } else {
                    const float2 y[4] = {gr[ix[r]], gr[ix[r + 1]], gr[ix[r + 2]], gr[ix[r + 3]]};
                    float2 x[4];
                    sfbc4(y, ch[ix[r]], ch[ps + ix[r + 2]], ch[2 * ps + ix[r]], ch[3 * ps + ix[r + 2]], x);
                    for (int q = 0; q < 4; q++) emit(r + q, x[q]);
                }
            }
        } else {
        // every thread handles up to RE_IT resource elements; indices, grid and estimate values are all fetched before the
        // arithmetic starts, so the global-memory latency is paid once, not per element
        constexpr int RE_IT = 4;
        for (int r0 = tid; r0 < M; r0 += RE_IT * nt) {
            int idx[RE_IT];
            float2 yv[RE_IT], hv[RE_IT];
#pragma unroll
            for (int u = 0; u < RE_IT; u++) idx[u] = (r0 + u * nt < M) ? ix[r0 + u * nt] : ix[r0];
#pragma unroll
            for (int u = 0; u < RE_IT; u++) { yv[u] = gr[idx[u]]; hv[u] = ch[idx[u]]; }
#pragma unroll
            for (int u = 0; u < RE_IT; u++) {
                const int r = r0 + u * nt;
                if (r < M) {
                    float xr, xi;
                    zf_1port(yv[u], hv[u], xr, xi);
                    int b[6];
                    demap_bits(xr, xi, A.mod, b);
                    for (int t = 0; t < A.Qm; t++) soft[r * A.Qm + t] = __fmul_rn((float)b[t], (float)(1 - 2 * (int)cs[r * A.Qm + t]));
                }
            }
        }
        }
        __syncthreads();
        // rate unmatching (liblte_phy.cc:11403-11415): first visit assigns, repeats are added in visiting order
        auto unmatch = [&](int fk) -> float {
            float w = LTEB200_RX_NULL_BIT;
            if (fk >= 0)
                for (int k = fk; k < E; k += A.period) {
                    const float x = soft[k];
                    if (w == LTEB200_RX_NULL_BIT) w = x; else if (x != LTEB200_RX_NULL_BIT) w = __fadd_rn(w, x);
                }
            return w;
        };
        if (nown == WIN) {
            const int4 *fk4 = reinterpret_cast<const int4 *>(A.first_k + 3 * WIN * tid);
#pragma unroll
            for (int j = 0; j < 3 * WIN / 4; j++) {
                const int4 f = fk4[j];
                v[4 * j] = unmatch(f.x); v[4 * j + 1] = unmatch(f.y); v[4 * j + 2] = unmatch(f.z); v[4 * j + 3] = unmatch(f.w);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 3 * WIN; j++) v[j] = (j < 3 * nown) ? unmatch(A.first_k[3 * WIN * tid + j]) : 0.f;
        }
        if (tid < 12) tail_v = unmatch(A.first_k[3 * K + tid]);
        __syncthreads();                                                   // soft[] is dead from here on
    } else if (nown == WIN) {
        const float4 *src = reinterpret_cast<const float4 *>(d + 3 * WIN * tid);
#pragma unroll
        for (int j = 0; j < 3 * WIN / 4; j++) {
            const float4 t = src[j];
            v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 3 * WIN; j++) v[j] = (j < 3 * nown) ? d[3 * WIN * tid + j] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 3 * WIN; j++) {
        if (v[j] == LTEB200_RX_NULL_BIT) v[j] = 0.f;
        m = fmaxf(m, fabsf(v[j]));
    }
    if (tid < 12) {                              // the 12 tail values take part in the maximum
        if (!FUSED) tail_v = d[3 * K + tid];
        if (tail_v == LTEB200_RX_NULL_BIT) tail_v = 0.f;
        m = fmaxf(m, fabsf(tail_v));
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((tid & 31) == 0) s_red[tid >> 5] = m;
    __syncthreads();
    if (tid < 32) {
        float t = tid < (nt >> 5) ? s_red[tid] : 0.f;
        for (int o = 16; o > 0; o >>= 1) t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, o));
        if (tid == 0) s_red[0] = t;
    }
    __syncthreads();
    const float maxv = s_red[0];
    // ---- quantise into the window-transposed arrays (thread w writes column w: conflict free); padding stays zero
    for (int p = tid; p < NP; p += nt) { Le1[p] = 0; Le2[p] = 0; }
    if (tid < NT) {
#pragma unroll
        for (int i = 0; i < WIN; i++) {
            const int p = i * NT + tid;
            const int q0 = (i < nown) ? qf(v[3 * i], maxv) : 0, q1 = (i < nown) ? qf(v[3 * i + 1], maxv) : 0, q2 = (i < nown) ? qf(v[3 * i + 2], maxv) : 0;
            Ls[p] = (int8_t)q0; Lp1[p] = (int8_t)q1; Lp2[p] = (int8_t)q2;
        }
    }
    if (tid < 12) s_tail[tid] = qf(tail_v, maxv);  // s_tail[3*r + x] = q(d_x[K+r])
    for (int i = tid; i < NT * 4; i += nt) bA1[i] = make_uint4(0, 0, 0, 0);  // bA1,bB1,bA2,bB2 are contiguous
    if constexpr (FUSED) {
        for (int p = tid; p < NP; p += nt) s_piT[p] = piT[p];
        for (int j = tid; j <= NP; j += nt) s_cs[j] = csr_start[j];
        for (int j = tid; j < NP; j += nt) s_cp[j] = j < K ? csr_pos[j] : (uint16_t)0;
        if (tid < 2) s_cp[NP + tid] = 0;                                    // s_cp[t0+1], s_cp[t0+2] are read unconditionally
        __syncthreads();
    } else {
        __syncthreads();
        mbar_wait(s_bar, 0);                                            // tables have landed (zero padding of csr_pos comes with the blob)
    }

    Pass P1, P2;
    P1.Ls = Ls; P1.LeSrc = Le2; P1.piT = s_piT; P1.cs = s_cs; P1.cp = s_cp; P1.Lp = Lp1; P1.LeOut = Le1; P1.st = st; P1.bA = bA1; P1.bB = bB1;
    P2.Ls = Ls; P2.LeSrc = Le1; P2.piT = s_piT; P2.cs = s_cs; P2.cp = s_cp; P2.Lp = Lp2; P2.LeOut = Le2; P2.st = st; P2.bA = bA2; P2.bB = bB2;
    // tail multiplexing (liblte_phy.cc:10575-10586): x = systematic/feedback bits, z = parity bits of the 3 tail steps
    P1.xt[0] = s_tail[0]; P1.xt[1] = s_tail[2]; P1.xt[2] = s_tail[4];
    P1.zt[0] = s_tail[1]; P1.zt[1] = s_tail[3]; P1.zt[2] = s_tail[5];
    P2.xt[0] = s_tail[6]; P2.xt[1] = s_tail[8]; P2.xt[2] = s_tail[10];
    P2.zt[0] = s_tail[7]; P2.zt[1] = s_tail[9]; P2.zt[2] = s_tail[11];

    for (int it = 0; it < n_iter; it++) {
        dec_pass<WIN, NT, false>(P1, tid, nw, K);
        dec_pass<WIN, NT, true>(P2, tid, nw, K);
    }
    if (!FUSED) {
        for (int j = tid; j < K; j += nt) {
            const int p = (j % WIN) * NT + (j / WIN);
            out[(size_t)cb * K + j] = ((int)Ls[p] + la_first(P1, p) + (int)Le1[p] >= 0) ? 0 : 1;
        }
    } else {
        // dlsch_channel_decode's tail (liblte_phy.cc:12841-12869): drop the F filler bits, CRC24A over [a | p], emit a.
        // The remainder of the tbs+24 bit message is the XOR of x^(distance from the end) mod g over its set bits.
        const int a_loc = cb / fp.n_sf, sf = cb - a_loc * fp.n_sf;
        const FusedAlloc &A = fp.al[a_loc];
        const int n = A.tbs + 24;
        uint8_t *o = fp.tb_out + ((size_t)sf * fp.n_alloc_total + A.a_idx) * fp.tb_stride;
        uint32_t rem = 0;
        if (tid < nw) {
            // thread w decides its own 16 steps (conflict-free column reads), j = 16w + i
            const int nown_e = min(WIN, K - WIN * tid);
            unsigned long long packed[2] = {0ull, 0ull};                  // one byte per bit, 16 bytes
#pragma unroll
            for (int i = 0; i < WIN; i++) {
                const int p = i * NT + tid;
                const int bit = (i < nown_e && ((int)Ls[p] + la_first(P1, p) + (int)Le1[p] < 0)) ? 1 : 0;
                packed[i >> 3] |= (unsigned long long)bit << (8 * (i & 7));
            }
            const int idx0 = WIN * tid - A.F;                              // index of this thread's first bit inside [a | p]
            const bool fast = A.F == 0 && nown_e == WIN && idx0 + WIN <= A.tbs && (fp.tb_stride & 7) == 0 && ((size_t)o & 7) == 0;
            if (fast) {
                reinterpret_cast<unsigned long long *>(o + idx0)[0] = packed[0];
                reinterpret_cast<unsigned long long *>(o + idx0)[1] = packed[1];
            } else {
#pragma unroll
                for (int i = 0; i < WIN; i++) {
                    const int idx = idx0 + i;
                    if (i < nown_e && idx >= 0 && idx < A.tbs) o[idx] = (uint8_t)((packed[i >> 3] >> (8 * (i & 7))) & 1);
                }
            }
#pragma unroll
            for (int i = 0; i < WIN; i++) {
                const int idx = idx0 + i;
                if (i < nown_e && idx >= 0 && ((packed[i >> 3] >> (8 * (i & 7))) & 1)) rem ^= fp.crc_tab[n - 1 - idx];
            }
        }
        for (int off = 16; off > 0; off >>= 1) rem ^= __shfl_xor_sync(0xffffffffu, rem, off);
        __syncthreads();
        if ((tid & 31) == 0) s_tail[tid >> 5] = (int)rem;                  // <= 12 warps
        __syncthreads();
        if (tid == 0) {
            uint32_t r = 0;
            for (int wv = 0; wv < (nt >> 5); wv++) r ^= (uint32_t)s_tail[wv];
            fp.crc_ok[(size_t)sf * fp.n_alloc_total + A.a_idx] = (r & 0xffffffu) == 0 ? 1 : 0;
        }
    }
}

}  // namespace

static int turbo_mlm_launch(lteb200_ctx *ctx, const float *d_llr, int K, int n_cb, int n_iter, uint8_t *d_bits, const FusedParams *fused, cudaStream_t stream);

int turbo_mlm_decode_batch(lteb200_ctx *ctx, const float *d_llr, int K, int n_cb, int n_iter, uint8_t *d_bits, cudaStream_t stream)
{
    return turbo_mlm_launch(ctx, d_llr, K, n_cb, n_iter, d_bits, nullptr, stream);
}

// Fused PDSCH receive: n_cb = fp.n_al * fp.n_sf transport blocks of the same K.
int turbo_mlm_fused_batch(lteb200_ctx *ctx, int K, int n_iter, const FusedParams &fp, cudaStream_t stream)
{
    return turbo_mlm_launch(ctx, nullptr, K, fp.n_al * fp.n_sf, n_iter, nullptr, &fp, stream);
}

static int turbo_mlm_launch(lteb200_ctx *ctx, const float *d_llr, int K, int n_cb, int n_iter, uint8_t *d_bits, const FusedParams *fused, cudaStream_t stream)
{
    if (K < 40 || (K & 7) || K > 6144) { lteb200_set_error(ctx, "turbo: illegal K=%d", K); return LTEB200_ERR_ARGS; }
    if (mlm2_handles(K) && !fused) return turbo_mlm2_launch(ctx, d_llr, nullptr, K, n_cb, n_iter, d_bits, nullptr, stream);   // two windows per thread, register-resident metrics
    if (n_iter < 1) n_iter = 1;
    QppTables qt;
    int rc = lteb200_get_qpp(ctx, K, &qt);  // validates K
    if (rc) return rc;
    constexpr int WIN = MLM_WIN;
    const int nw = (K + WIN - 1) / WIN;
    // size class: shared memory (and therefore blocks per SM) scales with the block size
    const int NT = mlm_size_class(K), NP = WIN * NT;
    MlmTables mt;
    {
        auto it = ctx->mlm.find(K | (ctx->qpp_spec << 16));
        if (it == ctx->mlm.end()) {
            std::vector<uint16_t> pi(K);
            LTE_CUDA_OK(ctx, cudaMemcpy(pi.data(), qt.pi, K * sizeof(uint16_t), cudaMemcpyDeviceToHost));
            auto pos = [&](int k) { return (uint16_t)((k % WIN) * NT + (k / WIN)); };
            std::vector<uint16_t> piT(NP, 0), cs(NP + 1, 0), cp(K);
            for (int k = 0; k < K; k++) piT[pos(k)] = pos(pi[k]);
            std::vector<int> cnt(NP, 0);
            for (int i = 0; i < K; i++) cnt[pos(pi[i])]++;
            for (int q = 0; q < NP; q++) cs[q + 1] = (uint16_t)(cs[q] + cnt[q]);
            std::vector<int> fill(cs.begin(), cs.end() - 1);
            for (int i = 0; i < K; i++) cp[fill[pos(pi[i])]++] = pos(i);   // increasing i inside each list
            // one blob, copied into shared memory by a single bulk copy: piT [NP] | csr_start [NP+8] | csr_pos [NP+8], zero padded
            std::vector<uint16_t> blob((size_t)3 * NP + 16, 0);
            memcpy(blob.data(), piT.data(), (size_t)NP * 2);
            memcpy(blob.data() + NP, cs.data(), (size_t)(NP + 1) * 2);
            for (int j = NP + 1; j < NP + 8; j++) blob[NP + j] = cs[NP];
            memcpy(blob.data() + 2 * NP + 8, cp.data(), (size_t)K * 2);
            LTE_CUDA_OK(ctx, cudaMalloc(&mt.piT, blob.size() * 2));
            LTE_CUDA_OK(ctx, cudaMemcpy(mt.piT, blob.data(), blob.size() * 2, cudaMemcpyHostToDevice));
            mt.csr_start = mt.piT + NP;
            mt.csr_pos = mt.csr_start + NP + 8;
            ctx->mlm[K | (ctx->qpp_spec << 16)] = mt;
        } else mt = it->second;
    }
    const size_t smem = (size_t)16 * NP + (size_t)4 * NT * 16 + (size_t)2 * NP * 2 + (size_t)(3 * NP + (fused ? 4 : 16)) * 2 + (size_t)3 * NP;
    const int threads = ((nw + 31) / 32) * 32;
    FusedParams none;
    memset(&none, 0, sizeof(none));
    const FusedParams &fp = fused ? *fused : none;
#define MLM_LAUNCH(NT_, FUSED_)                                                                                                     \
    do {                                                                                                                            \
        LTE_CUDA_OK(ctx, cudaFuncSetAttribute(turbo_mlm_kernel<WIN, NT_, FUSED_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        turbo_mlm_kernel<WIN, NT_, FUSED_><<<n_cb, threads, smem, stream>>>(d_llr, K, n_iter, mt.piT, mt.csr_start, mt.csr_pos, d_bits, fp); \
    } while (0)
    if (fused) { if (NT == 64) MLM_LAUNCH(64, true); else if (NT == 128) MLM_LAUNCH(128, true); else if (NT == 192) MLM_LAUNCH(192, true); else MLM_LAUNCH(384, true); }
    else       { if (NT == 64) MLM_LAUNCH(64, false); else if (NT == 128) MLM_LAUNCH(128, false); else if (NT == 192) MLM_LAUNCH(192, false); else MLM_LAUNCH(384, false); }
#undef MLM_LAUNCH
    ctx->launches += 1;
    LTE_CUDA_OK(ctx, cudaGetLastError());
    return LTEB200_OK;
}

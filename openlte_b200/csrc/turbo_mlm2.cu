// "Mode B" turbo decoder, second formulation: one trellis STATE per register, two decoder WINDOWS per thread in the two
// int16 lanes of every register, forward metrics of the window resident in REGISTERS.
//
// Numerically this is the decoder defined in oracle/port/turbo_mlm.cc (and implemented window-per-thread in turbo_mlm.cu):
// same quantisation, same 16-step windows with state-metric propagation between iterations, same re-referencing of the
// state vector on every 4th step, same extrinsic scaling and clamps -- the results are bit-identical, and the tests hold both
// kernels to the CPU twin.  What changes is the mapping onto the SM (ncu evidence of the first formulation, round 1: ALU pipe
// 61 %, 30 % of its ALU instructions byte-permutes that re-pair the (even, odd) state registers, 23 % of the shared-memory
// wavefronts bank conflicts of the CSR de-interleaver, one 203 KB CTA per SM):
//   * thread t owns the ADJACENT windows 2t and 2t+1: lane lo of every packed register is window 2t, lane hi is window 2t+1.
//     The add-compare-select of a state is then one VIADD.16x2 + one VIADDMNMX.S16x2 on whole registers -- no permutes at
//     all in the recursions -- and every window-transposed array is read and written as aligned (lo, hi) words;
//   * the code's branch metrics are (0, Lsa, Lp, Lsa+Lp): the four transitions with metric 0 need no add, so a forward step is
//     13 packed ALU instructions for two windows, a backward step with both LLR maxima 36;
//   * forward metrics never touch shared memory: steps 8..15 of the window stay in 64 registers, the backward recursion
//     consumes them, then steps 0..7 are recomputed from the window's start metrics (8 more forward steps, +10 % ALU work) and
//     consumed in turn.  Shared memory per block drops from 203 KB to 92 KB at K = 6144: two blocks per SM;
//   * the interleaver exchange needs ONE random shared-memory access per trellis step in each direction: the first decoder
//     publishes S1 = Lsys + Le1 (gathered through pi by the second), the second scatters Le2 with a shared-memory atomic add
//     into int32 accumulators indexed by pi (the reference's wrapped interleaver maps 0..3 positions onto one; the
//     first decoder reads, clamps and clears its accumulator) -- the CSR lists and their conflicts are gone;
//   * one barrier per half iteration: the boundary records of the NEXT pass are fetched before the barrier that ends this one.
// Block sizes with K mod 32 != 0 (all below 1024) keep the first formulation: a thread's two lanes must both hold a full window.
#include "internal.h"
#include "pdsch_dev.cuh"

namespace {

constexpr int WIN = 16;
constexpr int NEG = -8000;
constexpr int LE_LIM = 1023, LA_LIM = 1023;

__host__ __device__ constexpr int tr_a(int S, int u) { return u ^ ((S >> 1) & 1) ^ (S & 1); }
__host__ __device__ constexpr int tr_par(int S, int u) { return tr_a(S, u) ^ ((S >> 2) & 1) ^ (S & 1); }
__host__ __device__ constexpr int tr_nxt(int S, int u) { return (tr_a(S, u) << 2) | (S >> 1); }

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return max(lo, min(v, hi)); }
__device__ __forceinline__ int qf(float v, float maxv) { return __float2int_rz(__fdiv_rn(__fmul_rn(v, 127.0f), maxv)); }  // NaN -> 0

// packed int16 lanes: lo = window 2t, hi = window 2t+1
__device__ __forceinline__ uint32_t pk(int lo, int hi) { return __byte_perm((uint32_t)lo, (uint32_t)hi, 0x5410); }
__device__ __forceinline__ uint32_t vadd(uint32_t a, uint32_t b) { return __vadd2(a, b); }                          // VIADD.16x2
__device__ __forceinline__ uint32_t vsub(uint32_t a, uint32_t b) { return __vsub2(a, b); }
__device__ __forceinline__ uint32_t vmax(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }                         // VIMNMX.S16x2
__device__ __forceinline__ uint32_t vaddmax(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2(a, b, c); }  // max(a + b, c): VIADDMNMX.S16x2
// PRMT with sign replication (selector bit 3: the byte is replaced by its sign); __byte_perm() masks that bit away, hence the PTX
template <uint32_t SEL> __device__ __forceinline__ uint32_t prmt_s(uint32_t a)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %1, %2;" : "=r"(d) : "r"(a), "n"(SEL));
    return d;
}
__device__ __forceinline__ int lo16(uint32_t x) { return (int)prmt_s<0x9910u>(x); }   // sign-extended low half
__device__ __forceinline__ int hi16(uint32_t x) { return (int)x >> 16; }
// two int8 (bytes 0, 1) -> two sign-extended int16 lanes
__device__ __forceinline__ uint32_t s8x2_to_s16x2(uint32_t v) { return prmt_s<0x9180u>(v); }

// Forward step of both windows.  NORM: the state vector is re-referenced to state 0 first (every 4th step).
template <bool NORM>
__device__ __forceinline__ void fwd_step(uint32_t (&a)[8], uint32_t ls, uint32_t lp)
{
    if (NORM) {
        const uint32_t r = a[0];
#pragma unroll
        for (int s = 1; s < 8; s++) a[s] = vsub(a[s], r);
        a[0] = 0u;
    }
    const uint32_t g00 = vadd(ls, lp);
    const uint32_t n0 = vaddmax(a[0], g00, a[1]);                 // 0 <- (0, u=0, g00) | (1, u=1, 0)
    const uint32_t n4 = vaddmax(a[1], g00, a[0]);                 // 4 <- (1, u=0, g00) | (0, u=1, 0)
    const uint32_t n5 = vaddmax(a[3], lp, vadd(a[2], ls));        // 5 <- (2, u=0, Lsa) | (3, u=1, Lp)
    const uint32_t n1 = vaddmax(a[3], ls, vadd(a[2], lp));        // 1 <- (3, u=0, Lsa) | (2, u=1, Lp)
    const uint32_t n2 = vaddmax(a[5], lp, vadd(a[4], ls));        // 2 <- (4, u=0, Lsa) | (5, u=1, Lp)
    const uint32_t n6 = vaddmax(a[5], ls, vadd(a[4], lp));        // 6 <- (5, u=0, Lsa) | (4, u=1, Lp)
    const uint32_t n7 = vaddmax(a[6], g00, a[7]);                 // 7 <- (6, u=0, g00) | (7, u=1, 0)
    const uint32_t n3 = vaddmax(a[7], g00, a[6]);                 // 3 <- (7, u=0, g00) | (6, u=1, 0)
    a[0] = n0; a[1] = n1; a[2] = n2; a[3] = n3; a[4] = n4; a[5] = n5; a[6] = n6; a[7] = n7;
}

// Backward step of both windows; returns L(u=0) - L(u=1) in both lanes.  al = forward metrics at the step's start.
template <bool NORM>
__device__ __forceinline__ uint32_t bwd_step(uint32_t (&b)[8], const uint32_t (&al)[8], uint32_t ls, uint32_t lp)
{
    if (NORM) {
        const uint32_t r = b[0];
#pragma unroll
        for (int s = 1; s < 8; s++) b[s] = vsub(b[s], r);
        b[0] = 0u;
    }
    const uint32_t g00 = vadd(ls, lp);
    // gamma + beta' of the 16 transitions (x0: input bit 0, x1: input bit 1); the four with branch metric 0 are beta' itself
    const uint32_t x00 = vadd(g00, b[0]), x01 = vadd(g00, b[4]), x02 = vadd(ls, b[5]), x03 = vadd(ls, b[1]);
    const uint32_t x04 = vadd(ls, b[2]), x05 = vadd(ls, b[6]), x06 = vadd(g00, b[7]), x07 = vadd(g00, b[3]);
    const uint32_t x12 = vadd(lp, b[1]), x13 = vadd(lp, b[5]), x14 = vadd(lp, b[6]), x15 = vadd(lp, b[2]);
    const uint32_t x10 = b[4], x11 = b[0], x16 = b[3], x17 = b[7];
    uint32_t L0 = vadd(al[0], x00), L1 = vadd(al[0], x10);
    L0 = vaddmax(al[1], x01, L0); L1 = vaddmax(al[1], x11, L1);
    L0 = vaddmax(al[2], x02, L0); L1 = vaddmax(al[2], x12, L1);
    L0 = vaddmax(al[3], x03, L0); L1 = vaddmax(al[3], x13, L1);
    L0 = vaddmax(al[4], x04, L0); L1 = vaddmax(al[4], x14, L1);
    L0 = vaddmax(al[5], x05, L0); L1 = vaddmax(al[5], x15, L1);
    L0 = vaddmax(al[6], x06, L0); L1 = vaddmax(al[6], x16, L1);
    L0 = vaddmax(al[7], x07, L0); L1 = vaddmax(al[7], x17, L1);
    b[0] = vmax(x00, x10); b[1] = vmax(x01, x11); b[2] = vmax(x02, x12); b[3] = vmax(x03, x13);
    b[4] = vmax(x04, x14); b[5] = vmax(x05, x15); b[6] = vmax(x06, x16); b[7] = vmax(x07, x17);
    return vsub(L0, L1);
}

// ((d * 3) >> 2) of both lanes: the high product word of (lane << 16) * (3 << 14) is (3 lane) >> 2 with an arithmetic shift.  Two IMAD.HI
// (FMA pipe) instead of sign extension, multiply and shift per lane on the ALU pipe.
__device__ __forceinline__ void ext_pair(uint32_t d2, int &eA, int &eB)
{
    eA = __mulhi((int)(d2 << 16), 3 << 14);
    eB = __mulhi((int)(d2 & 0xffff0000u), 3 << 14);
}
constexpr uint32_t LE2 = (uint32_t)LE_LIM | ((uint32_t)LE_LIM << 16), NLE2 = (uint32_t)(-LE_LIM & 0xffff) | ((uint32_t)(-LE_LIM & 0xffff) << 16);
constexpr uint32_t LA2 = (uint32_t)LA_LIM | ((uint32_t)LA_LIM << 16), NLA2 = (uint32_t)(-LA_LIM & 0xffff) | ((uint32_t)(-LA_LIM & 0xffff) << 16);

// Shared-memory image of one code block (NT = row stride in windows of this size class, NT2 = NT/2 threads, NP = 16 NT positions;
// position of trellis step k: (k mod 16) * NT + k / 16, so thread t's two windows are the adjacent columns 2t, 2t+1):
template <int NT>
struct Smem {
    static constexpr int NT2 = NT / 2, NP = WIN * NT;
    // boundary records, lane-packed like the registers: recA[d][1 + t] = forward metrics at the END of windows (2t, 2t+1) of decoder d,
    // recA[d][0] = the block-start vector in its hi lanes; recB[d][t] = backward metrics at the START of windows (2t, 2t+1),
    // recB[d][n_threads] = the termination vector in its lo lanes
    uint4 recA[2][NT2 + 1][2];
    uint4 recB[2][NT2 + 1][2];
    uint4 scr[6][NT2];      // per-thread stash of the inputs of steps 0..7 and the start metrics while steps 15..8 run backward
    int acc[NP];            // sum of the second decoder's extrinsics per natural position (a-priori of the first decoder, unclamped)
    int16_t S1[NP];         // Lsys + Le1 (natural order): a-priori + systematic of the second decoder, read through pi
    uint16_t piT2[NP];      // 2 * position of pi(k), stored at the position of k (byte offset into S1)
    int8_t Ls[NP], Lp1[NP], Lp2[NP];   // quantised channel LLRs; with tail8 one contiguous image (the Q8 input format)
    int8_t tail8[16];                   // q(d_x[K + r]) at [3 r + x], 12 used
};

// Inputs of steps i0..i0+7 of both windows: (Lsa_lo, Lsa_hi) and (Lp_lo, Lp_hi) per step; 24 independent shared-memory loads.
//   first decoder : Lsa = Lsys + clamp(accumulated extrinsics of the second decoder); the accumulator is cleared for the next scatter
//   second decoder: Lsa = S1[pi] (= Lsys + Le1 at the interleaved position), one random 16-bit read per window
template <int NT, bool SECOND>
__device__ __forceinline__ void load_inputs(Smem<NT> &sm, const int t, const int i0, uint32_t (&ls)[8], uint32_t (&lp)[8])
{
    constexpr int NT2 = NT / 2;
    const uint16_t *Lp16 = reinterpret_cast<const uint16_t *>(SECOND ? sm.Lp2 : sm.Lp1) + t + i0 * NT2;
    if (SECOND) {
        const uint32_t *pi32 = reinterpret_cast<const uint32_t *>(sm.piT2) + t + i0 * NT2;
        const char *s1 = reinterpret_cast<const char *>(sm.S1);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t q2 = pi32[i * NT2];
            const int vA = *reinterpret_cast<const int16_t *>(s1 + (q2 & 0xffffu));
            const int vB = *reinterpret_cast<const int16_t *>(s1 + (q2 >> 16));
            ls[i] = pk(vA, vB);
            lp[i] = s8x2_to_s16x2(Lp16[i * NT2]);
        }
    } else {
        const uint16_t *Ls16 = reinterpret_cast<const uint16_t *>(sm.Ls) + t + i0 * NT2;
        int2 *acc2 = reinterpret_cast<int2 *>(sm.acc) + t + i0 * NT2;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int2 s = acc2[i * NT2];
            acc2[i * NT2] = make_int2(0, 0);
            const uint32_t la = vmax(__vmins2(pk(s.x, s.y), LA2), NLA2);      // each sum is at most three extrinsics: fits a lane, clamped packed
            ls[i] = vadd(s8x2_to_s16x2(Ls16[i * NT2]), la);
            lp[i] = s8x2_to_s16x2(Lp16[i * NT2]);
        }
    }
}
// keeps the compiler from moving memory operations across a phase boundary (hoisted loads lengthen live ranges and end in spills)
__device__ __forceinline__ void phase_fence() { asm volatile("" ::: "memory"); }

template <int NT, bool SECOND>
__device__ __forceinline__ void dec_pass(Smem<NT> &sm, const int t, const bool active, uint32_t (&a_in)[8], uint32_t (&b)[8])
{
    constexpr int NT2 = NT / 2;
    constexpr int D = SECOND ? 1 : 0;
    if (active) {
        uint4 *scr = &sm.scr[0][t];                           // private stash: [6][NT2] x 16 B, conflict free
        uint32_t a[8], A[8][8];
        uint32_t hi_ls[8], hi_lp[8];                          // inputs of steps 8..15, live from their prefetch to their backward step
        {
            uint32_t in_ls[8], in_lp[8];
            load_inputs<NT, SECOND>(sm, t, 0, in_ls, in_lp);
            // The inputs of steps 0..7 and the window's start metrics are needed again only after the backward recursion over steps
            // 15..8: they wait in shared memory meanwhile, which keeps the peak register demand (64 stored forward metrics + 8 backward
            // metrics + the inputs of 8 steps) under the 168 registers that two blocks per SM allow -- the first version spilled 30
            // registers per pass to local memory and spent a fifth of the pass waiting for them; now 123 registers, no spills
            // (profiles/r2_chain_ncu_summary.json).
            scr[0 * NT2] = make_uint4(in_ls[0], in_ls[1], in_ls[2], in_ls[3]);
            scr[1 * NT2] = make_uint4(in_ls[4], in_ls[5], in_ls[6], in_ls[7]);
            scr[2 * NT2] = make_uint4(in_lp[0], in_lp[1], in_lp[2], in_lp[3]);
            scr[3 * NT2] = make_uint4(in_lp[4], in_lp[5], in_lp[6], in_lp[7]);
            scr[4 * NT2] = make_uint4(a_in[0], a_in[1], a_in[2], a_in[3]);
            scr[5 * NT2] = make_uint4(a_in[4], a_in[5], a_in[6], a_in[7]);
#pragma unroll
            for (int s = 0; s < 8; s++) a[s] = a_in[s];
            // ---- forward over steps 0..7 (nothing kept)
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if ((i & 3) == 0) fwd_step<true>(a, in_ls[i], in_lp[i]); else fwd_step<false>(a, in_ls[i], in_lp[i]);
            }
        }
        phase_fence();
        load_inputs<NT, SECOND>(sm, t, 8, hi_ls, hi_lp);
        // ---- forward over steps 8..15, the metrics at the start of each step stay in registers
#pragma unroll
        for (int i = 0; i < 8; i++) {
#pragma unroll
            for (int s = 0; s < 8; s++) A[i][s] = a[s];
            if ((i & 3) == 0) fwd_step<true>(a, hi_ls[i], hi_lp[i]); else fwd_step<false>(a, hi_ls[i], hi_lp[i]);
        }
        sm.recA[D][1 + t][0] = make_uint4(a[0], a[1], a[2], a[3]);
        sm.recA[D][1 + t][1] = make_uint4(a[4], a[5], a[6], a[7]);
        const uint16_t *Ls16 = reinterpret_cast<const uint16_t *>(sm.Ls) + t;
        uint32_t *S1w = reinterpret_cast<uint32_t *>(sm.S1) + t;
        const uint32_t *pi32 = reinterpret_cast<const uint32_t *>(sm.piT2) + t;
        char *accb = reinterpret_cast<char *>(sm.acc);
        // backward step i of both windows + extrinsic output
        auto back = [&](const int i, const uint32_t (&al)[8], const uint32_t ls, const uint32_t lp) {
            uint32_t d2;
            if ((i & 3) == 3) d2 = bwd_step<true>(b, al, ls, lp); else d2 = bwd_step<false>(b, al, ls, lp);
            d2 = vsub(d2, ls);
            int eA, eB;
            ext_pair(d2, eA, eB);
            if (SECOND) {
                const uint32_t q2 = pi32[i * NT2];
                atomicAdd(reinterpret_cast<int *>(accb + 2 * (q2 & 0xffffu)), clampi(eA, -LE_LIM, LE_LIM));
                atomicAdd(reinterpret_cast<int *>(accb + 2 * (q2 >> 16)), clampi(eB, -LE_LIM, LE_LIM));
            } else {
                // |(3 d) >> 2| < 2^15 (|d| < 2^15), so the pair packs before the clamp
                S1w[i * NT2] = vadd(s8x2_to_s16x2(Ls16[i * NT2]), vmax(__vmins2(pk(eA, eB), LE2), NLE2));
            }
        };
        // ---- backward over steps 15..8
#pragma unroll
        for (int j = 7; j >= 0; j--) back(8 + j, A[j], hi_ls[j], hi_lp[j]);
        phase_fence();
        // ---- forward metrics of steps 0..7 again, from the stashed start metrics and inputs; then backward over steps 7..0
        uint32_t lo_ls[8], lo_lp[8];
        {
            const uint4 v0 = scr[0 * NT2], v1 = scr[1 * NT2], v2 = scr[2 * NT2], v3 = scr[3 * NT2], v4 = scr[4 * NT2], v5 = scr[5 * NT2];
            lo_ls[0] = v0.x; lo_ls[1] = v0.y; lo_ls[2] = v0.z; lo_ls[3] = v0.w; lo_ls[4] = v1.x; lo_ls[5] = v1.y; lo_ls[6] = v1.z; lo_ls[7] = v1.w;
            lo_lp[0] = v2.x; lo_lp[1] = v2.y; lo_lp[2] = v2.z; lo_lp[3] = v2.w; lo_lp[4] = v3.x; lo_lp[5] = v3.y; lo_lp[6] = v3.z; lo_lp[7] = v3.w;
            a[0] = v4.x; a[1] = v4.y; a[2] = v4.z; a[3] = v4.w; a[4] = v5.x; a[5] = v5.y; a[6] = v5.z; a[7] = v5.w;
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
#pragma unroll
            for (int s = 0; s < 8; s++) A[i][s] = a[s];
            if (i < 7) { if ((i & 3) == 0) fwd_step<true>(a, lo_ls[i], lo_lp[i]); else fwd_step<false>(a, lo_ls[i], lo_lp[i]); }
        }
#pragma unroll
        for (int j = 7; j >= 0; j--) back(j, A[j], lo_ls[j], lo_lp[j]);
        sm.recB[D][t][0] = make_uint4(b[0], b[1], b[2], b[3]);
        sm.recB[D][t][1] = make_uint4(b[4], b[5], b[6], b[7]);
        // ---- boundary metrics of the NEXT pass (the other decoder): nobody writes that decoder's records before the barrier below
        constexpr int O = 1 - D;
        {
            const uint4 p0 = sm.recA[O][t][0], p1 = sm.recA[O][t][1], c0 = sm.recA[O][1 + t][0], c1 = sm.recA[O][1 + t][1];
            // window 2t starts where window 2t-1 ended (hi lane of the previous record), window 2t+1 where window 2t ended (own lo lane)
            a_in[0] = __byte_perm(p0.x, c0.x, 0x5432); a_in[1] = __byte_perm(p0.y, c0.y, 0x5432);
            a_in[2] = __byte_perm(p0.z, c0.z, 0x5432); a_in[3] = __byte_perm(p0.w, c0.w, 0x5432);
            a_in[4] = __byte_perm(p1.x, c1.x, 0x5432); a_in[5] = __byte_perm(p1.y, c1.y, 0x5432);
            a_in[6] = __byte_perm(p1.z, c1.z, 0x5432); a_in[7] = __byte_perm(p1.w, c1.w, 0x5432);
            const uint4 o0 = sm.recB[O][t][0], o1 = sm.recB[O][t][1], n0 = sm.recB[O][t + 1][0], n1 = sm.recB[O][t + 1][1];
            // window 2t ends where window 2t+1 starts (own hi lane), window 2t+1 where window 2t+2 starts (lo lane of the next record)
            b[0] = __byte_perm(o0.x, n0.x, 0x5432); b[1] = __byte_perm(o0.y, n0.y, 0x5432);
            b[2] = __byte_perm(o0.z, n0.z, 0x5432); b[3] = __byte_perm(o0.w, n0.w, 0x5432);
            b[4] = __byte_perm(o1.x, n1.x, 0x5432); b[5] = __byte_perm(o1.y, n1.y, 0x5432);
            b[6] = __byte_perm(o1.z, n1.z, 0x5432); b[7] = __byte_perm(o1.w, n1.w, 0x5432);
        }
    }
    __syncthreads();
}

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) global -> shared with completion on an mbarrier
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

constexpr int MODE_F32 = 0;   // input: rate-unmatched float LLRs llr[n_cb][3(K+4)] (quantised here);  output: decoded bits out[n_cb][K]
constexpr int MODE_Q8 = 1;    // input: quantised, window-transposed LLR image q8[n_cb][3 NP + 16] written by pdsch_llr_q8_kernel (pdsch_llr.cu);
                              // output: transport block bits + CRC24A verdict (block index = local_alloc * n_sf + subframe)

// piT2 [16*NT] : 2 * pos(pi(k)) stored at pos(k)                                                     (global, per K)
template <int NT, int MODE>
__global__ void __launch_bounds__(NT / 2 < 32 ? 32 : NT / 2, NT == 384 ? 2 : (NT == 192 ? 4 : (NT == 128 ? 6 : 12)))
turbo_mlm2_kernel(const void *__restrict__ in, int K, int n_iter, const uint16_t *__restrict__ piT2g, uint8_t *__restrict__ out,
                  const __grid_constant__ FusedParams fp)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem<NT> &sm = *reinterpret_cast<Smem<NT> *>(smem_raw);
    constexpr int NT2 = NT / 2, NP = WIN * NT;
    __shared__ float s_red[32];
    __shared__ int s_tail[12];
    __shared__ __align__(8) uint64_t s_bar;

    const int cb = blockIdx.x, t = threadIdx.x, nt = blockDim.x;
    const int n_thr = K / (2 * WIN);                     // active threads: one per pair of windows (K is a multiple of 32)
    const bool active = t < n_thr;
    const int k0 = 2 * WIN * t;                          // first trellis step of window 2t; window 2t+1 starts at k0 + 16

    if (MODE == MODE_Q8) {
        // ---- input: two bulk copies (LLR image of this block, interleaver table) land while the threads clear the exchange state
        constexpr uint32_t IMG = 3 * NP + 16, TAB = 2 * NP;
        if (t == 0) {
            mbar_init(&s_bar, 1);
            mbar_expect_tx(&s_bar, IMG + TAB);
            bulk_copy_g2s(sm.Ls, static_cast<const int8_t *>(in) + (size_t)cb * IMG, IMG, &s_bar);
            bulk_copy_g2s(sm.piT2, piT2g, TAB, &s_bar);
        }
        for (int p = t; p < NP; p += nt) sm.acc[p] = 0;
        for (int p = t; p < 2 * (NT2 + 1) * 2; p += nt) {
            (&sm.recA[0][0][0])[p] = make_uint4(0, 0, 0, 0);
            (&sm.recB[0][0][0])[p] = make_uint4(0, 0, 0, 0);
        }
        __syncthreads();                                    // the barrier object is initialised before anybody polls it
        mbar_wait(&s_bar, 0);
        if (active) {
            const uint16_t *Ls16 = reinterpret_cast<const uint16_t *>(sm.Ls) + t;
            uint32_t *S1w = reinterpret_cast<uint32_t *>(sm.S1) + t;
#pragma unroll
            for (int i = 0; i < WIN; i++) S1w[i * NT2] = s8x2_to_s16x2(Ls16[i * NT2]);     // Le1 = 0: S1 = Lsys
        }
        if (t < 12) s_tail[t] = sm.tail8[t];
    } else {
        // ---- input, first pass: max |d| over all 3(K+4) inputs (sentinel -> 0).  Thread t owns 96 consecutive floats (384 B).
        const float *d = static_cast<const float *>(in) + (size_t)cb * 3 * (K + 4);
        float m = 0.f, tail_v = 0.f;
        if (active) {
            const float4 *src = reinterpret_cast<const float4 *>(d + 3 * k0);
#pragma unroll 6
            for (int j = 0; j < 6 * WIN / 4; j++) {
                float4 v = src[j];
                if (v.x == LTEB200_RX_NULL_BIT) v.x = 0.f;
                if (v.y == LTEB200_RX_NULL_BIT) v.y = 0.f;
                if (v.z == LTEB200_RX_NULL_BIT) v.z = 0.f;
                if (v.w == LTEB200_RX_NULL_BIT) v.w = 0.f;
                m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
            }
        }
        if (t < 12) {
            tail_v = d[3 * K + t];
            if (tail_v == LTEB200_RX_NULL_BIT) tail_v = 0.f;
            m = fmaxf(m, fabsf(tail_v));
        }
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if ((t & 31) == 0) s_red[t >> 5] = m;
        __syncthreads();
        if (t < 32) {
            float v = t < (nt >> 5) ? s_red[t] : 0.f;
            for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
            if (t == 0) s_red[31] = v;                       // (its own slot: lanes of this warp have just read s_red[0..5])
        }
        __syncthreads();
        const float maxv = s_red[31];
        // ---- second pass: quantise into the window-transposed arrays, thread t writes the (2t, 2t+1) word of every row
        if (active) {
            uint16_t *Ls16 = reinterpret_cast<uint16_t *>(sm.Ls) + t, *Lp116 = reinterpret_cast<uint16_t *>(sm.Lp1) + t, *Lp216 = reinterpret_cast<uint16_t *>(sm.Lp2) + t;
            uint32_t *S1w = reinterpret_cast<uint32_t *>(sm.S1) + t;
            const float *dw = d + 3 * k0;
#pragma unroll
            for (int c = 0; c < WIN / 4; c++) {              // 4 steps of both windows per round: 12 + 12 floats
                float va[12], vb[12];
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    const float4 x = reinterpret_cast<const float4 *>(dw + 12 * c)[j], y = reinterpret_cast<const float4 *>(dw + 3 * WIN + 12 * c)[j];
                    va[4 * j] = x.x; va[4 * j + 1] = x.y; va[4 * j + 2] = x.z; va[4 * j + 3] = x.w;
                    vb[4 * j] = y.x; vb[4 * j + 1] = y.y; vb[4 * j + 2] = y.z; vb[4 * j + 3] = y.w;
                }
#pragma unroll
                for (int j = 0; j < 12; j++) {
                    if (va[j] == LTEB200_RX_NULL_BIT) va[j] = 0.f;
                    if (vb[j] == LTEB200_RX_NULL_BIT) vb[j] = 0.f;
                }
#pragma unroll
                for (int ii = 0; ii < 4; ii++) {
                    const int i = 4 * c + ii;
                    const int sA = qf(va[3 * ii], maxv), sB = qf(vb[3 * ii], maxv);
                    const int pA = qf(va[3 * ii + 1], maxv), pB = qf(vb[3 * ii + 1], maxv);
                    const int rA = qf(va[3 * ii + 2], maxv), rB = qf(vb[3 * ii + 2], maxv);
                    Ls16[i * NT2] = (uint16_t)((sA & 0xff) | ((sB & 0xff) << 8));
                    Lp116[i * NT2] = (uint16_t)((pA & 0xff) | ((pB & 0xff) << 8));
                    Lp216[i * NT2] = (uint16_t)((rA & 0xff) | ((rB & 0xff) << 8));
                    S1w[i * NT2] = pk(sA, sB);
                }
            }
        }
        if (t < 12) s_tail[t] = qf(tail_v, maxv);            // s_tail[3*r + x] = q(d_x[K+r])
        // ---- interleaver table, accumulators, boundary records
        for (int p = t; p < NP / 2; p += nt) reinterpret_cast<uint32_t *>(sm.piT2)[p] = reinterpret_cast<const uint32_t *>(piT2g)[p];
        for (int p = t; p < NP; p += nt) sm.acc[p] = 0;
        for (int p = t; p < 2 * (NT2 + 1) * 2; p += nt) {
            (&sm.recA[0][0][0])[p] = make_uint4(0, 0, 0, 0);
            (&sm.recB[0][0][0])[p] = make_uint4(0, 0, 0, 0);
        }
    }
    __syncthreads();
    if (t < 2) {
        // thread d prepares decoder d: block-start vector (state 0 known) in the hi lanes of recA[d][0], termination vector in the lo lanes
        // of recB[d][n_thr].  beta at time K: three forced termination steps (u = feedback, so the register input is 0) ending in state 0;
        // tail multiplexing of liblte_phy.cc:10575-10586: x = systematic/feedback bits, z = parity bits of the 3 tail steps
        int xt[3], zt[3];
        if (t == 0) { xt[0] = s_tail[0]; xt[1] = s_tail[2]; xt[2] = s_tail[4]; zt[0] = s_tail[1]; zt[1] = s_tail[3]; zt[2] = s_tail[5]; }
        else        { xt[0] = s_tail[6]; xt[1] = s_tail[8]; xt[2] = s_tail[10]; zt[0] = s_tail[7]; zt[1] = s_tail[9]; zt[2] = s_tail[11]; }
        int bt[8];
        bt[0] = 0;
#pragma unroll
        for (int s = 1; s < 8; s++) bt[s] = NEG;
#pragma unroll
        for (int q = 2; q >= 0; q--) {
            int nb[8];
            const int ls = xt[q], lp = zt[q], b0 = bt[0];
#pragma unroll
            for (int S = 0; S < 8; S++) {
                const int u = ((S >> 1) & 1) ^ (S & 1);
                const int g = (u == 0 ? ls : 0) + (tr_par(S, u) == 0 ? lp : 0);
                nb[S] = max(g - b0 + bt[tr_nxt(S, u)], NEG);
            }
#pragma unroll
            for (int S = 0; S < 8; S++) bt[S] = nb[S];
        }
        sm.recB[t][n_thr][0] = make_uint4(pk(bt[0], 0), pk(bt[1], 0), pk(bt[2], 0), pk(bt[3], 0));
        sm.recB[t][n_thr][1] = make_uint4(pk(bt[4], 0), pk(bt[5], 0), pk(bt[6], 0), pk(bt[7], 0));
        const uint32_t neg_hi = pk(0, NEG);
        sm.recA[t][0][0] = make_uint4(pk(0, 0), neg_hi, neg_hi, neg_hi);
        sm.recA[t][0][1] = make_uint4(neg_hi, neg_hi, neg_hi, neg_hi);
    }
    __syncthreads();

    // boundary metrics of the first pass (first decoder): all-zero records, block start and termination
    uint32_t a_in[8], b[8];
    if (active) {
        const uint4 p0 = sm.recA[0][t][0], p1 = sm.recA[0][t][1], c0 = sm.recA[0][1 + t][0], c1 = sm.recA[0][1 + t][1];
        a_in[0] = __byte_perm(p0.x, c0.x, 0x5432); a_in[1] = __byte_perm(p0.y, c0.y, 0x5432);
        a_in[2] = __byte_perm(p0.z, c0.z, 0x5432); a_in[3] = __byte_perm(p0.w, c0.w, 0x5432);
        a_in[4] = __byte_perm(p1.x, c1.x, 0x5432); a_in[5] = __byte_perm(p1.y, c1.y, 0x5432);
        a_in[6] = __byte_perm(p1.z, c1.z, 0x5432); a_in[7] = __byte_perm(p1.w, c1.w, 0x5432);
        const uint4 o0 = sm.recB[0][t][0], o1 = sm.recB[0][t][1], n0 = sm.recB[0][t + 1][0], n1 = sm.recB[0][t + 1][1];
        b[0] = __byte_perm(o0.x, n0.x, 0x5432); b[1] = __byte_perm(o0.y, n0.y, 0x5432);
        b[2] = __byte_perm(o0.z, n0.z, 0x5432); b[3] = __byte_perm(o0.w, n0.w, 0x5432);
        b[4] = __byte_perm(o1.x, n1.x, 0x5432); b[5] = __byte_perm(o1.y, n1.y, 0x5432);
        b[6] = __byte_perm(o1.z, n1.z, 0x5432); b[7] = __byte_perm(o1.w, n1.w, 0x5432);
    } else {
#pragma unroll
        for (int s = 0; s < 8; s++) { a_in[s] = 0; b[s] = 0; }
    }
    __syncthreads();

    for (int it = 0; it < n_iter; it++) {
        dec_pass<NT, false>(sm, t, active, a_in, b);
        dec_pass<NT, true>(sm, t, active, a_in, b);
    }

    // ---- decisions of the thread's 32 steps: bit = (Lsys + Le1 + La1 < 0)
    uint32_t bits = 0;                                      // bit j = decision of step k0 + j
    if (active) {
        const uint32_t *S1w = reinterpret_cast<const uint32_t *>(sm.S1) + t;
        const int2 *acc2 = reinterpret_cast<const int2 *>(sm.acc) + t;
#pragma unroll
        for (int i = 0; i < WIN; i++) {
            const uint32_t s = S1w[i * NT2];
            const int2 la = acc2[i * NT2];
            if (lo16(s) + clampi(la.x, -LA_LIM, LA_LIM) < 0) bits |= 1u << i;
            if (hi16(s) + clampi(la.y, -LA_LIM, LA_LIM) < 0) bits |= 1u << (WIN + i);
        }
    }
    if (MODE == MODE_F32) {
        if (active) {
            // one byte per bit, 32 consecutive bytes
            uint32_t w[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t n = (bits >> (4 * j)) & 0xfu;
                w[j] = (n & 1u) | ((n & 2u) << 7) | ((n & 4u) << 14) | ((n & 8u) << 21);
            }
            uint4 *o = reinterpret_cast<uint4 *>(out + (size_t)cb * K + k0);
            o[0] = make_uint4(w[0], w[1], w[2], w[3]);
            o[1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
    } else {
        // dlsch_channel_decode's tail (liblte_phy.cc:12841-12869): drop the F filler bits, CRC24A over [a | p], emit a.
        // The remainder of the tbs+24 bit message is the XOR of x^(distance from the end) mod g over its set bits.
        const int a_loc = cb / fp.n_sf, sf = cb - a_loc * fp.n_sf;
        const FusedAlloc &A = fp.al[a_loc];
        const int n = A.tbs + 24;
        uint8_t *o = fp.tb_out + ((size_t)sf * fp.n_alloc_total + A.a_idx) * fp.tb_stride;
        uint32_t rem = 0;
        if (active) {
            const int idx0 = k0 - A.F;                                     // index of this thread's first bit inside [a | p]
            const bool fast = idx0 >= 0 && idx0 + 2 * WIN <= A.tbs && (((size_t)(o + idx0)) & 7) == 0;
            if (fp.tb_packed) {
                // eight bits per byte, first bit in the MSB (the launcher admits this format only without filler bits: idx0 = k0, 32 | k0)
                const uint32_t w = __byte_perm(__brev(bits), 0, 0x0123);   // message bit k0 + j -> byte j / 8, bit 7 - j % 8
                uint8_t *ob = o + (idx0 >> 3);
                if (idx0 + 2 * WIN <= A.tbs && (((size_t)ob) & 3) == 0) *reinterpret_cast<uint32_t *>(ob) = w;
                else
                    for (int j = 0; j < 4; j++)
                        if (idx0 + 8 * j < A.tbs) ob[j] = (uint8_t)(w >> (8 * j));    // transport block sizes are multiples of 8
            } else if (fast) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t b0 = (bits >> (8 * j)) & 0xfu, b1 = (bits >> (8 * j + 4)) & 0xfu;
                    const uint32_t w0 = (b0 & 1u) | ((b0 & 2u) << 7) | ((b0 & 4u) << 14) | ((b0 & 8u) << 21);
                    const uint32_t w1 = (b1 & 1u) | ((b1 & 2u) << 7) | ((b1 & 4u) << 14) | ((b1 & 8u) << 21);
                    reinterpret_cast<uint2 *>(o + idx0)[j] = make_uint2(w0, w1);
                }
            } else {
#pragma unroll 4
                for (int j = 0; j < 2 * WIN; j++) {
                    const int idx = idx0 + j;
                    if (idx >= 0 && idx < A.tbs) o[idx] = (uint8_t)((bits >> j) & 1u);
                }
            }
            // 32 independent (predicated) table reads, all issued before the first is consumed
            const uint32_t *ct = fp.crc_tab + (n - 1 - idx0);
#pragma unroll
            for (int j = 0; j < 2 * WIN; j++)
                if (idx0 + j >= 0 && ((bits >> j) & 1u)) rem ^= ct[-j];
        }
        for (int off = 16; off > 0; off >>= 1) rem ^= __shfl_xor_sync(0xffffffffu, rem, off);
        if ((t & 31) == 0) s_tail[t >> 5] = (int)rem;                      // <= 6 warps
        __syncthreads();
        if (t == 0) {
            uint32_t r = 0;
            for (int wv = 0; wv < (nt >> 5); wv++) r ^= (uint32_t)s_tail[wv];
            fp.crc_ok[(size_t)sf * fp.n_alloc_total + A.a_idx] = (r & 0xffffffu) == 0 ? 1 : 0;
        }
    }
}

}  // namespace

bool mlm2_handles(int K) { return K >= 64 && (K & 31) == 0 && getenv("LTEB200_MLM_V1") == nullptr; }
size_t mlm2_q8_stride(int K) { return (size_t)3 * WIN * mlm_size_class(K) + 16; }

// q8 == nullptr: float LLRs d_llr[n_cb][3(K+4)] in, decoded bits d_bits[n_cb][K] out.
// q8 != nullptr: quantised images in (mlm2_q8_stride(K) bytes per block), transport blocks + CRC verdicts out as `fused` describes.
int turbo_mlm2_launch(lteb200_ctx *ctx, const float *d_llr, const int8_t *q8, int K, int n_cb, int n_iter, uint8_t *d_bits, const FusedParams *fused,
                      cudaStream_t stream)
{
    if (!mlm2_handles(K) || K > 6144) { lteb200_set_error(ctx, "turbo (pair kernel): K=%d is not a multiple of 32", K); return LTEB200_ERR_ARGS; }
    if ((q8 != nullptr) != (fused != nullptr)) { lteb200_set_error(ctx, "turbo (pair kernel): inconsistent input/output mode"); return LTEB200_ERR_ARGS; }
    if (n_iter < 1) n_iter = 1;
    QppTables qt;
    int rc = lteb200_get_qpp(ctx, K, &qt);  // validates K
    if (rc) return rc;
    const int NT = mlm_size_class(K), NP = WIN * NT;
    uint16_t *piT2 = nullptr;
    {
        auto it = ctx->mlm2.find(K | (ctx->qpp_spec << 16));
        if (it == ctx->mlm2.end()) {
            std::vector<uint16_t> pi(K);
            LTE_CUDA_OK(ctx, cudaMemcpy(pi.data(), qt.pi, K * sizeof(uint16_t), cudaMemcpyDeviceToHost));
            auto pos = [&](int k) { return (k % WIN) * NT + (k / WIN); };
            std::vector<uint16_t> tab(NP, 0);
            for (int k = 0; k < K; k++) tab[pos(k)] = (uint16_t)(2 * pos(pi[k]));
            LTE_CUDA_OK(ctx, cudaMalloc(&piT2, (size_t)NP * 2));
            LTE_CUDA_OK(ctx, cudaMemcpy(piT2, tab.data(), (size_t)NP * 2, cudaMemcpyHostToDevice));
            ctx->mlm2[K | (ctx->qpp_spec << 16)] = piT2;
        } else piT2 = it->second;
    }
    const int n_thr = K / (2 * WIN);
    const int threads = ((n_thr + 31) / 32) * 32;
    FusedParams none;
    memset(&none, 0, sizeof(none));
    const FusedParams &fp = fused ? *fused : none;
    const void *in = q8 ? (const void *)q8 : (const void *)d_llr;
#define MLM2_LAUNCH(NT_, MODE_)                                                                                                        \
    do {                                                                                                                               \
        const size_t smem = sizeof(Smem<NT_>);                                                                                         \
        LTE_CUDA_OK(ctx, cudaFuncSetAttribute(turbo_mlm2_kernel<NT_, MODE_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        LTE_CUDA_OK(ctx, cudaFuncSetAttribute(turbo_mlm2_kernel<NT_, MODE_>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));    \
        turbo_mlm2_kernel<NT_, MODE_><<<n_cb, threads, smem, stream>>>(in, K, n_iter, piT2, d_bits, fp);                               \
    } while (0)
    if (q8) { if (NT == 64) MLM2_LAUNCH(64, MODE_Q8); else if (NT == 128) MLM2_LAUNCH(128, MODE_Q8); else if (NT == 192) MLM2_LAUNCH(192, MODE_Q8); else MLM2_LAUNCH(384, MODE_Q8); }
    else    { if (NT == 64) MLM2_LAUNCH(64, MODE_F32); else if (NT == 128) MLM2_LAUNCH(128, MODE_F32); else if (NT == 192) MLM2_LAUNCH(192, MODE_F32); else MLM2_LAUNCH(384, MODE_F32); }
#undef MLM2_LAUNCH
    ctx->launches += 1;
    LTE_CUDA_OK(ctx, cudaGetLastError());
    return LTEB200_OK;
}

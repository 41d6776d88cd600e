// Reference-exact turbo decoder ("mode A"): bit-for-bit twin of turbo_decode(),
// /root/reference/liblte/src/liblte_phy.cc:10620-10845, for batches of code blocks resident in HBM.
//
// What the reference computes: quantise the three LLR streams to int8, run three
// soft-output Viterbi passes (viterbi_decode_siso, liblte_phy.cc:10341-10529), re-encode softly
// (conv_encode_soft g=03, :10070-10151), merge, (de)interleave with the uint32-overflowing QPP
// (:10960-10985, :11047-11072) and take a 4-way int8 vote.  It is not iterative and not max-log-MAP.
//
// B200 mapping.  The only serial part is the 8-state add-compare-select recursion (K steps, each depending
// on the previous one, and not a semiring recursion: the survivor is chosen on the UNWEIGHTED sum while the
// WEIGHTED sum is stored, so it cannot be turned into a parallel scan without changing results).  Everything
// else is elementwise.  So:
//   * one THREAD per trellis keeps the 8 path metrics in registers (int32: all metrics are integers below
//     2^24, so integer arithmetic reproduces the reference's float arithmetic exactly) -- no shuffles, ILP 8,
//     ~60 instructions per trellis step; a warp advances 32 trellises per instruction stream, which is
//     fewer issue slots per trellis step than an 8-lanes-per-trellis shuffle layout (2 vs 2.5) and has no
//     shuffle latency on the critical path;
//   * everything of a step that does not depend on the path metrics (the signed weights W*A, W*B and the four compare
//     thresholds +-2A, +-2B) is computed by the elementwise kernels and staged as an 8-byte step descriptor in a
//     step-major, trellis-minor layout (16 B = 2 steps per trellis, a warp's loads are 512 B contiguous): the serial loop is
//     38 instructions per step instead of 60;
//   * survivor information is 4 compare bits per step (the traceback of the reference consults stored metrics of the two
//     predecessors, not survivor decisions), 3 KB per K=6144 trellis instead of the reference's 394 KB of metric history;
//   * the traceback is NOT serial: one step maps the state's low two bits m to ((m & 1) << 1 | compare bit of butterfly m), a
//     4 -> 4 map, and maps compose associatively.  The elementwise kernel that consumes the decisions builds the map of every
//     8-step word, suffix-scans the maps over the block (256 threads, 3 words each) and re-walks its words from the now known
//     entry states -- the trellis thread only runs the forward recursion;
//   * the soft output magnitude is state independent (127*W_t/W_max, W_t = |a_t|+|b_t|), so it is produced
//     elementwise outside the serial loop;
//   * the three SISO passes run as two launches (pass 1: n_cb trellises; passes 2 and 3 are independent of
//     each other and run together as 2*n_cb trellises).
// No tensor cores: there is no dense contraction anywhere on this path.
#include "internal.h"

namespace {

constexpr int PREP_THREADS = 256;

__device__ __forceinline__ int q8(float v, float maxv)
{
    // (int8)(d*127/max_value), liblte_phy.cc:10662: IEEE multiply, IEEE divide, truncate; NaN (0/0) -> 0 as on x86-64
    return __float2int_rz(__fdiv_rn(__fmul_rn(v, 127.0f), maxv));
}
__device__ __forceinline__ int soft_mag(int W, int Wmax)
{
    // (int8)(127 * (w/max_weight)), liblte_phy.cc:10512
    return __float2int_rz(__fmul_rn(127.0f, __fdiv_rn((float)W, (float)Wmax)));
}
__device__ __forceinline__ int iabs(int x) { return x < 0 ? -x : x; }
// PRMT with sign replication (selector bit 3); __byte_perm() masks that bit away
template <uint32_t SEL> __device__ __forceinline__ int prmt_sx(uint32_t a)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %1, %2;" : "=r"(d) : "r"(a), "n"(SEL));
    return (int)d;
}
// Step descriptor of the add-compare-select recursion (see Acs below): x = (W*A as int16, W*B as int16), y = (2A, -2A, 2B, -2B) as int8.
// a = parity input, b = systematic input of the step (liblte_phy.cc:10379-10408: sa, sb are the signs, W the unsigned weight).
__device__ __forceinline__ uint2 make_desc(int a, int b)
{
    const int sa = a >= 0 ? 1 : -1, sb = b >= 0 ? 1 : -1;
    const int W = iabs(a) + iabs(b);
    const int A = sa + sb, B = sb - sa;
    uint2 d;
    d.x = (uint32_t)((W * A) & 0xffff) | ((uint32_t)(W * B) << 16);
    d.y = (uint32_t)((2 * A) & 0xff) | ((uint32_t)((-2 * A) & 0xff) << 8) | ((uint32_t)((2 * B) & 0xff) << 16) | ((uint32_t)(-2 * B) << 24);
    return d;
}
// soft XOR of conv_encode_soft (liblte_phy.cc:10123-10147): magnitude (|x|+|y|)>>1, negative iff signs differ
__device__ __forceinline__ int sxor(int x, int y)
{
    int mag = (iabs(x) + iabs(y)) >> 1;
    return ((x < 0) != (y < 0)) ? -mag : mag;
}
// steps 3 / 10 / 11 of turbo_decode (liblte_phy.cc:10688-10707, 10777-10797, 10800-10820)
template <int VARIANT>
__device__ __forceinline__ int merge(int act, int fb, int alt)
{
    if (act >= 0 && fb >= 0) return (act + fb) >> 1;
    if (act < 0 && fb < 0) return (-act - fb) >> 1;
    if (act >= 0 && fb < 0) return -(((VARIANT == 1 ? alt : act) - fb) >> 1);
    if (VARIANT == 1) return -((-alt + fb) >> 1);
    if (VARIANT == 2) return -((-act - fb) >> 1);
    return -((-act + fb) >> 1);
}

// byte s (0..7) of an 8-byte chunk, sign-extended
__device__ __forceinline__ int sbyte(const uint2 v, const int s) { return (int)(int8_t)(((s < 4 ? v.x : v.y) >> (8 * (s & 3))) & 0xffu); }
__device__ __forceinline__ void put_byte(uint2 &v, const int s, const int x)
{
    if (s < 4) v.x |= (uint32_t)(x & 0xff) << (8 * s); else v.y |= (uint32_t)(x & 0xff) << (8 * (s - 4));
}
// elements i-2 and i-3 of a byte stream, i = 8c + s, from the chunk c (cur) and the chunk before it (prev)
__device__ __forceinline__ int back2(const uint2 prev, const uint2 cur, const int s) { return s >= 2 ? sbyte(cur, s - 2) : sbyte(prev, s + 6); }
__device__ __forceinline__ int back3(const uint2 prev, const uint2 cur, const int s) { return s >= 3 ? sbyte(cur, s - 3) : sbyte(prev, s + 5); }
constexpr uint32_t FILL127 = 0x7f7f7f7fu;     // conv_encode_soft's register preset: act[i < 0] = +127

__device__ __forceinline__ int block_max(int v, int *sh)
{
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        int w = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0;
        for (int o = 16; o > 0; o >>= 1) w = max(w, __shfl_xor_sync(0xffffffffu, w, o));
        if (threadIdx.x == 0) sh[0] = w;
    }
    __syncthreads();
    int r = sh[0];
    __syncthreads();
    return r;
}

// ---------------------------------------------------------------------------------------------------------
// Step 0/1 front end: sentinel -> 0, max |d| over i < K, quantise, build the pass-1 tiles (a = parity 1, b = systematic).
// One CTA per code block.  llr: [n_cb][3*(K+4)] interleaved d[3i+x].
__global__ void __launch_bounds__(PREP_THREADS)
turbo_ref_prep(const float *__restrict__ llr, int K, int n_cb, int8_t *__restrict__ q /*[3][n_cb][K]*/,
               uint4 *__restrict__ tile1 /*[K/2][n_cb]*/, int *__restrict__ wmax1 /*[n_cb]*/)
{
    __shared__ int sh[32];
    const int cb = blockIdx.x;
    const float *d = llr + (size_t)cb * 3 * (K + 4);
    float m = 0.f;
    const bool vec = (reinterpret_cast<size_t>(llr) & 15) == 0;        // rows are 12 K + 48 bytes apart: 16-byte aligned if the batch is
    if (vec) {
        const float4 *d4 = reinterpret_cast<const float4 *>(d);
        for (int idx = threadIdx.x; idx < 3 * K / 4; idx += PREP_THREADS) {
            float4 v = d4[idx];
            if (v.x == LTEB200_RX_NULL_BIT) v.x = 0.f;
            if (v.y == LTEB200_RX_NULL_BIT) v.y = 0.f;
            if (v.z == LTEB200_RX_NULL_BIT) v.z = 0.f;
            if (v.w == LTEB200_RX_NULL_BIT) v.w = 0.f;
            m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        }
    } else {
        for (int idx = threadIdx.x; idx < 3 * K; idx += PREP_THREADS) {
            float v = d[idx];
            if (v == LTEB200_RX_NULL_BIT) v = 0.f;
            m = fmaxf(m, fabsf(v));
        }
    }
    const float maxv = __int_as_float(block_max(__float_as_int(m), sh));  // non-negative floats order like ints
    int wm = 0;
    const int K8 = K >> 3;
    for (int c = threadIdx.x; c < K8; c += PREP_THREADS) {
        uint2 ds[8];
        uint32_t p0[2] = {0, 0}, p1[2] = {0, 0}, p2[2] = {0, 0};
        float f[24];                                            // the chunk's 8 x 3 inputs: 96 contiguous bytes
        if (vec) {
#pragma unroll
            for (int j = 0; j < 6; j++) {
                const float4 v = reinterpret_cast<const float4 *>(d + 24 * c)[j];
                f[4 * j] = v.x; f[4 * j + 1] = v.y; f[4 * j + 2] = v.z; f[4 * j + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 24; j++) f[j] = d[24 * c + j];
        }
#pragma unroll
        for (int s = 0; s < 8; s++) {
            float v0 = f[3 * s + 0], v1 = f[3 * s + 1], v2 = f[3 * s + 2];
            if (v0 == LTEB200_RX_NULL_BIT) v0 = 0.f;
            if (v1 == LTEB200_RX_NULL_BIT) v1 = 0.f;
            if (v2 == LTEB200_RX_NULL_BIT) v2 = 0.f;
            const int a0 = q8(v0, maxv), a1 = q8(v1, maxv), a2 = q8(v2, maxv);
            wm = max(wm, iabs(a1) + iabs(a0));
            p0[s >> 2] |= (uint32_t)(a0 & 0xff) << (8 * (s & 3));
            p1[s >> 2] |= (uint32_t)(a1 & 0xff) << (8 * (s & 3));
            p2[s >> 2] |= (uint32_t)(a2 & 0xff) << (8 * (s & 3));
            ds[s] = make_desc(a1, a0);
        }
        const size_t plane = (size_t)n_cb * K;
        *(uint2 *)(q + 0 * plane + (size_t)cb * K + 8 * c) = make_uint2(p0[0], p0[1]);
        *(uint2 *)(q + 1 * plane + (size_t)cb * K + 8 * c) = make_uint2(p1[0], p1[1]);
        *(uint2 *)(q + 2 * plane + (size_t)cb * K + 8 * c) = make_uint2(p2[0], p2[1]);
#pragma unroll
        for (int h = 0; h < 4; h++) tile1[(size_t)(4 * c + h) * n_cb + cb] = make_uint4(ds[2 * h].x, ds[2 * h].y, ds[2 * h + 1].x, ds[2 * h + 1].y);
    }
    wm = block_max(wm, sh);
    if (threadIdx.x == 0) wmax1[cb] = wm;
}

// The same front end fed by the quantised image of the symbol-level kernel (pdsch_llr.cu, written with max_no_tail = 1: quantised against
// the maximum over the 3 K code-block values, exactly the reference's (int8)(d*127/max_value)): no float LLRs, no separate rate-unmatch pass.
// img: [n_cb][3 NP + 16] int8, planes Ls | Lp1 | Lp2 of NP = 16 NT bytes, step k at (k mod 16) NT + k / 16.  Dynamic smem = 3 NP + 16 bytes.
__global__ void __launch_bounds__(PREP_THREADS)
turbo_ref_prep_q8(const int8_t *__restrict__ img, int K, int NT, int n_cb, int8_t *__restrict__ q /*[3][n_cb][K]*/,
                  uint4 *__restrict__ tile1 /*[K/2][n_cb]*/, int *__restrict__ wmax1 /*[n_cb]*/)
{
    extern __shared__ __align__(16) int8_t s_img[];
    __shared__ int sh[32];
    const int cb = blockIdx.x, NP = 16 * NT, bytes = 3 * NP + 16;
    const uint4 *src = reinterpret_cast<const uint4 *>(img + (size_t)cb * bytes);
    for (int p = threadIdx.x; p < bytes / 16; p += PREP_THREADS) reinterpret_cast<uint4 *>(s_img)[p] = src[p];
    __syncthreads();
    int wm = 0;
    const int K8 = K >> 3;
    for (int c = threadIdx.x; c < K8; c += PREP_THREADS) {
        uint2 ds[8];
        uint32_t p0[2] = {0, 0}, p1[2] = {0, 0}, p2[2] = {0, 0};
#pragma unroll
        for (int s = 0; s < 8; s++) {
            const int i = 8 * c + s, pos = (i & 15) * NT + (i >> 4);
            const int a0 = s_img[pos], a1 = s_img[NP + pos], a2 = s_img[2 * NP + pos];
            wm = max(wm, iabs(a1) + iabs(a0));
            p0[s >> 2] |= (uint32_t)(a0 & 0xff) << (8 * (s & 3));
            p1[s >> 2] |= (uint32_t)(a1 & 0xff) << (8 * (s & 3));
            p2[s >> 2] |= (uint32_t)(a2 & 0xff) << (8 * (s & 3));
            ds[s] = make_desc(a1, a0);
        }
        const size_t plane = (size_t)n_cb * K;
        *(uint2 *)(q + 0 * plane + (size_t)cb * K + 8 * c) = make_uint2(p0[0], p0[1]);
        *(uint2 *)(q + 1 * plane + (size_t)cb * K + 8 * c) = make_uint2(p1[0], p1[1]);
        *(uint2 *)(q + 2 * plane + (size_t)cb * K + 8 * c) = make_uint2(p2[0], p2[1]);
#pragma unroll
        for (int h = 0; h < 4; h++) tile1[(size_t)(4 * c + h) * n_cb + cb] = make_uint4(ds[2 * h].x, ds[2 * h].y, ds[2 * h + 1].x, ds[2 * h + 1].y);
    }
    wm = block_max(wm, sh);
    if (threadIdx.x == 0) wmax1[cb] = wm;
}

// ---------------------------------------------------------------------------------------------------------
// The serial core: forward add-compare-select over K steps.  One thread per trellis.
//
// States j = 0..7; predecessors of j are 2j mod 8 and 2j+1 mod 8, i.e. four butterflies m = 0..3 with
// predecessors (2m, 2m+1) and successors (m, m+4).  With sa = (a >= 0 ? +1 : -1), sb likewise, A = sa+sb,
// B = sb-sa, the branch value of butterfly m is M = {-A, +B, -B, +A}[m] (from the generator/register
// definition at liblte_phy.cc:10379-10408):   successor m   : +M from 2m, -M from 2m+1
//                                             successor m+4 : -M from 2m, +M from 2m+1
// Second predecessor wins iff its unweighted sum is strictly smaller; the weighted sum is stored (10452-10463).
// With E = P[2m+1] - P[2m]: "second wins" for successor m is E < -2M, for successor m+4 it is E < 2M; the compare bit kept for the
// traceback (metric[2m] > metric[2m+1]) is the sign of E, shifted into the word of the current 8 steps (first bit ends up in bit 31).
struct Acs {
    int P[8];
    uint32_t bits;
    __device__ __forceinline__ void step(uint32_t x, uint32_t y)
    {
        const int WA = prmt_sx<0x9910u>(x), WB = (int)x >> 16;
        const int A2 = prmt_sx<0x8880u>(y), nA2 = prmt_sx<0x9991u>(y), B2 = prmt_sx<0xAAA2u>(y), nB2 = (int)y >> 24;
        const int E0 = P[1] - P[0], E1 = P[3] - P[2], E2 = P[5] - P[4], E3 = P[7] - P[6];
        bits = __funnelshift_l((uint32_t)E0, bits, 1);
        bits = __funnelshift_l((uint32_t)E1, bits, 1);
        bits = __funnelshift_l((uint32_t)E2, bits, 1);
        bits = __funnelshift_l((uint32_t)E3, bits, 1);
        const int N0 = (E0 < nA2) ? P[1] + WA : P[0] - WA;   // butterfly 0, M = -A
        const int N4 = (E0 < A2) ? P[1] - WA : P[0] + WA;
        const int N1 = (E1 < B2) ? P[3] - WB : P[2] + WB;    // butterfly 1, M = +B
        const int N5 = (E1 < nB2) ? P[3] + WB : P[2] - WB;
        const int N2 = (E2 < nB2) ? P[5] + WB : P[4] - WB;   // butterfly 2, M = -B
        const int N6 = (E2 < B2) ? P[5] - WB : P[4] + WB;
        const int N3 = (E3 < A2) ? P[7] - WA : P[6] + WA;    // butterfly 3, M = +A
        const int N7 = (E3 < nA2) ? P[7] + WA : P[6] - WA;
        P[0] = N0; P[1] = N1; P[2] = N2; P[3] = N3; P[4] = N4; P[5] = N5; P[6] = N6; P[7] = N7;
    }
};

constexpr int PF = 8;     // words of descriptors in flight per trellis

__global__ void __launch_bounds__(32)
turbo_ref_siso(const uint4 *__restrict__ tile /*[K/2][n_tr]*/, int K8, int n_tr, uint32_t *__restrict__ dec /*[K8][n_tr]*/,
               uint8_t *__restrict__ st_end /*[n_tr]*/)
{
    const int t = blockIdx.x * 32 + threadIdx.x;
    if (t >= n_tr) return;
    Acs acs;
#pragma unroll
    for (int j = 0; j < 8; j++) acs.P[j] = 0;  // all-zero start metrics: no known start state (10411-10418)
    acs.bits = 0;
    // The recursion of one 8-step word is about 360 issue slots of a single warp, a fraction of one trip to HBM, and a warp has only
    // six scoreboards to track register loads in flight: so the descriptors travel through a shared-memory ring filled by cp.async
    // (LDGSTS), PF words ahead.  Every thread copies and reads only its own 16-byte columns: no barrier, just wait_group.
    __shared__ uint4 s_ring[PF][4][32];
    const int lane = threadIdx.x;
    const uint4 *tp = tile + t;
    auto fetch = [&](int slot, int c) {
#pragma unroll
        for (int h = 0; h < 4; h++) {
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&s_ring[slot][h][lane]);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(tp + (size_t)(4 * c + h) * n_tr) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int u = 0; u < PF; u++) fetch(u, min(u, K8 - 1));
    // (the loop is NOT unrolled over the ring: eight steps are 360 instructions, which a lone warp must be able to replay from the
    //  instruction cache closest to it -- unrolled over the ring the kernel spent half its cycles waiting for instruction fetch)
#pragma unroll 1
    for (int c = 0; c < K8; c++) {
        const int u = c & (PF - 1);
        asm volatile("cp.async.wait_group %0;" ::"n"(PF - 1) : "memory");
#pragma unroll
        for (int h = 0; h < 4; h++) {
            const uint4 g = s_ring[u][h][lane];
            acs.step(g.x, g.y);
            acs.step(g.z, g.w);
        }
        dec[(size_t)c * n_tr + t] = acs.bits;
        fetch(u, min(c + PF, K8 - 1));                          // (past the end: the last word again, never used)
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    // end state: first strict minimum over j = 0..7 (10467-10481)
    int st = 0, best = acs.P[0];
#pragma unroll
    for (int j = 1; j < 8; j++)
        if (acs.P[j] < best) { best = acs.P[j]; st = j; }
    st_end[t] = (uint8_t)st;
}

// ---------------------------------------------------------------------------------------------------------
// Traceback of one trellis by the PREP_THREADS threads of a block (liblte_phy.cc:10484-10528): from state s at time i+1 go to
// 2(s&3)+1 if metric[2(s&3)] > metric[2(s&3)+1] at time i, else to 2(s&3); the output bit of the transition INTO s is its MSB.
// Compare bit of (step 8c+s, butterfly m) = bit 31 - (4s + m) of word c.
//   walk_word: eight steps backward through one word from the full state `st` at the word's end; returns the state at its start and the
//              eight output bits (bit s = step 8c + s)
//   a "map" holds, for the four values of (state & 3) at a word's end, the full state at its start, one nibble each.
__device__ __forceinline__ int walk_word(uint32_t dw, int st, uint32_t &sb)
{
    sb = 0;
#pragma unroll
    for (int s = 7; s >= 0; s--) {
        sb |= (uint32_t)(st >> 2) << s;
        const int m = st & 3;
        st = 2 * m + (int)((dw >> (31 - 4 * s - m)) & 1u);
    }
    return st;
}
__device__ __forceinline__ uint32_t word_map(uint32_t dw)
{
    uint32_t sb, f = 0;
#pragma unroll
    for (int m = 0; m < 4; m++) f |= (uint32_t)walk_word(dw, m, sb) << (4 * m);
    return f;
}
// first F (later in time: the traceback passes it first), then G.  MAP_ID (bit 31 set) stands for "no words": composing with it must keep
// the other map's full 3-bit states, which a literal identity on (state & 3) would truncate.
__device__ __forceinline__ uint32_t compose(uint32_t G, uint32_t F)
{
    if (G & 0x80000000u) return F;
    uint32_t h = 0;
#pragma unroll
    for (int m = 0; m < 4; m++) h |= ((G >> (4 * ((F >> (4 * m)) & 3u))) & 7u) << (4 * m);
    return h;
}
constexpr uint32_t MAP_ID = 0x80003210u;
constexpr int TB_WORDS = 3;      // words per thread: K8 <= 768 = 3 * PREP_THREADS

// s_sgn[c] = output bits of word c (bit s = step 8c + s); s_warp: PREP_THREADS / 32 words of scratch.  Ends with a barrier.
__device__ void block_traceback(const uint32_t *__restrict__ dec, int n_tr, int tr, int K8, int st_end, uint8_t *s_sgn, uint32_t *s_warp)
{
    const int t = threadIdx.x, lane = t & 31, wp = t >> 5;
    const int c0 = TB_WORDS * t;
    uint32_t dw[TB_WORDS];
    uint32_t L = MAP_ID;
#pragma unroll
    for (int j = TB_WORDS - 1; j >= 0; j--) {
        dw[j] = (c0 + j < K8) ? dec[(size_t)(c0 + j) * n_tr + tr] : 0u;
        if (c0 + j < K8) L = compose(word_map(dw[j]), L);
    }
    // inclusive suffix scan inside the warp: L = maps of threads t .. warp end
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t other = __shfl_down_sync(0xffffffffu, L, o);
        if (lane + o < 32) L = compose(L, other);
    }
    if (lane == 0) s_warp[wp] = L;
    uint32_t X = __shfl_down_sync(0xffffffffu, L, 1);    // threads after this one inside the warp
    if (lane == 31) X = MAP_ID;
    __syncthreads();
    uint32_t later = MAP_ID;                             // warps after this one
    for (int w = PREP_THREADS / 32 - 1; w > wp; w--) later = compose(s_warp[w], later);
    X = compose(X, later);
    const int c_last = min(c0 + TB_WORDS, K8) - 1;
    if (c_last >= c0) {
        int st = (c_last == K8 - 1) ? st_end : (int)((X >> (4 * (st_end & 3))) & 7u);
#pragma unroll
        for (int j = TB_WORDS - 1; j >= 0; j--) {
            if (c0 + j <= c_last) {
                uint32_t sb;
                st = walk_word(dw[j], st, sb);
                s_sgn[c0 + j] = (uint8_t)sb;
            }
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------
// Steps 2-7 front end: act1 -> fb1 -> calc1, interleave q0 and calc1, build the tiles of passes 2 and 3.
// One CTA per code block; dynamic smem = 3*K bytes.
__global__ void __launch_bounds__(PREP_THREADS)
turbo_ref_mid(const int8_t *__restrict__ q, const uint32_t *__restrict__ dec1, const uint8_t *__restrict__ st_end1, const int *__restrict__ wmax1,
              const uint16_t *__restrict__ pi, int K, int n_cb, int8_t *__restrict__ act1_g, int8_t *__restrict__ calc1_g,
              int8_t *__restrict__ sysI_g, int8_t *__restrict__ c1I_g, uint4 *__restrict__ tile2 /*[K/2][2*n_cb]*/,
              int *__restrict__ wmax2 /*[2*n_cb]*/)
{
    extern __shared__ int8_t smem[];
    __shared__ int sh[32];
    __shared__ uint8_t s_sgn[TB_WORDS * PREP_THREADS];
    int8_t *s_q0 = smem, *s_act = smem + K, *s_calc = smem + 2 * K;
    const int cb = blockIdx.x;
    const size_t plane = (size_t)n_cb * K, row = (size_t)cb * K;
    const int wmx = wmax1[cb];
    block_traceback(dec1, n_cb, cb, K >> 3, st_end1[cb], s_sgn, reinterpret_cast<uint32_t *>(sh));
    // (all element-wise phases work on chunks of 8 consecutive elements = one 8-byte access per array)
    const int K8 = K >> 3, n_tr = 2 * n_cb;
    for (int c = threadIdx.x; c < K8; c += PREP_THREADS) {
        const uint2 vb = *(const uint2 *)(q + 0 * plane + row + 8 * c), va = *(const uint2 *)(q + 1 * plane + row + 8 * c);
        const uint32_t sg = s_sgn[c];
        uint2 act = make_uint2(0, 0);
#pragma unroll
        for (int s = 0; s < 8; s++) {
            const int mag = soft_mag(iabs(sbyte(va, s)) + iabs(sbyte(vb, s)), wmx);
            put_byte(act, s, ((sg >> s) & 1u) ? -mag : mag);
        }
        *(uint2 *)(s_q0 + 8 * c) = vb;
        *(uint2 *)(s_act + 8 * c) = act;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < K8; c += PREP_THREADS) {
        // fb[i] = sxor(act[i-2], act[i-3]) with act[<0] = +127 (conv_encode_soft, register preset to 127; fb[0] = 127 = sxor(127, 127))
        const uint2 cur = *(const uint2 *)(s_act + 8 * c);
        const uint2 prev = c > 0 ? *(const uint2 *)(s_act + 8 * c - 8) : make_uint2(FILL127, FILL127);
        uint2 calc = make_uint2(0, 0);
#pragma unroll
        for (int s = 0; s < 8; s++) put_byte(calc, s, merge<0>(sbyte(cur, s), sxor(back2(prev, cur, s), back3(prev, cur, s)), 0));
        *(uint2 *)(s_calc + 8 * c) = calc;
        *(uint2 *)(act1_g + row + 8 * c) = cur;
        *(uint2 *)(calc1_g + row + 8 * c) = calc;
    }
    __syncthreads();
    int wmA = 0, wmB = 0;
    for (int c = threadIdx.x; c < K8; c += PREP_THREADS) {
        uint2 da[8], db[8];
        uint32_t ps[2] = {0, 0}, pc[2] = {0, 0};
        const uint2 q2v = *(const uint2 *)(q + 2 * plane + row + 8 * c);
#pragma unroll
        for (int s = 0; s < 8; s++) {
            const int i = 8 * c + s;
            const int j = pi[i];
            const int a = (int)(int8_t)(((s < 4 ? q2v.x : q2v.y) >> (8 * (s & 3))) & 0xff);
            const int bs = s_q0[j], bc = s_calc[j];
            wmA = max(wmA, iabs(a) + iabs(bs));
            wmB = max(wmB, iabs(a) + iabs(bc));
            ps[s >> 2] |= (uint32_t)(bs & 0xff) << (8 * (s & 3));
            pc[s >> 2] |= (uint32_t)(bc & 0xff) << (8 * (s & 3));
            da[s] = make_desc(a, bs);
            db[s] = make_desc(a, bc);
        }
        *(uint2 *)(sysI_g + row + 8 * c) = make_uint2(ps[0], ps[1]);
        *(uint2 *)(c1I_g + row + 8 * c) = make_uint2(pc[0], pc[1]);
#pragma unroll
        for (int h = 0; h < 4; h++) {
            tile2[(size_t)(4 * c + h) * n_tr + 2 * cb + 0] = make_uint4(da[2 * h].x, da[2 * h].y, da[2 * h + 1].x, da[2 * h + 1].y);
            tile2[(size_t)(4 * c + h) * n_tr + 2 * cb + 1] = make_uint4(db[2 * h].x, db[2 * h].y, db[2 * h + 1].x, db[2 * h + 1].y);
        }
    }
    wmA = block_max(wmA, sh);
    wmB = block_max(wmB, sh);
    if (threadIdx.x == 0) { wmax2[2 * cb] = wmA; wmax2[2 * cb + 1] = wmB; }
}

// ---------------------------------------------------------------------------------------------------------
// Steps 8-14: actA/actB -> fb -> merge*/merge** -> deinterleave (last writer wins, holes = 0) -> 4-way vote.
// One CTA per code block; dynamic smem = 4*K bytes.
__global__ void __launch_bounds__(PREP_THREADS)
turbo_ref_final(const int8_t *__restrict__ q, const uint32_t *__restrict__ dec2, const uint8_t *__restrict__ st_end2, const int *__restrict__ wmax2,
                const int16_t *__restrict__ inv_last, const int8_t *__restrict__ act1_g, const int8_t *__restrict__ calc1_g,
                const int8_t *__restrict__ sysI_g, const int8_t *__restrict__ c1I_g, int K, int n_cb,
                uint8_t *__restrict__ out /*[n_cb][K]*/)
{
    extern __shared__ int8_t smem[];
    __shared__ uint8_t s_sgnA[TB_WORDS * PREP_THREADS], s_sgnB[TB_WORDS * PREP_THREADS];
    __shared__ uint32_t s_warp[PREP_THREADS / 32];
    int8_t *s_A = smem, *s_B = smem + K, *s_cA = smem + 2 * K, *s_cB = smem + 3 * K;
    const int cb = blockIdx.x;
    const size_t plane = (size_t)n_cb * K, row = (size_t)cb * K;
    const int n_tr = 2 * n_cb;
    const int wA = wmax2[2 * cb], wB = wmax2[2 * cb + 1];
    block_traceback(dec2, n_tr, 2 * cb, K >> 3, st_end2[2 * cb], s_sgnA, s_warp);
    block_traceback(dec2, n_tr, 2 * cb + 1, K >> 3, st_end2[2 * cb + 1], s_sgnB, s_warp);
    const int K8 = K >> 3;
    for (int c = threadIdx.x; c < K8; c += PREP_THREADS) {
        const uint2 va = *(const uint2 *)(q + 2 * plane + row + 8 * c);
        const uint2 vs = *(const uint2 *)(sysI_g + row + 8 * c), vc = *(const uint2 *)(c1I_g + row + 8 * c);
        const uint32_t sgA = s_sgnA[c], sgB = s_sgnB[c];
        uint2 A = make_uint2(0, 0), B = make_uint2(0, 0);
#pragma unroll
        for (int s = 0; s < 8; s++) {
            const int a = iabs(sbyte(va, s));
            const int magA = soft_mag(a + iabs(sbyte(vs, s)), wA), magB = soft_mag(a + iabs(sbyte(vc, s)), wB);
            put_byte(A, s, ((sgA >> s) & 1u) ? -magA : magA);
            put_byte(B, s, ((sgB >> s) & 1u) ? -magB : magB);
        }
        *(uint2 *)(s_A + 8 * c) = A;
        *(uint2 *)(s_B + 8 * c) = B;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < K8; c += PREP_THREADS) {
        const uint2 curA = *(const uint2 *)(s_A + 8 * c), curB = *(const uint2 *)(s_B + 8 * c);
        const uint2 prevA = c > 0 ? *(const uint2 *)(s_A + 8 * c - 8) : make_uint2(FILL127, FILL127);
        const uint2 prevB = c > 0 ? *(const uint2 *)(s_B + 8 * c - 8) : make_uint2(FILL127, FILL127);
        const uint2 a1 = *(const uint2 *)(act1_g + row + 8 * c);
        uint2 cA = make_uint2(0, 0), cB = make_uint2(0, 0);
#pragma unroll
        for (int s = 0; s < 8; s++) {
            const int fbA = sxor(back2(prevA, curA, s), back3(prevA, curA, s)), fbB = sxor(back2(prevB, curB, s), back3(prevB, curB, s));
            put_byte(cA, s, merge<1>(sbyte(curA, s), fbA, sbyte(a1, s)));  // step 10 reads td_in_act_1 (liblte_phy.cc:10791,10794)
            put_byte(cB, s, merge<2>(sbyte(curB, s), fbB, 0));             // step 11's last branch differs in sign (:10816)
        }
        *(uint2 *)(s_cA + 8 * c) = cA;
        *(uint2 *)(s_cB + 8 * c) = cB;
    }
    __syncthreads();
    const bool out_vec = (reinterpret_cast<size_t>(out) & 7) == 0;
    for (int c = threadIdx.x; c < K8; c += PREP_THREADS) {
        const uint4 jv = *(const uint4 *)(inv_last + 8 * c);                // eight int16 positions (-1: nobody writes this element)
        const uint2 v0 = *(const uint2 *)(q + 0 * plane + row + 8 * c), v1 = *(const uint2 *)(calc1_g + row + 8 * c);
        const uint32_t jw[4] = {jv.x, jv.y, jv.z, jv.w};
        uint2 o = make_uint2(0, 0);
#pragma unroll
        for (int s = 0; s < 8; s++) {
            const int j = (int)(int16_t)((jw[s >> 1] >> (16 * (s & 1))) & 0xffffu);
            const int c2 = j >= 0 ? s_cA[j] : 0, c3 = j >= 0 ? s_cB[j] : 0;
            const int tot = sbyte(v0, s) + sbyte(v1, s) + c2 + c3;
            put_byte(o, s, tot >= 0 ? 0 : 1);  // liblte_phy.cc:10832-10844
        }
        if (out_vec) *(uint2 *)(out + row + 8 * c) = o;
        else
            for (int s = 0; s < 8; s++) out[row + 8 * c + s] = (uint8_t)sbyte(o, s);
    }
}

}  // namespace

// Host launcher.  Splits very large batches so the scratch arena stays bounded.
// d_llr: float LLRs [n_cb][3 (K + 4)], or null with d_q8: quantised images [n_cb][mlm2_q8_stride(K)] (see turbo_ref_prep_q8)
int turbo_ref_decode_any(lteb200_ctx *ctx, const float *d_llr, const int8_t *d_q8, int K, int n_cb, uint8_t *d_bits, cudaStream_t stream)
{
    if (K < 40 || (K & 7)) { lteb200_set_error(ctx, "turbo: illegal K=%d", K); return LTEB200_ERR_ARGS; }
    QppTables qt;
    int rc = lteb200_get_qpp(ctx, K, &qt);
    if (rc) return rc;
    const int K8 = K / 8;
    const int max_chunk = 8192;
    const int chunk_cap = n_cb < max_chunk ? n_cb : max_chunk;
    const size_t nK = (size_t)chunk_cap * K;
    size_t bytes = 0;
    bytes += Bump::pad(3 * nK) + 4 * Bump::pad(nK);                     // q, act1, calc1, sysI, c1I
    bytes += Bump::pad(8 * nK) + Bump::pad(16 * nK);                    // tile1, tile2 (8 B per trellis step)
    bytes += Bump::pad(nK / 2) + Bump::pad(nK);                         // dec1, dec2
    bytes += 2 * Bump::pad(2 * (size_t)chunk_cap);                      // end states
    bytes += 2 * Bump::pad(sizeof(int) * 2 * (size_t)chunk_cap) + 4096;
    // scratch from the stream-ordered pool: calls on different streams of one context get different memory (a context-wide arena would be
    // shared by them), and with the pool's release threshold raised (ctx.cu) steady state allocates nothing
    void *arena = nullptr;
    LTE_CUDA_OK(ctx, cudaMallocAsync(&arena, bytes, stream));
    if (3 * K > 48 * 1024 || 4 * K > 48 * 1024) {
        cudaFuncSetAttribute(turbo_ref_mid, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * K);
        cudaFuncSetAttribute(turbo_ref_final, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * K);
    }
    for (int base = 0; base < n_cb; base += chunk_cap) {
        const int n = (n_cb - base) < chunk_cap ? (n_cb - base) : chunk_cap;
        Bump b(arena);
        int8_t *q = b.take<int8_t>(3 * nK);
        int8_t *act1 = b.take<int8_t>(nK), *calc1 = b.take<int8_t>(nK), *sysI = b.take<int8_t>(nK), *c1I = b.take<int8_t>(nK);
        uint4 *tile1 = (uint4 *)b.take<int8_t>(8 * nK);
        uint4 *tile2 = (uint4 *)b.take<int8_t>(16 * nK);
        uint32_t *dec1 = (uint32_t *)b.take<int8_t>(nK / 2), *dec2 = (uint32_t *)b.take<int8_t>(nK);
        uint8_t *end1 = b.take<uint8_t>(chunk_cap), *end2 = b.take<uint8_t>(2 * (size_t)chunk_cap);
        int *wmax1 = b.take<int>(chunk_cap), *wmax2 = b.take<int>(2 * (size_t)chunk_cap);
        if (d_q8) {
            const int NT = mlm_size_class(K);
            const size_t img_bytes = 3 * 16 * (size_t)NT + 16;
            turbo_ref_prep_q8<<<n, PREP_THREADS, img_bytes, stream>>>(d_q8 + (size_t)base * img_bytes, K, NT, n, q, tile1, wmax1);
        } else {
            const float *llr = d_llr + (size_t)base * 3 * (K + 4);
            turbo_ref_prep<<<n, PREP_THREADS, 0, stream>>>(llr, K, n, q, tile1, wmax1);
        }
        turbo_ref_siso<<<ceil_div(n, 32), 32, 0, stream>>>(tile1, K8, n, dec1, end1);
        turbo_ref_mid<<<n, PREP_THREADS, 3 * K, stream>>>(q, dec1, end1, wmax1, qt.pi, K, n, act1, calc1, sysI, c1I, tile2, wmax2);
        turbo_ref_siso<<<ceil_div(2 * n, 32), 32, 0, stream>>>(tile2, K8, 2 * n, dec2, end2);
        turbo_ref_final<<<n, PREP_THREADS, 4 * K, stream>>>(q, dec2, end2, wmax2, qt.inv_last, act1, calc1, sysI, c1I, K, n,
                                                            d_bits + (size_t)base * K);
        ctx->launches += 5;
    }
    cudaFreeAsync(arena, stream);
    LTE_CUDA_OK(ctx, cudaGetLastError());
    return LTEB200_OK;
}

int turbo_ref_decode_batch(lteb200_ctx *ctx, const float *d_llr, int K, int n_cb, uint8_t *d_bits, cudaStream_t stream)
{
    return turbo_ref_decode_any(ctx, d_llr, nullptr, K, n_cb, d_bits, stream);
}
int turbo_ref_decode_q8_batch(lteb200_ctx *ctx, const int8_t *d_q8, int K, int n_cb, uint8_t *d_bits, cudaStream_t stream)
{
    return turbo_ref_decode_any(ctx, nullptr, d_q8, K, n_cb, d_bits, stream);
}

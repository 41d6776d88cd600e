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
//   * inputs are staged in a step-major, trellis-minor tile layout (16 B = 8 steps per trellis) so that a
//     warp's loads are 512 B contiguous; survivor information is 4 compare bits per step (the traceback of
//     the reference consults stored metrics of the two predecessors, not survivor decisions), 3 KB per
//     K=6144 trellis instead of the reference's 394 KB of metric history;
//   * the soft output magnitude is state independent (127*W_t/W_max, W_t = |a_t|+|b_t|), so it is produced
//     elementwise outside the serial loop;
//   * the three SISO passes run as two launches (pass 1: n_cb trellises; passes 2 and 3 are independent of
//     each other and run together as 2*n_cb trellises).
// No tensor cores: there is no dense contraction anywhere on this path.
#include "common.cuh"

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
               uint4 *__restrict__ tile1 /*[K/8][n_cb]*/, int *__restrict__ wmax1 /*[n_cb]*/)
{
    __shared__ int sh[32];
    const int cb = blockIdx.x;
    const float *d = llr + (size_t)cb * 3 * (K + 4);
    float m = 0.f;
    for (int idx = threadIdx.x; idx < 3 * K; idx += PREP_THREADS) {
        float v = d[idx];
        if (v == LTEB200_RX_NULL_BIT) v = 0.f;
        m = fmaxf(m, fabsf(v));
    }
    const float maxv = __int_as_float(block_max(__float_as_int(m), sh));  // non-negative floats order like ints
    int wm = 0;
    const int K8 = K >> 3;
    for (int c = threadIdx.x; c < K8; c += PREP_THREADS) {
        uint32_t w[4];
        uint32_t p0[2] = {0, 0}, p1[2] = {0, 0}, p2[2] = {0, 0};
#pragma unroll
        for (int s = 0; s < 8; s++) {
            const int i = 8 * c + s;
            float v0 = d[3 * i + 0], v1 = d[3 * i + 1], v2 = d[3 * i + 2];
            if (v0 == LTEB200_RX_NULL_BIT) v0 = 0.f;
            if (v1 == LTEB200_RX_NULL_BIT) v1 = 0.f;
            if (v2 == LTEB200_RX_NULL_BIT) v2 = 0.f;
            const int a0 = q8(v0, maxv), a1 = q8(v1, maxv), a2 = q8(v2, maxv);
            wm = max(wm, iabs(a1) + iabs(a0));
            p0[s >> 2] |= (uint32_t)(a0 & 0xff) << (8 * (s & 3));
            p1[s >> 2] |= (uint32_t)(a1 & 0xff) << (8 * (s & 3));
            p2[s >> 2] |= (uint32_t)(a2 & 0xff) << (8 * (s & 3));
            const uint32_t pr = (uint32_t)(a1 & 0xff) | ((uint32_t)(a0 & 0xff) << 8);
            if (s & 1) w[s >> 1] |= pr << 16; else w[s >> 1] = pr;
        }
        const size_t plane = (size_t)n_cb * K;
        *(uint2 *)(q + 0 * plane + (size_t)cb * K + 8 * c) = make_uint2(p0[0], p0[1]);
        *(uint2 *)(q + 1 * plane + (size_t)cb * K + 8 * c) = make_uint2(p1[0], p1[1]);
        *(uint2 *)(q + 2 * plane + (size_t)cb * K + 8 * c) = make_uint2(p2[0], p2[1]);
        tile1[(size_t)c * n_cb + cb] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    wm = block_max(wm, sh);
    if (threadIdx.x == 0) wmax1[cb] = wm;
}

// ---------------------------------------------------------------------------------------------------------
// The serial core: forward add-compare-select over K steps, then traceback.  One thread per trellis.
//
// States j = 0..7; predecessors of j are 2j mod 8 and 2j+1 mod 8, i.e. four butterflies m = 0..3 with
// predecessors (2m, 2m+1) and successors (m, m+4).  With sa = (a >= 0 ? +1 : -1), sb likewise, A = sa+sb,
// B = sb-sa, the branch value of butterfly m is M = {-A, +B, -B, +A}[m] (from the generator/register
// definition at liblte_phy.cc:10379-10408):   successor m   : +M from 2m, -M from 2m+1
//                                             successor m+4 : -M from 2m, +M from 2m+1
// Second predecessor wins iff its unweighted sum is strictly smaller; the weighted sum is stored (10452-10463).
struct Acs {
    int P[8];
    __device__ __forceinline__ uint32_t step(int a, int b)
    {
        const int sa = a >= 0 ? 1 : -1, sb = b >= 0 ? 1 : -1;
        const int W = iabs(a) + iabs(b);
        const int A = sa + sb, B = sb - sa;
        const int WA = W * A, WB = W * B, A2 = 2 * A, B2 = 2 * B;
        int N[8];
        uint32_t cb = 0;
        // butterfly 0: M = -A
        {
            const int p0 = P[0], p1 = P[1], D = p0 - p1;
            cb |= (D > 0) ? 1u : 0u;
            N[0] = (D > A2) ? p1 + WA : p0 - WA;   // D + 2M > 0, M = -A
            N[4] = (D > -A2) ? p1 - WA : p0 + WA;  // D - 2M > 0
        }
        // butterfly 1: M = +B
        {
            const int p0 = P[2], p1 = P[3], D = p0 - p1;
            cb |= (D > 0) ? 2u : 0u;
            N[1] = (D > -B2) ? p1 - WB : p0 + WB;
            N[5] = (D > B2) ? p1 + WB : p0 - WB;
        }
        // butterfly 2: M = -B
        {
            const int p0 = P[4], p1 = P[5], D = p0 - p1;
            cb |= (D > 0) ? 4u : 0u;
            N[2] = (D > B2) ? p1 + WB : p0 - WB;
            N[6] = (D > -B2) ? p1 - WB : p0 + WB;
        }
        // butterfly 3: M = +A
        {
            const int p0 = P[6], p1 = P[7], D = p0 - p1;
            cb |= (D > 0) ? 8u : 0u;
            N[3] = (D > -A2) ? p1 - WA : p0 + WA;
            N[7] = (D > A2) ? p1 + WA : p0 - WA;
        }
#pragma unroll
        for (int j = 0; j < 8; j++) P[j] = N[j];
        return cb;
    }
};

__global__ void __launch_bounds__(32)
turbo_ref_siso(const uint4 *__restrict__ tile /*[K8][n_tr]*/, int K8, int n_tr, uint32_t *__restrict__ dec /*[K8][n_tr]*/,
               uint8_t *__restrict__ sgn /*[K8][n_tr]*/)
{
    const int t = blockIdx.x * 32 + threadIdx.x;
    if (t >= n_tr) return;
    Acs acs;
#pragma unroll
    for (int j = 0; j < 8; j++) acs.P[j] = 0;  // all-zero start metrics: no known start state (10411-10418)
    uint4 cur = tile[t];
    for (int c = 0; c < K8; c++) {
        const uint4 nxt = tile[(size_t)min(c + 1, K8 - 1) * n_tr + t];  // prefetch
        uint32_t dw = 0;
        const uint32_t w[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
        for (int h = 0; h < 4; h++) {
            const int a0 = (int)(int8_t)(w[h] & 0xff), b0 = (int)(int8_t)((w[h] >> 8) & 0xff);
            const int a1 = (int)(int8_t)((w[h] >> 16) & 0xff), b1 = (int)(int8_t)(w[h] >> 24);
            dw |= acs.step(a0, b0) << (8 * h);
            dw |= acs.step(a1, b1) << (8 * h + 4);
        }
        dec[(size_t)c * n_tr + t] = dw;
        cur = nxt;
    }
    // end state: first strict minimum over j = 0..7 (10467-10481)
    int st = 0, best = acs.P[0];
#pragma unroll
    for (int j = 1; j < 8; j++)
        if (acs.P[j] < best) { best = acs.P[j]; st = j; }
    // traceback (10484-10503): from state s at time i+1 go to 2(s&3)+1 if metric[2(s&3)] > metric[2(s&3)+1] at time i.
    // Output bit of the transition INTO s is its MSB (10506-10528).
    uint32_t dw = dec[(size_t)(K8 - 1) * n_tr + t];
    for (int c = K8 - 1; c >= 0; c--) {
        const uint32_t nxt = dec[(size_t)max(c - 1, 0) * n_tr + t];
        uint32_t sb = 0;
#pragma unroll
        for (int s = 7; s >= 0; s--) {
            sb |= (uint32_t)(st >> 2) << s;
            const int m = st & 3;
            st = 2 * m + (int)((dw >> (4 * s + m)) & 1u);
        }
        sgn[(size_t)c * n_tr + t] = (uint8_t)sb;
        dw = nxt;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Steps 2-7 front end: act1 -> fb1 -> calc1, interleave q0 and calc1, build the tiles of passes 2 and 3.
// One CTA per code block; dynamic smem = 3*K bytes.
__global__ void __launch_bounds__(PREP_THREADS)
turbo_ref_mid(const int8_t *__restrict__ q, const uint8_t *__restrict__ sgn1, const int *__restrict__ wmax1,
              const uint16_t *__restrict__ pi, int K, int n_cb, int8_t *__restrict__ act1_g, int8_t *__restrict__ calc1_g,
              int8_t *__restrict__ sysI_g, int8_t *__restrict__ c1I_g, uint4 *__restrict__ tile2 /*[K/8][2*n_cb]*/,
              int *__restrict__ wmax2 /*[2*n_cb]*/)
{
    extern __shared__ int8_t smem[];
    __shared__ int sh[32];
    int8_t *s_q0 = smem, *s_act = smem + K, *s_calc = smem + 2 * K;
    const int cb = blockIdx.x;
    const size_t plane = (size_t)n_cb * K, row = (size_t)cb * K;
    const int wmx = wmax1[cb];
    for (int i = threadIdx.x; i < K; i += PREP_THREADS) {
        const int b = q[0 * plane + row + i], a = q[1 * plane + row + i];
        const int mag = soft_mag(iabs(a) + iabs(b), wmx);
        const int bit = (sgn1[(size_t)(i >> 3) * n_cb + cb] >> (i & 7)) & 1;
        s_q0[i] = (int8_t)b;
        s_act[i] = (int8_t)(bit ? -mag : mag);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += PREP_THREADS) {
        // fb[0] = 127, fb[i] = sxor(act[i-2], act[i-3]) with act[<0] = +127 (conv_encode_soft, register preset to 127)
        const int act = s_act[i];
        const int fb = (i == 0) ? 127 : sxor(i >= 2 ? s_act[i - 2] : 127, i >= 3 ? s_act[i - 3] : 127);
        const int c1 = merge<0>(act, fb, 0);
        s_calc[i] = (int8_t)c1;
        act1_g[row + i] = (int8_t)act;
        calc1_g[row + i] = (int8_t)c1;
    }
    __syncthreads();
    int wmA = 0, wmB = 0;
    const int K8 = K >> 3, n_tr = 2 * n_cb;
    for (int c = threadIdx.x; c < K8; c += PREP_THREADS) {
        uint32_t wa[4], wb[4], ps[2] = {0, 0}, pc[2] = {0, 0};
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
            const uint32_t pa = (uint32_t)(a & 0xff) | ((uint32_t)(bs & 0xff) << 8);
            const uint32_t pb = (uint32_t)(a & 0xff) | ((uint32_t)(bc & 0xff) << 8);
            if (s & 1) { wa[s >> 1] |= pa << 16; wb[s >> 1] |= pb << 16; } else { wa[s >> 1] = pa; wb[s >> 1] = pb; }
        }
        *(uint2 *)(sysI_g + row + 8 * c) = make_uint2(ps[0], ps[1]);
        *(uint2 *)(c1I_g + row + 8 * c) = make_uint2(pc[0], pc[1]);
        tile2[(size_t)c * n_tr + 2 * cb + 0] = make_uint4(wa[0], wa[1], wa[2], wa[3]);
        tile2[(size_t)c * n_tr + 2 * cb + 1] = make_uint4(wb[0], wb[1], wb[2], wb[3]);
    }
    wmA = block_max(wmA, sh);
    wmB = block_max(wmB, sh);
    if (threadIdx.x == 0) { wmax2[2 * cb] = wmA; wmax2[2 * cb + 1] = wmB; }
}

// ---------------------------------------------------------------------------------------------------------
// Steps 8-14: actA/actB -> fb -> merge*/merge** -> deinterleave (last writer wins, holes = 0) -> 4-way vote.
// One CTA per code block; dynamic smem = 4*K bytes.
__global__ void __launch_bounds__(PREP_THREADS)
turbo_ref_final(const int8_t *__restrict__ q, const uint8_t *__restrict__ sgn2, const int *__restrict__ wmax2,
                const int16_t *__restrict__ inv_last, const int8_t *__restrict__ act1_g, const int8_t *__restrict__ calc1_g,
                const int8_t *__restrict__ sysI_g, const int8_t *__restrict__ c1I_g, int K, int n_cb,
                uint8_t *__restrict__ out /*[n_cb][K]*/)
{
    extern __shared__ int8_t smem[];
    int8_t *s_A = smem, *s_B = smem + K, *s_cA = smem + 2 * K, *s_cB = smem + 3 * K;
    const int cb = blockIdx.x;
    const size_t plane = (size_t)n_cb * K, row = (size_t)cb * K;
    const int n_tr = 2 * n_cb;
    const int wA = wmax2[2 * cb], wB = wmax2[2 * cb + 1];
    for (int i = threadIdx.x; i < K; i += PREP_THREADS) {
        const int a = q[2 * plane + row + i];
        const int bs = sysI_g[row + i], bc = c1I_g[row + i];
        const int magA = soft_mag(iabs(a) + iabs(bs), wA), magB = soft_mag(iabs(a) + iabs(bc), wB);
        const int bitA = (sgn2[(size_t)(i >> 3) * n_tr + 2 * cb] >> (i & 7)) & 1;
        const int bitB = (sgn2[(size_t)(i >> 3) * n_tr + 2 * cb + 1] >> (i & 7)) & 1;
        s_A[i] = (int8_t)(bitA ? -magA : magA);
        s_B[i] = (int8_t)(bitB ? -magB : magB);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += PREP_THREADS) {
        const int fbA = (i == 0) ? 127 : sxor(i >= 2 ? s_A[i - 2] : 127, i >= 3 ? s_A[i - 3] : 127);
        const int fbB = (i == 0) ? 127 : sxor(i >= 2 ? s_B[i - 2] : 127, i >= 3 ? s_B[i - 3] : 127);
        s_cA[i] = (int8_t)merge<1>(s_A[i], fbA, act1_g[row + i]);  // step 10 reads td_in_act_1 (liblte_phy.cc:10791,10794)
        s_cB[i] = (int8_t)merge<2>(s_B[i], fbB, 0);                // step 11's last branch differs in sign (:10816)
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += PREP_THREADS) {
        const int j = inv_last[i];
        const int c2 = j >= 0 ? s_cA[j] : 0, c3 = j >= 0 ? s_cB[j] : 0;
        const int tot = (int)q[0 * plane + row + i] + (int)calc1_g[row + i] + c2 + c3;
        out[row + i] = tot >= 0 ? 0 : 1;  // liblte_phy.cc:10832-10844
    }
}

}  // namespace

// Host launcher.  Splits very large batches so the scratch arena stays bounded.
int turbo_ref_decode_batch(lteb200_ctx *ctx, const float *d_llr, int K, int n_cb, uint8_t *d_bits, cudaStream_t stream)
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
    bytes += Bump::pad(2 * nK) + Bump::pad(4 * nK);                     // tile1, tile2
    bytes += Bump::pad(nK / 2) + Bump::pad(nK);                         // dec1, dec2
    bytes += Bump::pad(nK / 8) + Bump::pad(nK / 4);                     // sgn1, sgn2
    bytes += 2 * Bump::pad(sizeof(int) * 2 * (size_t)chunk_cap) + 4096;
    void *arena = nullptr;
    rc = lteb200_arena(ctx, bytes, &arena);
    if (rc) return rc;
    if (3 * K > 48 * 1024 || 4 * K > 48 * 1024) {
        cudaFuncSetAttribute(turbo_ref_mid, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * K);
        cudaFuncSetAttribute(turbo_ref_final, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * K);
    }
    for (int base = 0; base < n_cb; base += chunk_cap) {
        const int n = (n_cb - base) < chunk_cap ? (n_cb - base) : chunk_cap;
        Bump b(arena);
        int8_t *q = b.take<int8_t>(3 * nK);
        int8_t *act1 = b.take<int8_t>(nK), *calc1 = b.take<int8_t>(nK), *sysI = b.take<int8_t>(nK), *c1I = b.take<int8_t>(nK);
        uint4 *tile1 = (uint4 *)b.take<int8_t>(2 * nK);
        uint4 *tile2 = (uint4 *)b.take<int8_t>(4 * nK);
        uint32_t *dec1 = (uint32_t *)b.take<int8_t>(nK / 2), *dec2 = (uint32_t *)b.take<int8_t>(nK);
        uint8_t *sgn1 = b.take<uint8_t>(nK / 8), *sgn2 = b.take<uint8_t>(nK / 4);
        int *wmax1 = b.take<int>(chunk_cap), *wmax2 = b.take<int>(2 * (size_t)chunk_cap);
        const float *llr = d_llr + (size_t)base * 3 * (K + 4);
        turbo_ref_prep<<<n, PREP_THREADS, 0, stream>>>(llr, K, n, q, tile1, wmax1);
        turbo_ref_siso<<<ceil_div(n, 32), 32, 0, stream>>>(tile1, K8, n, dec1, sgn1);
        turbo_ref_mid<<<n, PREP_THREADS, 3 * K, stream>>>(q, sgn1, wmax1, qt.pi, K, n, act1, calc1, sysI, c1I, tile2, wmax2);
        turbo_ref_siso<<<ceil_div(2 * n, 32), 32, 0, stream>>>(tile2, K8, 2 * n, dec2, sgn2);
        turbo_ref_final<<<n, PREP_THREADS, 4 * K, stream>>>(q, sgn2, wmax2, qt.inv_last, act1, calc1, sysI, c1I, K, n,
                                                            d_bits + (size_t)base * K);
        ctx->launches += 5;
    }
    LTE_CUDA_OK(ctx, cudaGetLastError());
    return LTEB200_OK;
}

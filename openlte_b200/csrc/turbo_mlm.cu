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
// (LLRs int8, extrinsics int16, forward metrics of every window int16: 192 KB at K = 6144, so one CTA of
// 192 threads per SM); one thread per 32-step window holds the 8 forward/backward state metrics in
// registers (int32, DPX add-max), windows are coupled by state-metric propagation from the previous
// iteration, so all K/32 windows of a block advance in parallel without warm-up recursions.  Shared-memory
// arrays use a window-transposed layout pos(k) = (k mod 32)*nw + k/32 so that at every step the threads of a
// warp touch consecutive addresses (conflict free).  No tensor cores (no contraction on this path).
#include "common.cuh"

namespace {

constexpr int WIN = 32;
constexpr int NEG = -(1 << 28);
constexpr int LE_LIM = 1023, LA_LIM = 2047;

__host__ __device__ constexpr int tr_a(int S, int u) { return u ^ ((S >> 1) & 1) ^ (S & 1); }
__host__ __device__ constexpr int tr_par(int S, int u) { return tr_a(S, u) ^ ((S >> 2) & 1) ^ (S & 1); }
__host__ __device__ constexpr int tr_nxt(int S, int u) { return (tr_a(S, u) << 2) | (S >> 1); }

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return max(lo, min(v, hi)); }

struct Gam {
    int g[2][2];  // [u][p]
    __device__ __forceinline__ Gam(int ls, int lp)
    {
        g[0][0] = ls + lp; g[0][1] = ls; g[1][0] = lp; g[1][1] = 0;
    }
};

__device__ __forceinline__ int qf(float v, float maxv)
{
    return __float2int_rz(__fdiv_rn(__fmul_rn(v, 127.0f), maxv));  // NaN -> 0
}

struct Pass {
    // per-block shared arrays (window-transposed)
    const int16_t *Lsa;   // systematic + a-priori LLR, prepared before the pass
    const int8_t *Lp;
    int16_t *LeOut;
    int4 *st;             // [32][nw] packed 8 x int16
    int *bA, *bB;         // [nw][8]
    int xt[3], zt[3];
};

__device__ __forceinline__ int pack_sat(int lo, int hi)
{
    int r;
    asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(r) : "r"(hi), "r"(lo));  // {sat16(hi), sat16(lo)}
    return r;
}

__device__ __forceinline__ void dec_pass(const Pass &P, int w, int nw, int K)
{
    const bool active = w < nw;
    const int n = active ? min(WIN, K - WIN * w) : 0;
    int a[8], b[8];
    if (active) {
        if (w == 0) {
            a[0] = 0;
#pragma unroll
            for (int s = 1; s < 8; s++) a[s] = NEG;
        } else {
#pragma unroll
            for (int s = 0; s < 8; s++) a[s] = P.bA[w * 8 + s];
        }
        if (w == nw - 1) {
            // beta at time K: three forced termination steps ending in state 0
            b[0] = 0;
#pragma unroll
            for (int s = 1; s < 8; s++) b[s] = NEG;
#pragma unroll
            for (int t = 2; t >= 0; t--) {
                int nb[8];
                const Gam G(P.xt[t], P.zt[t]);
#pragma unroll
                for (int S = 0; S < 8; S++) {
                    const int u = ((S >> 1) & 1) ^ (S & 1);
                    nb[S] = max(G.g[u][tr_par(S, u)] + b[tr_nxt(S, u)], NEG);
                }
#pragma unroll
                for (int S = 0; S < 8; S++) b[S] = nb[S];
            }
        } else {
#pragma unroll
            for (int s = 0; s < 8; s++) b[s] = P.bB[w * 8 + s];
        }
    }
    __syncthreads();  // all boundaries of the previous iteration are read before any is overwritten

    // ---- forward (inputs of step i+1 are fetched while step i computes)
    int ls_n = 0, lp_n = 0;
    if (n > 0) { ls_n = P.Lsa[w]; lp_n = P.Lp[w]; }
    for (int i = 0; i < n; i++) {
        const int p = i * nw + w;
        const int ls = ls_n, lp = lp_n;
        if (i + 1 < n) { ls_n = P.Lsa[p + nw]; lp_n = P.Lp[p + nw]; }
        {
            const int a0 = a[0];
            int4 pk;
            pk.x = pack_sat(0, a[1] - a0); pk.y = pack_sat(a[2] - a0, a[3] - a0);
            pk.z = pack_sat(a[4] - a0, a[5] - a0); pk.w = pack_sat(a[6] - a0, a[7] - a0);
            P.st[p] = pk;
        }
        const Gam G(ls, lp);
        int na[8];
#pragma unroll
        for (int d = 0; d < 8; d++) {
            // predecessors of d: S = 2*(d&3) + s3 ; the input that leads S -> d is u = (d>>2) ^ s2 ^ s3
            const int S0 = 2 * (d & 3), S1 = S0 + 1;
            const int u0 = (d >> 2) ^ ((S0 >> 1) & 1) ^ (S0 & 1), u1 = (d >> 2) ^ ((S1 >> 1) & 1) ^ (S1 & 1);
            na[d] = __viaddmax_s32(a[S0], G.g[u0][tr_par(S0, u0)], a[S1] + G.g[u1][tr_par(S1, u1)]);
        }
#pragma unroll
        for (int s = 0; s < 8; s++) a[s] = na[s];
    }
    if (active && w + 1 < nw) {
        int mx = a[0];
#pragma unroll
        for (int s = 1; s < 8; s++) mx = max(mx, a[s]);
#pragma unroll
        for (int s = 0; s < 8; s++) P.bA[(w + 1) * 8 + s] = max(a[s] - mx, NEG);
    }
    // ---- backward + LLR
    int4 pk_n = make_int4(0, 0, 0, 0);
    if (n > 0) { const int p = (n - 1) * nw + w; ls_n = P.Lsa[p]; lp_n = P.Lp[p]; pk_n = P.st[p]; }
    for (int i = n - 1; i >= 0; i--) {
        const int p = i * nw + w;
        const int ls = ls_n, lp = lp_n;
        const int4 pk = pk_n;
        if (i > 0) { ls_n = P.Lsa[p - nw]; lp_n = P.Lp[p - nw]; pk_n = P.st[p - nw]; }
        const Gam G(ls, lp);
        int al[8];
        al[0] = (int)(short)(pk.x & 0xffff); al[1] = pk.x >> 16;
        al[2] = (int)(short)(pk.y & 0xffff); al[3] = pk.y >> 16;
        al[4] = (int)(short)(pk.z & 0xffff); al[5] = pk.z >> 16;
        al[6] = (int)(short)(pk.w & 0xffff); al[7] = pk.w >> 16;
        int nb[8], m0 = 0, m1 = 0;
#pragma unroll
        for (int S = 0; S < 8; S++) {
            const int g0 = G.g[0][tr_par(S, 0)] + b[tr_nxt(S, 0)];
            const int g1 = G.g[1][tr_par(S, 1)] + b[tr_nxt(S, 1)];
            nb[S] = max(g0, g1);
            if (S == 0) { m0 = al[0] + g0; m1 = al[0] + g1; }
            else { m0 = __viaddmax_s32(al[S], g0, m0); m1 = __viaddmax_s32(al[S], g1, m1); }
        }
#pragma unroll
        for (int S = 0; S < 8; S++) b[S] = nb[S];
        const int L = m0 - m1;
        P.LeOut[p] = (int16_t)clampi(((L - ls) * 3) >> 2, -LE_LIM, LE_LIM);
    }
    if (active && w > 0) {
        int mx = b[0];
#pragma unroll
        for (int s = 1; s < 8; s++) mx = max(mx, b[s]);
#pragma unroll
        for (int s = 0; s < 8; s++) P.bB[(w - 1) * 8 + s] = max(b[s] - mx, NEG);
    }
    __syncthreads();
}

// piT  [32*nw] : pos(pi(k)) stored at pos(k)           (global, per K)
// csr_start [K+1], csr_pos [K] : for natural index j, positions pos(i) of all i with pi(i) == j   (global, per K)
__global__ void __launch_bounds__(192)
turbo_mlm_kernel(const float *__restrict__ llr, int K, int n_iter, const uint16_t *__restrict__ piT,
                 const uint16_t *__restrict__ csr_start, const uint16_t *__restrict__ csr_pos, uint8_t *__restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nw = (K + WIN - 1) / WIN;
    const int NP = WIN * nw;  // padded length
    int4 *st = (int4 *)smem_raw;                                   // 16*NP bytes
    int *bA1 = (int *)(smem_raw + (size_t)16 * NP);                // 4 x nw*8 ints
    int *bB1 = bA1 + nw * 8, *bA2 = bB1 + nw * 8, *bB2 = bA2 + nw * 8;
    int16_t *Lsa1 = (int16_t *)(bB2 + nw * 8);   // Ls + a-priori of decoder 1 (natural order)
    int16_t *Lsa2 = Lsa1 + NP, *Le1 = Lsa2 + NP, *Le2 = Le1 + NP;
    int8_t *Ls = (int8_t *)(Le2 + NP);
    int8_t *LsI = Ls + NP, *Lp1 = LsI + NP, *Lp2 = Lp1 + NP;
    __shared__ float s_red[32];
    __shared__ int s_tail[12];

    const int cb = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const float *d = llr + (size_t)cb * 3 * (K + 4);
    // ---- max |d| over all 3(K+4) inputs (sentinel -> 0)
    float m = 0.f;
    for (int idx = tid; idx < 3 * (K + 4); idx += nt) {
        float v = d[idx];
        if (v == LTEB200_RX_NULL_BIT) v = 0.f;
        m = fmaxf(m, fabsf(v));
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((tid & 31) == 0) s_red[tid >> 5] = m;
    __syncthreads();
    if (tid < 32) {
        float v = tid < (nt >> 5) ? s_red[tid] : 0.f;
        for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (tid == 0) s_red[0] = v;
    }
    __syncthreads();
    const float maxv = s_red[0];
    // ---- quantise into the transposed arrays; zero the padding
    for (int k = tid; k < NP; k += nt) {
        const int p = (k & 31) * nw + (k >> 5);
        int q0 = 0, q1 = 0, q2 = 0;
        if (k < K) {
            float v0 = d[3 * k], v1 = d[3 * k + 1], v2 = d[3 * k + 2];
            if (v0 == LTEB200_RX_NULL_BIT) v0 = 0.f;
            if (v1 == LTEB200_RX_NULL_BIT) v1 = 0.f;
            if (v2 == LTEB200_RX_NULL_BIT) v2 = 0.f;
            q0 = qf(v0, maxv); q1 = qf(v1, maxv); q2 = qf(v2, maxv);
        }
        Ls[p] = (int8_t)q0; Lp1[p] = (int8_t)q1; Lp2[p] = (int8_t)q2;
        Lsa1[p] = (int16_t)q0; Lsa2[p] = 0; Le1[p] = 0; Le2[p] = 0;
    }
    if (tid < 12) {
        float v = d[3 * K + tid];
        if (v == LTEB200_RX_NULL_BIT) v = 0.f;
        s_tail[tid] = qf(v, maxv);  // s_tail[3*r + x] = q(d_x[K+r])
    }
    for (int i = tid; i < nw * 8 * 4; i += nt) bA1[i] = 0;  // bA1,bB1,bA2,bB2 are contiguous
    __syncthreads();
    for (int p = tid; p < NP; p += nt) LsI[p] = (p < NP) ? Ls[piT[p]] : 0;
    __syncthreads();

    Pass P1, P2;
    P1.Lsa = Lsa1; P1.Lp = Lp1; P1.LeOut = Le1; P1.st = st; P1.bA = bA1; P1.bB = bB1;
    P2.Lsa = Lsa2; P2.Lp = Lp2; P2.LeOut = Le2; P2.st = st; P2.bA = bA2; P2.bB = bB2;
    // tail multiplexing (liblte_phy.cc:10575-10586): x = systematic/feedback bits, z = parity bits of the 3 tail steps
    P1.xt[0] = s_tail[0]; P1.xt[1] = s_tail[2]; P1.xt[2] = s_tail[4];
    P1.zt[0] = s_tail[1]; P1.zt[1] = s_tail[3]; P1.zt[2] = s_tail[5];
    P2.xt[0] = s_tail[6]; P2.xt[1] = s_tail[8]; P2.xt[2] = s_tail[10];
    P2.zt[0] = s_tail[7]; P2.zt[1] = s_tail[9]; P2.zt[2] = s_tail[11];

    for (int it = 0; it < n_iter; it++) {
        dec_pass(P1, tid, nw, K);
        for (int p = tid; p < NP; p += nt) Lsa2[p] = (int16_t)((int)LsI[p] + (int)Le1[piT[p]]);   // interleave the extrinsic
        __syncthreads();
        dec_pass(P2, tid, nw, K);
        for (int j = tid; j < K; j += nt) {
            int acc = 0;
            const int e = csr_start[j + 1];
            for (int t = csr_start[j]; t < e; t++) acc += (int)Le2[csr_pos[t]];
            const int p = (j & 31) * nw + (j >> 5);
            Lsa1[p] = (int16_t)((int)Ls[p] + clampi(acc, -LA_LIM, LA_LIM));                      // de-interleave, combine
        }
        __syncthreads();
    }
    for (int j = tid; j < K; j += nt) {
        const int p = (j & 31) * nw + (j >> 5);
        out[(size_t)cb * K + j] = ((int)Lsa1[p] + (int)Le1[p] >= 0) ? 0 : 1;
    }
}

}  // namespace

int turbo_mlm_decode_batch(lteb200_ctx *ctx, const float *d_llr, int K, int n_cb, int n_iter, uint8_t *d_bits, cudaStream_t stream)
{
    if (K < 40 || (K & 7) || K > 6144) { lteb200_set_error(ctx, "turbo: illegal K=%d", K); return LTEB200_ERR_ARGS; }
    if (n_iter < 1) n_iter = 1;
    QppTables qt;
    int rc = lteb200_get_qpp(ctx, K, &qt);  // validates K
    if (rc) return rc;
    const int nw = (K + WIN - 1) / WIN, NP = WIN * nw;
    MlmTables mt;
    {
        auto it = ctx->mlm.find(K);
        if (it == ctx->mlm.end()) {
            std::vector<uint16_t> pi(K);
            LTE_CUDA_OK(ctx, cudaMemcpy(pi.data(), qt.pi, K * sizeof(uint16_t), cudaMemcpyDeviceToHost));
            auto pos = [&](int k) { return (uint16_t)((k & 31) * nw + (k >> 5)); };
            std::vector<uint16_t> piT(NP, 0), cs(K + 1, 0), cp(K);
            for (int k = 0; k < K; k++) piT[pos(k)] = pos(pi[k]);
            std::vector<int> cnt(K, 0);
            for (int i = 0; i < K; i++) cnt[pi[i]]++;
            for (int j = 0; j < K; j++) cs[j + 1] = (uint16_t)(cs[j] + cnt[j]);
            std::vector<int> fill(cs.begin(), cs.end() - 1);
            for (int i = 0; i < K; i++) cp[fill[pi[i]]++] = pos(i);
            LTE_CUDA_OK(ctx, cudaMalloc(&mt.piT, NP * 2));
            LTE_CUDA_OK(ctx, cudaMalloc(&mt.csr_start, (K + 1) * 2));
            LTE_CUDA_OK(ctx, cudaMalloc(&mt.csr_pos, K * 2));
            LTE_CUDA_OK(ctx, cudaMemcpy(mt.piT, piT.data(), NP * 2, cudaMemcpyHostToDevice));
            LTE_CUDA_OK(ctx, cudaMemcpy(mt.csr_start, cs.data(), (K + 1) * 2, cudaMemcpyHostToDevice));
            LTE_CUDA_OK(ctx, cudaMemcpy(mt.csr_pos, cp.data(), K * 2, cudaMemcpyHostToDevice));
            ctx->mlm[K] = mt;
        } else mt = it->second;
    }
    const size_t smem = (size_t)16 * NP + (size_t)4 * nw * 8 * 4 + (size_t)4 * NP * 2 + (size_t)4 * NP;
    LTE_CUDA_OK(ctx, cudaFuncSetAttribute(turbo_mlm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int threads = ((nw + 31) / 32) * 32;
    turbo_mlm_kernel<<<n_cb, threads, smem, stream>>>(d_llr, K, n_iter, mt.piT, mt.csr_start, mt.csr_pos, d_bits);
    ctx->launches += 1;
    LTE_CUDA_OK(ctx, cudaGetLastError());
    return LTEB200_OK;
}

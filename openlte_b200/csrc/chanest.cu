// CRS-based channel estimation for batches of subframes: twin of the estimator inside
// liblte_phy_get_dl_subframe_and_ce, /root/reference/liblte/src/liblte_phy.cc:5959-6193.
//
// Per antenna port: least-squares estimate at the CRS tones of 5 (ports 0/1) or 3 (ports 2/3) symbols, converted
// to polar form; phase unwrapped tone to tone (a dependent chain of 2*N_rb steps per CRS symbol, :6033-6036);
// linear interpolation across the 6-subcarrier gaps by repeated subtraction (:6038-6063), edges extrapolated
// with the neighbouring slope (the upper edge with the reference's sign, :6059-6063); linear interpolation in time
// in polar form, back to Cartesian with cosf/sinf (:6067-6193).  The arithmetic is restated operation by operation
// (IEEE +,-,*,/ and sqrt are bit-identical to the CPU; atan2f/cosf/sinf differ from glibc by a few ulp, which is
// the stated float tolerance of this stage).
//
// One CTA per (subframe, port); magnitudes and angles of the CRS symbols live in shared memory (48 KB).  Reads
// the 5 CRS rows of the grid once, writes 14*12*N_rb float2: HBM-bound.
#include "common.cuh"
#include <math.h>

namespace {

constexpr double PI_D = 3.14159265358979323846;

__device__ __forceinline__ float wrap_phase(float p1, float p2)
{
    // liblte_phy.cc:14105-14116: float difference, compared and corrected in double.  A float f satisfies (double)f >= pi exactly
    // when f >= 0x40490FDB (the float just above pi; the one below, 0x40490FDA, is smaller than pi), so the comparison is done in
    // float; the rarely taken correction stays in double.
    const float PI_UP = __uint_as_float(0x40490FDBu);
    while (__fsub_rn(p1, p2) >= PI_UP) p1 = (float)((double)p1 - 2 * PI_D);
    while (__fsub_rn(p1, p2) <= -PI_UP) p1 = (float)((double)p1 + 2 * PI_D);
    return p1;
}

// CRS symbol i of port p sits in grid symbol crs_sym(p, i) with frequency offset (v + N_id_cell mod 6) mod 6 (liblte_phy.cc:5973-6014)
__device__ __forceinline__ int crs_sym(int p, int i)
{
    if (p < 2) return i == 0 ? 0 : (i == 1 ? 4 : (i == 2 ? 7 : (i == 3 ? 11 : 14)));
    return i == 0 ? 1 : (i == 1 ? 8 : 15);
}
__device__ __forceinline__ int crs_off(int p, int i, int vsh)
{
    int v;
    if (p == 0) v = (i & 1) ? 3 : 0;
    else if (p == 1) v = (i & 1) ? 0 : 3;
    else if (p == 2) v = (i & 1) ? 3 : 0;
    else v = (i & 1) ? 6 : 3;
    return (v + vsh) % 6;
}

struct CeParams {
    int N_rb, nsc, N_ant, N_id_cell;
};

// crs_tab: [20 slots][3 (l = 0,1,4)][220] float2 for this cell.
__global__ void __launch_bounds__(256)
chan_est_kernel(const float2 *__restrict__ grid /*[n_sf][16][nsc]*/, const int *__restrict__ subfr_num, CeParams P,
                const float2 *__restrict__ crs_tab, float2 *__restrict__ ce /*[n_sf][N_ant][14][nsc]*/)
{
    extern __shared__ float smf[];
    float *Mg = smf, *Ag = smf + 5 * 1200;
    const int sf = blockIdx.x / P.N_ant, p = blockIdx.x % P.N_ant;
    const int sn = subfr_num[sf];
    const int vsh = P.N_id_cell % 6;
    const int n_crs = 2 * P.N_rb;
    const int N_sym = p < 2 ? 5 : 3;
    const float2 *g = grid + (size_t)sf * 16 * P.nsc;

    // 1. LS estimate at the CRS tones -> polar
    for (int idx = threadIdx.x; idx < N_sym * n_crs; idx += blockDim.x) {
        const int i = idx / n_crs, j = idx - i * n_crs;
        const int k = 6 * j + crs_off(p, i, vsh);
        const int s = crs_sym(p, i), slot = (2 * sn + s / 7) % 20, l = s % 7;
        const float2 rs = crs_tab[(slot * 3 + (l == 0 ? 0 : (l == 1 ? 1 : 2))) * 220 + j + 110 - P.N_rb];
        const float2 y = g[s * P.nsc + k];
        const float tr = __fadd_rn(__fmul_rn(y.x, rs.x), __fmul_rn(y.y, rs.y));
        const float ti = __fsub_rn(__fmul_rn(y.y, rs.x), __fmul_rn(y.x, rs.y));
        Mg[i * 1200 + k] = __fsqrt_rn(__fadd_rn(__fmul_rn(tr, tr), __fmul_rn(ti, ti)));
        Ag[i * 1200 + k] = atan2f(ti, tr);
    }
    __syncthreads();
    // 2. phase unwrap along frequency (liblte_phy.cc:6033-6036: every tone is brought within pi of its already unwrapped predecessor by adding or
    // subtracting 2 pi, in float, one turn at a time).  The number of turns applied to tone j is the running sum of the turns between neighbours,
    // d_j = -1 / 0 / +1 according to whether the RAW difference A[j] - A[j-1] is >= pi, in between, or <= -pi: a prefix sum instead of a chain of
    // 2 N_rb dependent steps.  One warp per CRS symbol, seven tones per lane; the turns are then applied to the raw value one at a time, which
    // reproduces the reference's sequence of float roundings.  (A decision can differ from the sequential form only when a difference lies within
    // an ulp of pi -- where the reference's own result is arbitrary.)  The round-1 kernel spent 43 % of its warp samples here, 251 of 256 threads
    // waiting at the barrier behind five (profiles/r2_chain_ncu_summary.json).
    {
        const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (w < N_sym) {
            float *A = Ag + w * 1200;
            const int of = crs_off(p, w, vsh);
            constexpr int PER = 7;                                   // 32 x 7 = 224 >= 2 x 100 tones
            const float PI_UP = __uint_as_float(0x40490FDBu);
            const int j0 = lane * PER;
            float v[PER + 1];
#pragma unroll
            for (int u = 0; u <= PER; u++) { const int j = j0 - 1 + u; v[u] = (j >= 0 && j < n_crs) ? A[6 * j + of] : 0.f; }
            int d[PER], tot = 0;
#pragma unroll
            for (int u = 0; u < PER; u++) {
                const int j = j0 + u;
                int t = 0;
                if (j >= 1 && j < n_crs) { const float df = __fsub_rn(v[u + 1], v[u]); t = df >= PI_UP ? -1 : (df <= -PI_UP ? 1 : 0); }
                tot += t; d[u] = tot;                                 // inclusive within the lane
            }
            int excl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int x = __shfl_up_sync(0xffffffffu, excl, o); if (lane >= o) excl += x; }
            excl -= tot;
            __syncwarp();                                             // every raw value has been read
#pragma unroll
            for (int u = 0; u < PER; u++) {
                const int j = j0 + u;
                if (j >= 1 && j < n_crs) {
                    int k = excl + d[u];
                    float ph = v[u + 1];
                    for (; k < 0; k++) ph = (float)((double)ph - 2 * PI_D);
                    for (; k > 0; k--) ph = (float)((double)ph + 2 * PI_D);
                    A[6 * j + of] = ph;
                }
            }
        }
    }
    __syncthreads();
    // 3. frequency interpolation by repeated subtraction
    for (int idx = threadIdx.x; idx < N_sym * (n_crs - 1); idx += blockDim.x) {
        const int i = idx / (n_crs - 1), j = idx - i * (n_crs - 1) + 1;
        float *M = Mg + i * 1200, *A = Ag + i * 1200;
        const int of = crs_off(p, i, vsh);
        const int k = 6 * j + of;
        const float fm = __fdiv_rn(__fsub_rn(M[k], M[k - 6]), 6.0f), fa = __fdiv_rn(__fsub_rn(A[k], A[k - 6]), 6.0f);
        float m = M[k], a = A[k];
#pragma unroll
        for (int z = 1; z < 6; z++) { m = __fsub_rn(m, fm); a = __fsub_rn(a, fa); M[k - z] = m; A[k - z] = a; }
        if (j == 1) {
            m = M[k - 6]; a = A[k - 6];
            for (int z = 1; z < of + 1; z++) { m = __fsub_rn(m, fm); a = __fsub_rn(a, fa); M[k - 6 - z] = m; A[k - 6 - z] = a; }
        }
        if (j == n_crs - 1) {
            m = M[k]; a = A[k];
            for (int z = 1; z < (5 - of) + 1; z++) { m = __fsub_rn(m, fm); a = __fsub_rn(a, fa); M[k + z] = m; A[k + z] = a; }
        }
    }
    __syncthreads();
    // 4. time interpolation, polar -> Cartesian
    float2 *out = ce + ((size_t)sf * P.N_ant + p) * 14 * P.nsc;
    for (int j = threadIdx.x; j < P.nsc; j += blockDim.x) {
        auto put = [&](int z, float m, float a) { float sn_, cs_; sincosf(a, &sn_, &cs_); out[z * P.nsc + j] = make_float2(__fmul_rn(m, cs_), __fmul_rn(m, sn_)); };
        float mg[5], an[5];
#pragma unroll
        for (int i = 0; i < 5; i++) { mg[i] = Mg[i * 1200 + j]; an[i] = Ag[i * 1200 + j]; }
        if (N_sym == 5) {
            put(0, mg[0], an[0]); put(4, mg[1], an[1]); put(7, mg[2], an[2]); put(11, mg[3], an[3]);
#pragma unroll
            for (int sgm = 0; sgm < 4; sgm++) {
                const float div = (sgm & 1) ? 3.0f : 4.0f;
                const int z_hi = (sgm == 0) ? 3 : (sgm == 1) ? 6 : (sgm == 2) ? 10 : 13;
                const int z_lo = (sgm == 0) ? 0 : (sgm == 1) ? 4 : (sgm == 2) ? 7 : 11;
                const float fm = __fdiv_rn(__fsub_rn(mg[sgm + 1], mg[sgm]), div);
                an[sgm + 1] = wrap_phase(an[sgm + 1], an[sgm]);
                float fa = __fsub_rn(an[sgm + 1], an[sgm]);
                fa = wrap_phase(fa, 0.0f);
                fa = __fdiv_rn(fa, div);
                float m = mg[sgm + 1], a = an[sgm + 1];
                for (int z = z_hi; z > z_lo; z--) { m = __fsub_rn(m, fm); a = __fsub_rn(a, fa); put(z, m, a); }
            }
        } else {
            put(1, mg[0], an[0]); put(8, mg[1], an[1]);
            float fm = __fdiv_rn(__fsub_rn(mg[1], mg[0]), 7.0f);
            an[1] = wrap_phase(an[1], an[0]);
            float fa = wrap_phase(__fsub_rn(an[1], an[0]), 0.0f);
            fa = __fdiv_rn(fa, 7.0f);
            float m = mg[1], a = an[1];
            for (int z = 7; z > 1; z--) { m = __fsub_rn(m, fm); a = __fsub_rn(a, fa); put(z, m, a); }
            put(0, __fsub_rn(mg[0], fm), __fsub_rn(an[0], fa));
            fm = __fdiv_rn(__fsub_rn(mg[2], mg[1]), 7.0f);
            an[2] = wrap_phase(an[2], an[1]);
            fa = wrap_phase(__fsub_rn(an[2], an[1]), 0.0f);
            fa = __fdiv_rn(fa, 7.0f);
            m = __fsub_rn(mg[2], fm); a = __fsub_rn(an[2], fa);
            for (int z = 13; z > 8; z--) { m = __fsub_rn(m, fm); a = __fsub_rn(a, fa); put(z, m, a); }
        }
    }
}

// Gold sequence c(n) on the host (generate_prs_c, liblte_phy.cc:9669-9704)
void prs_host(uint32_t c_init, int len, uint8_t *c)
{
    uint32_t x1 = 0x54D21B24u, x2 = c_init;
    for (int i = 0; i < 1600 - 31; i++) x2 = (x2 >> 1) | ((((x2 >> 3) ^ (x2 >> 2) ^ (x2 >> 1) ^ x2) & 1u) << 30);
    for (int i = 0; i < len; i++) {
        const uint32_t n1 = ((x1 >> 3) ^ x1) & 1u, n2 = ((x2 >> 3) ^ (x2 >> 2) ^ (x2 >> 1) ^ x2) & 1u;
        x1 = (x1 >> 1) | (n1 << 30);
        x2 = (x2 >> 1) | (n2 << 30);
        c[i] = (uint8_t)(n1 ^ n2);
    }
}

}  // namespace

static std::map<std::pair<int, int>, float2 *> g_crs;  // (device, N_id_cell) -> [20][3][220]
static std::mutex g_crs_mu;

int get_crs_table(lteb200_ctx *ctx, int N_id_cell, float2 **out)
{
    std::lock_guard<std::mutex> lk(g_crs_mu);
    auto key = std::make_pair(ctx->device, N_id_cell);
    auto it = g_crs.find(key);
    if (it != g_crs.end()) { *out = it->second; return LTEB200_OK; }
    std::vector<float2> h(20 * 3 * 220);
    const float r2 = (float)(1 / sqrt(2.0));
    const int Ls[3] = {0, 1, 4};
    uint8_t c[440];
    for (int ns = 0; ns < 20; ns++)
        for (int li = 0; li < 3; li++) {
            // generate_crs, liblte_phy.cc:8300-8333 (normal CP: N_cp = 1)
            const uint32_t c_init = 1024u * (7 * (ns + 1) + Ls[li] + 1) * (2 * N_id_cell + 1) + 2 * N_id_cell + 1;
            prs_host(c_init, 440, c);
            for (int i = 0; i < 220; i++)
                h[(ns * 3 + li) * 220 + i] = make_float2(r2 * (1 - 2 * (float)c[2 * i]), r2 * (1 - 2 * (float)c[2 * i + 1]));
        }
    float2 *d = nullptr;
    LTE_CUDA_OK(ctx, cudaMalloc(&d, sizeof(float2) * h.size()));
    LTE_CUDA_OK(ctx, cudaMemcpy(d, h.data(), sizeof(float2) * h.size(), cudaMemcpyHostToDevice));
    g_crs[key] = d;
    *out = d;
    return LTEB200_OK;
}

void prs_c_host(uint32_t c_init, int len, uint8_t *c) { prs_host(c_init, len, c); }

int chan_est_launch(lteb200_ctx *ctx, const float *d_grid, const int *d_subfr_num, int N_rb_dl, int N_id_cell, int N_ant, int n_sf,
                    float *d_ce, cudaStream_t stream)
{
    if (!(N_ant == 1 || N_ant == 2 || N_ant == 4) || N_id_cell < 0 || N_id_cell > 503) {
        lteb200_set_error(ctx, "chan_est: bad N_ant/N_id_cell");
        return LTEB200_ERR_ARGS;
    }
    if (N_rb_dl < 1 || N_rb_dl > 100) {      // rows of the shared-memory planes hold 1200 subcarriers
        lteb200_set_error(ctx, "chan_est: N_rb_dl = %d outside 1..100", N_rb_dl);
        return LTEB200_ERR_ARGS;
    }
    float2 *crs = nullptr;
    int rc = get_crs_table(ctx, N_id_cell, &crs);
    if (rc) return rc;
    CeParams P;
    P.N_rb = N_rb_dl; P.nsc = 12 * N_rb_dl; P.N_ant = N_ant; P.N_id_cell = N_id_cell;
    const size_t smem = sizeof(float) * 2 * 5 * 1200;
    chan_est_kernel<<<n_sf * N_ant, 256, smem, stream>>>((const float2 *)d_grid, d_subfr_num, P, crs, (float2 *)d_ce);
    ctx->launches += 1;
    LTE_CUDA_OK(ctx, cudaGetLastError());
    return LTEB200_OK;
}

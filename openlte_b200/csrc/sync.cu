// Cell search: coarse timing + fractional CFO, PSS + fine timing, SSS.  Host twins of
//   liblte_phy_dl_find_coarse_timing_and_freq_offset   /root/reference/liblte/src/liblte_phy.cc:5697-5852
//   liblte_phy_find_pss_and_fine_timing                liblte_phy.cc:5306-5510
//   liblte_phy_find_sss                                liblte_phy.cc:5578-5686
//
// The arithmetic-heavy parts run on the GPU: the cyclic-prefix autocorrelation at every offset of every slot
// (N_slots * samples_per_slot * CP complex MACs: 3.5e8 at 20 MHz / 160 slots, the reference's "hot loop #1"), the
// DFTs of the candidate symbols (84 + 80 + 1) and the 62-tap sequence correlations.  Each accumulation is done in the
// reference's own order (one thread per output, sequential over taps / slots), so the correlation values are
// bit-identical wherever only +,-,* are involved and every integer decision (peak positions, N_id_2, N_id_1,
// frame start) follows.  The O(samples_per_slot) peak picking and the first-match SSS rule are control flow and stay
// in host C++ below, restated from the cited lines.
#include "internal.h"
#include <math.h>

int fft_windows_launch(lteb200_ctx *ctx, const float *d_iq, long long n_samps, int fs_idx, int N_rb_dl, const long long *d_win_start, int n_win,
                       float *d_out, cudaStream_t stream);
int fs_legal(int fs_idx, int N_rb_dl);

namespace {

// corr[slot][i] = |sum_j x[idx] * conj(x[idx+N])|^2, idx = slot*S + i + j, j < cp   (liblte_phy.cc:5729-5743)
__global__ void __launch_bounds__(256)
cp_corr_kernel(const float *__restrict__ is, const float *__restrict__ qs, int S, int N, int cp, float *__restrict__ out /*[n_slots][S]*/)
{
    extern __shared__ float sh[];
    const int tile = 256 + cp;
    float *ai = sh, *aq = sh + tile, *bi = sh + 2 * tile, *bq = sh + 3 * tile;
    const int slot = blockIdx.y, i0 = blockIdx.x * 256;
    const size_t base = (size_t)slot * S + i0;
    for (int t = threadIdx.x; t < tile; t += 256) {
        ai[t] = is[base + t]; aq[t] = qs[base + t];
        bi[t] = is[base + N + t]; bq[t] = qs[base + N + t];
    }
    __syncthreads();
    const int i = i0 + threadIdx.x;
    if (i >= S) return;
    float cr = 0.f, ci = 0.f;
    for (int j = 0; j < cp; j++) {
        const int t = threadIdx.x + j;
        cr = __fadd_rn(cr, __fadd_rn(__fmul_rn(ai[t], bi[t]), __fmul_rn(aq[t], bq[t])));
        ci = __fadd_rn(ci, __fsub_rn(__fmul_rn(ai[t], bq[t]), __fmul_rn(aq[t], bi[t])));
    }
    out[(size_t)slot * S + i] = __fadd_rn(__fmul_rn(cr, cr), __fmul_rn(ci, ci));
}

// acc[i] = ((0 + c[0][i]) + c[1][i]) + ...   -- the reference's slot-by-slot accumulation order
__global__ void slot_sum_kernel(const float *__restrict__ c, int n_slots, int S, float *__restrict__ acc)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    float a = 0.f;
    for (int s = 0; s < n_slots; s++) a = __fadd_rn(a, c[(size_t)s * S + i]);
    acc[i] = a;
}

// per (slot, peak): atan2f of the CP correlation at the peak offset (liblte_phy.cc:5811-5824)
__global__ void cfo_kernel(const float *__restrict__ is, const float *__restrict__ qs, int S, int N, int cp, const int *__restrict__ peak, int n_peaks,
                           int n_slots, float *__restrict__ ang /*[n_slots][n_peaks]*/)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_slots * n_peaks) return;
    const int slot = t / n_peaks, k = t - slot * n_peaks;
    float cr = 0.f, ci = 0.f;
    for (int j = 0; j < cp; j++) {
        const size_t idx = (size_t)slot * S + peak[k] + j;
        cr = __fadd_rn(cr, __fadd_rn(__fmul_rn(is[idx], is[idx + N]), __fmul_rn(qs[idx], qs[idx + N])));
        ci = __fadd_rn(ci, __fsub_rn(__fmul_rn(is[idx], qs[idx + N]), __fmul_rn(qs[idx], is[idx + N])));
    }
    ang[t] = atan2f(ci, cr);
}

__global__ void interleave_kernel(const float *__restrict__ is, const float *__restrict__ qs, long long n, float2 *__restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_float2(is[i], qs[i]);
}

// |sum_j rx[first_sc[c] + j] * conj(seq[seq_id[c]][j])| over 62 taps, sequential (liblte_phy.cc:5385-5402, 5639-5650)
__global__ void corr62_kernel(const float2 *__restrict__ rows, int nsc, const int *__restrict__ row_of, const int *__restrict__ first_sc,
                              const int *__restrict__ seq_id, const float2 *__restrict__ seq /*[n_seq][62]*/, int n, float *__restrict__ out)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const float2 *r = rows + (size_t)row_of[c] * nsc + first_sc[c];
    const float2 *s = seq + (size_t)seq_id[c] * 62;
    float cr = 0.f, ci = 0.f;
    for (int j = 0; j < 62; j++) {
        cr = __fadd_rn(cr, __fadd_rn(__fmul_rn(r[j].x, s[j].x), __fmul_rn(r[j].y, s[j].y)));
        ci = __fadd_rn(ci, __fsub_rn(__fmul_rn(r[j].x, s[j].y), __fmul_rn(r[j].y, s[j].x)));
    }
    out[c] = __fsqrt_rn(__fadd_rn(__fmul_rn(cr, cr), __fmul_rn(ci, ci)));
}

struct FsP { int N, cp0, cpe, slot, frame; };
FsP fsp(int fs_idx) { FsP p; p.N = 128 << fs_idx; p.cp0 = 10 << fs_idx; p.cpe = 9 << fs_idx; p.slot = 960 << fs_idx; p.frame = p.slot * 20; return p; }

struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    template <typename T> T *as() { return (T *)p; }
};
static inline size_t at_least_one(size_t b) { return b ? b : 1; }
#define ALLOC(ctx, buf, bytes) LTE_CUDA_OK(ctx, cudaMalloc(&(buf).p, at_least_one((size_t)(bytes))))

}  // namespace

// (external linkage: the transmit side in control.cu maps the same sequences)
void gen_pss(int N_id_2, float2 *out)
{
    // generate_pss, liblte_phy.cc:8342-8368 (same expression, host libm)
    const float root = N_id_2 == 0 ? 25 : (N_id_2 == 1 ? 29 : 34);
    for (uint32_t i = 0; i < 31; i++) out[i] = make_float2(cosf(-M_PI * root * i * (i + 1) / 63), sinf(-M_PI * root * i * (i + 1) / 63));
    for (uint32_t i = 31; i < 62; i++) out[i] = make_float2(cosf(-M_PI * root * (i + 1) * (i + 2) / 63), sinf(-M_PI * root * (i + 1) * (i + 2) / 63));
}

void gen_sss(int N_id_1, int N_id_2, float2 *s0, float2 *s5)
{
    // generate_sss, liblte_phy.cc:8377-8474 (36.211 6.11.2.1)
    const int qp = N_id_1 / 30, q = (N_id_1 + (qp * (qp + 1) / 2)) / 30, mp = N_id_1 + (q * (q + 1) / 2);
    const int m0 = mp % 31, m1 = (m0 + (mp / 31) + 1) % 31;
    int xs[31] = {0}, xc[31] = {0}, xz[31] = {0};
    xs[4] = xc[4] = xz[4] = 1;
    for (int i = 0; i < 26; i++) {
        xs[i + 5] = (xs[i + 2] + xs[i]) % 2;
        xc[i + 5] = (xc[i + 3] + xc[i]) % 2;
        xz[i + 5] = (xz[i + 4] + xz[i + 2] + xz[i + 1] + xz[i]) % 2;
    }
    for (int i = 0; i < 31; i++) {
        const int a0 = 1 - 2 * xs[(i + m0) % 31], a1 = 1 - 2 * xs[(i + m1) % 31];
        const int c0 = 1 - 2 * xc[(i + N_id_2) % 31], c1 = 1 - 2 * xc[(i + N_id_2 + 3) % 31];
        const int z0 = 1 - 2 * xz[(i + (m0 % 8)) % 31], z1 = 1 - 2 * xz[(i + (m1 % 8)) % 31];
        s0[2 * i] = make_float2((float)(a0 * c0), 0.f); s0[2 * i + 1] = make_float2((float)(a1 * c1 * z0), 0.f);
        s5[2 * i] = make_float2((float)(a1 * c0), 0.f); s5[2 * i + 1] = make_float2((float)(a0 * c1 * z1), 0.f);
    }
}

namespace {

// uploads split I/Q [lo, hi) (plus `pad` zero samples behind it), optionally also as an interleaved float2 stream
int upload_iq(lteb200_ctx *ctx, const float *is, const float *qs, long long lo, long long hi, DevBuf &di, DevBuf &dq, DevBuf *diq, long long pad = 0)
{
    const long long n = hi - lo;
    ALLOC(ctx, di, sizeof(float) * (n + pad));
    ALLOC(ctx, dq, sizeof(float) * (n + pad));
    if (pad) {
        LTE_CUDA_OK(ctx, cudaMemset((float *)di.p + n, 0, sizeof(float) * pad));
        LTE_CUDA_OK(ctx, cudaMemset((float *)dq.p + n, 0, sizeof(float) * pad));
    }
    LTE_CUDA_OK(ctx, cudaMemcpy(di.p, is + lo, sizeof(float) * n, cudaMemcpyHostToDevice));
    LTE_CUDA_OK(ctx, cudaMemcpy(dq.p, qs + lo, sizeof(float) * n, cudaMemcpyHostToDevice));
    if (diq) {
        ALLOC(ctx, *diq, sizeof(float2) * n);
        interleave_kernel<<<(unsigned)((n + 255) / 256), 256>>>(di.as<float>(), dq.as<float>(), n, diq->as<float2>());
        ctx->launches += 1;
    }
    return LTEB200_OK;
}

// DFT + correlation of a list of candidates.  win[w]: absolute first sample of window w; cand: (row, first_sc, seq)
int run_corr(lteb200_ctx *ctx, const float *d_iq, long long n, int fs_idx, int N_rb_dl, const std::vector<long long> &win,
             const std::vector<int> &row_of, const std::vector<int> &first_sc, const std::vector<int> &seq_id, const std::vector<float2> &seq,
             std::vector<float> &out)
{
    const int nsc = 12 * N_rb_dl, nc = (int)row_of.size();
    DevBuf dwin, drows, dro, dfs, dsi, dseq, dout;
    ALLOC(ctx, dwin, sizeof(long long) * win.size());
    ALLOC(ctx, drows, sizeof(float2) * win.size() * nsc);
    ALLOC(ctx, dro, sizeof(int) * nc); ALLOC(ctx, dfs, sizeof(int) * nc); ALLOC(ctx, dsi, sizeof(int) * nc);
    ALLOC(ctx, dseq, sizeof(float2) * seq.size());
    ALLOC(ctx, dout, sizeof(float) * nc);
    LTE_CUDA_OK(ctx, cudaMemcpy(dwin.p, win.data(), sizeof(long long) * win.size(), cudaMemcpyHostToDevice));
    LTE_CUDA_OK(ctx, cudaMemcpy(dro.p, row_of.data(), sizeof(int) * nc, cudaMemcpyHostToDevice));
    LTE_CUDA_OK(ctx, cudaMemcpy(dfs.p, first_sc.data(), sizeof(int) * nc, cudaMemcpyHostToDevice));
    LTE_CUDA_OK(ctx, cudaMemcpy(dsi.p, seq_id.data(), sizeof(int) * nc, cudaMemcpyHostToDevice));
    LTE_CUDA_OK(ctx, cudaMemcpy(dseq.p, seq.data(), sizeof(float2) * seq.size(), cudaMemcpyHostToDevice));
    int rc = fft_windows_launch(ctx, d_iq, n, fs_idx, N_rb_dl, dwin.as<long long>(), (int)win.size(), drows.as<float>(), nullptr);
    if (rc) return rc;
    corr62_kernel<<<ceil_div(nc, 128), 128>>>(drows.as<float2>(), nsc, dro.as<int>(), dfs.as<int>(), dsi.as<int>(), dseq.as<float2>(), nc, dout.as<float>());
    ctx->launches += 1;
    out.resize(nc);
    LTE_CUDA_OK(ctx, cudaMemcpy(out.data(), dout.p, sizeof(float) * nc, cudaMemcpyDeviceToHost));
    return LTEB200_OK;
}

// first sample of the DFT window of samples_to_symbols_dl(slot_start_idx, symbol_offset = 0): index + CP_0 - 1 (liblte_phy.cc:8607-8622)
inline long long win_of(uint32_t slot_start_idx, const FsP &p) { return (long long)slot_start_idx + p.cp0 - 1; }

}  // namespace

extern "C" int lteb200_find_coarse_timing_and_freq_offset(lteb200_ctx *ctx, const float *i_samps, const float *q_samps, long long n_samps,
                                                          int fs_idx, uint32_t N_slots, float *freq_offset, uint32_t *symb_starts,
                                                          uint32_t *n_corr_peaks)
{
    if (!ctx || !i_samps || !q_samps || !freq_offset || !symb_starts || !n_corr_peaks || fs_idx < 0 || fs_idx > 4 || N_slots < 1) return LTEB200_ERR_ARGS;
    const FsP p = fsp(fs_idx);
    const long long need = (long long)N_slots * p.slot + p.slot + p.cpe + p.N;
    if (n_samps < need) { lteb200_set_error(ctx, "coarse timing reads %lld samples, buffer holds %lld", need, n_samps); return LTEB200_ERR_ARGS; }
    LTE_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    const int S = p.slot;
    DevBuf di, dq, dc, dacc;
    int rc = upload_iq(ctx, i_samps, q_samps, 0, need, di, dq, nullptr, 512);   // the last tile of a slot over-reads by < 256 samples
    if (rc) return rc;
    ALLOC(ctx, dc, sizeof(float) * (size_t)N_slots * S);
    ALLOC(ctx, dacc, sizeof(float) * S);
    dim3 grid(ceil_div(S, 256), N_slots);
    cp_corr_kernel<<<grid, 256, sizeof(float) * 4 * (256 + p.cpe)>>>(di.as<float>(), dq.as<float>(), S, p.N, p.cpe, dc.as<float>());
    slot_sum_kernel<<<ceil_div(S, 256), 256>>>(dc.as<float>(), (int)N_slots, S, dacc.as<float>());
    ctx->launches += 2;
    std::vector<float> ac(2 * 15360 * 2 + 16, 0.0f);
    LTE_CUDA_OK(ctx, cudaMemcpy(ac.data(), dacc.p, sizeof(float) * S, cudaMemcpyDeviceToHost));
    // ---- control flow of liblte_phy.cc:5746-5806 on the host
    const int per_else = p.N + p.cpe, blank = per_else / 10;
    float mean = 0;
    for (int i = 0; i < S; i++) { mean += ac[i]; ac[i + S] = ac[i]; }
    mean /= S;
    for (int i = 0; i < S; i++) if (ac[i] <= mean) { ac[i] = 0; ac[i + S] = 0; }
    const int lag = p.N + p.cp0 + per_else * 3;
    for (int i = 0; i < S; i++) ac[i] *= ac[lag + i];
    int peaks[5] = {0, 0, 0, 0, 0}, n = 5;
    for (int k = 0; k < 5; k++) {
        float mx = 0; peaks[k] = 0;
        for (int j = 0; j < S; j++) if (ac[j] > mx) { mx = ac[j]; peaks[k] = j; }
        if (mx == 0) { n = k; break; }
        int t = peaks[k];
        while (t > 0) t -= per_else;
        for (int j = 0; j < 7; j++) {
            t += per_else;
            for (int z = 0; z < blank; z++) {
                const uint32_t idx = (uint32_t)(t - (blank / 2) + z);   // unsigned in the reference: a negative index fails the bound test
                if (idx <= 15360u * 2) ac[idx] = 0;
            }
        }
    }
    *n_corr_peaks = (uint32_t)n;
    // ---- fractional frequency offset (liblte_phy.cc:5808-5830)
    if (n > 0) {
        DevBuf dpk, dang;
        ALLOC(ctx, dpk, sizeof(int) * 5);
        ALLOC(ctx, dang, sizeof(float) * N_slots * n);
        LTE_CUDA_OK(ctx, cudaMemcpy(dpk.p, peaks, sizeof(int) * 5, cudaMemcpyHostToDevice));
        cfo_kernel<<<ceil_div((int)N_slots * n, 128), 128>>>(di.as<float>(), dq.as<float>(), S, p.N, p.cpe, dpk.as<int>(), n, (int)N_slots, dang.as<float>());
        ctx->launches += 1;
        std::vector<float> ang((size_t)N_slots * n);
        LTE_CUDA_OK(ctx, cudaMemcpy(ang.data(), dang.p, sizeof(float) * ang.size(), cudaMemcpyDeviceToHost));
        float ferr[5] = {0, 0, 0, 0, 0};
        for (uint32_t s = 0; s < N_slots; s++)
            for (int k = 0; k < n; k++) ferr[k] += ang[(size_t)s * n + k] / (p.N * 2 * M_PI * (0.0005 / p.slot));
        for (int k = 0; k < n; k++) freq_offset[k] = ferr[k] / N_slots;
    }
    for (int k = 0; k < n; k++) {
        int t = peaks[k];
        while (t > 0) t -= per_else;
        for (int j = 0; j < 7; j++) symb_starts[k * 7 + j] = (uint32_t)(t + (j + 1) * per_else);
    }
    return LTEB200_OK;
}

extern "C" int lteb200_find_pss_and_fine_timing(lteb200_ctx *ctx, const float *i_samps, const float *q_samps, long long n_samps, int fs_idx,
                                                int N_rb_dl, uint32_t *symb_starts, uint32_t *N_id_2, uint32_t *pss_symb, float *pss_thresh,
                                                float *freq_offset)
{
    if (!ctx || !i_samps || !q_samps || !symb_starts || !N_id_2 || !pss_symb || !pss_thresh || !freq_offset) return LTEB200_ERR_ARGS;
    if (!fs_legal(fs_idx, N_rb_dl)) { lteb200_set_error(ctx, "illegal (fs, N_rb_dl)"); return LTEB200_ERR_ARGS; }
    const FsP p = fsp(fs_idx);
    LTE_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    const int nsc = 12 * N_rb_dl, c0 = nsc / 2 - 31;
    // the 84 + 80 windows lie within [0, max symb_start + 12 slots + N + CP + 40)
    uint32_t mx = 0;
    for (int j = 0; j < 7; j++) mx = symb_starts[j] > mx ? symb_starts[j] : mx;
    long long hi = (long long)mx + 12LL * p.slot + p.N + p.cp0 + 64;
    if (hi > n_samps) hi = n_samps;
    DevBuf di, dq, diq;
    int rc = upload_iq(ctx, i_samps, q_samps, 0, hi, di, dq, &diq);
    if (rc) return rc;
    std::vector<float2> seq(3 * 62);
    for (int k = 0; k < 3; k++) gen_pss(k, &seq[k * 62]);
    std::vector<long long> win;
    std::vector<int> row_of, first_sc, seq_id;
    for (int i = 0; i < 12; i++)
        for (int j = 0; j < 7; j++) {
            win.push_back(win_of(symb_starts[j] + (uint32_t)p.slot * i, p));
            for (int k = 0; k < 3; k++)
                for (int sh = -1; sh <= 1; sh++) { row_of.push_back(i * 7 + j); first_sc.push_back(c0 + sh); seq_id.push_back(k); }
        }
    std::vector<float> a;
    rc = run_corr(ctx, diq.as<float>(), hi, fs_idx, N_rb_dl, win, row_of, first_sc, seq_id, seq, a);
    if (rc) return rc;
    // selection in the reference's visiting order with strict '>' (liblte_phy.cc:5406-5427)
    float corr_max = 0; int which = 0;
    for (size_t c = 0; c < a.size(); c++)
        if (a[c] > corr_max) { corr_max = a[c]; which = first_sc[c] - c0; *pss_symb = (uint32_t)row_of[c]; *N_id_2 = (uint32_t)seq_id[c]; }
    *freq_offset = which * 15000.0f;
    // ---- fine timing: +-40 samples around the winner (liblte_phy.cc:5448-5491)
    const uint32_t N_s = *pss_symb / 7, N_symb = *pss_symb % 7;
    win.clear(); row_of.clear(); first_sc.clear(); seq_id.clear();
    for (int i = -40; i < 40; i++) {
        int idx = (int)(symb_starts[N_symb] + (uint32_t)p.slot * N_s);
        if (i < 0) { if (idx >= -i) idx += i; } else idx += i;
        win.push_back(win_of((uint32_t)idx, p));
        row_of.push_back(i + 40); first_sc.push_back(c0 + which); seq_id.push_back((int)*N_id_2);
    }
    rc = run_corr(ctx, diq.as<float>(), hi, fs_idx, N_rb_dl, win, row_of, first_sc, seq_id, seq, a);
    if (rc) return rc;
    corr_max = 0; int timing = 0;
    for (int i = 0; i < 80; i++) if (a[i] > corr_max) { corr_max = a[i]; timing = i - 40; }
    *pss_thresh = corr_max;
    uint32_t t = symb_starts[N_symb] + (uint32_t)p.slot * N_s + (uint32_t)timing;
    while (t + p.N + p.cpe < (uint32_t)p.slot) t += p.frame;
    symb_starts[0] = t + (p.N + p.cpe) * 1 - p.slot;
    for (int i = 1; i < 7; i++) symb_starts[i] = t + (p.N + p.cpe) * i + p.N + p.cp0 - p.slot;
    return LTEB200_OK;
}

extern "C" int lteb200_find_sss(lteb200_ctx *ctx, const float *i_samps, const float *q_samps, long long n_samps, int fs_idx, int N_rb_dl,
                                uint32_t N_id_2, uint32_t *symb_starts, float pss_thresh, uint32_t *N_id_1, uint32_t *frame_start_idx)
{
    if (!ctx || !i_samps || !q_samps || !symb_starts || !N_id_1 || !frame_start_idx || N_id_2 > 2) return LTEB200_ERR_ARGS;
    if (!fs_legal(fs_idx, N_rb_dl)) { lteb200_set_error(ctx, "illegal (fs, N_rb_dl)"); return LTEB200_ERR_ARGS; }
    const FsP p = fsp(fs_idx);
    LTE_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    const int nsc = 12 * N_rb_dl, c0 = nsc / 2 - 31;
    const long long w0 = win_of(symb_starts[5], p);
    if (w0 < 0 || w0 + p.N > n_samps) { lteb200_set_error(ctx, "SSS symbol lies outside the sample buffer"); return LTEB200_ERR_ARGS; }
    DevBuf di, dq, diq;
    int rc = upload_iq(ctx, i_samps, q_samps, w0, w0 + p.N, di, dq, &diq);
    if (rc) return rc;
    std::vector<float2> seq(336 * 62);
    std::vector<int> row_of(336, 0), first_sc(336, c0), seq_id(336);
    for (int i = 0; i < 168; i++) { gen_sss(i, (int)N_id_2, &seq[(2 * i) * 62], &seq[(2 * i + 1) * 62]); seq_id[2 * i] = 2 * i; seq_id[2 * i + 1] = 2 * i + 1; }
    std::vector<long long> win(1, 0);
    std::vector<float> a;
    rc = run_corr(ctx, diq.as<float>(), p.N, fs_idx, N_rb_dl, win, row_of, first_sc, seq_id, seq, a);
    if (rc) return rc;
    // first N_id_1 whose subframe-0 or subframe-5 sequence exceeds 0.9 * pss_thresh wins -- not the arg-max (liblte_phy.cc:5626-5683)
    const float thresh = pss_thresh * 0.9;
    const uint32_t back = (p.N + p.cpe) * 4 + p.N + p.cp0;
    for (int i = 0; i < 168; i++)
        for (int v = 0; v < 2; v++)
            if (a[2 * i + v] > thresh) {
                const uint32_t need = back + (v ? (uint32_t)p.slot * 10 : 0);
                while (symb_starts[5] < need) symb_starts[5] += p.frame;
                *N_id_1 = (uint32_t)i;
                *frame_start_idx = symb_starts[5] - need;
                return LTEB200_OK;
            }
    return LTEB200_ERR_DECODE;   // the reference returns its entry default (INVALID_INPUTS) when nothing matches; callers only test != SUCCESS
}

// C-ABI entry points: OFDM demod + channel estimate, whole-chain receive, host twins.
#include "internal.h"
#include <algorithm>

int ofdm_demod_launch(lteb200_ctx *ctx, const void *d_iq, int iq_fmt, long long n_samps, int fs_idx, int N_rb_dl,
                      const long long *d_start_idx, int n_sf, float *d_grid, cudaStream_t stream);
int chan_est_launch(lteb200_ctx *ctx, const float *d_grid, const int *d_subfr_num, int N_rb_dl, int N_id_cell, int N_ant, int n_sf,
                    float *d_ce, cudaStream_t stream);
int soa_to_grid(lteb200_ctx *ctx, const float *d_re, const float *d_im, int rows, int nsc, int src_stride, float *d_out, cudaStream_t stream);

namespace {
__global__ void split_grid_kernel(const float2 *__restrict__ in, int rows, int nsc, int dst_stride, float *__restrict__ re, float *__restrict__ im)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * nsc) return;
    const int r = i / nsc, k = i - r * nsc;
    const float2 v = in[i];
    re[r * dst_stride + k] = v.x;
    im[r * dst_stride + k] = v.y;
}
__global__ void interleave_iq_kernel(const float *__restrict__ i_s, const float *__restrict__ q_s, long long n, float2 *__restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_float2(i_s[i], q_s[i]);
}
}  // namespace

extern "C" int lteb200_get_dl_subframe_and_ce_batch(lteb200_ctx *ctx, const void *d_iq, int iq_fmt, long long n_samps, int fs_idx,
                                                    int N_rb_dl, int N_id_cell, int N_ant, const long long *start_idx, const int *subfr_num,
                                                    int n_sf, float *d_grid, float *d_ce, void *stream_)
{
    if (!ctx || !d_iq || !start_idx || !subfr_num || !d_grid || !d_ce || n_sf < 1 || n_samps < 1) return LTEB200_ERR_ARGS;
    cudaStream_t stream = (cudaStream_t)stream_;
    StreamScope scope;               // stream-ordered temporaries are released on every exit path
    LTE_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    for (int i = 0; i < n_sf; i++) if (subfr_num[i] < 0 || subfr_num[i] > 9) { lteb200_set_error(ctx, "subframe number out of range"); return LTEB200_ERR_ARGS; }
    long long *d_start = nullptr;
    int *d_sn = nullptr;
    LTE_CUDA_OK(ctx, scope.alloc(&d_start, sizeof(long long) * n_sf, stream));
    LTE_CUDA_OK(ctx, scope.alloc(&d_sn, sizeof(int) * n_sf, stream));
    if (int rc_ = lteb200_upload_small(ctx, start_idx, sizeof(long long) * n_sf, d_start, stream)) return rc_;
    if (int rc_ = lteb200_upload_small(ctx, subfr_num, sizeof(int) * n_sf, d_sn, stream)) return rc_;
    int rc = ofdm_demod_launch(ctx, d_iq, iq_fmt, n_samps, fs_idx, N_rb_dl, d_start, n_sf, d_grid, stream);
    if (rc == LTEB200_OK) rc = chan_est_launch(ctx, d_grid, d_sn, N_rb_dl, N_id_cell, N_ant, n_sf, d_ce, stream);
    return rc;
}

extern "C" int lteb200_ofdm_demod_batch(lteb200_ctx *ctx, const void *d_iq, int iq_fmt, long long n_samps, int fs_idx, int N_rb_dl,
                                        const long long *start_idx, int n_sf, float *d_grid, void *stream_)
{
    if (!ctx || !d_iq || !start_idx || !d_grid || n_sf < 1) return LTEB200_ERR_ARGS;
    cudaStream_t stream = (cudaStream_t)stream_;
    StreamScope scope;               // stream-ordered temporaries are released on every exit path
    LTE_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    long long *d_start = nullptr;
    LTE_CUDA_OK(ctx, scope.alloc(&d_start, sizeof(long long) * n_sf, stream));
    if (int rc_ = lteb200_upload_small(ctx, start_idx, sizeof(long long) * n_sf, d_start, stream)) return rc_;
    int rc = ofdm_demod_launch(ctx, d_iq, iq_fmt, n_samps, fs_idx, N_rb_dl, d_start, n_sf, d_grid, stream);
    return rc;
}

extern "C" int lteb200_chan_est_batch(lteb200_ctx *ctx, const float *d_grid, const int *subfr_num, int N_rb_dl, int N_id_cell, int N_ant, int n_sf,
                                      float *d_ce, void *stream_)
{
    if (!ctx || !d_grid || !subfr_num || !d_ce || n_sf < 1 || N_rb_dl < 6 || N_rb_dl > 100) return LTEB200_ERR_ARGS;
    cudaStream_t stream = (cudaStream_t)stream_;
    StreamScope scope;               // stream-ordered temporaries are released on every exit path
    LTE_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    for (int i = 0; i < n_sf; i++) if (subfr_num[i] < 0 || subfr_num[i] > 9) return LTEB200_ERR_ARGS;
    int *d_sn = nullptr;
    LTE_CUDA_OK(ctx, scope.alloc(&d_sn, sizeof(int) * n_sf, stream));
    if (int rc_ = lteb200_upload_small(ctx, subfr_num, sizeof(int) * n_sf, d_sn, stream)) return rc_;
    int rc = chan_est_launch(ctx, d_grid, d_sn, N_rb_dl, N_id_cell, N_ant, n_sf, d_ce, stream);
    return rc;
}

// Host twin of liblte_phy_get_dl_subframe_and_ce (liblte_phy.cc:5905): split float I/Q in, LIBLTE_PHY_SUBFRAME_STRUCT-shaped arrays out.
extern "C" int lteb200_get_dl_subframe_and_ce(lteb200_ctx *ctx, const float *i_samps, const float *q_samps, long long n_samps, int fs_idx,
                                              int N_rb_dl, uint32_t frame_start_idx, uint8_t subfr_num, uint32_t N_id_cell, uint8_t N_ant,
                                              float *rx_symb_re, float *rx_symb_im, float *rx_ce_re, float *rx_ce_im)
{
    if (!ctx || !i_samps || !q_samps || !rx_symb_re || !rx_symb_im || !rx_ce_re || !rx_ce_im) return LTEB200_ERR_ARGS;
    if (!(N_ant == 1 || N_ant == 2 || N_ant == 4) || fs_idx < 0 || fs_idx > 4) return LTEB200_ERR_ARGS;
    LTE_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    const int n_sf_samps = 1920 << fs_idx;
    // uint32 arithmetic as in the reference (liblte_phy.cc:5927): a frame that starts just before the buffer wraps back in
    const long long start = (long long)(uint32_t)(frame_start_idx + (uint32_t)subfr_num * (uint32_t)n_sf_samps);
    // only the window [start, start + subframe + 2 symbols + margin) travels to the device
    const long long lo = start > 0 ? start - 1 : 0;
    // last sample the reference reads: symbol 1 of the next subframe ends at start + subframe + (N + cp0) + (N + cpe) - 1 (liblte_phy.cc:5946-5957, 8621)
    long long hi = start + n_sf_samps + 2LL * (128 << fs_idx) + (10 << fs_idx) + (9 << fs_idx);
    if (hi > n_samps) hi = n_samps;
    if (lo >= hi) { lteb200_set_error(ctx, "subframe lies outside the sample buffer"); return LTEB200_ERR_ARGS; }
    const long long n = hi - lo;
    const int nsc = 12 * N_rb_dl;
    float *d_i = nullptr, *d_q = nullptr, *d_iq = nullptr, *d_grid = nullptr, *d_ce = nullptr, *d_soa = nullptr;
    LTE_CUDA_OK(ctx, cudaMalloc(&d_i, sizeof(float) * n));
    LTE_CUDA_OK(ctx, cudaMalloc(&d_q, sizeof(float) * n));
    LTE_CUDA_OK(ctx, cudaMalloc(&d_iq, sizeof(float) * 2 * n));
    LTE_CUDA_OK(ctx, cudaMalloc(&d_grid, sizeof(float) * 2 * 16 * nsc));
    LTE_CUDA_OK(ctx, cudaMalloc(&d_ce, sizeof(float) * 2 * N_ant * 14 * nsc));
    LTE_CUDA_OK(ctx, cudaMalloc(&d_soa, sizeof(float) * 2 * 4 * 16 * 1200));
    LTE_CUDA_OK(ctx, cudaMemcpy(d_i, i_samps + lo, sizeof(float) * n, cudaMemcpyHostToDevice));
    LTE_CUDA_OK(ctx, cudaMemcpy(d_q, q_samps + lo, sizeof(float) * n, cudaMemcpyHostToDevice));
    interleave_iq_kernel<<<(unsigned)((n + 255) / 256), 256>>>(d_i, d_q, n, (float2 *)d_iq);
    ctx->launches += 1;
    const long long st = start - lo;
    const int sn = subfr_num;
    // the uploaded window starts at sample `lo` of the caller's buffer: a fused frequency shift (lteb200_set_freq_offset) must number the samples
    // as freq_shift() does on the whole buffer (fs.cc:696-713)
    ctx->cfo_base += lo;
    int rc = lteb200_get_dl_subframe_and_ce_batch(ctx, d_iq, LTEB200_IQ_F32, n, fs_idx, N_rb_dl, (int)N_id_cell, N_ant, &st, &sn, 1, d_grid, d_ce, nullptr);
    ctx->cfo_base -= lo;
    if (rc == LTEB200_OK) {
        float *s_re = d_soa, *s_im = d_soa + 4 * 16 * 1200;
        cudaMemset(d_soa, 0, sizeof(float) * 2 * 4 * 16 * 1200);
        split_grid_kernel<<<ceil_div(16 * nsc, 256), 256>>>((const float2 *)d_grid, 16, nsc, 1200, s_re, s_im);
        cudaMemcpy(rx_symb_re, s_re, sizeof(float) * 16 * 1200, cudaMemcpyDeviceToHost);
        cudaMemcpy(rx_symb_im, s_im, sizeof(float) * 16 * 1200, cudaMemcpyDeviceToHost);
        for (int p = 0; p < N_ant; p++)
            split_grid_kernel<<<ceil_div(14 * nsc, 256), 256>>>((const float2 *)d_ce + (size_t)p * 14 * nsc, 14, nsc, 1200, s_re + p * 16 * 1200, s_im + p * 16 * 1200);
        ctx->launches += 1 + N_ant;
        // only rows 0..13 of rx_ce are defined by the reference; rows 14-15 of the caller's arrays are left untouched
        for (int p = 0; p < N_ant; p++) {
            cudaMemcpy(rx_ce_re + p * 16 * 1200, s_re + p * 16 * 1200, sizeof(float) * 14 * 1200, cudaMemcpyDeviceToHost);
            cudaMemcpy(rx_ce_im + p * 16 * 1200, s_im + p * 16 * 1200, sizeof(float) * 14 * 1200, cudaMemcpyDeviceToHost);
        }
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { lteb200_set_error(ctx, "get_dl_subframe_and_ce: %s", cudaGetErrorString(e)); rc = LTEB200_ERR_CUDA; }
    }
    cudaFree(d_i); cudaFree(d_q); cudaFree(d_iq); cudaFree(d_grid); cudaFree(d_ce); cudaFree(d_soa);
    return rc;
}

// Host twin of liblte_phy_pdsch_channel_decode (liblte_phy.cc:3690) for one allocation.
extern "C" int lteb200_pdsch_channel_decode(lteb200_ctx *ctx, const float *rx_symb_re, const float *rx_symb_im, const float *rx_ce_re,
                                            const float *rx_ce_im, int N_rb_dl, uint32_t subfr_num, const lteb200_alloc *alloc,
                                            uint32_t N_pdcch_symbs, uint32_t N_id_cell, uint8_t N_ant, int turbo_mode, int n_iter,
                                            uint8_t *out_bits, uint32_t *N_out_bits)
{
    // entry default of the reference: INVALID_INPUTS (liblte_phy.cc:3699-3705)
    if (!ctx || !rx_symb_re || !rx_symb_im || !rx_ce_re || !rx_ce_im || !alloc || !out_bits || !N_out_bits || N_id_cell > 503) return LTEB200_ERR_ARGS;
    TbBytesScope bytes_out(ctx);     // out_bits is one bit per byte like the reference's, whatever the context's transport-block format
    LTE_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    const int nsc = 12 * N_rb_dl;
    float *d_in = nullptr, *d_grid = nullptr, *d_ce = nullptr;
    uint8_t *d_tb = nullptr;
    int32_t *d_ok = nullptr;
    const size_t plane = 16 * 1200;
    LTE_CUDA_OK(ctx, cudaMalloc(&d_in, sizeof(float) * plane * (2 + 2 * N_ant)));
    LTE_CUDA_OK(ctx, cudaMalloc(&d_grid, sizeof(float) * 2 * 16 * nsc));
    LTE_CUDA_OK(ctx, cudaMalloc(&d_ce, sizeof(float) * 2 * N_ant * 14 * nsc));
    LTE_CUDA_OK(ctx, cudaMalloc(&d_tb, alloc->tbs > 0 ? alloc->tbs : 1));
    LTE_CUDA_OK(ctx, cudaMalloc(&d_ok, sizeof(int32_t)));
    LTE_CUDA_OK(ctx, cudaMemcpy(d_in, rx_symb_re, sizeof(float) * plane, cudaMemcpyHostToDevice));
    LTE_CUDA_OK(ctx, cudaMemcpy(d_in + plane, rx_symb_im, sizeof(float) * plane, cudaMemcpyHostToDevice));
    for (int p = 0; p < N_ant; p++) {
        LTE_CUDA_OK(ctx, cudaMemcpy(d_in + plane * (2 + 2 * p), rx_ce_re + p * plane, sizeof(float) * plane, cudaMemcpyHostToDevice));
        LTE_CUDA_OK(ctx, cudaMemcpy(d_in + plane * (3 + 2 * p), rx_ce_im + p * plane, sizeof(float) * plane, cudaMemcpyHostToDevice));
    }
    int rc = soa_to_grid(ctx, d_in, d_in + plane, 16, nsc, 1200, d_grid, nullptr);
    for (int p = 0; p < N_ant && rc == LTEB200_OK; p++)
        rc = soa_to_grid(ctx, d_in + plane * (2 + 2 * p), d_in + plane * (3 + 2 * p), 14, nsc, 1200, d_ce + (size_t)p * 2 * 14 * nsc, nullptr);
    const int sn = (int)subfr_num;
    if (rc == LTEB200_OK)
        rc = lteb200_pdsch_decode_batch(ctx, d_grid, d_ce, N_rb_dl, (int)N_id_cell, N_ant, &sn, 1, (int)N_pdcch_symbs, alloc, 1, turbo_mode, n_iter,
                                        d_tb, alloc->tbs, d_ok, nullptr);
    int32_t ok = 0;
    if (rc == LTEB200_OK) {
        cudaError_t e = cudaMemcpy(&ok, d_ok, sizeof(ok), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { lteb200_set_error(ctx, "pdsch_channel_decode: %s", cudaGetErrorString(e)); rc = LTEB200_ERR_CUDA; }
        else if (!ok) rc = LTEB200_ERR_DECODE;      // demodulated but CRC bad: DECODE_FAIL, outputs untouched (liblte_phy.cc:3720)
        else {
            cudaMemcpy(out_bits, d_tb, alloc->tbs, cudaMemcpyDeviceToHost);
            *N_out_bits = (uint32_t)alloc->tbs;
        }
    }
    cudaFree(d_in); cudaFree(d_grid); cudaFree(d_ce); cudaFree(d_tb); cudaFree(d_ok);
    return rc;
}

extern "C" int lteb200_dl_receive_batch(lteb200_ctx *ctx, const void *d_iq, int iq_fmt, long long n_samps, int fs_idx, int N_rb_dl,
                                        int N_id_cell, int N_ant, const long long *start_idx, const int *subfr_num, int n_sf,
                                        int N_pdcch_symbs, const lteb200_alloc *allocs, int n_alloc, int turbo_mode, int n_iter,
                                        uint8_t *d_tb_bits, int tb_stride, int32_t *d_crc_ok, void *stream_)
{
    if (!ctx || n_sf < 1) return LTEB200_ERR_ARGS;
    cudaStream_t stream = (cudaStream_t)stream_;
    StreamScope scope;               // stream-ordered temporaries are released on every exit path
    LTE_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    const int nsc = 12 * N_rb_dl;
    float *d_grid = nullptr, *d_ce = nullptr;
    LTE_CUDA_OK(ctx, scope.alloc(&d_grid, sizeof(float) * 2 * 16 * nsc * (size_t)n_sf, stream));
    LTE_CUDA_OK(ctx, scope.alloc(&d_ce, sizeof(float) * 2 * 14 * nsc * (size_t)n_sf * N_ant, stream));
    int rc = lteb200_get_dl_subframe_and_ce_batch(ctx, d_iq, iq_fmt, n_samps, fs_idx, N_rb_dl, N_id_cell, N_ant, start_idx, subfr_num, n_sf, d_grid,
                                                  d_ce, stream);
    if (rc == LTEB200_OK)
        rc = lteb200_pdsch_decode_batch(ctx, d_grid, d_ce, N_rb_dl, N_id_cell, N_ant, subfr_num, n_sf, N_pdcch_symbs, allocs, n_alloc, turbo_mode,
                                        n_iter, d_tb_bits, tb_stride, d_crc_ok, stream);
    return rc;
}

// ---- host pipeline plan (pure host logic, also exported for CPU tests as lteb200_e2e_plan) -------------------------------------------
// Chunk schedule.  Small chunks cost efficiency (partial waves, launch gaps), large ones expose the first copy in and the last
// copy out.  LTEB200_E2E_SCHEDULE="a,b,c" forces explicit sizes, LTEB200_E2E_CHUNK=n equal chunks of n subframes (development aids).
// Returns bounds: chunk c = subframes [bounds[c], bounds[c+1]).
static std::vector<int> plan_chunks(int n_sf, int n_alloc, int K_first, int turbo_mode, int n_sm)
{
    std::vector<int> bounds;
    bounds.push_back(0);
    if (const char *sched = getenv("LTEB200_E2E_SCHEDULE")) {         // explicit chunk sizes "a,b,c,..." (development aid); the rest goes into a last chunk
        int pos = 0;
        for (const char *q = sched; *q && pos < n_sf;) {
            const int sz = atoi(q);
            if (sz <= 0) break;
            pos += sz;
            if (pos < n_sf) bounds.push_back(pos);
            while (*q && *q != ',') q++;
            if (*q == ',') q++;
        }
    } else if (const char *e = getenv("LTEB200_E2E_CHUNK")) {
        int chunk = atoi(e);
        if (chunk < 64) chunk = 64;
        for (int b = chunk; b < n_sf; b += chunk) bounds.push_back(b);
    } else if (n_sf >= 256) {
        // With the round-2 kernels a 20 MHz subframe costs about as much on the GPU (1.3 us) as its 61 KB cost on PCIe (1.15 us), so the
        // pipeline is balanced: a chunk can only start once ALL its samples have arrived, which makes the first chunk's copy and every
        // growth step exposed time.  Chunks therefore start small and grow only as fast as the compute backlog allows
        // (size <= first + 0.12 x subframes already queued keeps the decode stream from ever waiting for a copy), and every size is a
        // whole number of decoder waves (148 SMs x resident blocks / transport blocks per subframe) so that no chunk ends in a
        // partial wave.  Measured on 4000 subframes (profiles/r2_e2e_schedule.txt): 10.8 ms in one chunk, 5.9-6.2 ms with equal or doubling chunks,
        // 5.7 ms with this schedule.
        const int nt_class = mlm_size_class(K_first);
        const int blocks_per_sm = nt_class == 384 ? 2 : (nt_class == 192 ? 4 : (nt_class == 128 ? 6 : 12));
        int wave = (n_sm * blocks_per_sm) / n_alloc;
        if (wave < 16) wave = 16;
        // a chunk costs about 1.28 us per subframe + 0.1 ms (kernel tails, launch gaps), the copy of the next one 1.2 us per subframe: the decode
        // stream never waits if next <= 1.07 x current + 83; the last chunk is kept small (its transport blocks are copied out after everything else)
        if (turbo_mode == LTEB200_TURBO_REF_EXACT) {
            // the reference-exact decoder runs one THREAD per trellis: its serial kernels take about 0.2 ms whatever the batch up to a few thousand
            // blocks, so small chunks cost it dearly; a few large ones instead (6, 14, 20, 20 ... waves; measured on 4000 subframes: 9.3 ms with the
            // schedule below it, 6.9 ms with this one)
            int pos = 0, sz = 6 * wave;
            while (n_sf - pos > sz + 4 * wave) {
                pos += sz;
                bounds.push_back(pos);
                sz = std::min(20 * wave, (int)(2.34 * sz) / wave * wave);
            }
        } else {
            const int first = 2 * wave, cap = 12 * wave;
            int pos = 0, sz = first;
            while (n_sf - pos > sz + 3 * wave) {
                pos += sz;
                bounds.push_back(pos);
                int nxt = (int)(1.07 * sz) + 83;
                nxt = (nxt / wave) * wave;
                if (nxt > cap) nxt = cap;
                sz = nxt;
            }
            if (n_sf - pos > 6 * wave) { pos = n_sf - 3 * wave; bounds.push_back(pos); }
        }
    }
    if (bounds.back() != n_sf) bounds.push_back(n_sf);
    return bounds;
}

// a subframe's 16 symbols end two symbols into the next subframe (the estimator's look-ahead): that halo travels with the chunk
static inline long long e2e_halo(int fs_idx) { return 2LL * ((128LL << fs_idx) + (10LL << fs_idx)) + 16; }

// One sample past the last sample chunk c's kernels may read = where the cumulative H2D copy must have got to before they start:
// everything the chunk's LAST subframe reads (its 14 symbols + the two look-ahead symbols), and at least up to the halo of the next chunk's
// first subframe; with gapped, duplicate or overlapping (non-decreasing) start indices the former is the larger one.  The last chunk takes
// the rest of the stream.
static long long chunk_copy_end(int c, const std::vector<int> &bounds, const long long *h_start_idx, int fs_idx, long long n_samps)
{
    const int n_chunks = (int)bounds.size() - 1;
    const long long halo = e2e_halo(fs_idx);
    long long hi = n_samps;
    if (c + 1 < n_chunks) {
        hi = h_start_idx[bounds[c + 1]] + halo;
        const long long last_end = h_start_idx[bounds[c + 1] - 1] + (1920LL << fs_idx) + halo;
        if (last_end > hi) hi = last_end;
    }
    if (hi > n_samps) hi = n_samps;
    return hi;
}

// HOST, no device needed: the plan lteb200_dl_receive_host follows for a batch -- chunk boundaries and, per chunk, how far the H2D copy has to
// have got before the chunk's kernels may start (copy_end[c], cumulative, in samples).  n_sm: multiprocessors of the device (148 on B200).
extern "C" int lteb200_e2e_plan(int n_sf, int n_alloc, int K_first, int turbo_mode, int n_sm, int fs_idx, long long n_samps, const long long *h_start_idx,
                                int *bounds_out, long long *copy_end_out, int capacity, int *n_chunks_out)
{
    if (n_sf < 1 || n_alloc < 1 || K_first < 40 || K_first > 6144 || n_sm < 1 || fs_idx < 0 || fs_idx > 4 || n_samps < 1 || !h_start_idx || !bounds_out ||
        !copy_end_out || !n_chunks_out)
        return LTEB200_ERR_ARGS;
    for (int i = 1; i < n_sf; i++) if (h_start_idx[i] < h_start_idx[i - 1]) return LTEB200_ERR_ARGS;
    const std::vector<int> bounds = plan_chunks(n_sf, n_alloc, K_first, turbo_mode, n_sm);
    const int n_chunks = (int)bounds.size() - 1;
    *n_chunks_out = n_chunks;
    if (n_chunks + 1 > capacity) return LTEB200_ERR_ARGS;
    long long copied = 0;
    for (int c = 0; c < n_chunks; c++) {
        bounds_out[c] = bounds[c];
        const long long hi = chunk_copy_end(c, bounds, h_start_idx, fs_idx, n_samps);
        if (hi > copied) copied = hi;
        copy_end_out[c] = copied;
    }
    bounds_out[n_chunks] = bounds[n_chunks];
    return LTEB200_OK;
}

// Host-pointer entry point.  The batch is cut into chunks of subframes that flow through a three-stage pipeline on
// three streams -- H2D copy of a chunk's samples, the receive kernels, D2H copy of its transport blocks -- so that PCIe
// traffic in both directions overlaps the kernels (pinned host buffers make the copies truly asynchronous; pageable
// ones still work, staged by the driver).  start_idx must be non-decreasing.
extern "C" int lteb200_dl_receive_host(lteb200_ctx *ctx, const void *h_iq, int iq_fmt, long long n_samps, int fs_idx, int N_rb_dl, int N_id_cell,
                                       int N_ant, const long long *h_start_idx, const int *h_subfr_num, int n_sf, int N_pdcch_symbs,
                                       const lteb200_alloc *allocs, int n_alloc, int turbo_mode, int n_iter, uint8_t *h_tb_bits, int tb_stride,
                                       int32_t *h_crc_ok)
{
    if (!ctx || !h_iq || !h_start_idx || !h_subfr_num || !allocs || !h_tb_bits || !h_crc_ok || n_sf < 1 || n_alloc < 1 || n_samps < 1 || fs_idx < 0 || fs_idx > 4)
        return LTEB200_ERR_ARGS;
    for (int i = 1; i < n_sf; i++)
        if (h_start_idx[i] < h_start_idx[i - 1]) { lteb200_set_error(ctx, "dl_receive_host: start_idx must be non-decreasing"); return LTEB200_ERR_ARGS; }
    LTE_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    const size_t samp_bytes = iq_fmt == LTEB200_IQ_I8 ? 2 : 8;
    if (!ctx->s_h2d) {
        LTE_CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
        // the decode kernels (critical path) outrank the front-end kernels of the NEXT chunk, which only fill the gaps they leave
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        LTE_CUDA_OK(ctx, cudaStreamCreateWithPriority(&ctx->s_comp, cudaStreamNonBlocking, prio_hi));
        LTE_CUDA_OK(ctx, cudaStreamCreateWithPriority(&ctx->s_front, cudaStreamNonBlocking, prio_lo));
        LTE_CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
    }
    int K_first = 6144;
    { AllocTables t0; if (get_alloc_tables(ctx, N_rb_dl, N_id_cell, N_ant, N_pdcch_symbs, allocs[0], &t0) == LTEB200_OK) K_first = t0.K; }
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, ctx->device);
    const std::vector<int> bounds = plan_chunks(n_sf, n_alloc, K_first, turbo_mode, n_sm);      // chunk c = subframes [bounds[c], bounds[c+1])
    const int n_chunks = (int)bounds.size() - 1;
    StreamScope scope;                 // every event and stream-ordered allocation below is released on every exit path
    std::vector<cudaEvent_t> ev_h(n_chunks), ev_c(n_chunks), ev_f(n_chunks);
    for (int c = 0; c < n_chunks; c++) {
        LTE_CUDA_OK(ctx, scope.event(&ev_h[c]));
        LTE_CUDA_OK(ctx, scope.event(&ev_c[c]));
        LTE_CUDA_OK(ctx, scope.event(&ev_f[c]));
    }
    cudaEvent_t ev_alloc;
    LTE_CUDA_OK(ctx, scope.event(&ev_alloc));
    // LTEB200_E2E_TRACE=1: device-side time line of the three streams, printed to stderr (development aid)
    const bool trace = getenv("LTEB200_E2E_TRACE") != nullptr;
    cudaEvent_t tr[6] = {};
    if (trace) for (auto &e : tr) LTE_CUDA_OK(ctx, scope.event(&e, cudaEventDefault));
    std::vector<cudaEvent_t> tr_c0, tr_c1;
    void *d_iq = nullptr;
    uint8_t *d_tb = nullptr;
    int32_t *d_ok = nullptr;
    const size_t tb_row = (size_t)n_alloc * tb_stride;
    LTE_CUDA_OK(ctx, scope.alloc(&d_iq, samp_bytes * n_samps, ctx->s_comp));
    LTE_CUDA_OK(ctx, scope.alloc(&d_tb, (size_t)n_sf * tb_row, ctx->s_comp));
    LTE_CUDA_OK(ctx, scope.alloc(&d_ok, sizeof(int32_t) * (size_t)n_sf * n_alloc, ctx->s_comp));
    if (trace) cudaEventRecord(tr[0], ctx->s_comp);
    LTE_CUDA_OK(ctx, cudaEventRecord(ev_alloc, ctx->s_comp));
    LTE_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->s_h2d, ev_alloc, 0));
    LTE_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->s_d2h, ev_alloc, 0));
    // stage 1: sample slices, back to back.  Chunk c's copy ends `halo` samples into chunk c+1; the next copy continues from there.
    long long copied = 0;
    for (int c = 0; c < n_chunks; c++) {
        const long long hi = chunk_copy_end(c, bounds, h_start_idx, fs_idx, n_samps);
        if (hi > copied) {
            LTE_CUDA_OK(ctx, cudaMemcpyAsync((char *)d_iq + samp_bytes * copied, (const char *)h_iq + samp_bytes * copied, samp_bytes * (size_t)(hi - copied),
                                            cudaMemcpyHostToDevice, ctx->s_h2d));
            copied = hi;
        }
        LTE_CUDA_OK(ctx, cudaEventRecord(ev_h[c], ctx->s_h2d));
        if (trace && c == 0) cudaEventRecord(tr[1], ctx->s_h2d);
    }
    if (trace) cudaEventRecord(tr[2], ctx->s_h2d);
    // stage 2 is split over two streams: OFDM + channel estimate of chunk c+1 (s_front) may run while the decode kernel of chunk c
    // (s_comp) drains its last wave, so the tail of one chunk is filled with the head of the next.
    int rc = LTEB200_OK;
    const int nsc = 12 * N_rb_dl;
    LTE_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->s_front, ev_alloc, 0));
    // start indices and subframe numbers of the whole batch go to the device once; the chunks use offsets into them
    for (int i = 0; i < n_sf; i++) if (h_subfr_num[i] < 0 || h_subfr_num[i] > 9) { lteb200_set_error(ctx, "subframe number out of range"); rc = LTEB200_ERR_ARGS; }
    long long *d_start_all = nullptr;
    int *d_sn_all = nullptr;
    if (rc == LTEB200_OK) {
        // (allocated in the order of the stream that fills them; the decode stream only touches them behind an event of that stream, and the
        // scope releases them after the final synchronisation)
        LTE_CUDA_OK(ctx, scope.alloc(&d_start_all, sizeof(long long) * n_sf, ctx->s_front));
        LTE_CUDA_OK(ctx, scope.alloc(&d_sn_all, sizeof(int) * n_sf, ctx->s_front));
        rc = lteb200_upload_small(ctx, h_start_idx, sizeof(long long) * n_sf, d_start_all, ctx->s_front);
        if (rc == LTEB200_OK) rc = lteb200_upload_small(ctx, h_subfr_num, sizeof(int) * n_sf, d_sn_all, ctx->s_front);
    }
    for (int c = 0; c < n_chunks && rc == LTEB200_OK; c++) {
        const int b0 = bounds[c], nb = bounds[c + 1] - b0;
        float *d_grid = nullptr, *d_ce = nullptr;
        LTE_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->s_front, ev_h[c], 0));
        LTE_CUDA_OK(ctx, scope.alloc(&d_grid, sizeof(float) * 2 * 16 * nsc * (size_t)nb, ctx->s_front));
        LTE_CUDA_OK(ctx, scope.alloc(&d_ce, sizeof(float) * 2 * 14 * nsc * (size_t)nb * N_ant, ctx->s_front));
        rc = ofdm_demod_launch(ctx, d_iq, iq_fmt, n_samps, fs_idx, N_rb_dl, d_start_all + b0, nb, d_grid, ctx->s_front);
        if (rc == LTEB200_OK) rc = chan_est_launch(ctx, d_grid, d_sn_all + b0, N_rb_dl, N_id_cell, N_ant, nb, d_ce, ctx->s_front);
        if (rc) break;
        LTE_CUDA_OK(ctx, cudaEventRecord(ev_f[c], ctx->s_front));
        LTE_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->s_comp, ev_f[c], 0));
        if (trace) { cudaEvent_t a, b; LTE_CUDA_OK(ctx, scope.event(&a, cudaEventDefault)); LTE_CUDA_OK(ctx, scope.event(&b, cudaEventDefault)); tr_c0.push_back(a); tr_c1.push_back(b); cudaEventRecord(a, ctx->s_comp); }
        rc = pdsch_decode_batch_dev(ctx, d_grid, d_ce, N_rb_dl, N_id_cell, N_ant, h_subfr_num + b0, d_sn_all + b0, nb, N_pdcch_symbs, allocs, n_alloc, turbo_mode,
                                    n_iter, d_tb + (size_t)b0 * tb_row, tb_stride, d_ok + (size_t)b0 * n_alloc, ctx->s_comp);
        if (rc) break;
        scope.release(d_grid, ctx->s_comp);       // stream-ordered after the decode kernels that read them
        scope.release(d_ce, ctx->s_comp);
        if (trace && c == 0) cudaEventRecord(tr[3], ctx->s_comp);
        if (trace) cudaEventRecord(tr_c1[c], ctx->s_comp);
        LTE_CUDA_OK(ctx, cudaEventRecord(ev_c[c], ctx->s_comp));
        LTE_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->s_d2h, ev_c[c], 0));
        LTE_CUDA_OK(ctx, cudaMemcpyAsync(h_tb_bits + (size_t)b0 * tb_row, d_tb + (size_t)b0 * tb_row, (size_t)nb * tb_row, cudaMemcpyDeviceToHost, ctx->s_d2h));
        LTE_CUDA_OK(ctx, cudaMemcpyAsync(h_crc_ok + (size_t)b0 * n_alloc, d_ok + (size_t)b0 * n_alloc, sizeof(int32_t) * (size_t)nb * n_alloc,
                                        cudaMemcpyDeviceToHost, ctx->s_d2h));
    }
    if (trace) { cudaEventRecord(tr[4], ctx->s_comp); cudaEventRecord(tr[5], ctx->s_d2h); }
    cudaError_t e1 = cudaStreamSynchronize(ctx->s_h2d), e2 = cudaStreamSynchronize(ctx->s_comp), e3 = cudaStreamSynchronize(ctx->s_d2h);
    { const cudaError_t e4 = cudaStreamSynchronize(ctx->s_front); if (e2 == cudaSuccess) e2 = e4; }
    if (trace) {
        float t[6] = {};
        for (int i = 1; i < 6; i++) cudaEventElapsedTime(&t[i], tr[0], tr[i]);
        fprintf(stderr, "e2e trace (ms from start): first H2D chunk done %.3f, all H2D done %.3f, first compute chunk done %.3f, all compute done %.3f, all D2H done %.3f (%d chunks, first %d subframes)\n",
                t[1], t[2], t[3], t[4], t[5], n_chunks, bounds[1]);
        for (size_t c = 0; c < tr_c0.size(); c++) {
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, tr[0], tr_c0[c]); cudaEventElapsedTime(&b, tr[0], tr_c1[c]);
            fprintf(stderr, "  chunk %zu: %d subframes, compute %.3f -> %.3f ms (%.3f ms, %.2f us/subframe)\n", c, bounds[c + 1] - bounds[c], a, b, b - a,
                    1e3 * (b - a) / (bounds[c + 1] - bounds[c]));
        }
    }
    if (rc == LTEB200_OK && (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)) {
        lteb200_set_error(ctx, "dl_receive_host: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3)));
        rc = LTEB200_ERR_CUDA;
    }
    return rc;
}

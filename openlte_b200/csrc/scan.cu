// The scan loop of LTE_fdd_dl_fs_samp_buf::work for a whole batch of subframes in one call
// (/root/reference/LTE_fdd_dl_file_scan/src/LTE_fdd_dl_fs_samp_buf.cc:491-515: get_dl_subframe_and_ce -> pdcch_channel_decode ->
// pdsch_channel_decode for every allocation the PDCCH announced):
//   1. OFDM demodulation + channel estimate of all subframes                         (ofdm.cu, chanest.cu)
//   2. PCFICH + PDCCH blind decode of all subframes                                   (control.cu), raw results to the host
//   3. DCI unpacking on the host (bit fields, a few hundred bytes per subframe)
//   4. subframes are grouped by identical allocation; each group's grid/estimate rows are gathered into a compact batch and go
//      through the batched PDSCH receive                                              (pdsch.cu, turbo_*.cu)
// Results are identical to running the reference's three calls subframe by subframe (tests/test_scan_gpu.py).
#include "internal.h"
#include <map>

namespace {
__global__ void gather_rows_kernel(const float4 *__restrict__ src, const int *__restrict__ rows, int row_f4, float4 *__restrict__ dst)
{
    const float4 *s = src + (size_t)rows[blockIdx.y] * row_f4;
    float4 *d = dst + (size_t)blockIdx.y * row_f4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < row_f4; i += gridDim.x * blockDim.x) d[i] = s[i];
}

struct AllocKey {
    int v[8 + 220];
    bool operator<(const AllocKey &o) const { return memcmp(v, o.v, sizeof(v)) < 0; }
};
}  // namespace

extern "C" int lteb200_dl_scan_batch(lteb200_ctx *ctx, const void *d_iq, int iq_fmt, long long n_samps, int fs_idx, int N_rb_dl, int N_id_cell, int N_ant,
                                     float phich_res, const long long *start_idx, const int *subfr_num, int n_sf, int turbo_mode, int n_iter,
                                     lteb200_scan_result *results, uint8_t *tb_bits, long long tb_capacity, long long *tb_used, void *stream_)
{
    if (!ctx || !d_iq || !start_idx || !subfr_num || !results || !tb_bits || !tb_used || n_sf < 1 || tb_capacity < 0) return LTEB200_ERR_ARGS;
    if (n_sf > 65535) { lteb200_set_error(ctx, "at most 65535 subframes per call (split the batch)"); return LTEB200_ERR_ARGS; }
    cudaStream_t stream = (cudaStream_t)stream_;
    StreamScope scope;               // stream-ordered temporaries are released on every exit path
    TbBytesScope bytes_out(ctx);     // tb_bits is one bit per byte, whatever the context's transport-block format
    LTE_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    *tb_used = 0;
    const int nsc = 12 * N_rb_dl;
    const size_t grid_row = (size_t)2 * 16 * nsc, ce_row = (size_t)2 * 14 * nsc * N_ant;     // floats per subframe
    float *d_grid = nullptr, *d_ce = nullptr;
    lteb200_pdcch_raw *d_raw = nullptr;
    LTE_CUDA_OK(ctx, scope.alloc(&d_grid, sizeof(float) * grid_row * n_sf, stream));
    LTE_CUDA_OK(ctx, scope.alloc(&d_ce, sizeof(float) * ce_row * n_sf, stream));
    LTE_CUDA_OK(ctx, scope.alloc(&d_raw, sizeof(lteb200_pdcch_raw) * n_sf, stream));
    auto release = [&]() { scope.release(d_grid, stream); scope.release(d_ce, stream); scope.release(d_raw, stream); };
    int rc = lteb200_get_dl_subframe_and_ce_batch(ctx, d_iq, iq_fmt, n_samps, fs_idx, N_rb_dl, N_id_cell, N_ant, start_idx, subfr_num, n_sf, d_grid, d_ce, stream);
    if (rc == LTEB200_OK) rc = lteb200_pdcch_decode_batch(ctx, d_grid, d_ce, N_rb_dl, N_id_cell, N_ant, phich_res, subfr_num, n_sf, d_raw, stream);
    std::vector<lteb200_pdcch_raw> raw(n_sf);
    if (rc == LTEB200_OK && cudaMemcpyAsync(raw.data(), d_raw, sizeof(lteb200_pdcch_raw) * n_sf, cudaMemcpyDeviceToHost, stream) != cudaSuccess) rc = LTEB200_ERR_CUDA;
    if (rc == LTEB200_OK && cudaStreamSynchronize(stream) != cudaSuccess) { lteb200_set_error(ctx, "dl_scan: %s", cudaGetErrorString(cudaGetLastError())); rc = LTEB200_ERR_CUDA; }
    if (rc != LTEB200_OK) { release(); return rc; }

    // ---- DCI unpacking; work list of (subframe, allocation) grouped by what has to be decoded
    std::map<AllocKey, std::vector<std::pair<int, int>>> groups;
    for (int b = 0; b < n_sf; b++) {
        lteb200_scan_result &R = results[b];
        memset(&R, 0, sizeof(R));
        for (int a = 0; a < 6; a++) R.tb_offset[a] = -1;
        R.cfi = raw[b].cfi; R.N_symbs = raw[b].N_symbs;
        int n = 0;
        R.pdcch_err = lteb200_pdcch_unpack(&raw[b], N_rb_dl, N_ant, R.dci, &n);
        R.N_alloc = n;
        if (R.pdcch_err != LTEB200_OK) continue;          // the scan only decodes PDSCH when the PDCCH call returned success (fs.cc:453, 499)
        for (int a = 0; a < n; a++) {
            const lteb200_alloc &al = R.dci[a].alloc;
            bool sane = al.N_prb >= 1 && al.N_prb <= N_rb_dl && al.tbs >= 1 && al.tbs + 24 <= 6144;
            for (int j = 0; j < al.N_prb && sane; j++) sane = al.prb[0][j] >= 0 && al.prb[0][j] < N_rb_dl && al.prb[1][j] >= 0 && al.prb[1][j] < N_rb_dl;
            if (!sane) continue;                          // nothing decodable was announced: crc_ok stays 0, as a failed decode
            AllocKey k;
            memset(&k, 0, sizeof(k));
            k.v[0] = al.mod_type; k.v[1] = al.N_prb; k.v[2] = al.tbs; k.v[3] = al.rv_idx; k.v[4] = al.tx_mode; k.v[5] = al.rnti; k.v[6] = R.N_symbs;
            for (int j = 0; j < al.N_prb && j < 110; j++) { k.v[8 + j] = al.prb[0][j]; k.v[118 + j] = al.prb[1][j]; }
            groups[k].push_back(std::make_pair(b, a));
        }
    }
    // ---- one batched PDSCH receive per distinct allocation
    long long used = 0;
    for (auto &kv : groups) {
        const std::vector<std::pair<int, int>> &members = kv.second;
        const int m = (int)members.size();
        const lteb200_alloc &al = results[members[0].first].dci[members[0].second].alloc;
        const int N_symbs = kv.first.v[6], tbs = al.tbs;
        if (used + (long long)m * tbs > tb_capacity) { lteb200_set_error(ctx, "dl_scan: tb_bits capacity exceeded"); rc = LTEB200_ERR_ARGS; break; }
        std::vector<int> rows(m), sn(m);
        for (int i = 0; i < m; i++) { rows[i] = members[i].first; sn[i] = subfr_num[members[i].first]; }
        int *d_rows = nullptr;
        float *g2 = nullptr, *c2 = nullptr;
        uint8_t *d_tb = nullptr;
        int32_t *d_ok = nullptr;
        LTE_CUDA_OK(ctx, scope.alloc(&d_rows, sizeof(int) * m, stream));
        LTE_CUDA_OK(ctx, scope.alloc(&g2, sizeof(float) * grid_row * m, stream));
        LTE_CUDA_OK(ctx, scope.alloc(&c2, sizeof(float) * ce_row * m, stream));
        LTE_CUDA_OK(ctx, scope.alloc(&d_tb, (size_t)m * tbs, stream));
        LTE_CUDA_OK(ctx, scope.alloc(&d_ok, sizeof(int32_t) * m, stream));
        rc = lteb200_upload_small(ctx, rows.data(), sizeof(int) * m, d_rows, stream);
        if (rc == LTEB200_OK) {
            gather_rows_kernel<<<dim3(8, m), 256, 0, stream>>>((const float4 *)d_grid, d_rows, (int)(grid_row / 4), (float4 *)g2);
            gather_rows_kernel<<<dim3(8, m), 256, 0, stream>>>((const float4 *)d_ce, d_rows, (int)(ce_row / 4), (float4 *)c2);
            ctx->launches += 2;
            rc = lteb200_pdsch_decode_batch(ctx, g2, c2, N_rb_dl, N_id_cell, N_ant, sn.data(), m, N_symbs, &al, 1, turbo_mode, n_iter, d_tb, tbs, d_ok, stream);
        }
        std::vector<int32_t> ok(m);
        if (rc == LTEB200_OK) {
            cudaMemcpyAsync(tb_bits + used, d_tb, (size_t)m * tbs, cudaMemcpyDeviceToHost, stream);
            cudaMemcpyAsync(ok.data(), d_ok, sizeof(int32_t) * m, cudaMemcpyDeviceToHost, stream);
            if (cudaStreamSynchronize(stream) != cudaSuccess) { lteb200_set_error(ctx, "dl_scan: %s", cudaGetErrorString(cudaGetLastError())); rc = LTEB200_ERR_CUDA; }
        } else if (rc == LTEB200_ERR_ARGS || rc == LTEB200_ERR_UNSUPPORTED) {
            rc = LTEB200_OK;                              // an allocation the receive path cannot take (e.g. no resource elements left): not decoded
            std::fill(ok.begin(), ok.end(), 0);
        }
        scope.release(d_rows, stream); scope.release(g2, stream); scope.release(c2, stream); scope.release(d_tb, stream); scope.release(d_ok, stream);
        if (rc != LTEB200_OK) break;
        for (int i = 0; i < m; i++) {
            lteb200_scan_result &R = results[members[i].first];
            R.crc_ok[members[i].second] = ok[i];
            R.tb_offset[members[i].second] = ok[i] ? used + (long long)i * tbs : -1;
        }
        used += (long long)m * tbs;
    }
    *tb_used = used;
    release();
    return rc;
}

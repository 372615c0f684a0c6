// aucell_rss.cu -- regulon specificity scores of an AUC matrix.
//
// Reference: pyscenic.rss.regulon_specificity_scores                                     src/pyscenic/rss.py:8-33
//     rss[t, r] = 1 - jensenshannon(auc[:, r] / sum(auc[:, r]), [label == t] / n_t)     (scipy: sqrt of the JS divergence,
//                                                                                         natural logarithm)
// With p_i = auc[i, r] / A_r, u = 1 / n_t and m = (p + q) / 2 the divergence of (regulon r, type t) splits into a part
// over the cells of type t and a closed form for all other cells (q_i = 0  =>  p_i log(p_i / m_i) = p_i log 2):
//     2 JSD = log 2 * (1 - P_t) + sum_{i in t} [ p_i log(2 p_i / (p_i + u)) + u log(2 u / (p_i + u)) ],   P_t = sum_{i in t} p_i
// so every cell contributes to exactly one (type, regulon) row: ONE pass over the AUC matrix after the column sums.
//
// Deterministic: cells are split into fixed chunks, each chunk is reduced sequentially by the thread that owns the
// regulon column (no atomics), and the chunk partials are added in chunk order.  fp64 throughout; the reference stores
// float32, which is what the caller casts to.
#include "launch.cuh"

namespace {

using namespace aucell_host;

constexpr int kRssThreads = 128;          // regulon columns per block (coalesced row-major reads)

// partial column sums: block (x = column tile, y = cell chunk) -> part[y][r]
__global__ void __launch_bounds__(kRssThreads) rss_colsum_kernel(const double* __restrict__ auc, int64_t n_cells,
                                                                 int64_t row_stride, int n_reg, int64_t chunk,
                                                                 double* __restrict__ part) {
    const int r = blockIdx.x * kRssThreads + threadIdx.x;
    if (r >= n_reg) return;
    const int64_t c0 = (int64_t)blockIdx.y * chunk, c1 = min(n_cells, c0 + chunk);
    double s = 0.0;
    for (int64_t i = c0; i < c1; ++i) s += auc[i * row_stride + r];
    part[(int64_t)blockIdx.y * n_reg + r] = s;
}

__global__ void __launch_bounds__(kRssThreads) rss_reduce_cols_kernel(const double* __restrict__ part, int n_chunks,
                                                                      int n_reg, double* __restrict__ colsum) {
    const int r = blockIdx.x * kRssThreads + threadIdx.x;
    if (r >= n_reg) return;
    double s = 0.0;
    for (int b = 0; b < n_chunks; ++b) s += part[(int64_t)b * n_reg + r];
    colsum[r] = s;
}

// per (chunk, type, regulon): P = sum p, S = sum [p log(2p/(p+u)) + u log(2u/(p+u))] over the chunk's cells of that type
__global__ void __launch_bounds__(kRssThreads) rss_partial_kernel(const double* __restrict__ auc,
                                                                  const int32_t* __restrict__ labels, int64_t n_cells,
                                                                  int64_t row_stride, int n_reg, int n_types,
                                                                  int64_t chunk, const double* __restrict__ colsum,
                                                                  const double* __restrict__ inv_count,
                                                                  double* __restrict__ part) {
    extern __shared__ double sm[];            // [n_types][kRssThreads][2]
    const int r = blockIdx.x * kRssThreads + threadIdx.x;
    for (int i = threadIdx.x; i < n_types * kRssThreads * 2; i += kRssThreads) sm[i] = 0.0;
    __syncthreads();
    if (r < n_reg) {
        const int64_t c0 = (int64_t)blockIdx.y * chunk, c1 = min(n_cells, c0 + chunk);
        const double inv_a = 1.0 / colsum[r];          // inf for an all-zero regulon: p becomes NaN/inf like the reference
        for (int64_t i = c0; i < c1; ++i) {
            const int t = labels[i];                   // warp-uniform: broadcast load
            const double u = inv_count[t];
            const double p = auc[i * row_stride + r] * inv_a;
            double s = u * log(2.0 * u / (p + u));
            if (p != 0.0) s += p * log(2.0 * p / (p + u));
            double* acc = sm + ((size_t)t * kRssThreads + threadIdx.x) * 2;   // only this thread touches column r
            acc[0] += p;
            acc[1] += s;
        }
        for (int t = 0; t < n_types; ++t) {
            const double* acc = sm + ((size_t)t * kRssThreads + threadIdx.x) * 2;
            double* o = part + (((int64_t)blockIdx.y * n_types + t) * n_reg + r) * 2;
            o[0] = acc[0];
            o[1] = acc[1];
        }
    }
}

__global__ void __launch_bounds__(kRssThreads) rss_finish_kernel(const double* __restrict__ part, int n_chunks,
                                                                 int n_reg, int n_types, double* __restrict__ out) {
    const int r = blockIdx.x * kRssThreads + threadIdx.x;
    const int t = blockIdx.y;
    if (r >= n_reg) return;
    double P = 0.0, S = 0.0;
    for (int b = 0; b < n_chunks; ++b) {
        const double* o = part + (((int64_t)b * n_types + t) * n_reg + r) * 2;
        P += o[0];
        S += o[1];
    }
    const double js = 0.5 * (0.6931471805599453 * (1.0 - P) + S);
    out[(int64_t)t * n_reg + r] = 1.0 - sqrt(js < 0.0 ? 0.0 : js);   // rounding can leave js = -1e-17 when p == q; NaN (all-zero regulon) passes through
}

// np.count_nonzero per row of a CSR matrix (derive_auc_threshold, reference aucell.py:67): NaN counts, -0.0 does not.
// One warp per row.
__global__ void __launch_bounds__(256) count_nonzero_csr_kernel(const int64_t* __restrict__ indptr, const float* __restrict__ data,
                                                                 int64_t n_rows, int32_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    for (int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < n_rows; row += (int64_t)gridDim.x * 8) {
        const int64_t b = indptr[row], e = indptr[row + 1];
        int c = 0;
        for (int64_t i = b + lane; i < e; i += 32) c += (data[i] != 0.0f) ? 1 : 0;
        c = __reduce_add_sync(0xffffffffu, c);
        if (lane == 0) out[row] = c;
    }
}

}  // namespace

extern "C" int aucell_count_nonzero_csr_f32(int device, const int64_t* d_indptr, const float* d_data, int64_t n_cells,
                                            int32_t* d_out_nnz, void* stream) {
    if (n_cells < 0) return fail(AUCELL_ERR_INVALID, "n_cells < 0");
    if (n_cells == 0) return AUCELL_OK;
    if (!d_indptr || !d_out_nnz) return fail(AUCELL_ERR_INVALID, "NULL pointer");
    DeviceGuard dg(device);
    if (!dg.ok) return fail(AUCELL_ERR_CUDA, "cudaSetDevice failed");
    const int grid = (int)std::min<int64_t>((n_cells + 7) / 8, 148 * 16);
    count_nonzero_csr_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_indptr, d_data, n_cells, d_out_nnz);
    launch_counter().fetch_add(1);
    CK(cudaGetLastError());
    return AUCELL_OK;
}

extern "C" int aucell_rss_f64(int device, const double* d_auc, int64_t n_cells, int64_t row_stride, int32_t n_regulons,
                              const int32_t* d_labels, int32_t n_types, const int64_t* h_type_counts, double* d_out,
                              void* stream) {
    if (!d_auc || !d_labels || !h_type_counts || !d_out) return fail(AUCELL_ERR_INVALID, "NULL pointer");
    if (n_cells <= 0 || n_regulons <= 0 || n_types <= 0 || row_stride < n_regulons)
        return fail(AUCELL_ERR_INVALID, "rss: need n_cells, n_regulons, n_types > 0 and row_stride >= n_regulons");
    int64_t tot = 0;
    for (int t = 0; t < n_types; ++t) {
        if (h_type_counts[t] <= 0) return fail(AUCELL_ERR_INVALID, "rss: every cell type needs at least one cell");
        tot += h_type_counts[t];
    }
    if (tot != n_cells) return fail(AUCELL_ERR_INVALID, "rss: type counts do not add up to n_cells");
    DeviceGuard dg(device);
    if (!dg.ok) return fail(AUCELL_ERR_CUDA, "cudaSetDevice failed");
    cudaStream_t st = (cudaStream_t)stream;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int tiles = (n_regulons + kRssThreads - 1) / kRssThreads;
    // enough chunks to fill the GPU a few times over, at least 256 cells each
    int64_t n_chunks = std::max<int64_t>(1, std::min<int64_t>((n_cells + 255) / 256, (int64_t)(8 * sms + tiles - 1) / tiles));
    const int64_t chunk = (n_cells + n_chunks - 1) / n_chunks;
    n_chunks = (n_cells + chunk - 1) / chunk;
    const size_t smem = (size_t)n_types * kRssThreads * 2 * sizeof(double);
    int optin = 0;
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (smem > (size_t)optin) return fail(AUCELL_ERR_UNSUPPORTED, "rss: too many cell types for the shared-memory accumulators");
    // scratch: colsum [R] | inv_count [T] | partials [n_chunks][T][R][2]  (column-sum partials reuse the last block)
    const size_t n_part = (size_t)n_chunks * n_types * n_regulons * 2;
    const size_t bytes = sizeof(double) * ((size_t)n_regulons + n_types + n_part);
    double* scratch = nullptr;
    CK(cudaMallocAsync((void**)&scratch, bytes, st));
    double* colsum = scratch;
    double* inv_count = scratch + n_regulons;
    double* part = inv_count + n_types;
    std::vector<double> h_inv((size_t)n_types);
    for (int t = 0; t < n_types; ++t) h_inv[t] = 1.0 / (double)h_type_counts[t];
    int rc = AUCELL_OK;
    do {
        cudaError_t e = cudaMemcpyAsync(inv_count, h_inv.data(), sizeof(double) * n_types, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);      // h_inv is a local: the copy must have left it
        if (e != cudaSuccess) { rc = fail_cuda(e, "rss: upload", __LINE__); break; }
        const dim3 grid(tiles, (unsigned)n_chunks);
        rss_colsum_kernel<<<grid, kRssThreads, 0, st>>>(d_auc, n_cells, row_stride, n_regulons, chunk, part);
        rss_reduce_cols_kernel<<<tiles, kRssThreads, 0, st>>>(part, (int)n_chunks, n_regulons, colsum);
        e = cudaFuncSetAttribute(rss_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { rc = fail_cuda(e, "rss: shared memory", __LINE__); break; }
        rss_partial_kernel<<<grid, kRssThreads, smem, st>>>(d_auc, d_labels, n_cells, row_stride, n_regulons, n_types,
                                                            chunk, colsum, inv_count, part);
        rss_finish_kernel<<<dim3(tiles, n_types), kRssThreads, 0, st>>>(part, (int)n_chunks, n_regulons, n_types, d_out);
        launch_counter().fetch_add(4);
        e = cudaGetLastError();
        if (e != cudaSuccess) rc = fail_cuda(e, "rss: launch", __LINE__);
    } while (0);
    cudaFreeAsync(scratch, st);
    return rc;
}

// aucell_f64_dense.cu -- float64 dense-input instantiations of the AUCell kernels; a translation unit of its own so that the kernel
// variants compile in parallel (see pyscenic_b200/_native.py: TRANSLATION_UNITS).
// float64 inputs whose values are not exactly representable in float32 must be ranked in float64: a cast would create
// ties that do not exist in the input.
#include "launch.cuh"

namespace aucell_host {

int rank_dense_f64(const aucell_plan* p, const double* d, int64_t n_cells, int64_t row_stride, uint32_t* out_rank,
                   int32_t* out_nnz, void* stream) {
    return run_rank<double, IN_DENSE>(p, d, row_stride, nullptr, nullptr, nullptr, n_cells, out_rank, out_nnz, stream);
}
int fused_dense_f64(const aucell_plan* p, const double* d, int64_t n_cells, int64_t row_stride, double* out_auc,
                    int32_t* out_nnz, void* stream) {
    return run_fused<double, IN_DENSE>(p, d, row_stride, nullptr, nullptr, nullptr, n_cells, out_auc, out_nnz, stream);
}

}  // namespace aucell_host

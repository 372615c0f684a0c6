// aucell_f32_dense.cu -- float32 dense-input instantiations of the AUCell kernels; a translation unit of its own so that the kernel
// variants compile in parallel (see pyscenic_b200/_native.py: TRANSLATION_UNITS).
#include "launch.cuh"

namespace aucell_host {

int rank_dense_f32(const aucell_plan* p, const float* d, int64_t n_cells, int64_t row_stride, uint32_t* out_rank,
                   int32_t* out_nnz, void* stream) {
    return run_rank<float, IN_DENSE>(p, d, row_stride, nullptr, nullptr, nullptr, n_cells, out_rank, out_nnz, stream);
}
int fused_dense_f32(const aucell_plan* p, const float* d, int64_t n_cells, int64_t row_stride, double* out_auc,
                    int32_t* out_nnz, void* stream) {
    return run_fused<float, IN_DENSE>(p, d, row_stride, nullptr, nullptr, nullptr, n_cells, out_auc, out_nnz, stream);
}

}  // namespace aucell_host

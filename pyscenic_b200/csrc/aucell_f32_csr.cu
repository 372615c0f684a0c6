// aucell_f32_csr.cu -- float32 CSR-input instantiations of the AUCell kernels; a translation unit of its own so that the kernel
// variants compile in parallel (see pyscenic_b200/_native.py: TRANSLATION_UNITS).
#include "launch.cuh"

namespace aucell_host {

int rank_csr_f32(const aucell_plan* p, const int64_t* indptr, const int32_t* indices, const float* data,
                 int64_t n_cells, uint32_t* out_rank, int32_t* out_nnz, void* stream) {
    return run_rank<float, IN_CSR>(p, nullptr, 0, indptr, indices, data, n_cells, out_rank, out_nnz, stream);
}
int fused_csr_f32(const aucell_plan* p, const int64_t* indptr, const int32_t* indices, const float* data,
                  int64_t n_cells, double* out_auc, int32_t* out_nnz, void* stream) {
    return run_fused<float, IN_CSR>(p, nullptr, 0, indptr, indices, data, n_cells, out_auc, out_nnz, stream);
}

int fused_csr_lut8(const aucell_plan* p, const int64_t* indptr, const int32_t* indices, const uint8_t* codes,
                   const float* lut, float* scratch_vals, int64_t n_cells, double* out_auc, int32_t* out_nnz,
                   void* stream) {
    return run_fused_lut8<IN_CSR_U8>(p, indptr, indices, codes, lut, scratch_vals, n_cells, out_auc, out_nnz, stream);
}

}  // namespace aucell_host

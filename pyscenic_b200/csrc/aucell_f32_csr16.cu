// aucell_f32_csr16.cu -- float32 CSR input with uint16 column indices (n_genes <= 65535): 6 instead of 8 bytes per stored
// entry.  What the host-buffer pipeline uploads (it narrows the caller's int32 indices chunk by chunk), and an entry
// point of its own for callers that already hold 16-bit indices.  A translation unit of its own so that the kernel
// variants compile in parallel (see pyscenic_b200/_native.py: TRANSLATION_UNITS).
#include "launch.cuh"

namespace aucell_host {

int fused_csr16_f32(const aucell_plan* p, const int64_t* indptr, const uint16_t* indices, const float* data,
                    int64_t n_cells, double* out_auc, int32_t* out_nnz, void* stream) {
    return run_fused<float, IN_CSR16>(p, nullptr, 0, indptr, reinterpret_cast<const int32_t*>(indices), data, n_cells,
                                      out_auc, out_nnz, stream);
}

int fused_csr16_lut8(const aucell_plan* p, const int64_t* indptr, const uint16_t* indices, const uint8_t* codes,
                     const float* lut, float* scratch_vals, int64_t n_cells, double* out_auc, int32_t* out_nnz,
                     void* stream) {
    return run_fused_lut8<IN_CSR16_U8>(p, indptr, indices, codes, lut, scratch_vals, n_cells, out_auc, out_nnz, stream);
}

}  // namespace aucell_host

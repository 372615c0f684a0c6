// aucell_b200.cu -- C-ABI of libaucell_b200.so (see include/aucell_b200.h): plan management, kernel launch
// configuration and the host-buffer streaming pipeline.  sm_100a only; no CPU fallback.
#include "aucell_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "cell_kernels.cuh"

using namespace aucell;

namespace {

thread_local std::string g_err;
std::atomic<int64_t> g_launches{0};

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

int fail_cuda(cudaError_t e, const char* what, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) at aucell_b200.cu:%d: %s", (int)e, cudaGetErrorString(e), line, what);
    g_err = buf;
    return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? AUCELL_ERR_NO_DEVICE : AUCELL_ERR_CUDA;
}

#define CK(call)                                                          \
    do {                                                                  \
        cudaError_t e_ = (call);                                          \
        if (e_ != cudaSuccess) return fail_cuda(e_, #call, __LINE__);     \
    } while (0)

constexpr int64_t kMaxGenes = 1 << 20;  // bitmap words / int positions; far above any gene annotation

int next_pow2(int64_t v) {
    int64_t p = 1;
    while (p < v) p <<= 1;
    return (int)p;
}

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess) ok = true;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

}  // namespace

struct aucell_plan {
    int device = 0;
    int sm_count = 0;
    int64_t smem_optin = 0;
    int n_genes = 0;
    int n_words = 0;
    int64_t rank_cutoff = 0;
    bool pos16 = true;        // shuffled positions / table entries fit uint16
    bool weighted = false;
    int n_regulons = 0;
    int64_t reg_nnz = 0;
    // device arrays
    int32_t* d_inv_perm = nullptr;
    int32_t* d_reg_indptr = nullptr;
    void* d_reg_pos = nullptr;
    double* d_reg_weight = nullptr;
    double* d_reg_max_auc = nullptr;
    uint8_t* d_reg_valid = nullptr;
    // overflow scratch lists, one per stream that ever needed one
    mutable std::mutex mu;
    mutable std::map<std::pair<cudaStream_t, int>, std::pair<void*, size_t>> scratch;
};

namespace {

RegulonsDev reg_of(const aucell_plan* p) {
    RegulonsDev r;
    r.n_regulons = p->n_regulons;
    r.indptr = p->d_reg_indptr;
    r.pos = p->d_reg_pos;
    r.weight = p->d_reg_weight;
    r.max_auc = p->d_reg_max_auc;
    r.valid = p->d_reg_valid;
    return r;
}

// Order the entries of one regulon so that consecutive groups of 32 (one warp load of the gather) touch
// distinct shared-memory banks of the per-cell table where possible: round-robin over banks.
void stagger_banks(std::vector<std::pair<int32_t, double>>& ent, bool pos16) {
    const int shift = pos16 ? 1 : 0;  // two uint16 table entries per 4-byte bank word
    std::vector<std::vector<std::pair<int32_t, double>>> bank(32);
    for (auto& e : ent) bank[(e.first >> shift) & 31].push_back(e);
    for (auto& b : bank) std::sort(b.begin(), b.end());
    size_t out = 0, level = 0;
    while (out < ent.size()) {
        for (int b = 0; b < 32; ++b)
            if (level < bank[b].size()) ent[out++] = bank[b][level];
        ++level;
    }
}

struct SmemLayout {
    int cap = 0;
    int off_key = 0, off_pos = 0, off_bits = 0, off_wpre = 0, off_tbl = 0;
    int total = 0;
    int ctas_per_sm = 1;
};

int align16(int v) { return (v + 15) & ~15; }

// Pick the smem list capacity and the number of resident CTAs per SM.
bool plan_smem(const aucell_plan* p, int key_bytes, bool fused, SmemLayout* L) {
    const int pos_bytes = p->pos16 ? 2 : 4;
    const int fixed = align16(p->n_words * 4) * 2 + (fused ? align16(p->n_genes * pos_bytes) : 0);
    const int per_entry = key_bytes + pos_bytes;
    const int want = std::min(next_pow2(p->n_genes), 4096);
    const int64_t sm_total = 228 * 1024;  // per SM, 1 KB reserved per resident CTA
    for (int k = 4; k >= 1; --k) {
        int64_t budget = std::min<int64_t>(p->smem_optin, sm_total / k - 1024);
        int64_t room = budget - fixed - 64;
        if (room < (int64_t)per_entry * 2) continue;
        int cap = 2;
        while ((int64_t)cap * 2 * per_entry <= room && cap * 2 <= next_pow2(p->n_genes)) cap <<= 1;
        if (cap < want && k > 1) continue;
        L->cap = cap;
        L->ctas_per_sm = k;
        int off = 0;
        L->off_key = off; off = align16(off + cap * key_bytes);
        L->off_pos = off; off = align16(off + cap * pos_bytes);
        L->off_bits = off; off = align16(off + p->n_words * 4);
        L->off_wpre = off; off = align16(off + p->n_words * 4);
        L->off_tbl = off; off = align16(off + (fused ? p->n_genes * pos_bytes : 0));
        L->total = off;
        return true;
    }
    return false;
}

int get_scratch(const aucell_plan* p, cudaStream_t st, int key_bytes, int grid, int scratch_cap, void** key,
                void** pos) {
    const int pos_bytes = p->pos16 ? 2 : 4;
    const size_t need = (size_t)grid * scratch_cap * (key_bytes + pos_bytes);
    std::lock_guard<std::mutex> lk(p->mu);
    auto& slot = p->scratch[{st, key_bytes}];
    if (slot.second < need) {
        if (slot.first) cudaFree(slot.first);
        slot = {nullptr, 0};
        void* ptr = nullptr;
        cudaError_t e = cudaMalloc(&ptr, need);
        if (e != cudaSuccess) return fail_cuda(e, "cudaMalloc(scratch)", __LINE__);
        slot = {ptr, need};
    }
    *key = slot.first;
    *pos = (char*)slot.first + (size_t)grid * scratch_cap * key_bytes;
    return AUCELL_OK;
}

template <typename ValT, typename PosT, int IN, int OUT>
int launch_cell(const aucell_plan* p, CellArgs<ValT>& a, const SmemLayout& L, int grid, cudaStream_t st) {
    if (OUT == OUT_AUC && p->weighted) {
        auto k = aucell_cell_kernel<ValT, PosT, IN, OUT, true>;
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
        k<<<grid, kBlock, L.total, st>>>(a);
    } else {
        auto k = aucell_cell_kernel<ValT, PosT, IN, OUT, false>;
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
        k<<<grid, kBlock, L.total, st>>>(a);
    }
    g_launches.fetch_add(1);
    CK(cudaGetLastError());
    return AUCELL_OK;
}

template <typename ValT, int IN, int OUT>
int run_cells(const aucell_plan* p, const ValT* dense, int64_t row_stride, const int64_t* indptr,
              const int32_t* indices, const ValT* data, int64_t n_cells, uint32_t* out_rank, double* out_auc,
              int32_t* out_nnz, void* stream) {
    if (!p) return fail(AUCELL_ERR_INVALID, "plan is NULL");
    if (n_cells < 0) return fail(AUCELL_ERR_INVALID, "n_cells < 0");
    if (IN == IN_DENSE && (!dense || row_stride < p->n_genes))
        if (n_cells > 0) return fail(AUCELL_ERR_INVALID, "dense input: NULL pointer or row_stride < n_genes");
    if (IN == IN_CSR && n_cells > 0 && (!indptr || (!indices && !data)))
        return fail(AUCELL_ERR_INVALID, "CSR input: NULL pointer");
    if (OUT == OUT_RANK && !out_rank && n_cells > 0) return fail(AUCELL_ERR_INVALID, "out_rank is NULL");
    if (OUT == OUT_AUC) {
        if (p->n_regulons <= 0) return fail(AUCELL_ERR_INVALID, "plan has no regulons");
        if (!out_auc && n_cells > 0) return fail(AUCELL_ERR_INVALID, "out_auc is NULL");
    }
    if (n_cells == 0) return AUCELL_OK;
    DeviceGuard dg(p->device);
    if (!dg.ok) return fail(AUCELL_ERR_CUDA, "cudaSetDevice failed");
    cudaStream_t st = (cudaStream_t)stream;

    SmemLayout L;
    const int key_bytes = (int)sizeof(typename KeyOf<ValT>::type);
    if (!plan_smem(p, key_bytes, OUT == OUT_AUC, &L))
        return fail(AUCELL_ERR_UNSUPPORTED, "n_genes too large for the shared-memory kernels");
    int grid = (int)std::min<int64_t>(n_cells, (int64_t)p->sm_count * L.ctas_per_sm);

    CellArgs<ValT> a;
    memset(&a, 0, sizeof a);
    a.dense = dense; a.row_stride = row_stride;
    a.indptr = indptr; a.indices = indices; a.data = data;
    a.n_cells = n_cells; a.n_genes = p->n_genes; a.n_words = p->n_words;
    a.inv_perm = p->d_inv_perm;
    a.rank_cutoff = (int)p->rank_cutoff;
    a.reg = reg_of(p);
    a.out_rank = out_rank; a.out_auc = out_auc; a.out_nnz = out_nnz;
    a.cap = L.cap;
    a.off_key = L.off_key; a.off_pos = L.off_pos; a.off_bits = L.off_bits; a.off_wpre = L.off_wpre;
    a.off_tbl = L.off_tbl;
    a.scratch_cap = next_pow2(p->n_genes);
    if (L.cap < a.scratch_cap) {
        int rc = get_scratch(p, st, key_bytes, grid, a.scratch_cap, &a.scratch_key, &a.scratch_pos);
        if (rc) return rc;
    }
    if (p->pos16) return launch_cell<ValT, uint16_t, IN, OUT>(p, a, L, grid, st);
    return launch_cell<ValT, uint32_t, IN, OUT>(p, a, L, grid, st);
}

}  // namespace

// ------------------------------------------------------------------------------------------------------
extern "C" {

int aucell_b200_abi_version(void) { return AUCELL_B200_ABI_VERSION; }

const char* aucell_b200_last_error(void) { return g_err.c_str(); }

int64_t aucell_b200_launch_count(void) { return g_launches.load(); }

int64_t aucell_b200_max_genes(void) { return 100000; }

int aucell_b200_device_info(int device, int* device_count, int* cc, int* sm_count, int64_t* smem_optin_bytes) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (device_count) *device_count = (e == cudaSuccess) ? n : -1;
    if (e != cudaSuccess) return fail_cuda(e, "cudaGetDeviceCount", __LINE__);
    if (n == 0) return fail(AUCELL_ERR_NO_DEVICE, "no CUDA device");
    if (device < 0 || device >= n) return fail(AUCELL_ERR_INVALID, "device index out of range");
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (cc) *cc = prop.major * 10 + prop.minor;
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (smem_optin_bytes) *smem_optin_bytes = (int64_t)prop.sharedMemPerBlockOptin;
    return AUCELL_OK;
}

int aucell_plan_create(int device, int64_t n_genes, const int32_t* h_perm, int32_t n_regulons,
                       const int64_t* h_reg_indptr, const int32_t* h_reg_cols, const double* h_reg_weights,
                       const double* h_reg_max_auc, const uint8_t* h_reg_valid, int64_t rank_cutoff,
                       aucell_plan_t** out_plan) {
    if (!out_plan) return fail(AUCELL_ERR_INVALID, "out_plan is NULL");
    *out_plan = nullptr;
    if (n_genes <= 0 || n_genes > kMaxGenes) return fail(AUCELL_ERR_INVALID, "n_genes out of range");
    if (n_regulons < 0) return fail(AUCELL_ERR_INVALID, "n_regulons < 0");
    if (n_regulons > 0) {
        if (!h_reg_indptr || !h_reg_max_auc || !h_reg_valid) return fail(AUCELL_ERR_INVALID, "regulon arrays are NULL");
        if (rank_cutoff < 0 || rank_cutoff >= n_genes)
            return fail(AUCELL_ERR_INVALID, "rank_cutoff must satisfy 0 <= rank_cutoff < n_genes");
        if (h_reg_indptr[0] != 0) return fail(AUCELL_ERR_INVALID, "reg_indptr[0] != 0");
        for (int r = 0; r < n_regulons; ++r)
            if (h_reg_indptr[r + 1] < h_reg_indptr[r]) return fail(AUCELL_ERR_INVALID, "reg_indptr not monotone");
        if (h_reg_indptr[n_regulons] > 0 && !h_reg_cols) return fail(AUCELL_ERR_INVALID, "reg_cols is NULL");
        if (h_reg_indptr[n_regulons] > INT32_MAX) return fail(AUCELL_ERR_UNSUPPORTED, "more than 2^31 regulon entries");
    }
    int count = 0, cc = 0, sms = 0;
    int64_t optin = 0;
    int rc = aucell_b200_device_info(device, &count, &cc, &sms, &optin);
    if (rc) return rc;
    if (cc < 100) return fail(AUCELL_ERR_NO_DEVICE, "device is not sm_100 (Blackwell); this library has no other code path");
    DeviceGuard dg(device);
    if (!dg.ok) return fail(AUCELL_ERR_CUDA, "cudaSetDevice failed");

    aucell_plan* p = new (std::nothrow) aucell_plan();
    if (!p) return fail(AUCELL_ERR_NOMEM, "out of host memory");
    p->device = device;
    p->sm_count = sms;
    p->smem_optin = optin;
    p->n_genes = (int)n_genes;
    p->n_words = (int)((n_genes + 31) / 32);
    p->rank_cutoff = n_regulons > 0 ? rank_cutoff : 0;
    p->pos16 = n_genes <= 65535;
    p->weighted = h_reg_weights != nullptr;
    p->n_regulons = n_regulons;

    // inverse permutation: gene column -> shuffled position
    std::vector<int32_t> inv((size_t)n_genes);
    if (h_perm) {
        std::vector<char> seen((size_t)n_genes, 0);
        for (int64_t j = 0; j < n_genes; ++j) {
            int32_t g = h_perm[j];
            if (g < 0 || g >= n_genes || seen[g]) {
                delete p;
                return fail(AUCELL_ERR_INVALID, "perm is not a permutation of 0..n_genes-1");
            }
            seen[g] = 1;
            inv[g] = (int32_t)j;
        }
    } else {
        for (int64_t g = 0; g < n_genes; ++g) inv[g] = (int32_t)g;
    }

#define PCK(call)                                                      \
    do {                                                               \
        cudaError_t e_ = (call);                                       \
        if (e_ != cudaSuccess) {                                       \
            int rc_ = fail_cuda(e_, #call, __LINE__);                  \
            aucell_plan_destroy(p);                                    \
            return rc_;                                                \
        }                                                              \
    } while (0)

    PCK(cudaMalloc(&p->d_inv_perm, sizeof(int32_t) * n_genes));
    PCK(cudaMemcpy(p->d_inv_perm, inv.data(), sizeof(int32_t) * n_genes, cudaMemcpyHostToDevice));

    if (n_regulons > 0) {
        const int64_t nnz = h_reg_indptr[n_regulons];
        p->reg_nnz = nnz;
        std::vector<int32_t> indptr32((size_t)n_regulons + 1);
        std::vector<int32_t> pos((size_t)std::max<int64_t>(nnz, 1));
        std::vector<double> wts((size_t)std::max<int64_t>(nnz, 1));
        std::vector<std::pair<int32_t, double>> ent;
        for (int r = 0; r < n_regulons; ++r) {
            indptr32[r] = (int32_t)h_reg_indptr[r];
            ent.clear();
            for (int64_t e = h_reg_indptr[r]; e < h_reg_indptr[r + 1]; ++e) {
                int32_t g = h_reg_cols[e];
                if (g < 0 || g >= n_genes) {
                    aucell_plan_destroy(p);
                    return fail(AUCELL_ERR_INVALID, "regulon gene column out of range");
                }
                ent.emplace_back(inv[g], h_reg_weights ? h_reg_weights[e] : 1.0);
            }
            stagger_banks(ent, p->pos16);
            for (size_t k = 0; k < ent.size(); ++k) {
                pos[h_reg_indptr[r] + k] = ent[k].first;
                wts[h_reg_indptr[r] + k] = ent[k].second;
            }
            if (h_reg_valid[r] && !(h_reg_max_auc[r] > 0.0)) {
                aucell_plan_destroy(p);
                return fail(AUCELL_ERR_INVALID, "max_auc must be > 0 for valid regulons");
            }
        }
        indptr32[n_regulons] = (int32_t)nnz;
        PCK(cudaMalloc(&p->d_reg_indptr, sizeof(int32_t) * (n_regulons + 1)));
        PCK(cudaMemcpy(p->d_reg_indptr, indptr32.data(), sizeof(int32_t) * (n_regulons + 1), cudaMemcpyHostToDevice));
        const size_t nalloc = (size_t)std::max<int64_t>(nnz, 1);
        if (p->pos16) {
            std::vector<uint16_t> pos16v(nalloc);
            for (int64_t e = 0; e < nnz; ++e) pos16v[e] = (uint16_t)pos[e];
            PCK(cudaMalloc(&p->d_reg_pos, sizeof(uint16_t) * nalloc));
            PCK(cudaMemcpy(p->d_reg_pos, pos16v.data(), sizeof(uint16_t) * nalloc, cudaMemcpyHostToDevice));
        } else {
            PCK(cudaMalloc(&p->d_reg_pos, sizeof(uint32_t) * nalloc));
            PCK(cudaMemcpy(p->d_reg_pos, pos.data(), sizeof(uint32_t) * nalloc, cudaMemcpyHostToDevice));
        }
        if (p->weighted) {
            PCK(cudaMalloc(&p->d_reg_weight, sizeof(double) * nalloc));
            PCK(cudaMemcpy(p->d_reg_weight, wts.data(), sizeof(double) * nalloc, cudaMemcpyHostToDevice));
        }
        PCK(cudaMalloc(&p->d_reg_max_auc, sizeof(double) * n_regulons));
        PCK(cudaMemcpy(p->d_reg_max_auc, h_reg_max_auc, sizeof(double) * n_regulons, cudaMemcpyHostToDevice));
        PCK(cudaMalloc(&p->d_reg_valid, n_regulons));
        PCK(cudaMemcpy(p->d_reg_valid, h_reg_valid, n_regulons, cudaMemcpyHostToDevice));
    }
#undef PCK
    *out_plan = p;
    return AUCELL_OK;
}

int aucell_plan_destroy(aucell_plan_t* p) {
    if (!p) return AUCELL_OK;
    DeviceGuard dg(p->device);
    cudaFree(p->d_inv_perm);
    cudaFree(p->d_reg_indptr);
    cudaFree(p->d_reg_pos);
    cudaFree(p->d_reg_weight);
    cudaFree(p->d_reg_max_auc);
    cudaFree(p->d_reg_valid);
    for (auto& kv : p->scratch) cudaFree(kv.second.first);
    delete p;
    return AUCELL_OK;
}

// ---- create_rankings ---------------------------------------------------------------------------------
int aucell_rank_dense_f32(const aucell_plan_t* plan, const float* d_expr, int64_t n_cells, int64_t row_stride,
                          uint32_t* d_out_rank, int32_t* d_out_nnz, void* stream) {
    return run_cells<float, IN_DENSE, OUT_RANK>(plan, d_expr, row_stride, nullptr, nullptr, nullptr, n_cells,
                                                 d_out_rank, nullptr, d_out_nnz, stream);
}
int aucell_rank_dense_f64(const aucell_plan_t* plan, const double* d_expr, int64_t n_cells, int64_t row_stride,
                          uint32_t* d_out_rank, int32_t* d_out_nnz, void* stream) {
    return run_cells<double, IN_DENSE, OUT_RANK>(plan, d_expr, row_stride, nullptr, nullptr, nullptr, n_cells,
                                                  d_out_rank, nullptr, d_out_nnz, stream);
}
int aucell_rank_csr_f32(const aucell_plan_t* plan, const int64_t* d_indptr, const int32_t* d_indices,
                        const float* d_data, int64_t n_cells, uint32_t* d_out_rank, int32_t* d_out_nnz,
                        void* stream) {
    return run_cells<float, IN_CSR, OUT_RANK>(plan, nullptr, 0, d_indptr, d_indices, d_data, n_cells, d_out_rank,
                                               nullptr, d_out_nnz, stream);
}
int aucell_rank_csr_f64(const aucell_plan_t* plan, const int64_t* d_indptr, const int32_t* d_indices,
                        const double* d_data, int64_t n_cells, uint32_t* d_out_rank, int32_t* d_out_nnz,
                        void* stream) {
    return run_cells<double, IN_CSR, OUT_RANK>(plan, nullptr, 0, d_indptr, d_indices, d_data, n_cells, d_out_rank,
                                                nullptr, d_out_nnz, stream);
}

// ---- fused aucell ------------------------------------------------------------------------------------
int aucell_dense_f32(const aucell_plan_t* plan, const float* d_expr, int64_t n_cells, int64_t row_stride,
                     double* d_out_auc, int32_t* d_out_nnz, void* stream) {
    return run_cells<float, IN_DENSE, OUT_AUC>(plan, d_expr, row_stride, nullptr, nullptr, nullptr, n_cells, nullptr,
                                                d_out_auc, d_out_nnz, stream);
}
int aucell_dense_f64(const aucell_plan_t* plan, const double* d_expr, int64_t n_cells, int64_t row_stride,
                     double* d_out_auc, int32_t* d_out_nnz, void* stream) {
    return run_cells<double, IN_DENSE, OUT_AUC>(plan, d_expr, row_stride, nullptr, nullptr, nullptr, n_cells, nullptr,
                                                 d_out_auc, d_out_nnz, stream);
}
int aucell_csr_f32(const aucell_plan_t* plan, const int64_t* d_indptr, const int32_t* d_indices,
                   const float* d_data, int64_t n_cells, double* d_out_auc, int32_t* d_out_nnz, void* stream) {
    return run_cells<float, IN_CSR, OUT_AUC>(plan, nullptr, 0, d_indptr, d_indices, d_data, n_cells, nullptr,
                                              d_out_auc, d_out_nnz, stream);
}
int aucell_csr_f64(const aucell_plan_t* plan, const int64_t* d_indptr, const int32_t* d_indices,
                   const double* d_data, int64_t n_cells, double* d_out_auc, int32_t* d_out_nnz, void* stream) {
    return run_cells<double, IN_CSR, OUT_AUC>(plan, nullptr, 0, d_indptr, d_indices, d_data, n_cells, nullptr,
                                               d_out_auc, d_out_nnz, stream);
}

// ---- aucell4r ----------------------------------------------------------------------------------------
int aucell_from_rankings_u32(const aucell_plan_t* p, const uint32_t* d_rank, int64_t n_cells, int64_t row_stride,
                             double* d_out_auc, void* stream) {
    if (!p) return fail(AUCELL_ERR_INVALID, "plan is NULL");
    if (p->n_regulons <= 0) return fail(AUCELL_ERR_INVALID, "plan has no regulons");
    if (n_cells < 0 || row_stride < p->n_genes) return fail(AUCELL_ERR_INVALID, "bad n_cells / row_stride");
    if (n_cells == 0) return AUCELL_OK;
    if (!d_rank || !d_out_auc) return fail(AUCELL_ERR_INVALID, "NULL pointer");
    DeviceGuard dg(p->device);
    if (!dg.ok) return fail(AUCELL_ERR_CUDA, "cudaSetDevice failed");
    cudaStream_t st = (cudaStream_t)stream;
    const int smem = align16(p->n_genes * (p->pos16 ? 2 : 4));
    if (smem > p->smem_optin) return fail(AUCELL_ERR_UNSUPPORTED, "n_genes too large for the shared-memory kernels");
    const int per_sm = std::max<int>(1, std::min<int64_t>(4, (228 * 1024) / (smem + 1024)));
    const int grid = (int)std::min<int64_t>(n_cells, (int64_t)p->sm_count * per_sm);
    RegulonsDev reg = reg_of(p);
#define LAUNCH_FR(PosT, W)                                                                              \
    do {                                                                                                \
        auto k = aucell_from_rankings_kernel<PosT, W>;                                                  \
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));                 \
        k<<<grid, kBlock, smem, st>>>(d_rank, n_cells, row_stride, p->n_genes, (int)p->rank_cutoff, reg, \
                                      d_out_auc);                                                       \
    } while (0)
    if (p->pos16) {
        if (p->weighted) LAUNCH_FR(uint16_t, true); else LAUNCH_FR(uint16_t, false);
    } else {
        if (p->weighted) LAUNCH_FR(uint32_t, true); else LAUNCH_FR(uint32_t, false);
    }
#undef LAUNCH_FR
    g_launches.fetch_add(1);
    CK(cudaGetLastError());
    return AUCELL_OK;
}

// ---- host-buffer streaming pipeline ------------------------------------------------------------------
namespace {

struct Slot {
    cudaStream_t st = nullptr;
    int64_t* d_indptr = nullptr;
    int32_t* d_indices = nullptr;
    float* d_vals = nullptr;   // CSR data or dense rows
    double* d_auc = nullptr;
    int32_t* d_nnz = nullptr;
};

void free_slots(std::vector<Slot>& s) {
    for (auto& x : s) {
        if (x.st) cudaStreamSynchronize(x.st);
        cudaFree(x.d_indptr); cudaFree(x.d_indices); cudaFree(x.d_vals); cudaFree(x.d_auc); cudaFree(x.d_nnz);
        if (x.st) cudaStreamDestroy(x.st);
    }
    s.clear();
}

}  // namespace

int aucell_csr_f32_host(const aucell_plan_t* p, const int64_t* h_indptr, const int32_t* h_indices,
                        const float* h_data, int64_t n_cells, double* h_out_auc, int32_t* h_out_nnz,
                        int64_t chunk_cells) {
    if (!p) return fail(AUCELL_ERR_INVALID, "plan is NULL");
    if (p->n_regulons <= 0) return fail(AUCELL_ERR_INVALID, "plan has no regulons");
    if (n_cells < 0) return fail(AUCELL_ERR_INVALID, "n_cells < 0");
    if (n_cells == 0) return AUCELL_OK;
    if (!h_indptr || !h_out_auc) return fail(AUCELL_ERR_INVALID, "NULL pointer");
    DeviceGuard dg(p->device);
    if (!dg.ok) return fail(AUCELL_ERR_CUDA, "cudaSetDevice failed");
    const int64_t R = p->n_regulons;
    const int64_t total_nnz = h_indptr[n_cells] - h_indptr[0];
    if (total_nnz > 0 && (!h_indices || !h_data)) return fail(AUCELL_ERR_INVALID, "NULL pointer");
    if (chunk_cells <= 0) {
        // ~192 MB of (index, value) pairs per chunk
        const double per_cell = std::max<double>(1.0, (double)total_nnz / (double)n_cells);
        chunk_cells = (int64_t)std::max<double>(1024.0, (192.0 * 1024 * 1024) / (8.0 * per_cell));
    }
    chunk_cells = std::min(chunk_cells, n_cells);
    // chunk nnz upper bound
    int64_t max_nnz = 0;
    for (int64_t c0 = 0; c0 < n_cells; c0 += chunk_cells) {
        int64_t c1 = std::min(n_cells, c0 + chunk_cells);
        max_nnz = std::max(max_nnz, h_indptr[c1] - h_indptr[c0]);
    }
    const int n_slots = (int)std::min<int64_t>(3, (n_cells + chunk_cells - 1) / chunk_cells);
    std::vector<Slot> slots(n_slots);
    int rc = AUCELL_OK;
#define HCK(call)                                                      \
    do {                                                               \
        cudaError_t e_ = (call);                                       \
        if (e_ != cudaSuccess) {                                       \
            rc = fail_cuda(e_, #call, __LINE__);                       \
            free_slots(slots);                                         \
            return rc;                                                 \
        }                                                              \
    } while (0)
    for (auto& s : slots) {
        HCK(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
        HCK(cudaMalloc(&s.d_indptr, sizeof(int64_t) * (chunk_cells + 1)));
        HCK(cudaMalloc(&s.d_indices, sizeof(int32_t) * std::max<int64_t>(max_nnz, 1)));
        HCK(cudaMalloc(&s.d_vals, sizeof(float) * std::max<int64_t>(max_nnz, 1)));
        HCK(cudaMalloc(&s.d_auc, sizeof(double) * chunk_cells * R));
        if (h_out_nnz) HCK(cudaMalloc(&s.d_nnz, sizeof(int32_t) * chunk_cells));
    }
    int k = 0;
    for (int64_t c0 = 0; c0 < n_cells; c0 += chunk_cells, k = (k + 1) % n_slots) {
        Slot& s = slots[k];
        const int64_t c1 = std::min(n_cells, c0 + chunk_cells);
        const int64_t nc = c1 - c0;
        const int64_t e0 = h_indptr[c0], ne = h_indptr[c1] - e0;
        // stream order on s.st protects the slot's buffers from the previous chunk that used it
        HCK(cudaMemcpyAsync(s.d_indptr, h_indptr + c0, sizeof(int64_t) * (nc + 1), cudaMemcpyHostToDevice, s.st));
        if (ne > 0) {
            HCK(cudaMemcpyAsync(s.d_indices, h_indices + e0, sizeof(int32_t) * ne, cudaMemcpyHostToDevice, s.st));
            HCK(cudaMemcpyAsync(s.d_vals, h_data + e0, sizeof(float) * ne, cudaMemcpyHostToDevice, s.st));
        }
        // the kernel indexes indices/data with the absolute offsets of h_indptr: shift the bases
        rc = aucell_csr_f32(p, s.d_indptr, s.d_indices - e0, s.d_vals - e0, nc, s.d_auc, s.d_nnz, s.st);
        if (rc) { free_slots(slots); return rc; }
        HCK(cudaMemcpyAsync(h_out_auc + c0 * R, s.d_auc, sizeof(double) * nc * R, cudaMemcpyDeviceToHost, s.st));
        if (h_out_nnz)
            HCK(cudaMemcpyAsync(h_out_nnz + c0, s.d_nnz, sizeof(int32_t) * nc, cudaMemcpyDeviceToHost, s.st));
    }
    for (auto& s : slots) HCK(cudaStreamSynchronize(s.st));
    free_slots(slots);
    return AUCELL_OK;
}

int aucell_dense_f32_host(const aucell_plan_t* p, const float* h_expr, int64_t n_cells, int64_t row_stride,
                          double* h_out_auc, int32_t* h_out_nnz, int64_t chunk_cells) {
    if (!p) return fail(AUCELL_ERR_INVALID, "plan is NULL");
    if (p->n_regulons <= 0) return fail(AUCELL_ERR_INVALID, "plan has no regulons");
    if (n_cells < 0 || row_stride < p->n_genes) return fail(AUCELL_ERR_INVALID, "bad n_cells / row_stride");
    if (n_cells == 0) return AUCELL_OK;
    if (!h_expr || !h_out_auc) return fail(AUCELL_ERR_INVALID, "NULL pointer");
    DeviceGuard dg(p->device);
    if (!dg.ok) return fail(AUCELL_ERR_CUDA, "cudaSetDevice failed");
    const int64_t R = p->n_regulons;
    if (chunk_cells <= 0)
        chunk_cells = std::max<int64_t>(256, (int64_t)(192.0 * 1024 * 1024 / (4.0 * (double)row_stride)));
    chunk_cells = std::min(chunk_cells, n_cells);
    const int n_slots = (int)std::min<int64_t>(3, (n_cells + chunk_cells - 1) / chunk_cells);
    std::vector<Slot> slots(n_slots);
    int rc = AUCELL_OK;
    for (auto& s : slots) {
        HCK(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
        HCK(cudaMalloc(&s.d_vals, sizeof(float) * chunk_cells * row_stride));
        HCK(cudaMalloc(&s.d_auc, sizeof(double) * chunk_cells * R));
        if (h_out_nnz) HCK(cudaMalloc(&s.d_nnz, sizeof(int32_t) * chunk_cells));
    }
    int k = 0;
    for (int64_t c0 = 0; c0 < n_cells; c0 += chunk_cells, k = (k + 1) % n_slots) {
        Slot& s = slots[k];
        const int64_t nc = std::min(n_cells, c0 + chunk_cells) - c0;
        HCK(cudaMemcpyAsync(s.d_vals, h_expr + c0 * row_stride, sizeof(float) * nc * row_stride,
                            cudaMemcpyHostToDevice, s.st));
        rc = aucell_dense_f32(p, s.d_vals, nc, row_stride, s.d_auc, s.d_nnz, s.st);
        if (rc) { free_slots(slots); return rc; }
        HCK(cudaMemcpyAsync(h_out_auc + c0 * R, s.d_auc, sizeof(double) * nc * R, cudaMemcpyDeviceToHost, s.st));
        if (h_out_nnz)
            HCK(cudaMemcpyAsync(h_out_nnz + c0, s.d_nnz, sizeof(int32_t) * nc, cudaMemcpyDeviceToHost, s.st));
    }
    for (auto& s : slots) HCK(cudaStreamSynchronize(s.st));
    free_slots(slots);
    return AUCELL_OK;
}
#undef HCK

}  // extern "C"

// launch.cuh -- plan object, error helpers and the kernel launchers shared by the translation units of
// libaucell_b200.so: aucell_b200.cu (C-ABI, plan, host pipelines) and one unit per input kind / dtype
// (aucell_{f32,f64}_{dense,csr}.cu, aucell_f32_csr16.cu) so that the kernel instantiations compile in parallel.
#pragma once
#include "aucell_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <type_traits>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "fused_kernel.cuh"
#include "tie_kernel.cuh"

namespace aucell_host {

using namespace aucell;

// defined once, in aucell_b200.cu
std::string& last_error();                 // thread-local message of the last failure
std::atomic<int64_t>& launch_counter();

inline int fail(int code, const std::string& msg) {
    last_error() = msg;
    return code;
}

inline int fail_cuda(cudaError_t e, const char* what, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) at libaucell_b200 line %d: %s", (int)e, cudaGetErrorString(e), line, what);
    last_error() = buf;
    return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? AUCELL_ERR_NO_DEVICE : AUCELL_ERR_CUDA;
}

#define CK(call)                                                          \
    do {                                                                  \
        cudaError_t e_ = (call);                                          \
        if (e_ != cudaSuccess) return fail_cuda(e_, #call, __LINE__);     \
    } while (0)

constexpr int64_t kMaxGenes = 1 << 20;  // bitmap words / int positions; far above any gene annotation

inline int next_pow2(int64_t v) {
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

}  // namespace aucell_host

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
    bool acc64 = false;       // 64-bit accumulators (weighted, or cutoff too large for exact u32 sums)
    int n_acc = 0;
    // device arrays
    int32_t* d_inv_perm = nullptr;
    uint16_t* d_inv_perm16 = nullptr;       // the same as uint16 (pos16 plans)
    uint32_t* d_tptr = nullptr;
    unsigned long long* d_tent = nullptr;   // weighted: (W << 16) | accumulator
    uint16_t* d_tacc16 = nullptr;           // unweighted: accumulator
    int32_t* d_acc_first = nullptr;
    uint8_t* d_acc_parts = nullptr;
    double* d_acc_scale = nullptr;
    double* d_reg_max_auc = nullptr;
    unsigned long long* d_stats = nullptr;   // path statistics of the fused kernel (5 counters)
    // tie-class kernel (tie_kernel.cuh): ELL regulon table, u32 limb accumulators.  Built when the plan qualifies
    // (16-bit positions, non-negative weights of bounded range); aucell_fused_kernel then takes only the to-do list.
    bool fast_ok = false;
    bool fast_enabled = true;                // aucell_plan_set_fast_path
    int tie_L = 0;                           // weighted: products are split at this bit
    int tie_slots = 4;                       // weighted: direct slots per table row (3: the fourth only holds the chain)
    uint4* d_ell_w = nullptr;
    uint4* d_ell_w16 = nullptr;              // weighted, compact rows (accumulator indices < 1024): see TieArgs::ell_w16
    uint2* d_xinfo = nullptr;
    uint2* d_xw = nullptr;
    uint2* d_ell_u = nullptr;
    uint16_t* d_xu = nullptr;
    unsigned long long* d_tie_stats = nullptr;   // [TIE_ST_N]
    void* d_reg_rec = nullptr;                   // [R] 32-byte epilogue records of aucell_tie_kernel
    // one staging slot of the host-buffer pipeline (the slots belong to the device and outlive the plans: aucell_b200.cu)
    struct HostSlot {
        cudaStream_t st = nullptr;
        char* buf = nullptr;       // one allocation: indptr | indices | values | auc | nnz
        size_t bytes = 0;
        uint16_t* pin_idx = nullptr;   // pinned host staging of a chunk's column indices narrowed to uint16
        size_t pin_bytes = 0;
        uint8_t* pin_codes = nullptr;  // pinned host staging of a chunk's values as one-byte codes, then the 1 KB table
        size_t pin_codes_bytes = 0;
        char* pin_out = nullptr;       // pageable callers: pinned landing zone of a chunk's AUC block (+ nnz), copied out
        size_t pin_out_bytes = 0;      // by host threads when the slot comes round again
        int64_t out_c0 = 0, out_nc = 0;   // the chunk waiting in pin_out (out_nc == 0: none)
        cudaEvent_t idx_done = nullptr;   // the upload out of pin_idx has finished
        bool idx_pending = false;
    };
    mutable std::mutex host_mu;
    mutable int64_t last_h2d_bytes = 0, last_d2h_bytes = 0;   // what the last host-buffer call moved over PCIe
    mutable bool hint_idx16 = true, hint_val8 = true;         // upload format the last host-buffer call settled on
    // overflow scratch lists, one per stream that ever needed one
    mutable std::mutex mu;
    mutable std::map<std::pair<cudaStream_t, int>, std::pair<void*, size_t>> scratch;
};

namespace aucell_host {

inline GeneTableDev tab_of(const aucell_plan* p) {
    GeneTableDev r;
    r.n_regulons = p->n_regulons;
    r.n_acc = p->n_acc;
    r.tptr = p->d_tptr;
    r.tent = p->d_tent;
    r.tacc16 = p->d_tacc16;
    r.acc_first = p->d_acc_first;
    r.acc_parts = p->d_acc_parts;
    r.acc_scale = p->d_acc_scale;
    r.max_auc = p->d_reg_max_auc;
    return r;
}

inline int align16(int v) { return (v + 15) & ~15; }

inline int get_scratch(const aucell_plan* p, cudaStream_t st, int key_bytes, int grid, int scratch_cap, void** key,
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

// to-do list of the tie-class kernel: [0] counter (int32), list from byte 16 on; one per stream, grown on demand
inline int get_todo(const aucell_plan* p, cudaStream_t st, int64_t n_cells, int32_t** todo, int32_t** todo_n) {
    const size_t need = 16 + 2 * sizeof(int32_t) * (size_t)n_cells;      // two lists: general kernel, long rows
    std::lock_guard<std::mutex> lk(p->mu);
    auto& slot = p->scratch[{st, -1}];
    if (slot.second < need) {
        if (slot.first) cudaFree(slot.first);
        slot = {nullptr, 0};
        void* ptr = nullptr;
        cudaError_t e = cudaMalloc(&ptr, need);
        if (e != cudaSuccess) return fail_cuda(e, "cudaMalloc(todo)", __LINE__);
        slot = {ptr, need};
    }
    *todo_n = reinterpret_cast<int32_t*>(slot.first);          // [0] counter, [1..2] value range (tie_range_kernel),
                                                               // [3] counter of the long-row list (behind the first list)
    *todo = reinterpret_cast<int32_t*>((char*)slot.first + 16);
    return AUCELL_OK;
}

// ---- tie-class kernel -----------------------------------------------------------------------------------------
struct TieLayout { int group_stride, groups, smem; };

// shared-memory carve-up of aucell_tie_kernel; fills the offsets of `a`
inline TieLayout tie_layout(const aucell_plan* p, TieArgs& a, int cap) {
    const int ip_bytes = align16(((p->n_genes + 7) / 8) * 16);
    // fixed-size arrays first, at the compile-time offsets the kernel uses (tie_kernel.cuh: kTieO*)
    a.o_misc = kTieOMisc; a.o_rep = kTieORep; a.o_code = kTieOCode; a.o_cnt = kTieOCnt; a.o_fix = kTieOFix;
    int off = kTieOA;
    a.o_A = off; off = align16(off + ((p->n_words + 3) / 4) * 16 + 128);    // + 32 spare words invalid lanes OR into
    a.o_wpre = off; off = align16(off + p->n_words * 2);
    a.o_bpp = off; off = align16(off + (cap + 1) * 2);                      // + the spare slot
    a.o_bpc = off; off = align16(off + cap + 1);
    a.o_acc = off; off = align16(off + a.acc_words * 4);
    TieLayout L;
    L.group_stride = off;
    const int64_t room = p->smem_optin - ip_bytes;
    L.groups = (int)std::max<int64_t>(0, std::min<int64_t>(kTieMaxGroups, room / off));
    L.smem = ip_bytes + L.groups * off;
    a.off_group0 = ip_bytes;
    a.group_stride = off;
    a.groups = L.groups;
    a.cap = cap;
    return L;
}

// shared-memory carve-up of aucell_tie2_kernel (class bitmaps)
inline TieLayout tie2_layout(const aucell_plan* p, TieArgs& a, int cap) {
    const int ip_bytes = align16(((p->n_genes + 7) / 8) * 16);
    a.nsw = (p->n_words + 3) / 4;
    a.spt = (a.nsw + kTG - 1) / kTG;
    int off = 0;
    a.o_D = off; off = align16(off + ((kT2C + 1) * a.nsw + 8) * 16);        // + 32 spare words invalid lanes OR into
    a.o_pre = off; off = align16(off + a.nsw * 16);
    a.o_rep = off; off = align16(off + (kTieNB + 4) * 4);
    a.o_code = off; off = align16(off + kTieNB);
    a.o_fix = off; off = align16(off + kT2Fix * 8);
    a.o_acc = off; off = align16(off + a.acc_words * 4);
    a.o_misc = off; off = align16(off + kTieMiscWords * 4);
    TieLayout L;
    L.group_stride = off;
    const int64_t room = p->smem_optin - ip_bytes;
    L.groups = (int)std::max<int64_t>(0, std::min<int64_t>(kTieMaxGroups, room / off));
    L.smem = ip_bytes + L.groups * off;
    a.off_group0 = ip_bytes;
    a.group_stride = off;
    a.groups = L.groups;
    a.cap = cap;
    return L;
}

// which tie-class kernel: 1 = position-ordered list (aucell_tie_kernel), 2 = class bitmaps (aucell_tie2_kernel),
// default: the first, with the rows beyond its capacity handed to the second
inline int tie_version() {
    const char* env = getenv("AUCELL_B200_TIE_KERNEL");
    if (env && env[0] == '1' && env[1] == 0) return 1;
    if (env && env[0] == '2' && env[1] == 0) return 2;
    return 12;
}

inline int tie2_cap() {
    const char* env = getenv("AUCELL_B200_TIE2_CAP");
    int cap = env && *env ? atoi(env) : 16384;         // (16-bit prefixes: < 65536)
    return std::max(256, std::min(32768, cap));
}

inline int tie_cap() {
    const char* env = getenv("AUCELL_B200_TIE_CAP");
    int cap = env && *env ? atoi(env) : kTieMaxCap;
    return std::max(256, std::min(kTieMaxCap, cap)) & ~15;
}

// Launch aucell_tie_kernel over all cells; cells it leaves are listed in *todo (count in **todo_n, device memory).
template <int IN>
int launch_tie(const aucell_plan* p, const int64_t* indptr, const void* indices, const void* data, int64_t n_cells,
               double* out_auc, int32_t* out_nnz, cudaStream_t st, int32_t** todo, int32_t** todo_n, bool* launched,
               const uint32_t* lut = nullptr) {
    *launched = false;
    TieArgs a;
    memset(&a, 0, sizeof a);
    a.acc_words = p->weighted ? 2 * p->n_acc + 64 : p->n_acc + 32;       // + 32 dummy accumulators (empty table slots)
    const int version = tie_version();              // 1: aucell_tie_kernel, 2: aucell_tie2_kernel, 12: both (long rows to the second)
    TieArgs a2;
    memcpy(&a2, &a, sizeof a);
    const TieLayout L1 = tie_layout(p, a, tie_cap());
    const TieLayout L2 = tie2_layout(p, a2, tie2_cap());
    // (the second kernel behind the first only when at least three of its cell groups fit an SM: with fewer -- more than
    // ~40,000 genes -- aucell_fused_kernel ranks the long rows faster)
    const bool use1 = version != 2 && L1.groups >= 1;
    const bool use2 = version != 1 && L2.groups >= (version == 2 || !use1 ? 1 : 3);
    if (!use1 && !use2) return AUCELL_OK;          // does not fit: the caller runs aucell_fused_kernel on everything
    int rc = get_todo(p, st, n_cells, todo, todo_n);
    if (rc) return rc;
    CK(cudaMemsetAsync(*todo_n, 0, 4 * sizeof(int32_t), st));
    for (TieArgs* x : {&a, &a2}) {
        x->indptr = indptr; x->indices = indices; x->data = data; x->lut = lut;
        x->n_cells = n_cells; x->n_genes = p->n_genes; x->n_words = p->n_words;
        x->wpt = ((p->n_words + kTG - 1) / kTG) | 1;      // words per thread in P2: odd, so that a warp's reads spread over the banks
        x->cutoff = (int)p->rank_cutoff;
        x->inv_perm = p->d_inv_perm16;
        x->ell_w = p->d_ell_w; x->xw = p->d_xw; x->ell_u = p->d_ell_u; x->xu = p->d_xu;
        x->ell_w16 = p->d_ell_w16; x->xinfo = p->d_xinfo; x->row16 = p->d_ell_w16 != nullptr ? 1 : 0;
        x->L = p->tie_L;
        x->slots4 = p->tie_slots == 4 ? 1 : 0;
        x->n_regulons = p->n_regulons;
        x->acc_first = p->d_acc_first; x->acc_parts = p->d_acc_parts; x->acc_scale = p->d_acc_scale;
        x->reg_rec = p->d_reg_rec;
        x->range = reinterpret_cast<const uint32_t*>(*todo_n) + 1;
        x->max_auc = p->d_reg_max_auc;
        x->out_auc = out_auc; x->out_nnz = out_nnz;
        x->todo = *todo; x->todo_n = *todo_n;
        x->stats = p->d_tie_stats;
    }
    {
        const char* envp = getenv("AUCELL_B200_TIE_PREFETCH");
        a.prefetch = (envp && envp[0] == '0') ? 0 : 1;
        const char* envk = getenv("AUCELL_B200_TIE_SKIP");
        a.skip = envk ? atoi(envk) : 0;
    }
    if (use1 && use2) {                            // rows beyond the first kernel's cap: a list for the second
        a.long_list = *todo + n_cells;
        a.long_n = *todo_n + 3;
        a.long_cap = a2.cap;
        a2.in_list = a.long_list;
        a2.in_n = a.long_n;
    }
    tie_range_kernel<IN><<<1, 1024, 0, st>>>(indptr, n_cells, data, lut, reinterpret_cast<uint32_t*>(*todo_n) + 1);
    launch_counter().fetch_add(1);
#define LAUNCH_TIE(KERNEL, ARGS, LAY, WORK)                                                       \
    do {                                                                                          \
        auto k = KERNEL;                                                                          \
        const int grid = (int)std::min<int64_t>(p->sm_count, ((WORK) + LAY.groups - 1) / LAY.groups); \
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, LAY.smem));       \
        k<<<grid, kTG * LAY.groups, LAY.smem, st>>>(ARGS);                                        \
        launch_counter().fetch_add(1);                                                            \
    } while (0)
    if (use1) {
        if (p->weighted) LAUNCH_TIE((aucell_tie_kernel<IN, false>), a, L1, n_cells);
        else LAUNCH_TIE((aucell_tie_kernel<IN, true>), a, L1, n_cells);
    }
    if (use2) {
        // (behind the first kernel the number of long rows is only known on the device: a full grid walks the list)
        const int64_t work2 = n_cells;
        if (p->weighted) LAUNCH_TIE((aucell_tie2_kernel<IN, false>), a2, L2, work2);
        else LAUNCH_TIE((aucell_tie2_kernel<IN, true>), a2, L2, work2);
    }
#undef LAUNCH_TIE
    launch_counter().fetch_add(1);
    CK(cudaGetLastError());
    *launched = true;
    return AUCELL_OK;
}

// ---- fused AUCell kernel ------------------------------------------------------------------------------------
// resident CTAs per SM of a kernel (registers, shared memory, threads): the persistent grid is a multiple of it
template <typename K>
int resident_ctas(K kernel, int threads, int smem) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem) != cudaSuccess || n < 1) n = 1;
    return n;
}

template <typename ValT, typename PosT, int IN>
int launch_fused(const aucell_plan* p, FusedArgs<ValT>& a, int smem, int64_t n_cells, cudaStream_t st, int key_bytes) {
    int grid = 0;
#define LAUNCH_FUSED(KERNEL)                                                                              \
    do {                                                                                                  \
        auto k = KERNEL;                                                                                  \
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));                   \
        grid = (int)std::min<int64_t>(n_cells, (int64_t)p->sm_count * resident_ctas(k, kF, smem));        \
        int rc = get_scratch(p, st, key_bytes, grid, a.scratch_cap, &a.scratch_key, &a.scratch_pos);      \
        if (rc) return rc;                                                                                \
        k<<<grid, kF, smem, st>>>(a);                                                                     \
    } while (0)
    if (p->acc64) LAUNCH_FUSED((aucell_fused_kernel<ValT, PosT, IN, true>));
    else LAUNCH_FUSED((aucell_fused_kernel<ValT, PosT, IN, false>));
#undef LAUNCH_FUSED
    launch_counter().fetch_add(1);
    CK(cudaGetLastError());
    return AUCELL_OK;
}

// pre_todo / pre_todo_n: a to-do list some other launch has already prepared on this stream (the byte-coded entry point:
// its tie-class kernel reads codes, this kernel the decoded values); the tie-class kernel is then not launched here.
template <typename ValT, int IN>
int run_fused(const aucell_plan* p, const ValT* dense, int64_t row_stride, const int64_t* indptr,
              const int32_t* indices, const ValT* data, int64_t n_cells, double* out_auc, int32_t* out_nnz,
              void* stream, int32_t* pre_todo = nullptr, int32_t* pre_todo_n = nullptr, bool allow_tie = true) {
    if (!p) return fail(AUCELL_ERR_INVALID, "plan is NULL");
    if (n_cells < 0) return fail(AUCELL_ERR_INVALID, "n_cells < 0");
    if (p->n_regulons <= 0) return fail(AUCELL_ERR_INVALID, "plan has no regulons");
    if (n_cells == 0) return AUCELL_OK;
    if (IN == IN_DENSE && (!dense || row_stride < p->n_genes))
        return fail(AUCELL_ERR_INVALID, "dense input: NULL pointer or row_stride < n_genes");
    if (IN != IN_DENSE && !indptr) return fail(AUCELL_ERR_INVALID, "CSR input: NULL indptr");
    if (IN == IN_CSR16 && !p->pos16) return fail(AUCELL_ERR_INVALID, "uint16 column indices need n_genes <= 65535");
    if (!out_auc) return fail(AUCELL_ERR_INVALID, "out_auc is NULL");
    // AUCell proper has rank_cutoff < n_genes; larger cutoffs (cisTarget on column-selected rankings) exist only for
    // aucell_from_rankings_u32.  Here they would truncate the 16-bit multipliers of the scatter queue.
    if (p->rank_cutoff > p->n_genes)
        return fail(AUCELL_ERR_INVALID, "rank_cutoff > n_genes: only aucell_from_rankings_u32 accepts such plans");
    DeviceGuard dg(p->device);
    if (!dg.ok) return fail(AUCELL_ERR_CUDA, "cudaSetDevice failed");
    cudaStream_t st = (cudaStream_t)stream;
    const int key_bytes = (int)sizeof(typename KeyOf<ValT>::type);
    const int pos_bytes = p->pos16 ? 2 : 4;

    FusedArgs<ValT> a;
    memset(&a, 0, sizeof a);
    // compile-time offsets up to the accumulators (fused_offsets), then the zero-fill bitmap + its prefixes
    const FusedOffsets L = fused_offsets(key_bytes, pos_bytes);
    a.off_h1 = L.h1; a.off_h2 = L.h2; a.off_fk = L.fk; a.off_fp = L.fp; a.off_mixed = L.mixed;
    a.off_acc = L.acc;
    int off = align16(L.acc + p->n_acc * 8);
    a.off_bits = off; off = align16(off + p->n_words * 4);
    a.off_wpre = off; off = align16(off + p->n_words * 4);
    {
        int g = 2;
        while ((int64_t)g * 2 * (key_bytes + pos_bytes) <= a.off_mixed - a.off_h1) g <<= 1;
        a.gen_cap = g;
    }
    const int smem = off;
    if (smem > p->smem_optin)
        return fail(AUCELL_ERR_UNSUPPORTED, "n_genes / n_regulons too large for the shared-memory kernels");

    a.dense = dense; a.row_stride = row_stride;
    a.indptr = indptr; a.indices = indices; a.data = data;
    a.n_cells = n_cells; a.n_genes = p->n_genes; a.n_words = p->n_words;
    a.inv_perm = p->pos16 ? (const void*)p->d_inv_perm16 : (const void*)p->d_inv_perm;
    a.rank_cutoff = (int)p->rank_cutoff;
    a.inv_ng = (uint32_t)std::min<uint64_t>(0xFFFFFFFFull, ((1ull << 32) + (uint64_t)p->n_genes - 1) / (uint64_t)p->n_genes);
    a.tab = tab_of(p);
    a.out_auc = out_auc; a.out_nnz = out_nnz;
    a.stats = p->d_stats;
    a.scratch_cap = next_pow2(p->n_genes);
    // Count-like cells go through the tie-class kernel; what it leaves (long rows, negative / NaN values, continuous
    // values) comes back as a to-do list for the kernel below.
    if (pre_todo != nullptr) {
        a.todo = pre_todo;
        a.todo_n = pre_todo_n;
    } else if constexpr (IN != IN_DENSE && std::is_same<ValT, float>::value) {
        if (allow_tie && p->fast_ok && p->fast_enabled && n_cells < (int64_t)INT32_MAX) {
            bool launched = false;
            int32_t *todo = nullptr, *todo_n = nullptr;
            int rc = launch_tie<IN>(p, indptr, indices, data, n_cells, out_auc, out_nnz, st, &todo, &todo_n, &launched);
            if (rc) return rc;
            if (launched) { a.todo = todo; a.todo_n = todo_n; }
        }
    }
    if constexpr (IN == IN_CSR16) {       // 16-bit indices imply 16-bit positions: one instantiation
        return launch_fused<ValT, uint16_t, IN>(p, a, smem, n_cells, st, key_bytes);
    } else {
        if (p->pos16) return launch_fused<ValT, uint16_t, IN>(p, a, smem, n_cells, st, key_bytes);
        return launch_fused<ValT, uint32_t, IN>(p, a, smem, n_cells, st, key_bytes);
    }
}

// CSR with one-byte value codes (a table of up to 256 float values) and uint16 (IN_TIE = IN_CSR16_U8) or int32
// (IN_CSR_U8) column indices: the tie-class kernel reads the codes; the values are decoded into `scratch_vals` (indexed
// like the codes) for the cells it hands over.
template <int IN_TIE>
int run_fused_lut8(const aucell_plan* p, const int64_t* indptr, const void* indices, const uint8_t* codes,
                   const float* lut, float* scratch_vals, int64_t n_cells, double* out_auc, int32_t* out_nnz,
                   void* stream) {
    constexpr int IN_GEN = IN_TIE == IN_CSR16_U8 ? IN_CSR16 : IN_CSR;
    if (!p) return fail(AUCELL_ERR_INVALID, "plan is NULL");
    if (n_cells < 0) return fail(AUCELL_ERR_INVALID, "n_cells < 0");
    if (p->n_regulons <= 0) return fail(AUCELL_ERR_INVALID, "plan has no regulons");
    if (n_cells == 0) return AUCELL_OK;
    if (!indptr || !lut || !scratch_vals || !out_auc) return fail(AUCELL_ERR_INVALID, "NULL pointer");
    if (IN_TIE == IN_CSR16_U8 && !p->pos16) return fail(AUCELL_ERR_INVALID, "uint16 column indices need n_genes <= 65535");
    DeviceGuard dg(p->device);
    if (!dg.ok) return fail(AUCELL_ERR_CUDA, "cudaSetDevice failed");
    cudaStream_t st = (cudaStream_t)stream;
    int32_t *todo = nullptr, *todo_n = nullptr;
    bool launched = false;
    if (p->fast_ok && p->fast_enabled && n_cells < (int64_t)INT32_MAX && p->rank_cutoff <= p->n_genes) {
        int rc = launch_tie<IN_TIE>(p, indptr, indices, codes, n_cells, out_auc, out_nnz, st, &todo, &todo_n, &launched,
                                    reinterpret_cast<const uint32_t*>(lut));
        if (rc) return rc;
    }
    tie_decode_kernel<IN_TIE><<<p->sm_count * 8, 256, 0, st>>>(codes, reinterpret_cast<const uint32_t*>(lut), scratch_vals, indptr, n_cells);
    launch_counter().fetch_add(1);
    CK(cudaGetLastError());
    return run_fused<float, IN_GEN>(p, nullptr, 0, indptr, reinterpret_cast<const int32_t*>(indices), scratch_vals, n_cells,
                                    out_auc, out_nnz, stream, launched ? todo : nullptr, launched ? todo_n : nullptr, false);
}

// create_rankings through the fused kernel in RANKS mode (bucket-sort ranking of every gene).
template <typename ValT, int IN>
int run_rank_fused(const aucell_plan* p, const ValT* dense, int64_t row_stride, const int64_t* indptr,
                   const int32_t* indices, const ValT* data, int64_t n_cells, uint32_t* out_rank, int32_t* out_nnz,
                   void* stream) {
    DeviceGuard dg(p->device);
    if (!dg.ok) return fail(AUCELL_ERR_CUDA, "cudaSetDevice failed");
    cudaStream_t st = (cudaStream_t)stream;
    const int key_bytes = (int)sizeof(typename KeyOf<ValT>::type);
    const int pos_bytes = p->pos16 ? 2 : 4;
    FusedArgs<ValT> a;
    memset(&a, 0, sizeof a);
    const FusedOffsets L = fused_offsets(key_bytes, pos_bytes);
    a.off_h1 = L.h1; a.off_h2 = L.h2; a.off_fk = L.fk; a.off_fp = L.fp; a.off_mixed = L.mixed;
    a.off_acc = L.acc;                    // no accumulators in RANKS mode
    int off = L.acc;
    {
        int g = 2;
        while ((int64_t)g * 2 * (key_bytes + pos_bytes) <= a.off_mixed - a.off_h1) g <<= 1;
        a.gen_cap = g;
    }
    a.off_bits = off; off = align16(off + p->n_words * 4);
    a.off_wpre = off; off = align16(off + p->n_words * 4);
    const int smem = off;
    if (smem > p->smem_optin)
        return fail(AUCELL_ERR_UNSUPPORTED, "n_genes too large for the shared-memory kernels");
    a.dense = dense; a.row_stride = row_stride;
    a.indptr = indptr; a.indices = indices; a.data = data;
    a.n_cells = n_cells; a.n_genes = p->n_genes; a.n_words = p->n_words;
    a.inv_perm = p->pos16 ? (const void*)p->d_inv_perm16 : (const void*)p->d_inv_perm;
    a.rank_cutoff = p->n_genes;
    a.inv_ng = (uint32_t)std::min<uint64_t>(0xFFFFFFFFull, ((1ull << 32) + (uint64_t)p->n_genes - 1) / (uint64_t)p->n_genes);
    memset(&a.tab, 0, sizeof a.tab);
    a.out_rank = out_rank; a.out_nnz = out_nnz;
    a.stats = nullptr;
    a.scratch_cap = next_pow2(p->n_genes);
    int grid = 0;
#define LAUNCH_RANK(PosT)                                                                                  \
    do {                                                                                                   \
        auto k = aucell_fused_kernel<ValT, PosT, IN, false, true>;                                         \
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));                    \
        grid = (int)std::min<int64_t>(n_cells, (int64_t)p->sm_count * resident_ctas(k, kF, smem));         \
        int rc = get_scratch(p, st, key_bytes, grid, a.scratch_cap, &a.scratch_key, &a.scratch_pos);       \
        if (rc) return rc;                                                                                 \
        k<<<grid, kF, smem, st>>>(a);                                                                      \
    } while (0)
    if (p->pos16) LAUNCH_RANK(uint16_t); else LAUNCH_RANK(uint32_t);
#undef LAUNCH_RANK
    launch_counter().fetch_add(1);
    CK(cudaGetLastError());
    return AUCELL_OK;
}


template <typename ValT, int IN>
int run_rank(const aucell_plan* p, const ValT* dense, int64_t row_stride, const int64_t* indptr,
             const int32_t* indices, const ValT* data, int64_t n_cells, uint32_t* out_rank, int32_t* out_nnz,
             void* stream) {
    if (!p) return fail(AUCELL_ERR_INVALID, "plan is NULL");
    if (n_cells < 0) return fail(AUCELL_ERR_INVALID, "n_cells < 0");
    if (n_cells == 0) return AUCELL_OK;
    if (IN == IN_DENSE && (!dense || row_stride < p->n_genes))
        return fail(AUCELL_ERR_INVALID, "dense input: NULL pointer or row_stride < n_genes");
    if (IN == IN_CSR && !indptr) return fail(AUCELL_ERR_INVALID, "CSR input: NULL indptr");
    if (!out_rank) return fail(AUCELL_ERR_INVALID, "out_rank is NULL");
    return run_rank_fused<ValT, IN>(p, dense, row_stride, indptr, indices, data, n_cells, out_rank, out_nnz, stream);
}

}  // namespace aucell_host

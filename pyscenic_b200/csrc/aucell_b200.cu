// aucell_b200.cu -- C-ABI of libaucell_b200.so (see include/aucell_b200.h): plan management, aucell4r and the host-buffer
// streaming pipeline.  The ranking / fused kernels are instantiated in aucell_{f32,f64}_{dense,csr}.cu.
// sm_100a only; no CPU fallback.
#include "launch.cuh"

#include <chrono>
#include <condition_variable>
#include <functional>
#include <memory>
#include <thread>
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#include <sys/mman.h>
#endif

using namespace aucell;
using namespace aucell_host;

namespace aucell_host {
std::string& last_error() {
    thread_local std::string err;
    return err;
}
std::atomic<int64_t>& launch_counter() {
    static std::atomic<int64_t> n{0};
    return n;
}
// kernel launchers, one translation unit per (value type, input kind)
#define DECL_DENSE(T, SUF)                                                                                         \
    int rank_dense_##SUF(const aucell_plan* p, const T* d, int64_t n_cells, int64_t row_stride, uint32_t* out_rank,  \
                         int32_t* out_nnz, void* stream);                                                          \
    int fused_dense_##SUF(const aucell_plan* p, const T* d, int64_t n_cells, int64_t row_stride, double* out_auc,    \
                          int32_t* out_nnz, void* stream);
#define DECL_CSR(T, SUF)                                                                                           \
    int rank_csr_##SUF(const aucell_plan* p, const int64_t* indptr, const int32_t* indices, const T* data,         \
                       int64_t n_cells, uint32_t* out_rank, int32_t* out_nnz, void* stream);                       \
    int fused_csr_##SUF(const aucell_plan* p, const int64_t* indptr, const int32_t* indices, const T* data,        \
                        int64_t n_cells, double* out_auc, int32_t* out_nnz, void* stream);
DECL_DENSE(float, f32)
DECL_DENSE(double, f64)
DECL_CSR(float, f32)
DECL_CSR(double, f64)
#undef DECL_DENSE
#undef DECL_CSR
int fused_csr16_f32(const aucell_plan* p, const int64_t* indptr, const uint16_t* indices, const float* data,
                    int64_t n_cells, double* out_auc, int32_t* out_nnz, void* stream);
int fused_csr16_lut8(const aucell_plan* p, const int64_t* indptr, const uint16_t* indices, const uint8_t* codes,
                     const float* lut, float* scratch_vals, int64_t n_cells, double* out_auc, int32_t* out_nnz,
                     void* stream);
int fused_csr_lut8(const aucell_plan* p, const int64_t* indptr, const int32_t* indices, const uint8_t* codes,
                   const float* lut, float* scratch_vals, int64_t n_cells, double* out_auc, int32_t* out_nnz,
                   void* stream);
}  // namespace aucell_host

// ------------------------------------------------------------------------------------------------------
extern "C" {

int aucell_b200_abi_version(void) { return AUCELL_B200_ABI_VERSION; }

const char* aucell_b200_last_error(void) { return last_error().c_str(); }

int64_t aucell_b200_launch_count(void) { return launch_counter().load(); }

int64_t aucell_b200_max_genes(void) {
    // per cell the kernels keep two n_genes-bit arrays (occupancy bitmap + its word prefix) in shared memory next to
    // ~72 KB of sort state; positions are 20-bit.  With the 227 KB a CTA may use on sm_100:
    const int64_t words = ((int64_t)227 * 1024 - (int64_t)80 * 1024) / 8;
    return std::min<int64_t>(kMaxGenes, words * 32);
}

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
        // AUCell has rank_cutoff < n_genes (ctxcore derive_rank_cutoff); ctxcore.recovery.aucs on a column-selected
        // ranking matrix (cisTarget) derives the cutoff from the whole genome, so larger values are legal here
        if (rank_cutoff < 0 || rank_cutoff > (int64_t)1 << 30)
            return fail(AUCELL_ERR_INVALID, "rank_cutoff must satisfy 0 <= rank_cutoff <= 2^30");
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

    PCK(cudaMalloc(&p->d_stats, 8 * sizeof(unsigned long long)));
    PCK(cudaMemset(p->d_stats, 0, 8 * sizeof(unsigned long long)));
    PCK(cudaMalloc(&p->d_inv_perm, sizeof(int32_t) * n_genes));
    PCK(cudaMemcpy(p->d_inv_perm, inv.data(), sizeof(int32_t) * n_genes, cudaMemcpyHostToDevice));
    if (p->pos16) {
        const size_t n16 = ((size_t)n_genes + 7) & ~(size_t)7;      // aucell_tie_kernel copies it in 128-bit vectors
        std::vector<uint16_t> inv16(n16, 0);
        for (int64_t g = 0; g < n_genes; ++g) inv16[g] = (uint16_t)inv[g];
        PCK(cudaMalloc(&p->d_inv_perm16, sizeof(uint16_t) * n16));
        PCK(cudaMemcpy(p->d_inv_perm16, inv16.data(), sizeof(uint16_t) * n16, cudaMemcpyHostToDevice));
    }

    if (n_regulons > 0) {
        const int64_t nnz = h_reg_indptr[n_regulons];
        p->reg_nnz = nnz;
        p->acc64 = p->weighted || rank_cutoff > 65535;
        // ---- accumulators and fixed-point weights (see fused_kernel.cuh SCATTER, tie_kernel.cuh P6) ------------
        // Two quantisations.  "tie": what aucell_tie_kernel can accumulate in pairs of u32 limbs -- non-negative
        // weights, W < 2^Bt per part, at most two parts per regulon, the smallest weight keeping >= 22 significant bits
        // (relative error <= 2.4e-7 on every single term; the sum of non-negative terms inherits it).  Both kernels
        // then add the SAME integers, so a cell gets the same bits whichever kernel ranks it.  "wide": 48-bit signed
        // weights, up to four parts, >= 24 bits (aucell_fused_kernel only) for everything else.
        std::vector<int32_t> acc_first;
        std::vector<uint8_t> acc_parts;
        std::vector<double> acc_scale;
        struct TEnt { int32_t pos; uint32_t acc; long long w; };
        std::vector<TEnt> ents;
        int Bt = 0, tieL = 0;
        bool tie_q = p->pos16 && rank_cutoff >= 1 && rank_cutoff <= 65535 && n_genes >= 1;
        {
            const char* env = getenv("AUCELL_B200_TIE");
            if (env && env[0] == '0') tie_q = false;
        }
        if (tie_q && p->weighted) {
            int64_t max_hits = 1;
            for (int r = 0; r < n_regulons; ++r)
                if (h_reg_valid[r])
                    max_hits = std::max<int64_t>(max_hits, h_reg_indptr[r + 1] - h_reg_indptr[r]);   // (a gene listed twice counts twice)
            int nrb = 0;
            while (((int64_t)1 << nrb) < max_hits) ++nrb;          // hits <= 2^nrb
            int c = 0;
            while (((int64_t)1 << c) <= rank_cutoff) ++c;          // multipliers cutoff - rank < 2^c
            tieL = 32 - nrb;                                       // sum of hits low limbs (< 2^L each) fits 32 bits
            Bt = std::min(31, 63 - c - 2 * nrb);                   // ... and so does the sum of the high limbs
            if (Bt < 24 || tieL < 1 || tieL > 31) tie_q = false;
            for (int64_t k = 0; tie_q && k < nnz; ++k)
                if (!(h_reg_weights[k] >= 0.0) || !std::isfinite(h_reg_weights[k])) tie_q = false;
        }
        auto quantise = [&](bool tie) -> int {
            acc_first.assign((size_t)n_regulons, -1);
            acc_parts.assign((size_t)n_regulons, 0);
            acc_scale.clear();
            ents.clear();
            ents.reserve((size_t)nnz + 16);
            for (int r = 0; r < n_regulons; ++r) {
                const int64_t b = h_reg_indptr[r], e = h_reg_indptr[r + 1];
                for (int64_t k = b; k < e; ++k)
                    if (h_reg_cols[k] < 0 || h_reg_cols[k] >= n_genes)
                        return fail(AUCELL_ERR_INVALID, "regulon gene column out of range");
                if (!h_reg_valid[r]) continue;
                if (!(h_reg_max_auc[r] > 0.0) || e == b)
                    return fail(AUCELL_ERR_INVALID, "valid regulons need max_auc > 0 and at least one gene");
                const int first = (int)acc_scale.size();
                acc_first[r] = first;
                if (!p->weighted) {
                    acc_parts[r] = 1;
                    acc_scale.push_back(1.0);
                    for (int64_t k = b; k < e; ++k) ents.push_back({inv[h_reg_cols[k]], (uint32_t)first, 1ll});
                    continue;
                }
                double wmax = 0.0, wmin = 0.0;
                for (int64_t k = b; k < e; ++k) {
                    const double aw = std::fabs(h_reg_weights[k]);
                    if (!std::isfinite(aw)) return fail(AUCELL_ERR_INVALID, "regulon weights must be finite");
                    if (aw > 0.0) {
                        wmax = std::max(wmax, aw);
                        wmin = (wmin == 0.0) ? aw : std::min(wmin, aw);
                    }
                }
                if (wmax == 0.0) {  // all-zero weights cannot give max_auc > 0, but stay safe
                    acc_parts[r] = 1;
                    acc_scale.push_back(1.0);
                    continue;
                }
                int B, max_parts;
                double need;                           // the smallest weight, scaled, must reach this
                if (tie) {
                    B = Bt; max_parts = 2; need = 2097152.0;            // 2^21: 22 significant bits
                } else {
                    // W = llrint(w * 2^s): |sum| <= hits * Wmax * cutoff must stay below 2^62
                    const double hits = (double)std::min<int64_t>(e - b, std::max<int64_t>(rank_cutoff, 1));
                    B = 62 - (int)std::ceil(std::log2(hits * (double)std::max<int64_t>(rank_cutoff, 1)));
                    B = std::max(24, std::min(B, 46));   // W is stored in 48 signed bits
                    max_parts = 4; need = 8388608.0;                    // 2^23: 24 significant bits
                }
                int ex = 0;
                std::frexp(wmax, &ex);                 // wmax = f * 2^ex, f in [0.5, 1)
                const int s1 = B - ex;                 // wmax * 2^s1 in [2^(B-1), 2^B)
                // parts: W_k = round(residual_{k-1} * 2^(s1 + (k-1)(B-1))), enough of them for the smallest weight
                int parts = 1;
                while (parts < max_parts && std::ldexp(wmin, s1 + (parts - 1) * (B - 1)) < need) ++parts;
                if (std::ldexp(wmin, s1 + (parts - 1) * (B - 1)) < need) {
                    if (tie) return 1;                 // not representable here: the caller falls back to "wide"
                    return fail(AUCELL_ERR_UNSUPPORTED,
                                "the weights of one regulon span more than ~45 orders of magnitude; cannot guarantee 1e-6");
                }
                acc_parts[r] = (uint8_t)parts;
                for (int k = 0; k < parts; ++k) acc_scale.push_back(std::ldexp(1.0, -(s1 + k * (B - 1))));
                for (int64_t k = b; k < e; ++k) {
                    double res = h_reg_weights[k];
                    for (int l = 0; l < parts; ++l) {
                        const int sl = s1 + l * (B - 1);
                        // unsigned limbs cannot take a negative residual: every part but the last rounds down
                        const long long Wl = (tie && l + 1 < parts) ? (long long)std::floor(std::ldexp(res, sl))
                                                                    : std::llrint(std::ldexp(res, sl));
                        if (Wl != 0) ents.push_back({inv[h_reg_cols[k]], (uint32_t)(first + l), Wl});
                        res -= std::ldexp((double)Wl, -sl);    // exact: Wl * 2^-sl is res on a coarser grid
                    }
                }
            }
            return AUCELL_OK;
        };
        {
            int qrc = tie_q ? quantise(true) : 1;
            if (qrc == 1) {
                if (p->weighted) tie_q = false;        // (unit weights never return 1)
                qrc = quantise(false);
            }
            if (qrc != AUCELL_OK) {
                aucell_plan_destroy(p);
                return qrc;
            }
        }
        p->n_acc = (int)acc_scale.size();
        // ---- gene-major table T: counting sort of the entries by shuffled position --------------------
        std::vector<uint32_t> tptr((size_t)n_genes + 1, 0u);
        for (auto& t : ents) tptr[(size_t)t.pos + 1]++;
        for (int64_t g = 0; g < n_genes; ++g) tptr[g + 1] += tptr[g];
        const size_t nT = ents.size(), nalloc = std::max<size_t>(nT, 1);
        std::vector<uint32_t> tacc(nalloc, 0u);
        std::vector<long long> tw(nalloc, 0ll);
        {
            std::vector<uint32_t> cur(tptr.begin(), tptr.end() - 1);
            for (auto& t : ents) {
                const uint32_t at = cur[t.pos]++;
                tacc[at] = t.acc;
                tw[at] = t.w;
            }
        }
        PCK(cudaMalloc(&p->d_tptr, sizeof(uint32_t) * (n_genes + 1)));
        PCK(cudaMemcpy(p->d_tptr, tptr.data(), sizeof(uint32_t) * (n_genes + 1), cudaMemcpyHostToDevice));
        if (p->n_acc > 65535) {
            aucell_plan_destroy(p);
            return fail(AUCELL_ERR_UNSUPPORTED, "more than 65535 regulon accumulators");
        }
        if (p->weighted) {
            std::vector<unsigned long long> tent(nalloc, 0ull);
            for (size_t k = 0; k < nT; ++k)
                tent[k] = ((unsigned long long)tw[k] << 16) | (unsigned long long)(tacc[k] & 0xFFFFu);
            PCK(cudaMalloc(&p->d_tent, sizeof(unsigned long long) * nalloc));
            PCK(cudaMemcpy(p->d_tent, tent.data(), sizeof(unsigned long long) * nalloc, cudaMemcpyHostToDevice));
        } else {
            std::vector<uint16_t> t16(nalloc, 0);
            for (size_t k = 0; k < nT; ++k) t16[k] = (uint16_t)tacc[k];
            PCK(cudaMalloc(&p->d_tacc16, sizeof(uint16_t) * nalloc));
            PCK(cudaMemcpy(p->d_tacc16, t16.data(), sizeof(uint16_t) * nalloc, cudaMemcpyHostToDevice));
        }
        // ---- ELL copy of T for aucell_tie_kernel: two (weighted) / four (unit) slots per gene, chained overflow ----
        p->fast_ok = false;
        if (tie_q && p->n_acc + 32 < (int)kTiePadU) {
            PCK(cudaMalloc(&p->d_tie_stats, TIE_ST_N * sizeof(unsigned long long)));
            PCK(cudaMemset(p->d_tie_stats, 0, TIE_ST_N * sizeof(unsigned long long)));
            if (p->weighted) {
                {
                    // (32-byte rows -- plans with more than 991 accumulators; smaller plans use the compact rows below --:
                    // four direct slots; AUCELL_B200_TIE_SLOTS=3 leaves the fourth to the chain, measured +4 % at 50..250
                    // regulons and -1 % at 500)
                    const char* envs = getenv("AUCELL_B200_TIE_SLOTS");
                    p->tie_slots = (envs && envs[0] == '3') ? 3 : 4;
                }
                std::vector<uint4> ell(2 * (size_t)n_genes);
                std::vector<uint2> xw;
                for (int64_t g = 0; g < n_genes; ++g) {
                    const uint32_t b = tptr[g], e = tptr[g + 1], len = e - b;
                    // empty slots: W = 0 into one of the 32 dummy accumulator pairs behind the real ones
                    const uint32_t pad = (uint32_t)(2 * p->n_acc + 2 * (g & 31)) * 4u;
                    uint32_t sw[4] = {0u, 0u, 0u, 0u}, so[4] = {pad, pad, pad, pad};
                    // (tie_slots == 3: the fourth slot only ever holds the chain, the kernel skips its atomics)
                    const uint32_t nslots = (uint32_t)p->tie_slots;
                    const uint32_t direct = len > nslots ? 3u : len;
                    // byte offset of an accumulator's low limb: the limbs swap places every 16 accumulators (tie_slot_w)
                    auto lo_off = [](uint32_t acc) { return acc * 8u + ((acc & 16u) ? 4u : 0u); };
                    for (uint32_t k = 0; k < direct; ++k) { sw[k] = (uint32_t)tw[b + k]; so[k] = lo_off(tacc[b + k]); }
                    if (len > nslots) {
                        sw[3] = (uint32_t)xw.size();
                        so[3] = 0x80000000u | (len - 3);
                        for (uint32_t k = b + 3; k < e; ++k) xw.push_back(make_uint2((uint32_t)tw[k], lo_off(tacc[k])));
                    }
                    ell[2 * (size_t)g] = make_uint4(sw[0], so[0], sw[1], so[1]);
                    ell[2 * (size_t)g + 1] = make_uint4(sw[2], so[2], sw[3], so[3]);
                }
                // compact rows (one 128-bit load per ranked gene instead of two: the row gathers are a quarter of the
                // kernel's L1 data-pipe traffic): three weights + three 10-bit accumulator indices, chains looked up by
                // position.  When the accumulators (and the 32 dummies behind them) can be named in 10 bits.
                {
                    const char* env16 = getenv("AUCELL_B200_TIE_ROW16");
                    if (p->n_acc + 32 <= 1023 && !(env16 && env16[0] == '0')) {
                        std::vector<uint4> e16((size_t)n_genes);
                        std::vector<uint2> xinfo((size_t)n_genes, make_uint2(0u, 0u));
                        std::vector<uint2> xw16;
                        auto lo_off16 = [](uint32_t acc) { return acc * 8u + ((acc & 16u) ? 4u : 0u); };
                        for (int64_t g = 0; g < n_genes; ++g) {
                            const uint32_t b = tptr[g], e = tptr[g + 1], len = e - b;
                            const uint32_t padi = (uint32_t)p->n_acc + (uint32_t)(g & 31);
                            uint32_t w3[3] = {0u, 0u, 0u}, i3[3] = {padi, padi, padi};
                            const uint32_t direct = std::min<uint32_t>(len, 3u);
                            for (uint32_t k = 0; k < direct; ++k) { w3[k] = (uint32_t)tw[b + k]; i3[k] = tacc[b + k]; }
                            uint32_t pk = i3[0] | (i3[1] << 10) | (i3[2] << 20);
                            if (len > 3) {
                                pk |= 0x80000000u;
                                xinfo[(size_t)g] = make_uint2((uint32_t)xw16.size(), len - 3);
                                for (uint32_t k = b + 3; k < e; ++k) xw16.push_back(make_uint2((uint32_t)tw[k], lo_off16(tacc[k])));
                            }
                            e16[(size_t)g] = make_uint4(w3[0], w3[1], w3[2], pk);
                        }
                        xw.swap(xw16);             // the chained slots of THIS format (the 32-byte rows stay as a fallback table)
                        PCK(cudaMalloc(&p->d_ell_w16, sizeof(uint4) * e16.size()));
                        PCK(cudaMemcpy(p->d_ell_w16, e16.data(), sizeof(uint4) * e16.size(), cudaMemcpyHostToDevice));
                        PCK(cudaMalloc(&p->d_xinfo, sizeof(uint2) * xinfo.size()));
                        PCK(cudaMemcpy(p->d_xinfo, xinfo.data(), sizeof(uint2) * xinfo.size(), cudaMemcpyHostToDevice));
                    }
                }
                if (xw.empty()) xw.push_back(make_uint2(0u, 0u));
                PCK(cudaMalloc(&p->d_ell_w, sizeof(uint4) * ell.size()));
                PCK(cudaMemcpy(p->d_ell_w, ell.data(), sizeof(uint4) * ell.size(), cudaMemcpyHostToDevice));
                PCK(cudaMalloc(&p->d_xw, sizeof(uint2) * xw.size()));
                PCK(cudaMemcpy(p->d_xw, xw.data(), sizeof(uint2) * xw.size(), cudaMemcpyHostToDevice));
                p->tie_L = tieL;
            } else {
                std::vector<uint2> ell((size_t)n_genes);
                std::vector<uint16_t> xu;
                for (int64_t g = 0; g < n_genes; ++g) {
                    const uint32_t b = tptr[g], e = tptr[g + 1], len = e - b;
                    const uint32_t padu = (uint32_t)p->n_acc + (uint32_t)(g & 31);     // a dummy accumulator
                    uint32_t s4[4] = {padu, padu, padu, padu};
                    for (uint32_t k = 0; k < std::min<uint32_t>(len, 4u); ++k) s4[k] = tacc[b + k];
                    uint2 row = make_uint2(s4[0] | (s4[1] << 16), s4[2] | (s4[3] << 16));
                    if (len - 2 > 65535u && len > 4) { xu.assign((size_t)1 << 31, 0); break; }   // (never: n_acc < 32767)
                    if (len > 4) {
                        row.y = 0x80000000u | (uint32_t)xu.size();
                        xu.push_back((uint16_t)(len - 2));
                        for (uint32_t k = b + 2; k < e; ++k) xu.push_back((uint16_t)tacc[k]);
                    }
                    ell[(size_t)g] = row;
                }
                if (xu.empty()) xu.push_back(0);
                if (xu.size() < ((size_t)1 << 31) && nT < ((size_t)1 << 31)) {
                    PCK(cudaMalloc(&p->d_ell_u, sizeof(uint2) * (size_t)n_genes));
                    PCK(cudaMemcpy(p->d_ell_u, ell.data(), sizeof(uint2) * (size_t)n_genes, cudaMemcpyHostToDevice));
                    PCK(cudaMalloc(&p->d_xu, sizeof(uint16_t) * xu.size()));
                    PCK(cudaMemcpy(p->d_xu, xu.data(), sizeof(uint16_t) * xu.size(), cudaMemcpyHostToDevice));
                }
            }
            p->fast_ok = p->weighted ? p->d_ell_w != nullptr : p->d_ell_u != nullptr;
        }
        const size_t nacc = std::max<size_t>(acc_scale.size(), 1);
        acc_scale.resize(nacc, 1.0);
        PCK(cudaMalloc(&p->d_acc_first, sizeof(int32_t) * n_regulons));
        PCK(cudaMemcpy(p->d_acc_first, acc_first.data(), sizeof(int32_t) * n_regulons, cudaMemcpyHostToDevice));
        PCK(cudaMalloc(&p->d_acc_parts, n_regulons));
        PCK(cudaMemcpy(p->d_acc_parts, acc_parts.data(), n_regulons, cudaMemcpyHostToDevice));
        PCK(cudaMalloc(&p->d_acc_scale, sizeof(double) * nacc));
        PCK(cudaMemcpy(p->d_acc_scale, acc_scale.data(), sizeof(double) * nacc, cudaMemcpyHostToDevice));
        {
            struct Rec { double s0, s1, mx; int32_t first, has2; };
            static_assert(sizeof(Rec) == 32, "epilogue record");
            std::vector<Rec> recs((size_t)n_regulons);
            for (int r = 0; r < n_regulons; ++r) {
                Rec rc{0.0, 0.0, 1.0, -1, 0};
                if (acc_first[r] >= 0) {
                    rc.first = acc_first[r];
                    rc.s0 = acc_scale[(size_t)acc_first[r]];
                    rc.mx = h_reg_max_auc[r];
                    if (acc_parts[r] > 1) { rc.has2 = 1; rc.s1 = acc_scale[(size_t)acc_first[r] + 1]; }
                }
                recs[(size_t)r] = rc;
            }
            PCK(cudaMalloc(&p->d_reg_rec, sizeof(Rec) * (size_t)n_regulons));
            PCK(cudaMemcpy(p->d_reg_rec, recs.data(), sizeof(Rec) * (size_t)n_regulons, cudaMemcpyHostToDevice));
        }
        PCK(cudaMalloc(&p->d_reg_max_auc, sizeof(double) * n_regulons));
        PCK(cudaMemcpy(p->d_reg_max_auc, h_reg_max_auc, sizeof(double) * n_regulons, cudaMemcpyHostToDevice));
    }
#undef PCK
    *out_plan = p;
    return AUCELL_OK;
}

int aucell_plan_destroy(aucell_plan_t* p) {
    if (!p) return AUCELL_OK;
    DeviceGuard dg(p->device);
    cudaFree(p->d_stats);
    cudaFree(p->d_inv_perm);
    cudaFree(p->d_inv_perm16);
    cudaFree(p->d_tptr);
    cudaFree(p->d_tent);
    cudaFree(p->d_tacc16);
    cudaFree(p->d_ell_w);
    cudaFree(p->d_ell_w16);
    cudaFree(p->d_xinfo);
    cudaFree(p->d_xw);
    cudaFree(p->d_ell_u);
    cudaFree(p->d_xu);
    cudaFree(p->d_tie_stats);
    cudaFree(p->d_reg_rec);
    cudaFree(p->d_acc_first);
    cudaFree(p->d_acc_parts);
    cudaFree(p->d_acc_scale);
    cudaFree(p->d_reg_max_auc);
    for (auto& kv : p->scratch) cudaFree(kv.second.first);
    delete p;          // (the staging buffers of the host-buffer calls belong to the device, not to the plan: staging_of)
    return AUCELL_OK;
}

int aucell_plan_last_host_bytes(const aucell_plan_t* p, int64_t* out2) {
    if (!p || !out2) return fail(AUCELL_ERR_INVALID, "NULL pointer");
    std::lock_guard<std::mutex> host_lock(p->host_mu);
    out2[0] = p->last_h2d_bytes;
    out2[1] = p->last_d2h_bytes;
    return AUCELL_OK;
}

int aucell_plan_path_stats(const aucell_plan_t* p, uint64_t* out5, int reset) {
    if (!p || !out5) return fail(AUCELL_ERR_INVALID, "NULL pointer");
    DeviceGuard dg(p->device);
    if (!dg.ok) return fail(AUCELL_ERR_CUDA, "cudaSetDevice failed");
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out5, p->d_stats, 5 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    if (reset) CK(cudaMemset(p->d_stats, 0, 8 * sizeof(unsigned long long)));
    return AUCELL_OK;
}

int aucell_plan_set_fast_path(aucell_plan_t* p, int enable, int* out_available) {
    if (!p) return fail(AUCELL_ERR_INVALID, "plan is NULL");
    p->fast_enabled = enable != 0;
    if (out_available) *out_available = p->fast_ok ? 1 : 0;
    return AUCELL_OK;
}

int aucell_plan_fast_stats(const aucell_plan_t* p, uint64_t* out8, int reset) {
    if (!p || !out8) return fail(AUCELL_ERR_INVALID, "NULL pointer");
    for (int k = 0; k < TIE_ST_N; ++k) out8[k] = 0;
    if (!p->d_tie_stats) return AUCELL_OK;
    DeviceGuard dg(p->device);
    if (!dg.ok) return fail(AUCELL_ERR_CUDA, "cudaSetDevice failed");
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out8, p->d_tie_stats, TIE_ST_N * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    if (reset) CK(cudaMemset(p->d_tie_stats, 0, TIE_ST_N * sizeof(unsigned long long)));
    return AUCELL_OK;
}

// ---- create_rankings ---------------------------------------------------------------------------------
int aucell_rank_dense_f32(const aucell_plan_t* plan, const float* d_expr, int64_t n_cells, int64_t row_stride,
                          uint32_t* d_out_rank, int32_t* d_out_nnz, void* stream) {
    return rank_dense_f32(plan, d_expr, n_cells, row_stride, d_out_rank, d_out_nnz, stream);
}
int aucell_rank_dense_f64(const aucell_plan_t* plan, const double* d_expr, int64_t n_cells, int64_t row_stride,
                          uint32_t* d_out_rank, int32_t* d_out_nnz, void* stream) {
    return rank_dense_f64(plan, d_expr, n_cells, row_stride, d_out_rank, d_out_nnz, stream);
}
int aucell_rank_csr_f32(const aucell_plan_t* plan, const int64_t* d_indptr, const int32_t* d_indices,
                        const float* d_data, int64_t n_cells, uint32_t* d_out_rank, int32_t* d_out_nnz,
                        void* stream) {
    return rank_csr_f32(plan, d_indptr, d_indices, d_data, n_cells, d_out_rank, d_out_nnz, stream);
}
int aucell_rank_csr_f64(const aucell_plan_t* plan, const int64_t* d_indptr, const int32_t* d_indices,
                        const double* d_data, int64_t n_cells, uint32_t* d_out_rank, int32_t* d_out_nnz,
                        void* stream) {
    return rank_csr_f64(plan, d_indptr, d_indices, d_data, n_cells, d_out_rank, d_out_nnz, stream);
}

// ---- fused aucell ------------------------------------------------------------------------------------
int aucell_dense_f32(const aucell_plan_t* plan, const float* d_expr, int64_t n_cells, int64_t row_stride,
                     double* d_out_auc, int32_t* d_out_nnz, void* stream) {
    return fused_dense_f32(plan, d_expr, n_cells, row_stride, d_out_auc, d_out_nnz, stream);
}
int aucell_dense_f64(const aucell_plan_t* plan, const double* d_expr, int64_t n_cells, int64_t row_stride,
                     double* d_out_auc, int32_t* d_out_nnz, void* stream) {
    return fused_dense_f64(plan, d_expr, n_cells, row_stride, d_out_auc, d_out_nnz, stream);
}
int aucell_csr_f32(const aucell_plan_t* plan, const int64_t* d_indptr, const int32_t* d_indices,
                   const float* d_data, int64_t n_cells, double* d_out_auc, int32_t* d_out_nnz, void* stream) {
    return fused_csr_f32(plan, d_indptr, d_indices, d_data, n_cells, d_out_auc, d_out_nnz, stream);
}
int aucell_csr_f64(const aucell_plan_t* plan, const int64_t* d_indptr, const int32_t* d_indices,
                   const double* d_data, int64_t n_cells, double* d_out_auc, int32_t* d_out_nnz, void* stream) {
    return fused_csr_f64(plan, d_indptr, d_indices, d_data, n_cells, d_out_auc, d_out_nnz, stream);
}
int aucell_csr16_f32(const aucell_plan_t* plan, const int64_t* d_indptr, const uint16_t* d_indices,
                     const float* d_data, int64_t n_cells, double* d_out_auc, int32_t* d_out_nnz, void* stream) {
    if (plan && !plan->pos16) return fail(AUCELL_ERR_INVALID, "uint16 column indices need n_genes <= 65535");
    return fused_csr16_f32(plan, d_indptr, d_indices, d_data, n_cells, d_out_auc, d_out_nnz, stream);
}

int aucell_csr16_lut8_f32(const aucell_plan_t* plan, const int64_t* d_indptr, const uint16_t* d_indices,
                          const uint8_t* d_codes, const float* d_lut, float* d_vals_scratch, int64_t n_cells,
                          double* d_out_auc, int32_t* d_out_nnz, void* stream) {
    return fused_csr16_lut8(plan, d_indptr, d_indices, d_codes, d_lut, d_vals_scratch, n_cells, d_out_auc, d_out_nnz, stream);
}

int aucell_csr_lut8_f32(const aucell_plan_t* plan, const int64_t* d_indptr, const int32_t* d_indices,
                        const uint8_t* d_codes, const float* d_lut, float* d_vals_scratch, int64_t n_cells,
                        double* d_out_auc, int32_t* d_out_nnz, void* stream) {
    return fused_csr_lut8(plan, d_indptr, d_indices, d_codes, d_lut, d_vals_scratch, n_cells, d_out_auc, d_out_nnz, stream);
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
    const int smem = align16(kHitCap * (int)sizeof(uint2) + p->n_acc * 8);     // hit list | accumulators
    if (smem > p->smem_optin) return fail(AUCELL_ERR_UNSUPPORTED, "too many regulons for the shared-memory accumulators");
    GeneTableDev reg = tab_of(p);
    if (p->acc64) {
        auto k = aucell_from_rankings_kernel<true>;
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        const int grid = (int)std::min<int64_t>(n_cells, (int64_t)p->sm_count * resident_ctas(k, kF, smem));
        k<<<grid, kF, smem, st>>>(d_rank, n_cells, row_stride, p->n_genes, (int)p->rank_cutoff, reg, d_out_auc);
    } else {
        auto k = aucell_from_rankings_kernel<false>;
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        const int grid = (int)std::min<int64_t>(n_cells, (int64_t)p->sm_count * resident_ctas(k, kF, smem));
        k<<<grid, kF, smem, st>>>(d_rank, n_cells, row_stride, p->n_genes, (int)p->rank_cutoff, reg, d_out_auc);
    }
    launch_counter().fetch_add(1);
    CK(cudaGetLastError());
    return AUCELL_OK;
}

// ---- host-buffer streaming pipeline ------------------------------------------------------------------
// Inside the chunk loops a failed CUDA call must not return at once: earlier chunks still have copies in flight that
// read the caller's arrays and write h_out_auc.  CKB records the error and breaks; every exit then passes the loop
// that synchronises the slot streams.
#define CKB(call)                                                         \
    {                                                                     \
        cudaError_t e_ = (call);                                          \
        if (e_ != cudaSuccess) {                                          \
            rc = fail_cuda(e_, #call, __LINE__);                          \
            break;                                                        \
        }                                                                 \
    }

namespace {

constexpr int kHostSlots = 4;

// ---- host worker threads --------------------------------------------------------------------------------------------
// The host-side passes (narrowing, byte-coding, staging copies, squeezing) are short: a chunk's pass takes 1-3 ms, and
// creating sixteen std::threads for it ~0.3 ms.  The workers are therefore created once and kept (detached, for the
// life of the process); parallel_tasks(n, f) runs f(0) .. f(n - 1), task 0 on the caller, one parallel region at a time
// (regions of different callers queue: they would share the cores anyway).  f must not call parallel_tasks itself.
extern "C++" {
class HostPool {
    std::mutex client_mu_;                    // one region at a time
    std::mutex mu_;
    std::condition_variable cv_work_, cv_done_;
    int n_workers_ = 0;
    const std::function<void(int)>* fn_ = nullptr;
    int n_tasks_ = 0, next_ = 0, done_ = 0;
    void worker() {
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            cv_work_.wait(lk, [this]() { return fn_ != nullptr && next_ < n_tasks_; });
            const int t = next_++;
            const std::function<void(int)>* f = fn_;
            lk.unlock();
            (*f)(t);
            lk.lock();
            if (++done_ == n_tasks_) cv_done_.notify_all();
        }
    }
public:
    void run(int n, const std::function<void(int)>& f) {
        if (n <= 1) {
            if (n == 1) f(0);
            return;
        }
        std::lock_guard<std::mutex> client(client_mu_);
        {
            std::unique_lock<std::mutex> lk(mu_);
            while (n_workers_ < n - 1 && n_workers_ < 63) {
                try {
                    std::thread([this]() { worker(); }).detach();
                    ++n_workers_;
                } catch (...) {
                    break;                    // no more threads to be had (container limits): the caller takes what is left
                }
            }
            fn_ = &f;
            n_tasks_ = n;
            next_ = 1;                        // task 0 runs on the caller
            done_ = 0;
        }
        cv_work_.notify_all();
        f(0);
        std::unique_lock<std::mutex> lk(mu_);
        ++done_;
        while (next_ < n_tasks_) {            // help out (also covers "fewer workers than tasks")
            const int t = next_++;
            lk.unlock();
            f(t);
            lk.lock();
            ++done_;
        }
        cv_done_.wait(lk, [this]() { return done_ == n_tasks_; });
        fn_ = nullptr;
    }
};
inline HostPool& host_pool() {
    static HostPool* pool = new HostPool();   // (never destroyed: its workers outlive main)
    return *pool;
}
template <typename F> inline void parallel_tasks(int n, F&& f) {
    const std::function<void(int)> fn(std::forward<F>(f));
    host_pool().run(n, fn);
}
}  // extern "C++"

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// Staging of the host-buffer calls: kHostSlots chunk slots (stream, device buffer, pinned host buffers) PER DEVICE,
// kept for the life of the process (aucell_b200_release_staging frees them).  A plan is made per aucell() call, and
// page-locking a few hundred MB costs more than ranking 300,000 cells: the buffers outlive the plans.  One host-buffer
// call at a time per device (they would share the PCIe link anyway).
struct DeviceStaging {
    std::mutex mu;
    std::vector<aucell_plan::HostSlot> slots;
};
std::mutex& staging_map_mu() { static std::mutex m; return m; }
std::map<int, std::unique_ptr<DeviceStaging>>& staging_map() { static std::map<int, std::unique_ptr<DeviceStaging>> m; return m; }
DeviceStaging& staging_of(int device) {
    std::lock_guard<std::mutex> lk(staging_map_mu());
    auto& e = staging_map()[device];
    if (!e) e.reset(new DeviceStaging());
    return *e;
}
void free_slots(std::vector<aucell_plan::HostSlot>& slots) {
    for (auto& hs : slots) {
        if (hs.st) { cudaStreamSynchronize(hs.st); cudaStreamDestroy(hs.st); }
        cudaFree(hs.buf);
        cudaFreeHost(hs.pin_idx);
        cudaFreeHost(hs.pin_codes);
        cudaFreeHost(hs.pin_out);
        if (hs.idx_done) cudaEventDestroy(hs.idx_done);
    }
    slots.clear();
}

// Make sure the device's kHostSlots staging slots hold at least `bytes` each (and their streams).
int ensure_host_slots(std::vector<aucell_plan::HostSlot>& slots, size_t bytes, size_t pin_idx_bytes = 0,
                      size_t pin_codes_bytes = 0, size_t pin_out_bytes = 0) {
    if (slots.empty()) slots.resize(kHostSlots);
    // grow with head-room: the next call's chunks are rarely the same size to the byte, and re-pinning costs ~0.1 s per GB
    auto roomy = [](size_t want) { return want == 0 ? (size_t)0 : ((want + want / 4 + ((size_t)8 << 20)) & ~(((size_t)1 << 20) - 1)); };
    for (auto& hs : slots) {
        if (!hs.st) CK(cudaStreamCreateWithFlags(&hs.st, cudaStreamNonBlocking));
        if (hs.bytes < bytes) {
            CK(cudaStreamSynchronize(hs.st));
            cudaFree(hs.buf);
            hs.buf = nullptr;
            hs.bytes = 0;
            CK(cudaMalloc(&hs.buf, roomy(bytes)));
            hs.bytes = roomy(bytes);
        }
        if (hs.pin_bytes < pin_idx_bytes) {
            CK(cudaStreamSynchronize(hs.st));
            cudaFreeHost(hs.pin_idx);
            hs.pin_idx = nullptr;
            hs.pin_bytes = 0;
            CK(cudaHostAlloc((void**)&hs.pin_idx, roomy(pin_idx_bytes), cudaHostAllocDefault));
            hs.pin_bytes = roomy(pin_idx_bytes);
        }
        if (hs.pin_codes_bytes < pin_codes_bytes) {
            CK(cudaStreamSynchronize(hs.st));
            cudaFreeHost(hs.pin_codes);
            hs.pin_codes = nullptr;
            hs.pin_codes_bytes = 0;
            CK(cudaHostAlloc((void**)&hs.pin_codes, roomy(pin_codes_bytes), cudaHostAllocDefault));
            hs.pin_codes_bytes = roomy(pin_codes_bytes);
        }
        if (hs.pin_out_bytes < pin_out_bytes) {
            CK(cudaStreamSynchronize(hs.st));
            cudaFreeHost(hs.pin_out);
            hs.pin_out = nullptr;
            hs.pin_out_bytes = 0;
            CK(cudaHostAlloc((void**)&hs.pin_out, roomy(pin_out_bytes), cudaHostAllocDefault));
            hs.pin_out_bytes = roomy(pin_out_bytes);
        }
        hs.out_nc = 0;
        if ((pin_idx_bytes || pin_codes_bytes) && !hs.idx_done) CK(cudaEventCreateWithFlags(&hs.idx_done, cudaEventDisableTiming));
        hs.idx_pending = false;
    }
    return AUCELL_OK;
}

// int32 -> uint16 column indices of one chunk, spread over host threads (the loop is memory bound: a PCIe 5 x16 link
// wants ~27 GB/s of int32 indices narrowed).  Returns false when an index does not fit 16 bits.
#if defined(__x86_64__) && defined(__GNUC__)
#define AUCELL_HAVE_AVX2_PATH 1
__attribute__((target("avx2"))) uint32_t narrow_span_avx2(const int32_t* src, uint16_t* dst, int64_t n) {
    __m256i bad = _mm256_setzero_si256();
    int64_t i = 0;
    for (; i + 16 <= n; i += 16) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 8));
        bad = _mm256_or_si256(bad, _mm256_or_si256(a, b));
        // packus works inside the 128-bit halves: [a0-3 b0-3 | a4-7 b4-7] -> reorder the 64-bit quarters
        const __m256i pk = _mm256_permute4x64_epi64(_mm256_packus_epi32(a, b), 0xD8);
        _mm256_storeu_si256(reinterpret_cast<__m256i*>(dst + i), pk);
    }
    alignas(32) uint32_t lanes[8];
    _mm256_store_si256(reinterpret_cast<__m256i*>(lanes), bad);
    uint32_t any = 0;
    for (int k = 0; k < 8; ++k) any |= lanes[k];
    for (; i < n; ++i) {
        const uint32_t v = (uint32_t)src[i];
        any |= v;
        dst[i] = (uint16_t)v;
    }
    return any >> 16;           // saturated lanes only occur together with a set high half, which is reported here
}
#endif

uint32_t narrow_span(const int32_t* src, uint16_t* dst, int64_t n) {
#ifdef AUCELL_HAVE_AVX2_PATH
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) return narrow_span_avx2(src, dst, n);
#endif
    uint32_t bad = 0;
    for (int64_t i = 0; i < n; ++i) {
        const uint32_t v = (uint32_t)src[i];
        bad |= v >> 16;
        dst[i] = (uint16_t)v;
    }
    return bad;
}

bool narrow_indices(const int32_t* src, uint16_t* dst, int64_t n, int threads) {
    if (threads <= 1 || n < (1 << 18)) return narrow_span(src, dst, n) == 0;
    std::vector<uint32_t> bad((size_t)threads, 0u);
    const int64_t per = ((n + threads - 1) / threads + 63) & ~int64_t(63);
    parallel_tasks(threads, [&bad, src, dst, n, per](int t) {
        const int64_t b = std::min(n, per * t), e = std::min(n, per * (t + 1));
        bad[(size_t)t] = b < e ? narrow_span(src + b, dst + b, e - b) : 0u;
    });
    uint32_t any = 0;
    for (uint32_t b : bad) any |= b;
    return any == 0;
}

// ---- one-byte value codes -----------------------------------------------------------------------------------------
// Expression matrices hold few distinct values (raw counts, log1p / size-factor normalised counts): an append-only
// dictionary of up to 256 float bit patterns turns the 4-byte values of the upload into 1-byte codes, losslessly
// (the GPU reads the value bits back from the table).  Open addressing over 1024 slots; the values seen first -- the
// frequent ones -- sit in their home slots, which is all the 8-wide AVX2 path probes.
struct ValueDict {
    static constexpr int kSlots = 1024, kMax = 256;
    alignas(64) uint32_t key[kSlots];
    alignas(64) uint32_t code[kSlots];      // 0xFFFFFFFF = empty
    uint32_t lut[kMax];
    int n = 0;
    ValueDict() { clear(); }
    void clear() {
        n = 0;
        for (int i = 0; i < kSlots; ++i) { key[i] = 0u; code[i] = 0xFFFFFFFFu; }
        for (int i = 0; i < kMax; ++i) lut[i] = 0u;
    }
    static uint32_t home(uint32_t bits) { return (bits * 0x9E3779B1u) >> 22; }
    int find(uint32_t bits) const {
        for (uint32_t h = home(bits);; h = (h + 1) & (kSlots - 1)) {
            if (code[h] == 0xFFFFFFFFu) return -1;
            if (key[h] == bits) return (int)code[h];
        }
    }
    bool insert(uint32_t bits) {            // false: full
        if (find(bits) >= 0) return true;
        if (n >= kMax) return false;
        uint32_t h = home(bits);
        while (code[h] != 0xFFFFFFFFu) h = (h + 1) & (kSlots - 1);
        key[h] = bits;
        code[h] = (uint32_t)n;
        lut[n++] = bits;
        return true;
    }
};

// Encode src[0..n) into dst; stops at the first value that is not in the dictionary and returns its index (n: done).
int64_t encode_span_scalar(const ValueDict& d, const uint32_t* src, uint8_t* dst, int64_t i, int64_t n) {
    for (; i < n; ++i) {
        const int c = d.find(src[i]);
        if (c < 0) return i;
        dst[i] = (uint8_t)c;
    }
    return n;
}

#ifdef AUCELL_HAVE_AVX2_PATH
// Eight values per step.  The dictionary's first kTop values (the most frequent ones of the seed sample: a handful of
// values make up > 99 % of a count matrix) are matched with compares against broadcast registers; a group holding any
// other value goes through the table (gathers), and a value the table does not hold ends the span.  Codes are written
// with streaming stores when `dst` is 32-byte aligned at a multiple of 32 values (the staging buffer is write-only).
__attribute__((target("avx2"))) int64_t encode_span_avx2(const ValueDict& d, const uint32_t* src, uint8_t* dst, int64_t n) {
    constexpr int kTop = 8;
    __m256i top[kTop];
    const int n_top = std::min(d.n, kTop);
    for (int t = 0; t < kTop; ++t) top[t] = _mm256_set1_epi32((int)d.lut[t < n_top ? t : 0]);
    const __m256i mul = _mm256_set1_epi32((int)0x9E3779B1u);
    const __m256i ones = _mm256_set1_epi32(-1);
    int64_t i = 0;
    alignas(32) uint8_t line[32];
    for (; i + 32 <= n; i += 32) {
        bool miss = false;
        for (int g = 0; g < 4 && !miss; ++g) {
            const __m256i bits = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 8 * g));
            __m256i c = _mm256_setzero_si256(), hit = _mm256_setzero_si256();
            for (int t = 0; t < kTop; ++t) {                 // (t >= n_top repeats value 0 with code 0: harmless)
                const __m256i eq = _mm256_cmpeq_epi32(bits, top[t]);
                hit = _mm256_or_si256(hit, eq);
                c = _mm256_or_si256(c, _mm256_and_si256(eq, _mm256_set1_epi32(t < n_top ? t : 0)));
            }
            if (_mm256_movemask_epi8(hit) != -1) {           // a rarer value among the eight: the table
                const __m256i h = _mm256_srli_epi32(_mm256_mullo_epi32(bits, mul), 22);
                const __m256i k = _mm256_i32gather_epi32(reinterpret_cast<const int*>(d.key), h, 4);
                const __m256i cc = _mm256_i32gather_epi32(reinterpret_cast<const int*>(d.code), h, 4);
                const __m256i ok = _mm256_or_si256(hit, _mm256_andnot_si256(_mm256_cmpeq_epi32(cc, ones), _mm256_cmpeq_epi32(k, bits)));
                if (_mm256_movemask_epi8(ok) != -1) {        // displaced or unknown: one by one
                    uint8_t tmp[8];
                    const int64_t stop = encode_span_scalar(d, src + i + 8 * g, tmp, 0, 8);
                    if (stop < 8) {                          // unknown value: flush what is done and stop there
                        memcpy(dst + i, line, (size_t)(8 * g));
                        memcpy(dst + i + 8 * g, tmp, (size_t)stop);
                        return i + 8 * g + stop;
                    }
                    memcpy(line + 8 * g, tmp, 8);
                    continue;
                }
                c = _mm256_blendv_epi8(cc, c, hit);
            }
            const __m256i p16 = _mm256_packus_epi32(c, c);
            const __m256i p8 = _mm256_packus_epi16(p16, p16);
            const uint32_t lo = (uint32_t)_mm256_extract_epi32(p8, 0), hi = (uint32_t)_mm256_extract_epi32(p8, 4);
            const uint64_t both = (uint64_t)lo | ((uint64_t)hi << 32);
            memcpy(line + 8 * g, &both, 8);
        }
        if ((reinterpret_cast<uintptr_t>(dst + i) & 31u) == 0)
            _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), _mm256_load_si256(reinterpret_cast<const __m256i*>(line)));
        else
            memcpy(dst + i, line, 32);
    }
    _mm_sfence();
    return encode_span_scalar(d, src, dst, i, n);
}
#endif

int64_t encode_span(const ValueDict& d, const uint32_t* src, uint8_t* dst, int64_t n) {
#ifdef AUCELL_HAVE_AVX2_PATH
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) return encode_span_avx2(d, src, dst, n);
#endif
    return encode_span_scalar(d, src, dst, 0, n);
}

// Encode a chunk's values on `threads` host threads; values the dictionary does not know yet are added (by the calling
// thread, between passes) while there is room.  Returns false when the chunk holds more than 256 distinct values.
bool encode_values(ValueDict& d, const float* src_f, uint8_t* dst, int64_t n, int threads) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(src_f);
    if (d.n == 0) {                         // seed: the distinct values of the first 64 k entries, most frequent first
        std::map<uint32_t, int64_t> freq;
        for (int64_t i = 0; i < std::min<int64_t>(n, 65536); ++i) {
            ++freq[src[i]];
            if (freq.size() > (size_t)ValueDict::kMax) return false;
        }
        std::vector<std::pair<int64_t, uint32_t>> by_count;
        for (auto& kv : freq) by_count.push_back({-kv.second, kv.first});
        std::sort(by_count.begin(), by_count.end());
        for (auto& kv : by_count)
            if (!d.insert(kv.second)) return false;
    }
    threads = std::max(1, n < (1 << 18) ? 1 : threads);
    const int64_t per = ((n + threads - 1) / threads + 63) & ~int64_t(63);
    std::vector<int64_t> pos((size_t)threads), end((size_t)threads);
    for (int t = 0; t < threads; ++t) { pos[(size_t)t] = std::min(n, per * t); end[(size_t)t] = std::min(n, per * (t + 1)); }
    for (;;) {
        auto work = [&d, src, dst, &pos, &end](int t) {
            const int64_t b = pos[(size_t)t], e = end[(size_t)t];
            if (b < e) pos[(size_t)t] = b + encode_span(d, src + b, dst + b, e - b);
        };
        parallel_tasks(threads, work);
        bool done = true;
        for (int t = 0; t < threads; ++t) {
            if (pos[(size_t)t] < end[(size_t)t]) {          // stopped at an unknown value: learn it, go on from there
                done = false;
                if (!d.insert(src[pos[(size_t)t]])) return false;
            }
        }
        if (done) return true;
    }
}

// Is this host pointer ordinary pageable memory (numpy / scipy arrays) rather than page-locked?  Copies from and to
// pageable memory go through the driver's own bounce buffer at a fraction of the link rate and block the caller.
bool is_pageable(const void* ptr) {
    if (!ptr) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return at.type == cudaMemoryTypeUnregistered;
}

// memcpy spread over host threads (one core streams ~10 GB/s; the link wants 50)
void copy_bytes(void* dst, const void* src, size_t n, int threads) {
    if (threads <= 1 || n < ((size_t)1 << 22)) {
        memcpy(dst, src, n);
        return;
    }
    const size_t per = ((n + threads - 1) / threads + 4095) & ~(size_t)4095;
    parallel_tasks(threads, [=](int t) {
        const size_t b = std::min(n, per * (size_t)t), e = std::min(n, per * ((size_t)t + 1));
        if (b < e) memcpy((char*)dst + b, (const char*)src + b, e - b);
    });
}

// ---- dense host rows -> CSR on host threads --------------------------------------------------------------------------
// A dense [cells x genes] float32 matrix of single-cell counts is ~93 % zeros: uploading it costs 4 bytes per GENE and
// cell, ranking it through the CSR kernels 3..8 bytes per EXPRESSED gene.  The host threads therefore squeeze the zeros
// out (one pass at memory speed, AVX2 compress through a permutation table) and the call continues as a CSR call.
// +0.0 / -0.0 are dropped (they rank as zeros either way), NaN is kept.
extern "C++" {
template <typename T> struct RawBuf {          // uninitialised growable array (std::vector would zero-fill what is about to be written)
    std::unique_ptr<T[]> p;
    size_t cap = 0;
    T* data() { return p.get(); }
    size_t size() const { return cap; }
    void grow(size_t n, size_t keep) {
        if (n <= cap) return;
        std::unique_ptr<T[]> q(new T[n]);
        if (keep) memcpy(q.get(), p.get(), sizeof(T) * keep);
        p.swap(q);
        cap = n;
    }
};
}  // extern "C++"
struct CompactLut {
    alignas(32) uint32_t perm[256][8];
    alignas(16) uint16_t lane[256][8];
    CompactLut() {
        for (int m = 0; m < 256; ++m) {
            int k = 0;
            for (int b = 0; b < 8; ++b)
                if (m & (1 << b)) { perm[m][k] = (uint32_t)b; lane[m][k] = (uint16_t)b; ++k; }
            for (; k < 8; ++k) { perm[m][k] = 0; lane[m][k] = 0; }
        }
    }
};
const CompactLut& compact_lut() { static const CompactLut l; return l; }

// non-zero entries of row[0..n) -> (idx, val), both with room for 8 entries past the result; returns their number
int64_t compact_row_scalar(const float* row, int64_t j0, int64_t n, int32_t* idx, float* val) {
    int64_t k = 0;
    for (int64_t j = j0; j < n; ++j) {
        const float v = row[j];
        if (v != 0.0f) { idx[k] = (int32_t)j; val[k] = v; ++k; }       // (NaN != 0: kept)
    }
    return k;
}
#ifdef AUCELL_HAVE_AVX2_PATH
__attribute__((target("avx2,popcnt"))) int64_t compact_row_avx2(const float* row, int64_t n, int32_t* idx, float* val) {
    const CompactLut& L = compact_lut();
    const __m256 zero = _mm256_setzero_ps();
    int64_t k = 0, j = 0;
    for (; j + 8 <= n; j += 8) {
        const __m256 v = _mm256_loadu_ps(row + j);
        const int m = _mm256_movemask_ps(_mm256_cmp_ps(v, zero, _CMP_NEQ_UQ));
        // branch-free: an all-zero group stores eight don't-cares that the next group overwrites (whether a group of
        // eight is empty is a coin flip at 7 % density: the mispredicted branch cost 2.7x the whole loop)
        const __m256i pm = _mm256_load_si256(reinterpret_cast<const __m256i*>(L.perm[m]));
        _mm256_storeu_ps(val + k, _mm256_permutevar8x32_ps(v, pm));
        _mm256_storeu_si256(reinterpret_cast<__m256i*>(idx + k), _mm256_add_epi32(pm, _mm256_set1_epi32((int)j)));
        k += __builtin_popcount((unsigned)m);
    }
    return k + compact_row_scalar(row, j, n, idx + k, val + k);
}
#endif
int64_t compact_row(const float* row, int64_t n, int32_t* idx, float* val) {
#ifdef AUCELL_HAVE_AVX2_PATH
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) return compact_row_avx2(row, n, idx, val);
#endif
    return compact_row_scalar(row, 0, n, idx, val);
}

constexpr int kNotCompacted = 1000;      // internal: the matrix is not sparse / not exact in float32; take another path

extern "C++" {
// value -> float32 when that is exact (a cast that changes a value may create ties the input does not have)
inline bool to_f32_exact(float v, float& f) { f = v; return true; }
inline bool to_f32_exact(double v, float& f) { f = (float)v; return (double)f == v || v != v; }
template <typename T> inline bool is_zero_bits(const T* p);
template <> inline bool is_zero_bits<float>(const float* p) { uint32_t u; memcpy(&u, p, 4); return (u << 1) == 0u; }
template <> inline bool is_zero_bits<double>(const double* p) { uint64_t u; memcpy(&u, p, 8); return (u << 1) == 0ull; }

// non-zero entries of one contiguous row; false: a value is not exact in float32
template <typename T> bool compact_row_t(const T* row, int64_t n, int32_t* idx, float* val, int64_t& k_out);
template <> bool compact_row_t<float>(const float* row, int64_t n, int32_t* idx, float* val, int64_t& k_out) {
    k_out = compact_row(row, n, idx, val);
    return true;
}
template <> bool compact_row_t<double>(const double* row, int64_t n, int32_t* idx, float* val, int64_t& k_out) {
    int64_t k = 0, j = 0;
    bool ok = true;
    for (; j + 4 <= n; j += 4) {                                 // (three groups of four in four hold only zeros)
        uint64_t u[4];
        memcpy(u, row + j, 32);
        if ((((u[0] | u[1]) | (u[2] | u[3])) << 1) == 0ull) continue;
        for (int q = 0; q < 4; ++q) {
            if ((u[q] << 1) == 0ull) continue;
            float f;
            ok &= to_f32_exact(row[j + q], f);
            idx[k] = (int32_t)(j + q);
            val[k] = f;
            ++k;
        }
    }
    for (; j < n; ++j) {
        if (is_zero_bits(row + j)) continue;
        float f;
        ok &= to_f32_exact(row[j], f);
        idx[k] = (int32_t)j;
        val[k] = f;
        ++k;
    }
    k_out = k;
    return ok;
}

// One tile of a gene-major block: `tn` consecutive cells x G genes; the non-zeros of gene g are appended to their cells'
// lists (cell i: tix / tvv [i * capR, (i + 1) * capR), length cur[i]).  AVX2: eight float32 / four float64 cells per
// compare, only the non-zero lanes are visited (93 % of the values are zeros).
template <typename T>
inline void tile_put(const T* col, int64_t i, int64_t rs, int64_t g, int64_t capR, int64_t* cur, int32_t* tix, float* tvv,
                     bool& ok, bool& overflow) {
    float f;
    ok &= to_f32_exact(col[i * rs], f);
    const int64_t k = cur[i];
    if (k >= capR) { overflow = true; return; }
    tix[i * capR + k] = (int32_t)g;
    tvv[i * capR + k] = f;
    cur[i] = k + 1;
}
template <typename T>
void scan_tile_scalar(const T* tile, int64_t rs, int64_t cs, int64_t tn, int64_t G, int64_t capR, int64_t* cur,
                      int32_t* tix, float* tvv, bool& ok, bool& overflow) {
    for (int64_t g = 0; g < G; ++g) {
        const T* col = tile + g * cs;
        for (int64_t i = 0; i < tn; ++i)
            if (!is_zero_bits(col + i * rs)) tile_put<T>(col, i, rs, g, capR, cur, tix, tvv, ok, overflow);
    }
}
#ifdef AUCELL_HAVE_AVX2_PATH
__attribute__((target("avx2"))) inline unsigned nz_lanes(const float* p) {
    return (unsigned)_mm256_movemask_ps(_mm256_cmp_ps(_mm256_loadu_ps(p), _mm256_setzero_ps(), _CMP_NEQ_UQ));
}
__attribute__((target("avx2"))) inline unsigned nz_lanes(const double* p) {
    return (unsigned)_mm256_movemask_pd(_mm256_cmp_pd(_mm256_loadu_pd(p), _mm256_setzero_pd(), _CMP_NEQ_UQ));
}
template <typename T>
__attribute__((target("avx2"))) void scan_tile_avx2(const T* tile, int64_t cs, int64_t tn, int64_t G, int64_t capR,
                                                    int64_t* cur, int32_t* tix, float* tvv, bool& ok, bool& overflow) {
    constexpr int64_t V = 32 / (int64_t)sizeof(T);            // cells per 256-bit vector (consecutive: rs == 1)
    constexpr int64_t kAhead = 12;                           // genes: their cache lines are a whole column apart, which
    const int64_t lines = (tn * (int64_t)sizeof(T) + 63) / 64;   // no hardware prefetcher follows
    for (int64_t g = 0; g < G; ++g) {
        const T* col = tile + g * cs;
        if (g + kAhead < G) {
            const char* nxt = reinterpret_cast<const char*>(tile + (g + kAhead) * cs);
            for (int64_t l = 0; l < lines; ++l) _mm_prefetch(nxt + 64 * l, _MM_HINT_T0);
        }
        // the non-zero cells of this gene as one bit mask over the tile (tn <= 128), then ONE loop over its set bits: a
        // loop per vector of eight mispredicts its exit every other vector (+33 % on one core)
        uint64_t mk[2] = {0ull, 0ull};
        int64_t i = 0;
        for (; i + V <= tn; i += V) mk[i >> 6] |= (uint64_t)nz_lanes(col + i) << (i & 63);
        for (; i < tn; ++i) mk[i >> 6] |= (uint64_t)(is_zero_bits(col + i) ? 0 : 1) << (i & 63);
        for (int h = 0; h < 2; ++h) {
            uint64_t m = mk[h];
            while (m) {
                const int64_t b = (int64_t)__builtin_ctzll(m) + 64 * h;
                m &= m - 1;
                tile_put<T>(col, b, 1, g, capR, cur, tix, tvv, ok, overflow);
            }
        }
    }
}
#endif
template <typename T>
void scan_tile(const T* tile, int64_t rs, int64_t cs, int64_t tn, int64_t G, int64_t capR, int64_t* cur, int32_t* tix,
               float* tvv, bool& ok, bool& overflow) {
#ifdef AUCELL_HAVE_AVX2_PATH
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2 && rs == 1) {
        scan_tile_avx2<T>(tile, cs, tn, G, capR, cur, tix, tvv, ok, overflow);
        return;
    }
#endif
    scan_tile_scalar<T>(tile, rs, cs, tn, G, capR, cur, tix, tvv, ok, overflow);
}

// Rows [c0, c0 + nc) of a strided host matrix (element (r, g) at base[r * rs + g * cs]; cs == 1: cell-major rows,
// rs == 1: gene-major, the layout of a DataFrame block) -> CSR with float32 values, on `threads` host threads.
// Returns false when a value is not exact in float32 or the rows turn out far denser than `density`.
template <typename T>
bool compact_block(const T* base, int64_t rs, int64_t cs, int64_t c0, int64_t nc, int64_t G, int threads, double density,
                   std::vector<RawBuf<int32_t>>& t_idx, std::vector<RawBuf<float>>& t_val, std::vector<int64_t>& indptr,
                   RawBuf<int32_t>& indices, RawBuf<float>& values) {
    indptr.assign((size_t)nc + 1, 0);
    std::vector<int> bad((size_t)threads, 0);
    auto work = [&](int t) {
        const int64_t r0 = nc * t / threads, r1 = nc * (t + 1) / threads;
        RawBuf<int32_t>& ix = t_idx[(size_t)t];
        RawBuf<float>& vv = t_val[(size_t)t];
        size_t used = 0;
        const size_t guess = (size_t)((double)(r1 - r0) * (double)G * std::min(1.0, density * 1.25 + 0.01)) + (size_t)G + 8;
        ix.grow(guess, 0);
        vv.grow(guess, 0);
        bool ok = true;
        if (cs == 1) {
            for (int64_t r = r0; r < r1; ++r) {
                if (ix.size() < used + (size_t)G + 8) {          // denser than the probe said: grow
                    const size_t want = ix.size() + ix.size() / 2 + (size_t)G + 8;
                    ix.grow(want, used);
                    vv.grow(want, used);
                }
                int64_t k = 0;
                ok &= compact_row_t<T>(base + (c0 + r) * rs, G, ix.data() + used, vv.data() + used, k);
                indptr[(size_t)r + 1] = k;                       // (row length for now)
                used += (size_t)k;
            }
        } else {
            // gene-major: tiles of kTile cells; a gene contributes kTile consecutive values, whose non-zeros are appended
            // to their cells' lists (fixed capacity per cell inside the tile: overflow = far denser than expected)
            constexpr int kTile = 128;            // (64 .. 512 measured: no clear difference; scan_tile_avx2 keeps a 128-bit mask)
            int64_t capR = std::min<int64_t>(G, (int64_t)((double)G * density * 2.0) + 256);
            RawBuf<int32_t> tix;
            RawBuf<float> tvv;
            tix.grow((size_t)(kTile * capR), 0);
            tvv.grow((size_t)(kTile * capR), 0);
            int64_t cur[kTile];
            for (int64_t t0 = r0; t0 < r1 && ok; t0 += kTile) {
                const int64_t tn = std::min<int64_t>(kTile, r1 - t0);
                for (;;) {
                    bool overflow = false;
                    for (int64_t i = 0; i < tn; ++i) cur[i] = 0;
                    scan_tile<T>(base + (c0 + t0) * rs, rs, cs, tn, G, capR, cur, tix.data(), tvv.data(), ok, overflow);
                    if (!overflow) break;
                    capR = std::min<int64_t>(G, capR * 2);       // a cell far denser than the sample: once more, with room
                    tix.grow((size_t)(kTile * capR), 0);
                    tvv.grow((size_t)(kTile * capR), 0);
                }
                size_t need = used;
                for (int64_t i = 0; i < tn; ++i) need += (size_t)cur[i];
                if (ix.size() < need + 8) {
                    const size_t want = std::max(need + 8, ix.size() + ix.size() / 2);
                    ix.grow(want, used);
                    vv.grow(want, used);
                }
                for (int64_t i = 0; i < tn; ++i) {
                    memcpy(ix.data() + used, tix.data() + i * capR, sizeof(int32_t) * (size_t)cur[i]);
                    memcpy(vv.data() + used, tvv.data() + i * capR, sizeof(float) * (size_t)cur[i]);
                    indptr[(size_t)(t0 + i) + 1] = cur[i];
                    used += (size_t)cur[i];
                }
            }
        }
        bad[(size_t)t] = ok ? 0 : 1;
    };
    {
        parallel_tasks(threads, work);
    }
    for (int b : bad) if (b) return false;
    for (int64_t r = 0; r < nc; ++r) indptr[(size_t)r + 1] += indptr[(size_t)r];
    const int64_t nnz = indptr[(size_t)nc];
    indices.grow((size_t)nnz + 8, 0);
    values.grow((size_t)nnz + 8, 0);
    auto gather = [&](int t) {
        const int64_t r0 = nc * t / threads, r1 = nc * (t + 1) / threads;
        const int64_t e0 = indptr[(size_t)r0], e1 = indptr[(size_t)r1];
        memcpy(indices.data() + e0, t_idx[(size_t)t].data(), sizeof(int32_t) * (size_t)(e1 - e0));
        memcpy(values.data() + e0, t_val[(size_t)t].data(), sizeof(float) * (size_t)(e1 - e0));
    };
    {
        parallel_tasks(threads, gather);
    }
    return true;
}
}  // extern "C++"

int host_threads();

extern "C++" {
// A sparse host matrix (dense storage, any of the two layouts, float32 or float64) through the CSR host call.
// kNotCompacted: more than 30 % of a sample of its entries are non-zero, or a value does not survive the cast to float32
// (nothing has been written to the outputs that the caller cannot overwrite).
template <typename T>
int dense_host_via_csr(const aucell_plan* p, const T* base, int64_t n_cells, int64_t rs, int64_t cs, double* h_out_auc,
                       int32_t* h_out_nnz) {
    const int64_t G = p->n_genes, R = p->n_regulons;
    if (G < 64 || n_cells * G < ((int64_t)1 << 22)) return kNotCompacted;
    // density of a sample: 64 evenly spaced cells x up to 4096 evenly spaced genes
    const int64_t pc = std::min<int64_t>(n_cells, 64), pg = std::min<int64_t>(G, 4096);
    int64_t nz = 0;
    for (int64_t a = 0; a < pc; ++a)
        for (int64_t b = 0; b < pg; ++b)
            nz += is_zero_bits(base + (n_cells * a / pc) * rs + (G * b / pg) * cs) ? 0 : 1;
    const double density = (double)nz / (double)(pc * pg);
    if (density > 0.30) return kNotCompacted;
    const int threads = host_threads();
    // blocks of cells whose dense image is <= 1 GB: the CSR copy of a block stays small
    const int64_t block = std::max<int64_t>(1024, ((int64_t)1 << 30) / (int64_t)sizeof(T) / std::max<int64_t>(G, 1));
    std::vector<RawBuf<int32_t>> t_idx((size_t)threads);
    std::vector<RawBuf<float>> t_val((size_t)threads);
    std::vector<int64_t> indptr;
    RawBuf<int32_t> indices;
    RawBuf<float> values;
    for (int64_t c0 = 0; c0 < n_cells; c0 += block) {
        const int64_t nc = std::min(n_cells, c0 + block) - c0;
        const auto tc0 = std::chrono::steady_clock::now();
        if (!compact_block<T>(base, rs, cs, c0, nc, G, threads, density, t_idx, t_val, indptr, indices, values)) {
            if (c0 == 0) return kNotCompacted;
            return fail(AUCELL_ERR_INVALID, "expression values that are not exact in float32 appear late in the matrix: "
                                            "pass float64 device chunks instead");
        }
        const auto tc1 = std::chrono::steady_clock::now();
        const int rc = aucell_csr_f32_host(p, indptr.data(), indices.data(), values.data(), nc, h_out_auc + c0 * R,
                                           h_out_nnz ? h_out_nnz + c0 : nullptr, 0);
        if (getenv("AUCELL_B200_DEBUG"))
            fprintf(stderr, "[aucell_b200] dense host call: %lld cells x %lld genes -> %lld stored entries (%.1f %%), "
                            "squeezed in %.3f s, CSR call %.3f s\n",
                    (long long)nc, (long long)G, (long long)indptr[(size_t)nc],
                    100.0 * (double)indptr[(size_t)nc] / ((double)nc * (double)G),
                    std::chrono::duration<double>(tc1 - tc0).count(),
                    std::chrono::duration<double>(std::chrono::steady_clock::now() - tc1).count());
        if (rc) return rc;
    }
    return AUCELL_OK;
}
}  // extern "C++"

int host_threads() {
    const char* env = getenv("AUCELL_B200_HOST_THREADS");
    if (env && *env) return std::max(1, std::min(64, atoi(env)));
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 4;
    // ranks of one job on this host share its cores (torchrun exports LOCAL_WORLD_SIZE)
    const char* lws = getenv("LOCAL_WORLD_SIZE");
    if (lws && atoi(lws) > 1) hw = std::max(1u, hw / (unsigned)atoi(lws));
    return (int)std::max(1u, std::min(16u, hw));
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
    DeviceStaging& staging = staging_of(p->device);
    std::lock_guard<std::mutex> staging_lock(staging.mu);
    std::lock_guard<std::mutex> host_lock(p->host_mu);
    std::vector<aucell_plan::HostSlot>& host_slots = staging.slots;
    const int64_t R = p->n_regulons;
    const int64_t total_nnz = h_indptr[n_cells] - h_indptr[0];
    if (total_nnz > 0 && (!h_indices || !h_data)) return fail(AUCELL_ERR_INVALID, "NULL pointer");
    if (chunk_cells <= 0) {
        // ~128 MB of (index, value) pairs per chunk; 48 MB for a pageable caller, whose chunks pass through pinned
        // buffers of the same size (page-locking them is the larger part of a first call on a few hundred thousand cells)
        const double chunk_mb = (total_nnz > 0 && (is_pageable(h_indices) || is_pageable(h_data) || is_pageable(h_out_auc))) ? 48.0 : 128.0;
        const double per_cell = std::max<double>(1.0, (double)total_nnz / (double)n_cells);
        chunk_cells = (int64_t)std::max<double>(1024.0, (chunk_mb * 1024 * 1024) / (8.0 * per_cell));
    }
    chunk_cells = std::min(chunk_cells, n_cells);
    int64_t max_nnz = 0;
    for (int64_t c0 = 0; c0 < n_cells; c0 += chunk_cells) {
        const int64_t c1 = std::min(n_cells, c0 + chunk_cells);
        max_nnz = std::max(max_nnz, h_indptr[c1] - h_indptr[c0]);
    }
    const size_t o_indptr = 0;
    const size_t o_indices = align256(o_indptr + sizeof(int64_t) * (chunk_cells + 1));
    const size_t o_vals = align256(o_indices + sizeof(int32_t) * std::max<int64_t>(max_nnz, 1));
    const size_t o_codes = align256(o_vals + sizeof(float) * std::max<int64_t>(max_nnz, 1));
    const size_t o_lut = align256(o_codes + std::max<int64_t>(max_nnz, 1));
    const size_t o_auc = align256(o_lut + sizeof(uint32_t) * ValueDict::kMax);
    const size_t o_nnz = align256(o_auc + sizeof(double) * chunk_cells * R);
    const size_t bytes = align256(o_nnz + sizeof(int32_t) * chunk_cells);
    // The link is the bottleneck of this call, so it sends less than the caller's arrays hold:
    //  - 16-bit positions: the column indices go up as uint16 (AUCELL_B200_HOST_IDX16=0 keeps int32);
    //  - the values go up as one-byte codes into a table of the distinct values (AUCELL_B200_HOST_VAL8=0 keeps float32),
    //    while the matrix holds at most 256 of them (counts and their normalisations do).
    // 3 instead of 8 bytes per stored entry; both narrowings run on host threads, chunk by chunk, ahead of the uploads.
    const char* env16 = getenv("AUCELL_B200_HOST_IDX16");
    const char* env8 = getenv("AUCELL_B200_HOST_VAL8");
    const bool idx16_allowed = p->pos16 && max_nnz > 0 && !(env16 && env16[0] == '0');
    const bool idx16_forced = env16 && env16[0] == '1';
    const bool val8_allowed = max_nnz > 0 && p->fast_ok && !(env8 && env8[0] == '0');
    const bool val8_forced = env8 && env8[0] == '1';
    // start from the format the previous call on this plan settled on (same host, same kind of matrix)
    bool idx16 = idx16_allowed && (p->hint_idx16 || idx16_forced);
    bool val8 = val8_allowed && (p->hint_val8 || val8_forced);
    std::unique_ptr<ValueDict> dict(val8 ? new (std::nothrow) ValueDict() : nullptr);
    if (val8 && !dict) val8 = false;
    const auto t_call0 = std::chrono::steady_clock::now();
    double t_alloc = 0.0, t_drain = 0.0;  // (debug) seconds spent growing the staging / copying AUC blocks out
    double narrow_ns = 0.0;               // time spent narrowing and entries narrowed so far in this call
    int64_t narrow_n = 0;
    int narrow_chunks = 0;
    int chunk_no = 0;                     // pipeline rate once the slots are full: entries enqueued per second
    int64_t rate_n = 0;
    std::chrono::steady_clock::time_point rate_t0;
    // Pageable caller (numpy / scipy arrays): nothing is copied straight out of or into its memory.  Inputs that are not
    // narrowed anyway are copied into the pinned staging by host threads, the AUC blocks land in pinned memory and are
    // copied out when their slot comes round again.
    const bool in_pageable = max_nnz > 0 && (is_pageable(h_indices) || is_pageable(h_data));
    const bool out_pageable = is_pageable(h_out_auc) || (h_out_nnz && is_pageable(h_out_nnz));
    const int n_threads = (idx16 || val8 || in_pageable || out_pageable) ? host_threads() : 1;
    const size_t out_block = sizeof(double) * (size_t)chunk_cells * R;
    const auto t_al0 = std::chrono::steady_clock::now();
    int rc = ensure_host_slots(host_slots, bytes,
                               in_pageable ? sizeof(int32_t) * (size_t)max_nnz : (idx16_allowed ? sizeof(uint16_t) * (size_t)max_nnz : 0),
                               in_pageable ? sizeof(float) * (size_t)max_nnz + sizeof(uint32_t) * ValueDict::kMax
                                           : (val8 ? (size_t)max_nnz + sizeof(uint32_t) * ValueDict::kMax : 0),
                               out_pageable ? out_block + sizeof(int32_t) * (size_t)chunk_cells : 0);
    if (rc) return rc;
    t_alloc = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_al0).count();
    auto drain_out = [&](aucell_plan::HostSlot& s) -> cudaError_t {      // the slot's pending AUC block -> the caller's arrays
        if (s.out_nc == 0) return cudaSuccess;
        const cudaError_t e = cudaStreamSynchronize(s.st);
        if (e == cudaSuccess) {
            const auto td0 = std::chrono::steady_clock::now();
            copy_bytes(h_out_auc + s.out_c0 * R, s.pin_out, sizeof(double) * (size_t)s.out_nc * R, n_threads);
            if (h_out_nnz) memcpy(h_out_nnz + s.out_c0, s.pin_out + out_block, sizeof(int32_t) * (size_t)s.out_nc);
            t_drain += std::chrono::duration<double>(std::chrono::steady_clock::now() - td0).count();
        }
        s.out_nc = 0;
        return e;
    };
    std::thread drainer;
    cudaError_t drain_err = cudaSuccess;
    // a fresh output array (np.zeros / np.empty) has no pages yet: let the kernel map them in bulk on helper threads
    // instead of one fault per 4 KB inside the copy-out (Linux >= 5.14; elsewhere the call fails and nothing is lost)
    std::vector<std::thread> prefault;
    if (out_pageable && is_pageable(h_out_auc) && sizeof(double) * (size_t)n_cells * R >= ((size_t)32 << 20)) {
        const uintptr_t b = ((uintptr_t)h_out_auc + 4095) & ~(uintptr_t)4095;
        const uintptr_t e = ((uintptr_t)h_out_auc + sizeof(double) * (size_t)n_cells * R) & ~(uintptr_t)4095;
        const int parts = 4;
        for (int q = 0; q < parts && e > b; ++q) {
            const uintptr_t qb = b + (((e - b) / parts * q) & ~(uintptr_t)4095);
            const uintptr_t qe = q + 1 == parts ? e : b + (((e - b) / parts * (q + 1)) & ~(uintptr_t)4095);
            try {
                prefault.emplace_back([qb, qe]() {
#ifdef __linux__
                    if (qe > qb) madvise((void*)qb, qe - qb, 23 /* MADV_POPULATE_WRITE */);
#endif
                });
            } catch (...) {
            }
        }
    }
    int k = 0;
    int64_t h2d_bytes = 0, d2h_bytes = 0;
    for (int64_t c0 = 0; c0 < n_cells; c0 += chunk_cells, k = (k + 1) % kHostSlots) {
        aucell_plan::HostSlot& s = host_slots[k];
        int64_t* d_indptr = reinterpret_cast<int64_t*>(s.buf + o_indptr);
        int32_t* d_indices = reinterpret_cast<int32_t*>(s.buf + o_indices);
        float* d_vals = reinterpret_cast<float*>(s.buf + o_vals);
        uint8_t* d_codes = reinterpret_cast<uint8_t*>(s.buf + o_codes);
        float* d_lut = reinterpret_cast<float*>(s.buf + o_lut);
        double* d_auc = reinterpret_cast<double*>(s.buf + o_auc);
        int32_t* d_nnz = h_out_nnz ? reinterpret_cast<int32_t*>(s.buf + o_nnz) : nullptr;
        const int64_t c1 = std::min(n_cells, c0 + chunk_cells);
        const int64_t nc = c1 - c0;
        const int64_t e0 = h_indptr[c0], ne = h_indptr[c1] - e0;
        bool chunk8 = false;
        // the AUC block waiting in this slot goes out to the caller's (pageable) array on a helper thread, while this
        // thread prepares and enqueues the next chunk; joined before the slot's landing zone is written again
        if (out_pageable && s.out_nc != 0) {
            if (drainer.joinable()) drainer.join();
            if (drain_err != cudaSuccess) { rc = fail_cuda(drain_err, "cudaStreamSynchronize", __LINE__); break; }
            aucell_plan::HostSlot* const ps = &s;
            const int dev = p->device;
            try {
                drainer = std::thread([&drain_err, &drain_out, ps, dev]() {
                    cudaSetDevice(dev);
                    drain_err = drain_out(*ps);
                });
            } catch (...) {
                CKB(drain_out(s));
            }
        }
        // stream order on the slot's stream protects its buffers from the previous chunk that used them
        CKB(cudaMemcpyAsync(d_indptr, h_indptr + c0, sizeof(int64_t) * (nc + 1), cudaMemcpyHostToDevice, s.st));
        h2d_bytes += (int64_t)sizeof(int64_t) * (nc + 1);
        d2h_bytes += (int64_t)sizeof(double) * nc * R + (h_out_nnz ? (int64_t)sizeof(int32_t) * nc : 0);
        if (ne > 0) {
            const bool staged = idx16 || val8 || in_pageable;
            if (staged) {
                if (s.idx_pending) CKB(cudaEventSynchronize(s.idx_done));      // staging buffers free again
                s.idx_pending = false;
            }
            const auto t0 = std::chrono::steady_clock::now();
            if (idx16 && !narrow_indices(h_indices + e0, s.pin_idx, ne, n_threads)) {
                rc = fail(AUCELL_ERR_INVALID, "CSR column index out of range");
                break;
            }
            if (val8) {
                chunk8 = encode_values(*dict, h_data + e0, s.pin_codes, ne, n_threads);
                if (!chunk8) val8 = false;                                     // more than 256 distinct values: float32 from here on
            }
            if (staged) {
                narrow_ns += std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - t0).count();
                narrow_n += ne;
                ++narrow_chunks;
            }
            if (idx16) {
                CKB(cudaMemcpyAsync(d_indices, s.pin_idx, sizeof(uint16_t) * ne, cudaMemcpyHostToDevice, s.st));
                h2d_bytes += (int64_t)sizeof(uint16_t) * ne;
            } else if (in_pageable) {
                copy_bytes(s.pin_idx, h_indices + e0, sizeof(int32_t) * (size_t)ne, n_threads);
                CKB(cudaMemcpyAsync(d_indices, s.pin_idx, sizeof(int32_t) * ne, cudaMemcpyHostToDevice, s.st));
                h2d_bytes += (int64_t)sizeof(int32_t) * ne;
            } else {
                CKB(cudaMemcpyAsync(d_indices, h_indices + e0, sizeof(int32_t) * ne, cudaMemcpyHostToDevice, s.st));
                h2d_bytes += (int64_t)sizeof(int32_t) * ne;
            }
            if (chunk8) {
                uint32_t* const pin_lut = reinterpret_cast<uint32_t*>(s.pin_codes + (in_pageable ? sizeof(float) : 1) * (size_t)max_nnz);
                memcpy(pin_lut, dict->lut, sizeof(uint32_t) * ValueDict::kMax);
                CKB(cudaMemcpyAsync(d_codes, s.pin_codes, (size_t)ne, cudaMemcpyHostToDevice, s.st));
                CKB(cudaMemcpyAsync(d_lut, pin_lut, sizeof(uint32_t) * ValueDict::kMax, cudaMemcpyHostToDevice, s.st));
                h2d_bytes += ne + (int64_t)sizeof(uint32_t) * ValueDict::kMax;
            } else if (in_pageable) {
                copy_bytes(s.pin_codes, h_data + e0, sizeof(float) * (size_t)ne, n_threads);
                CKB(cudaMemcpyAsync(d_vals, s.pin_codes, sizeof(float) * ne, cudaMemcpyHostToDevice, s.st));
                h2d_bytes += (int64_t)sizeof(float) * ne;
            } else {
                CKB(cudaMemcpyAsync(d_vals, h_data + e0, sizeof(float) * ne, cudaMemcpyHostToDevice, s.st));
                h2d_bytes += (int64_t)sizeof(float) * ne;
            }
            if (staged) {
                CKB(cudaEventRecord(s.idx_done, s.st));
                s.idx_pending = true;
            }
        }
        // the kernel indexes indices/data with the absolute offsets of h_indptr: shift the bases
        const bool chunk16 = idx16 && ne > 0;
        if (chunk8 && chunk16)
            rc = aucell_csr16_lut8_f32(p, d_indptr, reinterpret_cast<uint16_t*>(d_indices) - e0, d_codes - e0, d_lut,
                                       d_vals - e0, nc, d_auc, d_nnz, s.st);
        else if (chunk8)
            rc = aucell_csr_lut8_f32(p, d_indptr, d_indices - e0, d_codes - e0, d_lut, d_vals - e0, nc, d_auc, d_nnz, s.st);
        else if (chunk16)
            rc = aucell_csr16_f32(p, d_indptr, reinterpret_cast<uint16_t*>(d_indices) - e0, d_vals - e0, nc, d_auc,
                                  d_nnz, s.st);
        else
            rc = aucell_csr_f32(p, d_indptr, d_indices - e0, d_vals - e0, nc, d_auc, d_nnz, s.st);
        if (rc) break;
        // Preparing a chunk only pays while the host does it faster than the link would carry the next plainer format
        // (PCIe 5 x16 sustains ~50 GB/s: 0.10 ns per entry at 5 bytes, 0.12 at 6, 0.154 at 8).  Measured over three
        // chunks per format, the call steps down: both narrowings (3 bytes per entry, the host touches 11) -> byte codes
        // only (5, touches 5: the indices go up by DMA as they are) -> uint16 indices only (6, touches 6) -> as is (8).
        // (a pageable caller's arrays are copied by the host in every format: the smallest upload stays)
        if (!in_pageable && narrow_chunks == 3 && narrow_n > (1 << 22)) {
            const double t = narrow_ns / (double)narrow_n;
            bool changed = false;
            if (val8 && idx16) {
                if (t > 0.10 && !idx16_forced) { idx16 = false; changed = true; }
            } else if (val8) {
                if (t > 0.12 && !val8_forced) { val8 = false; idx16 = idx16_allowed; changed = true; }
            } else if (idx16) {
                if (t > 0.154 && !idx16_forced) { idx16 = false; changed = true; }
            }
            if (changed) { narrow_ns = 0.0; narrow_n = 0; narrow_chunks = 0; }
        }
        // ... and when the pipeline as a whole runs well below what the link carries (several GPUs behind one host
        // memory system: measured 8 ranks on 32 cores), narrowing only adds host memory traffic: stop it.  The enqueue
        // loop is throttled by the slots, so its period is the pipeline's.
        ++chunk_no;
        if (chunk_no == kHostSlots) rate_t0 = std::chrono::steady_clock::now();
        if (chunk_no > kHostSlots) rate_n += ne;
        if (idx16 && !idx16_forced && !val8 && !in_pageable && chunk_no == 3 * kHostSlots && rate_n > (1 << 24)) {
            const double ns = std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - rate_t0).count();
            if (ns > (double)rate_n / 4.5) idx16 = false;        // < 4.5 G entries/s
        }
        if (out_pageable) {
            if (drainer.joinable()) drainer.join();
            if (drain_err != cudaSuccess) { rc = fail_cuda(drain_err, "cudaStreamSynchronize", __LINE__); break; }
            CKB(cudaMemcpyAsync(s.pin_out, d_auc, sizeof(double) * nc * R, cudaMemcpyDeviceToHost, s.st));
            if (h_out_nnz) CKB(cudaMemcpyAsync(s.pin_out + out_block, d_nnz, sizeof(int32_t) * nc, cudaMemcpyDeviceToHost, s.st));
            s.out_c0 = c0;
            s.out_nc = nc;
        } else {
            CKB(cudaMemcpyAsync(h_out_auc + c0 * R, d_auc, sizeof(double) * nc * R, cudaMemcpyDeviceToHost, s.st));
            if (h_out_nnz) CKB(cudaMemcpyAsync(h_out_nnz + c0, d_nnz, sizeof(int32_t) * nc, cudaMemcpyDeviceToHost, s.st));
        }
    }
    if (drainer.joinable()) drainer.join();
    if (drain_err != cudaSuccess && rc == AUCELL_OK) rc = fail_cuda(drain_err, "cudaStreamSynchronize", __LINE__);
    for (auto& th : prefault) th.join();
    for (auto& s : host_slots) {
        cudaError_t e = cudaStreamSynchronize(s.st);
        if (e != cudaSuccess && rc == AUCELL_OK) rc = fail_cuda(e, "cudaStreamSynchronize", __LINE__);
        if (rc == AUCELL_OK && out_pageable) drain_out(s);
        s.out_nc = 0;
    }
    p->last_h2d_bytes = h2d_bytes;
    p->last_d2h_bytes = d2h_bytes;
    if (rc == AUCELL_OK && total_nnz > (1 << 24) && !in_pageable) {       // (a small call measures nothing)
        p->hint_idx16 = idx16 || !idx16_allowed;
        p->hint_val8 = val8 || !val8_allowed;
    }
    if (const char* dbg = getenv("AUCELL_B200_DEBUG"); dbg && dbg[0] == '1')
        fprintf(stderr, "[aucell_b200] csr host call: %lld cells, %lld entries, host prep %.3f s (%.2f G entries/s, %d threads), "
                        "idx16 %d val8 %d, h2d %.2f GB d2h %.2f GB; %.3f s in all (staging %.3f s, copy-out %.3f s, pageable in %d out %d, "
                        "chunks of %lld cells)\n", (long long)n_cells, (long long)total_nnz, narrow_ns * 1e-9,
                narrow_ns > 0 ? (double)narrow_n / narrow_ns : 0.0, n_threads, (int)idx16, (int)val8, h2d_bytes * 1e-9,
                d2h_bytes * 1e-9, std::chrono::duration<double>(std::chrono::steady_clock::now() - t_call0).count(), t_alloc,
                t_drain, (int)in_pageable, (int)out_pageable, (long long)chunk_cells);
    return rc;
}

int aucell_csr_is_canonical(const int64_t* h_indptr, const int32_t* h_indices, int64_t n_cells, int64_t n_genes) {
    if (!h_indptr || n_cells < 0) return fail(AUCELL_ERR_INVALID, "NULL pointer");
    if (n_cells == 0) return 1;
    if (!h_indices && h_indptr[n_cells] > h_indptr[0]) return fail(AUCELL_ERR_INVALID, "NULL pointer");
    const int threads = std::max(1, std::min(host_threads(), (int)std::min<int64_t>(n_cells, 64)));
    std::vector<int> bad((size_t)threads, 0);
    auto work = [&](int t) {
        const int64_t r0 = n_cells * t / threads, r1 = n_cells * (t + 1) / threads;
        int b = 0;
        for (int64_t r = r0; r < r1 && !b; ++r) {
            const int64_t s = h_indptr[r], e = h_indptr[r + 1];
            if (e < s) { b = 1; break; }
            if (e == s) continue;
            int32_t prev = h_indices[s];
            int ok = prev >= 0 ? 1 : 0;
            for (int64_t i = s + 1; i < e; ++i) {          // branch-free inner loop: the scan is memory bound
                const int32_t c = h_indices[i];
                ok &= (c > prev) ? 1 : 0;
                prev = c;
            }
            if (!ok || (int64_t)prev >= n_genes) b = 1;
        }
        bad[(size_t)t] = b;
    };
    parallel_tasks(threads, work);
    for (int b : bad) if (b) return 0;
    return 1;
}

int aucell_dense_strided_host(const aucell_plan_t* p, const void* h_expr, int is_f64, int64_t n_cells,
                              int64_t cell_stride, int64_t gene_stride, double* h_out_auc, int32_t* h_out_nnz) {
    if (!p) return fail(AUCELL_ERR_INVALID, "plan is NULL");
    if (p->n_regulons <= 0) return fail(AUCELL_ERR_INVALID, "plan has no regulons");
    if (n_cells < 0) return fail(AUCELL_ERR_INVALID, "n_cells < 0");
    if (n_cells == 0) return AUCELL_OK;
    if (!h_expr || !h_out_auc) return fail(AUCELL_ERR_INVALID, "NULL pointer");
    const bool cell_major = gene_stride == 1 && cell_stride >= p->n_genes;
    const bool gene_major = cell_stride == 1 && gene_stride >= n_cells;
    if (!cell_major && !gene_major) return fail(AUCELL_ERR_INVALID, "one of the two strides must be 1 and the other span the matrix");
    if (!p->fast_ok || !p->fast_enabled) return AUCELL_NOT_SPARSE;
    DeviceGuard dg(p->device);
    if (!dg.ok) return fail(AUCELL_ERR_CUDA, "cudaSetDevice failed");
    const int rc = is_f64 ? dense_host_via_csr<double>(p, static_cast<const double*>(h_expr), n_cells, cell_stride, gene_stride,
                                                       h_out_auc, h_out_nnz)
                          : dense_host_via_csr<float>(p, static_cast<const float*>(h_expr), n_cells, cell_stride, gene_stride,
                                                      h_out_auc, h_out_nnz);
    return rc == kNotCompacted ? AUCELL_NOT_SPARSE : rc;
}

int aucell_host_dense_to_csr(const void* h_expr, int is_f64, int64_t n_cells, int64_t n_genes, int64_t cell_stride,
                             int64_t gene_stride, int64_t* h_indptr, int32_t* h_indices, float* h_values, int64_t capacity,
                             int64_t* out_nnz) {
    if (n_cells < 0 || n_genes < 1 || n_genes > INT32_MAX) return fail(AUCELL_ERR_INVALID, "bad n_cells / n_genes");
    if (!h_indptr || !out_nnz || (n_cells > 0 && !h_expr)) return fail(AUCELL_ERR_INVALID, "NULL pointer");
    const bool cell_major = gene_stride == 1 && cell_stride >= n_genes;
    const bool gene_major = cell_stride == 1 && gene_stride >= n_cells;
    if (!cell_major && !gene_major) return fail(AUCELL_ERR_INVALID, "one of the two strides must be 1 and the other span the matrix");
    *out_nnz = 0;
    h_indptr[0] = 0;
    if (n_cells == 0) return AUCELL_OK;
    const int threads = (int)std::max<int64_t>(1, std::min<int64_t>(host_threads(), n_cells));
    // density of a sample: 64 evenly spaced cells x up to 4096 evenly spaced genes (only sizes the first buffers)
    const int64_t pc = std::min<int64_t>(n_cells, 64), pg = std::min<int64_t>(n_genes, 4096);
    int64_t nz = 0;
    for (int64_t a = 0; a < pc; ++a)
        for (int64_t b = 0; b < pg; ++b) {
            const int64_t off = (n_cells * a / pc) * cell_stride + (n_genes * b / pg) * gene_stride;
            nz += (is_f64 ? is_zero_bits(static_cast<const double*>(h_expr) + off) : is_zero_bits(static_cast<const float*>(h_expr) + off)) ? 0 : 1;
        }
    const double density = (double)nz / (double)(pc * pg);
    std::vector<RawBuf<int32_t>> t_idx((size_t)threads);
    std::vector<RawBuf<float>> t_val((size_t)threads);
    std::vector<int64_t> indptr;
    RawBuf<int32_t> indices;
    RawBuf<float> values;
    const bool ok = is_f64 ? compact_block<double>(static_cast<const double*>(h_expr), cell_stride, gene_stride, 0, n_cells, n_genes,
                                                   threads, density, t_idx, t_val, indptr, indices, values)
                           : compact_block<float>(static_cast<const float*>(h_expr), cell_stride, gene_stride, 0, n_cells, n_genes,
                                                  threads, density, t_idx, t_val, indptr, indices, values);
    if (!ok) return AUCELL_NOT_SPARSE;
    const int64_t nnz = indptr[(size_t)n_cells];
    *out_nnz = nnz;
    memcpy(h_indptr, indptr.data(), sizeof(int64_t) * (size_t)(n_cells + 1));
    if (nnz > capacity) return fail(AUCELL_ERR_NOMEM, "capacity of the output arrays is smaller than the number of non-zero values");
    if (nnz > 0 && (!h_indices || !h_values)) return fail(AUCELL_ERR_INVALID, "NULL pointer");
    if (nnz > 0) {
        memcpy(h_indices, indices.data(), sizeof(int32_t) * (size_t)nnz);
        memcpy(h_values, values.data(), sizeof(float) * (size_t)nnz);
    }
    return AUCELL_OK;
}

int aucell_cast_f64_to_f32(const double* h_src, float* h_dst, int64_t n) {
    if (n < 0 || (n > 0 && (!h_src || !h_dst))) return fail(AUCELL_ERR_INVALID, "NULL pointer");
    const int threads = n < ((int64_t)1 << 20) ? 1 : host_threads();
    std::vector<int> bad((size_t)threads, 0);
    auto work = [&](int t) {
        const int64_t b = n * t / threads, e = n * (t + 1) / threads;
        int inexact = 0;
        for (int64_t i = b; i < e; ++i) {
            const double v = h_src[i];
            const float f = (float)v;
            h_dst[i] = f;
            inexact |= ((double)f == v || v != v) ? 0 : 1;      // (NaN stays NaN)
        }
        bad[(size_t)t] = inexact;
    };
    parallel_tasks(threads, work);
    for (int b : bad) if (b) return 0;
    return 1;
}

int aucell_cast_i64_to_i32(const int64_t* h_src, int32_t* h_dst, int64_t n) {
    if (n < 0 || (n > 0 && (!h_src || !h_dst))) return fail(AUCELL_ERR_INVALID, "NULL pointer");
    const int threads = n < ((int64_t)1 << 20) ? 1 : host_threads();
    std::vector<int> bad((size_t)threads, 0);
    auto work = [&](int t) {
        const int64_t b = n * t / threads, e = n * (t + 1) / threads;
        uint64_t high = 0;
        for (int64_t i = b; i < e; ++i) {
            const int64_t v = h_src[i];
            h_dst[i] = (int32_t)v;
            high |= (uint64_t)v >> 31;                           // negative or >= 2^31
        }
        bad[(size_t)t] = high != 0 ? 1 : 0;
    };
    parallel_tasks(threads, work);
    for (int b : bad) if (b) return 0;
    return 1;
}

int aucell_b200_release_staging(int device) {
    std::vector<DeviceStaging*> which;
    {
        std::lock_guard<std::mutex> lk(staging_map_mu());
        for (auto& kv : staging_map())
            if (device < 0 || kv.first == device) which.push_back(kv.second.get());
    }
    int cur = 0;
    cudaGetDevice(&cur);
    for (DeviceStaging* st : which) {
        std::lock_guard<std::mutex> lk(st->mu);
        free_slots(st->slots);
    }
    cudaSetDevice(cur);
    cudaGetLastError();
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
    // Sparse rows (what expression matrices are): squeeze the zeros out on the host and continue as a CSR call.
    {
        const char* envc = getenv("AUCELL_B200_HOST_COMPACT");
        if (p->fast_ok && p->fast_enabled && chunk_cells <= 0 && !(envc && envc[0] == '0')) {
            const int rc = dense_host_via_csr<float>(p, h_expr, n_cells, row_stride, 1, h_out_auc, h_out_nnz);
            if (rc != kNotCompacted) return rc;
        }
    }
    DeviceStaging& staging = staging_of(p->device);
    std::lock_guard<std::mutex> staging_lock(staging.mu);
    std::lock_guard<std::mutex> host_lock(p->host_mu);
    std::vector<aucell_plan::HostSlot>& host_slots = staging.slots;
    if (chunk_cells <= 0)
        chunk_cells = std::max<int64_t>(256, (int64_t)(128.0 * 1024 * 1024 / (4.0 * (double)row_stride)));
    chunk_cells = std::min(chunk_cells, n_cells);
    const size_t o_vals = 0;
    const size_t o_auc = align256(o_vals + sizeof(float) * chunk_cells * row_stride);
    const size_t o_nnz = align256(o_auc + sizeof(double) * chunk_cells * R);
    const size_t bytes = align256(o_nnz + sizeof(int32_t) * chunk_cells);
    int rc = ensure_host_slots(host_slots, bytes);
    if (rc) return rc;
    int k = 0;
    p->last_h2d_bytes = 0;
    p->last_d2h_bytes = 0;
    for (int64_t c0 = 0; c0 < n_cells; c0 += chunk_cells, k = (k + 1) % kHostSlots) {
        const aucell_plan::HostSlot& s = host_slots[k];
        float* d_vals = reinterpret_cast<float*>(s.buf + o_vals);
        double* d_auc = reinterpret_cast<double*>(s.buf + o_auc);
        int32_t* d_nnz = h_out_nnz ? reinterpret_cast<int32_t*>(s.buf + o_nnz) : nullptr;
        const int64_t nc = std::min(n_cells, c0 + chunk_cells) - c0;
        CKB(cudaMemcpyAsync(d_vals, h_expr + c0 * row_stride, sizeof(float) * nc * row_stride, cudaMemcpyHostToDevice, s.st));
        p->last_h2d_bytes += (int64_t)sizeof(float) * nc * row_stride;
        p->last_d2h_bytes += (int64_t)sizeof(double) * nc * R + (h_out_nnz ? (int64_t)sizeof(int32_t) * nc : 0);
        rc = aucell_dense_f32(p, d_vals, nc, row_stride, d_auc, d_nnz, s.st);
        if (rc) break;
        CKB(cudaMemcpyAsync(h_out_auc + c0 * R, d_auc, sizeof(double) * nc * R, cudaMemcpyDeviceToHost, s.st));
        if (h_out_nnz) CKB(cudaMemcpyAsync(h_out_nnz + c0, d_nnz, sizeof(int32_t) * nc, cudaMemcpyDeviceToHost, s.st));
    }
    for (auto& s : host_slots) {
        cudaError_t e = cudaStreamSynchronize(s.st);
        if (e != cudaSuccess && rc == AUCELL_OK) rc = fail_cuda(e, "cudaStreamSynchronize", __LINE__);
    }
    return rc;
}

}  // extern "C"

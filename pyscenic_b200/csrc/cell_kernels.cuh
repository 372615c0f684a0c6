// cell_kernels.cuh -- one CTA per cell: ranking with seeded tie order, fused with the recovery-curve AUC.
//
// Reference path (aertslab/pySCENIC 0.12.1):
//   create_rankings                      src/pyscenic/aucell.py:25-51   (pandas sample + rank)
//   aucell4r / _enrichment               src/pyscenic/aucell.py:78-182
//   ctxcore.recovery.enrichment4cells -> aucs -> auc2d -> weighted_auc1d   (third-party)
//
// Per-cell algorithm (all in shared memory; the cells x genes ranking matrix is never written by the fused
// kernels):
//   1. INGEST   every value whose key differs from KEY_ZERO becomes an "explicit entry" (key, shuffled
//               position).  Zeros -- stored or implicit -- stay implicit: a zero at shuffled position p has rank
//               n_gt + Z(p), Z(p) = #zeros at positions < p = p - #explicit entries at positions < p.
//   2. SORT     explicit entries by (key, position): positives first (descending value), then negatives, NaN
//               last.  Sorted index i is the rank for i < n_gt and i + n_zero after the zero block.
//   3. TABLE    tbl[p] = max(0, rank_cutoff - rank(p)) for the genes ranked below the cutoff (u16/u32 smem).
//   4. GATHER   per regulon, sum_g w_g * tbl[pos_g] with a fixed-shape warp reduction, then / max_auc:
//               the closed form of weighted_auc1d's  sum(diff(x) * cumsum(w)) / max_auc.
//               Unweighted regulons accumulate in integers, so the result is bit-identical to the reference.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "keys.cuh"

namespace aucell {

constexpr int kBlock = 256;          // threads per CTA
constexpr int kWarps = kBlock / 32;

enum InputKind { IN_DENSE = 0, IN_CSR = 1 };
enum OutputKind { OUT_RANK = 0, OUT_AUC = 1 };

struct RegulonsDev {
    int n_regulons;
    const int32_t* indptr;   // [R+1] into pos/weight
    const void* pos;         // PosT[nnz] shuffled positions, bank-staggered order within each regulon
    const double* weight;    // [nnz], nullptr when unweighted
    const double* max_auc;   // [R]
    const uint8_t* valid;    // [R]
};

template <typename ValT>
struct CellArgs {
    // input
    const ValT* dense;       // IN_DENSE: [n_cells x row_stride]
    int64_t row_stride;
    const int64_t* indptr;   // IN_CSR
    const int32_t* indices;
    const ValT* data;
    int64_t n_cells;
    int n_genes;
    int n_words;             // ceil(n_genes / 32)
    const int32_t* inv_perm; // [n_genes] gene column -> shuffled position
    int rank_cutoff;
    RegulonsDev reg;
    // output
    uint32_t* out_rank;      // OUT_RANK: [n_cells x n_genes], shuffled column order
    double* out_auc;         // OUT_AUC:  [n_cells x n_regulons]
    int32_t* out_nnz;        // nullable
    // shared-memory carve-up (bytes from the dynamic smem base) and capacities
    int cap;                 // explicit-entry capacity of the smem list (power of two)
    int off_key, off_pos, off_bits, off_wpre, off_tbl;
    // global overflow list for cells with more than `cap` explicit entries: per CTA, scratch_cap entries
    void* scratch_key;
    void* scratch_pos;
    int scratch_cap;         // power of two >= n_genes
};

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// exclusive scan of one int per thread over the CTA; s_warp needs 33 ints.
__device__ __forceinline__ int block_excl_scan(int v, int* s_warp, int& total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    if (w == 0) {
        int x = (lane < kWarps) ? s_warp[lane] : 0;
        int xi = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, xi, o);
            if (lane >= o) xi += t;
        }
        s_warp[lane] = xi - x;
        if (lane == 31) s_warp[32] = xi;
    }
    __syncthreads();
    total = s_warp[32];
    int res = s_warp[w] + incl - v;
    __syncthreads();
    return res;
}

// Bitonic sort of P (power of two) (key, pos) pairs, ascending by key then pos.  Works on shared or global
// memory.  Pairs are unique (positions are), so the result does not depend on the input order.
template <typename KeyT, typename PosT>
__device__ void bitonic_sort_pairs(KeyT* __restrict__ k, PosT* __restrict__ p, int P) {
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < (P >> 1); t += kBlock) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool asc = ((lo & size) == 0);
                const KeyT ka = k[lo], kb = k[hi];
                const PosT pa = p[lo], pb = p[hi];
                const bool gt = (ka > kb) || (ka == kb && pa > pb);
                if (gt == asc) {
                    k[lo] = kb; k[hi] = ka;
                    p[lo] = pb; p[hi] = pa;
                }
            }
        }
    }
    __syncthreads();
}

// Append explicit entries of one cell to (lk, lp) (capacity cap; entries beyond it are dropped but counted).
// s_cnt[0] = number of explicit entries, s_cnt[1] = number with key < KEY_ZERO (values > 0).
template <typename ValT, typename PosT, int IN>
__device__ void ingest_cell(const CellArgs<ValT>& a, int64_t cell, typename KeyOf<ValT>::type* lk, PosT* lp,
                            int cap, int* s_cnt) {
    using KT = KeyOf<ValT>;
    using KeyT = typename KT::type;
    if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int64_t begin, count;
    const ValT* vals;
    const int32_t* cols = nullptr;
    if (IN == IN_DENSE) {
        begin = 0; count = a.n_genes;
        vals = a.dense + cell * a.row_stride;
    } else {
        begin = a.indptr[cell];
        count = a.indptr[cell + 1] - begin;
        vals = a.data + begin;
        cols = a.indices + begin;
    }
    for (int64_t base = 0; base < count; base += kBlock) {
        const int64_t i = base + threadIdx.x;
        const bool in = i < count;
        KeyT key = KT::KEY_ZERO;
        if (in) key = KT::make(vals[i]);
        const bool pred = in && key != KT::KEY_ZERO;
        const unsigned m = __ballot_sync(0xffffffffu, pred);
        if (m) {
            const unsigned mgt = __ballot_sync(0xffffffffu, pred && key < KT::KEY_ZERO);
            int wbase = 0;
            const int leader = __ffs(m) - 1;
            if (lane == leader) {
                wbase = atomicAdd(&s_cnt[0], __popc(m));
                if (mgt) atomicAdd(&s_cnt[1], __popc(mgt));
            }
            wbase = __shfl_sync(0xffffffffu, wbase, leader);
            if (pred) {
                const int idx = wbase + __popc(m & lanemask_lt());
                if (idx < cap) {
                    const int g = (IN == IN_DENSE) ? (int)i : cols[i];
                    lk[idx] = key;
                    lp[idx] = (PosT)a.inv_perm[g];
                }
            }
        }
    }
    __syncthreads();
}

// One CTA per cell (grid-stride).  See the file header for the algorithm.
template <typename ValT, typename PosT, int IN, int OUT, bool WEIGHTED>
__global__ void __launch_bounds__(kBlock) aucell_cell_kernel(const CellArgs<ValT> a) {
    using KT = KeyOf<ValT>;
    using KeyT = typename KT::type;
    extern __shared__ __align__(16) unsigned char smem[];
    KeyT* const s_key = reinterpret_cast<KeyT*>(smem + a.off_key);
    PosT* const s_pos = reinterpret_cast<PosT*>(smem + a.off_pos);
    uint32_t* const s_bits = reinterpret_cast<uint32_t*>(smem + a.off_bits);
    uint32_t* const s_wpre = reinterpret_cast<uint32_t*>(smem + a.off_wpre);
    PosT* const s_tbl = reinterpret_cast<PosT*>(smem + a.off_tbl);
    __shared__ int s_cnt[2];
    __shared__ int s_scan[33];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_genes = a.n_genes;
    const int cutoff = a.rank_cutoff;

    if (OUT == OUT_AUC) {
        for (int i = tid; i < n_genes; i += kBlock) s_tbl[i] = 0;
    }
    __syncthreads();

    for (int64_t cell = blockIdx.x; cell < a.n_cells; cell += gridDim.x) {
        // ---- 1. ingest ------------------------------------------------------------------------------
        KeyT* lk = s_key;
        PosT* lp = s_pos;
        int lcap = a.cap;
        ingest_cell<ValT, PosT, IN>(a, cell, lk, lp, lcap, s_cnt);
        int m = s_cnt[0];
        if (m > lcap) {  // rare: more explicit entries than the smem list holds -> global scratch list
            __syncthreads();
            lk = reinterpret_cast<KeyT*>(a.scratch_key) + (size_t)blockIdx.x * a.scratch_cap;
            lp = reinterpret_cast<PosT*>(a.scratch_pos) + (size_t)blockIdx.x * a.scratch_cap;
            lcap = a.scratch_cap;
            ingest_cell<ValT, PosT, IN>(a, cell, lk, lp, lcap, s_cnt);
            m = s_cnt[0];
        }
        const int n_gt = s_cnt[1];
        const int n_zero = n_genes - m;
        if (a.out_nnz != nullptr && tid == 0) a.out_nnz[cell] = m;

        // ---- 2. sort explicit entries by (key, shuffled position) -------------------------------------
        int P = 1;
        while (P < m) P <<= 1;
        if (P < 2) P = 2;
        for (int i = m + tid; i < P; i += kBlock) {
            lk[i] = KT::KEY_NAN;
            lp[i] = (PosT)~(PosT)0;
        }
        bitonic_sort_pairs<KeyT, PosT>(lk, lp, P);  // starts and ends with __syncthreads

        // zeros are needed when the whole ranking is written, or when they reach below the cutoff
        const bool need_zero = (OUT == OUT_RANK) || (n_gt < cutoff);
        if (need_zero) {
            for (int w = tid; w < a.n_words; w += kBlock) s_bits[w] = 0u;
            __syncthreads();
            for (int i = tid; i < m; i += kBlock) {
                const unsigned p = (unsigned)lp[i];
                atomicOr(&s_bits[p >> 5], 1u << (p & 31));
            }
            __syncthreads();
            // exclusive prefix of explicit-entry counts per bitmap word
            const int per = (a.n_words + kBlock - 1) / kBlock;
            const int w0 = tid * per;
            const int w1 = min(w0 + per, a.n_words);
            int local = 0;
            for (int w = w0; w < w1; ++w) local += __popc(s_bits[w]);
            int total;
            int run = block_excl_scan(local, s_scan, total);
            for (int w = w0; w < w1; ++w) {
                s_wpre[w] = (uint32_t)run;
                run += __popc(s_bits[w]);
            }
            __syncthreads();
        }

        if (OUT == OUT_RANK) {
            uint32_t* out = a.out_rank + cell * (int64_t)n_genes;
            for (int i = tid; i < m; i += kBlock)
                out[(unsigned)lp[i]] = (uint32_t)(i < n_gt ? i : i + n_zero);
            for (int w = tid; w < a.n_words; w += kBlock) {
                uint32_t zm = ~s_bits[w];
                const int rem = n_genes - (w << 5);
                if (rem < 32) zm &= (1u << rem) - 1u;
                uint32_t z = (uint32_t)((w << 5) - (int)s_wpre[w] + n_gt);
                while (zm) {
                    const int b = __ffs(zm) - 1;
                    out[(w << 5) + b] = z++;
                    zm &= zm - 1;
                }
            }
            __syncthreads();
        } else {
            // ---- 3. table of (cutoff - rank) for the genes ranked below the cutoff ----------------------
            const int top_pos = min(n_gt, cutoff);                   // positives: rank == sorted index
            for (int i = tid; i < top_pos; i += kBlock) s_tbl[(unsigned)lp[i]] = (PosT)(cutoff - i);
            const int q_zero = cutoff - n_gt;                        // zeros wanted (if > 0)
            const int q_neg = cutoff - n_gt - n_zero;                // negatives wanted (if > 0)
            const int neg_end = q_neg > 0 ? min(m, n_gt + q_neg) : n_gt;
            if (q_zero > 0) {
                for (int w = tid; w < a.n_words; w += kBlock) {
                    int z = (w << 5) - (int)s_wpre[w];               // zeros before this word
                    if (z >= q_zero) continue;
                    uint32_t zm = ~s_bits[w];
                    const int rem = n_genes - (w << 5);
                    if (rem < 32) zm &= (1u << rem) - 1u;
                    while (zm && z < q_zero) {
                        const int b = __ffs(zm) - 1;
                        s_tbl[(w << 5) + b] = (PosT)(q_zero - z);
                        ++z;
                        zm &= zm - 1;
                    }
                }
                for (int i = n_gt + tid; i < neg_end; i += kBlock)
                    s_tbl[(unsigned)lp[i]] = (PosT)(cutoff - (i + n_zero));
            }
            __syncthreads();

            // ---- 4. gather: one warp per regulon ------------------------------------------------------
            const PosT* __restrict__ rpos = reinterpret_cast<const PosT*>(a.reg.pos);
            double* out = a.out_auc + cell * (int64_t)a.reg.n_regulons;
            for (int r = warp; r < a.reg.n_regulons; r += kWarps) {
                const int b = a.reg.indptr[r], e = a.reg.indptr[r + 1];
                double res;
                if (WEIGHTED) {
                    double acc = 0.0;
                    for (int i = b + lane; i < e; i += 32)
                        acc = fma((double)s_tbl[rpos[i]], a.reg.weight[i], acc);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                    res = acc;
                } else {
                    unsigned long long acc = 0ull;
                    for (int i = b + lane; i < e; i += 32) acc += (unsigned long long)s_tbl[rpos[i]];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                    res = (double)acc;
                }
                if (lane == 0) out[r] = a.reg.valid[r] ? __ddiv_rn(res, a.reg.max_auc[r]) : 0.0;
            }
            __syncthreads();

            // ---- undo the table writes ----------------------------------------------------------------
            for (int i = tid; i < top_pos; i += kBlock) s_tbl[(unsigned)lp[i]] = 0;
            if (q_zero > 0) {
                for (int w = tid; w < a.n_words; w += kBlock) {
                    int z = (w << 5) - (int)s_wpre[w];
                    if (z >= q_zero) continue;
                    uint32_t zm = ~s_bits[w];
                    const int rem = n_genes - (w << 5);
                    if (rem < 32) zm &= (1u << rem) - 1u;
                    while (zm && z < q_zero) {
                        const int b = __ffs(zm) - 1;
                        s_tbl[(w << 5) + b] = 0;
                        ++z;
                        zm &= zm - 1;
                    }
                }
                for (int i = n_gt + tid; i < neg_end; i += kBlock) s_tbl[(unsigned)lp[i]] = 0;
            }
            __syncthreads();
        }
    }
}

// aucell4r path: AUCs from an existing ranking matrix (reference aucell.py:98-182; ctxcore auc2d).
// The table is filled straight from the ranking row, then the same gather runs.
template <typename PosT, bool WEIGHTED>
__global__ void __launch_bounds__(kBlock) aucell_from_rankings_kernel(const uint32_t* __restrict__ rnk,
                                                                      int64_t n_cells, int64_t row_stride,
                                                                      int n_genes, int cutoff, RegulonsDev reg,
                                                                      double* __restrict__ out_auc) {
    extern __shared__ __align__(16) unsigned char smem[];
    PosT* const s_tbl = reinterpret_cast<PosT*>(smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const PosT* __restrict__ rpos = reinterpret_cast<const PosT*>(reg.pos);
    for (int64_t cell = blockIdx.x; cell < n_cells; cell += gridDim.x) {
        const uint32_t* row = rnk + cell * row_stride;
        for (int j = tid; j < n_genes; j += kBlock) {
            const uint32_t r = row[j];
            s_tbl[j] = (PosT)(r < (uint32_t)cutoff ? (uint32_t)cutoff - r : 0u);
        }
        __syncthreads();
        double* out = out_auc + cell * (int64_t)reg.n_regulons;
        for (int r = warp; r < reg.n_regulons; r += kWarps) {
            const int b = reg.indptr[r], e = reg.indptr[r + 1];
            double res;
            if (WEIGHTED) {
                double acc = 0.0;
                for (int i = b + lane; i < e; i += 32) acc = fma((double)s_tbl[rpos[i]], reg.weight[i], acc);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                res = acc;
            } else {
                unsigned long long acc = 0ull;
                for (int i = b + lane; i < e; i += 32) acc += (unsigned long long)s_tbl[rpos[i]];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                res = (double)acc;
            }
            if (lane == 0) out[r] = reg.valid[r] ? __ddiv_rn(res, reg.max_auc[r]) : 0.0;
        }
        __syncthreads();
    }
}

}  // namespace aucell

// cell_kernels.cuh -- shared device helpers and the create_rankings kernel (full ranking rows).
//
// Reference: create_rankings  src/pyscenic/aucell.py:25-51
//     ex_mtx.sample(frac=1, axis=1, random_state=seed).rank(axis=1, ascending=False, method="first",
//     na_option="bottom").astype(uint32) - 1
//
// One CTA per cell.  Every value whose key differs from KEY_ZERO becomes an "explicit entry" (key, shuffled
// position); explicit entries are sorted by (key, position) with a bitonic network (shared memory, or global
// scratch for cells with more explicit entries than the shared list holds).  Sorted index i is the rank for the
// positives (i < n_gt) and i + n_zero for negatives / NaN.  Zeros -- stored or implicit -- are ranked analytically:
// a zero at shuffled position p has rank n_gt + Z(p), Z(p) = #zeros at positions < p, from a bitmap of the
// occupied positions and a prefix popcount.  The fused AUCell kernel lives in fused_kernel.cuh.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "keys.cuh"

namespace aucell {

constexpr int kBlock = 256;          // threads per CTA
constexpr int kWarps = kBlock / 32;

enum InputKind { IN_DENSE = 0, IN_CSR = 1 };

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
    // output
    uint32_t* out_rank;      // [n_cells x n_genes], shuffled column order
    int32_t* out_nnz;        // nullable
    // shared-memory carve-up (bytes from the dynamic smem base) and capacities
    int cap;                 // explicit-entry capacity of the smem list (power of two)
    int off_key, off_pos, off_bits, off_wpre;
    // global overflow list for cells with more than `cap` explicit entries: per CTA, scratch_cap entries
    void* scratch_key;
    void* scratch_pos;
    int scratch_cap;         // power of two >= n_genes
};

template <typename T> __device__ __forceinline__ T tmin(T a, T b) { return a < b ? a : b; }
template <typename T> __device__ __forceinline__ T tmax(T a, T b) { return a > b ? a : b; }
__device__ __forceinline__ void smem_atomic_min(uint32_t* p, uint32_t v) { atomicMin(p, v); }
__device__ __forceinline__ void smem_atomic_max(uint32_t* p, uint32_t v) { atomicMax(p, v); }
__device__ __forceinline__ void smem_atomic_min(uint64_t* p, uint64_t v) {
    atomicMin(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v);
}
__device__ __forceinline__ void smem_atomic_max(uint64_t* p, uint64_t v) {
    atomicMax(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v);
}

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
                            int cap, int* s_cnt, typename KeyOf<ValT>::type* s_kmm) {
    using KT = KeyOf<ValT>;
    using KeyT = typename KT::type;
    if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
    if (threadIdx.x == 0) { s_kmm[0] = KT::KEY_ZERO; s_kmm[1] = 0; }
    KeyT kmin = KT::KEY_ZERO, kmax = 0;   // over positive values (keys below KEY_ZERO)
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
        if (pred && key < KT::KEY_ZERO) { kmin = tmin(kmin, key); kmax = tmax(kmax, key); }
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
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        kmin = tmin(kmin, (KeyT)__shfl_xor_sync(0xffffffffu, kmin, o));
        kmax = tmax(kmax, (KeyT)__shfl_xor_sync(0xffffffffu, kmax, o));
    }
    if (lane == 0 && kmin < KT::KEY_ZERO) {
        smem_atomic_min(&s_kmm[0], kmin);
        smem_atomic_max(&s_kmm[1], kmax);
    }
    __syncthreads();
}

// create_rankings: one CTA per cell (grid-stride), full uint32 ranking row in shuffled column order.
template <typename ValT, typename PosT, int IN>
__global__ void __launch_bounds__(kBlock) aucell_rank_kernel(const CellArgs<ValT> a) {
    using KT = KeyOf<ValT>;
    using KeyT = typename KT::type;
    extern __shared__ __align__(16) unsigned char smem[];
    KeyT* const s_key = reinterpret_cast<KeyT*>(smem + a.off_key);
    PosT* const s_pos = reinterpret_cast<PosT*>(smem + a.off_pos);
    uint32_t* const s_bits = reinterpret_cast<uint32_t*>(smem + a.off_bits);
    uint32_t* const s_wpre = reinterpret_cast<uint32_t*>(smem + a.off_wpre);
    __shared__ int s_cnt[2];
    __shared__ int s_scan[33];
    __shared__ KeyT s_kmm[2];

    const int tid = threadIdx.x;
    const int n_genes = a.n_genes;

    for (int64_t cell = blockIdx.x; cell < a.n_cells; cell += gridDim.x) {
        // ---- 1. ingest ------------------------------------------------------------------------------
        KeyT* lk = s_key;
        PosT* lp = s_pos;
        int lcap = a.cap;
        ingest_cell<ValT, PosT, IN>(a, cell, lk, lp, lcap, s_cnt, s_kmm);
        int m = s_cnt[0];
        if (m > lcap) {  // more explicit entries than the smem list holds -> global scratch list
            __syncthreads();
            lk = reinterpret_cast<KeyT*>(a.scratch_key) + (size_t)blockIdx.x * a.scratch_cap;
            lp = reinterpret_cast<PosT*>(a.scratch_pos) + (size_t)blockIdx.x * a.scratch_cap;
            lcap = a.scratch_cap;
            ingest_cell<ValT, PosT, IN>(a, cell, lk, lp, lcap, s_cnt, s_kmm);
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

        // ---- 3. bitmap of occupied positions + exclusive prefix of their counts per word ---------------
        for (int w = tid; w < a.n_words; w += kBlock) s_bits[w] = 0u;
        __syncthreads();
        for (int i = tid; i < m; i += kBlock) {
            const unsigned p = (unsigned)lp[i];
            atomicOr(&s_bits[p >> 5], 1u << (p & 31));
        }
        __syncthreads();
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

        // ---- 4. write the ranking row ---------------------------------------------------------------
        uint32_t* out = a.out_rank + cell * (int64_t)n_genes;
        for (int i = tid; i < m; i += kBlock) out[(unsigned)lp[i]] = (uint32_t)(i < n_gt ? i : i + n_zero);
        for (int p = tid; p < n_genes; p += kBlock) {
            const uint32_t bw = s_bits[p >> 5];
            if (!((bw >> (p & 31)) & 1u)) {
                const int z = p - (int)s_wpre[p >> 5] - __popc(bw & ((1u << (p & 31)) - 1u));
                out[p] = (uint32_t)(n_gt + z);      // coalesced: consecutive threads, consecutive positions
            }
        }
        __syncthreads();
    }
}

}  // namespace aucell

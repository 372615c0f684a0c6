// fused_kernel.cuh -- the AUCell hot path: per-cell ranking fused with the recovery-curve AUC of every regulon
// (and, in RANKS mode, create_rankings: the same ranking written out as full uint32 rows).
//
// Reference: pyscenic.aucell.aucell = aucell4r(create_rankings(exp_mtx, seed), ...)   src/pyscenic/aucell.py:185-213
//            create_rankings  (pandas sample + rank(method="first", ascending=False))  src/pyscenic/aucell.py:25-51
//            ctxcore.recovery.enrichment4cells -> aucs -> auc2d -> weighted_auc1d     (third-party)
//
// One CTA (kF threads) per cell, grid-stride.  The cells x genes ranking matrix is never formed: only genes ranked
// below rank_cutoff contribute, AUC_r = sum_{g in r, rank_g < cutoff} w_g * (cutoff - rank_g) / max_auc_r
// (closed form of weighted_auc1d's sum(diff(x) * cumsum(w)) / max_auc).
//
// Per cell:
//   LOAD     the cell's explicit entries (values whose key != KEY_ZERO) live in REGISTERS, kItems per thread:
//            (key, shuffled position).  Zeros stay implicit (rank n_gt + #zeros at smaller positions).
//   RANK     positives are ranked by a two-level bucket sort on 32-bit shared atomics:
//              level 1: value bins over the cell's own key range (ascending key = descending value);
//                       bins starting at or beyond the cutoff are dropped;
//              level 2: every kept bin is split into as many sub-buckets as it has entries -- a bin with ONE
//                       distinct key (tie group) by shuffled position (positions are uniform -> ~1 entry each), a
//                       bin with several keys (continuous data) by the key's offset inside the bin;
//              exact:   entries are stored in sub-bucket order and each counts the members of its own sub-bucket
//                       that sort before it:  rank = base2[sub] + count.
//            Crowded sub-buckets (e.g. two adjacent values, hundreds of genes each, in one value bin), negatives
//            ranked below the cutoff, or a selection of more than kF * kItems entries send the cell to the general
//            path (bitonic sort in shared memory / global scratch).  Rows with more than kF * kItems explicit entries
//            are streamed ("select mode") and only the bins that start below the cutoff come to registers.
//   SCATTER  for each gene below the cutoff walk the gene-major regulon table T[pos] and add W * (cutoff - rank)
//            into 64-bit integer accumulators in shared memory (32-bit atomics + carry; fixed-point weights).
//            Integer adds commute, so the result does not depend on the order the atomics retire in.
//   ZEROS    when fewer than `cutoff` genes are positive, zeros fill the remaining ranks in position order
//            (bitmap of occupied positions + prefix popcount).
//   OUT      out[cell, r] = (double(acc) * 2^-s) / max_auc[r], coalesced.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "keys.cuh"

namespace aucell {

constexpr int kF = 512;                 // threads per CTA of the fused kernel
constexpr int kFWarps = kF / 32;
constexpr int kItems = 8;               // register-resident entries per thread
constexpr int kCapF = kF * kItems;      // 4096 explicit entries per cell on the fast path
constexpr int kNb1Log = 10;             // level-1 value bins
constexpr int kNb1 = 1 << kNb1Log;
constexpr int kMaxSub = 40;             // sub-bucket occupancy beyond which the cell goes to the general path

// Shared-memory carve-up of aucell_fused_kernel.  Everything up to the accumulators sits at COMPILE-TIME offsets (they
// depend on the key / position widths only), so the kernel addresses those arrays with immediates instead of
// re-deriving base pointers from kernel arguments all over the cell loop (measured: ~8 % of its instructions).
// The accumulators (2 * n_acc words) are followed by the zero-fill bitmap and its prefixes at run-time offsets.
struct FusedOffsets { int h1, h2, fk, fp, mixed, acc; };
__host__ __device__ constexpr int fused_align16(int v) { return (v + 15) & ~15; }
__host__ __device__ constexpr FusedOffsets fused_offsets(int key_bytes, int pos_bytes) {
    FusedOffsets o{};
    o.h1 = 0;
    o.h2 = fused_align16(o.h1 + (kNb1 + 1) * 4);
    const int h2_bytes = (kCapF + 1) * 4 > kNb1 * key_bytes ? (kCapF + 1) * 4 : kNb1 * key_bytes;   // also one key per bin
    o.fk = fused_align16(o.h2 + h2_bytes);
    o.fp = fused_align16(o.fk + kCapF * (key_bytes == 4 ? 8 : key_bytes));   // packed (key, pos) for 32-bit keys
    o.mixed = fused_align16(o.fp + kCapF * pos_bytes);
    o.acc = fused_align16(o.mixed + kNb1 / 8);
    return o;
}

// Gene-major regulon table with packed entries, indexed by shuffled position.
struct GeneTableDev {
    int n_regulons;
    int n_acc;
    const uint32_t* tptr;               // [n_genes + 1]
    const unsigned long long* tent;     // weighted: [nnz] (W << 16) | accumulator   (W signed 48-bit fixed point)
    const uint16_t* tacc16;             // unweighted: [nnz] accumulator
    const int32_t* acc_first;           // [R] first accumulator of regulon r, -1 = invalid regulon
    const uint8_t* acc_parts;           // [R] 1 or 2
    const double* acc_scale;            // [n_acc]
    const double* max_auc;              // [R]
};

template <typename ValT>
struct FusedArgs {
    const ValT* dense;                  // IN_DENSE
    int64_t row_stride;
    const int64_t* indptr;              // IN_CSR
    const int32_t* indices;             // IN_CSR16: really const uint16_t*
    const ValT* data;
    int64_t n_cells;
    int n_genes;
    int n_words;
    const void* inv_perm;               // [n_genes] shuffled position of a column, as PosT (uint16 when it fits:
                                        // half the footprint of the gather table in L1)
    int rank_cutoff;
    uint32_t inv_ng;                    // ceil(2^32 / n_genes)
    GeneTableDev tab;
    double* out_auc;
    uint32_t* out_rank;                 // RANKS mode (create_rankings): [n_cells x n_genes], shuffled column order
    int32_t* out_nnz;
    // dynamic shared memory carve-up
    int off_h1, off_h2, off_fk, off_fp, off_mixed, off_bits, off_wpre, off_acc;
    int gen_cap;                        // power of two: (key, pos) pairs the general path can sort in shared memory
                                        // (aliases the h1..fp regions)
    // global scratch for the general path: per CTA scratch_cap (key, pos) pairs
    void* scratch_key;
    void* scratch_pos;
    int scratch_cap;
    // to-do list (nullable): rank only the cells todo[0 .. *todo_n) -- what aucell_tie_kernel left over
    const int32_t* todo;
    const int32_t* todo_n;
    // path statistics (nullable): [0] cells, [1] general: too many entries, [2] general: negatives below the
    // cutoff, [3] general: crowded sub-bucket, [4] cells that needed zeros below the cutoff
    unsigned long long* stats;
};

// warp-wide reductions: one REDUX instruction for 32-bit operands, shuffles for 64-bit keys
__device__ __forceinline__ int warp_sum(int v) { return __reduce_add_sync(0xffffffffu, v); }
__device__ __forceinline__ uint32_t warp_min(uint32_t v) { return __reduce_min_sync(0xffffffffu, v); }
__device__ __forceinline__ uint32_t warp_max(uint32_t v) { return __reduce_max_sync(0xffffffffu, v); }
__device__ __forceinline__ uint64_t warp_min(uint64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = tmin(v, (uint64_t)__shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ uint64_t warp_max(uint64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = tmax(v, (uint64_t)__shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// TRAILING_BARRIER = false: the caller guarantees a block barrier before s_warp is written again.
template <int NT, bool TRAILING_BARRIER = true>
__device__ __forceinline__ int block_excl_scan_n(int v, int* s_warp, int& total) {
    constexpr int NW = NT / 32;
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
        int x = (lane < NW) ? s_warp[lane] : 0;
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
    const int res = s_warp[w] + incl - v;
    if (TRAILING_BARRIER) __syncthreads();
    return res;
}

// W * mval into accumulator(s) of every regulon holding the gene at shuffled position p.
template <bool ACC64>
__device__ __forceinline__ void scatter_packed(const GeneTableDev& t, uint32_t* __restrict__ acc_lo,
                                               uint32_t* __restrict__ acc_hi, unsigned p, unsigned mval) {
    const uint32_t b = __ldg(t.tptr + p), e = __ldg(t.tptr + p + 1);
    // (these row loops stay unrolled: the compiler batches the table loads of several entries, which is what
    //  aucell_from_rankings_kernel lives on -- rolled, it runs at half the rate)
    if (ACC64) {
        if (t.tent != nullptr) {
            for (uint32_t j = b; j < e; ++j) {
                const unsigned long long ent = __ldg(t.tent + j);
                const uint32_t r = (uint32_t)(ent & 0xFFFFull);
                const long long w = ((long long)ent) >> 16;
                const unsigned long long c = (unsigned long long)(w * (long long)mval);
                const uint32_t lo = (uint32_t)c;
                uint32_t hi = (uint32_t)(c >> 32);
                const uint32_t old = atomicAdd(&acc_lo[r], lo);
                hi += (old + lo < old) ? 1u : 0u;
                if (hi) atomicAdd(&acc_hi[r], hi);
            }
        } else {  // unit weights but a cutoff too large for 32-bit sums
            for (uint32_t j = b; j < e; ++j) {
                const uint32_t r = __ldg(t.tacc16 + j);
                const uint32_t old = atomicAdd(&acc_lo[r], mval);
                if (old + mval < old) atomicAdd(&acc_hi[r], 1u);
            }
        }
    } else {
        for (uint32_t j = b; j < e; ++j) atomicAdd(&acc_lo[__ldg(t.tacc16 + j)], mval);
    }
}

// One table entry j applied with multiplier mval (the body of scatter_packed, for the queued path).
template <bool ACC64>
__device__ __forceinline__ void apply_entry(const GeneTableDev& t, uint32_t* __restrict__ acc_lo,
                                            uint32_t* __restrict__ acc_hi, uint32_t j, unsigned mval) {
    if (ACC64) {
        if (t.tent != nullptr) {
            const unsigned long long ent = __ldg(t.tent + j);
            const uint32_t r = (uint32_t)(ent & 0xFFFFull);
            const long long w = ((long long)ent) >> 16;
            const unsigned long long c = (unsigned long long)(w * (long long)mval);
            const uint32_t lo = (uint32_t)c;
            uint32_t hi = (uint32_t)(c >> 32);
            const uint32_t old = atomicAdd(&acc_lo[r], lo);
            hi += (old + lo < old) ? 1u : 0u;
            if (hi) atomicAdd(&acc_hi[r], hi);
        } else {
            const uint32_t r = __ldg(t.tacc16 + j);
            const uint32_t old = atomicAdd(&acc_lo[r], mval);
            if (old + mval < old) atomicAdd(&acc_hi[r], 1u);
        }
    } else {
        atomicAdd(&acc_lo[__ldg(t.tacc16 + j)], mval);
    }
}

// Bucket-ordered storage of (key, pos): one 64-bit word when the key is 32-bit (single load + one compare in the
// rank loop), two arrays otherwise.
template <typename KeyT, typename PosT>
struct BucketStore {
    static constexpr bool packed = sizeof(KeyT) == 4;
    static constexpr int key_bytes = packed ? 8 : (int)sizeof(KeyT);
    unsigned long long* kp;   // packed
    KeyT* k;                  // unpacked
    PosT* p;
    __device__ __forceinline__ void put(uint32_t at, KeyT key, PosT pos) const {
        if (packed) kp[at] = ((unsigned long long)key << 32) | (unsigned long long)(uint32_t)pos;
        else { k[at] = key; p[at] = pos; }
    }
    // 1 when the stored pair at `at` sorts strictly before (key, pos)
    __device__ __forceinline__ uint32_t before(uint32_t at, KeyT key, PosT pos, unsigned long long mine) const {
        if (packed) return kp[at] < mine ? 1u : 0u;
        const KeyT ko = k[at];
        return (ko < key || (ko == key && p[at] < pos)) ? 1u : 0u;
    }
};

constexpr int kQ = kCapF / kFWarps;     // per-warp scatter queue (items): 256

// Drain a warp's queue of (table entry, multiplier) items with all 32 lanes busy.  The table entries of four items
// are loaded before any is applied, so four L2 round trips overlap instead of one per iteration.
template <bool ACC64, typename QmT>
__device__ __forceinline__ void flush_queue(const GeneTableDev& t, uint32_t* __restrict__ acc_lo,
                                            uint32_t* __restrict__ acc_hi, const uint32_t* __restrict__ qj,
                                            const QmT* __restrict__ qm, int qn) {
    __syncwarp();
    const int lane = threadIdx.x & 31;
    constexpr int U = 4;
#pragma unroll 1
    for (int base = 0; base < qn; base += 32 * U) {
        unsigned long long ent[U];
        unsigned mval[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = base + u * 32 + lane;
            ent[u] = 0ull;
            mval[u] = 0u;
            if (i < qn) {
                const uint32_t j = qj[i];
                mval[u] = (unsigned)qm[i];
                if (ACC64 && t.tent != nullptr) ent[u] = __ldg(t.tent + j);
                else ent[u] = (unsigned long long)__ldg(t.tacc16 + j);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (mval[u] == 0u) continue;          // multipliers are >= 1 for real items
            if (ACC64) {
                uint32_t r;
                unsigned long long c;
                if (t.tent != nullptr) {
                    r = (uint32_t)(ent[u] & 0xFFFFull);
                    c = (unsigned long long)((((long long)ent[u]) >> 16) * (long long)mval[u]);
                } else {
                    r = (uint32_t)ent[u];
                    c = (unsigned long long)mval[u];
                }
                const uint32_t lo = (uint32_t)c;
                uint32_t hi = (uint32_t)(c >> 32);
                const uint32_t old = atomicAdd(&acc_lo[r], lo);
                hi += (old + lo < old) ? 1u : 0u;
                atomicAdd(&acc_hi[r], hi);        // unconditional: hi is rarely 0, and a branch costs more than the add
            } else {
                atomicAdd(&acc_lo[(uint32_t)ent[u]], mval[u]);
            }
        }
    }
    __syncwarp();
}

// Apply the table rows (tb[k], tc[k]) with multipliers mv[k], LO <= k < HI, through the warp's queue: every lane lists
// its rows' entries at the offsets a warp scan assigns, then all 32 lanes apply them.  When the slots hold more entries
// than the queue (many regulons, deep cutoffs) the slot range is halved recursively.
template <bool ACC64, typename PosT, int NI, int QCAP, int LO, int HI>
struct ScatterRange {
    static __device__ __forceinline__ void run(const GeneTableDev& t, uint32_t* __restrict__ acc_lo,
                                               uint32_t* __restrict__ acc_hi, uint32_t* __restrict__ qj,
                                               PosT* __restrict__ qm, const uint32_t (&tb)[NI], const uint32_t (&tc)[NI],
                                               const uint32_t (&mv)[NI], int lane) {
        int mine_n = 0;
#pragma unroll
        for (int k = LO; k < HI; ++k) mine_n += (int)tc[k];
        int incl = mine_n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int tt = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += tt;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);     // warp-uniform
        if (total == 0) return;
        if (total <= QCAP) {
            int at = incl - mine_n;
#pragma unroll
            for (int k = LO; k < HI; ++k) {
#pragma unroll 1
                for (uint32_t i = 0; i < tc[k]; ++i) {          // a gene sits in ~2 regulons: keep the loop rolled
                    qj[at] = tb[k] + i;
                    qm[at] = (PosT)mv[k];
                    ++at;
                }
            }
            flush_queue<ACC64, PosT>(t, acc_lo, acc_hi, qj, qm, total);
            return;
        }
        if constexpr (HI - LO > 1) {
            ScatterRange<ACC64, PosT, NI, QCAP, LO, (LO + HI) / 2>::run(t, acc_lo, acc_hi, qj, qm, tb, tc, mv, lane);
            ScatterRange<ACC64, PosT, NI, QCAP, (LO + HI) / 2, HI>::run(t, acc_lo, acc_hi, qj, qm, tb, tc, mv, lane);
        } else {                          // one slot alone overflows the queue (genes in hundreds of regulons)
#pragma unroll 1
            for (uint32_t i = 0; i < tc[LO]; ++i) apply_entry<ACC64>(t, acc_lo, acc_hi, tb[LO] + i, mv[LO]);
        }
    }
};

template <bool ACC64, int NT>
__device__ __forceinline__ void write_aucs_packed(const GeneTableDev& t, uint32_t* __restrict__ acc_lo,
                                                  uint32_t* __restrict__ acc_hi, double* __restrict__ out) {
    for (int r = threadIdx.x; r < t.n_regulons; r += NT) {
        const int f = __ldg(t.acc_first + r);
        double res = 0.0;
        if (f >= 0) {
            const int parts = __ldg(t.acc_parts + r);
            double sum = 0.0;
            for (int k = 0; k < parts; ++k) {
                if (ACC64) {
                    const long long v = (long long)(((unsigned long long)acc_hi[f + k] << 32) | acc_lo[f + k]);
                    sum += (double)v * __ldg(t.acc_scale + f + k);
                    acc_hi[f + k] = 0u;
                } else {
                    sum += (double)acc_lo[f + k];
                }
                acc_lo[f + k] = 0u;
            }
            // (a zero numerator sends DDIV down its out-of-line special-case path for the whole warp)
            const double q = __ddiv_rn(nonzero_numerator(sum), __ldg(t.max_auc + r));
            res = sum == 0.0 ? 0.0 : q;
        }
        out[r] = res;
    }
}

// ---- general path: all explicit entries of the cell into global scratch, bitonic sort, scatter --------------
template <typename ValT, typename PosT, int IN, bool ACC64, int NT = kF, bool RANKS = false>
__device__ void general_cell(const FusedArgs<ValT>& a, int64_t cell, uint32_t* s_bits, uint32_t* s_wpre,
                             uint32_t* acc_lo, uint32_t* acc_hi, int* s_cnt, int* s_scan, unsigned char* smem) {
    using KT = KeyOf<ValT>;
    using KeyT = typename KT::type;
    const int tid = threadIdx.x, lane = tid & 31;
    // upper bound of the explicit entries: the row length.  Short rows are sorted in shared memory (the storage of
    // the fast path's histograms and bucket arrays), long ones in global scratch.
    int64_t bound = a.n_genes;
    if (IN != IN_DENSE) bound = tmin<int64_t>(tmax<int64_t>(a.indptr[cell + 1] - a.indptr[cell], 0), (int64_t)a.n_genes);
    KeyT* lk;
    PosT* lp;
    if (bound <= a.gen_cap) {
        lk = reinterpret_cast<KeyT*>(smem + a.off_h1);
        lp = reinterpret_cast<PosT*>(lk + a.gen_cap);
    } else {
        lk = reinterpret_cast<KeyT*>(a.scratch_key) + (size_t)blockIdx.x * a.scratch_cap;
        lp = reinterpret_cast<PosT*>(a.scratch_pos) + (size_t)blockIdx.x * a.scratch_cap;
    }
    if (tid < 2) s_cnt[tid] = 0;
    __syncthreads();
    int64_t count;
    const ValT* vals;
    using IdxT = typename IdxOf<IN>::type;
    const IdxT* cols = nullptr;
    if (IN == IN_DENSE) {
        count = a.n_genes;
        vals = a.dense + cell * a.row_stride;
    } else {
        const int64_t begin = a.indptr[cell];
        count = bound;                         // (clamped: a malformed row must not overrun the scratch)
        vals = a.data + begin;
        cols = reinterpret_cast<const IdxT*>(a.indices) + begin;
    }
    for (int64_t base = 0; base < count; base += NT) {
        const int64_t i = base + tid;
        const bool in = i < count;
        KeyT key = KT::KEY_ZERO;
        if (in) key = KT::make(vals[i]);
        if (IN != IN_DENSE) {                  // a column index out of range: memory-safe, the entry is dropped
            if (in && (unsigned)cols[i] >= (unsigned)a.n_genes) key = KT::KEY_ZERO;
        }
        const bool pred = in && key != KT::KEY_ZERO;
        const unsigned mk = __ballot_sync(0xffffffffu, pred);
        const unsigned mgt = __ballot_sync(0xffffffffu, pred && key < KT::KEY_ZERO);
        if (mk) {
            int wbase = 0;
            const int leader = __ffs(mk) - 1;
            if (lane == leader) {
                wbase = atomicAdd(&s_cnt[0], __popc(mk));
                if (mgt) atomicAdd(&s_cnt[1], __popc(mgt));
            }
            wbase = __shfl_sync(0xffffffffu, wbase, leader);
            if (pred) {
                const int idx = wbase + __popc(mk & lanemask_lt());
                const int g = (IN == IN_DENSE) ? (int)i : (int)cols[i];
                lk[idx] = key;
                lp[idx] = static_cast<const PosT*>(a.inv_perm)[g];
            }
        }
    }
    __syncthreads();
    const int m = s_cnt[0], n_gt = s_cnt[1];
    const int n_genes = a.n_genes, cutoff = RANKS ? a.n_genes : a.rank_cutoff;
    const int n_zero = n_genes - m;
    uint32_t* const out_rank = RANKS ? a.out_rank + cell * (int64_t)n_genes : nullptr;
    int P = 2;
    while (P < m) P <<= 1;
    for (int i = m + tid; i < P; i += NT) {
        lk[i] = KT::KEY_NAN;
        lp[i] = (PosT)~(PosT)0;
    }
    // bitonic sort by (key, position); pairs are unique so the result is order-independent
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = tid; t < (P >> 1); t += NT) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool asc = ((lo & size) == 0);
                const KeyT ka = lk[lo], kb = lk[hi];
                const PosT pa = lp[lo], pb = lp[hi];
                const bool gt = (ka > kb) || (ka == kb && pa > pb);
                if (gt == asc) {
                    lk[lo] = kb; lk[hi] = ka;
                    lp[lo] = pb; lp[hi] = pa;
                }
            }
        }
    }
    __syncthreads();
    if (RANKS) {
        // create_rankings: sorted index = rank for the positives, index + n_zero for negatives / NaN
        for (int i = tid; i < m; i += NT) out_rank[(unsigned)lp[i]] = (uint32_t)(i < n_gt ? i : i + n_zero);
    } else {
        const int top_pos = min(n_gt, cutoff);
        for (int i = tid; i < top_pos; i += NT)
            scatter_packed<ACC64>(a.tab, acc_lo, acc_hi, (unsigned)lp[i], (unsigned)(cutoff - i));
    }
    const int q_zero = cutoff - n_gt;
    if (q_zero > 0) {
        for (int w = tid; w < a.n_words; w += NT) s_bits[w] = 0u;
        __syncthreads();
        for (int i = tid; i < m; i += NT) {
            const unsigned p = (unsigned)lp[i];
            atomicOr(&s_bits[p >> 5], 1u << (p & 31));
        }
        __syncthreads();
        const int per = (a.n_words + NT - 1) / NT;
        const int w0 = tid * per, w1 = min(w0 + per, a.n_words);
        int local = 0;
        for (int w = w0; w < w1; ++w) local += __popc(s_bits[w]);
        int total;
        int run = block_excl_scan_n<NT>(local, s_scan, total);
        for (int w = w0; w < w1; ++w) {
            s_wpre[w] = (uint32_t)run;
            run += __popc(s_bits[w]);
        }
        __syncthreads();
        for (int p = tid; p < n_genes; p += NT) {
            const int w = p >> 5;
            const uint32_t bw = s_bits[w];
            const int z0 = (w << 5) - (int)s_wpre[w];
            if (z0 >= q_zero) break;                    // words are visited in increasing order by every thread
            if (!((bw >> (p & 31)) & 1u)) {
                const int z = z0 + (p & 31) - __popc(bw & ((1u << (p & 31)) - 1u));
                if (RANKS) out_rank[p] = (uint32_t)(n_gt + z);
                else if (z < q_zero) scatter_packed<ACC64>(a.tab, acc_lo, acc_hi, (unsigned)p, (unsigned)(q_zero - z));
            }
        }
        if (!RANKS) {
            const int q_neg = q_zero - n_zero;
            const int neg_end = q_neg > 0 ? min(m, n_gt + q_neg) : n_gt;
            for (int i = n_gt + tid; i < neg_end; i += NT)
                scatter_packed<ACC64>(a.tab, acc_lo, acc_hi, (unsigned)lp[i], (unsigned)(cutoff - (i + n_zero)));
        }
    }
    if (a.out_nnz != nullptr && tid == 0) a.out_nnz[cell] = m;
}

// Run f(SlotIdx<k>) for k = nk-1 ... 0 with ONE multi-way branch (Duff's device) instead of a `k < nk` test per slot:
// item slots >= nk hold nothing, and nk is uniform over the CTA.
template <int K>
struct SlotIdx { static constexpr int value = K; };

template <int NI, typename F>
__device__ __forceinline__ void for_slots(int nk, F&& f) {
    static_assert(NI == 8, "for_slots is written out for 8 item slots");
    switch (nk) {
        default: f(SlotIdx<7>{}); [[fallthrough]];
        case 7: f(SlotIdx<6>{}); [[fallthrough]];
        case 6: f(SlotIdx<5>{}); [[fallthrough]];
        case 5: f(SlotIdx<4>{}); [[fallthrough]];
        case 4: f(SlotIdx<3>{}); [[fallthrough]];
        case 3: f(SlotIdx<2>{}); [[fallthrough]];
        case 2: f(SlotIdx<1>{}); [[fallthrough]];
        case 1: f(SlotIdx<0>{}); [[fallthrough]];
        case 0: break;
    }
}

// ---- row access ---------------------------------------------------------------------------------------------
template <typename ValT, typename IdxT = int32_t>
struct RowView {
    const ValT* vals;
    const IdxT* cols;        // nullptr for dense rows (column == index)
    int count;               // stored entries (CSR) or n_genes (dense)
    int n_genes;             // CSR: a stored column index outside [0, n_genes) is treated like an explicit zero
};

// Visit every element of a row with all threads in lock step (uniform trip count, so `f` may use warp votes):
// f(is_explicit, key, column).  Dense rows are read with 128-bit loads over their 16-byte aligned interior.
template <int NT = kF, typename ValT, typename IdxT, typename F>
__device__ __forceinline__ void stream_row(const RowView<ValT, IdxT>& row, F f) {
    using KT = KeyOf<ValT>;
    using KeyT = typename KT::type;
    const int tid = (int)threadIdx.x;
    if (row.cols != nullptr) {
        for (int base = 0; base < row.count; base += NT) {
            const int i = base + tid;
            KeyT key = KT::KEY_ZERO;
            int col = 0;
            if (i < row.count) {
                key = KT::make(row.vals[i]);
                col = (int)row.cols[i];
                if ((unsigned)col >= (unsigned)row.n_genes) {      // malformed input: memory-safe, the entry is dropped
                    key = KT::KEY_ZERO;
                    col = 0;
                }
            }
            f(key != KT::KEY_ZERO, key, col);
        }
        return;
    }
    constexpr int V = 16 / (int)sizeof(ValT);                 // elements per 128-bit load
    struct alignas(16) Vec { ValT v[V]; };
    const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(row.vals) & 15u);
    const int head = min(row.count, (int)(((16u - mis) & 15u) / sizeof(ValT)));
    {   // ragged head (fewer than V elements): one pass
        KeyT key = KT::KEY_ZERO;
        if (tid < head) key = KT::make(row.vals[tid]);
        if (head > 0) f(key != KT::KEY_ZERO, key, tid);
    }
    const int nvec = (row.count - head) / V;
    const Vec* vp = reinterpret_cast<const Vec*>(row.vals + head);
    for (int base = 0; base < nvec; base += NT) {
        const int q = base + tid;
        Vec x;
#pragma unroll
        for (int j = 0; j < V; ++j) x.v[j] = (ValT)0;
        if (q < nvec) x = vp[q];
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const KeyT key = KT::make(x.v[j]);
            f(key != KT::KEY_ZERO, key, head + q * V + j);
        }
    }
    {   // ragged tail
        const int t0 = head + nvec * V;
        if (t0 < row.count) {
            KeyT key = KT::KEY_ZERO;
            const int i = t0 + tid;
            if (i < row.count) key = KT::make(row.vals[i]);
            f(key != KT::KEY_ZERO, key, i);
        }
    }
}

// Dense rows, one 128-bit vector per call: f(keys[V], first column).  Elements that do not exist (ragged head / tail,
// threads past the end) are KEY_ZERO.  Uniform trip count, so `f` may use warp votes.
template <int NT = kF, typename ValT, typename IdxT, typename F>
__device__ __forceinline__ void stream_dense_vec(const RowView<ValT, IdxT>& row, F f) {
    using KT = KeyOf<ValT>;
    using KeyT = typename KT::type;
    const int tid = (int)threadIdx.x;
    constexpr int V = 16 / (int)sizeof(ValT);
    struct alignas(16) Vec { ValT v[V]; };
    const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(row.vals) & 15u);
    const int head = min(row.count, (int)(((16u - mis) & 15u) / sizeof(ValT)));
    KeyT key[V];
    if (head > 0) {   // ragged head (fewer than V elements): one element per thread
#pragma unroll
        for (int j = 0; j < V; ++j) key[j] = KT::KEY_ZERO;
        if (tid < head) key[0] = KT::make(row.vals[tid]);
        f(key, tid);
    }
    const int nvec = (row.count - head) / V;
    const Vec* vp = reinterpret_cast<const Vec*>(row.vals + head);
    for (int base = 0; base < nvec; base += NT) {
        const int q = base + tid;
        Vec x;
#pragma unroll
        for (int j = 0; j < V; ++j) x.v[j] = (ValT)0;
        if (q < nvec) x = vp[q];
#pragma unroll
        for (int j = 0; j < V; ++j) key[j] = KT::make(x.v[j]);
        f(key, head + q * V);
    }
    const int t0 = head + nvec * V;
    if (t0 < row.count) {   // ragged tail
#pragma unroll
        for (int j = 0; j < V; ++j) key[j] = KT::KEY_ZERO;
        const int i = t0 + tid;
        if (i < row.count) key[0] = KT::make(row.vals[i]);
        f(key, i);
    }
}

// ---- the fused kernel ---------------------------------------------------------------------------------------
// RANKS = true turns the kernel into create_rankings: the cutoff is n_genes (every gene is ranked) and the ranks are
// written to out_rank instead of being scattered into regulon accumulators.
template <typename ValT, typename PosT, int IN, bool ACC64, bool RANKS = false>
__global__ void __launch_bounds__(kF, 2) aucell_fused_kernel(const FusedArgs<ValT> a) {
    using KT = KeyOf<ValT>;
    using KeyT = typename KT::type;
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr FusedOffsets L = fused_offsets((int)sizeof(KeyT), (int)sizeof(PosT));
    uint32_t* const h1 = reinterpret_cast<uint32_t*>(smem + L.h1);          // [kNb1 + 1] level-1 counts / bases
    uint32_t* const h2 = reinterpret_cast<uint32_t*>(smem + L.h2);          // [kCapF + 1] level-2 counts / bases
    KeyT* const rep = reinterpret_cast<KeyT*>(smem + L.h2);                 // [kNb1] alias of h2: one key per bin
    KeyT* const fk = reinterpret_cast<KeyT*>(smem + L.fk);                  // staging keys / packed bucket store
    PosT* const fp = reinterpret_cast<PosT*>(smem + L.fp);                  // staging positions / bucket store / queue
    uint32_t* const mixed = reinterpret_cast<uint32_t*>(smem + L.mixed);    // [kNb1 / 32] bins with > 1 key
    uint32_t* const acc_lo = reinterpret_cast<uint32_t*>(smem + L.acc);
    uint32_t* const s_bits = reinterpret_cast<uint32_t*>(smem + a.off_bits);
    uint32_t* const s_wpre = reinterpret_cast<uint32_t*>(smem + a.off_wpre);
    uint32_t* const acc_hi = acc_lo + a.tab.n_acc;
    BucketStore<KeyT, PosT> store;
    store.kp = reinterpret_cast<unsigned long long*>(smem + L.fk);
    store.k = fk;
    store.p = fp;
    __shared__ int s_scan[33];
    __shared__ int s_scan2[33];
    __shared__ int s_cnt[2];
    __shared__ int s_red_i[2][kFWarps];
    __shared__ KeyT s_red_k[2][kFWarps];
    __shared__ int s_flag[2];
    __shared__ int s_str[2];              // the value bin that straddles the cutoff: index, base rank

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_genes = a.n_genes, cutoff = RANKS ? a.n_genes : a.rank_cutoff;

    for (int i = tid; i < 2 * a.tab.n_acc; i += kF) acc_lo[i] = 0u;
    // epilogue constants of "my" regulon (thread r owns regulon r when there are at most kF regulons)
    const bool own_reg = !RANKS && a.tab.n_regulons <= kF;
    int my_first = -1, my_parts = 0;
    double my_scale0 = 1.0, my_scale1 = 0.0, my_max_auc = 1.0;
    if (own_reg && tid < a.tab.n_regulons) {
        my_first = __ldg(a.tab.acc_first + tid);
        if (my_first >= 0) {
            my_parts = __ldg(a.tab.acc_parts + tid);
            my_scale0 = __ldg(a.tab.acc_scale + my_first);
            if (my_parts > 1) my_scale1 = __ldg(a.tab.acc_scale + my_first + 1);
            my_max_auc = __ldg(a.tab.max_auc + tid);
        }
    }
    __syncthreads();

    const int64_t n_iter = a.todo != nullptr ? (int64_t)*a.todo_n : a.n_cells;
    for (int64_t it = blockIdx.x; it < n_iter; it += gridDim.x) {
        const int64_t cell = a.todo != nullptr ? (int64_t)a.todo[it] : it;
        using IdxT = typename IdxOf<IN>::type;
        RowView<ValT, IdxT> row;
        if (IN != IN_DENSE) {
            const int64_t begin = a.indptr[cell];
            row.vals = a.data + begin;
            row.cols = reinterpret_cast<const IdxT*>(a.indices) + begin;
            // (a canonical row holds at most n_genes entries; a longer one -- duplicates -- must not overrun the scratch
            // of the general path, a negative length must not be walked at all)
            row.count = (int)tmin<int64_t>(tmax<int64_t>(a.indptr[cell + 1] - begin, 0), (int64_t)n_genes);
        } else {
            row.vals = a.dense + cell * a.row_stride;
            row.cols = nullptr;
            row.count = n_genes;
        }
        row.n_genes = n_genes;

        // ---------------- LOAD ---------------------------------------------------------------------------------
        // direct mode: every explicit entry of the cell sits in registers (kItems per thread).
        // select mode (long or dense rows): the row is streamed from global memory / L2 for the statistics and the
        // level-1 histogram, and only the entries of the bins that start below the cutoff are brought to registers.
        KeyT key[kItems];
        PosT pos[kItems];
        int nk = 0;                       // item slots in use (block-uniform): slots >= nk hold nothing
        bool direct;
        bool general = false;
        int why = 1;                      // reason code for the general path (statistics only)
        int c_exp = 0, c_pos = 0;         // this thread's explicit / positive entries
        KeyT kmin = KT::KEY_ZERO, kmax = 0;
        if (IN != IN_DENSE) {
            direct = row.count <= kCapF;
            if (direct) {
                nk = (row.count + kF - 1) / kF;
                ValT v[kItems];
                int c[kItems];
#pragma unroll
                for (int k = 0; k < kItems; ++k) { v[k] = (ValT)0; c[k] = 0; key[k] = KT::KEY_ZERO; pos[k] = (PosT)0; }
                for_slots<kItems>(nk, [&](auto K) {
                    constexpr int k = decltype(K)::value;
                    const int i = k * kF + tid;
                    if (i < row.count) {
                        v[k] = row.vals[i];
                        c[k] = (int)row.cols[i];
                        if ((unsigned)c[k] >= (unsigned)n_genes) { v[k] = (ValT)0; c[k] = 0; }      // malformed input: dropped
                    }
                });
                for_slots<kItems>(nk, [&](auto K) {
                    constexpr int k = decltype(K)::value;
                    key[k] = KT::make(v[k]);
                    pos[k] = __ldg(static_cast<const PosT*>(a.inv_perm) + c[k]);      // unconditional: padding slots read inv_perm[0]
                });
            }
        } else {
            // dense row: compact the non-zero values into the staging arrays while taking the statistics
            if (tid == 0) s_cnt[0] = 0;
            __syncthreads();
            constexpr int V = 16 / (int)sizeof(ValT);
            stream_dense_vec(row, [&](const KeyT (&kk)[V], int col0) {
                // a thread's non-zeros (0..V) get consecutive staging slots: one offset computation per vector
                // (three votes) instead of one vote + atomic per element
                // branch-free statistics: zeros carry KEY_ZERO and negatives larger keys, so neither lowers kmin;
                // kmax only looks at positives
                int c = 0;
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    const bool ps = kk[j] < KT::KEY_ZERO;
                    c += (kk[j] != KT::KEY_ZERO) ? 1 : 0;
                    c_pos += ps ? 1 : 0;
                    kmin = tmin(kmin, kk[j]);
                    kmax = tmax(kmax, ps ? kk[j] : (KeyT)0);
                }
                c_exp += c;
                const unsigned b0 = __ballot_sync(0xffffffffu, c & 1);
                const unsigned b1 = __ballot_sync(0xffffffffu, c & 2);
                const unsigned b2 = __ballot_sync(0xffffffffu, c & 4);
                if (b0 | b1 | b2) {
                    const unsigned lt = lanemask_lt();
                    const int wtot = __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2);
                    int wbase = 0;
                    if (lane == 0) wbase = atomicAdd(&s_cnt[0], wtot);
                    wbase = __shfl_sync(0xffffffffu, wbase, 0);
                    int idx = wbase + __popc(b0 & lt) + 2 * __popc(b1 & lt) + 4 * __popc(b2 & lt);
                    const PosT* const ip = static_cast<const PosT*>(a.inv_perm) + col0;
#pragma unroll
                    for (int j = 0; j < V; ++j) {
                        const bool nz = kk[j] != KT::KEY_ZERO;
                        if (nz && idx < kCapF) {          // short body: predicated, no divergence bookkeeping
                            fk[idx] = kk[j];
                            fp[idx] = __ldg(ip + j);
                        }
                        idx += nz ? 1 : 0;
                    }
                }
            });
            __syncthreads();
            const int m_total = s_cnt[0];
            direct = m_total <= kCapF;
            if (direct) {
                nk = (m_total + kF - 1) / kF;
#pragma unroll
                for (int k = 0; k < kItems; ++k) {
                    const int i = k * kF + tid;
                    key[k] = (k < nk && i < m_total) ? fk[i] : KT::KEY_ZERO;
                    pos[k] = (k < nk && i < m_total) ? fp[i] : (PosT)0;
                }
            }
        }
        // ---------------- cell statistics: #explicit, #positive, key range of the positives --------------------
        if (direct) {
            if (IN != IN_DENSE) {
                for_slots<kItems>(nk, [&](auto K) {       // branch-free: zeros / negatives never lower kmin
                    constexpr int k = decltype(K)::value;
                    const bool ps = key[k] < KT::KEY_ZERO;
                    c_exp += (key[k] != KT::KEY_ZERO) ? 1 : 0;
                    c_pos += ps ? 1 : 0;
                    kmin = tmin(kmin, key[k]);
                    kmax = tmax(kmax, ps ? key[k] : (KeyT)0);
                });
            }
        } else if (IN != IN_DENSE) {
            stream_row(row, [&](bool ex, KeyT kk, int) {
                if (ex) {
                    ++c_exp;
                    if (kk < KT::KEY_ZERO) { ++c_pos; kmin = tmin(kmin, kk); kmax = tmax(kmax, kk); }
                }
            });
        }
        c_exp = warp_sum(c_exp);
        c_pos = warp_sum(c_pos);
        kmin = warp_min(kmin);
        kmax = warp_max(kmax);
        if (lane == 0) {
            s_red_i[0][warp] = c_exp; s_red_i[1][warp] = c_pos;
            s_red_k[0][warp] = kmin; s_red_k[1][warp] = kmax;
        }
#pragma unroll 1
        for (int b = tid; b <= kNb1; b += kF) h1[b] = 0u;     // clear the level-1 histogram while we are at it
        if (tid < (kNb1 >> 5)) mixed[tid] = 0u;
        if (tid == 0) { s_flag[0] = 0; s_flag[1] = 0x7FFFFFFF; s_cnt[1] = 0; }
        __syncthreads();
        int m, n_gt;
        {
            const bool has = lane < kFWarps;
            m = warp_sum(has ? s_red_i[0][lane] : 0);
            n_gt = warp_sum(has ? s_red_i[1][lane] : 0);
            kmin = warp_min(has ? s_red_k[0][lane] : (KeyT)KT::KEY_ZERO);
            kmax = warp_max(has ? s_red_k[1][lane] : (KeyT)0);
        }
        const int n_zero = n_genes - m;
        if (cutoff - n_gt - n_zero > 0) { general = true; why = 2; }   // negatives rank below the cutoff

        // ---------------- ZEROS: implicit zeros fill ranks n_gt .. cutoff-1 in position order ---------------------
        // (bitmap of the positions of explicit entries + prefix popcount).  AUC mode runs it after the scatter of the
        // positives; RANKS mode (cutoff = n_genes: every zero gets a rank) runs it BEFORE the ranking, see below.
        auto zeros_phase = [&]() {
            const int q_zero = cutoff - n_gt;
            if (a.stats != nullptr && tid == 0) atomicAdd(a.stats + 4, 1ull);
#pragma unroll 1
            for (int w = tid; w < a.n_words; w += kF) s_bits[w] = 0u;
            __syncthreads();
            if (direct) {
#pragma unroll
                for (int k = 0; k < kItems; ++k)
                    if (k < nk && key[k] != KT::KEY_ZERO) atomicOr(&s_bits[(unsigned)pos[k] >> 5], 1u << ((unsigned)pos[k] & 31));
            } else {
                stream_row(row, [&](bool ex, KeyT, int col) {
                    if (ex) {
                        const unsigned p = (unsigned)__ldg(static_cast<const PosT*>(a.inv_perm) + col);
                        atomicOr(&s_bits[p >> 5], 1u << (p & 31));
                    }
                });
            }
            __syncthreads();
            const int per = (a.n_words + kF - 1) / kF;
            const int w0 = tid * per, w1 = min(w0 + per, a.n_words);
            int local = 0;
#pragma unroll 1
            for (int w = w0; w < w1; ++w) local += __popc(s_bits[w]);
            int total;
            int run = block_excl_scan_n<kF>(local, s_scan, total);
#pragma unroll 1
            for (int w = w0; w < w1; ++w) {
                s_wpre[w] = (uint32_t)run;
                run += __popc(s_bits[w]);
            }
            __syncthreads();
            if (RANKS) {
                // every position gets n_gt + #zeros before it, four positions per 128-bit store; what lands on the
                // positions of explicit entries is overwritten by their own ranks later (S7 / the general path run
                // after this phase in RANKS mode, behind block barriers)
                uint32_t* const orow = a.out_rank + cell * (int64_t)n_genes;
                auto zrank = [&](int p) -> uint32_t {
                    const int w = p >> 5;
                    return (uint32_t)(n_gt + p - (int)s_wpre[w] - __popc(s_bits[w] & ((1u << (p & 31)) - 1u)));
                };
                const int oh = min(n_genes, (int)(((16u - (unsigned)(reinterpret_cast<uintptr_t>(orow) & 15u)) & 15u) >> 2));
                const int nq = (n_genes - oh) >> 2;
                if (tid < oh) orow[tid] = zrank(tid);
                uint4* const o4 = reinterpret_cast<uint4*>(orow + oh);
#pragma unroll 1
                for (int q = tid; q < nq; q += kF) {
                    const int p = oh + (q << 2);
                    o4[q] = make_uint4(zrank(p), zrank(p + 1), zrank(p + 2), zrank(p + 3));
                }
                for (int p = oh + (nq << 2) + tid; p < n_genes; p += kF) orow[p] = zrank(p);
            } else {
#pragma unroll 1
                for (int p = tid; p < n_genes; p += kF) {
                    const int w = p >> 5;
                    const uint32_t bw = s_bits[w];
                    const int z0 = (w << 5) - (int)s_wpre[w];
                    if (z0 >= q_zero) break;
                    if (!((bw >> (p & 31)) & 1u)) {
                        const int z = z0 + (p & 31) - __popc(bw & ((1u << (p & 31)) - 1u));
                        if (z < q_zero) scatter_packed<ACC64>(a.tab, acc_lo, acc_hi, (unsigned)p, (unsigned)(q_zero - z));
                    }
                }
            }
        };
        if (RANKS && !general && n_gt < cutoff) {
            zeros_phase();
            __syncthreads();              // orders the vector stores above before the positives' own rank stores
        }
        if (!general && n_gt > 0) {
            // ---------------- RANK: two-level bucket sort of the positives ------------------------------------
            const KeyT range = kmax - kmin;
            int nbits = 0;
            if (range) nbits = (sizeof(KeyT) == 8) ? 64 - __clzll((long long)range) : 32 - __clz((int)range);
            const int shift = max(0, nbits - kNb1Log);
            // S1: level-1 histogram, one representative key per bin
            if (direct) {
                for_slots<kItems>(nk, [&](auto K) {
                    constexpr int k = decltype(K)::value;
                    if (key[k] < KT::KEY_ZERO) {
                        const uint32_t b = (uint32_t)((key[k] - kmin) >> shift);
                        atomicAdd(&h1[b], 1u);
                        rep[b] = key[k];
                    }
                });
            } else {
                stream_row(row, [&](bool ex, KeyT kk, int) {
                    if (ex && kk < KT::KEY_ZERO) {
                        const uint32_t b = (uint32_t)((kk - kmin) >> shift);
                        atomicAdd(&h1[b], 1u);
                        rep[b] = kk;
                    }
                });
            }
            __syncthreads();
            // S2: bins holding more than one distinct key (direct mode; select mode does it while selecting);
            //     exclusive scan of the bin counts; n_s = entries in the bins that start below the cutoff
            if (direct) {
                for_slots<kItems>(nk, [&](auto K) {
                    constexpr int k = decltype(K)::value;
                    if (key[k] < KT::KEY_ZERO) {
                        const uint32_t b = (uint32_t)((key[k] - kmin) >> shift);
                        if (rep[b] != key[k]) atomicOr(&mixed[b >> 5], 1u << (b & 31));
                    }
                });
            }
            {
                constexpr int per = kNb1 / kF;
                const int b0 = tid * per;
                int cnt[per];
                int local = 0;
#pragma unroll
                for (int j = 0; j < per; ++j) { cnt[j] = (int)h1[b0 + j]; local += cnt[j]; }
                int total;
                int run = block_excl_scan_n<kF, false>(local, s_scan, total);
#pragma unroll
                for (int j = 0; j < per; ++j) {
                    h1[b0 + j] = (uint32_t)run;
                    if (run < cutoff && run + cnt[j] >= cutoff) {      // at most one writer
                        s_flag[1] = run + cnt[j];
                        s_str[0] = b0 + j;
                        s_str[1] = run;
                    }
                    run += cnt[j];
                }
                if (tid == 0) h1[kNb1] = (uint32_t)n_gt;
            }
            __syncthreads();
            const int n_s = min(s_flag[1], n_gt);      // sentinel when all positives rank below the cutoff
            int n_sub = n_s;                           // level-2 sub-buckets of the kept bins
            int n_ent = n_s;                           // entries of the kept bins that come to registers
            if (!direct) {
                // SELECT: bring the entries of the kept bins to registers (through the staging arrays).
                // When even the selection exceeds the capacity, the bin that straddles the cutoff is usually one huge
                // tie class (the lowest expressed value of a deeply sequenced cell).  Of a tie class only the q =
                // cutoff - base members at the smallest shuffled positions can rank below the cutoff, so only its
                // first q + 8 sqrt(q) + 32 position sub-buckets are kept (positions are a random permutation: the
                // count they hold is q + 8 sqrt(q) + 32 +- sqrt of that); the count is verified after the pass.
                bool pruned = false;
                uint32_t b_star = 0xFFFFFFFFu, c_star = 0u, s_star = 0u;
                int q_star = 0;
                if (n_s > kCapF && !RANKS && s_flag[1] != 0x7FFFFFFF) {
                    b_star = (uint32_t)s_str[0];
                    const int base_star = s_str[1];
                    c_star = (uint32_t)(s_flag[1] - base_star);
                    q_star = cutoff - base_star;
                    const int keep_sub = q_star + 8 * (int)sqrtf((float)q_star) + 32;
                    const int slack = 8 * (int)sqrtf((float)keep_sub) + 32;
                    if (keep_sub < (int)c_star && base_star + keep_sub + slack <= kCapF) {
                        pruned = true;
                        s_star = (uint32_t)keep_sub;
                        n_sub = base_star + keep_sub;
                    }
                }
                if (n_sub > kCapF) {
                    general = true;           // even the selection does not fit: general path
                } else {
                    stream_row(row, [&](bool ex, KeyT kk, int col) {
                        bool keep = false;
                        PosT ps = (PosT)0;
                        if (ex && kk < KT::KEY_ZERO) {
                            const uint32_t b = (uint32_t)((kk - kmin) >> shift);
                            keep = (int)h1[b] < cutoff;
                            if (keep) {
                                if (rep[b] != kk) atomicOr(&mixed[b >> 5], 1u << (b & 31));
                                ps = __ldg(static_cast<const PosT*>(a.inv_perm) + col);
                                if (b == b_star)      // pruned tie class: the sub-bucket S4 will compute for it
                                    keep = (uint32_t)(((unsigned long long)((unsigned)ps) * c_star * a.inv_ng) >> 32) < s_star;
                            }
                        }
                        const unsigned mk = __ballot_sync(0xffffffffu, keep);
                        if (mk) {
                            int wbase = 0;
                            const int leader = __ffs(mk) - 1;
                            if (lane == leader) wbase = atomicAdd(&s_cnt[1], __popc(mk));
                            wbase = __shfl_sync(0xffffffffu, wbase, leader);
                            if (keep) {
                                const int idx = wbase + __popc(mk & lanemask_lt());
                                if (idx < kCapF) {
                                    fk[idx] = kk;
                                    fp[idx] = ps;
                                }
                            }
                        }
                    });
                    __syncthreads();
                    if (pruned) {
                        n_ent = s_cnt[1];
                        // the pruning is only valid for a single-key bin that kept at least q members; the staging
                        // arrays must have held everything
                        if (((mixed[b_star >> 5] >> (b_star & 31)) & 1u) || n_ent > kCapF || n_ent - s_str[1] < q_star) {
                            general = true;
                            n_ent = 0;
                        }
                    }
                    nk = (n_ent + kF - 1) / kF;
#pragma unroll
                    for (int k = 0; k < kItems; ++k) {
                        const int i = k * kF + tid;
                        key[k] = (k < nk && i < n_ent) ? fk[i] : KT::KEY_ZERO;
                        pos[k] = (k < nk && i < n_ent) ? fp[i] : (PosT)0;
                    }
                }
                __syncthreads();              // staging and rep are free again
            }
            if (!general) {
                // S3: clear level-2 histogram (rep is dead: its last readers finished before the barriers above)
#pragma unroll 1
                for (int b = tid; b <= n_sub; b += kF) h2[b] = 0u;
                __syncthreads();
                // S4: level-2 histogram
                uint32_t sbs[kItems];     // (sub-bucket << 8) | arrival slot ; 0xFFFFFFFF = dropped
#pragma unroll
                for (int k = 0; k < kItems; ++k) sbs[k] = 0xFFFFFFFFu;
                for_slots<kItems>(nk, [&](auto K) {
                    constexpr int k = decltype(K)::value;
                    if (key[k] < KT::KEY_ZERO) {
                        const uint32_t bin = (uint32_t)((key[k] - kmin) >> shift);
                        const uint32_t base = h1[bin];
                        if ((int)base < cutoff) {
                            const uint32_t c = h1[bin + 1] - base;
                            uint32_t sub;
                            if (!((mixed[bin >> 5] >> (bin & 31)) & 1u)) {
                                // one key (tie class): split by shuffled position, ~1 entry per sub-bucket
                                sub = (uint32_t)(((unsigned long long)((unsigned)pos[k]) * c * a.inv_ng) >> 32);
                            } else {
                                // several keys (continuous data): split by the key's offset inside the bin, 20 bits of
                                // it -- monotone in the key, so sub-bucket order stays rank order
                                const KeyT off = (KeyT)(key[k] - kmin) & (((KeyT)1 << shift) - 1);
                                const uint32_t t20 = shift >= 20 ? (uint32_t)(off >> (shift - 20)) : (uint32_t)(off << (20 - shift));
                                sub = (uint32_t)(((unsigned long long)t20 * c) >> 20);
                            }
                            sub = min(sub, c - 1u);
                            const uint32_t sb = base + sub;
                            const uint32_t got = atomicAdd(&h2[sb], 1u);
                            if (got >= (uint32_t)kMaxSub) s_flag[0] = 1;
                            sbs[k] = (sb << 8) | min(got, 255u);
                        }
                    }
                });
                __syncthreads();
                if (s_flag[0]) {
                    general = true;       // crowded sub-bucket (uniform decision: read after the barrier)
                    why = 3;
                } else {
                    // S5: exclusive scan of level-2 counts
                    {
                        const int per = (n_sub + kF - 1) / kF;
                        const int b0 = tid * per, b1 = min(b0 + per, n_sub);
                        int local = 0;
#pragma unroll 1
                        for (int b = b0; b < b1; ++b) local += (int)h2[b];
                        int total;
                        int run = block_excl_scan_n<kF, false>(local, s_scan2, total);
#pragma unroll 1
                        for (int b = b0; b < b1; ++b) {
                            const int c = (int)h2[b];
                            h2[b] = (uint32_t)run;
                            run += c;
                        }
                        if (tid == 0) h2[n_sub] = (uint32_t)total;      // = entries in registers (n_sub unless pruned)
                    }
                    __syncthreads();
                    // S6: entries into sub-bucket order (sub-buckets starting at or beyond the cutoff are dropped)
                    for_slots<kItems>(nk, [&](auto K) {
                        constexpr int k = decltype(K)::value;
                        if (sbs[k] != 0xFFFFFFFFu) {
                            const uint32_t b2 = h2[sbs[k] >> 8];
                            if ((int)b2 < cutoff) store.put(b2 + (sbs[k] & 255u), key[k], pos[k]);
                            else sbs[k] = 0xFFFFFFFFu;
                        }
                    });
                    __syncthreads();
                    // S7: exact rank inside the sub-bucket -> multiplier cutoff - rank (0 = not below the cutoff),
                    //     and the gene's row of the regulon table
                    uint32_t mv[kItems], tb[kItems], tc[kItems];
#pragma unroll
                    for (int k = 0; k < kItems; ++k) { mv[k] = 0u; tb[k] = 0u; tc[k] = 0u; }
                    for_slots<kItems>(nk, [&](auto K) {
                        constexpr int k = decltype(K)::value;
                        if (sbs[k] != 0xFFFFFFFFu) {
                            const uint32_t sb = sbs[k] >> 8;
                            const uint32_t b = h2[sb], e = h2[sb + 1];
                            const unsigned long long mine = ((unsigned long long)key[k] << 32) | (unsigned long long)(uint32_t)pos[k];
                            uint32_t rank = b;
                            // sub-buckets hold ~1 entry: a plain loop beats the 8/4/2/1 unrolling ladder
#pragma unroll 1
                            for (uint32_t j = b; j < e; ++j) rank += store.before(j, key[k], pos[k], mine);
                            if (RANKS) {
                                a.out_rank[cell * (int64_t)n_genes + (unsigned)pos[k]] = rank;
                            } else if ((int)rank < cutoff) {
                                mv[k] = (uint32_t)(cutoff - (int)rank);
                                tb[k] = __ldg(a.tab.tptr + (unsigned)pos[k]);
                                tc[k] = __ldg(a.tab.tptr + (unsigned)pos[k] + 1) - tb[k];
                            }
                        }
                    });
                    __syncthreads();      // h2 / fp are dead: their storage becomes the per-warp scatter queues
                    // S8: scatter through the per-warp queue (multipliers are as wide as positions: a cutoff above
                    //     65535 implies 32-bit positions)
                    if (!RANKS)
                        ScatterRange<ACC64, PosT, kItems, kQ, 0, kItems>::run(a.tab, acc_lo, acc_hi, h2 + warp * kQ,
                                                                              fp + warp * kQ, tb, tc, mv, lane);
                }
            }
        }
        if (!RANKS && !general && n_gt < cutoff) zeros_phase();
        if (!general && a.out_nnz != nullptr && tid == 0) a.out_nnz[cell] = m;
        if (a.stats != nullptr && tid == 0) {
            atomicAdd(a.stats + 0, 1ull);
            if (general) atomicAdd(a.stats + why, 1ull);
        }
        if (general) {
            __syncthreads();
            general_cell<ValT, PosT, IN, ACC64, kF, RANKS>(a, cell, s_bits, s_wpre, acc_lo, acc_hi, s_cnt, s_scan, smem);
        }
        __syncthreads();
        if (RANKS) continue;              // nothing to reduce: the ranks are already in global memory
        if (own_reg) {
            if (tid < a.tab.n_regulons) {
                double res = 0.0;
                if (my_first >= 0) {
                    double sum;
                    if (ACC64) {
                        const long long v0 = (long long)(((unsigned long long)acc_hi[my_first] << 32) | acc_lo[my_first]);
                        sum = (double)v0 * my_scale0;
                        acc_hi[my_first] = 0u;
                        if (my_parts > 1) {
                            const long long v1 = (long long)(((unsigned long long)acc_hi[my_first + 1] << 32) | acc_lo[my_first + 1]);
                            sum += (double)v1 * my_scale1;
                            acc_hi[my_first + 1] = 0u;
                            acc_lo[my_first + 1] = 0u;
                            for (int k = 2; k < my_parts; ++k) {     // very wide weight ranges: up to 4 limbs
                                const long long vk = (long long)(((unsigned long long)acc_hi[my_first + k] << 32) | acc_lo[my_first + k]);
                                sum += (double)vk * __ldg(a.tab.acc_scale + my_first + k);
                                acc_hi[my_first + k] = 0u;
                                acc_lo[my_first + k] = 0u;
                            }
                        }
                    } else {
                        sum = (double)acc_lo[my_first];
                    }
                    acc_lo[my_first] = 0u;
                    // (a zero numerator sends DDIV down its out-of-line special-case path for the whole warp)
                    const double q = __ddiv_rn(nonzero_numerator(sum), my_max_auc);
                    res = sum == 0.0 ? 0.0 : q;
                }
                a.out_auc[cell * (int64_t)a.tab.n_regulons + tid] = res;
            }
        } else {
            write_aucs_packed<ACC64, kF>(a.tab, acc_lo, acc_hi, a.out_auc + cell * (int64_t)a.tab.n_regulons);
        }
        // no barrier here: the epilogue only touches the accumulators, and every path of the next cell passes
        // several block barriers before its first scatter
    }
}


// aucell4r path: AUCs from an existing ranking matrix (reference aucell.py:98-182; ctxcore auc2d): the same
// scatter, driven straight from the ranking row (T is indexed by the ranking matrix' own columns).
//
// Only ~auc_threshold of a row's ranks lie below the cutoff, spread at random over the row.  Scattering them where
// they are found leaves a warp waiting on one or two lanes' table rows at every step, so the hits are first collected
// into a compact list in shared memory ((column, cutoff - rank), slots handed out per 128-bit vector with three warp
// votes) and then applied one hit per thread with every lane busy.  Hits beyond the list's capacity (a row with more
// ranks below the cutoff than a ranking has: arbitrary input) are applied where they are found.
constexpr int kHitCap = 4096;             // (column, multiplier) pairs per cell in shared memory

template <bool ACC64>
__global__ void __launch_bounds__(kF) aucell_from_rankings_kernel(const uint32_t* __restrict__ rnk, int64_t n_cells,
                                                                  int64_t row_stride, int n_genes, int cutoff,
                                                                  GeneTableDev tab, double* __restrict__ out_auc) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint2* const hits = reinterpret_cast<uint2*>(smem);                       // [kHitCap]
    uint32_t* const acc_lo = reinterpret_cast<uint32_t*>(smem + kHitCap * sizeof(uint2));
    uint32_t* const acc_hi = acc_lo + tab.n_acc;
    __shared__ int s_n;
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < 2 * tab.n_acc; i += kF) acc_lo[i] = 0u;
    if (tid == 0) s_n = 0;
    __syncthreads();
    for (int64_t cell = blockIdx.x; cell < n_cells; cell += gridDim.x) {
        const uint32_t* row = rnk + cell * row_stride;
        // 128-bit loads over the 16-byte aligned interior of the row, scalar loads for the ragged edges
        const int head = (int)(((16u - (unsigned)(reinterpret_cast<uintptr_t>(row) & 15u)) & 15u) >> 2);
        const int h = min(head, n_genes);
        const int n4 = (n_genes - h) >> 2;
        if (tid < h) {
            const uint32_t r = row[tid];
            if (r < (uint32_t)cutoff) scatter_packed<ACC64>(tab, acc_lo, acc_hi, (unsigned)tid, (uint32_t)cutoff - r);
        }
        const uint4* row4 = reinterpret_cast<const uint4*>(row + h);
#pragma unroll 1
        for (int base = 0; base < n4; base += kF) {          // uniform trip count: the body votes
            const int q = base + tid;
            uint4 v = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
            if (q < n4) v = __ldg(row4 + q);
            const uint32_t r[4] = {v.x, v.y, v.z, v.w};
            int c = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) c += (r[k] < (uint32_t)cutoff) ? 1 : 0;
            const unsigned b0 = __ballot_sync(0xffffffffu, c & 1);
            const unsigned b1 = __ballot_sync(0xffffffffu, c & 2);
            const unsigned b2 = __ballot_sync(0xffffffffu, c & 4);
            if (b0 | b1 | b2) {
                const unsigned lt = lanemask_lt();
                int wbase = 0;
                if (lane == 0) wbase = atomicAdd(&s_n, __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2));
                wbase = __shfl_sync(0xffffffffu, wbase, 0);
                int idx = wbase + __popc(b0 & lt) + 2 * __popc(b1 & lt) + 4 * __popc(b2 & lt);
                const unsigned j = (unsigned)(h + (q << 2));
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (r[k] < (uint32_t)cutoff) {
                        if (idx < kHitCap) hits[idx] = make_uint2(j + k, (uint32_t)cutoff - r[k]);
                        else scatter_packed<ACC64>(tab, acc_lo, acc_hi, j + k, (uint32_t)cutoff - r[k]);
                        ++idx;
                    }
                }
            }
        }
        for (int j = h + (n4 << 2) + tid; j < n_genes; j += kF) {
            const uint32_t r = row[j];
            if (r < (uint32_t)cutoff) scatter_packed<ACC64>(tab, acc_lo, acc_hi, (unsigned)j, (uint32_t)cutoff - r);
        }
        __syncthreads();
        const int n_hits = min(s_n, kHitCap);
        for (int i = tid; i < n_hits; i += kF) {
            const uint2 e = hits[i];
            scatter_packed<ACC64>(tab, acc_lo, acc_hi, e.x, e.y);
        }
        __syncthreads();
        if (tid == 0) s_n = 0;
        write_aucs_packed<ACC64, kF>(tab, acc_lo, acc_hi, out_auc + cell * (int64_t)tab.n_regulons);
        __syncthreads();
    }
}

}  // namespace aucell

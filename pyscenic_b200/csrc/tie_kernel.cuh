// tie_kernel.cuh -- the AUCell hot path for cells whose values fall into a few tie classes (count data: raw UMI
// counts, log1p / CPM / size-factor normalised counts -- any cell with a few dozen distinct values).  This is the
// kernel the benchmark measures; cells it cannot take (more than `cap` stored entries, negative / NaN values,
// continuous values) are appended to a to-do list and ranked by aucell_fused_kernel (fused_kernel.cuh) in a second
// launch.  Both kernels accumulate the same integers, so which one ranked a cell does not show in the output.
//
// Reference: pyscenic.aucell.aucell = aucell4r(create_rankings(exp_mtx, seed), ...)   src/pyscenic/aucell.py:185-213
//            create_rankings  (pandas sample + rank(method="first", ascending=False))  src/pyscenic/aucell.py:25-51
//            ctxcore.recovery.enrichment4cells -> aucs -> auc2d -> weighted_auc1d     (third-party)
//
// Layout: ONE persistent CTA per SM; it keeps inv_perm (gene column -> shuffled position, uint16) resident in shared
// memory and runs `groups` (four) independent 256-thread cell groups, each with its own working set and its own named
// barrier (bar.sync id, 256), so no phase ever waits for more than eight warps.
//
// Two kernels live here: aucell_tie_kernel (position-ordered list + counting sort by class; rows up to 4,096 entries)
// and, further down, aucell_tie2_kernel (one position bitmap per value class; rows up to 16,384 entries), which runs
// behind the first on the rows beyond its capacity.  launch.cuh: launch_tie.
//
// rank(g) = #{genes with a larger value} + #{genes with the same value at a smaller shuffled position}.  Per cell:
//   P0  value range     of a sample of the whole matrix (tie_range_kernel, one small CTA before this kernel).  Cells with
//                       more than `cap` entries, negative or NaN values: to-do list.
//   P1  bitmap          A[pos] = 1 for every non-zero entry (one atomicOr each); rep[bin] = value bits, where
//                       bin = (max - bits) >> shift maps the sampled value range onto 512 bins -- float bits are
//                       piecewise linear in log2(value), so integer counts up to ~30 get a bin each.
//   P2  two scans       word prefix of A (position order for free: the j-th set bit is the j-th smallest position);
//                       occupied bins -> dense class index, descending value; the (kTieD - 1) LOWEST values get a
//                       class each, everything above them (few genes, many distinct values) is class 0, the tail.
//   P3  position order  every entry finds j = #entries at smaller positions (prefix + popcount) and stores
//                       (pos, class) at by_pos[j]; counts cnt[class][j / ch]; tail entries also go to a short list
//                       with their value bits.  A class bin holding two different values: to-do list.
//   P5  scan            exclusive scan of cnt in (class, chunk) order = rank of the first item of every chunk.
//   P6  rank + scatter  thread t walks chunk t of by_pos in position order with PRIVATE counters
//                       (rank = base[class][t]++: a stable counting sort by class, no atomics, no votes) and adds
//                       W * (cutoff - rank) for every regulon holding the gene: the regulon table is ELL (four slots
//                       per gene in two 128-bit loads, chained overflow), accumulators are pairs of u32 limbs updated
//                       with non-returning shared atomics (the product is split at bit L, so no carry is needed).
//   P7  tail            all pairs over the (<= 128) tail entries: rank = #{(value, pos) before mine}.
//   P8  zeros           fewer positives than the cutoff: zeros fill the remaining ranks in position order.
//   P9  out             out[cell, r] = (double(hi << L + lo) * 2^-s) / max_auc[r], coalesced.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace aucell {

constexpr int kTG = 256;                // threads per cell group
constexpr int kTGW = kTG / 32;
constexpr int kTieNBLog = 9;
constexpr int kTieNB = 1 << kTieNBLog;  // value bins over the cell's own range
constexpr int kTieD = 16;               // value classes per cell (class 0 = tail)
constexpr int kTieFix = 128;            // tail entries per cell
constexpr int kTieMaxGroups = 4;        // cell groups per CTA (1024 threads)
constexpr int kTieP6 = 2;               // items whose table rows are loaded together in P6
constexpr int kTieP6R = 4;              // items per thread and pass of P6 (their ranks are computed together)
constexpr int kTieMiscWords = 128;      // per group: reductions, scan words, flags, chain list lengths
constexpr int kTieU = 4;                // entries per thread and pass of the streaming loops (loads issued together)
constexpr int kTieU1 = 8;               // the same in P1 of aucell_tie_kernel
constexpr int kTieMaxCap = 4096;        // stored entries per cell (20-bit reciprocal in P3; u16 prefixes)
constexpr uint32_t kTiePadU = 0x7FFFu;  // unit weights: accumulator indices stay below this (bit 31 of a row flags a chain)

// aucell_tie_kernel: the arrays of a cell group whose size does not depend on the plan sit at compile-time offsets from
// the group's base (their addresses then cost nothing: immediate offsets instead of a constant load and an add)
constexpr int kTieOMisc = 0;                                         // kTieMiscWords words
constexpr int kTieORep = kTieOMisc + kTieMiscWords * 4;              // [kTieNB + 4] u32
constexpr int kTieOCode = kTieORep + (kTieNB + 4) * 4;               // [kTieNB] u8
constexpr int kTieOCnt = kTieOCode + kTieNB;                         // [kTieD][kTG] u32
constexpr int kTieOFix = kTieOCnt + kTieD * kTG * 4;                 // [kTieFix] uint2
constexpr int kTieOA = kTieOFix + kTieFix * 8;                       // bitmap; the plan-sized arrays follow
static_assert(kTieORep % 16 == 0 && kTieOCode % 16 == 0 && kTieOCnt % 16 == 0 && kTieOFix % 16 == 0 && kTieOA % 16 == 0,
              "16-byte aligned arrays");

// to-do reasons (statistics only)
enum { TIE_ST_FAST = 0, TIE_ST_LONG = 1, TIE_ST_BADVAL = 2, TIE_ST_DUP = 3, TIE_ST_MIXED = 4, TIE_ST_TAIL = 5,
       TIE_ST_SKIPPED = 6, TIE_ST_N = 8 };

struct TieArgs {
    const int64_t* indptr;
    const void* indices;                // int32 (IN_CSR) or uint16 (IN_CSR16 / IN_CSR16_U8)
    const void* data;                   // float, or uint8 codes (IN_CSR16_U8)
    const uint32_t* lut;                // IN_CSR16_U8: [256] float bit patterns of the codes
    int64_t n_cells;
    int n_genes, n_words, wpt;          // wpt = ceil(n_words / kTG)
    int cutoff;
    const uint16_t* inv_perm;           // [n_genes rounded up to 8]
    const uint4* ell_w;                 // weighted: [2 * n_genes] four slots {W, off} per gene; a gene in more than
                                        //           four regulons keeps three and {start, count | 1 << 31} of a chain
    const uint2* xw;                    //           chained slots {W, off}
    const uint4* ell_w16;               // weighted, compact rows (row16 != 0): [n_genes] {W0, W1, W2, idx0 | idx1 << 10 |
                                        //           idx2 << 20 | chain << 31} -- accumulator INDICES (< 1024), one 128-bit
                                        //           load per gene instead of two
    const uint2* xinfo;                 //           [n_genes] {start, count} of the chain of a gene whose row has the flag
    int row16;
    const uint2* ell_u;                 // unit weights: [n_genes] four uint16 accumulators, or two + chain
    const uint16_t* xu;                 //           chained: count, accumulators...
    int L;                              // weighted: products are split at bit L
    int slots4;                         // weighted: the fourth slot of a row may hold a regulon (else only a chain)
    int n_regulons;
    int acc_words;                      // u32 words of accumulators per group
    const int32_t* acc_first;           // [R] first accumulator of regulon r, -1 = invalid regulon
    const uint8_t* acc_parts;           // [R]
    const double* acc_scale;            // [n_acc]
    const void* reg_rec;                // [R] 32-byte records {double scale0, scale1 (0: none), max_auc; int32 first
                                        //     accumulator (-1: invalid regulon), int32 has a second accumulator}
    const uint32_t* range;              // [2] bits of the smallest / largest positive value of a sample (tie_range_kernel)
    const double* max_auc;              // [R]
    double* out_auc;
    int32_t* out_nnz;
    int32_t* todo;                      // cells left to aucell_fused_kernel
    int32_t* todo_n;
    int32_t* long_list;                 // aucell_tie_kernel: rows longer than its cap (up to long_cap entries) go here,
    int32_t* long_n;                    //   for aucell_tie2_kernel (nullptr: everything goes to `todo`)
    int long_cap;
    int prefetch;                       // aucell_tie_kernel: prefetch the next cell's entries into L2 during the scatter
    int skip;                           // timing experiments only (AUCELL_B200_TIE_SKIP): 1 no scatter, 2 no epilogue, stop after 4: P2, 8: P3, 16: P5
    const int32_t* in_list;             // aucell_tie2_kernel: the cells to rank (nullptr: all n_cells)
    const int32_t* in_n;
    unsigned long long* stats;          // [TIE_ST_N]
    int cap;                            // stored entries a group can hold
    int groups;
    // dynamic shared memory: inv_perm at 0, then `groups` blocks of group_stride bytes at off_group0
    int off_group0, group_stride;
    int o_A, o_wpre, o_rep, o_code, o_bpp, o_bpc, o_cnt, o_fix, o_acc, o_misc;
    // aucell_tie2_kernel (class bitmaps): nsw = superwords (4 bitmap words) per bitmap, spt = superwords per thread in P2
    int nsw, spt, o_D, o_pre;
};

// named barrier of one cell group (the builtin is convergent: the compiler reconverges the warp before it)
__device__ __forceinline__ void group_bar(int id) { __barrier_sync_count((unsigned)id, (unsigned)kTG); }

// exclusive scan over the kTG threads of a group; s_w: kTGW words nobody else writes until two barriers later
__device__ __forceinline__ uint32_t group_excl_scan(uint32_t v, uint32_t* s_w, int lane, int gw, int barid,
                                                    uint32_t& total) {
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_w[gw] = incl;
    group_bar(barid);
    uint32_t off = 0u, tot = 0u;
#pragma unroll
    for (int k = 0; k < kTGW; k += 4) {
        const uint4 w = *reinterpret_cast<const uint4*>(s_w + k);
        off += (gw > k ? w.x : 0u) + (gw > k + 1 ? w.y : 0u) + (gw > k + 2 ? w.z : 0u) + (gw > k + 3 ? w.w : 0u);
        tot += w.x + w.y + w.z + w.w;
    }
    total = tot;
    return off + incl - v;
}

// value bits of stored entry i (what orders the entries: for positive floats the bit pattern orders like the value)
template <int IN> struct TieVal {
    using type = float;
    static __device__ __forceinline__ uint32_t bits(const float* p, int64_t i, const uint32_t*) {
        return __float_as_uint(__ldg(p + i));
    }
};
// IN_CSR16_U8: one byte per stored value, an index into a table of up to 256 float bit patterns (the distinct values of
// the matrix: the host pipeline builds it while it narrows the upload to 3 bytes per entry).  1 KB: it lives in L1.
template <> struct TieVal<IN_CSR16_U8> {
    using type = uint8_t;
    static __device__ __forceinline__ uint32_t bits(const uint8_t* p, int64_t i, const uint32_t* lut) {
        return __ldg(lut + __ldg(p + i));
    }
};
template <> struct TieVal<IN_CSR_U8> : TieVal<IN_CSR16_U8> {};

// One gene's row of the regulon table: four slots.  Empty slots add zero to one of 32 dummy accumulators (picked by
// the gene's position, so that a warp's empty slots spread over the banks): the hot loop has no branch per slot.
template <bool UNIT> struct TieRow { uint4 a, b; };       // weighted: {W0, off0, W1, off1}, {W2, off2, W3, off3}
template <> struct TieRow<true> { uint2 a; };             // unit weights: four uint16 accumulator indices

template <bool UNIT>
__device__ __forceinline__ TieRow<UNIT> tie_load_row(const TieArgs& a, unsigned pos) {
    TieRow<UNIT> r;
    if constexpr (UNIT) {
        r.a = __ldg(a.ell_u + pos);
    } else if (a.row16) {
        r.a = __ldg(a.ell_w16 + pos);
        r.b = make_uint4(pos, 0u, 0u, 0u);                 // (the chain, if any, is looked up by position)
    } else {
        r.a = __ldg(a.ell_w + 2 * pos);
        r.b = __ldg(a.ell_w + 2 * pos + 1);
    }
    return r;
}

// acc[off] += low L bits of W * m, acc[off + 4] += the rest (non-returning shared atomics, no carry between them)
__device__ __forceinline__ void tie_slot_w(uint32_t acc_s, uint32_t W, uint32_t off, uint32_t m, int L, uint32_t maskL) {
    const unsigned long long p = (unsigned long long)W * m;
    const uint32_t lo = (uint32_t)p & maskL;
    const uint32_t hi = __funnelshift_r((uint32_t)p, (uint32_t)(p >> 32), L);
    // the two limbs of an accumulator are neighbours; which comes first alternates every 16 accumulators (the table's
    // offset points at the low limb), so that a warp's low-limb atomics spread over all 32 banks, not the even ones
    asm volatile("red.shared.add.u32 [%0], %1;\n\tred.shared.add.u32 [%2], %3;" ::"r"(acc_s + off), "r"(lo), "r"((acc_s + off) ^ 4u), "r"(hi) : "memory");
}
__device__ __forceinline__ void tie_slot_u(uint32_t acc_s, uint32_t a, uint32_t m) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(acc_s + (a << 2)), "r"(m) : "memory");
}

// The chained slots of genes that sit in more than four regulons, for a whole warp at once: every lane passes its own
// chain (cnt = 0: none).  The trip count is the warp's maximum, so that the warp leaves the loop together -- lanes
// leaving a loop one by one stay split for the rest of the cell (nothing downstream merges them again).
template <bool UNIT>
__device__ __forceinline__ void tie_apply_chain_warp(const TieArgs& a, uint32_t acc_s, uint32_t start, uint32_t cnt,
                                                     uint32_t m, uint32_t maskL) {
    const uint32_t maxc = __reduce_max_sync(0xffffffffu, cnt);
#pragma unroll 1
    for (uint32_t k = 0; k < maxc; ++k) {
        if (k < cnt) {
            if constexpr (UNIT) {
                tie_slot_u(acc_s, __ldg(a.xu + start + 1 + k), m);
            } else {
                const uint2 sl = __ldg(a.xw + start + k);
                tie_slot_w(acc_s, sl.x, sl.y, m, a.L, maskL);
            }
        }
    }
}

// The direct slots of a row (branch-free).  Returns true when the row continues in a chain: start / cnt then say where.
template <bool UNIT>
__device__ __forceinline__ bool tie_apply_direct(const TieArgs& a, uint32_t acc_s, const TieRow<UNIT>& row, uint32_t m,
                                                 uint32_t maskL, uint32_t dummy, uint32_t& start, uint32_t& cnt) {
    if constexpr (UNIT) {
        // (a row that continues in a chain keeps {start} in its last two slots: they add 0 to the first slot's
        // accumulator instead -- no per-lane dummy address to keep in a register)
        const bool chain = (row.a.y >> 31) != 0u;
        (void)dummy;
        tie_slot_u(acc_s, row.a.x & 0xFFFFu, m);
        tie_slot_u(acc_s, row.a.x >> 16, m);
        tie_slot_u(acc_s, chain ? (row.a.x & 0xFFFFu) : (row.a.y & 0xFFFFu), chain ? 0u : m);
        tie_slot_u(acc_s, chain ? (row.a.x & 0xFFFFu) : (row.a.y >> 16), chain ? 0u : m);
        start = row.a.y & 0x7FFFFFFFu;
        cnt = 0u;                                          // (read from xu[start] by the consumer)
        return chain;
    } else if (a.row16) {
        // compact row: three weights and three 10-bit accumulator indices; byte offset of an accumulator's low limb =
        // 8 idx (+ 4 when idx & 16: the limbs swap places every 16 accumulators, see tie_slot_w)
        (void)dummy;
        const uint32_t pk = row.a.w;
        const bool chain = (pk >> 31) != 0u;
        const uint32_t i0 = pk & 1023u, i1 = (pk >> 10) & 1023u, i2 = (pk >> 20) & 1023u;
        tie_slot_w(acc_s, row.a.x, (i0 << 3) + ((i0 & 16u) >> 2), m, a.L, maskL);
        tie_slot_w(acc_s, row.a.y, (i1 << 3) + ((i1 & 16u) >> 2), m, a.L, maskL);
        tie_slot_w(acc_s, row.a.z, (i2 << 3) + ((i2 & 16u) >> 2), m, a.L, maskL);
        start = 0u;
        cnt = 0u;
        if (chain) {
            const uint2 ci = __ldg(a.xinfo + row.b.x);
            start = ci.x;
            cnt = ci.y;
        }
        return chain;
    } else {
        const bool chain = (row.b.w >> 31) != 0u;
        tie_slot_w(acc_s, row.a.x, row.a.y, m, a.L, maskL);
        tie_slot_w(acc_s, row.a.z, row.a.w, m, a.L, maskL);
        tie_slot_w(acc_s, row.b.x, row.b.y, m, a.L, maskL);
        // (a chained row keeps {start, count} in its last slot: it adds 0 to the third slot's accumulator instead)
        (void)dummy;
        if (a.slots4) tie_slot_w(acc_s, chain ? 0u : row.b.z, chain ? row.b.y : row.b.w, m, a.L, maskL);
        start = row.b.z;
        cnt = row.b.w & 0x7FFFFFFFu;
        return chain;
    }
}

// whole row at once, for a whole warp (tail, zero fill): lanes without work pass m = 0 and any valid position
template <bool UNIT>
__device__ __forceinline__ void tie_apply_full_warp(const TieArgs& a, uint32_t acc_s, unsigned pos, uint32_t m,
                                                    uint32_t maskL, uint32_t dummy) {
    if (!__any_sync(0xffffffffu, m != 0u)) return;         // (warp-uniform) nobody has work
    const TieRow<UNIT> row = tie_load_row<UNIT>(a, pos);
    uint32_t start, cnt;
    const bool chain = tie_apply_direct<UNIT>(a, acc_s, row, m, maskL, dummy, start, cnt);
    if constexpr (UNIT) cnt = (chain && m != 0u) ? __ldg(a.xu + start) : 0u;
    tie_apply_chain_warp<UNIT>(a, acc_s, start, (chain && m != 0u) ? cnt : 0u, m, maskL);
}

constexpr int kTieChainCap = kTieNB * 4 / (kTGW * 8);      // chained rows a warp can defer (the list reuses rep[])

// Value range of the matrix from a sample of its stored values (one small CTA before the main kernel):
// range[0] = smallest, range[1] = largest positive value's bits among <= 64 k samples.  The main kernel maps this
// range onto its kTieNB bins; a value outside it lands in an end bin, which the one-value-per-bin check covers.
template <int IN>
__global__ void __launch_bounds__(1024) tie_range_kernel(const int64_t* __restrict__ indptr, int64_t n_cells,
                                                          const void* __restrict__ data, const uint32_t* __restrict__ lut,
                                                          uint32_t* __restrict__ range) {
    using VT = TieVal<IN>;
    const int64_t e0 = indptr[0], total = indptr[n_cells] - e0;
    const typename VT::type* const vals = reinterpret_cast<const typename VT::type*>(data) + e0;
    uint32_t umin = 0xFFFFFFFFu, umax = 0u;
    const int64_t n_s = total < 65536 ? total : 65536;
    for (int64_t k = threadIdx.x; k < n_s; k += 1024) {
        const int64_t i = total <= 65536 ? k : (int64_t)(((unsigned __int128)k * (unsigned __int128)total) >> 16);
        const uint32_t u = VT::bits(vals, i, lut);
        const bool ps = (u << 1) != 0u && u <= 0x7F800000u;
        umin = min(umin, ps ? u : 0xFFFFFFFFu);
        umax = max(umax, ps ? u : 0u);
    }
    __shared__ uint32_t s_min, s_max;
    if (threadIdx.x == 0) { s_min = 0xFFFFFFFFu; s_max = 0u; }
    __syncthreads();
    umin = __reduce_min_sync(0xffffffffu, umin);
    umax = __reduce_max_sync(0xffffffffu, umax);
    if ((threadIdx.x & 31) == 0) { atomicMin(&s_min, umin); atomicMax(&s_max, umax); }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_max < s_min) { s_min = 0x3F800000u; s_max = 0x3F800000u; }      // no positive sample
        range[0] = s_min;
        range[1] = s_max;
    }
}

// codes -> float values for the stored entries of the cells (absolute offsets, like the kernels index them): what the general
// kernel reads when cells of a byte-coded chunk land on the to-do list.
template <int IN>
__global__ void __launch_bounds__(256) tie_decode_kernel(const uint8_t* __restrict__ codes, const uint32_t* __restrict__ lut,
                                                         float* __restrict__ out, const int64_t* __restrict__ indptr,
                                                         int64_t n_cells) {
    const int64_t e0 = indptr[0], e1 = indptr[n_cells];
    for (int64_t i = e0 + (int64_t)blockIdx.x * 256 + threadIdx.x; i < e1; i += (int64_t)gridDim.x * 256)
        out[i] = __uint_as_float(__ldg(lut + __ldg(codes + i)));
}

template <int IN, bool UNIT>
__global__ void __launch_bounds__(kTG* kTieMaxGroups, 1) aucell_tie_kernel(const TieArgs a) {
    using IdxT = typename IdxOf<IN>::type;
    using VT = TieVal<IN>;
    using RowT = TieRow<UNIT>;
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, grp = tid / kTG;
    // the thread's index inside its cell group, made opaque: at 64 registers per thread the compiler would otherwise
    // re-derive it (S2R + shifts) and everything that follows from it all over the cell loop (7 % of the instructions)
    int gtid = tid % kTG;
    asm volatile("" : "+r"(gtid));
    const int lane = gtid & 31, gw = gtid >> 5;

    // inv_perm -> shared memory, once per CTA
    uint16_t* const ip = reinterpret_cast<uint16_t*>(smem);
    {
        const int nv = (a.n_genes + 7) >> 3;
        const uint4* const src = reinterpret_cast<const uint4*>(a.inv_perm);
        for (int i = tid; i < nv; i += blockDim.x) reinterpret_cast<uint4*>(ip)[i] = __ldg(src + i);
    }
    unsigned char* const gb = smem + a.off_group0 + grp * a.group_stride;
    uint32_t* const A = reinterpret_cast<uint32_t*>(gb + kTieOA);          // [n_words + pad] occupied positions
    uint16_t* const wpre = reinterpret_cast<uint16_t*>(gb + a.o_wpre);    // [n_words] set bits before the word
    uint32_t* const rep = reinterpret_cast<uint32_t*>(gb + kTieORep);      // [kTieNB + 4] value bits of the bin, 0 = empty;
                                                                          // from P6 on: the warps' chain lists
    uint8_t* const code_of = reinterpret_cast<uint8_t*>(gb + kTieOCode);   // [kTieNB] class of the bin
    uint16_t* const bpp = reinterpret_cast<uint16_t*>(gb + a.o_bpp);      // [cap + 1] position of the j-th entry (chunk-interleaved)
    uint8_t* const bpc = reinterpret_cast<uint8_t*>(gb + a.o_bpc);        // [cap + 1] its class
    uint32_t* const cnt = reinterpret_cast<uint32_t*>(gb + kTieOCnt);      // [kTieD][kTG] counters, then bases
    uint2* const fix = reinterpret_cast<uint2*>(gb + kTieOFix);            // [kTieFix] (value bits, pos) of the tail
    uint32_t* const acc = reinterpret_cast<uint32_t*>(gb + a.o_acc);      // accumulators, then 32 dummies
    uint32_t* const misc = reinterpret_cast<uint32_t*>(gb + kTieOMisc);    // kTieMiscWords words, see kM* below
    constexpr int kMScanA = 0;                 // [kTGW] scan words of P2
    constexpr int kMScanB = kTGW;              // [kTGW] scan words of P5
    constexpr int kMTail = 2 * kTGW;           // tail count
    constexpr int kMMixed = 2 * kTGW + 1;      // two values in one class bin
    constexpr int kMBad = 2 * kTGW + 2;        // negative / NaN value
    constexpr int kMValid = 2 * kTGW + 3;      // entries P1 put into the bitmap
    constexpr int kMChain = 2 * kTGW + 4;      // [kTGW] chain list lengths
    static_assert(kMChain + kTGW <= kTieMiscWords && kTGW % 4 == 0, "misc layout");
    for (int i = gtid; i < a.acc_words; i += kTG) acc[i] = 0u;
    if (gtid == 0) {                      // the dummy slot invalid lanes write to must always hold a valid (pos, class)
        bpp[a.cap] = 0;
        bpc[a.cap] = 0;
    }
    if (gtid < kTGW) misc[kMChain + gtid] = 0u;
    const uint32_t acc_s = (uint32_t)__cvta_generic_to_shared(acc);
    uint2* const chain_list = reinterpret_cast<uint2*>(rep) + gw * kTieChainCap;
    __syncthreads();

    const int barid = 1 + grp;
    const int n_genes = a.n_genes, n_words = a.n_words, cutoff = a.cutoff, R = a.n_regulons, cap = a.cap;
    // lanes without an entry OR 0 into a spare word behind the bitmap and add 0 to a counter of the tail class: one
    // spare PER LANE -- 32 lanes hitting one address would serialise the atomic 32-way
    const int dummy_word = ((n_words + 3) & ~3) + lane;
    const uint32_t maskL = UNIT ? 0u : ((1u << a.L) - 1u);
    // dummy accumulator of this lane (rows that continue in a chain; unit: index, weighted: byte offset)
    const uint32_t dummy = UNIT ? (uint32_t)(a.acc_words - 32 + lane) : (uint32_t)(a.acc_words - 64 + 2 * lane) * 4u;
    // value bits -> bin: (gmax - bits) >> shift, clamped (the range comes from a sample)
    const uint32_t gmax = __ldg(a.range + 1);
    const int shift = max(0, (32 - __clz((int)(gmax - __ldg(a.range)))) - kTieNBLog);
    int seen = 0, handed = 0;             // cells this group looked at / gave to the to-do list (group-uniform)

    auto hand_over = [&](int64_t cell, int why) {
        if (gtid == 0) {
            const int k = atomicAdd(a.todo_n, 1);
            a.todo[k] = (int32_t)cell;
            if (a.stats != nullptr) atomicAdd(a.stats + why, 1ull);
        }
        ++handed;
    };
    auto bin_of = [&](uint32_t u) -> uint32_t {
        const int d = max((int)(gmax - u), 0);
        return min((uint32_t)d >> shift, (uint32_t)(kTieNB - 1));
    };

    for (int64_t cell = (int64_t)blockIdx.x * a.groups + grp; cell < a.n_cells; cell += (int64_t)gridDim.x * a.groups) {
        // data this kernel does not suit (continuous values): stop looking, hand everything over
        if (seen >= 16 && handed * 8 >= seen * 7) {
            hand_over(cell, TIE_ST_SKIPPED);
            continue;
        }
        ++seen;
        const int64_t rb = __ldg(a.indptr + cell);
        const int64_t n64 = __ldg(a.indptr + cell + 1) - rb;
        if (n64 < 0 || n64 > (int64_t)cap) {
            if (a.long_list != nullptr && n64 > 0 && n64 <= (int64_t)a.long_cap) {      // aucell_tie2_kernel takes it
                if (gtid == 0) a.long_list[atomicAdd(a.long_n, 1)] = (int32_t)cell;
            } else {
                hand_over(cell, TIE_ST_LONG);
            }
            continue;
        }
        const int n = (int)n64;
        const int iters = (n + kTieU * kTG - 1) / (kTieU * kTG);          // uniform trip count of the entry loops
        const IdxT* const cols = reinterpret_cast<const IdxT*>(a.indices) + rb;
        const typename VT::type* const vals = reinterpret_cast<const typename VT::type*>(a.data) + rb;

        // ---------------- clear the cell's working set -----------------------------------------------------------
        {
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
            for (int i = gtid; i < ((n_words + 3) >> 2); i += kTG) reinterpret_cast<uint4*>(A)[i] = z;
#pragma unroll
            for (int i = gtid; i < kTieNB / 4; i += kTG) reinterpret_cast<uint4*>(rep)[i] = z;
#pragma unroll
            for (int i = gtid; i < kTieD * kTG / 4; i += kTG) reinterpret_cast<uint4*>(cnt)[i] = z;
            if (gtid == 0) *reinterpret_cast<uint4*>(misc + kMTail) = z;      // tail count, mixed, bad, valid
        }
        group_bar(barid);

        // ---------------- P1: position bitmap, one representative value per bin ----------------------------------
        // branch-free: lanes without an entry (past the row's end, explicit zeros, a column out of range) store into
        // the spare bin rep[kTieNB] and OR 0 into the spare word behind the bitmap
        {
            uint32_t bad = 0u;
            int nvalid = 0;
#pragma unroll 1
            // (kTieU1 loads of each array in flight per thread: this is the pass that waits for DRAM)
            const int iters1 = (n + kTieU1 * kTG - 1) / (kTieU1 * kTG);
            for (int it = 0; it < iters1; ++it) {
                const int base = it * (kTieU1 * kTG) + gtid;
                uint32_t u[kTieU1];
                unsigned col[kTieU1];
#pragma unroll
                for (int k = 0; k < kTieU1; ++k) {
                    const bool in = base + k * kTG < n;
                    u[k] = in ? VT::bits(vals, base + k * kTG, a.lut) : 0u;
                    col[k] = in ? (unsigned)__ldg(cols + base + k * kTG) : 0u;
                }
#pragma unroll
                for (int k = 0; k < kTieU1; ++k) {
                    const bool nz = (u[k] << 1) != 0u;                     // +0.0 / -0.0 are not entries
                    const bool valid = nz && col[k] < (unsigned)n_genes;
                    bad |= (nz && u[k] > 0x7F800000u) ? 1u : 0u;           // negative or NaN: not for this kernel
                    nvalid += nz ? 1 : 0;                                  // (an invalid column leaves the bit count short: P2 hands the cell over)
                    const unsigned pos = ip[valid ? col[k] : 0u];
                    rep[valid ? bin_of(u[k]) : (uint32_t)kTieNB] = u[k];
                    atomicOr(&A[valid ? (pos >> 5) : (unsigned)dummy_word], valid ? (1u << (pos & 31)) : 0u);
                }
            }
            nvalid = __reduce_add_sync(0xffffffffu, nvalid);
            bad = __reduce_or_sync(0xffffffffu, bad);
            if (lane == 0) {
                atomicAdd(&misc[kMValid], (uint32_t)nvalid);
                if (bad) misc[kMBad] = 1u;
            }
        }
        group_bar(barid);

        // ---------------- P2: word prefix of the bitmap; occupied bins -> classes ---------------------------------
        int d_total, npos;
        {
            const int w0 = gtid * a.wpt, w1 = min(w0 + a.wpt, n_words);
            uint32_t local = 0;
#pragma unroll 1
            for (int w = w0; w < w1; ++w) local += __popc(A[w]);
            constexpr int kBinsPT = kTieNB / kTG;         // bins per thread
            static_assert(kBinsPT == 2 || kBinsPT == 4, "two or four bins per thread");
            uint32_t occ;
            if constexpr (kBinsPT == 4) {
                const uint4 q0 = reinterpret_cast<const uint4*>(rep)[gtid];
                occ = (q0.x ? 1u : 0u) | (q0.y ? 2u : 0u) | (q0.z ? 4u : 0u) | (q0.w ? 8u : 0u);
            } else {
                const uint2 q0 = reinterpret_cast<const uint2*>(rep)[gtid];
                occ = (q0.x ? 1u : 0u) | (q0.y ? 2u : 0u);
            }
            // bin 0 also takes every value above the sampled range: it always belongs to the tail class (exact order)
            if (gtid == 0) occ &= ~1u;
            uint32_t total;
            const uint32_t excl = group_excl_scan((local << 16) | (uint32_t)__popc(occ), misc + kMScanA, lane, gw, barid, total);
            d_total = (int)(total & 0xFFFFu);
            npos = (int)(total >> 16);
            uint32_t run = excl >> 16;
#pragma unroll 1
            for (int w = w0; w < w1; ++w) {
                wpre[w] = (uint16_t)run;
                run += __popc(A[w]);
            }
            // class of my bins: index among the occupied bins (descending value) minus the tail's share
            const int excess = max(0, d_total - (kTieD - 1));
            int code = (int)(excl & 0xFFFFu);
            uint32_t packed = 0u;
#pragma unroll
            for (int k = 0; k < kBinsPT; ++k) {
                const int cls = code < excess ? 0 : code - excess + 1;
                packed |= ((occ >> k) & 1u) ? (uint32_t)cls << (8 * k) : 0u;
                code += (occ >> k) & 1u;
            }
            if constexpr (kBinsPT == 4) reinterpret_cast<uint32_t*>(code_of)[gtid] = packed;
            else reinterpret_cast<uint16_t*>(code_of)[gtid] = (uint16_t)packed;
            const bool bad = misc[kMBad] != 0u;
            const bool dup = (int)misc[kMValid] != npos;   // a column stored twice, or a column index out of range
            group_bar(barid);
            if (bad || dup) {
                hand_over(cell, bad ? TIE_ST_BADVAL : TIE_ST_DUP);
                continue;
            }
        }

        if (a.skip & 4) continue;                          // (timing experiments only: stop after P2)
        // ---------------- P3: entries into position order; class counts per chunk; the tail list ------------------
        const int ch = max(1, (npos + kTG - 1) / kTG);     // items per chunk (thread) in P6
        const uint32_t magic = ((1u << 20) + (uint32_t)ch - 1u) / (uint32_t)ch;
        {
            uint32_t mixed = 0u;
#pragma unroll 1
            for (int it = 0; it < iters; ++it) {
                const int base = it * (kTieU * kTG) + gtid;
                uint32_t u[kTieU];
                unsigned col[kTieU];
#pragma unroll
                for (int k = 0; k < kTieU; ++k) {
                    const bool in = base + k * kTG < n;
                    u[k] = in ? VT::bits(vals, base + k * kTG, a.lut) : 0u;
                    col[k] = in ? (unsigned)__ldg(cols + base + k * kTG) : 0u;
                }
                bool any_tail = false;
                uint32_t cls_k[kTieU], pos_k[kTieU];
#pragma unroll
                for (int k = 0; k < kTieU; ++k) {
                    const bool valid = (u[k] << 1) != 0u;  // (columns are in range: P2 checked the bit count)
                    const unsigned pos = ip[valid ? col[k] : 0u];
                    const unsigned w = pos >> 5;
                    const uint32_t jv = (uint32_t)wpre[w] + (uint32_t)__popc(A[w] & ((1u << (pos & 31)) - 1u));
                    const uint32_t bin = bin_of(u[k]);
                    const uint32_t cls = valid ? (uint32_t)code_of[bin] : 0u;
                    // by_pos is stored chunk-interleaved -- item k of chunk c at k * kTG + c --, so that the threads of a
                    // warp (one chunk each) read neighbouring elements in P6 instead of elements `ch` apart
                    const uint32_t chunk = valid ? ((jv * magic) >> 20) : (uint32_t)lane;
                    const uint32_t j = valid ? (jv - chunk * (uint32_t)ch) * kTG + chunk : (uint32_t)cap;   // else: the spare slot
                    bpp[j] = (uint16_t)pos;
                    bpc[j] = (uint8_t)cls;
                    atomicAdd(&cnt[cls * kTG + chunk], valid ? 1u : 0u);
                    mixed |= (valid && cls != 0u && rep[bin] != u[k]) ? 1u : 0u;   // two values in one class bin
                    any_tail |= valid && cls == 0u;
                    cls_k[k] = valid ? cls : 1u;
                    pos_k[k] = pos;
                }
                if (__any_sync(0xffffffffu, any_tail)) {   // rare: the largest values, beyond the kTieD - 1 classes
#pragma unroll
                    for (int k = 0; k < kTieU; ++k) {
                        if (cls_k[k] == 0u) {
                            const uint32_t f = atomicAdd(&misc[kMTail], 1u);
                            if (f < (uint32_t)kTieFix) fix[f] = make_uint2(u[k], pos_k[k]);
                        }
                    }
                }
            }
            if (mixed) misc[kMMixed] = 1u;
        }
        group_bar(barid);
        const int n_tail = (int)misc[kMTail];
        if (misc[kMMixed] != 0u || n_tail > kTieFix) {
            const int why = misc[kMMixed] != 0u ? TIE_ST_MIXED : TIE_ST_TAIL;
            group_bar(barid);                              // the flags are cleared by the next cell: read them first
            hand_over(cell, why);
            continue;
        }

        if (a.skip & 8) continue;                          // (timing experiments only: stop after P3)
        // ---------------- P5: exclusive scan of the counters in (class, chunk) order ------------------------------
        {
            // thread t scans the kTieD counters [kTieD t, kTieD t + kTieD) of class kTieD t / kTG: four 128-bit vectors.
            // 64 bytes per thread put four lanes of every quarter-warp on the same banks when all lanes touch "their
            // k-th vector" together (4x the wavefronts, 660 of the cell's 5,500): lane t therefore touches its vectors
            // in the rotated order (k + t / 2) mod 4, which spreads a quarter-warp over all eight bank groups.
            constexpr int kVec = kTieD / 4;               // 128-bit vectors of counters per thread
            static_assert(kVec == 4 && kTG % kTieD == 0, "counter slices");
            const int n_cls = min(d_total, kTieD - 1) + 1;
            const bool live = (gtid * kTieD) / kTG < n_cls;
            const int rot = (gtid >> 1) & 3;
            uint4* const mine = reinterpret_cast<uint4*>(cnt) + gtid * kVec;
            uint32_t q[kVec];                             // q[k]: sum of the vector touched at step k, vector (k + rot) & 3
#pragma unroll
            for (int k = 0; k < kVec; ++k) {
                const uint4 c = live ? mine[(k + rot) & 3] : make_uint4(0u, 0u, 0u, 0u);
                q[k] = c.x + c.y + c.z + c.w;
            }
            uint32_t total;
            const uint32_t run0 = group_excl_scan(q[0] + q[1] + q[2] + q[3], misc + kMScanB, lane, gw, barid, total);
            if (live) {
#pragma unroll
                for (int k = 0; k < kVec; ++k) {
                    const int jk = (k + rot) & 3;         // the vector of this step; what precedes it inside the slice:
                    uint32_t run = run0;
#pragma unroll
                    for (int k2 = 0; k2 < kVec; ++k2) run += (((k2 + rot) & 3) < jk) ? q[k2] : 0u;
                    const uint4 c = mine[jk];             // (again: cheaper than sixteen live registers across the scan)
                    uint4 o;
                    o.x = run; run += c.x;
                    o.y = run; run += c.y;
                    o.z = run; run += c.z;
                    o.w = run;
                    mine[jk] = o;
                }
            }
        }
        group_bar(barid);

        if (a.skip & 16) continue;                         // (timing experiments only: stop after P5)
        // the next cell's stored entries on their way from DRAM to L2 while this cell scatters: its first pass (P1) then
        // waits for L2, not for DRAM (one 128-byte line per thread and array)
        if (a.prefetch) {
            const int64_t nxt = cell + (int64_t)gridDim.x * a.groups;
            if (nxt < a.n_cells) {
                const int64_t nb = __ldg(a.indptr + nxt);
                const int nn = (int)min((int64_t)cap, __ldg(a.indptr + nxt + 1) - nb);
                const char* const pc = reinterpret_cast<const char*>(reinterpret_cast<const IdxT*>(a.indices) + nb);
                const char* const pv = reinterpret_cast<const char*>(reinterpret_cast<const typename VT::type*>(a.data) + nb);
                for (int o = gtid * 128; o < nn * (int)sizeof(IdxT); o += kTG * 128)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(pc + o));
                for (int o = gtid * 128; o < nn * (int)sizeof(typename VT::type); o += kTG * 128)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(pv + o));
            }
        }
        // ---------------- P6: rank in position order with private counters, scatter -------------------------------
        // kTieP6 items per pass, the same number of passes for every thread of the group (threads past their chunk's
        // end work on the spare slot with multiplier 0): ranks first -- the counters are private, so this is a short
        // dependent chain --, then the table rows, then the atomics.  No branch per slot; the few rows that continue
        // in a chain are put on the warp's list and applied afterwards.
        {
            const int j0 = gtid * ch, j1 = min(j0 + ch, npos);
            const int passes = (ch + kTieP6R - 1) / kTieP6R;
#pragma unroll 1
            for (int it = 0; it < passes; ++it) {
                const int j = j0 + it * kTieP6R;
                // ranks of kTieP6R items at once: their (position, class) and their class counters are loaded together
                // -- the counter of a class that appears twice in the batch is bumped in registers, so no load waits
                // for an earlier store that might alias it
                uint32_t mm[kTieP6R];
                unsigned ps[kTieP6R], cs[kTieP6R];
                int rr[kTieP6R];
#pragma unroll
                for (int k = 0; k < kTieP6R; ++k) {
                    const bool in = j + k < j1;
                    const int jj = in ? (it * kTieP6R + k) * kTG + gtid : cap;
                    ps[k] = bpp[jj];
                    cs[k] = in ? (unsigned)bpc[jj] : 0u;       // (the tail class: its counters are never read)
                }
#pragma unroll
                for (int k = 0; k < kTieP6R; ++k) rr[k] = (int)cnt[cs[k] * kTG + gtid];
#pragma unroll
                for (int k = 1; k < kTieP6R; ++k) {
                    int before = 0;
#pragma unroll
                    for (int q = 0; q < k; ++q) before += (cs[q] == cs[k]) ? 1 : 0;
                    rr[k] += before;
                }
#pragma unroll
                for (int k = 0; k < kTieP6R; ++k) {
                    cnt[cs[k] * kTG + gtid] = (uint32_t)(rr[k] + 1);      // (in order: the last of a class wins)
                    mm[k] = (cs[k] != 0u && rr[k] < cutoff && !(a.skip & 1)) ? (uint32_t)(cutoff - rr[k]) : 0u;
                }
#pragma unroll
                for (int h = 0; h < kTieP6R; h += kTieP6) {
                uint32_t m[kTieP6];
                RowT row[kTieP6];
#pragma unroll
                for (int k = 0; k < kTieP6; ++k) {
                    m[k] = mm[h + k];
                    if (m[k]) row[k] = tie_load_row<UNIT>(a, ps[h + k]);
                }
#pragma unroll
                for (int k = 0; k < kTieP6; ++k) {
                    if (m[k] == 0u) continue;               // not below the cutoff: nothing to add
                    uint32_t start, cntc;
                    const bool chain = tie_apply_direct<UNIT>(a, acc_s, row[k], m[k], maskL, dummy, start, cntc);
                    if (chain) {
                        const uint32_t at = atomicAdd(&misc[kMChain + gw], 1u);
                        if (at < (uint32_t)kTieChainCap) {
                            chain_list[at] = make_uint2(start, (cntc << 16) | m[k]);
                        } else {                            // list full (never seen): apply it here, lane by lane
                            if constexpr (UNIT) {
                                cntc = __ldg(a.xu + start);
                                for (uint32_t q = 1; q <= cntc; ++q) tie_slot_u(acc_s, __ldg(a.xu + start + q), m[k]);
                            } else {
                                for (uint32_t q = 0; q < cntc; ++q) {
                                    const uint2 sl = __ldg(a.xw + start + q);
                                    tie_slot_w(acc_s, sl.x, sl.y, m[k], a.L, maskL);
                                }
                            }
                        }
                    }
                }
                }
            }
            // the warp's deferred chains, one per lane
            __syncwarp();
            const int n_list = min((int)misc[kMChain + gw], kTieChainCap);
#pragma unroll 1
            for (int base = 0; base < n_list; base += 32) {
                const int e = base + lane;
                const uint2 ent = e < n_list ? chain_list[e] : make_uint2(0u, 0u);
                uint32_t cntc = ent.y >> 16;
                if constexpr (UNIT) cntc = e < n_list ? __ldg(a.xu + ent.x) : 0u;
                tie_apply_chain_warp<UNIT>(a, acc_s, ent.x, cntc, ent.y & 0xFFFFu, maskL);
            }
            __syncwarp();
            if (lane == 0) misc[kMChain + gw] = 0u;
        }
        // ---------------- P7: the tail (largest values): exact order by (value, position) -------------------------
#pragma unroll 1
        for (int f0 = 0; f0 < n_tail; f0 += kTG) {         // (uniform trip counts: see tie_apply_chain_warp)
            if (f0 + (gw << 5) >= n_tail) break;            // (warp-uniform) this warp has no tail entry
            const int f = f0 + gtid;
            const uint2 mine = f < n_tail ? fix[f] : make_uint2(0u, 0u);
            int r = 0;
#pragma unroll 4
            for (int g = 0; g < n_tail; ++g) {
                const uint2 o = fix[g];
                r += (o.x > mine.x || (o.x == mine.x && o.y < mine.y)) ? 1 : 0;
            }
            const uint32_t m = (f < n_tail && r < cutoff) ? (uint32_t)(cutoff - r) : 0u;
            // (idle lanes: distinct positions -- 32 lanes adding 0 to ONE accumulator would serialise 32-way)
            tie_apply_full_warp<UNIT>(a, acc_s, f < n_tail ? mine.y : (unsigned)min(lane, n_genes - 1), m, maskL, dummy);
        }
        // ---------------- P8: zeros fill the ranks npos .. cutoff-1 in position order ------------------------------
        // the q_zero-th zero sits at a position below q_zero + npos = cutoff: every thread scans positions [0, cutoff)
        const int q_zero = cutoff - npos;
        if (q_zero > 0) {
#pragma unroll 1
            for (int p0 = 0; p0 < cutoff; p0 += kTG) {
                const int p = min(p0 + gtid, n_genes - 1);
                const int w = p >> 5;
                const uint32_t bw = A[w];
                const int z = (w << 5) - (int)wpre[w] + (p & 31) - __popc(bw & ((1u << (p & 31)) - 1u));   // zeros before p
                const bool is_zero = p0 + gtid < cutoff && !((bw >> (p & 31)) & 1u) && z < q_zero;
                tie_apply_full_warp<UNIT>(a, acc_s, (unsigned)p, is_zero ? (uint32_t)(q_zero - z) : 0u, maskL, dummy);
            }
        }
        if (gtid == 0) {
            if (a.out_nnz != nullptr) a.out_nnz[cell] = npos;
            if (a.stats != nullptr) atomicAdd(a.stats + TIE_ST_FAST, 1ull);
        }
        // ---------------- P9: AUCs out, accumulators cleared -------------------------------------------------------
        // out[cell, r] = (double(acc) * 2^-s [+ double(acc') * 2^-s']) / max_auc[r].  Branch-free: a regulon without a
        // second accumulator reads one of the dummies (always zero) with scale 0.  The per-regulon constants (one 32-byte
        // record) do not depend on the accumulators: they are loaded before the barrier.  Same number of passes for
        // every thread, and the warp is reconverged first (P7 / P8 leave it split).
        __syncwarp();
        if (a.skip & 2) { group_bar(barid); continue; }
        double* const orow = a.out_auc + cell * (int64_t)R;
        const int passes9 = (R + kTG - 1) / kTG;
        int r9 = gtid;
        const uint4* const recs = reinterpret_cast<const uint4*>(a.reg_rec);
        const uint4 rec_none = make_uint4(0u, 0u, 0u, 0u), rec_none2 = make_uint4(0u, 0x3FF00000u, 0xFFFFFFFFu, 0u);
        uint4 ra_n = r9 < R ? __ldg(recs + 2 * r9) : rec_none;             // s0, s1
        uint4 rb_n = r9 < R ? __ldg(recs + 2 * r9 + 1) : rec_none2;        // max_auc, first accumulator (-1: none), has2
        group_bar(barid);
#pragma unroll 1
        for (int it = 0; it < passes9; ++it, r9 += kTG) {
            const uint4 ra = ra_n, rb = rb_n;
            if (it + 1 < passes9) {                        // next pass' constants while this one divides
                const int rn = r9 + kTG;
                ra_n = rn < R ? __ldg(recs + 2 * rn) : rec_none;
                rb_n = rn < R ? __ldg(recs + 2 * rn + 1) : rec_none2;
            }
            const int fx = (int)rb.z;
            const int f = fx < 0 ? 0 : fx;                 // (invalid regulons read accumulator 0 and select 0.0 below)
            const double mx = __hiloint2double((int)rb.y, (int)rb.x);
            double sum;
            if constexpr (UNIT) {
                sum = (double)acc[f];
                if (fx >= 0) acc[f] = 0u;
            } else {
                const bool has2 = rb.w != 0u;
                const int f2 = has2 ? f + 1 : ((a.acc_words - 64) >> 1) + lane;
                uint2 lh = *reinterpret_cast<const uint2*>(acc + 2 * f);
                uint2 l2 = *reinterpret_cast<const uint2*>(acc + 2 * f2);
                if (fx >= 0) *reinterpret_cast<uint2*>(acc + 2 * f) = make_uint2(0u, 0u);
                if (has2) *reinterpret_cast<uint2*>(acc + 2 * f2) = make_uint2(0u, 0u);
                if (f & 16) lh = make_uint2(lh.y, lh.x);                  // low limb second (see tie_slot_w)
                if (f2 & 16) l2 = make_uint2(l2.y, l2.x);
                const long long v0 = (long long)(((unsigned long long)lh.y << a.L) + (unsigned long long)lh.x);
                const long long v1 = (long long)(((unsigned long long)l2.y << a.L) + (unsigned long long)l2.x);
                sum = (double)v0 * __hiloint2double((int)ra.y, (int)ra.x);
                sum += (double)v1 * __hiloint2double((int)ra.w, (int)ra.z);
            }
            // a zero numerator (no gene of the regulon below the cutoff: common) would send DDIV down its out-of-line
            // special-case path for the whole warp: divide something harmless instead
            const double q = __ddiv_rn(nonzero_numerator(sum), mx);
            if (r9 < R) orow[r9] = (sum == 0.0 || fx < 0) ? 0.0 : q;
        }
        // no barrier: the next cell clears its working set and passes a group barrier before it touches anything else
    }
}

// =====================================================================================================================
// aucell_tie2_kernel -- the same job with one position bitmap PER VALUE CLASS instead of a position-ordered list:
//   rank(g) = base[class] + #{genes of the same class at smaller shuffled positions}
//           = base[class] + prefix[class][superword] + popcount(bitmap[class][superword] below the position).
// No list in position order, no counting sort, no second placement pass: an entry sets one bit (P1) and later reads one
// 128-bit superword of its class bitmap and one 16-bit prefix (P3); in between, one pass over the bitmaps builds the
// prefixes of all classes at once (P2).  kT2C classes (the lowest distinct values of the cell) have a bitmap; every
// value above them is the tail (<= kT2Fix entries, exact order by pairs).  Same accumulators, same regulon table,
// same epilogue as aucell_tie_kernel: the outputs are bit-identical.
//   P0  classes         rep[bin] = value bits over the cell's stored values; occupied bins -> class (descending value)
//   P1  bitmaps         D[class][pos] = 1 (one atomicOr per entry); tail entries also go to the pair list
//   P2  prefixes        per superword (128 positions): popcounts of the kT2C class words and of their union -> one
//                       packed scan (8 x 16 bit) -> pre[superword][class]; D[0] becomes the union (occupancy)
//   P3  rank + scatter  every entry: rank from its class bitmap, then its row of the regulon table (as P6 above)
//   P7-P9               tail, zero fill, AUCs out: as above
constexpr int kT2C = 7;                 // value classes with a bitmap (class 0: the tail; after P2: occupancy)
constexpr int kT2Fix = 256;             // tail entries per cell

// exclusive scan of four packed words over the kTG threads of a group; s_w: kTGW uint4
__device__ __forceinline__ uint4 group_excl_scan4(uint4 v, uint4* s_w, int lane, int gw, int barid, uint4& incl_out) {
    uint4 incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t tx = __shfl_up_sync(0xffffffffu, incl.x, o), ty = __shfl_up_sync(0xffffffffu, incl.y, o);
        const uint32_t tz = __shfl_up_sync(0xffffffffu, incl.z, o), tw = __shfl_up_sync(0xffffffffu, incl.w, o);
        if (lane >= o) { incl.x += tx; incl.y += ty; incl.z += tz; incl.w += tw; }
    }
    if (lane == 31) s_w[gw] = incl;
    group_bar(barid);
    uint4 off = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int k = 0; k < kTGW - 1; ++k) {
        if (k < gw) {                                      // (warp-uniform)
            const uint4 w = s_w[k];
            off.x += w.x; off.y += w.y; off.z += w.z; off.w += w.w;
        }
    }
    incl_out = make_uint4(off.x + incl.x, off.y + incl.y, off.z + incl.z, off.w + incl.w);
    return make_uint4(incl_out.x - v.x, incl_out.y - v.y, incl_out.z - v.z, incl_out.w - v.w);
}

// set bits of a 128-bit superword below bit `sh` (0 <= sh < 128)
__device__ __forceinline__ uint32_t popc_below(const uint4 q, int sh) {
    // __funnelshift_lc(~0, 0, n): the n low bits set, n clamped to 32
    const uint32_t m0 = __funnelshift_lc(0xFFFFFFFFu, 0u, (uint32_t)sh);
    const uint32_t m1 = __funnelshift_lc(0xFFFFFFFFu, 0u, (uint32_t)max(sh - 32, 0));
    const uint32_t m2 = __funnelshift_lc(0xFFFFFFFFu, 0u, (uint32_t)max(sh - 64, 0));
    const uint32_t m3 = __funnelshift_lc(0xFFFFFFFFu, 0u, (uint32_t)max(sh - 96, 0));
    return (uint32_t)(__popc(q.x & m0) + __popc(q.y & m1) + __popc(q.z & m2) + __popc(q.w & m3));
}

template <int IN, bool UNIT>
__global__ void __launch_bounds__(kTG* kTieMaxGroups, 1) aucell_tie2_kernel(const TieArgs a) {
    using IdxT = typename IdxOf<IN>::type;
    using VT = TieVal<IN>;
    using RowT = TieRow<UNIT>;
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, grp = tid / kTG, gtid = tid % kTG, lane = tid & 31, gw = gtid >> 5;

    uint16_t* const ip = reinterpret_cast<uint16_t*>(smem);
    {
        const int nv = (a.n_genes + 7) >> 3;
        const uint4* const src = reinterpret_cast<const uint4*>(a.inv_perm);
        for (int i = tid; i < nv; i += blockDim.x) reinterpret_cast<uint4*>(ip)[i] = __ldg(src + i);
    }
    unsigned char* const gb = smem + a.off_group0 + grp * a.group_stride;
    uint32_t* const D = reinterpret_cast<uint32_t*>(gb + a.o_D);          // [kT2C + 1][4 nsw] class bitmaps, then 32 spare words
    uint4* const D4 = reinterpret_cast<uint4*>(D);
    uint16_t* const pre = reinterpret_cast<uint16_t*>(gb + a.o_pre);      // [nsw][8] set bits before the superword, per class
    uint32_t* const rep = reinterpret_cast<uint32_t*>(gb + a.o_rep);      // [kTieNB + 4]; from P3 on: the warps' chain lists
    uint8_t* const code_of = reinterpret_cast<uint8_t*>(gb + a.o_code);   // [kTieNB] class of the bin
    uint2* const fix = reinterpret_cast<uint2*>(gb + a.o_fix);            // [kT2Fix] (value bits, pos) of the tail
    uint32_t* const acc = reinterpret_cast<uint32_t*>(gb + a.o_acc);
    uint32_t* const misc = reinterpret_cast<uint32_t*>(gb + a.o_misc);
    constexpr int kMScanA = 0;                 // [kTGW] scan words of P0
    constexpr int kMScan4 = kTGW;              // [kTGW] uint4 scan words of P2
    constexpr int kMBase = 5 * kTGW;           // [8] first rank of class c (c >= 1)
    constexpr int kMTail = 5 * kTGW + 8;       // tail count, mixed, bad, stored non-zero values (one uint4)
    constexpr int kMMixed = kMTail + 1;
    constexpr int kMBad = kMTail + 2;
    constexpr int kMValid = kMTail + 3;
    constexpr int kMChain = kMTail + 4;        // [kTGW] chain list lengths
    constexpr int kMNpos = kMChain + kTGW;     // occupied positions
    static_assert(kMNpos < kTieMiscWords && kTGW % 4 == 0 && (kMTail % 4) == 0, "misc layout");
    for (int i = gtid; i < a.acc_words; i += kTG) acc[i] = 0u;
    if (gtid < kTGW) misc[kMChain + gtid] = 0u;
    const uint32_t acc_s = (uint32_t)__cvta_generic_to_shared(acc);
    uint2* const chain_list = reinterpret_cast<uint2*>(rep) + gw * kTieChainCap;
    __syncthreads();

    const int barid = 1 + grp;
    const int n_genes = a.n_genes, cutoff = a.cutoff, R = a.n_regulons, cap = a.cap, nsw = a.nsw;
    const int dummy_word = (kT2C + 1) * 4 * nsw + lane;            // one spare word per lane behind the bitmaps
    const uint32_t maskL = UNIT ? 0u : ((1u << a.L) - 1u);
    const uint32_t dummy = UNIT ? (uint32_t)(a.acc_words - 32 + lane) : (uint32_t)(a.acc_words - 64 + 2 * lane) * 4u;
    const uint32_t gmax = __ldg(a.range + 1);
    const int shift = max(0, (32 - __clz((int)(gmax - __ldg(a.range)))) - kTieNBLog);
    int seen = 0, handed = 0;

    auto hand_over = [&](int64_t cell, int why) {
        if (gtid == 0) {
            const int k = atomicAdd(a.todo_n, 1);
            a.todo[k] = (int32_t)cell;
            if (a.stats != nullptr) atomicAdd(a.stats + why, 1ull);
        }
        ++handed;
    };
    auto bin_of = [&](uint32_t u) -> uint32_t {
        const int d = max((int)(gmax - u), 0);
        return min((uint32_t)d >> shift, (uint32_t)(kTieNB - 1));
    };
    // one item of the scatter: the direct slots now, a chained continuation onto the warp's list
    auto scatter_item = [&](const RowT& row, uint32_t m) {
        uint32_t start, cntc;
        const bool chain = tie_apply_direct<UNIT>(a, acc_s, row, m, maskL, dummy, start, cntc);
        if (chain) {
            const uint32_t at = atomicAdd(&misc[kMChain + gw], 1u);
            if (at < (uint32_t)kTieChainCap) {
                chain_list[at] = make_uint2(start, (cntc << 16) | m);
            } else {                                // list full (never seen): apply it here, lane by lane
                if constexpr (UNIT) {
                    cntc = __ldg(a.xu + start);
                    for (uint32_t q = 1; q <= cntc; ++q) tie_slot_u(acc_s, __ldg(a.xu + start + q), m);
                } else {
                    for (uint32_t q = 0; q < cntc; ++q) {
                        const uint2 sl = __ldg(a.xw + start + q);
                        tie_slot_w(acc_s, sl.x, sl.y, m, a.L, maskL);
                    }
                }
            }
        }
    };

    const int64_t n_work = a.in_list != nullptr ? (int64_t)__ldg(a.in_n) : a.n_cells;
    for (int64_t wi = (int64_t)blockIdx.x * a.groups + grp; wi < n_work; wi += (int64_t)gridDim.x * a.groups) {
        const int64_t cell = a.in_list != nullptr ? (int64_t)__ldg(a.in_list + wi) : wi;
        if (seen >= 16 && handed * 8 >= seen * 7) {        // data this kernel does not suit: stop looking
            hand_over(cell, TIE_ST_SKIPPED);
            continue;
        }
        ++seen;
        const int64_t rb = __ldg(a.indptr + cell);
        const int64_t n64 = __ldg(a.indptr + cell + 1) - rb;
        if (n64 < 0 || n64 > (int64_t)cap) {
            hand_over(cell, TIE_ST_LONG);
            continue;
        }
        const int n = (int)n64;
        const int iters = (n + kTieU * kTG - 1) / (kTieU * kTG);
        const IdxT* const cols = reinterpret_cast<const IdxT*>(a.indices) + rb;
        const typename VT::type* const vals = reinterpret_cast<const typename VT::type*>(a.data) + rb;

        // ---------------- clear ----------------------------------------------------------------------------------
        {
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
            const int nD4 = (kT2C + 1) * nsw + 8;
            for (int i = gtid; i < nD4; i += kTG) D4[i] = z;
#pragma unroll
            for (int i = gtid; i < kTieNB / 4; i += kTG) reinterpret_cast<uint4*>(rep)[i] = z;
            if (gtid == 0) *reinterpret_cast<uint4*>(misc + kMTail) = z;
        }
        group_bar(barid);

        // ---------------- P0: one representative value per bin ----------------------------------------------------
        {
#pragma unroll 1
            for (int it = 0; it < iters; ++it) {
                const int base = it * (kTieU * kTG) + gtid;
                uint32_t u[kTieU];
#pragma unroll
                for (int k = 0; k < kTieU; ++k) u[k] = base + k * kTG < n ? VT::bits(vals, base + k * kTG, a.lut) : 0u;
#pragma unroll
                for (int k = 0; k < kTieU; ++k) rep[(u[k] << 1) != 0u ? bin_of(u[k]) : (uint32_t)kTieNB] = u[k];
            }
        }
        group_bar(barid);
        // occupied bins -> classes: index among the occupied bins (descending value); the kT2C lowest values get a
        // bitmap each, everything above them is class 0, the tail
        {
            constexpr int kBinsPT = kTieNB / kTG;
            static_assert(kBinsPT == 2, "two bins per thread");
            const uint2 q0 = reinterpret_cast<const uint2*>(rep)[gtid];
            uint32_t occ = (q0.x ? 1u : 0u) | (q0.y ? 2u : 0u);
            if (gtid == 0) occ &= ~1u;         // bin 0 also takes every value above the sampled range: always tail
            uint32_t total;
            const uint32_t excl = group_excl_scan((uint32_t)__popc(occ), misc + kMScanA, lane, gw, barid, total);
            const int excess = max(0, (int)total - kT2C);
            int code = (int)excl;
            uint32_t packed = 0u;
#pragma unroll
            for (int k = 0; k < kBinsPT; ++k) {
                const int cls = code < excess ? 0 : code - excess + 1;
                packed |= ((occ >> k) & 1u) ? (uint32_t)cls << (8 * k) : 0u;
                code += (occ >> k) & 1u;
            }
            reinterpret_cast<uint16_t*>(code_of)[gtid] = (uint16_t)packed;
        }
        group_bar(barid);

        // ---------------- P1: class bitmaps; the tail list ---------------------------------------------------------
        {
            uint32_t bad = 0u, mixed = 0u;
            int nnz = 0;
#pragma unroll 1
            for (int it = 0; it < iters; ++it) {
                const int base = it * (kTieU * kTG) + gtid;
                uint32_t u[kTieU];
                unsigned col[kTieU];
#pragma unroll
                for (int k = 0; k < kTieU; ++k) {
                    const bool in = base + k * kTG < n;
                    u[k] = in ? VT::bits(vals, base + k * kTG, a.lut) : 0u;
                    col[k] = in ? (unsigned)__ldg(cols + base + k * kTG) : 0u;
                }
                bool any_tail = false;
                uint32_t cls_k[kTieU], pos_k[kTieU];
#pragma unroll
                for (int k = 0; k < kTieU; ++k) {
                    const bool nz = (u[k] << 1) != 0u;                     // +0.0 / -0.0 are not entries
                    const bool valid = nz && col[k] < (unsigned)n_genes;   // (an invalid column leaves the bit count short)
                    bad |= (nz && u[k] > 0x7F800000u) ? 1u : 0u;           // negative or NaN: not for this kernel
                    nnz += nz ? 1 : 0;
                    const unsigned pos = ip[valid ? col[k] : 0u];
                    const uint32_t bin = bin_of(u[k]);
                    const uint32_t cls = valid ? (uint32_t)code_of[bin] : 0u;
                    mixed |= (valid && cls != 0u && rep[bin] != u[k]) ? 1u : 0u;   // two values in one class bin
                    atomicOr(&D[valid ? cls * (4u * (unsigned)nsw) + (pos >> 5) : (unsigned)dummy_word],
                             valid ? (1u << (pos & 31)) : 0u);
                    any_tail |= valid && cls == 0u;
                    cls_k[k] = valid ? cls : 1u;
                    pos_k[k] = pos;
                }
                if (__any_sync(0xffffffffu, any_tail)) {   // rare: the largest values, beyond the kT2C classes
#pragma unroll
                    for (int k = 0; k < kTieU; ++k) {
                        if (cls_k[k] == 0u) {
                            const uint32_t f = atomicAdd(&misc[kMTail], 1u);
                            if (f < (uint32_t)kT2Fix) fix[f] = make_uint2(u[k], pos_k[k]);
                        }
                    }
                }
            }
            nnz = __reduce_add_sync(0xffffffffu, nnz);
            bad = __reduce_or_sync(0xffffffffu, bad | (mixed << 1));
            if (lane == 0) {
                atomicAdd(&misc[kMValid], (uint32_t)nnz);
                if (bad & 1u) misc[kMBad] = 1u;
                if (bad & 2u) misc[kMMixed] = 1u;
            }
        }
        group_bar(barid);

        // ---------------- P2: prefixes of all class bitmaps at once ------------------------------------------------
        {
            const int s0 = gtid * a.spt, s1 = min(s0 + a.spt, nsw);
            uint4 sum = make_uint4(0u, 0u, 0u, 0u);        // packed: {all | c1 << 16, c2 | c3 << 16, c4 | c5 << 16, c6 | c7 << 16}
            static_assert(kT2C == 7, "eight 16-bit counts in four words");
            auto count_sw = [&](int s, uint4& orw) -> uint4 {
                uint32_t c[kT2C + 1];
                orw = D4[s];
#pragma unroll
                for (int k = 1; k <= kT2C; ++k) {
                    const uint4 q = D4[k * nsw + s];
                    c[k] = (uint32_t)(__popc(q.x) + __popc(q.y) + __popc(q.z) + __popc(q.w));
                    orw.x |= q.x; orw.y |= q.y; orw.z |= q.z; orw.w |= q.w;
                }
                c[0] = (uint32_t)(__popc(orw.x) + __popc(orw.y) + __popc(orw.z) + __popc(orw.w));
                return make_uint4(c[0] | (c[1] << 16), c[2] | (c[3] << 16), c[4] | (c[5] << 16), c[6] | (c[7] << 16));
            };
#pragma unroll 1
            for (int s = s0; s < s1; ++s) {
                uint4 orw;
                const uint4 c = count_sw(s, orw);
                sum.x += c.x; sum.y += c.y; sum.z += c.z; sum.w += c.w;
            }
            uint4 incl;
            uint4 run = group_excl_scan4(sum, reinterpret_cast<uint4*>(misc + kMScan4), lane, gw, barid, incl);
#pragma unroll 1
            for (int s = s0; s < s1; ++s) {
                uint4 orw;
                const uint4 c = count_sw(s, orw);
                reinterpret_cast<uint4*>(pre)[s] = run;
                D4[s] = orw;                               // class 0's bitmap becomes the occupancy bitmap
                run.x += c.x; run.y += c.y; run.z += c.z; run.w += c.w;
            }
            if (gtid == kTG - 1) {                         // totals -> first rank of every class
                const uint32_t n_all = incl.x & 0xFFFFu;
                uint32_t sz[kT2C + 1] = {0u, incl.x >> 16, incl.y & 0xFFFFu, incl.y >> 16, incl.z & 0xFFFFu, incl.z >> 16,
                                         incl.w & 0xFFFFu, incl.w >> 16};
                uint32_t in_cls = 0u;
#pragma unroll
                for (int k = 1; k <= kT2C; ++k) in_cls += sz[k];
                uint32_t b = n_all - in_cls;               // the tail ranks first
                misc[kMBase] = 0u;
#pragma unroll
                for (int k = 1; k <= kT2C; ++k) {
                    misc[kMBase + k] = b;
                    b += sz[k];
                }
                misc[kMNpos] = n_all;
            }
        }
        group_bar(barid);
        const int npos = (int)misc[kMNpos];
        const int n_tail = (int)misc[kMTail];
        {
            const bool bad = misc[kMBad] != 0u, mixed = misc[kMMixed] != 0u;
            const bool dup = (int)misc[kMValid] != npos;   // a column stored twice, or a column index out of range
            if (bad || dup || mixed || n_tail > kT2Fix) {
                group_bar(barid);                          // the flags are cleared by the next cell: read them first
                hand_over(cell, bad ? TIE_ST_BADVAL : dup ? TIE_ST_DUP : mixed ? TIE_ST_MIXED : TIE_ST_TAIL);
                continue;
            }
        }

        // ---------------- P3: rank from the class bitmap, scatter --------------------------------------------------
        {
#pragma unroll 1
            for (int it = 0; it < iters; ++it) {
                const int base = it * (kTieU * kTG) + gtid;
                uint32_t u[kTieU];
                unsigned col[kTieU];
#pragma unroll
                for (int k = 0; k < kTieU; ++k) {
                    const bool in = base + k * kTG < n;
                    u[k] = in ? VT::bits(vals, base + k * kTG, a.lut) : 0u;
                    col[k] = in ? (unsigned)__ldg(cols + base + k * kTG) : 0u;
                }
                uint32_t m[kTieU];
                unsigned ps[kTieU];
#pragma unroll
                for (int k = 0; k < kTieU; ++k) {
                    const bool valid = (u[k] << 1) != 0u;  // (columns are in range: the bit count matched)
                    const unsigned pos = ip[valid ? col[k] : 0u];
                    const uint32_t cls = valid ? (uint32_t)code_of[bin_of(u[k])] : 0u;
                    const unsigned sw = pos >> 7;
                    const uint4 q = D4[cls * (unsigned)nsw + sw];
                    const int r = (int)(misc[kMBase + cls] + (uint32_t)pre[sw * 8u + cls] + popc_below(q, (int)(pos & 127u)));
                    m[k] = (cls != 0u && r < cutoff) ? (uint32_t)(cutoff - r) : 0u;
                    ps[k] = pos;
                }
#pragma unroll
                for (int k0 = 0; k0 < kTieU; k0 += kTieP6) {
                    RowT row[kTieP6];
#pragma unroll
                    for (int k = 0; k < kTieP6; ++k)
                        if (m[k0 + k]) row[k] = tie_load_row<UNIT>(a, ps[k0 + k]);
#pragma unroll
                    for (int k = 0; k < kTieP6; ++k)
                        if (m[k0 + k]) scatter_item(row[k], m[k0 + k]);
                }
            }
            // the warp's deferred chains, one per lane
            __syncwarp();
            const int n_list = min((int)misc[kMChain + gw], kTieChainCap);
#pragma unroll 1
            for (int base = 0; base < n_list; base += 32) {
                const int e = base + lane;
                const uint2 ent = e < n_list ? chain_list[e] : make_uint2(0u, 0u);
                uint32_t cntc = ent.y >> 16;
                if constexpr (UNIT) cntc = e < n_list ? __ldg(a.xu + ent.x) : 0u;
                tie_apply_chain_warp<UNIT>(a, acc_s, ent.x, cntc, ent.y & 0xFFFFu, maskL);
            }
            __syncwarp();
            if (lane == 0) misc[kMChain + gw] = 0u;
        }
        // ---------------- P7: the tail (largest values): exact order by (value, position) -------------------------
#pragma unroll 1
        for (int f0 = 0; f0 < n_tail; f0 += kTG) {
            if (f0 + (gw << 5) >= n_tail) break;            // (warp-uniform) this warp has no tail entry
            const int f = f0 + gtid;
            const uint2 mine = f < n_tail ? fix[f] : make_uint2(0u, 0u);
            int r = 0;
#pragma unroll 4
            for (int g = 0; g < n_tail; ++g) {
                const uint2 o = fix[g];
                r += (o.x > mine.x || (o.x == mine.x && o.y < mine.y)) ? 1 : 0;
            }
            const uint32_t m = (f < n_tail && r < cutoff) ? (uint32_t)(cutoff - r) : 0u;
            // (idle lanes: distinct positions -- 32 lanes adding 0 to ONE accumulator would serialise 32-way)
            tie_apply_full_warp<UNIT>(a, acc_s, f < n_tail ? mine.y : (unsigned)min(lane, n_genes - 1), m, maskL, dummy);
        }
        // ---------------- P8: zeros fill the ranks npos .. cutoff-1 in position order ------------------------------
        const int q_zero = cutoff - npos;
        if (q_zero > 0) {
#pragma unroll 1
            for (int p0 = 0; p0 < cutoff; p0 += kTG) {
                const int p = min(p0 + gtid, n_genes - 1);
                const unsigned sw = (unsigned)p >> 7;
                const uint4 q = D4[sw];
                const int z = p - (int)pre[sw * 8u] - (int)popc_below(q, p & 127);          // zeros before p
                const uint32_t bw = (p & 64) ? ((p & 32) ? q.w : q.z) : ((p & 32) ? q.y : q.x);
                const bool is_zero = p0 + gtid < cutoff && !((bw >> (p & 31)) & 1u) && z < q_zero;
                tie_apply_full_warp<UNIT>(a, acc_s, (unsigned)p, is_zero ? (uint32_t)(q_zero - z) : 0u, maskL, dummy);
            }
        }
        if (gtid == 0) {
            if (a.out_nnz != nullptr) a.out_nnz[cell] = npos;
            if (a.stats != nullptr) atomicAdd(a.stats + TIE_ST_FAST, 1ull);
        }
        // ---------------- P9: AUCs out, accumulators cleared (as aucell_tie_kernel) ---------------------------------
        __syncwarp();
        double* const orow = a.out_auc + cell * (int64_t)R;
        const int passes9 = (R + kTG - 1) / kTG;
        int r9 = gtid;
        const uint4* const recs = reinterpret_cast<const uint4*>(a.reg_rec);
        const uint4 rec_none = make_uint4(0u, 0u, 0u, 0u), rec_none2 = make_uint4(0u, 0x3FF00000u, 0xFFFFFFFFu, 0u);
        uint4 ra_n = r9 < R ? __ldg(recs + 2 * r9) : rec_none;
        uint4 rb_n = r9 < R ? __ldg(recs + 2 * r9 + 1) : rec_none2;
        group_bar(barid);
#pragma unroll 1
        for (int it = 0; it < passes9; ++it, r9 += kTG) {
            const uint4 ra = ra_n, rb = rb_n;
            if (it + 1 < passes9) {
                const int rn = r9 + kTG;
                ra_n = rn < R ? __ldg(recs + 2 * rn) : rec_none;
                rb_n = rn < R ? __ldg(recs + 2 * rn + 1) : rec_none2;
            }
            const int fx = (int)rb.z;
            const int f = fx < 0 ? 0 : fx;
            const double mx = __hiloint2double((int)rb.y, (int)rb.x);
            double sum;
            if constexpr (UNIT) {
                sum = (double)acc[f];
                if (fx >= 0) acc[f] = 0u;
            } else {
                const bool has2 = rb.w != 0u;
                const int f2 = has2 ? f + 1 : ((a.acc_words - 64) >> 1) + lane;
                uint2 lh = *reinterpret_cast<const uint2*>(acc + 2 * f);
                uint2 l2 = *reinterpret_cast<const uint2*>(acc + 2 * f2);
                if (fx >= 0) *reinterpret_cast<uint2*>(acc + 2 * f) = make_uint2(0u, 0u);
                if (has2) *reinterpret_cast<uint2*>(acc + 2 * f2) = make_uint2(0u, 0u);
                if (f & 16) lh = make_uint2(lh.y, lh.x);
                if (f2 & 16) l2 = make_uint2(l2.y, l2.x);
                const long long v0 = (long long)(((unsigned long long)lh.y << a.L) + (unsigned long long)lh.x);
                const long long v1 = (long long)(((unsigned long long)l2.y << a.L) + (unsigned long long)l2.x);
                sum = (double)v0 * __hiloint2double((int)ra.y, (int)ra.x);
                sum += (double)v1 * __hiloint2double((int)ra.w, (int)ra.z);
            }
            const double q = __ddiv_rn(nonzero_numerator(sum), mx);
            if (r9 < R) orow[r9] = (sum == 0.0 || fx < 0) ? 0.0 : q;
        }
    }
}

}  // namespace aucell

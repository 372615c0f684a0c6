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
// memory and runs `groups` independent 128-thread cell groups, each with its own working set and its own named
// barrier (bar.sync id, 128), so no phase ever waits for more than four warps.
//
// rank(g) = #{genes with a larger value} + #{genes with the same value at a smaller shuffled position}.  Per cell:
//   P0  statistics      min / max of the positive values (float bits order like the values), #non-zero; negative or
//                       NaN values, or more than `cap` entries: to-do list.
//   P1  bitmap          A[pos] = 1 for every non-zero entry (one atomicOr each); rep[bin] = value bits, where
//                       bin = (max - bits) >> shift maps the cell's own value range onto 1024 bins -- float bits are
//                       piecewise linear in log2(value), so integer counts up to ~60 get a bin each.
//   P2  two scans       word prefix of A (position order for free: the j-th set bit is the j-th smallest position);
//                       occupied bins -> dense class index, descending value; the (kTieD - 1) LOWEST values get a
//                       class each, everything above them (few genes, many distinct values) is class 0, the tail.
//   P3  position order  every entry finds j = #entries at smaller positions (prefix + popcount) and stores
//                       (pos, class) at by_pos[j]; counts cnt[class][j / ch]; tail entries also go to a short list
//                       with their value bits.  A class bin holding two different values: to-do list.
//   P5  scan            exclusive scan of cnt in (class, chunk) order = rank of the first item of every chunk.
//   P6  rank + scatter  thread t walks chunk t of by_pos in position order with PRIVATE counters
//                       (rank = base[class][t]++: a stable counting sort by class, no atomics, no votes) and adds
//                       W * (cutoff - rank) for every regulon holding the gene: the regulon table is ELL (two slots
//                       per gene in one 128-bit load, chained overflow), accumulators are pairs of u32 limbs updated
//                       with non-returning shared atomics (the product is split at bit L, so no carry is needed).
//   P7  tail            all pairs over the (<= 256) tail entries: rank = #{(value, pos) before mine}.
//   P8  zeros           fewer positives than the cutoff: zeros fill the remaining ranks in position order.
//   P9  out             out[cell, r] = (double(hi << L + lo) * 2^-s) / max_auc[r], coalesced.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace aucell {

constexpr int kTG = 128;                // threads per cell group
constexpr int kTGW = kTG / 32;
constexpr int kTieNBLog = 10;
constexpr int kTieNB = 1 << kTieNBLog;  // value bins over the cell's own range
constexpr int kTieD = 32;               // value classes per cell (class 0 = tail)
constexpr int kTieFix = 256;            // tail entries per cell
constexpr int kTieMaxGroups = 6;        // cell groups per CTA
constexpr int kTieMaxCap = 4096;        // stored entries per cell (20-bit reciprocal in P3; u16 prefixes)
constexpr uint32_t kTiePadU = 0x7FFFu;  // empty slot of a unit-weight table row

// to-do reasons (statistics only)
enum { TIE_ST_FAST = 0, TIE_ST_LONG = 1, TIE_ST_BADVAL = 2, TIE_ST_DUP = 3, TIE_ST_MIXED = 4, TIE_ST_TAIL = 5,
       TIE_ST_SKIPPED = 6, TIE_ST_N = 8 };

struct TieArgs {
    const int64_t* indptr;
    const void* indices;                // int32 (IN_CSR) or uint16 (IN_CSR16 / IN_CSR16_U8)
    const void* data;                   // float, or uint8 (IN_CSR16_U8)
    int64_t n_cells;
    int n_genes, n_words, wpt;          // wpt = ceil(n_words / kTG)
    int cutoff;
    const uint16_t* inv_perm;           // [n_genes rounded up to 8]
    const uint4* ell_w;                 // weighted: [n_genes] {W0, off0 | chain << 31, W1 or start, off1 or count}
    const uint2* xw;                    //           chained slots {W, off}
    const uint2* ell_u;                 // unit weights: [n_genes] four uint16 accumulators, or two + chain
    const uint16_t* xu;                 //           chained: count, accumulators...
    int L;                              // weighted: products are split at bit L
    int n_regulons;
    int acc_words;                      // u32 words of accumulators per group
    const int32_t* acc_first;           // [R] first accumulator of regulon r, -1 = invalid regulon
    const uint8_t* acc_parts;           // [R]
    const double* acc_scale;            // [n_acc]
    const double* max_auc;              // [R]
    double* out_auc;
    int32_t* out_nnz;
    int32_t* todo;                      // cells left to aucell_fused_kernel
    int32_t* todo_n;
    unsigned long long* stats;          // [TIE_ST_N]
    int cap;                            // stored entries a group can hold
    int groups;
    // dynamic shared memory: inv_perm at 0, then `groups` blocks of group_stride bytes at off_group0
    int off_group0, group_stride;
    int o_A, o_wpre, o_rep, o_code, o_bpp, o_bpc, o_cnt, o_fix, o_acc, o_misc;
};

__device__ __forceinline__ void group_bar(int id) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kTG) : "memory");
}

// exclusive scan over the kTG threads of a group; s_w: 4 words nobody else writes until two barriers later
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
    const uint4 w = *reinterpret_cast<const uint4*>(s_w);
    const uint32_t off = (gw > 0 ? w.x : 0u) + (gw > 1 ? w.y : 0u) + (gw > 2 ? w.z : 0u);
    total = w.x + w.y + w.z + w.w;
    return off + incl - v;
}

template <int IN> struct TieVal {
    using type = float;
    static __device__ __forceinline__ uint32_t bits(const float* p, int i) { return __float_as_uint(__ldg(p + i)); }
};
template <> struct TieVal<IN_CSR16_U8> {
    using type = uint8_t;
    // small integer counts stored as bytes; their float image orders like the counts
    static __device__ __forceinline__ uint32_t bits(const uint8_t* p, int i) {
        return __float_as_uint((float)__ldg(p + i));
    }
};

template <bool UNIT> struct TieRow { using type = uint4; };
template <> struct TieRow<true> { using type = uint2; };

template <bool UNIT>
__device__ __forceinline__ typename TieRow<UNIT>::type tie_load_row(const TieArgs& a, unsigned pos) {
    if constexpr (UNIT) return __ldg(a.ell_u + pos);
    else return __ldg(a.ell_w + pos);
}

__device__ __forceinline__ void tie_slot_w(uint32_t* acc, uint32_t W, uint32_t off, uint32_t m, int L, uint32_t maskL) {
    if (W) {
        const unsigned long long p = (unsigned long long)W * m;
        uint32_t* const at = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(acc) + off);
        atomicAdd(at, (uint32_t)p & maskL);
        atomicAdd(at + 1, (uint32_t)(p >> L));
    }
}

// add W * m to the accumulators of every regulon holding the gene whose table row is `row`
template <bool UNIT>
__device__ __forceinline__ void tie_apply(const TieArgs& a, uint32_t* acc, const typename TieRow<UNIT>::type& row,
                                          uint32_t m, uint32_t maskL) {
    if constexpr (UNIT) {
        const uint32_t a0 = row.x & 0xFFFFu, a1 = row.x >> 16;
        if (a0 != kTiePadU) atomicAdd(acc + a0, m);
        if (a1 != kTiePadU) atomicAdd(acc + a1, m);
        if (row.y >> 31) {
            const uint32_t start = row.y & 0x7FFFFFFFu;
            const uint32_t cnt = __ldg(a.xu + start);
#pragma unroll 1
            for (uint32_t k = 1; k <= cnt; ++k) atomicAdd(acc + __ldg(a.xu + start + k), m);
        } else {
            const uint32_t a2 = row.y & 0xFFFFu, a3 = row.y >> 16;
            if (a2 != kTiePadU) atomicAdd(acc + a2, m);
            if (a3 != kTiePadU) atomicAdd(acc + a3, m);
        }
    } else {
        tie_slot_w(acc, row.x, row.y & 0x7FFFFFFFu, m, a.L, maskL);
        if (row.y >> 31) {
#pragma unroll 1
            for (uint32_t k = 0; k < row.w; ++k) {
                const uint2 s = __ldg(a.xw + row.z + k);
                tie_slot_w(acc, s.x, s.y, m, a.L, maskL);
            }
        } else {
            tie_slot_w(acc, row.z, row.w, m, a.L, maskL);
        }
    }
}

template <int IN, bool UNIT>
__global__ void __launch_bounds__(kTG* kTieMaxGroups, 1) aucell_tie_kernel(const TieArgs a) {
    using IdxT = typename IdxOf<IN>::type;
    using VT = TieVal<IN>;
    using RowT = typename TieRow<UNIT>::type;
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, grp = tid / kTG, gtid = tid % kTG, lane = tid & 31, gw = gtid >> 5;

    // inv_perm -> shared memory, once per CTA
    uint16_t* const ip = reinterpret_cast<uint16_t*>(smem);
    {
        const int nv = (a.n_genes + 7) >> 3;
        const uint4* const src = reinterpret_cast<const uint4*>(a.inv_perm);
        for (int i = tid; i < nv; i += blockDim.x) reinterpret_cast<uint4*>(ip)[i] = __ldg(src + i);
    }
    unsigned char* const gb = smem + a.off_group0 + grp * a.group_stride;
    uint32_t* const A = reinterpret_cast<uint32_t*>(gb + a.o_A);          // [n_words] occupied positions
    uint16_t* const wpre = reinterpret_cast<uint16_t*>(gb + a.o_wpre);    // [n_words] set bits before the word
    uint32_t* const rep = reinterpret_cast<uint32_t*>(gb + a.o_rep);      // [kTieNB] value bits of the bin, 0 = empty
    uint8_t* const code_of = reinterpret_cast<uint8_t*>(gb + a.o_code);   // [kTieNB] class of the bin
    uint16_t* const bpp = reinterpret_cast<uint16_t*>(gb + a.o_bpp);      // [cap] position of the j-th entry
    uint8_t* const bpc = reinterpret_cast<uint8_t*>(gb + a.o_bpc);        // [cap] its class
    uint32_t* const cnt32 = reinterpret_cast<uint32_t*>(gb + a.o_cnt);    // [kTieD][kTG] uint16 counters / bases
    uint16_t* const cnt16 = reinterpret_cast<uint16_t*>(gb + a.o_cnt);
    uint2* const fix = reinterpret_cast<uint2*>(gb + a.o_fix);            // [kTieFix] (value bits, pos) of the tail
    uint32_t* const acc = reinterpret_cast<uint32_t*>(gb + a.o_acc);
    uint32_t* const misc = reinterpret_cast<uint32_t*>(gb + a.o_misc);    // [0..31] reductions (two parities),
                                                                          // [32..35], [36..39] scan words,
                                                                          // [40] tail count, [41] mixed-bin flag
    for (int i = gtid; i < a.acc_words; i += kTG) acc[i] = 0u;
    __syncthreads();

    const int barid = 1 + grp;
    const int n_genes = a.n_genes, n_words = a.n_words, cutoff = a.cutoff, R = a.n_regulons;
    const uint32_t maskL = UNIT ? 0u : ((1u << a.L) - 1u);
    int par = 0;
    int seen = 0, handed = 0;             // cells this group looked at / gave to the to-do list (group-uniform)

    auto hand_over = [&](int64_t cell, int why) {
        if (gtid == 0) {
            const int k = atomicAdd(a.todo_n, 1);
            a.todo[k] = (int32_t)cell;
            if (a.stats != nullptr) atomicAdd(a.stats + why, 1ull);
        }
        ++handed;
    };

    for (int64_t cell = (int64_t)blockIdx.x * a.groups + grp; cell < a.n_cells;
         cell += (int64_t)gridDim.x * a.groups, par ^= 1) {
        // data this kernel does not suit (continuous values): stop looking, hand everything over
        if (seen >= 16 && handed * 8 >= seen * 7) {
            hand_over(cell, TIE_ST_SKIPPED);
            continue;
        }
        ++seen;
        const int64_t rb = __ldg(a.indptr + cell);
        const int64_t n64 = __ldg(a.indptr + cell + 1) - rb;
        const bool fits = n64 >= 0 && n64 <= (int64_t)a.cap;
        const int n = fits ? (int)n64 : 0;
        const IdxT* const cols = reinterpret_cast<const IdxT*>(a.indices) + rb;
        const typename VT::type* const vals = reinterpret_cast<const typename VT::type*>(a.data) + rb;

        // ---------------- P0: statistics; clear the cell's working set -------------------------------------------
        uint32_t umin = 0xFFFFFFFFu, umax = 0u, bad = 0u;
        int npos = 0;
#pragma unroll 4
        for (int i = gtid; i < n; i += kTG) {
            const uint32_t u = VT::bits(vals, i);
            const bool nz = (u << 1) != 0u;                // +0.0 / -0.0 are not entries
            const bool ng = u > 0x7F800000u;               // negative or NaN
            const bool ps = nz && !ng;
            bad |= (nz && ng) ? 1u : 0u;
            npos += ps ? 1 : 0;
            umin = min(umin, ps ? u : 0xFFFFFFFFu);
            umax = max(umax, ps ? u : 0u);
        }
        umin = __reduce_min_sync(0xffffffffu, umin);
        umax = __reduce_max_sync(0xffffffffu, umax);
        npos = __reduce_add_sync(0xffffffffu, npos);
        bad = __reduce_or_sync(0xffffffffu, bad);
        uint32_t* const red = misc + par * 16;
        if (lane == 0) *reinterpret_cast<uint4*>(red + gw * 4) = make_uint4(umin, umax, (uint32_t)npos, bad);
        {
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
            for (int i = gtid; i < ((n_words + 3) >> 2); i += kTG) reinterpret_cast<uint4*>(A)[i] = z;
#pragma unroll
            for (int i = gtid; i < kTieNB / 4; i += kTG) reinterpret_cast<uint4*>(rep)[i] = z;
#pragma unroll
            for (int i = gtid; i < kTieD * kTG / 8; i += kTG) reinterpret_cast<uint4*>(cnt32)[i] = z;
            if (gtid == 0) { misc[40] = 0u; misc[41] = 0u; }
        }
        group_bar(barid);
        {
            const uint4 r0 = *reinterpret_cast<const uint4*>(red), r1 = *reinterpret_cast<const uint4*>(red + 4),
                        r2 = *reinterpret_cast<const uint4*>(red + 8), r3 = *reinterpret_cast<const uint4*>(red + 12);
            umin = min(min(r0.x, r1.x), min(r2.x, r3.x));
            umax = max(max(r0.y, r1.y), max(r2.y, r3.y));
            npos = (int)(r0.z + r1.z + r2.z + r3.z);
            bad = r0.w | r1.w | r2.w | r3.w;
        }
        if (!fits || bad) {
            hand_over(cell, fits ? TIE_ST_BADVAL : TIE_ST_LONG);
            continue;
        }
        const uint32_t range = umax - umin;               // npos == 0: unused
        const int shift = max(0, (32 - __clz((int)range)) - kTieNBLog);

        // ---------------- P1: position bitmap, one representative value per bin ----------------------------------
#pragma unroll 4
        for (int i = gtid; i < n; i += kTG) {
            const uint32_t u = VT::bits(vals, i);
            const unsigned col = (unsigned)cols[i];
            if ((u << 1) != 0u && col < (unsigned)n_genes) {
                const unsigned pos = ip[col];
                rep[(umax - u) >> shift] = u;
                atomicOr(&A[pos >> 5], 1u << (pos & 31));
            }
        }
        group_bar(barid);

        // ---------------- P2: word prefix of the bitmap; occupied bins -> classes ---------------------------------
        int d_total;
        {
            const int w0 = gtid * a.wpt, w1 = min(w0 + a.wpt, n_words);
            uint32_t local = 0;
#pragma unroll 1
            for (int w = w0; w < w1; ++w) local += __popc(A[w]);
            static_assert(kTieNB / kTG == 8, "eight bins per thread");
            const uint4 q0 = reinterpret_cast<const uint4*>(rep)[gtid * 2], q1 = reinterpret_cast<const uint4*>(rep)[gtid * 2 + 1];
            const uint32_t occ = (q0.x ? 1u : 0u) | (q0.y ? 2u : 0u) | (q0.z ? 4u : 0u) | (q0.w ? 8u : 0u) |
                                 (q1.x ? 16u : 0u) | (q1.y ? 32u : 0u) | (q1.z ? 64u : 0u) | (q1.w ? 128u : 0u);
            uint32_t total;
            const uint32_t excl = group_excl_scan((local << 16) | (uint32_t)__popc(occ), misc + 32, lane, gw, barid, total);
            d_total = (int)(total & 0xFFFFu);
            const int n_bits = (int)(total >> 16);
            uint32_t run = excl >> 16;
#pragma unroll 1
            for (int w = w0; w < w1; ++w) {
                wpre[w] = (uint16_t)run;
                run += __popc(A[w]);
            }
            // class of my bins: index among the occupied bins (descending value) minus the tail's share
            const int excess = max(0, d_total - (kTieD - 1));
            int code = (int)(excl & 0xFFFFu);
            unsigned long long packed = 0ull;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if ((occ >> k) & 1u) {
                    const int cls = code < excess ? 0 : code - excess + 1;
                    packed |= (unsigned long long)cls << (8 * k);
                    ++code;
                }
            }
            reinterpret_cast<unsigned long long*>(code_of)[gtid] = packed;
            group_bar(barid);
            if (n_bits != npos) {                          // a column stored twice, or a column index out of range
                hand_over(cell, TIE_ST_DUP);
                continue;
            }
        }

        // ---------------- P3: entries into position order; class counts per chunk; the tail list ------------------
        const int ch = max(1, (npos + kTG - 1) / kTG);     // items per chunk (thread) in P6
        const uint32_t magic = ((1u << 20) + (uint32_t)ch - 1u) / (uint32_t)ch;
#pragma unroll 4
        for (int i = gtid; i < n; i += kTG) {
            const uint32_t u = VT::bits(vals, i);
            const unsigned col = (unsigned)cols[i];
            if ((u << 1) != 0u) {
                const unsigned pos = ip[col];
                const unsigned w = pos >> 5;
                const uint32_t j = (uint32_t)wpre[w] + (uint32_t)__popc(A[w] & ((1u << (pos & 31)) - 1u));
                const uint32_t bin = (umax - u) >> shift;
                const uint32_t cls = code_of[bin];
                bpp[j] = (uint16_t)pos;
                bpc[j] = (uint8_t)cls;
                const uint32_t chunk = (j * magic) >> 20;
                atomicAdd(&cnt32[(cls * kTG + chunk) >> 1], 1u << ((chunk & 1u) << 4));
                if (cls == 0u) {
                    const uint32_t f = atomicAdd(&misc[40], 1u);
                    if (f < (uint32_t)kTieFix) fix[f] = make_uint2(u, pos);
                } else if (rep[bin] != u) {
                    misc[41] = 1u;                         // two different values in one class bin
                }
            }
        }
        group_bar(barid);
        const int n_tail = (int)misc[40];
        if (misc[41] != 0u || n_tail > kTieFix) {
            hand_over(cell, misc[41] != 0u ? TIE_ST_MIXED : TIE_ST_TAIL);
            group_bar(barid);                              // the flags are cleared by the next cell: read them first
            continue;
        }

        // ---------------- P5: exclusive scan of the counters in (class, chunk) order ------------------------------
        {
            static_assert(kTieD * kTG / kTG == 32, "thread t scans the 32 counters [32 t, 32 t + 32) of class t / 4");
            const int n_cls = min(d_total, kTieD - 1) + 1;
            const bool live = (gtid >> 2) < n_cls;
            uint4 c[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) c[k] = live ? reinterpret_cast<const uint4*>(cnt32)[gtid * 4 + k] : make_uint4(0u, 0u, 0u, 0u);
            uint32_t s = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) s += c[k].x + c[k].y + c[k].z + c[k].w;       // both halves stay below 2^16
            uint32_t total;
            uint32_t run = group_excl_scan((s & 0xFFFFu) + (s >> 16), misc + 36, lane, gw, barid, total);
            if (live) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint32_t* const wv = reinterpret_cast<uint32_t*>(&c[k]);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t lo = wv[q] & 0xFFFFu, hi = wv[q] >> 16;
                        wv[q] = run | ((run + lo) << 16);
                        run += lo + hi;
                    }
                    reinterpret_cast<uint4*>(cnt32)[gtid * 4 + k] = c[k];
                }
            }
        }
        group_bar(barid);

        // ---------------- P6: rank in position order with private counters, scatter -------------------------------
        {
            int j = gtid * ch;
            const int j1 = min(j + ch, npos);
            if (j < j1) {
                unsigned pos_n = bpp[j], cc_n = bpc[j];
                RowT row_n = tie_load_row<UNIT>(a, pos_n);
#pragma unroll 1
                for (; j < j1; ++j) {
                    const unsigned cc = cc_n;
                    const RowT row = row_n;
                    if (j + 1 < j1) {
                        pos_n = bpp[j + 1];
                        cc_n = bpc[j + 1];
                        row_n = tie_load_row<UNIT>(a, pos_n);
                    }
                    uint16_t* const ctr = cnt16 + cc * kTG + gtid;
                    const int r = (int)*ctr;
                    *ctr = (uint16_t)(r + 1);
                    if (cc != 0u && r < cutoff) tie_apply<UNIT>(a, acc, row, (uint32_t)(cutoff - r), maskL);
                }
            }
        }
        // ---------------- P7: the tail (largest values): exact order by (value, position) -------------------------
#pragma unroll 1
        for (int f = gtid; f < n_tail; f += kTG) {
            const uint2 mine = fix[f];
            int r = 0;
#pragma unroll 4
            for (int g = 0; g < n_tail; ++g) {
                const uint2 o = fix[g];
                r += (o.x > mine.x || (o.x == mine.x && o.y < mine.y)) ? 1 : 0;
            }
            if (r < cutoff) tie_apply<UNIT>(a, acc, tie_load_row<UNIT>(a, mine.y), (uint32_t)(cutoff - r), maskL);
        }
        // ---------------- P8: zeros fill the ranks npos .. cutoff-1 in position order ------------------------------
        const int q_zero = cutoff - npos;
        if (q_zero > 0) {
#pragma unroll 1
            for (int p = gtid; p < n_genes; p += kTG) {
                const int w = p >> 5;
                const int z0 = (w << 5) - (int)wpre[w];
                if (z0 >= q_zero) break;                   // words are visited in increasing order by every thread
                const uint32_t bw = A[w];
                if (!((bw >> (p & 31)) & 1u)) {
                    const int z = z0 + (p & 31) - __popc(bw & ((1u << (p & 31)) - 1u));
                    if (z < q_zero) tie_apply<UNIT>(a, acc, tie_load_row<UNIT>(a, (unsigned)p), (uint32_t)(q_zero - z), maskL);
                }
            }
        }
        if (gtid == 0) {
            if (a.out_nnz != nullptr) a.out_nnz[cell] = npos;
            if (a.stats != nullptr) atomicAdd(a.stats + TIE_ST_FAST, 1ull);
        }
        group_bar(barid);

        // ---------------- P9: AUCs out, accumulators cleared -------------------------------------------------------
        double* const orow = a.out_auc + cell * (int64_t)R;
#pragma unroll 1
        for (int r = gtid; r < R; r += kTG) {
            const int f = __ldg(a.acc_first + r);
            double res = 0.0;
            if (f >= 0) {
                double sum;
                if constexpr (UNIT) {
                    sum = (double)acc[f];
                    acc[f] = 0u;
                } else {
                    const int parts = __ldg(a.acc_parts + r);
                    sum = 0.0;
#pragma unroll 1
                    for (int k = 0; k < parts; ++k) {
                        const uint2 lh = *reinterpret_cast<const uint2*>(acc + 2 * (f + k));
                        const long long v = (long long)(((unsigned long long)lh.y << a.L) + (unsigned long long)lh.x);
                        sum += (double)v * __ldg(a.acc_scale + f + k);
                        *reinterpret_cast<uint2*>(acc + 2 * (f + k)) = make_uint2(0u, 0u);
                    }
                }
                res = __ddiv_rn(sum, __ldg(a.max_auc + r));
            }
            orow[r] = res;
        }
        // no barrier: the next cell passes several group barriers before its first scatter
    }
}

}  // namespace aucell

// common.cuh -- small device helpers shared by the AUCell kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "keys.cuh"

namespace aucell {

// IN_CSR16: CSR with uint16 column indices (n_genes <= 65535): 6 instead of 8 bytes per stored entry over PCIe / HBM
// IN_CSR16_U8: uint16 column indices and uint8 values (integer counts <= 255): 3 bytes per stored entry
// IN_CSR_U8: int32 column indices and uint8 values: 5 bytes per stored entry (the host only has to touch the values)
enum InputKind { IN_DENSE = 0, IN_CSR = 1, IN_CSR16 = 2, IN_CSR16_U8 = 3, IN_CSR_U8 = 4 };

template <int IN> struct IdxOf { using type = int32_t; };
template <> struct IdxOf<IN_CSR16> { using type = uint16_t; };
template <> struct IdxOf<IN_CSR16_U8> { using type = uint16_t; };

template <typename T> __device__ __forceinline__ T tmin(T a, T b) { return a < b ? a : b; }
template <typename T> __device__ __forceinline__ T tmax(T a, T b) { return a > b ? a : b; }

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// `sum`, or 1.0 when sum is +0.0 -- through the bit pattern, so that the compiler cannot fold the substitution back
// into the division it guards (a zero dividend sends DDIV down its out-of-line special-case path for the whole warp;
// the caller selects 0.0 for those lanes afterwards).  Sums here are never negative zero.
__device__ __forceinline__ double nonzero_numerator(double sum) {
    const long long b = __double_as_longlong(sum);
    return __longlong_as_double(b | (b == 0ll ? 0x3FF0000000000000ll : 0ll));
}

}  // namespace aucell

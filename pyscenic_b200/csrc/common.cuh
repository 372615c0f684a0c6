// common.cuh -- small device helpers shared by the AUCell kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "keys.cuh"

namespace aucell {

enum InputKind { IN_DENSE = 0, IN_CSR = 1 };

template <typename T> __device__ __forceinline__ T tmin(T a, T b) { return a < b ? a : b; }
template <typename T> __device__ __forceinline__ T tmax(T a, T b) { return a > b ? a : b; }

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

}  // namespace aucell

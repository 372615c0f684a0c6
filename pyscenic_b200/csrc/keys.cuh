// keys.cuh -- order-preserving integer keys for expression values.
//
// The reference ranks every cell with pandas
//     rank(axis=1, ascending=False, method="first", na_option="bottom")      (src/pyscenic/aucell.py:48)
// i.e. a stable DESCENDING sort with NaN after every number and -0.0 tied with +0.0.  We map every value to
// an unsigned key whose ASCENDING integer order is that order, so that
//     rank = #{smaller key} + #{equal key at a smaller shuffled position}.
//
//   value        float32 key (u32)            float64 key (u64)
//   +inf         0x007FFFFF                   0x000FFFFFFFFFFFFF
//   x > 0        0x7FFFFFFF - bits(x)         0x7FFF...F - bits(x)
//   +0.0, -0.0   0x7FFFFFFF  (KEY_ZERO)       0x7FFFFFFFFFFFFFFF
//   x < 0        bits(x)  (sign bit set)      bits(x)
//   -inf         0xFF800000                   0xFFF0000000000000
//   NaN          0xFFFFFFFF  (KEY_NAN)        0xFFFFFFFFFFFFFFFF
//
// Only bit operations are used, so denormals are kept apart exactly as the CPU does (no flush-to-zero).
#pragma once
#include <stdint.h>

namespace aucell {

template <typename ValT>
struct KeyOf;

template <>
struct KeyOf<float> {
    using type = uint32_t;
    static constexpr uint32_t KEY_ZERO = 0x7FFFFFFFu;
    static constexpr uint32_t KEY_NAN = 0xFFFFFFFFu;
    static __host__ __device__ __forceinline__ uint32_t make(float v) {
#ifdef __CUDA_ARCH__
        uint32_t u = __float_as_uint(v);
#else
        union { float f; uint32_t u; } cvt; cvt.f = v; uint32_t u = cvt.u;
#endif
        const uint32_t mag = u & 0x7FFFFFFFu;
        uint32_t key = (u >> 31) ? u : (0x7FFFFFFFu - u);      // selects, no branches: this runs once per stored value
        key = (mag == 0u) ? KEY_ZERO : key;
        key = (mag > 0x7F800000u) ? KEY_NAN : key;
        return key;
    }
};

template <>
struct KeyOf<double> {
    using type = uint64_t;
    static constexpr uint64_t KEY_ZERO = 0x7FFFFFFFFFFFFFFFull;
    static constexpr uint64_t KEY_NAN = 0xFFFFFFFFFFFFFFFFull;
    static __host__ __device__ __forceinline__ uint64_t make(double v) {
#ifdef __CUDA_ARCH__
        uint64_t u = (uint64_t)__double_as_longlong(v);
#else
        union { double f; uint64_t u; } cvt; cvt.f = v; uint64_t u = cvt.u;
#endif
        const uint64_t mag = u & 0x7FFFFFFFFFFFFFFFull;
        uint64_t key = (u >> 63) ? u : (0x7FFFFFFFFFFFFFFFull - u);
        key = (mag == 0ull) ? KEY_ZERO : key;
        key = (mag > 0x7FF0000000000000ull) ? KEY_NAN : key;
        return key;
    }
};

}  // namespace aucell

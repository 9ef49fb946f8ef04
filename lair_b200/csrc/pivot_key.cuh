// Pivot-search helpers shared by the batched and panel kernels.
// blas::iamax (src/blas/iamax.rs:6-21): first index of max |x|, strict `>`, NaN never wins.
// |x| is mapped to a monotone unsigned key (the IEEE bit pattern of a non-negative float
// orders like an unsigned integer); zero and NaN map to key 0, so "nothing exceeded 0"
// falls out as kmax == 0, and ties are broken towards the lowest logical row.
#pragma once
#include <cstdint>

namespace lair {

constexpr unsigned kFullMask = 0xffffffffu;

template <class T> struct PivotKey;
template <> struct PivotKey<float> {
    using type = uint32_t;
    __device__ static __forceinline__ type of(float x) {
        float a = fabsf(x);
        return (a > 0.f) ? __float_as_uint(a) : 0u;
    }
    __device__ static __forceinline__ type warp_max(type k) { return __reduce_max_sync(kFullMask, k); }
};
template <> struct PivotKey<double> {
    using type = unsigned long long;
    __device__ static __forceinline__ type of(double x) {
        double a = fabs(x);
        return (a > 0.0) ? (type)__double_as_longlong(a) : 0ull;
    }
    __device__ static __forceinline__ type warp_max(type k) {
        uint32_t hi = (uint32_t)(k >> 32);
        uint32_t mh = __reduce_max_sync(kFullMask, hi);
        uint32_t lo = (hi == mh) ? (uint32_t)k : 0u;
        uint32_t ml = __reduce_max_sync(kFullMask, lo);
        return ((type)mh << 32) | ml;
    }
};

}  // namespace lair

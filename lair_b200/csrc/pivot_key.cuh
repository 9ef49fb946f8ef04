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

// Exact arg-max of (key, pos) over a full warp: larger key wins, ties -> smaller pos.  A coarse
// pass on the top 32 bits of the key (one REDUX + one vote) decides almost every call; the full
// comparison (low word, then position) runs only among the lanes that tie on the coarse key.
// Returns the winning lane in `src` and its key / position in kbest / pbest (all lanes get them).
template <class KT>
__device__ __forceinline__ void warp_argmax(KT key, unsigned pos, KT& kbest, unsigned& pbest, int& src) {
    const unsigned lane_bit = 1u << (threadIdx.x & 31);
    const uint32_t hi = (sizeof(KT) == 8) ? (uint32_t)((unsigned long long)key >> 32) : (uint32_t)key;
    const uint32_t mh = __reduce_max_sync(kFullMask, hi);
    unsigned tie = __ballot_sync(kFullMask, hi == mh);
    if (__popc(tie) != 1) {
        if (sizeof(KT) == 8) {
            const uint32_t lo = (hi == mh) ? (uint32_t)key : 0u;
            const uint32_t ml = __reduce_max_sync(kFullMask, lo);
            tie = __ballot_sync(kFullMask, (hi == mh) && ((uint32_t)key == ml));
        }
        if (__popc(tie) != 1) {
            const unsigned pm = __reduce_min_sync(kFullMask, (tie & lane_bit) ? pos : 0xffffffffu);
            tie = __ballot_sync(kFullMask, ((tie & lane_bit) != 0) && pos == pm);
        }
    }
    src = __ffs(tie) - 1;
    kbest = (KT)__shfl_sync(kFullMask, (unsigned long long)key, src);
    pbest = __shfl_sync(kFullMask, pos, src);
}

}  // namespace lair

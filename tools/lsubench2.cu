// The batched-LU inner pattern in isolation: per pair 2 SHFL (winner's packed pair) -> FFMA2 -> FADD2 on the lane's own pair.
// Variants: pairs updated in place (as the kernel does), batch size of the shuffles, scalar instead of packed arithmetic.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
using u64 = unsigned long long;
__device__ __forceinline__ u64 pack32(unsigned lo, unsigned hi) { return ((u64)hi << 32) | lo; }
__device__ __forceinline__ u64 shfl64(u64 v, int src) {
    return pack32(__shfl_sync(0xffffffffu, (unsigned)v, src), __shfl_sync(0xffffffffu, (unsigned)(v >> 32), src));
}
template <int BATCH, int MODE>
__global__ void __launch_bounds__(32) k(float* out, long long* cyc, int iters, int zero, u64 nz) {
    u64 a[16];
    for (int i = 0; i < 16; ++i) a[i] = pack32(__float_as_uint(1.0f + threadIdx.x * 1e-3f + i), __float_as_uint(2.0f + i));
    int wl = (threadIdx.x * 7 + zero) & 31;
    u64 ll = pack32(__float_as_uint(1e-6f), __float_as_uint(1e-6f));
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        wl = (wl + 1 + zero) & 31;
        if (MODE == 0) {
#pragma unroll
            for (int p = 0; p < 16; p += BATCH) {
                u64 u[BATCH];
#pragma unroll
                for (int i = 0; i < BATCH; ++i) u[i] = shfl64(a[p + i], wl);
#pragma unroll
                for (int i = 0; i < BATCH; ++i) {
                    u64 t;
                    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(t) : "l"(ll), "l"(u[i]), "l"(nz));
                    asm("sub.rn.f32x2 %0, %0, %1;" : "+l"(a[p + i]) : "l"(t));
                }
            }
        } else if (MODE == 1) {  // shuffles only (results folded by xor)
#pragma unroll
            for (int p = 0; p < 16; ++p) a[p] ^= shfl64(a[p], wl) & 1ull;
        } else if (MODE == 2) {  // arithmetic only
#pragma unroll
            for (int p = 0; p < 16; ++p) {
                u64 t;
                asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(t) : "l"(ll), "l"(a[(p + 1) & 15]), "l"(nz));
                asm("sub.rn.f32x2 %0, %0, %1;" : "+l"(a[p]) : "l"(t));
            }
        } else {  // scalar arithmetic: 2 SHFL + 2 FMUL + 2 FADD per pair
#pragma unroll
            for (int p = 0; p < 16; p += BATCH) {
                u64 u[BATCH];
#pragma unroll
                for (int i = 0; i < BATCH; ++i) u[i] = shfl64(a[p + i], wl);
#pragma unroll
                for (int i = 0; i < BATCH; ++i) {
                    float l = __uint_as_float((unsigned)ll);
                    float x = __fsub_rn(__uint_as_float((unsigned)a[p + i]), __fmul_rn(l, __uint_as_float((unsigned)u[i])));
                    float y = __fsub_rn(__uint_as_float((unsigned)(a[p + i] >> 32)), __fmul_rn(l, __uint_as_float((unsigned)(u[i] >> 32))));
                    a[p + i] = pack32(__float_as_uint(x), __float_as_uint(y));
                }
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 16; ++i) s += __uint_as_float((unsigned)a[i]) + __uint_as_float((unsigned)(a[i] >> 32));
    out[blockIdx.x * 32 + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int BATCH, int MODE>
void run(const char* name, int ctas_per_sm, float* out, long long* cyc, long long* h) {
    const int iters = 4000, grid = 148 * ctas_per_sm;
    k<BATCH, MODE><<<grid, 32>>>(out, cyc, 10, 0, 0x8000000080000000ull);
    cudaDeviceSynchronize();
    k<BATCH, MODE><<<grid, 32>>>(out, cyc, iters, 0, 0x8000000080000000ull);
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("{\"pattern\": \"%s\", \"batch\": %d, \"warps_per_sm\": %d, \"cycles_per_pair_per_sm\": %.3f, \"err\": \"%s\"}\n", name, BATCH, ctas_per_sm,
           (double)mx / ((double)iters * 16 * ctas_per_sm), cudaGetErrorString(cudaGetLastError()));
}
int main() {
    float* out; long long *cyc, *h;
    cudaMalloc(&out, 148 * 32 * 32 * 4); cudaMalloc(&cyc, 148 * 32 * 8); h = (long long*)malloc(148 * 32 * 8);
    for (int w : {8, 16, 24, 32}) {
        run<1, 0>("2shfl+ffma2+fadd2", w, out, cyc, h);
        run<4, 0>("2shfl+ffma2+fadd2", w, out, cyc, h);
        run<16, 0>("2shfl+ffma2+fadd2", w, out, cyc, h);
        run<1, 1>("2shfl only", w, out, cyc, h);
        run<1, 2>("ffma2+fadd2 only", w, out, cyc, h);
        run<4, 3>("2shfl+2fmul+2fadd scalar", w, out, cyc, h);
    }
    return 0;
}

// LSU / MIO issue-rate microbenchmark: how many warp instructions per clock one SM retires for the
// shared-memory, shuffle, vote and reduction instructions the batched 32x32 LU kernel is built from.
// The batched kernel is bound by shared-memory INSTRUCTIONS (profiles/r1b_batched_v4.md); these
// numbers say what each one costs, which is what the mapping of batched_lu5.cu was chosen by.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lsubench lsubench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

using u64 = unsigned long long;

enum Op { LDS128_BCAST, LDS128_2ADDR, LDS128_FULL, LDS64_16ADDR, LDS32_BCAST, STS128_1LANE, STS128_2LANE, STS128_FULL,
          STS128_PREDOFF, STS64_16LANE, REDUX_MAX, VOTE_BALLOT, SHFL_IDX, LDS_BCAST_PLUS_FFMA, LDS_BCAST_PLUS_DFMA, PAIR_STS1_LDSB,
          FFMA2_ONLY, DFMA_ONLY, STS32_1LANE, STS64_1LANE, STS32_FULL, LDS64_BCAST, SHFL_INDEP, REDUX_INDEP, SEL_ONLY, SHFL_REGLANE, SHFL_VARLANE, SHFL_REG_PLUS_FFMA2, NOPS };
static const char* kNames[] = {"lds128_bcast", "lds128_2addr", "lds128_full(4wf)", "lds64_16addr", "lds32_bcast", "sts128_1lane", "sts128_2lane",
                               "sts128_full(4wf)", "sts128_pred_off", "sts64_16lane", "redux_max_u32", "vote_ballot", "shfl_idx",
                               "lds128_bcast+4ffma2", "lds128_bcast+2dfma", "sts128_1lane+lds128_bcast", "ffma2_only", "dfma_only", "sts32_1lane", "sts64_1lane", "sts32_full", "lds64_bcast", "shfl_idx_indep", "redux_indep", "sel_only", "shfl_idx_reg_uniform_lane", "shfl_idx_per_lane_src", "shfl_reg_lane+2ffma2"};

template <int OP>
__global__ void __launch_bounds__(256) bench(unsigned* out, long long* cyc, int iters, int zero) {
    __shared__ __align__(16) unsigned char buf[16384];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned base = (unsigned)__cvta_generic_to_shared(buf) + warp * 2048;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) reinterpret_cast<unsigned*>(buf)[i] = i;
    __syncthreads();
    unsigned addr_b = base;                          // broadcast
    unsigned addr_2 = base + (lane >> 4) * 16;       // two distinct 16-byte addresses
    unsigned addr_f = base + lane * 16;              // conflict-free full warp
    unsigned addr_16 = base + (lane & 15) * 8;       // 16 distinct 8-byte addresses
    int p1 = (lane == (zero + 5)), p2 = ((lane & 15) == (zero + 5)), poff = (zero != 0), p16 = (lane < 16 + zero);
    u64 x = lane, y = lane + 1;
    unsigned acc = lane;
    u64 f0 = 0x3f8000003f800000ull, f1 = f0, f2 = f0, f3 = f0, fm = 0x3f8000013f800001ull;
    double d0 = 1.0, d1 = 1.0, dm = 1.0000001;
    long long t0 = clock64();
    const unsigned a_b0 = addr_b, a_20 = addr_2, a_f0 = addr_f, a_160 = addr_16;
    for (int it = 0; it < iters; ++it) {
        const unsigned off = (unsigned)it * (unsigned)zero;  // runtime 0: keeps the loads inside the loop
        addr_b = a_b0 + off; addr_2 = a_20 + off; addr_f = a_f0 + off; addr_16 = a_160 + off;
#pragma unroll
        for (int u = 0; u < 32; ++u) {
            if (OP == LDS128_BCAST || OP == LDS_BCAST_PLUS_FFMA || OP == LDS_BCAST_PLUS_DFMA || OP == PAIR_STS1_LDSB) {
                u64 a, b;
                asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr_b + u * 16) : "memory");
                acc ^= (unsigned)a ^ (unsigned)(a >> 32) ^ (unsigned)b ^ (unsigned)(b >> 32);
            }
            if (OP == LDS128_2ADDR) {
                u64 a, b;
                asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr_2 + u * 32) : "memory");
                acc ^= (unsigned)a ^ (unsigned)(a >> 32) ^ (unsigned)b ^ (unsigned)(b >> 32);
            }
            if (OP == LDS128_FULL) {
                u64 a, b;
                asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr_f + (u & 3) * 512) : "memory");
                acc ^= (unsigned)a ^ (unsigned)(a >> 32) ^ (unsigned)b ^ (unsigned)(b >> 32);
            }
            if (OP == LDS64_16ADDR) {
                u64 a;
                asm volatile("ld.shared.b64 %0, [%1];" : "=l"(a) : "r"(addr_16 + (u & 15) * 128) : "memory");
                acc ^= (unsigned)a ^ (unsigned)(a >> 32);
            }
            if (OP == LDS32_BCAST) {
                unsigned a;
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(a) : "r"(addr_b + u * 4) : "memory");
                acc ^= a;
            }
            if (OP == STS128_1LANE || OP == PAIR_STS1_LDSB)
                asm volatile("{\n .reg .pred p;\n setp.ne.s32 p, %3, 0;\n @p st.shared.v2.b64 [%0], {%1, %2};\n}" ::"r"(addr_b + (u & 7) * 16 + 1024), "l"(x), "l"(y), "r"(p1) : "memory");
            if (OP == STS128_2LANE)
                asm volatile("{\n .reg .pred p;\n setp.ne.s32 p, %3, 0;\n @p st.shared.v2.b64 [%0], {%1, %2};\n}" ::"r"(addr_2 + (u & 7) * 32 + 1024), "l"(x), "l"(y), "r"(p2) : "memory");
            if (OP == STS128_FULL)
                asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(addr_f + (u & 3) * 512), "l"(x), "l"(y) : "memory");
            if (OP == STS128_PREDOFF)
                asm volatile("{\n .reg .pred p;\n setp.ne.s32 p, %3, 0;\n @p st.shared.v2.b64 [%0], {%1, %2};\n}" ::"r"(addr_f + (u & 3) * 512), "l"(x), "l"(y), "r"(poff) : "memory");
            if (OP == STS64_16LANE)
                asm volatile("{\n .reg .pred p;\n setp.ne.s32 p, %2, 0;\n @p st.shared.b64 [%0], %1;\n}" ::"r"(addr_16 + (u & 7) * 128), "l"(x), "r"(p16) : "memory");
            if (OP == REDUX_MAX) {
                unsigned r;
                asm volatile("redux.sync.max.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(acc + u));
                acc = r ^ lane;
            }
            if (OP == VOTE_BALLOT) {
                unsigned r;
                asm volatile("{\n .reg .pred p;\n setp.gt.u32 p, %1, %2;\n vote.sync.ballot.b32 %0, p, 0xffffffff;\n}" : "=r"(r) : "r"(acc), "r"(u * 977u));
                acc += r;
            }
            if (OP == SHFL_IDX) {
                unsigned r;
                asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(r) : "r"(acc), "r"(u & 31));
                acc = r + lane;
            }
            if (OP == STS32_1LANE)
                asm volatile("{\n .reg .pred p;\n setp.ne.s32 p, %2, 0;\n @p st.shared.b32 [%0], %1;\n}" ::"r"(addr_b + u * 4 + 1024), "r"(acc), "r"(p1) : "memory");
            if (OP == STS64_1LANE)
                asm volatile("{\n .reg .pred p;\n setp.ne.s32 p, %2, 0;\n @p st.shared.b64 [%0], %1;\n}" ::"r"(addr_b + u * 8 + 1024), "l"(x), "r"(p1) : "memory");
            if (OP == STS32_FULL)
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + lane * 4 + (u & 7) * 128), "r"(acc) : "memory");
            if (OP == LDS64_BCAST) {
                u64 a;
                asm volatile("ld.shared.b64 %0, [%1];" : "=l"(a) : "r"(addr_b + u * 8) : "memory");
                acc ^= (unsigned)a ^ (unsigned)(a >> 32);
            }
            if (OP == SHFL_INDEP) {
                unsigned r;
                asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(r) : "r"((unsigned)x + u), "r"(u & 31));
                acc ^= r;
            }
            if (OP == SHFL_REGLANE || OP == SHFL_REG_PLUS_FFMA2) {
                unsigned r;
                asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(r) : "r"((unsigned)x + u), "r"(zero + 7));
                acc ^= r;
            }
            if (OP == SHFL_VARLANE) {
                unsigned r;
                asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(r) : "r"((unsigned)x + u), "r"((lane * 5 + zero) & 31));
                acc ^= r;
            }
            if (OP == SHFL_REG_PLUS_FFMA2) {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(f0) : "l"(fm));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(f1) : "l"(fm));
            }
            if (OP == REDUX_INDEP) {
                unsigned r;
                asm volatile("redux.sync.max.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"((unsigned)x + u));
                acc ^= r;
            }
            if (OP == SEL_ONLY) {
                asm volatile("selp.b32 %0, %1, %0, %2;" : "+r"(acc) : "r"((unsigned)x + u), "r"(p1));
            }
            if (OP == LDS_BCAST_PLUS_FFMA || OP == FFMA2_ONLY) {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(f0) : "l"(fm));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(f1) : "l"(fm));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(f2) : "l"(fm));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(f3) : "l"(fm));
            }
            if (OP == LDS_BCAST_PLUS_DFMA || OP == DFMA_ONLY) {
                asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d0) : "d"(dm));
                asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d1) : "d"(dm));
            }
        }
    }
    long long t1 = clock64();
    acc ^= (unsigned)f0 ^ (unsigned)f1 ^ (unsigned)f2 ^ (unsigned)f3 ^ (unsigned)__double_as_longlong(d0 + d1);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(int warps_per_cta, int ctas_per_sm, unsigned* out, long long* cyc, long long* hcyc) {
    int dev, sms;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int iters = 2000, grid = sms * ctas_per_sm;
    bench<OP><<<grid, warps_per_cta * 32>>>(out, cyc, 10, 0);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<OP><<<grid, warps_per_cta * 32>>>(out, cyc, iters, 0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(hcyc, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < grid; ++i) mx = hcyc[i] > mx ? hcyc[i] : mx;
    // "op groups" per SM: one group = the body of one unrolled slot (1 instruction for the single-op cases)
    const double groups = (double)iters * 32 * warps_per_cta * ctas_per_sm;
    printf("{\"op\": \"%s\", \"warps_per_sm\": %d, \"cycles\": %lld, \"ms\": %.4f, \"cycles_per_group_per_sm\": %.3f, \"err\": \"%s\"}\n", kNames[OP],
           warps_per_cta * ctas_per_sm, mx, ms, mx / groups, cudaGetErrorString(cudaGetLastError()));
}

template <int OP>
void sweep(unsigned* out, long long* cyc, long long* hcyc) {
    run<OP>(4, 1, out, cyc, hcyc);
    run<OP>(8, 2, out, cyc, hcyc);
    run<OP>(8, 4, out, cyc, hcyc);
}

int main() {
    unsigned* out;
    long long *cyc, *hcyc;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(unsigned));
    cudaMalloc(&cyc, 148 * 8 * sizeof(long long));
    hcyc = (long long*)malloc(148 * 8 * sizeof(long long));
    sweep<LDS128_BCAST>(out, cyc, hcyc);
    sweep<LDS128_2ADDR>(out, cyc, hcyc);
    sweep<LDS128_FULL>(out, cyc, hcyc);
    sweep<LDS64_16ADDR>(out, cyc, hcyc);
    sweep<LDS32_BCAST>(out, cyc, hcyc);
    sweep<STS128_1LANE>(out, cyc, hcyc);
    sweep<STS128_2LANE>(out, cyc, hcyc);
    sweep<STS128_FULL>(out, cyc, hcyc);
    sweep<STS128_PREDOFF>(out, cyc, hcyc);
    sweep<STS64_16LANE>(out, cyc, hcyc);
    sweep<REDUX_MAX>(out, cyc, hcyc);
    sweep<VOTE_BALLOT>(out, cyc, hcyc);
    sweep<SHFL_IDX>(out, cyc, hcyc);
    sweep<LDS_BCAST_PLUS_FFMA>(out, cyc, hcyc);
    sweep<LDS_BCAST_PLUS_DFMA>(out, cyc, hcyc);
    sweep<PAIR_STS1_LDSB>(out, cyc, hcyc);
    sweep<FFMA2_ONLY>(out, cyc, hcyc);
    sweep<DFMA_ONLY>(out, cyc, hcyc);
    sweep<STS32_1LANE>(out, cyc, hcyc);
    sweep<STS64_1LANE>(out, cyc, hcyc);
    sweep<STS32_FULL>(out, cyc, hcyc);
    sweep<LDS64_BCAST>(out, cyc, hcyc);
    sweep<SHFL_INDEP>(out, cyc, hcyc);
    sweep<REDUX_INDEP>(out, cyc, hcyc);
    sweep<SHFL_REGLANE>(out, cyc, hcyc);
    sweep<SHFL_VARLANE>(out, cyc, hcyc);
    sweep<SHFL_REG_PLUS_FFMA2>(out, cyc, hcyc);
    return 0;
}

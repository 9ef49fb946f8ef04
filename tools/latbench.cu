// Latency microbenchmarks (single warp, dependent chains) for the instructions on the panel
// kernel's per-column critical path.  B200 / sm_100a.  Prints cycles per operation.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o latbench latbench.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

constexpr int N = 256;

__global__ void k_dfma(double* out, long long* cyc, double seed) {
    double x = seed + threadIdx.x;
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) x = fma(x, 1.0000001, 0.5);
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}
__global__ void k_dsetp(double* out, long long* cyc, double seed) {
    double x = seed + threadIdx.x, y = seed * 0.5;
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) x = (x > y) ? y + 1.0 : x + 2.0;  // DSETP + DADD/SEL dependent
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}
__global__ void k_redux(unsigned* out, long long* cyc, unsigned seed) {
    unsigned x = seed + threadIdx.x;
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) x = __reduce_max_sync(0xffffffffu, x + threadIdx.x) + 1;
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}
__global__ void k_shfl(unsigned* out, long long* cyc, unsigned seed) {
    unsigned x = seed + threadIdx.x;
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1) + 1;
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}
__global__ void k_ballot(unsigned* out, long long* cyc, unsigned seed) {
    unsigned x = seed + threadIdx.x;
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) x = __ballot_sync(0xffffffffu, x & 1) + threadIdx.x;
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}
__global__ void k_lds(unsigned* out, long long* cyc) {
    __shared__ unsigned buf[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = (i * 33 + 7) & 1023;
    __syncthreads();
    unsigned x = threadIdx.x;
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) x = buf[x];
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}
__global__ void k_ddiv(double* out, long long* cyc, double seed) {
    double x = seed + threadIdx.x;
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) x = 1.0 / x + 1.5;
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0) * (N / 64);
}
__global__ void k_drcp(double* out, long long* cyc, double seed) {
    double x = seed + threadIdx.x;
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) x = __drcp_rn(x) + 1.5;
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0) * (N / 64);
}
__global__ void k_sync(long long* cyc) {
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = (t1 - t0);
}
__global__ void k_cluster_sync(long long* cyc) {
    cg::cluster_group cl = cg::this_cluster();
    cl.sync();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) cl.sync();
    long long t1 = clock64();
    if (threadIdx.x == 0 && cl.block_rank() == 0) cyc[0] = (t1 - t0);
}
// remote store to every peer + cluster barrier + local read (the exchange pattern)
__global__ void k_cluster_push(long long* cyc, unsigned* out) {
    cg::cluster_group cl = cg::this_cluster();
    __shared__ unsigned long long slots[2][16][32];
    const int C = cl.num_blocks(), rank = cl.block_rank();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    cl.sync();
    unsigned long long acc = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
        for (int peer = warp; peer < C; peer += nw) {
            unsigned long long* dst = cl.map_shared_rank(&slots[i & 1][rank][0], peer);
            dst[lane] = acc + i;
        }
        cl.sync();
        acc += slots[i & 1][(rank + 1) % C][lane];
    }
    long long t1 = clock64();
    out[threadIdx.x] = (unsigned)acc;
    if (threadIdx.x == 0 && rank == 0) cyc[0] = (t1 - t0);
}
// st.async push of NV2 x 16 B to every peer, completion through the receiver's mbarrier (no
// cluster barrier): the candidate exchange of a panel column without cluster.sync
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa_u32(unsigned addr, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
template <int NV2>
__global__ void k_cluster_stasync(long long* cyc, unsigned* out) {
    cg::cluster_group cl = cg::this_cluster();
    __shared__ __align__(16) unsigned long long slots[2][16][2 * NV2];
    __shared__ __align__(8) unsigned long long mbar[2];
    const unsigned C = cl.num_blocks(), rank = cl.block_rank();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cl.sync();
    unsigned long long acc = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
        const int b = i & 1;
        const unsigned bar = smem_u32(&mbar[b]);
        if (threadIdx.x == 0) {
            unsigned long long st;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 %0, [%1], %2;" : "=l"(st) : "r"(bar), "r"(C * NV2 * 16u) : "memory");
        }
        for (unsigned idx = threadIdx.x; idx < C * NV2; idx += blockDim.x) {
            const unsigned peer = idx % C, chunk = idx / C;
            const unsigned raddr = mapa_u32(smem_u32(&slots[b][rank][2 * chunk]), peer);
            const unsigned rbar = mapa_u32(bar, peer);
            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];" ::"r"(raddr),
                         "l"(acc + i), "l"(acc), "r"(rbar)
                         : "memory");
        }
        const unsigned parity = (i >> 1) & 1;
        unsigned done = 0;
        while (!done) {
            asm volatile(
                "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                : "=r"(done)
                : "r"(bar), "r"(parity)
                : "memory");
        }
        acc += slots[b][(rank + 1) % C][threadIdx.x % (2 * NV2)];
        __syncthreads();
    }
    long long t1 = clock64();
    out[threadIdx.x] = (unsigned)acc;
    if (threadIdx.x == 0 && rank == 0) cyc[0] = (t1 - t0);
    cl.sync();
}
// remote load latency (pointer chase through a peer's shared memory)
__global__ void k_dsmem_load(long long* cyc, unsigned* out) {
    cg::cluster_group cl = cg::this_cluster();
    __shared__ unsigned buf[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = (i * 33 + 7) & 1023;
    cl.sync();
    const unsigned* remote = cl.map_shared_rank(buf, (cl.block_rank() + 1) % cl.num_blocks());
    unsigned x = threadIdx.x;
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) x = remote[x];
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0 && cl.block_rank() == 0) cyc[0] = (t1 - t0) * (N / 64);
    cl.sync();
}

template <class F>
void run(const char* name, F launch, long long* d_cyc) {
    launch();
    cudaDeviceSynchronize();
    launch();
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost);
    printf("{\"op\": \"%s\", \"cycles_per_op\": %.1f, \"status\": \"%s\"}\n", name, (double)c / N, cudaGetErrorString(e));
}

template <class K, class... Args>
void launch_cluster(K kern, int csize, int threads, Args... args) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(csize);
    cfg.blockDim = dim3(threads);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = csize;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, args...);
}

int main() {
    void* out;
    long long* cyc;
    cudaMalloc(&out, 1 << 20);
    cudaMalloc(&cyc, 64);
    run("dfma_dependent", [&] { k_dfma<<<1, 32>>>((double*)out, cyc, 1.0); }, cyc);
    run("dsetp_select_dependent", [&] { k_dsetp<<<1, 32>>>((double*)out, cyc, 1.0); }, cyc);
    run("redux_max_u32", [&] { k_redux<<<1, 32>>>((unsigned*)out, cyc, 1u); }, cyc);
    run("shfl_xor", [&] { k_shfl<<<1, 32>>>((unsigned*)out, cyc, 1u); }, cyc);
    run("ballot", [&] { k_ballot<<<1, 32>>>((unsigned*)out, cyc, 1u); }, cyc);
    run("lds_pointer_chase", [&] { k_lds<<<1, 32>>>((unsigned*)out, cyc); }, cyc);
    run("ddiv_1_over_x", [&] { k_ddiv<<<1, 32>>>((double*)out, cyc, 1.0); }, cyc);
    run("drcp_rn", [&] { k_drcp<<<1, 32>>>((double*)out, cyc, 1.0); }, cyc);
    run("syncthreads_128", [&] { k_sync<<<1, 128>>>(cyc); }, cyc);
    run("syncthreads_512", [&] { k_sync<<<1, 512>>>(cyc); }, cyc);
    for (int cs : {2, 4, 8, 16}) {
        char nm[64];
        snprintf(nm, sizeof nm, "cluster_sync_c%d_t128", cs);
        run(nm, [&] { launch_cluster(k_cluster_sync, cs, 128, cyc); }, cyc);
        snprintf(nm, sizeof nm, "cluster_sync_c%d_t512", cs);
        run(nm, [&] { launch_cluster(k_cluster_sync, cs, 512, cyc); }, cyc);
        snprintf(nm, sizeof nm, "cluster_push256B_sync_read_c%d_t128", cs);
        run(nm, [&] { launch_cluster(k_cluster_push, cs, 128, cyc, (unsigned*)out); }, cyc);
        snprintf(nm, sizeof nm, "stasync_push16B_mbar_wait_c%d_t128", cs);
        run(nm, [&] { launch_cluster(k_cluster_stasync<1>, cs, 128, cyc, (unsigned*)out); }, cyc);
        snprintf(nm, sizeof nm, "stasync_push16B_mbar_wait_c%d_t512", cs);
        run(nm, [&] { launch_cluster(k_cluster_stasync<1>, cs, 512, cyc, (unsigned*)out); }, cyc);
        snprintf(nm, sizeof nm, "stasync_push272B_mbar_wait_c%d_t512", cs);
        run(nm, [&] { launch_cluster(k_cluster_stasync<17>, cs, 512, cyc, (unsigned*)out); }, cyc);
        snprintf(nm, sizeof nm, "dsmem_load_chase_c%d", cs);
        run(nm, [&] { launch_cluster(k_dsmem_load, cs, 32, cyc, (unsigned*)out); }, cyc);
    }
    return 0;
}

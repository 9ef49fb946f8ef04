// Microbenchmarks for the roofline denominators the driver does not measure:
// FP64 DMMA (mma.sync.m8n8k4.f64) and DFMA issue-rate peaks on this B200.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dmma_peak(double* out, int iters) {
    double c[16][2];
    for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
    double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3 + 1.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dfma_peak(double* out, int iters) {
    double c[16];
    for (int i = 0; i < 16; ++i) c[i] = i;
    double a = 1.0000001, b = threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
    for (int i = 0; i < 16; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void ffma_peak(float* out, int iters) {
    float c[16];
    for (int i = 0; i < 16; ++i) c[i] = i;
    float a = 1.0000001f, b = threadIdx.x * 1e-9f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fmaf(c[i], a, b);
    }
    float s = 0;
    for (int i = 0; i < 16; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float time_ms(F f, int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0);
        f();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
    void* out;
    cudaMalloc(&out, (size_t)sms * 8 * 1024 * 8);
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        for (int bps : {1, 2}) {
            if (warps * bps > 64) continue;
            int threads = warps * 32;
            float ms = time_ms([&] { dmma_peak<<<sms * bps, threads>>>((double*)out, iters); }, 5);
            double flops = 2.0 * 256 * 16 * (double)iters * warps * bps * sms;
            printf("{\"bench\": \"dmma_m8n8k4\", \"warps_per_cta\": %d, \"ctas_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", warps, bps, ms,
                   flops / ms * 1e-9);
            ms = time_ms([&] { dfma_peak<<<sms * bps, threads>>>((double*)out, iters); }, 5);
            flops = 2.0 * 32 * 16 * (double)iters * warps * bps * sms;
            printf("{\"bench\": \"dfma\", \"warps_per_cta\": %d, \"ctas_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", warps, bps, ms,
                   flops / ms * 1e-9);
        }
    }
    {
        float ms = time_ms([&] { ffma_peak<<<sms * 2, 1024>>>((float*)out, iters); }, 5);
        double flops = 2.0 * 32 * 16 * (double)iters * 32 * 2 * sms;
        printf("{\"bench\": \"ffma\", \"ms\": %.3f, \"tflops\": %.2f}\n", ms, flops / ms * 1e-9);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e));
    return 0;
}

// Host <-> device marshalling for the host-pointer entry points: any (row_stride,
// col_stride) ndarray-style view (including negative and transposed strides, which the
// reference accepts -- src/lapack/getrf.rs:52-54, tests :430-483) is packed into the
// device layout (row-major, leading dimension ld) and scattered back through the same
// strides.
#include <cstring>
#include <vector>

#include "host_io.cuh"

namespace lair {

namespace {

template <class T>
__global__ void transpose_kernel(const T* __restrict__ src, long long lds, T* __restrict__ dst, long long ldd, int rows,
                                 int cols) {
    // dst[c][r] = src[r][c]; src is rows x cols
    __shared__ T tile[32][33];
    int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int r = r0 + i, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[i][threadIdx.x] = src[(long long)r * lds + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c = c0 + i, r = r0 + threadIdx.x;
        if (r < rows && c < cols) dst[(long long)c * ldd + r] = tile[threadIdx.x][i];
    }
}

template <class T>
int transpose_dev(const T* src, int64_t lds, T* dst, int64_t ldd, int64_t rows, int64_t cols, cudaStream_t s) {
    if (rows == 0 || cols == 0) return LAIR_B200_OK;
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
    LAIR_REQUIRE(grid.y <= 65535, "transpose: too many rows");
    transpose_kernel<T><<<grid, dim3(32, 8), 0, s>>>(src, lds, dst, ldd, (int)rows, (int)cols);
    LAIR_LAUNCH_CHECK();
    return LAIR_B200_OK;
}

__global__ void widen_ipiv_kernel(const int32_t* in, long long* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

}  // namespace

int DevicePool::get(int slot, size_t bytes, void** out) {
    if (slot < 0 || slot >= kSlots) return LAIR_B200_ERR_INVALID;
    if (bytes == 0) bytes = 16;
    if (cap_[slot] < bytes) {
        if (ptr_[slot]) {
            LAIR_CUDA_CHECK(cudaDeviceSynchronize());
            LAIR_CUDA_CHECK(cudaFree(ptr_[slot]));
            ptr_[slot] = nullptr;
            cap_[slot] = 0;
        }
        LAIR_CUDA_CHECK(cudaMalloc(&ptr_[slot], bytes));
        cap_[slot] = bytes;
    }
    *out = ptr_[slot];
    return LAIR_B200_OK;
}

void DevicePool::release() {
    for (int i = 0; i < kSlots; ++i) {
        if (ptr_[i]) cudaFree(ptr_[i]);
        ptr_[i] = nullptr;
        cap_[i] = 0;
    }
}

DevicePool& pool() {
    static DevicePool p;
    return p;
}

bool is_standard_layout(int64_t m, int64_t n, int64_t rs, int64_t cs) {
    // ndarray's ArrayBase::is_standard_layout for Ix2 (what getrf.rs:20 dispatches on)
    if (m == 0 || n == 0) return true;
    if (n != 1 && cs != 1) return false;
    if (m != 1 && rs != n) return false;
    return true;
}

template <class T>
int upload_matrix(const T* h, int64_t m, int64_t n, int64_t rs, int64_t cs, T* d, int64_t ld, int tmp_slot,
                  cudaStream_t s) {
    if (m == 0 || n == 0) return LAIR_B200_OK;
    if ((cs == 1 || n == 1) && (rs >= n || m == 1)) {
        // rows are contiguous: one strided DMA
        int64_t pitch = (m == 1) ? n : rs;
        LAIR_CUDA_CHECK(cudaMemcpy2DAsync(d, (size_t)ld * sizeof(T), h, (size_t)pitch * sizeof(T), (size_t)n * sizeof(T),
                                          (size_t)m, cudaMemcpyHostToDevice, s));
        return LAIR_B200_OK;
    }
    if ((rs == 1 || m == 1) && (cs >= m || n == 1)) {
        // columns are contiguous: DMA the transpose, transpose on the device
        void* tmp = nullptr;
        LAIR_CHECK(pool().get(tmp_slot, (size_t)n * m * sizeof(T), &tmp));
        int64_t pitch = (n == 1) ? m : cs;
        LAIR_CUDA_CHECK(cudaMemcpy2DAsync(tmp, (size_t)m * sizeof(T), h, (size_t)pitch * sizeof(T), (size_t)m * sizeof(T),
                                          (size_t)n, cudaMemcpyHostToDevice, s));
        return transpose_dev<T>((const T*)tmp, m, d, ld, n, m, s);
    }
    // anything else (negative / non-unit strides): gather on the host
    std::vector<T> stage((size_t)m * n);
    for (int64_t r = 0; r < m; ++r)
        for (int64_t c = 0; c < n; ++c) stage[(size_t)r * n + c] = h[r * rs + c * cs];
    LAIR_CUDA_CHECK(cudaMemcpy2DAsync(d, (size_t)ld * sizeof(T), stage.data(), (size_t)n * sizeof(T), (size_t)n * sizeof(T),
                                      (size_t)m, cudaMemcpyHostToDevice, s));
    LAIR_CUDA_CHECK(cudaStreamSynchronize(s));  // stage dies at scope exit
    return LAIR_B200_OK;
}

template <class T>
int download_matrix(T* h, int64_t m, int64_t n, int64_t rs, int64_t cs, const T* d, int64_t ld, int tmp_slot,
                    cudaStream_t s) {
    if (m == 0 || n == 0) return LAIR_B200_OK;
    if ((cs == 1 || n == 1) && (rs >= n || m == 1)) {
        int64_t pitch = (m == 1) ? n : rs;
        LAIR_CUDA_CHECK(cudaMemcpy2DAsync(h, (size_t)pitch * sizeof(T), d, (size_t)ld * sizeof(T), (size_t)n * sizeof(T),
                                          (size_t)m, cudaMemcpyDeviceToHost, s));
        return LAIR_B200_OK;
    }
    if ((rs == 1 || m == 1) && (cs >= m || n == 1)) {
        void* tmp = nullptr;
        LAIR_CHECK(pool().get(tmp_slot, (size_t)n * m * sizeof(T), &tmp));
        LAIR_CHECK(transpose_dev<T>(d, ld, (T*)tmp, m, m, n, s));
        int64_t pitch = (n == 1) ? m : cs;
        LAIR_CUDA_CHECK(cudaMemcpy2DAsync(h, (size_t)pitch * sizeof(T), tmp, (size_t)m * sizeof(T), (size_t)m * sizeof(T),
                                          (size_t)n, cudaMemcpyDeviceToHost, s));
        return LAIR_B200_OK;
    }
    std::vector<T> stage((size_t)m * n);
    LAIR_CUDA_CHECK(cudaMemcpy2DAsync(stage.data(), (size_t)n * sizeof(T), d, (size_t)ld * sizeof(T), (size_t)n * sizeof(T),
                                      (size_t)m, cudaMemcpyDeviceToHost, s));
    LAIR_CUDA_CHECK(cudaStreamSynchronize(s));
    for (int64_t r = 0; r < m; ++r)
        for (int64_t c = 0; c < n; ++c) h[r * rs + c * cs] = stage[(size_t)r * n + c];
    return LAIR_B200_OK;
}

int download_ipiv64(int64_t* h, const int32_t* d, int64_t n, int tmp_slot, cudaStream_t s) {
    if (n == 0) return LAIR_B200_OK;
    void* tmp = nullptr;
    LAIR_CHECK(pool().get(tmp_slot, (size_t)n * sizeof(long long), &tmp));
    widen_ipiv_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d, (long long*)tmp, (int)n);
    LAIR_LAUNCH_CHECK();
    LAIR_CUDA_CHECK(cudaMemcpyAsync(h, tmp, (size_t)n * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    return LAIR_B200_OK;
}

#define INST(T)                                                                                                \
    template int upload_matrix<T>(const T*, int64_t, int64_t, int64_t, int64_t, T*, int64_t, int, cudaStream_t); \
    template int download_matrix<T>(T*, int64_t, int64_t, int64_t, int64_t, const T*, int64_t, int, cudaStream_t);
INST(float)
INST(double)
INST(cxf)
INST(cxd)
#undef INST

}  // namespace lair

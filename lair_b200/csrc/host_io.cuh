// Host <-> device marshalling helpers (see host_io.cu).
#pragma once
#include "common.cuh"

namespace lair {

// Grow-only cache of device buffers for the host-pointer entry points, so steady-state
// calls do not pay cudaMalloc.  Slots are fixed roles (matrix, rhs, pivots, temporaries).
class DevicePool {
public:
    static constexpr int kSlots = 8;
    enum Slot { kMatrix = 0, kRhs = 1, kPivots = 2, kInfo = 3, kTmpA = 4, kTmpB = 5, kPivots64 = 6, kMisc = 7 };
    int get(int slot, size_t bytes, void** out);
    void release();

private:
    void* ptr_[kSlots] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t cap_[kSlots] = {0, 0, 0, 0, 0, 0, 0, 0};
};
DevicePool& pool();

bool is_standard_layout(int64_t m, int64_t n, int64_t rs, int64_t cs);

template <class T>
int upload_matrix(const T* h, int64_t m, int64_t n, int64_t rs, int64_t cs, T* d, int64_t ld, int tmp_slot, cudaStream_t s);
template <class T>
int download_matrix(T* h, int64_t m, int64_t n, int64_t rs, int64_t cs, const T* d, int64_t ld, int tmp_slot, cudaStream_t s);
int download_ipiv64(int64_t* h, const int32_t* d, int64_t n, int tmp_slot, cudaStream_t s);

}  // namespace lair

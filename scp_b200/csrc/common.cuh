// Shared helpers for the scp_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdarg.h>
#include <stdio.h>
#include <atomic>
#include "../../include/scp_b200.h"

namespace scp {

typedef unsigned long long u64;
typedef unsigned int u32;

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

// cudaMallocAsync on the device's default pool, configured once to keep freed blocks (release threshold = max): with
// the default threshold of 0 every stream/device synchronisation hands the pool back to the driver and the next step
// pays for re-mapping gigabytes of workspace (measured: +60 % step time, erratic).
cudaError_t malloc_async(void** p, size_t bytes, cudaStream_t st);

// Pinned host staging blocks, recycled: pinned_get() hands out a block whose previous upload has completed (its event
// is synchronised, normally long done); the user copies into it, issues cudaMemcpyAsync + cudaEventRecord(ev) on its
// stream and gives it back with pinned_put() at any later time.
struct PinnedBlock { void* p; size_t bytes; cudaEvent_t ev; };
bool pinned_get(size_t bytes, PinnedBlock* out);
void pinned_put(const PinnedBlock& b);
// host array -> fresh stream-ordered device buffer without synchronising the stream; free with cudaFreeAsync(*d, st)
cudaError_t upload_async(void** d, const void* h, size_t bytes, cudaStream_t st);

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

#define SCP_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            scp::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); \
            return SCP_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

#define SCP_LAUNCHED()                         \
    do {                                       \
        scp::g_launches.fetch_add(1);          \
        SCP_CUDA(cudaGetLastError());          \
    } while (0)

#define SCP_REQUIRE(cond, ...)                 \
    do {                                       \
        if (!(cond)) {                         \
            scp::set_error(__VA_ARGS__);       \
            return SCP_ERR_ARG;                \
        }                                      \
    } while (0)

// grow-only device buffer
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return SCP_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        SCP_CUDA(cudaMalloc(&p, want));
        cap = want;
        return SCP_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace scp

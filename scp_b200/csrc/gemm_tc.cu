// tcgen05 / TMEM / TMA GEMM engine (placeholder until the kernel below is validated on hardware).
#include "common.cuh"
namespace scp {
bool linear_tf32_ok(long long, long long, long long, int, int, const void*, const void*, const void*) { return false; }
int linear_tf32(const float*, long long, const float*, const float*, const float*, long long, float*, long long,
                long long, int, int, int, cudaStream_t) {
    set_error("tcgen05 engine not built");
    return SCP_ERR_STATE;
}
}  // namespace scp

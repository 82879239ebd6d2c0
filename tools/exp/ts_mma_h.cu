// Micro-experiment: tcgen05.mma kind::f16 with the A operand in TMEM ("TS" form), M=128, N=64, K=64 (4 x K16).
// A is written to TMEM with tcgen05.st (lane = row, one 32-bit column per PAIR of k: low half = even k if mode 0),
// B sits in shared memory (K-major SW128: 64 halfs = 128 B per row).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "../../scp_b200/csrc/tc.cuh"
using namespace scp;
namespace scp { void set_error(const char*, ...) {} std::atomic<long long> g_launches{0}; }

__device__ __forceinline__ void tc_mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}

// A [128][64] half, B [N][64] half (row-major, K contiguous), D [128][N] float
template <int N>
__global__ void __launch_bounds__(128, 1) k_ts(const __half* A, const __half* B, float* D, int mode) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + N * 128);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    if (warp == 0) {
        if (lane == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int f = t; f < N * 8; f += 128) {                 // B tile -> smem SW128 (rows = n, 8 chunks of 8 halfs)
        const int r = f >> 3, c = f & 7;
        const uint4 v = *reinterpret_cast<const uint4*>(B + r * 64 + c * 8);
        *reinterpret_cast<uint4*>(sm + r * 128 + ((c ^ (r & 7)) << 4)) = v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    uint32_t a[32];
    for (int c = 0; c < 32; ++c) {
        const uint32_t e = __half_as_ushort(A[t * 64 + 2 * c]), o = __half_as_ushort(A[t * 64 + 2 * c + 1]);
        a[c] = mode == 0 ? (e | (o << 16)) : (o | (e << 16));
    }
    tc_st32(trow + 256u, a);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (t == 0) {
        const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int ks = 0; ks < 4; ++ks)
            tc_mma_f16_ts(tmem, tmem + 256u + 8u * ks, make_smem_desc(sm) + 2 * ks, idesc, ks ? 1u : 0u);
        tc_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        tc_ld32(trow + c0, r);
        for (int j = 0; j < 32; ++j) D[t * N + c0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory"); }
}

template <int N>
int run(int mode) {
    std::vector<__half> A(128 * 64), B(N * 64);
    std::vector<float> D(128 * N), R(128 * N);
    srand(1);
    auto q = []() { return (float)((rand() % 2001) - 1000) / 256.0f; };     // exactly representable in fp16
    for (auto& v : A) v = __float2half(q());
    for (auto& v : B) v = __float2half(q());
    for (int i = 0; i < 128; ++i) for (int j = 0; j < N; ++j) { double s = 0; for (int k = 0; k < 64; ++k) s += (double)__half2float(A[i * 64 + k]) * __half2float(B[j * 64 + k]); R[i * N + j] = (float)s; }
    __half *dA, *dB; float* dD;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, D.size() * 4);
    const int smem = N * 128 + 2048 + 64;
    cudaFuncSetAttribute(k_ts<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k_ts<N><<<1, 128, smem>>>(dA, dB, dD, mode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d mode=%d CUDA error %s\n", N, mode, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double err = 0; for (size_t i = 0; i < D.size(); ++i) err = fmax(err, fabs((double)D[i] - R[i]));
    printf("f16 TS N=%d pack=%s max abs err %.3e  (D[0]=%f ref %f, D[last]=%f ref %f)\n", N, mode ? "odd-low" : "even-low", err, D[0], R[0], D.back(), R.back());
    return err < 1e-3 ? 0 : 2;
}

int main() {
    int a = run<64>(0), b = run<64>(1), c = run<128>(0), d = run<256>(0);
    printf("even-low %s, odd-low %s, N=128 %s, N=256 %s\n", a ? "WRONG" : "OK", b ? "WRONG" : "OK", c ? "WRONG" : "OK", d ? "WRONG" : "OK");
    return 0;
}

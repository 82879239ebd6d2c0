// Micro-experiment: tcgen05.mma with the A operand in TMEM ("TS" form), kind::tf32, M=128, N=64, K=32.
// A is written to TMEM with tcgen05.st (lane = row, one 32-bit column per k), B sits in shared memory (K-major SW128).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../../scp_b200/csrc/tc.cuh"
using namespace scp;
namespace scp { void set_error(const char*, ...) {} std::atomic<long long> g_launches{0}; }

__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),
          "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]),
          "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}

// A [128][32], B [N][32] (row-major, K contiguous), D [128][N]
template <int N>
__global__ void __launch_bounds__(128, 1) k_ts(const float* A, const float* B, float* D, int mode) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + N * 128 + 128 * 128);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    if (warp == 0) {
        if (lane == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // B tile -> smem SW128 (rows = n, 32 floats = 128 B per row)
    for (int f = t; f < N * 8; f += 128) {
        const int r = f >> 3, c = f & 7;
        const float4 v = *reinterpret_cast<const float4*>(B + r * 32 + c * 4);
        *reinterpret_cast<float4*>(sm + r * 128 + ((c ^ (r & 7)) << 4)) = v;
    }
    // A tile -> smem too (for the SS reference run), after the B tile
    uint8_t* sa = sm + N * 128;
    for (int f = t; f < 128 * 8; f += 128) {
        const int r = f >> 3, c = f & 7;
        const float4 v = *reinterpret_cast<const float4*>(A + r * 32 + c * 4);
        *reinterpret_cast<float4*>(sa + r * 128 + ((c ^ (r & 7)) << 4)) = v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    // A row of this thread -> TMEM columns [256, 288)
    uint32_t a[32];
    for (int k = 0; k < 32; ++k) a[k] = __float_as_uint(A[t * 32 + k]);
    tc_st32(trow + 256u, a);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (t == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int ks = 0; ks < 4; ++ks) {
            if (mode == 0) tc_mma_tf32_ts(tmem, tmem + 256u + 8u * ks, make_smem_desc(sm) + 2 * ks, idesc, ks ? 1u : 0u);
            else tc_mma_tf32(tmem, make_smem_desc(sa) + 2 * ks, make_smem_desc(sm) + 2 * ks, idesc, ks ? 1u : 0u);
        }
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
    std::vector<float> A(128 * 32), B(N * 32), D(128 * N), R(128 * N);
    srand(1);
    auto q = []() { return (float)((rand() % 2001) - 1000) / 256.0f; };     // exactly representable in tf32
    for (auto& v : A) v = q();
    for (auto& v : B) v = q();
    for (int i = 0; i < 128; ++i) for (int j = 0; j < N; ++j) { double s = 0; for (int k = 0; k < 32; ++k) s += (double)A[i * 32 + k] * B[j * 32 + k]; R[i * N + j] = (float)s; }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, D.size() * 4);
    const int smem = N * 128 + 128 * 128 + 1024 + 64;
    cudaFuncSetAttribute(k_ts<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k_ts<N><<<1, 128, smem>>>(dA, dB, dD, mode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d mode=%d CUDA error %s\n", N, mode, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double err = 0; for (size_t i = 0; i < D.size(); ++i) err = fmax(err, fabs((double)D[i] - R[i]));
    printf("N=%d mode=%s max abs err %.3e  (D[0]=%f ref %f, D[last]=%f ref %f)\n", N, mode ? "SS" : "TS", err, D[0], R[0], D.back(), R.back());
    return err < 1e-3 ? 0 : 2;
}

int main() {
    int rc = 0;
    rc |= run<64>(1); rc |= run<64>(0); rc |= run<32>(0); rc |= run<128>(0); rc |= run<256>(1); rc |= run<256>(0);
    printf(rc ? "FAILED\n" : "ALL OK\n");
    return rc;
}

// Micro-experiment: register <-> (lane, column) mapping of tcgen05.ld.16x256b / tcgen05.st.16x128b.
// TMEM is filled with value(lane, col) = lane * 1000 + col through 32x32b stores (lane = row: the known layout), read back
// with 16x256b.x8 from the 16-lane half at lane offset 16 h of every warp's quarter; then the inverse for 16x128b stores.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <atomic>
#include <cuda_runtime.h>
#include "../../scp_b200/csrc/tc.cuh"
using namespace scp;
namespace scp { void set_error(const char*, ...) {} std::atomic<long long> g_launches{0}; }

__device__ __forceinline__ void ld_16x256b_x8(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_16x128b_x8(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.16x128b.x8.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// out_ld[(w * 2 + h) * 32 * 32 + lane * 32 + reg] = value seen; out_st[lane128 * 32 + col] = value found after the 16x128b stores
__global__ void __launch_bounds__(128, 1) k_probe(uint32_t* out_ld, uint32_t* out_st) {
    __shared__ uint32_t slot;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    uint32_t a[32];
    for (int half = 0; half < 2; ++half) {                 // 64 columns: value = lane * 1000 + col
        for (int c = 0; c < 32; ++c) a[c] = (uint32_t)(t * 1000 + half * 32 + c);
        tc_st32(trow + (uint32_t)(half * 32), a);
    }
    tc_wait_st();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    for (int h = 0; h < 2; ++h) {
        uint32_t r[32];
        ld_16x256b_x8(tmem + ((uint32_t)(warp * 32 + 16 * h) << 16), r);
        for (int j = 0; j < 32; ++j) out_ld[((warp * 2 + h) * 32 + lane) * 32 + j] = r[j];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // inverse: 16x128b.x8 stores into columns [64, 96): register j of thread (warp, lane, h) carries the tag below
    for (int h = 0; h < 2; ++h) {
        uint32_t r[16];
        for (int j = 0; j < 16; ++j) r[j] = (uint32_t)(((warp * 2 + h) * 32 + lane) * 100 + j);
        st_16x128b_x8(tmem + ((uint32_t)(warp * 32 + 16 * h) << 16) + 64u, r);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    tc_ld32(trow + 64u, a);
    for (int c = 0; c < 32; ++c) out_st[t * 32 + c] = a[c];
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
    }
}

int main() {
    uint32_t *d_ld, *d_st;
    cudaMalloc(&d_ld, 8 * 32 * 32 * 4); cudaMalloc(&d_st, 128 * 32 * 4);
    k_probe<<<1, 128>>>(d_ld, d_st);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<uint32_t> ld(8 * 32 * 32), st(128 * 32);
    cudaMemcpy(ld.data(), d_ld, ld.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(st.data(), d_st, st.size() * 4, cudaMemcpyDeviceToHost);
    // hypothesis LD: reg 4n+0/1 -> (lane base + l/4, col 8n + 2(l%4) + {0,1}); reg 4n+2/3 -> row + 8
    int bad = 0;
    for (int wh = 0; wh < 8; ++wh) for (int l = 0; l < 32; ++l) for (int j = 0; j < 32; ++j) {
        const int n = j >> 2, k = j & 3;
        const int row = (wh >> 1) * 32 + (wh & 1) * 16 + l / 4 + (k >> 1) * 8, col = 8 * n + 2 * (l % 4) + (k & 1);
        const uint32_t want = (uint32_t)(row * 1000 + col), got = ld[(wh * 32 + l) * 32 + j];
        if (want != got && bad++ < 8) printf("LD mismatch warp-half %d lane %d reg %d: got lane %u col %u, hypothesis lane %d col %d\n", wh, l, j, got / 1000, got % 1000, row, col);
    }
    printf("16x256b.x8 load hypothesis: %s (%d mismatches)\n", bad ? "WRONG" : "confirmed", bad);
    for (int l = 0; l < 4; ++l) { printf("  warp 0 half 0 lane %d:", l); for (int j = 0; j < 8; ++j) printf(" (%u,%u)", ld[l * 32 + j] / 1000, ld[l * 32 + j] % 1000); printf("\n"); }
    // hypothesis ST 16x128b.x8: reg 2n+0 -> (row l/4, col 4n + l%4), reg 2n+1 -> (row l/4 + 8, same col)
    bad = 0;
    for (int row = 0; row < 128; ++row) for (int c = 0; c < 32; ++c) {
        const int w = row / 32, h = (row % 32) / 16, rr = row % 16, k = rr / 8, l = (rr % 8) * 4 + c % 4, n = c / 4;
        const uint32_t want = (uint32_t)(((w * 2 + h) * 32 + l) * 100 + 2 * n + k), got = st[row * 32 + c];
        if (want != got && bad++ < 8) printf("ST mismatch lane %d col %d: got thread %u reg %u, hypothesis thread %d reg %d\n", row, c, got / 100, got % 100, (w * 2 + h) * 32 + l, 2 * n + k);
    }
    printf("16x128b.x8 store hypothesis: %s (%d mismatches)\n", bad ? "WRONG" : "confirmed", bad);
    for (int c = 0; c < 8; ++c) printf("  lane 0 col %d <- thread %u reg %u;", c, st[c] / 100, st[c] % 100);
    printf("\n");
    for (int c = 0; c < 4; ++c) printf("  lane 8 col %d <- thread %u reg %u;", c, st[8 * 32 + c] / 100, st[8 * 32 + c] % 100);
    printf("\n");
    return 0;
}

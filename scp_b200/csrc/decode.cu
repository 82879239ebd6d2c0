// Device side of the decode path: the decoder rebuilds every octree level from the occupancy
// bytes it has decoded so far.  Two kernels:
//   k_decode_level_inputs  node state of a level -> the context bytes and normalised positions the entropy model is fed
//                          (exactly what k_context produces on the encode side: encode_dataset_ehem.py:54,66-72,86)
//   k_expand_children      decoded occupancy bytes -> the next level's nodes in BFS order (parents in order, child digit
//                          ascending; decode_ehem.py:116-140, Octree.py:68-99 DeOctree): per-parent popcount, an exclusive
//                          scan with a decoupled look-back over 1024-parent tiles, one thread per parent writes its children
#include "common.cuh"

namespace scp {

__global__ void __launch_bounds__(256) k_decode_level_inputs(const int* __restrict__ pos, const uint8_t* __restrict__ anc,
                                                              const uint8_t* __restrict__ octant, long long n, int level,
                                                              int clip, double mn, double den, uint8_t* __restrict__ ctx,
                                                              uint8_t* __restrict__ ctx_model, float* __restrict__ pos_norm) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    uint8_t c[12];
#pragma unroll
    for (int e = 0; e < 9; ++e) c[e] = anc[i * 9 + e];
    c[9] = (uint8_t)level; c[10] = octant[i]; c[11] = 255;                // self: occupancy unknown
#pragma unroll
    for (int e = 0; e < 12; ++e) ctx[i * 12 + e] = c[e];
#pragma unroll
    for (int k = 0; k < 4; ++k) c[3 * k] = (uint8_t)min((int)c[3 * k], clip);      // encode_dataset_ehem.py:86 (clip = 255: none)
#pragma unroll
    for (int e = 0; e < 12; ++e) ctx_model[i * 12 + e] = c[e];
#pragma unroll
    for (int a = 0; a < 3; ++a) pos_norm[i * 3 + a] = (float)(((double)pos[i * 3 + a] - mn) / den);     // encode_dataset_ehem.py:70-72
}

constexpr int EX_TILE = 1024;
constexpr u32 EX_AGG = 1u << 30, EX_INC = 1u << 31, EX_MASK = (1u << 30) - 1;

__global__ void __launch_bounds__(256) k_expand_children(const uint8_t* __restrict__ occ, const int* __restrict__ pos,
                                                          const uint8_t* __restrict__ ctx, long long n, int level, int cell,
                                                          int* __restrict__ cpos, uint8_t* __restrict__ canc,
                                                          uint8_t* __restrict__ coct, u32* desc, u32* ticket, u32* err) {
    __shared__ u32 s_w[8], s_base;
    __shared__ int s_tile;
    if (threadIdx.x == 0) s_tile = (int)atomicAdd(ticket, 1u);
    __syncthreads();
    const int tile = s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // thread t owns parents [tile*1024 + 4 t, +4)
    const long long p0 = (long long)tile * EX_TILE + 4 * threadIdx.x;
    u32 o[4], cnt = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) { o[q] = p0 + q < n ? occ[p0 + q] : 0u; cnt += __popc(o[q]); }
    u32 inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const u32 v = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += v; }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 total = 0;
        for (int w = 0; w < 8; ++w) total += s_w[w];
        volatile u32* vd = desc;
        vd[tile] = total | EX_AGG;
        u32 base = 0;
        for (int pt = tile - 1; pt >= 0; --pt) {                        // tickets are handed out in launch order
            u32 v;
            int spins = 0;
            do { v = vd[pt]; } while ((v & (EX_AGG | EX_INC)) == 0 && ++spins < (1 << 22));
            if ((v & (EX_AGG | EX_INC)) == 0) { atomicExch(err, 1u); break; }
            base += v & EX_MASK;
            if (v & EX_INC) break;
        }
        vd[tile] = ((base + total) & EX_MASK) | EX_INC;
        s_base = base;
    }
    __syncthreads();
    u32 at = s_base + inc - cnt;
    for (int w = 0; w < warp; ++w) at += s_w[w];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (o[q] == 0u) continue;
        const long long p = p0 + q;
        const int px = pos[p * 3], py = pos[p * 3 + 1], pz = pos[p * 3 + 2];
        uint8_t a[9];
#pragma unroll
        for (int e = 0; e < 6; ++e) a[e] = ctx[p * 12 + 3 + e];           // the parent's rows 1, 2 become rows 0, 1
        a[6] = (uint8_t)level; a[7] = ctx[p * 12 + 10]; a[8] = (uint8_t)(o[q] - 1u);     // the parent itself, occupancy now known
        u32 bits = o[q];
        while (bits) {                                                   // child digit ascending (bit d <-> digit d, Octree.py:175)
            const int d = __ffs(bits) - 1;
            bits &= bits - 1;
            cpos[(long long)at * 3] = px + ((d >> 2) & 1) * cell;
            cpos[(long long)at * 3 + 1] = py + ((d >> 1) & 1) * cell;
            cpos[(long long)at * 3 + 2] = pz + (d & 1) * cell;
#pragma unroll
            for (int e = 0; e < 9; ++e) canc[(long long)at * 9 + e] = a[e];
            coct[at] = (uint8_t)(d + 1);
            ++at;
        }
    }
}

}  // namespace scp

using namespace scp;

extern "C" {

int scp_decode_level_inputs(const int32_t* d_pos, const uint8_t* d_anc, const uint8_t* d_octant, int64_t n, int level,
                            int clip_level, double pos_min, double pos_den, uint8_t* d_ctx, uint8_t* d_ctx_model,
                            float* d_pos_norm, void* stream) {
    SCP_REQUIRE(d_pos && d_anc && d_octant && d_ctx && d_ctx_model && d_pos_norm && n >= 0, "scp_decode_level_inputs: bad argument");
    if (n == 0) return SCP_OK;
    k_decode_level_inputs<<<(unsigned)cdiv(n, 256), 256, 0, as_stream(stream)>>>(d_pos, d_anc, d_octant, n, level, clip_level,
                                                                                 pos_min, pos_den, d_ctx, d_ctx_model, d_pos_norm);
    SCP_LAUNCHED();
    return SCP_OK;
}

int scp_expand_children(const uint8_t* d_occ, const int32_t* d_pos, const uint8_t* d_ctx, int64_t n, int level, int cell,
                        int32_t* d_child_pos, uint8_t* d_child_anc, uint8_t* d_child_octant, void* stream) {
    SCP_REQUIRE(d_occ && d_pos && d_ctx && d_child_pos && d_child_anc && d_child_octant && n >= 0, "scp_expand_children: bad argument");
    if (n == 0) return SCP_OK;
    cudaStream_t st = as_stream(stream);
    const int n_tile = (int)cdiv(n, EX_TILE);
    u32* ws = nullptr;
    SCP_CUDA(malloc_async((void**)&ws, (size_t)(n_tile + 2) * 4, st));
    SCP_CUDA(cudaMemsetAsync(ws, 0, (size_t)(n_tile + 2) * 4, st));
    k_expand_children<<<n_tile, 256, 0, st>>>(d_occ, d_pos, d_ctx, n, level, cell, d_child_pos, d_child_anc, d_child_octant,
                                              ws, ws + n_tile, ws + n_tile + 1);
    SCP_LAUNCHED();
    SCP_CUDA(cudaFreeAsync(ws, st));
    return SCP_OK;
}

}  // extern "C"

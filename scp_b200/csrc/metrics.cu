// Distortion side of the encode report: dequantised voxel centres from the Morton keys of
// the octree build, and exact nearest-neighbour distances between two point sets -- what pt.py:88-95 (distChamfer,
// two KDTree queries) and the pc_error D1 metric need.  Integer de-interleave + float64 arithmetic; no tensor cores.
#include "common.cuh"

namespace scp {

static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// every third bit of a 63-bit Morton key, starting at bit `s`, packed into the low 21 bits
__device__ __forceinline__ u32 compact3(u64 k, int s) {
    u64 x = (k >> s) & 0x1249249249249249ull;
    x = (x | (x >> 2)) & 0x10c30c30c30c30c3ull;
    x = (x | (x >> 4)) & 0x100f00f00f00f00full;
    x = (x | (x >> 8)) & 0x001f0000ff0000ffull;
    x = (x | (x >> 16)) & 0x001f00000000ffffull;
    x = (x | (x >> 32)) & 0x00000000001fffffull;
    return (u32)x;
}

// v * steps + offset, then spher2cart / cylin2cart (data_preprocess.py:179-229), all in float64 like mul_proc_pc :160-167.
// Key bit triple b holds (x, y, z) at bits (3b+2, 3b+1, 3b), the digit order of Octree.cpp's Morton code.
__global__ void __launch_bounds__(256) k_dequantise_keys(const int64_t* __restrict__ keys, int64_t n, double sx, double sy,
                                                         double sz, double ox, double oy, double oz, int mode,
                                                         double* __restrict__ xyz) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const u64 k = (u64)keys[i];
    const double a = (double)compact3(k, 2) * sx + ox;
    const double b = (double)compact3(k, 1) * sy + oy;
    const double c = (double)compact3(k, 0) * sz + oz;
    double x = a, y = b, z = c;
    if (mode == SCP_MODE_SPHER) {
        double sb, cb, sc, cc;
        sincos(b, &sb, &cb);
        sincos(c, &sc, &cc);
        x = a * sc * cb;
        y = a * sc * sb;
        z = a * cc;
    } else if (mode == SCP_MODE_CYLIN) {
        double sb, cb;
        sincos(b, &sb, &cb);
        x = a * cb;
        y = a * sb;
    }
    xyz[3 * i + 0] = x;
    xyz[3 * i + 1] = y;
    xyz[3 * i + 2] = z;
}

constexpr int NN_TPB = 256;      // threads per block
constexpr int NN_QPT = 2;        // queries per thread: one shared-memory candidate read feeds two distance chains
constexpr int NN_TILE = 1024;    // candidates staged per round (24 KB of float64 coordinates)

// Brute-force exact nearest neighbour: d2[i] = min_j |q_i - c_j|^2 with the sum ((dx^2 + dy^2) + dz^2) rounded after every
// operation, i.e. the value a scalar float64 loop over the candidates produces.  blockIdx.y cuts the candidate set into
// chunks so that small query sets still fill the machine; the chunks meet in an atomicMin on the bit pattern (non-negative
// doubles order like their bits).  FP64-pipe bound: 7 DFMA-class instructions per (query, candidate) pair.
__global__ void __launch_bounds__(NN_TPB) k_nn_dist2(const double* __restrict__ q, int64_t nq, const double* __restrict__ c,
                                                     int64_t nc, int64_t chunk, u64* __restrict__ d2) {
    __shared__ double s_x[NN_TILE], s_y[NN_TILE], s_z[NN_TILE];
    const int64_t q0 = (int64_t)blockIdx.x * (NN_TPB * NN_QPT) + threadIdx.x;
    double qx[NN_QPT], qy[NN_QPT], qz[NN_QPT], best[NN_QPT];
#pragma unroll
    for (int j = 0; j < NN_QPT; ++j) {
        const int64_t i = q0 + (int64_t)j * NN_TPB;
        const bool ok = i < nq;
        qx[j] = ok ? q[3 * i + 0] : 0.0;
        qy[j] = ok ? q[3 * i + 1] : 0.0;
        qz[j] = ok ? q[3 * i + 2] : 0.0;
        best[j] = __longlong_as_double(0x7ff0000000000000ll);
    }
    const int64_t c_begin = (int64_t)blockIdx.y * chunk;
    const int64_t c_end = min(nc, c_begin + chunk);
    for (int64_t t = c_begin; t < c_end; t += NN_TILE) {
        const int m = (int)min((int64_t)NN_TILE, c_end - t);
        __syncthreads();
        for (int e = threadIdx.x; e < 3 * m; e += NN_TPB) {            // coalesced read of the [m,3] slab
            const double v = c[3 * t + e];
            const int p = e / 3, a = e - 3 * p;
            (a == 0 ? s_x : a == 1 ? s_y : s_z)[p] = v;
        }
        __syncthreads();
#pragma unroll 4
        for (int p = 0; p < m; ++p) {
            const double cx = s_x[p], cy = s_y[p], cz = s_z[p];
#pragma unroll
            for (int j = 0; j < NN_QPT; ++j) {
                const double dx = qx[j] - cx, dy = qy[j] - cy, dz = qz[j] - cz;
                const double d = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                best[j] = fmin(best[j], d);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NN_QPT; ++j) {
        const int64_t i = q0 + (int64_t)j * NN_TPB;
        if (i < nq) atomicMin(&d2[i], (u64)__double_as_longlong(best[j]));
    }
}

__global__ void k_fill_u64(u64* p, int64_t n, u64 v) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace scp

using namespace scp;

extern "C" {

int scp_dequantise_keys(const int64_t* d_keys, int64_t n, const double* h_steps, const double* h_offset, int mode,
                        double* d_xyz, void* stream) {
    SCP_REQUIRE(d_keys && h_steps && h_offset && d_xyz && n >= 0, "scp_dequantise_keys: bad argument");
    SCP_REQUIRE(mode == SCP_MODE_CART || mode == SCP_MODE_SPHER || mode == SCP_MODE_CYLIN, "scp_dequantise_keys: bad mode %d", mode);
    if (n == 0) return SCP_OK;
    k_dequantise_keys<<<(unsigned)cdiv64(n, 256), 256, 0, as_stream(stream)>>>(d_keys, n, h_steps[0], h_steps[1], h_steps[2],
                                                                              h_offset[0], h_offset[1], h_offset[2], mode, d_xyz);
    SCP_LAUNCHED();
    return SCP_OK;
}

int scp_nn_dist2(const double* d_query, int64_t n_query, const double* d_cand, int64_t n_cand, double* d_dist2, void* stream) {
    SCP_REQUIRE(d_query && d_cand && d_dist2 && n_query >= 0, "scp_nn_dist2: bad argument");
    SCP_REQUIRE(n_cand >= 1, "scp_nn_dist2: empty candidate set");
    if (n_query == 0) return SCP_OK;
    cudaStream_t st = as_stream(stream);
    int dev = 0, n_sm = 148;
    SCP_CUDA(cudaGetDevice(&dev));
    SCP_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    const int64_t gx = cdiv64(n_query, NN_TPB * NN_QPT);
    int64_t gy = cdiv64(4 * (int64_t)n_sm, gx);                       // >= 4 blocks per SM when the query set is small
    gy = max((int64_t)1, min(gy, cdiv64(n_cand, NN_TILE)));
    const int64_t chunk = cdiv64(cdiv64(n_cand, gy), NN_TILE) * NN_TILE;
    gy = cdiv64(n_cand, chunk);
    k_fill_u64<<<(unsigned)cdiv64(n_query, 256), 256, 0, st>>>((u64*)d_dist2, n_query, 0x7ff0000000000000ull);
    SCP_LAUNCHED();
    k_nn_dist2<<<dim3((unsigned)gx, (unsigned)gy), NN_TPB, 0, st>>>(d_query, n_query, d_cand, n_cand, chunk, (u64*)d_dist2);
    SCP_LAUNCHED();
    return SCP_OK;
}

}  // extern "C"

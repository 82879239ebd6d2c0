// Batched octree construction on B200.
//
//   points (float32) --K1--> per-frame rho max / z min
//                    --K2--> quantised (r,phi,theta|z) -> 63-bit Morton key per point   [20 B/pt]
//                    --K3--> segmented LSD radix sort, 8-bit digits, decoupled look-back [(1+2P)*8 B/pt]
//                    --K4--> "head level" of every sorted key + per-tile level histograms [8 B/pt]
//                    --K5--> level counts / offsets per job
//                    --K6--> node records of ALL levels in one pass over the sorted keys   [28 B/node]
//                    --K7a-> occupancy bytes from the children run of every node
//                    --K7b-> K=4 ancestor context bytes + level-normalised positions       [60 B/node]
//
// One pass over the sorted voxel keys yields every level at once: for consecutive distinct keys the
// number d of common leading octal digits says that the second key opens a new node on every level
// L >= d+2 ("head level" h = d+2).  A node's BFS index inside level L is the number of earlier keys
// with h <= L, i.e. one prefix sum per level, done with warp ballots.
//
// Reference behaviour reproduced (paths in luoao-kddi/SCP): data_preprocess.py:13-167,171-207;
// Octree.py:56-65,102-145,148-272; OctreeCPP/Octreewarpper.py; encode_dataset_ehem.py:52-105;
// encode_dataset_ehem_mullevel.py:47-85.
#include <vector>
#include <algorithm>
#include <string.h>
#include <stdlib.h>
#include "common.cuh"

namespace scp {

constexpr int TPB = 256;
constexpr int TILE = 2048;            // elements per tile for the streaming kernels
constexpr int NODE_TILE = 8192;       // nodes per tile of k_occupancy / k_context: the per-tile prologue (tile -> job -> level
                                      // tables, a dependent chain of global loads) is amortised over 8 rounds per thread
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = TPB * SORT_ITEMS;   // 4096 keys
constexpr int NBINS = 24;             // head-level histogram bins (0 = not a new voxel, 1..22)
constexpr int MAXL = SCP_MAX_DEPTH;   // 21
constexpr u64 SENTINEL = ~0ull;

struct Tile { int job; int begin; int count; int first; long long kb = 0; int pj = 0; int _pad = 0; };
// sort tiles only: kb = first key of the job, pj = digit passes of the job under the per-job schedule (0: all passes)

struct FrameDev { u32 rho_max_bits; u32 zmin_enc; u32 zmax_enc; u32 _pad; };

struct JobDev {
    int frame, path_len, path_bits, drop_last;
    double qs, cart_offset;
    int lidar_level, pos_eps_last;
    long long pt_begin;
    int n_points;
    long long key_begin;
    float bin_num;
    double step[3];
    double off[3];
    double inv_step[3];          // fast-path reciprocal steps
    double margin[3];            // fast-path safety margin (bins) around a .5 rounding boundary
    u32 qmax;
    u32 overflow;
    int depth;
    int depth_known;             // 1: `depth` was derived from the frame statistics BEFORE quantising (fused quantise path)
    int n_kept;                  // keys that take part in the sort (after the morton_path filter); = n_points without a filter
    int n_voxels, n_nodes, n_rows;
    int level_count[MAXL + 1];   // nodes on level L at [L-1]
    int level_start[MAXL + 2];   // node offset of level L inside the job at [L-1]
    long long node_start, row_start, vox_start;
    u32 pos_min[MAXL + 1];
    u32 pos_max[MAXL + 1];
};

// ------------------------------------------------------------------------------------------
// bit tricks
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ u32 spread3_11(u32 x) {   // 11 bits -> bits 0, 3, ..., 30 (32-bit operations only)
    x &= 0x7ffu;
    x = (x | x << 16) & 0x070000ffu;
    x = (x | x << 8) & 0x0700f00fu;
    x = (x | x << 4) & 0x430c30c3u;
    x = (x | x << 2) & 0x49249249u;
    return x;
}
__host__ __device__ __forceinline__ u64 spread3(u32 v) {      // 21 bits -> every third bit: low 11 bits + high 10 bits << 33
    return (u64)spread3_11(v) | ((u64)(spread3_11(v >> 11) << 1) << 32);
}
__host__ __device__ __forceinline__ u32 compact3(u64 x) {
    x &= 0x1249249249249249ull;
    x = (x ^ (x >> 2)) & 0x10c30c30c30c30c3ull;
    x = (x ^ (x >> 4)) & 0x100f00f00f00f00full;
    x = (x ^ (x >> 8)) & 0x1f0000ff0000ffull;
    x = (x ^ (x >> 16)) & 0x1f00000000ffffull;
    x = (x ^ (x >> 32)) & 0x1fffffull;
    return (u32)x;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ u32 enc_ordered(float f) {
    u32 b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec_ordered(u32 e) {
    return __uint_as_float((e & 0x80000000u) ? (e & 0x7fffffffu) : ~e);
}

// ------------------------------------------------------------------------------------------
// K1: per-frame rho max (and z min for cylindrical)            data_preprocess.py:44-46,49
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float rho_of(float x, float y, float z, int mode) {
    float s = __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y));          // x**2 + y**2 (float32, no FMA)
    if (mode == SCP_MODE_SPHER) s = __fadd_rn(s, __fmul_rn(z, z));
    return __fsqrt_rn(s);
}

__global__ void k_init_frames(FrameDev* fr, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { fr[i].rho_max_bits = 0u; fr[i].zmin_enc = 0xffffffffu; fr[i].zmax_enc = 0u; }
}

__global__ void __launch_bounds__(TPB) k_frame_stats(const float* __restrict__ xyz, int stride,
                                                      const Tile* __restrict__ tiles, const long long* __restrict__ frame_begin,
                                                      FrameDev* fr, int mode) {
    Tile t = tiles[blockIdx.x];
    const float* p = xyz + (frame_begin[t.job] + t.begin) * (long long)stride;
    float rmax = 0.f;
    u32 zmin = 0xffffffffu, zmax = 0u;
    const bool v4 = stride == 4 && (reinterpret_cast<uintptr_t>(p) & 15) == 0;      // KITTI rows (x, y, z, intensity): one 16-byte load
#pragma unroll 4
    for (int i = threadIdx.x; i < t.count; i += TPB) {
        float x, y, z;
        if (v4) { const float4 v = __ldg(reinterpret_cast<const float4*>(p) + i); x = v.x; y = v.y; z = v.z; }
        else { x = p[(long long)i * stride]; y = p[(long long)i * stride + 1]; z = p[(long long)i * stride + 2]; }
        rmax = fmaxf(rmax, rho_of(x, y, z, mode));
        const u32 ze = enc_ordered(z);
        zmin = min(zmin, ze); zmax = max(zmax, ze);
    }
    u32 rb = __reduce_max_sync(0xffffffffu, __float_as_uint(rmax));
    u32 zb = __reduce_min_sync(0xffffffffu, zmin), zt = __reduce_max_sync(0xffffffffu, zmax);
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&fr[t.job].rho_max_bits, rb);
        atomicMin(&fr[t.job].zmin_enc, zb);
        atomicMax(&fr[t.job].zmax_enc, zt);
    }
}

__global__ void k_job_setup(JobDev* jobs, int n_jobs, const FrameDev* fr, int mode) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_jobs) return;
    JobDev& J = jobs[j];
    J.qmax = 0; J.overflow = 0; J.depth = 0; J.depth_known = 0; J.n_kept = J.n_points;
    for (int l = 0; l <= MAXL; ++l) { J.pos_min[l] = 0xffffffffu; J.pos_max[l] = 0u; }
    if (mode == SCP_MODE_CART) {
        J.bin_num = 0.f;
        J.step[0] = J.step[1] = J.step[2] = J.qs;
        J.off[0] = J.off[1] = J.off[2] = J.cart_offset;
        return;
    }
    float rho_max = __uint_as_float(fr[J.frame].rho_max_bits);
    // bin_num = np.round(rho.max() / qs) + 1   (float32 arithmetic under numpy>=2)
    float bn = __fadd_rn(rintf(__fdiv_rn(rho_max, (float)J.qs)), 1.0f);
    J.bin_num = bn;
    float bm1 = __fsub_rn(bn, 1.0f);
    J.step[0] = J.qs;
    J.step[1] = (double)__fdiv_rn(6.2831854820251465f, bm1);     // float32(2*pi) / float32
    J.off[0] = 0.0; J.off[1] = 0.0; J.off[2] = 0.0;
    if (mode == SCP_MODE_SPHER) {
        J.step[2] = (double)__fdiv_rn(3.1415927410125732f, bm1);
    } else {
        J.step[2] = J.qs;
        J.off[2] = (double)dec_ordered(fr[J.frame].zmin_enc);
    }
    // Fast path of k_quantise_keys: float32 atan2f/acosf are within 2 ulp of the correctly rounded float32 angle the
    // slow (float64) path produces; with the +2*pi fold that is < 1.5e-6 rad, and multiplying by the reciprocal step
    // instead of dividing moves the bin coordinate by < 1e-10 bins.  Points whose fast bin coordinate is closer than
    // `margin` to a .5 boundary are recomputed exactly, everything else is provably identical.
    for (int c = 0; c < 3; ++c) J.inv_step[c] = 1.0 / J.step[c];
    J.margin[0] = 1e-7;
    J.margin[1] = 1.5e-6 * J.inv_step[1] + 1e-7;
    J.margin[2] = (mode == SCP_MODE_SPHER) ? 1.5e-6 * J.inv_step[2] + 1e-7 : 1e-7;
    // Depth of the octree (Octree.py:58: bit length of the largest quantised coordinate over all three axes, BEFORE any
    // morton_path filter) from the frame statistics alone, so that the fused quantise kernel can apply the filter -- which
    // tests bits depth-1-j of the radial coordinate -- before it spends any work on a point:
    //   radial axis: rint(rho / qs) is monotone in rho, so its maximum is rint(rho_max / qs), attained;
    //   cylindrical z axis: likewise rint((z_max - z_min) / qs);
    //   angular axes: phi <= float32(2 pi), theta <= float32(pi), and the steps are float32(2 pi | pi) / (bin_num - 1)
    //   rounded to float32, so q <= (bin_num - 1) * (1 + 2^-24) rounds to at most bin_num - 1 (bin_num <= 2^21).
    // lo = largest attained coordinate, hi = upper bound; when both have the same bit length the depth is known.  Otherwise
    // (bin_num - 1 a power of two and the radial maximum one short of it) depth_known stays 0 and the host takes the
    // two-kernel path (quantise everything, then filter).
    {
        const double rmax = (double)rho_max;
        long long lo = (long long)rint(rmax / J.step[0]);
        if (mode == SCP_MODE_CYLIN) {
            const double zspan = (double)dec_ordered(fr[J.frame].zmax_enc) - J.off[2];
            lo = max(lo, (long long)rint(zspan / J.step[2]));
        }
        const long long hi = max(lo, (long long)bm1);
        const int dl = lo > 0 ? 64 - __clzll(lo) : 0, dh = hi > 0 ? 64 - __clzll(hi) : 0;
        if (dl == dh && dl >= 1 && dl <= MAXL && lo < (1ll << 21)) { J.depth = dl; J.depth_known = 1; }
    }
}

// ------------------------------------------------------------------------------------------
// K2: coordinate transform + quantise + Morton key            data_preprocess.py:171-207, :56,:68-70
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool quantise_exact(float x, float y, float z, int mode, double s0, double s1, double s2, double o2,
                                               long long& q0, long long& q1, long long& q2) {
    float rho = rho_of(x, y, z, mode);
    float xe = __fadd_rn(x, 1e-9f);
    float phi = (float)atan2((double)y, (double)xe);           // correctly rounded float32 angle
    if (phi < 0.f) phi = __fadd_rn(phi, 6.2831854820251465f);
    float third = (mode == SCP_MODE_SPHER) ? (float)acos((double)__fdiv_rn(z, rho)) : z;
    q0 = (long long)rint((double)rho / s0);
    q1 = (long long)rint((double)phi / s1);
    q2 = (long long)rint(((double)third - o2) / s2);
    return true;
}

__global__ void __launch_bounds__(TPB) k_quantise_keys(const float* __restrict__ xyz, int stride,
                                                        const Tile* __restrict__ tiles, JobDev* jobs,
                                                        u64* __restrict__ keys, int mode) {
    __shared__ int s_slow[TILE];
    __shared__ int s_nslow;
    Tile t = tiles[blockIdx.x];
    JobDev& J = jobs[t.job];
    const float* p = xyz + (J.pt_begin + t.begin) * (long long)stride;
    u64* out = keys + J.key_begin + t.begin;
    const double s0 = J.step[0], s1 = J.step[1], s2 = J.step[2], o2 = J.off[2];
    const double i0 = J.inv_step[0], i1 = J.inv_step[1], i2 = J.inv_step[2];
    const double m0 = J.margin[0], m1 = J.margin[1], m2 = J.margin[2];
    const float coff = (float)J.cart_offset, cqs = (float)J.qs;
    u32 qmax = 0, ovf = 0;
    if (threadIdx.x == 0) s_nslow = 0;
    __syncthreads();
    auto emit = [&](int i, long long q0, long long q1, long long q2) {
        if (q0 < 0 || q1 < 0 || q2 < 0 || q0 >= (1 << 21) || q1 >= (1 << 21) || q2 >= (1 << 21)) {
            ovf = 1; q0 = q1 = q2 = 0;
        }
        qmax = max(qmax, (u32)max(q0, max(q1, q2)));
        out[i] = (spread3((u32)q0) << 2) | (spread3((u32)q1) << 1) | spread3((u32)q2);
    };
    for (int i = threadIdx.x; i < t.count; i += TPB) {
        float x = p[(long long)i * stride], y = p[(long long)i * stride + 1], z = p[(long long)i * stride + 2];
        long long q0, q1, q2;
        if (mode == SCP_MODE_CART) {
            // float32 chain: (p - offset) / qs with python-float scalars (NEP 50 keeps float32)
            q0 = (long long)rintf(__fdiv_rn(__fsub_rn(x, coff), cqs));
            q1 = (long long)rintf(__fdiv_rn(__fsub_rn(y, coff), cqs));
            q2 = (long long)rintf(__fdiv_rn(__fsub_rn(z, coff), cqs));
        } else {
            const float rho = rho_of(x, y, z, mode);
            const float xe = __fadd_rn(x, 1e-9f);
            float phi = atan2f(y, xe);
            if (phi < 0.f) phi = __fadd_rn(phi, 6.2831854820251465f);
            const float third = (mode == SCP_MODE_SPHER) ? acosf(__fdiv_rn(z, rho)) : z;
            const double u0 = (double)rho * i0, u1 = (double)phi * i1, u2 = ((double)third - o2) * i2;
            const double r0 = rint(u0), r1 = rint(u1), r2 = rint(u2);
            const bool safe = (0.5 - fabs(u0 - r0) > m0) && (0.5 - fabs(u1 - r1) > m1) && (0.5 - fabs(u2 - r2) > m2) &&
                              (phi > 1e-3f) && (phi < 6.28f);      // keep clear of the 0 / 2*pi fold
            if (!safe) { s_slow[atomicAdd(&s_nslow, 1)] = i; continue; }
            q0 = (long long)r0; q1 = (long long)r1; q2 = (long long)r2;
        }
        emit(i, q0, q1, q2);
    }
    __syncthreads();
    // exact float64 path for the few points near a rounding boundary, densely packed over the threads
    for (int e = threadIdx.x; e < s_nslow; e += TPB) {
        const int i = s_slow[e];
        float x = p[(long long)i * stride], y = p[(long long)i * stride + 1], z = p[(long long)i * stride + 2];
        long long q0, q1, q2;
        quantise_exact(x, y, z, mode, s0, s1, s2, o2, q0, q1, q2);
        emit(i, q0, q1, q2);
    }
    qmax = __reduce_max_sync(0xffffffffu, qmax);
    ovf = __reduce_max_sync(0xffffffffu, ovf);
    if ((threadIdx.x & 31) == 0) {
        if (qmax) atomicMax(&J.qmax, qmax);
        if (ovf) atomicMax(&J.overflow, 1u);
    }
}

// Same as k_quantise_keys for frames that carry several jobs (encode_mullevel: three sub-octrees with different steps
// over the SAME points): the float32 transform (sqrt, atan2, acos -- most of the instructions) is evaluated once per
// point, only the divide / round / Morton spread is per job.  One block per frame tile, up to QF_MAXJ jobs per frame.
constexpr int QF_MAXJ = 4;
__global__ void __launch_bounds__(TPB) k_quantise_frames(const float* __restrict__ xyz, int stride,
                                                          const Tile* __restrict__ ftiles, const long long* __restrict__ frame_begin,
                                                          const int* __restrict__ fj_start, const int* __restrict__ fj,
                                                          JobDev* jobs, u64* __restrict__ keys, int mode) {
    __shared__ int s_slow[TILE * QF_MAXJ];            // (job slot << 16) | point
    __shared__ int s_nslow;
    __shared__ long long s_kb[QF_MAXJ];
    const Tile t = ftiles[blockIdx.x];
    const int f = t.job;
    const int j0 = fj_start[f], nj = fj_start[f + 1] - j0;
    const float* p = xyz + (frame_begin[f] + t.begin) * (long long)stride;
    u32 qmax[QF_MAXJ], ovf[QF_MAXJ];
#pragma unroll
    for (int s = 0; s < QF_MAXJ; ++s) { qmax[s] = 0; ovf[s] = 0; }
    if (threadIdx.x == 0) s_nslow = 0;
    if (threadIdx.x < nj) {
        s_kb[threadIdx.x] = jobs[fj[j0 + threadIdx.x]].key_begin;
    }
    __syncthreads();
    auto emit = [&](int slot, int i, long long q0l, long long q1l, long long q2l) {
        u32 o = 0;
        // one unsigned test covers negative and too large values (they convert to >= 2^21 as unsigned 64-bit)
        if ((((unsigned long long)q0l | (unsigned long long)q1l | (unsigned long long)q2l) >> 21) != 0ull) { o = 1; q0l = q1l = q2l = 0; }
        const u32 q0 = (u32)q0l, q1 = (u32)q1l, q2 = (u32)q2l;
        const u32 qm = max(q0, max(q1, q2));
#pragma unroll
        for (int s = 0; s < QF_MAXJ; ++s) if (s == slot) { qmax[s] = max(qmax[s], qm); ovf[s] |= o; }
        keys[s_kb[slot] + t.begin + i] = (spread3(q0) << 2) | (spread3(q1) << 1) | spread3(q2);
    };
    for (int i = threadIdx.x; i < t.count; i += TPB) {
        if (i + 2 * TPB < t.count) prefetch_l2(p + (long long)(i + 2 * TPB) * stride);       // two rounds ahead (the loads below
                                                                                          // were 20 % of the stall samples)
        const float x = p[(long long)i * stride], y = p[(long long)i * stride + 1], z = p[(long long)i * stride + 2];
        const float rho = rho_of(x, y, z, mode);
        const float xe = __fadd_rn(x, 1e-9f);
        float phi = atan2f(y, xe);
        if (phi < 0.f) phi = __fadd_rn(phi, 6.2831854820251465f);
        const float third = (mode == SCP_MODE_SPHER) ? acosf(__fdiv_rn(z, rho)) : z;
        const bool clear = (phi > 1e-3f) && (phi < 6.28f);                 // keep clear of the 0 / 2*pi fold
#pragma unroll
        for (int slot = 0; slot < QF_MAXJ; ++slot) {
            if (slot >= nj) break;
            const JobDev& J = jobs[fj[j0 + slot]];                 // (L1-resident; a shared-memory copy cost 40 registers)
            const double u0 = (double)rho * J.inv_step[0], u1 = (double)phi * J.inv_step[1], u2 = ((double)third - J.off[2]) * J.inv_step[2];
            const double r0 = rint(u0), r1 = rint(u1), r2 = rint(u2);
            const bool safe = clear && (0.5 - fabs(u0 - r0) > J.margin[0]) && (0.5 - fabs(u1 - r1) > J.margin[1]) &&
                              (0.5 - fabs(u2 - r2) > J.margin[2]);
            if (!safe) { s_slow[atomicAdd(&s_nslow, 1)] = (slot << 16) | i; continue; }
            emit(slot, i, (long long)r0, (long long)r1, (long long)r2);
        }
    }
    __syncthreads();
    // exact float64 path for the few (point, job) pairs near a rounding boundary, densely packed over the threads
    for (int e = threadIdx.x; e < s_nslow; e += TPB) {
        const int slot = s_slow[e] >> 16, i = s_slow[e] & 0xffff;
        const JobDev& J = jobs[fj[j0 + slot]];
        const float x = p[(long long)i * stride], y = p[(long long)i * stride + 1], z = p[(long long)i * stride + 2];
        long long q0, q1, q2;
        quantise_exact(x, y, z, mode, J.step[0], J.step[1], J.step[2], J.off[2], q0, q1, q2);
        emit(slot, i, q0, q1, q2);
    }
#pragma unroll
    for (int s = 0; s < QF_MAXJ; ++s) {
        if (s >= nj) break;
        const u32 qm = __reduce_max_sync(0xffffffffu, qmax[s]), ov = __reduce_max_sync(0xffffffffu, ovf[s]);
        if ((threadIdx.x & 31) == 0) {
            JobDev& J = jobs[fj[j0 + s]];
            if (qm) atomicMax(&J.qmax, qm);
            if (ov) atomicMax(&J.overflow, 1u);
        }
    }
}

// Fused coordinate transform + quantise + morton_path filter + compaction (all jobs of a frame in one pass over its points).
// Needs every job's depth up front (JobDev::depth_known, see k_job_setup).  Per point the float32 transform is evaluated once;
// per job the radial coordinate is quantised first and the morton_path test (Octree.py:188: bits depth-1-j of the radial
// coordinate equal path[j]) rejects the point before the angular coordinates, the rounding-margin tests and the Morton spread
// are touched -- the second and third sub-octree of encode_mullevel keep 5 % and 0.4 % of a sweep.  Kept keys are staged in
// shared memory and appended to the job's key array at an offset reserved with ONE atomicAdd per (block, job): their order in
// front of the sort is irrelevant (equal keys are indistinguishable), so no ordered scan over the tiles is needed and the
// separate filter-count / compaction kernels (two more reads of every key) disappear.
// `par_pmax`: number of digit passes of the longest job; a job with fewer passes starts in the other ping-pong buffer so that
// all jobs end in the same one (see k_onesweep).
__device__ __forceinline__ int sort_passes(int depth) { return (3 * depth + 7) >> 3; }

struct QJob {                 // per-job constants of one block (shared memory)
    double inv0, inv1, inv2, off2, m0, m1, m2;
    double s0, s1, s2;
    long long key_begin;
    int fshift, fval, plen, job;
    int to_b, _pad;
};

constexpr int QHALF = TILE / 2;            // points per flush of the staged keys
__global__ void __launch_bounds__(TPB, 4) k_quantise_fused(const float* __restrict__ xyz, int stride,
                                                            const Tile* __restrict__ ftiles, const long long* __restrict__ frame_begin,
                                                            const int* __restrict__ fj_start, const int* __restrict__ fj,
                                                            JobDev* jobs, u64* __restrict__ keys_a, u64* __restrict__ keys_b,
                                                            int mode, int pmax) {
    __shared__ u64 s_keys[QF_MAXJ][QHALF];                              // kept keys of the current half tile, per job
    __shared__ QJob s_j[QF_MAXJ];
    __shared__ unsigned short s_slow[QHALF * QF_MAXJ];                  // (job slot << 12) | point inside the tile
    __shared__ int s_nslow;
    __shared__ u32 s_cnt[QF_MAXJ], s_base[QF_MAXJ];
    const Tile t = ftiles[blockIdx.x];
    const int f = t.job;
    const int j0 = fj_start[f], nj = fj_start[f + 1] - j0;
    const float* p = xyz + (frame_begin[f] + t.begin) * (long long)stride;
    const bool v4 = stride == 4 && (reinterpret_cast<uintptr_t>(p) & 15) == 0;      // KITTI rows (x, y, z, intensity): one 16-byte load
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < nj) {
        const int jid = fj[j0 + threadIdx.x];
        const JobDev& J = jobs[jid];
        QJob q;
        q.inv0 = J.inv_step[0]; q.inv1 = J.inv_step[1]; q.inv2 = J.inv_step[2]; q.off2 = J.off[2];
        q.m0 = J.margin[0]; q.m1 = J.margin[1]; q.m2 = J.margin[2];
        q.s0 = J.step[0]; q.s1 = J.step[1]; q.s2 = J.step[2];
        q.key_begin = J.key_begin; q.job = jid;
        // Octree.py:188: bit depth-1-j of the radial coordinate == path[j]; bits below bit 0 read as 0
        const int cmp = min(J.path_len, J.depth);
        int v = 0, tail_ok = 1;
        for (int j = 0; j < J.path_len; ++j) {
            const int b = (J.path_bits >> j) & 1;
            if (j < cmp) v = (v << 1) | b; else tail_ok &= (b == 0);
        }
        q.plen = tail_ok ? cmp : -1;
        q.fshift = J.depth - cmp;
        q.fval = v;
        q.to_b = (pmax - sort_passes(J.depth)) & 1;
        q._pad = 0;
        s_j[threadIdx.x] = q;
    }
    u32 qmax[QF_MAXJ], ovf[QF_MAXJ];
#pragma unroll
    for (int s = 0; s < QF_MAXJ; ++s) { qmax[s] = 0; ovf[s] = 0; }
    auto passes = [&](const QJob& q, long long q0) -> bool {          // radial filter on the exactly rounded coordinate
        if (q.plen == 0) return true;
        if (q.plen < 0) return false;
        return (q0 >> q.fshift) == (long long)q.fval;
    };
    // stages the kept keys of this warp: one shared-memory atomicAdd per (warp, job) reserves the run
    auto emit = [&](int slot, bool keep, long long q0l, long long q1l, long long q2l) {
        const u32 peers = __ballot_sync(0xffffffffu, keep);
        if (peers == 0) return;                                        // warp-uniform
        const int leader = __ffs(peers) - 1;
        u32 base = 0;
        if (lane == leader) base = atomicAdd(&s_cnt[slot], (u32)__popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (!keep) return;
        u32 o = 0;
        if ((((unsigned long long)q0l | (unsigned long long)q1l | (unsigned long long)q2l) >> 21) != 0ull) { o = 1; q0l = q1l = q2l = 0; }
        const u32 q0 = (u32)q0l, q1 = (u32)q1l, q2 = (u32)q2l;
        s_keys[slot][base + __popc(peers & ((1u << lane) - 1u))] = (spread3(q0) << 2) | (spread3(q1) << 1) | spread3(q2);
        const u32 qm = max(q0, max(q1, q2));
#pragma unroll
        for (int s = 0; s < QF_MAXJ; ++s) if (s == slot) { qmax[s] = max(qmax[s], qm); ovf[s] |= o; }
    };
    for (int h0 = 0; h0 < t.count; h0 += QHALF) {
        const int hend = min(t.count, h0 + QHALF);
        if (threadIdx.x < QF_MAXJ) s_cnt[threadIdx.x] = 0;
        if (threadIdx.x == 0) s_nslow = 0;
        __syncthreads();
        for (int i0 = h0; i0 < hend; i0 += TPB) {
            const int i = i0 + threadIdx.x;
            const bool valid = i < hend;
            float x = 0.f, y = 0.f, z = 1.f;
            if (valid) {
                if (i + 2 * TPB < t.count) prefetch_l2(p + (long long)(i + 2 * TPB) * stride);
                if (v4) { const float4 v = __ldg(reinterpret_cast<const float4*>(p) + i); x = v.x; y = v.y; z = v.z; }
                else { x = p[(long long)i * stride]; y = p[(long long)i * stride + 1]; z = p[(long long)i * stride + 2]; }
            }
            const float rho = rho_of(x, y, z, mode);
            const double drho = (double)rho;
            u32 want = 0;                                               // jobs that keep this point (radial test only)
            long long q0s[QF_MAXJ];
#pragma unroll
            for (int slot = 0; slot < QF_MAXJ; ++slot) {
                q0s[slot] = 0;
                if (slot >= nj || !valid) continue;
                const QJob& q = s_j[slot];
                const double u0 = drho * q.inv0, r0 = rint(u0);
                if (0.5 - fabs(u0 - r0) > q.m0) {
                    q0s[slot] = (long long)r0;
                    if (passes(q, q0s[slot])) want |= 1u << slot;
                } else {
                    s_slow[atomicAdd(&s_nslow, 1)] = (unsigned short)((slot << 12) | i);   // within the margin of a .5 boundary
                }
            }
            float phi = 0.f, third = 0.f;
            bool clear = false;
            if (want) {                                                 // the expensive part only for points somebody keeps
                const float xe = __fadd_rn(x, 1e-9f);
                phi = atan2f(y, xe);
                if (phi < 0.f) phi = __fadd_rn(phi, 6.2831854820251465f);
                third = (mode == SCP_MODE_SPHER) ? acosf(__fdiv_rn(z, rho)) : z;
                clear = (phi > 1e-3f) && (phi < 6.28f);                 // keep clear of the 0 / 2*pi fold
            }
#pragma unroll
            for (int slot = 0; slot < QF_MAXJ; ++slot) {
                if (slot >= nj) break;
                const QJob& q = s_j[slot];
                bool keep = false;
                long long q1 = 0, q2 = 0;
                if ((want >> slot) & 1u) {
                    const double u1 = (double)phi * q.inv1, u2 = ((double)third - q.off2) * q.inv2;
                    const double r1 = rint(u1), r2 = rint(u2);
                    if (clear && (0.5 - fabs(u1 - r1) > q.m1) && (0.5 - fabs(u2 - r2) > q.m2)) { keep = true; q1 = (long long)r1; q2 = (long long)r2; }
                    else s_slow[atomicAdd(&s_nslow, 1)] = (unsigned short)((slot << 12) | i);
                }
                emit(slot, keep, q0s[slot], q1, q2);
            }
        }
        __syncthreads();
        // exact float64 path for the (point, job) pairs near a rounding boundary, densely packed over the threads
        const int nslow = s_nslow;
        for (int e0 = 0; e0 < nslow; e0 += TPB) {
            const int e = e0 + threadIdx.x;
            bool keep = false;
            int slot = -1;
            long long q0 = 0, q1 = 0, q2 = 0;
            if (e < nslow) {
                slot = s_slow[e] >> 12;
                const int i = s_slow[e] & 0xfff;
                const QJob& q = s_j[slot];
                const float x = p[(long long)i * stride], y = p[(long long)i * stride + 1], z = p[(long long)i * stride + 2];
                quantise_exact(x, y, z, mode, q.s0, q.s1, q.s2, q.off2, q0, q1, q2);
                keep = passes(q, q0);
            }
#pragma unroll
            for (int sl = 0; sl < QF_MAXJ; ++sl) {
                if (sl >= nj) break;
                emit(sl, keep && slot == sl, q0, q1, q2);
            }
        }
        __syncthreads();
        // one atomicAdd per (block, job, half) reserves the run in the job's key array; coalesced copy out
        if (threadIdx.x < nj) s_base[threadIdx.x] = s_cnt[threadIdx.x] ? atomicAdd((u32*)&jobs[s_j[threadIdx.x].job].n_kept, s_cnt[threadIdx.x]) : 0u;
        __syncthreads();
        for (int slot = 0; slot < nj; ++slot) {
            const QJob& q = s_j[slot];
            u64* dst = (q.to_b ? keys_b : keys_a) + q.key_begin + s_base[slot];
            const int c = (int)s_cnt[slot];
            for (int i = threadIdx.x; i < c; i += TPB) dst[i] = s_keys[slot][i];
        }
        __syncthreads();
    }
#pragma unroll
    for (int s = 0; s < QF_MAXJ; ++s) {
        if (s >= nj) break;
        const u32 qm = __reduce_max_sync(0xffffffffu, qmax[s]), ov = __reduce_max_sync(0xffffffffu, ovf[s]);
        if (lane == 0) {
            JobDev& J = jobs[s_j[s].job];
            if (qm) atomicMax(&J.qmax, qm);
            if (ov) atomicMax(&J.overflow, 1u);
        }
    }
}

__global__ void k_job_prepare_fused(JobDev* jobs, int n_jobs) {      // the kept-key counters start at zero
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n_jobs) jobs[j].n_kept = 0;
}

__global__ void k_job_depth(JobDev* jobs, int n_jobs) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_jobs) return;
    u32 m = jobs[j].qmax;
    jobs[j].depth = m ? 32 - __clz(m) : 0;     // ceil(log2(max+1)), Octree.py:58
}

// ------------------------------------------------------------------------------------------
// K3: segmented radix sort (Onesweep: global digit histograms once, then P scatter passes whose
// tile offsets come from a decoupled look-back over single-word (flag|count) descriptors)
// ------------------------------------------------------------------------------------------
constexpr u32 FLAG_AGG = 1u << 30, FLAG_INC = 1u << 31, VAL_MASK = (1u << 30) - 1;

// Digit histograms of every job (the "1" of the sort's (1 + 2P) * 8 B/key).  pmax > 0: per-job pass schedule -- a job whose keys
// need fewer digit passes than the longest job starts in the other ping-pong buffer (see k_onesweep).
__global__ void __launch_bounds__(TPB) k_sort_hist(const u64* __restrict__ keys_a, const u64* __restrict__ keys_b,
                                                    const Tile* __restrict__ tiles, const JobDev* __restrict__ jobs,
                                                    u32* __restrict__ hist, int P, int pmax) {
    __shared__ u32 sh[8 * 256];
    Tile t = tiles[blockIdx.x];
    const JobDev& J = jobs[t.job];
    const int cnt = min(t.count, J.n_kept - t.begin);
    if (cnt <= 0) return;                                  // block-uniform
    const int pj = t.pj > 0 ? t.pj : P;
    for (int i = threadIdx.x; i < pj * 256; i += TPB) sh[i] = 0;
    __syncthreads();
    const u64* src = ((t.pj > 0 && ((pmax - pj) & 1)) ? keys_b : keys_a) + t.kb + t.begin;
    for (int i = threadIdx.x; i < cnt; i += TPB) {
        const u64 k = src[i];
        for (int p = 0; p < pj; ++p) atomicAdd(&sh[p * 256 + (u32)((k >> (8 * p)) & 0xff)], 1u);
    }
    __syncthreads();
    u32* h = hist + (size_t)t.job * P * 256;
    for (int i = threadIdx.x; i < pj * 256; i += TPB)
        if (sh[i]) atomicAdd(&h[i], sh[i]);
}

// morton_path filter of Octree.py:188 as a stream compaction in front of the sort (mullevel: each of the three jobs of a
// frame keeps a third of the points on average, so sorting the rejected keys as sentinels tripled the sort's traffic):
//   k_filter_count  kept keys per sort tile            k_kept_scan   per-job exclusive scan, n_kept
//   k_compact_hist  kept keys -> dense prefix of `out` (order inside a tile preserved) + the digit histograms of the sort
__device__ __forceinline__ bool path_keep(u64 k, int n, int plen, int pbits) {
    bool keep = true;
    for (int j = 0; j < plen; ++j) {
        const int b = n - 1 - j;
        const int bit = b >= 0 ? (int)((k >> (3 * b + 2)) & 1) : 0;
        keep &= (bit == ((pbits >> j) & 1));
    }
    return keep;
}

__global__ void __launch_bounds__(TPB) k_filter_count(const u64* __restrict__ keys, const Tile* __restrict__ tiles,
                                                       const JobDev* __restrict__ jobs, u32* __restrict__ tile_kept) {
    __shared__ u32 s_cnt;
    const Tile t = tiles[blockIdx.x];
    const JobDev& J = jobs[t.job];
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const u64* src = keys + J.key_begin + t.begin;
    u32 c = 0;
    for (int i = threadIdx.x; i < t.count; i += TPB) c += path_keep(src[i], J.depth, J.path_len, J.path_bits) ? 1u : 0u;
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_cnt, c);
    __syncthreads();
    if (threadIdx.x == 0) tile_kept[blockIdx.x] = s_cnt;
}

__global__ void k_kept_scan(JobDev* jobs, const int* __restrict__ job_tile_begin, u32* tile_kept /* in: counts, out: offsets */) {
    if (threadIdx.x) return;
    JobDev& J = jobs[blockIdx.x];
    u32 run = 0;
    for (int t = job_tile_begin[blockIdx.x]; t < job_tile_begin[blockIdx.x + 1]; ++t) { const u32 c = tile_kept[t]; tile_kept[t] = run; run += c; }
    J.n_kept = (int)run;
}

__global__ void __launch_bounds__(TPB) k_compact_hist(const u64* __restrict__ in, u64* __restrict__ out,
                                                       const Tile* __restrict__ tiles, const JobDev* __restrict__ jobs,
                                                       const u32* __restrict__ tile_off, u32* __restrict__ hist, int P) {
    __shared__ u32 sh[8 * 256];
    __shared__ u32 s_wbase[8];
    const Tile t = tiles[blockIdx.x];
    const JobDev& J = jobs[t.job];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < P * 256; i += TPB) sh[i] = 0;
    const u64* src = in + J.key_begin + t.begin;
    u64* dst = out + J.key_begin + tile_off[blockIdx.x];
    u32 base = 0;                                        // kept keys of the tile written so far
    for (int i0 = 0; i0 < t.count; i0 += TPB) {
        __syncthreads();
        const int i = i0 + threadIdx.x;
        u64 k = 0;
        bool keep = false;
        if (i < t.count) { k = src[i]; keep = path_keep(k, J.depth, J.path_len, J.path_bits); }
        const u32 bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_wbase[warp] = __popc(bal);
        __syncthreads();
        u32 wb = 0, tot = 0;
        for (int w = 0; w < 8; ++w) { const u32 c = s_wbase[w]; if (w < warp) wb += c; tot += c; }
        if (keep) {
            dst[base + wb + __popc(bal & ((1u << lane) - 1u))] = k;
            for (int p = 0; p < P; ++p) atomicAdd(&sh[p * 256 + (u32)((k >> (8 * p)) & 0xff)], 1u);
        }
        base += tot;
    }
    __syncthreads();
    u32* h = hist + (size_t)t.job * P * 256;
    for (int i = threadIdx.x; i < P * 256; i += TPB)
        if (sh[i]) atomicAdd(&h[i], sh[i]);
}

__global__ void __launch_bounds__(256) k_scan_hist(u32* hist) {   // one block per (job, pass): exclusive scan of 256 bins
    __shared__ u32 ws[8];
    u32* h = hist + (size_t)blockIdx.x * 256;
    u32 v = h[threadIdx.x];
    u32 inc = v;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { u32 n = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += n; }
    if (lane == 31) ws[w] = inc;
    __syncthreads();
    u32 base = 0;
    for (int i = 0; i < w; ++i) base += ws[i];
    h[threadIdx.x] = base + inc - v;
}

// Tiles with pj > 0 follow the per-job pass schedule (see below; result in (pmax & 1) ? B : A), tiles with pj == 0 take all P
// passes starting from A (scp_segmented_sort_u64, fallback path).
__global__ void __launch_bounds__(TPB, 4) k_onesweep(u64* __restrict__ buf_a, u64* __restrict__ buf_b,
                                                   const Tile* __restrict__ tiles, int n_tiles,
                                                   const JobDev* __restrict__ jobs, const u32* __restrict__ hist,
                                                   int P, int pass, int pmax, u32* desc, u32* ticket, u32* err) {
    __shared__ u64 s_keys[SORT_TILE];
    __shared__ u32 s_whist[8][257];
    __shared__ u32 s_dstart[256];
    __shared__ u32 s_gbase[256];
    __shared__ u32 s_ws[8];
    __shared__ int s_tile;
    if (threadIdx.x == 0) s_tile = (int)atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < 8 * 257; i += TPB) (&s_whist[0][0])[i] = 0;
    __syncthreads();
    const int t = s_tile;
    if (t >= n_tiles) return;
    Tile tl = tiles[t];
    const JobDev& J = jobs[tl.job];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int shift = pass * 8;
    const int wbase = warp * (32 * SORT_ITEMS);
    // per-job schedule (tl.pj > 0): a job of depth d needs pj = ceil(3d/8) passes; it skips the rest and starts in buffer B when
    // (pmax - pj) is odd, so that after pmax launches every job's sorted keys lie in the same buffer
    if (tl.pj > 0 && pass >= tl.pj) return;              // block-uniform: this job's keys are sorted already
    const int par = tl.pj > 0 ? ((pass + pmax - tl.pj) & 1) : (pass & 1);
    // The keys are addressed from the tile record alone (ticket -> tile -> keys): going through the job record for the key
    // offset (ticket -> tile -> job -> keys) put a fourth dependent L2 / DRAM latency at the head of every tile.  n_kept (a
    // device-side count) arrives in parallel; slots behind it hold stale keys of the job's own region and become SENTINEL.
    const int n_kept = J.n_kept;
    u64 key[SORT_ITEMS];
    {
        const u64* sa = (par ? buf_b : buf_a) + tl.kb + tl.begin;
#pragma unroll
        for (int i = 0; i < SORT_ITEMS; ++i) {
            const int idx = wbase + i * 32 + lane;
            key[i] = idx < tl.count ? sa[idx] : SENTINEL;
        }
    }
    tl.count = min(tl.count, n_kept - tl.begin);         // compacted jobs: the tail tiles are empty and nobody looks back at them
    if (tl.count <= 0) return;
    u64* out = par ? buf_a : buf_b;
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) if (wbase + i * 32 + lane >= tl.count) key[i] = SENTINEL;
    u32 rank[SORT_ITEMS];
    // the 16 match operations are independent of the histogram chain below: issue them back to back (each has ~40 cycles
    // of latency; inside the read-modify-write loop they were the largest stall of the kernel).  (Replacing the chain by one
    // shared-memory atomicAdd per digit group, whose return value is the count of the warp's earlier items, measured slower:
    // 2.16 against 1.63 ms for the seven passes.)
    u32 peers[SORT_ITEMS];
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        int idx = wbase + i * 32 + lane;
        u32 d = idx < tl.count ? (u32)((key[i] >> shift) & 0xff) : 256u;
        peers[i] = __match_any_sync(0xffffffffu, d);
    }
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        int idx = wbase + i * 32 + lane;
        u32 d = idx < tl.count ? (u32)((key[i] >> shift) & 0xff) : 256u;
        u32 lt = peers[i] & ((1u << lane) - 1u);
        u32 prev = s_whist[warp][d];
        __syncwarp();
        if (lt == 0) s_whist[warp][d] = prev + __popc(peers[i]);
        __syncwarp();
        rank[i] = prev + __popc(lt);
    }
    __syncthreads();
    const u32 d = threadIdx.x;
    u32 total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { u32 c = s_whist[w][d]; s_whist[w][d] = total; total += c; }
    volatile u32* vdesc = desc + ((size_t)pass * n_tiles) * 256;
    vdesc[(size_t)t * 256 + d] = total | FLAG_AGG;
    // exclusive scan of `total` over the 256 digits -> position of the digit's run inside the tile
    u32 inc = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { u32 n = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += n; }
    if (lane == 31) s_ws[warp] = inc;
    __syncthreads();
    u32 wb = 0;
    for (int i = 0; i < warp; ++i) wb += s_ws[i];
    s_dstart[d] = wb + inc - total;
    // decoupled look-back over the previous tiles of the same job
    u32 ex = 0;
    for (int p = t - 1; p >= tl.first; --p) {
        u32 v;
        int spins = 0;
        do { v = vdesc[(size_t)p * 256 + d]; } while ((v & (FLAG_AGG | FLAG_INC)) == 0 && ++spins < (1 << 22));
        if ((v & (FLAG_AGG | FLAG_INC)) == 0) { atomicExch(err, 1u); break; }   // never hang the GPU
        ex += v & VAL_MASK;
        if (v & FLAG_INC) break;
    }
    vdesc[(size_t)t * 256 + d] = ((ex + total) & VAL_MASK) | FLAG_INC;
    s_gbase[d] = hist[((size_t)tl.job * P + pass) * 256 + d] + ex;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        int idx = wbase + i * 32 + lane;
        if (idx < tl.count) {
            u32 dg = (u32)((key[i] >> shift) & 0xff);
            s_keys[s_dstart[dg] + s_whist[warp][dg] + rank[i]] = key[i];
        }
    }
    __syncthreads();
    u64* dst = out + tl.kb;
    for (int s = threadIdx.x; s < tl.count; s += TPB) {
        u64 k = s_keys[s];
        u32 dg = (u32)((k >> shift) & 0xff);
        dst[s_gbase[dg] + ((u32)s - s_dstart[dg])] = k;
    }
}

// ------------------------------------------------------------------------------------------
// K4: head levels + per-tile histograms
// ------------------------------------------------------------------------------------------
// h = 0: not a new voxel (duplicate / filtered); h = 1: first voxel of the job; otherwise d+2 where d is the
// number of leading octal digits (out of n) shared with the previous key.  h <= n+1.
__device__ __forceinline__ int head_level(u64 k, u64 prev, bool first, int n) {
    if (k == SENTINEL) return 0;
    if (first) return 1;
    if (k == prev) return 0;
    int hb = 63 - __clzll((long long)(k ^ prev));
    return n - hb / 3 + 1;      // d = n-1-hb/3
}

// A warp owns WKEYS consecutive sorted keys of the tile, 32 at a time: lane l holds key 32*it + l of the warp's run, its
// predecessor comes from lane l-1 by shuffle (one extra global load per warp, for the key in front of the run).
constexpr int WITER = TILE / TPB;            // 8 rounds of 32 keys per warp
constexpr int WKEYS = 32 * WITER;            // 256
__device__ __forceinline__ void load_heads(const u64* __restrict__ src, int tbegin, int cnt, int warp, int lane, int n,
                                           u64 (&k)[WITER], int (&h)[WITER]) {
    const int wbeg = warp * WKEYS;
#pragma unroll
    for (int it = 0; it < WITER; ++it) {
        const int idx = wbeg + it * 32 + lane;
        k[it] = idx < cnt ? __ldg(src + tbegin + idx) : SENTINEL;
    }
    const int g0 = tbegin + wbeg;
    u64 carry = (lane == 0 && g0 > 0 && wbeg < cnt) ? __ldg(src + g0 - 1) : 0ull;
#pragma unroll
    for (int it = 0; it < WITER; ++it) {
        u64 prev = __shfl_up_sync(0xffffffffu, k[it], 1);
        if (lane == 0) prev = carry;
        h[it] = head_level(k[it], prev, g0 + it * 32 + lane == 0, n);
        carry = __shfl_sync(0xffffffffu, k[it], 31);
    }
}

// One histogram per CHUNK = the WKEYS (256) consecutive sorted keys a warp owns (8 chunks per tile): with the exclusive prefix
// over a job's chunks (k_level_scan) every warp of the later passes knows the BFS rank of its first node on every level
// without any block-level scan.
constexpr int CHUNKS = TILE / WKEYS;         // 8
__global__ void __launch_bounds__(TPB) k_head_hist(const u64* __restrict__ keys, const Tile* __restrict__ tiles,
                                                    const JobDev* __restrict__ jobs, u32* __restrict__ chunk_hist) {
    __shared__ u32 sh[CHUNKS][NBINS];
    const Tile t = tiles[blockIdx.x];
    const JobDev& J = jobs[t.job];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane < NBINS) sh[warp][lane] = 0;
    __syncwarp();
    // keys past n_kept were filtered out before the sort (tail tiles are empty)
    const int cnt = min(t.count, J.n_kept - t.begin);
    if (cnt > 0) {
        u64 k[WITER];
        int h[WITER];
        load_heads(keys + J.key_begin, t.begin, cnt, warp, lane, J.depth, k, h);
#pragma unroll
        for (int it = 0; it < WITER; ++it) {
            const u32 peers = __match_any_sync(0xffffffffu, h[it]);
            if (h[it] > 0 && (peers & ((1u << lane) - 1u)) == 0) atomicAdd(&sh[warp][h[it]], (u32)__popc(peers));
        }
    }
    __syncwarp();
    if (lane < NBINS) chunk_hist[((size_t)blockIdx.x * CHUNKS + warp) * NBINS + lane] = sh[warp][lane];
}

// K5: one warp per job: exclusive scan of the tile histograms, level counts and offsets
__global__ void k_level_scan(JobDev* jobs, int n_jobs, const int* __restrict__ job_tile_begin,
                             u32* chunk_hist /* in: counts, out: exclusive prefix inside the job */) {
    int j = blockIdx.x;
    JobDev& J = jobs[j];
    __shared__ u32 tot[32];
    int b = threadIdx.x;
    u32 run = 0;
    if (b < NBINS) {
        // only the chunks that hold sorted keys (compacted jobs keep a fraction of their points)
        const int t0 = job_tile_begin[j] * CHUNKS;
        const int t1 = min(job_tile_begin[j + 1] * CHUNKS, t0 + (J.n_kept + WKEYS - 1) / WKEYS);
        int t = t0;
        for (; t + 8 <= t1; t += 8) {                   // eight independent loads in flight
            u32 c[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) c[q] = chunk_hist[(size_t)(t + q) * NBINS + b];
#pragma unroll
            for (int q = 0; q < 8; ++q) { chunk_hist[(size_t)(t + q) * NBINS + b] = run; run += c[q]; }
        }
        for (; t < t1; ++t) {
            u32 c = chunk_hist[(size_t)t * NBINS + b];
            chunk_hist[(size_t)t * NBINS + b] = run;
            run += c;
        }
    }
    tot[b] = run;
    __syncwarp();
    if (b == 0) {
        int n = J.depth;
        int nodes = 0, cum = 0, vox = 0;
        for (int h = 1; h < NBINS; ++h) vox += tot[h];
        for (int L = 1; L <= MAXL; ++L) {
            if (L <= n) cum += tot[L];
            int c = (L <= n) ? cum : 0;
            J.level_start[L - 1] = nodes;
            J.level_count[L - 1] = c;
            nodes += c;
        }
        J.level_start[MAXL] = nodes;
        J.n_voxels = vox;
        J.n_nodes = nodes;
        J.n_rows = (J.drop_last && nodes > 0) ? nodes - 1 : nodes;
    }
}

// exclusive scan of (n_nodes, n_rows, n_voxels) over the jobs: one block, each thread a contiguous run of jobs
__global__ void __launch_bounds__(1024) k_job_offsets(JobDev* jobs, int n_jobs) {
    __shared__ long long s[3][1024];
    const int per = (n_jobs + 1023) / 1024;
    const int j0 = min(n_jobs, (int)threadIdx.x * per), j1 = min(n_jobs, j0 + per);
    long long a[3] = {0, 0, 0};
    for (int j = j0; j < j1; ++j) { a[0] += jobs[j].n_nodes; a[1] += jobs[j].n_rows; a[2] += jobs[j].n_voxels; }
#pragma unroll
    for (int c = 0; c < 3; ++c) s[c][threadIdx.x] = a[c];
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        long long v[3] = {0, 0, 0};
        if ((int)threadIdx.x >= o)
            for (int c = 0; c < 3; ++c) v[c] = s[c][threadIdx.x - o];
        __syncthreads();
        for (int c = 0; c < 3; ++c) s[c][threadIdx.x] += v[c];
        __syncthreads();
    }
    long long node = s[0][threadIdx.x] - a[0], row = s[1][threadIdx.x] - a[1], vox = s[2][threadIdx.x] - a[2];
    for (int j = j0; j < j1; ++j) {
        jobs[j].node_start = node; jobs[j].row_start = row; jobs[j].vox_start = vox;
        node += jobs[j].n_nodes; row += jobs[j].n_rows; vox += jobs[j].n_voxels;
    }
}

// ------------------------------------------------------------------------------------------
// K6: node records for all levels in one pass
// ------------------------------------------------------------------------------------------
// Internal node records, structure of arrays (BFS order inside a job): 19 B/node + 1 B/voxel
struct NodeArrays {
    uint16_t* lo;               // level | octant << 8
    uint8_t* occ;
    u32* parent;
    u64* pos;                   // cell origin, 21 bits per axis: x | y << 21 | z << 42
    u32* fc;                    // first child (BFS index inside the job; voxel index for the deepest level)   [k_emit_nodes path]
    uint8_t* vdig;              // per voxel: child digit + 1 inside its level-n node (the "octant" of the voxel) [k_emit_nodes path]
    int morton;                 // 1: `pos` holds the node's Morton key masked to its level (k_level_pass path)
};
__device__ __forceinline__ u64 pack_pos(u32 x, u32 y, u32 z) { return (u64)x | ((u64)y << 21) | ((u64)z << 42); }
__device__ __forceinline__ void unpack_pos(u64 p, int morton, u32& x, u32& y, u32& z) {
    if (morton) { x = compact3(p >> 2); y = compact3(p >> 1); z = compact3(p); }
    else { x = (u32)p & 0x1fffffu; y = (u32)(p >> 21) & 0x1fffffu; z = (u32)(p >> 42); }
}

// One block iteration handles 256 consecutive sorted keys: per level one ballot per warp gives the rank of every head inside
// the warp, a 22-entry scan over the 8 warps gives the rank inside the block.  Only levels >= the smallest head level of
// the warp are visited (consecutive sorted keys share long prefixes: 3-4 levels out of up to 21).
__global__ void __launch_bounds__(TPB) k_emit_nodes(const u64* __restrict__ keys, const Tile* __restrict__ tiles,
                                                     JobDev* jobs, const u32* __restrict__ tile_base,
                                                     NodeArrays A, u64* __restrict__ vox_key) {
    __shared__ u32 s_base[MAXL + 2];        // running count of heads per level (index L), [0] = voxels
    __shared__ u32 s_wcnt[8][MAXL + 2];
    __shared__ u32 s_wpre[8][MAXL + 2];
    __shared__ u32 s_mm[4];                 // min / max coordinate over the tile's voxels: all, and all but the job's last voxel
    __shared__ u64 s_keep[MAXL + 2];        // packed-position bits that survive on level L (cell size 2^(n-L+1))
    const Tile t = tiles[blockIdx.x];
    JobDev& J = jobs[t.job];
    const int n = J.depth;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 ltmask = (1u << lane) - 1u;
    if (threadIdx.x >= 32 && threadIdx.x <= 32 + MAXL) {
        const int L = threadIdx.x - 32;
        const u32 m = (L >= 1 && L <= n) ? (~((1u << (n - L + 1)) - 1u)) & 0x1fffffu : 0u;
        s_keep[L] = pack_pos(m, m, m);
    }
    if (threadIdx.x <= MAXL) {
        int L = threadIdx.x;       // L = 0: voxel counter
        u32 s = 0;
        const u32* tb = tile_base + (size_t)(t.first + t.begin / TILE) * CHUNKS * NBINS;   // first chunk of the tile (tiles may be a compacted list)
        if (L == 0) { for (int h = 1; h < NBINS; ++h) s += tb[h]; }
        else if (L <= n) { for (int h = 1; h <= L; ++h) s += tb[h]; }
        s_base[L] = s;
    }
    if (threadIdx.x < 4) s_mm[threadIdx.x] = (threadIdx.x & 1) ? 0u : 0xffffffffu;
    __syncthreads();
    const u64* src = keys + J.key_begin;
    const long long node0 = J.node_start;
    const int n_vox = J.n_voxels;
    u32 cmin = 0xffffffffu, cmax = 0u, emin = 0xffffffffu, emax = 0u;
    const int cnt = min(t.count, J.n_kept - t.begin);          // keys past n_kept were filtered out before the sort
    if (cnt <= 0) return;                                      // block-uniform
    for (int i0 = 0; i0 < cnt; i0 += TPB) {
        const int i = i0 + threadIdx.x;
        int h = 0;
        u64 k = 0;
        if (i < cnt) {
            int g = t.begin + i;
            k = src[g];
            u64 prev = g > 0 ? src[g - 1] : 0ull;
            h = head_level(k, prev, g == 0, n);
            if (i + 2 * TPB < cnt) prefetch_l2(src + g + 2 * TPB);
        }
        const int hh = h == 0 ? 99 : h;
        // Only levels >= the smallest head level of the warp can open a node here: consecutive sorted keys share long
        // prefixes, so that is usually n-3 .. n and the 21-level ballot / emit loops shrink to a handful of iterations.
        const int wmin = __reduce_min_sync(0xffffffffu, hh);
        u32 rank[MAXL + 2];
        {
            u32 b = __ballot_sync(0xffffffffu, h > 0);
            rank[0] = __popc(b & ltmask);
            if (lane == 0) s_wcnt[warp][0] = __popc(b);
        }
#pragma unroll
        for (int L = 1; L <= MAXL; ++L) {
            if (L < wmin) {                                     // warp-uniform
                rank[L] = 0;
                if (lane == 0) s_wcnt[warp][L] = 0;
                continue;
            }
            u32 b = __ballot_sync(0xffffffffu, hh <= L);
            rank[L] = __popc(b & ltmask);
            if (lane == 0) s_wcnt[warp][L] = __popc(b);
        }
        __syncthreads();
        if (threadIdx.x <= MAXL) {
            u32 run = 0;
            for (int w = 0; w < 8; ++w) { s_wpre[w][threadIdx.x] = run; run += s_wcnt[w][threadIdx.x]; }
        }
        __syncthreads();
        const u32 x = compact3(k >> 2), y = compact3(k >> 1), z = compact3(k);
        const u64 P = pack_pos(x, y, z);
        const u32 v = s_base[0] + s_wpre[warp][0] + rank[0];
        if (h > 0) {
            if (vox_key) vox_key[J.vox_start + v] = k;
            A.vdig[J.vox_start + v] = (uint8_t)((k & 7) + 1);
            // x & m is monotone in x, so every level's min / max node coordinate follows from the voxel extremes
            const u32 lo = min(x, min(y, z)), hi = max(x, max(y, z));
            cmin = min(cmin, lo); cmax = max(cmax, hi);
            if ((int)v != n_vox - 1) { emin = min(emin, lo); emax = max(emax, hi); }
        }
#pragma unroll
        for (int L = 1; L <= MAXL; ++L) {
            if (L > n) break;                                   // uniform
            if (L < wmin) continue;                             // warp-uniform: nobody in this warp opens a node on level L
            const bool mine = (h > 0) && (L >= h);
            if (mine) {
                const u32 kL = s_base[L] + s_wpre[warp][L] + rank[L];
                const long long r = node0 + J.level_start[L - 1] + kL;
                A.lo[r] = (uint16_t)(L | (((L == 1) ? 1u : (u32)((k >> (3 * (n - L + 1))) & 7) + 1u) << 8));
                u32 par = 0;
                if (L > 1) {
                    u32 kp = s_base[L - 1] + s_wpre[warp][L - 1] + rank[L - 1] + ((hh <= L - 1) ? 1u : 0u) - 1u;
                    par = (u32)J.level_start[L - 2] + kp;
                }
                A.parent[r] = par;
                A.pos[r] = P & s_keep[L];
                A.fc[r] = (L < n) ? (u32)J.level_start[L] + s_base[L + 1] + s_wpre[warp][L + 1] + rank[L + 1] : v;
            }
        }
        __syncthreads();
        if (threadIdx.x <= MAXL) {
            u32 add = 0;
            for (int w = 0; w < 8; ++w) add += s_wcnt[w][threadIdx.x];
            s_base[threadIdx.x] += add;
        }
        __syncthreads();
    }
    cmin = __reduce_min_sync(0xffffffffu, cmin); cmax = __reduce_max_sync(0xffffffffu, cmax);
    emin = __reduce_min_sync(0xffffffffu, emin); emax = __reduce_max_sync(0xffffffffu, emax);
    if (lane == 0) {
        atomicMin(&s_mm[0], cmin); atomicMax(&s_mm[1], cmax); atomicMin(&s_mm[2], emin); atomicMax(&s_mm[3], emax);
    }
    __syncthreads();
    // level L (1-based): nodes are the voxels masked to the level's cell size.  The dropped last row (Octree.py:259-262)
    // is the last node of the LAST level only, i.e. the job's last voxel; every other level keeps all its nodes.
    if (threadIdx.x >= 1 && threadIdx.x <= n) {
        const int L = threadIdx.x;
        const u32 m = ~((1u << (n - L + 1)) - 1u);
        const bool excl = J.drop_last && L == n;
        const u32 mn = excl ? s_mm[2] : s_mm[0], mx = excl ? s_mm[3] : s_mm[1];
        if (mn != 0xffffffffu) {
            atomicMin(&J.pos_min[L - 1], mn & m);
            atomicMax(&J.pos_max[L - 1], mx & m);
        }
    }
}

// ------------------------------------------------------------------------------------------
// K6': bottom-up tree build, one pass per level ("per-level occupancy-byte emission")
// ------------------------------------------------------------------------------------------
// Pass p handles, for every job, the children on level Lc = depth + 1 - p (p = 0: the sorted voxel keys) and creates their
// parents on level Lc - 1: consecutive children with the same key prefix share a parent, so a parent's BFS index inside its
// level is the number of prefix changes in front of it (warp ballots + an 8-entry block scan + a decoupled look-back over
// one word per tile), and its occupancy byte is the OR of its children's digit bits -- a segmented OR over <= 8 adjacent
// lanes (three shuffle rounds).  Every child is read once (its masked Morton key, 8 B) and written once (parent index +
// level|octant, 6 B); every parent gets its key (8 B) and occupancy (1 B).  ~50 instructions per node, all accesses
// coalesced; the all-levels-in-one-pass kernel above spends ~200 per node because on the sparse upper levels most lanes of
// a level iteration idle, and needs k_occupancy (another ~110) behind it.  Level sizes come from k_head_hist / k_level_scan,
// so nodes land directly in BFS order.
// Segments cut by a 32-key boundary add their part with atomicOr on the aligned word of the (zeroed) occupancy array; the
// byte stores of complete segments next to them are safe (an L2 atomic ORs zeros into foreign bytes).
constexpr u32 PF_AGG = 1u << 30, PF_INC = 1u << 31, PF_MASK = (1u << 30) - 1;
// Everything a block of k_level_pass needs, prepared on the host (which has the level sizes after scp_octree_plan): one
// 64-byte load instead of the tile -> job -> level-table chain of dependent global loads, and exact per-pass tile lists
// (no empty blocks on the sparse upper levels).
struct PassTile {
    int job, begin, count, first;       // children [begin, begin+count) of the level; first = first tile of the job's level in the list
    int n_child, sh, Lc, depth;
    u32 lsp, lsc;                       // level starts (inside the job) of the parents' / children's level
    long long node0, src_off;           // job's first node; children keys: sorted + src_off (voxel pass) or pos + node0 + lsc
    long long vox_start;
    u32 drop_last, pad;
};

// MODE 0: inner level (children are nodes), 1: voxel pass (children are the sorted keys), 2: voxel pass + voxel_key output
template <int MODE>
__global__ void __launch_bounds__(TPB) k_level_pass(const u64* __restrict__ sorted, const PassTile* __restrict__ tiles, int n_tiles,
                                                     JobDev* jobs, NodeArrays A, u64* __restrict__ vox_key,
                                                     u32* desc /* [2][n_tiles] of this pass */, u32* ticket, u32* err) {
    __shared__ u32 s_wcnt[8], s_wvox[8], s_base[2];
    __shared__ u32 s_mm[4];
    __shared__ int s_tile;
    if (threadIdx.x == 0) s_tile = (int)atomicAdd(ticket, 1u);
    if (threadIdx.x < 4) s_mm[threadIdx.x] = (threadIdx.x & 1) ? 0u : 0xffffffffu;
    __syncthreads();
    const int tix = s_tile;
    if (tix >= n_tiles) return;
    const PassTile t = tiles[tix];
    const int n = t.depth;
    const int Lc = t.Lc, Lp = Lc - 1;                               // children / parents level (1-based; n+1 = voxels)
    constexpr bool vox = MODE != 0;
    const int Nc = t.n_child;
    const int cnt = t.count;
    const int sb = n - Lc + 1;                                      // coordinate bit that selects the child inside its parent
    const long long node0 = t.node0;
    const u32 lsp = t.lsp;
    // nodes carry their cell origin packed as x | y << 21 | z << 42: a parent's origin is its child's with one more bit
    // per axis cleared, two children share a parent iff their origins agree on the parent's bits
    const u32 km = (~((2u << sb) - 1u)) & 0x1fffffu;
    const u64 keep = pack_pos(km, km, km);
    constexpr bool want_vk = MODE == 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 lemask = (2u << lane) - 1u;                           // lanes <= me
    const int wbeg = warp * WKEYS;

    u64 c[WITER];                                                   // packed origin of the child (voxel pass: full-resolution position)
    u64 kraw[want_vk ? WITER : 1];                                  // voxel pass with voxel_key output only: the Morton keys
    u32 hflag = 0, dflag = 0;                                       // bit it: this lane's child opens a parent / is a new distinct voxel
    if (vox) {
        const u64* src = sorted + t.src_off + t.begin;
#pragma unroll
        for (int it = 0; it < WITER; ++it) {
            const int idx = wbeg + it * 32 + lane;
            const u64 k = idx < cnt ? __ldg(src + idx) : 0ull;
            if (want_vk) kraw[it] = k;
            c[it] = pack_pos(compact3(k >> 2), compact3(k >> 1), compact3(k));
        }
    } else {
        const u64* src = A.pos + node0 + t.lsc + t.begin;
#pragma unroll
        for (int it = 0; it < WITER; ++it) {
            const int idx = wbeg + it * 32 + lane;
            c[it] = idx < cnt ? __ldg(src + idx) : 0ull;
        }
    }
    {
        const int g0 = t.begin + wbeg;
        u64 carry = 0ull;
        if (lane == 0 && g0 > 0 && wbeg < cnt) {
            if (vox) { const u64 k = __ldg(sorted + t.src_off + g0 - 1); carry = pack_pos(compact3(k >> 2), compact3(k >> 1), compact3(k)); }
            else carry = __ldg(A.pos + node0 + t.lsc + g0 - 1);
        }
        u32 wc = 0, wv = 0;
#pragma unroll
        for (int it = 0; it < WITER; ++it) {
            u64 prev = __shfl_up_sync(0xffffffffu, c[it], 1);
            if (lane == 0) prev = carry;
            const int idx = wbeg + it * 32 + lane;
            const bool valid = idx < cnt, first = (g0 + it * 32 + lane == 0);
            const u64 df = c[it] ^ prev;
            const bool h = valid && (first || (df & keep) != 0ull);
            const bool d = valid && (first || df != 0ull);
            hflag |= h ? (1u << it) : 0u; dflag |= d ? (1u << it) : 0u;
            wc += __popc(__ballot_sync(0xffffffffu, h));
            if (want_vk) wv += __popc(__ballot_sync(0xffffffffu, d));
            carry = __shfl_sync(0xffffffffu, c[it], 31);
        }
        if (lane == 0) { s_wcnt[warp] = wc; s_wvox[warp] = wv; }
    }
    __syncthreads();
    if (threadIdx.x < 2) {                                          // thread 0: parents, thread 1: distinct voxels (only if wanted)
        const int q = threadIdx.x;
        u32 base = 0;
        if (q == 0 || want_vk) {
            u32 total = 0;
            for (int w = 0; w < 8; ++w) total += q ? s_wvox[w] : s_wcnt[w];
            volatile u32* vd = desc + (size_t)q * n_tiles;
            vd[tix] = total | PF_AGG;
            // tiles of the job's level in front of this one (tickets are handed out in launch order: they run or are done)
            for (int pt = tix - 1; pt >= t.first; --pt) {
                u32 v;
                int spins = 0;
                do { v = vd[pt]; } while ((v & (PF_AGG | PF_INC)) == 0 && ++spins < (1 << 22));
                if ((v & (PF_AGG | PF_INC)) == 0) { atomicExch(err, 1u); break; }      // never hang the GPU
                base += v & PF_MASK;
                if (v & PF_INC) break;
            }
            vd[tix] = ((base + total) & PF_MASK) | PF_INC;
        }
        s_base[q] = base;
    }
    __syncthreads();
    u32 r0 = s_base[0], v0 = s_base[1];
    for (int w = 0; w < warp; ++w) { r0 += s_wcnt[w]; v0 += s_wvox[w]; }
    u32 cmin = 0xffffffffu, cmax = 0u, emin = 0xffffffffu, emax = 0u;
    u32* occ32 = reinterpret_cast<u32*>(A.occ);
    uint8_t* occ_p = A.occ + node0 + lsp;
    u64* pos_p = A.pos + node0 + lsp;
    u32* par_c = A.parent + node0 + t.lsc + t.begin + wbeg + lane;
    uint16_t* lo_c = A.lo + node0 + t.lsc + t.begin + wbeg + lane;
    u64 clast = 0ull;
    if (vox && t.drop_last) { const u64 k = __ldg(sorted + t.src_off + Nc - 1); clast = pack_pos(compact3(k >> 2), compact3(k >> 1), compact3(k)); }
#pragma unroll
    for (int it = 0; it < WITER; ++it) {
        const int idx = wbeg + it * 32 + lane;
        const bool valid = idx < cnt;
        const bool h = (hflag >> it) & 1u;
        const u32 b = __ballot_sync(0xffffffffu, h), vm = __ballot_sync(0xffffffffu, valid);
        if (vm == 0) break;                                         // warp-uniform: past the end of the tile
        const u64 cc = c[it];
        const u32 r = r0 + __popc(b & lemask) - 1u;                 // parent of this child, index inside level Lp
        const u32 dig = ((u32)(cc >> sb) & 1u) << 2 | ((u32)(cc >> (21 + sb)) & 1u) << 1 | ((u32)(cc >> (42 + sb)) & 1u);
        u32 v = valid ? (1u << dig) : 0u;
#pragma unroll
        for (int o = 1; o <= 4; o <<= 1) {                           // segmented OR scan: a parent has <= 8 children
            const u32 uv = __shfl_up_sync(0xffffffffu, v, o), ur = __shfl_up_sync(0xffffffffu, r, o);
            if (lane >= o && ur == r) v |= uv;
        }
        const bool next_valid = lane < 31 && ((vm >> (lane + 1)) & 1u), next_head = lane < 31 && ((b >> (lane + 1)) & 1u);
        if (valid && (!next_valid || next_head)) {                  // last child of its parent among these 32
            const bool headed = (b & lemask) != 0;                  // the parent was opened inside these 32 children
            const bool closed = next_head || (t.begin + idx == Nc - 1);
            if (headed && closed) occ_p[r] = (uint8_t)v;
            else { const long long g = node0 + lsp + r; atomicOr(occ32 + (g >> 2), v << (8 * (int)(g & 3))); }
        }
        if (h) {
            pos_p[r] = cc & keep;
            if (Lp == 1) { A.parent[node0] = 0u; A.lo[node0] = (uint16_t)(1u | (1u << 8)); }      // root: level 1, octant 1
        }
        if (!vox) {
            if (valid) { par_c[it * 32] = lsp + r; lo_c[it * 32] = (uint16_t)((u32)Lc | ((dig + 1u) << 8)); }
        } else {
            const bool d = (dflag >> it) & 1u;
            if (d) {
                // x & m is monotone in x, so every level's min / max node coordinate follows from the voxel extremes
                const u32 x = (u32)cc & 0x1fffffu, y = (u32)(cc >> 21) & 0x1fffffu, z = (u32)(cc >> 42);
                const u32 lo = min(x, min(y, z)), hi = max(x, max(y, z));
                cmin = min(cmin, lo); cmax = max(cmax, hi);
                if (cc != clast) { emin = min(emin, lo); emax = max(emax, hi); }
            }
            if (want_vk) {
                const u32 bd = __ballot_sync(0xffffffffu, d);
                if (d) vox_key[t.vox_start + v0 + __popc(bd & lemask) - 1u] = kraw[want_vk ? it : 0];
                v0 += __popc(bd);
            }
        }
        r0 += __popc(b);
    }
    if (!vox) return;                                               // block-uniform
    cmin = __reduce_min_sync(0xffffffffu, cmin); cmax = __reduce_max_sync(0xffffffffu, cmax);
    emin = __reduce_min_sync(0xffffffffu, emin); emax = __reduce_max_sync(0xffffffffu, emax);
    if (lane == 0) {
        atomicMin(&s_mm[0], cmin); atomicMax(&s_mm[1], cmax); atomicMin(&s_mm[2], emin); atomicMax(&s_mm[3], emax);
    }
    __syncthreads();
    // level L (1-based): nodes are the voxels masked to the level's cell size.  The dropped last row (Octree.py:259-262)
    // is the last node of the LAST level only, i.e. the job's last voxel; every other level keeps all its nodes.
    if (threadIdx.x >= 1 && threadIdx.x <= n) {
        JobDev& J = jobs[t.job];
        const int L = threadIdx.x;
        const u32 m = ~((1u << (n - L + 1)) - 1u);
        const bool excl = t.drop_last && L == n;
        const u32 mn = excl ? s_mm[2] : s_mm[0], mx = excl ? s_mm[3] : s_mm[1];
        if (mn != 0xffffffffu) {
            atomicMin(&J.pos_min[L - 1], mn & m);
            atomicMax(&J.pos_max[L - 1], mx & m);
        }
    }
}

// ------------------------------------------------------------------------------------------
// K7a: occupancy byte = OR over the node's children run           Octree.py:175-176
// ------------------------------------------------------------------------------------------
// The <= 8 children of a node are consecutive bytes of `octant` (of `vdig` for the deepest level, whose children are the
// voxels); two nodes per thread are in flight.
constexpr int NPT = 4;
__global__ void __launch_bounds__(TPB) k_occupancy(const Tile* __restrict__ tiles, const JobDev* __restrict__ jobs, NodeArrays A) {
    __shared__ int s_ls[MAXL + 2];
    const Tile t = tiles[blockIdx.x];
    const JobDev& J = jobs[t.job];
    const int n = J.depth;
    if (threadIdx.x <= MAXL + 1) s_ls[threadIdx.x] = J.level_start[threadIdx.x];
    __syncthreads();
    const uint16_t* __restrict__ lvl = A.lo + J.node_start;
    const u32* __restrict__ fc = A.fc + J.node_start;
    const uint8_t* __restrict__ vd = A.vdig + J.vox_start;
    uint8_t* __restrict__ out = A.occ + J.node_start;
    const u32 n_vox = (u32)J.n_voxels;
    for (int i0 = threadIdx.x; i0 < t.count; i0 += 2 * TPB) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {                           // next round's records into L2 (see k_context_lean)
            const int i = i0 + (4 + u) * TPB;
            if (i < t.count) { prefetch_l2(fc + t.begin + i); if ((threadIdx.x & 15) == 0) prefetch_l2(lvl + t.begin + i); }
        }
        int L[2];
        u32 c0[2], c1[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = i0 + u * TPB;
            L[u] = 0; c0[u] = c1[u] = 0;
            if (i < t.count) { const int loc = t.begin + i; L[u] = lvl[loc] & 0xff; c0[u] = fc[loc]; c1[u] = fc[loc + 1]; }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (L[u] == 0) continue;
            const int loc = t.begin + i0 + u * TPB;
            const bool leaf = (L[u] == n);
            if (loc + 1 == s_ls[L[u]]) c1[u] = leaf ? n_vox : (u32)s_ls[L[u] + 1];       // level_start[L] = start of level L+1
            u32 occ = 0;
            if (leaf) { for (u32 c = c0[u]; c < c1[u]; ++c) occ |= 1u << (vd[c] - 1); }
            else { for (u32 c = c0[u]; c < c1[u]; ++c) occ |= 1u << ((lvl[c] >> 8) - 1); }
            out[loc] = (uint8_t)occ;
        }
    }
}

// ------------------------------------------------------------------------------------------
// K7b: K=4 ancestor context + level-normalised positions (+ optional reference-layout expansion)
//   Octree.py:102-137 gen_K_parent_seq ; encode_dataset_ehem.py:54,66-72,85-93
// ------------------------------------------------------------------------------------------
// float32((x - mn) / den) of encode_dataset_ehem.py:70-72 without float64: x - mn = d and max - min = D are integers below
// 2^21, den = D + 1e-9 (or D for the last level of a mullevel sub-octree).  The reference's float64 quotient d / den rounded
// to float32 equals the correctly rounded float32 quotient d / D: the 1e-9 moves the quotient by a relative 1e-9 / D, less
// than its distance 2^-25 / D from the nearest float32 rounding midpoint (d / D with D < 2^21 has at most 21 significant
// quotient bits, so it is never ON a midpoint), and the intermediate float64 rounding (2^-53) is smaller still.  D = 0 (a
// single-cell level): 0 / 1e-9 = 0 with the epsilon, 0 / 0 = NaN without -- `zero` carries that value.
__device__ __forceinline__ float norm_pos_f(u32 x, u32 mn, float D, float zero) {
    return D > 0.f ? __fdiv_rn(__uint2float_rn(x - mn), D) : zero;
}

// Lean variant for the encoder's outputs (occ, sym, ctx, pos_norm): own record + 3 dependent (parent, occupancy) gathers;
// the ancestors' level, octant and cell origin are derived from the own record (an ancestor's cell is the own cell with
// more low bits cleared, its octant is the coordinate bit triple of its level).  The kernel is instruction-issue bound
// (ncu: ~300 instructions per node before this version), hence the bit tricks: all four octants from one 12-bit word,
// byte permutes for the 12 context bytes, the divide out of line.  Four nodes per thread keep four gather chains in flight.
template <int NPT, int MINB>
__global__ void __launch_bounds__(TPB, MINB) k_context_lean(const Tile* __restrict__ tiles, const JobDev* __restrict__ jobs,
                                                             NodeArrays A, scp_octree_out O) {
    __shared__ u32 s_mn[MAXL + 1], s_lv[MAXL + 1];
    __shared__ float s_D[MAXL + 1], s_zero[MAXL + 1];
    __shared__ u32 s_stage[TPB / 32][2][96];
    const Tile t = tiles[blockIdx.x];
    const JobDev& J = jobs[t.job];
    const int n = J.depth;
    if (threadIdx.x >= 1 && threadIdx.x <= n) {
        const int L = threadIdx.x;
        s_mn[L] = J.pos_min[L - 1];
        s_D[L] = (float)(J.pos_max[L - 1] - J.pos_min[L - 1]);
        s_zero[L] = (L == n && !J.pos_eps_last) ? __int_as_float(0x7fc00000) : 0.f;
        // level bytes of (great-grandparent, grandparent, parent, self): 0 = missing; the last level is clipped to lidar_level
        u32 lv = 0;                                               // (encode_dataset_ehem.py:86)
        for (int q = 0; q < 4; ++q) {
            int Lk = max(L - (3 - q), 0);
            if (L == n && n > J.lidar_level) Lk = min(Lk, J.lidar_level);
            lv |= (u32)Lk << (8 * q);
        }
        s_lv[L] = lv;
    }
    __syncthreads();
    const long long node0 = J.node_start, row0 = J.row_start;
    const uint16_t* __restrict__ a_lvl = A.lo + node0;
    const uint8_t* __restrict__ a_occ = A.occ + node0;
    const u32* __restrict__ a_par = A.parent + node0;
    const u64* __restrict__ a_pos = A.pos + node0;
    const int cnt = min(t.count, J.n_rows - t.begin);             // the dropped last row (Octree.py:259-262)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int u = 0; u < NPT; ++u) {                                        // warm-up: round 1 (round 0 is loaded right away)
        const int i = (int)threadIdx.x + (NPT + u) * TPB;
        if (i < cnt) {
            const int loc = t.begin + i;
            prefetch_l2(a_par + loc); prefetch_l2(a_pos + loc);
            if ((lane & 7) == 0) { prefetch_l2(a_lvl + loc); prefetch_l2(a_occ + loc); }
        }
    }
    for (int i0 = threadIdx.x; i0 - lane < cnt; i0 += NPT * TPB) {        // warp-uniform trip count
        int L[NPT];
        u64 pp[NPT];
        u32 occp[NPT], a[NPT];                                     // occp: (occ-1) of ggp | gp << 8 | parent << 16 | self << 24
        // the own records of the NEXT round are pulled into L2 now: the three dependent gather rounds below would otherwise
        // leave the DRAM pipe idle for three quarters of every round
#pragma unroll
        for (int u = 0; u < NPT; ++u) {
            const int i = i0 + (2 * NPT + u) * TPB;                 // two rounds ahead
            if (i < cnt) {
                const int loc = t.begin + i;
                prefetch_l2(a_par + loc); prefetch_l2(a_pos + loc);
                if ((lane & 7) == 0) { prefetch_l2(a_lvl + loc); prefetch_l2(a_occ + loc); }
            }
        }
#pragma unroll
        for (int u = 0; u < NPT; ++u) {
            const int i = i0 + u * TPB;
            L[u] = 0; occp[u] = 0; a[u] = 0; pp[u] = 0;
            if (i < cnt) {
                const int loc = t.begin + i;
                L[u] = a_lvl[loc] & 0xff; a[u] = a_par[loc];
                occp[u] = (u32)a_occ[loc];
                pp[u] = a_pos[loc];
            }
        }
#pragma unroll
        for (int u = 0; u < NPT; ++u) occp[u] = ((occp[u] - 1u) & 0xffu) << 24;
#pragma unroll
        for (int k = 2; k >= 0; --k) {
#pragma unroll
            for (int u = 0; u < NPT; ++u) {
                u32 oc = 0x100u;                                    // missing ancestor: occ 256 -> byte 255
                if (L[u] - (3 - k) >= 1) {
                    oc = a_occ[a[u]];
                    if (k > 0) a[u] = a_par[a[u]];
                }
                occp[u] |= ((oc - 1u) & 0xffu) << (8 * k);
            }
        }
#pragma unroll
        for (int u = 0; u < NPT; ++u) {
            // a warp's 32 nodes are 32 consecutive output rows: ctx (12 B) and pos_norm (12 B) go through a per-warp staging
            // tile and leave as three fully coalesced 128-byte stores each instead of 4-byte stores 12 bytes apart
            const int wbase = i0 - lane + u * TPB;                 // tile-relative index of the warp's first node
            const int nvw = 3 * max(0, min(32, cnt - wbase));       // valid words of the staging tile
            if (nvw == 0) continue;                                 // warp-uniform
            const long long o0 = row0 + t.begin + wbase;
            const int Lu = L[u];
            u32 w0 = 0, w1 = 0, w2 = 0;
            float f0 = 0.f, f1 = 0.f, f2 = 0.f;
            if (Lu > 0) {
                const u32 self = occp[u] >> 24;
                if (O.occ) O.occ[o0 + lane] = (uint8_t)(self + 1);
                if (O.sym) O.sym[o0 + lane] = (int16_t)self;
                // bit j of (coordinate >> sh3) is the octant bit of the ancestor j levels up (sh3 = lowest bit of the own cell)
                const int sh3 = n - Lu + 1;
                u32 px, py, pz;
                unpack_pos(pp[u], A.morton, px, py, pz);
                const u32 W = (((px >> sh3) & 0xfu) << 8) | (((py >> sh3) & 0xfu) << 4) | ((pz >> sh3) & 0xfu);
                u32 OC = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int j = 3 - k, Lk = Lu - j;
                    u32 o = ((((W >> j) & 0x111u) * 0x124u) >> 8 & 7u) + 1u;       // 4*xbit + 2*ybit + zbit + 1
                    if (Lk <= 1) o = (Lk == 1) ? 1u : 0u;                           // root: octant 1; missing: 0
                    OC |= o << (8 * k);
                }
                const u32 LV = s_lv[Lu];
                // 12 context bytes (level, octant, occ-1) x 4 from the three byte-planes
                w0 = __byte_perm(__byte_perm(LV, OC, 0x1040), occp[u], 0x3410);
                w1 = __byte_perm(__byte_perm(LV, OC, 0x6205), occp[u], 0x3250);
                w2 = __byte_perm(__byte_perm(LV, OC, 0x0730), occp[u], 0x7216);
                if (O.pos_norm) {
                    const u32 mn = s_mn[Lu];
                    const float D = s_D[Lu], zero = s_zero[Lu];
                    f0 = norm_pos_f(px, mn, D, zero); f1 = norm_pos_f(py, mn, D, zero); f2 = norm_pos_f(pz, mn, D, zero);
                }
            }
            u32* sc = s_stage[warp][0];
            float* sp = reinterpret_cast<float*>(s_stage[warp][1]);
            sc[3 * lane] = w0; sc[3 * lane + 1] = w1; sc[3 * lane + 2] = w2;      // stride 3 words: conflict-free
            sp[3 * lane] = f0; sp[3 * lane + 1] = f1; sp[3 * lane + 2] = f2;
            __syncwarp();
            if (O.ctx) {
                u32* dst = reinterpret_cast<u32*>(O.ctx + 12 * o0);
#pragma unroll
                for (int j = 0; j < 3; ++j) if (32 * j + lane < nvw) dst[32 * j + lane] = sc[32 * j + lane];
            }
            if (O.pos_norm) {
                float* dst = O.pos_norm + 3 * o0;
#pragma unroll
                for (int j = 0; j < 3; ++j) if (32 * j + lane < nvw) dst[32 * j + lane] = sp[32 * j + lane];
            }
            __syncwarp();
        }
    }
}

// General variant: every optional output of scp_octree_out
__global__ void __launch_bounds__(TPB) k_context(const Tile* __restrict__ tiles, const JobDev* __restrict__ jobs,
                                                  NodeArrays A, scp_octree_out O) {
    const Tile t = tiles[blockIdx.x];
    const JobDev& J = jobs[t.job];
    const int n = J.depth;
    for (int i = threadIdx.x; i < t.count; i += TPB) {
        const int loc = t.begin + i;
        if (loc >= J.n_rows) continue;                    // the dropped last row (Octree.py:259-262)
        const long long r = J.node_start + loc;
        const long long o = J.row_start + loc;
        const u32 lo16 = A.lo[r];
        const int L = lo16 & 0xff;
        const bool last_block = (L == n);
        int lv[4], oc[4], occ[4];
        u32 px[4], py[4], pz[4];
        lv[3] = L; oc[3] = lo16 >> 8; occ[3] = A.occ[r];
        unpack_pos(A.pos[r], A.morton, px[3], py[3], pz[3]);
        u32 a = A.parent[r];
#pragma unroll
        for (int k = 2; k >= 0; --k) {
            const int Lk = L - (3 - k);
            if (Lk >= 1) {
                const long long ar = J.node_start + a;
                const int sh = n - Lk + 1;                      // lowest coordinate bit that survives on level Lk
                const u32 m = sh >= 32 ? 0u : ~((1u << sh) - 1u);
                lv[k] = Lk;
                occ[k] = A.occ[ar];
                a = A.parent[ar];
                px[k] = px[3] & m; py[k] = py[3] & m; pz[k] = pz[3] & m;
                oc[k] = (Lk == 1) ? 1 : (int)((((px[3] >> sh) & 1u) << 2) | (((py[3] >> sh) & 1u) << 1) | ((pz[3] >> sh) & 1u)) + 1;
            } else {
                lv[k] = 0; oc[k] = 0; occ[k] = 256; px[k] = py[k] = pz[k] = 0;
            }
        }
        if (O.occ) O.occ[o] = (uint8_t)occ[3];
        if (O.sym) O.sym[o] = (int16_t)(occ[3] - 1);
        if (O.level) O.level[o] = (uint8_t)lv[3];
        if (O.octant) O.octant[o] = (uint8_t)oc[3];
        if (O.parent) O.parent[o] = A.parent[r];
        if (O.pos) { O.pos[3 * o] = px[3]; O.pos[3 * o + 1] = py[3]; O.pos[3 * o + 2] = pz[3]; }
        if (O.ctx) {
            uint8_t c[12];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int l = lv[k];
                if (last_block && l > J.lidar_level) l = J.lidar_level;     // encode_dataset_ehem.py:86
                c[3 * k] = (uint8_t)l; c[3 * k + 1] = (uint8_t)oc[k]; c[3 * k + 2] = (uint8_t)(occ[k] - 1);
            }
            uint32_t* dst = reinterpret_cast<uint32_t*>(O.ctx + 12 * o);
            dst[0] = c[0] | (c[1] << 8) | (c[2] << 16) | ((u32)c[3] << 24);
            dst[1] = c[4] | (c[5] << 8) | (c[6] << 16) | ((u32)c[7] << 24);
            dst[2] = c[8] | (c[9] << 8) | (c[10] << 16) | ((u32)c[11] << 24);
        }
        if (O.pos_norm) {
            const double mn = (double)J.pos_min[L - 1];
            const double den = (double)(J.pos_max[L - 1] - J.pos_min[L - 1]) +
                               ((last_block && !J.pos_eps_last) ? 0.0 : 1e-9);
            O.pos_norm[3 * o] = (float)(((double)px[3] - mn) / den);
            O.pos_norm[3 * o + 1] = (float)(((double)py[3] - mn) / den);
            O.pos_norm[3 * o + 2] = (float)(((double)pz[3] - mn) / den);
        }
        if (O.ctx_pos) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                O.ctx_pos[12 * o + 3 * k] = px[k]; O.ctx_pos[12 * o + 3 * k + 1] = py[k]; O.ctx_pos[12 * o + 3 * k + 2] = pz[k];
            }
        }
        if (O.rows_i64) {
            long long* w = reinterpret_cast<long long*>(O.rows_i64) + 24 * o;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                w[6 * k] = occ[k]; w[6 * k + 1] = lv[k]; w[6 * k + 2] = oc[k];
                w[6 * k + 3] = px[k]; w[6 * k + 4] = py[k]; w[6 * k + 5] = pz[k];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// K6 + K7a in one pass over the SORTED KEYS (tree builder 2, the default for the encoder's outputs occ / sym / ctx / pos_norm).
// ------------------------------------------------------------------------------------------
// A node on level L is the run of sorted keys that share their first L-1 octal digits; it is "opened" by the first key of the
// run, i.e. by every key whose head level h is <= L.  With the per-level ballots of a warp a key knows, for EVERY level at
// once, the BFS rank of the node it opens there (rank of heads) and of the node that contains it (heads at or before it,
// minus one).  k_tree_occ<RECORDS> uses that twice per (key, level): the occupancy byte of a node = OR over the digits of the
// keys that open one of its children -- a segmented OR over the warp (segments = parents), plain byte stores for parents
// whose run lies inside the 32 keys, atomicOr on the aligned word for the <= 2 runs per level cut by the border -- and, with
// RECORDS, the node record (level | octant, parent index, cell origin) the context kernel gathers from.  No first-child
// array, no voxel-digit array, no separate occupancy kernel (k_emit_nodes + k_occupancy: 28 + 6 B/node more traffic and a
// dependent level -> first_child -> children load chain).  The voxel extremes that give every level's (min, max) coordinate
// for the normalisation come out of the same pass.
// Warp-autonomous: a warp owns one chunk (WKEYS = 256 consecutive sorted keys, 8 rounds of 32); the rank of its first node on
// every level comes from the chunk prefix table (k_head_hist / k_level_scan) -- no block-level scan, no __syncthreads
// inside the key loop.
// Register layout "lane = level": lane L keeps the ballot of level L (lanes that open a node there; lane 0: new voxels) and
// the running rank base[L] of the chunk, so the per-level state is two registers, the update after a round is ONE add per
// lane, and the level loops are ordinary runtime loops (an unrolled 21-level body overflowed the instruction cache: half of
// the stall samples of the first version were "no instruction").  b[L] / base[L] reach all lanes by a shuffle from lane L.
__device__ __forceinline__ u32 tree_ballots(int h, int n, int lane, int& wmin) {
    const int hh = h == 0 ? 99 : h;
    wmin = __reduce_min_sync(0xffffffffu, hh);
    const u32 b0 = __ballot_sync(0xffffffffu, h > 0);
    u32 bvec = lane == 0 ? b0 : 0u;
    for (int L = max(wmin, 1); L <= n; ++L) {                           // warp-uniform bounds
        const u32 b = __ballot_sync(0xffffffffu, hh <= L);
        if (lane == L) bvec = b;
    }
    return bvec;
}

// lane L: nodes the job opened on level L in front of this chunk (lane 0: voxels)
__device__ __forceinline__ u32 chunk_bases(const u32* __restrict__ cb, int n, int lane) {
    const u32 c = (lane >= 1 && lane < NBINS) ? __ldg(cb + lane) : 0u;   // lane h: heads of level h in front of the chunk
    u32 inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 up = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += up; }
    const u32 vox = __shfl_sync(0xffffffffu, inc, NBINS - 1);
    return lane == 0 ? vox : (lane <= n ? inc : 0u);
}

template <bool RECORDS>
__global__ void __launch_bounds__(TPB, 4) k_tree_occ(const u64* __restrict__ keys, const Tile* __restrict__ tiles, JobDev* jobs,
                                                   const u32* __restrict__ chunk_base, NodeArrays A, u64* __restrict__ vox_key) {
    __shared__ u32 s_mm[4];
    __shared__ u32 s_ls[MAXL + 2];
    __shared__ u64 s_keep[MAXL + 2];        // packed-position bits that survive on level L (cell size 2^(n-L+1))
    const Tile t = tiles[blockIdx.x];
    JobDev& J = jobs[t.job];
    const int n = J.depth;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 ltmask = (1u << lane) - 1u, lemask = (2u << lane) - 1u;
    if (threadIdx.x <= MAXL + 1) s_ls[threadIdx.x] = (u32)J.level_start[threadIdx.x];
    if (threadIdx.x >= 32 && threadIdx.x <= 32 + MAXL) {
        const int L = threadIdx.x - 32;
        const u32 m = (L >= 1 && L <= n) ? (~((1u << (n - L + 1)) - 1u)) & 0x1fffffu : 0u;
        s_keep[L] = pack_pos(m, m, m);
    }
    if (threadIdx.x >= 64 && threadIdx.x < 68) s_mm[threadIdx.x - 64] = (threadIdx.x & 1) ? 0u : 0xffffffffu;
    __syncthreads();
    const u64* src = keys + J.key_begin;
    const long long node0 = J.node_start;
    uint8_t* occ = A.occ + node0;
    u32* occ32 = reinterpret_cast<u32*>(A.occ);
    uint16_t* r_lo = A.lo + node0;
    u32* r_par = A.parent + node0;
    u64* r_pos = A.pos + node0;
    const int n_vox = J.n_voxels;
    u32 cmin = 0xffffffffu, cmax = 0u, emin = 0xffffffffu, emax = 0u;
    const int cnt = min(t.count, J.n_kept - t.begin);          // keys past n_kept were filtered out before the sort
    const int wbeg = warp * WKEYS;
    if (wbeg < cnt) {
        u32 basev = chunk_bases(chunk_base + ((size_t)(t.first + t.begin / TILE) * CHUNKS + warp) * NBINS, n, lane);
        const int g0 = t.begin + wbeg;
        u64 carry = (lane == 0 && g0 > 0) ? __ldg(src + g0 - 1) : 0ull;
        u64 knext = wbeg + lane < cnt ? __ldg(src + g0 + lane) : SENTINEL;
        for (int it = 0; it < WITER; ++it) {
            const int idx = wbeg + it * 32 + lane;
            if (wbeg + it * 32 >= cnt) break;                   // warp-uniform
            const u64 k = knext;
            if (it + 1 < WITER) knext = idx + 32 < cnt ? __ldg(src + t.begin + idx + 32) : SENTINEL;
            u64 prev = __shfl_up_sync(0xffffffffu, k, 1);
            if (lane == 0) prev = carry;
            const int h = idx < cnt ? head_level(k, prev, g0 + it * 32 + lane == 0, n) : 0;
            carry = __shfl_sync(0xffffffffu, k, 31);
            int wmin;
            const u32 bvec = tree_ballots(h, n, lane, wmin);
            const u32 b0 = __shfl_sync(0xffffffffu, bvec, 0);
            const u32 vbase = __shfl_sync(0xffffffffu, basev, 0);
            u64 P = 0;
            if (h > 0) {
                const u32 v = vbase + __popc(b0 & ltmask);
                if (vox_key) vox_key[J.vox_start + v] = k;
                // x & m is monotone in x, so every level's min / max node coordinate follows from the voxel extremes
                const u32 x = compact3(k >> 2), y = compact3(k >> 1), z = compact3(k);
                if (RECORDS) P = pack_pos(x, y, z);
                const u32 lo = min(x, min(y, z)), hi = max(x, max(y, z));
                cmin = min(cmin, lo); cmax = max(cmax, hi);
                if ((int)v != n_vox - 1) { emin = min(emin, lo); emax = max(emax, hi); }
                if (RECORDS && h == 1) { r_lo[0] = (uint16_t)(1u | (1u << 8)); r_par[0] = 0u; r_pos[0] = 0ull; }   // root: level 1, octant 1
            }
            // children on level Lc (Lc = n+1: the voxels) add their digit bit to their parent on level Lc-1; the children
            // that are nodes (Lc <= n) also get their record: level | octant, parent index, cell origin
            for (int Lc = max(wmin, 2); Lc <= n + 1; ++Lc) {     // warp-uniform bounds
                const u32 bc = Lc <= n ? __shfl_sync(0xffffffffu, bvec, Lc) : b0;
                if (bc == 0) continue;                           // warp-uniform
                const u32 bp = __shfl_sync(0xffffffffu, bvec, Lc - 1);   // parents opened in these 32 keys = segment starts
                const u32 pbase = __shfl_sync(0xffffffffu, basev, Lc - 1);
                const bool head = (bc >> lane) & 1u;
                const u32 dig = (u32)(k >> (3 * (n + 1 - Lc))) & 7u;
                u32 v = head ? (1u << dig) : 0u;
                const u32 below = bp & lemask;
                const int seg0 = below ? 31 - __clz(below) : 0;  // first lane of my parent's run inside these 32 keys
                const u32 pr = (u32)s_ls[Lc - 2] + pbase + __popc(below) - 1u;      // the parent that contains this key
                if (RECORDS && Lc <= n) {
                    const u32 cbase = __shfl_sync(0xffffffffu, basev, Lc);
                    if (head) {
                        const u32 r = (u32)s_ls[Lc - 1] + cbase + __popc(bc & ltmask);
                        r_lo[r] = (uint16_t)((u32)Lc | ((dig + 1u) << 8));
                        r_par[r] = pr;
                        r_pos[r] = P & s_keep[Lc];
                    }
                }
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const u32 up = __shfl_up_sync(0xffffffffu, v, o);
                    if (lane - o >= seg0) v |= up;
                }
                const bool last = (lane == 31) || ((bp >> (lane + 1)) & 1u);
                if (last && v) {
                    if (below != 0 && lane != 31) occ[pr] = (uint8_t)v;        // run inside these 32 keys: complete
                    else { const long long g = node0 + pr; atomicOr(occ32 + (g >> 2), v << (8 * (int)(g & 3))); }
                }
            }
            basev += __popc(bvec);                               // every lane advances its own level
        }
    }
    cmin = __reduce_min_sync(0xffffffffu, cmin); cmax = __reduce_max_sync(0xffffffffu, cmax);
    emin = __reduce_min_sync(0xffffffffu, emin); emax = __reduce_max_sync(0xffffffffu, emax);
    if (lane == 0 && cmin != 0xffffffffu) {
        atomicMin(&s_mm[0], cmin); atomicMax(&s_mm[1], cmax); atomicMin(&s_mm[2], emin); atomicMax(&s_mm[3], emax);
    }
    __syncthreads();
    if (threadIdx.x >= 1 && threadIdx.x <= n) {                 // see k_emit_nodes
        const int L = threadIdx.x;
        const u32 m = ~((1u << (n - L + 1)) - 1u);
        const bool excl = J.drop_last && L == n;
        const u32 mn = excl ? s_mm[2] : s_mm[0], mx = excl ? s_mm[3] : s_mm[1];
        if (mn != 0xffffffffu) {
            atomicMin(&J.pos_min[L - 1], mn & m);
            atomicMax(&J.pos_max[L - 1], mx & m);
        }
    }
}

}  // namespace scp

// ==========================================================================================
// host side
// ==========================================================================================
using namespace scp;

struct scp_octree {
    DevBuf keys_a, keys_b, tiles_pts, tiles_sort, tiles_node, tiles_emit, frames, jobs, frame_begin, hist, desc, misc,
        tile_hist, job_tile_begin, n_lo, n_occ, n_parent, n_pos, n_fc, n_vdig;
    std::vector<JobDev> hjobs;
    std::vector<Tile> h_tiles_pts, h_tiles_sort, h_tiles_node, h_tiles_emit;
    std::vector<PassTile> h_tiles_pass;
    std::vector<int> pass_begin;          // tiles of pass p: h_tiles_pass[pass_begin[p] .. pass_begin[p+1])
    int max_depth = 0;
    DevBuf pass_desc, tiles_pass;
    int n_jobs = 0, mode = 0, P = 0, nt_frame = 0;
    long long total_keys = 0, total_nodes = 0, total_rows = 0, total_vox = 0;
    bool planned = false, emitted = false, host_gap = false;
    cudaEvent_t ev[10] = {};        // [8], [9]: around the host synchronisation inside the quantise stage (fused path)
    bool ev_ok = false;
    u64* sorted = nullptr;
};

static void build_tiles(const std::vector<long long>& counts, int tile, std::vector<Tile>& out, std::vector<int>* job_begin) {
    out.clear();
    if (job_begin) job_begin->clear();
    for (size_t j = 0; j < counts.size(); ++j) {
        int first = (int)out.size();
        if (job_begin) job_begin->push_back(first);
        for (long long b = 0; b < counts[j]; b += tile)
            out.push_back(Tile{(int)j, (int)b, (int)std::min<long long>(tile, counts[j] - b), first});
    }
    if (job_begin) job_begin->push_back((int)out.size());
}

// keys -> sorted (ping-pong with tmp).  hist_ready: the digit histograms were already produced (k_compact_hist).
// pmax > 0: per-job pass schedule (keys of job j start in `keys` or `tmp` as k_quantise_fused left them), P == pmax.
static int run_sort(u64* keys, u64* tmp, const Tile* d_tiles, int n_tiles, const JobDev* d_jobs, int n_jobs, int P,
                    bool hist_ready, int pmax, DevBuf& hist, DevBuf& desc, DevBuf& misc, cudaStream_t st, u64** result) {
    if (P <= 0 || n_tiles == 0) { *result = keys; return SCP_OK; }
    if (int e = desc.reserve((size_t)P * n_tiles * 256 * 4)) return e;
    if (int e = misc.reserve(64 * 4)) return e;
    SCP_CUDA(cudaMemsetAsync(desc.p, 0, (size_t)P * n_tiles * 256 * 4, st));
    SCP_CUDA(cudaMemsetAsync(misc.p, 0, 64 * 4, st));
    if (!hist_ready) {
        if (int e = hist.reserve((size_t)n_jobs * P * 256 * 4)) return e;
        SCP_CUDA(cudaMemsetAsync(hist.p, 0, (size_t)n_jobs * P * 256 * 4, st));
        k_sort_hist<<<n_tiles, TPB, 0, st>>>(keys, tmp, d_tiles, d_jobs, hist.as<u32>(), P, pmax);
        SCP_LAUNCHED();
    }
    k_scan_hist<<<n_jobs * P, 256, 0, st>>>(hist.as<u32>());
    SCP_LAUNCHED();
    for (int p = 0; p < P; ++p) {
        k_onesweep<<<n_tiles, TPB, 0, st>>>(keys, tmp, d_tiles, n_tiles, d_jobs, hist.as<u32>(), P, p, pmax, desc.as<u32>(),
                                            misc.as<u32>() + 1 + p, misc.as<u32>());
        SCP_LAUNCHED();
    }
    *result = (P & 1) ? tmp : keys;
    return SCP_OK;
}

// 0 = node records, all levels in one pass (k_emit_nodes + k_occupancy + k_context*); 1 = node records, one pass per level
// (k_level_pass); 2 (default) = as 0, but the lean outputs (occ, sym, ctx, pos_norm) come from
// k_tree_occ (occupancy + records in one warp-autonomous key pass) + k_context_lean.  env SCP_TREE = records | level | keys
static int tree_builder_default() {
    const char* e = getenv("SCP_TREE");
    if (e && !strcmp(e, "level")) return 1;
    if (e && !strcmp(e, "records")) return 0;
    return 2;
}
static int g_tree_builder = tree_builder_default();

static int launch_context_lean(int nt_n, const Tile* tiles, const JobDev* jobs, NodeArrays A, const scp_octree_out& O, cudaStream_t st) {
    // two nodes per thread at 40 registers, six blocks per SM: measured 1.35 ms per 131 M nodes against 1.75 ms with four nodes
    // per thread at 64 registers / four blocks (the gather chains are latency-bound: resident warps beat unrolling)
    k_context_lean<2, 6><<<nt_n, TPB, 0, st>>>(tiles, jobs, A, O);
    SCP_LAUNCHED();
    return SCP_OK;
}

extern "C" {

int scp_set_tree_builder(int mode) { int old = g_tree_builder; g_tree_builder = mode < 0 ? 0 : (mode > 2 ? 2 : mode); return old; }

scp_octree* scp_octree_create(void) { return new scp_octree(); }

void scp_octree_destroy(scp_octree* t) {
    if (!t) return;
    DevBuf* bufs[] = {&t->keys_a, &t->keys_b, &t->tiles_pts, &t->tiles_sort, &t->tiles_node, &t->tiles_emit, &t->frames, &t->jobs,
                      &t->frame_begin, &t->hist, &t->desc, &t->misc, &t->tile_hist, &t->job_tile_begin, &t->n_lo,
                      &t->n_occ, &t->n_parent, &t->n_pos, &t->n_fc, &t->n_vdig, &t->pass_desc, &t->tiles_pass};
    for (DevBuf* b : bufs) b->release();
    if (t->ev_ok) for (auto& e : t->ev) cudaEventDestroy(e);
    delete t;
}

int scp_octree_plan(scp_octree* t, const float* d_xyz, int point_stride, const int64_t* h_frame_offsets, int n_frames,
                    const scp_job* h_jobs, int n_jobs, int mode, void* stream) {
    SCP_REQUIRE(t && d_xyz && h_frame_offsets && h_jobs, "scp_octree_plan: null argument");
    SCP_REQUIRE(point_stride >= 3, "scp_octree_plan: point_stride must be >= 3");
    SCP_REQUIRE(n_frames > 0 && n_jobs > 0, "scp_octree_plan: empty batch");
    SCP_REQUIRE(mode >= 0 && mode <= 2, "scp_octree_plan: bad mode %d", mode);
    cudaStream_t st = as_stream(stream);
    t->planned = t->emitted = false;
    t->mode = mode; t->n_jobs = n_jobs;
    if (!t->ev_ok) { for (auto& e : t->ev) SCP_CUDA(cudaEventCreate(&e)); t->ev_ok = true; }

    std::vector<long long> fcount(n_frames), fbegin(n_frames);
    for (int f = 0; f < n_frames; ++f) {
        fbegin[f] = h_frame_offsets[f];
        fcount[f] = h_frame_offsets[f + 1] - h_frame_offsets[f];
        SCP_REQUIRE(fcount[f] > 0 && fcount[f] < (1ll << 30), "scp_octree_plan: frame %d has %lld points", f, fcount[f]);
    }
    t->hjobs.assign(n_jobs, JobDev{});
    std::vector<long long> jcount(n_jobs);
    long long kb = 0;
    for (int j = 0; j < n_jobs; ++j) {
        const scp_job& s = h_jobs[j];
        SCP_REQUIRE(s.frame >= 0 && s.frame < n_frames, "scp_octree_plan: job %d frame out of range", j);
        SCP_REQUIRE(s.path_len >= 0 && s.path_len <= 8, "scp_octree_plan: job %d path_len", j);
        SCP_REQUIRE(s.qs > 0, "scp_octree_plan: job %d qs must be > 0", j);
        JobDev& J = t->hjobs[j];
        J.frame = s.frame; J.path_len = s.path_len; J.path_bits = s.path_bits; J.drop_last = s.drop_last;
        J.qs = s.qs; J.cart_offset = s.cart_offset; J.lidar_level = s.lidar_level; J.pos_eps_last = s.pos_eps_last;
        J.pt_begin = fbegin[s.frame]; J.n_points = (int)fcount[s.frame]; J.key_begin = kb;
        jcount[j] = fcount[s.frame];
        kb += fcount[s.frame];
    }
    t->total_keys = kb;

    std::vector<Tile> ftiles;
    std::vector<int> jtb;
    build_tiles(fcount, TILE, ftiles, nullptr);
    build_tiles(jcount, TILE, t->h_tiles_pts, &jtb);
    build_tiles(jcount, SORT_TILE, t->h_tiles_sort, nullptr);
    const int nt_f = (int)ftiles.size(), nt_p = (int)t->h_tiles_pts.size(), nt_s = (int)t->h_tiles_sort.size();

    if (int e = t->keys_a.reserve(kb * 8)) return e;
    if (int e = t->keys_b.reserve(kb * 8)) return e;
    if (int e = t->tiles_pts.reserve((size_t)(nt_f + nt_p) * sizeof(Tile))) return e;
    if (int e = t->tiles_sort.reserve((size_t)nt_s * sizeof(Tile))) return e;
    if (int e = t->frames.reserve((size_t)n_frames * sizeof(FrameDev))) return e;
    if (int e = t->frame_begin.reserve((size_t)n_frames * 8)) return e;
    if (int e = t->jobs.reserve((size_t)n_jobs * sizeof(JobDev))) return e;
    if (int e = t->tile_hist.reserve((size_t)nt_p * CHUNKS * NBINS * 4)) return e;
    if (int e = t->job_tile_begin.reserve((size_t)(n_jobs + 1) * 4)) return e;
    Tile* d_ftiles = t->tiles_pts.as<Tile>();
    t->nt_frame = nt_f;
    Tile* d_ptiles = d_ftiles + nt_f;
    SCP_CUDA(cudaMemcpyAsync(d_ftiles, ftiles.data(), nt_f * sizeof(Tile), cudaMemcpyHostToDevice, st));
    SCP_CUDA(cudaMemcpyAsync(d_ptiles, t->h_tiles_pts.data(), nt_p * sizeof(Tile), cudaMemcpyHostToDevice, st));
    for (Tile& tl : t->h_tiles_sort) tl.kb = t->hjobs[tl.job].key_begin;
    SCP_CUDA(cudaMemcpyAsync(t->tiles_sort.p, t->h_tiles_sort.data(), nt_s * sizeof(Tile), cudaMemcpyHostToDevice, st));
    SCP_CUDA(cudaMemcpyAsync(t->frame_begin.p, fbegin.data(), n_frames * 8, cudaMemcpyHostToDevice, st));
    SCP_CUDA(cudaMemcpyAsync(t->jobs.p, t->hjobs.data(), n_jobs * sizeof(JobDev), cudaMemcpyHostToDevice, st));
    SCP_CUDA(cudaMemcpyAsync(t->job_tile_begin.p, jtb.data(), (n_jobs + 1) * 4, cudaMemcpyHostToDevice, st));
    JobDev* d_jobs = t->jobs.as<JobDev>();

    SCP_CUDA(cudaEventRecord(t->ev[0], st));
    if (mode != SCP_MODE_CART) {
        k_init_frames<<<(int)cdiv(n_frames, 256), 256, 0, st>>>(t->frames.as<FrameDev>(), n_frames);
        SCP_LAUNCHED();
        k_frame_stats<<<nt_f, TPB, 0, st>>>(d_xyz, point_stride, d_ftiles, t->frame_begin.as<long long>(),
                                            t->frames.as<FrameDev>(), mode);
        SCP_LAUNCHED();
    }
    k_job_setup<<<(int)cdiv(n_jobs, 128), 128, 0, st>>>(d_jobs, n_jobs, t->frames.as<FrameDev>(), mode);
    SCP_LAUNCHED();
    // frames with several jobs over the same points (mullevel) share the transform: one pass over the FRAME's points
    std::vector<int> fj_start(n_frames + 1, 0), fj(n_jobs);
    for (int j = 0; j < n_jobs; ++j) fj_start[h_jobs[j].frame + 1]++;
    int max_per_frame = 0;
    for (int f = 0; f < n_frames; ++f) { max_per_frame = std::max(max_per_frame, fj_start[f + 1]); fj_start[f + 1] += fj_start[f]; }
    {
        std::vector<int> fill(fj_start.begin(), fj_start.end() - 1);
        for (int j = 0; j < n_jobs; ++j) fj[fill[h_jobs[j].frame]++] = j;
    }
    // Fused path (spherical / cylindrical): depth from the frame statistics, then ONE kernel does transform + quantise +
    // morton_path filter + compaction for all jobs of a frame; the sort follows with a per-job pass schedule.
    bool fused = mode != SCP_MODE_CART && max_per_frame >= 1 && max_per_frame <= QF_MAXJ && !getenv("SCP_QUANT_OLD");
    int max_depth = 0, any_filter = 0;
    t->host_gap = false;
    if (fused) {
        SCP_CUDA(cudaEventRecord(t->ev[8], st));
        SCP_CUDA(cudaMemcpyAsync(t->hjobs.data(), d_jobs, n_jobs * sizeof(JobDev), cudaMemcpyDeviceToHost, st));
        SCP_CUDA(cudaStreamSynchronize(st));
        for (int j = 0; j < n_jobs; ++j) {
            fused = fused && t->hjobs[j].depth_known;
            max_depth = std::max(max_depth, t->hjobs[j].depth);
        }
    }
    if (fused) {
        SCP_REQUIRE(max_depth >= 1 && max_depth <= MAXL, "octree depth %d outside [1,%d]", max_depth, MAXL);
        t->P = (3 * max_depth + 7) / 8;
        t->max_depth = max_depth;
        for (Tile& tl : t->h_tiles_sort) tl.pj = (3 * t->hjobs[tl.job].depth + 7) / 8;      // per-job pass schedule
        SCP_CUDA(cudaMemcpyAsync(t->tiles_sort.p, t->h_tiles_sort.data(), nt_s * sizeof(Tile), cudaMemcpyHostToDevice, st));
        int *d_fjs = nullptr, *d_fj = nullptr;
        SCP_CUDA(upload_async((void**)&d_fjs, fj_start.data(), (size_t)(n_frames + 1) * 4, st));
        SCP_CUDA(upload_async((void**)&d_fj, fj.data(), (size_t)n_jobs * 4, st));
        SCP_CUDA(cudaEventRecord(t->ev[9], st));
        t->host_gap = true;
        k_job_prepare_fused<<<(int)cdiv(n_jobs, 128), 128, 0, st>>>(d_jobs, n_jobs);
        SCP_LAUNCHED();
        k_quantise_fused<<<nt_f, TPB, 0, st>>>(d_xyz, point_stride, d_ftiles, t->frame_begin.as<long long>(), d_fjs, d_fj, d_jobs,
                                               t->keys_a.as<u64>(), t->keys_b.as<u64>(), mode, t->P);
        SCP_LAUNCHED();
        SCP_CUDA(cudaFreeAsync(d_fjs, st));
        SCP_CUDA(cudaFreeAsync(d_fj, st));
        SCP_CUDA(cudaEventRecord(t->ev[1], st));
        if (int e = run_sort(t->keys_a.as<u64>(), t->keys_b.as<u64>(), t->tiles_sort.as<Tile>(), nt_s, d_jobs, n_jobs, t->P,
                             false, t->P, t->hist, t->desc, t->misc, st, &t->sorted)) return e;
    } else {
    if (mode != SCP_MODE_CART && max_per_frame >= 2 && max_per_frame <= QF_MAXJ) {
        int *d_fjs = nullptr, *d_fj = nullptr;
        SCP_CUDA(upload_async((void**)&d_fjs, fj_start.data(), (size_t)(n_frames + 1) * 4, st));
        SCP_CUDA(upload_async((void**)&d_fj, fj.data(), (size_t)n_jobs * 4, st));
        k_quantise_frames<<<nt_f, TPB, 0, st>>>(d_xyz, point_stride, d_ftiles, t->frame_begin.as<long long>(), d_fjs, d_fj, d_jobs,
                                                t->keys_a.as<u64>(), mode);
        SCP_LAUNCHED();
        SCP_CUDA(cudaFreeAsync(d_fjs, st));
        SCP_CUDA(cudaFreeAsync(d_fj, st));
    } else {
        k_quantise_keys<<<nt_p, TPB, 0, st>>>(d_xyz, point_stride, d_ptiles, d_jobs, t->keys_a.as<u64>(), mode);
        SCP_LAUNCHED();
    }
    k_job_depth<<<(int)cdiv(n_jobs, 128), 128, 0, st>>>(d_jobs, n_jobs);
    SCP_LAUNCHED();
    SCP_CUDA(cudaEventRecord(t->ev[1], st));
    SCP_CUDA(cudaMemcpyAsync(t->hjobs.data(), d_jobs, n_jobs * sizeof(JobDev), cudaMemcpyDeviceToHost, st));
    SCP_CUDA(cudaStreamSynchronize(st));
    max_depth = 0;
    for (int j = 0; j < n_jobs; ++j) {
        if (t->hjobs[j].overflow) { set_error("job %d: quantised coordinate outside [0, 2^21)", j); return SCP_ERR_RANGE; }
        max_depth = std::max(max_depth, t->hjobs[j].depth);
        any_filter |= t->hjobs[j].path_len > 0;
    }
    SCP_REQUIRE(max_depth >= 1 && max_depth <= MAXL, "octree depth %d outside [1,%d]", max_depth, MAXL);
    t->P = (3 * max_depth + 7) / 8;
    t->max_depth = max_depth;
    if (any_filter) {
        // jobs with a morton_path (mullevel): drop the rejected keys before sorting
        std::vector<int> stb;
        { std::vector<Tile> tmp_tiles; build_tiles(jcount, SORT_TILE, tmp_tiles, &stb); }
        int* d_stb = nullptr;
        u32* d_kept = nullptr;
        SCP_CUDA(upload_async((void**)&d_stb, stb.data(), stb.size() * 4, st));
        SCP_CUDA(malloc_async((void**)&d_kept, (size_t)(nt_s + 1) * 4, st));
        if (int e = t->hist.reserve((size_t)n_jobs * t->P * 256 * 4)) return e;
        SCP_CUDA(cudaMemsetAsync(t->hist.p, 0, (size_t)n_jobs * t->P * 256 * 4, st));
        k_filter_count<<<nt_s, TPB, 0, st>>>(t->keys_a.as<u64>(), t->tiles_sort.as<Tile>(), d_jobs, d_kept);
        SCP_LAUNCHED();
        k_kept_scan<<<n_jobs, 32, 0, st>>>(d_jobs, d_stb, d_kept);
        SCP_LAUNCHED();
        k_compact_hist<<<nt_s, TPB, 0, st>>>(t->keys_a.as<u64>(), t->keys_b.as<u64>(), t->tiles_sort.as<Tile>(), d_jobs, d_kept,
                                             t->hist.as<u32>(), t->P);
        SCP_LAUNCHED();
        SCP_CUDA(cudaFreeAsync(d_stb, st));
        SCP_CUDA(cudaFreeAsync(d_kept, st));
        if (int e = run_sort(t->keys_b.as<u64>(), t->keys_a.as<u64>(), t->tiles_sort.as<Tile>(), nt_s, d_jobs, n_jobs, t->P,
                             true, 0, t->hist, t->desc, t->misc, st, &t->sorted)) return e;
    } else {
        if (int e = run_sort(t->keys_a.as<u64>(), t->keys_b.as<u64>(), t->tiles_sort.as<Tile>(), nt_s, d_jobs, n_jobs, t->P,
                             false, 0, t->hist, t->desc, t->misc, st, &t->sorted)) return e;
    }
    }
    SCP_CUDA(cudaEventRecord(t->ev[2], st));
    k_head_hist<<<nt_p, TPB, 0, st>>>(t->sorted, d_ptiles, d_jobs, t->tile_hist.as<u32>());
    SCP_LAUNCHED();
    k_level_scan<<<n_jobs, 32, 0, st>>>(d_jobs, n_jobs, t->job_tile_begin.as<int>(), t->tile_hist.as<u32>());
    SCP_LAUNCHED();
    k_job_offsets<<<1, 1024, 0, st>>>(d_jobs, n_jobs);
    SCP_LAUNCHED();
    SCP_CUDA(cudaEventRecord(t->ev[3], st));
    u32 err = 0;
    SCP_CUDA(cudaMemcpyAsync(t->hjobs.data(), d_jobs, n_jobs * sizeof(JobDev), cudaMemcpyDeviceToHost, st));
    SCP_CUDA(cudaMemcpyAsync(&err, t->misc.p, 4, cudaMemcpyDeviceToHost, st));
    SCP_CUDA(cudaStreamSynchronize(st));
    if (err) { set_error("radix sort look-back timed out"); return SCP_ERR_INTERNAL; }
    for (int j = 0; j < n_jobs; ++j) {
        if (t->hjobs[j].overflow) { set_error("job %d: quantised coordinate outside [0, 2^21)", j); return SCP_ERR_RANGE; }
        if (t->hjobs[j].depth_known && (t->hjobs[j].qmax >> t->hjobs[j].depth) != 0u) {
            set_error("job %d: a quantised coordinate (%u) exceeds the depth %d derived from the frame statistics", j, t->hjobs[j].qmax,
                      t->hjobs[j].depth);
            return SCP_ERR_INTERNAL;
        }
    }
    t->total_nodes = t->total_rows = t->total_vox = 0;
    std::vector<long long> ncount(n_jobs);
    for (int j = 0; j < n_jobs; ++j) {
        t->total_nodes += t->hjobs[j].n_nodes; t->total_rows += t->hjobs[j].n_rows; t->total_vox += t->hjobs[j].n_voxels;
        ncount[j] = t->hjobs[j].n_nodes;
    }
    build_tiles(ncount, NODE_TILE, t->h_tiles_node, nullptr);
    // key tiles for k_emit_nodes: only the tiles that hold sorted keys (compacted jobs keep a third of them); `first` stays
    // the job's first tile in the point-tile numbering, which is how the per-tile level prefixes are indexed
    t->h_tiles_emit.clear();
    for (int j = 0; j < n_jobs; ++j)
        for (int b = 0; b < t->hjobs[j].n_kept; b += TILE)
            t->h_tiles_emit.push_back(Tile{j, b, std::min(TILE, t->hjobs[j].n_kept - b), jtb[j]});
    // exact tile lists of the level passes: pass p handles the children on level depth + 1 - p of every job
    t->h_tiles_pass.clear();
    t->pass_begin.assign(1, 0);
    for (int p = 0; p < max_depth; ++p) {
        for (int j = 0; j < n_jobs; ++j) {
            const JobDev& J = t->hjobs[j];
            const int n = J.depth, Lc = n + 1 - p;
            if (Lc < 2) continue;
            const bool vx = Lc == n + 1;
            const int Nc = vx ? J.n_kept : J.level_count[Lc - 1];
            const int first = (int)t->h_tiles_pass.size() - t->pass_begin[p];
            for (int b = 0; b < Nc; b += TILE) {
                PassTile pt{};
                pt.job = j; pt.begin = b; pt.count = std::min(TILE, Nc - b); pt.first = first;
                pt.n_child = Nc; pt.sh = 3 * (n - Lc + 1); pt.Lc = Lc; pt.depth = n;
                pt.lsp = (u32)J.level_start[Lc - 2]; pt.lsc = vx ? 0u : (u32)J.level_start[Lc - 1];
                pt.node0 = J.node_start; pt.src_off = J.key_begin; pt.vox_start = J.vox_start; pt.drop_last = (u32)J.drop_last;
                t->h_tiles_pass.push_back(pt);
            }
        }
        t->pass_begin.push_back((int)t->h_tiles_pass.size());
    }
    t->planned = true;
    return SCP_OK;
}

int scp_octree_job_info(const scp_octree* t, int job, scp_job_info* out) {
    SCP_REQUIRE(t && out && t->planned, "scp_octree_job_info: plan first");
    SCP_REQUIRE(job >= 0 && job < t->n_jobs, "scp_octree_job_info: job out of range");
    const JobDev& J = t->hjobs[job];
    memset(out, 0, sizeof(*out));
    out->depth = J.depth; out->n_points = J.n_points; out->n_voxels = J.n_voxels; out->n_rows = J.n_rows;
    out->row_start = J.row_start; out->voxel_start = J.vox_start;
    for (int l = 0; l < MAXL; ++l) {
        int c = J.level_count[l];
        if (J.drop_last && l == J.depth - 1 && c > 0) c -= 1;
        out->level_rows[l] = c;
        out->pos_min[l] = J.pos_min[l] == 0xffffffffu ? 0 : (int64_t)J.pos_min[l];
        out->pos_max[l] = (int64_t)J.pos_max[l];
    }
    out->bin_num = J.bin_num;
    for (int c = 0; c < 3; ++c) { out->steps[c] = J.step[c]; out->offset[c] = J.off[c]; }
    return SCP_OK;
}

int64_t scp_octree_total_rows(const scp_octree* t) { return t && t->planned ? t->total_rows : -1; }
int64_t scp_octree_total_voxels(const scp_octree* t) { return t && t->planned ? t->total_vox : -1; }
int64_t scp_octree_total_kept(const scp_octree* t) {
    if (!t || !t->planned) return -1;
    long long s = 0;
    for (const JobDev& J : t->hjobs) s += J.n_kept;
    return s;
}

int scp_octree_emit(scp_octree* t, const scp_octree_out* d_out, void* stream) {
    SCP_REQUIRE(t && d_out && t->planned, "scp_octree_emit: plan first");
    cudaStream_t st = as_stream(stream);
    const long long N = t->total_nodes + 1;
    const bool lean_out = !d_out->level && !d_out->octant && !d_out->parent && !d_out->pos && !d_out->ctx_pos && !d_out->rows_i64;
    if (g_tree_builder == 2 && lean_out) {
        // one key pass writes occupancy + node records; the lean context kernel gathers from them
        const int nt_e = (int)t->h_tiles_emit.size(), nt_n = (int)t->h_tiles_node.size();
        if (int e = t->n_lo.reserve(N * 2)) return e;
        if (int e = t->n_occ.reserve(N + 8)) return e;
        if (int e = t->n_parent.reserve(N * 4)) return e;
        if (int e = t->n_pos.reserve(N * 8)) return e;
        if (int e = t->tiles_emit.reserve((size_t)(nt_e + 1) * sizeof(Tile))) return e;
        if (int e = t->tiles_node.reserve((size_t)(nt_n + 1) * sizeof(Tile))) return e;
        SCP_CUDA(cudaMemcpyAsync(t->tiles_emit.p, t->h_tiles_emit.data(), nt_e * sizeof(Tile), cudaMemcpyHostToDevice, st));
        SCP_CUDA(cudaMemcpyAsync(t->tiles_node.p, t->h_tiles_node.data(), nt_n * sizeof(Tile), cudaMemcpyHostToDevice, st));
        SCP_CUDA(cudaMemsetAsync(t->n_occ.p, 0, (size_t)N + 8, st));
        NodeArrays A{t->n_lo.as<uint16_t>(), t->n_occ.as<uint8_t>(), t->n_parent.as<u32>(), t->n_pos.as<u64>(), nullptr, nullptr, 0};
        JobDev* dj = t->jobs.as<JobDev>();
        const bool rows = d_out->occ || d_out->sym || d_out->ctx || d_out->pos_norm;
        SCP_CUDA(cudaEventRecord(t->ev[4], st));
        if (nt_e) {
            // instruction-issue bound (80 % issue active): four blocks per SM at 56 registers measured faster than five or six
            u64* vk = reinterpret_cast<u64*>(d_out->voxel_key);
            if (rows) k_tree_occ<true><<<nt_e, TPB, 0, st>>>(t->sorted, t->tiles_emit.as<Tile>(), dj, t->tile_hist.as<u32>(), A, vk);
            else k_tree_occ<false><<<nt_e, TPB, 0, st>>>(t->sorted, t->tiles_emit.as<Tile>(), dj, t->tile_hist.as<u32>(), A, vk);
            SCP_LAUNCHED();
        }
        SCP_CUDA(cudaEventRecord(t->ev[5], st));
        SCP_CUDA(cudaEventRecord(t->ev[6], st));
        if (nt_n && rows) {
            if (int e = launch_context_lean(nt_n, t->tiles_node.as<Tile>(), dj, A, *d_out, st)) return e;
        }
        SCP_CUDA(cudaEventRecord(t->ev[7], st));
        t->emitted = true;
        return SCP_OK;
    }
    if (int e = t->n_lo.reserve(N * 2)) return e;
    if (int e = t->n_occ.reserve(N)) return e;
    if (int e = t->n_parent.reserve(N * 4)) return e;
    if (int e = t->n_pos.reserve(N * 8)) return e;
    // tree builder: 0 = all levels in one pass over the sorted keys (k_emit_nodes + k_occupancy, default: faster), 1 = one
    // pass per level, bottom-up (k_level_pass); scp_set_tree_builder() / env SCP_TREE=level
    const bool by_level = g_tree_builder == 1;
    u64* vox = reinterpret_cast<u64*>(d_out->voxel_key);          // optional output
    if (!by_level) {
        if (int e = t->n_fc.reserve(N * 4)) return e;
        if (int e = t->n_vdig.reserve(t->total_vox + 1)) return e;
    }
    const int nt_n = (int)t->h_tiles_node.size(), nt_e = (int)t->h_tiles_emit.size(), nt_p = (int)t->h_tiles_pass.size();
    if (int e = t->tiles_node.reserve((size_t)(nt_n + 1) * sizeof(Tile))) return e;
    SCP_CUDA(cudaMemcpyAsync(t->tiles_node.p, t->h_tiles_node.data(), nt_n * sizeof(Tile), cudaMemcpyHostToDevice, st));
    if (by_level) {
        if (int e = t->tiles_pass.reserve((size_t)(nt_p + 1) * sizeof(PassTile))) return e;
        SCP_CUDA(cudaMemcpyAsync(t->tiles_pass.p, t->h_tiles_pass.data(), (size_t)nt_p * sizeof(PassTile), cudaMemcpyHostToDevice, st));
    } else {
        if (int e = t->tiles_emit.reserve((size_t)(nt_e + 1) * sizeof(Tile))) return e;
        SCP_CUDA(cudaMemcpyAsync(t->tiles_emit.p, t->h_tiles_emit.data(), nt_e * sizeof(Tile), cudaMemcpyHostToDevice, st));
    }
    NodeArrays A{t->n_lo.as<uint16_t>(), t->n_occ.as<uint8_t>(), t->n_parent.as<u32>(), t->n_pos.as<u64>(), t->n_fc.as<u32>(),
                 t->n_vdig.as<uint8_t>(), 0};
    JobDev* d_jobs = t->jobs.as<JobDev>();
    const int n_pass = t->max_depth;                              // children levels depth+1 (voxels) ... 2
    if (by_level && nt_p) {
        const size_t words = (size_t)2 * nt_p + n_pass + 64;
        if (int e = t->pass_desc.reserve(words * 4)) return e;
        SCP_CUDA(cudaMemsetAsync(t->pass_desc.p, 0, words * 4, st));
        SCP_CUDA(cudaMemsetAsync(t->n_occ.p, 0, (size_t)N, st));
    }
    SCP_CUDA(cudaEventRecord(t->ev[4], st));
    if (by_level && nt_p) {
        u32* desc = t->pass_desc.as<u32>();
        u32* ticket = desc + (size_t)2 * nt_p;
        for (int p = 0; p < n_pass; ++p) {
            const int b0 = t->pass_begin[p], ntp = t->pass_begin[p + 1] - b0;
            if (ntp == 0) continue;
            if (p > 0) k_level_pass<0><<<ntp, TPB, 0, st>>>(t->sorted, t->tiles_pass.as<PassTile>() + b0, ntp, d_jobs, A, vox,
                                                            desc + (size_t)2 * b0, ticket + p, t->misc.as<u32>());
            else if (!vox) k_level_pass<1><<<ntp, TPB, 0, st>>>(t->sorted, t->tiles_pass.as<PassTile>() + b0, ntp, d_jobs, A, vox,
                                                                desc + (size_t)2 * b0, ticket + p, t->misc.as<u32>());
            else k_level_pass<2><<<ntp, TPB, 0, st>>>(t->sorted, t->tiles_pass.as<PassTile>() + b0, ntp, d_jobs, A, vox,
                                                      desc + (size_t)2 * b0, ticket + p, t->misc.as<u32>());
            SCP_LAUNCHED();
        }
    } else if (!by_level && nt_e) {
        k_emit_nodes<<<nt_e, TPB, 0, st>>>(t->sorted, t->tiles_emit.as<Tile>(), d_jobs, t->tile_hist.as<u32>(), A, vox);
        SCP_LAUNCHED();
    }
    SCP_CUDA(cudaEventRecord(t->ev[5], st));
    if (!by_level && nt_n) {
        k_occupancy<<<nt_n, TPB, 0, st>>>(t->tiles_node.as<Tile>(), d_jobs, A);
        SCP_LAUNCHED();
    }
    SCP_CUDA(cudaEventRecord(t->ev[6], st));
    if (nt_n) {
        if (lean_out) { if (int e = launch_context_lean(nt_n, t->tiles_node.as<Tile>(), d_jobs, A, *d_out, st)) return e; }
        else { k_context<<<nt_n, TPB, 0, st>>>(t->tiles_node.as<Tile>(), d_jobs, A, *d_out); SCP_LAUNCHED(); }
    }
    SCP_CUDA(cudaEventRecord(t->ev[7], st));
    t->emitted = true;
    return SCP_OK;
}

int scp_octree_finish(scp_octree* t, void* stream) {
    SCP_REQUIRE(t && t->emitted, "scp_octree_finish: emit first");
    cudaStream_t st = as_stream(stream);
    SCP_CUDA(cudaMemcpyAsync(t->hjobs.data(), t->jobs.p, t->n_jobs * sizeof(JobDev), cudaMemcpyDeviceToHost, st));
    SCP_CUDA(cudaStreamSynchronize(st));
    return SCP_OK;
}

int scp_octree_stage_ms(scp_octree* t, float out[6]) {
    SCP_REQUIRE(t && t->emitted, "scp_octree_stage_ms: emit first");
    SCP_CUDA(cudaEventSynchronize(t->ev[7]));
    const int a[6] = {0, 1, 2, 4, 5, 6}, b[6] = {1, 2, 3, 5, 6, 7};
    for (int i = 0; i < 6; ++i) SCP_CUDA(cudaEventElapsedTime(&out[i], t->ev[a[i]], t->ev[b[i]]));
    if (t->host_gap) {                   // the device is idle while the host reads the job depths: not kernel time
        float gap = 0.f;
        SCP_CUDA(cudaEventElapsedTime(&gap, t->ev[8], t->ev[9]));
        out[0] -= gap;
    }
    return SCP_OK;
}

int scp_segmented_sort_u64(uint64_t* d_keys, uint64_t* d_tmp, const int64_t* h_seg_offsets, int n_seg, int key_bits,
                           void* stream) {
    SCP_REQUIRE(d_keys && d_tmp && h_seg_offsets && n_seg > 0, "scp_segmented_sort_u64: bad argument");
    SCP_REQUIRE(key_bits > 0 && key_bits <= 64, "scp_segmented_sort_u64: key_bits");
    cudaStream_t st = as_stream(stream);
    std::vector<JobDev> jobs(n_seg, JobDev{});
    std::vector<long long> cnt(n_seg);
    for (int j = 0; j < n_seg; ++j) {
        jobs[j].key_begin = h_seg_offsets[j];
        cnt[j] = h_seg_offsets[j + 1] - h_seg_offsets[j];
        jobs[j].n_points = (int)cnt[j];
        jobs[j].n_kept = (int)cnt[j];
    }
    std::vector<Tile> tiles;
    build_tiles(cnt, SORT_TILE, tiles, nullptr);
    for (Tile& tl : tiles) tl.kb = jobs[tl.job].key_begin;
    DevBuf dj, dt, hist, desc, misc;
    int rc = SCP_OK;
    u64* res = nullptr;
    const int P = (key_bits + 7) / 8;
    do {
        if ((rc = dj.reserve(n_seg * sizeof(JobDev)))) break;
        if ((rc = dt.reserve((tiles.size() + 1) * sizeof(Tile)))) break;
        if (cudaMemcpyAsync(dj.p, jobs.data(), n_seg * sizeof(JobDev), cudaMemcpyHostToDevice, st) != cudaSuccess ||
            cudaMemcpyAsync(dt.p, tiles.data(), tiles.size() * sizeof(Tile), cudaMemcpyHostToDevice, st) != cudaSuccess) {
            set_error("scp_segmented_sort_u64: upload failed"); rc = SCP_ERR_CUDA; break;
        }
        rc = run_sort((u64*)d_keys, (u64*)d_tmp, dt.as<Tile>(), (int)tiles.size(), dj.as<JobDev>(), n_seg, P, false, 0, hist, desc,
                      misc, st, &res);
        if (rc) break;
        long long total = h_seg_offsets[n_seg] - h_seg_offsets[0];
        if (res != (u64*)d_keys &&
            cudaMemcpyAsync(d_keys + h_seg_offsets[0], res + h_seg_offsets[0], total * 8, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
            set_error("scp_segmented_sort_u64: copy back failed"); rc = SCP_ERR_CUDA; break;
        }
        u32 err = 0;
        if (cudaMemcpyAsync(&err, misc.p, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess) { set_error("scp_segmented_sort_u64: sync failed"); rc = SCP_ERR_CUDA; break; }
        if (err) { set_error("radix sort look-back timed out"); rc = SCP_ERR_INTERNAL; }
    } while (0);
    cudaStreamSynchronize(st);
    dj.release(); dt.release(); hist.release(); desc.release(); misc.release();
    return rc;
}

}  // extern "C"

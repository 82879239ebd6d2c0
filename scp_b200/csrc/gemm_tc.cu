// tcgen05 / TMEM / TMA GEMM engine for nn.Linear:  Y[M,N] = act(X[M,K] W[N,K]^T + b) (+ R)   (sm_100a only)
//
// Both operands are K-major (row-major with K contiguous), which is exactly tcgen05's "K-major A, K-major B"
// form, so no transposes are needed.  Persistent, warp-specialised CTA (384 threads):
//   warp 0   TMA producer   cp.async.bulk.tensor.2d, 128B-swizzled [128 x 32] A tiles and [BN x 32] B tiles
//   warp 1   MMA issuer     one lane issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=BN, K=8), fp32 accumulate in TMEM
//   warp 2   TMEM allocator (2 accumulator stages x BN columns, so the epilogue of tile i overlaps the MMAs of tile i+1)
//   warps 4-11 epilogue     tcgen05.ld 32x32b.x32 -> bias / activation -> smem tile -> residual -> coalesced global stores
//                           (two warps per TMEM lane quarter: the epilogue is instruction-bound, esp. with erf-GELU)
// smem ring of STAGES x (A 16 KB + B BN*128 B) with full/empty mbarriers; tcgen05.commit releases stages.
// Operands are read as fp32 and rounded to TF32 by the tensor core (10-bit mantissa), accumulation is fp32.
//
// SPLIT = true is the error-compensated "3xTF32" mode that keeps fp32-class accuracy (needed for the 1e-3 PMF
// parity bound): x = x_hi + x_lo with x_hi = x & 0xffffe000 (exactly representable in TF32) and
//     D += A_hi B_hi + A_lo B_hi + A_hi B_lo          (the dropped A_lo B_lo term is ~2^-22 relative).
// W_hi / W_lo are split once per weight matrix (cached); A tiles are split in shared memory by warps 2-3 between the
// TMA arrival and the MMA (generic-proxy writes + fence.proxy.async), so activations are still read once from HBM.
#include <mutex>
#include <stdlib.h>
#include <unordered_map>
#include <stdio.h>
#include <cuda_fp16.h>
#include "tc.cuh"

namespace scp {

// erf for the GELU epilogue: Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7 absolute, the size of erff's own rounding) with
// one reciprocal and one exp2 on the special-function unit: 13 instructions instead of erff's ~25 -- the N = 1024 GELU layer
// is bound by its epilogue (20 % of all instructions of the kernel were erff).
__device__ __forceinline__ float tc_erf(float x) {
    const float ax = fabsf(x);
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.0f)));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(ax * ax * -1.4426950408889634f));
    return copysignf(fmaf(-p * t, e, 1.0f), x);
}
// Packed fp32 pairs (Blackwell FFMA2 / FMUL2 / FADD2: one issue slot for two IEEE-rounded results, bit-identical to the scalar
// instructions).  The GELU epilogue of the N = 1024 layer is issue-bound (ncu: 58 % issue-active, eligible warps not selected
// 0.65 per issue): the pair form needs ~8 issue slots per element instead of ~18.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
// GELU of two values, operation for operation the scalar 0.5 v (1 + tc_erf(v / sqrt 2)) above (the polynomial is evaluated with
// negated coefficients, which negates every intermediate exactly, so that -p t needs no separate sign flip)
__device__ __forceinline__ f32x2 tc_gelu2(f32x2 v) {
    const f32x2 x = mul2(v, pk2(0.70710678118654752440f, 0.70710678118654752440f));
    float x0, x1;
    upk2(x, x0, x1);
    const f32x2 ax = pk2(fabsf(x0), fabsf(x1));
    float d0, d1, t0, t1;
    upk2(fma2(pk2(0.3275911f, 0.3275911f), ax, pk2(1.0f, 1.0f)), d0, d1);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d1));
    const f32x2 t = pk2(t0, t1);
    f32x2 p = fma2(pk2(-1.061405429f, -1.061405429f), t, pk2(1.453152027f, 1.453152027f));
    p = fma2(p, t, pk2(-1.421413741f, -1.421413741f));
    p = fma2(p, t, pk2(0.284496736f, 0.284496736f));
    p = fma2(p, t, pk2(-0.254829592f, -0.254829592f));                       // = -p of the scalar form
    float a0, a1, e0, e1;
    upk2(mul2(mul2(ax, ax), pk2(-1.4426950408889634f, -1.4426950408889634f)), a0, a1);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
    float r0, r1;
    upk2(fma2(mul2(p, t), pk2(e0, e1), pk2(1.0f, 1.0f)), r0, r1);
    const f32x2 erf1 = add2(pk2(1.0f, 1.0f), pk2(copysignf(r0, x0), copysignf(r1, x1)));
    return mul2(mul2(pk2(0.5f, 0.5f), v), erf1);
}
__device__ __forceinline__ float tc_act(float v, int act) {
    switch (act) {
        case SCP_ACT_LEAKY001: return v > 0.f ? v : 0.01f * v;
        case SCP_ACT_GELU: return 0.5f * v * (1.0f + tc_erf(v * 0.70710678118654752440f));
        case SCP_ACT_RELU: return v > 0.f ? v : 0.f;
        default: return v;
    }
}

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
template <int BN, int STAGES, bool SPLIT>
__global__ void __launch_bounds__(384, 1) k_gemm_tf32(const __grid_constant__ CUtensorMap tmA,
                                                       const __grid_constant__ CUtensorMap tmB,
                                                       const __grid_constant__ CUtensorMap tmBlo,
                                                       const float* __restrict__ bias, const float* __restrict__ R,
                                                       long long ldr, float* __restrict__ Y, long long ldy, long long M, int N,
                                                       int K, int act) {
    constexpr int BM = 128, BK = 32;
    constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4;
    constexpr int STAGE_BYTES = SPLIT ? 2 * (A_BYTES + B_BYTES) : (A_BYTES + B_BYTES);   // [A | B] or [A_hi | A_lo | B_hi | B_lo]
    constexpr int TX_BYTES = SPLIT ? (A_BYTES + 2 * B_BYTES) : (A_BYTES + B_BYTES);
    constexpr int OFF_ALO = A_BYTES, OFF_B = SPLIT ? 2 * A_BYTES : A_BYTES, OFF_BLO = OFF_B + B_BYTES;
    constexpr uint32_t TMEM_COLS = 2 * BN;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by an OFFSET from the shared-space symbol: a pointer rebuilt from an integer would be generic
    // (LD/ST instead of LDS/STS and no alias information against global memory)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 2;
    uint64_t* ready = tempty + 2;                        // SPLIT: A tile split done (2 splitter warps)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ready + STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles_n = (N + BN - 1) / BN;
    const long long n_tiles = ((M + BM - 1) / BM) * n_tiles_n;
    const int n_kb = (K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&ready[s], 2); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        int stage = 0; uint32_t phase = 0;                  // all lanes run the loop, the elected lane issues (uniform operands)
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int m_blk = (int)(tile / n_tiles_n), n_blk = (int)(tile % n_tiles_n);
            for (int kb = 0; kb < n_kb; ++kb) {
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* a = smem + stage * STAGE_BYTES;
                if (elect_one()) {
                    mbar_expect_tx(&full[stage], TX_BYTES);
                    tma_load_2d(a, &tmA, &full[stage], kb * BK, m_blk * BM);
                    tma_load_2d(a + OFF_B, &tmB, &full[stage], kb * BK, n_blk * BN);
                    if (SPLIT) tma_load_2d(a + OFF_BLO, &tmBlo, &full[stage], kb * BK, n_blk * BN);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // all lanes run the loop (warp-uniform operands -> uniform registers), the elected lane issues; see tc.cuh elect_one()
        // instruction descriptor: D=f32, A=B=tf32, both K-major, N = BN, M = 128
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            mbar_wait(&tempty[acc], acc_phase ^ 1);
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
            for (int kb = 0; kb < n_kb; ++kb) {
                mbar_wait(SPLIT ? &ready[stage] : &full[stage], phase);
                tc_fence_after();
                const uint8_t* a = smem + stage * STAGE_BYTES;
                const uint64_t da = make_smem_desc(a), db = make_smem_desc(a + OFF_B);
                const uint64_t dal = make_smem_desc(a + OFF_ALO), dbl = make_smem_desc(a + OFF_BLO);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) {     // 8 tf32 = 32 bytes = 2 descriptor units per MMA
                        const uint64_t o = (uint64_t)(2 * k);
                        tc_mma_tf32(d_tmem, da + o, db + o, idesc, (kb | k) ? 1u : 0u);
                        if (SPLIT) {
                            tc_mma_tf32(d_tmem, dal + o, db + o, idesc, 1u);
                            tc_mma_tf32(d_tmem, da + o, dbl + o, idesc, 1u);
                        }
                    }
                    tc_commit(&empty[stage]);              // frees the smem stage when these MMAs retire
                    if (kb == n_kb - 1) tc_commit(&tfull[acc]);   // accumulator ready for the epilogue
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (SPLIT && (warp == 2 || warp == 3)) {
        // split the freshly landed A tile in place: A <- A_hi, A_lo tile next to it (same swizzled layout)
        const int tsp = (warp - 2) * 32 + lane;            // 0..63
        int stage = 0; uint32_t phase = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int kb = 0; kb < n_kb; ++kb) {
                mbar_wait(&full[stage], phase);
                uint4* hi = reinterpret_cast<uint4*>(smem + stage * STAGE_BYTES);
                uint4* lo = reinterpret_cast<uint4*>(smem + stage * STAGE_BYTES + OFF_ALO);
#pragma unroll 4
                for (int i = tsp; i < A_BYTES / 16; i += 64) {
                    uint4 v = hi[i], h, l;
                    h.x = v.x & 0xffffe000u; h.y = v.y & 0xffffe000u; h.z = v.z & 0xffffe000u; h.w = v.w & 0xffffe000u;
                    l.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(h.x));
                    l.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(h.y));
                    l.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(h.z));
                    l.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(h.w));
                    hi[i] = h; lo[i] = l;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to tcgen05.mma
                __syncwarp();
                if (lane == 0) mbar_arrive(&ready[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // Epilogue.  tcgen05.ld gives lane i the 32 consecutive columns of ROW i, but a row-per-lane global store touches
        // 32 different cache lines per instruction (the L1 pipe then costs more than the MMAs of a K=256 tile).  So the
        // chunk goes through a per-warp [32][36] shared-memory tile: bias + activation on the way in (lane = row), residual
        // add + 128-byte coalesced stores on the way out (8 lanes per row, 4 rows per instruction).
        const int w = (warp - 4) & 3;                      // TMEM lane quarter this warp may read (warp id mod 4)
        const int half = (warp - 4) >> 2;                  // two warps per quarter: even / odd 32-column chunks
        float* stg = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 512) + (warp - 4) * (32 * 32);
        int acc = 0; uint32_t acc_phase = 0;
        const bool vec_ok = (ldy % 4 == 0) && ((reinterpret_cast<uintptr_t>(Y) & 15) == 0) &&
                            (!R || ((ldr % 4 == 0) && ((reinterpret_cast<uintptr_t>(R) & 15) == 0))) &&
                            (!bias || ((reinterpret_cast<uintptr_t>(bias) & 15) == 0));
        const int sub_r = lane >> 3, sub_c = (lane & 7) << 2;
        constexpr int NCH = BN / 64;                       // 32-column chunks per warp (BN = 64: one chunk, half 1 idle on odd)
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int m_blk = (int)(tile / n_tiles_n), n_blk = (int)(tile % n_tiles_n);
            const long long row0 = (long long)m_blk * BM + w * 32;
            const uint32_t t_row = tmem_base + ((uint32_t)(w * 32) << 16) + (uint32_t)(acc * BN);
            const bool rows_ok = row0 < M;
            bool waited = false;
#pragma unroll
            for (int p = 0; p < (NCH + 1) / 2; ++p) {      // two chunks per round
                const int c0 = half * 32 + 128 * p, c1 = c0 + 64;
                const int n0 = n_blk * BN + c0, n1 = n_blk * BN + c1;
                const bool have0 = c0 < BN && n0 < N && rows_ok, have1 = c1 < BN && n1 < N && rows_ok;
                const bool vec0 = have0 && vec_ok && n0 + 32 <= N, vec1 = have1 && vec_ok && n1 + 32 <= N;
                // residual rows and bias of both chunks first (HBM latency), before the accumulator wait / TMEM loads
                float4 q0[8], q1[8], b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
                if (R) {
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const long long row = row0 + it * 4 + sub_r;
                        q0[it] = (vec0 && row < M) ? __ldcs(reinterpret_cast<const float4*>(R + row * ldr + n0 + sub_c)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        q1[it] = (vec1 && row < M) ? __ldcs(reinterpret_cast<const float4*>(R + row * ldr + n1 + sub_c)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                if (bias) {
                    if (vec0) b0 = __ldg(reinterpret_cast<const float4*>(bias + n0 + sub_c));
                    if (vec1) b1 = __ldg(reinterpret_cast<const float4*>(bias + n1 + sub_c));
                }
                if (!waited) { mbar_wait(&tfull[acc], acc_phase); tc_fence_after(); waited = true; }
                uint32_t r0[32], r1[32];
                if (c0 < BN) tc_ld32_nowait(t_row + (uint32_t)c0, r0);
                if (c1 < BN) tc_ld32_nowait(t_row + (uint32_t)c1, r1);
                tc_wait_ld();
                if (p == (NCH + 1) / 2 - 1) {               // accumulator is in registers: hand the TMEM stage back now
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_relaxed(&tempty[acc]);
                }
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const uint32_t* r = hh ? r1 : r0;
                    const float4* q = hh ? q1 : q0;
                    const float4 b4 = hh ? b1 : b0;
                    const int nn = hh ? n1 : n0;
                    if (!(hh ? have1 : have0)) continue;       // warp-uniform
                    if (hh ? vec1 : vec0) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            *reinterpret_cast<uint4*>(stg + lane * 32 + ((((j >> 2) ^ lane) & 7) << 2)) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
                        __syncwarp();
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const int rr = it * 4 + sub_r;
                            const long long row = row0 + rr;
                            if (row < M) {
                                float4 v = *reinterpret_cast<const float4*>(stg + rr * 32 + ((((sub_c >> 2) ^ rr) & 7) << 2));
                                v.x = tc_act(v.x + b4.x, act); v.y = tc_act(v.y + b4.y, act);
                                v.z = tc_act(v.z + b4.z, act); v.w = tc_act(v.w + b4.w, act);
                                if (R) { v.x += q[it].x; v.y += q[it].y; v.z += q[it].z; v.w += q[it].w; }
                                __stcs(reinterpret_cast<float4*>(Y + row * ldy + nn + sub_c), v);
                            }
                        }
                        __syncwarp();
                    } else {
                        const long long row = row0 + lane;
                        if (row < M) {
                            float* yrow = Y + row * ldy + nn;
                            const float* rrow = R ? R + row * ldr + nn : nullptr;
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                if (nn + j < N) {
                                    float v = __uint_as_float(r[j]);
                                    if (bias) v += bias[nn + j];
                                    v = tc_act(v, act);
                                    if (rrow) v += rrow[j];
                                    yrow[j] = v;
                                }
                            }
                        }
                    }
                }
            }
            if (!waited) {                                   // BN = 64 and this warp's half has no chunk: still part of the hand-shake
                mbar_wait(&tfull[acc], acc_phase);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_relaxed(&tempty[acc]);
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// x0, x1 -> packed fp16 pairs (x0 in the low half): hi = fp16(x) with saturation, lo = fp16(x - hi)
__device__ __forceinline__ void split_f16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
    float h0, h1;
    asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(h0), "=f"(h1) : "r"(hi));
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(x1 - h1), "f"(x0 - h0));
}

// ---------------------------------------------------------------------------------------------
// 3xTF32 kernel with the activation operand in TENSOR MEMORY ("TS" form of tcgen05.mma)
// ---------------------------------------------------------------------------------------------
// k_gemm_tf32<.., SPLIT> above keeps A_hi and A_lo in shared memory.  Per 32-wide K block the shared-memory port then
// carries the TMA writes (48 KB), the in-place split (16 KB read + 32 KB written) and the operand reads of twelve SS-form
// MMAs (12 x (4 KB A + 4 KB B) = 96 KB): 192 KB at 128 B/clk = 1500 clk against 768 clk of tensor work -- the kernel is
// shared-memory-bound at ~50 % tensor utilisation (ncu: 45-57 %).  Here four splitter warps (one TMEM lane quarter each,
// thread = row) read the landed fp32 A tile once and write A_hi / A_lo straight to TMEM with tcgen05.st, and the MMAs
// take A from TMEM: TMA 48 KB + split read 16 KB + B reads 48 KB = 112 KB = 875 clk per K block.  A_lo no longer
// occupies shared memory, so the ring is 4 stages deep instead of 3.
// Warps (512 threads, 128 registers): 0 TMA producer, 1 MMA issuer, 2 TMEM owner, 4-7 splitters, 8-15 epilogue.
// TMEM (512 columns): accumulators [0, 2 BN) | stage s: A_hi [256 + 64 s, +32), A_lo [+32, +64).
//
// H = true is the same kernel on the FP16 pipe ("3xFP16"): x = x_hi + x_lo with two fp16 numbers (11 + 11 mantissa bits like
// the tf32 split), kind::f16 MMAs at twice the tf32 rate, K blocks of 64 (two fp32 TMA boxes of A; W_hi/W_lo are stored as
// fp16, 64 per 128-byte row), A_hi/A_lo packed two per TMEM column (even k in the low half).  fp16 has 5 exponent bits:
// the weights are scaled by a power of two per matrix so that max|W| sits at 2^13..2^14 (the epilogue multiplies by the
// exact inverse, read from `oscale`), the activations are taken as they are -- |x| < 65504 converts with saturation,
// and what falls below the fp16 subnormal step (2^-24) is an ABSOLUTE error of 3e-8 per product term, far below the
// 2^-22 relative error of the split itself for the O(1) activations of this model.
//
// CL > 1 (A-stationary order only): a thread-block CLUSTER of CL CTAs walks the N tiles of CL different M blocks in lock-step and
// shares the weight stream -- every CTA fetches 1/CL of each W_hi / W_lo tile and TMA-multicasts it into the same stage of all
// CL CTAs.  The K <= 256 layers are bound by L2 bandwidth, not by the tensor pipe: every 128-row M block streams the whole
// weight matrix (hi + lo) from L2, 4.3 GB per call of the N = 1024 layer against 2.5 GB of activations, and the kernel time
// follows the L2 byte count at ~7.5 TB/s with or without the output stores (tools/exp_gemm_time.py, SCP_GEMM_DBG runs of
// round 2).  A stage may be refilled when the MMAs of ALL CL CTAs have read it: empty[] counts CL multicast commits.
template <int BN, int STAGES, bool H, int CL = 1>
__global__ void __launch_bounds__(512, 1) k_gemm_x3_ts(const __grid_constant__ CUtensorMap tmA,
                                                        const __grid_constant__ CUtensorMap tmB,
                                                        const __grid_constant__ CUtensorMap tmBlo,
                                                        const float* __restrict__ bias, const float* __restrict__ R,
                                                        long long ldr, float* __restrict__ Y, long long ldy, long long M, int N,
                                                        int K, int act, const float* __restrict__ oscale, int a_stationary,
                                                        long long* __restrict__ trace) {
    constexpr int BM = 128, BK = H ? 64 : 32;
    constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * 128;               // B rows: 32 tf32 or 64 fp16 = 128 bytes
    constexpr int STAGE_BYTES = A_BYTES + 2 * B_BYTES;                     // [A fp32 | B_hi | B_lo]
    constexpr int OFF_B = A_BYTES, OFF_BLO = A_BYTES + B_BYTES;
    constexpr uint32_t TMEM_COLS = 512, T_A = 256;
    static_assert(2 * BN <= 256 && STAGES * 64 <= 256, "TMEM budget");
    // A-stationary order: the ring is cut into stages of ONE operand K block each -- an A block (fp32, only for the first N tile of
    // an M block) or a [W_hi | W_lo] block -- instead of [A | W_hi | W_lo] stages whose A third stays empty for 5-7 of every 6-8
    // tiles.  Same shared memory, twice the stages in flight: the mainloop of the K <= 256 layers was bound by the latency of
    // the weight stream (one K block every ~950 cycles whether 12 MMAs or 4 were issued on it, and whether or not clusters halved
    // the L2 reads: three 32 KB loads in flight against ~2500 cycles from TMA issue to the stage's release).
    constexpr int AS_STAGE_BYTES = A_BYTES > 2 * B_BYTES ? A_BYTES : 2 * B_BYTES;
    constexpr int AS_STAGES = (STAGES * STAGE_BYTES) / AS_STAGE_BYTES;
    constexpr int NBAR = AS_STAGES > STAGES ? AS_STAGES : STAGES;
    static_assert((3 * NBAR + 4) * 8 + 8 <= 512, "barrier block");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty = full + NBAR;
    uint64_t* tfull = empty + NBAR;
    uint64_t* tempty = tfull + 2;
    uint64_t* ready = tempty + 2;                        // A_hi / A_lo of the stage (A-stationary: of K block kb) are in TMEM (4 splitter warps)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ready + NBAR);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles_n = (N + BN - 1) / BN;
    const long long n_tiles = ((M + BM - 1) / BM) * n_tiles_n;
    const int n_kb = (K + BK - 1) / BK;
    // Tile order.  Default: tiles round-robin over the CTAs.  A-stationary (K <= 4 K blocks, many M blocks): a CTA takes
    // whole M blocks and walks their N tiles back to back; A is loaded and split ONCE per M block and stays in TMEM
    // (columns [T_A + 64 kb, +64)) while only the weight tiles stream -- a 128x128 tile otherwise refetches and re-splits
    // its fp32 A block (128 KB at K = 256) for each of the 6-8 N tiles of the layer, which made the K = 256 layers
    // L2 -> shared-memory bound (8.4 TB/s of L2 reads) and kept the splitter warps as busy as the tensor pipe.
    const long long n_mblk = (M + BM - 1) / BM;
    const int cl_rank = CL > 1 ? (int)cluster_ctarank() : 0;
    constexpr uint16_t CL_MASK = (uint16_t)((1u << CL) - 1u);
    auto tile_of = [&](long long it, int& m_blk, int& n_blk) -> bool {
        if (CL > 1) {
            // cluster c takes the M blocks [CL (c + r n_clusters), +CL) in round r; a block past the end (odd tail) is a phantom:
            // its TMA boxes are zero-filled and its stores masked, and the CTA keeps the cluster's barriers in step
            const long long first = ((long long)(blockIdx.x / CL) + (it / n_tiles_n) * (gridDim.x / CL)) * CL;
            m_blk = (int)(first + cl_rank); n_blk = (int)(it % n_tiles_n);
            return first < n_mblk;
        }
        if (a_stationary) {
            const long long m = (long long)blockIdx.x + (it / n_tiles_n) * gridDim.x;
            m_blk = (int)m; n_blk = (int)(it % n_tiles_n);
            return m < n_mblk;
        }
        const long long tile = (long long)blockIdx.x + it * gridDim.x;
        m_blk = (int)(tile / n_tiles_n); n_blk = (int)(tile % n_tiles_n);
        return tile < n_tiles;
    };

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmBlo)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < NBAR; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CL); mbar_init(&ready[s], 4); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                        // every CTA's barriers exist before a peer's TMA or commit signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0 && a_stationary) {
        int stage = 0; uint32_t phase = 0;                  // all lanes run the loop, the elected lane issues (uniform operands)
        int m_blk, n_blk;
        for (long long it = 0; tile_of(it, m_blk, n_blk); ++it) {
            for (int kb = 0; kb < n_kb; ++kb) {
                if (n_blk == 0) {                           // the M block's activations: one stage per K block, consumed by the splitters
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* a = smem + stage * AS_STAGE_BYTES;
                    if (elect_one()) {
                        mbar_expect_tx(&full[stage], A_BYTES);
                        tma_load_2d(a, &tmA, &full[stage], kb * BK, m_blk * BM);
                        if (H) tma_load_2d(a + BM * 128, &tmA, &full[stage], kb * BK + 32, m_blk * BM);      // second 32-wide fp32 box
                    }
                    __syncwarp();
                    if (++stage == AS_STAGES) { stage = 0; phase ^= 1; }
                }
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* b = smem + stage * AS_STAGE_BYTES;
                if (elect_one()) {
                    mbar_expect_tx(&full[stage], 2 * B_BYTES);
                    if (CL > 1) {                          // this CTA's 1/CL of the weight tile, to every CTA of the cluster
                        constexpr int ROWS = BN / CL;
                        tma_load_2d_mc(b + cl_rank * ROWS * 128, &tmB, &full[stage], kb * BK, n_blk * BN + cl_rank * ROWS, CL_MASK);
                        tma_load_2d_mc(b + B_BYTES + cl_rank * ROWS * 128, &tmBlo, &full[stage], kb * BK, n_blk * BN + cl_rank * ROWS, CL_MASK);
                    } else {
                        tma_load_2d(b, &tmB, &full[stage], kb * BK, n_blk * BN);
                        tma_load_2d(b + B_BYTES, &tmBlo, &full[stage], kb * BK, n_blk * BN);
                    }
                }
                __syncwarp();
                if (++stage == AS_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 0) {
        // streaming order (K > 256, or too few M blocks): every tile loads [A | W_hi | W_lo] stages
        int stage = 0; uint32_t phase = 0;                  // all lanes run the loop, the elected lane issues (uniform operands)
        int m_blk, n_blk;
        for (long long it = 0; tile_of(it, m_blk, n_blk); ++it) {
            for (int kb = 0; kb < n_kb; ++kb) {
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* a = smem + stage * STAGE_BYTES;
                if (elect_one()) {
                    mbar_expect_tx(&full[stage], STAGE_BYTES);
                    tma_load_2d(a, &tmA, &full[stage], kb * BK, m_blk * BM);
                    if (H) tma_load_2d(a + BM * 128, &tmA, &full[stage], kb * BK + 32, m_blk * BM);   // second 32-wide fp32 box
                    tma_load_2d(a + OFF_B, &tmB, &full[stage], kb * BK, n_blk * BN);
                    tma_load_2d(a + OFF_BLO, &tmBlo, &full[stage], kb * BK, n_blk * BN);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && a_stationary) {
        // all lanes run the loop (warp-uniform operands -> uniform registers), the elected lane issues; see tc.cuh elect_one()
        const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);                        // f16 x f16 -> f32
        int stage = 0; uint32_t phase = 0, mb_phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        int m_blk, n_blk;
        int tr_n = 0;                                      // development aid (SCP_GEMM_TRACE=1): clock stamps of CTA 0's MMA warp
        const bool TR = trace && blockIdx.x == 0;
        for (long long it = 0; tile_of(it, m_blk, n_blk); ++it) {
            if (TR && lane == 0 && tr_n < 480) trace[tr_n++] = clock64();            // tile: before the accumulator wait
            mbar_wait(&tempty[acc], acc_phase ^ 1);
            if (TR && lane == 0 && tr_n < 480) trace[tr_n++] = clock64();            // tile: accumulator free
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
            for (int kb = 0; kb < n_kb; ++kb) {
                if (n_blk == 0) {
                    // A_hi / A_lo of this K block are in TMEM: the splitters have read the A stage, hand it back to the producer
                    // (in EVERY CTA of the cluster: the stage's next use may be a weight block that the peers multicast into it)
                    mbar_wait(&ready[kb], mb_phase);
                    // -- and through a multicast COMMIT, not a plain remote arrive: the arrivals of this CTA must reach a peer's
                    // barrier in the order of the stage's uses, and the commits of its earlier weight blocks are still in flight)
                    if (elect_one()) {
                        if (CL > 1) tc_commit_mc(&empty[stage], CL_MASK);
                        else mbar_arrive(&empty[stage]);
                    }
                    __syncwarp();
                    if (++stage == AS_STAGES) { stage = 0; phase ^= 1; }
                }
                mbar_wait(&full[stage], phase);
                if (TR && lane == 0 && tr_n < 480) trace[tr_n++] = clock64();        // K block: operands ready
                tc_fence_after();
                const uint8_t* b = smem + stage * AS_STAGE_BYTES;
                const uint64_t db = make_smem_desc(b), dbl = make_smem_desc(b + B_BYTES);
                const uint32_t ah = tmem_base + T_A + (uint32_t)(kb * 64), al = ah + 32u;
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {          // 16 fp16 = 8 TMEM columns of A = 2 descriptor units of B per MMA
                        const uint64_t o = (uint64_t)(2 * k);
                        tc_mma_f16_ts(d_tmem, ah + 8u * k, db + o, idesc, (kb | k) ? 1u : 0u);
                        tc_mma_f16_ts(d_tmem, al + 8u * k, db + o, idesc, 1u);
                        tc_mma_f16_ts(d_tmem, ah + 8u * k, dbl + o, idesc, 1u);
                    }
                    if (CL > 1) tc_commit_mc(&empty[stage], CL_MASK);   // in every CTA of the cluster (their TMA writes this stage too)
                    else tc_commit(&empty[stage]);         // frees the stage when these MMAs retire
                    if (kb == n_kb - 1) tc_commit(&tfull[acc]);   // accumulator ready for the epilogue
                }
                __syncwarp();
                if (++stage == AS_STAGES) { stage = 0; phase ^= 1; }
            }
            if (n_blk == n_tiles_n - 1) mb_phase ^= 1;
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp == 1) {
        // all lanes run the loop (warp-uniform operands -> uniform registers), the elected lane issues; see tc.cuh elect_one()
        const uint32_t idesc = H ? ((1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24))                  // f16 x f16 -> f32
                                 : ((1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24));
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        int m_blk, n_blk;
        int tr_n = 0;                                      // development aid (SCP_GEMM_TRACE=1): clock stamps of CTA 0's MMA warp
        const bool TR = trace && blockIdx.x == 0;
        for (long long it = 0; tile_of(it, m_blk, n_blk); ++it) {
            if (TR && lane == 0 && tr_n < 480) trace[tr_n++] = clock64();            // tile: before the accumulator wait
            mbar_wait(&tempty[acc], acc_phase ^ 1);
            if (TR && lane == 0 && tr_n < 480) trace[tr_n++] = clock64();            // tile: accumulator free
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
            for (int kb = 0; kb < n_kb; ++kb) {
                mbar_wait(&ready[stage], phase);
                if (TR && lane == 0 && tr_n < 480) trace[tr_n++] = clock64();        // K block: operands ready
                tc_fence_after();
                const uint8_t* a = smem + stage * STAGE_BYTES;
                const uint64_t db = make_smem_desc(a + OFF_B), dbl = make_smem_desc(a + OFF_BLO);
                const uint32_t ah = tmem_base + T_A + (uint32_t)(stage * 64), al = ah + 32u;
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {          // 8 tf32 / 16 fp16 = 8 TMEM columns of A = 2 descriptor units of B per MMA
                        const uint64_t o = (uint64_t)(2 * k);
                        if (H) {
                            tc_mma_f16_ts(d_tmem, ah + 8u * k, db + o, idesc, (kb | k) ? 1u : 0u);
                            tc_mma_f16_ts(d_tmem, al + 8u * k, db + o, idesc, 1u);
                            tc_mma_f16_ts(d_tmem, ah + 8u * k, dbl + o, idesc, 1u);
                        } else {
                            tc_mma_tf32_ts(d_tmem, ah + 8u * k, db + o, idesc, (kb | k) ? 1u : 0u);
                            tc_mma_tf32_ts(d_tmem, al + 8u * k, db + o, idesc, 1u);
                            tc_mma_tf32_ts(d_tmem, ah + 8u * k, dbl + o, idesc, 1u);
                        }
                    }
                    if (CL > 1) tc_commit_mc(&empty[stage], CL_MASK);   // ... in every CTA of the cluster (their TMA writes this stage too)
                    else
                    tc_commit(&empty[stage]);              // frees the smem stage and its TMEM A slot when these MMAs retire
                    if (kb == n_kb - 1) tc_commit(&tfull[acc]);   // accumulator ready for the epilogue
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= 4 && warp < 8 && a_stationary) {
        // splitter, A-stationary order: only the first N tile of an M block has A stages; thread = row of the A tile
        if constexpr (H) {
            const int row = (warp - 4) * 32 + lane;
            const uint32_t t_lane = tmem_base + ((uint32_t)((warp - 4) * 32) << 16) + T_A;
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            int m_blk, n_blk;
            for (long long it = 0; tile_of(it, m_blk, n_blk); ++it) {
                if (n_blk == 0) {
                    // (the resident A of the previous M block was read by the MMAs of its last tile: this warp has seen that
                    // tile's accumulator complete, below)
                    for (int kb = 0; kb < n_kb; ++kb) {
                        mbar_wait(&full[stage], phase);
                        const uint8_t* a = smem + stage * AS_STAGE_BYTES + row * 128;
                        uint32_t hi[32], lo[32];
                        // 64 fp32 of the row (two boxes) -> 32 + 32 packed fp16 pairs, even k in the low half
#pragma unroll
                        for (int c = 0; c < 16; ++c) {
                            const float4 v = *reinterpret_cast<const float4*>(a + (c >> 3) * (BM * 128) + (((c & 7) ^ (row & 7)) << 4));
                            split_f16x2(v.x, v.y, hi[2 * c], lo[2 * c]);
                            split_f16x2(v.z, v.w, hi[2 * c + 1], lo[2 * c + 1]);
                        }
                        tc_st32(t_lane + (uint32_t)(kb * 64), hi);
                        tc_st32(t_lane + (uint32_t)(kb * 64) + 32u, lo);
                        tc_wait_st();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&ready[kb]);           // (release: the stage's shared-memory reads are done as well)
                        if (++stage == AS_STAGES) { stage = 0; phase ^= 1; }      // the A stage ...
                        if (++stage == AS_STAGES) { stage = 0; phase ^= 1; }      // ... and the weight stage behind it
                    }
                } else {
                    for (int kb = 0; kb < n_kb; ++kb)
                        if (++stage == AS_STAGES) { stage = 0; phase ^= 1; }
                }
                // Follow EVERY tile's accumulator barrier: a parity wait only tells "the phase before the current one is over", so a
                // warp that skipped the five later tiles of an M block and then waited for the last one would be let through by the
                // first (and overwrite the resident A under the MMAs that still read it).
                mbar_wait(&tfull[acc], acc_phase);
                tc_fence_after();
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // splitter: thread = row of the A tile; its 128 bytes sit in 8 swizzled 16-byte chunks of the row's line
        const int row = (warp - 4) * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)((warp - 4) * 32) << 16) + T_A;
        int stage = 0; uint32_t phase = 0;
        int m_blk, n_blk;
        for (long long it = 0; tile_of(it, m_blk, n_blk); ++it) {
            for (int kb = 0; kb < n_kb; ++kb) {
                mbar_wait(&full[stage], phase);
                const uint8_t* a = smem + stage * STAGE_BYTES + row * 128;
                uint32_t hi[32], lo[32];
                if (H) {
                    // 64 fp32 of the row (two boxes) -> 32 + 32 packed fp16 pairs, even k in the low half
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        const float4 v = *reinterpret_cast<const float4*>(a + (c >> 3) * (BM * 128) + (((c & 7) ^ (row & 7)) << 4));
                        split_f16x2(v.x, v.y, hi[2 * c], lo[2 * c]);
                        split_f16x2(v.z, v.w, hi[2 * c + 1], lo[2 * c + 1]);
                    }
                } else {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint4 v = *reinterpret_cast<const uint4*>(a + ((c ^ (row & 7)) << 4));
                    hi[4 * c] = v.x & 0xffffe000u; hi[4 * c + 1] = v.y & 0xffffe000u;
                    hi[4 * c + 2] = v.z & 0xffffe000u; hi[4 * c + 3] = v.w & 0xffffe000u;
                    lo[4 * c] = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(hi[4 * c]));
                    lo[4 * c + 1] = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(hi[4 * c + 1]));
                    lo[4 * c + 2] = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(hi[4 * c + 2]));
                    lo[4 * c + 3] = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(hi[4 * c + 3]));
                }
                }
                tc_st32(t_lane + (uint32_t)(stage * 64), hi);
                tc_st32(t_lane + (uint32_t)(stage * 64) + 32u, lo);
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_relaxed(&ready[stage]);    // TMEM writes are ordered by wait::st + the tcgen05 fence; nothing in
                                                                         // generic memory to publish, and a release arrive costs a MEMBAR
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 8) {
        // epilogue: as in k_gemm_tf32, one 32-column chunk at a time (128-register budget)
        const int w = (warp - 8) & 3;                      // TMEM lane quarter this warp may read (warp id mod 4)
        const int half = (warp - 8) >> 2;                  // two warps per quarter: even / odd 32-column chunks
        float* stg = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 512) + (warp - 8) * (32 * 32);
        int acc = 0; uint32_t acc_phase = 0;
        const bool vec_ok = (ldy % 4 == 0) && ((reinterpret_cast<uintptr_t>(Y) & 15) == 0) &&
                            (!R || ((ldr % 4 == 0) && ((reinterpret_cast<uintptr_t>(R) & 15) == 0))) &&
                            (!bias || ((reinterpret_cast<uintptr_t>(bias) & 15) == 0));
        const float osc = H ? __ldg(oscale) : 1.0f;
        const int sub_r = lane >> 3, sub_c = (lane & 7) << 2;
        const bool ETR = trace && blockIdx.x == 0 && warp == 8 && lane == 0;
        int e_n = 0;
        int m_blk, n_blk;
        for (long long it = 0; tile_of(it, m_blk, n_blk); ++it) {
            const long long row0 = (long long)m_blk * BM + w * 32;
            const uint32_t t_row = tmem_base + ((uint32_t)(w * 32) << 16) + (uint32_t)(acc * BN);
            bool waited = false;
#pragma unroll
            for (int c0 = half * 32; c0 < BN; c0 += 64) {
                const int n0 = n_blk * BN + c0;
                const bool have = n0 < N && row0 < M;
                const bool vec = have && vec_ok && n0 + 32 <= N;
                float4 q[8], b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                const int n_row = (int)(M - row0 < 32 ? (M - row0 > 0 ? M - row0 : 0) : 32);       // rows of this warp's quarter inside the matrix
                if (R) {
                    const float* rp = R + (row0 + sub_r) * ldr + n0 + sub_c;
                    const long long rstep = 4 * ldr;
#pragma unroll
                    for (int it = 0; it < 8; ++it, rp += rstep)
                        q[it] = (vec && it * 4 + sub_r < n_row) ? __ldcs(reinterpret_cast<const float4*>(rp)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                if (bias && vec) b4 = __ldg(reinterpret_cast<const float4*>(bias + n0 + sub_c));
                if (ETR && e_n < 960) trace[1024 + e_n++] = clock64();                  // chunk: before the accumulator wait
                if (!waited) { mbar_wait(&tfull[acc], acc_phase); tc_fence_after(); waited = true; }
                if (ETR && e_n < 960) trace[1024 + e_n++] = clock64();                  // chunk: accumulator ready
                uint32_t r[32];
                tc_ld32(t_row + (uint32_t)c0, r);
                if (ETR && e_n < 960) trace[1024 + e_n++] = clock64();                  // chunk: scores in registers
                if (H) {
                    const f32x2 osc2 = pk2(osc, osc);
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {                                                        // exact: power of two
                        float a, b;
                        upk2(mul2(pk2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), osc2), a, b);
                        r[j] = __float_as_uint(a); r[j + 1] = __float_as_uint(b);
                    }
                }
                if (c0 + 64 >= BN) {                          // last chunk of this warp: the TMEM stage can go back now
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_relaxed(&tempty[acc]);
                }
                if (!have) continue;                          // warp-uniform
                if (vec) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<uint4*>(stg + lane * 32 + ((((j >> 2) ^ lane) & 7) << 2)) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
                    __syncwarp();
                    // the activation is chosen ONCE per chunk: a `switch (act)` per element was a third of the epilogue's
                    // instructions, and the eight epilogue warps set the pace of the K = 256 layers (MMA-warp trace: ~4500 of
                    // 9000 cycles per tile spent waiting for an accumulator to come back)
                    // (packed pairs: FADD2 / FMUL2 / FFMA2, same roundings as the scalar form)
                    const f32x2 b01 = pk2(b4.x, b4.y), b23 = pk2(b4.z, b4.w);
                    auto store_rows = [&](auto actf2) {
                        float* yp = Y + (row0 + sub_r) * ldy + n0 + sub_c;       // one 64-bit address, then a constant step per iteration
                        const long long ystep = 4 * ldy;
                        auto row4 = [&](int it, float* dst) {
                            const int rr = it * 4 + sub_r;
                            float4 v = *reinterpret_cast<const float4*>(stg + rr * 32 + ((((sub_c >> 2) ^ rr) & 7) << 2));
                            f32x2 lo2 = actf2(add2(pk2(v.x, v.y), b01)), hi2 = actf2(add2(pk2(v.z, v.w), b23));
                            if (R) { lo2 = add2(lo2, pk2(q[it].x, q[it].y)); hi2 = add2(hi2, pk2(q[it].z, q[it].w)); }
                            upk2(lo2, v.x, v.y); upk2(hi2, v.z, v.w);
                            __stcs(reinterpret_cast<float4*>(dst), v);
                        };
                        if (n_row == 32) {                      // warp-uniform; ONE basic block: the eight rows' dependent chains (LDS -> ~14
                                                                // packed ops + 2 MUFU -> STG, ~180 cycles each) interleave instead of running
                                                                // one after the other behind a per-row bounds branch
#pragma unroll
                            for (int it = 0; it < 8; ++it, yp += ystep) row4(it, yp);
                        } else {
#pragma unroll
                            for (int it = 0; it < 8; ++it, yp += ystep)
                                if (it * 4 + sub_r < n_row) row4(it, yp);
                        }
                    };
                    auto each = [](f32x2 v, auto f) { float a, b; upk2(v, a, b); return pk2(f(a), f(b)); };
                    switch (act) {                          // warp-uniform
                        case SCP_ACT_GELU: store_rows([](f32x2 v) { return tc_gelu2(v); }); break;
                        case SCP_ACT_LEAKY001: store_rows([&](f32x2 v) { return each(v, [](float x) { return x > 0.f ? x : 0.01f * x; }); }); break;
                        case SCP_ACT_RELU: store_rows([&](f32x2 v) { return each(v, [](float x) { return fmaxf(x, 0.f); }); }); break;
                        default: store_rows([](f32x2 v) { return v; }); break;
                    }
                    __syncwarp();
                    if (ETR && e_n < 960) trace[1024 + e_n++] = clock64();              // chunk: stored
                } else {
                    const long long row = row0 + lane;
                    if (row < M) {
                        float* yrow = Y + row * ldy + n0;
                        const float* rrow = R ? R + row * ldr + n0 : nullptr;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if (n0 + j < N) {
                                float v = __uint_as_float(r[j]);
                                if (bias) v += bias[n0 + j];
                                v = tc_act(v, act);
                                if (rrow) v += rrow[j];
                                yrow[j] = v;
                            }
                        }
                    }
                }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                        // no CTA leaves while a peer may still multicast into it or signal its barriers
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// host: tensor maps (cached) + launch
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

struct MapKey {
    const void* p; long long ld; long long rows; int cols; int box_rows; int f16;
    bool operator==(const MapKey& o) const { return p == o.p && ld == o.ld && rows == o.rows && cols == o.cols && box_rows == o.box_rows && f16 == o.f16; }
};
struct MapHash {
    size_t operator()(const MapKey& k) const {
        size_t h = reinterpret_cast<size_t>(k.p);
        h ^= (size_t)k.ld * 0x9e3779b97f4a7c15ull; h ^= (size_t)k.rows * 0xc2b2ae3d27d4eb4full;
        h ^= ((size_t)k.cols << 20) ^ (size_t)k.box_rows ^ ((size_t)k.f16 << 50);
        return h;
    }
};

static int get_tensor_map_2d_t(const void* p, long long ld, long long rows, int cols, int box_rows, int f16, CUtensorMap* out);
int get_tensor_map_2d(const float* p, long long ld, long long rows, int cols, int box_rows, CUtensorMap* out) {
    return get_tensor_map_2d_t(p, ld, rows, cols, box_rows, 0, out);
}
int get_tensor_map_2d_f16(const void* p, long long ld, long long rows, int cols, int box_rows, CUtensorMap* out) {
    return get_tensor_map_2d_t(p, ld, rows, cols, box_rows, 1, out);
}
// f16 = 0: float32 elements, boxes of 32 columns; f16 = 1: fp16 elements, boxes of 64 columns (128-byte rows either way)
static int get_tensor_map_2d_t(const void* p, long long ld, long long rows, int cols, int box_rows, int f16, CUtensorMap* out) {
    static std::unordered_map<MapKey, CUtensorMap, MapHash> cache;
    static std::mutex mu;
    MapKey key{p, ld, rows, cols, box_rows, f16};
    {
        std::lock_guard<std::mutex> g(mu);
        auto it = cache.find(key);
        if (it != cache.end()) { *out = it->second; return SCP_OK; }
    }
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return SCP_ERR_CUDA; }
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * (f16 ? 2 : 4)};
    cuuint32_t box[2] = {f16 ? 64u : 32u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = enc(out, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(p), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%d ld=%lld", (int)r, rows, cols, ld); return SCP_ERR_CUDA; }
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 4096) cache.clear();
    cache[key] = *out;
    return SCP_OK;
}

// W -> [W_hi ; W_lo] (two [N,K] matrices back to back), cached per weight pointer
__global__ void __launch_bounds__(256) k_split_weights(const float* __restrict__ w, long long n, float* __restrict__ hi,
                                                        float* __restrict__ lo) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const float v = w[i];
        const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        hi[i] = h; lo[i] = v - h;
    }
}

// W -> scale 2^e with max|W| 2^e in [2^13, 2^14), then [W_hi ; W_lo] as fp16 ([N,K] each) + the inverse scale (float) behind
// them; one block computes the maximum (weights are small and this runs once per matrix)
__global__ void __launch_bounds__(1024) k_weight_scale(const float* __restrict__ w, long long n, float* __restrict__ scales) {
    __shared__ float sm[32];
    float m = 0.f;
    for (long long i = threadIdx.x; i < n; i += 1024) m = fmaxf(m, fabsf(w[i]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = warp_max(sm[threadIdx.x]);
        if (threadIdx.x == 0) {
            int e = 0;
            if (m > 0.f && m < 3e38f) { frexpf(m, &e); e = 14 - e; }          // m = f * 2^(14-e), f in [0.5, 1)
            e = max(-100, min(100, e));
            scales[0] = ldexpf(1.0f, e);                                       // applied to W
            scales[1] = ldexpf(1.0f, -e);                                      // applied to the accumulators
        }
    }
}
__global__ void __launch_bounds__(256) k_split_weights_f16(const float* __restrict__ w, long long n, const float* __restrict__ scales,
                                                            __half* __restrict__ hi, __half* __restrict__ lo) {
    const float s = scales[0];
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const float v = w[i] * s;
        const __half h = __float2half_rn(v);
        hi[i] = h; lo[i] = __float2half_rn(v - __half2float(h));
    }
}

struct WKey { const void* p; int n, k; bool operator==(const WKey& o) const { return p == o.p && n == o.n && k == o.k; } };
struct WHash { size_t operator()(const WKey& k) const { return reinterpret_cast<size_t>(k.p) ^ ((size_t)k.n << 40) ^ ((size_t)k.k << 20); } };
static std::unordered_map<WKey, float*, WHash> g_wsplit;
static std::mutex g_wmu;

static std::unordered_map<WKey, __half*, WHash> g_wsplit_h;

void gemm_cache_clear() {
    std::lock_guard<std::mutex> g(g_wmu);
    for (auto& kv : g_wsplit) cudaFree(kv.second);
    g_wsplit.clear();
    for (auto& kv : g_wsplit_h) cudaFree(kv.second);
    g_wsplit_h.clear();
}

// drops the cached splits of one weight matrix (any shape registered under this pointer)
void gemm_cache_drop(const void* w) {
    std::lock_guard<std::mutex> g(g_wmu);
    for (auto it = g_wsplit.begin(); it != g_wsplit.end();) {
        if (it->first.p == w) { cudaFree(it->second); it = g_wsplit.erase(it); } else ++it;
    }
    for (auto it = g_wsplit_h.begin(); it != g_wsplit_h.end();) {
        if (it->first.p == w) { cudaFree(it->second); it = g_wsplit_h.erase(it); } else ++it;
    }
}

// fp16 split: [hi N*K halfs | lo N*K halfs | pad to 16 bytes | scale, 1/scale (floats)]
static int get_weight_split_f16(const float* w, int N, int K, cudaStream_t st, __half** out, const float** oscale) {
    std::lock_guard<std::mutex> g(g_wmu);
    WKey key{w, N, K};
    const long long n = (long long)N * K;
    const size_t off_scale = (size_t)((2 * n * 2 + 15) / 16) * 16;
    auto it = g_wsplit_h.find(key);
    if (it == g_wsplit_h.end()) {
        __half* buf = nullptr;
        SCP_CUDA(cudaMalloc((void**)&buf, off_scale + 16));
        float* sc = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(buf) + off_scale);
        k_weight_scale<<<1, 1024, 0, st>>>(w, n, sc);
        SCP_LAUNCHED();
        k_split_weights_f16<<<(unsigned)std::min<long long>(cdiv(n, 256), 1184), 256, 0, st>>>(w, n, sc, buf, buf + n);
        SCP_LAUNCHED();
        it = g_wsplit_h.emplace(key, buf).first;
    }
    *out = it->second;
    *oscale = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(it->second) + off_scale) + 1;
    return SCP_OK;
}

static int get_weight_split(const float* w, int N, int K, cudaStream_t st, float** out) {
    std::lock_guard<std::mutex> g(g_wmu);
    WKey key{w, N, K};
    auto it = g_wsplit.find(key);
    if (it != g_wsplit.end()) { *out = it->second; return SCP_OK; }
    float* buf = nullptr;
    const long long n = (long long)N * K;
    SCP_CUDA(cudaMalloc((void**)&buf, 2 * n * 4));
    k_split_weights<<<(unsigned)std::min<long long>(cdiv(n, 256), 1184), 256, 0, st>>>(w, n, buf, buf + n);
    SCP_LAUNCHED();
    g_wsplit[key] = buf;
    *out = buf;
    return SCP_OK;
}

bool linear_tf32_ok(long long ldx, long long ldy, long long M, int N, int K, const void* x, const void* w, const void* y) {
    (void)ldy; (void)y;
    if (M < 1 || N < 8 || K < 32) return false;
    if (K % 4 || ldx % 4) return false;
    if (x && (reinterpret_cast<uintptr_t>(x) & 15)) return false;
    if (w && (reinterpret_cast<uintptr_t>(w) & 15)) return false;
    return true;
}

template <int BN, int STAGES, bool SPLIT>
static int launch(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mbl, const float* bias, const float* res,
                  long long ldr, float* y, long long ldy, long long M, int N, int K, int act, cudaStream_t st) {
    constexpr int smem = STAGES * (128 * 128 + BN * 128) * (SPLIT ? 2 : 1) + 1024 + 512 + 8 * 32 * 32 * 4;
    static bool attr = false;
    if (!attr) {
        SCP_CUDA(cudaFuncSetAttribute(k_gemm_tf32<BN, STAGES, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = true;
    }
    static int n_sm = 0;
    if (!n_sm) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); if (n_sm <= 0) n_sm = 148; }
    const long long tiles = cdiv(M, 128) * cdiv(N, BN);
    const int grid = (int)std::min<long long>(tiles, n_sm);
    k_gemm_tf32<BN, STAGES, SPLIT><<<grid, 384, smem, st>>>(ma, mb, mbl, bias, res, ldr, y, ldy, M, N, K, act);
    SCP_LAUNCHED();
    return SCP_OK;
}

int g_gemm_cluster = getenv("SCP_GEMM_CL") ? atoi(getenv("SCP_GEMM_CL")) : 2;     // CTAs per cluster of the K <= 256 layers (1 | 2 | 4)

// Cluster launch of the A-stationary kernel (CL CTAs share the weight stream by TMA multicast).  `mb` / `mbl` have boxes of
// BN / CL rows.  Returns SCP_OK, or a negative value when the cluster cannot be scheduled (the caller falls back to CL = 1).
template <int BN, int STAGES, int CL>
static int launch_ts_cl(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mbl, const float* bias, const float* res,
                        long long ldr, float* y, long long ldy, long long M, int N, int K, int act, cudaStream_t st, const float* oscale) {
    constexpr int smem = STAGES * (128 * 256 + 2 * BN * 128) + 1024 + 512 + 8 * 32 * 32 * 4;
    auto kern = k_gemm_x3_ts<BN, STAGES, true, CL>;
    static int n_clusters = 0;                              // co-resident clusters of CL CTAs (one CTA per SM)
    if (!n_clusters) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) { cudaGetLastError(); n_clusters = -1; }
        else {
            cudaLaunchConfig_t q = {};
            q.gridDim = dim3(CL * 64); q.blockDim = dim3(512); q.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            q.attrs = at; q.numAttrs = 1;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, kern, &q) != cudaSuccess || n <= 0) { cudaGetLastError(); n_clusters = -1; }
            else n_clusters = n;
        }
    }
    if (n_clusters <= 0) return -1;
    const long long n_mblk = cdiv(M, 128);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(CL * std::min<long long>(cdiv(n_mblk, CL), n_clusters)));
    cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    long long* no_trace = nullptr;
    SCP_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mb, mbl, bias, res, ldr, y, ldy, M, N, K, act, oscale, 1, no_trace));
    SCP_LAUNCHED();
    return SCP_OK;
}

template <int BN, int STAGES, bool H>
static int launch_ts(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mbl, const float* bias, const float* res,
                     long long ldr, float* y, long long ldy, long long M, int N, int K, int act, cudaStream_t st,
                     const float* oscale = nullptr) {
    constexpr int smem = STAGES * (128 * (H ? 256 : 128) + 2 * BN * 128) + 1024 + 512 + 8 * 32 * 32 * 4;
    static_assert(smem <= 227 * 1024, "shared memory budget");
    static bool attr = false;
    if (!attr) {
        SCP_CUDA(cudaFuncSetAttribute(k_gemm_x3_ts<BN, STAGES, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = true;
    }
    static int n_sm = 0;
    if (!n_sm) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); if (n_sm <= 0) n_sm = 148; }
    const long long tiles = cdiv(M, 128) * cdiv(N, BN), n_mblk = cdiv(M, 128);
    // A-stationary tile order (see the kernel) when A fits the 256 spare TMEM columns and there are M blocks to spare
    // (SCP_GEMM_AS=0 restores the round-robin order.  While the epilogue set the pace it made no difference; with the
    // epilogue fixed it takes another 3 ms per step off nn.Linear)
    static const bool as_off = getenv("SCP_GEMM_AS") && atoi(getenv("SCP_GEMM_AS")) == 0;
    const int a_stationary = (H && !as_off && K <= 256 && cdiv(N, BN) >= 2 && n_mblk >= 2 * n_sm) ? 1 : 0;
    const int grid = (int)std::min<long long>(a_stationary ? n_mblk : tiles, n_sm);
    static long long* d_trace = nullptr;
    static const bool want_trace = getenv("SCP_GEMM_TRACE") != nullptr;
    if (want_trace && !d_trace) { cudaMalloc(&d_trace, 2048 * 8); }
    if (want_trace) cudaMemsetAsync(d_trace, 0, 2048 * 8, st);
    k_gemm_x3_ts<BN, STAGES, H><<<grid, 512, smem, st>>>(ma, mb, mbl, bias, res, ldr, y, ldy, M, N, K, act, oscale, a_stationary,
                                                         want_trace ? d_trace : nullptr);
    if (want_trace && M > 100000) {                      // development aid: MMA-warp timeline of CTA 0, first tiles
        long long hh[2048];
        cudaStreamSynchronize(st);
        cudaMemcpy(hh, d_trace, sizeof(hh), cudaMemcpyDeviceToHost);
        const int n_kb = (K + (H ? 64 : 32) - 1) / (H ? 64 : 32), per = 2 + n_kb;
        printf("GEMM TRACE M=%lld N=%d K=%d act=%d as=%d: per tile [wait_acc | kb waits... | tile total]\n", M, N, K, act, a_stationary);
        for (int t = 2; t < 14 && hh[(t + 1) * per]; ++t) {
            const long long* q = hh + t * per;
            printf("  tile %2d: acc_wait %5lld |", t, q[1] - q[0]);
            for (int kb = 0; kb < n_kb; ++kb) printf(" %5lld", q[2 + kb] - q[1 + kb]);
            printf(" | total %6lld\n", q[per] - q[0]);
        }
        printf("  epilogue warp 8, per chunk [wait_acc | tcgen05.ld | stage+store | gap to next chunk]\n");
        for (int c = 8; c < 20 && hh[1024 + 4 * (c + 1)]; ++c) {
            const long long* q = hh + 1024 + 4 * c;
            printf("  chunk %2d: %6lld %6lld %6lld %6lld\n", c, q[1] - q[0], q[2] - q[1], q[3] - q[2], q[4] - q[3]);
        }
    }
    SCP_LAUNCHED();
    return SCP_OK;
}

int linear_tf32(const float* x, long long ldx, const float* w, const float* bias, const float* res, long long ldr, float* y,
                long long ldy, long long M, int N, int K, int act, cudaStream_t st, int split) {
    CUtensorMap ma, mb, mbl;
    if (int e = get_tensor_map_2d(x, ldx, M, K, 128, &ma)) return e;
    if (split == 2 && K % 8 == 0) {                                       // 3xFP16 (fp16 rows need 16-byte strides)
        __half* wh = nullptr;
        const float* osc = nullptr;
        if (int e = get_weight_split_f16(w, N, K, st, &wh, &osc)) return e;
        const int BN = N > 64 ? 128 : 64;
        // K <= 256 layers with many M blocks: clusters share the weight stream (SCP_GEMM_CL = 1 | 2 | 4; default 2)
        const int cl = g_gemm_cluster;
        static int n_sm = 0;
        if (!n_sm) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); if (n_sm <= 0) n_sm = 148; }
        if (BN == 128 && cl > 1 && K <= 256 && N % 128 == 0 && N >= 256 && cdiv(M, 128) >= 2 * n_sm) {
            const int rows = 128 / (cl >= 4 ? 4 : 2);
            if (int e = get_tensor_map_2d_t(wh, K, N, K, rows, 1, &mb)) return e;
            if (int e = get_tensor_map_2d_t(wh + (long long)N * K, K, N, K, rows, 1, &mbl)) return e;
            const int r = cl >= 4 ? launch_ts_cl<128, 3, 4>(ma, mb, mbl, bias, res, ldr, y, ldy, M, N, K, act, st, osc)
                                  : launch_ts_cl<128, 3, 2>(ma, mb, mbl, bias, res, ldr, y, ldy, M, N, K, act, st, osc);
            if (r >= 0) return r;                            // < 0: clusters cannot be scheduled here -> single-CTA kernel below
        }
        if (int e = get_tensor_map_2d_t(wh, K, N, K, BN, 1, &mb)) return e;
        if (int e = get_tensor_map_2d_t(wh + (long long)N * K, K, N, K, BN, 1, &mbl)) return e;
        if (BN == 128) return launch_ts<128, 3, true>(ma, mb, mbl, bias, res, ldr, y, ldy, M, N, K, act, st, osc);
        return launch_ts<64, 3, true>(ma, mb, mbl, bias, res, ldr, y, ldy, M, N, K, act, st, osc);
    }
    if (split) {
        float* ws = nullptr;
        if (int e = get_weight_split(w, N, K, st, &ws)) return e;
        const int BN = N > 64 ? 128 : 64;
        if (int e = get_tensor_map_2d(ws, K, N, K, BN, &mb)) return e;
        if (int e = get_tensor_map_2d(ws + (long long)N * K, K, N, K, BN, &mbl)) return e;
        static const bool ss = getenv("SCP_GEMM_X3_SS") != nullptr;       // A/B switch: the older kernel with A_hi/A_lo in smem
        if (ss) {
            if (BN == 128) return launch<128, 3, true>(ma, mb, mbl, bias, res, ldr, y, ldy, M, N, K, act, st);
            return launch<64, 4, true>(ma, mb, mbl, bias, res, ldr, y, ldy, M, N, K, act, st);
        }
        if (BN == 128) return launch_ts<128, 4, false>(ma, mb, mbl, bias, res, ldr, y, ldy, M, N, K, act, st);
        return launch_ts<64, 4, false>(ma, mb, mbl, bias, res, ldr, y, ldy, M, N, K, act, st);
    }
    const int BN = N > 128 ? 256 : (N > 64 ? 128 : 64);
    if (int e = get_tensor_map_2d(w, K, N, K, BN, &mb)) return e;
    if (BN == 256) return launch<256, 4, false>(ma, mb, mb, bias, res, ldr, y, ldy, M, N, K, act, st);
    if (BN == 128) return launch<128, 6, false>(ma, mb, mb, bias, res, ldr, y, ldy, M, N, K, act, st);
    return launch<64, 8, false>(ma, mb, mb, bias, res, ldr, y, ldy, M, N, K, act, st);
}

}  // namespace scp

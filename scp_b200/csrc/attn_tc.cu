// 1-D shifted-window attention (swin_transformer.py:406-501,603-697) on the 5th-gen tensor cores.
//
// One CTA per (window, head, 128-query block); the 512 keys of the window stream through in 8 chunks of 64:
//     S_c = Q K_c^T          tcgen05.mma M128 x N64 x K64   (3xTF32: Q_hi K_hi + Q_lo K_hi + Q_hi K_lo), fp32 accum in TMEM
//     P_c = exp(S_c + bias + mask - m)      one query row per thread (tcgen05.ld), online softmax in registers
//     O_c = P_c V_c          tcgen05.mma M128 x N64 x K64   (P_hi V_hi + P_lo V_hi + P_hi V_lo), accum in TMEM,
//                            rescaled and added to the running output in registers
// The A operands live in TENSOR MEMORY (the "TS" form of tcgen05.mma): Q_hi/Q_lo are written there once with tcgen05.st,
// and the softmax threads write P_hi over the S columns they just read and P_lo next to them.  An MMA with N = 64 reads
// its 4 KB A slice from shared memory in 32 cycles but computes in 32: with A in TMEM only the 2 KB B slice crosses the
// shared-memory port and the kernel is MMA-bound instead of shared-memory-bound.
//
// TMEM map (512 columns): Q_hi [0,64) | Q_lo [64,128) | buffer b: S / P_hi [128+128b, +64), P_lo [+64, +128) | O_c b [384+64b, +64)
// Warp roles (512 threads, 128 registers): warps 0-7 softmax / output (two threads per query row: TMEM lane quarter =
// warp & 3, column / dim half = warp >> 2), warp 8 MMA issuer + TMEM owner, warps 9-12 K loader, warps 13-15 V loader.
// (A clock64 trace of one CTA showed the single-thread-per-row softmax at ~2600 clk and the 64-thread loaders at
// ~3000 clk per 64-key chunk against 1536 clk of MMAs.)  S/P, O_c, K and V stages are double-buffered and handed over through
// mbarriers, so the MMAs of chunk i+1 overlap the softmax of chunk i and the staging of chunk i+2.
// The loader warps write K_c and V_c^T in the canonical K-major 128B-swizzled layout: the roll by -shift, the zero
// padding of the sequence (padded tokens carry exactly the Linear biases), the hi/lo split and the V transpose happen
// on the way, so HBM/L2 is read once, in fp32, without a TMA descriptor for the gathered rows.
#include <stdlib.h>
#include <stdio.h>
#include "tc.cuh"

struct scp_seqs;

namespace scp {

constexpr int AT_WS = 512, AT_HD = 64, AT_BQ = 128, AT_BK = 64, AT_NC = AT_WS / AT_BK;
constexpr int AT_K_ATOM = AT_BK * 128;           // [64 keys x 128 B] one 32-float K-atom of the K operand      (8 KB)
constexpr int AT_V_ATOM = AT_HD * 128;           // [64 dims x 128 B] one 32-key  K-atom of the V^T operand     (8 KB)
// shared memory map (bytes, from a 1024-aligned base)
constexpr int AT_K_STAGE = 4 * AT_K_ATOM;        // K_hi (2 atoms) | K_lo (2 atoms)
constexpr int AT_V_STAGE = 4 * AT_V_ATOM;        // Vt_hi (2 atoms) | Vt_lo (2 atoms)
constexpr int AT_OFF_K = 0;
constexpr int AT_OFF_V = AT_OFF_K + 2 * AT_K_STAGE;
constexpr int AT_OFF_BIAS = AT_OFF_V + 2 * AT_V_STAGE;   // 1023 floats
constexpr int AT_OFF_XCH = AT_OFF_BIAS + 4096;                 // 3 x [2][128] floats: chunk-max exchange (2 slots) + row sums
constexpr int AT_OFF_BAR = AT_OFF_XCH + 3 * 1024;
constexpr int AT_SMEM = AT_OFF_BAR + 256 + 1024;
constexpr int AT_THREADS = 512;
constexpr uint32_t AT_TMEM_COLS = 512;
constexpr uint32_t AT_T_QH = 0, AT_T_QL = 64, AT_T_SP = 128, AT_T_O = 384;
constexpr float AT_LOG2E = 1.4426950408889634f;

__device__ __forceinline__ void split_tf32(const float4 v, uint4& hi, uint4& lo) {
    hi.x = __float_as_uint(v.x) & 0xffffe000u; hi.y = __float_as_uint(v.y) & 0xffffe000u;
    hi.z = __float_as_uint(v.z) & 0xffffe000u; hi.w = __float_as_uint(v.w) & 0xffffe000u;
    lo.x = __float_as_uint(v.x - __uint_as_float(hi.x)); lo.y = __float_as_uint(v.y - __uint_as_float(hi.y));
    lo.z = __float_as_uint(v.z - __uint_as_float(hi.z)); lo.w = __float_as_uint(v.w - __uint_as_float(hi.w));
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(AT_THREADS, 1) k_swin_attn_tc(const float* __restrict__ Q, long long ldq,
                                                                 const float* __restrict__ K, long long ldk,
                                                                 const float* __restrict__ V, long long ldv,
                                                                 const float* __restrict__ qb, const float* __restrict__ kb,
                                                                 const float* __restrict__ vb, const float* __restrict__ relpos,
                                                                 int heads, const long long* __restrict__ seq_off,
                                                                 const int* __restrict__ win_seq, const int* __restrict__ win_idx,
                                                                 int shift, float* __restrict__ O, long long ldo, long long* __restrict__ trace) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by an OFFSET from the shared-space symbol: a pointer rebuilt from an integer would be generic
    // (LD/ST instead of LDS/STS and no alias information against global memory)
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* s_bias = reinterpret_cast<float*>(sm + AT_OFF_BIAS);
    float* s_xch = reinterpret_cast<float*>(sm + AT_OFF_XCH);      // [2 slots][2 halves][128 rows] pair exchange
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + AT_OFF_BAR);
    uint64_t* k_full = bars;            // [2] K chunk staged                         (4 loader warps)
    uint64_t* v_full = bars + 2;        // [2] V chunk staged                         (3 loader warps)
    uint64_t* s_full = bars + 4;        // [2] S chunk in TMEM, K stage free          (tcgen05.commit)
    uint64_t* p_full = bars + 6;        // [2] P chunk written to TMEM                (8 softmax warps)
    uint64_t* pv_done = bars + 8;       // [2] O_c in TMEM, V stage and P_lo columns free (tcgen05.commit)
    uint64_t* o_empty = bars + 10;      // [2] O_c read back                          (8 softmax warps)
    uint64_t* q_full = bars + 12;       // [1] Q_hi/Q_lo written to TMEM              (8 softmax warps)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    // optional clock64 trace of one CTA (SCP_ATTN_TRACE=1, development aid): slot layout in swin_attn_tc()
    const bool TR = trace && blockIdx.x == 5 && blockIdx.y == 3;
#define AT_STAMP(slot) do { if (TR && lane == 0) trace[slot] = clock64(); } while (0)
    if (TR && t == 0) trace[0] = clock64();
    const int h = blockIdx.x % heads, qblk = blockIdx.x / heads;
    const int gw = blockIdx.y;
    const int s = win_seq[gw], w = win_idx[gw];
    const long long base = seq_off[s];
    const int S = (int)(seq_off[s + 1] - base);
    const int Sp = ((S + AT_WS - 1) / AT_WS) * AT_WS;
    const bool last_win = (w == Sp / AT_WS - 1) && shift > 0;
    // rolled position of this block's first query row; blocks are 128-aligned inside the 512-aligned padded sequence,
    // so a block never wraps and it has real (stored) rows iff its first row is real
    int q_start = w * AT_WS + qblk * AT_BQ + shift;
    if (q_start >= Sp) q_start -= Sp;
    if (q_start >= S) return;                                              // CTA-uniform, before any barrier / TMEM use

    if (warp == 8) {
        if (lane == 0) {
            for (int b = 0; b < 2; ++b) {
                mbar_init(&k_full[b], 4); mbar_init(&v_full[b], 3); mbar_init(&s_full[b], 1);
                mbar_init(&p_full[b], 8); mbar_init(&pv_done[b], 1); mbar_init(&o_empty[b], 8);
            }
            mbar_init(q_full, 8);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(AT_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int e = t; e < 2 * AT_WS - 1; e += AT_THREADS) s_bias[e] = relpos[e * heads + h] * AT_LOG2E;   // scores live in the log2 domain
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (TR && t == 0) trace[1] = clock64();

    if (warp < 8) {
        // ---------------- softmax + output: TWO threads per query row ----------------
        // row = 32 (warp & 3) + lane (the TMEM lane quarter of both warps); half = warp >> 2 owns score columns
        // [32 half, +32) of every chunk and output dims [32 half, +32).  The two threads agree on the chunk maximum through
        // a shared-memory slot + a 64-thread named barrier, everything else is private; the row sum is joined at the end.
        const int quarter = warp & 3, half = warp >> 2;
        const int row = quarter * 32 + lane;
        const uint32_t trow = tmem + ((uint32_t)(quarter * 32) << 16);
        const int bar_id = 1 + quarter;
        {   // half a Q row -> TMEM: 1/sqrt(64) and log2(e) folded in (softmax(x) = 2^(x log2e - max) / sum), hi/lo split
            const int u = q_start + row;
            const float4* src = reinterpret_cast<const float4*>(u < S ? Q + (base + u) * ldq + h * AT_HD : qb + h * AT_HD) + half * 8;
            uint32_t hi[32], lo[32];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float4 v = __ldg(src + c);
                const float qs = 0.125f * AT_LOG2E;
                v.x *= qs; v.y *= qs; v.z *= qs; v.w *= qs;
                uint4 h4, l4;
                split_tf32(v, h4, l4);
                hi[4 * c] = h4.x; hi[4 * c + 1] = h4.y; hi[4 * c + 2] = h4.z; hi[4 * c + 3] = h4.w;
                lo[4 * c] = l4.x; lo[4 * c + 1] = l4.y; lo[4 * c + 2] = l4.z; lo[4 * c + 3] = l4.w;
            }
            tc_st32(trow + AT_T_QH + (uint32_t)(half * 32), hi);
            tc_st32(trow + AT_T_QL + (uint32_t)(half * 32), lo);
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(q_full);
            if (warp == 0) AT_STAMP(2);
        }
        float o_acc[32];
#pragma unroll
        for (int d = 0; d < 32; ++d) o_acc[d] = 0.f;
        float m_run = -INFINITY, l_run = 0.f, m_acc = -INFINITY, m_hist0 = 0.f, m_hist1 = 0.f;
        const int pi = qblk * AT_BQ + row;                                 // window position of this row
#pragma unroll 1
        for (int i = 0; i < AT_NC; ++i) {
            const int b = i & 1, n = i >> 1;
            const uint32_t t_sp = trow + AT_T_SP + (uint32_t)(b * 128);
            mbar_wait(&s_full[b], n & 1);
            if (warp == 0) AT_STAMP(10 + 4 * i);
            tc_fence_after();
            uint32_t r[32];
            tc_ld32(t_sp + (uint32_t)(half * 32), r);
            // swin_transformer.py:620: -100 on the other half of the last (rolled) window.  A whole chunk is on one side, so
            // the offset is folded into the running-max bookkeeping instead of being added to all scores.
            const bool masked = last_win && ((pi < AT_WS / 2) != (i < AT_NC / 2));
            const float moff = masked ? -100.0f * AT_LOG2E : 0.0f;
            const float* bp = s_bias + (pi - i * AT_BK - half * 32 + AT_WS - 1);
            float cmax = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float v = __uint_as_float(r[j]) + bp[-j];
                r[j] = __float_as_uint(v);
                cmax = fmaxf(cmax, v);
            }
            // chunk maximum of the whole row: exchange with the partner thread (slot alternates per chunk, one barrier)
            float* slot = s_xch + (i & 1) * 256;
            slot[half * 128 + row] = cmax;
            asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
            if (warp == 0) AT_STAMP(11 + 4 * i);
            cmax = fmaxf(cmax, slot[(half ^ 1) * 128 + row]);
            const float mx = fmaxf(m_run, cmax + moff);
            const float alpha = ex2_approx(m_run - mx);
            m_run = mx;
            const float m_old = b ? m_hist1 : m_hist0;                     // max the in-flight O_c of this buffer refers to
            if (b) m_hist1 = mx; else m_hist0 = mx;
            const float sub = mx - moff;
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float p = ex2_approx(__uint_as_float(r[j]) - sub);
                sum += p;
                r[j] = __float_as_uint(p);
            }
            l_run = fmaf(l_run, alpha, sum);
            // PV(i-2) retired: its O_c is complete and the P_lo columns of buffer b are free again
            if (warp == 0) AT_STAMP(12 + 4 * i);
            if (i >= 2) { mbar_wait(&pv_done[b], (n - 1) & 1); tc_fence_after(); }
            // P_hi over the S columns, P_lo next to them (first, so that the MMA warp can go on), 16 columns at a time
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                uint32_t* src = r + 16 * q;
                uint32_t lo[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const uint32_t hi = src[j] & 0xffffe000u;
                    lo[j] = __float_as_uint(__uint_as_float(src[j]) - __uint_as_float(hi));
                    src[j] = hi;
                }
                tc_st16(t_sp + (uint32_t)(half * 32 + 16 * q), src);
                tc_st16(t_sp + 64u + (uint32_t)(half * 32 + 16 * q), lo);
            }
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[b]);
            if (warp == 0) AT_STAMP(13 + 4 * i);
            if (i >= 2) {                                                  // fold this thread's 32 dims of O_c(i-2) into the output
                const float sc = ex2_approx(m_acc - m_old);
                m_acc = m_old;
                uint32_t q0[32];
                tc_ld32(trow + AT_T_O + (uint32_t)(b * 64 + half * 32), q0);
#pragma unroll
                for (int d = 0; d < 32; ++d) o_acc[d] = fmaf(o_acc[d], sc, __uint_as_float(q0[d]));
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_relaxed(&o_empty[b]);
            }
        }
#pragma unroll 1
        for (int j = AT_NC - 2; j < AT_NC; ++j) {                          // drain the two in-flight O_c
            const int b = j & 1;
            mbar_wait(&pv_done[b], (j >> 1) & 1);
            tc_fence_after();
            const float m_j = b ? m_hist1 : m_hist0;
            const float sc = ex2_approx(m_acc - m_j);
            m_acc = m_j;
            uint32_t q0[32];
            tc_ld32(trow + AT_T_O + (uint32_t)(b * 64 + half * 32), q0);
#pragma unroll
            for (int d = 0; d < 32; ++d) o_acc[d] = fmaf(o_acc[d], sc, __uint_as_float(q0[d]));
        }
        // row sum = both halves (same running maximum, so the partial sums simply add)
        float* slot = s_xch + 512;
        slot[half * 128 + row] = l_run;
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
        const float l_tot = l_run + slot[(half ^ 1) * 128 + row];
        const int u = q_start + row;
        if (u < S) {
            const float inv = 1.0f / l_tot;                               // m_acc == m_run here
            float* dst = O + (base + u) * ldo + h * AT_HD + half * 32;
#pragma unroll
            for (int d = 0; d < 32; d += 4)
                *reinterpret_cast<float4*>(dst + d) = make_float4(o_acc[d] * inv, o_acc[d + 1] * inv, o_acc[d + 2] * inv, o_acc[d + 3] * inv);
        }
        if (warp == 0) AT_STAMP(3);
    } else if (warp == 8) {
        // ---------------- MMA issuer: all lanes run the loop, the elected lane issues (tc.cuh elect_one) ----------------
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(AT_BQ >> 4) << 24);
        mbar_wait(q_full, 0);
#pragma unroll 1
        for (int i = 0; i <= AT_NC; ++i) {
            if (i < AT_NC) {                                               // S(i) = Q K_i^T  (after PV(i-2) in program order,
                const int b = i & 1, n = i >> 1;                           //  which read P_hi from the same columns)
                mbar_wait(&k_full[b], n & 1);
                AT_STAMP(50 + 2 * i);
                tc_fence_after();
                const uint8_t* ks_ = sm + AT_OFF_K + b * AT_K_STAGE;
                const uint32_t d_tmem = tmem + AT_T_SP + (uint32_t)(b * 128);
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        const uint32_t ko = (ks >> 2) * AT_K_ATOM;
                        const uint64_t adv = (uint64_t)(2 * (ks & 3));
                        const uint64_t kh = make_smem_desc(ks_ + ko) + adv, kl = make_smem_desc(ks_ + 2 * AT_K_ATOM + ko) + adv;
                        tc_mma_tf32_ts(d_tmem, tmem + AT_T_QH + 8u * ks, kh, idesc, ks ? 1u : 0u);
                        tc_mma_tf32_ts(d_tmem, tmem + AT_T_QL + 8u * ks, kh, idesc, 1u);
                        tc_mma_tf32_ts(d_tmem, tmem + AT_T_QH + 8u * ks, kl, idesc, 1u);
                    }
                    tc_commit(&s_full[b]);
                }
                __syncwarp();
            }
            if (i >= 1) {                                                  // O_c(j) = P_j V_j
                const int j = i - 1, b = j & 1, n = j >> 1;
                mbar_wait(&v_full[b], n & 1);
                mbar_wait(&p_full[b], n & 1);
                if (n > 0) mbar_wait(&o_empty[b], (n - 1) & 1);
                AT_STAMP(51 + 2 * j);
                tc_fence_after();
                const uint8_t* vs_ = sm + AT_OFF_V + b * AT_V_STAGE;
                const uint32_t p_tmem = tmem + AT_T_SP + (uint32_t)(b * 128);
                const uint32_t d_tmem = tmem + AT_T_O + (uint32_t)(b * 64);
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        const uint32_t vo = (ks >> 2) * AT_V_ATOM;
                        const uint64_t adv = (uint64_t)(2 * (ks & 3));
                        const uint64_t vh = make_smem_desc(vs_ + vo) + adv, vl = make_smem_desc(vs_ + 2 * AT_V_ATOM + vo) + adv;
                        tc_mma_tf32_ts(d_tmem, p_tmem + 8u * ks, vh, idesc, ks ? 1u : 0u);
                        tc_mma_tf32_ts(d_tmem, p_tmem + 64u + 8u * ks, vh, idesc, 1u);
                        tc_mma_tf32_ts(d_tmem, p_tmem + 8u * ks, vl, idesc, 1u);
                    }
                    tc_commit(&pv_done[b]);
                }
                __syncwarp();
            }
        }
    } else if (warp < 13) {
        // ---------------- K loader (warps 9-12, 128 threads): chunk rows = keys, 64 dims = 2 atoms ----------------
        const int L = t - 288;
        const int c4 = (L & 15) << 2, r0 = L >> 4;
        const uint32_t col_off = (uint32_t)((c4 >> 5) * AT_K_ATOM);
        const int chunk = (c4 & 31) >> 2;
#pragma unroll 1
        for (int i = 0; i < AT_NC; ++i) {
            const int b = i & 1, n = i >> 1;
            float4 kv[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int r = r0 + 8 * e;
                int u = w * AT_WS + i * AT_BK + r + shift;
                if (u >= Sp) u -= Sp;
                kv[e] = u < S ? __ldg(reinterpret_cast<const float4*>(K + (base + u) * ldk + h * AT_HD + c4))
                              : __ldg(reinterpret_cast<const float4*>(kb + h * AT_HD + c4));
            }
            if (warp == 9) AT_STAMP(70 + 3 * i);
            if (n > 0) mbar_wait(&s_full[b], (n - 1) & 1);               // S(i-2) retired: K stage b is free
            if (warp == 9) AT_STAMP(71 + 3 * i);
            uint8_t* dst = sm + AT_OFF_K + b * AT_K_STAGE + col_off;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int r = r0 + 8 * e;
                uint4 hi, lo;
                split_tf32(kv[e], hi, lo);
                const uint32_t o = (uint32_t)(r * 128 + ((chunk ^ (r & 7)) << 4));
                *reinterpret_cast<uint4*>(dst + o) = hi;
                *reinterpret_cast<uint4*>(dst + 2 * AT_K_ATOM + o) = lo;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&k_full[b]);
            if (warp == 9) AT_STAMP(72 + 3 * i);
        }
    } else {
        // ---------------- V loader (warps 13-15, 96 threads): transposed tile, rows = dims, cols = keys ----------------
        // work unit = 4 keys x 4 dims: blk = atom(2) x dim group(16) x key group(8); consecutive lanes take consecutive key
        // groups of one dim group, which makes the transposed 16-byte stores conflict-free
        const int L = t - 416;
#pragma unroll 1
        for (int i = 0; i < AT_NC; ++i) {
            const int b = i & 1, n = i >> 1;
            float4 vv[3][4];
#pragma unroll
            for (int it = 0; it < 3; ++it) {
                const int blk = L + 96 * it;
                if (blk < 256) {
                    const int kg = blk & 7, d4 = (blk >> 3) & 15, at = blk >> 7;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        int u = w * AT_WS + i * AT_BK + at * 32 + 4 * kg + q + shift;
                        if (u >= Sp) u -= Sp;
                        vv[it][q] = u < S ? __ldg(reinterpret_cast<const float4*>(V + (base + u) * ldv + h * AT_HD + 4 * d4))
                                          : __ldg(reinterpret_cast<const float4*>(vb + h * AT_HD + 4 * d4));
                    }
                }
            }
            if (n > 0) mbar_wait(&pv_done[b], (n - 1) & 1);              // PV(i-2) retired: V stage b is free
#pragma unroll
            for (int it = 0; it < 3; ++it) {
                const int blk = L + 96 * it;
                if (blk < 256) {
                    const int kg = blk & 7, d4 = (blk >> 3) & 15, at = blk >> 7;
                    uint8_t* dst = sm + AT_OFF_V + b * AT_V_STAGE + at * AT_V_ATOM;
                    const float4 a0 = vv[it][0], a1 = vv[it][1], a2 = vv[it][2], a3 = vv[it][3];
                    const float4 rows[4] = {make_float4(a0.x, a1.x, a2.x, a3.x), make_float4(a0.y, a1.y, a2.y, a3.y),
                                            make_float4(a0.z, a1.z, a2.z, a3.z), make_float4(a0.w, a1.w, a2.w, a3.w)};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int rw = 4 * d4 + c;
                        uint4 hi, lo;
                        split_tf32(rows[c], hi, lo);
                        const uint32_t o = (uint32_t)(rw * 128 + ((kg ^ (rw & 7)) << 4));
                        *reinterpret_cast<uint4*>(dst + o) = hi;
                        *reinterpret_cast<uint4*>(dst + 2 * AT_V_ATOM + o) = lo;
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&v_full[b]);
            if (warp == 13) AT_STAMP(100 + i);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (TR && t == 0) trace[4] = clock64();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(AT_TMEM_COLS) : "memory");
    }
}

bool swin_attn_tc_ok(long long ldq, long long ldk, long long ldv, long long ldo, const void* q, const void* k, const void* v,
                     const void* o, const void* b0, const void* b1, const void* b2) {
    auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    return ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ldo % 4 == 0 && al(q) && al(k) && al(v) && al(o) && al(b0) && al(b1) && al(b2);
}

int swin_attn_tc(const float* q, long long ldq, const float* k, long long ldk, const float* v, long long ldv, const float* qb,
                 const float* kb, const float* vb, const float* relpos, int heads, const long long* d_off, const int* d_win_seq,
                 const int* d_win_idx, int n_win, int shift, float* out, long long ldo, cudaStream_t st) {
    static bool attr = false;
    if (!attr) { SCP_CUDA(cudaFuncSetAttribute(k_swin_attn_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM)); attr = true; }
    dim3 grid((AT_WS / AT_BQ) * heads, n_win);
    static long long* d_trace = nullptr;
    static const bool want_trace = getenv("SCP_ATTN_TRACE") != nullptr;
    if (want_trace && !d_trace) { cudaMalloc(&d_trace, 256 * 8); cudaMemset(d_trace, 0, 256 * 8); }
    k_swin_attn_tc<<<grid, AT_THREADS, AT_SMEM, st>>>(q, ldq, k, ldk, v, ldv, qb, kb, vb, relpos, heads, d_off, d_win_seq,
                                                      d_win_idx, shift, out, ldo, d_trace);
    if (want_trace && n_win > 8) {                       // development aid: cycle stamps of CTA (5, 3), relative to its start
        long long hh[256];
        cudaStreamSynchronize(st);
        cudaMemcpy(hh, d_trace, sizeof(hh), cudaMemcpyDeviceToHost);
        static int printed = 0;
        if (printed++ < 2) {
            const long long t0 = hh[0];
            printf("ATTN TRACE: setup %lld  q_in_tmem %lld  softmax_end %lld  cta_end %lld\n", hh[1] - t0, hh[2] - t0, hh[3] - t0, hh[4] - t0);
            for (int i = 0; i < AT_NC; ++i)
                printf("  chunk %d: softmax s_ready %lld  max_exchanged %lld  p_computed %lld  p_arrived %lld | mma S_issue %lld PV_issue %lld | "
                       "K loads_issued %lld stage_free %lld arrive %lld | V arrive %lld\n", i, hh[10 + 4 * i] - t0, hh[11 + 4 * i] - t0,
                       hh[12 + 4 * i] - t0, hh[13 + 4 * i] - t0, hh[50 + 2 * i] - t0, hh[51 + 2 * i] - t0, hh[70 + 3 * i] - t0,
                       hh[71 + 3 * i] - t0, hh[72 + 3 * i] - t0, hh[100 + i] - t0);
        }
    }
    SCP_LAUNCHED();
    return SCP_OK;
}

}  // namespace scp

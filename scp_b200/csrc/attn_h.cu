// Swin window attention (swin_transformer.py:406-501 + the shifted-window roll / mask of :583-706) on the FP16 tensor pipe,
// TWO CTAs per SM.
//
// Same math as k_swin_attn_tc (attn_tc.cu): S = Q K^T and O = P V as error-compensated products (x = x_hi + x_lo, three
// tcgen05.mma per product: hi.hi + lo.hi + hi.lo, fp32 accumulation), online softmax in the log2 domain.  Differences:
//  * hi / lo are FP16 (11 + 11 mantissa bits like the tf32 split; Q, K, V are O(1) activations, P is in [0,1]):
//    kind::f16 runs at twice the tf32 rate, and the operands take half the space -- Q_hi/Q_lo 64 TMEM columns instead of
//    128, P_hi/P_lo fit INTO the 64 columns of the S chunk they come from, K / V^T stages are 16 KB instead of 32 KB;
//  * the output accumulates in TMEM across the eight key chunks (PV with accumulate) and is rescaled in place when a row's
//    running maximum moves, instead of being folded into 32 registers per thread: the softmax threads fit in 80 registers;
//  * so a CTA needs 256 TMEM columns, 72 KB of shared memory and 384 threads, and TWO CTAs share an SM.  The tf32 kernel
//    was bound by its softmax warps (two per scheduler, ~2000 cycles of dependent issue per 64-key chunk against 1536
//    cycles of MMAs) and left the SM idle during its ~9000-cycle prologue (TMEM alloc, Q load, first K chunk); the second
//    CTA fills both.
//
// TMEM map (256 columns): Q_hi [0,32) | Q_lo [32,64) | S/P buffer b at [64 + 64 b, +64): S fp32, then P_hi [+0,+32) P_lo [+32,+64)
//                         | O [192,256)
// Warp roles (320 threads): warps 0-7 softmax / output (two threads per query row: TMEM lane quarter = warp & 3, column /
// dim half = warp >> 2), warp 8 MMA issuer + TMEM owner, warp 9 TMA producer.
// K and V are prepared ONCE per launch by k_attn_prep (roll by -shift, zero padding of the sequence = rows that carry
// exactly the Linear biases, fp16 hi/lo split, V transposed) into [padded token][head dim] / [head dim][padded token] fp16
// matrices, from which the chunks arrive by TMA in the K-major 128B-swizzled layout.  Staging the chunks with loader warps
// inside the kernel (as attn_tc.cu does) re-converts every chunk for each of the four query blocks of a window and, with
// the three loader warps that fit next to a second CTA, set the pace of the whole kernel (measured: 8.7 k cycles per chunk).
#include <algorithm>
#include <stdlib.h>
#include <cuda_fp16.h>
#include "tc.cuh"

struct scp_seqs;

namespace scp {

constexpr int AH_WS = 512, AH_HD = 64, AH_BQ = 128, AH_BK = 64, AH_NC = AH_WS / AH_BK;
constexpr int AH_TILE = 64 * 128;                // [64 rows x 128 B]: 64 keys x 64 dims (K) or 64 dims x 64 keys (V^T), fp16  (8 KB)
constexpr int AH_STAGE = 2 * AH_TILE;            // hi | lo
constexpr int AH_OFF_K = 0;
constexpr int AH_OFF_V = AH_OFF_K + 2 * AH_STAGE;
constexpr int AH_OFF_BIAS = AH_OFF_V + 2 * AH_STAGE;      // 1023 floats
constexpr int AH_OFF_XCH = AH_OFF_BIAS + 4096;            // 3 x [2][128] floats: chunk-max exchange (2 slots) + row sums
constexpr int AH_OFF_BAR = AH_OFF_XCH + 3 * 1024;
constexpr int AH_SMEM = AH_OFF_BAR + 256 + 1024;
constexpr int AH_THREADS = 320;
constexpr uint32_t AH_TMEM_COLS = 256;
constexpr uint32_t AH_T_QH = 0, AH_T_QL = 32, AH_T_SP = 64, AH_T_O = 192;
constexpr float AH_LOG2E = 1.4426950408889634f;

// x0, x1 -> packed fp16 pairs (x0 in the low half): hi = fp16(x) with saturation, lo = fp16(x - hi)
__device__ __forceinline__ void ah_split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
    float h0, h1;
    asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(h0), "=f"(h1) : "r"(hi));
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(x1 - h1), "f"(x0 - h0));
}
__device__ __forceinline__ float ah_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// 16 lanes x 64 columns: register 4 n + {0,1} = (lane l / 4, column 8 n + 2 (l % 4) + {0,1}), 4 n + {2,3} = the same of lane l / 4 + 8
__device__ __forceinline__ void ah_ld_16x256b_x8(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ah_st_16x256b_x8(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.16x256b.x8.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),
          "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]),
          "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
// 16 lanes x 32 columns: register 2 n + k = (lane l / 4 + 8 k, column 4 n + l % 4)
__device__ __forceinline__ void ah_st_16x128b_x8(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.16x128b.x8.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}

// K / V preparation: block = (64-key chunk of a window, head).  K16hi/lo [n_win*512][heads*64], V16Thi/lo [heads*64][n_win*512].
__global__ void __launch_bounds__(256) k_attn_prep(const float* __restrict__ K, long long ldk, const float* __restrict__ V,
                                                    long long ldv, const float* __restrict__ kb, const float* __restrict__ vb,
                                                    int heads, const long long* __restrict__ seq_off,
                                                    const int* __restrict__ win_seq, const int* __restrict__ win_idx, int shift,
                                                    int n_win, __half* __restrict__ k_hi, __half* __restrict__ k_lo,
                                                    __half* __restrict__ vt_hi, __half* __restrict__ vt_lo) {
    __shared__ float sv[64][65];                                           // V chunk [key][dim] for the transpose
    const int gw = blockIdx.x >> 3, i = blockIdx.x & 7, h = blockIdx.y;
    const int s = win_seq[gw], w = win_idx[gw];
    const long long base = seq_off[s];
    const int S = (int)(seq_off[s + 1] - base);
    const int Sp = ((S + AH_WS - 1) / AH_WS) * AH_WS;
    const long long prow0 = (long long)gw * AH_WS + i * AH_BK;              // first padded row of the chunk
    const long long ldp = (long long)heads * AH_HD, ldt = (long long)n_win * AH_WS;
#pragma unroll
    for (int e = 0; e < 4; ++e) {                                          // 1024 float4 units: key r, dims c4..c4+3
        const int unit = e * 256 + threadIdx.x;
        const int r = unit >> 4, c4 = (unit & 15) << 2;
        int u = w * AH_WS + i * AH_BK + r + shift;
        if (u >= Sp) u -= Sp;
        const float4 kv = u < S ? __ldg(reinterpret_cast<const float4*>(K + (base + u) * ldk + h * AH_HD + c4))
                                : __ldg(reinterpret_cast<const float4*>(kb + h * AH_HD + c4));
        const float4 vv = u < S ? __ldg(reinterpret_cast<const float4*>(V + (base + u) * ldv + h * AH_HD + c4))
                                : __ldg(reinterpret_cast<const float4*>(vb + h * AH_HD + c4));
        uint2 hi, lo;
        ah_split2(kv.x, kv.y, hi.x, lo.x);
        ah_split2(kv.z, kv.w, hi.y, lo.y);
        const long long o = (prow0 + r) * ldp + h * AH_HD + c4;
        *reinterpret_cast<uint2*>(k_hi + o) = hi;
        *reinterpret_cast<uint2*>(k_lo + o) = lo;
        sv[r][c4] = vv.x; sv[r][c4 + 1] = vv.y; sv[r][c4 + 2] = vv.z; sv[r][c4 + 3] = vv.w;
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 4; ++e) {                                          // 1024 units: dim d, keys k4..k4+3
        const int unit = e * 256 + threadIdx.x;
        const int d = unit >> 4, k4 = (unit & 15) << 2;
        uint2 hi, lo;
        ah_split2(sv[k4][d], sv[k4 + 1][d], hi.x, lo.x);
        ah_split2(sv[k4 + 2][d], sv[k4 + 3][d], hi.y, lo.y);
        const long long o = ((long long)h * AH_HD + d) * ldt + prow0 + k4;
        *reinterpret_cast<uint2*>(vt_hi + o) = hi;
        *reinterpret_cast<uint2*>(vt_lo + o) = lo;
    }
}

__global__ void __launch_bounds__(AH_THREADS, 2) k_swin_attn_h(const float* __restrict__ Q, long long ldq,
                                                                const __grid_constant__ CUtensorMap tmKh,
                                                                const __grid_constant__ CUtensorMap tmKl,
                                                                const __grid_constant__ CUtensorMap tmVh,
                                                                const __grid_constant__ CUtensorMap tmVl,
                                                                const float* __restrict__ qb, const float* __restrict__ relpos,
                                                                int heads, const long long* __restrict__ seq_off,
                                                                const int* __restrict__ win_seq, const int* __restrict__ win_idx,
                                                                int shift, float* __restrict__ O, long long ldo, int n_win) {
    // PERSISTENT: the grid is a multiple of `heads` CTAs (two per SM); CTA c keeps head c % heads (its relative-position table is
    // loaded once) and walks the (window, query block) pairs c / heads, + gridDim.x / heads, ...  TMEM, barriers and the table
    // are set up once per CTA instead of once per 128 queries (the set-up was ~9000 of the ~30 000 cycles of a work item).  Every
    // chunk barrier completes an even number of phases per work item (4), so the parities of the chunk loop repeat unchanged;
    // q_full (1 phase per item) and o_ready (7) carry the item count.
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* s_bias = reinterpret_cast<float*>(sm + AH_OFF_BIAS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + AH_OFF_BAR);
    uint64_t* k_full = bars;            // [2] K chunk landed                                   (TMA, expect_tx)
    uint64_t* v_full = bars + 2;        // [2] V chunk landed                                   (TMA, expect_tx)
    uint64_t* s_full = bars + 4;        // [2] S chunk in TMEM, K stage free                    (tcgen05.commit)
    uint64_t* p_full = bars + 6;        // [2] P chunk written to TMEM                          (8 softmax warps)
    uint64_t* pv_done = bars + 8;       // [2] PV of the chunk retired: O updated, V stage free (tcgen05.commit)
    uint64_t* o_ready = bars + 10;      // [1] O rescaled for the next chunk's maximum          (8 softmax warps)
    uint64_t* q_full = bars + 11;       // [1] Q_hi/Q_lo written to TMEM                        (8 softmax warps)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int h = blockIdx.x % heads;
    const int pair0 = blockIdx.x / heads, pair_step = gridDim.x / heads, n_pairs = n_win * (AH_WS / AH_BQ);
    // work item `pair`: window gw, query block qblk.  Returns false for a block without real (stored) rows: blocks are
    // 128-aligned inside the 512-aligned padded sequence, so a block never wraps and has real rows iff its first row is real
    struct Item { int gw, qblk, S, q_start; long long base; bool last_win; };
    auto item_of = [&](int pair, Item& I) -> bool {
        I.gw = pair / (AH_WS / AH_BQ); I.qblk = pair % (AH_WS / AH_BQ);
        const int s = win_seq[I.gw], w = win_idx[I.gw];
        I.base = seq_off[s];
        I.S = (int)(seq_off[s + 1] - I.base);
        const int Sp = ((I.S + AH_WS - 1) / AH_WS) * AH_WS;
        I.last_win = (w == Sp / AH_WS - 1) && shift > 0;
        int q_start = w * AH_WS + I.qblk * AH_BQ + shift;                  // rolled position of the block's first query row
        if (q_start >= Sp) q_start -= Sp;
        I.q_start = q_start;
        return q_start < I.S;
    };

    if (warp == 8) {
        if (lane == 0) {
            for (int b = 0; b < 2; ++b) {
                mbar_init(&k_full[b], 1); mbar_init(&v_full[b], 1); mbar_init(&s_full[b], 1);
                mbar_init(&p_full[b], 8); mbar_init(&pv_done[b], 1);
            }
            mbar_init(o_ready, 8);
            mbar_init(q_full, 8);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(AH_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int e = t; e < 2 * AH_WS - 1; e += AH_THREADS) s_bias[e] = relpos[e * heads + h] * AH_LOG2E;   // scores live in the log2 domain
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < 8) {
        // ---------------- softmax + output: a warp owns 16 query rows, FOUR threads per row ----------------
        // rows = TMEM lanes 32 (warp & 3) + 16 (warp >> 2) + [0, 16).  Scores are read with tcgen05.ld.16x256b (layout verified with
        // tools/exp/tmem_layout.cu): thread = (g = lane / 4, q = lane % 4) holds, for n = 0..7, columns 8 n + 2 q + {0, 1} of row g
        // (registers 4 n + 0, 1) and of row g + 8 (registers 4 n + 2, 3).  A row's maximum and sum are quad reductions (two
        // shuffles) -- the earlier two-threads-per-row form exchanged them through shared memory and a named barrier between two
        // warps of different schedulers, and that wait was the largest item of the ~2500-cycle softmax chain.  The pair
        // (column 2 w, 2 w + 1) is exactly one packed fp16 word of P, and tcgen05.st.16x128b puts register 2 n + k at
        // (row g + 8 k, word 4 n + q): P_hi / P_lo go back without any exchange, and no other warp touches these 16 lanes.
        const int quarter = warp & 3, rh = warp >> 2;
        const int g = lane >> 2, q4 = lane & 3;
        const int rowA = quarter * 32 + rh * 16 + g;                       // rowB = rowA + 8
        const uint32_t tbase = tmem + ((uint32_t)(quarter * 32 + rh * 16) << 16);
        Item I;
        uint32_t it = 0;                                                   // work items of this CTA so far
        for (int pair = pair0; pair < n_pairs; pair += pair_step) {
        if (!item_of(pair, I)) continue;
        const long long base = I.base;
        const int S = I.S, q_start = I.q_start, qblk = I.qblk;
        const bool last_win = I.last_win;
        const int uA = q_start + rowA, uB = uA + 8;
        {   // two Q rows -> TMEM: 1/sqrt(64) and log2(e) folded in (softmax(x) = 2^(x log2e - max) / sum), fp16 hi/lo pairs
            // (the S MMAs of the previous item, the last readers of Q, retired before its last softmax chunk started)
            const float* srcA = (uA < S ? Q + (base + uA) * ldq + h * AH_HD : qb + h * AH_HD) + 2 * q4;
            const float* srcB = (uB < S ? Q + (base + uB) * ldq + h * AH_HD : qb + h * AH_HD) + 2 * q4;
            uint32_t hi[16], lo[16];
            const float qs = 0.125f * AH_LOG2E;
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const float2 a = __ldg(reinterpret_cast<const float2*>(srcA + 8 * n));
                const float2 c = __ldg(reinterpret_cast<const float2*>(srcB + 8 * n));
                ah_split2(a.x * qs, a.y * qs, hi[2 * n], lo[2 * n]);
                ah_split2(c.x * qs, c.y * qs, hi[2 * n + 1], lo[2 * n + 1]);
            }
            ah_st_16x128b_x8(tbase + AH_T_QH, hi);
            ah_st_16x128b_x8(tbase + AH_T_QL, lo);
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(q_full);
        }
        float mA = -INFINITY, mB = -INFINITY, lA = 0.f, lB = 0.f;         // running maxima; PARTIAL row sums of this thread's columns
        const int piA = qblk * AH_BQ + rowA;                               // window position of row A (row B: + 8)
#pragma unroll 1
        for (int i = 0; i < AH_NC; ++i) {
            const int b = i & 1, n_ = i >> 1;
            const uint32_t t_sp = tbase + AH_T_SP + (uint32_t)(b * 64);
            mbar_wait(&s_full[b], n_ & 1);
            tc_fence_after();
            uint32_t r[32];
            ah_ld_16x256b_x8(t_sp, r);
            tc_wait_ld();
            // swin_transformer.py:620: -100 on the other half of the last (rolled) window.  A whole chunk is on one side (and so is
            // a 128-row query block), so the offset is folded into the running-max bookkeeping instead of being added to all scores.
            const bool masked = last_win && ((piA < AH_WS / 2) != (i < AH_NC / 2));
            const float moff = masked ? -100.0f * AH_LOG2E : 0.0f;
            const float* bpA = s_bias + (piA - i * AH_BK + AH_WS - 1 - 2 * q4);    // bias of (row, key column c) = bp[-c]
            float cA = -INFINITY, cB = -INFINITY;
            float e0 = bpA[8], e1 = bpA[7];                                // row B = row A + 8 sees row A's bias of the previous n
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const float d0 = bpA[-8 * n], d1 = bpA[-8 * n - 1];
                const float a0 = __uint_as_float(r[4 * n]) + d0, a1 = __uint_as_float(r[4 * n + 1]) + d1;
                const float b0 = __uint_as_float(r[4 * n + 2]) + e0, b1 = __uint_as_float(r[4 * n + 3]) + e1;
                e0 = d0; e1 = d1;
                r[4 * n] = __float_as_uint(a0); r[4 * n + 1] = __float_as_uint(a1);
                r[4 * n + 2] = __float_as_uint(b0); r[4 * n + 3] = __float_as_uint(b1);
                cA = fmaxf(cA, fmaxf(a0, a1)); cB = fmaxf(cB, fmaxf(b0, b1));
            }
            cA = fmaxf(cA, __shfl_xor_sync(0xffffffffu, cA, 1)); cB = fmaxf(cB, __shfl_xor_sync(0xffffffffu, cB, 1));
            cA = fmaxf(cA, __shfl_xor_sync(0xffffffffu, cA, 2)); cB = fmaxf(cB, __shfl_xor_sync(0xffffffffu, cB, 2));
            const float mxA = fmaxf(mA, cA + moff), mxB = fmaxf(mB, cB + moff);
            const float alA = ah_ex2(mA - mxA), alB = ah_ex2(mB - mxB);    // 0 on the first chunk (m = -inf)
            mA = mxA; mB = mxB;
            const float subA = mxA - moff, subB = mxB - moff;
            float sA = 0.f, sB = 0.f;
            uint32_t ph[16], pl[16];
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const float a0 = ah_ex2(__uint_as_float(r[4 * n]) - subA), a1 = ah_ex2(__uint_as_float(r[4 * n + 1]) - subA);
                const float b0 = ah_ex2(__uint_as_float(r[4 * n + 2]) - subB), b1 = ah_ex2(__uint_as_float(r[4 * n + 3]) - subB);
                sA += a0 + a1; sB += b0 + b1;
                ah_split2(a0, a1, ph[2 * n], pl[2 * n]);                   // word 4 n + q of row A: keys 8 n + 2 q, + 1 (even key low)
                ah_split2(b0, b1, ph[2 * n + 1], pl[2 * n + 1]);
            }
            lA = fmaf(lA, alA, sA); lB = fmaf(lB, alB, sB);
            ah_st_16x128b_x8(t_sp, ph);                                    // P_hi over columns [0, 32) of the chunk it came from,
            ah_st_16x128b_x8(t_sp + 32u, pl);                              // P_lo over [32, 64): only this warp reads these lanes
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[b]);
            if (i > 0) {
                // PV(i) accumulates into O: bring O to the new maximum first (PV(i-1) must have retired).  Skipped when no
                // row of the warp moved its maximum (alpha == 1 exactly), the common case after the first chunks.
                mbar_wait(&pv_done[b ^ 1], ((i - 1) >> 1) & 1);
                tc_fence_after();
                if (__any_sync(0xffffffffu, alA != 1.0f || alB != 1.0f)) {
                    uint32_t o[32];
                    ah_ld_16x256b_x8(tbase + AH_T_O, o);
                    tc_wait_ld();
#pragma unroll
                    for (int n = 0; n < 8; ++n) {
                        o[4 * n] = __float_as_uint(__uint_as_float(o[4 * n]) * alA); o[4 * n + 1] = __float_as_uint(__uint_as_float(o[4 * n + 1]) * alA);
                        o[4 * n + 2] = __float_as_uint(__uint_as_float(o[4 * n + 2]) * alB); o[4 * n + 3] = __float_as_uint(__uint_as_float(o[4 * n + 3]) * alB);
                    }
                    ah_st_16x256b_x8(tbase + AH_T_O, o);
                    tc_wait_st();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(o_ready);
            }
        }
        mbar_wait(&pv_done[(AH_NC - 1) & 1], ((AH_NC - 1) >> 1) & 1);      // PV of the last chunk
        tc_fence_after();
        // row sums: the four threads of a row hold partial sums under the same running maximum
        lA += __shfl_xor_sync(0xffffffffu, lA, 1); lB += __shfl_xor_sync(0xffffffffu, lB, 1);
        lA += __shfl_xor_sync(0xffffffffu, lA, 2); lB += __shfl_xor_sync(0xffffffffu, lB, 2);
        uint32_t o[32];
        ah_ld_16x256b_x8(tbase + AH_T_O, o);
        tc_wait_ld();
        if (uA < S) {
            const float inv = 1.0f / lA;
            float* dst = O + (base + uA) * ldo + h * AH_HD + 2 * q4;
#pragma unroll
            for (int n = 0; n < 8; ++n)
                *reinterpret_cast<float2*>(dst + 8 * n) = make_float2(__uint_as_float(o[4 * n]) * inv, __uint_as_float(o[4 * n + 1]) * inv);
        }
        if (uB < S) {
            const float inv = 1.0f / lB;
            float* dst = O + (base + uB) * ldo + h * AH_HD + 2 * q4;
#pragma unroll
            for (int n = 0; n < 8; ++n)
                *reinterpret_cast<float2*>(dst + 8 * n) = make_float2(__uint_as_float(o[4 * n + 2]) * inv, __uint_as_float(o[4 * n + 3]) * inv);
        }
        tc_fence_before();                                                 // O has been read: the next item's PV(0) may overwrite it
        ++it;                                                              // (ordered by that item's q_full arrival)
        }
    } else if (warp == 8) {
        // ---------------- MMA issuer: all lanes run the loop, the elected lane issues (tc.cuh elect_one) ----------------
        const uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(AH_BQ >> 4) << 24);     // f16 x f16 -> f32, N = 64
        Item I;
        uint32_t it = 0;
        for (int pair = pair0; pair < n_pairs; pair += pair_step) {
        if (!item_of(pair, I)) continue;
        mbar_wait(q_full, it & 1);                                         // Q of this item in TMEM, O of the previous one read
        tc_fence_after();
#pragma unroll 1
        for (int i = 0; i <= AH_NC; ++i) {
            if (i < AH_NC) {                                               // S(i) = Q K_i^T  (after PV(i-2) in program order,
                const int b = i & 1, n = i >> 1;                           //  which read P from the same columns)
                mbar_wait(&k_full[b], n & 1);
                tc_fence_after();
                const uint8_t* ks_ = sm + AH_OFF_K + b * AH_STAGE;
                const uint32_t d_tmem = tmem + AH_T_SP + (uint32_t)(b * 64);
                const uint64_t kh = make_smem_desc(ks_), kl = make_smem_desc(ks_ + AH_TILE);
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {                       // 16 dims = 8 TMEM columns of Q = 2 descriptor units of K
                        const uint64_t adv = (uint64_t)(2 * ks);
                        tc_mma_f16_ts(d_tmem, tmem + AH_T_QH + 8u * ks, kh + adv, idesc, ks ? 1u : 0u);
                        tc_mma_f16_ts(d_tmem, tmem + AH_T_QL + 8u * ks, kh + adv, idesc, 1u);
                        tc_mma_f16_ts(d_tmem, tmem + AH_T_QH + 8u * ks, kl + adv, idesc, 1u);
                    }
                    tc_commit(&s_full[b]);
                }
                __syncwarp();
            }
            if (i >= 1) {                                                  // O (+)= P_j V_j
                const int j = i - 1, b = j & 1, n = j >> 1;
                mbar_wait(&v_full[b], n & 1);
                mbar_wait(&p_full[b], n & 1);
                if (j > 0) mbar_wait(o_ready, (it + (uint32_t)(j - 1)) & 1);      // 7 phases per item
                tc_fence_after();
                const uint8_t* vs_ = sm + AH_OFF_V + b * AH_STAGE;
                const uint32_t p_tmem = tmem + AH_T_SP + (uint32_t)(b * 64);
                const uint32_t d_tmem = tmem + AH_T_O;
                const uint64_t vh = make_smem_desc(vs_), vl = make_smem_desc(vs_ + AH_TILE);
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {                       // 16 keys = 8 TMEM columns of P = 2 descriptor units of V^T
                        const uint64_t adv = (uint64_t)(2 * ks);
                        tc_mma_f16_ts(d_tmem, p_tmem + 8u * ks, vh + adv, idesc, (j | ks) ? 1u : 0u);
                        tc_mma_f16_ts(d_tmem, p_tmem + 32u + 8u * ks, vh + adv, idesc, 1u);
                        tc_mma_f16_ts(d_tmem, p_tmem + 8u * ks, vl + adv, idesc, 1u);
                    }
                    tc_commit(&pv_done[b]);
                }
                __syncwarp();
            }
        }
        ++it;
        }
    } else {
        // ---------------- TMA producer (warp 9): chunk i of the window = padded rows [gw*512 + 64 i, +64) ----------------
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmKh)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmKl)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmVh)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmVl)) : "memory");
        }
        Item I;
        uint32_t it = 0;
        for (int pair = pair0; pair < n_pairs; pair += pair_step) {
        if (!item_of(pair, I)) continue;
        const int gw = I.gw;
#pragma unroll 1
        for (int i = 0; i < AH_NC; ++i) {
            const int b = i & 1, n = i >> 1;
            const int prow = gw * AH_WS + i * AH_BK;
            uint8_t* kdst = sm + AH_OFF_K + b * AH_STAGE;
            uint8_t* vdst = sm + AH_OFF_V + b * AH_STAGE;
            // the stage's previous user: chunk i-2 of this item, or chunk 6 + b of the previous item (its 4th phase: parity 1)
            if (n > 0 || it > 0) mbar_wait(&s_full[b], (n - 1) & 1);     // S(i-2) retired: K stage b is free
            if (elect_one()) {
                mbar_expect_tx(&k_full[b], AH_STAGE);
                tma_load_2d(kdst, &tmKh, &k_full[b], h * AH_HD, prow);
                tma_load_2d(kdst + AH_TILE, &tmKl, &k_full[b], h * AH_HD, prow);
            }
            __syncwarp();
            if (n > 0 || it > 0) mbar_wait(&pv_done[b], (n - 1) & 1);    // PV(i-2) retired: V stage b is free
            if (elect_one()) {
                mbar_expect_tx(&v_full[b], AH_STAGE);
                tma_load_2d(vdst, &tmVh, &v_full[b], prow, h * AH_HD);
                tma_load_2d(vdst + AH_TILE, &tmVl, &v_full[b], prow, h * AH_HD);
            }
            __syncwarp();
        }
        ++it;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(AH_TMEM_COLS) : "memory");
    }
}

int swin_attn_h(const float* q, long long ldq, const float* k, long long ldk, const float* v, long long ldv, const float* qb,
                const float* kb, const float* vb, const float* relpos, int heads, const long long* d_off, const int* d_win_seq,
                const int* d_win_idx, int n_win, int shift, float* out, long long ldo, cudaStream_t st) {
    static bool attr = false;
    if (!attr) { SCP_CUDA(cudaFuncSetAttribute(k_swin_attn_h, cudaFuncAttributeMaxDynamicSharedMemorySize, AH_SMEM)); attr = true; }
    const long long rows = (long long)n_win * AH_WS, cols = (long long)heads * AH_HD;
    __half* buf = nullptr;
    SCP_CUDA(malloc_async((void**)&buf, (size_t)(4 * rows * cols) * sizeof(__half) + 1024, st));
    __half *k_hi = buf, *k_lo = buf + rows * cols, *vt_hi = buf + 2 * rows * cols, *vt_lo = buf + 3 * rows * cols;
    k_attn_prep<<<dim3((unsigned)(n_win * AH_NC), (unsigned)heads), 256, 0, st>>>(k, ldk, v, ldv, kb, vb, heads, d_off, d_win_seq,
                                                                                 d_win_idx, shift, n_win, k_hi, k_lo, vt_hi, vt_lo);
    SCP_LAUNCHED();
    CUtensorMap mkh, mkl, mvh, mvl;
    if (int e = get_tensor_map_2d_f16(k_hi, cols, rows, (int)cols, 64, &mkh)) return e;
    if (int e = get_tensor_map_2d_f16(k_lo, cols, rows, (int)cols, 64, &mkl)) return e;
    if (int e = get_tensor_map_2d_f16(vt_hi, rows, cols, (int)rows, 64, &mvh)) return e;
    if (int e = get_tensor_map_2d_f16(vt_lo, rows, cols, (int)rows, 64, &mvl)) return e;
    static int n_sm = 0;
    if (!n_sm) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); if (n_sm <= 0) n_sm = 148; }
    const long long n_items = (long long)n_win * (AH_WS / AH_BQ) * heads;
    const int slots = std::max(heads, (2 * n_sm / heads) * heads);            // two resident CTAs per SM, a multiple of `heads`
    const int grid = (int)std::min<long long>(n_items, slots);
    k_swin_attn_h<<<grid, AH_THREADS, AH_SMEM, st>>>(q, ldq, mkh, mkl, mvh, mvl, qb, relpos, heads, d_off, d_win_seq, d_win_idx,
                                                     shift, out, ldo, n_win);
    SCP_LAUNCHED();
    SCP_CUDA(cudaFreeAsync(buf, st));
    return SCP_OK;
}

}  // namespace scp

// kNN in learned feature space (dgcnn.py:10-28 on 144-/192-d activations) on the 5th-gen tensor cores.
//
// score(i,j) = 2 x_i.x_j - |x_j|^2 - |x_i|^2; the Gram tile x_i.x_j is a K-major x K-major tcgen05 GEMM of the
// token matrix with itself.  Neighbour sets must not move, so the dot products use the error-compensated split
// x = x_hi + x_lo (hi.hi + lo.hi + hi.lo: fp32-class accuracy) -- on the FP16 pipe: x_hi, x_lo are fp16 (11 + 11 mantissa
// bits, like the tf32 split) of X scaled by one power of two per call (max|X| -> [2^7, 2^8); the score undoes it exactly).
// kind::f16 runs at twice the tf32 rate and, more important here, the candidate tiles are HALF the bytes: every work item
// streams its whole window through L2 -> shared memory (ncu on the tf32 form: ~5 TB/s of L2 reads, the actual limiter).
// X is split ONCE per call by an elementwise kernel (the tiles are re-read ~64x from L2) and |x|^2 is fp32.
//
// One persistent CTA per work item = (window, 128-query tile):
//   * the query rows' X_hi / X_lo are written ONCE to tensor memory (tcgen05.st, lane = row) and serve as the A operand
//     of every MMA of the work item ("TS" form): only the candidate tiles cross L2 -> shared memory -> tensor core, which
//     halves the TMA traffic and takes the A reads (2/3 of an SS-form MMA's operand bytes at N = 64) off the
//     shared-memory port;
//   * candidate tiles of 64 rows stream through a 6-stage TMA ring; the 128 x 64 score tile accumulates in TMEM (2 stages);
//   * the epilogue warps (lane = query row) read the scores back with tcgen05.ld and keep, per row, a MIN-HEAP of the
//     kc = k + 8 best candidates in shared memory (4-ary, [entry][lane], conflict-free).  The heap root is the row's exact
//     running threshold, so a row accepts only ~kc (1 + ln(n / kc)) candidates over the whole scan and an accept costs
//     one sift-down (<= 5 levels).  (Earlier versions appended hits to per-row lists in global memory and compacted
//     them with a warp-wide radix select: 70 % of the kernel time went into those scattered 4-byte stores.)
// The kc survivors per row go to an exact float64 re-rank, which defines the final order and the tie rule.
#include <algorithm>
#include <stdlib.h>
#include <cuda_fp16.h>
#include "tc.cuh"

struct scp_seqs;

namespace scp {

constexpr int KT_BM = 128, KT_BN = 64, KT_BK = 64;                      // K blocks of 64 fp16 = 128-byte rows
// TWO CTAs per SM: the kernel is bound by its four epilogue warps (lane = query row is forced by the TMEM lane mapping, so
// a CTA cannot have more of them), each alone on its scheduler with nothing to hide its dependent-issue latency behind
// (ncu: 0.8 IPC per SM).  Half the TMEM (256 columns: one accumulator stage) and a 3-stage ring (one candidate tile) per
// CTA let a second CTA share the SM: two epilogue warps per scheduler, and each CTA's MMAs / TMA run under the other's
// epilogue.
// Round-2 timeline (clock64 stamps of CTA 0, `SCP_KNN_TRACE=1`; profiles/r02_knn_timeline.md) and the three single-CTA restructurings
// it led to, all measured SLOWER on the bench frames and therefore not kept:
//  * a (128-query, 64-candidate) tile takes 2200 cycles per SM for 36 MMAs; 1430 with the scan switched off -- that IS the tensor
//    pipe for this shape: N = 64 TS-form MMAs issued by ONE thread retire every 48-57 cycles (whatever the accumulator order),
//    by the two issuers of two co-resident CTAs (or two issuing warps of one CTA) every 40, never at the nominal 32;
//  * 256 query rows per CTA sharing every candidate tile (half the L2 bytes) changes nothing: L2 -> shared memory is not the
//    limiter (34 B/clk per SM at the rate above, cap ~42), and one issuer made it slower;
//  * four accumulator stages + two epilogue teams with a heap each (candidate tiles alternate between the teams, thresholds
//    shared, heaps merged at the end; two issuing warps) decouples the MMAs from the slowest epilogue warp and reaches the same
//    1430 without the scan, but every row then pays its accepts against two half-informed heaps: 44 ms per step against 36.
//  * eight epilogue warps with four threads per query row (tcgen05.ld.16x256b, one heap owner per row): no faster on the synthetic
//    walk, 11 % slower on the bench frames -- the scan and the accept rounds are bound by the SM's issue slots, which more warps
//    do not add to.
// What is left between 1430 and 2200 is the accept path itself (per-lane heap sift-downs, 150-250 dependent cycles per round,
// rounds = the maximum over the lanes of a warp).
constexpr int KT_STAGES = 3, KT_ACC = 1;
constexpr int KT_EXTRA = 8;                                    // approximate top-(k+8) is re-ranked exactly
constexpr int KT_MAXD = 192;                                   // A_hi + A_lo: 2 x 96 TMEM columns (two fp16 per column)
constexpr int KT_TILE_BYTES = KT_BN * KT_BK * 2;               // 8 KB: [64 candidates x 64 halfs]
constexpr int KT_STAGE_BYTES = 2 * KT_TILE_BYTES;              // B_hi | B_lo
constexpr uint32_t KT_TMEM_COLS = 256;
constexpr uint32_t KT_T_AH = KT_ACC * KT_BN, KT_T_AL = KT_T_AH + KT_MAXD / 2;   // TMEM: acc [0,64) | A_hi [64,160) | A_lo [160,256)
static_assert(KT_T_AL + KT_MAXD / 2 <= KT_TMEM_COLS, "TMEM budget");
// shared memory after the ring: barriers 256 B | candidate norms 4 x 64 f | score tiles 4 x [32][33] f | heaps 4 x [32][32] (f, i)
constexpr int KT_OFF_BAR = KT_STAGES * KT_STAGE_BYTES;
constexpr int KT_OFF_XC = KT_OFF_BAR + 256;
constexpr int KT_OFF_TR = KT_OFF_XC + 4 * 64 * 4;
constexpr int KT_OFF_HS = KT_OFF_TR + 4 * 32 * 33 * 4;
constexpr int KT_OFF_HI = KT_OFF_HS + 4 * 32 * 32 * 4;
constexpr int KT_SMEM = KT_OFF_HI + 4 * 32 * 32 * 4 + 1024;

// Candidate-tile visiting order for the query tile whose first 64-row tile is t0 (a query tile covers t0 and t0+1):
// own tiles first, then outwards by index distance, alternating left / right.  Tokens are in Morton order, so index
// distance tracks spatial distance: the heap threshold is tight after the first few tiles and the far tiles, visited
// last, hardly ever produce an accept (ascending order made every row accept ~190 candidates, this order ~60).
__host__ __device__ __forceinline__ int knn_tile_order(int i, int t0, int nt) {
    const int own = (t0 + 1 < nt) ? 2 : 1;
    if (i < own) return t0 + i;
    int j = i - own;                                                           // index among the other tiles
    const int L = t0, R = nt - t0 - own;                                       // tiles left of t0 / right of the own tiles
    const int m = L < R ? L : R;
    if (j < 2 * m) { const int k = (j >> 1) + 1; return (j & 1) ? t0 + own - 1 + k : t0 - k; }
    j -= 2 * m;
    return L > R ? t0 - m - 1 - j : t0 + own + m + j;
}

// max|X| of every SEQUENCE -> scales[2s] = 2^e with max * 2^e in [2^7, 2^8), scales[2s+1] = 2 * 2^(-2e) (the factor of the
// Gram term).  One scale per sequence (context window), not per call: the fp16 hi/lo parts of a window -- and with them its
// approximate candidate lists -- must not depend on which other windows share the ragged batch (the encoder batches the
// windows of whole frames, the decoder one level at a time, and both must see identical neighbour sets).
// One block per 128-row tile of the tile tables; 8 warps x 16 rows.
__global__ void __launch_bounds__(256) k_knn_absmax(const float* __restrict__ X, long long ldx, int d,
                                                     const long long* __restrict__ seq_off, const int* __restrict__ tile_seq,
                                                     const int* __restrict__ tile_start, unsigned* __restrict__ mx) {
    const int s = tile_seq[blockIdx.x];
    const long long base = seq_off[s] + tile_start[blockIdx.x];
    const int rows = (int)min((long long)KT_BM, seq_off[s + 1] - base);
    float m = 0.f;
    for (int r = threadIdx.x >> 5; r < rows; r += 8)
        for (int c = threadIdx.x & 31; c < d; c += 32) m = fmaxf(m, fabsf(X[(base + r) * ldx + c]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(mx + s, __float_as_uint(m));   // non-negative floats order like their bits
}
__global__ void __launch_bounds__(256) k_knn_scale(const unsigned* __restrict__ mx, int n_seq, float* __restrict__ scales) {
    const int s = blockIdx.x * 256 + threadIdx.x;
    if (s >= n_seq) return;
    const float m = __uint_as_float(mx[s]);
    int e = 0;
    if (m > 0.f && m < 3e38f) { frexpf(m, &e); e = 8 - e; }
    e = max(-60, min(60, e));
    scales[2 * s] = ldexpf(1.0f, e);
    scales[2 * s + 1] = ldexpf(2.0f, -2 * e);
}

__global__ void __launch_bounds__(256) k_split_rows(const float* __restrict__ X, long long ldx, int d, long long row0,
                                                     const long long* __restrict__ seq_off, const int* __restrict__ tile_seq,
                                                     const int* __restrict__ tile_start, const float* __restrict__ scales,
                                                     __half* __restrict__ hi, __half* __restrict__ lo, float* __restrict__ xx) {
    const int lane = threadIdx.x & 31;
    const int sq = tile_seq[blockIdx.x];
    const long long base = seq_off[sq] + tile_start[blockIdx.x];
    const int rows = (int)min((long long)KT_BM, seq_off[sq + 1] - base);
    const float sc = scales[2 * sq];
    for (int r = threadIdx.x >> 5; r < rows; r += 8) {
        const long long g = base + r, row = g - row0;               // global row; row inside the call's buffers
        float s = 0.f;
        for (int c = lane; c < d; c += 32) {
            const float v = X[g * ldx + c];
            const float vs = v * sc;
            const __half h = __float2half_rn(vs);
            hi[row * d + c] = h;
            lo[row * d + c] = __float2half_rn(vs - __half2float(h));
            s = fmaf(v, v, s);
        }
        s = warp_sum(s);
        if (lane == 0) xx[row] = s;
    }
}

__global__ void __launch_bounds__(256, 2) k_knn_tc(const __grid_constant__ CUtensorMap tmHi,
                                                    const __grid_constant__ CUtensorMap tmLo,
                                                    const __half* __restrict__ xhi, const __half* __restrict__ xlo,
                                                    const float* __restrict__ scales, const float* __restrict__ xx, const long long* __restrict__ seq_off,
                                                    const int* __restrict__ tile_seq, const int* __restrict__ tile_start,
                                                    int n_work, long long row0, int d, int kc, int* __restrict__ idx_out, int dbg,
                                                    long long* __restrict__ trace) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + KT_OFF_BAR);
    uint64_t* empty = full + KT_STAGES;
    uint64_t* tfull = empty + KT_STAGES;
    uint64_t* tempty = tfull + 2;
    uint64_t* a_ready = tempty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_kb = (d + KT_BK - 1) / KT_BK;
    constexpr uint32_t TMEM_COLS = KT_TMEM_COLS;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmHi)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmLo)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < KT_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < KT_ACC; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
        mbar_init(a_ready, 4);
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
        // ---------------- TMA producer: candidate tiles only (all lanes loop, the elected lane issues) ----------------
        int stage = 0; uint32_t phase = 0;
        for (int wk = blockIdx.x; wk < n_work; wk += gridDim.x) {
            const int s = tile_seq[wk];
            const long long base = seq_off[s] - row0;                   // row of the window inside the split copies
            const int n = (int)(seq_off[s + 1] - seq_off[s]);
            const int nt = (n + KT_BN - 1) / KT_BN;
            for (int ci = 0; ci < nt; ++ci) {
                const int c0 = knn_tile_order(ci, tile_start[wk] / KT_BN, nt) * KT_BN;
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* b = smem + stage * KT_STAGE_BYTES;
                    if (elect_one()) {
                        mbar_expect_tx(&full[stage], KT_STAGE_BYTES);
                        tma_load_2d(b, &tmHi, &full[stage], kb * KT_BK, (int)(base + c0));
                        tma_load_2d(b + KT_TILE_BYTES, &tmLo, &full[stage], kb * KT_BK, (int)(base + c0));
                    }
                    __syncwarp();
                    if (++stage == KT_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer: all lanes run the loop, the elected lane issues ----------------
        const uint32_t idesc = (1u << 4) | ((uint32_t)(KT_BN >> 3) << 17) | ((uint32_t)(KT_BM >> 4) << 24);     // f16 x f16 -> f32
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0, a_phase = 0;
        int tr_n = 0;                                      // development aid (SCP_KNN_TRACE=1): clock stamps of CTA 0
        const bool TR = trace && blockIdx.x == 0;
        for (int wk = blockIdx.x; wk < n_work; wk += gridDim.x) {
            const int s = tile_seq[wk];
            const int n = (int)(seq_off[s + 1] - seq_off[s]);
            const int nt = (n + KT_BN - 1) / KT_BN;
            mbar_wait(a_ready, a_phase);                                // this work item's query rows are in TMEM
            a_phase ^= 1;
            for (int ci = 0; ci < nt; ++ci) {
                if (TR && lane == 0 && tr_n < 900) trace[tr_n++] = clock64();            // tile: before the accumulator wait
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                if (TR && lane == 0 && tr_n < 900) trace[tr_n++] = clock64();            // tile: accumulator free
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * KT_BN);
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(&full[stage], phase);
                    if (TR && lane == 0 && tr_n < 900 && (kb == 0 || kb == n_kb - 1)) trace[tr_n++] = clock64();   // first / last K block landed
                    tc_fence_after();
                    const uint8_t* b = smem + stage * KT_STAGE_BYTES;
                    const uint64_t dbh = make_smem_desc(b), dbl = make_smem_desc(b + KT_TILE_BYTES);
                    const uint32_t ah = tmem_base + KT_T_AH + (uint32_t)(kb * 32), al = tmem_base + KT_T_AL + (uint32_t)(kb * 32);
                    const int kk_n = min(4, (d - kb * KT_BK + 15) >> 4);     // d = 144: the last K block holds 16 channels, the rest of
                                                                             // the box is TMA zero fill -- no MMAs for it
                    if (elect_one()) {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {                     // 16 fp16 = 8 TMEM columns of A = 2 descriptor units of B
                            if (kk >= kk_n) break;
                            const uint64_t o = (uint64_t)(2 * kk);
                            tc_mma_f16_ts(d_tmem, ah + 8u * kk, dbh + o, idesc, (kb | kk) ? 1u : 0u);
                            tc_mma_f16_ts(d_tmem, al + 8u * kk, dbh + o, idesc, 1u);
                            tc_mma_f16_ts(d_tmem, ah + 8u * kk, dbl + o, idesc, 1u);
                        }
                        tc_commit(&empty[stage]);
                        if (kb == n_kb - 1) tc_commit(&tfull[acc]);
                    }
                    __syncwarp();
                    if (++stage == KT_STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == KT_ACC) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ---------------- query loader + epilogue: warp w owns query rows 32w..32w+31, lane = row ----------------
        const int w = warp - 4;
        float* xcs = reinterpret_cast<float*>(smem + KT_OFF_XC) + w * 64;           // candidate norms of the current tile
        float* tr = reinterpret_cast<float*>(smem + KT_OFF_TR) + w * (32 * 33);     // this warp's parked score tile
        float2* hq = reinterpret_cast<float2*>(smem + KT_OFF_HS) + w * (32 * 32) + lane;  // heap (score, candidate), entry e at hq[32 e]
        const uint32_t t_lane = tmem_base + ((uint32_t)(w * 32) << 16);
        int acc = 0; uint32_t acc_phase = 0;
        const bool ETR = trace && blockIdx.x == 0 && w == 0 && lane == 0;
        int e_n = 0;
        for (int wk = blockIdx.x; wk < n_work; wk += gridDim.x) {
            const int s = tile_seq[wk];
            const long long gbase = seq_off[s];
            const int n = (int)(seq_off[s + 1] - gbase);
            const int q = tile_start[wk] + w * 32 + lane;                 // this lane's query row inside the window
            const bool rowv = q < n;
            // (every MMA of the previous work item has retired: this warp saw its last accumulator)
            {
                const __half* ph = xhi + (gbase - row0 + q) * (long long)d;
                const __half* pl = xlo + (gbase - row0 + q) * (long long)d;
                for (int kb = 0; kb < n_kb; ++kb) {
                    uint32_t h[32], l[32];                                // 64 halfs of the K block, two per word (even k low)
#pragma unroll
                    for (int c = 0; c < 32; c += 4) {
                        const int col = kb * KT_BK + 2 * c;
                        uint4 vh = make_uint4(0u, 0u, 0u, 0u), vl = vh;
                        if (rowv && col < d) {                            // d % 8 == 0: eight halfs never straddle the end
                            vh = *reinterpret_cast<const uint4*>(ph + col);
                            vl = *reinterpret_cast<const uint4*>(pl + col);
                        }
                        h[c] = vh.x; h[c + 1] = vh.y; h[c + 2] = vh.z; h[c + 3] = vh.w;
                        l[c] = vl.x; l[c + 1] = vl.y; l[c + 2] = vl.z; l[c + 3] = vl.w;
                    }
                    tc_st32(t_lane + KT_T_AH + (uint32_t)(kb * 32), h);
                    tc_st32(t_lane + KT_T_AL + (uint32_t)(kb * 32), l);
                }
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(a_ready);
            }
            const float two_g = __ldg(scales + 2 * s + 1);                       // 2 / scale^2: the Gram tile is of the scaled rows
            for (int e = 0; e < kc; ++e) hq[32 * e] = make_float2(-INFINITY, __int_as_float(-1));
            float th = (rowv && dbg == 0) ? -INFINITY : INFINITY;          // rows past the window never take a candidate
            const int nt = (n + KT_BN - 1) / KT_BN;
            const float* xw = xx + (gbase - row0);
            // candidate norms of a tile are fetched one tile ahead, so the global loads are never in flight at a barrier
            int c0 = knn_tile_order(0, tile_start[wk] / KT_BN, nt) * KT_BN;
            float nx0 = c0 + lane < n ? __ldg(xw + c0 + lane) : INFINITY;      // +inf norm -> score -inf
            float nx1 = c0 + lane + 32 < n ? __ldg(xw + c0 + lane + 32) : INFINITY;
            for (int ci = 0; ci < nt; ++ci) {
                __syncwarp();
                xcs[lane] = nx0; xcs[lane + 32] = nx1;
                __syncwarp();
                const int c0_next = ci + 1 < nt ? knn_tile_order(ci + 1, tile_start[wk] / KT_BN, nt) * KT_BN : 0;
                if (ci + 1 < nt) {
                    nx0 = c0_next + lane < n ? __ldg(xw + c0_next + lane) : INFINITY;
                    nx1 = c0_next + lane + 32 < n ? __ldg(xw + c0_next + lane + 32) : INFINITY;
                }
                if (ETR && e_n < 900) trace[1024 + e_n++] = clock64();                  // tile: before the score wait
                mbar_wait(&tfull[acc], acc_phase);
                if (ETR && e_n < 900) trace[1024 + e_n++] = clock64();                  // tile: scores ready
                tc_fence_after();
                uint32_t r[2][32];
                tc_ld32_nowait(t_lane + (uint32_t)(acc * KT_BN), r[0]);
                tc_ld32_nowait(t_lane + (uint32_t)(acc * KT_BN) + 32u, r[1]);
                tc_wait_ld();
                tc_fence_before();                                         // scores are in registers: the stage can be refilled
                __syncwarp();
                if (lane == 0) mbar_arrive_relaxed(&tempty[acc]);
                if (ETR && e_n < 900) trace[1024 + e_n++] = clock64();                  // tile: scores in registers, stage returned
                if (++acc == KT_ACC) { acc = 0; acc_phase ^= 1; }
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int cb = 32 * hh;
                    if (c0 + cb >= n || dbg >= 2) break;                   // warp-uniform
                    float smax = -INFINITY;
#pragma unroll
                    for (int j4 = 0; j4 < 32; j4 += 4) {
                        const float4 xc = *reinterpret_cast<const float4*>(xcs + cb + j4);
                        const float xcv[4] = {xc.x, xc.y, xc.z, xc.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            // 2 g - |c|^2 (2 g / scale^2 is exact): the score up to the row's own -|q|^2, which does not change
                            // the order inside a row -- the heap and the threshold live in this shifted scale (the scan is bound
                            // by issue slots: two instructions per candidate instead of three)
                            const float sc = fmaf(two_g, __uint_as_float(r[hh][j4 + e]), -xcv[e]);
                            r[hh][j4 + e] = __float_as_uint(sc);
                            smax = fmaxf(smax, sc);
                        }
                    }
                    if (__any_sync(0xffffffffu, smax > th)) {
                        // Accepts are per-lane events (lane = row): every lane builds the bit mask of its own candidates
                        // above the threshold, parks its 32 scores in shared memory ([lane][33], conflict-free) and pops
                        // only its own bits; the warp iterates max-over-lanes(accepts) times, usually 1-2.
                        uint32_t hm = 0u;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            hm |= (__uint_as_float(r[hh][j]) > th) ? (1u << j) : 0u;
                            tr[lane * 33 + j] = __uint_as_float(r[hh][j]);
                        }
                        while (hm) {
                            const int j = __ffs(hm) - 1;
                            hm &= hm - 1;
                            const float v = tr[lane * 33 + j];
                            if (!(v > th)) continue;                       // the threshold moved since the mask was built
                            const int vid = c0 + cb + j;
                            int i = 0;                                     // replace the root (the row's kc-th best) and sift down
                            for (;;) {                                     // 4-ary heap: 3 levels for kc <= 32, the four child
                                const int c0 = 4 * i + 1;                  // loads of a level are independent (one smem latency)
                                if (c0 >= kc) break;
                                const float2 inf2 = make_float2(INFINITY, 0.f);
                                const float2 e0 = hq[32 * c0];
                                const float2 e1 = c0 + 1 < kc ? hq[32 * (c0 + 1)] : inf2;
                                const float2 e2 = c0 + 2 < kc ? hq[32 * (c0 + 2)] : inf2;
                                const float2 e3 = c0 + 3 < kc ? hq[32 * (c0 + 3)] : inf2;
                                const bool b01 = e1.x < e0.x, b23 = e3.x < e2.x;
                                const float2 m01 = b01 ? e1 : e0, m23 = b23 ? e3 : e2;
                                const bool bm = m23.x < m01.x;
                                const float2 em = bm ? m23 : m01;
                                if (!(em.x < v)) break;
                                hq[32 * i] = em;
                                i = c0 + (bm ? (b23 ? 3 : 2) : (b01 ? 1 : 0));
                            }
                            hq[32 * i] = make_float2(v, __int_as_float(vid));
                            th = hq[0].x;
                        }
                    }
                }
                c0 = c0_next;
                if (ETR && e_n < 900) trace[1024 + e_n++] = clock64();                  // tile: scanned
            }
            if (rowv) {                                                    // kc survivors (unordered) for the exact re-rank
                int* dst = idx_out + (gbase - row0 + q) * 32;
#pragma unroll 4
                for (int e = 0; e < 32; ++e) {
                    const int id = e < kc ? __float_as_int(hq[32 * e].y) : -1;
                    dst[e] = id < 0 ? -1 : (int)(gbase + id);
                }
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// Exact re-rank: the tensor-core scores are accurate to ~1e-6 relative, which is enough to find the top-(k+8) but
// not to order near-ties reproducibly.  The final neighbour list is defined by the EXACT squared distance of the
// float32 rows (float64 accumulation in channel order), ties -> lowest index, independent of the engine.
__global__ void __launch_bounds__(256) k_knn_rerank(const float* __restrict__ X, long long ldx, int d, long long row0,
                                                     long long n, const int* __restrict__ cand, int kc, int k,
                                                     int* __restrict__ idx_out) {
    // One warp per query, one lane per candidate.  A lane walking its own candidate row with 16-byte loads makes every load
    // instruction touch 32 different rows (ncu: L1 at 99 %, 1536 wavefronts per query); instead the warp copies 32-channel
    // chunks of the candidate rows into shared memory with coalesced loads (eight lanes per row) and each lane then reads
    // its row from there -- same values, same channel order, same float64 arithmetic.
    __shared__ float s_c[8][32][33];
    __shared__ double s_q[8][32];                                         // query chunk, converted to float64 ONCE per warp
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long r = (long long)blockIdx.x * 8 + wid;
    if (r >= n) return;                                                   // warp-uniform; no block-wide barrier below
    const long long row = row0 + r;
    const int c = lane < kc ? cand[r * kc + lane] : -1;
    double dist = INFINITY;
    const float* a = X + row * ldx;
    const bool v4 = (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0) && (d % 4 == 0);
    if (v4) {
        float (*sc)[33] = s_c[wid];
        double acc = 0.0;
        const int sub = lane >> 3, q4 = (lane & 7) << 2;                  // this lane copies 4 floats of candidate 4 j + sub
        for (int c0 = 0; c0 < d; c0 += 32) {
            const int jn = min(32, d - c0);                               // (d = 144: the last chunk is 16 wide)
            __syncwarp();
            s_q[wid][lane] = lane < jn ? (double)a[c0 + lane] : 0.0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int cj = __shfl_sync(0xffffffffu, c, 4 * j + sub);
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (cj >= 0 && q4 < jn) v = __ldg(reinterpret_cast<const float4*>(X + (long long)cj * ldx + c0 + q4));
                float* dst = &sc[4 * j + sub][q4];
                dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
            }
            __syncwarp();
            if (c >= 0) {
#pragma unroll 8
                for (int j = 0; j < jn; ++j) {
                    const double df = s_q[wid][j] - (double)sc[lane][j];
                    acc = __dadd_rn(acc, __dmul_rn(df, df));
                }
            }
        }
        if (c >= 0) dist = acc;
    } else if (c >= 0) {
        const float* b = X + (long long)c * ldx;
        double acc = 0.0;
        for (int j = 0; j < d; ++j) { const double df = (double)a[j] - (double)b[j]; acc = __dadd_rn(acc, __dmul_rn(df, df)); }
        dist = acc;
    }
    // rank of this candidate among the warp's candidates by (dist, index)
    int rank = 0;
    for (int o = 0; o < 32; ++o) {
        const double od = __shfl_sync(0xffffffffu, dist, o);
        const int oc = __shfl_sync(0xffffffffu, c, o);
        if (oc >= 0 && (od < dist || (od == dist && oc < c))) ++rank;
    }
    if (c >= 0 && rank < k) idx_out[row * k + rank] = c;
    const int valid = __popc(__ballot_sync(0xffffffffu, c >= 0));
    if (lane >= valid && lane < k) idx_out[row * k + lane] = (int)row;         // short window: repeat self
}

// host ------------------------------------------------------------------------------------------
bool knn_tc_ok(int d, int k) { return d >= 32 && d <= KT_MAXD && d % 8 == 0 && k >= 1 && k + KT_EXTRA <= 32; }

int knn_tc(const float* d_x, long long ldx, int d, const long long* h_off, int n_seq, const long long* d_off,
           const int* d_tile_seq, const int* d_tile_start, int n_work, int k, int* d_idx, cudaStream_t st) {
    const long long row0 = h_off[0], total = h_off[n_seq] - h_off[0];
    __half *hi = nullptr, *lo = nullptr;
    float *xx = nullptr, *scales = nullptr;
    int* cand = nullptr;
    const int kc = k + KT_EXTRA;
    SCP_CUDA(malloc_async((void**)&cand, (size_t)total * 32 * 4 + 1024, st));
    SCP_CUDA(malloc_async((void**)&hi, (size_t)total * d * 2 + 1024, st));
    SCP_CUDA(malloc_async((void**)&lo, (size_t)total * d * 2 + 1024, st));
    SCP_CUDA(malloc_async((void**)&xx, (size_t)total * 4 + 1024, st));
    // per sequence: [n_seq][2] floats (scale, Gram factor) then [n_seq] running maxima
    const size_t sc_bytes = (size_t)n_seq * 12 + 64;
    SCP_CUDA(malloc_async((void**)&scales, sc_bytes, st));
    SCP_CUDA(cudaMemsetAsync(scales, 0, sc_bytes, st));
    unsigned* mx = reinterpret_cast<unsigned*>(scales + 2 * (size_t)n_seq);
    k_knn_absmax<<<n_work, 256, 0, st>>>(d_x, ldx, d, d_off, d_tile_seq, d_tile_start, mx);
    SCP_LAUNCHED();
    k_knn_scale<<<(unsigned)cdiv(n_seq, 256), 256, 0, st>>>(mx, n_seq, scales);
    SCP_LAUNCHED();
    k_split_rows<<<n_work, 256, 0, st>>>(d_x, ldx, d, row0, d_off, d_tile_seq, d_tile_start, scales, hi, lo, xx);
    SCP_LAUNCHED();
    CUtensorMap mh, ml;
    // the split buffers are transient: encode their maps every call (pointer reuse would alias a cached map only
    // when shape and address are identical, which is then also correct)
    if (int e = get_tensor_map_2d_f16(hi, d, total, d, KT_BN, &mh)) return e;
    if (int e = get_tensor_map_2d_f16(lo, d, total, d, KT_BN, &ml)) return e;
    static bool attr = false;
    if (!attr) { SCP_CUDA(cudaFuncSetAttribute(k_knn_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, KT_SMEM)); attr = true; }
    static int n_sm = 0;
    if (!n_sm) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); if (n_sm <= 0) n_sm = 148; }
    const int grid = std::min(n_work, 2 * n_sm);                      // two resident CTAs per SM
    const int dbg = getenv("SCP_KNN_DBG") ? atoi(getenv("SCP_KNN_DBG")) : 0;          // timing experiments only
    static long long* d_trace = nullptr;
    static const bool want_trace = getenv("SCP_KNN_TRACE") != nullptr;
    if (want_trace && !d_trace) cudaMalloc(&d_trace, 2048 * 8);
    if (want_trace) cudaMemsetAsync(d_trace, 0, 2048 * 8, st);
    k_knn_tc<<<grid, 256, KT_SMEM, st>>>(mh, ml, hi, lo, scales, xx, d_off, d_tile_seq, d_tile_start, n_work, row0, d, kc, cand, dbg,
                                         want_trace ? d_trace : nullptr);
    SCP_LAUNCHED();
    if (want_trace) {                                    // development aid: per-tile timeline of CTA 0 (MMA warp / epilogue warp 0)
        static long long hh[2048];
        cudaStreamSynchronize(st);
        cudaMemcpy(hh, d_trace, sizeof(hh), cudaMemcpyDeviceToHost);
        auto avg = [&](int base, int stride, int a, int b, int t0, int t1) {
            double s = 0; int n = 0;
            for (int t = t0; t < t1; ++t) { const long long x = hh[base + t * stride + a], y = hh[base + t * stride + b]; if (x && y) { s += (double)(y - x); ++n; } }
            return n ? s / n : 0.0;
        };
        for (int lo_t = 2; lo_t < 200; lo_t += 66)
            fprintf(stderr, "knn trace d=%d tiles %d-%d: MMA warp  acc-wait %.0f  kb0-wait %.0f  kb0->kbL %.0f  period %.0f | epilogue  score-wait %.0f  ld %.0f  scan %.0f  period %.0f\n",
                    d, lo_t, lo_t + 66, avg(0, 4, 0, 1, lo_t, lo_t + 66), avg(0, 4, 1, 2, lo_t, lo_t + 66), avg(0, 4, 2, 3, lo_t, lo_t + 66),
                    avg(0, 4, 0, 4, lo_t, lo_t + 66), avg(1024, 4, 0, 1, lo_t, lo_t + 66), avg(1024, 4, 1, 2, lo_t, lo_t + 66),
                    avg(1024, 4, 2, 3, lo_t, lo_t + 66), avg(1024, 4, 0, 4, lo_t, lo_t + 66));
    }
    k_knn_rerank<<<(unsigned)cdiv(total, 8), 256, 0, st>>>(d_x, ldx, d, row0, total, cand, 32, k, d_idx);
    SCP_LAUNCHED();
    SCP_CUDA(cudaFreeAsync(cand, st));
    SCP_CUDA(cudaFreeAsync(hi, st));
    SCP_CUDA(cudaFreeAsync(lo, st));
    SCP_CUDA(cudaFreeAsync(xx, st));
    SCP_CUDA(cudaFreeAsync(scales, st));
    return SCP_OK;
}

}  // namespace scp

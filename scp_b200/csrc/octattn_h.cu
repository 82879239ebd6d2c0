// Two-stream causal attention of SCP-OctAttention (attention_model.py:58-95) on the FP16 tensor pipe (tcgen05 / TMEM / TMA).
//
// Per head (4 x 150 dims, context <= 1024 tokens) the reference materialises S = QU K^T / sqrt(150) once and uses it twice:
//   known stream:    softmax(S + causal mask) V
//   unknown stream:  the same scores with the DIAGONAL replaced by QU . KU (the token's own key without its occupancy),
//                    softmax, then sum_{j<i} p_ij V_j + p_ii VU_i.
// Both streams therefore share everything strictly below the diagonal.  This kernel runs ONE flash-style pass over the
// strictly-lower triangle on the tensor cores -- S = QU K^T and O = P V as error-compensated products (x = x_hi + x_lo in
// fp16, three tcgen05.mma per product: hi.hi + lo.hi + hi.lo, fp32 accumulation in TMEM, the scheme of attn_h.cu) with an
// online softmax in the log2 domain -- and adds the two diagonal terms (two 150-dim dot products and one V / VU row per
// query, fp32 SIMT) in the epilogue, where each stream gets its own maximum and normaliser.  No subtraction of a diagonal
// term that was first summed in, hence no cancellation when a token attends mostly to itself.
//
// Shapes: head dim 150 is padded to 160 for the MMAs (K-steps of 16; Q / K storage is padded to 192 so that the 128-byte
// swizzled K tiles are three [64 keys x 64 dims] blocks); a CTA = (sequence, head, 128 query rows), key chunks of 64, only
// the chunks at or below the diagonal are visited (2 qb + 2 of them for query block qb).
// TMEM map (512 columns, one CTA per SM): Q_hi [0,96) | Q_lo [96,192) | S/P buffer b at [192 + 64 b, +64): S fp32, then
//                         P_hi [+0,+32) P_lo [+32,+64) | O [320,480)
// Warp roles (320 threads): warps 0-7 softmax / output (two threads per query row: TMEM lane quarter = warp & 3, column
// half = warp >> 2), warp 8 MMA issuer + TMEM owner, warp 9 TMA producer.
// K and V are prepared once per launch by k_octattn_prep (fp16 hi/lo split, V transposed, sequences padded to 64-key tiles).
#include <stdlib.h>
#include <vector>
#include <cuda_fp16.h>
#include "tc.cuh"

namespace scp {

constexpr int OH_HD = 150, OH_HDP = 160, OH_QP = 192, OH_BQ = 128, OH_BK = 64;
constexpr int OH_KBLK = 64 * 128;                  // [64 keys x 64 dims] fp16, 8 KB
constexpr int OH_KTILE = 3 * OH_KBLK;              // 192 padded dims
constexpr int OH_KSTAGE = 2 * OH_KTILE;            // hi | lo            48 KB
constexpr int OH_VTILE = OH_HDP * 128;             // [160 dims x 64 keys] fp16, 20 KB
constexpr int OH_VSTAGE = 2 * OH_VTILE;            // hi | lo            40 KB
constexpr int OH_OFF_K = 0;
constexpr int OH_OFF_V = OH_OFF_K + 2 * OH_KSTAGE;
constexpr int OH_OFF_XCH = OH_OFF_V + 2 * OH_VSTAGE;      // 5 x [2][128] floats: chunk max (2 slots), row sums, two dot products
constexpr int OH_OFF_BAR = OH_OFF_XCH + 5 * 1024;
constexpr int OH_SMEM = OH_OFF_BAR + 256 + 1024;
constexpr int OH_THREADS = 320;
constexpr uint32_t OH_TMEM_COLS = 512;
constexpr uint32_t OH_T_QH = 0, OH_T_QL = 96, OH_T_SP = 192, OH_T_O = 320;
constexpr float OH_LOG2E = 1.4426950408889634f;

__device__ __forceinline__ void oh_split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
    float h0, h1;
    asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(h0), "=f"(h1) : "r"(hi));
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(x1 - h1), "f"(x0 - h0));
}
__device__ __forceinline__ float oh_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void oh_pair_sync(int quarter) {
    switch (quarter) {
        case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
        case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
        case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
        default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
    }
}
__device__ __forceinline__ void oh_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}

// K / V preparation: block = (64-token tile of a sequence, head).  The padded row of token j of a sequence is
// 64 * (tile index) + j % 64 (tiles are numbered over all sequences: scp_seqs' 64-token tile table), so a key chunk never
// straddles two sequences; rows behind the end of a sequence and dims >= 150 are zero.
//   k_hi / k_lo  [n_tile * 64][heads * 192]       vt_hi / vt_lo  [heads * 160][n_tile * 64]
__global__ void __launch_bounds__(256) k_octattn_prep(const float* __restrict__ K, const float* __restrict__ V, long long ld,
                                                       int heads, const long long* __restrict__ seq_off,
                                                       const int* __restrict__ tile_seq, const int* __restrict__ tile_start,
                                                       int n_tile, __half* __restrict__ k_hi, __half* __restrict__ k_lo,
                                                       __half* __restrict__ vt_hi, __half* __restrict__ vt_lo) {
    __shared__ float sv[64][OH_HDP + 1];                                   // V tile [key][dim] for the transpose
    const int tile = blockIdx.x, h = blockIdx.y;
    const int s = tile_seq[tile];
    const long long base = seq_off[s];
    const int S = (int)(seq_off[s + 1] - base);
    const int j0 = tile_start[tile];
    const long long prow0 = (long long)tile * 64;
    const long long ldp = (long long)heads * OH_QP, ldt = (long long)n_tile * 64;
    for (int unit = threadIdx.x; unit < 64 * (OH_QP / 2); unit += 256) {   // key r, dims 2c, 2c+1
        const int r = unit / (OH_QP / 2), c = (unit % (OH_QP / 2)) * 2;
        const bool ok = j0 + r < S && c < OH_HD;
        float2 kv = make_float2(0.f, 0.f), vv = kv;
        if (ok) {
            kv = __ldg(reinterpret_cast<const float2*>(K + (base + j0 + r) * ld + h * OH_HD + c));
            vv = __ldg(reinterpret_cast<const float2*>(V + (base + j0 + r) * ld + h * OH_HD + c));
        }
        uint32_t hi, lo;
        oh_split2(kv.x, kv.y, hi, lo);
        const long long o = (prow0 + r) * ldp + h * OH_QP + c;
        *reinterpret_cast<uint32_t*>(k_hi + o) = hi;
        *reinterpret_cast<uint32_t*>(k_lo + o) = lo;
        if (c < OH_HDP) { sv[r][c] = vv.x; sv[r][c + 1] = vv.y; }
    }
    __syncthreads();
    for (int unit = threadIdx.x; unit < OH_HDP * 32; unit += 256) {        // dim d, keys 2k, 2k+1
        const int d = unit >> 5, k2 = (unit & 31) * 2;
        uint32_t hi, lo;
        oh_split2(sv[k2][d], sv[k2 + 1][d], hi, lo);
        const long long o = ((long long)h * OH_HDP + d) * ldt + prow0 + k2;
        *reinterpret_cast<uint32_t*>(vt_hi + o) = hi;
        *reinterpret_cast<uint32_t*>(vt_lo + o) = lo;
    }
}

__global__ void __launch_bounds__(OH_THREADS, 1) k_octattn_attn_h(const float* __restrict__ QU, const float* __restrict__ K,
                                                                   const float* __restrict__ KU, const float* __restrict__ V,
                                                                   const float* __restrict__ VU, long long ld,
                                                                   const __grid_constant__ CUtensorMap tmKh,
                                                                   const __grid_constant__ CUtensorMap tmKl,
                                                                   const __grid_constant__ CUtensorMap tmVh,
                                                                   const __grid_constant__ CUtensorMap tmVl,
                                                                   int heads, const long long* __restrict__ seq_off,
                                                                   const int* __restrict__ qt_seq, const int* __restrict__ qt_start,
                                                                   const int* __restrict__ first_tile,
                                                                   float* __restrict__ O, float* __restrict__ OU, long long ldo) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* s_xch = reinterpret_cast<float*>(sm + OH_OFF_XCH);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + OH_OFF_BAR);
    uint64_t* k_full = bars;            // [2] K chunk landed                                   (TMA, expect_tx)
    uint64_t* v_full = bars + 2;        // [2] V chunk landed                                   (TMA, expect_tx)
    uint64_t* s_full = bars + 4;        // [2] S chunk in TMEM, K stage free                    (tcgen05.commit)
    uint64_t* p_full = bars + 6;        // [2] P chunk written to TMEM                          (8 softmax warps)
    uint64_t* pv_done = bars + 8;       // [2] PV of the chunk retired: O updated, V stage free (tcgen05.commit)
    uint64_t* o_ready = bars + 10;      // [1] O rescaled for the next chunk's maximum          (8 softmax warps)
    uint64_t* q_full = bars + 11;       // [1] Q_hi/Q_lo written to TMEM                        (8 softmax warps)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int h = blockIdx.y;
    const int s = qt_seq[blockIdx.x];
    const long long base = seq_off[s];
    const int S = (int)(seq_off[s + 1] - base);
    const int q0 = qt_start[blockIdx.x];                                   // first query row of the block inside its sequence
    const int nc = min(q0 / OH_BK + 2, (S + OH_BK - 1) / OH_BK);           // key chunks at or below the diagonal
    const int tile0 = first_tile[s];                                       // padded row of key j: 64 * (tile0 + j / 64) + j % 64

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
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(OH_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < 8) {
        // ---------------- softmax + output: TWO threads per query row ----------------
        const int quarter = warp & 3, half = warp >> 2;
        const int row = quarter * 32 + lane;
        const int u = q0 + row;                                            // query index inside the sequence
        const bool valid = u < S;
        const uint32_t trow = tmem + ((uint32_t)(quarter * 32) << 16);
        const float qs = rsqrtf((float)OH_HD) * OH_LOG2E;                  // 1/sqrt(150) and log2(e): scores live in the log2 domain
        float dk = 0.f, dku = 0.f;                                         // this half's part of QU . K_own and QU . KU_own
        {   // half a Q row (dims [96 half, +96), zero behind dim 150) -> TMEM as fp16 hi/lo pairs
            uint32_t hi[48], lo[48];
            const float* qrow = QU + (base + u) * ld + h * OH_HD;
            const float* krow = K + (base + u) * ld + h * OH_HD;
            const float* kurow = KU + (base + u) * ld + h * OH_HD;
#pragma unroll
            for (int c = 0; c < 48; ++c) {
                const int d = half * 96 + 2 * c;
                float2 q = make_float2(0.f, 0.f);
                if (valid && d < OH_HD) {
                    q = __ldg(reinterpret_cast<const float2*>(qrow + d));
                    const float2 a = __ldg(reinterpret_cast<const float2*>(krow + d)), b = __ldg(reinterpret_cast<const float2*>(kurow + d));
                    dk = fmaf(q.x, a.x, fmaf(q.y, a.y, dk));
                    dku = fmaf(q.x, b.x, fmaf(q.y, b.y, dku));
                }
                oh_split2(q.x * qs, q.y * qs, hi[c], lo[c]);
            }
            tc_st32(trow + OH_T_QH + (uint32_t)(half * 48), reinterpret_cast<const uint32_t(&)[32]>(hi[0]));
            tc_st16(trow + OH_T_QH + (uint32_t)(half * 48 + 32), hi + 32);
            tc_st32(trow + OH_T_QL + (uint32_t)(half * 48), reinterpret_cast<const uint32_t(&)[32]>(lo[0]));
            tc_st16(trow + OH_T_QL + (uint32_t)(half * 48 + 32), lo + 32);
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(q_full);
        }
        float m_run = -INFINITY, l_run = 0.f;
#pragma unroll 1
        for (int i = 0; i < nc; ++i) {
            const int b = i & 1, n = i >> 1;
            const uint32_t t_sp = trow + OH_T_SP + (uint32_t)(b * 64);
            mbar_wait(&s_full[b], n & 1);
            tc_fence_after();
            uint32_t r[32];
            tc_ld32(t_sp + (uint32_t)(half * 32), r);
            float cmax = -INFINITY;
            if (i * OH_BK + OH_BK <= q0) {                                 // chunk entirely below the block's first row: no mask
#pragma unroll
                for (int j = 0; j < 32; ++j) cmax = fmaxf(cmax, __uint_as_float(r[j]));
            } else {                                                       // strictly-lower triangle: key < query
                const int jb = i * OH_BK + half * 32;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float v = (jb + j < u) ? __uint_as_float(r[j]) : -INFINITY;
                    r[j] = __float_as_uint(v);
                    cmax = fmaxf(cmax, v);
                }
            }
            float* slot = s_xch + (i & 1) * 256;
            slot[half * 128 + row] = cmax;
            oh_pair_sync(quarter);
            cmax = fmaxf(cmax, slot[(half ^ 1) * 128 + row]);
            const float mx = fmaxf(m_run, cmax);
            const float sub = mx == -INFINITY ? 0.f : mx;                  // a row without any key so far (row 0): all p = 0
            const float alpha = oh_ex2(m_run - sub);                       // 0 while m_run = -inf
            m_run = mx;
            float sum = 0.f;
            uint32_t lo[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float p0 = oh_ex2(__uint_as_float(r[2 * j]) - sub), p1 = oh_ex2(__uint_as_float(r[2 * j + 1]) - sub);
                sum += p0 + p1;
                oh_split2(p0, p1, r[j], lo[j]);                            // r[0..16) becomes P_hi (pairs of keys, even key low)
            }
            l_run = fmaf(l_run, alpha, sum);
            tc_st16(t_sp + (uint32_t)(half * 16), r);
            tc_st16(t_sp + 32u + (uint32_t)(half * 16), lo);
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[b]);
            if (i > 0) {
                // PV(i) accumulates into O: bring O to the new maximum first (PV(i-1) must have retired)
                mbar_wait(&pv_done[b ^ 1], ((i - 1) >> 1) & 1);
                tc_fence_after();
                if (__any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll 1
                    for (int q = 0; q < 5; ++q) {                          // this half's 80 output columns
                        uint32_t o16[16];
                        oh_ld16(trow + OH_T_O + (uint32_t)(half * 80 + 16 * q), o16);
                        tc_wait_ld();
#pragma unroll
                        for (int d = 0; d < 16; ++d) o16[d] = __float_as_uint(__uint_as_float(o16[d]) * alpha);
                        tc_st16(trow + OH_T_O + (uint32_t)(half * 80 + 16 * q), o16);
                    }
                    tc_wait_st();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(o_ready);
            }
        }
        mbar_wait(&pv_done[(nc - 1) & 1], ((nc - 1) >> 1) & 1);            // PV of the last chunk
        tc_fence_after();
        // join the halves: row sum (same running maximum in both threads) and the two diagonal dot products
        float* slot = s_xch + 512;
        slot[half * 128 + row] = l_run;
        slot[256 + half * 128 + row] = dk;
        slot[512 + half * 128 + row] = dku;
        oh_pair_sync(quarter);
        const float l_off = l_run + slot[(half ^ 1) * 128 + row];
        const float sii = (dk + slot[256 + (half ^ 1) * 128 + row]) * qs;     // known stream: own key with its occupancy
        const float siu = (dku + slot[512 + (half ^ 1) * 128 + row]) * qs;    // unknown stream: own key without
        // each stream: maximum over (strictly lower part, own diagonal term), weights of the two parts, normaliser
        const float M1 = fmaxf(m_run, sii), M2 = fmaxf(m_run, siu);
        const float wo1 = oh_ex2(m_run - M1), wd1 = oh_ex2(sii - M1), wo2 = oh_ex2(m_run - M2), wd2 = oh_ex2(siu - M2);
        const float inv1 = 1.0f / fmaf(l_off, wo1, wd1), inv2 = 1.0f / fmaf(l_off, wo2, wd2);
        const float a1 = wo1 * inv1, d1 = wd1 * inv1, a2 = wo2 * inv2, d2 = wd2 * inv2;
        const float* vrow = V + (base + u) * ld + h * OH_HD;
        const float* vurow = VU + (base + u) * ld + h * OH_HD;
        float* dst = O + (base + u) * ldo + h * OH_HD;
        float* dstu = OU + (base + u) * ldo + h * OH_HD;
#pragma unroll 1
        for (int q = 0; q < 5; ++q) {                                      // this half's output dims [80 half, +80) below 150
            uint32_t o16[16];
            oh_ld16(trow + OH_T_O + (uint32_t)(half * 80 + 16 * q), o16);  // .sync.aligned: the WHOLE warp, valid row or not
            tc_wait_ld();
            if (valid) {
#pragma unroll
                for (int e = 0; e < 16; e += 2) {
                    const int d = half * 80 + 16 * q + e;
                    if (d < OH_HD) {
                        const float2 vv = __ldg(reinterpret_cast<const float2*>(vrow + d)), vu = __ldg(reinterpret_cast<const float2*>(vurow + d));
                        const float o0 = __uint_as_float(o16[e]), o1 = __uint_as_float(o16[e + 1]);
                        *reinterpret_cast<float2*>(dst + d) = make_float2(fmaf(o0, a1, d1 * vv.x), fmaf(o1, a1, d1 * vv.y));
                        *reinterpret_cast<float2*>(dstu + d) = make_float2(fmaf(o0, a2, d2 * vu.x), fmaf(o1, a2, d2 * vu.y));
                    }
                }
            }
            __syncwarp();
        }
    } else if (warp == 8) {
        // ---------------- MMA issuer: all lanes run the loop, the elected lane issues (tc.cuh elect_one) ----------------
        const uint32_t idesc_s = (1u << 4) | ((uint32_t)(OH_BK >> 3) << 17) | ((uint32_t)(OH_BQ >> 4) << 24);     // f16 x f16 -> f32, N = 64
        const uint32_t idesc_o = (1u << 4) | ((uint32_t)(OH_HDP >> 3) << 17) | ((uint32_t)(OH_BQ >> 4) << 24);    // N = 160
        mbar_wait(q_full, 0);
#pragma unroll 1
        for (int i = 0; i <= nc; ++i) {
            if (i < nc) {                                                  // S(i) = Q K_i^T
                const int b = i & 1, n = i >> 1;
                mbar_wait(&k_full[b], n & 1);
                tc_fence_after();
                const uint8_t* ks_ = sm + OH_OFF_K + b * OH_KSTAGE;
                const uint32_t d_tmem = tmem + OH_T_SP + (uint32_t)(b * 64);
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < OH_HDP / 16; ++kk) {             // 10 K-steps of 16 dims = 8 TMEM columns of Q each
                        const int kb = kk >> 2, ks = kk & 3;               // 64-dim block of the K tile, step inside it
                        const uint64_t kh = make_smem_desc(ks_ + kb * OH_KBLK) + (uint64_t)(2 * ks);
                        const uint64_t kl = make_smem_desc(ks_ + OH_KTILE + kb * OH_KBLK) + (uint64_t)(2 * ks);
                        tc_mma_f16_ts(d_tmem, tmem + OH_T_QH + 8u * kk, kh, idesc_s, kk ? 1u : 0u);
                        tc_mma_f16_ts(d_tmem, tmem + OH_T_QL + 8u * kk, kh, idesc_s, 1u);
                        tc_mma_f16_ts(d_tmem, tmem + OH_T_QH + 8u * kk, kl, idesc_s, 1u);
                    }
                    tc_commit(&s_full[b]);
                }
                __syncwarp();
            }
            if (i >= 1) {                                                  // O (+)= P_j V_j
                const int j = i - 1, b = j & 1, n = j >> 1;
                mbar_wait(&v_full[b], n & 1);
                mbar_wait(&p_full[b], n & 1);
                if (j > 0) mbar_wait(o_ready, (j - 1) & 1);
                tc_fence_after();
                const uint8_t* vs_ = sm + OH_OFF_V + b * OH_VSTAGE;
                const uint32_t p_tmem = tmem + OH_T_SP + (uint32_t)(b * 64);
                const uint32_t d_tmem = tmem + OH_T_O;
                const uint64_t vh = make_smem_desc(vs_), vl = make_smem_desc(vs_ + OH_VTILE);
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {                       // 16 keys = 8 TMEM columns of P = 2 descriptor units of V^T
                        const uint64_t adv = (uint64_t)(2 * ks);
                        tc_mma_f16_ts(d_tmem, p_tmem + 8u * ks, vh + adv, idesc_o, (j | ks) ? 1u : 0u);
                        tc_mma_f16_ts(d_tmem, p_tmem + 32u + 8u * ks, vh + adv, idesc_o, 1u);
                        tc_mma_f16_ts(d_tmem, p_tmem + 8u * ks, vl + adv, idesc_o, 1u);
                    }
                    tc_commit(&pv_done[b]);
                }
                __syncwarp();
            }
        }
    } else {
        // ---------------- TMA producer (warp 9): chunk i = padded rows [64 (tile0 + i), +64) ----------------
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmKh)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmKl)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmVh)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmVl)) : "memory");
        }
#pragma unroll 1
        for (int i = 0; i < nc; ++i) {
            const int b = i & 1, n = i >> 1;
            const int prow = (tile0 + i) * OH_BK;
            uint8_t* kdst = sm + OH_OFF_K + b * OH_KSTAGE;
            uint8_t* vdst = sm + OH_OFF_V + b * OH_VSTAGE;
            if (n > 0) mbar_wait(&s_full[b], (n - 1) & 1);               // S(i-2) retired: K stage b is free
            if (elect_one()) {
                mbar_expect_tx(&k_full[b], OH_KSTAGE);
#pragma unroll
                for (int kb = 0; kb < 3; ++kb) {
                    tma_load_2d(kdst + kb * OH_KBLK, &tmKh, &k_full[b], h * OH_QP + kb * 64, prow);
                    tma_load_2d(kdst + OH_KTILE + kb * OH_KBLK, &tmKl, &k_full[b], h * OH_QP + kb * 64, prow);
                }
            }
            __syncwarp();
            if (n > 0) mbar_wait(&pv_done[b], (n - 1) & 1);              // PV(i-2) retired: V stage b is free
            if (elect_one()) {
                mbar_expect_tx(&v_full[b], OH_VSTAGE);
                tma_load_2d(vdst, &tmVh, &v_full[b], prow, h * OH_HDP);
                tma_load_2d(vdst + OH_VTILE, &tmVl, &v_full[b], prow, h * OH_HDP);
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(OH_TMEM_COLS) : "memory");
    }
}

bool octattn_h_ok(long long ld, long long ldo, int head_dim, const void* qu, const void* k, const void* ku, const void* v,
                  const void* vu, const void* o, const void* ou) {
    if (head_dim != OH_HD || (ld & 1) || (ldo & 1)) return false;
    const void* ps[] = {qu, k, ku, v, vu, o, ou};
    for (const void* p : ps)
        if (reinterpret_cast<uintptr_t>(p) & 7) return false;              // float2 accesses
    return true;
}

// h_off [n_seq + 1]: host copy of the sequence offsets; the 64- / 128-token tile tables are scp_seqs'
int octattn_attn_h(const float* qu, const float* k, const float* ku, const float* v, const float* vu, long long ld, int heads,
                   const long long* h_off, int n_seq, const long long* d_off, const int* d_tile_seq, const int* d_tile_start,
                   int n_tile, const int* d_tile128_seq, const int* d_tile128_start, int n_tile128, float* out, float* out_u,
                   long long ldo, cudaStream_t st) {
    static bool attr = false;
    if (!attr) { SCP_CUDA(cudaFuncSetAttribute(k_octattn_attn_h, cudaFuncAttributeMaxDynamicSharedMemorySize, OH_SMEM)); attr = true; }
    std::vector<int> first(n_seq);
    int acc = 0;
    for (int s = 0; s < n_seq; ++s) { first[s] = acc; acc += (int)((h_off[s + 1] - h_off[s] + 63) / 64); }
    if (acc != n_tile) { set_error("octattn_attn_h: tile table mismatch (%d vs %d)", acc, n_tile); return SCP_ERR_INTERNAL; }
    int* d_first = nullptr;
    SCP_CUDA(upload_async((void**)&d_first, first.data(), (size_t)n_seq * 4, st));
    const long long rows = (long long)n_tile * 64, kcols = (long long)heads * OH_QP, vrows = (long long)heads * OH_HDP;
    __half* buf = nullptr;
    SCP_CUDA(malloc_async((void**)&buf, (size_t)(2 * rows * kcols + 2 * vrows * rows) * sizeof(__half) + 1024, st));
    __half *k_hi = buf, *k_lo = buf + rows * kcols, *vt_hi = buf + 2 * rows * kcols, *vt_lo = vt_hi + vrows * rows;
    k_octattn_prep<<<dim3((unsigned)n_tile, (unsigned)heads), 256, 0, st>>>(k, v, ld, heads, d_off, d_tile_seq, d_tile_start, n_tile,
                                                                           k_hi, k_lo, vt_hi, vt_lo);
    SCP_LAUNCHED();
    CUtensorMap mkh, mkl, mvh, mvl;
    if (int e = get_tensor_map_2d_f16(k_hi, kcols, rows, (int)kcols, 64, &mkh)) return e;
    if (int e = get_tensor_map_2d_f16(k_lo, kcols, rows, (int)kcols, 64, &mkl)) return e;
    if (int e = get_tensor_map_2d_f16(vt_hi, rows, vrows, (int)rows, OH_HDP, &mvh)) return e;
    if (int e = get_tensor_map_2d_f16(vt_lo, rows, vrows, (int)rows, OH_HDP, &mvl)) return e;
    dim3 grid((unsigned)n_tile128, (unsigned)heads);
    k_octattn_attn_h<<<grid, OH_THREADS, OH_SMEM, st>>>(qu, k, ku, v, vu, ld, mkh, mkl, mvh, mvl, heads, d_off, d_tile128_seq,
                                                        d_tile128_start, d_first, out, out_u, ldo);
    SCP_LAUNCHED();
    SCP_CUDA(cudaFreeAsync(buf, st));
    SCP_CUDA(cudaFreeAsync(d_first, st));
    return SCP_OK;
}

}  // namespace scp

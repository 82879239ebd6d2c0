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
// Warp roles (320 threads): warps 0-7 softmax / output (a warp owns 16 query rows = TMEM lanes 32 (warp & 3) + 16 (warp >> 2) + [0,16),
// four threads per row), warp 8 MMA issuer + TMEM owner, warp 9 TMA producer.
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
// 16 lanes x 64 columns: register 4 n + {0,1} = (lane l / 4, column 8 n + 2 (l % 4) + {0,1}), 4 n + {2,3} = the same of lane l / 4 + 8
__device__ __forceinline__ void oh_ld_16x256b_x8(uint32_t taddr, uint32_t* r) {
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
// the same for 32 columns
__device__ __forceinline__ void oh_ld_16x256b_x4(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void oh_st_16x256b_x4(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
// 16 lanes x 32 columns: register 2 n + k = (lane l / 4 + 8 k, column 4 n + l % 4)
__device__ __forceinline__ void oh_st_16x128b_x8(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.16x128b.x8.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
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
        // ---------------- softmax + output: a warp owns 16 query rows, FOUR threads per row (see attn_h.cu) ----------------
        // rows = TMEM lanes 32 (warp & 3) + 16 (warp >> 2) + [0, 16); thread (g = lane / 4, q = lane % 4) holds, of rows g and g + 8,
        // the columns 8 n + 2 q + {0, 1} of every 64-column piece (tcgen05.ld.16x256b): row maxima, row sums and the two diagonal
        // dot products are quad reductions (two shuffles) instead of a shared-memory exchange between two warps, and a register
        // pair is one packed fp16 word of P / Q that tcgen05.st.16x128b puts back in the operand layout of the TS-form MMA.
        const int quarter = warp & 3, rh = warp >> 2;
        const int g = lane >> 2, q4 = lane & 3;
        const int rowA = quarter * 32 + rh * 16 + g;                       // rowB = rowA + 8
        const int uA = q0 + rowA, uB = uA + 8;                             // query indices inside the sequence
        const bool vA = uA < S, vB = uB < S;
        const uint32_t tbase = tmem + ((uint32_t)(quarter * 32 + rh * 16) << 16);
        const float qs = rsqrtf((float)OH_HD) * OH_LOG2E;                  // 1/sqrt(150) and log2(e): scores live in the log2 domain
        float dkA = 0.f, dkuA = 0.f, dkB = 0.f, dkuB = 0.f;                // this thread's part of QU . K_own and QU . KU_own
        {   // two Q rows (192 padded dims, zero behind dim 150) -> TMEM as fp16 hi/lo pairs, 32 words (64 dims) at a time
            const float* qA = QU + (base + uA) * ld + h * OH_HD;
            const float* kA = K + (base + uA) * ld + h * OH_HD;
            const float* kuA = KU + (base + uA) * ld + h * OH_HD;
            const float* qB = QU + (base + uB) * ld + h * OH_HD;
            const float* kB = K + (base + uB) * ld + h * OH_HD;
            const float* kuB = KU + (base + uB) * ld + h * OH_HD;
#pragma unroll 1
            for (int p = 0; p < 3; ++p) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                    const int d = 64 * p + 8 * n + 2 * q4;                 // word 32 p + 4 n + q: dims d, d + 1
                    float2 a = make_float2(0.f, 0.f), c = a;
                    if (d < OH_HD) {
                        if (vA) {
                            a = __ldg(reinterpret_cast<const float2*>(qA + d));
                            const float2 x = __ldg(reinterpret_cast<const float2*>(kA + d)), y = __ldg(reinterpret_cast<const float2*>(kuA + d));
                            dkA = fmaf(a.x, x.x, fmaf(a.y, x.y, dkA));
                            dkuA = fmaf(a.x, y.x, fmaf(a.y, y.y, dkuA));
                        }
                        if (vB) {
                            c = __ldg(reinterpret_cast<const float2*>(qB + d));
                            const float2 x = __ldg(reinterpret_cast<const float2*>(kB + d)), y = __ldg(reinterpret_cast<const float2*>(kuB + d));
                            dkB = fmaf(c.x, x.x, fmaf(c.y, x.y, dkB));
                            dkuB = fmaf(c.x, y.x, fmaf(c.y, y.y, dkuB));
                        }
                    }
                    oh_split2(a.x * qs, a.y * qs, hi[2 * n], lo[2 * n]);
                    oh_split2(c.x * qs, c.y * qs, hi[2 * n + 1], lo[2 * n + 1]);
                }
                oh_st_16x128b_x8(tbase + OH_T_QH + (uint32_t)(32 * p), hi);
                oh_st_16x128b_x8(tbase + OH_T_QL + (uint32_t)(32 * p), lo);
            }
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(q_full);
        }
        float mA = -INFINITY, mB = -INFINITY, lA = 0.f, lB = 0.f;         // running maxima; PARTIAL row sums of this thread's columns
#pragma unroll 1
        for (int i = 0; i < nc; ++i) {
            const int b = i & 1, n_ = i >> 1;
            const uint32_t t_sp = tbase + OH_T_SP + (uint32_t)(b * 64);
            mbar_wait(&s_full[b], n_ & 1);
            tc_fence_after();
            uint32_t r[32];
            oh_ld_16x256b_x8(t_sp, r);
            tc_wait_ld();
            float cA = -INFINITY, cB = -INFINITY;
            if (i * OH_BK + OH_BK <= q0) {                                 // chunk entirely below the block's first row: no mask
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                    cA = fmaxf(cA, fmaxf(__uint_as_float(r[4 * n]), __uint_as_float(r[4 * n + 1])));
                    cB = fmaxf(cB, fmaxf(__uint_as_float(r[4 * n + 2]), __uint_as_float(r[4 * n + 3])));
                }
            } else {                                                       // strictly-lower triangle: key < query
                const int jb = i * OH_BK + 2 * q4;
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                    const int j = jb + 8 * n;
                    const float a0 = j < uA ? __uint_as_float(r[4 * n]) : -INFINITY, a1 = j + 1 < uA ? __uint_as_float(r[4 * n + 1]) : -INFINITY;
                    const float b0 = j < uB ? __uint_as_float(r[4 * n + 2]) : -INFINITY, b1 = j + 1 < uB ? __uint_as_float(r[4 * n + 3]) : -INFINITY;
                    r[4 * n] = __float_as_uint(a0); r[4 * n + 1] = __float_as_uint(a1);
                    r[4 * n + 2] = __float_as_uint(b0); r[4 * n + 3] = __float_as_uint(b1);
                    cA = fmaxf(cA, fmaxf(a0, a1)); cB = fmaxf(cB, fmaxf(b0, b1));
                }
            }
            cA = fmaxf(cA, __shfl_xor_sync(0xffffffffu, cA, 1)); cB = fmaxf(cB, __shfl_xor_sync(0xffffffffu, cB, 1));
            cA = fmaxf(cA, __shfl_xor_sync(0xffffffffu, cA, 2)); cB = fmaxf(cB, __shfl_xor_sync(0xffffffffu, cB, 2));
            const float mxA = fmaxf(mA, cA), mxB = fmaxf(mB, cB);
            const float subA = mxA == -INFINITY ? 0.f : mxA, subB = mxB == -INFINITY ? 0.f : mxB;    // a row without any key so far: all p = 0
            const float alA = oh_ex2(mA - subA), alB = oh_ex2(mB - subB);  // 0 while m = -inf
            mA = mxA; mB = mxB;
            float sA = 0.f, sB = 0.f;
            uint32_t ph[16], pl[16];
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const float a0 = oh_ex2(__uint_as_float(r[4 * n]) - subA), a1 = oh_ex2(__uint_as_float(r[4 * n + 1]) - subA);
                const float b0 = oh_ex2(__uint_as_float(r[4 * n + 2]) - subB), b1 = oh_ex2(__uint_as_float(r[4 * n + 3]) - subB);
                sA += a0 + a1; sB += b0 + b1;
                oh_split2(a0, a1, ph[2 * n], pl[2 * n]);                   // word 4 n + q of row A: keys 8 n + 2 q, + 1 (even key low)
                oh_split2(b0, b1, ph[2 * n + 1], pl[2 * n + 1]);
            }
            lA = fmaf(lA, alA, sA); lB = fmaf(lB, alB, sB);
            oh_st_16x128b_x8(t_sp, ph);
            oh_st_16x128b_x8(t_sp + 32u, pl);
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[b]);
            if (i > 0) {
                // PV(i) accumulates into O: bring O to the new maximum first (PV(i-1) must have retired)
                mbar_wait(&pv_done[b ^ 1], ((i - 1) >> 1) & 1);
                tc_fence_after();
                if (__any_sync(0xffffffffu, alA != 1.0f || alB != 1.0f)) {
#pragma unroll 1
                    for (int p = 0; p < 5; ++p) {                          // 160 output columns, 32 at a time
                        uint32_t o[16];
                        oh_ld_16x256b_x4(tbase + OH_T_O + (uint32_t)(32 * p), o);
                        tc_wait_ld();
#pragma unroll
                        for (int n = 0; n < 4; ++n) {
                            o[4 * n] = __float_as_uint(__uint_as_float(o[4 * n]) * alA); o[4 * n + 1] = __float_as_uint(__uint_as_float(o[4 * n + 1]) * alA);
                            o[4 * n + 2] = __float_as_uint(__uint_as_float(o[4 * n + 2]) * alB); o[4 * n + 3] = __float_as_uint(__uint_as_float(o[4 * n + 3]) * alB);
                        }
                        oh_st_16x256b_x4(tbase + OH_T_O + (uint32_t)(32 * p), o);
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
        // join the four threads of a row: row sum (same running maximum in all of them) and the two diagonal dot products
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {
            lA += __shfl_xor_sync(0xffffffffu, lA, o); lB += __shfl_xor_sync(0xffffffffu, lB, o);
            dkA += __shfl_xor_sync(0xffffffffu, dkA, o); dkB += __shfl_xor_sync(0xffffffffu, dkB, o);
            dkuA += __shfl_xor_sync(0xffffffffu, dkuA, o); dkuB += __shfl_xor_sync(0xffffffffu, dkuB, o);
        }
        // each stream: maximum over (strictly lower part, own diagonal term), weights of the two parts, normaliser
        auto weights = [&](float m_run, float l_off, float dk, float dku, float& a1, float& d1, float& a2, float& d2) {
            const float sii = dk * qs, siu = dku * qs;                     // known stream: own key with its occupancy; unknown: without
            const float M1 = fmaxf(m_run, sii), M2 = fmaxf(m_run, siu);
            const float wo1 = oh_ex2(m_run - M1), wd1 = oh_ex2(sii - M1), wo2 = oh_ex2(m_run - M2), wd2 = oh_ex2(siu - M2);
            const float inv1 = 1.0f / fmaf(l_off, wo1, wd1), inv2 = 1.0f / fmaf(l_off, wo2, wd2);
            a1 = wo1 * inv1; d1 = wd1 * inv1; a2 = wo2 * inv2; d2 = wd2 * inv2;
        };
        float a1A, d1A, a2A, d2A, a1B, d1B, a2B, d2B;
        weights(mA, lA, dkA, dkuA, a1A, d1A, a2A, d2A);
        weights(mB, lB, dkB, dkuB, a1B, d1B, a2B, d2B);
        auto emit_row = [&](int u, const uint32_t* o, int k, int p, float a1, float d1, float a2, float d2) {
            const float* vrow = V + (base + u) * ld + h * OH_HD;
            const float* vurow = VU + (base + u) * ld + h * OH_HD;
            float* dst = O + (base + u) * ldo + h * OH_HD;
            float* dstu = OU + (base + u) * ldo + h * OH_HD;
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                const int d = 32 * p + 8 * n + 2 * q4;
                if (d < OH_HD) {
                    const float2 vv = __ldg(reinterpret_cast<const float2*>(vrow + d)), vu = __ldg(reinterpret_cast<const float2*>(vurow + d));
                    const float o0 = __uint_as_float(o[4 * n + 2 * k]), o1 = __uint_as_float(o[4 * n + 2 * k + 1]);
                    *reinterpret_cast<float2*>(dst + d) = make_float2(fmaf(o0, a1, d1 * vv.x), fmaf(o1, a1, d1 * vv.y));
                    *reinterpret_cast<float2*>(dstu + d) = make_float2(fmaf(o0, a2, d2 * vu.x), fmaf(o1, a2, d2 * vu.y));
                }
            }
        };
#pragma unroll 1
        for (int p = 0; p < 5; ++p) {                                      // output dims [32 p, +32) below 150
            uint32_t o[16];
            oh_ld_16x256b_x4(tbase + OH_T_O + (uint32_t)(32 * p), o);     // .sync.aligned: the WHOLE warp, valid rows or not
            tc_wait_ld();
            if (vA) emit_row(uA, o, 0, p, a1A, d1A, a2A, d2A);
            if (vB) emit_row(uB, o, 1, p, a1B, d1B, a2B, d2B);
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

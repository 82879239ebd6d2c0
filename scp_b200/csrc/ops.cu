// Entropy-model operators, fp32 SIMT engine.
// Everything here is hand-written CUDA; the tcgen05/TMEM/TMA GEMM lives in gemm_tc.cu.
//
// Reference behaviour reproduced (paths in luoao-kddi/SCP): models/dgcnn.py:10-154,
// models/swin_transformer.py:322-367,406-501,583-706, models/ehem.py:72-136,
// models/oct_attention.py:48-99, models/attention_model.py:6-155.
#include <vector>
#include <algorithm>
#include <math.h>
#include <mutex>
#include <string.h>
#include <stdlib.h>
#include "common.cuh"

struct scp_seqs {
    int n_seq = 0;
    long long total = 0;
    std::vector<long long> h_off;
    long long* d_off = nullptr;     // [n_seq+1]
    int n_win = 0;                  // 512-token attention windows over all (padded) sequences
    int* d_win_seq = nullptr;       // [n_win]
    int* d_win_idx = nullptr;       // [n_win] window index inside its sequence
    int n_tile = 0;                 // 64-token tiles over all sequences
    int* d_tile_seq = nullptr;      // [n_tile]
    int* d_tile_start = nullptr;    // [n_tile] first token of the tile inside its sequence
    int n_tile128 = 0;              // 128-token tiles
    int* d_tile128_seq = nullptr;
    int* d_tile128_start = nullptr;
    void* d_arena = nullptr;        // one stream-ordered allocation behind all the tables above
    void* h_pinned = nullptr;       // pinned staging block (returned to the pool on destroy)
    size_t pinned_bytes = 0;
    cudaEvent_t copied = nullptr;   // H2D of the tables finished: the staging block may be rewritten
    cudaStream_t stream = nullptr;
};

namespace scp {

int linear_tf32(const float* x, long long ldx, const float* w, const float* bias, const float* res, long long ldr,
                float* y, long long ldy, long long M, int N, int K, int act, cudaStream_t st, int split);   // gemm_tc.cu
void gemm_cache_clear();
void gemm_cache_drop(const void* w);
extern int g_gemm_cluster;
bool knn_tc_ok(int d, int k);                                                                      // knn_tc.cu
int knn_tc(const float* d_x, long long ldx, int d, const long long* h_off, int n_seq, const long long* d_off,
           const int* d_tile_seq, const int* d_tile_start, int n_work, int k, int* d_idx, cudaStream_t st);
bool linear_tf32_ok(long long ldx, long long ldy, long long M, int N, int K, const void* x, const void* w, const void* y);
bool swin_attn_tc_ok(long long ldq, long long ldk, long long ldv, long long ldo, const void* q, const void* k, const void* v,
                     const void* o, const void* b0, const void* b1, const void* b2);                      // attn_tc.cu
int swin_attn_tc(const float* q, long long ldq, const float* k, long long ldk, const float* v, long long ldv, const float* qb,
                 const float* kb, const float* vb, const float* relpos, int heads, const long long* d_off, const int* d_win_seq,
                 const int* d_win_idx, int n_win, int shift, float* out, long long ldo, cudaStream_t st);
bool octattn_h_ok(long long ld, long long ldo, int head_dim, const void* qu, const void* k, const void* ku, const void* v,
                  const void* vu, const void* o, const void* ou);                                          // octattn_h.cu
int octattn_attn_h(const float* qu, const float* k, const float* ku, const float* v, const float* vu, long long ld, int heads,
                   const long long* h_off, int n_seq, const long long* d_off, const int* d_tile_seq, const int* d_tile_start,
                   int n_tile, const int* d_tile128_seq, const int* d_tile128_start, int n_tile128, float* out, float* out_u,
                   long long ldo, cudaStream_t st);
int swin_attn_h(const float* q, long long ldq, const float* k, long long ldk, const float* v, long long ldv, const float* qb,
                const float* kb, const float* vb, const float* relpos, int heads, const long long* d_off, const int* d_win_seq,
                const int* d_win_idx, int n_win, int shift, float* out, long long ldo, cudaStream_t st);          // attn_h.cu

__device__ __forceinline__ float apply_act(float v, int act) {
    switch (act) {
        case SCP_ACT_LEAKY001: return v > 0.f ? v : 0.01f * v;
        case SCP_ACT_GELU: return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
        case SCP_ACT_RELU: return v > 0.f ? v : 0.f;
        default: return v;
    }
}

// ------------------------------------------------------------------------------------------
// fp32 GEMM  Y[M,N] = act(X[M,K] W[N,K]^T + b) (+ R)
// ------------------------------------------------------------------------------------------
template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(256) k_linear_simt(const float* __restrict__ X, long long ldx,
                                                      const float* __restrict__ W, const float* __restrict__ bias,
                                                      const float* __restrict__ R, long long ldr, float* __restrict__ Y,
                                                      long long ldy, long long M, int N, int K, int act, int vec4) {
    constexpr int BK = 16;
    static_assert((BM / TM) * (BN / TN) == 256, "256 threads");
    __shared__ __align__(16) float Xs[BK][BM + 4];
    __shared__ __align__(16) float Ws[BK][BN + 4];
    const int t = threadIdx.x;
    const long long m0 = (long long)blockIdx.y * BM;
    const int n0 = blockIdx.x * BN;
    const int ty = t / (BN / TN), tx = t % (BN / TN);
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += BK) {
        if (vec4) {
            for (int e = t; e < BM * 4; e += 256) {
                int r = e >> 2, kq = (e & 3) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (m0 + r < M && k0 + kq < K) v = *reinterpret_cast<const float4*>(X + (m0 + r) * ldx + k0 + kq);
                Xs[kq][r] = v.x; Xs[kq + 1][r] = v.y; Xs[kq + 2][r] = v.z; Xs[kq + 3][r] = v.w;
            }
            for (int e = t; e < BN * 4; e += 256) {
                int r = e >> 2, kq = (e & 3) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (n0 + r < N && k0 + kq < K) v = *reinterpret_cast<const float4*>(W + (long long)(n0 + r) * K + k0 + kq);
                Ws[kq][r] = v.x; Ws[kq + 1][r] = v.y; Ws[kq + 2][r] = v.z; Ws[kq + 3][r] = v.w;
            }
        } else {
            for (int e = t; e < BM * BK; e += 256) {
                int r = e / BK, kk = e % BK;
                Xs[kk][r] = (m0 + r < M && k0 + kk < K) ? X[(m0 + r) * ldx + k0 + kk] : 0.f;
            }
            for (int e = t; e < BN * BK; e += 256) {
                int r = e / BK, kk = e % BK;
                Ws[kk][r] = (n0 + r < N && k0 + kk < K) ? W[(long long)(n0 + r) * K + k0 + kk] : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; i += 4) *reinterpret_cast<float4*>(&a[i]) = *reinterpret_cast<const float4*>(&Xs[kk][ty * TM + i]);
#pragma unroll
            for (int j = 0; j < TN; j += 4) *reinterpret_cast<float4*>(&b[j]) = *reinterpret_cast<const float4*>(&Ws[kk][tx * TN + j]);
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const long long m = m0 + ty * TM + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + tx * TN + j;
            if (n >= N) continue;
            float v = acc[i][j];
            if (bias) v += bias[n];
            v = apply_act(v, act);
            if (R) v += R[m * ldr + n];
            Y[m * ldy + n] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------
// LayerNorm (optionally of x + res): one warp per row
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_layernorm(const float* __restrict__ X, long long ldx, const float* __restrict__ R,
                                                    long long ldr, const float* __restrict__ g, const float* __restrict__ b,
                                                    float* __restrict__ Y, long long ldy, long long M, int C, float eps) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= M) return;
    constexpr int MAXP = 20;
    float v[MAXP];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXP; ++i) {
        int c = lane + 32 * i;
        float x = 0.f;
        if (c < C) { x = X[row * ldx + c]; if (R) x += R[row * ldr + c]; }
        v[i] = x;
        s += x;
    }
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXP; ++i) {
        int c = lane + 32 * i;
        float dlt = c < C ? v[i] - mean : 0.f;
        q += dlt * dlt;
    }
    const float rs = 1.0f / sqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
    for (int i = 0; i < MAXP; ++i) {
        int c = lane + 32 * i;
        if (c < C) Y[row * ldy + c] = (v[i] - mean) * rs * g[c] + b[c];
    }
}

// vectorised variant for C = 128 * V4 (256 / 512 channel rows, 16-byte aligned): one warp per row, V4 float4 per lane,
// all loads of a row issued before the first reduction
template <int V4>
__global__ void __launch_bounds__(256) k_layernorm_v4(const float* __restrict__ X, long long ldx, const float* __restrict__ R,
                                                       long long ldr, const float* __restrict__ g, const float* __restrict__ b,
                                                       float* __restrict__ Y, long long ldy, long long M, float eps) {
    constexpr int C = 128 * V4;
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= M) return;
    float4 v[V4];
    const float4* x4 = reinterpret_cast<const float4*>(X + row * ldx);
#pragma unroll
    for (int i = 0; i < V4; ++i) v[i] = __ldcs(x4 + lane + 32 * i);
    if (R) {
        const float4* r4 = reinterpret_cast<const float4*>(R + row * ldr);
#pragma unroll
        for (int i = 0; i < V4; ++i) {
            const float4 r = __ldcs(r4 + lane + 32 * i);
            v[i].x += r.x; v[i].y += r.y; v[i].z += r.z; v[i].w += r.w;
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V4; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V4; ++i) {
        v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
        q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float rs = 1.0f / sqrtf(warp_sum(q) / (float)C + eps);
    float4* y4 = reinterpret_cast<float4*>(Y + row * ldy);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    const float4* b4 = reinterpret_cast<const float4*>(b);
#pragma unroll
    for (int i = 0; i < V4; ++i) {
        const float4 gg = __ldg(g4 + lane + 32 * i), bb = __ldg(b4 + lane + 32 * i);
        float4 o;
        o.x = v[i].x * rs * gg.x + bb.x; o.y = v[i].y * rs * gg.y + bb.y;
        o.z = v[i].z * rs * gg.z + bb.z; o.w = v[i].w * rs * gg.w + bb.w;
        y4[lane + 32 * i] = o;
    }
}

// same for any C = 4 * n4 <= 128 * NV (OctAttention: C = 600, NV = 5): lanes past the end of the row hold zeros
template <int NV>
__global__ void __launch_bounds__(256) k_layernorm_v4g(const float* __restrict__ X, long long ldx, const float* __restrict__ R,
                                                        long long ldr, const float* __restrict__ g, const float* __restrict__ b,
                                                        float* __restrict__ Y, long long ldy, long long M, int C, float eps) {
    const int lane = threadIdx.x & 31, n4 = C >> 2;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= M) return;
    float4 v[NV];
    const float4* x4 = reinterpret_cast<const float4*>(X + row * ldx);
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = lane + 32 * i < n4 ? __ldcs(x4 + lane + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (R) {
        const float4* r4 = reinterpret_cast<const float4*>(R + row * ldr);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            if (lane + 32 * i < n4) {
                const float4 r = __ldcs(r4 + lane + 32 * i);
                v[i].x += r.x; v[i].y += r.y; v[i].z += r.z; v[i].w += r.w;
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        if (lane + 32 * i < n4) {
            v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
            q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
        }
    }
    const float rs = 1.0f / sqrtf(warp_sum(q) / (float)C + eps);
    float4* y4 = reinterpret_cast<float4*>(Y + row * ldy);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    const float4* b4 = reinterpret_cast<const float4*>(b);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        if (lane + 32 * i < n4) {
            const float4 gg = __ldg(g4 + lane + 32 * i), bb = __ldg(b4 + lane + 32 * i);
            float4 o;
            o.x = v[i].x * rs * gg.x + bb.x; o.y = v[i].y * rs * gg.y + bb.y;
            o.z = v[i].z * rs * gg.z + bb.z; o.w = v[i].w * rs * gg.w + bb.w;
            y4[lane + 32 * i] = o;
        }
    }
}

// ------------------------------------------------------------------------------------------
// EHEM embedding
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ehem_embed(const uint8_t* __restrict__ ctx, long long n,
                                                     const float* __restrict__ occ_enc, const float* __restrict__ level_enc,
                                                     int n_level_rows, const float* __restrict__ octant_enc,
                                                     float* __restrict__ out, long long ldo) {
    const long long total = n * 80;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const long long tok = e / 80;
        const int c = (int)(e - tok * 80);
        const uint8_t* cx = ctx + tok * 12;
        float v;
        if (c < 48) { int k = c >> 4; v = occ_enc[(int)cx[3 * k + 2] * 16 + (c & 15)]; }          // ancestors' occupancy
        else if (c < 64) { int k = (c - 48) >> 2; int l = min((int)cx[3 * k], n_level_rows - 1); v = level_enc[l * 4 + (c & 3)]; }
        else { int k = (c - 64) >> 2; int o = min((int)cx[3 * k + 1], 8); v = octant_enc[o * 4 + (c & 3)]; }
        out[tok * ldo + c] = v;
    }
}

__global__ void __launch_bounds__(256) k_ehem_embed_occ(const uint8_t* __restrict__ ctx, long long n_even,
                                                         const float* __restrict__ occ_enc, float* __restrict__ out,
                                                         long long ldo) {
    const long long total = n_even * 16;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const long long i = e >> 4;
        const int c = (int)(e & 15);
        out[i * ldo + c] = occ_enc[(int)ctx[(2 * i) * 12 + 11] * 16 + c];       // data[:, ::2, -1, -1] (ehem.py:104)
    }
}

// ------------------------------------------------------------------------------------------
// kNN: scores 2 xi.xj - |xj|^2 - |xi|^2 by fp32 tiles + per-thread sorted top-k lists in shared memory
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_row_sqnorm(const float* __restrict__ X, long long ldx, int d, long long n,
                                                     float* __restrict__ xx) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= n) return;
    float s = 0.f;
    for (int c = lane; c < d; c += 32) { float v = X[row * ldx + c]; s = fmaf(v, v, s); }
    s = warp_sum(s);
    if (lane == 0) xx[row] = s;
}

constexpr int KQ = 64, KC = 128, KD = 32;

__global__ void __launch_bounds__(256) k_knn(const float* __restrict__ X, long long ldx, int d,
                                              const float* __restrict__ xx, const long long* __restrict__ seq_off,
                                              const int* __restrict__ tile_seq, const int* __restrict__ tile_start, int k,
                                              int* __restrict__ idx_out) {
    extern __shared__ __align__(16) float sm[];
    float (*Qs)[KQ + 4] = reinterpret_cast<float (*)[KQ + 4]>(sm);                        // [KD][68]
    float (*Cs)[KC + 4] = reinterpret_cast<float (*)[KC + 4]>(sm + KD * (KQ + 4));         // [KD][132]
    float (*S)[KC + 1] = reinterpret_cast<float (*)[KC + 1]>(sm + KD * (KQ + 4) + KD * (KC + 4));   // [KQ][129]
    float* ls = sm + KD * (KQ + 4) + KD * (KC + 4) + KQ * (KC + 1);                        // [k][256]
    int* li = reinterpret_cast<int*>(ls + k * 256);                                        // [k][256]
    const int t = threadIdx.x;
    const int s = tile_seq[blockIdx.x];
    const long long base = seq_off[s];
    const int n = (int)(seq_off[s + 1] - base);
    const int q0 = tile_start[blockIdx.x];
    const int tq = t >> 4, tc = t & 15;
    for (int j = 0; j < k; ++j) { ls[j * 256 + t] = -INFINITY; li[j * 256 + t] = -1; }
    float thresh = -INFINITY;
    const int myq = t >> 2, sub = t & 3;
    for (int c0 = 0; c0 < n; c0 += KC) {
        float acc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        for (int d0 = 0; d0 < d; d0 += KD) {
            __syncthreads();
            for (int e = t; e < KQ * KD; e += 256) {
                int r = e >> 5, dd = e & 31;
                Qs[dd][r] = (q0 + r < n && d0 + dd < d) ? X[(base + q0 + r) * ldx + d0 + dd] : 0.f;
            }
            for (int e = t; e < KC * KD; e += 256) {
                int r = e >> 5, dd = e & 31;
                Cs[dd][r] = (c0 + r < n && d0 + dd < d) ? X[(base + c0 + r) * ldx + d0 + dd] : 0.f;
            }
            __syncthreads();
            const int dmax = min(KD, d - d0);
            for (int dd = 0; dd < dmax; ++dd) {
                float a[4], b[8];
                *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&Qs[dd][tq * 4]);
                *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&Cs[dd][tc * 8]);
                *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(&Cs[dd][tc * 8 + 4]);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int q = q0 + tq * 4 + i;
            const float xq = q < n ? xx[base + q] : 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = c0 + tc * 8 + j;
                float sc = -INFINITY;
                if (q < n && c < n) sc = __fsub_rn(__fsub_rn(2.0f * acc[i][j], xx[base + c]), xq);   // dgcnn.py:18-20
                S[tq * 4 + i][tc * 8 + j] = sc;
            }
        }
        __syncthreads();
        // every thread scans a quarter of its query's row and keeps a private sorted top-k
        for (int c = sub; c < KC; c += 4) {
            const float sc = S[myq][c];
            if (sc > thresh) {
                int j = k - 1;
                while (j > 0 && ls[(j - 1) * 256 + t] < sc) {
                    ls[j * 256 + t] = ls[(j - 1) * 256 + t];
                    li[j * 256 + t] = li[(j - 1) * 256 + t];
                    --j;
                }
                ls[j * 256 + t] = sc;
                li[j * 256 + t] = c0 + c;
                thresh = ls[(k - 1) * 256 + t];
            }
        }
    }
    __syncthreads();
    if (sub == 0 && q0 + myq < n) {
        // 4-way merge of the sorted partial lists (score desc, index asc)
        int p[4] = {0, 0, 0, 0};
        const long long row = base + q0 + myq;
        for (int j = 0; j < k; ++j) {
            int best = -1;
            float bs = -INFINITY;
            int bi = 0x7fffffff;
            for (int u = 0; u < 4; ++u) {
                if (p[u] >= k) continue;
                const float sc = ls[p[u] * 256 + t + u];
                const int id = li[p[u] * 256 + t + u];
                if (id < 0) continue;
                if (sc > bs || (sc == bs && id < bi)) { bs = sc; bi = id; best = u; }
            }
            if (best < 0) { idx_out[row * k + j] = (int)row; continue; }     // sequence shorter than k: repeat self
            ++p[best];
            idx_out[row * k + j] = (int)(base + bi);
        }
    }
}

// d <= 4 (the 3-D position kNN): octree positions lie on a grid, so EXACT distance ties are the rule, and the
// reference's pick among tied neighbours is whatever torch.topk returns.  Here the rule is canonical: exact
// float64 squared distance ((dx^2 + dy^2) + dz^2, no FMA contraction), ties -> lowest index.  One thread per query.
constexpr int KS_Q = 128, KS_C = 256;
__host__ __device__ __forceinline__ int knn_tile_order(int i, int t0, int nt) {     // same rule as knn_tc.cu
    const int a = t0 > 0 ? t0 - 1 : 0, b = t0 + 1 < nt ? t0 + 1 : nt - 1;
    const int first[3] = {t0, t0 - 1, t0 + 1};
    int nf = 0;
#pragma unroll
    for (int u = 0; u < 3; ++u) {
        if (first[u] >= 0 && first[u] < nt) { if (i == nf) return first[u]; ++nf; }
    }
    const int j = i - nf;
    return j < a ? j : j + (b - a + 1);
}
// d <= 4 (positions): the neighbour list is DEFINED by the exact float64 squared distance ((dx^2 + dy^2) + dz^2 + dw^2,
// one rounding per operation, no FMA), ties -> lowest index.  Evaluating that for all n^2 pairs is bound by the FP64
// pipe (8 FP64 instructions per pair).  So every pair is first screened in float32 (the float32 value is within 4e-7
// relative of the exact one) against the row's current k-th distance inflated by 4e-6; only survivors -- a few hundred
// of 8192 per row -- take the exact path.  The screen can only pass extra candidates, never drop one.
// Bounding box of every 128-token tile (min x,y,z,w | max x,y,z,w): tokens are in Morton order, so a tile is a compact
// region and whole candidate chunks can be rejected against a query tile by their box distance.
__global__ void __launch_bounds__(256) k_tile_aabb(const float* __restrict__ X, long long ldx, int d,
                                                    const long long* __restrict__ seq_off, const int* __restrict__ tile_seq,
                                                    const int* __restrict__ tile_start, int n_tile, float* __restrict__ aabb) {
    const int tile = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (tile >= n_tile) return;
    const int s = tile_seq[tile];
    const long long base = seq_off[s];
    const int n = (int)(seq_off[s + 1] - base), t0 = tile_start[tile];
    float mn[4] = {INFINITY, INFINITY, INFINITY, INFINITY}, mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    for (int r = t0 + lane; r < min(n, t0 + 128); r += 32)
        for (int c = 0; c < 4; ++c) { const float v = c < d ? X[(base + r) * ldx + c] : 0.f; mn[c] = fminf(mn[c], v); mx[c] = fmaxf(mx[c], v); }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o)); mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o)); }
    }
    if (lane < 4) { aabb[tile * 8 + lane] = mn[lane]; aabb[tile * 8 + 4 + lane] = mx[lane]; }
}

__global__ void __launch_bounds__(KS_Q) k_knn_small(const float* __restrict__ X, long long ldx, int d,
                                                     const long long* __restrict__ seq_off, const int* __restrict__ tile_seq,
                                                     const int* __restrict__ tile_start, int k, int* __restrict__ idx_out,
                                                     const float* __restrict__ aabb) {
    extern __shared__ __align__(16) double smd[];
    float4* cs = reinterpret_cast<float4*>(smd);        // [KS_C] candidate coordinates (unused dims = 0)
    double* ls = smd + KS_C * 2;                        // [k][KS_Q]
    int* li = reinterpret_cast<int*>(ls + k * KS_Q);    // [k][KS_Q]
    const int t = threadIdx.x;
    const int s = tile_seq[blockIdx.x];
    const long long base = seq_off[s];
    const int n = (int)(seq_off[s + 1] - base);
    const int q = tile_start[blockIdx.x] + t;
    const bool active = q < n;
    float xf[4] = {0.f, 0.f, 0.f, 0.f};
    if (active) for (int c = 0; c < d; ++c) xf[c] = X[(base + q) * ldx + c];
    const double xq[4] = {(double)xf[0], (double)xf[1], (double)xf[2], (double)xf[3]};
    for (int j = 0; j < k; ++j) { ls[j * KS_Q + t] = INFINITY; li[j * KS_Q + t] = -1; }
    double worst = INFINITY;
    float wf = INFINITY;                                // float32 screen: >= worst * (1 + 4e-6)
    const int nchunk = (n + KS_C - 1) / KS_C;
    const int tile0 = (int)blockIdx.x - tile_start[blockIdx.x] / KS_Q;    // first 128-token tile of this sequence
    const float4 qmn = *reinterpret_cast<const float4*>(aabb + (size_t)blockIdx.x * 8);
    const float4 qmx = *reinterpret_cast<const float4*>(aabb + (size_t)blockIdx.x * 8 + 4);
    // the same test per warp (32 consecutive queries: a tighter box) decides whether the warp scans a chunk that was loaded
    float wmn[4], wmx[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        wmn[c] = active ? xf[c] : INFINITY; wmx[c] = active ? xf[c] : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { wmn[c] = fminf(wmn[c], __shfl_xor_sync(0xffffffffu, wmn[c], o)); wmx[c] = fmaxf(wmx[c], __shfl_xor_sync(0xffffffffu, wmx[c], o)); }
    }
    for (int ci = 0; ci < nchunk; ++ci) {
        // own neighbourhood first, then ascending (see knn_tile_order): ties are still resolved by (distance, index)
        const int c0 = knn_tile_order(ci, tile_start[blockIdx.x] / KS_C, nchunk) * KS_C;
        // box-to-box distance between this block's queries and the chunk (two 128-token tiles): if it exceeds every row's
        // screen threshold, no candidate of the chunk can pass the float32 screen below -- skip it unread
        float dmin2, wmin2;
        {
            const int ta = tile0 + c0 / KS_Q;
            float4 cmn = *reinterpret_cast<const float4*>(aabb + (size_t)ta * 8), cmx = *reinterpret_cast<const float4*>(aabb + (size_t)ta * 8 + 4);
            if (c0 + KS_Q < n) {
                const float4 m2 = *reinterpret_cast<const float4*>(aabb + (size_t)(ta + 1) * 8), x2 = *reinterpret_cast<const float4*>(aabb + (size_t)(ta + 1) * 8 + 4);
                cmn = make_float4(fminf(cmn.x, m2.x), fminf(cmn.y, m2.y), fminf(cmn.z, m2.z), fminf(cmn.w, m2.w));
                cmx = make_float4(fmaxf(cmx.x, x2.x), fmaxf(cmx.y, x2.y), fmaxf(cmx.z, x2.z), fmaxf(cmx.w, x2.w));
            }
            const float gx = fmaxf(0.f, fmaxf(qmn.x - cmx.x, cmn.x - qmx.x)), gy = fmaxf(0.f, fmaxf(qmn.y - cmx.y, cmn.y - qmx.y));
            const float gz = fmaxf(0.f, fmaxf(qmn.z - cmx.z, cmn.z - qmx.z)), gw = fmaxf(0.f, fmaxf(qmn.w - cmx.w, cmn.w - qmx.w));
            dmin2 = (gx * gx + gy * gy + gz * gz + gw * gw) * (1.0f - 2e-6f);
            const float hx = fmaxf(0.f, fmaxf(wmn[0] - cmx.x, cmn.x - wmx[0])), hy = fmaxf(0.f, fmaxf(wmn[1] - cmx.y, cmn.y - wmx[1]));
            const float hz = fmaxf(0.f, fmaxf(wmn[2] - cmx.z, cmn.z - wmx[2])), hw = fmaxf(0.f, fmaxf(wmn[3] - cmx.w, cmn.w - wmx[3]));
            wmin2 = (hx * hx + hy * hy + hz * hz + hw * hw) * (1.0f - 2e-6f);
        }
        if (__syncthreads_and(!active || dmin2 > wf)) continue;          // (also the barrier in front of rewriting `cs`)
        for (int r = t; r < KS_C; r += KS_Q) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c0 + r < n) {
                const float* px = X + (base + c0 + r) * ldx;
                v.x = px[0];
                if (d > 1) v.y = px[1];
                if (d > 2) v.z = px[2];
                if (d > 3) v.w = px[3];
            }
            cs[r] = v;
        }
        __syncthreads();
        if (!active) continue;
        if (__all_sync(__activemask(), wmin2 > wf)) continue;            // no row of this warp can take a candidate of the chunk
        const int cmax = min(KS_C, n - c0);
#pragma unroll 4
        for (int r = 0; r < cmax; ++r) {
            const float4 c = cs[r];
            const float fx = xf[0] - c.x, fy = xf[1] - c.y, fz = xf[2] - c.z, fw = xf[3] - c.w;
            const float d32 = fmaf(fw, fw, fmaf(fz, fz, fmaf(fy, fy, fx * fx)));
            if (!(d32 <= wf)) continue;
            double dist = 0.0;
            {
                double df = __dsub_rn(xq[0], (double)c.x); dist = __dadd_rn(dist, __dmul_rn(df, df));
                df = __dsub_rn(xq[1], (double)c.y); dist = __dadd_rn(dist, __dmul_rn(df, df));
                df = __dsub_rn(xq[2], (double)c.z); dist = __dadd_rn(dist, __dmul_rn(df, df));
                df = __dsub_rn(xq[3], (double)c.w); dist = __dadd_rn(dist, __dmul_rn(df, df));
            }
            const int cidx = c0 + r;
            if (dist < worst || (dist == worst && cidx < li[(k - 1) * KS_Q + t])) {
                int j = k - 1;
                while (j > 0 && (ls[(j - 1) * KS_Q + t] > dist || (ls[(j - 1) * KS_Q + t] == dist && li[(j - 1) * KS_Q + t] > cidx))) {
                    ls[j * KS_Q + t] = ls[(j - 1) * KS_Q + t];
                    li[j * KS_Q + t] = li[(j - 1) * KS_Q + t];
                    --j;
                }
                ls[j * KS_Q + t] = dist;
                li[j * KS_Q + t] = cidx;
                worst = ls[(k - 1) * KS_Q + t];
                wf = worst == INFINITY ? INFINITY : __double2float_ru(worst * (1.0 + 4e-6));
            }
        }
    }
    if (active) {
        const long long row = base + q;
        for (int j = 0; j < k; ++j) {
            const int id = li[j * KS_Q + t];
            idx_out[row * k + j] = id < 0 ? (int)row : (int)(base + id);
        }
    }
}

// ------------------------------------------------------------------------------------------
// edge conv gather: out = lrelu0.2( s * (sel_k uv[nbr] + uv_self[C:]) + t )
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_edge_gather(const float* __restrict__ uv, long long lduv, int C,
                                                      const int* __restrict__ idx, int k, long long n,
                                                      const float* __restrict__ bs, const float* __restrict__ bt,
                                                      float* __restrict__ out, long long ldo, float* __restrict__ out2,
                                                      long long ldo2) {
    const int per = 256 / C;            // points per block iteration (C in {64,128,256})
    const int c = threadIdx.x % C;
    const int pl = threadIdx.x / C;
    const float sc = bs[c], sh = bt[c];
    for (long long p = (long long)blockIdx.x * per + pl; p < n; p += (long long)gridDim.x * per) {
        float mx = -INFINITY, mn = INFINITY;
        for (int j = 0; j < k; ++j) {
            const int nb = idx[p * k + j];
            const float u = uv[(long long)nb * lduv + c];
            mx = fmaxf(mx, u); mn = fminf(mn, u);
        }
        const float sel = sc >= 0.f ? mx : mn;
        float v = fmaf(sc, sel + uv[p * lduv + C + c], sh);
        v = v > 0.f ? v : 0.2f * v;
        out[p * ldo + c] = v;
        if (out2) out2[p * ldo2 + c] = v;               // the same columns in a second concatenation (dgcnn.py feeds x_k to two cats)
    }
}

// ------------------------------------------------------------------------------------------
// shifted-window attention, one block per (window, head, 64-query chunk), online softmax over 8 key tiles
// ------------------------------------------------------------------------------------------
constexpr int WS = 512, HD = 64, AQ = 64, AK = 64, ALD = 68;

__global__ void __launch_bounds__(256) k_swin_attn(const float* __restrict__ Q, long long ldq, const float* __restrict__ K,
                                                    long long ldk, const float* __restrict__ V, long long ldv,
                                                    const float* __restrict__ qb, const float* __restrict__ kb,
                                                    const float* __restrict__ vb, const float* __restrict__ relpos, int heads,
                                                    const long long* __restrict__ seq_off, const int* __restrict__ win_seq,
                                                    const int* __restrict__ win_idx, int shift, float* __restrict__ O,
                                                    long long ldo) {
    extern __shared__ __align__(16) float sm[];
    float (*Qt)[ALD] = reinterpret_cast<float (*)[ALD]>(sm);                  // [d][q]
    float (*Kt)[ALD] = reinterpret_cast<float (*)[ALD]>(sm + HD * ALD);       // [d][key]
    float (*Vs)[ALD] = reinterpret_cast<float (*)[ALD]>(sm + 2 * HD * ALD);   // [key][d]
    float (*Pt)[ALD] = reinterpret_cast<float (*)[ALD]>(sm + 3 * HD * ALD);   // [key][q]
    __shared__ float s_bias[2 * WS];                                          // relpos[:, head], index (i-j)+511
    const int t = threadIdx.x;
    const int h = blockIdx.x % heads, qc = blockIdx.x / heads;
    const int gw = blockIdx.y;
    const int s = win_seq[gw], w = win_idx[gw];
    const long long base = seq_off[s];
    const int S = (int)(seq_off[s + 1] - base);
    const int Sp = ((S + WS - 1) / WS) * WS;
    const bool last_win = (w == Sp / WS - 1) && shift > 0;
    for (int e = t; e < 2 * WS - 1; e += 256) s_bias[e] = relpos[e * heads + h];
    // Q tile (scaled by 1/sqrt(64) = 0.125, exact)
    for (int e = t; e < AQ * HD; e += 256) {
        int r = e >> 6, dd = e & 63;
        int u = (w * WS + qc * AQ + r + shift) % Sp;
        float v = u < S ? Q[(base + u) * ldq + h * HD + dd] : qb[h * HD + dd];
        Qt[dd][r] = v * 0.125f;
    }
    const int ty = t >> 4, tx = t & 15;
    float o[4][4], m[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m[i] = -INFINITY; l[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
    }
    for (int kt = 0; kt < WS / AK; ++kt) {
        __syncthreads();
        for (int e = t; e < AK * HD; e += 256) {
            int r = e >> 6, dd = e & 63;
            int u = (w * WS + kt * AK + r + shift) % Sp;
            float kv = u < S ? K[(base + u) * ldk + h * HD + dd] : kb[h * HD + dd];
            float vv = u < S ? V[(base + u) * ldv + h * HD + dd] : vb[h * HD + dd];
            Kt[dd][r] = kv;
            Vs[r][dd] = vv;
        }
        __syncthreads();
        float sc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) sc[i][j] = 0.f;
#pragma unroll 8
        for (int dd = 0; dd < HD; ++dd) {
            float a[4], b[4];
            *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&Qt[dd][ty * 4]);
            *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&Kt[dd][tx * 4]);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) sc[i][j] = fmaf(a[i], b[j], sc[i][j]);
        }
        float alpha[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int pi = qc * AQ + ty * 4 + i;
            float rmax = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int pj = kt * AK + tx * 4 + j;
                float v = sc[i][j] + s_bias[pi - pj + WS - 1];
                if (last_win && ((pi < WS / 2) != (pj < WS / 2))) v += -100.0f;     // swin_transformer.py:620
                sc[i][j] = v;
                rmax = fmaxf(rmax, v);
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, off));
            const float mnew = fmaxf(m[i], rmax);
            alpha[i] = expf(m[i] - mnew);
            float rsum = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) { sc[i][j] = expf(sc[i][j] - mnew); rsum += sc[i][j]; }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) rsum += __shfl_xor_sync(0xffffffffu, rsum, off);
            l[i] = l[i] * alpha[i] + rsum;
            m[i] = mnew;
#pragma unroll
            for (int j = 0; j < 4; ++j) Pt[tx * 4 + j][ty * 4 + i] = sc[i][j];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) o[i][j] *= alpha[i];
#pragma unroll 8
        for (int kk = 0; kk < AK; ++kk) {
            float a[4], b[4];
            *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&Pt[kk][ty * 4]);
            *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&Vs[kk][tx * 4]);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) o[i][j] = fmaf(a[i], b[j], o[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int u = (w * WS + qc * AQ + ty * 4 + i + shift) % Sp;
        if (u >= S) continue;
        const float inv = 1.0f / l[i];
        float4 r = make_float4(o[i][0] * inv, o[i][1] * inv, o[i][2] * inv, o[i][3] * inv);
        *reinterpret_cast<float4*>(O + (base + u) * ldo + h * HD + tx * 4) = r;
    }
}

// ------------------------------------------------------------------------------------------
// small data-movement kernels
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pair_concat(const float* __restrict__ X, long long ldx,
                                                      const long long* __restrict__ soff, const long long* __restrict__ doff,
                                                      const int* __restrict__ tile_seq, const int* __restrict__ tile_start,
                                                      int C, float* __restrict__ out, long long ldo) {
    const int s = tile_seq[blockIdx.x];
    const int S = (int)(soff[s + 1] - soff[s]);
    const int Sd = (int)(doff[s + 1] - doff[s]);
    const int j0 = tile_start[blockIdx.x];
    const int C4 = C >> 2;
    for (int e = threadIdx.x; e < 64 * 2 * C4; e += 256) {
        const int r = e / (2 * C4), c4 = e % (2 * C4);
        const int j = j0 + r;
        if (j >= Sd) break;
        const int src = 2 * j + (c4 >= C4 ? 1 : 0);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (src < S) v = *reinterpret_cast<const float4*>(X + (soff[s] + src) * ldx + (c4 % C4) * 4);
        *reinterpret_cast<float4*>(out + (doff[s] + j) * ldo + c4 * 4) = v;
    }
}

__global__ void __launch_bounds__(256) k_upsample_cols(const float* __restrict__ X, long long ldx,
                                                        const long long* __restrict__ soff, const long long* __restrict__ doff,
                                                        const int* __restrict__ tile_seq, const int* __restrict__ tile_start,
                                                        int shift, int C, float* __restrict__ out, long long ldo, int col_off) {
    const int s = tile_seq[blockIdx.x];
    const int Sd = (int)(doff[s + 1] - doff[s]);
    const int j0 = tile_start[blockIdx.x];
    const int C4 = C >> 2;
    for (int e = threadIdx.x; e < 64 * C4; e += 256) {
        const int r = e / C4, c4 = e % C4;
        const int j = j0 + r;
        if (j >= Sd) break;
        const float4 v = *reinterpret_cast<const float4*>(X + (soff[s] + (j >> shift)) * ldx + c4 * 4);
        *reinterpret_cast<float4*>(out + (doff[s] + j) * ldo + col_off + c4 * 4) = v;
    }
}

__global__ void __launch_bounds__(256) k_copy_cols(const float* __restrict__ X, long long ldx, long long row_step,
                                                    long long row_off, long long rows, int C, float* __restrict__ out,
                                                    long long ldo, int col_off) {
    const long long total = rows * C;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const long long r = e / C;
        const int c = (int)(e - r * C);
        out[r * ldo + col_off + c] = X[(r * row_step + row_off) * ldx + c];
    }
}

__global__ void __launch_bounds__(256) k_add(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y,
                                              long long n) {
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < n; e += (long long)gridDim.x * 256) y[e] = a[e] + b[e];
}

// ------------------------------------------------------------------------------------------
// OctAttention: embedding of both streams, and two-stream causal attention
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_octattn_embed(const uint8_t* __restrict__ ctx, const u32* __restrict__ ctx_pos,
                                                        float pos_scale, int level_base, int max_lvl,
                                                        const long long* __restrict__ seq_off, const int* __restrict__ tile_seq,
                                                        const int* __restrict__ tile_start,
                                                        const float* __restrict__ occ_enc, const float* __restrict__ level_enc,
                                                        const float* __restrict__ octant_enc, const float* __restrict__ pw,
                                                        const float* __restrict__ pb, const float* __restrict__ pe,
                                                        float* __restrict__ E, float* __restrict__ EU) {
    const int s = tile_seq[blockIdx.x];
    const long long base = seq_off[s];
    const int S = (int)(seq_off[s + 1] - base);
    const int j0 = tile_start[blockIdx.x];
    const float scale = sqrtf(600.0f);
    for (int e = threadIdx.x; e < 64 * 600; e += 256) {
        const int r = e / 600, c = e % 600;
        const int j = j0 + r;
        if (j >= S) break;
        const long long tok = base + j;
        const int k = c / 150, f = c % 150;
        const uint8_t* cx = ctx + tok * 12 + 3 * k;
        const int self_level = ctx[tok * 12 + 9];
        const int sh = max(self_level - level_base, 0);                       // oct_attention.py:57-60
        int lvl = min(max((int)cx[0] - sh, 0), max_lvl);                      // :61
        float v, vu;
        if (f < 128) { v = occ_enc[(int)cx[2] * 128 + f]; vu = (k == 3) ? occ_enc[255 * 128 + f] : v; }
        else if (f < 134) { v = vu = level_enc[lvl * 6 + (f - 128)]; }
        else if (f < 138) { v = vu = octant_enc[min((int)cx[1], 8) * 4 + (f - 134)]; }
        else {
            const int o = f - 138;
            const u32* p = ctx_pos + tok * 12 + 3 * k;
            const float x = (float)p[0] * pos_scale, y = (float)p[1] * pos_scale, z = (float)p[2] * pos_scale;
            v = vu = pw[o * 3] * x + pw[o * 3 + 1] * y + pw[o * 3 + 2] * z + pb[o];
        }
        const float pos = pe ? pe[(long long)j * 600 + c] : 0.f;       // cfg.model.pos_embed False: no PositionalEncoding module
        E[tok * 600 + c] = v * scale + pos;
        EU[tok * 600 + c] = vu * scale + pos;
    }
}

// one warp per query row; keys in tiles of 32 staged in shared memory; both streams share the scores except
// on the diagonal (attention_model.py:82-93)
constexpr int OA_Q = 16, OA_K = 32, OA_LD = 151;
__global__ void __launch_bounds__(128) k_octattn_attn(const float* __restrict__ QU, const float* __restrict__ K,
                                                       const float* __restrict__ KU, const float* __restrict__ V,
                                                       const float* __restrict__ VU, long long ld, int heads, int hd,
                                                       const long long* __restrict__ seq_off, const int* __restrict__ tile_seq,
                                                       const int* __restrict__ tile_start, float* __restrict__ O,
                                                       float* __restrict__ OU, long long ldo) {
    __shared__ float Ks[OA_K][OA_LD];
    __shared__ float Vs[OA_K][OA_LD];
    __shared__ float Qs[OA_Q][OA_LD];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int h = blockIdx.y;
    // 64-token tiles are split into 4 sub-tiles of 16 queries
    const int tile = blockIdx.x >> 2, subt = blockIdx.x & 3;
    const int s = tile_seq[tile];
    const long long base = seq_off[s];
    const int S = (int)(seq_off[s + 1] - base);
    const int q0 = tile_start[tile] + subt * OA_Q;
    if (q0 >= S) return;
    const float inv = rsqrtf((float)hd);
    for (int e = threadIdx.x; e < OA_Q * hd; e += 128) {
        int r = e / hd, dd = e % hd;
        Qs[r][dd] = q0 + r < S ? QU[(base + q0 + r) * ld + h * hd + dd] : 0.f;
    }
    float m[4], l[4], mu[4], lu[4], o[4][5], ou[4][5];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m[i] = mu[i] = -INFINITY; l[i] = lu[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 5; ++j) o[i][j] = ou[i][j] = 0.f;
    }
    const int qlast = min(q0 + OA_Q, S) - 1;
    for (int k0 = 0; k0 <= qlast; k0 += OA_K) {
        __syncthreads();
        for (int e = threadIdx.x; e < OA_K * hd; e += 128) {
            int r = e / hd, dd = e % hd;
            bool ok = k0 + r < S;
            Ks[r][dd] = ok ? K[(base + k0 + r) * ld + h * hd + dd] : 0.f;
            Vs[r][dd] = ok ? V[(base + k0 + r) * ld + h * hd + dd] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int qi = q0 + warp * 4 + i;
            if (qi >= S || k0 > qi) continue;                      // warp-uniform
            const float* qrow = Qs[warp * 4 + i];
            const int kj = k0 + lane;
            float sc = 0.f;
            for (int dd = 0; dd < hd; ++dd) sc = fmaf(qrow[dd], Ks[lane][dd], sc);
            sc *= inv;
            float scu = sc;
            const bool diag = (kj == qi);
            if (diag) {                                             // unknown stream: own key/value have no occupancy
                float z = 0.f;
                for (int dd = 0; dd < hd; ++dd) z = fmaf(qrow[dd], KU[(base + qi) * ld + h * hd + dd], z);
                scu = z * inv;
            }
            const bool ok = kj <= qi;                               // causal mask (oct_attention.py:38-46)
            if (!ok) { sc = -INFINITY; scu = -INFINITY; }
            const float tmax = warp_max(sc), tmaxu = warp_max(scu);
            const float mn = fmaxf(m[i], tmax), mnu = fmaxf(mu[i], tmaxu);
            const float a = expf(m[i] - mn), au = expf(mu[i] - mnu);
            const float p = ok ? expf(sc - mn) : 0.f, pu = ok ? expf(scu - mnu) : 0.f;
            l[i] = l[i] * a + warp_sum(p);
            lu[i] = lu[i] * au + warp_sum(pu);
            m[i] = mn; mu[i] = mnu;
#pragma unroll
            for (int j = 0; j < 5; ++j) { o[i][j] *= a; ou[i][j] *= au; }
            for (int kk = 0; kk < OA_K; ++kk) {
                const float pk = __shfl_sync(0xffffffffu, p, kk), pku = __shfl_sync(0xffffffffu, pu, kk);
                if (pk == 0.f && pku == 0.f) continue;
                const bool dg = (k0 + kk == qi);
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    const int dd = lane + 32 * j;
                    if (dd < hd) {
                        const float vv = Vs[kk][dd];
                        o[i][j] = fmaf(pk, vv, o[i][j]);
                        const float vvu = dg ? VU[(base + qi) * ld + h * hd + dd] : vv;
                        ou[i][j] = fmaf(pku, vvu, ou[i][j]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int qi = q0 + warp * 4 + i;
        if (qi >= S) continue;
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const int dd = lane + 32 * j;
            if (dd < hd) {
                O[(base + qi) * ldo + h * hd + dd] = o[i][j] / l[i];
                OU[(base + qi) * ldo + h * hd + dd] = ou[i][j] / lu[i];
            }
        }
    }
}

static int g_knn_tc = 1;         // learned-feature kNN on the tensor cores (3xTF32 Gram + fused top-k); 0 = fp32 SIMT tiles
// window attention: 0 = fp32 SIMT, 1 = tensor cores 3xTF32 (attn_tc.cu), 2 = tensor cores 3xFP16, two CTAs per SM (attn_h.cu);
// env SCP_ATTN_ENGINE overrides the default
static int attn_engine_default() { const char* e = getenv("SCP_ATTN_ENGINE"); return e ? (atoi(e) < 0 ? 0 : (atoi(e) > 2 ? 2 : atoi(e))) : 2; }
static int g_attn_tc = attn_engine_default();
// SCP_GEMM_AUTO: 0 = fp32 SIMT, 1 = 3xTF32, 2 = 3xFP16 tcgen05 engine for the large layers (env SCP_AUTO_ENGINE overrides)
static int auto_engine_default() { const char* e = getenv("SCP_AUTO_ENGINE"); return e ? (atoi(e) < 0 ? 0 : (atoi(e) > 2 ? 2 : atoi(e))) : 2; }
static int g_auto_tf32 = auto_engine_default();

static inline int grid_for(long long work, int per_block, int cap = 148 * 16) {
    return (int)std::max<long long>(1, std::min<long long>(cdiv(work, per_block), cap));
}

}  // namespace scp

using namespace scp;

extern "C" {

// Creation must not synchronise: a cudaMalloc / pageable cudaMemcpy here drains the stream ~12 times per forward pass and
// runs host and GPU in lock-step.  The tables go through a recycled pinned staging block (common.cuh) instead.
scp_seqs* scp_seqs_create_async(const int64_t* h_offsets, int n_seq, void* stream) {
    if (!h_offsets || n_seq <= 0) { set_error("scp_seqs_create: bad argument"); return nullptr; }
    cudaStream_t st = as_stream(stream);
    auto* s = new scp_seqs();
    s->n_seq = n_seq;
    s->stream = st;
    s->h_off.assign(h_offsets, h_offsets + n_seq + 1);
    s->total = h_offsets[n_seq] - h_offsets[0];
    std::vector<int> wseq, widx, tseq, tstart, t2seq, t2start;
    for (int i = 0; i < n_seq; ++i) {
        long long len = h_offsets[i + 1] - h_offsets[i];
        if (len < 0 || len > (1 << 24)) { set_error("scp_seqs_create: sequence %d has length %lld", i, len); delete s; return nullptr; }
        for (long long w = 0; w < cdiv(len, 512); ++w) { wseq.push_back(i); widx.push_back((int)w); }
        for (long long t = 0; t < len; t += 64) { tseq.push_back(i); tstart.push_back((int)t); }
        for (long long t = 0; t < len; t += 128) { t2seq.push_back(i); t2start.push_back((int)t); }
    }
    s->n_win = (int)wseq.size();
    s->n_tile = (int)tseq.size();
    s->n_tile128 = (int)t2seq.size();
    // layout of the single arena: offsets (8-byte), then the six int tables, each 16-byte aligned
    auto al = [](size_t x) { return (x + 15) & ~(size_t)15; };
    const size_t o_off = 0, o_ws = al((size_t)(n_seq + 1) * 8), o_wi = o_ws + al((size_t)s->n_win * 4),
                 o_ts = o_wi + al((size_t)s->n_win * 4), o_tt = o_ts + al((size_t)s->n_tile * 4),
                 o_2s = o_tt + al((size_t)s->n_tile * 4), o_2t = o_2s + al((size_t)s->n_tile128 * 4),
                 bytes = o_2t + al((size_t)s->n_tile128 * 4) + 16;
    PinnedBlock pb;
    if (!pinned_get(bytes, &pb)) { set_error("scp_seqs_create: pinned staging allocation failed"); delete s; return nullptr; }
    uint8_t* h = static_cast<uint8_t*>(pb.p);
    memcpy(h + o_off, s->h_off.data(), (size_t)(n_seq + 1) * 8);
    if (s->n_win) { memcpy(h + o_ws, wseq.data(), (size_t)s->n_win * 4); memcpy(h + o_wi, widx.data(), (size_t)s->n_win * 4); }
    if (s->n_tile) { memcpy(h + o_ts, tseq.data(), (size_t)s->n_tile * 4); memcpy(h + o_tt, tstart.data(), (size_t)s->n_tile * 4); }
    if (s->n_tile128) { memcpy(h + o_2s, t2seq.data(), (size_t)s->n_tile128 * 4); memcpy(h + o_2t, t2start.data(), (size_t)s->n_tile128 * 4); }
    s->h_pinned = pb.p; s->pinned_bytes = pb.bytes; s->copied = pb.ev;
    bool ok = malloc_async(&s->d_arena, bytes, st) == cudaSuccess &&
              cudaMemcpyAsync(s->d_arena, h, bytes, cudaMemcpyHostToDevice, st) == cudaSuccess &&
              cudaEventRecord(pb.ev, st) == cudaSuccess;
    if (!ok) { set_error("scp_seqs_create: CUDA allocation/upload failed: %s", cudaGetErrorString(cudaGetLastError())); scp_seqs_destroy(s); return nullptr; }
    uint8_t* d = static_cast<uint8_t*>(s->d_arena);
    s->d_off = reinterpret_cast<long long*>(d + o_off);
    s->d_win_seq = reinterpret_cast<int*>(d + o_ws); s->d_win_idx = reinterpret_cast<int*>(d + o_wi);
    s->d_tile_seq = reinterpret_cast<int*>(d + o_ts); s->d_tile_start = reinterpret_cast<int*>(d + o_tt);
    s->d_tile128_seq = reinterpret_cast<int*>(d + o_2s); s->d_tile128_start = reinterpret_cast<int*>(d + o_2t);
    return s;
}

scp_seqs* scp_seqs_create(const int64_t* h_offsets, int n_seq) { return scp_seqs_create_async(h_offsets, n_seq, nullptr); }

void scp_seqs_destroy(scp_seqs* s) {
    if (!s) return;
    if (s->d_arena) cudaFreeAsync(s->d_arena, s->stream);                  // stream-ordered: later kernels of the stream are unaffected
    if (s->h_pinned) pinned_put(PinnedBlock{s->h_pinned, s->pinned_bytes, s->copied});
    delete s;
}

int64_t scp_seqs_total(const scp_seqs* s) { return s ? s->total : -1; }

void scp_gemm_cache_clear(void) { gemm_cache_clear(); }
void scp_gemm_cache_drop(const float* d_w) { gemm_cache_drop(d_w); }

int scp_set_knn_engine(int use_tensor_cores) { int old = g_knn_tc; g_knn_tc = use_tensor_cores ? 1 : 0; return old; }

int scp_set_attn_engine(int mode) { int old = g_attn_tc; g_attn_tc = mode < 0 ? 0 : (mode > 2 ? 2 : mode); return old; }

int scp_set_auto_engine(int mode) { int old = g_auto_tf32; g_auto_tf32 = mode < 0 ? 0 : (mode > 2 ? 2 : mode); return old; }

int scp_set_gemm_cluster(int ctas) { int old = g_gemm_cluster; g_gemm_cluster = ctas >= 4 ? 4 : (ctas >= 2 ? 2 : 1); return old; }

int scp_linear_tf32_supported(int64_t ldx, int64_t ldy, int64_t M, int N, int K) {
    return linear_tf32_ok(ldx, ldy, M, N, K, nullptr, nullptr, nullptr) ? 1 : 0;
}

int scp_linear(const float* d_x, int64_t ldx, const float* d_w, const float* d_bias, const float* d_res, int64_t ldr,
               float* d_y, int64_t ldy, int64_t M, int N, int K, int act, int engine, void* stream) {
    SCP_REQUIRE(d_x && d_w && d_y && M >= 0 && N > 0 && K > 0, "scp_linear: bad argument");
    SCP_REQUIRE(ldx >= K && ldy >= N && (!d_res || ldr >= N), "scp_linear: leading dimension too small");
    if (M == 0) return SCP_OK;
    cudaStream_t st = as_stream(stream);
    const bool tc_ok = linear_tf32_ok(ldx, ldy, M, N, K, d_x, d_w, d_y);
    // SCP_GEMM_TF32 / SCP_GEMM_TF32X3: tensor cores wherever the shape allows (tiny / unaligned layers stay on the fp32
    // tiles); SCP_GEMM_AUTO: error-compensated tensor cores for every layer with N >= 64.  The choice must not depend on
    // M: a window has to produce bit-identical logits whatever batch it is encoded in (frame partition = multi-GPU)
    if (tc_ok && engine == SCP_GEMM_TF32) return linear_tf32(d_x, ldx, d_w, d_bias, d_res, ldr, d_y, ldy, M, N, K, act, st, 0);
    if (tc_ok && (engine == SCP_GEMM_F16X3 || (engine == SCP_GEMM_AUTO && g_auto_tf32 == 2 && N >= 64)))
        return linear_tf32(d_x, ldx, d_w, d_bias, d_res, ldr, d_y, ldy, M, N, K, act, st, 2);
    if (tc_ok && (engine == SCP_GEMM_TF32X3 || (engine == SCP_GEMM_AUTO && g_auto_tf32 && N >= 64)))
        return linear_tf32(d_x, ldx, d_w, d_bias, d_res, ldr, d_y, ldy, M, N, K, act, st, 1);
    const int vec4 = (K % 4 == 0) && (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_x) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(d_w) & 15) == 0);
    if (N > 64 && M > 2048) {
        dim3 grid((unsigned)cdiv(N, 128), (unsigned)cdiv(M, 128));
        k_linear_simt<128, 128, 8, 8><<<grid, 256, 0, st>>>(d_x, ldx, d_w, d_bias, d_res, ldr, d_y, ldy, M, N, K, act, vec4);
    } else {
        dim3 grid((unsigned)cdiv(N, 64), (unsigned)cdiv(M, 64));
        k_linear_simt<64, 64, 4, 4><<<grid, 256, 0, st>>>(d_x, ldx, d_w, d_bias, d_res, ldr, d_y, ldy, M, N, K, act, vec4);
    }
    SCP_LAUNCHED();
    return SCP_OK;
}

int scp_layernorm(const float* d_x, int64_t ldx, const float* d_res, int64_t ldr, const float* d_gamma,
                  const float* d_beta, float* d_y, int64_t ldy, int64_t M, int C, float eps, void* stream) {
    SCP_REQUIRE(d_x && d_gamma && d_beta && d_y && C > 0 && C <= 640, "scp_layernorm: bad argument (C<=640)");
    if (M == 0) return SCP_OK;
    auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    const bool v4 = C % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && al(d_x) && al(d_y) && al(d_gamma) && al(d_beta) &&
                    (!d_res || (ldr % 4 == 0 && al(d_res)));
    if (v4 && C == 256) {
        k_layernorm_v4<2><<<(unsigned)cdiv(M, 8), 256, 0, as_stream(stream)>>>(d_x, ldx, d_res, ldr, d_gamma, d_beta, d_y, ldy, M, eps);
        SCP_LAUNCHED();
        return SCP_OK;
    }
    if (v4 && C == 512) {
        k_layernorm_v4<4><<<(unsigned)cdiv(M, 8), 256, 0, as_stream(stream)>>>(d_x, ldx, d_res, ldr, d_gamma, d_beta, d_y, ldy, M, eps);
        SCP_LAUNCHED();
        return SCP_OK;
    }
    if (v4) {                                    // other widths (OctAttention: 600) on the float4 path too
        k_layernorm_v4g<5><<<(unsigned)cdiv(M, 8), 256, 0, as_stream(stream)>>>(d_x, ldx, d_res, ldr, d_gamma, d_beta, d_y, ldy, M, C, eps);
        SCP_LAUNCHED();
        return SCP_OK;
    }
    k_layernorm<<<(unsigned)cdiv(M, 8), 256, 0, as_stream(stream)>>>(d_x, ldx, d_res, ldr, d_gamma, d_beta, d_y, ldy, M, C, eps);
    SCP_LAUNCHED();
    return SCP_OK;
}

int scp_ehem_embed(const uint8_t* d_ctx, int64_t n, const float* d_occ_enc, const float* d_level_enc, int n_level_rows,
                   const float* d_octant_enc, float* d_out, int64_t ldo, void* stream) {
    SCP_REQUIRE(d_ctx && d_occ_enc && d_level_enc && d_octant_enc && d_out && ldo >= 80, "scp_ehem_embed: bad argument");
    if (n == 0) return SCP_OK;
    k_ehem_embed<<<grid_for(n * 80, 256 * 4), 256, 0, as_stream(stream)>>>(d_ctx, n, d_occ_enc, d_level_enc, n_level_rows,
                                                                           d_octant_enc, d_out, ldo);
    SCP_LAUNCHED();
    return SCP_OK;
}

int scp_ehem_embed_occ(const uint8_t* d_ctx, int64_t n_even, const float* d_occ_enc, float* d_out, int64_t ldo,
                       void* stream) {
    SCP_REQUIRE(d_ctx && d_occ_enc && d_out && ldo >= 16, "scp_ehem_embed_occ: bad argument");
    if (n_even == 0) return SCP_OK;
    k_ehem_embed_occ<<<grid_for(n_even * 16, 256 * 4), 256, 0, as_stream(stream)>>>(d_ctx, n_even, d_occ_enc, d_out, ldo);
    SCP_LAUNCHED();
    return SCP_OK;
}

int scp_knn(const float* d_x, int64_t ldx, int d, const scp_seqs* seqs, int k, int32_t* d_idx, void* stream) {
    SCP_REQUIRE(d_x && seqs && d_idx && d > 0 && k > 0 && k <= 32, "scp_knn: bad argument (k<=32)");
    if (seqs->total == 0) return SCP_OK;
    cudaStream_t st = as_stream(stream);
    if (d <= 4) {
        const int smem_s = KS_C * 16 + k * KS_Q * 12;
        static int attr_s = 0;
        if (smem_s > attr_s) {
            SCP_CUDA(cudaFuncSetAttribute(k_knn_small, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_s));
            attr_s = smem_s;
        }
        float* aabb = nullptr;
        SCP_CUDA(malloc_async((void**)&aabb, (size_t)seqs->n_tile128 * 32 + 64, st));
        k_tile_aabb<<<(unsigned)cdiv(seqs->n_tile128, 8), 256, 0, st>>>(d_x, ldx, d, seqs->d_off, seqs->d_tile128_seq, seqs->d_tile128_start,
                                                                       seqs->n_tile128, aabb);
        SCP_LAUNCHED();
        k_knn_small<<<seqs->n_tile128, KS_Q, smem_s, st>>>(d_x, ldx, d, seqs->d_off, seqs->d_tile128_seq, seqs->d_tile128_start, k, d_idx, aabb);
        SCP_LAUNCHED();
        SCP_CUDA(cudaFreeAsync(aabb, st));
        return SCP_OK;
    }
    if (g_knn_tc && knn_tc_ok(d, k))
        return knn_tc(d_x, ldx, d, seqs->h_off.data(), seqs->n_seq, seqs->d_off, seqs->d_tile128_seq, seqs->d_tile128_start,
                      seqs->n_tile128, k, d_idx, st);
    float* xx = nullptr;
    SCP_CUDA(malloc_async((void**)&xx, seqs->total * 4, st));
    const float* x0 = d_x + seqs->h_off[0] * ldx;
    k_row_sqnorm<<<(unsigned)cdiv(seqs->total, 8), 256, 0, st>>>(x0, ldx, d, seqs->total, xx);
    SCP_LAUNCHED();
    const int smem = (KD * (KQ + 4) + KD * (KC + 4) + KQ * (KC + 1) + 2 * k * 256) * 4;
    static int attr_smem = 0;
    if (smem > attr_smem) {
        SCP_CUDA(cudaFuncSetAttribute(k_knn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_smem = smem;
    }
    // xx is indexed by global row: shift so that xx[row] works for rows starting at h_off[0]
    k_knn<<<seqs->n_tile, 256, smem, st>>>(d_x, ldx, d, xx - seqs->h_off[0], seqs->d_off, seqs->d_tile_seq,
                                           seqs->d_tile_start, k, d_idx);
    SCP_LAUNCHED();
    SCP_CUDA(cudaFreeAsync(xx, st));
    return SCP_OK;
}

int scp_edge_gather_max2(const float* d_uv, int64_t lduv, int C, const int32_t* d_idx, int k, int64_t n,
                         const float* d_bn_scale, const float* d_bn_shift, float* d_out, int64_t ldo, float* d_out2, int64_t ldo2,
                         void* stream) {
    SCP_REQUIRE(d_uv && d_idx && d_bn_scale && d_bn_shift && d_out, "scp_edge_gather_max: null argument");
    SCP_REQUIRE(C == 64 || C == 128 || C == 256, "scp_edge_gather_max: C must be 64, 128 or 256");
    if (n == 0) return SCP_OK;
    k_edge_gather<<<grid_for(n, 256 / C, 148 * 32), 256, 0, as_stream(stream)>>>(d_uv, lduv, C, d_idx, k, n, d_bn_scale,
                                                                                 d_bn_shift, d_out, ldo, d_out2, ldo2);
    SCP_LAUNCHED();
    return SCP_OK;
}

int scp_edge_gather_max(const float* d_uv, int64_t lduv, int C, const int32_t* d_idx, int k, int64_t n,
                        const float* d_bn_scale, const float* d_bn_shift, float* d_out, int64_t ldo, void* stream) {
    return scp_edge_gather_max2(d_uv, lduv, C, d_idx, k, n, d_bn_scale, d_bn_shift, d_out, ldo, nullptr, 0, stream);
}

int scp_swin_attention(const float* d_q, int64_t ldq, const float* d_k, int64_t ldk, const float* d_v, int64_t ldv,
                       const float* d_qb, const float* d_kb, const float* d_vb, const float* d_relpos, int heads,
                       const scp_seqs* seqs, int shift, float* d_out, int64_t ldo, void* stream) {
    SCP_REQUIRE(d_q && d_k && d_v && d_qb && d_kb && d_vb && d_relpos && seqs && d_out, "scp_swin_attention: null argument");
    SCP_REQUIRE(heads > 0 && heads <= 16 && (shift == 0 || shift == 256), "scp_swin_attention: heads/shift");
    SCP_REQUIRE(ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(d_out) & 15) == 0, "scp_swin_attention: out must be 16B aligned");
    if (seqs->n_win == 0) return SCP_OK;
    if (g_attn_tc == 2 && heads * 64 <= 1024 && swin_attn_tc_ok(ldq, ldk, ldv, ldo, d_q, d_k, d_v, d_out, d_qb, d_kb, d_vb))
        return swin_attn_h(d_q, ldq, d_k, ldk, d_v, ldv, d_qb, d_kb, d_vb, d_relpos, heads, seqs->d_off, seqs->d_win_seq,
                           seqs->d_win_idx, seqs->n_win, shift, d_out, ldo, as_stream(stream));
    if (g_attn_tc && heads * 64 <= 1024 && swin_attn_tc_ok(ldq, ldk, ldv, ldo, d_q, d_k, d_v, d_out, d_qb, d_kb, d_vb))
        return swin_attn_tc(d_q, ldq, d_k, ldk, d_v, ldv, d_qb, d_kb, d_vb, d_relpos, heads, seqs->d_off, seqs->d_win_seq,
                            seqs->d_win_idx, seqs->n_win, shift, d_out, ldo, as_stream(stream));
    const int smem = 4 * HD * ALD * 4;
    static bool attr = false;
    if (!attr) { SCP_CUDA(cudaFuncSetAttribute(k_swin_attn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr = true; }
    dim3 grid((WS / AQ) * heads, seqs->n_win);
    k_swin_attn<<<grid, 256, smem, as_stream(stream)>>>(d_q, ldq, d_k, ldk, d_v, ldv, d_qb, d_kb, d_vb, d_relpos, heads,
                                                        seqs->d_off, seqs->d_win_seq, seqs->d_win_idx, shift, d_out, ldo);
    SCP_LAUNCHED();
    return SCP_OK;
}

int scp_pair_concat(const float* d_x, int64_t ldx, const scp_seqs* src, const scp_seqs* dst, int C, float* d_out,
                    int64_t ldo, void* stream) {
    SCP_REQUIRE(d_x && src && dst && d_out && src->n_seq == dst->n_seq && C % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0,
                "scp_pair_concat: bad argument");
    if (dst->n_tile == 0) return SCP_OK;
    k_pair_concat<<<dst->n_tile, 256, 0, as_stream(stream)>>>(d_x, ldx, src->d_off, dst->d_off, dst->d_tile_seq,
                                                              dst->d_tile_start, C, d_out, ldo);
    SCP_LAUNCHED();
    return SCP_OK;
}

int scp_upsample_cols(const float* d_src, int64_t lds, const scp_seqs* src, const scp_seqs* dst, int shift, int C,
                      float* d_out, int64_t ldo, int col_off, void* stream) {
    SCP_REQUIRE(d_src && src && dst && d_out && src->n_seq == dst->n_seq && C % 4 == 0 && lds % 4 == 0 && ldo % 4 == 0 &&
                col_off % 4 == 0 && shift >= 0 && shift < 16, "scp_upsample_cols: bad argument");
    if (dst->n_tile == 0) return SCP_OK;
    k_upsample_cols<<<dst->n_tile, 256, 0, as_stream(stream)>>>(d_src, lds, src->d_off, dst->d_off, dst->d_tile_seq,
                                                                dst->d_tile_start, shift, C, d_out, ldo, col_off);
    SCP_LAUNCHED();
    return SCP_OK;
}

int scp_copy_cols(const float* d_src, int64_t lds, int64_t row_step, int64_t row_off, int64_t rows, int C, float* d_out,
                  int64_t ldo, int col_off, void* stream) {
    SCP_REQUIRE(d_src && d_out && rows >= 0 && C > 0, "scp_copy_cols: bad argument");
    if (rows == 0) return SCP_OK;
    k_copy_cols<<<grid_for(rows * C, 1024), 256, 0, as_stream(stream)>>>(d_src, lds, row_step, row_off, rows, C, d_out, ldo, col_off);
    SCP_LAUNCHED();
    return SCP_OK;
}

int scp_add(const float* d_a, const float* d_b, float* d_y, int64_t n, void* stream) {
    SCP_REQUIRE(d_a && d_b && d_y && n >= 0, "scp_add: bad argument");
    if (n == 0) return SCP_OK;
    k_add<<<grid_for(n, 1024), 256, 0, as_stream(stream)>>>(d_a, d_b, d_y, n);
    SCP_LAUNCHED();
    return SCP_OK;
}

int scp_octattn_embed(const uint8_t* d_ctx, const uint32_t* d_ctx_pos, float pos_scale, int level_base,
                      int max_octree_level, const scp_seqs* seqs, const float* d_occ_enc, const float* d_level_enc,
                      const float* d_octant_enc, const float* d_pos_w, const float* d_pos_b, const float* d_pe,
                      float* d_embed, float* d_embed_unknown, void* stream) {
    SCP_REQUIRE(d_ctx && d_ctx_pos && seqs && d_occ_enc && d_level_enc && d_octant_enc && d_pos_w && d_pos_b &&
                d_embed && d_embed_unknown, "scp_octattn_embed: null argument");      // d_pe may be NULL (pos_embed False)
    if (seqs->n_tile == 0) return SCP_OK;
    k_octattn_embed<<<seqs->n_tile, 256, 0, as_stream(stream)>>>(d_ctx, d_ctx_pos, pos_scale, level_base, max_octree_level,
                                                                 seqs->d_off, seqs->d_tile_seq, seqs->d_tile_start, d_occ_enc,
                                                                 d_level_enc, d_octant_enc, d_pos_w, d_pos_b, d_pe, d_embed,
                                                                 d_embed_unknown);
    SCP_LAUNCHED();
    return SCP_OK;
}

int scp_octattn_attention(const float* d_qu, const float* d_k, const float* d_ku, const float* d_v, const float* d_vu,
                          int64_t ld, int heads, int head_dim, const scp_seqs* seqs, float* d_out, float* d_out_u,
                          int64_t ldo, void* stream) {
    SCP_REQUIRE(d_qu && d_k && d_ku && d_v && d_vu && seqs && d_out && d_out_u, "scp_octattn_attention: null argument");
    SCP_REQUIRE(head_dim > 0 && head_dim <= 150 && heads > 0, "scp_octattn_attention: head_dim must be <= 150");
    if (seqs->n_tile == 0) return SCP_OK;
    // engine: 1 (default) = tcgen05 / TMEM / TMA flash kernel (octattn_h.cu), 0 = fp32 SIMT kernel (kept as the A/B reference)
    static const int engine = getenv("SCP_OCTATTN_ENGINE") ? atoi(getenv("SCP_OCTATTN_ENGINE")) : 1;
    if (engine == 1 && octattn_h_ok(ld, ldo, head_dim, d_qu, d_k, d_ku, d_v, d_vu, d_out, d_out_u))
        return octattn_attn_h(d_qu, d_k, d_ku, d_v, d_vu, ld, heads, seqs->h_off.data(), seqs->n_seq, seqs->d_off, seqs->d_tile_seq,
                              seqs->d_tile_start, seqs->n_tile, seqs->d_tile128_seq, seqs->d_tile128_start, seqs->n_tile128,
                              d_out, d_out_u, ldo, as_stream(stream));
    dim3 grid(seqs->n_tile * 4, heads);
    k_octattn_attn<<<grid, 128, 0, as_stream(stream)>>>(d_qu, d_k, d_ku, d_v, d_vu, ld, heads, head_dim, seqs->d_off,
                                                        seqs->d_tile_seq, seqs->d_tile_start, d_out, d_out_u, ldo);
    SCP_LAUNCHED();
    return SCP_OK;
}

}  // extern "C"

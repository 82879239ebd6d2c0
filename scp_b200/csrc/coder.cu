// Coding order (A7), softmax -> integer CDF (A13) and the host range coder (A14).
//
// Reference behaviour reproduced: encode.py:109-136, encode_mullevel.py:106-133 (even ids then odd ids per
// 8192-window); numpyAc/numpyAc.py:109-114 (float32 sequential cumsum, normalise, prepend 0) and :80-107
// (x 65281, rint, int16 wrap, + arange); numpyAc/backend/numpyAc_backend.cpp:245-323 (32-bit range coder).
#include <vector>
#include <string.h>
#include "common.cuh"

namespace scp {

// ------------------------------------------------------------------------------------------
// A7 coding order
// ------------------------------------------------------------------------------------------
struct Win { long long base; long long start; long long frame_base; int len; int single; };

__global__ void __launch_bounds__(256) k_coding_order(const Win* __restrict__ wins, int n_win,
                                                       const uint8_t* __restrict__ occ, long long* __restrict__ order,
                                                       int16_t* __restrict__ sym, int add_base_for_single) {
    for (int w = blockIdx.x; w < n_win; w += gridDim.x) {
        const Win W = wins[w];
        const int half = (W.len + 1) >> 1;
        for (int l = threadIdx.x; l < W.len; l += blockDim.x) {
            long long row = W.base + W.start + l;
            long long id = row;
            int p = (l & 1) ? half + (l >> 1) : (l >> 1);
            if (W.single) { p = 0; id = add_base_for_single ? row : (row - W.base + W.frame_base); }   // encode.py:123 vs encode_mullevel.py:120
            order[W.base + W.start + p] = id;
            if (sym) sym[W.base + W.start + p] = (int16_t)((int)occ[id] - 1);
        }
    }
}

// ------------------------------------------------------------------------------------------
// context-window assembly (all windows of a batch, odd windows get the pad token of ehem.py:92-99)
// ------------------------------------------------------------------------------------------
struct GWin { long long row; long long tok; int len; int pad; };

__global__ void __launch_bounds__(256) k_gather_windows(const GWin* __restrict__ wins, const uint8_t* __restrict__ ctx,
                                                         const float* __restrict__ pos, uint8_t* __restrict__ ctx_out,
                                                         float* __restrict__ pos_out, long long* __restrict__ row_even,
                                                         long long* __restrict__ row_odd) {
    const GWin W = wins[blockIdx.y];
    const int plen = W.len + (W.len & 1);
    const u32* ci = reinterpret_cast<const u32*>(ctx);
    u32* co = reinterpret_cast<u32*>(ctx_out);
    for (int l = blockIdx.x * 256 + threadIdx.x; l < plen; l += gridDim.x * 256) {
        const long long t = W.tok + l;
        const bool real = l < W.len;
        const long long r = W.row + l;
        if (real) {
            co[3 * t] = ci[3 * r]; co[3 * t + 1] = ci[3 * r + 1]; co[3 * t + 2] = ci[3 * r + 2];
            pos_out[3 * t] = pos[3 * r]; pos_out[3 * t + 1] = pos[3 * r + 1]; pos_out[3 * t + 2] = pos[3 * r + 2];
        } else {
            co[3 * t] = 0x00ff0000u; co[3 * t + 1] = 0x0000ff00u; co[3 * t + 2] = 0xff0000ffu;   // 4 x (0,0,255)
            pos_out[3 * t] = 0.f; pos_out[3 * t + 1] = 0.f; pos_out[3 * t + 2] = 0.f;
        }
        long long* dst = (l & 1) ? row_odd : row_even;
        if (dst) dst[t >> 1] = real ? r : -1;
    }
}

// OctAttention sequences (encode.py:23-82 / encode_dataset.py:31-55): [pad rows ; nodes of one row file] per sequence, the
// pad rows being (level 0, octant 0, occupancy 255) with zero positions and node id -1; ancestor positions are shifted so
// that one scale 2^-21 serves sequences of different depth (pos / 2^max_level == (pos << (21 - max_level)) / 2^21 exactly)
struct PSeq { long long dst; long long src; int len; int pad; int shift; int _; };

__global__ void __launch_bounds__(256) k_pad_gather_seqs(const PSeq* __restrict__ seqs, const uint8_t* __restrict__ ctx,
                                                          const u32* __restrict__ cpos, uint8_t* __restrict__ ctx_out,
                                                          u32* __restrict__ pos_out, long long* __restrict__ row_of) {
    const PSeq S = seqs[blockIdx.y];
    const u32* ci = reinterpret_cast<const u32*>(ctx);
    u32* co = reinterpret_cast<u32*>(ctx_out);
    for (int l = blockIdx.x * 256 + threadIdx.x; l < S.len; l += gridDim.x * 256) {
        const long long t = S.dst + l;
        const bool real = l >= S.pad;
        const long long r = S.src + (l - S.pad);
        if (real) {
#pragma unroll
            for (int j = 0; j < 3; ++j) co[3 * t + j] = ci[3 * r + j];
#pragma unroll
            for (int j = 0; j < 12; ++j) pos_out[12 * t + j] = cpos[12 * r + j] << S.shift;
        } else {
            co[3 * t] = 0x00ff0000u; co[3 * t + 1] = 0x0000ff00u; co[3 * t + 2] = 0xff0000ffu;   // 4 x (0,0,255)
#pragma unroll
            for (int j = 0; j < 12; ++j) pos_out[12 * t + j] = 0u;
        }
        if (row_of) row_of[t] = real ? r : -1;
    }
}

__global__ void __launch_bounds__(256) k_gather_rows8(const u64* __restrict__ in, const long long* __restrict__ idx,
                                                       long long n, u64* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) out[i] = in[idx[i]];
}

// ------------------------------------------------------------------------------------------
// A13 logits/PMF -> uint16 CDF
// ------------------------------------------------------------------------------------------
constexpr int CDF_R = 64;        // rows per block
constexpr int CDF_T = 256;       // threads per block: 8 warps load + softmax 8 rows each, 64 threads run the sequential cumsums
constexpr int CDF_LD = 257;      // padded row stride: (257*r + i) % 32 is conflict-free both ways

// Phase 1 (all 8 warps): a warp streams one row at a time -- 255 floats, lane + 32 j, fully coalesced -- and, for logits, does
// the softmax in registers (max and sum by xor-shuffle trees, IEEE divide), then parks the PMF row in shared memory.
// Phase 2 (64 threads, one per row): np.cumsum in float32 is a strictly sequential chain of 254 additions per row
// (numpyAc.py:111) and has to stay one to be bit-exact; the other warps of the block idle through it, the other resident
// blocks of the SM cover it.  With only the (c_low, c_high) interval requested -- the encoder's path -- the running sum is
// kept in a register and sampled at the symbol, nothing is written back.
// Phase 3: normalise (float32 divide by the last sum), float64 x 65281 + rint + int16 wrap (numpyAc.py:80-107).
__global__ void __launch_bounds__(CDF_T) k_pmf_to_cdf(const float* __restrict__ in, long long n, int is_logits,
                                                       const long long* __restrict__ row_of, const int16_t* __restrict__ sym,
                                                       uint16_t* __restrict__ cdf, u32* __restrict__ interval,
                                                       float* __restrict__ pmf) {
    extern __shared__ float s[];                 // [CDF_R][CDF_LD]
    __shared__ long long s_orow[CDF_R];
    const long long row0 = (long long)blockIdx.x * CDF_R;
    const int rows = (int)min((long long)CDF_R, n - row0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < rows) s_orow[threadIdx.x] = row_of ? row_of[row0 + threadIdx.x] : row0 + threadIdx.x;
    // four rows per warp in flight: all 32 loads are issued before the first softmax (one row at a time left the warp waiting
    // ~1 us of DRAM latency per row)
    constexpr int RB = 4;
    for (int r0 = warp * RB; r0 < rows; r0 += (CDF_T / 32) * RB) {
        float e[RB][8];
#pragma unroll
        for (int q = 0; q < RB; ++q) {
            const float* x = in + (row0 + r0 + q) * 255;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = lane + 32 * j;
                e[q][j] = (r0 + q < rows && c < 255) ? __ldg(x + c) : -INFINITY;
            }
        }
#pragma unroll
        for (int q = 0; q < RB; ++q) {
            if (r0 + q >= rows) break;                // warp-uniform
            if (is_logits) {                          // torch.softmax(output, 2)  (encode.py:126-127)
                float m = -INFINITY;
#pragma unroll
                for (int j = 0; j < 8; ++j) m = fmaxf(m, e[q][j]);
                m = warp_max(m);
                // 2^((x - max) log2 e) on the special-function unit and one reciprocal per row: the softmax was 80 % of the
                // kernel's instructions with expf / IEEE divides (0.16 ms of issue time per 514 k rows); the PMF differs from
                // torch.softmax by ~1e-7 either way, and encoder and decoder run this same kernel
                const float ml = m * 1.4426950408889634f;
                float sum = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float p;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(fmaf(e[q][j], 1.4426950408889634f, -ml)));
                    e[q][j] = lane + 32 * j < 255 ? p : 0.f;
                    sum += e[q][j];
                }
                sum = warp_sum(sum);
                const float inv = __frcp_rn(sum);
#pragma unroll
                for (int j = 0; j < 8; ++j) e[q][j] *= inv;
            }
            float* dst = s + (r0 + q) * CDF_LD;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = lane + 32 * j;
                if (c < 255) dst[c] = e[q][j];
            }
        }
    }
    __syncthreads();
    if (pmf) {
        for (int i = threadIdx.x; i < rows * 255; i += CDF_T) {
            int r = i / 255, c = i - r * 255;
            const long long orow = s_orow[r];
            if (orow < 0) continue;
            pmf[orow * 255 + c] = s[r * CDF_LD + c];
        }
    }
    if (!cdf) {
        // interval only: sequential float32 cumsum in a register, sampled at sym-1, sym and the end
        if (interval && threadIdx.x < rows) {
            const long long orow = s_orow[threadIdx.x];
            if (orow >= 0) {
                const float* x = s + threadIdx.x * CDF_LD;
                int sy = sym[orow];
                sy = sy < 0 ? 0 : (sy > 254 ? 254 : sy);
                float acc = x[0];
                float c_lo = 0.f, c_hi = acc;                     // cumsum[sy-1] (unused for sy == 0), cumsum[sy]
#pragma unroll 8
                for (int c = 1; c < 255; ++c) {
                    if (c == sy) c_lo = acc;
                    acc = __fadd_rn(acc, x[c]);
                    if (c == sy) c_hi = acc;
                }
                const float last = acc;
                const double Fl = sy == 0 ? 0.0 : (double)__fdiv_rn(c_lo, last);
                const u32 lo = (u32)(((long long)rint(Fl * 65281.0) + sy) & 0xffff);
                u32 hi = 0x10000u;                                // numpyAc_backend.cpp:277
                if (sy != 254) hi = (u32)(((long long)rint((double)__fdiv_rn(c_hi, last) * 65281.0) + sy + 1) & 0xffff);
                interval[2 * orow] = lo;
                interval[2 * orow + 1] = hi;
            }
        }
        return;
    }
    // np.cumsum(pdf, axis=1) in float32: strictly sequential adds (numpyAc.py:111)
    if (threadIdx.x < rows) {
        float* x = s + threadIdx.x * CDF_LD;
        float acc = x[0];
#pragma unroll 8
        for (int c = 1; c < 255; ++c) { acc = __fadd_rn(acc, x[c]); x[c] = acc; }
    }
    __syncthreads();
    // cdfF/cdfF[:, -1:] (float32) -> float64 [0, f...] * 65281 -> rint -> int16 wrap -> + arange(256)
    for (int i = threadIdx.x; i < rows * 128; i += CDF_T) {
        int r = i >> 7, c2 = (i & 127) * 2;
        const float* x = s + r * CDF_LD;
        const float last = x[254];
        u32 v[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            int c = c2 + k;                         // position in the 256-long CDF
            double F = c == 0 ? 0.0 : (double)__fdiv_rn(x[c - 1], last);
            long long q = (long long)rint(F * 65281.0) + c;
            v[k] = (u32)(q & 0xffff);
        }
        const long long orow = s_orow[r];
        if (orow < 0) continue;
        reinterpret_cast<u32*>(cdf + orow * 256)[c2 >> 1] = v[0] | (v[1] << 16);
    }
    if (interval && threadIdx.x < rows) {
        const float* x = s + threadIdx.x * CDF_LD;
        const float last = x[254];
        const long long orow = s_orow[threadIdx.x];
        if (orow < 0) return;
        int sy = sym[orow];
        sy = sy < 0 ? 0 : (sy > 254 ? 254 : sy);
        double Fl = sy == 0 ? 0.0 : (double)__fdiv_rn(x[sy - 1], last);
        u32 lo = (u32)(((long long)rint(Fl * 65281.0) + sy) & 0xffff);
        u32 hi = 0x10000u;                          // numpyAc_backend.cpp:277
        if (sy != 254) hi = (u32)(((long long)rint((double)__fdiv_rn(x[sy], last) * 65281.0) + sy + 1) & 0xffff);
        interval[2 * orow] = lo;
        interval[2 * orow + 1] = hi;
    }
}

// ------------------------------------------------------------------------------------------
// A14 range coder (host).  Classic 32-bit low/high coder with E1/E2 shifts and E3 pending bits.
// ------------------------------------------------------------------------------------------
struct BitSink {
    uint8_t* out; long long cap; long long n = 0; uint32_t acc = 0; int fill = 0; bool overflow = false;
    inline void put(int bit) {
        acc = (acc << 1) | (uint32_t)bit;
        if (++fill == 8) { if (out) { if (n < cap) out[n] = (uint8_t)acc; else overflow = true; } ++n; acc = 0; fill = 0; }
    }
    inline void put_with_pending(int bit, unsigned long long& pending) {
        put(bit);
        for (; pending > 0; --pending) put(!bit);
    }
    inline void flush() { while (fill != 0) put(0); }
};

template <class GetInterval>
static long long range_encode_impl(long long n, uint8_t* out, long long cap, GetInterval get) {
    BitSink sink{out, cap};
    uint32_t low = 0, high = 0xFFFFFFFFu;
    unsigned long long pending = 0;
    for (long long i = 0; i < n; ++i) {
        uint32_t c_low, c_high;
        get(i, c_low, c_high);
        const uint64_t span = (uint64_t)high - (uint64_t)low + 1;
        high = (low - 1) + (uint32_t)((span * (uint64_t)c_high) >> 16);
        low = low + (uint32_t)((span * (uint64_t)c_low) >> 16);
        for (;;) {
            if (high < 0x80000000u) {
                sink.put_with_pending(0, pending);
            } else if (low >= 0x80000000u) {
                sink.put_with_pending(1, pending);
            } else if (low >= 0x40000000u && high < 0xC0000000u) {
                ++pending;
                low = (low << 1) & 0x7FFFFFFFu;
                high = (high << 1) | 0x80000001u;
                continue;
            } else {
                break;
            }
            low <<= 1;
            high = (high << 1) | 1u;
        }
    }
    ++pending;
    sink.put_with_pending(low < 0x40000000u ? 0 : 1, pending);
    sink.flush();
    if (sink.overflow) { set_error("scp_range_encode: output buffer too small (%lld needed)", sink.n); return SCP_ERR_ARG; }
    return sink.n;
}


// Decoder mirror of range_encode_impl (numpyAc_backend.cpp:134-229 `decode`): same low/high/value registers, same E1/E2/E3
// renormalisation, symbol by binary search for the last CDF entry <= count.  Stateful, so that the caller can hand over the
// CDF rows in the order the symbols were coded, a window at a time (the entropy model needs decoded symbols to go on).
struct RangeDecoder {
    std::vector<uint8_t> in;
    size_t in_ptr = 0;
    uint8_t cache = 0;
    int cached_bits = 0;
    uint32_t low = 0, high = 0xFFFFFFFFu, value = 0;
    long long decoded = 0;
    inline void get() {
        if (cached_bits == 0) {
            if (in_ptr == in.size()) { value <<= 1; return; }      // past the end: zeros (numpyAc_backend.cpp:83-86)
            cache = in[in_ptr++];
            cached_bits = 8;
        }
        value = (value << 1) | (uint32_t)((cache >> (cached_bits - 1)) & 1);
        --cached_bits;
    }
    inline int decode_one(const uint16_t* cdf, int Lp) {
        const int max_symbol = Lp - 2;
        const uint64_t span = (uint64_t)high - (uint64_t)low + 1;
        const uint32_t count = (uint32_t)((((uint64_t)value - (uint64_t)low + 1) * 0x10000ull - 1) / span);
        // last s in [0, max_symbol] with cdf[s] <= count (cdf[0] = 0; entry Lp-1 wraps to 0 and is never read: the
        // upper bound of the last symbol is the constant 0x10000, numpyAc_backend.cpp:277)
        int lo = 0, hi = max_symbol + 1;
        while (lo + 1 < hi) {
            const int m = (lo + hi) >> 1;
            if ((uint32_t)cdf[m] <= count) lo = m; else hi = m;
        }
        const int s = lo;
        const uint32_t c_low = cdf[s];
        const uint32_t c_high = s == max_symbol ? 0x10000u : cdf[s + 1];
        high = (low - 1) + (uint32_t)((span * (uint64_t)c_high) >> 16);
        low = low + (uint32_t)((span * (uint64_t)c_low) >> 16);
        for (;;) {
            if (low >= 0x80000000u || high < 0x80000000u) {
                low <<= 1;
                high = (high << 1) | 1u;
                get();
            } else if (low >= 0x40000000u && high < 0xC0000000u) {
                low = (low << 1) & 0x7FFFFFFFu;
                high = (high << 1) | 0x80000001u;
                value -= 0x40000000u;
                get();
            } else {
                break;
            }
        }
        ++decoded;
        return s;
    }
};

}  // namespace scp

using namespace scp;

extern "C" {

int scp_coding_order(const int64_t* h_level_sizes, const uint8_t* h_level_restart, int n_levels, int context_size,
                     int add_base_for_single, const uint8_t* d_occ, int64_t* d_order, int16_t* d_sym, void* stream) {
    SCP_REQUIRE(h_level_sizes && d_order && n_levels > 0 && context_size > 0, "scp_coding_order: bad argument");
    SCP_REQUIRE(!d_sym || d_occ, "scp_coding_order: symbols need d_occ");
    cudaStream_t st = as_stream(stream);
    std::vector<Win> wins;
    long long base = 0, frame_base = 0;
    for (int l = 0; l < n_levels; ++l) {
        long long n = h_level_sizes[l];
        SCP_REQUIRE(n >= 0, "scp_coding_order: negative level size");
        if (h_level_restart && h_level_restart[l]) frame_base = base;
        for (long long i = 0; i < n; i += context_size)
            wins.push_back(Win{base, i, frame_base, (int)std::min<long long>(context_size, n - i), n == 1});
        base += n;
    }
    if (wins.empty()) return SCP_OK;
    Win* d_w = nullptr;
    SCP_CUDA(upload_async((void**)&d_w, wins.data(), wins.size() * sizeof(Win), st));
    int grid = (int)std::min<size_t>(wins.size(), 148 * 8);
    k_coding_order<<<grid, 256, 0, st>>>(d_w, (int)wins.size(), d_occ, (long long*)d_order, d_sym, add_base_for_single);
    SCP_LAUNCHED();
    SCP_CUDA(cudaFreeAsync(d_w, st));
    return SCP_OK;
}

int scp_gather_windows(const uint8_t* d_ctx, const float* d_pos, const int64_t* h_win_row, const int32_t* h_win_len,
                       const int64_t* h_win_tok, int n_win, uint8_t* d_ctx_out, float* d_pos_out, int64_t* d_row_even,
                       int64_t* d_row_odd, void* stream) {
    SCP_REQUIRE(d_ctx && d_pos && h_win_row && h_win_len && h_win_tok && d_ctx_out && d_pos_out && n_win >= 0,
                "scp_gather_windows: bad argument");
    if (n_win == 0) return SCP_OK;
    cudaStream_t st = as_stream(stream);
    std::vector<GWin> wins(n_win);
    int maxlen = 0;
    for (int w = 0; w < n_win; ++w) {
        SCP_REQUIRE(h_win_len[w] > 0 && (h_win_tok[w] & 1) == 0, "scp_gather_windows: window %d (len>0, even token start)", w);
        wins[w] = GWin{h_win_row[w], h_win_tok[w], h_win_len[w], 0};
        maxlen = std::max(maxlen, h_win_len[w] + 1);
    }
    GWin* d_w = nullptr;
    SCP_CUDA(upload_async((void**)&d_w, wins.data(), wins.size() * sizeof(GWin), st));
    dim3 grid((unsigned)std::min<long long>(cdiv(maxlen, 256), 32), (unsigned)n_win);
    k_gather_windows<<<grid, 256, 0, st>>>(d_w, d_ctx, d_pos, d_ctx_out, d_pos_out, (long long*)d_row_even, (long long*)d_row_odd);
    SCP_LAUNCHED();
    SCP_CUDA(cudaFreeAsync(d_w, st));
    return SCP_OK;
}

int scp_pad_gather_seqs(const uint8_t* d_ctx, const uint32_t* d_ctx_pos, const int64_t* h_dst_start, const int64_t* h_src_start,
                        const int32_t* h_len, const int32_t* h_shift, int n_seq, int pad, uint8_t* d_ctx_out,
                        uint32_t* d_pos_out, int64_t* d_row_of, void* stream) {
    SCP_REQUIRE(d_ctx && d_ctx_pos && h_dst_start && h_src_start && h_len && h_shift && d_ctx_out && d_pos_out && n_seq >= 0 &&
                pad >= 0, "scp_pad_gather_seqs: bad argument");
    if (n_seq == 0) return SCP_OK;
    cudaStream_t st = as_stream(stream);
    std::vector<PSeq> seqs(n_seq);
    int maxlen = 0;
    for (int i = 0; i < n_seq; ++i) {
        SCP_REQUIRE(h_len[i] >= pad && h_shift[i] >= 0 && h_shift[i] < 32, "scp_pad_gather_seqs: sequence %d (len >= pad, shift 0..31)", i);
        seqs[i] = PSeq{h_dst_start[i], h_src_start[i], h_len[i], pad, h_shift[i], 0};
        maxlen = std::max(maxlen, h_len[i]);
    }
    PSeq* d_s = nullptr;
    SCP_CUDA(upload_async((void**)&d_s, seqs.data(), seqs.size() * sizeof(PSeq), st));
    dim3 grid((unsigned)std::min<long long>(cdiv(maxlen, 256), 64), (unsigned)n_seq);
    k_pad_gather_seqs<<<grid, 256, 0, st>>>(d_s, d_ctx, d_ctx_pos, d_ctx_out, d_pos_out, (long long*)d_row_of);
    SCP_LAUNCHED();
    SCP_CUDA(cudaFreeAsync(d_s, st));
    return SCP_OK;
}

int scp_gather_rows8(const void* d_in, const int64_t* d_idx, int64_t n, void* d_out, void* stream) {
    SCP_REQUIRE(d_in && d_idx && d_out && n >= 0, "scp_gather_rows8: bad argument");
    if (n == 0) return SCP_OK;
    k_gather_rows8<<<(unsigned)std::min<long long>(cdiv(n, 256), 148 * 16), 256, 0, as_stream(stream)>>>(
        (const u64*)d_in, (const long long*)d_idx, n, (u64*)d_out);
    SCP_LAUNCHED();
    return SCP_OK;
}

int scp_pmf_to_cdf(const float* d_in, int64_t n, int is_logits, const int64_t* d_row_of, const int16_t* d_sym,
                   uint16_t* d_cdf, uint32_t* d_interval, float* d_pmf, void* stream) {
    SCP_REQUIRE(d_in && n >= 0, "scp_pmf_to_cdf: bad argument");
    SCP_REQUIRE(!d_interval || d_sym, "scp_pmf_to_cdf: intervals need the symbols");
    if (n == 0) return SCP_OK;
    static bool attr_done = false;
    const int smem = CDF_R * CDF_LD * 4;
    if (!attr_done) {
        SCP_CUDA(cudaFuncSetAttribute(k_pmf_to_cdf, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_done = true;
    }
    k_pmf_to_cdf<<<(unsigned)cdiv(n, CDF_R), CDF_T, smem, as_stream(stream)>>>(
        d_in, n, is_logits, (const long long*)d_row_of, d_sym, d_cdf, d_interval, d_pmf);
    SCP_LAUNCHED();
    return SCP_OK;
}

int64_t scp_range_encode(const uint32_t* h_interval, int64_t n, uint8_t* h_out, int64_t out_cap) {
    if (!h_interval || n < 0) { set_error("scp_range_encode: bad argument"); return SCP_ERR_ARG; }
    return range_encode_impl(n, h_out, out_cap, [&](long long i, uint32_t& lo, uint32_t& hi) {
        lo = h_interval[2 * i]; hi = h_interval[2 * i + 1];
    });
}

int64_t scp_range_encode_cdf(const uint16_t* h_cdf, const int16_t* h_sym, int64_t n, int Lp, uint8_t* h_out,
                             int64_t out_cap) {
    if (!h_cdf || !h_sym || n < 0 || Lp < 2) { set_error("scp_range_encode_cdf: bad argument"); return SCP_ERR_ARG; }
    const int max_symbol = Lp - 2;
    for (int64_t i = 0; i < n; ++i)
        if (h_sym[i] < 0 || h_sym[i] > max_symbol) { set_error("scp_range_encode_cdf: symbol %d out of range at %lld", (int)h_sym[i], (long long)i); return SCP_ERR_ARG; }
    return range_encode_impl(n, h_out, out_cap, [&](long long i, uint32_t& lo, uint32_t& hi) {
        const int s = h_sym[i];
        lo = h_cdf[i * Lp + s];
        hi = s == max_symbol ? 0x10000u : h_cdf[i * Lp + s + 1];
    });
}

scp_range_decoder* scp_range_decoder_create(const uint8_t* h_bytes, int64_t n_bytes) {
    if ((!h_bytes && n_bytes > 0) || n_bytes < 0) { set_error("scp_range_decoder_create: bad argument"); return nullptr; }
    RangeDecoder* d = new RangeDecoder();
    d->in.assign(h_bytes, h_bytes + n_bytes);
    for (int i = 0; i < 32; ++i) d->get();
    return reinterpret_cast<scp_range_decoder*>(d);
}

void scp_range_decoder_destroy(scp_range_decoder* d) { delete reinterpret_cast<RangeDecoder*>(d); }

int scp_range_decode(scp_range_decoder* dh, const uint16_t* h_cdf, int64_t n, int Lp, int16_t* h_sym) {
    RangeDecoder* d = reinterpret_cast<RangeDecoder*>(dh);
    if (!d || !h_cdf || !h_sym || n < 0 || Lp < 2) { set_error("scp_range_decode: bad argument"); return SCP_ERR_ARG; }
    for (int64_t i = 0; i < n; ++i) h_sym[i] = (int16_t)d->decode_one(h_cdf + i * Lp, Lp);
    return SCP_OK;
}

int64_t scp_range_decoder_count(const scp_range_decoder* d) { return d ? reinterpret_cast<const RangeDecoder*>(d)->decoded : -1; }

}  // extern "C"

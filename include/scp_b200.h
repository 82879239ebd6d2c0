/* scp_b200 -- C ABI of the B200-native encode-side hot path of SCP.
 *
 * Plain pointers and sizes only; no torch / C++ types cross this boundary.  Every function
 * returns 0 on success and a negative status on failure; scp_last_error() gives the message
 * (thread-local).  All `d_*` pointers are DEVICE pointers on the current CUDA device, all
 * `h_*` pointers are HOST pointers.  `stream` is a cudaStream_t passed as void* (0 = default).
 *
 * Each entry point names the reference interface it replaces (paths in luoao-kddi/SCP).
 */
#ifndef SCP_B200_H
#define SCP_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCP_OK 0
#define SCP_ERR_ARG (-1)
#define SCP_ERR_CUDA (-2)
#define SCP_ERR_RANGE (-3)   /* quantised coordinate does not fit 21 bits / depth > 21 */
#define SCP_ERR_STATE (-4)
#define SCP_ERR_INTERNAL (-5)

#define SCP_MAX_DEPTH 21
#define SCP_MODE_CART 0
#define SCP_MODE_SPHER 1
#define SCP_MODE_CYLIN 2

const char* scp_last_error(void);
int scp_version(void);
/* 1 if the library was compiled for sm_100a and a device with compute capability 10.x is current. */
int scp_device_ok(void);

/* ------------------------------------------------------------------------------------------
 * Octree construction (A1-A5).  Replaces, for a BATCH of jobs in one call:
 *   data_preproc/data_preprocess.py:13-92  proc_pc      (cart2spher/cart2cylin :171-207, quantise :42-70)
 *   data_preproc/data_preprocess.py:95-167 mul_proc_pc  (morton_path filter of Octree.py:188)
 *   data_preproc/OctreeCPP/Octree_python_lib.so::genOctreeInterface (Octreewarpper.py:28-29)
 *   data_preproc/Octree.py:102-137,224-272 gen_K_parent_seq[_mullevel]
 *   dataloaders/encode_dataset_ehem.py:52-105, encode_dataset_ehem_mullevel.py:47-85 (level split, pos normalise)
 *   dataloaders/encode_dataset.py:32-55 (OctAttention context, via ctx_pos)
 * A job = one octree over one frame's points: (frame, qs, morton_path, drop_last).
 * ---------------------------------------------------------------------------------------- */
typedef struct scp_octree scp_octree;

typedef struct {
    int32_t frame;        /* index into frame_offsets */
    int32_t path_len;     /* 0: whole cloud (proc_pc); >0: keep points whose top rho bits == path (Octree.py:188) */
    int32_t path_bits;    /* bit j = morton_path[j] */
    int32_t drop_last;    /* 1: drop the last BFS row (Octree.py:259-262) */
    double  qs;           /* quantisation step of the first axis (data_preprocess.py:44-50) */
    double  cart_offset;  /* SCP_MODE_CART only: scalar offset (data_preprocess.py:56) */
    int32_t lidar_level;  /* level clip of the last level block (encode_dataset_ehem.py:86) */
    int32_t pos_eps_last; /* 1: +1e-9 also in the last block (encode_dataset_ehem.py:92); 0: mullevel dataset (:80) */
} scp_job;

typedef struct {
    int32_t depth;                         /* n = ceil(log2(max(q)+1)), Octree.py:58 */
    int32_t n_points;                      /* points of the frame */
    int32_t n_voxels;                      /* unique occupied voxels after the path filter */
    int32_t n_rows;                        /* emitted rows (nodes, minus 1 if drop_last) */
    int64_t row_start;                     /* first row of this job in the batch outputs */
    int64_t voxel_start;                   /* first voxel of this job in `voxel_key` */
    int32_t level_rows[SCP_MAX_DEPTH + 1]; /* rows per level, index 0 = level 1 */
    float   bin_num;                       /* data_preprocess.py:44/49 (float32 like the reference) */
    double  steps[3];                      /* quantisation steps per axis */
    double  offset[3];                     /* offsets per axis (cylin: [0,0,min z]) */
    int64_t pos_min[SCP_MAX_DEPTH + 1];    /* scalar min of the level's (N_l,3) pos block (encode_dataset_ehem.py:71) */
    int64_t pos_max[SCP_MAX_DEPTH + 1];
} scp_job_info;

typedef struct {
    /* all optional (NULL = skip) except where noted; row-major; N = total rows of the batch */
    uint8_t*  occ;       /* [N]   occupancy byte 1..255 (Octree.py:175-176)            */
    uint8_t*  level;     /* [N]   1-based level                                          */
    uint8_t*  octant;    /* [N]   1..8 (root 1)                                          */
    uint32_t* parent;    /* [N]   row index of the parent inside the job (root: 0)       */
    uint32_t* pos;       /* [N,3] origin of the node's cell (Octree.py:140-145)          */
    uint8_t*  ctx;       /* [N,4,3] EHEM context bytes (level, octant, occ-1), ancestors great-grandparent..self,
                                   missing ancestor = (0,0,255) (encode_dataset_ehem.py:54,67)            */
    float*    pos_norm;  /* [N,3] min-max normalised own pos, float32 (encode_dataset_ehem.py:70-72)      */
    uint32_t* ctx_pos;   /* [N,4,3] pos of the 4 ancestors (0 = missing) (Octree.py:121-122)              */
    int64_t*  rows_i64;  /* [N,4,6] the reference's .npy layout [occ 1..256, level, octant, x, y, z]
                                   (data_preprocess.py:74); an expansion for parity dumps / np.save      */
    uint64_t* voxel_key; /* [V]   Morton keys of the unique voxels per job, ascending                     */
    int16_t*  sym;       /* [N]   coder symbol of the node = occ-1 (encode.py:98,148)                             */
} scp_octree_out;

scp_octree* scp_octree_create(void);
void        scp_octree_destroy(scp_octree* t);

/* Phase 1: quantise, Morton keys, segmented radix sort, unique, per-level node counts.
 * d_xyz: float32 points, `point_stride` floats per point (3, or 4 for KITTI .bin rows, pt.py:190-192).
 * h_frame_offsets[n_frames+1]: point ranges.  Synchronises `stream` once to return sizes. */
int scp_octree_plan(scp_octree* t, const float* d_xyz, int point_stride,
                    const int64_t* h_frame_offsets, int n_frames,
                    const scp_job* h_jobs, int n_jobs, int mode, void* stream);
int     scp_octree_job_info(const scp_octree* t, int job, scp_job_info* out);
int64_t scp_octree_total_rows(const scp_octree* t);
int64_t scp_octree_total_voxels(const scp_octree* t);
/* keys that took part in the sort (points that passed the morton_path filter, summed over the jobs) */
int64_t scp_octree_total_kept(const scp_octree* t);
/* Phase 2: node records, occupancy, K=4 ancestor context, level-wise normalised positions. Asynchronous. */
int scp_octree_emit(scp_octree* t, const scp_octree_out* d_out, void* stream);
/* After emit + stream sync: fills pos_min/pos_max of every job_info. */
int scp_octree_finish(scp_octree* t, void* stream);
/* Device time (ms) of the stages of the last plan+emit, measured with CUDA events on `stream`:
 * [0] quantise+keys [1] radix sort [2] heads/count [3] emit nodes [4] occupancy [5] context gather */
int scp_octree_stage_ms(scp_octree* t, float out[6]);
/* Tree builder used by scp_octree_emit (same outputs, bit for bit): 0 = node records, all levels in one pass over the sorted
 * keys; 1 = node records, one pass per level, bottom-up; 2 (default) = like 0, except that a request for the encoder's outputs
 * only (occ, sym, ctx, pos_norm, voxel_key) is served by ONE warp-autonomous pass over the sorted keys that writes occupancy
 * bytes and node records together (no first-child array, no separate occupancy kernel; [3] = that pass, [4] = 0 in
 * scp_octree_stage_ms), followed by the context gather.  Returns the old value. */
int scp_set_tree_builder(int mode);

/* Standalone pieces of the above, exposed for tests / profiling --------------------------- */
/* Segmented LSD radix sort of 64-bit keys (8-bit digits, decoupled look-back), in place.
 * h_seg_offsets[n_seg+1]; key_bits = number of low bits that take part. */
int scp_segmented_sort_u64(uint64_t* d_keys, uint64_t* d_tmp, const int64_t* h_seg_offsets, int n_seg,
                           int key_bits, void* stream);

/* ------------------------------------------------------------------------------------------
 * Legacy symbols of data_preproc/OctreeCPP/Octree_python_lib.so, exactly as bound by
 * Octreewarpper.py:17-39 (host API; runs the CUDA path above and copies the tree back).
 * ---------------------------------------------------------------------------------------- */
typedef struct { unsigned nodeid; unsigned octant; unsigned parent; uint8_t oct; unsigned pos[3]; } scp_legacy_node;
void* new_vector(void);
void  delete_vector(void* v);
int   vector_size(void* v);
void* vector_get(void* v, int i);
void  vector_push_back(void* v, int i);
void* genOctreeInterface(void* levels, const double* xyz, int n);
scp_legacy_node* Nodes_get(void* level, int i);
int   Nodes_size(void* level);
int   int_size(void* codes);
int   int_get(void* codes, int i);

/* ------------------------------------------------------------------------------------------
 * Coding order + symbols (A7): encode.py:109-136 / encode_mullevel.py:106-133.
 * For every level (h_level_sizes, rows consecutive) and window of `context_size`: even ids then odd ids.
 * d_order[N] receives row indices; d_sym[N] (optional) the symbols occ-1 in that order.
 * ---------------------------------------------------------------------------------------- */
int scp_coding_order(const int64_t* h_level_sizes, const uint8_t* h_level_restart, int n_levels, int context_size,
                     int add_base_for_single, const uint8_t* d_occ, int64_t* d_order, int16_t* d_sym, void* stream);
/* h_level_restart (optional): 1 where a level starts a new frame -- the reference restarts `coded_cnt` per frame,
 * which matters only for its single-node-level quirk (encode.py:123). */

/* Context-window assembly (encode.py:112-115 + the odd-length pad token of ehem.py:92-99) for ALL windows of a
 * batch: window w covers rows [h_win_row[w], h_win_row[w]+h_win_len[w]) and lands at token h_win_tok[w] of the
 * padded stream (every window padded to even length with ctx (0,0,255), pos 0).  Writes ctx [T,12], pos [T,3]
 * and, for the even / odd tokens, the row each logits row belongs to (-1 for the pad token). */
int scp_gather_windows(const uint8_t* d_ctx, const float* d_pos, const int64_t* h_win_row, const int32_t* h_win_len,
                       const int64_t* h_win_tok, int n_win, uint8_t* d_ctx_out, float* d_pos_out,
                       int64_t* d_row_even, int64_t* d_row_odd, void* stream);
/* OctAttention sequences (encode.py:23-82 with encode_dataset.py:31-55 / encode_dataset_mullevel.py:45-69): sequence i of the
 * output = `pad` pad rows ((level 0, octant 0, occupancy 255), zero positions, row_of -1) followed by h_len[i]-pad rows
 * of d_ctx [N,4,3] / d_ctx_pos [N,4,3] starting at row h_src_start[i], written at output row h_dst_start[i]; positions are
 * shifted left by h_shift[i] (= 21 - deepest level of the row file, so that pos / 2^max_level == out * 2^-21).
 * d_row_of (optional) [total] = source row of every output row. */
int scp_pad_gather_seqs(const uint8_t* d_ctx, const uint32_t* d_ctx_pos, const int64_t* h_dst_start, const int64_t* h_src_start,
                        const int32_t* h_len, const int32_t* h_shift, int n_seq, int pad, uint8_t* d_ctx_out,
                        uint32_t* d_pos_out, int64_t* d_row_of, void* stream);
/* out[i, :] = in[idx[i], :] for 8-byte rows (coding-order gather of the (c_low,c_high) intervals). */
int scp_gather_rows8(const void* d_in, const int64_t* d_idx, int64_t n, void* d_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * PMF / logits -> integer CDF (A13): torch.softmax (encode.py:126-127) +
 * numpyAc.py:109-114 pdf_convert_to_cdf_and_normalize + :80-107 _convert_to_int_and_normalize.
 * in: [n,255] float32 (logits if is_logits else PMF).  d_row_of[n] (optional) scatters row i of the
 * input to CDF row d_row_of[i] (negative = skip the row).  Outputs (each optional): d_cdf [n,256] uint16; d_interval [n,2] uint32
 * = (c_low, c_high) of symbol d_sym[row] with the coder's 0x10000 substitution (numpyAc_backend.cpp:271-277);
 * d_pmf [n,255] float32.
 * ---------------------------------------------------------------------------------------- */
int scp_pmf_to_cdf(const float* d_in, int64_t n, int is_logits, const int64_t* d_row_of,
                   const int16_t* d_sym, uint16_t* d_cdf, uint32_t* d_interval, float* d_pmf, void* stream);

/* ------------------------------------------------------------------------------------------
 * Range coder (A14, host): numpyAc_backend.cpp:245-323 `encode`.  Byte-identical output.
 * h_interval [n,2] uint32 (c_low,c_high) in coding order.  Returns the number of bytes, or <0.
 * If h_out is NULL only the size is computed.
 * ---------------------------------------------------------------------------------------- */
int64_t scp_range_encode(const uint32_t* h_interval, int64_t n, uint8_t* h_out, int64_t out_cap);
/* Same from a full CDF table like numpyAc_backend.encode_cdf(cdf[N,Lp], sym[N]) (:327-334). */
int64_t scp_range_encode_cdf(const uint16_t* h_cdf, const int16_t* h_sym, int64_t n, int Lp,
                             uint8_t* h_out, int64_t out_cap);

/* ------------------------------------------------------------------------------------------
 * Range decoder (row f-2 of SURVEY.md section 8, host): numpyAc_backend.cpp:134-229 `decode` / numpyAc.py:139-170
 * `arithmeticDeCoding`.  Stateful: the caller hands over the CDF rows (uint16 [n,Lp], the table scp_pmf_to_cdf
 * writes) in coding order, any number at a time, and gets the symbols back.
 * ---------------------------------------------------------------------------------------- */
typedef struct scp_range_decoder scp_range_decoder;
scp_range_decoder* scp_range_decoder_create(const uint8_t* h_bytes, int64_t n_bytes);
void    scp_range_decoder_destroy(scp_range_decoder* d);
int     scp_range_decode(scp_range_decoder* d, const uint16_t* h_cdf, int64_t n, int Lp, int16_t* h_sym);
int64_t scp_range_decoder_count(const scp_range_decoder* d);

/* Decode side, device: the decoder rebuilds every level from the occupancy bytes decoded so far (decode_ehem.py:100-140).
 * Node state of a level: d_pos int32 [n,3] cell origins, d_anc uint8 [n,3,3] (level, octant, occ-1) of the three ancestors
 * (missing = (0,0,255)), d_octant uint8 [n].
 * scp_decode_level_inputs: the context rows the entropy model is fed, as on the encode side -- d_ctx [n,4,3] with the self
 *   row (level, octant, 255), d_ctx_model = the same with the level column clipped to clip_level (encode_dataset_ehem.py:86;
 *   255 = no clip), d_pos_norm float32 [n,3] = float32((pos - pos_min) / pos_den) (encode_dataset_ehem.py:70-72).
 * scp_expand_children: d_occ uint8 [n] decoded occupancy bytes (0 = no children), d_ctx [n,4,3] of scp_decode_level_inputs;
 *   writes the next level's state in BFS order (parents in order, child digit ascending), sum(popcount(occ)) nodes;
 *   `level` = the parents' level, `cell` = the children's cell size. */
int scp_decode_level_inputs(const int32_t* d_pos, const uint8_t* d_anc, const uint8_t* d_octant, int64_t n, int level,
                            int clip_level, double pos_min, double pos_den, uint8_t* d_ctx, uint8_t* d_ctx_model,
                            float* d_pos_norm, void* stream);
int scp_expand_children(const uint8_t* d_occ, const int32_t* d_pos, const uint8_t* d_ctx, int64_t n, int level, int cell,
                        int32_t* d_child_pos, uint8_t* d_child_anc, uint8_t* d_child_octant, void* stream);

/* Distortion report.
 * scp_dequantise_keys: voxel Morton keys of a job (scp_octree_out.voxel_key; bit triple b = (x,y,z) at bits 3b+2,3b+1,3b)
 *   -> float64 [n,3] points `v * steps + offset` mapped back to Cartesian (spher2cart / cylin2cart,
 *   data_preprocess.py:179-229; what proc_pc :68-92 / mul_proc_pc :160-167 return as the quantised cloud).
 *   h_steps / h_offset: 3 host doubles each (scp_job_info.steps / .offset; offset zero for SCP_MODE_SPHER).
 * scp_nn_dist2: d_dist2[i] = min_j |query_i - cand_j|^2 in float64, exact (brute force) -- the two KDTree queries of
 *   pt.py:88-95 (distChamfer) and the nearest-neighbour pass of pc_error's D1 metric.  Points float64 [n,3]. */
int scp_dequantise_keys(const int64_t* d_keys, int64_t n, const double* h_steps, const double* h_offset, int mode,
                        double* d_xyz, void* stream);
int scp_nn_dist2(const double* d_query, int64_t n_query, const double* d_cand, int64_t n_cand, double* d_dist2, void* stream);

/* ------------------------------------------------------------------------------------------
 * Entropy-model operators (A8-A12).  Device pointers, float32 activations, row-major [tokens, channels].
 * Windows of ANY length are processed together as one ragged batch: a `scp_seqs` describes how the token
 * stream is cut into sequences (one per context window of encode.py:112-115).
 * ---------------------------------------------------------------------------------------- */
typedef struct scp_seqs scp_seqs;
/* h_offsets[n_seq+1]: token ranges of the sequences (host). Uploads the tables the kernels need. */
scp_seqs* scp_seqs_create(const int64_t* h_offsets, int n_seq);
/* Same, with the table upload ordered on `stream` (no host synchronisation; use the stream the operators run on). */
scp_seqs* scp_seqs_create_async(const int64_t* h_offsets, int n_seq, void* stream);
void      scp_seqs_destroy(scp_seqs* s);
int64_t   scp_seqs_total(const scp_seqs* s);

#define SCP_ACT_NONE 0
#define SCP_ACT_LEAKY001 1   /* nn.LeakyReLU() default slope 0.01 (ehem.py:36, dgcnn.py:94) */
#define SCP_ACT_GELU 2       /* exact erf GELU (swin_transformer.py:561, HF ACT2FN["gelu"]) */
#define SCP_ACT_RELU 3       /* oct_attention.py:26 */

#define SCP_GEMM_AUTO 0
#define SCP_GEMM_SIMT 1      /* fp32 FFMA tiles */
#define SCP_GEMM_TF32 2      /* tcgen05.mma kind::tf32, TMA-fed, TMEM accumulators (10-bit mantissa operands) */
#define SCP_GEMM_TF32X3 3    /* same engine, error-compensated split (x_hi + x_lo): fp32-class accuracy */
#define SCP_GEMM_F16X3 4     /* the same split on the fp16 pipe (kind::f16, twice the tf32 rate), weights scaled per matrix */

/* y[M,N] = act( x[M,K] @ W[N,K]^T + bias[N] ) (+ residual[M,N])   -- nn.Linear with fused epilogue.
 * bias, residual may be NULL.  ldx/ldy/ldr = row strides in floats (W is dense [N,K]). */
int scp_linear(const float* d_x, int64_t ldx, const float* d_w, const float* d_bias,
               const float* d_res, int64_t ldr, float* d_y, int64_t ldy,
               int64_t M, int N, int K, int act, int engine, void* stream);
/* What SCP_GEMM_AUTO means: 0 = fp32 SIMT everywhere, 1 = tcgen05 3xTF32 for the large layers. Returns the old value. */
int scp_set_auto_engine(int use_tf32);
/* CTAs per thread-block cluster of the K <= 256 tensor-core layers with many rows (1, 2 or 4; default 2, SCP_GEMM_CL): the CTAs of
   a cluster share the weight stream by TMA multicast.  Results do not depend on it (same MMAs, same order).  Returns the old value. */
int scp_set_gemm_cluster(int ctas);
/* Drops the cached hi/lo splits of weight matrices (call after weights changed in place). */
void scp_gemm_cache_clear(void);
/* Drops the cached splits of ONE weight matrix (keyed by its device pointer).  The cache cannot see contents: the host
 * side (scp_b200/ops.py) calls this whenever the tensor behind a pointer is a different object or was modified in place
 * since its split was taken. */
void scp_gemm_cache_drop(const float* d_w);
/* 1 if the tcgen05 engine can take this shape (alignment rules in DESIGN.md). */
int scp_linear_tf32_supported(int64_t ldx, int64_t ldy, int64_t M, int N, int K);

/* LayerNorm over the last dim: swin_transformer.py:591-593,340, attention_model.py:105-106.
 * Optional fused residual: y = LN(x + res) (attention_model.py:114-116). */
int scp_layernorm(const float* d_x, int64_t ldx, const float* d_res, int64_t ldr, const float* d_gamma,
                  const float* d_beta, float* d_y, int64_t ldy, int64_t M, int C, float eps, void* stream);

/* EHEM token embedding (dgcnn.py:122-129): ctx bytes [n,4,3] (level,octant,occ) -> [n,80]
 * = [occ_enc(occ of the 3 ancestors) 48 | level_enc(4) 16 | octant_enc(4) 16]. */
int scp_ehem_embed(const uint8_t* d_ctx, int64_t n, const float* d_occ_enc, const float* d_level_enc,
                   int n_level_rows, const float* d_octant_enc, float* d_out, int64_t ldo, void* stream);
/* occ_enc rows for pre_occ (dgcnn.py:153): out[i,:16] = occ_enc[ctx[2*i][3].occ]  (even tokens). */
int scp_ehem_embed_occ(const uint8_t* d_ctx, int64_t n_even, const float* d_occ_enc, float* d_out, int64_t ldo,
                       void* stream);

/* kNN (dgcnn.py:10-28): for every point the k nearest points (largest 2 xi.xj - |xj|^2 - |xi|^2, self
 * included) inside its own sequence.  x [total, d] (ldx).  idx [total, k] int32, GLOBAL row indices,
 * descending score; ties -> lower index first; sequences shorter than k repeat the point itself. */
int scp_knn(const float* d_x, int64_t ldx, int d, const scp_seqs* seqs, int k, int32_t* d_idx, void* stream);
/* d > 4: 1 (default) = tcgen05 3xTF32 Gram tiles + fused top-k, 0 = fp32 SIMT tiles.  Returns the old value. */
int scp_set_knn_engine(int use_tensor_cores);

/* Edge convolution (dgcnn.py:48-71 get_graph_feature + :79-87 conv/BN/LeakyReLU(0.2) + :134 max over k),
 * evaluated as max_k f(Wa x_nbr + (Wb-Wa) x_i):  uv [total, 2C] = x @ [Wa; Wb-Wa]^T comes from scp_linear,
 * then out[i,c] = lrelu02( s[c]*(sel_k uv[nbr_k, c] + uv[i, C+c]) + t[c] ), sel = max if s[c] >= 0 else min
 * (BatchNorm eval folded to s,t; monotone so the max commutes exactly). */
int scp_edge_gather_max(const float* d_uv, int64_t lduv, int C, const int32_t* d_idx, int k, int64_t n,
                        const float* d_bn_scale, const float* d_bn_shift, float* d_out, int64_t ldo, void* stream);
/* The same, writing the result into a second view as well (d_out2 may be NULL): dgcnn.py:118-150 concatenates every edge-conv
 * output twice (into the next layer's input and into the final feature), so the copy is made where the values are produced. */
int scp_edge_gather_max2(const float* d_uv, int64_t lduv, int C, const int32_t* d_idx, int k, int64_t n,
                         const float* d_bn_scale, const float* d_bn_shift, float* d_out, int64_t ldo, float* d_out2, int64_t ldo2,
                         void* stream);

/* 1-D shifted-window attention (swin_transformer.py:406-501 + :603-652,:684-697), heads x 64.
 * q [total, *] (ldq) queries, kv-side k,v [total, *] (ldk, ldv): projected tokens of every sequence.  Each
 * sequence of S tokens is logically zero-padded AFTER LayerNorm to Sp = ceil(S/512)*512 tokens, so padded
 * tokens carry exactly the Linear biases qb/kb/vb [heads*64] (what the reference computes); shift = 0 or 256
 * rolls by -shift before windowing and back afterwards; bias = relpos[(i-j)+511, head]; -100 shift mask on the
 * last window (:603-623).  out [total, heads*64] (ldo). */
int scp_swin_attention(const float* d_q, int64_t ldq, const float* d_k, int64_t ldk, const float* d_v, int64_t ldv,
                       const float* d_qb, const float* d_kb, const float* d_vb, const float* d_relpos, int heads,
                       const scp_seqs* seqs, int shift, float* d_out, int64_t ldo, void* stream);

/* 2 (default) = tcgen05 3xFP16 window attention, two CTAs per SM (attn_h.cu); 1 = tcgen05 3xTF32 (attn_tc.cu); 0 = fp32 SIMT
 * tiles.  Returns the old value. */
int scp_set_attn_engine(int mode);

/* Patch merging input (swin_transformer.py:350-362): per sequence, out[j] = [x[2j], x[2j+1]] (zeros if 2j+1 >= S),
 * j < ceil(S/2).  `dst` describes the halved sequences. */
int scp_pair_concat(const float* d_x, int64_t ldx, const scp_seqs* src, const scp_seqs* dst, int C,
                    float* d_out, int64_t ldo, void* stream);

/* EHEM.concat_states (ehem.py:72-86): out[t, col_off:col_off+C] = x[seq_start_src + ((t - seq_start_dst) >> shift)]
 * (nearest x2^shift upsample of a coarser stage, cropped). */
int scp_upsample_cols(const float* d_src, int64_t lds, const scp_seqs* src, const scp_seqs* dst, int shift, int C,
                      float* d_out, int64_t ldo, int col_off, void* stream);

/* Strided row copy: out[i, col_off:col_off+C] = src[i*row_step + row_off, :C], i < rows (even/odd split, concat). */
int scp_copy_cols(const float* d_src, int64_t lds, int64_t row_step, int64_t row_off, int64_t rows, int C,
                  float* d_out, int64_t ldo, int col_off, void* stream);

/* OctAttention embedding (oct_attention.py:52-79,85-99 + attention_model.py:20-22): both streams
 * embed / embed_unknown [n, 600] incl. *sqrt(600) and the sinusoidal PE (row = position inside the sequence).
 * ctx bytes are (level,octant,occ); ctx_pos u32 [n,4,3]; pos_scale = 1/2^max_level (encode_dataset.py:48).
 * d_pe may be NULL: cfg.model.pos_embed False, no PositionalEncoding module (attention_model.py:142-144,147-149). */
int scp_octattn_embed(const uint8_t* d_ctx, const uint32_t* d_ctx_pos, float pos_scale, int level_base,
                      int max_octree_level, const scp_seqs* seqs,
                      const float* d_occ_enc, const float* d_level_enc, const float* d_octant_enc,
                      const float* d_pos_w, const float* d_pos_b, const float* d_pe,
                      float* d_embed, float* d_embed_unknown, void* stream);

/* Two-stream causal attention of OctAttention (attention_model.py:58-95), heads x head_dim (4 x 150).
 * qu,k,ku,v,vu: [total, heads*head_dim] with row stride ld; out, out_u likewise (ldo). */
int scp_octattn_attention(const float* d_qu, const float* d_k, const float* d_ku, const float* d_v,
                          const float* d_vu, int64_t ld, int heads, int head_dim, const scp_seqs* seqs,
                          float* d_out, float* d_out_u, int64_t ldo, void* stream);

/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
int64_t scp_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SCP_B200_H */

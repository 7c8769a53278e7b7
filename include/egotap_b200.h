/* egotap_b200 -- C ABI of the B200-native EgoTAP heatmap->3D lifting path.
 *
 * The reference (tho-kn/EgoTAP) is pure Python/PyTorch and has no FFI of its own: the seam a
 * replacement plugs into is the factory `define_AutoEncoder(opt, model)`
 * (reference model/network.py:24-33) plus the nn.Module protocol of `EgoTAPAutoEncoder`
 * (reference model/net_architecture.py:579-758).  The Python module in egotap_b200/ mirrors that
 * surface and forwards every piece of arithmetic to the entry points below through ctypes
 * (INTEGRATION.md shows the binding).  Conventions:
 *   - every pointer is a DEVICE pointer unless the name says `host`; `stream` is a cudaStream_t
 *   - every function returns 0 on success, a negative EGOTAP_E_* for argument errors, or a
 *     positive cudaError_t; it never throws, never exits, and never synchronises unless stated
 *   - egotap_b200_last_error() returns a thread-local description of the last failure
 *   - there is no CPU fallback: without a CUDA device every compute entry fails
 */
#ifndef EGOTAP_B200_H
#define EGOTAP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EGOTAP_B200_ABI_VERSION 1

enum { EGOTAP_E_ARG = -1, EGOTAP_E_UNSUPPORTED = -2, EGOTAP_E_DRIVER = -3, EGOTAP_E_STATE = -4 };

/* presets (reference options/dataset_options.py:30-41, utils/util.py:51-52) */
enum { EGOTAP_PRESET_UNREALEGO = 0, EGOTAP_PRESET_EGOCAP = 1 };
/* operand precision of the tensor-core contractions (accumulation is always fp32) */
enum {
  EGOTAP_PREC_BF16X3 = 0, /* error-compensated bf16 hi/lo split, 3 MMAs: fp32-parity mode */
  EGOTAP_PREC_BF16 = 1    /* plain bf16 operands, 1 MMA: throughput mode, looser stated bound */
};

int egotap_b200_abi_version(void);
const char* egotap_b200_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches claim) */
long long egotap_b200_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Op level: one tcgen05 GEMM with fused epilogue.  D[g][m][n] = epi(sum_k A[g][m][k]*B[g][n][k]).
 * Replaces every F.linear / conv-as-GEMM / torch.matmul call site on the path
 * (reference model/modeling_vit.py:195,226-252,271,326,339; model/network_utils.py:123-142;
 *  model/custom_cells.py:99,104,107).
 * ------------------------------------------------------------------------------------------ */
typedef struct egotap_operand {
  const void* hi;      /* bf16 [g1][g0][rows][K], K contiguous */
  const void* lo;      /* bf16 residual x - hi, NULL in EGOTAP_PREC_BF16 */
  long long ld;        /* row stride, elements */
  long long rows;      /* rows per group (reads beyond are zero-filled) */
  long long g0_count, g0_stride, g1_count, g1_stride; /* group dims / strides, elements */
} egotap_operand;

enum { EGOTAP_ACT_NONE = 0, EGOTAP_ACT_GELU = 1, EGOTAP_ACT_LRELU = 2 };
enum { EGOTAP_STORE_ROWMAJOR = 0, EGOTAP_STORE_QKV = 1, EGOTAP_STORE_JOINT_REGROUP = 2, EGOTAP_STORE_HEAD_MERGE = 3 };

typedef struct egotap_epilogue {
  float alpha;            /* v = acc * alpha */
  const float* scale;     /* v *= scale[n]          (nullable) */
  const float* bias;      /* v += bias[n]           (nullable) */
  int act;                /* EGOTAP_ACT_* applied here */
  const float* resid;     /* v += resid[rrow][col]  (nullable) */
  long long resid_ld;
  int resid_mod;          /* rrow = resid_mod ? m % resid_mod : out row */
  int rows_in, rows_out;  /* out row = g*group_rows + (m/rows_in)*rows_out + m%rows_in  (rows_in==0: m) */
  long long group_rows;
  float* out_f32;         /* any subset of the outputs */
  void* out_hi;           /* bf16 */
  void* out_lo;           /* bf16 */
  long long ldo;
  int col_off;
  int store;              /* EGOTAP_STORE_* */
  int qk_cols, tokens;    /* STORE_QKV: cols >= qk_cols go transposed to vt_*[(frame*(N-qk_cols)+c)*tokens+tok] */
  void* vt_hi;
  void* vt_lo;
  int J;                  /* STORE_JOINT_REGROUP: m = frame*2J + view*J + j -> row frame*J + j, col += view*N */
  int heads;              /* STORE_HEAD_MERGE: g = frame*heads + h -> row frame*tokens + m, col += h*N */
} egotap_epilogue;

/* operand layout.  NT (default): A[g][m][k], B[g][n][k], K contiguous -- D = A B^T.
 * TN: A[k][m], B[k][n] row-major with the contraction along the ROWS and a group = a chunk of the contraction:
 *   D[g][m][n] = sum_{k < K} A[g K + k][m] B[g K + k][n]   (the weight-gradient GEMM dW = dY^T X from row-major activations;
 *   operand.rows = total rows of both matrices, reads beyond are zero-filled; group fields of the operands unused;
 *   N % 256 == 0, M % 64 == 0; fp32 / bf16 outputs through the same epilogue, one (M x N) block per group). */
enum { EGOTAP_GEMM_NT = 0, EGOTAP_GEMM_TN = 1 };

typedef struct egotap_gemm {
  egotap_operand a, b;
  int M, N, K, groups;    /* per-group extents; K % 64 == 0, N % 32 == 0 */
  int precision;          /* EGOTAP_PREC_* */
  int variant;            /* -1 = auto; else index into the compiled tile configurations */
  int layout;             /* EGOTAP_GEMM_* */
  egotap_epilogue epi;
} egotap_gemm;

int egotap_b200_gemm(const egotap_gemm* desc, void* stream);
/* number of compiled tile configurations and a printable name for each */
int egotap_b200_gemm_num_variants(void);
const char* egotap_b200_gemm_variant_name(int variant);

/* Op level: fused multi-head self-attention over 576 tokens, 8 heads x 128 (reference
 * model/modeling_vit.py:233-252).  qk: (frames*576, 2048) bf16 = [Q | K] per token, head h at columns h*128;
 * vt: (frames*8*128, 576) bf16 = V transposed per (frame, head) -- both as written by the QKV GEMM's
 * EGOTAP_STORE_QKV epilogue; ctx: (frames*576, 1024) bf16.  lo parts NULL in EGOTAP_PREC_BF16. */
int egotap_b200_attention(const void* qk_hi, const void* qk_lo, const void* vt_hi, const void* vt_lo, void* ctx_hi,
                          void* ctx_lo, int frames, int precision, void* stream);

/* Op level: the bandwidth-bound kernels and the persistent propagation chain, for op-level parity tests.
 *
 * ingest: (frames, 6J, 64, 64) fp32 heatmaps -> patch-embedding operand (frames*2J*16, 256) and limb operand
 *   (frames*2J, 8192) as bf16 hi/lo (reference model/net_architecture.py:688-694, :375-383, modeling_vit.py:195).
 * layernorm: LayerNorm over 1024 features of rows [frame*rows_in + tok], tok < rows_out, written compacted to
 *   row frame*rows_out + tok as bf16 hi/lo and/or fp32 (reference model/modeling_vit.py:367,378,609).
 * pu_permute_split: W_hh (2048 x 512 fp32, gate order f,i,g,o) -> gate-permuted bf16 hi/lo for pu_chain.
 * pu_chain: one propagation-unit layer over all J joints (reference model/custom_cells.py:94-120,149-197):
 *   gates[b,t] = G[b*G_rs + t*G_ts + :2048] + (sigmoid(F[b*F_rs + t*F_ts + :512]) * h[b,t-1]) . W_hh^T,
 *   h -> out[(b*J + t)*512 + :] (fp32, optional bf16 hi/lo); hg: scratch 2*frames*512 bf16 (hi, lo);
 *   counters: >= ceil(frames/256) 32-bit words; frames <= 1024 per call.
 * head: per-joint Linear(768->3) on [e[b*J+j, :256] | skel[b*J+j, :512]], optional global Linear(J*512->6) whose
 *   first 3 outputs are added to every joint and last 3 appended as the LAST joint
 *   (reference model/net_architecture.py:732-751). */
int egotap_b200_ingest(const float* heatmaps, int frames, int preset, void* patch_hi, void* patch_lo, void* limb_hi,
                       void* limb_lo, void* stream);
int egotap_b200_layernorm(const float* x, const float* weight, const float* bias, long long frames, int rows_in,
                          int rows_out, float eps, void* out_hi, void* out_lo, float* out_f32, void* stream);
int egotap_b200_pu_permute_split(const float* w_hh, void* w_hi, void* w_lo, void* stream);
int egotap_b200_pu_chain(const void* w_hi, const void* w_lo, const float* G, long long G_rs, long long G_ts,
                         const float* F, long long F_rs, long long F_ts, float* out, void* out_hi, void* out_lo,
                         void* hg_hi, void* hg_lo, void* counters, int frames, int J, int precision, void* stream);
int egotap_b200_head(const float* e, int e_ld, const float* skel, const float* Wp, const float* bp, const float* Wg,
                     const float* bg, long long frames, int J, float* pose, void* stream);

/* Evaluation metrics: per-frame MPJPE and Procrustes-aligned MPJPE of pred vs gt poses
 * (frames, joints, 3), multiplied by unit_scale (10 = cm -> mm).  Replaces batch_compute_similarity_transform_torch
 * + the per-frame metric loop (reference utils/util.py:328-379, utils/evaluate.py:54-73,
 * model/egotap_autoencoder_model.py:336-348). */
int egotap_b200_pose_metrics(const float* pred, const float* gt, long long frames, int joints, float unit_scale,
                             float* mpjpe, float* pa_mpjpe, void* stream);

/* Per-launch CUDA-event timing of every kernel launched between begin and end (bench.py's live roofline
 * measurement; events are recorded on the launching stream).  end() synchronises the device.  GEMM records
 * carry their shape; other kernels report M = N = K = 0. */
int egotap_b200_profile_begin(void);
int egotap_b200_profile_end(int* num_records);
int egotap_b200_profile_record(int index, const char** name, int* M, int* N, int* K, int* groups, int* variant, float* ms);

/* fp32 -> bf16 hi/lo split of a contiguous array (operand preparation; lo may be NULL) */
int egotap_b200_split_bf16(const float* src, void* hi, void* lo, long long n, void* stream);

/* ------------------------------------------------------------------------------------------
 * Model level: the whole lifting path.  Replaces EgoTAPAutoEncoder.forward
 * (reference model/net_architecture.py:682-758; called from model/egotap_autoencoder_model.py:223).
 *
 *   1. egotap_b200_plan_sizes  -> bytes of the packed-weight buffer and of the workspace for a
 *      preset / precision / maximum batch; the caller allocates both on the device (torch does)
 *   2. egotap_b200_plan_create -> lays both buffers out (HBM layout in DESIGN.md)
 *   3. egotap_b200_pack_weights: fp32 state_dict tensors (device pointers, in the canonical order
 *      egotap_b200_param_name() enumerates; names are the reference's state_dict keys) ->
 *      bf16 hi/lo operand matrices, QKV stacked, BN folded, position embeddings permuted to the
 *      heatmap-major token order.  Re-run after every load_state_dict.
 *   4. egotap_b200_forward: heatmaps (B, 6J, 64, 64) fp32 -> pose (B, num_joints, 3) fp32, async on stream
 * ------------------------------------------------------------------------------------------ */
typedef struct egotap_plan egotap_plan;

int egotap_b200_num_params(int preset);
const char* egotap_b200_param_name(int preset, int index);
int egotap_b200_plan_sizes(int preset, int precision, int max_batch, size_t* packed_bytes, size_t* workspace_bytes);
int egotap_b200_plan_create(int preset, int precision, int max_batch, void* packed, void* workspace, egotap_plan** out);
int egotap_b200_plan_destroy(egotap_plan* plan);
int egotap_b200_pack_weights(egotap_plan* plan, const float* const* params, int num_params, void* stream);

/* stages at which egotap_b200_forward can stop (op-level parity taps); -1 or EGOTAP_STAGE_POSE = full path */
enum {
  EGOTAP_STAGE_EMBED = 0, EGOTAP_STAGE_LAYER0 = 1, EGOTAP_STAGE_LAYER1 = 2, EGOTAP_STAGE_LAYER2 = 3,
  EGOTAP_STAGE_VIT_OUT = 4, EGOTAP_STAGE_JOINT_EMBED = 5, EGOTAP_STAGE_LIMB_EMBED = 6, EGOTAP_STAGE_CHAIN = 7,
  EGOTAP_STAGE_POSE = 8
};
int egotap_b200_forward(egotap_plan* plan, const float* heatmaps, int batch, float* pose, int last_stage, void* stream);
/* device pointer of a named intermediate ("hidden", "fin_hi", "fin_lo", "embed", "h0", "skel") for the parity tests */
int egotap_b200_plan_buffer(egotap_plan* plan, const char* name, void** ptr);

/* ------------------------------------------------------------------------------------------
 * Training step: the ops egotap_b200/training.py composes, together with the
 * GEMM / attention / layernorm / ingest / head entries above, into the train-mode forward, the loss, the full
 * backward and AdamW.  Replaces what torch.autograd + torch.optim.AdamW do for net_AutoEncoder in the reference's
 * optimize_parameters (model/egotap_autoencoder_model.py:284-323, model/network.py:72-78, utils/loss.py:44-85,
 * model/network_utils.py:123-142 for train-mode BatchNorm1d).  Exact semantics of each entry: oracle/op_oracle.py.
 * `scratch` is a caller-provided fp32 device buffer of `scratch_elems` elements (8-byte aligned) for the
 * deterministic two-stage reductions.  Pointers named *_host are HOST arrays.
 * ------------------------------------------------------------------------------------------ */
int egotap_b200_zero(void* ptr, size_t bytes, void* stream);
int egotap_b200_copy(void* dst, const void* src, size_t bytes, void* stream);
int egotap_b200_add3(const float* a, const float* b, const float* c /*nullable*/, float* out, int n, void* stream);
/* weight / operand preparation (exports of the packing kernels egotap_b200_pack_weights uses) */
int egotap_b200_split2d(const float* src, long long rows, long long cols, long long src_ld, void* hi, void* lo,
                        long long dst_ld, void* stream);
int egotap_b200_fill_dummy(float* hidden, const float* dummy, int frames, int tokens, int live, void* stream);
int egotap_b200_pos_permute(const float* pos, const float* mask_token, int grid, int n_hm, float* pos_perm, float* dummy,
                            void* stream);
int egotap_b200_pu_bridge_gate(const float* f, int f_ld, int f_col, const float* e, int e_ld, int X, long long rows, void* hi,
                               void* lo, void* stream);
/* gradient preparation: fp32 (rows x cols) -> row-major bf16 pair rm[r][c] and / or transposed pair t[c][r]
 * (t[c][rows..pad_rows) = 0); logical row r is read from source row (r / rows_out) * rows_in + r % rows_out when
 * rows_out > 0.  gelu_u (nullable): values are first multiplied by gelu'(gelu_u[r][c]) (same layout as src).
 * colsum_out (nullable): also out[c] = sum_r value[r][c] (needs scratch of ceil(rows/64) * cols floats). */
int egotap_b200_transpose_split(const float* src, long long rows, int cols, long long src_ld, int rows_in, int rows_out,
                                void* rm_hi, void* rm_lo, long long rm_ld, void* t_hi, void* t_lo, long long t_ld,
                                long long pad_rows, const float* gelu_u, float* colsum_out, float* scratch,
                                long long scratch_elems, void* stream);
/* batched bf16 transpose d[g1][g0][c][r] = s[g1][g0][r][c], zero padded to pad_rows */
int egotap_b200_transpose_bf16(const void* s_hi, const void* s_lo, long long rows, int cols, long long s_ld, int g0_count,
                               long long s_g0_stride, int g1_count, long long s_g1_stride, void* d_hi, void* d_lo,
                               long long d_ld, long long d_g0_stride, long long d_g1_stride, long long pad_rows, void* stream);
/* out[c] = sum_r src[row(r)][c]  (bias gradients; per-token sums over frames) */
int egotap_b200_colsum(const float* src, long long rows, int cols, long long ld, int rows_in, int rows_out, float* out,
                       float* scratch, long long scratch_elems, void* stream);
/* out[i] = sum_g partials[g*n + i]  (split-K partial products of the dW GEMMs) */
int egotap_b200_reduce_partials(const float* partials, int G, long long n, float* out, void* stream);
int egotap_b200_gelu_fwd(const float* u, long long n, void* out_hi, void* out_lo, void* stream);
int egotap_b200_gelu_bwd(float* dg /*in place*/, const float* u, long long n, void* stream);
int egotap_b200_layernorm_bwd(const float* dy, const float* x, const float* w, long long frames, int rows_in, int rows_out,
                              float eps, float* dx, int accumulate, float* dw, float* db, float* scratch,
                              long long scratch_elems, void* stream);
/* P = softmax(S) ; dS = P * (dP - rowsum(P * dP)) * scale ; S, dP fp32 (rows x 576) */
int egotap_b200_softmax_bwd(const float* S, const float* dP, long long rows, int cols, float scale, void* p_hi, void* p_lo,
                            void* ds_hi, void* ds_lo, void* stream);
/* Fused attention backward of the bf16-operand training mode (csrc/attention_bwd.cu; autograd of reference
 * model/modeling_vit.py:233-252).  attention_lse = egotap_b200_attention that also writes the per-row log-sum-exp in the exp2
 * domain, lse[(frame*8 + head)*576 + token] = max * c + log2(sum), c = log2(e)/sqrt(128).  attn_dsum: D = rowsum(dctx o ctx)
 * per (frame, head, token), same indexing (rows = frames*576).  attention_bwd: dqkv (frames*576, 3072) fp32 =
 * [dQ | dK | dV], head h at columns h*128 of each 1024-column third; scores and probabilities never leave the SM. */
int egotap_b200_attention_lse(const void* qk_hi, const void* qk_lo, const void* vt_hi, const void* vt_lo, void* ctx_hi,
                              void* ctx_lo, float* lse, int frames, int precision, void* stream);
int egotap_b200_attn_dsum(const void* ctx_hi, const void* ctx_lo, const void* dctx_hi, const void* dctx_lo, long long rows,
                          float* dsum, void* stream);
int egotap_b200_attention_bwd(const void* qk_hi, const void* vt_hi, const void* dctx_hi, const float* lse, const float* dsum,
                              float* dqkv, int frames, void* stream);
/* train-mode BatchNorm1d: batch statistics + running-buffer update + folded scale/shift; apply with LeakyReLU(0.2)
 * (optionally into the per-joint [left | right] layout); backward through LeakyReLU and the batch statistics */
int egotap_b200_bn_stats(const float* y, long long rows, int cols, const float* gamma, const float* beta, float* running_mean,
                         float* running_var, long long* num_batches_tracked /*nullable, device*/, float momentum, float eps,
                         float* mean, float* rstd, float* scale, float* shift, float* scratch, long long scratch_elems,
                         void* stream);
int egotap_b200_bn_apply(const float* y, long long rows, int cols, const float* scale, const float* shift, void* out_hi,
                         void* out_lo, long long out_ld, float* out_f32, long long f32_ld, int J, int col_off, void* stream);
int egotap_b200_bn_bwd(float* da /*in place -> dy*/, const float* y, long long rows, int cols, const float* scale,
                       const float* shift, const float* mean, const float* rstd, float* dgamma, float* dbeta, float* scratch,
                       long long scratch_elems, void* stream);
int egotap_b200_regroup_gather(const float* dE, long long e_ld, int col_off, long long frames, int J, int cols, float* out,
                               void* stream);
/* propagation-unit cell, one joint step, with the cell / hidden state of every step kept for the backward
 * (reference model/custom_cells.py:94-120), and its backward (BPTT, t = J-1 .. 0) */
int egotap_b200_pu_cell_fwd(const float* G, long long g_rs, long long g_ts, const float* F, long long f_rs, long long f_ts,
                            float* C, float* H, void* h_hi, void* h_lo, void* hg_hi, void* hg_lo, int t, int J,
                            long long frames, void* stream);
int egotap_b200_pu_cell_bwd(const float* G, long long g_rs, long long g_ts, const float* F, long long f_rs, long long f_ts,
                            const float* C, const float* H, const float* dOut, const float* dhg, float* dc, float* dG,
                            long long dg_rs, long long dg_ts, float* dF, long long df_rs, long long df_ts, void* dgp_hi,
                            void* dgp_lo, int t, int J, long long frames, void* stream);
/* the same BPTT as J x (pu_cell_bwd + dgates . W_hh GEMM) in ONE persistent launch (mirror of egotap_b200_pu_chain):
 * wT = W_hh^T (512 x 2048) bf16 hi/lo; x = exchange scratch [2][frames][2048] bf16 hi/lo; counters >= ceil(frames/256)
 * words; frames <= 1024 per call */
int egotap_b200_pu_chain_bwd(const void* wT_hi, const void* wT_lo, const float* G, long long g_rs, long long g_ts,
                             const float* F, long long f_rs, long long f_ts, const float* C, const float* H, const float* dOut,
                             float* dG, long long dg_rs, long long dg_ts, float* dF, long long df_rs, long long df_ts,
                             void* x_hi, void* x_lo, void* counters, int frames, int J, int precision, void* stream);
int egotap_b200_pu_bridge_gate_bwd(float* dE, long long e_ld, const float* F0, long long f_ld, int f_col, const float* E, int X,
                                   long long rows, float* dF, long long df_ld, void* stream);
int egotap_b200_head_bwd(const float* dpose, const float* e, long long e_ld, const float* skel, const float* Wp,
                         const float* Wg /*nullable*/, long long frames, int J, float* dE, long long de_ld, float* dSkel,
                         float* dWp, float* dbp, float* dWg, float* dbg, float* scratch, long long scratch_elems, void* stream);
int egotap_b200_embed_grads(const float* dpos_perm, int grid, int n_hm, float* dpos, float* dmask, void* stream);
/* loss[0] = total, [1] = lambda_mpjpe * MPJPE, [2] = lambda_cos * lambda_mpjpe * bone cosine; dpose = d total / d pred */
int egotap_b200_pose_loss(const float* pred, const float* gt, long long frames, int joints, const int* parents_host,
                          int n_parents, int drop_first, float lambda_mpjpe, float lambda_cos, float* loss, float* dpose,
                          float* scratch, long long scratch_elems, void* stream);
/* torch.optim.AdamW on `count` tensors; gradients are multiplied by grad_scale first (1 / world size after a SUM
 * all-reduce of the data-parallel gradients) */
int egotap_b200_adamw(float* const* params_host, const float* const* grads_host, float* const* m_host, float* const* v_host,
                      const long long* numel_host, int count, int step, double lr, double beta1, double beta2, double eps,
                      double weight_decay, double grad_scale, void* stream);

/* Ground-truth heatmap synthesis: keypoints -> the lifting input (frames, 6J, 64, 64) =
 * [joint L | joint R | cos L | sin L | cos R | sin R], as the reference builds it on the CPU under --use_gt_heatmap
 * (utils/projection.py:263-279, utils/data.py:175-252, dataloader/data_loader.py:127-132,193-199,
 * model/egotap_autoencoder_model.py:176-213).  pts2d: (frames, 2 views, J+1, 2) in 1024-pixel image coordinates;
 * pts3d_left: (frames, J+1, 3) = local pose + left pelvis (the limb elevation angle of BOTH views comes from it). */
int egotap_b200_gt_heatmaps(const float* pts2d, const float* pts3d_left, long long frames, int preset, float* out,
                            void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EGOTAP_B200_H */

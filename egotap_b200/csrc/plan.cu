// The lifting-path plan: packed weights, workspace layout and the forward pass
// (one call = reference EgoTAPAutoEncoder.forward(pose_only=True), model/net_architecture.py:682-751).
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "gemm.cuh"
#include "host_util.cuh"
#include "internal.h"

namespace eb {

// kernels.cu
int split2d_run(const float*, long long, long long, long long, __nv_bfloat16*, __nv_bfloat16*, long long, cudaStream_t);
int ingest_run(const float*, int, int, __nv_bfloat16*, __nv_bfloat16*, __nv_bfloat16*, __nv_bfloat16*, cudaStream_t);
int fill_dummy_run(float*, const float*, int, int, int, cudaStream_t, __nv_bfloat16* = nullptr, __nv_bfloat16* = nullptr, float2* = nullptr);
int ln_fold_pack_run(const float*, const float*, const float*, const float*, int, __nv_bfloat16*, __nv_bfloat16*, float*, float*, cudaStream_t);
int layernorm_run(const float*, const float*, const float*, long long, int, int, float, __nv_bfloat16*, __nv_bfloat16*,
                  float*, cudaStream_t);
int softmax_run(const float*, long long, int, __nv_bfloat16*, __nv_bfloat16*, cudaStream_t);
int pu_bridge_gate_run(const float*, int, int, const float*, int, int, long long, __nv_bfloat16*, __nv_bfloat16*,
                       cudaStream_t);
int pu_cell_run(const float*, long long, float*, const float*, int, int, int, int, long long, float*, __nv_bfloat16*,
                __nv_bfloat16*, __nv_bfloat16*, __nv_bfloat16*, cudaStream_t);
int head_run(const float*, int, const float*, const float*, const float*, const float*, const float*, long long, int, int,
             int, float*, cudaStream_t);
int bn_fold_run(const float*, const float*, const float*, const float*, const float*, int, float*, float*, cudaStream_t);
int vec_add3_run(const float*, const float*, const float*, float*, int, cudaStream_t);
int pos_permute_run(const float*, const float*, int, int, float*, float*, cudaStream_t);
// train_ops.cu (shared with the training step)
int reduce_partials_run(const float*, int, long long, float*, cudaStream_t);
int bn_apply_run(const float*, long long, int, const float*, const float*, __nv_bfloat16*, __nv_bfloat16*, long long, float*,
                 long long, int, int, cudaStream_t);

constexpr int HID = 1024, HEADS = 8, HDIM = 128, TOK = 576, MLP = 4096, EMB = 128, PUH = 512, PUX = 256;
constexpr int NLAYERS = 3;

static bool fused_attention();
static bool small_batch_splitk();
static bool layernorm_fold();
constexpr int SPLITK_MAX_FRAMES = 32, SPLITK_MAX_G = 8;

struct W2 {  // bf16 hi/lo weight matrix [N][K]
  __nv_bfloat16* hi = nullptr;
  __nv_bfloat16* lo = nullptr;
};

struct Bump {  // bump allocator over a caller-provided buffer (base == nullptr: size query only)
  uint8_t* base;
  size_t off = 0;
  explicit Bump(void* b) : base(static_cast<uint8_t*>(b)) {}
  template <class T>
  T* take(size_t n) {
    off = (off + 255) & ~size_t(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

// canonical parameter order of egotap_b200_pack_weights (names = reference state_dict keys)
static std::vector<std::string> param_names(int preset) {
  std::vector<std::string> n;
  const std::string v = "pos_heatmap_encoder.vit.";
  n.push_back(v + "embeddings.mask_token");
  n.push_back(v + "embeddings.position_embeddings");
  n.push_back(v + "embeddings.patch_embeddings.projection.weight");
  n.push_back(v + "embeddings.patch_embeddings.projection.bias");
  for (int l = 0; l < NLAYERS; ++l) {
    const std::string p = v + "encoder.layer." + std::to_string(l) + ".";
    for (const char* s : {"attention.attention.query", "attention.attention.key", "attention.attention.value",
                          "attention.output.dense", "intermediate.dense", "output.dense"}) {
      n.push_back(p + s + ".weight");
      n.push_back(p + s + ".bias");
    }
    for (const char* s : {"layernorm_before", "layernorm_after"}) {
      n.push_back(p + s + ".weight");
      n.push_back(p + s + ".bias");
    }
  }
  n.push_back(v + "layernorm.weight");
  n.push_back(v + "layernorm.bias");
  for (const char* enc : {"pos_heatmap_encoder.", "rot_heatmap_encoder."})
    for (const char* fc : {"fc1.", "fc2.", "fc3."})
      for (const char* s : {"fc.weight", "fc.bias", "bn.weight", "bn.bias", "bn.running_mean", "bn.running_var"})
        n.push_back(std::string(enc) + fc + s);
  const std::string pu = "skel_sequential_layer.lstm_custom.layers.";
  for (const char* s : {"0.x2f", "0.x2h", "0.b2h", "0.h2h", "1.x2f", "1.x2h", "1.h2h"}) {
    n.push_back(pu + s + ".weight");
    n.push_back(pu + s + ".bias");
  }
  n.push_back("pose_mlp.pose_fcs.0.weight");
  n.push_back("pose_mlp.pose_fcs.0.bias");
  if (preset == EGOTAP_PRESET_UNREALEGO) {
    n.push_back("global_mlp.pose_fcs.0.weight");
    n.push_back("global_mlp.pose_fcs.0.bias");
  }
  return n;
}

struct Plan {
  int preset, precision, nsplit, max_batch;
  int J, n_hm, grid, live, nj;
  bool global_head, packed_ok = false, fused_attention_layout = true, splitk_layout = false, ln_fold = false;
  std::vector<std::string> names;
  // ---- packed weights
  W2 w_patch;
  float *b_patch, *pos_perm, *dummy;
  struct Layer {
    W2 qkv, o, up, down;
    float *b_qkv, *b_o, *b_up, *b_down, *ln1w, *ln1b, *ln2w, *ln2b;
    float *s_qkv, *c_qkv, *s_up, *c_up;     // LayerNorm fold: row sums of the gamma-scaled weights, bias + beta . W
  } L[NLAYERS];
  float *lnfw, *lnfb;
  struct FC {
    W2 w;
    float *scale, *shift;
    int n, k;
  } pfc[3], rfc[3];
  W2 x2f0, xb0, hh0, cat1, hh1, hh0p, hh1p;   // hh*p: gate-permuted copies for the persistent chain kernel
  float *b_x2f0, *b_g0, *b_cat1;
  float *Wp, *bp, *Wg, *bg;
  size_t packed_bytes;
  // ---- workspace
  W2 a_patch, a_limb, ln, qk, vt, P, ctx, mlp, fin, f1, f2, xb, hg, h0b;
  float *hidden, *S, *E, *F0, *G0, *gates, *cst, *H0, *FG1, *skel, *skp = nullptr, *sky = nullptr;
  float2* stats;                            // LayerNorm fold: (sum, sum of squares) per token row and 128-column part
  unsigned int* counters;
  size_t workspace_bytes;

  W2 take2(Bump& b, size_t n) {
    W2 w;
    w.hi = b.take<__nv_bfloat16>(n);
    w.lo = nsplit == 3 ? b.take<__nv_bfloat16>(n) : nullptr;
    return w;
  }

  void layout(void* packed, void* workspace) {
    Bump p(packed);
    w_patch = take2(p, size_t(HID) * 256);
    b_patch = p.take<float>(HID);
    pos_perm = p.take<float>(size_t(TOK) * HID);
    dummy = p.take<float>(size_t(TOK - live) * HID + 4);
    for (auto& l : L) {
      l.qkv = take2(p, size_t(3 * HID) * HID);
      l.o = take2(p, size_t(HID) * HID);
      l.up = take2(p, size_t(MLP) * HID);
      l.down = take2(p, size_t(HID) * MLP);
      l.b_qkv = p.take<float>(3 * HID);
      l.b_o = p.take<float>(HID);
      l.b_up = p.take<float>(MLP);
      l.b_down = p.take<float>(HID);
      l.ln1w = p.take<float>(HID); l.ln1b = p.take<float>(HID);
      l.ln2w = p.take<float>(HID); l.ln2b = p.take<float>(HID);
      l.s_qkv = p.take<float>(3 * HID); l.c_qkv = p.take<float>(3 * HID);
      l.s_up = p.take<float>(MLP); l.c_up = p.take<float>(MLP);
    }
    lnfw = p.take<float>(HID); lnfb = p.take<float>(HID);
    const int dims[3][2] = {{2048, 0}, {512, 2048}, {EMB, 512}};
    for (int e = 0; e < 2; ++e)
      for (int i = 0; i < 3; ++i) {
        FC& f = e == 0 ? pfc[i] : rfc[i];
        f.n = dims[i][0];
        f.k = i == 0 ? (e == 0 ? 16 * HID : 2 * 64 * 64) : dims[i][1];
        f.w = take2(p, size_t(f.n) * f.k);
        f.scale = p.take<float>(f.n);
        f.shift = p.take<float>(f.n);
      }
    x2f0 = take2(p, size_t(PUH + PUX) * PUX);
    xb0 = take2(p, size_t(4 * PUH) * (2 * PUX));
    hh0 = take2(p, size_t(4 * PUH) * PUH);
    cat1 = take2(p, size_t(5 * PUH) * PUH);
    hh1 = take2(p, size_t(4 * PUH) * PUH);
    hh0p = take2(p, size_t(4 * PUH) * PUH);
    hh1p = take2(p, size_t(4 * PUH) * PUH);
    b_x2f0 = p.take<float>(PUH + PUX);
    b_g0 = p.take<float>(4 * PUH);
    b_cat1 = p.take<float>(5 * PUH);
    Wp = p.take<float>(3 * (PUX + PUH)); bp = p.take<float>(4);
    Wg = p.take<float>(size_t(6) * J * PUH); bg = p.take<float>(8);
    packed_bytes = p.off + 256;

    Bump w(workspace);
    const size_t B = size_t(max_batch);
    a_patch = take2(w, B * live * 256);
    a_limb = take2(w, B * n_hm * 8192);
    hidden = w.take<float>(B * TOK * HID);
    ln = take2(w, B * TOK * HID);
    stats = w.take<float2>(B * TOK * (HID / 128));
    qk = take2(w, B * TOK * 2 * HID);
    vt = take2(w, B * TOK * HID);
    if (!fused_attention_layout) {  // score / probability scratch of the unfused A/B path only
      S = w.take<float>(B * HEADS * TOK * TOK);
      P = take2(w, B * HEADS * TOK * TOK);
    } else {
      S = nullptr;
      P = W2();
    }
    ctx = take2(w, B * TOK * HID);
    mlp = take2(w, B * TOK * MLP);
    fin = take2(w, B * n_hm * 16 * HID);
    f1 = take2(w, B * n_hm * 2048);
    f2 = take2(w, B * n_hm * 512);
    E = w.take<float>(B * J * 2 * PUX);
    xb = take2(w, B * J * 2 * PUX);
    F0 = w.take<float>(B * J * (PUH + PUX));
    G0 = w.take<float>(B * J * 4 * PUH);
    gates = w.take<float>(B * 4 * PUH);
    cst = w.take<float>(B * PUH);
    hg = take2(w, 2 * B * PUH);
    counters = w.take<unsigned int>(64);
    H0 = w.take<float>(B * J * PUH);
    h0b = take2(w, B * J * PUH);
    FG1 = w.take<float>(B * J * 5 * PUH);
    skel = w.take<float>(B * J * PUH);
    if (splitk_layout) {   // split-K partial products / sums of the first FC block at small batches (default; EGOTAP_SPLITK=0 disables)
      const size_t Bs = B < size_t(SPLITK_MAX_FRAMES) ? B : size_t(SPLITK_MAX_FRAMES);
      skp = w.take<float>(Bs * n_hm * 2048 * SPLITK_MAX_G);
      sky = w.take<float>(Bs * n_hm * 2048);
    }
    workspace_bytes = w.off + 256;
  }
};

static int plan_init(Plan& pl, int preset, int precision, int max_batch) {
  EB_REQUIRE(preset == EGOTAP_PRESET_UNREALEGO || preset == EGOTAP_PRESET_EGOCAP, "unknown preset %d", preset);
  EB_REQUIRE(precision == EGOTAP_PREC_BF16X3 || precision == EGOTAP_PREC_BF16, "unknown precision %d", precision);
  EB_REQUIRE(max_batch > 0, "max_batch must be positive");
  pl.preset = preset;
  pl.precision = precision;
  pl.nsplit = precision == EGOTAP_PREC_BF16X3 ? 3 : 1;
  pl.max_batch = max_batch;
  pl.J = preset == EGOTAP_PRESET_UNREALEGO ? 15 : 17;
  pl.global_head = preset == EGOTAP_PRESET_UNREALEGO;
  pl.nj = pl.global_head ? pl.J + 1 : pl.J;
  pl.n_hm = 2 * pl.J;
  pl.grid = 6;  // int(sqrt(n_hm - 1)) + 1 for n_hm = 30 and 34 (reference model/net_architecture.py:328)
  pl.live = pl.n_hm * 16;
  pl.names = param_names(preset);
  pl.fused_attention_layout = fused_attention();
  pl.splitk_layout = small_batch_splitk();
  pl.ln_fold = layernorm_fold();
  return 0;
}

// LayerNorm fold (opt-in: EGOTAP_LN=fold; the default keeps the six in-layer LayerNorm passes).  LN(x) W^T + b =
// rstd (x (gamma o W)^T - mean s) + c with s[n] = sum_k gamma_k W[n][k], c[n] = b[n] + sum_k beta_k W[n][k]: the GEMM that
// writes the residual stream (patch embedding, out-projection, MLP-down) also stores the row as the bf16 (hi / lo) operand and
// its (sum, sum of squares) per 128 columns; the GEMM that consumed the LayerNorm output (QKV, MLP-up) multiplies the
// un-normalised row with the gamma-scaled weights and applies mean / rstd per row in its epilogue.  Operand precision is
// unchanged (the row is split exactly as the normalised row was; the mean term is removed with the row sums of the ROUNDED
// weights): measured on the B200, pose error vs the oracle 8.4e-5 folded vs 8.6e-5 (parity mode), ViT stages 1.8-2.0e-5 both.
// Why it is not the default (profiles/r02i_layernorm_fold.md): it removes 1.1 ms of HBM-roofline kernels from a 31 ms step
// (the sum of the kernel times drops 29.6 -> 28.2 ms) and the STEP does not get faster (31.4 vs 31.3 ms, three interleaved
// pairs; bf16 mode 13.43 vs 13.54): the step runs against the 1000 W cap, the LayerNorm passes are its low-power phases, and
// without them the clock governor settles lower (1.33 -> 1.29 GHz).  Energy per frame, not kernel time, bounds this step.
static bool layernorm_fold() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("EGOTAP_LN");
    v = (e && std::string(e) == "fold") ? 1 : 0;
  }
  return v == 1;
}

// EGOTAP_ATTN=unfused selects the three-kernel attention (score GEMM, softmax, context GEMM) kept for A/B checks
static bool fused_attention() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("EGOTAP_ATTN");
    v = (e && std::string(e) == "unfused") ? 0 : 1;
  }
  return v == 1;
}

// At small batches (<= 32 frames) the first FC block of both encoders (K = 16384 / 8192, only B*2J rows) is a handful of
// tiles; its reduction dimension is cut into G groups so that >= 128 CTAs stream the weight matrix, the partial products are
// summed and the folded BatchNorm + LeakyReLU applied in a separate bandwidth-bound pass.  Measured on the B200
// (profiles/r02_small_batch.json): batch 1 1.59 -> 1.28 ms, batch 8 1.99 -> 1.69, batch 32 4.55 -> 4.43.  EGOTAP_SPLITK=0
// turns it off (A/B check).
static bool small_batch_splitk() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("EGOTAP_SPLITK");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// EGOTAP_SKIP_DUMMY=0 computes the last layer's dummy-token rows too (A/B check; results are identical)
static bool skip_dummy_rows() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("EGOTAP_SKIP_DUMMY");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// EGOTAP_PU=steps selects the per-joint launch sequence (h2h GEMM + cell kernel per step) kept for A/B checks
static bool persistent_chain() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("EGOTAP_PU");
    v = (e && std::string(e) == "steps") ? 0 : 1;
  }
  return v == 1;
}

static GemmOperand opnd(const W2& w, long long ld, long long rows) { return GemmOperand{w.hi, w.lo, ld, rows, 1, 0, 1, 0}; }

static EpiParams epi0() {
  EpiParams e;
  memset(&e, 0, sizeof(e));
  e.alpha = 1.0f;
  return e;
}

// y = epi(A[M,K] . W[N,K]^T) for plain (ungrouped) operands
static int linear(const Plan& pl, const W2& a, long long lda, int M, int K, const W2& w, int N, const EpiParams& e,
                  cudaStream_t st) {
  GemmShape s{M, N, K, 1, 1, 0};
  return gemm_run(opnd(a, lda, M), opnd(w, K, N), s, e, pl.nsplit, -1, st);
}

// Same, but only over the first `live` of every frame's 576 token rows (the dummy/mask tokens of the last layer
// are attended to as keys/values but their own rows are never read again: SURVEY 0.7).  One group per frame.
static int linear_live(const Plan& pl, const W2& a, long long lda, int B, int live, int K, const W2& w, int N,
                       EpiParams e, cudaStream_t st) {
  GemmOperand A{a.hi, a.lo, lda, live, B, (long long)TOK * lda, 1, 0};
  GemmOperand Wt{w.hi, w.lo, K, N, 1, 0, 1, 0};      // weights: every group multiplies the same matrix
  e.group_rows = TOK;
  GemmShape s{live, N, K, B, B, 1};
  return gemm_run(A, Wt, s, e, pl.nsplit, -1, st);
}

// LayerNorm fold (layernorm_fold() above): the producer also stores the bf16 operand and the row statistics, the consumer
// applies them
static void produce_ln_operand(const Plan& pl, EpiParams& e) {
  e.out_hi = pl.ln.hi; e.out_lo = pl.ln.lo;
  e.stats_out = pl.stats; e.stats_parts = HID / 128;
}
static void consume_ln_operand(const Plan& pl, EpiParams& e, const float* s, const float* c) {
  e.scale = s; e.bias = c;
  e.stats_in = pl.stats; e.stats_parts = HID / 128; e.ln_cols = HID; e.ln_eps = 1e-12f;
}

static int pack(Plan& pl, const float* const* P, int n, cudaStream_t st) {
  EB_REQUIRE(n == int(pl.names.size()), "pack_weights: expected %d parameter pointers, got %d", int(pl.names.size()), n);
  for (int i = 0; i < n; ++i) EB_REQUIRE(P[i] != nullptr, "pack_weights: parameter %d (%s) is null", i, pl.names[i].c_str());
  int rc, i = 0;
#define RC(x) do { if ((rc = (x))) return rc; } while (0)
#define COPYF(dst, src, cnt) EB_CUDA(cudaMemcpyAsync(dst, src, size_t(cnt) * 4, cudaMemcpyDeviceToDevice, st))
  const float* mask = P[i++];
  const float* pos = P[i++];
  RC(pos_permute_run(pos, mask, pl.grid, pl.n_hm, pl.pos_perm, pl.dummy, st));
  RC(split2d_run(P[i++], HID, 256, 256, pl.w_patch.hi, pl.w_patch.lo, 256, st));
  COPYF(pl.b_patch, P[i++], HID);
  for (auto& l : pl.L) {
    // canonical order inside a layer: q, k, v, attention-output, MLP-up, MLP-down (weight, bias each), then the two LayerNorms
    const float* const* lp = P + i;
    const float *ln1w = lp[12], *ln1b = lp[13], *ln2w = lp[14], *ln2b = lp[15];
    for (int q = 0; q < 3; ++q) {  // query, key, value stacked along N
      __nv_bfloat16* hi = l.qkv.hi + size_t(q) * HID * HID;
      __nv_bfloat16* lo = l.qkv.lo ? l.qkv.lo + size_t(q) * HID * HID : nullptr;
      if (pl.ln_fold) RC(ln_fold_pack_run(P[i], P[i + 1], ln1w, ln1b, HID, hi, lo, l.s_qkv + q * HID, l.c_qkv + q * HID, st));
      else RC(split2d_run(P[i], HID, HID, HID, hi, lo, HID, st));
      COPYF(l.b_qkv + q * HID, P[i + 1], HID);
      i += 2;
    }
    RC(split2d_run(P[i++], HID, HID, HID, l.o.hi, l.o.lo, HID, st));
    COPYF(l.b_o, P[i++], HID);
    if (pl.ln_fold) RC(ln_fold_pack_run(P[i], P[i + 1], ln2w, ln2b, MLP, l.up.hi, l.up.lo, l.s_up, l.c_up, st));
    else RC(split2d_run(P[i], MLP, HID, HID, l.up.hi, l.up.lo, HID, st));
    ++i;
    COPYF(l.b_up, P[i++], MLP);
    RC(split2d_run(P[i++], HID, MLP, MLP, l.down.hi, l.down.lo, MLP, st));
    COPYF(l.b_down, P[i++], HID);
    COPYF(l.ln1w, P[i++], HID); COPYF(l.ln1b, P[i++], HID);
    COPYF(l.ln2w, P[i++], HID); COPYF(l.ln2b, P[i++], HID);
  }
  COPYF(pl.lnfw, P[i++], HID); COPYF(pl.lnfb, P[i++], HID);
  for (int e = 0; e < 2; ++e)
    for (int k = 0; k < 3; ++k) {
      Plan::FC& f = e == 0 ? pl.pfc[k] : pl.rfc[k];
      RC(split2d_run(P[i], f.n, f.k, f.k, f.w.hi, f.w.lo, f.k, st));
      RC(bn_fold_run(P[i + 1], P[i + 2], P[i + 3], P[i + 4], P[i + 5], f.n, f.scale, f.shift, st));
      i += 6;
    }
  // propagation unit (reference model/custom_cells.py:80-86)
  const float *x2f0 = P[i], *bx2f0 = P[i + 1], *x2h0 = P[i + 2], *bx2h0 = P[i + 3], *b2h0 = P[i + 4], *bb2h0 = P[i + 5],
              *h2h0 = P[i + 6], *bh2h0 = P[i + 7], *x2f1 = P[i + 8], *bx2f1 = P[i + 9], *x2h1 = P[i + 10],
              *bx2h1 = P[i + 11], *h2h1 = P[i + 12], *bh2h1 = P[i + 13];
  i += 14;
  RC(split2d_run(x2f0, PUH + PUX, PUX, PUX, pl.x2f0.hi, pl.x2f0.lo, PUX, st));
  COPYF(pl.b_x2f0, bx2f0, PUH + PUX);
  // layer-0 x-side + bridge-side gate projections share one GEMM: K-concatenated [x2h | b2h]
  RC(split2d_run(x2h0, 4 * PUH, PUX, PUX, pl.xb0.hi, pl.xb0.lo, 2 * PUX, st));
  RC(split2d_run(b2h0, 4 * PUH, PUX, PUX, pl.xb0.hi + PUX, pl.xb0.lo ? pl.xb0.lo + PUX : nullptr, 2 * PUX, st));
  RC(vec_add3_run(bx2h0, bb2h0, bh2h0, pl.b_g0, 4 * PUH, st));
  RC(split2d_run(h2h0, 4 * PUH, PUH, PUH, pl.hh0.hi, pl.hh0.lo, PUH, st));
  RC(pu_permute_split_run(h2h0, pl.hh0p.hi, pl.hh0p.lo, st));
  // layer-1 x-side: N-concatenated [x2f ; x2h]
  RC(split2d_run(x2f1, PUH, PUH, PUH, pl.cat1.hi, pl.cat1.lo, PUH, st));
  RC(split2d_run(x2h1, 4 * PUH, PUH, PUH, pl.cat1.hi + size_t(PUH) * PUH,
                 pl.cat1.lo ? pl.cat1.lo + size_t(PUH) * PUH : nullptr, PUH, st));
  COPYF(pl.b_cat1, bx2f1, PUH);
  RC(vec_add3_run(bx2h1, bh2h1, nullptr, pl.b_cat1 + PUH, 4 * PUH, st));
  RC(split2d_run(h2h1, 4 * PUH, PUH, PUH, pl.hh1.hi, pl.hh1.lo, PUH, st));
  RC(pu_permute_split_run(h2h1, pl.hh1p.hi, pl.hh1p.lo, st));
  COPYF(pl.Wp, P[i++], 3 * (PUX + PUH));
  COPYF(pl.bp, P[i++], 3);
  if (pl.global_head) {
    COPYF(pl.Wg, P[i++], 6 * pl.J * PUH);
    COPYF(pl.bg, P[i++], 6);
  }
#undef COPYF
  pl.packed_ok = true;
  return 0;
}

// stages of the forward pass, for op-level parity taps (egotap_b200_forward's `last_stage`)
enum { ST_EMBED = 0, ST_LAYER0 = 1, ST_LAYER1 = 2, ST_LAYER2 = 3, ST_VIT_OUT = 4, ST_JOINT_EMBED = 5, ST_LIMB_EMBED = 6,
       ST_CHAIN = 7, ST_POSE = 8 };

static int forward(Plan& pl, const float* x, int B, float* pose, int last_stage, cudaStream_t st) {
  EB_REQUIRE(pl.packed_ok, "forward: weights have not been packed");
  EB_REQUIRE(x && pose, "forward: null pointer");
  EB_REQUIRE(B > 0 && B <= pl.max_batch, "forward: batch %d outside (0, %d]", B, pl.max_batch);
  int rc;
  const int J = pl.J, n_hm = pl.n_hm, live = pl.live;
  // ---- K1: ingest
  RC(ingest_run(x, B, J, pl.a_patch.hi, pl.a_patch.lo, pl.a_limb.hi, pl.a_limb.lo, st));
  // ---- K2+K3: patch embedding GEMM (+bias +permuted pos-emb), dummy rows
  {
    EpiParams e = epi0();
    e.bias = pl.b_patch;
    e.resid = pl.pos_perm; e.resid_ld = HID; e.resid_mod = live;
    e.rows_in = live; e.rows_out = TOK;
    e.out_f32 = pl.hidden; e.ldo = HID;
    if (pl.ln_fold) produce_ln_operand(pl, e);
    RC(linear(pl, pl.a_patch, 256, B * live, 256, pl.w_patch, HID, e, st));
    if (pl.ln_fold) RC(fill_dummy_run(pl.hidden, pl.dummy, B, TOK, live, st, pl.ln.hi, pl.ln.lo, pl.stats));
    else RC(fill_dummy_run(pl.hidden, pl.dummy, B, TOK, live, st));
  }
  if (last_stage == ST_EMBED) return 0;
  const int M = B * TOK;
  for (int l = 0; l < NLAYERS; ++l) {
    Plan::Layer& L = pl.L[l];
    if (!pl.ln_fold) RC(layernorm_run(pl.hidden, L.ln1w, L.ln1b, B, TOK, TOK, 1e-12f, pl.ln.hi, pl.ln.lo, nullptr, st));
    {  // K5: fused QKV projection; Q|K row-major, V transposed per (frame, head)
      EpiParams e = epi0();
      e.bias = L.b_qkv;
      if (pl.ln_fold) consume_ln_operand(pl, e, L.s_qkv, L.c_qkv);
      e.store = STORE_QKV; e.qk_cols = 2 * HID; e.tokens = TOK;
      e.out_hi = pl.qk.hi; e.out_lo = pl.qk.lo; e.ldo = 2 * HID;
      e.vt_hi = pl.vt.hi; e.vt_lo = pl.vt.lo;
      RC(linear(pl, pl.ln, HID, M, HID, L.qkv, 3 * HID, e, st));
    }
    if (fused_attention()) {
      // K6: fused softmax attention, scores stay in TMEM / shared memory
      // in the last layer only the live tokens need a context row (their dummy-row successors are never read)
      const int qrows = (l == NLAYERS - 1 && skip_dummy_rows()) ? live : TOK;
      RC(attention_run(pl.qk.hi, pl.qk.lo, pl.vt.hi, pl.vt.lo, pl.ctx.hi, pl.ctx.lo, B, pl.nsplit, qrows, st));
    } else {
      {  // K6a: scores S[b,h] = Q K^T / sqrt(128)
        EpiParams e = epi0();
        e.alpha = 0.08838834764831845f;
        e.out_f32 = pl.S; e.ldo = TOK; e.group_rows = TOK;
        GemmOperand q{pl.qk.hi, pl.qk.lo, 2 * HID, TOK, HEADS, HDIM, B, (long long)TOK * 2 * HID};
        GemmOperand k{pl.qk.hi + HID, pl.qk.lo ? pl.qk.lo + HID : nullptr, 2 * HID, TOK, HEADS, HDIM, B,
                      (long long)TOK * 2 * HID};
        GemmShape s{TOK, TOK, HDIM, B * HEADS, HEADS, 0};
        RC(gemm_run(q, k, s, e, pl.nsplit, -1, st));
      }
      RC(softmax_run(pl.S, (long long)B * HEADS * TOK, TOK, pl.P.hi, pl.P.lo, st));
      {  // K6b: context = P V, heads merged back to (B*576, 1024)
        EpiParams e = epi0();
        e.store = STORE_HEAD_MERGE; e.heads = HEADS; e.tokens = TOK;
        e.out_hi = pl.ctx.hi; e.out_lo = pl.ctx.lo; e.ldo = HID;
        GemmOperand p{pl.P.hi, pl.P.lo, TOK, TOK, (long long)B * HEADS, (long long)TOK * TOK, 1, 0};
        GemmOperand v{pl.vt.hi, pl.vt.lo, TOK, HDIM, (long long)B * HEADS, (long long)HDIM * TOK, 1, 0};
        GemmShape s{TOK, HDIM, TOK, B * HEADS, B * HEADS, 0};
        RC(gemm_run(p, v, s, e, pl.nsplit, -1, st));
      }
    }
    // worthwhile only when the per-frame tiling of the live rows issues fewer MMA rows than the dense tiling
    const bool last = (l == NLAYERS - 1) && skip_dummy_rows() && ((live + 255) / 256) * 256 < TOK;
    {  // K7: output projection + residual (in place on the fp32 residual stream)
      EpiParams e = epi0();
      e.bias = L.b_o; e.resid = pl.hidden; e.resid_ld = HID; e.out_f32 = pl.hidden; e.ldo = HID;
      if (pl.ln_fold) produce_ln_operand(pl, e);
      if (last) RC(linear_live(pl, pl.ctx, HID, B, live, HID, L.o, HID, e, st));
      else RC(linear(pl, pl.ctx, HID, M, HID, L.o, HID, e, st));
    }
    if (!pl.ln_fold) RC(layernorm_run(pl.hidden, L.ln2w, L.ln2b, B, TOK, TOK, 1e-12f, pl.ln.hi, pl.ln.lo, nullptr, st));
    {  // K8: MLP up + exact GELU
      EpiParams e = epi0();
      e.bias = L.b_up; e.act = ACT_GELU;
      if (pl.ln_fold) consume_ln_operand(pl, e, L.s_up, L.c_up); e.out_hi = pl.mlp.hi; e.out_lo = pl.mlp.lo; e.ldo = MLP;
      if (last) RC(linear_live(pl, pl.ln, HID, B, live, HID, L.up, MLP, e, st));
      else RC(linear(pl, pl.ln, HID, M, HID, L.up, MLP, e, st));
    }
    {  // K9: MLP down + residual
      EpiParams e = epi0();
      e.bias = L.b_down; e.resid = pl.hidden; e.resid_ld = HID; e.out_f32 = pl.hidden; e.ldo = HID;
      if (pl.ln_fold && l + 1 < NLAYERS) produce_ln_operand(pl, e);     // the last layer feeds the final LayerNorm kernel
      if (last) RC(linear_live(pl, pl.mlp, MLP, B, live, MLP, L.down, HID, e, st));
      else RC(linear(pl, pl.mlp, MLP, M, MLP, L.down, HID, e, st));
    }
    if (last_stage == ST_LAYER0 + l) return 0;
  }
  // ---- final LayerNorm, live tokens only, compacted to (B*n_hm, 16*1024) = per-heatmap 4x4 patch blocks
  RC(layernorm_run(pl.hidden, pl.lnfw, pl.lnfb, B, TOK, live, 1e-12f, pl.fin.hi, pl.fin.lo, nullptr, st));
  if (last_stage == ST_VIT_OUT) return 0;
  // ---- K11 / K12: the two FC encoders (Linear + folded BN + LeakyReLU), K13 regroup fused in fc3's store
  for (int enc = 0; enc < 2; ++enc) {
    Plan::FC* f = enc == 0 ? pl.pfc : pl.rfc;
    const W2& a0 = enc == 0 ? pl.fin : pl.a_limb;
    const int R = B * n_hm;
    EpiParams e = epi0();
    int G = 1;
    if (pl.splitk_layout && B <= SPLITK_MAX_FRAMES) {
      const int tiles = ((R + 127) / 128) * (2048 / 128);
      G = (148 + tiles - 1) / tiles;
      if (G > SPLITK_MAX_G) G = SPLITK_MAX_G;
      while (G > 1 && f[0].k % (64 * G) != 0) --G;
    }
    if (G > 1) {
      const int Kc = f[0].k / G;
      GemmOperand A{a0.hi, a0.lo, f[0].k, R, G, Kc, 1, 0};
      GemmOperand Wt{f[0].w.hi, f[0].w.lo, f[0].k, 2048, G, Kc, 1, 0};
      e.out_f32 = pl.skp; e.ldo = 2048; e.group_rows = R;
      GemmShape s{R, 2048, Kc, G, G, 0};
      RC(gemm_run(A, Wt, s, e, pl.nsplit, -1, st));
      RC(reduce_partials_run(pl.skp, G, (long long)R * 2048, pl.sky, st));
      RC(bn_apply_run(pl.sky, R, 2048, f[0].scale, f[0].shift, pl.f1.hi, pl.f1.lo, 2048, nullptr, 0, 0, 0, st));
      e = epi0();
    } else {
      e.scale = f[0].scale; e.bias = f[0].shift; e.act = ACT_LRELU; e.out_hi = pl.f1.hi; e.out_lo = pl.f1.lo; e.ldo = 2048;
      RC(linear(pl, a0, f[0].k, R, f[0].k, f[0].w, 2048, e, st));
    }
    e.act = ACT_LRELU;
    e.scale = f[1].scale; e.bias = f[1].shift; e.out_hi = pl.f2.hi; e.out_lo = pl.f2.lo; e.ldo = 512;
    RC(linear(pl, pl.f1, 2048, R, 2048, f[1].w, 512, e, st));
    e.scale = f[2].scale; e.bias = f[2].shift;
    e.store = STORE_JOINT_REGROUP; e.J = J; e.ldo = 2 * PUX;
    e.out_f32 = pl.E; e.col_off = enc == 0 ? 0 : PUX;
    e.out_hi = enc == 0 ? pl.xb.hi : nullptr; e.out_lo = enc == 0 ? pl.xb.lo : nullptr;
    RC(linear(pl, pl.f2, 512, R, 512, f[2].w, EMB, e, st));
    if (last_stage == ST_JOINT_EMBED + enc) return 0;
  }
  // ---- K14: propagation chain (chain semantics: joint t continues from joint t-1)
  const int R = B * J;
  {
    EpiParams e = epi0();
    e.bias = pl.b_x2f0; e.out_f32 = pl.F0; e.ldo = PUH + PUX;
    RC(linear(pl, pl.xb, 2 * PUX, R, PUX, pl.x2f0, PUH + PUX, e, st));
    RC(pu_bridge_gate_run(pl.F0, PUH + PUX, PUH, pl.E, 2 * PUX, PUX, R, pl.xb.hi, pl.xb.lo, st));
    e = epi0();
    e.bias = pl.b_g0; e.out_f32 = pl.G0; e.ldo = 4 * PUH;
    RC(linear(pl, pl.xb, 2 * PUX, R, 2 * PUX, pl.xb0, 4 * PUH, e, st));
  }
  for (int layer = 0; layer < 2; ++layer) {
    const float* G = layer == 0 ? pl.G0 : pl.FG1 + PUH;
    const int G_ld = layer == 0 ? 4 * PUH : 5 * PUH;
    const float* F = layer == 0 ? pl.F0 : pl.FG1;
    const int F_ld = layer == 0 ? PUH + PUX : 5 * PUH;
    const W2& hh = layer == 0 ? pl.hh0 : pl.hh1;
    float* out = layer == 0 ? pl.H0 : pl.skel;
    if (persistent_chain()) {
      // one persistent launch per layer (per chunk of up to 1024 frames = 128 co-resident CTAs)
      const W2& hp = layer == 0 ? pl.hh0p : pl.hh1p;
      for (int b0 = 0; b0 < B; b0 += 1024) {
        const int bc = B - b0 < 1024 ? B - b0 : 1024;
        RC(pu_chain_run(hp.hi, hp.lo, G + (long long)b0 * J * G_ld, (long long)J * G_ld, G_ld,
                        F + (long long)b0 * J * F_ld, (long long)J * F_ld, F_ld, out + (long long)b0 * J * PUH,
                        layer == 0 ? pl.h0b.hi + (long long)b0 * J * PUH : nullptr,
                        layer == 0 && pl.h0b.lo ? pl.h0b.lo + (long long)b0 * J * PUH : nullptr, pl.hg.hi, pl.hg.lo,
                        pl.counters, bc, J, pl.nsplit, st));
      }
    } else {
    EB_CUDA(cudaMemsetAsync(pl.cst, 0, size_t(B) * PUH * 4, st));
    for (int t = 0; t < J; ++t) {
      const float* gates = G + (long long)t * G_ld;   // t == 0: h = 0, so gates are the x-side term alone
      long long gates_ld = (long long)J * G_ld;
      if (t > 0) {
        EpiParams e = epi0();
        e.resid = G + (long long)t * G_ld; e.resid_ld = (long long)J * G_ld;
        e.out_f32 = pl.gates; e.ldo = 4 * PUH;
        RC(linear(pl, pl.hg, PUH, B, PUH, hh, 4 * PUH, e, st));
        gates = pl.gates; gates_ld = 4 * PUH;
      }
      RC(pu_cell_run(gates, gates_ld, pl.cst, F, F_ld, t, J, PUH, B, out, layer == 0 ? pl.h0b.hi : nullptr,
                     layer == 0 ? pl.h0b.lo : nullptr, pl.hg.hi, pl.hg.lo, st));
    }
    }
    if (layer == 0) {
      EpiParams e = epi0();
      e.bias = pl.b_cat1; e.out_f32 = pl.FG1; e.ldo = 5 * PUH;
      RC(linear(pl, pl.h0b, PUH, R, PUH, pl.cat1, 5 * PUH, e, st));
    }
  }
  if (last_stage == ST_CHAIN) return 0;
  // ---- K15: regression head (+ global offset, head joint last)
  RC(head_run(pl.E, 2 * PUX, pl.skel, pl.Wp, pl.bp, pl.global_head ? pl.Wg : nullptr, pl.global_head ? pl.bg : nullptr, B, J,
              PUX, PUH, pose, st));
  return 0;
#undef RC
}

}  // namespace eb

using namespace eb;

extern "C" int egotap_b200_num_params(int preset) {
  if (preset != EGOTAP_PRESET_UNREALEGO && preset != EGOTAP_PRESET_EGOCAP) return fail(EGOTAP_E_ARG, "unknown preset");
  return int(param_names(preset).size());
}
extern "C" const char* egotap_b200_param_name(int preset, int index) {
  static thread_local std::string s;
  if (preset != EGOTAP_PRESET_UNREALEGO && preset != EGOTAP_PRESET_EGOCAP) return "";
  auto n = param_names(preset);
  if (index < 0 || index >= int(n.size())) return "";
  s = n[index];
  return s.c_str();
}
extern "C" int egotap_b200_plan_sizes(int preset, int precision, int max_batch, size_t* packed_bytes,
                                      size_t* workspace_bytes) {
  Plan pl;
  int rc = plan_init(pl, preset, precision, max_batch);
  if (rc) return rc;
  pl.layout(nullptr, nullptr);
  if (packed_bytes) *packed_bytes = pl.packed_bytes;
  if (workspace_bytes) *workspace_bytes = pl.workspace_bytes;
  return 0;
}
extern "C" int egotap_b200_plan_create(int preset, int precision, int max_batch, void* packed, void* workspace,
                                       egotap_plan** out) {
  if (!out || !packed || !workspace) return fail(EGOTAP_E_ARG, "plan_create: null pointer");
  Plan* pl = new Plan();
  int rc = plan_init(*pl, preset, precision, max_batch);
  if (rc) { delete pl; return rc; }
  pl->layout(packed, workspace);
  *out = reinterpret_cast<egotap_plan*>(pl);
  return 0;
}
extern "C" int egotap_b200_plan_destroy(egotap_plan* p) {
  delete reinterpret_cast<Plan*>(p);
  return 0;
}
extern "C" int egotap_b200_pack_weights(egotap_plan* p, const float* const* params, int n, void* stream) {
  if (!p || !params) return fail(EGOTAP_E_ARG, "pack_weights: null pointer");
  return pack(*reinterpret_cast<Plan*>(p), params, n, (cudaStream_t)stream);
}
extern "C" int egotap_b200_forward(egotap_plan* p, const float* heatmaps, int batch, float* pose, int last_stage,
                                   void* stream) {
  if (!p) return fail(EGOTAP_E_ARG, "forward: null plan");
  return forward(*reinterpret_cast<Plan*>(p), heatmaps, batch, pose, last_stage < 0 ? ST_POSE : last_stage,
                 (cudaStream_t)stream);
}
extern "C" int egotap_b200_plan_buffer(egotap_plan* p, const char* name, void** ptr) {
  if (!p || !name || !ptr) return fail(EGOTAP_E_ARG, "plan_buffer: null pointer");
  Plan& pl = *reinterpret_cast<Plan*>(p);
  const std::string n(name);
  void* r = nullptr;
  if (n == "hidden") r = pl.hidden;
  else if (n == "fin_hi") r = pl.fin.hi;
  else if (n == "fin_lo") r = pl.fin.lo;
  else if (n == "embed") r = pl.E;
  else if (n == "skel") r = pl.skel;
  else if (n == "h0") r = pl.H0;
  else return fail(EGOTAP_E_ARG, "plan_buffer: unknown buffer '%s'", name);
  *ptr = r;
  return 0;
}

// ---- op-level entry points (parity tests) ----------------------------------------------------------------
extern "C" int egotap_b200_ingest(const float* x, int frames, int preset, void* p_hi, void* p_lo, void* l_hi, void* l_lo,
                                  void* stream) {
  if (preset != EGOTAP_PRESET_UNREALEGO && preset != EGOTAP_PRESET_EGOCAP) return fail(EGOTAP_E_ARG, "unknown preset");
  return ingest_run(x, frames, preset == EGOTAP_PRESET_UNREALEGO ? 15 : 17, (__nv_bfloat16*)p_hi, (__nv_bfloat16*)p_lo,
                    (__nv_bfloat16*)l_hi, (__nv_bfloat16*)l_lo, (cudaStream_t)stream);
}
extern "C" int egotap_b200_layernorm(const float* x, const float* w, const float* b, long long frames, int rows_in,
                                     int rows_out, float eps, void* hi, void* lo, float* out_f32, void* stream) {
  return layernorm_run(x, w, b, frames, rows_in, rows_out, eps, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, out_f32,
                       (cudaStream_t)stream);
}
extern "C" int egotap_b200_pu_permute_split(const float* w, void* hi, void* lo, void* stream) {
  if (!w || !hi) return fail(EGOTAP_E_ARG, "pu_permute_split: null pointer");
  return pu_permute_split_run(w, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, (cudaStream_t)stream);
}
extern "C" int egotap_b200_pu_chain(const void* w_hi, const void* w_lo, const float* G, long long G_rs, long long G_ts,
                                    const float* F, long long F_rs, long long F_ts, float* out, void* out_hi, void* out_lo,
                                    void* hg_hi, void* hg_lo, void* counters, int frames, int J, int precision,
                                    void* stream) {
  return pu_chain_run((const __nv_bfloat16*)w_hi, (const __nv_bfloat16*)w_lo, G, G_rs, G_ts, F, F_rs, F_ts, out,
                      (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, (__nv_bfloat16*)hg_hi, (__nv_bfloat16*)hg_lo,
                      (unsigned int*)counters, frames, J, precision == EGOTAP_PREC_BF16 ? 1 : 3, (cudaStream_t)stream);
}
extern "C" int egotap_b200_head(const float* e, int e_ld, const float* skel, const float* Wp, const float* bp,
                                const float* Wg, const float* bg, long long frames, int J, float* pose, void* stream) {
  return head_run(e, e_ld, skel, Wp, bp, Wg, bg, frames, J, PUX, PUH, pose, (cudaStream_t)stream);
}

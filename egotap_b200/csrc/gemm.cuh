// Persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[g][m][n] = epilogue( sum_k A[g][m][k] * B[g][n][k] )        A, B: bf16, K-major; accumulate fp32 in TMEM
//
// * operands arrive by TMA (4-D tiled maps, 128-byte swizzle) into a multi-stage shared-memory ring
// * one elected thread issues tcgen05.mma; CG = 2 pairs two SMs on one 256 x BN tile (cta_group::2)
// * NSPLIT = 3 is the error-compensated mode: A = Ah + Al, B = Bh + Bl (bf16 hi/lo splits) and
//   D = Ah*Bh + Ah*Bl + Al*Bh, which restores ~16 mantissa bits per operand (fp32-parity mode);
//   NSPLIT = 1 is the plain bf16-operand mode
// * accumulators are double-buffered in TMEM so the epilogue of tile i overlaps the main loop of tile i+1
// * the epilogue (bias / scale-shift / GELU / LeakyReLU / residual / re-layout / hi-lo split) is fused
//
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2.. = epilogue (4 warps for BN = 128,
// 8 for BN = 256: warp w may only touch TMEM lanes [32*(w%4), 32*(w%4)+32), so each group of 4 consecutive
// warps covers all 128 lanes and the two groups split the tile's columns).
#pragma once
#include "ptx.cuh"

namespace eb {

enum : int { ACT_NONE = 0, ACT_GELU = 1, ACT_LRELU = 2 };
enum : int { STORE_ROWMAJOR = 0, STORE_QKV = 1, STORE_JOINT_REGROUP = 2, STORE_HEAD_MERGE = 3 };

struct EpiParams {
  float alpha;              // v = acc * alpha
  const float* scale;       // v *= scale[n]            (nullable)
  const float* bias;        // v += bias[n]             (nullable)
  int act;                  // ACT_*
  const float* resid;       // v += resid[rrow][n]      (nullable); rrow = resid_mod ? m % resid_mod : out row
  long long resid_ld;
  int resid_mod;
  int rows_in, rows_out;    // out row = g*group_rows + (m / rows_in) * rows_out + m % rows_in  (rows_in == 0: m)
  long long group_rows;
  float* out_f32;           // any subset of the three outputs may be set
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
  long long ldo;
  int col_off;
  int store;                // STORE_*
  // STORE_QKV: columns [0, qk_cols) row-major as above; columns [qk_cols, N) are V and are written transposed
  // per (frame, head): vt[((frame*heads + h)*128 + d) * tokens + tok]
  int qk_cols, tokens;
  __nv_bfloat16* vt_hi;
  __nv_bfloat16* vt_lo;
  // STORE_JOINT_REGROUP: m = frame*2J + view*J + j  ->  out row = frame*J + j, column += view*N
  int J;
  // STORE_HEAD_MERGE: group g = frame*heads + h  ->  out row = frame*tokens + m, column += h*N
  int heads;
  // LayerNorm folded into the GEMMs around it (plan.cu, "LayerNorm fold"):
  //  * producer (the GEMM that writes the residual stream): stats_out[out_row * stats_parts + c / 128] = (sum, sum of squares) of
  //    the values it stores in columns [128 (c / 128), +128) of that row -- one entry per epilogue warp and tile
  //  * consumer (the GEMM whose A operand is the un-normalised row, its weights pre-multiplied by gamma):
  //    v = (acc * alpha - mean_r * scale[n]) * rstd_r + bias[n], mean / rstd from stats_in[out_row * stats_parts + 0 .. parts)
  //    over ln_cols columns (scale[n] = sum_k gamma_k W[n][k], bias[n] = b[n] + sum_k beta_k W[n][k])
  float2* stats_out;
  const float2* stats_in;
  int stats_parts, ln_cols;
  float ln_eps;
};

// Tensor maps of the TMA epilogue (TEPI): the fp32 output and the fp32 residual / additive table, described as bf16 matrices
// of twice the width (a 64 x 32 box of bf16 = 32 fp32 columns x 32 rows = one warp's accumulator chunk).
struct EpiMaps {
  CUtensorMap out, res;
};

struct GemmShape {
  int M, N, K;      // per group
  int groups;       // tiles enumerate (g, m_blk, n_blk), n fastest
  int gdiv;         // TMA coords: c2 = g % gdiv, c3 = g / gdiv
  int b_shared;     // 1: every group multiplies the same B matrix (B's group coordinates stay 0)
};
// TN = true (template): the operands are ROW-major with the CONTRACTION along the rows -- A[k][m], B[k][n], D = A^T B -- and a
// group is a chunk of the contraction: D[g][m][n] = sum_{k < K} A[g K + k][m] B[g K + k][n].  This is the weight-gradient GEMM
// dW = dY^T X straight from the row-major activations and gradients of the training step (no transposed copies): the TMA
// boxes are 64 contraction rows x 64 columns and the MMA reads them through MN-major shared-memory descriptors
// (canonical layout ((8,8,m),(8,k)) : ((1,8,LBO),(64,SBO)) in bf16 elements: LBO = one 8 KB box, SBO = 1 KB, 2 KB per k-step).

// Per-thread row mapping of the epilogue (computed once per tile): output row, residual row, column shift.
struct EpiRow {
  long long orow, rrow;
  int col_shift;
};
__device__ __forceinline__ EpiRow epi_row(const EpiParams& p, int g, int m, int N) {
  EpiRow r;
  r.col_shift = 0;
  if (p.store == STORE_JOINT_REGROUP) {
    const int frame = m / (2 * p.J), rem = m % (2 * p.J);
    const int view = rem / p.J, j = rem % p.J;
    r.orow = (long long)frame * p.J + j;
    r.col_shift = view * N;
  } else if (p.store == STORE_HEAD_MERGE) {
    r.orow = (long long)(g / p.heads) * p.tokens + m;
    r.col_shift = (g % p.heads) * N;
  } else {
    r.orow = (p.rows_in > 0) ? (long long)(m / p.rows_in) * p.rows_out + (m % p.rows_in) : (long long)m;
    r.orow += (long long)g * p.group_rows;
  }
  r.rrow = p.resid_mod > 0 ? (long long)(m % p.resid_mod) : r.orow;
  return r;
}
// Residual / additive-table chunk of 32 columns; issued BEFORE the accumulator chunk is waited for so the
// (row-strided, DRAM-latency) loads overlap the TMEM read and the previous chunk's math.
__device__ __forceinline__ void epi_load_resid(const EpiParams& p, const EpiRow& row, int n0, float4 (&t)[8], bool wide) {
  const float* r = p.resid + row.rrow * p.resid_ld + n0 + p.col_off + row.col_shift;
  if (wide) {            // 32-byte loads: 4 instead of 8 LSU instructions per 32-column chunk of a row
#pragma unroll
    for (int j = 0; j < 4; ++j) ld_global_256(r + 8 * j, t[2 * j], t[2 * j + 1]);
  } else {
    const float4* r4 = reinterpret_cast<const float4*>(r);
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = r4[j];
  }
}
// all row starts of the residual / fp32 / bf16 outputs are 32-byte aligned (warp-uniform, evaluated once per kernel)
__device__ __forceinline__ bool epi_wide_ok(const EpiParams& p) {
  auto ok = [](const void* ptr, long long ld_bytes) {
    return ptr == nullptr || (((reinterpret_cast<uintptr_t>(ptr)) | uintptr_t(ld_bytes)) & 31) == 0;
  };
  return ok(p.resid, p.resid_ld * 4) && ok(p.out_f32, p.ldo * 4) && ok(p.out_hi, p.ldo * 2) && ok(p.out_lo, p.ldo * 2);
}

// Per-column vectors of the epilogue (folded scale, bias) for the 32 columns of a chunk.  They are staged once per tile in
// shared memory by the epilogue threads (one column each, requested BEFORE the accumulator wait) and read back as warp-wide
// broadcasts: the first version fetched them with __ldg inside the chunk loop, and the first use of every chunk sat on that
// L2 round trip (ncu, bf16 mode: 13-17 % of all samples of the QKV / MLP-up GEMMs on the first bias FADD of a chunk).
struct EpiLn {             // per-thread (= per-row) LayerNorm statistics of the consumer form; rstd == 0: not folded
  float mean, rstd;
};
__device__ __forceinline__ EpiLn epi_ln_row(const EpiParams& p, long long orow, bool row_ok) {
  EpiLn ln{0.f, 0.f};
  if (p.stats_in != nullptr && row_ok) {
    const float2* st = p.stats_in + orow * p.stats_parts;
    float sum = 0.f, sq = 0.f;
    for (int i = 0; i < p.stats_parts; ++i) { const float2 t = __ldg(st + i); sum += t.x; sq += t.y; }
    const float inv_n = 1.0f / float(p.ln_cols);
    ln.mean = sum * inv_n;
    const float var = fmaxf(sq * inv_n - ln.mean * ln.mean, 0.f);
    ln.rstd = 1.0f / sqrtf(var + p.ln_eps);
  }
  return ln;
}
__device__ __forceinline__ void epi_scale_bias(const EpiParams& p, const float* sv_scale, const float* sv_bias, float (&v)[32],
                                               const EpiLn ln = EpiLn{0.f, 0.f}) {
  if (p.stats_in != nullptr) {     // LayerNorm fold: (acc - mean * s[n]) * rstd + c[n]
    const float4* s4 = reinterpret_cast<const float4*>(sv_scale);
    const float4* b4 = reinterpret_cast<const float4*>(sv_bias);
    const float nm = -ln.mean;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 s = s4[j], b = b4[j];
      v[4 * j] = fmaf(fmaf(nm, s.x, v[4 * j]), ln.rstd, b.x);
      v[4 * j + 1] = fmaf(fmaf(nm, s.y, v[4 * j + 1]), ln.rstd, b.y);
      v[4 * j + 2] = fmaf(fmaf(nm, s.z, v[4 * j + 2]), ln.rstd, b.z);
      v[4 * j + 3] = fmaf(fmaf(nm, s.w, v[4 * j + 3]), ln.rstd, b.w);
    }
    return;
  }
  if (p.scale) {
    const float4* s4 = reinterpret_cast<const float4*>(sv_scale);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 s = s4[j];
      v[4 * j] *= s.x; v[4 * j + 1] *= s.y; v[4 * j + 2] *= s.z; v[4 * j + 3] *= s.w;
    }
  }
  if (p.bias) {
    const float4* b4 = reinterpret_cast<const float4*>(sv_bias);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = b4[j];
      v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
    }
  }
}

template <bool RESID = true>
__device__ __forceinline__ void epi_apply(const EpiParams& p, const EpiRow& row, int m, int n0, int N, uint32_t (&r)[32],
                                          const float4 (&t)[8], const float* sv_scale, const float* sv_bias, bool wide,
                                          const EpiLn ln = EpiLn{0.f, 0.f}) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
  epi_scale_bias(p, sv_scale, sv_bias, v, ln);
  if (p.act == ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
  } else if (p.act == ACT_LRELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : 0.2f * v[j];
  }
  const long long orow = row.orow;
  const int col = n0 + p.col_off + row.col_shift;
  if (RESID && p.resid) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[4 * j] += t[j].x; v[4 * j + 1] += t[j].y; v[4 * j + 2] += t[j].z; v[4 * j + 3] += t[j].w;
    }
  }
  if (p.store == STORE_QKV && n0 >= p.qk_cols) {
    // V, transposed store: for a fixed column the 32 lanes of the warp hold 32 consecutive tokens
    int frame = m / p.tokens, tok = m % p.tokens;
    int hd = n0 - p.qk_cols;  // h*128 + d0
    long long base = ((long long)frame * (N - p.qk_cols) + hd) * p.tokens + tok;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      __nv_bfloat16 hi, lo;
      split_bf16(v[j], hi, lo);
      p.vt_hi[base + (long long)j * p.tokens] = hi;
      if (p.vt_lo) p.vt_lo[base + (long long)j * p.tokens] = lo;
    }
    return;
  }
  if (p.out_f32) {
    float* o = p.out_f32 + orow * p.ldo + col;
    if (wide) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        st_global_256(o + 8 * j, __float_as_uint(v[8 * j]), __float_as_uint(v[8 * j + 1]), __float_as_uint(v[8 * j + 2]),
                      __float_as_uint(v[8 * j + 3]), __float_as_uint(v[8 * j + 4]), __float_as_uint(v[8 * j + 5]),
                      __float_as_uint(v[8 * j + 6]), __float_as_uint(v[8 * j + 7]));
    } else {
      float4* o4 = reinterpret_cast<float4*>(o);
#pragma unroll
      for (int j = 0; j < 8; ++j) o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
  }
  if (p.out_hi) {
    uint32_t h[16], l[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) split_pack2(v[2 * j], v[2 * j + 1], h[j], l[j]);
    __nv_bfloat16* oh = p.out_hi + orow * p.ldo + col;
    __nv_bfloat16* ol = p.out_lo ? p.out_lo + orow * p.ldo + col : nullptr;
    if (wide) {
#pragma unroll
      for (int j = 0; j < 2; ++j)
        st_global_256(oh + 16 * j, h[8 * j], h[8 * j + 1], h[8 * j + 2], h[8 * j + 3], h[8 * j + 4], h[8 * j + 5], h[8 * j + 6], h[8 * j + 7]);
      if (ol) {
#pragma unroll
        for (int j = 0; j < 2; ++j)
          st_global_256(ol + 16 * j, l[8 * j], l[8 * j + 1], l[8 * j + 2], l[8 * j + 3], l[8 * j + 4], l[8 * j + 5], l[8 * j + 6], l[8 * j + 7]);
      }
    } else {
      uint4* oh4 = reinterpret_cast<uint4*>(oh);
#pragma unroll
      for (int j = 0; j < 4; ++j) oh4[j] = make_uint4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
      if (ol) {
        uint4* ol4 = reinterpret_cast<uint4*>(ol);
#pragma unroll
        for (int j = 0; j < 4; ++j) ol4[j] = make_uint4(l[4 * j], l[4 * j + 1], l[4 * j + 2], l[4 * j + 3]);
      }
    }
  }
}

// ---- coalesced variant of the epilogue (COAL = true; opt-in EGOTAP_EPI=coalesced) ---------------------------------
// tcgen05.ld hands every thread one ROW of the tile (32 consecutive columns), so the row-strided stores / residual loads
// above touch 32 different 128-byte lines per warp instruction: on the K = 1024 GEMMs (out-projection 607 vs MLP-down
// 1346 TFLOP/s in bf16 mode, profiles/r01c) the epilogue is bound by LSU wavefronts, not by HBM.  Here the 32 x 32 chunk
// of a warp goes through a 4 KB shared-memory staging block (16-byte chunks XOR-swizzled by row, conflict-free per
// quarter warp both ways) and comes back with lane l holding the 16-byte segment l % 8 of rows 4 i + l / 8: every global
// access of the warp then covers 4 full 128-byte lines (fp32) or 4 x 64 contiguous bytes (bf16 hi / lo).  The pointwise
// part runs in the row domain, the residual add and the stores in the transposed domain; results are identical.
struct EpiRowsT {            // transposed-domain addressing of the 8 rows a lane touches, fixed per tile
  int orow[8];               // output row of tile row 4 i + lane / 8 (row counts stay far below 2^31)
  unsigned ok;               // bit i: the row exists
};
__device__ __forceinline__ EpiRowsT epi_rows_t(const EpiRow& row, bool row_ok, int lane) {
  EpiRowsT t;
  t.ok = 0u;
  const int sub = lane >> 3;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = 4 * i + sub;
    t.orow[i] = __shfl_sync(0xffffffffu, int(row.orow), rr);
    t.ok |= unsigned(__shfl_sync(0xffffffffu, row_ok ? 1 : 0, rr)) << i;
  }
  return t;
}
// Residual row and column shift of tile row rr: equal to the output row / zero in the common layouts, fetched from the row's
// owner (warp shuffle, all lanes participate: the conditions are warp-uniform) only in the modes where they differ.
__device__ __forceinline__ void epi_row_extras(const EpiParams& p, const EpiRow& row, int orow_i, int rr, int& rrow, int& cs) {
  rrow = orow_i;
  cs = 0;
  if (p.resid_mod > 0) rrow = __shfl_sync(0xffffffffu, int(row.rrow), rr);
  if (p.store == STORE_JOINT_REGROUP || p.store == STORE_HEAD_MERGE) cs = __shfl_sync(0xffffffffu, row.col_shift, rr);
}
__device__ __forceinline__ void epi_load_resid_t(const EpiParams& p, const EpiRow& row, const EpiRowsT& rows, int n0, int lane,
                                                 float4 (&t)[8]) {
  const int c0 = n0 + p.col_off + (lane & 7) * 4;
  const int sub = lane >> 3;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int rrow, cs;
    epi_row_extras(p, row, rows.orow[i], 4 * i + sub, rrow, cs);
    if ((rows.ok >> i) & 1u) t[i] = *reinterpret_cast<const float4*>(p.resid + (long long)rrow * p.resid_ld + c0 + cs);
  }
}
__device__ __forceinline__ void epi_apply_coalesced(const EpiParams& p, const EpiRow& row, const EpiRowsT& rows, int n0,
                                                    uint32_t (&r)[32], const float4 (&t)[8], float4* stg, int lane,
                                                    const float* sv_scale, const float* sv_bias, const EpiLn ln,
                                                    float (&st_sum)[8], float (&st_sq)[8]) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
  epi_scale_bias(p, sv_scale, sv_bias, v, ln);
  if (p.act == ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
  } else if (p.act == ACT_LRELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : 0.2f * v[j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) stg[lane * 8 + (j ^ (lane & 7))] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
  const int seg = lane & 7, sub = lane >> 3;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = 4 * i + sub;
    float4 x = stg[rr * 8 + (seg ^ (rr & 7))];
    int rrow_unused, cs;
    epi_row_extras(p, row, rows.orow[i], rr, rrow_unused, cs);
    if ((rows.ok >> i) & 1u) {
      if (p.resid) { x.x += t[i].x; x.y += t[i].y; x.z += t[i].z; x.w += t[i].w; }
      if (p.stats_out) {           // this lane's 4 columns of row 4 i + sub; the 8 lanes of the row are summed at the end of the tile
        st_sum[i] += (x.x + x.y) + (x.z + x.w);
        st_sq[i] += fmaf(x.x, x.x, x.y * x.y) + fmaf(x.z, x.z, x.w * x.w);
      }
      const long long o = (long long)rows.orow[i] * p.ldo + n0 + p.col_off + cs + seg * 4;
      if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + o) = x;
      if (p.out_hi) {
        uint32_t h0, l0, h1, l1;
        split_pack2(x.x, x.y, h0, l0);
        split_pack2(x.z, x.w, h1, l1);
        *reinterpret_cast<uint2*>(p.out_hi + o) = make_uint2(h0, h1);
        if (p.out_lo) *reinterpret_cast<uint2*>(p.out_lo + o) = make_uint2(l0, l1);
      }
    }
  }
  __syncwarp();                 // the staging block is rewritten by the next chunk
}

// EW16 = true: 16 epilogue warps instead of 8 (four per tensor-memory lane quarter, 64 columns each).  The row-domain epilogues
// that only write bf16 operands (QKV, MLP-up + GELU) are bound by the latency of their own instruction stream with two warps per
// scheduler (ncu: issue slots 35-47 % used, tensor pipe 55-75 %); four warps per scheduler hide it.  112 registers per thread.
//
// TEPI = true: TMA epilogue for fp32 outputs (residual stream, patch embedding, fp32 partial products).  The coalesced form above
// still executes ~450 instructions per 32 x 32 chunk and keeps only one chunk of residual loads (4 KB per warp) in flight: at
// K = 1024 in bf16 mode the epilogue needed ~21,000 cycles per tile against 8,192 of main loop (ncu, round 2: out-projection at
// 43 % tensor-pipe activity).  Here each epilogue warp owns a ring of NBOX 4 KB boxes: the residual chunk arrives by TMA
// (issued NBOX - 1 chunks ahead, across tile boundaries, i.e. normally a whole main loop ahead), the thread that owns row r adds
// its accumulator row in place (the 128-byte-swizzle position of a 16-byte piece depends on the row only: no bank conflicts, no
// cross-lane exchange) and one lane stores the box by TMA.  No global load / store instruction is left in the epilogue.
// Eligibility (host, gemm_launch.cu): row-major fp32 output only, 32-row boxes map to 32 consecutive output / residual rows.
template <int CG, int BN, int NSPLIT, int STAGES, bool COAL = false, bool TN = false, bool EW16 = false, bool TEPI = false>
struct GemmCfg {
  static constexpr int BM = 128;                     // rows per CTA (tile rows = BM * CG)
  static constexpr int BK = 64;                      // bf16 elements = one 128-byte swizzle row
  static constexpr int BNL = BN / CG;                // B rows loaded by each CTA
  static constexpr int NOPS = (NSPLIT == 1) ? 1 : 2; // hi (+ lo) copies of each operand
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BNL * BK * 2;
  static constexpr int STAGE_BYTES = NOPS * (A_BYTES + B_BYTES);
  static constexpr int BAR_BYTES = TEPI ? 512 : 256;  // TEPI: + one mbarrier per epilogue warp and box (second half)
  static constexpr int TMEM_COLS = 2 * BN;           // double-buffered fp32 accumulator
  static constexpr int EPI_WARPS = EW16 ? 16 : (BN >= 256 ? 8 : 4);
  static_assert(!EW16 || (BN == 256 && !COAL), "16 epilogue warps: BN = 256, row-domain epilogue");
  static constexpr int VEC_BYTES = 2 * BN * 4;       // the tile's columns of the epilogue scale / bias vectors
  static constexpr int STG_BYTES = COAL ? EPI_WARPS * 4096 : 0;   // coalesced epilogue: one 32 x 32 fp32 block per warp
  static constexpr int VEC_OFF = STAGES * STAGE_BYTES + BAR_BYTES;
  static constexpr int STG_OFF = VEC_OFF + VEC_BYTES;
  static constexpr int NBOX = TEPI ? (NSPLIT == 1 ? 3 : 1) : 0;   // boxes per epilogue warp (what fits next to the operand ring)
  static constexpr int BOX_OFF = (STG_OFF + STG_BYTES + 1023) / 1024 * 1024;
  static constexpr int BOX_BYTES = EPI_WARPS * NBOX * 4096;
  static constexpr int SMEM_BYTES = TEPI ? BOX_OFF + BOX_BYTES : STG_OFF + STG_BYTES;   // the dynamic shared memory is declared 1 KB aligned
  static_assert(!TEPI || (!COAL && !TN && !EW16), "TMA epilogue: own form");
  static constexpr int COLS_PER_EPI_GROUP = BN / (EPI_WARPS / 4);
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
  static_assert(TMEM_COLS == 256 || TMEM_COLS == 512 || TMEM_COLS == 128, "TMEM columns must be a power of two");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
  static_assert(BN <= 32 * EPI_WARPS, "one epilogue thread per tile column (staging of the scale / bias vectors)");
};

template <int CG, int BN, int NSPLIT, int STAGES, bool COAL = false, bool TN = false, bool EW16 = false, bool TEPI = false>
__global__ void __launch_bounds__((GemmCfg<CG, BN, NSPLIT, STAGES, COAL, TN, EW16, TEPI>::THREADS), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
               const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
               const GemmShape s, const EpiParams ep, const __grid_constant__ EpiMaps maps) {
  using C = GemmCfg<CG, BN, NSPLIT, STAGES, COAL, TN, EW16, TEPI>;
  EB_DYN_SMEM_1K(smem);
  if ((smem_u32(smem) & 1023u) != 0) __trap();   // 128-byte-swizzle tiles need a 1 KB aligned base
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * C::STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta_rank = (CG == 2) ? int(cluster_ctarank()) : 0;
  const int cluster_id = blockIdx.x / CG;
  const int num_clusters = gridDim.x / CG;

  const int nM = (s.M + C::BM * CG - 1) / (C::BM * CG);
  const int nN = (s.N + BN - 1) / BN;
  const int nK = s.K / C::BK;
  const int num_tiles = s.groups * nM * nN;

  if (warp == 0 && lane == 0) {
    if constexpr (TEPI) { tma_prefetch_desc(&maps.out); tma_prefetch_desc(&maps.res); }
    tma_prefetch_desc(&tmAh);
    tma_prefetch_desc(&tmBh);
    if (NSPLIT > 1) { tma_prefetch_desc(&tmAl); tma_prefetch_desc(&tmBl); }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], C::EPI_WARPS * CG); }
    if constexpr (TEPI) {
      uint64_t* rbar = reinterpret_cast<uint64_t*>(smem + STAGES * C::STAGE_BYTES + 256);
      for (int i = 0; i < C::EPI_WARPS * C::NBOX; ++i) mbar_init(&rbar[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<CG>(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int n_blk = tile % nN; const int t2 = tile / nN;
        const int m_blk = t2 % nM;   const int g = t2 / nM;
        const int g0 = g % s.gdiv, g1 = g / s.gdiv;
        const int bg0 = s.b_shared ? 0 : g0, bg1 = s.b_shared ? 0 : g1;
        const int row_a = m_blk * C::BM * CG + cta_rank * C::BM;
        const int row_b = n_blk * BN + cta_rank * C::BNL;
        for (int kb = 0; kb < nK; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + C::NOPS * C::A_BYTES;
          if constexpr (TN) {
            // 64 contraction rows x 64 columns per box: BM / 64 boxes of A, BNL / 64 boxes of B (8 KB each), group = row chunk
            const int krow = g * s.K + kb * C::BK;
            if (CG == 1 || cta_rank == 0) mbar_expect_tx(&full[stage], C::STAGE_BYTES * CG);
#pragma unroll
            for (int h = 0; h < C::BM / 64; ++h) {
              if (CG == 1) tma_load_4d(sa + h * 8192, &tmAh, &full[stage], row_a + h * 64, krow, 0, 0);
              else tma_load_4d_2sm(sa + h * 8192, &tmAh, &full[stage], row_a + h * 64, krow, 0, 0);
              if (NSPLIT > 1) {
                if (CG == 1) tma_load_4d(sa + C::A_BYTES + h * 8192, &tmAl, &full[stage], row_a + h * 64, krow, 0, 0);
                else tma_load_4d_2sm(sa + C::A_BYTES + h * 8192, &tmAl, &full[stage], row_a + h * 64, krow, 0, 0);
              }
            }
#pragma unroll
            for (int h = 0; h < C::BNL / 64; ++h) {
              if (CG == 1) tma_load_4d(sb + h * 8192, &tmBh, &full[stage], row_b + h * 64, krow, 0, 0);
              else tma_load_4d_2sm(sb + h * 8192, &tmBh, &full[stage], row_b + h * 64, krow, 0, 0);
              if (NSPLIT > 1) {
                if (CG == 1) tma_load_4d(sb + C::B_BYTES + h * 8192, &tmBl, &full[stage], row_b + h * 64, krow, 0, 0);
                else tma_load_4d_2sm(sb + C::B_BYTES + h * 8192, &tmBl, &full[stage], row_b + h * 64, krow, 0, 0);
              }
            }
          } else if (CG == 1) {
            mbar_expect_tx(&full[stage], C::STAGE_BYTES);
            tma_load_4d(sa, &tmAh, &full[stage], kb * C::BK, row_a, g0, g1);
            tma_load_4d(sb, &tmBh, &full[stage], kb * C::BK, row_b, bg0, bg1);
            if (NSPLIT > 1) {
              tma_load_4d(sa + C::A_BYTES, &tmAl, &full[stage], kb * C::BK, row_a, g0, g1);
              tma_load_4d(sb + C::B_BYTES, &tmBl, &full[stage], kb * C::BK, row_b, bg0, bg1);
            }
          } else {
            // both CTAs load their halves; all bytes are accounted on the leader's barrier
            if (cta_rank == 0) mbar_expect_tx(&full[stage], C::STAGE_BYTES * 2);
            tma_load_4d_2sm(sa, &tmAh, &full[stage], kb * C::BK, row_a, g0, g1);
            tma_load_4d_2sm(sb, &tmBh, &full[stage], kb * C::BK, row_b, bg0, bg1);
            if (NSPLIT > 1) {
              tma_load_4d_2sm(sa + C::A_BYTES, &tmAl, &full[stage], kb * C::BK, row_a, g0, g1);
              tma_load_4d_2sm(sb + C::B_BYTES, &tmBl, &full[stage], kb * C::BK, row_b, bg0, bg1);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (cta_rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(C::BM * CG, BN, TN ? 1 : 0, TN ? 1 : 0);
      const uint32_t smem_base = smem_u32(smem);
      int stage = 0; uint32_t phase = 0; int it = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
        const int as = it & 1; const uint32_t aph = (it >> 1) & 1;
        mbar_wait(&tempty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < nK; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          {
            // K-major: 32 bytes per k-step inside the 128-byte rows; MN-major (TN): 16 contraction rows = 2 KB per k-step and
            // LBO = one 8 KB box between 64-column groups
            const uint32_t stage_addr = smem_base + stage * C::STAGE_BYTES;
            const uint32_t a_lo = TN ? (((stage_addr >> 4) & 0x3FFFu) | (uint32_t(8192 >> 4) << 16)) : sdesc_lo(stage_addr);
            constexpr uint32_t B_OFF = C::NOPS * C::A_BYTES;
            constexpr int KSTEP = TN ? 2048 : 32;
#pragma unroll
            for (int k = 0; k < C::BK / 16; ++k) {
              const uint64_t dah = sdesc_at(a_lo, k * KSTEP);
              const uint64_t dbh = sdesc_at(a_lo, B_OFF + k * KSTEP);
              umma_bf16<CG>(d_tmem, dah, dbh, idesc, (kb | k) != 0 ? 1u : 0u);
              if (NSPLIT > 1) {
                const uint64_t dal = sdesc_at(a_lo, C::A_BYTES + k * KSTEP);
                const uint64_t dbl = sdesc_at(a_lo, B_OFF + C::B_BYTES + k * KSTEP);
                umma_bf16<CG>(d_tmem, dah, dbl, idesc, 1u);
                umma_bf16<CG>(d_tmem, dal, dbh, idesc, 1u);
              }
            }
            umma_commit<CG>(&empty[stage]);
            if (kb == nK - 1) umma_commit<CG>(&tfull[as]);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3;
    const int col_base = ((warp - 2) >> 2) * C::COLS_PER_EPI_GROUP;
    const int et = int(threadIdx.x) - 64;          // epilogue thread = the tile column whose scale / bias it stages
    const bool wide = epi_wide_ok(ep);             // 32-byte global accesses in the row-domain epilogue
    float* sv_scale = reinterpret_cast<float*>(smem + C::VEC_OFF);
    float* sv_bias = sv_scale + BN;
    int it = 0;
    if constexpr (TEPI) {
      constexpr int NCH = C::COLS_PER_EPI_GROUP / 32;           // chunks of this warp per tile
      uint8_t* box0 = smem + C::BOX_OFF + (warp - 2) * C::NBOX * 4096;
      uint64_t* rbar = reinterpret_cast<uint64_t*>(smem + STAGES * C::STAGE_BYTES + 256) + (warp - 2) * C::NBOX;
      const bool has_resid = ep.resid != nullptr;
      // this warp's 32 x 32 boxes of a tile: first tile row, first column, number of boxes inside the matrix, group
      auto tile_boxes = [&](int tile, int& g, int& m0, int& n_first) -> int {
        const int n_blk = tile % nN; const int t2 = tile / nN;
        const int m_blk = t2 % nM;   g = t2 / nM;
        m0 = m_blk * C::BM * CG + cta_rank * C::BM + q * 32;
        n_first = n_blk * BN + col_base;
        if (m0 >= s.M || n_first >= s.N) return 0;
        const int left = (s.N - n_first) / 32;
        return left < NCH ? left : NCH;
      };
      // residual boxes are requested NBOX - 1 chunks ahead of their use, along the sequence of this warp's chunks over all its tiles
      int pf_tile = cluster_id, pf_c = 0, pf_nc = 0, pf_g = 0, pf_m0 = 0, pf_nf = 0;
      uint32_t pf_n = 0, n = 0;
      auto pf_settle = [&]() {
        while (pf_tile < num_tiles) {
          if (pf_c == 0) pf_nc = tile_boxes(pf_tile, pf_g, pf_m0, pf_nf);
          if (pf_c < pf_nc) return;
          pf_tile += num_clusters; pf_c = 0;
        }
      };
      auto prefetch_to = [&](uint32_t limit) {      // requests the residual boxes of chunks [pf_n, limit)
        while (pf_n < limit && pf_tile < num_tiles) {
          if (lane == 0) {
            const EpiRow row = epi_row(ep, pf_g, pf_m0, s.N);
            const uint32_t slot = pf_n % C::NBOX;
            mbar_expect_tx(&rbar[slot], 4096);
            tma_load_4d(box0 + slot * 4096, &maps.res, &rbar[slot], 2 * (pf_nf + pf_c * 32 + ep.col_off), int(row.rrow), 0, 0);
          }
          ++pf_n; ++pf_c;
          pf_settle();
        }
      };
      pf_settle();
      if (has_resid) prefetch_to(C::NBOX - 1);
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
        const int n_blk = tile % nN;
        const int as = it & 1; const uint32_t aph = (it >> 1) & 1;
        {
          const int ncol = n_blk * BN + et;
          float my_scale = 1.0f, my_bias = 0.0f;
          if (ep.scale != nullptr && et < BN && ncol < s.N) my_scale = __ldg(ep.scale + ncol);
          if (ep.bias != nullptr && et < BN && ncol < s.N) my_bias = __ldg(ep.bias + ncol);
          named_bar_sync<32 * C::EPI_WARPS>(1);
          if (et < BN) { sv_scale[et] = my_scale; sv_bias[et] = my_bias; }
        }
        mbar_wait(&tfull[as], aph);
        tc_fence_after();
        named_bar_sync<32 * C::EPI_WARPS>(1);
        int g, m0, n_first;
        const int nc = tile_boxes(tile, g, m0, n_first);
        const EpiRow row0 = epi_row(ep, g, m0 < s.M ? m0 : 0, s.N);
        const long long orow_lane = row0.orow + lane;            // boxes map to 32 consecutive output rows (eligibility)
        const EpiLn ln = epi_ln_row(ep, orow_lane, nc > 0);
        float st_sum = 0.f, st_sq = 0.f;                         // LayerNorm-fold producer: this row over the warp's columns
        const uint32_t t_addr = tmem_base + (uint32_t(q * 32) << 16) + as * BN;
#pragma unroll 1
        for (int c = 0; c < nc; ++c, ++n) {
          const int n0 = n_first + c * 32;
          const uint32_t slot = n % C::NBOX;
          uint32_t r[32];
          tmem_ld32(t_addr + col_base + c * 32, r);
          tmem_ld_wait();
          if (c == nc - 1) {                    // the accumulator buffer is free once its last chunk is in registers
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (CG == 2) mbar_arrive_remote(&tempty[as], 0); else mbar_arrive(&tempty[as]); }
          }
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * ep.alpha;
          epi_scale_bias(ep, sv_scale + col_base + c * 32, sv_bias + col_base + c * 32, v, ln);
          if (ep.act == ACT_GELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
          } else if (ep.act == ACT_LRELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : 0.2f * v[j];
          }
          // the box that takes chunk n + NBOX - 1 is the one chunk n - 1 was stored from: its store must have read it
          if (lane == 0) { if (has_resid) bulk_wait_read<0>(); else bulk_wait_read<C::NBOX - 1>(); }
          __syncwarp();
          if (has_resid) {
            prefetch_to(n + C::NBOX);
            mbar_wait(&rbar[slot], (n / C::NBOX) & 1);
          }
          float4* rowp = reinterpret_cast<float4*>(box0 + slot * 4096 + lane * 128);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 x = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            float4* pp = rowp + (j ^ (lane & 7));           // 128-byte swizzle: 16-byte piece j of row r sits at piece j ^ (r & 7)
            if (has_resid) { const float4 t = *pp; x.x += t.x; x.y += t.y; x.z += t.z; x.w += t.w; }
            *pp = x;
            v[4 * j] = x.x; v[4 * j + 1] = x.y; v[4 * j + 2] = x.z; v[4 * j + 3] = x.w;
          }
          if (ep.stats_out != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) { st_sum += v[j]; st_sq = fmaf(v[j], v[j], st_sq); }
          }
          if (ep.out_hi != nullptr) {             // the stored row also as a bf16 (hi / lo) operand: 64 bytes per thread and part
            uint32_t hh[16], ll[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) split_pack2(v[2 * j], v[2 * j + 1], hh[j], ll[j]);
            const long long o = orow_lane * ep.ldo + n0 + ep.col_off;
#pragma unroll
            for (int e = 0; e < 2; ++e)
              st_global_256(ep.out_hi + o + 16 * e, hh[8 * e], hh[8 * e + 1], hh[8 * e + 2], hh[8 * e + 3], hh[8 * e + 4], hh[8 * e + 5],
                            hh[8 * e + 6], hh[8 * e + 7]);
            if (ep.out_lo != nullptr) {
#pragma unroll
              for (int e = 0; e < 2; ++e)
                st_global_256(ep.out_lo + o + 16 * e, ll[8 * e], ll[8 * e + 1], ll[8 * e + 2], ll[8 * e + 3], ll[8 * e + 4], ll[8 * e + 5],
                              ll[8 * e + 6], ll[8 * e + 7]);
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&maps.out, box0 + slot * 4096, 2 * (n0 + ep.col_off), int(row0.orow), 0, 0);
            bulk_commit();
          }
        }
        if (ep.stats_out != nullptr && nc > 0) ep.stats_out[orow_lane * ep.stats_parts + n_first / 128] = make_float2(st_sum, st_sq);
        if (nc == 0) {                          // a warp whose rows / columns lie outside the matrix still hands the buffer back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (CG == 2) mbar_arrive_remote(&tempty[as], 0); else mbar_arrive(&tempty[as]); }
        }
      }
      if (lane == 0) bulk_wait<0>();
      __syncwarp();
    } else {
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
      const int n_blk = tile % nN; const int t2 = tile / nN;
      const int m_blk = t2 % nM;   const int g = t2 / nM;
      const int as = it & 1; const uint32_t aph = (it >> 1) & 1;
      {
        // this tile's columns of the scale / bias vectors: requested before the accumulator wait, staged in shared memory
        const int ncol = n_blk * BN + et;
        float my_scale = 1.0f, my_bias = 0.0f;
        if (ep.scale != nullptr && et < BN && ncol < s.N) my_scale = __ldg(ep.scale + ncol);
        if (ep.bias != nullptr && et < BN && ncol < s.N) my_bias = __ldg(ep.bias + ncol);
        named_bar_sync<32 * C::EPI_WARPS>(1);      // every epilogue warp has finished reading the previous tile's vectors
        if (et < BN) {
          sv_scale[et] = my_scale;
          sv_bias[et] = my_bias;
        }
      }
      mbar_wait(&tfull[as], aph);
      tc_fence_after();
      named_bar_sync<32 * C::EPI_WARPS>(1);        // staged vectors visible to all epilogue warps
      const int m = m_blk * C::BM * CG + cta_rank * C::BM + q * 32 + lane;
      const uint32_t t_addr = tmem_base + (uint32_t(q * 32) << 16) + as * BN;
      const bool row_ok = m < s.M;
      const bool use_resid = ep.resid != nullptr && row_ok;
      const EpiRow row = epi_row(ep, g, row_ok ? m : 0, s.N);
      const EpiLn ln = epi_ln_row(ep, row.orow, row_ok);
      float4 t_cur[8], t_nxt[8];
      const int n_first = n_blk * BN + col_base;
      if constexpr (COAL) {
        // transposed-domain epilogue (see epi_apply_coalesced); the transposed V store of STORE_QKV keeps the row-domain path
        float4* stg = reinterpret_cast<float4*>(smem + C::STG_OFF) + (warp - 2) * 256;
        const EpiRowsT rows = epi_rows_t(row, row_ok, lane);
        // coalesced residual loads run ONE CHUNK AHEAD of the math (ncu, bf16 mode: with the loads issued in the chunk that
        // uses them, 22 % of the out-projection's samples sat on the first residual FADD of a chunk)
        const bool has_resid = ep.resid != nullptr && ep.store != STORE_QKV;
        if (has_resid && n_first < s.N) epi_load_resid_t(ep, row, rows, n_first, lane, t_cur);
        float st_sum[8], st_sq[8];                 // LayerNorm-fold producer: row sums of this warp's columns (transposed domain)
#pragma unroll
        for (int i = 0; i < 8; ++i) { st_sum[i] = 0.f; st_sq[i] = 0.f; }
#pragma unroll 1
        for (int c = 0; c < C::COLS_PER_EPI_GROUP / 32; ++c) {
          const int n0 = n_first + c * 32;
          if (n0 >= s.N) break;
          const bool v_part = ep.store == STORE_QKV && n0 >= ep.qk_cols;
          uint32_t r[32];
          tmem_ld32(t_addr + col_base + c * 32, r);
          const bool more = (c + 1 < C::COLS_PER_EPI_GROUP / 32) && (n0 + 32 < s.N);
          if (has_resid && more) epi_load_resid_t(ep, row, rows, n0 + 32, lane, t_nxt);
          tmem_ld_wait();
          const float* svs = sv_scale + col_base + c * 32;
          const float* svb = sv_bias + col_base + c * 32;
          if (v_part) {
            if (row_ok) epi_apply(ep, row, m, n0, s.N, r, t_cur, svs, svb, wide, ln);
          } else {
            epi_apply_coalesced(ep, row, rows, n0, r, t_cur, stg, lane, svs, svb, ln, st_sum, st_sq);
          }
          if (has_resid && more) {
#pragma unroll
            for (int j = 0; j < 8; ++j) t_cur[j] = t_nxt[j];
          }
        }
        if (ep.stats_out != nullptr && n_first < s.N) {
          // the 8 lanes that hold the 16-byte pieces of a row (same lane / 8) add up; piece 0's lane writes the entry of the
          // warp's 128-column part
#pragma unroll
          for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int d = 1; d < 8; d <<= 1) {
              st_sum[i] += __shfl_xor_sync(0xffffffffu, st_sum[i], d);
              st_sq[i] += __shfl_xor_sync(0xffffffffu, st_sq[i], d);
            }
            if ((lane & 7) == 0 && ((rows.ok >> i) & 1u))
              ep.stats_out[(long long)rows.orow[i] * ep.stats_parts + n_first / 128] = make_float2(st_sum[i], st_sq[i]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 2) mbar_arrive_remote(&tempty[as], 0); else mbar_arrive(&tempty[as]);
        }
        continue;
      }
      if constexpr (EW16) {
        // 16-warp form: launched only for epilogues without a residual (the launcher's policy), so none is compiled in
#pragma unroll 1
        for (int c = 0; c < C::COLS_PER_EPI_GROUP / 32; ++c) {
          const int n0 = n_first + c * 32;
          if (n0 >= s.N) break;
          uint32_t r[32];
          tmem_ld32(t_addr + col_base + c * 32, r);
          tmem_ld_wait();
          if (row_ok) epi_apply<false>(ep, row, m, n0, s.N, r, t_cur, sv_scale + col_base + c * 32, sv_bias + col_base + c * 32, wide, ln);
        }
      } else {
      if (use_resid && n_first < s.N) epi_load_resid(ep, row, n_first, t_cur, wide);
#pragma unroll 1
      for (int c = 0; c < C::COLS_PER_EPI_GROUP / 32; ++c) {
        const int n0 = n_first + c * 32;
        if (n0 >= s.N) break;
        uint32_t r[32];
        tmem_ld32(t_addr + col_base + c * 32, r);
        const bool more = (c + 1 < C::COLS_PER_EPI_GROUP / 32) && (n0 + 32 < s.N);
        if (use_resid && more) epi_load_resid(ep, row, n0 + 32, t_nxt, wide);   // prefetch the next chunk's residual
        tmem_ld_wait();
        if (row_ok) epi_apply(ep, row, m, n0, s.N, r, t_cur, sv_scale + col_base + c * 32, sv_bias + col_base + c * 32, wide, ln);
        if (use_resid && more) {
#pragma unroll
          for (int j = 0; j < 8; ++j) t_cur[j] = t_nxt[j];
        }
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_remote(&tempty[as], 0); else mbar_arrive(&tempty[as]);
      }
    }
    }   // !TEPI
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) tmem_dealloc<CG>(tmem_base, C::TMEM_COLS);
}

}  // namespace eb

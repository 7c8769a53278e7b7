// Bandwidth-bound kernels of the lifting path (everything that is not a tensor-core contraction):
// operand preparation, input ingest, LayerNorm, softmax, propagation-unit gates, regression head.
// All are coalesced, 16-byte vectorised where the layout allows, and sized in multiples of the SM count.
#include "host_util.cuh"
#include "internal.h"
#ifdef EB_HOST_EMU          // CUDA-on-CPU emulation of the tests (tests/cuda_emu): no inline PTX
#include "numeric.cuh"
#else
#include "ptx.cuh"
#endif

namespace eb {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ void store_split4(__nv_bfloat16* hi, __nv_bfloat16* lo, long long off, float4 v) {
  uint32_t h0, h1, l0, l1;
  split_pack2(v.x, v.y, h0, l0);
  split_pack2(v.z, v.w, h1, l1);
  *reinterpret_cast<uint2*>(hi + off) = make_uint2(h0, h1);
  if (lo) *reinterpret_cast<uint2*>(lo + off) = make_uint2(l0, l1);
}

// ---------------------------------------------------------------------------------------------
// fp32 -> bf16 hi/lo split of a (rows x cols) block into a (possibly larger) destination matrix
// (weight packing: plain split, concatenation along N or K; test operand preparation)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) split2d_kernel(const float* __restrict__ src, long long rows, int cols4,
                                                      long long src_ld, __nv_bfloat16* __restrict__ hi,
                                                      __nv_bfloat16* __restrict__ lo, long long dst_ld) {
  const long long n4 = rows * cols4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols4;
    const int c = int(i % cols4) * 4;
    float4 v = __ldg(reinterpret_cast<const float4*>(src + r * src_ld + c));
    store_split4(hi, lo, r * dst_ld + c, v);
  }
}

int split2d_run(const float* src, long long rows, long long cols, long long src_ld, __nv_bfloat16* hi,
                __nv_bfloat16* lo, long long dst_ld, cudaStream_t stream) {
  EB_REQUIRE(src && hi, "split: null pointer");
  EB_REQUIRE(cols % 4 == 0 && src_ld % 4 == 0 && dst_ld % 4 == 0, "split: cols/ld must be multiples of 4");
  if (rows * cols == 0) return 0;
  long long n4 = rows * (cols / 4);
  long long blocks = (n4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  ProfScope prof("split2d_kernel", stream);
  EB_LAUNCH(split2d_kernel, (unsigned)blocks, 256, stream, src, rows, int(cols / 4), src_ld, hi, lo, dst_ld);
  EB_CHECK_LAUNCH("split2d_kernel");
  return 0;
}

int split_bf16_run(const float* src, __nv_bfloat16* hi, __nv_bfloat16* lo, long long n, cudaStream_t stream) {
  EB_REQUIRE(n % 4 == 0, "split_bf16: n (%lld) must be a multiple of 4", n);
  return split2d_run(src, 1, n, n, hi, lo, n, stream);
}

// ---------------------------------------------------------------------------------------------
// Input ingest (reference model/net_architecture.py:688-694 + :375-383 + modeling_vit.py:195):
// one pass over the (B, 6J, 64, 64) fp32 input producing the two GEMM A-operands
//   patches[(b*2J + n)*16 + pr*4 + pc][py*16 + px] = x[b][n][pr*16+py][pc*16+px]      (n < 2J)
//   limbs  [b*2J + view*J + j][d*4096 + pix]       = x[b][2J + view*2J + d*J + j][pix]
// as bf16 hi/lo.  One CTA per 64x64 channel; reads are whole 256-byte image rows.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ingest_kernel(const float* __restrict__ x, int J,
                                                     __nv_bfloat16* __restrict__ p_hi, __nv_bfloat16* __restrict__ p_lo,
                                                     __nv_bfloat16* __restrict__ l_hi, __nv_bfloat16* __restrict__ l_lo) {
  const int C = 6 * J;
  const int c = blockIdx.x % C;
  const long long b = blockIdx.x / C;
  const float4* src = reinterpret_cast<const float4*>(x + (b * C + c) * 4096ll);
  if (c < 2 * J) {
    const long long row0 = (b * 2 * J + c) * 16;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx4 = threadIdx.x + i * 256;
      const int y = idx4 >> 4, x4 = idx4 & 15;
      const int pr = y >> 4, py = y & 15, pc = x4 >> 2, px0 = (x4 & 3) * 4;
      store_split4(p_hi, p_lo, (row0 + pr * 4 + pc) * 256 + py * 16 + px0, __ldg(src + idx4));
    }
  } else {
    const int cc = c - 2 * J;
    const int view = cc / (2 * J), d = (cc % (2 * J)) / J, j = cc % J;
    const long long base = (b * 2 * J + view * J + j) * 8192ll + d * 4096;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx4 = threadIdx.x + i * 256;
      store_split4(l_hi, l_lo, base + idx4 * 4, __ldg(src + idx4));
    }
  }
}

int ingest_run(const float* x, int B, int J, __nv_bfloat16* p_hi, __nv_bfloat16* p_lo, __nv_bfloat16* l_hi,
               __nv_bfloat16* l_lo, cudaStream_t stream) {
  EB_REQUIRE(x && p_hi && l_hi && B > 0, "ingest: bad arguments");
  ProfScope prof("ingest_kernel", stream);
  EB_LAUNCH(ingest_kernel, B * 6 * J, 256, stream, x, J, p_hi, p_lo, l_hi, l_lo);
  EB_CHECK_LAUNCH("ingest_kernel");
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Dummy (mask) token rows: hidden[b][live + i][:] = dummy[i][:]  (= mask_token + permuted pos-emb,
// precomputed at pack time; reference model/modeling_vit.py:137-153)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fill_dummy_kernel(float4* __restrict__ hidden, const float4* __restrict__ dummy,
                                                         int tokens, int live, __nv_bfloat16* __restrict__ xb_hi,
                                                         __nv_bfloat16* __restrict__ xb_lo, float2* __restrict__ stats) {
  const int nd = tokens - live;
  const long long b = blockIdx.x / nd;
  const int i = blockIdx.x % nd;
  const long long row = b * tokens + live + i;
  const float4 v = __ldg(dummy + i * 256 + threadIdx.x);
  hidden[row * 256 + threadIdx.x] = v;
  if (xb_hi != nullptr) {            // LayerNorm fold (plan.cu): the row also as the bf16 operand + its (sum, sum of squares) per 128 columns
    store_split4(xb_hi, xb_lo, row * 1024 + threadIdx.x * 4, v);
    const float s = warp_sum((v.x + v.y) + (v.z + v.w));
    const float q = warp_sum(fmaf(v.x, v.x, v.y * v.y) + fmaf(v.z, v.z, v.w * v.w));
    if ((threadIdx.x & 31) == 0) stats[row * 8 + (threadIdx.x >> 5)] = make_float2(s, q);     // warp w holds columns [128 w, +128)
  }
}

int fill_dummy_run(float* hidden, const float* dummy, int B, int tokens, int live, cudaStream_t stream, __nv_bfloat16* xb_hi,
                   __nv_bfloat16* xb_lo, float2* stats) {
  if (tokens == live) return 0;
  ProfScope prof("fill_dummy_kernel", stream);
  EB_LAUNCH_COOP(fill_dummy_kernel, B * (tokens - live), 256, stream, (float4*)hidden, (const float4*)dummy, tokens, live, xb_hi, xb_lo, stats);
  EB_CHECK_LAUNCH("fill_dummy_kernel");
  return 0;
}

// ---------------------------------------------------------------------------------------------
// LayerNorm fold, weight packing (plan.cu): W'[n][k] = W[n][k] gamma[k] as bf16 hi(/lo);
//   s[n] = sum_k W'[n][k] (of the ROUNDED operand the tensor cores will see, so that mean * s[n] cancels exactly what the GEMM
//   accumulated), c[n] = b[n] + sum_k beta[k] W[n][k].   K = 1024, one CTA per output row.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ln_fold_pack_kernel(const float* __restrict__ W, const float* __restrict__ bias,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                           float* __restrict__ s_out, float* __restrict__ c_out) {
  __shared__ double red[2][8];
  const long long n = blockIdx.x;
  const int k = threadIdx.x * 4;
  const float4 w = __ldg(reinterpret_cast<const float4*>(W + n * 1024 + k));
  const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + k));
  const float4 b = __ldg(reinterpret_cast<const float4*>(beta + k));
  const float4 wp = make_float4(w.x * g.x, w.y * g.y, w.z * g.z, w.w * g.w);
  store_split4(hi, lo, n * 1024 + k, wp);
  const float e[4] = {wp.x, wp.y, wp.z, wp.w};
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat16 h, l;
    split_bf16(e[j], h, l);
    s += double(__bfloat162float(h)) + (lo != nullptr ? double(__bfloat162float(l)) : 0.0);
  }
  double c = double(w.x) * b.x + double(w.y) * b.y + double(w.z) * b.z + double(w.w) * b.w;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ss = 0.0, cc = 0.0;
    for (int i = 0; i < 8; ++i) { ss += red[0][i]; cc += red[1][i]; }
    s_out[n] = float(ss);
    c_out[n] = float(cc + double(__ldg(bias + n)));
  }
}

int ln_fold_pack_run(const float* W, const float* bias, const float* gamma, const float* beta, int N, __nv_bfloat16* hi,
                     __nv_bfloat16* lo, float* s_out, float* c_out, cudaStream_t stream) {
  EB_REQUIRE(W && bias && gamma && beta && hi && s_out && c_out && N > 0, "ln_fold_pack: bad arguments");
  ProfScope prof("ln_fold_pack_kernel", stream);
  EB_LAUNCH_COOP(ln_fold_pack_kernel, N, 256, stream, W, bias, gamma, beta, hi, lo, s_out, c_out);
  EB_CHECK_LAUNCH("ln_fold_pack_kernel");
  return 0;
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over 1024 features, eps 1e-12, fp32 in -> bf16 hi/lo out (the next GEMM's A operand).
// One warp per row, two-pass statistics in registers.  Optional per-frame row compaction
// (input row = frame*rows_in + tok, output row = frame*rows_out + tok, tok < rows_out) drops the
// dummy tokens after the last layer.  Reference model/modeling_vit.py:367,378,609.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layernorm1024_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias, long long out_rows,
                                                            int rows_in, int rows_out, float eps,
                                                            __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                            float* __restrict__ out_f32) {
  const int lane = threadIdx.x & 31;
  const long long r = blockIdx.x * 8ll + (threadIdx.x >> 5);
  if (r >= out_rows) return;
  const long long in_row = (r / rows_out) * rows_in + (r % rows_out);
  const float4* src = reinterpret_cast<const float4*>(x + in_row * 1024);
  float4 v[8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] = src[lane + 32 * i];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.0f / 1024.0f);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / 1024.0f) + eps);
  const float4* w4 = reinterpret_cast<const float4*>(w);
  const float4* b4 = reinterpret_cast<const float4*>(bias);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 ww = __ldg(w4 + lane + 32 * i), bb = __ldg(b4 + lane + 32 * i);
    float4 y = make_float4(v[i].x * rstd * ww.x + bb.x, v[i].y * rstd * ww.y + bb.y, v[i].z * rstd * ww.z + bb.z,
                           v[i].w * rstd * ww.w + bb.w);
    const long long off = r * 1024 + (lane + 32 * i) * 4;
    if (hi) store_split4(hi, lo, off, y);
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + off) = y;
  }
}

int layernorm_run(const float* x, const float* w, const float* b, long long frames, int rows_in, int rows_out, float eps,
                  __nv_bfloat16* hi, __nv_bfloat16* lo, float* out_f32, cudaStream_t stream) {
  EB_REQUIRE(x && w && b && (hi || out_f32), "layernorm: null pointer");
  EB_REQUIRE(rows_out > 0 && rows_out <= rows_in, "layernorm: rows_out must be in (0, rows_in]");
  const long long out_rows = frames * rows_out;
  if (out_rows == 0) return 0;
  ProfScope prof("layernorm1024_kernel", stream);
  EB_LAUNCH_COOP(layernorm1024_kernel, (unsigned)((out_rows + 7) / 8), 256, stream, x, w, b, out_rows, rows_in, rows_out, eps, hi,
                                                                         lo, out_f32);
  EB_CHECK_LAUNCH("layernorm1024_kernel");
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Row softmax over `cols` (= 576 keys) scaled scores, fp32 in -> bf16 hi/lo probabilities out.
// One warp per row.  Reference model/modeling_vit.py:239-242.
// ---------------------------------------------------------------------------------------------
template <int PER_LANE2>  // float2 per lane; cols = 64 * PER_LANE2
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ s, long long rows,
                                                           __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const int lane = threadIdx.x & 31;
  const long long r = blockIdx.x * 8ll + (threadIdx.x >> 5);
  if (r >= rows) return;
  constexpr int cols = 64 * PER_LANE2;
  const float2* src = reinterpret_cast<const float2*>(s + r * cols);
  float2 v[PER_LANE2];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < PER_LANE2; ++i) {
    v[i] = src[lane + 32 * i];
    m = fmaxf(m, fmaxf(v[i].x, v[i].y));
  }
  m = warp_max(m);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < PER_LANE2; ++i) {
    v[i].x = expf(v[i].x - m);
    v[i].y = expf(v[i].y - m);
    sum += v[i].x + v[i].y;
  }
  const float inv = 1.0f / warp_sum(sum);
#pragma unroll
  for (int i = 0; i < PER_LANE2; ++i) {
    __nv_bfloat16 h0, h1, l0, l1;
    split_bf16(v[i].x * inv, h0, l0);
    split_bf16(v[i].y * inv, h1, l1);
    const long long off = r * cols + (lane + 32 * i) * 2;
    *reinterpret_cast<uint32_t*>(hi + off) = pack_bf16x2(h0, h1);
    if (lo) *reinterpret_cast<uint32_t*>(lo + off) = pack_bf16x2(l0, l1);
  }
}

int softmax_run(const float* s, long long rows, int cols, __nv_bfloat16* hi, __nv_bfloat16* lo, cudaStream_t stream) {
  EB_REQUIRE(s && hi, "softmax: null pointer");
  EB_REQUIRE(cols == 576, "softmax: only 576 columns (24x24 tokens) are compiled, got %d", cols);
  if (rows == 0) return 0;
  ProfScope prof("softmax_rows_kernel", stream);
  EB_LAUNCH_COOP(softmax_rows_kernel<9>, (unsigned)((rows + 7) / 8), 256, stream, s, rows, hi, lo);
  EB_CHECK_LAUNCH("softmax_rows_kernel");
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Propagation unit, input-driven bridge gate (reference model/custom_cells.py:99-102):
//   b' = sigmoid(x2f(x)[:, H:H+X]) * bridge          -> bf16 hi/lo, columns [X, 2X) of the layer-0 x|b operand
// f: (rows, f_ld) fp32 (x2f output), e: (rows, e_ld) fp32 with the bridge (limb embedding) in columns [X, 2X)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pu_bridge_gate_kernel(const float* __restrict__ f, int f_ld, int f_col,
                                                             const float* __restrict__ e, int e_ld, int X, long long rows,
                                                             __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const int x4 = X / 4;
  const long long n4 = rows * x4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / x4;
    const int c = int(i % x4) * 4;
    const float4 g = *reinterpret_cast<const float4*>(f + r * f_ld + f_col + c);
    const float4 b = *reinterpret_cast<const float4*>(e + r * e_ld + X + c);
    store_split4(hi, lo, r * e_ld + X + c,
                 make_float4(sigmoidf_(g.x) * b.x, sigmoidf_(g.y) * b.y, sigmoidf_(g.z) * b.z, sigmoidf_(g.w) * b.w));
  }
}

int pu_bridge_gate_run(const float* f, int f_ld, int f_col, const float* e, int e_ld, int X, long long rows,
                       __nv_bfloat16* hi, __nv_bfloat16* lo, cudaStream_t stream) {
  EB_REQUIRE(f && e && hi, "pu_bridge_gate: null pointer");
  long long n4 = rows * (X / 4);
  long long blocks = (n4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  ProfScope prof("pu_bridge_gate_kernel", stream);
  EB_LAUNCH(pu_bridge_gate_kernel, (unsigned)blocks, 256, stream, f, f_ld, f_col, e, e_ld, X, rows, hi, lo);
  EB_CHECK_LAUNCH("pu_bridge_gate_kernel");
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Propagation unit cell update for one joint step t (reference model/custom_cells.py:109-120):
//   (fg, ig, cg, og) = gates[b][0:H | H:2H | 2H:3H | 3H:4H]
//   c = c * sigmoid(fg) + sigmoid(ig) * tanh(cg) ;  h = sigmoid(og) * tanh(c)
// and the NEXT step's pre-gated recurrent operand (custom_cells.py:101):
//   hg_next = sigmoid(F[b*J + t + 1][0:H]) * h        -> bf16 hi/lo (A operand of the next h2h GEMM)
// h is written to out[(b*J + t)][0:H] as fp32 (and bf16 hi/lo when it feeds the next layer's GEMMs).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pu_cell_kernel(const float* __restrict__ gates, long long gates_ld,
                                                      float* __restrict__ c, const float* __restrict__ F, int F_ld, int t,
                                                      int J, int H, long long B, float* __restrict__ out,
                                                      __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo,
                                                      __nv_bfloat16* __restrict__ hg_hi, __nv_bfloat16* __restrict__ hg_lo) {
  const int h4 = H / 4;
  const long long n4 = B * h4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / h4;
    const int u = int(i % h4) * 4;
    const float* g = gates + b * gates_ld + u;
    const float4 fg = *reinterpret_cast<const float4*>(g);
    const float4 ig = *reinterpret_cast<const float4*>(g + H);
    const float4 cg = *reinterpret_cast<const float4*>(g + 2 * H);
    const float4 og = *reinterpret_cast<const float4*>(g + 3 * H);
    float4 cv = *reinterpret_cast<float4*>(c + b * H + u);
    cv.x = cv.x * sigmoidf_(fg.x) + sigmoidf_(ig.x) * tanhf(cg.x);
    cv.y = cv.y * sigmoidf_(fg.y) + sigmoidf_(ig.y) * tanhf(cg.y);
    cv.z = cv.z * sigmoidf_(fg.z) + sigmoidf_(ig.z) * tanhf(cg.z);
    cv.w = cv.w * sigmoidf_(fg.w) + sigmoidf_(ig.w) * tanhf(cg.w);
    *reinterpret_cast<float4*>(c + b * H + u) = cv;
    const float4 hv = make_float4(sigmoidf_(og.x) * tanhf(cv.x), sigmoidf_(og.y) * tanhf(cv.y),
                                  sigmoidf_(og.z) * tanhf(cv.z), sigmoidf_(og.w) * tanhf(cv.w));
    const long long orow = b * J + t;
    *reinterpret_cast<float4*>(out + orow * H + u) = hv;
    if (out_hi) store_split4(out_hi, out_lo, orow * H + u, hv);
    if (t + 1 < J) {
      const float4 f = *reinterpret_cast<const float4*>(F + (orow + 1) * F_ld + u);
      store_split4(hg_hi, hg_lo, b * H + u,
                   make_float4(sigmoidf_(f.x) * hv.x, sigmoidf_(f.y) * hv.y, sigmoidf_(f.z) * hv.z, sigmoidf_(f.w) * hv.w));
    }
  }
}

int pu_cell_run(const float* gates, long long gates_ld, float* c, const float* F, int F_ld, int t, int J, int H,
                long long B, float* out, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, __nv_bfloat16* hg_hi,
                __nv_bfloat16* hg_lo, cudaStream_t stream) {
  EB_REQUIRE(gates && c && F && out && hg_hi, "pu_cell: null pointer");
  long long n4 = B * (H / 4);
  long long blocks = (n4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  ProfScope prof("pu_cell_kernel", stream);
  EB_LAUNCH(pu_cell_kernel, (unsigned)blocks, 256, stream, gates, gates_ld, c, F, F_ld, t, J, H, B, out, out_hi, out_lo, hg_hi,
                                                      hg_lo);
  EB_CHECK_LAUNCH("pu_cell_kernel");
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Regression head (reference model/net_architecture.py:732-751):
//   pose[b][j] = Wp . [pos_embed[b][j] (X) | skel[b][j] (H)] + bp                       j < J
//   UnrealEgo: o = Wg . skel[b].flatten() + bg ; pose[b][j] += o[0:3] ; pose[b][J] = o[3:6]   (head joint LAST)
// One CTA per frame, one warp per output scalar (shuffle reduction).  e: (B*J, e_ld) with pos_embed in [0, X).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) head_kernel(const float* __restrict__ e, int e_ld, const float* __restrict__ skel,
                                                   const float* __restrict__ Wp, const float* __restrict__ bp,
                                                   const float* __restrict__ Wg, const float* __restrict__ bg, int J, int X,
                                                   int H, float* __restrict__ pose) {
  __shared__ float s_off[8];
  const long long b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nj = Wg ? J + 1 : J;
  const float* sk = skel + b * J * H;
  if (Wg) {
    if (warp < 6) {
      const float4* w4 = reinterpret_cast<const float4*>(Wg + (long long)warp * J * H);
      const float4* s4 = reinterpret_cast<const float4*>(sk);
      float acc = 0.f;
      for (int i = lane; i < J * H / 4; i += 32) {
        const float4 a = __ldg(w4 + i), v = s4[i];
        acc += (a.x * v.x + a.y * v.y) + (a.z * v.z + a.w * v.w);
      }
      acc = warp_sum(acc);
      if (lane == 0) s_off[warp] = acc + bg[warp];
    }
    __syncthreads();
    if (threadIdx.x < 3) pose[(b * nj + J) * 3 + threadIdx.x] = s_off[3 + threadIdx.x];
  }
  for (int o = warp; o < J * 3; o += 8) {
    const int j = o / 3, k = o % 3;
    const float4* w4 = reinterpret_cast<const float4*>(Wp + k * (X + H));
    const float4* p4 = reinterpret_cast<const float4*>(e + (b * J + j) * e_ld);
    const float4* s4 = reinterpret_cast<const float4*>(sk + j * H);
    float acc = 0.f;
    for (int i = lane; i < X / 4; i += 32) {
      const float4 a = __ldg(w4 + i), v = p4[i];
      acc += (a.x * v.x + a.y * v.y) + (a.z * v.z + a.w * v.w);
    }
    for (int i = lane; i < H / 4; i += 32) {
      const float4 a = __ldg(w4 + X / 4 + i), v = s4[i];
      acc += (a.x * v.x + a.y * v.y) + (a.z * v.z + a.w * v.w);
    }
    acc = warp_sum(acc);
    if (lane == 0) pose[(b * nj + j) * 3 + k] = acc + bp[k] + (Wg ? s_off[k] : 0.f);
  }
}

int head_run(const float* e, int e_ld, const float* skel, const float* Wp, const float* bp, const float* Wg,
             const float* bg, long long B, int J, int X, int H, float* pose, cudaStream_t stream) {
  EB_REQUIRE(e && skel && Wp && bp && pose, "head: null pointer");
  if (B == 0) return 0;
  ProfScope prof("head_kernel", stream);
  EB_LAUNCH_COOP(head_kernel, (unsigned)B, 256, stream, e, e_ld, skel, Wp, bp, Wg, bg, J, X, H, pose);
  EB_CHECK_LAUNCH("head_kernel");
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Weight-packing helpers (run once per load_state_dict)
// ---------------------------------------------------------------------------------------------
// eval-mode BatchNorm1d folded into the preceding Linear (reference model/network_utils.py:123-142):
//   y = ((Wx + b) - mu) / sqrt(var + 1e-5) * gamma + beta = (Wx) * scale + shift
__global__ void bn_fold_kernel(const float* fcb, const float* gamma, const float* beta, const float* mean, const float* var,
                               int n, float eps, float* scale, float* shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float s = gamma[i] / sqrtf(var[i] + eps);
  scale[i] = s;
  shift[i] = (fcb[i] - mean[i]) * s + beta[i];
}
int bn_fold_run(const float* fcb, const float* gamma, const float* beta, const float* mean, const float* var, int n,
                float* scale, float* shift, cudaStream_t stream) {
  ProfScope prof("bn_fold_kernel", stream);
  EB_LAUNCH(bn_fold_kernel, (n + 255) / 256, 256, stream, fcb, gamma, beta, mean, var, n, 1e-5f, scale, shift);
  EB_CHECK_LAUNCH("bn_fold_kernel");
  return 0;
}

// out[i] = a[i] + b[i] + c[i]   (b, c nullable)
__global__ void vec_add3_kernel(const float* a, const float* b, const float* c, float* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = a[i] + (b ? b[i] : 0.f) + (c ? c[i] : 0.f);
}
int vec_add3_run(const float* a, const float* b, const float* c, float* out, int n, cudaStream_t stream) {
  ProfScope prof("vec_add3_kernel", stream);
  EB_LAUNCH(vec_add3_kernel, (n + 255) / 256, 256, stream, a, b, c, out, n);
  EB_CHECK_LAUNCH("vec_add3_kernel");
  return 0;
}

// Position embeddings permuted from the mosaic's raster order to heatmap-major token order
// (token t' = n*16 + pr*4 + pc  <-  raster ((n/grid)*4 + pr)*side + (n%grid)*4 + pc; reference
// model/net_architecture.py:397-402), plus the constant dummy rows mask_token + pos for n >= n_hm.
__global__ void pos_permute_kernel(const float* __restrict__ pos, const float* __restrict__ mask_token, int grid, int n_hm,
                                   float* __restrict__ pos_perm, float* __restrict__ dummy) {
  const int tp = blockIdx.x;  // heatmap-major token
  const int n = tp >> 4, pr = (tp >> 2) & 3, pc = tp & 3;
  const int side = grid * 4;
  const int raster = ((n / grid) * 4 + pr) * side + (n % grid) * 4 + pc;
  for (int c = threadIdx.x; c < 1024; c += blockDim.x) {
    const float p = pos[raster * 1024 + c];
    pos_perm[tp * 1024 + c] = p;
    if (n >= n_hm) dummy[(tp - n_hm * 16) * 1024 + c] = p + mask_token[c];
  }
}
int pos_permute_run(const float* pos, const float* mask_token, int grid, int n_hm, float* pos_perm, float* dummy,
                    cudaStream_t stream) {
  ProfScope prof("pos_permute_kernel", stream);
  EB_LAUNCH(pos_permute_kernel, grid * grid * 16, 256, stream, pos, mask_token, grid, n_hm, pos_perm, dummy);
  EB_CHECK_LAUNCH("pos_permute_kernel");
  return 0;
}

}  // namespace eb

extern "C" int egotap_b200_split_bf16(const float* src, void* hi, void* lo, long long n, void* stream) {
  return eb::split_bf16_run(src, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, n, (cudaStream_t)stream);
}

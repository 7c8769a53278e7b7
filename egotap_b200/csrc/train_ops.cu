// Bandwidth-bound kernels of the TRAINING step that are not tied to one layer type: transposed / split operand
// preparation for the dX and dW contractions, column reductions (bias gradients), split-K partial sums, GELU,
// LayerNorm backward, softmax backward, train-mode BatchNorm1d (statistics, apply, backward).
// Host orchestration: egotap_b200/training.py; semantics of every entry: oracle/op_oracle.py (test infrastructure).
// All kernels are plain data-parallel kernels (no inter-CTA synchronisation), coalesced, vectorised where the layout
// allows, reductions in two deterministic stages (no floating-point atomics).
#include "host_util.cuh"
#include "numeric.cuh"

namespace eb {
namespace {

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ void st_split4(__nv_bfloat16* hi, __nv_bfloat16* lo, long long off, float4 v) {
  uint32_t h0, h1, l0, l1;
  split_pack2(v.x, v.y, h0, l0);
  split_pack2(v.z, v.w, h1, l1);
  *reinterpret_cast<uint2*>(hi + off) = make_uint2(h0, h1);
  if (lo) *reinterpret_cast<uint2*>(lo + off) = make_uint2(l0, l1);
}
__device__ __forceinline__ long long src_row(long long r, int rows_in, int rows_out) {
  return rows_out > 0 ? (r / rows_out) * rows_in + (r % rows_out) : r;
}
inline unsigned blocks_for(long long n, int per_block, long long cap) {
  long long b = (n + per_block - 1) / per_block;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return unsigned(b);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// transpose_split: the "gradient preparation" pass.  fp32 (rows x cols) -> row-major bf16 pair rm[r][c] and/or
// transposed pair t[c][r] (t[c][rows .. pad_rows) = 0), optionally with
//   * gelu_u != null: the value is first multiplied by gelu'(u[r][c]) (u has the source's layout) -- the backward of the
//     MLP activation fused into the pass that consumes it, so d u is never written in fp32
//   * colpart != null: per-tile column sums colpart[blockIdx.y][c] (the bias gradient's first reduction stage)
// Logical row r lives at source row (r/rows_out)*rows_in + r%rows_out (rows_out > 0).
// 64 x 64 tiles through shared memory; all outputs are written with full-width coalesced stores.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

__global__ void __launch_bounds__(256) transpose_split_kernel(const float* __restrict__ src, long long rows, int cols,
                                                              long long src_ld, int rows_in, int rows_out,
                                                              __nv_bfloat16* __restrict__ rm_hi,
                                                              __nv_bfloat16* __restrict__ rm_lo, long long rm_ld,
                                                              __nv_bfloat16* __restrict__ t_hi,
                                                              __nv_bfloat16* __restrict__ t_lo, long long t_ld,
                                                              long long pad_rows, const float* __restrict__ gelu_u,
                                                              float* __restrict__ colpart) {
  __shared__ float tile[64][65];
  const int c0 = blockIdx.x * 64;
  const long long r0 = (long long)blockIdx.y * 64;
  {
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // 16 float4 columns x 16 rows per pass
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rr = ty + 16 * i;
      const long long r = r0 + rr;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < rows) {
        const long long so = src_row(r, rows_in, rows_out) * src_ld + c0 + tx * 4;
        v = *reinterpret_cast<const float4*>(src + so);
        if (gelu_u) {
          const float4 u = *reinterpret_cast<const float4*>(gelu_u + so);
          v.x *= gelu_grad(u.x); v.y *= gelu_grad(u.y); v.z *= gelu_grad(u.z); v.w *= gelu_grad(u.w);
        }
        if (rm_hi) st_split4(rm_hi, rm_lo, r * rm_ld + c0 + tx * 4, v);
      }
      tile[rr][tx * 4 + 0] = v.x; tile[rr][tx * 4 + 1] = v.y; tile[rr][tx * 4 + 2] = v.z; tile[rr][tx * 4 + 3] = v.w;
    }
  }
  if (!t_hi && !colpart) return;
  __syncthreads();
  if (colpart && threadIdx.x < 64 && r0 < rows) {             // rows beyond `rows` hold zeros
    float acc = 0.f;
#pragma unroll 8
    for (int rr = 0; rr < 64; ++rr) acc += tile[rr][threadIdx.x];
    colpart[(long long)blockIdx.y * cols + c0 + threadIdx.x] = acc;
  }
  if (t_hi) {
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 row pairs x 8 columns per pass
    const long long r = r0 + tx * 2;
    if (r < pad_rows) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int cc = ty + 8 * i;
        uint32_t h, l;
        split_pack2(tile[tx * 2][cc], tile[tx * 2 + 1][cc], h, l);
        const long long off = (long long)(c0 + cc) * t_ld + r;
        *reinterpret_cast<uint32_t*>(t_hi + off) = h;
        if (t_lo) *reinterpret_cast<uint32_t*>(t_lo + off) = l;
      }
    }
  }
}

int reduce_partials_run(const float* partials, int G, long long n, float* out, cudaStream_t st);

int transpose_split_run(const float* src, long long rows, int cols, long long src_ld, int rows_in, int rows_out,
                        __nv_bfloat16* rm_hi, __nv_bfloat16* rm_lo, long long rm_ld, __nv_bfloat16* t_hi,
                        __nv_bfloat16* t_lo, long long t_ld, long long pad_rows, const float* gelu_u, float* colsum_out,
                        float* scratch, long long scratch_elems, cudaStream_t st) {
  EB_REQUIRE(src && (rm_hi || t_hi || colsum_out), "transpose_split: null pointer");
  EB_REQUIRE(rows > 0 && cols > 0 && cols % 64 == 0, "transpose_split: cols (%d) must be a positive multiple of 64", cols);
  EB_REQUIRE(src_ld % 4 == 0 && (!rm_hi || rm_ld % 4 == 0), "transpose_split: leading dimensions must be multiples of 4");
  EB_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (!gelu_u || (reinterpret_cast<uintptr_t>(gelu_u) & 15) == 0),
             "transpose_split: source must be 16-byte aligned");
  long long extent = rows;
  if (t_hi) {
    EB_REQUIRE(pad_rows >= rows && pad_rows % 2 == 0 && t_ld % 2 == 0 && t_ld >= pad_rows,
               "transpose_split: bad padding (rows %lld pad %lld ld %lld)", rows, pad_rows, t_ld);
    extent = pad_rows;
  }
  const long long gy = (extent + 63) / 64, gy_rows = (rows + 63) / 64;
  EB_REQUIRE(gy <= 65535, "transpose_split: too many rows (%lld)", extent);
  float* colpart = nullptr;
  if (colsum_out) {
    EB_REQUIRE(scratch && gy_rows * (long long)cols <= scratch_elems, "transpose_split: scratch too small for the column sums "
               "(%lld x %d floats needed)", gy_rows, cols);
    colpart = gy_rows == 1 ? colsum_out : scratch;
  }
  {
    ProfScope prof("transpose_split_kernel", st);
    EB_LAUNCH_COOP(transpose_split_kernel, dim3(cols / 64, (unsigned)gy), 256, st, src, rows, cols, src_ld, rows_in, rows_out,
                   rm_hi, rm_lo, rm_ld, t_hi, t_lo, t_ld, pad_rows, gelu_u, colpart);
    EB_CHECK_LAUNCH("transpose_split_kernel");
  }
  if (colsum_out && gy_rows > 1) return reduce_partials_run(scratch, int(gy_rows), cols, colsum_out, st);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// transpose_bf16: batched bf16 transpose d[g1][g0][c][r] = s[g1][g0][r][c], d[..][c][rows .. pad_rows) = 0;
// blockIdx.z enumerates (part, g1, g0) with part = hi / lo.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) transpose_bf16_kernel(const uint16_t* __restrict__ s_hi, const uint16_t* __restrict__ s_lo,
                                                             long long rows, int cols, long long s_ld, int g0c,
                                                             long long s_g0s, long long s_g1s, uint16_t* __restrict__ d_hi,
                                                             uint16_t* __restrict__ d_lo, long long d_ld, long long d_g0s,
                                                             long long d_g1s, long long pad_rows, int groups) {
  __shared__ uint16_t tile[64][66];
  const int part = blockIdx.z / groups, g = blockIdx.z % groups;
  const uint16_t* s = part == 0 ? s_hi : s_lo;
  uint16_t* d = part == 0 ? d_hi : d_lo;
  const long long goff_s = (long long)(g / g0c) * s_g1s + (long long)(g % g0c) * s_g0s;
  const long long goff_d = (long long)(g / g0c) * d_g1s + (long long)(g % g0c) * d_g0s;
  const int c0 = blockIdx.x * 64;
  const long long r0 = (long long)blockIdx.y * 64;
  {
    const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;    // 8 x 16-byte columns, 32 rows per pass
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int rr = ty + 32 * i;
      const long long r = r0 + rr;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (s != nullptr && r < rows) v = *reinterpret_cast<const uint4*>(s + goff_s + r * s_ld + c0 + tx * 8);
      uint32_t* trow = reinterpret_cast<uint32_t*>(&tile[rr][tx * 8]);
      trow[0] = v.x; trow[1] = v.y; trow[2] = v.z; trow[3] = v.w;
    }
  }
  __syncthreads();
  {
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const long long r = r0 + tx * 2;
    if (r < pad_rows) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int cc = ty + 8 * i;
        const uint32_t v = uint32_t(tile[tx * 2][cc]) | (uint32_t(tile[tx * 2 + 1][cc]) << 16);
        *reinterpret_cast<uint32_t*>(d + goff_d + (long long)(c0 + cc) * d_ld + r) = v;
      }
    }
  }
}

int transpose_bf16_run(const void* s_hi, const void* s_lo, long long rows, int cols, long long s_ld, int g0c,
                       long long s_g0s, int g1c, long long s_g1s, void* d_hi, void* d_lo, long long d_ld, long long d_g0s,
                       long long d_g1s, long long pad_rows, cudaStream_t st) {
  EB_REQUIRE(s_hi && d_hi, "transpose_bf16: null pointer");
  EB_REQUIRE(rows > 0 && cols > 0 && cols % 64 == 0, "transpose_bf16: cols (%d) must be a positive multiple of 64", cols);
  EB_REQUIRE(s_ld % 8 == 0 && s_g0s % 8 == 0 && s_g1s % 8 == 0, "transpose_bf16: source strides must be multiples of 8");
  EB_REQUIRE((reinterpret_cast<uintptr_t>(s_hi) & 15) == 0 && (!s_lo || (reinterpret_cast<uintptr_t>(s_lo) & 15) == 0),
             "transpose_bf16: source must be 16-byte aligned");
  EB_REQUIRE(pad_rows >= rows && pad_rows % 2 == 0 && d_ld % 2 == 0 && d_ld >= pad_rows && d_g0s % 2 == 0 && d_g1s % 2 == 0,
             "transpose_bf16: bad destination layout (rows %lld pad %lld ld %lld)", rows, pad_rows, d_ld);
  if (g0c < 1) g0c = 1;
  if (g1c < 1) g1c = 1;
  const int groups = g0c * g1c;
  // a destination lo without a source lo (cannot happen from training.py) would be left undefined: zero-fill it
  const int parts = d_lo ? 2 : 1;
  const long long gy = (pad_rows + 63) / 64;
  EB_REQUIRE(gy <= 65535 && (long long)groups * parts <= 65535, "transpose_bf16: grid too large");
  ProfScope prof("transpose_bf16_kernel", st);
  EB_LAUNCH_COOP(transpose_bf16_kernel, dim3(cols / 64, (unsigned)gy, groups * parts), 256, st, 
      (const uint16_t*)s_hi, (const uint16_t*)s_lo, rows, cols, s_ld, g0c, s_g0s, s_g1s, (uint16_t*)d_hi, (uint16_t*)d_lo,
      d_ld, d_g0s, d_g1s, pad_rows, groups);
  EB_CHECK_LAUNCH("transpose_bf16_kernel");
  return 0;
}

// ---------------------------------------------------------------------------------------------
// colsum: out[c] = sum_r src[srow(r)][c].  Stage 1: grid (cols/128, S) -- 32 float4 column lanes x 8 row lanes per
// CTA, each CTA reduces its slice of the rows; stage 2 sums the S partial rows (fixed order: deterministic).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ src, long long rows, int cols,
                                                             long long ld, int rows_in, int rows_out, long long chunk,
                                                             float* __restrict__ out) {
  __shared__ float4 red[8][32];
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + lx) * 4;
  const long long rbeg = (long long)blockIdx.y * chunk;
  long long rend = rbeg + chunk;
  if (rend > rows) rend = rows;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < cols) {
    for (long long r = rbeg + ly; r < rend; r += 8) {
      const float4 v = *reinterpret_cast<const float4*>(src + src_row(r, rows_in, rows_out) * ld + c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  red[ly][lx] = acc;
  __syncthreads();
  if (ly == 0 && c < cols) {
#pragma unroll
    for (int i = 1; i < 8; ++i) {
      const float4 v = red[i][lx];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(out + (long long)blockIdx.y * cols + c) = acc;
  }
}

__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ p, int G, long long n4,
                                                              float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 acc = reinterpret_cast<const float4*>(p)[i];
    for (int g = 1; g < G; ++g) {
      const float4 v = reinterpret_cast<const float4*>(p)[(long long)g * n4 + i];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(out)[i] = acc;
  }
}

// Tall reductions (G in the thousands, n a few thousand: the per-tile column sums behind every bias gradient): one thread per
// output element would leave ~1 k threads walking 2 k rows each (measured on the B200: 0.47 ms for 2304 x 4096 floats = 80 GB/s).
// Here a CTA owns 64 output floats and its 16 row lanes walk the G rows interleaved; the 16 lane sums are folded in a fixed
// order (deterministic), so n / 64 CTAs x 256 threads stream the partials.
__global__ void __launch_bounds__(256) reduce_partials_tall_kernel(const float* __restrict__ p, int G, long long n4,
                                                                   float* __restrict__ out) {
  __shared__ float4 part[16][16];
  const int cl = threadIdx.x & 15, rl = threadIdx.x >> 4;
  const long long i = (long long)blockIdx.x * 16 + cl;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < n4) {
    for (int g = rl; g < G; g += 16) {
      const float4 v = reinterpret_cast<const float4*>(p)[(long long)g * n4 + i];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  part[rl][cl] = acc;
  __syncthreads();
  if (rl == 0 && i < n4) {
    for (int r = 1; r < 16; ++r) {
      const float4 v = part[r][cl];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(out)[i] = acc;
  }
}

int reduce_partials_run(const float* partials, int G, long long n, float* out, cudaStream_t st) {
  EB_REQUIRE(partials && out && G >= 1 && n > 0 && n % 4 == 0, "reduce_partials: bad arguments (G %d n %lld)", G, n);
  ProfScope prof("reduce_partials_kernel", st);
  const long long n4 = n / 4;
  if (G >= 64 && n4 <= 148 * 8 * 256 / 4) {       // tall and narrow: the flat kernel would not fill the machine
    EB_LAUNCH_COOP(reduce_partials_tall_kernel, (unsigned)((n4 + 15) / 16), 256, st, partials, G, n4, out);
  } else {
    EB_LAUNCH(reduce_partials_kernel, (blocks_for(n4, 256, 148 * 8)), 256, st, partials, G, n4, out);
  }
  EB_CHECK_LAUNCH("reduce_partials_kernel");
  return 0;
}

int colsum_run(const float* src, long long rows, int cols, long long ld, int rows_in, int rows_out, float* out,
               float* scratch, long long scratch_elems, cudaStream_t st) {
  EB_REQUIRE(src && out, "colsum: null pointer");
  EB_REQUIRE(rows > 0 && cols > 0 && cols % 4 == 0 && ld % 4 == 0, "colsum: cols / ld must be multiples of 4");
  const int gx = (cols + 127) / 128;
  long long S = (148 * 8) / gx;
  if (S > (rows + 63) / 64) S = (rows + 63) / 64;
  if (S < 1) S = 1;
  if (S > 1 && (!scratch || S * (long long)cols > scratch_elems)) S = scratch ? scratch_elems / cols : 1;
  if (S < 1) S = 1;
  if (S > 65535) S = 65535;
  long long chunk = (rows + S - 1) / S;
  S = (rows + chunk - 1) / chunk;
  {
    ProfScope prof("colsum_partial_kernel", st);
    EB_LAUNCH_COOP(colsum_partial_kernel, dim3(gx, (unsigned)S), 256, st, src, rows, cols, ld, rows_in, rows_out, chunk,
                                                                 S == 1 ? out : scratch);
    EB_CHECK_LAUNCH("colsum_partial_kernel");
  }
  if (S > 1) return reduce_partials_run(scratch, int(S), cols, out, st);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// GELU (exact erf, reference ACT2FN['gelu'], model/modeling_vit.py:326)
//   forward:  g = gelu(u) -> bf16 pair        backward (in place): dg <- dg * (Phi(u) + u * phi(u))
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gelu_fwd_kernel(const float4* __restrict__ u, long long n4,
                                                       __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = u[i];
    st_split4(hi, lo, i * 4, make_float4(gelu_erf(v.x), gelu_erf(v.y), gelu_erf(v.z), gelu_erf(v.w)));
  }
}
__global__ void __launch_bounds__(256) gelu_bwd_kernel(float4* __restrict__ dg, const float4* __restrict__ u, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 x = u[i];
    float4 d = dg[i];
    d.x *= gelu_grad(x.x); d.y *= gelu_grad(x.y); d.z *= gelu_grad(x.z); d.w *= gelu_grad(x.w);
    dg[i] = d;
  }
}
int gelu_fwd_run(const float* u, long long n, __nv_bfloat16* hi, __nv_bfloat16* lo, cudaStream_t st) {
  EB_REQUIRE(u && hi && n > 0 && n % 4 == 0, "gelu_fwd: bad arguments");
  ProfScope prof("gelu_fwd_kernel", st);
  EB_LAUNCH(gelu_fwd_kernel, (blocks_for(n / 4, 256, 148 * 16)), 256, st, (const float4*)u, n / 4, hi, lo);
  EB_CHECK_LAUNCH("gelu_fwd_kernel");
  return 0;
}
int gelu_bwd_run(float* dg, const float* u, long long n, cudaStream_t st) {
  EB_REQUIRE(u && dg && n > 0 && n % 4 == 0, "gelu_bwd: bad arguments");
  ProfScope prof("gelu_bwd_kernel", st);
  EB_LAUNCH(gelu_bwd_kernel, (blocks_for(n / 4, 256, 148 * 16)), 256, st, (float4*)dg, (const float4*)u, n / 4);
  EB_CHECK_LAUNCH("gelu_bwd_kernel");
  return 0;
}

// ---------------------------------------------------------------------------------------------
// LayerNorm(1024) backward.  One warp per row, persistent over the rows; the per-column sums for d gamma / d beta
// stay in registers across a warp's rows and are combined per CTA through shared memory into scratch[2][ctas][1024],
// which reduce_partials folds in a second launch.
//   xhat = (x - mean) * rstd ; g = dy * w ; dx = rstd * (g - mean(g) - xhat * mean(g * xhat))
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                            const float* __restrict__ w, long long out_rows, int rows_in,
                                                            int rows_out, float eps, float* __restrict__ dx, int accumulate,
                                                            float* __restrict__ part_w, float* __restrict__ part_b) {
  __shared__ float red[8][1024];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 aw[8], ab[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { aw[i] = make_float4(0.f, 0.f, 0.f, 0.f); ab[i] = aw[i]; }
  const float4* w4 = reinterpret_cast<const float4*>(w);
  for (long long r = blockIdx.x * 8ll + warp; r < out_rows; r += (long long)gridDim.x * 8) {
    const long long in_row = src_row(r, rows_in, rows_out);
    const float4* xs = reinterpret_cast<const float4*>(x + in_row * 1024);
    const float4* gs = reinterpret_cast<const float4*>(dy + r * 1024);
    float4 v[8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[i] = xs[lane + 32 * i];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = wsum(s) * (1.0f / 1024.0f);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float rstd = 1.0f / sqrtf(wsum(q) * (1.0f / 1024.0f) + eps);
    float4 g[8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;          // xhat
      const float4 d = gs[lane + 32 * i];
      const float4 ww = __ldg(w4 + lane + 32 * i);
      aw[i].x += d.x * v[i].x; aw[i].y += d.y * v[i].y; aw[i].z += d.z * v[i].z; aw[i].w += d.w * v[i].w;
      ab[i].x += d.x; ab[i].y += d.y; ab[i].z += d.z; ab[i].w += d.w;
      g[i] = make_float4(d.x * ww.x, d.y * ww.y, d.z * ww.z, d.w * ww.w);
      s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      s2 += (g[i].x * v[i].x + g[i].y * v[i].y) + (g[i].z * v[i].z + g[i].w * v[i].w);
    }
    s1 = wsum(s1) * (1.0f / 1024.0f);
    s2 = wsum(s2) * (1.0f / 1024.0f);
    float4* o = reinterpret_cast<float4*>(dx + in_row * 1024);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 d = make_float4(rstd * (g[i].x - s1 - v[i].x * s2), rstd * (g[i].y - s1 - v[i].y * s2),
                             rstd * (g[i].z - s1 - v[i].z * s2), rstd * (g[i].w - s1 - v[i].w * s2));
      if (accumulate) {
        const float4 p = o[lane + 32 * i];
        d.x += p.x; d.y += p.y; d.z += p.z; d.w += p.w;
      }
      o[lane + 32 * i] = d;
    }
  }
  // per-CTA combination of the 8 warps' column sums: first d gamma, then d beta
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i)
      *reinterpret_cast<float4*>(&red[warp][(lane + 32 * i) * 4]) = pass == 0 ? aw[i] : ab[i];
    __syncthreads();
    float4 acc = *reinterpret_cast<float4*>(&red[0][threadIdx.x * 4]);
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      const float4 t = *reinterpret_cast<float4*>(&red[k][threadIdx.x * 4]);
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    float* dst = (pass == 0 ? part_w : part_b) + (long long)blockIdx.x * 1024;
    *reinterpret_cast<float4*>(dst + threadIdx.x * 4) = acc;
  }
}

int layernorm_bwd_run(const float* dy, const float* x, const float* w, long long frames, int rows_in, int rows_out,
                      float eps, float* dx, int accumulate, float* dw, float* db, float* scratch, long long scratch_elems,
                      cudaStream_t st) {
  EB_REQUIRE(dy && x && w && dx && dw && db && scratch, "layernorm_bwd: null pointer");
  EB_REQUIRE(rows_out > 0 && rows_out <= rows_in, "layernorm_bwd: rows_out must be in (0, rows_in]");
  const long long out_rows = frames * rows_out;
  EB_REQUIRE(out_rows > 0, "layernorm_bwd: no rows");
  long long ctas = (out_rows + 7) / 8;
  if (ctas > 148 * 2) ctas = 148 * 2;
  EB_REQUIRE(2 * ctas * 1024 <= scratch_elems, "layernorm_bwd: scratch too small");
  float* part_w = scratch;
  float* part_b = scratch + ctas * 1024;
  {
    ProfScope prof("layernorm_bwd_kernel", st);
    EB_LAUNCH_COOP(layernorm_bwd_kernel, (unsigned)ctas, 256, st, dy, x, w, out_rows, rows_in, rows_out, eps, dx, accumulate, part_w,
                                                        part_b);
    EB_CHECK_LAUNCH("layernorm_bwd_kernel");
  }
  int rc = reduce_partials_run(part_w, int(ctas), 1024, dw, st);
  if (rc) return rc;
  return reduce_partials_run(part_b, int(ctas), 1024, db, st);
}

// ---------------------------------------------------------------------------------------------
// softmax backward with recomputation, one warp per row of 576 scaled scores:
//   P = softmax(S) ; dS = P * (dP - sum(P * dP)) * scale       -> both as bf16 pairs
// ---------------------------------------------------------------------------------------------
template <int PER_LANE2>
__global__ void __launch_bounds__(256) softmax_bwd_kernel(const float* __restrict__ S, const float* __restrict__ dP,
                                                          long long rows, float scale, __nv_bfloat16* __restrict__ p_hi,
                                                          __nv_bfloat16* __restrict__ p_lo, __nv_bfloat16* __restrict__ ds_hi,
                                                          __nv_bfloat16* __restrict__ ds_lo) {
  const int lane = threadIdx.x & 31;
  const long long r = blockIdx.x * 8ll + (threadIdx.x >> 5);
  if (r >= rows) return;
  constexpr int cols = 64 * PER_LANE2;
  const float2* s2 = reinterpret_cast<const float2*>(S + r * cols);
  const float2* d2 = reinterpret_cast<const float2*>(dP + r * cols);
  float2 v[PER_LANE2], d[PER_LANE2];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < PER_LANE2; ++i) {
    v[i] = s2[lane + 32 * i];
    d[i] = d2[lane + 32 * i];
    m = fmaxf(m, fmaxf(v[i].x, v[i].y));
  }
  m = wmax(m);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < PER_LANE2; ++i) {
    v[i].x = expf(v[i].x - m);
    v[i].y = expf(v[i].y - m);
    sum += v[i].x + v[i].y;
  }
  const float inv = 1.0f / wsum(sum);
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < PER_LANE2; ++i) {
    v[i].x *= inv; v[i].y *= inv;
    dot += v[i].x * d[i].x + v[i].y * d[i].y;
  }
  dot = wsum(dot);
#pragma unroll
  for (int i = 0; i < PER_LANE2; ++i) {
    const long long off = r * cols + (lane + 32 * i) * 2;
    uint32_t h, l;
    split_pack2(v[i].x, v[i].y, h, l);
    *reinterpret_cast<uint32_t*>(p_hi + off) = h;
    if (p_lo) *reinterpret_cast<uint32_t*>(p_lo + off) = l;
    split_pack2(v[i].x * (d[i].x - dot) * scale, v[i].y * (d[i].y - dot) * scale, h, l);
    *reinterpret_cast<uint32_t*>(ds_hi + off) = h;
    if (ds_lo) *reinterpret_cast<uint32_t*>(ds_lo + off) = l;
  }
}

int softmax_bwd_run(const float* S, const float* dP, long long rows, int cols, float scale, __nv_bfloat16* p_hi,
                    __nv_bfloat16* p_lo, __nv_bfloat16* ds_hi, __nv_bfloat16* ds_lo, cudaStream_t st) {
  EB_REQUIRE(S && dP && p_hi && ds_hi, "softmax_bwd: null pointer");
  EB_REQUIRE(cols == 576, "softmax_bwd: only 576 columns (24x24 tokens) are compiled, got %d", cols);
  EB_REQUIRE(rows > 0, "softmax_bwd: no rows");
  ProfScope prof("softmax_bwd_kernel", st);
  EB_LAUNCH_COOP(softmax_bwd_kernel<9>, (unsigned)((rows + 7) / 8), 256, st, S, dP, rows, scale, p_hi, p_lo, ds_hi, ds_lo);
  EB_CHECK_LAUNCH("softmax_bwd_kernel");
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Train-mode BatchNorm1d over the rows of y (rows x cols fp32), reference model/network_utils.py:123-142.
// Column reductions of two quantities in fp64, two deterministic stages; layout of the fp64 scratch:
// part[s][q][cols], q = 0 / 1.
//   MODE 0 (statistics): q0 = y, q1 = y*y
//   MODE 1 (backward):   z = y*scale + shift ; dz = da * (z > 0 ? 1 : 0.2) ; q0 = dz ; q1 = dz * (y - mean) * rstd
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) bn_colreduce_kernel(const float* __restrict__ y, const float* __restrict__ da,
                                                           long long rows, int cols, long long chunk,
                                                           const float* __restrict__ scale, const float* __restrict__ shift,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           double* __restrict__ part) {
  __shared__ double red[2][8][32];
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lx;
  const long long rbeg = (long long)blockIdx.y * chunk;
  long long rend = rbeg + chunk;
  if (rend > rows) rend = rows;
  double a0 = 0.0, a1 = 0.0;
  if (c < cols) {
    float sc = 0.f, sh = 0.f, mu = 0.f, rs = 0.f;
    if (MODE == 1) { sc = scale[c]; sh = shift[c]; mu = mean[c]; rs = rstd[c]; }
    for (long long r = rbeg + ly; r < rend; r += 8) {
      const float v = y[r * cols + c];
      if (MODE == 0) {
        a0 += double(v);
        a1 += double(v) * double(v);
      } else {
        const float z = v * sc + sh;
        const float dz = da[r * cols + c] * (z > 0.f ? 1.0f : 0.2f);
        a0 += double(dz);
        a1 += double(dz) * double((v - mu) * rs);
      }
    }
  }
  red[0][ly][lx] = a0;
  red[1][ly][lx] = a1;
  __syncthreads();
  if (ly == 0 && c < cols) {
#pragma unroll
    for (int i = 1; i < 8; ++i) { a0 += red[0][i][lx]; a1 += red[1][i][lx]; }
    part[((long long)blockIdx.y * 2 + 0) * cols + c] = a0;
    part[((long long)blockIdx.y * 2 + 1) * cols + c] = a1;
  }
}

__global__ void bn_stats_finalize_kernel(const double* __restrict__ part, int S, long long rows, int cols,
                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                         float* __restrict__ running_mean, float* __restrict__ running_var,
                                         long long* __restrict__ num_batches_tracked, float momentum, float eps,
                                         float* __restrict__ mean, float* __restrict__ rstd, float* __restrict__ scale,
                                         float* __restrict__ shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && num_batches_tracked) *num_batches_tracked += 1;
  if (c >= cols) return;
  double s0 = 0.0, s1 = 0.0;
  for (int s = 0; s < S; ++s) {
    s0 += part[((long long)s * 2 + 0) * cols + c];
    s1 += part[((long long)s * 2 + 1) * cols + c];
  }
  const double mu = s0 / double(rows);
  double var = s1 / double(rows) - mu * mu;
  if (var < 0.0) var = 0.0;
  const double r = 1.0 / sqrt(var + double(eps));
  const double sc = double(gamma[c]) * r;
  mean[c] = float(mu);
  rstd[c] = float(r);
  scale[c] = float(sc);
  shift[c] = float(double(beta[c]) - mu * sc);
  const double unbiased = var * double(rows) / double(rows > 1 ? rows - 1 : 1);
  running_mean[c] = float((1.0 - momentum) * double(running_mean[c]) + momentum * mu);
  running_var[c] = float((1.0 - momentum) * double(running_var[c]) + momentum * unbiased);
}

__global__ void bn_bwd_finalize_kernel(const double* __restrict__ part, int S, int cols, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  double s0 = 0.0, s1 = 0.0;
  for (int s = 0; s < S; ++s) {
    s0 += part[((long long)s * 2 + 0) * cols + c];
    s1 += part[((long long)s * 2 + 1) * cols + c];
  }
  dbeta[c] = float(s0);
  dgamma[c] = float(s1);
}

// dy = scale * (dz - dbeta/rows - xhat * dgamma/rows), in place over da
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(float* __restrict__ da, const float* __restrict__ y, long long rows,
                                                           int cols4, const float* __restrict__ scale,
                                                           const float* __restrict__ shift, const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, const float* __restrict__ dgamma,
                                                           const float* __restrict__ dbeta, float inv_rows) {
  const long long n4 = rows * cols4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c = int(i % cols4) * 4;
    const float4 v = reinterpret_cast<const float4*>(y)[i];
    float4 d = reinterpret_cast<float4*>(da)[i];
    const float4 sc = *reinterpret_cast<const float4*>(scale + c), sh = *reinterpret_cast<const float4*>(shift + c);
    const float4 mu = *reinterpret_cast<const float4*>(mean + c), rs = *reinterpret_cast<const float4*>(rstd + c);
    const float4 dg = *reinterpret_cast<const float4*>(dgamma + c), db = *reinterpret_cast<const float4*>(dbeta + c);
#define EB_BN1(f)                                                             \
    {                                                                         \
      const float z = v.f * sc.f + sh.f;                                      \
      const float dz = d.f * (z > 0.f ? 1.0f : 0.2f);                         \
      const float xh = (v.f - mu.f) * rs.f;                                   \
      d.f = sc.f * (dz - db.f * inv_rows - xh * dg.f * inv_rows);             \
    }
    EB_BN1(x) EB_BN1(y) EB_BN1(z) EB_BN1(w)
#undef EB_BN1
    reinterpret_cast<float4*>(da)[i] = d;
  }
}

// a = LeakyReLU_0.2(y * scale + shift) -> bf16 pair and / or fp32; J > 0 regroups rows frame*2J + view*J + j to
// frame*J + j with column += view*cols + col_off (reference model/net_architecture.py:699-705)
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ y, long long rows, int cols4,
                                                       const float* __restrict__ scale, const float* __restrict__ shift,
                                                       __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                       long long out_ld, float* __restrict__ out_f32, long long f32_ld, int J,
                                                       int col_off) {
  const long long n4 = rows * cols4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols4;
    const int c = int(i % cols4) * 4;
    const float4 v = reinterpret_cast<const float4*>(y)[i];
    const float4 sc = *reinterpret_cast<const float4*>(scale + c), sh = *reinterpret_cast<const float4*>(shift + c);
    float4 a = make_float4(v.x * sc.x + sh.x, v.y * sc.y + sh.y, v.z * sc.z + sh.z, v.w * sc.w + sh.w);
    a.x = a.x > 0.f ? a.x : 0.2f * a.x; a.y = a.y > 0.f ? a.y : 0.2f * a.y;
    a.z = a.z > 0.f ? a.z : 0.2f * a.z; a.w = a.w > 0.f ? a.w : 0.2f * a.w;
    long long orow = r;
    int oc = c + col_off;
    if (J > 0) {
      const long long frame = r / (2 * J);
      const int rem = int(r % (2 * J));
      orow = frame * J + rem % J;
      oc += (rem / J) * cols4 * 4;
    }
    if (hi) st_split4(hi, lo, orow * out_ld + oc, a);
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + orow * f32_ld + oc) = a;
  }
}

static int bn_reduce_plan(long long rows, int cols, long long scratch_elems, int* S, long long* chunk) {
  const int gx = (cols + 31) / 32;
  long long s = (148 * 8) / gx;
  if (s > (rows + 31) / 32) s = (rows + 31) / 32;
  if (s < 1) s = 1;
  const long long cap = scratch_elems / 2 / (2ll * cols);     // fp64 pairs inside an fp32 scratch buffer
  if (s > cap) s = cap;
  EB_REQUIRE(s >= 1, "batchnorm: scratch too small for %d columns", cols);
  *chunk = (rows + s - 1) / s;
  *S = int((rows + *chunk - 1) / *chunk);
  return 0;
}

int bn_stats_run(const float* y, long long rows, int cols, const float* gamma, const float* beta, float* running_mean,
                 float* running_var, long long* nbt, float momentum, float eps, float* mean, float* rstd, float* scale,
                 float* shift, float* scratch, long long scratch_elems, cudaStream_t st) {
  EB_REQUIRE(y && gamma && beta && running_mean && running_var && mean && rstd && scale && shift && scratch,
             "bn_stats: null pointer");
  EB_REQUIRE(rows > 0 && cols > 0, "bn_stats: empty input");
  EB_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 7) == 0, "bn_stats: scratch must be 8-byte aligned");
  int S;
  long long chunk;
  int rc = bn_reduce_plan(rows, cols, scratch_elems, &S, &chunk);
  if (rc) return rc;
  double* part = reinterpret_cast<double*>(scratch);
  {
    ProfScope prof("bn_colreduce_kernel", st);
    EB_LAUNCH_COOP(bn_colreduce_kernel<0>, dim3((cols + 31) / 32, S), 256, st, y, nullptr, rows, cols, chunk, nullptr, nullptr,
                                                                      nullptr, nullptr, part);
    EB_CHECK_LAUNCH("bn_colreduce_kernel");
  }
  ProfScope prof("bn_stats_finalize_kernel", st);
  EB_LAUNCH(bn_stats_finalize_kernel, (cols + 255) / 256, 256, st, part, S, rows, cols, gamma, beta, running_mean, running_var,
                                                              nbt, momentum, eps, mean, rstd, scale, shift);
  EB_CHECK_LAUNCH("bn_stats_finalize_kernel");
  return 0;
}

int bn_apply_run(const float* y, long long rows, int cols, const float* scale, const float* shift, __nv_bfloat16* hi,
                 __nv_bfloat16* lo, long long out_ld, float* out_f32, long long f32_ld, int J, int col_off, cudaStream_t st) {
  EB_REQUIRE(y && scale && shift && (hi || out_f32), "bn_apply: null pointer");
  EB_REQUIRE(rows > 0 && cols > 0 && cols % 4 == 0 && col_off % 4 == 0 && (!hi || out_ld % 4 == 0) &&
                 (!out_f32 || f32_ld % 4 == 0), "bn_apply: cols / ld / col_off must be multiples of 4");
  EB_REQUIRE(J == 0 || rows % (2 * J) == 0, "bn_apply: rows (%lld) must be a multiple of 2J", rows);
  ProfScope prof("bn_apply_kernel", st);
  EB_LAUNCH(bn_apply_kernel, (blocks_for(rows * (cols / 4), 256, 148 * 8)), 256, st, y, rows, cols / 4, scale, shift, hi, lo, out_ld,
                                                                              out_f32, f32_ld, J, col_off);
  EB_CHECK_LAUNCH("bn_apply_kernel");
  return 0;
}

int bn_bwd_run(float* da, const float* y, long long rows, int cols, const float* scale, const float* shift, const float* mean,
               const float* rstd, float* dgamma, float* dbeta, float* scratch, long long scratch_elems, cudaStream_t st) {
  EB_REQUIRE(da && y && scale && shift && mean && rstd && dgamma && dbeta && scratch, "bn_bwd: null pointer");
  EB_REQUIRE(rows > 0 && cols > 0 && cols % 4 == 0, "bn_bwd: cols must be a multiple of 4");
  EB_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 7) == 0, "bn_bwd: scratch must be 8-byte aligned");
  int S;
  long long chunk;
  int rc = bn_reduce_plan(rows, cols, scratch_elems, &S, &chunk);
  if (rc) return rc;
  double* part = reinterpret_cast<double*>(scratch);
  {
    ProfScope prof("bn_colreduce_kernel", st);
    EB_LAUNCH_COOP(bn_colreduce_kernel<1>, dim3((cols + 31) / 32, S), 256, st, y, da, rows, cols, chunk, scale, shift, mean, rstd, part);
    EB_CHECK_LAUNCH("bn_colreduce_kernel");
  }
  {
    ProfScope prof("bn_bwd_finalize_kernel", st);
    EB_LAUNCH(bn_bwd_finalize_kernel, (cols + 255) / 256, 256, st, part, S, cols, dgamma, dbeta);
    EB_CHECK_LAUNCH("bn_bwd_finalize_kernel");
  }
  ProfScope prof("bn_bwd_apply_kernel", st);
  EB_LAUNCH(bn_bwd_apply_kernel, (blocks_for(rows * (cols / 4), 256, 148 * 8)), 256, st, da, y, rows, cols / 4, scale, shift, mean,
                                                                                  rstd, dgamma, dbeta, 1.0f / float(rows));
  EB_CHECK_LAUNCH("bn_bwd_apply_kernel");
  return 0;
}

}  // namespace eb

using namespace eb;

// ---- C ABI (include/egotap_b200.h, "training" section) -------------------------------------------------------
extern "C" int egotap_b200_transpose_split(const float* src, long long rows, int cols, long long src_ld, int rows_in,
                                           int rows_out, void* rm_hi, void* rm_lo, long long rm_ld, void* t_hi, void* t_lo,
                                           long long t_ld, long long pad_rows, const float* gelu_u, float* colsum_out,
                                           float* scratch, long long scratch_elems, void* stream) {
  return transpose_split_run(src, rows, cols, src_ld, rows_in, rows_out, (__nv_bfloat16*)rm_hi, (__nv_bfloat16*)rm_lo, rm_ld,
                             (__nv_bfloat16*)t_hi, (__nv_bfloat16*)t_lo, t_ld, pad_rows, gelu_u, colsum_out, scratch,
                             scratch_elems, (cudaStream_t)stream);
}
extern "C" int egotap_b200_transpose_bf16(const void* s_hi, const void* s_lo, long long rows, int cols, long long s_ld,
                                          int g0_count, long long s_g0_stride, int g1_count, long long s_g1_stride, void* d_hi,
                                          void* d_lo, long long d_ld, long long d_g0_stride, long long d_g1_stride,
                                          long long pad_rows, void* stream) {
  return transpose_bf16_run(s_hi, s_lo, rows, cols, s_ld, g0_count, s_g0_stride, g1_count, s_g1_stride, d_hi,
                            s_lo ? d_lo : nullptr, d_ld, d_g0_stride, d_g1_stride, pad_rows, (cudaStream_t)stream);
}
extern "C" int egotap_b200_colsum(const float* src, long long rows, int cols, long long ld, int rows_in, int rows_out,
                                  float* out, float* scratch, long long scratch_elems, void* stream) {
  return colsum_run(src, rows, cols, ld, rows_in, rows_out, out, scratch, scratch_elems, (cudaStream_t)stream);
}
extern "C" int egotap_b200_reduce_partials(const float* partials, int G, long long n, float* out, void* stream) {
  return reduce_partials_run(partials, G, n, out, (cudaStream_t)stream);
}
extern "C" int egotap_b200_gelu_fwd(const float* u, long long n, void* out_hi, void* out_lo, void* stream) {
  return gelu_fwd_run(u, n, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, (cudaStream_t)stream);
}
extern "C" int egotap_b200_gelu_bwd(float* dg, const float* u, long long n, void* stream) {
  return gelu_bwd_run(dg, u, n, (cudaStream_t)stream);
}
extern "C" int egotap_b200_layernorm_bwd(const float* dy, const float* x, const float* w, long long frames, int rows_in,
                                         int rows_out, float eps, float* dx, int accumulate, float* dw, float* db,
                                         float* scratch, long long scratch_elems, void* stream) {
  return layernorm_bwd_run(dy, x, w, frames, rows_in, rows_out, eps, dx, accumulate, dw, db, scratch, scratch_elems,
                           (cudaStream_t)stream);
}
extern "C" int egotap_b200_softmax_bwd(const float* S, const float* dP, long long rows, int cols, float scale, void* p_hi,
                                       void* p_lo, void* ds_hi, void* ds_lo, void* stream) {
  return softmax_bwd_run(S, dP, rows, cols, scale, (__nv_bfloat16*)p_hi, (__nv_bfloat16*)p_lo, (__nv_bfloat16*)ds_hi,
                         (__nv_bfloat16*)ds_lo, (cudaStream_t)stream);
}
extern "C" int egotap_b200_bn_stats(const float* y, long long rows, int cols, const float* gamma, const float* beta,
                                    float* running_mean, float* running_var, long long* num_batches_tracked, float momentum,
                                    float eps, float* mean, float* rstd, float* scale, float* shift, float* scratch,
                                    long long scratch_elems, void* stream) {
  return bn_stats_run(y, rows, cols, gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps, mean, rstd,
                      scale, shift, scratch, scratch_elems, (cudaStream_t)stream);
}
extern "C" int egotap_b200_bn_apply(const float* y, long long rows, int cols, const float* scale, const float* shift,
                                    void* out_hi, void* out_lo, long long out_ld, float* out_f32, long long f32_ld, int J,
                                    int col_off, void* stream) {
  return bn_apply_run(y, rows, cols, scale, shift, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, out_ld, out_f32, f32_ld, J,
                      col_off, (cudaStream_t)stream);
}
extern "C" int egotap_b200_bn_bwd(float* da, const float* y, long long rows, int cols, const float* scale, const float* shift,
                                  const float* mean, const float* rstd, float* dgamma, float* dbeta, float* scratch,
                                  long long scratch_elems, void* stream) {
  return bn_bwd_run(da, y, rows, cols, scale, shift, mean, rstd, dgamma, dbeta, scratch, scratch_elems, (cudaStream_t)stream);
}

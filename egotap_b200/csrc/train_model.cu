// Training-step kernels tied to the model's own layers: propagation-unit cell (forward with saved state, backward),
// bridge-gate backward, regression-head backward, embedding-parameter gradients, the pose loss and AdamW.
// Host orchestration: egotap_b200/training.py; semantics of every entry: oracle/op_oracle.py (test infrastructure).
#include "host_util.cuh"
#include "numeric.cuh"

namespace eb {

// kernels.cu / train_ops.cu
int split2d_run(const float*, long long, long long, long long, __nv_bfloat16*, __nv_bfloat16*, long long, cudaStream_t);
int fill_dummy_run(float*, const float*, int, int, int, cudaStream_t, __nv_bfloat16* = nullptr, __nv_bfloat16* = nullptr, float2* = nullptr);
int pos_permute_run(const float*, const float*, int, int, float*, float*, cudaStream_t);
int pu_bridge_gate_run(const float*, int, int, const float*, int, int, long long, __nv_bfloat16*, __nv_bfloat16*,
                       cudaStream_t);
int vec_add3_run(const float*, const float*, const float*, float*, int, cudaStream_t);
int reduce_partials_run(const float*, int, long long, float*, cudaStream_t);

namespace {

__device__ __forceinline__ float sigm(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ void st_split4(__nv_bfloat16* hi, __nv_bfloat16* lo, long long off, float4 v) {
  uint32_t h0, h1, l0, l1;
  split_pack2(v.x, v.y, h0, l0);
  split_pack2(v.z, v.w, h1, l1);
  *reinterpret_cast<uint2*>(hi + off) = make_uint2(h0, h1);
  if (lo) *reinterpret_cast<uint2*>(lo + off) = make_uint2(l0, l1);
}
inline unsigned blocks_for(long long n, int per_block, long long cap) {
  long long b = (n + per_block - 1) / per_block;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return unsigned(b);
}
constexpr int PH = 512;   // propagation-unit hidden size

}  // namespace

// ---------------------------------------------------------------------------------------------
// Propagation-unit cell, one joint step t, with the state kept for the backward
// (reference model/custom_cells.py:109-120, gate order forget, in, cell, out):
//   gates = G[b*g_rs + t*g_ts + :2048] ; c_prev = C[b*J + t - 1] (0 at t = 0)
//   c = c_prev * s(fg) + s(ig) * tanh(cg) -> C[b*J + t] ;  h = s(og) * tanh(c) -> H[b*J + t] (+ bf16 pair)
//   hg[b*J + t + 1] = s(F[b*f_rs + (t+1)*f_ts + :512]) * h        (the next step's recurrent GEMM operand, :101)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pu_cell_fwd_kernel(const float* __restrict__ G, long long g_rs, long long g_ts,
                                                          const float* __restrict__ F, long long f_rs, long long f_ts,
                                                          float* __restrict__ C, float* __restrict__ H,
                                                          __nv_bfloat16* __restrict__ h_hi, __nv_bfloat16* __restrict__ h_lo,
                                                          __nv_bfloat16* __restrict__ hg_hi, __nv_bfloat16* __restrict__ hg_lo,
                                                          int t, int J, long long B) {
  const long long n4 = B * (PH / 4);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / (PH / 4);
    const int u = int(i % (PH / 4)) * 4;
    const float* g = G + b * g_rs + t * g_ts + u;
    const float4 fg = *reinterpret_cast<const float4*>(g);
    const float4 ig = *reinterpret_cast<const float4*>(g + PH);
    const float4 cg = *reinterpret_cast<const float4*>(g + 2 * PH);
    const float4 og = *reinterpret_cast<const float4*>(g + 3 * PH);
    const long long row = (b * J + t) * PH + u;
    float4 cv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t > 0) cv = *reinterpret_cast<const float4*>(C + row - PH);
    cv.x = cv.x * sigm(fg.x) + sigm(ig.x) * tanhf(cg.x);
    cv.y = cv.y * sigm(fg.y) + sigm(ig.y) * tanhf(cg.y);
    cv.z = cv.z * sigm(fg.z) + sigm(ig.z) * tanhf(cg.z);
    cv.w = cv.w * sigm(fg.w) + sigm(ig.w) * tanhf(cg.w);
    *reinterpret_cast<float4*>(C + row) = cv;
    const float4 hv = make_float4(sigm(og.x) * tanhf(cv.x), sigm(og.y) * tanhf(cv.y), sigm(og.z) * tanhf(cv.z),
                                  sigm(og.w) * tanhf(cv.w));
    *reinterpret_cast<float4*>(H + row) = hv;
    if (h_hi) st_split4(h_hi, h_lo, row, hv);
    if (t + 1 < J) {
      const float4 f = *reinterpret_cast<const float4*>(F + b * f_rs + (t + 1) * f_ts + u);
      st_split4(hg_hi, hg_lo, row + PH,
                make_float4(sigm(f.x) * hv.x, sigm(f.y) * hv.y, sigm(f.z) * hv.z, sigm(f.w) * hv.w));
    }
  }
}

int pu_cell_fwd_run(const float* G, long long g_rs, long long g_ts, const float* F, long long f_rs, long long f_ts, float* C,
                    float* H, __nv_bfloat16* h_hi, __nv_bfloat16* h_lo, __nv_bfloat16* hg_hi, __nv_bfloat16* hg_lo, int t,
                    int J, long long B, cudaStream_t st) {
  EB_REQUIRE(G && F && C && H && hg_hi, "pu_cell_fwd: null pointer");
  EB_REQUIRE(t >= 0 && t < J && B > 0, "pu_cell_fwd: bad step %d of %d", t, J);
  EB_REQUIRE(g_rs % 4 == 0 && g_ts % 4 == 0 && f_rs % 4 == 0 && f_ts % 4 == 0, "pu_cell_fwd: strides must be multiples of 4");
  ProfScope prof("pu_cell_fwd_kernel", st);
  EB_LAUNCH(pu_cell_fwd_kernel, (blocks_for(B * (PH / 4), 256, 148 * 8)), 256, st, G, g_rs, g_ts, F, f_rs, f_ts, C, H, h_hi, h_lo,
                                                                            hg_hi, hg_lo, t, J, B);
  EB_CHECK_LAUNCH("pu_cell_fwd_kernel");
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Backward of one cell step (BPTT walks t = J-1 .. 0):
//   dh = dOut[b*J+t] + (t+1 < J ? dhg[b] * s(F[b,t+1]) : 0) ;  dF[b,t+1] = dhg[b] * h_t * s'(F[b,t+1]) ; dF[b,0] = 0
//   dc_tot = (t+1 < J ? dc[b] : 0) + dh * o * (1 - tanh(c_t)^2)
//   dgates = [dc_tot*c_prev*f', dc_tot*tanh(cg)*i', dc_tot*i*(1-tanh(cg)^2), dh*tanh(c_t)*o']
//   dc[b] <- dc_tot * f ;  dgates -> dG[b*dg_rs + t*dg_ts + :2048] (fp32) and dgp[b][:2048] (bf16 pair, GEMM operand)
// ---------------------------------------------------------------------------------------------
struct Gate4 { float4 s, d; };   // sigmoid / tanh value and its derivative
__device__ __forceinline__ Gate4 sig4(float4 x) {
  Gate4 r;
  r.s = make_float4(sigm(x.x), sigm(x.y), sigm(x.z), sigm(x.w));
  r.d = make_float4(r.s.x * (1.f - r.s.x), r.s.y * (1.f - r.s.y), r.s.z * (1.f - r.s.z), r.s.w * (1.f - r.s.w));
  return r;
}
__device__ __forceinline__ Gate4 tanh4(float4 x) {
  Gate4 r;
  r.s = make_float4(tanhf(x.x), tanhf(x.y), tanhf(x.z), tanhf(x.w));
  r.d = make_float4(1.f - r.s.x * r.s.x, 1.f - r.s.y * r.s.y, 1.f - r.s.z * r.s.z, 1.f - r.s.w * r.s.w);
  return r;
}
__device__ __forceinline__ float4 mul4(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

__global__ void __launch_bounds__(256) pu_cell_bwd_kernel(const float* __restrict__ G, long long g_rs, long long g_ts,
                                                          const float* __restrict__ F, long long f_rs, long long f_ts,
                                                          const float* __restrict__ C, const float* __restrict__ H,
                                                          const float* __restrict__ dOut, const float* __restrict__ dhg,
                                                          float* __restrict__ dc, float* __restrict__ dG, long long dg_rs,
                                                          long long dg_ts, float* __restrict__ dF, long long df_rs,
                                                          long long df_ts, __nv_bfloat16* __restrict__ dgp_hi,
                                                          __nv_bfloat16* __restrict__ dgp_lo, int t, int J, long long B) {
  const long long n4 = B * (PH / 4);
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / (PH / 4);
    const int u = int(i % (PH / 4)) * 4;
    const float* g = G + b * g_rs + t * g_ts + u;
    const Gate4 fg = sig4(*reinterpret_cast<const float4*>(g));
    const Gate4 ig = sig4(*reinterpret_cast<const float4*>(g + PH));
    const Gate4 cg = tanh4(*reinterpret_cast<const float4*>(g + 2 * PH));
    const Gate4 og = sig4(*reinterpret_cast<const float4*>(g + 3 * PH));
    const long long row = (b * J + t) * PH + u;
    const Gate4 tc = tanh4(*reinterpret_cast<const float4*>(C + row));
    const float4 c_prev = t > 0 ? *reinterpret_cast<const float4*>(C + row - PH) : zero4;
    const float4 h_t = *reinterpret_cast<const float4*>(H + row);
    float4 dh = *reinterpret_cast<const float4*>(dOut + row);
    float4 dc_tot = zero4;
    if (t + 1 < J) {
      const Gate4 sf = sig4(*reinterpret_cast<const float4*>(F + b * f_rs + (t + 1) * f_ts + u));
      const float4 dv = *reinterpret_cast<const float4*>(dhg + b * PH + u);
      dh = add4(dh, mul4(dv, sf.s));
      *reinterpret_cast<float4*>(dF + b * df_rs + (t + 1) * df_ts + u) = mul4(mul4(dv, h_t), sf.d);
      dc_tot = *reinterpret_cast<const float4*>(dc + b * PH + u);
    }
    if (t == 0) *reinterpret_cast<float4*>(dF + b * df_rs + u) = zero4;
    dc_tot = add4(dc_tot, mul4(mul4(dh, og.s), tc.d));
    const float4 dfg = mul4(mul4(dc_tot, c_prev), fg.d);
    const float4 dig = mul4(mul4(dc_tot, cg.s), ig.d);
    const float4 dcg = mul4(mul4(dc_tot, ig.s), cg.d);
    const float4 dog = mul4(mul4(dh, tc.s), og.d);
    *reinterpret_cast<float4*>(dc + b * PH + u) = mul4(dc_tot, fg.s);
    float* o = dG + b * dg_rs + t * dg_ts + u;
    *reinterpret_cast<float4*>(o) = dfg;
    *reinterpret_cast<float4*>(o + PH) = dig;
    *reinterpret_cast<float4*>(o + 2 * PH) = dcg;
    *reinterpret_cast<float4*>(o + 3 * PH) = dog;
    const long long po = b * 4 * PH + u;
    st_split4(dgp_hi, dgp_lo, po, dfg);
    st_split4(dgp_hi, dgp_lo, po + PH, dig);
    st_split4(dgp_hi, dgp_lo, po + 2 * PH, dcg);
    st_split4(dgp_hi, dgp_lo, po + 3 * PH, dog);
  }
}

int pu_cell_bwd_run(const float* G, long long g_rs, long long g_ts, const float* F, long long f_rs, long long f_ts,
                    const float* C, const float* H, const float* dOut, const float* dhg, float* dc, float* dG, long long dg_rs,
                    long long dg_ts, float* dF, long long df_rs, long long df_ts, __nv_bfloat16* dgp_hi, __nv_bfloat16* dgp_lo,
                    int t, int J, long long B, cudaStream_t st) {
  EB_REQUIRE(G && F && C && H && dOut && dhg && dc && dG && dF && dgp_hi, "pu_cell_bwd: null pointer");
  EB_REQUIRE(t >= 0 && t < J && B > 0, "pu_cell_bwd: bad step %d of %d", t, J);
  EB_REQUIRE(g_rs % 4 == 0 && g_ts % 4 == 0 && f_rs % 4 == 0 && f_ts % 4 == 0 && dg_rs % 4 == 0 && dg_ts % 4 == 0 &&
                 df_rs % 4 == 0 && df_ts % 4 == 0, "pu_cell_bwd: strides must be multiples of 4");
  ProfScope prof("pu_cell_bwd_kernel", st);
  EB_LAUNCH(pu_cell_bwd_kernel, (blocks_for(B * (PH / 4), 256, 148 * 8)), 256, st, G, g_rs, g_ts, F, f_rs, f_ts, C, H, dOut, dhg, dc,
                                                                            dG, dg_rs, dg_ts, dF, df_rs, df_ts, dgp_hi, dgp_lo,
                                                                            t, J, B);
  EB_CHECK_LAUNCH("pu_cell_bwd_kernel");
  return 0;
}

// ---------------------------------------------------------------------------------------------
// bridge-gate backward (forward: b' = s(F0[:, f_col:f_col+X]) * E[:, X:2X], kernels.cu pu_bridge_gate_kernel):
//   in: dE[r][X + c] = d b'   out: dE[r][X + c] = d b' * s(Fb) ; dF[r][f_col + c] = d b' * bridge * s'(Fb)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pu_bridge_gate_bwd_kernel(float* __restrict__ dE, long long e_ld,
                                                                 const float* __restrict__ F0, long long f_ld, int f_col,
                                                                 const float* __restrict__ E, int X, long long rows,
                                                                 float* __restrict__ dF, long long df_ld) {
  const int x4 = X / 4;
  const long long n4 = rows * x4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / x4;
    const int c = int(i % x4) * 4;
    float4* de = reinterpret_cast<float4*>(dE + r * e_ld + X + c);
    const float4 db = *de;
    const float4 f = *reinterpret_cast<const float4*>(F0 + r * f_ld + f_col + c);
    const float4 br = *reinterpret_cast<const float4*>(E + r * e_ld + X + c);
    const float4 s = make_float4(sigm(f.x), sigm(f.y), sigm(f.z), sigm(f.w));
    *de = make_float4(db.x * s.x, db.y * s.y, db.z * s.z, db.w * s.w);
    *reinterpret_cast<float4*>(dF + r * df_ld + f_col + c) =
        make_float4(db.x * br.x * s.x * (1.f - s.x), db.y * br.y * s.y * (1.f - s.y), db.z * br.z * s.z * (1.f - s.z),
                    db.w * br.w * s.w * (1.f - s.w));
  }
}

int pu_bridge_gate_bwd_run(float* dE, long long e_ld, const float* F0, long long f_ld, int f_col, const float* E, int X,
                           long long rows, float* dF, long long df_ld, cudaStream_t st) {
  EB_REQUIRE(dE && F0 && E && dF && rows > 0, "pu_bridge_gate_bwd: bad arguments");
  EB_REQUIRE(X % 4 == 0 && e_ld % 4 == 0 && f_ld % 4 == 0 && f_col % 4 == 0 && df_ld % 4 == 0,
             "pu_bridge_gate_bwd: sizes must be multiples of 4");
  ProfScope prof("pu_bridge_gate_bwd_kernel", st);
  EB_LAUNCH(pu_bridge_gate_bwd_kernel, (blocks_for(rows * (X / 4), 256, 148 * 8)), 256, st, dE, e_ld, F0, f_ld, f_col, E, X, rows, dF,
                                                                                     df_ld);
  EB_CHECK_LAUNCH("pu_bridge_gate_bwd_kernel");
  return 0;
}

// out[f*2J + v*J + j][c] = dE[f*J + j][col_off + v*cols + c]   (inverse of bn_apply's regrouped store)
__global__ void __launch_bounds__(256) regroup_gather_kernel(const float* __restrict__ dE, long long e_ld, int col_off,
                                                             long long rows, int J, int cols4, float* __restrict__ out) {
  const long long n4 = rows * cols4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols4;
    const int c = int(i % cols4) * 4;
    const long long frame = r / (2 * J);
    const int rem = int(r % (2 * J));
    reinterpret_cast<float4*>(out)[i] =
        *reinterpret_cast<const float4*>(dE + (frame * J + rem % J) * e_ld + col_off + (rem / J) * cols4 * 4 + c);
  }
}

int regroup_gather_run(const float* dE, long long e_ld, int col_off, long long frames, int J, int cols, float* out,
                       cudaStream_t st) {
  EB_REQUIRE(dE && out && frames > 0 && J > 0, "regroup_gather: bad arguments");
  EB_REQUIRE(cols % 4 == 0 && e_ld % 4 == 0 && col_off % 4 == 0, "regroup_gather: sizes must be multiples of 4");
  ProfScope prof("regroup_gather_kernel", st);
  EB_LAUNCH(regroup_gather_kernel, (blocks_for(frames * 2 * J * (cols / 4), 256, 148 * 8)), 256, st, dE, e_ld, col_off, frames * 2 * J,
                                                                                             J, cols / 4, out);
  EB_CHECK_LAUNCH("regroup_gather_kernel");
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Regression-head backward (forward: kernels.cu head_kernel; reference model/net_architecture.py:732-751)
//   d_in[b,j,:768] = sum_k dpose[b,j,k] * Wp[k,:]  (+ sum_k6 do[b,k6] * Wg[k6, j*512 + :] on the skel part)
//   do[b,:3] = sum_j dpose[b,j,:] ; do[b,3:6] = dpose[b,J,:]
//   writes dE[r][:256] = d_in[:, :256], dE[r][256:512] = 0, dSkel[r][:512] = d_in[:, 256:], do -> scratch[b][8]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) head_bwd_data_kernel(const float* __restrict__ dpose, const float* __restrict__ Wp,
                                                            const float* __restrict__ Wg, int J, float* __restrict__ dE,
                                                            long long de_ld, float* __restrict__ dSkel,
                                                            float* __restrict__ dout) {
  __shared__ float s_dp[32 * 3];
  __shared__ float s_do[8];
  const long long b = blockIdx.x;
  const int nj = Wg ? J + 1 : J;
  for (int i = threadIdx.x; i < nj * 3; i += blockDim.x) s_dp[i] = dpose[b * nj * 3 + i];
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = 0.f;
    if (Wg) {
      if (threadIdx.x < 3) {
        for (int j = 0; j < J; ++j) v += s_dp[j * 3 + threadIdx.x];
      } else if (threadIdx.x < 6) {
        v = s_dp[J * 3 + threadIdx.x - 3];
      }
    }
    s_do[threadIdx.x] = v;
    dout[b * 8 + threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < J * 768; i += blockDim.x) {
    const int j = i / 768, c = i % 768;
    float v = s_dp[j * 3] * __ldg(Wp + c) + s_dp[j * 3 + 1] * __ldg(Wp + 768 + c) + s_dp[j * 3 + 2] * __ldg(Wp + 1536 + c);
    const long long r = b * J + j;
    if (c < 256) {
      dE[r * de_ld + c] = v;
      dE[r * de_ld + 256 + c] = 0.f;
    } else {
      if (Wg) {
        const long long col = (long long)j * 512 + c - 256;
#pragma unroll
        for (int k = 0; k < 6; ++k) v += s_do[k] * __ldg(Wg + (long long)k * J * 512 + col);
      }
      dSkel[r * 512 + c - 256] = v;
    }
  }
}

// dW[k][c] partial over a slice of the rows: in[row][c] = c < 256 ? e[row*e_ld + c] : skel[row*512 + c - 256]
// (NK = 3, cols = 768, weight = dpose[b,j,k]) or in[b][c] = skel[b*cols + c] (NK = 6, weight = do[b][k])
template <int NK>
__global__ void __launch_bounds__(256) head_bwd_weight_kernel(const float* __restrict__ wsrc, int w_stride, int J, int nj,
                                                              const float* __restrict__ e, long long e_ld,
                                                              const float* __restrict__ skel, long long rows, int cols,
                                                              long long chunk, float* __restrict__ part) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const long long rbeg = (long long)blockIdx.y * chunk;
  long long rend = rbeg + chunk;
  if (rend > rows) rend = rows;
  float acc[NK];
#pragma unroll
  for (int k = 0; k < NK; ++k) acc[k] = 0.f;
  for (long long r = rbeg; r < rend; ++r) {
    float x;
    const float* w;
    if (NK == 3) {
      x = c < 256 ? e[r * e_ld + c] : skel[r * 512 + c - 256];
      w = wsrc + ((r / J) * nj + r % J) * 3;          // dpose[b][j][:]
    } else {
      x = skel[r * cols + c];
      w = wsrc + r * w_stride;                        // do[b][:]
    }
#pragma unroll
    for (int k = 0; k < NK; ++k) acc[k] += __ldg(w + k) * x;
  }
#pragma unroll
  for (int k = 0; k < NK; ++k) part[((long long)blockIdx.y * NK + k) * cols + c] = acc[k];
}

// dbp[k] = sum_{b, j < J} dpose[b,j,k] ; dbg[k6] = sum_b do[b,k6]   (one CTA, deterministic tree)
__global__ void __launch_bounds__(256) head_bwd_bias_kernel(const float* __restrict__ dpose, const float* __restrict__ dout,
                                                            long long B, int J, int nj, float* __restrict__ dbp,
                                                            float* __restrict__ dbg) {
  __shared__ float red[9][256];
  float a[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) a[k] = 0.f;
  for (long long r = threadIdx.x; r < B * J; r += blockDim.x) {
    const float* p = dpose + ((r / J) * nj + r % J) * 3;
    a[0] += p[0]; a[1] += p[1]; a[2] += p[2];
  }
  if (dbg)
    for (long long b = threadIdx.x; b < B; b += blockDim.x) {
#pragma unroll
      for (int k = 0; k < 6; ++k) a[3 + k] += dout[b * 8 + k];
    }
#pragma unroll
  for (int k = 0; k < 9; ++k) red[k][threadIdx.x] = a[k];
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
#pragma unroll
      for (int k = 0; k < 9; ++k) red[k][threadIdx.x] += red[k][threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x < 3) dbp[threadIdx.x] = red[threadIdx.x][0];
  if (dbg && threadIdx.x >= 3 && threadIdx.x < 9) dbg[threadIdx.x - 3] = red[threadIdx.x][0];
}

int head_bwd_run(const float* dpose, const float* e, long long e_ld, const float* skel, const float* Wp, const float* Wg,
                 long long B, int J, float* dE, long long de_ld, float* dSkel, float* dWp, float* dbp, float* dWg, float* dbg,
                 float* scratch, long long scratch_elems, cudaStream_t st) {
  EB_REQUIRE(dpose && e && skel && Wp && dE && dSkel && dWp && dbp && scratch, "head_bwd: null pointer");
  EB_REQUIRE(!Wg || (dWg && dbg), "head_bwd: global head needs dWg / dbg");
  EB_REQUIRE(B > 0 && J > 0 && J < 32, "head_bwd: bad sizes");
  const int nj = Wg ? J + 1 : J;
  float* dout = scratch;                                  // [B][8]
  float* part = scratch + ((B * 8 + 63) / 64) * 64;
  const long long part_elems = scratch_elems - (part - scratch);
  {
    ProfScope prof("head_bwd_data_kernel", st);
    EB_LAUNCH_COOP(head_bwd_data_kernel, (unsigned)B, 256, st, dpose, Wp, Wg, J, dE, de_ld, dSkel, dout);
    EB_CHECK_LAUNCH("head_bwd_data_kernel");
  }
  {  // dWp: reduce over the B*J rows
    const long long rows = B * J;
    long long S = rows / 16 > 0 ? rows / 16 : 1;
    if (S > 148) S = 148;
    EB_REQUIRE(S * 3 * 768 <= part_elems, "head_bwd: scratch too small");
    long long chunk = (rows + S - 1) / S;
    S = (rows + chunk - 1) / chunk;
    {
      ProfScope prof("head_bwd_weight_kernel", st);
      EB_LAUNCH(head_bwd_weight_kernel<3>, dim3(3, (unsigned)S), 256, st, dpose, 0, J, nj, e, e_ld, skel, rows, 768, chunk, part);
      EB_CHECK_LAUNCH("head_bwd_weight_kernel");
    }
    int rc = reduce_partials_run(part, int(S), 3 * 768, dWp, st);
    if (rc) return rc;
  }
  if (Wg) {  // dWg: reduce over the B frames
    const int cols = J * 512;
    long long S = B / 8 > 0 ? B / 8 : 1;
    if (S > 32) S = 32;
    EB_REQUIRE(S * 6 * cols <= part_elems, "head_bwd: scratch too small");
    long long chunk = (B + S - 1) / S;
    S = (B + chunk - 1) / chunk;
    {
      ProfScope prof("head_bwd_weight_kernel", st);
      EB_LAUNCH(head_bwd_weight_kernel<6>, dim3((cols + 255) / 256, (unsigned)S), 256, st, dout, 8, J, nj, nullptr, 0, skel, B, cols,
                                                                                      chunk, part);
      EB_CHECK_LAUNCH("head_bwd_weight_kernel");
    }
    int rc = reduce_partials_run(part, int(S), 6ll * cols, dWg, st);
    if (rc) return rc;
  }
  ProfScope prof("head_bwd_bias_kernel", st);
  EB_LAUNCH_COOP(head_bwd_bias_kernel, 1, 256, st, dpose, dout, B, J, nj, dbp, Wg ? dbg : nullptr);
  EB_CHECK_LAUNCH("head_bwd_bias_kernel");
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Embedding parameters: dpos_perm (576 x 1024, heatmap-major tokens, already summed over frames) ->
// dpos in the mosaic's raster order (inverse of pos_permute_kernel) and dmask = sum over the dummy tokens
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) embed_grads_kernel(const float* __restrict__ dpp, int grid, int n_hm,
                                                          float* __restrict__ dpos, float* __restrict__ dmask) {
  const int tokens = grid * grid * 16;
  if (int(blockIdx.x) < tokens) {
    const int tp = blockIdx.x;
    const int n = tp >> 4, pr = (tp >> 2) & 3, pc = tp & 3;
    const int raster = ((n / grid) * 4 + pr) * (grid * 4) + (n % grid) * 4 + pc;
    reinterpret_cast<float4*>(dpos)[raster * 256 + threadIdx.x] = reinterpret_cast<const float4*>(dpp)[tp * 256 + threadIdx.x];
  } else {
    const int c = (blockIdx.x - tokens) * 256 + threadIdx.x;      // 4 CTAs cover the 1024 columns
    float acc = 0.f;
    for (int tp = n_hm * 16; tp < tokens; ++tp) acc += dpp[tp * 1024 + c];
    dmask[c] = acc;
  }
}

int embed_grads_run(const float* dpp, int grid, int n_hm, float* dpos, float* dmask, cudaStream_t st) {
  EB_REQUIRE(dpp && dpos && dmask, "embed_grads: null pointer");
  EB_REQUIRE(grid * grid * 16 == 576 && n_hm > 0 && n_hm <= grid * grid, "embed_grads: bad geometry");
  ProfScope prof("embed_grads_kernel", st);
  EB_LAUNCH(embed_grads_kernel, grid * grid * 16 + 4, 256, st, dpp, grid, n_hm, dpos, dmask);
  EB_CHECK_LAUNCH("embed_grads_kernel");
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Pose loss (reference model/egotap_autoencoder_model.py:284-296, utils/loss.py:44-85):
//   total = lm * mean_{b,j} |gt - pred| + lc * lm * mean_b sum_bones cos(pred bone, gt bone)
// one CTA of 64 threads per frame computes the frame's two sums and d total / d pred; a second kernel folds the
// frames in a fixed order.  drop_first (EgoCap): a zero root joint is prepended and the first bone is not counted.
// ---------------------------------------------------------------------------------------------
struct LossCfg {
  int nj, drop_first, n_ext;          // n_ext = joints incl. the optional prepended root
  int parents[24];
  float w_mpjpe, w_cos;               // lm / (B*nj), lc*lm / B
};

__global__ void __launch_bounds__(64) pose_loss_frame_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                             LossCfg cfg, float* __restrict__ frame_sums,
                                                             float* __restrict__ dpose) {
  __shared__ float p[24 * 3], g[24 * 3], d[24 * 3];
  __shared__ float s_mp[32], s_cs[32];
  const long long b = blockIdx.x;
  const int t = threadIdx.x, shift = cfg.drop_first ? 1 : 0;
  for (int i = t; i < cfg.n_ext * 3; i += blockDim.x) {
    const int j = i / 3 - shift;
    p[i] = j >= 0 ? pred[(b * cfg.nj + j) * 3 + i % 3] : 0.f;
    g[i] = j >= 0 ? gt[(b * cfg.nj + j) * 3 + i % 3] : 0.f;
    d[i] = 0.f;
  }
  if (t < 32) { s_mp[t] = 0.f; s_cs[t] = 0.f; }
  __syncthreads();
  if (t < cfg.nj) {                                    // MPJPE term of joint t
    const int e = t + shift;
    const float dx = p[e * 3] - g[e * 3], dy = p[e * 3 + 1] - g[e * 3 + 1], dz = p[e * 3 + 2] - g[e * 3 + 2];
    const float n = sqrtf(dx * dx + dy * dy + dz * dz);
    s_mp[t] = n;
    if (n > 0.f) {
      const float w = cfg.w_mpjpe / n;
      atomicAdd(&d[e * 3], w * dx); atomicAdd(&d[e * 3 + 1], w * dy); atomicAdd(&d[e * 3 + 2], w * dz);
    }
  }
  if (t >= 1 + shift && t < cfg.n_ext) {               // bone t (extended index) -> parent
    const int e = t, q = cfg.parents[t];
    const float ax = p[e * 3] - p[q * 3], ay = p[e * 3 + 1] - p[q * 3 + 1], az = p[e * 3 + 2] - p[q * 3 + 2];
    const float bx = g[e * 3] - g[q * 3], by = g[e * 3 + 1] - g[q * 3 + 1], bz = g[e * 3 + 2] - g[q * 3 + 2];
    const float eps = 1e-8f;
    const float na_raw = sqrtf(ax * ax + ay * ay + az * az), nb_raw = sqrtf(bx * bx + by * by + bz * bz);
    const float na = fmaxf(na_raw, eps), nb = fmaxf(nb_raw, eps);
    const float dot = ax * bx + ay * by + az * bz;
    const float cs = dot / (na * nb);
    s_cs[t] = cs;
    // d cos / d a = b / (na nb) - cos * a / na^2   (the second term vanishes while the norm is clamped)
    const float k1 = cfg.w_cos / (na * nb), k2 = na_raw > eps ? cfg.w_cos * cs / (na * na) : 0.f;
    const float gx = k1 * bx - k2 * ax, gy = k1 * by - k2 * ay, gz = k1 * bz - k2 * az;
    atomicAdd(&d[e * 3], gx); atomicAdd(&d[e * 3 + 1], gy); atomicAdd(&d[e * 3 + 2], gz);
    atomicAdd(&d[q * 3], -gx); atomicAdd(&d[q * 3 + 1], -gy); atomicAdd(&d[q * 3 + 2], -gz);
  }
  __syncthreads();
  for (int i = t; i < cfg.nj * 3; i += blockDim.x) dpose[b * cfg.nj * 3 + i] = d[i + shift * 3];
  if (t == 0) {
    float mp = 0.f, cs = 0.f;
    for (int i = 0; i < 32; ++i) { mp += s_mp[i]; cs += s_cs[i]; }
    frame_sums[b * 2] = mp;
    frame_sums[b * 2 + 1] = cs;
  }
}

__global__ void __launch_bounds__(256) pose_loss_reduce_kernel(const float* __restrict__ frame_sums, long long B, float w_mpjpe,
                                                               float w_cos, float* __restrict__ loss) {
  __shared__ double red[2][256];
  double a = 0.0, c = 0.0;
  for (long long b = threadIdx.x; b < B; b += blockDim.x) { a += frame_sums[b * 2]; c += frame_sums[b * 2 + 1]; }
  red[0][threadIdx.x] = a;
  red[1][threadIdx.x] = c;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) { red[0][threadIdx.x] += red[0][threadIdx.x + s]; red[1][threadIdx.x] += red[1][threadIdx.x + s]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float mp = float(red[0][0] * w_mpjpe), cs = float(red[1][0] * w_cos);
    loss[0] = mp + cs;
    loss[1] = mp;
    loss[2] = cs;
  }
}

int pose_loss_run(const float* pred, const float* gt, long long B, int nj, const int* parents, int n_parents, int drop_first,
                  float lambda_mpjpe, float lambda_cos, float* loss, float* dpose, float* scratch, long long scratch_elems,
                  cudaStream_t st) {
  EB_REQUIRE(pred && gt && parents && loss && dpose && scratch, "pose_loss: null pointer");
  const int n_ext = nj + (drop_first ? 1 : 0);
  EB_REQUIRE(B > 0 && nj >= 2 && n_ext <= 24 && n_parents == n_ext, "pose_loss: %d joints / %d parents not supported", nj,
             n_parents);
  EB_REQUIRE(2 * B <= scratch_elems, "pose_loss: scratch too small");
  LossCfg cfg;
  cfg.nj = nj; cfg.drop_first = drop_first ? 1 : 0; cfg.n_ext = n_ext;
  for (int i = 0; i < 24; ++i) cfg.parents[i] = i < n_parents ? parents[i] : 0;
  for (int i = 0; i < n_parents; ++i) EB_REQUIRE(parents[i] >= 0 && parents[i] < n_ext, "pose_loss: bad parent index");
  cfg.w_mpjpe = lambda_mpjpe / float(B * nj);
  cfg.w_cos = lambda_cos * lambda_mpjpe / float(B);
  {
    ProfScope prof("pose_loss_frame_kernel", st);
    EB_LAUNCH_COOP(pose_loss_frame_kernel, (unsigned)B, 64, st, pred, gt, cfg, scratch, dpose);
    EB_CHECK_LAUNCH("pose_loss_frame_kernel");
  }
  ProfScope prof("pose_loss_reduce_kernel", st);
  EB_LAUNCH_COOP(pose_loss_reduce_kernel, 1, 256, st, scratch, B, cfg.w_mpjpe, cfg.w_cos, loss);
  EB_CHECK_LAUNCH("pose_loss_reduce_kernel");
  return 0;
}

// ---------------------------------------------------------------------------------------------
// AdamW (torch.optim.AdamW semantics; reference model/network.py:72-78), multi-tensor: up to 64 tensors per launch,
// blockIdx.y selects the tensor.
// ---------------------------------------------------------------------------------------------
struct AdamTable {
  float* p[64];
  const float* g[64];
  float* m[64];
  float* v[64];
  long long n[64];
};

__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, float decay, float beta1, float omb1,
                                          float beta2, float omb2, float step, float inv_sqrt_bc2, float eps) {
  m = beta1 * m + omb1 * g;
  v = beta2 * v + omb2 * g * g;
  p = p * decay - step * m / (sqrtf(v) * inv_sqrt_bc2 + eps);
}

__global__ void __launch_bounds__(256) adamw_kernel(AdamTable tab, float decay, float beta1, float omb1, float beta2,
                                                    float omb2, float step, float inv_sqrt_bc2, float eps, float grad_scale) {
  const int k = blockIdx.y;
  float* __restrict__ p = tab.p[k];
  const float* __restrict__ g = tab.g[k];
  float* __restrict__ m = tab.m[k];
  float* __restrict__ v = tab.v[k];
  const long long n = tab.n[k];
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nthreads = (long long)gridDim.x * blockDim.x;
  // 16-byte vector body (every tensor's slot in the flat gradient / state buffers is 256-byte aligned; torch aligns
  // parameter storage), scalar tail
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  const long long n4 = vec ? n / 4 : 0;
  for (long long i = tid; i < n4; i += nthreads) {
    float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    adamw_one(pp.x, gg.x * grad_scale, mm.x, vv.x, decay, beta1, omb1, beta2, omb2, step, inv_sqrt_bc2, eps);
    adamw_one(pp.y, gg.y * grad_scale, mm.y, vv.y, decay, beta1, omb1, beta2, omb2, step, inv_sqrt_bc2, eps);
    adamw_one(pp.z, gg.z * grad_scale, mm.z, vv.z, decay, beta1, omb1, beta2, omb2, step, inv_sqrt_bc2, eps);
    adamw_one(pp.w, gg.w * grad_scale, mm.w, vv.w, decay, beta1, omb1, beta2, omb2, step, inv_sqrt_bc2, eps);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (long long i = n4 * 4 + tid; i < n; i += nthreads) {
    float pi = p[i], mi = m[i], vi = v[i];
    adamw_one(pi, g[i] * grad_scale, mi, vi, decay, beta1, omb1, beta2, omb2, step, inv_sqrt_bc2, eps);
    p[i] = pi; m[i] = mi; v[i] = vi;
  }
}

int adamw_run(float* const* p, const float* const* g, float* const* m, float* const* v, const long long* n, int count, int step,
              double lr, double beta1, double beta2, double eps, double weight_decay, double grad_scale, cudaStream_t st) {
  EB_REQUIRE(p && g && m && v && n && count > 0 && step >= 1, "adamw: bad arguments");
  // hyper-parameters arrive in double, as torch.optim.AdamW holds them; every derived constant is formed in double
  // and rounded to fp32 once (1 - beta2 formed in fp32 would be off by 5e-5 relative)
  const double bc1 = 1.0 - pow(beta1, double(step)), bc2 = 1.0 - pow(beta2, double(step));
  for (int base = 0; base < count; base += 64) {
    AdamTable tab;
    memset(&tab, 0, sizeof(tab));
    const int cnt = count - base < 64 ? count - base : 64;
    long long biggest = 0;
    for (int i = 0; i < cnt; ++i) {
      EB_REQUIRE(p[base + i] && g[base + i] && m[base + i] && v[base + i] && n[base + i] > 0, "adamw: null tensor %d", base + i);
      tab.p[i] = p[base + i]; tab.g[i] = g[base + i]; tab.m[i] = m[base + i]; tab.v[i] = v[base + i]; tab.n[i] = n[base + i];
      if (n[base + i] > biggest) biggest = n[base + i];
    }
    ProfScope prof("adamw_kernel", st);
    EB_LAUNCH(adamw_kernel, dim3(blocks_for(biggest, 256 * 8, 148 * 2), cnt), 256, st, tab, float(1.0 - lr * weight_decay),
              float(beta1), float(1.0 - beta1), float(beta2), float(1.0 - beta2), float(lr / bc1), float(1.0 / sqrt(bc2)),
              float(eps), float(grad_scale));
    EB_CHECK_LAUNCH("adamw_kernel");
  }
  return 0;
}

}  // namespace eb

using namespace eb;

// ---- C ABI (include/egotap_b200.h, "training" section) -------------------------------------------------------
extern "C" int egotap_b200_zero(void* ptr, size_t bytes, void* stream) {
  if (!ptr) return fail(EGOTAP_E_ARG, "zero: null pointer");
  EB_CUDA(cudaMemsetAsync(ptr, 0, bytes, (cudaStream_t)stream));
  return 0;
}
extern "C" int egotap_b200_copy(void* dst, const void* src, size_t bytes, void* stream) {
  if (!dst || !src) return fail(EGOTAP_E_ARG, "copy: null pointer");
  EB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}
extern "C" int egotap_b200_add3(const float* a, const float* b, const float* c, float* out, int n, void* stream) {
  if (!a || !b || !out) return fail(EGOTAP_E_ARG, "add3: null pointer");
  return vec_add3_run(a, b, c, out, n, (cudaStream_t)stream);
}
extern "C" int egotap_b200_split2d(const float* src, long long rows, long long cols, long long src_ld, void* hi, void* lo,
                                   long long dst_ld, void* stream) {
  return split2d_run(src, rows, cols, src_ld, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, dst_ld, (cudaStream_t)stream);
}
extern "C" int egotap_b200_fill_dummy(float* hidden, const float* dummy, int frames, int tokens, int live, void* stream) {
  if (!hidden || !dummy) return fail(EGOTAP_E_ARG, "fill_dummy: null pointer");
  return fill_dummy_run(hidden, dummy, frames, tokens, live, (cudaStream_t)stream);
}
extern "C" int egotap_b200_pos_permute(const float* pos, const float* mask_token, int grid, int n_hm, float* pos_perm,
                                       float* dummy, void* stream) {
  if (!pos || !mask_token || !pos_perm || !dummy) return fail(EGOTAP_E_ARG, "pos_permute: null pointer");
  if (grid * grid * 16 != 576 || n_hm <= 0 || n_hm > grid * grid) return fail(EGOTAP_E_ARG, "pos_permute: bad geometry");
  return pos_permute_run(pos, mask_token, grid, n_hm, pos_perm, dummy, (cudaStream_t)stream);
}
extern "C" int egotap_b200_pu_bridge_gate(const float* f, int f_ld, int f_col, const float* e, int e_ld, int X, long long rows,
                                          void* hi, void* lo, void* stream) {
  return pu_bridge_gate_run(f, f_ld, f_col, e, e_ld, X, rows, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, (cudaStream_t)stream);
}
extern "C" int egotap_b200_pu_cell_fwd(const float* G, long long g_rs, long long g_ts, const float* F, long long f_rs,
                                       long long f_ts, float* C, float* H, void* h_hi, void* h_lo, void* hg_hi, void* hg_lo,
                                       int t, int J, long long frames, void* stream) {
  return pu_cell_fwd_run(G, g_rs, g_ts, F, f_rs, f_ts, C, H, (__nv_bfloat16*)h_hi, (__nv_bfloat16*)h_lo, (__nv_bfloat16*)hg_hi,
                         (__nv_bfloat16*)hg_lo, t, J, frames, (cudaStream_t)stream);
}
extern "C" int egotap_b200_pu_cell_bwd(const float* G, long long g_rs, long long g_ts, const float* F, long long f_rs,
                                       long long f_ts, const float* C, const float* H, const float* dOut, const float* dhg,
                                       float* dc, float* dG, long long dg_rs, long long dg_ts, float* dF, long long df_rs,
                                       long long df_ts, void* dgp_hi, void* dgp_lo, int t, int J, long long frames,
                                       void* stream) {
  return pu_cell_bwd_run(G, g_rs, g_ts, F, f_rs, f_ts, C, H, dOut, dhg, dc, dG, dg_rs, dg_ts, dF, df_rs, df_ts,
                         (__nv_bfloat16*)dgp_hi, (__nv_bfloat16*)dgp_lo, t, J, frames, (cudaStream_t)stream);
}
extern "C" int egotap_b200_pu_bridge_gate_bwd(float* dE, long long e_ld, const float* F0, long long f_ld, int f_col,
                                              const float* E, int X, long long rows, float* dF, long long df_ld, void* stream) {
  return pu_bridge_gate_bwd_run(dE, e_ld, F0, f_ld, f_col, E, X, rows, dF, df_ld, (cudaStream_t)stream);
}
extern "C" int egotap_b200_regroup_gather(const float* dE, long long e_ld, int col_off, long long frames, int J, int cols,
                                          float* out, void* stream) {
  return regroup_gather_run(dE, e_ld, col_off, frames, J, cols, out, (cudaStream_t)stream);
}
extern "C" int egotap_b200_head_bwd(const float* dpose, const float* e, long long e_ld, const float* skel, const float* Wp,
                                    const float* Wg, long long frames, int J, float* dE, long long de_ld, float* dSkel,
                                    float* dWp, float* dbp, float* dWg, float* dbg, float* scratch, long long scratch_elems,
                                    void* stream) {
  return head_bwd_run(dpose, e, e_ld, skel, Wp, Wg, frames, J, dE, de_ld, dSkel, dWp, dbp, dWg, dbg, scratch, scratch_elems,
                      (cudaStream_t)stream);
}
extern "C" int egotap_b200_embed_grads(const float* dpos_perm, int grid, int n_hm, float* dpos, float* dmask, void* stream) {
  return embed_grads_run(dpos_perm, grid, n_hm, dpos, dmask, (cudaStream_t)stream);
}
extern "C" int egotap_b200_pose_loss(const float* pred, const float* gt, long long frames, int joints, const int* parents_host,
                                     int n_parents, int drop_first, float lambda_mpjpe, float lambda_cos, float* loss,
                                     float* dpose, float* scratch, long long scratch_elems, void* stream) {
  return pose_loss_run(pred, gt, frames, joints, parents_host, n_parents, drop_first, lambda_mpjpe, lambda_cos, loss, dpose,
                       scratch, scratch_elems, (cudaStream_t)stream);
}
extern "C" int egotap_b200_adamw(float* const* params_host, const float* const* grads_host, float* const* m_host,
                                 float* const* v_host, const long long* numel_host, int count, int step, double lr, double beta1,
                                 double beta2, double eps, double weight_decay, double grad_scale, void* stream) {
  return adamw_run(params_host, grads_host, m_host, v_host, numel_host, count, step, lr, beta1, beta2, eps, weight_decay,
                   grad_scale, (cudaStream_t)stream);
}

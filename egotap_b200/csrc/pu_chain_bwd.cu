// Persistent BPTT kernel of the propagation chain (training step): one launch walks the J joints of
// one propagation-unit layer BACKWARDS (reference: torch.autograd through model/custom_cells.py:94-120,149-197).  It is the
// mirror image of pu_chain.cu and replaces the per-joint sequence (pu_cell_bwd kernel + dgates . W_hh GEMM) of
// train_model.cu -- same arithmetic, no per-joint launch.  Per joint t (descending):
//     dh      = dOut[b, t] + dhg_{t+1} * sigmoid(F[b, t+1])            dF[b, t+1] = dhg_{t+1} * h_t * sigmoid'(F[b, t+1])
//     dc_tot  = dc + dh * o * (1 - tanh(c_t)^2)                         dc <- dc_tot * f
//     dgates  = [dc_tot c_{t-1} f', dc_tot tanh(g) i', dc_tot i (1 - tanh(g)^2), dh tanh(c_t) o']   -> dG[b, t] (fp32)
//     dhg_t   = dgates . W_hh                                           (B x 2048 . 2048 x 512, the recurrent gradient)
// Work split: 32 CTAs per batch group of up to 256 frames; CTA `sl` owns hidden units [16 sl, 16 sl + 16): it computes the
// gate gradients of those units (64 of the 2048 gate columns) for every frame, keeps the 16 x 2048 slice of W_hh^T (bf16
// hi/lo) RESIDENT IN SHARED MEMORY, and after a group barrier contracts the full dgates rows (streamed by TMA from the
// exchange buffer every CTA wrote its columns to) with that slice: 128 x 16 x 16 tcgen05 MMAs into TMEM.  The carried
// cell-state gradient stays in registers across joints.
#include "gemm.cuh"
#include "host_util.cuh"
#include "internal.h"

namespace eb {

constexpr int PB_H = 512, PB_U = 16, PB_SLICES = PB_H / PB_U, PB_G = 4 * PB_H, PB_KB = PB_G / 64, PB_RPG = 256;

template <int NSPLIT>
struct PbCfg {
  static constexpr int NOPS = NSPLIT == 1 ? 1 : 2;
  static constexpr int W_BLK = PB_U * 64 * 2;                 // one 64-wide K block of the W_hh^T slice (2 KB = 2 swizzle atoms)
  static constexpr int W_BYTES = NOPS * PB_KB * W_BLK;        // 64 / 128 KB resident
  static constexpr int A_BLK = 128 * 64 * 2;                  // 16 KB
  static constexpr int STAGE_BYTES = NOPS * A_BLK;
  static constexpr int STAGES = NSPLIT == 1 ? 4 : 2;
  static constexpr int BAR_OFF = W_BYTES + STAGES * STAGE_BYTES;
  static constexpr int SMEM_BYTES = BAR_OFF + 256;
  static constexpr int TMEM_COLS = 64;                        // 2 row tiles at columns 0 / 32, 16 accumulator columns each
  static constexpr int T_STRIDE = 32;                         // (the epilogue reads them with one 32-column tcgen05.ld)
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

struct PbParams {
  const float* G; long long G_rs, G_ts;     // complete gate pre-activations: G[b*G_rs + t*G_ts + gate*512 + u]
  const float* F; long long F_rs, F_ts;     // pre-sigmoid forget gate on h
  const float* C;                           // cell states   C[(b*J + t)*512 + u]
  const float* H;                           // hidden states H[(b*J + t)*512 + u]
  const float* dOut;                        // gradient wrt the layer's outputs, same layout as H
  float* dG; long long dG_rs, dG_ts;        // gate gradients (fp32), layout as G
  float* dF; long long dF_rs, dF_ts;        // gradient wrt F (columns [0, 512) of the row)
  __nv_bfloat16* x_hi; __nv_bfloat16* x_lo; // exchange buffer [2][B][2048]: bf16 (hi/lo) dgates of the previous step
  unsigned int* counters;                   // one per batch group, zeroed before launch
  int B, J;
};

__device__ __forceinline__ float pb_sigm(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int NSPLIT>
__global__ void __launch_bounds__(192, 1)
pu_chain_bwd_kernel(const __grid_constant__ CUtensorMap tmWh, const __grid_constant__ CUtensorMap tmWl,
                    const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl, const PbParams p) {
  using C = PbCfg<NSPLIT>;
  EB_DYN_SMEM_1K(smem);
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sW = smem;
  uint8_t* sA = smem + C::W_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
  uint64_t* w_full = bars;
  uint64_t* a_full = bars + 1;
  uint64_t* a_empty = a_full + C::STAGES;
  uint64_t* acc_full = a_empty + C::STAGES;
  uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bg = blockIdx.x / PB_SLICES, sl = blockIdx.x % PB_SLICES;
  const int r0 = bg * PB_RPG;
  const int rows = min(PB_RPG, p.B - r0);
  const int nmt = (rows + 127) / 128;
  const int u0 = sl * PB_U;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmWh); tma_prefetch_desc(&tmAh);
    if (NSPLIT > 1) { tma_prefetch_desc(&tmWl); tma_prefetch_desc(&tmAl); }
  }
  if (warp == 1 && lane == 0) {
    mbar_init(w_full, 1);
    for (int i = 0; i < C::STAGES; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    mbar_init(acc_full, 1); mbar_init(acc_empty, 4);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<1>(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // resident weight slice: rows [16 sl, 16 sl + 16) of W_hh^T (512 x 2048), all 32 K blocks
  if (warp == 0 && lane == 0) {
    mbar_expect_tx(w_full, C::W_BYTES);
    for (int kb = 0; kb < PB_KB; ++kb) {
      tma_load_4d(sW + kb * C::W_BLK, &tmWh, w_full, kb * 64, sl * PB_U, 0, 0);
      if (NSPLIT > 1) tma_load_4d(sW + (PB_KB + kb) * C::W_BLK, &tmWl, w_full, kb * 64, sl * PB_U, 0, 0);
    }
  }

  uint32_t ring = 0;          // producer / MMA ring counter (same sequence in both roles)
  float dc[2][PB_U];          // carried cell-state gradient of this thread's rows, persistent across joints
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int u = 0; u < PB_U; ++u) dc[m][u] = 0.f;

  for (int s = 0; s < p.J; ++s) {            // step s handles joint t = J-1-s
    const int t = p.J - 1 - s;
    const int rd = s & 1, wr = rd ^ 1;       // exchange buffer read this step / written for the next step
    if (warp == 0) {
      // -------------------------------------------------------------- TMA producer: full dgates rows of joint t+1
      if (lane == 0 && s > 0) {
        for (int mt = 0; mt < nmt; ++mt)
          for (int kb = 0; kb < PB_KB; ++kb) {
            const uint32_t st = ring % C::STAGES, ph = (ring / C::STAGES) & 1;
            mbar_wait(&a_empty[st], ph ^ 1);
            mbar_expect_tx(&a_full[st], C::STAGE_BYTES);
            uint8_t* dst = sA + st * C::STAGE_BYTES;
            tma_load_4d(dst, &tmAh, &a_full[st], kb * 64, rd * p.B + r0 + mt * 128, 0, 0);
            if (NSPLIT > 1) tma_load_4d(dst + C::A_BLK, &tmAl, &a_full[st], kb * 64, rd * p.B + r0 + mt * 128, 0, 0);
            ++ring;
          }
      }
    } else if (warp == 1) {
      // -------------------------------------------------------------- MMA issuer: dhg_{t+1} = dgates_{t+1} . W_hh[:, own units]
      if (s > 0) {
        constexpr uint32_t idesc = make_idesc_bf16(128, PB_U);
        if (s == 1) mbar_wait(w_full, 0);
        mbar_wait(acc_empty, (s - 1) & 1);   // epilogue of step s-1 has drained the accumulators
        tc_fence_after();
        const uint32_t w_lo = sdesc_lo(smem_u32(sW)), a_base = smem_u32(sA);
        for (int mt = 0; mt < nmt; ++mt)
          for (int kb = 0; kb < PB_KB; ++kb) {
            const uint32_t st = ring % C::STAGES, ph = (ring / C::STAGES) & 1;
            mbar_wait(&a_full[st], ph);
            tc_fence_after();
            const uint32_t a_lo = sdesc_lo(a_base + st * C::STAGE_BYTES);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t ah = sdesc_at(a_lo, kk * 32);
              const uint64_t wh = sdesc_at(w_lo, kb * C::W_BLK + kk * 32);
              umma_bf16<1>(tmem_base + mt * C::T_STRIDE, ah, wh, idesc, (kb | kk) != 0 ? 1u : 0u);
              if (NSPLIT > 1) {
                const uint64_t al = sdesc_at(a_lo, C::A_BLK + kk * 32);
                const uint64_t wl = sdesc_at(w_lo, (PB_KB + kb) * C::W_BLK + kk * 32);
                umma_bf16<1>(tmem_base + mt * C::T_STRIDE, ah, wl, idesc, 1u);
                umma_bf16<1>(tmem_base + mt * C::T_STRIDE, al, wh, idesc, 1u);
              }
            }
            umma_commit<1>(&a_empty[st]);
            ++ring;
          }
        umma_commit<1>(acc_full);
      }
    } else {
      // -------------------------------------------------------------- gate gradients (4 warps = 128 rows per tile)
      const int q = warp & 3;
      const uint32_t lane_sel = uint32_t(q * 32) << 16;
      if (s > 0) {
        mbar_wait(acc_full, (s - 1) & 1);
        tc_fence_after();
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {          // unrolled so the register-resident carried gradient is statically indexed
        if (mt >= nmt) continue;
        const int b = r0 + mt * 128 + q * 32 + lane;
        const bool valid = (mt * 128 + q * 32 + lane) < rows;
        float dhg[PB_U];
        if (s > 0) {
          uint32_t acc[32];                   // the tile's 16 accumulator columns + 16 spare columns of its 32-column slot
          tmem_ld32(tmem_base + lane_sel + mt * C::T_STRIDE, acc);
          tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < PB_U; ++u) dhg[u] = __uint_as_float(acc[u]);
        } else {
#pragma unroll
          for (int u = 0; u < PB_U; ++u) dhg[u] = 0.f;
        }
        if (valid) {
          const float* g = p.G + (long long)b * p.G_rs + (long long)t * p.G_ts + u0;
          const long long row = ((long long)b * p.J + t) * PB_H + u0;
          float dgt[4][PB_U];
#pragma unroll
          for (int v4 = 0; v4 < PB_U / 4; ++v4) {
            const float4 gf = *reinterpret_cast<const float4*>(g + v4 * 4);
            const float4 gi = *reinterpret_cast<const float4*>(g + PB_H + v4 * 4);
            const float4 gg = *reinterpret_cast<const float4*>(g + 2 * PB_H + v4 * 4);
            const float4 go = *reinterpret_cast<const float4*>(g + 3 * PB_H + v4 * 4);
            const float4 ct4 = *reinterpret_cast<const float4*>(p.C + row + v4 * 4);
            float4 cp4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t > 0) cp4 = *reinterpret_cast<const float4*>(p.C + row - PB_H + v4 * 4);
            const float4 ht4 = *reinterpret_cast<const float4*>(p.H + row + v4 * 4);
            const float4 do4 = *reinterpret_cast<const float4*>(p.dOut + row + v4 * 4);
            float4 fn4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (s > 0) fn4 = *reinterpret_cast<const float4*>(p.F + (long long)b * p.F_rs + (long long)(t + 1) * p.F_ts + u0 + v4 * 4);
            const float xf[4] = {gf.x, gf.y, gf.z, gf.w}, xi[4] = {gi.x, gi.y, gi.z, gi.w};
            const float xg[4] = {gg.x, gg.y, gg.z, gg.w}, xo[4] = {go.x, go.y, go.z, go.w};
            const float ct[4] = {ct4.x, ct4.y, ct4.z, ct4.w}, cp[4] = {cp4.x, cp4.y, cp4.z, cp4.w};
            const float ht[4] = {ht4.x, ht4.y, ht4.z, ht4.w}, dout[4] = {do4.x, do4.y, do4.z, do4.w};
            const float fn[4] = {fn4.x, fn4.y, fn4.z, fn4.w};
            float dfn[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int u = v4 * 4 + e;
              float dh = dout[e];
              dfn[e] = 0.f;
              if (s > 0) {
                const float sf = pb_sigm(fn[e]);
                dh += dhg[u] * sf;
                dfn[e] = dhg[u] * ht[e] * sf * (1.f - sf);
              }
              const float sfg = pb_sigm(xf[e]), sig = pb_sigm(xi[e]), sog = pb_sigm(xo[e]);
              const float tcg = tanhf(xg[e]), tc = tanhf(ct[e]);
              const float dct = (s > 0 ? dc[mt][u] : 0.f) + dh * sog * (1.f - tc * tc);
              dgt[0][u] = dct * cp[e] * sfg * (1.f - sfg);
              dgt[1][u] = dct * tcg * sig * (1.f - sig);
              dgt[2][u] = dct * sig * (1.f - tcg * tcg);
              dgt[3][u] = dh * tc * sog * (1.f - sog);
              dc[mt][u] = dct * sfg;
            }
            if (s > 0)
              *reinterpret_cast<float4*>(p.dF + (long long)b * p.dF_rs + (long long)(t + 1) * p.dF_ts + u0 + v4 * 4) =
                  make_float4(dfn[0], dfn[1], dfn[2], dfn[3]);
            if (t == 0)
              *reinterpret_cast<float4*>(p.dF + (long long)b * p.dF_rs + u0 + v4 * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          float* og = p.dG + (long long)b * p.dG_rs + (long long)t * p.dG_ts + u0;
#pragma unroll
          for (int gt = 0; gt < 4; ++gt)
#pragma unroll
            for (int v4 = 0; v4 < PB_U / 4; ++v4)
              *reinterpret_cast<float4*>(og + gt * PB_H + v4 * 4) =
                  make_float4(dgt[gt][4 * v4], dgt[gt][4 * v4 + 1], dgt[gt][4 * v4 + 2], dgt[gt][4 * v4 + 3]);
          if (t > 0) {   // this CTA's 64 columns of the next step's MMA operand
            const long long xrow = ((long long)wr * p.B + b) * PB_G + u0;
#pragma unroll
            for (int gt = 0; gt < 4; ++gt) {
              uint32_t hh[8], ll[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) split_pack2(dgt[gt][2 * e], dgt[gt][2 * e + 1], hh[e], ll[e]);
              *reinterpret_cast<uint4*>(p.x_hi + xrow + gt * PB_H) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
              *reinterpret_cast<uint4*>(p.x_hi + xrow + gt * PB_H + 8) = make_uint4(hh[4], hh[5], hh[6], hh[7]);
              if (NSPLIT > 1) {
                *reinterpret_cast<uint4*>(p.x_lo + xrow + gt * PB_H) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
                *reinterpret_cast<uint4*>(p.x_lo + xrow + gt * PB_H + 8) = make_uint4(ll[4], ll[5], ll[6], ll[7]);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
    }
    // ---------------------------------------------------------------- group barrier between joints
    if (s + 1 < p.J) {
      fence_proxy_async_all();   // this thread's st.global -> later TMA (async proxy) reads
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) {
        fence_proxy_async_all();
        atomicAdd(p.counters + bg, 1u);
        const unsigned int want = PB_SLICES * (s + 1);
        const long long t0 = clock64();
        unsigned int seen;
        do {
          seen = ld_acquire_gpu_u32(p.counters + bg);
          if (clock64() - t0 > EB_WAIT_TIMEOUT_CYCLES) { printf("egotap_b200: pu_chain_bwd group barrier timeout\n"); __trap(); }
        } while (seen < want);
        fence_proxy_async_all();
      }
      __syncthreads();
    }
  }

  if (warp == 0 && lane == 0 && p.J < 2) mbar_wait(w_full, 0);   // never leave with the weight TMA in flight
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, C::TMEM_COLS);
}

template <int NSPLIT>
static int launch_pb(const CUtensorMap* tm, const PbParams& p, int groups, cudaStream_t stream) {
  using C = PbCfg<NSPLIT>;
  auto kern = pu_chain_bwd_kernel<NSPLIT>;
  static bool attr_done[64] = {false};
  const int dev_ = current_device();
  if (!attr_done[dev_]) {
    EB_CUDA(EB_SET_MAX_SMEM(kern, C::SMEM_BYTES));
    attr_done[dev_] = true;
  }
  ProfScope prof("pu_chain_bwd_kernel", stream);
  EB_LAUNCH_GRID_SYNC(kern, groups * PB_SLICES, 192, C::SMEM_BYTES, stream, tm[0], tm[1], tm[2], tm[3], p);
  EB_CHECK_LAUNCH("pu_chain_bwd_kernel");
  return 0;
}

// BPTT of one propagation-unit layer over all J joints.  wT: W_hh^T (512 x 2048) bf16 hi/lo; x: exchange scratch
// [2][B][2048] bf16 hi/lo; counters: >= ceil(B/256) words; B <= 1024 per call (all CTAs must be co-resident).
int pu_chain_bwd_run(const __nv_bfloat16* wT_hi, const __nv_bfloat16* wT_lo, const float* G, long long G_rs, long long G_ts,
                     const float* F, long long F_rs, long long F_ts, const float* Cs, const float* H, const float* dOut,
                     float* dG, long long dG_rs, long long dG_ts, float* dF, long long dF_rs, long long dF_ts,
                     __nv_bfloat16* x_hi, __nv_bfloat16* x_lo, unsigned int* counters, int B, int J, int nsplit,
                     cudaStream_t stream) {
  EB_REQUIRE(wT_hi && G && F && Cs && H && dOut && dG && dF && x_hi && counters && B > 0 && J > 0, "pu_chain_bwd: bad arguments");
  EB_REQUIRE(nsplit == 1 || (wT_lo && x_lo), "pu_chain_bwd: bf16x3 mode needs the lo parts");
  EB_REQUIRE(G_rs % 4 == 0 && G_ts % 4 == 0 && F_rs % 4 == 0 && F_ts % 4 == 0 && dG_rs % 4 == 0 && dG_ts % 4 == 0 &&
                 dF_rs % 4 == 0 && dF_ts % 4 == 0, "pu_chain_bwd: strides must be multiples of 4");
  const int groups = (B + PB_RPG - 1) / PB_RPG;
  EB_REQUIRE(groups * PB_SLICES <= num_sms(), "pu_chain_bwd: batch %d needs %d co-resident CTAs (> %d SMs); chunk the batch", B,
             groups * PB_SLICES, num_sms());
  CUtensorMap tm[4];
  int rc;
  if ((rc = make_operand_tmap(&tm[0], wT_hi, PB_G, PB_H, PB_G, 1, 0, 1, 0, PB_U))) return rc;
  if ((rc = make_operand_tmap(&tm[2], x_hi, PB_G, 2ll * B, PB_G, 1, 0, 1, 0, 128))) return rc;
  if (nsplit == 3) {
    if ((rc = make_operand_tmap(&tm[1], wT_lo, PB_G, PB_H, PB_G, 1, 0, 1, 0, PB_U))) return rc;
    if ((rc = make_operand_tmap(&tm[3], x_lo, PB_G, 2ll * B, PB_G, 1, 0, 1, 0, 128))) return rc;
  } else {
    tm[1] = tm[0];
    tm[3] = tm[2];
  }
  EB_CUDA(cudaMemsetAsync(counters, 0, sizeof(unsigned int) * groups, stream));
  PbParams p{G, G_rs, G_ts, F, F_rs, F_ts, Cs, H, dOut, dG, dG_rs, dG_ts, dF, dF_rs, dF_ts, x_hi, x_lo, counters, B, J};
  return nsplit == 3 ? launch_pb<3>(tm, p, groups, stream) : launch_pb<1>(tm, p, groups, stream);
}

}  // namespace eb

extern "C" int egotap_b200_pu_chain_bwd(const void* wT_hi, const void* wT_lo, const float* G, long long g_rs, long long g_ts,
                                        const float* F, long long f_rs, long long f_ts, const float* C, const float* H,
                                        const float* dOut, float* dG, long long dg_rs, long long dg_ts, float* dF,
                                        long long df_rs, long long df_ts, void* x_hi, void* x_lo, void* counters, int frames,
                                        int J, int precision, void* stream) {
  return eb::pu_chain_bwd_run((const __nv_bfloat16*)wT_hi, (const __nv_bfloat16*)wT_lo, G, g_rs, g_ts, F, f_rs, f_ts, C, H, dOut,
                              dG, dg_rs, dg_ts, dF, df_rs, df_ts, (__nv_bfloat16*)x_hi, (__nv_bfloat16*)x_lo,
                              (unsigned int*)counters, frames, J, precision == EGOTAP_PREC_BF16 ? 1 : 3, (cudaStream_t)stream);
}

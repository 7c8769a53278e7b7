// Persistent propagation-chain kernel: one launch walks all J joints of one propagation-unit layer
// (reference model/custom_cells.py:94-120,149-197 driven by model/net_architecture.py:539-568; chain semantics,
// SURVEY 0.4).  Per joint t the recurrence is
//     gates = G[b, t] + (sigmoid(F[b, t]) * h_{t-1}) . W_hh^T          (G, F: batched x-side projections, fp32)
//     c = c * sig(f) + sig(i) * tanh(g) ;  h = sig(o) * tanh(c)         (gate order f, i, g, o)
// Work split: 32 CTAs per batch group of up to 256 frames; CTA `sl` owns hidden units [16 sl, 16 sl + 16), i.e. 64
// gate columns, and keeps that 64 x 512 slice of W_hh (bf16 hi/lo) RESIDENT IN SHARED MEMORY for all joints.
// Each joint: TMA streams the pre-gated state (B x 512 bf16 hi/lo, written by all CTAs in the previous step) through
// a ring, one thread issues tcgen05 MMAs (128 x 64 x 16) into TMEM, four epilogue warps apply the gate math with
// the cell state held in registers, publish h (fp32 + operand copies) and the next pre-gated state, and the CTAs of
// the group meet at a global-memory barrier.  No per-joint kernel launch.
#include "gemm.cuh"
#include "host_util.cuh"
#include "internal.h"

namespace eb {

constexpr int PC_H = 512, PC_U = 16, PC_SLICES = PC_H / PC_U, PC_N = 4 * PC_U, PC_KB = PC_H / 64, PC_RPG = 256;

template <int NSPLIT>
struct PuCfg {
  static constexpr int NOPS = NSPLIT == 1 ? 1 : 2;
  static constexpr int W_BLK = PC_N * 64 * 2;                 // one 64-wide K block of the slice (8 KB)
  static constexpr int W_BYTES = NOPS * PC_KB * W_BLK;        // 64 / 128 KB resident
  static constexpr int A_BLK = 128 * 64 * 2;                  // 16 KB
  static constexpr int STAGE_BYTES = NOPS * A_BLK;
  static constexpr int STAGES = NSPLIT == 1 ? 4 : 2;
  static constexpr int BAR_OFF = W_BYTES + STAGES * STAGE_BYTES;
  static constexpr int SMEM_BYTES = BAR_OFF + 256;
  static constexpr int TMEM_COLS = 128;                       // 2 row tiles x 64 gate columns
};

__device__ __forceinline__ float fast_sigmoid(float x) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float fast_tanh(float x) { return fmaf(2.0f, fast_sigmoid(2.0f * x), -1.0f); }

struct PuParams {
  const float* G; long long G_rs, G_ts;     // gates x-side term: G[b*G_rs + t*G_ts + gate*512 + u]
  const float* F; long long F_rs, F_ts;     // pre-sigmoid forget gate on h: F[b*F_rs + t*F_ts + u]
  float* out;                               // h: out[(b*J + t)*512 + u]
  __nv_bfloat16* out_hi; __nv_bfloat16* out_lo;   // optional operand copies of h (next layer's GEMM input)
  __nv_bfloat16* hg_hi; __nv_bfloat16* hg_lo;     // ping-pong pre-gated state, [2][B][512]
  unsigned int* counters;                   // one per batch group, zeroed before launch
  int B, J;
};

template <int NSPLIT>
__global__ void __launch_bounds__(192, 1)
pu_chain_kernel(const __grid_constant__ CUtensorMap tmWh, const __grid_constant__ CUtensorMap tmWl,
                const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl, const PuParams p) {
  using C = PuCfg<NSPLIT>;
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
  const int bg = blockIdx.x / PC_SLICES, sl = blockIdx.x % PC_SLICES;
  const int r0 = bg * PC_RPG;
  const int rows = min(PC_RPG, p.B - r0);
  const int nmt = (rows + 127) / 128;
  const int u0 = sl * PC_U;

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

  // resident weight slice: rows [64 sl, 64 sl + 64) of the gate-permuted W_hh (pack time), all 8 K blocks
  if (warp == 0 && lane == 0) {
    mbar_expect_tx(w_full, C::W_BYTES);
    for (int kb = 0; kb < PC_KB; ++kb) {
      tma_load_4d(sW + kb * C::W_BLK, &tmWh, w_full, kb * 64, sl * PC_N, 0, 0);
      if (NSPLIT > 1) tma_load_4d(sW + (PC_KB + kb) * C::W_BLK, &tmWl, w_full, kb * 64, sl * PC_N, 0, 0);
    }
  }

  uint32_t ring = 0;          // producer / MMA ring counter (same sequence in both roles)
  float c[2][PC_U];           // cell state of this thread's rows (epilogue warps), persistent across joints
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int u = 0; u < PC_U; ++u) c[m][u] = 0.f;

  for (int t = 0; t < p.J; ++t) {
    const int rd = t & 1, wr = rd ^ 1;       // hg buffer read this step / written for the next step
    if (warp == 0) {
      // -------------------------------------------------------------- TMA producer: pre-gated state tiles
      if (lane == 0 && t > 0) {
        for (int mt = 0; mt < nmt; ++mt)
          for (int kb = 0; kb < PC_KB; ++kb) {
            const uint32_t s = ring % C::STAGES, ph = (ring / C::STAGES) & 1;
            mbar_wait(&a_empty[s], ph ^ 1);
            mbar_expect_tx(&a_full[s], C::STAGE_BYTES);
            uint8_t* dst = sA + s * C::STAGE_BYTES;
            tma_load_4d(dst, &tmAh, &a_full[s], kb * 64, rd * p.B + r0 + mt * 128, 0, 0);
            if (NSPLIT > 1) tma_load_4d(dst + C::A_BLK, &tmAl, &a_full[s], kb * 64, rd * p.B + r0 + mt * 128, 0, 0);
            ++ring;
          }
      }
    } else if (warp == 1) {
      // -------------------------------------------------------------- MMA issuer
      if (t > 0) {
        constexpr uint32_t idesc = make_idesc_bf16(128, PC_N);
        if (t == 1) mbar_wait(w_full, 0);
        mbar_wait(acc_empty, (t - 1) & 1);   // epilogue of step t-1 has drained the accumulators
        tc_fence_after();
        const uint32_t w_lo = sdesc_lo(smem_u32(sW)), a_base = smem_u32(sA);
        for (int mt = 0; mt < nmt; ++mt)
          for (int kb = 0; kb < PC_KB; ++kb) {
            const uint32_t s = ring % C::STAGES, ph = (ring / C::STAGES) & 1;
            mbar_wait(&a_full[s], ph);
            tc_fence_after();
            const uint32_t a_lo = sdesc_lo(a_base + s * C::STAGE_BYTES);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t ah = sdesc_at(a_lo, kk * 32);
              const uint64_t wh = sdesc_at(w_lo, kb * C::W_BLK + kk * 32);
              umma_bf16<1>(tmem_base + mt * PC_N, ah, wh, idesc, (kb | kk) != 0 ? 1u : 0u);
              if (NSPLIT > 1) {
                const uint64_t al = sdesc_at(a_lo, C::A_BLK + kk * 32);
                const uint64_t wl = sdesc_at(w_lo, (PC_KB + kb) * C::W_BLK + kk * 32);
                umma_bf16<1>(tmem_base + mt * PC_N, ah, wl, idesc, 1u);
                umma_bf16<1>(tmem_base + mt * PC_N, al, wh, idesc, 1u);
              }
            }
            umma_commit<1>(&a_empty[s]);
            ++ring;
          }
        umma_commit<1>(acc_full);
      }
    } else {
      // -------------------------------------------------------------- gate math (4 warps = 128 rows per tile)
      const int q = warp & 3;
      const uint32_t lane_sel = uint32_t(q * 32) << 16;
      if (t > 0) {
        mbar_wait(acc_full, (t - 1) & 1);
        tc_fence_after();
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {          // unrolled so the register-resident cell state is statically indexed
        if (mt >= nmt) continue;
        const int b = r0 + mt * 128 + q * 32 + lane;
        const bool valid = (mt * 128 + q * 32 + lane) < rows;
        uint32_t acc[64];
        if (t > 0) {
          tmem_ld32(tmem_base + lane_sel + mt * PC_N, reinterpret_cast<uint32_t(&)[32]>(acc[0]));
          tmem_ld32(tmem_base + lane_sel + mt * PC_N + 32, reinterpret_cast<uint32_t(&)[32]>(acc[32]));
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 64; ++i) acc[i] = 0u;
        }
        if (valid) {
          const float* g = p.G + (long long)b * p.G_rs + (long long)t * p.G_ts + u0;
          float h[PC_U];
#pragma unroll
          for (int v4 = 0; v4 < PC_U / 4; ++v4) {
            const float4 gf = *reinterpret_cast<const float4*>(g + v4 * 4);
            const float4 gi = *reinterpret_cast<const float4*>(g + PC_H + v4 * 4);
            const float4 gg = *reinterpret_cast<const float4*>(g + 2 * PC_H + v4 * 4);
            const float4 go = *reinterpret_cast<const float4*>(g + 3 * PC_H + v4 * 4);
            const float xf[4] = {gf.x, gf.y, gf.z, gf.w}, xi[4] = {gi.x, gi.y, gi.z, gi.w};
            const float xg[4] = {gg.x, gg.y, gg.z, gg.w}, xo[4] = {go.x, go.y, go.z, go.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int u = v4 * 4 + e;
              const float fg = __uint_as_float(acc[u]) + xf[e];
              const float ig = __uint_as_float(acc[PC_U + u]) + xi[e];
              const float cg = __uint_as_float(acc[2 * PC_U + u]) + xg[e];
              const float og = __uint_as_float(acc[3 * PC_U + u]) + xo[e];
              const float cn = c[mt][u] * fast_sigmoid(fg) + fast_sigmoid(ig) * fast_tanh(cg);
              c[mt][u] = cn;
              h[u] = fast_sigmoid(og) * fast_tanh(cn);
            }
          }
          const long long orow = ((long long)b * p.J + t) * PC_H + u0;
#pragma unroll
          for (int v4 = 0; v4 < PC_U / 4; ++v4)
            *reinterpret_cast<float4*>(p.out + orow + v4 * 4) = make_float4(h[4 * v4], h[4 * v4 + 1], h[4 * v4 + 2], h[4 * v4 + 3]);
          if (p.out_hi) {
            uint32_t hh[8], ll[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) split_pack2(h[2 * e], h[2 * e + 1], hh[e], ll[e]);
            *reinterpret_cast<uint4*>(p.out_hi + orow) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
            *reinterpret_cast<uint4*>(p.out_hi + orow + 8) = make_uint4(hh[4], hh[5], hh[6], hh[7]);
            if (p.out_lo) {
              *reinterpret_cast<uint4*>(p.out_lo + orow) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
              *reinterpret_cast<uint4*>(p.out_lo + orow + 8) = make_uint4(ll[4], ll[5], ll[6], ll[7]);
            }
          }
          if (t + 1 < p.J) {
            const float* f = p.F + (long long)b * p.F_rs + (long long)(t + 1) * p.F_ts + u0;
            uint32_t hh[8], ll[8];
#pragma unroll
            for (int v4 = 0; v4 < PC_U / 4; ++v4) {
              const float4 ff = *reinterpret_cast<const float4*>(f + v4 * 4);
              const float a0 = fast_sigmoid(ff.x) * h[4 * v4], a1 = fast_sigmoid(ff.y) * h[4 * v4 + 1];
              const float a2 = fast_sigmoid(ff.z) * h[4 * v4 + 2], a3 = fast_sigmoid(ff.w) * h[4 * v4 + 3];
              split_pack2(a0, a1, hh[2 * v4], ll[2 * v4]);
              split_pack2(a2, a3, hh[2 * v4 + 1], ll[2 * v4 + 1]);
            }
            const long long hrow = ((long long)wr * p.B + b) * PC_H + u0;
            *reinterpret_cast<uint4*>(p.hg_hi + hrow) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
            *reinterpret_cast<uint4*>(p.hg_hi + hrow + 8) = make_uint4(hh[4], hh[5], hh[6], hh[7]);
            if (NSPLIT > 1) {
              *reinterpret_cast<uint4*>(p.hg_lo + hrow) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
              *reinterpret_cast<uint4*>(p.hg_lo + hrow + 8) = make_uint4(ll[4], ll[5], ll[6], ll[7]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
    }
    // ---------------------------------------------------------------- group barrier between joints
    if (t + 1 < p.J) {
      fence_proxy_async_all();   // this thread's st.global -> later TMA (async proxy) reads
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) {
        fence_proxy_async_all();
        atomicAdd(p.counters + bg, 1u);
        const unsigned int want = PC_SLICES * (t + 1);
        const long long t0 = clock64();
        unsigned int seen;
        do {
          seen = ld_acquire_gpu_u32(p.counters + bg);
          if (clock64() - t0 > EB_WAIT_TIMEOUT_CYCLES) { printf("egotap_b200: pu_chain group barrier timeout\n"); __trap(); }
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

// W_hh rows permuted so that CTA slice sl owns 64 consecutive rows: perm[sl*64 + gate*16 + u] = W[gate*512 + sl*16 + u]
__global__ void __launch_bounds__(128) pu_permute_split_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ hi,
                                                               __nv_bfloat16* __restrict__ lo) {
  const int prow = blockIdx.x;                       // 0..2047
  const int sl = prow / PC_N, gate = (prow % PC_N) / PC_U, u = prow % PC_U;
  const float4 v = reinterpret_cast<const float4*>(W + (long long)(gate * PC_H + sl * PC_U + u) * PC_H)[threadIdx.x];
  uint32_t h0, h1, l0, l1;
  split_pack2(v.x, v.y, h0, l0);
  split_pack2(v.z, v.w, h1, l1);
  reinterpret_cast<uint2*>(hi + (long long)prow * PC_H)[threadIdx.x] = make_uint2(h0, h1);
  if (lo) reinterpret_cast<uint2*>(lo + (long long)prow * PC_H)[threadIdx.x] = make_uint2(l0, l1);
}

int pu_permute_split_run(const float* W, __nv_bfloat16* hi, __nv_bfloat16* lo, cudaStream_t stream) {
  ProfScope prof("pu_permute_split_kernel", stream);
  EB_LAUNCH(pu_permute_split_kernel, 4 * PC_H, 128, stream, W, hi, lo);
  EB_CHECK_LAUNCH("pu_permute_split_kernel");
  return 0;
}

template <int NSPLIT>
static int launch_pu(const CUtensorMap* tm, const PuParams& p, int groups, cudaStream_t stream) {
  using C = PuCfg<NSPLIT>;
  auto kern = pu_chain_kernel<NSPLIT>;
  // the opt-in to > 48 KB dynamic shared memory is a per-device function attribute
  static bool attr_done[64] = {false};
  const int dev_ = current_device();
  if (!attr_done[dev_]) {
    EB_CUDA(EB_SET_MAX_SMEM(kern, C::SMEM_BYTES));
    attr_done[dev_] = true;
  }
  ProfScope prof("pu_chain_kernel", stream);
  EB_LAUNCH_GRID_SYNC(kern, groups * PC_SLICES, 192, C::SMEM_BYTES, stream, tm[0], tm[1], tm[2], tm[3], p);
  EB_CHECK_LAUNCH("pu_chain_kernel");
  return 0;
}

// One propagation-unit layer over all J joints.  w_perm: gate-permuted W_hh (2048 x 512) bf16 hi/lo;
// hg: scratch [2][B][512] bf16 hi/lo; counters: >= ceil(B/256) zero-initialisable words.
int pu_chain_run(const __nv_bfloat16* w_hi, const __nv_bfloat16* w_lo, const float* G, long long G_rs, long long G_ts,
                 const float* F, long long F_rs, long long F_ts, float* out, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo,
                 __nv_bfloat16* hg_hi, __nv_bfloat16* hg_lo, unsigned int* counters, int B, int J, int nsplit,
                 cudaStream_t stream) {
  EB_REQUIRE(w_hi && G && F && out && hg_hi && counters && B > 0 && J > 0, "pu_chain: bad arguments");
  EB_REQUIRE(nsplit == 1 || (w_lo && hg_lo), "pu_chain: bf16x3 mode needs the lo parts");
  const int groups = (B + PC_RPG - 1) / PC_RPG;
  EB_REQUIRE(groups * PC_SLICES <= num_sms(), "pu_chain: batch %d needs %d co-resident CTAs (> %d SMs); chunk the batch", B,
             groups * PC_SLICES, num_sms());
  CUtensorMap tm[4];
  int rc;
  if ((rc = make_operand_tmap(&tm[0], w_hi, PC_H, 4 * PC_H, PC_H, 1, 0, 1, 0, PC_N))) return rc;
  if ((rc = make_operand_tmap(&tm[2], hg_hi, PC_H, 2ll * B, PC_H, 1, 0, 1, 0, 128))) return rc;
  if (nsplit == 3) {
    if ((rc = make_operand_tmap(&tm[1], w_lo, PC_H, 4 * PC_H, PC_H, 1, 0, 1, 0, PC_N))) return rc;
    if ((rc = make_operand_tmap(&tm[3], hg_lo, PC_H, 2ll * B, PC_H, 1, 0, 1, 0, 128))) return rc;
  } else {
    tm[1] = tm[0];
    tm[3] = tm[2];
  }
  EB_CUDA(cudaMemsetAsync(counters, 0, sizeof(unsigned int) * groups, stream));
  PuParams p{G, G_rs, G_ts, F, F_rs, F_ts, out, out_hi, out_lo, hg_hi, hg_lo, counters, B, J};
  return nsplit == 3 ? launch_pu<3>(tm, p, groups, stream) : launch_pu<1>(tm, p, groups, stream);
}

}  // namespace eb

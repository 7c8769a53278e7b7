// Fused backward of the multi-head softmax attention (bf16-operand training mode), sm_100a.
//   forward (reference model/modeling_vit.py:233-252):  S = Q K^T / sqrt(128),  P = softmax(S),  O = P V
//   backward:  dV = P^T dO,   dP = dO V^T,   dS = P o (dP - D) / sqrt(128) with D = rowsum(dO o O),   dQ = dS K,   dK = dS^T Q
//
// It replaces, per ViT layer, what the first training path did with library-style pieces (TrainEngine._attention_bwd): two
// grouped GEMMs that wrote the 576 x 576 scores and dP of every (frame, head) to HBM in fp32, softmax_bwd, two 576 x 576 bf16
// transposes, four operand transposes and three more grouped GEMMs -- 27 % of the first measured training step (B200, batch 256:
// 9.6 + 4.5 ms of GEMMs at 109 / 346 TFLOP/s, 3.5 ms softmax_bwd, ~3.5 ms transposes).  Here the scores, probabilities and their
// gradients exist only in tensor memory / registers; HBM sees Q, K, V^T, dO once per tile pass and dQ, dK, dV once.
//
// Two kernels, both persistent, one CTA per SM, 128 x 128 tiles, all MMAs 128 x 128 x 16:
//   attn_bwd_dkv_kernel  work item = (frame, head, 128-key block j); loops over the query blocks i.  Tensor-memory lanes = keys:
//       S^T = K_j Q_i^T and dP^T = V_j dO_i^T (SS MMAs), the compute warps turn them into P^T and dS^T IN PLACE (packed bf16 over
//       the fp32 columns they just read), then dV_j += P^T dO_i and dK_j += dS^T Q_i (TS MMAs: A from tensor memory).
//   attn_bwd_dq_kernel   work item = (frame, head, 128-query block i); loops over the key blocks j.  Lanes = queries:
//       S = Q_i K_j^T, dP = dO_i V_j^T, dS in place, dQ_i += dS K_j.
// The scores are computed twice (7 MMA groups per tile pair instead of 5) so that every accumulator has ONE owner: no atomics,
// no partial-sum buffers, bit-identical runs.
//
// No operand is ever transposed in memory.  Every tile is loaded by TMA exactly as the forward loads it (rows x 64 bf16, 128-byte
// swizzle) and used in BOTH roles by choosing the major-ness in the MMA descriptor:
//   K-major  (rows = the MMA's M / N index, the contraction runs along the 128-byte row):  Q_i, K_j, dO_i in S / S^T / dP / dP^T
//   MN-major (rows = the contraction index, M / N runs along the row):  dO_i in dV, Q_i in dK, K_j in dQ (N = d), the stored V^T
//            tile (128 d rows x keys) as V_j in dP^T (A operand, M = keys) and dP (B operand, N = keys)
// MN-major, 128-byte swizzle (canonical layout ((8,8,m),(8,k)) : ((1,8,LBO),(64,SBO)) in bf16 elements): 64 M/N-elements per
// 128-byte row, 8 rows = one 1 KB swizzle atom (SBO = 1024), the next 64 M/N-elements = the next TMA box (LBO = 16 KB here);
// a k-step of 16 advances the start address by 16 rows = 2 KB.
//
// Row statistics come from the forward: lse[q] = max * c + log2(sum) in the exp2 domain (c = log2(e) / sqrt(128)), written by
// attention_kernel when asked to, and D[q] from attn_dsum_kernel below.
//   warp 0  TMA producer     warp 1  MMA issuer     warps 2-9  compute + epilogue (pairs split the 128 columns of a tile)
#include "gemm.cuh"
#include "host_util.cuh"
#include "internal.h"

namespace eb {

constexpr int AB_TOK = 576, AB_HEADS = 8, AB_D = 128, AB_T = 128;
constexpr int AB_NB = (AB_TOK + AB_T - 1) / AB_T;              // 5 blocks of 128 tokens, the last one has 64 valid
constexpr int AB_BLK = 128 * 64 * 2;                            // one TMA box: 128 rows x 64 bf16 = 16 KB
constexpr int AB_TILE = 2 * AB_BLK;                             // a 128 x 128 operand tile = two boxes (column blocks)
constexpr int AB_THREADS = 64 + 256;
constexpr float AB_SCALE = 0.08838834764831845f;                // 1 / sqrt(128)
constexpr float AB_C = 0.08838834764831845f * 1.4426950408889634f;

// shared-memory descriptors (high word: SBO = 1024 B, descriptor version 1, 128-byte swizzle -- see ptx.cuh)
__device__ __forceinline__ uint32_t ab_lo_kmajor(uint32_t smem_addr) { return sdesc_lo(smem_addr); }
__device__ __forceinline__ uint32_t ab_lo_mnmajor(uint32_t smem_addr) {          // LBO = one TMA box (16 KB) between 64-wide M/N groups
  return ((smem_addr >> 4) & 0x3FFFu) | (uint32_t(AB_BLK >> 4) << 16);
}
// K-major tile (two column boxes): k-step kk of 16 covers columns 16 kk .. 16 kk + 15 = box kk / 4, 32 bytes per step inside a row
__device__ __forceinline__ uint64_t ab_desc_k(uint32_t lo, int kk) { return sdesc_at(lo, (kk >> 2) * AB_BLK + (kk & 3) * 32); }
// MN-major tile: k-step kk covers rows 16 kk .. 16 kk + 15 = 2 KB further
__device__ __forceinline__ uint64_t ab_desc_mn(uint32_t lo, int kk) { return sdesc_at(lo, kk * 2048); }

struct AttnBwdSmem {
  static constexpr int RES_OFF = 0;                             // resident pair of tiles (K_j, V^T_j  |  Q_i, dO_i)
  static constexpr int RING_OFF = 2 * AB_TILE;                  // 2 stages x 2 streamed tiles
  static constexpr int BAR_OFF = RING_OFF + 4 * AB_TILE;        // 192 KB of tiles
  static constexpr int STAT_OFF = BAR_OFF + 256;                // [2][128] lse, [2][128] D * scale (dkv kernel)
  static constexpr int BYTES = STAT_OFF + 2048;
  static_assert(BYTES <= 227 * 1024, "shared memory budget");
};

// ------------------------------------------------------------------------------------------------------------------
// dK / dV
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmQK, const __grid_constant__ CUtensorMap tmV,
                    const __grid_constant__ CUtensorMap tmDO, const float* __restrict__ lse, const float* __restrict__ dsum,
                    float* __restrict__ dqkv, int num_items) {
  using S = AttnBwdSmem;
  EB_DYN_SMEM_1K(smem);
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sK = smem + S::RES_OFF;                  // K_j   [128 keys][128 d]    (two d boxes)
  uint8_t* sV = sK + AB_TILE;                       // V^T_j [128 d][128 keys]    (two key boxes)
  uint8_t* sRing = smem + S::RING_OFF;              // stage s: Q_i [128 q][128 d], dO_i [128 q][128 d]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  uint64_t* kv_full = bars;            // [1]
  uint64_t* kv_empty = bars + 1;       // [1]  last S^T / dP^T of the item issued and complete
  uint64_t* qd_full = bars + 2;        // [2]
  uint64_t* qd_empty = bars + 4;       // [2]  dV / dK MMAs of that stage complete
  uint64_t* s_full = bars + 6;         // [1]  S^T_i and dP^T_i in tensor memory
  uint64_t* p_full = bars + 7;         // [1]  P^T_i and dS^T_i in tensor memory (8 warps)
  uint64_t* acc_full = bars + 8;       // [1]  dV_j, dK_j complete
  uint64_t* acc_empty = bars + 9;      // [1]  ... drained (8 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  float* sL = reinterpret_cast<float*>(smem + S::STAT_OFF);      // [2][128]
  float* sD = sL + 256;                                           // [2][128]  D * scale
  constexpr uint32_t T_ST = 0, T_DPT = 128, T_DV = 256, T_DK = 384;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int my_items = (num_items - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmQK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO); }
  if (warp == 1 && lane == 0) {
    mbar_init(kv_full, 1); mbar_init(kv_empty, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&qd_full[i], 1); mbar_init(&qd_empty[i], 1); }
    mbar_init(s_full, 1); mbar_init(p_full, 8); mbar_init(acc_full, 1); mbar_init(acc_empty, 8);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<1>(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (*tmem_slot != 0) __trap();

  auto coords = [&](int it, int& j, int& h, int& b, int& bh) {
    const int item = int(blockIdx.x) + it * int(gridDim.x);
    j = item % AB_NB; bh = item / AB_NB; h = bh % AB_HEADS; b = bh / AB_HEADS;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t n = 0;                                   // streamed stage counter
      for (int it = 0; it < my_items; ++it) {
        int j, h, b, bh;
        coords(it, j, h, b, bh);
        mbar_wait(kv_empty, (it & 1) ^ 1);
        mbar_expect_tx(kv_full, 2 * AB_TILE);
        for (int kb = 0; kb < 2; ++kb) {
          tma_load_4d(sK + kb * AB_BLK, &tmQK, kv_full, kb * 64, j * AB_T, AB_HEADS + h, b);      // K_j, d box kb
          tma_load_4d(sV + kb * AB_BLK, &tmV, kv_full, j * AB_T + kb * 64, bh * AB_D, 0, 0);      // V^T_j, key box kb
        }
        for (int i = 0; i < AB_NB; ++i, ++n) {
          const uint32_t st = n & 1;
          mbar_wait(&qd_empty[st], ((n >> 1) & 1) ^ 1);
          uint8_t* q = sRing + st * 2 * AB_TILE;
          mbar_expect_tx(&qd_full[st], 2 * AB_TILE);
          for (int kb = 0; kb < 2; ++kb) {
            tma_load_4d(q + kb * AB_BLK, &tmQK, &qd_full[st], kb * 64, i * AB_T, h, b);            // Q_i
            tma_load_4d(q + AB_TILE + kb * AB_BLK, &tmDO, &qd_full[st], kb * 64, i * AB_T, h, b);  // dO_i
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t id_kk = make_idesc_bf16(AB_T, AB_T, 0, 0);      // A K-major, B K-major
    constexpr uint32_t id_mk = make_idesc_bf16(AB_T, AB_T, 1, 0);      // A MN-major (V^T tile as V), B K-major
    constexpr uint32_t id_xm = make_idesc_bf16(AB_T, AB_D, 0, 1);      // A from tensor memory, B MN-major
    const uint32_t k_lo = ab_lo_kmajor(smem_u32(sK)), v_lo = ab_lo_mnmajor(smem_u32(sV));
    uint32_t n = 0;
    for (int it = 0; it < my_items; ++it) {
      mbar_wait(kv_full, it & 1);
      for (int i = 0; i < AB_NB; ++i, ++n) {
        const uint32_t st = n & 1;
        const uint32_t q_addr = smem_u32(sRing + st * 2 * AB_TILE);
        const uint32_t qk_lo = ab_lo_kmajor(q_addr), dok_lo = ab_lo_kmajor(q_addr + AB_TILE);
        const uint32_t qm_lo = ab_lo_mnmajor(q_addr), dom_lo = ab_lo_mnmajor(q_addr + AB_TILE);
        mbar_wait(&qd_full[st], (n >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)          // S^T = K_j Q_i^T
          umma_bf16<1>(T_ST, ab_desc_k(k_lo, kk), ab_desc_k(qk_lo, kk), id_kk, kk != 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)          // dP^T = V_j dO_i^T
          umma_bf16<1>(T_DPT, ab_desc_mn(v_lo, kk), ab_desc_k(dok_lo, kk), id_mk, kk != 0 ? 1u : 0u);
        umma_commit<1>(s_full);
        if (i == AB_NB - 1) umma_commit<1>(kv_empty);
        mbar_wait(p_full, n & 1);
        if (i == 0) mbar_wait(acc_empty, (it & 1) ^ 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {        // k-step = 16 queries: 8 packed columns, the pair's halves start at +0 / +64
          const uint32_t a_off = uint32_t((kk >> 2) * 64 + (kk & 3) * 8);
          umma_bf16_ts(T_DV, T_ST + a_off, ab_desc_mn(dom_lo, kk), id_xm, (i | kk) != 0 ? 1u : 0u);    // dV += P^T dO_i
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t a_off = uint32_t((kk >> 2) * 64 + (kk & 3) * 8);
          umma_bf16_ts(T_DK, T_DPT + a_off, ab_desc_mn(qm_lo, kk), id_xm, (i | kk) != 0 ? 1u : 0u);    // dK += dS^T Q_i
        }
        umma_commit<1>(&qd_empty[st]);
      }
      umma_commit<1>(acc_full);
    }
  } else {
    // ------------------------------------------------------------------ compute warps: lanes = keys, columns = queries
    const int q = warp & 3, hf = (warp - 2) >> 2;
    const uint32_t lane_sel = uint32_t(q * 32) << 16;
    const int ct = int(threadIdx.x) - 64;               // 0 .. 255
    uint32_t n = 0;
    for (int it = 0; it < my_items; ++it) {
      int j, h, b, bh;
      coords(it, j, h, b, bh);
      for (int i = 0; i < AB_NB; ++i, ++n) {
        // row statistics of the 128 queries of block i: one value per thread, staged in shared memory (double-buffered)
        {
          const int qq = ct & 127, tok = i * AB_T + qq;
          const long long o = (long long)bh * AB_TOK + tok;
          float v;
          if (ct < 128) v = tok < AB_TOK ? lse[o] : 1e30f;               // exp2(-1e30) = 0: no contribution
          else v = tok < AB_TOK ? dsum[o] * AB_SCALE : 0.0f;
          (ct < 128 ? sL : sD)[(n & 1) * 128 + qq] = v;
        }
        named_bar_sync<256>(1);
        mbar_wait(s_full, n & 1);
        tc_fence_after();
        const float* L = sL + (n & 1) * 128 + hf * 64;
        const float* Dq = sD + (n & 1) * 128 + hf * 64;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t s[32], dp[32];
          tmem_ld32(T_ST + lane_sel + hf * 64 + c * 32, s);
          tmem_ld32(T_DPT + lane_sel + hf * 64 + c * 32, dp);
          tmem_ld_wait();
          uint32_t pp[16], ds[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float2 l2 = *reinterpret_cast<const float2*>(L + c * 32 + 2 * e);
            const float2 d2 = *reinterpret_cast<const float2*>(Dq + c * 32 + 2 * e);
            const float p0 = ex2_approx(fmaf(__uint_as_float(s[2 * e]), AB_C, -l2.x));
            const float p1 = ex2_approx(fmaf(__uint_as_float(s[2 * e + 1]), AB_C, -l2.y));
            const float g0 = p0 * fmaf(__uint_as_float(dp[2 * e]), AB_SCALE, -d2.x);
            const float g1 = p1 * fmaf(__uint_as_float(dp[2 * e + 1]), AB_SCALE, -d2.y);
            pp[e] = cvt_bf16x2(p0, p1);
            ds[e] = cvt_bf16x2(g0, g1);
          }
          // packed over the fp32 columns this warp has already read (its own 64-column half)
          tmem_st16(T_ST + lane_sel + hf * 64 + c * 16, pp);
          tmem_st16(T_DPT + lane_sel + hf * 64 + c * 16, ds);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
      }
      // ---- epilogue: dV_j, dK_j rows (keys) -> dqkv[(b, key), 2048 + h*128 + d] / [.., 1024 + h*128 + d], fp32
      mbar_wait(acc_full, it & 1);
      tc_fence_after();
      const int key = j * AB_T + q * 32 + lane;
      float* orow = dqkv + ((long long)b * AB_TOK + key) * (3 * AB_HEADS * AB_D) + h * AB_D + hf * 64;
#pragma unroll 1
      for (int m = 0; m < 2; ++m) {              // 0: dV, 1: dK
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t o[32];
          tmem_ld32((m == 0 ? T_DV : T_DK) + lane_sel + hf * 64 + c * 32, o);
          tmem_ld_wait();
          if (key < AB_TOK) {
            float* dst = orow + (m == 0 ? 2 : 1) * (AB_HEADS * AB_D) + c * 32;
#pragma unroll
            for (int e = 0; e < 4; ++e)
              st_global_256(dst + 8 * e, o[8 * e], o[8 * e + 1], o[8 * e + 2], o[8 * e + 3], o[8 * e + 4], o[8 * e + 5], o[8 * e + 6],
                            o[8 * e + 7]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(0u, 512);
}

// ------------------------------------------------------------------------------------------------------------------
// dQ
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQK, const __grid_constant__ CUtensorMap tmV,
                   const __grid_constant__ CUtensorMap tmDO, const float* __restrict__ lse, const float* __restrict__ dsum,
                   float* __restrict__ dqkv, int num_items) {
  using S = AttnBwdSmem;
  EB_DYN_SMEM_1K(smem);
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem + S::RES_OFF;                  // Q_i  [128 q][128 d]
  uint8_t* sDO = sQ + AB_TILE;                      // dO_i [128 q][128 d]
  uint8_t* sRing = smem + S::RING_OFF;              // stage s: K_j [128 keys][128 d], V^T_j [128 d][128 keys]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  uint64_t* q_full = bars;             // [1]
  uint64_t* q_empty = bars + 1;        // [1]  last S / dP of the item complete
  uint64_t* kv_full = bars + 2;        // [2]
  uint64_t* kv_empty = bars + 4;       // [2]  dQ MMAs of that stage complete
  uint64_t* s_full = bars + 6;
  uint64_t* p_full = bars + 7;         // (8 warps)
  uint64_t* acc_full = bars + 8;
  uint64_t* acc_empty = bars + 9;      // (8 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  constexpr uint32_t T_S = 0, T_DP = 128, T_DQ = 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int my_items = (num_items - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmQK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO); }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1); mbar_init(q_empty, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(s_full, 1); mbar_init(p_full, 8); mbar_init(acc_full, 1); mbar_init(acc_empty, 8);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<1>(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (*tmem_slot != 0) __trap();

  auto coords = [&](int it, int& i, int& h, int& b, int& bh) {
    const int item = int(blockIdx.x) + it * int(gridDim.x);
    i = item % AB_NB; bh = item / AB_NB; h = bh % AB_HEADS; b = bh / AB_HEADS;
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t n = 0;
      for (int it = 0; it < my_items; ++it) {
        int i, h, b, bh;
        coords(it, i, h, b, bh);
        mbar_wait(q_empty, (it & 1) ^ 1);
        mbar_expect_tx(q_full, 2 * AB_TILE);
        for (int kb = 0; kb < 2; ++kb) {
          tma_load_4d(sQ + kb * AB_BLK, &tmQK, q_full, kb * 64, i * AB_T, h, b);
          tma_load_4d(sDO + kb * AB_BLK, &tmDO, q_full, kb * 64, i * AB_T, h, b);
        }
        for (int j = 0; j < AB_NB; ++j, ++n) {
          const uint32_t st = n & 1;
          mbar_wait(&kv_empty[st], ((n >> 1) & 1) ^ 1);
          uint8_t* k = sRing + st * 2 * AB_TILE;
          mbar_expect_tx(&kv_full[st], 2 * AB_TILE);
          for (int kb = 0; kb < 2; ++kb) {
            tma_load_4d(k + kb * AB_BLK, &tmQK, &kv_full[st], kb * 64, j * AB_T, AB_HEADS + h, b);            // K_j
            tma_load_4d(k + AB_TILE + kb * AB_BLK, &tmV, &kv_full[st], j * AB_T + kb * 64, bh * AB_D, 0, 0);  // V^T_j
          }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t id_kk = make_idesc_bf16(AB_T, AB_T, 0, 0);      // S  = Q_i K_j^T
    constexpr uint32_t id_km = make_idesc_bf16(AB_T, AB_T, 0, 1);      // dP = dO_i V_j^T: B = V^T tile, MN-major (N = keys)
    constexpr uint32_t id_xm = make_idesc_bf16(AB_T, AB_D, 0, 1);      // dQ += dS K_j: A from tensor memory, B = K_j MN-major (N = d)
    const uint32_t q_lo = ab_lo_kmajor(smem_u32(sQ)), do_lo = ab_lo_kmajor(smem_u32(sDO));
    uint32_t n = 0;
    for (int it = 0; it < my_items; ++it) {
      mbar_wait(q_full, it & 1);
      for (int j = 0; j < AB_NB; ++j, ++n) {
        const uint32_t st = n & 1;
        const uint32_t k_addr = smem_u32(sRing + st * 2 * AB_TILE);
        const uint32_t kk_lo = ab_lo_kmajor(k_addr), km_lo = ab_lo_mnmajor(k_addr), vm_lo = ab_lo_mnmajor(k_addr + AB_TILE);
        mbar_wait(&kv_full[st], (n >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_bf16<1>(T_S, ab_desc_k(q_lo, kk), ab_desc_k(kk_lo, kk), id_kk, kk != 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_bf16<1>(T_DP, ab_desc_k(do_lo, kk), ab_desc_mn(vm_lo, kk), id_km, kk != 0 ? 1u : 0u);
        umma_commit<1>(s_full);
        if (j == AB_NB - 1) umma_commit<1>(q_empty);
        mbar_wait(p_full, n & 1);
        if (j == 0) mbar_wait(acc_empty, (it & 1) ^ 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t a_off = uint32_t((kk >> 2) * 64 + (kk & 3) * 8);
          umma_bf16_ts(T_DQ, T_DP + a_off, ab_desc_mn(km_lo, kk), id_xm, (j | kk) != 0 ? 1u : 0u);
        }
        umma_commit<1>(&kv_empty[st]);
      }
      umma_commit<1>(acc_full);
    }
  } else {
    // compute warps: lanes = queries, columns = keys
    const int q = warp & 3, hf = (warp - 2) >> 2;
    const uint32_t lane_sel = uint32_t(q * 32) << 16;
    uint32_t n = 0;
    for (int it = 0; it < my_items; ++it) {
      int i, h, b, bh;
      coords(it, i, h, b, bh);
      const int tok = i * AB_T + q * 32 + lane;
      const long long so = (long long)bh * AB_TOK + tok;
      const float l2 = tok < AB_TOK ? lse[so] : 1e30f;
      const float dsc = tok < AB_TOK ? dsum[so] * AB_SCALE : 0.0f;
      for (int j = 0; j < AB_NB; ++j, ++n) {
        mbar_wait(s_full, n & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t s[32], dp[32];
          tmem_ld32(T_S + lane_sel + hf * 64 + c * 32, s);
          tmem_ld32(T_DP + lane_sel + hf * 64 + c * 32, dp);
          tmem_ld_wait();
          const bool keys_valid = j * AB_T + hf * 64 + c * 32 < AB_TOK;      // the tail block has 64 keys (warp-uniform)
          uint32_t ds[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float p0 = ex2_approx(fmaf(__uint_as_float(s[2 * e]), AB_C, -l2));
            const float p1 = ex2_approx(fmaf(__uint_as_float(s[2 * e + 1]), AB_C, -l2));
            const float g0 = p0 * fmaf(__uint_as_float(dp[2 * e]), AB_SCALE, -dsc);
            const float g1 = p1 * fmaf(__uint_as_float(dp[2 * e + 1]), AB_SCALE, -dsc);
            ds[e] = keys_valid ? cvt_bf16x2(g0, g1) : 0u;
          }
          tmem_st16(T_DP + lane_sel + hf * 64 + c * 16, ds);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
      }
      // ---- epilogue: dQ_i rows -> dqkv[(b, q), h*128 + d]
      mbar_wait(acc_full, it & 1);
      tc_fence_after();
      float* orow = dqkv + ((long long)b * AB_TOK + tok) * (3 * AB_HEADS * AB_D) + h * AB_D + hf * 64;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t o[32];
        tmem_ld32(T_DQ + lane_sel + hf * 64 + c * 32, o);
        tmem_ld_wait();
        if (tok < AB_TOK) {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            st_global_256(orow + c * 32 + 8 * e, o[8 * e], o[8 * e + 1], o[8 * e + 2], o[8 * e + 3], o[8 * e + 4], o[8 * e + 5],
                          o[8 * e + 6], o[8 * e + 7]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(0u, 512);
}

// ------------------------------------------------------------------------------------------------------------------
// D[b, h, q] = sum_d dO[b, q, h, d] * O[b, q, h, d]   (one warp per (token, head); O and dO as bf16 hi (+ lo) rows of 1024)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attn_dsum_kernel(const __nv_bfloat16* __restrict__ o_hi, const __nv_bfloat16* __restrict__ o_lo,
                                                        const __nv_bfloat16* __restrict__ do_hi, const __nv_bfloat16* __restrict__ do_lo,
                                                        long long rows, float* __restrict__ dsum) {
  const long long w = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);      // (token row, head)
  const int lane = threadIdx.x & 31;
  if (w >= rows * AB_HEADS) return;
  const long long r = w / AB_HEADS;
  const int h = int(w % AB_HEADS);
  const long long off = r * (AB_HEADS * AB_D) + h * AB_D + lane * 4;
  auto ld4 = [&](const __nv_bfloat16* p, float (&v)[4]) {
    const uint2 u = *reinterpret_cast<const uint2*>(p + off);
    v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
    v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
  };
  float a[4], g[4], t[4];
  ld4(o_hi, a);
  ld4(do_hi, g);
  if (o_lo) { ld4(o_lo, t); for (int i = 0; i < 4; ++i) a[i] += t[i]; }
  if (do_lo) { ld4(do_lo, t); for (int i = 0; i < 4; ++i) g[i] += t[i]; }
  float s = a[0] * g[0] + a[1] * g[1] + a[2] * g[2] + a[3] * g[3];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  if (lane == 0) {
    const long long b = r / AB_TOK, tok = r % AB_TOK;
    dsum[(b * AB_HEADS + h) * AB_TOK + tok] = s;
  }
}

int attention_bwd_run(const __nv_bfloat16* qk, const __nv_bfloat16* vt, const __nv_bfloat16* dctx, const float* lse,
                      const float* dsum, float* dqkv, int B, cudaStream_t stream) {
  EB_REQUIRE(qk && vt && dctx && lse && dsum && dqkv && B > 0, "attention_bwd: bad arguments");
  EB_REQUIRE((reinterpret_cast<uintptr_t>(dqkv) & 31) == 0, "attention_bwd: dqkv must be 32-byte aligned");
  CUtensorMap tmQK, tmV, tmDO;
  int rc;
  const long long fs = (long long)AB_TOK * 2 * AB_HEADS * AB_D;
  if ((rc = make_operand_tmap(&tmQK, qk, AB_D, AB_TOK, 2 * AB_HEADS * AB_D, 2 * AB_HEADS, AB_D, B, fs, 128))) return rc;
  if ((rc = make_operand_tmap(&tmV, vt, AB_TOK, (long long)B * AB_HEADS * AB_D, AB_TOK, 1, 0, 1, 0, 128))) return rc;
  if ((rc = make_operand_tmap(&tmDO, dctx, AB_D, AB_TOK, AB_HEADS * AB_D, AB_HEADS, AB_D, B, (long long)AB_TOK * AB_HEADS * AB_D, 128)))
    return rc;
  static bool attr_done[64] = {false};
  const int dev_ = current_device();
  if (!attr_done[dev_]) {
    EB_CUDA(EB_SET_MAX_SMEM(attn_bwd_dkv_kernel, AttnBwdSmem::BYTES));
    EB_CUDA(EB_SET_MAX_SMEM(attn_bwd_dq_kernel, AttnBwdSmem::BYTES));
    attr_done[dev_] = true;
  }
  const int items = B * AB_HEADS * AB_NB;
  const int grid = items < num_sms() ? items : num_sms();
  {
    ProfScope prof("attn_bwd_dkv_kernel", stream);
    EB_LAUNCH_SMEM(attn_bwd_dkv_kernel, grid, AB_THREADS, AttnBwdSmem::BYTES, stream, tmQK, tmV, tmDO, lse, dsum, dqkv, items);
    EB_CHECK_LAUNCH("attn_bwd_dkv_kernel");
  }
  {
    ProfScope prof("attn_bwd_dq_kernel", stream);
    EB_LAUNCH_SMEM(attn_bwd_dq_kernel, grid, AB_THREADS, AttnBwdSmem::BYTES, stream, tmQK, tmV, tmDO, lse, dsum, dqkv, items);
    EB_CHECK_LAUNCH("attn_bwd_dq_kernel");
  }
  return 0;
}

int attn_dsum_run(const __nv_bfloat16* o_hi, const __nv_bfloat16* o_lo, const __nv_bfloat16* do_hi, const __nv_bfloat16* do_lo,
                  long long rows, float* dsum, cudaStream_t stream) {
  EB_REQUIRE(o_hi && do_hi && dsum && rows > 0 && rows % AB_TOK == 0, "attn_dsum: bad arguments (rows %lld)", rows);
  ProfScope prof("attn_dsum_kernel", stream);
  EB_LAUNCH_COOP(attn_dsum_kernel, (unsigned)((rows * AB_HEADS + 7) / 8), 256, stream, o_hi, o_lo, do_hi, do_lo, rows, dsum);
  EB_CHECK_LAUNCH("attn_dsum_kernel");
  return 0;
}

}  // namespace eb

/* C ABI (include/egotap_b200.h) */
extern "C" int egotap_b200_attention_bwd(const void* qk_hi, const void* vt_hi, const void* dctx_hi, const float* lse,
                                         const float* dsum, float* dqkv, int frames, void* stream) {
  return eb::attention_bwd_run((const __nv_bfloat16*)qk_hi, (const __nv_bfloat16*)vt_hi, (const __nv_bfloat16*)dctx_hi, lse, dsum,
                               dqkv, frames, (cudaStream_t)stream);
}
extern "C" int egotap_b200_attn_dsum(const void* ctx_hi, const void* ctx_lo, const void* dctx_hi, const void* dctx_lo,
                                     long long rows, float* dsum, void* stream) {
  return eb::attn_dsum_run((const __nv_bfloat16*)ctx_hi, (const __nv_bfloat16*)ctx_lo, (const __nv_bfloat16*)dctx_hi,
                           (const __nv_bfloat16*)dctx_lo, rows, dsum, (cudaStream_t)stream);
}

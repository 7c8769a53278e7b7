// Fused multi-head self-attention for the ViT heatmap encoder (576 tokens, 8 heads x 128), sm_100a.
//   ctx[b, q, h*128 + :] = softmax(Q K^T / sqrt(128)) V        reference model/modeling_vit.py:233-252
//
// One persistent CTA per SM walks work items (frame b, head h, 128-row query tile).  Per item the 576 keys
// are processed in 9 tiles of 64:
//   warp 0   TMA producer: Q tile once, then K_j / V^T_j tiles through a shared-memory ring
//   warp 1   MMA issuer:   S_j = Q K_j^T  (tcgen05, fp32 in TMEM, double-buffered),  O += P_j V_j
//   warps 2-9 softmax:     (two warps per TMEM lane quarter, splitting the 64 keys of a tile)
//                          tcgen05.ld S_j -> online softmax in the exp2 domain (lazy rescale of O in TMEM only
//                          when a row maximum grows by more than 2^8) -> P_j written back to TENSOR MEMORY as the
//                          packed-bf16 A operand of the P.V MMA (tcgen05.mma TS form: no smem traffic for P)
//   warps 10-13 epilogue:  O (double-buffered in TMEM across items) / l -> bf16 hi/lo context rows
// NSPLIT = 3: Q, K, V and P are bf16 hi/lo pairs and every product is 3 MMAs (fp32-parity mode).
// The score matrix never leaves the SM: HBM traffic is Q, K, V in and ctx out.
#include "gemm.cuh"
#include "host_util.cuh"
#include "internal.h"

#include <cstdlib>
#include <cstring>

namespace eb {

// tmem_st32 / tmem_st16 / umma_bf16_ts / tmem_st_wait: ptx.cuh

// Debug-only wait-cycle accounting (tools/attn_trace.sh builds a separate library with -DEB_ATTN_TRACE).
#ifdef EB_ATTN_TRACE
__device__ unsigned long long g_attn_trace[32];
#define TR_DECL long long tr_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long tr_t0_ = 0; (void)tr_t0_;
#define TR_WAIT(i, stmt) { tr_t0_ = clock64(); stmt; tr_[i] += clock64() - tr_t0_; }
#define TR_FLUSH(base, cond) if (cond) { for (int i_ = 0; i_ < 8; ++i_) atomicAdd(&g_attn_trace[(base) + i_], (unsigned long long)tr_[i_]); }
#else
#define TR_DECL
#define TR_WAIT(i, stmt) { stmt; }
#define TR_FLUSH(base, cond)
#endif

constexpr int AT_TOK = 576, AT_HEADS = 8, AT_D = 128, AT_QT = 128, AT_KT = 64;
constexpr int AT_NT = AT_TOK / AT_KT;                        // 9 key tiles
constexpr int AT_THREADS = 64 + 256 + 128;                   // TMA + MMA warps, 8 softmax warps, 4 epilogue warps

template <int NSPLIT>
struct AttnCfg {
  static constexpr int NOPS = NSPLIT == 1 ? 1 : 2;
  static constexpr int Q_BLK = AT_QT * 64 * 2;               // one 64-wide d block of the Q tile (16 KB)
  static constexpr int Q_BYTES = NOPS * 2 * Q_BLK;
  static constexpr int K_BLK = AT_KT * 64 * 2;               // one 64-wide d block of a K tile (8 KB)
  static constexpr int V_BLK = AT_D * 64 * 2;                // V^T tile: 128 d rows x 64 keys (16 KB)
  static constexpr int SLOT_BYTES = NOPS * V_BLK;            // K tile (2 d blocks) and V^T tile have the same size
  // The ring carries 2 * 9 = 18 tiles per item; a slot count that divides 18 makes the slot of every tile a
  // compile-time constant in the unrolled MMA issue loop.
  static constexpr int NSLOTS = NSPLIT == 1 ? 6 : 3;
  static constexpr int USES = 2 * AT_NT / NSLOTS;            // ring-slot uses per item
  static_assert((2 * AT_NT) % NSLOTS == 0, "ring slots must divide the tiles per item");
  static constexpr int BAR_OFF = Q_BYTES + NSLOTS * SLOT_BYTES;
  static constexpr int XCHG_OFF = BAR_OFF + 256;             // [2][2][128] row maxima, then [2][2][128] row sums
  static constexpr int SMEM_BYTES = XCHG_OFF + 4096;         // base must be 1 KB aligned (checked)
  // TMEM (all 512 columns, so the base is column 0): S 2 x 64 at [0,128), O 2 x 128 at [128,384),
  // P 2 x (32 hi + 32 lo packed bf16x2 columns) at [384,512)
  static constexpr int TMEM_COLS = 512;
  static constexpr uint32_t T_S = 0, T_O = 128, T_P = 384;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

// S / P double buffers: tile j of an item uses buffer j & 1.  Buffer 0 is used 5 times per item (tiles 0,2,4,6,8),
// buffer 1 four times, so the mbarrier phase parity of tile j's use in the it-th item of this CTA is:
__device__ __forceinline__ uint32_t buf_parity(int j, int it) { return uint32_t((j >> 1) + ((j & 1) ? 0 : it)) & 1u; }

template <int NSPLIT>
__global__ void __launch_bounds__(AT_THREADS, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQh, const __grid_constant__ CUtensorMap tmQl,
                 const __grid_constant__ CUtensorMap tmKh, const __grid_constant__ CUtensorMap tmKl,
                 const __grid_constant__ CUtensorMap tmVh, const __grid_constant__ CUtensorMap tmVl,
                 __nv_bfloat16* __restrict__ ctx_hi, __nv_bfloat16* __restrict__ ctx_lo, int num_items, int qtiles) {
  using C = AttnCfg<NSPLIT>;
  EB_DYN_SMEM_1K(smem);
  if ((smem_u32(smem) & 1023u) != 0) __trap();   // 128-byte-swizzle tiles need a 1 KB aligned base
  uint8_t* sQ = smem;
  uint8_t* sRing = smem + C::Q_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
  uint64_t* q_full = bars;                  // [1]
  uint64_t* q_empty = bars + 1;             // [1]
  uint64_t* kv_full = bars + 2;             // [NSLOTS]
  uint64_t* kv_empty = kv_full + C::NSLOTS;
  uint64_t* s_full = kv_empty + C::NSLOTS;  // [2]  S_j in TMEM
  uint64_t* s_empty = s_full + 2;           // [2]  softmax has read S_j
  uint64_t* p_full = s_empty + 2;           // [2]  P_j in TMEM
  uint64_t* pv_done = p_full + 2;           // [2]  PV_j finished: P buffer free, O up to date
  uint64_t* o_full = pv_done + 2;           // [2]  last PV of an item finished
  uint64_t* o_empty = o_full + 2;           // [2]  epilogue has drained that O buffer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 2);
  float* xmax = reinterpret_cast<float*>(smem + C::XCHG_OFF);     // [2][2][128]
  float* lsum = xmax + 512;                                        // [2][2][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQh); tma_prefetch_desc(&tmKh); tma_prefetch_desc(&tmVh);
    if (NSPLIT > 1) { tma_prefetch_desc(&tmQl); tma_prefetch_desc(&tmKl); tma_prefetch_desc(&tmVl); }
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1); mbar_init(q_empty, 1);
    for (int i = 0; i < C::NSLOTS; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 8);
      mbar_init(&p_full[i], 8); mbar_init(&pv_done[i], 1);
      mbar_init(&o_full[i], 1); mbar_init(&o_empty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<1>(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (*tmem_slot != 0) __trap();            // the whole TMEM was allocated: base is column 0 by construction

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      TR_DECL
      uint32_t rc = 0;  // ring counter: 18 tiles per item, NSLOTS | 18, so slot = step % NSLOTS as in the MMA warp
      int it = 0;
#ifdef EB_ATTN_TRACE
      const long long tr_start_ = clock64();
#endif
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const int qt = item % qtiles, bh = item / qtiles;
        const int h = bh % AT_HEADS, b = bh / AT_HEADS;
        TR_WAIT(1, mbar_wait(q_empty, (it & 1) ^ 1));
        mbar_expect_tx(q_full, C::Q_BYTES);
        for (int kb = 0; kb < 2; ++kb) {
          tma_load_4d(sQ + kb * C::Q_BLK, &tmQh, q_full, kb * 64, qt * AT_QT, h, b);
          if (NSPLIT > 1) tma_load_4d(sQ + (2 + kb) * C::Q_BLK, &tmQl, q_full, kb * 64, qt * AT_QT, h, b);
        }
        // ring order = MMA consumption order: K0, K1, V0, K2, V1, ..., K8, V7, V8
        for (int step = 0; step < 2 * AT_NT; ++step) {
          const bool is_k = (step == 0) || (step < 2 * AT_NT - 1 && (step & 1));
          const int j = (step == 0) ? 0 : (is_k ? (step + 1) / 2 : (step == 2 * AT_NT - 1 ? AT_NT - 1 : step / 2 - 1));
          const uint32_t slot = rc % C::NSLOTS, ph = (rc / C::NSLOTS) & 1;
          TR_WAIT(2, mbar_wait(&kv_empty[slot], ph ^ 1));
          uint8_t* dst = sRing + slot * C::SLOT_BYTES;
          mbar_expect_tx(&kv_full[slot], C::SLOT_BYTES);
          if (is_k) {
            for (int kb = 0; kb < 2; ++kb) {
              tma_load_4d(dst + kb * C::K_BLK, &tmKh, &kv_full[slot], kb * 64, j * AT_KT, AT_HEADS + h, b);
              if (NSPLIT > 1)
                tma_load_4d(dst + (2 + kb) * C::K_BLK, &tmKl, &kv_full[slot], kb * 64, j * AT_KT, AT_HEADS + h, b);
            }
          } else {
            tma_load_4d(dst, &tmVh, &kv_full[slot], j * AT_KT, bh * AT_D, 0, 0);
            if (NSPLIT > 1) tma_load_4d(dst + C::V_BLK, &tmVl, &kv_full[slot], j * AT_KT, bh * AT_D, 0, 0);
          }
          ++rc;
        }
      }
#ifdef EB_ATTN_TRACE
      tr_[0] = clock64() - tr_start_;
#endif
      TR_FLUSH(16, true)
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // Program order = ring order: S0, S1, PV0, S2, PV1, ..., S8, PV7, PV8.  The loop is fully unrolled; tile index,
    // ring slot, S / P buffer are compile-time constants, only the mbarrier parities depend on the item count.
    constexpr uint32_t idesc_s = make_idesc_bf16(AT_QT, AT_KT);
    constexpr uint32_t idesc_o = make_idesc_bf16(AT_QT, AT_D);
    const uint32_t q_lo = sdesc_lo(smem_u32(sQ)), ring_lo = sdesc_lo(smem_u32(sRing));
    int it = 0;
    TR_DECL
#ifdef EB_ATTN_TRACE
    const long long tr_start_ = clock64();
#endif
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const uint32_t t_o = C::T_O + (it & 1) * AT_D;      // O accumulator of this item
      TR_WAIT(1, mbar_wait(q_full, it & 1));
      tc_fence_after();
#pragma unroll
      for (int step = 0; step < 2 * AT_NT; ++step) {
        const bool is_s = (step == 0) || (step < 2 * AT_NT - 1 && (step & 1));
        const int j = (step == 0) ? 0 : (is_s ? (step + 1) / 2 : (step == 2 * AT_NT - 1 ? AT_NT - 1 : step / 2 - 1));
        const int slot = step % C::NSLOTS;
        const uint32_t slot_lo = ring_lo + ((slot * C::SLOT_BYTES) >> 4);
        TR_WAIT(2, mbar_wait(&kv_full[slot], (step / C::NSLOTS + C::USES * it) & 1));
        const int buf = j & 1;
        if (is_s) {
          TR_WAIT(3, mbar_wait(&s_empty[buf], buf_parity(j, it) ^ 1));
          tc_fence_after();
          const uint32_t d = C::T_S + buf * AT_KT;
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t qh = sdesc_at(q_lo, kb * C::Q_BLK + kk * 32);
              const uint64_t kh = sdesc_at(slot_lo, kb * C::K_BLK + kk * 32);
              umma_bf16<1>(d, qh, kh, idesc_s, (kb | kk) != 0 ? 1u : 0u);
              if (NSPLIT > 1) {
                const uint64_t ql = sdesc_at(q_lo, (2 + kb) * C::Q_BLK + kk * 32);
                const uint64_t kl = sdesc_at(slot_lo, (2 + kb) * C::K_BLK + kk * 32);
                umma_bf16<1>(d, qh, kl, idesc_s, 1u);
                umma_bf16<1>(d, ql, kh, idesc_s, 1u);
              }
            }
          umma_commit<1>(&kv_empty[slot]);
          umma_commit<1>(&s_full[buf]);
          if (j == AT_NT - 1) umma_commit<1>(q_empty);
        } else {
          if (j == 0) {   // the epilogue warps have drained this O buffer (used two items ago)
            TR_WAIT(5, mbar_wait(&o_empty[it & 1], ((it >> 1) & 1) ^ 1));
          }
          TR_WAIT(4, mbar_wait(&p_full[buf], buf_parity(j, it)));
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t vh = sdesc_at(slot_lo, kk * 32);
            const uint32_t pa = C::T_P + buf * 64 + kk * 8;      // 16 keys = 8 packed columns per k-step
            umma_bf16_ts(t_o, pa, vh, idesc_o, (j == 0 && kk == 0) ? 0u : 1u);
            if (NSPLIT > 1) {
              const uint64_t vl = sdesc_at(slot_lo, C::V_BLK + kk * 32);
              umma_bf16_ts(t_o, pa, vl, idesc_o, 1u);
              umma_bf16_ts(t_o, pa + 32, vh, idesc_o, 1u);
            }
          }
          umma_commit<1>(&kv_empty[slot]);
          umma_commit<1>(&pv_done[buf]);
          if (j == AT_NT - 1) umma_commit<1>(&o_full[it & 1]);
        }
      }
    }
#ifdef EB_ATTN_TRACE
    tr_[0] = clock64() - tr_start_;
#endif
    TR_FLUSH(0, lane == 0)
  } else if (warp < 10) {
    // ------------------------------------------------------------------ softmax warps (8)
    // Warps w and w+4 own the same TMEM lane quarter (query rows) and split each 64-key tile into two 32-column
    // halves; the pair agrees on the row maximum through shared memory and a 64-thread named barrier.
    const int q = warp & 3;                       // TMEM lane quarter
    const int hf = (warp - 2) >> 2;               // column half handled by this warp
    const int row = q * 32 + lane;                // query row inside the tile
    const uint32_t lane_sel = uint32_t(q * 32) << 16;
    const float c = 0.08838834764831845f * 1.4426950408889634f;   // log2(e) / sqrt(128)
    int it = 0;
    uint32_t gt = 0;                              // global tile counter (exchange-buffer parity)
    TR_DECL
#ifdef EB_ATTN_TRACE
    const long long tr_start_ = clock64();
#endif
    auto pair_sync = [&]() { named_bar_sync<64>(1 + q); };
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const uint32_t t_o = C::T_O + (it & 1) * AT_D;
      float m_ref = 0.f, l = 0.f;
#pragma unroll 1
      for (int j = 0; j < AT_NT; ++j, ++gt) {
        const uint32_t sb = j & 1;
        TR_WAIT(1, mbar_wait(&s_full[sb], buf_parity(j, it)));
        tc_fence_after();
        uint32_t r[32];
        tmem_ld32(C::T_S + lane_sel + sb * AT_KT + hf * 32, r);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[sb]);
        float mt = __uint_as_float(r[0]);
#pragma unroll
        for (int i = 1; i < 32; ++i) mt = fmaxf(mt, __uint_as_float(r[i]));
        float* xm = xmax + (gt & 1) * 256;
        xm[hf * 128 + row] = mt;
        TR_WAIT(2, pair_sync());
        mt = fmaxf(mt, xm[(hf ^ 1) * 128 + row]);
        if (j == 0) m_ref = mt;
        const bool need = (j > 0) && ((mt - m_ref) * c > 8.0f);
        // probabilities first (registers only) ...
        float f = 1.0f;
        if (need) { f = ex2_approx((m_ref - mt) * c); m_ref = mt; l *= f; }
        const float mc = m_ref * c;
        uint32_t hh[16], ll[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(r[2 * e]), c, -mc));
          const float p1 = ex2_approx(fmaf(__uint_as_float(r[2 * e + 1]), c, -mc));
          l += p0 + p1;
          if (NSPLIT > 1) split_pack2(p0, p1, hh[e], ll[e]); else hh[e] = cvt_bf16x2(p0, p1);
        }
        // ... then wait until PV_{j-2} has completed (this P buffer is free); only a rescale of O additionally
        // needs PV_{j-1}
        const uint32_t pb = j & 1;
        TR_WAIT(3, mbar_wait(&pv_done[pb], buf_parity(j, it) ^ 1));
        if (__any_sync(0xffffffffu, need)) {
          mbar_wait(&pv_done[pb ^ 1], buf_parity(j - 1, it));
          tc_fence_after();
#pragma unroll 1
          for (int ch = 0; ch < 2; ++ch) {
            uint32_t o[32];
            tmem_ld32(t_o + lane_sel + hf * 64 + ch * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            tmem_st32(t_o + lane_sel + hf * 64 + ch * 32, o);
          }
        }
        // this thread's 32 keys = 16 packed columns of its row (lane) of the TMEM-resident A operand of P.V
        tmem_st16(C::T_P + lane_sel + pb * 64 + hf * 16, hh);
        if (NSPLIT > 1) tmem_st16(C::T_P + lane_sel + pb * 64 + 32 + hf * 16, ll);
        tmem_st_wait();
        if (j == AT_NT - 1) lsum[(it & 1) * 256 + hf * 128 + row] = l;   // for the epilogue warps, ordered by p_full
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[pb]);
      }
    }
#ifdef EB_ATTN_TRACE
    tr_[0] = clock64() - tr_start_;
#endif
    TR_FLUSH(8, warp == 2 && lane == 0)
  } else {
    // ------------------------------------------------------------------ epilogue warps (4): O / l -> ctx rows
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_sel = uint32_t(q * 32) << 16;
    int it = 0;
    TR_DECL
#ifdef EB_ATTN_TRACE
    const long long tr_start_ = clock64();
#endif
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const int qt = item % qtiles, bh = item / qtiles;
      const int h = bh % AT_HEADS, b = bh / AT_HEADS;
      const int ob = it & 1;
      TR_WAIT(1, mbar_wait(&o_full[ob], (it >> 1) & 1));
      tc_fence_after();
      const float inv = 1.0f / (lsum[ob * 256 + row] + lsum[ob * 256 + 128 + row]);
      const int tok = qt * AT_QT + row;
      const long long orow = (long long)b * AT_TOK + tok;
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t o[32];
        tmem_ld32(C::T_O + ob * AT_D + lane_sel + ch * 32, o);
        tmem_ld_wait();
        if (tok < AT_TOK) {
          uint32_t hh[16], ll[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float v0 = __uint_as_float(o[2 * e]) * inv, v1 = __uint_as_float(o[2 * e + 1]) * inv;
            if (NSPLIT > 1) split_pack2(v0, v1, hh[e], ll[e]); else hh[e] = cvt_bf16x2(v0, v1);
          }
          const long long off = orow * (AT_HEADS * AT_D) + h * AT_D + ch * 32;
          uint4* oh = reinterpret_cast<uint4*>(ctx_hi + off);
#pragma unroll
          for (int e = 0; e < 4; ++e) oh[e] = make_uint4(hh[4 * e], hh[4 * e + 1], hh[4 * e + 2], hh[4 * e + 3]);
          if (NSPLIT > 1) {
            uint4* ol = reinterpret_cast<uint4*>(ctx_lo + off);
#pragma unroll
            for (int e = 0; e < 4; ++e) ol[e] = make_uint4(ll[4 * e], ll[4 * e + 1], ll[4 * e + 2], ll[4 * e + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[ob]);
    }
#ifdef EB_ATTN_TRACE
    tr_[0] = clock64() - tr_start_;
#endif
    TR_FLUSH(24, warp == 10 && lane == 0)
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(0u, C::TMEM_COLS);
}

template <int NSPLIT>
static int launch_attention(const CUtensorMap* tm, __nv_bfloat16* ctx_hi, __nv_bfloat16* ctx_lo, int B, int qtiles,
                            cudaStream_t stream) {
  using C = AttnCfg<NSPLIT>;
  auto kern = attention_kernel<NSPLIT>;
  // the opt-in to > 48 KB dynamic shared memory is a per-device function attribute
  static bool attr_done[64] = {false};
  const int dev_ = current_device();
  if (!attr_done[dev_]) {
    EB_CUDA(EB_SET_MAX_SMEM(kern, C::SMEM_BYTES));
    attr_done[dev_] = true;
  }
  const int items = B * AT_HEADS * qtiles;
  const int grid = items < num_sms() ? items : num_sms();
  ProfScope prof("attention_kernel", stream);
  EB_LAUNCH_SMEM(kern, grid, AT_THREADS, C::SMEM_BYTES, stream, tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], ctx_hi, ctx_lo, items, qtiles);
  EB_CHECK_LAUNCH("attention_kernel");
  return 0;
}

// qk: (B*576, 2048) = [Q | K] per token, head h at columns h*128; vt: (B*8*128, 576) = V^T per (frame, head)
// query_rows: only the first query_rows tokens of every frame get a context row (all 576 are keys / values)
int attention_run(const __nv_bfloat16* qk_hi, const __nv_bfloat16* qk_lo, const __nv_bfloat16* vt_hi,
                  const __nv_bfloat16* vt_lo, __nv_bfloat16* ctx_hi, __nv_bfloat16* ctx_lo, int B, int nsplit,
                  int query_rows, cudaStream_t stream) {
  // EGOTAP_ATTN=wide (opt-in, A/B): the 128-key-tile structure of attention_wide.cu; read per call so that one process
  // can compare the two
  if (const char* e = getenv("EGOTAP_ATTN"))
    if (strcmp(e, "wide") == 0)
      return attention_wide_run(qk_hi, qk_lo, vt_hi, vt_lo, ctx_hi, ctx_lo, B, nsplit, query_rows, stream);
  EB_REQUIRE(query_rows > 0 && query_rows <= AT_TOK, "attention: query_rows must be in (0, 576]");
  const int qtiles = (query_rows + AT_QT - 1) / AT_QT;
  EB_REQUIRE(qk_hi && vt_hi && ctx_hi && B > 0, "attention: bad arguments");
  EB_REQUIRE(nsplit == 1 || (qk_lo && vt_lo && ctx_lo), "attention: bf16x3 mode needs the lo parts");
  CUtensorMap tm[6];
  int rc;
  const long long fs = (long long)AT_TOK * 2 * AT_HEADS * AT_D;  // frame stride of the qk buffer
  for (int part = 0; part < (nsplit == 3 ? 2 : 1); ++part) {
    const __nv_bfloat16* qk = part == 0 ? qk_hi : qk_lo;
    const __nv_bfloat16* vt = part == 0 ? vt_hi : vt_lo;
    if ((rc = make_operand_tmap(&tm[0 + part], qk, AT_D, AT_TOK, 2 * AT_HEADS * AT_D, 2 * AT_HEADS, AT_D, B, fs, AT_QT)))
      return rc;
    if ((rc = make_operand_tmap(&tm[2 + part], qk, AT_D, AT_TOK, 2 * AT_HEADS * AT_D, 2 * AT_HEADS, AT_D, B, fs, AT_KT)))
      return rc;
    if ((rc = make_operand_tmap(&tm[4 + part], vt, AT_TOK, (long long)B * AT_HEADS * AT_D, AT_TOK, 1, 0, 1, 0, AT_D)))
      return rc;
  }
  if (nsplit == 1) { tm[1] = tm[0]; tm[3] = tm[2]; tm[5] = tm[4]; }
  return nsplit == 3 ? launch_attention<3>(tm, ctx_hi, ctx_lo, B, qtiles, stream)
                     : launch_attention<1>(tm, ctx_hi, ctx_lo, B, qtiles, stream);
}

}  // namespace eb

#ifdef EB_ATTN_TRACE
extern "C" int egotap_b200_attn_trace(unsigned long long* out32, int reset) {
  if (out32) cudaMemcpyFromSymbol(out32, eb::g_attn_trace, sizeof(unsigned long long) * 32);
  if (reset) { unsigned long long z[32] = {0}; cudaMemcpyToSymbol(eb::g_attn_trace, z, sizeof(z)); }
  return 0;
}
#endif

extern "C" int egotap_b200_attention(const void* qk_hi, const void* qk_lo, const void* vt_hi, const void* vt_lo,
                                     void* ctx_hi, void* ctx_lo, int frames, int precision, void* stream) {
  return eb::attention_run((const __nv_bfloat16*)qk_hi, (const __nv_bfloat16*)qk_lo, (const __nv_bfloat16*)vt_hi,
                           (const __nv_bfloat16*)vt_lo, (__nv_bfloat16*)ctx_hi, (__nv_bfloat16*)ctx_lo, frames,
                           precision == EGOTAP_PREC_BF16 ? 1 : 3, eb::AT_TOK, (cudaStream_t)stream);
}

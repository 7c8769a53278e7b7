// Fused multi-head self-attention, sm_100a.
//   ctx[b, q, h*128 + :] = softmax(Q K^T / sqrt(128)) V        reference model/modeling_vit.py:233-252
//
// Round 1's first fused kernel used 64-key score tiles: every score MMA was 128 x 64 x 16 = 32 tensor-pipe cycles, below the
// ~55 cycles the issuing warp needs per tcgen05.mma (measured), with ~40 mbarrier round trips per work item; it ran at 50.6 %
// tensor-pipe activity (1.058 ms per layer at batch 256, parity mode) and was deleted in round 2 when this structure measured
// 65.7 % / 0.994 ms (0.484 vs 0.523 ms in bf16 mode; profiles/r02_attention_ab.md).  Here every MMA is 128 x 128 x 16
// (64 cycles of math) and the hand-offs are cut to ~26 per item.
//
// Measured limits of this structure (round 2, profiles/r02c_attention_analysis.md): per 128-key tile the tensor pipe needs
// 3,072 cycles in the parity mode (1,024 in bf16 mode) and the period is T_PV + max(T_S, T_softmax) with T_softmax ~ 3,100
// cycles for the 8 softmax warps -- ~50 cycles per score element and thread, set by tcgen05.ld / tcgen05.st round trips
// that share the tensor-memory ports with the running MMAs, not by MUFU (XU pipe 20 % busy) or issue slots (29 %).  A variant
// with two softmax groups alternating over the tiles (one query row per thread, no intra-tile exchange, epilogue folded into
// the groups) was built, verified and timed: 1.016 / 0.503 ms vs 0.994 / 0.484 ms here (parity / bf16, batch 256) -- the
// per-element cost did not change, so the longer per-tile chain of a 4-warp group cancelled the overlap.  It was dropped.
//   * 128-key score tiles (576 keys = 4 full tiles + one 64-key tile, issued with N = 64);
//   * P_g is written over the columns of S_g (same tensor-memory buffer: S 128 fp32 columns -> P 64 hi + 64 lo packed
//     bf16x2 columns), so TMEM holds S/P 2 x 128 + O 2 x 128 = 512 columns;
//   * no "S buffer free" / "P buffer free" barriers: S_{g+2} is issued after PV_g in program order and tcgen05.mma
//     instructions of one CTA execute in issue order, so S_{g+2} cannot overwrite P_g before PV_g has read it
//     (a hardware property this kernel relies on, confirmed on the B200; the softmax warps never touch buffer
//     g & 1 between their p_full arrive for tile g and the s_full of tile g + 2);
//   * one software pipeline over the GLOBAL tile sequence of a persistent CTA (5 tiles per work item, buffers alternate
//     across item boundaries):  S_0; for g: { S_{g+1}; PV_g }  -- the score MMAs of the next item's first tile run
//     while the softmax warps finish the current item.
// Shared memory: Q tile (hi/lo; two of them in bf16 mode) + a ring of 32 KB granules (one 64-wide d block of a 128-key K tile, or one 64-key
// block of a V^T tile), consumed in MMA order.
//   warp 0   TMA producer      warps 1, 2   MMA issuers (even / odd key tiles)      warp 3   tensor-memory allocation
//   warps 4-11  softmax (pairs split the keys of a tile)
//   warps 12-15  epilogue: O (double-buffered across items) / l -> bf16 hi/lo context rows
// Two issuing warps (round 2; ncu source page and cycle counters of the single-issuer kernel, profiles/r02g_attention_two_issuers.md):
// every tcgen05.mma costs its issuing warp ~80 cycles of election / descriptor / R2UR instructions against 64 cycles of tensor-pipe
// work per 128 x 128 x 16 MMA; with the 48 MMAs of a parity-mode tile on ONE warp that warp was busy ~3,800 of the ~4,300 cycles of
// a tile period and stalled on the MMA queue only 12 % of its time -- instruction issue, not the tensor pipe or the softmax, set
// the period.  Warp 1 now issues S_g and P V_g of the even tiles, warp 2 of the odd tiles.  Each S/P buffer is touched by one
// thread's MMAs only, so S_{g+2} still follows P V_g in that thread's program order (no buffer-free barrier); the accumulation
// order of O across the two threads is enforced with the existing pv_done barriers (P V_{g+1} is issued after P V_g has completed,
// which it has long before P_{g+1} arrives).  Measured: bf16 mode 0.460 -> 0.424 ms per layer; the parity mode stays at 0.944 ms,
// now bound by the latency of the L2 -> shared-memory stream of K / V against the 4-slot ring (at most 128 KB in flight per SM).
#include "gemm.cuh"
#include "host_util.cuh"
#include "internal.h"

namespace eb {

// A/B builds (tools/gpu_job_r2*.sh): -DEB_ATTN_NO_REGSPLIT keeps the launch-time register allocation and the compiler's own
// handling of the role constants
#ifdef EB_ATTN_NO_REGSPLIT
template <int N> __device__ __forceinline__ void aw_reg_dec() {}
template <int N> __device__ __forceinline__ void aw_reg_inc() {}
__device__ __forceinline__ int aw_pin(int v) { return v; }
#else
template <int N> __device__ __forceinline__ void aw_reg_dec() { warpgroup_reg_dec<N>(); }
template <int N> __device__ __forceinline__ void aw_reg_inc() { warpgroup_reg_inc<N>(); }
__device__ __forceinline__ int aw_pin(int v) { return pin_reg(v); }
#endif

// Debug-only wait-cycle accounting (tools/attn_trace.py builds a separate library with -DEB_ATTN_TRACE); no-ops otherwise.
#ifdef EB_ATTN_TRACE
__device__ unsigned long long g_attn_trace[32];
#define TW_DECL long long tw_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long tw_t0_ = 0; (void)tw_t0_; const long long tw_start_ = clock64();
#define TW_WAIT(i, stmt) { tw_t0_ = clock64(); stmt; tw_[i] += clock64() - tw_t0_; }
#define TW_FLUSH(base, cond) { tw_[0] = clock64() - tw_start_; if (cond) { for (int i_ = 0; i_ < 8; ++i_) atomicAdd(&g_attn_trace[(base) + i_], (unsigned long long)tw_[i_]); } }
#else
#define TW_DECL
#define TW_WAIT(i, stmt) { stmt; }
#define TW_FLUSH(base, cond)
#endif

constexpr int AW_TOK = 576, AW_HEADS = 8, AW_D = 128, AW_QT = 128, AW_KT = 128;
constexpr int AW_NT = (AW_TOK + AW_KT - 1) / AW_KT;          // 5 key tiles, the last one has 64 keys
constexpr int AW_LAST_KEYS = AW_TOK - (AW_NT - 1) * AW_KT;   // 64
constexpr int AW_THREADS = 128 + 256 + 128;                  // TMA / 2 MMA / allocation warps, 8 softmax warps, 4 epilogue warps
static_assert(AW_LAST_KEYS == 64, "the last key tile is issued with N = 64");

template <int NSPLIT>
struct AttnWideCfg {
  static constexpr int NOPS = NSPLIT == 1 ? 1 : 2;
  static constexpr int BLK = 128 * 64 * 2;                   // 128 rows x 64 bf16 (128-byte rows): 16 KB
  static constexpr int Q_BYTES = NOPS * 2 * BLK;             // [hi d0-63][hi d64-127][lo d0-63][lo d64-127]
  static constexpr int GRAN_BYTES = NOPS * BLK;              // K granule: [hi][lo] of one d block; V^T granule: [hi][lo] of one key block
  static constexpr int NSLOTS = NSPLIT == 1 ? 8 : 4;
  // bf16 mode has room for two Q tiles: the next item's Q is loaded a whole item ahead.  In the parity mode the single Q
  // buffer is reloaded when the item's last score tile has been computed; the tile is prefetched into L2 an item ahead.
  static constexpr int QBUF = NSPLIT == 1 ? 2 : 1;
  static constexpr int BAR_OFF = QBUF * Q_BYTES + NSLOTS * GRAN_BYTES;
  static constexpr int XCHG_OFF = BAR_OFF + 512;             // [2][2][128] row maxima, then [2][2][128] row sums
  static constexpr int SMEM_BYTES = XCHG_OFF + 5120;         // base must be 1 KB aligned (checked); + [2][128] final row maxima
  static constexpr int TMEM_COLS = 512;
  static constexpr uint32_t T_S = 0, T_O = 256;              // S/P buffer b at T_S + 128 b; O buffer ob at T_O + 128 ob
  static constexpr uint32_t P_LO = 64;                       // packed lo columns of P inside its S/P buffer
  // registers per thread after the role split (launched with 128): TMA / MMA / allocation warpgroup, the two softmax warpgroups
  // (64 score registers + 32 packed probabilities live at once: 128 spilled and re-derived its role constants), epilogue
  static constexpr int REG_ISSUE = 96, REG_SOFTMAX = 152, REG_EPILOGUE = 112;
  static_assert(128 * REG_ISSUE + 256 * REG_SOFTMAX + 128 * REG_EPILOGUE <= 65536, "register file");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
  static_assert(2 * QBUF + 3 * NSLOTS + 12 + 1 <= 64, "barrier block");
};

template <int NSPLIT>
__global__ void __launch_bounds__(AW_THREADS, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQh, const __grid_constant__ CUtensorMap tmQl,
                      const __grid_constant__ CUtensorMap tmVh, const __grid_constant__ CUtensorMap tmVl,
                      __nv_bfloat16* __restrict__ ctx_hi, __nv_bfloat16* __restrict__ ctx_lo, float* __restrict__ lse,
                      int num_items, int qtiles) {
  using C = AttnWideCfg<NSPLIT>;
  EB_DYN_SMEM_1K(smem);
  if ((smem_u32(smem) & 1023u) != 0) __trap();   // 128-byte-swizzle tiles need a 1 KB aligned base
  uint8_t* sQ = smem;
  uint8_t* sRing = smem + C::QBUF * C::Q_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
  uint64_t* q_full = bars;                  // [QBUF]
  uint64_t* q_empty = bars + C::QBUF;       // [QBUF]  last score tile of an item has been computed
  uint64_t* kv_full = bars + 2 * C::QBUF;   // [2][NSLOTS]: one set per issuing warp (below)
  uint64_t* kv_empty = kv_full + 2 * C::NSLOTS;
  uint64_t* s_full = kv_empty + C::NSLOTS;  // [2]  S_g in TMEM
  uint64_t* p_full = s_full + 2;            // [2]  P_g in TMEM (over S_g)
  uint64_t* pv_done = p_full + 2;           // [2]  PV_g finished: O up to date (needed only for a rescale of O)
  uint64_t* o_full = pv_done + 2;           // [2]  last PV of an item finished
  uint64_t* o_empty = o_full + 2;           // [2]  epilogue has drained that O buffer
  uint64_t* p_half = o_empty + 2;           // [2]  first 32 keys of every softmax warp's half of P_g in TMEM (full tiles)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_half + 2);
  float* xmax = reinterpret_cast<float*>(smem + C::XCHG_OFF);     // [2][2][128]
  float* lsum = xmax + 512;                                        // [2][2][128]
  float* mfin = lsum + 512;                                        // [2][128] reference maximum at the end of an item (for lse)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int my_items = (num_items - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);   // grid <= num_items
  const int my_tiles = my_items * AW_NT;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQh); tma_prefetch_desc(&tmVh);
    if (NSPLIT > 1) { tma_prefetch_desc(&tmQl); tma_prefetch_desc(&tmVl); }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::QBUF; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 2); }   // q_empty: one commit per issuing warp
    for (int i = 0; i < C::NSLOTS; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_full[C::NSLOTS + i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 8); mbar_init(&pv_done[i], 1);
      mbar_init(&o_full[i], 1); mbar_init(&o_empty[i], 4); mbar_init(&p_half[i], 8);
    }
    fence_mbar_init();
  }
  if (warp == 3) tmem_alloc<1>(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (*tmem_slot != 0) __trap();            // the whole TMEM was allocated: base is column 0 by construction

  // (each warpgroup re-sizes its registers at the top of its own branch: ptxas budgets the code that follows a setmaxnreg)
  if (warp < 4) {
  aw_reg_dec<C::REG_ISSUE>();
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // ring order = MMA consumption order: K(0); then per global tile g: K(g+1), V(g).  Q is outside the ring.
    if (lane == 0) {
      TW_DECL
      uint32_t rc = 0;
      auto item_coords = [&](int it, int& qt, int& h, int& b, int& bh) {
        const int item = int(blockIdx.x) + it * int(gridDim.x);
        qt = item % qtiles; bh = item / qtiles; h = bh % AW_HEADS; b = bh / AW_HEADS;
      };
      auto load_q = [&](int it) {
        int qt, h, b, bh;
        item_coords(it, qt, h, b, bh);
        const int qb = it % C::QBUF;
        uint8_t* dst = sQ + qb * C::Q_BYTES;
        TW_WAIT(1, mbar_wait(&q_empty[qb], ((it / C::QBUF) & 1) ^ 1));
        mbar_expect_tx(&q_full[qb], C::Q_BYTES);
        for (int kb = 0; kb < 2; ++kb) {
          tma_load_4d(dst + kb * C::BLK, &tmQh, &q_full[qb], kb * 64, qt * AW_QT, h, b);
          if (NSPLIT > 1) tma_load_4d(dst + (2 + kb) * C::BLK, &tmQl, &q_full[qb], kb * 64, qt * AW_QT, h, b);
        }
      };
      auto prefetch_q = [&](int it) {      // L2 prefetch of a Q tile that will be loaded when its buffer is free
        int qt, h, b, bh;
        item_coords(it, qt, h, b, bh);
        for (int kb = 0; kb < 2; ++kb) {
          tma_prefetch_4d(&tmQh, kb * 64, qt * AW_QT, h, b);
          if (NSPLIT > 1) tma_prefetch_4d(&tmQl, kb * 64, qt * AW_QT, h, b);
        }
      };
      // A granule's arrival is signalled on the barrier set of the warp that consumes it (tile parity).  With ONE set, a warp that
      // consumes only every other filling of a slot would wait for "its" phase while the previous filling (the other warp's) is
      // still in flight -- and mbarrier parity waits succeed at once when the barrier is a phase behind: seen on the B200 in the
      // parity mode (4 slots, loads often late) as non-deterministic results and launch failures.  Per warp, every phase of its
      // barriers is waited for in order.
      auto slot_acquire = [&](uint64_t*& full, int g) -> uint8_t* {
        const uint32_t slot = rc % C::NSLOTS, ph = (rc / C::NSLOTS) & 1;
        ++rc;
        TW_WAIT(2, mbar_wait(&kv_empty[slot], ph ^ 1));
        full = &kv_full[(g & 1) * C::NSLOTS + slot];
        mbar_expect_tx(full, C::GRAN_BYTES);
        return sRing + slot * C::GRAN_BYTES;
      };
      auto load_k = [&](int g) {       // two granules: d blocks 0 and 1 of the 128-key tile (keys past 576 are zero-filled)
        int qt, h, b, bh;
        item_coords(g / AW_NT, qt, h, b, bh);
        const int j = g % AW_NT;
        for (int kb = 0; kb < 2; ++kb) {
          uint64_t* full;
          uint8_t* dst = slot_acquire(full, g);
          tma_load_4d(dst, &tmQh, full, kb * 64, j * AW_KT, AW_HEADS + h, b);
          if (NSPLIT > 1) tma_load_4d(dst + C::BLK, &tmQl, full, kb * 64, j * AW_KT, AW_HEADS + h, b);
        }
      };
      auto load_v = [&](int g) {       // one granule per 64 keys of the tile
        int qt, h, b, bh;
        item_coords(g / AW_NT, qt, h, b, bh);
        const int j = g % AW_NT;
        const int ngran = j == AW_NT - 1 ? AW_LAST_KEYS / 64 : AW_KT / 64;
        for (int kg = 0; kg < ngran; ++kg) {
          uint64_t* full;
          uint8_t* dst = slot_acquire(full, g);
          tma_load_4d(dst, &tmVh, full, j * AW_KT + kg * 64, bh * AW_D, 0, 0);
          if (NSPLIT > 1) tma_load_4d(dst + C::BLK, &tmVl, full, j * AW_KT + kg * 64, bh * AW_D, 0, 0);
        }
      };
      if (my_tiles > 0) { load_q(0); load_k(0); }
      for (int g = 0; g < my_tiles; ++g) {
        const int it = g / AW_NT;
        if (g % AW_NT == 0 && it + 1 < my_items) {
          if (C::QBUF == 2) load_q(it + 1); else prefetch_q(it + 1);
        }
        if (g + 1 < my_tiles) {
          if (C::QBUF == 1 && (g + 1) % AW_NT == 0) load_q((g + 1) / AW_NT);
          load_k(g + 1);
        }
        load_v(g);
      }
      TW_FLUSH(16, true)
    }
  } else if (warp == 1 || warp == 2) {
    // ------------------------------------------------------------------ MMA issuers (all lanes run the loop; the wrappers
    // elect one lane).  Warp 1 owns the even global tiles, warp 2 the odd ones; program order of each: S_p; { PV_g; S_{g+2} }.
    const int par = warp - 1;
    constexpr uint32_t idesc_s = make_idesc_bf16(AW_QT, AW_KT);
    constexpr uint32_t idesc_s_last = make_idesc_bf16(AW_QT, AW_LAST_KEYS);
    constexpr uint32_t idesc_o = make_idesc_bf16(AW_QT, AW_D);
    const uint32_t q_lo0 = sdesc_lo(smem_u32(sQ)), ring_lo = sdesc_lo(smem_u32(sRing));
    TW_DECL
    // position of a tile's granules in the ring sequence K(0); { K(g+1); V(g) }: a full tile has two V^T granules, the 64-key
    // tail tile of every item one
    auto seq_k = [&](int g) -> uint32_t { return g == 0 ? 0u : uint32_t(2 + 4 * (g - 1) - (g - 1) / AW_NT); };
    auto seq_v = [&](int g) -> uint32_t { return uint32_t(2 + 4 * g - g / AW_NT + (g + 1 < my_tiles ? 2 : 0)); };
    uint32_t full_ph = 0;       // bit s: parity of this warp's next wait on its kv_full barrier of slot s
    auto slot_wait = [&](uint32_t seq, uint32_t& slot) -> uint32_t {
      slot = seq % C::NSLOTS;
      TW_WAIT(2, mbar_wait(&kv_full[par * C::NSLOTS + slot], (full_ph >> slot) & 1u));
      full_ph ^= 1u << slot;
      return ring_lo + ((slot * C::GRAN_BYTES) >> 4);
    };
    auto issue_s = [&](int g) {
      const int it = g / AW_NT, j = g % AW_NT;
      const uint32_t d = C::T_S + uint32_t(g & 1) * 128u;
      const uint32_t idesc = j == AW_NT - 1 ? idesc_s_last : idesc_s;
      const int qb = it % C::QBUF;
      const uint32_t q_lo = q_lo0 + ((qb * C::Q_BYTES) >> 4);
      // both warps read the item's Q tile (the buffer is not reloaded before both have committed q_empty, so the phase
      // waited for is the current or the previous one of the barrier: no aliasing)
      TW_WAIT(1, mbar_wait(&q_full[qb], (it / C::QBUF) & 1));
      const uint32_t seq = seq_k(g);
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
        uint32_t slot;
        const uint32_t g_lo = slot_wait(seq + kb, slot);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t qh = sdesc_at(q_lo, kb * C::BLK + kk * 32);
          const uint64_t kh = sdesc_at(g_lo, kk * 32);
          umma_bf16<1>(d, qh, kh, idesc, (kb | kk) != 0 ? 1u : 0u);
          if (NSPLIT > 1) {
            const uint64_t ql = sdesc_at(q_lo, (2 + kb) * C::BLK + kk * 32);
            const uint64_t kl = sdesc_at(g_lo, C::BLK + kk * 32);
            umma_bf16<1>(d, qh, kl, idesc, 1u);
            umma_bf16<1>(d, ql, kh, idesc, 1u);
          }
        }
        umma_commit<1>(&kv_empty[slot]);
      }
      umma_commit<1>(&s_full[g & 1]);
      // a commit covers the MMAs of its own thread: each warp releases Q after its last score tile of the item (j = 3 and 4)
      if (j >= AW_NT - 2) umma_commit<1>(&q_empty[qb]);
    };
    auto issue_pv = [&](int g) {
      const int it = g / AW_NT, j = g % AW_NT;
      const uint32_t t_o = C::T_O + uint32_t(it & 1) * AW_D;
      const uint32_t t_p = C::T_S + uint32_t(g & 1) * 128u;
      if (j == 0) TW_WAIT(5, mbar_wait(&o_empty[it & 1], ((it >> 1) & 1) ^ 1));   // the epilogue drained this O buffer (two items ago)
      // P_g is handed over in two steps (full tiles): each softmax warp publishes the first 32 of its 64 keys (p_half), then the
      // rest (p_full), so the k-steps over keys {0-31, 64-95} run while the exponentials of {32-63, 96-127} are still being
      // computed -- the P V MMAs of a tile no longer wait for the whole softmax of that tile.
      auto pv_steps = [&](uint32_t g_lo, int kg, int kk0, int kk1) {
#pragma unroll
        for (int kk = kk0; kk < kk1; ++kk) {
          const uint64_t vh = sdesc_at(g_lo, kk * 32);
          const uint32_t pa = t_p + uint32_t(kg * 32 + kk * 8);      // 16 keys = 8 packed columns per k-step
          umma_bf16_ts(t_o, pa, vh, idesc_o, (j == 0 && kg == 0 && kk == 0) ? 0u : 1u);
          if (NSPLIT > 1) {
            const uint64_t vl = sdesc_at(g_lo, C::BLK + kk * 32);
            umma_bf16_ts(t_o, pa, vl, idesc_o, 1u);
            umma_bf16_ts(t_o, pa + C::P_LO, vh, idesc_o, 1u);
          }
        }
      };
      const uint32_t seq = seq_v(g);
      // O is accumulated by the MMAs of two threads: P V_{g-1} (the other warp's) must have completed.  The barrier cannot run
      // ahead: its next phase is P V_{g+1}, which waits for this tile's.
      auto order_after_previous_pv = [&]() {
        if (g > 0) { TW_WAIT(3, mbar_wait(&pv_done[(g - 1) & 1], uint32_t((g - 1) >> 1) & 1u)); }
      };
      if (j != AW_NT - 1) {
        uint32_t slot0, slot1;
        const uint32_t g0_lo = slot_wait(seq, slot0);
        const uint32_t g1_lo = slot_wait(seq + 1, slot1);
        order_after_previous_pv();
#ifndef EB_ATTN_NO_PHALF
        TW_WAIT(4, mbar_wait(&p_half[g & 1], (g >> 1) & 1));
#else
        TW_WAIT(4, mbar_wait(&p_full[g & 1], (g >> 1) & 1));
#endif
        tc_fence_after();
        pv_steps(g0_lo, 0, 0, 2);
        pv_steps(g1_lo, 1, 0, 2);
        TW_WAIT(4, mbar_wait(&p_full[g & 1], (g >> 1) & 1));
        tc_fence_after();
        pv_steps(g0_lo, 0, 2, 4);
        pv_steps(g1_lo, 1, 2, 4);
        umma_commit<1>(&kv_empty[slot0]);
        umma_commit<1>(&kv_empty[slot1]);
      } else {                                  // 64-key tail tile: one V^T granule, every warp's 32 keys arrive at once
        uint32_t slot;
        const uint32_t g_lo = slot_wait(seq, slot);
        order_after_previous_pv();
        TW_WAIT(4, mbar_wait(&p_full[g & 1], (g >> 1) & 1));
        tc_fence_after();
        pv_steps(g_lo, 0, 0, 4);
        umma_commit<1>(&kv_empty[slot]);
      }
      umma_commit<1>(&pv_done[g & 1]);
      if (j == AW_NT - 1) umma_commit<1>(&o_full[it & 1]);
    };
    if (par < my_tiles) issue_s(par);
    for (int g = par; g < my_tiles; g += 2) {
      issue_pv(g);
      if (g + 2 < my_tiles) issue_s(g + 2);
    }
    TW_FLUSH(0, lane == 0 && warp == 1)
  }   // warp 3 (allocation warp) has nothing to do between allocation and release
  } else if (warp < 12) {
    // ------------------------------------------------------------------ softmax warps (8)
    aw_reg_inc<C::REG_SOFTMAX>();
    // Warps w and w+4 own the same TMEM lane quarter (query rows) and split the keys of a tile into two halves; the
    // pair agrees on the row maximum through shared memory and a 64-thread named barrier -- which also separates the
    // pair's loads of S from its stores of P into the same columns.
    const int q = aw_pin(warp & 3);              // TMEM lane quarter
    const int hf = aw_pin((warp - 4) >> 2);      // key half handled by this warp
    const int row = aw_pin(q * 32 + lane);       // query row inside the tile
    const bool lane0 = aw_pin(lane == 0 ? 1 : 0) != 0;
    const uint32_t lane_sel = uint32_t(q * 32) << 16;
    const float c = 0.08838834764831845f * 1.4426950408889634f;   // log2(e) / sqrt(128)
    float m_ref = 0.f, l = 0.f;
    const int n_tiles = aw_pin(my_tiles);
    TW_DECL
#pragma unroll 1
    for (int g = 0; g < n_tiles; ++g) {
      const int it = g / AW_NT, j = g % AW_NT;
      const uint32_t sb = uint32_t(g & 1), par = uint32_t(g >> 1) & 1u;
      const uint32_t t_s = C::T_S + sb * 128u + lane_sel;
      const uint32_t t_o = C::T_O + uint32_t(it & 1) * AW_D + lane_sel;
      const bool full_tile = j != AW_NT - 1;      // this warp: 64 keys of a full tile, 32 of the last one
      const uint32_t col0 = full_tile ? hf * 64 : hf * 32;
      if (j == 0) l = 0.f;
      TW_WAIT(1, mbar_wait(&s_full[sb], par));
      tc_fence_after();
      uint32_t r0[32], r1[32];
      tmem_ld32(t_s + col0, r0);
      if (full_tile) tmem_ld32(t_s + col0 + 32, r1);
      tmem_ld_wait();
      float mt = __uint_as_float(r0[0]);
#pragma unroll
      for (int i = 1; i < 32; ++i) mt = fmaxf(mt, __uint_as_float(r0[i]));
      if (full_tile) {
#pragma unroll
        for (int i = 0; i < 32; ++i) mt = fmaxf(mt, __uint_as_float(r1[i]));
      }
      float* xm = xmax + (g & 1) * 256;
      xm[hf * 128 + row] = mt;
      tc_fence_before();
      TW_WAIT(2, named_bar_sync<64>(1 + q));
      tc_fence_after();
      mt = fmaxf(mt, xm[(hf ^ 1) * 128 + row]);
      if (j == 0) m_ref = mt;
      const bool need = (j > 0) && ((mt - m_ref) * c > 8.0f);
      float f = 1.0f;
      if (need) { f = ex2_approx((m_ref - mt) * c); m_ref = mt; l *= f; }
      const float mc = m_ref * c;
      if (__any_sync(0xffffffffu, need)) {        // lazy rescale of this warp's 64 O columns: needs PV_{g-1} complete
        TW_WAIT(3, mbar_wait(&pv_done[sb ^ 1], uint32_t((g - 1) >> 1) & 1u));
        tc_fence_after();
#pragma unroll 1
        for (int ch = 0; ch < 2; ++ch) {
          uint32_t o[32];
          tmem_ld32(t_o + hf * 64 + ch * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
          tmem_st32(t_o + hf * 64 + ch * 32, o);
        }
      }
      // probabilities: 32 keys = 16 packed columns per chunk, written over S (hi at [0,64), lo at [64,128))
      const uint32_t pcol0 = full_tile ? hf * 32 : hf * 16;
      {
        uint32_t hh[16], ll[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(r0[2 * e]), c, -mc));
          const float p1 = ex2_approx(fmaf(__uint_as_float(r0[2 * e + 1]), c, -mc));
          l += p0 + p1;
          if (NSPLIT > 1) split_pack2(p0, p1, hh[e], ll[e]); else hh[e] = cvt_bf16x2(p0, p1);
        }
        tmem_st16(t_s + pcol0, hh);
        if (NSPLIT > 1) tmem_st16(t_s + C::P_LO + pcol0, ll);
      }
      // first 32 keys of this warp are in tensor memory: the issuing warp may start the P V k-steps that read them (every tile
      // arrives here, so the barrier's phase follows the tile index also across the 64-key tail tiles)
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane0) mbar_arrive(&p_half[sb]);
      if (full_tile) {
        uint32_t hh[16], ll[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(r1[2 * e]), c, -mc));
          const float p1 = ex2_approx(fmaf(__uint_as_float(r1[2 * e + 1]), c, -mc));
          l += p0 + p1;
          if (NSPLIT > 1) split_pack2(p0, p1, hh[e], ll[e]); else hh[e] = cvt_bf16x2(p0, p1);
        }
        tmem_st16(t_s + pcol0 + 16, hh);
        if (NSPLIT > 1) tmem_st16(t_s + C::P_LO + pcol0 + 16, ll);
      }
      tmem_st_wait();
      if (j == AW_NT - 1) {                       // for the epilogue warps, ordered by p_full
        lsum[(it & 1) * 256 + hf * 128 + row] = l;
        if (hf == 0) mfin[(it & 1) * 128 + row] = m_ref;
      }
      tc_fence_before();
      __syncwarp();
      if (lane0) mbar_arrive(&p_full[sb]);
    }
    TW_FLUSH(8, warp == 4 && lane == 0)
  } else {
    // ------------------------------------------------------------------ epilogue warps (4): O / l -> ctx rows
    aw_reg_dec<C::REG_EPILOGUE>();
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_sel = uint32_t(q * 32) << 16;
    int it = 0;
    TW_DECL
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const int qt = item % qtiles, bh = item / qtiles;
      const int h = bh % AW_HEADS, b = bh / AW_HEADS;
      const int ob = it & 1;
      TW_WAIT(1, mbar_wait(&o_full[ob], (it >> 1) & 1));
      tc_fence_after();
      const float l_tot = lsum[ob * 256 + row] + lsum[ob * 256 + 128 + row];
      const float inv = 1.0f / l_tot;
      const int tok = qt * AW_QT + row;
      const long long orow = (long long)b * AW_TOK + tok;
      // training: log-sum-exp of the row in the exp2 domain, P = exp2(s * c - lse), for the fused backward (attention_bwd.cu)
      if (lse != nullptr && tok < AW_TOK)
        lse[(long long)bh * AW_TOK + tok] = mfin[ob * 128 + row] * (0.08838834764831845f * 1.4426950408889634f) + log2f(l_tot);
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t o[32];
        tmem_ld32(C::T_O + ob * AW_D + lane_sel + ch * 32, o);
        tmem_ld_wait();
        if (tok < AW_TOK) {
          uint32_t hh[16], ll[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float v0 = __uint_as_float(o[2 * e]) * inv, v1 = __uint_as_float(o[2 * e + 1]) * inv;
            if (NSPLIT > 1) split_pack2(v0, v1, hh[e], ll[e]); else hh[e] = cvt_bf16x2(v0, v1);
          }
          const long long off = orow * (AW_HEADS * AW_D) + h * AW_D + ch * 32;
          // 32-byte stores (row bases are 64-byte aligned: 2048-byte rows, 32-column chunks)
#pragma unroll
          for (int e = 0; e < 2; ++e)
            st_global_256(ctx_hi + off + 16 * e, hh[8 * e], hh[8 * e + 1], hh[8 * e + 2], hh[8 * e + 3], hh[8 * e + 4], hh[8 * e + 5],
                          hh[8 * e + 6], hh[8 * e + 7]);
          if (NSPLIT > 1) {
#pragma unroll
            for (int e = 0; e < 2; ++e)
              st_global_256(ctx_lo + off + 16 * e, ll[8 * e], ll[8 * e + 1], ll[8 * e + 2], ll[8 * e + 3], ll[8 * e + 4], ll[8 * e + 5],
                            ll[8 * e + 6], ll[8 * e + 7]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[ob]);
    }
    TW_FLUSH(24, warp == 12 && lane == 0)
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 3) tmem_dealloc<1>(0u, C::TMEM_COLS);
}

template <int NSPLIT>
static int launch_attention(const CUtensorMap* tm, __nv_bfloat16* ctx_hi, __nv_bfloat16* ctx_lo, float* lse, int B, int qtiles,
                            cudaStream_t stream) {
  using C = AttnWideCfg<NSPLIT>;
  auto kern = attention_kernel<NSPLIT>;
  static bool attr_done[64] = {false};
  const int dev_ = current_device();
  if (!attr_done[dev_]) {
    EB_CUDA(EB_SET_MAX_SMEM(kern, C::SMEM_BYTES));
    attr_done[dev_] = true;
  }
  const int items = B * AW_HEADS * qtiles;
  const int grid = items < num_sms() ? items : num_sms();
  ProfScope prof("attention_kernel", stream);
  EB_LAUNCH_SMEM(kern, grid, AW_THREADS, C::SMEM_BYTES, stream, tm[0], tm[1], tm[2], tm[3], ctx_hi, ctx_lo, lse, items, qtiles);
  EB_CHECK_LAUNCH("attention_kernel");
  return 0;
}

// qk: (B*576, 2048) = [Q | K] per token, head h at columns h*128; vt: (B*8*128, 576) = V^T per (frame, head)
// query_rows: only the first query_rows tokens of every frame get a context row (all 576 are keys / values)
int attention_run(const __nv_bfloat16* qk_hi, const __nv_bfloat16* qk_lo, const __nv_bfloat16* vt_hi,
                  const __nv_bfloat16* vt_lo, __nv_bfloat16* ctx_hi, __nv_bfloat16* ctx_lo, int B, int nsplit,
                  int query_rows, cudaStream_t stream, float* lse) {
  EB_REQUIRE(query_rows > 0 && query_rows <= AW_TOK, "attention: query_rows must be in (0, 576]");
  const int qtiles = (query_rows + AW_QT - 1) / AW_QT;
  EB_REQUIRE(qk_hi && vt_hi && ctx_hi && B > 0, "attention: bad arguments");
  EB_REQUIRE(nsplit == 1 || (qk_lo && vt_lo && ctx_lo), "attention: bf16x3 mode needs the lo parts");
  EB_REQUIRE(((reinterpret_cast<uintptr_t>(ctx_hi) | reinterpret_cast<uintptr_t>(ctx_lo)) & 31) == 0,
             "attention: the context buffers must be 32-byte aligned");
  CUtensorMap tm[4];
  int rc;
  const long long fs = (long long)AW_TOK * 2 * AW_HEADS * AW_D;  // frame stride of the qk buffer
  for (int part = 0; part < (nsplit == 3 ? 2 : 1); ++part) {
    const __nv_bfloat16* qk = part == 0 ? qk_hi : qk_lo;
    const __nv_bfloat16* vt = part == 0 ? vt_hi : vt_lo;
    // Q and K tiles are both 128 rows x 64 of the [Q | K] buffer: one map, the K tiles at head coordinate 8 + h
    if ((rc = make_operand_tmap(&tm[0 + part], qk, AW_D, AW_TOK, 2 * AW_HEADS * AW_D, 2 * AW_HEADS, AW_D, B, fs, 128)))
      return rc;
    if ((rc = make_operand_tmap(&tm[2 + part], vt, AW_TOK, (long long)B * AW_HEADS * AW_D, AW_TOK, 1, 0, 1, 0, AW_D)))
      return rc;
  }
  if (nsplit == 1) { tm[1] = tm[0]; tm[3] = tm[2]; }
  return nsplit == 3 ? launch_attention<3>(tm, ctx_hi, ctx_lo, lse, B, qtiles, stream)
                     : launch_attention<1>(tm, ctx_hi, ctx_lo, lse, B, qtiles, stream);
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
                           precision == EGOTAP_PREC_BF16 ? 1 : 3, eb::AW_TOK, (cudaStream_t)stream);
}

/* the same with the per-row log-sum-exp (exp2 domain) written to lse[(frame * 8 + head) * 576 + token]: training forward */
extern "C" int egotap_b200_attention_lse(const void* qk_hi, const void* qk_lo, const void* vt_hi, const void* vt_lo,
                                         void* ctx_hi, void* ctx_lo, float* lse, int frames, int precision, void* stream) {
  return eb::attention_run((const __nv_bfloat16*)qk_hi, (const __nv_bfloat16*)qk_lo, (const __nv_bfloat16*)vt_hi,
                           (const __nv_bfloat16*)vt_lo, (__nv_bfloat16*)ctx_hi, (__nv_bfloat16*)ctx_lo, frames,
                           precision == EGOTAP_PREC_BF16 ? 1 : 3, eb::AW_TOK, (cudaStream_t)stream, lse);
}

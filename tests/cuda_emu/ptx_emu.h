// TEST INFRASTRUCTURE ONLY -- functional model of the sm_100a features behind egotap_b200/csrc/ptx.cuh (same function
// names and signatures), for the CUDA-on-CPU emulation (cuda_emu.h).  Semantics follow the PTX ISA 8.7 descriptions the
// product code cites; the model is calibrated by running the GEMM and attention kernels that were verified on the B200.
//   shared-memory addresses  : byte offset inside the CTA's dynamic shared memory; bits [24, 28) select a peer CTA of the
//                              cluster (rank + 1; 0 = the executing CTA) as mapa() would
//   mbarrier (8 bytes)       : pending arrivals, transaction bytes, phase bit; a phase completes when both counts reach 0
//   TMA tiled load           : 64-element (128-byte) rows, 128-byte swizzle (16-byte chunk index ^= row % 8), zero fill
//                              outside the tensor, complete_tx of the whole box on the given barrier
//   tensor memory            : 128 lanes x 512 32-bit columns per CTA; address = lane << 16 | column
//   tcgen05.mma kind::f16    : D[m][n] (+)= sum_k A[m][k] * B[n][k], K = 16 per instruction, operands located through the
//                              real matrix descriptor (start address, SBO, 128-byte swizzle on the address bits) or, for
//                              the TS form, packed bf16 pairs in tensor memory; cta_group::2 spans both CTAs of the pair
// Asynchronous operations (TMA, MMA, commit) are queued at issue and complete, in order, one scheduler round later; a
// TMA destination holds bf16 NaN in between.  A missing wait therefore shows up as NaN / stale results, a wait that can
// never be satisfied as a deadlock (fiber scheduler).  Proxy fences and finer memory-ordering rules are not modelled.
#pragma once
#include <deque>
#include <map>
#include <memory>
#include "cuda_emu.h"

#include <cstdio>
#include <cstdlib>

#define EB_WAIT_TIMEOUT_CYCLES (1LL << 62)
#define EB_DYN_SMEM(name) uint8_t* name = eb_emu::dyn_smem()
#define EB_DYN_SMEM_1K(name) uint8_t* name = eb_emu::dyn_smem()

namespace eb {

inline uint32_t smem_u32(const void* p) {
  const ptrdiff_t off = reinterpret_cast<const uint8_t*>(p) - eb_emu::dyn_smem();
  if (off < 0 || off >= (1 << 18)) { fprintf(stderr, "ptx_emu: pointer is not in dynamic shared memory\n"); abort(); }
  return uint32_t(off);
}
inline uint8_t* smem_ptr(uint32_t addr) {          // shared::cluster address -> host pointer
  const int sel = int(addr >> 24) & 15;
  return eb_emu::smem_of(sel ? sel - 1 : eb_emu::cta_rank()) + (addr & 0xFFFFFF);
}
inline uint32_t lane_id() { return uint32_t(eb_emu::lane()); }
inline bool elect_one() { return eb_emu::lane() == 0; }
inline uint32_t cluster_ctarank() { return uint32_t(eb_emu::cta_rank()); }
inline void cluster_sync_all() { eb_emu::cluster_sync(); }
inline uint32_t mapa(uint32_t addr, uint32_t cta) { return (addr & 0xFFFFFF) | ((cta + 1) << 24); }
inline void fence_proxy_async_smem() {}
inline void fence_proxy_async_all() {}
inline uint32_t ld_acquire_gpu_u32(const unsigned int* p) { eb_emu::yield_wait(); return *p; }   // every poll yields
template <int N>
inline void named_bar_sync(int id) { eb_emu::named_barrier(id, N); }

// ---- mbarrier ---------------------------------------------------------------------------------------------------
struct EmuMbar2 { uint16_t init; uint16_t pending; int32_t tx; };
static_assert(sizeof(EmuMbar2) == 8, "mbarrier storage is 8 bytes");
// the phase bit lives in the top bit of `init` (counts are < 2^15)
inline EmuMbar2* mb(void* p) { return reinterpret_cast<EmuMbar2*>(p); }
inline void mbar_check(EmuMbar2* b) {
  if (b->pending == 0 && b->tx == 0) {
    b->init ^= 0x8000;
    b->pending = b->init & 0x7fff;
    eb_emu::note_progress();
  }
}
inline void mbar_init(uint64_t* bar, uint32_t count) { EmuMbar2* b = mb(bar); b->init = uint16_t(count); b->pending = uint16_t(count); b->tx = 0; }
inline void fence_mbar_init() {}
inline void mbar_arrive_at(void* p) {
  EmuMbar2* b = mb(p);
  if (b->pending == 0) { fprintf(stderr, "ptx_emu: mbarrier over-arrival\n"); abort(); }
  --b->pending;
  eb_emu::note_progress();
  mbar_check(b);
}
inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { mb(bar)->tx += int32_t(bytes); mbar_arrive_at(bar); }
inline void mbar_arrive(uint64_t* bar) { mbar_arrive_at(bar); }
inline void mbar_arrive_remote(uint64_t* bar, uint32_t cta) { mbar_arrive_at(smem_ptr(mapa(smem_u32(bar), cta))); }
inline void mbar_complete_tx(void* p, uint32_t bytes) { EmuMbar2* b = mb(p); b->tx -= int32_t(bytes); eb_emu::note_progress(); mbar_check(b); }
inline bool mbar_try_wait(uint64_t* bar, uint32_t parity) { return uint32_t((mb(bar)->init >> 15) & 1) != (parity & 1); }
inline void mbar_wait(uint64_t* bar, uint32_t parity) { eb_emu::wait_phase(&mb(bar)->init, parity); }

// ---- TMA --------------------------------------------------------------------------------------------------------
struct EmuTmap {            // lives in the 128 bytes of a CUtensorMap (filled by host_util.cuh make_operand_tmap)
  const uint8_t* base;
  uint64_t dims[4];         // elements: K, rows, g0, g1
  uint64_t strides[3];      // bytes: row, g0, g1
  uint32_t box_rows;
  uint32_t magic;
};
static_assert(sizeof(EmuTmap) <= sizeof(CUtensorMap), "emulated tensor map must fit the opaque storage");
inline void tma_prefetch_desc(const CUtensorMap*) {}
inline void tma_load_now(uint8_t* dst, const EmuTmap* tm, void* bar, int c0, int c1, int c2, int c3) {
  for (uint32_t r = 0; r < tm->box_rows; ++r)
    for (uint32_t i = 0; i < 64; ++i) {
      const uint64_t k = uint64_t(c0) + i, row = uint64_t(c1) + r;
      uint16_t v = 0;
      if (c0 >= 0 && c1 >= 0 && k < tm->dims[0] && row < tm->dims[1] && uint64_t(c2) < tm->dims[2] && uint64_t(c3) < tm->dims[3])
        memcpy(&v, tm->base + k * 2 + row * tm->strides[0] + uint64_t(c2) * tm->strides[1] + uint64_t(c3) * tm->strides[2], 2);
      uint32_t off = r * 128 + i * 2;
      off ^= ((off >> 7) & 7) << 4;                       // 128-byte swizzle
      memcpy(dst + off, &v, 2);
    }
  mbar_complete_tx(bar, tm->box_rows * 128);
}
// asynchronous: the destination is poisoned (bf16 NaN) at issue and filled, with complete_tx, one scheduler round later --
// a consumer that does not wait on the barrier computes on NaN
inline void tma_load_to(uint8_t* dst, const CUtensorMap* tm_, void* bar, int c0, int c1, int c2, int c3) {
  const EmuTmap tm = *reinterpret_cast<const EmuTmap*>(tm_);
  if (tm.magic != 0x7e4a0001u) { fprintf(stderr, "ptx_emu: not an emulated tensor map\n"); abort(); }
  const uint32_t doff = uint32_t(dst - eb_emu::dyn_smem());
  if (doff % 1024) { fprintf(stderr, "ptx_emu: TMA destination must be 1024-byte aligned for the 128B swizzle\n"); abort(); }
  for (uint32_t i = 0; i < tm.box_rows * 64; ++i) { const uint16_t nan = 0x7fc0; memcpy(dst + i * 2, &nan, 2); }
  eb_emu::defer_tma([=]() { tma_load_now(dst, &tm, bar, c0, c1, c2, c3); });
}
inline void tma_prefetch_4d(const CUtensorMap*, int, int, int, int) {}   // L2 prefetch: no functional effect
// ---- TMA store + bulk async-groups.  The shared-memory source is read WHEN THE STORE EXECUTES (a later scheduler round), so a
// box that is rewritten before cp.async.bulk.wait_group.read allowed it shows up as corrupted output.
struct EmuBulkGroup { int outstanding = 0; };
struct EmuBulkState { std::shared_ptr<EmuBulkGroup> open; std::deque<std::shared_ptr<EmuBulkGroup>> committed; };
inline EmuBulkState& emu_bulk_state() {
  static std::map<uint64_t, EmuBulkState> st;
  return st[(uint64_t(blockIdx.x) << 32) | uint64_t(threadIdx.x)];
}
inline void tma_store_now(const uint8_t* src, const EmuTmap* tm, int c0, int c1, int c2, int c3) {
  for (uint32_t r = 0; r < tm->box_rows; ++r)
    for (uint32_t i = 0; i < 64; ++i) {
      const uint64_t k = uint64_t(c0) + i, row = uint64_t(c1) + r;
      if (!(c0 >= 0 && c1 >= 0 && k < tm->dims[0] && row < tm->dims[1] && uint64_t(c2) < tm->dims[2] && uint64_t(c3) < tm->dims[3])) continue;
      uint32_t off = r * 128 + i * 2;
      off ^= ((off >> 7) & 7) << 4;                       // 128-byte swizzle
      memcpy(const_cast<uint8_t*>(tm->base) + k * 2 + row * tm->strides[0] + uint64_t(c2) * tm->strides[1] + uint64_t(c3) * tm->strides[2],
             src + off, 2);
    }
}
inline void tma_store_4d(const CUtensorMap* tm_, const void* src_, int c0, int c1, int c2, int c3) {
  const EmuTmap tm = *reinterpret_cast<const EmuTmap*>(tm_);
  if (tm.magic != 0x7e4a0001u) { fprintf(stderr, "ptx_emu: not an emulated tensor map\n"); abort(); }
  const uint8_t* src = reinterpret_cast<const uint8_t*>(src_);
  if (uint32_t(src - eb_emu::dyn_smem()) % 1024) { fprintf(stderr, "ptx_emu: TMA source must be 1024-byte aligned for the 128B swizzle\n"); abort(); }
  EmuBulkState& st = emu_bulk_state();
  if (!st.open) st.open = std::make_shared<EmuBulkGroup>();
  std::shared_ptr<EmuBulkGroup> grp = st.open;
  ++grp->outstanding;
  eb_emu::defer_tma([=]() { tma_store_now(src, &tm, c0, c1, c2, c3); --grp->outstanding; });
}
inline void bulk_commit() {
  EmuBulkState& st = emu_bulk_state();
  if (!st.open) st.open = std::make_shared<EmuBulkGroup>();
  st.committed.push_back(st.open);
  st.open.reset();
}
template <int N>
inline void bulk_wait_read() {
  EmuBulkState& st = emu_bulk_state();
  for (;;) {
    while (!st.committed.empty() && st.committed.front()->outstanding == 0) st.committed.pop_front();
    int pending = 0;
    for (auto& g : st.committed) pending += g->outstanding > 0 ? 1 : 0;
    // groups complete in order in this model: "at most N pending" = at most N committed groups left
    if (int(st.committed.size()) <= N || pending == 0) return;
    eb_emu::yield_wait();
  }
}
template <int N>
inline void bulk_wait() { bulk_wait_read<N>(); }
inline void tma_load_4d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
  tma_load_to(reinterpret_cast<uint8_t*>(dst), tm, bar, c0, c1, c2, c3);
}
inline void tma_load_4d_2sm(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
  tma_load_to(reinterpret_cast<uint8_t*>(dst), tm, smem_ptr(mapa(smem_u32(bar), 0)), c0, c1, c2, c3);   // leader's barrier
}

// ---- tensor memory ------------------------------------------------------------------------------------------------
template <int CG>
inline void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  if (eb_emu::lane() != 0) return;
  uint32_t& next = eb_emu::tmem_next_col();
  if (ncols < 32 || (ncols & (ncols - 1)) || next + ncols > 512) { fprintf(stderr, "ptx_emu: bad TMEM allocation of %u columns\n", ncols); abort(); }
  *dst_smem = next;
  next += ncols;
}
template <int CG>
inline void tmem_dealloc(uint32_t, uint32_t) {}
inline void tc_fence_before() {}
inline void tc_fence_after() {}

inline void st_global_256(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f, uint32_t g, uint32_t h) {
  if (reinterpret_cast<uintptr_t>(p) % 32) { fprintf(stderr, "ptx_emu: st.global.v8 address not 32-byte aligned\n"); abort(); }
  const uint32_t v[8] = {a, b, c, d, e, f, g, h};
  memcpy(p, v, 32);
}
inline void ld_global_256(const void* p, float4& lo, float4& hi) {
  if (reinterpret_cast<uintptr_t>(p) % 32) { fprintf(stderr, "ptx_emu: ld.global.v8 address not 32-byte aligned\n"); abort(); }
  memcpy(&lo, p, 16);
  memcpy(&hi, static_cast<const char*>(p) + 16, 16);
}

constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major = 0, int b_mn_major = 0) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(a_mn_major) << 15) | (uint32_t(b_mn_major) << 16) |
         (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}
inline uint64_t make_sdesc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3FFFF) >> 4);
  d |= uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;
  return d;
}
inline uint32_t sdesc_lo(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3FFFu) | (1u << 16); }
inline uint64_t sdesc_at(uint32_t base_lo, uint32_t byte_off) { return (uint64_t(0x40004040u) << 32) | uint64_t(base_lo + (byte_off >> 4)); }

inline float emu_bf16_at(const uint8_t* smem, uint32_t addr) {
  addr ^= ((addr >> 7) & 7) << 4;                         // the swizzle is a function of the shared-memory address bits
  uint16_t v;
  memcpy(&v, smem + addr, 2);
  const uint32_t u = uint32_t(v) << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
// operand tile (rows x 16) of a K-major, 128B-swizzled shared-memory matrix descriptor, as floats
inline void emu_load_smem_operand(const uint8_t* smem, uint64_t desc, int rows, float* out) {
  if (((desc >> 61) & 7) != 2 || ((desc >> 46) & 3) != 1) { fprintf(stderr, "ptx_emu: unsupported matrix descriptor\n"); abort(); }
  const uint32_t start = uint32_t(desc & 0x3FFF) << 4, sbo = uint32_t((desc >> 32) & 0x3FFF) << 4;
  for (int r = 0; r < rows; ++r)
    for (int ch = 0; ch < 2; ++ch) {                      // the 16 K-elements of a row are two 16-byte chunks
      uint32_t addr = start + uint32_t(r / 8) * sbo + uint32_t(r % 8) * 128 + uint32_t(ch) * 16;
      addr ^= ((addr >> 7) & 7) << 4;                     // 128-byte swizzle on the shared-memory address bits
      uint16_t v[8];
      memcpy(v, smem + addr, 16);
      for (int k = 0; k < 8; ++k) out[r * 16 + ch * 8 + k] = __uint_as_float(uint32_t(v[k]) << 16);
    }
}
// operand tile (rows x 16) of an MN-major, 128B-swizzled descriptor: canonical layout ((8,8,m),(8,k)) : ((1,8,LBO),(64,SBO)) in
// bf16 elements (CUTLASS cute/atom/mma_traits_sm100.hpp, make_umma_desc<Major::MN>): the M/N index runs along the 128-byte row
// (64 elements), the next 64 M/N elements are LBO bytes further; the K index selects the row: 8 rows per 1 KB swizzle atom, the
// next 8 rows SBO bytes further.  out[r * 16 + k] like the K-major loader.
inline void emu_load_smem_operand_mn(const uint8_t* smem, uint64_t desc, int rows, float* out) {
  if (((desc >> 61) & 7) != 2 || ((desc >> 46) & 3) != 1) { fprintf(stderr, "ptx_emu: unsupported matrix descriptor\n"); abort(); }
  const uint32_t start = uint32_t(desc & 0x3FFF) << 4, lbo = uint32_t((desc >> 16) & 0x3FFF) << 4, sbo = uint32_t((desc >> 32) & 0x3FFF) << 4;
  for (int r = 0; r < rows; ++r)
    for (int k = 0; k < 16; ++k) {
      uint32_t addr = start + uint32_t(r / 64) * lbo + uint32_t(k / 8) * sbo + uint32_t(k % 8) * 128 + uint32_t(r % 64) * 2;
      out[r * 16 + k] = emu_bf16_at(smem, addr);
    }
}
inline void emu_load_operand(const uint8_t* smem, uint64_t desc, int rows, bool mn_major, float* out) {
  if (mn_major) emu_load_smem_operand_mn(smem, desc, rows, out); else emu_load_smem_operand(smem, desc, rows, out);
}
__attribute__((no_sanitize("alignment"))) inline void emu_mma_accumulate(uint32_t* tmem, uint32_t lane0, uint32_t col, int M,
                                                                         int N, const float* A, const float* B, bool acc) {
  static float Bt[16 * 256];                              // B transposed to [k][n]: the n loop then vectorises
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < 16; ++k) Bt[k * 256 + n] = B[n * 16 + k];
  float row[256];
  for (int m = 0; m < M; ++m) {
    for (int n = 0; n < N; ++n) row[n] = 0.f;
    for (int k = 0; k < 16; ++k) {
      const float a = A[m * 16 + k];
      const float* b = Bt + k * 256;
      for (int n = 0; n < N; ++n) row[n] += a * b[n];
    }
    float* __restrict__ d = reinterpret_cast<float*>(tmem + (lane0 + m) * 512 + col);   // (built with -fno-strict-aliasing)
    if (acc) for (int n = 0; n < N; ++n) d[n] += row[n];
    else for (int n = 0; n < N; ++n) d[n] = row[n];
  }
}
template <int CG>
inline void umma_bf16_now(int rank, uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  const int M = int((idesc >> 24) & 31) << 4, N = int((idesc >> 17) & 63) << 3;
  const uint32_t col = tmem_d & 0xFFFF, lane0 = tmem_d >> 16;
  const bool a_mn = ((idesc >> 15) & 1) != 0, b_mn = ((idesc >> 16) & 1) != 0;
  if (col + N > 512 || lane0 != 0) { fprintf(stderr, "ptx_emu: unsupported MMA (idesc %x tmem %x)\n", idesc, tmem_d); abort(); }
  static float A[256 * 16], B[256 * 16];
  if (CG == 1) {
    if (M != 128) { fprintf(stderr, "ptx_emu: cta_group::1 MMA with M = %d\n", M); abort(); }
    const uint8_t* smem = eb_emu::smem_of(rank);
    emu_load_operand(smem, adesc, M, a_mn, A);
    emu_load_operand(smem, bdesc, N, b_mn, B);
    emu_mma_accumulate(eb_emu::tmem_of(rank), 0, col, M, N, A, B, accumulate != 0);
  } else {
    if (M != 256 || rank != 0) { fprintf(stderr, "ptx_emu: cta_group::2 MMA must be issued by the leader with M = 256\n"); abort(); }
    for (int r = 0; r < 2; ++r) {       // A rows and B rows are split across the pair, at the same shared-memory offsets
      emu_load_operand(eb_emu::smem_of(r), adesc, 128, a_mn, A + r * 128 * 16);
      emu_load_operand(eb_emu::smem_of(r), bdesc, N / 2, b_mn, B + r * (N / 2) * 16);
    }
    for (int r = 0; r < 2; ++r) emu_mma_accumulate(eb_emu::tmem_of(r), 0, col, 128, N, A + r * 128 * 16, B, accumulate != 0);
  }
}
// asynchronous: queued at issue, executed in issue order one scheduler round later (operands are read THEN, so a ring slot
// that is recycled before the commit that guards it has arrived feeds the MMA the wrong data)
template <int CG>
inline void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (eb_emu::lane() != 0) return;                        // elect.sync: one issuing thread
  const int rank = eb_emu::cta_rank();
  eb_emu::defer([=]() { umma_bf16_now<CG>(rank, tmem_d, adesc, bdesc, idesc, accumulate); });
}
template <int CG>
inline void umma_commit(uint64_t* bar) {              // arrives once every previously issued MMA has completed
  if (eb_emu::lane() != 0) return;
  if (CG == 1) {
    eb_emu::defer([=]() { mbar_arrive_at(bar); });
  } else {
    void* b0 = smem_ptr(mapa(smem_u32(bar), 0));
    void* b1 = smem_ptr(mapa(smem_u32(bar), 1));
    eb_emu::defer([=]() { mbar_arrive_at(b0); mbar_arrive_at(b1); });
  }
}
inline void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  const uint32_t* t = eb_emu::tmem_of(eb_emu::cta_rank());
  const uint32_t lane = (taddr >> 16) + uint32_t(eb_emu::lane()), col = taddr & 0xFFFF;
  if (lane >= 128 || col + 32 > 512) { fprintf(stderr, "ptx_emu: tcgen05.ld out of range (%x)\n", taddr); abort(); }
  for (int j = 0; j < 32; ++j) r[j] = t[lane * 512 + col + j];
}
inline void tmem_ld_wait() {}
inline void tmem_st_n(uint32_t taddr, const uint32_t* r, int n) {
  uint32_t* t = eb_emu::tmem_of(eb_emu::cta_rank());
  const uint32_t lane = (taddr >> 16) + uint32_t(eb_emu::lane()), col = taddr & 0xFFFF;
  if (lane >= 128 || col + n > 512) { fprintf(stderr, "ptx_emu: tcgen05.st out of range (%x)\n", taddr); abort(); }
  for (int j = 0; j < n; ++j) t[lane * 512 + col + j] = r[j];
  eb_emu::note_progress();
}
inline void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) { tmem_st_n(taddr, r, 32); }
inline void tmem_st16(uint32_t taddr, const uint32_t* r) { tmem_st_n(taddr, r, 16); }
inline void tmem_st_wait() {}
template <int N> inline void warpgroup_reg_dec() {}
template <int N> inline void warpgroup_reg_inc() {}
inline int pin_reg(int v) { return v; }
// TS form, cta_group::1: A (128 x 16 bf16) from tensor memory -- lane = row, 8 consecutive 32-bit columns, each holding
// the K-elements (2c, 2c + 1) as a packed bf16 pair with the even one in the low half; B from shared memory
inline void umma_bf16_ts_now(int rank, uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  const int M = int((idesc >> 24) & 31) << 4, N = int((idesc >> 17) & 63) << 3;
  const uint32_t col = tmem_d & 0xFFFF, acol = tmem_a & 0xFFFF;
  if (M != 128 || (tmem_d >> 16) != 0 || (tmem_a >> 16) != 0 || col + N > 512 || acol + 8 > 512) { fprintf(stderr, "ptx_emu: unsupported TS MMA\n"); abort(); }
  static float A[128 * 16], B[256 * 16];
  uint32_t* t = eb_emu::tmem_of(rank);
  for (int m = 0; m < 128; ++m)
    for (int c = 0; c < 8; ++c) {
      const uint32_t w = t[m * 512 + acol + c];
      A[m * 16 + 2 * c] = __uint_as_float(w << 16);
      A[m * 16 + 2 * c + 1] = __uint_as_float(w & 0xffff0000u);
    }
  emu_load_operand(eb_emu::smem_of(rank), bdesc, N, ((idesc >> 16) & 1) != 0, B);
  emu_mma_accumulate(t, 0, col, M, N, A, B, accumulate != 0);
}
inline void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (eb_emu::lane() != 0) return;
  const int rank = eb_emu::cta_rank();
  eb_emu::defer([=]() { umma_bf16_ts_now(rank, tmem_d, tmem_a, bdesc, idesc, accumulate); });
}

}  // namespace eb

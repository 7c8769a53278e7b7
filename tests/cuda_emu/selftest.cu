// TEST INFRASTRUCTURE ONLY -- self-test kernels of the emulation (compiled into the emulation library only): a
// producer / consumer hand-off through two mbarriers written with the product's own wrappers (ptx.cuh names).  The
// correct protocol must run; the broken one (consumer waits on the wrong phase parity once) must be reported as a
// deadlock by the fiber scheduler instead of hanging or silently passing.
#include "ptx.cuh"

namespace {
void handoff_kernel(int rounds, int break_at, int* out) {
  using namespace eb;
  EB_DYN_SMEM(smem);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + 1;
  int* slot = reinterpret_cast<int*>(smem + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(full, 1); mbar_init(empty, 1); fence_mbar_init(); }
  __syncthreads();
  if (warp == 0) {
    if (lane == 0)
      for (int i = 0; i < rounds; ++i) {
        mbar_wait(empty, (i & 1) ^ 1);
        *slot = 100 + i;
        mbar_arrive(full);
      }
  } else if (lane == 0) {
    int sum = 0;
    for (int i = 0; i < rounds; ++i) {
      mbar_wait(full, (i == break_at) ? ((i & 1) ^ 1) : (i & 1));   // break_at >= 0: one wait on the wrong parity
      sum += *slot;
      mbar_arrive(empty);
    }
    *out = sum;
  }
}
}  // namespace

extern "C" int emu_selftest_handoff(int rounds, int break_at, int* out) {
  eb_emu::launch_ex(dim3(1), dim3(64), 1, 1024, [=]() { handoff_kernel(rounds, break_at, out); });
  return 0;
}

// asynchrony self-test: a TMA load whose consumer does / does not wait on the barrier.  Without the wait the consumer
// must see the poison the emulation puts into the destination at issue (bf16 NaN), not the data.
#include "host_util.cuh"
namespace {
void tma_kernel(CUtensorMap tm, int wait, float* out) {
  using namespace eb;
  EB_DYN_SMEM_1K(smem);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 64 * 128);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
    mbar_expect_tx(bar, 64 * 128);
    tma_load_4d(smem, &tm, bar, 0, 0, 0, 0);
    if (wait) mbar_wait(bar, 0);
    uint16_t v;
    memcpy(&v, smem, 2);                       // element (row 0, k 0): chunk 0 of row 0 is not moved by the swizzle
    *out = __uint_as_float(uint32_t(v) << 16);
  }
}
}  // namespace

extern "C" int emu_selftest_tma(const void* src_bf16_64x64, int wait, float* out) {
  CUtensorMap tm;
  if (eb::make_operand_tmap(&tm, src_bf16_64x64, 64, 64, 64, 1, 0, 1, 0, 64)) return -1;
  eb_emu::launch_ex(dim3(1), dim3(32), 1, 64 * 128 + 64, [=]() { tma_kernel(tm, wait, out); });
  return 0;
}

// schedule self-test: a neighbour exchange through shared memory WITHOUT the barrier between write and read (sync = 0).
// Which neighbour's value a thread sees then depends on the order the threads run in: the ascending schedule hides the
// dependence on lower-numbered threads, the descending one the dependence on higher-numbered threads, so the three
// schedules of the emulation give different answers for the racy kernel and the same answer for the correct one.
namespace {
void neighbour_kernel(int sync, int* out) {
  __shared__ int buf[64];
  const int t = threadIdx.x;
  buf[t] = -1;
  __syncthreads();
  buf[t] = t;
  if (sync) __syncthreads();
  out[t] = buf[(t + 1) & 63] + 1000 * buf[(t + 63) & 63];
}
}  // namespace

extern "C" int emu_selftest_neighbours(int sync, int* out64) {
  eb_emu::launch(dim3(1), dim3(64), true, [=]() { neighbour_kernel(sync, out64); });
  return 0;
}

// TEST INFRASTRUCTURE ONLY -- runtime of the CUDA-on-CPU emulation (see cuda_emu.h) plus the few symbols the emulated
// translation units expect from the parts of libegotap_b200.so that are not emulated.
#include "cuda_emu.h"

#include <ucontext.h>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <memory>
#include <deque>
#include <vector>

namespace eb_emu {

uint3 g_threadIdx, g_blockIdx;
dim3 g_blockDim, g_gridDim;

extern int g_num_sms;
namespace {
constexpr size_t kStack = 512 * 1024;
constexpr int kMaxWarps = 64, kNamed = 16;
struct Fiber {
  ucontext_t ctx;
  std::unique_ptr<char[]> stack;          // not zero-filled: pages are touched only as far as the fiber really uses them
  bool done = false;
  int cta = 0, warp = 0, lane = 0, tid = 0;
  // what a blocked fiber waits for, so the scheduler does not switch to it in vain: a barrier generation counter that
  // must move on, or an mbarrier phase bit (top bit of its first 16-bit word) that must differ from `parity`
  const unsigned* wait_gen = nullptr;
  unsigned wait_val = 0;
  const uint16_t* wait_phase = nullptr;
  unsigned wait_parity = 0;
};
struct Cta {
  std::vector<uint8_t> smem_store;
  uint8_t* smem = nullptr;
  std::vector<uint32_t> tmem;
  uint32_t tmem_next = 0;
  uint3 block_idx;
  int live = 0, block_waiting = 0;
  unsigned block_gen = 0;
  int warp_live[kMaxWarps], warp_waiting[kMaxWarps];
  unsigned warp_gen[kMaxWarps];
  double warp_buf[kMaxWarps][32];
  int named_waiting[kNamed], named_need[kNamed];
  unsigned named_gen[kNamed];
};
std::vector<Fiber> fibers;
std::vector<Cta> ctas;
ucontext_t sched_ctx;
Fiber* cur = nullptr;
bool in_coop = false;
const std::function<void()>* cur_body = nullptr;
int cluster_live = 0, cluster_waiting = 0;
unsigned cluster_gen = 0;
unsigned long long progress = 0;

// Thread schedule: 0 = ascending thread index (default), 1 = descending, 2 = reshuffled every scheduler round.  CUDA
// promises no order between threads that are not separated by a barrier, so results must not depend on this; running the
// same kernel under all three exposes a missing __syncthreads / __syncwarp (or an in-place hazard between the threads of an
// independent-thread kernel) whichever direction the dependence has.
int sched_mode = 0;
uint64_t sched_rng = 0x9E3779B97F4A7C15ull;
uint32_t rnd() {
  sched_rng ^= sched_rng << 13; sched_rng ^= sched_rng >> 7; sched_rng ^= sched_rng << 17;
  return uint32_t(sched_rng >> 32);
}
void make_order(std::vector<int>& order, int n) {
  order.resize(n);
  for (int i = 0; i < n; ++i) order[i] = sched_mode == 1 ? n - 1 - i : i;
  if (sched_mode == 2)
    for (int i = n - 1; i > 0; --i) std::swap(order[i], order[rnd() % unsigned(i + 1)]);
}

// Launch limits of the real device (sm_100a), checked on every launch; with dry_run set the launch ends here, so the host
// side of a whole step (argument checks, scratch sizing, tensor maps, grid shapes) can be exercised at the full batch sizes
// of BASELINE.json without executing -- or even touching the memory of -- the kernels.
int dry_run = 0;
long long dry_launches = 0;
void check_launch_limits(dim3 grid, dim3 block, size_t dyn_smem_bytes) {
  const unsigned long long threads = 1ull * block.x * block.y * block.z;
  const bool ok = grid.x >= 1 && grid.y >= 1 && grid.z >= 1 && grid.x <= 2147483647u && grid.y <= 65535u && grid.z <= 65535u &&
                  threads >= 1 && threads <= 1024 && block.z <= 64 && dyn_smem_bytes <= 227u * 1024u;
  if (!ok) {
    fprintf(stderr, "cuda_emu: launch outside the device limits: grid (%u, %u, %u) block (%u, %u, %u) dynamic smem %zu\n", grid.x,
            grid.y, grid.z, block.x, block.y, block.z, dyn_smem_bytes);
    abort();
  }
}

Cta& C() { return ctas[cur->cta]; }
void yield() { swapcontext(&cur->ctx, &sched_ctx); }

void trampoline() {
  (*cur_body)();
  cur->done = true;
  swapcontext(&cur->ctx, &sched_ctx);
}

void release_ready_barriers() {
  for (Cta& c : ctas) {
    if (c.block_waiting > 0 && c.block_waiting == c.live) { c.block_waiting = 0; ++c.block_gen; ++progress; }
    for (int w = 0; w < kMaxWarps; ++w)
      if (c.warp_waiting[w] > 0 && c.warp_waiting[w] == c.warp_live[w]) { c.warp_waiting[w] = 0; ++c.warp_gen[w]; ++progress; }
    for (int i = 0; i < kNamed; ++i)
      if (c.named_waiting[i] > 0 && c.named_waiting[i] >= c.named_need[i]) { c.named_waiting[i] = 0; ++c.named_gen[i]; ++progress; }
  }
  if (cluster_waiting > 0 && cluster_waiting == cluster_live) { cluster_waiting = 0; ++cluster_gen; ++progress; }
}

void need_coop(const char* what) {
  if (!in_coop) { fprintf(stderr, "cuda_emu: %s in a kernel launched with EB_LAUNCH\n", what); abort(); }
}

void wait_gen(const unsigned* gen, unsigned val) {
  while (*gen == val) {
    cur->wait_gen = gen;
    cur->wait_val = val;
    yield();
  }
}

void warp_barrier() {
  Cta& c = C();
  const int w = cur->warp;
  const unsigned gen = c.warp_gen[w];
  ++c.warp_waiting[w];
  wait_gen(&c.warp_gen[w], gen);
}
}  // namespace

namespace {
std::vector<std::function<void()>> deferred;
// "late TMA" mode (emu_set_tma_latency): bulk-tensor loads complete a random number of scheduler rounds after issue (in issue
// order), so protocols that only work while loads land promptly -- e.g. an mbarrier parity wait by a thread that has not seen
// the barrier's previous phase complete -- fail here as they do on a busy GPU
struct LateOp { std::function<void()> op; unsigned long long due; };
std::deque<LateOp> late_ops;
unsigned long long round_no = 0, tma_rng = 88172645463325252ull;
int tma_latency_max = 0;
}
void defer(std::function<void()> op) { deferred.push_back(std::move(op)); }
void defer_tma(std::function<void()> op) {
  if (tma_latency_max <= 0) { defer(std::move(op)); return; }
  tma_rng ^= tma_rng << 13; tma_rng ^= tma_rng >> 7; tma_rng ^= tma_rng << 17;
  unsigned long long due = round_no + 1 + tma_rng % (unsigned long long)tma_latency_max;
  if (!late_ops.empty() && late_ops.back().due > due) due = late_ops.back().due;
  late_ops.push_back({std::move(op), due});
}
static void run_deferred() {
  // asynchronous operations (TMA, tcgen05.mma, tcgen05.commit) issued during the previous round complete now, in issue
  // order: a consumer that did not wait for them has already run on stale / poisoned data
  ++round_no;
  for (size_t i = 0; i < deferred.size(); ++i) { deferred[i](); ++progress; }
  deferred.clear();
  while (!late_ops.empty() && late_ops.front().due <= round_no) { late_ops.front().op(); late_ops.pop_front(); ++progress; }
  if (!late_ops.empty()) ++progress;     // loads in flight: time passing is progress (not a deadlock)
}
void note_progress() { ++progress; }
void yield_wait() { need_coop("a spinning wait"); yield(); }
void wait_phase(const void* mbar_word, unsigned parity) {
  need_coop("an mbarrier wait");
  const uint16_t* w = static_cast<const uint16_t*>(mbar_word);
  while (unsigned((*w >> 15) & 1) == (parity & 1)) {
    cur->wait_phase = w;
    cur->wait_parity = parity & 1;
    yield();
  }
}
uint8_t* dyn_smem() { return C().smem; }
uint8_t* smem_of(int r) { return ctas[r].smem; }
uint32_t* tmem_of(int r) { return ctas[r].tmem.data(); }
uint32_t& tmem_next_col() { return C().tmem_next; }
int cta_rank() { return cur ? cur->cta : 0; }
int lane() { return cur ? cur->lane : 0; }
int g_num_sms = 6;                 // a small "device": persistent kernels then loop over several tiles per CTA
int num_sms() { return g_num_sms; }

void syncthreads() {
  need_coop("__syncthreads()");
  Cta& c = C();
  const unsigned gen = c.block_gen;
  ++c.block_waiting;
  wait_gen(&c.block_gen, gen);
}
void syncwarp() { need_coop("__syncwarp()"); warp_barrier(); }
void named_barrier(int id, int nthreads) {
  need_coop("bar.sync");
  Cta& c = C();
  if (id < 0 || id >= kNamed) { fprintf(stderr, "cuda_emu: named barrier id %d\n", id); abort(); }
  const unsigned gen = c.named_gen[id];
  c.named_need[id] = nthreads;
  ++c.named_waiting[id];
  wait_gen(&c.named_gen[id], gen);
}
void cluster_sync() {
  need_coop("cluster barrier");
  const unsigned gen = cluster_gen;
  ++cluster_waiting;
  wait_gen(&cluster_gen, gen);
}

double shfl_xor(double v, int lane_mask) {
  need_coop("warp shuffle");
  Cta& c = C();
  c.warp_buf[cur->warp][cur->lane] = v;
  warp_barrier();
  const double r = c.warp_buf[cur->warp][cur->lane ^ lane_mask];
  warp_barrier();
  return r;
}

double shfl_idx(double v, int src_lane) {
  need_coop("warp shuffle");
  Cta& c = C();
  c.warp_buf[cur->warp][cur->lane] = v;
  warp_barrier();
  const double r = c.warp_buf[cur->warp][src_lane & 31];
  warp_barrier();
  return r;
}

bool any_sync(bool pred) {
  need_coop("warp vote");
  Cta& c = C();
  c.warp_buf[cur->warp][cur->lane] = pred ? 1.0 : 0.0;
  warp_barrier();
  bool r = false;
  for (int i = 0; i < 32; ++i) r = r || c.warp_buf[cur->warp][i] != 0.0;   // (all kernels here use full warps)
  warp_barrier();
  return r;
}

void launch(dim3 grid, dim3 block, bool cooperative, const std::function<void()>& body) {
  if (cooperative) { launch_ex(grid, block, 1, 0, body); return; }
  check_launch_limits(grid, block, 0);
  if (dry_run) { ++dry_launches; return; }
  g_blockDim = block;
  g_gridDim = grid;
  in_coop = false;
  cur = nullptr;
  const int nthreads = int(block.x * block.y * block.z);
  const long long nblocks = (long long)grid.x * grid.y * grid.z;
  std::vector<int> torder;
  make_order(torder, nthreads);
  for (long long i = 0; i < nblocks; ++i) {
    // blocks: ascending, descending, or a stride walk coprime with the block count from a random start
    long long b = sched_mode == 1 ? nblocks - 1 - i : i;
    if (sched_mode == 2) {
      static const long long strides[] = {7919, 104729, 1299709};
      long long st = 1;
      for (long long c : strides) if (nblocks % c != 0) { st = c; break; }     // prime and not a divisor: coprime
      b = (i * (st % nblocks) + 17) % nblocks;
    }
    g_blockIdx = make_uint3(unsigned(b % grid.x), unsigned((b / grid.x) % grid.y), unsigned(b / ((long long)grid.x * grid.y)));
    if (sched_mode == 2 && (i & 7) == 0) make_order(torder, nthreads);
    for (int t = 0; t < nthreads; ++t) {
      g_threadIdx = make_uint3(unsigned(torder[t]), 0, 0);
      body();
    }
  }
}

void launch_ex(dim3 grid, dim3 block, int cluster, size_t dyn_smem_bytes, const std::function<void()>& body) {
  check_launch_limits(grid, block, dyn_smem_bytes);
  if (dry_run) { ++dry_launches; return; }
  g_blockDim = block;
  g_gridDim = grid;
  const int nthreads = int(block.x * block.y * block.z);
  if (block.y != 1 || block.z != 1 || nthreads > 2048 || cluster < 1 || (cluster > 1 && (grid.y != 1 || grid.z != 1 || grid.x % cluster))) {
    fprintf(stderr, "cuda_emu: unsupported launch shape\n");
    abort();
  }
  const long long nblocks = (long long)grid.x * grid.y * grid.z;
  ctas.resize(cluster);
  for (Cta& c : ctas) {
    c.smem_store.assign(dyn_smem_bytes + 2048, 0xCD);                      // poisoned: uninitialised reads stand out
    c.smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(c.smem_store.data()) + 1023) & ~uintptr_t(1023));
    c.tmem.assign(128 * 512, 0x7fc00000u);                                 // NaN-filled tensor memory
  }
  if (int(fibers.size()) < cluster * nthreads) fibers.resize(cluster * nthreads);
  in_coop = true;
  cur_body = &body;
  for (long long b0 = 0; b0 < nblocks; b0 += cluster) {
    cluster_live = cluster * nthreads;
    cluster_waiting = 0;
    for (int k = 0; k < cluster; ++k) {
      Cta& c = ctas[k];
      const long long b = b0 + k;
      c.block_idx = make_uint3(unsigned(b % grid.x), unsigned((b / grid.x) % grid.y), unsigned(b / ((long long)grid.x * grid.y)));
      c.live = nthreads;
      c.block_waiting = 0;
      c.tmem_next = 0;
      for (int w = 0; w < kMaxWarps; ++w) { c.warp_live[w] = 0; c.warp_waiting[w] = 0; }
      for (int i = 0; i < kNamed; ++i) { c.named_waiting[i] = 0; c.named_need[i] = 1 << 30; }
      for (int t = 0; t < nthreads; ++t) {
        Fiber& f = fibers[k * nthreads + t];
        if (!f.stack) f.stack.reset(new char[kStack]);
        f.done = false;
        f.wait_gen = nullptr; f.wait_phase = nullptr;
        f.cta = k; f.tid = t; f.warp = t / 32; f.lane = t % 32;
        ++c.warp_live[f.warp];
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack.get();
        f.ctx.uc_stack.ss_size = kStack;
        f.ctx.uc_link = &sched_ctx;
        makecontext(&f.ctx, trampoline, 0);
      }
    }
    int idle_rounds = 0;
    std::vector<int> order;
    make_order(order, cluster * nthreads);
    while (cluster_live > 0) {
      const unsigned long long before = progress;
      run_deferred();
      if (sched_mode == 2) make_order(order, cluster * nthreads);
      for (int oi = 0; oi < cluster * nthreads; ++oi) {
        Fiber& f = fibers[order[oi]];
        if (f.done) continue;
        if (f.wait_gen) {
          if (*f.wait_gen == f.wait_val) continue;
          f.wait_gen = nullptr;
        }
        if (f.wait_phase) {
          if (unsigned((*f.wait_phase >> 15) & 1) == f.wait_parity) continue;
          f.wait_phase = nullptr;
        }
        cur = &f;
        g_threadIdx = make_uint3(unsigned(f.tid), 0, 0);
        g_blockIdx = ctas[f.cta].block_idx;
        swapcontext(&sched_ctx, &f.ctx);
        if (f.done) {
          --cluster_live; --ctas[f.cta].live; --ctas[f.cta].warp_live[f.warp]; ++progress;
        }
        release_ready_barriers();
      }
      if (progress == before) {
        if (++idle_rounds > 3) {
          fprintf(stderr, "cuda_emu: DEADLOCK -- every live thread of the cluster is waiting (block %u); waiting threads:",
                  ctas[0].block_idx.x);
          int shown = 0;
          for (int i = 0; i < cluster * nthreads && shown < 16; ++i)
            if (!fibers[i].done) { fprintf(stderr, " cta%d/t%d", fibers[i].cta, fibers[i].tid); ++shown; }
          fprintf(stderr, "\n");
          abort();
        }
      } else {
        idle_rounds = 0;
      }
    }
  }
  run_deferred();
  while (!late_ops.empty()) { late_ops.front().op(); late_ops.pop_front(); }   // nothing may leak into the next launch
  in_coop = false;
  cur = nullptr;
}

}  // namespace eb_emu

// ---- symbols the emulated translation units reference but that live in files that are not emulated -------------
#include "../../egotap_b200/csrc/host_util.cuh"
namespace eb {
bool& prof_on() { static bool off = false; return off; }
void prof_push(const ProfRec&) {}
}  // namespace eb
extern "C" void emu_set_num_sms(int n) { eb_emu::g_num_sms = n; }
extern "C" long long emu_set_dry_run(int on) {
  const long long n = eb_emu::dry_launches;
  eb_emu::dry_run = on;
  eb_emu::dry_launches = 0;
  return n;          // launches that were checked (not executed) since the last call
}
extern "C" void emu_set_tma_latency(int max_rounds, unsigned long long seed) {
  eb_emu::tma_latency_max = max_rounds;
  eb_emu::tma_rng = seed ? seed : 88172645463325252ull;
}
extern "C" void emu_set_schedule(int mode, unsigned long long seed) {
  eb_emu::sched_mode = mode;
  eb_emu::sched_rng = seed ? seed : 0x9E3779B97F4A7C15ull;
}
extern "C" const char* egotap_b200_last_error(void) { return eb::err_buf(); }
extern "C" long long egotap_b200_launch_count(void) { return eb::launch_counter().load(); }


// TEST INFRASTRUCTURE ONLY -- runtime of the CUDA-on-CPU emulation (see cuda_emu.h) plus the few symbols the emulated
// translation units expect from the rest of libegotap_b200.so.
#include "cuda_emu.h"

#include <ucontext.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace eb_emu {

uint3 g_threadIdx, g_blockIdx;
dim3 g_blockDim, g_gridDim;

namespace {
constexpr size_t kStack = 256 * 1024;
struct Fiber {
  ucontext_t ctx;
  std::vector<char> stack;
  bool done = false;
  int warp = 0, lane = 0;
};
std::vector<Fiber> fibers;
ucontext_t sched_ctx;
Fiber* cur = nullptr;
bool in_coop = false;
const std::function<void()>* cur_body = nullptr;
// barrier state of the running CTA
int live = 0, block_waiting = 0;
unsigned block_gen = 0;
int warp_live[64], warp_waiting[64];
unsigned warp_gen[64];
float warp_buf[64][32];

void yield() { swapcontext(&cur->ctx, &sched_ctx); }

void trampoline() {
  (*cur_body)();
  cur->done = true;
  swapcontext(&cur->ctx, &sched_ctx);
}

void release_ready_barriers() {
  if (block_waiting > 0 && block_waiting == live) { block_waiting = 0; ++block_gen; }
  for (int w = 0; w < 64; ++w)
    if (warp_waiting[w] > 0 && warp_waiting[w] == warp_live[w]) { warp_waiting[w] = 0; ++warp_gen[w]; }
}

void warp_barrier() {
  const int w = cur->warp;
  const unsigned gen = warp_gen[w];
  ++warp_waiting[w];
  while (warp_gen[w] == gen) yield();
}
}  // namespace

void syncthreads() {
  if (!in_coop) { fprintf(stderr, "cuda_emu: __syncthreads() in a kernel launched with EB_LAUNCH\n"); abort(); }
  const unsigned gen = block_gen;
  ++block_waiting;
  while (block_gen == gen) yield();
}

float shfl_xor(float v, int lane_mask) {
  if (!in_coop) { fprintf(stderr, "cuda_emu: warp shuffle in a kernel launched with EB_LAUNCH\n"); abort(); }
  warp_buf[cur->warp][cur->lane] = v;
  warp_barrier();
  const float r = warp_buf[cur->warp][cur->lane ^ lane_mask];
  warp_barrier();
  return r;
}

void launch(dim3 grid, dim3 block, bool cooperative, const std::function<void()>& body) {
  g_blockDim = block;
  g_gridDim = grid;
  const int nthreads = int(block.x * block.y * block.z);
  if (block.y != 1 || block.z != 1 || nthreads > 2048) { fprintf(stderr, "cuda_emu: unsupported block shape\n"); abort(); }
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        g_blockIdx = make_uint3(bx, by, bz);
        if (!cooperative) {
          in_coop = false;
          for (int t = 0; t < nthreads; ++t) {
            g_threadIdx = make_uint3(unsigned(t), 0, 0);
            body();
          }
          continue;
        }
        in_coop = true;
        cur_body = &body;
        if (int(fibers.size()) < nthreads) fibers.resize(nthreads);
        live = nthreads;
        block_waiting = 0;
        for (int w = 0; w < 64; ++w) { warp_live[w] = 0; warp_waiting[w] = 0; }
        for (int t = 0; t < nthreads; ++t) {
          Fiber& f = fibers[t];
          if (f.stack.empty()) f.stack.resize(kStack);
          f.done = false;
          f.warp = t / 32;
          f.lane = t % 32;
          ++warp_live[f.warp];
          getcontext(&f.ctx);
          f.ctx.uc_stack.ss_sp = f.stack.data();
          f.ctx.uc_stack.ss_size = f.stack.size();
          f.ctx.uc_link = &sched_ctx;
          makecontext(&f.ctx, trampoline, 0);
        }
        long spins = 0;
        while (live > 0) {
          bool progressed = false;
          for (int t = 0; t < nthreads; ++t) {
            Fiber& f = fibers[t];
            if (f.done) continue;
            cur = &f;
            g_threadIdx = make_uint3(unsigned(t), 0, 0);
            const unsigned bg = block_gen, wg = warp_gen[f.warp];
            const int bw = block_waiting, ww = warp_waiting[f.warp];
            swapcontext(&sched_ctx, &f.ctx);
            if (f.done) { --live; --warp_live[f.warp]; progressed = true; }
            if (bg != block_gen || wg != warp_gen[f.warp] || bw != block_waiting || ww != warp_waiting[f.warp]) progressed = true;
            release_ready_barriers();
          }
          if (!progressed && ++spins > 4) { fprintf(stderr, "cuda_emu: deadlock (divergent barrier?)\n"); abort(); }
          if (progressed) spins = 0;
        }
        in_coop = false;
      }
}

}  // namespace eb_emu

// ---- symbols the emulated translation units reference but that live in files that are not emulated -------------
#include "../../egotap_b200/csrc/host_util.cuh"
namespace eb {
bool& prof_on() { static bool off = false; return off; }
void prof_push(const ProfRec&) {}
static int not_emulated(const char* what) { return fail(EGOTAP_E_UNSUPPORTED, "%s is not part of the CPU emulation", what); }
int split2d_run(const float*, long long, long long, long long, __nv_bfloat16*, __nv_bfloat16*, long long, cudaStream_t) { return not_emulated("split2d"); }
int fill_dummy_run(float*, const float*, int, int, int, cudaStream_t) { return not_emulated("fill_dummy"); }
int pos_permute_run(const float*, const float*, int, int, float*, float*, cudaStream_t) { return not_emulated("pos_permute"); }
int pu_bridge_gate_run(const float*, int, int, const float*, int, int, long long, __nv_bfloat16*, __nv_bfloat16*, cudaStream_t) { return not_emulated("pu_bridge_gate"); }
int vec_add3_run(const float*, const float*, const float*, float*, int, cudaStream_t) { return not_emulated("add3"); }
}  // namespace eb
extern "C" const char* egotap_b200_last_error(void) { return eb::err_buf(); }
extern "C" long long egotap_b200_launch_count(void) { return eb::launch_counter().load(); }

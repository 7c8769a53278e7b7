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
// declared locally in plan.cu (which holds the tensor-core plan and cannot be emulated)
int ingest_run(const float*, int, int, __nv_bfloat16*, __nv_bfloat16*, __nv_bfloat16*, __nv_bfloat16*, cudaStream_t);
int layernorm_run(const float*, const float*, const float*, long long, int, int, float, __nv_bfloat16*, __nv_bfloat16*,
                  float*, cudaStream_t);
int head_run(const float*, int, const float*, const float*, const float*, const float*, const float*, long long, int, int,
             int, float*, cudaStream_t);
}  // namespace eb
extern "C" const char* egotap_b200_last_error(void) { return eb::err_buf(); }
extern "C" long long egotap_b200_launch_count(void) { return eb::launch_counter().load(); }

// the op-level entries whose extern "C" wrappers live in plan.cu: same one-line forwarding as there
extern "C" int egotap_b200_ingest(const float* x, int frames, int preset, void* p_hi, void* p_lo, void* l_hi, void* l_lo,
                                  void* stream) {
  return eb::ingest_run(x, frames, preset == EGOTAP_PRESET_UNREALEGO ? 15 : 17, (__nv_bfloat16*)p_hi, (__nv_bfloat16*)p_lo,
                        (__nv_bfloat16*)l_hi, (__nv_bfloat16*)l_lo, (cudaStream_t)stream);
}
extern "C" int egotap_b200_layernorm(const float* x, const float* w, const float* b, long long frames, int rows_in,
                                     int rows_out, float eps, void* hi, void* lo, float* out_f32, void* stream) {
  return eb::layernorm_run(x, w, b, frames, rows_in, rows_out, eps, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, out_f32,
                           (cudaStream_t)stream);
}
extern "C" int egotap_b200_head(const float* e, int e_ld, const float* skel, const float* Wp, const float* bp,
                                const float* Wg, const float* bg, long long frames, int J, float* pose, void* stream) {
  return eb::head_run(e, e_ld, skel, Wp, bp, Wg, bg, frames, J, 256, 512, pose, (cudaStream_t)stream);
}

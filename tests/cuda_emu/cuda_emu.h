// TEST INFRASTRUCTURE ONLY -- a minimal CUDA-on-CPU execution model, so that the SOURCE of the product's kernels can be
// compiled by g++ (-DEB_HOST_EMU) and run in the GPU-less build container against the op oracle.
//
//   * bandwidth-bound kernels (csrc/kernels.cu, train_ops.cu, train_model.cu, gt_heatmap.cu): exact execution of the
//     kernel code -- indexing, layouts, reductions, arithmetic, alignment (-fsanitize=alignment)
//   * tensor-core kernels (csrc/gemm.cuh, attention.cu): a FUNCTIONAL model of the sm_100a features they use -- mbarrier
//     (arrival counts, transaction bytes, phase parity), TMA tiled loads with the 128-byte swizzle and zero fill,
//     tensor memory, tcgen05.mma (SS and TS forms, cta_group 1 and 2) decoding the real shared-memory / instruction
//     descriptors, tcgen05.ld / st, clusters -- see ptx_emu.h.  It checks protocol (no deadlock, every wait satisfied),
//     addressing and arithmetic of the kernel code.  Asynchronous operations (TMA loads, MMAs, commits) complete one
//     scheduler round after issue and a TMA destination is poisoned in between, so a consumer that does not wait computes
//     on NaN / stale data; proxy fences and memory-ordering subtleties are NOT modelled, and nothing about speed is.
//     The model was calibrated on kernels verified on the B200.
//
// Execution model: thread-block clusters run one after another; every thread of a cooperative launch is a fiber
// (ucontext), so __syncthreads(), named barriers, warp shuffles, __syncwarp() and mbarrier waits have their real meaning
// (threads that return early drop out of the barriers).  EB_LAUNCH kernels (independent threads) run as a plain loop.
// Static __shared__ variables become function-local statics (valid because such kernels run one CTA at a time);
// dynamic shared memory is a per-CTA 1024-byte aligned buffer.
#pragma once
#include <cuda.h>           // CUtensorMap (opaque 128-byte storage)
#include <cuda_runtime.h>   // host-side vector types, dim3, cudaStream_t, error codes
#include <cuda_bf16.h>      // __nv_bfloat16 (2-byte storage type; host-compilable)
#include <stdint.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>

namespace eb_emu {
extern uint3 g_threadIdx, g_blockIdx;
extern dim3 g_blockDim, g_gridDim;
void launch(dim3 grid, dim3 block, bool cooperative, const std::function<void()>& body);
void launch_ex(dim3 grid, dim3 block, int cluster, size_t dyn_smem, const std::function<void()>& body);
void syncthreads();
void syncwarp();
void named_barrier(int id, int nthreads);
void cluster_sync();
double shfl_xor(double v, int lane_mask);
inline float shfl_xor(float v, int lane_mask) { return float(shfl_xor(double(v), lane_mask)); }   // exact round trip
double shfl_idx(double v, int src_lane);
inline int shfl_idx(int v, int src_lane) { return int(shfl_idx(double(v), src_lane)); }
inline long long shfl_idx(long long v, int src_lane) {       // exact for |v| < 2^53 (row / element offsets)
  if (v > (1ll << 53) || v < -(1ll << 53)) { fprintf(stderr, "cuda_emu: 64-bit shuffle value out of the exact range\n"); abort(); }
  return (long long)shfl_idx(double(v), src_lane);
}
bool any_sync(bool pred);
void wait_phase(const void* mbar_first_word, unsigned parity);   // block until the mbarrier phase bit != parity
void yield_wait();              // a spinning wait gives the other fibers a turn (and feeds the deadlock detector)
void defer(std::function<void()> op);   // an asynchronous operation: runs at the start of the next scheduler round
void defer_tma(std::function<void()> op);   // a TMA load: as defer(), or a random number of rounds later (emu_set_tma_latency)
void note_progress();           // any state change another fiber may be waiting for
uint8_t* dyn_smem();            // dynamic shared memory of the running CTA
uint8_t* smem_of(int cta_rank); // ... of a CTA of the running cluster
uint32_t* tmem_of(int cta_rank);// tensor memory (128 lanes x 512 columns of 32 bits) of a CTA of the running cluster
uint32_t& tmem_next_col();      // allocation cursor of the running CTA
int cta_rank();
int lane();
int num_sms();
}  // namespace eb_emu

#define threadIdx eb_emu::g_threadIdx
#define blockIdx eb_emu::g_blockIdx
#define blockDim eb_emu::g_blockDim
#define gridDim eb_emu::g_gridDim
#undef __shared__
#define __shared__ static
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __grid_constant__
#define __grid_constant__
#define __syncthreads() eb_emu::syncthreads()
#define __syncwarp() eb_emu::syncwarp()
#define __shfl_xor_sync(mask, v, o) eb_emu::shfl_xor((v), (o))
#define __shfl_sync(mask, v, src) eb_emu::shfl_idx((v), (src))
#define __any_sync(mask, p) eb_emu::any_sync(p)
#define __trap() (fprintf(stderr, "cuda_emu: __trap() at %s:%d\n", __FILE__, __LINE__), abort())
#define __ldg(p) (*(p))
#define __expf(x) expf(x)
inline float atomicAdd(float* p, float v) { const float old = *p; *p = old + v; return old; }
inline unsigned atomicAdd(unsigned* p, unsigned v) { const unsigned old = *p; *p = old + v; eb_emu::note_progress(); return old; }
#define __threadfence() ((void)0)
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
#define clock64() (0LL)
inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
#define cudaMemsetAsync(p, v, n, s) (memset((p), (v), (n)), cudaSuccess)
#define cudaMemcpyAsync(d, s, n, kind, st) (memcpy((d), (s), (n)), cudaSuccess)

// TEST INFRASTRUCTURE ONLY -- a minimal CUDA-on-CPU execution model, so that the SOURCE of the bandwidth-bound
// training kernels (egotap_b200/csrc/train_ops.cu, train_model.cu) can be compiled by g++ (-DEB_HOST_EMU) and run in
// the GPU-less build container against the op oracle.  It checks indexing, layouts, reductions and arithmetic of the
// kernel code itself; it says nothing about performance, memory coalescing or hardware-specific behaviour, and the
// tensor-core kernels (tcgen05 / TMA / mbarrier) are out of its reach -- those were verified on the B200 in round 1.
//
// Execution model: CTAs run one after another.  Kernels launched with EB_LAUNCH have independent threads and run as a
// plain loop; EB_LAUNCH_COOP kernels run every thread of the CTA as a fiber (ucontext) so __syncthreads() and
// __shfl_xor_sync() have their real meaning (threads that return early drop out of the barriers, as on the GPU).
// __shared__ variables become function-local statics (one CTA at a time, one OS thread).
#pragma once
#include <cuda_runtime.h>   // host-side vector types, dim3, cudaStream_t, error codes
#include <stdint.h>
#include <cmath>
#include <cstring>
#include <functional>

#include <cuda_bf16.h>   // __nv_bfloat16 (2-byte storage type; host-compilable)

namespace eb_emu {
extern uint3 g_threadIdx, g_blockIdx;
extern dim3 g_blockDim, g_gridDim;
void launch(dim3 grid, dim3 block, bool cooperative, const std::function<void()>& body);
void syncthreads();
float shfl_xor(float v, int lane_mask);
}  // namespace eb_emu

#define threadIdx eb_emu::g_threadIdx
#define blockIdx eb_emu::g_blockIdx
#define blockDim eb_emu::g_blockDim
#define gridDim eb_emu::g_gridDim
#undef __shared__
#define __shared__ static
#undef __launch_bounds__
#define __launch_bounds__(...)
#define __syncthreads() eb_emu::syncthreads()
#define __shfl_xor_sync(mask, v, o) eb_emu::shfl_xor((v), (o))
#define __ldg(p) (*(p))
#define __expf(x) expf(x)
inline float atomicAdd(float* p, float v) { const float old = *p; *p = old + v; return old; }
inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
#define cudaMemsetAsync(p, v, n, s) (memset((p), (v), (n)), cudaSuccess)
#define cudaMemcpyAsync(d, s, n, kind, st) (memcpy((d), (s), (n)), cudaSuccess)

// Host-side helpers: error plumbing, TMA descriptor creation, launch accounting.
#pragma once
#ifdef EB_HOST_EMU
#include "cuda_emu.h"
#endif
#include <cuda.h>
#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "../../include/egotap_b200.h"

namespace eb {

inline char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}
inline std::atomic<long long>& launch_counter() {
  static std::atomic<long long> c{0};
  return c;
}

#define EB_CUDA(expr)                                                                           \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return eb::fail(int(_e), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define EB_CHECK_LAUNCH(name)                                                                   \
  do {                                                                                          \
    cudaError_t _e = cudaGetLastError();                                                        \
    if (_e != cudaSuccess)                                                                      \
      return eb::fail(int(_e), "launch of %s failed: %s", name, cudaGetErrorString(_e));        \
    eb::launch_counter().fetch_add(1, std::memory_order_relaxed);                               \
  } while (0)

// Kernel launch.  EB_LAUNCH_COOP marks kernels whose threads communicate inside a CTA (__syncthreads, warp shuffles,
// shared-memory atomics); EB_LAUNCH kernels have fully independent threads.  Both are the plain <<<>>> launch here --
// the distinction only matters to the CUDA-on-CPU emulation the tests build with -DEB_HOST_EMU (tests/cuda_emu), which
// runs independent-thread kernels as a loop and cooperative ones on fibers.
// EB_LAUNCH_CLUSTER: kernels with dynamic shared memory and (optionally) a thread-block cluster; EB_SET_MAX_SMEM: the
// per-device opt-in to > 48 KB of dynamic shared memory
#ifdef EB_HOST_EMU
#define EB_SET_MAX_SMEM(kernel, bytes) cudaSuccess
#define EB_LAUNCH_CLUSTER(kernel, grid, block, smem, cluster, stream, ...) \
  (eb_emu::launch_ex(dim3(grid), dim3(block), (cluster), (smem), [=]() { kernel(__VA_ARGS__); }), cudaSuccess)
#define EB_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...) \
  eb_emu::launch_ex(dim3(grid), dim3(block), 1, (smem), [=]() { kernel(__VA_ARGS__); })
#define EB_LAUNCH_GRID_SYNC(kernel, grid, block, smem, stream, ...) \
  eb_emu::launch_ex(dim3(grid), dim3(block), int(grid), (smem), [=]() { kernel(__VA_ARGS__); })
#else
#define EB_SET_MAX_SMEM(kernel, bytes) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)
#define EB_LAUNCH_CLUSTER(kernel, grid, block, smem, cluster, stream, ...) \
  eb::launch_cluster(kernel, grid, block, smem, cluster, stream, __VA_ARGS__)
#define EB_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...) kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
// kernels whose CTAs synchronise through global memory: every CTA of the grid must be co-resident.  A cooperative launch is
// all-or-nothing for residency (the driver rejects a grid that cannot be co-resident and never schedules part of it next to
// another stream's kernel), so two such kernels on different streams cannot dead-lock each other; the emulation runs the whole
// grid concurrently
#define EB_LAUNCH_GRID_SYNC(kernel, grid, block, smem, stream, ...) \
  eb::launch_cooperative(kernel, grid, block, smem, stream, __VA_ARGS__)
#endif

#ifdef EB_HOST_EMU
#define EB_LAUNCH(kernel, grid, block, stream, ...) \
  eb_emu::launch(dim3(grid), dim3(block), false, [=]() { kernel(__VA_ARGS__); })
#define EB_LAUNCH_COOP(kernel, grid, block, stream, ...) \
  eb_emu::launch(dim3(grid), dim3(block), true, [=]() { kernel(__VA_ARGS__); })
#undef EB_CHECK_LAUNCH
#define EB_CHECK_LAUNCH(name) eb::launch_counter().fetch_add(1, std::memory_order_relaxed)
#else
#define EB_LAUNCH(kernel, grid, block, stream, ...) kernel<<<grid, block, 0, stream>>>(__VA_ARGS__)
#define EB_LAUNCH_COOP(kernel, grid, block, stream, ...) kernel<<<grid, block, 0, stream>>>(__VA_ARGS__)
#endif

#define EB_REQUIRE(cond, ...)                                   \
  do {                                                          \
    if (!(cond)) return eb::fail(EGOTAP_E_ARG, __VA_ARGS__);    \
  } while (0)

// Optional per-launch CUDA-event timing (egotap_b200_profile_*): a ProfScope around a launch records start/end
// events on the launching stream while profiling is on, and costs one branch otherwise.
struct ProfRec {
  cudaEvent_t e0, e1;
  const char* name;
  int M, N, K, groups, variant;
};
bool& prof_on();
void prof_push(const ProfRec& r);
struct ProfScope {
  ProfRec r;
  cudaStream_t st;
  bool on;
#ifdef EB_HOST_EMU
  ProfScope(const char*, cudaStream_t stream, int = 0, int = 0, int = 0, int = 0, int = -1) : st(stream), on(false) {}
#else
  ProfScope(const char* name, cudaStream_t stream, int M = 0, int N = 0, int K = 0, int groups = 0, int variant = -1)
      : st(stream), on(prof_on()) {
    if (!on) return;
    r = ProfRec{nullptr, nullptr, name, M, N, K, groups, variant};
    cudaEventCreate(&r.e0);
    cudaEventCreate(&r.e1);
    cudaEventRecord(r.e0, st);
  }
#endif
  ~ProfScope() {
    if (!on) return;
    cudaEventRecord(r.e1, st);
    prof_push(r);
  }
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

#ifdef EB_HOST_EMU
// emulation: the same arguments and checks, recorded in the opaque CUtensorMap storage for tests/cuda_emu/ptx_emu.h
inline int make_operand_tmap(CUtensorMap* tm, const void* base, long long K, long long rows, long long ld,
                             long long g0_count, long long g0_stride, long long g1_count, long long g1_stride,
                             int box_rows);
#else
// 4-D bf16 tensor map [g1][g0][rows][K] (K contiguous), box = 64 x box_rows x 1 x 1, 128-byte swizzle.
inline int make_operand_tmap(CUtensorMap* tm, const void* base, long long K, long long rows, long long ld,
                             long long g0_count, long long g0_stride, long long g1_count, long long g1_stride,
                             int box_rows) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail(EGOTAP_E_DRIVER, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  if (g0_count < 1) g0_count = 1;
  if (g1_count < 1) g1_count = 1;
  if (g0_stride <= 0) g0_stride = rows * ld;
  if (g1_stride <= 0) g1_stride = g0_stride * g0_count;
  cuuint64_t dims[4] = {cuuint64_t(K), cuuint64_t(rows), cuuint64_t(g0_count), cuuint64_t(g1_count)};
  cuuint64_t strides[3] = {cuuint64_t(ld) * 2, cuuint64_t(g0_stride) * 2, cuuint64_t(g1_stride) * 2};
  cuuint32_t box[4] = {64, cuuint32_t(box_rows), 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (strides[0] & 15) || (strides[1] & 15) || (strides[2] & 15))
    return fail(EGOTAP_E_ARG, "TMA operand must be 16-byte aligned (base %p ld %lld)", base, ld);
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(EGOTAP_E_DRIVER, "cuTensorMapEncodeTiled failed (%d): K %lld rows %lld ld %lld box %d", int(r), K,
                rows, ld, box_rows);
  return 0;
}

#endif  // EB_HOST_EMU (make_operand_tmap)

#ifdef EB_HOST_EMU
inline int current_device() { return 0; }
inline int num_sms() { return eb_emu::num_sms(); }
#else
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < 64) ? dev : 0;
}

inline int num_sms() {
  static int n[64] = {0};
  const int dev = current_device();
  if (!n[dev]) cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
  return n[dev];
}
#endif

#ifndef EB_HOST_EMU
// cudaLaunchKernelEx with a cluster dimension attribute (cluster = 1: a plain launch with dynamic shared memory)
template <class Kernel, class... Args>
inline cudaError_t launch_cluster(Kernel kernel, unsigned grid, unsigned block, size_t smem, int cluster, cudaStream_t stream,
                                  Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = unsigned(cluster);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

template <class Kernel, class... Args>
inline cudaError_t launch_cooperative(Kernel kernel, unsigned grid, unsigned block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}
#endif

}  // namespace eb

#ifdef EB_HOST_EMU
#include "ptx_emu.h"
namespace eb {
inline int make_operand_tmap(CUtensorMap* tm, const void* base, long long K, long long rows, long long ld,
                             long long g0_count, long long g0_stride, long long g1_count, long long g1_stride,
                             int box_rows) {
  if (g0_count < 1) g0_count = 1;
  if (g1_count < 1) g1_count = 1;
  if (g0_stride <= 0) g0_stride = rows * ld;
  if (g1_stride <= 0) g1_stride = g0_stride * g0_count;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld * 2) & 15) || ((g0_stride * 2) & 15) || ((g1_stride * 2) & 15))
    return fail(EGOTAP_E_ARG, "TMA operand must be 16-byte aligned (base %p ld %lld)", base, ld);
  // the driver's limits: box <= 256 per dimension, every global dimension <= 2^32, every stride < 2^40 bytes
  const long long dmax = 1ll << 32, smax = 1ll << 40;
  if (K < 1 || rows < 1 || box_rows < 1 || box_rows > 256 || K > dmax || rows > dmax || g0_count > dmax || g1_count > dmax ||
      ld * 2 >= smax || g0_stride * 2 >= smax || g1_stride * 2 >= smax)
    return fail(EGOTAP_E_DRIVER, "cuTensorMapEncodeTiled (emulated) failed: K %lld rows %lld ld %lld box %d", K, rows, ld, box_rows);
  memset(tm, 0, sizeof(*tm));
  EmuTmap* e = reinterpret_cast<EmuTmap*>(tm);
  e->base = static_cast<const uint8_t*>(base);
  e->dims[0] = uint64_t(K); e->dims[1] = uint64_t(rows); e->dims[2] = uint64_t(g0_count); e->dims[3] = uint64_t(g1_count);
  e->strides[0] = uint64_t(ld) * 2; e->strides[1] = uint64_t(g0_stride) * 2; e->strides[2] = uint64_t(g1_stride) * 2;
  e->box_rows = uint32_t(box_rows);
  e->magic = 0x7e4a0001u;
  return 0;
}
}  // namespace eb
#endif

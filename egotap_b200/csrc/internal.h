// Internal (non-ABI) interfaces shared by the translation units of libegotap_b200.so.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace eb {

struct EpiParams;
struct GemmShape;

struct GemmOperand {
  const __nv_bfloat16* hi;
  const __nv_bfloat16* lo;
  long long ld;
  long long rows;
  long long g0_count, g0_stride, g1_count, g1_stride;
};

int pick_variant(int M, int N, int groups, int nsplit);
int gemm_run(const GemmOperand& a, const GemmOperand& b, const GemmShape& s, const EpiParams& ep, int nsplit,
             int variant, cudaStream_t stream);

// attention.cu
int attention_run(const __nv_bfloat16* qk_hi, const __nv_bfloat16* qk_lo, const __nv_bfloat16* vt_hi,
                  const __nv_bfloat16* vt_lo, __nv_bfloat16* ctx_hi, __nv_bfloat16* ctx_lo, int B, int nsplit,
                  int query_rows, cudaStream_t stream, float* lse = nullptr);
// attention_bwd.cu (bf16-operand training mode)
int attention_bwd_run(const __nv_bfloat16* qk, const __nv_bfloat16* vt, const __nv_bfloat16* dctx, const float* lse,
                      const float* dsum, float* dqkv, int B, cudaStream_t stream);
int attn_dsum_run(const __nv_bfloat16* o_hi, const __nv_bfloat16* o_lo, const __nv_bfloat16* do_hi, const __nv_bfloat16* do_lo,
                  long long rows, float* dsum, cudaStream_t stream);

// pu_chain.cu
int pu_permute_split_run(const float* W, __nv_bfloat16* hi, __nv_bfloat16* lo, cudaStream_t stream);
int pu_chain_run(const __nv_bfloat16* w_hi, const __nv_bfloat16* w_lo, const float* G, long long G_rs, long long G_ts,
                 const float* F, long long F_rs, long long F_ts, float* out, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo,
                 __nv_bfloat16* hg_hi, __nv_bfloat16* hg_lo, unsigned int* counters, int B, int J, int nsplit,
                 cudaStream_t stream);

// kernels.cu
int split_bf16_run(const float* src, __nv_bfloat16* hi, __nv_bfloat16* lo, long long n, cudaStream_t stream);

}  // namespace eb

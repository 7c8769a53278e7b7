// On-GPU evaluation metrics: per-frame MPJPE and Procrustes-aligned MPJPE in mm.
//   reference utils/loss.py:79-85 (MPJPE = mean joint L2 distance), utils/util.py:328-379
//   (batch_compute_similarity_transform_torch: similarity Procrustes via the SVD of the 3x3 cross-covariance),
//   utils/evaluate.py:54-73 and model/egotap_autoencoder_model.py:329-350 (per-frame loop, cm -> mm x10).
// One warp per frame: lane j holds joint j; means, variance and the 3x3 cross-covariance are warp-shuffle
// reductions; the 3x3 SVD (one-sided Jacobi, fp64, in registers) replaces the batched torch.svd + per-frame Python
// loop with its device->host synchronisations.
#include "host_util.cuh"
#include "internal.h"

namespace eb {

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// K = U diag(S) V^T by one-sided Jacobi (Hestenes): rotate column pairs of A = K until orthogonal.
__device__ void svd3x3(const double K[3][3], double U[3][3], double S[3], double V[3][3]) {
  double A[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) { A[i][j] = K[i][j]; V[i][j] = (i == j) ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 12; ++sweep) {
    double off = 0.0;
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
      double alpha = 0, beta = 0, gamma = 0;
#pragma unroll
      for (int i = 0; i < 3; ++i) { alpha += A[i][p] * A[i][p]; beta += A[i][q] * A[i][q]; gamma += A[i][p] * A[i][q]; }
      off = fmax(off, fabs(gamma) / (sqrt(alpha * beta) + 1e-300));
      if (fabs(gamma) > 1e-18 * sqrt(alpha * beta)) {
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const double ap = A[i][p], aq = A[i][q];
          A[i][p] = c * ap - s * aq; A[i][q] = s * ap + c * aq;
          const double vp = V[i][p], vq = V[i][q];
          V[i][p] = c * vp - s * vq; V[i][q] = s * vp + c * vq;
        }
      }
    }
    if (off < 1e-15) break;
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    S[j] = sqrt(A[0][j] * A[0][j] + A[1][j] * A[1][j] + A[2][j] * A[2][j]);
    const double inv = S[j] > 1e-300 ? 1.0 / S[j] : 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) U[i][j] = A[i][j] * inv;
  }
}

__device__ __forceinline__ double det3(const double M[3][3]) {
  return M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
         M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
}

__global__ void __launch_bounds__(256) pose_metrics_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                           long long frames, int nj, float unit_scale,
                                                           float* __restrict__ mpjpe, float* __restrict__ pa_mpjpe) {
  const int lane = threadIdx.x & 31;
  const long long f = blockIdx.x * 8ll + (threadIdx.x >> 5);
  if (f >= frames) return;
  const bool on = lane < nj;
  double p[3] = {0, 0, 0}, g[3] = {0, 0, 0};
  if (on) {
#pragma unroll
    for (int a = 0; a < 3; ++a) { p[a] = pred[(f * nj + lane) * 3 + a]; g[a] = gt[(f * nj + lane) * 3 + a]; }
  }
  const double inv_n = 1.0 / nj;
  // MPJPE
  const double d0 = sqrt((p[0] - g[0]) * (p[0] - g[0]) + (p[1] - g[1]) * (p[1] - g[1]) + (p[2] - g[2]) * (p[2] - g[2]));
  const double m = warp_sum_d(on ? d0 : 0.0) * inv_n;
  // Procrustes: S1 = pred, S2 = gt
  double mu1[3], mu2[3], x1[3], x2[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    mu1[a] = warp_sum_d(p[a]) * inv_n; mu2[a] = warp_sum_d(g[a]) * inv_n;
    x1[a] = on ? p[a] - mu1[a] : 0.0; x2[a] = on ? g[a] - mu2[a] : 0.0;
  }
  const double var1 = warp_sum_d(x1[0] * x1[0] + x1[1] * x1[1] + x1[2] * x1[2]);
  double K[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) K[a][b] = warp_sum_d(x1[a] * x2[b]);
  double U[3][3], S[3], V[3][3];
  svd3x3(K, U, S, V);
  // R = V Z U^T with Z flipping the direction of the SMALLEST singular value when det(U V^T) < 0
  const double sgn = (det3(U) * det3(V)) < 0.0 ? -1.0 : 1.0;
  int jmin = 0;
  if (S[1] < S[jmin]) jmin = 1;
  if (S[2] < S[jmin]) jmin = 2;
  double R[3][3], tr = 0.0;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      double r = 0.0;
#pragma unroll
      for (int j = 0; j < 3; ++j) r += (j == jmin ? sgn : 1.0) * V[a][j] * U[b][j];
      R[a][b] = r;
    }
#pragma unroll
  for (int j = 0; j < 3; ++j) tr += (j == jmin ? sgn : 1.0) * S[j];
  const double scale = tr / var1;
  double e2 = 0.0;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const double h = scale * (R[a][0] * x1[0] + R[a][1] * x1[1] + R[a][2] * x1[2]) + mu2[a] - g[a];
    e2 += h * h;
  }
  const double pa = warp_sum_d(on ? sqrt(e2) : 0.0) * inv_n;
  if (lane == 0) {
    mpjpe[f] = float(m * unit_scale);
    pa_mpjpe[f] = float(pa * unit_scale);
  }
}

int pose_metrics_run(const float* pred, const float* gt, long long frames, int nj, float unit_scale, float* mpjpe,
                     float* pa_mpjpe, cudaStream_t stream) {
  EB_REQUIRE(pred && gt && mpjpe && pa_mpjpe, "pose_metrics: null pointer");
  EB_REQUIRE(nj >= 3 && nj <= 32, "pose_metrics: joints must be in [3, 32], got %d", nj);
  if (frames == 0) return 0;
  ProfScope prof("pose_metrics_kernel", stream);
  EB_LAUNCH_COOP(pose_metrics_kernel, (unsigned)((frames + 7) / 8), 256, stream, pred, gt, frames, nj, unit_scale, mpjpe, pa_mpjpe);
  EB_CHECK_LAUNCH("pose_metrics_kernel");
  return 0;
}

}  // namespace eb

extern "C" int egotap_b200_pose_metrics(const float* pred, const float* gt, long long frames, int joints, float unit_scale,
                                        float* mpjpe, float* pa_mpjpe, void* stream) {
  return eb::pose_metrics_run(pred, gt, frames, joints, unit_scale, mpjpe, pa_mpjpe, (cudaStream_t)stream);
}

// Ground-truth heatmap synthesis on the GPU: 2-D / 3-D keypoints -> the lifting
// network's (B, 6J, 64, 64) input [joint L | joint R | cos L | sin L | cos R | sin R], replacing the reference's per-frame
// numpy / scipy / skimage pipeline under --use_gt_heatmap (utils/projection.py:263-279, utils/data.py:175-252,
// dataloader/data_loader.py:127-132,193-199, model/egotap_autoencoder_model.py:176-213) and the 1.47 MB/frame host->device
// copy of its result: only ~0.5 KB of keypoints per frame crosses PCIe.
// One CTA per output heatmap (joint channels) or per limb (its cos and sin channels); HBM-bound on the 1.47 MB/frame
// written (coalesced float4 stores).  Semantics: oracle/gt_heatmap_oracle.py (test infrastructure).
#include "host_util.cuh"
#include "numeric.cuh"

namespace eb {
namespace {

struct GtCfg {
  int J, n;               // heatmaps per view, keypoints per view (J + 1)
  int parents[24];
  float w[5];             // normalised 9-tap Gaussian (sigma 1, truncate 4), w[|k|]
};
constexpr int R = 64;

}  // namespace

__global__ void __launch_bounds__(256) gt_heatmap_kernel(const float* __restrict__ pts2d, const float* __restrict__ pts3d_left,
                                                         GtCfg cfg, float* __restrict__ out) {
  __shared__ float tile[R][R + 1];
  __shared__ float tmp[R][R + 1];
  __shared__ float s_cs[2];
  const int J = cfg.J, n = cfg.n;
  const long long b = blockIdx.x / (4 * J);
  const int k = blockIdx.x % (4 * J);
  const int tid = threadIdx.x;
  float* frame_out = out + b * 6 * J * (long long)(R * R);
  if (k < 2 * J) {
    // ---- joint heatmap: separable Gaussian around the truncated pixel position, / (1 / 2 pi)
    const int view = k / J, j = k % J;
    const float* p = pts2d + ((b * 2 + view) * n + (j + 1)) * 2;
    const float x = p[0] / 1024.0f * float(R), y = p[1] / 1024.0f * float(R);
    const bool on = (-4.0f <= y) && (y < float(R + 4)) && (-4.0f <= x) && (x < float(R));      // (sic) no +4 on x
    const int ix = int(x), iy = int(y);                                                           // truncation toward zero
    float4* o = reinterpret_cast<float4*>(frame_out + (long long)k * (R * R));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx4 = tid + i * 256;
      const int r = idx4 >> 4, c0 = (idx4 & 15) * 4;
      float v[4];
      const int dr = r - iy < 0 ? iy - r : r - iy;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int dc = c0 + q - ix < 0 ? ix - c0 - q : c0 + q - ix;
        v[q] = (on && dr <= 4 && dc <= 4) ? (cfg.w[dr] * cfg.w[dc]) / 0.15915589174187972f : 0.0f;
      }
      o[idx4] = make_float4(v[0], v[1], v[2], v[3]);
    }
    return;
  }
  // ---- limb heatmap: anti-aliased line between the rounded endpoints, 9-tap blur (zero boundary), x2, cos / sin
  const int kk = k - 2 * J, view = kk / J, limb = kk % J;
  const int joint = limb + 1, parent = cfg.parents[joint];
  for (int i = tid; i < R * R; i += 256) tile[i / R][i % R] = 0.0f;
  __syncthreads();
  if (tid == 0) {
    const float* pp = pts2d + ((b * 2 + view) * n + parent) * 2;
    const float* pc = pts2d + ((b * 2 + view) * n + joint) * 2;
    // skimage.draw.line_aa(r0, c0, r1, c1) with (r, c) = (x, y) as the reference calls it (utils/data.py:181-183)
    const int r0 = int(rintf(pp[0] / 16.0f)), c0 = int(rintf(pp[1] / 16.0f));
    const int r1 = int(rintf(pc[0] / 16.0f)), c1 = int(rintf(pc[1] / 16.0f));
    const int dc = c0 < c1 ? c1 - c0 : c0 - c1, dr = r0 < r1 ? r1 - r0 : r0 - r1;
    const int sign_c = c0 < c1 ? 1 : -1, sign_r = r0 < r1 ? 1 : -1;
    float err = float(dc - dr);
    const float ed = (dc + dr == 0) ? 1.0f : sqrtf(float(dc) * float(dc) + float(dr) * float(dr));
    int c = c0, r = r0;
#define EB_PLOT(xx, yy, vv)                                                        \
    do {                                                                           \
      if ((xx) >= 0 && (xx) < R && (yy) >= 0 && (yy) < R) tile[(yy)][(xx)] = 1.0f - (vv); \
    } while (0)
    for (int guard = 0; guard < (1 << 20); ++guard) {     // bounded: garbage keypoints must not hang the device
      EB_PLOT(r, c, fabsf(err - float(dc) + float(dr)) / ed);
      const float err_prime = err;
      const int c_prime = c;
      if (2.0f * err_prime >= -float(dc)) {
        if (c == c1) break;
        if (err_prime + float(dr) < ed) EB_PLOT(r + sign_r, c, fabsf(err_prime + float(dr)) / ed);
        err -= float(dr);
        c += sign_c;
      }
      if (2.0f * err_prime <= float(dr)) {
        if (r == r1) break;
        if (float(dc) - err_prime < ed) EB_PLOT(r, c_prime + sign_c, fabsf(float(dc) - err_prime) / ed);
        err += float(dc);
        r += sign_r;
      }
    }
#undef EB_PLOT
    // elevation angle of the limb from the LEFT view's 3-D points (utils/data.py:222-226, 254-262)
    const float* a = pts3d_left + (b * n + parent) * 3;
    const float* q = pts3d_left + (b * n + joint) * 3;
    const float lx = a[0] - q[0], ly = a[1] - q[1], lz = a[2] - q[2];
    const float theta = atanf(lz / sqrtf(lx * lx + ly * ly));
    s_cs[0] = cosf(theta);
    s_cs[1] = sinf(theta);
  }
  __syncthreads();
  for (int i = tid; i < R * R; i += 256) {       // axis 0 (y)
    const int yy = i / R, xx = i % R;
    float acc = 0.0f;
#pragma unroll
    for (int d = -4; d <= 4; ++d) {
      const int ys = yy + d;
      if (ys >= 0 && ys < R) acc += cfg.w[d < 0 ? -d : d] * tile[ys][xx];
    }
    tmp[yy][xx] = acc;
  }
  __syncthreads();
  float* o_cos = frame_out + (long long)(2 * J + view * 2 * J + limb) * (R * R);
  float* o_sin = o_cos + (long long)J * (R * R);
  const float cs = s_cs[0], sn = s_cs[1];
  for (int i = tid; i < R * R; i += 256) {       // axis 1 (x), then sigma (= 1) * 2 * cos / sin
    const int yy = i / R, xx = i % R;
    float acc = 0.0f;
#pragma unroll
    for (int d = -4; d <= 4; ++d) {
      const int xs = xx + d;
      if (xs >= 0 && xs < R) acc += cfg.w[d < 0 ? -d : d] * tmp[yy][xs];
    }
    const float raw = acc * 2.0f;
    o_cos[i] = raw * cs;
    o_sin[i] = raw * sn;
  }
}

int gt_heatmaps_run(const float* pts2d, const float* pts3d_left, long long B, int preset, float* out, cudaStream_t st) {
  EB_REQUIRE(pts2d && pts3d_left && out, "gt_heatmaps: null pointer");
  EB_REQUIRE(preset == EGOTAP_PRESET_UNREALEGO || preset == EGOTAP_PRESET_EGOCAP, "gt_heatmaps: unknown preset %d", preset);
  EB_REQUIRE(B > 0 && B * 68 < 2147483647ll, "gt_heatmaps: bad frame count %lld", B);
  static const int ue[16] = {0, 0, 1, 1, 2, 3, 4, 5, 2, 3, 8, 9, 10, 11, 12, 13};                 // reference utils/util.py:51
  static const int ec[18] = {0, 0, 1, 2, 3, 4, 1, 6, 7, 8, 2, 10, 11, 12, 6, 14, 15, 16};         // :52
  GtCfg cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.n = preset == EGOTAP_PRESET_UNREALEGO ? 16 : 18;
  cfg.J = cfg.n - 1;
  for (int i = 0; i < cfg.n; ++i) cfg.parents[i] = preset == EGOTAP_PRESET_UNREALEGO ? ue[i] : ec[i];
  double w[5], sum = 0.0;
  for (int k = 0; k < 5; ++k) { w[k] = exp(-0.5 * k * k); sum += (k == 0 ? 1.0 : 2.0) * w[k]; }
  for (int k = 0; k < 5; ++k) cfg.w[k] = float(w[k] / sum);
  ProfScope prof("gt_heatmap_kernel", st);
  EB_LAUNCH_COOP(gt_heatmap_kernel, (unsigned)(B * 4 * cfg.J), 256, st, pts2d, pts3d_left, cfg, out);
  EB_CHECK_LAUNCH("gt_heatmap_kernel");
  return 0;
}

}  // namespace eb

extern "C" int egotap_b200_gt_heatmaps(const float* pts2d, const float* pts3d_left, long long frames, int preset, float* out,
                                       void* stream) {
  return eb::gt_heatmaps_run(pts2d, pts3d_left, frames, preset, out, (cudaStream_t)stream);
}

// Backward of the alpha composite (reference models/rendering.py:116-143), the entry of
// the training step's gradient into the per-point MLP outputs.
//
// Forward (per ray, samples s = 0..S-1):
//   delta_s = z_{s+1} - z_s (last 1e2)            x_s = sigma_s + noise_s
//   alpha_s = 1 - exp(-delta_s relu(x_s))         T_s = prod_{j<s} (1 - alpha_j)
//   w_s = alpha_s T_s      feature = sum_s w_s f_s      depth = sum_s w_s z_s
// with f_s = sigmoid(rgb_pre_s) (64 channels) and sigma_s = softplus(sigma_pre_s).
//
// Given g_feature (N,64), g_weights (N,S), g_depth (N) (any may be NULL) this computes
//   gw_s        = <g_feature, f_s> + g_weights_s + g_depth z_s          (dL/dw_s)
//   R_s         = sum_{k>s} gw_k alpha_k prod_{s<j<k} (1 - alpha_j)      (R_{S-1} = 0,
//                 R_s = gw_{s+1} alpha_{s+1} + (1 - alpha_{s+1}) R_{s+1}: no division, exact
//                 even when a sample saturates to alpha = 1)
//   dL/dalpha_s = T_s (gw_s - R_s)
//   d_sigma_pre = dL/dalpha_s * delta_s (1 - alpha_s) [x_s > 0] * (1 - exp(-sigma_s))
//   d_rgb_pre   = w_s g_feature f_s (1 - f_s)
// z, noise and rays receive no gradient (the reference detaches the importance samples,
// rendering.py:184, and its inputs do not require grad).
//
// One warp per ray: lanes stride the 64 channels (coalesced rows of the (P,65) buffer) for the
// dot products and the output rows; the two length-S recurrences run on lane 0 over shared
// memory (S <= 1024; training uses 64 / 128).
#include "common.h"

namespace crnerf {
namespace {

constexpr int kMaxS = 1024;

__global__ void __launch_bounds__(128)
composite_backward_kernel(const float* __restrict__ raw, const float* __restrict__ z,
                          const float* __restrict__ noise, const float* __restrict__ g_feature,
                          const float* __restrict__ g_weights, const float* __restrict__ g_depth,
                          int n_rays, int S, float* __restrict__ d_rgb_pre,
                          float* __restrict__ d_sigma_pre) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ray = blockIdx.x * 4 + warp;
  if (ray >= n_rays) return;
  float* gw = sm + warp * 4 * S;  // dL/dw_s, later dL/dalpha_s
  float* al = gw + S;             // alpha_s
  float* wt = al + S;             // w_s
  float* dl = wt + S;             // delta_s (1 - alpha_s) [x_s > 0]
  const long long p0 = (long long)ray * S;
  const float g0 = g_feature ? g_feature[(long long)ray * 64 + lane] : 0.f;
  const float g1 = g_feature ? g_feature[(long long)ray * 64 + 32 + lane] : 0.f;
  const float gd = g_depth ? g_depth[ray] : 0.f;

  // pass 1 (parallel over samples in groups of one row per iteration): dL/dw_s and alpha_s
  for (int s = 0; s < S; ++s) {
    const float* row = raw + (p0 + s) * 65;
    float dot = g0 * row[lane] + g1 * row[32 + lane];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, d);
    if (lane == 0) {
      const float zs = z[p0 + s];
      const float delta = s + 1 < S ? __fsub_rn(z[p0 + s + 1], zs) : 1e2f;
      const float x = row[64] + (noise ? noise[p0 + s] : 0.f);
      const float e = expf(-(delta * fmaxf(x, 0.f)));  // 1 - alpha
      al[s] = 1.f - e;
      dl[s] = x > 0.f ? delta * e : 0.f;
      gw[s] = dot + (g_weights ? g_weights[p0 + s] : 0.f) + gd * zs;
    }
  }
  __syncwarp();
  // pass 2 (lane 0): T_s forward, R_s backward
  if (lane == 0) {
    float T = 1.f;
    for (int s = 0; s < S; ++s) {
      wt[s] = al[s] * T;
      const float t_next = T * (1.f - al[s]);
      dl[s] *= T;  // T_s * d alpha_s / d x_s
      T = t_next;
    }
    float R = 0.f;
    for (int s = S - 1; s >= 0; --s) {
      const float gws = gw[s];
      gw[s] = gws - R;  // (dL/dalpha_s) / T_s
      R = gws * al[s] + (1.f - al[s]) * R;
    }
  }
  __syncwarp();
  // pass 3: outputs
  for (int s = 0; s < S; ++s) {
    const float* row = raw + (p0 + s) * 65;
    const float w = wt[s];
    const float f0 = row[lane], f1 = row[32 + lane];
    float* o = d_rgb_pre + (p0 + s) * 64;
    o[lane] = w * g0 * f0 * (1.f - f0);
    o[32 + lane] = w * g1 * f1 * (1.f - f1);
    if (lane == 0) {
      const float sigma = row[64];
      d_sigma_pre[p0 + s] = gw[s] * dl[s] * (1.f - expf(-sigma));
    }
  }
}

}  // namespace

int composite_backward(const float* raw, const float* z, const float* noise, const float* g_feature,
                       const float* g_weights, const float* g_depth, int n_rays, int n_samples,
                       float* d_rgb_pre, float* d_sigma_pre, cudaStream_t st) {
  CRNERF_REQUIRE(raw && z && d_rgb_pre && d_sigma_pre, "null argument");
  CRNERF_REQUIRE(n_samples >= 1 && n_samples <= kMaxS, "n_samples=%d unsupported by the backward (<= %d)",
                 n_samples, kMaxS);
  if (n_rays == 0) return CRNERF_OK;
  const size_t smem = (size_t)4 * 4 * n_samples * sizeof(float);
  if (smem > 48 * 1024)
    CRNERF_CUDA(cudaFuncSetAttribute(composite_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
  composite_backward_kernel<<<(n_rays + 3) / 4, 128, smem, st>>>(raw, z, noise, g_feature, g_weights, g_depth,
                                                                 n_rays, n_samples, d_rgb_pre, d_sigma_pre);
  count_launch();
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

}  // namespace crnerf

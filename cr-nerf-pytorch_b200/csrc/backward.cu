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
// One block of four warps per ray, 32 samples per warp and pass.  Dot products: a warp's 32 rows at a time,
// lanes over the channels (coalesced 256-byte rows, 64 independent loads in flight per lane), then a
// transposing butterfly (31 shuffles for 32 rows) leaves row r's dot product on lane r, so everything per
// sample - z, noise, sigma, the exponentials - runs with lanes over samples.  Only the two length-S
// recurrences are serial (one thread over shared memory; S <= 1024, training uses 64 / 128).  For
// S <= 128 (kKeep) a warp's rows stay in registers between the passes: the features are read once.
#include <algorithm>
#include "common.h"

namespace crnerf {
namespace {

constexpr int kMaxS = 1024;
constexpr int kCbWarps = 4;   // warps per ray

template <bool kKeep>
__global__ void __launch_bounds__(32 * kCbWarps)
composite_backward_kernel(const float* __restrict__ raw, const float* __restrict__ z,
                          const float* __restrict__ noise, const float* __restrict__ g_feature,
                          const float* __restrict__ g_weights, const float* __restrict__ g_depth,
                          int n_rays, int S, int split, float* __restrict__ d_rgb_pre,
                          float* __restrict__ d_sigma_pre, float* __restrict__ amax) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ray = blockIdx.x;
  float* gw = sm;       // dL/dw_s, later (dL/dalpha_s) / T_s
  float* al = gw + S;   // alpha_s
  float* wt = al + S;   // w_s
  float* dl = wt + S;   // delta_s (1 - alpha_s) [x_s > 0], later times T_s
  float* sg = dl + S;   // 1 - exp(-sigma_s)
  const long long p0 = (long long)ray * S;
  // split: the training forward's layout ((P,64) feature rows, then P sigmas: aligned rows, whole-sector stores);
  // otherwise (P,65) interleaved rows as NeRF_sigma.forward returns them
  const int rs = split ? 64 : 65;
  const float* sig = split ? raw + (long long)n_rays * S * 64 : raw + 64;
  const int ss = split ? 1 : 65;
  const float g0 = g_feature ? g_feature[(long long)ray * 64 + lane] : 0.f;
  const float g1 = g_feature ? g_feature[(long long)ray * 64 + 32 + lane] : 0.f;
  const float gd = g_depth ? g_depth[ray] : 0.f;
  const int n_chunks = (S + 31) >> 5;

  // pass 1: dL/dw_s, alpha_s and the per-sample factors
  float f0[32], f1[32];
  for (int c = warp; c < n_chunks; c += kCbWarps) {
    const int s0 = c * 32;
    float p[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) {
      const int s = min(s0 + r, S - 1);  // rows past the end repeat the last one; their lanes write nothing
      const float* row = raw + (p0 + s) * rs;
      f0[r] = __ldg(row + lane);
      f1[r] = __ldg(row + 32 + lane);
    }
#pragma unroll
    for (int r = 0; r < 32; ++r) p[r] = g0 * f0[r] + g1 * f1[r];
    // after the step with distance d, bit log2(d) of the row a lane still carries equals that bit of the lane
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      const bool up = (lane & d) != 0;
#pragma unroll
      for (int i = 0; i < d; ++i) {
        const float keep = up ? p[i + d] : p[i], send = up ? p[i] : p[i + d];
        p[i] = keep + __shfl_xor_sync(0xffffffffu, send, d);
      }
    }
    const int s = s0 + lane;
    if (s < S) {
      const float zs = z[p0 + s];
      const float delta = s + 1 < S ? __fsub_rn(z[p0 + s + 1], zs) : 1e2f;
      const float sigma = __ldg(sig + (p0 + s) * ss);
      const float x = sigma + (noise ? noise[p0 + s] : 0.f);
      const float e = expf(-(delta * fmaxf(x, 0.f)));  // 1 - alpha
      al[s] = 1.f - e;
      dl[s] = x > 0.f ? delta * e : 0.f;
      sg[s] = 1.f - expf(-sigma);
      gw[s] = p[0] + (g_weights ? g_weights[p0 + s] : 0.f) + gd * zs;
    }
  }
  __syncthreads();
  // pass 2 (one thread): T_s forward, R_s backward
  if (threadIdx.x == 0) {
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
  __syncthreads();
  // pass 3: outputs
  float am = 0.f;   // max magnitude of everything this thread writes (the backward chain's first scale)
  for (int s = threadIdx.x; s < S; s += 32 * kCbWarps) {
    const float v = gw[s] * dl[s] * sg[s];
    d_sigma_pre[p0 + s] = v;
    am = fmaxf(am, fabsf(v));
  }
  for (int c = warp; c < n_chunks; c += kCbWarps) {
    const int s0 = c * 32;
#pragma unroll
    for (int r = 0; r < 32; ++r) {
      const int s = s0 + r;
      if (s < S) {
        float a0 = f0[r], a1 = f1[r];
        if (!kKeep) {   // more than one chunk per warp: the registers hold the last one only
          const float* row = raw + (p0 + s) * rs;
          a0 = __ldg(row + lane);
          a1 = __ldg(row + 32 + lane);
        }
        const float w = wt[s];
        float* o = d_rgb_pre + (p0 + s) * 64;
        const float o0 = w * g0 * a0 * (1.f - a0), o1 = w * g1 * a1 * (1.f - a1);
        o[lane] = o0;
        o[32 + lane] = o1;
        am = fmaxf(am, fmaxf(fabsf(o0), fabsf(o1)));
      }
    }
  }
  if (amax != nullptr) {   // non-negative floats order like their bit patterns; NaN / inf are left out
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, d));
    if (lane == 0 && am > 0.f && am < 3.0e38f) atomicMax(reinterpret_cast<unsigned int*>(amax), __float_as_uint(am));
  }
}

// g (P, C) fp32 *= (act > 0) in place (act: the layer's saved 16-bit post-ReLU output, or
// NULL for a layer without activation), and partial column sums for the bias gradient.
// One pass over g instead of the mask multiply + a strided column reduction.
// 256 threads: column = tid % C, row lane = tid / C (C in {64, 128, 256}).
__global__ void __launch_bounds__(256)
relu_bias_grad_kernel(float* __restrict__ g, const uint16_t* __restrict__ act, long long P, int C,
                      float* __restrict__ partial) {
  __shared__ float red[256];
  const int c = threadIdx.x % C, rl = threadIdx.x / C, rows_per_it = 256 / C;
  const long long rows_per_block = (P + gridDim.x - 1) / gridDim.x;
  const long long r0 = blockIdx.x * rows_per_block, r1 = min(P, r0 + rows_per_block);
  float acc = 0.f;
  // 8 independent rows in flight per thread: the kernel is a pure HBM stream and one
  // outstanding load per thread would leave it latency-bound
  constexpr int kU = 8;
  long long r = r0 + rl;
  for (; r + (long long)(kU - 1) * rows_per_it < r1; r += (long long)kU * rows_per_it) {
    float v[kU];
    uint16_t a[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) v[u] = g[(r + (long long)u * rows_per_it) * C + c];
    if (act) {
#pragma unroll
      for (int u = 0; u < kU; ++u) a[u] = act[(r + (long long)u * rows_per_it) * C + c];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        if (!(a[u] != 0 && a[u] < 0x8000)) v[u] = 0.f;  // post-ReLU value not strictly positive
        g[(r + (long long)u * rows_per_it) * C + c] = v[u];  // unconditional: whole-line writes
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) acc += v[u];
  }
  for (; r < r1; r += rows_per_it) {
    float v = g[r * C + c];
    if (act) {
      const uint16_t a = act[r * C + c];
      if (!(a != 0 && a < 0x8000)) {
        v = 0.f;
        g[r * C + c] = 0.f;
      }
    }
    acc += v;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < C) {
    float t = 0.f;
    for (int k = 0; k < rows_per_it; ++k) t += red[k * C + threadIdx.x];
    partial[(long long)blockIdx.x * C + threadIdx.x] = t;
  }
}

// gb[c] = sum_b partial[b][c]; one block per column, fixed summation tree (deterministic)
__global__ void __launch_bounds__(128)
bias_reduce_kernel(const float* __restrict__ partial, int n_parts, int C, float* __restrict__ gb) {
  __shared__ float red[4];
  const int c = blockIdx.x;
  float acc = 0.f;
  for (int b = threadIdx.x; b < n_parts; b += 128) acc += partial[(long long)b * C + c];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) gb[c] = (red[0] + red[1]) + (red[2] + red[3]);
}

}  // namespace

int relu_bias_grad(float* g, const void* act, int64_t n_points, int width, float* gb, float* scratch,
                   cudaStream_t st) {
  CRNERF_REQUIRE(g && gb && scratch, "null argument");
  CRNERF_REQUIRE(width == 64 || width == 128 || width == 256, "width must be 64, 128 or 256");
  if (n_points <= 0) {
    CRNERF_CUDA(cudaMemsetAsync(gb, 0, sizeof(float) * width, st));
    return CRNERF_OK;
  }
  const int nb = (int)std::min<long long>(8LL * num_sms(), (n_points + 63) / 64);
  relu_bias_grad_kernel<<<nb, 256, 0, st>>>(g, static_cast<const uint16_t*>(act), n_points, width, scratch);
  bias_reduce_kernel<<<width, 128, 0, st>>>(scratch, nb, width, gb);
  count_launch(2);
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

int composite_backward(const float* raw, const float* z, const float* noise, const float* g_feature,
                       const float* g_weights, const float* g_depth, int n_rays, int n_samples,
                       float* d_rgb_pre, float* d_sigma_pre, cudaStream_t st, int split, float* amax) {
  CRNERF_REQUIRE(raw && z && d_rgb_pre && d_sigma_pre, "null argument");
  CRNERF_REQUIRE(n_samples >= 1 && n_samples <= kMaxS, "n_samples=%d unsupported by the backward (<= %d)",
                 n_samples, kMaxS);
  if (n_rays == 0) return CRNERF_OK;
  const size_t smem = (size_t)5 * n_samples * sizeof(float);   // <= 20 KB
  if (n_samples <= 32 * kCbWarps)
    composite_backward_kernel<true><<<n_rays, 32 * kCbWarps, smem, st>>>(raw, z, noise, g_feature, g_weights, g_depth,
                                                                        n_rays, n_samples, split, d_rgb_pre, d_sigma_pre, amax);
  else
    composite_backward_kernel<false><<<n_rays, 32 * kCbWarps, smem, st>>>(raw, z, noise, g_feature, g_weights, g_depth,
                                                                         n_rays, n_samples, split, d_rgb_pre, d_sigma_pre, amax);
  count_launch();
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

}  // namespace crnerf

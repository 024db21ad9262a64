// Adam update of the training step (reference utils/__init__.py:31-32: torch.optim.Adam(parameters,
// lr, eps=1e-8, weight_decay) stepped once per batch by train_mask_grid_sample.py's optimizer) as
// ONE launch per 48 parameter tensors instead of the tensor library's ~160 (21 multi-tensor launches
// plus two scalar pow kernels per parameter in its graph-capturable form).
//
// Math of torch.optim.Adam (amsgrad = False), element by element in fp32:
//   g   = grad (+ weight_decay * p)            (negated first when maximize)
//   m   = m + (1 - beta1) (g - m)              (lerp)
//   v   = beta2 v + (1 - beta2) g g
//   p   = p - (lr / (1 - beta1^t)) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)
// t = *step + 1 is read from device memory (the caller increments *step after the launch), so the
// launch is capturable in a CUDA graph; the bias corrections are evaluated in double like the
// optimizer's Python scalars.  The kernel is a pure HBM stream: 4 reads + 3 writes of 4 B per element.
#include <algorithm>
#include "common.h"

namespace crnerf {
namespace {

constexpr int kAdamTensors = 48;     // per launch (table lives in the kernel parameters)
constexpr int kAdamChunk = 4096;     // elements per block

struct AdamTable {
  float* p[kAdamTensors];
  const float* g[kAdamTensors];
  float* m[kAdamTensors];
  float* v[kAdamTensors];
  long long n[kAdamTensors];
  int first_block[kAdamTensors + 1];
  int count;
};

struct AdamHyper {
  double lr, beta1, beta2, eps, weight_decay;
  const float* lr_dev;   // overrides lr when non-null (a tensor learning rate)
  const float* step;     // steps taken so far
  int maximize;
};

__global__ void __launch_bounds__(256) adam_kernel(const __grid_constant__ AdamTable T, const __grid_constant__ AdamHyper H) {
  __shared__ float s_step_size, s_rsqrt_bc2;
  int t = 0;
  while (t + 1 < T.count && (int)blockIdx.x >= T.first_block[t + 1]) ++t;
  if (threadIdx.x == 0) {
    const double step = (double)*H.step + 1.0;
    const double lr = H.lr_dev ? (double)*H.lr_dev : H.lr;
    s_step_size = (float)(lr / (1.0 - pow(H.beta1, step)));
    s_rsqrt_bc2 = (float)sqrt(1.0 - pow(H.beta2, step));
  }
  __syncthreads();
  const float step_size = s_step_size, bc2_sqrt = s_rsqrt_bc2;
  const float w1 = (float)(1.0 - H.beta1), b2 = (float)H.beta2, w2 = (float)(1.0 - H.beta2);
  const float eps = (float)H.eps, wd = (float)H.weight_decay;
  float* __restrict__ p = T.p[t];
  const float* __restrict__ g = T.g[t];
  float* __restrict__ m = T.m[t];
  float* __restrict__ v = T.v[t];
  const long long n = T.n[t];
  const long long e0 = (long long)((int)blockIdx.x - T.first_block[t]) * kAdamChunk;
  const long long e1 = min(n, e0 + kAdamChunk);
  auto update = [&](float& pv, float gv, float& mv, float& vv) {
    if (H.maximize) gv = -gv;
    if (wd != 0.f) gv = fmaf(wd, pv, gv);
    mv = fmaf(w1, gv - mv, mv);
    vv = fmaf(w2 * gv, gv, vv * b2);
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pv = fmaf(-step_size, mv / denom, pv);
  };
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  if (vec) {
    const long long q1 = e0 + ((e1 - e0) & ~3LL);
    for (long long e = e0 + 4LL * threadIdx.x; e < q1; e += 4 * 256) {
      float4 pv = *reinterpret_cast<float4*>(p + e), mv = *reinterpret_cast<float4*>(m + e),
             vv = *reinterpret_cast<float4*>(v + e);
      const float4 gv = __ldg(reinterpret_cast<const float4*>(g + e));
      update(pv.x, gv.x, mv.x, vv.x);
      update(pv.y, gv.y, mv.y, vv.y);
      update(pv.z, gv.z, mv.z, vv.z);
      update(pv.w, gv.w, mv.w, vv.w);
      *reinterpret_cast<float4*>(p + e) = pv;
      *reinterpret_cast<float4*>(m + e) = mv;
      *reinterpret_cast<float4*>(v + e) = vv;
    }
    for (long long e = q1 + threadIdx.x; e < e1; e += 256) update(p[e], g[e], m[e], v[e]);
  } else {
    for (long long e = e0 + threadIdx.x; e < e1; e += 256) update(p[e], g[e], m[e], v[e]);
  }
}

}  // namespace

int adam_step(int n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
              float* const* exp_avg_sq, const int64_t* numel, const float* step, const float* lr_dev, double lr,
              double beta1, double beta2, double eps, double weight_decay, int maximize, cudaStream_t st) {
  CRNERF_REQUIRE(n_tensors >= 0, "n_tensors=%d", n_tensors);
  CRNERF_REQUIRE(step, "null step counter");
  CRNERF_REQUIRE(n_tensors == 0 || (params && grads && exp_avg && exp_avg_sq && numel), "null table");
  CRNERF_REQUIRE(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0, "betas (%g, %g) outside [0, 1)", beta1, beta2);
  CRNERF_REQUIRE(eps >= 0.0 && weight_decay >= 0.0, "eps=%g weight_decay=%g", eps, weight_decay);
  AdamHyper H{lr, beta1, beta2, eps, weight_decay, lr_dev, step, maximize ? 1 : 0};
  int i = 0;
  while (i < n_tensors) {
    AdamTable T;
    int c = 0;
    long long blocks = 0;
    for (; i < n_tensors && c < kAdamTensors; ++i) {
      CRNERF_REQUIRE(numel[i] >= 0, "numel[%d]=%lld", i, (long long)numel[i]);
      if (numel[i] == 0) continue;
      CRNERF_REQUIRE(params[i] && grads[i] && exp_avg[i] && exp_avg_sq[i], "null pointer for tensor %d", i);
      const long long nb = (numel[i] + kAdamChunk - 1) / kAdamChunk;
      if (blocks + nb > 0x7fffffffLL) break;   // next launch
      T.p[c] = params[i]; T.g[c] = grads[i]; T.m[c] = exp_avg[i]; T.v[c] = exp_avg_sq[i];
      T.n[c] = numel[i];
      T.first_block[c] = (int)blocks;
      blocks += nb;
      ++c;
    }
    if (c == 0) {
      CRNERF_REQUIRE(i >= n_tensors, "tensor %d needs more blocks than one launch holds", i);
      break;
    }
    T.first_block[c] = (int)blocks;
    T.count = c;
    adam_kernel<<<(unsigned)blocks, 256, 0, st>>>(T, H);
    count_launch();
  }
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

}  // namespace crnerf

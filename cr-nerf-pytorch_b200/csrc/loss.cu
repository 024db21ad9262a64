// Loss + mask tail of the training step: the reference evaluates
// CRNeRFLoss (losses.py:50-89) and the mask lookup (train_mask_grid_sample.py:170-176) as a few
// dozen launch-bound elementwise/reduction calls on (1024, 3) tensors.  Here:
//   ray_loss_fwd   c_l, f_l, r_ms, r_md in one launch (deterministic two-level sum)
//   ray_loss_bwd   their gradients w.r.t. rgb_coarse, rgb_fine and the mask in one launch
//   pair_loss_fwd/bwd   the embedding terms kl_a / rec_a_random / content_constraint (up to 4 per launch)
//   mask_sample_fwd/bwd bilinear upsample (align_corners=False) evaluated only at the sampled pixels
// Every reduction is block partials -> last-arriving block sums them in block order, so a
// result does not depend on scheduling.
#include <algorithm>
#include "common.h"

namespace crnerf {
namespace {

constexpr int kThreads = 256;
constexpr float kFocusEps = 0.02f;  // losses.py:74

// Sum `v[0..K)` over the block; valid in thread 0.
template <int K>
__device__ __forceinline__ void block_sum(float (&v)[K], float* red /* [K][kThreads/32] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], d);
    if (lane == 0) red[k * (kThreads / 32) + warp] = v[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      float t = 0.f;
      for (int w = 0; w < kThreads / 32; ++w) t += red[k * (kThreads / 32) + w];
      v[k] = t;
    }
  }
}

// Write this block's K partials, then let the last block to arrive add all partials in block
// order.  Returns true (in thread 0 of that last block only) with the totals in v.
template <int K>
__device__ __forceinline__ bool grid_sum(float (&v)[K], float* partial, unsigned int* counter, int nb,
                                         int block) {
  __shared__ bool last;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) partial[block * K + k] = v[k];
    __threadfence();
    last = atomicAdd(counter, 1u) == (unsigned)nb - 1;
  }
  __syncthreads();
  if (!last || threadIdx.x != 0) return false;
  __threadfence();
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = 0.f;
  for (int b = 0; b < nb; ++b)
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] += __ldcg(partial + b * K + k);
  *counter = 0;  // ready for the next call on this scratch
  return true;
}

__global__ void __launch_bounds__(kThreads)
ray_loss_fwd_kernel(const float* __restrict__ coarse, const float* __restrict__ fine,
                    const float* __restrict__ target, const float* __restrict__ mask, long long n,
                    float coef, float size_delta, float digit_delta, float* __restrict__ out,
                    float* partial, unsigned int* counter) {
  __shared__ float red[4 * (kThreads / 32)];
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
    const float m = mask ? mask[i] : 0.f;
    const float keep = 1.f - m;
    const float t0 = target[3 * i], t1 = target[3 * i + 1], t2 = target[3 * i + 2];
    {
      const float d0 = coarse[3 * i] - t0, d1 = coarse[3 * i + 1] - t1, d2 = coarse[3 * i + 2] - t2;
      acc[0] += keep * (d0 * d0) + keep * (d1 * d1) + keep * (d2 * d2);
    }
    if (fine) {
      const float d0 = fine[3 * i] - t0, d1 = fine[3 * i + 1] - t1, d2 = fine[3 * i + 2] - t2;
      acc[1] += keep * (d0 * d0) + keep * (d1 * d1) + keep * (d2 * d2);
    }
    if (mask) {
      acc[2] += m * m;
      const float c = m - 0.5f;
      acc[3] += 1.f / (c * c + kFocusEps);
    }
  }
  block_sum<4>(acc, red);
  if (grid_sum<4>(acc, partial, counter, gridDim.x, blockIdx.x)) {
    const float n3 = 3.f * (float)n, n1 = (float)n;
    out[0] = coef * (0.5f * (acc[0] / n3));
    out[1] = coef * (0.5f * (acc[1] / n3));
    out[2] = coef * ((acc[2] / n1) * size_delta);
    out[3] = coef * ((acc[3] / n1) * digit_delta);
  }
}

// go: upstream gradients of the 4 outputs (device, or NULL = all ones)
__global__ void __launch_bounds__(kThreads)
ray_loss_bwd_kernel(const float* __restrict__ coarse, const float* __restrict__ fine,
                    const float* __restrict__ target, const float* __restrict__ mask, long long n,
                    float coef, float size_delta, float digit_delta, const float* __restrict__ go,
                    float* __restrict__ g_coarse, float* __restrict__ g_fine, float* __restrict__ g_mask) {
  const float go0 = go ? go[0] : 1.f, go1 = go ? go[1] : 1.f, go2 = go ? go[2] : 1.f, go3 = go ? go[3] : 1.f;
  const float inv3 = 1.f / (3.f * (float)n), inv1 = 1.f / (float)n;
  const float kc = go0 * coef * inv3, kf = go1 * coef * inv3;  // 0.5 * 2 = 1
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
    const float m = mask ? mask[i] : 0.f;
    const float keep = 1.f - m;
    const float t0 = target[3 * i], t1 = target[3 * i + 1], t2 = target[3 * i + 2];
    if (g_coarse) {
      g_coarse[3 * i] = kc * keep * (coarse[3 * i] - t0);
      g_coarse[3 * i + 1] = kc * keep * (coarse[3 * i + 1] - t1);
      g_coarse[3 * i + 2] = kc * keep * (coarse[3 * i + 2] - t2);
    }
    float sq = 0.f;
    if (fine) {
      const float d0 = fine[3 * i] - t0, d1 = fine[3 * i + 1] - t1, d2 = fine[3 * i + 2] - t2;
      sq = d0 * d0 + d1 * d1 + d2 * d2;
      if (g_fine) {
        g_fine[3 * i] = kf * keep * d0;
        g_fine[3 * i + 1] = kf * keep * d1;
        g_fine[3 * i + 2] = kf * keep * d2;
      }
    }
    if (mask && g_mask) {
      // c_l sees mask.detach() (losses.py:64); f_l, r_ms and r_md do not
      const float c = m - 0.5f, q = c * c + kFocusEps;
      g_mask[i] = -0.5f * kf * sq + go2 * coef * size_delta * inv1 * 2.f * m -
                  go3 * coef * digit_delta * inv1 * 2.f * c / (q * q);
    }
  }
}

struct PairTerm {
  const float* a;
  const float* b;  // NULL for mode 0
  float* ga;       // backward only (may be NULL)
  float* gb;
  long long n;
  int mode;  // 0 mean(a^2), 1 mean|a-b|, 2 mean((a-b)^2)
  float scale;
};
struct PairTerms {
  PairTerm t[4];
};

__global__ void __launch_bounds__(kThreads)
pair_loss_fwd_kernel(PairTerms terms, float* __restrict__ out, float* partial, unsigned int* counter) {
  __shared__ float red[kThreads / 32];
  const PairTerm& T = terms.t[blockIdx.y];
  float acc[1] = {0.f};
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < T.n; i += (long long)gridDim.x * kThreads) {
    const float d = T.mode == 0 ? T.a[i] : T.a[i] - T.b[i];
    acc[0] += T.mode == 1 ? fabsf(d) : d * d;
  }
  block_sum<1>(acc, red);
  if (grid_sum<1>(acc, partial + (size_t)blockIdx.y * gridDim.x, counter + blockIdx.y, gridDim.x, blockIdx.x))
    out[blockIdx.y] = (acc[0] / (float)T.n) * T.scale;
}

__global__ void __launch_bounds__(kThreads)
pair_loss_bwd_kernel(PairTerms terms, const float* __restrict__ go) {
  const PairTerm& T = terms.t[blockIdx.y];
  const float k = (go ? go[blockIdx.y] : 1.f) * T.scale / (float)T.n;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < T.n; i += (long long)gridDim.x * kThreads) {
    const float d = T.mode == 0 ? T.a[i] : T.a[i] - T.b[i];
    const float g = T.mode == 1 ? (d > 0.f ? k : (d < 0.f ? -k : 0.f)) : 2.f * k * d;
    if (T.ga) T.ga[i] = g;
    if (T.gb) T.gb[i] = -g;
  }
}

// torch's bilinear source index for align_corners=False (aten UpSample.h
// area_pixel_compute_source_index): scale*(dst+0.5)-0.5, clamped at 0
__device__ __forceinline__ void bilinear_tap(int dst, float scale, int in, int& i0, int& i1, float& l1) {
  float src = __fsub_rn(__fmul_rn(scale, (float)dst + 0.5f), 0.5f);
  if (src < 0.f) src = 0.f;
  i0 = min((int)src, in - 1);
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
  l1 = __fsub_rn(src, (float)i0);
}

// pred (C, h, w) -> out (n, C): bilinear upsample to (H, W) read at flat pixel indices idx
// (idx == NULL: pixel i), i.e. interpolate(...)[0].permute(1,2,0).reshape(-1,C)[idx]
__global__ void __launch_bounds__(kThreads)
mask_sample_fwd_kernel(const float* __restrict__ pred, int C, int h, int w, int H, int W,
                       const long long* __restrict__ idx, long long n, float* __restrict__ out) {
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n * C; i += (long long)gridDim.x * kThreads) {
    const long long r = i / C;
    const int c = (int)(i - r * C);
    const long long p = idx ? idx[r] : r;
    const int y = (int)(p / W), x = (int)(p - (long long)y * W);
    int y0, y1, x0, x1;
    float ly, lx;
    bilinear_tap(y, sy, h, y0, y1, ly);
    bilinear_tap(x, sx, w, x0, x1, lx);
    const float* pc = pred + (size_t)c * h * w;
    const float top = __fadd_rn(__fmul_rn(1.f - lx, pc[y0 * w + x0]), __fmul_rn(lx, pc[y0 * w + x1]));
    const float bot = __fadd_rn(__fmul_rn(1.f - lx, pc[y1 * w + x0]), __fmul_rn(lx, pc[y1 * w + x1]));
    out[i] = __fadd_rn(__fmul_rn(1.f - ly, top), __fmul_rn(ly, bot));
  }
}

__global__ void __launch_bounds__(kThreads)
mask_sample_bwd_kernel(const float* __restrict__ g_out, int C, int h, int w, int H, int W,
                       const long long* __restrict__ idx, long long n, float* __restrict__ g_pred) {
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n * C; i += (long long)gridDim.x * kThreads) {
    const long long r = i / C;
    const int c = (int)(i - r * C);
    const long long p = idx ? idx[r] : r;
    const int y = (int)(p / W), x = (int)(p - (long long)y * W);
    int y0, y1, x0, x1;
    float ly, lx;
    bilinear_tap(y, sy, h, y0, y1, ly);
    bilinear_tap(x, sx, w, x0, x1, lx);
    float* pc = g_pred + (size_t)c * h * w;
    const float g = g_out[i];
    atomicAdd(pc + y0 * w + x0, g * (1.f - ly) * (1.f - lx));
    atomicAdd(pc + y0 * w + x1, g * (1.f - ly) * lx);
    atomicAdd(pc + y1 * w + x0, g * ly * (1.f - lx));
    atomicAdd(pc + y1 * w + x1, g * ly * lx);
  }
}

int blocks_for(long long n) {
  return (int)std::max<long long>(1, std::min<long long>(2LL * num_sms(), (n + kThreads * 4 - 1) / (kThreads * 4)));
}

}  // namespace

// scratch layout (floats): [0, 4) counters (as uint32, zero before the first use; the kernels
// leave them zero), [4, ...) block partials.  kLossScratchFloats covers the largest grid.
size_t loss_scratch_floats() { return 4 + (size_t)4 * 2 * num_sms() * 4; }

int ray_loss_forward(const float* coarse, const float* fine, const float* target, const float* mask,
                     int64_t n_rays, float coef, float size_delta, float digit_delta, float* out4,
                     float* scratch, cudaStream_t st) {
  CRNERF_REQUIRE(coarse && target && out4 && scratch, "null argument");
  CRNERF_REQUIRE(n_rays > 0, "n_rays must be positive");
  const int nb = blocks_for(n_rays);
  ray_loss_fwd_kernel<<<nb, kThreads, 0, st>>>(coarse, fine, target, mask, n_rays, coef, size_delta, digit_delta,
                                               out4, scratch + 4, reinterpret_cast<unsigned int*>(scratch));
  count_launch();
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

int ray_loss_backward(const float* coarse, const float* fine, const float* target, const float* mask,
                      int64_t n_rays, float coef, float size_delta, float digit_delta, const float* go4,
                      float* g_coarse, float* g_fine, float* g_mask, cudaStream_t st) {
  CRNERF_REQUIRE(coarse && target, "null argument");
  CRNERF_REQUIRE(n_rays > 0, "n_rays must be positive");
  CRNERF_REQUIRE(!g_fine || fine, "g_fine without rgb_fine");
  ray_loss_bwd_kernel<<<blocks_for(n_rays), kThreads, 0, st>>>(coarse, fine, target, mask, n_rays, coef, size_delta,
                                                               digit_delta, go4, g_coarse, g_fine, g_mask);
  count_launch();
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

static int fill_terms(PairTerms& T, int n_terms, const float* const* a, const float* const* b, const int64_t* n,
                      const int* mode, const float* scale, long long& n_max) {
  CRNERF_REQUIRE(n_terms >= 1 && n_terms <= 4, "n_terms must be 1..4");
  n_max = 0;
  for (int k = 0; k < n_terms; ++k) {
    CRNERF_REQUIRE(a[k] && n[k] > 0, "term %d: null or empty", k);
    CRNERF_REQUIRE(mode[k] >= 0 && mode[k] <= 2, "term %d: mode %d", k, mode[k]);
    CRNERF_REQUIRE(mode[k] == 0 || b[k], "term %d: mode %d needs b", k, mode[k]);
    T.t[k] = PairTerm{a[k], b[k], nullptr, nullptr, (long long)n[k], mode[k], scale[k]};
    n_max = std::max<long long>(n_max, n[k]);
  }
  return CRNERF_OK;
}

int pair_loss_forward(int n_terms, const float* const* a, const float* const* b, const int64_t* n, const int* mode,
                      const float* scale, float* out, float* scratch, cudaStream_t st) {
  CRNERF_REQUIRE(a && b && n && mode && scale && out && scratch, "null argument");
  PairTerms T{};
  long long n_max;
  int rc = fill_terms(T, n_terms, a, b, n, mode, scale, n_max);
  if (rc) return rc;
  const int nb = blocks_for(n_max);
  pair_loss_fwd_kernel<<<dim3(nb, n_terms), kThreads, 0, st>>>(T, out, scratch + 4,
                                                               reinterpret_cast<unsigned int*>(scratch));
  count_launch();
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

int pair_loss_backward(int n_terms, const float* const* a, const float* const* b, const int64_t* n, const int* mode,
                       const float* scale, const float* go, float* const* ga, float* const* gb, cudaStream_t st) {
  CRNERF_REQUIRE(a && b && n && mode && scale && ga && gb, "null argument");
  PairTerms T{};
  long long n_max;
  int rc = fill_terms(T, n_terms, a, b, n, mode, scale, n_max);
  if (rc) return rc;
  for (int k = 0; k < n_terms; ++k) {
    T.t[k].ga = ga[k];
    T.t[k].gb = mode[k] == 0 ? nullptr : gb[k];
  }
  pair_loss_bwd_kernel<<<dim3(blocks_for(n_max), n_terms), kThreads, 0, st>>>(T, go);
  count_launch();
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

int mask_sample_forward(const float* pred, int channels, int h, int w, int H, int W, const int64_t* idx, int64_t n,
                        float* out, cudaStream_t st) {
  CRNERF_REQUIRE(pred && out, "null argument");
  CRNERF_REQUIRE(channels >= 1 && h >= 1 && w >= 1 && H >= 1 && W >= 1, "bad shape");
  CRNERF_REQUIRE(idx || n == (int64_t)H * W, "without idx, n must be H*W");
  if (n <= 0) return CRNERF_OK;
  mask_sample_fwd_kernel<<<blocks_for(n * channels), kThreads, 0, st>>>(
      pred, channels, h, w, H, W, reinterpret_cast<const long long*>(idx), n, out);
  count_launch();
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

int mask_sample_backward(const float* g_out, int channels, int h, int w, int H, int W, const int64_t* idx, int64_t n,
                         float* g_pred, cudaStream_t st) {
  CRNERF_REQUIRE(g_out && g_pred, "null argument");
  CRNERF_REQUIRE(channels >= 1 && h >= 1 && w >= 1 && H >= 1 && W >= 1, "bad shape");
  CRNERF_REQUIRE(idx || n == (int64_t)H * W, "without idx, n must be H*W");
  CRNERF_CUDA(cudaMemsetAsync(g_pred, 0, sizeof(float) * channels * h * w, st));
  if (n <= 0) return CRNERF_OK;
  mask_sample_bwd_kernel<<<blocks_for(n * channels), kThreads, 0, st>>>(
      g_out, channels, h, w, H, W, reinterpret_cast<const long long*>(idx), n, g_pred);
  count_launch();
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

}  // namespace crnerf

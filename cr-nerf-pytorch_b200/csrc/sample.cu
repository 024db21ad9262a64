// Depth sampling kernels: PosEmbedding, stratified coarse depths, and the
// inverse-CDF importance sampler fused with the coarse/fine merge.
//   pos_embed        <- PosEmbedding.forward        (reference models/nerf.py:17-30)
//   coarse_z         <- render_rays_cross_ray       (reference models/rendering.py:161-176)
//   sample_pdf(+sort)<- sample_pdf + cat + sort     (reference models/rendering.py:7-46, 183-187)
// These are latency/HBM-bound (< 1% of the path); the design goals are one
// launch each, coalesced row access, and the reference's rounding order
// (explicit __f*_rn so nvcc does not contract mul+add into FMA).
#include <math_constants.h>
#include "common.h"

namespace crnerf {
namespace {

// ---------------------------------------------------------------------------
__global__ void pos_embed_kernel(const float* __restrict__ x, long long n, int n_freqs,
                                 float* __restrict__ out) {
  const int width = 3 + 6 * n_freqs;
  const long long total = n * (n_freqs + 1);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / (n_freqs + 1);
    const int k = (int)(i - row * (n_freqs + 1));  // 0: identity, k>=1: band k-1
    const float* xr = x + row * 3;
    float* o = out + row * width;
    if (k == 0) {
      o[0] = xr[0];
      o[1] = xr[1];
      o[2] = xr[2];
    } else {
      const float f = __int_as_float((127 + (k - 1)) << 23);  // 2^(k-1), exact
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float s, cs;
        sincosf(__fmul_rn(f, xr[c]), &s, &cs);
        o[3 + 6 * (k - 1) + c] = s;
        o[6 + 6 * (k - 1) + c] = cs;
      }
    }
  }
}

// ---------------------------------------------------------------------------
__device__ __forceinline__ float z_at(float near, float far, float t, int use_disp) {
  const float omt = __fsub_rn(1.f, t);
  if (!use_disp) return __fadd_rn(__fmul_rn(near, omt), __fmul_rn(far, t));
  const float a = __fmul_rn(__fdiv_rn(1.f, near), omt);
  const float b = __fmul_rn(__fdiv_rn(1.f, far), t);
  return __fdiv_rn(1.f, __fadd_rn(a, b));
}

__global__ void coarse_z_kernel(const float* __restrict__ rays, const float* __restrict__ t_steps,
                                const float* __restrict__ perturb_rand, int n_rays, int S,
                                int use_disp, float* __restrict__ z_out) {
  const long long total = (long long)n_rays * S;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ray = (int)(i / S);
    const int s = (int)(i - (long long)ray * S);
    const float near = __ldg(rays + (long long)ray * 8 + 6), far = __ldg(rays + (long long)ray * 8 + 7);
    const float z = z_at(near, far, __ldg(t_steps + s), use_disp);
    if (!perturb_rand) {
      z_out[i] = z;
      continue;
    }
    // stratified jitter: lower + (upper - lower) * rand  (rendering.py:169-176)
    const float zp = s > 0 ? z_at(near, far, __ldg(t_steps + s - 1), use_disp) : z;
    const float zn = s + 1 < S ? z_at(near, far, __ldg(t_steps + s + 1), use_disp) : z;
    const float lower = s > 0 ? __fmul_rn(0.5f, __fadd_rn(zp, z)) : z;
    const float upper = s + 1 < S ? __fmul_rn(0.5f, __fadd_rn(z, zn)) : z;
    z_out[i] = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), __ldg(perturb_rand + i)));
  }
}

// ---------------------------------------------------------------------------
// One warp per ray.  Shared memory per warp: cdf[m+1] | sort buffer[pow2(n_coarse+n_imp)].
// kMerge: bins are the midpoints of z (n_coarse = m+2 depths), weights =
// weights_coarse[:, 1:-1]; output is sort(cat(z, samples)).
template <bool kMerge>
__global__ void __launch_bounds__(128)
sample_pdf_kernel(const float* __restrict__ bins_or_z, const float* __restrict__ weights,
                  const float* __restrict__ u, long long u_stride, int n_rays, int m,
                  int n_imp, float eps, float* __restrict__ samples_out,
                  float* __restrict__ sorted_out, int cdf_pad, int sort_pad) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ray = blockIdx.x * 4 + warp;
  if (ray >= n_rays) return;
  float* cdf = sm + warp * (cdf_pad + sort_pad);
  float* buf = cdf + cdf_pad;
  const int n_coarse = m + 2;
  const float* wrow = kMerge ? weights + (long long)ray * n_coarse + 1 : weights + (long long)ray * m;
  const float* brow = kMerge ? bins_or_z + (long long)ray * n_coarse : bins_or_z + (long long)ray * (m + 1);

  // pdf normaliser.  torch sums/accumulates in higher precision on the CPU
  // (cumsum uses a double accumulator); doing the same here keeps the cdf
  // order-independent to the last bit in nearly every case.
  double part = 0.0;
  for (int j = lane; j < m; j += 32) part += (double)__fadd_rn(wrow[j], eps);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
  const float total = (float)part;

  // cdf[0] = 0, cdf[j+1] = sum_{i<=j} pdf_i  (warp scan over 32-wide blocks)
  double carry = 0.0;
  if (lane == 0) cdf[0] = 0.f;
  for (int base = 0; base < m; base += 32) {
    const int j = base + lane;
    double v = j < m ? (double)__fdiv_rn(__fadd_rn(wrow[j], eps), total) : 0.0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const double o = __shfl_up_sync(0xffffffffu, v, d);
      if (lane >= d) v += o;
    }
    v += carry;
    if (j < m) cdf[j + 1] = (float)v;
    carry = __shfl_sync(0xffffffffu, v, 31);
  }
  __syncwarp();

  for (int i = lane; i < n_imp; i += 32) {
    const float ui = __ldg(u + (long long)ray * u_stride + i);
    // searchsorted(cdf, u, right=True): first index with cdf[idx] > u, over m+1 entries
    int lo = 0, hi = m + 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cdf[mid] <= ui) lo = mid + 1; else hi = mid;
    }
    const int below = max(lo - 1, 0), above = min(lo, m);
    const float c0 = cdf[below], c1 = cdf[above];
    float b0, b1;
    if (kMerge) {
      b0 = __fmul_rn(0.5f, __fadd_rn(brow[below], brow[below + 1]));
      b1 = __fmul_rn(0.5f, __fadd_rn(brow[above], brow[above + 1]));
    } else {
      b0 = brow[below];
      b1 = brow[above];
    }
    float denom = __fsub_rn(c1, c0);
    if (denom < eps) denom = 1.f;
    const float smp =
        __fadd_rn(b0, __fmul_rn(__fdiv_rn(__fsub_rn(ui, c0), denom), __fsub_rn(b1, b0)));
    if (samples_out) samples_out[(long long)ray * n_imp + i] = smp;
    if (kMerge) buf[n_coarse + i] = smp;
  }
  if (!kMerge) return;

  const int n_tot = n_coarse + n_imp;
  for (int i = lane; i < n_coarse; i += 32) buf[i] = brow[i];
  __syncwarp();
  // Fast path: both runs are already sorted (coarse depths always are; the new samples are
  // whenever u is non-decreasing, e.g. the shared linspace of eval mode) -> merge by rank:
  // an element's final position is its index in its own run plus the number of elements of
  // the other run that precede it (ties: coarse first).  Two binary searches per lane-step
  // instead of a 36-stage bitonic network.  Any other order of equal values would produce the
  // same sorted row, so this is bit-identical to sort(cat(...)).
  bool sorted_runs = true;
  for (int i = lane; i + 1 < n_imp; i += 32) sorted_runs &= buf[n_coarse + i] <= buf[n_coarse + i + 1];
  for (int i = lane; i + 1 < n_coarse; i += 32) sorted_runs &= buf[i] <= buf[i + 1];
  if (__all_sync(0xffffffffu, sorted_runs)) {
    float* orow = sorted_out + (long long)ray * n_tot;
    const float* sc = buf;             // coarse run
    const float* sn = buf + n_coarse;  // new samples
    for (int i = lane; i < n_coarse; i += 32) {   // # new samples strictly below sc[i]
      const float v = sc[i];
      int lo = 0, hi = n_imp;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (sn[mid] < v) lo = mid + 1; else hi = mid;
      }
      orow[i + lo] = v;
    }
    for (int j = lane; j < n_imp; j += 32) {      // # coarse depths not above sn[j]
      const float v = sn[j];
      int lo = 0, hi = n_coarse;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (sc[mid] <= v) lo = mid + 1; else hi = mid;
      }
      orow[j + lo] = v;
    }
    return;
  }
  for (int i = n_tot + lane; i < sort_pad; i += 32) buf[i] = CUDART_INF_F;
  __syncwarp();
  // bitonic sort of sort_pad (power of two) values held in shared memory
  for (int k = 2; k <= sort_pad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < sort_pad; i += 32) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const float a = buf[i], b = buf[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            buf[i] = b;
            buf[ixj] = a;
          }
        }
      }
      __syncwarp();
    }
  }
  for (int i = lane; i < n_tot; i += 32) sorted_out[(long long)ray * n_tot + i] = buf[i];
}

// ---------------------------------------------------------------------------
// Camera rays of a pinhole frame (reference datasets/ray_utils.py:5-52 + the (h*w, 8) row the
// datasets build, phototourism_mask_grid_sample.py:300-307): pixel (i, j) -> camera direction
// ((i-cx)/fx, -(j-cy)/fy, -1), rotated by c2w[:, :3], normalised; origin c2w[:, 3]; near, far.
struct RayGenParams {
  float fx, fy, cx, cy;
  float c2w[12];
  float near, far;
  int H, W;
};
__global__ void generate_rays_kernel(const __grid_constant__ RayGenParams P, float* __restrict__ rays) {
  const long long total = (long long)P.H * P.W;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx / P.W), i = (int)(idx - (long long)j * P.W);
    const float dx = __fdiv_rn(__fsub_rn((float)i, P.cx), P.fx);
    const float dy = -__fdiv_rn(__fsub_rn((float)j, P.cy), P.fy);
    const float dz = -1.f;
    float d[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)   // directions @ c2w[:, :3].T
      d[k] = __fadd_rn(__fadd_rn(__fmul_rn(dx, P.c2w[4 * k]), __fmul_rn(dy, P.c2w[4 * k + 1])),
                       __fmul_rn(dz, P.c2w[4 * k + 2]));
    const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])),
                                           __fmul_rn(d[2], d[2])));
    float4* o = reinterpret_cast<float4*>(rays + idx * 8);
    o[0] = make_float4(P.c2w[3], P.c2w[7], P.c2w[11], __fdiv_rn(d[0], nrm));
    o[1] = make_float4(__fdiv_rn(d[1], nrm), __fdiv_rn(d[2], nrm), P.near, P.far);
  }
}

// ---------------------------------------------------------------------------
// Output stage of the eval loop (reference eval.py:295-297): rgb (3, n) planar fp32 in [0,1] ->
// (n, 3) interleaved uint8 = uint8(clip(x, 0, 1) * 255) (truncation, as numpy's astype).
__global__ void rgb_to_u8_kernel(const float* __restrict__ rgb, long long n, uint8_t* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = fminf(fmaxf(rgb[(long long)c * n + i], 0.f), 1.f);
      out[i * 3 + c] = (uint8_t)(int)__fmul_rn(v, 255.f);
    }
  }
}

// ---------------------------------------------------------------------------
// Grid-sampled training patch (reference datasets/phototourism_mask_grid_sample.py:241-275): a
// g x g lattice (g = sqrt(batch_size)) laid over one training image at a random scale / offset;
// lattice point (i, j) - i along the width, j along the height - lands on pixel
//   w = floor((lin_w[i]*scale + w_off) * img_w),  h = floor((lin_h[j]*scale + h_off) * img_h)
// and output row r = j*g + i (the reference's .permute(1, 0).view(-1)) gathers that pixel's
// cached ray row (o3 d3 near far | image id, the (N,9) ray cache) and rgb.  Every product and sum is
// rounded separately, as torch's elementwise ops do, so indices and uv match bit for bit; the two
// linspace vectors come from the caller (torch.linspace, as the reference builds them).
// The reference keeps the image sizes as fp32 (all_imgs_wh is a float tensor, :197), so the
// offset of the image inside the cache is an fp32 sum and the cache row is formed in fp32
// (int64 index + 0-dim float tensor promotes to float32, :266): beyond 2^24 rows that rounds, and
// the same rounded row is gathered here.
struct GridPatchParams {
  const float* lin_w;
  const float* lin_h;
  const float* all_rays;   // (M, 9)
  const float* all_rgbs;   // (M, 3)
  long long n_cache_rows;
  float image_offset;      // fp32, as the reference computes it
  float scale, h_off, w_off;
  float img_w, img_h;
  int g;
  float* rays;        // (g*g, 8)
  long long* ts;      // (g*g)
  float* rgbs;        // (g*g, 3)
  long long* rgb_idx; // (g*g)  index inside the image
  float* uv;          // (g*g, 2)  [h_sb, w_sb]
  int* status;        // set to 1 if a gathered row falls outside the cache (never silently clamped)
};
__global__ void grid_patch_kernel(const __grid_constant__ GridPatchParams P) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= P.g * P.g) return;
  const int j = r / P.g, i = r - j * P.g;
  const float h_sb = __fadd_rn(__fmul_rn(P.lin_h[j], P.scale), P.h_off);
  const float w_sb = __fadd_rn(__fmul_rn(P.lin_w[i], P.scale), P.w_off);
  const float h = floorf(__fmul_rn(h_sb, P.img_h));
  const float w = floorf(__fmul_rn(w_sb, P.img_w));
  const long long idx = (long long)__fadd_rn(w, __fmul_rn(h, P.img_w));   // fp32, then .long()
  const long long row = (long long)__fadd_rn((float)idx, P.image_offset);  // (idx + offset_f32).long()
  P.rgb_idx[r] = idx;
  P.uv[2 * r] = h_sb;
  P.uv[2 * r + 1] = w_sb;
  if (row < 0 || row >= P.n_cache_rows) {
    if (P.status) *P.status = 1;
    return;
  }
  const float* src = P.all_rays + row * 9;
#pragma unroll
  for (int k = 0; k < 8; ++k) P.rays[(long long)r * 8 + k] = src[k];
  P.ts[r] = (long long)src[8];
#pragma unroll
  for (int k = 0; k < 3; ++k) P.rgbs[(long long)r * 3 + k] = P.all_rgbs[row * 3 + k];
}

int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = (long long)num_sms() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

int pos_embed(const float* x, int64_t n, int n_freqs, float* out, cudaStream_t st) {
  CRNERF_REQUIRE(x && out, "null argument");
  CRNERF_REQUIRE(n_freqs >= 0 && n_freqs <= 32, "n_freqs out of range");
  if (n == 0) return CRNERF_OK;
  pos_embed_kernel<<<grid_for(n * (n_freqs + 1), 256), 256, 0, st>>>(x, n, n_freqs, out);
  count_launch();
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

int rgb_to_u8(const float* rgb, int64_t n, uint8_t* out, cudaStream_t st) {
  CRNERF_REQUIRE(rgb && out, "null argument");
  if (n <= 0) return CRNERF_OK;
  rgb_to_u8_kernel<<<grid_for(n, 256), 256, 0, st>>>(rgb, n, out);
  count_launch();
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

int generate_rays(const float* intr4_host, const float* c2w12_host, float near, float far, int H, int W,
                  float* rays, cudaStream_t st) {
  CRNERF_REQUIRE(intr4_host && c2w12_host && rays, "null argument");
  CRNERF_REQUIRE(H >= 0 && W >= 0, "negative frame size");
  CRNERF_REQUIRE((reinterpret_cast<uintptr_t>(rays) & 15) == 0, "rays must be 16-byte aligned");
  if ((long long)H * W == 0) return CRNERF_OK;
  RayGenParams P;
  P.fx = intr4_host[0]; P.fy = intr4_host[1]; P.cx = intr4_host[2]; P.cy = intr4_host[3];
  for (int i = 0; i < 12; ++i) P.c2w[i] = c2w12_host[i];
  P.near = near; P.far = far; P.H = H; P.W = W;
  generate_rays_kernel<<<grid_for((long long)H * W, 256), 256, 0, st>>>(P, rays);
  count_launch();
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

int coarse_z(const float* rays, const float* t_steps, const float* perturb_rand, int n_rays,
             int n_samples, int use_disp, float* z, cudaStream_t st) {
  CRNERF_REQUIRE(rays && t_steps && z, "null argument");
  CRNERF_REQUIRE(n_samples >= 1, "n_samples must be positive");
  if (n_rays == 0) return CRNERF_OK;
  coarse_z_kernel<<<grid_for((long long)n_rays * n_samples, 256), 256, 0, st>>>(
      rays, t_steps, perturb_rand, n_rays, n_samples, use_disp, z);
  count_launch();
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

int sample_pdf(const float* bins_or_z, const float* weights, const float* u, int64_t u_stride,
               int n_rays, int m, int n_imp, float eps, float* samples, float* sorted, bool merge,
               cudaStream_t st) {
  CRNERF_REQUIRE(bins_or_z && weights && u, "null argument");
  CRNERF_REQUIRE(m >= 1 && n_imp >= 1, "need at least one bin and one sample");
  CRNERF_REQUIRE(!merge || sorted, "merge needs an output buffer");
  CRNERF_REQUIRE(merge || samples, "samples output is null");
  if (n_rays == 0) return CRNERF_OK;
  const int cdf_pad = (m + 1 + 3) & ~3;
  const int sort_pad = merge ? next_pow2(m + 2 + n_imp) : 0;
  CRNERF_REQUIRE(sort_pad <= 4096 && cdf_pad <= 4096, "too many samples per ray (max 4096)");
  const size_t smem = 4 * (size_t)(cdf_pad + sort_pad) * sizeof(float);
  const int grid = (n_rays + 3) / 4;
  if (merge) {
    if (smem > 48 * 1024)
      CRNERF_CUDA(cudaFuncSetAttribute(sample_pdf_kernel<true>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sample_pdf_kernel<true><<<grid, 128, smem, st>>>(bins_or_z, weights, u, u_stride, n_rays, m,
                                                     n_imp, eps, samples, sorted, cdf_pad, sort_pad);
  } else {
    sample_pdf_kernel<false><<<grid, 128, smem, st>>>(bins_or_z, weights, u, u_stride, n_rays, m,
                                                      n_imp, eps, samples, sorted, cdf_pad, sort_pad);
  }
  count_launch();
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

int grid_patch(const float* lin_w, const float* lin_h, int g, float img_w, float img_h, float scale, float h_off,
               float w_off, const float* all_rays, const float* all_rgbs, long long n_cache_rows,
               float image_offset, float* rays, long long* ts, float* rgbs, long long* rgb_idx, float* uv,
               int* status, cudaStream_t st) {
  CRNERF_REQUIRE(lin_w && lin_h && all_rays && all_rgbs && rays && ts && rgbs && rgb_idx && uv, "null argument");
  CRNERF_REQUIRE(g >= 1 && g <= 4096 && img_w >= 1.f && img_h >= 1.f, "bad lattice / image size");
  CRNERF_REQUIRE(n_cache_rows >= 1 && image_offset >= 0.f, "bad ray cache extent");
  GridPatchParams P{lin_w, lin_h, all_rays, all_rgbs, n_cache_rows, image_offset, scale, h_off, w_off,
                    img_w, img_h, g, rays, ts, rgbs, rgb_idx, uv, status};
  grid_patch_kernel<<<(g * g + 127) / 128, 128, 0, st>>>(P);
  count_launch();
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

}  // namespace crnerf

// encoder_sameoutputsize.forward (reference models/linearStyleTransfer.py:208-276; SURVEY.md 8f
// rank 1): the style/content encoder enc_a / enc_cont, inference only.
//
//   conv1 1x1 3->3 . reflect-pad . conv2 3x3 3->64 . LeakyReLU           enc_first_kernel (fp32 CUDA cores, 0.9 % of the MACs)
//   [pad . conv3 64->64 . LReLU . maxpool2] [pad . conv4 64->128 . LReLU]
//   [pad . conv5 128->128 . LReLU . maxpool2] [pad . conv6 128->128 . LReLU]   enc_conv_tc_kernel (tcgen05 implicit GEMM)
//   adaptive-avg-pool 32x32 . conv7 1x1 128->64 . LeakyReLU               enc_tail_kernel
//
// The 3x3 convolutions are 99 % of the work (103 k MAC per input pixel).  They run as implicit
// GEMMs on the tensor cores: D[128 pixels x Cout] += A_tap[128 x 64] * W_tap[Cout x 64]^T over
// 9 taps x Cin/64 channel blocks.  fp32-class accuracy comes from splitting both operands into
// fp16 hi + lo and issuing three MMAs per product (hi*hi + lo*hi + hi*lo, fp32 accumulate), as the
// cross-ray Gram kernel does.
//
// Activation layout between layers ("planes"): [C/8][H+2][W+2][8] fp16, twice (hi, lo), with the
// reflection halo materialised.  One 16-byte element holds 8 channels of one pixel, so
//   * a run of pixels of one channel chunk is contiguous -> the producer warp stages a tile with
//     plain cp.async.bulk copies (130 pixels x 8 chunks x {hi, lo} per tap row);
//   * in shared memory every pixel of a chunk is 16 B after its neighbour -> the SWIZZLE_NONE
//     K-major descriptor (LBO = chunk stride, SBO = 128 B) reads ANY 128 consecutive pixels, and a
//     3x3 tap is the row slot dy with the descriptor start address shifted by dx * 16 B
//     (validated bit-exactly by tools/nosw_probe.cu).
// A tile is 128 consecutive positions of the flattened padded grid, so rows of any width pack
// tiles densely (W / (W+2) useful rows); halo positions compute garbage that is not stored.
// The producer stages one tap row per ring slot (the issuer reads one row at a time, so the other
// slots are prefetch).  Layers followed by a max-pool write fp32 NHWC and enc_prep_kernel (one HBM
// pass) pools, splits and adds the halo; the others write the next layer's planes directly.
#include <cuda_fp16.h>
#include <algorithm>
#include "common.h"
#include "ptx.cuh"

namespace crnerf {
namespace {

constexpr float kSlope = 0.2f;

__host__ __device__ __forceinline__ int reflect_idx(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }
__device__ __forceinline__ float lrelu(float v) { return v > 0.f ? v : v * kSlope; }


// hi/lo split of 8 consecutive channels of one pixel -> two 16-byte plane elements
__device__ __forceinline__ void split8(const float (&v)[8], uint4& h, uint4& l) {
  uint32_t hh[4], ll[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __half h0 = __float2half_rn(v[2 * j]), h1 = __float2half_rn(v[2 * j + 1]);
    const __half l0 = __float2half_rn(v[2 * j] - __half2float(h0));
    const __half l1 = __float2half_rn(v[2 * j + 1] - __half2float(h1));
    hh[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
    ll[j] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
  }
  h = make_uint4(hh[0], hh[1], hh[2], hh[3]);
  l = make_uint4(ll[0], ll[1], ll[2], ll[3]);
}

// Padded positions that hold interior pixel (yp, xp) (padded coordinates, 1-based interior) of an
// H x W layer: itself plus its reflection-halo copies (row 0 mirrors row 2, row H+1 mirrors H-1).
struct HaloTargets {
  int ys[3], xs[3], ny, nx;
  __device__ __forceinline__ HaloTargets(int yp, int xp, int H, int W) {
    ny = nx = 0;
    ys[ny++] = yp;
    if (yp == 2) ys[ny++] = 0;
    if (yp == H - 1) ys[ny++] = H + 1;
    xs[nx++] = xp;
    if (xp == 2) xs[nx++] = 0;
    if (xp == W - 1) xs[nx++] = W + 1;
  }
};
__device__ __forceinline__ void store_plane_elem(__half* hi, __half* lo, int Hp, int Wp, int chunk,
                                                 const HaloTargets& t, const uint4& h, const uint4& l) {
  for (int a = 0; a < t.ny; ++a)
    for (int b = 0; b < t.nx; ++b) {
      const size_t o = (((size_t)chunk * Hp + t.ys[a]) * Wp + t.xs[b]) * 8;
      *reinterpret_cast<uint4*>(hi + o) = h;
      *reinterpret_cast<uint4*>(lo + o) = l;
    }
}

// ---- packed weight image ----------------------------------------------------------------------
// [conv3 | conv4 | conv5 | conv6] tensor-core chunks, then an fp32 blob.
// chunk (kb, tap) of a layer = [hi: Cout rows x 128 B, SWIZZLE_128B][lo: same]; row co, k = ci - 64 kb.
struct TcLayer {
  int cin, cout;
  size_t offset;  // bytes from the start of the image
  __host__ __device__ size_t chunk_bytes() const { return (size_t)2 * cout * 128; }
  __host__ __device__ size_t bytes() const { return chunk_bytes() * 9 * (cin / 64); }
};
constexpr int kTcCin[4] = {64, 64, 128, 128};
constexpr int kTcCout[4] = {64, 128, 128, 128};

struct Blob {  // float offsets inside the fp32 blob
  static constexpr int w1 = 0, b1 = 9, w2t = 12 /* [27][64] */, b2 = w2t + 27 * 64, b3 = b2 + 64, b4 = b3 + 64,
                       b5 = b4 + 128, b6 = b5 + 128, w7 = b6 + 128 /* [64][128] */, b7 = w7 + 64 * 128,
                       total = b7 + 64;
};

void tc_layers(TcLayer (&L)[4], size_t& blob_offset) {
  size_t off = 0;
  for (int i = 0; i < 4; ++i) {
    L[i] = TcLayer{kTcCin[i], kTcCout[i], off};
    off += L[i].bytes();
  }
  blob_offset = off;
}

__global__ void enc_pack_tc_kernel(const float* __restrict__ w, int cin, int cout, uint8_t* __restrict__ img) {
  const long long total = (long long)9 * cin * cout;
  const size_t chunk = (size_t)2 * cout * 128;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % 64);
    long long r = i / 64;
    const int co = (int)(r % cout);
    r /= cout;
    const int tap = (int)(r % 9), kb = (int)(r / 9);
    const float v = w[((size_t)co * cin + kb * 64 + k) * 9 + tap];
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    uint8_t* base = img + (size_t)(kb * 9 + tap) * chunk + sw128_offset(co, k >> 3) + (k & 7) * 2;
    *reinterpret_cast<__half*>(base) = hi;
    *reinterpret_cast<__half*>(base + (size_t)cout * 128) = lo;
  }
}

__global__ void enc_pack_blob_kernel(const float* w1, const float* b1, const float* w2, const float* b2,
                                     const float* b3, const float* b4, const float* b5, const float* b6,
                                     const float* w7, const float* b7, float* __restrict__ blob) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Blob::total; i += gridDim.x * blockDim.x) {
    float v;
    if (i < Blob::b1) v = w1[i];
    else if (i < Blob::w2t) v = b1[i - Blob::b1];
    else if (i < Blob::b2) {
      const int k = (i - Blob::w2t) / 64, co = (i - Blob::w2t) % 64;
      v = w2[co * 27 + k];
    } else if (i < Blob::b3) v = b2[i - Blob::b2];
    else if (i < Blob::b4) v = b3[i - Blob::b3];
    else if (i < Blob::b5) v = b4[i - Blob::b4];
    else if (i < Blob::b6) v = b5[i - Blob::b5];
    else if (i < Blob::w7) v = b6[i - Blob::b6];
    else if (i < Blob::b7) v = w7[i - Blob::w7];
    else v = b7[i - Blob::b7];
    blob[i] = v;
  }
}

// ---- conv1 + pad + conv2 + LeakyReLU: img (3,H,W) -> planes (64 ch, H, W) hi/lo with halo ------------
__global__ void __launch_bounds__(128)
enc_first_kernel(const float* __restrict__ img, int H, int W, const float* __restrict__ blob,
                 __half* __restrict__ out_hi, __half* __restrict__ out_lo) {
  __shared__ __align__(16) float s_w2t[27 * 64];
  __shared__ float s_b2[64], s_w1[9], s_b1[3];
  for (int i = threadIdx.x; i < 27 * 64; i += blockDim.x) s_w2t[i] = blob[Blob::w2t + i];
  if (threadIdx.x < 64) s_b2[threadIdx.x] = blob[Blob::b2 + threadIdx.x];
  if (threadIdx.x < 9) s_w1[threadIdx.x] = blob[Blob::w1 + threadIdx.x];
  if (threadIdx.x < 3) s_b1[threadIdx.x] = blob[Blob::b1 + threadIdx.x];
  __syncthreads();
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= (long long)H * W) return;
  const int y = (int)(p / W), x = (int)(p - (long long)y * W);
  float in[27];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int sy = reflect_idx(y + ky - 1, H), sx = reflect_idx(x + kx - 1, W);
      const size_t o = (size_t)sy * W + sx;
      const float v0 = img[o], v1 = img[(size_t)H * W + o], v2 = img[(size_t)2 * H * W + o];
#pragma unroll
      for (int c = 0; c < 3; ++c)
        in[c * 9 + ky * 3 + kx] = s_b1[c] + (s_w1[c * 3] * v0 + s_w1[c * 3 + 1] * v1 + s_w1[c * 3 + 2] * v2);
    }
  const HaloTargets tg(y + 1, x + 1, H, W);
#pragma unroll
  for (int c0 = 0; c0 < 64; c0 += 16) {
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = s_b2[c0 + j];
#pragma unroll
    for (int k = 0; k < 27; ++k) {
      const float4* wr = reinterpret_cast<const float4*>(s_w2t + k * 64 + c0);
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) {
        const float4 w = wr[j4];
        acc[j4 * 4 + 0] = fmaf(in[k], w.x, acc[j4 * 4 + 0]);
        acc[j4 * 4 + 1] = fmaf(in[k], w.y, acc[j4 * 4 + 1]);
        acc[j4 * 4 + 2] = fmaf(in[k], w.z, acc[j4 * 4 + 2]);
        acc[j4 * 4 + 3] = fmaf(in[k], w.w, acc[j4 * 4 + 3]);
      }
    }
#pragma unroll
    for (int h8 = 0; h8 < 2; ++h8) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = lrelu(acc[h8 * 8 + j]);
      uint4 h, l;
      split8(v, h, l);
      store_plane_elem(out_hi, out_lo, H + 2, W + 2, c0 / 8 + h8, tg, h, l);
    }
  }
}

// ---- F (Hi,Wi,C) fp32 -> [maxpool 2x2] -> planes [C/8][Ho+2][Wo+2][8] hi, lo with reflection halo ----
template <bool kPool>
__global__ void __launch_bounds__(256)
enc_prep_kernel(const float* __restrict__ f, int Wi, int C, int Ho, int Wo, __half* __restrict__ hi,
                __half* __restrict__ lo) {
  const int Hp = Ho + 2, Wp = Wo + 2;
  const long long plane = (long long)Hp * Wp, total = plane * (C / 8);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i / plane);
    const long long q = i - c * plane;
    const int yp = (int)(q / Wp), xp = (int)(q - (long long)yp * Wp);
    const int y = reflect_idx(yp - 1, Ho), x = reflect_idx(xp - 1, Wo);
    float v[8];
    if constexpr (kPool) {
      const float* s = f + ((size_t)(2 * y) * Wi + 2 * x) * C + c * 8;
      const float4 a0 = *reinterpret_cast<const float4*>(s), a1 = *reinterpret_cast<const float4*>(s + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(s + C), b1 = *reinterpret_cast<const float4*>(s + C + 4);
      const float* t = s + (size_t)Wi * C;
      const float4 c0 = *reinterpret_cast<const float4*>(t), c1 = *reinterpret_cast<const float4*>(t + 4);
      const float4 d0 = *reinterpret_cast<const float4*>(t + C), d1 = *reinterpret_cast<const float4*>(t + C + 4);
      v[0] = fmaxf(fmaxf(a0.x, b0.x), fmaxf(c0.x, d0.x));
      v[1] = fmaxf(fmaxf(a0.y, b0.y), fmaxf(c0.y, d0.y));
      v[2] = fmaxf(fmaxf(a0.z, b0.z), fmaxf(c0.z, d0.z));
      v[3] = fmaxf(fmaxf(a0.w, b0.w), fmaxf(c0.w, d0.w));
      v[4] = fmaxf(fmaxf(a1.x, b1.x), fmaxf(c1.x, d1.x));
      v[5] = fmaxf(fmaxf(a1.y, b1.y), fmaxf(c1.y, d1.y));
      v[6] = fmaxf(fmaxf(a1.z, b1.z), fmaxf(c1.z, d1.z));
      v[7] = fmaxf(fmaxf(a1.w, b1.w), fmaxf(c1.w, d1.w));
    } else {
      const float* s = f + ((size_t)y * Wi + x) * C + c * 8;
      const float4 a0 = *reinterpret_cast<const float4*>(s), a1 = *reinterpret_cast<const float4*>(s + 4);
      v[0] = a0.x, v[1] = a0.y, v[2] = a0.z, v[3] = a0.w, v[4] = a1.x, v[5] = a1.y, v[6] = a1.z, v[7] = a1.w;
    }
    uint4 h, l;
    split8(v, h, l);
    *reinterpret_cast<uint4*>(hi + (size_t)i * 8) = h;
    *reinterpret_cast<uint4*>(lo + (size_t)i * 8) = l;
  }
}

// ---- 3x3 convolution + bias + LeakyReLU on the tensor cores ----------------------------------------
__device__ __forceinline__ uint64_t make_sdesc_k_nosw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3ffff) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fff) << 16;   // K-adjacent core matrices
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fff) << 32;   // 8-row groups
  d |= static_cast<uint64_t>(1) << 46;
  return d;  // layout type 0: no swizzle
}

constexpr int kTilePix = 128;
constexpr int kRun = kTilePix + 2;         // pixels staged per tap row
constexpr int kChunkStride = kRun * 16;    // bytes between K-adjacent core matrices (LBO)
constexpr int kRowHalf = 8 * kChunkStride; // one tap row, one of {hi, lo}: 8 channel chunks x 130 pixels
constexpr int kRowBytes = 2 * kRowHalf;    // hi then lo
constexpr int kRingW = 3;                  // weight chunks in flight
constexpr int kConvThreads = 256;

// tap rows in flight: the issuer reads one at a time, so every further slot is prefetch
template <int COUT>
__host__ __device__ constexpr int ring_a() {
  return COUT == 64 ? 5 : 3;
}
template <int COUT>
constexpr int conv_smem_bytes() {
  return kRingW * 2 * COUT * 128 + ring_a<COUT>() * kRowBytes + 256 + 1024;
}

// kPlanes: write the next layer's input directly (hi/lo planes with the reflection halo, same
// H x W) instead of fp32 NHWC rows (which enc_prep_kernel then pools and splits).
template <int CIN, int COUT, bool kPlanes>
__global__ void __launch_bounds__(kConvThreads, 1)
enc_conv_tc_kernel(const __half* __restrict__ in_hi, const __half* __restrict__ in_lo,
                   const uint8_t* __restrict__ wimg, const float* __restrict__ bias, float* __restrict__ out,
                   __half* __restrict__ out_hi, __half* __restrict__ out_lo, int H, int W, int n_tiles) {
  constexpr int kKB = CIN / 64;
  constexpr int kRingA = ring_a<COUT>();
  constexpr uint32_t kChunk = 2 * COUT * 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_ring = smem;
  uint8_t* s_rows = s_ring + kRingW * kChunk;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_rows + kRingA * kRowBytes);
  uint64_t* a_full = bars;                     // [kRingA]
  uint64_t* a_empty = a_full + kRingA;         // [kRingA]
  uint64_t* w_full = a_empty + kRingA;         // [kRingW]
  uint64_t* w_empty = w_full + kRingW;         // [kRingW]
  uint64_t* d_full = w_empty + kRingW;         // [2]
  uint64_t* d_empty = d_full + 2;              // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_empty + 2);
  __shared__ float s_bias[COUT];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Wp = W + 2;
  const long long plane_len = (long long)(H + 2) * Wp;
  if (threadIdx.x < COUT) s_bias[threadIdx.x] = bias[threadIdx.x];
  if (threadIdx.x == 0) {
    for (int s = 0; s < kRingA; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < kRingW; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&d_full[b], 1);
      mbar_init(&d_empty[b], 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<2 * COUT>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ---- A producer: one tap row (130 pixels x 8 channel chunks x {hi, lo}) per ring slot
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long q0 = (long long)Wp + 1 + (long long)kTilePix * tile;
        for (int kb = 0; kb < kKB; ++kb)
          for (int dy = 0; dy < 3; ++dy, ++it) {
            const long long start = q0 + (long long)(dy - 1) * Wp - 1;
            const long long room = plane_len - start;
            const uint32_t len = (uint32_t)(room < kRun ? room : kRun);
            const uint32_t s = it % kRingA;
            if (it >= kRingA) mbar_wait(&a_empty[s], (it / kRingA - 1) & 1, 11);
            mbar_arrive_expect_tx(&a_full[s], len * 16 * 8 * 2);
            uint8_t* dst = s_rows + s * kRowBytes;
            for (int c = 0; c < 8; ++c) {
              const size_t src = ((size_t)(kb * 8 + c) * plane_len + start) * 8;  // in halfs
              bulk_g2s(dst + c * kChunkStride, in_hi + src, len * 16, &a_full[s]);
              bulk_g2s(dst + kRowHalf + c * kChunkStride, in_lo + src, len * 16, &a_full[s]);
            }
          }
      }
    }
  } else if (warp == 1) {
    // ---- weight producer: one (channel block, tap) chunk per ring slot
    if (lane == 0) {
      const uint64_t policy = l2_policy_evict_last();
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
        for (int c = 0; c < kKB * 9; ++c, ++it) {
          const uint32_t s = it % kRingW;
          if (it >= kRingW) mbar_wait(&w_empty[s], (it / kRingW - 1) & 1, 12);
          mbar_arrive_expect_tx(&w_full[s], kChunk);
          bulk_g2s_hint(s_ring + s * kChunk, wimg + (size_t)c * kChunk, kChunk, &w_full[s], policy);
        }
    }
  } else if (warp == 2) {
    // ---- MMA issuer
    constexpr uint32_t idesc = make_idesc_f16(128, COUT, 0);
    uint32_t a_it = 0, w_it = 0, local = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++local) {
      const uint32_t buf = local & 1;
      if (local >= 2) mbar_wait(&d_empty[buf], (local / 2 - 1) & 1, 13);
      const uint32_t d_tmem = tmem + buf * COUT;
      uint32_t acc = 0;
      for (int kb = 0; kb < kKB; ++kb) {
        for (int tap = 0; tap < 9; ++tap, ++w_it) {
          const uint32_t sa = a_it % kRingA;
          if (tap % 3 == 0) mbar_wait(&a_full[sa], (a_it / kRingA) & 1, 14);
          const uint32_t s = w_it % kRingW;
          mbar_wait(&w_full[s], (w_it / kRingW) & 1, 15);
          tc_fence_after_sync();
          if (elect_one()) {
            const uint32_t a_hi = smem_u32(s_rows + sa * kRowBytes) + (tap % 3) * 16, a_lo = a_hi + kRowHalf;
            const uint32_t b_hi = smem_u32(s_ring + s * kChunk), b_lo = b_hi + COUT * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint64_t ah = make_sdesc_k_nosw(a_hi + (2 * j) * kChunkStride, kChunkStride, 128);
              const uint64_t al = make_sdesc_k_nosw(a_lo + (2 * j) * kChunkStride, kChunkStride, 128);
              const uint64_t bh = make_sdesc_k_sw128(b_hi + j * 32, 1024);
              const uint64_t bl = make_sdesc_k_sw128(b_lo + j * 32, 1024);
              umma_ss(d_tmem, ah, bh, idesc, acc);
              acc = 1;
              umma_ss(d_tmem, al, bh, idesc, 1);
              umma_ss(d_tmem, ah, bl, idesc, 1);
            }
            umma_commit(&w_empty[s]);
            if (tap % 3 == 2) umma_commit(&a_empty[sa]);
            if (tap == 8 && kb == kKB - 1) umma_commit(&d_full[buf]);
          }
          __syncwarp();
          if (tap % 3 == 2) ++a_it;
        }
      }
    }
  } else if (warp >= 4) {
    // ---- epilogue: + bias, LeakyReLU, store
    const int quarter = warp & 3;
    uint32_t local = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++local) {
      const uint32_t buf = local & 1;
      mbar_wait(&d_full[buf], (local / 2) & 1, 16);
      tc_fence_after_sync();
      const long long q = (long long)Wp + 1 + (long long)kTilePix * tile + quarter * 32 + lane;
      const int yp = (int)(q / Wp), xp = (int)(q - (long long)yp * Wp);
      const bool valid = xp >= 1 && xp <= W && yp <= H;
      float4* o4 = reinterpret_cast<float4*>(out + ((size_t)(yp - 1) * W + (xp - 1)) * COUT);
      const HaloTargets tg(yp, xp, H, W);
#pragma unroll 1
      for (int c0 = 0; c0 < COUT; c0 += 32) {
        uint32_t v[32];
        tmem_ld_x32(tmem + (static_cast<uint32_t>(quarter * 32) << 16) + buf * COUT + c0, v);
        tmem_ld_wait();
        if (valid) {
          if constexpr (kPlanes) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float f[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = lrelu(__uint_as_float(v[g * 8 + j]) + s_bias[c0 + g * 8 + j]);
              uint4 h, l;
              split8(f, h, l);
              store_plane_elem(out_hi, out_lo, H + 2, Wp, c0 / 8 + g, tg, h, l);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              o4[(c0 + j) / 4] = make_float4(lrelu(__uint_as_float(v[j]) + s_bias[c0 + j]),
                                             lrelu(__uint_as_float(v[j + 1]) + s_bias[c0 + j + 1]),
                                             lrelu(__uint_as_float(v[j + 2]) + s_bias[c0 + j + 2]),
                                             lrelu(__uint_as_float(v[j + 3]) + s_bias[c0 + j + 3]));
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&d_empty[buf]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc<2 * COUT>(tmem);
}

// ---- adaptive avg-pool to 32x32 + conv7 1x1 128->64 + LeakyReLU: F (H4,W4,128) -> out (64,32,32) ----
__global__ void __launch_bounds__(128)
enc_tail_kernel(const float* __restrict__ f, int H4, int W4, const float* __restrict__ blob, float* __restrict__ out) {
  __shared__ float pooled[128];
  const int bi = blockIdx.x / 32, bj = blockIdx.x % 32;
  // torch adaptive pooling bins: [floor(i*in/out), ceil((i+1)*in/out))
  const int y0 = (bi * H4) / 32, y1 = ((bi + 1) * H4 + 31) / 32;
  const int x0 = (bj * W4) / 32, x1 = ((bj + 1) * W4 + 31) / 32;
  float acc = 0.f;
  for (int y = y0; y < y1; ++y)
    for (int x = x0; x < x1; ++x) acc += f[((size_t)y * W4 + x) * 128 + threadIdx.x];
  pooled[threadIdx.x] = acc / (float)((y1 - y0) * (x1 - x0));
  __syncthreads();
  if (threadIdx.x < 64) {
    const float* w = blob + Blob::w7 + threadIdx.x * 128;
    float s = blob[Blob::b7 + threadIdx.x];
#pragma unroll 8
    for (int c = 0; c < 128; ++c) s = fmaf(pooled[c], w[c], s);
    out[(size_t)threadIdx.x * 1024 + blockIdx.x] = lrelu(s);
  }
}

template <int CIN, int COUT, bool kPlanes>
int launch_conv(const __half* planes, long long plane_len, const uint8_t* wimg, const float* bias, float* out,
                __half* out_planes, int H, int W, cudaStream_t st) {
  const long long span = (long long)(H - 1) * (W + 2) + W;
  const int n_tiles = (int)((span + kTilePix - 1) / kTilePix);
  constexpr int smem = conv_smem_bytes<COUT>();
  auto kern = enc_conv_tc_kernel<CIN, COUT, kPlanes>;
  CRNERF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int grid = std::min(n_tiles, num_sms());
  kern<<<grid, kConvThreads, smem, st>>>(planes, planes + (size_t)CIN * plane_len, wimg, bias, out, out_planes,
                                         out_planes ? out_planes + (size_t)COUT * plane_len : nullptr, H, W, n_tiles);
  count_launch();
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

int launch_prep(const float* f, int Wi, int C, int Ho, int Wo, bool pool, __half* planes, cudaStream_t st) {
  const long long plane_len = (long long)(Ho + 2) * (Wo + 2), total = plane_len * (C / 8);
  const int grid = (int)std::min<long long>((total + 255) / 256, 16LL * num_sms());
  __half* lo = planes + (size_t)C * plane_len;
  if (pool)
    enc_prep_kernel<true><<<grid, 256, 0, st>>>(f, Wi, C, Ho, Wo, planes, lo);
  else
    enc_prep_kernel<false><<<grid, 256, 0, st>>>(f, Wi, C, Ho, Wo, planes, lo);
  count_launch();
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

size_t align256(size_t b) { return (b + 255) & ~size_t(255); }
// scratch = [F: fp32 NHWC, largest H*W*64][Pa: planes, largest of (64, H, W) / (128, H/2, W/2)][Pb: planes (128, H/2, W/2)]
size_t f_bytes(int H, int W) { return align256((size_t)H * W * 64 * sizeof(float)); }
size_t pb_bytes(int H, int W) { return align256((size_t)2 * 128 * (H / 2 + 2) * (W / 2 + 2) * sizeof(__half)); }
size_t pa_bytes(int H, int W) {
  return std::max(align256((size_t)2 * 64 * (H + 2) * (W + 2) * sizeof(__half)), pb_bytes(H, W));
}

}  // namespace

size_t encoder_packed_bytes() {
  TcLayer L[4];
  size_t blob;
  tc_layers(L, blob);
  return blob + Blob::total * sizeof(float);
}

size_t encoder_scratch_bytes(int H, int W) { return f_bytes(H, W) + pa_bytes(H, W) + pb_bytes(H, W); }

int encoder_pack(const crnerf_encoder_weights* w, void* packed, size_t packed_bytes, cudaStream_t st) {
  CRNERF_REQUIRE(w && packed, "null argument");
  for (int i = 0; i < 7; ++i) CRNERF_REQUIRE(w->weight[i] && w->bias[i], "conv%d: null weight or bias", i + 1);
  CRNERF_REQUIRE(packed_bytes >= encoder_packed_bytes(), "packed buffer too small");
  TcLayer L[4];
  size_t blob;
  tc_layers(L, blob);
  uint8_t* img = static_cast<uint8_t*>(packed);
  for (int i = 0; i < 4; ++i)
    enc_pack_tc_kernel<<<2 * num_sms(), 256, 0, st>>>(w->weight[2 + i], L[i].cin, L[i].cout, img + L[i].offset);
  enc_pack_blob_kernel<<<32, 256, 0, st>>>(w->weight[0], w->bias[0], w->weight[1], w->bias[1], w->bias[2], w->bias[3],
                                           w->bias[4], w->bias[5], w->weight[6], w->bias[6],
                                           reinterpret_cast<float*>(img + blob));
  count_launch(5);
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

int encoder_forward(const void* packed, const float* img, int H, int W, float* out, void* scratch,
                    size_t scratch_bytes, cudaStream_t st) {
  CRNERF_REQUIRE(packed && img && out && scratch, "null argument");
  CRNERF_REQUIRE(H >= 8 && W >= 8 && H <= 8192 && W <= 8192, "image %dx%d unsupported (8..8192 per side)", H, W);
  CRNERF_REQUIRE(scratch_bytes >= encoder_scratch_bytes(H, W), "scratch too small");
  CRNERF_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 255) == 0 && (reinterpret_cast<uintptr_t>(packed) & 127) == 0,
                 "scratch must be 256-byte and packed 128-byte aligned");
  TcLayer L[4];
  size_t blob_off;
  tc_layers(L, blob_off);
  const uint8_t* wimg = static_cast<const uint8_t*>(packed);
  const float* blob = reinterpret_cast<const float*>(wimg + blob_off);
  float* F = static_cast<float*>(scratch);
  __half* Pa = reinterpret_cast<__half*>(static_cast<uint8_t*>(scratch) + f_bytes(H, W));
  __half* Pb = reinterpret_cast<__half*>(static_cast<uint8_t*>(scratch) + f_bytes(H, W) + pa_bytes(H, W));
  const int H2 = H / 2, W2 = W / 2, H4 = H2 / 2, W4 = W2 / 2;
  auto plane = [](int h, int w) { return (long long)(h + 2) * (w + 2); };
  int rc;

  enc_first_kernel<<<(unsigned)(((long long)H * W + 127) / 128), 128, 0, st>>>(img, H, W, blob, Pa,
                                                                               Pa + (size_t)64 * plane(H, W));
  count_launch();
  CRNERF_CUDA(cudaGetLastError());
  if ((rc = launch_conv<64, 64, false>(Pa, plane(H, W), wimg + L[0].offset, blob + Blob::b3, F, nullptr, H, W, st)))
    return rc;
  if ((rc = launch_prep(F, W, 64, H2, W2, true, Pa, st))) return rc;
  if ((rc = launch_conv<64, 128, true>(Pa, plane(H2, W2), wimg + L[1].offset, blob + Blob::b4, nullptr, Pb, H2, W2,
                                       st)))
    return rc;
  if ((rc = launch_conv<128, 128, false>(Pb, plane(H2, W2), wimg + L[2].offset, blob + Blob::b5, F, nullptr, H2, W2,
                                         st)))
    return rc;
  if ((rc = launch_prep(F, W2, 128, H4, W4, true, Pa, st))) return rc;
  if ((rc = launch_conv<128, 128, false>(Pa, plane(H4, W4), wimg + L[3].offset, blob + Blob::b6, F, nullptr, H4, W4,
                                         st)))
    return rc;
  enc_tail_kernel<<<1024, 128, 0, st>>>(F, H4, W4, blob, out);
  count_launch();
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

}  // namespace crnerf

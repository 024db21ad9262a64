// encoder_sameoutputsize.forward (reference models/linearStyleTransfer.py:208-276; SURVEY.md 8f
// rank 1): the style/content encoder enc_a / enc_cont.  Inference below; the training step's forward and
// backward at the end of the file (kernels: encoder_train.cuh).
//
//   conv1 1x1 3->3 . reflect-pad                                         enc_conv1_planes_kernel (fp32)
//   conv2 3x3 3->64 . LeakyReLU                                          enc_conv_tc_kernel<8, 64> (two taps per K = 16 step)
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
// Columns past the row end compute garbage (whatever follows in the flattened plane) that is not stored.
// The producer stages one input row per ring slot; a step computes two vertically adjacent
// 128-pixel tiles that share every weight chunk and three of their four input rows.  The epilogue
// applies bias, LeakyReLU and - for conv3 / conv5 - the 2x2 max-pool in registers, and writes the
// next layer's planes (halo included) directly: no fp32 activation ever goes to HBM except
// conv6's small output, which the tail kernel pools.
#include <cuda_fp16.h>
#include <algorithm>
#include <cstdlib>
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
// H x W layer: itself plus its reflection-halo copies (row 0 mirrors row 2, row H+1 mirrors H-1,
// same for columns).  Flags, not index arrays: the common interior case is one predicated-off branch.
struct HaloTargets {
  int yp, xp, H, W;
  bool top, bot, left, right, any;
  __device__ __forceinline__ HaloTargets(int yp_, int xp_, int H_, int W_) : yp(yp_), xp(xp_), H(H_), W(W_) {
    top = yp == 2;
    bot = yp == H - 1;
    left = xp == 2;
    right = xp == W - 1;
    any = top || bot || left || right;
  }
};
__device__ __forceinline__ void store_plane_at(__half* hi, __half* lo, size_t plane_base, int Wp, int y, int x,
                                               const uint4& h, const uint4& l) {
  const size_t o = plane_base + ((size_t)y * Wp + x) * 8;
  *reinterpret_cast<uint4*>(hi + o) = h;
  *reinterpret_cast<uint4*>(lo + o) = l;
}
__device__ __forceinline__ void store_plane_elem(__half* hi, __half* lo, int Hp, int Wp, int chunk,
                                                 const HaloTargets& t, const uint4& h, const uint4& l) {
  const size_t pb = (size_t)chunk * Hp * Wp * 8;
  store_plane_at(hi, lo, pb, Wp, t.yp, t.xp, h, l);
  if (t.any) {
    if (t.left) store_plane_at(hi, lo, pb, Wp, t.yp, 0, h, l);
    if (t.right) store_plane_at(hi, lo, pb, Wp, t.yp, t.W + 1, h, l);
    if (t.top) {
      store_plane_at(hi, lo, pb, Wp, 0, t.xp, h, l);
      if (t.left) store_plane_at(hi, lo, pb, Wp, 0, 0, h, l);
      if (t.right) store_plane_at(hi, lo, pb, Wp, 0, t.W + 1, h, l);
    }
    if (t.bot) {
      store_plane_at(hi, lo, pb, Wp, t.H + 1, t.xp, h, l);
      if (t.left) store_plane_at(hi, lo, pb, Wp, t.H + 1, 0, h, l);
      if (t.right) store_plane_at(hi, lo, pb, Wp, t.H + 1, t.W + 1, h, l);
    }
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

// ---- conv1 (1x1, 3 -> 3) + reflection halo: img (3,H,W) -> planes (one 8-channel chunk: 3 real + 5 zero) -----
__global__ void __launch_bounds__(256)
enc_conv1_planes_kernel(const float* __restrict__ img, int H, int W, const float* __restrict__ blob,
                        __half* __restrict__ out_hi, __half* __restrict__ out_lo) {
  const int Wp = W + 2;
  const long long total = (long long)(H + 2) * Wp;
  const float w0 = blob[Blob::w1], w1 = blob[Blob::w1 + 1], w2 = blob[Blob::w1 + 2], w3 = blob[Blob::w1 + 3],
              w4 = blob[Blob::w1 + 4], w5 = blob[Blob::w1 + 5], w6 = blob[Blob::w1 + 6], w7 = blob[Blob::w1 + 7],
              w8 = blob[Blob::w1 + 8], b0 = blob[Blob::b1], b1 = blob[Blob::b1 + 1], b2 = blob[Blob::b1 + 2];
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    const int yp = (int)(q / Wp), xp = (int)(q - (long long)yp * Wp);
    const size_t o = (size_t)reflect_idx(yp - 1, H) * W + reflect_idx(xp - 1, W);
    const float v0 = img[o], v1 = img[(size_t)H * W + o], v2 = img[(size_t)2 * H * W + o];
    const float v[8] = {b0 + (w0 * v0 + w1 * v1 + w2 * v2), b1 + (w3 * v0 + w4 * v1 + w5 * v2),
                        b2 + (w6 * v0 + w7 * v1 + w8 * v2), 0.f, 0.f, 0.f, 0.f, 0.f};
    uint4 h, l;
    split8(v, h, l);
    *reinterpret_cast<uint4*>(out_hi + (size_t)q * 8) = h;
    *reinterpret_cast<uint4*>(out_lo + (size_t)q * 8) = l;
  }
}

// conv2 weights (64,3,3,3) -> 3 chunks (one per tap row dy) of [hi: 64 rows x 128 B, SW128][lo: same];
// K index = dx * 8 + channel for dx = 0..3 (dx = 3 and channels 3..7 are zero), the rest of the row unused
__global__ void enc_pack_first_kernel(const float* __restrict__ w, uint8_t* __restrict__ img) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 3 * 64 * 64; i += gridDim.x * blockDim.x) {
    const int k = i % 64, co = (i / 64) % 64, dy = i / 4096;
    const int dx = k >> 3, ch = k & 7;
    const float v = (dx < 3 && ch < 3) ? w[((co * 3 + ch) * 3 + dy) * 3 + dx] : 0.f;
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    uint8_t* base = img + (size_t)dy * (2 * 64 * 128) + sw128_offset(co, k >> 3) + (k & 7) * 2;
    *reinterpret_cast<__half*>(base) = hi;
    *reinterpret_cast<__half*>(base + 64 * 128) = lo;
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
constexpr int kRowHalf = 8 * kChunkStride; // one input row, one of {hi, lo}: 8 channel chunks x 130 pixels
constexpr int kRowBytes = 2 * kRowHalf;    // hi then lo
constexpr int kRingW = 3;                  // weight chunks in flight
constexpr int kConvThreads = 384;         // producers + issuer (warps 0-3), 8 epilogue warps (lane quarter x column half)

// input rows in flight: a tap row pair (dy, dy+1) is live, every further slot is prefetch
template <int COUT>
__host__ __device__ constexpr int ring_a() {
  return COUT == 64 ? 5 : 4;
}
template <int COUT>
constexpr int conv_smem_bytes() {
  return kRingW * 2 * COUT * 128 + ring_a<COUT>() * kRowBytes + 256 + COUT * 4;  // rings, barriers, bias
}

enum ConvOut { kOutRows = 0, kOutPlanes = 1, kOutPoolPlanes = 2, kOutRaw = 3 };

// One step = a "band tile": 128 consecutive positions of the flattened (band, padded column) grid,
// for output rows 2b and 2b+1 -> two accumulators that share every weight chunk (half the L2 weight
// traffic per pixel) and three of their four input rows.  Runs cross band ends, so layers of any
// width fill the 128-row MMA densely (W / (W+2)); halo columns compute garbage that is not stored.
// kOut selects the epilogue:
//   kOutRows        fp32 NHWC rows (H, W, COUT)                          (conv6 -> tail kernel)
//   kOutPlanes      the next layer's hi/lo planes with reflection halo    (conv4 -> conv5)
//   kOutPoolPlanes  2x2 max-pool in registers (vertical: the two accumulators, horizontal: the
//                   neighbour lane), then planes of the pooled layer       (conv3, conv5)
//   kOutRaw         fp32 NHWC rows without bias / activation, and the launch's max |value| folded into
//                   *maxbits (the backward's input-gradient convolutions, see "training" below)
template <int CIN, int COUT, int kOut>
__global__ void __launch_bounds__(kConvThreads, 1)
enc_conv_tc_kernel(const __half* __restrict__ in_hi, const __half* __restrict__ in_lo,
                   const uint8_t* __restrict__ wimg, const float* __restrict__ bias, float* __restrict__ out,
                   __half* __restrict__ out_hi, __half* __restrict__ out_lo, int H, int W, int Ws, int n_tiles,
                   unsigned* __restrict__ maxbits) {
  // CIN == 8: the first 3x3 convolution (conv2, 3 real channels padded to one 8-channel chunk).  A K = 16
  // step then spans TWO horizontally adjacent taps: the descriptor's K-chunk stride (LBO) is 16 B, i.e.
  // the next pixel, so a tap row is 2 steps (dx = 0,1 | 2, and a phantom dx = 3 whose weights are zero).
  constexpr bool kFirst = CIN == 8;
  constexpr int kKB = kFirst ? 1 : CIN / 64;
  constexpr int kTaps = kFirst ? 3 : 9;          // weight chunks per channel block
  constexpr int kChunks = kFirst ? 1 : 8;        // 8-channel chunks per input row
  constexpr int kRunLoad = kFirst ? kRun + 1 : kRun;
  constexpr int kRingA = ring_a<COUT>();
  constexpr uint32_t kChunk = 2 * COUT * 128;
  // COUT == 64: the hi and lo weight images of a chunk are adjacent, i.e. one 128-row operand, so
  // hi*[hi; lo] is ONE N = 128 MMA whose two column halves the epilogue adds (an N = 64 MMA is bound
  // by its shared-memory operand reads, 6 KB per 32 tensor cycles); lo*hi stays an N = 64 MMA.
  constexpr bool kStack = COUT == 64;
  constexpr uint32_t kDCols = kStack ? 128 : COUT;   // TMEM columns per accumulator
  extern __shared__ __align__(1024) uint8_t smem[];
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
  float* s_bias = reinterpret_cast<float*>(s_rows + kRingA * kRowBytes + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Wp = W + 2;
  const long long plane_len = (long long)(H + 2) * Wp;
  if (threadIdx.x < COUT) s_bias[threadIdx.x] = bias[threadIdx.x];
  if constexpr (kFirst) {
    // the phantom tap multiplies whatever the row slot holds by zero weights: it must be finite
    for (int i = threadIdx.x; i < kRingA * kRowBytes / 16; i += kConvThreads)
      reinterpret_cast<uint4*>(s_rows)[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
  }
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
      mbar_init(&d_empty[b], 8);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<4 * kDCols>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ---- A producer: for a step's 130 flat entries, input rows 2b .. 2b+3 (padded) of every band b the
    // run touches, one row index per ring slot (a run that crosses a band end is filled in pieces)
    // Lane l < 16 issues the copy of (channel chunk l & 7, hi / lo = l >> 3): a row slot is 16-32 small
    // bulk copies, and one thread issuing them all was slower than the MMAs that consume them.
    {
      uint32_t it = 0;
      const int c = lane % kChunks;
      const bool lo_half = (lane / kChunks) & 1;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long i_lo = (long long)kTilePix * tile;   // flat index of smem entry 0
        const int b_lo = (int)(i_lo / Ws), b_hi = (int)((i_lo + kRunLoad - 1) / Ws);
        for (int kb = 0; kb < kKB; ++kb)
          for (int r = 0; r < 4; ++r, ++it) {
            const uint32_t s = it % kRingA;
            if (it >= kRingA) mbar_wait(&a_empty[s], (it / kRingA - 1) & 1, 11);
            uint32_t total = 0;
            for (int b = b_lo; b <= b_hi; ++b) {
              const long long lo = max(i_lo, (long long)b * Ws), hi = min(i_lo + kRunLoad, (long long)b * Ws + Wp);
              if (hi > lo && 2 * b + r <= H + 1) total += (uint32_t)(hi - lo);
            }
            if (lane == 0) mbar_arrive_expect_tx(&a_full[s], total * 16 * kChunks * 2);
            __syncwarp();
            if (lane < 2 * kChunks) {
              uint8_t* dst = s_rows + s * kRowBytes + (lo_half ? kRowHalf : 0) + c * kChunkStride;
              const __half* srcp = lo_half ? in_lo : in_hi;
              for (int b = b_lo; b <= b_hi; ++b) {
                const long long lo = max(i_lo, (long long)b * Ws), hi = min(i_lo + kRunLoad, (long long)b * Ws + Wp);
                if (hi <= lo || 2 * b + r > H + 1) continue;
                const uint32_t bytes = (uint32_t)(hi - lo) * 16, e = (uint32_t)(lo - i_lo) * 16;
                const long long src_px = (long long)(2 * b + r) * Wp + (lo - (long long)b * Ws);
                bulk_g2s(dst + e, srcp + ((size_t)(kb * 8 + c) * plane_len + src_px) * 8, bytes, &a_full[s]);
              }
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
        for (int c = 0; c < kKB * kTaps; ++c, ++it) {
          const uint32_t s = it % kRingW;
          if (it >= kRingW) mbar_wait(&w_empty[s], (it / kRingW - 1) & 1, 12);
          mbar_arrive_expect_tx(&w_full[s], kChunk);
          bulk_g2s_hint(s_ring + s * kChunk, wimg + (size_t)c * kChunk, kChunk, &w_full[s], policy);
        }
    }
  } else if (warp == 2) {
    // ---- MMA issuer
    constexpr uint32_t idesc = make_idesc_f16(128, COUT, 0);
    uint32_t a_it = 0, w_it = 0, local = 0;  // a_it: sequence number of input row 0 of the current (tile, kb)
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++local) {
      const uint32_t buf = local & 1;
      if (local >= 2) mbar_wait(&d_empty[buf], (local / 2 - 1) & 1, 13);
      const uint32_t d0 = tmem + buf * 2 * kDCols, d1 = d0 + kDCols;
      uint32_t acc = 0;
      for (int kb = 0; kb < kKB; ++kb, a_it += 4) {
        for (int tap = 0; tap < kTaps; ++tap, ++w_it) {
          const int dy = kFirst ? tap : tap / 3, dx = kFirst ? 0 : tap - dy * 3;
          const uint32_t r0 = a_it + dy, r1 = r0 + 1;
          if (dx == 0) {
            if (dy == 0) mbar_wait(&a_full[r0 % kRingA], (r0 / kRingA) & 1, 14);
            mbar_wait(&a_full[r1 % kRingA], (r1 / kRingA) & 1, 14);
          }
          const uint32_t s = w_it % kRingW;
          mbar_wait(&w_full[s], (w_it / kRingW) & 1, 15);
          tc_fence_after_sync();
          if (elect_one()) {
            // descriptors differ only in the 14-bit address field: build one per operand and add offsets
            constexpr uint32_t kLbo = kFirst ? 16 : kChunkStride;   // K-adjacent chunk: next pixel | next channel chunk
            const uint64_t a0h = make_sdesc_k_nosw(smem_u32(s_rows + (r0 % kRingA) * kRowBytes) + dx * 16, kLbo, 128);
            const uint64_t a1h = make_sdesc_k_nosw(smem_u32(s_rows + (r1 % kRingA) * kRowBytes) + dx * 16, kLbo, 128);
            const uint64_t bh = make_sdesc_k_sw128(smem_u32(s_ring + s * kChunk), 1024);
            constexpr uint64_t kLo = kRowHalf >> 4, kBLo = (COUT * 128) >> 4;
            // all MMAs of one accumulator back to back, then the other: alternating accumulators
            // instruction by instruction costs ~110 cycles per MMA instead of 64 / 48
            const uint32_t acc0 = acc;
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              const uint32_t d = t ? d1 : d0;
              const uint64_t ah = t ? a1h : a0h;
#pragma unroll
              for (int j = 0; j < (kFirst ? 2 : 4); ++j) {
                const uint64_t ja = (uint64_t)((2 * j * kLbo) >> 4), jb = (uint64_t)((j * 32) >> 4);
                if constexpr (kStack) {
                  constexpr uint32_t idesc2 = make_idesc_f16(128, 128, 0);
                  umma_ss(d, ah + ja, bh + jb, idesc2, j ? 1u : acc0);   // hi * [hi; lo]
                  umma_ss(d, ah + kLo + ja, bh + jb, idesc, 1);          // lo * hi
                } else {
                  umma_ss(d, ah + ja, bh + jb, idesc, j ? 1u : acc0);
                  umma_ss(d, ah + kLo + ja, bh + jb, idesc, 1);
                  umma_ss(d, ah + ja, bh + kBLo + jb, idesc, 1);
                }
              }
            }
            acc = 1;
            umma_commit(&w_empty[s]);
            if (dx == 2 || kFirst) {
              umma_commit(&a_empty[r0 % kRingA]);                 // row dy is done after tap row dy
              if (dy == 2) umma_commit(&a_empty[r1 % kRingA]);    // and row 3 with it
            }
            if (tap == kTaps - 1 && kb == kKB - 1) umma_commit(&d_full[buf]);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp >= 4) {
    // ---- epilogue: + bias, LeakyReLU, (pool,) store.  Two warps per lane quarter, each taking every
    // other 32-channel block: with Cout = 64 a step is only ~8 k tensor cycles and one warp per
    // quarter (a single warp on its scheduler, nothing to hide latency) could not keep up.
    const int quarter = warp & 3, chalf = (warp - 4) >> 2;
    uint32_t local = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++local) {
      const uint32_t buf = local & 1;
      mbar_wait(&d_full[buf], (local / 2) & 1, 16);
      tc_fence_after_sync();
      const long long fi = 1 + (long long)kTilePix * tile + quarter * 32 + lane;   // flat index = band * Ws + padded column
      const int band = (int)(fi / Ws), xp = (int)(fi - (long long)band * Ws);
      const int x = xp - 1;                                // interior column; even x <-> even lane (Ws is even)
      const int y = 2 * band;                              // interior row of accumulator 0
      const bool in_row = xp >= 1 && xp <= W;
      const uint32_t t0 = tmem + (static_cast<uint32_t>(quarter * 32) << 16) + buf * 2 * kDCols;
      if constexpr (kOut == kOutPoolPlanes) {
        const int Ho = H / 2, Wo = W / 2, xo = x >> 1;
        const bool valid = in_row && band < Ho && xo < Wo && (x & 1) == 0;
        const HaloTargets tg(band + 1, xo + 1, Ho, Wo);
#pragma unroll 1
        for (int c0 = chalf * 32; c0 < COUT; c0 += 64) {
          uint32_t v[32], u[32];
          tmem_ld_x32(t0 + c0, v);
          tmem_ld_x32(t0 + kDCols + c0, u);
          tmem_ld_wait();
          if constexpr (kStack) {   // add the hi*lo column half
            uint32_t v2[32], u2[32];
            tmem_ld_x32(t0 + 64 + c0, v2);
            tmem_ld_x32(t0 + kDCols + 64 + c0, u2);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
              u[j] = __float_as_uint(__uint_as_float(u[j]) + __uint_as_float(u2[j]));
            }
          }
          float m[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float a = fmaxf(__uint_as_float(v[j]), __uint_as_float(u[j]));
            m[j] = lrelu(fmaxf(a, __shfl_xor_sync(0xffffffffu, a, 1)) + s_bias[c0 + j]);
          }
          if (valid) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float f[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = m[g * 8 + j];
              uint4 h, l;
              split8(f, h, l);
              store_plane_elem(out_hi, out_lo, Ho + 2, Wo + 2, c0 / 8 + g, tg, h, l);
            }
          }
        }
      } else {
        float vmax = 0.f;
#pragma unroll 1
        for (int r = 0; r < 2; ++r) {
          const bool valid = in_row && y + r < H;
          const HaloTargets tg(y + r + 1, x + 1, H, W);
          float4* o4 = reinterpret_cast<float4*>(out + ((size_t)(y + r) * W + x) * COUT);
#pragma unroll 1
          for (int c0 = chalf * 32; c0 < COUT; c0 += 64) {
            uint32_t v[32];
            tmem_ld_x32(t0 + r * kDCols + c0, v);
            tmem_ld_wait();
            if constexpr (kStack) {   // add the hi*lo column half
              uint32_t v2[32];
              tmem_ld_x32(t0 + r * kDCols + 64 + c0, v2);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
            }
            if (valid) {
              if constexpr (kOut == kOutPlanes) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                  float f[8];
#pragma unroll
                  for (int j = 0; j < 8; ++j) f[j] = lrelu(__uint_as_float(v[g * 8 + j]) + s_bias[c0 + g * 8 + j]);
                  uint4 h, l;
                  split8(f, h, l);
                  store_plane_elem(out_hi, out_lo, H + 2, Wp, c0 / 8 + g, tg, h, l);
                }
              } else if constexpr (kOut == kOutRaw) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  const float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                               __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                  vmax = fmaxf(fmaxf(vmax, fmaxf(fabsf(o.x), fabsf(o.y))), fmaxf(fabsf(o.z), fabsf(o.w)));
                  o4[(c0 + j) / 4] = o;
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
        }
        if constexpr (kOut == kOutRaw) {
          // non-negative floats order like their bit patterns: one atomic per warp and tile
#pragma unroll
          for (int o = 16; o; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
          if (lane == 0 && vmax > 0.f) atomicMax(maxbits, __float_as_uint(vmax));
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&d_empty[buf]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc<4 * kDCols>(tmem);
}

// ---- adaptive avg-pool to 32x32 + conv7 1x1 128->64 + LeakyReLU: F (H4,W4,128) -> out (64,32,32) ----
__global__ void __launch_bounds__(512)
enc_tail_kernel(const float* __restrict__ f, int H4, int W4, const float* __restrict__ blob, float* __restrict__ out,
                float* __restrict__ pooled_out) {
  __shared__ float part[4][128];
  __shared__ float pooled[128];
  const int bi = blockIdx.x / 32, bj = blockIdx.x % 32;
  // torch adaptive pooling bins: [floor(i*in/out), ceil((i+1)*in/out))
  const int y0 = (bi * H4) / 32, y1 = ((bi + 1) * H4 + 31) / 32;
  const int x0 = (bj * W4) / 32, x1 = ((bj + 1) * W4 + 31) / 32;
  const int c = threadIdx.x & 127, g = threadIdx.x >> 7;   // 4 pixel groups x 128 channels
  const int bw = x1 - x0, n = (y1 - y0) * bw;
  float acc = 0.f;
  for (int p = g; p < n; p += 4) {
    const int y = y0 + p / bw, x = x0 + p % bw;
    acc += f[((size_t)y * W4 + x) * 128 + c];
  }
  part[g][c] = acc;
  __syncthreads();
  if (threadIdx.x < 128) {
    pooled[c] = ((part[0][c] + part[1][c]) + (part[2][c] + part[3][c])) / (float)n;
    if (pooled_out) pooled_out[(size_t)blockIdx.x * 128 + c] = pooled[c];   // training: conv7's input for its weight gradient
  }
  __syncthreads();
  // conv7: 64 outputs x 128 inputs; 8 threads per output, 16 inputs each
  const int co = threadIdx.x >> 3, k0 = (threadIdx.x & 7) * 16;
  const float* w = blob + Blob::w7 + co * 128 + k0;
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 16; ++k) s = fmaf(pooled[k0 + k], w[k], s);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  if ((threadIdx.x & 7) == 0) out[(size_t)co * 1024 + blockIdx.x] = lrelu(s + blob[Blob::b7 + co]);
}

template <int CIN, int COUT, int kOut>
int launch_conv(const __half* planes, const uint8_t* wimg, const float* bias, float* out, __half* out_planes,
                long long out_plane_len, int H, int W, cudaStream_t st, unsigned* maxbits = nullptr) {
  const long long plane_len = (long long)(H + 2) * (W + 2);
  // flat index space: band b (output rows 2b, 2b+1) x padded column, band stride Ws (even, so that
  // 2x2 pool partners share a lane pair); tiles are runs of 128 consecutive indices starting at 1
  const int Ws = (W + 3) & ~1;
  const int n_tiles = (int)((((long long)(H + 1) / 2) * Ws - 1 + kTilePix - 1) / kTilePix);
  constexpr int smem = conv_smem_bytes<COUT>();
  auto kern = enc_conv_tc_kernel<CIN, COUT, kOut>;
  CRNERF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int grid = std::min(n_tiles, num_sms());
  kern<<<grid, kConvThreads, smem, st>>>(planes, planes + (size_t)CIN * plane_len, wimg, bias, out, out_planes,
                                         out_planes ? out_planes + (size_t)COUT * out_plane_len : nullptr, H, W,
                                         Ws, n_tiles, maxbits);
  count_launch();
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

size_t align256(size_t b) { return (b + 255) & ~size_t(255); }
// scratch = [F: conv6 output, fp32 NHWC (H/4, W/4, 128)][Pa: planes (64, H, W), later (128, H/4, W/4)]
//           [Pb: planes (64, H/2, W/2), later (128, H/2, W/2) at +Pb2]
size_t planes_bytes(int C, int h, int w) { return align256((size_t)2 * C * (h + 2) * (w + 2) * sizeof(__half)); }
size_t f_bytes(int H, int W) { return align256((size_t)(H / 4) * (W / 4) * 128 * sizeof(float)); }
size_t pa_bytes(int H, int W) { return std::max(planes_bytes(64, H, W), planes_bytes(128, H / 4, W / 4)); }
size_t pb_bytes(int H, int W) { return planes_bytes(64, H / 2, W / 2); }
size_t pc_bytes(int H, int W) { return planes_bytes(128, H / 2, W / 2); }

#include "encoder_train.cuh"

}  // namespace

// conv2's tensor-core image (3 chunks of 16 KB) follows the fp32 blob, 128-byte aligned
static size_t first_image_offset(size_t blob_off) { return (blob_off + Blob::total * sizeof(float) + 127) & ~size_t(127); }
constexpr size_t kFirstImageBytes = 3 * 2 * 64 * 128;

// training: weight images of the four input-gradient convolutions (conv3..conv6 with channels swapped and
// taps flipped) follow conv2's image; same sizes as the forward images
static size_t dgrad_image_offset(size_t blob_off, int i) {
  TcLayer L[4];
  size_t b;
  tc_layers(L, b);
  size_t off = (first_image_offset(blob_off) + kFirstImageBytes + 127) & ~size_t(127);
  for (int j = 0; j < i; ++j) off += L[j].bytes();
  return off;
}

size_t encoder_packed_bytes() {
  TcLayer L[4];
  size_t blob;
  tc_layers(L, blob);
  return dgrad_image_offset(blob, 4);
}

size_t encoder_scratch_bytes(int H, int W) {
  return f_bytes(H, W) + pa_bytes(H, W) + pb_bytes(H, W) + pc_bytes(H, W);
}

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
  enc_pack_first_kernel<<<48, 256, 0, st>>>(w->weight[1], img + first_image_offset(blob));
  for (int i = 0; i < 4; ++i)
    enc_pack_tc_dgrad_kernel<<<2 * num_sms(), 256, 0, st>>>(w->weight[2 + i], L[i].cin, L[i].cout,
                                                            img + dgrad_image_offset(blob, i));
  count_launch(10);
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
  uint8_t* base = static_cast<uint8_t*>(scratch);
  float* F = reinterpret_cast<float*>(base);
  __half* Pa = reinterpret_cast<__half*>(base + f_bytes(H, W));
  __half* Pb = reinterpret_cast<__half*>(base + f_bytes(H, W) + pa_bytes(H, W));
  __half* Pc = reinterpret_cast<__half*>(base + f_bytes(H, W) + pa_bytes(H, W) + pb_bytes(H, W));
  const int H2 = H / 2, W2 = W / 2, H4 = H2 / 2, W4 = W2 / 2;
  auto plane = [](int h, int w) { return (long long)(h + 2) * (w + 2); };
  int rc;

  // conv1 + halo -> P0 (one 8-channel chunk, H, W; lives in Pb's region, dead before conv3 writes Pb)
  __half* P0 = Pb;
  {
    const long long total = plane(H, W);
    const int grid = (int)std::min<long long>((total + 255) / 256, 16LL * num_sms());
    enc_conv1_planes_kernel<<<grid, 256, 0, st>>>(img, H, W, blob, P0, P0 + (size_t)8 * total);
    count_launch();
    CRNERF_CUDA(cudaGetLastError());
  }
  // conv2: P0 -> Pa (64, H, W)
  if ((rc = launch_conv<8, 64, kOutPlanes>(P0, wimg + first_image_offset(blob_off), blob + Blob::b2, nullptr, Pa,
                                           plane(H, W), H, W, st)))
    return rc;
  // conv3 + pool: Pa -> Pb (64, H/2, W/2)
  if ((rc = launch_conv<64, 64, kOutPoolPlanes>(Pa, wimg + L[0].offset, blob + Blob::b3, nullptr, Pb, plane(H2, W2), H,
                                                W, st)))
    return rc;
  // conv4: Pb -> Pc (128, H/2, W/2)
  if ((rc = launch_conv<64, 128, kOutPlanes>(Pb, wimg + L[1].offset, blob + Blob::b4, nullptr, Pc, plane(H2, W2), H2,
                                             W2, st)))
    return rc;
  // conv5 + pool: Pc -> Pa (128, H/4, W/4)
  if ((rc = launch_conv<128, 128, kOutPoolPlanes>(Pc, wimg + L[2].offset, blob + Blob::b5, nullptr, Pa, plane(H4, W4),
                                                  H2, W2, st)))
    return rc;
  // conv6: Pa -> F (H/4, W/4, 128) fp32
  if ((rc = launch_conv<128, 128, kOutRows>(Pa, wimg + L[3].offset, blob + Blob::b6, F, nullptr, 0, H4, W4, st)))
    return rc;
  enc_tail_kernel<<<1024, 512, 0, st>>>(F, H4, W4, blob, out, nullptr);
  count_launch();
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

}  // namespace crnerf

// ======================================================================================================
// training: forward that keeps what the backward needs, and the backward (kernels: encoder_train.cuh)
// ======================================================================================================
namespace crnerf {
namespace {

size_t tape_planes(int C, int h, int w) { return align256((size_t)2 * C * (h + 2) * (w + 2) * sizeof(__half) + 256); }
struct Tape {   // byte offsets: activation planes kept by the training forward (+256 B: a K step may read past a row)
  size_t p0, a2, a3, q3, a4, a5, q5, f, pooled, total;
};
Tape tape_layout(int H, int W) {
  const int H2 = H / 2, W2 = W / 2, H4 = H2 / 2, W4 = W2 / 2;
  Tape t;
  size_t o = 0;
  t.p0 = o; o += tape_planes(8, H, W);        // conv1 output (3 real channels)
  t.a2 = o; o += tape_planes(64, H, W);       // LeakyReLU(conv2)
  t.a3 = o; o += tape_planes(64, H, W);       // LeakyReLU(conv3), before the pool
  t.q3 = o; o += tape_planes(64, H2, W2);     // pooled
  t.a4 = o; o += tape_planes(128, H2, W2);
  t.a5 = o; o += tape_planes(128, H2, W2);    // before the pool
  t.q5 = o; o += tape_planes(128, H4, W4);
  t.f = o; o += f_bytes(H, W);                // LeakyReLU(conv6), fp32 NHWC
  t.pooled = o; o += align256((size_t)1024 * 128 * sizeof(float));
  t.total = o;
  return t;
}

size_t grad_planes(int C, int h, int w) { return align256((size_t)2 * C * (h + 4) * grad_stride(w) * sizeof(__half) + 256); }
size_t dx_rows(int C, int h, int w) { return align256((size_t)(h + 2) * (grad_stride(w) - 2) * C * sizeof(float)); }
constexpr int kPrepBlocks = 296;     // x-dimension of the elementwise backward grids (bias partials per chunk)
constexpr int kFirstBlocks = 296;
struct BwdScratch {
  size_t ga, gb, dx, part, dpooled, dpre7, dbpart, part2, small, total;
};
BwdScratch bwd_layout(int H, int W) {
  const int H2 = H / 2, W2 = W / 2, H4 = H2 / 2, W4 = W2 / 2;
  BwdScratch s;
  size_t o = 0;
  s.ga = o; o += std::max({grad_planes(128, H4, W4), grad_planes(128, H2, W2), grad_planes(64, H, W)});
  s.gb = o; o += std::max(grad_planes(128, H2, W2), grad_planes(64, H, W));
  s.dx = o; o += std::max({dx_rows(128, H4, W4), dx_rows(128, H2, W2), dx_rows(64, H2, W2), dx_rows(64, H, W)});
  s.part = o; o += align256(std::max((size_t)148 * 3 * 128 * 128, (size_t)kFirstBlocks * 1728) * sizeof(float));
  s.dpooled = o; o += align256((size_t)1024 * 128 * sizeof(float));
  s.dpre7 = o; o += align256((size_t)1024 * 64 * sizeof(float));
  s.dbpart = o; o += align256((size_t)16 * 4 * kPrepBlocks * 8 * sizeof(float));
  s.part2 = o; o += align256((size_t)kFirstBlocks * 12 * sizeof(float));
  s.small = o; o += 256;                       // scales[8] floats | maxbits[8]
  s.total = o;
  return s;
}

template <int CIN, int COUT>
int launch_wgrad(const __half* G, const __half* X, int h, int w, float* part, const float* scale, float* gw,
                 cudaStream_t st) {
  using Cfg = WgradCfg<CIN, COUT>;
  const int Wg = grad_stride(w);
  const long long g_plane = (long long)(h + 4) * Wg, x_plane = (long long)(h + 2) * (w + 2);
  const int nseg = h * ((w + Cfg::kSeg - 1) / Cfg::kSeg);
  const int grid = 3 * std::min(std::min(num_sms(), 148) / 3, nseg);
  auto kern = enc_wgrad_tc_kernel<CIN, COUT>;
  CRNERF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
  kern<<<grid, kWgradThreads, Cfg::kSmem, st>>>(G, G + (size_t)COUT * g_plane, g_plane, Wg, X,
                                                X + (size_t)CIN * x_plane, x_plane, h, w, part);
  enc_wgrad_reduce_kernel<CIN, COUT><<<(9 * COUT * CIN + 255) / 256, 256, 0, st>>>(part, grid, scale, gw);
  count_launch(2);
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

int launch_prep(bool pool, const float* dx, int C, int hq, int wq, const __half* act, int H, int W, __half* G,
                const unsigned* maxbits_in, const float* scale_in, float* scale_out, float* dbpart, float* gb,
                cudaStream_t st) {
  PrepArgs a;
  a.dx = dx;
  a.wdx = grad_stride(wq) - 2;
  a.C = C;
  a.hq = hq;
  a.wq = wq;
  a.act_hi = act;
  a.act_lo = act + (size_t)C * (H + 2) * (W + 2);
  a.H = H;
  a.W = W;
  a.Wg = grad_stride(W);
  a.g_hi = G;
  a.g_lo = G + (size_t)C * (H + 4) * a.Wg;
  a.maxbits_in = maxbits_in;
  a.scale_in = scale_in;
  a.scale_out = scale_out;
  a.db_part = dbpart;
  CRNERF_CUDA(cudaMemsetAsync(G, 0, (size_t)2 * C * (H + 4) * a.Wg * sizeof(__half), st));
  const int gx = pool ? (int)std::min<long long>(((long long)hq * wq + 255) / 256, kPrepBlocks)
                      : (int)std::min<long long>(((long long)hq * wq * (C / 8) + 255) / 256, 4 * kPrepBlocks);
  if (pool) enc_grad_prep_kernel<true><<<dim3(gx, C / 8), 256, 0, st>>>(a);
  else enc_grad_prep_kernel<false><<<gx, 256, 0, st>>>(a);
  enc_db_reduce_kernel<<<(C * 32 + 255) / 256, 256, 0, st>>>(dbpart, gx, C, scale_out, gb);
  count_launch(2);
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

}  // namespace

size_t encoder_tape_bytes(int H, int W) { return tape_layout(H, W).total; }
size_t encoder_backward_scratch_bytes(int H, int W) { return bwd_layout(H, W).total; }

// forward of the training step: the same kernels, every layer's activation planes kept in `tape`
int encoder_forward_train(const void* packed, const float* img, int H, int W, float* out, void* tape,
                          size_t tape_bytes, cudaStream_t st) {
  CRNERF_REQUIRE(packed && img && out && tape, "null argument");
  CRNERF_REQUIRE(H >= 8 && W >= 8 && H <= 8192 && W <= 8192, "image %dx%d unsupported (8..8192 per side)", H, W);
  const Tape T = tape_layout(H, W);
  CRNERF_REQUIRE(tape_bytes >= T.total, "tape too small");
  CRNERF_REQUIRE((reinterpret_cast<uintptr_t>(tape) & 255) == 0 && (reinterpret_cast<uintptr_t>(packed) & 127) == 0,
                 "tape must be 256-byte and packed 128-byte aligned");
  TcLayer L[4];
  size_t blob_off;
  tc_layers(L, blob_off);
  const uint8_t* wimg = static_cast<const uint8_t*>(packed);
  const float* blob = reinterpret_cast<const float*>(wimg + blob_off);
  uint8_t* base = static_cast<uint8_t*>(tape);
  auto P = [&](size_t off) { return reinterpret_cast<__half*>(base + off); };
  const int H2 = H / 2, W2 = W / 2, H4 = H2 / 2, W4 = W2 / 2;
  auto plane = [](int h, int w) { return (long long)(h + 2) * (w + 2); };
  auto pool = [&](const __half* in, int C, int h, int w, __half* o) {
    const long long total = (long long)(h / 2) * (w / 2) * (C / 8);
    const int grid = (int)std::min<long long>((total + 255) / 256, 16LL * num_sms());
    enc_pool_planes_kernel<<<grid, 256, 0, st>>>(in, in + (size_t)C * plane(h, w), C, h, w, o,
                                                 o + (size_t)C * plane(h / 2, w / 2));
    count_launch();
  };
  int rc;
  {
    const long long total = plane(H, W);
    const int grid = (int)std::min<long long>((total + 255) / 256, 16LL * num_sms());
    enc_conv1_planes_kernel<<<grid, 256, 0, st>>>(img, H, W, blob, P(T.p0), P(T.p0) + (size_t)8 * total);
    count_launch();
    CRNERF_CUDA(cudaGetLastError());
  }
  if ((rc = launch_conv<8, 64, kOutPlanes>(P(T.p0), wimg + first_image_offset(blob_off), blob + Blob::b2, nullptr,
                                           P(T.a2), plane(H, W), H, W, st)))
    return rc;
  if ((rc = launch_conv<64, 64, kOutPlanes>(P(T.a2), wimg + L[0].offset, blob + Blob::b3, nullptr, P(T.a3), plane(H, W),
                                            H, W, st)))
    return rc;
  pool(P(T.a3), 64, H, W, P(T.q3));
  if ((rc = launch_conv<64, 128, kOutPlanes>(P(T.q3), wimg + L[1].offset, blob + Blob::b4, nullptr, P(T.a4),
                                             plane(H2, W2), H2, W2, st)))
    return rc;
  if ((rc = launch_conv<128, 128, kOutPlanes>(P(T.a4), wimg + L[2].offset, blob + Blob::b5, nullptr, P(T.a5),
                                              plane(H2, W2), H2, W2, st)))
    return rc;
  pool(P(T.a5), 128, H2, W2, P(T.q5));
  float* F = reinterpret_cast<float*>(base + T.f);
  if ((rc = launch_conv<128, 128, kOutRows>(P(T.q5), wimg + L[3].offset, blob + Blob::b6, F, nullptr, 0, H4, W4, st)))
    return rc;
  enc_tail_kernel<<<1024, 512, 0, st>>>(F, H4, W4, blob, out, reinterpret_cast<float*>(base + T.pooled));
  count_launch();
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

// backward: grad_out (64,32,32) -> gradients of the 14 parameter tensors (torch layouts) and, optionally, of img
int encoder_backward(const void* packed, const float* img, int H, int W, const float* out, const float* grad_out,
                     const void* tape, float* const* gw, float* const* gb, float* grad_img, void* scratch,
                     size_t scratch_bytes, cudaStream_t st) {
  CRNERF_REQUIRE(packed && img && out && grad_out && tape && gw && gb && scratch, "null argument");
  for (int i = 0; i < 7; ++i) CRNERF_REQUIRE(gw[i] && gb[i], "conv%d: null gradient buffer", i + 1);
  CRNERF_REQUIRE(H >= 8 && W >= 8 && H <= 8192 && W <= 8192, "image %dx%d unsupported (8..8192 per side)", H, W);
  const Tape T = tape_layout(H, W);
  const BwdScratch S = bwd_layout(H, W);
  CRNERF_REQUIRE(scratch_bytes >= S.total, "scratch too small");
  CRNERF_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 255) == 0 && (reinterpret_cast<uintptr_t>(tape) & 255) == 0,
                 "scratch and tape must be 256-byte aligned");
  TcLayer L[4];
  size_t blob_off;
  tc_layers(L, blob_off);
  const uint8_t* wimg = static_cast<const uint8_t*>(packed);
  const float* blob = reinterpret_cast<const float*>(wimg + blob_off);
  const uint8_t* tb = static_cast<const uint8_t*>(tape);
  uint8_t* sb = static_cast<uint8_t*>(scratch);
  auto P = [&](size_t off) { return reinterpret_cast<const __half*>(tb + off); };
  __half* Ga = reinterpret_cast<__half*>(sb + S.ga);
  __half* Gb = reinterpret_cast<__half*>(sb + S.gb);
  float* DX = reinterpret_cast<float*>(sb + S.dx);
  float* part = reinterpret_cast<float*>(sb + S.part);
  float* dpooled = reinterpret_cast<float*>(sb + S.dpooled);
  float* dpre7 = reinterpret_cast<float*>(sb + S.dpre7);
  float* dbpart = reinterpret_cast<float*>(sb + S.dbpart);
  float* part2 = reinterpret_cast<float*>(sb + S.part2);
  float* scales = reinterpret_cast<float*>(sb + S.small);           // [layer 1..7]
  unsigned* maxbits = reinterpret_cast<unsigned*>(sb + S.small + 32);  // [stage]
  const int H2 = H / 2, W2 = W / 2, H4 = H2 / 2, W4 = W2 / 2;
  const float* F = reinterpret_cast<const float*>(tb + T.f);
  const float* pooled = reinterpret_cast<const float*>(tb + T.pooled);
  int rc;
  CRNERF_CUDA(cudaMemsetAsync(sb + S.small, 0, 256, st));

  // conv7 / pool tail
  enc_tail_bwd_kernel<<<1024, 128, 0, st>>>(grad_out, out, blob, dpre7, dpooled, maxbits + 6);
  enc_w7_grad_kernel<<<64, 1024, 0, st>>>(dpre7, pooled, gw[6], gb[6]);
  count_launch(2);
  // dZ6 planes
  {
    auto ilog2c = [](int v) { int e = 0; while ((1 << e) < v) ++e; return e; };
    const int fold = ilog2c((32 / H4 + 3) * (32 / W4 + 3));
    const int Wg = grad_stride(W4);
    CRNERF_CUDA(cudaMemsetAsync(Ga, 0, (size_t)2 * 128 * (H4 + 4) * Wg * sizeof(__half), st));
    const int gx = (int)std::min<long long>(((long long)H4 * W4 + 255) / 256, kPrepBlocks);
    enc_grad_top_kernel<<<dim3(gx, 16), 256, 0, st>>>(dpooled, F, H4, W4, Ga, Ga + (size_t)128 * (H4 + 4) * Wg, Wg,
                                                      maxbits + 6, fold, scales + 6, dbpart);
    enc_db_reduce_kernel<<<16, 256, 0, st>>>(dbpart, gx, 128, scales + 6, gb[5]);
    count_launch(2);
    CRNERF_CUDA(cudaGetLastError());
  }
  static const int stop = getenv("CRNERF_ENC_BWD_STOP") ? atoi(getenv("CRNERF_ENC_BWD_STOP")) : 0;   // debug: keep a stage's planes
  if (stop == 1) return CRNERF_OK;
  // conv6 (128 -> 128 at 1/4): input Q5
  if ((rc = launch_wgrad<128, 128>(Ga, P(T.q5), H4, W4, part, scales + 6, gw[5], st))) return rc;
  if ((rc = launch_conv<128, 128, kOutRaw>(Ga, wimg + dgrad_image_offset(blob_off, 3), blob + Blob::b6, DX, nullptr, 0,
                                           H4 + 2, grad_stride(W4) - 2, st, maxbits + 5)))
    return rc;
  if ((rc = launch_prep(true, DX, 128, H4, W4, P(T.a5), H2, W2, Gb, maxbits + 5, scales + 6, scales + 5, dbpart, gb[4], st)))
    return rc;
  if (stop == 2) return CRNERF_OK;
  // conv5 (128 -> 128 at 1/2): input A4
  if ((rc = launch_wgrad<128, 128>(Gb, P(T.a4), H2, W2, part, scales + 5, gw[4], st))) return rc;
  if ((rc = launch_conv<128, 128, kOutRaw>(Gb, wimg + dgrad_image_offset(blob_off, 2), blob + Blob::b5, DX, nullptr, 0,
                                           H2 + 2, grad_stride(W2) - 2, st, maxbits + 4)))
    return rc;
  if ((rc = launch_prep(false, DX, 128, H2, W2, P(T.a4), H2, W2, Ga, maxbits + 4, scales + 5, scales + 4, dbpart, gb[3], st)))
    return rc;
  if (stop == 3) return CRNERF_OK;
  // conv4 (64 -> 128 at 1/2): input Q3
  if ((rc = launch_wgrad<64, 128>(Ga, P(T.q3), H2, W2, part, scales + 4, gw[3], st))) return rc;
  if ((rc = launch_conv<128, 64, kOutRaw>(Ga, wimg + dgrad_image_offset(blob_off, 1), blob + Blob::b3, DX, nullptr, 0,
                                          H2 + 2, grad_stride(W2) - 2, st, maxbits + 3)))
    return rc;
  if ((rc = launch_prep(true, DX, 64, H2, W2, P(T.a3), H, W, Gb, maxbits + 3, scales + 4, scales + 3, dbpart, gb[2], st)))
    return rc;
  if (stop == 4) return CRNERF_OK;
  // conv3 (64 -> 64 at full resolution): input A2
  if ((rc = launch_wgrad<64, 64>(Gb, P(T.a2), H, W, part, scales + 3, gw[2], st))) return rc;
  if ((rc = launch_conv<64, 64, kOutRaw>(Gb, wimg + dgrad_image_offset(blob_off, 0), blob + Blob::b3, DX, nullptr, 0,
                                         H + 2, grad_stride(W) - 2, st, maxbits + 2)))
    return rc;
  if ((rc = launch_prep(false, DX, 64, H, W, P(T.a2), H, W, Ga, maxbits + 2, scales + 3, scales + 2, dbpart, gb[1], st)))
    return rc;
  // conv2 (3 -> 64) and conv1 (1x1) on the CUDA cores
  {
    const int Wg = grad_stride(W);
    const long long g_plane = (long long)(H + 4) * Wg;
    const __half* p0 = P(T.p0);
    const int nseg = H * ((W + kFwSeg - 1) / kFwSeg);
    const int g1 = std::min(nseg, kFirstBlocks);
    enc_first_wgrad_kernel<<<g1, 256, 0, st>>>(Ga, Ga + (size_t)64 * g_plane, g_plane, Wg, p0,
                                               p0 + (size_t)8 * (H + 2) * (W + 2), H, W, part);
    enc_part_reduce_kernel<<<1728 / 8, 256, 0, st>>>(part, g1, 1728, 0, 1728, scales + 2, gw[1]);
    // conv2's input gradient at every padded position (fp32, in the rows buffer), then fold + conv1
    float4* dp0 = reinterpret_cast<float4*>(DX);
    const long long quads = (long long)(H + 2) * ((W + 2 + 3) / 4);
    enc_first_dgrad_kernel<<<(int)std::min<long long>((quads + 127) / 128, 8LL * num_sms()), 128, 0, st>>>(
        Ga, Ga + (size_t)64 * g_plane, g_plane, Wg, blob, H, W, dp0);
    const int g2 = (int)std::min<long long>(((long long)H * W + 255) / 256, kFirstBlocks);
    enc_first_finish_kernel<<<g2, 256, 0, st>>>(dp0, blob, img, H, W, scales + 2, part2, grad_img);
    enc_part_reduce_kernel<<<2, 256, 0, st>>>(part2, g2, 12, 0, 9, scales + 2, gw[0]);
    enc_part_reduce_kernel<<<1, 256, 0, st>>>(part2, g2, 12, 9, 3, scales + 2, gb[0]);
    count_launch(6);
    CRNERF_CUDA(cudaGetLastError());
  }
  return CRNERF_OK;
}

}  // namespace crnerf

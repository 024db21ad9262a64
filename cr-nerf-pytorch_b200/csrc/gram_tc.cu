// Gram statistics of the cross-ray block on the tensor core.
//
// CNN.forward of the reference (models/linearStyleTransfer.py:28-37) pushes every pixel of the
// (mean-free) feature map through three 1x1 convolutions 64 -> 128 -> 64 -> 32 with
// LeakyReLU(0.2) in between and accumulates the 32x32 Gram matrix of the result over ALL
// pixels: 18.4 kMAC per pixel against 256 B of input, i.e. the only compute-heavy pass of
// the otherwise HBM-streaming block (see crossray.cu for the other passes).
//
// Per 128-pixel tile (one pixel per TMEM lane / epilogue thread):
//   X (128x64)  -> smem, K-major SW128                                 (threads, from HBM)
//   D1 = X  W1^T  (N=128)  -> +b1, LeakyReLU -> H1 in TMEM              tcgen05 SS
//   D2 = H1 W2^T  (N=64)   -> +b2, LeakyReLU -> H2 in TMEM              tcgen05 TS
//   D3 = H2 W3^T  (N=32)   -> +b3 -> Y; Y^T (32x128, pixels along K) -> smem
//   G += Y^T Y    (M=128 with only rows 0..31 meaningful, N=32, K=128)  tcgen05 SS
// G lives in TMEM for the CTA's whole lifetime and is written out once.
//
// Precision: the result feeds a 1e-4 parity bar through two 1024x1024 FCs, so 16-bit
// operands alone are not enough.  Every operand is split into fp16 hi + fp16 lo
// (x = hi + lo to ~2^-22) and every product is issued as three MMAs
// (hi*hi + hi*lo + lo*hi, fp32 accumulate): fp32-class accuracy at 3x a tensor cost that is
// still far below the pass's HBM time.
#include <cuda_fp16.h>
#include <algorithm>
#include "common.h"
#include "ptx.cuh"

namespace crnerf {
namespace {

constexpr int kGThreads = 288;  // warps 0-7: (lane quarter, column half) of a pixel row; warp 8: MMA issuer + TMEM allocator
// shared memory map (bytes).  The Gram A operand is addressed as a 128-row tile although only
// 32 rows (channels) exist: rows 32..127 alias whatever follows (the other Y^T slabs, the X
// tile) and only feed accumulator lanes 32..127, which are never read.
constexpr int kYtOff = 0;       // Y^T: hi slab0, hi slab1, lo slab0, lo slab1 (32 rows x 128 B each)
constexpr int kYtSlab = 4096;
constexpr int kXOff = 16384;    // X hi, X lo (128 rows x 128 B each)
constexpr int kW1Off = 49152;   // W1 hi, lo: 128 rows x 128 B
constexpr int kW2Off = 81920;   // W2 hi (2 slabs x 64 rows x 128 B), lo
constexpr int kW3Off = 114688;  // W3 hi, lo: 32 rows x 128 B
constexpr int kFOff = 122880;   // fp32: b1[128] b2[64] b3[32] mean[64]
constexpr int kBarOff2 = kFOff + 288 * 4;  // 8 mbarriers + tmem slot
constexpr int kGramSmem = kBarOff2 + 8 * 8 + 16;
// TMEM columns
constexpr uint32_t cD1 = 0, cA1h = 128, cA1l = 192, cD2 = 256, cA2h = 320, cA2l = 352, cD3 = 384, cG = 416;

__device__ __forceinline__ float lrelu02(float v) { return v > 0.f ? v : 0.2f * v; }

// (a, b) -> packed fp16 hi pair and packed fp16 lo pair (a = hi + lo to ~2^-22 relative)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack2<0, false>(a, b);
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  lo = pack2<0, false>(a - f.x, b - f.y);
}

// fp32 row-major weight (rows x cols) -> SW128 K-major fp16 hi/lo slabs of 64 columns
__device__ void stage_weight(const float* __restrict__ w, int rows, int cols, uint8_t* hi, uint8_t* lo) {
  const int slab_bytes = rows * 128;
  for (int e = threadIdx.x; e < rows * cols / 2; e += kGThreads) {
    const int r = e / (cols / 2), c = 2 * (e % (cols / 2));
    uint32_t h, l;
    split2(w[r * cols + c], w[r * cols + c + 1], h, l);
    const uint32_t off = (uint32_t)(c >> 6) * slab_bytes + sw128_offset(r, (c & 63) >> 3) + (c & 7) * 2;
    *reinterpret_cast<uint32_t*>(hi + off) = h;
    *reinterpret_cast<uint32_t*>(lo + off) = l;
  }
}

struct GramParams {
  const float* g;
  long long n, pix_stride, ch_stride;
  const float* mean;
  crnerf_cnn_weights w;
  float* partial;
  int vec;  // rows are contiguous, 16-byte aligned and a multiple of 4 floats apart: float4 loads
};

__global__ void __launch_bounds__(kGThreads, 1) gram_tc_kernel(const __grid_constant__ GramParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* fblob = reinterpret_cast<float*>(smem + kFOff);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBarOff2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  enum { X_FULL = 0, D1_FULL, A1_FULL, D2_FULL, A2_FULL, D3_FULL, YT_FULL, G_DONE };
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(&bars[X_FULL], 8);
    mbar_init(&bars[D1_FULL], 1);
    mbar_init(&bars[A1_FULL], 8);
    mbar_init(&bars[D2_FULL], 1);
    mbar_init(&bars[A2_FULL], 8);
    mbar_init(&bars[D3_FULL], 1);
    mbar_init(&bars[YT_FULL], 8);
    mbar_init(&bars[G_DONE], 1);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc<512>(tmem_slot);
  stage_weight(P.w.conv_w[0], 128, 64, smem + kW1Off, smem + kW1Off + 16384);
  stage_weight(P.w.conv_w[1], 64, 128, smem + kW2Off, smem + kW2Off + 16384);
  stage_weight(P.w.conv_w[2], 32, 64, smem + kW3Off, smem + kW3Off + 4096);
  for (int i = tid; i < 288; i += kGThreads)
    fblob[i] = i < 128 ? P.w.conv_b[0][i]
                       : (i < 192 ? P.w.conv_b[1][i - 128] : (i < 224 ? P.w.conv_b[2][i - 192] : P.mean[i - 224]));
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const long long n_tiles = (P.n + 127) / 128;
  const float* b1 = fblob, *b2 = fblob + 128, *b3 = fblob + 192, *mean = fblob + 224;

  if (warp == 8) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO 1024 | version | SW128
    auto desc = [](uint32_t saddr) {
      return (static_cast<uint64_t>(kHi) << 32) | (((saddr & 0x3ffffu) >> 4) | (1u << 16));
    };
    const uint32_t s0 = smem_u32(smem);
    // layer 1 of tile i: D1 = X W1^T, three split terms x 4 k-steps (SS).  Issued one stage
    // ahead (right after layer 2 of the previous tile): X(i) is staged as soon as D1(i-1) has
    // been drained, so these MMAs run under the previous tile's layer-3 / Gram phases.
    auto layer1 = [&](uint32_t i) {
      mbar_wait(&bars[X_FULL], i & 1, 60);
      tc_fence_after_sync();
      if (elect_one()) {
        const uint32_t id = make_idesc_f16(128, 128, 0);
#pragma unroll
        for (int term = 0; term < 3; ++term) {
          const uint32_t a = s0 + kXOff + (term == 2 ? 16384 : 0);    // hi, hi, lo
          const uint32_t bq = s0 + kW1Off + (term == 1 ? 16384 : 0);  // hi, lo, hi
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tmem + cD1, desc(a + 32 * k), desc(bq + 32 * k), id, (term | k) ? 1u : 0u);
        }
        umma_commit(&bars[D1_FULL]);
      }
      __syncwarp();
    };
    uint32_t it = 0;
    if ((long long)blockIdx.x < n_tiles) layer1(0);
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const uint32_t par = it & 1;
      // ---- layer 2: D2 = H1 W2^T, K = 128 (TS)
      mbar_wait(&bars[A1_FULL], par, 61);
      tc_fence_after_sync();
      if (elect_one()) {
        const uint32_t id = make_idesc_f16(128, 64, 0);
#pragma unroll
        for (int term = 0; term < 3; ++term) {
          const uint32_t a = tmem + (term == 2 ? cA1l : cA1h);
          const uint32_t bq = s0 + kW2Off + (term == 1 ? 16384 : 0);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_ts(tmem + cD2, a + 8 * k, desc(bq + (k >> 2) * 8192 + 32 * (k & 3)), id, (term | k) ? 1u : 0u);
        }
        umma_commit(&bars[D2_FULL]);
      }
      __syncwarp();
      if (t + gridDim.x < n_tiles) layer1(it + 1);   // A1_FULL(it) above also means D1 is drained
      // ---- layer 3: D3 = H2 W3^T, K = 64 (TS)
      mbar_wait(&bars[A2_FULL], par, 62);
      tc_fence_after_sync();
      if (elect_one()) {
        const uint32_t id = make_idesc_f16(128, 32, 0);
#pragma unroll
        for (int term = 0; term < 3; ++term) {
          const uint32_t a = tmem + (term == 2 ? cA2l : cA2h);
          const uint32_t bq = s0 + kW3Off + (term == 1 ? 4096 : 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ts(tmem + cD3, a + 8 * k, desc(bq + 32 * k), id, (term | k) ? 1u : 0u);
        }
        umma_commit(&bars[D3_FULL]);
      }
      __syncwarp();
      // ---- Gram: G += Y^T Y, K = the tile's 128 pixels (SS; A rows 32..127 are don't-care)
      mbar_wait(&bars[YT_FULL], par, 63);
      tc_fence_after_sync();
      if (elect_one()) {
        const uint32_t id = make_idesc_f16(128, 32, 0);
#pragma unroll
        for (int term = 0; term < 3; ++term) {
          const uint32_t a = s0 + kYtOff + (term == 2 ? 2 * kYtSlab : 0);
          const uint32_t bq = s0 + kYtOff + (term == 1 ? 2 * kYtSlab : 0);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_ss(tmem + cG, desc(a + (k >> 2) * kYtSlab + 32 * (k & 3)),
                    desc(bq + (k >> 2) * kYtSlab + 32 * (k & 3)), id, (it | term | k) ? 1u : 0u);
        }
        umma_commit(&bars[G_DONE]);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ pixel rows
    // warp w: TMEM lane quarter q = w & 3 (rows 32q..32q+31, one per lane), column half ch = w >> 2
    const int q = warp & 3, ch = warp >> 2;
    const int row = 32 * q + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    // this thread's half of a pixel row (32 channels) as 8 float4; the NEXT tile's values are
    // requested while the current tile is being processed (registers are plentiful here)
    float4 xr[8];
    auto load_x = [&](long long t) {
      const long long p = t * 128 + row;
      const bool valid = t < n_tiles && p < P.n;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (valid && P.vec) {
          xr[i] = __ldg(reinterpret_cast<const float4*>(P.g + p * P.pix_stride + 32 * ch) + i);
        } else if (valid) {
          const float* src = P.g + p * P.pix_stride + (long long)(32 * ch + 4 * i) * P.ch_stride;
          xr[i] = make_float4(__ldg(src), __ldg(src + P.ch_stride), __ldg(src + 2 * P.ch_stride),
                              __ldg(src + 3 * P.ch_stride));
        } else {
          xr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    };
    // X row (held in xr) of tile tt -> smem (hi, lo), then signal the issuer
    auto stage_x = [&](long long tt) {
      const bool valid = tt * 128 + row < P.n;
      {
        uint8_t* xh = smem + kXOff, *xl = xh + 16384;
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          const float4 q0 = xr[2 * c8], q1 = xr[2 * c8 + 1];
          const float v[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
          const int c0 = 32 * ch + 8 * c8;
          uint32_t h[4], l[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float a = valid ? v[2 * j] - mean[c0 + 2 * j] : 0.f;
            const float b = valid ? v[2 * j + 1] - mean[c0 + 2 * j + 1] : 0.f;
            split2(a, b, h[j], l[j]);
          }
          const uint32_t off = sw128_offset(row, 4 * ch + c8);
          *reinterpret_cast<uint4*>(xh + off) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4*>(xl + off) = make_uint4(l[0], l[1], l[2], l[3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[X_FULL]);
      }
    };
    load_x(blockIdx.x);
    if ((long long)blockIdx.x < n_tiles) {
      stage_x(blockIdx.x);
      load_x(blockIdx.x + gridDim.x);
    }
    uint32_t it = 0;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const uint32_t par = it & 1;
      const long long p = t * 128 + row;
      const bool valid = p < P.n;
      // ---- epilogue 1: H1 = LeakyReLU(D1 + b1) -> A1 hi / lo (128 values = 64 + 64 columns)
      mbar_wait(&bars[D1_FULL], par, 64);
      tc_fence_after_sync();
#pragma unroll
      for (int qq = 0; qq < 2; ++qq) {
        const int blk = 2 * ch + qq;
        uint32_t v[32], h[16], l[16];
        tmem_ld_x32(lane_base + cD1 + 32 * blk, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          split2(lrelu02(__uint_as_float(v[2 * j]) + b1[32 * blk + 2 * j]),
                 lrelu02(__uint_as_float(v[2 * j + 1]) + b1[32 * blk + 2 * j + 1]), h[j], l[j]);
        tmem_st_x16p(lane_base + cA1h + 16 * blk, h);
        tmem_st_x16p(lane_base + cA1l + 16 * blk, l);
      }
      tmem_st_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[A1_FULL]);
      // D1(t) is drained and its MMAs have retired: stage the next tile's X now, so its layer 1
      // runs under this tile's remaining phases; then request the tile after that
      if (t + gridDim.x < n_tiles) {
        stage_x(t + gridDim.x);
        load_x(t + 2 * gridDim.x);
      }
      // ---- epilogue 2: H2 = LeakyReLU(D2 + b2) -> A2 hi / lo (64 values = 32 + 32 columns)
      mbar_wait(&bars[D2_FULL], par, 65);
      tc_fence_after_sync();
      {
        uint32_t v[32], h[16], l[16];
        tmem_ld_x32(lane_base + cD2 + 32 * ch, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          split2(lrelu02(__uint_as_float(v[2 * j]) + b2[32 * ch + 2 * j]),
                 lrelu02(__uint_as_float(v[2 * j + 1]) + b2[32 * ch + 2 * j + 1]), h[j], l[j]);
        tmem_st_x16p(lane_base + cA2h + 16 * ch, h);
        tmem_st_x16p(lane_base + cA2l + 16 * ch, l);
      }
      tmem_st_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[A2_FULL]);
      // ---- epilogue 3: Y = D3 + b3 (no activation, linearStyleTransfer.py:15), zero for pixels
      // beyond n; Y^T hi / lo -> smem with the pixel index along K (this thread: 16 channels)
      mbar_wait(&bars[D3_FULL], par, 66);
      tc_fence_after_sync();
      {
        uint32_t v[16];
        tmem_ld_x16(lane_base + cD3 + 16 * ch, v);
        tmem_ld_wait();
        if (it > 0) mbar_wait(&bars[G_DONE], (it - 1) & 1, 67);  // previous Gram MMAs done reading Y^T
        uint8_t* yh = smem + kYtOff + (row >> 6) * kYtSlab;
        uint8_t* yl = yh + 2 * kYtSlab;
        const uint32_t kk = row & 63;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int c = 16 * ch + j;
          const float y = valid ? __uint_as_float(v[j]) + b3[c] : 0.f;
          const __half hi = __float2half_rn(y);
          const __half lo = __float2half_rn(y - __half2float(hi));
          const uint32_t off = sw128_offset(c, kk >> 3) + (kk & 7) * 2;
          *reinterpret_cast<__half*>(yh + off) = hi;
          *reinterpret_cast<__half*>(yl + off) = lo;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[YT_FULL]);
      }
    }
    // ---- G (rows 0..31) -> this block's partial
    if (it > 0) {
      mbar_wait(&bars[G_DONE], (it - 1) & 1, 68);
      tc_fence_after_sync();
    }
    if (warp == 0) {
      uint32_t v[32];
      if (it > 0) {
        tmem_ld_x32(lane_base + cG, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) P.partial[(long long)blockIdx.x * 1024 + row * 32 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 8) tmem_dealloc<512>(tmem);
}

}  // namespace

// partial[nb][1024] <- per-block un-normalised Gram partials; returns nb through *n_blocks
int gram_tc(const crnerf_cnn_weights& cw, const float* g, int64_t n, int64_t ps, int64_t cs, const float* mean,
            float* partial, int max_blocks, int* n_blocks, cudaStream_t st) {
  const long long tiles = (n + 127) / 128;
  const int nb = (int)std::max<long long>(1, std::min<long long>(tiles, std::min(max_blocks, num_sms())));
  GramParams P;
  P.g = g;
  P.n = n;
  P.pix_stride = ps;
  P.ch_stride = cs;
  P.mean = mean;
  P.w = cw;
  P.partial = partial;
  P.vec = cs == 1 && (ps & 3) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0;
  CRNERF_CUDA(cudaFuncSetAttribute(gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGramSmem));
  gram_tc_kernel<<<nb, kGThreads, kGramSmem, st>>>(P);
  count_launch();
  CRNERF_CUDA(cudaGetLastError());
  *n_blocks = nb;
  return CRNERF_OK;
}

}  // namespace crnerf

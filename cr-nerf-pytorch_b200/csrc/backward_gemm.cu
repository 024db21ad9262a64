// Backward of the NeRF_sigma MLP on the tensor cores (sm_100a): the 11 dgrad and 13 wgrad GEMMs of
// one render pass of the training step (reference train_mask_grid_sample.py:186-197 under
// autograd; the forward is models/nerf.py:157-182).  Replaces the cuBLAS calls of round 1.
//
// Data layout ("tiled16"): every (points x features) 16-bit operand - the activations saved by
// the training forward and the gradients flowing down the chain - is stored per 128-point tile
// and 64-feature slab as a 16 KB block of 128-byte rows (one per point) with the 16-byte chunks
// XOR-swizzled by (row % 8).  Such a block is, byte for byte, both
//   * a K-major SWIZZLE_128B tcgen05 operand with M = points, K = features     (dgrad's A), and
//   * an MN-major SWIZZLE_128B operand with M/N = features, K = points          (wgrad's A and B;
//     validated bit-exactly by tools/mn_probe.cu),
// so a tile is fetched with plain bulk copies (cp.async.bulk, no tensor map) and no transpose is
// ever materialised.
//
// Gradients are 16-bit in the operand format of the forward (fp16 or bf16), each stage of the
// chain scaled by its own power of two chosen ON THE DEVICE: the kernel that produces a stage
// measures its max magnitude (atomicMax on the bit pattern), and the kernel that produces the
// next stage re-centres so that this max would land in [32, 64) - 2^10 of headroom for growth
// inside one layer, 2^-20 of the max still a normal fp16 number.  (One scale for the whole chain
// is not enough: magnitudes fall ~0.4x per layer at default init and the bottom layers end up in
// fp16's subnormals - measured: 1e-2 relative error at layer 1 instead of 1e-3.)  Scales are
// undone when weight and bias gradients are accumulated in fp32.  State words (device floats):
// st[k] = max |stored value| of stage k, st[16 + k] = its scale, st[31] = max |top gradient|.
//
// Per layer l (top to bottom), Gm_l = dL/d(pre-activation of layer l), X_l = the layer's input:
//   wgrad:  dW_l (out x in) = Gm_l^T X_l         M = out features, N = in features, K = points;
//           each CTA reduces its share of the tiles in TMEM (256 x 256 fp32 = all 512 columns) and
//           stores its partial; ONE reduce kernel per pass then sums the partials of all 13
//           products in CTA order (deterministic; atomics on 148 x 65 k addresses per layer cost
//           more than the GEMM - measured 47 us per layer against 20 us of HBM time)
//   dgrad:  Gm_{l-1} = (Gm_l W_l) * relu'(X_l)   M = points, N = in features, K = out features;
//           W_l^T stays resident in shared memory, the epilogue adds the sigma-head term at h8,
//           applies the ReLU mask from the saved activation, accumulates the bias gradient of
//           layer l-1 (column sums) and writes the 16-bit tile of the next step.
// Both kernels are HBM-bound (134 / 200 MB per 256-wide layer at 131 k points).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <algorithm>
#include <cstring>
#include "common.h"
#include "nerf_layout.h"
#include "ptx.cuh"

namespace crnerf {
namespace {

constexpr int kSlab = 16384;   // 128 rows x 128 B
constexpr int kHalf = 8192;    // 64 rows

constexpr int kStScale = 16, kStTop = 31;
// power of two that moves a max magnitude `a` into [32, 64)
__device__ __forceinline__ float recentre(float a) {
  if (!(a > 0.f) || !(a < 3.0e38f)) return 1.f;
  int e;
  frexpf(a, &e);                 // a = m * 2^e, m in [0.5, 1)
  e = max(-100, min(100, e));
  return exp2f((float)(6 - e));
}

// MN-major SWIZZLE_128B descriptor: lbo = bytes between 64-feature slabs, sbo = 1024 (8-point groups)
__device__ __forceinline__ uint64_t desc_mn(uint32_t addr, uint32_t lbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3ffff) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3fff) << 16;
  d |= static_cast<uint64_t>(1024u >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__device__ __forceinline__ uint32_t idesc_mn(uint32_t M, uint32_t N, uint32_t fmt) {
  return make_idesc_f16(M, N, fmt) | (1u << 15) | (1u << 16);
}

// one mbarrier arrival on behalf of a converged warp
__device__ __forceinline__ void warp_arrive_bar(uint64_t* bar) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}
// column sums over the 32 lanes of a warp for 32 values per lane (transposing butterfly, 31
// shuffles): lane L returns sum over lanes of v[L]
template <int kHalfW>
__device__ __forceinline__ void bfly(float (&v)[32], int lane) {
  const bool up = (lane & kHalfW) != 0;
#pragma unroll
  for (int i = 0; i < kHalfW; ++i) {
    const float send = up ? v[i] : v[i + kHalfW];
    const float keep = up ? v[i + kHalfW] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, kHalfW);
  }
}
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
  bfly<16>(v, lane);
  bfly<8>(v, lane);
  bfly<4>(v, lane);
  bfly<2>(v, lane);
  bfly<1>(v, lane);
  return v[0];
}

// rows [r0, 128) of one slab := 0 (the point count's tail inside the last tile)
__global__ void zero_tail_kernel(uint8_t* base, int n_slabs, long long slab_stride, int r0) {
  uint4* dst = reinterpret_cast<uint4*>(base + (size_t)blockIdx.x * slab_stride + (size_t)r0 * 128);
  const int n16 = (128 - r0) * 8;
  for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = make_uint4(0, 0, 0, 0);
}

// ------------------------------------------------------------------------------------------
// top of the chain: d_rgb (P,64) fp32 + d_sigma (P) fp32 -> Gm_rgb tiles (one slab, scaled) and
// Gsig tiles (one slab whose column 0 is d_sigma, scaled); bias gradients of static_rgb / static_sigma
// ------------------------------------------------------------------------------------------
template <int kFmt>
__global__ void __launch_bounds__(128) top_pack_kernel(const float* __restrict__ d_rgb, const float* __restrict__ d_sig,
                                                       long long P, float* __restrict__ st,
                                                       uint8_t* __restrict__ g_rgb, uint8_t* __restrict__ g_sig,
                                                       float* __restrict__ db_rgb, float* __restrict__ db_sig) {
  __shared__ float red[4][65];
  const float top = st[kStTop];
  const float S = recentre(top);
  if (blockIdx.x == 0 && threadIdx.x == 0) {   // stage 0 = the two top tiles
    st[kStScale] = S;
    st[0] = top * S;
  }
  const int r = threadIdx.x, warp = r >> 5, lane = r & 31;
  const long long p = (long long)blockIdx.x * 128 + r;
  const bool valid = p < P;
  float v[64];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float4 t = valid ? __ldg(reinterpret_cast<const float4*>(d_rgb + p * 64) + i) : make_float4(0, 0, 0, 0);
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
  const float ds = valid ? __ldg(d_sig + p) : 0.f;
  uint4* row = reinterpret_cast<uint4*>(g_rgb + (size_t)blockIdx.x * kSlab + r * 128);
  uint4* rows = reinterpret_cast<uint4*>(g_sig + (size_t)blockIdx.x * kSlab + r * 128);
  const uint32_t rx = r & 7;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint32_t w[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) w[q] = pack2<kFmt, false>(v[16 * k + 2 * q] * S, v[16 * k + 2 * q + 1] * S);
    stg_row_pair(row, (uint32_t)k, rx, make_uint4(w[0], w[1], w[2], w[3]), make_uint4(w[4], w[5], w[6], w[7]));
    stg_row_pair(rows, (uint32_t)k, rx, k == 0 ? make_uint4(pack2<kFmt, false>(ds * S, 0.f), 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u),
                 make_uint4(0u, 0u, 0u, 0u));
  }
  // bias gradients (unscaled fp32): column sums over the block's rows, then one atomic per column
  float dsum = ds;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, d);
#pragma unroll
  for (int c = 0; c < 64; ++c) {
    float t = v[c];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
    if (lane == 0) red[warp][c] = t;
  }
  if (lane == 0) red[warp][64] = dsum;
  __syncthreads();
  if (r < 65) {
    const float t = (red[0][r] + red[1][r]) + (red[2][r] + red[3][r]);
    if (r < 64)
      atomicAdd(db_rgb + r, t);
    else
      atomicAdd(db_sig, t);
  }
}

// ------------------------------------------------------------------------------------------
// weights for dgrad: W_l^T as K-major SWIZZLE_128B slabs.  Slab j of layer l holds, for every input
// feature k (row, 128 B), the 64 output features n = 64 j .. 64 j + 63:  element (k, n) = W[n][k0 + k]
// ------------------------------------------------------------------------------------------
struct BwdLayer {
  int w_index;    // index into the 12 weight pointers
  int n_out;      // N (64 / 128 / 256)
  int k_in;       // rows of the image = input features that receive a gradient (128 / 256)
  int k0;         // first such column of W
  int ld;         // row stride of W
  int offset;     // byte offset of the layer's image
};
constexpr int kBwdLayers = 10;   // layers 1..7 (trunk 2-8), final, dir, rgb; layer 0 has no dgrad
struct BwdPackParams {
  const float* w[12];
  uint8_t* img;
  BwdLayer L[kBwdLayers];
  int fmt;
  int total_bytes;
};

__global__ void bwd_pack_kernel(const __grid_constant__ BwdPackParams P) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.total_bytes / 16; i += gridDim.x * blockDim.x) {
    const int byte = i * 16;
    int li = 0;
    while (li + 1 < kBwdLayers && P.L[li + 1].offset <= byte) ++li;
    const BwdLayer L = P.L[li];
    const int within = byte - L.offset;
    const int slab_bytes = L.k_in * 128;
    const int j = within / slab_bytes, rem = within % slab_bytes;
    const int k = rem >> 7, pos = (rem & 127) >> 4;
    const int n0 = 64 * j + ((pos ^ (k & 7)) & 7) * 8;
    const float* w = P.w[L.w_index] + (size_t)(L.k0 + k);
    uint32_t out[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float v0 = w[(size_t)(n0 + 2 * e) * L.ld], v1 = w[(size_t)(n0 + 2 * e + 1) * L.ld];
      out[e] = P.fmt == 0 ? pack2<0, false>(v0, v1) : pack2<1, false>(v0, v1);
    }
    *reinterpret_cast<uint4*>(P.img + byte) = make_uint4(out[0], out[1], out[2], out[3]);
  }
}

// ------------------------------------------------------------------------------------------
// wgrad
// ------------------------------------------------------------------------------------------
struct WgradParams {
  const uint8_t* g;     // tiled16 (P x 64*g_slabs)
  const uint8_t* x;     // tiled16, x_slabs_total slabs per tile; the kernel reads slabs [xs0, xs0 + nxs)
  float* partial;       // fp32 [gridDim.x][nrows][64 * nxs]: this launch's per-CTA partial products
  int g_slabs, x_slabs_total, xs0, nxs;
  int nrows;            // valid rows (output features)
  int n_tiles;
  int x_dead;           // no later kernel of this pass reads these X slabs: stream them through the L2
};
// one product's share of the reduce kernel
struct ReduceJob {
  const float* partial;
  float* dw;            // fp32 (rows x ldw)
  const float* scale;   // scale of G's stage
  int n_parts, nrows, n_mma;
  int ldw, c0;          // output column of X-window column xcol0
  int xcol0, ncols;     // valid columns of the X window
};
constexpr int kMaxJobs = 16;
struct ReduceParams {
  ReduceJob job[kMaxJobs];
  int n_jobs;
};

// dW[row][c0 + col - xcol0] = (sum over CTAs of partial[cta][row][col]) / scale, fixed order.
// One thread per four consecutive columns (16-byte loads), eight partials in flight.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const __grid_constant__ ReduceParams P) {
  pdl_wait();   // the last wgrad of the pass (and, through the chain, every kernel before it)
  const ReduceJob& J = P.job[blockIdx.y];
  const int n = J.nrows * J.n_mma;
  const int i = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (i >= n) return;
  int row = i / J.n_mma, col = i % J.n_mma;
  if (J.nrows % 128 == 0) {   // tiled partial (wgrad_kernel's staged epilogue): [128-row block][4-column group][row][4]
    const int f4 = i >> 2, blk = (J.n_mma / 4) * 128, rem = f4 % blk;
    row = (f4 / blk) * 128 + rem % 128;
    col = (rem / 128) * 4;
  }
  if (col + 3 < J.xcol0 || col >= J.xcol0 + J.ncols) return;
  const float4* p = reinterpret_cast<const float4*>(J.partial + i);
  const size_t stride = (size_t)n / 4;
  float4 acc[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  int b = 0;
  for (; b + 7 < J.n_parts; b += 8) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float4 v = __ldcs(p + (size_t)(b + u) * stride);
      acc[u].x += v.x; acc[u].y += v.y; acc[u].z += v.z; acc[u].w += v.w;
    }
  }
  for (; b < J.n_parts; ++b) {
    const float4 v = __ldcs(p + (size_t)b * stride);
    acc[0].x += v.x; acc[0].y += v.y; acc[0].z += v.z; acc[0].w += v.w;
  }
  const float inv = 1.f / *J.scale;
  float r[4];
  r[0] = ((acc[0].x + acc[1].x) + (acc[2].x + acc[3].x)) + ((acc[4].x + acc[5].x) + (acc[6].x + acc[7].x));
  r[1] = ((acc[0].y + acc[1].y) + (acc[2].y + acc[3].y)) + ((acc[4].y + acc[5].y) + (acc[6].y + acc[7].y));
  r[2] = ((acc[0].z + acc[1].z) + (acc[2].z + acc[3].z)) + ((acc[4].z + acc[5].z) + (acc[6].z + acc[7].z));
  r[3] = ((acc[0].w + acc[1].w) + (acc[2].w + acc[3].w)) + ((acc[4].w + acc[5].w) + (acc[6].w + acc[7].w));
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int c = col + e;
    if (c >= J.xcol0 && c < J.xcol0 + J.ncols) J.dw[(size_t)row * J.ldw + J.c0 + (c - J.xcol0)] = r[e] * inv;
  }
}
constexpr int kWgStages = 3;

template <int kFmt>
__global__ void __launch_bounds__(192, 1) wgrad_kernel(const __grid_constant__ WgradParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // stage: [G half-tile: ga slabs x 8 KB][X half-tile: nxs x 8 KB]; ga = max(2, g_slabs) so that M = 128
  // MMAs always find two slabs (the second one is zero for the 64-wide rgb gradient)
  const int ga = P.g_slabs < 2 ? 2 : P.g_slabs;
  const int stage_bytes = (ga + P.nxs) * kHalf;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWgStages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kWgStages;
  uint64_t* done = bars + 2 * kWgStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kWgStages + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kWgStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (P.g_slabs < 2) {   // zero the phantom second G slab of every stage once
    for (int i = threadIdx.x; i < kWgStages * (kHalf / 16); i += blockDim.x) {
      const int st = i / (kHalf / 16), o = i % (kHalf / 16);
      reinterpret_cast<uint4*>(smem + st * stage_bytes + kHalf)[o] = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();      // everything above overlapped the previous kernel's tail
  pdl_trigger();
  // this CTA's half-tiles (64 points each): blockIdx.x, blockIdx.x + gridDim.x, ... of the 2 * n_tiles - halves, not
  // tiles, are dealt out, so that 512 tiles over 148 CTAs cost 7 steps per CTA instead of 8
  int n_steps = 0;
  for (int t = blockIdx.x; t < 2 * P.n_tiles; t += gridDim.x) ++n_steps;
  const int m_halves = (P.g_slabs + 1) / 2;   // 128-row blocks of dW
  const int n_mma = P.nxs * 64;

  if (warp == 0) {
    if (elect_one()) {
      // Tiles in DESCENDING order: the dgrad kernel that ran just before this one wrote G in ascending order, so
      // the highest tiles are what the L2 still holds; the dgrad that follows re-reads G and X from the low end,
      // where this kernel ends (G + X of one 256-wide layer at 131 k points = 134 MB against 126 MB of L2).
      // x_dead: no later kernel reads these X slabs (evict-first).
      const uint64_t pol_x = P.x_dead ? l2_policy_evict_first() : 0;
      for (int s = 0; s < n_steps; ++s) {
        const int st = s % kWgStages, n = s / kWgStages;
        mbar_wait(&empty[st], (n & 1) ^ 1, 1);
        const size_t half = (size_t)2 * P.n_tiles - 1 - ((size_t)blockIdx.x + (size_t)s * gridDim.x);
        const size_t tile = half >> 1;
        const int h = (int)(half & 1);
        uint8_t* dst = smem + st * stage_bytes;
        mbar_arrive_expect_tx(&full[st], (uint32_t)((P.g_slabs + P.nxs) * kHalf));
        for (int j = 0; j < P.g_slabs; ++j)
          bulk_g2s(dst + j * kHalf, P.g + (tile * P.g_slabs + j) * kSlab + h * kHalf, kHalf, &full[st]);
        for (int j = 0; j < P.nxs; ++j) {
          const uint8_t* src = P.x + (tile * P.x_slabs_total + P.xs0 + j) * kSlab + h * kHalf;
          if (P.x_dead)
            bulk_g2s_hint(dst + (ga + j) * kHalf, src, kHalf, &full[st], pol_x);
          else
            bulk_g2s(dst + (ga + j) * kHalf, src, kHalf, &full[st]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    const uint32_t idesc = idesc_mn(128, (uint32_t)n_mma, kFmt);
    for (int s = 0; s < n_steps; ++s) {
      const int st = s % kWgStages, n = s / kWgStages;
      mbar_wait(&full[st], n & 1, 2);
      tc_fence_after_sync();
      if (elect_one()) {
        const uint32_t base = smem_u32(smem + st * stage_bytes);
        for (int mh = 0; mh < m_halves; ++mh) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {   // 64 points = 4 k-steps of 16 points = 2048 B down every slab
            const uint64_t ad = desc_mn(base + (uint32_t)(mh * 2) * kHalf + ks * 2048u, kHalf);
            const uint64_t bd = desc_mn(base + (uint32_t)ga * kHalf + ks * 2048u, kHalf);
            umma_ss(tmem + (uint32_t)mh * 256u, ad, bd, idesc, (s | ks) ? 1u : 0u);
          }
        }
        umma_commit(&empty[st]);
        if (s == n_steps - 1) umma_commit(done);
      }
      __syncwarp();
    }
  } else {
    // ---- epilogue (warps 2..5: TMEM lane quarter = warp % 4): this CTA's partial product
    const int q = warp & 3;
    float* part = P.partial + (size_t)blockIdx.x * P.nrows * n_mma;
    if (n_steps > 0) {
      mbar_wait(done, 0, 3);
      tc_fence_after_sync();
    }
    if (P.nrows == m_halves * 128) {
      // Whole 128-row blocks: staged through the (now idle) operand ring and written with bulk copies.  A chunk =
      // 32 columns of one block = 16 KB laid out [4-column group (8)][row (128)][16 B]: a lane's stores are 16 B
      // apart (conflict-free) and the chunk leaves as full-line writes (a thread-per-row store to the 1 KB rows of
      // a plain [row][column] partial touches 32 lines per request: 9 us of a 29 us kernel).  The partial keeps this
      // tiled order, [block][4-column group][row][4]; wgrad_reduce_kernel decodes it.  Four buffers in rotation.
      const int trow = q * 32 + lane;
      const bool issuer = warp == 2 && lane == 0;
      int ci = 0;
      for (int mh = 0; mh < m_halves; ++mh) {
        for (int c0 = 0; c0 < n_mma; c0 += 32, ++ci) {
          uint32_t v[32];
          if (n_steps > 0) {
            tmem_ld_x32(tmem + (uint32_t)mh * 256u + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0u;
          }
          uint8_t* buf = smem + (ci & 3) * 16384;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(buf + j * 2048 + trow * 16) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          fence_proxy_async_smem();
          if (issuer) bulk_wait_group_read<2>();   // the buffer of chunk ci + 1 (written after this barrier) has drained
          named_bar_sync(1, 128);
          if (issuer) {
            bulk_s2g(part + ((size_t)mh * (n_mma / 4) + c0 / 4) * 512, buf, 16384);
            bulk_commit_group();
          }
        }
      }
      if (issuer) bulk_wait_group_all();
    } else
    for (int mh = 0; mh < m_halves; ++mh) {
      const int rowf = mh * 128 + q * 32 + lane;
      for (int c0 = 0; c0 < n_mma; c0 += 32) {
        uint32_t v[32];
        if (n_steps > 0) {
          tmem_ld_x32(tmem + (uint32_t)mh * 256u + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
        if (rowf < P.nrows) {
          uint4* dst = reinterpret_cast<uint4*>(part + (size_t)rowf * n_mma + c0);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            stg_v8_cs(dst + 2 * j, make_uint4(v[8 * j], v[8 * j + 1], v[8 * j + 2], v[8 * j + 3]),   // read once, by the
                      make_uint4(v[8 * j + 4], v[8 * j + 5], v[8 * j + 6], v[8 * j + 7]));               // pass's last kernel
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------
// dgrad
// ------------------------------------------------------------------------------------------
struct DgradParams {
  const uint8_t* g;      // tiled16 (P x 64*g_slabs): Gm of this layer (scaled)
  const uint8_t* wimg;   // this layer's W^T image: g_slabs slabs of (k_in x 128 B)
  const uint8_t* act;    // tiled16 saved activation of the layer below (ReLU mask), or nullptr
  uint8_t* out;          // tiled16 (P x k_in): Gm of the layer below (scaled)
  float* db;             // fp32 (k_in): bias gradient of the layer below, accumulated with atomics
  const float* dsig;     // fp32 (P): sigma-head gradient joined at h8 (d_sigma x W_sigma), or nullptr
  const float* wsig;     // fp32 (k_in): static_sigma weight row
  float* st;             // scale state (see the header comment)
  int stage;             // stage index of g; the output is stage + 1
  long long n_points;
  int g_slabs, k_in, act_slabs_total, act_s0;
  int n_tiles;
};
constexpr int kDgStages = 3;
constexpr int kDgStageBytes = 2 * kSlab;   // up to two 64-feature slabs of G per stage
constexpr int kDgThreads = 576;            // producer, issuer, 16 epilogue warps (4 lane quarters x 4 output slabs)

template <int kFmt>
__global__ void __launch_bounds__(kDgThreads, 1) dgrad_kernel(const __grid_constant__ DgradParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int w_bytes = P.g_slabs * P.k_in * 128;
  uint8_t* sW = smem;
  uint8_t* ring = smem + w_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + kDgStages * kDgStageBytes);
  uint64_t* full = bars;                 // [kDgStages]
  uint64_t* empty = bars + kDgStages;    // [kDgStages]
  uint64_t* w_full = bars + 2 * kDgStages;
  uint64_t* d_full = w_full + 1;         // [2]
  uint64_t* d_empty = d_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_empty + 2);
  float* s_wsig = reinterpret_cast<float*>(tmem_slot + 2);   // [256]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kDgStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&d_full[i], 1);
      mbar_init(&d_empty[i], (uint32_t)(4 * (P.k_in / 64)));   // epilogue warps that own a slab
    }
    fence_mbar_init();
  }
  if (P.dsig)
    for (int i = threadIdx.x; i < P.k_in; i += blockDim.x) s_wsig[i] = P.wsig[i];
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  // the weight image was packed at the head of the pass (several kernels ago): its load, like everything above,
  // overlaps the previous kernel's tail
  if (warp == 0 && elect_one()) {
    mbar_arrive_expect_tx(w_full, (uint32_t)w_bytes);
    for (int o = 0; o < w_bytes; o += 32768) bulk_g2s(sW + o, P.wimg + o, (uint32_t)min(32768, w_bytes - o), w_full);
  }
  __syncwarp();
  pdl_wait();
  pdl_trigger();
  int my_tiles = 0;
  for (int t = blockIdx.x; t < P.n_tiles; t += gridDim.x) ++my_tiles;
  const int spt = (P.g_slabs + 1) / 2;           // stages per tile
  const int slabs_last = P.g_slabs - 2 * (spt - 1);

  if (warp == 0) {
    if (elect_one()) {
      // G is dead after this kernel (its wgrad ran before), and so is the mask source: evict-first, so that what
      // stays in the L2 is the output tiles, which the next wgrad reads from the high end, where this kernel ends
      const uint64_t pol_g = l2_policy_evict_first();
      int s = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const size_t tile = (size_t)blockIdx.x + (size_t)i * gridDim.x;
        for (int k = 0; k < spt; ++k, ++s) {
          const int st = s % kDgStages, n = s / kDgStages;
          mbar_wait(&empty[st], (n & 1) ^ 1, 1);
          const int ns = k == spt - 1 ? slabs_last : 2;
          mbar_arrive_expect_tx(&full[st], (uint32_t)(ns * kSlab));
          bulk_g2s_hint(ring + st * kDgStageBytes, P.g + (tile * P.g_slabs + 2 * k) * kSlab, (uint32_t)(ns * kSlab), &full[st],
                        pol_g);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_f16(128, (uint32_t)P.k_in, kFmt);
    mbar_wait(w_full, 0, 2);
    int s = 0;
    for (int i = 0; i < my_tiles; ++i) {
      const int buf = i & 1;
      mbar_wait(&d_empty[buf], ((i >> 1) & 1) ^ 1, 3);
      for (int k = 0; k < spt; ++k, ++s) {
        const int st = s % kDgStages, n = s / kDgStages;
        mbar_wait(&full[st], n & 1, 4);
        tc_fence_after_sync();
        if (elect_one()) {
          const int ns = k == spt - 1 ? slabs_last : 2;
          for (int j = 0; j < ns; ++j) {
            const uint32_t a0 = smem_u32(ring + st * kDgStageBytes + j * kSlab);
            const uint32_t b0 = smem_u32(sW + (2 * k + j) * P.k_in * 128);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_ss(tmem + (uint32_t)buf * 256u, make_sdesc_k_sw128(a0 + ks * 32u, 1024), make_sdesc_k_sw128(b0 + ks * 32u, 1024),
                      idesc, (k | j | ks) ? 1u : 0u);
          }
          umma_commit(&empty[st]);
          if (k == spt - 1) umma_commit(&d_full[buf]);
        }
        __syncwarp();
      }
    }
  } else {
    // ---- epilogue: 16 warps = TMEM lane quarter (warp % 4) x 64-column slab of the output (4 for a 256-wide
    // layer below, 2 for a 128-wide one: the other warps idle).  A warp walks its slab in two 32-column halves, so
    // that 18 warps fit the register file (112 registers each); with 8 warps x two slabs the epilogue - a
    // dependent chain accumulator -> mask rows (an L2 / HBM round trip) -> mask, rescale, pack -> store per slab -
    // was what bounded the kernel (~5.5 us per tile against 1.1 us of MMAs).
    const int q = warp & 3, slab = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    if (slab < P.k_in / 64) {
    const float s_in = P.st[kStScale + P.stage];
    // stored units of the input stage -> of the output stage.  The sigma-head term joined below
    // (|d_sigma| <= the top max, st[kStTop]) must fit as well.
    float in_max = P.st[P.stage];
    if (P.dsig) {
      float wm = 0.f;
      for (int i = 0; i < P.k_in; ++i) wm = fmaxf(wm, fabsf(s_wsig[i]));
      in_max = fmaxf(in_max, P.st[kStTop] * s_in * wm);
    }
    const float r = recentre(in_max);
    if (blockIdx.x == 0 && warp == 2 && lane == 0) P.st[kStScale + P.stage + 1] = s_in * r;
    float dbacc[2] = {0.f, 0.f};
    float amax_out = 0.f;
    const uint64_t pol_dead = l2_policy_evict_first();   // the mask source's last reader (its wgrad ran before)
    const uint32_t rx = (uint32_t)(row & 7);
    for (int i = 0; i < my_tiles; ++i) {
      const int buf = i & 1;
      const size_t tile = (size_t)blockIdx.x + (size_t)i * gridDim.x;
      const long long p = (long long)tile * 128 + row;
      const bool valid = p < P.n_points;
      const float ds = (P.dsig && valid) ? __ldg(P.dsig + p) * s_in : 0.f;
      // the mask rows do not depend on the MMAs: requested before the wait for the accumulator
      uint4 m[8];
      if (P.act) {
        const uint4* arow = reinterpret_cast<const uint4*>(
            P.act + ((size_t)tile * P.act_slabs_total + P.act_s0 + slab) * kSlab + row * 128);
#pragma unroll
        for (int k = 0; k < 4; ++k) ldg_row_pair(arow, (uint32_t)k, rx, m[2 * k], m[2 * k + 1], pol_dead);
      }
      uint4* orow = reinterpret_cast<uint4*>(P.out + ((size_t)tile * (P.k_in / 64) + slab) * kSlab + row * 128);
      mbar_wait(&d_full[buf], (i >> 1) & 1, 5);
      tc_fence_after_sync();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int col0 = slab * 64 + h * 32;
        uint32_t v[32];
        tmem_ld_x32(tmem + (uint32_t)buf * 256u + ((uint32_t)(q * 32) << 16) + (uint32_t)col0, v);
        tmem_ld_wait();
        if (h == 1) {   // last read of this accumulator
          tc_fence_before_sync();
          warp_arrive_bar(&d_empty[buf]);
        }
        float x[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
        if (P.dsig) {
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = fmaf(ds, s_wsig[col0 + j], x[j]);
        }
        if (P.act) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint4 mc = m[4 * h + c];
            const uint32_t w4[4] = {mc.x, mc.y, mc.z, mc.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const uint32_t lo = w4[e] & 0xffffu, hi = w4[e] >> 16;
              // saved post-ReLU activation strictly positive <=> the unit was active
              if (!(lo != 0u && lo < 0x8000u)) x[8 * c + 2 * e] = 0.f;
              if (!(hi != 0u && hi < 0x8000u)) x[8 * c + 2 * e + 1] = 0.f;
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          x[j] = valid ? x[j] * r : 0.f;
          amax_out = fmaxf(amax_out, fabsf(x[j]));
        }
        // 16-bit tile of the layer below
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          uint32_t w[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) w[e] = pack2<kFmt, false>(x[16 * k + 2 * e], x[16 * k + 2 * e + 1]);
          stg_row_pair(orow, (uint32_t)(2 * h + k), rx, make_uint4(w[0], w[1], w[2], w[3]), make_uint4(w[4], w[5], w[6], w[7]));
        }
        // bias gradient: column sums over the warp's 32 rows (transposing butterfly, in place)
        if (P.db) dbacc[h] += warp_colsum32(x, lane);
      }
    }
    if (P.db && my_tiles > 0) {
      const float inv = 1.f / (s_in * r);
#pragma unroll
      for (int h = 0; h < 2; ++h) atomicAdd(P.db + slab * 64 + h * 32 + lane, dbacc[h] * inv);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) amax_out = fmaxf(amax_out, __shfl_xor_sync(0xffffffffu, amax_out, d));
    if (lane == 0 && amax_out > 0.f) {
      amax_out = fminf(amax_out, 65504.f);
      atomicMax(reinterpret_cast<unsigned int*>(P.st + P.stage + 1), __float_as_uint(amax_out));
    }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static void bwd_layers(int e_xyz, BwdLayer* L, int* total) {
  // order: trunk layers 1..7, final (8), dir (9), rgb (10)
  int off = 0, n = 0;
  auto add = [&](int wi, int n_out, int k_in, int k0, int ld) {
    L[n] = BwdLayer{wi, n_out, k_in, k0, ld, off};
    off += (n_out / 64) * k_in * 128;
    ++n;
  };
  for (int l = 1; l < 8; ++l) add(l, 256, 256, l == kSkipLayer ? e_xyz : 0, l == kSkipLayer ? e_xyz + 256 : 256);
  add(kLFinal, 256, 256, 0, 256);
  add(kLDir, 128, 256, 0, -1);   // ld patched by the caller (256 + e_dir)
  add(kLRgb, 64, 128, 0, 128);
  *total = off;
}

size_t bwd_packed_bytes(int e_xyz) {
  BwdLayer L[kBwdLayers];
  int total;
  bwd_layers(e_xyz, L, &total);
  return (size_t)total;
}
size_t bwd_tiled_bytes(int64_t n_points, int features) {
  return (size_t)((n_points + 127) / 128) * (size_t)((features + 63) / 64) * kSlab;
}
// scratch of one backward pass: d_rgb fp32, d_sig fp32, two ping-pong gradient buffers (256 wide),
// the rgb / sigma top tiles, the amax word
struct BwdScratch {
  size_t d_rgb, d_sig, g0, g1, g_rgb, g_sig, amax, partial, total;
};
// floats of partial products one CTA writes over a whole pass (13 products; see backward_chain)
constexpr size_t kPartialFloatsPerCta = 1 * 256 + 64 * 128 + 128 * 256 + 128 * 64 + 8 * 256 * 256 + 2 * 256 * 128;
static BwdScratch bwd_scratch(int64_t P) {
  BwdScratch s;
  size_t o = 0;
  auto take = [&](size_t b) {
    const size_t r = o;
    o += (b + 255) / 256 * 256;
    return r;
  };
  s.d_rgb = take((size_t)P * 64 * 4);
  s.d_sig = take((size_t)P * 4);
  s.g0 = take(bwd_tiled_bytes(P, 256));
  s.g1 = take(bwd_tiled_bytes(P, 256));
  s.g_rgb = take(bwd_tiled_bytes(P, 64));
  s.g_sig = take(bwd_tiled_bytes(P, 64));
  s.amax = take(256);
  const int64_t tiles = (P + 127) / 128;
  s.partial = take((size_t)std::min<int64_t>(num_sms(), tiles) * kPartialFloatsPerCta * sizeof(float));
  s.total = o;
  return s;
}
size_t bwd_scratch_bytes(int64_t n_points) { return bwd_scratch(n_points).total; }

// saved-activation buffer of the training forward: slots 0..8 (256 wide), dir (128), embedding (128)
size_t render_acts_bytes(int64_t n_points) {
  return n_points <= 0 ? 0 : (size_t)((n_points + 127) / 128) * (9 * 4 + 2 + 2) * kSlab;
}
// the forward writes one row per valid point; rows past the last point inside the last tile stay
// untouched - zero them, wgrad sums over whole tiles
int render_acts_zero_tail(void* acts, int64_t n_points, cudaStream_t st) {
  const int r0 = (int)(n_points % 128);
  if (n_points <= 0 || r0 == 0) return CRNERF_OK;
  const size_t T = (size_t)((n_points + 127) / 128);
  uint8_t* A = static_cast<uint8_t*>(acts);
  for (int k = 0; k < 11; ++k) {
    const int slabs = k < 9 ? 4 : 2;
    uint8_t* slot = A + (k < 9 ? (size_t)k * T * 4 : (size_t)9 * T * 4 + (size_t)(k - 9) * T * 2) * kSlab;
    zero_tail_kernel<<<slabs, 128, 0, st>>>(slot + (T - 1) * slabs * kSlab, slabs, kSlab, r0);
  }
  count_launch(11);
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

int composite_backward(const float* raw, const float* z, const float* noise, const float* g_feature,
                       const float* g_weights, const float* g_depth, int n_rays, int n_samples,
                       float* d_rgb_pre, float* d_sigma_pre, cudaStream_t st, int split, float* amax);

// launch with programmatic stream serialization (the kernel calls pdl_wait() before its first dependent access)
template <typename Kern, typename Params>
static cudaError_t launch_pdl(Kern kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, const Params& p) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, p);
}

template <int kFmt>
static int run_wgrad(const WgradParams& w, cudaStream_t st) {
  const int ga = w.g_slabs < 2 ? 2 : w.g_slabs;
  const size_t smem = (size_t)kWgStages * (ga + w.nxs) * kHalf + 256;
  CRNERF_CUDA(cudaFuncSetAttribute(wgrad_kernel<kFmt>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = std::min(num_sms(), w.n_tiles);
  CRNERF_CUDA(launch_pdl(wgrad_kernel<kFmt>, dim3(grid), dim3(192), smem, st, w));
  count_launch();
  return CRNERF_OK;
}
template <int kFmt>
static int run_dgrad(const DgradParams& d, cudaStream_t st) {
  const size_t smem = (size_t)d.g_slabs * d.k_in * 128 + kDgStages * kDgStageBytes + 256 + 1024;
  CRNERF_CUDA(cudaFuncSetAttribute(dgrad_kernel<kFmt>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = std::min(num_sms(), d.n_tiles);
  CRNERF_CUDA(launch_pdl(dgrad_kernel<kFmt>, dim3(grid), dim3(kDgThreads), smem, st, d));
  count_launch();
  return CRNERF_OK;
}

template <int kFmt>
static int backward_chain(const crnerf_mlp_weights* w, const void* acts, const float* raw, const float* z,
                          const float* noise, const float* g_feature, const float* g_weights, const float* g_depth,
                          int n_rays, int n_samples, void* bwd_weights, void* scratch, float* const* gw,
                          float* const* gb, cudaStream_t st) {
  const int64_t P = (int64_t)n_rays * n_samples;
  const int T = (int)((P + 127) / 128);
  const BwdScratch sc = bwd_scratch(P);
  uint8_t* base = static_cast<uint8_t*>(scratch);
  float* d_rgb = reinterpret_cast<float*>(base + sc.d_rgb);
  float* d_sig = reinterpret_cast<float*>(base + sc.d_sig);
  uint8_t* gbuf[2] = {base + sc.g0, base + sc.g1};
  uint8_t* g_rgb = base + sc.g_rgb;
  uint8_t* g_sig = base + sc.g_sig;
  float* stw = reinterpret_cast<float*>(base + sc.amax);    // scale state words
  const int e_xyz = w->e_xyz, e_dir = w->e_dir;

  // 1. weights for dgrad (the optimizer changed them since the last step)
  BwdPackParams bp;
  memset(&bp, 0, sizeof(bp));
  for (int i = 0; i < 12; ++i) bp.w[i] = w->weight[i];
  bp.img = static_cast<uint8_t*>(bwd_weights);
  bwd_layers(e_xyz, bp.L, &bp.total_bytes);
  bp.L[8].ld = 256 + e_dir;
  bp.fmt = kFmt;
  bwd_pack_kernel<<<num_sms(), 256, 0, st>>>(bp);
  count_launch();

  // 2. composite backward -> d_rgb, d_sigma (fp32), their max magnitude, top tiles
  //    (the composite kernel also measures the max magnitude of what it writes: the chain's first scale)
  CRNERF_CUDA(cudaMemsetAsync(stw, 0, 32 * sizeof(float), st));
  int rc = composite_backward(raw, z, noise, g_feature, g_weights, g_depth, n_rays, n_samples, d_rgb, d_sig, st, 1,
                              stw + kStTop);
  if (rc) return rc;
  top_pack_kernel<kFmt><<<T, 128, 0, st>>>(d_rgb, d_sig, P, stw, g_rgb, g_sig, gb[kLRgb], gb[kLSigma]);
  count_launch(1);

  // saved activations (tiled16): slots 0..8 (256 wide), 9 = dir (128), 10 = embedding tile (128)
  const uint8_t* A = static_cast<const uint8_t*>(acts);
  auto slot = [&](int k) -> const uint8_t* {
    return A + (k < 9 ? (size_t)k * T * 4 : (size_t)9 * T * 4 + (size_t)(k - 9) * T * 2) * kSlab;
  };
  int stage = 0;   // stage of the gradient tiles currently at the head of the chain
  ReduceParams red;
  memset(&red, 0, sizeof(red));
  const int wg_grid = std::min(num_sms(), T);
  float* part_next = reinterpret_cast<float*>(base + sc.partial);
  auto wgrad = [&](const uint8_t* g, int g_slabs, const uint8_t* x, int xtot, int xs0, int nxs, float* dw, int ldw,
                   int c0, int xcol0, int ncols, int nrows, int x_dead = 0) {
    WgradParams wp{g, x, part_next, g_slabs, xtot, xs0, nxs, nrows, T, x_dead};
    red.job[red.n_jobs++] = ReduceJob{part_next, dw, stw + kStScale + stage, wg_grid, nrows, nxs * 64, ldw, c0, xcol0, ncols};
    part_next += (size_t)wg_grid * nrows * nxs * 64;
    return run_wgrad<kFmt>(wp, st);
  };
  auto dgrad = [&](const uint8_t* g, int g_slabs, int layer_idx, const uint8_t* act, int atot, int as0, uint8_t* out,
                   float* db, const float* dsig) {
    DgradParams dp{g, static_cast<const uint8_t*>(bwd_weights) + bp.L[layer_idx].offset, act, out, db, dsig,
                   w->weight[kLSigma], stw, stage, P, g_slabs, bp.L[layer_idx].k_in, atot, as0, T};
    const int rc2 = run_dgrad<kFmt>(dp, st);
    ++stage;       // the output tiles are the next stage
    return rc2;
  };
#define CK_(x) do { rc = (x); if (rc) return rc; } while (0)
  // static_sigma: dW = d_sigma^T h8 (row 0 of a 64-row product)
  CK_(wgrad(g_sig, 1, slot(7), 4, 0, 4, gw[kLSigma], 256, 0, 0, 256, 1));
  // static_rgb (64 x 128): X = dir_out
  CK_(wgrad(g_rgb, 1, slot(9), 2, 0, 2, gw[kLRgb], 128, 0, 0, 128, 64));
  CK_(dgrad(g_rgb, 1, 9, slot(9), 2, 0, gbuf[0], gb[kLDir], nullptr));                 // -> Gm_dir (P x 128)
  // dir_encoding (128 x (256 + e_dir)): X = [final | dir embedding = columns 96.. of the embedding tile]
  CK_(wgrad(gbuf[0], 2, slot(8), 4, 0, 4, gw[kLDir], 256 + e_dir, 0, 0, 256, 128, 1));
  CK_(wgrad(gbuf[0], 2, slot(10), 2, 1, 1, gw[kLDir], 256 + e_dir, 256, kDirCol0 - 64, e_dir, 128));
  CK_(dgrad(gbuf[0], 2, 8, nullptr, 0, 0, gbuf[1], gb[kLFinal], nullptr));             // -> Gm_final (P x 256)
  // xyz_encoding_final (256 x 256): X = h8; the sigma head joins below it
  CK_(wgrad(gbuf[1], 4, slot(7), 4, 0, 4, gw[kLFinal], 256, 0, 0, 256, 256));
  CK_(dgrad(gbuf[1], 4, 7, slot(7), 4, 0, gbuf[0], gb[7], d_sig));                      // -> Gm of trunk layer 8
  int cur = 0;
  for (int l = 7; l >= 1; --l) {
    // trunk layer l (0-based): X = h_l (slot l-1) [+ xyz embedding in front for the skip layer]
    const int ld = l == kSkipLayer ? e_xyz + 256 : 256, c0 = l == kSkipLayer ? e_xyz : 0;
    CK_(wgrad(gbuf[cur], 4, slot(l - 1), 4, 0, 4, gw[l], ld, c0, 0, 256, 256));
    if (l == kSkipLayer) CK_(wgrad(gbuf[cur], 4, slot(10), 2, 0, 2, gw[l], ld, 0, 0, e_xyz, 256));
    CK_(dgrad(gbuf[cur], 4, l - 1, slot(l - 1), 4, 0, gbuf[cur ^ 1], gb[l - 1], nullptr));
    cur ^= 1;
  }
  CK_(wgrad(gbuf[cur], 4, slot(10), 2, 0, 2, gw[0], e_xyz, 0, 0, e_xyz, 256, 1));       // layer 0: X = xyz embedding
#undef CK_
  // all 13 weight gradients: sum the per-CTA partials (fixed order)
  CRNERF_CUDA(launch_pdl(wgrad_reduce_kernel, dim3(64, red.n_jobs), dim3(256), 0, st, red));
  count_launch();
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

int render_backward(const crnerf_mlp_weights* w, int operand, const void* acts, const float* raw, const float* z,
                    const float* noise, const float* g_feature, const float* g_weights, const float* g_depth,
                    int n_rays, int n_samples, void* bwd_weights, void* scratch, float* const* gw,
                    float* const* gb, cudaStream_t st) {
  if (operand == 0)
    return backward_chain<0>(w, acts, raw, z, noise, g_feature, g_weights, g_depth, n_rays, n_samples, bwd_weights,
                             scratch, gw, gb, st);
  return backward_chain<1>(w, acts, raw, z, noise, g_feature, g_weights, g_depth, n_rays, n_samples, bwd_weights,
                           scratch, gw, gb, st);
}

}  // namespace crnerf

// Cross-ray fusion + decoder: style_net.forward of the reference
// (models/linearStyleTransfer.py:284-291 -> MulLayer.forward :58-90 -> CNN.forward
// :28-37 -> NeuralRenderer.forward, models/nerf_decoder_stylenerf.py:279-291 with
// n_blocks == 0).
//
// The block is linear per pixel once its two 32x32 matrices are known:
//   rgb = sigmoid( Wr ( Wu ( S C ( Wc (x - mu_c) + bc ) ) + bu + mu_s ) + br )
//       = sigmoid( A x + a0 ),  A = Wr Wu (S C) Wc  (3x64)
// where C = fc_c(Gram(cnet(x - mu_c))/HW) needs statistics over ALL pixels (the
// "cross-ray" part) and S the same over the style feature.  So the data path is
//   pass 1  channel sums - normally NOT a pass: the render kernel's epilogue emits per-CTA partial
//           sums of the features it writes (crnerf_render_pass_opts) and the Gram kernel reduces
//           them in its prologue; sums_rows_kernel / sums_kernel exist for maps that come from
//           elsewhere (read the feature map once)
//   pass 2  pixel MLP 64-128-64-32 + 32x32 Gram partials, content and style map in ONE launch
//           (tensor core, gram_tc.cu; walks the map backwards: its tail is what the producer left in L2)
//   tiny    reduce the Gram partials (both maps, one launch), two 1024x1024 GEMVs (one launch)
//   pass 3  compose the 3x64 map in the prologue (18 kFMA per CTA), apply it + sigmoid (second
//           read of the map, forwards: its head is what pass 2 left in L2), write RGB
// = 4 launches.  All reductions use per-block partials combined in a fixed order, so results
// are deterministic run to run.  Feature maps are read in place through
// (pixel stride, channel stride), i.e. both the renderer's (N,64) rows and
// contiguous NCHW.
#include <algorithm>
#include "common.h"

#include "gram_tc.h"

namespace crnerf {
namespace {

constexpr int kC = 64;     // feature channels (MulLayer hard-codes 64, linearStyleTransfer.py:46-47)
constexpr int kTP = 64;    // pixels per tile
constexpr int kTPS = 68;   // padded tile row (float4-aligned)
constexpr int kMaxBlocks = 296;


// tile loader: xin[c][px] = x[p0+px][c] - mean[c]   (zero for px beyond n)
__device__ __forceinline__ void load_tile(const float* __restrict__ g, long long n, long long p0,
                                          long long pix_stride, long long ch_stride,
                                          const float* mean, float (*xin)[kTPS]) {
  const int tid = threadIdx.x;
  if (ch_stride == 1) {  // rows layout: a pixel's 64 channels are contiguous
    const int c = tid & 63, pl = tid >> 6;
    const float m = mean ? mean[c] : 0.f;
#pragma unroll 4
    for (int pp = 0; pp < kTP / 4; ++pp) {
      const int px = pl + 4 * pp;
      const long long p = p0 + px;
      xin[c][px] = p < n ? __ldg(g + p * pix_stride + c) - m : 0.f;
    }
  } else {  // planar (NCHW) or generic strides: consecutive threads -> consecutive pixels
    const int px = tid & 63, cl = tid >> 6;
    const long long p = p0 + px;
#pragma unroll 4
    for (int cc = 0; cc < kC / 4; ++cc) {
      const int c = cl + 4 * cc;
      xin[c][px] = p < n ? __ldg(g + p * pix_stride + c * ch_stride) - (mean ? mean[c] : 0.f) : 0.f;
    }
  }
}

// ---------------------------------------------------------------- pass 1
// partial[block][c] = sum over this block's pixels of x[p][c]
__global__ void __launch_bounds__(256)
sums_kernel(const float* __restrict__ g, long long n, long long pix_stride, long long ch_stride,
            float* __restrict__ partial) {
  __shared__ float xin[kC][kTPS];
  __shared__ float red[4][kC];
  const int tid = threadIdx.x;
  const int c = tid & 63, pl = tid >> 6;
  float acc = 0.f;
  const long long n_tiles = (n + kTP - 1) / kTP;
  for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    load_tile(g, n, t * kTP, pix_stride, ch_stride, nullptr, xin);
    __syncthreads();
#pragma unroll
    for (int pp = 0; pp < kTP / 4; ++pp) acc += xin[c][pl * (kTP / 4) + pp];
    __syncthreads();
  }
  red[pl][c] = acc;
  __syncthreads();
  if (tid < kC) partial[blockIdx.x * kC + tid] = (red[0][tid] + red[1][tid]) + (red[2][tid] + red[3][tid]);
}

// Same for the renderer's row layout (a pixel's 64 channels contiguous, 16-byte aligned): a pure
// HBM stream.  16 threads cover one pixel with float4 loads, 16 pixels per pass, 8 passes in
// flight per thread; fixed-order block reduction.
__global__ void __launch_bounds__(256)
sums_rows_kernel(const float* __restrict__ g, long long n, long long pix_stride, float* __restrict__ partial) {
  __shared__ float4 red[16][16];
  const int c4 = threadIdx.x & 15, rl = threadIdx.x >> 4;
  const long long per = (n + gridDim.x - 1) / gridDim.x;
  const long long r0 = blockIdx.x * per, r1 = min(n, r0 + per);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  long long r = r0 + rl;
  for (; r + 7 * 16 < r1; r += 8 * 16) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(g + (r + 16 * u) * pix_stride) + c4);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
    }
  }
  for (; r < r1; r += 16) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(g + r * pix_stride) + c4);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  red[rl][c4] = acc;
  __syncthreads();
  if (threadIdx.x < 64) {
    const int c = threadIdx.x;
    float t = 0.f;
    for (int k = 0; k < 16; ++k) t += reinterpret_cast<const float*>(&red[k][c >> 2])[c & 3];
    partial[blockIdx.x * kC + c] = t;
  }
}

// out[i] = scale * sum_b partial[b][i], fixed order
__global__ void reduce_partials_kernel(const float* __restrict__ partial, int n_parts, int len,
                                       float scale, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  float acc = 0.f;
  for (int b = 0; b < n_parts; ++b) acc += partial[(long long)b * len + i];
  out[i] = acc * scale;
}

// same result layout, for many partials: one warp per output element, lanes stride the partials
// (independent loads in flight) and combine in a fixed shuffle tree - still deterministic
__global__ void __launch_bounds__(256)
reduce_partials_warp_kernel(const float* __restrict__ partial, int n_parts, int len, float scale,
                            float* __restrict__ out) {
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= len) return;
  float acc = 0.f;
  for (int b = lane; b < n_parts; b += 32) acc += partial[(long long)b * len + i];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
  if (lane == 0) out[i] = acc * scale;
}

// ---------------------------------------------------------------- pass 2
// pixel MLP 64 -> 128 -> 64 -> 32 + Gram partials: gram_tc.cu (tcgen05)

// ---------------------------------------------------------------- tiny stage
// rows [0,1024): cnet.fc(gram_c); rows [1024,2048): snet.fc(gram_s).  One warp per row.
__global__ void __launch_bounds__(256)
fc_kernel(const float* __restrict__ wc, const float* __restrict__ bc, const float* __restrict__ gc,
          const float* __restrict__ ws, const float* __restrict__ bs, const float* __restrict__ gs,
          float* __restrict__ out_c, float* __restrict__ out_s) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  const bool is_s = row >= 1024;
  const int r = is_s ? row - 1024 : row;
  const float* W = (is_s ? ws : wc) + (long long)r * 1024;
  const float* v = is_s ? gs : gc;
  if (v == nullptr) return;
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(W) + lane + 32 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(v) + lane + 32 * i);
    acc = fmaf(a.x, b.x, acc);
    acc = fmaf(a.y, b.y, acc);
    acc = fmaf(a.z, b.z, acc);
    acc = fmaf(a.w, b.w, acc);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
  if (lane == 0) (is_s ? out_s : out_c)[r] = acc + (is_s ? bs : bc)[r];
}

// Compose the per-pixel linear map.  Layout of `map` (floats):
//   [0,192) A (3x64) | [192,195) a0 | [256, 256+4096) M64 | [4352, 4416) v64 (fused offset)
constexpr int kMapA = 0, kMapA0 = 192, kMapM = 256, kMapV = 4352, kMapFloats = 4416;

__global__ void __launch_bounds__(256)
compose_kernel(const float* __restrict__ cmat, const float* __restrict__ smat,
               crnerf_style_weights w, const float* __restrict__ mean_c,
               const float* __restrict__ mean_s, float* __restrict__ map,
               float* __restrict__ trans_out) {
  __shared__ float T[32][33], TW[32][65], Tb[32], M[64][65], V[64];
  const int tid = threadIdx.x;
  // T = S C  (MulLayer.forward :86)
  for (int e = tid; e < 1024; e += 256) {
    const int i = e >> 5, j = e & 31;
    float acc = 0.f;
    for (int k = 0; k < 32; ++k) acc = fmaf(smat[i * 32 + k], cmat[k * 32 + j], acc);
    T[i][j] = acc;
    if (trans_out) trans_out[e] = acc;
  }
  __syncthreads();
  // T Wc, T bc   (compress, :54,:76)
  for (int e = tid; e < 32 * 64; e += 256) {
    const int i = e >> 6, c = e & 63;
    float acc = 0.f;
    for (int k = 0; k < 32; ++k) acc = fmaf(T[i][k], w.compress_w[k * 64 + c], acc);
    TW[i][c] = acc;
  }
  if (tid < 32) {
    float acc = 0.f;
    for (int k = 0; k < 32; ++k) acc = fmaf(T[tid][k], w.compress_b[k], acc);
    Tb[tid] = acc;
  }
  __syncthreads();
  // M = Wu (T Wc), v = Wu (T bc) + bu + mu_s   (unzip :55,:88 and "+ sMeanC" :89)
  for (int e = tid; e < 64 * 64; e += 256) {
    const int o = e >> 6, c = e & 63;
    float acc = 0.f;
    for (int i = 0; i < 32; ++i) acc = fmaf(w.unzip_w[o * 32 + i], TW[i][c], acc);
    M[o][c] = acc;
  }
  if (tid < 64) {
    float acc = 0.f;
    for (int i = 0; i < 32; ++i) acc = fmaf(w.unzip_w[tid * 32 + i], Tb[i], acc);
    V[tid] = acc + w.unzip_b[tid] + mean_s[tid];
  }
  __syncthreads();
  // fold the content mean: M (x - mu_c) + v = M x + (v - M mu_c)
  if (tid < 64) {
    float acc = 0.f;
    for (int c = 0; c < 64; ++c) acc = fmaf(M[tid][c], mean_c[c], acc);
    V[tid] -= acc;
  }
  __syncthreads();
  for (int e = tid; e < 4096; e += 256) map[kMapM + e] = M[e >> 6][e & 63];
  if (tid < 64) map[kMapV + tid] = V[tid];
  // A = Wr M, a0 = Wr v + br   (NeuralRenderer.forward, nerf_decoder_stylenerf.py:280)
  if (tid < 192) {
    const int r = tid >> 6, c = tid & 63;
    float acc = 0.f;
    for (int o = 0; o < 64; ++o) acc = fmaf(w.rgb_w[r * 64 + o], M[o][c], acc);
    map[kMapA + tid] = acc;
  }
  if (tid >= 192 && tid < 195) {
    const int r = tid - 192;
    float acc = 0.f;
    for (int o = 0; o < 64; ++o) acc = fmaf(w.rgb_w[r * 64 + o], V[o], acc);
    map[kMapA0 + r] = acc + w.rgb_b[r];
  }
}

// Gram partials of both maps -> normalised Gram vectors, one launch: output e < 1024 belongs to
// job 0, e >= 1024 to job 1 (absent when part1 == nullptr).  One warp per element, lanes stride
// the partials (independent loads in flight), fixed shuffle tree.
__global__ void __launch_bounds__(256)
reduce_gram_kernel(const float* __restrict__ part0, int n0, float scale0, float* __restrict__ out0,
                   const float* __restrict__ part1, int n1, float scale1, float* __restrict__ out1) {
  const int e = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  const bool second = e >= 1024;
  const float* part = second ? part1 : part0;
  if (part == nullptr) return;
  const int n_parts = second ? n1 : n0, i = e & 1023;
  float acc = 0.f;
  for (int b = lane; b < n_parts; b += 32) acc += part[(long long)b * 1024 + i];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
  if (lane == 0) (second ? out1 : out0)[i] = acc * (second ? scale1 : scale0);
}

// ---------------------------------------------------------------- pass 3
// The rgb map of the block, composed right to left so that no 64x64 matrix is ever formed:
//   A  = Wr Wu S C Wc                      = ((((Wr Wu) S) C) Wc)                 (3x64, 18 kFMA)
//   a0 = R3 bc + Wr (bu + mu_s) + br - A mu_c,   R3 = Wr Wu S C                   (3)
// (MulLayer.forward :76-89 + NeuralRenderer.forward).  Every apply CTA runs this in its prologue
// from the two FC outputs; both apply kernels share it, so they use bit-identical maps.
struct MapSrc {
  const float* cmat;     // cnet.fc output (1024) = C row-major; nullptr: decoder only (A = Wr, a0 = br)
  const float* smat;     // snet.fc output = S
  const float* mean_c;
  const float* mean_s;
  crnerf_style_weights w;
  float* trans_out;      // optional (1024): T = S C, written by block 0
};

struct MapSmem {
  float S[32][33], C[32][33], Wu[64][33], Wc[32][65], Wr[3][64];
  float R[2][3][32];
  float A[3][64], a0[3];
};

__device__ void compose_rgb_map(const MapSrc& m, MapSmem& sm) {
  const int tid = threadIdx.x, nt = blockDim.x;
  if (m.cmat == nullptr) {
    for (int e = tid; e < 192; e += nt) sm.A[e >> 6][e & 63] = m.w.rgb_w[e];
    if (tid < 3) sm.a0[tid] = m.w.rgb_b[tid];
    __syncthreads();
    return;
  }
  for (int e = tid; e < 1024; e += nt) {
    sm.S[e >> 5][e & 31] = m.smat[e];
    sm.C[e >> 5][e & 31] = m.cmat[e];
  }
  for (int e = tid; e < 2048; e += nt) {
    sm.Wu[e >> 5][e & 31] = m.w.unzip_w[e];      // (64, 32)
    sm.Wc[e >> 6][e & 63] = m.w.compress_w[e];   // (32, 64)
  }
  for (int e = tid; e < 192; e += nt) sm.Wr[e >> 6][e & 63] = m.w.rgb_w[e];
  __syncthreads();
  if (m.trans_out && blockIdx.x == 0) {          // transmatrix = S C (MulLayer.forward :86)
    for (int e = tid; e < 1024; e += nt) {
      const int i = e >> 5, j = e & 31;
      float acc = 0.f;
#pragma unroll 8
      for (int k = 0; k < 32; ++k) acc = fmaf(sm.S[i][k], sm.C[k][j], acc);
      m.trans_out[e] = acc;
    }
  }
  const int r = tid >> 5, i = tid & 31;          // threads 0..95: one element of a 3x32 row block
  if (tid < 96) {                                // R1 = Wr Wu
    float acc = 0.f;
#pragma unroll 8
    for (int o = 0; o < 64; ++o) acc = fmaf(sm.Wr[r][o], sm.Wu[o][i], acc);
    sm.R[0][r][i] = acc;
  }
  __syncthreads();
  if (tid < 96) {                                // R2 = R1 S
    float acc = 0.f;
#pragma unroll 8
    for (int k = 0; k < 32; ++k) acc = fmaf(sm.R[0][r][k], sm.S[k][i], acc);
    sm.R[1][r][i] = acc;
  }
  __syncthreads();
  if (tid < 96) {                                // R3 = R2 C
    float acc = 0.f;
#pragma unroll 8
    for (int k = 0; k < 32; ++k) acc = fmaf(sm.R[1][r][k], sm.C[k][i], acc);
    sm.R[0][r][i] = acc;
  }
  __syncthreads();
  if (tid < 192) {                               // A = R3 Wc
    const int rr = tid >> 6, c = tid & 63;
    float acc = 0.f;
#pragma unroll 8
    for (int k = 0; k < 32; ++k) acc = fmaf(sm.R[0][rr][k], sm.Wc[k][c], acc);
    sm.A[rr][c] = acc;
  }
  __syncthreads();
  if (tid < 3) {
    float acc = 0.f;
    for (int k = 0; k < 32; ++k) acc = fmaf(sm.R[0][tid][k], m.w.compress_b[k], acc);
    float acc2 = 0.f;
    for (int o = 0; o < 64; ++o) acc2 = fmaf(sm.Wr[tid][o], m.w.unzip_b[o] + m.mean_s[o], acc2);
    float acc3 = 0.f;
    for (int c = 0; c < 64; ++c) acc3 = fmaf(sm.A[tid][c], m.mean_c[c], acc3);
    sm.a0[tid] = ((acc + acc2) + m.w.rgb_b[tid]) - acc3;
  }
  __syncthreads();
}

// rgb[r][p] = sigmoid(A[r] . x[p] + a0[r]); optional fused[o][p] = M[o] . x[p] + v[o] (M, v from
// compose_kernel's `map`).  Any layout; the row layout normally takes apply_rows_kernel.
__global__ void __launch_bounds__(256)
apply_kernel(const float* __restrict__ g, long long n, long long pix_stride, long long ch_stride,
             const __grid_constant__ MapSrc src, const float* __restrict__ map, float* __restrict__ rgb,
             float* __restrict__ fused) {
  __shared__ float xin[kC][kTPS];
  __shared__ MapSmem sm;
  const int tid = threadIdx.x;
  compose_rgb_map(src, sm);
  const long long n_tiles = (n + kTP - 1) / kTP;
  for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const long long p0 = t * kTP;
    __syncthreads();
    load_tile(g, n, p0, pix_stride, ch_stride, nullptr, xin);
    __syncthreads();
    if (tid < 3 * kTP) {
      const int r = tid / kTP, px = tid % kTP;
      float acc = 0.f;
#pragma unroll 8
      for (int c = 0; c < 64; ++c) acc = fmaf(sm.A[r][c], xin[c][px], acc);
      acc += sm.a0[r];
      if (p0 + px < n) rgb[(long long)r * n + p0 + px] = 1.f / (1.f + expf(-acc));
    }
    if (fused) {
      for (int e = tid; e < 64 * kTP; e += 256) {
        const int o = e / kTP, px = e % kTP;
        float acc = 0.f;
        for (int c = 0; c < 64; ++c) acc = fmaf(map[kMapM + o * 64 + c], xin[c][px], acc);
        if (p0 + px < n) fused[(long long)o * n + p0 + px] = acc + map[kMapV + o];
      }
    }
  }
}

// pass 3 for the renderer's row layout: a pure HBM stream.  16 lanes cover one pixel with one
// float4 load each (a warp-wide load = two whole 256-byte rows), 8 such loads in flight per
// thread; every lane keeps its 4 channels of A in registers.  The 16 partial dot products of a
// pixel are combined by a transposing butterfly: each exchange halves the number of (pixel, r)
// values a lane still owns, 24 shuffles for 16 pixels x 3 outputs, fixed order.  Persistent
// blocks sweep the map front to back.
__global__ void __launch_bounds__(256, 2)
apply_rows_kernel(const float* __restrict__ g, long long n, long long pix_stride,
                  const __grid_constant__ MapSrc src, float* __restrict__ rgb) {
  __shared__ MapSmem sm;
  compose_rgb_map(src, sm);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int l = lane & 15, half = lane >> 4;
  float a[3][4];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int k = 0; k < 4; ++k) a[r][k] = sm.A[r][4 * l + k];
  const float a0[3] = {sm.a0[0], sm.a0[1], sm.a0[2]};
  const bool b3 = l & 8, b2 = l & 4, b1 = l & 2;
  // software pipeline: the next step's 8 loads are in flight while this step's 16 pixels are reduced
  const long long step = (long long)gridDim.x * 128;
  auto load16 = [&](long long base, float4 (&x)[8]) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const long long p = base + 2 * u + half;
      x[u] = p < n ? __ldcs(reinterpret_cast<const float4*>(g + p * pix_stride) + l) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  long long base = ((long long)blockIdx.x * 8 + warp) * 16;
  float4 x[8], xn[8];
  if (base < n) load16(base, x);
  for (; base < n; base += step) {
    if (base + step < n) load16(base + step, xn);
    float v[8][3];
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int r = 0; r < 3; ++r)
        v[u][r] = fmaf(a[r][3], x[u].w, fmaf(a[r][2], x[u].z, fmaf(a[r][1], x[u].y, a[r][0] * x[u].x)));
    // xor 8: lanes with bit 3 clear keep u 0..3, the others u 4..7
    float w4[4][3];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const float keep = b3 ? v[u + 4][r] : v[u][r], send = b3 ? v[u][r] : v[u + 4][r];
        w4[u][r] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
    float w2[2][3];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const float keep = b2 ? w4[u + 2][r] : w4[u][r], send = b2 ? w4[u][r] : w4[u + 2][r];
        w2[u][r] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
    float w1[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float keep = b1 ? w2[1][r] : w2[0][r], send = b1 ? w2[0][r] : w2[1][r];
      w1[r] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) w1[r] += __shfl_xor_sync(0xffffffffu, w1[r], 1);
    // this lane pair owns u = (bit3, bit2, bit1) of the lane index
    const long long p = base + 2 * ((l >> 1) & 7) + half;
    if ((l & 1) == 0 && p < n) {
#pragma unroll
      for (int r = 0; r < 3; ++r) rgb[(long long)r * n + p] = 1.f / (1.f + expf(-(w1[r] + a0[r])));
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) x[u] = xn[u];
  }
}

static bool rows_fast(const float* g, long long ps, long long cs) {
  return cs == 1 && (ps & 3) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0;
}

int blocks_for_pixels(long long n) {
  long long tiles = (n + kTP - 1) / kTP;
  long long cap = std::min<long long>(kMaxBlocks, 2LL * num_sms());
  return (int)std::max<long long>(1, std::min(tiles, cap));
}

}  // namespace

// scratch layout (floats)
//   [0, P1)            sums partials   kMaxBlocks*64
//   [P1, P1+P2)        gram partials   kMaxBlocks*1024 (content job: first half, style job: second half)
//   then mean_c 64 | mean_s 64 | gram_c 1024 | gram_s 1024 | cmat 1024 | smat 1024 | map 4416
constexpr size_t kOffSumPart = 0;
constexpr size_t kOffGramPart = kOffSumPart + (size_t)kMaxBlocks * 64;
constexpr size_t kOffMeanC = kOffGramPart + (size_t)kMaxBlocks * 1024;
constexpr size_t kOffMeanS = kOffMeanC + 64;
constexpr size_t kOffGramC = kOffMeanS + 64;
constexpr size_t kOffGramS = kOffGramC + 1024;
constexpr size_t kOffCmat = kOffGramS + 1024;
constexpr size_t kOffSmat = kOffCmat + 1024;
constexpr size_t kOffMap = kOffSmat + 1024;
constexpr size_t kScratchFloats = kOffMap + kMapFloats;
constexpr int kGramBlocksPerJob = kMaxBlocks / 2;
constexpr long long kSelfMeanMaxPixels = 4096;   // maps this small: every Gram CTA sums them itself

size_t style_scratch_floats(int64_t) { return kScratchFloats; }

static int check_feat(const float* p, int64_t n, int64_t ps, int64_t cs, const char* what) {
  CRNERF_REQUIRE(p != nullptr, "%s is null", what);
  CRNERF_REQUIRE(n >= 1, "%s has no pixels", what);
  CRNERF_REQUIRE(ps >= 1 && cs >= 1, "%s strides must be positive", what);
  return CRNERF_OK;
}

// per-block partial channel sums of a map -> partial (n_blocks, 64); returns n_blocks
static int sum_partials(const float* x, int64_t n, int64_t ps, int64_t cs, float* partial, int* n_blocks,
                        cudaStream_t st, int cap = kMaxBlocks) {
  int nb = std::min(cap, blocks_for_pixels(n));
  if (rows_fast(x, ps, cs)) {
    nb = (int)std::max<long long>(1, std::min<long long>((n + 127) / 128, cap));
    sums_rows_kernel<<<nb, 256, 0, st>>>(x, n, ps, partial);
  } else {
    sums_kernel<<<nb, 256, 0, st>>>(x, n, ps, cs, partial);
  }
  count_launch();
  CRNERF_CUDA(cudaGetLastError());
  *n_blocks = nb;
  return CRNERF_OK;
}

int style_stats1(const float* content, int64_t n, int64_t ps, int64_t cs, float* sums,
                 float* partial, cudaStream_t st) {
  int nb = 0;
  int rc = sum_partials(content, n, ps, cs, partial, &nb, st);
  if (rc) return rc;
  reduce_partials_warp_kernel<<<8, 256, 0, st>>>(partial, nb, 64, 1.f, sums);
  count_launch();
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

// sums[len] = sum of the rows of parts (n_parts, len) in a fixed order (the render kernel's
// per-CTA channel sums -> the 64 channel sums of a rank's block of rays)
int sum_rows(const float* parts, int n_parts, int len, float* out, cudaStream_t st) {
  CRNERF_REQUIRE(parts && out && n_parts >= 1 && len >= 1, "bad argument");
  reduce_partials_warp_kernel<<<(len + 7) / 8, 256, 0, st>>>(parts, n_parts, len, 1.f, out);
  count_launch();
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

static GramJob make_job(const crnerf_cnn_weights& cw, const float* x, int64_t n, int64_t ps, int64_t cs,
                        const float* sum_parts, int n_parts, float mean_scale, float* partial, float* mean_out,
                        int reverse) {
  GramJob j{};
  j.g = x;
  j.n = n;
  j.pix_stride = ps;
  j.ch_stride = cs;
  j.sum_parts = sum_parts;
  j.n_parts = n_parts;
  j.mean_scale = mean_scale;
  j.self_mean = sum_parts == nullptr;
  j.w = cw;
  j.partial = partial;
  j.mean_out = mean_out;
  j.reverse = reverse;
  return j;
}

int style_stats2(const crnerf_cnn_weights& cw, const float* content, int64_t n, int64_t ps,
                 int64_t cs, const float* mean, float* gram, float* partial, float scale,
                 cudaStream_t st) {
  // pixel MLP + Gram on the tensor core (gram_tc.cu), then the fixed-order sum of the per-block partials
  GramJob j = make_job(cw, content, n, ps, cs, mean, 1, 1.f, partial, nullptr, 1);
  int rc = gram_tc_launch(&j, 1, kGramBlocksPerJob, st);
  if (rc) return rc;
  reduce_gram_kernel<<<128, 256, 0, st>>>(partial, j.n_blocks, scale, gram, nullptr, 0, 0.f, nullptr);
  count_launch();
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

// FC GEMVs + the apply pass, given both normalised Gram vectors and both means
static int fc_and_apply(const crnerf_style_weights* w, const float* content, int64_t n, int64_t ps, int64_t cs,
                        const float* mean_c, const float* gram_c, const float* mean_s, const float* gram_s,
                        float* rgb, float* transmatrix, float* fused, float* scratch, cudaStream_t st) {
  float* cmat = scratch + kOffCmat;
  float* smat = scratch + kOffSmat;
  float* map = scratch + kOffMap;
  fc_kernel<<<256, 256, 0, st>>>(w->cnet.fc_w, w->cnet.fc_b, gram_c, w->snet.fc_w, w->snet.fc_b, gram_s, cmat, smat);
  count_launch();
  MapSrc src{cmat, smat, mean_c, mean_s, *w, transmatrix};
  if (fused == nullptr && rows_fast(content, ps, cs)) {
    const int grid = (int)std::max<long long>(1, std::min<long long>((n + 127) / 128, 4LL * num_sms()));
    apply_rows_kernel<<<grid, 256, 0, st>>>(content, n, ps, src, rgb);
    count_launch();
  } else {
    if (fused) {   // the 64x64 map of the fused feature itself (tests / debugging output)
      compose_kernel<<<1, 256, 0, st>>>(cmat, smat, *w, mean_c, mean_s, map, nullptr);
      count_launch();
    }
    apply_kernel<<<blocks_for_pixels(n), 256, 0, st>>>(content, n, ps, cs, src, map, rgb, fused);
    count_launch();
  }
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

// everything after the content Gram is known (sharded path: mean and Gram were all-reduced):
// style branch, FCs, apply
int style_finish(const crnerf_style_weights* w, const float* content, int64_t n, int64_t ps,
                 int64_t cs, const float* mean_c, const float* gram_c_normalised,
                 const float* style, int64_t ns, int64_t sps, int64_t scs, float* rgb,
                 float* transmatrix, float* fused, float* scratch, cudaStream_t st) {
  float* gram_part = scratch + kOffGramPart;
  float* mean_s = scratch + kOffMeanS;
  float* gram_s = scratch + kOffGramS;
  int rc;
  GramJob js;
  if (ns <= kSelfMeanMaxPixels) {
    js = make_job(w->snet, style, ns, sps, scs, nullptr, 0, 1.f / (float)ns, gram_part, mean_s, 0);
  } else {
    int nb = 0;
    rc = sum_partials(style, ns, sps, scs, scratch + kOffSumPart, &nb, st);
    if (rc) return rc;
    js = make_job(w->snet, style, ns, sps, scs, scratch + kOffSumPart, nb, 1.f / (float)ns, gram_part, mean_s, 1);
  }
  rc = gram_tc_launch(&js, 1, kGramBlocksPerJob, st);
  if (rc) return rc;
  reduce_gram_kernel<<<128, 256, 0, st>>>(gram_part, js.n_blocks, 1.f / (float)ns, gram_s, nullptr, 0, 0.f, nullptr);
  count_launch();
  return fc_and_apply(w, content, n, ps, cs, mean_c, gram_c_normalised, mean_s, gram_s, rgb, transmatrix, fused,
                      scratch, st);
}

// CNN.forward alone (linearStyleTransfer.py:28-37): convs -> Gram/(h*w) -> fc, no mean removal
int cnn_forward(const crnerf_cnn_weights* cw, const float* x, int64_t n, int64_t ps, int64_t cs,
                float* out, float* scratch, cudaStream_t st) {
  CRNERF_REQUIRE(cw && out && scratch, "null argument");
  int rc = check_feat(x, n, ps, cs, "x");
  if (rc) return rc;
  float* zero_mean = scratch + kOffMeanC;
  float* gram = scratch + kOffGramC;
  CRNERF_CUDA(cudaMemsetAsync(zero_mean, 0, 64 * sizeof(float), st));
  rc = style_stats2(*cw, x, n, ps, cs, zero_mean, gram, scratch + kOffGramPart, 1.f / (float)n, st);
  if (rc) return rc;
  fc_kernel<<<128, 256, 0, st>>>(cw->fc_w, cw->fc_b, gram, nullptr, nullptr, nullptr, out, nullptr);
  count_launch();
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

// style_net.forward.  content_sum_parts (n_parts, 64), optional: partial channel sums of the
// content map whose rows add up to its channel sums (the render kernel's epilogue emits them);
// without them the map is read once more to form them.
int style_forward(const crnerf_style_weights* w, const float* content, int64_t n, int64_t ps,
                  int64_t cs, const float* style, int64_t ns, int64_t sps, int64_t scs,
                  const float* content_sum_parts, int n_parts, float* rgb,
                  float* transmatrix, float* fused, float* scratch, cudaStream_t st) {
  CRNERF_REQUIRE(w && rgb && scratch, "null argument");
  int rc = check_feat(content, n, ps, cs, "content");
  if (rc) return rc;
  if (style == nullptr) {   // decoder only (type == "content", linearStyleTransfer.py:285-287)
    MapSrc src{nullptr, nullptr, nullptr, nullptr, *w, nullptr};
    if (rows_fast(content, ps, cs)) {
      const int grid = (int)std::max<long long>(1, std::min<long long>((n + 127) / 128, 4LL * num_sms()));
      apply_rows_kernel<<<grid, 256, 0, st>>>(content, n, ps, src, rgb);
    } else {
      apply_kernel<<<blocks_for_pixels(n), 256, 0, st>>>(content, n, ps, cs, src, nullptr, rgb, nullptr);
    }
    count_launch();
    CRNERF_CUDA(cudaGetLastError());
    return CRNERF_OK;
  }
  rc = check_feat(style, ns, sps, scs, "style");
  if (rc) return rc;
  CRNERF_REQUIRE(content_sum_parts == nullptr || n_parts >= 1, "content_sum_parts without rows");
  float* mean_c = scratch + kOffMeanC;
  float* mean_s = scratch + kOffMeanS;
  float* gram_c = scratch + kOffGramC;
  float* gram_s = scratch + kOffGramS;
  float* gram_part = scratch + kOffGramPart;
  // channel means: from the caller's partial sums, by the Gram CTAs themselves (small maps), or
  // from one extra pass over the map
  GramJob jobs[2];
  int nb = 0;
  if (content_sum_parts == nullptr && n > kSelfMeanMaxPixels) {
    rc = sum_partials(content, n, ps, cs, scratch + kOffSumPart, &nb, st, kMaxBlocks / 2);
    if (rc) return rc;
    content_sum_parts = scratch + kOffSumPart;
    n_parts = nb;
  }
  jobs[0] = make_job(w->cnet, content, n, ps, cs, content_sum_parts, n_parts, 1.f / (float)n, gram_part, mean_c, 1);
  if (ns <= kSelfMeanMaxPixels) {
    jobs[1] = make_job(w->snet, style, ns, sps, scs, nullptr, 0, 1.f / (float)ns,
                       gram_part + (size_t)kGramBlocksPerJob * 1024, mean_s, 0);
  } else {
    float* sp = scratch + kOffSumPart + (size_t)(kMaxBlocks / 2) * 64;   // second half of the sums region
    rc = sum_partials(style, ns, sps, scs, sp, &nb, st, kMaxBlocks / 2);
    if (rc) return rc;
    jobs[1] = make_job(w->snet, style, ns, sps, scs, sp, nb, 1.f / (float)ns,
                       gram_part + (size_t)kGramBlocksPerJob * 1024, mean_s, 1);
  }
  rc = gram_tc_launch(jobs, 2, kGramBlocksPerJob, st);
  if (rc) return rc;
  reduce_gram_kernel<<<256, 256, 0, st>>>(jobs[0].partial, jobs[0].n_blocks, 1.f / (float)n, gram_c,
                                          jobs[1].partial, jobs[1].n_blocks, 1.f / (float)ns, gram_s);
  count_launch();
  return fc_and_apply(w, content, n, ps, cs, mean_c, gram_c, mean_s, gram_s, rgb, transmatrix, fused, scratch, st);
}

// Training forward of the block: style_forward + what the hand-written backward (style_backward.cu)
// needs from it, copied out of the scratch buffer: aux = [mean_c 64 | mean_s 64 | gram_c 1024 | gram_s 1024 |
// cmat 1024 | smat 1024] (one contiguous run of the scratch layout) | T = S C (1024).
size_t style_aux_floats();
int style_forward_train(const crnerf_style_weights* w, const float* content, int64_t n, int64_t ps, int64_t cs,
                        const float* style, int64_t ns, int64_t sps, int64_t scs, const float* content_sum_parts,
                        int n_parts, float* rgb, float* aux, float* scratch, cudaStream_t st) {
  CRNERF_REQUIRE(aux != nullptr && style != nullptr, "the training forward needs a style map and an aux buffer");
  static_assert(kOffMeanS == kOffMeanC + 64 && kOffGramC == kOffMeanS + 64 && kOffGramS == kOffGramC + 1024 &&
                kOffCmat == kOffGramS + 1024 && kOffSmat == kOffCmat + 1024, "aux is one contiguous run of the scratch");
  int rc = style_forward(w, content, n, ps, cs, style, ns, sps, scs, content_sum_parts, n_parts, rgb, aux + 4224,
                         nullptr, scratch, st);
  if (rc) return rc;
  CRNERF_CUDA(cudaMemcpyAsync(aux, scratch + kOffMeanC, 4224 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return CRNERF_OK;
}

}  // namespace crnerf

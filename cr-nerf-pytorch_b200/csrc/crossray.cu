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
//   pass 1  channel sums               (read feature map once)
//   pass 2  pixel MLP 64-128-64-32 + 32x32 Gram partials (read it again; tensor core, gram_tc.cu)
//   tiny    partial reductions, two 1024x1024 GEMVs, compose A / a0
//   pass 3  apply the 3x64 map + sigmoid (read it a third time, write RGB)
// All reductions use per-block partials combined in a fixed order, so results
// are deterministic run to run.  Feature maps are read in place through
// (pixel stride, channel stride), i.e. both the renderer's (N,64) rows and
// contiguous NCHW.
#include <algorithm>
#include "common.h"

namespace crnerf {
// gram_tc.cu
int gram_tc(const crnerf_cnn_weights& cw, const float* g, int64_t n, int64_t ps, int64_t cs, const float* mean,
            float* partial, int max_blocks, int* n_blocks, cudaStream_t st);
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

// decoder-only map (style is None and type == "content", linearStyleTransfer.py:285-287)
__global__ void content_map_kernel(crnerf_style_weights w, float* __restrict__ map) {
  const int tid = threadIdx.x;
  if (tid < 192) map[kMapA + tid] = w.rgb_w[tid];
  if (tid < 3) map[kMapA0 + tid] = w.rgb_b[tid];
}

// ---------------------------------------------------------------- pass 3
// rgb[r][p] = sigmoid(A[r] . x[p] + a0[r]); optional fused[o][p] = M[o] . x[p] + v[o]
__global__ void __launch_bounds__(256)
apply_kernel(const float* __restrict__ g, long long n, long long pix_stride, long long ch_stride,
             const float* __restrict__ map, float* __restrict__ rgb, float* __restrict__ fused) {
  __shared__ float xin[kC][kTPS];
  __shared__ float A[3][64], a0[3];
  const int tid = threadIdx.x;
  if (tid < 192) A[tid >> 6][tid & 63] = map[kMapA + tid];
  if (tid < 3) a0[tid] = map[kMapA0 + tid];
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
      for (int c = 0; c < 64; ++c) acc = fmaf(A[r][c], xin[c][px], acc);
      acc += a0[r];
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

// pass 3 for the row layout: one pixel per thread, the whole row requested up front
// (16 x 16-byte loads in flight per thread), the 3x64 map broadcast from shared memory; same
// fmaf order as apply_kernel, so both give identical bits.
__global__ void __launch_bounds__(256)
apply_rows_kernel(const float* __restrict__ g, long long n, long long pix_stride, const float* __restrict__ map,
                  float* __restrict__ rgb) {
  __shared__ float A[3][64], a0[3];
  const int tid = threadIdx.x;
  if (tid < 192) A[tid >> 6][tid & 63] = map[kMapA + tid];
  if (tid < 3) a0[tid] = map[kMapA0 + tid];
  __syncthreads();
  const long long p = blockIdx.x * 256LL + tid;
  if (p >= n) return;
  float4 x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = __ldg(reinterpret_cast<const float4*>(g + p * pix_stride) + i);
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      acc = fmaf(A[r][4 * i], x[i].x, acc);
      acc = fmaf(A[r][4 * i + 1], x[i].y, acc);
      acc = fmaf(A[r][4 * i + 2], x[i].z, acc);
      acc = fmaf(A[r][4 * i + 3], x[i].w, acc);
    }
    acc += a0[r];
    rgb[(long long)r * n + p] = 1.f / (1.f + expf(-acc));
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
//   [P1, P1+P2)        gram partials   kMaxBlocks*1024
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

size_t style_scratch_floats(int64_t) { return kScratchFloats; }

static int check_feat(const float* p, int64_t n, int64_t ps, int64_t cs, const char* what) {
  CRNERF_REQUIRE(p != nullptr, "%s is null", what);
  CRNERF_REQUIRE(n >= 1, "%s has no pixels", what);
  CRNERF_REQUIRE(ps >= 1 && cs >= 1, "%s strides must be positive", what);
  return CRNERF_OK;
}

int style_stats1(const float* content, int64_t n, int64_t ps, int64_t cs, float* sums,
                 float* partial, cudaStream_t st) {
  int nb = blocks_for_pixels(n);
  if (rows_fast(content, ps, cs)) {
    nb = (int)std::max<long long>(1, std::min<long long>((n + 127) / 128, kMaxBlocks));
    sums_rows_kernel<<<nb, 256, 0, st>>>(content, n, ps, partial);
  } else {
    sums_kernel<<<nb, 256, 0, st>>>(content, n, ps, cs, partial);
  }
  reduce_partials_warp_kernel<<<8, 256, 0, st>>>(partial, nb, 64, 1.f, sums);
  count_launch(2);
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

int style_stats2(const crnerf_cnn_weights& cw, const float* content, int64_t n, int64_t ps,
                 int64_t cs, const float* mean, float* gram, float* partial, float scale,
                 cudaStream_t st) {
  // pixel MLP + Gram on the tensor core (gram_tc.cu), then the fixed-order sum of the per-block partials
  int nb = 0;
  int rc = gram_tc(cw, content, n, ps, cs, mean, partial, kMaxBlocks, &nb, st);
  if (rc) return rc;
  reduce_partials_warp_kernel<<<128, 256, 0, st>>>(partial, nb, 1024, scale, gram);
  count_launch(1);
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

// everything after the content Gram is known: style branch, FCs, compose, apply
int style_finish(const crnerf_style_weights* w, const float* content, int64_t n, int64_t ps,
                 int64_t cs, const float* mean_c, const float* gram_c_normalised,
                 const float* style, int64_t ns, int64_t sps, int64_t scs, float* rgb,
                 float* transmatrix, float* fused, float* scratch, cudaStream_t st) {
  float* sum_part = scratch + kOffSumPart;
  float* gram_part = scratch + kOffGramPart;
  float* mean_s = scratch + kOffMeanS;
  float* gram_s = scratch + kOffGramS;
  float* cmat = scratch + kOffCmat;
  float* smat = scratch + kOffSmat;
  float* map = scratch + kOffMap;
  int rc = style_stats1(style, ns, sps, scs, mean_s, sum_part, st);
  if (rc) return rc;
  reduce_partials_kernel<<<1, 64, 0, st>>>(mean_s, 1, 64, 1.f / (float)ns, mean_s);
  count_launch();
  rc = style_stats2(w->snet, style, ns, sps, scs, mean_s, gram_s, gram_part, 1.f / (float)ns, st);
  if (rc) return rc;
  fc_kernel<<<256, 256, 0, st>>>(w->cnet.fc_w, w->cnet.fc_b, gram_c_normalised, w->snet.fc_w,
                                 w->snet.fc_b, gram_s, cmat, smat);
  compose_kernel<<<1, 256, 0, st>>>(cmat, smat, *w, mean_c, mean_s, map, transmatrix);
  if (fused == nullptr && rows_fast(content, ps, cs))
    apply_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(content, n, ps, map, rgb);
  else
    apply_kernel<<<blocks_for_pixels(n), 256, 0, st>>>(content, n, ps, cs, map, rgb, fused);
  count_launch(3);
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
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

int style_forward(const crnerf_style_weights* w, const float* content, int64_t n, int64_t ps,
                  int64_t cs, const float* style, int64_t ns, int64_t sps, int64_t scs, float* rgb,
                  float* transmatrix, float* fused, float* scratch, cudaStream_t st) {
  CRNERF_REQUIRE(w && rgb && scratch, "null argument");
  int rc = check_feat(content, n, ps, cs, "content");
  if (rc) return rc;
  float* map = scratch + kOffMap;
  if (style == nullptr) {
    content_map_kernel<<<1, 192, 0, st>>>(*w, map);
    if (rows_fast(content, ps, cs))
      apply_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(content, n, ps, map, rgb);
    else
      apply_kernel<<<blocks_for_pixels(n), 256, 0, st>>>(content, n, ps, cs, map, rgb, nullptr);
    count_launch(2);
    CRNERF_CUDA(cudaGetLastError());
    return CRNERF_OK;
  }
  rc = check_feat(style, ns, sps, scs, "style");
  if (rc) return rc;
  float* mean_c = scratch + kOffMeanC;
  float* gram_c = scratch + kOffGramC;
  rc = style_stats1(content, n, ps, cs, mean_c, scratch + kOffSumPart, st);
  if (rc) return rc;
  reduce_partials_kernel<<<1, 64, 0, st>>>(mean_c, 1, 64, 1.f / (float)n, mean_c);
  count_launch();
  rc = style_stats2(w->cnet, content, n, ps, cs, mean_c, gram_c, scratch + kOffGramPart,
                    1.f / (float)n, st);
  if (rc) return rc;
  return style_finish(w, content, n, ps, cs, mean_c, gram_c, style, ns, sps, scs, rgb,
                      transmatrix, fused, scratch, st);
}

}  // namespace crnerf

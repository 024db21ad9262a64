// Backward of the cross-ray fusion + decoder (style_net.forward under autograd: the decode() of
// the training step, reference train_mask_grid_sample.py:127-149 with
// models/linearStyleTransfer.py:28-37, 58-90, 284-291 and models/nerf_decoder_stylenerf.py:279-291).
//
// The training step decodes 32x32 patches (1,024 pixels): everything here is a few MFLOP, so the
// kernels are fp32 CUDA-core code organised for few launches and deterministic sums, not for the
// tensor core.  Nothing but the two FC outputs (C, S), T = S C and the channel means is kept from
// the forward (crnerf_style_forward_train's `aux`); per-pixel activations are recomputed.
//
// Per pixel the block is two dense chains over the centred features xc = x - mean:
//   head : xc -Wc,bc-> comp(32) -T-> z(32) -Wu,bu+mu_s-> u(64) -Wr,br-> sigmoid -> rgb(3)
//   cnn  : xc -W1,b1-> lrelu(128) -W2,b2-> lrelu(64) -W3,b3-> y(32);  G = sum_p y y^T / n
// chain_backward_kernel runs either one on 8-pixel tiles: forward recompute into shared memory,
// then top-down input gradients and outer-product weight gradients, accumulated into the CTA's own
// slot of a partial buffer (plain read-modify-write, no atomics; a fixed-order reduction follows).
// Between the two chains sit the 32x32 algebra (dS = dT C^T, dC = S^T dT) and the two FC layers
// (d vec(G) = F^T d c, dF = d c (x) vec(G): style_fc_backward_kernel).  The mean subtraction's
// backward (dx = dxc - mean_p dxc) and the style mean's gradient (mu_s enters the unzip bias) are
// applied by style_backward_finish_kernel.
#include <algorithm>
#include "common.h"

namespace crnerf {
namespace {

constexpr int kTP = 8;         // pixels per tile: a 32x32 training patch spreads over 128 CTAs per map (the kernel is
                               // latency-bound per tile: 16-pixel tiles on 64 CTAs took 1.8x as long)
constexpr int kActStride = 297;  // floats per pixel of the activation stash (>= 64+128+64+32, odd: no bank conflicts)
constexpr int kGStride = 129;
constexpr int kMaxLayers = 4;

struct ChainLayer {
  const float* W;    // (N, K) row-major
  const float* b;    // (N) or nullptr
  const float* b2;   // optional second vector added to the bias (the style mean for `unzip`)
  int K, N;
  int act;           // 0 none, 1 LeakyReLU(0.2), 2 sigmoid
  int w_off, b_off;  // offsets of dW / db inside a partial slot (b_off < 0: no bias)
};

struct ChainJob {
  ChainLayer L[kMaxLayers];
  int n_layers;
  const float* x;
  long long n, ps, cs;
  const float* mean;       // (64)
  int top_mode;            // 0: g_top (N_last, n) planar; 1: g_y = top_scale * (gg + gg^T) y
  const float* g_top;
  const float* gg;         // (N_last, N_last): d vec(G) before symmetrisation
  float top_scale;
  float* gx;               // out (n, 64) rows: gradient w.r.t. xc
  float* partial;          // (n_blocks, slot_len)
  int slot_len, colsum_off;
  int first_block, n_blocks;
};
struct ChainParams {
  ChainJob job[2];
  int n_jobs;
};

__device__ __forceinline__ float act_fn(float v, int act) {
  if (act == 1) return v > 0.f ? v : 0.2f * v;
  if (act == 2) return 1.f / (1.f + expf(-v));
  return v;
}
// derivative expressed through the OUTPUT of the activation (LeakyReLU keeps the sign)
__device__ __forceinline__ float act_grad(float out, int act) {
  if (act == 1) return out > 0.f ? 1.f : 0.2f;
  if (act == 2) return out * (1.f - out);
  return 1.f;
}

__global__ void __launch_bounds__(256)
chain_backward_kernel(const __grid_constant__ ChainParams P) {
  __shared__ float acts[kTP * kActStride];
  __shared__ float gA[kTP * kGStride], gB[kTP * kGStride];
  __shared__ float gsym[32 * 33];
  const ChainJob& J = P.job[(P.n_jobs == 2 && (int)blockIdx.x >= P.job[1].first_block) ? 1 : 0];
  const int lb = (int)blockIdx.x - J.first_block, tid = threadIdx.x;
  float* slot = J.partial + (long long)lb * J.slot_len;
  const int nl = J.n_layers, n_last = J.L[nl - 1].N;
  if (J.top_mode == 1) {
    for (int e = tid; e < n_last * n_last; e += 256) {
      const int j = e / n_last, i = e % n_last;
      gsym[j * 33 + i] = (J.gg[j * n_last + i] + J.gg[i * n_last + j]) * J.top_scale;
    }
  }
  int off[kMaxLayers + 1];
  off[0] = 0;
  for (int l = 0; l < nl; ++l) off[l + 1] = off[l] + J.L[l].K;   // input of layer l at off[l]; final output at off[nl]
  // The chain's weights, staged once: a CTA touches every weight only twice per tile (forward, input gradient), so
  // from global memory each of those reads was a cold L2 round trip inside a dependent FMA chain - the kernel ran at
  // L2 latency.  Rows padded to K + 1 floats: the forward's four rows per warp land in different banks.
  extern __shared__ float wsm[];
  int woff[kMaxLayers];
  {
    int o = 0;
    for (int l = 0; l < nl; ++l) {
      woff[l] = o;
      const ChainLayer& L = J.L[l];
      for (int e = tid; e < L.N * L.K; e += 256) wsm[o + (e / L.K) * (L.K + 1) + e % L.K] = __ldg(L.W + e);
      o += L.N * (L.K + 1);
    }
  }
  const long long n_tiles = (J.n + kTP - 1) / kTP;
  bool first = true;
  for (long long t = lb; t < n_tiles; t += J.n_blocks, first = false) {
    const long long p0 = t * kTP;
    __syncthreads();
    // ---- centred input
    for (int e = tid; e < kTP * 64; e += 256) {
      const int p = e >> 6, c = e & 63;
      const long long px = p0 + p;
      acts[p * kActStride + c] = px < J.n ? J.x[px * J.ps + (long long)c * J.cs] - J.mean[c] : 0.f;
    }
    __syncthreads();
    // ---- forward recompute
    for (int l = 0; l < nl; ++l) {
      const ChainLayer& L = J.L[l];
      for (int e = tid; e < kTP * L.N; e += 256) {
        const int p = e % kTP, j = e / kTP;
        float acc = L.b ? L.b[j] : 0.f;
        if (L.b2) acc += L.b2[j];
        const float* w = wsm + woff[l] + j * (L.K + 1);
        const float* in = acts + p * kActStride + off[l];
#pragma unroll 8
        for (int k = 0; k < L.K; ++k) acc = fmaf(w[k], in[k], acc);
        acts[p * kActStride + off[l + 1] + j] = act_fn(acc, L.act);
      }
      __syncthreads();
    }
    // ---- top gradient (rows beyond n contribute nothing)
    float* g_out = gA;
    float* g_in = gB;
    for (int e = tid; e < kTP * n_last; e += 256) {
      const int p = e % kTP, j = e / kTP;
      const long long px = p0 + p;
      float g = 0.f;
      if (px < J.n) {
        if (J.top_mode == 0) {
          g = J.g_top[(long long)j * J.n + px];
        } else {
          const float* y = acts + p * kActStride + off[nl];
          for (int i = 0; i < n_last; ++i) g = fmaf(gsym[j * 33 + i], y[i], g);
        }
      }
      g_out[p * kGStride + j] = g;
    }
    __syncthreads();
    // ---- top-down
    for (int l = nl - 1; l >= 0; --l) {
      const ChainLayer& L = J.L[l];
      for (int e = tid; e < kTP * L.N; e += 256) {   // through the activation
        const int p = e % kTP, j = e / kTP;
        g_out[p * kGStride + j] *= act_grad(acts[p * kActStride + off[l + 1] + j], L.act);
      }
      __syncthreads();
      if (L.b_off >= 0) {
        for (int j = tid; j < L.N; j += 256) {
          float acc = 0.f;
          for (int p = 0; p < kTP; ++p) acc += g_out[p * kGStride + j];
          slot[L.b_off + j] = first ? acc : slot[L.b_off + j] + acc;
        }
      }
      for (int e = tid; e < L.N * L.K; e += 256) {   // dW[j][k] = sum_p g[p][j] in[p][k]
        const int j = e / L.K, k = e % L.K;
        float acc = 0.f;
#pragma unroll
        for (int p = 0; p < kTP; ++p) acc = fmaf(g_out[p * kGStride + j], acts[p * kActStride + off[l] + k], acc);
        slot[L.w_off + e] = first ? acc : slot[L.w_off + e] + acc;
      }
      for (int e = tid; e < kTP * L.K; e += 256) {   // g_in[p][k] = sum_j W[j][k] g[p][j]
        const int p = e % kTP, k = e / kTP;
        float acc = 0.f;
        const float* w = wsm + woff[l] + k;
#pragma unroll 8
        for (int j = 0; j < L.N; ++j) acc = fmaf(w[j * (L.K + 1)], g_out[p * kGStride + j], acc);
        g_in[p * kGStride + k] = acc;
      }
      __syncthreads();
      float* tmp = g_out;
      g_out = g_in;
      g_in = tmp;
    }
    // ---- gradient w.r.t. the centred input (now in g_out) + its column sums
    for (int e = tid; e < kTP * 64; e += 256) {
      const int p = e >> 6, c = e & 63;
      if (p0 + p < J.n) J.gx[(p0 + p) * 64 + c] = g_out[p * kGStride + c];
    }
    if (tid < 64) {
      float acc = 0.f;
      for (int p = 0; p < kTP; ++p) acc += g_out[p * kGStride + tid];   // rows beyond n are zero
      slot[J.colsum_off + tid] = first ? acc : slot[J.colsum_off + tid] + acc;
    }
  }
}

// out[i] = sum_b partial[b][i], fixed order: one thread per element (coalesced across the block), eight slots in
// flight per thread, partial sums combined in a fixed tree
__global__ void __launch_bounds__(128)
reduce_slots_kernel(const float* __restrict__ partial, int n_slots, int len, float* __restrict__ out) {
  const int i = blockIdx.x * 128 + threadIdx.x;
  if (i >= len) return;
  const float* p = partial + i;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int b = 0;
  for (; b + 7 < n_slots; b += 8) {
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] += __ldcs(p + (long long)(b + u) * len);
  }
  for (; b < n_slots; ++b) acc[0] += __ldcs(p + (long long)b * len);
  out[i] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
}

// dS = dT C^T, dC = S^T dT  (T = S C, MulLayer.forward :86); also the fc bias gradients (= dc, ds)
__global__ void __launch_bounds__(256)
style_mid_kernel(const float* __restrict__ dT, const float* __restrict__ C, const float* __restrict__ S,
                 float* __restrict__ dC, float* __restrict__ dS, float* __restrict__ dbc, float* __restrict__ dbs) {
  __shared__ float sT[32][33], sC[32][33], sS[32][33];
  const int tid = threadIdx.x;
  for (int e = tid; e < 1024; e += 256) {
    sT[e >> 5][e & 31] = dT[e];
    sC[e >> 5][e & 31] = C[e];
    sS[e >> 5][e & 31] = S[e];
  }
  __syncthreads();
  for (int e = tid; e < 1024; e += 256) {
    const int a = e >> 5, b = e & 31;
    float gs = 0.f, gc = 0.f;
    for (int j = 0; j < 32; ++j) {
      gs = fmaf(sT[a][j], sC[b][j], gs);   // dS[a][b] = sum_j dT[a][j] C[b][j]
      gc = fmaf(sS[j][a], sT[j][b], gc);   // dC[a][b] = sum_i S[i][a] dT[i][b]
    }
    dS[e] = gs;
    dC[e] = gc;
    dbs[e] = gs;
    dbc[e] = gc;
  }
}

// Both FC layers (c = F vec(G) + fb): block (chunk, net) takes 32 rows r of F_net:
//   dF[r][k] = dc[r] * G[k]                      (written in full)
//   partial[chunk][net*1024 + k] = sum_{r in chunk} F[r][k] dc[r]     (d vec(G), reduced afterwards)
__global__ void __launch_bounds__(256)
style_fc_backward_kernel(const float* __restrict__ Fc, const float* __restrict__ Fs, const float* __restrict__ dc,
                         const float* __restrict__ ds, const float* __restrict__ Gc, const float* __restrict__ Gs,
                         float* __restrict__ dFc, float* __restrict__ dFs, float* __restrict__ partial) {
  const int chunk = blockIdx.x, net = blockIdx.y, tid = threadIdx.x;
  const float* F = net ? Fs : Fc;
  const float* d = net ? ds : dc;
  const float* G = net ? Gs : Gc;
  float* dF = net ? dFs : dFc;
  __shared__ float sd[32];
  if (tid < 32) sd[tid] = d[chunk * 32 + tid];
  __syncthreads();
  float g[4], acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int u = 0; u < 4; ++u) g[u] = G[tid + 256 * u];
  for (int r = 0; r < 32; ++r) {
    const long long row = (long long)(chunk * 32 + r) * 1024;
    const float dr = sd[r];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = tid + 256 * u;
      acc[u] = fmaf(F[row + k], dr, acc[u]);
      dF[row + k] = dr * g[u];
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) partial[(long long)chunk * 2048 + net * 1024 + tid + 256 * u] = acc[u];
}

// dx[p][c] = (gxa[p][c] + gxb[p][c]) - (suma[c] + sumb[c]) / n + extra[c] * extra_scale   (gxa / suma / extra may be null)
__global__ void __launch_bounds__(256)
style_backward_finish_kernel(const float* __restrict__ gxa, const float* __restrict__ suma, const float* __restrict__ gxb,
                             const float* __restrict__ sumb, const float* __restrict__ extra, float extra_scale,
                             long long n, float* __restrict__ out) {
  const long long total = n * 64;
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
    const int c = (int)(e & 63);
    float v = gxb[e], s = sumb[c];
    if (gxa) {
      v += gxa[e];
      s += suma[c];
    }
    v -= s / (float)n;
    if (extra) v = fmaf(extra[c], extra_scale, v);
    out[e] = v;
  }
}

// slot layouts (floats)
constexpr int kA_dWc = 0, kA_dbc = 2048, kA_dT = 2080, kA_dWu = 3104, kA_dbu = 5152, kA_dWr = 5216, kA_dbr = 5408,
              kA_colsum = 5412, kA_len = 5476;
constexpr int kB_dW1 = 0, kB_db1 = 8192, kB_dW2 = 8320, kB_db2 = 16512, kB_dW3 = 16576, kB_db3 = 18624,
              kB_colsum = 18656, kB_len = 18720;
constexpr int kMaxSlots = 128;

}  // namespace

// aux (floats) written by the training forward: the contiguous tail of the style scratch
// [mean_c 64 | mean_s 64 | gram_c 1024 | gram_s 1024 | cmat 1024 | smat 1024] followed by T (1024)
constexpr int kAuxMeanC = 0, kAuxMeanS = 64, kAuxGramC = 128, kAuxGramS = 1152, kAuxC = 2176, kAuxS = 3200, kAuxT = 4224,
              kAuxFloats = 5248;
size_t style_aux_floats() { return kAuxFloats; }

// flat gradient buffer (floats): head block | cnet block | snet block | dFc | dfbc | dFs | dfbs
constexpr size_t kG_head = 0, kG_cnet = kA_len, kG_snet = kG_cnet + kB_len, kG_dFc = kG_snet + kB_len,
                 kG_dfbc = kG_dFc + 1048576, kG_dFs = kG_dfbc + 1024, kG_dfbs = kG_dFs + 1048576,
                 kG_floats = kG_dfbs + 1024;
size_t style_backward_grads_floats() { return kG_floats; }

// offsets of the 22 parameter gradients inside the flat buffer, in crnerf_style_weights order:
// cnet {conv_w[3], conv_b[3], fc_w, fc_b}, snet {same}, compress_w, compress_b, unzip_w, unzip_b, rgb_w, rgb_b;
// then the two extras: dT (transmatrix) and the column sums are internal.
void style_backward_layout(int64_t* out22) {
  const int64_t net[8] = {kB_dW1, kB_dW2, kB_dW3, kB_db1, kB_db2, kB_db3, 0, 0};
  for (int i = 0; i < 6; ++i) {
    out22[i] = (int64_t)kG_cnet + net[i];
    out22[8 + i] = (int64_t)kG_snet + net[i];
  }
  out22[6] = kG_dFc;
  out22[7] = kG_dfbc;
  out22[14] = kG_dFs;
  out22[15] = kG_dfbs;
  out22[16] = kG_head + kA_dWc;
  out22[17] = kG_head + kA_dbc;
  out22[18] = kG_head + kA_dWu;
  out22[19] = kG_head + kA_dbu;
  out22[20] = kG_head + kA_dWr;
  out22[21] = kG_head + kA_dbr;
}

// scratch (floats): partial A | partial B (2 nets) | fc partial (32 x 2048) | dvecG (2048) | dC, dS (2 x 1024) |
//                   gxa (n x 64) | gxb_c (n x 64) | gxb_s (m x 64)
static size_t bwd_off_partB() { return (size_t)kMaxSlots * kA_len; }
static size_t bwd_off_fcpart() { return bwd_off_partB() + (size_t)2 * kMaxSlots * kB_len; }
static size_t bwd_off_dvecg() { return bwd_off_fcpart() + 32 * 2048; }
static size_t bwd_off_dcs() { return bwd_off_dvecg() + 2048; }
static size_t bwd_off_gx() { return bwd_off_dcs() + 2048; }
size_t style_backward_scratch_floats(int64_t n, int64_t m) {
  return bwd_off_gx() + (size_t)(2 * n + m) * 64;
}

// dynamic shared memory of chain_backward_kernel: the job's weights with rows padded by one float
static size_t chain_wsm_bytes(const ChainJob& j) {
  size_t f = 0;
  for (int l = 0; l < j.n_layers; ++l) f += (size_t)j.L[l].N * (j.L[l].K + 1);
  return f * sizeof(float);
}

int style_backward(const crnerf_style_weights* w, const float* content, int64_t n, int64_t ps, int64_t cs,
                   const float* style, int64_t m, int64_t sps, int64_t scs, const float* aux, const float* g_rgb,
                   float* g_content, float* g_style, float* grads, float* scratch, cudaStream_t st) {
  CRNERF_REQUIRE(w && content && style && aux && g_rgb && g_content && g_style && grads && scratch, "null argument");
  CRNERF_REQUIRE(n >= 1 && m >= 1, "empty map");
  // 22 KB static + up to 75 KB of staged weights: two CTAs per SM
  CRNERF_CUDA(cudaFuncSetAttribute(chain_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
  float* partA = scratch;
  float* partB = scratch + bwd_off_partB();
  float* fcpart = scratch + bwd_off_fcpart();
  float* dvecg = scratch + bwd_off_dvecg();
  float* dC = scratch + bwd_off_dcs();
  float* dS = dC + 1024;
  float* gxa = scratch + bwd_off_gx();
  float* gxb_c = gxa + (size_t)n * 64;
  float* gxb_s = gxb_c + (size_t)n * 64;
  auto blocks_for = [](int64_t px) { return (int)std::max<int64_t>(1, std::min<int64_t>((px + kTP - 1) / kTP, kMaxSlots)); };

  // ---- head chain on the content map
  ChainParams A{};
  A.n_jobs = 1;
  ChainJob& ja = A.job[0];
  ja.n_layers = 4;
  ja.L[0] = {w->compress_w, w->compress_b, nullptr, 64, 32, 0, kA_dWc, kA_dbc};
  ja.L[1] = {aux + kAuxT, nullptr, nullptr, 32, 32, 0, kA_dT, -1};
  ja.L[2] = {w->unzip_w, w->unzip_b, aux + kAuxMeanS, 32, 64, 0, kA_dWu, kA_dbu};
  ja.L[3] = {w->rgb_w, w->rgb_b, nullptr, 64, 3, 2, kA_dWr, kA_dbr};
  ja.x = content;
  ja.n = n;
  ja.ps = ps;
  ja.cs = cs;
  ja.mean = aux + kAuxMeanC;
  ja.top_mode = 0;
  ja.g_top = g_rgb;
  ja.gx = gxa;
  ja.partial = partA;
  ja.slot_len = kA_len;
  ja.colsum_off = kA_colsum;
  ja.first_block = 0;
  ja.n_blocks = blocks_for(n);
  chain_backward_kernel<<<ja.n_blocks, 256, chain_wsm_bytes(ja), st>>>(A);
  reduce_slots_kernel<<<(kA_len + 127) / 128, 128, 0, st>>>(partA, ja.n_blocks, kA_len, grads + kG_head);
  // ---- 32x32 algebra and the FC layers
  float* head = grads + kG_head;
  style_mid_kernel<<<1, 256, 0, st>>>(head + kA_dT, aux + kAuxC, aux + kAuxS, dC, dS, grads + kG_dfbc, grads + kG_dfbs);
  style_fc_backward_kernel<<<dim3(32, 2), 256, 0, st>>>(w->cnet.fc_w, w->snet.fc_w, dC, dS, aux + kAuxGramC,
                                                        aux + kAuxGramS, grads + kG_dFc, grads + kG_dFs, fcpart);
  reduce_slots_kernel<<<2048 / 128, 128, 0, st>>>(fcpart, 32, 2048, dvecg);
  // ---- pixel-MLP chains: content with cnet, style with snet, one launch
  ChainParams B{};
  B.n_jobs = 2;
  for (int net = 0; net < 2; ++net) {
    const crnerf_cnn_weights& cw = net ? w->snet : w->cnet;
    ChainJob& jb = B.job[net];
    jb.n_layers = 3;
    jb.L[0] = {cw.conv_w[0], cw.conv_b[0], nullptr, 64, 128, 1, kB_dW1, kB_db1};
    jb.L[1] = {cw.conv_w[1], cw.conv_b[1], nullptr, 128, 64, 1, kB_dW2, kB_db2};
    jb.L[2] = {cw.conv_w[2], cw.conv_b[2], nullptr, 64, 32, 0, kB_dW3, kB_db3};
    jb.x = net ? style : content;
    jb.n = net ? m : n;
    jb.ps = net ? sps : ps;
    jb.cs = net ? scs : cs;
    jb.mean = aux + (net ? kAuxMeanS : kAuxMeanC);
    jb.top_mode = 1;
    jb.gg = dvecg + net * 1024;
    jb.top_scale = 1.f / (float)(net ? m : n);     // G = sum y y^T / n
    jb.gx = net ? gxb_s : gxb_c;
    jb.partial = partB + (size_t)net * kMaxSlots * kB_len;
    jb.slot_len = kB_len;
    jb.colsum_off = kB_colsum;
    jb.n_blocks = blocks_for(jb.n);
    jb.first_block = net ? B.job[0].n_blocks : 0;
  }
  chain_backward_kernel<<<B.job[0].n_blocks + B.job[1].n_blocks, 256, chain_wsm_bytes(B.job[0]), st>>>(B);
  reduce_slots_kernel<<<(kB_len + 127) / 128, 128, 0, st>>>(B.job[0].partial, B.job[0].n_blocks, kB_len, grads + kG_cnet);
  reduce_slots_kernel<<<(kB_len + 127) / 128, 128, 0, st>>>(B.job[1].partial, B.job[1].n_blocks, kB_len, grads + kG_snet);
  // ---- mean-subtraction backward; the style mean also enters through the unzip bias (d mu_s = d bu)
  const int gc = (int)std::max<int64_t>(1, std::min<int64_t>((n * 64 + 255) / 256, 4 * num_sms()));
  const int gs = (int)std::max<int64_t>(1, std::min<int64_t>((m * 64 + 255) / 256, 4 * num_sms()));
  style_backward_finish_kernel<<<gc, 256, 0, st>>>(gxa, head + kA_colsum, gxb_c, grads + kG_cnet + kB_colsum, nullptr,
                                                   0.f, n, g_content);
  style_backward_finish_kernel<<<gs, 256, 0, st>>>(nullptr, nullptr, gxb_s, grads + kG_snet + kB_colsum,
                                                   head + kA_dbu, 1.f / (float)m, m, g_style);
  count_launch(10);
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

}  // namespace crnerf

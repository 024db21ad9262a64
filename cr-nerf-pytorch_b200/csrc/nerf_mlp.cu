// Fused NeRF_sigma volume-rendering pass for sm_100a.
//
// Replaces, in one persistent kernel, the reference's `inference` closure
// (models/rendering.py:82-145): positional encoding (models/nerf.py:17-30),
// the 11-layer NeRF_sigma MLP (models/nerf.py:157-182) and the alpha composite
// (rendering.py:121-143).  No (points x width) intermediate touches HBM.
//
// Structure (one CTA per SM, 384 threads):
//   warp 0      producer : streams the packed weight image L2 -> smem ring with
//                          cp.async.bulk (TMA engine), mbarrier complete_tx
//   warp 1      issuer   : one thread issues tcgen05.mma (kind::f16, M=128,
//                          N=128/64, K=16); activations are the A operand read
//                          from TMEM (TS form), the embedding is read from smem
//   warp 2      TMEM allocator
//   warps 4-7   epilogue group X, warps 8-11 epilogue group Y: each group owns
//               one 128-point tile (one point per thread = one TMEM lane):
//               embedding -> smem, per layer tcgen05.ld -> +bias -> ReLU ->
//               cvt.f16x2 -> tcgen05.st back as next layer's A operand; sigma
//               head as an fp32 dot in the layer-8 epilogue; segmented
//               warp-shuffle transmittance scan; weighted feature reduction.
// Two tiles (X, Y) are in flight per CTA and share every weight chunk; the
// issuer alternates X/Y per 128-wide output half so each epilogue runs under
// the other tile's MMAs.  TMEM: X {A: cols 0-127, D: 128-255}, Y {A: 256-383,
// D: 384-511}; the first output half is held in registers until the layer's
// second half has been issued, so A needs no double buffer.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <algorithm>
#include <cstring>
#include <mutex>
#include "common.h"
#include "nerf_layout.h"
#include "ptx.cuh"

namespace crnerf {

__constant__ Chunk c_chunks[kMaxChunks];
__constant__ Unit c_units[kMaxUnits];

namespace {

constexpr int kSlots = 8;
constexpr int kSlotBytes = 16384;
constexpr int kEmbBufBytes = 32768;
constexpr int kThreads = 384;
constexpr int kMaxSeg = 10;  // ray segments per 128-row tile (n_samples >= 16)

constexpr int kRingOff = 0;
constexpr int kEmbOff = kSlots * kSlotBytes;        // 131072
constexpr int kBlobOff = kEmbOff + 2 * kEmbBufBytes;  // 196608
constexpr int kMiscOff = kBlobOff + kBlobFloats * 4;  // 207648

struct Misc {
  uint64_t ring_full[kSlots];
  uint64_t ring_empty[kSlots];
  uint64_t emb_full[2], a_full[2], d_full[2], d_empty[2], carry_a[2], carry_b[2];
  uint32_t tmem_base;
  uint32_t pad0;
  float scan_p[2][4];
  float scan_d[2][4];
  int scan_f[2][4];
  float carry_T[2];
  float carry_depth[2];
  float carry_feat[2][64];
  float part[2][2][kMaxSeg][64];
};
constexpr int kSmemBytes = kMiscOff + (int)sizeof(Misc);
static_assert(kSmemBytes <= 232448, "exceeds 227 KB of shared memory");
static_assert(kMiscOff % 16 == 0, "misc alignment");

enum : int { kModeEmbedded = 1, kModeRaw = 2, kModeSigmaOnly = 4 };

struct RenderParams {
  const uint8_t* wimg;
  const float* blob;
  const float* rays;
  const float* view_dir;
  const float* z_vals;
  const float* noise;
  const float* x;
  float* weights;
  float* feature;
  float* depth;
  float* raw;
  float* dbg;
  long long n_points;
  long long pts_per_cta;
  int x_stride;
  int dbg_layer;
  int S;
  int n_freq_xyz, n_freq_dir, e_xyz, e_dir;
  int n_chunks, n_units;
  int mode;
};

__device__ __forceinline__ void group_sync(int b) {
  asm volatile("bar.sync %0, 128;" ::"r"(1 + b) : "memory");
}

template <int kFmt>
__device__ __forceinline__ uint16_t to_operand(float v) {
  if constexpr (kFmt == 0) {
    // clamp to the finite fp16 range so an outlier cannot turn into inf
    v = fminf(fmaxf(v, -65504.f), 65504.f);
    return __half_as_ushort(__float2half_rn(v));
  } else {
    return __bfloat16_as_ushort(__float2bfloat16_rn(v));
  }
}

// element (row, col) of a tile's embedding buffer: two K-major SW128 slabs
template <int kFmt>
__device__ __forceinline__ void emb_put(uint8_t* buf, int row, int col, float v) {
  const uint32_t off = (uint32_t)(col >> 6) * 16384u + sw128_offset(row, (col & 63) >> 3) +
                       (uint32_t)(col & 7) * 2u;
  *reinterpret_cast<uint16_t*>(buf + off) = to_operand<kFmt>(v);
}

// sin/cos of 2^k * x with x/(2*pi) supplied as an unevaluated sum hi+lo.
// The argument 2^k*x is exact in fp32 (power-of-two scale), so the only error
// is in the range reduction - done here to ~2^-45 of a revolution in
// double-float arithmetic - and the SFU evaluation on |a| <= pi (~5e-7 abs),
// far below the 16-bit operand rounding (2^-11) applied right after.
__device__ __forceinline__ void sincos_band(float x, float hi, float lo, int k, float& s, float& c) {
  const float sc = __int_as_float((127 + k) << 23);
  const float hk = hi * sc, lk = lo * sc;
  if (fabsf(hk) < 4194304.f) {
    const float n = rintf(hk);
    const float r = (hk - n) + lk;
    const float a = r * 6.283185307179586f;
    s = __sinf(a);
    c = __cosf(a);
  } else {  // |x| beyond any scene scale: full-range library path
    sincosf(x * sc, &s, &c);
  }
}

// [v, sin(2^0 v), cos(2^0 v), ...] for a 3-vector, written at buffer column col0
// (reference column order, models/nerf.py:25-30).
template <int kFmt>
__device__ __forceinline__ void embed3(uint8_t* buf, int row, int col0, const float (&v)[3],
                                       int n_freqs) {
  float hi[3], lo[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    emb_put<kFmt>(buf, row, col0 + i, v[i]);
    const float chi = 0.15915493667125702f;  // fl32(1/2pi)
    const float clo = 6.420638316725915e-09f;  // 1/2pi - chi
    hi[i] = v[i] * chi;
    lo[i] = fmaf(v[i], clo, fmaf(v[i], chi, -hi[i]));
  }
  for (int k = 0; k < n_freqs; ++k) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      float s, c;
      sincos_band(v[i], hi[i], lo[i], k, s, c);
      emb_put<kFmt>(buf, row, col0 + 3 + 6 * k + i, s);
      emb_put<kFmt>(buf, row, col0 + 6 + 6 * k + i, c);
    }
  }
}

__device__ __forceinline__ float softplus_ref(float x) {
  // torch.nn.Softplus(beta=1, threshold=20) (models/nerf.py:149)
  return x > 20.f ? x : log1pf(expf(x));
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
      "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]),
      "r"(v[15])
      : "memory");
}

// One 32-column slice of an accumulator: +bias, optional ReLU, pack to 16 words.
// kSigma additionally accumulates the fp32 sigma-head dot product.
template <int kFmt, bool kRelu, bool kSigma>
__device__ __forceinline__ void bias_act_pack(const uint32_t (&v)[32], const float* bias,
                                              const float* wsig, uint32_t* out, float& sig_acc) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float2 bb = *reinterpret_cast<const float2*>(bias + 2 * j);
    const float2 a =
        add2(make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), bb);
    if constexpr (kSigma) {
      const float2 ws = *reinterpret_cast<const float2*>(wsig + 2 * j);
      sig_acc = fmaf(fmaxf(a.x, 0.f), ws.x, sig_acc);
      sig_acc = fmaf(fmaxf(a.y, 0.f), ws.y, sig_acc);
    }
    out[j] = pack2<kFmt, kRelu>(a.x, a.y);
  }
}

template <int kFmt>
__global__ void __launch_bounds__(kThreads, 1)
render_fused_kernel(const __grid_constant__ RenderParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* ring = smem + kRingOff;
  uint8_t* emb = smem + kEmbOff;
  float* blob = reinterpret_cast<float*>(smem + kBlobOff);
  Misc* M = reinterpret_cast<Misc*>(smem + kMiscOff);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long p0 = (long long)blockIdx.x * P.pts_per_cta;
  const long long p1 = min(p0 + P.pts_per_cta, P.n_points);
  if (p0 >= p1) return;
  const int n_tiles = (int)((p1 - p0 + 127) >> 7);
  const int n_pairs = (n_tiles + 1) >> 1;

  if (tid == 0) {
    for (int i = 0; i < kSlots; ++i) {
      mbar_init(&M->ring_full[i], 1);
      mbar_init(&M->ring_empty[i], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&M->emb_full[b], 128);
      mbar_init(&M->a_full[b], 128);
      mbar_init(&M->d_full[b], 1);
      mbar_init(&M->d_empty[b], 128);
      mbar_init(&M->carry_a[b], 1);
      mbar_init(&M->carry_b[b], 64);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(&M->tmem_base);
  for (int i = tid; i < kBlobFloats; i += kThreads) blob[i] = P.blob[i];
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = M->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    if (elect_one()) {
      const uint64_t pol = l2_policy_evict_last();
      uint32_t g = 0;
      for (int pair = 0; pair < n_pairs; ++pair) {
        for (int c = 0; c < P.n_chunks; ++c, ++g) {
          const uint32_t slot = g % kSlots, n = g / kSlots;
          mbar_wait(&M->ring_empty[slot], (n & 1) ^ 1, 1);
          const uint32_t bytes = (uint32_t)c_chunks[c].bytes;
          mbar_arrive_expect_tx(&M->ring_full[slot], bytes);
          bulk_g2s_hint(ring + slot * kSlotBytes, P.wimg + c_chunks[c].offset, bytes,
                        &M->ring_full[slot], pol);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // -------------------------------------------------------------------- issuer
    if (elect_one()) {
      const uint32_t ring_addr = smem_u32(ring);
      const uint32_t emb_addr = smem_u32(emb);
      uint32_t ucount[2] = {0, 0}, acount[2] = {0, 0}, tcount[2] = {0, 0};
      uint32_t g_base = 0;
      for (int pair = 0; pair < n_pairs; ++pair) {
        const bool valid1 = (2 * pair + 1) < n_tiles;
        const int last_b = valid1 ? 1 : 0;
        for (int u = 0; u < P.n_units; ++u) {
          const Unit un = c_units[u];
          const uint32_t idesc = make_idesc_f16(128, (uint32_t)un.n, kFmt);
          for (int b = 0; b <= last_b; ++b) {
            const uint32_t tA = tmem + (b ? 256u : 0u);
            const uint32_t tD = tmem + (b ? 384u : 128u);
            if (un.first_of_layer) {
              if (un.layer == 0) {
                mbar_wait(&M->emb_full[b], tcount[b] & 1, 2);
              } else {
                mbar_wait(&M->a_full[b], acount[b] & 1, 3);
                acount[b]++;
              }
            }
            mbar_wait(&M->d_empty[b], (ucount[b] & 1) ^ 1, 4);
            tc_fence_after_sync();
            for (int j = 0; j < un.nchunks; ++j) {
              const uint32_t g = g_base + (uint32_t)(un.chunk0 + j);
              const uint32_t slot = g % kSlots, n = g / kSlots;
              const Chunk ch = c_chunks[un.chunk0 + j];
              if (b == 0) {
                mbar_wait(&M->ring_full[slot], n & 1, 5);
                tc_fence_after_sync();
              }
              const uint32_t bslot = ring_addr + slot * kSlotBytes;
              for (int k = 0; k < ch.nk; ++k) {
                const uint64_t bdesc = make_sdesc_k_sw128(bslot + (uint32_t)k * 32u, 1024);
                const uint32_t acc = (j | k) ? 1u : 0u;
                const int ak = ch.a_k0 + k;
                if (ch.a_src == kSrcEmb) {
                  const uint32_t a_smem = emb_addr + (uint32_t)b * kEmbBufBytes +
                                          (uint32_t)(ak >> 2) * 16384u + (uint32_t)(ak & 3) * 32u;
                  umma_ss(tD, make_sdesc_k_sw128(a_smem, 1024), bdesc, idesc, acc);
                } else {
                  umma_ts(tD, tA + (uint32_t)ak * 8u, bdesc, idesc, acc);
                }
              }
              if (b == last_b) umma_commit(&M->ring_empty[slot]);
            }
            umma_commit(&M->d_full[b]);
            ucount[b]++;
          }
        }
        tcount[0]++;
        tcount[1]++;
        g_base += (uint32_t)P.n_chunks;
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int b = (warp - 4) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int gtid = tid - 128 - b * 128;  // 0..127 inside the group (== row)
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t tA = tmem + (b ? 256u : 0u) + lane_off;
    const uint32_t tD = tmem + (b ? 384u : 128u) + lane_off;
    uint8_t* my_emb = emb + b * kEmbBufBytes;
    float* staging = reinterpret_cast<float*>(my_emb);
    const bool ray_mode = !(P.mode & kModeEmbedded);
    const bool raw_mode = (P.mode & kModeRaw) != 0;
    const float* wsig = blob + kSigmaWOff;
    uint32_t ud = 0;

    for (int pair = 0; pair < n_pairs; ++pair) {
      const int t = 2 * pair + b;
      if (t >= n_tiles) break;
      const long long tile_p0 = p0 + (long long)t * 128;
      const int nvalid = (int)min((long long)128, p1 - tile_p0);
      const long long p = tile_p0 + row;
      const bool valid = row < nvalid;

      // ---- tile inputs + embedding -> smem (A operand of layers 1, 5 and dir)
      float z = 0.f, delta = 0.f, nz = 0.f;
      int s = 0;
      long long ray = 0;
      if (ray_mode) {
        float xyz[3] = {0.f, 0.f, 0.f}, vd[3] = {0.f, 0.f, 0.f};
        if (valid) {
          ray = p / P.S;
          s = (int)(p - ray * P.S);
          z = __ldg(P.z_vals + p);
          delta = (s + 1 < P.S) ? __fsub_rn(__ldg(P.z_vals + p + 1), z) : 1e2f;
          nz = P.noise ? __ldg(P.noise + p) : 0.f;
          const float4 r0 = __ldg(reinterpret_cast<const float4*>(P.rays + ray * 8));
          const float4 r1 = __ldg(reinterpret_cast<const float4*>(P.rays + ray * 8) + 1);
          // xyz = o + d*z with separate roundings, as the reference's broadcast
          // mul then add (rendering.py:178)
          xyz[0] = __fadd_rn(r0.x, __fmul_rn(r0.w, z));
          xyz[1] = __fadd_rn(r0.y, __fmul_rn(r1.x, z));
          xyz[2] = __fadd_rn(r0.z, __fmul_rn(r1.y, z));
          if (P.view_dir) {
            vd[0] = __ldg(P.view_dir + ray * 3 + 0);
            vd[1] = __ldg(P.view_dir + ray * 3 + 1);
            vd[2] = __ldg(P.view_dir + ray * 3 + 2);
          } else {
            vd[0] = r0.w;
            vd[1] = r1.x;
            vd[2] = r1.y;
          }
        }
        embed3<kFmt>(my_emb, row, 0, xyz, P.n_freq_xyz);
        embed3<kFmt>(my_emb, row, kDirCol0, vd, P.n_freq_dir);
      } else {
        const float* xr = P.x + p * P.x_stride;
        for (int c = 0; c < P.e_xyz; ++c) emb_put<kFmt>(my_emb, row, c, valid ? __ldg(xr + c) : 0.f);
        for (int c = 0; c < P.e_dir; ++c)
          emb_put<kFmt>(my_emb, row, kDirCol0 + c,
                        (valid && !(P.mode & kModeSigmaOnly)) ? __ldg(xr + P.e_xyz + c) : 0.f);
      }
      for (int c = P.e_xyz; c < kDirCol0; ++c) emb_put<kFmt>(my_emb, row, c, 0.f);
      for (int c = kDirCol0 + P.e_dir; c < kEmbCols; ++c) emb_put<kFmt>(my_emb, row, c, 0.f);
      fence_proxy_async_smem();
      mbar_arrive(&M->emb_full[b]);

      uint32_t staged[64];
      float sig_acc = 0.f, sigma = 0.f, w_ray = 0.f;

      for (int u = 0; u < P.n_units; ++u) {
        const Unit un = c_units[u];
        mbar_wait(&M->d_full[b], ud & 1, 16 + u);
        ud++;
        tc_fence_after_sync();
        const float* bias = blob + kBiasOff(un.layer) + un.half * 128;
        const bool dump = P.dbg != nullptr && P.dbg_layer == un.layer && valid;

        if (un.layer <= kLFinal) {
          // ------------------------------------------------ 256-wide layers
          if (un.half == 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint32_t v[32];
              tmem_ld_x32(tD + 32 * c, v);
              tmem_ld_wait();
              if (c == 3) {
                tc_fence_before_sync();
                mbar_arrive(&M->d_empty[b]);
              }
              if (dump) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  float a = __uint_as_float(v[j]) + bias[32 * c + j];
                  P.dbg[p * 256 + 32 * c + j] = un.layer < 8 ? fmaxf(a, 0.f) : a;
                }
              }
              if (un.layer == 7)
                bias_act_pack<kFmt, true, true>(v, bias + 32 * c, wsig + 32 * c, &staged[16 * c],
                                                sig_acc);
              else if (un.layer < 8)
                bias_act_pack<kFmt, true, false>(v, bias + 32 * c, wsig, &staged[16 * c], sig_acc);
              else
                bias_act_pack<kFmt, false, false>(v, bias + 32 * c, wsig, &staged[16 * c], sig_acc);
            }
          } else {
            // every MMA of this layer has retired: A may be overwritten
#pragma unroll
            for (int c = 0; c < 4; ++c) tmem_st16(tA + 16 * c, &staged[16 * c]);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint32_t v[32], cur[16];
              tmem_ld_x32(tD + 32 * c, v);
              tmem_ld_wait();
              if (c == 3) {
                tc_fence_before_sync();
                mbar_arrive(&M->d_empty[b]);
              }
              if (dump) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  float a = __uint_as_float(v[j]) + bias[32 * c + j];
                  P.dbg[p * 256 + 128 + 32 * c + j] = un.layer < 8 ? fmaxf(a, 0.f) : a;
                }
              }
              if (un.layer == 7)
                bias_act_pack<kFmt, true, true>(v, bias + 32 * c, wsig + 128 + 32 * c, cur, sig_acc);
              else if (un.layer < 8)
                bias_act_pack<kFmt, true, false>(v, bias + 32 * c, wsig, cur, sig_acc);
              else
                bias_act_pack<kFmt, false, false>(v, bias + 32 * c, wsig, cur, sig_acc);
              tmem_st16(tA + 64 + 16 * c, cur);
            }
            tmem_st_wait();
            tc_fence_before_sync();
            mbar_arrive(&M->a_full[b]);

            if (un.layer == 7) {
              // ---------------- sigma head + alpha composite (rendering.py:121-143)
              sigma = softplus_ref(sig_acc + blob[kSigmaBOff]);
              if (!raw_mode) {
                const int s_first = (int)(tile_p0 % P.S);
                const float alpha =
                    valid ? 1.f - expf(-(delta * fmaxf(sigma + nz, 0.f))) : 0.f;
                const float om = 1.f - alpha;
                const int f0 = (valid && s == 0) ? 1 : 0;
                // inclusive segmented product over the warp's 32 rows
                float Pp = om;
                int F = f0;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                  const float pn = __shfl_up_sync(0xffffffffu, Pp, d);
                  const int fn = __shfl_up_sync(0xffffffffu, F, d);
                  if (lane >= d) {
                    if (!F) Pp *= pn;
                    F |= fn;
                  }
                }
                float Pe = __shfl_up_sync(0xffffffffu, Pp, 1);
                int Fe = __shfl_up_sync(0xffffffffu, F, 1);
                if (lane == 0) {
                  Pe = 1.f;
                  Fe = 0;
                }
                if (lane == 31) {
                  M->scan_p[b][q] = Pp;
                  M->scan_f[b][q] = F;
                }
                group_sync(b);
                // carry of the ray that straddles the previous tile boundary
                float cin_T = 1.f, cin_d = 0.f;
                if (t > 0) {
                  const uint32_t par = b ? (uint32_t)(pair & 1) : (uint32_t)((pair - 1) & 1);
                  mbar_wait(&M->carry_a[1 - b], par, 40);
                  if (s_first != 0) {
                    cin_T = M->carry_T[1 - b];
                    cin_d = M->carry_depth[1 - b];
                  }
                }
                float pre = cin_T;
                for (int w2 = 0; w2 < q; ++w2)
                  pre = M->scan_f[b][w2] ? M->scan_p[b][w2] : pre * M->scan_p[b][w2];
                const float T = f0 ? 1.f : (Fe ? Pe : pre * Pe);
                w_ray = alpha * T;
                if (valid) P.weights[p] = w_ray;
                // inclusive segmented sum of w*z for the depth map
                float Sd = w_ray * z;
                int F2 = f0;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                  const float sn = __shfl_up_sync(0xffffffffu, Sd, d);
                  const int fn = __shfl_up_sync(0xffffffffu, F2, d);
                  if (lane >= d) {
                    if (!F2) Sd += sn;
                    F2 |= fn;
                  }
                }
                if (lane == 31) M->scan_d[b][q] = Sd;
                group_sync(b);
                float pre_d = cin_d;
                for (int w2 = 0; w2 < q; ++w2)
                  pre_d = M->scan_f[b][w2] ? M->scan_d[b][w2] : pre_d + M->scan_d[b][w2];
                const float D_incl = F2 ? Sd : pre_d + Sd;
                const bool ray_end = valid && (s == P.S - 1);
                if (ray_end) P.depth[ray] = D_incl;
                if (row == nvalid - 1) {
                  M->carry_T[b] = ray_end ? 1.f : T * om;
                  M->carry_depth[b] = ray_end ? 0.f : D_incl;
                  mbar_arrive(&M->carry_a[b]);
                }
              }
            }
          }
        } else if (un.layer == kLDir) {
          // ------------------------------------------------ dir layer (128, ReLU)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t v[32], cur[16];
            tmem_ld_x32(tD + 32 * c, v);
            tmem_ld_wait();
            if (c == 3) {
              tc_fence_before_sync();
              mbar_arrive(&M->d_empty[b]);
            }
            if (dump) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                P.dbg[p * 256 + 32 * c + j] = fmaxf(__uint_as_float(v[j]) + bias[32 * c + j], 0.f);
            }
            bias_act_pack<kFmt, true, false>(v, bias + 32 * c, wsig, cur, sig_acc);
            tmem_st16(tA + 16 * c, cur);
          }
          tmem_st_wait();
          tc_fence_before_sync();
          mbar_arrive(&M->a_full[b]);
        } else {
          // ------------------------------------------------ rgb layer (64, sigmoid)
          // the embedding buffer is dead (dir layer retired): reuse it as the
          // (row, channel) staging area, XOR-swizzled so both the row-wise
          // writes and the channel-wise reads are bank-conflict free
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tmem_ld_x32(tD + 32 * c, v);
            tmem_ld_wait();
            if (c == 1) {
              tc_fence_before_sync();
              mbar_arrive(&M->d_empty[b]);
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float a = __uint_as_float(v[j]) + bias[32 * c + j];
              const float f = __fdividef(1.f, 1.f + __expf(-a));
              const int ch = 32 * c + j;
              if (raw_mode) {
                if (valid && !(P.mode & kModeSigmaOnly)) P.raw[p * 65 + ch] = f;
              } else {
                staging[row * 64 + (ch ^ (row & 31))] = w_ray * f;
              }
              if (dump) P.dbg[p * 256 + ch] = f;
            }
          }
          if (raw_mode) {
            if (valid) {
              if (P.mode & kModeSigmaOnly)
                P.raw[p] = sigma;
              else
                P.raw[p * 65 + 64] = sigma;
            }
            group_sync(b);  // staging/embedding buffer hand-over is uniform in both modes
          } else {
            group_sync(b);
            const int s_first = (int)(tile_p0 % P.S);
            const long long ray_first = tile_p0 / P.S;
            const int n_seg = (s_first + nvalid + P.S - 1) / P.S;
            const int cc = gtid & 63, hh = gtid >> 6;
            for (int sg = 0; sg < n_seg; ++sg) {
              const int r_beg = max(0, sg * P.S - s_first);
              const int r_end = min(nvalid, (sg + 1) * P.S - s_first);
              const int lo = max(r_beg, 64 * hh), hi = min(r_end, 64 * hh + 64);
              float acc = 0.f;
              for (int r = lo; r < hi; ++r) acc += staging[r * 64 + (cc ^ (r & 31))];
              M->part[b][hh][sg][cc] = acc;
            }
            group_sync(b);
            if (hh == 0) {
              if (t > 0) {
                const uint32_t par = b ? (uint32_t)(pair & 1) : (uint32_t)((pair - 1) & 1);
                mbar_wait(&M->carry_b[1 - b], par, 41);
              }
              float carry_out = 0.f;
              for (int sg = 0; sg < n_seg; ++sg) {
                float tot = (sg == 0 && s_first != 0) ? M->carry_feat[1 - b][cc] : 0.f;
                tot += M->part[b][0][sg][cc];
                tot += M->part[b][1][sg][cc];
                const bool ends = ((sg + 1) * P.S - s_first) <= nvalid;
                if (ends)
                  P.feature[(ray_first + sg) * 64 + cc] = tot;
                else
                  carry_out = tot;
              }
              M->carry_feat[b][cc] = carry_out;
              mbar_arrive(&M->carry_b[b]);
            }
          }
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem);
}

// ---------------------------------------------------------------------------
// weight packer: fp32 nn.Linear tensors -> swizzled 16-bit chunk image + blob
// ---------------------------------------------------------------------------
struct PackParams {
  const float* w[12];
  const float* b[12];
  uint8_t* img;
  float* blob;
  int32_t* status;
  int e_xyz, e_dir, n_chunks, image_bytes, fmt;
};

__global__ void pack_kernel(const __grid_constant__ PackParams P) {
  const int total16 = P.image_bytes / 16;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total16; i += gridDim.x * blockDim.x) {
    const int byte = i * 16;
    int ci = 0;
    while (ci + 1 < P.n_chunks && c_chunks[ci + 1].offset <= byte) ++ci;
    const Chunk ch = c_chunks[ci];
    const int within = byte - ch.offset;
    const int row = within >> 7;
    const int pos = (within & 127) >> 4;
    const int k0 = ((pos ^ (row & 7)) & 7) * 8;
    const int in_f = layer_in_features(ch.layer, P.e_xyz, P.e_dir);
    const float* wrow = P.w[ch.layer] + (long long)(ch.row0 + row) * in_f + ch.wcol0;
    uint32_t out[4];
    bool over = false;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float v0 = (k0 + 2 * e < ch.wcols) ? wrow[k0 + 2 * e] : 0.f;
      float v1 = (k0 + 2 * e + 1 < ch.wcols) ? wrow[k0 + 2 * e + 1] : 0.f;
      if (P.fmt == 0) {
        over |= fabsf(v0) > 65504.f || fabsf(v1) > 65504.f;
        out[e] = pack2<0, false>(v0, v1);
      } else {
        out[e] = pack2<1, false>(v0, v1);
      }
    }
    if (over && P.status) atomicExch(P.status, 1);
    *reinterpret_cast<uint4*>(P.img + byte) = make_uint4(out[0], out[1], out[2], out[3]);
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kBlobFloats; i += gridDim.x * blockDim.x) {
    float v = 0.f;
    if (i < kSigmaWOff) {
      const int layer = i < 2048 ? i / 256 : i < 2304 ? kLFinal : i < 2432 ? kLDir : kLRgb;
      const int j = i - kBiasOff(layer);
      if (j < layer_out_features(layer)) v = P.b[layer][j];
    } else if (i < kSigmaBOff) {
      v = P.w[kLSigma][i - kSigmaWOff];
    } else if (i == kSigmaBOff) {
      v = P.b[kLSigma][0];
    }
    P.blob[i] = v;
  }
}

// program tables live in __constant__ memory of this TU; upload once per
// (device, e_xyz, e_dir) and re-upload when the embedding widths change.
std::mutex g_prog_mu;
int g_prog_dev = -1, g_prog_exyz = -1, g_prog_edir = -1;
Program g_prog;

int ensure_program(int e_xyz, int e_dir, cudaStream_t st, const Program** out) {
  CRNERF_REQUIRE(e_xyz >= 3 && e_xyz <= kMaxExyz && e_dir >= 0 && e_dir <= kMaxEdir,
                 "embedding widths out of range: e_xyz=%d (<=96), e_dir=%d (<=32)", e_xyz, e_dir);
  int dev = 0;
  CRNERF_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_prog_mu);
  if (dev != g_prog_dev || e_xyz != g_prog_exyz || e_dir != g_prog_edir) {
    build_program(e_xyz, e_dir, &g_prog);
    // stream-ordered so that kernels already queued keep the tables they were launched with
    CRNERF_CUDA(cudaMemcpyToSymbolAsync(c_chunks, g_prog.chunks, sizeof(Chunk) * kMaxChunks, 0,
                                        cudaMemcpyHostToDevice, st));
    CRNERF_CUDA(cudaMemcpyToSymbolAsync(c_units, g_prog.units, sizeof(Unit) * kMaxUnits, 0,
                                        cudaMemcpyHostToDevice, st));
    // the tables are read by every later launch on any stream of this device
    CRNERF_CUDA(cudaStreamSynchronize(st));
    g_prog_dev = dev;
    g_prog_exyz = e_xyz;
    g_prog_edir = e_dir;
  }
  *out = &g_prog;
  return CRNERF_OK;
}

}  // namespace

size_t mlp_packed_bytes(int e_xyz, int e_dir) {
  Program p;
  build_program(e_xyz, e_dir, &p);
  return (size_t)p.image_bytes + sizeof(float) * kBlobFloats;
}

int mlp_pack(const crnerf_mlp_weights* w, int operand, void* packed, size_t packed_bytes,
             int32_t* status_dev, cudaStream_t st) {
  CRNERF_REQUIRE(w && packed, "null argument");
  CRNERF_REQUIRE(operand == 0 || operand == 1, "operand must be 0 (fp16) or 1 (bf16)");
  for (int i = 0; i < 12; ++i)
    CRNERF_REQUIRE(w->weight[i] && w->bias[i], "weight/bias pointer %d is null", i);
  const Program* prog;
  int rc = ensure_program(w->e_xyz, w->e_dir, st, &prog);
  if (rc) return rc;
  const size_t need = (size_t)prog->image_bytes + sizeof(float) * kBlobFloats;
  CRNERF_REQUIRE(packed_bytes >= need, "packed buffer too small: %zu < %zu", packed_bytes, need);
  CRNERF_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 15) == 0, "packed buffer must be 16-byte aligned");
  PackParams P;
  for (int i = 0; i < 12; ++i) {
    P.w[i] = w->weight[i];
    P.b[i] = w->bias[i];
  }
  P.img = static_cast<uint8_t*>(packed);
  P.blob = reinterpret_cast<float*>(P.img + prog->image_bytes);
  P.status = status_dev;
  P.e_xyz = w->e_xyz;
  P.e_dir = w->e_dir;
  P.n_chunks = prog->n_chunks;
  P.image_bytes = prog->image_bytes;
  P.fmt = operand;
  if (status_dev) CRNERF_CUDA(cudaMemsetAsync(status_dev, 0, sizeof(int32_t), st));
  pack_kernel<<<148, 256, 0, st>>>(P);
  count_launch();
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

// host-only: the chunk/unit program as flat int32 (tests emulate the kernel's tiling with it)
int debug_program(int e_xyz, int e_dir, int32_t* out, int cap) {
  CRNERF_REQUIRE(out && e_xyz >= 3 && e_xyz <= kMaxExyz && e_dir >= 0 && e_dir <= kMaxEdir, "bad argument");
  Program p;
  build_program(e_xyz, e_dir, &p);
  const int need = 3 + p.n_chunks * 10 + p.n_units * 7;
  CRNERF_REQUIRE(cap >= need, "buffer too small: need %d ints", need);
  int k = 0;
  out[k++] = p.n_chunks;
  out[k++] = p.n_units;
  out[k++] = p.image_bytes;
  for (int i = 0; i < p.n_chunks; ++i) {
    const Chunk& c = p.chunks[i];
    const int v[10] = {c.offset, c.bytes, c.layer, c.rows, c.row0, c.wcol0, c.wcols, c.a_src, c.a_k0, c.nk};
    for (int j = 0; j < 10; ++j) out[k++] = v[j];
  }
  for (int i = 0; i < p.n_units; ++i) {
    const Unit& u = p.units[i];
    const int v[7] = {u.layer, u.half, u.n, u.chunk0, u.nchunks, u.first_of_layer, u.last_of_layer};
    for (int j = 0; j < 7; ++j) out[k++] = v[j];
  }
  return need;
}

static int gcd_int(int a, int b) { return b ? gcd_int(b, a % b) : a; }

int launch_render(const RenderArgs& a, cudaStream_t st) {
  const int operand = a.operand, e_xyz = a.e_xyz, e_dir = a.e_dir;
  const Program* prog;
  int rc = ensure_program(e_xyz, e_dir, st, &prog);
  if (rc) return rc;
  RenderParams P;
  memset(&P, 0, sizeof(P));
  P.wimg = static_cast<const uint8_t*>(a.packed);
  P.blob = reinterpret_cast<const float*>(P.wimg + prog->image_bytes);
  P.rays = a.rays;
  P.view_dir = a.view_dir;
  P.z_vals = a.z_vals;
  P.noise = a.noise;
  P.x = a.x;
  P.x_stride = a.x_stride;
  P.weights = a.weights;
  P.feature = a.feature;
  P.depth = a.depth;
  P.raw = a.raw;
  P.dbg = g_dbg_buf;
  P.dbg_layer = g_dbg_layer;
  P.n_points = a.n_points;
  P.S = a.n_samples > 0 ? a.n_samples : 1;
  P.n_freq_xyz = a.n_freq_xyz;
  P.n_freq_dir = a.n_freq_dir;
  P.e_xyz = e_xyz;
  P.e_dir = e_dir;
  P.n_chunks = prog->n_chunks;
  P.n_units = prog->n_units;
  const int sms = num_sms();
  int grid;
  if (a.x) {
    P.mode = kModeEmbedded | kModeRaw | (a.sigma_only ? kModeSigmaOnly : 0);
    const long long tiles = (a.n_points + 127) / 128;
    long long tpc = (tiles + sms - 1) / sms;
    tpc += tpc & 1;  // whole X/Y pairs
    P.pts_per_cta = tpc * 128;
    grid = (int)((tiles + tpc - 1) / tpc);
  } else {
    P.mode = 0;
    const int S = a.n_samples;
    const int m = 128 / gcd_int(S, 128);  // rays per whole number of tiles
    long long rpc = (a.n_rays + sms - 1) / sms;
    if (rpc >= 2 * m) rpc = (rpc + m - 1) / m * m;
    P.pts_per_cta = rpc * S;
    grid = (int)((a.n_rays + rpc - 1) / rpc);
  }
  auto kern = operand == 0 ? render_fused_kernel<0> : render_fused_kernel<1>;
  CRNERF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  kern<<<grid, kThreads, kSmemBytes, st>>>(P);
  count_launch();
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

}  // namespace crnerf

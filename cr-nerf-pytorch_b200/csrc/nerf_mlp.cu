// Fused NeRF_sigma volume-rendering pass for sm_100a.
//
// Replaces, in one persistent kernel, the reference's `inference` closure
// (models/rendering.py:82-145): positional encoding (models/nerf.py:17-30),
// the 11-layer NeRF_sigma MLP (models/nerf.py:157-182) and the alpha composite
// (rendering.py:121-143).  No (points x width) intermediate touches HBM.
//
// Structure (one persistent CTA per SM, 640 threads = 20 warps):
//   warp 0      producer : streams the packed weight image L2 -> 8-slot x 16 KB smem ring with
//                          cp.async.bulk (TMA engine), mbarrier complete_tx
//   warps 1, 3  issuers  : one per tile stream (X, Y); an elected thread issues
//                          tcgen05.mma (kind::f16, M=128, N=128/64, K=16); activations
//                          are the A operand read from TMEM (TS form), the embedding is
//                          read from smem (SS form).  The two streams alternate whole
//                          units (layer halves) through a shared-memory turn word; every
//                          wait of a unit is taken before its turn, so a turn is one
//                          uninterrupted burst of 16 MMAs.
//   warp 2      TMEM allocator
//   warps 4-11  epilogue group X, warps 12-19 epilogue group Y (setmaxnreg 112): each group
//               owns one 128-point tile; a warp owns a TMEM lane quarter (32 rows, one per
//               lane) and a 64-column half of every 128-wide accumulator:
//               embedding -> smem; per layer half tcgen05.ld -> + fp32 bias -> ReLU ->
//               cvt.f16x2 -> (first halves wait in registers) -> tcgen05.st as the next
//               layer's A operand; sigma head as an fp32 dot in the layer-8 epilogue;
//               segmented warp-shuffle transmittance scan; in-warp butterfly feature sums.
// Two tiles (X, Y) are in flight per CTA and share every weight chunk.
// TMEM: X {A: cols 0-127, D: 128-255}, Y {A: 256-383, D: 384-511}; the first output half of
// a layer is held in registers until the layer's second half has retired, so A needs no
// double buffer, and the next layer may start on that half (a_half) while the second is
// still being packed.
//
// Split-precision variant (kSplit, operand "fp16x3"): every operand is x = hi + lo (two fp16
// values) and every product three MMAs (hi*hi + lo*hi + hi*lo, fp32 accumulate): fp32-class
// accuracy for trained weights, whose cancellation 11-bit operands cannot hold to 1e-4 (measured:
// 1-2e-3 on the trained-like set, tests/test_trained_weights.py).  A tile then needs A_hi, A_lo
// and D in TMEM, so ONE tile is in flight per CTA: stream X only, with the Y stream's columns
// re-used - TMEM {A_hi: 0-127, D0: 128-255, A_lo: 256-383, D1: 384-511}; the two output halves of
// a 256-wide layer accumulate in D0 / D1 and are drained together once both have retired (no
// register staging).  The ring becomes 4 slots x 32 KB (W_hi slab | W_lo slab), the embedding
// buffers of X / Y hold the embedding's hi / lo parts.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <algorithm>
#include <cstring>
#include <mutex>
#include "common.h"
#include "nerf_layout.h"
#include "ptx.cuh"

namespace crnerf {

namespace {

constexpr int kSlots = 8;
constexpr int kSlotBytes = 16384;
constexpr int kEmbBufBytes = 32768;
constexpr int kThreads = 640;        // 4 control warps + 2 tile groups x 8 epilogue warps
constexpr int kCtrlRegs = 32, kEpiRegs = 112;  // 4*32*32 + 16*32*112 = 61,440 = 640*96
constexpr int kMaxSeg = 10;  // ray segments per 128-row tile (n_samples >= 16)

constexpr int kRingOff = 0;
constexpr int kEmbOff = kSlots * kSlotBytes;        // 131072
constexpr int kBlobOff = kEmbOff + 2 * kEmbBufBytes;  // 196608
constexpr int kMiscOff = kBlobOff + kBlobFloats * 4;  // 207648

struct Misc {
  uint64_t ring_full[kSlots];
  uint64_t ring_empty[kSlots];
  uint64_t emb_full[2], a_full[2], a_half[2], d_full[2], d_empty[2], carry_a[2], carry_b[2];
  uint32_t tmem_base;
  uint32_t pipe_turn;  // whose unit goes down the tensor pipe next (even: stream X, odd: Y)
  // the issuers' copy of the program: constant-bank lookups indexed by a run-time value are
  // slow (tens to hundreds of cycles each), shared-memory loads are not
  //   unit_tab: n | first_of_layer << 8 | (layer == 0) << 9 | standard << 10 | half << 11 |
  //             chunk0 << 16 | nchunks << 24; "standard" = exactly four full activation slabs
  //   meta_tab: a_src | a_k0 << 8 | nk << 16 per chunk
  //   chunk_ob: byte offset in the image | (rows == 64) << 31, for the producer
  uint32_t unit_tab[kMaxUnits];
  uint32_t meta_tab[kMaxChunks];
  uint32_t chunk_ob[kMaxChunks];
  float scan_p[2][4];
  float scan_d[2][4];
  int scan_f[2][4];
  float carry_T[2];
  float carry_depth[2];
  float carry_feat[2][64];
  float sig_part[2][128];  // sigma-head partial sums of the second column-half warps
  float wray[2][128];      // per-row composite weights for the second column-half warps
  float part[2][4][kMaxSeg][64];
};
constexpr int kSmemBytes = kMiscOff + (int)sizeof(Misc);
static_assert(kSmemBytes <= 232448, "exceeds 227 KB of shared memory");
static_assert(kMiscOff % 16 == 0, "misc alignment");

enum : int { kModeEmbedded = 1, kModeRaw = 2, kModeSigmaOnly = 4 };

struct RenderParams {
  const uint8_t* wimg;
  const uint8_t* wimg_lo;  // split format: the W_lo image (same layout as wimg)
  const Tables* tab;       // program tables stored behind the blob in the packed buffer
  const float* jitter;     // optional (n_points, 3) added to xyz (args.pertubeCord, rendering.py:102-104)
  float* chan_part;        // optional (2*grid, 64): per (CTA, tile group) sums of the feature rows written
  int32_t* overflow;       // optional: OR-ed with 1 when an fp16 operand saturated (|x| >= 65504)
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
  uint16_t* acts;    // training: per-layer A-operand activations (see crnerf_render_pass_train)
  float* raw_save;   // training: (n_points, 64) sigmoid features, then the n_points softplus sigmas
  float* dbg;
  long long* prof;  // optional cycle counters of CTA 0 (tests / profiling only)
  int exp;          // profiling experiments (debug instantiation only): 1 = epilogue skips TMEM
                    // traffic and math, 2 = producer skips the weight copies
  long long n_points;
  long long pts_per_cta;
  int x_stride;
  int dbg_layer;
  int S;
  int n_freq_xyz, n_freq_dir, e_xyz, e_dir;
  int n_chunks, n_units;
  int mode;
};

// mbar_wait that adds the blocked cycles to a counter when profiling is on
__device__ __forceinline__ void timed_wait(uint64_t* bar, uint32_t parity, uint32_t tag, bool prof,
                                           long long& acc) {
  if (!prof) {
    mbar_wait(bar, parity, tag);
    return;
  }
  const long long t0 = clock64();
  mbar_wait(bar, parity, tag);
  acc += clock64() - t0;
}

// bounded spin on the pipe-turn word (same policy as mbar_wait: trap instead of hanging;
// a shared-memory load spins ~30 cycles per iteration, so 2^27 spins is seconds)
__device__ __forceinline__ void turn_wait(const uint32_t* turn, uint32_t want) {
  const volatile uint32_t* t = turn;
  uint32_t spins = 0;
  while (*t != want) {
    if (++spins == (1u << 27)) {
      g_wait_timeout_tag = 0x80000000u | (50u << 16) | (blockIdx.x & 0xffff);
      __trap();
    }
  }
}

// one arrival on behalf of a converged warp
__device__ __forceinline__ void warp_arrive(uint64_t* bar) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}

// named barriers: 1+b all 256 threads of tile group b; 3+b its 128 "column-half 0" threads;
// 5+b / 7+b producer-consumer pairs between the two column halves (arrive + sync = 256)
__device__ __forceinline__ void group_sync(int b) {
  asm volatile("bar.sync %0, 256;" ::"r"(1 + b) : "memory");
}
__device__ __forceinline__ void half_sync(int b) {
  asm volatile("bar.sync %0, 128;" ::"r"(3 + b) : "memory");
}
__device__ __forceinline__ void pc_arrive(int id) {
  asm volatile("bar.arrive %0, 256;" ::"r"(id) : "memory");
}
__device__ __forceinline__ void pc_sync(int id) {
  asm volatile("bar.sync %0, 256;" ::"r"(id) : "memory");
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
template <int kFmt, bool kSplit = false>
__device__ __forceinline__ void emb_put(uint8_t* buf, int row, int col, float v) {
  const uint32_t off = (uint32_t)(col >> 6) * 16384u + sw128_offset(row, (col & 63) >> 3) +
                       (uint32_t)(col & 7) * 2u;
  const uint16_t h = to_operand<kFmt>(v);
  *reinterpret_cast<uint16_t*>(buf + off) = h;
  if constexpr (kSplit)   // lo part into the second embedding buffer
    *reinterpret_cast<uint16_t*>(buf + kEmbBufBytes + off) =
        to_operand<0>(fminf(fmaxf(v, -65504.f), 65504.f) - __half2float(__ushort_as_half(h)));
}

// ---------------------------------------------------------------------------
// Positional encoding (reference models/nerf.py:17-30).
// sin/cos of 2^k * x with x/(2*pi) held as an unevaluated sum hi+lo.  The
// argument 2^k*x is exact in fp32 (power-of-two scale), so the only error is in
// the range reduction - done to ~2^-45 of a revolution in double-float
// arithmetic - and the SFU evaluation on |a| <= pi (~5e-7 abs), far below the
// 16-bit operand rounding (2^-11) applied right after.
// ---------------------------------------------------------------------------
struct EmbIn {
  float v[3], hi[3], lo[3];
};
__device__ __forceinline__ void emb_prepare(EmbIn& e) {
  const float chi = 0.15915493667125702f;    // fl32(1/2pi)
  const float clo = 6.420638316725915e-09f;  // 1/2pi - chi
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    e.hi[i] = e.v[i] * chi;
    e.lo[i] = fmaf(e.v[i], clo, fmaf(e.v[i], chi, -e.hi[i]));
  }
}
// fast path is valid while 2^14 |x| / 2pi < 2^22
__device__ __forceinline__ bool emb_fast_ok(const EmbIn& e) {
  return fmaxf(fmaxf(fabsf(e.v[0]), fabsf(e.v[1])), fabsf(e.v[2])) < 1024.f;
}
// sin/cos of band k, coordinate i.  Every third band is evaluated on the SFU from its own
// reduced argument; the two bands above it come from the double-angle identities
// (sin 2a = 2 sin a cos a, cos 2a = 1 - 2 sin^2 a: error x2 per step, <= 4e-6 after two,
// still far below the 2^-11 operand rounding).  The SFU is quarter rate and four warps per
// scheduler embed at the same time, so MUFU count - 114 per point without this, 48 with -
// is what bounds the embedding phase.  k and i are constants after unrolling; the compiler
// shares each (band, coordinate) evaluation between the columns that use it (CSE).
template <int k>
__device__ __forceinline__ void emb_sincos(const EmbIn& e, int i, float& s, float& c) {
  if constexpr (k % 3 == 0) {
    const float sc = __int_as_float((127 + k) << 23);
    const float hk = e.hi[i] * sc, lk = e.lo[i] * sc;
    const float a = ((hk - rintf(hk)) + lk) * 6.283185307179586f;
    s = __sinf(a);
    c = __cosf(a);
  } else {
    float s0, c0;
    emb_sincos<k - 1>(e, i, s0, c0);
    s = 2.f * s0 * c0;
    c = fmaf(-2.f * s0, s0, 1.f);
  }
}
template <int k>
__device__ __forceinline__ float emb_band(const EmbIn& e, int r) {
  float s, c;
  emb_sincos<k>(e, r % 3, s, c);
  return r < 3 ? s : c;
}
// column c of [v, sin(2^0 v), cos(2^0 v), ...]; c is a constant after unrolling
template <int kNFreq>
__device__ __forceinline__ float emb_col(const EmbIn& e, int c) {
  if (c < 3) return e.v[c];
  if (c >= 3 + 6 * kNFreq) return 0.f;
  const int k = (c - 3) / 6, r = (c - 3) % 6;
  switch (k) {
    case 0: return emb_band<0>(e, r);
    case 1: return emb_band<1>(e, r);
    case 2: return emb_band<2>(e, r);
    case 3: return emb_band<3>(e, r);
    case 4: return emb_band<4>(e, r);
    case 5: return emb_band<5>(e, r);
    case 6: return emb_band<6>(e, r);
    case 7: return emb_band<7>(e, r);
    case 8: return emb_band<8>(e, r);
    case 9: return emb_band<9>(e, r);
    case 10: return emb_band<10>(e, r);
    case 11: return emb_band<11>(e, r);
    case 12: return emb_band<12>(e, r);
    case 13: return emb_band<13>(e, r);
    default: return emb_band<14>(e, r);
  }
}
// 8*kNChunks consecutive columns -> kNChunks 16-byte stores into the swizzled buffer
// embedding-column chunks [kSrc0, kSrc0+kNChunks) -> buffer chunks [kDst0, ...)
// kSplit: also the lo parts (v - fp16(v)) into the second embedding buffer, kEmbBufBytes further on
template <int kFmt, int kNFreq, int kSrc0, int kDst0, int kNChunks, bool kSplit = false>
__device__ __forceinline__ void emb_write(uint8_t* buf, uint32_t row_off, uint32_t row_xor,
                                          const EmbIn& e) {
#pragma unroll
  for (int m = 0; m < kNChunks; ++m) {
    uint32_t w[4], wl[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float a = emb_col<kNFreq>(e, 8 * (kSrc0 + m) + 2 * q);
      const float b = emb_col<kNFreq>(e, 8 * (kSrc0 + m) + 2 * q + 1);
      w[q] = pack2<kFmt, false>(a, b);
      if constexpr (kSplit) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[q]));
        wl[q] = pack2<0, false>(a - f.x, b - f.y);
      }
    }
    const uint32_t cm = kDst0 + m;
    const uint32_t off = (cm >> 3) * 16384u + row_off + (((cm & 7u) << 4) ^ row_xor);
    *reinterpret_cast<uint4*>(buf + off) = make_uint4(w[0], w[1], w[2], w[3]);
    if constexpr (kSplit)
      *reinterpret_cast<uint4*>(buf + kEmbBufBytes + off) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
  }
}
// any band count / any magnitude: library sincos, element stores (rare path, kept out of line)
template <int kFmt, bool kSplit>
__device__ __noinline__ void embed3_generic(uint8_t* buf, int row, int col0, int col_end, float v0,
                                            float v1, float v2, int n_freqs) {
  const float v[3] = {v0, v1, v2};
#pragma unroll
  for (int i = 0; i < 3; ++i) emb_put<kFmt, kSplit>(buf, row, col0 + i, v[i]);
  for (int k = 0; k < n_freqs; ++k) {
    const float sc = __int_as_float((127 + k) << 23);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      float s, c;
      sincosf(v[i] * sc, &s, &c);
      emb_put<kFmt, kSplit>(buf, row, col0 + 3 + 6 * k + i, s);
      emb_put<kFmt, kSplit>(buf, row, col0 + 6 + 6 * k + i, c);
    }
  }
  for (int c = col0 + 3 + 6 * n_freqs; c < col_end; ++c)
    emb_put<kFmt, kSplit>(buf, row, c, 0.f);
}

__device__ __forceinline__ float softplus_ref(float x) {
  // torch.nn.Softplus(beta=1, threshold=20) (models/nerf.py:149)
  return x > 20.f ? x : log1pf(expf(x));
}

// Epilogue of one warp's 64 accumulator columns (bias already inside, see nerf_layout.h):
// one tcgen05.ld.x64, optional ReLU, pack to 32 words.  kSigma additionally accumulates
// this warp's share of the fp32 sigma-head dot product of the layer-8 activations.
// epi_stage / epi_flush below combine two slices into a layer-half epilogue.

// fp16 activations beyond 65504 must not saturate silently.  CRNERF_OVF_MODE selects how the
// epilogue notices (A/B-timed on B200, see profiles/README.md):
//   2 (default): activations are packed WITHOUT satfinite.  An overflow becomes +inf (post-ReLU)
//      or +-inf (no activation); inf is sticky through the remaining layers: every unit of the
//      next layer receives +-inf or NaN, cvt.relu keeps +inf and NaN, and whatever the route, the
//      rgb head's pre-sigmoid accumulators - fed by every earlier layer through xyz_encoding_final
//      and dir_encoding - end up non-finite.  The rgb epilogue folds them into one sticky value
//      (s = fma(a, 0, s): NaN iff some a is inf/NaN; 32 FFMA per thread and tile) and raises
//      RenderParams::overflow.  (The sigmoid itself would hide it: sigmoid(+-inf) is 1 or 0.)
//   1: packed with satfinite and a running maximum of the packed words (VIMNMX3.U16x2, one
//      instruction per four activations): a lane that reaches 0x7bff raises the flag;
//   0: satfinite, no report (the round-1 behaviour).
#ifndef CRNERF_OVF_MODE
#define CRNERF_OVF_MODE 2
#endif
constexpr bool kActSat = CRNERF_OVF_MODE != 2;
template <int kFmt, bool kRelu>
__device__ __forceinline__ void track_amax(uint32_t& amax, uint32_t w0, uint32_t w1) {
  if constexpr (kFmt == 0 && CRNERF_OVF_MODE == 1) {
    if constexpr (kRelu)
      amax = __vmaxu2(__vmaxu2(amax, w0), w1);   // post-ReLU: sign bits are clear
    else
      amax = __vmaxu2(__vmaxu2(amax, w0 & 0x7fff7fffu), w1 & 0x7fff7fffu);
  }
}
__device__ __forceinline__ bool amax_saturated(uint32_t amax) {
  return (amax & 0xffffu) >= 0x7bffu || (amax >> 16) >= 0x7bffu;
}

template <int kFmt, bool kRelu, bool kSigma, bool kDbg, int kOff, int kN, int kCols = 32>
__device__ __forceinline__ void epi_slice32(const uint32_t (&v)[kCols], const float* __restrict__ blob_g,
                                            uint32_t boff, const float* wsig, uint32_t (&out)[kN],
                                            float& sig_acc, float* dbg, uint32_t& amax) {
  // blob_g is the CTA's shared-memory copy of the side blob, boff an element offset
  const float* bias = blob_g + boff;
  // bias: consecutive fp32 values of the side blob in shared memory (same address in every
  // lane -> one broadcast wavefront per 16 bytes), added in fp32 before the activation
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 bv = *(reinterpret_cast<const float4*>(bias) + j);
    const float x0 = __uint_as_float(v[4 * j]) + bv.x, x1 = __uint_as_float(v[4 * j + 1]) + bv.y;
    const float x2 = __uint_as_float(v[4 * j + 2]) + bv.z, x3 = __uint_as_float(v[4 * j + 3]) + bv.w;
    if constexpr (kDbg) {
      if (dbg) {
        dbg[4 * j] = kRelu ? fmaxf(x0, 0.f) : x0;
        dbg[4 * j + 1] = kRelu ? fmaxf(x1, 0.f) : x1;
        dbg[4 * j + 2] = kRelu ? fmaxf(x2, 0.f) : x2;
        dbg[4 * j + 3] = kRelu ? fmaxf(x3, 0.f) : x3;
      }
    }
    if constexpr (kSigma) {
      const float4 ws = *reinterpret_cast<const float4*>(wsig + 4 * j);
      sig_acc = fmaf(fmaxf(x0, 0.f), ws.x, sig_acc);
      sig_acc = fmaf(fmaxf(x1, 0.f), ws.y, sig_acc);
      sig_acc = fmaf(fmaxf(x2, 0.f), ws.z, sig_acc);
      sig_acc = fmaf(fmaxf(x3, 0.f), ws.w, sig_acc);
    }
    out[kOff + 2 * j] = pack2<kFmt, kRelu, kActSat>(x0, x1);
    out[kOff + 2 * j + 1] = pack2<kFmt, kRelu, kActSat>(x2, x3);
    track_amax<kFmt, kRelu>(amax, out[kOff + 2 * j], out[kOff + 2 * j + 1]);
  }
}

// First half of a 256-wide layer: drain this warp's 64 accumulator columns into 32 packed
// words that stay in registers (A is still being read by the layer's second half).
// 16 packed words (32 activations of one row) -> 64 contiguous bytes of the saved-activation buffer
// Saved activations live in the backward kernels' operand layout (csrc/backward_gemm.cu,
// "tiled16"): per 128-point tile and 64-column slab a 16 KB block of 128-byte rows whose
// 16-byte chunks are XOR-swizzled with (row % 8) - readable by bulk copies as a SWIZZLE_128B
// tcgen05 operand, K-major (dgrad) and MN-major (wgrad) alike.  A thread owns one row.
struct SaveRow {
  uint4* base;   // the row's 128 bytes inside its slab, or nullptr
  uint32_t rx;   // global row % 8
};
// 16 packed words = chunks [chunk0, chunk0 + 4) of the row
__device__ __forceinline__ void save16(const SaveRow& r, int chunk0, const uint32_t* w) {
#pragma unroll
  for (int i = 0; i < 2; ++i)   // whole 32-byte sectors per thread (STG.256, see stg_v8)
    stg_row_pair(r.base, (uint32_t)(chunk0 >> 1) + i, r.rx, make_uint4(w[8 * i], w[8 * i + 1], w[8 * i + 2], w[8 * i + 3]),
                 make_uint4(w[8 * i + 4], w[8 * i + 5], w[8 * i + 6], w[8 * i + 7]));
}

template <int kFmt, bool kRelu, bool kSigma, bool kDbg, bool kSave>
__device__ __forceinline__ void epi_stage(uint32_t tD_ch, const float* blob_g, uint32_t boff, const float* wsig_ch,
                                          uint32_t (&staged)[32], float& sig_acc, uint64_t* d_empty,
                                          float* dbg, bool skip, const SaveRow& asave, uint32_t& amax) {
  if (kDbg && skip) {
    tc_fence_before_sync();
    warp_arrive(d_empty);
    return;
  }
  // both loads first and the "drained" signal right behind them: the issuer's next unit waits
  // on it, the packing below does not
  uint32_t va[32], vb[32];
  tmem_ld_x32(tD_ch, va);
  tmem_ld_x32(tD_ch + 32, vb);
  tmem_ld_wait();
  tc_fence_before_sync();
  warp_arrive(d_empty);  // accumulator drained: the issuer may overwrite it
  epi_slice32<kFmt, kRelu, kSigma, kDbg, 0, 32>(va, blob_g, boff, wsig_ch, staged, sig_acc, dbg, amax);
  epi_slice32<kFmt, kRelu, kSigma, kDbg, 16, 32>(vb, blob_g, boff + 32u, wsig_ch + 32, staged, sig_acc,
                                                 dbg ? dbg + 32 : nullptr, amax);
  if constexpr (kSave) {
    if (asave.base) {
      save16(asave, 0, staged);
      save16(asave, 4, staged + 16);
    }
  }
}

// Second half (kDirect == false): every MMA of the layer has retired, so the staged first
// half goes to A columns [32ch, 32ch+32) and this half's 64 columns to [64+32ch, ...).
// kDirect (dir layer, 128 wide): this warp's 64 columns go to A columns [32ch, 32ch+32).
template <int kFmt, bool kRelu, bool kSigma, bool kDbg, bool kDirect, bool kSave>
__device__ __forceinline__ void epi_flush(uint32_t tD_ch, uint32_t tA_ch, const float* blob_g, uint32_t boff,
                                          const float* wsig_ch, const uint32_t (&staged)[32], float& sig_acc,
                                          uint64_t* d_empty, uint64_t* a_full, uint64_t* a_half, float* dbg,
                                          bool skip, const SaveRow& asave, uint32_t& amax) {
  if (kDbg && skip) {
    tc_fence_before_sync();
    warp_arrive(d_empty);
    warp_arrive(a_half);
    warp_arrive(a_full);
    return;
  }
  if constexpr (!kDirect) tmem_st_x32(tA_ch, staged);
  const uint32_t a_dst = tA_ch + (kDirect ? 0u : 64u);
  uint32_t va[32], vb[32], out[16];
  tmem_ld_x32(tD_ch, va);
  tmem_ld_wait();
  tmem_ld_x32(tD_ch + 32, vb);
  epi_slice32<kFmt, kRelu, kSigma, kDbg, 0, 16>(va, blob_g, boff, wsig_ch, out, sig_acc, dbg, amax);
  tmem_st_x16p(a_dst, out);
  if constexpr (kSave) {
    if (asave.base) save16(asave, 0, out);
  }
  tmem_ld_wait();
  tc_fence_before_sync();
  warp_arrive(d_empty);  // accumulator drained: the issuer may overwrite it
  if constexpr (!kDirect) {
    // the staged first half (A columns 0..63 = the next layer's K 0..127) is in place: the
    // issuer may start the next layer's first two slabs while this half is still being packed
    tmem_st_wait();
    tc_fence_before_sync();
    warp_arrive(a_half);
  }
  epi_slice32<kFmt, kRelu, kSigma, kDbg, 0, 16>(vb, blob_g, boff + 32u, wsig_ch + 32, out, sig_acc,
                                                dbg ? dbg + 32 : nullptr, amax);
  tmem_st_x16p(a_dst + 16, out);
  if constexpr (kSave) {
    if (asave.base) save16(asave, 4, out);
  }
  // A holds the next layer's full input
  tmem_st_wait();
  tc_fence_before_sync();
  if constexpr (kDirect) warp_arrive(a_half);
  warp_arrive(a_full);
}

// Split-precision epilogue of one warp's 64 accumulator columns (one output half of a layer):
// bias, activation, x = hi + lo with hi = fp16(x), lo = fp16(x - hi); hi goes to A_hi columns
// [0,32) behind tAh, lo to the same columns of A_lo.  The accumulator is released (d_empty) as soon
// as both 32-column loads have landed.
template <bool kRelu, bool kSigma>
__device__ __forceinline__ void epi_split_half(uint32_t tD_ch, uint32_t tAh, uint32_t tAl, const float* blob_s,
                                               uint32_t boff, const float* wsig_ch, float& sig_acc,
                                               uint64_t* d_empty, uint32_t& amax) {
#pragma unroll
  for (int sl = 0; sl < 2; ++sl) {
    uint32_t v[32];
    tmem_ld_x32(tD_ch + 32u * sl, v);
    tmem_ld_wait();
    if (sl == 1) {
      tc_fence_before_sync();
      warp_arrive(d_empty);
    }
    uint32_t oh[16], ol[16];
    const float* bias = blob_s + boff + 32u * sl;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 bv = *(reinterpret_cast<const float4*>(bias) + j);
      float x0 = __uint_as_float(v[4 * j]) + bv.x, x1 = __uint_as_float(v[4 * j + 1]) + bv.y;
      float x2 = __uint_as_float(v[4 * j + 2]) + bv.z, x3 = __uint_as_float(v[4 * j + 3]) + bv.w;
      if constexpr (kRelu) {   // x < 0 ? 0 : x keeps NaN (fmaxf would drop it; see CRNERF_OVF_MODE)
        x0 = x0 < 0.f ? 0.f : x0;
        x1 = x1 < 0.f ? 0.f : x1;
        x2 = x2 < 0.f ? 0.f : x2;
        x3 = x3 < 0.f ? 0.f : x3;
      }
      if constexpr (kSigma) {
        const float4 ws = *reinterpret_cast<const float4*>(wsig_ch + 32 * sl + 4 * j);
        sig_acc = fmaf(x0, ws.x, sig_acc);
        sig_acc = fmaf(x1, ws.y, sig_acc);
        sig_acc = fmaf(x2, ws.z, sig_acc);
        sig_acc = fmaf(x3, ws.w, sig_acc);
      }
      const uint32_t h0 = pack2<0, false, kActSat>(x0, x1), h1 = pack2<0, false, kActSat>(x2, x3);
      const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&h0));
      const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&h1));
      oh[2 * j] = h0;
      oh[2 * j + 1] = h1;
      ol[2 * j] = pack2<0, false>(x0 - f0.x, x1 - f0.y);
      ol[2 * j + 1] = pack2<0, false>(x2 - f1.x, x3 - f1.y);
      track_amax<0, false>(amax, h0, h1);
    }
    tmem_st_x16p(tAh + 16u * sl, oh);
    tmem_st_x16p(tAl + 16u * sl, ol);
  }
}

// Column sums over the 32 lanes of a warp for 32 values per lane, by a transposing butterfly:
// after the step with distance h a lane keeps the half of its values whose index has bit h
// equal to the lane's bit h, so after five steps lane L holds sum_over_lanes(v[L]).
// 31 shuffles + 31 adds instead of a shared-memory transpose and two barriers.
template <int kHalf>
__device__ __forceinline__ void bfly_step(float (&v)[32], int lane) {
  const bool up = (lane & kHalf) != 0;
#pragma unroll
  for (int i = 0; i < kHalf; ++i) {
    const float send = up ? v[i] : v[i + kHalf];
    const float keep = up ? v[i + kHalf] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, kHalf);
  }
}
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
  bfly_step<16>(v, lane);
  bfly_step<8>(v, lane);
  bfly_step<4>(v, lane);
  bfly_step<2>(v, lane);
  bfly_step<1>(v, lane);
  return v[0];
}

// kVariant: 0 production, 1 debug/profiling instrumentation, 2 training forward (also stores
// every layer's A-operand activations and the per-point [features | sigma] for the backward)
template <int kFmt, int kVariant, bool kSplit>
__global__ void __launch_bounds__(kThreads, 1)
render_fused_kernel(const __grid_constant__ RenderParams P) {
  constexpr bool kDbg = kVariant == 1;
  constexpr bool kSave = kVariant == 2;
  static_assert(!kSplit || (kFmt == 0 && kVariant == 0), "the split format is fp16 hi/lo, inference only");
  constexpr int kStreams = kSplit ? 1 : 2;          // tiles in flight per CTA
  constexpr int kRing = kSplit ? 4 : kSlots;        // ring slots (split: 32 KB = W_hi slab | W_lo slab)
  constexpr int kRingSlot = kSplit ? 2 * kSlotBytes : kSlotBytes;
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
  const int n_pairs = (n_tiles + kStreams - 1) / kStreams;   // rounds of kStreams tiles

  if (tid == 0) {
    for (int i = 0; i < kSlots; ++i) {
      mbar_init(&M->ring_full[i], 1);
      // released by stream Y alone: the turn order puts X's use of a chunk ahead of Y's on the
      // in-order tensor pipe, so Y's commit covers both (X commits only when it runs solo)
      mbar_init(&M->ring_empty[i], 1);
    }
    for (int b = 0; b < 2; ++b) {
      // epilogue-side barriers take ONE arrival per warp (an elected lane after the warp's
      // collective tcgen05.wait / __syncwarp): a 32-lane arrive is 32 serialized smem atomics
      mbar_init(&M->emb_full[b], 8);
      mbar_init(&M->a_full[b], 8);
      mbar_init(&M->a_half[b], 8);
      mbar_init(&M->d_full[b], 1);
      mbar_init(&M->d_empty[b], 8);
      mbar_init(&M->carry_a[b], 1);
      mbar_init(&M->carry_b[b], 2);
    }
    M->pipe_turn = 0u;
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(&M->tmem_base);
  for (int i = tid; i < kBlobFloats; i += kThreads) blob[i] = P.blob[i];
  for (int i = tid; i < P.n_units; i += kThreads) {
    const Unit un = P.tab->units[i];
    bool standard = un.nchunks == 4;
    for (int j = 0; standard && j < 4; ++j) {
      const Chunk c = P.tab->chunks[un.chunk0 + j];
      standard = c.a_src == kSrcAct && c.a_k0 == 4 * j && c.nk == 4;
    }
    M->unit_tab[i] = (uint32_t)un.n | ((uint32_t)(un.first_of_layer != 0) << 8) |
                     ((uint32_t)(un.layer == 0) << 9) | ((uint32_t)standard << 10) |
                     ((uint32_t)(un.half != 0) << 11) | ((uint32_t)un.chunk0 << 16) |
                     ((uint32_t)un.nchunks << 24);
  }
  for (int i = tid; i < P.n_chunks; i += kThreads) {
    const Chunk c = P.tab->chunks[i];
    M->meta_tab[i] = (uint32_t)c.a_src | ((uint32_t)c.a_k0 << 8) | ((uint32_t)c.nk << 16);
    M->chunk_ob[i] = (uint32_t)c.offset | (c.rows == 64 ? 0x80000000u : 0u);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = M->tmem_base;

  // Register rebalancing inside the CTA's launch allocation (640 threads x 96 registers =
  // 61,440; setmaxnreg cannot draw from outside it): the 4 control warps drop to kCtrlRegs,
  // the 16 epilogue warps rise to kEpiRegs.  setmaxnreg is a warpgroup-wide instruction:
  // warps 0-3 (one warpgroup) must all use the same value.  Each call sits at the top of
  // its role branch - ptxas only raises a region's register budget when the instruction
  // dominates it.
  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    setmaxnreg_dec<kCtrlRegs>();
    if (elect_one()) {
      const uint64_t pol = l2_policy_evict_last();
      uint32_t g = 0;
      for (int pair = 0; pair < n_pairs; ++pair) {
        for (int c = 0; c < P.n_chunks; ++c, ++g) {
          const uint32_t slot = g % kRing, n = g / kRing;
          mbar_wait(&M->ring_empty[slot], (n & 1) ^ 1, 1);
          const uint32_t ob = M->chunk_ob[c];
          const uint32_t bytes = (ob >> 31) ? 8192u : 16384u, off = ob & 0x7fffffffu;
          if (kDbg && (P.exp & 2)) {
            mbar_arrive(&M->ring_full[slot]);
            continue;
          }
          mbar_arrive_expect_tx(&M->ring_full[slot], kSplit ? 2u * bytes : bytes);
          bulk_g2s_hint(ring + slot * kRingSlot, P.wimg + off, bytes, &M->ring_full[slot], pol);
          if constexpr (kSplit)
            bulk_g2s_hint(ring + slot * kRingSlot + kSlotBytes, P.wimg_lo + off, bytes, &M->ring_full[slot], pol);
        }
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    setmaxnreg_dec<kCtrlRegs>();  // TMEM allocator warp: idle until teardown
  } else if (warp == 1 || warp == 3) {
    // ------------------------------------------------------------------- issuers
    setmaxnreg_dec<kCtrlRegs>();
    if constexpr (kSplit) {
      // ---- split-precision issuer: stream X only (warp 3 idles), table-driven, three MMAs per
      // k-step: A_hi*W_hi, A_lo*W_hi, A_hi*W_lo.  The two output halves of a 256-wide layer go
      // to D0 / D1 (stream Y's accumulator columns and barriers), so the issuer runs a whole
      // layer ahead of the epilogue.  Weight chunks are awaited one by one inside the unit (the
      // 4-slot ring is smaller than the skip layer's six chunks) and released per chunk.
      if (warp == 1) {
        constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO | version | SW128
        const uint32_t ring_lo = ((smem_u32(ring) & 0x3ffffu) >> 4) | (1u << 16);
        const uint32_t emb_lo = ((smem_u32(emb) & 0x3ffffu) >> 4) | (1u << 16);
        auto desc = [](uint32_t lo) { return (static_cast<uint64_t>(kDescHi) << 32) | lo; };
        const uint32_t tAh = tmem, tAl = tmem + 256u;
        uint32_t ucnt[2] = {0u, 0u}, acount = 0, g_base = 0;
        for (int pair = 0; pair < n_pairs; ++pair) {
          for (int u = 0; u < P.n_units; ++u) {
            const uint32_t ut = M->unit_tab[u];
            const uint32_t idesc = make_idesc_f16(128, ut & 0xffu, 0);
            const int chunk0 = (int)((ut >> 16) & 0xffu), nch = (int)(ut >> 24);
            const uint32_t g0 = g_base + (uint32_t)chunk0;
            const int acc = (ut >> 11) & 1;                   // output half 1 -> D1
            const uint32_t tDu = tmem + (acc ? 384u : 128u);
            if (ut & 0x100u) {
              if (ut & 0x200u) {
                mbar_wait(&M->emb_full[0], (uint32_t)pair & 1, 2);
              } else {
                mbar_wait(&M->a_full[0], acount & 1, 3);
                acount++;
              }
            }
            mbar_wait(&M->d_empty[acc], (ucnt[acc] & 1) ^ 1, 4);
            tc_fence_after_sync();
            if (elect_one()) {
              for (int j = 0; j < nch; ++j) {
                const uint32_t gj = g0 + (uint32_t)j, slot = gj % kRing;
                mbar_wait(&M->ring_full[slot], (gj / kRing) & 1, 5);
                tc_fence_after_sync();
                const uint32_t meta = M->meta_tab[chunk0 + j];
                const uint32_t bh = ring_lo + slot * (kRingSlot >> 4), bl = bh + (kSlotBytes >> 4);
                const uint32_t a_k0 = (meta >> 8) & 0xffu;
                const uint32_t acc0 = j ? 1u : 0u;
                const int nk = (int)((meta >> 16) & 0xffu);
                if ((meta & 0x7fu) == (uint32_t)kSrcEmb) {
                  const uint32_t ah = emb_lo + (a_k0 >> 2) * 1024u + (a_k0 & 3u) * 2u;
                  const uint32_t al = ah + (kEmbBufBytes >> 4);
                  for (int k = 0; k < nk; ++k) {
                    umma_ss(tDu, desc(ah + 2u * k), desc(bh + 2u * k), idesc, k ? 1u : acc0);
                    umma_ss(tDu, desc(al + 2u * k), desc(bh + 2u * k), idesc, 1u);
                    umma_ss(tDu, desc(ah + 2u * k), desc(bl + 2u * k), idesc, 1u);
                  }
                } else {
                  for (int k = 0; k < nk; ++k) {
                    const uint32_t ao = (a_k0 + (uint32_t)k) * 8u;
                    umma_ts(tDu, tAh + ao, desc(bh + 2u * k), idesc, k ? 1u : acc0);
                    umma_ts(tDu, tAl + ao, desc(bh + 2u * k), idesc, 1u);
                    umma_ts(tDu, tAh + ao, desc(bl + 2u * k), idesc, 1u);
                  }
                }
                umma_commit(&M->ring_empty[slot]);
              }
              umma_commit(&M->d_full[acc]);
            }
            __syncwarp();
            ucnt[acc]++;
          }
          g_base += (uint32_t)P.n_chunks;
        }
      }
    } else {
    // One issuer warp per tile stream (warp 1 -> X, warp 3 -> Y).  The tcgen05 issue
    // queue is only 1-2 instructions deep (measured: tools/umma_probe issue timestamps),
    // so the tensor pipe runs only while some thread is actually issuing; with two
    // independent streams one warp's barrier waits and bookkeeping hide under the other
    // warp's MMAs.  X and Y touch disjoint TMEM regions, so no ordering between the two
    // streams is needed; a ring slot is released when BOTH streams have committed it.
    // The whole warp walks the (warp-uniform) program so table loads and descriptor
    // arithmetic stay in uniform registers; tcgen05 instructions run under elect_one().
    const int b = warp == 1 ? 0 : 1;
    constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO | version | SW128
    const uint32_t ring_lo = ((smem_u32(ring) & 0x3ffffu) >> 4) | (1u << 16);
    const uint32_t emb_lo = ((smem_u32(emb) & 0x3ffffu) >> 4) | (1u << 16);
    auto desc = [](uint32_t lo) { return (static_cast<uint64_t>(kDescHi) << 32) | lo; };
    const uint32_t tA = tmem + (b ? 256u : 0u);
    const uint32_t tD = tmem + (b ? 384u : 128u);
    uint32_t ucount = 0, acount = 0;
    uint32_t g_base = 0;
    uint32_t my_turn = (uint32_t)b;  // pipe_turn value at which this stream may issue its next unit
    const bool prof = kDbg && P.prof != nullptr && blockIdx.x == 0;
    long long w_emb = 0, w_a = 0, w_d = 0, w_ring = 0, w_lock = 0, t_burst = 0;
    const long long t_start = clock64();
    for (int pair = 0; pair < n_pairs; ++pair) {
      if (2 * pair + b >= n_tiles) break;  // odd tile count: X finishes the last pair solo
      const bool solo = 2 * pair + 2 > n_tiles;  // odd tile count: X owns the last pair alone
      for (int u = 0; u < P.n_units; ++u) {
        const uint32_t ut = M->unit_tab[u];
        const uint32_t idesc = make_idesc_f16(128, ut & 0xffu, kFmt);
        const int chunk0 = (int)((ut >> 16) & 0xffu), nch = (int)(ut >> 24);
        const uint32_t g0 = g_base + (uint32_t)chunk0;
        uint32_t a_par = 0;
        // ---- everything this unit depends on is awaited BEFORE the pipe turn, so the turn
        // holder issues one uninterrupted burst: the tcgen05 queue is only 1-2 MMAs deep and
        // the pipe idles whenever the issuing thread does anything else for long.
        // weight chunks first: they were requested when the partner's unit before last freed
        // their slots and have normally landed long ago, so these waits cost nothing and are
        // off the epilogue -> issuer critical path
        for (int j = 0; j < nch; ++j)
          timed_wait(&M->ring_full[(g0 + j) % kSlots], ((g0 + j) / kSlots) & 1, 5, prof, w_ring);
        if (ut & 0x100u) {
          if (ut & 0x200u) {
            timed_wait(&M->emb_full[b], (uint32_t)pair & 1, 2, prof, w_emb);
          } else {
            // the previous layer's first output half (A columns 0..63) suffices to start a
            // standard unit; its second half is awaited inside the burst, two slabs later
            timed_wait(&M->a_half[b], acount & 1, 3, prof, w_a);
            if (!(ut & 0x400u)) timed_wait(&M->a_full[b], acount & 1, 3, prof, w_a);
            a_par = acount & 1;
            acount++;
          }
        }
        timed_wait(&M->d_empty[b], (ucount & 1) ^ 1, 4, prof, w_d);
        // Take the tensor pipe for this whole unit, strictly alternating X, Y, X, ...
        // Left alone the two streams issue MMA by MMA in lock step, finish their units
        // together and then both sit in their epilogues with the pipe idle; with unit-sized
        // turns one tile's epilogue runs under the other tile's MMAs.  (Strict order, not a
        // free-for-all lock: the weight ring holds two units, so a stream that got a whole
        // unit ahead while holding the pipe would wait for a slot its partner can only free
        // after taking the pipe.)
        {
          const long long t_l0 = prof ? clock64() : 0;
          if (lane == 0) turn_wait(&M->pipe_turn, my_turn);
          __syncwarp();
          if (prof) w_lock += clock64() - t_l0;
        }
        tc_fence_after_sync();
        const long long t_b0 = prof ? clock64() : 0;
        // ring slots are released by stream Y's commits (see ring_empty init)
        const bool release = b == 1 || solo;
        if (elect_one()) {
          if (ut & 0x400u) {
            // ---- standard unit: 16 TS MMAs over the four activation slabs, straight-line
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (j == 2 && (ut & 0x100u) && !(ut & 0x200u)) {
                mbar_wait(&M->a_full[b], a_par, 3);   // second half of the previous layer's output
                tc_fence_after_sync();
              }
              const uint32_t slot = (g0 + (uint32_t)j) % kSlots;
              const uint32_t b_lo = ring_lo + slot * (kSlotBytes >> 4);
              const uint32_t a_t = tA + 32u * j;
              umma_ts(tD, a_t, desc(b_lo), idesc, j ? 1u : 0u);
              umma_ts(tD, a_t + 8u, desc(b_lo + 2u), idesc, 1u);
              umma_ts(tD, a_t + 16u, desc(b_lo + 4u), idesc, 1u);
              umma_ts(tD, a_t + 24u, desc(b_lo + 6u), idesc, 1u);
            }
            // all 16 MMAs are queued: pass the pipe on before the (slow-to-issue) commits, so
            // the partner's first MMA follows ours without a gap.  Slots are released at unit
            // granularity - the ring holds two whole units, so nothing waits on the finer one.
            *(volatile uint32_t*)&M->pipe_turn = my_turn + (solo ? 2u : 1u);
            umma_commit(&M->d_full[b]);   // the epilogue's wake-up first, the ring releases after it
            if (release) {
#pragma unroll
              for (int j = 0; j < 4; ++j) umma_commit(&M->ring_empty[(g0 + (uint32_t)j) % kSlots]);
            }
          } else {
            // ---- embedding-fed and narrow units (layer 1, skip layer, dir, rgb): table-driven
            for (int j = 0; j < nch; ++j) {
              const uint32_t slot = (g0 + (uint32_t)j) % kSlots;
              const uint32_t meta = M->meta_tab[chunk0 + j];
              const uint32_t b_lo = ring_lo + slot * (kSlotBytes >> 4);
              const uint32_t a_k0 = (meta >> 8) & 0xffu;
              const uint32_t acc0 = j ? 1u : 0u;
              const int nk = (int)((meta >> 16) & 0xffu);
              if ((meta & 0x7fu) == (uint32_t)kSrcEmb) {
                const uint32_t a_lo = emb_lo + (uint32_t)b * (kEmbBufBytes >> 4) + (a_k0 >> 2) * 1024u +
                                      (a_k0 & 3u) * 2u;
                for (int k = 0; k < nk; ++k)
                  umma_ss(tD, desc(a_lo + 2u * k), desc(b_lo + 2u * k), idesc, k ? 1u : acc0);
              } else {
                const uint32_t a_t = tA + a_k0 * 8u;
                for (int k = 0; k < nk; ++k)
                  umma_ts(tD, a_t + 8u * k, desc(b_lo + 2u * k), idesc, k ? 1u : acc0);
              }
              // the last chunk's MMAs are queued: hand the pipe over before its commit
              if (j == nch - 1) *(volatile uint32_t*)&M->pipe_turn = my_turn + (solo ? 2u : 1u);
              if (release) umma_commit(&M->ring_empty[slot]);
            }
            umma_commit(&M->d_full[b]);
          }
        }
        __syncwarp();
        if (prof) t_burst += clock64() - t_b0;
        my_turn += 2u;
        // (the turn was passed inside the burst by the elected lane; writing it again here could
        // undo the partner's next hand-over)
        ucount++;
      }
      g_base += (uint32_t)P.n_chunks;
    }
    if (prof && lane == 0) {
      long long* o = P.prof + (b ? 24 : 0);
      o[0] = clock64() - t_start;  // issuer lifetime
      o[1] = w_emb;
      o[2] = w_a;
      o[3] = w_d;
      o[4] = w_ring;
      o[5] = n_pairs;
      o[6] = w_lock;
      o[7] = t_burst;  // turn acquired -> unit committed (includes the ring waits inside)
    }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue
    setmaxnreg_inc<kEpiRegs>();
    // tile group b (8 warps): warp gw owns TMEM lane quarter q = gw & 3 (rows 32q..32q+31,
    // one row per lane) and column half ch = gw >> 2 of every 128-wide accumulator.
    const int b = (warp - 4) >> 3;
    const int gw = (warp - 4) & 7;
    const int q = gw & 3, ch = gw >> 2;
    const int row = q * 32 + lane;
    const int gtid = gw * 32 + lane;  // 0..255 inside the group
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t tA = tmem + (b ? 256u : 0u) + lane_off;
    const uint32_t tD = tmem + (b ? 384u : 128u) + lane_off;
    // split format: group X works alone; A_lo and D1 live in stream Y's columns
    const bool active = !kSplit || b == 0;
    const uint32_t tAl = tmem + 256u + lane_off, tD1 = tmem + 384u + lane_off;
    uint32_t amax = 0;     // running max of packed fp16 magnitudes (overflow report)
    float csum = 0.f;      // channel sum of the feature rows this thread wrote (gtid < 64)
    uint8_t* my_emb = emb + b * kEmbBufBytes;
    const bool ray_mode = !(P.mode & kModeEmbedded);
    const bool skip = kDbg && (P.exp & 1);
    const bool raw_mode = (P.mode & kModeRaw) != 0 || skip;
    const bool fast_emb = P.n_freq_xyz == 15 && P.n_freq_dir == 4;
    const float* wsig = blob + kSigmaWOff;
    const uint32_t row_off = (uint32_t)row * 128u, row_xor = (uint32_t)(row & 7) << 4;
    uint32_t ud = 0, ud1 = 0;
    const bool prof = kDbg && P.prof != nullptr && blockIdx.x == 0 && gtid == 0;
    long long w_dfull = 0, t_emb = 0, t_comp = 0, t_red = 0, t_stage = 0, t_flush = 0;
    const long long t_start = clock64();

    // ---- tile inputs + embedding -> smem (A operand of layers 1, 5 and dir) for tile tt of
    // this stream.  Software-pipelined: tile t+2's embedding is produced at the end of tile t's
    // epilogue (its buffer is free once the dir unit has retired), before the cross-warp part
    // of tile t's feature reduction, so the issuer can start tile t+2 as soon as the rgb
    // accumulator is drained instead of waiting for reduction + loads + sin/cos.
    auto embed_tile = [&](int tt) {
      const long long tile_p0 = p0 + (long long)tt * 128;
      const int nvalid = (int)min((long long)128, p1 - tile_p0);
      const long long p = tile_p0 + row;
      const bool valid = row < nvalid;
        // ---- tile inputs + embedding -> smem (A operand of layers 1, 5, dir and of every
        // bias k-step).  Column half 0 writes buffer columns 0..63, half 1 columns 64..127.
        const long long t_e0 = prof ? clock64() : 0;
        if (ray_mode) {
          EmbIn ex, ed;
          ex.v[0] = ex.v[1] = ex.v[2] = 0.f;
          ed.v[0] = ed.v[1] = ed.v[2] = 0.f;
          if (valid) {
            const long long ray = p / P.S;
            const float z = __ldg(P.z_vals + p);
            const float4 r0 = __ldg(reinterpret_cast<const float4*>(P.rays + ray * 8));
            const float4 r1 = __ldg(reinterpret_cast<const float4*>(P.rays + ray * 8) + 1);
            // xyz = o + d*z with separate roundings, as the reference's broadcast
            // mul then add (rendering.py:178)
            ex.v[0] = __fadd_rn(r0.x, __fmul_rn(r0.w, z));
            ex.v[1] = __fadd_rn(r0.y, __fmul_rn(r1.x, z));
            ex.v[2] = __fadd_rn(r0.z, __fmul_rn(r1.y, z));
            if (P.jitter) {   // xyz_ += pertube_ratio * rand (rendering.py:102-104); scaled by the caller
              ex.v[0] = __fadd_rn(ex.v[0], __ldg(P.jitter + p * 3 + 0));
              ex.v[1] = __fadd_rn(ex.v[1], __ldg(P.jitter + p * 3 + 1));
              ex.v[2] = __fadd_rn(ex.v[2], __ldg(P.jitter + p * 3 + 2));
            }
            if (P.view_dir) {
              ed.v[0] = __ldg(P.view_dir + ray * 3 + 0);
              ed.v[1] = __ldg(P.view_dir + ray * 3 + 1);
              ed.v[2] = __ldg(P.view_dir + ray * 3 + 2);
            } else {
              ed.v[0] = r0.w;
              ed.v[1] = r1.x;
              ed.v[2] = r1.y;
            }
          }
          // warp-uniform choice so both halves of a row agree on the path
          const bool fast = fast_emb && __all_sync(0xffffffffu, emb_fast_ok(ex) && emb_fast_ok(ed));
          if (fast) {
            emb_prepare(ex);
            if (ch == 0) {
              emb_write<kFmt, 15, 0, 0, 8, kSplit>(my_emb, row_off, row_xor, ex);    // columns 0..63
            } else {
              emb_prepare(ed);
              emb_write<kFmt, 15, 8, 8, 4, kSplit>(my_emb, row_off, row_xor, ex);    // columns 64..95
              emb_write<kFmt, 4, 0, 12, 4, kSplit>(my_emb, row_off, row_xor, ed);   // columns 96..127
            }
          } else if (ch == 0) {
            embed3_generic<kFmt, kSplit>(my_emb, row, 0, kDirCol0, ex.v[0], ex.v[1], ex.v[2], P.n_freq_xyz);
          } else {
            embed3_generic<kFmt, kSplit>(my_emb, row, kDirCol0, kEmbCols, ed.v[0], ed.v[1], ed.v[2],
                                         P.n_freq_dir);
          }
          if constexpr (kFmt == 0) {   // a coordinate beyond the fp16 range is clamped by the operand cast
            if (fmaxf(fmaxf(fabsf(ex.v[0]), fabsf(ex.v[1])), fabsf(ex.v[2])) > 65504.f) amax = 0x7bff7bffu;
          }
        } else {
          const float* xr = P.x + p * P.x_stride;
          float big = 0.f;
          if (ch == 0) {
            for (int c = 0; c < P.e_xyz; ++c) {
              const float v = valid ? __ldg(xr + c) : 0.f;
              big = fmaxf(big, fabsf(v));
              emb_put<kFmt, kSplit>(my_emb, row, c, v);
            }
            for (int c = P.e_xyz; c < kDirCol0; ++c) emb_put<kFmt, kSplit>(my_emb, row, c, 0.f);
          } else {
            for (int c = 0; c < P.e_dir; ++c) {
              const float v = (valid && !(P.mode & kModeSigmaOnly)) ? __ldg(xr + P.e_xyz + c) : 0.f;
              big = fmaxf(big, fabsf(v));
              emb_put<kFmt, kSplit>(my_emb, row, kDirCol0 + c, v);
            }
            for (int c = kDirCol0 + P.e_dir; c < kEmbCols; ++c) emb_put<kFmt, kSplit>(my_emb, row, c, 0.f);
          }
          if constexpr (kFmt == 0) {
            if (big > 65504.f) amax = 0x7bff7bffu;
          }
        }
      fence_proxy_async_smem();
      warp_arrive(&M->emb_full[b]);
      if (prof) t_emb += clock64() - t_e0;
    };
    if (active && b < n_tiles) embed_tile(b);

    for (int pair = 0; active && pair < n_pairs; ++pair) {
      const int t = kStreams * pair + b;
      if (t >= n_tiles) break;
      const long long tile_p0 = p0 + (long long)t * 128;
      const int nvalid = (int)min((long long)128, p1 - tile_p0);
      const long long p = tile_p0 + row;
      const bool valid = row < nvalid;

      uint32_t staged[32];
      float sig_acc = 0.f, sigma = 0.f, w_ray = 0.f;
      const uint32_t tD_ch = tD + 64u * ch, tA_ch = tA + 32u * ch;
      // this warp's first column in the blob's bias table (layer l starts at 256 l, second
      // halves at +128; dir at 9*256, rgb at 9*256+128 - see bias_offset())
      const uint32_t bias_ch = (uint32_t)(kBiasOff + 64 * ch);
      uint64_t* const d_full = &M->d_full[b];
      uint64_t* const d_empty = &M->d_empty[b];
      uint64_t* const a_full = &M->a_full[b];
      uint64_t* const a_half = &M->a_half[b];
      // wait for the next accumulator of this tile (units arrive in program order)
      auto wait_d = [&]() {
        timed_wait(d_full, ud & 1, 16, prof, w_dfull);
        ud++;
        tc_fence_after_sync();
      };
      // split format: accumulator acc (0: D0, 1: D1) has its own full/empty pair and counter
      auto wait_d2 = [&](int acc) {
        mbar_wait(&M->d_full[acc], (acc ? ud1 : ud) & 1, 16);
        if (acc) ud1++; else ud++;
        tc_fence_after_sync();
      };
      // one 256-wide layer in the split format: both halves have retired -> A_hi / A_lo
      auto split_layer = [&](auto relu_tag, auto sigma_tag, uint32_t boff) {
        constexpr bool kR = decltype(relu_tag)::value, kS = decltype(sigma_tag)::value;
        wait_d2(0);
        wait_d2(1);
        epi_split_half<kR, kS>(tD_ch, tA_ch, tAl + 32u * ch, blob, boff, wsig + 64 * ch, sig_acc,
                               &M->d_empty[0], amax);
        epi_split_half<kR, kS>(tD1 + 64u * ch, tA_ch + 64u, tAl + 32u * ch + 64u, blob, boff + 128u,
                               wsig + 128 + 64 * ch, sig_acc, &M->d_empty[1], amax);
        tmem_st_wait();
        tc_fence_before_sync();
        warp_arrive(a_full);
      };
      // debug dump target of (layer, half) for this warp's 64 columns, or nullptr
      auto dbg_at = [&](int layer, int half) -> float* {
        if constexpr (kDbg) {
          if (P.dbg != nullptr && P.dbg_layer == layer && valid) return P.dbg + p * 256 + half * 128 + 64 * ch;
        }
        return nullptr;
      };
      // training: this warp's 64 activations of (layer, half) = one row of one slab of the
      // saved-activation buffer (tiled16; slots 0..8 = trunk 1-8 + final, four slabs per tile;
      // slot 9 = dir, two slabs; slot 10 = the embedding tile, two slabs - see save_emb)
      auto save_at = [&](int layer, int half) -> SaveRow {
        SaveRow r{nullptr, 0u};
        if constexpr (kSave) {
          if (P.acts != nullptr && valid) {
            const size_t T = (size_t)((P.n_points + 127) >> 7), tile = (size_t)(p >> 7);
            const uint32_t rg = (uint32_t)(p & 127);
            const size_t off = layer < 9 ? ((size_t)layer * T * 4 + tile * 4 + half * 2 + ch) * 16384
                                         : ((size_t)9 * T * 4 + ((layer - 9) * T + tile) * 2 + ch) * 16384;
            r.base = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(P.acts) + off + rg * 128u);
            r.rx = rg & 7u;
          }
        }
        return r;
      };
      // the tile's embedding rows (this thread wrote its own half-row in embed_tile) -> slot 10
      if constexpr (kSave) {
        const SaveRow er = save_at(10, 0);
        if (er.base) {
          const uint4* src = reinterpret_cast<const uint4*>(my_emb + ch * 16384 + row * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k) stg_row_pair(er.base, (uint32_t)k, er.rx, src[(2 * k) ^ (row & 7)], src[(2 * k + 1) ^ (row & 7)]);
        }
      }
      long long t_a = 0;

      // ---- trunk layers 1..7: ReLU.  The layer structure is spelled out here (the issuer
      // walks the same program from the tables) so that the 32 staged words live in
      // registers between a layer's two halves and nothing is indexed at run time.
#pragma unroll 1
      for (int layer = 0; layer < 7; ++layer) {
        const uint32_t bl = bias_ch + 256u * layer;
        if constexpr (kSplit) {
          split_layer(std::true_type{}, std::false_type{}, bl);
        } else {
          wait_d();
          if (prof) t_a = clock64();
          epi_stage<kFmt, true, false, kDbg, kSave>(tD_ch, blob, bl, wsig, staged, sig_acc, d_empty, dbg_at(layer, 0),
                                                    skip, save_at(layer, 0), amax);
          if (prof) t_stage += clock64() - t_a;
          wait_d();
          if (prof) t_a = clock64();
          epi_flush<kFmt, true, false, kDbg, false, kSave>(tD_ch, tA_ch, blob, bl + 128u, wsig, staged, sig_acc,
                                                           d_empty, a_full, a_half, dbg_at(layer, 1), skip,
                                                           save_at(layer, 1), amax);
          if (prof) t_flush += clock64() - t_a;
        }
      }
      // composite inputs of this row, requested now so the loads are long complete when the
      // sigma head is (they sit on this tile's critical path: the composite runs between
      // layer 8 and xyz_encoding_final on the same warps)
      const int s_first = (int)(tile_p0 % P.S);
      const long long ray_first = tile_p0 / P.S;
      const int seg_of_row = (s_first + row) / P.S;
      const int s_of_row = s_first + row - seg_of_row * P.S;
      float pre_z = 0.f, pre_z1 = 0.f, pre_nz = 0.f;
      if (!raw_mode && ch == 0 && valid) {
        pre_z = __ldg(P.z_vals + p);
        if (s_of_row + 1 < P.S) pre_z1 = __ldg(P.z_vals + p + 1);
        if (P.noise) pre_nz = __ldg(P.noise + p);
      }
      // ---- layer 8: ReLU + this warp's share of the fp32 sigma-head dot product
      if constexpr (kSplit) {
        split_layer(std::true_type{}, std::true_type{}, bias_ch + 256u * 7);
      } else {
        wait_d();
        if (prof) t_a = clock64();
        epi_stage<kFmt, true, true, kDbg, kSave>(tD_ch, blob, bias_ch + 256u * 7, wsig + 64 * ch, staged, sig_acc,
                                                 d_empty, dbg_at(7, 0), skip, save_at(7, 0), amax);
        if (prof) t_stage += clock64() - t_a;
        wait_d();
        if (prof) t_a = clock64();
        epi_flush<kFmt, true, true, kDbg, false, kSave>(tD_ch, tA_ch, blob, bias_ch + 256u * 7 + 128u,
                                                        wsig + 128 + 64 * ch, staged, sig_acc, d_empty, a_full,
                                                        a_half, dbg_at(7, 1), skip, save_at(7, 1), amax);
        if (prof) t_flush += clock64() - t_a;
      }
      {
        const long long t_c0 = prof ? clock64() : 0;
        // ---------------- sigma head + alpha composite (rendering.py:121-143), done by
        // the column-half-0 warps (one thread per row); half 1 hands over its partial dot
        if (ch == 1) {
          M->sig_part[b][row] = sig_acc;
          pc_arrive(7 + b);
        } else {
          pc_sync(7 + b);
          sigma = softplus_ref(sig_acc + M->sig_part[b][row] + blob[kSigmaBOff]);
          if constexpr (kSave) {
            if (P.raw_save != nullptr && valid) P.raw_save[(long long)P.n_points * 64 + p] = sigma;   // sigmas behind the feature rows
          }
          if (!raw_mode) {
            const long long ray = valid ? ray_first + seg_of_row : 0;
            const int s = valid ? s_of_row : 0;
            const float z = pre_z;
            const float delta = valid ? ((s + 1 < P.S) ? __fsub_rn(pre_z1, z) : 1e2f) : 0.f;
            const float nz = pre_nz;
            const float alpha = valid ? 1.f - expf(-(delta * fmaxf(sigma + nz, 0.f))) : 0.f;
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
            half_sync(b);
            // carry of the ray that straddles the previous tile boundary
            float cin_T = 1.f, cin_d = 0.f;
            if (t > 0) {
              // the previous tile belongs to the other group - or, in the split format, to this
              // group's previous iteration (ordered by the barriers in between)
              constexpr int kPrevOther = kSplit ? 0 : 1;
              const int pb = kPrevOther ? 1 - b : b;
              if constexpr (!kSplit) {
                const uint32_t par = b ? (uint32_t)(pair & 1) : (uint32_t)((pair - 1) & 1);
                mbar_wait(&M->carry_a[pb], par, 40);
              }
              if (s_first != 0) {
                cin_T = M->carry_T[pb];
                cin_d = M->carry_depth[pb];
              }
            }
            float pre = cin_T;
            for (int w2 = 0; w2 < q; ++w2)
              pre = M->scan_f[b][w2] ? M->scan_p[b][w2] : pre * M->scan_p[b][w2];
            const float T = f0 ? 1.f : (Fe ? Pe : pre * Pe);
            w_ray = alpha * T;
            if (valid) P.weights[p] = w_ray;
            M->wray[b][row] = w_ray;
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
            half_sync(b);
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
            pc_arrive(5 + b);  // wray[] is published for the half-1 warps
          }
        }
        if (prof) t_comp += clock64() - t_c0;
      }
      // ---- xyz_encoding_final: no activation
      if constexpr (kSplit) {
        split_layer(std::false_type{}, std::false_type{}, bias_ch + 256u * kLFinal);
        // ---- dir layer (128 wide, ReLU, one accumulator): A columns [0,64), hi and lo
        wait_d2(0);
        epi_split_half<true, false>(tD_ch, tA_ch, tAl + 32u * ch, blob, bias_ch + 256u * 9, wsig, sig_acc,
                                    &M->d_empty[0], amax);
        tmem_st_wait();
        tc_fence_before_sync();
        warp_arrive(a_full);
        // ---- rgb layer
        wait_d2(0);
      } else {
        wait_d();
        if (prof) t_a = clock64();
        epi_stage<kFmt, false, false, kDbg, kSave>(tD_ch, blob, bias_ch + 256u * kLFinal, wsig, staged, sig_acc,
                                                   d_empty, dbg_at(kLFinal, 0), skip, save_at(kLFinal, 0), amax);
        if (prof) t_stage += clock64() - t_a;
        wait_d();
        if (prof) t_a = clock64();
        epi_flush<kFmt, false, false, kDbg, false, kSave>(tD_ch, tA_ch, blob, bias_ch + 256u * kLFinal + 128u, wsig,
                                                          staged, sig_acc, d_empty, a_full, a_half, dbg_at(kLFinal, 1),
                                                          skip, save_at(kLFinal, 1), amax);
        if (prof) t_flush += clock64() - t_a;
        // ---- dir layer (128 wide, ReLU): straight to A columns [0,64)
        wait_d();
        if (prof) t_a = clock64();
        epi_flush<kFmt, true, false, kDbg, true, kSave>(tD_ch, tA_ch, blob, bias_ch + 256u * 9, wsig, staged, sig_acc,
                                                        d_empty, a_full, a_half, dbg_at(kLDir, 0), skip,
                                                        save_at(kLDir, 0), amax);
        if (prof) t_flush += clock64() - t_a;
        // ---- rgb layer
        wait_d();
      }
      {
        float* dbg_row = nullptr;
        if constexpr (kDbg) {
          if (P.dbg != nullptr && P.dbg_layer == kLRgb && valid) dbg_row = P.dbg + p * 256;
        }
        // ------------------------------------------------ rgb layer (64, sigmoid)
        // Each column half handles 32 of the 64 channels; w*f stays in registers and is summed
        // over the rows of each ray segment inside the warp (warp_transpose_sum).
        if (!raw_mode && ch == 1) {
          pc_sync(5 + b);
          w_ray = M->wray[b][row];
        }
        uint32_t v[32];
        float wf[32];
        [[maybe_unused]] float fsave[8];
        float sticky = 0.f;     // NaN once any rgb pre-activation is inf / NaN (fp16 operand overflow upstream)
        if (!skip) {
          tmem_ld_x32(tD + 32 * ch, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0;
        }
        tc_fence_before_sync();
        warp_arrive(&M->d_empty[b]);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int chn = 32 * ch + j;
          const float a = __uint_as_float(v[j]) + blob[bias_offset(kLRgb) + chn];
          if constexpr (kFmt == 0 && CRNERF_OVF_MODE == 2) sticky = fmaf(a, 0.f, sticky);
          const float f = __fdividef(1.f, 1.f + __expf(-a));
          if constexpr (kSave) {
            // saved features: (n_points, 64) rows, this warp's 32 channels as four whole 32-byte sectors
            fsave[j & 7] = f;
            if ((j & 7) == 7 && P.raw_save != nullptr && valid)
              stg_v8(P.raw_save + p * 64 + (chn - 7),
                     make_uint4(__float_as_uint(fsave[0]), __float_as_uint(fsave[1]), __float_as_uint(fsave[2]), __float_as_uint(fsave[3])),
                     make_uint4(__float_as_uint(fsave[4]), __float_as_uint(fsave[5]), __float_as_uint(fsave[6]), __float_as_uint(fsave[7])));
          }
          if (raw_mode) {
            if (valid && !skip && !(P.mode & kModeSigmaOnly)) P.raw[p * 65 + chn] = f;
          }
          wf[j] = w_ray * f;
          if constexpr (kDbg) {
            if (dbg_row) dbg_row[chn] = f;
          }
        }
        if (raw_mode) {
          if (valid && !skip && ch == 0) {
            if (P.mode & kModeSigmaOnly)
              P.raw[p] = sigma;
            else
              P.raw[p * 65 + 64] = sigma;
          }
        }
        if constexpr (kFmt == 0 && CRNERF_OVF_MODE == 2) {
          if (sticky != sticky) amax = 0x7bff7bffu;
        }
        const int n_seg = (s_first + nvalid + P.S - 1) / P.S;
        const long long t_r0 = prof ? clock64() : 0;
        if (!raw_mode) {
          // per-warp partial sums of every ray segment that crosses this warp's 32 rows
          for (int sg = 0; sg < n_seg; ++sg) {
            const int r_beg = max(0, sg * P.S - s_first);
            const int r_end = min(nvalid, (sg + 1) * P.S - s_first);
            float tot = 0.f;
            if (r_beg < 32 * q + 32 && r_end > 32 * q) {   // warp-uniform
              const bool in = row >= r_beg && row < r_end;
              float m[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) m[j] = in ? wf[j] : 0.f;
              tot = warp_transpose_sum(m, lane);
            }
            M->part[b][q][sg][32 * ch + lane] = tot;
          }
        }
        if (prof) t_red += clock64() - t_r0;
        // next tile of this stream: its embedding goes out before the cross-warp combine
        if (t + kStreams < n_tiles) embed_tile(t + kStreams);
        if (!raw_mode) {
          const long long t_r1 = prof ? clock64() : 0;
          const int cc = gtid & 63, hh = gtid >> 6;  // channel, row quarter
          group_sync(b);
          if (hh == 0) {
            if constexpr (!kSplit) {
              if (t > 0) {
                const uint32_t par = b ? (uint32_t)(pair & 1) : (uint32_t)((pair - 1) & 1);
                mbar_wait(&M->carry_b[1 - b], par, 41);
              }
            }
            float carry_out = 0.f;
            for (int sg = 0; sg < n_seg; ++sg) {
              float tot = (sg == 0 && s_first != 0) ? M->carry_feat[kSplit ? b : 1 - b][cc] : 0.f;
              tot += M->part[b][0][sg][cc];
              tot += M->part[b][1][sg][cc];
              tot += M->part[b][2][sg][cc];
              tot += M->part[b][3][sg][cc];
              const bool ends = ((sg + 1) * P.S - s_first) <= nvalid;
              if (ends) {
                P.feature[(ray_first + sg) * 64 + cc] = tot;
                csum += tot;
              } else {
                carry_out = tot;
              }
            }
            M->carry_feat[b][cc] = carry_out;
            warp_arrive(&M->carry_b[b]);
          }
          if (prof) t_red += clock64() - t_r1;
        }
      }
    }
    // per (CTA, tile group) sums of the feature rows written above: the cross-ray block's channel
    // mean (linearStyleTransfer.py:62) rides on this kernel instead of a pass over the feature map
    if (P.chan_part != nullptr && gtid < 64) P.chan_part[((size_t)blockIdx.x * 2 + b) * 64 + gtid] = csum;
    if constexpr (kFmt == 0) {
      // the flag may live in mapped pinned host memory: every writer stores the same 1
      if (P.overflow != nullptr && __any_sync(0xffffffffu, amax_saturated(amax)) && lane == 0) {
        *reinterpret_cast<volatile int32_t*>(P.overflow) = 1;
        __threadfence_system();
      }
    }
    if (prof) {
      long long* o = P.prof + 8 + 8 * b;
      o[0] = clock64() - t_start;  // epilogue-thread lifetime
      o[1] = w_dfull;              // blocked waiting for accumulators
      o[2] = t_emb;                // embedding phases
      o[3] = t_comp;               // sigma/alpha scan
      o[4] = t_red;                // feature reduction
      o[5] = t_stage + t_flush;    // all layer epilogues (wake-up -> a_full/d_empty arrive)
      o[6] = t_stage;              // ... of which first halves (drain to registers)
      o[7] = t_flush;              // ... of which second halves / dir (drain + write A)
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
  uint8_t* img;      // W (or W_hi) image
  uint8_t* img_lo;   // split format: W_lo image, else nullptr
  float* blob;
  Tables* tab_out;   // where the tables are stored in the packed buffer
  int32_t* status;
  int e_xyz, e_dir, fmt;
  Tables tab;        // the program, by value (constant bank): nothing to upload
};
static_assert(sizeof(PackParams) <= 4096, "kernel parameter space");

__global__ void pack_kernel(const __grid_constant__ PackParams P) {
  const int total16 = P.tab.image_bytes / 16;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total16; i += gridDim.x * blockDim.x) {
    const int byte = i * 16;
    int ci = 0;
    while (ci + 1 < P.tab.n_chunks && P.tab.chunks[ci + 1].offset <= byte) ++ci;
    const Chunk ch = P.tab.chunks[ci];
    const int within = byte - ch.offset;
    const int row = within >> 7;
    const int pos = (within & 127) >> 4;
    const int k0 = ((pos ^ (row & 7)) & 7) * 8;
    const int in_f = layer_in_features(ch.layer, P.e_xyz, P.e_dir);
    const float* wrow = P.w[ch.layer] + (long long)(ch.row0 + row) * in_f + ch.wcol0;
    uint32_t out[4], out_lo[4];
    bool over = false;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float v0 = (k0 + 2 * e < ch.wcols) ? wrow[k0 + 2 * e] : 0.f;
      float v1 = (k0 + 2 * e + 1 < ch.wcols) ? wrow[k0 + 2 * e + 1] : 0.f;
      if (P.fmt == 0) {
        over |= fabsf(v0) > 65504.f || fabsf(v1) > 65504.f;
        out[e] = pack2<0, false>(v0, v1);
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&out[e]));
        out_lo[e] = pack2<0, false>(v0 - f.x, v1 - f.y);   // used by the split format only
      } else {
        out[e] = pack2<1, false>(v0, v1);
        out_lo[e] = 0u;
      }
    }
    if (over && P.status) atomicExch(P.status, 1);
    *reinterpret_cast<uint4*>(P.img + byte) = make_uint4(out[0], out[1], out[2], out[3]);
    if (P.img_lo)
      *reinterpret_cast<uint4*>(P.img_lo + byte) = make_uint4(out_lo[0], out_lo[1], out_lo[2], out_lo[3]);
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kBlobFloats; i += gridDim.x * blockDim.x) {
    float v = 0.f;
    if (i < kSigmaBOff) {
      v = P.w[kLSigma][i - kSigmaWOff];
    } else if (i == kSigmaBOff) {
      v = P.b[kLSigma][0];
    } else if (i >= kBiasOff) {
      const int k = i - kBiasOff;
      const int layer = k < 9 * 256 ? k >> 8 : (k < 9 * 256 + 128 ? kLDir : kLRgb);
      v = P.b[layer][i - bias_offset(layer)];
    }
    P.blob[i] = v;
  }
  // the program tables travel with the image
  const uint32_t* src = reinterpret_cast<const uint32_t*>(&P.tab);
  uint32_t* dst = reinterpret_cast<uint32_t*>(P.tab_out);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (int)(sizeof(Tables) / 4); i += gridDim.x * blockDim.x)
    dst[i] = src[i];
}

// packed buffer: [W image][W_lo image (split format)][fp32 blob][tables]
struct PackedLayout {
  size_t img_lo, blob, tab, total;
};
int make_layout(int e_xyz, int e_dir, int operand, Program* prog, PackedLayout* L) {
  CRNERF_REQUIRE(e_xyz >= 3 && e_xyz <= kMaxExyz && e_dir >= 0 && e_dir <= kMaxEdir,
                 "embedding widths out of range: e_xyz=%d (<=96), e_dir=%d (<=32)", e_xyz, e_dir);
  CRNERF_REQUIRE(operand >= 0 && operand < kNumOperands, "operand must be 0 (fp16), 1 (bf16) or 2 (fp16x3)");
  build_program(e_xyz, e_dir, prog);
  const int n_img = operand_images(operand);
  L->img_lo = n_img == 2 ? (size_t)prog->image_bytes : 0;
  L->blob = (size_t)n_img * prog->image_bytes;
  L->tab = L->blob + sizeof(float) * kBlobFloats;
  L->total = L->tab + sizeof(Tables);
  return CRNERF_OK;
}

}  // namespace

size_t mlp_packed_bytes(int e_xyz, int e_dir, int operand) {
  Program p;
  PackedLayout L;
  if (make_layout(e_xyz, e_dir, operand, &p, &L)) return 0;
  return L.total;
}

int mlp_pack(const crnerf_mlp_weights* w, int operand, void* packed, size_t packed_bytes,
             int32_t* status_dev, cudaStream_t st) {
  CRNERF_REQUIRE(w && packed, "null argument");
  for (int i = 0; i < 12; ++i)
    CRNERF_REQUIRE(w->weight[i] && w->bias[i], "weight/bias pointer %d is null", i);
  Program prog;
  PackedLayout L;
  int rc = make_layout(w->e_xyz, w->e_dir, operand, &prog, &L);
  if (rc) return rc;
  CRNERF_REQUIRE(packed_bytes >= L.total, "packed buffer too small: %zu < %zu", packed_bytes, L.total);
  CRNERF_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 15) == 0, "packed buffer must be 16-byte aligned");
  PackParams P;
  memset(&P, 0, sizeof(P));
  for (int i = 0; i < 12; ++i) {
    P.w[i] = w->weight[i];
    P.b[i] = w->bias[i];
  }
  P.img = static_cast<uint8_t*>(packed);
  P.img_lo = L.img_lo ? P.img + L.img_lo : nullptr;
  P.blob = reinterpret_cast<float*>(P.img + L.blob);
  P.tab_out = reinterpret_cast<Tables*>(P.img + L.tab);
  P.status = status_dev;
  P.e_xyz = w->e_xyz;
  P.e_dir = w->e_dir;
  P.fmt = operand_fmt(operand);
  memcpy(P.tab.chunks, prog.chunks, sizeof(prog.chunks));
  memcpy(P.tab.units, prog.units, sizeof(prog.units));
  P.tab.n_chunks = prog.n_chunks;
  P.tab.n_units = prog.n_units;
  P.tab.image_bytes = prog.image_bytes;
  P.tab.n_images = operand_images(operand);
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
  const int need = 3 + p.n_chunks * 11 + p.n_units * 7;
  CRNERF_REQUIRE(cap >= need, "buffer too small: need %d ints", need);
  int k = 0;
  out[k++] = p.n_chunks;
  out[k++] = p.n_units;
  out[k++] = p.image_bytes;
  for (int i = 0; i < p.n_chunks; ++i) {
    const Chunk& c = p.chunks[i];
    const int v[11] = {c.offset, c.bytes, c.layer, c.rows, c.row0, c.wcol0, c.wcols, c.a_src, c.a_k0, c.nk, c.kind};
    for (int j = 0; j < 11; ++j) out[k++] = v[j];
  }
  for (int i = 0; i < p.n_units; ++i) {
    const Unit& u = p.units[i];
    const int v[7] = {u.layer, u.half, u.n, u.chunk0, u.nchunks, u.first_of_layer, u.last_of_layer};
    for (int j = 0; j < 7; ++j) out[k++] = v[j];
  }
  return need;
}

static int gcd_int(int a, int b) { return b ? gcd_int(b, a % b) : a; }

// work split of a ray batch: whole rays per CTA, a multiple of the rays that fill whole tiles
static int ray_grid(int n_rays, int S, long long* pts_per_cta) {
  const int sms = num_sms();
  const int m = 128 / gcd_int(S, 128);  // rays per whole number of tiles
  long long rpc = ((long long)n_rays + sms - 1) / sms;
  if (rpc < 1) rpc = 1;
  if (rpc >= 2 * m) rpc = (rpc + m - 1) / m * m;
  if (pts_per_cta) *pts_per_cta = rpc * S;
  return (int)((n_rays + rpc - 1) / rpc);
}

int render_partial_rows(int n_rays, int n_samples) {
  if (n_rays <= 0 || n_samples <= 0) return 0;
  return 2 * ray_grid(n_rays, n_samples, nullptr);
}

int launch_render(const RenderArgs& a, cudaStream_t st) {
  const int operand = a.operand, e_xyz = a.e_xyz, e_dir = a.e_dir;
  Program prog;
  PackedLayout L;
  int rc = make_layout(e_xyz, e_dir, operand, &prog, &L);
  if (rc) return rc;
  RenderParams P;
  memset(&P, 0, sizeof(P));
  P.wimg = static_cast<const uint8_t*>(a.packed);
  P.wimg_lo = L.img_lo ? P.wimg + L.img_lo : nullptr;
  P.blob = reinterpret_cast<const float*>(P.wimg + L.blob);
  P.tab = reinterpret_cast<const Tables*>(P.wimg + L.tab);
  P.jitter = a.jitter;
  P.chan_part = a.chan_part;
  P.overflow = a.overflow;
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
  P.dbg = g_dbg_layer >= 0 ? g_dbg_buf : nullptr;
  P.dbg_layer = g_dbg_layer;
  P.prof = g_dbg_layer <= -2 ? reinterpret_cast<long long*>(g_dbg_buf) : nullptr;
  P.exp = g_dbg_layer <= -2 ? (-2 - g_dbg_layer) : 0;  // -2: profile, -3: exp 1, -4: exp 2, -6: exp 4 (bit mask)
  P.n_points = a.n_points;
  P.S = a.n_samples > 0 ? a.n_samples : 1;
  P.n_freq_xyz = a.n_freq_xyz;
  P.n_freq_dir = a.n_freq_dir;
  P.e_xyz = e_xyz;
  P.e_dir = e_dir;
  P.n_chunks = prog.n_chunks;
  P.n_units = prog.n_units;
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
    grid = ray_grid(a.n_rays, a.n_samples, &P.pts_per_cta);
  }
  P.acts = static_cast<uint16_t*>(a.acts);
  P.raw_save = a.raw_save;
  const bool dbg = P.dbg != nullptr || P.prof != nullptr;
  const bool save = !dbg && (a.acts != nullptr || a.raw_save != nullptr);
  void (*kern)(const RenderParams);
  if (operand == 2) {
    CRNERF_REQUIRE(!dbg && !save, "the fp16x3 operand format is inference only (no activation dump / saved activations)");
    kern = render_fused_kernel<0, 0, true>;
  } else if (operand == 0) {
    kern = dbg ? render_fused_kernel<0, 1, false> : (save ? render_fused_kernel<0, 2, false> : render_fused_kernel<0, 0, false>);
  } else {
    kern = dbg ? render_fused_kernel<1, 1, false> : (save ? render_fused_kernel<1, 2, false> : render_fused_kernel<1, 0, false>);
  }
  CRNERF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  kern<<<grid, kThreads, kSmemBytes, st>>>(P);
  count_launch();
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

}  // namespace crnerf

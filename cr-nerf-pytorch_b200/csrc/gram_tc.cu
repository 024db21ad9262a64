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
//   D1 = X  W1^T  (N=128)  -> +b1, LeakyReLU -> H1, written IN PLACE over D1   tcgen05 SS
//   D2 = H1 W2^T  (N=64)   -> +b2, LeakyReLU -> H2, in place over D2            tcgen05 TS
//   D3 = H2 W3^T  (N=32)   -> +b3 -> Y; Y^T (32x128, pixels along K) -> smem    tcgen05 TS
//   G1 += Yh^T Yh, G2 += Yh^T Yl  (M=128 with only rows 0..31 meaningful, N=32, K=128)  SS
// G1 / G2 live in TMEM for the CTA's whole lifetime; G = G1 + G2 + G2^T is written once.
//
// Two tiles ("streams") are in flight per CTA, each with its own operand buffers and TMEM
// regions (2 x 224 columns + 64 for G1/G2 = all 512), so one stream's MMAs run under the other's
// epilogues; the 16 epilogue warps work through the two streams' phases alternately, and an
// issuer warp polls both streams' barriers and issues whichever MMA group is ready.  In-place activations are what makes two streams fit: a 32-column
// block of fp32 accumulators becomes 16 columns of packed fp16 hi words + 16 of lo words in the
// same columns (each thread rewrites its own lane).
//
// Precision: the result feeds a 1e-4 parity bar through two 1024x1024 FCs, and on trained
// weights the layers cancel (a CPU emulation with 16-bit activations misses the bar by 50-200x,
// see DESIGN.md), so every operand is split into fp16 hi + fp16 lo (x = hi + lo to ~2^-22) and
// every product issued as three MMAs (hi*hi + hi*lo + lo*hi, fp32 accumulate); the Gram uses its
// symmetry (Yh^T Yl + Yl^T Yh = G2 + G2^T) to get by with two.
//
// One launch serves up to two "jobs" (the content map with cnet, the style map with snet): CTAs
// past job 0's share work on job 1.  Each job's channel mean comes from per-block partial sums
// (the render kernel's epilogue or sums_rows_kernel), from an explicit vector (sharded path:
// 1 partial), or - small maps - is computed by the CTA itself.
#include <cuda_fp16.h>
#include <algorithm>
#include "common.h"
#include "ptx.cuh"
#include "gram_tc.h"

namespace crnerf {
namespace {

constexpr int kEpiWarps = 16;                       // lane quarter x column quarter
// + warps 16, 17: MMA issuers (16 also allocates TMEM); warps 18, 19 only donate registers: the register file is
// allocated in units of 4 warps, so 17 warps cost as much as 20, and setmaxnreg moves what the
// four control warps do not need (4 x 32 x 32) to the 16 epilogue warps (96 -> 104 each); the issuer
// keeps 64: its spills would go to L2 (the streaming loads thrash the small L1) and stall every MMA group
constexpr int kGThreads = (kEpiWarps + 4) * 32;
// shared memory map (bytes).  The Gram A operand is addressed as a 128-row tile although only
// 32 rows (channels) exist: rows 32..127 alias whatever follows (the other Y^T slabs, the X
// tile) and only feed accumulator lanes that are never read.
constexpr int kStreamBytes = 49152;   // per stream: Y^T hi slab0, hi slab1, lo slab0, lo slab1 (4 KB each) | X hi, X lo (16 KB each)
constexpr int kYtSlab = 4096;
constexpr int kXOffS = 16384;
constexpr int kW1Off = 2 * kStreamBytes;   // W1 hi, lo: 128 rows x 128 B
constexpr int kW2Off = kW1Off + 32768;     // W2 hi (2 slabs x 64 rows x 128 B), lo
constexpr int kW3Off = kW2Off + 32768;     // W3 hi, lo: 32 rows x 128 B
constexpr int kFOff = kW3Off + 8192;       // fp32: b1[128] b2[64] b3[32] mean[64]
constexpr int kRedOff = kFOff + 288 * 4;   // 8 x 64 floats (mean partials), reused as the 32 x 33 transpose tile
constexpr int kBarOff2 = kRedOff + 33 * 32 * 4;  // 16 mbarriers + tmem slot
constexpr int kGramSmem = kBarOff2 + 16 * 8 + 16;
// TMEM columns: stream s at 224*s: P (128: D1 -> H1 in place), Q (64: D2 -> H2 in place), R (32: D3)
constexpr uint32_t kStreamCols = 224, cP = 0, cQ = 128, cR = 192, cG1 = 448, cG2 = 480;
enum { X_FULL = 0, D1_FULL, A1_FULL, D2_FULL, A2_FULL, D3_FULL, YT_FULL, G_DONE, kBarsPerStream };

__device__ __forceinline__ float lrelu02(float v) { return fmaxf(v, 0.2f * v); }   // = v > 0 ? v : 0.2 v

// (a, b) -> packed fp16 hi pair and packed fp16 lo pair (a = hi + lo to ~2^-22 relative)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack2<0, false>(a, b);
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  lo = pack2<0, false>(a - f.x, b - f.y);
}

__device__ __forceinline__ uint32_t mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}

// Blocking wait with a suspend-time hint: the hardware parks the warp until the phase completes (or
// the hint expires) instead of letting it spin - a quarter of this kernel's issued instructions were
// try_wait spin loops competing with the epilogue arithmetic for the same schedulers.  Bounded like
// mbar_wait: a protocol bug traps instead of hanging.
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, uint32_t parity, uint32_t tag) {
  uint32_t spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
        : "memory");
    if (ok) return;
    if (++spins == (1u << 18)) {
      g_wait_timeout_tag = 0x80000000u | (tag << 16) | (blockIdx.x & 0xffff);
      __trap();
    }
  }
}

// fp32 row-major weight (rows x cols) -> SW128 K-major fp16 hi/lo slabs of 64 columns.  Each thread
// requests all its float4s first (independent loads, one L2 round trip), then splits and stores.
template <int kRows, int kCols>
__device__ __forceinline__ void stage_weight(const float* __restrict__ w, uint8_t* hi, uint8_t* lo) {
  constexpr int kQuads = kRows * kCols / 4, kIter = (kQuads + kGThreads - 1) / kGThreads, kSlab = kRows * 128;
  float4 v[kIter];
#pragma unroll
  for (int i = 0; i < kIter; ++i) {
    const int e = threadIdx.x + i * kGThreads;
    v[i] = e < kQuads ? __ldg(reinterpret_cast<const float4*>(w) + e) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int i = 0; i < kIter; ++i) {
    const int e = threadIdx.x + i * kGThreads;
    if (e >= kQuads) break;
    const int r = e / (kCols / 4), c = 4 * (e % (kCols / 4));
    uint32_t h0, l0, h1, l1;
    split2(v[i].x, v[i].y, h0, l0);
    split2(v[i].z, v[i].w, h1, l1);
    const uint32_t off = (uint32_t)(c >> 6) * kSlab + sw128_offset(r, (c & 63) >> 3) + (c & 7) * 2;
    *reinterpret_cast<uint2*>(hi + off) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(lo + off) = make_uint2(l0, l1);
  }
}

struct GramParams {
  GramJob job[2];
  int n_jobs;
  long long* ts;   // CRNERF_GRAM_TIMING builds: clock64 stamps of block 0 (tools/gram_timing.py)
};

#ifdef CRNERF_GRAM_TIMING
#define TSTAMP(slot)                                                                   \
  do {                                                                                 \
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && P.ts) P.ts[slot] = clock64();    \
  } while (0)
// per-warp event trace (warps 0 and 15 of block 0): (tag << 48) | clock, appended at ts[256 + 128*w + i]
#define TRACE(tag)                                                                                   \
  do {                                                                                               \
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && P.ts && (warp == 0 || warp == 15) && trace_n < 127) \
      P.ts[256 + 128 * (warp == 15) + trace_n++] = ((long long)(tag) << 48) | (clock64() & 0xffffffffffffLL); \
  } while (0)
// all-warp snapshot of one event in round 2: ts[800 + 20*k + warp]
#define SNAP(k, it)                                                                               \
  do {                                                                                            \
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && P.ts && (it) == 2) P.ts[800 + 20 * (k) + warp] = clock64(); \
  } while (0)
// fine-grained stamps of ONE untraced warp (5) in round 2: ts[900 + k]
#define FINE(k, it)                                                                               \
  do {                                                                                            \
    if (blockIdx.x == 0 && threadIdx.x == 5 * 32 && P.ts && (it) == 2) P.ts[900 + (k)] = clock64(); \
  } while (0)
#else
#define TSTAMP(slot) do { } while (0)
#define TRACE(tag) do { } while (0)
#define SNAP(k, it) do { } while (0)
#define FINE(k, it) do { } while (0)
#endif

// channel mean of the job's map -> mean[64] (shared), fixed summation order.  All threads call.
__device__ void job_mean(const GramJob& J, float* mean, float* red) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (J.self_mean && J.ch_stride != 1) {
    // planar / generic strides: a warp per channel, lanes along the pixels (coalesced)
    if (warp < kEpiWarps) {    // 16 warps x 4 channels (warp, +16, +32, +48), 4 x 4 loads in flight per lane
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const float* base = J.g + (long long)warp * J.ch_stride;
      long long p = lane;
      for (; p + 96 < J.n; p += 128) {
        float v[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int k = 0; k < 4; ++k) v[u][k] = __ldg(base + (p + 32 * u) * J.pix_stride + (long long)(16 * k) * J.ch_stride);
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[k] += v[u][k];
      }
      for (; p < J.n; p += 32)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] += __ldg(base + p * J.pix_stride + (long long)(16 * k) * J.ch_stride);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], d);
        if (lane == 0) mean[warp + 16 * k] = acc[k] * J.mean_scale;
      }
    }
    return;   // caller's __syncthreads publishes mean[]
  }
  // rows of 64 contiguous channels: the map itself (self_mean) or per-block partial sums
  const float* src = J.self_mean ? J.g : J.sum_parts;
  const long long rows = J.self_mean ? J.n : (long long)J.n_parts;
  const long long stride = J.self_mean ? J.pix_stride : 64;
  if (tid < 512) {
    const int c = tid & 63, rl = tid >> 6;
    float acc = 0.f;
#pragma unroll 8
    for (long long r = rl; r < rows; r += 8) acc += __ldg(src + r * stride + c);
    red[rl * 64 + c] = acc;
  }
  __syncthreads();
  if (tid < 64)
    mean[tid] = (((red[tid] + red[64 + tid]) + (red[128 + tid] + red[192 + tid])) +
                 ((red[256 + tid] + red[320 + tid]) + (red[384 + tid] + red[448 + tid]))) * J.mean_scale;
}

__global__ void __launch_bounds__(kGThreads, 1) gram_tc_kernel(const __grid_constant__ GramParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if (threadIdx.x == 0) TSTAMP(0);
  float* fblob = reinterpret_cast<float*>(smem + kFOff);
  float* red = reinterpret_cast<float*>(smem + kRedOff);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBarOff2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kBarsPerStream);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const GramJob& J = P.job[(P.n_jobs == 2 && (int)blockIdx.x >= P.job[1].first_block) ? 1 : 0];
  const int lb = (int)blockIdx.x - J.first_block;      // this CTA's index inside its job

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      uint64_t* b = bars + s * kBarsPerStream;
      mbar_init(&b[X_FULL], 16);
      mbar_init(&b[D1_FULL], 1);
      mbar_init(&b[A1_FULL], 16);
      mbar_init(&b[D2_FULL], 1);
      mbar_init(&b[A2_FULL], 16);
      mbar_init(&b[D3_FULL], 1);
      mbar_init(&b[YT_FULL], 16);
      mbar_init(&b[G_DONE], 1);
    }
    fence_mbar_init();
  }
  if (warp == kEpiWarps) tmem_alloc<512>(tmem_slot);
  stage_weight<128, 64>(J.w.conv_w[0], smem + kW1Off, smem + kW1Off + 16384);
  stage_weight<64, 128>(J.w.conv_w[1], smem + kW2Off, smem + kW2Off + 16384);
  stage_weight<32, 64>(J.w.conv_w[2], smem + kW3Off, smem + kW3Off + 4096);
  if (threadIdx.x == 0) TSTAMP(1);
  for (int i = tid; i < 224; i += kGThreads)
    fblob[i] = i < 128 ? J.w.conv_b[0][i] : (i < 192 ? J.w.conv_b[1][i - 128] : J.w.conv_b[2][i - 192]);
  job_mean(J, fblob + 224, red);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) TSTAMP(2);
  const long long n_tiles = (J.n + 127) / 128;
  const float* b1 = fblob, *b2 = fblob + 128, *b3 = fblob + 192, *mean = fblob + 224;
  if (lb == 0 && J.mean_out && tid < 64) J.mean_out[tid] = mean[tid];
  // CTA-local tile j (stream j & 1) -> tile of the map; `reverse` walks the map from its end, so a
  // pass that follows a forward pass over the same map starts on what is still in L2
  const int nb = J.n_blocks;
  auto tile_of = [&](long long j) -> long long {
    const long long t = lb + j * (long long)nb;
    return t < n_tiles ? (J.reverse ? n_tiles - 1 - t : t) : -1;
  };
  // number of tiles of stream s
  auto count_of = [&](int s) -> uint32_t {
    const long long mine = lb < n_tiles ? (n_tiles - 1 - lb) / nb + 1 : 0;   // tiles of this CTA
    return (uint32_t)((mine + 1 - s) / 2);
  };

  if (warp >= kEpiWarps) {
    setmaxnreg_dec<64>();     // the whole warpgroup (warps 16-19) must execute the same setmaxnreg
    // ------------------------------------------------------------------ MMA issuers
    // Two issuer warps, split by urgency rather than by stream: warp 16 issues the layer-2 / layer-3
    // groups (what the epilogue warps are waiting for), warp 17 the layer-1 prefetch of a stream's
    // next tile and the Gram groups (a whole tile of slack).  A burst blocks its issuer for its whole
    // execution time (the issue queue is shallow); with one issuer a short critical group queued up
    // behind the ISSUE of a long prefetch burst (measured: 1.2 - 2.2 k cycles from "activations
    // ready" to "group issued"), with two the pipe interleaves them.  All Gram MMAs come from one
    // thread, so the accumulators G1 / G2 see a single ordered stream; everything else is ordered
    // through the barriers.  Each warp runs its scheduler with warp-uniform state (barrier tests are
    // combined by a vote, so descriptors and counters stay in uniform registers: a thread-private
    // scheduler pays ~20 instructions of register shuffling per MMA) and one elected lane issues a
    // group as straight-line code.
    constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO 1024 | version | SW128
    auto desc = [](uint32_t saddr) {
      return (static_cast<uint64_t>(kHi) << 32) | (((saddr & 0x3ffffu) >> 4) | (1u << 16));
    };
    auto ready = [](uint64_t* bar, uint32_t parity) -> bool {
      return __all_sync(0xffffffffu, mbar_test(bar, parity) != 0);
    };
    const uint32_t s0 = smem_u32(smem);
    const uint32_t cnt[2] = {count_of(0), count_of(1)};
    [[maybe_unused]] int itr_n = 0;
#ifdef CRNERF_GRAM_TIMING
#define ITRACE(tag)                                                                                  \
  do {                                                                                               \
    if (blockIdx.x == 0 && lane == 0 && P.ts && itr_n < 100)                                         \
      P.ts[600 + 100 * (warp - kEpiWarps) + itr_n++] = ((long long)(tag) << 48) | (clock64() & 0xffffffffffffLL); \
  } while (0)
#else
#define ITRACE(tag) do { } while (0)
#endif
    if (warp == kEpiWarps) {
      uint32_t m_it[2] = {0, 0};     // tile of the chain
      int m_op[2] = {0, 0};          // 0: layer 2, 1: layer 3
      uint32_t idle = 0;
      while (m_it[0] < cnt[0] || m_it[1] < cnt[1]) {
        bool progressed = false;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          uint64_t* b = bars + s * kBarsPerStream;
          const uint32_t tb = tmem + s * kStreamCols;
          if (m_it[s] >= cnt[s]) continue;
          const uint32_t par = m_it[s] & 1;
          if (m_op[s] == 0) {
            // ---- layer 2: D2 = H1 W2^T, K = 128 (TS; H1 hi / lo words interleaved per 32-column block)
            if (!ready(&b[A1_FULL], par)) continue;
            tc_fence_after_sync();
            if (elect_one()) {
              const uint32_t id = make_idesc_f16(128, 64, 0);
#pragma unroll
              for (int term = 0; term < 3; ++term) {
                const uint32_t bq = s0 + kW2Off + (term == 1 ? 16384 : 0);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                  umma_ts(tb + cQ, tb + cP + 32 * (k >> 1) + 8 * (k & 1) + (term == 2 ? 16 : 0),
                          desc(bq + (k >> 2) * 8192 + 32 * (k & 3)), id, (term | k) ? 1u : 0u);
              }
              umma_commit(&b[D2_FULL]);
            }
            __syncwarp();
            ITRACE(40 + 4 * s + 0);
            m_op[s] = 1;
            progressed = true;
          } else {
            // ---- layer 3: D3 = H2 W3^T, K = 64 (TS; H2 per 16-column group: 8 hi words | 8 lo words)
            if (!ready(&b[A2_FULL], par)) continue;
            tc_fence_after_sync();
            if (elect_one()) {
              const uint32_t id = make_idesc_f16(128, 32, 0);
#pragma unroll
              for (int term = 0; term < 3; ++term) {
                const uint32_t bq = s0 + kW3Off + (term == 1 ? 4096 : 0);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_ts(tb + cR, tb + cQ + 16 * k + (term == 2 ? 8 : 0), desc(bq + 32 * k), id,
                          (term | k) ? 1u : 0u);
              }
              umma_commit(&b[D3_FULL]);
            }
            __syncwarp();
            ITRACE(40 + 4 * s + 1);
            m_op[s] = 0;
            ++m_it[s];
            progressed = true;
          }
        }
        if (progressed) {
          idle = 0;
        } else if (__nanosleep(20), ++idle == (1u << 24)) {   // protocol bug: fail the launch instead of hanging the box
          g_wait_timeout_tag = 0x80000000u | (69u << 16) | (blockIdx.x & 0xffff);
          __trap();
        }
      }
    } else if (warp == kEpiWarps + 1) {
      uint32_t l1_it[2] = {0, 0};    // next tile (stream-local) whose layer 1 is to be issued
      uint32_t g_it[2] = {0, 0};     // next tile whose Gram group is to be issued
      uint32_t g_acc = 0;            // 1 once the Gram accumulators hold something
      uint32_t idle = 0;
      while (g_it[0] < cnt[0] || g_it[1] < cnt[1]) {
        bool progressed = false;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          uint64_t* b = bars + s * kBarsPerStream;
          const uint32_t sb = s0 + s * kStreamBytes, tb = tmem + s * kStreamCols;
          // ---- layer 1 of tile l1_it: D1 = X W1^T, three split terms x 4 k-steps (SS).  Needs the
          // staged X and region P free (layer 2 of the previous tile, which reads H1 there, retired).
          if (l1_it[s] < cnt[s] && ready(&b[X_FULL], l1_it[s] & 1) &&
              (l1_it[s] == 0 || ready(&b[D2_FULL], (l1_it[s] - 1) & 1))) {
            tc_fence_after_sync();
            if (elect_one()) {
              const uint32_t id = make_idesc_f16(128, 128, 0);
#pragma unroll
              for (int term = 0; term < 3; ++term) {
                const uint32_t a = sb + kXOffS + (term == 2 ? 16384 : 0);   // hi, hi, lo
                const uint32_t bq = s0 + kW1Off + (term == 1 ? 16384 : 0);  // hi, lo, hi
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_ss(tb + cP, desc(a + 32 * k), desc(bq + 32 * k), id, (term | k) ? 1u : 0u);
              }
              umma_commit(&b[D1_FULL]);
            }
            __syncwarp();
            ITRACE(40 + 4 * s + 3);
            ++l1_it[s];
            progressed = true;
          }
          // ---- Gram: G1 += Yh^T Yh, G2 += Yh^T Yl, K = the tile's 128 pixels (SS; A rows 32..127 don't care)
          if (g_it[s] < cnt[s] && ready(&b[YT_FULL], g_it[s] & 1)) {
            tc_fence_after_sync();
            if (elect_one()) {
              const uint32_t id = make_idesc_f16(128, 32, 0);
              const uint32_t yh = sb, yl = sb + 2 * kYtSlab;
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const uint32_t o = (k >> 2) * kYtSlab + 32 * (k & 3);
                umma_ss(tmem + cG1, desc(yh + o), desc(yh + o), id, k ? 1u : g_acc);
              }
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const uint32_t o = (k >> 2) * kYtSlab + 32 * (k & 3);
                umma_ss(tmem + cG2, desc(yh + o), desc(yl + o), id, k ? 1u : g_acc);
              }
              umma_commit(&b[G_DONE]);
            }
            __syncwarp();
            ITRACE(40 + 4 * s + 2);
            g_acc = 1;
            ++g_it[s];
            progressed = true;
          }
        }
        if (progressed) {
          idle = 0;
        } else if (__nanosleep(20), ++idle == (1u << 24)) {
          g_wait_timeout_tag = 0x80000000u | (70u << 16) | (blockIdx.x & 0xffff);
          __trap();
        }
      }
    }
  } else {
    setmaxnreg_inc<104>();
    // ------------------------------------------------------------------ pixel rows
    // All 16 warps work on ONE phase of ONE stream at a time (TMEM lane quarter q = rows 32q..32q+31,
    // one per lane; column quarter cq) and alternate between the two streams phase by phase:
    //   ep1(0) ep1(1) ep2(0) ep2(1) ep3(0) ep3(1) | next pair of tiles ...
    // so every scheduler always has four warps on the same short phase while the other stream's
    // MMAs run on the tensor pipe (with eight private warps per stream, a stream waiting for its
    // MMAs idles half the SM and the other half runs at two warps per scheduler).
    const int q = warp & 3, cq = warp >> 2;
    const int row = 32 * q + lane;
    [[maybe_unused]] int trace_n = 0;
    // this thread's quarter of a pixel row (16 channels) of each stream's NEXT tile, requested one
    // tile ahead
    float4 xr[2][4];
    auto load_x = [&](int s, long long t) {
      const long long p = t * 128 + row;
      const bool valid = t >= 0 && p < J.n;
      if (valid && J.vec) {
        ldg_stream_v8(J.g + p * J.pix_stride + 16 * cq, xr[s][0], xr[s][1]);
        ldg_stream_v8(J.g + p * J.pix_stride + 16 * cq + 8, xr[s][2], xr[s][3]);
        return;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (valid) {
          const float* src = J.g + p * J.pix_stride + (long long)(16 * cq + 4 * i) * J.ch_stride;
          xr[s][i] = make_float4(__ldg(src), __ldg(src + J.ch_stride), __ldg(src + 2 * J.ch_stride),
                                 __ldg(src + 3 * J.ch_stride));
        } else {
          xr[s][i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    };
    // X row quarter (held in xr[s]) of tile tt -> smem (hi, lo), then signal the issuer
    auto stage_x = [&](int s, long long tt) {
      const bool valid = tt * 128 + row < J.n;
      uint8_t* xh = smem + s * kStreamBytes + kXOffS, *xl = xh + 16384;
#pragma unroll
      for (int c8 = 0; c8 < 2; ++c8) {
        const float4 q0 = xr[s][2 * c8], q1 = xr[s][2 * c8 + 1];
        const float v[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
        const int c0 = 16 * cq + 8 * c8;
        uint32_t h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float a = valid ? v[2 * j] - mean[c0 + 2 * j] : 0.f;
          const float bb = valid ? v[2 * j + 1] - mean[c0 + 2 * j + 1] : 0.f;
          split2(a, bb, h[j], l[j]);
        }
        const uint32_t off = sw128_offset(row, 2 * cq + c8);
        *reinterpret_cast<uint4*>(xh + off) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(xl + off) = make_uint4(l[0], l[1], l[2], l[3]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[s * kBarsPerStream + X_FULL]);
    };
    // ---- the three epilogues of a tile of stream s (it = the stream's tile counter)
    // epilogue 1: H1 = LeakyReLU(D1 + b1), fp16 hi / lo words written over the same 32 columns; then
    // (D1 drained, its MMAs retired: X smem is free) stage the stream's next tile and request the one after
    auto ep1 = [&](int s, uint32_t it, long long t_next, long long t_after) {
      uint64_t* b = bars + s * kBarsPerStream;
      const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16) + s * kStreamCols;
      TRACE(10 + s);
      if (s == 0) FINE(0, it);
      mbar_wait_parked(&b[D1_FULL], it & 1, 64);
      tc_fence_after_sync();
      if (s == 0) FINE(1, it);
      TRACE(12 + s);
      if (s == 0) SNAP(0, it);
      {
        uint32_t v[32], w[32];
        tmem_ld_x32(lane_base + cP + 32 * cq, v);
        tmem_ld_wait();
#pragma unroll
        for (int jj = 0; jj < 16; ++jj)
          split2(lrelu02(__uint_as_float(v[2 * jj]) + b1[32 * cq + 2 * jj]),
                 lrelu02(__uint_as_float(v[2 * jj + 1]) + b1[32 * cq + 2 * jj + 1]), w[jj], w[16 + jj]);
        tmem_st_x32(lane_base + cP + 32 * cq, w);
      }
      tmem_st_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&b[A1_FULL]);
      TRACE(14 + s);
      if (s == 0) SNAP(1, it);
      if (t_next >= 0) {
        stage_x(s, t_next);
        if (s == 0) FINE(5, it);
        load_x(s, t_after);
      }
      if (s == 0) FINE(6, it);
      TRACE(16 + s);
      if (s == 0) SNAP(2, it);
    };
    // epilogue 2: H2 = LeakyReLU(D2 + b2), in place: 16 columns -> 8 hi words | 8 lo words
    auto ep2 = [&](int s, uint32_t it) {
      uint64_t* b = bars + s * kBarsPerStream;
      const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16) + s * kStreamCols;
      mbar_wait_parked(&b[D2_FULL], it & 1, 65);
      tc_fence_after_sync();
      TRACE(20 + s);
      {
        uint32_t v[16], w[16];
        tmem_ld_x16(lane_base + cQ + 16 * cq, v);
        tmem_ld_wait();
#pragma unroll
        for (int jj = 0; jj < 8; ++jj)
          split2(lrelu02(__uint_as_float(v[2 * jj]) + b2[16 * cq + 2 * jj]),
                 lrelu02(__uint_as_float(v[2 * jj + 1]) + b2[16 * cq + 2 * jj + 1]), w[jj], w[8 + jj]);
        tmem_st_x16p(lane_base + cQ + 16 * cq, w);
      }
      tmem_st_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&b[A2_FULL]);
      TRACE(22 + s);
      if (s == 0) SNAP(3, it);
    };
    // epilogue 3: Y = D3 + b3 (no activation, linearStyleTransfer.py:15), zero for pixels beyond n;
    // Y^T hi / lo -> smem with the pixel index along K (this thread: 8 channels)
    auto ep3 = [&](int s, uint32_t it, long long tile) {
      uint64_t* b = bars + s * kBarsPerStream;
      const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16) + s * kStreamCols;
      const bool valid = tile * 128 + row < J.n;
      if (s == 0) FINE(10, it);
      mbar_wait_parked(&b[D3_FULL], it & 1, 66);
      tc_fence_after_sync();
      TRACE(30 + s);
      if (s == 0) FINE(11, it);
      uint32_t v[8];
      tmem_ld_x8(lane_base + cR + 8 * cq, v);
      tmem_ld_wait();
      if (s == 0) FINE(12, it);
      if (it > 0) mbar_wait_parked(&b[G_DONE], (it - 1) & 1, 67);  // the stream's previous Gram MMAs are done reading Y^T
      if (s == 0) FINE(13, it);
      uint8_t* yh = smem + s * kStreamBytes + (row >> 6) * kYtSlab;
      uint8_t* yl = yh + 2 * kYtSlab;
      const uint32_t kk = row & 63;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int c = 8 * cq + jj;
        const float y = valid ? __uint_as_float(v[jj]) + b3[c] : 0.f;
        const __half hi = __float2half_rn(y);
        const __half lo = __float2half_rn(y - __half2float(hi));
        const uint32_t off = sw128_offset(c, kk >> 3) + (kk & 7) * 2;
        *reinterpret_cast<__half*>(yh + off) = hi;
        *reinterpret_cast<__half*>(yl + off) = lo;
      }
      if (s == 0) FINE(14, it);
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&b[YT_FULL]);
      if (s == 0) FINE(15, it);
      TRACE(32 + s);
    };
    long long t0 = tile_of(0), t1 = tile_of(1);
    load_x(0, t0);
    if (t0 >= 0) {
      stage_x(0, t0);
      load_x(0, tile_of(2));
    }
    load_x(1, t1);
    if (t1 >= 0) {
      stage_x(1, t1);
      load_x(1, tile_of(3));
    }
    // The two streams run half a tile apart, so that each MMA group of one stream executes under an
    // epilogue of the other:   ep1(0) | ep3(1, previous tile) | ep2(0) | ep1(1) | ep3(0) | ep2(1)
    long long t1_prev = -1;
    for (uint32_t r = 0; t0 >= 0 || t1_prev >= 0; ++r) {
      const long long j = 2 * (long long)r;
      const long long t0_next = tile_of(j + 2), t1_next = tile_of(j + 3);
      if (t0 >= 0) ep1(0, r, t0_next, tile_of(j + 4));
      if (t1_prev >= 0) ep3(1, r - 1, t1_prev);
      if (t0 >= 0) ep2(0, r);
      if (t1 >= 0) ep1(1, r, t1_next, tile_of(j + 5));
      if (t0 >= 0) ep3(0, r, t0);
      if (t1 >= 0) ep2(1, r);
      t1_prev = t1;
      t0 = t0_next;
      t1 = t1_next;
    }
    // ---- G = G1 + G2 + G2^T (rows 0..31) -> this block's partial
    if (warp == 0) {
      const uint32_t c0 = count_of(0), c1 = count_of(1);
      if (c0) mbar_wait(&bars[G_DONE], (c0 - 1) & 1, 68);
      if (c1) mbar_wait(&bars[kBarsPerStream + G_DONE], (c1 - 1) & 1, 68);
      tc_fence_after_sync();
      TSTAMP(3);
      uint32_t g1[32], g2[32];
      if (c0 + c1) {
        tmem_ld_x32(tmem + cG1, g1);
        tmem_ld_x32(tmem + cG2, g2);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) g1[jj] = g2[jj] = 0u;
      }
#pragma unroll
      for (int jj = 0; jj < 32; ++jj) red[lane * 33 + jj] = __uint_as_float(g2[jj]);
      __syncwarp();
#pragma unroll
      for (int jj = 0; jj < 32; ++jj)
        J.partial[(long long)lb * 1024 + lane * 32 + jj] =
            (__uint_as_float(g1[jj]) + __uint_as_float(g2[jj])) + red[jj * 33 + lane];
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x == 0) TSTAMP(4);
  if (warp == kEpiWarps) tmem_dealloc<512>(tmem);
}

}  // namespace

int gram_blocks(int64_t n, int max_blocks) {
  const long long tiles = (n + 127) / 128;
  return (int)std::max<long long>(1, std::min<long long>(tiles, std::min(max_blocks, num_sms())));
}

// One launch for up to two jobs.  job[i].first_block / n_blocks are filled here: when both jobs
// together need more CTAs than there are SMs, job 1 (the small one: the style map) keeps its tile
// count and job 0 gets the remaining SMs.
int gram_tc_launch(GramJob* jobs, int n_jobs, int max_blocks, cudaStream_t st) {
  CRNERF_REQUIRE(n_jobs == 1 || n_jobs == 2, "one or two Gram jobs per launch");
  GramParams P;
  P.n_jobs = n_jobs;
#ifdef CRNERF_GRAM_TIMING
  P.ts = reinterpret_cast<long long*>(g_dbg_buf);
#else
  P.ts = nullptr;
#endif
  const int sms = num_sms();
  int want[2] = {0, 0};
  for (int i = 0; i < n_jobs; ++i) {
    CRNERF_REQUIRE(jobs[i].g && jobs[i].partial && jobs[i].n >= 1, "bad Gram job");
    CRNERF_REQUIRE(jobs[i].self_mean || (jobs[i].sum_parts && jobs[i].n_parts >= 1), "Gram job has no mean source");
    want[i] = gram_blocks(jobs[i].n, max_blocks);
  }
  if (n_jobs == 2 && want[0] + want[1] > sms) want[0] = std::max(1, sms - want[1]);
  int first = 0;
  for (int i = 0; i < n_jobs; ++i) {
    jobs[i].first_block = first;
    jobs[i].n_blocks = want[i];
    jobs[i].vec = jobs[i].ch_stride == 1 && (jobs[i].pix_stride & 3) == 0 &&
                  (reinterpret_cast<uintptr_t>(jobs[i].g) & 15) == 0;
    first += want[i];
    P.job[i] = jobs[i];
  }
  CRNERF_CUDA(cudaFuncSetAttribute(gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGramSmem));
  gram_tc_kernel<<<first, kGThreads, kGramSmem, st>>>(P);
  count_launch();
  CRNERF_CUDA(cudaGetLastError());
  return CRNERF_OK;
}

}  // namespace crnerf

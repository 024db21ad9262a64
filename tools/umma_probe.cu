// umma_probe: validates, on a real B200, every tcgen05 / bulk-copy primitive
// the fused render kernel relies on, with the SAME wrappers (csrc/ptx.cuh):
//   * K-major SWIZZLE_128B smem descriptors, K-advance inside a slab, slab stride
//   * SS MMA (A,B from smem) and TS MMA (A from TMEM), mixed into one accumulator
//   * N = 64 / 128 / 256, D at a column offset, accumulate flag
//   * tcgen05.st/ld 32x32b, pack2 ordering, FADD2
//   * MMA issue rate (cycles per instruction) for SS vs TS
// Inputs are small integers so fp16 x fp16 -> fp32 sums are exact and the
// comparison with the host reference is bit-exact.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I cr-nerf-pytorch_b200/csrc
//        tools/umma_probe.cu -o tools/umma_probe
// run  : tools/umma_probe <mode 0=SS 1=TS 2=mixed> <N> <K> <fmt 0=f16 1=bf16> [reps]
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "ptx.cuh"

using namespace crnerf;

#define CK(x)                                                                     \
  do {                                                                            \
    cudaError_t e = (x);                                                          \
    if (e != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(2);                                                                    \
    }                                                                             \
  } while (0)

struct Params {
  const uint8_t* a_img;   // swizzled A slabs: [K/64][128 rows][128 B]
  const uint8_t* b_img;   // swizzled B slabs: [K/64][N rows][128 B]
  const uint16_t* a_rm;   // row-major A (128 x K) 16-bit
  float* d_out;           // 128 x N fp32 row-major
  long long* cycles;      // [0] = cycles for the timed MMA loop
  int mode, N, K, fmt, reps, dcol;
  int grid;
  int commit_every;  // extra tcgen05.commit to a scratch mbarrier every n MMAs (0 = never)
  int alt;           // alternate between two accumulators every 16 MMAs
  int acol;          // first TMEM column of the A operand (TS modes)
};

__global__ void __launch_bounds__(672, 1) probe_kernel(Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // layout: A slabs (K/64 x 16 KB) | B slabs (K/64 x N*128 B) | barriers
  const int kslabs = p.K / 64;
  uint8_t* sA = smem;
  uint8_t* sB = smem + kslabs * 16384;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + kslabs * p.N * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);  // bulk copy landed
    mbar_init(&bars[1], 1);  // MMA done
    mbar_init(&bars[2], 1);  // scratch target of the extra commits
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc<512>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t a_tmem = tmem + p.acol;  // A operand columns (K/2 <= 128 cols)
  const uint32_t d_tmem = tmem + p.dcol;

  if (threadIdx.x == 0) {
    uint32_t bytes = kslabs * 16384 + kslabs * p.N * 128;
    mbar_arrive_expect_tx(&bars[0], bytes);
    bulk_g2s(sA, p.a_img, kslabs * 16384, &bars[0]);
    bulk_g2s(sB, p.b_img, kslabs * p.N * 128, &bars[0]);
  }
  // TS / mixed: stage A into TMEM through registers (row = lane of warp w)
  if (warp < 4 && p.mode != 0) {
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < p.K; c0 += 32) {
      uint32_t v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        uint16_t lo = p.a_rm[row * p.K + c0 + 2 * j];
        uint16_t hi = p.a_rm[row * p.K + c0 + 2 * j + 1];
        v[j] = (static_cast<uint32_t>(hi) << 16) | lo;
      }
      tmem_st_x16(a_tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0 / 2, v);
    }
    tmem_st_wait();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();

  if (warp == 4) {
    mbar_wait(&bars[0], 0, 1);
    tc_fence_after_sync();
    if (elect_one()) {
      const uint32_t idesc = make_idesc_f16(128, p.N, p.fmt);
      const uint32_t sbo = 1024;
      long long t0 = clock64();
      int issued = 0;
      for (int r = 0; r < p.reps; ++r) {
        const uint32_t d_tmem = (p.alt && (r & 1)) ? tmem + p.dcol + 128 : tmem + p.dcol;
        for (int ks = 0; ks < p.K / 16; ++ks) {
          const int slab = ks / 4, kin = ks % 4;
          uint64_t bdesc = make_sdesc_k_sw128(smem_u32(sB + slab * p.N * 128) + kin * 32, sbo);
          const uint32_t acc = (ks > 0) ? 1u : 0u;
          bool use_ts = (p.mode == 1) || (p.mode == 2 && (ks & 1));
          if (use_ts) {
            umma_ts(d_tmem, a_tmem + ks * 8, bdesc, idesc, acc);
          } else {
            uint64_t adesc = make_sdesc_k_sw128(smem_u32(sA + slab * 16384) + kin * 32, sbo);
            umma_ss(d_tmem, adesc, bdesc, idesc, acc);
          }
          if (++issued == p.commit_every) { umma_commit(&bars[2]); issued = 0; }
          if (blockIdx.x == 0 && r < 3) p.cycles[256 + r * 16 + ks] = clock64() - t0;
        }
      }
      umma_commit(&bars[1]);
      mbar_wait(&bars[1], 0, 2);
      long long t1 = clock64();
      p.cycles[blockIdx.x] = t1 - t0;
    }
    __syncwarp();
  }
  if (warp < 4) {
    mbar_wait(&bars[1], 0, 3);
    tc_fence_after_sync();
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < p.N; c0 += 32) {
      uint32_t v[32];
      tmem_ld_x32(d_tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
      tmem_ld_wait();
      if (blockIdx.x == 0) {
#pragma unroll
        for (int j = 0; j < 32; ++j) p.d_out[row * p.N + c0 + j] = __uint_as_float(v[j]);
      }
    }
  }
  if (warp > 4) {  // extra pollers: spin on the "MMA done" barrier like the product kernel's epilogue warps
    mbar_wait(&bars[1], 0, 4);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc<512>(tmem);
}

// pack2 / add2 semantics
__global__ void misc_kernel(float* out) {
  uint32_t r = pack2<0, true>(1.5f, -2.0f);  // lo=1.5, hi=-2 -> relu -> hi=0
  out[0] = __half2float(__ushort_as_half(static_cast<unsigned short>(r & 0xffff)));
  out[1] = __half2float(__ushort_as_half(static_cast<unsigned short>(r >> 16)));
  uint32_t s = pack2<0, false>(70000.f, -3.25f);  // satfinite: lo -> 65504
  out[2] = __half2float(__ushort_as_half(static_cast<unsigned short>(s & 0xffff)));
  out[3] = __half2float(__ushort_as_half(static_cast<unsigned short>(s >> 16)));
  float2 c = add2(make_float2(1.f, 2.f), make_float2(10.f, 20.f));
  out[4] = c.x;
  out[5] = c.y;
  uint32_t b = pack2<1, true>(0.3f, 3.0f);
  out[6] = __bfloat162float(__ushort_as_bfloat16(static_cast<unsigned short>(b & 0xffff)));
  out[7] = __bfloat162float(__ushort_as_bfloat16(static_cast<unsigned short>(b >> 16)));
}

static uint16_t to16(float v, int fmt) {
  if (fmt == 0) {
    __half h = __float2half(v);
    uint16_t u;
    memcpy(&u, &h, 2);
    return u;
  }
  __nv_bfloat16 h = __float2bfloat16(v);
  uint16_t u;
  memcpy(&u, &h, 2);
  return u;
}

int main(int argc, char** argv) {
  int mode = argc > 1 ? atoi(argv[1]) : 0;
  int N = argc > 2 ? atoi(argv[2]) : 128;
  int K = argc > 3 ? atoi(argv[3]) : 128;
  int fmt = argc > 4 ? atoi(argv[4]) : 0;
  int reps = argc > 5 ? atoi(argv[5]) : 1;
  int dcol = argc > 6 ? atoi(argv[6]) : 0;
  int grid = argc > 7 ? atoi(argv[7]) : 1;
  int commit_every = argc > 8 ? atoi(argv[8]) : 0;
  int alt = argc > 9 ? atoi(argv[9]) : 0;
  int acol = argc > 10 ? atoi(argv[10]) : 384;
  const int M = 128;
  if (K % 64 || K > 256 || N % 16 || N > 256 || dcol + N > 512 || acol + K / 2 > 512 ||
      (acol < dcol + N * (alt ? 2 : 1) && dcol < acol + K / 2)) {
    printf("bad args\n");
    return 2;
  }
  srand(1234 + N + K + mode);
  std::vector<float> A(M * K), B(N * K);
  const int rnd = argc > 11 ? atoi(argv[11]) : 0;  // 1: dense random mantissas (power test; compare is loose)
  for (auto& v : A) v = rnd ? float(rand() % 20001 - 10000) * 1.37e-4f : float(rand() % 9 - 4);
  for (auto& v : B) v = rnd ? float(rand() % 20001 - 10000) * 0.93e-5f : float(rand() % 9 - 4) * 0.125f;
  std::vector<uint16_t> a_rm(M * K);
  for (int i = 0; i < M * K; ++i) a_rm[i] = to16(A[i], fmt);
  const int kslabs = K / 64;
  std::vector<uint8_t> a_img(kslabs * 16384, 0), b_img(kslabs * N * 128, 0);
  for (int s = 0; s < kslabs; ++s) {
    for (int r = 0; r < M; ++r)
      for (int k = 0; k < 64; ++k) {
        uint16_t v = to16(A[r * K + s * 64 + k], fmt);
        uint32_t off = sw128_offset(r, k / 8) + (k % 8) * 2;
        memcpy(&a_img[s * 16384 + off], &v, 2);
      }
    for (int r = 0; r < N; ++r)
      for (int k = 0; k < 64; ++k) {
        uint16_t v = to16(B[r * K + s * 64 + k], fmt);
        uint32_t off = sw128_offset(r, k / 8) + (k % 8) * 2;
        memcpy(&b_img[s * N * 128 + off], &v, 2);
      }
  }
  std::vector<float> ref(M * N, 0.f);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float acc = 0;
      for (int k = 0; k < K; ++k) acc += A[m * K + k] * B[n * K + k];
      ref[m * N + n] = acc;
    }

  Params p{};
  uint8_t *da, *db;
  uint16_t* darm;
  float* dd;
  long long* dc;
  CK(cudaMalloc(&da, a_img.size()));
  CK(cudaMalloc(&db, b_img.size()));
  CK(cudaMalloc(&darm, a_rm.size() * 2));
  CK(cudaMalloc(&dd, M * N * 4));
  CK(cudaMalloc(&dc, 8 * 1024));
  CK(cudaMemcpy(da, a_img.data(), a_img.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, b_img.data(), b_img.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(darm, a_rm.data(), a_rm.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dd, 0xff, M * N * 4));
  p.a_img = da;
  p.b_img = db;
  p.a_rm = darm;
  p.d_out = dd;
  p.cycles = dc;
  p.mode = mode;
  p.N = N;
  p.K = K;
  p.fmt = fmt;
  p.reps = reps;
  p.dcol = dcol;
  p.grid = grid;
  p.commit_every = commit_every;
  p.alt = alt;
  p.acol = acol;
  size_t smem = kslabs * 16384 + kslabs * N * 128 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int pollers = argc > 12 ? atoi(argv[12]) : 0;  // extra warps spinning on an mbarrier
  probe_kernel<<<grid, 160 + 32 * pollers, smem>>>(p);
  CK(cudaGetLastError());
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    unsigned int tag = 0;
    printf("probe mode=%d N=%d K=%d fmt=%d: LAUNCH FAILED: %s\n", mode, N, K, fmt,
           cudaGetErrorString(e));
    (void)tag;
    return 1;
  }
  std::vector<float> out(M * N);
  long long cyc = 0;
  CK(cudaMemcpy(out.data(), dd, M * N * 4, cudaMemcpyDeviceToHost));
  {
    std::vector<long long> cy(grid);
    CK(cudaMemcpy(cy.data(), dc, 8 * grid, cudaMemcpyDeviceToHost));
    for (auto v : cy) cyc = v > cyc ? v : cyc;   // slowest CTA
  }
  double maxerr = 0;
  int bad = 0;
  for (int i = 0; i < M * N; ++i) {
    double d = fabs((double)out[i] - (double)ref[i]);
    if (!(d <= maxerr)) maxerr = d;
    if (!(d == 0)) bad++;
  }
  int nmma = reps * (K / 16);
  if (reps >= 3 && K == 256) {
    std::vector<long long> ts(48);
    CK(cudaMemcpy(ts.data(), dc + 256, 8 * 48, cudaMemcpyDeviceToHost));
    printf("issue timestamps (cycles since start) of the first 48 MMAs:");
    for (int i = 0; i < 48; ++i) printf(" %lld", ts[i]);
    printf("\n");
  }
  printf("probe acol=%d commit_every=%d alt=%d grid=%d mode=%d N=%d K=%d fmt=%d dcol=%d reps=%d: max_abs_err=%g mismatches=%d/%d  cycles=%lld (%.1f per MMA)  %s\n",
         acol, commit_every, alt, grid, mode, N, K, fmt, dcol, reps, maxerr, bad, M * N, cyc, double(cyc) / nmma,
         bad == 0 ? "OK" : "FAIL");
  if (bad && bad < M * N) {
    int shown = 0;
    for (int i = 0; i < M * N && shown < 8; ++i)
      if (out[i] != ref[i]) {
        printf("   [%d,%d] got %g want %g\n", i / N, i % N, out[i], ref[i]);
        shown++;
      }
  }
  if (mode == 0 && N == 128 && K == 128 && reps == 1 && fmt == 0) {
    float* dm;
    CK(cudaMalloc(&dm, 64));
    misc_kernel<<<1, 1>>>(dm);
    float hm[8];
    CK(cudaMemcpy(hm, dm, 32, cudaMemcpyDeviceToHost));
    printf("misc: pack2<f16,relu>(lo=1.5,hi=-2) -> lo=%g hi=%g (want 1.5, 0) | sat(lo=70000,hi=-3.25) -> "
           "lo=%g hi=%g (want 65504,-3.25) | add2 -> %g %g (want 11 22) | bf16 relu(0.3,3) -> %g %g\n",
           hm[0], hm[1], hm[2], hm[3], hm[4], hm[5], hm[6], hm[7]);
  }
  return bad ? 1 : 0;
}

// mn_nosw_probe: validates the MN-major SWIZZLE_NONE tcgen05 shared-memory operand form the
// encoder's weight-gradient kernel (csrc/encoder.cu, enc_wgrad_tc_kernel) relies on.
//
// The encoder's activation / gradient "planes" are [C/8 chunks][pixels][8 channels] 16-bit: one
// 16-byte element = 8 channels of one pixel, consecutive pixels of a chunk 16 B apart.  Read as an
// MN-major operand (M/N = channels, K = pixels) that is the canonical no-swizzle layout
//   ((8 ch, m chunks), (8 px, k groups)) : ((2 B, SBO), (16 B, LBO)),  SBO = chunk stride, LBO = 128 B
// (CUTLASS cute/atom/mma_traits_sm100.hpp, make_umma_desc<Major::MN>, LayoutType::INTERLEAVE),
// and a 3x3 tap is the descriptor start address shifted by whole pixels (16 B each).
// The probe computes D[m][n] = sum_p A[p + sa][m] * B[p + sb][n] over K = 64 pixels with both
// operands MN-major (M = 128 channels of A, N channels of B) and compares bit-exactly with the
// host (small-integer inputs).  variant 1 swaps LBO / SBO so a wrong reading shows in one run.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I cr-nerf-pytorch_b200/csrc
//        tools/mn_nosw_probe.cu -o tools/mn_nosw_probe
// run  : tools/mn_nosw_probe [N=128] [variant=0] [shift_a=0] [shift_b=0] [reps: time 16 x reps MMAs]
#include <cuda_fp16.h>
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

constexpr int kPix = 64, kRun = 80;   // K pixels per product, pixels staged per chunk (room for shifts)

__device__ __forceinline__ uint64_t desc_mn_nosw(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3ffff) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3fff) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3fff) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

__global__ void __launch_bounds__(160, 1)
probe_kernel(const uint8_t* a_img, const uint8_t* b_img, float* d_out, int N, int variant, int sa, int sb, int reps,
             long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t a_bytes = 16 * kRun * 16, b_bytes = (N / 8) * kRun * 16;
  uint8_t* sA = smem;
  uint8_t* sB = smem + a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + b_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc<512>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bars[0], a_bytes + b_bytes);
    bulk_g2s(sA, a_img, a_bytes, &bars[0]);
    bulk_g2s(sB, b_img, b_bytes, &bars[0]);
  }
  if (warp == 4) {
    mbar_wait(&bars[0], 0, 1);
    tc_fence_after_sync();
    if (elect_one()) {
      const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
      uint32_t lbo = 128, sbo = kRun * 16;
      if (variant == 1) { lbo = kRun * 16; sbo = 128; }
      for (int ks = 0; ks < kPix / 16; ++ks) {
        const uint64_t ad = desc_mn_nosw(smem_u32(sA) + (ks * 16 + sa) * 16, lbo, sbo);
        const uint64_t bd = desc_mn_nosw(smem_u32(sB) + (ks * 16 + sb) * 16, lbo, sbo);
        umma_ss(tmem, ad, bd, idesc, ks ? 1u : 0u);
      }
      umma_commit(&bars[1]);
      // issue rate of the MN-major no-swizzle form (results discarded): `reps` x 16 MMAs back to back
      if (reps > 0) {
        mbar_wait(&bars[1], 0, 5);
        const uint64_t ad = desc_mn_nosw(smem_u32(sA), lbo, sbo), bd = desc_mn_nosw(smem_u32(sB), lbo, sbo);
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
          for (int j = 0; j < 16; ++j) umma_ss(tmem + 128, ad + (uint64_t)((j & 3) * 16), bd + (uint64_t)((j & 3) * 16 + (j >> 2)), idesc, 1u);
        }
        umma_commit(&bars[2]);
        mbar_wait(&bars[2], 0, 6);
        cycles[0] = clock64() - t0;
      }
    }
    __syncwarp();
  }
  if (warp < 4) {
    mbar_wait(&bars[1], 0, 3);
    tc_fence_after_sync();
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t v[32];
      tmem_ld_x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) d_out[row * N + c0 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc<512>(tmem);
}

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 128;
  const int variant = argc > 2 ? atoi(argv[2]) : 0;
  const int sa = argc > 3 ? atoi(argv[3]) : 0, sb = argc > 4 ? atoi(argv[4]) : 0;
  const int reps = argc > 5 ? atoi(argv[5]) : 0;
  const int M = 128;
  if (N % 16 || N > 256 || sa < 0 || sb < 0 || sa + kPix > kRun || sb + kPix > kRun) {
    printf("bad args\n");
    return 2;
  }
  srand(91 + N);
  std::vector<float> A(kRun * M), B(kRun * N);   // [pixel][channel]
  for (auto& v : A) v = float(rand() % 9 - 4);
  for (auto& v : B) v = float(rand() % 9 - 4) * 0.125f;
  auto image = [&](const std::vector<float>& X, int C) {
    std::vector<uint8_t> img((C / 8) * kRun * 16, 0);
    for (int c = 0; c < C; ++c)
      for (int p = 0; p < kRun; ++p) {
        const __half h = __float2half(X[p * C + c]);
        memcpy(&img[((c / 8) * kRun + p) * 16 + (c % 8) * 2], &h, 2);
      }
    return img;
  };
  auto a_img = image(A, M), b_img = image(B, N);
  std::vector<float> ref(M * N, 0.f);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float acc = 0;
      for (int p = 0; p < kPix; ++p) acc += A[(p + sa) * M + m] * B[(p + sb) * N + n];
      ref[m * N + n] = acc;
    }
  uint8_t *da, *db;
  float* dd;
  CK(cudaMalloc(&da, a_img.size()));
  CK(cudaMalloc(&db, b_img.size()));
  CK(cudaMalloc(&dd, M * N * 4));
  CK(cudaMemcpy(da, a_img.data(), a_img.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, b_img.data(), b_img.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dd, 0xff, M * N * 4));
  const size_t smem = a_img.size() + b_img.size() + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long* dcyc;
  CK(cudaMalloc(&dcyc, 8));
  CK(cudaMemset(dcyc, 0, 8));
  probe_kernel<<<1, 160, smem>>>(da, db, dd, N, variant, sa, sb, reps, dcyc);
  CK(cudaGetLastError());
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("mn_nosw_probe N=%d variant=%d: LAUNCH FAILED: %s\n", N, variant, cudaGetErrorString(e));
    return 1;
  }
  std::vector<float> out(M * N);
  CK(cudaMemcpy(out.data(), dd, M * N * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int i = 0; i < M * N; ++i) bad += out[i] != ref[i];
  printf("mn_nosw_probe N=%d variant=%d shift_a=%d shift_b=%d: mismatches=%d/%d %s\n", N, variant, sa, sb, bad, M * N,
         bad == 0 ? "OK" : "FAIL");
  if (reps > 0) {
    long long cyc = 0;
    CK(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost));
    printf("   MN-major no-swizzle SS MMA M=128 N=%d K=16: %.1f cycles per MMA (%d back to back)\n", N,
           double(cyc) / (16.0 * reps), 16 * reps);
  }
  if (bad) {
    int shown = 0;
    for (int i = 0; i < M * N && shown < 6; ++i)
      if (out[i] != ref[i]) {
        printf("   [%d,%d] got %g want %g\n", i / N, i % N, out[i], ref[i]);
        shown++;
      }
  }
  return bad ? 1 : 0;
}

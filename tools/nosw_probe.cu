// nosw_probe: validates the SWIZZLE_NONE K-major shared-memory descriptor the encoder's
// implicit-GEMM convolution relies on: the A operand lives in "channel-chunk planes"
// [K/8][rows][16 B], so every row of a K-chunk is 16 B after the previous one and a 3x3 tap is
// just a row shift of the descriptor start address.  B stays in the SWIZZLE_128B form the
// other kernels use.  Integer inputs -> the fp32 result is exact and compared bit for bit.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I cr-nerf-pytorch_b200/csrc
//        tools/nosw_probe.cu -o tools/nosw_probe
// run  : tools/nosw_probe <shift rows> <plane rows> <variant 0: LBO=plane stride, SBO=128  1: swapped> [reps: time SS MMAs]
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "ptx.cuh"

using namespace crnerf;

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e = (x);                                                             \
    if (e != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(2);                                                                       \
    }                                                                                \
  } while (0)

__device__ __forceinline__ uint64_t make_sdesc_k_nosw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3ffff) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;  // layout type 0 = no swizzle
}

constexpr int K = 64, N = 128;

__global__ void __launch_bounds__(160, 1)
probe(const uint16_t* a_planes, int plane_rows, const uint8_t* b_img, int shift, int variant, float* d_out,
      long long* cycles, int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sB = smem;                      // N x 128 B, swizzled
  uint8_t* sA = smem + N * 128;            // [K/8][plane_rows][16 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + (K / 8) * plane_rows * 16);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc<128>(slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t bytes = N * 128 + (K / 8) * plane_rows * 16;
    mbar_arrive_expect_tx(&bars[0], bytes);
    bulk_g2s(sB, b_img, N * 128, &bars[0]);
    bulk_g2s(sA, a_planes, (K / 8) * plane_rows * 16, &bars[0]);
  }
  if (warp == 4) {
    mbar_wait(&bars[0], 0, 1);
    tc_fence_after_sync();
    if (elect_one()) {
      const uint32_t idesc = make_idesc_f16(128, N, 0);
      const uint32_t plane = plane_rows * 16;
      for (int ks = 0; ks < K / 16; ++ks) {
        const uint32_t a_addr = smem_u32(sA) + (2 * ks) * plane + shift * 16;
        const uint64_t adesc = variant == 0 ? make_sdesc_k_nosw(a_addr, plane, 128) : make_sdesc_k_nosw(a_addr, 128, plane);
        const uint64_t bdesc = make_sdesc_k_sw128(smem_u32(sB) + ks * 32, 1024);
        umma_ss(tmem, adesc, bdesc, idesc, ks > 0);
      }
      umma_commit(&bars[1]);
      // issue-rate comparison (results discarded): N = 128 and N = 64, A no-swizzle planes vs A SWIZZLE_128B
      if (reps > 0) {
        mbar_wait(&bars[1], 0, 5);
        uint32_t par = 1;
        for (int mode = 0; mode < 4; ++mode) {
          const uint32_t n = (mode & 1) ? 64 : 128;
          const uint32_t id = make_idesc_f16(128, n, 0);
          const long long t0 = clock64();
          for (int r = 0; r < reps; ++r)
            for (int ks = 0; ks < K / 16; ++ks) {
              const uint64_t adesc = (mode & 2) ? make_sdesc_k_sw128(smem_u32(sA) + ks * 32, 1024)
                                                : make_sdesc_k_nosw(smem_u32(sA) + (2 * ks) * plane, plane, 128);
              umma_ss(tmem, adesc, make_sdesc_k_sw128(smem_u32(sB) + ks * 32, 1024), id, 1);
            }
          umma_commit(&bars[1]);
          mbar_wait(&bars[1], par, 6);
          par ^= 1;
          cycles[mode] = clock64() - t0;
        }
      }
    }
    __syncwarp();
  }
  if (warp < 4) {
    mbar_wait(&bars[1], 0, 2);   // (the timing rounds later overwrite D; d_out is read right here)
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
  if (warp == 4) tmem_dealloc<128>(tmem);
}

static uint16_t h16(float v) {
  __half h = __float2half(v);
  uint16_t u;
  memcpy(&u, &h, 2);
  return u;
}

int main(int argc, char** argv) {
  const int shift = argc > 1 ? atoi(argv[1]) : 0;
  const int plane_rows = argc > 2 ? atoi(argv[2]) : 390;
  const int variant = argc > 3 ? atoi(argv[3]) : 0;
  if (shift + 128 > plane_rows) {
    printf("bad args\n");
    return 2;
  }
  srand(7 + shift);
  std::vector<float> A(plane_rows * K), B(N * K);
  for (auto& v : A) v = float(rand() % 9 - 4);
  for (auto& v : B) v = float(rand() % 9 - 4) * 0.125f;
  std::vector<uint16_t> planes((K / 8) * plane_rows * 8);
  for (int r = 0; r < plane_rows; ++r)
    for (int k = 0; k < K; ++k) planes[((k / 8) * plane_rows + r) * 8 + (k % 8)] = h16(A[r * K + k]);
  std::vector<uint8_t> bimg(N * 128);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      uint16_t u = h16(B[n * K + k]);
      memcpy(&bimg[sw128_offset(n, k / 8) + (k % 8) * 2], &u, 2);
    }
  uint16_t* d_planes;
  uint8_t* d_b;
  float* d_out;
  CK(cudaMalloc(&d_planes, planes.size() * 2));
  CK(cudaMalloc(&d_b, bimg.size()));
  CK(cudaMalloc(&d_out, 128 * N * 4));
  CK(cudaMemcpy(d_planes, planes.data(), planes.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_b, bimg.data(), bimg.size(), cudaMemcpyHostToDevice));
  const int smem = N * 128 + (K / 8) * plane_rows * 16 + 64;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int reps = argc > 4 ? atoi(argv[4]) : 0;
  long long* d_cyc;
  CK(cudaMalloc(&d_cyc, 4 * sizeof(long long)));
  probe<<<1, 160, smem>>>(d_planes, plane_rows, d_b, shift, variant, d_out, d_cyc, reps);
  CK(cudaDeviceSynchronize());
  std::vector<float> out(128 * N);
  CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      float ref = 0.f;
      for (int k = 0; k < K; ++k) ref += A[(m + shift) * K + k] * B[n * K + k];
      if (ref != out[m * N + n]) ++bad;
    }
  if (reps > 0) {
    long long cyc[4];
    CK(cudaMemcpy(cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost));
    const double n = 4.0 * reps;
    printf("cycles per SS MMA (K16): A no-swizzle N128 %.1f  N64 %.1f | A SW128 N128 %.1f  N64 %.1f\n", cyc[0] / n,
           cyc[1] / n, cyc[2] / n, cyc[3] / n);
  }
  if (reps > 0)
    printf("nosw_probe: timing run (the timed MMAs overwrite the accumulator; run without reps to compare)\n");
  else
    printf("nosw_probe shift=%d plane_rows=%d variant=%d: %s (%d / %d mismatches)\n", shift, plane_rows, variant,
           bad ? "FAIL" : "ok", bad, 128 * N);
  return 0;
}

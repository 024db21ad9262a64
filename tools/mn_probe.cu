// mn_probe: validates MN-major (transposed) tcgen05 shared-memory operands on a real B200 -
// the form the weight-gradient kernel (csrc/backward_gemm.cu) relies on.
//
// A [points x features] 16-bit tile stored as 64-feature slabs of 128-byte rows (one row per
// point, 16-byte chunks XOR-swizzled with row % 8) is byte-for-byte BOTH
//   * a K-major SWIZZLE_128B operand with M/N = points, K = features (what dgrad reads), and
//   * an MN-major SWIZZLE_128B operand with M/N = features, K = points (what wgrad reads):
//     canonical layout ((8 chunks, m groups), (8 rows, k groups)) : ((16 B, LBO), (128 B, SBO)),
//     LBO = byte distance between 64-feature slabs, SBO = 1024 B between 8-point row groups
//     (CUTLASS cute/atom/mma_traits_sm100.hpp, make_umma_desc<Major::MN>).
// The probe computes D[m][n] = sum_p A[p][m] * B[p][n] (M = 128 features of A, N = 64..256
// features of B, K = 128 points) with both operands MN-major and compares bit-exactly with the
// host (small-integer inputs).  Variants (argv[2]) try the other plausible LBO/SBO readings so a
// wrong guess is diagnosed in ONE run.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I cr-nerf-pytorch_b200/csrc
//        tools/mn_probe.cu -o tools/mn_probe
// run  : tools/mn_probe [N=128] [variant=0] [fmt_a=0] [fmt_b=0]
#include <cuda_bf16.h>
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

constexpr int kPoints = 128;

struct Params {
  const uint8_t* a_img;  // [M/64 slabs][128 points][128 B]
  const uint8_t* b_img;  // [N/64 slabs][128 points][128 B]
  float* d_out;          // M x N
  int M, N, variant, fmt_a, fmt_b;
};

__device__ __forceinline__ uint64_t desc_mn(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3ffff) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3fff) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3fff) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(160, 1) probe_kernel(Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int aslabs = p.M / 64, bslabs = p.N / 64;
  uint8_t* sA = smem;
  uint8_t* sB = smem + aslabs * 16384;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + bslabs * 16384);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc<512>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bars[0], (aslabs + bslabs) * 16384);
    bulk_g2s(sA, p.a_img, aslabs * 16384, &bars[0]);
    bulk_g2s(sB, p.b_img, bslabs * 16384, &bars[0]);
  }
  if (warp == 4) {
    mbar_wait(&bars[0], 0, 1);
    tc_fence_after_sync();
    if (elect_one()) {
      // kind::f16, fp32 accumulate, a_major = b_major = MN (bits 15, 16)
      const uint32_t idesc = (1u << 4) | ((uint32_t)p.fmt_a << 7) | ((uint32_t)p.fmt_b << 10) | (1u << 15) |
                             (1u << 16) | ((uint32_t)(p.N >> 3) << 17) | ((128u >> 4) << 24);
      uint32_t lbo = 16384, sbo = 1024;
      if (p.variant == 1) { lbo = 1024; sbo = 16384; }
      for (int mh = 0; mh < p.M / 128; ++mh) {
        for (int ks = 0; ks < kPoints / 16; ++ks) {
          // 16 points per k-step = two 8-row groups = 2048 B further down every slab
          const uint64_t ad = desc_mn(smem_u32(sA + mh * 2 * 16384) + ks * 2048, lbo, sbo);
          const uint64_t bd = desc_mn(smem_u32(sB) + ks * 2048, lbo, sbo);
          umma_ss(tmem + mh * 256, ad, bd, idesc, ks ? 1u : 0u);
        }
      }
      umma_commit(&bars[1]);
    }
    __syncwarp();
  }
  if (warp < 4) {
    mbar_wait(&bars[1], 0, 3);
    tc_fence_after_sync();
    const int row = warp * 32 + lane;
    for (int mh = 0; mh < p.M / 128; ++mh)
      for (int c0 = 0; c0 < p.N; c0 += 32) {
        uint32_t v[32];
        tmem_ld_x32(tmem + mh * 256 + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) p.d_out[(mh * 128 + row) * p.N + c0 + j] = __uint_as_float(v[j]);
      }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc<512>(tmem);
}

static uint16_t to16(float v, int fmt) {
  uint16_t u;
  if (fmt == 0) {
    __half h = __float2half(v);
    memcpy(&u, &h, 2);
  } else {
    __nv_bfloat16 h = __float2bfloat16(v);
    memcpy(&u, &h, 2);
  }
  return u;
}

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 128;
  const int variant = argc > 2 ? atoi(argv[2]) : 0;
  const int fmt_a = argc > 3 ? atoi(argv[3]) : 0, fmt_b = argc > 4 ? atoi(argv[4]) : 0;
  const int M = argc > 5 ? atoi(argv[5]) : 128;
  if (N % 64 || N > 256 || (M != 128 && M != 256)) {
    printf("bad args\n");
    return 2;
  }
  srand(77 + N);
  std::vector<float> A(kPoints * M), B(kPoints * N);   // [point][feature]
  for (auto& v : A) v = float(rand() % 9 - 4);
  for (auto& v : B) v = float(rand() % 9 - 4) * 0.125f;
  auto image = [&](const std::vector<float>& X, int F, int fmt) {
    std::vector<uint8_t> img((F / 64) * 16384, 0);
    for (int s = 0; s < F / 64; ++s)
      for (int r = 0; r < kPoints; ++r)
        for (int c = 0; c < 64; ++c) {
          const uint16_t v = to16(X[r * F + s * 64 + c], fmt);
          memcpy(&img[s * 16384 + sw128_offset(r, c / 8) + (c % 8) * 2], &v, 2);
        }
    return img;
  };
  auto a_img = image(A, M, fmt_a), b_img = image(B, N, fmt_b);
  std::vector<float> ref(M * N, 0.f);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float acc = 0;
      for (int pnt = 0; pnt < kPoints; ++pnt) acc += A[pnt * M + m] * B[pnt * N + n];
      ref[m * N + n] = acc;
    }
  Params p{};
  uint8_t *da, *db;
  float* dd;
  CK(cudaMalloc(&da, a_img.size()));
  CK(cudaMalloc(&db, b_img.size()));
  CK(cudaMalloc(&dd, M * N * 4));
  CK(cudaMemcpy(da, a_img.data(), a_img.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, b_img.data(), b_img.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dd, 0xff, M * N * 4));
  p.a_img = da;
  p.b_img = db;
  p.d_out = dd;
  p.M = M;
  p.N = N;
  p.variant = variant;
  p.fmt_a = fmt_a;
  p.fmt_b = fmt_b;
  const size_t smem = a_img.size() + b_img.size() + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<<<1, 160, smem>>>(p);
  CK(cudaGetLastError());
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("mn_probe M=%d N=%d variant=%d: LAUNCH FAILED: %s\n", M, N, variant, cudaGetErrorString(e));
    return 1;
  }
  std::vector<float> out(M * N);
  CK(cudaMemcpy(out.data(), dd, M * N * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int i = 0; i < M * N; ++i) bad += out[i] != ref[i];
  printf("mn_probe M=%d N=%d variant=%d fmt_a=%d fmt_b=%d: mismatches=%d/%d %s\n", M, N, variant, fmt_a, fmt_b, bad,
         M * N, bad == 0 ? "OK" : "FAIL");
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

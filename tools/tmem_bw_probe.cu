// tmem_bw_probe: how fast can warps drain TMEM accumulators (tcgen05.ld), alone and
// while the tensor pipe is busy?  Decides the epilogue design of the fused kernel.
//   tools/tmem_bw_probe <shape 16|32|64> <warps 4|8> <mma 0|1> <iters> [grid]
// Each epilogue warp reads its 32 lanes x 128 columns per iteration (4 warps cover one
// 128x128 fp32 accumulator = 64 KB; with 8 warps two accumulators are read concurrently).
// With mma=1 a fifth/ninth warp issues back-to-back M128 N128 K16 TS MMAs into other columns.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ptx.cuh"
using namespace crnerf;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(2);} } while (0)

__device__ __forceinline__ void ld64(uint32_t taddr, uint32_t (&v)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
      "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
      : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]),
        "=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31]),
        "=r"(v[32]),"=r"(v[33]),"=r"(v[34]),"=r"(v[35]),"=r"(v[36]),"=r"(v[37]),"=r"(v[38]),"=r"(v[39]),"=r"(v[40]),"=r"(v[41]),"=r"(v[42]),"=r"(v[43]),"=r"(v[44]),"=r"(v[45]),"=r"(v[46]),"=r"(v[47]),
        "=r"(v[48]),"=r"(v[49]),"=r"(v[50]),"=r"(v[51]),"=r"(v[52]),"=r"(v[53]),"=r"(v[54]),"=r"(v[55]),"=r"(v[56]),"=r"(v[57]),"=r"(v[58]),"=r"(v[59]),"=r"(v[60]),"=r"(v[61]),"=r"(v[62]),"=r"(v[63])
      : "r"(taddr) : "memory");
}

struct P { long long* out; int shape, warps, mma, iters; };

__global__ void __launch_bounds__(320, 1) k(P p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 16384);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 4);
  volatile int* stop = reinterpret_cast<volatile int*>(slot + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bars[0], 1); *stop = 0; fence_mbar_init(); }
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;  // B operand = zeros
  if (warp == 9) tmem_alloc<512>(slot);
  fence_proxy_async_smem();
  tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
  const uint32_t tmem = *slot;
  if (warp < p.warps) {
    const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (warp >= 4 ? 128u : 0u);
    uint32_t acc = 0;
    __syncwarp();
    const long long t0 = clock64();
    for (int it = 0; it < p.iters; ++it) {
      if (p.shape == 16) {
#pragma unroll
        for (int c = 0; c < 8; ++c) { uint32_t v[16]; tmem_ld_x16(base + 16 * c, v); tmem_ld_wait(); acc += v[0] + v[15]; }
      } else if (p.shape == 32) {
#pragma unroll
        for (int c = 0; c < 4; ++c) { uint32_t v[32]; tmem_ld_x32(base + 32 * c, v); tmem_ld_wait(); acc += v[0] + v[31]; }
      } else if (p.shape == 64) {
#pragma unroll
        for (int c = 0; c < 2; ++c) { uint32_t v[64]; ld64(base + 64 * c, v); tmem_ld_wait(); acc += v[0] + v[63]; }
      } else if (p.shape == 100) {  // store test: 64 words per iteration (2 x st.x32) + wait::st
        uint32_t v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = acc + j;
        tmem_st_x32(base, v);
        tmem_st_x32(base + 32, v);
        tmem_st_wait();
        acc += 1;
      } else if (p.shape == 101) {  // ld x32 + st x16 + wait::st (the flush pattern)
        uint32_t v[32];
        tmem_ld_x32(base, v);
        tmem_ld_wait();
        tmem_st_x16p(base + 64, v);
        tmem_ld_x32(base + 32, v);
        tmem_ld_wait();
        tmem_st_x16p(base + 80, v);
        tmem_st_wait();
        acc += v[0];
      } else {  // 32 columns, two loads in flight before each wait
#pragma unroll
        for (int c = 0; c < 2; ++c) { uint32_t v[32], w[32]; tmem_ld_x32(base + 64 * c, v); tmem_ld_x32(base + 64 * c + 32, w); tmem_ld_wait(); acc += v[0] + w[31]; }
      }
    }
    const long long t1 = clock64();
    if (lane == 0) { p.out[blockIdx.x * 16 + warp] = t1 - t0; if (acc == 0x12345678u) p.out[0] = 0; }
    __syncwarp();
    if (warp == 0 && lane == 0) *stop = 1;
  } else if (warp == 8 && p.mma) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_f16(128, 128, 0);
      long long n = 0;
      const long long t0 = clock64();
      while (!*stop) {
        for (int ks = 0; ks < 16; ++ks)
          umma_ts(tmem + 256, tmem + 384 + (ks & 7) * 8, make_sdesc_k_sw128(smem_u32(smem) + (ks & 3) * 32, 1024), idesc, ks ? 1u : 0u);
        n += 16;
      }
      umma_commit(&bars[0]);
      mbar_wait(&bars[0], 0, 1);
      p.out[blockIdx.x * 16 + 8] = clock64() - t0;
      p.out[blockIdx.x * 16 + 9] = n;
    }
    __syncwarp();
  }
  tc_fence_before_sync(); __syncthreads();
  if (warp == 9) tmem_dealloc<512>(tmem);
}

int main(int argc, char** argv) {
  P p{};
  p.shape = argc > 1 ? atoi(argv[1]) : 32;
  p.warps = argc > 2 ? atoi(argv[2]) : 4;
  p.mma = argc > 3 ? atoi(argv[3]) : 0;
  p.iters = argc > 4 ? atoi(argv[4]) : 200;
  int grid = argc > 5 ? atoi(argv[5]) : 1;
  CK(cudaMalloc(&p.out, grid * 16 * 8));
  CK(cudaMemset(p.out, 0, grid * 16 * 8));
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 20480));
  k<<<grid, 320, 20480>>>(p);
  CK(cudaDeviceSynchronize());
  std::vector<long long> h(grid * 16);
  CK(cudaMemcpy(h.data(), p.out, grid * 16 * 8, cudaMemcpyDeviceToHost));
  long long worst = 0;
  for (int w = 0; w < p.warps; ++w) worst = h[w] > worst ? h[w] : worst;
  const double bytes_per_warp_iter = 32.0 * 128 * 4;
  const double total_bytes = bytes_per_warp_iter * p.warps * p.iters;
  printf("shape x%d warps=%d mma=%d: %lld cycles for %d iters -> %.0f cyc per 128-col drain per warp, "
         "%.1f B/cycle/SM aggregate", p.shape, p.warps, p.mma, worst, p.iters, (double)worst / p.iters,
         total_bytes / worst);
  if (p.mma) printf("; concurrent MMA rate %.1f cyc/MMA", (double)h[8] / (double)(h[9] ? h[9] : 1));
  printf("\n");
  return 0;
}

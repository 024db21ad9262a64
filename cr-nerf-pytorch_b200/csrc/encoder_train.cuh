// Training kernels of encoder_sameoutputsize (reference models/linearStyleTransfer.py:208-276 under
// autograd: train_mask_grid_sample.py back-propagates through enc_a / enc_cont every step).
// Included by encoder.cu inside its anonymous namespace (shares the convolution kernel, the plane
// helpers and the weight blob).
//
// Backward of one reflection-padded 3x3 convolution Z = conv(pad(X), W), with dZ given:
//   * input gradient  dXpad = full correlation of dZ with the flipped, transposed weights = the
//     FORWARD tensor-core kernel run on "gradient planes" (dZ with a zero halo of 2, hi/lo split) and
//     a re-packed weight image (enc_pack_tc_dgrad_kernel), fp32 rows out (kOutRaw);
//     enc_grad_prep_kernel then folds the reflection halo back into the interior, routes through the
//     2x2 max-pool where there is one (argmax recomputed from the saved pre-pool planes, torch's scan
//     order), applies LeakyReLU' from the saved activation's sign and writes the next layer's
//     gradient planes - one elementwise pass between two convolutions;
//   * weight gradient dW[co][ci][ky][kx] = sum_p dZ[co][p] Xpad[ci][p + (ky,kx)]: a GEMM with
//     K = pixels.  The planes layout [C/8][pixels][8] is byte for byte an MN-major SWIZZLE_NONE
//     tcgen05 operand (M/N = channels, SBO = chunk stride, LBO = 128 B: 8 pixels) and a tap is the
//     operand's start address shifted by whole pixels (tools/mn_nosw_probe.cu: bit-exact, shifts
//     included), so enc_wgrad_tc_kernel stages row segments of dZ and of the saved input planes with
//     plain bulk copies and accumulates the three taps of one kernel row in TMEM for the CTA's whole
//     lifetime (hi/lo operands, three MMAs per product as in the forward); per-CTA partials, one
//     fixed-order reduction.
// Gradient planes carry a power-of-two scale chosen on the device from the previous stage's
// max magnitude (fp16 range), undone by the reductions: no host synchronisation anywhere.

// gradient planes: [C/8][h + 4][Wg][8] fp16 (hi, then lo), zero halo of 2; the row stride leaves the
// zeros a 16-pixel K step may run into after the last interior column
__host__ __device__ inline int grad_stride(int w) { return ((w + 15) & ~15) + 4; }

__device__ __forceinline__ void unpack8(const uint4& h, const uint4& l, float (&f)[8]) {
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f[2 * j] = __half2float(__ushort_as_half((unsigned short)(hw[j] & 0xffffu))) +
               __half2float(__ushort_as_half((unsigned short)(lw[j] & 0xffffu)));
    f[2 * j + 1] = __half2float(__ushort_as_half((unsigned short)(hw[j] >> 16))) +
                   __half2float(__ushort_as_half((unsigned short)(lw[j] >> 16)));
  }
}
// torch's max-pool scan order over the window (0,0) (0,1) (1,0) (1,1): the first maximum wins
__device__ __forceinline__ int pool_winner(float f0, float f1, float f2, float f3) {
  int k = 0;
  float b = f0;
  if (f1 > b) { b = f1; k = 1; }
  if (f2 > b) { b = f2; k = 2; }
  if (f3 > b) { k = 3; }
  return k;
}
__device__ __forceinline__ uint32_t word_of(const uint4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
__device__ __forceinline__ uint32_t sel16(const uint4& v, int j) {   // 16-bit element j of a plane element
  const uint32_t w = word_of(v, j >> 1);
  return (j & 1) ? (w >> 16) : (w & 0xffffu);
}

// ---- forward (training): 2x2 max-pool of activation planes -> planes of the pooled layer ----------
// The training forward keeps the pre-pool activation planes (the backward recomputes the argmax and the
// LeakyReLU sign from them), so the pool is its own pass here instead of the conv epilogue's.
__global__ void __launch_bounds__(256)
enc_pool_planes_kernel(const __half* __restrict__ in_hi, const __half* __restrict__ in_lo, int C, int H, int W,
                       __half* __restrict__ out_hi, __half* __restrict__ out_lo) {
  const int Ho = H / 2, Wo = W / 2, Wp = W + 2;
  const long long per = (long long)Ho * Wo, total = per * (C / 8);
  const uint4* ih = reinterpret_cast<const uint4*>(in_hi);
  const uint4* il = reinterpret_cast<const uint4*>(in_lo);
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    const int chunk = (int)(q / per);
    const long long rem = q - (long long)chunk * per;
    const int yo = (int)(rem / Wo), xo = (int)(rem - (long long)yo * Wo);
    const size_t base = ((size_t)chunk * (H + 2) + (2 * yo + 1)) * Wp + 2 * xo + 1;
    const uint4 h[4] = {ih[base], ih[base + 1], ih[base + Wp], ih[base + Wp + 1]};
    const uint4 l[4] = {il[base], il[base + 1], il[base + Wp], il[base + Wp + 1]};
    float f[4][8];
#pragma unroll
    for (int k = 0; k < 4; ++k) unpack8(h[k], l[k], f[k]);
    uint32_t oh[4] = {0, 0, 0, 0}, ol[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = pool_winner(f[0][j], f[1][j], f[2][j], f[3][j]);
      const uint32_t hv = k == 0 ? sel16(h[0], j) : k == 1 ? sel16(h[1], j) : k == 2 ? sel16(h[2], j) : sel16(h[3], j);
      const uint32_t lv = k == 0 ? sel16(l[0], j) : k == 1 ? sel16(l[1], j) : k == 2 ? sel16(l[2], j) : sel16(l[3], j);
      oh[j >> 1] |= hv << ((j & 1) * 16);
      ol[j >> 1] |= lv << ((j & 1) * 16);
    }
    store_plane_elem(out_hi, out_lo, Ho + 2, Wo + 2, chunk, HaloTargets(yo + 1, xo + 1, Ho, Wo),
                     make_uint4(oh[0], oh[1], oh[2], oh[3]), make_uint4(ol[0], ol[1], ol[2], ol[3]));
  }
}

// ---- weight image of an input-gradient convolution: CIN' = cout, COUT' = cin, taps flipped ---------
__global__ void enc_pack_tc_dgrad_kernel(const float* __restrict__ w, int cin, int cout, uint8_t* __restrict__ img) {
  const long long total = (long long)9 * cin * cout;
  const size_t chunk = (size_t)2 * cin * 128;     // [hi: cin rows x 128 B][lo]
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % 64);
    long long r = i / 64;
    const int ci = (int)(r % cin);
    r /= cin;
    const int tap = (int)(r % 9), kb = (int)(r / 9);
    const float v = w[((size_t)(kb * 64 + k) * cin + ci) * 9 + (8 - tap)];
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    uint8_t* base = img + (size_t)(kb * 9 + tap) * chunk + sw128_offset(ci, k >> 3) + (k & 7) * 2;
    *reinterpret_cast<__half*>(base) = hi;
    *reinterpret_cast<__half*>(base + (size_t)cin * 128) = lo;
  }
}

// power-of-two rescale that brings a stage bounded by `fold` x (max of the previous stage) to <= 2^12
__device__ __forceinline__ float stage_rescale(unsigned maxbits, int fold_log2, float scale_in) {
  const float m = __uint_as_float(maxbits);
  if (!(m > 0.f) || !isfinite(m)) return 1.f;
  const float r = scalbnf(1.f, 12 - (ilogbf(m) + 1) - fold_log2);
  return isfinite(scale_in * r) && scale_in * r > 0.f ? r : 1.f;
}

// sum 8 per-thread values over the block (256 threads), fixed order; result valid in threads 0..7
__device__ __forceinline__ float block_sum8(float (&s)[8], float (*sh)[8]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
#pragma unroll
    for (int o = 16; o; o >>= 1) s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
  }
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) sh[warp][j] = s[j];
  }
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < 8) {
    for (int w = 0; w < 8; ++w) t += sh[w][threadIdx.x];
  }
  return t;
}

// ---- tail backward: LeakyReLU' . conv7 (1x1) -> d pooled; one block per 32x32 bin -------------------
__global__ void __launch_bounds__(128)
enc_tail_bwd_kernel(const float* __restrict__ g_out, const float* __restrict__ out, const float* __restrict__ blob,
                    float* __restrict__ dpre7, float* __restrict__ dpooled, unsigned* __restrict__ maxbits) {
  __shared__ float dp[64];
  const int bin = blockIdx.x, t = threadIdx.x;
  if (t < 64) {
    const float g = g_out[(size_t)t * 1024 + bin] * (out[(size_t)t * 1024 + bin] > 0.f ? 1.f : kSlope);
    dp[t] = g;
    dpre7[(size_t)bin * 64 + t] = g;
  }
  __syncthreads();
  float s = 0.f;
  const float* w = blob + Blob::w7 + t;
#pragma unroll 8
  for (int co = 0; co < 64; ++co) s = fmaf(w[co * 128], dp[co], s);
  dpooled[(size_t)bin * 128 + t] = s;
  float m = fabsf(s);
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((t & 31) == 0 && m > 0.f) atomicMax(maxbits, __float_as_uint(m));
}
// dW7[co][k] = sum_bin dpre7[bin][co] pooled[bin][k], db7[co] = sum_bin dpre7[bin][co]; one block per co,
// 8 groups of 128 bins each, combined in group order
__global__ void __launch_bounds__(1024)
enc_w7_grad_kernel(const float* __restrict__ dpre7, const float* __restrict__ pooled, float* __restrict__ gw7,
                   float* __restrict__ gb7) {
  __shared__ float d[1024];
  __shared__ float red[8][128];
  const int co = blockIdx.x, t = threadIdx.x, k = t & 127, grp = t >> 7;
  d[t] = dpre7[(size_t)t * 64 + co];
  __syncthreads();
  float s = 0.f;
  for (int b = grp * 128; b < grp * 128 + 128; ++b) s = fmaf(d[b], pooled[(size_t)b * 128 + k], s);
  red[grp][k] = s;
  __syncthreads();
  if (t < 128) {
    float a = 0.f;
#pragma unroll
    for (int g2 = 0; g2 < 8; ++g2) a += red[g2][t];
    gw7[(size_t)co * 128 + t] = a;
  } else if (t < 160) {
    const int lane = t & 31;
    float a = 0.f;
    for (int b = lane; b < 1024; b += 32) a += d[b];
#pragma unroll
    for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) gb7[co] = a;
  }
}

// ---- top of the chain: adaptive-avg-pool backward . LeakyReLU'(conv6) -> gradient planes of conv6 ---
__global__ void __launch_bounds__(256)
enc_grad_top_kernel(const float* __restrict__ dpooled, const float* __restrict__ F, int H4, int W4,
                    __half* __restrict__ g_hi, __half* __restrict__ g_lo, int Wg, const unsigned* __restrict__ maxbits_in,
                    int fold_log2, float* __restrict__ scale_out, float* __restrict__ db_part) {
  __shared__ float sh[8][8];
  const int chunk = blockIdx.y;
  const float r = stage_rescale(*maxbits_in, fold_log2, 1.f);
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *scale_out = r;
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const long long total = (long long)H4 * W4;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(p / W4), x = (int)(p - (long long)y * W4);
    float d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int bi0 = max(0, (y * 32) / H4 - 1), bi1 = min(31, ((y + 1) * 32 + H4 - 1) / H4);
    const int bj0 = max(0, (x * 32) / W4 - 1), bj1 = min(31, ((x + 1) * 32 + W4 - 1) / W4);
    for (int bi = bi0; bi <= bi1; ++bi) {
      const int y0 = (bi * H4) / 32, y1 = ((bi + 1) * H4 + 31) / 32;
      if (y < y0 || y >= y1) continue;
      for (int bj = bj0; bj <= bj1; ++bj) {
        const int x0 = (bj * W4) / 32, x1 = ((bj + 1) * W4 + 31) / 32;
        if (x < x0 || x >= x1) continue;
        const float inv = 1.f / (float)((y1 - y0) * (x1 - x0));
        const float4* dp = reinterpret_cast<const float4*>(dpooled + (size_t)(bi * 32 + bj) * 128 + chunk * 8);
        const float4 a = dp[0], b = dp[1];
        d[0] += a.x * inv; d[1] += a.y * inv; d[2] += a.z * inv; d[3] += a.w * inv;
        d[4] += b.x * inv; d[5] += b.y * inv; d[6] += b.z * inv; d[7] += b.w * inv;
      }
    }
    const float4* fp = reinterpret_cast<const float4*>(F + (size_t)p * 128 + chunk * 8);
    const float4 fa = fp[0], fb = fp[1];
    const float fv[8] = {fa.x, fa.y, fa.z, fa.w, fb.x, fb.y, fb.z, fb.w};
    float g[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      g[j] = d[j] * (fv[j] > 0.f ? 1.f : kSlope) * r;
      s[j] += g[j];
    }
    uint4 h, l;
    split8(g, h, l);
    const size_t o = (((size_t)chunk * (H4 + 4) + y + 2) * Wg + x + 2) * 8;
    *reinterpret_cast<uint4*>(g_hi + o) = h;
    *reinterpret_cast<uint4*>(g_lo + o) = l;
  }
  const float t = block_sum8(s, sh);
  if (threadIdx.x < 8) db_part[((size_t)chunk * gridDim.x + blockIdx.x) * 8 + threadIdx.x] = t;
}

// ---- between two convolutions of the backward: fold . (pool routing) . LeakyReLU' -> gradient planes --
struct PrepArgs {
  const float* dx;      // fp32 rows (hq + 2, wdx, C): gradient w.r.t. the PADDED input of the layer above, x scale_in
  int wdx, C, hq, wq;   // hq x wq: interior of that input (the pooled size when kPool)
  const __half* act_hi; // saved activation planes (C, H, W) of the layer whose dZ is produced (pre-pool when kPool)
  const __half* act_lo;
  int H, W;
  __half* g_hi;         // gradient planes out (C, H + 4, Wg)
  __half* g_lo;
  int Wg;
  const unsigned* maxbits_in;
  const float* scale_in;
  float* scale_out;
  float* db_part;       // [C/8][gridDim.x][8] per-block bias-gradient partials (scaled)
};
// Thread mapping by what dominates the traffic.  No pool: thread = (pixel, channel chunk), chunk fastest - a pixel's
// fp32 row of dx is one contiguous run over C/8 neighbouring lanes (grid 1-D).  Pool: the 4 + 4 plane reads and 8
// plane writes dominate - blockIdx.y = chunk, lanes = consecutive pixels, so every plane access of a warp is one
// contiguous run.
template <bool kPool>
__global__ void __launch_bounds__(256) enc_grad_prep_kernel(const PrepArgs a) {
  __shared__ float sh[8][16][8];
  const int nch = kPool ? 1 : a.C >> 3;           // chunks interleaved in the thread index (8 or 16: divides the strides)
  const int chunk = kPool ? (int)blockIdx.y : (int)(threadIdx.x & (nch - 1));
  const float sc_in = *a.scale_in;
  const float r = stage_rescale(*a.maxbits_in, 2, sc_in);   // a fold sums at most 4 entries
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *a.scale_out = sc_in * r;
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const long long total = (long long)a.hq * a.wq * nch;
  const uint4* ah = reinterpret_cast<const uint4*>(a.act_hi);
  const uint4* al = reinterpret_cast<const uint4*>(a.act_lo);
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    const long long p = q / nch;
    const int yq = (int)(p / a.wq), xq = (int)(p - (long long)yq * a.wq);
    // reflection pad 1: padded row 0 mirrors interior row 1, padded row hq + 1 mirrors row hq - 2
    int ys[3], xs[3], ny = 1, nx = 1;
    ys[0] = yq + 1;
    xs[0] = xq + 1;
    if (yq == 1) ys[ny++] = 0;
    if (yq == a.hq - 2) ys[ny++] = a.hq + 1;
    if (xq == 1) xs[nx++] = 0;
    if (xq == a.wq - 2) xs[nx++] = a.wq + 1;
    float d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int iy = 0; iy < ny; ++iy)
      for (int ix = 0; ix < nx; ++ix) {
        const float4* sp = reinterpret_cast<const float4*>(a.dx + ((size_t)ys[iy] * a.wdx + xs[ix]) * a.C + chunk * 8);
        const float4 u = sp[0], v = sp[1];
        d[0] += u.x; d[1] += u.y; d[2] += u.z; d[3] += u.w;
        d[4] += v.x; d[5] += v.y; d[6] += v.z; d[7] += v.w;
      }
    if constexpr (!kPool) {
      const size_t ai = ((size_t)chunk * (a.H + 2) + yq + 1) * (a.W + 2) + xq + 1;
      const uint4 h = ah[ai];
      float g[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float act = __half2float(__ushort_as_half((unsigned short)sel16(h, j)));
        g[j] = d[j] * (act > 0.f ? 1.f : kSlope) * r;
        s[j] += g[j];
      }
      uint4 gh, gl;
      split8(g, gh, gl);
      const size_t o = (((size_t)chunk * (a.H + 4) + yq + 2) * a.Wg + xq + 2) * 8;
      *reinterpret_cast<uint4*>(a.g_hi + o) = gh;
      *reinterpret_cast<uint4*>(a.g_lo + o) = gl;
    } else {
      const int Wp = a.W + 2;
      const size_t ai = ((size_t)chunk * (a.H + 2) + 2 * yq + 1) * Wp + 2 * xq + 1;
      const uint4 h[4] = {ah[ai], ah[ai + 1], ah[ai + Wp], ah[ai + Wp + 1]};
      const uint4 l[4] = {al[ai], al[ai + 1], al[ai + Wp], al[ai + Wp + 1]};
      float f[4][8];
#pragma unroll
      for (int k = 0; k < 4; ++k) unpack8(h[k], l[k], f[k]);
      float g[4][8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = pool_winner(f[0][j], f[1][j], f[2][j], f[3][j]);
        const float fw = k == 0 ? f[0][j] : k == 1 ? f[1][j] : k == 2 ? f[2][j] : f[3][j];
        const float gv = d[j] * (fw > 0.f ? 1.f : kSlope) * r;
        s[j] += gv;
#pragma unroll
        for (int q2 = 0; q2 < 4; ++q2) g[q2][j] = q2 == k ? gv : 0.f;
      }
#pragma unroll
      for (int q2 = 0; q2 < 4; ++q2) {
        uint4 gh, gl;
        split8(g[q2], gh, gl);
        const size_t o = (((size_t)chunk * (a.H + 4) + 2 * yq + (q2 >> 1) + 2) * a.Wg + 2 * xq + (q2 & 1) + 2) * 8;
        *reinterpret_cast<uint4*>(a.g_hi + o) = gh;
        *reinterpret_cast<uint4*>(a.g_lo + o) = gl;
      }
    }
  }
  // bias-gradient partials: lanes of equal chunk first (fixed order), then the 8 warps
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s[j] += __shfl_xor_sync(0xffffffffu, s[j], 16);
    if (nch <= 8) s[j] += __shfl_xor_sync(0xffffffffu, s[j], 8);
    if (nch == 1) {
      s[j] += __shfl_xor_sync(0xffffffffu, s[j], 4);
      s[j] += __shfl_xor_sync(0xffffffffu, s[j], 2);
      s[j] += __shfl_xor_sync(0xffffffffu, s[j], 1);
    }
  }
  if (lane < nch) {
#pragma unroll
    for (int j = 0; j < 8; ++j) sh[warp][lane][j] = s[j];
  }
  __syncthreads();
  if (threadIdx.x < nch * 8) {
    const int c = threadIdx.x >> 3, j = threadIdx.x & 7;
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += sh[w][c][j];
    a.db_part[((size_t)(kPool ? chunk : c) * gridDim.x + blockIdx.x) * 8 + j] = t;
  }
}

// bias gradient: db[c] = (1 / scale) * sum over blocks; one warp per channel, fixed order
__global__ void __launch_bounds__(256)
enc_db_reduce_kernel(const float* __restrict__ part, int nblk, int C, const float* __restrict__ scale,
                     float* __restrict__ out) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= C) return;
  const float* p = part + (size_t)(c >> 3) * nblk * 8 + (c & 7);
  float s = 0.f;
  for (int b = lane; b < nblk; b += 32) s += p[(size_t)b * 8];
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[c] = s / *scale;
}
// generic: out[i] = (1 / scale) * sum_b part[b * row_stride + offset + i]; one warp per output, fixed order
__global__ void __launch_bounds__(256)
enc_part_reduce_kernel(const float* __restrict__ part, int nblk, int row_stride, int offset, int n,
                       const float* __restrict__ scale, float* __restrict__ out) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= n) return;
  float s = 0.f;
  for (int b = lane; b < nblk; b += 32) s += part[(size_t)b * row_stride + offset + i];
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[i] = s / *scale;
}

// ---- weight gradient of a 3x3 convolution on the tensor cores ---------------------------------------
template <int CIN, int COUT>
struct WgradCfg {
  static constexpr int kSeg = (CIN == 64 && COUT == 64) ? 128 : 64;   // pixels of one output row per stage
  static constexpr int kGChunks = COUT / 8 * 2;                       // hi chunks, then lo chunks
  static constexpr int kXChunks = CIN / 8 * 2;
  static constexpr int kGStride = kSeg * 16;                          // bytes between channel chunks (SBO)
  static constexpr int kXStride = (kSeg + 2) * 16;
  static constexpr int kGBytes = kGChunks * kGStride;
  static constexpr int kXBytes = kXChunks * kXStride;
  static constexpr int kStageBytes = kGBytes + kXBytes;
  static constexpr int kStages = (200 * 1024) / kStageBytes > 4 ? 4 : (200 * 1024) / kStageBytes;
  static constexpr int kSmem = kStages * kStageBytes + 256;
};
constexpr int kWgradThreads = 256;

__device__ __forceinline__ uint64_t make_sdesc_mn_nosw(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3ffff) >> 4);
  d |= static_cast<uint64_t>(128 >> 4) << 16;                      // LBO: the next 8 pixels (K)
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fff) << 32;     // SBO: the next 8-channel chunk (M / N)
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

// CTA b: kernel row ky = b % 3, row segments b / 3, b / 3 + gridDim.x / 3, ... of the output; accumulators
// D[kx] (128 lanes x CIN columns, kx = 0..2) live in TMEM until the end.  COUT = 64: the hi and the lo
// gradient chunks of a stage are adjacent, i.e. ONE 128-row operand [dZ_hi; dZ_lo] - two M = 128 MMAs per
// product (rows 64..127 hold the lo terms, added by the reduction) instead of three half-empty M = 64 ones.
// part: [gridDim.x][3][128][CIN] fp32.
template <int CIN, int COUT>
__global__ void __launch_bounds__(kWgradThreads, 1)
enc_wgrad_tc_kernel(const __half* __restrict__ g_hi, const __half* __restrict__ g_lo, long long g_plane, int Wg,
                    const __half* __restrict__ x_hi, const __half* __restrict__ x_lo, long long x_plane, int H, int W,
                    float* __restrict__ part) {
  using Cfg = WgradCfg<CIN, COUT>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* full = bars;                 // [kStages]
  uint64_t* empty = full + kStages;      // [kStages]
  uint64_t* d_full = empty + kStages;    // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#if defined(CRNERF_WGRAD_DIAG) && CRNERF_WGRAD_DIAG == 3
  __shared__ long long ts[8];
#define WSTAMP(k) do { if (lane == 0) ts[k] = clock64(); } while (0)
  if (threadIdx.x == 0) ts[0] = clock64();
#else
#define WSTAMP(k) do { } while (0)
#endif
  // stale shared memory may hold NaN patterns; operands past a row's end are multiplied by zero gradients
  for (int i = threadIdx.x; i < kStages * Cfg::kStageBytes / 16; i += kWgradThreads)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(d_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  if (warp == 3) WSTAMP(1);

  const int ky = blockIdx.x % 3, grp = blockIdx.x / 3, ngrp = gridDim.x / 3;
  const int nseg_row = (W + Cfg::kSeg - 1) / Cfg::kSeg, nseg = H * nseg_row;
  const int Wp = W + 2;

  if (warp == 0) {
    uint32_t it = 0;
    for (int sg = grp; sg < nseg; sg += ngrp, ++it) {
      const uint32_t st = it % kStages;
      if (it >= (uint32_t)kStages) mbar_wait(&empty[st], (it / kStages - 1) & 1, 21);
      const int y = sg / nseg_row, sx = (sg - y * nseg_row) * Cfg::kSeg;
      const int nks = (min(Cfg::kSeg, W - sx) + 15) >> 4;
      // the X run stops at the row's end: entries past it keep stale (finite) data and meet zero gradients
      const uint32_t gbytes = (uint32_t)nks * 256, xbytes = (uint32_t)min(nks * 16 + 2, Wp - sx) * 16;
      uint8_t* stage = smem + st * Cfg::kStageBytes;
#if defined(CRNERF_WGRAD_DIAG) && CRNERF_WGRAD_DIAG == 1   // diagnostic: no operand traffic after the first ring fill
      if (it >= (uint32_t)kStages) {
        if (lane == 0) mbar_arrive(&full[st]);
        continue;
      }
#endif
      if (lane == 0) mbar_arrive_expect_tx(&full[st], Cfg::kGChunks * gbytes + Cfg::kXChunks * xbytes);
      __syncwarp();
      for (int c = lane; c < Cfg::kGChunks; c += 32) {
        const int half = c / (COUT / 8), ch = c - half * (COUT / 8);
        const __half* src = (half ? g_lo : g_hi) + ((size_t)ch * g_plane + (size_t)(y + 2) * Wg + sx + 2) * 8;
        bulk_g2s(stage + c * Cfg::kGStride, src, gbytes, &full[st]);
      }
      for (int c = lane; c < Cfg::kXChunks; c += 32) {
        const int half = c / (CIN / 8), ch = c - half * (CIN / 8);
        const __half* src = (half ? x_lo : x_hi) + ((size_t)ch * x_plane + (size_t)(y + ky) * Wp + sx) * 8;
        bulk_g2s(stage + Cfg::kGBytes + c * Cfg::kXStride, src, xbytes, &full[st]);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_f16(128, CIN, 0) | (1u << 15) | (1u << 16);   // both operands MN-major
    uint32_t it = 0;
    for (int sg = grp; sg < nseg; sg += ngrp, ++it) {
      const uint32_t st = it % kStages;
      mbar_wait(&full[st], (it / kStages) & 1, 22);
      tc_fence_after_sync();
      if (it == 0) WSTAMP(2);
      const int y = sg / nseg_row, sx = (sg - y * nseg_row) * Cfg::kSeg;
      const int nks = (min(Cfg::kSeg, W - sx) + 15) >> 4;
      if (elect_one()) {
        const uint32_t gbase = smem_u32(smem + st * Cfg::kStageBytes), xbase = gbase + Cfg::kGBytes;
        const uint64_t ah = make_sdesc_mn_nosw(gbase, Cfg::kGStride);
        const uint64_t xh = make_sdesc_mn_nosw(xbase, Cfg::kXStride);
        constexpr uint64_t kALo = (uint64_t)((COUT / 8) * Cfg::kGStride) >> 4;
        constexpr uint64_t kXLo = (uint64_t)((CIN / 8) * Cfg::kXStride) >> 4;
        // all MMAs of one accumulator back to back (alternating accumulators per instruction is slower)
#if defined(CRNERF_WGRAD_DIAG) && CRNERF_WGRAD_DIAG == 2   // diagnostic: operand traffic only
        if (false)
#endif
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const uint32_t d = tmem + kx * CIN;
          for (int ks = 0; ks < nks; ++ks) {
            const uint64_t ja = (uint64_t)(ks * 16), jx = (uint64_t)(ks * 16 + kx);   // 16-byte units = pixels
            const uint32_t first = (it | (uint32_t)ks) ? 1u : 0u;
            if constexpr (COUT == 64) {
              umma_ss(d, ah + ja, xh + jx, idesc, first);            // [hi; lo] * hi
              umma_ss(d, ah + ja, xh + kXLo + jx, idesc, 1u);        // [hi; lo] * lo
            } else {
              umma_ss(d, ah + ja, xh + jx, idesc, first);            // hi * hi
              umma_ss(d, ah + kALo + ja, xh + jx, idesc, 1u);        // lo * hi
              umma_ss(d, ah + ja, xh + kXLo + jx, idesc, 1u);        // hi * lo
            }
          }
        }
        umma_commit(&empty[st]);
      }
      __syncwarp();
    }
    WSTAMP(3);
    if (elect_one()) umma_commit(d_full);
    __syncwarp();
  } else if (warp >= 4) {
    const int quarter = warp & 3;
    const bool any = grp < nseg;      // a CTA without segments has issued nothing: its partial is zero
    mbar_wait(d_full, 0, 23);
    tc_fence_after_sync();
    if (warp == 4) WSTAMP(4);
    const int row = quarter * 32 + lane;
#pragma unroll 1
    for (int kx = 0; kx < 3; ++kx) {
      float4* o = reinterpret_cast<float4*>(part + (((size_t)blockIdx.x * 3 + kx) * 128 + row) * CIN);
#pragma unroll 1
      for (int c0 = 0; c0 < CIN; c0 += 32) {
        uint32_t v[32];
        if (any) {
          tmem_ld_x32(tmem + (static_cast<uint32_t>(quarter * 32) << 16) + kx * CIN + c0, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          o[(c0 + j) / 4] = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                        __uint_as_float(v[j + 3]));
      }
    }
  }
  if (warp == 4) WSTAMP(5);
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem);
#if defined(CRNERF_WGRAD_DIAG) && CRNERF_WGRAD_DIAG == 3
  if (threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == 100))
    printf("wgrad<%d,%d> cta %d (%d x %d): prologue %lld | first operands +%lld | issue loop end +%lld | accumulators done +%lld | epilogue +%lld | total %lld\n",
           CIN, COUT, (int)blockIdx.x, H, W, ts[1] - ts[0], ts[2] - ts[1], ts[3] - ts[2], ts[4] - ts[3], ts[5] - ts[4], clock64() - ts[0]);
#endif
}

// dW[co][ci][ky][kx] = (1 / scale) * sum over the CTAs of kernel row ky (CTA order) of their partials
template <int CIN, int COUT>
__global__ void __launch_bounds__(256)
enc_wgrad_reduce_kernel(const float* __restrict__ part, int ncta, const float* __restrict__ scale, float* __restrict__ gw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // (ky, kx, co, ci), ci fastest: coalesced partial reads
  if (i >= 9 * COUT * CIN) return;
  const int ci = i % CIN, co = (i / CIN) % COUT, tap = i / (CIN * COUT), ky = tap / 3, kx = tap - 3 * ky;
  float s = 0.f;
  for (int b = ky; b < ncta; b += 3) {
    const float* p = part + (((size_t)b * 3 + kx) * 128 + co) * CIN + ci;
    s += COUT == 64 ? p[0] + p[(size_t)64 * CIN] : p[0];
  }
  gw[((size_t)co * CIN + ci) * 9 + tap] = s / *scale;
}

// ---- conv2 (3 -> 64) weight gradient on the CUDA cores (N = 27 is no tensor-core shape) --------------
// block: 64-pixel row segments; lane = one of the 27 (ci, tap) pairs (5 lanes idle), warp = 8 output channels:
// per pixel a thread reads its input value once and the warp's 8 gradient values as two broadcast vectors.
// part: [gridDim.x][64 * 27]
constexpr int kFwSeg = 64;
__global__ void __launch_bounds__(256)
enc_first_wgrad_kernel(const __half* __restrict__ g_hi, const __half* __restrict__ g_lo, long long g_plane, int Wg,
                       const __half* __restrict__ p_hi, const __half* __restrict__ p_lo, int H, int W,
                       float* __restrict__ part) {
  __shared__ __align__(16) float g_s[kFwSeg][68];       // [pixel][co]; 68: consecutive pixels' 16-byte stores hit distinct banks
  __shared__ float p_s[3][3][kFwSeg + 2];               // [ci][row][pixel]
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int cc = min(lane, 26), ci = cc / 9, tap = cc - ci * 9, ky = tap / 3, kx = tap - ky * 3;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int nseg_row = (W + kFwSeg - 1) / kFwSeg, nseg = H * nseg_row, Wp = W + 2;
  const uint4* gh = reinterpret_cast<const uint4*>(g_hi);
  const uint4* gl = reinterpret_cast<const uint4*>(g_lo);
  const uint4* ph = reinterpret_cast<const uint4*>(p_hi);
  const uint4* pl = reinterpret_cast<const uint4*>(p_lo);
  for (int sg = blockIdx.x; sg < nseg; sg += gridDim.x) {
    const int y = sg / nseg_row, sx = (sg - y * nseg_row) * kFwSeg, valid = min(kFwSeg, W - sx);
    __syncthreads();
    for (int i = t; i < 8 * kFwSeg; i += 256) {
      const int px = i % kFwSeg, ch = i / kFwSeg;
      float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (px < valid) {
        const size_t o = (size_t)ch * g_plane + (size_t)(y + 2) * Wg + sx + 2 + px;
        unpack8(gh[o], gl[o], f);
      }
      *reinterpret_cast<float4*>(&g_s[px][ch * 8]) = make_float4(f[0], f[1], f[2], f[3]);
      *reinterpret_cast<float4*>(&g_s[px][ch * 8 + 4]) = make_float4(f[4], f[5], f[6], f[7]);
    }
    for (int i = t; i < 3 * (kFwSeg + 2); i += 256) {
      const int px = i % (kFwSeg + 2), r = i / (kFwSeg + 2);
      float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (px < valid + 2) {
        const size_t o = (size_t)(y + r) * Wp + sx + px;
        unpack8(ph[o], pl[o], f);
      }
      p_s[0][r][px] = f[0];
      p_s[1][r][px] = f[1];
      p_s[2][r][px] = f[2];
    }
    __syncthreads();
    const float* pp = &p_s[ci][ky][kx];
#pragma unroll 4
    for (int p = 0; p < valid; ++p) {
      const float x = pp[p];
      const float4 a = *reinterpret_cast<const float4*>(&g_s[p][warp * 8]);
      const float4 b = *reinterpret_cast<const float4*>(&g_s[p][warp * 8 + 4]);
      acc[0] = fmaf(a.x, x, acc[0]); acc[1] = fmaf(a.y, x, acc[1]); acc[2] = fmaf(a.z, x, acc[2]); acc[3] = fmaf(a.w, x, acc[3]);
      acc[4] = fmaf(b.x, x, acc[4]); acc[5] = fmaf(b.y, x, acc[5]); acc[6] = fmaf(b.z, x, acc[6]); acc[7] = fmaf(b.w, x, acc[7]);
    }
  }
  if (lane < 27) {
#pragma unroll
    for (int j = 0; j < 8; ++j) part[(size_t)blockIdx.x * 1728 + (warp * 8 + j) * 27 + lane] = acc[j];
  }
}

// ---- conv2 input gradient (64 -> 3) at every PADDED position, four horizontally adjacent ones per thread ---
// dp0: [(H + 2)][(W + 2)] float4 (3 used), x scale
__global__ void __launch_bounds__(128)
enc_first_dgrad_kernel(const __half* __restrict__ g_hi, const __half* __restrict__ g_lo, long long g_plane, int Wg,
                       const float* __restrict__ blob, int H, int W, float4* __restrict__ dp0) {
  __shared__ float4 w_s[9][64];       // [tap][co] -> (ci 0, 1, 2, -)
  for (int i = threadIdx.x; i < 9 * 64; i += 128) {
    const int tap = i / 64, co = i % 64;
    const float* w2t = blob + Blob::w2t;    // [ci * 9 + tap][co]
    w_s[tap][co] = make_float4(w2t[(0 * 9 + tap) * 64 + co], w2t[(1 * 9 + tap) * 64 + co], w2t[(2 * 9 + tap) * 64 + co], 0.f);
  }
  __syncthreads();
  const uint4* gh = reinterpret_cast<const uint4*>(g_hi);
  const uint4* gl = reinterpret_cast<const uint4*>(g_lo);
  const int Wp = W + 2, nq = (Wp + 3) / 4;
  const long long total = (long long)(H + 2) * nq;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    const int yp = (int)(q / nq), xp0 = (int)(q - (long long)yp * nq) * 4;
    float d[4][3];
#pragma unroll
    for (int i = 0; i < 4; ++i) d[i][0] = d[i][1] = d[i][2] = 0.f;
    // padded input (yp, xp) receives dZ2 at (yp - ky, xp - kx): gradient-plane position (yp - ky + 2, xp - kx + 2)
#pragma unroll 1
    for (int ky = 0; ky < 3; ++ky) {
      const size_t row = (size_t)(yp - ky + 2) * Wg + xp0;
#pragma unroll 1
      for (int ch = 0; ch < 8; ++ch) {
        float f[6][8];
#pragma unroll
        for (int c = 0; c < 6; ++c) unpack8(gh[(size_t)ch * g_plane + row + c], gl[(size_t)ch * g_plane + row + c], f[c]);
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 w = w_s[ky * 3 + kx][ch * 8 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float g = f[i - kx + 2][j];
              d[i][0] = fmaf(g, w.x, d[i][0]);
              d[i][1] = fmaf(g, w.y, d[i][1]);
              d[i][2] = fmaf(g, w.z, d[i][2]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (xp0 + i < Wp) dp0[(size_t)yp * Wp + xp0 + i] = make_float4(d[i][0], d[i][1], d[i][2], 0.f);
  }
}

// ---- fold of dp0 . conv1 (1x1) gradients, one thread per pixel -----------------------------------------
// part: [gridDim.x][12] = dW1 (9, scaled) | db1 (3, scaled); gimg (3, H, W) optional, unscaled
__global__ void __launch_bounds__(256)
enc_first_finish_kernel(const float4* __restrict__ dp0, const float* __restrict__ blob, const float* __restrict__ img,
                        int H, int W, const float* __restrict__ scale, float* __restrict__ part,
                        float* __restrict__ gimg) {
  __shared__ float red[8][12];
  float s[12] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const float inv = 1.f / *scale;
  const float* w1 = blob + Blob::w1;
  const long long total = (long long)H * W;
  const int Wp = W + 2;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(p / W), x = (int)(p - (long long)y * W);
    int ys[3], xs[3], ny = 1, nx = 1;
    ys[0] = y + 1;
    xs[0] = x + 1;
    if (y == 1) ys[ny++] = 0;
    if (y == H - 2) ys[ny++] = H + 1;
    if (x == 1) xs[nx++] = 0;
    if (x == W - 2) xs[nx++] = W + 1;
    float d0 = 0.f, d1 = 0.f, d2 = 0.f;
    for (int iy = 0; iy < ny; ++iy)
      for (int ix = 0; ix < nx; ++ix) {
        const float4 v = dp0[(size_t)ys[iy] * Wp + xs[ix]];
        d0 += v.x;
        d1 += v.y;
        d2 += v.z;
      }
    const float v0 = img[p], v1 = img[total + p], v2 = img[2 * total + p];
    s[0] += d0 * v0; s[1] += d0 * v1; s[2] += d0 * v2;
    s[3] += d1 * v0; s[4] += d1 * v1; s[5] += d1 * v2;
    s[6] += d2 * v0; s[7] += d2 * v1; s[8] += d2 * v2;
    s[9] += d0; s[10] += d1; s[11] += d2;
    if (gimg) {
      gimg[p] = (w1[0] * d0 + w1[3] * d1 + w1[6] * d2) * inv;
      gimg[total + p] = (w1[1] * d0 + w1[4] * d1 + w1[7] * d2) * inv;
      gimg[2 * total + p] = (w1[2] * d0 + w1[5] * d1 + w1[8] * d2) * inv;
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < 12; ++j) {
#pragma unroll
    for (int o = 16; o; o >>= 1) s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
  }
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < 12; ++j) red[warp][j] = s[j];
  }
  __syncthreads();
  if (threadIdx.x < 12) {
    float a = 0.f;
    for (int w = 0; w < 8; ++w) a += red[w][threadIdx.x];
    part[(size_t)blockIdx.x * 12 + threadIdx.x] = a;
  }
}

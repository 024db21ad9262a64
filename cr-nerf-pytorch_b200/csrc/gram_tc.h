// Interface between crossray.cu (orchestration of the cross-ray block) and gram_tc.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/crnerf_b200.h"

namespace crnerf {

// One feature map whose pixel-MLP Gram statistic is wanted (CNN.forward, linearStyleTransfer.py:28-37).
struct GramJob {
  const float* g;                 // element (pixel p, channel c) at g[p*pix_stride + c*ch_stride]
  long long n, pix_stride, ch_stride;
  // channel mean subtracted from every pixel: mean_scale * (sum of the n_parts rows of sum_parts
  // (n_parts, 64)), or - self_mean - mean_scale * (sum over the map itself; small maps only)
  const float* sum_parts;
  int n_parts;
  float mean_scale;
  int self_mean;
  crnerf_cnn_weights w;
  float* partial;                 // out: (n_blocks, 1024) un-normalised Gram partials
  float* mean_out;                // out, optional: the 64 means
  int reverse;                    // walk the map from its end (L2 reuse after a forward pass)
  // filled by gram_tc_launch
  int vec, first_block, n_blocks;
};

int gram_blocks(int64_t n_pixels, int max_blocks);
int gram_tc_launch(GramJob* jobs, int n_jobs, int max_blocks, cudaStream_t st);

}  // namespace crnerf

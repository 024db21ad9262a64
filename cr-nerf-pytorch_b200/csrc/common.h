// Internal helpers shared by the translation units of libcrnerf_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/crnerf_b200.h"

namespace crnerf {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);  // records message, returns CRNERF_ERR_DEVICE
void count_launch(int n = 1);
int num_sms();  // SM count of the current device (cached per device)

// debug hook (tests): per-layer activation dump of the fused kernel
extern float* g_dbg_buf;
extern int g_dbg_layer;

// ---- internal entry points implemented in the kernel translation units ----
struct RenderArgs {
  const void* packed;
  int operand;       // crnerf_operand
  int e_xyz, e_dir;  // embedding widths the weights were packed for
  const float* rays;
  const float* view_dir;
  const float* z_vals;
  const float* noise;
  const float* x;  // pre-embedded rows (mlp_forward) or nullptr
  int x_stride;
  int sigma_only;
  int64_t n_points;  // rays * samples, or rows of x
  int n_rays, n_samples;
  int n_freq_xyz, n_freq_dir;
  float* weights;
  float* feature;
  float* depth;
  float* raw;  // (n,65) or (n,1) output of mlp_forward
  void* acts;       // training forward: saved activations (crnerf_render_acts_bytes), or nullptr
  float* raw_save;  // training forward: (n_points, 65), or nullptr
  const float* jitter;   // optional (n_points, 3) xyz jitter (crnerf_render_opts)
  float* chan_part;      // optional (render_partial_rows(), 64) per-CTA channel sums of `feature`
  int32_t* overflow;     // optional fp16 saturation flag
};
int launch_render(const RenderArgs& a, cudaStream_t st);
int render_partial_rows(int n_rays, int n_samples);
size_t mlp_packed_bytes(int e_xyz, int e_dir, int operand);
int debug_program(int e_xyz, int e_dir, int32_t* out, int cap);
int mlp_pack(const crnerf_mlp_weights* w, int operand, void* packed, size_t packed_bytes,
             int32_t* status_dev, cudaStream_t st);
// backward_gemm.cu: the training step's backward through one render pass
size_t render_acts_bytes(int64_t n_points);
int render_acts_zero_tail(void* acts, int64_t n_points, cudaStream_t st);
size_t bwd_packed_bytes(int e_xyz);
size_t bwd_scratch_bytes(int64_t n_points);
int render_backward(const crnerf_mlp_weights* w, int operand, const void* acts, const float* raw, const float* z,
                    const float* noise, const float* g_feature, const float* g_weights, const float* g_depth,
                    int n_rays, int n_samples, void* bwd_weights, void* scratch, float* const* gw,
                    float* const* gb, cudaStream_t st);

}  // namespace crnerf

#define CRNERF_CUDA(x)                                          \
  do {                                                          \
    cudaError_t e_ = (x);                                       \
    if (e_ != cudaSuccess) return crnerf::cuda_fail(e_, #x);    \
  } while (0)

#define CRNERF_REQUIRE(cond, ...)        \
  do {                                   \
    if (!(cond)) {                       \
      crnerf::set_error(__VA_ARGS__);    \
      return CRNERF_ERR_ARG;             \
    }                                    \
  } while (0)

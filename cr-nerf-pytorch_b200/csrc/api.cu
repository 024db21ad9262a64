// extern "C" surface of libcrnerf_b200.so (declared in include/crnerf_b200.h).
// Argument validation lives here; kernels live in nerf_mlp.cu / sample.cu /
// crossray.cu.  No torch types, no allocation, no host synchronisation.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include "common.h"

namespace crnerf {

constexpr int kMaxExyzPublic = 96;
static thread_local char t_err[512] = "";
static std::atomic<uint64_t> g_launches{0};
float* g_dbg_buf = nullptr;
int g_dbg_layer = -1;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error '%s' in %s", cudaGetErrorString(e), what);
  return CRNERF_ERR_DEVICE;
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int num_sms() {
  static std::mutex mu;
  static int cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  std::lock_guard<std::mutex> lk(mu);
  if (cache[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cache[dev] = v;
  }
  return cache[dev];
}

static int device_check() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute");
  if (major != 10) {
    set_error("crnerf_b200 kernels are built for sm_100a only; device %d has compute capability %d.x "
              "(there is no fallback path)", dev, major);
    return CRNERF_ERR_DEVICE;
  }
  return CRNERF_OK;
}

// implemented in sample.cu / crossray.cu
int pos_embed(const float* x, int64_t n, int n_freqs, float* out, cudaStream_t st);
int coarse_z(const float* rays, const float* t_steps, const float* perturb_rand, int n_rays,
             int n_samples, int use_disp, float* z, cudaStream_t st);
int sample_pdf(const float* bins_or_z, const float* weights, const float* u, int64_t u_stride,
               int n_rays, int m, int n_imp, float eps, float* samples, float* sorted, bool merge,
               cudaStream_t st);
size_t style_scratch_floats(int64_t);
int style_stats1(const float* content, int64_t n, int64_t ps, int64_t cs, float* sums,
                 float* partial, cudaStream_t st);
int style_stats2(const crnerf_cnn_weights& cw, const float* content, int64_t n, int64_t ps,
                 int64_t cs, const float* mean, float* gram, float* partial, float scale,
                 cudaStream_t st);
int style_finish(const crnerf_style_weights* w, const float* content, int64_t n, int64_t ps,
                 int64_t cs, const float* mean_c, const float* gram_c_normalised,
                 const float* style, int64_t ns, int64_t sps, int64_t scs, float* rgb,
                 float* transmatrix, float* fused, float* scratch, cudaStream_t st);
int style_forward(const crnerf_style_weights* w, const float* content, int64_t n, int64_t ps,
                  int64_t cs, const float* style, int64_t ns, int64_t sps, int64_t scs,
                  const float* content_sum_parts, int n_parts, float* rgb,
                  float* transmatrix, float* fused, float* scratch, cudaStream_t st);
int sum_rows(const float* parts, int n_parts, int len, float* out, cudaStream_t st);
size_t style_aux_floats();
size_t style_backward_grads_floats();
size_t style_backward_scratch_floats(int64_t n, int64_t m);
void style_backward_layout(int64_t* out22);
int style_forward_train(const crnerf_style_weights* w, const float* content, int64_t n, int64_t ps, int64_t cs,
                        const float* style, int64_t ns, int64_t sps, int64_t scs, const float* content_sum_parts,
                        int n_parts, float* rgb, float* aux, float* scratch, cudaStream_t st);
int style_backward(const crnerf_style_weights* w, const float* content, int64_t n, int64_t ps, int64_t cs,
                   const float* style, int64_t m, int64_t sps, int64_t scs, const float* aux, const float* g_rgb,
                   float* g_content, float* g_style, float* grads, float* scratch, cudaStream_t st);
int cnn_forward(const crnerf_cnn_weights* cw, const float* x, int64_t n, int64_t ps, int64_t cs,
                float* out, float* scratch, cudaStream_t st);
int grid_patch(const float* lin_w, const float* lin_h, int g, float img_w, float img_h, float scale, float h_off,
               float w_off, const float* all_rays, const float* all_rgbs, long long n_cache_rows,
               float image_offset, float* rays, long long* ts, float* rgbs, long long* rgb_idx, float* uv,
               int* status, cudaStream_t st);
int generate_rays(const float* intr4_host, const float* c2w12_host, float near, float far, int H, int W,
                  float* rays, cudaStream_t st);
int rgb_to_u8(const float* rgb, int64_t n, uint8_t* out, cudaStream_t st);
// implemented in backward.cu
int composite_backward(const float* raw, const float* z, const float* noise, const float* g_feature,
                       const float* g_weights, const float* g_depth, int n_rays, int n_samples,
                       float* d_rgb_pre, float* d_sigma_pre, cudaStream_t st, int split = 0, float* amax = nullptr);
int relu_bias_grad(float* g, const void* act, int64_t n_points, int width, float* gb, float* scratch,
                   cudaStream_t st);
// implemented in optim.cu
int adam_step(int n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
              float* const* exp_avg_sq, const int64_t* numel, const float* step, const float* lr_dev, double lr,
              double beta1, double beta2, double eps, double weight_decay, int maximize, cudaStream_t st);
// implemented in encoder.cu
size_t encoder_packed_bytes();
size_t encoder_scratch_bytes(int H, int W);
int encoder_pack(const crnerf_encoder_weights* w, void* packed, size_t packed_bytes, cudaStream_t st);
size_t encoder_tape_bytes(int H, int W);
size_t encoder_backward_scratch_bytes(int H, int W);
int encoder_forward_train(const void* packed, const float* img, int H, int W, float* out, void* tape,
                          size_t tape_bytes, cudaStream_t st);
int encoder_backward(const void* packed, const float* img, int H, int W, const float* out, const float* grad_out,
                     const void* tape, float* const* gw, float* const* gb, float* grad_img, void* scratch,
                     size_t scratch_bytes, cudaStream_t st);
int encoder_forward(const void* packed, const float* img, int H, int W, float* out, void* scratch,
                    size_t scratch_bytes, cudaStream_t st);
// implemented in loss.cu
size_t loss_scratch_floats();
int ray_loss_forward(const float* coarse, const float* fine, const float* target, const float* mask,
                     int64_t n_rays, float coef, float size_delta, float digit_delta, float* out4,
                     float* scratch, cudaStream_t st);
int ray_loss_backward(const float* coarse, const float* fine, const float* target, const float* mask,
                      int64_t n_rays, float coef, float size_delta, float digit_delta, const float* go4,
                      float* g_coarse, float* g_fine, float* g_mask, cudaStream_t st);
int pair_loss_forward(int n_terms, const float* const* a, const float* const* b, const int64_t* n, const int* mode,
                      const float* scale, float* out, float* scratch, cudaStream_t st);
int pair_loss_backward(int n_terms, const float* const* a, const float* const* b, const int64_t* n, const int* mode,
                       const float* scale, const float* go, float* const* ga, float* const* gb, cudaStream_t st);
int mask_sample_forward(const float* pred, int channels, int h, int w, int H, int W, const int64_t* idx, int64_t n,
                        float* out, cudaStream_t st);
int mask_sample_backward(const float* g_out, int channels, int h, int w, int H, int W, const int64_t* idx, int64_t n,
                         float* g_pred, cudaStream_t st);

}  // namespace crnerf

using namespace crnerf;

extern "C" {

const char* crnerf_last_error(void) { return t_err; }
int crnerf_abi_version(void) { return CRNERF_ABI_VERSION; }
int crnerf_device_ok(void) { return device_check() == CRNERF_OK ? 1 : 0; }
uint64_t crnerf_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int crnerf_debug_set(float* dbg_buf, int layer) {
  g_dbg_buf = dbg_buf;
  g_dbg_layer = layer;
  return CRNERF_OK;
}

int crnerf_debug_program(int e_xyz, int e_dir, int32_t* out_host, int cap) {
  return debug_program(e_xyz, e_dir, out_host, cap);
}

size_t crnerf_mlp_packed_bytes(int e_xyz, int e_dir) { return mlp_packed_bytes(e_xyz, e_dir, 0); }
size_t crnerf_mlp_packed_bytes_op(int e_xyz, int e_dir, int operand) {
  return mlp_packed_bytes(e_xyz, e_dir, operand);
}
int crnerf_render_partial_rows(int n_rays, int n_samples) { return render_partial_rows(n_rays, n_samples); }

int crnerf_mlp_pack(const crnerf_mlp_weights* w, int operand, void* packed, size_t packed_bytes,
                    int32_t* status_dev, void* stream) {
  int rc = device_check();
  if (rc) return rc;
  return mlp_pack(w, operand, packed, packed_bytes, status_dev, (cudaStream_t)stream);
}

static int render_pass_impl(const void* packed, int operand, const float* rays, const float* view_dir,
                            const float* z_vals, const float* noise, int n_rays, int n_samples,
                            int n_freq_xyz, int n_freq_dir, float* weights, float* feature,
                            float* depth, void* acts, float* raw_save, const crnerf_render_opts* opts,
                            void* stream);

int crnerf_render_pass(const void* packed, int operand, const float* rays, const float* view_dir,
                       const float* z_vals, const float* noise, int n_rays, int n_samples,
                       int n_freq_xyz, int n_freq_dir, float* weights, float* feature,
                       float* depth, void* stream) {
  return render_pass_impl(packed, operand, rays, view_dir, z_vals, noise, n_rays, n_samples, n_freq_xyz,
                          n_freq_dir, weights, feature, depth, nullptr, nullptr, nullptr, stream);
}

int crnerf_render_pass_opts(const void* packed, int operand, const float* rays, const float* view_dir,
                            const float* z_vals, const float* noise, int n_rays, int n_samples,
                            int n_freq_xyz, int n_freq_dir, float* weights, float* feature,
                            float* depth, const crnerf_render_opts* opts, void* stream) {
  return render_pass_impl(packed, operand, rays, view_dir, z_vals, noise, n_rays, n_samples, n_freq_xyz,
                          n_freq_dir, weights, feature, depth, nullptr, nullptr, opts, stream);
}

size_t crnerf_render_acts_bytes(int64_t n_points) { return render_acts_bytes(n_points); }
size_t crnerf_render_backward_weights_bytes(int e_xyz) { return bwd_packed_bytes(e_xyz); }
size_t crnerf_render_backward_scratch_bytes(int64_t n_points) { return bwd_scratch_bytes(n_points); }

int crnerf_render_backward(const crnerf_mlp_weights* w, int operand, const void* acts, const float* raw,
                           const float* z_vals, const float* noise, const float* g_feature,
                           const float* g_weights, const float* g_depth, int n_rays, int n_samples,
                           void* bwd_weights, void* scratch, float* const* grad_weight, float* const* grad_bias,
                           void* stream) {
  int rc = device_check();
  if (rc) return rc;
  CRNERF_REQUIRE(w && acts && raw && z_vals && bwd_weights && scratch && grad_weight && grad_bias, "null argument");
  CRNERF_REQUIRE(operand == 0 || operand == 1, "the backward takes operand 0 (fp16) or 1 (bf16)");
  CRNERF_REQUIRE(n_rays >= 0 && n_samples >= 16 && n_samples <= 1024, "n_samples=%d unsupported (16..1024)", n_samples);
  CRNERF_REQUIRE(w->e_xyz >= 3 && w->e_xyz <= kMaxExyzPublic && w->e_dir >= 0 && w->e_dir <= 32, "embedding widths out of range");
  for (int i = 0; i < 12; ++i)
    CRNERF_REQUIRE(w->weight[i] && grad_weight[i] && grad_bias[i], "weight / gradient pointer %d is null", i);
  if (n_rays == 0) return CRNERF_OK;
  return render_backward(w, operand, acts, raw, z_vals, noise, g_feature, g_weights, g_depth, n_rays, n_samples,
                         bwd_weights, scratch, grad_weight, grad_bias, (cudaStream_t)stream);
}

int crnerf_render_pass_train(const void* packed, int operand, const float* rays, const float* view_dir,
                             const float* z_vals, const float* noise, int n_rays, int n_samples,
                             int n_freq_xyz, int n_freq_dir, float* weights, float* feature,
                             float* depth, void* acts, float* raw, void* stream) {
  CRNERF_REQUIRE(acts && raw, "acts and raw are required (use crnerf_render_pass for inference)");
  CRNERF_REQUIRE((reinterpret_cast<uintptr_t>(acts) & 15) == 0, "acts must be 16-byte aligned");
  CRNERF_REQUIRE(operand == 0 || operand == 1, "the training forward takes operand 0 (fp16) or 1 (bf16)");
  int rc = render_pass_impl(packed, operand, rays, view_dir, z_vals, noise, n_rays, n_samples, n_freq_xyz,
                            n_freq_dir, weights, feature, depth, acts, raw, nullptr, stream);
  if (rc) return rc;
  return render_acts_zero_tail(acts, (int64_t)n_rays * n_samples, (cudaStream_t)stream);
}

int crnerf_render_pass_train_opts(const void* packed, int operand, const float* rays, const float* view_dir,
                                  const float* z_vals, const float* noise, int n_rays, int n_samples,
                                  int n_freq_xyz, int n_freq_dir, float* weights, float* feature,
                                  float* depth, void* acts, float* raw, const crnerf_render_opts* opts,
                                  void* stream) {
  CRNERF_REQUIRE(acts && raw, "acts and raw are required (use crnerf_render_pass for inference)");
  CRNERF_REQUIRE((reinterpret_cast<uintptr_t>(acts) & 15) == 0, "acts must be 16-byte aligned");
  CRNERF_REQUIRE(operand == 0 || operand == 1, "the training forward takes operand 0 (fp16) or 1 (bf16)");
  int rc = render_pass_impl(packed, operand, rays, view_dir, z_vals, noise, n_rays, n_samples, n_freq_xyz,
                            n_freq_dir, weights, feature, depth, acts, raw, opts, stream);
  if (rc) return rc;
  return render_acts_zero_tail(acts, (int64_t)n_rays * n_samples, (cudaStream_t)stream);
}

int crnerf_composite_backward(const float* raw, const float* z_vals, const float* noise,
                              const float* g_feature, const float* g_weights, const float* g_depth,
                              int n_rays, int n_samples, float* d_rgb_pre, float* d_sigma_pre,
                              void* stream) {
  int rc = device_check();
  if (rc) return rc;
  return composite_backward(raw, z_vals, noise, g_feature, g_weights, g_depth, n_rays, n_samples,
                            d_rgb_pre, d_sigma_pre, (cudaStream_t)stream);
}

int crnerf_adam_step(int n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                     float* const* exp_avg_sq, const int64_t* numel, const float* step, const float* lr_dev,
                     double lr, double beta1, double beta2, double eps, double weight_decay, int maximize,
                     void* stream) {
  int rc = device_check();
  if (rc) return rc;
  return adam_step(n_tensors, params, grads, exp_avg, exp_avg_sq, numel, step, lr_dev, lr, beta1, beta2, eps,
                   weight_decay, maximize, (cudaStream_t)stream);
}

size_t crnerf_relu_bias_grad_scratch_floats(int width) {
  return (size_t)8 * (size_t)num_sms() * (size_t)(width > 0 ? width : 0);   // one row of partials per block
}

int crnerf_relu_bias_grad(float* g, const void* act, int64_t n_points, int width, float* gb,
                          float* scratch, void* stream) {
  int rc = device_check();
  if (rc) return rc;
  return relu_bias_grad(g, act, n_points, width, gb, scratch, (cudaStream_t)stream);
}

static int render_pass_impl(const void* packed, int operand, const float* rays, const float* view_dir,
                            const float* z_vals, const float* noise, int n_rays, int n_samples,
                            int n_freq_xyz, int n_freq_dir, float* weights, float* feature,
                            float* depth, void* acts, float* raw_save, const crnerf_render_opts* opts,
                            void* stream) {
  int rc = device_check();
  if (rc) return rc;
  CRNERF_REQUIRE(packed && rays && z_vals && weights && feature && depth, "null argument");
  CRNERF_REQUIRE(operand >= 0 && operand <= 2, "operand must be 0 (fp16), 1 (bf16) or 2 (fp16x3)");
  CRNERF_REQUIRE(n_rays >= 0, "n_rays must be non-negative");
  CRNERF_REQUIRE(n_samples >= 16 && n_samples <= 4096,
                 "n_samples=%d unsupported (16 <= n_samples <= 4096)", n_samples);
  CRNERF_REQUIRE(n_freq_xyz >= 0 && n_freq_xyz <= 15 && n_freq_dir >= 0 && n_freq_dir <= 4,
                 "positional-encoding bands out of range (xyz <= 15, dir <= 4)");
  CRNERF_REQUIRE((reinterpret_cast<uintptr_t>(rays) & 15) == 0, "rays must be 16-byte aligned");
  if (n_rays == 0) return CRNERF_OK;
  RenderArgs a{};
  a.packed = packed;
  a.operand = operand;
  a.e_xyz = 3 + 6 * n_freq_xyz;
  a.e_dir = 3 + 6 * n_freq_dir;
  a.rays = rays;
  a.view_dir = view_dir;
  a.z_vals = z_vals;
  a.noise = noise;
  a.n_points = (int64_t)n_rays * n_samples;
  a.n_rays = n_rays;
  a.n_samples = n_samples;
  a.n_freq_xyz = n_freq_xyz;
  a.n_freq_dir = n_freq_dir;
  a.weights = weights;
  a.feature = feature;
  a.depth = depth;
  a.acts = acts;
  a.raw_save = raw_save;
  if (opts) {
    a.jitter = opts->xyz_jitter;
    a.chan_part = opts->channel_partials;
    a.overflow = opts->overflow_flag;
  }
  return launch_render(a, (cudaStream_t)stream);
}

int crnerf_mlp_forward(const void* packed, int operand, int e_xyz, int e_dir, const float* x,
                       int64_t n, int x_stride, int sigma_only, float* out, void* stream) {
  int rc = device_check();
  if (rc) return rc;
  CRNERF_REQUIRE(packed && x && out, "null argument");
  CRNERF_REQUIRE(operand >= 0 && operand <= 2, "operand must be 0 (fp16), 1 (bf16) or 2 (fp16x3)");
  CRNERF_REQUIRE(n >= 0, "n must be non-negative");
  CRNERF_REQUIRE(x_stride >= (sigma_only ? e_xyz : e_xyz + e_dir), "x_stride smaller than the row width");
  if (n == 0) return CRNERF_OK;
  RenderArgs a{};
  a.packed = packed;
  a.operand = operand;
  a.e_xyz = e_xyz;
  a.e_dir = e_dir;
  a.x = x;
  a.x_stride = x_stride;
  a.sigma_only = sigma_only;
  a.n_points = n;
  a.n_samples = 1;
  a.raw = out;
  return launch_render(a, (cudaStream_t)stream);
}

int crnerf_generate_rays(const float* intrinsics_host, const float* c2w_host, float near, float far,
                         int height, int width, float* rays, void* stream) {
  int rc = device_check();
  if (rc) return rc;
  return generate_rays(intrinsics_host, c2w_host, near, far, height, width, rays, (cudaStream_t)stream);
}

int crnerf_grid_patch(const float* lin_w, const float* lin_h, int grid, float img_w, float img_h, float scale,
                      float h_offset, float w_offset, const float* all_rays, const float* all_rgbs,
                      int64_t n_cache_rows, float image_offset, float* rays, int64_t* ts, float* rgbs,
                      int64_t* rgb_idx, float* uv_sample, int32_t* status_dev, void* stream) {
  int rc = device_check();
  if (rc) return rc;
  return grid_patch(lin_w, lin_h, grid, img_w, img_h, scale, h_offset, w_offset, all_rays, all_rgbs,
                    (long long)n_cache_rows, image_offset, rays, (long long*)ts, rgbs,
                    (long long*)rgb_idx, uv_sample, status_dev, (cudaStream_t)stream);
}

int crnerf_rgb_to_u8(const float* rgb, int64_t n_pixels, uint8_t* out, void* stream) {
  int rc = device_check();
  if (rc) return rc;
  return rgb_to_u8(rgb, n_pixels, out, (cudaStream_t)stream);
}

size_t crnerf_encoder_packed_bytes(void) { return encoder_packed_bytes(); }
size_t crnerf_encoder_scratch_bytes(int height, int width) {
  return height >= 8 && width >= 8 ? encoder_scratch_bytes(height, width) : 0;
}
int crnerf_encoder_pack(const crnerf_encoder_weights* w, void* packed, size_t packed_bytes, void* stream) {
  int rc = device_check();
  if (rc) return rc;
  return encoder_pack(w, packed, packed_bytes, (cudaStream_t)stream);
}
int crnerf_encoder_forward(const void* packed, const float* img, int height, int width, float* out,
                           void* scratch, size_t scratch_bytes, void* stream) {
  int rc = device_check();
  if (rc) return rc;
  return encoder_forward(packed, img, height, width, out, scratch, scratch_bytes, (cudaStream_t)stream);
}
size_t crnerf_encoder_tape_bytes(int height, int width) {
  return height >= 8 && width >= 8 ? encoder_tape_bytes(height, width) : 0;
}
size_t crnerf_encoder_backward_scratch_bytes(int height, int width) {
  return height >= 8 && width >= 8 ? encoder_backward_scratch_bytes(height, width) : 0;
}
int crnerf_encoder_forward_train(const void* packed, const float* img, int height, int width, float* out,
                                 void* tape, size_t tape_bytes, void* stream) {
  int rc = device_check();
  if (rc) return rc;
  return encoder_forward_train(packed, img, height, width, out, tape, tape_bytes, (cudaStream_t)stream);
}
int crnerf_encoder_backward(const void* packed, const float* img, int height, int width, const float* out,
                            const float* grad_out, const void* tape, const crnerf_encoder_grads* grads,
                            float* grad_img, void* scratch, size_t scratch_bytes, void* stream) {
  int rc = device_check();
  if (rc) return rc;
  CRNERF_REQUIRE(grads, "null argument");
  return encoder_backward(packed, img, height, width, out, grad_out, tape, grads->weight, grads->bias, grad_img,
                          scratch, scratch_bytes, (cudaStream_t)stream);
}

int crnerf_ray_loss_forward(const float* rgb_coarse, const float* rgb_fine, const float* targets,
                            const float* mask, int64_t n_rays, float coef, float size_delta,
                            float digit_delta, float* out4, float* scratch, void* stream) {
  int rc = device_check();
  if (rc) return rc;
  return ray_loss_forward(rgb_coarse, rgb_fine, targets, mask, n_rays, coef, size_delta, digit_delta, out4,
                          scratch, (cudaStream_t)stream);
}

int crnerf_ray_loss_backward(const float* rgb_coarse, const float* rgb_fine, const float* targets,
                             const float* mask, int64_t n_rays, float coef, float size_delta,
                             float digit_delta, const float* grad_out4, float* g_rgb_coarse,
                             float* g_rgb_fine, float* g_mask, void* stream) {
  int rc = device_check();
  if (rc) return rc;
  return ray_loss_backward(rgb_coarse, rgb_fine, targets, mask, n_rays, coef, size_delta, digit_delta,
                           grad_out4, g_rgb_coarse, g_rgb_fine, g_mask, (cudaStream_t)stream);
}

int crnerf_pair_loss_forward(int n_terms, const float* const* a, const float* const* b,
                             const int64_t* n, const int* mode, const float* scale, float* out,
                             float* scratch, void* stream) {
  int rc = device_check();
  if (rc) return rc;
  return pair_loss_forward(n_terms, a, b, n, mode, scale, out, scratch, (cudaStream_t)stream);
}

int crnerf_pair_loss_backward(int n_terms, const float* const* a, const float* const* b,
                              const int64_t* n, const int* mode, const float* scale,
                              const float* grad_out, float* const* ga, float* const* gb, void* stream) {
  int rc = device_check();
  if (rc) return rc;
  return pair_loss_backward(n_terms, a, b, n, mode, scale, grad_out, ga, gb, (cudaStream_t)stream);
}

size_t crnerf_loss_scratch_floats(void) { return device_check() == CRNERF_OK ? loss_scratch_floats() : 0; }

int crnerf_mask_sample_forward(const float* pred, int channels, int h, int w, int H, int W,
                               const int64_t* idx, int64_t n, float* out, void* stream) {
  int rc = device_check();
  if (rc) return rc;
  return mask_sample_forward(pred, channels, h, w, H, W, idx, n, out, (cudaStream_t)stream);
}

int crnerf_mask_sample_backward(const float* g_out, int channels, int h, int w, int H, int W,
                                const int64_t* idx, int64_t n, float* g_pred, void* stream) {
  int rc = device_check();
  if (rc) return rc;
  return mask_sample_backward(g_out, channels, h, w, H, W, idx, n, g_pred, (cudaStream_t)stream);
}

int crnerf_pos_embed(const float* x, int64_t n, int n_freqs, float* out, void* stream) {
  int rc = device_check();
  if (rc) return rc;
  return pos_embed(x, n, n_freqs, out, (cudaStream_t)stream);
}

int crnerf_coarse_z(const float* rays, const float* t_steps, const float* perturb_rand, int n_rays,
                    int n_samples, int use_disp, float* z_vals, void* stream) {
  int rc = device_check();
  if (rc) return rc;
  return coarse_z(rays, t_steps, perturb_rand, n_rays, n_samples, use_disp, z_vals,
                  (cudaStream_t)stream);
}

int crnerf_sample_pdf_merge(const float* z_coarse, const float* weights_coarse, const float* u,
                            int64_t u_stride, int n_rays, int n_samples, int n_importance,
                            float eps, float* z_fine, float* z_new, void* stream) {
  int rc = device_check();
  if (rc) return rc;
  CRNERF_REQUIRE(n_samples >= 3, "sample_pdf needs at least 3 coarse samples");
  return sample_pdf(z_coarse, weights_coarse, u, u_stride, n_rays, n_samples - 2, n_importance, eps,
                    z_new, z_fine, true, (cudaStream_t)stream);
}

int crnerf_sample_pdf(const float* bins, const float* weights, const float* u, int64_t u_stride,
                      int n_rays, int m, int n_importance, float eps, float* samples,
                      void* stream) {
  int rc = device_check();
  if (rc) return rc;
  return sample_pdf(bins, weights, u, u_stride, n_rays, m, n_importance, eps, samples, nullptr,
                    false, (cudaStream_t)stream);
}

size_t crnerf_style_scratch_floats(int64_t n_pixels) { return style_scratch_floats(n_pixels); }

int crnerf_style_forward(const crnerf_style_weights* w, const float* content, int64_t n_pixels,
                         int64_t c_pix_stride, int64_t c_ch_stride, const float* style,
                         int64_t n_style_pixels, int64_t s_pix_stride, int64_t s_ch_stride,
                         float* rgb, float* transmatrix, float* fused, float* scratch,
                         void* stream) {
  int rc = device_check();
  if (rc) return rc;
  return style_forward(w, content, n_pixels, c_pix_stride, c_ch_stride, style, n_style_pixels,
                       s_pix_stride, s_ch_stride, nullptr, 0, rgb, transmatrix, fused, scratch,
                       (cudaStream_t)stream);
}

int crnerf_style_forward_sums(const crnerf_style_weights* w, const float* content, int64_t n_pixels,
                              int64_t c_pix_stride, int64_t c_ch_stride, const float* style,
                              int64_t n_style_pixels, int64_t s_pix_stride, int64_t s_ch_stride,
                              const float* content_sum_partials, int n_partials, float* rgb,
                              float* transmatrix, float* fused, float* scratch, void* stream) {
  int rc = device_check();
  if (rc) return rc;
  return style_forward(w, content, n_pixels, c_pix_stride, c_ch_stride, style, n_style_pixels,
                       s_pix_stride, s_ch_stride, content_sum_partials, n_partials, rgb, transmatrix, fused,
                       scratch, (cudaStream_t)stream);
}

size_t crnerf_style_aux_floats(void) { return style_aux_floats(); }
size_t crnerf_style_backward_grads_floats(void) { return style_backward_grads_floats(); }
size_t crnerf_style_backward_scratch_floats(int64_t n_pixels, int64_t n_style_pixels) {
  return style_backward_scratch_floats(n_pixels, n_style_pixels);
}
void crnerf_style_backward_layout(int64_t* offsets22) { style_backward_layout(offsets22); }

int crnerf_style_forward_train(const crnerf_style_weights* w, const float* content, int64_t n_pixels,
                               int64_t c_pix_stride, int64_t c_ch_stride, const float* style,
                               int64_t n_style_pixels, int64_t s_pix_stride, int64_t s_ch_stride,
                               const float* content_sum_partials, int n_partials, float* rgb, float* aux,
                               float* scratch, void* stream) {
  int rc = device_check();
  if (rc) return rc;
  return style_forward_train(w, content, n_pixels, c_pix_stride, c_ch_stride, style, n_style_pixels, s_pix_stride,
                             s_ch_stride, content_sum_partials, n_partials, rgb, aux, scratch, (cudaStream_t)stream);
}

int crnerf_style_backward(const crnerf_style_weights* w, const float* content, int64_t n_pixels,
                          int64_t c_pix_stride, int64_t c_ch_stride, const float* style, int64_t n_style_pixels,
                          int64_t s_pix_stride, int64_t s_ch_stride, const float* aux, const float* g_rgb,
                          float* g_content, float* g_style, float* grads, float* scratch, void* stream) {
  int rc = device_check();
  if (rc) return rc;
  return style_backward(w, content, n_pixels, c_pix_stride, c_ch_stride, style, n_style_pixels, s_pix_stride,
                        s_ch_stride, aux, g_rgb, g_content, g_style, grads, scratch, (cudaStream_t)stream);
}

int crnerf_sum_rows(const float* parts, int n_parts, int len, float* out, void* stream) {
  int rc = device_check();
  if (rc) return rc;
  return sum_rows(parts, n_parts, len, out, (cudaStream_t)stream);
}

int crnerf_cnn_forward(const crnerf_cnn_weights* w, const float* x, int64_t n_pixels,
                       int64_t pix_stride, int64_t ch_stride, float* out, float* scratch,
                       void* stream) {
  int rc = device_check();
  if (rc) return rc;
  return cnn_forward(w, x, n_pixels, pix_stride, ch_stride, out, scratch, (cudaStream_t)stream);
}

int crnerf_style_stats1(const float* content, int64_t n_pixels, int64_t pix_stride,
                        int64_t ch_stride, float* sums, float* scratch, void* stream) {
  int rc = device_check();
  if (rc) return rc;
  CRNERF_REQUIRE(content && sums && scratch && n_pixels >= 1, "bad argument");
  return style_stats1(content, n_pixels, pix_stride, ch_stride, sums, scratch, (cudaStream_t)stream);
}

int crnerf_style_stats2(const crnerf_style_weights* w, const float* content, int64_t n_pixels,
                        int64_t pix_stride, int64_t ch_stride, const float* mean, float* gram,
                        float* scratch, void* stream) {
  int rc = device_check();
  if (rc) return rc;
  CRNERF_REQUIRE(w && content && mean && gram && scratch && n_pixels >= 1, "bad argument");
  // partials go to the Gram-partial region of the scratch buffer (offset = kMaxBlocks*64 floats)
  return style_stats2(w->cnet, content, n_pixels, pix_stride, ch_stride, mean, gram,
                      scratch + 296 * 64, 1.f, (cudaStream_t)stream);
}

int crnerf_style_apply(const crnerf_style_weights* w, const float* content, int64_t n_pixels,
                       int64_t pix_stride, int64_t ch_stride, const float* mean,
                       const float* gram_normalised, const float* style, int64_t n_style_pixels,
                       int64_t s_pix_stride, int64_t s_ch_stride, float* rgb, float* transmatrix,
                       float* scratch, void* stream) {
  int rc = device_check();
  if (rc) return rc;
  CRNERF_REQUIRE(w && content && mean && gram_normalised && style && rgb && scratch, "null argument");
  return style_finish(w, content, n_pixels, pix_stride, ch_stride, mean, gram_normalised, style,
                      n_style_pixels, s_pix_stride, s_ch_stride, rgb, transmatrix, nullptr, scratch,
                      (cudaStream_t)stream);
}

}  // extern "C"

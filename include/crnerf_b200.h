/* crnerf_b200 - C ABI of the B200-native CR-NeRF volume-rendering path.
 *
 * Drop-in boundary for the hot path of YifYang993/CR-NeRF-PyTorch
 * (models/rendering.py::render_rays_cross_ray and the NeRF_sigma / style_net /
 * NeuralRenderer forwards it feeds).  The reference is pure Python/PyTorch and has
 * no FFI of its own, so these entry points are what a ctypes binding placed at
 * the reference's call sites would bind; the repo's own
 * Python mirror (cr-nerf-pytorch_b200/models/) is exactly such a binding.
 *
 * Conventions
 *   - All pointers are DEVICE pointers on the current CUDA device unless the
 *     name ends in _host; tensors are dense row-major fp32 unless stated.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *     Every call only enqueues work on that stream: no host synchronisation, no
 *     internal streams, no allocation (callers own every buffer; scratch sizes
 *     are queryable).
 *   - Return value: 0 on success, negative crnerf_status on error;
 *     crnerf_last_error() returns a thread-local message.  There is NO CPU
 *     fallback: on a machine without an sm_100 GPU every compute call fails with
 *     CRNERF_ERR_DEVICE.
 *   - Thread safety: calls may be made from any host thread; a packed-weights
 *     buffer may be shared by concurrent calls (read-only).
 */
#ifndef CRNERF_B200_H_
#define CRNERF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRNERF_ABI_VERSION 1

typedef enum {
  CRNERF_OK = 0,
  CRNERF_ERR_ARG = -1,     /* bad shape / null pointer / unsupported configuration */
  CRNERF_ERR_DEVICE = -2,  /* no sm_100 device, or a CUDA runtime error */
  CRNERF_ERR_RANGE = -3,   /* a weight does not fit the operand format (see pack) */
  CRNERF_ERR_KERNEL = -4   /* device-side protocol timeout (internal error) */
} crnerf_status;

/* Operand format of the tensor-core MLP.  Accumulation is always fp32.
 *   FP16  : 10-bit mantissa (same as TF32).  Meets the 1e-4 parity bar at default-init weights
 *           (measured 3-4e-5); on trained weights, whose layers cancel large terms, the operand
 *           rounding shows as 1-2e-3 relative (PSNR-level parity: >= 70 dB against the fp32
 *           reference).  |w| and activations must stay below 65504 (reported, never silent).
 *   BF16  : 7-bit mantissa, fp32 range - PSNR-level parity only; the training format.
 *   FP16X3: every operand split as hi + lo (two fp16 values), every product three MMAs
 *           (hi*hi + lo*hi + hi*lo): fp32-class accuracy (measured <= 1e-5 on trained weights) at
 *           three times the tensor work.  Inference only. */
typedef enum { CRNERF_OPERAND_FP16 = 0, CRNERF_OPERAND_BF16 = 1, CRNERF_OPERAND_FP16X3 = 2 } crnerf_operand;

const char* crnerf_last_error(void);
int crnerf_abi_version(void);
/* 1 if the current device can run the kernels (compute capability 10.x). */
int crnerf_device_ok(void);
/* number of kernels this library has launched since load (all threads) */
uint64_t crnerf_launch_count(void);

/* ------------------------------------------------------------------------
 * NeRF_sigma weights (reference models/nerf.py:116-154, state_dict keys
 * xyz_encoding_{1..8}.0, xyz_encoding_final, dir_encoding.0, static_rgb.0,
 * static_sigma.0).  nn.Linear layout: weight (out,in) row-major, bias (out).
 * Index: 0..7 trunk, 8 final, 9 dir, 10 rgb, 11 sigma.
 * ---------------------------------------------------------------------- */
typedef struct {
  const float* weight[12];
  const float* bias[12];
  int32_t e_xyz; /* in_channels_xyz, <= 96 (93 for N_emb_xyz=15) */
  int32_t e_dir; /* in_channels_dir, <= 32 (27 for N_emb_dir=4)  */
} crnerf_mlp_weights;

/* Bytes of the packed form (16-bit swizzled weight image(s) + fp32 bias blob + the kernel's
 * program tables): crnerf_mlp_packed_bytes for operands 0 / 1, _op for any operand
 * (CRNERF_OPERAND_FP16X3 stores two images, W_hi and W_lo = fp16(W - W_hi)). */
size_t crnerf_mlp_packed_bytes(int e_xyz, int e_dir);
size_t crnerf_mlp_packed_bytes_op(int e_xyz, int e_dir, int operand);
/* Pack once per weight version.  `status_dev` (int32, device, may be NULL)
 * receives 1 if some |w| exceeded the operand format's finite range (the value
 * is clamped); the Python mirror turns that into an error. */
int crnerf_mlp_pack(const crnerf_mlp_weights* w, int operand, void* packed, size_t packed_bytes,
                    int32_t* status_dev, void* stream);

/* ------------------------------------------------------------------------
 * One volume-rendering pass = the reference's `inference` closure
 * (models/rendering.py:82-145) with PosEmbedding (models/nerf.py:17-30) and
 * NeRF_sigma.forward (models/nerf.py:157-182) fused in: for every ray,
 * embed xyz = o + d*z and the view direction, run the MLP on every sample,
 * alpha-composite.  No per-point intermediate is written to memory.
 *   rays      (n_rays, 8)  [o3, d3, near, far]   (rendering.py:151-153)
 *   view_dir  (n_rays, 3) or NULL -> rays[:,3:6]  (kwargs['view_dir'], :155)
 *   z_vals    (n_rays, n_samples) sorted depths
 *   noise     (n_rays, n_samples) already scaled by noise_std, or NULL (:125)
 * outputs
 *   weights   (n_rays, n_samples), feature (n_rays, 64), depth (n_rays)
 * `operand` must be the crnerf_operand the weights were packed with, and the
 * weights must have been packed for e_xyz = 3+6*n_freq_xyz, e_dir = 3+6*n_freq_dir.
 * 16 <= n_samples <= 4096.  n_freq_xyz <= 15, n_freq_dir <= 4 (PosEmbedding(L-1, L)).
 * ---------------------------------------------------------------------- */
int crnerf_render_pass(const void* packed, int operand, const float* rays, const float* view_dir,
                       const float* z_vals, const float* noise, int n_rays, int n_samples,
                       int n_freq_xyz, int n_freq_dir, float* weights, float* feature,
                       float* depth, void* stream);

/* The same pass with optional extras (every member may be NULL):
 *   xyz_jitter        (n_rays*n_samples, 3) added to xyz = o + d*z before the embedding: the
 *                     reference's args.pertubeCord path, `xyz_ += pertube_ratio * rand(...)`
 *                     (models/rendering.py:102-104); the caller draws and scales the noise.
 *   channel_partials  out, (crnerf_render_partial_rows(n_rays, n_samples), 64): partial column
 *                     sums of `feature`, one row per (CTA, tile group), written in full (no
 *                     zeroing needed).  Their sum is the per-channel sum over all rays that
 *                     MulLayer.forward's mean needs (models/linearStyleTransfer.py:62), so the
 *                     cross-ray block does not have to read the feature map for it
 *                     (crnerf_style_forward_sums / crnerf_style_stats1_from_partials).
 *   overflow_flag     int32 visible to the device (device memory or mapped pinned host memory),
 *                     set to 1 if an fp16 operand (activation or input) reached the format's
 *                     finite limit 65504 and was clamped: the result is then NOT within
 *                     tolerance and the caller should switch to CRNERF_OPERAND_BF16.  Never
 *                     written otherwise; bf16 passes never write it. */
typedef struct {
  const float* xyz_jitter;
  float* channel_partials;
  int32_t* overflow_flag;
} crnerf_render_opts;
int crnerf_render_partial_rows(int n_rays, int n_samples);
int crnerf_render_pass_opts(const void* packed, int operand, const float* rays, const float* view_dir,
                            const float* z_vals, const float* noise, int n_rays, int n_samples,
                            int n_freq_xyz, int n_freq_dir, float* weights, float* feature,
                            float* depth, const crnerf_render_opts* opts, void* stream);

/* ------------------------------------------------------------------------
 * Training step (train_mask_grid_sample.py:186-197 calls render_rays_cross_ray under autograd).
 * crnerf_render_pass_train = crnerf_render_pass that additionally stores what the backward
 * needs, so nothing is recomputed and no (points x width) fp32 tensor is ever built:
 *   acts  crnerf_render_acts_bytes(n_rays*n_samples) bytes of 16-bit values (the operand format)
 *         in the backward kernels' operand layout: per 128-point tile and 64-feature slab a 16 KB
 *         block of 128-byte rows (one per point, 16-byte chunks XOR-swizzled with row % 8).
 *         Slots: 0..8 = outputs of xyz_encoding_1..8 (post-ReLU) and xyz_encoding_final (256
 *         wide each), 9 = dir_encoding output (post-ReLU, 128 wide), 10 = the embedding tile
 *         (columns [0, e_xyz) xyz embedding, [96, 96 + e_dir) direction embedding).
 *   raw   65 * n_points fp32: the sigmoid features as (n_points, 64) rows, then the n_points softplus
 *         sigmas (models/nerf.py:180-181's columns, split so that a point's features are whole
 *         32-byte sectors); consumed by crnerf_render_backward only
 * crnerf_composite_backward is the backward of rendering.py:116-143 (raw here = (n_points, 65) rows
 * [features | sigma] as NeRF_sigma.forward returns them): from the gradients of
 * feature (n_rays,64), weights (n_rays,n_samples), depth (n_rays) (each may be NULL) to the
 * gradients of the pre-sigmoid features d_rgb_pre (n_points,64) and of the pre-softplus
 * density d_sigma_pre (n_points).  n_samples <= 1024.
 * crnerf_render_backward is the whole backward of one pass on the tensor cores: composite
 * backward, then for every layer the weight gradient dW += G^T X and the input gradient
 * G' = (G W) * relu'(X) as tcgen05 GEMMs over the saved activations (no library GEMM, no fp32
 * (points x width) tensor).  `w` holds the CURRENT fp32 weights (their transposes are re-packed
 * into `bwd_weights`, crnerf_render_backward_weights_bytes(e_xyz) bytes); `scratch` is
 * crnerf_render_backward_scratch_bytes(n_points) bytes; grad_weight[i] / grad_bias[i] (fp32, the
 * shapes of w->weight[i] / w->bias[i], i in the order of crnerf_mlp_weights): weight gradients are
 * WRITTEN (every element; per-CTA partial products summed in a fixed order - run-to-run identical),
 * bias gradients are ACCUMULATED with atomics (zero them first).  16 <= n_samples <= 1024.
 * ---------------------------------------------------------------------- */
size_t crnerf_render_acts_bytes(int64_t n_points);
size_t crnerf_render_backward_weights_bytes(int e_xyz);
size_t crnerf_render_backward_scratch_bytes(int64_t n_points);
int crnerf_render_backward(const crnerf_mlp_weights* w, int operand, const void* acts, const float* raw,
                           const float* z_vals, const float* noise, const float* g_feature,
                           const float* g_weights, const float* g_depth, int n_rays, int n_samples,
                           void* bwd_weights, void* scratch, float* const* grad_weight,
                           float* const* grad_bias, void* stream);
int crnerf_render_pass_train(const void* packed, int operand, const float* rays, const float* view_dir,
                             const float* z_vals, const float* noise, int n_rays, int n_samples,
                             int n_freq_xyz, int n_freq_dir, float* weights, float* feature,
                             float* depth, void* acts, float* raw, void* stream);
/* crnerf_render_pass_train with the options of crnerf_render_pass_opts (args.pertubeCord jitter of the
 * sample positions, models/rendering.py:102-104, in the training step; the saved embedding tile is the
 * jittered one, so the backward needs nothing else). */
int crnerf_render_pass_train_opts(const void* packed, int operand, const float* rays, const float* view_dir,
                                  const float* z_vals, const float* noise, int n_rays, int n_samples,
                                  int n_freq_xyz, int n_freq_dir, float* weights, float* feature,
                                  float* depth, void* acts, float* raw, const crnerf_render_opts* opts,
                                  void* stream);
int crnerf_composite_backward(const float* raw, const float* z_vals, const float* noise,
                              const float* g_feature, const float* g_weights, const float* g_depth,
                              int n_rays, int n_samples, float* d_rgb_pre, float* d_sigma_pre,
                              void* stream);

/* Backward of a layer's ReLU + bias: g (n_points, width) fp32 is multiplied in place by
 * (act > 0), act being that layer's slice of `acts` (16-bit, same shape; NULL = no activation),
 * and gb (width) receives the column sums = the bias gradient.  width in {64,128,256};
 * scratch: crnerf_relu_bias_grad_scratch_floats(width) floats (sized from the current device's
 * SM count). */
size_t crnerf_relu_bias_grad_scratch_floats(int width);
int crnerf_relu_bias_grad(float* g, const void* act, int64_t n_points, int width, float* gb,
                          float* scratch, void* stream);

/* ---- loss + mask tail of the training step ------------------------------
 * All pointers are device pointers unless marked HOST.  `scratch` for the two *_loss_forward
 * calls: crnerf_loss_scratch_floats() floats whose first 16 bytes are zero before the first
 * use (the kernels leave them zero); one scratch per stream.
 *
 * CRNeRFLoss.forward's per-ray terms (losses.py:62-76) in one launch:
 *   out4[0] c_l  = coef * 0.5 * mean((1-mask) * (rgb_coarse-targets)^2)     (losses.py:63-66)
 *   out4[1] f_l  = the same on rgb_fine (0 if rgb_fine == NULL)              (losses.py:70-74)
 *   out4[2] r_ms = coef * size_delta * mean(mask^2)                          (mask_regularize, :80-84)
 *   out4[3] r_md = coef * digit_delta * mean(1/((mask-0.5)^2 + 0.02))        (:86-87)
 * rgb_* / targets (n_rays,3), mask (n_rays) or NULL (then (1-mask) = 1 and r_ms = r_md = 0). */
int crnerf_ray_loss_forward(const float* rgb_coarse, const float* rgb_fine, const float* targets,
                            const float* mask, int64_t n_rays, float coef, float size_delta,
                            float digit_delta, float* out4, float* scratch, void* stream);
/* Gradients of sum_k grad_out4[k]*out4[k] (grad_out4 == NULL: all ones).  c_l sees the mask
 * detached (losses.py:64); g_* may be NULL to skip an output. */
int crnerf_ray_loss_backward(const float* rgb_coarse, const float* rgb_fine, const float* targets,
                             const float* mask, int64_t n_rays, float coef, float size_delta,
                             float digit_delta, const float* grad_out4, float* g_rgb_coarse,
                             float* g_rgb_fine, float* g_mask, void* stream);
/* Up to 4 embedding terms per launch, out[k] = scale[k] * mean(f(a_k, b_k)) over n[k] elements:
 *   mode 0: a^2           kl_a = _l2_regularize(a_embedded)                  (losses.py:53, :91-94)
 *   mode 1: |a - b|       rec_a_random, L1 form                              (losses.py:57)
 *   mode 2: (a - b)^2     rec_a_random MSE form / content_constraint         (losses.py:56, :68)
 * a, b, n, mode, scale are HOST arrays of n_terms entries (a[k], b[k] device pointers). */
int crnerf_pair_loss_forward(int n_terms, const float* const* a, const float* const* b,
                             const int64_t* n, const int* mode, const float* scale, float* out,
                             float* scratch, void* stream);
/* ga[k] / gb[k] (device, n[k] floats, either may be NULL) receive grad_out[k] * d out[k] / d a_k, b_k. */
int crnerf_pair_loss_backward(int n_terms, const float* const* a, const float* const* b,
                              const int64_t* n, const int* mode, const float* scale,
                              const float* grad_out, float* const* ga, float* const* gb, void* stream);
size_t crnerf_loss_scratch_floats(void);

/* The mask lookup of NeRFSystem.forward (train_mask_grid_sample.py:171-175):
 *   interpolate(pred (1,C,h,w), size=(H,W), mode='bilinear', align_corners=False)
 *   -> rearrange '1 n h w -> (h w) n' -> [rgb_idx]
 * evaluated only at the n sampled pixels: out (n, C).  idx (n) int64 flat pixel indices in
 * [0, H*W), or NULL for every pixel in order (validation; then n == H*W). */
int crnerf_mask_sample_forward(const float* pred, int channels, int h, int w, int H, int W,
                               const int64_t* idx, int64_t n, float* out, void* stream);
/* g_pred (C,h,w) = the adjoint scatter of g_out (n, C) (zeroed inside, fp32 atomics). */
int crnerf_mask_sample_backward(const float* g_out, int channels, int h, int w, int H, int W,
                                const int64_t* idx, int64_t n, float* g_pred, void* stream);

/* NeRF_sigma.forward on pre-embedded rows (models/nerf.py:157-182):
 *   x (n, x_stride) with [0,e_xyz) xyz embedding, [e_xyz, e_xyz+e_dir) dir
 *   embedding; out (n, 65) = [sigmoid features(64) | softplus sigma].
 * If sigma_only != 0, x holds only the xyz embedding and out is (n, 1). */
int crnerf_mlp_forward(const void* packed, int operand, int e_xyz, int e_dir, const float* x,
                       int64_t n, int x_stride, int sigma_only, float* out, void* stream);

/* PosEmbedding.forward (models/nerf.py:17-30): x (n,3) -> out (n, 3+6*n_freqs),
 * [x, sin(2^0 x), cos(2^0 x), ...]. */
int crnerf_pos_embed(const float* x, int64_t n, int n_freqs, float* out, void* stream);

/* Camera rays of one pinhole frame, written straight into device memory as the (height*width, 8)
 * rows [o3, d3, near, far] the renderer consumes - replaces get_ray_directions + get_rays
 * (datasets/ray_utils.py:5-52) and the row assembly of the datasets
 * (datasets/phototourism_mask_grid_sample.py:300-307), i.e. the per-frame CPU meshgrid and the
 * 32 B/ray host->device copy of eval.py:279.  intrinsics_host = {fx, fy, cx, cy} (K[0,0], K[1,1],
 * K[0,2], K[1,2]); c2w_host = the 3x4 camera-to-world matrix, row-major; both HOST pointers.
 * Pixel (i = column, j = row) -> direction ((i-cx)/fx, -(j-cy)/fy, -1) rotated by c2w[:, :3] and
 * normalised, origin c2w[:, 3]; row index j*width + i. */
int crnerf_generate_rays(const float* intrinsics_host, const float* c2w_host, float near, float far,
                         int height, int width, float* rays, void* stream);

/* Grid-sampled training patch (datasets/phototourism_mask_grid_sample.py:241-275): the g x g
 * lattice (g = sqrt(batch_size)) of one training image at a random scale / offset, gathered from
 * the ray cache resident in GPU memory - replaces the DataLoader worker's index arithmetic, the
 * three fancy-index gathers and the per-step host->device copy of the batch.
 *   lin_w, lin_h   (grid) device: linspace(0, 1-1/img_w, g) and linspace(0, 1-1/img_h, g) (:246-247)
 *   scale, h_offset, w_offset    the three host-side uniform draws (:252-254)
 *   all_rays (n_cache_rows, 9) = [o3 d3 near far image_id]  (the .npy ray cache, :204-208),
 *   all_rgbs (n_cache_rows, 3); image_offset = sum of w*h of the images before this one (:266).
 *   The reference holds image sizes as fp32 (:197), so img_w / img_h / image_offset are floats and
 *   the cache row is (float(rgb_idx) + image_offset) rounded to fp32, then truncated - reproduced
 *   as is (it only rounds once the cache exceeds 2^24 rows).
 * Outputs, row r = j*g + i (lattice column i along the width, row j along the height):
 *   rays (g*g, 8), ts (g*g) int64, rgbs (g*g, 3), rgb_idx (g*g) int64 (pixel index w + h*img_w
 *   inside the image, computed in fp32 as the reference does), uv_sample (g*g, 2) = [h_sb, w_sb].
 * Bit-exact with the reference.  status_dev (may be NULL) is set to 1 if a row index leaves the
 * cache (inconsistent img_w/img_h/offset); such rows are left unwritten, never clamped. */
int crnerf_grid_patch(const float* lin_w, const float* lin_h, int grid, float img_w, float img_h, float scale,
                      float h_offset, float w_offset, const float* all_rays, const float* all_rgbs,
                      int64_t n_cache_rows, float image_offset, float* rays, int64_t* ts, float* rgbs,
                      int64_t* rgb_idx, float* uv_sample, int32_t* status_dev, void* stream);

/* Output stage of the eval loop (eval.py:295-297): rgb (3, n_pixels) planar fp32 (what
 * crnerf_style_forward writes) -> out (n_pixels, 3) interleaved uint8 = uint8(clip(x,0,1)*255),
 * so a frame leaves the GPU as 3 B/pixel instead of 12. */
int crnerf_rgb_to_u8(const float* rgb, int64_t n_pixels, uint8_t* out, void* stream);

/* Coarse depths (models/rendering.py:161-176): z = near*(1-t)+far*t (or the
 * disparity form).  t_steps (n_samples) is the caller's linspace(0,1,n_samples)
 * (rendering.py:161; passed in so the grid is bit-identical to the one the
 * reference builds on the same device).  If perturb_rand != NULL (n_rays,
 * n_samples, already multiplied by `perturb`) applies the stratified jitter
 * lower + (upper-lower)*perturb_rand. */
int crnerf_coarse_z(const float* rays, const float* t_steps, const float* perturb_rand, int n_rays,
                    int n_samples, int use_disp, float* z_vals, void* stream);

/* sample_pdf + merge (models/rendering.py:7-46 and :183-187):
 *   bins = midpoints of z_coarse, weights = weights_coarse[:,1:-1]; draws
 *   n_importance samples by inverse CDF at the uniforms u[ray*u_stride + i]
 *   (u_stride = n_importance for per-ray draws, rendering.py:30; u_stride = 0
 *   to share one row, the det case u = linspace(0,1,n_importance), :27-28) and
 *   writes z_fine (n_rays, n_samples+n_importance) = sort(cat(z_coarse, samples)).
 *   z_new (n_rays, n_importance), optional, receives the unsorted samples.
 *   n_samples + n_importance <= 4096. */
int crnerf_sample_pdf_merge(const float* z_coarse, const float* weights_coarse, const float* u,
                            int64_t u_stride, int n_rays, int n_samples, int n_importance,
                            float eps, float* z_fine, float* z_new, void* stream);

/* Stand-alone sample_pdf (models/rendering.py:7-46): bins (n, m+1), weights (n, m). */
int crnerf_sample_pdf(const float* bins, const float* weights, const float* u, int64_t u_stride,
                      int n_rays, int m, int n_importance, float eps, float* samples,
                      void* stream);

/* ------------------------------------------------------------------------
 * Cross-ray fusion + decoder = style_net.forward
 * (models/linearStyleTransfer.py:284-291 -> MulLayer.forward :58-90 ->
 *  CNN.forward :28-37 -> NeuralRenderer.forward nerf_decoder_stylenerf.py:279-291).
 * Feature maps are addressed as element (pixel p, channel c) at
 *   base[p * pix_stride + c * ch_stride]
 * so both the renderer's (N,64) rows (pix_stride 64, ch_stride 1: what the
 * callers' rearrange produces as a view) and contiguous NCHW
 * (pix_stride 1, ch_stride H*W) are read in place.
 * ---------------------------------------------------------------------- */
typedef struct {
  const float* conv_w[3]; /* convs.0 (128,64) convs.2 (64,128) convs.4 (32,64) */
  const float* conv_b[3];
  const float* fc_w; /* (1024,1024) */
  const float* fc_b; /* (1024)      */
} crnerf_cnn_weights;

typedef struct {
  crnerf_cnn_weights cnet, snet;
  const float* compress_w; /* (32,64) */
  const float* compress_b;
  const float* unzip_w; /* (64,32) */
  const float* unzip_b;
  const float* rgb_w; /* decoder.feat_2_rgb_list.0 (3,64) */
  const float* rgb_b;
} crnerf_style_weights;

/* scratch floats needed by crnerf_style_forward */
size_t crnerf_style_scratch_floats(int64_t n_pixels);
/* content (n_pixels x 64), style (n_style_pixels x 64) or NULL (type=="content":
 * decoder only) -> rgb written as (3, n_pixels) planar (= (1,3,H,W) contiguous).
 * Optionally writes transmatrix (32x32) and fused (64 x n_pixels planar). */
int crnerf_style_forward(const crnerf_style_weights* w, const float* content,
                         int64_t n_pixels, int64_t c_pix_stride, int64_t c_ch_stride,
                         const float* style, int64_t n_style_pixels, int64_t s_pix_stride,
                         int64_t s_ch_stride, float* rgb, float* transmatrix, float* fused,
                         float* scratch, void* stream);

/* Same, with the content map's channel sums supplied as partial sums: content_sum_partials
 * (n_partials, 64), rows adding up to the per-channel sums over all n_pixels - what
 * crnerf_render_pass_opts writes as channel_partials (concatenate the rows of several calls).
 * The block then reads the map twice (Gram pass, apply pass) instead of three times.  NULL / 0
 * = crnerf_style_forward. */
int crnerf_style_forward_sums(const crnerf_style_weights* w, const float* content, int64_t n_pixels,
                              int64_t c_pix_stride, int64_t c_ch_stride, const float* style,
                              int64_t n_style_pixels, int64_t s_pix_stride, int64_t s_ch_stride,
                              const float* content_sum_partials, int n_partials, float* rgb,
                              float* transmatrix, float* fused, float* scratch, void* stream);
/* out (len) = column sums of parts (n_parts, len), fixed order (e.g. channel_partials -> the
 * 64 channel sums of a rank's rays, the input of the sharded form's first all-reduce). */
int crnerf_sum_rows(const float* parts, int n_parts, int len, float* out, void* stream);

/* ---- the same block under autograd: the decode() of the training step
 * (train_mask_grid_sample.py:127-149) ----------------------------------------------------------
 * crnerf_style_forward_train = crnerf_style_forward_sums plus `aux`
 * (crnerf_style_aux_floats() floats: channel means, normalised Gram vectors, the two FC outputs and
 * transmatrix - everything the backward keeps from the forward; per-pixel activations are recomputed).
 * crnerf_style_backward: g_rgb (3, n_pixels) planar -> g_content (n_pixels, 64) rows, g_style
 * (n_style_pixels, 64) rows and the 22 parameter gradients in ONE flat buffer `grads`
 * (crnerf_style_backward_grads_floats() floats; crnerf_style_backward_layout fills the 22 offsets in
 * crnerf_style_weights order: cnet {conv_w[0..2], conv_b[0..2], fc_w, fc_b}, snet {same},
 * compress_w, compress_b, unzip_w, unzip_b, rgb_w, rgb_b).  fp32, deterministic (per-block partials
 * summed in block order).  scratch: crnerf_style_backward_scratch_floats(n_pixels, n_style_pixels). */
size_t crnerf_style_aux_floats(void);
size_t crnerf_style_backward_grads_floats(void);
size_t crnerf_style_backward_scratch_floats(int64_t n_pixels, int64_t n_style_pixels);
void crnerf_style_backward_layout(int64_t* offsets22);
int crnerf_style_forward_train(const crnerf_style_weights* w, const float* content, int64_t n_pixels,
                               int64_t c_pix_stride, int64_t c_ch_stride, const float* style,
                               int64_t n_style_pixels, int64_t s_pix_stride, int64_t s_ch_stride,
                               const float* content_sum_partials, int n_partials, float* rgb, float* aux,
                               float* scratch, void* stream);
int crnerf_style_backward(const crnerf_style_weights* w, const float* content, int64_t n_pixels,
                          int64_t c_pix_stride, int64_t c_ch_stride, const float* style, int64_t n_style_pixels,
                          int64_t s_pix_stride, int64_t s_ch_stride, const float* aux, const float* g_rgb,
                          float* g_content, float* g_style, float* grads, float* scratch, void* stream);

/* CNN.forward alone (models/linearStyleTransfer.py:28-37): x (n_pixels x 64) ->
 * out (1024) = fc(flatten(convs(x) convs(x)^T / n_pixels)).  Same scratch buffer. */
int crnerf_cnn_forward(const crnerf_cnn_weights* w, const float* x, int64_t n_pixels,
                       int64_t pix_stride, int64_t ch_stride, float* out, float* scratch,
                       void* stream);

/* The sharded (multi-GPU) form of the same block, SURVEY.md 8(e) scheme B:
 *   stats1: per-channel sums of this rank's pixels          -> sums (64)
 *   [all-reduce sums; mean = sums / total_pixels]
 *   stats2: un-normalised Gram of cnet.convs(content - mean) -> gram (32x32)
 *   [all-reduce gram]
 *   apply : build the fused 64->3 map from gram / total_pixels and the style
 *           branch, apply it to this rank's pixels            -> rgb (3, n_pixels)
 * `scratch` is the same buffer crnerf_style_forward takes. */
int crnerf_style_stats1(const float* content, int64_t n_pixels, int64_t pix_stride,
                        int64_t ch_stride, float* sums, float* scratch, void* stream);
int crnerf_style_stats2(const crnerf_style_weights* w, const float* content, int64_t n_pixels,
                        int64_t pix_stride, int64_t ch_stride, const float* mean, float* gram,
                        float* scratch, void* stream);
/* gram_normalised = (all-reduced gram) / total_pixels, computed by the caller */
int crnerf_style_apply(const crnerf_style_weights* w, const float* content, int64_t n_pixels,
                       int64_t pix_stride, int64_t ch_stride, const float* mean,
                       const float* gram_normalised, const float* style, int64_t n_style_pixels,
                       int64_t s_pix_stride, int64_t s_ch_stride, float* rgb, float* transmatrix,
                       float* scratch, void* stream);

/* ---- style/content encoder ---------------------------------------
 * encoder_sameoutputsize.forward (models/linearStyleTransfer.py:250-276), inference only:
 * conv1..conv7 in the module's order; weight[i] / bias[i] are the fp32 tensors of
 * conv{i+1} exactly as its state_dict holds them ((3,3,1,1), (64,3,3,3), (64,64,3,3),
 * (128,64,3,3), (128,128,3,3), (128,128,3,3), (64,128,1,1)). */
typedef struct {
  const float* weight[7];
  const float* bias[7];
} crnerf_encoder_weights;
size_t crnerf_encoder_packed_bytes(void);
/* one-time re-layout (fp16 hi/lo tensor-core images of conv3..conv6 + an fp32 blob); packed must be
 * 128-byte aligned */
int crnerf_encoder_pack(const crnerf_encoder_weights* w, void* packed, size_t packed_bytes, void* stream);
size_t crnerf_encoder_scratch_bytes(int height, int width);
/* img (3, height, width) fp32 (= the (1,3,H,W) tensor enc_a receives, eval.py:278) ->
 * out (64, 32, 32) fp32 (= (1,64,32,32)).  height, width in [8, 8192]; scratch 256-byte aligned. */
int crnerf_encoder_forward(const void* packed, const float* img, int height, int width, float* out,
                           void* scratch, size_t scratch_bytes, void* stream);

/* Training step (train_mask_grid_sample.py back-propagates through enc_a every step; the module under
 * autograd, models/linearStyleTransfer.py:250-276).  forward_train = the same kernels, every layer's
 * activation planes kept in `tape` (crnerf_encoder_tape_bytes, 256-byte aligned); backward = input-
 * gradient convolutions on the forward's tensor-core kernel (flipped weight images inside `packed`),
 * weight gradients as tcgen05 GEMMs with K = pixels, fixed-order reductions (deterministic), no library
 * convolution.  grad_out (64,32,32); grads.weight[i] / grads.bias[i] receive the gradients of
 * conv{i+1} in the state_dict layouts (overwritten, not accumulated); grad_img (3,H,W) or NULL.
 * `packed` must be the image of the weights the forward ran with. */
typedef struct {
  float* weight[7];
  float* bias[7];
} crnerf_encoder_grads;
size_t crnerf_encoder_tape_bytes(int height, int width);
size_t crnerf_encoder_backward_scratch_bytes(int height, int width);
int crnerf_encoder_forward_train(const void* packed, const float* img, int height, int width, float* out,
                                 void* tape, size_t tape_bytes, void* stream);
int crnerf_encoder_backward(const void* packed, const float* img, int height, int width, const float* out,
                            const float* grad_out, const void* tape, const crnerf_encoder_grads* grads,
                            float* grad_img, void* scratch, size_t scratch_bytes, void* stream);

/* ---- optimizer update of the training step ----------------------------------------------
 * The Adam optimizer as the reference builds it (utils/__init__.py:31-32 `Adam(parameters, lr=hparams.lr,
 * eps=eps, weight_decay=hparams.weight_decay)`, stepped once per batch by the training loop of
 * train_mask_grid_sample.py) over a table of tensors: one launch per 48 tensors.  params / grads /
 * exp_avg / exp_avg_sq / numel are HOST arrays of n_tensors entries holding device pointers (fp32,
 * contiguous) and element counts; `step` is a DEVICE float holding the number of steps already taken
 * (the update uses *step + 1; the caller increments it afterwards, so the call can be captured in a CUDA
 * graph); lr_dev, when non-NULL, is a DEVICE float that overrides `lr` (a scheduler can change it
 * between graph replays).  Element-wise fp32 math of the PyTorch optimizer with amsgrad = False
 * (m += (1-b1)(g-m); v = b2 v + (1-b2) g g; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)), bias
 * corrections in double. */
int crnerf_adam_step(int n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                     float* const* exp_avg_sq, const int64_t* numel, const float* step, const float* lr_dev,
                     double lr, double beta1, double beta2, double eps, double weight_decay, int maximize,
                     void* stream);

/* debug (tests only): dump the post-activation values of `layer` (0..10) for every
 * point of later fused launches into dbg_buf (n_points x 256 floats); NULL disables. */
int crnerf_debug_set(float* dbg_buf, int layer);
/* debug (tests only, host-only, no GPU needed): the weight-chunk program of the
 * fused kernel as flat int32: [n_chunks, n_units, image_bytes, 11 ints per chunk
 * (offset, bytes, layer, rows, row0, wcol0, wcols, a_src, a_k0, nk, kind), 7 ints per
 * unit (layer, half, n, chunk0, nchunks, first_of_layer, last_of_layer)].
 * Returns the number of ints written, or a negative crnerf_status. */
int crnerf_debug_program(int e_xyz, int e_dir, int32_t* out_host, int cap);

#ifdef __cplusplus
}
#endif
#endif /* CRNERF_B200_H_ */

/*
 * gsr_b200.h — C ABI of the B200-native differentiable 3D-Gaussian-splatting rasterizer.
 *
 * Drop-in boundary for the hot path of RPL-CS-UCL/gs_localization (LoGS): the depth+alpha
 * fork of diff-gaussian-rasterization.  Every entry point replaces one host entry of the
 * reference's native library; citations are relative to
 *   gaussian_splatting/submodules/diff-gaussian-rasterization/
 *
 * Conventions (same as the reference, rasterize_points.cu:96-115):
 *   - all pointers are DEVICE pointers to contiguous float32 / int32 data unless stated;
 *   - absent optional inputs are NULL (the reference passes the data pointer of an empty tensor);
 *   - matrices are 4x4 in the transposed storage the reference uses (scene/cameras.py:56-59);
 *   - the caller owns all memory.  Scratch is three opaque byte buffers ("geometry",
 *     "binning", "image") that the forward sizes through caller-supplied allocation
 *     callbacks and the backward re-reads, exactly like the reference's
 *     std::function<char*(size_t)> resize callbacks (cuda_rasterizer/rasterizer.h:24-27,
 *     rasterize_points.cu:27-33).  Their layout is private to this library;
 *   - `stream` is a cudaStream_t (NULL = legacy default stream, which is what the
 *     reference always uses).
 *
 * There is no CPU fallback: every function launches sm_100a kernels or fails.
 */
#ifndef GSR_B200_H_
#define GSR_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define GSR_API __attribute__((visibility("default")))
#else
#define GSR_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define GSR_ABI_VERSION 2

/* error codes (negative return values) */
#define GSR_OK 0
#define GSR_ERR_INVALID_ARGUMENT (-1) /* reference: AT_ERROR in rasterize_points.cu:57-59 */
#define GSR_ERR_CUDA (-2)             /* reference: CHECK_CUDA throw, auxiliary.h:166-173 */
#define GSR_ERR_ALLOC (-3)            /* an allocation callback returned NULL */
#define GSR_ERR_UNSUPPORTED (-4)      /* reference: "For non-RGB, provide precomputed Gaussian colors!" rasterizer_impl.cu:243-246 */

/* Scratch allocation callback: must return a device pointer to at least `bytes` bytes,
 * 128-byte aligned, that stays valid until the matching backward has run.
 * Mirrors std::function<char*(size_t N)> (rasterizer.h:24-27). */
typedef char* (*gsr_alloc_fn)(size_t bytes, void* user);

GSR_API int gsr_abi_version(void);
/* Human-readable description of the last error on the calling thread. */
GSR_API const char* gsr_last_error(void);
/* Number of kernels this library has launched since load (all threads). bench.py's gpu_launches. */
GSR_API unsigned long long gsr_launch_count(void);

/* Scratch sizes.  Replace CudaRasterizer::required<GeometryState|ImageState|BinningState>
 * (rasterizer_impl.h:66-72).  The forward calls the callbacks with exactly these values. */
GSR_API size_t gsr_geometry_bytes(int P);
GSR_API size_t gsr_image_bytes(int width, int height);
GSR_API size_t gsr_binning_bytes(long long num_rendered, int width, int height);

/*
 * Forward.  Replaces CudaRasterizer::Rasterizer::forward (rasterizer_impl.cu:197-339,
 * declared rasterizer.h:35-60) as called from RasterizeGaussiansCUDA (rasterize_points.cu:86-116).
 *
 *   P, D, M            number of Gaussians, active SH degree, SH coefficients stored per Gaussian
 *   background[3], means3D[P,3], shs[P,M,3] | colors_precomp[P,3], opacities[P],
 *   scales[P,3] + rotations[P,4] | cov3D_precomp[P,6], viewmatrix[16], projmatrix[16], cam_pos[3]
 *   out_color[3,H,W], out_depth[1,H,W], out_alpha[1,H,W], radii[P] (int32)
 *   n_touched[P] (int32, may be NULL): pose-variant extra output
 *
 * Returns num_rendered (>= 0) — the number of (Gaussian, tile) instances — or a negative
 * error code.  Like the reference it waits once for num_rendered to reach the host
 * (rasterizer_impl.cu:282).  With no history for this (P, W, H) it does so before sizing the binning
 * buffer, exactly like the reference; afterwards it launches the rest of the forward speculatively into
 * a buffer sized from recent calls (binning_alloc is then called with that larger size, and once more
 * with the exact size in the rare case the guess was too small), so the GPU does not idle during the wait.
 */
GSR_API long long gsr_rasterize_forward(
    gsr_alloc_fn geometry_alloc, gsr_alloc_fn binning_alloc, gsr_alloc_fn image_alloc, void* user,
    int P, int D, int M,
    const float* background, int width, int height,
    const float* means3D, const float* shs, const float* colors_precomp, const float* opacities,
    const float* scales, float scale_modifier, const float* rotations, const float* cov3D_precomp,
    const float* viewmatrix, const float* projmatrix, const float* cam_pos,
    float tan_fovx, float tan_fovy, int prefiltered,
    float* out_color, float* out_depth, float* out_alpha, int* radii, int* n_touched,
    int debug, void* stream);

/*
 * Sync-free forward for callers that own persistent scratch (the fused pose-refinement loop): same kernels
 * and results as gsr_rasterize_forward, but no allocation callbacks and no host wait, so the whole call can
 * be captured into a CUDA graph.  geometry_buffer / image_buffer must hold gsr_geometry_bytes(P) /
 * gsr_image_bytes(W,H); binning_buffer must hold gsr_binning_bytes(binning_capacity, W, H).  global_sort != 0
 * selects the global radix-sort layout (needed for speed only when tile lists exceed 4096 entries).
 * num_rendered stays on the device: gsr_read_counters returns {num_rendered, overflow, longest tile list};
 * overflow != 0 means num_rendered exceeded binning_capacity and the outputs of that forward are invalid.
 * gsr_rasterize_backward accepts binning_capacity as its num_rendered argument for such a forward.
 */
GSR_API int gsr_rasterize_forward_async(
    char* geometry_buffer, char* binning_buffer, long long binning_capacity, int global_sort, char* image_buffer,
    int P, int D, int M,
    const float* background, int width, int height,
    const float* means3D, const float* shs, const float* colors_precomp, const float* opacities,
    const float* scales, float scale_modifier, const float* rotations, const float* cov3D_precomp,
    const float* viewmatrix, const float* projmatrix, const float* cam_pos,
    float tan_fovx, float tan_fovy,
    float* out_color, float* out_depth, float* out_alpha, int* radii, int* n_touched,
    const float* cull_records /* gsr_build_cull_records output for these means3D / scales / rotations, or NULL */, void* stream);
GSR_API int gsr_read_counters(const char* geometry_buffer, int P, unsigned int* out3, void* stream);
/* Static map, packed once at load (LoGS localizes every query of a scene against one read-only map loaded by
 * scene/gaussian_model.py:215-256): records[P][4] = mean x, y, z and the static factor of the conservative screen-radius
 * bound.  Passed as `cull_records` to gsr_rasterize_forward_async, the cull pass streams 16 bytes per Gaussian instead
 * of 40 from three arrays; results are identical.  Rebuild after any change of means3D / scales / rotations. */
GSR_API int gsr_build_cull_records(int P, const float* means3D, const float* scales, const float* rotations, float* records, void* stream);
/* The overflow flag reported by gsr_read_counters is STICKY: it stays set until cleared, however many forwards ran on the
 * geometry buffer in between (a CUDA graph replays many forwards between two reads).  Clear it once after allocating the
 * buffer (its contents are otherwise undefined) and whenever a new query starts. */
GSR_API int gsr_clear_overflow(char* geometry_buffer, int P, void* stream);

/*
 * Backward.  Replaces CudaRasterizer::Rasterizer::backward (rasterizer_impl.cu:343-444,
 * declared rasterizer.h:62-88) as called from RasterizeGaussiansBackwardCUDA
 * (rasterize_points.cu:170-203).
 *
 * Zero-fill contract: output buffers that lie within 128 bytes of each other (one arena carved into 128-byte
 * aligned slices, or a framework allocator's padding) are zero-filled as ONE span, gap included — do not keep live
 * data in a gap smaller than 128 bytes between two gradient outputs.
 *
 * Gradient outputs are written densely for ALL P Gaussians (invisible ones get zeros), so
 * the caller does NOT need to zero-fill them first (the reference needs nine torch::zeros,
 * rasterize_points.cu:158-166).  Any output pointer may be NULL to skip that gradient.
 *   dL_dmean2D[P,3], dL_dconic[P,4] (x,y,_,w as the reference's float4), dL_dopacity[P],
 *   dL_dcolor[P,3], dL_dmean3D[P,3], dL_dcov3D[P,6], dL_dsh[P,M,3], dL_dscale[P,3], dL_drot[P,4]
 *
 * Upstream gradients: dL_dpix[3,H,W] is required; dL_ddepth[H,W] and dL_dalpha[H,W] may be NULL, meaning no
 * gradient flows into that output (the same result as a tensor of zeros, without the tensor).
 *
 * Pose extension (diff_gaussian_rasterization_pose surface used by
 * gs_localization/pipelines/tools/__init__.py:58-141): when dL_dtau is non-NULL, six floats
 * [d/drho(3), d/dtheta(3)] of the loss w.r.t. the left perturbation T_w2c <- exp(tau) T_w2c
 * (tools/pose_utils.py:90-122) are written there; projmatrix_raw (transposed storage) is then
 * required.
 */
GSR_API int gsr_rasterize_backward(
    int P, int D, int M, long long num_rendered,
    const float* background, int width, int height,
    const float* means3D, const float* shs, const float* colors_precomp, const float* out_alpha,
    const float* scales, float scale_modifier, const float* rotations, const float* cov3D_precomp,
    const float* viewmatrix, const float* projmatrix, const float* projmatrix_raw, const float* cam_pos,
    float tan_fovx, float tan_fovy, const int* radii,
    char* geometry_buffer, char* binning_buffer, char* image_buffer,
    const float* dL_dpix, const float* dL_ddepth, const float* dL_dalpha,
    float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
    float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot,
    float* dL_dtau,
    int debug, void* stream);

/* Replaces CudaRasterizer::Rasterizer::markVisible (rasterizer_impl.cu:141-153): present[i] = z_view > 0.2. */
GSR_API int gsr_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     unsigned char* present, void* stream);

/*
 * Introspection for parity tests (not on the hot path): copy the library's private
 * intermediates out of the scratch buffers in the REFERENCE's formats
 * (GeometryState / BinningState / ImageState, rasterizer_impl.h:29-64).  Any pointer may be NULL.
 *   depths[P] means2D[P,2] cov3D[P,6] conic_opacity[P,4] rgb[P,3] clamped[P,3](u8) tiles_touched[P]
 *   (pre-zeroed by the caller; only rows of visible Gaussians are written);
 *   keys[R] = sorted tile<<32|depth keys, list[R] = sorted Gaussian ids (point_list); ranges[T,2]; n_contrib[H*W]
 * The library never materialises the reference's unsorted key array or the per-Gaussian offsets.
 */
GSR_API int gsr_export_state(
    int P, long long num_rendered, int width, int height,
    const char* geometry_buffer, const char* binning_buffer, const char* image_buffer,
    float* depths, float* means2D, float* cov3D, float* conic_opacity, float* rgb, unsigned char* clamped,
    uint32_t* tiles_touched, uint64_t* keys, uint32_t* list,
    uint32_t* ranges, uint32_t* n_contrib, void* stream);

/*
 * Per-stage device timing for bench.py's stage split.  While enabled, every
 * forward/backward brackets its stages with CUDA events on the launching stream and
 * synchronises at the end to accumulate them — measurement only, never on in a timed run.
 * Stage order: preprocess(+scan), duplicate_with_keys, radix_sort, tile_ranges, render,
 * render_backward, preprocess_backward.  gsr_stage_times returns the number of stages.
 */
GSR_API void gsr_stage_timing(int enable);
GSR_API int gsr_stage_times(double* total_ms, unsigned long long* calls, int n);

/*
 * Device-side glue of LoGS' pose-refinement loop (gradient_decent,
 * gs_localization/pipelines/7scenes_localize_full_dslam.py:66-91), so that an iteration is
 * forward -> loss -> backward -> pose update with no host round trip and no framework ops.
 *
 * gsr_l1_loss_grad: tracking loss with all-ones masks (tools/descent_utils.py:85-123) and its
 *   gradient in one pass: *loss_accum += weight * mean|image - target| ; dL_dimage = weight * sign(.)/n.
 *   loss_accum must be zeroed by the caller.
 * gsr_pose_adam_step: torch.optim.Adam (betas 0.9/0.999, eps 1e-8, per-group lr) on the six pose deltas
 *   given dL_dtau = [d/drho, d/dtheta]; then T_w2c <- SE3_exp(tau) T_w2c (tools/pose_utils.py:54-122) and
 *   the per-view constants viewmatrix = T_w2c^T, projmatrix = viewmatrix @ projmatrix_raw, campos
 *   (tools/camera_utils.py:144-158).  adam_m[6], adam_v[6], step_count[1], w2c[16] (row-major) are state
 *   updated in place; tau_norm[1] (optional) receives |tau| for the caller's convergence test.
 */
GSR_API int gsr_l1_loss_grad(const float* image, const float* target, float* dL_dimage, long long n, float weight,
                             float* loss_accum, void* stream);
/* Photometric loss of LoGS map training and its gradient in two passes over the image
 * (gs_localization/gs/7scenes_gs_full_dslam.py:165-166; gaussian_splatting/utils/loss_utils.py:17-64):
 *   *loss_accum += (1-lambda) * mean|image - target| + lambda * (1 - mean SSIM(image, target))
 *   dL_dimage[C,H,W] = its gradient w.r.t. image (what the rasterizer's backward takes as dL_dpix).
 * 11x11 Gaussian window (sigma 1.5), zero padding, per channel.  scratch: 3*C*H*W + 2 floats. */
GSR_API int gsr_l1_ssim_loss_grad(const float* image, const float* target, int channels, int height, int width,
                                  float lambda_dssim, float* loss_accum, float* dL_dimage, float* scratch, void* stream);
/* Optimiser side of one map-training iteration, fused (gs_localization/gs/7scenes_gs_full_dslam.py:225-242;
 * gaussian_splatting/scene/gaussian_model.py:44-58 activations, :152-168 Adam groups, :405-407 statistics):
 *   do_stats: over the visible set (radii > 0): max_radii2D = max(., radii); xyz_gradient_accum += |dL_dmeans2D.xy|;
 *             denom += 1.
 *   do_adam : torch.optim.Adam, step numbers steps[6] (>= 1, one per group like torch's per-parameter counters), on the five tensors
 *             params[] = { xyz [P,3], features [P,M,3], opacity [P] (logit), scaling [P,3] (log), rotation [P,4] }
 *             with grads[] taken w.r.t. what the rasterizer consumed (activated opacity / scales / rotations; the
 *             chain rule through sigmoid / exp / normalize is applied here), state exp_avg[] / exp_avg_sq[], and
 *             lrs[6] = { xyz, f_dc, f_rest, opacity, scaling, rotation } (f_dc = SH coefficient 0 of `features`,
 *             f_rest the others).  A negative learning rate leaves that group untouched (the reference skips groups
 *             whose parameter was just replaced, e.g. after reset_opacity).  The activated copies
 *             opacity_act [P], scaling_act [P,3], rotation_act [P,4] are refreshed for the next render.
 * All tables hold device pointers; the tables themselves are host arrays. */
GSR_API int gsr_map_adam_step(int P, int M, int do_stats, int do_adam, const int* steps, const float* lrs, float beta1, float beta2,
                              float eps, float* const* params, const float* const* grads, float* const* exp_avg,
                              float* const* exp_avg_sq, float* opacity_act, float* scaling_act, float* rotation_act,
                              const float* dL_dmeans2D, const int* radii, float* max_radii2D, float* xyz_gradient_accum,
                              float* denom, void* stream);
/* Visible-rows gradient exchange of data-parallel map training (the dense all-reduce moves 59 floats
 * per Gaussian although one view touches a few percent of the map).  gsr_pack_gradient_rows gathers rows row_ids[n_rows]
 * of the five gradient tensors grads[] = { xyz [P,3], features [P,M,3], opacity [P], scaling [P,3], rotation [P,4] }
 * into table[padded_rows][1 + 11 + 3M]: column 0 is the row id (int bits; -1 for padding rows).  gsr_add_gradient_rows
 * adds a (received) table back into dense gradients; ids must be unique within a table.  Host tables of device pointers. */
GSR_API int gsr_pack_gradient_rows(const long long* row_ids, int n_rows, int padded_rows, int M, float* const* grads,
                                   float* table, void* stream);
GSR_API int gsr_add_gradient_rows(const float* table, int padded_rows, int M, int P, float* const* grads, void* stream);
/* The same exchange without any host round trip.  gsr_pack_visible_rows appends the rows with radii[i] > 0 through the
 * device counter `count` (zeroed by the call) into table[capacity_rows + 1][1 + 11 + 3M + 3]: columns = row id (int bits),
 * the gradients as above, then dL_dmeans2D x, y and the radius of the view (the inputs of the densification statistics,
 * gaussian_model.py:405-407).  Row `capacity_rows` is a header whose word 0 holds the number of visible rows of the view
 * (it may exceed capacity_rows: the surplus rows were dropped and the caller must treat the step as failed).
 * gsr_add_counted_rows applies a (received) table: adds the gradient rows when add_gradients != 0 (a rank passes 0 for its
 * own table), and, when max_radii2D / xyz_gradient_accum / denom are given, updates the statistics for the table's rows.
 * Row ids outside [0, P) are ignored. */
GSR_API int gsr_pack_visible_rows(const int* radii, int P, int M, float* const* grads, const float* dL_dmeans2D, float* table,
                                  int capacity_rows, unsigned int* count, void* stream);
GSR_API int gsr_add_counted_rows(const float* table, int capacity_rows, int M, int P, float* const* grads, int add_gradients,
                                 float* max_radii2D, float* xyz_gradient_accum, float* denom, void* stream);
/* distCUDA2 of simple-knn (gaussian_splatting/submodules/simple-knn/spatial.cu:15-26, simple_knn.cu:185-221):
 * mean_dists[i] = mean of the squared distances from points[i] to its three nearest other points (exact search;
 * FLT_MAX stands in for missing neighbours when n_points < 4, as in the reference).  points: [n,3] float32.
 * workspace: gsr_knn_workspace_bytes(n) bytes of device memory. */
GSR_API size_t gsr_knn_workspace_bytes(long long n_points);
GSR_API int gsr_dist2_knn3(const float* points, long long n_points, float* mean_dists, char* workspace, void* stream);
/* Depth terms of the map-training loss (gs_localization/gs/7scenes_gs_full_dslam.py:168-184) and their gradient:
 *   *loss_accum += pearson_weight * min(1 - r(-m, d), 1 - r(inv_numerator / (m + 200), d))
 *                + l1_weight * mean|d*mask - gt*mask|,  mask = gt_depth > 0
 * d = depth[n] (rendered), m = pseudo_depth[n] (monocular estimate; NULL drops the term), gt_depth[n] (NULL drops
 * the L1 term); r = Pearson correlation (torchmetrics.functional.pearson_corrcoef in the reference; restated from
 * its definition, centred sums in double).  Reference weights: 0.01, inv_numerator 1000, 0.05.
 * dL_ddepth[n] is written.  scratch: 9 doubles. */
GSR_API int gsr_depth_loss_grad(const float* depth, const float* pseudo_depth, const float* gt_depth, long long n,
                                float inv_numerator, float pearson_weight, float l1_weight, float* loss_accum,
                                float* dL_ddepth, double* scratch, void* stream);
/* Full tracking loss of LoGS and its gradient in one pass (tools/descent_utils.py:85-123):
 *   image_ab = exp(exposure[0]) * image + exposure[1]
 *   L = mean_{3HW} om * |image_ab*gm - gt_image*gm| + depth_weight * mean_{HW} |depth*dm - gt_depth*dm|
 *   om = opacity > opacity_threshold, gm = grad_mask (float 0/1, NULL = ones), dm = (gt_depth > 0.01) * om * gm.
 * depth_weight = 1 - config.Training.alpha; gt_depth NULL = monocular.  exposure NULL = (0, 0).
 * *loss_accum += L; dL_dimage[3,H,W] and dL_ddepth[H,W] (if non-NULL) are written; dL_dexposure[2] (optional) is
 * accumulated into.  gsr_exposure_adam_step is torch.optim.Adam on the two exposure scalars
 * (7scenes_localize_full_dslam.py:48-61) and clears dL_dexposure. */
GSR_API int gsr_tracking_loss_grad(const float* image, const float* depth, const float* opacity, const float* gt_image,
                                   const float* gt_depth, const float* grad_mask, const float* exposure, int height, int width,
                                   float opacity_threshold, float depth_weight, float* loss_accum, float* dL_dimage,
                                   float* dL_ddepth, float* dL_dexposure, void* stream);
GSR_API int gsr_exposure_adam_step(float* exposure, float* dL_dexposure, float* adam_m, float* adam_v, float* step_count,
                                   float lr, void* stream);
GSR_API int gsr_pose_adam_step(const float* dL_dtau, float* adam_m, float* adam_v, float* step_count, float lr_trans,
                               float lr_rot, float* w2c, const float* projmatrix_raw, float* viewmatrix, float* projmatrix,
                               float* campos, float* tau_norm, void* stream);

/* Stand-alone stable LSD radix sort of (u64 key, u32 value) pairs on bits [0, end_bit) —
 * the hand-written onesweep that replaces cub::DeviceRadixSort::SortPairs
 * (rasterizer_impl.cu:304-309).  temp must hold gsr_sort_temp_bytes(n) bytes. */
GSR_API size_t gsr_sort_temp_bytes(long long n);
GSR_API int gsr_sort_pairs(const uint64_t* keys_in, uint64_t* keys_out, const uint32_t* vals_in, uint32_t* vals_out,
                   uint64_t* keys_tmp, uint32_t* vals_tmp, long long n, int end_bit, char* temp, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GSR_B200_H_ */

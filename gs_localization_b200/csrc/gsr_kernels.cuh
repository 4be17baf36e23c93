// Kernel-launcher interface between api.cu and the stage files.
#pragma once
#include <atomic>
#include <cstdio>
#include "gsr_common.cuh"

namespace gsr {

extern std::atomic<unsigned long long> g_launch_count;
inline void count_launch(int n = 1) { g_launch_count.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

// ---- programmatic dependent launch (griddepcontrol): a kernel launched through launch_pdl may be scheduled while its
// predecessor in the stream is still draining; it must call pdl_wait() before it touches anything the predecessor
// wrote (or anything the predecessor still reads), after which the predecessor has completed and its writes are visible.
// pdl_trigger() at the top lets the NEXT kernel's CTAs be scheduled as soon as this grid's last CTAs are resident.
// -DGSR_AB_NO_PDL builds the plain <<<>>> launches for measurements.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() {
#ifndef GSR_AB_NO_PDL
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_trigger() {
#ifndef GSR_AB_NO_PDL
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
#ifdef GSR_AB_NO_PDL
  kernel<<<grid, block, smem, stream>>>(args...);
#else
  // Inside a CUDA-graph capture the programmatic edges measured SLOWER than plain ones (headline pose iteration 0.451 vs
  // 0.384 ms, profiles/r2_pdl_ab.txt): graph launches already have no per-kernel launch latency to hide, and early-scheduled
  // dependents only take SM resources from the running kernel.  Eager launches gain 5.8 % per fwd+bwd.
  cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(stream, &capturing);
  if (capturing != cudaStreamCaptureStatusNone) {
    kernel<<<grid, block, smem, stream>>>(args...);
    return;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, args...);
#endif
}
#endif

struct PreprocessParams {
  int P, D, M, W, H;
  uint32_t grid_x, grid_y;
  const float* means3D;
  const float* scales;
  const float* rotations;
  const float* opacities;
  const float* shs;
  const float* cov3D_precomp;
  const float* colors_precomp;
  const float4* cull_rec;       // optional: packed static map (mean, static radius-bound factor) for the cull pass
  const float* viewmatrix;
  const float* projmatrix;
  const float* campos;
  float scale_modifier, tan_fovx, tan_fovy, focal_x, focal_y;
  int prefiltered;
  int sh_vec4;          // SH rows are 16-byte aligned and a multiple of 4 floats long
  int* radii;
  int* n_touched;       // may be null
  int* tile_diff;       // [(grid_y+1)(grid_x+1)], zeroed by preprocess_cull_kernel
  GeometryView geom;
};

void launch_preprocess_fwd(const PreprocessParams& p, cudaStream_t stream);
// colours of the visible slots (needs preprocess; only the blend needs it)
void launch_color_fwd(const PreprocessParams& p, cudaStream_t stream);
void launch_build_cull_records(int P, const float* means3D, const float* scales, const float* rotations, float4* rec, cudaStream_t stream);
void launch_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present, cudaStream_t stream);

// ---- binning (binning.cu)
constexpr int SORT_RADIX_BITS = 8;
constexpr int SORT_RADIX = 1 << SORT_RADIX_BITS;
constexpr int SORT_MAX_PASSES = 8;
constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;

struct SortTemp {
  uint32_t* hist;     // [SORT_MAX_PASSES][256]  digit counts -> exclusive bases
  uint32_t* status;   // [passes][ntiles][256]   decoupled look-back words
};
inline int sort_passes(int end_bit) { return (end_bit + SORT_RADIX_BITS - 1) / SORT_RADIX_BITS; }
inline size_t sort_num_tiles(long long n) { return (size_t)((n + SORT_TILE - 1) / SORT_TILE); }
size_t sort_temp_bytes(long long n, int passes);
void carve_sort_temp(char* base, long long n, int passes, SortTemp& t);
// zero hist/status (one memset)
void sort_temp_reset(char* base, long long n, int passes, cudaStream_t stream);

// coverage grid -> per-tile list lengths -> ranges / bucket cursors / num_rendered (counters[1]) / longest list (counters[4])
// + tile_order: tiles by descending list length (launch order of the blend CTAs)
void launch_scan_tiles(int* tile_diff, uint32_t gx, uint32_t gy, uint2* ranges, uint32_t* cursor, uint32_t* tile_order,
                       uint32_t* counters, uint32_t capacity, cudaStream_t stream);
// (Gaussian, tile) instances -> tile buckets as depth_bits << 32 | slot; also zeroes the slots' backward accumulators
void launch_scatter(int P, const GeometryView& g, uint32_t* cursor, uint64_t* comp, uint32_t grid_x, BinHeader hv,
                    BinHeader* header, cudaStream_t stream);
// long-list path (binning.cu "long lists"): Gaussian-level depth sort, then a tile-major scan straight into point_list
void launch_long_bin(int P, int T, uint32_t grid_x, uint32_t grid_y, const uint2* ranges, const GeometryView& g, const BinningView& bl,
                     long long capacity, BinHeader hv, BinHeader* header, cudaStream_t stream);
void launch_reset_cursors(int T, const uint2* ranges, uint32_t* cursor, cudaStream_t stream);
// per-tile sort of the buckets on the composite key; writes the sorted slots to point_list
void launch_tile_sort(int num_tiles, const uint2* ranges, uint64_t* comp, uint32_t* point_list, uint32_t capacity, const uint32_t* tile_order,
                      cudaStream_t stream);
// library radix sort (gsr_sort_pairs): the element count is a host upper bound and, optionally, a device pointer
void launch_sort_histogram(const uint64_t* keys, const uint32_t* n_ptr, long long n_host, int lo_bit, int end_bit, uint32_t* hist,
                           cudaStream_t stream);
// exclusive scan of the histograms + all onesweep passes; returns index (0/1) of the buffer holding the result
int launch_onesweep(uint64_t* keys[2], uint32_t* vals[2], const uint32_t* n_ptr, long long n_host, int end_bit,
                    const SortTemp& t, cudaStream_t stream);

// ---- blending (render.cu)
struct RenderParams {
  int W, H;
  uint32_t grid_x, grid_y;
  const uint2* ranges;
  const uint32_t* tile_order;   // CTA index -> tile (longest lists first)
  const uint32_t* point_list;   // sorted slots
  const float4* mean_tau;      // (x, y, 2 tau, -) per slot
  const float4* conic_opacity;
  const float4* rgbd;        // (r, g, b, depth) per slot
  const uint32_t* gid;       // slot -> Gaussian id (n_touched only)
  const float* bg;
  float* out_color;
  float* out_depth;
  float* out_alpha;
  uint32_t* n_contrib;
  int* n_touched;            // may be null
  uint32_t capacity;         // entries of point_list that exist (speculative launch: ranges may exceed it)
  float4* final_cd;          // [N] final (C0, C1, C2, D) for the segment-parallel backward
  uint4* units;              // backward work units appended here, *unit_count of them
  float* ckpt;               // pixel state at segment boundaries
  float4* rec;               // staged records of every blended segment, for the backward's bulk copies
  uint32_t units_cap;        // entries per unit array
  uint32_t* unit_count;      // counters + 5: [0] units, [3..6] units per cost class
};
void launch_render_fwd(const RenderParams& p, cudaStream_t stream);

// Dense zero rows the API owes for culled Gaussians, written by the blend backward itself between its work units
// (a memset on a side stream cannot overlap: the persistent blend CTAs leave it no SM slots).
struct FillSpans {
  float4* base[9];
  unsigned long long n4[9];   // 16-byte words per span
  int count;
};

struct RenderBwdParams {
  FillSpans fills;
  int W, H;
  uint32_t grid_x, grid_y;
  const char* binning_base;     // BinHeader at offset 0: where the units, checkpoints and records of this forward live
  const uint32_t* unit_count;
  uint32_t* queue;              // [2] ticket counter + finished-CTA counter of the unit queue, zero between launches
  const float4* final_cd;
  uint32_t max_units;           // launch bound: 4 (floor(R / SEG) + T + 2)
  const float* bg;
  const float* out_alpha;
  const uint32_t* n_contrib;
  const float* dL_dpix;
  const float* dL_ddepth;
  const float* dL_dalpha;
  float* grad_acc;           // [slots][12]: Q*sum u*(dx, dy), sum u*(dx^2, dx dy, dy^2, 1) (-> mean2D, conic, opacity), rgb, depth, pad, pad
};
void launch_render_bwd(const RenderBwdParams& p, cudaStream_t stream);

// SM count of the current device (cached per device)
int sm_count();

// ---- preprocess backward (backward_pre.cu)
struct PreBwdParams {
  int P, D, M, W, H;
  const float* means3D;
  const int* radii;
  const float* shs;
  const float* scales;
  const float* rotations;
  const float* cov3D_precomp;
  const float* viewmatrix;
  const float* projmatrix;
  const float* projmatrix_raw;  // pose mode only
  const float* campos;
  float scale_modifier, tan_fovx, tan_fovy, focal_x, focal_y;
  GeometryView geom;
  float* dL_dmean2D;   // [P,3]
  float* dL_dconic;    // [P,4]
  float* dL_dopacity;  // [P]
  float* dL_dcolor;    // [P,3]
  float* dL_dmean3D;   // [P,3]
  float* dL_dcov3D;    // [P,6]
  float* dL_dsh;       // [P,M,3]
  float* dL_dscale;    // [P,3]
  float* dL_drot;      // [P,4]
  float* dL_dtau;      // [6] or null
};
// one CTA per preprocess slot segment, threads beyond the segment's visible count idle
void launch_preprocess_bwd(const PreBwdParams& p, cudaStream_t stream);

// ---- pose-refinement glue (pose_step.cu)
void launch_l1_loss_grad(const float* image, const float* target, float* dL_dimage, size_t n, float weight, float* loss_out,
                         cudaStream_t stream);
void launch_tracking_loss_grad(const float* image, const float* depth, const float* opacity, const float* gt_image, const float* gt_depth,
                               const float* grad_mask, const float* exposure, int npix, float opacity_threshold, float depth_weight,
                               float* dL_dimage, float* dL_ddepth, float* loss_out, float* dL_dexposure, cudaStream_t stream);
void launch_exposure_adam_step(float* exposure, float* dL_dexposure, float* adam_m, float* adam_v, float* step_count, float lr,
                               cudaStream_t stream);
void launch_pose_adam_step(const float* dL_dtau, float* adam_m, float* adam_v, float* step_count, float lr_trans, float lr_rot,
                           float* w2c, const float* raw, float* viewmatrix, float* projmatrix, float* campos, float* tau_norm,
                           cudaStream_t stream);

// ---- photometric loss of map training (losses.cu); scratch holds 3*C*H*W + 2 floats
void launch_l1_ssim_loss_grad(const float* img1, const float* img2, int C, int H, int W, float lambda, float* loss, float* dL_dimg1,
                              float* scratch, cudaStream_t stream);

// ---- optimiser side of a map-training iteration (map_step.cu); a negative learning rate skips that group
struct MapStepParams {
  int P, M;
  int do_stats, do_adam;
  // raw parameters, updated in place; features is [P,M,3] (f_dc = coefficient 0, f_rest = the others)
  float *xyz, *features, *opacity, *scaling, *rotation;
  // gradients w.r.t. what the rasterizer consumed (activated opacity / scales / rotations)
  const float *g_xyz, *g_features, *g_opacity, *g_scaling, *g_rotation;
  float *m_xyz, *m_features, *m_opacity, *m_scaling, *m_rotation;
  float *v_xyz, *v_features, *v_opacity, *v_scaling, *v_rotation;
  float lr_xyz, lr_f_dc, lr_f_rest, lr_opacity, lr_scaling, lr_rotation;
  // refreshed activations for the next render
  float *opacity_act, *scaling_act, *rotation_act;
  // densification statistics
  const float* g_means2D;
  const int* radii;
  float *max_radii2D, *xyz_gradient_accum, *denom;
};
struct GradRowTensors { float* g[5]; };   // xyz, features, opacity, scaling, rotation gradients
void launch_pack_gradient_rows(const long long* idx, int k, int K, int M, const GradRowTensors& t, float* table, cudaStream_t stream);
void launch_add_gradient_rows(const float* table, int K, int M, const GradRowTensors& t, int P, cudaStream_t stream);
void launch_pack_visible_rows(const int* radii, int P, int M, const GradRowTensors& t, const float* g_means2D, float* table, int capacity,
                              unsigned int* count, cudaStream_t stream);
void launch_add_counted_rows(const float* table, int capacity, int M, int P, const GradRowTensors& t, int add_grads, float* max_radii2D,
                             float* xyz_gradient_accum, float* denom, cudaStream_t stream);
void launch_map_step(const MapStepParams& s, float beta1, float beta2, float eps, const int* steps, cudaStream_t stream);

// ---- simple-knn (knn.cu)
size_t knn_workspace_bytes(long long P);
int launch_dist2_knn3(const float* points, long long P, float* mean_dists, char* workspace, cudaStream_t stream);

// depth terms of the map-training loss; scratch holds 9 doubles
void launch_depth_loss_grad(const float* depth, const float* pseudo, const float* gt, int n, float k, float w_pearson, float w_l1,
                            float* loss, float* dL_ddepth, double* scratch, cudaStream_t stream);

}  // namespace gsr

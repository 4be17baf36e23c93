// Shared definitions of the B200 rasterizer kernels: scratch layout, rounding-pinned math
// helpers and launch bookkeeping.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace gsr {

constexpr int TILE_X = 16, TILE_Y = 16;     // reference config.h:15-17 (BLOCK_X/BLOCK_Y)
constexpr int TILE_PIX = TILE_X * TILE_Y;
constexpr int PRE_THREADS = 256;            // Gaussians per preprocess CTA

// SH constants, reference auxiliary.h:22-39
__device__ constexpr float SH_C0 = 0.28209479177387814f;
__device__ constexpr float SH_C1 = 0.4886025119029199f;
__device__ constexpr float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                       -1.0925484305920792f, 0.5462742152960396f};
__device__ constexpr float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                       0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                       -0.5900435899266435f};

// ---------------------------------------------------------------------------------------
// Scratch layout (private contract between this library's forward and backward).
// All slabs 128-byte aligned.  SoA, sized for P Gaussians / R instances / N pixels.
// ---------------------------------------------------------------------------------------
// Visible Gaussians are compacted per preprocess CTA: CTA b owns the 256 "slots"
// [256 b, 256 b + 256) and packs its visible Gaussians into the first block_vis[b] of them, in
// Gaussian order.  Slot ids are what the sort carries and what the blend kernels gather by;
// slot order == Gaussian order, so ties keep the reference's order.  No CTA depends on another.
struct GeometryView {           // replaces GeometryState (reference rasterizer_impl.h:29-44)
  float* depths;                // [slots] view-space z
  float2* means2D;              // [slots] pixel-space centre
  float4* conic_opacity;        // [slots] (conic.x, conic.y, conic.z, opacity)
  float4* rgbd;                 // [slots] (r, g, b, depth): one 16-byte gather for the blend kernels
  float* cov3D;                 // [6 slots]
  uint2* rect;                  // [slots] tile rectangle: x = minx | maxx<<16, y = miny | maxy<<16
  uint8_t* clamped;             // [slots] bit c set <=> channel c was clamped at 0
  uint32_t* gid;                // [slots] Gaussian id held by the slot
  uint32_t* block_vis;          // [ceil(P/256)] visible Gaussians of each preprocess CTA
  uint32_t* block_tiles;        // [ceil(P/256)+1] instances of each CTA
  uint32_t* block_cand;         // [ceil(P/256)+1] candidates of each segment after the cheap cull (preprocess pass 1)
  uint8_t* cand;                // [S] their positions inside the segment, ascending
  uint32_t* block_off;          // [ceil(P/256)+1] their exclusive prefix sum (ordered-emission path only)
  uint32_t* counters;           // [32] 1: num_rendered, 4: largest tile list
  float* grad_acc;              // [12 slots] backward accumulators, zero between uses
};

// Bucket cursors are hammered by one atomic per instance; packed, the T cursors of a frame share ~40 cache lines and
// the L2 atomic units serialise on them.  One cursor per 32-byte sector spreads them over all slices.
constexpr int CURSOR_STRIDE = 8;
constexpr int DIFF_STRIDE = 8;     // same for the cells of the coverage difference grid (4 atomics per visible Gaussian)

struct ImageView {              // replaces ImageState (reference rasterizer_impl.h:46-52)
  uint2* ranges;                // [T]   [start, end) of each tile in the sorted list, (0,0) if empty
  uint32_t* n_contrib;          // [N]
  int* tile_diff;               // [(gy+1)(gx+1) * DIFF_STRIDE] 2-D difference grid of tile coverage -> per-tile counts
  uint32_t* tile_cursor;        // [T * CURSOR_STRIDE] bucket write cursors, one per 32-byte sector
  uint32_t* tile_order;         // [T]   tiles by descending list length: launch order of the blend CTAs
  float4* final_cd;             // [N]   forward's final (C0, C1, C2, D) without the background term
};

struct BinningView {            // replaces BinningState (reference rasterizer_impl.h:54-64)
  uint32_t* point_list;         // [cap] sorted slots (same offset in both layouts: the backward needs nothing else)
  uint64_t* comp;               // tile-local path: [cap] depth bits << 32 | slot, bucketed by tile
  // global radix-sort path (tile lists too long for shared memory): ping-pong key/value arrays + sort temp
  uint64_t* keys[2];
  uint32_t* vals_other;
  char* sort_temp;
  // backward work units and the forward's per-segment pixel checkpoints (both layouts; offsets are from the buffer base
  // and are recorded in the header by the scatter kernel, because the backward is not told the capacity)
  uint2* units;                 // [max_units] (tile, segment)
  float* ckpt;                  // [max_units][5][256]  T, C0, C1, C2, D of the tile's pixels at the start of a segment
  size_t units_off, ckpt_off;
};

// The blend backward runs one CTA per (tile, segment of SEG list entries); the forward leaves the pixel state at every
// segment boundary so that a segment can be walked back to front without the ones behind it.
constexpr int SEG = 256;
constexpr int CKPT_FLOATS = 5 * 256;
struct BinHeader { unsigned long long capacity, units_off, ckpt_off; };
inline size_t max_units(long long cap, size_t tiles) { return (size_t)(cap / SEG) + tiles + 2; }
// slot of tile t's checkpoint for segment r: sum_{i<t} ceil(len_i / SEG) <= floor(start_t / SEG) + t, and consecutive
// tiles never overlap, so no prefix sum over segment counts is needed
__host__ __device__ inline uint32_t ckpt_slot(uint32_t range_start, uint32_t tile, uint32_t seg) { return range_start / SEG + tile + seg; }

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

template <typename T> inline void carve(char*& p, T*& out, size_t count) {
  p = (char*)align_up((size_t)p, 128);
  out = (T*)p;
  p += count * sizeof(T);
}

inline int num_pre_blocks(int P) { return (P + PRE_THREADS - 1) / PRE_THREADS; }

inline char* carve_geometry(char* base, int P, GeometryView& g) {
  char* p = base;
  const size_t S = (size_t)num_pre_blocks(P) * PRE_THREADS;
  carve(p, g.depths, S);
  carve(p, g.means2D, S);
  carve(p, g.conic_opacity, S);
  carve(p, g.rgbd, S);
  carve(p, g.cov3D, 6 * S);
  carve(p, g.rect, S);
  carve(p, g.clamped, S);
  carve(p, g.gid, S);
  carve(p, g.block_vis, (size_t)num_pre_blocks(P) + 1);
  carve(p, g.block_tiles, (size_t)num_pre_blocks(P) + 1);
  carve(p, g.block_cand, (size_t)num_pre_blocks(P) + 1);
  carve(p, g.cand, S);
  carve(p, g.block_off, (size_t)num_pre_blocks(P) + 1);
  carve(p, g.counters, (size_t)32);
  carve(p, g.grad_acc, 12 * S);
  return p;
}

inline char* carve_image(char* base, int W, int H, ImageView& im) {
  char* p = base;
  const size_t gx = (W + TILE_X - 1) / TILE_X, gy = (H + TILE_Y - 1) / TILE_Y;
  carve(p, im.ranges, gx * gy);
  carve(p, im.n_contrib, (size_t)W * H);
  carve(p, im.tile_diff, (gx + 1) * (gy + 1) * DIFF_STRIDE);
  carve(p, im.tile_cursor, gx * gy * CURSOR_STRIDE);
  carve(p, im.tile_order, gx * gy);
  carve(p, im.final_cd, (size_t)W * H);
  return p;
}

// reference rasterizer_impl.cu:35-50
inline uint32_t get_higher_msb(uint32_t n) {
  uint32_t msb = sizeof(n) * 4, step = msb;
  while (step > 1) {
    step /= 2;
    if (n >> msb) msb += step; else msb -= step;
  }
  if (n >> msb) msb++;
  return msb;
}

// ---------------------------------------------------------------------------------------
// Rounding-pinned arithmetic.  The binning decisions (radius, tile rectangle, depth key)
// and n_contrib must be bit-identical to the reference build, so the expressions that feed
// them use explicit IEEE intrinsics in exactly the order nvcc contracted the reference's
// source into FMAs on sm_100a (oracle/_ref/forward.sass).  These do
// not depend on -fmad / -use_fast_math.
// ---------------------------------------------------------------------------------------
// a0*b0 + a1*b1 + a2*b2 as the reference compiles it: second product rounded, others fused.
__device__ __forceinline__ float dot3c(float a0, float b0, float a1, float b1, float a2, float b2) {
  float t = __fmul_rn(a1, b1);
  t = __fmaf_rn(a0, b0, t);
  t = __fmaf_rn(a2, b2, t);
  return t;
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

}  // namespace gsr

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
  float4* mean_tau;             // [slots] (pixel-space centre x, y, 2 ln(255 o) inflated: the opacity-aware culling bound, slot bits)
  float4* conic_opacity;        // [slots] (conic.x, conic.y, conic.z, opacity)
  float4* rgbd;                 // [slots] (r, g, b, depth): one 16-byte gather for the blend kernels
  float4* msr;                  // [3 slots] (mean xyz, scale x) (scale y z, rot w x) (rot y z, -, -): what the preprocess backward needs
                                //           of the map, kept per visible slot so that it does not gather it by Gaussian id again
  float4* shd;                  // [3 slots] d(rgb before clamping)/d(unit view direction): (dr/dx dg/dx db/dx dr/dy) (dg/dy db/dy dr/dz dg/dz) (db/dz - - -),
                                //           left by the forward's colour kernel so that the backward never reads the SH rows again
  float* cov3D;                 // [6 slots]
  uint2* rect;                  // [slots] tile rectangle: x = minx | maxx<<16, y = miny | maxy<<16
  uint8_t* clamped;             // [slots] bit c set <=> channel c was clamped at 0
  uint32_t* gid;                // [slots] Gaussian id held by the slot
  uint32_t* block_vis;          // [ceil(P/256)] visible Gaussians of each preprocess CTA
  uint32_t* block_tiles;        // [ceil(P/256)+1] instances of each CTA
  uint32_t* block_cand;         // [ceil(P/256)+1] candidates of each segment after the cheap cull (preprocess pass 1)
  uint8_t* cand;                // [S] their positions inside the segment, ascending
  uint32_t* block_off;          // [ceil(P/256)+1] exclusive prefix sum of block_vis (long-list path)
  uint32_t* counters;           // [32] 1: num_rendered, 3: overflow, 4: largest tile list, 5: backward units, 6-7: backward queue,
                                //      8-11: units per cost class, 13: visible Gaussians (long-list path),
                                //      16: sticky overflow
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
  // long-list path (tile lists too long for shared memory): scratch of the Gaussian-level depth sort
  // (gcap = min(slots, capacity) entries); the instances themselves go straight to point_list
  uint64_t* keys[2];       // unused (kept for layout helpers)
  char* sort_temp;
  uint64_t* gkeys[2];
  uint32_t* gvals[2];
  char* gsort_temp;
  long long gcap;
  // backward work units and the forward's per-segment pixel checkpoints (both layouts; offsets are from the buffer base
  // and are recorded in the header by the scatter kernel, because the backward is not told the capacity)
  uint4* units;                 // [2][units_cap] (tile | quadrant << 30, segment, range start, range end): one per 8x8 pixel quadrant and
                                //   piece, in four cost classes (render.cu); units_cap = 16 max_units
  float* ckpt;                  // [4 max_units][5][256]  T, C0, C1, C2, D of the tile's pixels in front of a backward piece
  float4* rec;                  // [4 max_units][BSEG][3] the forward's staged 48-byte splat records of every batch it blended, contiguous:
                                //                     the backward fetches a piece with one bulk copy (cp.async.bulk, 3 KB)
  size_t units_off, ckpt_off, rec_off, units_cap;
};

// The forward blends a tile's list in batches of SEG entries; the backward walks it in independent pieces of BSEG
// entries (one warp per (tile, 8x8 pixel quadrant, piece)): the forward leaves the pixel state at every BSEG boundary so
// that a piece can be walked back to front without the ones behind it.
constexpr int SEG = 256;
constexpr int BSEG = 64;
constexpr int BSEG_PER_SEG = SEG / BSEG;
constexpr int CKPT_FLOATS = 5 * 256;
constexpr int REC_BYTES = 48;               // staged splat record: (x, y, 2 tau, slot bits) (conic x, y, z, opacity) (r, g, b, depth)
constexpr int REC_FLOAT4 = SEG * REC_BYTES / 16;   // float4 words per batch of records
constexpr int BREC_FLOAT4 = BSEG * REC_BYTES / 16; // float4 words per backward piece
struct BinHeader { unsigned long long capacity, units_off, ckpt_off, rec_off, units_cap; };
inline size_t max_units(long long cap, size_t tiles) { return (size_t)(cap / SEG) + tiles + 2; }
// slot of tile t's batch r: sum_{i<t} ceil(len_i / SEG) <= floor(start_t / SEG) + t, and consecutive
// tiles never overlap, so no prefix sum over segment counts is needed
__host__ __device__ inline uint32_t ckpt_slot(uint32_t range_start, uint32_t tile, uint32_t seg) { return range_start / SEG + tile + seg; }
// slot of tile t's backward piece s (list positions [s BSEG, (s+1) BSEG)): 4 (floor(start_t / SEG) + t) + s; the slots of a tile
// end at most 4 floor(len_t / SEG) + 4 after its base, which is where the next tile's begin at the earliest
__host__ __device__ inline uint32_t piece_slot(uint32_t range_start, uint32_t tile, uint32_t piece) {
  return BSEG_PER_SEG * (range_start / SEG + tile) + piece;
}

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
  carve(p, g.mean_tau, S);
  carve(p, g.conic_opacity, S);
  carve(p, g.rgbd, S);
  carve(p, g.msr, 3 * S);
  carve(p, g.shd, 3 * S);
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

// A splat can only reach alpha = min(0.99, o * exp(power)) >= 1/255 where power >= -tau, tau = ln(255 o),
// i.e. inside the ellipse Q(d) = a dx^2 + 2 b dx dy + c dy^2 <= 2 tau around its centre.  two_tau is
// inflated (0.1 % + 1e-3) so that rounding in the exact per-pixel test can never accept a pixel this bound
// rejects; < 0 means "never visible" (o < 1/255), +inf disables culling (degenerate conic).
__device__ __forceinline__ float splat_two_tau(float a, float b, float c, float opacity) {
  const float o255 = opacity * 255.0f;
  if (!(o255 >= 1.0f)) return -1.0f;
  if (!(a * c - b * b > 0.f) || !(a > 0.f) || !(c > 0.f)) return __int_as_float(0x7f800000);
  return 2.0f * __logf(o255) * 1.001f + 1e-3f;
}

// Real SH basis of group g (coefficients 4g .. 4g+3) at the unit direction (x, y, z): values w and their derivatives
// wx, wy, wz with respect to the direction components (reference forward.cu:20-71, backward.cu:20-139).
__device__ __forceinline__ void sh_basis_group(int g, float x, float y, float z, float (&w)[4], float (&wx)[4], float (&wy)[4], float (&wz)[4]) {
  const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
  if (g == 0) {
    w[0] = SH_C0;       wx[0] = 0.f;    wy[0] = 0.f;    wz[0] = 0.f;
    w[1] = -SH_C1 * y;  wx[1] = 0.f;    wy[1] = -SH_C1; wz[1] = 0.f;
    w[2] = SH_C1 * z;   wx[2] = 0.f;    wy[2] = 0.f;    wz[2] = SH_C1;
    w[3] = -SH_C1 * x;  wx[3] = -SH_C1; wy[3] = 0.f;    wz[3] = 0.f;
  } else if (g == 1) {
    w[0] = SH_C2[0] * xy;                     wx[0] = SH_C2[0] * y;        wy[0] = SH_C2[0] * x;        wz[0] = 0.f;
    w[1] = SH_C2[1] * yz;                     wx[1] = 0.f;                 wy[1] = SH_C2[1] * z;        wz[1] = SH_C2[1] * y;
    w[2] = SH_C2[2] * (2.f * zz - xx - yy);   wx[2] = SH_C2[2] * -2.f * x; wy[2] = SH_C2[2] * -2.f * y; wz[2] = SH_C2[2] * 4.f * z;
    w[3] = SH_C2[3] * xz;                     wx[3] = SH_C2[3] * z;        wy[3] = 0.f;                 wz[3] = SH_C2[3] * x;
  } else if (g == 2) {
    w[0] = SH_C2[4] * (xx - yy);              wx[0] = SH_C2[4] * 2.f * x;  wy[0] = SH_C2[4] * -2.f * y; wz[0] = 0.f;
    w[1] = SH_C3[0] * y * (3.f * xx - yy);    wx[1] = SH_C3[0] * 6.f * xy; wy[1] = SH_C3[0] * 3.f * (xx - yy); wz[1] = 0.f;
    w[2] = SH_C3[1] * xy * z;                 wx[2] = SH_C3[1] * yz;       wy[2] = SH_C3[1] * xz;       wz[2] = SH_C3[1] * xy;
    w[3] = SH_C3[2] * y * (4.f * zz - xx - yy);
    wx[3] = SH_C3[2] * -2.f * xy; wy[3] = SH_C3[2] * (-3.f * yy + 4.f * zz - xx); wz[3] = SH_C3[2] * 8.f * yz;
  } else {
    w[0] = SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
    wx[0] = SH_C3[3] * -6.f * xz; wy[0] = SH_C3[3] * -6.f * yz; wz[0] = SH_C3[3] * 3.f * (2.f * zz - xx - yy);
    w[1] = SH_C3[4] * x * (4.f * zz - xx - yy);
    wx[1] = SH_C3[4] * (-3.f * xx + 4.f * zz - yy); wy[1] = SH_C3[4] * -2.f * xy; wz[1] = SH_C3[4] * 8.f * xz;
    w[2] = SH_C3[5] * z * (xx - yy);          wx[2] = SH_C3[5] * 2.f * xz; wy[2] = SH_C3[5] * -2.f * yz; wz[2] = SH_C3[5] * (xx - yy);
    w[3] = SH_C3[6] * x * (xx - 3.f * yy);    wx[3] = SH_C3[6] * 3.f * (xx - yy); wy[3] = SH_C3[6] * -6.f * xy; wz[3] = 0.f;
  }
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

}  // namespace gsr

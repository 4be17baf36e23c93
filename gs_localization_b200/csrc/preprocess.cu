// Forward preprocess fused with the tile-count prefix sum (one kernel, one pass over the map).
//
// Replaces, from the reference (gaussian_splatting/submodules/diff-gaussian-rasterization):
//   preprocessCUDA            cuda_rasterizer/forward.cu:155-256
//   in_frustum                cuda_rasterizer/auxiliary.h:139-164
//   computeCov3D / Cov2D      cuda_rasterizer/forward.cu:118-152, 74-113
//   computeColorFromSH        cuda_rasterizer/forward.cu:20-71
//   cub::DeviceScan::InclusiveSum + its temp storage   cuda_rasterizer/rasterizer_impl.cu:278
//   checkFrustum              cuda_rasterizer/rasterizer_impl.cu:54-66
//
// B200 design: 256 Gaussians per CTA and NO dependency between CTAs.  Means are staged through
// shared memory with 128-bit coalesced loads (the reference issues stride-3 scalar loads).  The
// CTA packs its visible Gaussians, in order, into the first slots of its own 256-slot segment
// (ballot ranks, no scan), so everything downstream runs on dense records; colour (SH) is then
// evaluated only for those, on dense lanes.  Instead of a prefix sum over tile counts
// (cub::DeviceScan + its 4P+4P bytes + a chained look-back) each visible Gaussian adds its tile
// rectangle to a 2-D difference grid with four atomics; the per-tile list lengths, the tile
// ranges and num_rendered all fall out of one tiny prefix-sum kernel over the tile grid
// (binning.cu).  Arithmetic that decides binning is pinned with IEEE intrinsics to the
// reference's sm_100a rounding sequence (gsr_common.cuh).
#include <cstdlib>

#include "gsr_kernels.cuh"

namespace gsr {

__device__ __forceinline__ void compute_cov3D(const float3 scale, float mod, const float4 rot, float* cov3D) {
  // reference forward.cu:118-152; rounding per oracle/_ref/forward.sass 0x0a50-0x0ef0
  const float sx = __fmul_rn(scale.x, mod), sy = __fmul_rn(scale.y, mod), sz = __fmul_rn(scale.z, mod);
  const float r = rot.x, x = rot.y, y = rot.z, z = rot.w;  // unnormalised (forward.cu:127)
  const float xz = __fmul_rn(x, z), rx = __fmul_rn(r, x), rz = __fmul_rn(r, z);
  const float yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
  const float xz_p_ry = __fmaf_rn(r, y, xz), xz_m_ry = __fmaf_rn(-r, y, xz);
  const float yz_m_rx = __fmaf_rn(y, z, -rx), yz_p_rx = __fmaf_rn(y, z, rx);
  const float xy_m_rz = __fmaf_rn(x, y, -rz), xy_p_rz = __fmaf_rn(x, y, rz);
  const float xx_p_yy = __fmaf_rn(x, x, yy), yy_p_zz = __fadd_rn(yy, zz), xx_p_zz = __fmaf_rn(x, x, zz);
  const float R00 = __fadd_rn(1.f, -__fadd_rn(yy_p_zz, yy_p_zz));
  const float R11 = __fadd_rn(1.f, -__fadd_rn(xx_p_zz, xx_p_zz));
  const float R22 = __fadd_rn(1.f, -__fadd_rn(xx_p_yy, xx_p_yy));
  // M = S*R (glm, column-major): entry (col j, row i) = s_i * R[j][i].  The reference also adds
  // 0*x terms of the diagonal S; for finite inputs they only affect the sign of zero.
  const float m00 = __fmul_rn(sx, R00), m01 = __fmul_rn(sx, __fadd_rn(xy_p_rz, xy_p_rz)),
              m02 = __fmul_rn(sx, __fadd_rn(xz_m_ry, xz_m_ry));
  const float m10 = __fmul_rn(sy, __fadd_rn(xy_m_rz, xy_m_rz)), m11 = __fmul_rn(sy, R11),
              m12 = __fmul_rn(sy, __fadd_rn(yz_p_rx, yz_p_rx));
  const float m20 = __fmul_rn(sz, __fadd_rn(xz_p_ry, xz_p_ry)), m21 = __fmul_rn(sz, __fadd_rn(yz_m_rx, yz_m_rx)),
              m22 = __fmul_rn(sz, R22);
  cov3D[0] = dot3c(m00, m00, m10, m10, m20, m20);
  cov3D[1] = dot3c(m00, m01, m10, m11, m20, m21);
  cov3D[2] = dot3c(m00, m02, m10, m12, m20, m22);
  cov3D[3] = dot3c(m01, m01, m11, m11, m21, m21);
  cov3D[4] = dot3c(m01, m02, m11, m12, m21, m22);
  cov3D[5] = dot3c(m02, m02, m12, m12, m22, m22);
}

// reference forward.cu:74-113; rounding per oracle/_ref/forward.sass 0x1040-0x1a90
__device__ __forceinline__ float3 compute_cov2D(float tx, float ty, float tz, float focal_x, float focal_y,
                                                float tan_fovx, float tan_fovy, const float* c, const float* vm) {
  const float limx = __fmul_rn(1.3f, tan_fovx), limy = __fmul_rn(1.3f, tan_fovy);
  const float txtz = __fdiv_rn(tx, tz), tytz = __fdiv_rn(ty, tz);
  const float cxv = fminf(fmaxf(txtz, -limx), limx), cyv = fminf(fmaxf(tytz, -limy), limy);
  tx = __fmul_rn(cxv, tz);
  ty = __fmul_rn(cyv, tz);
  const float tz2 = __fmul_rn(tz, tz);
  const float J00 = __fdiv_rn(focal_x, tz), J02 = __fdiv_rn(__fmul_rn(-tx, focal_x), tz2);
  const float J11 = __fdiv_rn(focal_y, tz), J12 = __fdiv_rn(__fmul_rn(-ty, focal_y), tz2);
  float T0[3], T1[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const float W0 = vm[4 * i + 0], W1 = vm[4 * i + 1], W2 = vm[4 * i + 2];
    // reference: W0*J00 + W1*0 + W2*J02 and W0*0 + W1*J11 + W2*J12 (zero terms dropped: finite inputs)
    T0[i] = __fmaf_rn(W2, J02, __fmul_rn(W0, J00));
    T1[i] = __fmaf_rn(W2, J12, __fmul_rn(W1, J11));
  }
  const float V[3][3] = {{c[0], c[1], c[2]}, {c[1], c[3], c[4]}, {c[2], c[4], c[5]}};
  float A0[3], A1[3];
#pragma unroll
  for (int j = 0; j < 3; j++) {
    A0[j] = dot3c(T0[0], V[0][j], T0[1], V[1][j], T0[2], V[2][j]);
    A1[j] = dot3c(T1[0], V[0][j], T1[1], V[1][j], T1[2], V[2][j]);
  }
  float3 cov;
  cov.x = __fadd_rn(dot3c(A0[0], T0[0], A0[1], T0[1], A0[2], T0[2]), 0.3f);
  cov.y = dot3c(A1[0], T0[0], A1[1], T0[1], A1[2], T0[2]);
  cov.z = __fadd_rn(dot3c(A1[0], T1[0], A1[1], T1[1], A1[2], T1[2]), 0.3f);
  return cov;
}

// reference forward.cu:20-71.  `sh` points at this Gaussian's M*3 coefficients.  Evaluated in
// groups of four coefficients (= three 16-byte loads) to keep register pressure low.  Also returns
// d(rgb before clamping)/d(unit view direction) (dcol[0..2] = d/dx of r, g, b; [3..5] d/dy; [6..8] d/dz): the SH part of
// the backward (reference backward.cu:20-139) needs the coefficients only through these nine sums.
template <bool VEC4>
__device__ __forceinline__ float3 color_from_sh(int deg, float3 pos, float3 campos, const float* __restrict__ sh,
                                                uint8_t& clamp_mask, float (&dcol)[9]) {
  float3 dir = {pos.x - campos.x, pos.y - campos.y, pos.z - campos.z};
  const float len = sqrtf(__fmaf_rn(dir.z, dir.z, __fmaf_rn(dir.x, dir.x, dir.y * dir.y)));
  const float x = dir.x / len, y = dir.y / len, z = dir.z / len;
  float r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll
  for (int i = 0; i < 9; i++) dcol[i] = 0.f;
  const int ngroups = deg == 0 ? 1 : deg == 1 ? 1 : deg == 2 ? 3 : 4;   // groups of 4 coefficients
#pragma unroll
  for (int g = 0; g < 4; g++) {
    if (g < ngroups) {
      float w[4], wx[4], wy[4], wz[4];
      sh_basis_group(g, x, y, z, w, wx, wy, wz);
      // number of valid coefficients in this group for the active degree
      const int ncoef = (deg + 1) * (deg + 1);
      float c[12];
      if (VEC4) {
        const float4* s4 = reinterpret_cast<const float4*>(sh + 12 * g);
        const float4 v0 = __ldg(s4), v1 = __ldg(s4 + 1), v2 = __ldg(s4 + 2);
        c[0] = v0.x, c[1] = v0.y, c[2] = v0.z, c[3] = v0.w, c[4] = v1.x, c[5] = v1.y, c[6] = v1.z, c[7] = v1.w;
        c[8] = v2.x, c[9] = v2.y, c[10] = v2.z, c[11] = v2.w;
      } else {
#pragma unroll
        for (int q = 0; q < 12; q++) c[q] = (4 * g + q / 3 < ncoef) ? __ldg(sh + 12 * g + q) : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 4; k++) {
        if (4 * g + k < ncoef) {
          r0 += w[k] * c[3 * k];
          r1 += w[k] * c[3 * k + 1];
          r2 += w[k] * c[3 * k + 2];
          dcol[0] += wx[k] * c[3 * k]; dcol[1] += wx[k] * c[3 * k + 1]; dcol[2] += wx[k] * c[3 * k + 2];
          dcol[3] += wy[k] * c[3 * k]; dcol[4] += wy[k] * c[3 * k + 1]; dcol[5] += wy[k] * c[3 * k + 2];
          dcol[6] += wz[k] * c[3 * k]; dcol[7] += wz[k] * c[3 * k + 1]; dcol[8] += wz[k] * c[3 * k + 2];
        }
      }
    }
  }
  r0 += 0.5f, r1 += 0.5f, r2 += 0.5f;
  clamp_mask = (r0 < 0 ? 1 : 0) | (r1 < 0 ? 2 : 0) | (r2 < 0 ? 4 : 0);
  return make_float3(fmaxf(r0, 0.f), fmaxf(r1, 0.f), fmaxf(r2, 0.f));
}

__device__ __forceinline__ float rcp_mufu(float x) {
  float r;
  asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float sqrt_mufu(float x) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// static factor of the conservative radius bound: s_max^2 |R|_F^2 with |R|_F^2 = 1 + 2 (1 - 2|v|^2)^2 + 8 w^2 |v|^2 for the
// reference's unnormalised-quaternion matrix (forward.cu:127-140)
__device__ __forceinline__ float cull_static_factor(float3 sc, float4 q) {
  const float smax = fmaxf(fmaxf(fabsf(sc.x), fabsf(sc.y)), fabsf(sc.z));
  const float v2 = q.y * q.y + q.z * q.z + q.w * q.w;
  const float r2 = 1.0f + 2.0f * (1.0f - 2.0f * v2) * (1.0f - 2.0f * v2) + 8.0f * q.x * q.x * v2;
  return smax * smax * r2;
}

// Static map packed once at load (LoGS localizes hundreds of queries against one read-only map,
// scene/gaussian_model.py:215-256 load_ply): 16 bytes per Gaussian = mean + the static factor of the radius bound.
__global__ void __launch_bounds__(256) build_cull_records_kernel(int P, const float* __restrict__ means3D, const float* __restrict__ scales,
                                                                 const float* __restrict__ rotations, float4* __restrict__ rec) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const float3 sc = make_float3(scales[3 * (size_t)i], scales[3 * (size_t)i + 1], scales[3 * (size_t)i + 2]);
  const float4 q = reinterpret_cast<const float4*>(rotations)[i];
  rec[i] = make_float4(means3D[3 * (size_t)i], means3D[3 * (size_t)i + 1], means3D[3 * (size_t)i + 2], cull_static_factor(sc, q));
}
void launch_build_cull_records(int P, const float* means3D, const float* scales, const float* rotations, float4* rec, cudaStream_t stream) {
  if (P <= 0) return;
  build_cull_records_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, means3D, scales, rotations, rec);
  count_launch();
}

// Pass 1, one thread per Gaussian: depth cull + conservative screen-radius bound; the survivors ("candidates") of
// each 256-Gaussian segment are compacted, in order, into the segment's candidate list.
__global__ void __launch_bounds__(PRE_THREADS) preprocess_cull_kernel(const PreprocessParams p) {
  __shared__ uint32_t s_warp_near[PRE_THREADS / 32];
  __shared__ float s_cam[16 + 16];
  __shared__ float s_w2;
  pdl_trigger();
  pdl_wait();
  const int tid = threadIdx.x;
  const uint32_t lane = tid & 31, warp = tid >> 5;
  const uint32_t block = blockIdx.x;
  const int base = (int)block * PRE_THREADS;
  const int P = p.P;
  // the frame's counters and coverage grid start at zero: done here (this is the first kernel of a forward and the
  // grid's first atomics come from the next one) instead of two memset nodes in front of it
  // (words 16..31 are sticky across forwards: 16 = "some forward overflowed the binning capacity", cleared by gsr_clear_overflow)
  if (block == 0 && tid < 16) p.geom.counters[tid] = 0;
  {
    const uint32_t n_diff = (p.grid_x + 1) * (p.grid_y + 1) * (uint32_t)DIFF_STRIDE;
    for (uint32_t i = block * PRE_THREADS + tid; i < n_diff; i += gridDim.x * PRE_THREADS) p.tile_diff[i] = 0;
  }
  // ---- everything cheap.  The reference culls only on view-space depth
  //      (auxiliary.h:139-164; its x/y frustum test is commented out), so about half of a room-scale map survives
  //      the cull and goes through the covariance pipeline only to end with an empty tile rectangle.  Here a
  //      conservative bound of the screen radius decides first whether the rectangle CAN be non-empty:
  //        lambda_max(cov2D) <= |W|_F^2 |J|_F^2 s_max^2 |R|_F^2 + 0.3,   radius <= 3 sqrt(lambda_max + sqrt(0.1)) + 1
  //      (J with the clamped t of forward.cu:86-91, R the reference's unnormalised-quaternion matrix, whose Frobenius
  //      norm is 1 + 2 (1 - 2|v|^2)^2 + 8 w^2 |v|^2).  A Gaussian whose bounded rectangle misses the image by a
  //      pixel of margin gets radius 0 exactly as the full computation would give it; NaNs fall through to phase 2.
  //      All per-Gaussian inputs are requested up front so that one DRAM round trip covers them.
  bool near = false;
  {
    const int idx = base + tid;
    float px = 0.f, py = 0.f, pz = 0.f;
    float3 sc = {0, 0, 0};
    float4 q = {0, 0, 0, 0};
    float s2r2 = 0.f;          // (largest scale)^2 x squared Frobenius norm of the rotation matrix: the static factor of the radius bound
    if (idx < P) {
      if (p.cull_rec) {
        // static map, packed once at load (gsr_build_cull_records): one 16-byte load instead of 40 bytes in three streams
        const float4 r = __ldg(p.cull_rec + idx);
        px = r.x, py = r.y, pz = r.z, s2r2 = r.w;
      } else {
        px = __ldg(p.means3D + 3 * (size_t)idx), py = __ldg(p.means3D + 3 * (size_t)idx + 1), pz = __ldg(p.means3D + 3 * (size_t)idx + 2);
        if (!p.cov3D_precomp) {
          sc = make_float3(__ldg(p.scales + 3 * (size_t)idx), __ldg(p.scales + 3 * (size_t)idx + 1), __ldg(p.scales + 3 * (size_t)idx + 2));
          q = __ldg(reinterpret_cast<const float4*>(p.rotations) + idx);
          s2r2 = cull_static_factor(sc, q);
        }
      }
    }
    if (tid < 16) s_cam[tid] = p.viewmatrix[tid];
    else if (tid < 32) s_cam[tid] = p.projmatrix[tid - 16];
    else if (tid == 32) {       // squared Frobenius norm of the view rotation: the same for every Gaussian of the launch
      float w2 = 0.f;
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) { const float v = p.viewmatrix[4 * c + r]; w2 += v * v; }
      s_w2 = w2;
    }
    __syncthreads();
    if (idx < P) {
      const float* vm = s_cam;
      const float* pm = s_cam + 16;
      // in_frustum: p_view.z <= 0.2 culls (NaN culls too)
      const float vz = __fadd_rn(dot3c(px, vm[2], py, vm[6], pz, vm[10]), vm[14]);
      if (vz > 0.2f) {
        near = true;
        if (!p.cov3D_precomp) {
          const float hx = __fadd_rn(dot3c(px, pm[0], py, pm[4], pz, pm[8]), pm[12]);
          const float hy = __fadd_rn(dot3c(px, pm[1], py, pm[5], pz, pm[9]), pm[13]);
          const float hw = __fadd_rn(dot3c(px, pm[3], py, pm[7], pz, pm[11]), pm[15]);
          // bare MUFU reciprocals / square root (1-2 ulp): this is a BOUND with 1 % / 2 % / 2 px of slack built in, the exact
          // pipeline runs in pass 2 for everything that survives
          const float p_w = rcp_mufu(hw + 0.0000001f);
          const float cx = ((hx * p_w + 1.0f) * (float)p.W - 1.0f) * 0.5f, cy = ((hy * p_w + 1.0f) * (float)p.H - 1.0f) * 0.5f;
          const float w2 = s_w2;
          const float limx = 1.3f * p.tan_fovx, limy = 1.3f * p.tan_fovy;
          const float iz = rcp_mufu(vz);
          const float j2 = (p.focal_x * iz) * (p.focal_x * iz) * (1.0f + limx * limx) + (p.focal_y * iz) * (p.focal_y * iz) * (1.0f + limy * limy);
          const float lam = w2 * j2 * (s2r2 * p.scale_modifier * p.scale_modifier) * 1.02f + 0.3f + 0.32f;
          const float rb = 3.0f * sqrt_mufu(lam) * 1.01f + 2.0f;
          // outside for sure: the whole [c - rb, c + rb + 15] interval maps to tile index <= 0 or >= grid on one axis
          const bool outside = (cx + rb + 15.0f < 0.0f) || (cx - rb >= 16.0f * (float)p.grid_x) || (cy + rb + 15.0f < 0.0f) ||
                               (cy - rb >= 16.0f * (float)p.grid_y);
          near = !outside;   // NaN/Inf anywhere -> comparisons false -> kept
        }
      } else if (p.prefiltered) {
        // reference auxiliary.h:156-160
        printf("Point is filtered although prefiltered is set. This shouldn't happen!");
        __trap();
      }
      if (!near) p.radii[idx] = 0;
      if (p.n_touched) p.n_touched[idx] = 0;
    }
  }
  const uint32_t m = __ballot_sync(0xffffffffu, near);
  if (lane == 0) s_warp_near[warp] = __popc(m);
  __syncthreads();
  uint32_t off = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < PRE_THREADS / 32; w++) {
    const uint32_t c = s_warp_near[w];
    if (w < (int)warp) off += c;
    tot += c;
  }
  if (near) p.geom.cand[(size_t)base + off + __popc(m & ((1u << lane) - 1u))] = (uint8_t)tid;
  if (tid == 0) p.geom.block_cand[block] = tot;
}

// Pass 2, one WARP per segment, lanes over its candidates: the pinned-rounding covariance pipeline on dense warps, no
// block barrier after the camera constants are staged.  Visible Gaussians are packed into the segment's slots in order
// (ballot ranks + a running count over the warp's iterations).
constexpr int PRE2_WARPS = 8;
__global__ void __launch_bounds__(PRE2_WARPS * 32, 4) preprocess_fwd_kernel(const PreprocessParams p) {
  __shared__ float s_cam[16 + 16];
  pdl_trigger();
  pdl_wait();
  const int tid = threadIdx.x;
  const uint32_t lane = tid & 31, warp = tid >> 5;
  if (tid < 16) s_cam[tid] = p.viewmatrix[tid];
  else if (tid < 32) s_cam[tid] = p.projmatrix[tid - 16];
  __syncthreads();
  const uint32_t block = blockIdx.x * PRE2_WARPS + warp;      // segment of this warp
  if (block >= (uint32_t)((p.P + PRE_THREADS - 1) / PRE_THREADS)) return;
  const int base = (int)block * PRE_THREADS;
  const uint32_t n_cand = p.geom.block_cand[block];
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t nvis = 0, ntiles = 0;
  for (uint32_t c0 = 0; c0 < n_cand; c0 += 32) {
    const bool work = c0 + lane < n_cand;
    const int idx = work ? base + (int)p.geom.cand[(size_t)base + c0 + lane] : p.P;
    float px = 0.f, py = 0.f, pz = 0.f, opacity = 0.f;
    float3 sc = {0, 0, 0};
    float4 q = {0, 0, 0, 0};
    float cov3D[6] = {0, 0, 0, 0, 0, 0};
    if (work) {
      px = __ldg(p.means3D + 3 * (size_t)idx), py = __ldg(p.means3D + 3 * (size_t)idx + 1), pz = __ldg(p.means3D + 3 * (size_t)idx + 2);
      opacity = __ldg(p.opacities + idx);
      if (p.cov3D_precomp) {
#pragma unroll
        for (int k = 0; k < 6; k++) cov3D[k] = __ldg(p.cov3D_precomp + 6 * (size_t)idx + k);
      } else {
        sc = make_float3(__ldg(p.scales + 3 * (size_t)idx), __ldg(p.scales + 3 * (size_t)idx + 1), __ldg(p.scales + 3 * (size_t)idx + 2));
        q = __ldg(reinterpret_cast<const float4*>(p.rotations) + idx);
      }
    }
    uint32_t tiles = 0;
    int radius = 0;
    float vz = 0.f, pix_x = 0.f, pix_y = 0.f;
    float3 conic = {0, 0, 0};
    uint2 rect = {0, 0};
    if (work) {
      const float* vm = s_cam;
      const float* pm = s_cam + 16;
      vz = __fadd_rn(dot3c(px, vm[2], py, vm[6], pz, vm[10]), vm[14]);
      {
      const float vx = __fadd_rn(dot3c(px, vm[0], py, vm[4], pz, vm[8]), vm[12]);
      const float vy = __fadd_rn(dot3c(px, vm[1], py, vm[5], pz, vm[9]), vm[13]);
      const float hx = __fadd_rn(dot3c(px, pm[0], py, pm[4], pz, pm[8]), pm[12]);
      const float hy = __fadd_rn(dot3c(px, pm[1], py, pm[5], pz, pm[9]), pm[13]);
      const float hw = __fadd_rn(dot3c(px, pm[3], py, pm[7], pz, pm[11]), pm[15]);
      const float p_w = __frcp_rn(__fadd_rn(hw, 0.0000001f));
      const float projx = __fmul_rn(hx, p_w), projy = __fmul_rn(hy, p_w);
      if (!p.cov3D_precomp) compute_cov3D(sc, p.scale_modifier, q, cov3D);
      const float3 cov = compute_cov2D(vx, vy, vz, p.focal_x, p.focal_y, p.tan_fovx, p.tan_fovy, cov3D, vm);
      const float det = __fmaf_rn(cov.x, cov.z, -__fmul_rn(cov.y, cov.y));
      if (det != 0.0f) {
        const float det_inv = __frcp_rn(det);
        conic = make_float3(__fmul_rn(cov.z, det_inv), __fmul_rn(cov.y, -det_inv), __fmul_rn(cov.x, det_inv));
        const float mid = __fmul_rn(__fadd_rn(cov.x, cov.z), 0.5f);
        const float s = __fsqrt_rn(fmaxf(__fmaf_rn(mid, mid, -det), 0.1f));
        const float lambda1 = __fadd_rn(mid, s), lambda2 = __fadd_rn(mid, -s);
        const int my_radius = __float2int_ru(__fmul_rn(__fsqrt_rn(fmaxf(lambda1, lambda2)), 3.0f));
        // ndc2Pix in double with one DFMA (auxiliary.h:41-44)
        pix_x = (float)(__dmul_rn(__fma_rn(__dadd_rn((double)projx, 1.0), (double)p.W, -1.0), 0.5));
        pix_y = (float)(__dmul_rn(__fma_rn(__dadd_rn((double)projy, 1.0), (double)p.H, -1.0), 0.5));
        // getRect (auxiliary.h:46-56)
        const float rf = (float)my_radius;
        const uint32_t gx = p.grid_x, gy = p.grid_y;
        const uint32_t minx = min(gx, (uint32_t)max(0, __float2int_rz(__fmul_rn(__fadd_rn(pix_x, -rf), 0.0625f))));
        const uint32_t miny = min(gy, (uint32_t)max(0, __float2int_rz(__fmul_rn(__fadd_rn(pix_y, -rf), 0.0625f))));
        const uint32_t maxx = min(gx, (uint32_t)max(0, __float2int_rz(__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(pix_x, rf), 16.0f), -1.0f), 0.0625f))));
        const uint32_t maxy = min(gy, (uint32_t)max(0, __float2int_rz(__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(pix_y, rf), 16.0f), -1.0f), 0.0625f))));
        const uint32_t cnt = (maxx - minx) * (maxy - miny);
        if (cnt != 0) {
          tiles = cnt;
          radius = my_radius;
          rect = make_uint2(minx | (maxx << 16), miny | (maxy << 16));
        }
      }
      }
      p.radii[idx] = radius;
    }
    const bool vis = tiles != 0;
    const uint32_t vis_mask = __ballot_sync(0xffffffffu, vis);
    ntiles += __reduce_add_sync(0xffffffffu, tiles);
    if (vis) {
      const uint32_t k = (uint32_t)base + nvis + __popc(vis_mask & lt);   // slot: segment start + rank inside the segment
      p.geom.depths[k] = vz;
      p.geom.mean_tau[k] = make_float4(pix_x, pix_y, splat_two_tau(conic.x, conic.y, conic.z, opacity), __uint_as_float(k));   // w: the slot itself, so a staged record is three plain 16-byte copies
      p.geom.conic_opacity[k] = make_float4(conic.x, conic.y, conic.z, opacity);
      p.geom.rect[k] = rect;
      p.geom.gid[k] = (uint32_t)idx;
      p.geom.msr[3 * (size_t)k] = make_float4(px, py, pz, sc.x);
      p.geom.msr[3 * (size_t)k + 1] = make_float4(sc.y, sc.z, q.x, q.y);
      p.geom.msr[3 * (size_t)k + 2] = make_float4(q.z, q.w, 0.f, 0.f);
      if (!p.cov3D_precomp) {
#pragma unroll
        for (int i = 0; i < 6; i++) p.geom.cov3D[6 * (size_t)k + i] = cov3D[i];
      }
      // tile coverage: +1/-1 at the rectangle corners of the 2-D difference grid
      const uint32_t minx = rect.x & 0xffffu, maxx = rect.x >> 16, miny = rect.y & 0xffffu, maxy = rect.y >> 16;
      const uint32_t stride = p.grid_x + 1;
      atomicAdd(p.tile_diff + (size_t)(miny * stride + minx) * DIFF_STRIDE, 1);
      atomicAdd(p.tile_diff + (size_t)(miny * stride + maxx) * DIFF_STRIDE, -1);
      atomicAdd(p.tile_diff + (size_t)(maxy * stride + minx) * DIFF_STRIDE, -1);
      atomicAdd(p.tile_diff + (size_t)(maxy * stride + maxx) * DIFF_STRIDE, 1);
    }
    nvis += __popc(vis_mask);
  }
  if (lane == 0) p.geom.block_vis[block] = nvis, p.geom.block_tiles[block] = ntiles;
}

void launch_preprocess_fwd(const PreprocessParams& p, cudaStream_t stream) {
  const int nb = num_pre_blocks(p.P);
  launch_pdl(preprocess_cull_kernel, dim3(nb), dim3(PRE_THREADS), 0, stream, p);
  launch_pdl(preprocess_fwd_kernel, dim3((nb + PRE2_WARPS - 1) / PRE2_WARPS), dim3(PRE2_WARPS * 32), 0, stream, p);
  count_launch(2);
}

// Colour of the visible Gaussians (SH -> RGB, reference forward.cu:20-71, or the caller's precomputed colours),
// one single-warp CTA per slot segment, dense lanes.  Nothing before the blend needs colours, so this kernel
// runs on a side stream concurrently with the binning kernels (scan_tiles / scatter / tile_sort).
constexpr int COLOR_THREADS = 32;
__global__ void __launch_bounds__(COLOR_THREADS) color_fwd_kernel(const PreprocessParams p) {
  const uint32_t nvis = p.geom.block_vis[blockIdx.x];
  if (nvis == 0) return;
  const float3 cp = {__ldg(p.campos), __ldg(p.campos + 1), __ldg(p.campos + 2)};
  for (uint32_t t = threadIdx.x; t < nvis; t += COLOR_THREADS) {
    const uint32_t k = blockIdx.x * PRE_THREADS + t;
    const size_t g = __ldg(p.geom.gid + k);
    float3 rgb;
    uint8_t cm = 0;
    float dcol[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (p.colors_precomp) {
      rgb = make_float3(__ldg(p.colors_precomp + 3 * g), __ldg(p.colors_precomp + 3 * g + 1), __ldg(p.colors_precomp + 3 * g + 2));
    } else {
      const float3 pos = {__ldg(p.means3D + 3 * g), __ldg(p.means3D + 3 * g + 1), __ldg(p.means3D + 3 * g + 2)};
      const float* sh = p.shs + g * p.M * 3;
      if (p.sh_vec4) rgb = color_from_sh<true>(p.D, pos, cp, sh, cm, dcol);
      else rgb = color_from_sh<false>(p.D, pos, cp, sh, cm, dcol);
      p.geom.shd[3 * (size_t)k] = make_float4(dcol[0], dcol[1], dcol[2], dcol[3]);
      p.geom.shd[3 * (size_t)k + 1] = make_float4(dcol[4], dcol[5], dcol[6], dcol[7]);
      p.geom.shd[3 * (size_t)k + 2] = make_float4(dcol[8], 0.f, 0.f, 0.f);
    }
    p.geom.rgbd[k] = make_float4(rgb.x, rgb.y, rgb.z, p.geom.depths[k]);
    p.geom.clamped[k] = cm;
  }
}

void launch_color_fwd(const PreprocessParams& p, cudaStream_t stream) {
  color_fwd_kernel<<<num_pre_blocks(p.P), COLOR_THREADS, 0, stream>>>(p);
  count_launch();
}

// reference rasterizer_impl.cu:54-66 (checkFrustum): present[i] = p_view.z > 0.2
__global__ void __launch_bounds__(256) mark_visible_kernel(int P, const float* __restrict__ means3D,
                                                           const float* __restrict__ vm, uint8_t* __restrict__ present) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  const float px = __ldg(means3D + 3 * (size_t)idx), py = __ldg(means3D + 3 * (size_t)idx + 1),
              pz = __ldg(means3D + 3 * (size_t)idx + 2);
  const float vz = __fadd_rn(dot3c(px, __ldg(vm + 2), py, __ldg(vm + 6), pz, __ldg(vm + 10)), __ldg(vm + 14));
  present[idx] = vz > 0.2f ? 1 : 0;
}

void launch_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present, cudaStream_t stream) {
  mark_visible_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, means3D, viewmatrix, present);
  count_launch();
}

}  // namespace gsr

// Backward of the per-Gaussian preprocessing, one fused kernel.
//
// Replaces, from the reference (gaussian_splatting/submodules/diff-gaussian-rasterization):
//   computeCov2DCUDA (backward)      cuda_rasterizer/backward.cu:144-274
//   preprocessCUDA (backward)        cuda_rasterizer/backward.cu:346-396
//   computeColorFromSH (backward)    cuda_rasterizer/backward.cu:20-139  (+ dnormvdv auxiliary.h:107-117)
//   computeCov3D (backward)          cuda_rasterizer/backward.cu:278-341
//   the nine torch::zeros fills      rasterize_points.cu:158-166
//
// B200 design: the forward packed the visible Gaussians of every 256-Gaussian segment into the
// segment's first slots, so each CTA here runs only ceil(visible/32) dense warps (the reference
// launches P threads, most of which exit at once); the two reference kernels are fused, so dL_dcov3D / dL_dmean3D never
// round-trip through HBM between them; the dense zero rows the API promises for culled
// Gaussians are written by the blend backward (render.cu), and the rows of the visible ones here, as
// complete 32-byte sectors (store_rows).  Nothing of the map is read: the forward left mean / scale /
// rotation and the SH direction derivatives in the slot.  With dL_dtau != nullptr the SE(3) chain rule is fused in: each
// thread forms its 6-vector contribution to dL/d(rho, theta) for the left perturbation
// T_w2c <- exp(tau) T_w2c (gs_localization/pipelines/tools/pose_utils.py:90-122), the CTA
// reduces it with shuffles and issues 6 atomics.
#include <algorithm>

#include "gsr_kernels.cuh"

namespace gsr {

constexpr int BW_THREADS = 128;  // one warp per preprocess slot segment, looping over its visible slots; four segments per CTA
                                 // so that the pose gradient leaves as 6 atomics per four segments (they all hit one cache line)

// reference auxiliary.h:107-117
__device__ __forceinline__ float3 dnormvdv(float3 v, float3 dv) {
  const float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
  const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
  float3 o;
  o.x = ((+sum2 - v.x * v.x) * dv.x - v.y * v.x * dv.y - v.z * v.x * dv.z) * invsum32;
  o.y = (-v.x * v.y * dv.x + (sum2 - v.y * v.y) * dv.y - v.z * v.y * dv.z) * invsum32;
  o.z = (-v.x * v.z * dv.x - v.y * v.z * dv.y + (sum2 - v.z * v.z) * dv.z) * invsum32;
  return o;
}

// SH backward (reference backward.cu:20-139).  Writes this Gaussian's dL_dsh row to `out` (may be null) and returns
// dL_dmean through the view direction.  The coefficients themselves are not read: dL_dsh_k = basis_k * dL_dRGB needs
// only the direction, and the direction gradient needs them only through the nine sums d(rgb)/d(dir) that the
// forward's colour kernel left in the slot (`dcol`).  Rows are written in groups of four coefficients (three 16-byte
// stores).
__device__ __forceinline__ float3 sh_backward(int deg, float3 pos, float3 campos, const float (&dcol)[9], float3 dL_dRGB, float* out,
                                              bool vec4) {
  const float3 dir_orig = {pos.x - campos.x, pos.y - campos.y, pos.z - campos.z};
  const float inv_len = 1.0f / sqrtf(dir_orig.x * dir_orig.x + dir_orig.y * dir_orig.y + dir_orig.z * dir_orig.z);
  const float x = dir_orig.x * inv_len, y = dir_orig.y * inv_len, z = dir_orig.z * inv_len;
  const int ncoef = (deg + 1) * (deg + 1);
  const int ngroups = (ncoef + 3) / 4;
  if (out) {
#pragma unroll
    for (int g = 0; g < 4; g++) {
      if (g < ngroups) {
        float w[4], wx[4], wy[4], wz[4];
        sh_basis_group(g, x, y, z, w, wx, wy, wz);
        float o[12];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const float wk = 4 * g + k < ncoef ? w[k] : 0.f;
          o[3 * k] = wk * dL_dRGB.x, o[3 * k + 1] = wk * dL_dRGB.y, o[3 * k + 2] = wk * dL_dRGB.z;
        }
        if (vec4) {
          float4* o4 = reinterpret_cast<float4*>(out + 12 * g);
          o4[0] = make_float4(o[0], o[1], o[2], o[3]);
          o4[1] = make_float4(o[4], o[5], o[6], o[7]);
          o4[2] = make_float4(o[8], o[9], o[10], o[11]);
        } else {
#pragma unroll
          for (int q = 0; q < 12; q++)
            if (4 * g + q / 3 < ncoef) out[12 * g + q] = o[q];
        }
      }
    }
  }
  const float3 dL_ddir = {dcol[0] * dL_dRGB.x + dcol[1] * dL_dRGB.y + dcol[2] * dL_dRGB.z,
                          dcol[3] * dL_dRGB.x + dcol[4] * dL_dRGB.y + dcol[5] * dL_dRGB.z,
                          dcol[6] * dL_dRGB.x + dcol[7] * dL_dRGB.y + dcol[8] * dL_dRGB.z};
  return dnormvdv(dir_orig, dL_ddir);
}

// Rows of one gradient tensor ([P, R] floats) for the warp's visible Gaussians, written as FULL 32-byte sectors.
// A row is 4-24 bytes at stride 4R: written on its own it is a partial sector, which L2 has to complete from DRAM before
// it can merge the bytes (the zero rows around it were streamed out long ago) — those read-modify-writes were 16 us of
// this kernel's 37 at the headline.  The rest of a row's sector is known, though: rows of culled Gaussians are zero, and the
// visible neighbours are in the adjacent lanes (slot order is Gaussian order).  The lanes stage their rows in a
// shared-memory window of the segment (256 rows, zeroed around the rows only), and every lane then writes the one or two
// sectors its row touches, complete.  Lanes whose rows share a sector write the same bytes.  Sectors that hold a row of
// another iteration of the warp's loop over the segment (`taint`), that reach behind the tensor's end, or a tensor that is
// not 32-byte aligned take the plain element stores.
#ifndef GSR_PBW_GRAN
#define GSR_PBW_GRAN 8
#endif
constexpr int PBW_GRAN = GSR_PBW_GRAN;          // floats per completed block: 8 = one 32-byte sector
template <int R>
__device__ __forceinline__ void store_rows(float* __restrict__ out, float* __restrict__ stage, uint32_t seg_row0, uint32_t l, bool visible,
                                           const float (&v)[R], int taint_lo_row, int taint_hi_row, uint32_t P) {
  constexpr int G = PBW_GRAN, G4 = G / 4;
  const bool sectors = ((uintptr_t)out & (4u * G - 1u)) == 0;
  const uint32_t f0 = l * R, s0 = f0 / G, s1 = (f0 + R - 1) / G;           // floats / sectors of the row inside the window
  // sectors <= taint_lo or >= taint_hi hold (part of) a row that another iteration of the loop writes
  const int taint_lo = taint_lo_row < 0 ? -1 : (int)(((uint32_t)taint_lo_row * R + R - 1) / G);
  const int taint_hi = taint_hi_row < 0 ? 0x7fffffff : (int)(((uint32_t)taint_hi_row * R) / G);
  const size_t base = (size_t)seg_row0 * R, total = (size_t)P * R;
  auto plain = [&](uint32_t sct) { return !sectors || (int)sct <= taint_lo || (int)sct >= taint_hi || base + (size_t)sct * G + G > total; };
  float4* st4 = reinterpret_cast<float4*>(stage);
  if (visible) {
#pragma unroll
    for (int i = 0; i < G4; i++) st4[G4 * s0 + i] = make_float4(0.f, 0.f, 0.f, 0.f), st4[G4 * s1 + i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncwarp();
  if (visible) {
#pragma unroll
    for (int c = 0; c < R; c++) stage[f0 + c] = v[c];
  }
  __syncwarp();
  if (visible) {
#pragma unroll
    for (int w = 0; w < 2; w++) {
      const uint32_t sct = w ? s1 : s0;
      if (w && s1 == s0) break;
      if (plain(sct)) {
#pragma unroll
        for (int c = 0; c < R; c++)
          if ((f0 + c) / G == sct) out[base + f0 + c] = v[c];
      } else {
        float4* dst = reinterpret_cast<float4*>(out + base + (size_t)sct * G);
#pragma unroll
        for (int i = 0; i < G4; i++) dst[i] = st4[G4 * sct + i];
      }
    }
  }
  __syncwarp();      // the window is reused by the next tensor
}

__global__ void __launch_bounds__(BW_THREADS) preprocess_bwd_kernel(const PreBwdParams p) {
  __shared__ float s_tau[BW_THREADS / 32][6];
  pdl_trigger();
  pdl_wait();

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t seg = blockIdx.x * (BW_THREADS / 32) + warp;
  const uint32_t nvis = seg < (uint32_t)((p.P + PRE_THREADS - 1) / PRE_THREADS) ? p.geom.block_vis[seg] : 0u;
  const int M = p.M;
  const int row = 3 * M;                 // floats per SH row
  // camera constants through shared memory: read through the global pointers they would be re-loaded after every
  // store of the loop (possible aliasing), a dozen serialised round trips per Gaussian
  __shared__ float s_cam[16 + 16 + 16 + 4];
  if (tid < 16) s_cam[tid] = p.viewmatrix[tid];
  else if (tid < 32) s_cam[tid] = p.projmatrix[tid - 16];
  else if (tid < 48) s_cam[tid] = p.projmatrix_raw ? p.projmatrix_raw[tid - 32] : 0.f;
  else if (tid < 51) s_cam[tid] = p.campos[tid - 48];
  __syncthreads();
  const float* vm = s_cam;
  const float* proj = s_cam + 16;
  float tau[6] = {0, 0, 0, 0, 0, 0};

  __shared__ __align__(16) float s_stage[BW_THREADS / 32][PRE_THREADS * 6];     // per warp: a segment's rows of the widest small tensor
  float* const stage = s_stage[warp];
  int prev_last_row = -1;                                // segment-local row of the previous iteration's last Gaussian
  for (uint32_t t0 = 0; t0 < nvis; t0 += 32) {          // warp-uniform trip count: the row stores below are warp-cooperative
  const uint32_t t_in_seg = t0 + lane;
  const bool visible = t_in_seg < nvis;
  const uint32_t k = seg * PRE_THREADS + (visible ? t_in_seg : 0u);     // slot
  // first Gaussian of the next iteration, if there is one (its sectors are left to the plain stores, like the previous one's)
  const int next_first_row = t0 + 32 < nvis ? (int)(__ldg(p.geom.gid + seg * PRE_THREADS + t0 + 32) - seg * PRE_THREADS) : -1;
  float3 g_mean2D = {0, 0, 0};
  float4 g_conic = {0, 0, 0, 0};
  float g_opacity = 0.f;
  float3 g_color = {0, 0, 0};
  float3 g_mean = {0, 0, 0};
  float g_cov[6] = {0, 0, 0, 0, 0, 0};
  float3 g_scale = {0, 0, 0};
  float4 g_rot = {0, 0, 0, 0};
  size_t idx = 0;

  if (visible) {
    idx = __ldg(p.geom.gid + k);
    float4* acc = reinterpret_cast<float4*>(p.geom.grad_acc + 12 * (size_t)k);
    const float4 a0 = acc[0], a1 = acc[1], a2 = acc[2];
    acc[0] = acc[1] = acc[2] = make_float4(0.f, 0.f, 0.f, 0.f);   // leave the row zero for the next backward
    // moments -> gradients: the blend backward accumulates Q S1 (first moments sum u (dx, dy), conic applied per lane) and
    // the second and zeroth moments sum u * (dx^2, dx dy, dy^2, 1) with
    // u = G * dL/dalpha; the per-Gaussian factors of backward.cu:549-578 (dL_dG = opacity * ..., dG/ddel through the conic,
    // ddelx_dx = W/2, the -0.5 of the conic terms) are applied here, once per Gaussian instead of once per pixel.
    {
      const float4 co = __ldg(p.geom.conic_opacity + k);
      const float ddelx_dx = 0.5f * p.W, ddely_dy = 0.5f * p.H;
      g_mean2D = make_float3(ddelx_dx * co.w * -a0.x, ddely_dy * co.w * -a0.y, 0.f);   // a0.xy = Q S, the conic already applied per lane
      const float mh = -0.5f * co.w;
      g_conic = make_float4(mh * a0.z, mh * a0.w, 0.f, mh * a1.x);
      g_opacity = a1.y;
    }
    g_color = make_float3(a1.z, a1.w, a2.x);
    const float g_depth = a2.y;

    // everything this Gaussian needs is in its slot (the forward left mean / scale / rotation and the SH direction
    // derivatives there): coalesced loads, one round trip; the map itself is not read
    float dcol[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (p.shs) {
      const float4 d0 = __ldg(p.geom.shd + 3 * (size_t)k), d1 = __ldg(p.geom.shd + 3 * (size_t)k + 1), d2 = __ldg(p.geom.shd + 3 * (size_t)k + 2);
      dcol[0] = d0.x, dcol[1] = d0.y, dcol[2] = d0.z, dcol[3] = d0.w, dcol[4] = d1.x, dcol[5] = d1.y, dcol[6] = d1.z, dcol[7] = d1.w, dcol[8] = d2.x;
    }
    // mean, scale and rotation were left in the slot by the forward: three coalesced 16-byte loads, no second gather by id
    const float4 m0 = __ldg(p.geom.msr + 3 * (size_t)k), m1 = __ldg(p.geom.msr + 3 * (size_t)k + 1), m2 = __ldg(p.geom.msr + 3 * (size_t)k + 2);
    const float3 mean = {m0.x, m0.y, m0.z};
    const float3 sc_in = {m0.w, m1.x, m1.y};
    const float4 q_in = {m1.z, m1.w, m2.x, m2.y};
    const uint8_t cm = p.shs ? __ldg(p.geom.clamped + k) : (uint8_t)0;
    const float* c3 = p.cov3D_precomp ? p.cov3D_precomp + 6 * idx : p.geom.cov3D + 6 * (size_t)k;
    float cov3D[6];
#pragma unroll
    for (int q = 0; q < 6; q++) cov3D[q] = __ldg(c3 + q);

    // ---------------- computeCov2DCUDA backward (backward.cu:144-274)
    float3 t = {vm[0] * mean.x + vm[4] * mean.y + vm[8] * mean.z + vm[12],
                vm[1] * mean.x + vm[5] * mean.y + vm[9] * mean.z + vm[13],
                vm[2] * mean.x + vm[6] * mean.y + vm[10] * mean.z + vm[14]};
    const float3 t_orig = t;
    const float limx = 1.3f * p.tan_fovx, limy = 1.3f * p.tan_fovy;
    const float txtz = t.x / t.z, tytz = t.y / t.z;
    t.x = fminf(limx, fmaxf(-limx, txtz)) * t.z;
    t.y = fminf(limy, fmaxf(-limy, tytz)) * t.z;
    const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
    const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
    const float h_x = p.focal_x, h_y = p.focal_y;
    const float J00 = h_x / t.z, J02 = -(h_x * t.x) / (t.z * t.z), J11 = h_y / t.z, J12 = -(h_y * t.y) / (t.z * t.z);
    // Rw[k][i] = W2C rotation entry (row k, col i) = vm[4*i + k]
    float T0[3], T1[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      T0[i] = vm[4 * i + 0] * J00 + vm[4 * i + 2] * J02;
      T1[i] = vm[4 * i + 1] * J11 + vm[4 * i + 2] * J12;
    }
    const float V[3][3] = {{cov3D[0], cov3D[1], cov3D[2]}, {cov3D[1], cov3D[3], cov3D[4]}, {cov3D[2], cov3D[4], cov3D[5]}};
    float TV0[3], TV1[3];   // T0 . V[:,j], T1 . V[:,j]
#pragma unroll
    for (int j = 0; j < 3; j++) {
      TV0[j] = T0[0] * V[0][j] + T0[1] * V[1][j] + T0[2] * V[2][j];
      TV1[j] = T1[0] * V[0][j] + T1[1] * V[1][j] + T1[2] * V[2][j];
    }
    const float a = TV0[0] * T0[0] + TV0[1] * T0[1] + TV0[2] * T0[2] + 0.3f;
    const float b = TV1[0] * T0[0] + TV1[1] * T0[1] + TV1[2] * T0[2];
    const float c = TV1[0] * T1[0] + TV1[1] * T1[1] + TV1[2] * T1[2] + 0.3f;
    const float denom = a * c - b * b;
    float dL_da = 0, dL_db = 0, dL_dc = 0;
    const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
    const float3 dL_dconic = {g_conic.x, g_conic.y, g_conic.w};
    if (denom2inv != 0) {
      dL_da = denom2inv * (-c * c * dL_dconic.x + 2 * b * c * dL_dconic.y + (denom - a * c) * dL_dconic.z);
      dL_dc = denom2inv * (-a * a * dL_dconic.z + 2 * a * b * dL_dconic.y + (denom - a * c) * dL_dconic.x);
      dL_db = denom2inv * 2 * (b * c * dL_dconic.x - (denom + 2 * b * b) * dL_dconic.y + a * b * dL_dconic.z);
      g_cov[0] = (T0[0] * T0[0] * dL_da + T0[0] * T1[0] * dL_db + T1[0] * T1[0] * dL_dc);
      g_cov[3] = (T0[1] * T0[1] * dL_da + T0[1] * T1[1] * dL_db + T1[1] * T1[1] * dL_dc);
      g_cov[5] = (T0[2] * T0[2] * dL_da + T0[2] * T1[2] * dL_db + T1[2] * T1[2] * dL_dc);
      g_cov[1] = 2 * T0[0] * T0[1] * dL_da + (T0[0] * T1[1] + T0[1] * T1[0]) * dL_db + 2 * T1[0] * T1[1] * dL_dc;
      g_cov[2] = 2 * T0[0] * T0[2] * dL_da + (T0[0] * T1[2] + T0[2] * T1[0]) * dL_db + 2 * T1[0] * T1[2] * dL_dc;
      g_cov[4] = 2 * T0[2] * T0[1] * dL_da + (T0[1] * T1[2] + T0[2] * T1[1]) * dL_db + 2 * T1[1] * T1[2] * dL_dc;
    }
    float dL_dT0[3], dL_dT1[3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
      dL_dT0[j] = 2 * TV0[j] * dL_da + TV1[j] * dL_db;
      dL_dT1[j] = 2 * TV1[j] * dL_dc + TV0[j] * dL_db;
    }
    const float dL_dJ00 = vm[0] * dL_dT0[0] + vm[4] * dL_dT0[1] + vm[8] * dL_dT0[2];
    const float dL_dJ02 = vm[2] * dL_dT0[0] + vm[6] * dL_dT0[1] + vm[10] * dL_dT0[2];
    const float dL_dJ11 = vm[1] * dL_dT1[0] + vm[5] * dL_dT1[1] + vm[9] * dL_dT1[2];
    const float dL_dJ12 = vm[2] * dL_dT1[0] + vm[6] * dL_dT1[1] + vm[10] * dL_dT1[2];
    const float tz = 1.f / t.z, tz2 = tz * tz, tz3 = tz2 * tz;
    const float dL_dtx = x_grad_mul * -h_x * tz2 * dL_dJ02;
    const float dL_dty = y_grad_mul * -h_y * tz2 * dL_dJ12;
    const float dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * t.x) * tz3 * dL_dJ02 + (2 * h_y * t.y) * tz3 * dL_dJ12;
    // transformVec4x3Transpose (auxiliary.h:89-97)
    g_mean.x = vm[0] * dL_dtx + vm[1] * dL_dty + vm[2] * dL_dtz;
    g_mean.y = vm[4] * dL_dtx + vm[5] * dL_dty + vm[6] * dL_dtz;
    g_mean.z = vm[8] * dL_dtx + vm[9] * dL_dty + vm[10] * dL_dtz;

    // ---------------- mean2D -> mean3D through projmatrix (backward.cu:370-387)
    const float hw = proj[3] * mean.x + proj[7] * mean.y + proj[11] * mean.z + proj[15];
    const float m_w = 1.0f / (hw + 0.0000001f);
    const float mul1 = (proj[0] * mean.x + proj[4] * mean.y + proj[8] * mean.z + proj[12]) * m_w * m_w;
    const float mul2 = (proj[1] * mean.x + proj[5] * mean.y + proj[9] * mean.z + proj[13]) * m_w * m_w;
    g_mean.x += (proj[0] * m_w - proj[3] * mul1) * g_mean2D.x + (proj[1] * m_w - proj[3] * mul2) * g_mean2D.y;
    g_mean.y += (proj[4] * m_w - proj[7] * mul1) * g_mean2D.x + (proj[5] * m_w - proj[7] * mul2) * g_mean2D.y;
    g_mean.z += (proj[8] * m_w - proj[11] * mul1) * g_mean2D.x + (proj[9] * m_w - proj[11] * mul2) * g_mean2D.y;

    // ---------------- SH (backward.cu:389-391)
    float3 g_sh_mean = {0, 0, 0};
    if (p.shs) {
      const float3 dL_dRGB = {(cm & 1) ? 0.f : g_color.x, (cm & 2) ? 0.f : g_color.y, (cm & 4) ? 0.f : g_color.z};
      const float3 cp = {s_cam[48], s_cam[49], s_cam[50]};
      // rows of the SH gradient were zero-filled; coefficients above the active degree stay zero
      const bool vec4 = (M % 4 == 0) && ((uintptr_t)p.dL_dsh % 16 == 0);
      g_sh_mean = sh_backward(p.D, mean, cp, dcol, dL_dRGB, p.dL_dsh ? p.dL_dsh + idx * row : nullptr, vec4);
      g_mean.x += g_sh_mean.x; g_mean.y += g_sh_mean.y; g_mean.z += g_sh_mean.z;
    }

    // ---------------- computeCov3D backward (backward.cu:278-341)
    if (p.scales) {
      const float3 sc = sc_in;
      const float4 q = q_in;
      const float r = q.x, x = q.y, y = q.z, z = q.w;
      // glm columns of R
      const float Rm[3][3] = {{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                              {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                              {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
      const float s[3] = {p.scale_modifier * sc.x, p.scale_modifier * sc.y, p.scale_modifier * sc.z};
      const float dS[3][3] = {{g_cov[0], 0.5f * g_cov[1], 0.5f * g_cov[2]},
                              {0.5f * g_cov[1], g_cov[3], 0.5f * g_cov[4]},
                              {0.5f * g_cov[2], 0.5f * g_cov[4], g_cov[5]}};
      // dL_dM[j][i] = 2 * sum_k M[k][i] dS[j][k], M[k][i] = s_i Rm[k][i];  dMt[j][i] = dM[i][j]
      float dMt[3][3];
#pragma unroll
      for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < 3; i++)
          dMt[i][j] = 2.f * s[i] * (Rm[0][i] * dS[j][0] + Rm[1][i] * dS[j][1] + Rm[2][i] * dS[j][2]);
      g_scale.x = Rm[0][0] * dMt[0][0] + Rm[1][0] * dMt[0][1] + Rm[2][0] * dMt[0][2];
      g_scale.y = Rm[0][1] * dMt[1][0] + Rm[1][1] * dMt[1][1] + Rm[2][1] * dMt[1][2];
      g_scale.z = Rm[0][2] * dMt[2][0] + Rm[1][2] * dMt[2][1] + Rm[2][2] * dMt[2][2];
#pragma unroll
      for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < 3; i++) dMt[j][i] *= s[j];
      g_rot.x = 2 * z * (dMt[0][1] - dMt[1][0]) + 2 * y * (dMt[2][0] - dMt[0][2]) + 2 * x * (dMt[1][2] - dMt[2][1]);
      g_rot.y = 2 * y * (dMt[1][0] + dMt[0][1]) + 2 * z * (dMt[2][0] + dMt[0][2]) + 2 * r * (dMt[1][2] - dMt[2][1]) - 4 * x * (dMt[2][2] + dMt[1][1]);
      g_rot.z = 2 * x * (dMt[1][0] + dMt[0][1]) + 2 * r * (dMt[2][0] - dMt[0][2]) + 2 * z * (dMt[1][2] + dMt[2][1]) - 4 * y * (dMt[2][2] + dMt[0][0]);
      g_rot.w = 2 * r * (dMt[0][1] - dMt[1][0]) + 2 * x * (dMt[2][0] + dMt[0][2]) + 2 * y * (dMt[1][2] + dMt[2][1]) - 4 * z * (dMt[1][1] + dMt[0][0]);
    }

    // ---------------- SE(3) chain rule (pose extension), camera-space derivation
    if (p.dL_dtau) {
      const float* raw = s_cam + 32;
      // (1) screen-space mean through projmatrix_raw applied to the camera-space point
      const float pcx = t_orig.x, pcy = t_orig.y, pcz = t_orig.z;
      const float rw = raw[3] * pcx + raw[7] * pcy + raw[11] * pcz + raw[15];
      const float r_w = 1.0f / (rw + 0.0000001f);
      const float rmul1 = (raw[0] * pcx + raw[4] * pcy + raw[8] * pcz + raw[12]) * r_w * r_w;
      const float rmul2 = (raw[1] * pcx + raw[5] * pcy + raw[9] * pcz + raw[13]) * r_w * r_w;
      float gx = (raw[0] * r_w - raw[3] * rmul1) * g_mean2D.x + (raw[1] * r_w - raw[3] * rmul2) * g_mean2D.y;
      float gy = (raw[4] * r_w - raw[7] * rmul1) * g_mean2D.x + (raw[5] * r_w - raw[7] * rmul2) * g_mean2D.y;
      float gz = (raw[8] * r_w - raw[11] * rmul1) * g_mean2D.x + (raw[9] * r_w - raw[11] * rmul2) * g_mean2D.y;
      // (2) covariance path through t, (3) rendered depth through p_c.z
      gx += dL_dtx; gy += dL_dty; gz += dL_dtz + g_depth;
      tau[0] += gx; tau[1] += gy; tau[2] += gz;
      tau[3] += pcy * gz - pcz * gy;
      tau[4] += pcz * gx - pcx * gz;
      tau[5] += pcx * gy - pcy * gx;
      // (4) covariance path through the rotation W: G[m][i] = dL/dR[m][i], Q = R G^T, dtheta = axial(Q - Q^T)
      float G[3][3];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        G[0][i] = J00 * dL_dT0[i];
        G[1][i] = J11 * dL_dT1[i];
        G[2][i] = J02 * dL_dT0[i] + J12 * dL_dT1[i];
      }
      auto Q = [&](int i, int j) {   // sum_m R[i][m] G[j][m], R[i][m] = vm[4*m + i]
        return vm[i] * G[j][0] + vm[4 + i] * G[j][1] + vm[8 + i] * G[j][2];
      };
      tau[3] += Q(1, 2) - Q(2, 1);
      tau[4] += Q(2, 0) - Q(0, 2);
      tau[5] += Q(0, 1) - Q(1, 0);
      // (5) SH view direction: the camera centre moves by -R^T rho
      tau[0] += vm[0] * g_sh_mean.x + vm[4] * g_sh_mean.y + vm[8] * g_sh_mean.z;
      tau[1] += vm[1] * g_sh_mean.x + vm[5] * g_sh_mean.y + vm[9] * g_sh_mean.z;
      tau[2] += vm[2] * g_sh_mean.x + vm[6] * g_sh_mean.y + vm[10] * g_sh_mean.z;
    }
  }

  // ---------------- rows of the visible Gaussians (culled rows were zero-filled by the blend backward)
  {
    const uint32_t row0 = seg * PRE_THREADS, l = visible ? (uint32_t)idx - row0 : 0u;
    const uint32_t Pn = (uint32_t)p.P;
    if (p.dL_dmean2D) { const float v[3] = {g_mean2D.x, g_mean2D.y, 0.f}; store_rows<3>(p.dL_dmean2D, stage, row0, l, visible, v, prev_last_row, next_first_row, Pn); }
    if (p.dL_dconic) { const float v[4] = {g_conic.x, g_conic.y, g_conic.z, g_conic.w}; store_rows<4>(p.dL_dconic, stage, row0, l, visible, v, prev_last_row, next_first_row, Pn); }
    if (p.dL_dopacity) { const float v[1] = {g_opacity}; store_rows<1>(p.dL_dopacity, stage, row0, l, visible, v, prev_last_row, next_first_row, Pn); }
    if (p.dL_dcolor) { const float v[3] = {g_color.x, g_color.y, g_color.z}; store_rows<3>(p.dL_dcolor, stage, row0, l, visible, v, prev_last_row, next_first_row, Pn); }
    if (p.dL_dmean3D) { const float v[3] = {g_mean.x, g_mean.y, g_mean.z}; store_rows<3>(p.dL_dmean3D, stage, row0, l, visible, v, prev_last_row, next_first_row, Pn); }
    if (p.dL_dcov3D) { const float v[6] = {g_cov[0], g_cov[1], g_cov[2], g_cov[3], g_cov[4], g_cov[5]}; store_rows<6>(p.dL_dcov3D, stage, row0, l, visible, v, prev_last_row, next_first_row, Pn); }
    if (p.dL_dscale) { const float v[3] = {g_scale.x, g_scale.y, g_scale.z}; store_rows<3>(p.dL_dscale, stage, row0, l, visible, v, prev_last_row, next_first_row, Pn); }
    if (p.dL_drot) { const float v[4] = {g_rot.x, g_rot.y, g_rot.z, g_rot.w}; store_rows<4>(p.dL_drot, stage, row0, l, visible, v, prev_last_row, next_first_row, Pn); }
    prev_last_row = (int)__shfl_sync(0xffffffffu, l, 31);      // only used when there is a next iteration (lane 31 is then visible)
  }
  }  // loop over the segment's visible slots

  // ---------------- pose gradient: CTA reduction, 6 atomics
  if (p.dL_dtau) {
#pragma unroll
    for (int c = 0; c < 6; c++) {
      float v = tau[c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) s_tau[warp][c] = v;
    }
    __syncthreads();
    if (tid < 6) {
      float v = 0.f;
      for (int w = 0; w < BW_THREADS / 32; w++) v += s_tau[w][tid];
      if (v != 0.f) atomicAdd(p.dL_dtau + tid, v);
    }
  }
}

void launch_preprocess_bwd(const PreBwdParams& p, cudaStream_t stream) {
  if (p.P <= 0) return;
  launch_pdl(preprocess_bwd_kernel, dim3((num_pre_blocks(p.P) + BW_THREADS / 32 - 1) / (BW_THREADS / 32)), dim3(BW_THREADS), 0, stream, p);
  count_launch();
}

}  // namespace gsr

// Optimiser side of a LoGS map-training iteration, fused (gs_localization/gs/7scenes_gs_full_dslam.py:225-242 with
// gaussian_splatting/scene/gaussian_model.py:44-58,96-115,152-168,405-407):
//
//   * chain rule through the parameter activations (sigmoid opacity, exp scaling, normalised rotation), which the
//     reference leaves to autograd as ~10 element-wise kernels over P-sized tensors;
//   * torch.optim.Adam (betas 0.9/0.999, eps 1e-15, one learning rate per parameter group) on all six groups;
//   * re-activation of the updated parameters into the buffers the rasterizer reads next iteration;
//   * the densification statistics (max_radii2D, xyz_gradient_accum, denom over the visible set).
//
// Two launches: one thread per Gaussian for the 11 geometric scalars + statistics, one float4 per thread for the
// SH features, whose DC and higher-order coefficients live in ONE [P,M,3] tensor (no torch.cat per render, no split
// of its gradient) with the group's learning rate chosen per coefficient.  HBM-bound: 7 passes over 59 floats per
// Gaussian (read p, g, m, v; write p, m, v) — 1.65 KB per Gaussian at SH degree 3.
#include <algorithm>
#include <cmath>

#include "gsr_kernels.cuh"

namespace gsr {

struct AdamScalars {
  float w1;           // 1 - beta1   (exp_avg.lerp_(grad, 1 - beta1))
  float b2, w2;       // beta2, 1 - beta2
  float eps;
  // per group (xyz, f_dc, f_rest, opacity, scaling, rotation): torch keeps one step counter per parameter, and a group
  // whose parameter was just replaced (reset_opacity) misses a step
  float inv_bc1[6];   // 1 / (1 - beta1^t)
  float bc2_sqrt[6];  // sqrt(1 - beta2^t)
};

__device__ __forceinline__ float adam_update(float p, float g, float& m, float& v, float lr, const AdamScalars& a, int group) {
  m = m + a.w1 * (g - m);
  v = v * a.b2 + a.w2 * g * g;
  const float denom = sqrtf(v) / a.bc2_sqrt[group] + a.eps;
  return p - (lr * a.inv_bc1[group]) * (m / denom);
}

__global__ void __launch_bounds__(256) map_geometry_step_kernel(MapStepParams s, AdamScalars a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= s.P) return;
  if (s.do_stats) {
    const int r = s.radii[i];
    if (r > 0) {
      s.max_radii2D[i] = fmaxf(s.max_radii2D[i], (float)r);
      const float gx = s.g_means2D[3 * (size_t)i], gy = s.g_means2D[3 * (size_t)i + 1];
      s.xyz_gradient_accum[i] += sqrtf(gx * gx + gy * gy);
      s.denom[i] += 1.f;
    }
  }
  if (!s.do_adam) return;
  if (s.lr_xyz >= 0.f) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const size_t e = 3 * (size_t)i + c;
      float m = s.m_xyz[e], v = s.v_xyz[e];
      s.xyz[e] = adam_update(s.xyz[e], s.g_xyz[e], m, v, s.lr_xyz, a, 0);
      s.m_xyz[e] = m, s.v_xyz[e] = v;
    }
  }
  if (s.lr_opacity >= 0.f) {
    const float raw = s.opacity[i];
    const float sig = 1.f / (1.f + expf(-raw));
    float m = s.m_opacity[i], v = s.v_opacity[i];
    const float nraw = adam_update(raw, s.g_opacity[i] * sig * (1.f - sig), m, v, s.lr_opacity, a, 3);
    s.opacity[i] = nraw, s.m_opacity[i] = m, s.v_opacity[i] = v;
    s.opacity_act[i] = 1.f / (1.f + expf(-nraw));
  }
  if (s.lr_scaling >= 0.f) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const size_t e = 3 * (size_t)i + c;
      const float raw = s.scaling[e];
      float m = s.m_scaling[e], v = s.v_scaling[e];
      const float nraw = adam_update(raw, s.g_scaling[e] * expf(raw), m, v, s.lr_scaling, a, 4);
      s.scaling[e] = nraw, s.m_scaling[e] = m, s.v_scaling[e] = v;
      s.scaling_act[e] = expf(nraw);
    }
  }
  if (s.lr_rotation >= 0.f) {
    const float4 q = reinterpret_cast<const float4*>(s.rotation)[i];
    const float4 g = reinterpret_cast<const float4*>(s.g_rotation)[i];
    float4 m = reinterpret_cast<float4*>(s.m_rotation)[i], v = reinterpret_cast<float4*>(s.v_rotation)[i];
    // backward of F.normalize (eps 1e-12): (g - n (n.g)) / max(|q|, eps)
    const float len = fmaxf(sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e-12f);
    const float4 n = make_float4(q.x / len, q.y / len, q.z / len, q.w / len);
    const float ng = n.x * g.x + n.y * g.y + n.z * g.z + n.w * g.w;
    float4 nq;
    nq.x = adam_update(q.x, (g.x - n.x * ng) / len, m.x, v.x, s.lr_rotation, a, 5);
    nq.y = adam_update(q.y, (g.y - n.y * ng) / len, m.y, v.y, s.lr_rotation, a, 5);
    nq.z = adam_update(q.z, (g.z - n.z * ng) / len, m.z, v.z, s.lr_rotation, a, 5);
    nq.w = adam_update(q.w, (g.w - n.w * ng) / len, m.w, v.w, s.lr_rotation, a, 5);
    reinterpret_cast<float4*>(s.rotation)[i] = nq;
    reinterpret_cast<float4*>(s.m_rotation)[i] = m;
    reinterpret_cast<float4*>(s.v_rotation)[i] = v;
    const float nl = fmaxf(sqrtf(nq.x * nq.x + nq.y * nq.y + nq.z * nq.z + nq.w * nq.w), 1e-12f);
    reinterpret_cast<float4*>(s.rotation_act)[i] = make_float4(nq.x / nl, nq.y / nl, nq.z / nl, nq.w / nl);
  }
}

// SH features [P,M,3]: coefficient 0 is the f_dc group, 1..M-1 the f_rest group
__global__ void __launch_bounds__(256) map_features_step_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m,
                                                                float4* __restrict__ v, size_t n, int floats_per_gaussian, float lr_dc,
                                                                float lr_rest, AdamScalars a) {
  const size_t n4 = n >> 2;
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {   // tail when P*M*3 is not a multiple of 4 (SH degree 0 or 2)
    const size_t e = 4 * n4 + threadIdx.x;
    const float lr = (int)(e % (size_t)floats_per_gaussian) < 3 ? lr_dc : lr_rest;
    float* ps = reinterpret_cast<float*>(p);
    float* ms = reinterpret_cast<float*>(m);
    float* vs = reinterpret_cast<float*>(v);
    if (lr >= 0.f) ps[e] = adam_update(ps[e], reinterpret_cast<const float*>(g)[e], ms[e], vs[e], lr, a, (int)(e % (size_t)floats_per_gaussian) < 3 ? 1 : 2);
  }
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += (size_t)gridDim.x * blockDim.x) {
    const int within = (int)((4 * t) % (size_t)floats_per_gaussian);
    float4 pp = p[t], mm = m[t], vv = v[t];
    const float4 gg = g[t];
    float* pa = &pp.x;
    float* ma = &mm.x;
    float* va = &vv.x;
    const float* ga = &gg.x;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const bool dc = ((within + k) % floats_per_gaussian) < 3;
      const float lr = dc ? lr_dc : lr_rest;
      if (lr >= 0.f) pa[k] = adam_update(pa[k], ga[k], ma[k], va[k], lr, a, dc ? 1 : 2);
    }
    p[t] = pp, m[t] = mm, v[t] = vv;
  }
}

void launch_map_step(const MapStepParams& s, float beta1, float beta2, float eps, const int* steps, cudaStream_t stream) {
  if (s.P <= 0) return;
  AdamScalars a;
  a.w1 = (float)(1.0 - (double)beta1);
  a.b2 = beta2;
  a.w2 = (float)(1.0 - (double)beta2);
  a.eps = eps;
  for (int g = 0; g < 6; g++) {
    const double t = (double)(steps ? std::max(steps[g], 1) : 1);
    a.inv_bc1[g] = (float)(1.0 / (1.0 - pow((double)beta1, t)));
    a.bc2_sqrt[g] = (float)sqrt(1.0 - pow((double)beta2, t));
  }
  map_geometry_step_kernel<<<(s.P + 255) / 256, 256, 0, stream>>>(s, a);
  count_launch();
  if (s.do_adam && s.features && (s.lr_f_dc >= 0.f || s.lr_f_rest >= 0.f)) {
    const size_t n = (size_t)s.P * s.M * 3;
    const int blocks = (int)std::min<size_t>((n / 4 + 255) / 256 + 1, (size_t)sm_count() * 16);
    map_features_step_kernel<<<blocks, 256, 0, stream>>>((float4*)s.features, (const float4*)s.g_features, (float4*)s.m_features,
                                                         (float4*)s.v_features, n, s.M * 3, s.lr_f_dc, s.lr_f_rest, a);
    count_launch();
  }
}

// ---- visible-rows gradient exchange of data-parallel map training (parallel.SparseGradientExchange)
// table row = [row id as int bits | xyz 3 | features 3M | opacity 1 | scaling 3 | rotation 4]; one warp per row.
__device__ __forceinline__ void row_column(int c, int M, int& tensor, int& within) {
  const int f = 3 * M;
  if (c < 3) tensor = 0, within = c;
  else if (c < 3 + f) tensor = 1, within = c - 3;
  else if (c < 4 + f) tensor = 2, within = 0;
  else if (c < 7 + f) tensor = 3, within = c - 4 - f;
  else tensor = 4, within = c - 7 - f;
}

__global__ void __launch_bounds__(256) pack_gradient_rows_kernel(const long long* __restrict__ idx, int k, int K, int M, GradRowTensors t,
                                                                 float* __restrict__ table) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= K) return;
  const int F = 11 + 3 * M;
  float* out = table + (size_t)row * (1 + F);
  if (row >= k) {
    if (lane == 0) out[0] = __int_as_float(-1);
    return;
  }
  const long long g = idx[row];
  if (lane == 0) out[0] = __int_as_float((int)g);
  const int width[5] = {3, 3 * M, 1, 3, 4};
  for (int c = lane; c < F; c += 32) {
    int ten, w;
    row_column(c, M, ten, w);
    out[1 + c] = t.g[ten][(size_t)g * width[ten] + w];
  }
}

__global__ void __launch_bounds__(256) add_gradient_rows_kernel(const float* __restrict__ table, int K, int M, GradRowTensors t, int P) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= K) return;
  const int F = 11 + 3 * M;
  const float* in = table + (size_t)row * (1 + F);
  const int g = __float_as_int(in[0]);
  if (g < 0) return;
  const int width[5] = {3, 3 * M, 1, 3, 4};
  for (int c = lane; c < F; c += 32) {
    int ten, w;
    row_column(c, M, ten, w);
    t.g[ten][(size_t)g * width[ten] + w] += in[1 + c];   // row ids are unique within one rank's table
  }
}

// ---- the same exchange without a host round trip: the visible rows (radii > 0) are appended through a device counter
// into a table of fixed capacity; row `capacity` is the header (word 0: number of rows the view produced, which the
// receivers clamp to the capacity and the host checks one step later).  Columns after the gradients: the screen-space
// gradient (x, y) and the radius of the view, i.e. what the densification statistics of gaussian_model.py:405-407 need,
// so that every replica accumulates the statistics of ALL views of the step and takes identical densification decisions.
constexpr int PACK_ROWS = 4;
__global__ void __launch_bounds__(256) pack_visible_rows_kernel(const int* __restrict__ radii, int P, int M, GradRowTensors t,
                                                                const float* __restrict__ g_means2D, float* __restrict__ table,
                                                                int capacity, unsigned int* __restrict__ count) {
  const int F = 11 + 3 * M, W = 1 + F + 3;
  const int lane = threadIdx.x & 31;
  const int warp_first = (blockIdx.x * blockDim.x + threadIdx.x) & ~31;
  if (warp_first >= P) return;
  const int i = warp_first + lane;
  const bool vis = i < P && radii[i] > 0;
  const unsigned int bal = __ballot_sync(0xffffffffu, vis);
  if (!bal) return;
  unsigned int base = 0;
  if (lane == 0) base = atomicAdd(count, (unsigned int)__popc(bal));
  base = __shfl_sync(0xffffffffu, base, 0);
  // The warp writes its visible rows PACK_ROWS at a time: all their columns are requested first (read-only path, so the
  // loads do not wait for the stores of the previous group), then stored — one DRAM round trip per group instead of per row.
  const int width[5] = {3, 3 * M, 1, 3, 4};
  unsigned int m = bal, k = 0;
  while (m) {
    int g[PACK_ROWS];
    unsigned int row[PACK_ROWS];
#pragma unroll
    for (int r = 0; r < PACK_ROWS; r++) {
      const bool have = m != 0;
      g[r] = have ? warp_first + (__ffs(m) - 1) : -1;
      m &= m - 1;                                 // no-op on 0
      row[r] = base + k;
      if (have) k++;
      if (row[r] >= (unsigned int)capacity) g[r] = -1;
    }
    for (int c0 = 0; c0 < F; c0 += 32) {
      const int c = c0 + lane;
      int ten = 0, w = 0;
      if (c < F) row_column(c, M, ten, w);
      float v[PACK_ROWS];
#pragma unroll
      for (int r = 0; r < PACK_ROWS; r++) v[r] = (c < F && g[r] >= 0) ? __ldg(t.g[ten] + (size_t)g[r] * width[ten] + w) : 0.f;
#pragma unroll
      for (int r = 0; r < PACK_ROWS; r++)
        if (c < F && g[r] >= 0) table[(size_t)row[r] * W + 1 + c] = v[r];
    }
#pragma unroll
    for (int r = 0; r < PACK_ROWS; r++) {
      if (g[r] < 0) continue;
      float* out = table + (size_t)row[r] * W;
      if (lane == 0) out[0] = __int_as_float(g[r]);
      if (lane < 2) out[1 + F + lane] = g_means2D ? __ldg(g_means2D + 3 * (size_t)g[r] + lane) : 0.f;
      if (lane == 2) out[1 + F + 2] = (float)radii[g[r]];
    }
  }
}
__global__ void pack_visible_header_kernel(float* table, int capacity, int W, const unsigned int* count) {
  table[(size_t)capacity * W] = __int_as_float((int)*count);
}

// adds a received table: gradients (unless it is this rank's own table, whose rows are already in place) and the
// densification statistics; rows beyond the header's count or with ids outside [0, P) are ignored.
// The table may live in a PEER's memory (VisibleRowExchange with symmetric memory: gather and add are this one kernel,
// the loads go over NVLink): a warp therefore takes ADD_ROWS rows at a time and requests all their ids, then all their
// columns, before it touches the local rows — four times the bytes in flight per warp for a round trip of microseconds.
constexpr int ADD_ROWS = 4;
// The table may live in a peer's memory and was written a barrier (and a kernel boundary) ago; it is read-only for the
// lifetime of this kernel, so it goes through the read-only path: L1 is invalidated between kernels, and its 128-byte line
// fills are the request size NVLink moves best — ld.global.cg / ld.relaxed.sys (32-byte sectors) ran the peer pull 25 %
// slower (tests/tools/exchange_peer_probe.py, which changes the table contents every step and checks the sums).
__device__ __forceinline__ float ld_sys(const float* p) { return __ldg(p); }
__global__ void __launch_bounds__(256) add_counted_rows_kernel(const float* __restrict__ table, int capacity, int M, int P, GradRowTensors t,
                                                               int add_grads, float* __restrict__ max_radii2D,
                                                               float* __restrict__ xyz_gradient_accum, float* __restrict__ denom) {
  const int F = 11 + 3 * M, W = 1 + F + 3;
  const int n = min(__float_as_int(ld_sys(table + (size_t)capacity * W)), capacity);
  const int row0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * ADD_ROWS, lane = threadIdx.x & 31;
  if (row0 >= n) return;
  // ids, gradient columns and the three statistics columns are requested together (a column's address depends on the
  // row index only, not on the id): ONE round trip to the peer per group of rows
  int g[ADD_ROWS];
#pragma unroll
  for (int k = 0; k < ADD_ROWS; k++) g[k] = row0 + k < n ? __float_as_int(ld_sys(table + (size_t)(row0 + k) * W)) : -1;
  constexpr int MAX_PASSES = 2;                 // 1 + F + 3 = W <= 65 floats for SH degree <= 3 (M <= 16): columns 0 .. F + 2 in two passes
  const bool two_pass = F + 3 <= MAX_PASSES * 32;
  const int width[5] = {3, 3 * M, 1, 3, 4};
  float v[MAX_PASSES][ADD_ROWS];
#pragma unroll
  for (int ps = 0; ps < MAX_PASSES; ps++) {
    const int c = ps * 32 + lane;
    const bool want = two_pass ? (c < F + 3 && (add_grads || c >= F)) : (add_grads && c < F);
#pragma unroll
    for (int k = 0; k < ADD_ROWS; k++) v[ps][k] = (want && row0 + k < n) ? ld_sys(table + (size_t)(row0 + k) * W + 1 + c) : 0.f;
  }
#pragma unroll
  for (int k = 0; k < ADD_ROWS; k++)
    if (g[k] >= P) g[k] = -1;
  if (add_grads) {
#pragma unroll
    for (int ps = 0; ps < MAX_PASSES; ps++) {
      const int c = ps * 32 + lane;
      if (c < F) {
        int ten, w;
        row_column(c, M, ten, w);
        // Row ids are unique within one rank's table and the tables of a step are added one kernel after the other, so
        // every address receives ONE addition per launch: the reduction instruction gives the bits of `+=` (a subnormal
        // sum flushes to zero), but nothing comes back to the SM — no read-modify-write round trip per row.
#pragma unroll
        for (int k = 0; k < ADD_ROWS; k++)
          if (g[k] >= 0) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(t.g[ten] + (size_t)g[k] * width[ten] + w), "f"(v[ps][k]) : "memory");
      }
    }
    for (int c0 = MAX_PASSES * 32; c0 < F; c0 += 32) {      // rows wider than 64 columns (not reachable with SH degree <= 3)
      const int c = c0 + lane;
      if (c < F) {
        int ten, w;
        row_column(c, M, ten, w);
#pragma unroll
        for (int k = 0; k < ADD_ROWS; k++)
          if (g[k] >= 0) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(t.g[ten] + (size_t)g[k] * width[ten] + w), "f"(ld_sys(table + (size_t)(row0 + k) * W + 1 + c)) : "memory");
      }
    }
  }
  if (max_radii2D) {
    // densification statistics of row k on lane k: its (gx, gy, radius) sit in the lanes that loaded columns F .. F + 2
    float sx = 0.f, sy = 0.f, sr = 0.f;
    int gk = -1;
#pragma unroll
    for (int k = 0; k < ADD_ROWS; k++) {
      float x, y, r;
      if (two_pass) {
        const float x0 = __shfl_sync(0xffffffffu, v[0][k], F & 31), x1 = __shfl_sync(0xffffffffu, v[1][k], F & 31);
        const float y0 = __shfl_sync(0xffffffffu, v[0][k], (F + 1) & 31), y1 = __shfl_sync(0xffffffffu, v[1][k], (F + 1) & 31);
        const float r0 = __shfl_sync(0xffffffffu, v[0][k], (F + 2) & 31), r1 = __shfl_sync(0xffffffffu, v[1][k], (F + 2) & 31);
        x = F < 32 ? x0 : x1, y = F + 1 < 32 ? y0 : y1, r = F + 2 < 32 ? r0 : r1;
      } else {
        const float* in = table + (size_t)(row0 + k) * W;
        const bool have = row0 + k < n;
        x = have ? ld_sys(in + 1 + F) : 0.f, y = have ? ld_sys(in + 1 + F + 1) : 0.f, r = have ? ld_sys(in + 1 + F + 2) : 0.f;
      }
      if (lane == k) sx = x, sy = y, sr = r, gk = g[k];
    }
    if (lane < ADD_ROWS && gk >= 0) {
      max_radii2D[gk] = fmaxf(max_radii2D[gk], sr);
      xyz_gradient_accum[gk] += sqrtf(sx * sx + sy * sy);
      denom[gk] += 1.f;
    }
  }
}

void launch_pack_visible_rows(const int* radii, int P, int M, const GradRowTensors& t, const float* g_means2D, float* table, int capacity,
                              unsigned int* count, cudaStream_t stream) {
  cudaMemsetAsync(count, 0, sizeof(unsigned int), stream);
  if (P > 0) pack_visible_rows_kernel<<<(P + 255) / 256, 256, 0, stream>>>(radii, P, M, t, g_means2D, table, capacity, count);
  pack_visible_header_kernel<<<1, 1, 0, stream>>>(table, capacity, 1 + 11 + 3 * M + 3, count);
  count_launch(2);
}
void launch_add_counted_rows(const float* table, int capacity, int M, int P, const GradRowTensors& t, int add_grads, float* max_radii2D,
                             float* xyz_gradient_accum, float* denom, cudaStream_t stream) {
  if (capacity <= 0) return;
  add_counted_rows_kernel<<<(capacity + 8 * ADD_ROWS - 1) / (8 * ADD_ROWS), 256, 0, stream>>>(table, capacity, M, P, t, add_grads, max_radii2D, xyz_gradient_accum, denom);
  count_launch();
}

void launch_pack_gradient_rows(const long long* idx, int k, int K, int M, const GradRowTensors& t, float* table, cudaStream_t stream) {
  if (K <= 0) return;
  pack_gradient_rows_kernel<<<(K + 7) / 8, 256, 0, stream>>>(idx, k, K, M, t, table);
  count_launch();
}
void launch_add_gradient_rows(const float* table, int K, int M, const GradRowTensors& t, int P, cudaStream_t stream) {
  if (K <= 0) return;
  add_gradient_rows_kernel<<<(K + 7) / 8, 256, 0, stream>>>(table, K, M, t, P);
  count_launch();
}

}  // namespace gsr

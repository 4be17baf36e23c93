// distCUDA2 of simple-knn (gaussian_splatting/submodules/simple-knn/simple_knn.cu:185-221; spatial.cu:15-26):
// for every point the mean of the squared distances to its three nearest other points — the scale
// initialisation of GaussianModel.create_from_pcd (scene/gaussian_model.py:135-136).
//
// The reference searches exactly (Morton order, 1024-point boxes, one thread walking whole boxes serially), so its
// result is the multiset of the three smallest pair distances and does not depend on traversal order.  This
// implementation keeps the per-pair arithmetic of the reference build bit for bit (dy*dy rounded, then two FMAs;
// (b0+b1)+b2 divided by 3 with IEEE division) and replaces the search: 63-bit Morton sort with the library's own radix
// sort, 32-point boxes grouped 32 to a super-box, one warp per box of 32 queries, box pruning by warp votes against
// the warp's query bounds, candidates staged through a per-warp shared-memory slab and read as broadcast LDS.128.
#include <cfloat>

#include "gsr_kernels.cuh"

namespace gsr {

constexpr int KNN_BOX = 32;      // points per box (= one warp-wide load)
constexpr int KNN_SUPER = 32;    // boxes per super-box
constexpr int KNN_WARPS = 8;     // query warps per CTA

struct KnnWorkspace {
  uint32_t* bounds;   // [6] order-preserving uint images of min xyz, max xyz
  uint64_t* keys[2];
  uint32_t* vals[2];
  float4* sorted;     // [nb*32]  x y z idx-bits, padded with +inf
  float4* box_min;    // [nb] ; [nsb] super boxes follow
  float4* box_max;
  float4* sbox_min;
  float4* sbox_max;
  char* sort_temp;
};

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t carve_knn(char* base, long long P, KnnWorkspace& w) {
  const size_t nb = (size_t)((P + KNN_BOX - 1) / KNN_BOX), nsb = (nb + KNN_SUPER - 1) / KNN_SUPER;
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += align256(bytes); return p; };
  w.bounds = (uint32_t*)take(6 * sizeof(uint32_t));
  for (int i = 0; i < 2; i++) w.keys[i] = (uint64_t*)take(sizeof(uint64_t) * (size_t)P);
  for (int i = 0; i < 2; i++) w.vals[i] = (uint32_t*)take(sizeof(uint32_t) * (size_t)P);
  w.sorted = (float4*)take(sizeof(float4) * nb * KNN_BOX);
  w.box_min = (float4*)take(sizeof(float4) * nb);
  w.box_max = (float4*)take(sizeof(float4) * nb);
  w.sbox_min = (float4*)take(sizeof(float4) * nsb);
  w.sbox_max = (float4*)take(sizeof(float4) * nsb);
  w.sort_temp = take(sort_temp_bytes(P, 8));
  return off;
}

size_t knn_workspace_bytes(long long P) {
  KnnWorkspace w;
  return carve_knn(nullptr, P, w);
}

__device__ __forceinline__ uint32_t float_order(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float order_float(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

__global__ void knn_bounds_init_kernel(uint32_t* bounds) {
  if (threadIdx.x < 3) bounds[threadIdx.x] = 0xffffffffu;
  else if (threadIdx.x < 6) bounds[threadIdx.x] = 0u;
}

__global__ void __launch_bounds__(256) knn_bounds_kernel(const float* __restrict__ pts, int P, uint32_t* __restrict__ bounds) {
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x)
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const float v = __ldg(pts + 3 * (size_t)i + a);
      mn[a] = fminf(mn[a], v), mx[a] = fmaxf(mx[a], v);
    }
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
  if ((threadIdx.x & 31) == 0)
#pragma unroll
    for (int a = 0; a < 3; a++) {
      atomicMin(bounds + a, float_order(mn[a]));
      atomicMax(bounds + 3 + a, float_order(mx[a]));
    }
}

// 21 bits per axis (the reference's 10-bit codes, simple_knn.cu:44-61, put whole dense clusters of an SfM cloud into
// one cell and leave them unordered inside it; the order only affects speed, never the result)
__device__ __forceinline__ uint64_t spread21(uint64_t x) {
  x &= 0x1fffffull;
  x = (x | (x << 32)) & 0x1f00000000ffffull;
  x = (x | (x << 16)) & 0x1f0000ff0000ffull;
  x = (x | (x << 8)) & 0x100f00f00f00f00full;
  x = (x | (x << 4)) & 0x10c30c30c30c30c3ull;
  x = (x | (x << 2)) & 0x1249249249249249ull;
  return x;
}

__global__ void __launch_bounds__(256) knn_morton_kernel(const float* __restrict__ pts, int P, const uint32_t* __restrict__ bounds,
                                                         uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  uint64_t code = 0;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const float lo = order_float(bounds[a]), hi = order_float(bounds[3 + a]);
    const float ext = hi - lo;
    const float u = ext > 0.f ? (__ldg(pts + 3 * (size_t)i + a) - lo) / ext * 2097151.f : 0.f;
    code |= spread21((uint64_t)fminf(fmaxf(u, 0.f), 2097151.f)) << a;
  }
  keys[i] = code;
  vals[i] = (uint32_t)i;
}

// one warp per box: gather the box's points in Morton order, reduce their bounds
__global__ void __launch_bounds__(256) knn_boxes_kernel(const float* __restrict__ pts, int P, const uint32_t* __restrict__ order,
                                                        float4* __restrict__ sorted, float4* __restrict__ box_min,
                                                        float4* __restrict__ box_max, int nb) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= nb) return;
  const int s = b * KNN_BOX + lane;
  float4 p = make_float4(INFINITY, INFINITY, INFINITY, 0.f);
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  if (s < P) {
    const uint32_t id = order[s];
    p = make_float4(__ldg(pts + 3 * (size_t)id), __ldg(pts + 3 * (size_t)id + 1), __ldg(pts + 3 * (size_t)id + 2), __uint_as_float(id));
    mn[0] = mx[0] = p.x, mn[1] = mx[1] = p.y, mn[2] = mx[2] = p.z;
  }
  sorted[s] = p;
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
  if (lane == 0) box_min[b] = make_float4(mn[0], mn[1], mn[2], 0.f), box_max[b] = make_float4(mx[0], mx[1], mx[2], 0.f);
}

__global__ void __launch_bounds__(256) knn_super_boxes_kernel(const float4* __restrict__ box_min, const float4* __restrict__ box_max, int nb,
                                                              float4* __restrict__ sbox_min, float4* __restrict__ sbox_max, int nsb) {
  const int sb = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (sb >= nsb) return;
  const int b = sb * KNN_SUPER + lane;
  float4 mn = make_float4(FLT_MAX, FLT_MAX, FLT_MAX, 0.f), mx = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, 0.f);
  if (b < nb) mn = box_min[b], mx = box_max[b];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn.x = fminf(mn.x, __shfl_xor_sync(0xffffffffu, mn.x, o)), mn.y = fminf(mn.y, __shfl_xor_sync(0xffffffffu, mn.y, o));
    mn.z = fminf(mn.z, __shfl_xor_sync(0xffffffffu, mn.z, o)), mx.x = fmaxf(mx.x, __shfl_xor_sync(0xffffffffu, mx.x, o));
    mx.y = fmaxf(mx.y, __shfl_xor_sync(0xffffffffu, mx.y, o)), mx.z = fmaxf(mx.z, __shfl_xor_sync(0xffffffffu, mx.z, o));
  }
  if (lane == 0) sbox_min[sb] = mn, sbox_max[sb] = mx;
}

// squared distance in the reference build's operation order (simple_knn.cu:137-141 as compiled: FMUL on y, FFMA x, FFMA z)
__device__ __forceinline__ float dist2_ref(float dx, float dy, float dz) {
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}
// gap between two axis-aligned boxes (a point is a degenerate box); never above the distance of any pair they hold,
// because every step is monotone in the per-axis gaps
__device__ __forceinline__ float box_gap2(const float4& amin, const float4& amax, const float4& bmin, const float4& bmax) {
  const float dx = fmaxf(0.f, fmaxf(__fsub_rn(bmin.x, amax.x), __fsub_rn(amin.x, bmax.x)));
  const float dy = fmaxf(0.f, fmaxf(__fsub_rn(bmin.y, amax.y), __fsub_rn(amin.y, bmax.y)));
  const float dz = fmaxf(0.f, fmaxf(__fsub_rn(bmin.z, amax.z), __fsub_rn(amin.z, bmax.z)));
  return dist2_ref(dx, dy, dz);
}

__device__ __forceinline__ void keep3(float d, float& b0, float& b1, float& b2) {   // updateKBest<3>
  if (d < b2) {
    if (d < b1) {
      b2 = b1;
      if (d < b0) b1 = b0, b0 = d;
      else b1 = d;
    } else b2 = d;
  }
}

__global__ void __launch_bounds__(KNN_WARPS * 32) knn_query_kernel(const float4* __restrict__ sorted, int P, int nb, int nsb,
                                                                  const float4* __restrict__ box_min, const float4* __restrict__ box_max,
                                                                  const float4* __restrict__ sbox_min, const float4* __restrict__ sbox_max,
                                                                  float* __restrict__ mean_dists) {
  __shared__ float4 slab[KNN_WARPS][KNN_BOX];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int own = blockIdx.x * KNN_WARPS + warp;
  if (own >= nb) return;
  const int s = own * KNN_BOX + lane;
  const bool valid = s < P;
  const float4 q = sorted[s];
  float b0 = FLT_MAX, b1 = FLT_MAX, b2 = FLT_MAX;

  auto scan_box = [&](int b) {
    slab[warp][lane] = sorted[b * KNN_BOX + lane];
    __syncwarp();
#pragma unroll 8
    for (int j = 0; j < KNN_BOX; j++) {
      const float4 c = slab[warp][j];
      const float d = dist2_ref(__fsub_rn(c.x, q.x), __fsub_rn(c.y, q.y), __fsub_rn(c.z, q.z));
      if (!(b == own && j == lane)) keep3(d, b0, b1, b2);
    }
    __syncwarp();
  };
  scan_box(own);

  // the warp's query bounds: its own box
  const float4 wmin = box_min[own], wmax = box_max[own];
  auto warp_bound = [&]() {
    float m = valid ? b2 : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    return m;
  };
  float bound = warp_bound();

  auto visit_super_box = [&](int sb) {
    const int b_lane = sb * KNN_SUPER + lane;
    bool cand = false;
    float4 cmin = make_float4(0, 0, 0, 0), cmax = cmin;
    if (b_lane < nb && b_lane != own) {
      cmin = box_min[b_lane], cmax = box_max[b_lane];
      cand = !(box_gap2(wmin, wmax, cmin, cmax) > bound);
    }
    unsigned mask = __ballot_sync(0xffffffffu, cand);
    while (mask) {
      const int j = __ffs(mask) - 1;
      mask &= mask - 1;
      float4 bmn, bmx;
      bmn.x = __shfl_sync(0xffffffffu, cmin.x, j), bmn.y = __shfl_sync(0xffffffffu, cmin.y, j), bmn.z = __shfl_sync(0xffffffffu, cmin.z, j);
      bmx.x = __shfl_sync(0xffffffffu, cmax.x, j), bmx.y = __shfl_sync(0xffffffffu, cmax.y, j), bmx.z = __shfl_sync(0xffffffffu, cmax.z, j);
      const bool need = valid && !(box_gap2(q, q, bmn, bmx) > b2);
      if (!__any_sync(0xffffffffu, need)) continue;
      scan_box(sb * KNN_SUPER + j);
      bound = warp_bound();
    }
  };
  // our own super-box first (Morton neighbours tighten the bound fastest), then the rest 32 super-boxes per vote,
  // chunks visited outward from ours
  const int own_sb = own / KNN_SUPER;
  visit_super_box(own_sb);
  const int nchunks = (nsb + 31) >> 5, own_chunk = own_sb >> 5;
  for (int step = 0; step < 2 * nchunks; step++) {
    const int chunk = (step & 1) ? own_chunk + ((step + 1) >> 1) : own_chunk - (step >> 1);
    if (chunk < 0 || chunk >= nchunks) continue;
    const int sb_lane = (chunk << 5) + lane;
    const bool hit = sb_lane < nsb && sb_lane != own_sb && !(box_gap2(wmin, wmax, sbox_min[sb_lane], sbox_max[sb_lane]) > bound);
    unsigned smask = __ballot_sync(0xffffffffu, hit);
    while (smask) {
      const int k = __ffs(smask) - 1;
      smask &= smask - 1;
      const int sb = (chunk << 5) + k;
      if (box_gap2(wmin, wmax, sbox_min[sb], sbox_max[sb]) > bound) continue;   // the bound may have tightened since the vote
      visit_super_box(sb);
    }
  }
  if (valid) mean_dists[__float_as_uint(q.w)] = __fdiv_rn(__fadd_rn(__fadd_rn(b0, b1), b2), 3.0f);
}

int launch_dist2_knn3(const float* points, long long P, float* mean_dists, char* workspace, cudaStream_t stream) {
  KnnWorkspace w;
  carve_knn(workspace, P, w);
  const int n = (int)P;
  const int nb = (n + KNN_BOX - 1) / KNN_BOX, nsb = (nb + KNN_SUPER - 1) / KNN_SUPER;
  knn_bounds_init_kernel<<<1, 32, 0, stream>>>(w.bounds);
  knn_bounds_kernel<<<std::min((n + 255) / 256, sm_count() * 8), 256, 0, stream>>>(points, n, w.bounds);
  knn_morton_kernel<<<(n + 255) / 256, 256, 0, stream>>>(points, n, w.bounds, w.keys[0], w.vals[0]);
  SortTemp st;
  const int end_bit = 63, passes = sort_passes(end_bit);
  carve_sort_temp(w.sort_temp, P, passes, st);
  sort_temp_reset(w.sort_temp, P, passes, stream);
  launch_sort_histogram(w.keys[0], nullptr, P, 0, end_bit, st.hist, stream);
  launch_onesweep(w.keys, w.vals, nullptr, P, end_bit, st, stream);
  const uint32_t* order = w.vals[passes & 1];
  knn_boxes_kernel<<<(nb + 7) / 8, 256, 0, stream>>>(points, n, order, w.sorted, w.box_min, w.box_max, nb);
  knn_super_boxes_kernel<<<(nsb + 7) / 8, 256, 0, stream>>>(w.box_min, w.box_max, nb, w.sbox_min, w.sbox_max, nsb);
  knn_query_kernel<<<(nb + KNN_WARPS - 1) / KNN_WARPS, KNN_WARPS * 32, 0, stream>>>(w.sorted, n, nb, nsb, w.box_min, w.box_max, w.sbox_min,
                                                                                   w.sbox_max, mean_dists);
  count_launch(6);
  return 0;
}

}  // namespace gsr

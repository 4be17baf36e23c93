// Tile binning: per-tile list lengths and ranges from the coverage grid, bucketing of the
// (Gaussian, tile) instances, a tile-local sort — and the hand-written onesweep LSD radix sort.
//
// Replaces, from the reference (gaussian_splatting/submodules/diff-gaussian-rasterization):
//   cub::DeviceScan::InclusiveSum             cuda_rasterizer/rasterizer_impl.cu:278
//   duplicateWithKeys                         cuda_rasterizer/rasterizer_impl.cu:70-111
//   cub::DeviceRadixSort::SortPairs           cuda_rasterizer/rasterizer_impl.cu:304-309 (size query :187-190)
//   cudaMemset(ranges) + identifyTileRanges   cuda_rasterizer/rasterizer_impl.cu:311-318, 116-138
//
// B200 design.  The reference sorts all R instances globally on a 64-bit tile|depth key (six
// 8-bit passes, each reading and writing 12R bytes) and then recovers the tile ranges from the
// sorted keys.  The tile half of that key is known before sorting, so here:
//   1. scan_tiles: a 2-D prefix sum over the tile-coverage difference grid (written by
//      preprocess, four atomics per Gaussian) gives every tile's list length; their exclusive
//      scan IS the tile ranges and num_rendered.  One small CTA.
//   2. scatter: CTAs expand their slot segment cooperatively (output slot j finds its owner by
//      binary search over CTA-local prefix sums, so a Gaussian covering thousands of tiles
//      costs no more per instance than one covering four) and drop each instance into its
//      tile's bucket through an atomic cursor, as the composite depth_bits << 32 | slot.
//   3. tile_sort: one CTA per tile sorts its bucket with a register / shuffle bitonic network.
//      Slot order is Gaussian order, so ascending (depth bits, slot) is exactly the order the
//      reference's stable LSD sort produces — ties included — while each instance crosses HBM
//      once instead of 12 times.  The network runs on 32-bit words (quantised depth | index in
//      the bucket: a third of the 64-bit network's instructions); runs of equal quantised depth
//      are resolved exactly afterwards (tile_sort_small).  Buckets larger than the shared-memory
//      budget are sorted by the generic 64-bit network directly in global memory (L2).
// The onesweep radix sort (Adinets & Merrill) that replaces cub::DeviceRadixSort as a library
// primitive is kept below and exported as gsr_sort_pairs.
#include <cstdlib>

#include "gsr_kernels.cuh"

namespace gsr {

// ------------------------------------------------------------------ sort temp layout
size_t sort_temp_bytes(long long n, int passes) {
  size_t b = 0;
  b += align_up(sizeof(uint32_t) * SORT_MAX_PASSES * SORT_RADIX, 128);
  b += align_up(sizeof(uint32_t) * (size_t)passes * sort_num_tiles(n) * SORT_RADIX, 128);
  return b + 128;
}
void carve_sort_temp(char* base, long long n, int passes, SortTemp& t) {
  char* p = base;
  carve(p, t.hist, (size_t)SORT_MAX_PASSES * SORT_RADIX);
  carve(p, t.status, (size_t)passes * sort_num_tiles(n) * SORT_RADIX);
}
void sort_temp_reset(char* base, long long n, int passes, cudaStream_t stream) {
  SortTemp t;
  carve_sort_temp(base, n, passes, t);
  char* end = (char*)(t.status + (size_t)passes * sort_num_tiles(n) * SORT_RADIX);
  cudaMemsetAsync(t.hist, 0, (size_t)(end - (char*)t.hist), stream);
}

__device__ __forceinline__ uint32_t digit_of(uint64_t key, int shift, uint32_t mask) {
  return (uint32_t)(key >> shift) & mask;
}

// ------------------------------------------------------------------ block-wide bitonic network
// Ascending bitonic network in which every merge starts with a mirrored compare (i <-> i ^ (k-1))
// followed by half-cleaners, so ALL compare-exchanges put the smaller key at the lower index and
// virtual +inf padding beyond `n` never has to be stored or moved.
template <typename T, typename Ptr>
__device__ __forceinline__ void bitonic_sort_block(Ptr a, uint32_t n, uint32_t n2) {
  const uint32_t half = n2 >> 1;
  for (uint32_t k = 2; k <= n2; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      const bool mirror = j == (k >> 1);
      for (uint32_t t = threadIdx.x; t < half; t += blockDim.x) {
        const uint32_t i0 = ((t & ~(j - 1)) << 1) | (t & (j - 1));          // t with a zero inserted at bit j
        const uint32_t i1 = mirror ? (i0 & ~(k - 1)) | ((k - 1) - (i0 & (k - 1))) : (i0 | j);
        if (i1 < n) {                                                        // beyond n: virtual +inf, already in place
          const T x = a[i0], y = a[i1];
          if (x > y) { a[i0] = y; a[i1] = x; }
        }
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------ tile counts -> ranges
// One CTA.  diff is the (gy+1) x (gx+1) difference grid; after the 2-D inclusive prefix sum
// diff[y][x] is the number of Gaussians whose rectangle covers tile (x, y).
constexpr uint32_t ST_SMEM_CELLS = 10240;   // difference grids up to this many cells are scanned in shared memory
constexpr int ST_THREADS = 1024;
constexpr int ST_BUCKETS = 64;

// One CTA.  IN_SMEM: the whole grid fits in shared memory (any image up to ~2.5 Mpixel).
template <bool IN_SMEM>
__global__ void __launch_bounds__(ST_THREADS) scan_tiles_kernel(int* __restrict__ diff_g, uint32_t gx, uint32_t gy,
                                                                uint2* __restrict__ ranges, uint32_t* __restrict__ cursor,
                                                                uint32_t* __restrict__ tile_order,
                                                                uint32_t* __restrict__ counters, uint32_t capacity) {
  __shared__ int s_grid[IN_SMEM ? ST_SMEM_CELLS : 1];
  __shared__ uint32_t s_part[ST_THREADS];
  pdl_trigger();
  pdl_wait();
  __shared__ uint32_t s_max[32];
  __shared__ uint32_t s_bucket[ST_BUCKETS];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr uint32_t nt = ST_THREADS;
  const uint32_t stride = gx + 1;
  const uint32_t cells = stride * (gy + 1);
  // logical cell i lives at diff_g[i * DIFF_STRIDE]; the shared copy is packed
  int* diff = IN_SMEM ? s_grid : diff_g;
  constexpr uint32_t DS = IN_SMEM ? 1u : (uint32_t)DIFF_STRIDE;
  if (IN_SMEM) {
    for (uint32_t i = tid; i < cells; i += nt) s_grid[i] = diff_g[(size_t)i * DIFF_STRIDE];
  }
  if (tid < ST_BUCKETS) s_bucket[tid] = 0;
  __syncthreads();
  // prefix along rows: one warp per row, shuffle scan over 32-wide pieces
  for (uint32_t y = warp; y <= gy; y += nt / 32) {
    int carry = 0;
    int* row = diff + (size_t)y * stride * DS;
    for (uint32_t x0 = 0; x0 <= gx; x0 += 32) {
      const uint32_t x = x0 + lane;
      int v = x <= gx ? row[x * DS] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int n = __shfl_up_sync(0xffffffffu, v, o); if (lane >= (uint32_t)o) v += n; }
      v += carry;
      if (x <= gx) row[x * DS] = v;
      carry = __shfl_sync(0xffffffffu, v, 31);
    }
  }
  __syncthreads();
  // prefix along columns: one thread per column
  for (uint32_t x = tid; x <= gx; x += nt) {
    int run = 0;
    for (uint32_t y = 0; y <= gy; y++) { run += diff[(size_t)(y * stride + x) * DS]; diff[(size_t)(y * stride + x) * DS] = run; }
  }
  __syncthreads();
  // exclusive scan of the T counts in tile order: each thread owns a contiguous chunk
  const uint32_t T = gx * gy;
  const uint32_t chunk = (T + nt - 1) / nt;
  const uint32_t t0 = min(T, tid * chunk), t1 = min(T, t0 + chunk);
  uint32_t sum = 0, mx = 0;
  {
    uint32_t ty = t0 / gx, tx = t0 - ty * gx;
    for (uint32_t t = t0; t < t1; t++) {
      const uint32_t c = (uint32_t)diff[(size_t)(ty * stride + tx) * DS];
      sum += c;
      mx = max(mx, c);
      if (++tx == gx) tx = 0, ty++;
    }
  }
  // block scan of the per-thread sums
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += n; }
  mx = __reduce_max_sync(0xffffffffu, mx);
  if (lane == 31) s_part[warp] = incl;
  if (lane == 0) s_max[warp] = mx;
  __syncthreads();
  if (warp == 0) {
    const uint32_t w = s_part[lane];
    uint32_t wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= (uint32_t)o) wi += n; }
    s_part[lane] = wi - w;                       // exclusive offset of each warp
    const uint32_t m = __reduce_max_sync(0xffffffffu, s_max[lane]);
    if (lane == 31) {
      counters[1] = wi;                          // num_rendered
      if (capacity && wi > capacity) counters[3] = 1, counters[16] = 1;   // sync-free forward: the caller's binning buffer is too small (16: sticky copy)
    }
    if (lane == 0) counters[4] = m, s_max[0] = m;  // longest tile list
  }
  __syncthreads();
  const uint32_t longest = s_max[0];
  const float bucket_scale = (float)ST_BUCKETS / (float)(longest + 1u);
  // ranges + bucket cursors + launch order of the blend CTAs (longest lists first: counting sort on 64 length classes)
  uint32_t run = s_part[warp] + incl - sum;
  uint32_t my_bucket[(8192 + ST_THREADS - 1) / ST_THREADS + 1];
  const bool order_ok = chunk <= sizeof(my_bucket) / sizeof(uint32_t);
  {
    uint32_t ty = t0 / gx, tx = t0 - ty * gx;
    for (uint32_t t = t0, i = 0; t < t1; t++, i++) {
      const uint32_t c = (uint32_t)diff[(size_t)(ty * stride + tx) * DS];
      ranges[t] = c ? make_uint2(run, run + c) : make_uint2(0u, 0u);   // untouched tiles stay (0,0) like the reference
      cursor[(size_t)t * CURSOR_STRIDE] = run;
      run += c;
      if (order_ok) {
        const uint32_t bk = (ST_BUCKETS - 1) - min((uint32_t)(ST_BUCKETS - 1), (uint32_t)((float)c * bucket_scale));   // a launch-order heuristic: float is exact enough
        my_bucket[i] = bk;
        atomicAdd(&s_bucket[bk], 1u);
      }
      if (++tx == gx) tx = 0, ty++;
    }
  }
  __syncthreads();
  if (!order_ok) {
    for (uint32_t t = tid; t < T; t += nt) tile_order[t] = t;
    return;
  }
  if (warp == 0) {   // exclusive scan of the 64 class sizes
    const uint32_t a = s_bucket[lane], b = s_bucket[lane + 32];
    uint32_t ia = a, ib = b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t na = __shfl_up_sync(0xffffffffu, ia, o), nb = __shfl_up_sync(0xffffffffu, ib, o);
      if (lane >= (uint32_t)o) ia += na, ib += nb;
    }
    const uint32_t tot_a = __shfl_sync(0xffffffffu, ia, 31);
    s_bucket[lane] = ia - a;
    s_bucket[lane + 32] = tot_a + ib - b;
  }
  __syncthreads();
  for (uint32_t t = t0, i = 0; t < t1; t++, i++) tile_order[atomicAdd(&s_bucket[my_bucket[i]], 1u)] = t;
}

void launch_scan_tiles(int* tile_diff, uint32_t gx, uint32_t gy, uint2* ranges, uint32_t* cursor, uint32_t* tile_order,
                       uint32_t* counters, uint32_t capacity, cudaStream_t stream) {
  if ((gx + 1) * (gy + 1) <= ST_SMEM_CELLS)
    launch_pdl(scan_tiles_kernel<true>, dim3(1), dim3(ST_THREADS), 0, stream, tile_diff, gx, gy, ranges, cursor, tile_order, counters, capacity);
  else
    launch_pdl(scan_tiles_kernel<false>, dim3(1), dim3(ST_THREADS), 0, stream, tile_diff, gx, gy, ranges, cursor, tile_order, counters, capacity);
  count_launch();
}

// cursors back to the start of each tile's range (re-run of the tail after a speculative launch overflowed)
__global__ void reset_cursors_kernel(int T, const uint2* __restrict__ ranges, const int* __restrict__ diff, uint32_t gx,
                                     uint32_t* __restrict__ cursor) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  // empty tiles have range (0,0) but their cursor was the running prefix; they receive no instance, any value works
  cursor[(size_t)t * CURSOR_STRIDE] = ranges[t].x;
  (void)diff, (void)gx;
}
void launch_reset_cursors(int T, const uint2* ranges, uint32_t* cursor, cudaStream_t stream) {
  if (T <= 0) return;
  reset_cursors_kernel<<<(T + 255) / 256, 256, 0, stream>>>(T, ranges, nullptr, 0, cursor);
  count_launch();
}

// ------------------------------------------------------------------ instances -> tile buckets
// One CTA per preprocess slot segment: atomic cursor per tile, composite depth|slot.  Also zeroes the backward
// accumulator rows of its slots.
__global__ void __launch_bounds__(256) scatter_kernel(const uint32_t* __restrict__ block_vis, const uint2* __restrict__ rects,
                                                      const float* __restrict__ depths, uint32_t* __restrict__ cursor,
                                                      uint64_t* __restrict__ comp, float* __restrict__ grad_acc,
                                                      uint32_t grid_x, BinHeader hv, BinHeader* header, uint32_t* unit_count) {
  const uint32_t capacity = (uint32_t)hv.capacity;
  __shared__ uint32_t s_end[256];     // CTA-local inclusive prefix of tile counts
  __shared__ uint2 s_rect[256];
  __shared__ uint32_t s_depth[256];
  __shared__ uint32_t s_wsum[8];
  pdl_trigger();
  pdl_wait();
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (header) *header = hv;
    unit_count[0] = 0;   // the forward blend appends the backward's work units ...
    unit_count[3] = unit_count[4] = unit_count[5] = unit_count[6] = 0;   // ... in four cost classes
  }
  const uint32_t cnt = block_vis[blockIdx.x];
  if (cnt == 0) return;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t first = blockIdx.x * 256u;
  uint32_t tiles = 0;
  if (tid < cnt) {
    const uint2 rc = __ldg(rects + first + tid);
    s_rect[tid] = rc;
    s_depth[tid] = __float_as_uint(__ldg(depths + first + tid));
    tiles = ((rc.x >> 16) - (rc.x & 0xffffu)) * ((rc.y >> 16) - (rc.y & 0xffffu));
    float4* acc = reinterpret_cast<float4*>(grad_acc + 12 * (size_t)(first + tid));
    acc[0] = acc[1] = acc[2] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  uint32_t incl = tiles;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += v; }
  if (lane == 31) s_wsum[warp] = incl;
  __syncthreads();
  uint32_t woff = 0;
  for (uint32_t w = 0; w < warp; w++) woff += s_wsum[w];
  s_end[tid] = woff + incl;
  __syncthreads();
  const uint32_t total = s_end[255];
  for (uint32_t j = tid; j < total; j += 256u) {
    uint32_t lo = 0, hi = cnt - 1;      // smallest t with s_end[t] > j
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (s_end[mid] > j) hi = mid; else lo = mid + 1;
    }
    const uint32_t within = j - (lo == 0 ? 0u : s_end[lo - 1]);
    const uint2 rc = s_rect[lo];
    const uint32_t minx = rc.x & 0xffffu, w = (rc.x >> 16) - minx, miny = rc.y & 0xffffu;
    const uint32_t row = within / w, col = within - row * w;
    const uint32_t tile = (miny + row) * grid_x + (minx + col);
    const uint32_t pos = atomicAdd(cursor + (size_t)tile * CURSOR_STRIDE, 1u);
    if (pos < capacity) comp[pos] = ((uint64_t)s_depth[lo] << 32) | (first + lo);
  }
}

// exclusive scan, in place, of the per-CTA instance counts (one CTA; n = ceil(P/256) values)
__global__ void __launch_bounds__(1024) scan_blocks_kernel(const uint32_t* __restrict__ a, uint32_t* __restrict__ out, uint32_t n,
                                                           uint32_t* __restrict__ total_out) {
  __shared__ uint32_t s_warp[32];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t chunk = (n + 1023) / 1024;
  const uint32_t i0 = min(n, tid * chunk), i1 = min(n, i0 + chunk);
  uint32_t sum = 0;
  for (uint32_t i = i0; i < i1; i++) sum += a[i];
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += v; }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const uint32_t w = s_warp[lane];
    uint32_t wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= (uint32_t)o) wi += v; }
    s_warp[lane] = wi - w;
  }
  __syncthreads();
  uint32_t run = s_warp[warp] + incl - sum;
  for (uint32_t i = i0; i < i1; i++) { const uint32_t c = a[i]; out[i] = run; run += c; }
  if (tid == 1023) {           // the last thread's running sum is the total
    out[n] = run;
    if (total_out) *total_out = run;
  }
}

void launch_scatter(int P, const GeometryView& g, uint32_t* cursor, uint64_t* comp, uint32_t grid_x, BinHeader hv,
                    BinHeader* header, cudaStream_t stream) {
  if (P <= 0) return;
  launch_pdl(scatter_kernel, dim3(num_pre_blocks(P)), dim3(256), 0, stream, (const uint32_t*)g.block_vis, (const uint2*)g.rect,
             (const float*)g.depths, cursor, comp, g.grad_acc, grid_x, hv, header, g.counters + 5);
  count_launch();
}

// ------------------------------------------------------------------ tile-local sort (kernel)
constexpr int TS_THREADS = 256;
constexpr uint32_t TS_SMEM_KEYS = 4096;   // 32 KB of composite keys in shared memory
constexpr int TS_RUN_SCAN = 16;           // longest run of equal quantised depths that is ranked in place
#ifndef GSR_TS_CTAS_PER_SM
#define GSR_TS_CTAS_PER_SM 5
#endif
constexpr int TS_CTAS_PER_SM = GSR_TS_CTAS_PER_SM;

__device__ __forceinline__ uint64_t shfl_xor_key(uint64_t v, int mask) {
  const uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, mask);
  const uint32_t hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), mask);
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint32_t shfl_xor_key(uint32_t v, int mask) { return __shfl_xor_sync(0xffffffffu, v, mask); }

// Register-resident bitonic sort of N2 = 256*E keys held E per thread (thread t owns indices [tE, tE+E)).
// Same network as bitonic_sort_block (mirror step, then half-cleaners, always ascending), but a
// compare-exchange whose partner lives in the same thread is a register swap, one whose partner lives in
// the same warp is a shuffle, and only the few steps that cross warps go through shared memory.
// For N2 = 1024 that is 6 shared-memory steps out of 55.  Padding is real +inf keys in registers.
// The k / j loops stay rolled (the fully unrolled network is ~30k instructions and thrashes the
// instruction cache); only the per-thread element loops are unrolled.
template <typename K, int E, int J>
__device__ __forceinline__ void cx_in_thread(K (&v)[E]) {   // partner r ^ J, J < E
#pragma unroll
  for (int r = 0; r < E; r++) {
    if ((r & J) == 0 && (r | J) < E) {
      const K a = v[r], b = v[r | J];
      v[r] = min(a, b);
      v[r | J] = max(a, b);
    }
  }
}
template <typename K, int E, int KK>
__device__ __forceinline__ void mirror_in_thread(K (&v)[E]) {   // partner r ^ (KK-1), KK <= E
#pragma unroll
  for (int r = 0; r < E; r++) {
    const int rp = r ^ (KK - 1);
    if (r < rp && rp < E) {
      const K a = v[r], b = v[rp];
      v[r] = min(a, b);
      v[rp] = max(a, b);
    }
  }
}

template <typename K, int E>
__device__ __forceinline__ void bitonic_sort_regs(K (&v)[E], K* s) {
  constexpr int N2 = TS_THREADS * E;
  const uint32_t t = threadIdx.x;
  for (int k = 2; k <= N2; k <<= 1) {
    // ---- mirror step: i <-> i ^ (k-1)
    if (k <= E) {
      switch (k) {
        case 2: mirror_in_thread<K, E, 2>(v); break;
        case 4: mirror_in_thread<K, E, 4>(v); break;
        case 8: mirror_in_thread<K, E, 8>(v); break;
        default: mirror_in_thread<K, E, 16>(v); break;
      }
    } else if (k <= 32 * E) {
      const int m = k / E - 1;
      const bool lower = (t & (k / (2 * E))) == 0;
      K w[E];
#pragma unroll
      for (int r = 0; r < E; r++) w[r] = v[r];
#pragma unroll
      for (int r = 0; r < E; r++) {
        const K pv = shfl_xor_key(w[E - 1 - r], m);
        v[r] = lower ? min(w[r], pv) : max(w[r], pv);
      }
    } else {
      __syncthreads();
#pragma unroll
      for (int r = 0; r < E; r++) s[t * E + r] = v[r];
      __syncthreads();
#pragma unroll
      for (int r = 0; r < E; r++) {
        const uint32_t i = t * E + r, ip = i ^ (uint32_t)(k - 1);
        const K pv = s[ip];
        v[r] = (i < ip) ? min(v[r], pv) : max(v[r], pv);
      }
    }
    // ---- half-cleaners: i <-> i ^ j
    for (int j = k >> 2; j > 0; j >>= 1) {
      if (j < E) {
        switch (j) {
          case 1: cx_in_thread<K, E, 1>(v); break;
          case 2: cx_in_thread<K, E, 2>(v); break;
          case 4: cx_in_thread<K, E, 4>(v); break;
          default: cx_in_thread<K, E, 8>(v); break;
        }
      } else if (j < 32 * E) {
        const int m = j / E;
        const bool lower = (t & m) == 0;
#pragma unroll
        for (int r = 0; r < E; r++) {
          const K pv = shfl_xor_key(v[r], m);
          v[r] = lower ? min(v[r], pv) : max(v[r], pv);
        }
      } else {
        __syncthreads();
#pragma unroll
        for (int r = 0; r < E; r++) s[t * E + r] = v[r];
        __syncthreads();
#pragma unroll
        for (int r = 0; r < E; r++) {
          const uint32_t i = t * E + r, ip = i ^ (uint32_t)j;
          const K pv = s[ip];
          v[r] = (i < ip) ? min(v[r], pv) : max(v[r], pv);
        }
      }
    }
  }
}

// Sort a bucket of n <= 256 E composites (depth bits << 32 | slot) and write the slots in that order.
// The network runs on 32-bit words, a third of the 64-bit network's instructions (one shuffle and one VIMNMX per
// compare-exchange instead of two shuffles, two compares and two selects): word = quantised depth << b | index in the
// bucket, b = log2(256 E).  The quantisation is a shift of (depth bits - smallest depth bits of the bucket) chosen so that
// the bucket's depth range fills the 31 - b bits that are left — an integer, monotone map of the composite's upper half,
// so the words order the bucket exactly except inside runs of equal quantised depth (a handful of pairs per frame).  Those
// are found by comparing neighbours after the sort and resolved exactly: every member of a run counts the members whose
// full composite is smaller than its own and takes that place in the run.  The result is the ascending order of the
// 64-bit composites — the reference's stable (depth, Gaussian id) order — for any input; a bucket with a run longer
// than TS_RUN_SCAN (many equal depths: nothing a camera produces) is handed to the generic 64-bit network in place
// (register-light, so the rare path does not set the kernel's register count).
template <int E>
__device__ __forceinline__ void tile_sort_small(uint64_t* __restrict__ comp, uint32_t* __restrict__ point_list,
                                                uint32_t start, uint32_t n, uint64_t* s64) {
  constexpr uint32_t N2 = TS_THREADS * E;
  constexpr int B = (E == 1 ? 8 : E == 2 ? 9 : E == 4 ? 10 : E == 8 ? 11 : 12);
  constexpr uint32_t IDX_MASK = N2 - 1u;
  uint32_t* s32 = reinterpret_cast<uint32_t*>(s64);
  __shared__ uint32_t s_red[2][TS_THREADS / 32];
  const uint32_t t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const uint32_t base = t * E;
  uint32_t d[E];
  uint32_t dmin = 0xffffffffu, dmax = 0u;
#pragma unroll
  for (int r = 0; r < E; r++) {
    const bool ok = base + r < n;
    d[r] = ok ? (uint32_t)(__ldg(comp + start + base + r) >> 32) : 0u;
    if (ok) dmin = min(dmin, d[r]), dmax = max(dmax, d[r]);
  }
  dmin = __reduce_min_sync(0xffffffffu, dmin), dmax = __reduce_max_sync(0xffffffffu, dmax);
  if (lane == 0) s_red[0][warp] = dmin, s_red[1][warp] = dmax;
  __syncthreads();
#pragma unroll
  for (int w = 0; w < TS_THREADS / 32; w++) dmin = min(dmin, s_red[0][w]), dmax = max(dmax, s_red[1][w]);
  // (dmax - dmin) >> sh < 2^(31 - B): the top bit stays clear, so the all-ones padding word sorts behind every key
  const int sh = max(0, (32 - __clz(dmax - dmin)) - (31 - B));
  uint32_t v[E];
#pragma unroll
  for (int r = 0; r < E; r++) v[r] = (base + r < n) ? (((d[r] - dmin) >> sh) << B) | (base + r) : 0xffffffffu;
  bitonic_sort_regs<uint32_t, E>(v, s32);
  // publish the sorted words; pair (p, p+1) with equal quantised depth = a run
  __syncthreads();
#pragma unroll
  for (int r = 0; r < E; r++) s32[base + r] = v[r];
  __syncthreads();
  bool run = false;
#pragma unroll
  for (int r = 0; r < E; r++) {
    const uint32_t p = base + r;
    if (p + 1 < n) {
      const uint32_t nx = (r + 1 < E) ? v[(r + 1) % E] : s32[p + 1];
      run |= (nx >> B) == (v[r] >> B);
    }
  }
  if (!__syncthreads_or(run)) {
    // the common case: the words are distinct in their depth part, position p takes the slot of bucket entry (word & mask)
#pragma unroll
    for (int r = 0; r < E; r++)
      if (base + r < n) point_list[start + base + r] = __ldg(reinterpret_cast<const uint32_t*>(comp + start + (v[r] & IDX_MASK)));
    return;
  }
  bool too_long = false;
  uint32_t dst[E], slot[E];
#pragma unroll
  for (int r = 0; r < E; r++) {
    const uint32_t p = base + r;
    dst[r] = p, slot[r] = 0;
    if (p < n) {
      const uint32_t q = v[r] >> B;
      const uint64_t f = __ldg(comp + start + (v[r] & IDX_MASK));
      slot[r] = (uint32_t)f;
      int steps = 0;
      for (uint32_t j = p; j > 0 && (s32[j - 1] >> B) == q; j--) {        // members in front of p that belong behind it
        if (++steps > TS_RUN_SCAN) { too_long = true; break; }
        if (__ldg(comp + start + (s32[j - 1] & IDX_MASK)) > f) dst[r]--;
      }
      steps = 0;
      for (uint32_t j = p + 1; j < n && (s32[j] >> B) == q; j++) {        // members behind p that belong in front of it
        if (++steps > TS_RUN_SCAN) { too_long = true; break; }
        if (__ldg(comp + start + (s32[j] & IDX_MASK)) < f) dst[r]++;
      }
    }
  }
  if (__syncthreads_or(too_long)) {
    uint64_t* a = comp + start;
    bitonic_sort_block<uint64_t>(a, n, N2);
    for (uint32_t i = t; i < n; i += TS_THREADS) point_list[start + i] = (uint32_t)a[i];
    return;
  }
#pragma unroll
  for (int r = 0; r < E; r++)
    if (base + r < n) point_list[start + dst[r]] = slot[r];
}

__global__ void __launch_bounds__(TS_THREADS, TS_CTAS_PER_SM) tile_sort_kernel(const uint2* __restrict__ ranges, uint64_t* __restrict__ comp,
                                                               uint32_t* __restrict__ point_list, uint32_t capacity,
                                                               const uint32_t* __restrict__ tile_order) {
  __shared__ uint64_t s_keys[TS_SMEM_KEYS];
  pdl_trigger();
  pdl_wait();
  // longest lists first (the blend kernels' launch order): a 4096-key network takes many times longer than the median
  // tile's, and started late it finishes the kernel alone
  const uint2 rg = ranges[tile_order ? tile_order[blockIdx.x] : blockIdx.x];
  const uint32_t start = min(rg.x, capacity), end = min(rg.y, capacity);
  if (end <= start) return;
  const uint32_t n = end - start;
  if (n == 1) {
    if (threadIdx.x == 0) point_list[start] = (uint32_t)comp[start];
  } else if (n <= 1 * TS_THREADS) {
    tile_sort_small<1>(comp, point_list, start, n, s_keys);
  } else if (n <= 2 * TS_THREADS) {
    tile_sort_small<2>(comp, point_list, start, n, s_keys);
  } else if (n <= 4 * TS_THREADS) {
    tile_sort_small<4>(comp, point_list, start, n, s_keys);
  } else if (n <= 8 * TS_THREADS) {
    tile_sort_small<8>(comp, point_list, start, n, s_keys);
  } else if (n <= 16 * TS_THREADS) {
    tile_sort_small<16>(comp, point_list, start, n, s_keys);
  } else {
    uint32_t n2 = 1;
    while (n2 < n) n2 <<= 1;
    uint64_t* a = comp + start;      // large bucket: the generic network straight on global memory (L2 resident)
    bitonic_sort_block<uint64_t>(a, n, n2);
    for (uint32_t i = threadIdx.x; i < n; i += TS_THREADS) point_list[start + i] = (uint32_t)a[i];
  }
}

void launch_tile_sort(int num_tiles, const uint2* ranges, uint64_t* comp, uint32_t* point_list, uint32_t capacity,
                      const uint32_t* tile_order, cudaStream_t stream) {
  if (num_tiles <= 0) return;
  launch_pdl(tile_sort_kernel, dim3(num_tiles), dim3(TS_THREADS), 0, stream, ranges, comp, point_list, capacity, tile_order);
  count_launch();
}

// ------------------------------------------------------------------ digit histograms of all passes
// Keys of one Gaussian are adjacent and share every depth digit, so equal digits are first
// aggregated inside the warp (match_any) and cost one shared-memory atomic per group.
__global__ void __launch_bounds__(256) sort_histogram_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ n_ptr,
                                                             uint32_t n_host, int lo_bit, int end_bit, uint32_t* __restrict__ hist) {
  __shared__ uint32_t s_hist[SORT_MAX_PASSES * SORT_RADIX];
  const uint32_t n = n_ptr ? min(*n_ptr, n_host) : n_host;
  const int passes = (end_bit - lo_bit + SORT_RADIX_BITS - 1) / SORT_RADIX_BITS;
  for (int i = threadIdx.x; i < passes * SORT_RADIX; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t stride = gridDim.x * blockDim.x;
  // warp-uniform trip count so that match_any sees full warps
  for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += stride) {
    const uint32_t i = base + lane;
    const bool valid = i < n;
    const uint64_t k = valid ? __ldg(keys + i) : 0ull;
    const uint32_t act = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      for (int p = 0; p < passes; p++) {
        const int shift = lo_bit + SORT_RADIX_BITS * p;
        const uint32_t mask = (1u << min(SORT_RADIX_BITS, end_bit - shift)) - 1u;
        const uint32_t d = digit_of(k, shift, mask);
        const uint32_t peers = __match_any_sync(act, d);
        if (lane == (uint32_t)(__ffs(peers) - 1)) atomicAdd(&s_hist[p * SORT_RADIX + d], (uint32_t)__popc(peers));
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < passes * SORT_RADIX; i += blockDim.x) {
    const uint32_t c = s_hist[i];
    if (c) atomicAdd(&hist[i], c);
  }
}
void launch_sort_histogram(const uint64_t* keys, const uint32_t* n_ptr, long long n_host, int lo_bit, int end_bit, uint32_t* hist,
                           cudaStream_t stream) {
  if (n_host <= 0) return;
  const int blocks = (int)std::min<long long>((n_host + 256 * 8 - 1) / (256 * 8), sm_count() * 8);
  sort_histogram_kernel<<<blocks, 256, 0, stream>>>(keys, n_ptr, (uint32_t)n_host, lo_bit, end_bit, hist);
  count_launch();
}

// counts -> exclusive bases, one CTA of 256 threads per pass
__global__ void __launch_bounds__(SORT_RADIX) sort_scan_hist_kernel(uint32_t* __restrict__ hist) {
  __shared__ uint32_t s_warp[SORT_RADIX / 32];
  uint32_t* h = hist + blockIdx.x * SORT_RADIX;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t c = h[threadIdx.x];
  uint32_t incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  uint32_t base = 0;
  for (uint32_t w = 0; w < warp; w++) base += s_warp[w];
  h[threadIdx.x] = base + incl - c;
}

// ------------------------------------------------------------------ onesweep pass
constexpr uint32_t FLAG_AGG = 1u << 30, FLAG_INCL = 2u << 30, VALUE_MASK = (1u << 30) - 1u;

template <bool PAIRS>
struct __align__(16) SortSmemT {
  uint64_t keys[SORT_TILE];
  uint32_t vals[PAIRS ? SORT_TILE : 1];
  uint32_t warp_hist[SORT_THREADS / 32][SORT_RADIX];
  uint32_t local_start[SORT_RADIX];
  uint32_t adj[SORT_RADIX];
  uint32_t warp_tot[SORT_RADIX / 32];
};
using SortSmem = SortSmemT<true>;

// MODE 0: (key, value) pairs.  MODE 1: keys only (the value rides in the key's low bits).  MODE 2: keys only, last pass:
// the low `out_bits` of every key are written, as a u32, to vout instead of the key to kout.
template <int MODE>
__global__ void __launch_bounds__(SORT_THREADS) onesweep_pass_kernel(const uint64_t* __restrict__ kin, uint64_t* __restrict__ kout,
                                                                    const uint32_t* __restrict__ vin, uint32_t* __restrict__ vout,
                                                                    const uint32_t* __restrict__ n_ptr, uint32_t n_host, int shift,
                                                                    uint32_t mask, const uint32_t* __restrict__ bases,
                                                                    volatile uint32_t* status, int out_bits) {
  extern __shared__ __align__(16) char smem_raw[];
  SortSmemT<MODE == 0>& s = *reinterpret_cast<SortSmemT<MODE == 0>*>(smem_raw);
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = SORT_THREADS / 32;

  // CTAs are dispatched in blockIdx order: every predecessor the look-back waits on is resident.
  const uint32_t n = n_ptr ? min(*n_ptr, n_host) : n_host;
  const uint32_t tile = blockIdx.x;
  const uint32_t tile_base = tile * SORT_TILE;
  if (tile_base >= n) return;
  for (int i = tid; i < NW * SORT_RADIX; i += SORT_THREADS) (&s.warp_hist[0][0])[i] = 0;
  __syncthreads();
  const uint32_t count = min((uint32_t)SORT_TILE, n - tile_base);

  // ---- load (warp-striped: item i of lane l sits at warp_base + 32 i + l, so index order = (i, lane))
  uint64_t key[SORT_ITEMS];
  uint16_t rank[SORT_ITEMS];
  const uint32_t warp_base = warp * (32 * SORT_ITEMS);
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; i++) {
    const uint32_t pos = warp_base + 32 * i + lane;
    key[i] = pos < count ? __ldg(kin + tile_base + pos) : ~0ull;
  }
  // ---- stable ranking inside the warp with match_any
  const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; i++) {
    const uint32_t pos = warp_base + 32 * i + lane;
    const bool valid = pos < count;
    const uint32_t d = valid ? digit_of(key[i], shift, mask) : SORT_RADIX;  // invalid lanes form their own group
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    const uint32_t leader = __ffs(peers) - 1;
    uint32_t old = 0;
    if (valid && lane == leader) {
      old = s.warp_hist[warp][d];
      s.warp_hist[warp][d] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[i] = (uint16_t)(old + __popc(peers & lt_mask));
    __syncwarp();
  }
  __syncthreads();

  // ---- per-digit: exclusive offsets over warps, tile total, look-back
  {
    const uint32_t d = tid;  // SORT_THREADS == SORT_RADIX
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) {
      const uint32_t c = s.warp_hist[w][d];
      s.warp_hist[w][d] = run;
      run += c;
    }
    const uint32_t total = run;
    // decoupled look-back
    uint32_t excl = 0;
    volatile uint32_t* st = status + (size_t)tile * SORT_RADIX + d;
    if (tile == 0) {
      *st = total | FLAG_INCL;
    } else {
      *st = total | FLAG_AGG;
      int look = (int)tile - 1;
      while (true) {
        uint32_t w;
        do {
          w = status[(size_t)look * SORT_RADIX + d];
        } while ((w >> 30) == 0);
        excl += w & VALUE_MASK;
        if ((w >> 30) == 2) break;
        look--;
      }
      *st = (excl + total) | FLAG_INCL;
    }
    // CTA exclusive scan of totals over digits
    uint32_t incl = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s.warp_tot[warp] = incl;
    __syncthreads();
    uint32_t wbase = 0;
    for (uint32_t w = 0; w < warp; w++) wbase += s.warp_tot[w];
    const uint32_t lstart = wbase + incl - total;
    s.local_start[d] = lstart;
    s.adj[d] = __ldg(bases + d) + excl - lstart;
  }
  __syncthreads();

  // ---- place the tile in shared memory in sorted order
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; i++) {
    const uint32_t pos = warp_base + 32 * i + lane;
    if (pos < count) {
      const uint32_t d = digit_of(key[i], shift, mask);
      const uint32_t lp = s.local_start[d] + s.warp_hist[warp][d] + rank[i];
      s.keys[lp] = key[i];
      if (MODE == 0) s.vals[lp] = __ldg(vin + tile_base + pos);
    }
  }
  __syncthreads();
  // ---- write runs out
  const uint64_t out_mask = out_bits >= 64 ? ~0ull : (1ull << out_bits) - 1ull;
  for (uint32_t j = tid; j < count; j += SORT_THREADS) {
    const uint64_t k = s.keys[j];
    const uint32_t dst = s.adj[digit_of(k, shift, mask)] + j;
    if (MODE == 2) {
      vout[dst] = (uint32_t)(k & out_mask);
    } else {
      kout[dst] = k;
      if (MODE == 0) vout[dst] = s.vals[j];
    }
  }
}

template <int MODE> static void onesweep_set_smem() {
  static bool attr_set[64] = {};   // per device: the attribute belongs to the device's copy of the kernel
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaFuncSetAttribute(onesweep_pass_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SortSmemT<MODE == 0>));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
}

int launch_onesweep(uint64_t* keys[2], uint32_t* vals[2], const uint32_t* n_ptr, long long n, int end_bit, const SortTemp& t,
                    cudaStream_t stream) {
  const int passes = sort_passes(end_bit);
  if (n <= 0 || passes == 0) return 0;
  onesweep_set_smem<0>();
  sort_scan_hist_kernel<<<passes, SORT_RADIX, 0, stream>>>(t.hist);
  count_launch();
  const size_t ntiles = sort_num_tiles(n);
  int cur = 0;
  for (int p = 0; p < passes; p++) {
    const int shift = SORT_RADIX_BITS * p;
    const uint32_t mask = (1u << std::min(SORT_RADIX_BITS, end_bit - shift)) - 1u;
    onesweep_pass_kernel<0><<<(unsigned)ntiles, SORT_THREADS, sizeof(SortSmem), stream>>>(
        keys[cur], keys[cur ^ 1], vals[cur], vals[cur ^ 1], n_ptr, (uint32_t)n, shift, mask, t.hist + p * SORT_RADIX,
        t.status + (size_t)p * ntiles * SORT_RADIX, 64);
    count_launch();
    cur ^= 1;
  }
  return cur;
}

// ------------------------------------------------------------------ long lists
// Frames with tile lists beyond the shared-memory sort (config C5: 36 K entries per tile, 2.9e8 instances).  The
// reference materialises all instances as (tile|depth, id) pairs and radix-sorts them: six passes over 2.9e8 pairs.
// Here no instance is ever written before its final place:
//   1. the visible slots are compacted and sorted by (depth bits, slot) — the pair onesweep on the ~6e5 GAUSSIANS, of
//      which there are hundreds of times fewer than instances; their tile rectangles are gathered into that order;
//   2. tile-major scan: one CTA per super-tile of 4 x 4 tiles streams the sorted rectangles once, keeps — in order — the
//      candidates that overlap the super-tile in shared memory, and each of its 16 warps appends the candidates that
//      contain ITS tile to the tile's range of point_list (ballot + prefix: order preserved).  The k-th entry of a tile
//      is the k-th Gaussian in (depth, slot) order whose rectangle contains it: exactly the reference's sorted list.
// Per instance: 4 bytes written, once (the reference moves 12 + 8 + 6 x 24 = 164 B per instance).
constexpr int TB_THREADS = 512;
constexpr int TB_WARPS = TB_THREADS / 32;
constexpr int TB_SUPER = 4;                  // super-tile edge in tiles (TB_SUPER^2 == TB_WARPS)
constexpr int TB_CAND = 3840;                // candidate buffer entries (45 KB)
static_assert(TB_SUPER * TB_SUPER == TB_WARPS, "one warp per tile of the super-tile");

__device__ __forceinline__ uint32_t rect_tiles(uint2 rc) { return ((rc.x >> 16) - (rc.x & 0xffffu)) * ((rc.y >> 16) - (rc.y & 0xffffu)); }

// visible slots -> dense (depth bits, slot) pairs; also the per-forward duties of the bucket scatter: zero the backward
// accumulator rows of the visible slots, record the buffer header, reset the backward's unit counters
__global__ void __launch_bounds__(256) compact_visible_kernel(const uint32_t* __restrict__ block_vis, const uint32_t* __restrict__ vis_off,
                                                              const float* __restrict__ depths, uint64_t* __restrict__ gkeys,
                                                              uint32_t* __restrict__ gvals, uint32_t gcap, float* __restrict__ grad_acc,
                                                              BinHeader hv, BinHeader* header, uint32_t* unit_count) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (header) *header = hv;
    unit_count[0] = 0;
    unit_count[3] = unit_count[4] = unit_count[5] = unit_count[6] = 0;
  }
  const uint32_t cnt = block_vis[blockIdx.x];
  if (threadIdx.x >= cnt) return;
  const uint32_t slot = blockIdx.x * 256u + threadIdx.x, i = vis_off[blockIdx.x] + threadIdx.x;
  float4* acc = reinterpret_cast<float4*>(grad_acc + 12 * (size_t)slot);
  acc[0] = acc[1] = acc[2] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < gcap) {
    gkeys[i] = (uint64_t)__float_as_uint(__ldg(depths + slot));     // visible depths are > 0.2: the bits order like the values
    gvals[i] = slot;
  }
}

// tile rectangles in depth order (so that the scan below streams them with coalesced loads)
__global__ void __launch_bounds__(256) gather_sorted_rects_kernel(const uint32_t* __restrict__ order, const uint32_t* __restrict__ n_vis_ptr,
                                                                  uint32_t gcap, const uint2* __restrict__ rects, uint2* __restrict__ srect) {
  const uint32_t n_vis = min(*n_vis_ptr, gcap);
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;
  if (i < n_vis) srect[i] = __ldg(rects + __ldg(order + i));
}

constexpr int TB_ITEMS = 2;                  // sorted Gaussians per thread and chunk (consecutive: order = thread order)
constexpr int TB_CHUNK = TB_THREADS * TB_ITEMS;
static_assert(TB_CAND >= 2 * TB_CHUNK - TB_CHUNK / 2, "the candidate buffer must take a chunk on top of a partly filled buffer");

__global__ void __launch_bounds__(TB_THREADS, 4) tile_scan_bin_kernel(const uint32_t* __restrict__ order, const uint2* __restrict__ srect,
                                                                      const uint32_t* __restrict__ n_vis_ptr, uint32_t gcap,
                                                                      const uint2* __restrict__ ranges, uint32_t* __restrict__ point_list,
                                                                      uint32_t grid_x, uint32_t grid_y, uint32_t capacity) {
  __shared__ uint2 s_rect[TB_CAND];
  __shared__ uint32_t s_slot[TB_CAND];
  __shared__ uint32_t s_wcount[2][TB_WARPS];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  const uint32_t n_vis = min(*n_vis_ptr, gcap);
  // super-tile of this CTA and tile of this warp
  const uint32_t sgx = (grid_x + TB_SUPER - 1) / TB_SUPER;
  const uint32_t sx0 = (blockIdx.x % sgx) * TB_SUPER, sy0 = (blockIdx.x / sgx) * TB_SUPER;
  const uint32_t sx1 = min(sx0 + TB_SUPER, grid_x), sy1 = min(sy0 + TB_SUPER, grid_y);
  const uint32_t tx = sx0 + (warp % TB_SUPER), ty = sy0 + (warp / TB_SUPER);
  const bool have_tile = tx < grid_x && ty < grid_y;
  uint32_t cursor = have_tile ? ranges[ty * grid_x + tx].x : 0u;      // next free position of the tile's list (warp-uniform)

  auto flush = [&](uint32_t ncand) {
    // every warp walks the candidates in order and appends those whose rectangle contains its tile
    if (have_tile) {
      for (uint32_t c0 = 0; c0 < ncand; c0 += 32) {
        const uint32_t c = c0 + lane;
        bool h = false;
        if (c < ncand) {
          const uint2 rc = s_rect[c];
          h = tx >= (rc.x & 0xffffu) && tx < (rc.x >> 16) && ty >= (rc.y & 0xffffu) && ty < (rc.y >> 16);
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, h);
        if (h) {
          const uint32_t pos = cursor + __popc(bal & lt);
          if (pos < capacity) point_list[pos] = s_slot[c];
        }
        cursor += __popc(bal);
      }
    }
  };

  uint32_t ncand = 0;      // candidates buffered (identical in all threads)
  uint32_t par = 0;
  // Each thread owns TB_ITEMS consecutive entries of a chunk (two 16-byte loads of rectangles, one of slots), so the
  // order of the candidates is thread order.  The next chunk is requested one iteration ahead: the loop has a block
  // barrier per chunk and would otherwise pay a memory round trip between every two of them.
  uint2 rc_next[TB_ITEMS];
  uint32_t slot_next[TB_ITEMS];
  auto request = [&](uint32_t first) {
#pragma unroll
    for (int k = 0; k < TB_ITEMS; k++) {
      const uint32_t i = first + k;
      rc_next[k] = i < n_vis ? __ldg(srect + i) : make_uint2(0u, 0u);     // empty rectangle: never a candidate
      slot_next[k] = i < n_vis ? __ldg(order + i) : 0u;
    }
  };
  request(tid * TB_ITEMS);
  for (uint32_t base = 0; base < n_vis; base += TB_CHUNK, par ^= 1u) {
    uint2 rc[TB_ITEMS];
    uint32_t slot[TB_ITEMS];
    uint32_t mine = 0, hits = 0;
#pragma unroll
    for (int k = 0; k < TB_ITEMS; k++) {
      rc[k] = rc_next[k], slot[k] = slot_next[k];
      // rectangles are [min, max) in tiles
      const bool hit = (rc[k].x & 0xffffu) < sx1 && (rc[k].x >> 16) > sx0 && (rc[k].y & 0xffffu) < sy1 && (rc[k].y >> 16) > sy0;
      mine |= (hit ? 1u : 0u) << k;
      hits += hit ? 1u : 0u;
    }
    if (base + TB_CHUNK < n_vis) request(base + TB_CHUNK + tid * TB_ITEMS);
    // exclusive prefix of the per-thread hit counts inside the warp
    uint32_t incl = hits;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += v; }
    if (lane == 31) s_wcount[par][warp] = incl;
    __syncthreads();                                   // counts visible (double-buffered: one barrier per chunk)
    // prefix over the 16 warps, computed by every warp with shuffles
    uint32_t wc = lane < TB_WARPS ? s_wcount[par][lane] : 0u, wincl = wc;
#pragma unroll
    for (int o = 1; o < TB_WARPS; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, wincl, o); if (lane >= (uint32_t)o) wincl += v; }
    const uint32_t total = __shfl_sync(0xffffffffu, wincl, TB_WARPS - 1);
    const uint32_t before = __shfl_sync(0xffffffffu, wincl - wc, warp);
    if (ncand + total > TB_CAND) {                     // does not fit on top of what is buffered: drain first (uniform decision)
      flush(ncand);
      ncand = 0;
      __syncthreads();                                 // every warp is done reading the buffer
    }
    uint32_t pos = ncand + before + incl - hits;
#pragma unroll
    for (int k = 0; k < TB_ITEMS; k++) {
      if (mine & (1u << k)) {
        s_rect[pos] = rc[k];
        s_slot[pos] = slot[k];
        pos++;
      }
    }
    ncand += total;
    if (ncand > TB_CAND - TB_CHUNK) {                  // the next chunk might not fit: make the candidates visible for the next drain
      __syncthreads();
      flush(ncand);
      ncand = 0;
      __syncthreads();
    }
  }
  __syncthreads();
  flush(ncand);
}

void launch_long_bin(int P, int T, uint32_t grid_x, uint32_t grid_y, const uint2* ranges, const GeometryView& g, const BinningView& bl,
                     long long capacity, BinHeader hv, BinHeader* header, cudaStream_t stream) {
  if (P <= 0 || T <= 0) return;
  const uint32_t nblocks = (uint32_t)num_pre_blocks(P);
  const uint32_t gcap = (uint32_t)bl.gcap;
  const uint32_t gblocks = (gcap + 255u) / 256u;
  uint32_t* n_vis = g.counters + 13;
  // 1. visible slots, compacted, keyed by depth bits
  scan_blocks_kernel<<<1, 1024, 0, stream>>>(g.block_vis, g.block_off, nblocks, n_vis);
  compact_visible_kernel<<<nblocks, 256, 0, stream>>>(g.block_vis, g.block_off, g.depths, bl.gkeys[0], bl.gvals[0], gcap, g.grad_acc, hv,
                                                     header, g.counters + 5);
  count_launch(2);
  // 2. sorted by (depth bits, slot): stable pair sort, the input is in slot order
  SortTemp st;
  carve_sort_temp(bl.gsort_temp, gcap, sort_passes(32), st);
  sort_temp_reset(bl.gsort_temp, gcap, sort_passes(32), stream);
  uint64_t* gk[2] = {bl.gkeys[0], bl.gkeys[1]};
  uint32_t* gv[2] = {bl.gvals[0], bl.gvals[1]};
  launch_sort_histogram(gk[0], n_vis, gcap, 0, 32, st.hist, stream);
  const int cur = launch_onesweep(gk, gv, n_vis, gcap, 32, st, stream);
  const uint32_t* order = gv[cur];
  uint2* srect = reinterpret_cast<uint2*>(gk[cur ^ 1]);      // the other key buffer is free now: 8 bytes per entry
  gather_sorted_rects_kernel<<<gblocks, 256, 0, stream>>>(order, n_vis, gcap, g.rect, srect);
  // 3. tile-major scan into point_list
  const uint32_t sgx = (grid_x + TB_SUPER - 1) / TB_SUPER, sgy = (grid_y + TB_SUPER - 1) / TB_SUPER;
  tile_scan_bin_kernel<<<sgx * sgy, TB_THREADS, 0, stream>>>(order, srect, n_vis, gcap, ranges, bl.point_list, grid_x, grid_y,
                                                           (uint32_t)capacity);
  count_launch(2);
}

}  // namespace gsr

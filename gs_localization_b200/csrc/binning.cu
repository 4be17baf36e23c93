// Tile binning: key generation, a hand-written stable onesweep LSD radix sort of
// (u64 tile|depth key, u32 Gaussian id) pairs, and tile-range extraction.
//
// Replaces, from the reference (gaussian_splatting/submodules/diff-gaussian-rasterization):
//   duplicateWithKeys                         cuda_rasterizer/rasterizer_impl.cu:70-111
//   cub::DeviceRadixSort::SortPairs           cuda_rasterizer/rasterizer_impl.cu:304-309 (size query :187-190)
//   cudaMemset(ranges) + identifyTileRanges   cuda_rasterizer/rasterizer_impl.cu:311-318, 116-138
//
// B200 design
//   * duplicate_with_keys also accumulates the per-digit histograms every sort pass needs
//     (one shared-memory histogram per CTA, one atomic per Gaussian for the four depth
//     digits because all of a Gaussian's instances share its depth), so the sort needs no
//     histogram pass over the 8R bytes of keys.
//   * onesweep (Adinets & Merrill): one kernel per 8-bit digit; each CTA ranks a 4096-key
//     tile with warp match_any (stable), chains its per-digit counts to its predecessors
//     with a decoupled look-back, stages the tile in shared memory in sorted order and
//     writes runs of equal digits out contiguously.  Each pass reads and writes every
//     pair exactly once (12R + 12R bytes).
//   * Stability is required: instances with equal (tile, depth bits) must stay in Gaussian
//     order as with the reference's CUB sort, otherwise point_list diverges.
#include "gsr_kernels.cuh"

namespace gsr {

// ------------------------------------------------------------------ sort temp layout
size_t sort_temp_bytes(long long n, int passes) {
  size_t b = 0;
  b += align_up(sizeof(uint32_t) * SORT_MAX_PASSES * SORT_RADIX, 128);
  b += align_up(sizeof(uint32_t) * 32, 128);
  b += align_up(sizeof(uint32_t) * (size_t)passes * sort_num_tiles(n) * SORT_RADIX, 128);
  return b + 128;
}
void carve_sort_temp(char* base, long long n, int passes, SortTemp& t) {
  char* p = base;
  carve(p, t.hist, (size_t)SORT_MAX_PASSES * SORT_RADIX);
  carve(p, t.tickets, (size_t)32);
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

// ------------------------------------------------------------------ key generation
// One CTA expands 256 consecutive visible Gaussians.  The CTA's output range is contiguous
// (prefix sums), so all 256 threads walk it together: output slot j finds its owner with a
// binary search over the CTA-local prefix sums in shared memory, decodes its tile from the
// owner's rectangle (rows then columns, like the reference's nested loop) and writes
// key = tile << 32 | depth bits, value = visible rank.  Writes are fully coalesced and a
// Gaussian covering thousands of tiles costs no more per key than one covering four.
__global__ void __launch_bounds__(256) emit_keys_kernel(const uint32_t* __restrict__ counters,
                                                        const uint32_t* __restrict__ tiles_touched,
                                                        const uint32_t* __restrict__ point_offsets,
                                                        const uint2* __restrict__ rects, const float* __restrict__ depths,
                                                        uint64_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                                        uint32_t grid_x, uint32_t capacity) {
  __shared__ uint32_t s_end[256];     // CTA-local inclusive prefix of tiles
  __shared__ uint2 s_rect[256];
  __shared__ uint32_t s_depth[256];
  const uint32_t Pv = counters[2];
  const uint32_t first = blockIdx.x * 256u;
  if (first >= Pv) return;
  const uint32_t tid = threadIdx.x;
  const uint32_t k = first + tid;
  const uint32_t cnt = min(256u, Pv - first);
  const uint32_t block_begin = first == 0 ? 0u : __ldg(point_offsets + first - 1);
  if (tid < cnt) {
    s_end[tid] = __ldg(point_offsets + k) - block_begin;
    s_rect[tid] = __ldg(rects + k);
    s_depth[tid] = __float_as_uint(__ldg(depths + k));
  }
  __syncthreads();
  const uint32_t total = s_end[cnt - 1];
  for (uint32_t j = tid; j < total; j += 256u) {
    // smallest t with s_end[t] > j
    uint32_t lo = 0, hi = cnt - 1;
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (s_end[mid] > j) hi = mid; else lo = mid + 1;
    }
    const uint32_t begin = lo == 0 ? 0u : s_end[lo - 1];
    const uint32_t within = j - begin;
    const uint2 rc = s_rect[lo];
    const uint32_t minx = rc.x & 0xffffu, w = (rc.x >> 16) - minx, miny = rc.y & 0xffffu;
    const uint32_t row = within / w, col = within - row * w;
    const uint32_t tile = (miny + row) * grid_x + (minx + col);
    const uint32_t dst = block_begin + j;
    if (dst < capacity) {
      keys[dst] = ((uint64_t)tile << 32) | s_depth[lo];
      vals[dst] = first + lo;
    }
  }
}

void launch_emit_keys(int max_visible, const GeometryView& g, uint64_t* keys, uint32_t* vals, uint32_t grid_x,
                      uint32_t capacity, cudaStream_t stream) {
  if (max_visible <= 0) return;
  emit_keys_kernel<<<(max_visible + 255) / 256, 256, 0, stream>>>(g.counters, g.tiles_touched, g.point_offsets, g.rect,
                                                               g.depths, keys, vals, grid_x, capacity);
  count_launch();
}

// ------------------------------------------------------------------ digit histograms of all passes
// Keys of one Gaussian are adjacent and share every depth digit, so equal digits are first
// aggregated inside the warp (match_any) and cost one shared-memory atomic per group.
__global__ void __launch_bounds__(256) sort_histogram_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ n_ptr,
                                                             uint32_t n_host, int end_bit, uint32_t* __restrict__ hist) {
  __shared__ uint32_t s_hist[SORT_MAX_PASSES * SORT_RADIX];
  const uint32_t n = n_ptr ? min(*n_ptr, n_host) : n_host;
  const int passes = (end_bit + SORT_RADIX_BITS - 1) / SORT_RADIX_BITS;
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
        const int shift = SORT_RADIX_BITS * p;
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
void launch_sort_histogram(const uint64_t* keys, const uint32_t* n_ptr, long long n_host, int end_bit, uint32_t* hist,
                           cudaStream_t stream) {
  if (n_host <= 0) return;
  const int blocks = (int)std::min<long long>((n_host + 256 * 8 - 1) / (256 * 8), 148 * 4);
  sort_histogram_kernel<<<blocks, 256, 0, stream>>>(keys, n_ptr, (uint32_t)n_host, end_bit, hist);
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

struct __align__(16) SortSmem {
  uint64_t keys[SORT_TILE];
  uint32_t vals[SORT_TILE];
  uint32_t warp_hist[SORT_THREADS / 32][SORT_RADIX];
  uint32_t local_start[SORT_RADIX];
  uint32_t adj[SORT_RADIX];
  uint32_t warp_tot[SORT_RADIX / 32];
  uint32_t tile;
};

__global__ void __launch_bounds__(SORT_THREADS) onesweep_pass_kernel(const uint64_t* __restrict__ kin, uint64_t* __restrict__ kout,
                                                                    const uint32_t* __restrict__ vin, uint32_t* __restrict__ vout,
                                                                    const uint32_t* __restrict__ n_ptr, uint32_t n_host, int shift,
                                                                    uint32_t mask, const uint32_t* __restrict__ bases,
                                                                    volatile uint32_t* status) {
  extern __shared__ __align__(16) char smem_raw[];
  SortSmem& s = *reinterpret_cast<SortSmem*>(smem_raw);
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
      s.vals[lp] = __ldg(vin + tile_base + pos);
    }
  }
  __syncthreads();
  // ---- write runs out
  for (uint32_t j = tid; j < count; j += SORT_THREADS) {
    const uint64_t k = s.keys[j];
    const uint32_t dst = s.adj[digit_of(k, shift, mask)] + j;
    kout[dst] = k;
    vout[dst] = s.vals[j];
  }
}

int launch_onesweep(uint64_t* keys[2], uint32_t* vals[2], const uint32_t* n_ptr, long long n, int end_bit, const SortTemp& t,
                    cudaStream_t stream) {
  const int passes = sort_passes(end_bit);
  if (n <= 0 || passes == 0) return 0;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(onesweep_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SortSmem));
    attr_set = true;
  }
  sort_scan_hist_kernel<<<passes, SORT_RADIX, 0, stream>>>(t.hist);
  count_launch();
  const size_t ntiles = sort_num_tiles(n);
  int cur = 0;
  for (int p = 0; p < passes; p++) {
    const int shift = SORT_RADIX_BITS * p;
    const uint32_t mask = (1u << std::min(SORT_RADIX_BITS, end_bit - shift)) - 1u;
    onesweep_pass_kernel<<<(unsigned)ntiles, SORT_THREADS, sizeof(SortSmem), stream>>>(
        keys[cur], keys[cur ^ 1], vals[cur], vals[cur ^ 1], n_ptr, (uint32_t)n, shift, mask, t.hist + p * SORT_RADIX,
        t.status + (size_t)p * ntiles * SORT_RADIX);
    count_launch();
    cur ^= 1;
  }
  return cur;
}

// ------------------------------------------------------------------ tile ranges
__global__ void __launch_bounds__(256) identify_tile_ranges_kernel(const uint32_t* __restrict__ n_ptr, uint32_t n_host,
                                                                   const uint64_t* __restrict__ keys,
                                                                   uint2* __restrict__ ranges) {
  const uint32_t L = n_ptr ? min(*n_ptr, n_host) : n_host;
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= L) return;
  const uint32_t cur = (uint32_t)(__ldg(keys + idx) >> 32);
  if (idx == 0) {
    ranges[cur].x = 0;
  } else {
    const uint32_t prev = (uint32_t)(__ldg(keys + idx - 1) >> 32);
    if (cur != prev) {
      ranges[prev].y = idx;
      ranges[cur].x = idx;
    }
  }
  if (idx == L - 1) ranges[cur].y = L;
}

void launch_identify_tile_ranges(const uint32_t* n_ptr, long long n, const uint64_t* keys, uint2* ranges, int num_tiles,
                                 cudaStream_t stream) {
  cudaMemsetAsync(ranges, 0, sizeof(uint2) * (size_t)num_tiles, stream);
  if (n > 0) {
    identify_tile_ranges_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(n_ptr, (uint32_t)n, keys, ranges);
    count_launch();
  }
}

}  // namespace gsr

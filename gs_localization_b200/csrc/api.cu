// C ABI of the rasterizer (include/gsr_b200.h): host orchestration of the sm_100a kernels.
//
// Replaces CudaRasterizer::Rasterizer::{forward, backward, markVisible}
// (reference cuda_rasterizer/rasterizer_impl.cu:197-339, 343-444, 141-153) and the
// buffer-carving helpers (rasterizer_impl.cu:155-193, rasterizer_impl.h:21-27,66-72).
#include <algorithm>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "../../include/gsr_b200.h"
#include "gsr_kernels.cuh"

namespace gsr {
std::atomic<unsigned long long> g_launch_count{0};

int sm_count() {
  static std::atomic<int> cached[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int n = cached[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}
}
using namespace gsr;

namespace {
thread_local std::string t_error;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  t_error = buf;
  return code;
}

#define GSR_CUDA(expr)                                                                         \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess) return fail(GSR_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

// mirrors CHECK_CUDA (reference auxiliary.h:166-173): with debug, synchronise after each stage and surface errors
#define GSR_STAGE(name, debug, stream)                                                          \
  do {                                                                                          \
    cudaError_t e__ = cudaGetLastError();                                                       \
    if (e__ == cudaSuccess && (debug)) e__ = cudaStreamSynchronize(stream);                     \
    if (e__ != cudaSuccess) return fail(GSR_ERR_CUDA, "stage %s: %s", name, cudaGetErrorString(e__)); \
  } while (0)

// Binning buffer: [header 128 B][point_list u32[cap]] then, depending on the path,
//   tile-local sort : [comp u64[cap]]
//   long lists      : [scratch of the Gaussian-level depth sort] (instances go straight to point_list)
// The backward only needs point_list, whose offset depends on neither the path nor the capacity the
// forward happened to allocate (speculative launches over-allocate).
char* carve_binning(char* base, long long cap, BinningView& b, bool global_path = false, int W = 0, int H = 0, int P = 0) {
  char* p = base + 128;
  carve(p, b.point_list, (size_t)cap);
  b.keys[0] = b.keys[1] = nullptr, b.sort_temp = nullptr, b.comp = nullptr;
  b.gkeys[0] = b.gkeys[1] = nullptr, b.gvals[0] = b.gvals[1] = nullptr, b.gsort_temp = nullptr, b.gcap = 0;
  if (!global_path) {
    carve(p, b.comp, (size_t)cap);
  } else {
    // Gaussian-level depth sort: at most min(slots, capacity) visible Gaussians (each has at least one instance)
    b.gcap = std::min<long long>((long long)num_pre_blocks(P) * PRE_THREADS, cap);
    carve(p, b.gkeys[0], (size_t)b.gcap);
    carve(p, b.gkeys[1], (size_t)b.gcap);
    carve(p, b.gvals[0], (size_t)b.gcap);
    carve(p, b.gvals[1], (size_t)b.gcap);
    p = (char*)align_up((size_t)p, 128);
    b.gsort_temp = p;
    p += sort_temp_bytes(b.gcap, sort_passes(32));
  }
  b.units = nullptr, b.ckpt = nullptr, b.rec = nullptr, b.units_off = b.ckpt_off = b.rec_off = b.units_cap = 0;
  if (W > 0 && H > 0) {
    const size_t tiles = (size_t)((W + TILE_X - 1) / TILE_X) * ((H + TILE_Y - 1) / TILE_Y);
    const size_t nu = max_units(cap, tiles);
    b.units_cap = 4 * BSEG_PER_SEG * nu;
    carve(p, b.units, 2 * b.units_cap);
    carve(p, b.ckpt, BSEG_PER_SEG * nu * CKPT_FLOATS);
    carve(p, b.rec, nu * REC_FLOAT4);
    b.units_off = (size_t)((char*)b.units - base), b.ckpt_off = (size_t)((char*)b.ckpt - base);
    b.rec_off = (size_t)((char*)b.rec - base);
  }
  return p;
}
constexpr long long LOCAL_SORT_MAX = 4096;   // longest tile list the shared-memory tile sort handles (binning.cu TS_SMEM_KEYS)

// ---- optional per-stage device timing (bench.py's stage split; off on the hot path)
enum Stage { ST_PREPROCESS, ST_DUPLICATE, ST_SORT, ST_RANGES, ST_RENDER, ST_BWD_RENDER, ST_BWD_PREPROCESS, ST_COUNT };
struct StageTimer {
  bool enabled = false;
  cudaEvent_t ev[ST_COUNT][2] = {};
  bool used[ST_COUNT] = {};
  double total_ms[ST_COUNT] = {};
  unsigned long long calls[ST_COUNT] = {};
  std::mutex mu;
};
StageTimer g_timer;

struct StageScope {
  int st;
  cudaStream_t stream;
  bool on;
  StageScope(int st_, cudaStream_t s) : st(st_), stream(s), on(g_timer.enabled) {
    if (!on) return;
    if (!g_timer.ev[st][0]) {
      cudaEventCreate(&g_timer.ev[st][0]);
      cudaEventCreate(&g_timer.ev[st][1]);
    }
    cudaEventRecord(g_timer.ev[st][0], stream);
  }
  ~StageScope() {
    if (!on) return;
    cudaEventRecord(g_timer.ev[st][1], stream);
    g_timer.used[st] = true;
  }
};
void stage_collect(cudaStream_t stream) {
  if (!g_timer.enabled) return;
  cudaStreamSynchronize(stream);
  std::lock_guard<std::mutex> lk(g_timer.mu);
  for (int i = 0; i < ST_COUNT; i++) {
    if (!g_timer.used[i]) continue;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g_timer.ev[i][0], g_timer.ev[i][1]) == cudaSuccess) {
      g_timer.total_ms[i] += ms;
      g_timer.calls[i]++;
    }
    g_timer.used[i] = false;
  }
}

// Side stream for the colour (SH) kernel of the forward, which nothing before the blend depends on: it runs
// concurrently with scan_tiles / scatter / tile_sort and re-joins before the blend.  Fork/join through events,
// which is also legal while the caller's stream is being captured into a CUDA graph.
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
SideStream* side_stream() {
  // per host thread and device: the autograd engine runs backward on its own thread, and two threads must never
  // re-record each other's fork/join events
  thread_local SideStream per_device[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  SideStream& s = per_device[dev];
  if (!s.stream) {
    if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  }
  return &s;
}

// pinned mailbox for num_rendered + the event that says it has landed, one per host thread
struct Mailbox {
  uint32_t* value = nullptr;
  cudaEvent_t ready = nullptr;
};
Mailbox* pinned_mailbox() {
  thread_local Mailbox box;
  if (!box.value) {
    if (cudaHostAlloc((void**)&box.value, 64, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&box.ready, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  }
  return &box;
}

// Binning-capacity history for the speculative launch: the last 16 num_rendered values seen per (P, W, H),
// per host thread.  The guess is their maximum plus 25 % headroom: consecutive refinement iterations barely
// change num_rendered, and a caller cycling through a handful of views is covered by the maximum.
struct CountHistory {
  int P = -1, W = 0, H = 0, n = 0, pos = 0;
  long long R[16] = {}, longest[16] = {};
};
thread_local CountHistory t_hist[4];
// Capacities are rounded up to 1/8-octave steps so that successive forwards ask the caller's allocator for a small
// set of sizes (a caching allocator then reuses its blocks instead of growing the buffer a few percent at a time).
long long round_capacity(long long c) {
  if (c <= 4096) return 4096;
  long long step = 1;
  while ((step << 4) <= c) step <<= 1;   // step = 2^(floor(log2 c) - 3)
  return (c + step - 1) / step * step;
}
void capacity_guess(int P, int W, int H, long long& cap, long long& longest) {
  cap = longest = 0;
#ifdef GSR_AB_NO_SPECULATION      // measurement build: wait for num_rendered like the reference
  return;
#endif
  for (auto& h : t_hist)
    if (h.P == P && h.W == W && h.H == H && h.n > 0) {
      long long m = 0, l = 0;
      for (int i = 0; i < h.n; i++) m = std::max(m, h.R[i]), l = std::max(l, h.longest[i]);
      cap = round_capacity(m + m / 4 + 4096);
      longest = l + l / 4;
      return;
    }
}
void remember_count(int P, int W, int H, long long R, long long longest) {
  for (auto& h : t_hist)
    if (h.P == P && h.W == W && h.H == H) {
      h.R[h.pos] = R, h.longest[h.pos] = longest;
      h.pos = (h.pos + 1) % 16;
      h.n = std::min(h.n + 1, 16);
      return;
    }
  static thread_local int next = 0;
  CountHistory& h = t_hist[next];
  h = CountHistory{};
  h.P = P, h.W = W, h.H = H, h.n = 1, h.pos = 1, h.R[0] = R, h.longest[0] = longest;
  next = (next + 1) % 4;
}
}  // namespace

extern "C" {

int gsr_abi_version(void) { return GSR_ABI_VERSION; }
const char* gsr_last_error(void) { return t_error.c_str(); }
unsigned long long gsr_launch_count(void) { return g_launch_count.load(); }

size_t gsr_geometry_bytes(int P) {
  GeometryView g;
  char* end = carve_geometry(nullptr, P, g);
  return (size_t)end + 128;
}
size_t gsr_image_bytes(int width, int height) {
  ImageView im;
  char* end = carve_image(nullptr, width, height, im);
  return (size_t)end + 128;
}
size_t gsr_binning_bytes(long long num_rendered, int width, int height) {
  // worst case over the two layouts (the long-list one with as many visible Gaussians as instances)
  BinningView b;
  const int p_bound = num_rendered > 0x7fffff00ll ? 0x7fffff00 : (int)num_rendered;
  char* end_long = carve_binning(nullptr, num_rendered, b, true, width, height, p_bound);
  char* end_local = carve_binning(nullptr, num_rendered, b, false, width, height, p_bound);
  return (size_t)std::max(end_long, end_local) + 128;
}
static size_t binning_bytes_for(long long cap, bool global_path, int width, int height, int P) {
  BinningView b;
  char* end = carve_binning(nullptr, cap, b, global_path, width, height, P);
  return (size_t)end + 128;
}

long long gsr_rasterize_forward(gsr_alloc_fn geometry_alloc, gsr_alloc_fn binning_alloc, gsr_alloc_fn image_alloc, void* user,
                                int P, int D, int M, const float* background, int width, int height, const float* means3D,
                                const float* shs, const float* colors_precomp, const float* opacities, const float* scales,
                                float scale_modifier, const float* rotations, const float* cov3D_precomp,
                                const float* viewmatrix, const float* projmatrix, const float* cam_pos, float tan_fovx,
                                float tan_fovy, int prefiltered, float* out_color, float* out_depth, float* out_alpha,
                                int* radii, int* n_touched, int debug, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (P < 0 || width <= 0 || height <= 0) return fail(GSR_ERR_INVALID_ARGUMENT, "bad sizes P=%d W=%d H=%d", P, width, height);
  if (!geometry_alloc || !binning_alloc || !image_alloc) return fail(GSR_ERR_INVALID_ARGUMENT, "allocation callbacks are required");
  if (!background || !viewmatrix || !projmatrix || !cam_pos || !out_color || !out_depth || !out_alpha)
    return fail(GSR_ERR_INVALID_ARGUMENT, "null required pointer");
  if (P > 0) {
    if (!means3D || !opacities || !radii) return fail(GSR_ERR_INVALID_ARGUMENT, "means3D, opacities and radii are required");
    if (!shs && !colors_precomp) return fail(GSR_ERR_UNSUPPORTED, "provide either SHs or precomputed colors");
    if (!cov3D_precomp && (!scales || !rotations)) return fail(GSR_ERR_INVALID_ARGUMENT, "provide scales+rotations or cov3D_precomp");
    if (!colors_precomp && (M <= 0 || (D + 1) * (D + 1) > M || D > 3))
      return fail(GSR_ERR_INVALID_ARGUMENT, "SH degree %d needs %d coefficients, %d stored", D, (D + 1) * (D + 1), M);
    if (width > 16 * 65535 || height > 16 * 65535) return fail(GSR_ERR_INVALID_ARGUMENT, "image too large");
  }
  const uint32_t gx = (width + TILE_X - 1) / TILE_X, gy = (height + TILE_Y - 1) / TILE_Y;
  const int T = (int)(gx * gy);

  char* img_base = image_alloc(gsr_image_bytes(width, height), user);
  if (!img_base) return fail(GSR_ERR_ALLOC, "image_alloc returned NULL");
  ImageView im;
  carve_image(img_base, width, height, im);

  long long R = 0;
  GeometryView g{};
  BinningView bl{};
  Mailbox* box = nullptr;
  SideStream* color_side = nullptr;
  bool color_joined = false, count_waited = false;
  // Whatever path leaves this function (error returns included): the num_rendered copy into this thread's pinned mailbox
  // has landed — a later forward must never see a stale copy arrive after it wrote its sentinel — and the colour side
  // stream has re-joined the caller's stream, so that its kernel cannot race with the caller freeing the scratch.
  struct Cleanup {
    Mailbox*& box; SideStream*& side; bool& joined; bool& waited; cudaStream_t stream;
    ~Cleanup() {
      if (box && !waited) cudaEventSynchronize(box->ready);
      if (side && !joined) cudaStreamWaitEvent(stream, side->join, 0);
    }
  } cleanup{box, color_side, color_joined, count_waited, stream};
  if (P > 0) {
    char* geom_base = geometry_alloc(gsr_geometry_bytes(P), user);
    if (!geom_base) return fail(GSR_ERR_ALLOC, "geometry_alloc returned NULL");
    carve_geometry(geom_base, P, g);
    // counters and the coverage grid are zeroed by the first preprocess kernel

    PreprocessParams pp{};
    pp.P = P, pp.D = D, pp.M = M, pp.W = width, pp.H = height, pp.grid_x = gx, pp.grid_y = gy;
    pp.means3D = means3D, pp.scales = scales, pp.rotations = rotations, pp.opacities = opacities, pp.shs = shs;
    pp.cov3D_precomp = cov3D_precomp, pp.colors_precomp = colors_precomp;
    pp.viewmatrix = viewmatrix, pp.projmatrix = projmatrix, pp.campos = cam_pos;
    pp.scale_modifier = scale_modifier, pp.tan_fovx = tan_fovx, pp.tan_fovy = tan_fovy;
    pp.focal_y = height / (2.0f * tan_fovy);   // reference rasterizer_impl.cu:223-224
    pp.focal_x = width / (2.0f * tan_fovx);
    pp.prefiltered = prefiltered;
    pp.sh_vec4 = shs && (M % 4 == 0) && ((uintptr_t)shs % 16 == 0);
    pp.radii = radii, pp.n_touched = n_touched, pp.tile_diff = im.tile_diff, pp.geom = g;
    {
      StageScope ts(ST_PREPROCESS, stream);
      launch_preprocess_fwd(pp, stream);
      // colours on the side stream, concurrent with the binning kernels; joined before the blend
      color_side = side_stream();
      if (color_side && !debug && !g_timer.enabled) {
        GSR_CUDA(cudaEventRecord(color_side->fork, stream));
        GSR_CUDA(cudaStreamWaitEvent(color_side->stream, color_side->fork, 0));
        launch_color_fwd(pp, color_side->stream);
        GSR_CUDA(cudaEventRecord(color_side->join, color_side->stream));
      } else {
        color_side = nullptr;
        launch_color_fwd(pp, stream);
      }
    }
    GSR_STAGE("preprocess", debug, stream);
    {
      StageScope ts(ST_RANGES, stream);
      launch_scan_tiles(im.tile_diff, gx, gy, im.ranges, im.tile_cursor, im.tile_order, g.counters, 0, stream);
    }
    GSR_STAGE("scan_tiles", debug, stream);

    // num_rendered -> pinned host mailbox, asynchronously; `r_ready` fires when it has landed
    box = pinned_mailbox();
    if (!box) return fail(GSR_ERR_CUDA, "cudaHostAlloc failed");
    *reinterpret_cast<volatile uint32_t*>(box->value) = 0xffffffffu;   // sentinel: num_rendered is always < 2^31
    GSR_CUDA(cudaMemcpyAsync(box->value + 4, g.counters + 4, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));   // longest tile list
    GSR_CUDA(cudaMemcpyAsync(box->value, g.counters + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));       // num_rendered (last: it is the flag)
    GSR_CUDA(cudaEventRecord(box->ready, stream));
  } else {
    GSR_CUDA(cudaMemsetAsync(im.ranges, 0, sizeof(uint2) * (size_t)T, stream));
  }

  // The reference blocks here until num_rendered is on the host (rasterizer_impl.cu:282) and only then
  // sizes the binning buffer and launches the rest, so the GPU idles for a host round trip every
  // forward.  Here the rest of the forward is launched SPECULATIVELY into a binning buffer sized from the
  // previous call with the same (P, W, H); the host then waits for num_rendered while the GPU keeps
  // working.  If the guess was too small (kernels clamp to the capacity, nothing overruns) the tail is
  // simply re-run with the exact size.  First call / no history: the reference's order.
  BinHeader* bin_header = nullptr;
  auto run_tail = [&](long long capacity, bool global_path) -> int {
    if (capacity > 0 && !global_path) {
      {
        StageScope ts(ST_DUPLICATE, stream);
        launch_scatter(P, g, im.tile_cursor, bl.comp, gx, BinHeader{(unsigned long long)capacity, bl.units_off, bl.ckpt_off, bl.rec_off, bl.units_cap}, bin_header, stream);
      }
      GSR_STAGE("scatter", debug, stream);
      {
        StageScope ts(ST_SORT, stream);
        launch_tile_sort(T, im.ranges, bl.comp, bl.point_list, (uint32_t)capacity, im.tile_order, stream);
      }
      GSR_STAGE("tile_sort", debug, stream);
    } else if (capacity > 0) {
      // tile lists too long for the shared-memory sort (binning.cu, "long lists")
      {
        StageScope ts(ST_SORT, stream);
        launch_long_bin(P, T, gx, gy, im.ranges, g, bl, capacity, BinHeader{(unsigned long long)capacity, bl.units_off, bl.ckpt_off, bl.rec_off, bl.units_cap}, bin_header, stream);
      }
      GSR_STAGE("radix_sort", debug, stream);
    }
    RenderParams rp{};
    rp.W = width, rp.H = height, rp.grid_x = gx, rp.grid_y = gy;
    rp.ranges = im.ranges, rp.point_list = bl.point_list, rp.tile_order = (P > 0) ? im.tile_order : nullptr;
    rp.mean_tau = g.mean_tau, rp.conic_opacity = g.conic_opacity, rp.rgbd = g.rgbd, rp.gid = g.gid;
    rp.bg = background, rp.out_color = out_color, rp.out_depth = out_depth, rp.out_alpha = out_alpha;
    rp.n_contrib = im.n_contrib, rp.n_touched = (P > 0) ? n_touched : nullptr;
    rp.capacity = (uint32_t)capacity;
    rp.final_cd = im.final_cd, rp.units = bl.units, rp.ckpt = bl.ckpt, rp.rec = bl.rec, rp.units_cap = (uint32_t)bl.units_cap, rp.unit_count = (P > 0) ? g.counters + 5 : nullptr;
    if (color_side && !color_joined) {
      GSR_CUDA(cudaStreamWaitEvent(stream, color_side->join, 0));
      color_joined = true;
    }
    {
      StageScope ts(ST_RENDER, stream);
      launch_render_fwd(rp, stream);
    }
    GSR_STAGE("render", debug, stream);
    return GSR_OK;
  };
  auto alloc_binning = [&](long long cap, bool global_path) -> int {
    char* bin_base = binning_alloc(binning_bytes_for(cap, global_path, width, height, P), user);
    if (!bin_base) return fail(GSR_ERR_ALLOC, "binning_alloc returned NULL");
    carve_binning(bin_base, cap, bl, global_path, width, height, P);
    bin_header = reinterpret_cast<BinHeader*>(bin_base);   // capacity and the unit/checkpoint offsets are recorded there by the scatter kernel
    return GSR_OK;
  };

  long long guess = 0, guess_longest = 0;
  if (P > 0 && !debug && !g_timer.enabled) capacity_guess(P, width, height, guess, guess_longest);
  bool speculated = false, spec_global = guess_longest > LOCAL_SORT_MAX;
  if (guess > 0) {
    if (guess >= (1ll << 30)) return fail(GSR_ERR_UNSUPPORTED, "more than 2^30 instances");
    if (int rc = alloc_binning(guess, spec_global)) return rc;
    if (int rc = run_tail(guess, spec_global)) return rc;
    speculated = true;
  }
  long long longest = 0;
  if (P > 0) {
    // poll the pinned word the copy engine writes (cheaper than a driver-level wait), with the event as a safety net
    {
      volatile uint32_t* v = reinterpret_cast<volatile uint32_t*>(box->value);
      int spins = 0;
      while (*v == 0xffffffffu) {
        if (++spins > 2000) {
          const cudaError_t q = cudaEventQuery(box->ready);
          if (q == cudaSuccess) break;
          if (q != cudaErrorNotReady) return fail(GSR_ERR_CUDA, "waiting for num_rendered: %s", cudaGetErrorString(q));
          spins = 0;
        }
      }
      if (*v == 0xffffffffu) GSR_CUDA(cudaEventSynchronize(box->ready));
    }
    count_waited = true;
    R = (long long)box->value[0];
    longest = (long long)box->value[4];
    remember_count(P, width, height, R, longest);
    if (R >= (1ll << 30)) return fail(GSR_ERR_UNSUPPORTED, "more than 2^30 instances (%lld)", R);
  }
  if (!speculated || R > guess) {
    if (speculated) {
      // overflow: the scatter consumed the cursors; rebuild them from the ranges before re-running
      GSR_CUDA(cudaStreamSynchronize(stream));
      launch_reset_cursors(T, im.ranges, im.tile_cursor, stream);
      if (n_touched) GSR_CUDA(cudaMemsetAsync(n_touched, 0, sizeof(int) * (size_t)P, stream));
    }
    const bool global_path = longest > LOCAL_SORT_MAX;
    const long long cap = R > 0 ? round_capacity(R) : 0;
    if (int rc = alloc_binning(cap, global_path)) return rc;
    if (int rc = run_tail(cap, global_path)) return rc;
  }
  stage_collect(stream);
  return R;
}


// Sync-free forward for callers that keep their own buffers (the fused pose-refinement loop): no allocation
// callbacks, no host wait, nothing that cannot be captured into a CUDA graph.  num_rendered stays on the device
// (read it later with gsr_read_counters); if it exceeds binning_capacity the kernels clamp and counters[3] is set.
int gsr_rasterize_forward_async(char* geometry_buffer, char* binning_buffer, long long binning_capacity, int global_sort,
                                char* image_buffer, int P, int D, int M, const float* background, int width, int height,
                                const float* means3D, const float* shs, const float* colors_precomp, const float* opacities,
                                const float* scales, float scale_modifier, const float* rotations, const float* cov3D_precomp,
                                const float* viewmatrix, const float* projmatrix, const float* cam_pos, float tan_fovx,
                                float tan_fovy, float* out_color, float* out_depth, float* out_alpha, int* radii, int* n_touched,
                                const float* cull_records, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (P <= 0 || width <= 0 || height <= 0 || binning_capacity <= 0 || binning_capacity >= (1ll << 30))
    return fail(GSR_ERR_INVALID_ARGUMENT, "bad sizes");
  if (!geometry_buffer || !binning_buffer || !image_buffer || !background || !means3D || !opacities || !viewmatrix || !projmatrix ||
      !cam_pos || !out_color || !out_depth || !out_alpha || !radii || (!shs && !colors_precomp) ||
      (!cov3D_precomp && (!scales || !rotations)))
    return fail(GSR_ERR_INVALID_ARGUMENT, "null required pointer");
  const uint32_t gx = (width + TILE_X - 1) / TILE_X, gy = (height + TILE_Y - 1) / TILE_Y;
  const int T = (int)(gx * gy);
  GeometryView g;
  carve_geometry(geometry_buffer, P, g);
  ImageView im;
  carve_image(image_buffer, width, height, im);
  BinningView bl;
  carve_binning(binning_buffer, binning_capacity, bl, global_sort != 0, width, height, P);
  // counters and the coverage grid are zeroed by the first preprocess kernel
  PreprocessParams pp{};
  pp.P = P, pp.D = D, pp.M = M, pp.W = width, pp.H = height, pp.grid_x = gx, pp.grid_y = gy;
  pp.means3D = means3D, pp.scales = scales, pp.rotations = rotations, pp.opacities = opacities, pp.shs = shs;
  pp.cov3D_precomp = cov3D_precomp, pp.colors_precomp = colors_precomp;
  pp.viewmatrix = viewmatrix, pp.projmatrix = projmatrix, pp.campos = cam_pos;
  pp.scale_modifier = scale_modifier, pp.tan_fovx = tan_fovx, pp.tan_fovy = tan_fovy;
  pp.focal_y = height / (2.0f * tan_fovy), pp.focal_x = width / (2.0f * tan_fovx);
  pp.prefiltered = 0;
  pp.cull_rec = cov3D_precomp ? nullptr : reinterpret_cast<const float4*>(cull_records);
  pp.sh_vec4 = shs && (M % 4 == 0) && ((uintptr_t)shs % 16 == 0);
  pp.radii = radii, pp.n_touched = n_touched, pp.tile_diff = im.tile_diff, pp.geom = g;
  launch_preprocess_fwd(pp, stream);
  SideStream* ss = side_stream();
  if (ss) {
    GSR_CUDA(cudaEventRecord(ss->fork, stream));
    GSR_CUDA(cudaStreamWaitEvent(ss->stream, ss->fork, 0));
    launch_color_fwd(pp, ss->stream);
    GSR_CUDA(cudaEventRecord(ss->join, ss->stream));
  } else {
    launch_color_fwd(pp, stream);
  }
  launch_scan_tiles(im.tile_diff, gx, gy, im.ranges, im.tile_cursor, im.tile_order, g.counters, (uint32_t)binning_capacity, stream);
  BinHeader* hdr = reinterpret_cast<BinHeader*>(binning_buffer);
  const BinHeader hv{(unsigned long long)binning_capacity, bl.units_off, bl.ckpt_off, bl.rec_off, bl.units_cap};
  if (!global_sort) {
    launch_scatter(P, g, im.tile_cursor, bl.comp, gx, hv, hdr, stream);
    launch_tile_sort(T, im.ranges, bl.comp, bl.point_list, (uint32_t)binning_capacity, im.tile_order, stream);
  } else {
    launch_long_bin(P, T, gx, gy, im.ranges, g, bl, binning_capacity, hv, hdr, stream);
  }
  RenderParams rp{};
  rp.W = width, rp.H = height, rp.grid_x = gx, rp.grid_y = gy;
  rp.ranges = im.ranges, rp.point_list = bl.point_list, rp.tile_order = im.tile_order;
  rp.final_cd = im.final_cd, rp.units = bl.units, rp.ckpt = bl.ckpt, rp.rec = bl.rec, rp.units_cap = (uint32_t)bl.units_cap, rp.unit_count = g.counters + 5;
  rp.mean_tau = g.mean_tau, rp.conic_opacity = g.conic_opacity, rp.rgbd = g.rgbd, rp.gid = g.gid;
  rp.bg = background, rp.out_color = out_color, rp.out_depth = out_depth, rp.out_alpha = out_alpha;
  rp.n_contrib = im.n_contrib, rp.n_touched = n_touched, rp.capacity = (uint32_t)binning_capacity;
  if (ss) GSR_CUDA(cudaStreamWaitEvent(stream, ss->join, 0));
  launch_render_fwd(rp, stream);
  GSR_STAGE("forward_async", 0, stream);
  return GSR_OK;
}

// counters of the last forward run on this geometry buffer: out[0] = num_rendered, out[1] = overflow flag,
// out[2] = longest tile list.  Synchronises the stream.
int gsr_read_counters(const char* geometry_buffer, int P, unsigned int* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!geometry_buffer || P <= 0 || !out) return fail(GSR_ERR_INVALID_ARGUMENT, "bad arguments");
  GeometryView g;
  carve_geometry(const_cast<char*>(geometry_buffer), P, g);
  uint32_t h[32];
  GSR_CUDA(cudaMemcpyAsync(h, g.counters, sizeof(h), cudaMemcpyDeviceToHost, stream));
  GSR_CUDA(cudaStreamSynchronize(stream));
  out[0] = h[1], out[1] = (h[3] | h[16]) ? 1u : 0u, out[2] = h[4];
  return GSR_OK;
}

// The overflow flag returned by gsr_read_counters is sticky across forwards on the same geometry buffer (a CUDA graph
// replays many forwards between two reads); this clears it.  Call once after allocating the buffer and per query.
int gsr_clear_overflow(char* geometry_buffer, int P, void* stream_) {
  if (!geometry_buffer || P <= 0) return fail(GSR_ERR_INVALID_ARGUMENT, "bad arguments");
  GeometryView g;
  carve_geometry(geometry_buffer, P, g);
  GSR_CUDA(cudaMemsetAsync(g.counters + 16, 0, 16 * sizeof(uint32_t), (cudaStream_t)stream_));
  return GSR_OK;
}

int gsr_rasterize_backward(int P, int D, int M, long long R, const float* background, int width, int height,
                           const float* means3D, const float* shs, const float* colors_precomp, const float* out_alpha,
                           const float* scales, float scale_modifier, const float* rotations, const float* cov3D_precomp,
                           const float* viewmatrix, const float* projmatrix, const float* projmatrix_raw, const float* cam_pos,
                           float tan_fovx, float tan_fovy, const int* radii, char* geometry_buffer, char* binning_buffer,
                           char* image_buffer, const float* dL_dpix, const float* dL_ddepth, const float* dL_dalpha,
                           float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor, float* dL_dmean3D,
                           float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot, float* dL_dtau, int debug,
                           void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (P < 0 || width <= 0 || height <= 0 || R < 0) return fail(GSR_ERR_INVALID_ARGUMENT, "bad sizes");
  if (dL_dtau) GSR_CUDA(cudaMemsetAsync(dL_dtau, 0, 6 * sizeof(float), stream));
  if (P == 0) return GSR_OK;   // reference rasterize_points.cu:168
  if (!geometry_buffer || !image_buffer || !binning_buffer) return fail(GSR_ERR_INVALID_ARGUMENT, "scratch buffers are required");
  if (!background || !means3D || !out_alpha || !viewmatrix || !projmatrix || !cam_pos || !radii || !dL_dpix)   // dL_ddepth / dL_dalpha may be NULL: no upstream gradient on that output
    return fail(GSR_ERR_INVALID_ARGUMENT, "null required pointer");
  if (dL_dtau && !projmatrix_raw) return fail(GSR_ERR_INVALID_ARGUMENT, "projmatrix_raw is required for the pose gradient");
  (void)colors_precomp;
  const uint32_t gx = (width + TILE_X - 1) / TILE_X, gy = (height + TILE_Y - 1) / TILE_Y;

  GeometryView g;
  carve_geometry(geometry_buffer, P, g);
  ImageView im;
  carve_image(image_buffer, width, height, im);
  BinningView bl;
  carve_binning(binning_buffer, R, bl);

  // Dense outputs: zero rows for culled Gaussians (the reference's nine torch::zeros, rasterize_points.cu:158-166).
  // Outputs that sit next to each other in memory (a caller carving them from one arena; gaps of alignment padding
  // below 128 bytes are zeroed with them — documented in include/gsr_b200.h) are merged.  When the blend backward runs, it writes the zeros itself between its work units — a memset
  // on a side stream cannot overlap with it, the persistent blend CTAs leave it no SM slots (measured: the fills
  // cost 45 us of a 0.49 ms step that way).  Otherwise (nothing rendered, unaligned spans) plain memsets.
  struct Span { char* lo; char* hi; } spans[9];
  int ns = 0;
  {
    const size_t Pz = (size_t)P;
    struct { float* p; size_t n; } fills[] = {{dL_dmean2D, 3 * Pz}, {dL_dconic, 4 * Pz}, {dL_dopacity, Pz}, {dL_dcolor, 3 * Pz},
                                             {dL_dmean3D, 3 * Pz}, {dL_dcov3D, 6 * Pz}, {dL_dsh, 3 * (size_t)M * Pz},
                                             {dL_dscale, 3 * Pz}, {dL_drot, 4 * Pz}};
    Span raw[9];
    int nr = 0;
    for (auto& f : fills)
      if (f.p && f.n) raw[nr++] = Span{(char*)f.p, (char*)f.p + sizeof(float) * f.n};
    std::sort(raw, raw + nr, [](const Span& a, const Span& b) { return a.lo < b.lo; });
    for (int i = 0; i < nr;) {
      Span m = raw[i];
      int j = i + 1;
      while (j < nr && raw[j].lo >= m.hi && raw[j].lo - m.hi < 128) m.hi = raw[j++].hi;   // contract: include/gsr_b200.h
      spans[ns++] = m;
      i = j;
    }
  }
  FillSpans fused{};
#ifdef GSR_AB_FILL_MEMSET          // measurement build: zero rows by cudaMemsetAsync
  bool fuse_fill = false;
#else
  bool fuse_fill = R > 0;
#endif
  for (int i = 0; i < ns && fuse_fill; i++) fuse_fill = ((uintptr_t)spans[i].lo % 16 == 0);
  if (fuse_fill) {
    for (int i = 0; i < ns; i++) {
      const size_t bytes = (size_t)(spans[i].hi - spans[i].lo);
      fused.base[i] = reinterpret_cast<float4*>(spans[i].lo);
      fused.n4[i] = bytes / 16;
      if (bytes % 16) GSR_CUDA(cudaMemsetAsync(spans[i].lo + bytes / 16 * 16, 0, bytes % 16, stream));   // at most 12 bytes
    }
    fused.count = ns;
  } else {
    for (int i = 0; i < ns; i++) GSR_CUDA(cudaMemsetAsync(spans[i].lo, 0, (size_t)(spans[i].hi - spans[i].lo), stream));
  }

  // the accumulator rows of all visible slots are zero here: the forward's scatter kernel zeroed
  // them and every preprocess-backward leaves them zero again
  if (R > 0) {
    StageScope ts(ST_BWD_RENDER, stream);
    RenderBwdParams rb{};
    rb.fills = fused;
    rb.W = width, rb.H = height, rb.grid_x = gx, rb.grid_y = gy;
    rb.binning_base = binning_buffer, rb.unit_count = g.counters + 5, rb.queue = g.counters + 6, rb.final_cd = im.final_cd;
    rb.max_units = (uint32_t)(4 * BSEG_PER_SEG * max_units(R, (size_t)gx * gy));
    rb.bg = background, rb.out_alpha = out_alpha, rb.n_contrib = im.n_contrib;
    rb.dL_dpix = dL_dpix, rb.dL_ddepth = dL_ddepth, rb.dL_dalpha = dL_dalpha, rb.grad_acc = g.grad_acc;
    launch_render_bwd(rb, stream);
  }
  GSR_STAGE("render_backward", debug, stream);

  PreBwdParams pb{};
  pb.P = P, pb.D = D, pb.M = M, pb.W = width, pb.H = height;
  pb.means3D = means3D, pb.radii = radii, pb.shs = shs, pb.scales = scales, pb.rotations = rotations;
  pb.cov3D_precomp = cov3D_precomp, pb.viewmatrix = viewmatrix, pb.projmatrix = projmatrix;
  pb.projmatrix_raw = projmatrix_raw, pb.campos = cam_pos;
  pb.scale_modifier = scale_modifier, pb.tan_fovx = tan_fovx, pb.tan_fovy = tan_fovy;
  pb.focal_y = height / (2.0f * tan_fovy), pb.focal_x = width / (2.0f * tan_fovx);
  pb.geom = g;
  pb.dL_dmean2D = dL_dmean2D, pb.dL_dconic = dL_dconic, pb.dL_dopacity = dL_dopacity, pb.dL_dcolor = dL_dcolor;
  pb.dL_dmean3D = dL_dmean3D, pb.dL_dcov3D = dL_dcov3D, pb.dL_dsh = dL_dsh, pb.dL_dscale = dL_dscale, pb.dL_drot = dL_drot;
  pb.dL_dtau = dL_dtau;
  {
    StageScope ts(ST_BWD_PREPROCESS, stream);
    if (R > 0) launch_preprocess_bwd(pb, stream);
  }
  GSR_STAGE("preprocess_backward", debug, stream);
  stage_collect(stream);
  return GSR_OK;
}

int gsr_build_cull_records(int P, const float* means3D, const float* scales, const float* rotations, float* records, void* stream_) {
  if (P < 0) return fail(GSR_ERR_INVALID_ARGUMENT, "P < 0");
  if (P == 0) return GSR_OK;
  if (!means3D || !scales || !rotations || !records) return fail(GSR_ERR_INVALID_ARGUMENT, "null pointer");
  if ((uintptr_t)records % 16 || (uintptr_t)rotations % 16) return fail(GSR_ERR_INVALID_ARGUMENT, "records and rotations must be 16-byte aligned");
  launch_build_cull_records(P, means3D, scales, rotations, reinterpret_cast<float4*>(records), (cudaStream_t)stream_);
  GSR_STAGE("build_cull_records", 0, (cudaStream_t)stream_);
  return GSR_OK;
}

int gsr_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix, unsigned char* present,
                     void* stream_) {
  (void)projmatrix;   // the reference's test only uses the view matrix (auxiliary.h:154)
  if (P < 0) return fail(GSR_ERR_INVALID_ARGUMENT, "P < 0");
  if (P == 0) return GSR_OK;
  if (!means3D || !viewmatrix || !present) return fail(GSR_ERR_INVALID_ARGUMENT, "null pointer");
  launch_mark_visible(P, means3D, viewmatrix, present, (cudaStream_t)stream_);
  GSR_STAGE("mark_visible", 0, (cudaStream_t)stream_);
  return GSR_OK;
}

}  // extern "C"

// ----------------------------------------------------------------------------- state export (tests)
namespace {
// scatter the per-slot records back to the reference's per-Gaussian arrays (outputs pre-zeroed)
__global__ void export_geometry_kernel(GeometryView g, float* depths, float* means2D, float* cov3D, float* conic_opacity,
                                       float* rgb, unsigned char* clamped, uint32_t* tiles_touched) {
  if (threadIdx.x >= g.block_vis[blockIdx.x]) return;
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t i = g.gid[k];
  if (depths) depths[i] = g.depths[k];
  if (means2D) means2D[2 * i] = g.mean_tau[k].x, means2D[2 * i + 1] = g.mean_tau[k].y;
  if (cov3D)
    for (int q = 0; q < 6; q++) cov3D[6 * i + q] = g.cov3D[6 * (size_t)k + q];
  if (conic_opacity) {
    const float4 c = g.conic_opacity[k];
    conic_opacity[4 * i] = c.x, conic_opacity[4 * i + 1] = c.y, conic_opacity[4 * i + 2] = c.z, conic_opacity[4 * i + 3] = c.w;
  }
  if (rgb) {
    const float4 c = g.rgbd[k];
    rgb[3 * i] = c.x, rgb[3 * i + 1] = c.y, rgb[3 * i + 2] = c.z;
  }
  if (clamped) {
    const uint8_t m = g.clamped[k];
    clamped[3 * i] = m & 1, clamped[3 * i + 1] = (m >> 1) & 1, clamped[3 * i + 2] = (m >> 2) & 1;
  }
  if (tiles_touched) {
    const uint2 rc = g.rect[k];
    tiles_touched[i] = ((rc.x >> 16) - (rc.x & 0xffffu)) * ((rc.y >> 16) - (rc.y & 0xffffu));
  }
}
// the reference's sorted arrays: keys = tile << 32 | depth bits, point_list = Gaussian ids
__global__ void export_sorted_kernel(const uint2* ranges, const uint32_t* point_list, const float* depths, const uint32_t* gid,
                                     uint64_t* keys, uint32_t* list) {
  const uint2 rg = ranges[blockIdx.x];
  for (uint32_t i = rg.x + threadIdx.x; i < rg.y; i += blockDim.x) {
    const uint32_t slot = point_list[i];
    if (keys) keys[i] = ((uint64_t)blockIdx.x << 32) | __float_as_uint(depths[slot]);
    if (list) list[i] = gid[slot];
  }
}
}  // namespace

extern "C" {

int gsr_export_state(int P, long long R, int width, int height, const char* geometry_buffer, const char* binning_buffer,
                     const char* image_buffer, float* depths, float* means2D, float* cov3D, float* conic_opacity, float* rgb,
                     unsigned char* clamped, uint32_t* tiles_touched, uint64_t* keys, uint32_t* list, uint32_t* ranges,
                     uint32_t* n_contrib, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const uint32_t gx = (width + TILE_X - 1) / TILE_X, gy = (height + TILE_Y - 1) / TILE_Y;
  GeometryView g{};
  ImageView im{};
  if (image_buffer) carve_image(const_cast<char*>(image_buffer), width, height, im);
  if (P > 0 && geometry_buffer) {
    carve_geometry(const_cast<char*>(geometry_buffer), P, g);
    export_geometry_kernel<<<num_pre_blocks(P), PRE_THREADS, 0, stream>>>(g, depths, means2D, cov3D, conic_opacity, rgb, clamped,
                                                                       tiles_touched);
  }
  if (R > 0 && binning_buffer && image_buffer && P > 0 && geometry_buffer && (keys || list)) {
    BinningView bl;
    unsigned long long cap = 0;
    GSR_CUDA(cudaStreamSynchronize(stream));
    GSR_CUDA(cudaMemcpy(&cap, binning_buffer, sizeof(cap), cudaMemcpyDeviceToHost));
    carve_binning(const_cast<char*>(binning_buffer), (long long)cap, bl);
    export_sorted_kernel<<<gx * gy, 256, 0, stream>>>(im.ranges, bl.point_list, g.depths, g.gid, keys, list);
  }
  if (image_buffer) {
    if (ranges) GSR_CUDA(cudaMemcpyAsync(ranges, im.ranges, sizeof(uint2) * (size_t)gx * gy, cudaMemcpyDeviceToDevice, stream));
    if (n_contrib) GSR_CUDA(cudaMemcpyAsync(n_contrib, im.n_contrib, sizeof(uint32_t) * (size_t)width * height, cudaMemcpyDeviceToDevice, stream));
  }
  GSR_STAGE("export_state", 1, stream);
  return GSR_OK;
}

void gsr_stage_timing(int enable) {
  std::lock_guard<std::mutex> lk(g_timer.mu);
  g_timer.enabled = enable != 0;
  for (int i = 0; i < ST_COUNT; i++) g_timer.total_ms[i] = 0.0, g_timer.calls[i] = 0, g_timer.used[i] = false;
}
int gsr_stage_times(double* total_ms, unsigned long long* calls, int n) {
  std::lock_guard<std::mutex> lk(g_timer.mu);
  for (int i = 0; i < n && i < ST_COUNT; i++) {
    if (total_ms) total_ms[i] = g_timer.total_ms[i];
    if (calls) calls[i] = g_timer.calls[i];
  }
  return ST_COUNT;
}

int gsr_l1_loss_grad(const float* image, const float* target, float* dL_dimage, long long n, float weight, float* loss_accum,
                     void* stream_) {
  if (n < 0 || (n > 0 && (!image || !target || !dL_dimage || !loss_accum))) return fail(GSR_ERR_INVALID_ARGUMENT, "bad arguments");
  launch_l1_loss_grad(image, target, dL_dimage, (size_t)n, weight, loss_accum, (cudaStream_t)stream_);
  GSR_STAGE("l1_loss_grad", 0, (cudaStream_t)stream_);
  return GSR_OK;
}

int gsr_l1_ssim_loss_grad(const float* image, const float* target, int channels, int height, int width, float lambda_dssim,
                          float* loss_accum, float* dL_dimage, float* scratch, void* stream_) {
  if (channels <= 0 || height <= 0 || width <= 0 || !image || !target || !loss_accum || !dL_dimage || !scratch)
    return fail(GSR_ERR_INVALID_ARGUMENT, "bad arguments");
  launch_l1_ssim_loss_grad(image, target, channels, height, width, lambda_dssim, loss_accum, dL_dimage, scratch, (cudaStream_t)stream_);
  GSR_STAGE("l1_ssim_loss_grad", 0, (cudaStream_t)stream_);
  return GSR_OK;
}

int gsr_map_adam_step(int P, int M, int do_stats, int do_adam, const int* steps, const float* lrs, float beta1, float beta2, float eps,
                      float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                      float* opacity_act, float* scaling_act, float* rotation_act, const float* dL_dmeans2D, const int* radii,
                      float* max_radii2D, float* xyz_gradient_accum, float* denom, void* stream_) {
  if (P < 0 || M < 1) return fail(GSR_ERR_INVALID_ARGUMENT, "bad P/M");
  if (P == 0) return GSR_OK;
  MapStepParams s{};
  s.P = P, s.M = M, s.do_stats = do_stats, s.do_adam = do_adam;
  if (do_stats && (!dL_dmeans2D || !radii || !max_radii2D || !xyz_gradient_accum || !denom))
    return fail(GSR_ERR_INVALID_ARGUMENT, "statistics need dL_dmeans2D, radii, max_radii2D, xyz_gradient_accum, denom");
  s.g_means2D = dL_dmeans2D, s.radii = radii, s.max_radii2D = max_radii2D, s.xyz_gradient_accum = xyz_gradient_accum, s.denom = denom;
  if (do_adam) {
    if (!steps || !lrs || !params || !grads || !exp_avg || !exp_avg_sq) return fail(GSR_ERR_INVALID_ARGUMENT, "adam needs steps, lrs and the tensor tables");
    for (int i = 0; i < 6; i++)
      if (lrs[i] >= 0.f && steps[i] < 1) return fail(GSR_ERR_INVALID_ARGUMENT, "steps[%d] must be >= 1", i);
    float lr[6];
    for (int i = 0; i < 6; i++) lr[i] = lrs[i];
    // order: xyz, f_dc+f_rest (one tensor), opacity, scaling, rotation
    for (int t = 0; t < 5; t++) {
      const bool used = t == 1 ? (lr[1] >= 0.f || lr[2] >= 0.f) : lr[t == 0 ? 0 : t + 1] >= 0.f;
      if (used && (!params[t] || !grads[t] || !exp_avg[t] || !exp_avg_sq[t])) return fail(GSR_ERR_INVALID_ARGUMENT, "null tensor in group %d", t);
    }
    if ((lr[3] >= 0.f && !opacity_act) || (lr[4] >= 0.f && !scaling_act) || (lr[5] >= 0.f && !rotation_act))
      return fail(GSR_ERR_INVALID_ARGUMENT, "activation outputs missing");
    s.xyz = params[0], s.features = params[1], s.opacity = params[2], s.scaling = params[3], s.rotation = params[4];
    s.g_xyz = grads[0], s.g_features = grads[1], s.g_opacity = grads[2], s.g_scaling = grads[3], s.g_rotation = grads[4];
    s.m_xyz = exp_avg[0], s.m_features = exp_avg[1], s.m_opacity = exp_avg[2], s.m_scaling = exp_avg[3], s.m_rotation = exp_avg[4];
    s.v_xyz = exp_avg_sq[0], s.v_features = exp_avg_sq[1], s.v_opacity = exp_avg_sq[2], s.v_scaling = exp_avg_sq[3], s.v_rotation = exp_avg_sq[4];
    s.lr_xyz = lr[0], s.lr_f_dc = lr[1], s.lr_f_rest = lr[2], s.lr_opacity = lr[3], s.lr_scaling = lr[4], s.lr_rotation = lr[5];
    s.opacity_act = opacity_act, s.scaling_act = scaling_act, s.rotation_act = rotation_act;
  }
  launch_map_step(s, beta1, beta2, eps, steps, (cudaStream_t)stream_);
  GSR_STAGE("map_adam_step", 0, (cudaStream_t)stream_);
  return GSR_OK;
}

int gsr_pack_gradient_rows(const long long* row_ids, int n_rows, int padded_rows, int M, float* const* grads, float* table, void* stream_) {
  if (n_rows < 0 || padded_rows < n_rows || M < 1 || !grads || !table || (n_rows > 0 && !row_ids)) return fail(GSR_ERR_INVALID_ARGUMENT, "bad arguments");
  GradRowTensors t;
  for (int i = 0; i < 5; i++) {
    if (!grads[i]) return fail(GSR_ERR_INVALID_ARGUMENT, "null gradient tensor %d", i);
    t.g[i] = grads[i];
  }
  launch_pack_gradient_rows(row_ids, n_rows, padded_rows, M, t, table, (cudaStream_t)stream_);
  GSR_STAGE("pack_gradient_rows", 0, (cudaStream_t)stream_);
  return GSR_OK;
}

int gsr_pack_visible_rows(const int* radii, int P, int M, float* const* grads, const float* dL_dmeans2D, float* table, int capacity_rows,
                          unsigned int* count, void* stream_) {
  if (P < 0 || M < 1 || capacity_rows < 1 || !radii || !grads || !table || !count) return fail(GSR_ERR_INVALID_ARGUMENT, "bad arguments");
  GradRowTensors t;
  for (int i = 0; i < 5; i++) {
    if (!grads[i]) return fail(GSR_ERR_INVALID_ARGUMENT, "null gradient tensor %d", i);
    t.g[i] = grads[i];
  }
  launch_pack_visible_rows(radii, P, M, t, dL_dmeans2D, table, capacity_rows, count, (cudaStream_t)stream_);
  GSR_STAGE("pack_visible_rows", 0, (cudaStream_t)stream_);
  return GSR_OK;
}

int gsr_add_counted_rows(const float* table, int capacity_rows, int M, int P, float* const* grads, int add_gradients, float* max_radii2D,
                         float* xyz_gradient_accum, float* denom, void* stream_) {
  if (capacity_rows < 1 || M < 1 || P < 0 || !grads || !table) return fail(GSR_ERR_INVALID_ARGUMENT, "bad arguments");
  if (max_radii2D && (!xyz_gradient_accum || !denom)) return fail(GSR_ERR_INVALID_ARGUMENT, "statistics need all three arrays");
  GradRowTensors t;
  for (int i = 0; i < 5; i++) {
    if (!grads[i]) return fail(GSR_ERR_INVALID_ARGUMENT, "null gradient tensor %d", i);
    t.g[i] = grads[i];
  }
  launch_add_counted_rows(table, capacity_rows, M, P, t, add_gradients, max_radii2D, xyz_gradient_accum, denom, (cudaStream_t)stream_);
  GSR_STAGE("add_counted_rows", 0, (cudaStream_t)stream_);
  return GSR_OK;
}

int gsr_add_gradient_rows(const float* table, int padded_rows, int M, int P, float* const* grads, void* stream_) {
  if (padded_rows < 0 || M < 1 || P < 0 || !grads || !table) return fail(GSR_ERR_INVALID_ARGUMENT, "bad arguments");
  GradRowTensors t;
  for (int i = 0; i < 5; i++) {
    if (!grads[i]) return fail(GSR_ERR_INVALID_ARGUMENT, "null gradient tensor %d", i);
    t.g[i] = grads[i];
  }
  launch_add_gradient_rows(table, padded_rows, M, t, P, (cudaStream_t)stream_);
  GSR_STAGE("add_gradient_rows", 0, (cudaStream_t)stream_);
  return GSR_OK;
}

size_t gsr_knn_workspace_bytes(long long n_points) { return n_points > 0 ? knn_workspace_bytes(n_points) : 0; }

int gsr_dist2_knn3(const float* points, long long n_points, float* mean_dists, char* workspace, void* stream_) {
  if (n_points < 0 || n_points >= (1ll << 30)) return fail(GSR_ERR_INVALID_ARGUMENT, "n_points must be in [0, 2^30)");
  if (n_points == 0) return GSR_OK;
  if (!points || !mean_dists || !workspace) return fail(GSR_ERR_INVALID_ARGUMENT, "null pointer");
  launch_dist2_knn3(points, n_points, mean_dists, workspace, (cudaStream_t)stream_);
  GSR_STAGE("dist2_knn3", 0, (cudaStream_t)stream_);
  return GSR_OK;
}

int gsr_depth_loss_grad(const float* depth, const float* pseudo_depth, const float* gt_depth, long long n, float inv_numerator,
                        float pearson_weight, float l1_weight, float* loss_accum, float* dL_ddepth, double* scratch, void* stream_) {
  if (n <= 0 || n >= (1ll << 31) || !depth || !loss_accum || !dL_ddepth || !scratch) return fail(GSR_ERR_INVALID_ARGUMENT, "bad arguments");
  launch_depth_loss_grad(depth, pseudo_depth, gt_depth, (int)n, inv_numerator, pearson_weight, l1_weight, loss_accum, dL_ddepth,
                         scratch, (cudaStream_t)stream_);
  GSR_STAGE("depth_loss_grad", 0, (cudaStream_t)stream_);
  return GSR_OK;
}

int gsr_tracking_loss_grad(const float* image, const float* depth, const float* opacity, const float* gt_image, const float* gt_depth,
                           const float* grad_mask, const float* exposure, int height, int width, float opacity_threshold,
                           float depth_weight, float* loss_accum, float* dL_dimage, float* dL_ddepth, float* dL_dexposure,
                           void* stream_) {
  if (height <= 0 || width <= 0 || !image || !opacity || !gt_image || !loss_accum || !dL_dimage)
    return fail(GSR_ERR_INVALID_ARGUMENT, "bad arguments");
  if (gt_depth && (!depth || !dL_ddepth)) return fail(GSR_ERR_INVALID_ARGUMENT, "gt_depth needs depth and dL_ddepth");
  launch_tracking_loss_grad(image, depth, opacity, gt_image, gt_depth, grad_mask, exposure, height * width, opacity_threshold,
                            depth_weight, dL_dimage, dL_ddepth, loss_accum, dL_dexposure, (cudaStream_t)stream_);
  GSR_STAGE("tracking_loss_grad", 0, (cudaStream_t)stream_);
  return GSR_OK;
}

int gsr_exposure_adam_step(float* exposure, float* dL_dexposure, float* adam_m, float* adam_v, float* step_count, float lr,
                           void* stream_) {
  if (!exposure || !dL_dexposure || !adam_m || !adam_v || !step_count) return fail(GSR_ERR_INVALID_ARGUMENT, "null pointer");
  launch_exposure_adam_step(exposure, dL_dexposure, adam_m, adam_v, step_count, lr, (cudaStream_t)stream_);
  GSR_STAGE("exposure_adam_step", 0, (cudaStream_t)stream_);
  return GSR_OK;
}

int gsr_pose_adam_step(const float* dL_dtau, float* adam_m, float* adam_v, float* step_count, float lr_trans, float lr_rot,
                       float* w2c, const float* projmatrix_raw, float* viewmatrix, float* projmatrix, float* campos,
                       float* tau_norm, void* stream_) {
  if (!dL_dtau || !adam_m || !adam_v || !step_count || !w2c || !projmatrix_raw || !viewmatrix || !projmatrix || !campos)
    return fail(GSR_ERR_INVALID_ARGUMENT, "null pointer");
  launch_pose_adam_step(dL_dtau, adam_m, adam_v, step_count, lr_trans, lr_rot, w2c, projmatrix_raw, viewmatrix, projmatrix, campos,
                        tau_norm, (cudaStream_t)stream_);
  GSR_STAGE("pose_adam_step", 0, (cudaStream_t)stream_);
  return GSR_OK;
}

size_t gsr_sort_temp_bytes(long long n) { return sort_temp_bytes(n, SORT_MAX_PASSES); }

int gsr_sort_pairs(const uint64_t* keys_in, uint64_t* keys_out, const uint32_t* vals_in, uint32_t* vals_out, uint64_t* keys_tmp,
                   uint32_t* vals_tmp, long long n, int end_bit, char* temp, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || end_bit < 1 || end_bit > 64) return fail(GSR_ERR_INVALID_ARGUMENT, "bad n/end_bit");
  if (n == 0) return GSR_OK;
  if (n >= (1ll << 30)) return fail(GSR_ERR_INVALID_ARGUMENT, "n must be < 2^30");
  const int passes = sort_passes(end_bit);
  SortTemp st;
  carve_sort_temp(temp, n, passes, st);
  sort_temp_reset(temp, n, passes, stream);
  launch_sort_histogram(keys_in, nullptr, n, 0, end_bit, st.hist, stream);
  // arrange the ping-pong so that the last pass lands in keys_out: odd passes in->out directly
  uint64_t* kb[2];
  uint32_t* vb[2];
  if (passes & 1) {
    // pass 0 must read keys_in: copy-free only if we may treat keys_in as buffer 0; it is const, so stage through tmp
    GSR_CUDA(cudaMemcpyAsync(keys_tmp, keys_in, sizeof(uint64_t) * (size_t)n, cudaMemcpyDeviceToDevice, stream));
    GSR_CUDA(cudaMemcpyAsync(vals_tmp, vals_in, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToDevice, stream));
    kb[0] = keys_tmp, kb[1] = keys_out, vb[0] = vals_tmp, vb[1] = vals_out;
  } else {
    GSR_CUDA(cudaMemcpyAsync(keys_out, keys_in, sizeof(uint64_t) * (size_t)n, cudaMemcpyDeviceToDevice, stream));
    GSR_CUDA(cudaMemcpyAsync(vals_out, vals_in, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToDevice, stream));
    kb[0] = keys_out, kb[1] = keys_tmp, vb[0] = vals_out, vb[1] = vals_tmp;
  }
  launch_onesweep(kb, vb, nullptr, n, end_bit, st, stream);
  GSR_STAGE("sort_pairs", 0, stream);
  return GSR_OK;
}

}  // extern "C"

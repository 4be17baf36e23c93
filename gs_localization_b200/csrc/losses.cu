// Fused photometric loss of LoGS map training, forward and backward in two passes over the image:
//   L = (1 - lambda) * mean|x - y| + lambda * (1 - mean(SSIM(x, y)))
// (gs_localization/gs/7scenes_gs_full_dslam.py:165-166 with gaussian_splatting/utils/loss_utils.py:17-64:
// 11x11 Gaussian window, sigma 1.5, zero padding, C1 = 0.01^2, C2 = 0.03^2, per-channel).
//
// The reference evaluates this with five grouped conv2d calls plus ~20 element-wise kernels, and as many
// again in autograd's backward.  Here pass 1 computes the five windowed moments with a separable filter in
// shared memory, the SSIM map, the loss sums and the three derivative maps (d ssim / d mu1, d E[xx], d E[xy],
// pre-multiplied by dL/d ssim); pass 2 filters those maps with the same (symmetric) window and adds the L1 term,
// yielding dL/dx directly — which is exactly the dL_dpix the rasterizer's backward consumes.
#include <algorithm>

#include "gsr_kernels.cuh"

namespace gsr {

constexpr int SSIM_R = 5;                 // window radius (11 taps)
constexpr int ST = 16;                    // output tile edge
constexpr int SH_ = ST + 2 * SSIM_R;      // tile edge with halo

struct SsimWindow { float w[2 * SSIM_R + 1]; };

__device__ __forceinline__ float load_or_zero(const float* __restrict__ img, int x, int y, int W, int H) {
  return (x >= 0 && x < W && y >= 0 && y < H) ? __ldg(img + (size_t)y * W + x) : 0.f;
}

__global__ void __launch_bounds__(ST * ST) ssim_fwd_kernel(const float* __restrict__ img1, const float* __restrict__ img2, int W, int H,
                                                          int C, SsimWindow win, float g_ssim, float* __restrict__ maps,
                                                          float* __restrict__ sums /* [0]=sum ssim, [1]=sum |x-y| */) {
  __shared__ float s1[SH_][SH_ + 1], s2[SH_][SH_ + 1];
  __shared__ float h[5][SH_][ST + 1];
  __shared__ float s_red[2][ST * ST / 32];
  const int c = blockIdx.z;
  const size_t plane = (size_t)W * H;
  const float* a = img1 + c * plane;
  const float* b = img2 + c * plane;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * ST + tx;
  const int x0 = blockIdx.x * ST - SSIM_R, y0 = blockIdx.y * ST - SSIM_R;
  for (int i = tid; i < SH_ * SH_; i += ST * ST) {
    const int ly = i / SH_, lx = i - ly * SH_;
    s1[ly][lx] = load_or_zero(a, x0 + lx, y0 + ly, W, H);
    s2[ly][lx] = load_or_zero(b, x0 + lx, y0 + ly, W, H);
  }
  __syncthreads();
  // horizontal pass: SH_ rows x ST columns
  for (int i = tid; i < SH_ * ST; i += ST * ST) {
    const int ly = i / ST, lx = i - ly * ST;
    float m1 = 0, m2 = 0, e11 = 0, e22 = 0, e12 = 0;
#pragma unroll
    for (int k = 0; k <= 2 * SSIM_R; k++) {
      const float w = win.w[k], u = s1[ly][lx + k], v = s2[ly][lx + k];
      m1 += w * u; m2 += w * v; e11 += w * u * u; e22 += w * v * v; e12 += w * u * v;
    }
    h[0][ly][lx] = m1; h[1][ly][lx] = m2; h[2][ly][lx] = e11; h[3][ly][lx] = e22; h[4][ly][lx] = e12;
  }
  __syncthreads();
  const int x = blockIdx.x * ST + tx, y = blockIdx.y * ST + ty;
  float ssim_v = 0.f, l1_v = 0.f;
  if (x < W && y < H) {
    float mu1 = 0, mu2 = 0, e11 = 0, e22 = 0, e12 = 0;
#pragma unroll
    for (int k = 0; k <= 2 * SSIM_R; k++) {
      const float w = win.w[k];
      mu1 += w * h[0][ty + k][tx]; mu2 += w * h[1][ty + k][tx];
      e11 += w * h[2][ty + k][tx]; e22 += w * h[3][ty + k][tx]; e12 += w * h[4][ty + k][tx];
    }
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
    const float sig1 = e11 - mu1_sq, sig2 = e22 - mu2_sq, sig12 = e12 - mu12;
    const float A1 = 2.f * mu12 + C1, A2 = 2.f * sig12 + C2, B1 = mu1_sq + mu2_sq + C1, B2 = sig1 + sig2 + C2;
    const float inv = 1.f / (B1 * B2);
    ssim_v = A1 * A2 * inv;
    // partial derivatives w.r.t. the windowed moments of img1 (E[xx], E[xy] held fixed for d/dmu1)
    const float d_mu1 = 2.f * mu2 * (A2 - A1) * inv - ssim_v * (2.f * mu1 / B1 - 2.f * mu1 / B2);
    const float d_e11 = -ssim_v / B2;
    const float d_e12 = 2.f * A1 * inv;
    const size_t p = c * plane + (size_t)y * W + x;
    const size_t vol = plane * C;
    maps[p] = g_ssim * d_mu1;
    maps[vol + p] = g_ssim * d_e11;
    maps[2 * vol + p] = g_ssim * d_e12;
    l1_v = fabsf(s1[ty + SSIM_R][tx + SSIM_R] - s2[ty + SSIM_R][tx + SSIM_R]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ssim_v += __shfl_xor_sync(0xffffffffu, ssim_v, o);
    l1_v += __shfl_xor_sync(0xffffffffu, l1_v, o);
  }
  if ((tid & 31) == 0) s_red[0][tid >> 5] = ssim_v, s_red[1][tid >> 5] = l1_v;
  __syncthreads();
  if (tid == 0) {
    float s = 0, l = 0;
    for (int w = 0; w < ST * ST / 32; w++) s += s_red[0][w], l += s_red[1][w];
    atomicAdd(sums + 0, s);
    atomicAdd(sums + 1, l);
  }
}

__global__ void __launch_bounds__(ST * ST) ssim_bwd_kernel(const float* __restrict__ img1, const float* __restrict__ img2, int W, int H,
                                                          int C, SsimWindow win, float g_l1, const float* __restrict__ maps,
                                                          float* __restrict__ dL_dimg1) {
  __shared__ float sm[3][SH_][SH_ + 1];
  __shared__ float h[3][SH_][ST + 1];
  const int c = blockIdx.z;
  const size_t plane = (size_t)W * H, vol = plane * C;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * ST + tx;
  const int x0 = blockIdx.x * ST - SSIM_R, y0 = blockIdx.y * ST - SSIM_R;
  for (int i = tid; i < SH_ * SH_; i += ST * ST) {
    const int ly = i / SH_, lx = i - ly * SH_;
#pragma unroll
    for (int m = 0; m < 3; m++) sm[m][ly][lx] = load_or_zero(maps + m * vol + c * plane, x0 + lx, y0 + ly, W, H);
  }
  __syncthreads();
  for (int i = tid; i < SH_ * ST; i += ST * ST) {
    const int ly = i / ST, lx = i - ly * ST;
    float a0 = 0, a1 = 0, a2 = 0;
#pragma unroll
    for (int k = 0; k <= 2 * SSIM_R; k++) {
      const float w = win.w[k];
      a0 += w * sm[0][ly][lx + k]; a1 += w * sm[1][ly][lx + k]; a2 += w * sm[2][ly][lx + k];
    }
    h[0][ly][lx] = a0; h[1][ly][lx] = a1; h[2][ly][lx] = a2;
  }
  __syncthreads();
  const int x = blockIdx.x * ST + tx, y = blockIdx.y * ST + ty;
  if (x < W && y < H) {
    float a0 = 0, a1 = 0, a2 = 0;
#pragma unroll
    for (int k = 0; k <= 2 * SSIM_R; k++) {
      const float w = win.w[k];
      a0 += w * h[0][ty + k][tx]; a1 += w * h[1][ty + k][tx]; a2 += w * h[2][ty + k][tx];
    }
    const size_t p = c * plane + (size_t)y * W + x;
    const float u = __ldg(img1 + p), v = __ldg(img2 + p);
    const float d = u - v;
    dL_dimg1[p] = a0 + 2.f * u * a1 + v * a2 + (d > 0.f ? g_l1 : (d < 0.f ? -g_l1 : 0.f));
  }
}

// finishes the loss scalar: loss += (1-lambda) * sum|x-y| / n + lambda * (1 - sum ssim / n)
__global__ void ssim_finish_kernel(const float* __restrict__ sums, float n, float lambda, float* __restrict__ loss) {
  if (threadIdx.x == 0 && blockIdx.x == 0) loss[0] += (1.f - lambda) * sums[1] / n + lambda * (1.f - sums[0] / n);
}

void launch_l1_ssim_loss_grad(const float* img1, const float* img2, int C, int H, int W, float lambda, float* loss, float* dL_dimg1,
                              float* scratch, cudaStream_t stream) {
  SsimWindow win;
  {   // gaussian(11, 1.5) normalised, in float32 like loss_utils.py:23-25
    float g[2 * SSIM_R + 1], s = 0.f;
    for (int i = 0; i <= 2 * SSIM_R; i++) { g[i] = (float)exp(-(double)((i - SSIM_R) * (i - SSIM_R)) / (2.0 * 1.5 * 1.5)); s += g[i]; }
    for (int i = 0; i <= 2 * SSIM_R; i++) win.w[i] = g[i] / s;
  }
  const size_t vol = (size_t)C * H * W;
  float* maps = scratch;              // 3 * vol
  float* sums = scratch + 3 * vol;    // 2
  cudaMemsetAsync(sums, 0, 2 * sizeof(float), stream);
  const dim3 grid((W + ST - 1) / ST, (H + ST - 1) / ST, C), block(ST, ST);
  const float n = (float)vol;
  ssim_fwd_kernel<<<grid, block, 0, stream>>>(img1, img2, W, H, C, win, -lambda / n, maps, sums);
  ssim_bwd_kernel<<<grid, block, 0, stream>>>(img1, img2, W, H, C, win, (1.f - lambda) / n, maps, dL_dimg1);
  ssim_finish_kernel<<<1, 32, 0, stream>>>(sums, n, lambda, loss);
  count_launch(3);
}

// ---------------------------------------------------------------------------------------------------------------
// Depth terms of the map-training loss (gs_localization/gs/7scenes_gs_full_dslam.py:168-184):
//   pseudo = min(1 - r(-m, d), 1 - r(k / (m + 200), d))      r = Pearson correlation over all pixels, m = monocular
//                                                            (MiDaS) depth, d = rendered depth, k = 1000 (1 in train.py)
//   L = w_p * pseudo + w_l1 * mean|d * mask - gt * mask|,    mask = gt > 0
// r comes from torchmetrics.functional.pearson_corrcoef (third-party, not vendored, version unpinned): restated
// here from its definition r = S_ad / sqrt(S_aa S_dd) with centred sums accumulated in double.
// Pass 1 reduces the nine raw sums, pass 2 recomputes the two coefficients per thread and writes dL/dd.
__global__ void __launch_bounds__(256) depth_loss_sums_kernel(const float* __restrict__ depth, const float* __restrict__ pseudo,
                                                              const float* __restrict__ gt, int n, float k, double* __restrict__ sums) {
  // 0:Sd 1:Sdd 2:Sa 3:Saa 4:Sad 5:Sb 6:Sbb 7:Sbd 8:S|l1|
  double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double d = depth[i];
    if (pseudo) {
      const float m = __ldg(pseudo + i);
      const double a = -(double)m, b = (double)(k / (m + 200.f));
      acc[0] += d; acc[1] += d * d; acc[2] += a; acc[3] += a * a; acc[4] += a * d; acc[5] += b; acc[6] += b * b; acc[7] += b * d;
    }
    if (gt) {
      const float g = __ldg(gt + i);
      if (g > 0.f) acc[8] += fabs(d - (double)g);
    }
  }
  __shared__ double s_part[9][8];
#pragma unroll
  for (int q = 0; q < 9; q++) {
    double v = acc[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_part[q][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    double t = 0;
    for (int w = 0; w < 8; w++) t += s_part[threadIdx.x][w];
    atomicAdd(sums + threadIdx.x, t);
  }
}

__global__ void __launch_bounds__(256) depth_loss_grad_kernel(const float* __restrict__ depth, const float* __restrict__ pseudo,
                                                              const float* __restrict__ gt, int n, float k, float w_pearson, float w_l1,
                                                              const double* __restrict__ sums, float* __restrict__ dL_ddepth,
                                                              float* __restrict__ loss) {
  const double N = (double)n;
  double mean_d = 0, mean_x = 0, c_x = 0, c_d = 0;   // selected regressor x: dL/dd_i = c_x (x_i - mean_x) + c_d (d_i - mean_d)
  int pick = 0;
  double pseudo_loss = 0;
  if (pseudo) {
    mean_d = sums[0] / N;
    const double Sdd = sums[1] - sums[0] * mean_d;
    const double Saa = sums[3] - sums[2] * sums[2] / N, Sad = sums[4] - sums[2] * mean_d;
    const double Sbb = sums[6] - sums[5] * sums[5] / N, Sbd = sums[7] - sums[5] * mean_d;
    const double ra = Sad / sqrt(Saa * Sdd), rb = Sbd / sqrt(Sbb * Sdd);
    pick = (1.0 - rb) < (1.0 - ra);                   // python min(): the second only if strictly smaller
    const double r = pick ? rb : ra, Sxx = pick ? Sbb : Saa;
    mean_x = (pick ? sums[5] : sums[2]) / N;
    pseudo_loss = 1.0 - r;
    c_x = -(double)w_pearson / sqrt(Sxx * Sdd);
    c_d = (double)w_pearson * r / Sdd;
  }
  const float g_l1 = w_l1 / (float)n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float d = depth[i];
    double g = 0;
    if (pseudo) {
      const float m = __ldg(pseudo + i);
      const double x = pick ? (double)(k / (m + 200.f)) : -(double)m;
      g = c_x * (x - mean_x) + c_d * ((double)d - mean_d);
    }
    if (gt) {
      const float t = __ldg(gt + i);
      if (t > 0.f) g += d > t ? g_l1 : (d < t ? -g_l1 : 0.f);
    }
    dL_ddepth[i] = (float)g;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) loss[0] += (float)((double)w_pearson * pseudo_loss + (gt ? (double)w_l1 * sums[8] / N : 0.0));
}

void launch_depth_loss_grad(const float* depth, const float* pseudo, const float* gt, int n, float k, float w_pearson, float w_l1,
                            float* loss, float* dL_ddepth, double* scratch, cudaStream_t stream) {
  cudaMemsetAsync(scratch, 0, 9 * sizeof(double), stream);
  const int blocks = std::min((n + 255) / 256, sm_count() * 4);
  depth_loss_sums_kernel<<<blocks, 256, 0, stream>>>(depth, pseudo, gt, n, k, scratch);
  depth_loss_grad_kernel<<<blocks, 256, 0, stream>>>(depth, pseudo, gt, n, k, w_pearson, w_l1, scratch, dL_ddepth, loss);
  count_launch(2);
}

}  // namespace gsr

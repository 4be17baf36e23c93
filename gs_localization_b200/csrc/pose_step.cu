// The two small device-side pieces that sit either side of the rasterizer in LoGS' pose-refinement
// loop (`gradient_decent`, gs_localization/pipelines/7scenes_localize_full_dslam.py:66-91), fused so
// that an iteration needs no host round trip and no framework ops:
//
//   l1_loss_grad_kernel   tracking loss with all-ones masks (tools/descent_utils.py:85-123):
//                         L = mean|I - I*| (+ w mean|D - D*|); writes dL/dI (and dL/dD) directly,
//                         i.e. forward and backward of the loss in one pass over the image.
//   pose_adam_step_kernel torch.optim.Adam on the six pose deltas (lr per group, betas 0.9/0.999,
//                         eps 1e-8; the deltas are zero before every step, tools/pose_utils.py:120-121),
//                         tau = [rho, theta], T_w2c <- SE3_exp(tau) T_w2c (tools/pose_utils.py:54-122),
//                         then the rasterizer's per-view constants: viewmatrix = T_w2c^T,
//                         projmatrix = viewmatrix @ projmatrix_raw, campos = -R^T t
//                         (tools/camera_utils.py:144-158).  One thread.
#include "gsr_kernels.cuh"

namespace gsr {

__global__ void __launch_bounds__(256) l1_loss_grad_kernel(const float* __restrict__ image, const float* __restrict__ target,
                                                           float* __restrict__ dL_dimage, size_t n, float scale,
                                                           float* __restrict__ loss_out) {
  __shared__ float s_part[8];
  float acc = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float d = image[i] - __ldg(target + i);
    acc += fabsf(d);
    dL_dimage[i] = d > 0.f ? scale : (d < 0.f ? -scale : 0.f);   // torch.sign semantics
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; w++) t += s_part[w];
    atomicAdd(loss_out, t * scale);
  }
}

void launch_l1_loss_grad(const float* image, const float* target, float* dL_dimage, size_t n, float weight, float* loss_out,
                         cudaStream_t stream) {
  if (n == 0) return;
  const int blocks = (int)std::min<size_t>((n + 256 * 4 - 1) / (256 * 4), sm_count() * 4);
  l1_loss_grad_kernel<<<blocks, 256, 0, stream>>>(image, target, dL_dimage, n, weight / (float)n, loss_out);
  count_launch();
}

// Full tracking loss of LoGS (tools/descent_utils.py:85-123) and its gradient in one pass over the pixels:
//   image_ab = exp(a) * image + b                                     (:86)
//   L_rgb    = mean_{3HW} om * |image_ab * gm - gt * gm|              (:104-106)   om = opacity > threshold
//   L_depth  = mean_{HW} |depth * dm - gt_depth * dm|,  dm = (gt_depth > 0.01) * om * gm      (:116-123)
//   L        = L_rgb + depth_weight * L_depth       (depth_weight = 1 - alpha; 0 / no gt_depth = monocular)
// The masks are thresholds, so no gradient reaches opacity.  One thread per pixel, all three channels.
__global__ void __launch_bounds__(256) tracking_loss_grad_kernel(const float* __restrict__ image, const float* __restrict__ depth,
                                                                 const float* __restrict__ opacity, const float* __restrict__ gt_image,
                                                                 const float* __restrict__ gt_depth, const float* __restrict__ grad_mask,
                                                                 const float* __restrict__ exposure, int npix, float opacity_threshold,
                                                                 float depth_weight, float* __restrict__ dL_dimage,
                                                                 float* __restrict__ dL_ddepth, float* __restrict__ loss_out,
                                                                 float* __restrict__ dL_dexposure) {
  __shared__ float s_part[3][8];
  const float ea = exposure ? expf(exposure[0]) : 1.f, eb = exposure ? exposure[1] : 0.f;
  const float s_rgb = 1.f / (3.f * (float)npix), s_d = depth_weight / (float)npix;
  float acc = 0.f, acc_a = 0.f, acc_b = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
    const float gm = grad_mask ? __ldg(grad_mask + i) : 1.f;
    const float om = opacity[i] > opacity_threshold ? 1.f : 0.f;
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const float I = image[(size_t)c * npix + i];
      const float d = (ea * I + eb) * gm - __ldg(gt_image + (size_t)c * npix + i) * gm;
      const float sg = (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * om * gm;
      acc += om * fabsf(d) * s_rgb;
      dL_dimage[(size_t)c * npix + i] = sg * ea * s_rgb;
      acc_a += sg * ea * I;
      acc_b += sg;
    }
    if (dL_ddepth) {
      float g = 0.f;
      if (gt_depth) {
        const float gd = __ldg(gt_depth + i);
        const float dm = (gd > 0.01f ? 1.f : 0.f) * om * gm;
        const float d = depth[i] * dm - gd * dm;
        acc += fabsf(d) * s_d;
        g = (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * dm * s_d;
      }
      dL_ddepth[i] = g;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
    acc_a += __shfl_xor_sync(0xffffffffu, acc_a, o);
    acc_b += __shfl_xor_sync(0xffffffffu, acc_b, o);
  }
  if ((threadIdx.x & 31) == 0) s_part[0][threadIdx.x >> 5] = acc, s_part[1][threadIdx.x >> 5] = acc_a, s_part[2][threadIdx.x >> 5] = acc_b;
  __syncthreads();
  if (threadIdx.x < 3) {
    float t = 0.f;
    for (int w = 0; w < 8; w++) t += s_part[threadIdx.x][w];
    if (threadIdx.x == 0) atomicAdd(loss_out, t);
    else if (dL_dexposure) atomicAdd(dL_dexposure + (threadIdx.x - 1), t * s_rgb);
  }
}

void launch_tracking_loss_grad(const float* image, const float* depth, const float* opacity, const float* gt_image, const float* gt_depth,
                               const float* grad_mask, const float* exposure, int npix, float opacity_threshold, float depth_weight,
                               float* dL_dimage, float* dL_ddepth, float* loss_out, float* dL_dexposure, cudaStream_t stream) {
  if (npix <= 0) return;
  const int blocks = std::min((npix + 255) / 256, sm_count() * 4);
  tracking_loss_grad_kernel<<<blocks, 256, 0, stream>>>(image, depth, opacity, gt_image, gt_depth, grad_mask, exposure, npix,
                                                        opacity_threshold, depth_weight, dL_dimage, dL_ddepth, loss_out, dL_dexposure);
  count_launch();
}

// torch.optim.Adam on the two exposure scalars (7scenes_localize_full_dslam.py:48-61: lr 1e-3 each); consumes and
// clears their gradient accumulators so that the next iteration's loss kernel can add into them.
__global__ void exposure_adam_step_kernel(float* __restrict__ exposure, float* __restrict__ dL_dexposure, float* __restrict__ adam_m,
                                          float* __restrict__ adam_v, float* __restrict__ step_count, float lr) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
  const float t = step_count[0] + 1.0f;
  step_count[0] = t;
  const float bc1 = 1.0f - powf(b1, t), bc2 = 1.0f - powf(b2, t);
  for (int i = 0; i < 2; i++) {
    const float g = dL_dexposure[i];
    dL_dexposure[i] = 0.f;
    const float m = b1 * adam_m[i] + (1.0f - b1) * g;
    const float v = b2 * adam_v[i] + (1.0f - b2) * g * g;
    adam_m[i] = m;
    adam_v[i] = v;
    exposure[i] -= (lr / bc1) * (m / (sqrtf(v) / sqrtf(bc2) + eps));
  }
}

void launch_exposure_adam_step(float* exposure, float* dL_dexposure, float* adam_m, float* adam_v, float* step_count, float lr,
                               cudaStream_t stream) {
  exposure_adam_step_kernel<<<1, 32, 0, stream>>>(exposure, dL_dexposure, adam_m, adam_v, step_count, lr);
  count_launch();
}

// state: m[6], v[6], step (as float), then w2c[16] row-major, raw projection (transposed storage) [16]
__global__ void pose_adam_step_kernel(const float* __restrict__ dL_dtau, float* __restrict__ adam_m, float* __restrict__ adam_v,
                                      float* __restrict__ step_count, float lr_trans, float lr_rot, float* __restrict__ w2c,
                                      const float* __restrict__ raw, float* __restrict__ viewmatrix,
                                      float* __restrict__ projmatrix, float* __restrict__ campos, float* __restrict__ tau_norm) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
  const float t = step_count[0] + 1.0f;
  step_count[0] = t;
  const float bc1 = 1.0f - powf(b1, t), bc2 = 1.0f - powf(b2, t);
  float tau[6];
  for (int i = 0; i < 6; i++) {
    const float g = dL_dtau[i];
    const float m = b1 * adam_m[i] + (1.0f - b1) * g;
    const float v = b2 * adam_v[i] + (1.0f - b2) * g * g;
    adam_m[i] = m;
    adam_v[i] = v;
    const float lr = i < 3 ? lr_trans : lr_rot;
    const float denom = sqrtf(v) / sqrtf(bc2) + eps;      // torch.optim.Adam
    tau[i] = -(lr / bc1) * (m / denom);                   // the delta parameters start every step at zero
  }
  if (tau_norm) tau_norm[0] = sqrtf(tau[0] * tau[0] + tau[1] * tau[1] + tau[2] * tau[2] + tau[3] * tau[3] + tau[4] * tau[4] + tau[5] * tau[5]);
  // SE3_exp (tools/pose_utils.py:54-102)
  const float rx = tau[3], ry = tau[4], rz = tau[5];
  const float Wm[3][3] = {{0.f, -rz, ry}, {rz, 0.f, -rx}, {-ry, rx, 0.f}};
  float W2[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) W2[i][j] = Wm[i][0] * Wm[0][j] + Wm[i][1] * Wm[1][j] + Wm[i][2] * Wm[2][j];
  const float a = sqrtf(rx * rx + ry * ry + rz * rz);
  float cA, cB, cC, cD;   // R = I + cA W + cB W2 ; V = I + cC W + cD W2
  if (a < 1e-5f) {
    cA = 1.f, cB = 0.5f, cC = 0.5f, cD = 1.0f / 6.0f;
  } else {
    cA = sinf(a) / a, cB = (1.f - cosf(a)) / (a * a), cC = cB, cD = (a - sinf(a)) / (a * a * a);
  }
  float E[4][4] = {{0}};
  for (int i = 0; i < 3; i++) {
    float vt = 0.f;
    for (int j = 0; j < 3; j++) {
      E[i][j] = (i == j ? 1.f : 0.f) + cA * Wm[i][j] + cB * W2[i][j];
      vt += ((i == j ? 1.f : 0.f) + cC * Wm[i][j] + cD * W2[i][j]) * tau[j];
    }
    E[i][3] = vt;
  }
  E[3][3] = 1.f;
  float N[4][4];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      float s = 0.f;
      for (int k = 0; k < 4; k++) s += E[i][k] * w2c[4 * k + j];
      N[i][j] = s;
    }
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      w2c[4 * i + j] = N[i][j];
      viewmatrix[4 * j + i] = N[i][j];        // transposed storage
    }
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      float s = 0.f;
      for (int k = 0; k < 4; k++) s += N[k][i] * raw[4 * k + j];   // viewmatrix[i][k] = N[k][i]
      projmatrix[4 * i + j] = s;
    }
  for (int i = 0; i < 3; i++) campos[i] = -(N[0][i] * N[0][3] + N[1][i] * N[1][3] + N[2][i] * N[2][3]);
}

void launch_pose_adam_step(const float* dL_dtau, float* adam_m, float* adam_v, float* step_count, float lr_trans, float lr_rot,
                           float* w2c, const float* raw, float* viewmatrix, float* projmatrix, float* campos, float* tau_norm,
                           cudaStream_t stream) {
  pose_adam_step_kernel<<<1, 32, 0, stream>>>(dL_dtau, adam_m, adam_v, step_count, lr_trans, lr_rot, w2c, raw, viewmatrix,
                                              projmatrix, campos, tau_norm);
  count_launch();
}

}  // namespace gsr

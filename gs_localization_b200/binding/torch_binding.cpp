// Compiled host side of the drop-in `_C` module: the torch-facing glue the reference keeps in
// rasterize_points.cu (RasterizeGaussiansCUDA :35-119, RasterizeGaussiansBackwardCUDA :121-206, markVisible
// :208-227) and ext.cpp, written against the C ABI of libgsr_b200.so (include/gsr_b200.h).  It owns nothing but
// tensors and the current stream: outputs and the three scratch buffers are torch tensors grown through the
// allocation callbacks (the reference's resizeFunctional, rasterize_points.cu:27-33); every kernel is behind
// the C ABI.  The Python module gs_localization_b200/diff_gaussian_rasterization/_C.py wraps these entry points
// with the reference's argument order; the same file holds the equivalent ctypes binding.
#include <torch/extension.h>

#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include <algorithm>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/gsr_b200.h"

namespace {

struct Scratch {
  torch::Tensor geom, binning, img;
  torch::TensorOptions opts;
};
char* grow(torch::Tensor& t, const torch::TensorOptions& o, size_t n) {
  t = torch::empty({static_cast<int64_t>(n)}, o);
  return reinterpret_cast<char*>(t.data_ptr());
}
char* alloc_geom(size_t n, void* u) { auto* s = static_cast<Scratch*>(u); return grow(s->geom, s->opts, n); }
char* alloc_binning(size_t n, void* u) { auto* s = static_cast<Scratch*>(u); return grow(s->binning, s->opts, n); }
char* alloc_img(size_t n, void* u) { auto* s = static_cast<Scratch*>(u); return grow(s->img, s->opts, n); }

// .contiguous().data<float>() of the reference (rasterize_points.cu:96-115); empty placeholders become NULL
struct Arg {
  torch::Tensor keep;
  const float* ptr = nullptr;
  Arg(const torch::Tensor& t, const torch::Device& dev, torch::ScalarType st = torch::kFloat32) {
    if (!t.defined() || t.numel() == 0) return;
    if (t.device() == dev && t.scalar_type() == st && t.is_contiguous()) keep = t;
    else keep = t.to(torch::TensorOptions().device(dev).dtype(st)).contiguous();
    ptr = reinterpret_cast<const float*>(keep.data_ptr());
  }
};

[[noreturn]] void raise(const char* what) { throw std::runtime_error(std::string(what) + ": " + gsr_last_error()); }

std::tuple<int64_t, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
forward(const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& colors, const torch::Tensor& opacity,
        const torch::Tensor& scales, const torch::Tensor& rotations, double scale_modifier, const torch::Tensor& cov3D_precomp,
        const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix, double tan_fovx, double tan_fovy, int64_t image_height,
        int64_t image_width, const torch::Tensor& sh, int64_t degree, const torch::Tensor& campos, bool prefiltered, bool debug,
        bool want_n_touched) {
  if (means3D.ndimension() != 2 || means3D.size(1) != 3) throw std::runtime_error("means3D must have dimensions (num_points, 3)");
  if (!means3D.is_cuda()) throw std::runtime_error("gs_localization_b200: tensors must be CUDA tensors (no CPU fallback exists)");
  const auto dev = means3D.device();
  const int64_t P = means3D.size(0), H = image_height, W = image_width;
  const auto f32 = torch::TensorOptions().device(dev).dtype(torch::kFloat32);
  const auto i32 = torch::TensorOptions().device(dev).dtype(torch::kInt32);
  Scratch s;
  s.opts = torch::TensorOptions().device(dev).dtype(torch::kUInt8);
  torch::Tensor n_touched = torch::zeros({want_n_touched ? P : 0}, i32);
  if (P == 0) {   // rasterize_points.cu:83
    auto e = torch::empty({0}, s.opts);
    return {0, torch::zeros({3, H, W}, f32), torch::zeros({1, H, W}, f32), torch::zeros({1, H, W}, f32), torch::zeros({0}, i32), e, e, e, n_touched};
  }
  c10::cuda::CUDAGuard guard(dev);
  auto color = torch::empty({3, H, W}, f32), depth = torch::empty({1, H, W}, f32), alpha = torch::empty({1, H, W}, f32);
  auto radii = torch::empty({P}, i32);
  const int M = sh.defined() && sh.numel() != 0 ? (int)sh.size(1) : 0;
  Arg bg(background, dev), m3(means3D, dev), shc(sh, dev), col(colors, dev), opa(opacity, dev), sc(scales, dev), rot(rotations, dev),
      cov(cov3D_precomp, dev), vm(viewmatrix, dev), pm(projmatrix, dev), cp(campos, dev);
  const int64_t R = gsr_rasterize_forward(
      alloc_geom, alloc_binning, alloc_img, &s, (int)P, (int)degree, M, bg.ptr, (int)W, (int)H, m3.ptr, shc.ptr, col.ptr, opa.ptr, sc.ptr,
      (float)scale_modifier, rot.ptr, cov.ptr, vm.ptr, pm.ptr, cp.ptr, (float)tan_fovx, (float)tan_fovy, prefiltered ? 1 : 0,
      color.data_ptr<float>(), depth.data_ptr<float>(), alpha.data_ptr<float>(), radii.data_ptr<int>(),
      want_n_touched ? n_touched.data_ptr<int>() : nullptr, debug ? 1 : 0, c10::cuda::getCurrentCUDAStream(dev.index()).stream());
  if (R < 0) raise("gsr_rasterize_forward");
  auto e = torch::empty({0}, s.opts);
  return {R, color, depth, alpha, radii, s.geom.defined() ? s.geom : e, s.binning.defined() ? s.binning : e, s.img.defined() ? s.img : e, n_touched};
}

// needs: bit i set -> allocate and compute gradient i of {means3D, means2D, sh, colors, opacity, scales, rotations, cov3D}
std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
backward(const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& radii, const torch::Tensor& colors,
         const torch::Tensor& scales, const torch::Tensor& rotations, double scale_modifier, const torch::Tensor& cov3D_precomp,
         const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix, double tan_fovx, double tan_fovy, const torch::Tensor& dL_dout_color,
         const torch::Tensor& dL_dout_depth, const torch::Tensor& dL_dout_alpha, const torch::Tensor& sh, int64_t degree,
         const torch::Tensor& campos, const torch::Tensor& geomBuffer, int64_t R, const torch::Tensor& binningBuffer,
         const torch::Tensor& imageBuffer, const torch::Tensor& alpha, bool debug, const torch::Tensor& projmatrix_raw, bool want_pose,
         int64_t needs) {
  if (!means3D.is_cuda()) throw std::runtime_error("gs_localization_b200: tensors must be CUDA tensors (no CPU fallback exists)");
  const auto dev = means3D.device();
  const int64_t P = means3D.size(0), H = dL_dout_color.size(1), W = dL_dout_color.size(2);
  const int M = sh.defined() && sh.numel() != 0 ? (int)sh.size(1) : 0;
  const auto f32 = torch::TensorOptions().device(dev).dtype(torch::kFloat32);
  // Every row is written by the library (zeros for culled Gaussians): empty, not the reference's nine zeros.  The
  // gradients are carved out of ONE allocation, 128-byte aligned slices in a fixed order, so that the library's dense
  // zero fill is a single memset over the arena instead of one per tensor.
  const int64_t widths[8] = {3, 3, 3 * (int64_t)M, 3, 1, 3, 4, 6};   // means3D, means2D, sh, colors, opacity, scales, rotations, cov3D
  int64_t offs[8], total = 0;
  for (int b = 0; b < 8; b++) {
    offs[b] = total;
    if ((needs >> b) & 1) total += (P * widths[b] + 31) / 32 * 32;
  }
  torch::Tensor arena = P > 0 ? torch::empty({std::max<int64_t>(total, 1)}, f32) : torch::zeros({std::max<int64_t>(total, 1)}, f32);
  auto mk = [&](int bit, std::vector<int64_t> shape) -> torch::Tensor {
    if (!((needs >> bit) & 1)) return torch::Tensor();
    return arena.narrow(0, offs[bit], P * widths[bit]).view(shape);
  };
  auto dL_dmeans3D = mk(0, {P, 3}), dL_dmeans2D = mk(1, {P, 3}), dL_dsh = mk(2, {P, M, 3}), dL_dcolors = mk(3, {P, 3}),
       dL_dopacity = mk(4, {P, 1}), dL_dscales = mk(5, {P, 3}), dL_drotations = mk(6, {P, 4}), dL_dcov3D = mk(7, {P, 6});
  torch::Tensor dL_dtau = want_pose ? torch::zeros({6}, f32) : torch::Tensor();
  if (P > 0) {
    c10::cuda::CUDAGuard guard(dev);
    Arg bg(background, dev), m3(means3D, dev), shc(sh, dev), col(colors, dev), alp(alpha, dev), sc(scales, dev), rot(rotations, dev),
        cov(cov3D_precomp, dev), vm(viewmatrix, dev), pm(projmatrix, dev), praw(projmatrix_raw, dev), cp(campos, dev),
        rad(radii, dev, torch::kInt32), gC(dL_dout_color, dev), gD(dL_dout_depth, dev), gA(dL_dout_alpha, dev);
    auto fp = [](torch::Tensor& t) { return t.defined() ? t.data_ptr<float>() : nullptr; };
    auto bp = [](const torch::Tensor& t) { return t.defined() && t.numel() ? reinterpret_cast<char*>(t.data_ptr()) : nullptr; };
    const int rc = gsr_rasterize_backward(
        (int)P, (int)degree, M, R, bg.ptr, (int)W, (int)H, m3.ptr, shc.ptr, col.ptr, alp.ptr, sc.ptr, (float)scale_modifier, rot.ptr, cov.ptr,
        vm.ptr, pm.ptr, praw.ptr, cp.ptr, (float)tan_fovx, (float)tan_fovy, reinterpret_cast<const int*>(rad.ptr), bp(geomBuffer),
        bp(binningBuffer), bp(imageBuffer), gC.ptr, gD.ptr, gA.ptr, fp(dL_dmeans2D), nullptr, fp(dL_dopacity), fp(dL_dcolors),
        fp(dL_dmeans3D), fp(dL_dcov3D), fp(dL_dsh), fp(dL_dscales), fp(dL_drotations), fp(dL_dtau), debug ? 1 : 0,
        c10::cuda::getCurrentCUDAStream(dev.index()).stream());
    if (rc != 0) raise("gsr_rasterize_backward");
  }
  // the reference's return order (rasterize_points.cu:205) + the pose gradient
  return {dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations, dL_dtau};
}

torch::Tensor mark_visible(const torch::Tensor& means3D, const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix) {
  if (!means3D.is_cuda()) throw std::runtime_error("gs_localization_b200: tensors must be CUDA tensors (no CPU fallback exists)");
  const auto dev = means3D.device();
  const int64_t P = means3D.size(0);
  auto present = torch::zeros({P}, torch::TensorOptions().device(dev).dtype(torch::kBool));
  if (P != 0) {
    c10::cuda::CUDAGuard guard(dev);
    Arg m3(means3D, dev), vm(viewmatrix, dev), pm(projmatrix, dev);
    if (gsr_mark_visible((int)P, m3.ptr, vm.ptr, pm.ptr, reinterpret_cast<unsigned char*>(present.data_ptr()),
                         c10::cuda::getCurrentCUDAStream(dev.index()).stream()) != 0)
      raise("gsr_mark_visible");
  }
  return present;
}

constexpr auto forward_impl = &forward;
constexpr auto backward_impl = &backward;

// The autograd node of the plain (non-pose) rasterizer in C++: same contract as the Python
// `_RasterizeGaussians` (reference diff_gaussian_rasterization/__init__.py:44-158: 9 inputs, 4 outputs, radii
// non-differentiable, gradients in input order), without the Python trampoline on either side of the autograd engine.
// The Python class remains for debug mode (its snapshot files) and for the ctypes binding.
struct RasterizeFn : public torch::autograd::Function<RasterizeFn> {
  static torch::autograd::variable_list forward(torch::autograd::AutogradContext* ctx, torch::Tensor means3D, torch::Tensor means2D,
                                                torch::Tensor sh, torch::Tensor colors_precomp, torch::Tensor opacities,
                                                torch::Tensor scales, torch::Tensor rotations, torch::Tensor cov3Ds_precomp,
                                                torch::Tensor bg, double scale_modifier, torch::Tensor viewmatrix,
                                                torch::Tensor projmatrix, double tanfovx, double tanfovy, int64_t image_height,
                                                int64_t image_width, int64_t sh_degree, torch::Tensor campos, bool prefiltered) {
    (void)means2D;   // gradient sink only, never read (reference __init__.py:60-80)
    ctx->set_materialize_grads(false);   // outputs the loss does not use arrive as undefined gradients, not as zero tensors + fill kernels
    auto out = forward_impl(bg, means3D, colors_precomp, opacities, scales, rotations, scale_modifier, cov3Ds_precomp, viewmatrix, projmatrix,
                         tanfovx, tanfovy, image_height, image_width, sh, sh_degree, campos, prefiltered, false, false);
    auto& color = std::get<1>(out);
    auto& depth = std::get<2>(out);
    auto& alpha = std::get<3>(out);
    auto& radii = std::get<4>(out);
    ctx->save_for_backward({colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, std::get<5>(out), std::get<6>(out),
                            std::get<7>(out), alpha, bg, viewmatrix, projmatrix, campos});
    ctx->saved_data["R"] = std::get<0>(out);
    ctx->saved_data["scale_modifier"] = scale_modifier;
    ctx->saved_data["tanfovx"] = tanfovx;
    ctx->saved_data["tanfovy"] = tanfovy;
    ctx->saved_data["H"] = image_height;
    ctx->saved_data["W"] = image_width;
    ctx->saved_data["degree"] = sh_degree;
    ctx->mark_non_differentiable({radii});
    return {color, radii, depth, alpha};
  }

  static torch::autograd::variable_list backward(torch::autograd::AutogradContext* ctx, torch::autograd::variable_list grad_out) {
    const auto saved = ctx->get_saved_variables();
    const auto& colors_precomp = saved[0];
    const auto& means3D = saved[1];
    const auto& scales = saved[2];
    const auto& rotations = saved[3];
    const auto& cov3Ds_precomp = saved[4];
    const auto& radii = saved[5];
    const auto& sh = saved[6];
    const auto& alpha = saved[10];
    const int64_t H = ctx->saved_data["H"].toInt(), W = ctx->saved_data["W"].toInt();
    const auto f32 = torch::TensorOptions().device(means3D.device()).dtype(torch::kFloat32);
    torch::Tensor gC = grad_out[0].defined() ? grad_out[0] : torch::zeros({3, H, W}, f32);
    const torch::Tensor& gD = grad_out[2];   // undefined == no upstream gradient: the library takes NULL for these two
    const torch::Tensor& gA = grad_out[3];
    // forward inputs 0..7: means3D, means2D, sh, colors, opacities, scales, rotations, cov3D == the bits of `needs`
    int64_t needs = 0;
    for (int i = 0; i < 8; i++)
      if (ctx->needs_input_grad(i)) needs |= 1ll << i;
    auto res = backward_impl(saved[11], means3D, radii, colors_precomp, scales, rotations, ctx->saved_data["scale_modifier"].toDouble(),
                          cov3Ds_precomp, saved[12], saved[13], ctx->saved_data["tanfovx"].toDouble(), ctx->saved_data["tanfovy"].toDouble(),
                          gC, gD, gA, sh, ctx->saved_data["degree"].toInt(), saved[14], saved[7], ctx->saved_data["R"].toInt(), saved[8],
                          saved[9], alpha, false, torch::Tensor(), false, needs);
    // res: (means2D, colors, opacity, means3D, cov3D, sh, scales, rotations, tau); gradients in forward-input order
    torch::Tensor none;
    return {std::get<3>(res), std::get<0>(res), std::get<5>(res), std::get<1>(res), std::get<2>(res), std::get<6>(res), std::get<7>(res),
            std::get<4>(res), none, none, none, none, none, none, none, none, none, none, none};
  }
};

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor> rasterize(
    torch::Tensor means3D, torch::Tensor means2D, torch::Tensor sh, torch::Tensor colors_precomp, torch::Tensor opacities,
    torch::Tensor scales, torch::Tensor rotations, torch::Tensor cov3Ds_precomp, torch::Tensor bg, double scale_modifier,
    torch::Tensor viewmatrix, torch::Tensor projmatrix, double tanfovx, double tanfovy, int64_t image_height, int64_t image_width,
    int64_t sh_degree, torch::Tensor campos, bool prefiltered) {
  auto out = RasterizeFn::apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, bg, scale_modifier,
                                viewmatrix, projmatrix, tanfovx, tanfovy, image_height, image_width, sh_degree, campos, prefiltered);
  return {out[0], out[1], out[2], out[3]};
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("forward", &forward);
  m.def("backward", &backward);
  m.def("mark_visible", &mark_visible);
  m.def("rasterize", &rasterize);
  m.def("abi_version", []() { return gsr_abi_version(); });
}

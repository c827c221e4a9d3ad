// O3DCudaInfluence.h - the host side of Omega3D's `gpu_cuda` influence arm, over the C ABI (include/o3d_cuda.h).
//
// Header-only C++17. It is written against the ACCESSOR NAMES of the reference's element containers
// (Points<S>, Surfaces<S>: get_pos/get_str/get_rad/get_vel/get_velgrad/get_idx/get_area/..., reference
// src/Points.h, src/Surfaces.h, src/ElementBase.h) but includes none of the reference's files: the
// functions are templates over the container types, so the same header serves the patched reference
// (integration/omega3d_use_cuda.patch adds one `#include` and one `if (env.get_instrs() == gpu_cuda)` arm per
// routine) and any other host that exposes the same SoA accessors.
//
// One function per reference routine, same argument meaning, same accumulate-into-target semantics:
//   cuda_points_affect_points   <- points_affect_points<S,A>    src/Influence.h:67-551
//   cuda_panels_affect_points   <- panels_affect_points<S,A>    src/Influence.h:557-1099
//   cuda_points_affect_panels   <- points_affect_panels<S,A>    src/Influence.h:1107-1221
//   cuda_panels_on_panels_coeff <- panels_on_panels_coeff<S>    src/Coefficients.h:169-483
// (panels_affect_panels, src/Influence.h:1224-1245, needs no arm of its own: it calls panels_affect_points.)
// Each returns the reference's own flop estimate for the call so the caller can print its usual
// "[%.4f] seconds at %.3f GFlop/s" line. Errors: the reference aborts on broken preconditions (assert is live
// in its Release builds, CMakeLists.txt:60); the C ABI returns codes, and this layer turns a non-zero code
// into the same behaviour - message on stderr, then abort().
#pragma once

#include <array>
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "o3d_cuda.h"

namespace o3d {

// The core function of the reference build this header is compiled into: src/CoreFunc.h:35-38 leaves exactly one
// USE_*_KERNEL defined (USE_WL_KERNEL as shipped) and src/Kernels.h includes it before the patched Influence.h /
// Coefficients.h include this header, so the CUDA arm follows an edit of that line like every CPU arm does.
#if defined(USE_RM_KERNEL)
constexpr int kCudaCoreFunc = O3D_CORE_RM;
#elif defined(USE_EXPONENTIAL_KERNEL)
constexpr int kCudaCoreFunc = O3D_CORE_EXP;
#elif defined(USE_V2_KERNEL)
constexpr int kCudaCoreFunc = O3D_CORE_V2;
#else
constexpr int kCudaCoreFunc = O3D_CORE_WL;
#endif

// One context per process, created on first use. Devices: all visible GPUs, or the first
// $O3D_CUDA_NDEV of them (targets are partitioned across them inside the library).
inline o3d_ctx* cuda_context() {
  struct Holder {
    o3d_ctx* ctx = nullptr;
    Holder() {
      int ndev = o3d_cuda_device_count();
      if (const char* e = std::getenv("O3D_CUDA_NDEV")) {
        const int want = std::atoi(e);
        if (want >= 1 && want < ndev) ndev = want;
      }
      const int rc = ndev >= 1 ? o3d_cuda_create(&ctx, ndev, nullptr) : O3D_ERR_NODEVICE;
      if (rc != O3D_OK) {
        std::fprintf(stderr, "Omega3D gpu_cuda arm: no usable sm_100 device (o3d_cuda_create -> %d); there is no CPU fallback in this arm\n", rc);
        std::abort();
      }
      o3d_cuda_set_core_func(ctx, kCudaCoreFunc);
    }
    ~Holder() { o3d_cuda_destroy(ctx); }
  };
  static Holder h;
  return h.ctx;
}

inline void cuda_check(int rc, const char* what) {
  if (rc == O3D_OK) return;
  std::fprintf(stderr, "Omega3D gpu_cuda arm: %s failed with code %d: %s\n", what, rc, o3d_cuda_last_error(cuda_context()));
  std::abort();
}

// While a convection step runs on RESIDENT particles around a body (O3DCudaConvection.h), the host copy of the particle
// collection is stale. The reference's solve_bem (src/BEMHelper.h:83-94) still asks for points_affect_panels(vort, bdry) at that
// moment; the sums it wants have just been formed on the device from the resident state. This slot carries them: when it
// is armed, the gpu_cuda arm of points_affect_panels adds them into the target instead of evaluating the (stale) host arrays.
struct ResidentPanelSums {
  const float* raw = nullptr;    // 3 x np floats (u | v | w rows): zero minus the particle sums, un-normalised
  int64_t np = 0;
};
inline ResidentPanelSums& resident_panel_sums() {
  static thread_local ResidentPanelSums s;
  return s;
}

// `want_grad`: the caller's results type asks for gradients (ResultsType::compute_grad / get_type()==velandgrad).
template <class PointsT>
double cuda_points_affect_points(const PointsT& src, PointsT& targ, const bool want_grad) {
  using S = std::remove_cv_t<std::remove_reference_t<decltype(src.get_rad()[0])>>;
  static_assert(std::is_same<S, float>::value, "the CUDA arm stores float, like the reference's STORE type");
  const auto& sx = src.get_pos();
  const auto& ss = src.get_str();
  const auto& sr = src.get_rad();
  const auto& tx = targ.get_pos();
  auto& tu = targ.get_vel();
  auto& opttug = targ.get_velgrad();
  const float* tr = targ.is_inert() ? nullptr : targ.get_rad().data();
  float* tug[9];
  bool grad = false;
  if (want_grad) {
    // inert targets without gradient storage + velandgrad is the reference's assert(false), src/Influence.h:368-370
    if (!opttug) cuda_check(O3D_ERR_UNSUPPORTED, "points_affect_points (velandgrad on a target without gradient storage)");
    for (int k = 0; k < 9; ++k) tug[k] = (*opttug)[k].data();
    grad = true;
  }
  double flops = 0.0;
  cuda_check(o3d_cuda_pts_on_pts(cuda_context(), (int64_t)src.get_n(), sx[0].data(), sx[1].data(), sx[2].data(), sr.data(),
                                 ss[0].data(), ss[1].data(), ss[2].data(), (int64_t)targ.get_n(), tx[0].data(), tx[1].data(),
                                 tx[2].data(), tr, tu[0].data(), tu[1].data(), tu[2].data(), grad ? tug : nullptr, &flops),
             "points_affect_points");
  return flops;
}

template <class SurfacesT, class PointsT>
double cuda_panels_affect_points(const SurfacesT& src, PointsT& targ) {
  const auto& sx = src.get_pos();
  const auto& si = src.get_idx();
  const auto& ss = src.get_str();
  const auto& sa = src.get_area();
  const auto& tx = targ.get_pos();
  auto& tu = targ.get_vel();
  auto& opttug = targ.get_velgrad();
  const float* sss = src.have_src_str() ? src.get_src_str().data() : nullptr;
  float* tug[9];
  if (opttug)  // gradients iff the target stores them (src/Influence.h:652,875)
    for (int k = 0; k < 9; ++k) tug[k] = (*opttug)[k].data();
  double flops = 0.0;
  cuda_check(o3d_cuda_pan_on_pts(cuda_context(), (int64_t)sx[0].size(), sx[0].data(), sx[1].data(), sx[2].data(),
                                 (int64_t)src.get_npanels(), si.data(), ss[0].data(), ss[1].data(), ss[2].data(), sa.data(), sss,
                                 (int64_t)targ.get_n(), tx[0].data(), tx[1].data(), tx[2].data(), tu[0].data(), tu[1].data(),
                                 tu[2].data(), opttug ? tug : nullptr, &flops),
             "panels_affect_points");
  return flops;
}

template <class PointsT, class SurfacesT>
double cuda_points_affect_panels(const PointsT& src, SurfacesT& targ) {
  const auto& sx = src.get_pos();
  const auto& ss = src.get_str();
  const auto& tx = targ.get_pos();
  const auto& ti = targ.get_idx();
  const auto& ta = targ.get_area();
  auto& tu = targ.get_vel();
  double flops = 0.0;
  const ResidentPanelSums& rs = resident_panel_sums();
  if (rs.raw) {                  // a resident step is in flight: its device state is the source, not src's host arrays
    if ((int64_t)targ.get_npanels() != rs.np) cuda_check(O3D_ERR_INVALID, "points_affect_panels (resident sums for another surface)");
    for (int d = 0; d < 3; ++d)
      for (int64_t i = 0; i < rs.np; ++i) tu[d][i] += rs.raw[(size_t)d * rs.np + i];
    return 3.0 * (double)rs.np;
  }
  cuda_check(o3d_cuda_pts_on_pan(cuda_context(), (int64_t)src.get_n(), sx[0].data(), sx[1].data(), sx[2].data(), ss[0].data(),
                                 ss[1].data(), ss[2].data(), (int64_t)tx[0].size(), tx[0].data(), tx[1].data(), tx[2].data(),
                                 (int64_t)targ.get_npanels(), ti.data(), ta.data(), tu[0].data(), tu[1].data(), tu[2].data(), &flops),
             "points_affect_panels");
  return flops;
}

// Returns the column-major (3 ntarg) x (3 nsrc) block. Requires 3 unknowns per panel on both sides
// (the reference's default, source_str_is_unknown = true, src/Surfaces.h:71,261).
template <class SurfacesT>
std::vector<float> cuda_panels_on_panels_coeff(const SurfacesT& src, SurfacesT& targ, double* flops_out = nullptr) {
  if (src.num_unknowns_per_panel() != 3 || targ.num_unknowns_per_panel() != 3)
    cuda_check(O3D_ERR_UNSUPPORTED, "panels_on_panels_coeff (CUDA arm builds the 3-unknown block only)");
  const auto& sx = src.get_pos();
  const auto& tx = targ.get_pos();
  const size_t ns = src.get_npanels(), nt = targ.get_npanels();
  // bases arrive as 3 separate vectors per direction; the ABI takes x|y|z concatenated
  auto flat3 = [](const auto& b, size_t n) {
    std::vector<float> f(3 * n);
    for (int d = 0; d < 3; ++d) std::copy(b[d].begin(), b[d].begin() + n, f.begin() + d * n);
    return f;
  };
  const std::vector<float> sb1 = flat3(src.get_x1(), ns), sb2 = flat3(src.get_x2(), ns);
  const std::vector<float> tb1 = flat3(targ.get_x1(), nt), tb2 = flat3(targ.get_x2(), nt), tn = flat3(targ.get_norm(), nt);
  std::vector<float> coeffs(9 * ns * nt);
  double flops = 0.0;
  cuda_check(o3d_cuda_pan_on_pan_coeff(cuda_context(), (int64_t)sx[0].size(), sx[0].data(), sx[1].data(), sx[2].data(), (int64_t)ns,
                                       src.get_idx().data(), sb1.data(), sb2.data(), src.get_area().data(), (int64_t)tx[0].size(),
                                       tx[0].data(), tx[1].data(), tx[2].data(), (int64_t)nt, targ.get_idx().data(), tb1.data(),
                                       tb2.data(), tn.data(), targ.get_area().data(), (&src == &targ) ? 1 : 0, coeffs.data(), &flops),
             "panels_on_panels_coeff");
  if (flops_out) *flops_out = flops;
  return coeffs;
}

}  // namespace o3d

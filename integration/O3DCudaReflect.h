// O3DCudaReflect.h - the host side of Omega3D's `gpu_cuda` arm for the particle-vs-body clean-up loops, over the
// C ABI (include/o3d_cuda.h). Companion of O3DCudaInfluence.h; same conventions (templates over the reference's
// container accessors, no reference file included, non-zero return codes abort like the reference's asserts).
//
//   cuda_reflect_panp2      <- reflect_panp2<S>       reference src/Reflect.h:194-311
//   cuda_clear_inner_panp2  <- clear_inner_panp2<S>   reference src/Reflect.h:446-620 (method 1, the only one called)
// Both return the number of particles moved so the caller can print its usual "reflected N particles" line.
#pragma once

#include <vector>

#include "O3DCudaInfluence.h"

namespace o3d {

template <class SurfacesT>
inline std::vector<float> flat_normals(const SurfacesT& src) {
  const auto& sn = src.get_norm();
  const size_t np = src.get_npanels();
  std::vector<float> f(3 * np);
  for (int d = 0; d < 3; ++d) std::copy(sn[d].begin(), sn[d].begin() + np, f.begin() + d * np);
  return f;
}

template <class SurfacesT, class PointsT>
inline int64_t cuda_reflect_panp2(const SurfacesT& src, PointsT& targ) {
  const auto& sx = src.get_pos();
  auto& tx = targ.get_pos();
  const std::vector<float> nrm = flat_normals(src);
  int64_t moved = 0;
  cuda_check(o3d_cuda_reflect_pts(cuda_context(), (int64_t)sx[0].size(), sx[0].data(), sx[1].data(), sx[2].data(),
                                  (int64_t)src.get_npanels(), src.get_idx().data(), nrm.data(), (int64_t)targ.get_n(), tx[0].data(),
                                  tx[1].data(), tx[2].data(), &moved),
             "reflect_panp2");
  return moved;
}

template <class SurfacesT, class PointsT>
inline int64_t cuda_clear_inner_panp2(const int method, const SurfacesT& src, PointsT& targ, const float cutoff_mult, const float ips) {
  const auto& sx = src.get_pos();
  auto& tx = targ.get_pos();
  const std::vector<float> nrm = flat_normals(src);
  int64_t moved = 0;
  cuda_check(o3d_cuda_clear_inner_pts(cuda_context(), method, (int64_t)sx[0].size(), sx[0].data(), sx[1].data(), sx[2].data(),
                                      (int64_t)src.get_npanels(), src.get_idx().data(), nrm.data(), (int64_t)targ.get_n(),
                                      tx[0].data(), tx[1].data(), tx[2].data(), cutoff_mult, ips, &moved),
             "clear_inner_panp2");
  return moved;
}

}  // namespace o3d

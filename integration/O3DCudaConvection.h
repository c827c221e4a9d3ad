// O3DCudaConvection.h - the host side of Omega3D's `gpu_cuda` convection arm, over the C ABI (include/o3d_cuda.h).
//
// Companion of O3DCudaInfluence.h. The influence arm alone already moves >95 % of a step to the GPU, but it pays
// a host round trip of every array for each of the 2-3 evaluations of a Runge-Kutta step, and the O(N) work between
// them (zero_vels, finalize_vels, Points::move) stays on one host thread. For a system made of ONE vortex-particle
// collection and nothing else (no boundaries, no field points: every `*_nv.json` example) this header runs the whole
//     Convection<S,A,I>::advect(time, dt, fs, ips, vort, bdry, fldpt, bem)        reference src/Convection.h:208-228
// on the device: one upload, `order` evaluations + moves in HBM, one download.
//
// Written against the accessor names of the reference's Points<S> (get_pos/get_str/get_rad/get_elong/get_vel/
// get_velgrad/get_elemt/get_movet/update_max_str: src/Points.h, src/ElementBase.h); includes none of its files.
#pragma once

#include <array>

#include "O3DCudaInfluence.h"

namespace o3d {

// True when `advect` for this system can run entirely on the device.
template <class CollectionVec>
inline bool cuda_can_advect(const CollectionVec& vort, const CollectionVec& bdry, const CollectionVec& fldpt) {
  return vort.size() == 1 && bdry.empty() && fldpt.empty();
}

// Convection::advect for one active, lagrangian Points collection. Leaves in `pts` exactly what the reference's
// advect_1st / advect_2nd_ralston / advect_3rd leave: new positions, stretched strengths, elongation, the (combined)
// velocity, the first-stage velocity gradient, and the time-averaged peak strength (update_max_str, src/Points.h:535).
template <class PointsT>
double cuda_advect_particles(PointsT& pts, const int order, const double time, const double dt, const std::array<double, 3>& fs) {
  o3d_ctx* ctx = cuda_context();
  struct Holder {
    o3d_particles* p = nullptr;
    ~Holder() { if (p) o3d_cuda_particles_destroy(cuda_context(), p); }
  };
  static Holder h;   // the device-side collection lives across steps; only its contents are re-uploaded
  if (!h.p) cuda_check(o3d_cuda_particles_create(ctx, &h.p), "particles_create");
  auto& x = pts.get_pos();
  auto& s = pts.get_str();
  auto& r = pts.get_rad();
  auto& e = pts.get_elong();
  auto& u = pts.get_vel();
  auto& optug = pts.get_velgrad();
  const int64_t n = (int64_t)pts.get_n();
  cuda_check(o3d_cuda_particles_upload(ctx, h.p, n, x[0].data(), x[1].data(), x[2].data(), s[0].data(), s[1].data(), s[2].data(),
                                       r.data(), e.data()),
             "particles_upload");
  double flops = 0.0;
  cuda_check(o3d_cuda_particles_advect(ctx, h.p, order, time, dt, fs.data(), 1, &flops), "particles_advect");
  float* ug[9];
  if (optug)
    for (int k = 0; k < 9; ++k) ug[k] = (*optug)[k].data();
  cuda_check(o3d_cuda_particles_download(ctx, h.p, x[0].data(), x[1].data(), x[2].data(), s[0].data(), s[1].data(), s[2].data(),
                                         nullptr, e.data(), u[0].data(), u[1].data(), u[2].data(), optug ? ug : nullptr),
             "particles_download");
  pts.update_max_str();   // every Points::move ends with it (src/Points.h:350,438,519)
  return flops;
}

}  // namespace o3d

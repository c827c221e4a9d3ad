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

#include <algorithm>
#include <array>
#include <cmath>
#include <vector>

#include "O3DCudaInfluence.h"

namespace o3d {

// True when `advect` for this system can keep its particles on the device: one particle collection, no field points, and
// either no boundary (every `*_nv.json` example) or ONE boundary collection (3Dexamples/flow_over_sphere.json) - which the
// caller then checks to be a static Surfaces before taking cuda_advect_particles_body.
template <class CollectionVec>
inline bool cuda_can_advect(const CollectionVec& vort, const CollectionVec& bdry, const CollectionVec& fldpt) {
  return vort.size() == 1 && bdry.size() <= 1 && fldpt.empty();
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

// Convection::advect for one active, lagrangian Points collection around ONE static body (a Surfaces<S> whose Body, if any,
// does not move). The particles stay on the device for the whole step; the reference's BEM stays the reference's:
// `solve_bem_rest()` is called once per derivative evaluation, while resident_panel_sums() holds the panel-centre sums the
// device formed from the state being evaluated, and must run the reference's own
//     solve_bem<S,A,I>(time, fs, vort, bdry, bem)                                   src/BEMHelper.h:44-262
// unchanged - its points_affect_panels(vort, bdry) lands in the gpu_cuda arm, which delivers those sums; finalize_vels,
// the right-hand side, the A matrix (first call), the GMRES solve and set_str are the reference's host code. The strengths
// it leaves in `surf` go back to the device. After every move the device clears the inner layer (clear_inner_layer(1, bdry,
// vort, 0.5/sqrt(2 pi), ips), src/Convection.h:258,372,405).
template <class PointsT, class SurfacesT, class SolveRest>
double cuda_advect_particles_body(PointsT& pts, SurfacesT& surf, const int order, const double time, const double dt,
                                  const std::array<double, 3>& fs, const float ips, SolveRest&& solve_bem_rest) {
  o3d_ctx* ctx = cuda_context();
  struct Holder {
    o3d_particles* p = nullptr;
    ~Holder() { if (p) o3d_cuda_particles_destroy(cuda_context(), p); }
  };
  static Holder h;
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
  // the body: geometry once per call (it is static; a few KB), strengths per evaluation through the callback
  const auto& bx = surf.get_pos();
  const auto& bn = surf.get_norm();
  const int64_t np = (int64_t)surf.get_npanels();
  std::vector<float> nrm(3 * (size_t)np);
  for (int d = 0; d < 3; ++d) std::copy(bn[d].begin(), bn[d].begin() + np, nrm.begin() + (size_t)d * np);
  struct Ctx {
    SurfacesT* surf;
    std::remove_reference_t<SolveRest>* rest;
  } cb{&surf, &solve_bem_rest};
  auto trampoline = [](void* user, int64_t npan, const float* pu, float* tsx, float* tsy, float* tsz, float* sss, int* have_source) -> int {
    Ctx* c = static_cast<Ctx*>(user);
    ResidentPanelSums& slot = resident_panel_sums();
    slot.raw = pu;
    slot.np = npan;
    (*c->rest)();                  // the reference's solve_bem: leaves the solved strengths in the surface
    slot.raw = nullptr;
    const auto& ts = c->surf->get_str();
    std::copy(ts[0].begin(), ts[0].begin() + npan, tsx);
    std::copy(ts[1].begin(), ts[1].begin() + npan, tsy);
    std::copy(ts[2].begin(), ts[2].begin() + npan, tsz);
    *have_source = c->surf->have_src_str() ? 1 : 0;
    if (*have_source) {
      const auto& q = c->surf->get_src_str();
      std::copy(q.begin(), q.begin() + npan, sss);
    }
    return 0;
  };
  const float cutoff_mult = (float)(0.5 / std::sqrt(2.0 * 3.14159265358979323846));
  cuda_check(o3d_cuda_particles_set_body(ctx, h.p, (int64_t)bx[0].size(), bx[0].data(), bx[1].data(), bx[2].data(), np,
                                         surf.get_idx().data(), surf.get_area().data(), nrm.data(), cutoff_mult, ips,
                                         +trampoline, &cb),
             "particles_set_body");
  double flops = 0.0;
  cuda_check(o3d_cuda_particles_advect(ctx, h.p, order, time, dt, fs.data(), 1, &flops), "particles_advect");
  cuda_check(o3d_cuda_particles_clear_body(ctx, h.p), "particles_clear_body");
  float* ug[9];
  if (optug)
    for (int k = 0; k < 9; ++k) ug[k] = (*optug)[k].data();
  cuda_check(o3d_cuda_particles_download(ctx, h.p, x[0].data(), x[1].data(), x[2].data(), s[0].data(), s[1].data(), s[2].data(),
                                         nullptr, e.data(), u[0].data(), u[1].data(), u[2].data(), optug ? ug : nullptr),
             "particles_download");
  pts.update_max_str();
  return flops;
}

}  // namespace o3d

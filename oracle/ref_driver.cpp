// oracle/ref_driver.cpp - C-callable wrapper around the REFERENCE's own hot-path templates.
//
// TEST INFRASTRUCTURE ONLY. This file contains no Biot-Savart arithmetic of its own: it
// builds the reference's containers (Points<float>, Surfaces<float>) from flat arrays and
// calls the reference's loop nests, compiled from the sources where they lie under
// /root/reference/src (never copied into this repo):
//     points_affect_points<float,double>   src/Influence.h:67-551
//     panels_affect_points<float,double>   src/Influence.h:557-1099
//     points_affect_panels<float,double>   src/Influence.h:1107-1221
//     panels_on_panels_coeff<float>        src/Coefficients.h:169-483
// with ExecEnv(true,true,direct,cpu_x86), i.e. the scalar "float kernel, double accumulator"
// arm (src/Simulation.h:41-47 without USE_VC). The result, oracle/_ref/libo3d_ref.so, is the
// parity oracle for tests/ and the CPU baseline for bench.py. Nothing in omega3d_b200/
// may link or load it.
//
// Build: see oracle/Makefile (target _ref/libo3d_ref.so).

#ifndef VERBOSE
#define VERBOSE false  // CMake normally passes -DVERBOSE (CMakeLists.txt:31-35)
#endif
#include "Collection.h"  // must precede Influence.h (ElementBase.h -> GlComputeState.h -> Collection.h cycle)
#include "Influence.h"
#include "Coefficients.h"
#include <chrono>
#include "Reflect.h"
#include "StatusFile.h"          // src/StatusFile.cpp is compiled where it lies (oracle/Makefile)
#include "RHS.h"                 // vels_to_rhs_panels: the projection solve_bem applies (src/BEMHelper.h:103)
#ifdef USE_CUDA
#include "O3DCudaConvection.h"   // what the patched Convection.h includes (integration/omega3d_use_cuda.patch)
#endif

#include <omp.h>
#include <unistd.h>
#include <fcntl.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

// Body.cpp needs the real Eigen and is not built. ElementBase::move references these two members for
// body-bound collections; the oracle never attaches a Body (always nullptr), so they are link-time stand-ins
// that abort if anything ever reaches them.
void Body::transform(const double) { std::abort(); }
Trans Body::get_transform_mat() { std::abort(); }
std::string Body::get_name() { std::abort(); }      // (Surfaces::to_string, only with a Body attached)

namespace {

// The reference chats on stdout/stderr from every routine and ctor; mute fds 1 and 2 for
// the duration of a call so harness output (one JSON line) stays clean.
struct Mute {
  int so = -1, se = -1;
  bool on;
  explicit Mute(bool enable) : on(enable) {
    if (!on) return;
    fflush(stdout); fflush(stderr); std::cout.flush(); std::cerr.flush();
    so = dup(1); se = dup(2);
    int nul = open("/dev/null", O_WRONLY);
    dup2(nul, 1); dup2(nul, 2); close(nul);
  }
  ~Mute() {
    if (!on) return;
    fflush(stdout); fflush(stderr); std::cout.flush(); std::cerr.flush();
    dup2(so, 1); dup2(se, 2); close(so); close(se);
  }
};

bool g_mute = true;
int g_device_convect = 1;  // drop-in build only: route whole convection steps to the device arm
int g_accel = cpu_x86;  // accel_t handed to ExecEnv; gpu_cuda only means something in the -DUSE_CUDA (drop-in) build

Points<float> make_points(int n, const float* x, const float* y, const float* z,
                          const float* str3 /*SoA 3 x n or NULL*/, const float* rad /*or NULL*/,
                          elem_t e, move_t m) {
  std::vector<float> px(3 * (size_t)n), val;
  for (int i = 0; i < n; ++i) { px[3*i] = x[i]; px[3*i+1] = y[i]; px[3*i+2] = z[i]; }
  if (e != inert) {
    val.resize(3 * (size_t)n);
    for (int i = 0; i < n; ++i) for (int d = 0; d < 3; ++d) val[3*i+d] = str3[(size_t)d*n + i];
  }
  ElementPacket<float> pk(px, std::vector<Int>(), val, (size_t)n, 0);
  Points<float> p(pk, e, m, nullptr, 0.0f);
  if (e != inert && rad) {
    Vector<float>& r = p.get_rad();
    for (int i = 0; i < n; ++i) r[i] = rad[i];
  }
  return p;
}

Surfaces<float> make_surfaces(int nn, const float* nodes /*3*nn interleaved*/, int np,
                              const uint32_t* idx /*3*np*/, const float* val /*3*np interleaved*/,
                              elem_t e) {
  std::vector<float> x(nodes, nodes + 3 * (size_t)nn);
  std::vector<Int> id(idx, idx + 3 * (size_t)np);
  std::vector<float> v(val, val + 3 * (size_t)np);
  ElementPacket<float> pk(x, id, v, (size_t)np, 2);
  return Surfaces<float>(pk, e, fixed, nullptr);
}

// target kinds understood by the *_pts entry points
//   0: inert + fixed      -> no radius, HAS velgrad storage   (field points)
//   1: inert + lagrangian -> no radius, NO velgrad storage    (tracers)
//   2: active + lagrangian-> radius and velgrad storage       (vortex particles)
Points<float> make_targets(int kind, int nt, const float* tx, const float* ty, const float* tz,
                           const float* tr) {
  if (kind == 2) {
    std::vector<float> zero(3 * (size_t)nt, 0.0f);
    return make_points(nt, tx, ty, tz, zero.data(), tr, active, lagrangian);
  }
  return make_points(nt, tx, ty, tz, nullptr, nullptr, inert, kind == 0 ? fixed : lagrangian);
}

void load_results(Points<float>& t, int nt, const float* tu, const float* tug) {
  auto& u = t.get_vel();
  for (int d = 0; d < 3; ++d) std::memcpy(u[d].data(), tu + (size_t)d*nt, sizeof(float)*nt);
  auto& og = t.get_velgrad();
  if (og && tug) for (int d = 0; d < 9; ++d) std::memcpy((*og)[d].data(), tug + (size_t)d*nt, sizeof(float)*nt);
}
void store_results(Points<float>& t, int nt, float* tu, float* tug) {
  auto& u = t.get_vel();
  for (int d = 0; d < 3; ++d) std::memcpy(tu + (size_t)d*nt, u[d].data(), sizeof(float)*nt);
  auto& og = t.get_velgrad();
  if (og && tug) for (int d = 0; d < 9; ++d) std::memcpy(tug + (size_t)d*nt, (*og)[d].data(), sizeof(float)*nt);
}

}  // namespace

extern "C" {

void o3d_ref_set_mute(int on) { g_mute = on != 0; }
// 1 = cpu_x86 (the oracle), 4 = gpu_cuda (dispatches into integration/O3DCudaInfluence.h when built with
// the patched headers and -DUSE_CUDA: oracle/_ref/libo3d_dropin.so)
void o3d_ref_set_accel(int a) { g_accel = a; }
void o3d_ref_set_device_convect(int on) { g_device_convect = on; }
int o3d_ref_built_with_cuda() {
#ifdef USE_CUDA
  return 1;
#else
  return 0;
#endif
}
int o3d_ref_max_threads() { return omp_get_max_threads(); }
void o3d_ref_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }

// src (ns particles, SoA) -> targets; results ACCUMULATE into tu (3 x nt SoA) and tug (9 x nt SoA).
// want_grad selects ResultsType velandgrad / velonly. Returns 0, or -1 on a combination the
// reference itself asserts on (src/Influence.h:368-370).
int o3d_ref_pts_on_pts(int ns, const float* sx, const float* sy, const float* sz, const float* sr,
                       const float* ss3, int targ_kind, int nt, const float* tx, const float* ty,
                       const float* tz, const float* tr, int want_grad, float* tu, float* tug) {
  if (targ_kind == 1 && want_grad) return -1;
  Mute m(g_mute);
  Points<float> src = make_points(ns, sx, sy, sz, ss3, sr, active, lagrangian);
  Points<float> targ = make_targets(targ_kind, nt, tx, ty, tz, tr);
  load_results(targ, nt, tu, tug);
  ExecEnv env(true, true, direct, (accel_t)g_accel);
  points_affect_points<float, double>(src, targ, ResultsType(want_grad ? velandgrad : velonly), env);
  store_results(targ, nt, tu, tug);
  return 0;
}

// Derived panel quantities exactly as the reference's Surfaces ctor computes them
// (compute_bases, vortex_sheet_to_panel_strength: src/Surfaces.h:309-335).
// val = (vortex-sheet strength along x1, along x2, source-sheet strength) per panel.
// Outputs (all SoA): area[np], ts[3*np] total vortex strength, b1/b2/nrm[3*np] basis vectors.
int o3d_ref_surface_props(int nn, const float* nodes, int np, const uint32_t* idx, const float* val,
                          float* area, float* ts, float* b1, float* b2, float* nrm) {
  Mute m(g_mute);
  Surfaces<float> s = make_surfaces(nn, nodes, np, idx, val, active);
  for (int i = 0; i < np; ++i) area[i] = s.get_area()[i];
  for (int d = 0; d < 3; ++d) for (int i = 0; i < np; ++i) {
    ts[(size_t)d*np + i]  = s.get_str()[d][i];
    b1[(size_t)d*np + i]  = s.get_x1()[d][i];
    b2[(size_t)d*np + i]  = s.get_x2()[d][i];
    nrm[(size_t)d*np + i] = s.get_norm()[d][i];
  }
  return 0;
}

// panels -> points. Gradients are produced iff the target kind has velgrad storage
// (src/Influence.h:652,875), independent of restype.
int o3d_ref_pan_on_pts(int nn, const float* nodes, int np, const uint32_t* idx, const float* val,
                       int targ_kind, int nt, const float* tx, const float* ty, const float* tz,
                       const float* tr, float* tu, float* tug) {
  Mute m(g_mute);
  Surfaces<float> src = make_surfaces(nn, nodes, np, idx, val, active);
  Points<float> targ = make_targets(targ_kind, nt, tx, ty, tz, tr);
  load_results(targ, nt, tu, tug);
  ExecEnv env(true, true, direct, (accel_t)g_accel);
  panels_affect_points<float, double>(src, targ, ResultsType(velonly), env);
  store_results(targ, nt, tu, tug);
  return 0;
}

// particles -> panel centres (BEM right-hand side). pu (3 x np SoA) is DECREMENTED
// (src/Influence.h:1210-1212).
int o3d_ref_pts_on_pan(int ns, const float* sx, const float* sy, const float* sz, const float* sr,
                       const float* ss3, int nn, const float* nodes, int np, const uint32_t* idx,
                       const float* val, float* pu) {
  Mute m(g_mute);
  Points<float> src = make_points(ns, sx, sy, sz, ss3, sr, active, lagrangian);
  Surfaces<float> targ = make_surfaces(nn, nodes, np, idx, val, active);
  auto& u = targ.get_vel();
  for (int d = 0; d < 3; ++d) std::memcpy(u[d].data(), pu + (size_t)d*np, sizeof(float)*np);
  ExecEnv env(true, true, direct, (accel_t)g_accel);
  points_affect_panels<float, double>(src, targ, ResultsType(velonly), env);
  for (int d = 0; d < 3; ++d) std::memcpy(pu + (size_t)d*np, u[d].data(), sizeof(float)*np);
  return 0;
}

// panels -> panels velocity via the reference's colocation-point trick (src/Influence.h:1224-1245).
int o3d_ref_pan_on_pan(int nn, const float* nodes, int np, const uint32_t* idx, const float* val,
                       int tnn, const float* tnodes, int tnp, const uint32_t* tidx, const float* tval,
                       float* pu) {
  Mute m(g_mute);
  Surfaces<float> src = make_surfaces(nn, nodes, np, idx, val, active);
  Surfaces<float> targ = make_surfaces(tnn, tnodes, tnp, tidx, tval, reactive);
  ExecEnv env(true, true, direct, (accel_t)g_accel);
  panels_affect_panels<float, double>(src, targ, ResultsType(velonly), env);
  auto& u = targ.get_vel();
  for (int d = 0; d < 3; ++d) std::memcpy(pu + (size_t)d*tnp, u[d].data(), sizeof(float)*tnp);
  return 0;
}

// BEM influence block of a reactive surface on itself (self != 0) or on a second surface.
// Returns the number of floats written to coeffs (column-major, (nunk*ntarg) x (nunk*nsrc)),
// or the required size when coeffs == NULL.
long o3d_ref_pan_on_pan_coeff(int nn, const float* nodes, int np, const uint32_t* idx, const float* bc,
                              int self, int tnn, const float* tnodes, int tnp, const uint32_t* tidx,
                              const float* tbc, float* coeffs) {
  Mute m(g_mute);
  Surfaces<float> src = make_surfaces(nn, nodes, np, idx, bc, reactive);
  if (self) {
    Vector<float> c = panels_on_panels_coeff<float>(src, src);
    if (coeffs) std::memcpy(coeffs, c.data(), sizeof(float)*c.size());
    return (long)c.size();
  }
  Surfaces<float> targ = make_surfaces(tnn, tnodes, tnp, tidx, tbc, reactive);
  Vector<float> c = panels_on_panels_coeff<float>(src, targ);
  if (coeffs) std::memcpy(coeffs, c.data(), sizeof(float)*c.size());
  return (long)c.size();
}

// Single-interaction known-answer hooks straight into src/Kernels.h (double accumulators).
void o3d_ref_kernel_0v_0bg(const float* s7 /*x y z r wx wy wz*/, const float* t4 /*x y z r*/, double* out12) {
  for (int i = 0; i < 12; ++i) out12[i] = 0.0;
  kernel_0v_0bg<float, double>(s7[0], s7[1], s7[2], s7[3], s7[4], s7[5], s7[6], t4[0], t4[1], t4[2], t4[3],
      out12+0, out12+1, out12+2, out12+3, out12+4, out12+5, out12+6, out12+7, out12+8, out12+9, out12+10, out12+11);
}
int o3d_ref_rkernel_2vs_0p(const float* tri9 /*x0 y0 z0 x1 ..*/, const float* str4, const float* t3,
                           float sa, double* out3) {
  out3[0] = out3[1] = out3[2] = 0.0;
  return rkernel_2vs_0p<float, double>(tri9[0], tri9[1], tri9[2], tri9[3], tri9[4], tri9[5], tri9[6], tri9[7],
      tri9[8], str4[0], str4[1], str4[2], str4[3], t3[0], t3[1], t3[2], sa, 0, RECURSIVE_LEVELS,
      out3, out3+1, out3+2);
}
int o3d_ref_rkernel_2vs_0pg(const float* tri9, const float* str4, const float* t3, float sa, double* out12) {
  for (int i = 0; i < 12; ++i) out12[i] = 0.0;
  return rkernel_2vs_0pg<float, double>(tri9[0], tri9[1], tri9[2], tri9[3], tri9[4], tri9[5], tri9[6], tri9[7],
      tri9[8], str4[0], str4[1], str4[2], str4[3], t3[0], t3[1], t3[2], sa, 0, RECURSIVE_LEVELS,
      out12+0, out12+1, out12+2, out12+3, out12+4, out12+5, out12+6, out12+7, out12+8, out12+9, out12+10, out12+11);
}

}  // extern "C"

// ---- convection: the reference's own Points<float>::zero_vels / finalize_vels / move (src/Points.h:252-520,
// src/ElementBase.h:170-336) driven in the order Convection<S,A,I>::advect_* drives them for a system with no
// boundaries and no field points (src/Convection.h:130-184 find_vels, :232-262 advect_1st, :349-425
// advect_2nd_ralston, :431-556 advect_3rd). Convection.h itself cannot be compiled here (it includes BEM.h, which
// needs the real Eigen), so only that call ORDER is restated below; every arithmetic step is the reference's.
namespace {

void load_state(Points<float>& p, int n, const float* x, const float* s, const float* elong) {
  auto& px = p.get_pos();
  for (int d = 0; d < 3; ++d) std::memcpy(px[d].data(), x + (size_t)d*n, sizeof(float)*n);
  if (s) { auto& ps = p.get_str(); for (int d = 0; d < 3; ++d) std::memcpy(ps[d].data(), s + (size_t)d*n, sizeof(float)*n); }
  if (elong) std::memcpy(p.get_elong().data(), elong, sizeof(float)*n);
}
void store_state(Points<float>& p, int n, float* x, float* s, float* elong) {
  auto& px = p.get_pos();
  for (int d = 0; d < 3; ++d) std::memcpy(x + (size_t)d*n, px[d].data(), sizeof(float)*n);
  if (s) { auto& ps = p.get_str(); for (int d = 0; d < 3; ++d) std::memcpy(s + (size_t)d*n, ps[d].data(), sizeof(float)*n); }
  if (elong) std::memcpy(elong, p.get_elong().data(), sizeof(float)*n);
}

// Convection::find_vels(_fs, _vort, _bdry = {}, _targets = _vort) for one collection
void find_vels_one(Points<float>& p, const std::array<double,3>& fs, const ExecEnv& env) {
  p.zero_vels();
  points_affect_points<float, double>(p, p, ResultsType(velandgrad), env);
  p.finalize_vels(fs);
}

void advect_once(Points<float>& vort, int order, double time, double dt, const std::array<double,3>& fs, const ExecEnv& env) {
#ifdef USE_CUDA
  // the arm integration/omega3d_use_cuda.patch adds to Convection::advect (src/Convection.h:218): one particle
  // collection and nothing else -> the whole step on the device. g_device_convect = 0 keeps the reference's host
  // sequencing below with only the influence sums on the GPU (the Influence.h arms), for comparison.
  if (g_device_convect && env.is_internal() && env.get_instrs() == gpu_cuda &&
      vort.get_elemt() == active && vort.get_movet() == lagrangian) {
    (void) o3d::cuda_advect_particles(vort, order, time, dt, fs);
    return;
  }
#endif
  find_vels_one(vort, fs, env);
  if (order == 1) {
    vort.move(time, dt, 1.0, vort);
  } else if (order == 2) {
    const double twothirds = 2.0/3.0;
    Points<float> interim = vort;
    interim.move(time, twothirds*dt, 1.0, interim);
    find_vels_one(interim, fs, env);
    vort.move(time, dt, 0.25, vort, 0.75, interim);
  } else {
    Points<float> v1 = vort;
    v1.move(time, 0.5*dt, 1.0, v1);
    find_vels_one(v1, fs, env);
    Points<float> v2 = vort;
    v2.move(time, 0.75*dt, 1.0, v1);
    find_vels_one(v2, fs, env);
    vort.move(time, dt, 2.0/9.0, vort, 3.0/9.0, v1, 4.0/9.0, v2);
  }
}

}  // namespace

extern "C" {

// Points::finalize_vels on flat arrays: u (3 x n), ug (9 x n or NULL)
void o3d_ref_finalize_vels(int n, float* u, float* ug, const double* fs) {
  Mute m(g_mute);
  std::vector<float> zx(n, 0.0f), zs(3*(size_t)n, 0.0f);
  Points<float> p = make_points(n, zx.data(), zx.data(), zx.data(), zs.data(), nullptr, active, lagrangian);
  load_results(p, n, u, ug);
  if (!ug) p.get_velgrad().reset();
  p.finalize_vels({fs[0], fs[1], fs[2]});
  store_results(p, n, u, ug);
}

// Points::move with `order` stages. State x, s (3 x n), elong (n) in/out. stage k: uk (3 x n), ugk (9 x n or NULL).
// For order 1 the reference stretches with the moving object's OWN gradient: pass it as ug0 (u0 is the velocity of
// the object handed to move(), which may be another collection - advect_3rd does exactly that).
// uout (3 x n, may be NULL): this->u after the call (order >= 2 overwrites it with the combined velocity).
void o3d_ref_move(int order, int n, double dt, const double* wt, const float* u0, const float* ug0, const float* u1,
                  const float* ug1, const float* u2, const float* ug2, float* x, float* s, float* elong, float* uout) {
  Mute m(g_mute);
  std::vector<float> zx(n, 0.0f), zs(3*(size_t)n, 0.0f);
  Points<float> self = make_points(n, zx.data(), zx.data(), zx.data(), zs.data(), nullptr, active, lagrangian);
  load_state(self, n, x, s, elong);
  auto stage = [&](const float* u, const float* ug) {
    Points<float> p = make_points(n, zx.data(), zx.data(), zx.data(), zs.data(), nullptr, active, lagrangian);
    load_results(p, n, u, ug);
    if (!ug) p.get_velgrad().reset();
    return p;
  };
  if (order == 1) {
    Points<float> a = stage(u0, nullptr);
    if (ug0) load_results(self, n, u0, ug0); else self.get_velgrad().reset();
    self.move(0.0, dt, wt[0], a);
  } else if (order == 2) {
    load_results(self, n, u0, ug0);           // _u1 is *this in the reference's calls; keep that aliasing
    if (!ug0) self.get_velgrad().reset();
    Points<float> b = stage(u1, ug1);
    self.move(0.0, dt, wt[0], self, wt[1], b);
  } else {
    load_results(self, n, u0, ug0);
    if (!ug0) self.get_velgrad().reset();
    Points<float> b = stage(u1, ug1), c = stage(u2, ug2);
    self.move(0.0, dt, wt[0], self, wt[1], b, wt[2], c);
  }
  store_state(self, n, x, s, elong);
  if (uout) { auto& u = self.get_vel(); for (int d = 0; d < 3; ++d) std::memcpy(uout + (size_t)d*n, u[d].data(), sizeof(float)*n); }
}

// nsteps x Convection::advect(order) on one vortex-particle collection. x, s (3 x n), r, elong (n) in/out;
// u (3 x n), ug (9 x n) out = what the collection holds afterwards. With the drop-in build and accel 4 the
// influence sums run through the CUDA arm of the patched points_affect_points.
void o3d_ref_advect(int order, int nsteps, double dt, const double* fs, int n, float* x, float* s, const float* r,
                    float* elong, float* u, float* ug) {
  Mute m(g_mute);
  Points<float> vort = make_points(n, x, x + n, x + 2*(size_t)n, s, r, active, lagrangian);
  load_state(vort, n, x, s, elong);
  ExecEnv env(true, true, direct, (accel_t)g_accel);
  const std::array<double,3> f = {fs[0], fs[1], fs[2]};
  double time = 0.0;
  for (int k = 0; k < nsteps; ++k) { advect_once(vort, order, time, dt, f, env); time += dt; }
  store_state(vort, n, x, s, elong);
  store_results(vort, n, u, ug);
}

// One Convection::advect step of a particle collection around ONE static body, the way the patched Convection.h dispatches it
// (integration/omega3d_use_cuda.patch): with the drop-in build and accel 4, o3d::cuda_advect_particles_body keeps the particles
// on the device and calls `rest` once per derivative evaluation. `rest` here is solve_bem (src/BEMHelper.h:44-262) line for
// line for one surface - zero_vels, points_affect_panels through the reference's own dispatch (whose gpu_cuda arm delivers the
// device's sums), finalize_vels, vels_to_rhs_panels, set_str - EXCEPT the linear solve, for which BEM.h needs Eigen: the caller's
// `solve(rhs, n, strengths)` stands in for BEM::solve. Without CUDA (accel 1) the same `rest` runs inside the reference's own
// host sequencing: find_derivs, move, clear_inner_layer (src/Convection.h:232-425), orders 1 and 2.
typedef void (*o3d_ref_dense_solve_fn)(const float* rhs, int n, float* strengths);
long o3d_ref_advect_body(int order, double dt, const double* fs, float ips, int n, float* x, float* s, const float* r, float* elong,
                         int nn, const float* nodes, int np, const uint32_t* idx, o3d_ref_dense_solve_fn solve, float* ts_out) {
  Mute m(g_mute);
  Points<float> vort = make_points(n, x, x + n, x + 2*(size_t)n, s, r, active, lagrangian);
  load_state(vort, n, x, s, elong);
  std::vector<float> bc(3 * (size_t)np, 0.0f);
  Surfaces<float> surf = make_surfaces(nn, nodes, np, idx, bc.data(), reactive);
  ExecEnv env(true, true, direct, (accel_t)g_accel);
  const std::array<double,3> f = {fs[0], fs[1], fs[2]};
  const float cut = 0.5/std::sqrt(2.0*M_PI);
  long solves = 0;
  auto rest_for = [&](Points<float>& state) {
    surf.zero_vels();
    points_affect_panels<float, double>(state, surf, ResultsType(velonly), env);
    surf.finalize_vels(f);
    std::vector<float> rhs = vels_to_rhs_panels<float>(surf);
    Vector<float> sol(rhs.size());
    solve(rhs.data(), (int)rhs.size(), sol.data());
    surf.set_str(0, sol.size(), sol);
    ++solves;
  };
  bool on_device = false;
#ifdef USE_CUDA
  if (g_device_convect && env.get_instrs() == gpu_cuda) {
    auto rest = [&]() { rest_for(vort); };     // (vort's host arrays are stale here: the gpu_cuda arm never reads them)
    (void) o3d::cuda_advect_particles_body(vort, surf, order, 0.0, dt, f, ips, rest);
    on_device = true;
  }
#endif
  if (!on_device) {
    auto derivs = [&](Points<float>& state) {
      rest_for(state);
      state.zero_vels();
      points_affect_points<float, double>(state, state, ResultsType(velandgrad), env);
      panels_affect_points<float, double>(surf, state, ResultsType(velandgrad), env);
      state.finalize_vels(f);
    };
    derivs(vort);
    if (order == 1) {
      vort.move(0.0, dt, 1.0, vort);
    } else {
      Points<float> interim = vort;
      interim.move(0.0, (2.0/3.0)*dt, 1.0, interim);
      (void) clear_inner_panp2<float>(1, surf, interim, cut, ips);
      derivs(interim);
      vort.move(0.0, dt, 0.25, vort, 0.75, interim);
    }
    (void) clear_inner_panp2<float>(1, surf, vort, cut, ips);
  }
  store_state(vort, n, x, s, elong);
  if (ts_out) for (int d = 0; d < 3; ++d) for (int i = 0; i < np; ++i) ts_out[(size_t)d*np + i] = surf.get_str()[d][i];
  return solves;
}

// ElementBase::get_max_str and Points::get_max_elong
void o3d_ref_stats(int n, const float* s, const float* elong, float* max_str, float* max_elong) {
  Mute m(g_mute);
  std::vector<float> zx(n, 0.0f);
  Points<float> p = make_points(n, zx.data(), zx.data(), zx.data(), s, nullptr, active, lagrangian);
  std::memcpy(p.get_elong().data(), elong, sizeof(float)*n);
  *max_str = p.get_max_str();
  *max_elong = p.get_max_elong();
}

// ElementBase::get_total_circ (src/ElementBase.h:354-378) and Points::get_total_impulse (src/Points.h:547-563)
void o3d_ref_totals(int n, const float* x /*3 x n*/, const float* s /*3 x n*/, float* circ, float* impulse) {
  Mute m(g_mute);
  Points<float> p = make_points(n, x, x + n, x + 2 * (size_t)n, s, nullptr, active, lagrangian);
  const std::array<float,3> c = p.get_total_circ(0.0);
  const std::array<float,3> i = p.get_total_impulse();
  for (int d = 0; d < 3; ++d) { circ[d] = c[d]; impulse[d] = i[d]; }
}

// The reference's StatusFile driven the way Simulation::dump_stats_to_status drives it (src/Simulation.cpp:851-897): `nlines`
// lines of (time, Nv, gx gy gz, fx fy fz) appended to `path`; reset_before[k] != 0 calls reset_sim() before line k (a new run
// in the same process: src/Simulation.cpp:575). vals: nlines x 7 floats (time, g[3], f[3]); nv: nlines ints.
int o3d_ref_status_lines(const char* path, int csv_format, int nlines, const float* vals, const int* nv, const int* reset_before) {
  StatusFile sf;
  sf.set_filename(path);
  sf.format = csv_format ? csv : dat;
  for (int k = 0; k < nlines; ++k) {
    if (reset_before && reset_before[k]) sf.reset_sim();
    const float* v = vals + 7 * (size_t)k;
    sf.append_value("time", v[0]);
    sf.append_value("Nv", nv[k]);
    sf.append_value("gx", v[1]); sf.append_value("gy", v[2]); sf.append_value("gz", v[3]);
    sf.append_value("fx", v[4]); sf.append_value("fy", v[5]); sf.append_value("fz", v[6]);
    sf.write_line();
  }
  return 0;
}

}  // extern "C"

// ---- initial conditions of the example cases, from the reference's own feature generators ------------
// (src/FlowFeature.cpp, compiled where it lies; its three GUI draw-geometry helpers live in GeomHelper.cpp, which
// needs igl/Eigen, so they are link-time stand-ins below - init_elements never calls them).
#ifdef O3D_REF_WITH_FEATURES
#include "FlowFeature.h"
#include "GeomHelper.h"
ElementPacket<float> generate_ovoid(const float, const float, const float, const float) { return ElementPacket<float>(); }
ElementPacket<float> generate_cuboid(const float, const float, const float, const float) { return ElementPacket<float>(); }
ElementPacket<float> generate_torus(const float, const float, const float) { return ElementPacket<float>(); }

namespace {
long emit_packet(const ElementPacket<float>& pk, float* x, float* s, long cap) {
  const long n = (long)pk.nelem;
  if (!x || !s || cap < n) return n;
  for (long i = 0; i < n; ++i)
    for (int d = 0; d < 3; ++d) { x[(size_t)d*n + i] = pk.x[3*i + d]; s[(size_t)d*n + i] = pk.val[3*i + d]; }
  return n;
}
}  // namespace

extern "C" {
int o3d_ref_has_features() { return 1; }
// SingularRing::init_elements(ips) (src/FlowFeature.cpp:789-832) -> SoA x, s (3 x n). Returns n; call with
// x == NULL to size the arrays.
long o3d_ref_singular_ring(const float* c3, const float* n3, float majrad, float circ, float ips, float* x, float* s, long cap) {
  Mute m(g_mute);
  SingularRing r(c3[0], c3[1], c3[2], n3[0], n3[1], n3[2], majrad, circ);
  return emit_packet(r.init_elements(ips), x, s, cap);
}
// ThickRing::init_elements(ips) (src/FlowFeature.cpp:941-1019)
long o3d_ref_thick_ring(const float* c3, const float* n3, float majrad, float minrad, float circ, float ips, float* x, float* s, long cap) {
  Mute m(g_mute);
  ThickRing r(c3[0], c3[1], c3[2], n3[0], n3[1], n3[2], majrad, minrad, circ);
  return emit_packet(r.init_elements(ips), x, s, cap);
}
}  // extern "C"
#else
extern "C" int o3d_ref_has_features() { return 0; }
#endif

// ---- particle x panel closest-point loops: the reference's reflect_panp2 / clear_inner_panp2 (src/Reflect.h:194-311,
// :446-620) on its real containers. x is 3 x nt SoA, updated in place. Returns the number of particles whose position
// changed (the reference only prints its count).
extern "C" {
long o3d_ref_reflect(int nn, const float* nodes, int np, const uint32_t* idx, int nt, float* x) {
  Mute m(g_mute);
  std::vector<float> val(3 * (size_t)np, 0.0f), zs(3 * (size_t)nt, 0.0f);
  Surfaces<float> surf = make_surfaces(nn, nodes, np, idx, val.data(), reactive);
  Points<float> pts = make_points(nt, x, x + nt, x + 2 * (size_t)nt, zs.data(), nullptr, active, lagrangian);
  reflect_panp2<float>(surf, pts);
  long moved = 0;
  auto& px = pts.get_pos();
  for (int i = 0; i < nt; ++i) {
    bool ch = false;
    for (int d = 0; d < 3; ++d) { ch = ch || px[d][i] != x[(size_t)d*nt + i]; }
    moved += ch;
  }
  for (int d = 0; d < 3; ++d) std::memcpy(x + (size_t)d*nt, px[d].data(), sizeof(float)*nt);
  return moved;
}
long o3d_ref_clear_inner(int method, int nn, const float* nodes, int np, const uint32_t* idx, int nt, float* x,
                         const float* rad, float cutoff_mult, float ips) {
  Mute m(g_mute);
  std::vector<float> val(3 * (size_t)np, 0.0f), zs(3 * (size_t)nt, 0.0f);
  Surfaces<float> surf = make_surfaces(nn, nodes, np, idx, val.data(), reactive);
  Points<float> pts = make_points(nt, x, x + nt, x + 2 * (size_t)nt, zs.data(), rad, active, lagrangian);
  clear_inner_panp2<float>(method, surf, pts, cutoff_mult, ips);
  long moved = 0;
  auto& px = pts.get_pos();
  for (int i = 0; i < nt; ++i) {
    bool ch = false;
    for (int d = 0; d < 3; ++d) { ch = ch || px[d][i] != x[(size_t)d*nt + i]; }
    moved += ch;
  }
  for (int d = 0; d < 3; ++d) std::memcpy(x + (size_t)d*nt, px[d].data(), sizeof(float)*nt);
  return moved;
}
}  // extern "C"

// ---- Points<float>::write_vtk (src/Points.h:851-1039): writes part_<index>_<frame>.vtu into `dir`. Returns 0.
extern "C" int o3d_ref_write_vtk(int n, const float* x, const float* s, const float* r, const float* u, int index, int frameno,
                                 double time, const char* dir) {
  Mute m(g_mute);
  Points<float> p = make_points(n, x, x + n, x + 2 * (size_t)n, s, r, active, lagrangian);
  auto& pu = p.get_vel();
  for (int d = 0; d < 3; ++d) std::memcpy(pu[d].data(), u + (size_t)d*n, sizeof(float)*n);
  char cwd[4096];
  if (!getcwd(cwd, sizeof cwd) || chdir(dir) != 0) return -1;
  p.write_vtk((size_t)index, (size_t)frameno, time);
  return chdir(cwd);
}

/* oracle/biot_oracle.h - plain-C CPU restatement of Omega3D's direct Biot-Savart hot path.
 *
 * TEST INFRASTRUCTURE ONLY: used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * as the CHECKER. Nothing under omega3d_b200/ may include, link or load it.
 *
 * Parity status: PINNED. The reference ships no tests or golden vectors, so the
 * restatement is pinned against outputs of the reference's own templates compiled here from
 * /root/reference/src (oracle/ref_driver.cpp -> oracle/_ref/libo3d_ref.so): bit-identical on every
 * committed fixture in tests/golden/ (tests/test_oracle.py), both built with -ffp-contract=off.
 *
 * Arithmetic scheme = the reference's non-Vc build (src/Simulation.h:41-47): every pairwise
 * quantity in float, in the reference's operation order; per-target sums in double for the
 * influence routines, float for the BEM coefficient block (rkernel_2vs_2p<S,S>, src/Coefficients.h:356).
 * All arrays are SoA float32; outputs ACCUMULATE into the caller's arrays, un-normalised (no 1/4pi).
 */
#ifndef O3D_BIOT_ORACLE_H
#define O3D_BIOT_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* single interactions (src/Kernels.h:50-345), double accumulators */
void o3d_oracle_kernel_0v_0b(const float s[7], const float t[4], double u[3]);
void o3d_oracle_kernel_0v_0p(const float s[7], const float t[3], double u[3]);
void o3d_oracle_kernel_0v_0bg(const float s[7], const float t[4], double out[12]);
void o3d_oracle_kernel_0v_0pg(const float s[7], const float t[3], double out[12]);
/* recursive panel -> point (src/Kernels.h:1028-1211); returns the reference's flop count */
int o3d_oracle_rkernel_2vs_0p(const float tri[9], const float str[4], const float t[3], float sa, double u[3]);
int o3d_oracle_rkernel_2vs_0pg(const float tri[9], const float str[4], const float t[3], float sa, double out[12]);

/* particles -> points (src/Influence.h:67-551). tr == NULL: singular targets (kernel_0v_0p[g]);
 * tug == NULL: velocity only. tu is 3 x nt, tug is 9 x nt (row k = d u_{k%3} / d x_{k/3}). */
void o3d_oracle_pts_on_pts(int64_t ns, const float* sx, const float* sy, const float* sz, const float* sr,
                           const float* ssx, const float* ssy, const float* ssz,
                           int64_t nt, const float* tx, const float* ty, const float* tz, const float* tr,
                           float* tu, float* tug);

/* the same for a build of the reference with another core function (src/CoreFunc.h:35-38):
 * core 0 Winckelmans-Leonard (shipped), 1 Rosenhead-Moore, 2 exponential, 3 Vatistas n=2. */
void o3d_oracle_pts_on_pts_core(int core, int64_t ns, const float* sx, const float* sy, const float* sz, const float* sr,
                                const float* ssx, const float* ssy, const float* ssz,
                                int64_t nt, const float* tx, const float* ty, const float* tz, const float* tr,
                                float* tu, float* tug);

/* panels -> points (src/Influence.h:557-1099). nodes SoA (nx,ny,nz), idx 3 per panel, ts = total
 * vortex strength SoA 3 x np, area, sss = source-sheet strength or NULL. */
void o3d_oracle_pan_on_pts(int64_t np, const float* nx, const float* ny, const float* nz, const uint32_t* idx,
                           const float* ts, const float* area, const float* sss,
                           int64_t nt, const float* tx, const float* ty, const float* tz,
                           float* tu, float* tug);

/* particles -> panel centres, SUBTRACTED from pu (src/Influence.h:1107-1221). */
void o3d_oracle_pts_on_pan(int64_t ns, const float* sx, const float* sy, const float* sz,
                           const float* ssx, const float* ssy, const float* ssz,
                           int64_t np, const float* nx, const float* ny, const float* nz, const uint32_t* idx,
                           const float* area, float* pu);

/* BEM influence block (src/Coefficients.h:169-483), nunk = 3 unknowns per panel, column-major
 * (3*ntp) x (3*nsp); self != 0 applies the diagonal override (:414-436). b1/b2/nrm are SoA 3 x n. */
void o3d_oracle_pan_on_pan_coeff(int64_t nsp, const float* snx, const float* sny, const float* snz,
                                 const uint32_t* sidx, const float* sb1, const float* sb2, const float* sarea,
                                 int64_t ntp, const float* tnx, const float* tny, const float* tnz,
                                 const uint32_t* tidx, const float* tb1, const float* tb2, const float* tnrm,
                                 const float* tarea, int self, float* coeffs);

/* ---- convection: finalize_vels, Points::move, Convection::advect for a lone particle collection ---- */
/* src/ElementBase.h:187-192, src/Points.h:265-277. u 3 x n, ug 9 x n or NULL. */
void o3d_oracle_finalize_vels(int64_t n, float* u, float* ug, const double* fs);
/* Points::move, 1-3 stages (src/ElementBase.h:253-336, src/Points.h:288-520). u[k] 3 x n, ug[k] 9 x n or NULL. */
void o3d_oracle_move(int order, int64_t n, double dt, const double* wt, const float* const* u, const float* const* ug,
                     float* x, float* s, float* elong, float* uout);
/* nsteps x Convection::advect (src/Convection.h:232-262, :349-425, :431-556) with no boundaries / field points. */
void o3d_oracle_advect(int order, int nsteps, double dt, const double* fs, int64_t n, float* x, float* s, const float* r,
                       float* elong, float* u, float* ug);
/* core function used by o3d_oracle_advect's evaluations (0 = Winckelmans-Leonard, the default; see o3d_oracle_pts_on_pts_core) */
void o3d_oracle_set_advect_core(int core);
void o3d_oracle_stats(int64_t n, const float* s, const float* elong, float* max_str, float* max_elong);
/* ElementBase::get_total_circ (src/ElementBase.h:354-378) and Points::get_total_impulse (src/Points.h:547-563); x, s: 3 x n */
void o3d_oracle_totals(int64_t n, const float* x, const float* s, float* circ, float* impulse);

/* ---- particle x panel closest-point loops: reflect_panp2 (mode 0, src/Reflect.h:194-311) and clear_inner_panp2 with
 * _method 1 (mode 1, :446-620). nodes SoA, idx 3 per panel, nrm SoA 3 x np, x 3 x nt in/out. Returns particles moved. */
int64_t o3d_oracle_closest_pass(int mode, int64_t np, const float* nx, const float* ny, const float* nz, const uint32_t* idx,
                                const float* nrm, int64_t nt, float* x, float cutoff_mult, float ips);

void o3d_oracle_set_threads(int n);
int o3d_oracle_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif

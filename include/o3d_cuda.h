/* o3d_cuda.h - C ABI of the B200 (sm_100a) Biot-Savart back end for Omega3D.
 *
 * This is the drop-in boundary: the entry points below are what the reference's `gpu_cuda` arm
 * (accel_t::gpu_cuda, /root/reference/src/ExecEnv.h:34-39 - declared, never dispatched) binds. The
 * precedent for the shape of the interface is the reference's own external-solver hook
 * `external_vel_solver_f_` (src/Influence.h:34-47,97-103): SoA float arrays by pointer, counts by value,
 * results accumulated into caller-owned arrays. INTEGRATION.md shows the reference-side patch.
 *
 * Conventions shared by every entry point (they are the reference's):
 *   - all arrays are float32 SoA in the layout of the reference's containers (Points: x[3], s[3], r -
 *     src/Points.h; Surfaces: node x[3], idx[3*np], area, ts[3] - src/Surfaces.h);
 *   - results ACCUMULATE into the caller's arrays (`tu[d][i] += sum`, src/Influence.h:296-307,462-473;
 *     `-=` for particles->panels, :1210-1212), un-normalised (no 1/4pi, no freestream: the caller's
 *     finalize_vels does that, src/Points.h:265-277);
 *   - self interactions are NOT skipped (src/Kernels.h:184-192); sources may alias targets;
 *   - pairwise arithmetic in float, per-target sums carried to double before the final `+=`
 *     (the reference's non-Vc scheme, src/Simulation.h:41-47);
 *   - every function returns 0 on success or an O3D_ERR_* code; nothing aborts or throws across
 *     the boundary (the reference asserts; the integration arm asserts on non-zero);
 *   - a context is single-caller (the reference calls the influence routines from one thread at a time,
 *     src/Simulation.cpp:776,802); host pointers are never retained past the return;
 *   - `flops_out` (may be NULL) receives the reference's own flop estimate for the call so the caller
 *     can print its usual "[%.4f] seconds at %.3f GFlop/s" line (src/Influence.h:310,366,475,534).
 *
 * There is no CPU fallback: without a usable sm_100 device o3d_cuda_create fails with O3D_ERR_NODEVICE.
 */
#ifndef O3D_CUDA_H
#define O3D_CUDA_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define O3D_CUDA_ABI_VERSION 3

enum {
  O3D_OK = 0,
  O3D_ERR_INVALID = 1,     /* bad argument (NULL where an array is required, negative count, ...) */
  O3D_ERR_CUDA = 2,        /* a CUDA runtime call failed; see o3d_cuda_last_error */
  O3D_ERR_NOMEM = 3,       /* device or pinned-host allocation failed */
  O3D_ERR_NODEVICE = 4,    /* no CUDA device / wrong architecture */
  O3D_ERR_UNSUPPORTED = 5  /* combination the reference itself asserts on (src/Influence.h:368-370) */
};

typedef struct o3d_ctx o3d_ctx;

/* ---- context ------------------------------------------------------------------------------------ */
int o3d_cuda_abi_version(void);
int o3d_cuda_device_count(void);
/* One context drives `ndev` GPUs of this process (devices == NULL: 0..ndev-1). With ndev > 1 the host
 * entry points partition the TARGETS across the devices and replicate the sources to each
 * (SURVEY.md section 8e). One-process-per-GPU jobs create a 1-device context per rank and use the *_dev
 * entry points with an NCCL all-gather of the packed source records in between. */
int o3d_cuda_create(o3d_ctx** ctx, int ndev, const int* devices);
void o3d_cuda_destroy(o3d_ctx* ctx);
const char* o3d_cuda_last_error(const o3d_ctx* ctx);
int o3d_cuda_num_devices(const o3d_ctx* ctx);
/* SM count, SM clock (kHz) and FP32 FMA peak (flop/s at that clock: SMs x 128 x 2 x f) of device k. */
int o3d_cuda_device_props(const o3d_ctx* ctx, int k, int* sm_count, int* clock_khz, double* fp32_peak);
/* Device-side time of the LAST host entry point on this context, max over its devices, from CUDA events
 * on the launching streams: influence kernels only / host->device / device->host, in milliseconds;
 * `launches` = number of kernels this library launched for that call. Any pointer may be NULL. */
int o3d_cuda_last_timing(const o3d_ctx* ctx, double* kernel_ms, double* h2d_ms, double* d2h_ms, int* launches);

/* ---- host-pointer entry points: one per reference influence routine -------------------------------- */

/* particles -> points. Replaces the CPU block of points_affect_points<S,A> (src/Influence.h:202-551).
 *   sources: position sx,sy,sz, radius sr, strength ssx,ssy,ssz                       (ns each)
 *   targets: position tx,ty,tz; tr = radius, or NULL for singular (inert) targets     (nt each)
 *   tu,tv,tw: velocity, += ; tug: 9 arrays, slot 3*j+i = d u_i / d x_j, += ; NULL => velocity only
 * Kernel selected as the reference does (src/Influence.h:213-533):
 *   tr && tug -> kernel_0v_0bg   tr && !tug -> kernel_0v_0b   !tr && tug -> kernel_0v_0pg   else kernel_0v_0p */
int o3d_cuda_pts_on_pts(o3d_ctx* ctx, int64_t ns, const float* sx, const float* sy, const float* sz,
                        const float* sr, const float* ssx, const float* ssy, const float* ssz, int64_t nt,
                        const float* tx, const float* ty, const float* tz, const float* tr, float* tu, float* tv,
                        float* tw, float* const* tug, double* flops_out);

/* triangular panels -> points. Replaces panels_affect_points<S,A> (src/Influence.h:557-1099): per
 * (target, panel) the recursive rkernel_2vs_0p / rkernel_2vs_0pg (src/Kernels.h:1028-1211), maxlev 3.
 *   nodes: nx,ny,nz (nn each); idx: 3 node indices per panel; ts: total vortex strength tsx,tsy,tsz (np each);
 *   area (np); sss: source-sheet strength per panel or NULL (src/Influence.h:577-579)
 *   tug != NULL selects the gradient kernel (the reference keys this on the target having gradient
 *   storage, not on the results type: src/Influence.h:652,875). Target radius never enters (:873,893). */
int o3d_cuda_pan_on_pts(o3d_ctx* ctx, int64_t nn, const float* nx, const float* ny, const float* nz, int64_t np,
                        const uint32_t* idx, const float* tsx, const float* tsy, const float* tsz,
                        const float* area, const float* sss, int64_t nt, const float* tx, const float* ty,
                        const float* tz, float* tu, float* tv, float* tw, float* const* tug, double* flops_out);

/* particles -> panel centres (the BEM right-hand side). Replaces points_affect_panels<S,A>
 * (src/Influence.h:1107-1221): the same recursive kernel with the panel as geometry and the particle
 * strength / panel area as sheet strength; the result is SUBTRACTED from pu,pv,pw (np each). */
int o3d_cuda_pts_on_pan(o3d_ctx* ctx, int64_t ns, const float* sx, const float* sy, const float* sz,
                        const float* ssx, const float* ssy, const float* ssz, int64_t nn, const float* nx,
                        const float* ny, const float* nz, int64_t np, const uint32_t* idx, const float* area,
                        float* pu, float* pv, float* pw, double* flops_out);

/* panels -> panels BEM influence block. Replaces panels_on_panels_coeff<S> (src/Coefficients.h:169-483),
 * 3 unknowns per panel (vortex x1, vortex x2, source), rkernel_2vs_2p (src/Kernels.h:1217-1315).
 * Source panels: nodes + idx + basis sb1, sb2 (3 arrays of nsp each, SoA x|y|z) + area.
 * Target panels: nodes + idx + basis tb1, tb2, tnrm + area. self != 0: source and target are the same
 * surface, apply the diagonal override (:414-436). coeffs: column-major (3*ntp) x (3*nsp), OVERWRITTEN,
 * already scaled by 1/4pi (:448-451). */
int o3d_cuda_pan_on_pan_coeff(o3d_ctx* ctx, int64_t snn, const float* snx, const float* sny, const float* snz,
                              int64_t nsp, const uint32_t* sidx, const float* sb1, const float* sb2,
                              const float* sarea, int64_t tnn, const float* tnx, const float* tny,
                              const float* tnz, int64_t ntp, const uint32_t* tidx, const float* tb1,
                              const float* tb2, const float* tnrm, const float* tarea, int self, float* coeffs,
                              double* flops_out);

/* ---- device-pointer entry points (arrays already resident in HBM on the context's device 0) ------- */
/* `stream` is a cudaStream_t passed as an opaque pointer (NULL = the legacy default stream). All calls
 * are asynchronous with respect to the host. ONE stream per context at a time: the launches share the context's
 * scratch (radius-range block, stream-K workspace), so calls on different streams must be ordered by the caller, or use
 * one context per stream. The stream must belong to the context's first device, which must be the CURRENT device of the
 * calling thread (O3D_ERR_INVALID otherwise - the library does not switch devices under a caller that owns them). */

/* Number of 32-byte records the packed source stream holds for ns sources (padded to whole tiles). */
int64_t o3d_cuda_packed_records(int64_t ns);

/* The launch shape the library picks for particles -> points on a device with sm_count SMs - host arithmetic only, no
 * device needed (capi.cu: pp_shape). The kernels run as PERSISTENT CTAs (128 threads x 2 targets with gradients, x 4
 * without) over a static stream-K partition: the (target block, source tile) units, block-major, are dealt out in equal
 * contiguous shares to `grid` = min(units, 3 x sm_count) CTAs. `split_blocks` target blocks are shared by more than one
 * CTA; their FP64 partial sums meet in a fixed workspace of `workspace_bytes` and are added in unit order. `balance` =
 * mean / max tiles per CTA (1 = perfectly even). Any output pointer may be NULL. The reference's counterpart is the
 * static OpenMP schedule over targets (src/Influence.h:281,405,445). */
int o3d_cuda_plan_pts_on_pts(int sm_count, int64_t ns, int64_t nt, int want_grad, int64_t* grid, int* split_blocks,
                             double* balance, int64_t* workspace_bytes);
/* Host replay of that launch's bookkeeping (every CTA's segments, every workspace slot, every fix-up) - 0 when each unit
 * is consumed exactly once and each target block is finished exactly once; otherwise the number of the failed check. */
int o3d_cuda_plan_check(int sm_count, int64_t ns, int64_t nt, int want_grad);
/* SoA sources -> packed record stream `packed` (device memory, nrec * 32 bytes). nrec = 0 means
 * o3d_cuda_packed_records(ns); a larger whole number of tiles is filled up with zero-strength records
 * (ranks of a sharded job all contribute equally sized streams to one all-gather). */
int o3d_cuda_pack_sources_dev(o3d_ctx* ctx, void* stream, int64_t ns, const float* sx, const float* sy,
                              const float* sz, const float* sr, const float* ssx, const float* ssy,
                              const float* ssz, int64_t nrec, void* packed);
/* Packed sources (nrec records, any whole number of tiles - e.g. an all-gathered concatenation of every
 * rank's padded stream) -> targets. tug = base of a 9 x tug_stride float block or NULL. `workspace` is
 * used when the library splits the source range across CTAs for small target counts: pass NULL to let
 * the context own it. */
int o3d_cuda_pts_on_pts_dev(o3d_ctx* ctx, void* stream, int64_t nrec, const void* packed, int64_t nt,
                            const float* tx, const float* ty, const float* tz, const float* tr, float* tu,
                            float* tv, float* tw, float* tug, int64_t tug_stride);


/* ---- convection on the device: the O(N) steps around the influence sums --- */

/* One Runge-Kutta stage as the reference's Points::move sees it: velocity u[3] and, optionally, the 9 x ug_stride
 * velocity-gradient block of a Points object evaluated at that stage (device pointers). */
typedef struct {
  const float* u[3];
  const float* ug;      /* NULL: this stage carries no gradients => no stretching (src/Points.h:296,366,453) */
  int64_t ug_stride;
} o3d_stage;

/* finalize_vels on device arrays: u = fs + u/(4 pi) in double, grads *= float(1/(4 pi))
 * (src/ElementBase.h:187-192, src/Points.h:265-277). ug may be NULL. */
int o3d_cuda_pts_finalize_dev(o3d_ctx* ctx, void* stream, int64_t n, float* u, float* v, float* w, float* ug,
                              int64_t ug_stride, const double* fs);
/* Points::move with `order` = 1, 2 or 3 stages (src/ElementBase.h:253-336 advection, src/Points.h:288-520
 * stretching and elongation), rounding as the reference's scalar build does. State is read from xin/sin/ein and
 * written to xout/sout/eout (may alias; arrays of 3 device pointers; sin/sout NULL: inert points; ein/eout NULL:
 * elongation not tracked). order 1: position moves with stages[0].u, strengths stretch with stages[0].ug (the
 * caller passes the object's OWN gradients there, as src/Points.h:296-332 uses them). order >= 2 also stores the
 * combined velocity in uout (3 pointers), as the reference does into this->u. wt: `order` stage weights. */
int o3d_cuda_pts_move_dev(o3d_ctx* ctx, void* stream, int64_t n, int order, double dt, const double* wt,
                          const o3d_stage* stages, const float* const* xin, const float* const* sin,
                          const float* ein, float* const* xout, float* const* sout, float* eout, float* const* uout);

/* A vortex-particle collection resident in HBM (the reference's Points<S>, active + lagrangian: position,
 * strength, radius, elongation, velocity, velocity gradient). With a multi-device context the particles are
 * block-partitioned over the devices in whole 512-particle tiles; each device evaluates and moves its own block
 * and the packed source records are exchanged device-to-device (NVLink peer copies) once per evaluation. */
typedef struct o3d_particles o3d_particles;
int o3d_cuda_particles_create(o3d_ctx* ctx, o3d_particles** out);
void o3d_cuda_particles_destroy(o3d_ctx* ctx, o3d_particles* p);
int64_t o3d_cuda_particles_count(const o3d_particles* p);
/* host -> device. elong == NULL: 1 everywhere (a fresh collection, src/Points.h:120-127). */
int o3d_cuda_particles_upload(o3d_ctx* ctx, o3d_particles* p, int64_t n, const float* x, const float* y,
                              const float* z, const float* sx, const float* sy, const float* sz, const float* r,
                              const float* elong);
/* device -> host; any pointer may be NULL (skipped). ug: 9 host arrays or NULL. */
int o3d_cuda_particles_download(o3d_ctx* ctx, o3d_particles* p, float* x, float* y, float* z, float* sx, float* sy,
                                float* sz, float* r, float* elong, float* u, float* v, float* w, float* const* ug);
/* Convection::find_vels(fs, vort, {}, vort) for this collection (src/Convection.h:130-184): zero_vels,
 * particles -> themselves (velocity [+ gradient]), finalize_vels(fs). */
int o3d_cuda_particles_find_vels(o3d_ctx* ctx, o3d_particles* p, const double* fs, int want_grad, double* flops_out);
/* `nsteps` calls of Convection::advect(time, dt, fs, ...) for a particle-only system (no boundaries, no field
 * points): order 1 = advect_1st (src/Convection.h:232-262), 2 = advect_2nd_ralston (:349-425), 3 = advect_3rd
 * (:431-556). Nothing leaves the device between steps; steps of unchanged size replay one captured CUDA graph.
 * Afterwards velocity holds what the reference leaves in it (order 1: u at the start of the last step; order >= 2:
 * the last combined stage velocity) and the gradient the first-stage gradient of the last step. */
int o3d_cuda_particles_advect(o3d_ctx* ctx, o3d_particles* p, int order, double time, double dt, const double* fs,
                              int nsteps, double* flops_out);
/* panels -> points (o3d_cuda_pan_on_pts, bodies attached to resident collections) pools the (point, panel) pairs that need
 * subdivision per warp and lets all 32 lanes drain the pool (default); off = every lane walks only its own pairs, the
 * round-1 kernel, kept for A/B measurements. Same leaves, same counts; FP32 terms of a tile regrouped.
 * on = 1 (default): panels -> points only; 2: particles -> panels as well (o3d_cuda_pts_on_pan, the BEM right-hand side of a
 * resident body: measured no faster than the per-lane kernel, so not the default); 0: neither. */
int o3d_cuda_set_panel_queue(o3d_ctx* ctx, int on);
/* Host arrays of the entry points above may be pageable (the reference's std::vector storage) or pinned. Pageable arrays
 * travel through a ring of pinned slots the context owns (4 x 8 MB per device; helper threads fill the next slot while the
 * DMA engine moves the previous one; O3D_CUDA_COPY_THREADS sets their number); pinned ones go to the DMA engine directly.
 * off = hand every pointer to cudaMemcpyAsync as it comes (the driver's own bounce buffers) - for A/B measurements.
 * With ndev > 1 the sources go to device 0 once and reach the other devices as packed records over NVLink; the environment
 * variable O3D_CUDA_SOURCES_PCIE_ALL=1 (read by o3d_cuda_create) makes every device upload them itself instead (A/B only). */
int o3d_cuda_set_host_staging(o3d_ctx* ctx, int on);
/* CUDA-graph replay of repeated steps is on by default; off = launch every kernel individually (same results,
 * bit for bit - tests compare the two). o3d_cuda_particles_graph_active: 1 if the collection holds a captured step. */
int o3d_cuda_set_graphs(o3d_ctx* ctx, int on);
int o3d_cuda_particles_graph_active(const o3d_particles* p);
/* sqrt(max |s|^2) (ElementBase::get_max_str, src/ElementBase.h:339-351) and max elongation (src/Points.h:523-532). */
int o3d_cuda_particles_stats(o3d_ctx* ctx, o3d_particles* p, float* max_str, float* max_elong);


/* ---- a static body attached to a resident collection (Convection::find_vels / advect with boundaries) ---------------
 * With a body attached, o3d_cuda_particles_find_vels adds panels -> particles to the particle sums (src/Convection.h:157-167)
 * and o3d_cuda_particles_advect runs the reference's sequence for a system with a boundary: before every derivative
 * evaluation the device forms what solve_bem (src/BEMHelper.h:44-262) starts with - panel-centre velocities zeroed, then
 * points_affect_panels from the state's particles (:83-94; un-normalised, subtracted as the reference subtracts them) - and
 * hands them to `solve`, the REST of solve_bem: finalize_vels(fs), right-hand side (src/RHS.h), A, the solve (src/BEM.h,
 * Eigen GMRES), set_str. That is the reference's host code and stays there; it returns the panels' total vortex strengths
 * (and source strengths). After every move clear_inner_layer(1, body, particles, cutoff_mult, ips) (src/Reflect.h:625-655)
 * runs on the moved state. Particle arrays never leave the device; per evaluation 3 np floats go to the host and 4 np
 * come back.
 *
 * solve(user, np, pu, tsx, tsy, tsz, sss, have_source): pu = 3 x np floats (u | v | w rows), the raw sums, IN; tsx.. = np
 * floats each OUT (Surfaces::get_str after set_str, src/Surfaces.h:267-335); sss = np source strengths OUT, used iff
 * *have_source is set non-zero. Returns 0, anything else aborts the call with O3D_ERR_CUDA. solve may be NULL: the
 * strengths then stay what o3d_cuda_particles_set_body_strengths last set (zero after set_body).
 * nodes SoA (nn), idx 3 per panel, area np, nrm 3 x np (x | y | z rows) as the reference's Surfaces holds them. */
typedef int (*o3d_bem_solve_fn)(void* user, int64_t np, const float* pu, float* tsx, float* tsy, float* tsz, float* sss, int* have_source);
int o3d_cuda_particles_set_body(o3d_ctx* ctx, o3d_particles* p, int64_t nn, const float* nx, const float* ny, const float* nz,
                                int64_t np, const uint32_t* idx, const float* area, const float* nrm, float cutoff_mult, float ips,
                                o3d_bem_solve_fn solve, void* user);
int o3d_cuda_particles_clear_body(o3d_ctx* ctx, o3d_particles* p);
int o3d_cuda_particles_set_body_strengths(o3d_ctx* ctx, o3d_particles* p, const float* tsx, const float* tsy, const float* tsz,
                                          const float* sss /* NULL: no source sheet */);
/* Those raw panel-centre sums of the current state alone (no solve, strengths untouched): np floats each. */
int o3d_cuda_particles_body_vels(o3d_ctx* ctx, o3d_particles* p, float* pu, float* pv, float* pw);
/* clear_inner_layer on the resident positions, outside a step (src/Simulation.cpp:839 after diffusion). */
int o3d_cuda_particles_clear_inner(o3d_ctx* ctx, o3d_particles* p, int64_t* num_moved);
/* Of the last o3d_cuda_particles_advect: particles pushed out by its clear-inner passes, BEM solves requested. */
int o3d_cuda_particles_body_counters(const o3d_particles* p, int64_t* moved, int* solves);


/* ---- status-file quantities and writer (src/StatusFile.cpp, src/Simulation.cpp:851-924) ------------ */
/* Total circulation sum_i s_i (ElementBase::get_total_circ, src/ElementBase.h:354-378) and linear impulse
 * sum_i (s1 x2 - s2 x1, s2 x0 - s0 x2, s0 x1 - s1 x0) (Points::get_total_impulse, src/Points.h:547-563) of a resident
 * collection, reduced on the device(s): per-particle terms in float as the reference forms them, sums in FP64 over a
 * fixed-shape tree (deterministic). circ, impulse: 3 doubles each, either may be NULL. */
int o3d_cuda_particles_totals(o3d_ctx* ctx, o3d_particles* p, double* circ, double* impulse);

/* The line-per-step status file: StatusFile's behaviour byte for byte (header of value names per data set, "# " and
 * spaces for .dat, commas for csv != 0, values printed as by operator<<, an empty line between data sets, append mode).
 * Host I/O only - these five need no device. append_*: name NULL = StatusFile's anonymous "float" / "int". */
typedef struct o3d_status o3d_status;
int o3d_cuda_status_open(const char* path, int csv, o3d_status** out);
void o3d_cuda_status_close(o3d_status* st);
int o3d_cuda_status_reset_sim(o3d_status* st);                                   /* StatusFile::reset_sim */
int o3d_cuda_status_append_float(o3d_status* st, const char* name, float value);  /* StatusFile::append_value */
int o3d_cuda_status_append_int(o3d_status* st, const char* name, int value);
int o3d_cuda_status_write_line(o3d_status* st);                                   /* StatusFile::write_line */
/* Simulation::dump_stats_to_status for a system that is one resident particle collection: appends time, Nv, gx gy gz
 * (o3d_cuda_particles_totals) and fx fy fz (calculate_simple_forces: the one-sided time difference of the total impulse,
 * whose previous sample the status object keeps; time < 0.1 dt restarts it) and writes the line. */
int o3d_cuda_particles_write_status(o3d_ctx* ctx, o3d_particles* p, o3d_status* st, double time, double dt);


/* ---- matrix-free BEM operator ----------------------------------------------------------- */
/* y = A x where A is exactly the (3 ntp) x (3 nsp) block o3d_cuda_pan_on_pan_coeff builds (same arguments), applied
 * without being stored: what the GMRES of BEM<S,I>::solve needs from A (src/BEM.h:182-202, `A * x`), for panel counts
 * whose dense matrix no longer fits (the reference caps them for that reason, src/Simulation.cpp:675). create uploads
 * and packs the geometry once; apply takes host vectors x (3 nsp) and y (3 ntp, overwritten). With a multi-device
 * context the rows (target panels) are partitioned over the devices. */
typedef struct o3d_bem_op o3d_bem_op;
int o3d_cuda_bem_op_create(o3d_ctx* ctx, int64_t snn, const float* snx, const float* sny, const float* snz, int64_t nsp,
                           const uint32_t* sidx, const float* sb1, const float* sb2, const float* sarea, int64_t tnn,
                           const float* tnx, const float* tny, const float* tnz, int64_t ntp, const uint32_t* tidx,
                           const float* tb1, const float* tb2, const float* tnrm, const float* tarea, int self,
                           o3d_bem_op** out);
int o3d_cuda_bem_op_apply(o3d_ctx* ctx, o3d_bem_op* op, const float* x, float* y, double* flops_out);
void o3d_cuda_bem_op_destroy(o3d_ctx* ctx, o3d_bem_op* op);

/* ---- particle x panel closest-point loops ----------------------------------------------- */
/* Panels: nodes nx,ny,nz (nn), 3 node indices per panel, unit normals nrm as x|y|z rows of np (Surfaces::get_norm).
 * Particle positions tx,ty,tz (nt) are updated in place; *num = how many were moved. Results are bit-identical to the
 * reference's scalar loops (same panel order, same unfused float arithmetic).
 * o3d_cuda_reflect_pts: reflect_panp2<S> (src/Reflect.h:194-311) - particles under the surface are mirrored out.
 * o3d_cuda_clear_inner_pts: clear_inner_panp2<S> (src/Reflect.h:446-620) with _method = 1, the only one the
 * reference calls - particles lower than cutoff_mult * ips above the surface are pushed out to that height.
 * Any other method: O3D_ERR_UNSUPPORTED. */
int o3d_cuda_reflect_pts(o3d_ctx* ctx, int64_t nn, const float* nx, const float* ny, const float* nz, int64_t np,
                         const uint32_t* idx, const float* nrm, int64_t nt, float* tx, float* ty, float* tz,
                         int64_t* num_reflected);
int o3d_cuda_clear_inner_pts(o3d_ctx* ctx, int method, int64_t nn, const float* nx, const float* ny, const float* nz,
                             int64_t np, const uint32_t* idx, const float* nrm, int64_t nt, float* tx, float* ty,
                             float* tz, float cutoff_mult, float ips, int64_t* num_moved);

/* ---- field output -------------------------------------------------------------------------- */
/* The `.vtu` file of Points<S>::write_vtk (src/Points.h:851-1039, src/VtkXmlWriter.h:36-148), byte for byte: FieldData
 * TimeValue, positions, vertex cells, circulation (strengths), radius, velocity, base64 DataArrays. Host arrays, SoA.
 * sx/sy/sz and r NULL: inert points (the reference's "fldpt_" files). o3d_cuda_particles_write_vtu downloads a resident
 * collection and writes it. File naming ("part_%02d_%05d.vtu") is the caller's. */
int o3d_cuda_write_points_vtu(const char* path, int64_t n, const float* x, const float* y, const float* z,
                              const float* sx, const float* sy, const float* sz, const float* r, const float* u,
                              const float* v, const float* w, double time);
int o3d_cuda_particles_write_vtu(o3d_ctx* ctx, o3d_particles* p, const char* path, double time);

/* ---- kernel selection ---------------------------------------------------------------------------------- */
/* The particles-on-points kernel exists twice in the library, from the same source: as compiled, and post-processed at
 * the SASS level (tools/sass_patch.py: longer operand-reuse chains; same instructions, same results bit for bit, ~2 %
 * faster). The post-processed copy is the default (environment O3D_CUDA_TUNED=0 starts contexts on the other);
 * this switch exists so that tests and profiles can compare the two. o3d_cuda_tuned_kernels: 1 if on. */
int o3d_cuda_set_tuned_kernels(o3d_ctx* ctx, int on);
int o3d_cuda_tuned_kernels(const o3d_ctx* ctx);

/* Core function of the particle kernels. The reference selects it per BUILD by moving the one active
 * "#define USE_*_KERNEL" in src/CoreFunc.h:35-38; here it is a property of the context, Winckelmans-Leonard (the
 * define the reference ships with) by default. integration/O3DCudaInfluence.h passes the value that matches the
 * define its translation unit was compiled with, so the CUDA arm always follows the reference build it sits in.
 * Applies to particles -> points (all four variants, host and *_dev entry points, resident particle collections)
 * and to every flops_out figure (flops_t{v,p}_{grads,nograds}, src/CoreFunc.h); panel leaves evaluate the cores
 * at zero radius, where all four coincide. A packed source stream (o3d_cuda_pack_sources_dev) carries the radius
 * term of the core that was set when it was packed: set the core first, then pack, then evaluate.
 * o3d_cuda_core_func returns the current value (-1 on a NULL context). */
enum o3d_core_func {
  O3D_CORE_WL = 0,   /* Winckelmans-Leonard      src/CoreFunc.h:241-289 */
  O3D_CORE_RM = 1,   /* Rosenhead-Moore          src/CoreFunc.h:43-83   */
  O3D_CORE_EXP = 2,  /* exponential              src/CoreFunc.h:86-238  */
  O3D_CORE_V2 = 3    /* Vatistas n = 2           src/CoreFunc.h:292-341 */
};
int o3d_cuda_set_core_func(o3d_ctx* ctx, int core);
int o3d_cuda_core_func(const o3d_ctx* ctx);

/* ---- measurement helpers ----------------------------------------------------------------------------- */
/* When on, o3d_cuda_pts_on_pts_dev brackets its dominant kernel with CUDA events on the launching stream;
 * o3d_cuda_dev_kernel_ms waits for the last such launch and returns its device time in milliseconds. */
int o3d_cuda_set_profiling(o3d_ctx* ctx, int on);
int o3d_cuda_dev_kernel_ms(o3d_ctx* ctx, double* ms);
/* Runs a dependency-free packed-FMA (fma.rn.f32x2) loop on every SM of device 0 for a few tens of
 * milliseconds and reports the FP32 rate the GPU sustains at its real clocks, in TFLOP/s. */
int o3d_cuda_probe_fp32_peak(o3d_ctx* ctx, double* tflops, double* ms);

#ifdef __cplusplus
}
#endif
#endif /* O3D_CUDA_H */

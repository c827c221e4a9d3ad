/* oracle/biot_oracle.c - plain-C restatement of Omega3D's direct Biot-Savart path (see biot_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY - the checker, never the product. Compile with -ffp-contract=off so
 * every float operation rounds exactly as the reference's scalar build does with the same flag.
 * Each function cites the reference file:line it follows (paths relative to /root/reference).
 */
#include "biot_oracle.h"
#include <math.h>
#include <omp.h>
#include <stddef.h>

#define O3D_MAXLEV 3 /* RECURSIVE_LEVELS, src/Influence.h:23 and src/Coefficients.h:23 */

/* ---- src/MathHelper.h:122-135 (scalar specialisations): x^-2.5 by one sqrt and one divide ---- */
static inline float inv_pow_2p5(float x) { return 1.0f / (x * x * sqrtf(x)); }

/* ---- Winckelmans-Leonard core, src/CoreFunc.h:245-288 ---- */
/* velocity only; blob target (:245-250) and singular target (:255-259) */
static inline float wl_core_blob(float distsq, float sr, float tr) {
  const float r2 = sr * sr + tr * tr;
  const float d2 = distsq + r2;
  return (distsq + 2.5f * r2) * inv_pow_2p5(d2);
}
static inline float wl_core_point(float distsq, float sr) {
  const float r2 = sr * sr;
  const float d2 = distsq + r2;
  return (distsq + 2.5f * r2) * inv_pow_2p5(d2);
}
/* with the gradient factor; blob (:265-274) and singular (:279-287) */
static inline void wl_core_grad_r2(float distsq, float r2, float* r3, float* bbb) {
  const float d2 = distsq + r2;
  const float d2top = distsq + 2.5f * r2;
  const float dn5 = inv_pow_2p5(d2);
  *r3 = d2top * dn5;
  *bbb = 2.0f * dn5 - 5.0f * d2top * dn5 / d2;
}
static inline void wl_core_blob_grad(float distsq, float sr, float tr, float* r3, float* bbb) {
  wl_core_grad_r2(distsq, sr * sr + tr * tr, r3, bbb);
}
static inline void wl_core_point_grad(float distsq, float sr, float* r3, float* bbb) {
  wl_core_grad_r2(distsq, sr * sr, r3, bbb);
}

/* ---- pairwise kernels, src/Kernels.h ---- */
typedef struct { float x, y, z, r, wx, wy, wz, q; } src_t; /* q = scalar source strength (vs kernels) */

/* velocity: common tail of kernel_0v_0b (:50-69), kernel_0v_0p (:94-112), kernel_0vs_0p (:115-133) */
#define VEL_BODY(ACC_T, CORE, WITH_Q)                                           \
  const float dx = tx - s->x, dy = ty - s->y, dz = tz - s->z;                  \
  const float k = CORE;                                                        \
  float cx = dz * s->wy - dy * s->wz;                                          \
  float cy = dx * s->wz - dz * s->wx;                                          \
  float cz = dy * s->wx - dx * s->wy;                                          \
  if (WITH_Q) { cx = cx + dx * s->q; cy = cy + dy * s->q; cz = cz + dz * s->q; } \
  acc[0] += (ACC_T)(k * cx);                                                   \
  acc[1] += (ACC_T)(k * cy);                                                   \
  acc[2] += (ACC_T)(k * cz);

static inline void k_0v_0b(const src_t* s, float tx, float ty, float tz, float tr, double* acc) {
  VEL_BODY(double, wl_core_blob(dx * dx + dy * dy + dz * dz, s->r, tr), 0)
}
static inline void k_0v_0p(const src_t* s, float tx, float ty, float tz, double* acc) {
  VEL_BODY(double, wl_core_point(dx * dx + dy * dy + dz * dz, s->r), 0)
}
static inline void k_0vs_0p(const src_t* s, float tx, float ty, float tz, double* acc) {
  VEL_BODY(double, wl_core_point(dx * dx + dy * dy + dz * dz, s->r), 1)
}
static inline void k_0vs_0p_f(const src_t* s, float tx, float ty, float tz, float* acc) {
  VEL_BODY(float, wl_core_point(dx * dx + dy * dy + dz * dz, s->r), 1)
}

/* velocity + 9 gradients: kernel_0v_0bg (:155-193), kernel_0v_0pg (:253-290), kernel_0vs_0pg (:294-345).
 * acc layout: u v w | ux vx wx | uy vy wy | uz vz wz  (gradient slot 3*j+i = d u_i / d x_j). */
static inline void grad_body(const src_t* s, float dx, float dy, float dz, float r3, float bbb,
                             int with_q, double* acc) {
  float cx = dz * s->wy - dy * s->wz;
  float cy = dx * s->wz - dz * s->wx;
  float cz = dy * s->wx - dx * s->wy;
  if (with_q) {
    acc[0] += (double)(r3 * (cx + dx * s->q));
    acc[1] += (double)(r3 * (cy + dy * s->q));
    acc[2] += (double)(r3 * (cz + dz * s->q));
  } else {
    acc[0] += (double)(r3 * cx);
    acc[1] += (double)(r3 * cy);
    acc[2] += (double)(r3 * cz);
  }
  cx *= bbb; cy *= bbb; cz *= bbb;
  acc[3]  += (double)(dx * cx);
  acc[4]  += (double)(dx * cy + s->wz * r3);
  acc[5]  += (double)(dx * cz - s->wy * r3);
  acc[6]  += (double)(dy * cx - s->wz * r3);
  acc[7]  += (double)(dy * cy);
  acc[8]  += (double)(dy * cz + s->wx * r3);
  acc[9]  += (double)(dz * cx + s->wy * r3);
  acc[10] += (double)(dz * cy - s->wx * r3);
  acc[11] += (double)(dz * cz);
  if (with_q) { /* gradient of the source-strength part, :327-344 */
    const float gx = dx * bbb * s->q, gy = dy * bbb * s->q, gz = dz * bbb * s->q;
    const float iso = s->q * r3;
    acc[3]  += (double)(dx * gx + iso);
    acc[4]  += (double)(dx * gy);
    acc[5]  += (double)(dx * gz);
    acc[6]  += (double)(dy * gx);
    acc[7]  += (double)(dy * gy + iso);
    acc[8]  += (double)(dy * gz);
    acc[9]  += (double)(dz * gx);
    acc[10] += (double)(dz * gy);
    acc[11] += (double)(dz * gz + iso);
  }
}
static inline void k_0v_0bg(const src_t* s, float tx, float ty, float tz, float tr, double* acc) {
  const float dx = tx - s->x, dy = ty - s->y, dz = tz - s->z;
  float r3, bbb;
  wl_core_blob_grad(dx * dx + dy * dy + dz * dz, s->r, tr, &r3, &bbb);
  grad_body(s, dx, dy, dz, r3, bbb, 0, acc);
}
static inline void k_0v_0pg(const src_t* s, float tx, float ty, float tz, double* acc) {
  const float dx = tx - s->x, dy = ty - s->y, dz = tz - s->z;
  float r3, bbb;
  wl_core_point_grad(dx * dx + dy * dy + dz * dz, s->r, &r3, &bbb);
  grad_body(s, dx, dy, dz, r3, bbb, 0, acc);
}
static inline void k_0vs_0pg(const src_t* s, float tx, float ty, float tz, double* acc) {
  const float dx = tx - s->x, dy = ty - s->y, dz = tz - s->z;
  float r3, bbb;
  wl_core_point_grad(dx * dx + dy * dy + dz * dz, s->r, &r3, &bbb);
  grad_body(s, dx, dy, dz, r3, bbb, 1, acc);
}

/* ---- recursive panel kernels, src/Kernels.h:1028-1315 ---- */
typedef struct { float x[3], y[3], z[3]; } tri_t;

/* the 4 children of a flat triangle from its 3 corners + 3 edge midpoints (:1081-1089) */
static void tri_split(const tri_t* p, tri_t c[4]) {
  const float nx[6] = {p->x[0], 0.5f * (p->x[0] + p->x[1]), p->x[1], 0.5f * (p->x[0] + p->x[2]), 0.5f * (p->x[1] + p->x[2]), p->x[2]};
  const float ny[6] = {p->y[0], 0.5f * (p->y[0] + p->y[1]), p->y[1], 0.5f * (p->y[0] + p->y[2]), 0.5f * (p->y[1] + p->y[2]), p->y[2]};
  const float nz[6] = {p->z[0], 0.5f * (p->z[0] + p->z[1]), p->z[1], 0.5f * (p->z[0] + p->z[2]), 0.5f * (p->z[1] + p->z[2]), p->z[2]};
  static const int child[4][3] = {{0, 1, 3}, {1, 2, 4}, {1, 4, 3}, {3, 4, 5}};
  for (int k = 0; k < 4; ++k)
    for (int v = 0; v < 3; ++v) {
      c[k].x[v] = nx[child[k][v]];
      c[k].y[v] = ny[child[k][v]];
      c[k].z[v] = nz[child[k][v]];
    }
}
static inline void tri_centroid(const tri_t* p, float* cx, float* cy, float* cz) {
  *cx = (p->x[0] + p->x[1] + p->x[2]) / 3.0f;
  *cy = (p->y[0] + p->y[1] + p->y[2]) / 3.0f;
  *cz = (p->z[0] + p->z[1] + p->z[2]) / 3.0f;
}
static inline float dist3(float dx, float dy, float dz) { return sqrtf(dx * dx + dy * dy + dz * dz); }

/* panel -> point. grads != 0 selects rkernel_2vs_0pg (:1120-1211), else rkernel_2vs_0p (:1028-1115).
 * Flop bookkeeping mirrors the reference so the returned count can be compared too. */
static int rk_2vs_0(const tri_t* p, float gx, float gy, float gz, float gs, float tx, float ty, float tz,
                    float sa, int lev, int grads, double* acc) {
  int flops = 0;
  if (lev == 0) { gx *= sa; gy *= sa; gz *= sa; gs *= sa; flops += 4; }
  float cx, cy, cz;
  tri_centroid(p, &cx, &cy, &cz);
  flops += 9;
  const float trisize = sqrtf(sa);
  const float dist = dist3(tx - cx, ty - cy, tz - cz);
  flops += 10;
  const int wellsep = dist > trisize * 4.0f; /* my_well_sep, :1000-1002 */
  flops += 1;
  if (wellsep || lev == O3D_MAXLEV) {
    const src_t leaf = {cx, cy, cz, 0.0f, gx, gy, gz, gs};
    if (grads) { k_0vs_0pg(&leaf, tx, ty, tz, acc); flops += 79 + 14; }
    else       { k_0vs_0p(&leaf, tx, ty, tz, acc);  flops += 29 + 8; }
  } else {
    gx *= 0.25f; gy *= 0.25f; gz *= 0.25f; gs *= 0.25f;
    const float ca = 0.25f * sa;
    flops += 5 + 18;
    tri_t c[4];
    tri_split(p, c);
    for (int k = 0; k < 4; ++k) flops += rk_2vs_0(&c[k], gx, gy, gz, gs, tx, ty, tz, ca, lev + 1, grads, acc);
  }
  return flops;
}

/* panel -> panel, float accumulators (:1217-1315) */
static void rk_2vs_2(const tri_t* p, float gx, float gy, float gz, float gs, const tri_t* q, float sa, float ta,
                     int lev, float* acc) {
  if (lev == 0) { gx *= sa; gy *= sa; gz *= sa; gs *= sa; }
  float sx, sy, sz, tx, ty, tz;
  tri_centroid(p, &sx, &sy, &sz);
  tri_centroid(q, &tx, &ty, &tz);
  const float trisize = sqrtf(sa) + sqrtf(ta);
  const float dist = dist3(tx - sx, ty - sy, tz - sz);
  if (dist > trisize * 4.0f || lev == O3D_MAXLEV) {
    const src_t leaf = {sx, sy, sz, 0.0f, gx, gy, gz, gs};
    k_0vs_0p_f(&leaf, tx, ty, tz, acc);
  } else {
    gx *= 0.0625f; gy *= 0.0625f; gz *= 0.0625f; gs *= 0.0625f;
    const float sca = 0.25f * sa, tca = 0.25f * ta;
    tri_t sc[4], tc[4];
    tri_split(p, sc);
    tri_split(q, tc);
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) rk_2vs_2(&sc[i], gx, gy, gz, gs, &tc[j], sca, tca, lev + 1, acc);
  }
}

static inline void load_tri(tri_t* t, const float* nx, const float* ny, const float* nz, const uint32_t* idx, int64_t p) {
  for (int v = 0; v < 3; ++v) {
    const uint32_t n = idx[3 * p + v];
    t->x[v] = nx[n]; t->y[v] = ny[n]; t->z[v] = nz[n];
  }
}

/* ---- exported single-interaction hooks ---- */
static src_t src_from7(const float s[7]) { src_t r = {s[0], s[1], s[2], s[3], s[4], s[5], s[6], 0.0f}; return r; }
void o3d_oracle_kernel_0v_0b(const float s[7], const float t[4], double u[3]) {
  const src_t a = src_from7(s); u[0] = u[1] = u[2] = 0.0; k_0v_0b(&a, t[0], t[1], t[2], t[3], u);
}
void o3d_oracle_kernel_0v_0p(const float s[7], const float t[3], double u[3]) {
  const src_t a = src_from7(s); u[0] = u[1] = u[2] = 0.0; k_0v_0p(&a, t[0], t[1], t[2], u);
}
void o3d_oracle_kernel_0v_0bg(const float s[7], const float t[4], double out[12]) {
  const src_t a = src_from7(s); for (int i = 0; i < 12; ++i) out[i] = 0.0; k_0v_0bg(&a, t[0], t[1], t[2], t[3], out);
}
void o3d_oracle_kernel_0v_0pg(const float s[7], const float t[3], double out[12]) {
  const src_t a = src_from7(s); for (int i = 0; i < 12; ++i) out[i] = 0.0; k_0v_0pg(&a, t[0], t[1], t[2], out);
}
static tri_t tri_from9(const float v[9]) {
  tri_t t; for (int k = 0; k < 3; ++k) { t.x[k] = v[3 * k]; t.y[k] = v[3 * k + 1]; t.z[k] = v[3 * k + 2]; } return t;
}
int o3d_oracle_rkernel_2vs_0p(const float tri[9], const float str[4], const float t[3], float sa, double u[3]) {
  const tri_t p = tri_from9(tri); u[0] = u[1] = u[2] = 0.0;
  return rk_2vs_0(&p, str[0], str[1], str[2], str[3], t[0], t[1], t[2], sa, 0, 0, u);
}
int o3d_oracle_rkernel_2vs_0pg(const float tri[9], const float str[4], const float t[3], float sa, double out[12]) {
  const tri_t p = tri_from9(tri); for (int i = 0; i < 12; ++i) out[i] = 0.0;
  return rk_2vs_0(&p, str[0], str[1], str[2], str[3], t[0], t[1], t[2], sa, 0, 1, out);
}

/* ---- loop nests ---- */
static inline void flush_acc(int nacc, const double* acc, int64_t nt, int64_t i, float* tu, float* tug, double sign) {
  for (int d = 0; d < 3; ++d) tu[d * nt + i] = (float)((double)tu[d * nt + i] + sign * acc[d]);
  if (nacc == 12) for (int d = 0; d < 9; ++d) tug[d * nt + i] = (float)((double)tug[d * nt + i] + acc[3 + d]);
}

/* src/Influence.h:278-309 (0pg), :351-365 (0p), :443-474 (0bg), :518-533 (0b): OpenMP over targets,
 * sources innermost in index order, one double accumulator set per target, then tu[d][i] += acc. */
void o3d_oracle_pts_on_pts(int64_t ns, const float* sx, const float* sy, const float* sz, const float* sr,
                           const float* ssx, const float* ssy, const float* ssz,
                           int64_t nt, const float* tx, const float* ty, const float* tz, const float* tr,
                           float* tu, float* tug) {
  const int nacc = tug ? 12 : 3;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nt; ++i) {
    double acc[12] = {0};
    for (int64_t j = 0; j < ns; ++j) {
      const src_t s = {sx[j], sy[j], sz[j], sr[j], ssx[j], ssy[j], ssz[j], 0.0f};
      if (tug) { if (tr) k_0v_0bg(&s, tx[i], ty[i], tz[i], tr[i], acc); else k_0v_0pg(&s, tx[i], ty[i], tz[i], acc); }
      else     { if (tr) k_0v_0b(&s, tx[i], ty[i], tz[i], tr[i], acc);  else k_0v_0p(&s, tx[i], ty[i], tz[i], acc); }
    }
    flush_acc(nacc, acc, nt, i, tu, tug, 1.0);
  }
}

/* ---- the alternate core functions of src/CoreFunc.h (one is chosen per BUILD of the reference by moving the active
 * "#define USE_*_KERNEL", :35-38). core: 0 Winckelmans-Leonard (:241-289, the shipped choice), 1 Rosenhead-Moore
 * (:43-83), 2 exponential (:86-238), 3 Vatistas n=2 (:292-341). Scalar helpers: src/MathHelper.h:55-181. ---- */
static inline float inv_pow_1p5(float x) { return 1.0f / (x * sqrtf(x)); }                               /* oor1p5 */
static inline float inv_pow_0p75(float x) { const float sqd = sqrtf(x); return 1.0f / (sqd * sqrtf(sqd)); } /* oor0p75 */
/* exp_cond (:114-128) and exp_bbb (:158-172), generic non-Vc templates */
static inline float exp_cond(float ood3, float corefac, float reld3) {
  if (reld3 > 16.0f) return ood3;
  else if (reld3 < 0.001f) return corefac;
  else return ood3 * (1.0f - expf(-reld3));
}
static inline float exp_bbb(float r3, float corefac, float reld3, float dist, float distsq) {
  if (reld3 > 16.0f) return -3.0f * r3 / distsq;
  else if (reld3 < 0.001f) return -1.5f * dist * r3 * r3;
  else { const float e = expf(-reld3); return 3.0f * (corefac * e - r3) / distsq; }
}
/* velocity only; blob != 0: core_func(distsq, sr, tr), else core_func(distsq, sr) */
static inline float alt_core(int core, int blob, float distsq, float sr, float tr) {
  if (core == 1) {
    const float r2 = blob ? distsq + sr * sr + tr * tr : distsq + sr * sr;
    return inv_pow_1p5(r2);
  } else if (core == 2) {
    const float dist = sqrtf(distsq);
    const float ood3 = 1.0f / (distsq * dist);
    const float corefac = blob ? 1.0f / (sr * sr * sr + tr * tr * tr) : 1.0f / (sr * sr * sr);
    const float reld3 = corefac / ood3;
    return exp_cond(ood3, corefac, reld3);
  } else if (core == 3) {
    const float s2 = sr * sr, t2 = tr * tr;
    const float denom = blob ? distsq * distsq + s2 * s2 + t2 * t2 : distsq * distsq + s2 * s2;
    return inv_pow_0p75(denom);
  }
  return blob ? wl_core_blob(distsq, sr, tr) : wl_core_point(distsq, sr);
}
/* with the gradient factor */
static inline void alt_core_grad(int core, int blob, float distsq, float sr, float tr, float* r3, float* bbb) {
  if (core == 1) {
    const float r2 = blob ? distsq + sr * sr + tr * tr : distsq + sr * sr;
    *r3 = inv_pow_1p5(r2);
    *bbb = -3.0f * (*r3) * (1.0f / r2);
  } else if (core == 2) {
    const float dist = sqrtf(distsq);
    const float corefac = blob ? 1.0f / (sr * sr * sr + tr * tr * tr) : 1.0f / (sr * sr * sr);
    const float d3 = distsq * dist;
    const float reld3 = d3 * corefac;
    const float ood3 = 1.0f / d3;
    *r3 = exp_cond(ood3, corefac, reld3);
    *bbb = exp_bbb(*r3, corefac, reld3, dist, distsq);
  } else if (core == 3) {
    const float s2 = sr * sr, t2 = tr * tr;
    const float denom = blob ? distsq * distsq + s2 * s2 + t2 * t2 : distsq * distsq + s2 * s2;
    *r3 = inv_pow_0p75(denom);
    *bbb = -3.0f * (*r3) * (1.0f / sqrtf(denom));
  } else if (blob) wl_core_blob_grad(distsq, sr, tr, r3, bbb);
  else wl_core_point_grad(distsq, sr, r3, bbb);
}

/* particles -> points as o3d_oracle_pts_on_pts, for a build of the reference with core function `core` */
void o3d_oracle_pts_on_pts_core(int core, int64_t ns, const float* sx, const float* sy, const float* sz, const float* sr,
                                const float* ssx, const float* ssy, const float* ssz,
                                int64_t nt, const float* tx, const float* ty, const float* tz, const float* tr,
                                float* tu, float* tug) {
  const int nacc = tug ? 12 : 3;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nt; ++i) {
    double acc[12] = {0};
    const float trad = tr ? tr[i] : 0.0f;
    for (int64_t j = 0; j < ns; ++j) {
      const src_t s = {sx[j], sy[j], sz[j], sr[j], ssx[j], ssy[j], ssz[j], 0.0f};
      const float dx = tx[i] - s.x, dy = ty[i] - s.y, dz = tz[i] - s.z;
      const float distsq = dx * dx + dy * dy + dz * dz;
      if (tug) {
        float r3, bbb;
        alt_core_grad(core, tr != NULL, distsq, s.r, trad, &r3, &bbb);
        grad_body(&s, dx, dy, dz, r3, bbb, 0, acc);
      } else {
        const float k = alt_core(core, tr != NULL, distsq, s.r, trad);
        const float cx = dz * s.wy - dy * s.wz, cy = dx * s.wz - dz * s.wx, cz = dy * s.wx - dx * s.wy;
        acc[0] += (double)(k * cx); acc[1] += (double)(k * cy); acc[2] += (double)(k * cz);
      }
    }
    flush_acc(nacc, acc, nt, i, tu, tug, 1.0);
  }
}

/* src/Influence.h:728-775 (grads) and :826-864 (vel only); blob-target branches :989-1094 are the
 * same arithmetic (target radius is never used by panel kernels). Sheet strength = ts/area. */
void o3d_oracle_pan_on_pts(int64_t np, const float* nx, const float* ny, const float* nz, const uint32_t* idx,
                           const float* ts, const float* area, const float* sss,
                           int64_t nt, const float* tx, const float* ty, const float* tz,
                           float* tu, float* tug) {
  const int nacc = tug ? 12 : 3;
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t i = 0; i < nt; ++i) {
    double acc[12] = {0};
    for (int64_t j = 0; j < np; ++j) {
      tri_t p; load_tri(&p, nx, ny, nz, idx, j);
      rk_2vs_0(&p, ts[j] / area[j], ts[np + j] / area[j], ts[2 * np + j] / area[j], sss ? sss[j] : 0.0f,
               tx[i], ty[i], tz[i], area[j], 0, tug != NULL, acc);
    }
    flush_acc(nacc, acc, nt, i, tu, tug, 1.0);
  }
}

/* src/Influence.h:1186-1214: panel geometry plays "source", particle position plays "target",
 * particle strength / panel area is the sheet strength, result subtracted from pu. */
void o3d_oracle_pts_on_pan(int64_t ns, const float* sx, const float* sy, const float* sz,
                           const float* ssx, const float* ssy, const float* ssz,
                           int64_t np, const float* nx, const float* ny, const float* nz, const uint32_t* idx,
                           const float* area, float* pu) {
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t i = 0; i < np; ++i) {
    double acc[3] = {0};
    tri_t p; load_tri(&p, nx, ny, nz, idx, i);
    for (int64_t j = 0; j < ns; ++j)
      rk_2vs_0(&p, ssx[j] / area[i], ssy[j] / area[i], ssz[j] / area[i], 0.0f, sx[j], sy[j], sz[j], area[i], 0, 0, acc);
    flush_acc(3, acc, np, i, pu, NULL, -1.0);
  }
}

/* src/Coefficients.h:214-446 (scalar arm :327-411). Column block j = source panel: unit sheet strength
 * along its x1, x2, then unit source; rows = target panel i projected on (t1, t2, n). */
void o3d_oracle_pan_on_pan_coeff(int64_t nsp, const float* snx, const float* sny, const float* snz,
                                 const uint32_t* sidx, const float* sb1, const float* sb2, const float* sarea,
                                 int64_t ntp, const float* tnx, const float* tny, const float* tnz,
                                 const uint32_t* tidx, const float* tb1, const float* tb2, const float* tnrm,
                                 const float* tarea, int self, float* coeffs) {
  const int64_t nrows = 3 * ntp;
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t j = 0; j < nsp; ++j) {
    tri_t p; load_tri(&p, snx, sny, snz, sidx, j);
    const float dir[3][4] = {{sb1[j], sb1[nsp + j], sb1[2 * nsp + j], 0.0f},
                             {sb2[j], sb2[nsp + j], sb2[2 * nsp + j], 0.0f},
                             {0.0f, 0.0f, 0.0f, 1.0f}};
    for (int64_t i = 0; i < ntp; ++i) {
      tri_t q; load_tri(&q, tnx, tny, tnz, tidx, i);
      for (int k = 0; k < 3; ++k) {
        float r[3] = {0.0f, 0.0f, 0.0f};
        rk_2vs_2(&p, dir[k][0], dir[k][1], dir[k][2], dir[k][3], &q, sarea[j], tarea[i], 0, r);
        float* col = coeffs + (3 * j + k) * nrows + 3 * i;
        col[0] = r[0] * tb1[i] + r[1] * tb1[ntp + i] + r[2] * tb1[2 * ntp + i];
        col[1] = r[0] * tb2[i] + r[1] * tb2[ntp + i] + r[2] * tb2[2 * ntp + i];
        col[2] = r[0] * tnrm[i] + r[1] * tnrm[ntp + i] + r[2] * tnrm[2 * ntp + i];
      }
    }
    if (self) { /* :414-436 */
      float* d0 = coeffs + (3 * j) * nrows + 3 * j;
      d0[0] = 0.0f; d0[1] = (float)(2.0 * M_PI); d0[2] = 0.0f;
      float* d1 = d0 + nrows;
      d1[0] = (float)(-2.0 * M_PI); d1[1] = 0.0f; d1[2] = 0.0f;
      float* d2 = d1 + nrows;
      d2[0] = 0.0f; d2[1] = 0.0f; d2[2] = (float)(2.0 * M_PI);
    }
  }
  const float fac = (float)(1.0 / (4.0 * M_PI)); /* :448-451 */
  for (int64_t k = 0; k < nrows * 3 * nsp; ++k) coeffs[k] = coeffs[k] * fac;
}

void o3d_oracle_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int o3d_oracle_max_threads(void) { return omp_get_max_threads(); }

/* =====================================================================================================
 * Convection: the O(N) steps around the influence sums. Every expression keeps the reference's operand
 * types (S = float; dt, weights, fs = double) and its operation order; -ffp-contract=off keeps them unfused.
 * ===================================================================================================== */
#include <stdlib.h>
#include <string.h>

/* src/ElementBase.h:187-192 (u = fs + u * factor, factor a double) and src/Points.h:269-276 (grads * S factor) */
void o3d_oracle_finalize_vels(int64_t n, float* u, float* ug, const double* fs) {
  const double factor = 0.25 / M_PI;
  for (int d = 0; d < 3; ++d)
    for (int64_t i = 0; i < n; ++i) u[d * n + i] = (float)(fs[d] + u[d * n + i] * factor);
  if (ug) {
    const float ff = (float)(0.25 / M_PI);
    for (int64_t k = 0; k < 9 * n; ++k) ug[k] = ug[k] * ff;
  }
}

/* w . grad u, src/Points.h:316-318 */
static inline void stretch_term(const float* ug, int64_t n, int64_t i, const float s[3], float wdu[3]) {
  for (int k = 0; k < 3; ++k) wdu[k] = s[0] * ug[k * n + i] + s[1] * ug[(3 + k) * n + i] + s[2] * ug[(6 + k) * n + i];
}

/* Points::move with 1, 2 or 3 stages: src/ElementBase.h:253-336 (advection), src/Points.h:288-520 (stretch).
 * u[k] is 3 x n, ug[k] 9 x n or NULL. order 1 stretches with ug[0] = the moving object's own gradient.
 * uout (3 x n, may be NULL or alias u[0]) receives the combined velocity for order >= 2. elong may be NULL. */
void o3d_oracle_move(int order, int64_t n, double dt, const double* wt, const float* const* u, const float* const* ug,
                     float* x, float* s, float* elong, float* uout) {
  const float dtf = (float)dt;
  int have = s != NULL;
  for (int k = 0; k < order; ++k) have = have && ug[k] != NULL;
  for (int64_t i = 0; i < n; ++i) {
    float un[3];
    if (order == 1) {
      for (int d = 0; d < 3; ++d) x[d * n + i] = (float)(x[d * n + i] + dtf * wt[0] * u[0][d * n + i]);   /* :264 */
    } else {
      for (int d = 0; d < 3; ++d) {
        double c = wt[0] * u[0][d * n + i] + wt[1] * u[1][d * n + i];                                     /* :287 */
        if (order == 3) c = c + wt[2] * u[2][d * n + i];                                                  /* :319 */
        un[d] = (float)c;
      }
      for (int d = 0; d < 3; ++d) {
        x[d * n + i] = x[d * n + i] + dtf * un[d];                                                        /* :294,326 */
        if (uout) uout[d * n + i] = un[d];
      }
    }
    if (!have) continue;
    const float ts[3] = {s[i], s[n + i], s[2 * n + i]};
    float wdu[3];
    stretch_term(ug[0], n, i, ts, wdu);
    if (order >= 2) {
      float w2[3], w3[3] = {0, 0, 0};
      stretch_term(ug[1], n, i, ts, w2);
      if (order == 3) stretch_term(ug[2], n, i, ts, w3);
      for (int k = 0; k < 3; ++k) {
        double c = wt[0] * wdu[k] + wt[1] * w2[k];                                                        /* src/Points.h:399-402 */
        if (order == 3) c = c + wt[2] * w3[k];                                                            /* :498-501 */
        wdu[k] = (float)c;
      }
    }
    const float circ = ts[0] * ts[0] + ts[1] * ts[1] + ts[2] * ts[2];
    if (elong && circ > 0.0f) {
      const float sd = ts[0] * wdu[0] + ts[1] * wdu[1] + ts[2] * wdu[2];
      float ef;
      if (order == 1) ef = (float)(dtf * wt[0] * sd / circ);                                              /* :321 */
      else ef = dtf * sd / circ;                                                                          /* :407,505 */
      elong[i] = (float)(elong[i] * (1.0 + ef));                                                          /* :322 */
    }
    for (int d = 0; d < 3; ++d) {
      if (order == 1) s[d * n + i] = (float)(ts[d] + dt * wt[0] * wdu[d]);                                /* :330-332 */
      else s[d * n + i] = (float)(ts[d] + dt * wdu[d]);                                                   /* :414-416 */
    }
  }
}

/* Convection::find_vels for a lone vortex-particle collection, src/Convection.h:130-184 */
/* core function of the convection sequence below (a build-time choice in the reference, src/CoreFunc.h:35-38) */
static int g_advect_core = 0;
void o3d_oracle_set_advect_core(int core) { g_advect_core = core; }

static void find_vels_one(int64_t n, const float* x, const float* s, const float* r, float* u, float* ug, const double* fs) {
  memset(u, 0, sizeof(float) * 3 * n);
  memset(ug, 0, sizeof(float) * 9 * n);
  if (g_advect_core) o3d_oracle_pts_on_pts_core(g_advect_core, n, x, x + n, x + 2 * n, r, s, s + n, s + 2 * n, n, x, x + n, x + 2 * n, r, u, ug);
  else o3d_oracle_pts_on_pts(n, x, x + n, x + 2 * n, r, s, s + n, s + 2 * n, n, x, x + n, x + 2 * n, r, u, ug);
  o3d_oracle_finalize_vels(n, u, ug, fs);
}

/* nsteps x Convection::advect: order 1 src/Convection.h:232-262, 2 (Ralston) :349-425, 3 :431-556; no boundaries,
 * no field points. x, s 3 x n; r, elong n; u 3 x n and ug 9 x n hold what the collection holds afterwards. */
void o3d_oracle_advect(int order, int nsteps, double dt, const double* fs, int64_t n, float* x, float* s, const float* r,
                       float* elong, float* u, float* ug) {
  float* x1 = malloc(sizeof(float) * 3 * n), *s1 = malloc(sizeof(float) * 3 * n);
  float* u1 = malloc(sizeof(float) * 3 * n), *g1 = malloc(sizeof(float) * 9 * n);
  float* x2 = malloc(sizeof(float) * 3 * n), *s2 = malloc(sizeof(float) * 3 * n);
  float* u2 = malloc(sizeof(float) * 3 * n), *g2 = malloc(sizeof(float) * 9 * n);
  const double one = 1.0;
  for (int step = 0; step < nsteps; ++step) {
    find_vels_one(n, x, s, r, u, ug, fs);
    if (order == 1) {
      const float* uu[1] = {u}; const float* gg[1] = {ug};
      o3d_oracle_move(1, n, dt, &one, uu, gg, x, s, elong, NULL);
    } else if (order == 2) {
      memcpy(x1, x, sizeof(float) * 3 * n); memcpy(s1, s, sizeof(float) * 3 * n);
      const float* uu[2] = {u, u1}; const float* gg[2] = {ug, g1};
      o3d_oracle_move(1, n, (2.0 / 3.0) * dt, &one, uu, gg, x1, s1, NULL, NULL);   /* interim copy's elongation is discarded */
      find_vels_one(n, x1, s1, r, u1, g1, fs);
      const double wt[2] = {0.25, 0.75};
      o3d_oracle_move(2, n, dt, wt, uu, gg, x, s, elong, u);
    } else {
      memcpy(x1, x, sizeof(float) * 3 * n); memcpy(s1, s, sizeof(float) * 3 * n);
      const float* uu[3] = {u, u1, u2}; const float* gg[3] = {ug, g1, g2};
      o3d_oracle_move(1, n, 0.5 * dt, &one, uu, gg, x1, s1, NULL, NULL);
      find_vels_one(n, x1, s1, r, u1, g1, fs);
      memcpy(x2, x, sizeof(float) * 3 * n); memcpy(s2, s, sizeof(float) * 3 * n);
      const float* u_1[1] = {u1}; const float* g_0[1] = {ug};   /* vort2.move(.., vort1): vort1's velocity, its OWN (= the original's) gradient */
      o3d_oracle_move(1, n, 0.75 * dt, &one, u_1, g_0, x2, s2, NULL, NULL);
      find_vels_one(n, x2, s2, r, u2, g2, fs);
      const double wt[3] = {2.0 / 9.0, 3.0 / 9.0, 4.0 / 9.0};
      o3d_oracle_move(3, n, dt, wt, uu, gg, x, s, elong, u);
    }
  }
  free(x1); free(s1); free(u1); free(g1); free(x2); free(s2); free(u2); free(g2);
}

/* ElementBase::get_max_str (src/ElementBase.h:339-351) and Points::get_max_elong (src/Points.h:523-532) */
void o3d_oracle_stats(int64_t n, const float* s, const float* elong, float* max_str, float* max_elong) {
  float ms = 0.0f, me = 0.0f;
  for (int64_t i = 0; i < n; ++i) {
    const float t = s[i] * s[i] + s[n + i] * s[n + i] + s[2 * n + i] * s[2 * n + i];
    if (t > ms) ms = t;
    if (elong[i] > me) me = elong[i];
  }
  *max_str = sqrtf(ms);
  *max_elong = me;
}

/* ElementBase::get_total_circ (src/ElementBase.h:354-378: std::accumulate with a DOUBLE init value over the float strengths, so
 * a sequential double sum rounded to float once; the "< 40000" test reads the size of the 3-array, always the first branch)
 * and Points::get_total_impulse (src/Points.h:547-563: float terms summed sequentially in float). x, s: 3 x n rows. */
void o3d_oracle_totals(int64_t n, const float* x, const float* s, float* circ, float* impulse) {
  for (int d = 0; d < 3; ++d) {
    double acc = 0.0;
    for (int64_t i = 0; i < n; ++i) acc += s[d * n + i];
    circ[d] = (float)acc;
  }
  float i0 = 0.0f, i1 = 0.0f, i2 = 0.0f;
  for (int64_t i = 0; i < n; ++i) {
    i0 += s[n + i] * x[2 * n + i] - s[2 * n + i] * x[n + i];
    i1 += s[2 * n + i] * x[i] - s[i] * x[2 * n + i];
    i2 += s[i] * x[n + i] - s[n + i] * x[i];
  }
  impulse[0] = i0; impulse[1] = i1; impulse[2] = i2;
}

/* =====================================================================================================
 * Particle x panel closest-point loops (src/Reflect.h). Float throughout; "1.0 / x" is a double division
 * stored into a float, as in the reference.
 * ===================================================================================================== */
#include <float.h>

typedef struct { float distsq, cpx, cpy, cpz; } closest_t;

static inline float dot3(const float a[3], const float b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; } /* src/MathHelper.h:195-198 */
static inline void cross3(const float a[3], const float b[3], float r[3]) {                                       /* :208-214 */
  r[0] = a[1] * b[2] - a[2] * b[1];
  r[1] = a[2] * b[0] - a[0] * b[2];
  r[2] = a[0] * b[1] - a[1] * b[0];
}

/* one edge test of panel_point_distance, src/Reflect.h:107-159 */
static inline void closest_edge(const float a[3], const float e[3], const float dt[3], closest_t* r) {
  const float inv = (float)(1.0 / dot3(e, e));
  float rx[3];
  cross3(e, dt, rx);
  const float d = dot3(rx, rx) * inv;
  if (d < r->distsq) {
    const float t = dot3(e, dt) * inv;
    if (0.0 < t && t < 1.0) {
      r->distsq = d;
      r->cpx = a[0] + t * e[0]; r->cpy = a[1] + t * e[1]; r->cpz = a[2] + t * e[2];
    }
  }
}

/* src/Reflect.h:55-188 */
static closest_t panel_point_distance(const float s0[3], const float s1[3], const float s2[3], const float n[3], const float t[3]) {
  closest_t r = {9.9e+9f, 0.f, 0.f, 0.f};
  const float dt0[3] = {t[0] - s0[0], t[1] - s0[1], t[2] - s0[2]};
  const float d0 = dot3(dt0, dt0);
  if (d0 < r.distsq) { r.distsq = d0; r.cpx = s0[0]; r.cpy = s0[1]; r.cpz = s0[2]; }
  const float dt1[3] = {t[0] - s1[0], t[1] - s1[1], t[2] - s1[2]};
  const float d1 = dot3(dt1, dt1);
  if (d1 < r.distsq) { r.distsq = d1; r.cpx = s1[0]; r.cpy = s1[1]; r.cpz = s1[2]; }
  const float dt2[3] = {t[0] - s2[0], t[1] - s2[1], t[2] - s2[2]};
  const float d2 = dot3(dt2, dt2);
  if (d2 < r.distsq) { r.distsq = d2; r.cpx = s2[0]; r.cpy = s2[1]; r.cpz = s2[2]; }
  const float e01[3] = {s1[0] - s0[0], s1[1] - s0[1], s1[2] - s0[2]};
  closest_edge(s0, e01, dt0, &r);
  const float e12[3] = {s2[0] - s1[0], s2[1] - s1[1], s2[2] - s1[2]};
  closest_edge(s1, e12, dt1, &r);
  const float e20[3] = {s0[0] - s2[0], s0[1] - s2[1], s0[2] - s2[2]};
  closest_edge(s2, e20, dt2, &r);
  float in[3];
  cross3(n, e01, in); const float in01 = dot3(dt0, in);
  cross3(n, e12, in); const float in12 = dot3(dt1, in);
  cross3(n, e20, in); const float in20 = dot3(dt2, in);
  if (in01 > 0.0 && in12 > 0.0 && in20 > 0.0) {
    const float td = dot3(dt0, n);
    r.distsq = td * td;
    r.cpx = t[0] - n[0] * td; r.cpy = t[1] - n[1] * td; r.cpz = t[2] - n[2] * td;
  }
  return r;
}

/* mode 0: reflect_panp2 (src/Reflect.h:194-311); mode 1: clear_inner_panp2 with _method 1 (:446-620).
 * nodes SoA, idx 3 per panel, nrm SoA 3 x np; x 3 x nt SoA in/out. Returns the number of particles moved. */
int64_t o3d_oracle_closest_pass(int mode, int64_t np, const float* nx, const float* ny, const float* nz, const uint32_t* idx,
                                const float* nrm, int64_t nt, float* x, float cutoff_mult, float ips) {
  const float eps = 10.0f * FLT_EPSILON;
  int64_t moved = 0;
#pragma omp parallel for schedule(static) reduction(+ : moved)
  for (int64_t i = 0; i < nt; ++i) {
    const float t[3] = {x[i], x[nt + i], x[2 * nt + i]};
    float mindist = FLT_MAX, cnt = 0.0f, dfirst = 0.0f;
    float ns[3] = {0, 0, 0}, cs[3] = {0, 0, 0};
    for (int64_t j = 0; j < np; ++j) {
      const uint32_t a = idx[3 * j], b = idx[3 * j + 1], c = idx[3 * j + 2];
      const float s0[3] = {nx[a], ny[a], nz[a]}, s1[3] = {nx[b], ny[b], nz[b]}, s2[3] = {nx[c], ny[c], nz[c]};
      const float n[3] = {nrm[j], nrm[np + j], nrm[2 * np + j]};
      const closest_t r = panel_point_distance(s0, s1, s2, n, t);
      if (r.distsq < mindist - eps) {                 /* :226-234 the hit list restarts */
        mindist = r.distsq; cnt = 1.0f; dfirst = r.distsq;
        ns[0] = n[0]; ns[1] = n[1]; ns[2] = n[2];
        cs[0] = r.cpx; cs[1] = r.cpy; cs[2] = r.cpz;
      } else if (r.distsq < mindist + eps) {           /* :236-242 a tie joins it */
        cnt = cnt + 1.0f;
        ns[0] = ns[0] + n[0]; ns[1] = ns[1] + n[1]; ns[2] = ns[2] + n[2];
        cs[0] = cs[0] + r.cpx; cs[1] = cs[1] + r.cpy; cs[2] = cs[2] + r.cpz;
      }
    }
    if (cnt == 0.0f) continue;
    const float len = (float)(1.0 / sqrtf(ns[0] * ns[0] + ns[1] * ns[1] + ns[2] * ns[2]));   /* normalizeVec */
    const float m[3] = {ns[0] * len, ns[1] * len, ns[2] * len};
    const float cp[3] = {cs[0] / cnt, cs[1] / cnt, cs[2] / cnt};
    const float dx[3] = {t[0] - cp[0], t[1] - cp[1], t[2] - cp[2]};
    if (mode == 0) {
      const float dotp = dot3(m, dx);
      if (dotp < 0.0) {
        const float dist = sqrtf(dfirst);
        for (int d = 0; d < 3; ++d) x[d * nt + i] = cp[d] + dist * m[d];
        moved += 1;
      }
    } else {
      const float dotp = dot3(m, dx) - cutoff_mult * ips;
      if (dotp < 0.0) {
        for (int d = 0; d < 3; ++d) x[d * nt + i] = t[d] - dotp * m[d];
        moved += 1;
      }
    }
  }
  return moved;
}

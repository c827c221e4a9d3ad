// biot_panel.cuh - flat triangular panels <-> points for sm_100a (B200).
//
// Replaces the loop nests of panels_affect_points<S,A> (reference src/Influence.h:728-775, :826-864 and
// the identical blob-target arms :952-1094), points_affect_panels<S,A> (:1186-1214) and
// panels_on_panels_coeff<S> (src/Coefficients.h:327-446), together with the recursive kernels they call:
// rkernel_2vs_0p / rkernel_2vs_0pg (src/Kernels.h:1028-1211) and rkernel_2vs_2p (:1217-1315), whose
// leaves are kernel_0vs_0p / kernel_0vs_0pg (:115-133, :294-345) with a zero source radius.
//
// What the reference computes per (panel, point) pair: an adaptive 4-way midpoint subdivision, to at
// most 3 levels, that stops at a node when |point - centroid| > 4 sqrt(area); every stopped node adds
// one singular point-vortex(+source) interaction from its centroid carrying area-fraction of the panel
// strength. Panel-panel pairs subdivide both triangles (16 children), size = sqrt(sa) + sqrt(ta).
//
// Design:
//   * The stop predicate is a hard branch that changes how many leaves are summed, so it is evaluated
//     bit-for-bit as the reference's scalar build rounds it: unfused __fmul_rn/__fadd_rn in the
//     reference's operand order, IEEE __fdiv_rn for the /3 and __fsqrt_rn for the distance. The
//     threshold 4*sqrt(area*4^-l) equals 2^-l * 4*sqrt(area) exactly, so it is formed once per panel.
//   * The recursion is unrolled into three statically nested loops (no stack, no local memory): the
//     parent's three edge midpoints are formed once and a child is a register select.
//   * Leaves run in FP32 with MUFU.RSQ: r3 = rs^3, bbb = -3 rs^5 (the WL core at zero radius, src/
//     CoreFunc.h:255-259,279-287). Strengths enter as totals: the reference's (ts/area)*area round trip
//     (src/Influence.h:736, Kernels.h:1043-1048) differs from ts by <= 1 ulp.
//   * A thread owns one target (a point for panels->points, a panel for points->panels and for the
//     coefficient block) and its sums: FP32 across one shared-memory tile of sources, FP64 across tiles,
//     the reference's float-kernel/double-accumulator scheme. Small target counts split the source
//     range over gridDim.y; each slice stores its FP64 partials in its own slab and pp_finish_kernel adds
//     the slabs in slice order (deterministic), as in biot_pp.cuh.
//   * Leaf/split counters (for the reference's flop report) are reduced per warp with
//     __reduce_add_sync and added once per warp.
#pragma once
#include "biot_pp.cuh"

namespace o3d {

constexpr int kPanRec = 5;        // float4 per packed panel record (80 bytes)
constexpr int kPanTile = 64;      // panels per shared-memory tile (5 KB)
constexpr int kMaxLev = 3;        // RECURSIVE_LEVELS, src/Influence.h:23, src/Coefficients.h:23

struct Tri {
  float x0, y0, z0, x1, y1, z1, x2, y2, z2;
};

// s / 3 correctly rounded (== __fdiv_rn(s, 3.0f), the reference's IEEE division) in three FP32 instructions instead of the
// ~9 of the generic division (FCHK, MUFU-seeded FFMA chain, convergence barrier around a slow-path call): with
// y = RN(1/3), q = RN(s y) is a faithful quotient (|s y - s/3| <= 2^-25 |s/3|: at most a quarter ulp off before rounding),
// r = s - 3 q is exact in one FMA, and q + r y rounds to RN(s/3) (Markstein's theorem). Zeros, infinities and the subnormal
// range aside - an exhaustive sweep of all 2^32 inputs on the device (csrc/microbench/div3_check.cu,
// profiles/r01_div3_check.txt) finds no other difference; -0 gives +0, which a centroid coordinate cannot tell apart.
__device__ __forceinline__ float div3_rn(float s) {
  const float y = 0.3333333432674407958984375f;   // 0x3EAAAAAB
  const float q = __fmul_rn(s, y);
  const float r = __fmaf_rn(-3.0f, q, s);
  return __fmaf_rn(r, y, q);
}
// reference operand order: (a + b + c) / 3  (src/Kernels.h:1052-1054)
__device__ __forceinline__ float third_sum(float a, float b, float c) {
  return div3_rn(__fadd_rn(__fadd_rn(a, b), c));
}
__device__ __forceinline__ float mid(float a, float b) { return __fmul_rn(0.5f, __fadd_rn(a, b)); }
// my_dist (src/Kernels.h:979-982): sqrt(dx*dx + dy*dy + dz*dz), unfused, left to right
__device__ __forceinline__ float sumsq_rn(float dx, float dy, float dz) {
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// sumsq_rn for two lanes at once. The products are packed; the sums are NOT: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into
// FFMA2 (it leaves the scalar .rn forms alone, and -fmad=false does not stop it), which would change the bits of the squared
// distance and with them the reference's stop decision. Scalar adds on the halves of packed products stay unfused (checked in
// the SASS: FMUL2, FMUL2, FADD, FADD, FMUL2, FADD, FADD; and by the leaf-for-leaf counts of tests/test_gpu_parity.py).
__device__ __forceinline__ float2 sumsq2_rn(float2 dx, float2 dy, float2 dz) {
  const float2 xx = __fmul2_rn(dx, dx), yy = __fmul2_rn(dy, dy), zz = __fmul2_rn(dz, dz);
  return make_float2(__fadd_rn(__fadd_rn(xx.x, yy.x), zz.x), __fadd_rn(__fadd_rn(xx.y, yy.y), zz.y));
}

// The stop predicate without the square root. The reference tests  my_dist = sqrt(dx^2+dy^2+dz^2) > 4 sqrt(area)  with a
// correctly rounded sqrt (src/Kernels.h:979-1002); that function is monotonic, so for every threshold thr there is one float
// T = max { x : sqrt_rn(x) <= thr }  with  sqrt_rn(x) > thr  <=>  x > T  for all x >= 0: the SAME decision, bit for bit, from
// the squared distance the leaf needs anyway, without an IEEE square root (MUFU + fix-up, ~8 instructions) per visited node.
// T is found once per panel (pan_pack_kernel) by stepping from fl(thr*thr) through neighbouring floats; a level down, the
// threshold halves and T quarters, both exactly (sqrt_rn(4^-l x) = 2^-l sqrt_rn(x)). Checked by brute force over the
// neighbourhood of T for thousands of thresholds (tests/test_abi.py::test_squared_threshold_identity), and on the device by
// the leaf-for-leaf flop counts of the golden panels (tests/test_gpu_parity.py::test_panels_golden).
__device__ __forceinline__ float sq_threshold(float thr) {
  float x = __fmul_rn(thr, thr);
  for (int it = 0; it < 8 && x > 0.0f && __fsqrt_rn(x) > thr; ++it) x = __uint_as_float(__float_as_uint(x) - 1u);
  for (int it = 0; it < 8; ++it) {
    const float nx = __uint_as_float(__float_as_uint(x) + 1u);
    if (!(__fsqrt_rn(nx) <= thr)) break;
    x = nx;
  }
  return x;
}

struct Mids {
  float ax, ay, az;  // mid(v0,v1)
  float bx, by, bz;  // mid(v0,v2)
  float cx, cy, cz;  // mid(v1,v2)
};
__device__ __forceinline__ Mids tri_mids(const Tri& p) {
  Mids m;
  m.ax = mid(p.x0, p.x1); m.ay = mid(p.y0, p.y1); m.az = mid(p.z0, p.z1);
  m.bx = mid(p.x0, p.x2); m.by = mid(p.y0, p.y2); m.bz = mid(p.z0, p.z2);
  m.cx = mid(p.x1, p.x2); m.cy = mid(p.y1, p.y2); m.cz = mid(p.z1, p.z2);
  return m;
}
// children {0,1,3},{1,2,4},{1,4,3},{3,4,5} of nodes [v0, m01, v1, m02, m12, v2] (src/Kernels.h:1081-1089)
__device__ __forceinline__ Tri tri_child(const Tri& p, const Mids& m, int k) {
  Tri c;
  if (k == 0)      c = Tri{p.x0, p.y0, p.z0, m.ax, m.ay, m.az, m.bx, m.by, m.bz};
  else if (k == 1) c = Tri{m.ax, m.ay, m.az, p.x1, p.y1, p.z1, m.cx, m.cy, m.cz};
  else if (k == 2) c = Tri{m.ax, m.ay, m.az, m.cx, m.cy, m.cz, m.bx, m.by, m.bz};
  else             c = Tri{m.bx, m.by, m.bz, m.cx, m.cy, m.cz, p.x2, p.y2, p.z2};
  return c;
}

// Accumulator layout for panel -> point: [0..2] u v w | [3..11] d_j * (bbb e_i) | [12..14] sum w r3 |
// [15] sum q r3 (the isotropic part of the source-sheet gradient, src/Kernels.h:327-344)
template <bool GRAD> struct PanAcc { static constexpr int N = GRAD ? 16 : 3; };

// one leaf: singular vortex (wx,wy,wz) + source q at (cx,cy,cz) acting on the point (tx,ty,tz)
template <bool GRAD>
__device__ __forceinline__ void pan_leaf(float dx, float dy, float dz, float distsq, float wx, float wy, float wz,
                                         float q, float (&acc)[PanAcc<GRAD>::N]) {
  const float rs = rsqrt_approx(distsq);
  const float rs2 = rs * rs;
  const float r3 = rs2 * rs;
  float ex = fmaf(dz, wy, -(dy * wz));
  float ey = fmaf(dx, wz, -(dz * wx));
  float ez = fmaf(dy, wx, -(dx * wy));
  ex = fmaf(dx, q, ex); ey = fmaf(dy, q, ey); ez = fmaf(dz, q, ez);
  acc[0] = fmaf(r3, ex, acc[0]);
  acc[1] = fmaf(r3, ey, acc[1]);
  acc[2] = fmaf(r3, ez, acc[2]);
  if constexpr (GRAD) {
    const float bbb = -3.0f * (r3 * rs2);
    ex *= bbb; ey *= bbb; ez *= bbb;
    acc[3]  = fmaf(dx, ex, acc[3]);
    acc[4]  = fmaf(dx, ey, acc[4]);
    acc[5]  = fmaf(dx, ez, acc[5]);
    acc[6]  = fmaf(dy, ex, acc[6]);
    acc[7]  = fmaf(dy, ey, acc[7]);
    acc[8]  = fmaf(dy, ez, acc[8]);
    acc[9]  = fmaf(dz, ex, acc[9]);
    acc[10] = fmaf(dz, ey, acc[10]);
    acc[11] = fmaf(dz, ez, acc[11]);
    acc[12] = fmaf(wx, r3, acc[12]);
    acc[13] = fmaf(wy, r3, acc[13]);
    acc[14] = fmaf(wz, r3, acc[14]);
    acc[15] = fmaf(q, r3, acc[15]);
  }
}

// Two leaves per instruction (packed FP32, as in biot_pp.cuh): lane halves .x / .y are two source panels against one point.
// rs = 0 in a half switches that half off (its r3 and bbb vanish and every contribution is 0 * finite). Same operations per
// half as pan_leaf; the two halves' sums meet once per tile.
template <bool GRAD, bool SRC = true>
__device__ __forceinline__ void pan_leaf2(float2 dx, float2 dy, float2 dz, float2 rs, float2 wx, float2 wy, float2 wz,
                                          float2 q, float2 (&acc)[PanAcc<GRAD>::N]) {
  const float2 rs2 = __fmul2_rn(rs, rs);
  const float2 r3 = __fmul2_rn(rs2, rs);
  float2 ex = __ffma2_rn(dz, wy, neg2(__fmul2_rn(dy, wz)));
  float2 ey = __ffma2_rn(dx, wz, neg2(__fmul2_rn(dz, wx)));
  float2 ez = __ffma2_rn(dy, wx, neg2(__fmul2_rn(dx, wy)));
  if constexpr (SRC) { ex = __ffma2_rn(dx, q, ex); ey = __ffma2_rn(dy, q, ey); ez = __ffma2_rn(dz, q, ez); }   // source sheet
  acc[0] = __ffma2_rn(r3, ex, acc[0]);
  acc[1] = __ffma2_rn(r3, ey, acc[1]);
  acc[2] = __ffma2_rn(r3, ez, acc[2]);
  if constexpr (GRAD) {
    const float2 bbb = __fmul2_rn(f2(-3.0f, -3.0f), __fmul2_rn(r3, rs2));
    ex = __fmul2_rn(ex, bbb); ey = __fmul2_rn(ey, bbb); ez = __fmul2_rn(ez, bbb);
    acc[3]  = __ffma2_rn(dx, ex, acc[3]);
    acc[4]  = __ffma2_rn(dx, ey, acc[4]);
    acc[5]  = __ffma2_rn(dx, ez, acc[5]);
    acc[6]  = __ffma2_rn(dy, ex, acc[6]);
    acc[7]  = __ffma2_rn(dy, ey, acc[7]);
    acc[8]  = __ffma2_rn(dy, ez, acc[8]);
    acc[9]  = __ffma2_rn(dz, ex, acc[9]);
    acc[10] = __ffma2_rn(dz, ey, acc[10]);
    acc[11] = __ffma2_rn(dz, ez, acc[11]);
    acc[12] = __ffma2_rn(wx, r3, acc[12]);
    acc[13] = __ffma2_rn(wy, r3, acc[13]);
    acc[14] = __ffma2_rn(wz, r3, acc[14]);
    acc[15] = __ffma2_rn(q, r3, acc[15]);
  }
}

// Node test + leaf for a triangle whose centroid is (cx,cy,cz): returns true when the node was
// consumed as a leaf (well separated, or deepest level). thrsq = sq_threshold(4 sqrt(area of the node)).
template <bool GRAD>
__device__ __forceinline__ bool pan_node(float cx, float cy, float cz, float thrsq, bool deepest, float tx, float ty,
                                         float tz, float wx, float wy, float wz, float q,
                                         float (&acc)[PanAcc<GRAD>::N]) {
  const float dx = __fsub_rn(tx, cx), dy = __fsub_rn(ty, cy), dz = __fsub_rn(tz, cz);
  const float distsq = sumsq_rn(dx, dy, dz);
  if (distsq > thrsq || deepest) {      // == sqrt_rn(distsq) > thr, see sq_threshold
    pan_leaf<GRAD>(dx, dy, dz, distsq, wx, wy, wz, q, acc);
    return true;
  }
  return false;
}

// s / 3 for two lanes: div3_rn on FMUL2 / FFMA2 (nothing here is a mul feeding an add, so nothing can be contracted)
__device__ __forceinline__ float2 div3_rn2(float2 s) {
  const float2 y = f2(0.3333333432674407958984375f, 0.3333333432674407958984375f);
  const float2 q = __fmul2_rn(s, y);
  const float2 r = __ffma2_rn(f2(-3.0f, -3.0f), q, s);
  return __ffma2_rn(r, y, q);
}
// Two deepest-level children at once: vertex k of the two children in (ak, bk, ck).x/.y per coordinate. Always leaves.
template <bool GRAD>
__device__ __forceinline__ void pan_deepest2(float2 x0, float2 y0, float2 z0, float2 x1, float2 y1, float2 z1, float2 x2, float2 y2,
                                             float2 z2, float tx, float ty, float tz, float wx, float wy, float wz, float q,
                                             float2 (&acc)[PanAcc<GRAD>::N]) {
  const float2 cx = div3_rn2(__fadd2_rn(__fadd2_rn(x0, x1), x2));     // (a + b + c) / 3 in the reference's order
  const float2 cy = div3_rn2(__fadd2_rn(__fadd2_rn(y0, y1), y2));
  const float2 cz = div3_rn2(__fadd2_rn(__fadd2_rn(z0, z1), z2));
  const float2 m1 = f2(-1.0f, -1.0f);
  const float2 dx = __ffma2_rn(cx, m1, f2(tx, tx)), dy = __ffma2_rn(cy, m1, f2(ty, ty)), dz = __ffma2_rn(cz, m1, f2(tz, tz));   // t - c, exactly
  const float2 d2 = sumsq2_rn(dx, dy, dz);
  const float2 rs = f2(rsqrt_approx(d2.x), rsqrt_approx(d2.y));
  pan_leaf2<GRAD>(dx, dy, dz, rs, f2(wx, wx), f2(wy, wy), f2(wz, wz), f2(q, q), acc);
}

// Levels 1..3 of a (panel, point) pair whose level-0 node was NOT well separated. (wx,wy,wz,q) = total panel
// strengths; thr0 = sq_threshold(4 sqrt(area)), the level-0 squared threshold. counts[0] += leaves, counts[1] += splits (the level-0 split included).
template <bool GRAD>
__device__ __forceinline__ void pan_subdivide(const Tri& p0, float thr0, float wx, float wy, float wz, float q, float tx,
                                              float ty, float tz, float (&acc)[PanAcc<GRAD>::N], unsigned (&counts)[2]) {
  counts[1] += 1;
  float2 acc2[PanAcc<GRAD>::N];                     // the deepest level's leaves, two per instruction
#pragma unroll
  for (int k = 0; k < PanAcc<GRAD>::N; ++k) acc2[k] = f2(0.f, 0.f);
  const Mids m0 = tri_mids(p0);
  const float w1x = wx * 0.25f, w1y = wy * 0.25f, w1z = wz * 0.25f, q1 = q * 0.25f, thr1 = thr0 * 0.25f;
#pragma unroll 1   // (unrolling this level too costs registers: 255 + spills in pts_pan_kernel)
  for (int k1 = 0; k1 < 4; ++k1) {
    const Tri p1 = tri_child(p0, m0, k1);
    if (pan_node<GRAD>(third_sum(p1.x0, p1.x1, p1.x2), third_sum(p1.y0, p1.y1, p1.y2), third_sum(p1.z0, p1.z1, p1.z2),
                       thr1, false, tx, ty, tz, w1x, w1y, w1z, q1, acc)) {
      counts[0] += 1;
      continue;
    }
    counts[1] += 1;
    const Mids m1 = tri_mids(p1);
    const float w2x = w1x * 0.25f, w2y = w1y * 0.25f, w2z = w1z * 0.25f, q2 = q1 * 0.25f, thr2 = thr1 * 0.25f;
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) {
      const Tri p2 = tri_child(p1, m1, k2);
      if (pan_node<GRAD>(third_sum(p2.x0, p2.x1, p2.x2), third_sum(p2.y0, p2.y1, p2.y2),
                         third_sum(p2.z0, p2.z1, p2.z2), thr2, false, tx, ty, tz, w2x, w2y, w2z, q2, acc)) {
        counts[0] += 1;
        continue;
      }
      counts[1] += 1;
      const Mids m2 = tri_mids(p2);
      const float w3x = w2x * 0.25f, w3y = w2y * 0.25f, w3z = w2z * 0.25f, q3 = q2 * 0.25f;
      // deepest level, most of the visited nodes, and every one of them a leaf
      if constexpr (!GRAD) {
        // velocity only: children {0,1} and {2,3} go through the packed arithmetic two at a time (centroids by the exact packed /3,
        // d = t - c as fma(c, -1, t), unfused squared distance). With gradients the 16 extra packed accumulators cost a resident
        // CTA per SM (192 registers) or spills, and the gain is gone (measured: 3.19 -> 3.57 / 3.28 ms at 5 120 x 262 144).
        pan_deepest2<GRAD>(f2(p2.x0, m2.ax), f2(p2.y0, m2.ay), f2(p2.z0, m2.az), f2(m2.ax, p2.x1), f2(m2.ay, p2.y1), f2(m2.az, p2.z1),
                           f2(m2.bx, m2.cx), f2(m2.by, m2.cy), f2(m2.bz, m2.cz), tx, ty, tz, w3x, w3y, w3z, q3, acc2);
        pan_deepest2<GRAD>(f2(m2.ax, m2.bx), f2(m2.ay, m2.by), f2(m2.az, m2.bz), f2(m2.cx, m2.cx), f2(m2.cy, m2.cy), f2(m2.cz, m2.cz),
                           f2(m2.bx, p2.x2), f2(m2.by, p2.y2), f2(m2.bz, p2.z2), tx, ty, tz, w3x, w3y, w3z, q3, acc2);
      } else {
#pragma unroll   // unrolled so that each child is a static choice of registers, not selects
        for (int k3 = 0; k3 < 4; ++k3) {
          const Tri p3 = tri_child(p2, m2, k3);
          pan_node<GRAD>(third_sum(p3.x0, p3.x1, p3.x2), third_sum(p3.y0, p3.y1, p3.y2), third_sum(p3.z0, p3.z1, p3.z2),
                         0.0f, true, tx, ty, tz, w3x, w3y, w3z, q3, acc);
        }
      }
      counts[0] += 4;
    }
  }
  if constexpr (!GRAD) {
#pragma unroll
    for (int k = 0; k < PanAcc<GRAD>::N; ++k) acc[k] += acc2[k].x + acc2[k].y;
  }
}

// Whole (panel, point) pair: level-0 node, then the subdivision if it was not well separated.
// (thr0 is the SQUARED level-0 threshold, sq_threshold(4 sqrt(area)).)
template <bool GRAD>
__device__ __forceinline__ void pan_on_point(const Tri& p0, float c0x, float c0y, float c0z, float thr0, float wx,
                                             float wy, float wz, float q, float tx, float ty, float tz,
                                             float (&acc)[PanAcc<GRAD>::N], unsigned (&counts)[2]) {
  if (pan_node<GRAD>(c0x, c0y, c0z, thr0, false, tx, ty, tz, wx, wy, wz, q, acc)) {
    counts[0] += 1;
    return;
  }
  pan_subdivide<GRAD>(p0, thr0, wx, wy, wz, q, tx, ty, tz, acc, counts);
}

// Fold a tile's FP32 partials into the FP64 sums (u v w | ux vx wx | uy vy wy | uz vz wz) and clear them.
template <bool GRAD>
__device__ __forceinline__ void pan_promote(float (&acc)[PanAcc<GRAD>::N], double (&sum)[GRAD ? 12 : 3]) {
  sum[0] += (double)acc[0]; sum[1] += (double)acc[1]; sum[2] += (double)acc[2];
  if constexpr (GRAD) {
    const float ax = acc[12], ay = acc[13], az = acc[14], s = acc[15];
    sum[3]  += (double)(acc[3] + s);
    sum[4]  += (double)(acc[4] + az);
    sum[5]  += (double)(acc[5] - ay);
    sum[6]  += (double)(acc[6] - az);
    sum[7]  += (double)(acc[7] + s);
    sum[8]  += (double)(acc[8] + ax);
    sum[9]  += (double)(acc[9] + ay);
    sum[10] += (double)(acc[10] - ax);
    sum[11] += (double)(acc[11] + s);
  }
#pragma unroll
  for (int k = 0; k < PanAcc<GRAD>::N; ++k) acc[k] = 0.0f;
}

// ---- packed panel records ---------------------------------------------------------------------------
//   r[0] = { x0 y0 z0 x1 }  r[1] = { y1 z1 x2 y2 }  r[2] = { z2 wx wy wz }  r[3] = { q cx cy cz }
//   r[4] = { thr0 = 4 sqrt(area), area, sq_threshold(thr0), 0 }
// Padding records (j >= np) sit far away with zero strength and thr0 = 0: always one zero-valued leaf.
__global__ void pan_pack_kernel(int64_t np, int64_t np_pad, const float* nx, const float* ny, const float* nz,
                                const uint32_t* idx, const float* tsx, const float* tsy, const float* tsz,
                                const float* area, const float* sss, float4* out) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= np_pad) return;
  float4 r0 = make_float4(1e18f, 1e18f, 1e18f, 1e18f), r1 = r0, r2 = make_float4(1e18f, 0.f, 0.f, 0.f);
  float4 r3 = make_float4(0.f, 1e18f, 1e18f, 1e18f), r4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (j < np) {
    const uint32_t a = idx[3 * j], b = idx[3 * j + 1], c = idx[3 * j + 2];
    const float x0 = nx[a], y0 = ny[a], z0 = nz[a], x1 = nx[b], y1 = ny[b], z1 = nz[b], x2 = nx[c], y2 = ny[c], z2 = nz[c];
    const float sa = area[j];
    r0 = make_float4(x0, y0, z0, x1);
    r1 = make_float4(y1, z1, x2, y2);
    r2 = make_float4(z2, tsx ? tsx[j] : 0.f, tsy ? tsy[j] : 0.f, tsz ? tsz[j] : 0.f);
    r3 = make_float4(sss ? __fmul_rn(sss[j], sa) : 0.f, third_sum(x0, x1, x2), third_sum(y0, y1, y2), third_sum(z0, z1, z2));
    const float thr0 = __fmul_rn(__fsqrt_rn(sa), 4.0f);
    r4 = make_float4(thr0, sa, sq_threshold(thr0), 0.f);
  }
  float4* o = out + (size_t)j * kPanRec;
  o[0] = r0; o[1] = r1; o[2] = r2; o[3] = r3; o[4] = r4;
}

__host__ __device__ inline int64_t padded_panels(int64_t np) { return ((np + kPanTile - 1) / kPanTile) * kPanTile; }

struct PanPtsArgs {
  const float4* pan;     // packed panel records, padded to whole tiles
  int ntiles;            // tiles in the stream
  int nsplit;            // gridDim.y slices of the tile range
  int64_t nt;
  const float* tx; const float* ty; const float* tz;
  float* tu; float* tv; float* tw;
  float* tug; int64_t tug_stride;
  double* partial;       // nsplit > 1: [nsplit][12|3][nt], one slab per panel-tile slice
  unsigned long long* counts;  // [0] leaves, [1] splits (may be nullptr)
};

__device__ __forceinline__ void add_counts(unsigned long long* g, unsigned (&counts)[2]) {
  if (!g) return;
  const unsigned l = __reduce_add_sync(0xffffffffu, counts[0]);
  const unsigned s = __reduce_add_sync(0xffffffffu, counts[1]);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(g, (unsigned long long)l);
    atomicAdd(g + 1, (unsigned long long)s);
  }
}

// panels -> points: one target point per thread, panel tiles through shared memory.
template <bool GRAD, int BLOCK>
__global__ void __launch_bounds__(BLOCK) pan_pts_kernel(const PanPtsArgs p) {
  constexpr int NA = PanAcc<GRAD>::N;
  constexpr int NS = GRAD ? 12 : 3;
  // every WARP streams the panel tiles through its own 5 KB of shared memory and synchronises only with itself:
  // the work per tile varies from warp to warp (how many of its pairs subdivide), and a CTA-wide barrier per tile
  // made every warp wait for the slowest one (ncu: 4.2 barrier-stall cycles per issued instruction, profiles/
  // r01_pan_pts_ncu.txt). The second copy of a tile comes out of L2.
  __shared__ alignas(16) float4 tiles[BLOCK / 32][kPanTile * kPanRec];
  float4* tile = tiles[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;

  const int per = (p.ntiles + p.nsplit - 1) / p.nsplit;
  const int k0 = blockIdx.y * per;
  const int k1 = min(p.ntiles, k0 + per);

  const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
  const int64_t ic = min(i, p.nt - 1);
  const float tx = p.tx[ic], ty = p.ty[ic], tz = p.tz[ic];

  float acc[NA];
  double sum[NS];
  unsigned counts[2] = {0u, 0u};
#pragma unroll
  for (int k = 0; k < NA; ++k) acc[k] = 0.0f;
#pragma unroll
  for (int k = 0; k < NS; ++k) sum[k] = 0.0;

  for (int k = k0; k < k1; ++k) {
    __syncwarp();
    const float4* g = p.pan + (size_t)k * (kPanTile * kPanRec);
#pragma unroll
    for (int e = lane; e < kPanTile * kPanRec; e += 32) tile[e] = g[e];
    __syncwarp();
    // Two phases per tile, so that lanes whose pair needs the deep (divergent) subdivision run it TOGETHER instead of
    // one or two at a time while the rest of the warp waits: (A) every panel's level-0 test and, where it is well
    // separated, its single leaf - convergent, warp-broadcast reads; pairs that are not are remembered in a 64-bit
    // mask (kPanTile = 64); (B) each lane walks its own mask in ascending panel order. Per target this only reorders
    // the FP32 terms inside one tile (far leaves first); leaf and split counts are unchanged.
    unsigned long long near = 0ull;
#pragma unroll 2
    for (int j = 0; j < kPanTile; ++j) {
      const float4 r2 = tile[j * kPanRec + 2], r3 = tile[j * kPanRec + 3], r4 = tile[j * kPanRec + 4];
      if (pan_node<GRAD>(r3.y, r3.z, r3.w, r4.z, false, tx, ty, tz, r2.y, r2.z, r2.w, r3.x, acc)) counts[0] += 1;
      else near |= 1ull << j;
    }
    while (near) {
      const int j = __ffsll((long long)near) - 1;
      near &= near - 1ull;
      const float4 r0 = tile[j * kPanRec], r1 = tile[j * kPanRec + 1], r2 = tile[j * kPanRec + 2],
                   r3 = tile[j * kPanRec + 3], r4 = tile[j * kPanRec + 4];
      const Tri t{r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x};
      pan_subdivide<GRAD>(t, r4.z, r2.y, r2.z, r2.w, r3.x, tx, ty, tz, acc, counts);
    }
    pan_promote<GRAD>(acc, sum);
  }

  // padding records and clamped duplicate threads are not part of the reference's count
  if (i >= p.nt) counts[0] = counts[1] = 0u;
  add_counts(p.counts, counts);
  if (i >= p.nt) return;
  if (p.nsplit > 1) {
    double* slab = p.partial + (size_t)blockIdx.y * NS * p.nt;
#pragma unroll
    for (int k = 0; k < NS; ++k) slab[(size_t)k * p.nt + i] = sum[k];
  } else {
    p.tu[i] = (float)((double)p.tu[i] + sum[0]);
    p.tv[i] = (float)((double)p.tv[i] + sum[1]);
    p.tw[i] = (float)((double)p.tw[i] + sum[2]);
    if constexpr (GRAD) {
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        float* g = p.tug + (size_t)k * p.tug_stride + i;
        *g = (float)((double)*g + sum[3 + k]);
      }
    }
  }
}

// panels -> points with a WARP-LEVEL WORK QUEUE for the pairs that subdivide (the product kernel; pan_pts_kernel above is kept
// as the measured baseline - o3d_cuda_set_panel_queue). Phase A is the same convergent pass over the tile. What it defers
// differs from lane to lane by an order of magnitude - a point next to the body is near dozens of panels, a point in the wake
// near none - and a lane that walks only its own list leaves the others idle (ncu on the baseline: 16 of 32 lanes active,
// profiles/r01_pan_pts_ncu_v9.txt). Here the warp pools its deferred (point, panel) items: every lane writes its items to a
// shared list (owner lane | panel, in owner order), the warp takes them 32 at a time - lane l walks item base + l for
// WHOSEVER point it belongs to - and the item's partial sums go back through shared memory, where every owner adds the
// partials of its own items in item order. Per target the FP32 terms of a tile are only regrouped (one partial sum per deferred
// panel instead of one running sum); leaf and split counts are those of the reference.
template <bool GRAD, int BLOCK>
__global__ void __launch_bounds__(BLOCK) pan_pts_queue_kernel(const PanPtsArgs p) {
  constexpr int NA = PanAcc<GRAD>::N;
  constexpr int NS = GRAD ? 12 : 3;
  constexpr int NW = BLOCK / 32;
  constexpr int RET = NA + 1;                                   // row stride of the return buffer: odd, conflict-free columns
  // per-warp tile of phase A, pair-interleaved so that one packed instruction carries two panels (built while staging):
  //   pair[4 m] = { -cx0 -cx1 -cy0 -cy1 }  [4 m + 1] = { -cz0 -cz1 T0 T1 }  [4 m + 2] = { wx0 wx1 wy0 wy1 }  [4 m + 3] = { wz0 wz1 q0 q1 }
  // for panels 2 m, 2 m + 1 of the tile (T = squared level-0 threshold). The vertices are not staged: only the few pairs that
  // subdivide need them, and read them from the record array (L1 / L2).
  __shared__ alignas(16) float4 pairs[NW][(kPanTile / 2) * 4];  // 2 KB per warp
  __shared__ unsigned short lists[NW][32 * kPanTile];           // per-warp item list: owner << 6 | panel (4 KB)
  __shared__ float rets[NW][32 * RET];                          // per-warp partial sums of the 32 items in flight
  static_assert(kPanTile == 64, "one lane stages one panel pair");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4* pair = pairs[warp];
  unsigned short* list = lists[warp];
  float* ret = rets[warp];

  const int per = (p.ntiles + p.nsplit - 1) / p.nsplit;
  const int k0 = blockIdx.y * per;
  const int k1 = min(p.ntiles, k0 + per);

  const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
  const int64_t ic = min(i, p.nt - 1);
  const float tx = p.tx[ic], ty = p.ty[ic], tz = p.tz[ic];
  const float2 tx2 = f2(tx, tx), ty2 = f2(ty, ty), tz2 = f2(tz, tz);
  const bool live = i < p.nt;                                   // clamped duplicate threads defer nothing and count nothing

  float2 acc2[NA];
  float acc[NA];
  double sum[NS];
  unsigned counts[2] = {0u, 0u};
#pragma unroll
  for (int k = 0; k < NA; ++k) { acc2[k] = f2(0.f, 0.f); acc[k] = 0.0f; }
#pragma unroll
  for (int k = 0; k < NS; ++k) sum[k] = 0.0;

  for (int k = k0; k < k1; ++k) {
    __syncwarp();
    const float4* g = p.pan + (size_t)k * (kPanTile * kPanRec);
    {
      const float4* ra = g + (size_t)(2 * lane) * kPanRec;      // this lane stages panels 2 lane, 2 lane + 1
      const float4* rb = ra + kPanRec;
      const float4 a2 = ra[2], a3 = ra[3], a4 = ra[4], b2 = rb[2], b3 = rb[3], b4 = rb[4];
      pair[4 * lane + 0] = make_float4(-a3.y, -b3.y, -a3.z, -b3.z);
      pair[4 * lane + 1] = make_float4(-a3.w, -b3.w, a4.z, b4.z);
      pair[4 * lane + 2] = make_float4(a2.y, b2.y, a2.z, b2.z);
      pair[4 * lane + 3] = make_float4(a2.w, b2.w, a3.x, b3.x);
    }
    __syncwarp();
    // (A) every panel's level-0 test and, where the pair is well separated, its single leaf - two panels per instruction.
    //     The decision is the reference's, bit for bit: t + (-c) is t - c, the squared distance is formed unfused in its order.
    unsigned long long near = 0ull;
    unsigned leaves0 = 0u;
#pragma unroll 2
    for (int m = 0; m < kPanTile / 2; ++m) {
      const float4 q0 = pair[4 * m], q1 = pair[4 * m + 1], q2 = pair[4 * m + 2], q3 = pair[4 * m + 3];
      const float2 dx = __fadd2_rn(tx2, f2(q0.x, q0.y)), dy = __fadd2_rn(ty2, f2(q0.z, q0.w)), dz = __fadd2_rn(tz2, f2(q1.x, q1.y));
      const float2 d2 = sumsq2_rn(dx, dy, dz);
      const bool far0 = d2.x > q1.z, far1 = d2.y > q1.w;
      leaves0 += (far0 ? 1u : 0u) + (far1 ? 1u : 0u);
      near |= (unsigned long long)((far0 ? 0u : 1u) | (far1 ? 0u : 2u)) << (2 * m);
      const float2 rs = f2(far0 ? rsqrt_approx(d2.x) : 0.0f, far1 ? rsqrt_approx(d2.y) : 0.0f);
      pan_leaf2<GRAD>(dx, dy, dz, rs, f2(q2.x, q2.y), f2(q2.z, q2.w), f2(q3.x, q3.y), f2(q3.z, q3.w), acc2);
    }
#pragma unroll
    for (int q = 0; q < NA; ++q) { acc[q] = acc2[q].x + acc2[q].y; acc2[q] = f2(0.f, 0.f); }
    if (live) counts[0] += leaves0;
    else near = 0ull;
    // (B) pool the deferred pairs of the warp
    const int mine = __popcll(near);
    int first = mine;                                            // exclusive prefix over the lanes
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, first, off);
      if (lane >= off) first += v;
    }
    const int total = __shfl_sync(0xffffffffu, first, 31);
    first -= mine;
    if (total == 0) {
      pan_promote<GRAD>(acc, sum);
      continue;
    }
    {
      unsigned long long m = near;
      int at = first;
      while (m) {
        const int j = __ffsll((long long)m) - 1;
        m &= m - 1ull;
        list[at++] = (unsigned short)((lane << 6) | j);
      }
    }
    __syncwarp();
    for (int base = 0; base < total; base += 32) {
      const int gi = base + lane;
      const bool have = gi < total;
      const unsigned e = have ? list[gi] : (unsigned)(lane << 6);
      const int owner = (int)(e >> 6), j = (int)(e & 63u);
      const float ox = __shfl_sync(0xffffffffu, tx, owner), oy = __shfl_sync(0xffffffffu, ty, owner), oz = __shfl_sync(0xffffffffu, tz, owner);
      float part[NA];
#pragma unroll
      for (int q = 0; q < NA; ++q) part[q] = 0.0f;
      if (have) {
        const float4* r = g + (size_t)j * kPanRec;
        const float4 r0 = r[0], r1 = r[1], r2 = r[2], r3 = r[3], r4 = r[4];
        const Tri t{r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x};
        pan_subdivide<GRAD>(t, r4.z, r2.y, r2.z, r2.w, r3.x, ox, oy, oz, part, counts);
      }
      __syncwarp();                                              // the previous batch's partials have been consumed
#pragma unroll
      for (int q = 0; q < NA; ++q) ret[lane * RET + q] = part[q];
      __syncwarp();
      // every owner adds the partials of its own items of this batch, in item order: slots [a, b) of the batch
      const int a = max(first, base) - base, b = min(first + mine, base + 32) - base;
      for (int sl = a; sl < b; ++sl) {
#pragma unroll
        for (int q = 0; q < NA; ++q) acc[q] += ret[sl * RET + q];
      }
    }
    pan_promote<GRAD>(acc, sum);
  }

  add_counts(p.counts, counts);
  if (!live) return;
  if (p.nsplit > 1) {
    double* slab = p.partial + (size_t)blockIdx.y * NS * p.nt;
#pragma unroll
    for (int k = 0; k < NS; ++k) slab[(size_t)k * p.nt + i] = sum[k];
  } else {
    p.tu[i] = (float)((double)p.tu[i] + sum[0]);
    p.tv[i] = (float)((double)p.tv[i] + sum[1]);
    p.tw[i] = (float)((double)p.tw[i] + sum[2]);
    if constexpr (GRAD) {
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        float* g2 = p.tug + (size_t)k * p.tug_stride + i;
        *g2 = (float)((double)*g2 + sum[3 + k]);
      }
    }
  }
}

// ---- particles -> panels (BEM right-hand side) ---------------------------------------------------------
// One target PANEL per thread (its triangle stays in registers); the particles stream through shared
// memory in the packed pair-interleaved layout of biot_pp.cuh (positions negated). The particle axis is
// split over gridDim.y; per-slice FP64 slabs are summed in order by pp_finish_kernel, which applies the `-=`.
struct PtsPanArgs {
  const float4* src;     // packed particle stream (pp_pack2_kernel)
  int64_t ns;            // real particle count (padding records are skipped)
  int ntiles, nsplit;
  int64_t np;            // target panels
  const float4* pan;     // packed panel records (strength fields unused)
  double* partial;       // [nsplit][3][np], one slab per particle slice
  unsigned long long* counts;
};

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) pts_pan_kernel(const PtsPanArgs p) {
  // warp-private particle tiles, synchronised per warp (see pan_pts_kernel)
  __shared__ alignas(128) float4 tiles[BLOCK / 32][kTile * 2];
  float4* tile = tiles[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const int per = (p.ntiles + p.nsplit - 1) / p.nsplit;
  const int k0 = blockIdx.y * per;
  const int k1 = min(p.ntiles, k0 + per);

  const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
  const int64_t ic = min(i, p.np - 1);
  const float4* r = p.pan + (size_t)ic * kPanRec;
  const float4 r0 = r[0], r1 = r[1], r2 = r[2], r3 = r[3], r4 = r[4];
  const Tri t{r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x};
  const float cx = r3.y, cy = r3.z, cz = r3.w, thr0 = r4.z;   // the squared level-0 threshold (sq_threshold)

  float acc[3] = {0.f, 0.f, 0.f};
  double sum[3] = {0.0, 0.0, 0.0};
  unsigned counts[2] = {0u, 0u};

  for (int k = k0; k < k1; ++k) {
    __syncwarp();
    const float4* g = p.src + (size_t)k * (kTile * 2);
#pragma unroll 4
    for (int e = lane; e < kTile * 2; e += 32) tile[e] = g[e];
    __syncwarp();
    const int cnt = (int)min((int64_t)kTile, p.ns - (int64_t)k * kTile);
    // same two phases as pan_pts_kernel, 64 particles at a time: (A) level-0 test + leaf where well separated,
    // (B) every lane subdivides against the particles it marked, together
#pragma unroll 1
    for (int c0 = 0; c0 < cnt; c0 += 64) {
      const int cn = min(64, cnt - c0);
      unsigned long long near = 0ull;
      // c0 is even and a record pair holds particles 2 pr, 2 pr + 1: one set of four LDS.128 serves both (the stream is
      // padded to whole tiles, so the second half of the last pair is readable; it is skipped when jj + 1 == cn)
      // two particles per packed instruction: the record pair holds (-x0 -x1 | -y0 -y1 | -z0 -z1), so nd = (-x) + c is -(x - c)
      // exactly, the squared distance - and with it the reference's decision - has the same bits, and the sign goes into rs
      // (the velocity sums are linear in r3 = rs^3)
      const float2 cx2 = f2(cx, cx), cy2 = f2(cy, cy), cz2 = f2(cz, cz);
      float2 acc2[3] = {f2(0.f, 0.f), f2(0.f, 0.f), f2(0.f, 0.f)};
      unsigned leaves0 = 0u;
#pragma unroll 2
      for (int jj = 0; jj < cn; jj += 2) {
        const int pr = (c0 + jj) >> 1;
        const float4 q0 = tile[4 * pr], q1 = tile[4 * pr + 1], q2 = tile[4 * pr + 2], q3 = tile[4 * pr + 3];
        const float2 dx = __fadd2_rn(f2(q0.x, q0.y), cx2), dy = __fadd2_rn(f2(q0.z, q0.w), cy2), dz = __fadd2_rn(f2(q1.x, q1.y), cz2);
        const float2 d2 = sumsq2_rn(dx, dy, dz);
        const bool two = jj + 1 < cn;                            // the stream is padded to whole tiles: an odd count ends on half a pair
        const bool far0 = d2.x > thr0, far1 = d2.y > thr0;
        leaves0 += (far0 ? 1u : 0u) + (two && far1 ? 1u : 0u);
        near |= (unsigned long long)((far0 ? 0u : 1u) | (two && !far1 ? 2u : 0u)) << jj;
        const float2 rs = f2(far0 ? -rsqrt_approx(d2.x) : 0.0f, two && far1 ? -rsqrt_approx(d2.y) : 0.0f);
        pan_leaf2<false, false>(dx, dy, dz, rs, f2(q2.x, q2.y), f2(q2.z, q2.w), f2(q3.x, q3.y), f2(0.f, 0.f), acc2);
      }
      counts[0] += leaves0;
      acc[0] += acc2[0].x + acc2[0].y; acc[1] += acc2[1].x + acc2[1].y; acc[2] += acc2[2].x + acc2[2].y;
      while (near) {
        const int jj = __ffsll((long long)near) - 1;
        near &= near - 1ull;
        const int j = c0 + jj, pr = j >> 1, h = j & 1;
        const float4 q0 = tile[4 * pr], q1 = tile[4 * pr + 1], q2 = tile[4 * pr + 2], q3 = tile[4 * pr + 3];
        const float px = -(h ? q0.y : q0.x), py = -(h ? q0.w : q0.z), pz = -(h ? q1.y : q1.x);
        const float wx = h ? q2.y : q2.x, wy = h ? q2.w : q2.z, wz = h ? q3.y : q3.x;
        pan_subdivide<false>(t, thr0, wx, wy, wz, 0.0f, px, py, pz, acc, counts);
      }
    }
    sum[0] += (double)acc[0]; sum[1] += (double)acc[1]; sum[2] += (double)acc[2];
    acc[0] = acc[1] = acc[2] = 0.f;
  }
  if (i >= p.np) counts[0] = counts[1] = 0u;
  add_counts(p.counts, counts);
  if (i >= p.np) return;
  double* slab = p.partial + (size_t)blockIdx.y * 3 * p.np;
#pragma unroll
  for (int k = 0; k < 3; ++k) slab[(size_t)k * p.np + i] = sum[k];
}

// particles -> panels with the warp-level work queue of pan_pts_queue_kernel, roles swapped: a lane owns a target PANEL, the
// deferred items are (panel, particle) pairs of one 64-particle chunk; whichever lane takes an item fetches the owner's
// triangle by shuffles, subdivides it against the particle (read from the warp's tile) and returns three partial sums.
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) pts_pan_queue_kernel(const PtsPanArgs p) {
  constexpr int NW = BLOCK / 32;
  constexpr int RET = 5;                                         // row stride of the return buffer (3 sums, odd padding)
  __shared__ alignas(128) float4 tiles[NW][kTile * 2];
  __shared__ unsigned short lists[NW][32 * 64];                  // owner << 6 | particle within the chunk
  __shared__ float rets[NW][32 * RET];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4* tile = tiles[warp];
  unsigned short* list = lists[warp];
  float* ret = rets[warp];
  const int per = (p.ntiles + p.nsplit - 1) / p.nsplit;
  const int k0 = blockIdx.y * per;
  const int k1 = min(p.ntiles, k0 + per);

  const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
  const int64_t ic = min(i, p.np - 1);
  const bool live = i < p.np;
  const float4* r = p.pan + (size_t)ic * kPanRec;
  const float4 r0 = r[0], r1 = r[1], r2 = r[2], r3 = r[3], r4 = r[4];
  const Tri t{r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x};
  const float cx = r3.y, cy = r3.z, cz = r3.w, thr0 = r4.z;   // the squared level-0 threshold (sq_threshold)

  float acc[3] = {0.f, 0.f, 0.f};
  double sum[3] = {0.0, 0.0, 0.0};
  unsigned counts[2] = {0u, 0u};

  for (int k = k0; k < k1; ++k) {
    __syncwarp();
    const float4* g = p.src + (size_t)k * (kTile * 2);
#pragma unroll 4
    for (int e = lane; e < kTile * 2; e += 32) tile[e] = g[e];
    __syncwarp();
    const int cnt = (int)min((int64_t)kTile, p.ns - (int64_t)k * kTile);
#pragma unroll 1
    for (int c0 = 0; c0 < cnt; c0 += 64) {
      const int cn = min(64, cnt - c0);
      unsigned long long near = 0ull;
      unsigned leaves0 = 0u;
#pragma unroll 1
      for (int jj = 0; jj < cn; jj += 2) {
        const int pr = (c0 + jj) >> 1;
        const float4 q0 = tile[4 * pr], q1 = tile[4 * pr + 1], q2 = tile[4 * pr + 2], q3 = tile[4 * pr + 3];
        if (pan_node<false>(cx, cy, cz, thr0, false, -q0.x, -q0.z, -q1.x, q2.x, q2.z, q3.x, 0.0f, acc)) leaves0 += 1;
        else near |= 1ull << jj;
        if (jj + 1 < cn) {
          if (pan_node<false>(cx, cy, cz, thr0, false, -q0.y, -q0.w, -q1.y, q2.y, q2.w, q3.y, 0.0f, acc)) leaves0 += 1;
          else near |= 1ull << (jj + 1);
        }
      }
      if (live) counts[0] += leaves0;
      else near = 0ull;
      const int mine = __popcll(near);
      int first = mine;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, first, off);
        if (lane >= off) first += v;
      }
      const int total = __shfl_sync(0xffffffffu, first, 31);
      first -= mine;
      if (total == 0) continue;
      {
        unsigned long long m = near;
        int at = first;
        while (m) {
          const int jj = __ffsll((long long)m) - 1;
          m &= m - 1ull;
          list[at++] = (unsigned short)((lane << 6) | jj);
        }
      }
      __syncwarp();
      for (int base = 0; base < total; base += 32) {
        const int gi = base + lane;
        const bool have = gi < total;
        const unsigned e = have ? list[gi] : (unsigned)(lane << 6);
        const int owner = (int)(e >> 6), jj = (int)(e & 63u);
        Tri o;
        o.x0 = __shfl_sync(0xffffffffu, t.x0, owner); o.y0 = __shfl_sync(0xffffffffu, t.y0, owner); o.z0 = __shfl_sync(0xffffffffu, t.z0, owner);
        o.x1 = __shfl_sync(0xffffffffu, t.x1, owner); o.y1 = __shfl_sync(0xffffffffu, t.y1, owner); o.z1 = __shfl_sync(0xffffffffu, t.z1, owner);
        o.x2 = __shfl_sync(0xffffffffu, t.x2, owner); o.y2 = __shfl_sync(0xffffffffu, t.y2, owner); o.z2 = __shfl_sync(0xffffffffu, t.z2, owner);
        const float othr = __shfl_sync(0xffffffffu, thr0, owner);
        float part[3] = {0.f, 0.f, 0.f};
        if (have) {
          const int j = c0 + jj, pr = j >> 1, h = j & 1;
          const float4 q0 = tile[4 * pr], q1 = tile[4 * pr + 1], q2 = tile[4 * pr + 2], q3 = tile[4 * pr + 3];
          const float px = -(h ? q0.y : q0.x), py = -(h ? q0.w : q0.z), pz = -(h ? q1.y : q1.x);
          const float wx = h ? q2.y : q2.x, wy = h ? q2.w : q2.z, wz = h ? q3.y : q3.x;
          pan_subdivide<false>(o, othr, wx, wy, wz, 0.0f, px, py, pz, part, counts);
        }
        __syncwarp();
        ret[lane * RET + 0] = part[0]; ret[lane * RET + 1] = part[1]; ret[lane * RET + 2] = part[2];
        __syncwarp();
        const int a = max(first, base) - base, b = min(first + mine, base + 32) - base;
        for (int sl = a; sl < b; ++sl) {
          acc[0] += ret[sl * RET + 0]; acc[1] += ret[sl * RET + 1]; acc[2] += ret[sl * RET + 2];
        }
      }
    }
    sum[0] += (double)acc[0]; sum[1] += (double)acc[1]; sum[2] += (double)acc[2];
    acc[0] = acc[1] = acc[2] = 0.f;
  }
  add_counts(p.counts, counts);
  if (!live) return;
  double* slab = p.partial + (size_t)blockIdx.y * 3 * p.np;
#pragma unroll
  for (int k = 0; k < 3; ++k) slab[(size_t)k * p.np + i] = sum[k];
}

// ---- panel -> panel BEM coefficient block -------------------------------------------------------------
// rkernel_2vs_2p (src/Kernels.h:1217-1315): both triangles subdivide, 16 children per level, the
// strength falls by 1/16, size = sqrt(sa) + sqrt(ta). The reference runs it three times per pair with
// unit sheet strength along the source's x1, x2 and a unit source sheet (src/Coefficients.h:356-405);
// the traversal does not depend on the strength, so one traversal carries all three: per leaf
//   R1 += s r3 (d x b1),  R2 += s r3 (d x b2),  R3 += s r3 d      (s = sa 16^-l, exact)
// accumulated in FP32 in the reference's leaf order (it accumulates this block in S = float).
struct PanCoefArgs {
  const float4* spn;  // packed source panels
  const float4* tpn;  // packed target panels
  int64_t nsp, ntp;
  int64_t j0, j1;     // source-panel (column) range of this launch
  const float* sb1; const float* sb2;              // source bases, SoA x|y|z with stride nsp
  const float* tb1; const float* tb2; const float* tnrm;  // target bases, stride ntp
  int self;
  float* coeffs;      // column-major (3 ntp) x (3 nsp); this launch writes columns 3 j0 .. 3 j1
  int64_t col_offset; // column index of j0 inside `coeffs` (0 when the buffer holds only this range)
  unsigned long long* counts;  // [0] leaves, [1] splits
};

struct Coef9 {
  float v[9];
};

__device__ __forceinline__ void coef_leaf(float dx, float dy, float dz, float distsq, float s, const float (&b1)[3],
                                          const float (&b2)[3], Coef9& R) {
  const float rs = rsqrt_approx(distsq);
  const float k = s * (rs * rs * rs);
  R.v[0] = fmaf(k, fmaf(dz, b1[1], -(dy * b1[2])), R.v[0]);
  R.v[1] = fmaf(k, fmaf(dx, b1[2], -(dz * b1[0])), R.v[1]);
  R.v[2] = fmaf(k, fmaf(dy, b1[0], -(dx * b1[1])), R.v[2]);
  R.v[3] = fmaf(k, fmaf(dz, b2[1], -(dy * b2[2])), R.v[3]);
  R.v[4] = fmaf(k, fmaf(dx, b2[2], -(dz * b2[0])), R.v[4]);
  R.v[5] = fmaf(k, fmaf(dy, b2[0], -(dx * b2[1])), R.v[5]);
  R.v[6] = fmaf(k, dx, R.v[6]);
  R.v[7] = fmaf(k, dy, R.v[7]);
  R.v[8] = fmaf(k, dz, R.v[8]);
}

// SQ: thr is the squared threshold of sq_threshold (levels 1..3); !SQ: thr is 4 (sqrt(sa) + sqrt(ta)) itself and the
// distance takes the reference's correctly rounded square root (level 0: most pairs end there, and finding the squared
// threshold costs more than the one square root it would save).
template <bool SQ>
__device__ __forceinline__ bool coef_node(const Tri& s, const Tri& t, float thr, bool deepest, float str,
                                          const float (&b1)[3], const float (&b2)[3], Coef9& R) {
  const float sx = third_sum(s.x0, s.x1, s.x2), sy = third_sum(s.y0, s.y1, s.y2), sz = third_sum(s.z0, s.z1, s.z2);
  const float tx = third_sum(t.x0, t.x1, t.x2), ty = third_sum(t.y0, t.y1, t.y2), tz = third_sum(t.z0, t.z1, t.z2);
  const float dx = __fsub_rn(tx, sx), dy = __fsub_rn(ty, sy), dz = __fsub_rn(tz, sz);
  const float distsq = sumsq_rn(dx, dy, dz);
  if ((SQ ? distsq : __fsqrt_rn(distsq)) > thr || deepest) {
    coef_leaf(dx, dy, dz, distsq, str, b1, b2, R);
    return true;
  }
  return false;
}

// The 3 x 3 influence block of source panel j on target panel i: rows = target (t1, t2, n) components, columns = the
// three unknowns of the source panel (vortex x1, vortex x2, source), already scaled by 1/4pi, with the self-block
// override applied. m[k][r] = A[3i + r, 3j + k]. One traversal carries all three unit strengths.
__device__ __forceinline__ void coef_block(const PanCoefArgs& p, const int64_t i, const int64_t j, unsigned (&counts)[2],
                                           float (&m)[3][3]) {
  const float4* sr = p.spn + (size_t)j * kPanRec;
  const float4 a0 = sr[0], a1 = sr[1], a2 = sr[2], a4 = sr[4];
  const Tri s0{a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x};
  const float sa = a4.y;
  const float4* tr = p.tpn + (size_t)i * kPanRec;
  const float4 c0 = tr[0], c1 = tr[1], c2 = tr[2], c4 = tr[4];
  const Tri t0{c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w, c2.x};
  const float b1[3] = {p.sb1[j], p.sb1[p.nsp + j], p.sb1[2 * p.nsp + j]};
  const float b2[3] = {p.sb2[j], p.sb2[p.nsp + j], p.sb2[2 * p.nsp + j]};
  // trisize = sqrt(sa) + sqrt(ta); threshold 4 * trisize, halves exactly per level
  const float thr0 = __fmul_rn(__fadd_rn(__fsqrt_rn(sa), __fsqrt_rn(c4.y)), 4.0f);

  Coef9 R;
#pragma unroll
  for (int k = 0; k < 9; ++k) R.v[k] = 0.0f;

  if (coef_node<false>(s0, t0, thr0, false, sa, b1, b2, R)) {
    counts[0] += 1;
  } else {
    counts[1] += 1;
    const Mids sm0 = tri_mids(s0), tm0 = tri_mids(t0);
    const float str1 = sa * 0.0625f, thr1 = sq_threshold(thr0) * 0.25f;   // squared thresholds from here down
#pragma unroll 1
    for (int e1 = 0; e1 < 16; ++e1) {
      const Tri s1 = tri_child(s0, sm0, e1 >> 2), t1 = tri_child(t0, tm0, e1 & 3);
      if (coef_node<true>(s1, t1, thr1, false, str1, b1, b2, R)) { counts[0] += 1; continue; }
      counts[1] += 1;
      const Mids sm1 = tri_mids(s1), tm1 = tri_mids(t1);
      const float str2 = str1 * 0.0625f, thr2 = thr1 * 0.25f;
#pragma unroll 1   // (unroll 4 here: 254 registers, coefficient block 5.2 -> 7.7 ms at 5120 panels: measured, not adopted)
      for (int e2 = 0; e2 < 16; ++e2) {
        const Tri s2 = tri_child(s1, sm1, e2 >> 2), t2 = tri_child(t1, tm1, e2 & 3);
        if (coef_node<true>(s2, t2, thr2, false, str2, b1, b2, R)) { counts[0] += 1; continue; }
        counts[1] += 1;
        const Mids sm2 = tri_mids(s2), tm2 = tri_mids(t2);
        const float str3 = str2 * 0.0625f;
#pragma unroll     // both children (e3 >> 2, e3 & 3) become static choices of registers
        for (int e3 = 0; e3 < 16; ++e3) {
          const Tri s3 = tri_child(s2, sm2, e3 >> 2), t3 = tri_child(t2, tm2, e3 & 3);
          coef_node<true>(s3, t3, 0.0f, true, str3, b1, b2, R);
          counts[0] += 1;
        }
      }
    }
  }

  const float t1x = p.tb1[i], t1y = p.tb1[p.ntp + i], t1z = p.tb1[2 * p.ntp + i];
  const float t2x = p.tb2[i], t2y = p.tb2[p.ntp + i], t2z = p.tb2[2 * p.ntp + i];
  const float tnx = p.tnrm[i], tny = p.tnrm[p.ntp + i], tnz = p.tnrm[2 * p.ntp + i];
  const float fac = (float)(1.0 / (4.0 * 3.14159265358979323846));  // src/Coefficients.h:448
  const float twopi = (float)(2.0 * 3.14159265358979323846);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float ru = R.v[3 * k], rv = R.v[3 * k + 1], rw = R.v[3 * k + 2];
    float m0 = __fadd_rn(__fadd_rn(__fmul_rn(ru, t1x), __fmul_rn(rv, t1y)), __fmul_rn(rw, t1z));
    float m1 = __fadd_rn(__fadd_rn(__fmul_rn(ru, t2x), __fmul_rn(rv, t2y)), __fmul_rn(rw, t2z));
    float m2 = __fadd_rn(__fadd_rn(__fmul_rn(ru, tnx), __fmul_rn(rv, tny)), __fmul_rn(rw, tnz));
    if (p.self && i == j) {  // src/Coefficients.h:414-436
      m0 = k == 1 ? -twopi : 0.0f;
      m1 = k == 0 ? twopi : 0.0f;
      m2 = k == 2 ? twopi : 0.0f;
    }
    m[k][0] = m0 * fac; m[k][1] = m1 * fac; m[k][2] = m2 * fac;
  }
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) pan_coef_kernel(const PanCoefArgs p) {
  const int64_t j = p.j0 + blockIdx.x;                              // source panel = column block
  const int64_t i = (int64_t)blockIdx.y * BLOCK + threadIdx.x;      // target panel = row block
  const int64_t ic = min(i, p.ntp - 1);
  unsigned counts[2] = {0u, 0u};
  float m[3][3];
  coef_block(p, ic, j, counts, m);
  if (i >= p.ntp) counts[0] = counts[1] = 0u;
  add_counts(p.counts, counts);
  if (i >= p.ntp) return;
  const size_t nrows = (size_t)3 * p.ntp;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float* col = p.coeffs + ((size_t)(p.col_offset + (j - p.j0)) * 3 + k) * nrows + (size_t)3 * i;
    col[0] = m[k][0]; col[1] = m[k][1]; col[2] = m[k][2];
  }
}

// Matrix-free product y = A x with the same blocks: the (3 ntp) x (3 nsp) influence matrix of
// panels_on_panels_coeff (src/Coefficients.h:169-483) applied to a vector of panel unknowns without ever being stored -
// what BEM<S,I>::solve's GMRES needs from A (src/BEM.h:182-202) once (3 np)^2 floats no longer fit (the reference caps
// the panel count for that reason, src/Simulation.cpp:675). One thread owns one target panel (3 rows) and walks a slice
// [j0, j1) of the source panels in index order, FP64 row sums; blockIdx.y slices meet in slabs added in slice order by
// pan_matvec_finish_kernel - deterministic, no atomics.
struct PanMatvecArgs {
  PanCoefArgs c;         // geometry (coeffs / col_offset unused)
  const float* x;        // 3 nsp unknowns
  int64_t i0, ni;        // this launch's rows: target panels [i0, i0 + ni)
  double* partial;       // [nsplit][3][ni]
  int nsplit;
};

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) pan_matvec_kernel(const PanMatvecArgs p) {
  const int64_t il = (int64_t)blockIdx.x * BLOCK + threadIdx.x;   // row block local to this launch
  const int64_t ic = p.i0 + min(il, p.ni - 1);
  const int64_t per = (p.c.nsp + p.nsplit - 1) / p.nsplit;
  const int64_t j0 = (int64_t)blockIdx.y * per, j1 = min(p.c.nsp, j0 + per);
  unsigned counts[2] = {0u, 0u};
  double y0 = 0.0, y1 = 0.0, y2 = 0.0;
  for (int64_t j = j0; j < j1; ++j) {
    float m[3][3];
    coef_block(p.c, ic, j, counts, m);
    const float x0 = p.x[3 * j], x1 = p.x[3 * j + 1], x2 = p.x[3 * j + 2];
    y0 += (double)m[0][0] * x0 + (double)m[1][0] * x1 + (double)m[2][0] * x2;
    y1 += (double)m[0][1] * x0 + (double)m[1][1] * x1 + (double)m[2][1] * x2;
    y2 += (double)m[0][2] * x0 + (double)m[1][2] * x1 + (double)m[2][2] * x2;
  }
  if (il >= p.ni) counts[0] = counts[1] = 0u;
  add_counts(p.c.counts, counts);
  if (il >= p.ni) return;
  double* slab = p.partial + (size_t)blockIdx.y * 3 * p.ni;
  slab[il] = y0; slab[p.ni + il] = y1; slab[2 * p.ni + il] = y2;
}

__global__ void pan_matvec_finish_kernel(int nsplit, int64_t ntp, const double* partial, float* y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ntp) return;
  for (int r = 0; r < 3; ++r) {
    double acc = 0.0;
    for (int s = 0; s < nsplit; ++s) acc += partial[((size_t)s * 3 + r) * ntp + i];
    y[3 * i + r] = (float)acc;
  }
}

}  // namespace o3d

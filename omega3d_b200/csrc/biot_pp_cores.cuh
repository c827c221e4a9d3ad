// biot_pp_cores.cuh - particles -> points with the ALTERNATE core functions of the reference's src/CoreFunc.h.
//
// The reference picks its core function per build by moving one "#define USE_*_KERNEL" (src/CoreFunc.h:35-38);
// Winckelmans-Leonard is the shipped one and has its own searched, SASS-post-processed kernel (biot_pp.cuh:
// pp2_kernel). The other three - Rosenhead-Moore (:43-83), exponential (:86-238), Vatistas n=2 (:292-341) - run
// through ppc_kernel below: the same tile ring (cp.async.bulk + mbarrier), the same packed record stream, the same
// register blocking, FP32 tile sums promoted to FP64 once per tile and the same read-modify-write epilogue; only the
// radial factors (r3, bbb) differ. Everything after them - c = w x d, u += r3 c, G += d (x) (bbb c), the
// antisymmetric A = sum w r3 and the trace-free ninth slot - is core-independent (src/Kernels.h:155-193).
//
// The radius lane of a packed record holds what the selected core adds per SOURCE (pp_pack2_kernel, formed exactly
// as the reference forms it) and each thread keeps the matching per-TARGET term:
//     Rosenhead-Moore   r2 = |d|^2 + sr*sr + tr*tr            lane sr*sr          target tr*tr
//     exponential       corefac = 1/(sr*sr*sr + tr*tr*tr)     lane sr*sr*sr       target tr*tr*tr
//     Vatistas n=2      denom = |d|^4 + s2*s2 + t2*t2         lane (sr*sr)^2      target (tr*tr)^2
// Singular targets (core_func(distsq, sr)) are the same formulas with a zero target term.
#pragma once
#include "biot_pp.cuh"

#ifndef O3D_PPC_NOBAR
#define O3D_PPC_NOBAR 0   // as O3D_PP_NOBAR (biot_pp.cuh) for the alternate-core kernels: 1 = the velocity+gradient ones run without the per-tile barrier
#endif

namespace o3d {

constexpr int kCoreWL = 0, kCoreRM = 1, kCoreEXP = 2, kCoreV2 = 3;   // == O3D_CORE_* (include/o3d_cuda.h)

__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// what one particle's core radius contributes to the pair term, in the reference's operation order
__host__ __device__ __forceinline__ float core_radius_term(const int core, const float r) {
  const float r2 = r * r;
  return core == kCoreEXP ? r2 * r : core == kCoreV2 ? r2 * r2 : r2;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Exponential core, two lanes (src/CoreFunc.h:114-128 exp_cond, :158-172 exp_bbb, :213-238):
//   dist = sqrt(|d|^2), d3 = |d|^2 dist, corefac = 1/(sr^3 + tr^3), reld3 = d3 corefac, ood3 = 1/d3
//   r3  = reld3 > 16 ? ood3 : reld3 < 0.001 ? corefac : ood3 (1 - exp(-reld3))
//   bbb = reld3 > 16 ? -3 r3/|d|^2 : reld3 < 0.001 ? -1.5 dist r3 r3 : 3 (corefac exp(-reld3) - r3)/|d|^2
// All three arms are evaluated with packed arithmetic in the reference's operation order and the reference's comparisons
// pick one per lane (FSEL). At |d| = 0 the unselected arms hold inf / NaN exactly as the reference's scalar code would
// have produced had it evaluated them; selects do not propagate them (reld3 = 0 takes the "< 0.001" arm: r3 = corefac,
// bbb = -1.5 * 0 * r3 * r3 = -0). Four MUFU per lane: SQRT, two RCP, EX2 (exp(-x) = 2^(-x log2 e)).
// UNI: one core radius in the system - st is a kernel-wide constant and arrives as its reciprocal cf already (one MUFU.RCP per
// lane less, and no add).
template <bool GRAD, bool UNI>
__device__ __forceinline__ void exp_core2(const float2 dsq, const float2 st, float2& r3, float2& bbb) {
  const float2 dist = f2(sqrt_approx(dsq.x), sqrt_approx(dsq.y));
  const float2 d3 = __fmul2_rn(dsq, dist);
  const float2 cf = UNI ? st : f2(rcp_approx(st.x), rcp_approx(st.y));
  const float2 reld3 = __fmul2_rn(d3, cf);
  const float2 ood3 = f2(rcp_approx(d3.x), rcp_approx(d3.y));
  const float2 xe = __fmul2_rn(reld3, f2(-1.4426950408889634f, -1.4426950408889634f));
  const float2 e = f2(ex2_approx(xe.x), ex2_approx(xe.y));
  const float2 mid = __fmul2_rn(ood3, __fadd2_rn(f2(1.0f, 1.0f), neg2(e)));
  const bool far0 = reld3.x > 16.0f, near0 = reld3.x < 0.001f;
  const bool far1 = reld3.y > 16.0f, near1 = reld3.y < 0.001f;
  r3.x = far0 ? ood3.x : near0 ? cf.x : mid.x;
  r3.y = far1 ? ood3.y : near1 ? cf.y : mid.y;
  if constexpr (GRAD) {
    const float2 oodsq = __fmul2_rn(ood3, dist);                                          // 1 / |d|^2
    const float2 bfar = __fmul2_rn(__fmul2_rn(f2(-3.0f, -3.0f), r3), oodsq);
    const float2 bnear = __fmul2_rn(__fmul2_rn(f2(-1.5f, -1.5f), dist), __fmul2_rn(r3, r3));
    const float2 bmid = __fmul2_rn(__fmul2_rn(f2(3.0f, 3.0f), __ffma2_rn(cf, e, neg2(r3))), oodsq);
    bbb.x = far0 ? bfar.x : near0 ? bnear.x : bmid.x;
    bbb.y = far1 ? bfar.y : near1 ? bnear.y : bmid.y;
  }
}

// Two sources (one packed record pair) on one target; tt = the target's term of the table above. The Rosenhead-Moore and
// Vatistas bodies live in their own files: like pp_interact2's, their statement ORDER decides how many packed instructions
// pay a third register-file cycle, and tools/tune_order.py (O3D_TUNE_VARIANT=rm|rmvel|v2|v2vel) searches it.
//   Rosenhead-Moore:  st = tt + lane; d2 = |d|^2 + st; rs = rsqrt(d2); r3 = rs^3; bbb = (-3 rs^2) r3
//   Vatistas n=2:     st = tt + lane; den = |d|^4 + st; rq = rsqrt(den); r3 = rq sqrt(rq); bbb = (-3 rq) r3
// followed by c = (dz wy - dy wz, dx wz - dz wx, dy wx - dx wy), A += r3 w, u += r3 c, G += d (x) (bbb c) without the
// wz slot (trace-free: recovered as -(ux + vy) per tile, d . (d x w) = 0 whatever the core).
#ifndef O3D_PPC_BODY_RM_GRAD
#define O3D_PPC_BODY_RM_GRAD "ppc_body_rm_velgrad.inc"
#endif
#ifndef O3D_PPC_BODY_RM_VEL
#define O3D_PPC_BODY_RM_VEL "ppc_body_rm_vel.inc"
#endif
#ifndef O3D_PPC_BODY_V2_GRAD
#define O3D_PPC_BODY_V2_GRAD "ppc_body_v2_velgrad.inc"
#endif
#ifndef O3D_PPC_BODY_V2_VEL
#define O3D_PPC_BODY_V2_VEL "ppc_body_v2_vel.inc"
#endif
// UNI (every source and every target radius equal, found by pp_scan_kernel as in pp2_kernel): tt already holds the pair term
// st = lane + tt (exponential core: its reciprocal), the same bits every pair would have formed.
template <int CORE, bool GRAD, bool UNI>
__device__ __forceinline__ void ppc_interact2(const float4 q0, const float4 q1, const float4 q2, const float4 q3,
                                              const float2 tx, const float2 ty, const float2 tz, const float2 tt,
                                              float2 (&acc)[PPAcc<GRAD>::N]) {
  if constexpr (CORE == kCoreRM && GRAD) {
#include O3D_PPC_BODY_RM_GRAD
    return;
  }
  if constexpr (CORE == kCoreRM && !GRAD) {
#include O3D_PPC_BODY_RM_VEL
    return;
  }
  if constexpr (CORE == kCoreV2 && GRAD) {
#include O3D_PPC_BODY_V2_GRAD
    return;
  }
  if constexpr (CORE == kCoreV2 && !GRAD) {
#include O3D_PPC_BODY_V2_VEL
    return;
  }
  // exponential core: all three arms packed, one select per lane and factor
  const float2 dx = __fadd2_rn(tx, f2(q0.x, q0.y));
  const float2 dy = __fadd2_rn(ty, f2(q0.z, q0.w));
  const float2 dz = __fadd2_rn(tz, f2(q1.x, q1.y));
  const float2 st = UNI ? tt : __fadd2_rn(tt, f2(q1.z, q1.w));
  const float2 wx = f2(q2.x, q2.y), wy = f2(q2.z, q2.w), wz = f2(q3.x, q3.y);
  float2 r3, bbb = f2(0.f, 0.f);
  const float2 dsq = __ffma2_rn(dx, dx, __ffma2_rn(dy, dy, __fmul2_rn(dz, dz)));
  exp_core2<GRAD, UNI>(dsq, st, r3, bbb);
  const float2 t1 = __fmul2_rn(dy, wz);
  const float2 t2 = __fmul2_rn(dx, wz);
  const float2 t3 = __fmul2_rn(dx, wy);
  float2 cx = __ffma2_rn(dz, wy, neg2(t1));
  float2 cy = __ffma2_rn(neg2(dz), wx, t2);
  float2 cz = __ffma2_rn(dy, wx, neg2(t3));
  if constexpr (GRAD) {
    acc[12] = __ffma2_rn(r3, wx, acc[12]);
    acc[13] = __ffma2_rn(r3, wy, acc[13]);
    acc[14] = __ffma2_rn(r3, wz, acc[14]);
  }
  acc[0] = __ffma2_rn(r3, cx, acc[0]);
  acc[1] = __ffma2_rn(r3, cy, acc[1]);
  acc[2] = __ffma2_rn(r3, cz, acc[2]);
  if constexpr (GRAD) {
    cz = __fmul2_rn(bbb, cz);
    cy = __fmul2_rn(bbb, cy);
    cx = __fmul2_rn(bbb, cx);
    acc[3]  = __ffma2_rn(dx, cx, acc[3]);
    acc[4]  = __ffma2_rn(dx, cy, acc[4]);
    acc[5]  = __ffma2_rn(dx, cz, acc[5]);
    acc[8]  = __ffma2_rn(dy, cz, acc[8]);
    acc[7]  = __ffma2_rn(dy, cy, acc[7]);
    acc[6]  = __ffma2_rn(dy, cx, acc[6]);
    acc[9]  = __ffma2_rn(dz, cx, acc[9]);
    acc[10] = __ffma2_rn(dz, cy, acc[10]);
  }
}

// One tile of a persistent CTA's walk out of ring buffer BUF: pp2_tile (biot_pp.cuh) with this file's interaction.
template <int BUF, int CORE, int T, bool GRAD, bool UNI, int BLOCK>
__device__ __forceinline__ void ppc_tile(const PPArgs& p, PPWalk& w, PPBlock& s, float4 (&tile)[2][kTile * 2], PPSync& sy, const float ttu,
                                         float2 (&tx)[T], float2 (&ty)[T], float2 (&tz)[T], float2 (&tt)[T],
                                         float2 (&acc)[T][PPAcc<GRAD>::N], double (&sum)[T][GRAD ? 12 : 3]) {
  constexpr int NS = GRAD ? 12 : 3;
  constexpr int NA = PPAcc<GRAD>::N;
  if (s.fresh) {
    const int64_t base = (int64_t)s.b * (BLOCK * T) + threadIdx.x;
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int64_t i = min(base + (int64_t)t * BLOCK, p.nt - 1);
      tx[t] = f2(p.tx[i], p.tx[i]); ty[t] = f2(p.ty[i], p.ty[i]); tz[t] = f2(p.tz[i], p.tz[i]);
      const float term = UNI ? ttu : core_radius_term(CORE, p.tr ? p.tr[i] : 0.0f);
      tt[t] = f2(term, term);
#pragma unroll
      for (int k = 0; k < NS; ++k) sum[t][k] = 0.0;
    }
  }
  pp_ring_wait(sy, w.kring);
  const float4* __restrict__ src = tile[BUF];
#pragma unroll(CORE == kCoreEXP ? 2 : GRAD ? kPPUnrollGrad : kPPUnrollVel)
  for (int j = 0; j < kTile / 2; ++j) {
    const float4 q0 = src[4 * j], q1 = src[4 * j + 1], q2 = src[4 * j + 2], q3 = src[4 * j + 3];
#pragma unroll
    for (int t = 0; t < T; ++t) ppc_interact2<CORE, GRAD, UNI>(q0, q1, q2, q3, tx[t], ty[t], tz[t], tt[t], acc[t]);
  }
#pragma unroll
  for (int t = 0; t < T; ++t) {
    float h[NA];
#pragma unroll
    for (int q = 0; q < NA; ++q) { h[q] = acc[t][q].x + acc[t][q].y; acc[t][q] = f2(0.f, 0.f); }
    if constexpr (GRAD) h[11] = -(h[3] + h[7]);
    pp_promote<GRAD>(h, sum[t]);
  }
  constexpr bool NOBAR = (O3D_PPC_NOBAR == 1 && GRAD) || O3D_PPC_NOBAR == 2 || O3D_PP_NOBAR == 2;
  pp_ring_refill<BLOCK, NOBAR>(p, w, tile[BUF], &sy.full[BUF], &sy.released[BUF]);         // ++w.kring
  pp_segment_end<T, GRAD, BLOCK>(p, w, s, sum);
}

template <int CORE, int T, bool GRAD, bool UNI, int BLOCK>
__device__ __forceinline__ void ppc_walk(const PPArgs& p, PPWalk& w, float4 (&tile)[2][kTile * 2], PPSync& sy, const float ttu_in) {
  constexpr int NS = GRAD ? 12 : 3;
  constexpr int NA = PPAcc<GRAD>::N;
  // (the constant passes through a warp reduction so that ptxas holds it in a uniform register, as in pp2_walk)
  const float ttu = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(ttu_in)));
  float2 tx[T], ty[T], tz[T], tt[T];
  double sum[T][NS];
  float2 acc[T][NA];
#pragma unroll
  for (int t = 0; t < T; ++t) {
#pragma unroll
    for (int k = 0; k < NA; ++k) acc[t][k] = f2(0.f, 0.f);
  }
  PPBlock s = pp_first_block(w);
  while (w.kring < w.nk) {
    ppc_tile<0, CORE, T, GRAD, UNI, BLOCK>(p, w, s, tile, sy, ttu, tx, ty, tz, tt, acc, sum);
    if (w.kring < w.nk) ppc_tile<1, CORE, T, GRAD, UNI, BLOCK>(p, w, s, tile, sy, ttu, tx, ty, tz, tt, acc, sum);
  }
}

template <int CORE, int T, bool GRAD, int BLOCK>
__global__ void __launch_bounds__(BLOCK, kPPWarpsPerSM * 32 / BLOCK) ppc_kernel(const PPArgs p) {
  __shared__ alignas(128) float4 tile[2][kTile * 2];
  __shared__ alignas(8) PPSync sy;

  // persistent CTA: one loop over the tiles of its share, two per trip (ring buffer 0, 1), the target block changing at
  // segment boundaries (pp2_walk, biot_pp.cuh)
  PPWalk w = pp_ring_start<BLOCK>(p, tile, sy);
  // radius scan (pp_scan_kernel over this core's radius lane): uniform <=> one lane value among the sources that carry
  // strength and one target radius. The pair term st = lane + core_radius_term(tr) is then a constant of the launch.
  bool uni = false;
  float ttu = 0.0f;
  if (p.radius_range) {
    const uint32_t s0 = ~p.radius_range[0], s1 = p.radius_range[1], t0 = ~p.radius_range[2], t1 = p.radius_range[3];
    const float st = __fadd_rn(core_radius_term(CORE, p.tr ? __uint_as_float(t0) : 0.0f), __uint_as_float(s0));   // tt + lane, as every pair forms it
    uni = s0 == s1 && (!p.tr || t0 == t1) && s0 != 0xffffffffu && st > 0.0f;
    ttu = CORE == kCoreEXP ? rcp_approx(st) : st;
  }
  if (uni) ppc_walk<CORE, T, GRAD, true, BLOCK>(p, w, tile, sy, ttu);
  else     ppc_walk<CORE, T, GRAD, false, BLOCK>(p, w, tile, sy, 0.0f);
}

// pp_pack2_kernel's record layout with the radius lane of the selected core (core_radius_term); padding records keep
// zero strength and a unit lane: they add exactly 0 under every core.
__global__ void ppc_pack2_kernel(int core, int64_t ns, int64_t ns_pad, const float* sx, const float* sy, const float* sz,
                                 const float* sr, const float* wx, const float* wy, const float* wz, float4* out) {
  const int64_t pr = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // pair index
  if (2 * pr >= ns_pad) return;
  float x[2], y[2], z[2], rt[2], a[2], b[2], c[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int64_t j = 2 * pr + h;
    x[h] = y[h] = z[h] = 0.f; rt[h] = 1.f; a[h] = b[h] = c[h] = 0.f;
    if (j < ns) {
      x[h] = -sx[j]; y[h] = -sy[j]; z[h] = -sz[j]; rt[h] = core_radius_term(core, sr ? sr[j] : 0.0f);
      a[h] = wx[j]; b[h] = wy[j]; c[h] = wz[j];
    }
  }
  out[4 * pr + 0] = make_float4(x[0], x[1], y[0], y[1]);
  out[4 * pr + 1] = make_float4(z[0], z[1], rt[0], rt[1]);
  out[4 * pr + 2] = make_float4(a[0], a[1], b[0], b[1]);
  out[4 * pr + 3] = make_float4(c[0], c[1], 0.f, 0.f);
}

}  // namespace o3d

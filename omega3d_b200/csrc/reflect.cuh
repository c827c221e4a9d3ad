// reflect.cuh - particle x panel closest-point loops.
//
// Replaces the O(N_particles x N_panels) OpenMP loops the reference runs every step when bodies exist
// (paths relative to /root/reference):
//   panel_point_distance<S>   src/Reflect.h:55-188   closest point of a triangle (node / edge / face) to a point
//   reflect_panp2<S>          src/Reflect.h:194-311  particles found under the surface are mirrored back out
//   clear_inner_panp2<S>      src/Reflect.h:446-620  method 1: particles closer than cutoff_mult * ips are pushed out
// called from reflect_interior (:313-335, Diffusion.h:292) and clear_inner_layer (:625-655, Convection.h:260-556,
// Diffusion.h:306, Simulation.cpp:839).
//
// The outcome per particle is decided by hard comparisons (which feature of which panel is closest, ties within
// 10 epsilon, the sign of a dot product), and the reference's scan over the panels is ORDER DEPENDENT: a candidate
// within eps of the running minimum joins the hit list, one clearly below it restarts the list. So this is
// integer-like work: the kernel keeps the reference's panel order per particle and spells every float operation
// unfused, in the reference's operation order (its `1.0 / x` is a double division rounded to float), and returns the
// same bits - positions and counts (tests/test_gpu_reflect.py). One thread owns one particle; panels stream through
// shared memory as 144-byte records, every LDS a warp-wide broadcast. Everything panel_point_distance computes from the
// PANEL ALONE - the three edge vectors, 1 / |edge|^2 (the double divisions) and the three normal x edge vectors of the
// prism test, 48 of its ~150 flops - is formed once per panel by ref_pack_kernel, with the same unfused operations in the
// same order, and read from the record: the same bits reach every comparison. Bound: FP32/issue, parallel over particles.
#pragma once
#include "o3d_common.cuh"

namespace o3d {

constexpr int kRefTile = 256;   // panels per shared-memory tile (36 KB)
constexpr int kRefRec = 9;      // float4 per packed panel: 3 nodes + normal | 3 x (edge, 1/|edge|^2) | 3 x (normal x edge)

struct Closest {
  float distsq, cpx, cpy, cpz;
};

__device__ __forceinline__ float dot3_rn(float a0, float a1, float a2, float b0, float b1, float b2) {
  return __fadd_rn(__fadd_rn(__fmul_rn(a0, b0), __fmul_rn(a1, b1)), __fmul_rn(a2, b2));   // src/MathHelper.h:195-198
}
__device__ __forceinline__ void cross3_rn(float a0, float a1, float a2, float b0, float b1, float b2, float& r0, float& r1, float& r2) {
  r0 = __fsub_rn(__fmul_rn(a1, b2), __fmul_rn(a2, b1));                                    // src/MathHelper.h:208-214
  r1 = __fsub_rn(__fmul_rn(a2, b0), __fmul_rn(a0, b2));
  r2 = __fsub_rn(__fmul_rn(a0, b1), __fmul_rn(a1, b0));
}
// "const S r = 1.0 / x" : a double division stored into a float
__device__ __forceinline__ float recip_via_double(float x) { return __double2float_rn(__ddiv_rn(1.0, (double)x)); }

// one edge of panel_point_distance (src/Reflect.h:107-159): a -> b, dt = t - a; inv = 1 / |e|^2 from the record
__device__ __forceinline__ void closest_edge(float ax, float ay, float az, float ex, float ey, float ez, float inv, float dx, float dy,
                                             float dz, Closest& r) {
  float rx, ry, rz;
  cross3_rn(ex, ey, ez, dx, dy, dz, rx, ry, rz);
  const float d = __fmul_rn(dot3_rn(rx, ry, rz, rx, ry, rz), inv);
  if (d < r.distsq) {
    const float t = __fmul_rn(dot3_rn(ex, ey, ez, dx, dy, dz), inv);
    if (0.0f < t && t < 1.0f) {
      r.distsq = d;
      r.cpx = __fadd_rn(ax, __fmul_rn(t, ex));
      r.cpy = __fadd_rn(ay, __fmul_rn(t, ey));
      r.cpz = __fadd_rn(az, __fmul_rn(t, ez));
    }
  }
}

// src/Reflect.h:55-188; rec = the panel's kRefRec float4 (ref_pack_kernel)
__device__ __forceinline__ Closest panel_point_distance(const float4* __restrict__ rec, float tx, float ty, float tz) {
  const float4 p0 = rec[0], p1 = rec[1], p2 = rec[2];
  const float x0 = p0.x, y0 = p0.y, z0 = p0.z, x1 = p0.w, y1 = p1.x, z1 = p1.y, x2 = p1.z, y2 = p1.w, z2 = p2.x;
  const float nx = p2.y, ny = p2.z, nz = p2.w;
  Closest r;
  r.distsq = 9.9e+9f;
  r.cpx = r.cpy = r.cpz = 0.0f;
  // the three corners (:67-99)
  const float d0x = __fsub_rn(tx, x0), d0y = __fsub_rn(ty, y0), d0z = __fsub_rn(tz, z0);
  const float d0 = dot3_rn(d0x, d0y, d0z, d0x, d0y, d0z);
  if (d0 < r.distsq) { r.distsq = d0; r.cpx = x0; r.cpy = y0; r.cpz = z0; }
  const float d1x = __fsub_rn(tx, x1), d1y = __fsub_rn(ty, y1), d1z = __fsub_rn(tz, z1);
  const float d1 = dot3_rn(d1x, d1y, d1z, d1x, d1y, d1z);
  if (d1 < r.distsq) { r.distsq = d1; r.cpx = x1; r.cpy = y1; r.cpz = z1; }
  const float d2x = __fsub_rn(tx, x2), d2y = __fsub_rn(ty, y2), d2z = __fsub_rn(tz, z2);
  const float d2 = dot3_rn(d2x, d2y, d2z, d2x, d2y, d2z);
  if (d2 < r.distsq) { r.distsq = d2; r.cpx = x2; r.cpy = y2; r.cpz = z2; }
  // the three edges (:107-159)
  const float4 e01 = rec[3], e12 = rec[4], e20 = rec[5];
  closest_edge(x0, y0, z0, e01.x, e01.y, e01.z, e01.w, d0x, d0y, d0z, r);
  closest_edge(x1, y1, z1, e12.x, e12.y, e12.z, e12.w, d1x, d1y, d1z, r);
  closest_edge(x2, y2, z2, e20.x, e20.y, e20.z, e20.w, d2x, d2y, d2z, r);
  // inside the panel's prism (:163-184); normal x edge from the record
  const float4 i01 = rec[6], i12 = rec[7], i20 = rec[8];
  const float in01 = dot3_rn(d0x, d0y, d0z, i01.x, i01.y, i01.z);
  const float in12 = dot3_rn(d1x, d1y, d1z, i12.x, i12.y, i12.z);
  const float in20 = dot3_rn(d2x, d2y, d2z, i20.x, i20.y, i20.z);
  if (in01 > 0.0f && in12 > 0.0f && in20 > 0.0f) {
    const float td = dot3_rn(d0x, d0y, d0z, nx, ny, nz);
    r.distsq = __fmul_rn(td, td);
    r.cpx = __fsub_rn(tx, __fmul_rn(nx, td));
    r.cpy = __fsub_rn(ty, __fmul_rn(ny, td));
    r.cpz = __fsub_rn(tz, __fmul_rn(nz, td));
  }
  return r;
}

// nodes SoA + connectivity + normals SoA (3 x np, stride np) -> kRefRec float4 per panel, padded to whole tiles (padding
// records are never read: the kernel stops at the last real panel). The derived fields are the reference's own
// per-pair expressions (src/Reflect.h:107-113 edge and 1.0/dot(edge,edge), :163-170 cross(normal, edge)), evaluated once.
__global__ void ref_pack_kernel(int64_t np, int64_t np_pad, const float* nx, const float* ny, const float* nz, const uint32_t* idx,
                                const float* nrm, float4* out) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= np_pad) return;
  float4 a = make_float4(3.0e18f, 0.f, 0.f, 3.0e18f), b = make_float4(1.f, 0.f, 3.0e18f, 0.f), c = make_float4(1.f, 0.f, 0.f, 1.f);
  if (j < np) {
    const uint32_t i0 = idx[3 * j], i1 = idx[3 * j + 1], i2 = idx[3 * j + 2];
    a = make_float4(nx[i0], ny[i0], nz[i0], nx[i1]);
    b = make_float4(ny[i1], nz[i1], nx[i2], ny[i2]);
    c = make_float4(nz[i2], nrm[j], nrm[np + j], nrm[2 * np + j]);
  }
  float4* o = out + (size_t)j * kRefRec;
  o[0] = a; o[1] = b; o[2] = c;
  const float x0 = a.x, y0 = a.y, z0 = a.z, x1 = a.w, y1 = b.x, z1 = b.y, x2 = b.z, y2 = b.w, z2 = c.x;
  const float e[3][3] = {{__fsub_rn(x1, x0), __fsub_rn(y1, y0), __fsub_rn(z1, z0)},
                         {__fsub_rn(x2, x1), __fsub_rn(y2, y1), __fsub_rn(z2, z1)},
                         {__fsub_rn(x0, x2), __fsub_rn(y0, y2), __fsub_rn(z0, z2)}};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    o[3 + k] = make_float4(e[k][0], e[k][1], e[k][2], recip_via_double(dot3_rn(e[k][0], e[k][1], e[k][2], e[k][0], e[k][1], e[k][2])));
    float ix, iy, iz;
    cross3_rn(c.y, c.z, c.w, e[k][0], e[k][1], e[k][2], ix, iy, iz);
    o[6 + k] = make_float4(ix, iy, iz, 0.f);
  }
}

struct ReflectArgs {
  const float4* pan;      // packed panels, whole tiles
  int64_t np;             // real panels (padding beyond is skipped)
  int ntiles;
  int64_t nt;
  float* tx; float* ty; float* tz;   // particle positions, updated in place
  int mode;               // 0: reflect_panp2; 1: clear_inner_panp2 method 1
  float cutoff;           // mode 1: _cutoff_mult * _ips (formed by the caller as the reference forms it, in float)
  unsigned long long* count;         // particles moved
};

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) reflect_kernel(const ReflectArgs p) {
  __shared__ float4 tile[kRefTile * kRefRec];
  const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
  const int64_t ic = min(i, p.nt - 1);
  const float tx = p.tx[ic], ty = p.ty[ic], tz = p.tz[ic];
  const float eps = 10.0f * 1.1920928955078125e-07f;        // 10 * numeric_limits<float>::epsilon(), src/Reflect.h:208
  float mindist = 3.402823466e+38f;
  float cnt = 0.0f, d_first = 0.0f;
  float nsx = 0.f, nsy = 0.f, nsz = 0.f, csx = 0.f, csy = 0.f, csz = 0.f;

  for (int k = 0; k < p.ntiles; ++k) {
    __syncthreads();
    for (int q = threadIdx.x; q < kRefTile * kRefRec; q += BLOCK) tile[q] = p.pan[(size_t)k * kRefTile * kRefRec + q];
    __syncthreads();
    const int nj = (int)min((int64_t)kRefTile, p.np - (int64_t)k * kRefTile);
#pragma unroll 2
    for (int j = 0; j < nj; ++j) {
      const float4 p2 = tile[kRefRec * j + 2];
      const Closest r = panel_point_distance(tile + kRefRec * j, tx, ty, tz);
      // the reference's hit list (src/Reflect.h:226-243), reduced on the fly to what is read from it afterwards:
      // the sums of normals and contact points in hit order, the count, and the first hit's distance
      if (r.distsq < __fsub_rn(mindist, eps)) {
        mindist = r.distsq;
        cnt = 1.0f; d_first = r.distsq;
        nsx = p2.y; nsy = p2.z; nsz = p2.w;
        csx = r.cpx; csy = r.cpy; csz = r.cpz;
      } else if (r.distsq < __fadd_rn(mindist, eps)) {
        cnt = __fadd_rn(cnt, 1.0f);
        nsx = __fadd_rn(nsx, p2.y); nsy = __fadd_rn(nsy, p2.z); nsz = __fadd_rn(nsz, p2.w);
        csx = __fadd_rn(csx, r.cpx); csy = __fadd_rn(csy, r.cpy); csz = __fadd_rn(csz, r.cpz);
      }
    }
  }

  bool moved = false;
  if (i < p.nt && cnt > 0.0f) {
    // mean normal (normalizeVec, src/MathHelper.h:186-192) and mean contact point (src/Reflect.h:283-285)
    const float len = recip_via_double(__fsqrt_rn(dot3_rn(nsx, nsy, nsz, nsx, nsy, nsz)));
    const float mx = __fmul_rn(nsx, len), my = __fmul_rn(nsy, len), mz = __fmul_rn(nsz, len);
    const float cx = __fdiv_rn(csx, cnt), cy = __fdiv_rn(csy, cnt), cz = __fdiv_rn(csz, cnt);
    float dotp = dot3_rn(mx, my, mz, __fsub_rn(tx, cx), __fsub_rn(ty, cy), __fsub_rn(tz, cz));
    if (p.mode == 0) {
      if (dotp < 0.0f) {                                    // under the surface: mirror off the first hit (:291-299)
        const float dist = __fsqrt_rn(d_first);
        p.tx[i] = __fadd_rn(cx, __fmul_rn(dist, mx));
        p.ty[i] = __fadd_rn(cy, __fmul_rn(dist, my));
        p.tz[i] = __fadd_rn(cz, __fmul_rn(dist, mz));
        moved = true;
      }
    } else {
      dotp = __fsub_rn(dotp, p.cutoff);                     // height above the cutoff layer (:537)
      if (dotp < 0.0f) {                                    // method 1: push out, keep all strength (:569-578)
        p.tx[i] = __fsub_rn(tx, __fmul_rn(dotp, mx));
        p.ty[i] = __fsub_rn(ty, __fmul_rn(dotp, my));
        p.tz[i] = __fsub_rn(tz, __fmul_rn(dotp, mz));
        moved = true;
      }
    }
  }
  const unsigned m = __popc(__ballot_sync(0xffffffffu, moved));
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(p.count, (unsigned long long)m);
}

}  // namespace o3d

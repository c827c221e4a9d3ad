// convect.cuh - the O(N) steps that bracket every Biot-Savart evaluation, kept on the device so that a
// convection step never leaves HBM.
//
// Replaces, for vortex-particle collections (paths relative to /root/reference):
//   zero_vels       src/ElementBase.h:170-184, src/Points.h:252-263
//   finalize_vels   src/ElementBase.h:187-209 (u = fs + u * 1/4pi, in double), src/Points.h:265-277 (grads * float(1/4pi))
//   move (1 stage)  src/ElementBase.h:253-273 + src/Points.h:288-351   Euler advection + vortex stretching
//   move (2 stage)  src/ElementBase.h:276-304 + src/Points.h:356-439   RK2 combination
//   move (3 stage)  src/ElementBase.h:307-336 + src/Points.h:444-520   RK3 combination
//   get_max_str     src/ElementBase.h:339-351, get_max_elong src/Points.h:523-532
//
// These are bandwidth-trivial (tens of bytes per particle against N x 70 flops per particle in the influence
// kernel); what matters is that they round EXACTLY as the reference's scalar build does, so that a device-resident
// step and a host step fed the same velocities produce the same bits. Every operation below is therefore spelled
// with an explicit round-to-nearest intrinsic in the reference's own operation order and operand types
// (float S, double weights/dt): nvcc must not contract a*b+c into an FMA here.
#pragma once
#include "o3d_common.cuh"

namespace o3d {

// one Runge-Kutta stage as the reference's move() sees it: the velocity and (optionally) the velocity gradient
// of a Points object evaluated at that stage
struct StageRef {
  const float* u[3];
  const float* ug;        // 9 rows of stride ug_stride (row 3*j+i = d u_i / d x_j) or nullptr
  int64_t ug_stride;
};

struct MoveArgs {
  int64_t n;
  double dt;
  double wt[3];
  StageRef st[3];
  const float* xin[3];    // state read ...
  const float* sin[3];    // ... strengths (nullptr: inert points, no stretching)
  const float* ein;       // elongation (nullptr: not tracked - interim copies of a Runge-Kutta step)
  float* xout[3];         // ... and written (may alias the inputs)
  float* sout[3];
  float* eout;
  float* uout[3];         // ORDER >= 2 stores the combined velocity here, as the reference does into this->u
};

__device__ __forceinline__ float f_dot3(float a0, float b0, float a1, float b1, float a2, float b2) {
  // (a0*b0 + a1*b1) + a2*b2, three products and two sums, each rounded to float
  return __fadd_rn(__fadd_rn(__fmul_rn(a0, b0), __fmul_rn(a1, b1)), __fmul_rn(a2, b2));
}

// w . grad u for one stage: wdu_i = s0 ug[i] + s1 ug[3+i] + s2 ug[6+i]   (src/Points.h:316-318)
__device__ __forceinline__ void stretch_term(const StageRef& st, int64_t i, float s0, float s1, float s2, float (&wdu)[3]) {
  const float* g = st.ug + i;
  const int64_t p = st.ug_stride;
#pragma unroll
  for (int k = 0; k < 3; ++k) wdu[k] = f_dot3(s0, g[(size_t)k * p], s1, g[(size_t)(3 + k) * p], s2, g[(size_t)(6 + k) * p]);
}

template <int ORDER>
__global__ void __launch_bounds__(256) pts_move_kernel(const MoveArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const float dtf = __double2float_rn(a.dt);   // (S)_dt

  // ---- advection ----
  float unew[3];
  if constexpr (ORDER == 1) {
    // x += (S)_dt * _wt1 * u1 : ((double)(float)dt * wt1) * (double)u, added to x in double (src/ElementBase.h:264)
    const double w = __dmul_rn((double)dtf, a.wt[0]);
#pragma unroll
    for (int d = 0; d < 3; ++d)
      a.xout[d][i] = __double2float_rn(__dadd_rn((double)a.xin[d][i], __dmul_rn(w, (double)a.st[0].u[d][i])));
  } else {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      // u = wt1*u1 + wt2*u2 (+ wt3*u3) in double, stored as float (src/ElementBase.h:287,319)
      double c = __dadd_rn(__dmul_rn(a.wt[0], (double)a.st[0].u[d][i]), __dmul_rn(a.wt[1], (double)a.st[1].u[d][i]));
      if constexpr (ORDER == 3) c = __dadd_rn(c, __dmul_rn(a.wt[2], (double)a.st[2].u[d][i]));
      unew[d] = __double2float_rn(c);
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      // x += (S)_dt * u : float product, float sum (src/ElementBase.h:294,326)
      a.xout[d][i] = __fadd_rn(a.xin[d][i], __fmul_rn(dtf, unew[d]));
      a.uout[d][i] = unew[d];
    }
  }

  // ---- stretching: active particles whose stages all carry gradients (src/Points.h:296,366,453) ----
  if (a.sin[0] == nullptr) return;
  const float s0 = a.sin[0][i], s1 = a.sin[1][i], s2 = a.sin[2][i];
  bool have = a.st[0].ug != nullptr;
  if constexpr (ORDER >= 2) have = have && a.st[1].ug != nullptr;
  if constexpr (ORDER == 3) have = have && a.st[2].ug != nullptr;
  if (!have) {
    if (a.sout[0] != a.sin[0]) { a.sout[0][i] = s0; a.sout[1][i] = s1; a.sout[2][i] = s2; }
    if (a.eout && a.eout != a.ein) a.eout[i] = a.ein[i];
    return;
  }

  float wdu[3];
  stretch_term(a.st[0], i, s0, s1, s2, wdu);
  if constexpr (ORDER >= 2) {
    float w2[3], w3[3];
    stretch_term(a.st[1], i, s0, s1, s2, w2);
    if constexpr (ORDER == 3) stretch_term(a.st[2], i, s0, s1, s2, w3);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      // wdu = _wt1*wdu1 + _wt2*wdu2 (+ _wt3*wdu3): double arithmetic stored into a float array (src/Points.h:399-402,498-501)
      double c = __dadd_rn(__dmul_rn(a.wt[0], (double)wdu[k]), __dmul_rn(a.wt[1], (double)w2[k]));
      if constexpr (ORDER == 3) c = __dadd_rn(c, __dmul_rn(a.wt[2], (double)w3[k]));
      wdu[k] = __double2float_rn(c);
    }
  }

  // elongation (src/Points.h:321-325,405-409,504-508)
  if (a.eout) {
    float e = a.ein[i];
    const float circ = f_dot3(s0, s0, s1, s1, s2, s2);
    if (circ > 0.0f) {
      const float sd = f_dot3(s0, wdu[0], s1, wdu[1], s2, wdu[2]);
      float ef;
      if constexpr (ORDER == 1)   // (S)_dt * _wt1 * sd / circ : double until the store into "const S"
        ef = __double2float_rn(__ddiv_rn(__dmul_rn(__dmul_rn((double)dtf, a.wt[0]), (double)sd), (double)circ));
      else                        // (S)_dt * sd / circ : all float
        ef = __fdiv_rn(__fmul_rn(dtf, sd), circ);
      e = __double2float_rn(__dmul_rn((double)e, __dadd_rn(1.0, (double)ef)));   // elong *= 1.0 + elongfactor
    }
    a.eout[i] = e;
  }

  // strengths: s = this_s + _dt * [_wt1 *] wdu in double, stored as float (src/Points.h:330-332,414-416,513-515)
  const double w = ORDER == 1 ? __dmul_rn(a.dt, a.wt[0]) : a.dt;
  a.sout[0][i] = __double2float_rn(__dadd_rn((double)s0, __dmul_rn(w, (double)wdu[0])));
  a.sout[1][i] = __double2float_rn(__dadd_rn((double)s1, __dmul_rn(w, (double)wdu[1])));
  a.sout[2][i] = __double2float_rn(__dadd_rn((double)s2, __dmul_rn(w, (double)wdu[2])));
}

// u = fs + u * (0.25/pi) in double; grads *= float(0.25/pi)
__global__ void __launch_bounds__(256) pts_finalize_kernel(int64_t n, float* u, float* v, float* w, float* ug, int64_t ug_stride,
                                                            double fs0, double fs1, double fs2) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double factor = 0.25 / 3.14159265358979323846;
  u[i] = __double2float_rn(__dadd_rn(fs0, __dmul_rn((double)u[i], factor)));
  v[i] = __double2float_rn(__dadd_rn(fs1, __dmul_rn((double)v[i], factor)));
  w[i] = __double2float_rn(__dadd_rn(fs2, __dmul_rn((double)w[i], factor)));
  if (ug) {
    const float ff = __double2float_rn(factor);
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      float* g = ug + (size_t)k * ug_stride + i;
      *g = __fmul_rn(*g, ff);
    }
  }
}

// rows x n block set to `value` (rows of stride `stride`); value = 0 is zero_vels
__global__ void __launch_bounds__(256) pts_fill_kernel(int64_t n, int rows, float* base, int64_t stride, float value) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int k = 0; k < rows; ++k) base[(size_t)k * stride + i] = value;
}

// out[0] = max over particles of |s|^2 as the reference forms it ((s0*s0 + s1*s1) + s2*s2), out[1] = max elongation.
// Both are non-negative floats, whose bit patterns order like unsigned integers; zero-initialise `out`.
__global__ void __launch_bounds__(256) pts_stats_kernel(int64_t n, const float* s0, const float* s1, const float* s2,
                                                         const float* elong, uint32_t* out) {
  float ms = 0.0f, me = 0.0f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (s0) ms = fmaxf(ms, f_dot3(s0[i], s0[i], s1[i], s1[i], s2[i], s2[i]));
    if (elong) me = fmaxf(me, elong[i]);
  }
  const uint32_t a = __reduce_max_sync(0xffffffffu, __float_as_uint(ms));
  const uint32_t b = __reduce_max_sync(0xffffffffu, __float_as_uint(me));
  if ((threadIdx.x & 31) == 0) {
    atomicMax(out + 0, a);
    atomicMax(out + 1, b);
  }
}

// Total circulation and linear impulse of a collection (the status-file quantities):
//   circ = sum_i s_i                                        ElementBase::get_total_circ  (src/ElementBase.h:354-378)
//   imp  = sum_i (s1 x2 - s2 x1, s2 x0 - s0 x2, s0 x1 - s1 x0)   Points::get_total_impulse  (src/Points.h:547-563)
// The per-particle terms are formed in float with the reference's unfused operations; the SUMS are FP64 trees of fixed shape
// (grid-stride per thread, shuffle tree per warp, warps in order, blocks in order in pts_totals_finish_kernel): deterministic,
// and exact to ~1e-16 where the reference's own sums carry a double sequential (circulation) or a float sequential (impulse)
// rounding. part: [gridDim.x][6] doubles.
constexpr int kTotalsBlock = 256;
__global__ void __launch_bounds__(kTotalsBlock) pts_totals_kernel(int64_t n, const float* x0, const float* x1, const float* x2, const float* s0,
                                                                  const float* s1, const float* s2, double* part) {
  double a[6] = {0, 0, 0, 0, 0, 0};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float px = x0[i], py = x1[i], pz = x2[i], wx = s0[i], wy = s1[i], wz = s2[i];
    a[0] += (double)wx; a[1] += (double)wy; a[2] += (double)wz;
    a[3] += (double)__fsub_rn(__fmul_rn(wy, pz), __fmul_rn(wz, py));
    a[4] += (double)__fsub_rn(__fmul_rn(wz, px), __fmul_rn(wx, pz));
    a[5] += (double)__fsub_rn(__fmul_rn(wx, py), __fmul_rn(wy, px));
  }
  __shared__ double warp_sum[kTotalsBlock / 32][6];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) a[k] += __shfl_down_sync(0xffffffffu, a[k], off);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5][k] = a[k];
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double t = 0.0;
    for (int w = 0; w < kTotalsBlock / 32; ++w) t += warp_sum[w][threadIdx.x];
    part[(size_t)blockIdx.x * 6 + threadIdx.x] = t;
  }
}
__global__ void pts_totals_finish_kernel(int nblocks, const double* part, double* out) {
  if (threadIdx.x < 6) {
    double t = 0.0;
    for (int b = 0; b < nblocks; ++b) t += part[(size_t)b * 6 + threadIdx.x];
    out[threadIdx.x] = t;
  }
}

}  // namespace o3d

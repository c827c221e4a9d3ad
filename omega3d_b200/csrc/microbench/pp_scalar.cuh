// pp_scalar.cuh - the scalar-FFMA particles -> points kernel (one source per instruction) with the gridDim.y source split.
// NOT a product kernel: kept for microbench/kbench.cu, whose shoot-out against the packed-FP32 kernel (../biot_pp.cuh) is the
// "alternatives measured and rejected" table of DESIGN.md section 7.
#pragma once
#include "../biot_pp.cuh"

namespace o3d {

struct PPScalarArgs {
  const float4* src;      // unpaired stream (pp_pack_kernel): 2 float4 per source, padded to whole tiles
  int ntiles;             // tiles in the whole stream
  int nsplit;             // source slices = gridDim.y
  int64_t nt;             // targets
  const float* tx; const float* ty; const float* tz;
  const float* tr;        // nullptr => singular targets (tr = 0)
  float* tu; float* tv; float* tw;   // velocity, read-modify-write
  float* tug;             // 9 rows of stride tug_stride, or nullptr
  int64_t tug_stride;
  double* partial;        // nsplit > 1: [nsplit][12 or 3][nt] FP64 workspace, one slab per source slice
  float sign;
};

// One source record (a = x y z sr^2, b = wx wy wz -) on one target. acc layout:
//   [0..2] u v w | [3..11] G[3j+i] = sum d_j * bbb*c_i | [12..14] A = sum w * r3
template <bool GRAD>
__device__ __forceinline__ void pp_interact(const float4 a, const float4 b, const float tx, const float ty,
                                            const float tz, const float tr2, float (&acc)[PPAcc<GRAD>::N]) {
  const float dx = tx - a.x, dy = ty - a.y, dz = tz - a.z;
  const float r2 = a.w + tr2;
  const float d2 = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, r2)));
  const float top = fmaf(1.5f, r2, d2);
  const float rs = rsqrt_approx(d2);
  const float rs2 = rs * rs;
  const float rs4 = rs2 * rs2;
  const float dn5 = rs4 * rs;
  const float r3 = top * dn5;
  float cx = fmaf(dz, b.y, -(dy * b.z));
  float cy = fmaf(dx, b.z, -(dz * b.x));
  float cz = fmaf(dy, b.x, -(dx * b.y));
  acc[0] = fmaf(r3, cx, acc[0]);
  acc[1] = fmaf(r3, cy, acc[1]);
  acc[2] = fmaf(r3, cz, acc[2]);
  if constexpr (GRAD) {
    const float bbb = dn5 * fmaf(-5.0f, top * rs2, 2.0f);
    cx *= bbb; cy *= bbb; cz *= bbb;
    acc[3]  = fmaf(dx, cx, acc[3]);
    acc[4]  = fmaf(dx, cy, acc[4]);
    acc[5]  = fmaf(dx, cz, acc[5]);
    acc[6]  = fmaf(dy, cx, acc[6]);
    acc[7]  = fmaf(dy, cy, acc[7]);
    acc[8]  = fmaf(dy, cz, acc[8]);
    acc[9]  = fmaf(dz, cx, acc[9]);
    acc[10] = fmaf(dz, cy, acc[10]);
    acc[11] = fmaf(dz, cz, acc[11]);
    acc[12] = fmaf(b.x, r3, acc[12]);
    acc[13] = fmaf(b.y, r3, acc[13]);
    acc[14] = fmaf(b.z, r3, acc[14]);
  }
}

// Scalar-FFMA kernel: T register-blocked targets per thread.
template <int T, bool GRAD, int BLOCK>
__global__ void __launch_bounds__(BLOCK) pp_kernel(const PPScalarArgs p) {
  constexpr int NA = PPAcc<GRAD>::N;
  constexpr int NS = GRAD ? 12 : 3;
  __shared__ alignas(128) float4 tile[2][kTile * 2];
  __shared__ alignas(8) uint64_t full[2];

  // this CTA's slice of the source stream
  const int per = (p.ntiles + p.nsplit - 1) / p.nsplit;
  const int k0 = blockIdx.y * per;
  const int k1 = min(p.ntiles, k0 + per);
  const int nk = k1 - k0;

  if (threadIdx.x == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < 2; ++s)
      if (s < nk) {
        mbar_expect_tx(&full[s], kTileBytes);
        bulk_g2s(tile[s], p.src + (size_t)(k0 + s) * (kTile * 2), kTileBytes, &full[s]);
      }
  }

  float tx[T], ty[T], tz[T], tr2[T];
  const int64_t base = (int64_t)blockIdx.x * (BLOCK * T) + threadIdx.x;
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const int64_t i = min(base + (int64_t)t * BLOCK, p.nt - 1);
    tx[t] = p.tx[i]; ty[t] = p.ty[i]; tz[t] = p.tz[i];
    const float r = p.tr ? p.tr[i] : 0.0f;
    tr2[t] = r * r;
  }

  float acc[T][NA];
  double sum[T][NS];
#pragma unroll
  for (int t = 0; t < T; ++t) {
#pragma unroll
    for (int k = 0; k < NA; ++k) acc[t][k] = 0.0f;
#pragma unroll
    for (int k = 0; k < NS; ++k) sum[t][k] = 0.0;
  }

  for (int k = 0; k < nk; ++k) {
    const int buf = k & 1;
    mbar_wait(&full[buf], (k >> 1) & 1);
    const float4* __restrict__ s = tile[buf];
#pragma unroll 4
    for (int j = 0; j < kTile; ++j) {
      const float4 a = s[2 * j], b = s[2 * j + 1];
#pragma unroll
      for (int t = 0; t < T; ++t) pp_interact<GRAD>(a, b, tx[t], ty[t], tz[t], tr2[t], acc[t]);
    }
#pragma unroll
    for (int t = 0; t < T; ++t) pp_promote<GRAD>(acc[t], sum[t]);
    __syncthreads();  // every warp is done with tile[buf]; safe to refill
    if (threadIdx.x == 0 && k + 2 < nk) {
      mbar_expect_tx(&full[buf], kTileBytes);
      bulk_g2s(tile[buf], p.src + (size_t)(k0 + k + 2) * (kTile * 2), kTileBytes, &full[buf]);
    }
  }

#pragma unroll
  for (int t = 0; t < T; ++t) {
    const int64_t i = base + (int64_t)t * BLOCK;
    if (i >= p.nt) continue;
    if (p.nsplit > 1) {
      // slice blockIdx.y owns its own [NS][nt] slab: plain stores, summed in slice order by
      // pp_finish_kernel, so the result does not depend on CTA scheduling
      double* slab = p.partial + (size_t)blockIdx.y * NS * p.nt;
#pragma unroll
      for (int k = 0; k < NS; ++k) slab[(size_t)k * p.nt + i] = sum[t][k];
    } else {
      const double sg = (double)p.sign;
      p.tu[i] = (float)((double)p.tu[i] + sg * sum[t][0]);
      p.tv[i] = (float)((double)p.tv[i] + sg * sum[t][1]);
      p.tw[i] = (float)((double)p.tw[i] + sg * sum[t][2]);
      if constexpr (GRAD) {
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          float* g = p.tug + (size_t)k * p.tug_stride + i;
          *g = (float)((double)*g + sum[t][3 + k]);
        }
      }
    }
  }
}

// SoA (the reference's Points layout) -> packed record stream, with zero-strength padding records.
__global__ void pp_pack_kernel(int64_t ns, int64_t ns_pad, const float* sx, const float* sy, const float* sz,
                               const float* sr, const float* wx, const float* wy, const float* wz, float4* out) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ns_pad) return;
  float4 a = make_float4(0.f, 0.f, 0.f, 1.0f), b = make_float4(0.f, 0.f, 0.f, 0.f);
  if (j < ns) {
    const float r = sr ? sr[j] : 0.0f;
    a = make_float4(sx[j], sy[j], sz[j], r * r);
    b = make_float4(wx[j], wy[j], wz[j], 0.f);
  }
  out[2 * j] = a;
  out[2 * j + 1] = b;
}

}  // namespace o3d

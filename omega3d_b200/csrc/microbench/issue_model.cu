// issue_model.cu - what limits a packed-FP32 (FFMA2) instruction stream on sm_100a besides the FMA pipe?
// Each kernel runs 16 FFMA2 per loop trip in 8 independent chains plus a variable "rider":
//   base      : acc = acc*a + b           (operands a, b served by the operand-reuse cache: 1 fresh 64-bit read)
//   three     : acc = x_i*y_j + acc       (3 distinct register pairs, no reuse possible)
//   two       : acc = x_i*x_i + acc       (2 distinct pairs)
//   +ffma     : base + 8 independent scalar FFMA          (does the second FMA pipe take them for free?)
//   +mufu     : base + 2 MUFU.RSQ                         (XU pipe)
//   +lds      : base + 2 LDS.128                          (LSU)
//   +alu      : base + 4 integer LOP3/IADD3                (ALU pipe)
// Reported: cycles per trip per warp scheduler (clock64), so DVFS does not matter. Development tool.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
constexpr int ITERS = 8192;

template <int MODE>
__global__ void k(float* out, long long* cyc, float a, float b, int dummy) {
  __shared__ float4 sm[256];
  float2 acc[8], x[8], y[8];
  float sacc[8];
  float m0 = 1.5f + threadIdx.x, m1 = 2.5f + threadIdx.x;
  int ia = threadIdx.x, ib = dummy;
  float4 l0 = make_float4(0, 0, 0, 0), l1 = l0;
  sm[threadIdx.x & 255] = make_float4(a, b, a, b);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
    x[i] = make_float2(a + i * 1e-4f, a - i * 1e-4f);
    y[i] = make_float2(b + i * 1e-4f, b - i * 1e-4f);
    sacc[i] = threadIdx.x * 3e-3f + i;
  }
  const float2 aa = make_float2(a, a * 1.0001f), bb = make_float2(b, b * 0.999f);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 1) acc[i] = __ffma2_rn(x[i], y[(i + r + 1) & 7], acc[i]);
        else if (MODE == 2) acc[i] = __ffma2_rn(x[(i + r) & 7], x[(i + r) & 7], acc[i]);
        else acc[i] = __ffma2_rn(acc[i], aa, bb);
        if (MODE == 3 && r == 0) sacc[i] = fmaf(sacc[i], a, b);
        if (MODE == 7) sacc[i] = fmaf(sacc[i], a, b);
      }
      if (MODE == 4) { if (r == 0) m0 = rsqrtf(m0 + 1.0f); else m1 = rsqrtf(m1 + 1.0f); }
      if (MODE == 5) { if (r == 0) l0 = sm[(ia + it) & 255]; else l1 = sm[(ia + it + 7) & 255]; }
      if (MODE == 6) { ia = (ia ^ ib) + it; ib = (ib & ia) + r; }
    }
    if (MODE == 5) { acc[0].x += l0.x + l1.y; }
  }
  long long t1 = clock64();
  float s = m0 + m1 + ia + ib + l0.x + l1.x;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y + sacc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  cudaDeviceProp prop; CHECK(cudaGetDeviceProperties(&prop, 0));
  const int blocks = prop.multiProcessorCount, threads = 512;   // 4 warps per scheduler
  float* out; long long* cyc; CHECK(cudaMalloc(&out, blocks * threads * 4)); CHECK(cudaMalloc(&cyc, blocks * 8));
  const char* names[] = {"base (reuse, 1 fresh read)", "three distinct pairs", "two distinct pairs", "base + 8 scalar FFMA", "base + 2 MUFU.RSQ",
                         "base + 2 LDS.128", "base + 4 ALU int ops", "base + 16 scalar FFMA"};
  void (*ks[])(float*, long long*, float, float, int) = {k<0>, k<1>, k<2>, k<3>, k<4>, k<5>, k<6>, k<7>};
  for (int m = 0; m < 8; ++m) {
    for (int rep = 0; rep < 2; ++rep) ks[m]<<<blocks, threads>>>(out, cyc, 0.999f, 1e-3f, 3);
    CHECK(cudaDeviceSynchronize());
    long long h[1024]; CHECK(cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
    // 16 warps per SM = 4 per scheduler; each trip issues 16 FFMA2 per warp
    const double cyc_per_trip_per_sched = avg / ITERS;   // wall cycles per trip with 4 warps sharing a scheduler
    printf("%-28s cycles/trip (4 warps/scheduler) %7.2f  => %5.2f cycles per FFMA2 per scheduler\n", names[m], cyc_per_trip_per_sched,
           cyc_per_trip_per_sched / (16.0 * 4));
  }
  return 0;
}

// fp32_pipes.cu - issue-rate microbenchmarks behind the kernel design choices in DESIGN.md.
// Measures, per SM and per clock (clock64 deltas inside the kernel, so DVFS does not matter):
//   ffma      : 3-register scalar FFMA, 16 independent chains per thread
//   ffma_xyz  : FFMA with three DISTINCT registers per instruction (x_i*y_i+acc_i), the shape the
//               Biot-Savart inner loop has
//   ffma2     : packed fma.rn.f32x2 (sm_100 FFMA2), 8 independent 64-bit chains
//   mix40     : 40 FFMA : 1 MUFU.RSQ, the Biot-Savart velocity+gradient instruction mix
//   mix40_2   : 20 FFMA2 : 1 MUFU.RSQ ... same flops through the packed pipe
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp32_pipes fp32_pipes.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int ITERS = 4096;

__global__ void k_ffma(float* out, long long* cyc, float a, float b) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3f + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_ffma_xyz(float* out, long long* cyc, float a) {
  float acc[8], x[8], y[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { acc[i] = 0.f; x[i] = threadIdx.x * 1e-3f + i; y[i] = a + i * 1e-4f; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(x[i], y[(i + r) & 7], acc[i]);
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_ffma2(float* out, long long* cyc, float a, float b) {
  float2 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
  const float2 aa = make_float2(a, a * 1.0001f), bb = make_float2(b, b * 0.999f);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = __ffma2_rn(acc[i], aa, bb);
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_ffma2_xyz(float* out, long long* cyc, float a) {
  float2 acc[8], x[8], y[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc[i] = make_float2(0.f, 0.f);
    x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
    y[i] = make_float2(a + i * 1e-4f, a - i * 1e-4f);
  }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = __ffma2_rn(x[i], y[(i + r) & 7], acc[i]);
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// 40 FFMA + 1 MUFU.RSQ per group, 2 independent groups
__global__ void k_mix40(float* out, long long* cyc, float a, float b) {
  float acc[16], m[2];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3f + i;
  m[0] = 1.5f + threadIdx.x; m[1] = 2.5f + threadIdx.x;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      m[g] = rsqrtf(m[g] + 1.0f);
#pragma unroll
      for (int r = 0; r < 5; ++r) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[g * 8 + i] = fmaf(acc[g * 8 + i], a, m[g]);
      }
      m[g] = m[g] + acc[g * 8] * 1e-9f;
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + m[0] + m[1] + b;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// same flops through the packed pipe: 20 FFMA2 + 1 MUFU per group (per 40 scalar-equivalent FMAs)
__global__ void k_mix40_2(float* out, long long* cyc, float a, float b) {
  float2 acc[8];
  float m[2];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
  m[0] = 1.5f + threadIdx.x; m[1] = 2.5f + threadIdx.x;
  const float2 aa = make_float2(a, a * 1.0001f);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      m[g] = rsqrtf(m[g] + 1.0f);
      const float2 mm = make_float2(m[g], m[g]);
#pragma unroll
      for (int r = 0; r < 5; ++r) {
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[g * 4 + i] = __ffma2_rn(acc[g * 4 + i], aa, mm);
      }
      m[g] = m[g] + acc[g * 4].x * 1e-9f;
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + m[0] + m[1] + b;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <class F>
static void run(const char* name, F launch, int nsm, double fma_per_thread, int threads, int blocks_per_sm) {
  float* out; long long* cyc;
  const int blocks = nsm * blocks_per_sm;
  CHECK(cudaMalloc(&out, sizeof(float) * blocks * threads));
  CHECK(cudaMalloc(&cyc, sizeof(long long) * blocks));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(blocks, threads, out, cyc);  // warm
  CHECK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  launch(blocks, threads, out, cyc);
  cudaEventRecord(e1);
  CHECK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long* h = (long long*)malloc(sizeof(long long) * blocks);
  CHECK(cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost));
  double avg = 0; for (int i = 0; i < blocks; ++i) avg += (double)h[i]; avg /= blocks;
  const double fma_per_clk_sm = fma_per_thread * threads * blocks_per_sm / avg;
  const double tflops = 2.0 * fma_per_thread * threads * blocks / (ms * 1e-3) * 1e-12;
  printf("%-10s thr=%4d bps=%d  cycles=%9.0f  FMA/clk/SM=%7.2f  wall=%7.3f ms  %.2f TFLOP/s (=> %.0f MHz eff)\n",
         name, threads, blocks_per_sm, avg, fma_per_clk_sm, ms, tflops, avg / (ms * 1e-3) * 1e-6);
  free(h); cudaFree(out); cudaFree(cyc);
}

int main() {
  cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, 0));
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("device %s  SMs=%d  clockRate=%d kHz  cc=%d.%d\n", p.name, p.multiProcessorCount, clk, p.major, p.minor);
  const int nsm = p.multiProcessorCount;
  for (int cfg = 0; cfg < 3; ++cfg) {
    const int threads = cfg == 0 ? 1024 : (cfg == 1 ? 512 : 256);
    const int bps = cfg == 0 ? 1 : (cfg == 1 ? 2 : 2);
    run("ffma", [&](int b, int t, float* o, long long* c) { k_ffma<<<b, t>>>(o, c, 1.0001f, 1e-6f); }, nsm, 16.0 * ITERS, threads, bps);
    run("ffma_xyz", [&](int b, int t, float* o, long long* c) { k_ffma_xyz<<<b, t>>>(o, c, 1.0001f); }, nsm, 16.0 * ITERS, threads, bps);
    run("ffma2", [&](int b, int t, float* o, long long* c) { k_ffma2<<<b, t>>>(o, c, 1.0001f, 1e-6f); }, nsm, 32.0 * ITERS, threads, bps);
    run("ffma2_xyz", [&](int b, int t, float* o, long long* c) { k_ffma2_xyz<<<b, t>>>(o, c, 1.0001f); }, nsm, 32.0 * ITERS, threads, bps);
    run("mix40", [&](int b, int t, float* o, long long* c) { k_mix40<<<b, t>>>(o, c, 1.0001f, 1e-6f); }, nsm, 80.0 * (ITERS / 4), threads, bps);
    run("mix40_2", [&](int b, int t, float* o, long long* c) { k_mix40_2<<<b, t>>>(o, c, 1.0001f, 1e-6f); }, nsm, 80.0 * (ITERS / 4), threads, bps);
  }
  return 0;
}

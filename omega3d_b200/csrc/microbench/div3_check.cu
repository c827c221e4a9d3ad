// div3_check.cu - exhaustive check of biot_panel.cuh's div3_rn against the IEEE division it replaces (development tool).
// Sweeps all 2^32 float bit patterns on the device and reports every class of input on which
// div3_rn(x) and __fdiv_rn(x, 3.0f) differ in their bits. Build: make -C omega3d_b200/csrc div3   Run: microbench/div3_check
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../biot_panel.cuh"

__global__ void sweep(unsigned long long* out) {
  // out[0] mismatches among normal finite inputs whose quotient is normal, [1] mismatches with a subnormal input or quotient,
  // [2] zeros (sign only), [3] inf / NaN inputs, [4] smallest |x| bits of a class-0 mismatch, [5] largest
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  unsigned long long c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < (1ull << 32); b += stride) {
    const float x = __uint_as_float((uint32_t)b);
    const uint32_t ref = __float_as_uint(__fdiv_rn(x, 3.0f)), got = __float_as_uint(o3d::div3_rn(x));
    if (ref == got) continue;
    const uint32_t ax = (uint32_t)b & 0x7fffffffu, aq = ref & 0x7fffffffu;
    if (ax >= 0x7f800000u) c3++;
    else if (ax == 0u) c2++;
    else if (ax < 0x00800000u || aq < 0x00800000u) c1++;
    else {
      c0++;
      atomicMin(out + 4, (unsigned long long)ax);
      atomicMax(out + 5, (unsigned long long)ax);
    }
  }
  atomicAdd(out + 0, c0); atomicAdd(out + 1, c1); atomicAdd(out + 2, c2); atomicAdd(out + 3, c3);
}

int main() {
  unsigned long long* d;
  unsigned long long h[6] = {0, 0, 0, 0, ~0ull, 0};
  cudaMalloc(&d, sizeof h);
  cudaMemcpy(d, h, sizeof h, cudaMemcpyHostToDevice);
  sweep<<<148 * 8, 256>>>(d);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed\n"); return 1; }
  cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  printf("div3_rn vs __fdiv_rn(x, 3.0f) over all 2^32 inputs:\n");
  printf("  normal input, normal quotient : %llu mismatches", h[0]);
  if (h[0]) printf(" (|x| bits %#llx .. %#llx)", h[4], h[5]);
  printf("\n  subnormal input or quotient   : %llu\n  zeros (sign of zero)          : %llu\n  inf / NaN inputs              : %llu\n", h[1], h[2], h[3]);
  return h[0] ? 2 : 0;
}

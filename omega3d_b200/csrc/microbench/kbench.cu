// kbench.cu - times the points->points kernel variants against each other on one GPU and checks
// them against a double-precision host loop on a few targets. Development harness only (the product
// entry points are in ../capi.cu). Build: see ../Makefile (target kbench).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include <algorithm>
#include <string>
#include <cstring>
#include <cuda.h>
#include "pp_scalar.cuh"   // the scalar-FFMA kernel with its gridDim.y split (not a product kernel) + ../biot_pp.cuh
#include "../biot_pp_cores.cuh"   // KBENCH_CORE: the alternate-core kernels from cubin files

using namespace o3d;
#define O3D_STR2(x) #x
#define O3D_STR(x) O3D_STR2(x)
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

static void host_ref(int ns, const float* sx, const float* sy, const float* sz, const float* sr, const float* wx,
                     const float* wy, const float* wz, float tx, float ty, float tz, float tr, double* o) {
  for (int k = 0; k < 12; ++k) o[k] = 0;
  for (int j = 0; j < ns; ++j) {
    double dx = (double)tx - sx[j], dy = (double)ty - sy[j], dz = (double)tz - sz[j];
    double r2 = (double)sr[j] * sr[j] + (double)tr * tr, ds = dx * dx + dy * dy + dz * dz, d2 = ds + r2;
    double top = ds + 2.5 * r2, dn5 = 1.0 / (d2 * d2 * std::sqrt(d2)), r3 = top * dn5, bbb = 2 * dn5 - 5 * top * dn5 / d2;
    double cx = dz * wy[j] - dy * wz[j], cy = dx * wz[j] - dz * wx[j], cz = dy * wx[j] - dx * wy[j];
    o[0] += r3 * cx; o[1] += r3 * cy; o[2] += r3 * cz;
    cx *= bbb; cy *= bbb; cz *= bbb;
    o[3] += dx * cx; o[4] += dx * cy + wz[j] * r3; o[5] += dx * cz - wy[j] * r3;
    o[6] += dy * cx - wz[j] * r3; o[7] += dy * cy; o[8] += dy * cz + wx[j] * r3;
    o[9] += dz * cx + wy[j] * r3; o[10] += dz * cy - wx[j] * r3; o[11] += dz * cz;
  }
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 262144;
  const int reps = argc > 2 ? atoi(argv[2]) : 3;
  cudaDeviceProp prop; CHECK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s SMs=%d  N=%d\n", prop.name, prop.multiProcessorCount, n);
  std::mt19937_64 gen(20240517 + n);
  std::uniform_real_distribution<float> U(-0.5f, 0.5f);
  std::vector<float> h[7];
  for (int c = 0; c < 7; ++c) h[c].resize(n);
  for (int i = 0; i < n; ++i) { h[0][i] = U(gen); h[1][i] = U(gen); h[2][i] = U(gen); }
  for (int i = 0; i < n; ++i) { h[4][i] = U(gen) / n; h[5][i] = U(gen) / n; h[6][i] = U(gen) / n; h[3][i] = 1.5f * std::pow((float)n, -1.f / 3.f); }
  float* d[7];
  for (int c = 0; c < 7; ++c) { CHECK(cudaMalloc(&d[c], n * 4)); CHECK(cudaMemcpy(d[c], h[c].data(), n * 4, cudaMemcpyHostToDevice)); }
  const int64_t npad = padded_sources(n);
  float4 *pk, *pk2;
  CHECK(cudaMalloc(&pk, npad * 32)); CHECK(cudaMalloc(&pk2, npad * 32));
  pp_pack_kernel<<<(npad + 255) / 256, 256>>>(n, npad, d[0], d[1], d[2], d[3], d[4], d[5], d[6], pk);
  // KBENCH_CORE=1|2|3 (Rosenhead-Moore | exponential | Vatistas): the cubins hold ppc_kernel<core, ..>, the stream carries that core's radius lane
  const int kcore = getenv("KBENCH_CORE") ? atoi(getenv("KBENCH_CORE")) : 0;
  if (kcore) ppc_pack2_kernel<<<(npad / 2 + 255) / 256, 256>>>(kcore, n, npad, d[0], d[1], d[2], d[3], d[4], d[5], d[6], pk2);
  else pp_pack2_kernel<<<(npad / 2 + 255) / 256, 256>>>(n, npad, d[0], d[1], d[2], d[3], d[4], d[5], d[6], pk2);
  float* out; CHECK(cudaMalloc(&out, (size_t)n * 12 * 4));
  double* partial; CHECK(cudaMalloc(&partial, (size_t)n * 12 * 8 * 4));  // scalar kernel: up to 4 source slices
  // packed kernel: persistent CTAs, 3 per SM (capi.cu: pp_shape), 2 workspace slots each
  const int resident = prop.multiProcessorCount * 3;
  double* ppwork; CHECK(cudaMalloc(&ppwork, (size_t)resident * kPPSlots * 12 * 256 * 8));
  uint32_t* range; CHECK(cudaMalloc(&range, 16));
  CHECK(cudaMemset(range, 0, 16));
  pp_scan_kernel<<<prop.multiProcessorCount * 4, 256>>>(npad, pk2, n, d[3], range);
  const bool no_uniform = getenv("KBENCH_NO_UNIFORM") != nullptr;   // force the general-radius path
  CHECK(cudaDeviceSynchronize());

  // host reference on 8 targets
  const int nchk = 8;
  std::vector<double> ref(nchk * 12);
  double umax = 0, gmax = 0;
  for (int c = 0; c < nchk; ++c) {
    const int i = (int)((long long)c * n / nchk);
    host_ref(n, h[0].data(), h[1].data(), h[2].data(), h[3].data(), h[4].data(), h[5].data(), h[6].data(), h[0][i], h[1][i], h[2][i], h[3][i], &ref[c * 12]);
    for (int k = 0; k < 3; ++k) umax = std::fmax(umax, std::fabs(ref[c * 12 + k]));
    for (int k = 3; k < 12; ++k) gmax = std::fmax(gmax, std::fabs(ref[c * 12 + k]));
  }

  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto report = [&](const char* name, int T, int BLOCK, bool grad, int nsplit, int grid, float best, bool fnv) {
    std::vector<float> o((size_t)n * 12);
    CHECK(cudaMemcpy(o.data(), out, (size_t)n * 12 * 4, cudaMemcpyDeviceToHost));
    double eu = 0, eg = 0;
    for (int c = 0; c < nchk; ++c) {
      const int i = (int)((long long)c * n / nchk);
      for (int k = 0; k < 3; ++k) eu = std::fmax(eu, std::fabs(o[(size_t)k * n + i] - ref[c * 12 + k]));
      if (grad) for (int k = 3; k < 12; ++k) eg = std::fmax(eg, std::fabs(o[(size_t)k * n + i] - ref[c * 12 + k]));
    }
    const double ips = (double)n * n / (best * 1e-3);
    printf("%-36s T=%d B=%3d split=%2d grid=%6d  %8.3f ms  %.3e int/s  %6.2f TFLOP/s@%d  err u %.2e g %.2e", name, T, BLOCK, nsplit,
           grid, best, ips, ips * (grad ? 70 : 33) * 1e-12, grad ? 70 : 33, eu / umax, eg / gmax);
    if (fnv) {   // checksum over ALL outputs: a patched cubin must reproduce the linked kernel bit for bit
      unsigned long long h = 1469598103934665603ull;
      const size_t cnt = grad ? o.size() : (size_t)n * 3;
      for (size_t q = 0; q < cnt; ++q) { unsigned v; memcpy(&v, &o[q], 4); h = (h ^ v) * 1099511628211ull; }
      printf("  fnv %016llx", h);
    }
    printf("\n");
  };
  // scalar-FFMA kernel: one target block per CTA, gridDim.y source slices (pp_scalar.cuh)
  auto run_scalar = [&](const char* name, auto kern, int T, int BLOCK, bool grad, int nsplit) {
    PPScalarArgs a{};
    a.src = pk; a.ntiles = (int)(npad / kTile); a.nsplit = nsplit; a.nt = n;
    a.tx = d[0]; a.ty = d[1]; a.tz = d[2]; a.tr = d[3];
    a.tu = out; a.tv = out + n; a.tw = out + 2 * (size_t)n; a.tug = out + 3 * (size_t)n; a.tug_stride = n;
    a.partial = partial; a.sign = 1.0f;
    dim3 grid((n + BLOCK * T - 1) / (BLOCK * T), nsplit);
    float best = 1e30f;
    for (int r = 0; r < reps + 1; ++r) {
      CHECK(cudaMemset(out, 0, (size_t)n * 12 * 4));
      cudaEventRecord(e0);
      kern<<<grid, BLOCK>>>(a);
      if (nsplit > 1) pp_finish_kernel<<<(n + 255) / 256, 256>>>(grad ? 12 : 3, nsplit, n, partial, a.tu, a.tv, a.tw, a.tug, n, 1.0f);
      cudaEventRecord(e1);
      CHECK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (r > 0 && ms < best) best = ms;
    }
    report(name, T, BLOCK, grad, nsplit, grid.x * grid.y, best, false);
  };
  // packed kernel: persistent CTAs over the stream-K partition + fix-up of the shared target blocks (what capi.cu launches).
  // `fn` != nullptr: the copy of the kernel in a cubin file, through the driver API.
  auto run_packed = [&](const char* name, auto kern, CUfunction fn, int T, int BLOCK, bool grad, int per_sm) {
    PPArgs a{};
    a.src = pk2; a.ntiles = (int)(npad / kTile); a.nt = n;
    a.nblocks = (n + BLOCK * T - 1) / (BLOCK * T);
    a.tx = d[0]; a.ty = d[1]; a.tz = d[2]; a.tr = d[3];
    a.tu = out; a.tv = out + n; a.tw = out + 2 * (size_t)n; a.tug = grad ? out + 3 * (size_t)n : nullptr; a.tug_stride = n;
    a.partial = ppwork; a.sign = 1.0f;
    a.radius_range = no_uniform ? nullptr : range;
    a.slots = prop.multiProcessorCount * per_sm;
    const PPPlan plan = pp_make_plan(a.slots, a.nblocks, a.ntiles);
    const int grid = plan.P;
    if ((int64_t)grid * BLOCK * T * (grad ? 12 : 3) > (int64_t)resident * 12 * 256) { printf("%s: workspace too small for this shape\n", name); return; }
    float best = 1e30f;
    for (int r = 0; r < reps + 1; ++r) {
      CHECK(cudaMemset(out, 0, (size_t)n * 12 * 4));
      cudaEventRecord(e0);
      if (fn) {
        void* params[] = {&a};
        if (cuLaunchKernel(fn, grid, 1, 1, BLOCK, 1, 1, 0, 0, params, nullptr) != CUDA_SUCCESS) { printf("launch failed\n"); return; }
      } else {
        kern<<<grid, BLOCK>>>(a);
      }
      if (plan.Pt > 1) pp_fixup_kernel<<<plan.Pt - 1, 256>>>(grad ? 12 : 3, BLOCK * T, plan, n, ppwork, a.tu, a.tv, a.tw, a.tug, n, 1.0f, nullptr, 0);
      cudaEventRecord(e1);
      CHECK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (r > 0 && ms < best) best = ms;
    }
    report(name, T, BLOCK, grad, 1, grid, best, true);
#ifdef O3D_PP_ENDTIME
    if (!fn) {   // when did each CTA of the last launch end, and on which SM: do co-resident CTAs progress equally?
      std::vector<unsigned long long> te(2048); std::vector<unsigned> sm(2048);
      CHECK(cudaMemcpyFromSymbol(te.data(), pp_end_time, 2048 * 8)); CHECK(cudaMemcpyFromSymbol(sm.data(), pp_end_smid, 2048 * 4));
      unsigned long long tmax = 0; for (int c = 0; c < grid; ++c) tmax = std::max(tmax, te[c]);
      std::vector<double> lag(grid); for (int c = 0; c < grid; ++c) lag[c] = (double)(tmax - te[c]) * 1e-6;   // ms before the last CTA
      std::vector<double> s2 = lag; std::sort(s2.begin(), s2.end());
      printf("   CTA end times, ms before the last one: min %.3f  p25 %.3f  median %.3f  p75 %.3f  max %.3f\n", s2[0], s2[grid / 4], s2[grid / 2], s2[3 * grid / 4], s2[grid - 1]);
      printf("   SM 0..3 CTAs (lag ms):");
      for (unsigned q = 0; q < 4; ++q) { printf(" ["); for (int c = 0; c < grid; ++c) if (sm[c] == q) printf(" %.2f", lag[c]); printf(" ]"); }
      printf("\n");
      if (BLOCK == 384) {   // per-tile skew of the 12 warps of CTA 0: when does each warp finish a tile's arithmetic, relative to the first
        long long tc[12][64];
        CHECK(cudaMemcpyFromSymbol(tc, pp_tile_clock, sizeof tc));
        for (int k = 8; k < 14; ++k) {
          long long t0 = tc[0][k]; for (int wp = 0; wp < 12; ++wp) t0 = std::min(t0, tc[wp][k]);
          printf("   tile %2d: length %lld clk; warps done at (clk after the first, scheduler = warp %% 4):", k, tc[0][k] - tc[0][k - 1]);
          for (int q = 0; q < 4; ++q) { printf("  [s%d", q); for (int wp = q; wp < 12; wp += 4) printf(" %lld", tc[wp][k] - t0); printf("]"); }
          printf("\n");
        }
      }
    }
#endif
  };
#define RUN_S(T, B, G, SPLIT) run_scalar("scalar" #G, pp_kernel<T, G, B>, T, B, G, SPLIT)
#define RUN_P(T, B, G, PER_SM) run_packed("packed" #G " stage" O3D_STR(O3D_PP_STAGE), pp2_kernel<T, G, B>, nullptr, T, B, G, PER_SM)
  // KBENCH_CUBIN=a.cubin[:b.cubin...]: time the two product instantiations of pp2_kernel loaded from cubin files
  // (tools/sass_patch.py output) through the driver API, next to the copies linked into this binary.
  if (const char* list = getenv("KBENCH_CUBIN")) {
    std::string all(list);
    size_t pos = 0;
    while (pos <= all.size()) {
      size_t e = all.find(':', pos);
      if (e == std::string::npos) e = all.size();
      const std::string path = all.substr(pos, e - pos);
      pos = e + 1;
      if (path.empty()) continue;
      CUmodule mod; CUfunction fn;
      if (cuModuleLoad(&mod, path.c_str()) != CUDA_SUCCESS) { printf("cannot load %s\n", path.c_str()); continue; }
      char sg[96], sv[96];
      if (kcore) {
        snprintf(sg, sizeof sg, "_ZN3o3d10ppc_kernelILi%dELi2ELb1ELi" O3D_STR(O3D_PP_BLOCK) "EEEvNS_6PPArgsE", kcore);
        snprintf(sv, sizeof sv, "_ZN3o3d10ppc_kernelILi%dELi4ELb0ELi" O3D_STR(O3D_PP_BLOCK) "EEEvNS_6PPArgsE", kcore);
      } else {
        snprintf(sg, sizeof sg, "_ZN3o3d10pp2_kernelILi2ELb1ELi" O3D_STR(O3D_PP_BLOCK) "EEEvNS_6PPArgsE");
        snprintf(sv, sizeof sv, "_ZN3o3d10pp2_kernelILi4ELb0ELi" O3D_STR(O3D_PP_BLOCK) "EEEvNS_6PPArgsE");
      }
      struct { const char* sym; int T; bool grad; } kinds[] = {{sg, 2, true}, {sv, 4, false}};
      for (auto& kd : kinds) {
        if (cuModuleGetFunction(&fn, mod, kd.sym) != CUDA_SUCCESS) continue;
        const std::string label = "cubin " + path + (kd.grad ? " velgrad" : " vel");
        run_packed(label.c_str(), pp2_kernel<2, true, kPPBlock>, fn, kd.T, kPPBlock, kd.grad, kPPResident);
      }
    }
    if (getenv("KBENCH_CUBIN_ONLY")) return 0;
  }
  if (getenv("KBENCH_PRODUCT_ONLY")) {     // the two product shapes only (staging-variant builds: make kbench_stage)
    RUN_P(2, kPPBlock, true, kPPResident);
    RUN_P(4, kPPBlock, false, kPPResident);
#ifdef O3D_PP_ENDTIME
    RUN_P(2, 128, true, 3);     // three CTAs per SM: the oldest finishes first (strict age priority between co-resident CTAs)
    RUN_P(4, 128, false, 3);
#endif
    return 0;
  }
  RUN_S(1, 256, true, 1);
  RUN_S(2, 256, true, 1);
  RUN_S(2, 128, true, 1);
  RUN_S(4, 256, true, 1);
  RUN_S(4, 128, true, 1);
  RUN_S(3, 128, true, 1);
  RUN_S(3, 256, true, 1);
  RUN_S(2, 256, true, 4);
  RUN_P(2, 128, true, 3);
  RUN_P(2, 384, true, 1);
  RUN_S(4, 256, false, 1);
  RUN_S(8, 128, false, 1);
  RUN_P(4, 128, false, 3);
  RUN_P(4, 384, false, 1);
  return 0;
}

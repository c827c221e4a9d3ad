// capi.cu - the C ABI (include/o3d_cuda.h) over the sm_100a kernels. Host-side plumbing only: device
// selection, buffers that grow geometrically and live in the context, H2D/D2H staging, target partition
// across the context's GPUs, launch-shape selection and event timing. No arithmetic of the path happens
// on the host and there is no CPU fallback: every entry point either runs the CUDA kernels or fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/o3d_cuda.h"
#include "biot_panel.cuh"
#include "biot_pp.cuh"

using namespace o3d;

__global__ void fma_probe_kernel(float* out, float a, float b, int iters) {
  // 8 independent packed chains acc = acc * aa + bb: two 64-bit register reads per FFMA2 after operand
  // reuse, which is what the register file can feed at full FFMA2 rate (tools/sass_rf_model.py)
  float2 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
  const float2 aa = make_float2(a, a * 1.0001f), bb = make_float2(b, b * 0.999f);
#pragma unroll 4
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = __ffma2_rn(acc[i], aa, bb);
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  // grow-only, geometric: particle counts change every step (VRM / split / merge), SURVEY.md 8b
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    size_t want = std::max(bytes, cap + cap / 2);
    want = (want + 255) & ~size_t(255);
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {  // retry with the exact size before giving up
      cudaGetLastError();
      want = (bytes + 255) & ~size_t(255);
      e = cudaMalloc(&p, want);
    }
    cap = e == cudaSuccess ? want : 0;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct Device {
  int id = 0;
  int sm_count = 0;
  int clock_khz = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // h2d start, compute start, compute end, d2h end
  cudaEvent_t evk[2] = {nullptr, nullptr};                   // around the dominant kernel of a *_dev call
  bool profile = false;
  DevBuf src, packed, targ, out, work, geom, panels, tpanels, cnt, rng;
  unsigned long long counts[2] = {0, 0};  // leaves, splits of the last panel call on this device
  // result of the last call on this device
  float kernel_ms = 0, h2d_ms = 0, d2h_ms = 0;
  int launches = 0;
  cudaError_t status = cudaSuccess;
  const char* where = "";
};

}  // namespace

struct o3d_ctx {
  std::vector<Device> dev;
  std::string err;
  double kernel_ms = 0, h2d_ms = 0, d2h_ms = 0;
  int launches = 0;
};

namespace {

#define O3D_TRY(dv, call)                 \
  do {                                    \
    cudaError_t e_ = (call);              \
    if (e_ != cudaSuccess) {              \
      (dv).status = e_;                   \
      (dv).where = #call;                 \
      return false;                       \
    }                                     \
  } while (0)

int fail(o3d_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  return code;
}

int collect(o3d_ctx* c) {
  c->kernel_ms = c->h2d_ms = c->d2h_ms = 0;
  c->launches = 0;
  for (Device& d : c->dev) {
    if (d.status != cudaSuccess) {
      char buf[512];
      snprintf(buf, sizeof buf, "device %d: %s failed: %s", d.id, d.where, cudaGetErrorString(d.status));
      const int code = d.status == cudaErrorMemoryAllocation ? O3D_ERR_NOMEM : O3D_ERR_CUDA;
      d.status = cudaSuccess;
      cudaGetLastError();
      return fail(c, code, buf);
    }
    c->kernel_ms = std::max(c->kernel_ms, (double)d.kernel_ms);
    c->h2d_ms = std::max(c->h2d_ms, (double)d.h2d_ms);
    c->d2h_ms = std::max(c->d2h_ms, (double)d.d2h_ms);
    c->launches += d.launches;
  }
  c->err.clear();
  return O3D_OK;
}

// contiguous block partition of n items over the context's devices
void partition(int64_t n, int ndev, int k, int64_t* b, int64_t* e) {
  const int64_t per = (n + ndev - 1) / ndev;
  *b = std::min(n, per * k);
  *e = std::min(n, per * (k + 1));
}

template <class F> void for_each_device(o3d_ctx* c, F&& f) {
  const int nd = (int)c->dev.size();
  if (nd == 1) {
    f(0);
    return;
  }
  std::vector<std::thread> th;
  th.reserve(nd);
  for (int k = 0; k < nd; ++k) th.emplace_back([&f, k] { f(k); });
  for (auto& t : th) t.join();
}

bool finish_timing(Device& d) {
  O3D_TRY(d, cudaStreamSynchronize(d.stream));
  O3D_TRY(d, cudaEventElapsedTime(&d.h2d_ms, d.ev[0], d.ev[1]));
  O3D_TRY(d, cudaEventElapsedTime(&d.kernel_ms, d.ev[1], d.ev[2]));
  O3D_TRY(d, cudaEventElapsedTime(&d.d2h_ms, d.ev[2], d.ev[3]));
  return true;
}

// ---- launch-shape selection for particles -> points ------------------------------------------------
// Product configuration (profiles/r01_*: packed FFMA2 kernel, 128-thread CTAs): vel+grad keeps 2 targets
// per thread, velocity-only 4. When the target count cannot fill the GPU the source range is split over
// gridDim.y and the per-slice FP64 partial sums meet in a workspace.
struct PPShape {
  int nsplit;
  dim3 grid;
  size_t work_bytes;
};
constexpr int kPPBlock = 128;
constexpr int kPPTgrad = 2;
constexpr int kPPTvel = 4;

// CTAs resident per SM (register-limited: 162 / 128 regs x 128 threads) for the two product kernels
constexpr int kPPResidentGrad = 3;
constexpr int kPPResidentVel = 4;

PPShape pp_shape(const Device& d, int64_t ntiles, int64_t nt, bool grad) {
  const int per_cta = kPPBlock * (grad ? kPPTgrad : kPPTvel);
  const int64_t gx = (nt + per_cta - 1) / per_cta;
  const int64_t slots = (int64_t)d.sm_count * (grad ? kPPResidentGrad : kPPResidentVel);
  // Pick the source split that wastes the least of the last wave: efficiency = CTAs / (waves * slots).
  // Large target counts (>= 16 waves) never split; tiny ones split until the GPU is covered twice.
  int64_t best = 1;
  if (gx < 16 * slots) {
    double best_eff = 0.0;
    const int64_t max_split = std::min<int64_t>(ntiles, 64);
    for (int64_t sp = 1; sp <= max_split; ++sp) {
      const int64_t ctas = gx * sp;
      const int64_t waves = (ctas + slots - 1) / slots;
      double eff = (double)ctas / (double)(waves * slots);
      if (ctas >= 2 * slots) eff += 1e-3 * (1.0 / sp);  // among equals prefer fewer slices
      if (eff > best_eff + 0.02) best_eff = eff, best = sp;
    }
  }
  PPShape s;
  s.nsplit = (int)best;
  s.grid = dim3((unsigned)gx, (unsigned)best, 1);
  s.work_bytes = best > 1 ? (size_t)best * (grad ? 12 : 3) * nt * sizeof(double) : 0;
  return s;
}

bool launch_pp(Device& d, cudaStream_t st, int64_t nrec, const float4* packed, int64_t nt, const float* tx,
               const float* ty, const float* tz, const float* tr, float* tu, float* tv, float* tw, float* tug,
               int64_t tug_stride, double* workspace) {
  const bool grad = tug != nullptr;
  const int64_t ntiles = nrec / kTile;
  const PPShape s = pp_shape(d, ntiles, nt, grad);
  PPArgs a{};
  a.src = packed;
  a.ntiles = (int)ntiles;
  a.nsplit = s.nsplit;
  a.nt = nt;
  a.tx = tx; a.ty = ty; a.tz = tz; a.tr = tr;
  a.tu = tu; a.tv = tv; a.tw = tw;
  a.tug = tug;
  a.tug_stride = tug_stride;
  a.sign = 1.0f;
  a.partial = nullptr;
  if (s.nsplit > 1) {
    if (!workspace) {
      O3D_TRY(d, d.work.ensure(s.work_bytes));
      workspace = d.work.as<double>();
    }
    a.partial = workspace;
  }
  // radius ranges for the uniform-radius fast path (one pass over the records' r^2 lane and the target radii)
  static const uint32_t kRangeInit[4] = {0xffffffffu, 0u, 0xffffffffu, 0u};
  O3D_TRY(d, d.rng.ensure(sizeof kRangeInit));
  O3D_TRY(d, cudaMemcpyAsync(d.rng.p, kRangeInit, sizeof kRangeInit, cudaMemcpyHostToDevice, st));
  pp_scan_kernel<<<d.sm_count * 4, 256, 0, st>>>(nrec, packed, nt, tr, d.rng.as<uint32_t>());
  O3D_TRY(d, cudaGetLastError());
  d.launches += 1;
  a.radius_range = d.rng.as<uint32_t>();
  if (d.profile) O3D_TRY(d, cudaEventRecord(d.evk[0], st));
  if (grad)
    pp2_kernel<kPPTgrad, true, kPPBlock><<<s.grid, kPPBlock, 0, st>>>(a);
  else
    pp2_kernel<kPPTvel, false, kPPBlock><<<s.grid, kPPBlock, 0, st>>>(a);
  O3D_TRY(d, cudaGetLastError());
  if (d.profile) O3D_TRY(d, cudaEventRecord(d.evk[1], st));
  d.launches += 1;
  if (s.nsplit > 1) {
    pp_finish_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, st>>>(grad ? 12 : 3, s.nsplit, nt, workspace, tu, tv, tw,
                                                                  tug, tug_stride, 1.0f);
    O3D_TRY(d, cudaGetLastError());
    d.launches += 1;
  }
  return true;
}

bool launch_pack(Device& d, cudaStream_t st, int64_t ns, const float* sx, const float* sy, const float* sz,
                 const float* sr, const float* wx, const float* wy, const float* wz, float4* packed, int64_t nrec = 0) {
  const int64_t npad = nrec > 0 ? nrec : padded_sources(ns);
  pp_pack2_kernel<<<(unsigned)((npad / 2 + 255) / 256), 256, 0, st>>>(ns, npad, sx, sy, sz, sr, wx, wy, wz, packed);
  O3D_TRY(d, cudaGetLastError());
  d.launches += 1;
  return true;
}

// Dependent-free packed-FMA loop (fma.rn.f32x2, three distinct register pairs per instruction - the operand
// shape of the Biot-Savart inner loop): what the FP32 pipe sustains on this GPU at the clocks it actually
// holds. bench.py reports it beside the nominal SMs x 128 x 2 x f_max peak.
bool run_fma_probe(Device& d, double* tflops, double* ms_out) {
  const int threads = 256, blocks = d.sm_count * 8, iters = 1 << 15;
  O3D_TRY(d, cudaSetDevice(d.id));
  O3D_TRY(d, d.work.ensure((size_t)blocks * threads * sizeof(float)));
  cudaStream_t st = d.stream;
  fma_probe_kernel<<<blocks, threads, 0, st>>>(d.work.as<float>(), 0.999f, 1e-3f, 256);  // warm-up, clocks ramp
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    O3D_TRY(d, cudaEventRecord(d.ev[0], st));
    fma_probe_kernel<<<blocks, threads, 0, st>>>(d.work.as<float>(), 0.999f, 1e-3f, iters);
    O3D_TRY(d, cudaEventRecord(d.ev[1], st));
    O3D_TRY(d, cudaStreamSynchronize(st));
    float ms = 0;
    O3D_TRY(d, cudaEventElapsedTime(&ms, d.ev[0], d.ev[1]));
    best = std::min(best, ms);
    d.launches += 1;
  }
  O3D_TRY(d, cudaGetLastError());
  const double flops = (double)blocks * threads * (double)iters * 16.0 * 2.0 * 2.0;  // 16 FFMA2 per iteration
  *tflops = flops / (best * 1e-3) * 1e-12;
  *ms_out = best;
  return true;
}

bool check_counts(o3d_ctx* c, int64_t a, int64_t b) {
  return c && a >= 0 && b >= 0 && a < (int64_t(1) << 31) && b < (int64_t(1) << 31);
}

}  // namespace

extern "C" {

int o3d_cuda_abi_version(void) { return O3D_CUDA_ABI_VERSION; }

int o3d_cuda_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int o3d_cuda_create(o3d_ctx** out, int ndev, const int* devices) {
  if (!out || ndev < 1) return O3D_ERR_INVALID;
  *out = nullptr;
  const int have = o3d_cuda_device_count();
  if (have < 1 || ndev > have) return O3D_ERR_NODEVICE;
  o3d_ctx* c = new o3d_ctx();
  c->dev.resize(ndev);
  for (int k = 0; k < ndev; ++k) {
    Device& d = c->dev[k];
    d.id = devices ? devices[k] : k;
    cudaDeviceProp prop;
    if (d.id < 0 || d.id >= have || cudaGetDeviceProperties(&prop, d.id) != cudaSuccess || prop.major < 10) {
      o3d_cuda_destroy(c);  // built for sm_100a only
      return O3D_ERR_NODEVICE;
    }
    d.sm_count = prop.multiProcessorCount;
    cudaDeviceGetAttribute(&d.clock_khz, cudaDevAttrClockRate, d.id);
    bool ok = cudaSetDevice(d.id) == cudaSuccess &&
              cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int e = 0; ok && e < 4; ++e) ok = cudaEventCreate(&d.ev[e]) == cudaSuccess;
    for (int e = 0; ok && e < 2; ++e) ok = cudaEventCreate(&d.evk[e]) == cudaSuccess;
    if (!ok) {
      o3d_cuda_destroy(c);
      return O3D_ERR_CUDA;
    }
  }
  *out = c;
  return O3D_OK;
}

void o3d_cuda_destroy(o3d_ctx* c) {
  if (!c) return;
  for (Device& d : c->dev) {
    cudaSetDevice(d.id);
    for (DevBuf* b : {&d.src, &d.packed, &d.targ, &d.out, &d.work, &d.geom, &d.panels, &d.tpanels, &d.cnt, &d.rng}) b->release();
    for (cudaEvent_t e : d.ev)
      if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : d.evk)
      if (e) cudaEventDestroy(e);
    if (d.stream) cudaStreamDestroy(d.stream);
  }
  delete c;
}

const char* o3d_cuda_last_error(const o3d_ctx* c) { return c ? c->err.c_str() : "null context"; }
int o3d_cuda_num_devices(const o3d_ctx* c) { return c ? (int)c->dev.size() : 0; }

int o3d_cuda_device_props(const o3d_ctx* c, int k, int* sm_count, int* clock_khz, double* fp32_peak) {
  if (!c || k < 0 || k >= (int)c->dev.size()) return O3D_ERR_INVALID;
  const Device& d = c->dev[k];
  if (sm_count) *sm_count = d.sm_count;
  if (clock_khz) *clock_khz = d.clock_khz;
  if (fp32_peak) *fp32_peak = (double)d.sm_count * 128.0 * 2.0 * (double)d.clock_khz * 1e3;
  return O3D_OK;
}

int o3d_cuda_last_timing(const o3d_ctx* c, double* kernel_ms, double* h2d_ms, double* d2h_ms, int* launches) {
  if (!c) return O3D_ERR_INVALID;
  if (kernel_ms) *kernel_ms = c->kernel_ms;
  if (h2d_ms) *h2d_ms = c->h2d_ms;
  if (d2h_ms) *d2h_ms = c->d2h_ms;
  if (launches) *launches = c->launches;
  return O3D_OK;
}

// ---------------------------------------------------------------------------------------------------
int o3d_cuda_pts_on_pts(o3d_ctx* c, int64_t ns, const float* sx, const float* sy, const float* sz, const float* sr,
                        const float* ssx, const float* ssy, const float* ssz, int64_t nt, const float* tx,
                        const float* ty, const float* tz, const float* tr, float* tu, float* tv, float* tw,
                        float* const* tug, double* flops_out) {
  if (!check_counts(c, ns, nt)) return fail(c, O3D_ERR_INVALID, "pts_on_pts: bad context or counts");
  if (ns > 0 && (!sx || !sy || !sz || !sr || !ssx || !ssy || !ssz))
    return fail(c, O3D_ERR_INVALID, "pts_on_pts: NULL source array");
  if (nt > 0 && (!tx || !ty || !tz || !tu || !tv || !tw)) return fail(c, O3D_ERR_INVALID, "pts_on_pts: NULL target array");
  if (tug)
    for (int k = 0; k < 9; ++k)
      if (nt > 0 && !tug[k]) return fail(c, O3D_ERR_INVALID, "pts_on_pts: NULL gradient array");
  const bool grad = tug != nullptr;
  if (flops_out) {
    // src/Influence.h:310 (0pg: 68), :366 (0p: 31), :475 (0bg: 70), :534 (0b: 33)
    const double per = grad ? (tr ? 70.0 : 68.0) : (tr ? 33.0 : 31.0);
    *flops_out = (double)nt * ((grad ? 12.0 : 3.0) + per * (double)ns);
  }
  for (Device& d : c->dev) d.kernel_ms = d.h2d_ms = d.d2h_ms = 0, d.launches = 0;
  if (ns == 0 || nt == 0) return collect(c);

  const int ndev = (int)c->dev.size();
  for_each_device(c, [&](int k) {
    Device& d = c->dev[k];
    int64_t t0, t1;
    partition(nt, ndev, k, &t0, &t1);
    const int64_t n = t1 - t0;
    if (n == 0) return true;
    O3D_TRY(d, cudaSetDevice(d.id));
    const int64_t nrec = padded_sources(ns);
    const int nout = grad ? 12 : 3;
    O3D_TRY(d, d.src.ensure((size_t)7 * ns * 4));
    O3D_TRY(d, d.packed.ensure((size_t)nrec * 32));
    O3D_TRY(d, d.targ.ensure((size_t)4 * n * 4));
    O3D_TRY(d, d.out.ensure((size_t)nout * n * 4));
    cudaStream_t st = d.stream;
    float* ds = d.src.as<float>();
    float* dt = d.targ.as<float>();
    float* dout = d.out.as<float>();
    O3D_TRY(d, cudaEventRecord(d.ev[0], st));
    const float* hs[7] = {sx, sy, sz, sr, ssx, ssy, ssz};
    for (int a = 0; a < 7; ++a) O3D_TRY(d, cudaMemcpyAsync(ds + (size_t)a * ns, hs[a], (size_t)ns * 4, cudaMemcpyHostToDevice, st));
    const float* ht[4] = {tx, ty, tz, tr};
    for (int a = 0; a < 4; ++a)
      if (ht[a]) O3D_TRY(d, cudaMemcpyAsync(dt + (size_t)a * n, ht[a] + t0, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    float* ho[12] = {tu, tv, tw};
    for (int a = 0; a < 9; ++a) ho[3 + a] = grad ? tug[a] : nullptr;
    for (int a = 0; a < nout; ++a) O3D_TRY(d, cudaMemcpyAsync(dout + (size_t)a * n, ho[a] + t0, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    O3D_TRY(d, cudaEventRecord(d.ev[1], st));
    if (!launch_pack(d, st, ns, ds, ds + ns, ds + 2 * ns, ds + 3 * ns, ds + 4 * ns, ds + 5 * ns, ds + 6 * ns, d.packed.as<float4>()))
      return false;
    if (!launch_pp(d, st, nrec, d.packed.as<float4>(), n, dt, dt + n, dt + 2 * n, tr ? dt + 3 * n : nullptr, dout, dout + n,
                   dout + 2 * n, grad ? dout + 3 * n : nullptr, n, nullptr))
      return false;
    O3D_TRY(d, cudaEventRecord(d.ev[2], st));
    for (int a = 0; a < nout; ++a) O3D_TRY(d, cudaMemcpyAsync(ho[a] + t0, dout + (size_t)a * n, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    O3D_TRY(d, cudaEventRecord(d.ev[3], st));
    return finish_timing(d);
  });
  return collect(c);
}

// ---------------------------------------------------------------------------------------------------
// Shared by the three panel entry points: upload nodes + connectivity (+ strengths) and build the packed
// panel records on device `d` into `dst`.
namespace {
bool upload_panels(Device& d, cudaStream_t st, DevBuf& dst, int64_t nn, const float* nx, const float* ny,
                   const float* nz, int64_t np, const uint32_t* idx, const float* tsx, const float* tsy,
                   const float* tsz, const float* area, const float* sss) {
  const int64_t npad = padded_panels(np);
  // geom block: nodes 3*nn | idx 3*np | ts 3*np | area np | sss np
  const size_t words = (size_t)3 * nn + (size_t)8 * np;
  O3D_TRY(d, d.geom.ensure(words * 4));
  O3D_TRY(d, dst.ensure((size_t)npad * kPanRec * sizeof(float4)));
  float* g = d.geom.as<float>();
  float* gnx = g; float* gny = g + nn; float* gnz = g + 2 * nn;
  uint32_t* gidx = reinterpret_cast<uint32_t*>(g + 3 * nn);
  float* gts = g + 3 * nn + 3 * np;
  float* garea = gts + 3 * np;
  float* gsss = garea + np;
  O3D_TRY(d, cudaMemcpyAsync(gnx, nx, (size_t)nn * 4, cudaMemcpyHostToDevice, st));
  O3D_TRY(d, cudaMemcpyAsync(gny, ny, (size_t)nn * 4, cudaMemcpyHostToDevice, st));
  O3D_TRY(d, cudaMemcpyAsync(gnz, nz, (size_t)nn * 4, cudaMemcpyHostToDevice, st));
  O3D_TRY(d, cudaMemcpyAsync(gidx, idx, (size_t)3 * np * 4, cudaMemcpyHostToDevice, st));
  const float* hts[3] = {tsx, tsy, tsz};
  for (int a = 0; a < 3; ++a)
    if (hts[a]) O3D_TRY(d, cudaMemcpyAsync(gts + (size_t)a * np, hts[a], (size_t)np * 4, cudaMemcpyHostToDevice, st));
  O3D_TRY(d, cudaMemcpyAsync(garea, area, (size_t)np * 4, cudaMemcpyHostToDevice, st));
  if (sss) O3D_TRY(d, cudaMemcpyAsync(gsss, sss, (size_t)np * 4, cudaMemcpyHostToDevice, st));
  pan_pack_kernel<<<(unsigned)((npad + 127) / 128), 128, 0, st>>>(np, npad, gnx, gny, gnz, gidx, tsx ? gts : nullptr,
                                                                 tsy ? gts + np : nullptr, tsz ? gts + 2 * np : nullptr,
                                                                 garea, sss ? gsss : nullptr, dst.as<float4>());
  O3D_TRY(d, cudaGetLastError());
  d.launches += 1;
  return true;
}

bool zero_counts(Device& d, cudaStream_t st) {
  O3D_TRY(d, d.cnt.ensure(2 * sizeof(unsigned long long)));
  O3D_TRY(d, cudaMemsetAsync(d.cnt.p, 0, 2 * sizeof(unsigned long long), st));
  return true;
}
bool fetch_counts(Device& d, cudaStream_t st) {
  O3D_TRY(d, cudaMemcpyAsync(d.counts, d.cnt.p, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  return true;
}

// gridDim.y split of the source tiles when the target axis alone cannot fill the GPU
int pan_nsplit(const Device& d, int64_t gx, int64_t ntiles) {
  const int64_t fill = (int64_t)d.sm_count * 8;
  int64_t n = 1;
  if (gx < fill) n = std::min<int64_t>(std::min<int64_t>(ntiles, (fill + gx - 1) / gx), 4096);
  return (int)std::max<int64_t>(n, 1);
}

bool valid_indices(const uint32_t* idx, int64_t np, int64_t nn) {
  for (int64_t k = 0; k < 3 * np; ++k)
    if ((int64_t)idx[k] >= nn) return false;
  return true;
}
}  // namespace

int o3d_cuda_pan_on_pts(o3d_ctx* c, int64_t nn, const float* nx, const float* ny, const float* nz, int64_t np,
                        const uint32_t* idx, const float* tsx, const float* tsy, const float* tsz, const float* area,
                        const float* sss, int64_t nt, const float* tx, const float* ty, const float* tz, float* tu,
                        float* tv, float* tw, float* const* tug, double* flops_out) {
  if (!check_counts(c, np, nt) || nn < 0 || nn >= (int64_t(1) << 31)) return fail(c, O3D_ERR_INVALID, "pan_on_pts: bad context or counts");
  if (np > 0 && (!nx || !ny || !nz || !idx || !tsx || !tsy || !tsz || !area))
    return fail(c, O3D_ERR_INVALID, "pan_on_pts: NULL panel array");
  if (nt > 0 && (!tx || !ty || !tz || !tu || !tv || !tw)) return fail(c, O3D_ERR_INVALID, "pan_on_pts: NULL target array");
  if (tug)
    for (int k = 0; k < 9; ++k)
      if (nt > 0 && !tug[k]) return fail(c, O3D_ERR_INVALID, "pan_on_pts: NULL gradient array");
  if (np > 0 && !valid_indices(idx, np, nn)) return fail(c, O3D_ERR_INVALID, "pan_on_pts: node index out of range");
  const bool grad = tug != nullptr;
  for (Device& d : c->dev) d.kernel_ms = d.h2d_ms = d.d2h_ms = 0, d.launches = 0, d.counts[0] = d.counts[1] = 0;
  if (flops_out) *flops_out = (double)nt * (grad ? 12.0 : 3.0);
  if (np == 0 || nt == 0) return collect(c);

  const int ndev = (int)c->dev.size();
  const int64_t npad = padded_panels(np);
  for_each_device(c, [&](int k) {
    Device& d = c->dev[k];
    int64_t t0, t1;
    partition(nt, ndev, k, &t0, &t1);
    const int64_t n = t1 - t0;
    if (n == 0) return true;
    O3D_TRY(d, cudaSetDevice(d.id));
    cudaStream_t st = d.stream;
    const int nout = grad ? 12 : 3;
    O3D_TRY(d, d.targ.ensure((size_t)3 * n * 4));
    O3D_TRY(d, d.out.ensure((size_t)nout * n * 4));
    float* dt = d.targ.as<float>();
    float* dout = d.out.as<float>();
    O3D_TRY(d, cudaEventRecord(d.ev[0], st));
    const float* ht[3] = {tx, ty, tz};
    for (int a = 0; a < 3; ++a) O3D_TRY(d, cudaMemcpyAsync(dt + (size_t)a * n, ht[a] + t0, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    float* ho[12] = {tu, tv, tw};
    for (int a = 0; a < 9; ++a) ho[3 + a] = grad ? tug[a] : nullptr;
    for (int a = 0; a < nout; ++a) O3D_TRY(d, cudaMemcpyAsync(dout + (size_t)a * n, ho[a] + t0, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    if (!zero_counts(d, st)) return false;
    O3D_TRY(d, cudaEventRecord(d.ev[1], st));
    if (!upload_panels(d, st, d.panels, nn, nx, ny, nz, np, idx, tsx, tsy, tsz, area, sss)) return false;
    PanPtsArgs a{};
    a.pan = d.panels.as<float4>();
    a.ntiles = (int)(npad / kPanTile);
    a.nt = n;
    a.tx = dt; a.ty = dt + n; a.tz = dt + 2 * n;
    a.tu = dout; a.tv = dout + n; a.tw = dout + 2 * n;
    a.tug = grad ? dout + 3 * n : nullptr;
    a.tug_stride = n;
    a.counts = d.cnt.as<unsigned long long>();
    constexpr int B = 128;
    const int64_t gx = (n + B - 1) / B;
    a.nsplit = pan_nsplit(d, gx, a.ntiles);
    if (a.nsplit > 1) {
      O3D_TRY(d, d.work.ensure((size_t)a.nsplit * nout * n * sizeof(double)));
      a.partial = d.work.as<double>();
    }
    const dim3 grid((unsigned)gx, (unsigned)a.nsplit);
    if (grad) pan_pts_kernel<true, B><<<grid, B, 0, st>>>(a);
    else      pan_pts_kernel<false, B><<<grid, B, 0, st>>>(a);
    O3D_TRY(d, cudaGetLastError());
    d.launches += 1;
    if (a.nsplit > 1) {
      pp_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(nout, a.nsplit, n, a.partial, a.tu, a.tv, a.tw, a.tug, n, 1.0f);
      O3D_TRY(d, cudaGetLastError());
      d.launches += 1;
    }
    O3D_TRY(d, cudaEventRecord(d.ev[2], st));
    for (int a2 = 0; a2 < nout; ++a2) O3D_TRY(d, cudaMemcpyAsync(ho[a2] + t0, dout + (size_t)a2 * n, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    if (!fetch_counts(d, st)) return false;
    O3D_TRY(d, cudaEventRecord(d.ev[3], st));
    return finish_timing(d);
  });
  const int rc = collect(c);
  if (rc == O3D_OK && flops_out) {
    // the reference's own bookkeeping (src/Kernels.h:1036-1112): +4 once per pair, +20 per visited node,
    // + leaf kernel (37 | 93) per leaf, +23 per split; +3 | +12 per target (src/Influence.h:774,862)
    double leaves = 0, splits = 0;
    for (Device& d : c->dev) leaves += (double)d.counts[0], splits += (double)d.counts[1];
    leaves -= (double)nt * (double)(npad - np);  // padding records are always one leaf
    *flops_out = (double)nt * (double)np * 4.0 + (leaves + splits) * 20.0 + leaves * (grad ? 93.0 : 37.0) +
                 splits * 23.0 + (double)nt * (grad ? 12.0 : 3.0);
  }
  return rc;
}

int o3d_cuda_pts_on_pan(o3d_ctx* c, int64_t ns, const float* sx, const float* sy, const float* sz, const float* ssx,
                        const float* ssy, const float* ssz, int64_t nn, const float* nx, const float* ny,
                        const float* nz, int64_t np, const uint32_t* idx, const float* area, float* pu, float* pv,
                        float* pw, double* flops_out) {
  if (!check_counts(c, ns, np) || nn < 0 || nn >= (int64_t(1) << 31)) return fail(c, O3D_ERR_INVALID, "pts_on_pan: bad context or counts");
  if (ns > 0 && (!sx || !sy || !sz || !ssx || !ssy || !ssz)) return fail(c, O3D_ERR_INVALID, "pts_on_pan: NULL source array");
  if (np > 0 && (!nx || !ny || !nz || !idx || !area || !pu || !pv || !pw)) return fail(c, O3D_ERR_INVALID, "pts_on_pan: NULL panel array");
  if (np > 0 && !valid_indices(idx, np, nn)) return fail(c, O3D_ERR_INVALID, "pts_on_pan: node index out of range");
  for (Device& d : c->dev) d.kernel_ms = d.h2d_ms = d.d2h_ms = 0, d.launches = 0, d.counts[0] = d.counts[1] = 0;
  if (flops_out) *flops_out = 3.0 * (double)np;
  if (ns == 0 || np == 0) return collect(c);

  const int ndev = (int)c->dev.size();
  for_each_device(c, [&](int k) {
    Device& d = c->dev[k];
    int64_t p0, p1;
    partition(np, ndev, k, &p0, &p1);
    const int64_t n = p1 - p0;
    if (n == 0) return true;
    O3D_TRY(d, cudaSetDevice(d.id));
    cudaStream_t st = d.stream;
    const int64_t nrec = padded_sources(ns);
    O3D_TRY(d, d.src.ensure((size_t)6 * ns * 4));
    O3D_TRY(d, d.packed.ensure((size_t)nrec * 32));
    O3D_TRY(d, d.out.ensure((size_t)3 * n * 4));
    float* ds = d.src.as<float>();
    float* dout = d.out.as<float>();
    O3D_TRY(d, cudaEventRecord(d.ev[0], st));
    const float* hs[6] = {sx, sy, sz, ssx, ssy, ssz};
    for (int a = 0; a < 6; ++a) O3D_TRY(d, cudaMemcpyAsync(ds + (size_t)a * ns, hs[a], (size_t)ns * 4, cudaMemcpyHostToDevice, st));
    float* ho[3] = {pu, pv, pw};
    for (int a = 0; a < 3; ++a) O3D_TRY(d, cudaMemcpyAsync(dout + (size_t)a * n, ho[a] + p0, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    if (!zero_counts(d, st)) return false;
    O3D_TRY(d, cudaEventRecord(d.ev[1], st));
    // this device's slice of the target panels (connectivity offset by p0; nodes whole)
    if (!upload_panels(d, st, d.tpanels, nn, nx, ny, nz, n, idx + 3 * p0, nullptr, nullptr, nullptr, area + p0, nullptr)) return false;
    if (!launch_pack(d, st, ns, ds, ds + ns, ds + 2 * ns, nullptr, ds + 3 * ns, ds + 4 * ns, ds + 5 * ns, d.packed.as<float4>())) return false;
    PtsPanArgs a{};
    a.src = d.packed.as<float4>();
    a.ns = ns;
    a.ntiles = (int)(nrec / kTile);
    a.np = n;
    a.pan = d.tpanels.as<float4>();
    a.counts = d.cnt.as<unsigned long long>();
    constexpr int B = 64;
    const int64_t gx = (n + B - 1) / B;
    a.nsplit = pan_nsplit(d, gx, a.ntiles);
    O3D_TRY(d, d.work.ensure((size_t)a.nsplit * 3 * n * sizeof(double)));
    a.partial = d.work.as<double>();
    pts_pan_kernel<B><<<dim3((unsigned)gx, (unsigned)a.nsplit), B, 0, st>>>(a);
    O3D_TRY(d, cudaGetLastError());
    d.launches += 1;
    pp_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(3, a.nsplit, n, a.partial, dout, dout + n, dout + 2 * n, nullptr, n, -1.0f);
    O3D_TRY(d, cudaGetLastError());
    d.launches += 1;
    O3D_TRY(d, cudaEventRecord(d.ev[2], st));
    for (int a2 = 0; a2 < 3; ++a2) O3D_TRY(d, cudaMemcpyAsync(ho[a2] + p0, dout + (size_t)a2 * n, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    if (!fetch_counts(d, st)) return false;
    O3D_TRY(d, cudaEventRecord(d.ev[3], st));
    return finish_timing(d);
  });
  const int rc = collect(c);
  if (rc == O3D_OK && flops_out) {
    double leaves = 0, splits = 0;
    for (Device& d : c->dev) leaves += (double)d.counts[0], splits += (double)d.counts[1];
    *flops_out = (double)ns * (double)np * 4.0 + (leaves + splits) * 20.0 + leaves * 37.0 + splits * 23.0 + 3.0 * (double)np;
  }
  return rc;
}

int o3d_cuda_pan_on_pan_coeff(o3d_ctx* c, int64_t snn, const float* snx, const float* sny, const float* snz, int64_t nsp,
                              const uint32_t* sidx, const float* sb1, const float* sb2, const float* sarea, int64_t tnn,
                              const float* tnx, const float* tny, const float* tnz, int64_t ntp, const uint32_t* tidx,
                              const float* tb1, const float* tb2, const float* tnrm, const float* tarea, int self,
                              float* coeffs, double* flops_out) {
  if (!check_counts(c, nsp, ntp) || snn < 0 || tnn < 0 || snn >= (int64_t(1) << 31) || tnn >= (int64_t(1) << 31))
    return fail(c, O3D_ERR_INVALID, "pan_on_pan_coeff: bad context or counts");
  if (nsp > 0 && (!snx || !sny || !snz || !sidx || !sb1 || !sb2 || !sarea)) return fail(c, O3D_ERR_INVALID, "pan_on_pan_coeff: NULL source array");
  if (ntp > 0 && (!tnx || !tny || !tnz || !tidx || !tb1 || !tb2 || !tnrm || !tarea)) return fail(c, O3D_ERR_INVALID, "pan_on_pan_coeff: NULL target array");
  if (nsp > 0 && ntp > 0 && !coeffs) return fail(c, O3D_ERR_INVALID, "pan_on_pan_coeff: NULL output");
  if (self && nsp != ntp) return fail(c, O3D_ERR_INVALID, "pan_on_pan_coeff: self block must be square");
  if ((nsp > 0 && !valid_indices(sidx, nsp, snn)) || (ntp > 0 && !valid_indices(tidx, ntp, tnn)))
    return fail(c, O3D_ERR_INVALID, "pan_on_pan_coeff: node index out of range");
  for (Device& d : c->dev) d.kernel_ms = d.h2d_ms = d.d2h_ms = 0, d.launches = 0, d.counts[0] = d.counts[1] = 0;
  if (flops_out) *flops_out = 2.0;
  if (nsp == 0 || ntp == 0) return collect(c);

  const int ndev = (int)c->dev.size();
  const size_t nrows = (size_t)3 * ntp;
  for_each_device(c, [&](int k) {
    Device& d = c->dev[k];
    int64_t j0, j1;
    partition(nsp, ndev, k, &j0, &j1);  // the reference's own parallel axis: source columns (src/Coefficients.h:215)
    const int64_t n = j1 - j0;
    if (n == 0) return true;
    O3D_TRY(d, cudaSetDevice(d.id));
    cudaStream_t st = d.stream;
    O3D_TRY(d, d.src.ensure(((size_t)6 * nsp + (size_t)9 * ntp) * 4));
    O3D_TRY(d, d.out.ensure((size_t)3 * n * nrows * 4));
    float* db = d.src.as<float>();
    float* dsb1 = db; float* dsb2 = db + 3 * nsp;
    float* dtb1 = db + 6 * nsp; float* dtb2 = dtb1 + 3 * ntp; float* dtn = dtb2 + 3 * ntp;
    O3D_TRY(d, cudaEventRecord(d.ev[0], st));
    O3D_TRY(d, cudaMemcpyAsync(dsb1, sb1, (size_t)3 * nsp * 4, cudaMemcpyHostToDevice, st));
    O3D_TRY(d, cudaMemcpyAsync(dsb2, sb2, (size_t)3 * nsp * 4, cudaMemcpyHostToDevice, st));
    O3D_TRY(d, cudaMemcpyAsync(dtb1, tb1, (size_t)3 * ntp * 4, cudaMemcpyHostToDevice, st));
    O3D_TRY(d, cudaMemcpyAsync(dtb2, tb2, (size_t)3 * ntp * 4, cudaMemcpyHostToDevice, st));
    O3D_TRY(d, cudaMemcpyAsync(dtn, tnrm, (size_t)3 * ntp * 4, cudaMemcpyHostToDevice, st));
    if (!zero_counts(d, st)) return false;
    O3D_TRY(d, cudaEventRecord(d.ev[1], st));
    if (!upload_panels(d, st, d.panels, snn, snx, sny, snz, nsp, sidx, nullptr, nullptr, nullptr, sarea, nullptr)) return false;
    if (!upload_panels(d, st, d.tpanels, tnn, tnx, tny, tnz, ntp, tidx, nullptr, nullptr, nullptr, tarea, nullptr)) return false;
    PanCoefArgs a{};
    a.spn = d.panels.as<float4>();
    a.tpn = d.tpanels.as<float4>();
    a.nsp = nsp; a.ntp = ntp;
    a.j0 = j0; a.j1 = j1;
    a.sb1 = dsb1; a.sb2 = dsb2; a.tb1 = dtb1; a.tb2 = dtb2; a.tnrm = dtn;
    a.self = self;
    a.coeffs = d.out.as<float>();
    a.col_offset = 0;
    a.counts = d.cnt.as<unsigned long long>();
    constexpr int B = 64;
    pan_coef_kernel<B><<<dim3((unsigned)n, (unsigned)((ntp + B - 1) / B)), B, 0, st>>>(a);
    O3D_TRY(d, cudaGetLastError());
    d.launches += 1;
    O3D_TRY(d, cudaEventRecord(d.ev[2], st));
    O3D_TRY(d, cudaMemcpyAsync(coeffs + (size_t)3 * j0 * nrows, d.out.p, (size_t)3 * n * nrows * 4, cudaMemcpyDeviceToHost, st));
    if (!fetch_counts(d, st)) return false;
    O3D_TRY(d, cudaEventRecord(d.ev[3], st));
    return finish_timing(d);
  });
  const int rc = collect(c);
  if (rc == O3D_OK && flops_out) {
    // three rkernel_2vs_2p calls per pair (src/Coefficients.h:356-405; per node 31, leaf +37, split +42:
    // src/Kernels.h:1250-1292) + the final scaling pass (:452)
    double leaves = 0, splits = 0;
    for (Device& d : c->dev) leaves += (double)d.counts[0], splits += (double)d.counts[1];
    *flops_out = 3.0 * ((leaves + splits) * 31.0 + leaves * 37.0 + splits * 42.0) + 2.0 + 9.0 * (double)nsp * (double)ntp;
  }
  return rc;
}

// ---------------------------------------------------------------------------------------------------
int64_t o3d_cuda_packed_records(int64_t ns) { return ns < 0 ? 0 : padded_sources(ns); }

int o3d_cuda_pack_sources_dev(o3d_ctx* c, void* stream, int64_t ns, const float* sx, const float* sy, const float* sz,
                              const float* sr, const float* ssx, const float* ssy, const float* ssz, int64_t nrec,
                              void* packed) {
  if (!c || ns < 0 || ns >= (int64_t(1) << 31)) return fail(c, O3D_ERR_INVALID, "pack_sources_dev: bad context or count");
  if (nrec != 0 && (nrec < padded_sources(ns) || nrec % kTile != 0)) return fail(c, O3D_ERR_INVALID, "pack_sources_dev: bad nrec");
  if (ns == 0 && nrec == 0) return O3D_OK;
  if (!packed || (ns > 0 && (!sx || !sy || !sz || !ssx || !ssy || !ssz))) return fail(c, O3D_ERR_INVALID, "pack_sources_dev: NULL array");
  Device& d = c->dev[0];
  d.launches = 0;
  launch_pack(d, (cudaStream_t)stream, ns, sx, sy, sz, sr, ssx, ssy, ssz, (float4*)packed, nrec);
  return collect(c);
}

int o3d_cuda_pts_on_pts_dev(o3d_ctx* c, void* stream, int64_t nrec, const void* packed, int64_t nt, const float* tx,
                            const float* ty, const float* tz, const float* tr, float* tu, float* tv, float* tw,
                            float* tug, int64_t tug_stride) {
  if (!check_counts(c, nrec, nt) || nrec % kTile != 0) return fail(c, O3D_ERR_INVALID, "pts_on_pts_dev: bad context or counts");
  if (nrec == 0 || nt == 0) return O3D_OK;
  if (!packed || !tx || !ty || !tz || !tu || !tv || !tw || (tug && tug_stride < nt))
    return fail(c, O3D_ERR_INVALID, "pts_on_pts_dev: NULL array");
  Device& d = c->dev[0];
  d.launches = 0;
  launch_pp(d, (cudaStream_t)stream, nrec, (const float4*)packed, nt, tx, ty, tz, tr, tu, tv, tw, tug, tug_stride, nullptr);
  return collect(c);
}


int o3d_cuda_set_profiling(o3d_ctx* c, int on) {
  if (!c) return O3D_ERR_INVALID;
  for (Device& d : c->dev) d.profile = on != 0;
  return O3D_OK;
}

int o3d_cuda_dev_kernel_ms(o3d_ctx* c, double* ms) {
  if (!c || !ms) return O3D_ERR_INVALID;
  Device& d = c->dev[0];
  if (!d.profile) return fail(c, O3D_ERR_INVALID, "dev_kernel_ms: profiling is off");
  float t = 0;
  if (cudaEventSynchronize(d.evk[1]) != cudaSuccess || cudaEventElapsedTime(&t, d.evk[0], d.evk[1]) != cudaSuccess) {
    cudaGetLastError();
    return fail(c, O3D_ERR_CUDA, "dev_kernel_ms: no profiled launch on record");
  }
  *ms = t;
  return O3D_OK;
}

int o3d_cuda_probe_fp32_peak(o3d_ctx* c, double* tflops, double* ms_out) {
  if (!c || !tflops) return O3D_ERR_INVALID;
  Device& d = c->dev[0];
  double ms = 0;
  d.launches = 0;
  if (!run_fma_probe(d, tflops, &ms)) return collect(c);
  if (ms_out) *ms_out = ms;
  return collect(c);
}

}  // extern "C"

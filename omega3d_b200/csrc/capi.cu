// capi.cu - the C ABI (include/o3d_cuda.h) over the sm_100a kernels. Host-side plumbing only: device
// selection, buffers that grow geometrically and live in the context, H2D/D2H staging, target partition
// across the context's GPUs, launch-shape selection and event timing. No arithmetic of the path happens
// on the host and there is no CPU fallback: every entry point either runs the CUDA kernels or fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/o3d_cuda.h"
#include "biot_panel.cuh"
#include "biot_pp.cuh"
#include "biot_pp_cores.cuh"
#include "convect.cuh"
#include "reflect.cuh"
#include "vtu_writer.h"
#include "status_writer.h"
#include "pp2_tuned_cubin.h"   // generated (csrc/Makefile): pp_tuned.cu -> cubin -> tools/sass_patch.py -> byte array

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

// ---- host staging for pageable callers --------------------------------------------------------------------------------
// The reference hands this library plain std::vector<float> storage: pageable memory, which the driver can only move
// through its own bounce buffers, synchronously and at a few GB/s. The context therefore owns a small ring of pinned
// slots per device; large host arrays travel through it in chunks - a few helper threads copy the next chunk into (out of)
// a slot while the DMA engine moves the previous one - and arrays that already are pinned go to the DMA engine directly.
class CopyPool {
 public:
  explicit CopyPool(int helpers) {
    for (int k = 0; k < helpers; ++k) th_.emplace_back([this, k] { run(k); });
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> l(m_);
      stop_ = true;
    }
    go_.notify_all();
    for (auto& t : th_) t.join();
  }
  // memcpy split over the helpers and the calling thread; returns when every byte is in place
  void copy(void* dst, const void* src, size_t bytes) {
    const size_t parts = th_.size() + 1;
    if (th_.empty() || bytes < (size_t(1) << 20)) {
      std::memcpy(dst, src, bytes);
      return;
    }
    const size_t per = ((bytes + parts - 1) / parts + 4095) & ~size_t(4095);
    {
      std::lock_guard<std::mutex> l(m_);
      dst_ = (char*)dst; src_ = (const char*)src; bytes_ = bytes; per_ = per;
      pending_ = (int)th_.size();
      ++gen_;
    }
    go_.notify_all();
    if (bytes > per * th_.size()) std::memcpy((char*)dst + per * th_.size(), (const char*)src + per * th_.size(), bytes - per * th_.size());
    std::unique_lock<std::mutex> l(m_);
    done_.wait(l, [this] { return pending_ == 0; });
  }

 private:
  void run(int k) {
    uint64_t seen = 0;
    for (;;) {
      std::unique_lock<std::mutex> l(m_);
      go_.wait(l, [&] { return stop_ || gen_ != seen; });
      if (stop_) return;
      seen = gen_;
      char* d = dst_; const char* s = src_;
      const size_t b = bytes_, per = per_;
      l.unlock();
      const size_t off = per * k;
      if (off < b) std::memcpy(d + off, s + off, std::min(per, b - off));
      l.lock();
      if (--pending_ == 0) done_.notify_one();
    }
  }
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable go_, done_;
  char* dst_ = nullptr; const char* src_ = nullptr;
  size_t bytes_ = 0, per_ = 0;
  int pending_ = 0;
  uint64_t gen_ = 0;
  bool stop_ = false;
};

constexpr int kStageSlots = 4;
constexpr size_t kStageBytes = size_t(8) << 20;
struct Stage {
  void* slot[kStageSlots] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t idle[kStageSlots] = {nullptr, nullptr, nullptr, nullptr};   // the last DMA out of / into the slot has finished
  int next = 0;
  CopyPool* pool = nullptr;
  bool enabled = true;       // o3d_cuda_set_host_staging(ctx, 0): hand every pointer to cudaMemcpyAsync as it comes (A/B runs)
  cudaError_t ensure() {
    for (int k = 0; k < kStageSlots; ++k) {
      if (slot[k]) continue;
      cudaError_t e = cudaHostAlloc(&slot[k], kStageBytes, cudaHostAllocPortable);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&idle[k], cudaEventDisableTiming);
      if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
  }
  void release() {
    for (int k = 0; k < kStageSlots; ++k) {
      if (slot[k]) cudaFreeHost(slot[k]);
      if (idle[k]) cudaEventDestroy(idle[k]);
      slot[k] = nullptr; idle[k] = nullptr;
    }
    delete pool;
    pool = nullptr;
  }
};

bool host_is_pinned(const void* p) {
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

struct Device {
  int id = 0;
  int sm_count = 0;
  int clock_khz = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;                            // uploads that overlap the kernel running on `stream`
  cudaEvent_t evx[2] = {nullptr, nullptr};                   // cross-stream / cross-device ordering (no timing)
  Stage stage;
  DevBuf acc;                                                // FP64 sums of a host call (PPArgs::acc64)
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // h2d start, compute start, compute end, d2h end
  cudaEvent_t evk[2] = {nullptr, nullptr};                   // around the dominant kernel of a *_dev call
  bool profile = false;
  bool tuned = true;                                         // launch pp2_kernel from the post-processed cubin (TunedKernels)
  int pan_queue = 1;                                         // 1: panels -> points with the warp-level work pool (pan_pts_queue_kernel);
                                                             // 2: particles -> panels too (pts_pan_queue_kernel: measured no faster); 0: neither
  int core = O3D_CORE_WL;                                    // core function of the particle kernels (o3d_cuda_set_core_func)
  DevBuf src, packed, targ, out, work, ppwork, geom, panels, tpanels, cnt, rng;
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
  bool use_graphs = true;   // o3d_cuda_set_graphs
  bool sources_pcie_all = false;   // O3D_CUDA_SOURCES_PCIE_ALL=1 (A/B measurement): every device of a multi-device context uploads and
                                   // packs the sources itself (round 1) instead of device 0 uploading once and the peers pulling the
                                   // packed records over NVLink
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

// host -> device of `bytes` on stream st: straight to the DMA engine when the caller's memory is pinned, else through the
// device's pinned ring (helper threads fill slot k+1 while slot k is in flight)
bool h2d(Device& d, cudaStream_t st, void* dst, const void* src, size_t bytes) {
  if (bytes == 0) return true;
  Stage& g = d.stage;
  if (!g.enabled || bytes < (size_t(256) << 10) || host_is_pinned(src)) {
    O3D_TRY(d, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
    return true;
  }
  O3D_TRY(d, g.ensure());
  for (size_t off = 0; off < bytes; off += kStageBytes) {
    const size_t len = std::min(kStageBytes, bytes - off);
    const int k = g.next;
    g.next = (g.next + 1) % kStageSlots;
    O3D_TRY(d, cudaEventSynchronize(g.idle[k]));
    g.pool->copy(g.slot[k], (const char*)src + off, len);
    O3D_TRY(d, cudaMemcpyAsync((char*)dst + off, g.slot[k], len, cudaMemcpyHostToDevice, st));
    O3D_TRY(d, cudaEventRecord(g.idle[k], st));
  }
  return true;
}

// device -> host, complete on return for pageable destinations (the DMA into slot k+1.. runs while slot k is copied out);
// pinned destinations are only enqueued - the caller synchronises the stream as before
bool d2h(Device& d, cudaStream_t st, void* dst, const void* src, size_t bytes) {
  if (bytes == 0) return true;
  Stage& g = d.stage;
  if (!g.enabled || bytes < (size_t(256) << 10) || host_is_pinned(dst)) {
    O3D_TRY(d, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
    return true;
  }
  O3D_TRY(d, g.ensure());
  struct Pending { int k; size_t off, len; };
  std::deque<Pending> fifo;
  auto drain_one = [&]() {
    const Pending q = fifo.front();
    fifo.pop_front();
    O3D_TRY(d, cudaEventSynchronize(g.idle[q.k]));
    g.pool->copy((char*)dst + q.off, g.slot[q.k], q.len);
    return true;
  };
  for (size_t off = 0; off < bytes; off += kStageBytes) {
    const size_t len = std::min(kStageBytes, bytes - off);
    if ((int)fifo.size() == kStageSlots && !drain_one()) return false;
    const int k = g.next;
    g.next = (g.next + 1) % kStageSlots;
    O3D_TRY(d, cudaEventSynchronize(g.idle[k]));            // (a slot last used by an upload)
    O3D_TRY(d, cudaMemcpyAsync(g.slot[k], (const char*)src + off, len, cudaMemcpyDeviceToHost, st));
    O3D_TRY(d, cudaEventRecord(g.idle[k], st));
    fifo.push_back({k, off, len});
  }
  while (!fifo.empty())
    if (!drain_one()) return false;
  return true;
}

// ---- the post-processed copies of the two product pp2_kernel instantiations ---------------------------------
// Same source (csrc/pp_tuned.cu -> biot_pp.cuh), compiled to a cubin and passed through tools/sass_patch.py, which
// only lengthens the operand-reuse chains of adjacent packed FP32 instructions: identical instructions and results,
// about 2 % fewer FMA-pipe cycles. Loaded once per process; a load failure is an error of o3d_cuda_create, not a
// silent switch to the linked copies (those stay reachable through o3d_cuda_set_tuned_kernels(ctx, 0) for A/B tests).
static_assert(kPPTgrad == 2 && kPPTvel == 4 && kPPBlock == 384 && kPPSmallBlock == 128, "kernel names below spell these template arguments");
struct TunedKernels {
  cudaLibrary_t lib = nullptr;
  cudaKernel_t grad = nullptr, vel = nullptr;
  cudaKernel_t grad_small = nullptr, vel_small = nullptr;   // the 128-thread instantiations (systems below one target block)
  cudaKernel_t core[4][2] = {};   // ppc_kernel<core, .., grad?>: [O3D_CORE_RM .. O3D_CORE_V2][0 = velocity only, 1 = with gradients]
  cudaError_t status = cudaSuccess;
  const char* where = "";
};
const TunedKernels& tuned_kernels() {
  static const TunedKernels t = [] {
    TunedKernels k;
    k.status = cudaLibraryLoadData(&k.lib, o3d_pp2_tuned_cubin, nullptr, nullptr, 0, nullptr, nullptr, 0);
    k.where = "cudaLibraryLoadData(pp2_tuned.cubin)";
    if (k.status == cudaSuccess) {
      k.status = cudaLibraryGetKernel(&k.grad, k.lib, "_ZN3o3d10pp2_kernelILi2ELb1ELi384EEEvNS_6PPArgsE");
      k.where = "cudaLibraryGetKernel(pp2_kernel<2,true,384>)";
    }
    if (k.status == cudaSuccess) {
      k.status = cudaLibraryGetKernel(&k.vel, k.lib, "_ZN3o3d10pp2_kernelILi4ELb0ELi384EEEvNS_6PPArgsE");
      k.where = "cudaLibraryGetKernel(pp2_kernel<4,false,384>)";
    }
    if (k.status == cudaSuccess) {
      k.status = cudaLibraryGetKernel(&k.grad_small, k.lib, "_ZN3o3d10pp2_kernelILi2ELb1ELi128EEEvNS_6PPArgsE");
      k.where = "cudaLibraryGetKernel(pp2_kernel<2,true,128>)";
    }
    if (k.status == cudaSuccess) {
      k.status = cudaLibraryGetKernel(&k.vel_small, k.lib, "_ZN3o3d10pp2_kernelILi4ELb0ELi128EEEvNS_6PPArgsE");
      k.where = "cudaLibraryGetKernel(pp2_kernel<4,false,128>)";
    }
    for (int core = O3D_CORE_RM; core <= O3D_CORE_V2 && k.status == cudaSuccess; ++core)
      for (int g = 0; g < 2 && k.status == cudaSuccess; ++g) {
        char name[96];
        snprintf(name, sizeof name, "_ZN3o3d10ppc_kernelILi%dELi%dELb%dELi384EEEvNS_6PPArgsE", core, g ? kPPTgrad : kPPTvel, g);
        k.status = cudaLibraryGetKernel(&k.core[core][g], k.lib, name);
        k.where = "cudaLibraryGetKernel(ppc_kernel<core,T,grad,384>)";
      }
    if (k.status != cudaSuccess) cudaGetLastError();
    return k;
  }();
  return t;
}

// ---- launch shape for particles -> points ---------------------------------------------------------
// Product configuration: packed FFMA2 kernels, one PERSISTENT 384-thread CTA per SM, 2 targets per thread with gradients,
// 4 without, over a static stream-K partition of the (target block, source tile) units (csrc/biot_pp.cuh: PPPlan):
// every CTA the same number of tiles (+-1) whatever the target count.
struct PPShape {
  int block;          // threads per CTA: kPPBlock, or kPPSmallBlock for a system of at most one product-size target block
  int nblocks;        // target blocks of block * T targets
  int grid;           // CTAs = min(units, resident slots)
  int64_t units;      // nblocks * ntiles
  int split_blocks;   // tail blocks shared by more than one CTA (finished by pp_fixup_kernel)
  PPPlan plan;
};
// (kPPResident = 1 CTA of 384 threads per SM, register-limited: <= 168 registers per thread, lib/ptxas.log; csrc/biot_pp.cuh)
// workspace of one launch: 2 slots per CTA x (12 rows x 768 targets | 3 rows x 1536 targets) FP64 - a constant of the device
constexpr size_t pp_workspace_bytes(int sm_count) {
  return (size_t)sm_count * kPPResident * kPPSlots * 12 * (kPPBlock * kPPTgrad) * sizeof(double);
}
static_assert(12 * kPPTgrad >= 3 * kPPTvel, "the gradient kernel's slots are the larger ones");

// pair-term flops of the reference's core functions, src/CoreFunc.h: flops_tv_grads / flops_tp_grads /
// flops_tv_nograds / flops_tp_nograds for WL (:251-288), Rosenhead-Moore (:50-82), exponential (:202-237), Vatistas (:304-340)
double core_flops(int core, bool grad, bool blob) {
  static const double t[4][4] = {{16, 14, 10, 8}, {9, 7, 7, 5}, {14, 11, 12, 9}, {13, 10, 11, 8}};
  return t[core][(grad ? 0 : 2) + (blob ? 0 : 1)];
}
// flops per interaction as the reference's routines report them: flops_0v_0bg = 54 + tv_grads, _0pg = 54 + tp_grads,
// _0b = 23 + tv_nograds, _0p = 23 + tp_nograds (src/Kernels.h:50,94,155,253); WL: 70 / 68 / 33 / 31
double pp_pair_flops(int core, bool grad, bool blob) { return (grad ? 54.0 : 23.0) + core_flops(core, grad, blob); }
// ... and of the panel leaves: flops_0vs_0pg = 79 + tp_grads, flops_0vs_0p = 29 + tp_nograds (src/Kernels.h:115,294); WL: 93 / 37
double leaf_flops(int core, bool grad) { return (grad ? 79.0 : 29.0) + core_flops(core, grad, false); }

PPShape pp_shape(int sm_count, int64_t ntiles, int64_t nt, bool grad, bool wl_core = true) {
  const int T = grad ? kPPTgrad : kPPTvel;
  PPShape s;
  s.block = wl_core && nt <= (int64_t)kPPBlock * T ? kPPSmallBlock : kPPBlock;      // (the alternate cores ship as 384-thread kernels only)
  const int per_cta = s.block * T;
  s.nblocks = (int)((nt + per_cta - 1) / per_cta);
  s.units = (int64_t)s.nblocks * ntiles;
  s.plan = pp_make_plan(sm_count * (kPPWarpsPerSM * 32 / s.block), s.nblocks, (int)ntiles);
  s.grid = s.plan.P;
  s.split_blocks = 0;
  for (int j = 1; j < s.plan.Pt; ++j) {   // pp_fixup_kernel's own test: a boundary inside a block, the first one there
    const int64_t cut = s.plan.begin(j), start = cut / ntiles * ntiles;
    if (cut != start && s.plan.begin(j - 1) <= start) ++s.split_blocks;
  }
  return s;
}

bool launch_pp(Device& d, cudaStream_t st, int64_t nrec, const float4* packed, int64_t nt, const float* tx,
               const float* ty, const float* tz, const float* tr, float* tu, float* tv, float* tw, float* tug,
               int64_t tug_stride, double* acc64 = nullptr, int64_t acc_stride = 0, int want_grad = -1) {
  // acc64 != nullptr: store the FP64 sums there instead of adding them into tu.. (which may then be NULL; want_grad says
  // whether gradients are computed)
  const bool grad = want_grad < 0 ? tug != nullptr : want_grad != 0;
  const int64_t ntiles = nrec / kTile;
  const PPShape s = pp_shape(d.sm_count, ntiles, nt, grad, d.core == O3D_CORE_WL);
  PPArgs a{};
  a.src = packed;
  a.ntiles = (int)ntiles;
  a.nblocks = s.nblocks;
  a.slots = d.sm_count * (kPPWarpsPerSM * 32 / s.block);
  a.nt = nt;
  a.tx = tx; a.ty = ty; a.tz = tz; a.tr = tr;
  a.tu = tu; a.tv = tv; a.tw = tw;
  a.tug = tug;
  a.tug_stride = tug_stride;
  a.sign = 1.0f;
  a.acc64 = acc64;
  a.acc_stride = acc_stride;
  // fixed-size, allocated once per device and never moved: captured CUDA graphs may hold its address
  O3D_TRY(d, d.ppwork.ensure(pp_workspace_bytes(d.sm_count)));
  a.partial = d.ppwork.as<double>();
  {
    // radius ranges for the uniform-radius fast path (one pass over the records' radius lane - r^2, r^3 or r^4 by core - and
    // the target radii)
    O3D_TRY(d, d.rng.ensure(4 * sizeof(uint32_t)));
    O3D_TRY(d, cudaMemsetAsync(d.rng.p, 0, 4 * sizeof(uint32_t), st));
    pp_scan_kernel<<<d.sm_count * 4, 256, 0, st>>>(nrec, packed, nt, tr, d.rng.as<uint32_t>());
    O3D_TRY(d, cudaGetLastError());
    d.launches += 1;
    a.radius_range = d.rng.as<uint32_t>();
  }
  const dim3 grid((unsigned)s.grid);
  if (d.profile) O3D_TRY(d, cudaEventRecord(d.evk[0], st));
  if (d.core != O3D_CORE_WL) {
    // the alternate core functions of src/CoreFunc.h (csrc/biot_pp_cores.cuh); the stream was packed for d.core
    if (d.tuned) {   // the post-processed copy (same instructions and results; tools/sass_patch.py)
      void* params[] = {&a};
      O3D_TRY(d, cudaLaunchKernel((const void*)tuned_kernels().core[d.core][grad ? 1 : 0], grid, dim3(kPPBlock), params, 0, st));
    } else {
#define O3D_PPC_LAUNCH(CORE)                                                                        \
  if (grad) ppc_kernel<CORE, kPPTgrad, true, kPPBlock><<<grid, kPPBlock, 0, st>>>(a);               \
  else      ppc_kernel<CORE, kPPTvel, false, kPPBlock><<<grid, kPPBlock, 0, st>>>(a)
    if (d.core == O3D_CORE_RM) { O3D_PPC_LAUNCH(kCoreRM); }
    else if (d.core == O3D_CORE_EXP) { O3D_PPC_LAUNCH(kCoreEXP); }
    else { O3D_PPC_LAUNCH(kCoreV2); }
#undef O3D_PPC_LAUNCH
    }
  } else if (d.tuned) {
    const TunedKernels& tk = tuned_kernels();
    void* params[] = {&a};
    const bool small = s.block == kPPSmallBlock;
    O3D_TRY(d, cudaLaunchKernel((const void*)(grad ? (small ? tk.grad_small : tk.grad) : (small ? tk.vel_small : tk.vel)), grid, dim3(s.block), params, 0, st));
  } else if (s.block == kPPSmallBlock) {
    if (grad) pp2_kernel<kPPTgrad, true, kPPSmallBlock><<<grid, kPPSmallBlock, 0, st>>>(a);
    else      pp2_kernel<kPPTvel, false, kPPSmallBlock><<<grid, kPPSmallBlock, 0, st>>>(a);
  } else if (grad) {
    pp2_kernel<kPPTgrad, true, kPPBlock><<<grid, kPPBlock, 0, st>>>(a);
  } else {
    pp2_kernel<kPPTvel, false, kPPBlock><<<grid, kPPBlock, 0, st>>>(a);
  }
  O3D_TRY(d, cudaGetLastError());
  if (d.profile) O3D_TRY(d, cudaEventRecord(d.evk[1], st));
  d.launches += 1;
  if (s.plan.Pt > 1) {
    // the tail blocks cut by a CTA-range boundary: add their pieces in unit order (one CTA per boundary)
    pp_fixup_kernel<<<(unsigned)(s.plan.Pt - 1), 256, 0, st>>>(grad ? 12 : 3, s.block * (grad ? kPPTgrad : kPPTvel), s.plan, nt, a.partial, tu, tv, tw,
                                                             tug, tug_stride, 1.0f, a.acc64, a.acc_stride);
    O3D_TRY(d, cudaGetLastError());
    d.launches += 1;
  }
  return true;
}

bool launch_pack(Device& d, cudaStream_t st, int64_t ns, const float* sx, const float* sy, const float* sz,
                 const float* sr, const float* wx, const float* wy, const float* wz, float4* packed, int64_t nrec = 0) {
  const int64_t npad = nrec > 0 ? nrec : padded_sources(ns);
  const unsigned blocks = (unsigned)((npad / 2 + 255) / 256);
  if (d.core == O3D_CORE_WL) pp_pack2_kernel<<<blocks, 256, 0, st>>>(ns, npad, sx, sy, sz, sr, wx, wy, wz, packed);
  else ppc_pack2_kernel<<<blocks, 256, 0, st>>>(d.core, ns, npad, sx, sy, sz, sr, wx, wy, wz, packed);   // radius lane of that core
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

// ---- matrix-free BEM operator (include/o3d_cuda.h: o3d_bem_op) ---------------------------------------
struct BemDev {
  DevBuf bases, spanels, tpanels, x, y, work, cnt;
  int64_t i0 = 0, ni = 0;    // target-panel rows of this device
};

}  // namespace

struct o3d_bem_op {
  int64_t nsp = 0, ntp = 0;
  int self = 0;
  std::vector<BemDev> dev;
};

namespace {

// ---- device-resident particle collections (include/o3d_cuda.h: o3d_particles) ------------------------
// Row layout of one per-device state block (rows of `cap` floats; only this device's slice is stored):
enum { kRowX = 0, kRowS = 3, kRowR = 6, kRowE = 7, kRowU = 8, kRowG = 11, kRowsMain = 20 };
// an interim Runge-Kutta copy keeps position, strength, velocity, gradient: x 0-2, s 3-5, u 6-8, ug 9-17
enum { kIntX = 0, kIntS = 3, kIntU = 6, kIntG = 9, kRowsInterim = 18 };

// A static body attached to a resident collection (include/o3d_cuda.h: o3d_cuda_particles_set_body): its geometry and the
// packed records of the three panel kernels live on every device of the context for as long as it is attached.
struct BodyDev {
  DevBuf geom;      // nodes 3*nn | idx 3*np | ts 3*np | area np | sss np | normals 3*np
  DevBuf panels;    // pan_pts / pts_pan records (80 B per panel, carry the current strengths)
  DevBuf refl;      // closest-point records of reflect_kernel (144 B per panel)
  DevBuf pu;        // 3 x np floats: panel-centre velocities (the BEM right-hand side before projection)
  DevBuf work;      // FP64 slabs of the panel kernels
  DevBuf cnt;       // leaf / split / moved counters
};

struct PartDev {
  BodyDev body;
  DevBuf main, interim[2], packed, stats, totals;
  cudaEvent_t packed_ready = nullptr;   // this device's slice of the packed stream is written
  cudaEvent_t pulled = nullptr;         // this device has copied every peer's slice
  cudaEvent_t ev[2] = {nullptr, nullptr};
  cudaGraphExec_t graph = nullptr;      // one captured advect step (single-device contexts)
  int64_t graph_n = -1;
  int graph_order = 0;
  int graph_core = -1;                  // core function / kernel copy the captured launches belong to
  bool graph_tuned = true;
  double graph_dt = 0, graph_fs[3] = {0, 0, 0};
  bool graph_failed = false;
  int launches_per_step = 0;
  int64_t t0 = 0, n = 0, cap = 0;       // slice [t0, t0 + n) of the collection, row stride
};

}  // namespace

struct o3d_particles {
  // the attached body (has_body): panel counts, the host callback that turns panel-centre velocities into strengths (the
  // reference's BEM solve stays host code), the clear-inner-layer parameters of Convection::advect, host staging
  bool has_body = false;
  int64_t b_nn = 0, b_np = 0;
  o3d_bem_solve_fn solve = nullptr;
  void* solve_user = nullptr;
  float cutoff_mult = 0.0f, ips = 0.0f;
  std::vector<float> h_pu, h_str;
  int64_t moved = 0;        // particles pushed out by clear-inner passes of the last call
  int solves = 0;           // callbacks made by the last call
  int64_t n = 0;
  int64_t per = 0;          // particles per device, a whole number of tiles
  int64_t nrec = 0;         // packed records in the whole stream = ndev slices of `per`, the last one padded
  std::vector<PartDev> dev;
  bool peers_enabled = false;
};

namespace {

float* prow(DevBuf& b, int64_t cap, int row) { return b.as<float>() + (size_t)row * cap; }

bool part_layout(o3d_ctx* c, o3d_particles* p, int64_t n) {
  const int nd = (int)c->dev.size();
  p->n = n;
  p->per = padded_sources((n + nd - 1) / nd);   // tile-aligned blocks: the packed stream is the same for any device count
  p->nrec = 0;
  for (int k = 0; k < nd; ++k) {
    PartDev& q = p->dev[k];
    q.t0 = std::min(n, p->per * k);
    q.n = std::min(n, p->per * (k + 1)) - q.t0;
    if (q.n > 0) p->nrec = q.t0 + padded_sources(q.n);
  }
  for (int k = 0; k < nd; ++k) {
    Device& d = c->dev[k];
    PartDev& q = p->dev[k];
    O3D_TRY(d, cudaSetDevice(d.id));
    const int64_t cap = std::max<int64_t>(q.n, 1);
    if (cap > q.cap || cap * 2 < q.cap) q.cap = (cap + cap / 8 + 63) & ~int64_t(63);   // head-room for growth, 256-byte rows
    O3D_TRY(d, q.main.ensure((size_t)kRowsMain * q.cap * 4));
    O3D_TRY(d, q.packed.ensure((size_t)std::max<int64_t>(p->nrec, kTile) * 32));
    O3D_TRY(d, q.stats.ensure(2 * sizeof(uint32_t)));
    if (!q.packed_ready) {
      O3D_TRY(d, cudaEventCreateWithFlags(&q.packed_ready, cudaEventDisableTiming));
      O3D_TRY(d, cudaEventCreateWithFlags(&q.pulled, cudaEventDisableTiming));
      O3D_TRY(d, cudaEventCreate(&q.ev[0]));
      O3D_TRY(d, cudaEventCreate(&q.ev[1]));
    }
  }
  if (nd > 1 && !p->peers_enabled) {
    for (int k = 0; k < nd; ++k) {
      cudaSetDevice(c->dev[k].id);
      for (int j = 0; j < nd; ++j)
        if (j != k) {
          int can = 0;
          cudaDeviceCanAccessPeer(&can, c->dev[k].id, c->dev[j].id);
          if (can && cudaDeviceEnablePeerAccess(c->dev[j].id, 0) != cudaSuccess) cudaGetLastError();  // already enabled is fine
        }
    }
    p->peers_enabled = true;
  }
  return true;
}

// A view of one Points-like state on one device: where its x, s, r, u, ug rows live.
struct PartView {
  float* x[3]; float* s[3]; float* r; float* u[3]; float* ug; int64_t stride;
};
PartView view_main(PartDev& q) {
  PartView v;
  for (int d = 0; d < 3; ++d) v.x[d] = prow(q.main, q.cap, kRowX + d), v.s[d] = prow(q.main, q.cap, kRowS + d), v.u[d] = prow(q.main, q.cap, kRowU + d);
  v.r = prow(q.main, q.cap, kRowR);
  v.ug = prow(q.main, q.cap, kRowG);
  v.stride = q.cap;
  return v;
}
PartView view_interim(PartDev& q, int which) {
  PartView v;
  DevBuf& b = q.interim[which];
  for (int d = 0; d < 3; ++d) v.x[d] = prow(b, q.cap, kIntX + d), v.s[d] = prow(b, q.cap, kIntS + d), v.u[d] = prow(b, q.cap, kIntU + d);
  v.r = prow(q.main, q.cap, kRowR);   // radii never change inside a convection step
  v.ug = prow(b, q.cap, kIntG);
  v.stride = q.cap;
  return v;
}

// gridDim.y split of the source tiles when the target axis alone cannot fill the GPU
int pan_nsplit(const Device& d, int64_t gx, int64_t ntiles) {
  const int64_t fill = (int64_t)d.sm_count * 8;
  int64_t n = 1;
  if (gx < fill) n = std::min<int64_t>(std::min<int64_t>(ntiles, (fill + gx - 1) / gx), 4096);
  return (int)std::max<int64_t>(n, 1);
}

// panels -> points, 128-thread CTAs: the warp-queue kernel (product) or the per-lane baseline (o3d_cuda_set_panel_queue)
void launch_pan_pts(const Device& d, cudaStream_t st, const PanPtsArgs& a, dim3 grid, bool grad) {
  constexpr int B = 128;
  if (d.pan_queue) {
    if (grad) pan_pts_queue_kernel<true, B><<<grid, B, 0, st>>>(a);
    else      pan_pts_queue_kernel<false, B><<<grid, B, 0, st>>>(a);
  } else {
    if (grad) pan_pts_kernel<true, B><<<grid, B, 0, st>>>(a);
    else      pan_pts_kernel<false, B><<<grid, B, 0, st>>>(a);
  }
}

struct BodyDev;
bool body_solve(o3d_ctx* c, o3d_particles* p, const double* fs);
bool body_on_particles(Device& d, cudaStream_t st, o3d_particles* p, PartDev& q, const PartView& v, bool grad);

// Convection::find_vels for one state, on every device of the context (src/Convection.h:130-184 with no boundaries):
// zero_vels; pack own slice; exchange packed slices; particles -> own targets; finalize_vels(fs).
// `sel`: 0 = main state, 1/2 = interim copy 0/1. Everything is enqueued on the per-device streams; no host sync.
bool part_find_vels(o3d_ctx* c, o3d_particles* p, int sel, const double* fs, bool grad, bool in_capture, bool solve_body = false) {
  const int nd = (int)c->dev.size();
  // 1. each device: (wait until every peer has finished reading its previous packed slice) zero, pack own slice
  for (int k = 0; k < nd; ++k) {
    Device& d = c->dev[k];
    PartDev& q = p->dev[k];
    if (q.n == 0) continue;
    O3D_TRY(d, cudaSetDevice(d.id));
    cudaStream_t st = d.stream;
    if (nd > 1)
      for (int j = 0; j < nd; ++j)
        if (j != k && p->dev[j].n > 0) O3D_TRY(d, cudaStreamWaitEvent(st, p->dev[j].pulled, 0));
    PartView v = sel == 0 ? view_main(q) : view_interim(q, sel - 1);
    const unsigned gb = (unsigned)((q.n + 255) / 256);
    pts_fill_kernel<<<gb, 256, 0, st>>>(q.n, 3, v.u[0], v.stride, 0.0f);
    if (grad) pts_fill_kernel<<<gb, 256, 0, st>>>(q.n, 9, v.ug, v.stride, 0.0f);
    O3D_TRY(d, cudaGetLastError());
    d.launches += grad ? 2 : 1;
    float4* slice = q.packed.as<float4>() + (size_t)q.t0 * 2;
    if (!launch_pack(d, st, q.n, v.x[0], v.x[1], v.x[2], v.r, v.s[0], v.s[1], v.s[2], slice, padded_sources(q.n))) return false;
    if (nd > 1) O3D_TRY(d, cudaEventRecord(q.packed_ready, st));
  }
  // 2. each device pulls every peer's slice of the packed stream (NVLink peer copy)
  for (int k = 0; k < nd && nd > 1; ++k) {
    Device& d = c->dev[k];
    PartDev& q = p->dev[k];
    if (q.n == 0) continue;
    O3D_TRY(d, cudaSetDevice(d.id));
    cudaStream_t st = d.stream;
    for (int j = 0; j < nd; ++j) {
      PartDev& o = p->dev[j];
      if (j == k || o.n == 0) continue;
      O3D_TRY(d, cudaStreamWaitEvent(st, o.packed_ready, 0));
      const size_t off = (size_t)o.t0 * 32, bytes = (size_t)padded_sources(o.n) * 32;
      O3D_TRY(d, cudaMemcpyPeerAsync((char*)q.packed.p + off, d.id, (const char*)o.packed.p + off, c->dev[j].id, bytes, st));
    }
    O3D_TRY(d, cudaEventRecord(q.pulled, st));
  }
  // 2b. with a body whose strengths are solved for: the BEM right-hand side from this state, the host solve, new panel records
  if (solve_body && p->has_body && p->solve && !body_solve(c, p, fs)) return false;
  // 3. each device evaluates its targets: particles (then the body's panels) on its particles, finalize_vels(fs)
  for (int k = 0; k < nd; ++k) {
    Device& d = c->dev[k];
    PartDev& q = p->dev[k];
    if (q.n == 0) continue;
    O3D_TRY(d, cudaSetDevice(d.id));
    cudaStream_t st = d.stream;
    PartView v = sel == 0 ? view_main(q) : view_interim(q, sel - 1);
    if (d.profile && !in_capture) O3D_TRY(d, cudaEventRecord(q.ev[0], st));
    if (!launch_pp(d, st, p->nrec, q.packed.as<float4>(), q.n, v.x[0], v.x[1], v.x[2], v.r, v.u[0], v.u[1], v.u[2],
                   grad ? v.ug : nullptr, v.stride))
      return false;
    if (p->has_body && !body_on_particles(d, st, p, q, v, grad)) return false;
    pts_finalize_kernel<<<(unsigned)((q.n + 255) / 256), 256, 0, st>>>(q.n, v.u[0], v.u[1], v.u[2], grad ? v.ug : nullptr, v.stride,
                                                                      fs[0], fs[1], fs[2]);
    O3D_TRY(d, cudaGetLastError());
    d.launches += 1;
  }
  return true;
}

// ---- resident body --------------------------------------------------------------------------------------------------
float* body_row(BodyDev& b, int64_t nn, int64_t np, int which) {
  // geom layout: nodes 3*nn | idx 3*np | ts 3*np | area np | sss np | normals 3*np
  float* g = b.geom.as<float>();
  switch (which) {
    case 0: return g;                        // node x (y at +nn, z at +2nn)
    case 1: return g + 3 * nn;               // idx (uint32)
    case 2: return g + 3 * nn + 3 * np;      // ts x (y, z at +np, +2np)
    case 3: return g + 3 * nn + 6 * np;      // area
    case 4: return g + 3 * nn + 7 * np;      // sss
    default: return g + 3 * nn + 8 * np;     // normals x (y, z at +np, +2np)
  }
}

bool body_repack(Device& d, cudaStream_t st, o3d_particles* p, BodyDev& b, bool with_source) {
  const int64_t nn = p->b_nn, np = p->b_np, npad = padded_panels(np);
  float* ts = body_row(b, nn, np, 2);
  pan_pack_kernel<<<(unsigned)((npad + 127) / 128), 128, 0, st>>>(np, npad, body_row(b, nn, np, 0), body_row(b, nn, np, 0) + nn,
                                                                 body_row(b, nn, np, 0) + 2 * nn, (const uint32_t*)body_row(b, nn, np, 1), ts,
                                                                 ts + np, ts + 2 * np, body_row(b, nn, np, 3),
                                                                 with_source ? body_row(b, nn, np, 4) : nullptr, b.panels.as<float4>());
  O3D_TRY(d, cudaGetLastError());
  d.launches += 1;
  return true;
}

// panels -> the particles of one state on one device, added into its u (and ug): the second half of Convection::find_vels
bool body_on_particles(Device& d, cudaStream_t st, o3d_particles* p, PartDev& q, const PartView& v, bool grad) {
  BodyDev& b = q.body;
  const int64_t npad = padded_panels(p->b_np);
  PanPtsArgs a{};
  a.pan = b.panels.as<float4>();
  a.ntiles = (int)(npad / kPanTile);
  a.nt = q.n;
  a.tx = v.x[0]; a.ty = v.x[1]; a.tz = v.x[2];
  a.tu = v.u[0]; a.tv = v.u[1]; a.tw = v.u[2];
  a.tug = grad ? v.ug : nullptr;
  a.tug_stride = v.stride;
  a.counts = nullptr;
  constexpr int B = 128;
  const int64_t gx = (q.n + B - 1) / B;
  const int nout = grad ? 12 : 3;
  a.nsplit = pan_nsplit(d, gx, a.ntiles);
  if (a.nsplit > 1) {
    O3D_TRY(d, b.work.ensure((size_t)a.nsplit * nout * q.n * sizeof(double)));
    a.partial = b.work.as<double>();
  }
  const dim3 grid((unsigned)gx, (unsigned)a.nsplit);
  launch_pan_pts(d, st, a, grid, grad);
  O3D_TRY(d, cudaGetLastError());
  d.launches += 1;
  if (a.nsplit > 1) {
    pp_finish_kernel<<<(unsigned)((q.n + 255) / 256), 256, 0, st>>>(nout, a.nsplit, q.n, a.partial, a.tu, a.tv, a.tw, a.tug, v.stride, 1.0f);
    O3D_TRY(d, cudaGetLastError());
    d.launches += 1;
  }
  return true;
}

// clear_inner_layer(1, bdry, state, cutoff_mult, ips) on one state of every device (src/Reflect.h:625-655, :446-620)
bool body_clear_inner(o3d_ctx* c, o3d_particles* p, int sel) {
  const int64_t np = p->b_np, npad = ((np + kRefTile - 1) / kRefTile) * kRefTile;
  for (size_t k = 0; k < c->dev.size(); ++k) {
    Device& d = c->dev[k];
    PartDev& q = p->dev[k];
    if (q.n == 0) continue;
    O3D_TRY(d, cudaSetDevice(d.id));
    PartView v = sel == 0 ? view_main(q) : view_interim(q, sel - 1);
    ReflectArgs a{};
    a.pan = q.body.refl.as<float4>();
    a.np = np;
    a.ntiles = (int)(npad / kRefTile);
    a.nt = q.n;
    a.tx = v.x[0]; a.ty = v.x[1]; a.tz = v.x[2];
    a.mode = 1;
    a.cutoff = p->cutoff_mult * p->ips;      // float product, as "_cutoff_mult*_ips" with S = float (src/Reflect.h:537)
    a.count = q.body.cnt.as<unsigned long long>() + 2;
    constexpr int B = 128;
    reflect_kernel<B><<<(unsigned)((q.n + B - 1) / B), B, 0, d.stream>>>(a);
    O3D_TRY(d, cudaGetLastError());
    d.launches += 1;
  }
  return true;
}

// Right-hand side of the BEM solve for one state (solve_bem, src/BEMHelper.h:83-94): zero the panel-centre velocities,
// particles -> panels (subtracting, src/Influence.h:1210-1212) on device 0, whose packed stream holds every particle of the
// state; then the host callback - the rest of solve_bem: finalize_vels(fs), right-hand side, solve, set_str; the reference's
// BEM stays host code - returns the strengths and every device repacks its panel records with them. Needs the state's packed
// stream in place (part_find_vels steps 1-2).
bool body_solve(o3d_ctx* c, o3d_particles* p, const double* fs) {
  (void)fs;   // the freestream enters in the host's finalize_vels
  const int nd = (int)c->dev.size();
  const int64_t np = p->b_np, nn = p->b_nn;
  Device& d = c->dev[0];
  PartDev& q = p->dev[0];
  BodyDev& b = q.body;
  O3D_TRY(d, cudaSetDevice(d.id));
  cudaStream_t st = d.stream;
  float* pu = b.pu.as<float>();
  pts_fill_kernel<<<(unsigned)((np + 255) / 256), 256, 0, st>>>(np, 3, pu, np, 0.0f);
  O3D_TRY(d, cudaGetLastError());
  d.launches += 1;
  for (int k = 0; k < nd; ++k) {             // one launch per device slice of the stream: padding records are never read
    const PartDev& o = p->dev[k];
    if (o.n == 0) continue;
    PtsPanArgs a{};
    a.src = q.packed.as<float4>() + (size_t)o.t0 * 2;
    a.ns = o.n;
    a.ntiles = (int)(padded_sources(o.n) / kTile);
    a.np = np;
    a.pan = b.panels.as<float4>();
    a.counts = nullptr;
    constexpr int B = 64;
    const int64_t gx = (np + B - 1) / B;
    a.nsplit = pan_nsplit(d, gx, a.ntiles);
    O3D_TRY(d, b.work.ensure((size_t)a.nsplit * 3 * np * sizeof(double)));
    a.partial = b.work.as<double>();
    if (d.pan_queue >= 2) pts_pan_queue_kernel<B><<<dim3((unsigned)gx, (unsigned)a.nsplit), B, 0, st>>>(a);
    else             pts_pan_kernel<B><<<dim3((unsigned)gx, (unsigned)a.nsplit), B, 0, st>>>(a);
    O3D_TRY(d, cudaGetLastError());
    pp_finish_kernel<<<(unsigned)((np + 255) / 256), 256, 0, st>>>(3, a.nsplit, np, a.partial, pu, pu + np, pu + 2 * np, nullptr, np, -1.0f);
    O3D_TRY(d, cudaGetLastError());
    d.launches += 2;
  }
  p->h_pu.resize((size_t)3 * np);
  p->h_str.assign((size_t)4 * np, 0.0f);
  O3D_TRY(d, cudaMemcpyAsync(p->h_pu.data(), pu, (size_t)3 * np * 4, cudaMemcpyDeviceToHost, st));
  O3D_TRY(d, cudaStreamSynchronize(st));
  int have_source = 0;
  float* hs = p->h_str.data();
  if (p->solve(p->solve_user, np, p->h_pu.data(), hs, hs + np, hs + 2 * np, hs + 3 * np, &have_source) != 0) {
    d.status = cudaErrorUnknown;
    d.where = "the BEM solve callback reported a failure";
    return false;
  }
  p->solves += 1;
  for (int k = 0; k < nd; ++k) {
    Device& dk = c->dev[k];
    PartDev& qk = p->dev[k];
    if (qk.n == 0 && k != 0) continue;
    O3D_TRY(dk, cudaSetDevice(dk.id));
    O3D_TRY(dk, cudaMemcpyAsync(body_row(qk.body, nn, np, 2), hs, (size_t)3 * np * 4, cudaMemcpyHostToDevice, dk.stream));
    O3D_TRY(dk, cudaMemcpyAsync(body_row(qk.body, nn, np, 4), hs + 3 * np, (size_t)np * 4, cudaMemcpyHostToDevice, dk.stream));
    if (!body_repack(dk, dk.stream, p, qk.body, have_source != 0)) return false;
  }
  return true;
}

bool launch_move(Device& d, cudaStream_t st, const MoveArgs& a, int order) {
  if (a.n == 0) return true;
  const unsigned gb = (unsigned)((a.n + 255) / 256);
  if (order == 1) pts_move_kernel<1><<<gb, 256, 0, st>>>(a);
  else if (order == 2) pts_move_kernel<2><<<gb, 256, 0, st>>>(a);
  else pts_move_kernel<3><<<gb, 256, 0, st>>>(a);
  O3D_TRY(d, cudaGetLastError());
  d.launches += 1;
  return true;
}

StageRef stage_of(const PartView& v) {
  StageRef s;
  for (int d = 0; d < 3; ++d) s.u[d] = v.u[d];
  s.ug = v.ug;
  s.ug_stride = v.stride;
  return s;
}

// `dst` <- Euler move of `src` over dt with velocity `vel`'s u and gradient `own`'s ug (src/Points.h:288-351: the
// one-stage move stretches with the moving object's OWN gradient). Elongation is only tracked on the main state.
MoveArgs euler_args(const PartDev& q, const PartView& src, const PartView& dst, const PartView& vel, const PartView& own, double dt) {
  MoveArgs a{};
  a.n = q.n;
  a.dt = dt;
  a.wt[0] = 1.0;
  a.st[0] = stage_of(vel);
  a.st[0].ug = own.ug;
  for (int d = 0; d < 3; ++d) {
    a.xin[d] = src.x[d]; a.sin[d] = src.s[d];
    a.xout[d] = dst.x[d]; a.sout[d] = dst.s[d];
    a.uout[d] = nullptr;
  }
  a.ein = nullptr; a.eout = nullptr;
  return a;
}

// One Convection::advect call (src/Convection.h:208-228) for one particle collection, optionally around a static body,
// enqueued on every device. With a body every derivative evaluation is find_derivs (:186-202) = solve the BEM for the state,
// then velocities from particles and panels; and every move is followed by clear_inner_layer on the moved state (:258,
// :372, :405, advect_3rd).
bool part_advect_once(o3d_ctx* c, o3d_particles* p, int order, double dt, const double* fs, bool in_capture) {
  const int nd = (int)c->dev.size();
  auto each = [&](auto&& f) {
    for (int k = 0; k < nd; ++k) {
      Device& d = c->dev[k];
      PartDev& q = p->dev[k];
      if (q.n == 0) continue;
      if (cudaSetDevice(d.id) != cudaSuccess) return false;
      if (!f(d, q)) return false;
    }
    return true;
  };
  auto clear = [&](int sel) { return !p->has_body || body_clear_inner(c, p, sel); };
  // find_derivs at the current state
  if (!part_find_vels(c, p, 0, fs, true, in_capture, true)) return false;
  if (order == 1) {
    // advect_1st :247-250 - elem.move(time, dt, 1.0, elem)
    return each([&](Device& d, PartDev& q) {
      PartView m = view_main(q);
      MoveArgs a = euler_args(q, m, m, m, m, dt);
      a.ein = a.eout = prow(q.main, q.cap, kRowE);
      return launch_move(d, d.stream, a, 1);
    }) && clear(0);
  }
  if (order == 2) {
    // advect_2nd_ralston :366-405 - interim = copy moved by 2/3 dt; derivatives there; combine 1/4, 3/4
    const double twothirds = 2.0 / 3.0;
    if (!each([&](Device& d, PartDev& q) {
          PartView m = view_main(q), i1 = view_interim(q, 0);
          return launch_move(d, d.stream, euler_args(q, m, i1, m, m, twothirds * dt), 1);
        }))
      return false;
    if (!clear(1)) return false;
    if (!part_find_vels(c, p, 1, fs, true, in_capture, true)) return false;
    return each([&](Device& d, PartDev& q) {
      PartView m = view_main(q), i1 = view_interim(q, 0);
      MoveArgs a{};
      a.n = q.n; a.dt = dt; a.wt[0] = 0.25; a.wt[1] = 0.75;
      a.st[0] = stage_of(m); a.st[1] = stage_of(i1);
      for (int k = 0; k < 3; ++k) { a.xin[k] = a.xout[k] = m.x[k]; a.sin[k] = a.sout[k] = m.s[k]; a.uout[k] = m.u[k]; }
      a.ein = a.eout = prow(q.main, q.cap, kRowE);
      return launch_move(d, d.stream, a, 2);
    }) && clear(0);
  }
  // advect_3rd :446-532 - vort1 = copy moved 1/2 dt with its own derivatives; vort2 = copy of the ORIGINAL moved
  // 3/4 dt with vort1's velocity (and, being a one-stage move, its own = the original's gradient); combine 2/9, 3/9, 4/9
  if (!each([&](Device& d, PartDev& q) {
        PartView m = view_main(q), i1 = view_interim(q, 0);
        return launch_move(d, d.stream, euler_args(q, m, i1, m, m, 0.5 * dt), 1);
      }))
    return false;
  if (!clear(1)) return false;
  if (!part_find_vels(c, p, 1, fs, true, in_capture, true)) return false;
  if (!each([&](Device& d, PartDev& q) {
        PartView m = view_main(q), i1 = view_interim(q, 0), i2 = view_interim(q, 1);
        return launch_move(d, d.stream, euler_args(q, m, i2, i1, m, 0.75 * dt), 1);
      }))
    return false;
  if (!clear(2)) return false;
  if (!part_find_vels(c, p, 2, fs, true, in_capture, true)) return false;
  return each([&](Device& d, PartDev& q) {
    PartView m = view_main(q), i1 = view_interim(q, 0), i2 = view_interim(q, 1);
    MoveArgs a{};
    a.n = q.n; a.dt = dt; a.wt[0] = 2.0 / 9.0; a.wt[1] = 3.0 / 9.0; a.wt[2] = 4.0 / 9.0;
    a.st[0] = stage_of(m); a.st[1] = stage_of(i1); a.st[2] = stage_of(i2);
    for (int k = 0; k < 3; ++k) { a.xin[k] = a.xout[k] = m.x[k]; a.sin[k] = a.sout[k] = m.s[k]; a.uout[k] = m.u[k]; }
    a.ein = a.eout = prow(q.main, q.cap, kRowE);
    return launch_move(d, d.stream, a, 3);
  }) && clear(0);
}

bool part_sync_all(o3d_ctx* c, o3d_particles* p) {
  for (size_t k = 0; k < c->dev.size(); ++k) {
    Device& d = c->dev[k];
    O3D_TRY(d, cudaSetDevice(d.id));
    O3D_TRY(d, cudaStreamSynchronize(d.stream));
  }
  return true;
}

void part_release_graph(PartDev& q) {
  if (q.graph) cudaGraphExecDestroy(q.graph);
  q.graph = nullptr;
  q.graph_n = -1;
}

// *_dev entry points launch on a caller stream: the context's first device must be the caller's current device
bool on_current_device(const o3d_ctx* c) {
  int cur = -1;
  return c && !c->dev.empty() && cudaGetDevice(&cur) == cudaSuccess && cur == c->dev[0].id;
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
  { const char* e = getenv("O3D_CUDA_SOURCES_PCIE_ALL"); c->sources_pcie_all = e && e[0] == '1'; }
  c->dev.resize(ndev);
  const char* env_tuned = getenv("O3D_CUDA_TUNED");          // "0": start with the linked copies of pp2_kernel
  const bool want_tuned = !(env_tuned && env_tuned[0] == '0');
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
    ok = ok && cudaStreamCreateWithFlags(&d.stream2, cudaStreamNonBlocking) == cudaSuccess;
    for (int e = 0; ok && e < 4; ++e) ok = cudaEventCreate(&d.ev[e]) == cudaSuccess;
    for (int e = 0; ok && e < 2; ++e) ok = cudaEventCreate(&d.evk[e]) == cudaSuccess;
    for (int e = 0; ok && e < 2; ++e) ok = cudaEventCreateWithFlags(&d.evx[e], cudaEventDisableTiming) == cudaSuccess;
    if (ok) {
      // helper threads for staging copies: O3D_CUDA_COPY_THREADS, default min(6, cores / devices) - 1 beside the caller
      const char* env_ct = getenv("O3D_CUDA_COPY_THREADS");
      int total = env_ct ? atoi(env_ct) : std::min(6, std::max(1, (int)std::thread::hardware_concurrency() / ndev));
      d.stage.pool = new CopyPool(std::max(0, std::min(total, 16) - 1));
    }
    if (!ok) {
      o3d_cuda_destroy(c);
      return O3D_ERR_CUDA;
    }
    d.tuned = want_tuned;
  }
  if (ndev > 1) {   // peers read each other's packed source stream over NVLink (host entry points, resident collections)
    for (int k = 0; k < ndev; ++k) {
      cudaSetDevice(c->dev[k].id);
      for (int j = 0; j < ndev; ++j) {
        int can = 0;
        if (j != k && cudaDeviceCanAccessPeer(&can, c->dev[k].id, c->dev[j].id) == cudaSuccess && can &&
            cudaDeviceEnablePeerAccess(c->dev[j].id, 0) != cudaSuccess)
          cudaGetLastError();   // already enabled is fine
      }
    }
  }
  if (tuned_kernels().status != cudaSuccess) {               // no silent fallback: the library is broken
    fprintf(stderr, "o3d_cuda_create: %s failed: %s\n", tuned_kernels().where, cudaGetErrorString(tuned_kernels().status));
    o3d_cuda_destroy(c);
    return O3D_ERR_CUDA;
  }
  *out = c;
  return O3D_OK;
}

void o3d_cuda_destroy(o3d_ctx* c) {
  if (!c) return;
  for (Device& d : c->dev) {
    cudaSetDevice(d.id);
    d.stage.release();
    d.acc.release();
    for (cudaEvent_t e : d.evx)
      if (e) cudaEventDestroy(e);
    if (d.stream2) cudaStreamDestroy(d.stream2);
    for (DevBuf* b : {&d.src, &d.packed, &d.targ, &d.out, &d.work, &d.ppwork, &d.geom, &d.panels, &d.tpanels, &d.cnt, &d.rng}) b->release();
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
    // src/Influence.h:310 (0pg: 68), :366 (0p: 31), :475 (0bg: 70), :534 (0b: 33) with the WL core
    const double per = pp_pair_flops(c->dev[0].core, grad, tr != nullptr);
    *flops_out = (double)nt * ((grad ? 12.0 : 3.0) + per * (double)ns);
  }
  for (Device& d : c->dev) d.kernel_ms = d.h2d_ms = d.d2h_ms = 0, d.launches = 0;
  if (ns == 0 || nt == 0) return collect(c);

  // Per device (one host thread each): its slice of the targets against all sources.
  //   sources  -> device 0 only (PCIe once), packed there; the other devices copy the PACKED stream from device 0 over NVLink
  //   targets  -> each device its slice
  //   kernel   -> stores FP64 sums (PPArgs::acc64) and never reads the outputs, so ...
  //   outputs  -> ... the caller's initial values are uploaded on a second stream WHILE the kernel runs, then one
  //               O(n) pass adds the sums into them with the rounding of the fused epilogue, and they travel back.
  const int ndev = (int)c->dev.size();
  const int64_t nrec = padded_sources(ns);
  std::atomic<int> packed_posted{0};        // 1: device 0 has recorded evx[0] after its pack; -1: it failed before that
  for_each_device(c, [&](int k) {
    Device& d = c->dev[k];
    auto bail = [&]() {
      if (k == 0 && packed_posted.load() == 0) packed_posted.store(-1);
      return false;
    };
    int64_t t0, t1;
    partition(nt, ndev, k, &t0, &t1);
    const int64_t n = t1 - t0;
    if (n == 0 && k != 0) return true;       // device 0 always packs: its peers wait for it
    if (cudaSetDevice(d.id) != cudaSuccess) { d.status = cudaGetLastError(); d.where = "cudaSetDevice"; return bail(); }
    const int nout = grad ? 12 : 3;
    const bool own_sources = k == 0 || c->sources_pcie_all;
    auto prep = [&]() {
      if (own_sources) O3D_TRY(d, d.src.ensure((size_t)7 * ns * 4));
      O3D_TRY(d, d.packed.ensure((size_t)nrec * 32));
      O3D_TRY(d, d.targ.ensure((size_t)4 * std::max<int64_t>(n, 1) * 4));
      O3D_TRY(d, d.out.ensure((size_t)nout * std::max<int64_t>(n, 1) * 4));
      O3D_TRY(d, d.acc.ensure((size_t)nout * std::max<int64_t>(n, 1) * sizeof(double)));
      return true;
    };
    if (!prep()) return bail();
    cudaStream_t st = d.stream;
    float* ds = d.src.as<float>();
    float* dt = d.targ.as<float>();
    float* dout = d.out.as<float>();
    auto upload_sources = [&]() {
      O3D_TRY(d, cudaEventRecord(d.ev[0], st));
      if (own_sources) {
        const float* hs[7] = {sx, sy, sz, sr, ssx, ssy, ssz};
        for (int a = 0; a < 7; ++a)
          if (!h2d(d, st, ds + (size_t)a * ns, hs[a], (size_t)ns * 4)) return false;
        if (!launch_pack(d, st, ns, ds, ds + ns, ds + 2 * ns, ds + 3 * ns, ds + 4 * ns, ds + 5 * ns, ds + 6 * ns, d.packed.as<float4>()))
          return false;
        if (ndev > 1 && k == 0) O3D_TRY(d, cudaEventRecord(d.evx[0], st));
      }
      return true;
    };
    const bool up = upload_sources();
    if (k == 0 && ndev > 1) packed_posted.store(up ? 1 : -1);
    if (!up) return bail();
    if (n == 0) {
      O3D_TRY(d, cudaStreamSynchronize(st));
      return true;
    }
    if (!own_sources) {
      int state;
      while ((state = packed_posted.load()) == 0) std::this_thread::yield();
      if (state < 0) return true;            // device 0 reports the error
      Device& d0 = c->dev[0];
      O3D_TRY(d, cudaStreamWaitEvent(st, d0.evx[0], 0));
      O3D_TRY(d, cudaMemcpyPeerAsync(d.packed.p, d.id, d0.packed.p, d0.id, (size_t)nrec * 32, st));
    }
    const float* ht[4] = {tx, ty, tz, tr};
    for (int a = 0; a < 4; ++a)
      if (ht[a] && !h2d(d, st, dt + (size_t)a * n, ht[a] + t0, (size_t)n * 4)) return false;
    O3D_TRY(d, cudaEventRecord(d.ev[1], st));
    if (!launch_pp(d, st, nrec, d.packed.as<float4>(), n, dt, dt + n, dt + 2 * n, tr ? dt + 3 * n : nullptr, nullptr, nullptr, nullptr,
                   nullptr, n, d.acc.as<double>(), n, grad ? 1 : 0))
      return false;
    // the caller's initial values go up while the kernel runs
    float* ho[12] = {tu, tv, tw};
    for (int a = 0; a < 9; ++a) ho[3 + a] = grad ? tug[a] : nullptr;
    for (int a = 0; a < nout; ++a)
      if (!h2d(d, d.stream2, dout + (size_t)a * n, ho[a] + t0, (size_t)n * 4)) return false;
    O3D_TRY(d, cudaEventRecord(d.evx[1], d.stream2));
    O3D_TRY(d, cudaStreamWaitEvent(st, d.evx[1], 0));
    pp_accumulate_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(nout, n, d.acc.as<double>(), n, dout, dout + n, dout + 2 * n,
                                                                     grad ? dout + 3 * n : nullptr, n, 1.0f);
    O3D_TRY(d, cudaGetLastError());
    d.launches += 1;
    O3D_TRY(d, cudaEventRecord(d.ev[2], st));
    for (int a = 0; a < nout; ++a)
      if (!d2h(d, st, ho[a] + t0, dout + (size_t)a * n, (size_t)n * 4)) return false;
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
  if (!h2d(d, st, gnx, nx, (size_t)nn * 4)) return false;
  if (!h2d(d, st, gny, ny, (size_t)nn * 4)) return false;
  if (!h2d(d, st, gnz, nz, (size_t)nn * 4)) return false;
  if (!h2d(d, st, gidx, idx, (size_t)3 * np * 4)) return false;
  const float* hts[3] = {tsx, tsy, tsz};
  for (int a = 0; a < 3; ++a)
    if (hts[a]) if (!h2d(d, st, gts + (size_t)a * np, hts[a], (size_t)np * 4)) return false;
  if (!h2d(d, st, garea, area, (size_t)np * 4)) return false;
  if (sss) if (!h2d(d, st, gsss, sss, (size_t)np * 4)) return false;
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
  if (!d2h(d, st, d.counts, d.cnt.p, 2 * sizeof(unsigned long long))) return false;
  return true;
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
    for (int a = 0; a < 3; ++a) if (!h2d(d, st, dt + (size_t)a * n, ht[a] + t0, (size_t)n * 4)) return false;
    float* ho[12] = {tu, tv, tw};
    for (int a = 0; a < 9; ++a) ho[3 + a] = grad ? tug[a] : nullptr;
    for (int a = 0; a < nout; ++a) if (!h2d(d, st, dout + (size_t)a * n, ho[a] + t0, (size_t)n * 4)) return false;
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
    launch_pan_pts(d, st, a, grid, grad);
    O3D_TRY(d, cudaGetLastError());
    d.launches += 1;
    if (a.nsplit > 1) {
      pp_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(nout, a.nsplit, n, a.partial, a.tu, a.tv, a.tw, a.tug, n, 1.0f);
      O3D_TRY(d, cudaGetLastError());
      d.launches += 1;
    }
    O3D_TRY(d, cudaEventRecord(d.ev[2], st));
    for (int a2 = 0; a2 < nout; ++a2) if (!d2h(d, st, ho[a2] + t0, dout + (size_t)a2 * n, (size_t)n * 4)) return false;
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
    *flops_out = (double)nt * (double)np * 4.0 + (leaves + splits) * 20.0 + leaves * leaf_flops(c->dev[0].core, grad) +
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
    for (int a = 0; a < 6; ++a) if (!h2d(d, st, ds + (size_t)a * ns, hs[a], (size_t)ns * 4)) return false;
    float* ho[3] = {pu, pv, pw};
    for (int a = 0; a < 3; ++a) if (!h2d(d, st, dout + (size_t)a * n, ho[a] + p0, (size_t)n * 4)) return false;
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
    if (d.pan_queue >= 2) pts_pan_queue_kernel<B><<<dim3((unsigned)gx, (unsigned)a.nsplit), B, 0, st>>>(a);
    else             pts_pan_kernel<B><<<dim3((unsigned)gx, (unsigned)a.nsplit), B, 0, st>>>(a);
    O3D_TRY(d, cudaGetLastError());
    d.launches += 1;
    pp_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(3, a.nsplit, n, a.partial, dout, dout + n, dout + 2 * n, nullptr, n, -1.0f);
    O3D_TRY(d, cudaGetLastError());
    d.launches += 1;
    O3D_TRY(d, cudaEventRecord(d.ev[2], st));
    for (int a2 = 0; a2 < 3; ++a2) if (!d2h(d, st, ho[a2] + p0, dout + (size_t)a2 * n, (size_t)n * 4)) return false;
    if (!fetch_counts(d, st)) return false;
    O3D_TRY(d, cudaEventRecord(d.ev[3], st));
    return finish_timing(d);
  });
  const int rc = collect(c);
  if (rc == O3D_OK && flops_out) {
    double leaves = 0, splits = 0;
    for (Device& d : c->dev) leaves += (double)d.counts[0], splits += (double)d.counts[1];
    *flops_out = (double)ns * (double)np * 4.0 + (leaves + splits) * 20.0 + leaves * leaf_flops(c->dev[0].core, false) +
                 splits * 23.0 + 3.0 * (double)np;
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
    if (!h2d(d, st, dsb1, sb1, (size_t)3 * nsp * 4)) return false;
    if (!h2d(d, st, dsb2, sb2, (size_t)3 * nsp * 4)) return false;
    if (!h2d(d, st, dtb1, tb1, (size_t)3 * ntp * 4)) return false;
    if (!h2d(d, st, dtb2, tb2, (size_t)3 * ntp * 4)) return false;
    if (!h2d(d, st, dtn, tnrm, (size_t)3 * ntp * 4)) return false;
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
    if (!d2h(d, st, coeffs + (size_t)3 * j0 * nrows, d.out.p, (size_t)3 * n * nrows * 4)) return false;
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

int o3d_cuda_plan_pts_on_pts(int sm_count, int64_t ns, int64_t nt, int want_grad, int64_t* grid, int* split_blocks,
                             double* balance, int64_t* workspace_bytes) {
  if (sm_count < 1 || ns < 1 || nt < 1) return O3D_ERR_INVALID;
  const bool grad = want_grad != 0;
  const int64_t ntiles = padded_sources(ns) / kTile;
  const PPShape s = pp_shape(sm_count, ntiles, nt, grad);
  if (grid) *grid = s.grid;
  if (split_blocks) *split_blocks = s.split_blocks;
  if (balance) {   // mean / max tiles per CTA: what fraction of the launch's duration the average CTA is busy
    int64_t most = 0;
    for (int c = 0; c < s.grid; ++c) most = std::max(most, s.plan.tiles_of(c));
    *balance = (double)s.units / ((double)most * s.grid);
  }
  if (workspace_bytes) *workspace_bytes = (int64_t)pp_workspace_bytes(sm_count);
  return O3D_OK;
}

// Host replay of one launch's bookkeeping, for tests without a device: walks every CTA's tiles exactly as the kernels do
// (pp_ring_start, pp2_walk) and every boundary exactly as pp_fixup_kernel does, and checks that each (target block,
// source tile) unit is consumed once, that every partial segment has its own workspace slot, and that every shared block is
// finished by exactly one fix-up CTA from exactly the slots that were written. Returns 0, or the number of the failed check.
int o3d_cuda_plan_check(int sm_count, int64_t ns, int64_t nt, int want_grad) {
  if (sm_count < 1 || ns < 1 || nt < 1) return -1;
  const bool grad = want_grad != 0;
  const int ntiles = (int)(padded_sources(ns) / kTile);
  const PPShape s = pp_shape(sm_count, ntiles, nt, grad);
  const PPPlan plan = s.plan;
  if (s.grid < 1 || s.grid > sm_count * (kPPWarpsPerSM * 32 / s.block) || plan.begin(0) != 0 || plan.begin(s.grid) != plan.Wt || plan.Pt > plan.P) return 1;
  if ((int64_t)plan.P * plan.full * ntiles + plan.Wt != s.units) return 1;
  std::vector<int> tiles_done(s.nblocks, 0), whole(s.nblocks, 0), fixed(s.nblocks, 0);
  std::vector<int> slot_block((size_t)s.grid * kPPSlots, -1), slot_tiles((size_t)s.grid * kPPSlots, 0), slot_read((size_t)s.grid * kPPSlots, 0);
  for (int c = 0; c < s.grid; ++c) {
    const int64_t u0 = plan.begin(c), u1 = plan.begin(c + 1);
    const int nA = plan.full * ntiles, nk = nA + (int)(u1 - u0);                              // pp_ring_start
    const int bA0 = c * plan.full, bB0 = plan.tail_block0() + (int)(u0 / ntiles), ktB0 = (int)(u0 % ntiles);
    if (nk < 1 || nk != plan.tiles_of(c)) return 2;
    int b = nA > 0 ? bA0 : bB0, kt = nA > 0 ? 0 : ktB0, kring = 0, seg_tiles = 0;             // pp_first_block
    bool seg_first = nA == 0;
    int tnext = nA > 0 ? 0 : ktB0, fetched = 0;                                               // the ring's prefetch order
    auto fetch = [&]() { const int t = tnext; tnext = fetched + 1 == nA ? ktB0 : (tnext + 1 == ntiles ? 0 : tnext + 1); ++fetched; return t; };
    int ring[2] = {-1, -1};
    for (int q = 0; q < 2 && q < nk; ++q) ring[q] = fetch();
    while (kring < nk) {                                                                      // pp2_walk / pp_segment_end
      if (b >= s.nblocks) return 3;
      if (ring[kring & 1] != kt) return 4;                                                    // the tile in the buffer is the unit's
      if (kring >= nA && (int64_t)(b - plan.tail_block0()) * ntiles + kt != u0 + (kring - nA)) return 4;
      if (kring < nA && (b != bA0 + kring / ntiles || kt != kring % ntiles)) return 4;
      if (kring + 2 < nk) ring[kring & 1] = fetch();
      ++kring; ++kt; ++seg_tiles;
      if (kt == ntiles || kring == nk) {
        const bool is_whole = kt == ntiles && (!seg_first || ktB0 == 0);
        if (is_whole != (seg_tiles == ntiles)) return 5;
        tiles_done[b] += seg_tiles;
        if (is_whole) whole[b] += 1;
        else {
          const size_t slot = (size_t)c * kPPSlots + (seg_first ? 0 : 1);
          if (slot_block[slot] != -1) return 6;                             // a slot is written once per launch
          slot_block[slot] = b;
          slot_tiles[slot] = seg_tiles;
        }
        seg_first = false;
        kt = 0;
        seg_tiles = 0;
        ++b;
        if (kring == nA) { b = bB0; kt = ktB0; seg_first = true; }
      }
    }
  }
  for (int j = 1; j < plan.Pt; ++j) {                                     // pp_fixup_kernel
    const int64_t cut = plan.begin(j);
    const int64_t bt = cut / ntiles, start = bt * ntiles, end = start + ntiles;
    if (cut == start) continue;
    const int64_t prev = plan.begin(j - 1);
    if (prev > start) continue;
    const int64_t b = plan.tail_block0() + bt;
    int got = 0;
    auto take = [&](size_t slot) {
      if (slot_block[slot] != (int)b) return false;
      got += slot_tiles[slot];
      slot_read[slot] += 1;
      return true;
    };
    if (!take((size_t)(j - 1) * kPPSlots + (prev == start ? 0 : 1))) return 7;
    for (int c = j; c < plan.Pt && plan.begin(c) < end; ++c)
      if (!take((size_t)c * kPPSlots)) return 8;
    if (got != ntiles) return 9;
    fixed[b] += 1;
  }
  int nsplit = 0;
  for (int b = 0; b < s.nblocks; ++b) {
    if (tiles_done[b] != ntiles) return 10;
    if (whole[b] + fixed[b] != 1) return 11;
    nsplit += fixed[b];
  }
  for (size_t k = 0; k < slot_block.size(); ++k)
    if ((slot_block[k] != -1) != (slot_read[k] == 1)) return 12;
  return nsplit == s.split_blocks ? 0 : 13;
}

int o3d_cuda_pack_sources_dev(o3d_ctx* c, void* stream, int64_t ns, const float* sx, const float* sy, const float* sz,
                              const float* sr, const float* ssx, const float* ssy, const float* ssz, int64_t nrec,
                              void* packed) {
  if (!c || ns < 0 || ns >= (int64_t(1) << 31)) return fail(c, O3D_ERR_INVALID, "pack_sources_dev: bad context or count");
  if (nrec != 0 && (nrec < padded_sources(ns) || nrec % kTile != 0)) return fail(c, O3D_ERR_INVALID, "pack_sources_dev: bad nrec");
  if (ns == 0 && nrec == 0) return O3D_OK;
  if (!packed || (ns > 0 && (!sx || !sy || !sz || !ssx || !ssy || !ssz))) return fail(c, O3D_ERR_INVALID, "pack_sources_dev: NULL array");
  if (!on_current_device(c)) return fail(c, O3D_ERR_INVALID, "pack_sources_dev: the context's device is not the current device");
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
  if (!on_current_device(c)) return fail(c, O3D_ERR_INVALID, "pts_on_pts_dev: the context's device is not the current device");
  Device& d = c->dev[0];
  d.launches = 0;
  launch_pp(d, (cudaStream_t)stream, nrec, (const float4*)packed, nt, tx, ty, tz, tr, tu, tv, tw, tug, tug_stride);
  return collect(c);
}


// ---------------------------------------------------------------------------------------------------
// Convection on the device
int o3d_cuda_pts_finalize_dev(o3d_ctx* c, void* stream, int64_t n, float* u, float* v, float* w, float* ug,
                              int64_t ug_stride, const double* fs) {
  if (!c || n < 0 || n >= (int64_t(1) << 31)) return fail(c, O3D_ERR_INVALID, "pts_finalize_dev: bad context or count");
  if (n == 0) return O3D_OK;
  if (!u || !v || !w || !fs || (ug && ug_stride < n)) return fail(c, O3D_ERR_INVALID, "pts_finalize_dev: NULL array");
  if (!on_current_device(c)) return fail(c, O3D_ERR_INVALID, "pts_finalize_dev: the context's device is not the current device");
  Device& d = c->dev[0];
  d.launches = 0;
  pts_finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, u, v, w, ug, ug_stride, fs[0], fs[1], fs[2]);
  d.status = cudaGetLastError();
  d.where = "pts_finalize_kernel";
  d.launches = 1;
  return collect(c);
}

int o3d_cuda_pts_move_dev(o3d_ctx* c, void* stream, int64_t n, int order, double dt, const double* wt,
                          const o3d_stage* stages, const float* const* xin, const float* const* sin, const float* ein,
                          float* const* xout, float* const* sout, float* eout, float* const* uout) {
  if (!c || n < 0 || n >= (int64_t(1) << 31) || order < 1 || order > 3)
    return fail(c, O3D_ERR_INVALID, "pts_move_dev: bad context, count or order");
  if (n == 0) return O3D_OK;
  if (!wt || !stages || !xin || !xout || (order > 1 && !uout) || ((sin == nullptr) != (sout == nullptr)) ||
      ((ein == nullptr) != (eout == nullptr)))
    return fail(c, O3D_ERR_INVALID, "pts_move_dev: NULL argument");
  MoveArgs a{};
  a.n = n;
  a.dt = dt;
  for (int k = 0; k < order; ++k) {
    a.wt[k] = wt[k];
    for (int d = 0; d < 3; ++d) {
      if (!stages[k].u[d]) return fail(c, O3D_ERR_INVALID, "pts_move_dev: NULL stage velocity");
      a.st[k].u[d] = stages[k].u[d];
    }
    if (stages[k].ug && stages[k].ug_stride < n) return fail(c, O3D_ERR_INVALID, "pts_move_dev: gradient stride < n");
    a.st[k].ug = stages[k].ug;
    a.st[k].ug_stride = stages[k].ug_stride;
  }
  for (int d = 0; d < 3; ++d) {
    if (!xin[d] || !xout[d] || (sin && (!sin[d] || !sout[d])) || (order > 1 && !uout[d]))
      return fail(c, O3D_ERR_INVALID, "pts_move_dev: NULL state array");
    a.xin[d] = xin[d]; a.xout[d] = xout[d];
    a.sin[d] = sin ? sin[d] : nullptr; a.sout[d] = sout ? sout[d] : nullptr;
    a.uout[d] = order > 1 ? uout[d] : nullptr;
  }
  a.ein = ein; a.eout = eout;
  if (!on_current_device(c)) return fail(c, O3D_ERR_INVALID, "pts_move_dev: the context's device is not the current device");
  Device& d = c->dev[0];
  d.launches = 0;
  launch_move(d, (cudaStream_t)stream, a, order);
  return collect(c);
}

int o3d_cuda_particles_create(o3d_ctx* c, o3d_particles** out) {
  if (!c || !out) return O3D_ERR_INVALID;
  o3d_particles* p = new o3d_particles();
  p->dev.resize(c->dev.size());
  *out = p;
  return O3D_OK;
}

void o3d_cuda_particles_destroy(o3d_ctx* c, o3d_particles* p) {
  if (!p) return;
  for (size_t k = 0; k < p->dev.size(); ++k) {
    if (c && k < c->dev.size()) cudaSetDevice(c->dev[k].id);
    PartDev& q = p->dev[k];
    part_release_graph(q);
    for (DevBuf* b : {&q.main, &q.interim[0], &q.interim[1], &q.packed, &q.stats, &q.totals, &q.body.geom, &q.body.panels, &q.body.refl,
                      &q.body.pu, &q.body.work, &q.body.cnt})
      b->release();
    for (cudaEvent_t e : {q.packed_ready, q.pulled, q.ev[0], q.ev[1]})
      if (e) cudaEventDestroy(e);
  }
  delete p;
}

int64_t o3d_cuda_particles_count(const o3d_particles* p) { return p ? p->n : 0; }

int o3d_cuda_particles_upload(o3d_ctx* c, o3d_particles* p, int64_t n, const float* x, const float* y, const float* z,
                              const float* sx, const float* sy, const float* sz, const float* r, const float* elong) {
  if (!c || !p || n < 0 || n >= (int64_t(1) << 31) || p->dev.size() != c->dev.size())
    return fail(c, O3D_ERR_INVALID, "particles_upload: bad context, collection or count");
  if (n > 0 && (!x || !y || !z || !sx || !sy || !sz || !r)) return fail(c, O3D_ERR_INVALID, "particles_upload: NULL array");
  for (Device& d : c->dev) d.kernel_ms = d.h2d_ms = d.d2h_ms = 0, d.launches = 0;
  if (!part_layout(c, p, n)) return collect(c);
  const float* rows[8] = {x, y, z, sx, sy, sz, r, elong};
  for (size_t k = 0; k < c->dev.size(); ++k) {
    Device& d = c->dev[k];
    PartDev& q = p->dev[k];
    if (q.n == 0) continue;
    auto go = [&]() {
      O3D_TRY(d, cudaSetDevice(d.id));
      for (int a = 0; a < 8; ++a) {
        float* dst = prow(q.main, q.cap, a);
        if (rows[a]) if (!h2d(d, d.stream, dst, rows[a] + q.t0, (size_t)q.n * 4)) return false;
      }
      if (!elong) {   // a fresh collection: elong = 1 (src/Points.h:120-127)
        pts_fill_kernel<<<(unsigned)((q.n + 255) / 256), 256, 0, d.stream>>>(q.n, 1, prow(q.main, q.cap, kRowE), q.cap, 1.0f);
        O3D_TRY(d, cudaGetLastError());
        d.launches += 1;
      }
      pts_fill_kernel<<<(unsigned)((q.n + 255) / 256), 256, 0, d.stream>>>(q.n, 12, prow(q.main, q.cap, kRowU), q.cap, 0.0f);
      O3D_TRY(d, cudaGetLastError());
      d.launches += 1;
      O3D_TRY(d, cudaStreamSynchronize(d.stream));
      return true;
    };
    if (!go()) break;
  }
  return collect(c);
}

int o3d_cuda_particles_download(o3d_ctx* c, o3d_particles* p, float* x, float* y, float* z, float* sx, float* sy, float* sz,
                                float* r, float* elong, float* u, float* v, float* w, float* const* ug) {
  if (!c || !p || p->dev.size() != c->dev.size()) return fail(c, O3D_ERR_INVALID, "particles_download: bad context or collection");
  for (Device& d : c->dev) d.kernel_ms = d.h2d_ms = d.d2h_ms = 0, d.launches = 0;
  float* rows[20] = {x, y, z, sx, sy, sz, r, elong, u, v, w};
  for (int a = 0; a < 9; ++a) rows[11 + a] = ug ? ug[a] : nullptr;
  for (size_t k = 0; k < c->dev.size(); ++k) {
    Device& d = c->dev[k];
    PartDev& q = p->dev[k];
    if (q.n == 0) continue;
    auto go = [&]() {
      O3D_TRY(d, cudaSetDevice(d.id));
      for (int a = 0; a < 20; ++a)
        if (rows[a]) if (!d2h(d, d.stream, rows[a] + q.t0, prow(q.main, q.cap, a), (size_t)q.n * 4)) return false;
      O3D_TRY(d, cudaStreamSynchronize(d.stream));
      return true;
    };
    if (!go()) break;
  }
  return collect(c);
}

int o3d_cuda_particles_find_vels(o3d_ctx* c, o3d_particles* p, const double* fs, int want_grad, double* flops_out) {
  if (!c || !p || !fs || p->dev.size() != c->dev.size()) return fail(c, O3D_ERR_INVALID, "particles_find_vels: bad argument");
  for (Device& d : c->dev) d.kernel_ms = d.h2d_ms = d.d2h_ms = 0, d.launches = 0;
  const double n = (double)p->n;
  if (flops_out)   // src/Influence.h:475,534
    *flops_out = want_grad ? n * (12.0 + pp_pair_flops(c->dev[0].core, true, true) * n) : n * (3.0 + pp_pair_flops(c->dev[0].core, false, true) * n);
  if (p->n == 0) return collect(c);
  if (part_find_vels(c, p, 0, fs, want_grad != 0, false)) part_sync_all(c, p);
  return collect(c);
}

int o3d_cuda_particles_advect(o3d_ctx* c, o3d_particles* p, int order, double time, double dt, const double* fs, int nsteps,
                              double* flops_out) {
  (void)time;   // no time-dependent boundary motion in a particle-only system
  if (!c || !p || !fs || order < 1 || order > 3 || nsteps < 0 || p->dev.size() != c->dev.size())
    return fail(c, O3D_ERR_INVALID, "particles_advect: bad argument");
  for (Device& d : c->dev) d.kernel_ms = d.h2d_ms = d.d2h_ms = 0, d.launches = 0;
  const double n = (double)p->n;
  if (flops_out) *flops_out = (double)nsteps * order * n * (12.0 + pp_pair_flops(c->dev[0].core, true, true) * n);
  p->solves = 0;
  p->moved = 0;
  if (p->n == 0 || nsteps == 0) return collect(c);
  if (p->has_body) {
    for (size_t k = 0; k < c->dev.size(); ++k) {
      if (p->dev[k].n == 0) continue;
      cudaSetDevice(c->dev[k].id);
      cudaMemsetAsync(p->dev[k].body.cnt.p, 0, 4 * sizeof(unsigned long long), c->dev[k].stream);
    }
  }
  const int nd = (int)c->dev.size();
  auto run = [&]() {
    for (int k = 0; k < nd; ++k) {
      Device& d = c->dev[k];
      PartDev& q = p->dev[k];
      O3D_TRY(d, cudaSetDevice(d.id));
      for (int i = 0; i + 1 < order; ++i) O3D_TRY(d, q.interim[i].ensure((size_t)kRowsInterim * q.cap * 4));
      O3D_TRY(d, cudaEventRecord(q.ev[0], d.stream));
    }
    int done = 0;
    // Single-device contexts replay one captured CUDA graph per step (the step is ~20 small launches around
    // 1-3 big ones: launch-bound for small collections). The first step always runs eagerly so that every
    // grow-only buffer has its final size before capture.
    Device& d0 = c->dev[0];
    PartDev& q0 = p->dev[0];
    const bool same = c->use_graphs && !p->has_body && q0.graph && q0.graph_n == p->n && q0.graph_order == order && q0.graph_dt == dt &&
                      q0.graph_core == d0.core && q0.graph_tuned == d0.tuned &&
                      q0.graph_fs[0] == fs[0] && q0.graph_fs[1] == fs[1] && q0.graph_fs[2] == fs[2];
    if (!same) part_release_graph(q0);
    int per_step = 0;
    if (!same) {
      const int before = d0.launches;
      if (!part_advect_once(c, p, order, dt, fs, false)) return false;
      per_step = d0.launches - before;
      done = 1;
    }
    if (nd == 1 && c->use_graphs && !p->has_body && done < nsteps && !q0.graph_failed) {   // (a body's solve is a host call)
      if (!q0.graph) {
        const bool prof = d0.profile;
        d0.profile = false;
        cudaGraph_t g = nullptr;
        const int before = d0.launches;
        bool ok = cudaStreamBeginCapture(d0.stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
          const bool enq = part_advect_once(c, p, order, dt, fs, true);
          ok = cudaStreamEndCapture(d0.stream, &g) == cudaSuccess && enq && g;
        }
        per_step = d0.launches - before;
        d0.launches = before;
        if (ok) ok = cudaGraphInstantiate(&q0.graph, g, 0) == cudaSuccess;
        if (g) cudaGraphDestroy(g);
        d0.profile = prof;
        if (!ok) {   // still the GPU path: just launch-by-launch
          cudaGetLastError();
          d0.status = cudaSuccess;
          q0.graph = nullptr;
          q0.graph_failed = true;
        } else {
          q0.graph_n = p->n; q0.graph_order = order; q0.graph_dt = dt;
          q0.graph_core = d0.core; q0.graph_tuned = d0.tuned;
          for (int a = 0; a < 3; ++a) q0.graph_fs[a] = fs[a];
          q0.launches_per_step = per_step;
        }
      }
      if (q0.graph) {
        for (; done < nsteps; ++done) {
          O3D_TRY(d0, cudaGraphLaunch(q0.graph, d0.stream));
          d0.launches += q0.launches_per_step;
        }
      }
    }
    for (; done < nsteps; ++done)
      if (!part_advect_once(c, p, order, dt, fs, false)) return false;
    for (int k = 0; k < nd; ++k) {
      Device& d = c->dev[k];
      PartDev& q = p->dev[k];
      O3D_TRY(d, cudaSetDevice(d.id));
      O3D_TRY(d, cudaEventRecord(q.ev[1], d.stream));
    }
    if (!part_sync_all(c, p)) return false;
    for (int k = 0; k < nd; ++k) {
      Device& d = c->dev[k];
      O3D_TRY(d, cudaEventElapsedTime(&d.kernel_ms, p->dev[k].ev[0], p->dev[k].ev[1]));
      if (p->has_body && p->dev[k].n > 0) {
        unsigned long long m = 0;
        O3D_TRY(d, cudaSetDevice(d.id));
        O3D_TRY(d, cudaMemcpy(&m, p->dev[k].body.cnt.as<unsigned long long>() + 2, sizeof m, cudaMemcpyDeviceToHost));
        p->moved += (int64_t)m;
      }
    }
    return true;
  };
  run();
  return collect(c);
}

int o3d_cuda_particles_stats(o3d_ctx* c, o3d_particles* p, float* max_str, float* max_elong) {
  if (!c || !p || p->dev.size() != c->dev.size()) return fail(c, O3D_ERR_INVALID, "particles_stats: bad argument");
  for (Device& d : c->dev) d.kernel_ms = d.h2d_ms = d.d2h_ms = 0, d.launches = 0;
  float ms = 0.0f, me = 0.0f;
  for (size_t k = 0; k < c->dev.size(); ++k) {
    Device& d = c->dev[k];
    PartDev& q = p->dev[k];
    if (q.n == 0) continue;
    uint32_t h[2] = {0, 0};
    auto go = [&]() {
      O3D_TRY(d, cudaSetDevice(d.id));
      O3D_TRY(d, cudaMemsetAsync(q.stats.p, 0, 2 * sizeof(uint32_t), d.stream));
      pts_stats_kernel<<<d.sm_count * 2, 256, 0, d.stream>>>(q.n, prow(q.main, q.cap, kRowS), prow(q.main, q.cap, kRowS + 1),
                                                             prow(q.main, q.cap, kRowS + 2), prow(q.main, q.cap, kRowE), q.stats.as<uint32_t>());
      O3D_TRY(d, cudaGetLastError());
      d.launches += 1;
      if (!d2h(d, d.stream, h, q.stats.p, sizeof h)) return false;
      O3D_TRY(d, cudaStreamSynchronize(d.stream));
      return true;
    };
    if (!go()) break;
    float a, b;
    std::memcpy(&a, &h[0], 4);
    std::memcpy(&b, &h[1], 4);
    ms = std::max(ms, a);
    me = std::max(me, b);
  }
  if (max_str) *max_str = std::sqrt(ms);   // ElementBase::get_max_str returns sqrt of the largest |s|^2
  if (max_elong) *max_elong = me;
  return collect(c);
}

// ---- a static body attached to a resident collection (SURVEY.md 8 f: Convection::find_vels / advect with boundaries) ----
int o3d_cuda_particles_set_body(o3d_ctx* c, o3d_particles* p, int64_t nn, const float* nx, const float* ny, const float* nz, int64_t np,
                                const uint32_t* idx, const float* area, const float* nrm, float cutoff_mult, float ips,
                                o3d_bem_solve_fn solve, void* user) {
  if (!c || !p || p->dev.size() != c->dev.size() || nn < 1 || np < 1 || nn >= (int64_t(1) << 31) || np >= (int64_t(1) << 31))
    return fail(c, O3D_ERR_INVALID, "particles_set_body: bad argument");
  if (!nx || !ny || !nz || !idx || !area || !nrm) return fail(c, O3D_ERR_INVALID, "particles_set_body: NULL array");
  if (!valid_indices(idx, np, nn)) return fail(c, O3D_ERR_INVALID, "particles_set_body: node index out of range");
  for (Device& d : c->dev) d.kernel_ms = d.h2d_ms = d.d2h_ms = 0, d.launches = 0;
  const int64_t npad = padded_panels(np), rpad = ((np + kRefTile - 1) / kRefTile) * kRefTile;
  p->b_nn = nn; p->b_np = np;
  p->solve = solve; p->solve_user = user;
  p->cutoff_mult = cutoff_mult; p->ips = ips;
  for (size_t k = 0; k < c->dev.size(); ++k) {
    Device& d = c->dev[k];
    BodyDev& b = p->dev[k].body;
    part_release_graph(p->dev[k]);
    auto go = [&]() {
      O3D_TRY(d, cudaSetDevice(d.id));
      cudaStream_t st = d.stream;
      O3D_TRY(d, b.geom.ensure(((size_t)3 * nn + (size_t)11 * np) * 4));
      O3D_TRY(d, b.panels.ensure((size_t)npad * kPanRec * sizeof(float4)));
      O3D_TRY(d, b.refl.ensure((size_t)rpad * kRefRec * sizeof(float4)));
      O3D_TRY(d, b.pu.ensure((size_t)3 * np * 4));
      O3D_TRY(d, b.cnt.ensure(4 * sizeof(unsigned long long)));
      float* gx = body_row(b, nn, np, 0);
      O3D_TRY(d, cudaMemcpyAsync(gx, nx, (size_t)nn * 4, cudaMemcpyHostToDevice, st));
      O3D_TRY(d, cudaMemcpyAsync(gx + nn, ny, (size_t)nn * 4, cudaMemcpyHostToDevice, st));
      O3D_TRY(d, cudaMemcpyAsync(gx + 2 * nn, nz, (size_t)nn * 4, cudaMemcpyHostToDevice, st));
      O3D_TRY(d, cudaMemcpyAsync(body_row(b, nn, np, 1), idx, (size_t)3 * np * 4, cudaMemcpyHostToDevice, st));
      O3D_TRY(d, cudaMemsetAsync(body_row(b, nn, np, 2), 0, (size_t)3 * np * 4, st));        // strengths start at zero
      O3D_TRY(d, cudaMemcpyAsync(body_row(b, nn, np, 3), area, (size_t)np * 4, cudaMemcpyHostToDevice, st));
      O3D_TRY(d, cudaMemsetAsync(body_row(b, nn, np, 4), 0, (size_t)np * 4, st));
      O3D_TRY(d, cudaMemcpyAsync(body_row(b, nn, np, 5), nrm, (size_t)3 * np * 4, cudaMemcpyHostToDevice, st));
      O3D_TRY(d, cudaMemsetAsync(b.cnt.p, 0, 4 * sizeof(unsigned long long), st));
      if (!body_repack(d, st, p, b, false)) return false;
      ref_pack_kernel<<<(unsigned)((rpad + 127) / 128), 128, 0, st>>>(np, rpad, gx, gx + nn, gx + 2 * nn, (const uint32_t*)body_row(b, nn, np, 1),
                                                                     body_row(b, nn, np, 5), b.refl.as<float4>());
      O3D_TRY(d, cudaGetLastError());
      d.launches += 1;
      O3D_TRY(d, cudaStreamSynchronize(st));
      return true;
    };
    if (!go()) break;
  }
  const int rc = collect(c);
  p->has_body = rc == O3D_OK;
  return rc;
}

int o3d_cuda_particles_clear_body(o3d_ctx* c, o3d_particles* p) {
  if (!c || !p || p->dev.size() != c->dev.size()) return fail(c, O3D_ERR_INVALID, "particles_clear_body: bad argument");
  p->has_body = false;
  p->solve = nullptr;
  for (size_t k = 0; k < p->dev.size(); ++k) {
    cudaSetDevice(c->dev[k].id);
    BodyDev& b = p->dev[k].body;
    for (DevBuf* q : {&b.geom, &b.panels, &b.refl, &b.pu, &b.work, &b.cnt}) q->release();
  }
  return O3D_OK;
}

int o3d_cuda_particles_set_body_strengths(o3d_ctx* c, o3d_particles* p, const float* tsx, const float* tsy, const float* tsz, const float* sss) {
  if (!c || !p || p->dev.size() != c->dev.size() || !p->has_body || !tsx || !tsy || !tsz)
    return fail(c, O3D_ERR_INVALID, "particles_set_body_strengths: bad argument or no body attached");
  for (Device& d : c->dev) d.kernel_ms = d.h2d_ms = d.d2h_ms = 0, d.launches = 0;
  const int64_t nn = p->b_nn, np = p->b_np;
  for (size_t k = 0; k < c->dev.size(); ++k) {
    Device& d = c->dev[k];
    BodyDev& b = p->dev[k].body;
    auto go = [&]() {
      O3D_TRY(d, cudaSetDevice(d.id));
      cudaStream_t st = d.stream;
      float* ts = body_row(b, nn, np, 2);
      O3D_TRY(d, cudaMemcpyAsync(ts, tsx, (size_t)np * 4, cudaMemcpyHostToDevice, st));
      O3D_TRY(d, cudaMemcpyAsync(ts + np, tsy, (size_t)np * 4, cudaMemcpyHostToDevice, st));
      O3D_TRY(d, cudaMemcpyAsync(ts + 2 * np, tsz, (size_t)np * 4, cudaMemcpyHostToDevice, st));
      if (sss) O3D_TRY(d, cudaMemcpyAsync(body_row(b, nn, np, 4), sss, (size_t)np * 4, cudaMemcpyHostToDevice, st));
      if (!body_repack(d, st, p, b, sss != nullptr)) return false;
      O3D_TRY(d, cudaStreamSynchronize(st));
      return true;
    };
    if (!go()) break;
  }
  return collect(c);
}

// The raw panel-centre sums of the CURRENT state without solving: zero, particles -> panels (what solve_bem holds before finalize_vels)
int o3d_cuda_particles_body_vels(o3d_ctx* c, o3d_particles* p, float* pu, float* pv, float* pw) {
  const double fs[3] = {0.0, 0.0, 0.0};
  if (!c || !p || !pu || !pv || !pw || p->dev.size() != c->dev.size() || !p->has_body)
    return fail(c, O3D_ERR_INVALID, "particles_body_vels: bad argument or no body attached");
  for (Device& d : c->dev) d.kernel_ms = d.h2d_ms = d.d2h_ms = 0, d.launches = 0;
  if (p->n == 0) return collect(c);
  struct Grab { float* out[3]; int64_t np; };
  Grab g{{pu, pv, pw}, p->b_np};
  const o3d_bem_solve_fn keep = p->solve;
  void* keep_user = p->solve_user;
  p->solve = [](void* user, int64_t np, const float* v, float*, float*, float*, float*, int*) {
    Grab* q = (Grab*)user;
    for (int a = 0; a < 3; ++a) std::memcpy(q->out[a], v + (size_t)a * np, (size_t)np * 4);
    return 1;     // "do not touch the strengths": reported as a refusal, recognised below
  };
  p->solve_user = &g;
  // pack + exchange the main state, then the right-hand side; the evaluation that normally follows is skipped
  const int nd = (int)c->dev.size();
  bool ok = true;
  for (int k = 0; k < nd && ok; ++k) {
    Device& d = c->dev[k];
    PartDev& q = p->dev[k];
    if (q.n == 0) continue;
    auto go = [&]() {
      O3D_TRY(d, cudaSetDevice(d.id));
      PartView v = view_main(q);
      float4* slice = q.packed.as<float4>() + (size_t)q.t0 * 2;
      if (!launch_pack(d, d.stream, q.n, v.x[0], v.x[1], v.x[2], v.r, v.s[0], v.s[1], v.s[2], slice, padded_sources(q.n))) return false;
      if (nd > 1) O3D_TRY(d, cudaEventRecord(q.packed_ready, d.stream));
      return true;
    };
    ok = go();
  }
  if (ok && nd > 1) {
    Device& d = c->dev[0];
    PartDev& q = p->dev[0];
    auto pull = [&]() {
      O3D_TRY(d, cudaSetDevice(d.id));
      for (int j = 1; j < nd; ++j) {
        PartDev& o = p->dev[j];
        if (o.n == 0) continue;
        O3D_TRY(d, cudaStreamWaitEvent(d.stream, o.packed_ready, 0));
        const size_t off = (size_t)o.t0 * 32, bytes = (size_t)padded_sources(o.n) * 32;
        O3D_TRY(d, cudaMemcpyPeerAsync((char*)q.packed.p + off, d.id, (const char*)o.packed.p + off, c->dev[j].id, bytes, d.stream));
      }
      return true;
    };
    ok = pull();
  }
  if (ok) {
    body_solve(c, p, fs);
    Device& d0 = c->dev[0];
    if (d0.status == cudaErrorUnknown) d0.status = cudaSuccess;     // the grabber's "refusal" is not an error
  }
  p->solve = keep;
  p->solve_user = keep_user;
  if (ok) part_sync_all(c, p);
  return collect(c);
}

int o3d_cuda_particles_clear_inner(o3d_ctx* c, o3d_particles* p, int64_t* num_moved) {
  if (!c || !p || p->dev.size() != c->dev.size() || !p->has_body) return fail(c, O3D_ERR_INVALID, "particles_clear_inner: bad argument or no body attached");
  for (Device& d : c->dev) d.kernel_ms = d.h2d_ms = d.d2h_ms = 0, d.launches = 0;
  if (num_moved) *num_moved = 0;
  if (p->n == 0) return collect(c);
  for (size_t k = 0; k < c->dev.size(); ++k) {
    if (p->dev[k].n == 0) continue;
    cudaSetDevice(c->dev[k].id);
    cudaMemsetAsync(p->dev[k].body.cnt.p, 0, 4 * sizeof(unsigned long long), c->dev[k].stream);
  }
  if (body_clear_inner(c, p, 0) && part_sync_all(c, p)) {
    for (size_t k = 0; k < c->dev.size(); ++k) {
      if (p->dev[k].n == 0) continue;
      unsigned long long m = 0;
      cudaSetDevice(c->dev[k].id);
      if (cudaMemcpy(&m, p->dev[k].body.cnt.as<unsigned long long>() + 2, sizeof m, cudaMemcpyDeviceToHost) == cudaSuccess && num_moved) *num_moved += (int64_t)m;
    }
  }
  return collect(c);
}

int o3d_cuda_particles_body_counters(const o3d_particles* p, int64_t* moved, int* solves) {
  if (!p) return O3D_ERR_INVALID;
  if (moved) *moved = p->moved;
  if (solves) *solves = p->solves;
  return O3D_OK;
}

// Status-file quantities of a resident collection: total circulation and linear impulse
int o3d_cuda_particles_totals(o3d_ctx* c, o3d_particles* p, double* circ, double* impulse) {
  if (!c || !p || p->dev.size() != c->dev.size()) return fail(c, O3D_ERR_INVALID, "particles_totals: bad argument");
  for (Device& d : c->dev) d.kernel_ms = d.h2d_ms = d.d2h_ms = 0, d.launches = 0;
  double tot[6] = {0, 0, 0, 0, 0, 0};
  constexpr int kBlocks = 296;   // fixed, not the SM count: the shape of the summation tree is part of the result
  for (size_t k = 0; k < c->dev.size(); ++k) {       // devices in order: the collection's blocks in particle order
    Device& d = c->dev[k];
    PartDev& q = p->dev[k];
    if (q.n == 0) continue;
    double h[6];
    auto go = [&]() {
      O3D_TRY(d, cudaSetDevice(d.id));
      O3D_TRY(d, q.totals.ensure((size_t)(kBlocks + 1) * 6 * sizeof(double)));
      double* part = q.totals.as<double>();
      pts_totals_kernel<<<kBlocks, kTotalsBlock, 0, d.stream>>>(q.n, prow(q.main, q.cap, kRowX), prow(q.main, q.cap, kRowX + 1),
                                                                prow(q.main, q.cap, kRowX + 2), prow(q.main, q.cap, kRowS),
                                                                prow(q.main, q.cap, kRowS + 1), prow(q.main, q.cap, kRowS + 2), part);
      pts_totals_finish_kernel<<<1, 32, 0, d.stream>>>(kBlocks, part, part + (size_t)kBlocks * 6);
      O3D_TRY(d, cudaGetLastError());
      d.launches += 2;
      if (!d2h(d, d.stream, h, part + (size_t)kBlocks * 6, sizeof h)) return false;
      O3D_TRY(d, cudaStreamSynchronize(d.stream));
      return true;
    };
    if (!go()) break;
    for (int a = 0; a < 6; ++a) tot[a] += h[a];
  }
  for (int a = 0; a < 3; ++a) {
    if (circ) circ[a] = tot[a];
    if (impulse) impulse[a] = tot[3 + a];
  }
  return collect(c);
}

// ---- status file (csrc/status_writer.h) ---------------------------------------------------------------------
int o3d_cuda_status_open(const char* path, int csv, o3d_status** out) {
  if (!path || !path[0] || !out) return O3D_ERR_INVALID;
  o3d_status* st = new o3d_status();
  st->fn = path;
  st->csv = csv != 0;
  *out = st;
  return O3D_OK;
}
void o3d_cuda_status_close(o3d_status* st) { delete st; }
int o3d_cuda_status_reset_sim(o3d_status* st) {
  if (!st) return O3D_ERR_INVALID;
  st->reset_sim();
  return O3D_OK;
}
int o3d_cuda_status_append_float(o3d_status* st, const char* name, float value) {
  if (!st) return O3D_ERR_INVALID;
  st->append(name ? name : "float", value);       // StatusFile::append_value(float) names the column "float"
  return O3D_OK;
}
int o3d_cuda_status_append_int(o3d_status* st, const char* name, int value) {
  if (!st) return O3D_ERR_INVALID;
  st->append(name ? name : "int", value);
  return O3D_OK;
}
int o3d_cuda_status_write_line(o3d_status* st) {
  if (!st) return O3D_ERR_INVALID;
  return st->write_line() ? O3D_OK : O3D_ERR_INVALID;
}

// Simulation::dump_stats_to_status (src/Simulation.cpp:851-897) for a system that is one resident particle collection:
// time, Nv, total circulation, and the force estimate of calculate_simple_forces (:900-924) - the time derivative of the
// total impulse by a one-sided difference whose previous sample lives in the status object.
int o3d_cuda_particles_write_status(o3d_ctx* c, o3d_particles* p, o3d_status* st, double time, double dt) {
  if (!c || !p || !st) return fail(c, O3D_ERR_INVALID, "particles_write_status: bad argument");
  double circ[3], imp[3];
  const int rc = o3d_cuda_particles_totals(c, p, circ, imp);
  if (rc != O3D_OK) return rc;
  st->append("time", (float)time);
  st->append("Nv", (int)p->n);
  st->append("gx", (float)circ[0]);
  st->append("gy", (float)circ[1]);
  st->append("gz", (float)circ[2]);
  if (time < 0.1 * dt) {           // a new run: the "last" sample is zero impulse one step before the start
    st->last_time = -dt;
    st->last_impulse[0] = st->last_impulse[1] = st->last_impulse[2] = 0.0f;
  }
  float force[3];
  for (int a = 0; a < 3; ++a) {
    const float now = (float)imp[a];
    force[a] = (float)((now - st->last_impulse[a]) / (time - st->last_time));   // float difference over a double interval
    st->last_impulse[a] = now;
  }
  st->last_time = time;
  st->append("fx", force[0]);
  st->append("fy", force[1]);
  st->append("fz", force[2]);
  if (!st->write_line()) return fail(c, O3D_ERR_INVALID, "particles_write_status: cannot write " + st->fn);
  return O3D_OK;
}

// ---------------------------------------------------------------------------------------------------
// Matrix-free BEM operator: y = A x with A the panels_on_panels_coeff block
int o3d_cuda_bem_op_create(o3d_ctx* c, int64_t snn, const float* snx, const float* sny, const float* snz, int64_t nsp,
                           const uint32_t* sidx, const float* sb1, const float* sb2, const float* sarea, int64_t tnn,
                           const float* tnx, const float* tny, const float* tnz, int64_t ntp, const uint32_t* tidx,
                           const float* tb1, const float* tb2, const float* tnrm, const float* tarea, int self,
                           o3d_bem_op** out) {
  if (!out) return fail(c, O3D_ERR_INVALID, "bem_op_create: NULL output");
  *out = nullptr;
  if (!check_counts(c, nsp, ntp) || nsp == 0 || ntp == 0 || snn < 0 || tnn < 0 || snn >= (int64_t(1) << 31) || tnn >= (int64_t(1) << 31))
    return fail(c, O3D_ERR_INVALID, "bem_op_create: bad context or counts");
  if (!snx || !sny || !snz || !sidx || !sb1 || !sb2 || !sarea || !tnx || !tny || !tnz || !tidx || !tb1 || !tb2 || !tnrm || !tarea)
    return fail(c, O3D_ERR_INVALID, "bem_op_create: NULL array");
  if (self && nsp != ntp) return fail(c, O3D_ERR_INVALID, "bem_op_create: self block must be square");
  if (!valid_indices(sidx, nsp, snn) || !valid_indices(tidx, ntp, tnn)) return fail(c, O3D_ERR_INVALID, "bem_op_create: node index out of range");
  for (Device& d : c->dev) d.kernel_ms = d.h2d_ms = d.d2h_ms = 0, d.launches = 0;
  o3d_bem_op* op = new o3d_bem_op();
  op->nsp = nsp; op->ntp = ntp; op->self = self;
  op->dev.resize(c->dev.size());
  const int ndev = (int)c->dev.size();
  for (int k = 0; k < ndev; ++k) {
    Device& d = c->dev[k];
    BemDev& q = op->dev[k];
    int64_t i1;
    partition(ntp, ndev, k, &q.i0, &i1);
    q.ni = i1 - q.i0;
    if (q.ni == 0) continue;
    auto go = [&]() {
      O3D_TRY(d, cudaSetDevice(d.id));
      cudaStream_t st = d.stream;
      O3D_TRY(d, q.bases.ensure(((size_t)6 * nsp + (size_t)9 * ntp) * 4));
      O3D_TRY(d, q.x.ensure((size_t)3 * nsp * 4));
      O3D_TRY(d, q.y.ensure((size_t)3 * q.ni * 4));
      O3D_TRY(d, q.cnt.ensure(2 * sizeof(unsigned long long)));
      float* db = q.bases.as<float>();
      if (!h2d(d, st, db, sb1, (size_t)3 * nsp * 4)) return false;
      if (!h2d(d, st, db + 3 * nsp, sb2, (size_t)3 * nsp * 4)) return false;
      if (!h2d(d, st, db + 6 * nsp, tb1, (size_t)3 * ntp * 4)) return false;
      if (!h2d(d, st, db + 6 * nsp + 3 * ntp, tb2, (size_t)3 * ntp * 4)) return false;
      if (!h2d(d, st, db + 6 * nsp + 6 * ntp, tnrm, (size_t)3 * ntp * 4)) return false;
      if (!upload_panels(d, st, q.spanels, snn, snx, sny, snz, nsp, sidx, nullptr, nullptr, nullptr, sarea, nullptr)) return false;
      O3D_TRY(d, cudaStreamSynchronize(st));   // d.geom is reused by the second upload
      if (!upload_panels(d, st, q.tpanels, tnn, tnx, tny, tnz, ntp, tidx, nullptr, nullptr, nullptr, tarea, nullptr)) return false;
      O3D_TRY(d, cudaStreamSynchronize(st));
      return true;
    };
    if (!go()) break;
  }
  const int rc = collect(c);
  if (rc != O3D_OK) {
    o3d_cuda_bem_op_destroy(c, op);
    return rc;
  }
  *out = op;
  return O3D_OK;
}

void o3d_cuda_bem_op_destroy(o3d_ctx* c, o3d_bem_op* op) {
  if (!op) return;
  for (size_t k = 0; k < op->dev.size(); ++k) {
    if (c && k < c->dev.size()) cudaSetDevice(c->dev[k].id);
    BemDev& q = op->dev[k];
    for (DevBuf* b : {&q.bases, &q.spanels, &q.tpanels, &q.x, &q.y, &q.work, &q.cnt}) b->release();
  }
  delete op;
}

int o3d_cuda_bem_op_apply(o3d_ctx* c, o3d_bem_op* op, const float* x, float* y, double* flops_out) {
  if (!c || !op || !x || !y || op->dev.size() != c->dev.size()) return fail(c, O3D_ERR_INVALID, "bem_op_apply: bad argument");
  for (Device& d : c->dev) d.kernel_ms = d.h2d_ms = d.d2h_ms = 0, d.launches = 0, d.counts[0] = d.counts[1] = 0;
  const int64_t nsp = op->nsp, ntp = op->ntp;
  const int ndev = (int)c->dev.size();
  // enqueue on every device, then wait: the devices work concurrently
  for (int k = 0; k < ndev; ++k) {
    Device& d = c->dev[k];
    BemDev& q = op->dev[k];
    if (q.ni == 0) continue;
    auto go = [&]() {
      O3D_TRY(d, cudaSetDevice(d.id));
      cudaStream_t st = d.stream;
      O3D_TRY(d, cudaEventRecord(d.ev[0], st));
      if (!h2d(d, st, q.x.p, x, (size_t)3 * nsp * 4)) return false;
      O3D_TRY(d, cudaMemsetAsync(q.cnt.p, 0, 2 * sizeof(unsigned long long), st));
      O3D_TRY(d, cudaEventRecord(d.ev[1], st));
      float* db = q.bases.as<float>();
      PanMatvecArgs a{};
      a.c.spn = q.spanels.as<float4>();
      a.c.tpn = q.tpanels.as<float4>();
      a.c.nsp = nsp; a.c.ntp = ntp;
      a.c.sb1 = db; a.c.sb2 = db + 3 * nsp;
      a.c.tb1 = db + 6 * nsp; a.c.tb2 = db + 6 * nsp + 3 * ntp; a.c.tnrm = db + 6 * nsp + 6 * ntp;
      a.c.self = op->self;
      a.c.counts = q.cnt.as<unsigned long long>();
      a.x = q.x.as<float>();
      a.i0 = q.i0; a.ni = q.ni;
      constexpr int B = 64;
      const int64_t gx = (q.ni + B - 1) / B;
      a.nsplit = pan_nsplit(d, gx, nsp);
      O3D_TRY(d, q.work.ensure((size_t)a.nsplit * 3 * q.ni * sizeof(double)));
      a.partial = q.work.as<double>();
      pan_matvec_kernel<B><<<dim3((unsigned)gx, (unsigned)a.nsplit), B, 0, st>>>(a);
      O3D_TRY(d, cudaGetLastError());
      pan_matvec_finish_kernel<<<(unsigned)((q.ni + 255) / 256), 256, 0, st>>>(a.nsplit, q.ni, a.partial, q.y.as<float>());
      O3D_TRY(d, cudaGetLastError());
      d.launches += 2;
      O3D_TRY(d, cudaEventRecord(d.ev[2], st));
      if (!d2h(d, st, y + 3 * q.i0, q.y.p, (size_t)3 * q.ni * 4)) return false;
      if (!d2h(d, st, d.counts, q.cnt.p, 2 * sizeof(unsigned long long))) return false;
      O3D_TRY(d, cudaEventRecord(d.ev[3], st));
      return true;
    };
    if (!go()) break;
  }
  for (int k = 0; k < ndev; ++k) {
    Device& d = c->dev[k];
    if (op->dev[k].ni == 0 || d.status != cudaSuccess) continue;
    cudaSetDevice(d.id);
    finish_timing(d);
  }
  const int rc = collect(c);
  if (rc == O3D_OK && flops_out) {
    // the traversal costs what panels_on_panels_coeff costs (src/Kernels.h:1250-1292 counts, three unit strengths),
    // plus 18 flops per block for the product
    double leaves = 0, splits = 0;
    for (Device& d : c->dev) leaves += (double)d.counts[0], splits += (double)d.counts[1];
    *flops_out = 3.0 * ((leaves + splits) * 31.0 + leaves * 37.0 + splits * 42.0) + 27.0 * (double)nsp * (double)ntp;
  }
  return rc;
}

// ---------------------------------------------------------------------------------------------------
// Particle x panel closest-point loops: reflect_panp2 and clear_inner_panp2 (method 1)
namespace {
int closest_point_pass(o3d_ctx* c, const char* who, int mode, float cutoff, int64_t nn, const float* nx, const float* ny,
                       const float* nz, int64_t np, const uint32_t* idx, const float* nrm, int64_t nt, float* tx, float* ty,
                       float* tz, int64_t* num_moved) {
  if (!check_counts(c, np, nt) || nn < 0 || nn >= (int64_t(1) << 31)) return fail(c, O3D_ERR_INVALID, std::string(who) + ": bad context or counts");
  if (np > 0 && (!nx || !ny || !nz || !idx || !nrm)) return fail(c, O3D_ERR_INVALID, std::string(who) + ": NULL panel array");
  if (nt > 0 && (!tx || !ty || !tz)) return fail(c, O3D_ERR_INVALID, std::string(who) + ": NULL particle array");
  if (np > 0 && !valid_indices(idx, np, nn)) return fail(c, O3D_ERR_INVALID, std::string(who) + ": node index out of range");
  for (Device& d : c->dev) d.kernel_ms = d.h2d_ms = d.d2h_ms = 0, d.launches = 0, d.counts[0] = d.counts[1] = 0;
  if (num_moved) *num_moved = 0;
  if (np == 0 || nt == 0) return collect(c);
  const int ndev = (int)c->dev.size();
  const int64_t npad = ((np + kRefTile - 1) / kRefTile) * kRefTile;
  for_each_device(c, [&](int k) {
    Device& d = c->dev[k];
    int64_t t0, t1;
    partition(nt, ndev, k, &t0, &t1);
    const int64_t n = t1 - t0;
    if (n == 0) return true;
    O3D_TRY(d, cudaSetDevice(d.id));
    cudaStream_t st = d.stream;
    O3D_TRY(d, d.geom.ensure(((size_t)3 * nn + (size_t)6 * np) * 4));
    O3D_TRY(d, d.panels.ensure((size_t)npad * kRefRec * sizeof(float4)));
    O3D_TRY(d, d.targ.ensure((size_t)3 * n * 4));
    float* g = d.geom.as<float>();
    uint32_t* gidx = reinterpret_cast<uint32_t*>(g + 3 * nn);
    float* gnrm = g + 3 * nn + 3 * np;
    float* dt = d.targ.as<float>();
    O3D_TRY(d, cudaEventRecord(d.ev[0], st));
    if (!h2d(d, st, g, nx, (size_t)nn * 4)) return false;
    if (!h2d(d, st, g + nn, ny, (size_t)nn * 4)) return false;
    if (!h2d(d, st, g + 2 * nn, nz, (size_t)nn * 4)) return false;
    if (!h2d(d, st, gidx, idx, (size_t)3 * np * 4)) return false;
    if (!h2d(d, st, gnrm, nrm, (size_t)3 * np * 4)) return false;
    float* hx[3] = {tx, ty, tz};
    for (int a = 0; a < 3; ++a) if (!h2d(d, st, dt + (size_t)a * n, hx[a] + t0, (size_t)n * 4)) return false;
    if (!zero_counts(d, st)) return false;
    O3D_TRY(d, cudaEventRecord(d.ev[1], st));
    ref_pack_kernel<<<(unsigned)((npad + 127) / 128), 128, 0, st>>>(np, npad, g, g + nn, g + 2 * nn, gidx, gnrm, d.panels.as<float4>());
    O3D_TRY(d, cudaGetLastError());
    ReflectArgs a{};
    a.pan = d.panels.as<float4>();
    a.np = np;
    a.ntiles = (int)(npad / kRefTile);
    a.nt = n;
    a.tx = dt; a.ty = dt + n; a.tz = dt + 2 * n;
    a.mode = mode;
    a.cutoff = cutoff;
    a.count = d.cnt.as<unsigned long long>();
    constexpr int B = 128;
    reflect_kernel<B><<<(unsigned)((n + B - 1) / B), B, 0, st>>>(a);
    O3D_TRY(d, cudaGetLastError());
    d.launches += 2;
    O3D_TRY(d, cudaEventRecord(d.ev[2], st));
    for (int a2 = 0; a2 < 3; ++a2) if (!d2h(d, st, hx[a2] + t0, dt + (size_t)a2 * n, (size_t)n * 4)) return false;
    if (!fetch_counts(d, st)) return false;
    O3D_TRY(d, cudaEventRecord(d.ev[3], st));
    return finish_timing(d);
  });
  const int rc = collect(c);
  if (rc == O3D_OK && num_moved) {
    unsigned long long m = 0;
    for (Device& d : c->dev) m += d.counts[0];
    *num_moved = (int64_t)m;
  }
  return rc;
}
}  // namespace

int o3d_cuda_reflect_pts(o3d_ctx* c, int64_t nn, const float* nx, const float* ny, const float* nz, int64_t np,
                         const uint32_t* idx, const float* nrm, int64_t nt, float* tx, float* ty, float* tz, int64_t* num_reflected) {
  return closest_point_pass(c, "reflect_pts", 0, 0.0f, nn, nx, ny, nz, np, idx, nrm, nt, tx, ty, tz, num_reflected);
}

int o3d_cuda_clear_inner_pts(o3d_ctx* c, int method, int64_t nn, const float* nx, const float* ny, const float* nz, int64_t np,
                             const uint32_t* idx, const float* nrm, int64_t nt, float* tx, float* ty, float* tz, float cutoff_mult,
                             float ips, int64_t* num_moved) {
  // method 0 (table-driven cropping of strength, src/Reflect.h:547-567) has no caller in the reference: every call site
  // passes 1 (src/Convection.h:260-556, src/Diffusion.h:306, src/Simulation.cpp:839)
  if (method != 1) return fail(c, O3D_ERR_UNSUPPORTED, "clear_inner_pts: only method 1 (push out, keep strength) is implemented");
  const float cutoff = cutoff_mult * ips;   // float product, as "_cutoff_mult*_ips" with S = float (src/Reflect.h:537)
  return closest_point_pass(c, "clear_inner_pts", 1, cutoff, nn, nx, ny, nz, np, idx, nrm, nt, tx, ty, tz, num_moved);
}

// ---------------------------------------------------------------------------------------------------
// Field output: the reference's Points::write_vtk file, byte for byte (csrc/vtu_writer.h)
int o3d_cuda_write_points_vtu(const char* path, int64_t n, const float* x, const float* y, const float* z, const float* sx,
                              const float* sy, const float* sz, const float* r, const float* u, const float* v, const float* w,
                              double time) {
  if (!path || n <= 0 || n >= (int64_t(1) << 31) || !x || !y || !z || !u || !v || !w) return O3D_ERR_INVALID;
  if ((sx || sy || sz) && !(sx && sy && sz)) return O3D_ERR_INVALID;
  const float* xs[3] = {x, y, z};
  const float* ss[3] = {sx, sy, sz};
  const float* us[3] = {u, v, w};
  return write_points_vtu(path, n, xs, sx ? ss : nullptr, r, us, time) ? O3D_OK : O3D_ERR_INVALID;
}

int o3d_cuda_particles_write_vtu(o3d_ctx* c, o3d_particles* p, const char* path, double time) {
  if (!c || !p || !path || p->n <= 0) return fail(c, O3D_ERR_INVALID, "particles_write_vtu: bad argument or empty collection");
  const size_t n = (size_t)p->n;
  std::vector<float> h(10 * n);
  float* a = h.data();
  const int rc = o3d_cuda_particles_download(c, p, a, a + n, a + 2 * n, a + 3 * n, a + 4 * n, a + 5 * n, a + 6 * n, nullptr, a + 7 * n,
                                             a + 8 * n, a + 9 * n, nullptr);
  if (rc != O3D_OK) return rc;
  if (o3d_cuda_write_points_vtu(path, p->n, a, a + n, a + 2 * n, a + 3 * n, a + 4 * n, a + 5 * n, a + 6 * n, a + 7 * n, a + 8 * n, a + 9 * n,
                                time) != O3D_OK)
    return fail(c, O3D_ERR_INVALID, std::string("particles_write_vtu: cannot write ") + path);
  return O3D_OK;
}

int o3d_cuda_set_graphs(o3d_ctx* c, int on) {
  if (!c) return O3D_ERR_INVALID;
  c->use_graphs = on != 0;
  return O3D_OK;
}

int o3d_cuda_set_panel_queue(o3d_ctx* c, int on) {
  if (!c) return O3D_ERR_INVALID;
  for (Device& d : c->dev) d.pan_queue = on < 0 ? 0 : on > 2 ? 2 : on;
  return O3D_OK;
}

int o3d_cuda_set_host_staging(o3d_ctx* c, int on) {
  if (!c) return O3D_ERR_INVALID;
  for (Device& d : c->dev) d.stage.enabled = on != 0;
  return O3D_OK;
}

int o3d_cuda_particles_graph_active(const o3d_particles* p) { return p && !p->dev.empty() && p->dev[0].graph != nullptr; }

int o3d_cuda_set_tuned_kernels(o3d_ctx* c, int on) {
  if (!c) return O3D_ERR_INVALID;
  for (Device& d : c->dev) d.tuned = on != 0;
  return O3D_OK;
}

int o3d_cuda_tuned_kernels(const o3d_ctx* c) { return c && !c->dev.empty() && c->dev[0].tuned; }

int o3d_cuda_set_core_func(o3d_ctx* c, int core) {
  if (!c) return O3D_ERR_INVALID;
  if (core < O3D_CORE_WL || core > O3D_CORE_V2) return fail(c, O3D_ERR_INVALID, "set_core_func: unknown core function");
  for (Device& d : c->dev) d.core = core;
  return O3D_OK;
}

int o3d_cuda_core_func(const o3d_ctx* c) { return c && !c->dev.empty() ? c->dev[0].core : -1; }

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

// biot_pp.cuh - particles -> points direct Biot-Savart sums for sm_100a (B200).
//
// Replaces the loop nests of points_affect_points<S,A> (reference src/Influence.h:278-309,
// :351-365, :443-474, :518-533) and the pairwise kernels they call (src/Kernels.h:50-69 kernel_0v_0b,
// :94-112 kernel_0v_0p, :155-193 kernel_0v_0bg, :253-290 kernel_0v_0pg) with the Winckelmans-Leonard
// core (src/CoreFunc.h:245-288). Written from the formulas, not from those loops:
//
//   d = t - s,  r2 = sr^2 + tr^2,  d2 = |d|^2 + r2,  top = |d|^2 + 2.5 r2 = d2 + 1.5 r2
//   dn5 = d2^(-5/2) = rs^5 with rs = MUFU.RSQ(d2)          (reference: 1/(d2*d2*sqrt(d2)))
//   r3  = top * dn5,   bbb = dn5 * (2 - 5 top / d2) = dn5 * (2 - 5 top rs^2)
//   c   = (dz wy - dy wz, dx wz - dz wx, dy wx - dx wy)
//   u  += r3 c ;   G[3j+i] += d_j (bbb c_i)  (+/- w_k r3 on the six off-diagonals)
//
// Design:
//   * each thread owns T targets in registers; every source record is read once from shared memory
//     (one broadcast LDS.128 per warp) and applied to all T targets, two sources per packed FP32
//     instruction: 35 FFMA2/FMUL2/FADD2 + 1 MUFU.RSQ per interaction for velocity+gradient;
//   * the antisymmetric +/- w_k r3 part of the gradient is accumulated once as A = sum w r3 (3 FMA)
//     and applied in the epilogue instead of 6 FMA per interaction;
//   * source tiles (512 records = 16 KB, contiguous) arrive by cp.async.bulk (TMA, SASS UBLKCP)
//     into a 2-deep mbarrier ring - no thread spends registers or issue slots on the copy;
//   * ONE PERSISTENT 384-thread CTA per SM (the 12 warps x 168 registers the register file holds; co-resident CTAs are served
//     in age order) over a static partition of the nblocks x ntiles grid of (target block, source tile) units
//     (PPPlan below): every CTA first finishes nblocks / P WHOLE target blocks in place - all CTAs walk the source stream
//     together, so a tile fetched from HBM by one is an L2 hit for the others - then takes an equal share of the units of the
//     remaining < P blocks (stream-K). Every CTA streams the same number of tiles (+-1) whatever the target count - no
//     last-wave quantisation, no source split - and the TMA ring runs straight through block and phase boundaries. The (at most
//     P-1) tail blocks cut by a share boundary leave FP64 partial sums in a fixed-size workspace (2 slots per CTA) which
//     pp_fixup_kernel adds in unit order: deterministic, no atomics, a few MB of traffic whatever the problem size;
//   * in the velocity+gradient kernel no CTA-wide barrier per tile: every warp counts itself out of a ring buffer and the last
//     one out issues the refill (O3D_PP_NOBAR);
//   * sums are FP32 FMA chains inside one tile, promoted to FP64 once per tile (B200 keeps a full FP64
//     pipe; 12 DADD per 512 interactions), mirroring the reference's float-kernel/double-accumulator
//     scheme (src/Simulation.h:41-47) to ~1e-7 relative;
//   * outputs are read-modify-write, un-normalised: tu[i] = float(double(tu[i]) + sum), exactly the
//     reference's "tu[0][i] += accumu" (src/Influence.h:462-473).
#pragma once
#include "o3d_common.cuh"

#ifndef O3D_PP_TRACE
#define O3D_PP_TRACE 1    // 1: recover the wz gradient slot from the trace-free identity instead of accumulating it
#endif
#ifndef O3D_PP_POW
#define O3D_PP_POW 2      // 1: uniform-radius path forms r3 and bbb from odd powers of rs (two packed instructions fewer);
#endif                    // 2: ... with the two constants as 32-bit broadcast operands and the searched statement order
#ifndef O3D_PP_BODY_FILE
#define O3D_PP_BODY_FILE "pp_body_velgrad_uni.inc"       // velocity + gradient, uniform radii (the 1M-16M benchmark path)
#endif
#ifndef O3D_PP_BODY_FILE_GEN
#define O3D_PP_BODY_FILE_GEN "pp_body_velgrad_gen.inc"   // velocity + gradient, per-particle radii
#endif
#ifndef O3D_PP_BODY_FILE_VEL
#define O3D_PP_BODY_FILE_VEL "pp_body_vel_uni.inc"       // velocity only, uniform radii
#endif
#ifndef O3D_PP_BODY_FILE_VELGEN
#define O3D_PP_BODY_FILE_VELGEN "pp_body_vel_gen.inc"    // velocity only, per-particle radii
#endif
#ifndef O3D_PP_JOINT
#define O3D_PP_JOINT 0    // 1 (tools/tune_order.py, O3D_TUNE_JOINT=1 only): the velocity+gradient, uniform-radius loop body covers
#endif                    //    BOTH register-blocked targets in one searched statement order, written by the tool to O3D_PP_JOINT_FILE
#ifndef O3D_PP_JOINT_FILE
#define O3D_PP_JOINT_FILE "pp_body_velgrad_uni_joint.inc"
#endif
#ifndef O3D_PP_BLOCK
#define O3D_PP_BLOCK 384  // threads per CTA. The kernels need 12 warps per SM (168 registers each: the whole register file) and run them as
#endif                    // ONE persistent CTA per SM. 128 (three CTAs per SM; microbench/kbench only) loses 2-3 %: the warp schedulers serve
                          // co-resident CTAs in strict age order - the oldest CTA of an SM finishes its static share at 39 % of the
                          // kernel's duration, the second at 70 %, and the youngest then runs alone, one warp per scheduler
                          // (profiles/r02_cta_end_times.txt). Inside one CTA the per-tile barrier keeps the twelve warps together.
#ifndef O3D_PP_UNROLL_GRAD
#define O3D_PP_UNROLL_GRAD 4   // source pairs per trip of the packed inner loop, velocity+gradient kernel
#endif
#ifndef O3D_PP_STAGE
#define O3D_PP_STAGE 0    // how a source tile reaches shared memory. 0 (product): one cp.async.bulk (TMA) per tile by one elected
#endif                    // thread + mbarrier; 1: cp.async 16 B x 8 per thread; 2: LDG.128 -> STS.128 through registers. 1 and 2 exist
                          // for microbench/kbench only (make kbench_stage): DESIGN.md section 7 "staging" rows
#ifndef O3D_PP_TWOBUF
#define O3D_PP_TWOBUF 1   // 1 (product): the walk alternates two copies of the tile body, one per ring buffer, each with static
#endif                    // shared-memory addresses; 0: one copy, the buffer's shared address handed to ptxas through a warp
                          // reduction once per tile (microbench/kbench only: DESIGN.md section 7)
#ifndef O3D_PP_NOBAR
#define O3D_PP_NOBAR 1    // no CTA-wide barrier per tile: every warp counts itself out of a ring buffer (shared-memory counter) and the last
#endif                    // one out refills it, so the warps of a CTA may drift one tile apart. 0: barrier in every kernel; 1 (product): the
                          // velocity+gradient pp2_kernel goes without (-0.8 %; the velocity-only kernel measures 1.1 % SLOWER without its
                          // barrier: profiles/r02_kbench_nobar.txt); 2: every kernel without (microbench/kbench only)
#ifndef O3D_PP_PURE
#define O3D_PP_PURE 0     // 1 (microbench/kbench only): no whole-block phase - the pure stream-K partition of every unit
#endif
#ifndef O3D_PP_SKEW
#define O3D_PP_SKEW 0     // > 0 (microbench/kbench only): CTA c starts c * O3D_PP_SKEW nanoseconds late
#endif
#ifndef O3D_PP_STAGGER
#define O3D_PP_STAGGER 0  // > 0 (microbench/kbench only): the k-th CTA to arrive on an SM starts (k mod 3) * O3D_PP_STAGGER clocks late, so
#endif                    // that the co-resident CTAs do not reach their tile boundaries together
#ifndef O3D_PP_UNROLL_VEL
#define O3D_PP_UNROLL_VEL 2    // ... velocity-only kernel
#endif

namespace o3d {

#if O3D_PP_STAGGER
__device__ unsigned pp_stagger_arrivals[1024];
#endif
#ifdef O3D_PP_ENDTIME
__device__ unsigned long long pp_end_time[2048];   // microbench only: globaltimer at the end of every CTA
__device__ unsigned pp_end_smid[2048];
__device__ long long pp_tile_clock[12][64];        // ... and clock64 of every warp of CTA 0 when it has finished the arithmetic of its k-th tile
#endif

// Product launch configuration of pp2_kernel (capi.cu launches these two instantiations; pp_tuned.cu compiles the same
// two into the cubin that tools/sass_patch.py post-processes): one 384-thread CTA per SM, 2 register-blocked targets per
// thread with gradients, 4 without.
constexpr int kPPBlock = O3D_PP_BLOCK;
constexpr int kPPWarpsPerSM = 12;                       // 65536 registers / (168 x 32)
constexpr int kPPResident = kPPWarpsPerSM * 32 / kPPBlock;   // persistent CTAs per SM
static_assert(kPPResident * kPPBlock == kPPWarpsPerSM * 32, "CTA size must divide 384 threads");
// Systems of at most ONE product-size target block (C1 as shipped: 210 particles) run the same kernels as 128-thread CTAs, three per
// SM: a 384-thread CTA would carry the whole system on one SM with most of its warps summing for no target (0.19 ms instead of
// 0.12 ms per RK2 step at 210 particles). Which CTA of an SM runs first does not matter at that size.
constexpr int kPPSmallBlock = 128;
constexpr int kPPTgrad = 2;
constexpr int kPPTvel = 4;

constexpr int kPPUnrollGrad = O3D_PP_UNROLL_GRAD;
constexpr int kPPUnrollVel = O3D_PP_UNROLL_VEL;

__device__ __forceinline__ float rsqrt_approx(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <bool GRAD> struct PPAcc { static constexpr int N = GRAD ? 15 : 3; };

// Fold the per-tile FP32 partials into the 3 (or 12) FP64 running sums and clear them.
template <bool GRAD>
__device__ __forceinline__ void pp_promote(float (&acc)[PPAcc<GRAD>::N], double (&sum)[GRAD ? 12 : 3]) {
  sum[0] += (double)acc[0]; sum[1] += (double)acc[1]; sum[2] += (double)acc[2];
  if constexpr (GRAD) {
    const float ax = acc[12], ay = acc[13], az = acc[14];
    sum[3]  += (double)acc[3];            // ux
    sum[4]  += (double)(acc[4] + az);     // vx = sum dx*cy + wz*r3
    sum[5]  += (double)(acc[5] - ay);     // wx = sum dx*cz - wy*r3
    sum[6]  += (double)(acc[6] - az);     // uy
    sum[7]  += (double)acc[7];            // vy
    sum[8]  += (double)(acc[8] + ax);     // wy
    sum[9]  += (double)(acc[9] + ay);     // uz
    sum[10] += (double)(acc[10] - ax);    // vz
    sum[11] += (double)acc[11];           // wz
  }
#pragma unroll
  for (int k = 0; k < PPAcc<GRAD>::N; ++k) acc[k] = 0.0f;
}


// ---- the static partition: whole blocks first, a stream-K tail -------------------------------------------------------
// Units are (target block b, source tile k). With P persistent CTAs and nblocks target blocks:
//   phase A  CTA c owns the `full` = nblocks / P WHOLE blocks [c full, (c+1) full): every CTA starts every block at tile 0 at
//            (nearly) the same time, so the P CTAs walk the source stream together and a tile fetched from HBM by one of them
//            is an L2 hit for the others - one pass over the sources per sweep of P blocks, whatever their size (with a pure
//            stream-K partition the CTAs sit at P different places of the stream, and a source set larger than L2 is re-read
//            from HBM by each of them: 368 GB per launch measured at 4 M particles, profiles/r02_pp2_dram_traffic.txt);
//   phase B  the remaining nblocks - P full < P blocks are dealt out stream-K: their Wt = (nblocks - P full) ntiles units in
//            block-major order, CTA c < Pt owns [Wt c / Pt, Wt (c+1) / Pt), Pt = min(P, Wt). Every CTA streams the same number
//            of tiles (+-1) whatever the target count - no last-wave quantisation - and the TMA ring runs straight through
//            block and phase boundaries.
// Host and device use the same arithmetic (capi.cu: launch shape, o3d_cuda_plan_pts_on_pts, o3d_cuda_plan_check; the kernels;
// pp_fixup_kernel).
struct PPPlan {
  int nblocks, ntiles;
  int P;          // CTAs
  int full;       // whole blocks per CTA (phase A)
  int Pt;         // CTAs that take part in the tail (phase B); 0: no tail
  int64_t Wt;     // tail units
  __host__ __device__ int tail_block0() const { return P * full; }
  __host__ __device__ int64_t begin(int c) const { return Pt ? Wt * (int64_t)(c < Pt ? c : Pt) / Pt : 0; }   // first tail unit of CTA c
  __host__ __device__ int64_t tiles_of(int c) const { return (int64_t)full * ntiles + begin(c + 1) - begin(c); }
};
// slots = resident CTA slots of the device (SMs x kPPResident)
__host__ __device__ inline PPPlan pp_make_plan(int slots, int nblocks, int ntiles) {
  PPPlan q;
  q.nblocks = nblocks; q.ntiles = ntiles;
  const int64_t W = (int64_t)nblocks * ntiles;
  q.P = (int)(W < slots ? W : slots);
  q.full = O3D_PP_PURE ? 0 : nblocks / q.P;
  q.Wt = (int64_t)(nblocks - q.P * q.full) * ntiles;
  q.Pt = (int)(q.Wt < q.P ? q.Wt : q.P);
  return q;
}
// A CTA's tail range is cut into segments at target-block boundaries. Only its FIRST and its LAST tail segment can cover a
// block partially; a partial first segment leaves its sums in slot 0 of the CTA's workspace pair, a partial last segment that
// is not also the first in slot 1.
constexpr int kPPSlots = 2;

struct PPArgs {
  const float4* src;      // packed source stream, 2 float4 per source, padded to whole tiles
  int ntiles;             // tiles in the whole stream
  int nblocks;            // target blocks of BLOCK * T targets
  int slots;              // resident CTA slots the launch was planned for (gridDim.x = pp_make_plan(slots, ..).P)
  int64_t nt;             // targets
  const float* tx; const float* ty; const float* tz;
  const float* tr;        // nullptr => singular targets (tr = 0)
  float* tu; float* tv; float* tw;   // velocity, read-modify-write
  float* tug;             // 9 rows of stride tug_stride, or nullptr
  int64_t tug_stride;
  double* partial;        // [gridDim.x][kPPSlots][12 or 3][BLOCK * T] FP64: sums of the target blocks this CTA shares
  double* acc64;          // nullptr: results are added into tu/tv/tw/tug (read-modify-write). Otherwise the FP64 sums are STORED
                          // here, [12 or 3][acc_stride], and the outputs are not touched: the host entry points add them to
                          // the caller's initial values afterwards (pp_accumulate_kernel), so that the upload of those values
                          // overlaps the kernel instead of preceding it
  int64_t acc_stride;
  float sign;             // +1
  const uint32_t* radius_range;  // pp_scan_kernel output, or nullptr: no uniform-radius fast path
};

// The walk of one persistent CTA through its share. Everything here is a function of blockIdx and the kernel parameters
// only, so ptxas keeps it in UNIFORM registers: the inner loops address shared memory as [UR + imm] and branch on uniform
// predicates exactly as a one-block-per-CTA kernel would (values re-read from shared memory would count as divergent and move
// the loop counters and LDS addresses into the vector register file, whose read bandwidth is what bounds the hot loop).
// mbarriers of the two ring buffers, and (O3D_PP_NOBAR) how many of the CTA's warps have left each
struct PPSync { uint64_t full[2]; unsigned released[2]; };

struct PPWalk {
  int nk;           // tiles in the CTA's share: nA + its tail units
  int nA;           // ... of which phase A (whole blocks)
  int bA0;          // first whole block
  int bB0, ktB0;    // target block and source tile of its first tail unit
  int kring;        // tiles consumed so far; tile kring lives in ring buffer kring & 1
  int tnext;        // source tile the next refill fetches (wraps from the last tile of a block to tile 0 of the next)
};
__device__ __forceinline__ int pp_wrap(int t, int ntiles) { return t + 1 == ntiles ? 0 : t + 1; }

#if O3D_PP_STAGE != 0
// microbench-only staging variants: every thread moves its 128-byte share of the tile (8 x 16 B, coalesced per 16-byte column)
template <int BLOCK>
__device__ __forceinline__ void pp_stage_copy(float4* dst, const float4* __restrict__ src) {
  constexpr int NQ = (kTile * 2 + BLOCK - 1) / BLOCK;
#if O3D_PP_STAGE == 1
#pragma unroll
  for (int q = 0; q < NQ; ++q)
    if (q * BLOCK + (int)threadIdx.x < kTile * 2)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + q * BLOCK + threadIdx.x)), "l"(src + q * BLOCK + threadIdx.x) : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
#else
  float4 v[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) if (q * BLOCK + (int)threadIdx.x < kTile * 2) v[q] = __ldg(src + q * BLOCK + threadIdx.x);
#pragma unroll
  for (int q = 0; q < NQ; ++q) if (q * BLOCK + (int)threadIdx.x < kTile * 2) dst[q * BLOCK + threadIdx.x] = v[q];
#endif
}
#endif

// fetch source tile w.tnext into ring buffer `buf` (barrier `bar`) and advance
template <int BLOCK>
__device__ __forceinline__ void pp_ring_fetch(const PPArgs& p, PPWalk& w, float4* buf, uint64_t* bar) {
#if O3D_PP_STAGE == 0
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, kTileBytes);
    bulk_g2s(buf, p.src + (size_t)w.tnext * (kTile * 2), kTileBytes, bar);
  }
#else
  pp_stage_copy<BLOCK>(buf, p.src + (size_t)w.tnext * (kTile * 2));
#endif
}
// the source tile of ring position `pos` has been requested: which one does position pos + 1 hold? (the walk jumps from the last
// tile of phase A to the CTA's first tail unit)
__device__ __forceinline__ void pp_ring_advance(const PPArgs& p, PPWalk& w, const int pos) {
  w.tnext = pos + 1 == w.nA ? w.ktB0 : pp_wrap(w.tnext, p.ntiles);
}

template <int BLOCK>
__device__ __forceinline__ PPWalk pp_ring_start(const PPArgs& p, float4 (&tile)[2][kTile * 2], PPSync& sy) {
  const PPPlan plan = pp_make_plan(p.slots, p.nblocks, p.ntiles);
  const int64_t u0 = plan.begin(blockIdx.x), u1 = plan.begin(blockIdx.x + 1);
  // (the 64-bit divisions run as a subroutine on the vector datapath, after which ptxas no longer knows the quotients to be
  // warp-uniform; a warp reduction - REDUX writes a uniform register - hands them back as provably uniform values)
  const int bt = (int)(u0 / p.ntiles);
  PPWalk w;
  w.nA = __reduce_max_sync(0xffffffffu, plan.full * p.ntiles);
  w.nk = __reduce_max_sync(0xffffffffu, plan.full * p.ntiles + (int)(u1 - u0));
  w.bA0 = __reduce_max_sync(0xffffffffu, (int)blockIdx.x * plan.full);
  w.bB0 = __reduce_max_sync(0xffffffffu, plan.tail_block0() + bt);
  w.ktB0 = __reduce_max_sync(0xffffffffu, (int)(u0 - (int64_t)bt * p.ntiles));
  w.kring = 0;
  w.tnext = w.nA > 0 ? 0 : w.ktB0;
#if O3D_PP_STAGE == 0
  if (threadIdx.x == 0) {
    mbar_init(&sy.full[0], 1);
    mbar_init(&sy.full[1], 1);
    mbar_fence_init();
    sy.released[0] = sy.released[1] = 0u;
  }
  __syncthreads();
#endif
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    if (s < w.nk) { pp_ring_fetch<BLOCK>(p, w, tile[s], &sy.full[s]); pp_ring_advance(p, w, s); }
#if O3D_PP_STAGE == 1
    else asm volatile("cp.async.commit_group;" ::: "memory");   // keep the group count in step with the tile count
#endif
  }
  return w;
}
// tile w.kring of the CTA's share has landed in its ring buffer
__device__ __forceinline__ void pp_ring_wait(PPSync& sy, const int kring) {
#if O3D_PP_STAGE == 0
  mbar_wait(&sy.full[kring & 1], (kring >> 1) & 1);
#else
#if O3D_PP_STAGE == 1
  asm volatile("cp.async.wait_group 1;" ::: "memory");   // all but the newest group
#endif
  __syncthreads();
#endif
}
// every warp is done with the buffer of tile w.kring: refill it with the tile two ahead, step to the next tile
template <int BLOCK, bool NOBAR>
__device__ __forceinline__ void pp_ring_refill(const PPArgs& p, PPWalk& w, float4* buf, uint64_t* bar, unsigned* released) {
#if O3D_PP_STAGE == 0
  if constexpr (NOBAR) {
  // each warp counts itself out of the buffer; the last of the CTA's warps to leave it issues the refill
  __syncwarp();
  const bool more = w.kring + 2 < w.nk;
  if ((threadIdx.x & 31) == 0) {
    if (atomicAdd(released, 1u) == BLOCK / 32 - 1) {
      *released = 0u;
      if (more) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, kTileBytes);
        bulk_g2s(buf, p.src + (size_t)w.tnext * (kTile * 2), kTileBytes, bar);
      }
    }
  }
  if (more) pp_ring_advance(p, w, w.kring + 2);
  ++w.kring;
  return;
  }
#endif
  (void)released;
  __syncthreads();
  if (w.kring + 2 < w.nk) { pp_ring_fetch<BLOCK>(p, w, buf, bar); pp_ring_advance(p, w, w.kring + 2); }
#if O3D_PP_STAGE == 1
  else asm volatile("cp.async.commit_group;" ::: "memory");
#endif
  ++w.kring;
}
// Epilogue of one segment for one thread's T targets: a whole block is finished in place, a partial one parks its sums.
template <int T, bool GRAD, int BLOCK>
__device__ __forceinline__ void pp_store(const PPArgs& p, const int b, const bool whole, const int slot,
                                         const double (&sum)[T][GRAD ? 12 : 3]) {
  constexpr int NS = GRAD ? 12 : 3;
  constexpr int PER = BLOCK * T;
  if (!whole) {
    double* w = p.partial + ((size_t)blockIdx.x * kPPSlots + slot) * (NS * PER) + threadIdx.x;
#pragma unroll
    for (int t = 0; t < T; ++t) {
#pragma unroll
      for (int k = 0; k < NS; ++k) w[k * PER + t * BLOCK] = sum[t][k];
    }
    return;
  }
  const int64_t base = (int64_t)b * PER + threadIdx.x;
  if (p.acc64) {
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int64_t i = base + (int64_t)t * BLOCK;
      if (i >= p.nt) continue;
#pragma unroll
      for (int k = 0; k < NS; ++k) p.acc64[(size_t)k * p.acc_stride + i] = sum[t][k];
    }
    return;
  }
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const int64_t i = base + (int64_t)t * BLOCK;
    if (i >= p.nt) continue;
    const double sg = (double)p.sign;
    p.tu[i] = (float)((double)p.tu[i] + sg * sum[t][0]);
    p.tv[i] = (float)((double)p.tv[i] + sg * sum[t][1]);
    p.tw[i] = (float)((double)p.tw[i] + sg * sum[t][2]);
    if constexpr (GRAD) {
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        float* g = p.tug + (size_t)k * p.tug_stride + i;
        *g = (float)((double)*g + sg * sum[t][3 + k]);
      }
    }
  }
}

// Per-target-block state of a persistent CTA's walk (pp2_walk / ppc_kernel)
struct PPBlock {
  int b, kt;        // target block and source tile of the current unit
  bool seg_first;   // the current segment is the first of the CTA's tail (the only one of its FIRST segments that may be partial)
  bool fresh;       // the current tile starts a segment: (re)load the targets, clear the FP64 sums
};

// The tile just consumed may end a segment: store it, move on to the next block - or from the last whole block to the CTA's tail.
template <int T, bool GRAD, int BLOCK>
__device__ __forceinline__ void pp_segment_end(const PPArgs& p, const PPWalk& w, PPBlock& s, const double (&sum)[T][GRAD ? 12 : 3]) {
  ++s.kt;
  s.fresh = s.kt == p.ntiles || w.kring == w.nk;
  if (s.fresh) {
    // whole block <=> the segment ran from tile 0 to the last tile: kt == ntiles and (not a first tail segment, or one that began at tile 0)
    const bool whole = s.kt == p.ntiles && (!s.seg_first || w.ktB0 == 0);
    pp_store<T, GRAD, BLOCK>(p, s.b, whole, s.seg_first ? 0 : 1, sum);
    s.seg_first = false;
    s.kt = 0;
    ++s.b;
    if (w.kring == w.nA) { s.b = w.bB0; s.kt = w.ktB0; s.seg_first = true; }   // phase A is done (nA > 0 here: kring >= 1)
  }
}
__device__ __forceinline__ PPBlock pp_first_block(const PPWalk& w) {
  return w.nA > 0 ? PPBlock{w.bA0, 0, false, true} : PPBlock{w.bB0, w.ktB0, true, true};
}

// One CTA per TAIL-range boundary j = blockIdx.x + 1 (1 .. Pt-1) of the launch that just ran: if that boundary is the first one
// inside its target block, add the block's pieces in unit order - the last segment of CTA j-1, then the first tail segments of
// CTAs j, j+1, ... that start inside the block - and finish the block: out = float(double(out) + sign * sum).
// per = targets per block (BLOCK * T of the main kernel).
__global__ void pp_fixup_kernel(const int nrows, const int per, const PPPlan plan, const int64_t nt, const double* __restrict__ partial,
                                float* tu, float* tv, float* tw, float* tug, const int64_t tug_stride, const float sign,
                                double* acc64, const int64_t acc_stride) {
  const int j = blockIdx.x + 1;
  const int64_t cut = plan.begin(j);
  const int64_t bt = cut / plan.ntiles, start = bt * plan.ntiles, end = start + plan.ntiles;   // tail block, its unit range
  if (cut == start) return;                 // the boundary coincides with a block edge: nothing is shared here
  const int64_t prev = plan.begin(j - 1);
  if (prev > start) return;                 // an earlier boundary inside this block owns it
  const int64_t b = plan.tail_block0() + bt;
  const size_t slot_elems = (size_t)nrows * per;
  for (int l = threadIdx.x; l < per; l += blockDim.x) {
    const int64_t i = b * per + l;
    if (i >= nt) return;
    for (int k = 0; k < nrows; ++k) {
      double acc = partial[((size_t)(j - 1) * kPPSlots + (prev == start ? 0 : 1)) * slot_elems + (size_t)k * per + l];
      for (int c = j; c < plan.Pt && plan.begin(c) < end; ++c)
        acc += partial[((size_t)c * kPPSlots) * slot_elems + (size_t)k * per + l];
      if (acc64) {
        acc64[(size_t)k * acc_stride + i] = acc;
        continue;
      }
      float* o = k == 0 ? tu + i : k == 1 ? tv + i : k == 2 ? tw + i : tug + (size_t)(k - 3) * tug_stride + i;
      *o = (float)((double)*o + (double)sign * acc);
    }
  }
}

// out[k][i] = float(double(out[k][i]) + sign * acc64[k][i]): the read-modify-write of the kernels' epilogue as its own pass, for
// launches that stored their FP64 sums (PPArgs::acc64). Same operation, same rounding: the bits are those of the fused path.
__global__ void pp_accumulate_kernel(const int nrows, const int64_t nt, const double* __restrict__ acc64, const int64_t acc_stride,
                                     float* tu, float* tv, float* tw, float* tug, const int64_t tug_stride, const float sign) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nt) return;
  for (int k = 0; k < nrows; ++k) {
    float* o = k == 0 ? tu + i : k == 1 ? tv + i : k == 2 ? tw + i : tug + (size_t)(k - 3) * tug_stride + i;
    *o = (float)((double)*o + (double)sign * acc64[(size_t)k * acc_stride + i]);
  }
}

// nsplit-slab epilogue of the PANEL kernels (csrc/biot_panel.cuh): out[i] = float(double(out[i]) + sign * sum over slices (in slice order) of partial).
__global__ void pp_finish_kernel(int nrows, int nsplit, int64_t nt, const double* partial, float* tu, float* tv, float* tw,
                                 float* tug, int64_t tug_stride, float sign) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nt) return;
  for (int k = 0; k < nrows; ++k) {
    double acc = 0.0;
    for (int s = 0; s < nsplit; ++s) acc += partial[((size_t)s * nrows + k) * nt + i];
    float* o = k == 0 ? tu + i : k == 1 ? tv + i : k == 2 ? tw + i : tug + (size_t)(k - 3) * tug_stride + i;
    *o = (float)((double)*o + (double)sign * acc);
  }
}

// =============================================================================================
// Packed-FP32 variant (sm_100 FFMA2 / FMUL2 / FADD2: fma.rn.f32x2): each instruction carries TWO
// sources against one target. Tile layout is pair-interleaved (built by pp_pack2_kernel):
//     q[4p+0] = { -x0, -x1, -y0, -y1 }   q[4p+1] = { -z0, -z1, r0^2, r1^2 }
//     q[4p+2] = { wx0, wx1, wy0, wy1 }   q[4p+3] = { wz0, wz1, 0, 0 }
// Positions are stored negated so d = t + (-s) is one FADD2 (the packed ops have no negate modifier
// in PTX). Kept as a measured alternative - see DESIGN.md section 5 for the verdict.
// =============================================================================================
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }

// Instruction ORDER matters here. A packed FP32 instruction holds the FMA pipe for 2 cycles and the register
// file feeds two fresh 64-bit operands in that time; a third distinct operand costs a third cycle unless the
// operand-reuse cache serves it, which happens when the PREVIOUS instruction used the same register in the same
// operand slot (tools/sass_rf_model.py; measured in profiles/r01_issue_model_microbench.txt). ptxas largely
// keeps the source order of independent instructions, so the sequence below is written as a chain in which
// every instruction shares slot 0 or slot 1 with its predecessor (marked <). Measured gain over formula order:
// 1.5-3 % (profiles/r01_variants_256k.txt).
//
// UNI: every source radius equals one value and every target radius equals one value (the usual case: particles
// are created with one core size), so r2 = sr^2 + tr^2 is a kernel-wide constant and the per-pair FADD2 goes.
// TRACE (O3D_PP_TRACE): the vortex-only gradient is trace free, d.(d x w) = 0, so the wz slot is recovered in
// the epilogue as -(ux + vy) instead of being accumulated: one FFMA2 less per interaction.
template <bool GRAD, bool UNI>
__device__ __forceinline__ void pp_interact2(const float4 q0, const float4 q1, const float4 q2, const float4 q3,
                                             const float2 tx, const float2 ty, const float2 tz, const float2 tr2,
                                             const float k15, const float k75, float2 (&acc)[PPAcc<GRAD>::N]) {
#if O3D_PP_POW
#if O3D_PP_POW == 2
  // k15 = 1.5 r2, k75 = -7.5 r2 (UNI only): scalar constants, which ptxas reads as 32-bit broadcast operands (.F32) instead of
  // 64-bit pairs. They arrive as arguments, formed once per kernel and passed through a warp reduction (pp2_walk): inside the
  // loop they would be re-derived from r2 on every trip (two FMUL, and moves between the register files) to save two registers
  const float2 tk = f2(k15, k15), tk5 = f2(k75, k75);
#else
  const float2 tk = __fmul2_rn(f2(1.5f, 1.5f), tr2), tk5 = __fmul2_rn(f2(-7.5f, -7.5f), tr2);   // UNI only; loop-invariant (hoisted)
#endif
#endif
#if O3D_PP_POW == 2
  // the velocity+gradient, uniform-radius body lives in its own file: its statement ORDER is machine-searched
  // (tools/tune_order.py overrides O3D_PP_BODY_FILE with candidate orders)
  if constexpr (GRAD && UNI) {
#include O3D_PP_BODY_FILE
    return;
  }
  // the other three instantiations, each in its own searched order (same statements as the formula order below, which
  // is what O3D_PP_POW < 2 builds and what documents the dataflow)
  if constexpr (GRAD && !UNI) {
#include O3D_PP_BODY_FILE_GEN
    return;
  }
  if constexpr (!GRAD && UNI) {
#include O3D_PP_BODY_FILE_VEL
    return;
  }
  if constexpr (!GRAD && !UNI) {
#include O3D_PP_BODY_FILE_VELGEN
    return;
  }
#endif
  const float2 dx = __fadd2_rn(tx, f2(q0.x, q0.y));
  const float2 dy = __fadd2_rn(ty, f2(q0.z, q0.w));
  const float2 dz = __fadd2_rn(tz, f2(q1.x, q1.y));
  const float2 r2 = UNI ? tr2 : __fadd2_rn(tr2, f2(q1.z, q1.w));   // UNI: tr2 already holds sr^2 + tr^2
  const float2 wx = f2(q2.x, q2.y), wy = f2(q2.z, q2.w), wz = f2(q3.x, q3.y);
  const float2 d2 = __ffma2_rn(dx, dx, __ffma2_rn(dy, dy, __ffma2_rn(dz, dz, r2)));
  const float2 rs = f2(rsqrt_approx(d2.x), rsqrt_approx(d2.y));
  const float2 rs2 = __fmul2_rn(rs, rs);
#if O3D_PP_POW
  // UNI: k = 1.5 r2 is a kernel-wide constant (tk), and with top = d2 + k:
  //   r3  = top rs^5           = rs^3 + k rs^5                  (d2 rs^5 = rs^3)
  //   bbb = rs^5 (2 - 5 top rs^2) = rs^5 (-3 - 5 k rs^2)
  // two packed instructions fewer than forming top, rs^4 and top*rs^2; every term keeps one sign (no cancellation).
  float2 dn5, r3, top;
  if constexpr (UNI) {
    const float2 rs3 = __fmul2_rn(rs2, rs);
    dn5 = __fmul2_rn(rs3, rs2);
    r3 = __ffma2_rn(tk, dn5, rs3);
  } else {
    top = __ffma2_rn(f2(1.5f, 1.5f), r2, d2);
    const float2 rs4 = __fmul2_rn(rs2, rs2);
    dn5 = __fmul2_rn(rs4, rs);
    r3 = __fmul2_rn(top, dn5);
  }
#else
  const float2 top = __ffma2_rn(f2(1.5f, 1.5f), r2, d2);
  const float2 rs4 = __fmul2_rn(rs2, rs2);
  const float2 dn5 = __fmul2_rn(rs4, rs);
  const float2 r3 = __fmul2_rn(top, dn5);
#endif
  // c = (dz wy - dy wz, dx wz - dz wx, dy wx - dx wy)
  const float2 t1 = __fmul2_rn(dy, wz);
  const float2 t2 = __fmul2_rn(dx, wz);                       // < wz
  const float2 t3 = __fmul2_rn(dx, wy);                       // < dx
  float2 cx = __ffma2_rn(dz, wy, neg2(t1));                   // < wy
  float2 cy = __ffma2_rn(neg2(dz), wx, t2);                   // < dz
  float2 cz = __ffma2_rn(dy, wx, neg2(t3));                   // < wx
  if constexpr (GRAD) {
    acc[12] = __ffma2_rn(r3, wx, acc[12]);                    // < wx
    acc[13] = __ffma2_rn(r3, wy, acc[13]);                    // < r3
    acc[14] = __ffma2_rn(r3, wz, acc[14]);                    // < r3
  }
  acc[0] = __ffma2_rn(r3, cx, acc[0]);                        // < r3
  acc[1] = __ffma2_rn(r3, cy, acc[1]);                        // < r3
  acc[2] = __ffma2_rn(r3, cz, acc[2]);                        // < r3
  if constexpr (GRAD) {
#if O3D_PP_POW
    float2 bbb;
    if constexpr (UNI) bbb = __fmul2_rn(dn5, __ffma2_rn(tk5, rs2, f2(-3.0f, -3.0f)));        // tk5 = -5 k
    else bbb = __fmul2_rn(dn5, __ffma2_rn(f2(-5.0f, -5.0f), __fmul2_rn(top, rs2), f2(2.0f, 2.0f)));
#else
    const float2 bbb = __fmul2_rn(dn5, __ffma2_rn(f2(-5.0f, -5.0f), __fmul2_rn(top, rs2), f2(2.0f, 2.0f)));
#endif
    cz = __fmul2_rn(bbb, cz);                                 // < cz
    cy = __fmul2_rn(bbb, cy);                                 // < bbb
    cx = __fmul2_rn(bbb, cx);                                 // < bbb
    acc[3]  = __ffma2_rn(dx, cx, acc[3]);
    acc[4]  = __ffma2_rn(dx, cy, acc[4]);                     // < dx
    acc[5]  = __ffma2_rn(dx, cz, acc[5]);                     // < dx
    acc[8]  = __ffma2_rn(dy, cz, acc[8]);                     // < cz   (snake through the 3x3 outer product)
    acc[7]  = __ffma2_rn(dy, cy, acc[7]);                     // < dy
    acc[6]  = __ffma2_rn(dy, cx, acc[6]);                     // < dy
    acc[9]  = __ffma2_rn(dz, cx, acc[9]);                     // < cx
    acc[10] = __ffma2_rn(dz, cy, acc[10]);                    // < dz
#if !O3D_PP_TRACE
    acc[11] = __ffma2_rn(dz, cz, acc[11]);                    // < dz
#endif
  }
}

// One tile of the walk, out of ring buffer BUF. BUF is a template argument - the walk alternates two copies of this body - so
// that every shared-memory address of the inner loop is [uniform base + immediate] with a base that does not depend on the tile:
// with a run-time buffer index ptxas re-derives the buffer's address (a chain of five uniform-datapath instructions) at the
// top of every trip of the inner loop, ahead of its first LDS, and at three warps per scheduler that latency is exposed (+3 %).
template <int BUF, int T, bool GRAD, bool UNI, int BLOCK>
__device__ __forceinline__ void pp2_tile(const PPArgs& p, PPWalk& w, PPBlock& s, float4 (&tile)[2][kTile * 2], PPSync& sy,
                                         const float r2u, const float k15, const float k75, float2 (&tx)[T], float2 (&ty)[T],
                                         float2 (&tz)[T], float2 (&tr2)[T],
                                         float2 (&acc)[T][PPAcc<GRAD>::N], double (&sum)[T][GRAD ? 12 : 3]) {
  constexpr int NA = PPAcc<GRAD>::N;
  constexpr int NS = GRAD ? 12 : 3;
  constexpr int U = GRAD ? kPPUnrollGrad : kPPUnrollVel;
  if (s.fresh) {
    const int64_t base = (int64_t)s.b * (BLOCK * T) + threadIdx.x;
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int64_t i = min(base + (int64_t)t * BLOCK, p.nt - 1);
      tx[t] = f2(p.tx[i], p.tx[i]); ty[t] = f2(p.ty[i], p.ty[i]); tz[t] = f2(p.tz[i], p.tz[i]);
      const float r = p.tr ? p.tr[i] : 0.0f;
      tr2[t] = UNI ? f2(r2u, r2u) : f2(r * r, r * r);
#pragma unroll
      for (int k = 0; k < NS; ++k) sum[t][k] = 0.0;
    }
  }
  pp_ring_wait(sy, w.kring);
#if O3D_PP_TWOBUF
  const float4* __restrict__ src = tile[BUF];
#else
  // one copy of the body for both buffers: the buffer's shared-memory address, formed once per tile and passed through a warp
  // reduction (REDUX writes a uniform register) so that ptxas holds it instead of re-deriving it on every trip of the loop
  const uint32_t sbase = __reduce_max_sync(0xffffffffu, smem_u32(tile[w.kring & 1]));
#endif
#pragma unroll(U)
  for (int j = 0; j < kTile / 2; ++j) {
#if O3D_PP_TWOBUF
    const float4 q0 = src[4 * j], q1 = src[4 * j + 1], q2 = src[4 * j + 2], q3 = src[4 * j + 3];
#else
    float4 q0, q1, q2, q3;
    const uint32_t sa = sbase + (uint32_t)j * 64u;
    asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(q0.x), "=f"(q0.y), "=f"(q0.z), "=f"(q0.w) : "r"(sa));
    asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+16];" : "=f"(q1.x), "=f"(q1.y), "=f"(q1.z), "=f"(q1.w) : "r"(sa));
    asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+32];" : "=f"(q2.x), "=f"(q2.y), "=f"(q2.z), "=f"(q2.w) : "r"(sa));
    asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+48];" : "=f"(q3.x), "=f"(q3.y), "=f"(q3.z), "=f"(q3.w) : "r"(sa));
#endif
#if O3D_PP_JOINT
    if constexpr (GRAD && UNI && T == 2) {
      const float2 tk = f2(k15, k15), tk5 = f2(k75, k75);          // 32-bit broadcast operands, as in pp_interact2
#include O3D_PP_JOINT_FILE
    } else
#endif
    {
#pragma unroll
      for (int t = 0; t < T; ++t) pp_interact2<GRAD, UNI>(q0, q1, q2, q3, tx[t], ty[t], tz[t], tr2[t], k15, k75, acc[t]);
    }
  }
#pragma unroll
  for (int t = 0; t < T; ++t) {
    float h[NA];
#pragma unroll
    for (int q = 0; q < NA; ++q) { h[q] = acc[t][q].x + acc[t][q].y; acc[t][q] = f2(0.f, 0.f); }
#if O3D_PP_TRACE
    if constexpr (GRAD) h[11] = -(h[3] + h[7]);
#endif
    pp_promote<GRAD>(h, sum[t]);
  }
#ifdef O3D_PP_ENDTIME
  if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && w.kring < 64 && (threadIdx.x >> 5) < 12) pp_tile_clock[threadIdx.x >> 5][w.kring] = clock64();
#endif
  constexpr bool NOBAR = (O3D_PP_NOBAR == 1 && GRAD) || O3D_PP_NOBAR == 2;
#if O3D_PP_TWOBUF
  pp_ring_refill<BLOCK, NOBAR>(p, w, tile[BUF], &sy.full[BUF], &sy.released[BUF]);         // ++w.kring
#else
  pp_ring_refill<BLOCK, NOBAR>(p, w, tile[w.kring & 1], &sy.full[w.kring & 1], &sy.released[w.kring & 1]);
#endif
  pp_segment_end<T, GRAD, BLOCK>(p, w, s, sum);
}

// The whole share of one persistent CTA: ONE loop over its tiles, two per trip (ring buffer 0, ring buffer 1), so that the nest
// is two deep - tiles, source pairs - like a one-block-per-CTA kernel's, which is the shape ptxas keeps the inner loop's counter
// and LDS addresses in uniform registers for; the target block changes inside that loop, at the tiles where a segment starts /
// ends. The UNI flag is warp-uniform: the two instantiations are two straight-line copies selected once per kernel.
template <int T, bool GRAD, bool UNI, int BLOCK>
__device__ __forceinline__ void pp2_walk(const PPArgs& p, PPWalk& w, float4 (&tile)[2][kTile * 2], PPSync& sy, const float r2u) {
  constexpr int NA = PPAcc<GRAD>::N;
  constexpr int NS = GRAD ? 12 : 3;
  float2 tx[T], ty[T], tz[T], tr2[T];
  double sum[T][NS];
  float2 acc[T][NA];
#pragma unroll
  for (int t = 0; t < T; ++t) {
#pragma unroll
    for (int k = 0; k < NA; ++k) acc[t][k] = f2(0.f, 0.f);
  }
  // the two constants of the uniform-radius core: the same value in every thread; a warp reduction (REDUX writes a uniform
  // register) hands them to ptxas as values it neither has to keep in the vector register file nor can re-derive in the loop
  const float k15 = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(1.5f * r2u)));
  const float k75 = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(-7.5f * r2u)));
  PPBlock s = pp_first_block(w);
  while (w.kring < w.nk) {
    pp2_tile<0, T, GRAD, UNI, BLOCK>(p, w, s, tile, sy, r2u, k15, k75, tx, ty, tz, tr2, acc, sum);
#if O3D_PP_TWOBUF
    if (w.kring < w.nk) pp2_tile<1, T, GRAD, UNI, BLOCK>(p, w, s, tile, sy, r2u, k15, k75, tx, ty, tz, tr2, acc, sum);
#endif
  }
}

template <int T, bool GRAD, int BLOCK>
__global__ void __launch_bounds__(BLOCK, kPPWarpsPerSM * 32 / BLOCK) pp2_kernel(const PPArgs p) {
  __shared__ alignas(128) float4 tile[2][kTile * 2];
  __shared__ alignas(8) PPSync sy;

#if O3D_PP_STAGGER
  {
    __shared__ unsigned slot;
    if (threadIdx.x == 0) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      slot = atomicAdd(&pp_stagger_arrivals[smid], 1u) % 3u;
    }
    __syncthreads();
    const long long until = clock64() + (long long)slot * O3D_PP_STAGGER;
    while (clock64() < until) {}
  }
#endif
#if O3D_PP_SKEW
  {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); } while (t1 - t0 < (unsigned long long)blockIdx.x * O3D_PP_SKEW);
  }
#endif
  PPWalk w = pp_ring_start<BLOCK>(p, tile, sy);

  // radius scan (pp_scan_kernel): [0] ~min, [1] max of sr^2 bit patterns over sources that carry strength,
  // [2] ~min, [3] max of tr bit patterns. Uniform <=> both ranges collapse and the constant r2 is positive.
  bool uni = false;
  float r2u = 0.0f;
  if (p.radius_range) {
    const uint32_t s0 = ~p.radius_range[0], s1 = p.radius_range[1], t0 = ~p.radius_range[2], t1 = p.radius_range[3];
    const float tr = p.tr ? __uint_as_float(t0) : 0.0f;
    r2u = __fadd_rn(__uint_as_float(s0), __fmul_rn(tr, tr));   // sr*sr + tr*tr, the reference's r2 (src/CoreFunc.h:267)
    uni = s0 == s1 && (!p.tr || t0 == t1) && r2u > 0.0f && s0 != 0xffffffffu;
  }
  if (uni) pp2_walk<T, GRAD, true, BLOCK>(p, w, tile, sy, r2u);
  else     pp2_walk<T, GRAD, false, BLOCK>(p, w, tile, sy, r2u);
#ifdef O3D_PP_ENDTIME
  if (threadIdx.x == 0 && blockIdx.x < 2048) {
    unsigned long long t; unsigned smid;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    pp_end_time[blockIdx.x] = t; pp_end_smid[blockIdx.x] = smid;
  }
#endif
}

// Radius ranges for the uniform-radius fast path of pp2_kernel. range[0..1]: min/max of the sr^2 bit patterns
// (non-negative floats order like unsigned integers) over records that carry strength - zero-strength records,
// padding included, contribute exactly 0 whatever their radius; range[2..3]: min/max of the tr bit patterns.
// The minima are kept COMPLEMENTED (range[0] = max of ~bits) so that the whole block starts as zeros: one
// cudaMemsetAsync, which - unlike a copy from pageable host memory - can be captured into a CUDA graph.
__global__ void pp_scan_kernel(int64_t nrec, const float4* packed, int64_t nt, const float* tr, uint32_t* range) {
  uint32_t smin = 0xffffffffu, smax = 0u, tmin = 0xffffffffu, tmax = 0u;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t pr = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pr < nrec / 2; pr += stride) {
    const float4 q1 = packed[4 * pr + 1], q2 = packed[4 * pr + 2], q3 = packed[4 * pr + 3];
    if (q2.x != 0.f || q2.z != 0.f || q3.x != 0.f) { const uint32_t b = __float_as_uint(q1.z); smin = min(smin, b); smax = max(smax, b); }
    if (q2.y != 0.f || q2.w != 0.f || q3.y != 0.f) { const uint32_t b = __float_as_uint(q1.w); smin = min(smin, b); smax = max(smax, b); }
  }
  if (tr)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nt; i += stride) {
      const uint32_t b = __float_as_uint(tr[i]);
      tmin = min(tmin, b); tmax = max(tmax, b);
    }
  smin = __reduce_min_sync(0xffffffffu, smin); smax = __reduce_max_sync(0xffffffffu, smax);
  tmin = __reduce_min_sync(0xffffffffu, tmin); tmax = __reduce_max_sync(0xffffffffu, tmax);
  if ((threadIdx.x & 31) == 0) {
    atomicMax(range + 0, ~smin); atomicMax(range + 1, smax);
    atomicMax(range + 2, ~tmin); atomicMax(range + 3, tmax);
  }
}

__global__ void pp_pack2_kernel(int64_t ns, int64_t ns_pad, const float* sx, const float* sy, const float* sz,
                                const float* sr, const float* wx, const float* wy, const float* wz, float4* out) {
  const int64_t pr = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // pair index
  if (2 * pr >= ns_pad) return;
  float x[2], y[2], z[2], r2[2], a[2], b[2], c[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int64_t j = 2 * pr + h;
    x[h] = y[h] = z[h] = 0.f; r2[h] = 1.f; a[h] = b[h] = c[h] = 0.f;
    if (j < ns) {
      const float r = sr ? sr[j] : 0.0f;
      x[h] = -sx[j]; y[h] = -sy[j]; z[h] = -sz[j]; r2[h] = r * r;
      a[h] = wx[j]; b[h] = wy[j]; c[h] = wz[j];
    }
  }
  out[4 * pr + 0] = make_float4(x[0], x[1], y[0], y[1]);
  out[4 * pr + 1] = make_float4(z[0], z[1], r2[0], r2[1]);
  out[4 * pr + 2] = make_float4(a[0], a[1], b[0], b[1]);
  out[4 * pr + 3] = make_float4(c[0], c[1], 0.f, 0.f);
}

}  // namespace o3d

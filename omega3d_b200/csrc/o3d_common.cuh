// o3d_common.cuh - shared device helpers for the sm_100a Biot-Savart kernels.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace o3d {

// ---------------------------------------------------------------------------------------------
// Packed source stream. The reference keeps sources as 7 separate float vectors
// (x[3], r, s[3]: src/Points.h:54-140). For the GPU they are re-packed once per evaluation into
// 32-byte records so that a whole tile is ONE contiguous block that a single bulk-async (TMA)
// copy can drop into shared memory:
//     rec[2*j+0] = { x, y, z, sr*sr }     (sr*sr is exactly the product the reference forms,
//     rec[2*j+1] = { wx, wy, wz, 0 }       src/CoreFunc.h:267)
// The stream is padded to a whole number of tiles with zero-strength records (they add exactly 0).
// ---------------------------------------------------------------------------------------------
#ifndef O3D_TILE
#define O3D_TILE 512
#endif
constexpr int kTile = O3D_TILE;                  // sources per shared-memory tile (512: 16 KB)
constexpr int kTileBytes = kTile * 32;

__host__ __device__ inline int64_t padded_sources(int64_t ns) { return ((ns + kTile - 1) / kTile) * kTile; }

// ---- mbarrier + bulk-async copy (cp.async.bulk => SASS UBLKCP), PTX ISA 8.x, sm_90+ ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

}  // namespace o3d

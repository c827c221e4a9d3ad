// pp_tuned.cu - the two product instantiations of pp2_kernel (csrc/biot_pp.cuh) and the six of ppc_kernel (the alternate
// core functions, csrc/biot_pp_cores.cuh) as a standalone cubin.
//
// Build (csrc/Makefile): nvcc -cubin -> lib/pp2_base.cubin -> tools/sass_patch.py -> lib/pp2_tuned.cubin -> embedded in
// libo3d_cuda.so as a byte array (lib/pp2_tuned_cubin.h) and loaded with cudaLibraryLoadData (capi.cu: TunedKernels).
// The post-pass only sets operand-reuse bits / clears yield hints on adjacent packed FP32 instructions: same
// instructions, same operands, same results bit for bit (tests/test_gpu_parity.py::test_tuned_kernels_bit_identical),
// fewer third register-file cycles.
#include "biot_pp_cores.cuh"

template __global__ void o3d::pp2_kernel<o3d::kPPTgrad, true, o3d::kPPBlock>(const o3d::PPArgs);
template __global__ void o3d::pp2_kernel<o3d::kPPTvel, false, o3d::kPPBlock>(const o3d::PPArgs);
// ... and as 128-thread CTAs for systems smaller than one product-size target block
template __global__ void o3d::pp2_kernel<o3d::kPPTgrad, true, o3d::kPPSmallBlock>(const o3d::PPArgs);
template __global__ void o3d::pp2_kernel<o3d::kPPTvel, false, o3d::kPPSmallBlock>(const o3d::PPArgs);

// the alternate core functions (o3d_cuda_set_core_func): Rosenhead-Moore, exponential, Vatistas n=2
template __global__ void o3d::ppc_kernel<o3d::kCoreRM, o3d::kPPTgrad, true, o3d::kPPBlock>(const o3d::PPArgs);
template __global__ void o3d::ppc_kernel<o3d::kCoreRM, o3d::kPPTvel, false, o3d::kPPBlock>(const o3d::PPArgs);
template __global__ void o3d::ppc_kernel<o3d::kCoreEXP, o3d::kPPTgrad, true, o3d::kPPBlock>(const o3d::PPArgs);
template __global__ void o3d::ppc_kernel<o3d::kCoreEXP, o3d::kPPTvel, false, o3d::kPPBlock>(const o3d::PPArgs);
template __global__ void o3d::ppc_kernel<o3d::kCoreV2, o3d::kPPTgrad, true, o3d::kPPBlock>(const o3d::PPArgs);
template __global__ void o3d::ppc_kernel<o3d::kCoreV2, o3d::kPPTvel, false, o3d::kPPBlock>(const o3d::PPArgs);

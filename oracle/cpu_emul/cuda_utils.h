/* TEST INFRASTRUCTURE — CPU emulation prelude for the reference's CUDA kernels.
 *
 * The reference kernels (u_net_arch/pt_custom_ops/_ext_src/src/*_gpu.cu) are plain grid-stride
 * CUDA C whose only device-specific constructs are __global__, __restrict__,
 * blockIdx/threadIdx/blockDim, atomicAdd and an in-kernel thrust::sort_by_key(thrust::device, ...).
 * This header shadows the reference's include/cuda_utils.h (it is first on the -I path) so that the
 * kernel *definitions* compile as ordinary C++ with Thrust's sequential CPP backend (the same stable
 * merge sort the device build resolves to without -rdc).  The kernel sources themselves are never
 * copied: oracle/Makefile streams them from /root/reference and cuts each host-side <<<...>>> wrapper.
 */
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
struct d3d_emul_dim3 { int x, y, z; };
extern thread_local d3d_emul_dim3 blockIdx, threadIdx, blockDim, gridDim;
#define __global__
#define __restrict__
#define __device__
#define __host__
static inline float atomicAdd(float* p, float v) { float o = *p; *p = o + v; return o; }
using std::floor;

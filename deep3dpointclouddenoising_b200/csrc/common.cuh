// Shared device helpers for the sm_100a kernels behind include/d3d_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/d3d_b200.h"

#define D3D_FULL_MASK 0xffffffffu

#define D3D_REQUIRE(cond) \
  do {                    \
    if (!(cond)) return D3D_ERR_BAD_ARG; \
  } while (0)

static inline int d3d_launch_status() { return (int)cudaGetLastError(); }

// Process-wide count of kernels this library has launched (d3d_kernel_launches(); bench.py reports it).
void d3d_note_launches(int n);

static inline int d3d_ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// Squared distance exactly as the reference kernels compute it on the device.  nvcc contracts
//   (qx-x)*(qx-x) + (qy-y)*(qy-y) + (qz-z)*(qz-z)
// (masked_ordered_ball_query_gpu.cu:57-58, masked_nearest_query_gpu.cu:48-49) into
//   FMUL dy*dy ; FFMA dx*dx + . ; FFMA dz*dz + .
// (read from the SASS of the reference kernels built for sm_100a).  Spelled with intrinsics so that
// no compiler flag can re-associate it: a different rounding flips neighbour sets at the radius.
__device__ __forceinline__ float d3d_dist2(float qx, float qy, float qz, float sx, float sy, float sz) {
  const float dx = __fsub_rn(qx, sx), dy = __fsub_rn(qy, sy), dz = __fsub_rn(qz, sz);
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// Valid-prefix length of every mask row: vlen[b] = index of the first 0 in mask[b, :] (N if none).
// All reference kernels stop scanning at the first 0 (see d3d_b200.h).  One block per cloud.
void d3d_launch_prefix_len(const int* mask, int B, int N, int* vlen, cudaStream_t st);

__device__ __forceinline__ int d3d_clamp_index(int i, int N) {
  // pt_utils.py:126-127 zeroes idx > N and idx < 0; idx == N cannot be produced by the ball query,
  // so the in-kernel clamp sends everything outside [0, N) to 0.
  return ((unsigned)i < (unsigned)N) ? i : 0;
}

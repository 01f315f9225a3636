// Spatial (Morton) processing order of a batch of point sets.
//
// The staged-tile aggregation kernels (pospool_tiles.cu) process 128 rows per CTA and stage the UNION of the rows
// their neighbourhoods gather; the union is small only if the 128 rows are close in space.  Clouds arrive in arbitrary
// index order (the reference's patches are sorted by distance from the patch centre, offset_dataset.py:630-656, the
// synthetic ones are shuffled), so every point set gets a permutation `order` once per forward: points sorted by the
// Morton code of their position on a 64^3 grid over the cloud's bounding box, ties by index.  The permutation only
// decides WHICH rows share a CTA — results are written back to the rows' own positions, nothing is reordered in HBM.
//
// One 1024-thread block per cloud: bounding box, 18-bit Morton code | 14-bit index as one 32-bit key, bitonic sort in
// shared memory (N <= 16384).
#include "common.cuh"

namespace {

constexpr int kMaxN = 16384;

__device__ __forceinline__ unsigned spread6(unsigned v) {  // 6 bits -> every third bit
  return (v & 1u) | ((v & 2u) << 2) | ((v & 4u) << 4) | ((v & 8u) << 6) | ((v & 16u) << 8) | ((v & 32u) << 10);
}

__global__ void __launch_bounds__(1024)
spatial_order_kernel(const float* __restrict__ xyz, int N, int npow2, int* __restrict__ order) {
  extern __shared__ unsigned keys[];
  __shared__ float red[6][32];
  __shared__ float box[4];  // min x, y, z, scale
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* P = xyz + (size_t)b * N * 3;
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int i = tid; i < N; i += 1024)
    for (int d = 0; d < 3; ++d) {
      const float v = P[3 * (size_t)i + d];
      lo[d] = fminf(lo[d], v);
      hi[d] = fmaxf(hi[d], v);
    }
  for (int d = 0; d < 3; ++d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[d] = fminf(lo[d], __shfl_xor_sync(D3D_FULL_MASK, lo[d], o));
      hi[d] = fmaxf(hi[d], __shfl_xor_sync(D3D_FULL_MASK, hi[d], o));
    }
    if (lane == 0) { red[d][warp] = lo[d]; red[3 + d][warp] = hi[d]; }
  }
  __syncthreads();
  if (warp == 0) {
    float ext = 0.f, mn[3];
    for (int d = 0; d < 3; ++d) {
      float a = red[d][lane], z = red[3 + d][lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a = fminf(a, __shfl_xor_sync(D3D_FULL_MASK, a, o));
        z = fmaxf(z, __shfl_xor_sync(D3D_FULL_MASK, z, o));
      }
      mn[d] = a;
      ext = fmaxf(ext, z - a);
    }
    if (lane == 0) {
      box[0] = mn[0]; box[1] = mn[1]; box[2] = mn[2];
      box[3] = (ext > 0.f && ext < INFINITY) ? 63.999f / ext : 0.f;
    }
  }
  __syncthreads();
  const float s = box[3];
  for (int i = tid; i < npow2; i += 1024) {
    unsigned key = 0xffffffffu;  // padding sorts last
    if (i < N) {
      const unsigned cx = (unsigned)min(max((int)((P[3 * (size_t)i] - box[0]) * s), 0), 63);
      const unsigned cy = (unsigned)min(max((int)((P[3 * (size_t)i + 1] - box[1]) * s), 0), 63);
      const unsigned cz = (unsigned)min(max((int)((P[3 * (size_t)i + 2] - box[2]) * s), 0), 63);
      key = ((spread6(cx) | (spread6(cy) << 1) | (spread6(cz) << 2)) << 14) | (unsigned)i;
    }
    keys[i] = key;
  }
  __syncthreads();
  for (int k = 2; k <= npow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (npow2 >> 1); t += 1024) {
        const int i = 2 * t - (t & (j - 1));
        const int l = i + j;
        const unsigned a = keys[i], c = keys[l];
        const bool up = (i & k) == 0;
        if ((a > c) == up) { keys[i] = c; keys[l] = a; }
      }
      __syncthreads();
    }
  }
  int* out = order + (size_t)b * N;
  for (int i = tid; i < N; i += 1024) out[i] = (int)(keys[i] & 0x3fffu);
}

}  // namespace

extern "C" {

int d3d_spatial_order(const float* xyz, int B, int N, int* order, void* stream) {
  D3D_REQUIRE(xyz && order);
  D3D_REQUIRE(B >= 0 && N > 0);
  if (N > kMaxN) return D3D_ERR_UNSUPPORTED;
  if (B == 0) return 0;
  int npow2 = 2;
  while (npow2 < N) npow2 <<= 1;
  const size_t smem = (size_t)npow2 * sizeof(unsigned);
  cudaError_t e = cudaFuncSetAttribute(spatial_order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  spatial_order_kernel<<<B, 1024, smem, (cudaStream_t)stream>>>(xyz, N, npow2, order);
  d3d_note_launches(1);
  return d3d_launch_status();
}

}  // extern "C"

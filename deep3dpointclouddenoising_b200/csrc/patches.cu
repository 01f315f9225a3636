// Full-shape inference support: patch centres, radius patches, vote averaging.
//
//   ref: offset_dataset.py:540-561   patch centres = CPU grid_subsampling(cloud, 0.05) barycentres -> nearest real point
//   ref: u_net_arch/cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.cpp:25-103   voxel arithmetic
//   ref: offset_dataset.py:630-656   patch = KDTree.query_radius(centre, r, sort_results=True)[:num_points]
//   ref: u_net_arch/qualitative_inference_test.py:325-342   votes: sum of predictions per point / (count + 1e-7)
#include "common.cuh"

namespace {

// voxel id of every point with the reference's float expressions; origin/NX/NY are computed on the host
__global__ void voxel_id_kernel(const float* __restrict__ pts, int N, float ox, float oy, float oz, float dl, int NX, int NY,
                                int n_cells, int* __restrict__ ids) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  // grid_subsampling.cpp:52-55  iX = floor((p.x - origin.x) / sampleDl)   (host build: separate sub, IEEE div)
  const int ix = (int)floorf(__fdiv_rn(__fsub_rn(pts[3 * (size_t)i], ox), dl));
  const int iy = (int)floorf(__fdiv_rn(__fsub_rn(pts[3 * (size_t)i + 1], oy), dl));
  const int iz = (int)floorf(__fdiv_rn(__fsub_rn(pts[3 * (size_t)i + 2], oz), dl));
  const long long id = (long long)ix + (long long)NX * iy + (long long)NX * NY * iz;
  ids[i] = (int)min(max(id, 0ll), (long long)n_cells - 1);
}

// thread per cell: members in ascending point index (the inverse map is sorted), float sum, times 1/count
__global__ void barycentre_kernel(const float* __restrict__ pts, const int* __restrict__ rowptr, const int* __restrict__ entries,
                                  int n_cells, float* __restrict__ bary, int* __restrict__ counts) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  const int beg = rowptr[c], end = rowptr[c + 1];
  float sx = 0.f, sy = 0.f, sz = 0.f;
  for (int e = beg; e < end; ++e) {
    const int i = entries[e] >> 8;
    sx = __fadd_rn(sx, pts[3 * (size_t)i]); sy = __fadd_rn(sy, pts[3 * (size_t)i + 1]); sz = __fadd_rn(sz, pts[3 * (size_t)i + 2]);
  }
  const int n = end - beg;
  counts[c] = n;
  if (n > 0) {
    const float inv = (float)(1.0 / (double)n);  // grid_subsampling.cpp:87  point * (1.0 / count)
    bary[3 * (size_t)c] = __fmul_rn(sx, inv); bary[3 * (size_t)c + 1] = __fmul_rn(sy, inv); bary[3 * (size_t)c + 2] = __fmul_rn(sz, inv);
  }
}

// ---- radius patches: block per centre over a uniform grid of the cloud -----------------------------------------
struct PatchGrid {
  float min_x, min_y, min_z, cell, inv_cell;
  int G;
};

constexpr int kPatchThreads = 512;
constexpr int kPatchSmemKeys = 12288;  // candidates sorted in shared memory (d2 as fp64 bits + index: 12 B each)
constexpr int kSelBins = 512;          // selection histogram (few results out of many candidates)
constexpr int kSelKeys = 2048;         // capacity of the selected set

__device__ __forceinline__ int pcoord(float v, float mn, float inv_cell, int G) {
  return min(max((int)floorf((v - mn) * inv_cell), 0), G - 1);
}

__global__ void __launch_bounds__(kPatchThreads)
radius_patch_kernel(const float* __restrict__ centres, int P, float radius, int num_points, PatchGrid g,
                    const int* __restrict__ cell_start, int n_cells, int N, const float4* __restrict__ sorted,
                    int* __restrict__ out_idx /* (P, num_points), -1 padded */, int* __restrict__ out_count,
                    unsigned long long* __restrict__ scratch_keys, int* __restrict__ scratch_idx, int scratch_stride,
                    int smem_keys) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ int n_cand;
  const int pidx = blockIdx.x;
  const float cx = centres[3 * (size_t)pidx], cy = centres[3 * (size_t)pidx + 1], cz = centres[3 * (size_t)pidx + 2];
  unsigned long long* keys = scratch_stride > 0 ? scratch_keys + (size_t)pidx * scratch_stride : reinterpret_cast<unsigned long long*>(smem);
  int* vals = scratch_stride > 0 ? scratch_idx + (size_t)pidx * scratch_stride : reinterpret_cast<int*>(smem + (size_t)smem_keys * 8);
  const int cap = scratch_stride > 0 ? scratch_stride : smem_keys;
  if (threadIdx.x == 0) n_cand = 0;
  __syncthreads();
  const double r2 = (double)radius * (double)radius;
  const int x0 = pcoord(cx - radius, g.min_x, g.inv_cell, g.G), x1 = pcoord(cx + radius, g.min_x, g.inv_cell, g.G);
  const int y0 = pcoord(cy - radius, g.min_y, g.inv_cell, g.G), y1 = pcoord(cy + radius, g.min_y, g.inv_cell, g.G);
  const int z0 = pcoord(cz - radius, g.min_z, g.inv_cell, g.G), z1 = pcoord(cz + radius, g.min_z, g.inv_cell, g.G);
  const int nx = x1 - x0 + 1, ny = y1 - y0 + 1, nz = z1 - z0 + 1;
  // cells of one x-row are contiguous in memory: a (y, z) row is one contiguous range of sorted points.  Rows are
  // short (a few points), so every thread takes whole rows.
  for (int row = threadIdx.x; row < ny * nz; row += kPatchThreads) {
    const int y = y0 + row % ny, z = z0 + row / ny;
    const int c0 = (z * g.G + y) * g.G + x0, c1 = c0 + nx;
    const int beg = cell_start[c0], end = c1 < n_cells ? cell_start[c1] : N;
    for (int t = beg; t < end; ++t) {
      const float4 s = __ldg(sorted + t);
      // distances in fp64 like the KD-tree the reference uses (sklearn promotes to float64)
      const double dx = (double)s.x - (double)cx, dy = (double)s.y - (double)cy, dz = (double)s.z - (double)cz;
      const double d2 = dx * dx + dy * dy + dz * dz;
      if (d2 <= r2) {
        const int pos = atomicAdd(&n_cand, 1);
        if (pos < cap) { keys[pos] = (unsigned long long)__double_as_longlong(d2); vals[pos] = __float_as_int(s.w); }
      }
    }
  }
  __syncthreads();
  int n = min(n_cand, cap);
  // ---- few results out of many candidates (neighbour lists: 26-52 nearest of hundreds to thousands in the ball): select
  // before sorting.  A 512-bin histogram of d2 / r2 (monotone in the key) gives the bin b* where the cumulative count
  // reaches num_points; every candidate of a bin <= b* is copied to a small buffer and only that buffer is sorted —
  // the same first num_points entries as the full sort (ties in b* are all kept and ordered by the sort).
  __shared__ int hist[kSelBins];
  __shared__ int sel_bin, sel_n;
  __shared__ unsigned long long sel_keys[kSelKeys];
  __shared__ int sel_vals[kSelKeys];
  if (n_cand <= cap && n > 4 * num_points && n > 256) {
    for (int i = threadIdx.x; i < kSelBins; i += kPatchThreads) hist[i] = 0;
    if (threadIdx.x == 0) { sel_bin = kSelBins - 1; sel_n = 0; }
    __syncthreads();
    const double to_bin = (double)kSelBins / r2;
    for (int i = threadIdx.x; i < n; i += kPatchThreads)
      atomicAdd(&hist[min(kSelBins - 1, (int)(__longlong_as_double((long long)keys[i]) * to_bin))], 1);
    __syncthreads();
    if (threadIdx.x < 32) {  // inclusive scan of the 512 bins by one warp, 16 bins per lane
      int local[kSelBins / 32], sum = 0;
#pragma unroll
      for (int k = 0; k < kSelBins / 32; ++k) { sum += hist[threadIdx.x * (kSelBins / 32) + k]; local[k] = sum; }
      int incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)threadIdx.x >= o) incl += up;
      }
      const int base = incl - sum;
#pragma unroll
      for (int k = 0; k < kSelBins / 32; ++k) {
        const int c = base + local[k], before = c - hist[threadIdx.x * (kSelBins / 32) + k];
        if (before < num_points && c >= num_points) sel_bin = threadIdx.x * (kSelBins / 32) + k;
      }
      __syncwarp();
      int upto = 0;  // candidates in bins <= sel_bin
      const int sb = sel_bin;
#pragma unroll
      for (int k = 0; k < kSelBins / 32; ++k)
        if (threadIdx.x * (kSelBins / 32) + k <= sb) upto = base + local[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) upto = max(upto, __shfl_xor_sync(0xffffffffu, upto, o));
      if (threadIdx.x == 0) sel_n = upto;
    }
    __syncthreads();
    if (sel_n <= kSelKeys) {  // (a bin with thousands of equal distances would not fit: the full sort handles it)
      __shared__ int fill;
      if (threadIdx.x == 0) fill = 0;
      __syncthreads();
      const int sb = sel_bin;
      for (int i = threadIdx.x; i < n; i += kPatchThreads) {
        const unsigned long long key = keys[i];
        if (min(kSelBins - 1, (int)(__longlong_as_double((long long)key) * to_bin)) <= sb) {
          const int pos = atomicAdd(&fill, 1);
          sel_keys[pos] = key; sel_vals[pos] = vals[i];
        }
      }
      __syncthreads();
      n = sel_n;
      keys = sel_keys; vals = sel_vals;
    }
  }
  int P2 = 1;
  while (P2 < n) P2 <<= 1;
  if (keys != sel_keys) P2 = min(P2, cap);  // cap is a power of two or the candidates fit
  for (int i = n + threadIdx.x; i < P2; i += kPatchThreads) { keys[i] = ~0ull; vals[i] = 0x7fffffff; }
  __syncthreads();
  // bitonic sort by (distance, index): ties (duplicate points) resolve to the lower index
  for (int k = 2; k <= P2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (P2 >> 1); t += kPatchThreads) {
        const int a = ((t & ~(j - 1)) << 1) | (t & (j - 1)), c = a | j;
        const unsigned long long ka = keys[a], kc = keys[c];
        const int va = vals[a], vc = vals[c];
        const bool greater = ka > kc || (ka == kc && va > vc);
        if (greater == ((a & k) == 0)) { keys[a] = kc; keys[c] = ka; vals[a] = vc; vals[c] = va; }
      }
      __syncthreads();
    }
  const int take = min(min(n_cand, cap), num_points);
  for (int i = threadIdx.x; i < num_points; i += kPatchThreads) out_idx[(size_t)pidx * num_points + i] = i < take ? vals[i] : -1;
  if (threadIdx.x == 0) out_count[pidx] = n_cand;  // > cap means the candidate buffer overflowed (caller re-runs with scratch)
}

// votes: thread per cloud point over its inverse-map segment; entry (j, k) -> flat slot j * 128 + k of pred (P, 3, n)
__global__ void vote_kernel(const float* __restrict__ pred, const int* __restrict__ rowptr, const int* __restrict__ entries,
                            int N, int num_points, float* __restrict__ mean_offset, float* __restrict__ votes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int beg = rowptr[i], end = rowptr[i + 1];
  float sx = 0.f, sy = 0.f, sz = 0.f;
  for (int e = beg; e < end; ++e) {
    const int packed = entries[e];
    const long long flat = (long long)(packed >> 8) * 128 + (packed & 255);
    const long long patch = flat / num_points, slot = flat - patch * num_points;
    const float* p = pred + (size_t)patch * 3 * num_points + slot;
    sx += p[0]; sy += p[num_points]; sz += p[2 * (size_t)num_points];
  }
  const float cnt = (float)(end - beg) + 1e-7f;  // qualitative_inference_test.py:286,342
  mean_offset[3 * (size_t)i] = sx / cnt; mean_offset[3 * (size_t)i + 1] = sy / cnt; mean_offset[3 * (size_t)i + 2] = sz / cnt;
  if (votes) votes[i] = (float)(end - beg);
}

}  // namespace

// grid construction shared with chamfer.cu
int d3d_internal_build_grid(const float* s, int N, int G_override, void* ws, cudaStream_t st, void** gp_dev, int** cell_start,
                            float4** sorted, int* G_out);
size_t d3d_internal_grid_bytes(int N, int G_override);

extern "C" {

int d3d_voxel_ids(const float* points, int N, float origin_x, float origin_y, float origin_z, float dl, int NX, int NY,
                  int n_cells, int* ids, void* stream) {
  D3D_REQUIRE(points && ids && N > 0 && dl > 0.f && NX > 0 && NY > 0 && n_cells > 0);
  voxel_id_kernel<<<d3d_ceil_div(N, 256), 256, 0, (cudaStream_t)stream>>>(points, N, origin_x, origin_y, origin_z, dl, NX, NY,
                                                                        n_cells, ids);
  d3d_note_launches(1);
  return d3d_launch_status();
}

int d3d_voxel_barycentres(const float* points, const int* rowptr, const int* entries, int n_cells, float* bary, int* counts,
                          void* stream) {
  D3D_REQUIRE(points && rowptr && entries && bary && counts && n_cells > 0);
  barycentre_kernel<<<d3d_ceil_div(n_cells, 128), 128, 0, (cudaStream_t)stream>>>(points, rowptr, entries, n_cells, bary, counts);
  d3d_note_launches(1);
  return d3d_launch_status();
}

int d3d_radius_patches_tier(const float* points, int N, const float* centres, int P, float radius, int num_points,
                            int smem_keys, int overflow_stride, int* out_idx, int* out_count, void* ws, size_t ws_bytes,
                            void* stream);

size_t d3d_radius_patches_workspace_bytes(int N, int P, int overflow_stride) {
  if (N <= 0 || P <= 0) return 0;
  size_t b = d3d_internal_grid_bytes(N, 0);
  if (overflow_stride > 0) b += (size_t)P * overflow_stride * 12 + 512;
  return b;
}

/* out_idx (P, num_points): indices of the points within `radius` of each centre, ascending distance, -1 padded;
 * out_count (P): number of points in the ball.  overflow_stride = 0 sorts in shared memory (<= 12288 candidates per
 * patch); if some out_count exceeds that, call again with overflow_stride = a power of two >= max(out_count). */
int d3d_radius_patches(const float* points, int N, const float* centres, int P, float radius, int num_points,
                       int overflow_stride, int* out_idx, int* out_count, void* ws, size_t ws_bytes, void* stream) {
  return d3d_radius_patches_tier(points, N, centres, P, radius, num_points, kPatchSmemKeys, overflow_stride, out_idx, out_count,
                                 ws, ws_bytes, stream);
}

/* smem_keys: candidates a block can hold in shared memory — 12288 (one block per SM), or a smaller power of two
 * (>= 256) when the balls are known to be small: 12 bytes per key, so 2048 keys let 6 blocks share an SM. */
int d3d_radius_patches_tier(const float* points, int N, const float* centres, int P, float radius, int num_points,
                            int smem_keys, int overflow_stride, int* out_idx, int* out_count, void* ws, size_t ws_bytes,
                            void* stream) {
  D3D_REQUIRE(points && centres && out_idx && out_count && N > 0 && P > 0 && radius > 0.f && num_points > 0);
  D3D_REQUIRE(smem_keys == kPatchSmemKeys || (smem_keys >= 256 && smem_keys < kPatchSmemKeys && (smem_keys & (smem_keys - 1)) == 0));
  D3D_REQUIRE(overflow_stride == 0 || (overflow_stride & (overflow_stride - 1)) == 0);
  if (!ws || ws_bytes < d3d_radius_patches_workspace_bytes(N, P, overflow_stride)) return D3D_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  void* gp_dev; int* cell_start; float4* sorted; int G;
  const int rc = d3d_internal_build_grid(points, N, 0, ws, st, &gp_dev, &cell_start, &sorted, &G);
  if (rc != 0) return rc;
  // the kernel needs the grid parameters by value: they live on the device, so fetch them (inference set-up path)
  PatchGrid g;
  cudaError_t e = cudaMemcpyAsync(&g, gp_dev, sizeof(PatchGrid), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return (int)e;
  unsigned long long* sk = nullptr;
  int* si = nullptr;
  size_t smem = (size_t)smem_keys * 12;
  if (overflow_stride > 0) {
    unsigned char* p = (unsigned char*)ws + d3d_internal_grid_bytes(N, 0);
    p = (unsigned char*)(((uintptr_t)p + 255) & ~(uintptr_t)255);
    sk = (unsigned long long*)p;
    si = (int*)(p + (size_t)P * overflow_stride * 8);
    smem = 0;
  }
  e = cudaFuncSetAttribute(radius_patch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)kPatchSmemKeys * 12));
  if (e != cudaSuccess) return (int)e;
  radius_patch_kernel<<<P, kPatchThreads, smem, st>>>(centres, P, radius, num_points, g, cell_start, G * G * G, N, sorted, out_idx,
                                                      out_count, sk, si, overflow_stride, smem_keys);
  d3d_note_launches(1);
  return d3d_launch_status();
}

int d3d_vote_mean(const float* pred, const int* rowptr, const int* entries, int N, int num_points, float* mean_offset,
                  float* votes, void* stream) {
  D3D_REQUIRE(pred && rowptr && entries && mean_offset && N > 0 && num_points > 0);
  vote_kernel<<<d3d_ceil_div(N, 256), 256, 0, (cudaStream_t)stream>>>(pred, rowptr, entries, N, num_points, mean_offset, votes);
  d3d_note_launches(1);
  return d3d_launch_status();
}

}  // extern "C"

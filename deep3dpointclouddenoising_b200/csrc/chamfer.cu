// Exact nearest-neighbour squared distances on large clouds and the Chamfer distance built on them
// (SURVEY.md §8 row f3).
//
//   ref: u_net_arch/compute_cd.py:74-75  chamfer_distance(clean, denoised, batch_reduction="mean",
//        point_reduction="mean", norm_type="L2")
//   ref: u_net_arch/models/losses/chamfer_distance_aux.py:154-155,166-167,216-246: two K=1 nearest-neighbour
//        searches (pytorch3d.ops.knn_points, squared distances), mean over the points of each cloud, sum of both.
// pytorch3d's knn is brute force: 10^12 pair tests for two 1M-point clouds.  Here the supports are binned into a
// uniform grid (counting sort by cell: integer histogram, device-wide scan, cursor fill) and every query walks
// cube shells of cells around its own cell until the best distance found is <= the distance to the next shell:
// exact, and ~30-300 pair tests per query on surface-like clouds.  Ties in distance resolve to the lowest support
// index, so the result does not depend on the (atomic) order inside a cell.
#include "common.cuh"

namespace {

struct GridParams {
  float min_x, min_y, min_z;
  float cell, inv_cell;
  int G;
};

__device__ __forceinline__ float warp_min_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(D3D_FULL_MASK, v, o));
  return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(D3D_FULL_MASK, v, o));
  return v;
}

// one block: bounding box of the supports -> grid parameters (cubic cells, G per axis)
__global__ void __launch_bounds__(1024)
bbox_kernel(const float* __restrict__ s, int N, int G, GridParams* __restrict__ gp) {
  __shared__ float red[6][32];
  float lo[3] = {s[0], s[1], s[2]}, hi[3] = {s[0], s[1], s[2]};
  for (int i = threadIdx.x; i < N; i += blockDim.x)
    for (int d = 0; d < 3; ++d) {
      const float v = s[3 * (size_t)i + d];
      lo[d] = fminf(lo[d], v);
      hi[d] = fmaxf(hi[d], v);
    }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int d = 0; d < 3; ++d) {
    lo[d] = warp_min_f(lo[d]);
    hi[d] = warp_max_f(hi[d]);
    if (lane == 0) { red[d][warp] = lo[d]; red[3 + d][warp] = hi[d]; }
  }
  __syncthreads();
  if (warp == 0) {
    float ext = 0.f, mn[3];
    for (int d = 0; d < 3; ++d) {
      const float a = warp_min_f(red[d][lane]), b = warp_max_f(red[3 + d][lane]);
      mn[d] = a;
      ext = fmaxf(ext, b - a);
    }
    if (lane == 0) {
      const float cell = fmaxf(ext, 1e-12f) * (1.0f + 1e-5f) / (float)G;
      gp->min_x = mn[0]; gp->min_y = mn[1]; gp->min_z = mn[2];
      gp->cell = cell;
      gp->inv_cell = 1.0f / cell;
      gp->G = G;
    }
  }
}

__device__ __forceinline__ int cell_coord(float v, float mn, float inv_cell, int G) {
  const int c = (int)floorf((v - mn) * inv_cell);
  return min(max(c, 0), G - 1);
}

__global__ void cell_count_kernel(const float* __restrict__ s, int N, const GridParams* __restrict__ gp,
                                  int* __restrict__ cell_of, int* __restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const GridParams p = *gp;
  const int cx = cell_coord(s[3 * (size_t)i], p.min_x, p.inv_cell, p.G), cy = cell_coord(s[3 * (size_t)i + 1], p.min_y, p.inv_cell, p.G),
            cz = cell_coord(s[3 * (size_t)i + 2], p.min_z, p.inv_cell, p.G);
  const int c = (cz * p.G + cy) * p.G + cx;
  cell_of[i] = c;
  atomicAdd(&counts[c], 1);
}

// device-wide exclusive scan in three launches: per-block scan + block totals, scan of the totals, add back
__global__ void __launch_bounds__(1024)
scan_blocks_kernel(const int* __restrict__ in, int n, int* __restrict__ out, int* __restrict__ block_tot) {
  __shared__ int warp_tot[32];
  const int i = blockIdx.x * 1024 + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int v = i < n ? in[i] : 0;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int up = __shfl_up_sync(D3D_FULL_MASK, incl, o);
    if (lane >= o) incl += up;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const int w = warp_tot[lane];
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(D3D_FULL_MASK, wi, o);
      if (lane >= o) wi += up;
    }
    warp_tot[lane] = wi - w;
    if (lane == 31) block_tot[blockIdx.x] = wi;
  }
  __syncthreads();
  if (i < n) out[i] = warp_tot[warp] + incl - v;
}

__global__ void __launch_bounds__(1024)
scan_totals_kernel(int* __restrict__ block_tot, int nblk) {  // single block, in place, exclusive
  __shared__ int warp_tot[32];
  __shared__ int carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblk; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < nblk ? block_tot[i] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(D3D_FULL_MASK, incl, o);
      if (lane >= o) incl += up;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const int w = warp_tot[lane];
      int wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(D3D_FULL_MASK, wi, o);
        if (lane >= o) wi += up;
      }
      warp_tot[lane] = wi - w;
    }
    __syncthreads();
    const int excl = carry + warp_tot[warp] + incl - v;
    if (i < nblk) block_tot[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
}

__global__ void scan_add_kernel(int* __restrict__ out, int n, const int* __restrict__ block_tot, int* __restrict__ cursor) {
  const int i = blockIdx.x * 1024 + threadIdx.x;
  if (i >= n) return;
  const int v = out[i] + block_tot[blockIdx.x];
  out[i] = v;
  cursor[i] = v;
}

__global__ void cell_fill_kernel(const float* __restrict__ s, int N, const int* __restrict__ cell_of, int* __restrict__ cursor,
                                 float4* __restrict__ sorted /* xyz + original index */) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int pos = atomicAdd(&cursor[cell_of[i]], 1);
  sorted[pos] = make_float4(s[3 * (size_t)i], s[3 * (size_t)i + 1], s[3 * (size_t)i + 2], __int_as_float(i));
}

// thread per query: expanding cube shells of cells.  kDouble: distances in fp64 (what a KD-tree on float64 copies of
// the data decides, used where the INDEX must match the reference's sklearn search; fp32 otherwise).
template <bool kDouble>
__global__ void __launch_bounds__(128)
nn_query_kernel(const float* __restrict__ q, int M, const GridParams* __restrict__ gp, const int* __restrict__ cell_start,
                int n_cells, int N, const float4* __restrict__ sorted, float* __restrict__ out_d2, int* __restrict__ out_idx) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  const GridParams p = *gp;
  const int G = p.G;
  const float qx = q[3 * (size_t)j], qy = q[3 * (size_t)j + 1], qz = q[3 * (size_t)j + 2];
  const int cx = cell_coord(qx, p.min_x, p.inv_cell, G), cy = cell_coord(qy, p.min_y, p.inv_cell, G), cz = cell_coord(qz, p.min_z, p.inv_cell, G);
  float best = INFINITY;
  double best_d = INFINITY;
  int best_i = -1;
  auto visit = [&](int x, int y, int z) {
    const int c = (z * G + y) * G + x;
    const int beg = cell_start[c], end = c + 1 < n_cells ? cell_start[c + 1] : N;
    for (int t = beg; t < end; ++t) {
      const float4 s = __ldg(sorted + t);
      const int si = __float_as_int(s.w);
      if (kDouble) {
        const double dx = (double)qx - (double)s.x, dy = (double)qy - (double)s.y, dz = (double)qz - (double)s.z;
        const double d2 = dx * dx + dy * dy + dz * dz;
        if (d2 < best_d || (d2 == best_d && si < best_i)) { best_d = d2; best = (float)d2; best_i = si; }
      } else {
        const float dx = qx - s.x, dy = qy - s.y, dz = qz - s.z;
        const float d2 = dx * dx + dy * dy + dz * dz;
        if (d2 < best || (d2 == best && si < best_i)) { best = d2; best_i = si; }
      }
    }
  };
  for (int r = 0; r < G; ++r) {
    // The query (or, when it lies outside the grid box, its projection onto the box, which is never farther from a
    // support than the query itself) is inside the centre cell, so every support of shell r is farther than
    // (r - 1) * cell: once best <= ((r - 1) * cell)^2 no outer shell can win.
    // Compared in fp64 with a 1e-6 relative margin (like ball_nearest_kernel): an fp32 bound can cut the search one
    // shell short when the winner sits exactly on a cell boundary.
    if (r >= 1) {
      const double reach = (double)(r - 1) * (double)p.cell * (1.0 - 1e-6);
      if ((double)best <= reach * reach) break;
    }
    if (cx - r < 0 && cx + r >= G && cy - r < 0 && cy + r >= G && cz - r < 0 && cz + r >= G) break;  // shell outside the grid
    for (int z = max(cz - r, 0); z <= min(cz + r, G - 1); ++z) {
      const bool z_face = (z == cz - r) || (z == cz + r);
      for (int y = max(cy - r, 0); y <= min(cy + r, G - 1); ++y) {
        if (z_face || y == cy - r || y == cy + r) {
          for (int x = max(cx - r, 0); x <= min(cx + r, G - 1); ++x) visit(x, y, z);
        } else {  // interior row of the cube: only its two end cells belong to the shell
          if (cx - r >= 0) visit(cx - r, y, z);
          if (cx + r < G) visit(cx + r, y, z);
        }
      }
    }
  }
  out_d2[j] = best;
  if (out_idx) out_idx[j] = best_i;
}

// deterministic mean: per-block fp64 partial sums, then one block adds them in order
__global__ void __launch_bounds__(256)
sum_partials_kernel(const float* __restrict__ v, int n, double* __restrict__ partial) {
  __shared__ double red[8];
  double s = 0.0;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) s += (double)v[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(D3D_FULL_MASK, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    partial[blockIdx.x] = t;
  }
}

__global__ void chamfer_finish_kernel(const double* __restrict__ px, int nbx, int nx, const double* __restrict__ py, int nby,
                                      int ny, float* __restrict__ out) {
  double sx = 0.0, sy = 0.0;
  for (int i = 0; i < nbx; ++i) sx += px[i];
  for (int i = 0; i < nby; ++i) sy += py[i];
  const double cx = sx / (double)nx, cy = sy / (double)ny;
  out[0] = (float)(cx + cy);  // chamfer_distance_aux.py:244  cham_dist = cham_x + cham_y
  out[1] = (float)cx;
  out[2] = (float)cy;
}

int grid_size_for(int N) {
  int G = (int)sqrt((double)N / 8.0);
  if (G < 4) G = 4;
  if (G > 256) G = 256;
  return G;
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct NnWs {
  GridParams* gp;
  int* cell_of;
  int* cell_start;
  int* cursor;
  int* block_tot;
  float4* sorted;
  size_t bytes;
};

NnWs carve_nn(void* ws, int N) {
  const int G = grid_size_for(N);
  const size_t cells = (size_t)G * G * G;
  unsigned char* p = (unsigned char*)ws;
  NnWs w;
  w.gp = (GridParams*)p; p += 256;
  w.cell_of = (int*)p; p += align256((size_t)N * sizeof(int));
  w.cell_start = (int*)p; p += align256(cells * sizeof(int));
  w.cursor = (int*)p; p += align256(cells * sizeof(int));
  w.block_tot = (int*)p; p += align256(((cells + 1023) / 1024 + 1) * sizeof(int));
  w.sorted = (float4*)p; p += align256((size_t)N * sizeof(float4));
  w.bytes = (size_t)(p - (unsigned char*)ws);
  return w;
}

int build_grid(const float* s, int N, void* ws, cudaStream_t st, NnWs* out) {
  NnWs w = carve_nn(ws, N);
  const int G = grid_size_for(N);
  const int cells = G * G * G;
  const int nblk = (cells + 1023) / 1024;
  bbox_kernel<<<1, 1024, 0, st>>>(s, N, G, w.gp);
  cudaError_t e = cudaMemsetAsync(w.cursor, 0, (size_t)cells * sizeof(int), st);  // used as the histogram first
  if (e != cudaSuccess) return (int)e;
  cell_count_kernel<<<d3d_ceil_div(N, 256), 256, 0, st>>>(s, N, w.gp, w.cell_of, w.cursor);
  scan_blocks_kernel<<<nblk, 1024, 0, st>>>(w.cursor, cells, w.cell_start, w.block_tot);
  scan_totals_kernel<<<1, 1024, 0, st>>>(w.block_tot, nblk);
  scan_add_kernel<<<nblk, 1024, 0, st>>>(w.cell_start, cells, w.block_tot, w.cursor);
  cell_fill_kernel<<<d3d_ceil_div(N, 256), 256, 0, st>>>(s, N, w.cell_of, w.cursor, w.sorted);
  d3d_note_launches(6);
  *out = w;
  return d3d_launch_status();
}

int nn_search(const float* q, const float* s, int M, int N, float* out_d2, int* out_idx, void* ws, cudaStream_t st,
              bool precise = false) {
  NnWs w;
  const int rc = build_grid(s, N, ws, st, &w);
  if (rc != 0) return rc;
  const int G = grid_size_for(N);
  if (precise)
    nn_query_kernel<true><<<d3d_ceil_div(M, 128), 128, 0, st>>>(q, M, w.gp, w.cell_start, G * G * G, N, w.sorted, out_d2, out_idx);
  else
    nn_query_kernel<false><<<d3d_ceil_div(M, 128), 128, 0, st>>>(q, M, w.gp, w.cell_start, G * G * G, N, w.sorted, out_d2, out_idx);
  d3d_note_launches(1);
  return d3d_launch_status();
}

constexpr int kSumBlocks = 512;

}  // namespace

// shared with patches.cu (radius patches walk the same grid)
int d3d_internal_build_grid(const float* s, int N, int G_override, void* ws, cudaStream_t st, void** gp_dev, int** cell_start,
                            float4** sorted, int* G_out) {
  (void)G_override;
  NnWs w;
  const int rc = build_grid(s, N, ws, st, &w);
  *gp_dev = w.gp; *cell_start = w.cell_start; *sorted = w.sorted; *G_out = grid_size_for(N);
  return rc;
}
size_t d3d_internal_grid_bytes(int N, int G_override) {
  (void)G_override;
  return carve_nn(nullptr, N).bytes;
}

extern "C" {

size_t d3d_nn_workspace_bytes(int N) {
  if (N <= 0) return 0;
  return carve_nn(nullptr, N).bytes;
}

int d3d_nn_sqdist(const float* query_xyz, const float* support_xyz, int M, int N, int precise, float* out_d2, int* out_idx,
                  void* ws, size_t ws_bytes, void* stream) {
  D3D_REQUIRE(query_xyz && support_xyz && out_d2);
  D3D_REQUIRE(M >= 0 && N > 0);
  if (M == 0) return 0;
  if (!ws || ws_bytes < d3d_nn_workspace_bytes(N)) return D3D_ERR_WORKSPACE;
  return nn_search(query_xyz, support_xyz, M, N, out_d2, out_idx, ws, (cudaStream_t)stream, precise != 0);
}

size_t d3d_chamfer_workspace_bytes(int Nx, int Ny) {
  if (Nx <= 0 || Ny <= 0) return 0;
  const size_t nn = d3d_nn_workspace_bytes(Nx > Ny ? Nx : Ny);
  return nn + align256((size_t)Nx * sizeof(float)) + align256((size_t)Ny * sizeof(float)) + 2 * align256(kSumBlocks * sizeof(double));
}

int d3d_chamfer_l2(const float* x, const float* y, int Nx, int Ny, float* out3, void* ws, size_t ws_bytes, void* stream) {
  D3D_REQUIRE(x && y && out3 && Nx > 0 && Ny > 0);
  if (!ws || ws_bytes < d3d_chamfer_workspace_bytes(Nx, Ny)) return D3D_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* p = (unsigned char*)ws;
  void* nn_ws = p; p += d3d_nn_workspace_bytes(Nx > Ny ? Nx : Ny);
  float* dx = (float*)p; p += align256((size_t)Nx * sizeof(float));
  float* dy = (float*)p; p += align256((size_t)Ny * sizeof(float));
  double* px = (double*)p; p += align256(kSumBlocks * sizeof(double));
  double* py = (double*)p;
  int rc = nn_search(x, y, Nx, Ny, dx, nullptr, nn_ws, st);  // every x to its nearest y
  if (rc != 0) return rc;
  rc = nn_search(y, x, Ny, Nx, dy, nullptr, nn_ws, st);      // every y to its nearest x
  if (rc != 0) return rc;
  sum_partials_kernel<<<kSumBlocks, 256, 0, st>>>(dx, Nx, px);
  sum_partials_kernel<<<kSumBlocks, 256, 0, st>>>(dy, Ny, py);
  chamfer_finish_kernel<<<1, 1, 0, st>>>(px, kSumBlocks, Nx, py, kSumBlocks, Ny, out3);
  d3d_note_launches(3);
  return d3d_launch_status();
}

}  // extern "C"

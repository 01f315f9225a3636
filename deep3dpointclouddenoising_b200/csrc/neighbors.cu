// Masked ordered ball query and masked nearest query for sm_100a.
//
// Semantics (bit-exact with the reference):
//   ref: u_net_arch/pt_custom_ops/_ext_src/src/masked_ordered_ball_query_gpu.cu:37-94
//   ref: u_net_arch/pt_custom_ops/_ext_src/src/masked_nearest_query_gpu.cu:30-61
//
// Design (not the reference's one-thread-per-query global-scratch scan):
//   * one WARP owns QW queries; the 32 lanes test 32 consecutive supports per step, so a
//     __ballot_sync + prefix popcount gives every in-radius support its slot in ASCENDING INDEX
//     order — exactly the reference's append order — without any atomics;
//   * supports are staged tile by tile in shared memory in SoA form (conflict-free, shared by the
//     16 warps = 32 queries of the block), the next tile travelling through registers meanwhile;
//   * the reference's "nearest in-radius support over ALL supports" (it replaces the last candidate
//     when the list overflowed) comes from an exact uniform-grid search per cloud (ball_grid_build +
//     ball_nearest) — or, for support sets of at most kScanMinN points, from the fill scan itself;
//   * candidates live in shared memory as 64-bit keys (d2 bits << 32 | index): d2 >= 0 so the IEEE
//     bit pattern is monotonic and one integer compare is the reference's stable sort-by-distance
//     order (ties -> lower index first);
//   * the <= 3*nsample candidates are ordered by warp-cooperative rank counting and written as
//     coalesced rows.  No (B, M, 3*nsample) global scratch, no zero-fill.
#include <stdlib.h>

#include "common.cuh"

namespace {

// tuning knobs (overridable through the environment for experiments: D3D_BQ_WARPS in {4,8,16}, D3D_BQ_TILE)
constexpr int kBqDefaultWarps = 16;  // measured on B200 (tools/bq_sweep.py): 16 warps x 2048-support tiles is the fastest of the sweep
constexpr int kBqDefaultTile = 2048;  // supports staged per shared-memory tile
constexpr int kScanMinN = 2048;       // at most this many supports: the fill kernel scans them all and finds the nearest itself

__global__ void prefix_len_kernel(const int* __restrict__ mask, int N, int* __restrict__ vlen) {
  __shared__ int first_zero;
  const int b = blockIdx.x;
  if (threadIdx.x == 0) first_zero = N;
  __syncthreads();
  const int* row = mask + (size_t)b * N;
  int local = N;
  for (int i = threadIdx.x; i < N; i += blockDim.x)
    if (row[i] == 0) { local = i; break; }  // per-thread indices ascend, first hit is its minimum
  if (local < N) atomicMin(&first_zero, local);
  __syncthreads();
  if (threadIdx.x == 0) vlen[b] = first_zero;
}

__device__ __forceinline__ unsigned long long make_key(float d2, int k) {
  return ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)k;
}

// ---- pass 1 of the ball query: the reference's running (min_dist, min_idx) over ALL in-radius supports (:59-62) ----
// = the nearest in-radius support, lowest index among equal distances.  Found exactly with a small uniform grid per
// cloud (built by one block per cloud in shared memory, cells sized ~2 supports) and expanding cube shells of
// cells: ~100 pair tests per query instead of N.  Distances use the reference's float expression (d3d_dist2).
constexpr int kNgMaxG = 16;  // up to 16^3 cells per cloud

struct CloudGrid {
  float min_x, min_y, min_z, cell, inv_cell;
  int G;
};

__device__ __forceinline__ int ng_coord(float v, float mn, float inv_cell, int G) {
  return min(max((int)floorf((v - mn) * inv_cell), 0), G - 1);
}

__global__ void __launch_bounds__(1024)
ball_grid_build_kernel(const float* __restrict__ support_xyz, const int* __restrict__ vlen, int N,
                       CloudGrid* __restrict__ grids, int* __restrict__ cell_start /* (B, 16^3 + 1) */,
                       float4* __restrict__ sorted /* (B, N) */) {
  __shared__ int counts[kNgMaxG * kNgMaxG * kNgMaxG];
  __shared__ float red[6][32];
  __shared__ int warp_tot[32];
  __shared__ CloudGrid g;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int v = vlen[b];
  const float* S = support_xyz + (size_t)b * N * 3;
  int* cs = cell_start + (size_t)b * (kNgMaxG * kNgMaxG * kNgMaxG + 1);
  float4* out = sorted + (size_t)b * N;
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int i = tid; i < v; i += 1024)
    for (int d = 0; d < 3; ++d) {
      const float x = S[3 * (size_t)i + d];
      lo[d] = fminf(lo[d], x);
      hi[d] = fmaxf(hi[d], x);
    }
  for (int d = 0; d < 3; ++d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[d] = fminf(lo[d], __shfl_xor_sync(D3D_FULL_MASK, lo[d], o));
      hi[d] = fmaxf(hi[d], __shfl_xor_sync(D3D_FULL_MASK, hi[d], o));
    }
    if (lane == 0) { red[d][warp] = lo[d]; red[3 + d][warp] = hi[d]; }
  }
  for (int c = tid; c < kNgMaxG * kNgMaxG * kNgMaxG; c += 1024) counts[c] = 0;
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
      int G = (int)cbrtf((float)max(v, 1) * 0.5f);
      G = min(max(G, 1), kNgMaxG);
      if (!(ext > 0.f) || !(ext < INFINITY)) { ext = 1.f; G = 1; }
      if (v == 0) { mn[0] = mn[1] = mn[2] = 0.f; }
      g.min_x = mn[0]; g.min_y = mn[1]; g.min_z = mn[2];
      g.cell = ext * (1.0f + 1e-5f) / (float)G;
      g.inv_cell = 1.0f / g.cell;
      g.G = G;
      grids[b] = g;
    }
  }
  __syncthreads();
  const int G = g.G, n_cells = G * G * G;
  for (int i = tid; i < v; i += 1024) {
    const int c = (ng_coord(S[3 * (size_t)i + 2], g.min_z, g.inv_cell, G) * G + ng_coord(S[3 * (size_t)i + 1], g.min_y, g.inv_cell, G)) * G +
                  ng_coord(S[3 * (size_t)i], g.min_x, g.inv_cell, G);
    atomicAdd(&counts[c], 1);
  }
  __syncthreads();
  // exclusive scan of counts[0 .. 4096): 4 cells per thread
  int loc[4], sum = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) { loc[k] = counts[tid * 4 + k]; sum += loc[k]; }
  int incl = sum;
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
  int run = warp_tot[warp] + incl - sum;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = tid * 4 + k;
    cs[c] = run;
    counts[c] = run;  // becomes the fill cursor
    run += loc[k];
  }
  if (tid == 1023) cs[kNgMaxG * kNgMaxG * kNgMaxG] = run;
  __syncthreads();
  (void)n_cells;
  for (int i = tid; i < v; i += 1024) {
    const float x = S[3 * (size_t)i], y = S[3 * (size_t)i + 1], z = S[3 * (size_t)i + 2];
    const int c = (ng_coord(z, g.min_z, g.inv_cell, G) * G + ng_coord(y, g.min_y, g.inv_cell, G)) * G + ng_coord(x, g.min_x, g.inv_cell, G);
    out[atomicAdd(&counts[c], 1)] = make_float4(x, y, z, __int_as_float(i));
  }
}

__global__ void __launch_bounds__(128)
ball_nearest_kernel(const float* __restrict__ query_xyz, const CloudGrid* __restrict__ grids,
                    const int* __restrict__ cell_start, const float4* __restrict__ sorted, int M, int N, float radius,
                    float* __restrict__ best_d2, int* __restrict__ best_k) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * 128 + threadIdx.x;
  if (j >= M) return;
  const CloudGrid g = grids[b];
  const int G = g.G;
  const int* cs = cell_start + (size_t)b * (kNgMaxG * kNgMaxG * kNgMaxG + 1);
  const float4* pts = sorted + (size_t)b * N;
  const float* q = query_xyz + ((size_t)b * M + j) * 3;
  const float qx = q[0], qy = q[1], qz = q[2];
  const float r2 = __fmul_rn(radius, radius);
  float best = r2;  // :45-47 min_dist = radius2, min_idx = 0
  int bk = 0;
  const int cx = ng_coord(qx, g.min_x, g.inv_cell, G), cy = ng_coord(qy, g.min_y, g.inv_cell, G), cz = ng_coord(qz, g.min_z, g.inv_cell, G);
  auto visit = [&](int x, int y, int z) {
    const int c = (z * G + y) * G + x;
    const int beg = cs[c], end = cs[c + 1];
    for (int t = beg; t < end; ++t) {
      const float4 s = __ldg(pts + t);
      const float d2 = d3d_dist2(qx, qy, qz, s.x, s.y, s.z);
      const int k = __float_as_int(s.w);
      // running minimum over ascending index == (smaller distance) or (equal distance and lower index)
      if (d2 < best || (d2 == best && best < r2 && k < bk)) { best = d2; bk = k; }
    }
  };
  for (int r = 0; r < G; ++r) {
    // The query's projection onto the grid box lies in the centre cell and is never farther from a support than the
    // query: every support of shell r is farther than (r - 1) * cell.  Compare in fp64 so that rounding cannot cut
    // the search short; best starts at radius^2, so shells beyond the radius are never visited.
    if (r >= 1) {
      const double reach = (double)(r - 1) * (double)g.cell * (1.0 - 1e-6);
      if ((double)best <= reach * reach) break;
    }
    if (cx - r < 0 && cx + r >= G && cy - r < 0 && cy + r >= G && cz - r < 0 && cz + r >= G) break;
    for (int z = max(cz - r, 0); z <= min(cz + r, G - 1); ++z) {
      const bool z_face = (z == cz - r) || (z == cz + r);
      for (int y = max(cy - r, 0); y <= min(cy + r, G - 1); ++y) {
        if (z_face || y == cy - r || y == cy + r) {
          for (int x = max(cx - r, 0); x <= min(cx + r, G - 1); ++x) visit(x, y, z);
        } else {
          if (cx - r >= 0) visit(cx - r, y, z);
          if (cx + r < G) visit(cx + r, y, z);
        }
      }
    }
  }
  best_d2[(size_t)b * M + j] = best;
  best_k[(size_t)b * M + j] = bk;
}

// far-away filler for the unused tail of a tile: its squared distance overflows to +inf, which fails
// every "d2 < something" test, so the scan loops need no bounds predicate
constexpr float kFar = 1.0e30f;

// ---- pass 2: candidate fill + ordered selection ----
//   fill : 32 supports per step, ballot + prefix popcount append the in-radius ones to the candidate list in
//          ascending index order; stops as soon as the list holds 3*nsample entries (on the BASELINE patches after
//          ~1/4 of the supports on average); a block leaves the tile loop once all of its lists are full;
//   then : nearest-swap with pass 1's result, histogram pre-selection, rank sort, coalesced row emission.
// kScanMin (small support sets, N <= kScanMinN): the kernel scans every support itself and keeps the running nearest
// in-radius support — one launch instead of grid build + grid search + this kernel, which are pure latency there.
template <int QW, int kBqWarps, bool kScanMin>
__global__ void __launch_bounds__(kBqWarps * 32)
ball_query_kernel(const float* __restrict__ query_xyz, const float* __restrict__ support_xyz,
                  const int* __restrict__ query_mask, const int* __restrict__ vlen,
                  const float* __restrict__ best_d2, const int* __restrict__ best_k, int M, int N, int tile,
                  float radius, int nsample, int* __restrict__ idx, int* __restrict__ idx_mask,
                  int* __restrict__ nvalid, int* __restrict__ idx_by_support) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int cap = 3 * nsample;
  float* sx = reinterpret_cast<float*>(smem_raw);
  float* sy = sx + tile;
  float* sz = sy + tile;
  unsigned long long* lists = reinterpret_cast<unsigned long long*>(sz + tile);
  int* sorted_all = reinterpret_cast<int*>(lists + (size_t)kBqWarps * QW * cap);
  int* hist_all = sorted_all + (size_t)kBqWarps * QW * nsample;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int q0 = (blockIdx.x * kBqWarps + warp) * QW;
  const int v = vlen[b];
  const float r2 = __fmul_rn(radius, radius);
  const unsigned lt_mask = (1u << lane) - 1u;

  const float* Q = query_xyz + (size_t)b * M * 3;
  const float* S = support_xyz + (size_t)b * N * 3;
  unsigned long long* my_list = lists + (size_t)warp * QW * cap;
  int* my_sorted = sorted_all + (size_t)warp * QW * nsample;
  int* my_hist = hist_all + warp * 32;

  float qx[QW], qy[QW], qz[QW];
  int cnt[QW];
  float near_d2[QW];  // kScanMin: this lane's running nearest in-radius support (:45-47, :59-62)
  int near_k[QW];
#pragma unroll
  for (int u = 0; u < QW; ++u) {
    const int j = min(q0 + u, M - 1);
    qx[u] = Q[3 * j + 0]; qy[u] = Q[3 * j + 1]; qz[u] = Q[3 * j + 2];
    cnt[u] = 0;
    near_d2[u] = r2;
    near_k[u] = 0;
  }

  // Tiles of supports go global -> registers -> shared memory; the NEXT tile's loads are issued before the current
  // tile is scanned, so their L2 latency (39 % of the stall samples of the single-buffered version, ncu) overlaps with
  // the scan.  kStage points per thread: 4 with 16 warps and 2048-point tiles.
  constexpr int kStage = (kBqDefaultTile / (kBqWarps * 32) <= 8) ? kBqDefaultTile / (kBqWarps * 32) : 0;
  const bool staged = kStage > 0 && tile <= kStage * (int)blockDim.x;
  float rx[kStage > 0 ? kStage : 1], ry[kStage > 0 ? kStage : 1], rz[kStage > 0 ? kStage : 1];
  auto fetch = [&](int base) {  // tile starting at `base` -> registers (kFar beyond the cloud)
#pragma unroll
    for (int k = 0; k < kStage; ++k) {
      const int i = threadIdx.x + k * blockDim.x;
      float x = kFar, y = kFar, z = kFar;
      if (i < tile && base + i < v) {
        const float* p = S + (size_t)(base + i) * 3;
        x = p[0]; y = p[1]; z = p[2];
      }
      rx[k] = x; ry[k] = y; rz[k] = z;
    }
  };
  if (staged && v > 0) fetch(0);

  for (int base = 0; base < v; base += tile) {
    bool filling = false;
#pragma unroll
    for (int u = 0; u < QW; ++u) filling = filling || ((kScanMin || cnt[u] < cap) && q0 + u < M);
    if (!__syncthreads_or(filling ? 1 : 0)) break;  // every list of the block is full (also: previous tile consumed)
    const int tile_n = min(tile, v - base);
    if (staged) {
#pragma unroll
      for (int k = 0; k < kStage; ++k) {
        const int i = threadIdx.x + k * blockDim.x;
        if (i < tile) { sx[i] = rx[k]; sy[i] = ry[k]; sz[i] = rz[k]; }
      }
      __syncthreads();
      if (base + tile < v) fetch(base + tile);  // in flight during the scan below
    } else {
      for (int i = threadIdx.x; i < tile; i += blockDim.x) {
        float x = kFar, y = kFar, z = kFar;
        if (i < tile_n) {
          const float* p = S + (size_t)(base + i) * 3;
          x = p[0]; y = p[1]; z = p[2];
        }
        sx[i] = x; sy[i] = y; sz[i] = z;
      }
      __syncthreads();
    }
    const int steps = (tile_n + 31) >> 5;
    // ---- fill pass, one query at a time, until its list is full
#pragma unroll
    for (int u = 0; u < QW; ++u) {
      for (int st = 0; st < steps && (kScanMin || cnt[u] < cap); ++st) {
        const int i = (st << 5) + lane;
        const float d2 = d3d_dist2(qx[u], qy[u], qz[u], sx[i], sy[i], sz[i]);
        const bool inr = d2 < r2;
        if (kScanMin && d2 < near_d2[u]) { near_d2[u] = d2; near_k[u] = base + i; }  // a lane's indices ascend
        const unsigned ball = __ballot_sync(D3D_FULL_MASK, inr);
        if (ball) {
          const int pos = cnt[u] + __popc(ball & lt_mask);
          if (inr && pos < cap) my_list[u * cap + pos] = make_key(d2, base + i);
          cnt[u] += __popc(ball);
        }
      }
    }
  }

#pragma unroll
  for (int u = 0; u < QW; ++u) {
    const int j = q0 + u;
    if (j >= M) break;  // warp-uniform
    unsigned long long* list = my_list + u * cap;
    int* sorted = my_sorted + u * nsample;
    const int c = min(cnt[u], cap);
    __syncwarp();
    if (cnt[u] >= cap) {
      // :72-75  the global nearest in-radius support (pass 1) lies beyond the last slot -> it replaces the last slot
      float bd;
      int bk;
      if (kScanMin) {  // warp minimum of (d2, index): smaller distance, then lower index — the scan order's winner
        unsigned long long key = make_key(near_d2[u], near_k[u]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const unsigned long long other = __shfl_xor_sync(D3D_FULL_MASK, key, o);
          key = other < key ? other : key;
        }
        bd = __uint_as_float((unsigned)(key >> 32));
        bk = (int)(unsigned)(key & 0xffffffffull);
      } else {
        bd = best_d2[(size_t)b * M + j];
        bk = best_k[(size_t)b * M + j];
      }
      if (lane == 0 && bk > (int)(unsigned)(list[cap - 1] & 0xffffffffull)) list[cap - 1] = make_key(bd, bk);
      __syncwarp();
    }
    // Pre-selection: only the nsample smallest keys are emitted.  d2 -> bin is monotone, so every key in a
    // bin below B* is smaller than every key above it: rank only the keys of bins <= B*.
    int n_rank = c;
    if (c > nsample) {
      my_hist[lane] = 0;
      __syncwarp();
      const float scale = 32.0f / r2;
      for (int t = lane; t < c; t += 32) {
        const int bin = min(31, (int)(__uint_as_float((unsigned)(list[t] >> 32)) * scale));
        atomicAdd(&my_hist[bin], 1);
      }
      __syncwarp();
      int incl = my_hist[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(D3D_FULL_MASK, incl, o);
        if (lane >= o) incl += up;
      }
      const unsigned reach = __ballot_sync(D3D_FULL_MASK, incl >= nsample);
      const int bstar = __ffs(reach) - 1;  // exists: incl[31] == c > nsample
      n_rank = __shfl_sync(D3D_FULL_MASK, incl, bstar);
      if (n_rank < c) {  // in-place left compaction, round by round (reads of a round precede its writes)
        int kept = 0;
        for (int t0 = 0; t0 < c; t0 += 32) {
          const int t = t0 + lane;
          unsigned long long key = 0;
          bool keep = false;
          if (t < c) {
            key = list[t];
            keep = min(31, (int)(__uint_as_float((unsigned)(key >> 32)) * scale)) <= bstar;
          }
          const unsigned kb = __ballot_sync(D3D_FULL_MASK, keep);
          __syncwarp();
          if (keep) list[kept + __popc(kb & lt_mask)] = key;
          kept += __popc(kb);
          __syncwarp();
        }
      }
    }
    // rank counting among the n_rank selected keys (unique): rank = number of smaller keys
    // idx_by_support (optional): the same nsample-or-fewer winners in ASCENDING SUPPORT INDEX (the list is in that order
    // already), each as (distance rank << 16) | index, -1 padded — what the staged-tile aggregation kernels walk
    int* srow = idx_by_support != nullptr ? idx_by_support + ((size_t)b * M + j) * nsample : nullptr;
    int emitted = 0;
    for (int t0 = 0; t0 < n_rank; t0 += 64) {
      const int ta = t0 + lane, tb = t0 + 32 + lane;
      const unsigned long long own_a = ta < n_rank ? list[ta] : ~0ull;
      const unsigned long long own_b = tb < n_rank ? list[tb] : ~0ull;
      int ra = 0, rb = 0;
#pragma unroll 4
      for (int jj = 0; jj < n_rank; ++jj) {
        const unsigned long long kj = list[jj];  // broadcast
        ra += (kj < own_a) ? 1 : 0;
        rb += (kj < own_b) ? 1 : 0;
      }
      const bool ka = ta < n_rank && ra < nsample, kb = tb < n_rank && rb < nsample;
      if (ka) sorted[ra] = (int)(unsigned)(own_a & 0xffffffffull);
      if (kb) sorted[rb] = (int)(unsigned)(own_b & 0xffffffffull);
      if (srow != nullptr) {
        const unsigned ba = __ballot_sync(D3D_FULL_MASK, ka), bb = __ballot_sync(D3D_FULL_MASK, kb);
        if (ka) srow[emitted + __popc(ba & lt_mask)] = (ra << 16) | (int)(unsigned)(own_a & 0xffffu);
        if (kb) srow[emitted + __popc(ba) + __popc(bb & lt_mask)] = (rb << 16) | (int)(unsigned)(own_b & 0xffffu);
        emitted += __popc(ba) + __popc(bb);
      }
    }
    if (srow != nullptr)
      for (int i = emitted + lane; i < nsample; i += 32) srow[i] = -1;
    __syncwarp();
    const int qm = query_mask[(size_t)b * M + j];
    int* orow = idx + ((size_t)b * M + j) * nsample;
    int* mrow = idx_mask + ((size_t)b * M + j) * nsample;
    for (int i = lane; i < nsample; i += 32) {
      int o = 0, mk = 0;
      if (c > 0) {
        o = sorted[i < c ? i : i % c];  // :83-86 cyclic padding
        mk = (i < c && qm != 0) ? 1 : 0;   // :88-93 padded query -> whole row masked
      }
      orow[i] = o;
      mrow[i] = mk;
    }
    if (nvalid != nullptr && lane == 0) nvalid[(size_t)b * M + j] = min(c, nsample);
  }
}

constexpr int kNnThreads = 128;
constexpr int kNnTile = 2048;

__global__ void __launch_bounds__(kNnThreads)
nearest_query_kernel(const float* __restrict__ query_xyz, const float* __restrict__ support_xyz,
                     const int* __restrict__ query_mask, const int* __restrict__ vlen, int M, int N,
                     int* __restrict__ idx, int* __restrict__ idx_mask) {
  __shared__ float sx[kNnTile], sy[kNnTile], sz[kNnTile];
  const int b = blockIdx.y;
  const int j = blockIdx.x * kNnThreads + threadIdx.x;
  const int v = vlen[b];
  const float* Q = query_xyz + (size_t)b * M * 3;
  const float* S = support_xyz + (size_t)b * N * 3;
  const int jj = min(j, M - 1);
  const float qx = Q[3 * jj], qy = Q[3 * jj + 1], qz = Q[3 * jj + 2];
  float best = 100.0f;  // masked_nearest_query_gpu.cu:36
  int bestk = -1;
  for (int base = 0; base < v; base += kNnTile) {
    __syncthreads();
    const int tile_n = min(kNnTile, v - base);
    for (int f = threadIdx.x; f < 3 * tile_n; f += blockDim.x) {
      const float val = S[(size_t)base * 3 + f];
      const int i = f / 3, c = f - 3 * i;
      (c == 0 ? sx : (c == 1 ? sy : sz))[i] = val;
    }
    __syncthreads();
#pragma unroll 8
    for (int i = 0; i < tile_n; ++i) {  // every lane reads the same word: shared-memory broadcast
      const float d2 = d3d_dist2(qx, qy, qz, sx[i], sy[i], sz[i]);
      if (d2 < best) { best = d2; bestk = base + i; }  // strict <, ascending i: lowest index wins
    }
  }
  if (j < M) {
    idx[(size_t)b * M + j] = bestk;
    idx_mask[(size_t)b * M + j] = query_mask[(size_t)b * M + j] != 0 ? 1 : 0;
  }
}

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// experiment knobs, read ONCE per process (not per launch)
struct BqKnobs {
  int tile, warps, scanmin_n;
  BqKnobs() : tile(env_int("D3D_BQ_TILE", kBqDefaultTile)), warps(env_int("D3D_BQ_WARPS", kBqDefaultWarps)),
              scanmin_n(env_int("D3D_BQ_SCANMIN_N", kScanMinN)) {}
};
const BqKnobs& bq_knobs() {
  static const BqKnobs k;
  return k;
}

int bq_tile(int N) {
  const int t = (N + 31) & ~31;
  const int mx = (bq_knobs().tile + 31) & ~31;
  return t < mx ? t : mx;
}

size_t bq_smem_bytes(int warps, int qw, int nsample, int tile) {
  return (size_t)3 * tile * sizeof(float) + (size_t)warps * qw * 3 * nsample * sizeof(unsigned long long) +
         (size_t)warps * qw * nsample * sizeof(int) + (size_t)warps * 32 * sizeof(int);
}

template <int QW, int WARPS, bool SCAN>
int launch_ball_query(const float* q, const float* s, const int* qm, const int* vlen, const float* best_d2,
                      const int* best_k, int B, int M, int N, float radius, int nsample, int* idx, int* idx_mask,
                      int* nvalid, int* by_support, cudaStream_t st) {
  const int tile = bq_tile(N);
  const size_t smem = bq_smem_bytes(WARPS, QW, nsample, tile);
  cudaError_t e = cudaFuncSetAttribute(ball_query_kernel<QW, WARPS, SCAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  dim3 grid(d3d_ceil_div(M, WARPS * QW), B);
  ball_query_kernel<QW, WARPS, SCAN><<<grid, WARPS * 32, smem, st>>>(q, s, qm, vlen, best_d2, best_k, M, N, tile, radius, nsample,
                                                             idx, idx_mask, nvalid, by_support);
  d3d_note_launches(1);
  return d3d_launch_status();
}

template <int WARPS, bool SCAN>
int dispatch_ball_query(const float* q, const float* s, const int* qm, const int* vlen, const float* best_d2,
                        const int* best_k, int B, int M, int N, float radius, int nsample, int* idx, int* idx_mask,
                        int* nvalid, int* by_support, cudaStream_t st) {
  const size_t budget = 220 * 1024;  // opt-in shared memory per block on sm_100a is 227 KB
  const int tile = bq_tile(N);
  if (bq_smem_bytes(WARPS, 4, nsample, tile) <= budget / 2)
    return launch_ball_query<4, WARPS, SCAN>(q, s, qm, vlen, best_d2, best_k, B, M, N, radius, nsample, idx, idx_mask, nvalid, by_support, st);
  if (bq_smem_bytes(WARPS, 2, nsample, tile) <= budget / 2)
    return launch_ball_query<2, WARPS, SCAN>(q, s, qm, vlen, best_d2, best_k, B, M, N, radius, nsample, idx, idx_mask, nvalid, by_support, st);
  if (bq_smem_bytes(WARPS, 1, nsample, tile) <= budget)
    return launch_ball_query<1, WARPS, SCAN>(q, s, qm, vlen, best_d2, best_k, B, M, N, radius, nsample, idx, idx_mask, nvalid, by_support, st);
  return D3D_ERR_UNSUPPORTED;
}

}  // namespace

void d3d_launch_prefix_len(const int* mask, int B, int N, int* vlen, cudaStream_t st) {
  prefix_len_kernel<<<B, 256, 0, st>>>(mask, N, vlen);
  d3d_note_launches(1);
}

extern "C" {

static size_t bq_align(size_t x) { return (x + 255) & ~(size_t)255; }

/* [vlen: B int][grids: B][cell_start: B x (16^3 + 1)][sorted supports: B x N float4][best d2, best k: B x M each] */
size_t d3d_ball_query_workspace_bytes(int B, int M, int N) {
  if (B <= 0 || M < 0 || N <= 0) return 0;
  return bq_align((size_t)B * sizeof(int)) + bq_align((size_t)B * sizeof(CloudGrid)) +
         bq_align((size_t)B * (kNgMaxG * kNgMaxG * kNgMaxG + 1) * sizeof(int)) + bq_align((size_t)B * N * sizeof(float4)) +
         2 * bq_align((size_t)B * M * sizeof(float));
}

int d3d_ball_query(const float* query_xyz, const float* support_xyz, const int* query_mask,
                   const int* support_mask, int B, int M, int N, float radius, int nsample, int* idx,
                   int* idx_mask, int* nvalid, int* idx_by_support, void* ws, size_t ws_bytes, void* stream) {
  int* by_support = (N <= 65536 && nsample <= 255) ? idx_by_support : nullptr;  // 16-bit index, 8-bit rank
  if (idx_by_support != nullptr && by_support == nullptr) return D3D_ERR_UNSUPPORTED;
  D3D_REQUIRE(query_xyz && support_xyz && query_mask && support_mask && idx && idx_mask);
  D3D_REQUIRE(B >= 0 && M >= 0 && N > 0 && nsample > 0 && nsample <= D3D_MAX_NSAMPLE);
  if (B == 0 || M == 0) return 0;
  if (!ws || ws_bytes < d3d_ball_query_workspace_bytes(B, M, N)) return D3D_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* p = (unsigned char*)ws;
  int* vlen = (int*)p; p += bq_align((size_t)B * sizeof(int));
  CloudGrid* grids = (CloudGrid*)p; p += bq_align((size_t)B * sizeof(CloudGrid));
  int* cell_start = (int*)p; p += bq_align((size_t)B * (kNgMaxG * kNgMaxG * kNgMaxG + 1) * sizeof(int));
  float4* sorted = (float4*)p; p += bq_align((size_t)B * N * sizeof(float4));
  float* best_d2 = (float*)p; p += bq_align((size_t)B * M * sizeof(float));
  int* best_k = (int*)p;
  d3d_launch_prefix_len(support_mask, B, N, vlen, st);
  const int warps = bq_knobs().warps;
  if (N <= bq_knobs().scanmin_n) {  // small support set: one launch, the kernel finds the nearest itself
    switch (warps) {
      case 4: return dispatch_ball_query<4, true>(query_xyz, support_xyz, query_mask, vlen, best_d2, best_k, B, M, N, radius, nsample, idx, idx_mask, nvalid, by_support, st);
      case 16: return dispatch_ball_query<16, true>(query_xyz, support_xyz, query_mask, vlen, best_d2, best_k, B, M, N, radius, nsample, idx, idx_mask, nvalid, by_support, st);
      default: return dispatch_ball_query<8, true>(query_xyz, support_xyz, query_mask, vlen, best_d2, best_k, B, M, N, radius, nsample, idx, idx_mask, nvalid, by_support, st);
    }
  }
  ball_grid_build_kernel<<<B, 1024, 0, st>>>(support_xyz, vlen, N, grids, cell_start, sorted);
  ball_nearest_kernel<<<dim3(d3d_ceil_div(M, 128), B), 128, 0, st>>>(query_xyz, grids, cell_start, sorted, M, N, radius, best_d2,
                                                                   best_k);
  d3d_note_launches(2);
  switch (warps) {
    case 4: return dispatch_ball_query<4, false>(query_xyz, support_xyz, query_mask, vlen, best_d2, best_k, B, M, N, radius, nsample, idx, idx_mask, nvalid, by_support, st);
    case 16: return dispatch_ball_query<16, false>(query_xyz, support_xyz, query_mask, vlen, best_d2, best_k, B, M, N, radius, nsample, idx, idx_mask, nvalid, by_support, st);
    default: return dispatch_ball_query<8, false>(query_xyz, support_xyz, query_mask, vlen, best_d2, best_k, B, M, N, radius, nsample, idx, idx_mask, nvalid, by_support, st);
  }
}

size_t d3d_nearest_query_workspace_bytes(int B) { return (size_t)(B > 0 ? B : 0) * sizeof(int); }

int d3d_nearest_query(const float* query_xyz, const float* support_xyz, const int* query_mask,
                      const int* support_mask, int B, int M, int N, int* idx, int* idx_mask, void* ws,
                      size_t ws_bytes, void* stream) {
  D3D_REQUIRE(query_xyz && support_xyz && query_mask && support_mask && idx && idx_mask);
  D3D_REQUIRE(B >= 0 && M >= 0 && N > 0);
  if (B == 0 || M == 0) return 0;
  if (!ws || ws_bytes < d3d_nearest_query_workspace_bytes(B)) return D3D_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  int* vlen = (int*)ws;
  d3d_launch_prefix_len(support_mask, B, N, vlen, st);
  dim3 grid(d3d_ceil_div(M, kNnThreads), B);
  nearest_query_kernel<<<grid, kNnThreads, 0, st>>>(query_xyz, support_xyz, query_mask, vlen, M, N, idx, idx_mask);
  d3d_note_launches(1);
  return d3d_launch_status();
}

}  // extern "C"

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
//     8 warps = 32 queries of the block);
//   * candidates live in shared memory as 64-bit keys (d2 bits << 32 | index): d2 >= 0 so the IEEE
//     bit pattern is monotonic and one integer compare is the reference's stable sort-by-distance
//     order (ties -> lower index first);
//   * the <= 3*nsample candidates are ordered by warp-cooperative rank counting and written as
//     coalesced rows.  No (B, M, 3*nsample) global scratch, no zero-fill.
#include "common.cuh"

namespace {

constexpr int kBqWarps = 8;
constexpr int kBqTile = 1024;  // supports per shared-memory tile

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

template <int QW>
__global__ void __launch_bounds__(kBqWarps * 32)
ball_query_kernel(const float* __restrict__ query_xyz, const float* __restrict__ support_xyz,
                  const int* __restrict__ query_mask, const int* __restrict__ vlen, int M, int N,
                  float radius, int nsample, int* __restrict__ idx, int* __restrict__ idx_mask,
                  int* __restrict__ nvalid) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int cap = 3 * nsample;
  float* sx = reinterpret_cast<float*>(smem_raw);
  float* sy = sx + kBqTile;
  float* sz = sy + kBqTile;
  unsigned long long* lists = reinterpret_cast<unsigned long long*>(sz + kBqTile);
  int* sorted_all = reinterpret_cast<int*>(lists + (size_t)kBqWarps * QW * cap);

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

  float qx[QW], qy[QW], qz[QW], best[QW];
  int bestk[QW], cnt[QW];
#pragma unroll
  for (int u = 0; u < QW; ++u) {
    const int j = min(q0 + u, M - 1);
    qx[u] = Q[3 * j + 0]; qy[u] = Q[3 * j + 1]; qz[u] = Q[3 * j + 2];
    best[u] = r2; bestk[u] = 0; cnt[u] = 0;
  }

  for (int base = 0; base < v; base += kBqTile) {
    __syncthreads();  // everyone is done with the previous tile
    const int tile_n = min(kBqTile, v - base);
    // coalesced flat copy of 3*tile_n floats, de-interleaved into SoA
    for (int f = threadIdx.x; f < 3 * tile_n; f += blockDim.x) {
      const float val = S[(size_t)base * 3 + f];
      const int i = f / 3, c = f - 3 * i;
      (c == 0 ? sx : (c == 1 ? sy : sz))[i] = val;
    }
    __syncthreads();
    for (int i0 = 0; i0 < tile_n; i0 += 32) {
      const int i = i0 + lane;
      const bool in_tile = i < tile_n;
      const int k = base + i;
      const float x = sx[i], y = sy[i], z = sz[i];  // i < kBqTile always (tile is a multiple of 32)
#pragma unroll
      for (int u = 0; u < QW; ++u) {
        const float d2 = d3d_dist2(qx[u], qy[u], qz[u], x, y, z);
        const bool inr = in_tile && (d2 < r2);
        if (inr && d2 < best[u]) { best[u] = d2; bestk[u] = k; }  // strict: earliest index per lane
        const unsigned ball = __ballot_sync(D3D_FULL_MASK, inr);
        if (ball) {
          if (cnt[u] < cap) {
            const int pos = cnt[u] + __popc(ball & lt_mask);
            if (inr && pos < cap) my_list[u * cap + pos] = make_key(d2, k);
          }
          cnt[u] += __popc(ball);
        }
      }
    }
  }

#pragma unroll
  for (int u = 0; u < QW; ++u) {
    const int j = q0 + u;
    if (j >= M) break;  // warp-uniform
    // global nearest in-radius support: min d2, lowest index among equal d2 (:59-62)
    float bd = best[u];
    int bk = bestk[u];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float od = __shfl_xor_sync(D3D_FULL_MASK, bd, off);
      const int ok = __shfl_xor_sync(D3D_FULL_MASK, bk, off);
      if (od < bd || (od == bd && ok < bk)) { bd = od; bk = ok; }
    }
    unsigned long long* list = my_list + u * cap;
    int* sorted = my_sorted + u * nsample;
    const int c = min(cnt[u], cap);
    __syncwarp();
    if (lane == 0 && cnt[u] >= cap && cap > 0) {  // :72-75 nearest-swap into the last slot
      const int last_k = (int)(unsigned)(list[cap - 1] & 0xffffffffull);
      if (bk > last_k) list[cap - 1] = make_key(bd, bk);
    }
    __syncwarp();
    // rank counting: keys are unique, rank = number of smaller keys = position after the stable sort
    for (int t0 = 0; t0 < c; t0 += 128) {
      unsigned long long own[4];
      int rank[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int t = t0 + e * 32 + lane;
        own[e] = t < c ? list[t] : ~0ull;
        rank[e] = 0;
      }
#pragma unroll 4
      for (int jj = 0; jj < c; ++jj) {
        const unsigned long long kj = list[jj];  // broadcast
#pragma unroll
        for (int e = 0; e < 4; ++e) rank[e] += (kj < own[e]) ? 1 : 0;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int t = t0 + e * 32 + lane;
        if (t < c && rank[e] < nsample) sorted[rank[e]] = (int)(unsigned)(own[e] & 0xffffffffull);
      }
    }
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

template <int QW>
int launch_ball_query(const float* q, const float* s, const int* qm, const int* vlen, int B, int M, int N,
                      float radius, int nsample, int* idx, int* idx_mask, int* nvalid, size_t smem,
                      cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(ball_query_kernel<QW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  dim3 grid(d3d_ceil_div(M, kBqWarps * QW), B);
  ball_query_kernel<QW><<<grid, kBqWarps * 32, smem, st>>>(q, s, qm, vlen, M, N, radius, nsample, idx, idx_mask,
                                                          nvalid);
  d3d_note_launches(1);
  return d3d_launch_status();
}

size_t bq_smem_bytes(int qw, int nsample) {
  return (size_t)3 * kBqTile * sizeof(float) + (size_t)kBqWarps * qw * 3 * nsample * sizeof(unsigned long long) +
         (size_t)kBqWarps * qw * nsample * sizeof(int);
}

}  // namespace

void d3d_launch_prefix_len(const int* mask, int B, int N, int* vlen, cudaStream_t st) {
  prefix_len_kernel<<<B, 256, 0, st>>>(mask, N, vlen);
  d3d_note_launches(1);
}

extern "C" {

size_t d3d_ball_query_workspace_bytes(int B) { return (size_t)(B > 0 ? B : 0) * sizeof(int); }

int d3d_ball_query(const float* query_xyz, const float* support_xyz, const int* query_mask,
                   const int* support_mask, int B, int M, int N, float radius, int nsample, int* idx,
                   int* idx_mask, int* nvalid, void* ws, size_t ws_bytes, void* stream) {
  D3D_REQUIRE(query_xyz && support_xyz && query_mask && support_mask && idx && idx_mask);
  D3D_REQUIRE(B >= 0 && M >= 0 && N > 0 && nsample > 0 && nsample <= D3D_MAX_NSAMPLE);
  if (B == 0 || M == 0) return 0;
  if (!ws || ws_bytes < d3d_ball_query_workspace_bytes(B)) return D3D_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  int* vlen = (int*)ws;
  d3d_launch_prefix_len(support_mask, B, N, vlen, st);
  const size_t budget = 100 * 1024;  // keep >= 2 blocks per SM
  if (bq_smem_bytes(4, nsample) <= budget)
    return launch_ball_query<4>(query_xyz, support_xyz, query_mask, vlen, B, M, N, radius, nsample, idx, idx_mask,
                                nvalid, bq_smem_bytes(4, nsample), st);
  if (bq_smem_bytes(2, nsample) <= budget)
    return launch_ball_query<2>(query_xyz, support_xyz, query_mask, vlen, B, M, N, radius, nsample, idx, idx_mask,
                                nvalid, bq_smem_bytes(2, nsample), st);
  if (bq_smem_bytes(1, nsample) <= 200 * 1024)
    return launch_ball_query<1>(query_xyz, support_xyz, query_mask, vlen, B, M, N, radius, nsample, idx, idx_mask,
                                nvalid, bq_smem_bytes(1, nsample), st);
  return D3D_ERR_UNSUPPORTED;
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

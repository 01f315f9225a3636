// Inverse neighbour map: CSR by support point of an idx tensor (B, M, nsample).
//
// Every backward kernel of the path (gather grad, PosPool, PseudoGrid, max-pool, nearest upsample)
// is a sum over "all (query, slot) pairs that gathered support i".  The reference scatters with
// float atomicAdd (ref: u_net_arch/pt_custom_ops/_ext_src/src/group_points_gpu.cu:58-68) and is
// therefore non-deterministic; here the pairs of each support are listed once, in ascending
// (query, slot) order, and every backward kernel reduces its segment sequentially: no float atomics,
// bit-reproducible gradients.
//
// Entries are (query << 8) | slot.  Two builders:
//   chunked (default, N <= kChunkMaxN): every cloud's entry list is cut into <= 16 chunks; a block
//     histograms its chunk in SHARED memory and writes one row of a (cloud, chunk, support) count matrix;
//     a scan turns the matrix into rowptr and per-(chunk, support) write cursors; the fill pass hands out
//     slots with shared-memory atomics.  No global atomics; a segment comes out as <= 16 pieces in chunk
//     order, each piece (a few entries) is then rank-sorted by a warp.
//   global (fallback for very large N): global integer histogram -> scan -> cursor fill -> per-segment
//     sort (a warp for short segments, a block-wide bitonic network for long ones).
#include "common.cuh"

namespace {

constexpr int kShortLen = 64;   // segments up to this length are sorted by one warp
constexpr int kSortSmem = 8192;  // longer ones by a block, in shared memory up to this many entries
constexpr int kSortThreads = 256;

__global__ void count_kernel(const int* __restrict__ idx, long long total, int P, int N, int* __restrict__ counts) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int b = (int)(e / P);
  atomicAdd(&counts[(size_t)b * N + d3d_clamp_index(idx[e], N)], 1);
}

// one block per cloud: rowptr[b*N + i] = b*P + exclusive prefix of counts[b, :]; cursor <- rowptr
__global__ void __launch_bounds__(1024)
scan_kernel(int* __restrict__ counts_then_cursor, int N, int P, int B, int* __restrict__ rowptr) {
  __shared__ int warp_tot[32];
  __shared__ int carry;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int* cnt = counts_then_cursor + (size_t)b * N;
  int* rp = rowptr + (size_t)b * N;
  if (tid == 0) carry = b * P;
  __syncthreads();
  for (int base = 0; base < N; base += 1024) {
    const int i = base + tid;
    const int c = i < N ? cnt[i] : 0;
    int incl = c;
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
    const int excl = carry + warp_tot[warp] + incl - c;
    if (i < N) { rp[i] = excl; cnt[i] = excl; }
    __syncthreads();
    if (tid == 1023) carry = excl + c;
    __syncthreads();
  }
  if (b == B - 1 && tid == 0) rowptr[(size_t)B * N] = B * P;
}

__global__ void fill_kernel(const int* __restrict__ idx, long long total, int P, int N, int nsample,
                            int* __restrict__ cursor, int* __restrict__ unsorted) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int b = (int)(e / P);
  const int p = (int)(e - (long long)b * P);
  const int j = p / nsample, k = p - j * nsample;
  const int pos = atomicAdd(&cursor[(size_t)b * N + d3d_clamp_index(idx[e], N)], 1);
  unsorted[pos] = (j << 8) | k;
}

// warp per support: segments <= kShortLen are rank-sorted here; longer ones are queued for the block sorter
__global__ void sort_short_kernel(const int* __restrict__ rowptr, int rows, const int* __restrict__ unsorted,
                                  int* __restrict__ entries, int* __restrict__ long_rows, int* __restrict__ long_count) {
  const int lane = threadIdx.x & 31;
  const int row = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (row >= rows) return;
  const int beg = rowptr[row], len = rowptr[row + 1] - beg;
  if (len > kShortLen) {
    if (lane == 0) long_rows[atomicAdd(long_count, 1)] = row;
    return;
  }
  // entries are unique -> rank = number of smaller entries
  int own0 = lane < len ? unsorted[beg + lane] : 0x7fffffff;
  int own1 = lane + 32 < len ? unsorted[beg + lane + 32] : 0x7fffffff;
  int r0 = 0, r1 = 0;
  for (int t = 0; t < len; ++t) {
    const int other = t < 32 ? __shfl_sync(D3D_FULL_MASK, own0, t) : __shfl_sync(D3D_FULL_MASK, own1, t - 32);
    r0 += other < own0 ? 1 : 0;
    r1 += other < own1 ? 1 : 0;
  }
  if (lane < len) entries[beg + r0] = own0;
  if (lane + 32 < len) entries[beg + r1] = own1;
}

__global__ void __launch_bounds__(kSortThreads)
sort_long_kernel(const int* __restrict__ rowptr, const int* __restrict__ long_rows, const int* __restrict__ long_count,
                 int* unsorted, int* __restrict__ entries) {
  __shared__ int buf[kSortSmem];
  const int n_long = *long_count;
  for (int it = blockIdx.x; it < n_long; it += gridDim.x) {
    const int row = long_rows[it];
    const int beg = rowptr[row], len = rowptr[row + 1] - beg;
    int P2 = 1;
    while (P2 < len) P2 <<= 1;
    if (P2 <= kSortSmem) {
      for (int i = threadIdx.x; i < P2; i += kSortThreads) buf[i] = i < len ? unsorted[beg + i] : 0x7fffffff;
      __syncthreads();
      for (int k = 2; k <= P2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
          for (int t = threadIdx.x; t < (P2 >> 1); t += kSortThreads) {
            const int a = ((t & ~(j - 1)) << 1) | (t & (j - 1)), c = a | j;
            const int va = buf[a], vc = buf[c];
            if ((va > vc) == ((a & k) == 0)) { buf[a] = vc; buf[c] = va; }
          }
          __syncthreads();
        }
      for (int i = threadIdx.x; i < len; i += kSortThreads) entries[beg + i] = buf[i];
      __syncthreads();
    } else {
      // very long segment: odd-even merge is overkill; rank counting straight from global memory
      // (L1/L2 resident), out of place.  O(len^2 / 256) per thread, only for > 8192 entries.
      for (int i = threadIdx.x; i < len; i += kSortThreads) {
        const int own = unsorted[beg + i];
        int r = 0;
        for (int t = 0; t < len; ++t) r += unsorted[beg + t] < own ? 1 : 0;
        entries[beg + r] = own;
      }
    }
  }
}


// ---------------------------------------------------------------------------------------------------
// chunked builder
// ---------------------------------------------------------------------------------------------------
constexpr int kChunkMaxN = 49152;  // N ints of shared memory per block (192 KB)
constexpr int kMaxChunks = 16;

// block = (chunk, cloud): histogram of idx[b, chunk range] over supports, written as one matrix row
__global__ void __launch_bounds__(512)
chunk_count_kernel(const int* __restrict__ idx, int P, int N, int n_chunks, int chunk_len, int* __restrict__ matrix) {
  extern __shared__ int hist[];
  const int chunk = blockIdx.x, b = blockIdx.y;
  for (int i = threadIdx.x; i < N; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  const int beg = chunk * chunk_len, end = min(beg + chunk_len, P);
  const int* src = idx + (size_t)b * P;
  for (int e = beg + threadIdx.x; e < end; e += blockDim.x) atomicAdd(&hist[d3d_clamp_index(src[e], N)], 1);
  __syncthreads();
  int* row = matrix + ((size_t)b * n_chunks + chunk) * N;
  for (int i = threadIdx.x; i < N; i += blockDim.x) row[i] = hist[i];
}

// block per cloud: counts -> rowptr, and the matrix is rewritten in place as write cursors
__global__ void __launch_bounds__(1024)
chunk_scan_kernel(int* __restrict__ matrix, int N, int P, int B, int n_chunks, int* __restrict__ rowptr) {
  __shared__ int warp_tot[32];
  __shared__ int carry;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int* mat = matrix + (size_t)b * n_chunks * N;
  int* rp = rowptr + (size_t)b * N;
  if (tid == 0) carry = b * P;
  __syncthreads();
  for (int base = 0; base < N; base += 1024) {
    const int i = base + tid;
    int c = 0;
    if (i < N)
      for (int ch = 0; ch < n_chunks; ++ch) c += mat[(size_t)ch * N + i];
    int incl = c;
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
    const int excl = carry + warp_tot[warp] + incl - c;
    if (i < N) {
      rp[i] = excl;
      int run = excl;
      for (int ch = 0; ch < n_chunks; ++ch) {  // chunk ch writes its entries of support i from here on
        const int k = mat[(size_t)ch * N + i];
        mat[(size_t)ch * N + i] = run;
        run += k;
      }
    }
    __syncthreads();
    if (tid == 1023) carry = excl + c;
    __syncthreads();
  }
  if (b == B - 1 && tid == 0) rowptr[(size_t)B * N] = B * P;
}

__global__ void __launch_bounds__(512)
chunk_fill_kernel(const int* __restrict__ idx, int P, int N, int nsample, int n_chunks, int chunk_len,
                  const int* __restrict__ matrix, int* __restrict__ unsorted) {
  extern __shared__ int cursor[];
  const int chunk = blockIdx.x, b = blockIdx.y;
  const int* row = matrix + ((size_t)b * n_chunks + chunk) * N;
  for (int i = threadIdx.x; i < N; i += blockDim.x) cursor[i] = row[i];
  __syncthreads();
  const int beg = chunk * chunk_len, end = min(beg + chunk_len, P);
  const int* src = idx + (size_t)b * P;
  for (int e = beg + threadIdx.x; e < end; e += blockDim.x) {
    const int pos = atomicAdd(&cursor[d3d_clamp_index(src[e], N)], 1);
    const int j = e / nsample;
    unsorted[pos] = (j << 8) | (e - j * nsample);
  }
}

// warp per support: its segment is n_chunks pieces (already in chunk order); rank-sort each piece
__global__ void __launch_bounds__(256)
chunk_sort_kernel(const int* __restrict__ rowptr, const int* __restrict__ matrix, int B, int N, int n_chunks,
                  const int* __restrict__ unsorted, int* __restrict__ entries) {
  const int lane = threadIdx.x & 31;
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= (long long)B * N) return;
  const int b = (int)(row / N), i = (int)(row - (long long)b * N);
  const int seg_beg = rowptr[row], seg_end = rowptr[row + 1];
  const int seg_len = seg_end - seg_beg;
  if (seg_len <= 1) {
    if (seg_len == 1 && lane == 0) entries[seg_beg] = unsorted[seg_beg];
    return;
  }
  // piece boundaries: lane ch holds where piece ch starts (the fill cursors' initial values)
  const int* mat = matrix + (size_t)b * n_chunks * N + i;
  const int my_start = lane < n_chunks ? (lane == 0 ? seg_beg : mat[(size_t)lane * N]) : seg_end;
  if (seg_len <= 32) {
    // whole segment in one pass: rank among the entries of the same piece, offset by the piece start
    const int pos = seg_beg + lane;
    const int own = lane < seg_len ? unsorted[pos] : 0x7fffffff;
    // piece of this lane's entry = number of piece starts <= pos, minus 1
    int piece = 0;
    for (int ch = 1; ch < n_chunks; ++ch) piece += __shfl_sync(D3D_FULL_MASK, my_start, ch) <= pos ? 1 : 0;
    const int pstart = __shfl_sync(D3D_FULL_MASK, my_start, piece);
    int r = 0;
    for (int t = 0; t < seg_len; ++t) {
      const int other = __shfl_sync(D3D_FULL_MASK, own, t);
      const int opiece = __shfl_sync(D3D_FULL_MASK, piece, t);
      r += (opiece == piece && other < own) ? 1 : 0;
    }
    if (lane < seg_len) entries[pstart + r] = own;
    return;
  }
  int beg = seg_beg;
  for (int ch = 0; ch < n_chunks; ++ch) {
    const int end = __shfl_sync(D3D_FULL_MASK, my_start, ch + 1 < 32 ? ch + 1 : 31);  // lanes >= n_chunks hold seg_end
    const int len = end - beg;
    if (len == 1) {
      if (lane == 0) entries[beg] = unsorted[beg];
    } else if (len > 1 && len <= 64) {
      const int own0 = lane < len ? unsorted[beg + lane] : 0x7fffffff;
      const int own1 = lane + 32 < len ? unsorted[beg + lane + 32] : 0x7fffffff;
      int r0 = 0, r1 = 0;
      for (int t = 0; t < len; ++t) {  // entries are unique -> rank = number of smaller entries
        const int other = t < 32 ? __shfl_sync(D3D_FULL_MASK, own0, t) : __shfl_sync(D3D_FULL_MASK, own1, t - 32);
        r0 += other < own0 ? 1 : 0;
        r1 += other < own1 ? 1 : 0;
      }
      if (lane < len) entries[beg + r0] = own0;
      if (lane + 32 < len) entries[beg + r1] = own1;
    } else if (len > 64) {
      for (int t = lane; t < len; t += 32) {  // rare: out-of-place rank counting straight from L1/L2
        const int own = unsorted[beg + t];
        int r = 0;
        for (int u = 0; u < len; ++u) r += unsorted[beg + u] < own ? 1 : 0;
        entries[beg + r] = own;
      }
    }
    beg = end;
  }
}

struct InvWs {
  int* cursor;
  int* unsorted;
  int* long_rows;
  int* long_count;
};

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

InvWs carve(void* ws, int B, int N, int M, int nsample) {
  unsigned char* p = (unsigned char*)ws;
  InvWs w;
  w.cursor = (int*)p; p += align256((size_t)B * N * sizeof(int));
  w.unsorted = (int*)p; p += align256((size_t)B * M * nsample * sizeof(int));
  w.long_rows = (int*)p; p += align256((size_t)B * N * sizeof(int));
  w.long_count = (int*)p;
  return w;
}

}  // namespace

extern "C" {

size_t d3d_inverse_map_workspace_bytes(int B, int N, int M, int nsample) {
  if (B <= 0 || N <= 0 || M <= 0 || nsample <= 0) return 0;
  const size_t global_path = 2 * align256((size_t)B * N * sizeof(int)) + align256((size_t)B * M * nsample * sizeof(int)) + 256;
  const size_t chunk_path = align256((size_t)B * kMaxChunks * N * sizeof(int)) + align256((size_t)B * M * nsample * sizeof(int));
  return global_path > chunk_path ? global_path : chunk_path;
}

int d3d_build_inverse_map(const int* idx, int B, int N, int M, int nsample, int* rowptr, int* entries, void* ws,
                          size_t ws_bytes, void* stream) {
  D3D_REQUIRE(idx && rowptr && entries);
  D3D_REQUIRE(B >= 0 && N > 0 && M >= 0 && nsample > 0 && nsample <= 256);
  const long long total = (long long)B * M * nsample;
  if (total >= (1ll << 31) || (long long)M >= (1ll << 23)) return D3D_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0) return 0;
  if (!ws || ws_bytes < d3d_inverse_map_workspace_bytes(B, N, M, nsample)) return D3D_ERR_WORKSPACE;
  const int P = M * nsample;
  if (N <= kChunkMaxN && total > 0) {
    int n_chunks = (P + 8191) / 8192;
    if (n_chunks > kMaxChunks) n_chunks = kMaxChunks;
    const int chunk_len = (P + n_chunks - 1) / n_chunks;
    int* matrix = (int*)ws;
    int* unsorted = (int*)((unsigned char*)ws + align256((size_t)B * kMaxChunks * N * sizeof(int)));
    const size_t smem = (size_t)N * sizeof(int);
    cudaError_t e1 = cudaFuncSetAttribute(chunk_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(chunk_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e1 != cudaSuccess) return (int)e1;
    dim3 grid(n_chunks, B);
    chunk_count_kernel<<<grid, 512, smem, st>>>(idx, P, N, n_chunks, chunk_len, matrix);
    chunk_scan_kernel<<<B, 1024, 0, st>>>(matrix, N, P, B, n_chunks, rowptr);
    chunk_fill_kernel<<<grid, 512, smem, st>>>(idx, P, N, nsample, n_chunks, chunk_len, matrix, unsorted);
    chunk_sort_kernel<<<d3d_ceil_div((long long)B * N * 32, 256), 256, 0, st>>>(rowptr, matrix, B, N, n_chunks, unsorted, entries);
    d3d_note_launches(4);
    return d3d_launch_status();
  }
  InvWs w = carve(ws, B, N, M, nsample);
  cudaError_t e = cudaMemsetAsync(w.cursor, 0, (size_t)B * N * sizeof(int), st);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemsetAsync(w.long_count, 0, sizeof(int), st);
  if (e != cudaSuccess) return (int)e;
  d3d_note_launches(total > 0 ? 5 : 1);
  if (total > 0) count_kernel<<<d3d_ceil_div(total, 256), 256, 0, st>>>(idx, total, P, N, w.cursor);
  scan_kernel<<<B, 1024, 0, st>>>(w.cursor, N, P, B, rowptr);
  if (total > 0) {
    fill_kernel<<<d3d_ceil_div(total, 256), 256, 0, st>>>(idx, total, P, N, nsample, w.cursor, w.unsorted);
    const int rows = B * N;
    sort_short_kernel<<<d3d_ceil_div((long long)rows * 32, 256), 256, 0, st>>>(rowptr, rows, w.unsorted, entries,
                                                                             w.long_rows, w.long_count);
    sort_long_kernel<<<148 * 4, kSortThreads, 0, st>>>(rowptr, w.long_rows, w.long_count, w.unsorted, entries);
  }
  return d3d_launch_status();
}

}  // extern "C"

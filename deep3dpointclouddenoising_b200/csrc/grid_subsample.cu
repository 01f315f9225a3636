// Masked grid subsampling for sm_100a: one 1024-thread block per cloud, everything on chip.
//
// Semantics (bit-exact with the reference):
//   ref: u_net_arch/pt_custom_ops/_ext_src/src/masked_grid_subsampling_gpu.cu:31-152
// The reference runs the whole algorithm in ONE thread per cloud with two in-thread merge sorts in
// global scratch.  Here:
//   1. bounding box over all N points (block min/max reduction; min/max are order independent);
//   2. voxel id per valid point with the reference's exact float expression (see voxel_coord);
//   3. 64-bit composite keys (voxel id, point index) sorted by a shared-memory bitonic network —
//      the composite key is unique, so the order equals the reference's STABLE sort by voxel id;
//   4. cell boundaries by a block-wide scan; the thread that owns a cell start adds the members in
//      ascending point index and divides once (same float summation order as the reference);
//   5. the reference's LCG(17,139,256) key + stable sort "shuffle" in closed form: the LCG has full
//      period 256, so the key of cell ordinal c depends on c mod 256 only and the sorted position is
//      start[key] + c / 256;
//   6. cyclic padding up to m rows.
#include "common.cuh"

namespace {

constexpr int kGsThreads = 1024;
constexpr int kGsSmemPoints = 16384;  // clouds up to this many points are sorted in shared memory

__device__ __forceinline__ int voxel_coord(float x, float fl, float dl) {
  // reference: (int)floor((x - origin) / dl) with origin = floor(min * (1/dl)) * dl.  nvcc contracts
  // "x - floor(.)*dl" into one FFMA (SASS of the reference kernel: FFMA R, -Rfloor, Rdl, Rx), the
  // division is IEEE (masked_grid_subsampling_gpu.cu:48-50, 67-69).
  return (int)floorf(__fdiv_rn(__fmaf_rn(-fl, dl, x), dl));
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(D3D_FULL_MASK, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(D3D_FULL_MASK, v, o));
  return v;
}

__global__ void __launch_bounds__(kGsThreads)
grid_subsample_kernel(const float* __restrict__ xyz, const int* __restrict__ mask, int N, int m, float dl,
                      float* sub_xyz, int* __restrict__ sub_mask, unsigned long long* global_keys, size_t key_stride,
                      int keys_in_smem) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float red[6][32];
  __shared__ float box[6];
  __shared__ int s_first_zero;
  __shared__ int warp_tot[32];
  __shared__ int s_ncell;
  __shared__ int cyc[256], inv_pos[256], start[256];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  const float* P = xyz + (size_t)b * N * 3;
  const int* mk = mask + (size_t)b * N;
  float* out = sub_xyz + (size_t)b * m * 3;
  int* outm = sub_mask + (size_t)b * m;

  // ---- 1. bounding box over ALL points, padding included (:31-46), and the valid prefix length
  if (tid == 0) s_first_zero = N;
  float lo0 = P[0], lo1 = P[1], lo2 = P[2], hi0 = lo0, hi1 = lo1, hi2 = lo2;
  int first_zero = N;
  for (int i = tid; i < N; i += kGsThreads) {
    const float x = P[3 * i], y = P[3 * i + 1], z = P[3 * i + 2];
    lo0 = fminf(lo0, x); lo1 = fminf(lo1, y); lo2 = fminf(lo2, z);
    hi0 = fmaxf(hi0, x); hi1 = fmaxf(hi1, y); hi2 = fmaxf(hi2, z);
    if (first_zero == N && mk[i] == 0) first_zero = i;
  }
  lo0 = warp_min(lo0); lo1 = warp_min(lo1); lo2 = warp_min(lo2);
  hi0 = warp_max(hi0); hi1 = warp_max(hi1); hi2 = warp_max(hi2);
  if (lane == 0) { red[0][warp] = lo0; red[1][warp] = lo1; red[2][warp] = lo2; red[3][warp] = hi0; red[4][warp] = hi1; red[5][warp] = hi2; }
  __syncthreads();
  if (first_zero < N) atomicMin(&s_first_zero, first_zero);
  if (warp < 6) {
    float v = red[warp][lane];
    v = warp < 3 ? warp_min(v) : warp_max(v);
    if (lane == 0) box[warp] = v;
  }
  __syncthreads();
  int v = s_first_zero;
  const bool empty = (v == 0);  // reference reads its zero-filled scratch: one cell made of point 0
  if (empty) v = 1;

  const float inv = __fdiv_rn(1.0f, dl);
  const float fl0 = floorf(__fmul_rn(box[0], inv)), fl1 = floorf(__fmul_rn(box[1], inv)), fl2 = floorf(__fmul_rn(box[2], inv));
  const int NX = voxel_coord(box[3], fl0, dl) + 1;
  const int NY = voxel_coord(box[4], fl1, dl) + 1;

  // ---- 2./3. composite keys, padded to a power of two with +inf sentinels, bitonic sort
  int P2 = 1;
  while (P2 < v) P2 <<= 1;
  unsigned long long* keys = keys_in_smem ? reinterpret_cast<unsigned long long*>(smem_raw)
                                          : global_keys + (size_t)b * key_stride;
  for (int i = tid; i < P2; i += kGsThreads) {
    unsigned long long key = ~0ull;
    if (i < v) {
      int id = 0;
      if (!empty) {
        const int ix = voxel_coord(P[3 * i], fl0, dl), iy = voxel_coord(P[3 * i + 1], fl1, dl), iz = voxel_coord(P[3 * i + 2], fl2, dl);
        id = ix + NX * iy + NX * NY * iz;  // int32 arithmetic like the reference (:71)
      }
      key = ((unsigned long long)((unsigned)id ^ 0x80000000u) << 32) | (unsigned)i;
    }
    keys[i] = key;
  }
  __syncthreads();
  for (int k = 2; k <= P2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (P2 >> 1); t += kGsThreads) {
        const int a = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int c = a | j;
        const unsigned long long ka = keys[a], kc = keys[c];
        const bool up = (a & k) == 0;
        if ((ka > kc) == up) { keys[a] = kc; keys[c] = ka; }
      }
      __syncthreads();
    }
  }

  // ---- 4. cell starts -> ordinals (block exclusive scan over per-thread counts)
  const int chunk = (P2 + kGsThreads - 1) / kGsThreads;
  const int i_begin = tid * chunk, i_end = min(i_begin + chunk, v);
  int local = 0;
  for (int i = i_begin; i < i_end; ++i)
    local += (i == 0 || (unsigned)(keys[i] >> 32) != (unsigned)(keys[i - 1] >> 32)) ? 1 : 0;
  int incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int up = __shfl_up_sync(D3D_FULL_MASK, incl, o);
    if (lane >= o) incl += up;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = warp_tot[lane], wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(D3D_FULL_MASK, wi, o);
      if (lane >= o) wi += up;
    }
    warp_tot[lane] = wi - w;  // exclusive
    if (lane == 31) s_ncell = wi;
  }
  __syncthreads();
  const int ncell = s_ncell;
  int ordinal = warp_tot[warp] + incl - local;

  // ---- 5. closed-form position of every cell ordinal after the LCG "shuffle" (:124-135)
  if (tid < 256) {
    int k0 = (int)((unsigned)(keys[0] >> 32) ^ 0x80000000u) % 256;
    if (k0 < 0) k0 += 256;  // only after int32 overflow of the voxel id; outside the supported domain
    int x = k0;
    for (int s = 0; s < tid; ++s) x = (17 * x + 139) & 255;
    cyc[tid] = x;
    inv_pos[x] = tid;  // full-period LCG: a bijection on 0..255
  }
  __syncthreads();
  if (tid < 256) {
    int acc = 0;
    for (int val = 0; val < tid; ++val) {
      const int p = inv_pos[val];
      acc += p < ncell ? (ncell - 1 - p) / 256 + 1 : 0;
    }
    start[tid] = acc;
  }
  __syncthreads();

  // ---- per-cell barycentre: members in ascending point index, one division (:79-122)
  for (int i = i_begin; i < i_end; ++i) {
    const unsigned cell = (unsigned)(keys[i] >> 32);
    if (i != 0 && cell == (unsigned)(keys[i - 1] >> 32)) continue;
    const int c = ordinal++;
    const int pos = start[cyc[c & 255]] + (c >> 8);
    if (pos >= m) continue;  // shuffled tail dropped when there are more cells than m (:137)
    int p = (int)(unsigned)(keys[i] & 0xffffffffull);
    float xs = P[3 * p], ys = P[3 * p + 1], zs = P[3 * p + 2], pnum = 1.0f;
    for (int t = i + 1; t < v && (unsigned)(keys[t] >> 32) == cell; ++t) {
      p = (int)(unsigned)(keys[t] & 0xffffffffull);
      xs = __fadd_rn(xs, P[3 * p]); ys = __fadd_rn(ys, P[3 * p + 1]); zs = __fadd_rn(zs, P[3 * p + 2]);
      pnum += 1.0f;
    }
    out[3 * pos] = __fdiv_rn(xs, pnum); out[3 * pos + 1] = __fdiv_rn(ys, pnum); out[3 * pos + 2] = __fdiv_rn(zs, pnum);
    outm[pos] = 1;
  }
  __syncthreads();
  // ---- 6. cyclic padding with real sub points (:145-151)
  for (int i = ncell + tid; i < m; i += kGsThreads) {
    const int src = i % ncell;
    out[3 * i] = out[3 * src]; out[3 * i + 1] = out[3 * src + 1]; out[3 * i + 2] = out[3 * src + 2];
    outm[i] = 0;
  }
}

size_t pow2_at_least(size_t n) {
  size_t p = 1;
  while (p < n) p <<= 1;
  return p;
}

}  // namespace

extern "C" {

size_t d3d_grid_subsample_workspace_bytes(int B, int N) {
  if (B <= 0 || N <= kGsSmemPoints) return 0;
  return (size_t)B * pow2_at_least((size_t)N) * sizeof(unsigned long long);
}

int d3d_grid_subsample(const float* xyz, const int* mask, int B, int N, int m, float sample_dl, float* sub_xyz,
                       int* sub_mask, void* ws, size_t ws_bytes, void* stream) {
  D3D_REQUIRE(xyz && mask && sub_xyz && sub_mask);
  D3D_REQUIRE(B >= 0 && N > 0 && m > 0 && sample_dl > 0.0f);
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const bool in_smem = N <= kGsSmemPoints;
  size_t smem = 0;
  if (in_smem) {
    smem = pow2_at_least((size_t)N) * sizeof(unsigned long long);
    cudaError_t e = cudaFuncSetAttribute(grid_subsample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  } else if (!ws || ws_bytes < d3d_grid_subsample_workspace_bytes(B, N)) {
    return D3D_ERR_WORKSPACE;
  }
  grid_subsample_kernel<<<B, kGsThreads, smem, st>>>(xyz, mask, N, m, sample_dl, sub_xyz, sub_mask,
                                                     (unsigned long long*)ws, pow2_at_least((size_t)N), in_smem ? 1 : 0);
  d3d_note_launches(1);
  return d3d_launch_status();
}

}  // extern "C"

// group_points / group_points_grad in the reference's own layouts, plus the (B,C,N) <-> (B,N,C)
// layout changes the fused kernels need.
//   ref: u_net_arch/pt_custom_ops/_ext_src/src/group_points_gpu.cu:13-33 (gather), :48-69 (atomicAdd scatter)
//
// Gather: the (B, C, M*nsample) output is the whole cost (1.96 GB at B=16, C=72, M=8192, ns=52), so a
// thread owns 4 consecutive output positions, keeps their 4 indices in registers for every channel
// (the reference re-reads idx per channel) and writes one coalesced float4 per channel.
// Gradient: a segmented sum over the inverse map — deterministic, no float atomics.
#include "common.cuh"

namespace {

template <int VEC>
__global__ void __launch_bounds__(256)
group_points_kernel(const float* __restrict__ points, const int* __restrict__ idx, int C, int N, int P,
                    float* __restrict__ out) {
  const int b = blockIdx.y;
  const long long p0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (p0 >= P) return;
  const int* id = idx + (size_t)b * P + p0;
  int i[VEC];
  if (VEC == 4) {
    const int4 v = *reinterpret_cast<const int4*>(id);
    i[0] = v.x; i[1 % VEC] = v.y; i[2 % VEC] = v.z; i[3 % VEC] = v.w;
  } else {
    i[0] = id[0];
  }
#pragma unroll
  for (int e = 0; e < VEC; ++e) i[e] = d3d_clamp_index(i[e], N);
  const float* src = points + (size_t)b * C * N;
  float* dst = out + (size_t)b * C * P + p0;
#pragma unroll 4
  for (int c = 0; c < C; ++c) {
    const float* row = src + (size_t)c * N;
    if (VEC == 4) {
      float4 v;
      v.x = __ldg(row + i[0]); v.y = __ldg(row + i[1 % VEC]); v.z = __ldg(row + i[2 % VEC]); v.w = __ldg(row + i[3 % VEC]);
      __stcs(reinterpret_cast<float4*>(dst + (size_t)c * P), v);  // streamed: never re-read by this kernel
    } else {
      dst[(size_t)c * P] = __ldg(row + i[0]);
    }
  }
}

constexpr int kGradChan = 8;

// thread = (support i, chunk of kGradChan channels): walks the support's sorted segment
__global__ void __launch_bounds__(128)
group_points_grad_kernel(const float* __restrict__ grad_out, const int* __restrict__ rowptr,
                         const int* __restrict__ entries, int C, int N, int P, int nsample,
                         float* __restrict__ grad_points) {
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * kGradChan;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int row = b * N + i;
  const int beg = rowptr[row], end = rowptr[row + 1];
  float acc[kGradChan];
#pragma unroll
  for (int c = 0; c < kGradChan; ++c) acc[c] = 0.0f;
  const float* g = grad_out + ((size_t)b * C + c0) * P;
  const int nc = min(kGradChan, C - c0);
  for (int e = beg; e < end; ++e) {
    const int packed = entries[e];
    const int p = (packed >> 8) * nsample + (packed & 255);
#pragma unroll
    for (int c = 0; c < kGradChan; ++c)
      if (c < nc) acc[c] += __ldg(g + (size_t)c * P + p);
  }
#pragma unroll
  for (int c = 0; c < kGradChan; ++c)
    if (c < nc) grad_points[((size_t)b * C + c0 + c) * N + i] = acc[c];
}

// Fast form of the gather gradient: one block per (b, c) plane.  The N sums of the plane live in shared memory, the
// (M, nsample) gradients and indices stream through coalesced (float4 / int4) and are added with shared-memory atomics —
// the reference's own formulation (atomicAdd, group_points_gpu.cu:61-66) without its global-memory contention.
// Reproducible to rounding only; d3d_group_points_grad (inverse map, fixed order) is the bit-reproducible form.
__global__ void __launch_bounds__(1024)
group_points_grad_plane_kernel(const float* __restrict__ grad_out, const int* __restrict__ idx, int C, int N, int P,
                               float* __restrict__ grad_points) {
  extern __shared__ float acc[];
  const int c = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  for (int i = tid; i < N; i += blockDim.x) acc[i] = 0.0f;
  __syncthreads();
  const float* g = grad_out + ((size_t)b * C + c) * P;
  const int* ix = idx + (size_t)b * P;
  if ((P & 3) == 0) {
    const float4* g4 = reinterpret_cast<const float4*>(g);
    const int4* i4 = reinterpret_cast<const int4*>(ix);
    for (int p = tid; p < P / 4; p += blockDim.x) {
      const float4 v = __ldg(g4 + p);
      const int4 j = __ldg(i4 + p);
      if ((unsigned)j.x < (unsigned)N) atomicAdd(&acc[j.x], v.x);
      if ((unsigned)j.y < (unsigned)N) atomicAdd(&acc[j.y], v.y);
      if ((unsigned)j.z < (unsigned)N) atomicAdd(&acc[j.z], v.z);
      if ((unsigned)j.w < (unsigned)N) atomicAdd(&acc[j.w], v.w);
    }
  } else {
    for (int p = tid; p < P; p += blockDim.x) {
      const int j = __ldg(ix + p);
      if ((unsigned)j < (unsigned)N) atomicAdd(&acc[j], __ldg(g + p));
    }
  }
  __syncthreads();
  float* o = grad_points + ((size_t)b * C + c) * N;
  for (int i = tid; i < N; i += blockDim.x) o[i] = acc[i];
}

// (rows x cols) -> (cols x rows) per batch through a padded 32x32 shared tile
__global__ void __launch_bounds__(256)
transpose_kernel(const float* __restrict__ src, int rows, int cols, float* __restrict__ dst) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const float* s = src + (size_t)b * rows * cols;
  float* d = dst + (size_t)b * rows * cols;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int r = r0 + ty + k, c = c0 + tx;
    if (r < rows && c < cols) tile[ty + k][tx] = s[(size_t)r * cols + c];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int c = c0 + ty + k, r = r0 + tx;
    if (r < rows && c < cols) d[(size_t)c * rows + r] = tile[tx][ty + k];
  }
}

int launch_transpose(const float* src, int B, int rows, int cols, float* dst, cudaStream_t st) {
  if (B == 0 || rows == 0 || cols == 0) return 0;
  dim3 grid(d3d_ceil_div(cols, 32), d3d_ceil_div(rows, 32), B);
  transpose_kernel<<<grid, 256, 0, st>>>(src, rows, cols, dst);
  d3d_note_launches(1);
  return d3d_launch_status();
}

}  // namespace

extern "C" {

int d3d_group_points(const float* points, const int* idx, int B, int C, int N, int M, int nsample, float* out,
                     void* stream) {
  D3D_REQUIRE(points && idx && out);
  D3D_REQUIRE(B >= 0 && C >= 0 && N > 0 && M >= 0 && nsample > 0);
  const long long P = (long long)M * nsample;
  if (P >= (1ll << 31)) return D3D_ERR_UNSUPPORTED;
  if (B == 0 || C == 0 || P == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (P % 4 == 0) {
    dim3 grid(d3d_ceil_div(P / 4, 256), B);
    group_points_kernel<4><<<grid, 256, 0, st>>>(points, idx, C, N, (int)P, out);
  } else {
    dim3 grid(d3d_ceil_div(P, 256), B);
    group_points_kernel<1><<<grid, 256, 0, st>>>(points, idx, C, N, (int)P, out);
  }
  d3d_note_launches(1);
  return d3d_launch_status();
}

size_t d3d_group_points_grad_workspace_bytes(int B, int N, int M, int nsample) {
  if (B <= 0 || N <= 0 || M <= 0 || nsample <= 0) return 0;
  const size_t rowptr = (((size_t)B * N + 1) * sizeof(int) + 255) & ~(size_t)255;
  const size_t entries = ((size_t)B * M * nsample * sizeof(int) + 255) & ~(size_t)255;
  return rowptr + entries + d3d_inverse_map_workspace_bytes(B, N, M, nsample);
}

int d3d_group_points_grad(const float* grad_out, const int* idx, int B, int C, int N, int M, int nsample,
                          float* grad_points, void* ws, size_t ws_bytes, void* stream) {
  D3D_REQUIRE(grad_out && idx && grad_points);
  D3D_REQUIRE(B >= 0 && C >= 0 && N > 0 && M >= 0 && nsample > 0 && nsample <= 256);
  if (B == 0 || C == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (M == 0) return (int)cudaMemsetAsync(grad_points, 0, (size_t)B * C * N * sizeof(float), st);
  if (!ws || ws_bytes < d3d_group_points_grad_workspace_bytes(B, N, M, nsample)) return D3D_ERR_WORKSPACE;
  unsigned char* p = (unsigned char*)ws;
  int* rowptr = (int*)p;
  p += (((size_t)B * N + 1) * sizeof(int) + 255) & ~(size_t)255;
  int* entries = (int*)p;
  p += ((size_t)B * M * nsample * sizeof(int) + 255) & ~(size_t)255;
  const int rc = d3d_build_inverse_map(idx, B, N, M, nsample, rowptr, entries, p,
                                       ws_bytes - (size_t)(p - (unsigned char*)ws), stream);
  if (rc != 0) return rc;
  dim3 grid(d3d_ceil_div(N, 128), d3d_ceil_div(C, kGradChan), B);
  group_points_grad_kernel<<<grid, 128, 0, st>>>(grad_out, rowptr, entries, C, N, M * nsample, nsample, grad_points);
  d3d_note_launches(1);
  return d3d_launch_status();
}

int d3d_group_points_grad_atomic(const float* grad_out, const int* idx, int B, int C, int N, int M, int nsample,
                                 float* grad_points, void* stream) {
  D3D_REQUIRE(grad_out && idx && grad_points);
  D3D_REQUIRE(B >= 0 && C >= 0 && N > 0 && M >= 0 && nsample > 0);
  if (B == 0 || C == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const long long P = (long long)M * nsample;
  if (P >= (1ll << 31) || (size_t)N * sizeof(float) > 200 * 1024) return D3D_ERR_UNSUPPORTED;
  if (M == 0) return (int)cudaMemsetAsync(grad_points, 0, (size_t)B * C * N * sizeof(float), st);
  const size_t smem = (size_t)N * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(group_points_grad_plane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  dim3 grid(C, B);
  group_points_grad_plane_kernel<<<grid, 1024, smem, st>>>(grad_out, idx, C, N, (int)P, grad_points);
  d3d_note_launches(1);
  return d3d_launch_status();
}

int d3d_cm_to_cl(const float* src_cm, int B, int C, int N, float* dst_cl, void* stream) {
  D3D_REQUIRE(src_cm && dst_cl && B >= 0 && C >= 0 && N >= 0);
  return launch_transpose(src_cm, B, C, N, dst_cl, (cudaStream_t)stream);
}

int d3d_cl_to_cm(const float* src_cl, int B, int C, int N, float* dst_cm, void* stream) {
  D3D_REQUIRE(src_cl && dst_cm && B >= 0 && C >= 0 && N >= 0);
  return launch_transpose(src_cl, B, N, C, dst_cm, (cudaStream_t)stream);
}

static long long g_launches = 0;
long long d3d_kernel_launches(void) { return g_launches; }

int d3d_abi_version(void) { return 4; }

const char* d3d_error_string(int code) {
  switch (code) {
    case 0: return "ok";
    case D3D_ERR_BAD_ARG: return "d3d: bad argument (null pointer or size out of range)";
    case D3D_ERR_UNSUPPORTED: return "d3d: size not supported by the sm_100a kernels";
    case D3D_ERR_WORKSPACE: return "d3d: workspace missing or too small";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "d3d: unknown error";
  }
}

}  // extern "C"

void d3d_note_launches(int n) { __atomic_fetch_add(&g_launches, (long long)n, __ATOMIC_RELAXED); }

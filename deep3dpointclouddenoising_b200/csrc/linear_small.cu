// 1x1 convolutions with a tiny channel count on one side — the U-Net's input layer (3 -> 72) and its last layer
// (-> 3 offsets): row-major y[R x N] = x[R x K] . w[N x K]^T with K <= 4 or N <= 4.
//
//   ref: u_net_arch/models/backbones/resnet.py:100-103 (conv1 on the 3 input features)
//   ref: u_net_arch/models/heads/multi_dimensional_head.py:45-50 (head: ... -> Conv1d(width, num_classes = 3))
//
// TMA needs 16-byte rows, so the tensor-core GEMM (gemm.cu) does not take these shapes; they are pure streaming work
// (the wide side is read or written once), done here on the CUDA cores in fp32:
//   linear_small_k : K <= 4.  thread = (row, 4 output channels); also the DATA gradient of the last layer
//                    (dx[R x K'] = dy[R x 3] . W[3 x K']: the weight is read through strides, no transposition)
//   linear_small_n : N <= 4.  warp = row, lanes over K, shuffle reduction
//   wgrad_small    : out[Nb x Ks] (strided) = sum_r big[r, Nb] * small[r, Ks] — the weight gradients of both layers;
//                    row-range partials + a fixed-order second pass (no float atomics)
#include "common.cuh"

namespace {

constexpr int kMaxSmall = 4;

__global__ void __launch_bounds__(256)
linear_small_k_kernel(const float* __restrict__ x, const float* __restrict__ w, long long w_sn, long long w_sk,
                      const float* __restrict__ bias, long long R, int K, int N, float* __restrict__ y) {
  extern __shared__ float sw[];  // [N][kMaxSmall] weights, [N] bias
  for (int i = threadIdx.x; i < N * kMaxSmall; i += blockDim.x) {
    const int n = i / kMaxSmall, k = i % kMaxSmall;
    sw[i] = k < K ? w[n * w_sn + k * w_sk] : 0.0f;
  }
  for (int i = threadIdx.x; i < N; i += blockDim.x) sw[N * kMaxSmall + i] = bias ? bias[i] : 0.0f;
  __syncthreads();
  const int groups = N >> 2;
  const int lanes = blockDim.x / groups;          // rows per pass of the block
  const int g = threadIdx.x % groups, rl = threadIdx.x / groups;
  if (rl >= lanes) return;
  float wreg[4][kMaxSmall], breg[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    breg[j] = sw[N * kMaxSmall + 4 * g + j];
#pragma unroll
    for (int k = 0; k < kMaxSmall; ++k) wreg[j][k] = sw[(4 * g + j) * kMaxSmall + k];
  }
  for (long long r = (long long)blockIdx.x * lanes + rl; r < R; r += (long long)gridDim.x * lanes) {
    float xv[kMaxSmall];
#pragma unroll
    for (int k = 0; k < kMaxSmall; ++k) xv[k] = k < K ? __ldg(x + r * K + k) : 0.0f;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float acc = breg[j];
#pragma unroll
      for (int k = 0; k < kMaxSmall; ++k) acc = fmaf(xv[k], wreg[j][k], acc);
      o[j] = acc;
    }
    reinterpret_cast<float4*>(y + r * N)[g] = make_float4(o[0], o[1], o[2], o[3]);
  }
}

constexpr int kSnRows = 64;  // rows per block pass of the small-N kernel

__global__ void __launch_bounds__(256)
linear_small_n_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, long long R,
                      int K, int N, float* __restrict__ y) {
  extern __shared__ float tile[];  // [kSnRows][K + 1] rows of x (padded: conflict-free column walks), then [N][K] weights
  float* sw = tile + (size_t)kSnRows * (K + 1);
  for (int i = threadIdx.x; i < N * K; i += blockDim.x) sw[i] = w[i];
  const int k4 = K >> 2;
  const int row = threadIdx.x & (kSnRows - 1), n = threadIdx.x / kSnRows;  // 256 threads = 64 rows x 4 outputs
  for (long long r0 = (long long)blockIdx.x * kSnRows; r0 < R; r0 += (long long)gridDim.x * kSnRows) {
    __syncthreads();
    const int rows_here = (int)min((long long)kSnRows, R - r0);
    for (int i = threadIdx.x; i < rows_here * k4; i += blockDim.x) {  // coalesced float4 loads of the row block
      const int rr = i / k4, c = i - rr * k4;
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + (r0 + rr) * K) + c);
      float* d = tile + (size_t)rr * (K + 1) + 4 * c;
      d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
    __syncthreads();
    if (n < N && row < rows_here) {
      const float* xr = tile + (size_t)row * (K + 1);
      const float* wr = sw + (size_t)n * K;
      float acc = bias ? bias[n] : 0.0f;
      for (int k = 0; k < K; ++k) acc = fmaf(xr[k], wr[k], acc);
      y[(r0 + row) * N + n] = acc;
    }
  }
}

// partial[blk][nb][ks] over the block's row range; thread = (row lane, 4 channels of the wide operand)
__global__ void __launch_bounds__(256)
wgrad_small_partial_kernel(const float* __restrict__ big, const float* __restrict__ small, long long R, int Nb, int Ks,
                           long long rows_per_block, float* __restrict__ partial) {
  extern __shared__ float red[];  // [lanes][Nb][kMaxSmall]
  const int groups = Nb >> 2, lanes = blockDim.x / groups;
  const int g = threadIdx.x % groups, rl = threadIdx.x / groups;
  const long long beg = (long long)blockIdx.x * rows_per_block, end = min(R, beg + rows_per_block);
  float acc[4][kMaxSmall];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < kMaxSmall; ++k) acc[j][k] = 0.0f;
  if (rl < lanes)
    for (long long r = beg + rl; r < end; r += lanes) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(big + r * Nb) + g);
      float s[kMaxSmall];
#pragma unroll
      for (int k = 0; k < kMaxSmall; ++k) s[k] = k < Ks ? __ldg(small + r * Ks + k) : 0.0f;
#pragma unroll
      for (int k = 0; k < kMaxSmall; ++k) {
        acc[0][k] = fmaf(v.x, s[k], acc[0][k]); acc[1][k] = fmaf(v.y, s[k], acc[1][k]);
        acc[2][k] = fmaf(v.z, s[k], acc[2][k]); acc[3][k] = fmaf(v.w, s[k], acc[3][k]);
      }
    }
  if (rl < lanes)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < kMaxSmall; ++k) red[((size_t)rl * Nb + 4 * g + j) * kMaxSmall + k] = acc[j][k];
  __syncthreads();
  for (int i = threadIdx.x; i < Nb * kMaxSmall; i += blockDim.x) {
    float sum = 0.0f;
    for (int l = 0; l < lanes; ++l) sum += red[(size_t)l * Nb * kMaxSmall + i];  // fixed order
    partial[(size_t)blockIdx.x * Nb * kMaxSmall + i] = sum;
  }
}

__global__ void __launch_bounds__(256)
wgrad_small_reduce_kernel(const float* __restrict__ partial, int nblk, int Nb, int Ks, long long out_sn, long long out_sk,
                          int accumulate, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Nb * Ks) return;
  const int n = i / Ks, k = i % Ks;
  double sum = 0.0;
  for (int b = 0; b < nblk; ++b) sum += (double)partial[((size_t)b * Nb + n) * kMaxSmall + k];  // fixed order
  float* o = out + n * out_sn + k * out_sk;
  *o = (accumulate ? *o : 0.0f) + (float)sum;
}

constexpr int kWgradBlocks = 148 * 2;

}  // namespace

extern "C" {

int d3d_linear_small_k(const float* x, const float* w, long long w_stride_n, long long w_stride_k, const float* bias,
                       long long R, int K, int N, float* y, void* stream) {
  D3D_REQUIRE(x && w && y && R >= 0 && K > 0 && K <= kMaxSmall && N > 0 && N % 4 == 0);
  if (((uintptr_t)y & 15) != 0) return D3D_ERR_UNSUPPORTED;
  if (R == 0) return 0;
  if (N / 4 > 256) return D3D_ERR_UNSUPPORTED;
  const int lanes = 256 / (N / 4);
  const int blocks = (int)min((R + lanes - 1) / lanes, (long long)148 * 8);
  const size_t smem = (size_t)N * (kMaxSmall + 1) * sizeof(float);
  linear_small_k_kernel<<<blocks, 256, smem, (cudaStream_t)stream>>>(x, w, w_stride_n, w_stride_k, bias, R, K, N, y);
  d3d_note_launches(1);
  return d3d_launch_status();
}

int d3d_linear_small_n(const float* x, const float* w, const float* bias, long long R, int K, int N, float* y, void* stream) {
  D3D_REQUIRE(x && w && y && R >= 0 && K > 0 && K % 4 == 0 && N > 0 && N <= kMaxSmall);
  if ((((uintptr_t)x | (uintptr_t)w) & 15) != 0) return D3D_ERR_UNSUPPORTED;
  if (R == 0) return 0;
  const int blocks = (int)min((R + kSnRows - 1) / kSnRows, (long long)148 * 4);
  const size_t smem = ((size_t)kSnRows * (K + 1) + (size_t)N * K) * sizeof(float);
  if (smem > 200 * 1024) return D3D_ERR_UNSUPPORTED;  // K beyond ~750: not a shape of this path
  cudaError_t e = cudaFuncSetAttribute(linear_small_n_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  linear_small_n_kernel<<<blocks, 256, smem, (cudaStream_t)stream>>>(x, w, bias, R, K, N, y);
  d3d_note_launches(1);
  return d3d_launch_status();
}

size_t d3d_wgrad_small_workspace_bytes(int Nb) { return Nb > 0 ? (size_t)kWgradBlocks * Nb * kMaxSmall * sizeof(float) : 0; }

/* out[n * out_stride_n + k * out_stride_k] (+)= sum_r big[r, n] * small[r, k];  big (R, Nb) with Nb % 4 == 0, small (R, Ks), Ks <= 4 */
int d3d_wgrad_small(const float* big, const float* small, long long R, int Nb, int Ks, float* out, long long out_stride_n,
                    long long out_stride_k, int accumulate, void* ws, size_t ws_bytes, void* stream) {
  D3D_REQUIRE(big && small && out && R > 0 && Nb > 0 && Nb % 4 == 0 && Nb <= 1024 && Ks > 0 && Ks <= kMaxSmall);
  if (((uintptr_t)big & 15) != 0) return D3D_ERR_UNSUPPORTED;
  if (!ws || ws_bytes < d3d_wgrad_small_workspace_bytes(Nb)) return D3D_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const int groups = Nb / 4;
  if (groups > 256) return D3D_ERR_UNSUPPORTED;
  const int lanes = 256 / groups;
  long long nblk = min((long long)kWgradBlocks, (R + 4LL * lanes - 1) / (4LL * lanes));
  const long long rows_per_block = (R + nblk - 1) / nblk;
  nblk = (R + rows_per_block - 1) / rows_per_block;
  const size_t smem = (size_t)lanes * Nb * kMaxSmall * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(wgrad_small_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  wgrad_small_partial_kernel<<<(unsigned)nblk, 256, smem, st>>>(big, small, R, Nb, Ks, rows_per_block, (float*)ws);
  wgrad_small_reduce_kernel<<<d3d_ceil_div(Nb * Ks, 256), 256, 0, st>>>((const float*)ws, (int)nblk, Nb, Ks, out_stride_n,
                                                                      out_stride_k, accumulate, out);
  d3d_note_launches(2);
  return d3d_launch_status();
}

}  // extern "C"

// Fused BatchNorm1d (+ residual add) (+ ReLU) on channel-major (B, C, N) activations, training and eval mode.
//
// the 1x1 conv / BN / ReLU sandwich around every local aggregation
//   ref: u_net_arch/models/backbones/resnet.py:32-45, 58-66  (Conv1d -> BatchNorm1d -> ReLU; conv2 -> BN, + identity, ReLU)
//   ref: u_net_arch/models/local_aggregation_operators.py:121-123 (out_transform = BatchNorm1d + ReLU)
// The reference runs these as separate cuDNN / ATen kernels: BN forward (2 reads + 1 write of the activation),
// ReLU (read + write), and in backward ReLU-grad (2 reads + write) + BN-grad (cuDNN bn_bw: ~1.1 TB/s measured on
// B200 for (16,144,8192)).  Here: statistics pass (1 read) + apply pass (1 read, 1 write, ReLU and the residual add
// folded in); backward: reduction pass (2 reads) + apply pass (2 reads, 1 write).  All passes are HBM-bound
// streaming kernels with float4 accesses.
//
// Semantics = torch.nn.BatchNorm1d: biased variance for normalisation, unbiased for running_var, momentum update,
// eps inside the sqrt.  Accumulation: per-thread fp32 partial sums of the SHIFTED data (x - x[0] of the channel),
// combined in fp64 — stable for |mean| >> std.
#include "common.cuh"

namespace {

constexpr int kStatThreads = 512;

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(D3D_FULL_MASK, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double t = lane < (blockDim.x >> 5) ? red[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(D3D_FULL_MASK, t, o);
    if (lane == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}

// one block per channel: mean / inverse std over (B, N), running-stat update
__global__ void __launch_bounds__(kStatThreads)
bn_stats_kernel(const float* __restrict__ x, int B, int C, int N, float eps, float momentum,
                float* __restrict__ running_mean, float* __restrict__ running_var, float* __restrict__ save_mean,
                float* __restrict__ save_invstd) {
  __shared__ double red[32];
  const int c = blockIdx.x;
  const float shift = x[(size_t)c * N];
  float s1 = 0.f, s2 = 0.f;
  const int n4 = (N % 4 == 0) ? N / 4 : 0;
  for (int b = 0; b < B; ++b) {
    const float* row = x + ((size_t)b * C + c) * N;
    if (n4) {
      const float4* r4 = reinterpret_cast<const float4*>(row);
      for (int i = threadIdx.x; i < n4; i += kStatThreads) {
        const float4 v = __ldg(r4 + i);
        const float a0 = v.x - shift, a1 = v.y - shift, a2 = v.z - shift, a3 = v.w - shift;
        s1 += (a0 + a1) + (a2 + a3);
        s2 += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
      }
    } else {
      for (int i = threadIdx.x; i < N; i += kStatThreads) {
        const float a = row[i] - shift;
        s1 += a;
        s2 += a * a;
      }
    }
  }
  const double t1 = block_sum((double)s1, red);
  const double t2 = block_sum((double)s2, red);
  if (threadIdx.x == 0) {
    const double n = (double)B * N;
    const double m = t1 / n;                       // mean of the shifted data
    double var = t2 / n - m * m;                   // biased variance
    if (var < 0) var = 0;
    const float mean = (float)(m + (double)shift);
    save_mean[c] = mean;
    save_invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) {
      const double unbiased = n > 1 ? var * n / (n - 1) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
  }
}

// y = act((x - mean) * invstd * gamma + beta [+ residual]);  grid (chunks of N, B*C)
template <bool kVec>
__global__ void __launch_bounds__(256)
bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ residual, const float* __restrict__ gamma,
                const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ invstd,
                int C, int N, int relu, float* __restrict__ y) {
  const int bc = blockIdx.y, c = bc % C;
  const float scale = invstd[c] * (gamma ? gamma[c] : 1.f);
  const float offset = (beta ? beta[c] : 0.f) - mean[c] * scale;
  const size_t base = (size_t)bc * N;
  if (kVec) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i * 4 >= N) return;
    float4 v = __ldg(reinterpret_cast<const float4*>(x + base) + i);
    v.x = v.x * scale + offset; v.y = v.y * scale + offset; v.z = v.z * scale + offset; v.w = v.w * scale + offset;
    if (residual) {
      const float4 r = __ldg(reinterpret_cast<const float4*>(residual + base) + i);
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    reinterpret_cast<float4*>(y + base)[i] = v;
  } else {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    float v = x[base + i] * scale + offset;
    if (residual) v += residual[base + i];
    y[base + i] = relu ? fmaxf(v, 0.f) : v;
  }
}

// one block per channel: sum(dyr), sum(dyr * xhat) with dyr = dy masked by the ReLU; dgamma / dbeta out
__global__ void __launch_bounds__(kStatThreads)
bn_bwd_reduce_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ y,
                     const float* __restrict__ mean, const float* __restrict__ invstd, int B, int C, int N, int relu,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ sums /* (C, 2) */) {
  __shared__ double red[32];
  const int c = blockIdx.x;
  const float m = mean[c], is = invstd[c];
  float s1 = 0.f, s2 = 0.f;
  const int n4 = (N % 4 == 0) ? N / 4 : 0;
  for (int b = 0; b < B; ++b) {
    const size_t base = ((size_t)b * C + c) * N;
    if (n4) {
      const float4* d4 = reinterpret_cast<const float4*>(dy + base);
      const float4* x4 = reinterpret_cast<const float4*>(x + base);
      const float4* y4 = reinterpret_cast<const float4*>(y + base);
      for (int i = threadIdx.x; i < n4; i += kStatThreads) {
        float4 d = __ldg(d4 + i);
        const float4 xv = __ldg(x4 + i);
        if (relu) {
          const float4 yv = __ldg(y4 + i);
          d.x = yv.x > 0.f ? d.x : 0.f; d.y = yv.y > 0.f ? d.y : 0.f; d.z = yv.z > 0.f ? d.z : 0.f; d.w = yv.w > 0.f ? d.w : 0.f;
        }
        s1 += (d.x + d.y) + (d.z + d.w);
        s2 += (d.x * (xv.x - m) + d.y * (xv.y - m)) + (d.z * (xv.z - m) + d.w * (xv.w - m));
      }
    } else {
      for (int i = threadIdx.x; i < N; i += kStatThreads) {
        float d = dy[base + i];
        if (relu && !(y[base + i] > 0.f)) d = 0.f;
        s1 += d;
        s2 += d * (x[base + i] - m);
      }
    }
  }
  const double t1 = block_sum((double)s1, red);
  const double t2 = block_sum((double)s2, red) * (double)is;  // sum(dyr * xhat)
  if (threadIdx.x == 0) {
    if (dbeta) dbeta[c] = (float)t1;
    if (dgamma) dgamma[c] = (float)t2;
    sums[2 * c] = (float)t1;
    sums[2 * c + 1] = (float)t2;
  }
}

// dx = gamma * invstd * (dyr - sum_dy / n - xhat * sum_dy_xhat / n);  dres = dyr  (training mode)
// eval mode (use_batch_stats == 0): dx = gamma * invstd * dyr
template <bool kVec>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ y,
                    const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ invstd,
                    const float* __restrict__ sums, int C, int N, float inv_count, int relu, int use_batch_stats,
                    float* __restrict__ dx, float* __restrict__ dres) {
  const int bc = blockIdx.y, c = bc % C;
  const float m = mean[c], is = invstd[c];
  const float g = (gamma ? gamma[c] : 1.f) * is;
  const float k1 = use_batch_stats ? sums[2 * c] * inv_count : 0.f;
  const float k2 = use_batch_stats ? sums[2 * c + 1] * inv_count * is : 0.f;  // multiplies (x - mean)
  const size_t base = (size_t)bc * N;
  if (kVec) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i * 4 >= N) return;
    float4 d = __ldg(reinterpret_cast<const float4*>(dy + base) + i);
    const float4 xv = __ldg(reinterpret_cast<const float4*>(x + base) + i);
    if (relu) {
      const float4 yv = __ldg(reinterpret_cast<const float4*>(y + base) + i);
      d.x = yv.x > 0.f ? d.x : 0.f; d.y = yv.y > 0.f ? d.y : 0.f; d.z = yv.z > 0.f ? d.z : 0.f; d.w = yv.w > 0.f ? d.w : 0.f;
    }
    if (dres) reinterpret_cast<float4*>(dres + base)[i] = d;
    float4 o;
    o.x = g * (d.x - k1 - (xv.x - m) * k2); o.y = g * (d.y - k1 - (xv.y - m) * k2);
    o.z = g * (d.z - k1 - (xv.z - m) * k2); o.w = g * (d.w - k1 - (xv.w - m) * k2);
    reinterpret_cast<float4*>(dx + base)[i] = o;
  } else {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    float d = dy[base + i];
    if (relu && !(y[base + i] > 0.f)) d = 0.f;
    if (dres) dres[base + i] = d;
    dx[base + i] = g * (d - k1 - (x[base + i] - m) * k2);
  }
}

__global__ void bn_eval_stats_kernel(const float* __restrict__ running_mean, const float* __restrict__ running_var, int C,
                                     float eps, float* __restrict__ mean, float* __restrict__ invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  mean[c] = running_mean[c];
  invstd[c] = rsqrtf(running_var[c] + eps);
}

bool vec_ok(int N, const void* a, const void* b, const void* c, const void* d) {
  auto al = [](const void* p) { return p == nullptr || ((uintptr_t)p & 15) == 0; };
  return N % 4 == 0 && al(a) && al(b) && al(c) && al(d);
}

}  // namespace

extern "C" {

int d3d_bn_act_fwd(const float* x, const float* residual, const float* gamma, const float* beta, float* running_mean,
                   float* running_var, int B, int C, int N, float eps, float momentum, int training, int relu, float* y,
                   float* save_mean, float* save_invstd, void* stream) {
  D3D_REQUIRE(x && y && save_mean && save_invstd);
  D3D_REQUIRE(B > 0 && C > 0 && N > 0);
  D3D_REQUIRE(training || (running_mean && running_var));
  cudaStream_t st = (cudaStream_t)stream;
  if (training)
    bn_stats_kernel<<<C, kStatThreads, 0, st>>>(x, B, C, N, eps, momentum, running_mean, running_var, save_mean, save_invstd);
  else
    bn_eval_stats_kernel<<<d3d_ceil_div(C, 256), 256, 0, st>>>(running_mean, running_var, C, eps, save_mean, save_invstd);
  if (vec_ok(N, x, residual, y, nullptr)) {
    dim3 grid(d3d_ceil_div(N / 4, 256), B * C);
    bn_apply_kernel<true><<<grid, 256, 0, st>>>(x, residual, gamma, beta, save_mean, save_invstd, C, N, relu, y);
  } else {
    dim3 grid(d3d_ceil_div(N, 256), B * C);
    bn_apply_kernel<false><<<grid, 256, 0, st>>>(x, residual, gamma, beta, save_mean, save_invstd, C, N, relu, y);
  }
  d3d_note_launches(2);
  return d3d_launch_status();
}

size_t d3d_bn_act_bwd_workspace_bytes(int C) { return C > 0 ? (size_t)C * 2 * sizeof(float) : 0; }

int d3d_bn_act_bwd(const float* dy, const float* x, const float* y, const float* gamma, const float* save_mean,
                   const float* save_invstd, int B, int C, int N, int training, int relu, float* dx, float* dres,
                   float* dgamma, float* dbeta, void* ws, size_t ws_bytes, void* stream) {
  D3D_REQUIRE(dy && x && save_mean && save_invstd && dx);
  D3D_REQUIRE(B > 0 && C > 0 && N > 0);
  D3D_REQUIRE(!relu || y);
  if (!ws || ws_bytes < d3d_bn_act_bwd_workspace_bytes(C)) return D3D_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  float* sums = (float*)ws;
  bn_bwd_reduce_kernel<<<C, kStatThreads, 0, st>>>(dy, x, y, save_mean, save_invstd, B, C, N, relu, dgamma, dbeta, sums);
  const float inv_count = 1.0f / ((float)B * (float)N);
  if (vec_ok(N, dy, x, y, dx) && vec_ok(N, dres, nullptr, nullptr, nullptr)) {
    dim3 grid(d3d_ceil_div(N / 4, 256), B * C);
    bn_bwd_apply_kernel<true><<<grid, 256, 0, st>>>(dy, x, y, gamma, save_mean, save_invstd, sums, C, N, inv_count, relu,
                                                    training, dx, dres);
  } else {
    dim3 grid(d3d_ceil_div(N, 256), B * C);
    bn_bwd_apply_kernel<false><<<grid, 256, 0, st>>>(dy, x, y, gamma, save_mean, save_invstd, sums, C, N, inv_count, relu,
                                                     training, dx, dres);
  }
  d3d_note_launches(2);
  return d3d_launch_status();
}

}  // extern "C"

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

// Channel reductions are split over S blocks per channel (grid (C, S)) so that even C = 72 fills the 148 SMs; every
// block writes its fp64 partials, the LAST block of a channel to finish (ticket counter) adds the S partials in
// fixed order — deterministic — and finalises.  The counters are zeroed by the finaliser for the next launch.
__device__ __forceinline__ bool last_block_of_channel(unsigned* counters, int c, int S) {
  __shared__ bool is_last;
  __threadfence();  // this block's partials are visible before the ticket is taken
  if (threadIdx.x == 0) {
    const unsigned ticket = atomicAdd(&counters[c], 1u);
    is_last = (ticket == (unsigned)S - 1u);
    if (is_last) counters[c] = 0u;
  }
  __syncthreads();
  if (is_last) __threadfence();
  return is_last;
}

// mean / inverse std over (B, N), running-stat update.  Rows (b, c, :) are dealt round-robin to the S blocks in
// segments of N / seg_per_row elements.
__global__ void __launch_bounds__(kStatThreads)
bn_stats_kernel(const float* __restrict__ x, int B, int C, int N, int S, float eps, float momentum,
                float* __restrict__ running_mean, float* __restrict__ running_var, float* __restrict__ save_mean,
                float* __restrict__ save_invstd, double* __restrict__ partials, unsigned* __restrict__ counters) {
  __shared__ double red[32];
  const int c = blockIdx.x, sblk = blockIdx.y;
  const float shift = x[(size_t)c * N];
  float s1 = 0.f, s2 = 0.f;
  const bool vec = (N % 4 == 0);
  // work units: (b, chunk) with chunks of kStatThreads*4 elements
  const int chunk = kStatThreads * 4;
  const int chunks_per_row = (N + chunk - 1) / chunk;
  const int units = B * chunks_per_row;
  for (int u = sblk; u < units; u += S) {
    const int b = u / chunks_per_row, ch = u - b * chunks_per_row;
    const float* row = x + ((size_t)b * C + c) * N;
    const int i0 = ch * chunk + threadIdx.x * 4;
    if (vec) {
      if (i0 < N) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(row + i0));
        const float a0 = v.x - shift, a1 = v.y - shift, a2 = v.z - shift, a3 = v.w - shift;
        s1 += (a0 + a1) + (a2 + a3);
        s2 += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
      }
    } else {
      for (int i = i0; i < min(i0 + 4, N); ++i) {
        const float a = row[i] - shift;
        s1 += a;
        s2 += a * a;
      }
    }
  }
  const double t1 = block_sum((double)s1, red);
  const double t2 = block_sum((double)s2, red);
  if (threadIdx.x == 0) {
    partials[((size_t)c * S + sblk) * 2] = t1;
    partials[((size_t)c * S + sblk) * 2 + 1] = t2;
  }
  if (!last_block_of_channel(counters, c, S)) return;
  if (threadIdx.x == 0) {
    double a1 = 0.0, a2 = 0.0;
    for (int k = 0; k < S; ++k) { a1 += partials[((size_t)c * S + k) * 2]; a2 += partials[((size_t)c * S + k) * 2 + 1]; }
    const double n = (double)B * N;
    const double m = a1 / n;                       // mean of the shifted data
    double var = a2 / n - m * m;                   // biased variance
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

// Small activations (B*N <= kSmallCount elements per channel, the deep U-Net levels with hundreds to thousands of
// channels): one WARP per channel, shuffle reduction only — no block barrier, no partials, no ticket.
constexpr int kSmallCount = 2048;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(D3D_FULL_MASK, v, o);
  return v;
}

__global__ void __launch_bounds__(256)
bn_stats_small_kernel(const float* __restrict__ x, int B, int C, int N, float eps, float momentum,
                      float* __restrict__ running_mean, float* __restrict__ running_var, float* __restrict__ save_mean,
                      float* __restrict__ save_invstd) {
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= C) return;
  const float shift = x[(size_t)c * N];
  float s1 = 0.f, s2 = 0.f;
  const bool vec = (N % 4 == 0);
  for (int b = 0; b < B; ++b) {
    const float* row = x + ((size_t)b * C + c) * N;
    if (vec) {
      for (int i = lane * 4; i < N; i += 128) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(row + i));
        const float a0 = v.x - shift, a1 = v.y - shift, a2 = v.z - shift, a3 = v.w - shift;
        s1 += (a0 + a1) + (a2 + a3);
        s2 += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
      }
    } else {
      for (int i = lane; i < N; i += 32) {
        const float a = row[i] - shift;
        s1 += a;
        s2 += a * a;
      }
    }
  }
  const double a1 = warp_sum((double)s1), a2 = warp_sum((double)s2);
  if (lane == 0) {
    const double n = (double)B * N;
    const double m = a1 / n;
    double var = a2 / n - m * m;
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

__global__ void __launch_bounds__(256)
bn_bwd_reduce_small_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ y,
                           const float* __restrict__ gamma, const float* __restrict__ beta,
                           const float* __restrict__ mean, const float* __restrict__ invstd, int B, int C, int N, int relu,
                           float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ sums) {
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= C) return;
  const float m = mean[c], is = invstd[c];
  const float scale = is * (gamma ? gamma[c] : 1.f);
  const float offset = (beta ? beta[c] : 0.f) - m * scale;
  float s1 = 0.f, s2 = 0.f;
  if (N % 4 == 0) {
    const int n4 = N / 4;
    for (int u = lane; u < B * n4; u += 32) {
      const int b = u / n4, i = u - b * n4;
      const size_t base = ((size_t)b * C + c) * N + 4 * i;
      float4 d = __ldg(reinterpret_cast<const float4*>(dy + base));
      const float4 xv = __ldg(reinterpret_cast<const float4*>(x + base));
      if (relu == 2) {
        const float4 yv = __ldg(reinterpret_cast<const float4*>(y + base));
        d.x = yv.x > 0.f ? d.x : 0.f; d.y = yv.y > 0.f ? d.y : 0.f; d.z = yv.z > 0.f ? d.z : 0.f; d.w = yv.w > 0.f ? d.w : 0.f;
      } else if (relu == 1) {
        d.x = xv.x * scale + offset > 0.f ? d.x : 0.f; d.y = xv.y * scale + offset > 0.f ? d.y : 0.f;
        d.z = xv.z * scale + offset > 0.f ? d.z : 0.f; d.w = xv.w * scale + offset > 0.f ? d.w : 0.f;
      }
      s1 += (d.x + d.y) + (d.z + d.w);
      s2 += (d.x * (xv.x - m) + d.y * (xv.y - m)) + (d.z * (xv.z - m) + d.w * (xv.w - m));
    }
  } else {
    for (int b = 0; b < B; ++b) {
      const size_t base = ((size_t)b * C + c) * N;
      for (int i = lane; i < N; i += 32) {
        float d = dy[base + i];
        const float xv = x[base + i];
        if (relu == 2 && !(y[base + i] > 0.f)) d = 0.f;
        if (relu == 1 && !(xv * scale + offset > 0.f)) d = 0.f;
        s1 += d;
        s2 += d * (xv - m);
      }
    }
  }
  const double a1 = warp_sum((double)s1), a2 = warp_sum((double)s2) * (double)is;
  if (lane == 0) {
    if (dbeta) dbeta[c] = (float)a1;
    if (dgamma) dgamma[c] = (float)a2;
    sums[2 * c] = (float)a1;
    sums[2 * c + 1] = (float)a2;
  }
}

// y = act((x - mean) * invstd * gamma + beta [+ residual]);  grid (chunks of N, B*C)
template <bool kVec>
__global__ void __launch_bounds__(256)
bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ residual, const float* __restrict__ gamma,
                const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ invstd,
                int C, int N, long long total_rows, int relu, float* __restrict__ y) {
  // flattened over (b*C + c, element): small N (deep levels) still gives full blocks
  const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
  const int per_row = kVec ? N / 4 : N;
  const long long bc_ll = e / per_row;
  if (bc_ll >= total_rows) return;
  const int bc = (int)bc_ll, c = bc % C;
  const int i = (int)(e - bc_ll * per_row);
  const float scale = invstd[c] * (gamma ? gamma[c] : 1.f);
  const float offset = (beta ? beta[c] : 0.f) - mean[c] * scale;
  const size_t base = (size_t)bc * N;
  if (kVec) {
    float4 v = __ldg(reinterpret_cast<const float4*>(x + base) + i);
    v.x = v.x * scale + offset; v.y = v.y * scale + offset; v.z = v.z * scale + offset; v.w = v.w * scale + offset;
    if (residual) {
      const float4 r = __ldg(reinterpret_cast<const float4*>(residual + base) + i);
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    reinterpret_cast<float4*>(y + base)[i] = v;
  } else {
    float v = x[base + i] * scale + offset;
    if (residual) v += residual[base + i];
    y[base + i] = relu ? fmaxf(v, 0.f) : v;
  }
}

// sum(dyr), sum(dyr * xhat) with dyr = dy masked by the ReLU; dgamma / dbeta out.
// relu: 0 = none, 1 = mask recomputed from x (y = relu(x * scale + offset), no residual), 2 = mask read from y
__global__ void __launch_bounds__(kStatThreads)
bn_bwd_reduce_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ y,
                     const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean,
                     const float* __restrict__ invstd, int B, int C, int N, int S, int relu, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, float* __restrict__ sums /* (C, 2) */, double* __restrict__ partials,
                     unsigned* __restrict__ counters) {
  __shared__ double red[32];
  const int c = blockIdx.x, sblk = blockIdx.y;
  const float m = mean[c], is = invstd[c];
  const float scale = is * (gamma ? gamma[c] : 1.f);
  const float offset = (beta ? beta[c] : 0.f) - m * scale;
  float s1 = 0.f, s2 = 0.f;
  const bool vec = (N % 4 == 0);
  const int chunk = kStatThreads * 4;
  const int chunks_per_row = (N + chunk - 1) / chunk;
  const int units = B * chunks_per_row;
  for (int u = sblk; u < units; u += S) {
    const int b = u / chunks_per_row, ch = u - b * chunks_per_row;
    const size_t base = ((size_t)b * C + c) * N;
    const int i0 = ch * chunk + threadIdx.x * 4;
    if (vec) {
      if (i0 < N) {
        float4 d = __ldg(reinterpret_cast<const float4*>(dy + base + i0));
        const float4 xv = __ldg(reinterpret_cast<const float4*>(x + base + i0));
        if (relu == 2) {
          const float4 yv = __ldg(reinterpret_cast<const float4*>(y + base + i0));
          d.x = yv.x > 0.f ? d.x : 0.f; d.y = yv.y > 0.f ? d.y : 0.f; d.z = yv.z > 0.f ? d.z : 0.f; d.w = yv.w > 0.f ? d.w : 0.f;
        } else if (relu == 1) {  // same expression as bn_apply_kernel: bit-identical mask
          d.x = xv.x * scale + offset > 0.f ? d.x : 0.f; d.y = xv.y * scale + offset > 0.f ? d.y : 0.f;
          d.z = xv.z * scale + offset > 0.f ? d.z : 0.f; d.w = xv.w * scale + offset > 0.f ? d.w : 0.f;
        }
        s1 += (d.x + d.y) + (d.z + d.w);
        s2 += (d.x * (xv.x - m) + d.y * (xv.y - m)) + (d.z * (xv.z - m) + d.w * (xv.w - m));
      }
    } else {
      for (int i = i0; i < min(i0 + 4, N); ++i) {
        float d = dy[base + i];
        const float xv = x[base + i];
        if (relu == 2 && !(y[base + i] > 0.f)) d = 0.f;
        if (relu == 1 && !(xv * scale + offset > 0.f)) d = 0.f;
        s1 += d;
        s2 += d * (xv - m);
      }
    }
  }
  const double t1 = block_sum((double)s1, red);
  const double t2 = block_sum((double)s2, red);
  if (threadIdx.x == 0) {
    partials[((size_t)c * S + sblk) * 2] = t1;
    partials[((size_t)c * S + sblk) * 2 + 1] = t2;
  }
  if (!last_block_of_channel(counters, c, S)) return;
  if (threadIdx.x == 0) {
    double a1 = 0.0, a2 = 0.0;
    for (int k = 0; k < S; ++k) { a1 += partials[((size_t)c * S + k) * 2]; a2 += partials[((size_t)c * S + k) * 2 + 1]; }
    a2 *= (double)is;  // sum(dyr * xhat)
    if (dbeta) dbeta[c] = (float)a1;
    if (dgamma) dgamma[c] = (float)a2;
    sums[2 * c] = (float)a1;
    sums[2 * c + 1] = (float)a2;
  }
}

// dx = gamma * invstd * (dyr - sum_dy / n - xhat * sum_dy_xhat / n);  dres = dyr  (training mode)
// eval mode (use_batch_stats == 0): dx = gamma * invstd * dyr
template <bool kVec>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ y,
                    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean,
                    const float* __restrict__ invstd, const float* __restrict__ sums, int C, int N, long long total_rows,
                    float inv_count, int relu, int use_batch_stats, float* __restrict__ dx, float* __restrict__ dres) {
  const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
  const int per_row = kVec ? N / 4 : N;
  const long long bc_ll = e / per_row;
  if (bc_ll >= total_rows) return;
  const int bc = (int)bc_ll, c = bc % C;
  const int i = (int)(e - bc_ll * per_row);
  const float m = mean[c], is = invstd[c];
  const float g = (gamma ? gamma[c] : 1.f) * is;
  const float offset = (beta ? beta[c] : 0.f) - m * g;  // y = relu(x * g + offset) when there is no residual
  const float k1 = use_batch_stats ? sums[2 * c] * inv_count : 0.f;
  const float k2 = use_batch_stats ? sums[2 * c + 1] * inv_count * is : 0.f;  // multiplies (x - mean)
  const size_t base = (size_t)bc * N;
  if (kVec) {
    float4 d = __ldg(reinterpret_cast<const float4*>(dy + base) + i);
    const float4 xv = __ldg(reinterpret_cast<const float4*>(x + base) + i);
    if (relu == 2) {
      const float4 yv = __ldg(reinterpret_cast<const float4*>(y + base) + i);
      d.x = yv.x > 0.f ? d.x : 0.f; d.y = yv.y > 0.f ? d.y : 0.f; d.z = yv.z > 0.f ? d.z : 0.f; d.w = yv.w > 0.f ? d.w : 0.f;
    } else if (relu == 1) {
      d.x = xv.x * g + offset > 0.f ? d.x : 0.f; d.y = xv.y * g + offset > 0.f ? d.y : 0.f;
      d.z = xv.z * g + offset > 0.f ? d.z : 0.f; d.w = xv.w * g + offset > 0.f ? d.w : 0.f;
    }
    if (dres) reinterpret_cast<float4*>(dres + base)[i] = d;
    float4 o;
    o.x = g * (d.x - k1 - (xv.x - m) * k2); o.y = g * (d.y - k1 - (xv.y - m) * k2);
    o.z = g * (d.z - k1 - (xv.z - m) * k2); o.w = g * (d.w - k1 - (xv.w - m) * k2);
    reinterpret_cast<float4*>(dx + base)[i] = o;
  } else {
    float d = dy[base + i];
    if (relu == 2 && !(y[base + i] > 0.f)) d = 0.f;
    if (relu == 1 && !(x[base + i] * g + offset > 0.f)) d = 0.f;
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



int stat_splits(int C) {  // blocks per channel so that the grid covers the machine ~3x
  int S = (148 * 3 + C - 1) / C;
  if (S < 1) S = 1;
  if (S > 64) S = 64;
  return S;
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

extern "C" {

/* workspace layout (both directions): [sums: 2*C float][partials: C*S*2 double][counters: C unsigned, must be ZERO
 * before the first use; the kernels leave them zero] */
size_t d3d_bn_act_workspace_bytes(int C) {
  if (C <= 0) return 0;
  return align256((size_t)C * 2 * sizeof(float)) + align256((size_t)C * 64 * 2 * sizeof(double)) + align256((size_t)C * sizeof(unsigned));
}

static void carve(void* ws, int C, float** sums, double** partials, unsigned** counters) {
  unsigned char* p = (unsigned char*)ws;
  *sums = (float*)p; p += align256((size_t)C * 2 * sizeof(float));
  *partials = (double*)p; p += align256((size_t)C * 64 * 2 * sizeof(double));
  *counters = (unsigned*)p;
}

int d3d_bn_act_fwd(const float* x, const float* residual, const float* gamma, const float* beta, float* running_mean,
                   float* running_var, int B, int C, int N, float eps, float momentum, int training, int relu, float* y,
                   float* save_mean, float* save_invstd, void* ws, size_t ws_bytes, void* stream) {
  D3D_REQUIRE(x && y && save_mean && save_invstd);
  D3D_REQUIRE(B > 0 && C > 0 && N > 0);
  D3D_REQUIRE(training || (running_mean && running_var));
  cudaStream_t st = (cudaStream_t)stream;
  if (training) {
    if (!ws || ws_bytes < d3d_bn_act_workspace_bytes(C)) return D3D_ERR_WORKSPACE;
    float* sums; double* partials; unsigned* counters;
    carve(ws, C, &sums, &partials, &counters);
    if ((long long)B * N <= kSmallCount) {
      bn_stats_small_kernel<<<d3d_ceil_div(C, 8), 256, 0, st>>>(x, B, C, N, eps, momentum, running_mean, running_var,
                                                                save_mean, save_invstd);
    } else {
      const int S = stat_splits(C);
      bn_stats_kernel<<<dim3(C, S), kStatThreads, 0, st>>>(x, B, C, N, S, eps, momentum, running_mean, running_var,
                                                           save_mean, save_invstd, partials, counters);
    }
  } else {
    bn_eval_stats_kernel<<<d3d_ceil_div(C, 256), 256, 0, st>>>(running_mean, running_var, C, eps, save_mean, save_invstd);
  }
  const long long rows = (long long)B * C;
  if (vec_ok(N, x, residual, y, nullptr)) {
    bn_apply_kernel<true><<<d3d_ceil_div(rows * (N / 4), 256), 256, 0, st>>>(x, residual, gamma, beta, save_mean, save_invstd,
                                                                            C, N, rows, relu, y);
  } else {
    bn_apply_kernel<false><<<d3d_ceil_div(rows * N, 256), 256, 0, st>>>(x, residual, gamma, beta, save_mean, save_invstd, C, N,
                                                                       rows, relu, y);
  }
  d3d_note_launches(2);
  return d3d_launch_status();
}

/* relu: 0 = none, 1 = ReLU without residual (mask recomputed from x: y may be NULL), 2 = mask read from y */
int d3d_bn_act_bwd(const float* dy, const float* x, const float* y, const float* gamma, const float* beta,
                   const float* save_mean, const float* save_invstd, int B, int C, int N, int training, int relu,
                   float* dx, float* dres, float* dgamma, float* dbeta, void* ws, size_t ws_bytes, void* stream) {
  D3D_REQUIRE(dy && x && save_mean && save_invstd && dx);
  D3D_REQUIRE(B > 0 && C > 0 && N > 0 && relu >= 0 && relu <= 2);
  D3D_REQUIRE(relu != 2 || y);
  if (!ws || ws_bytes < d3d_bn_act_workspace_bytes(C)) return D3D_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  float* sums; double* partials; unsigned* counters;
  carve(ws, C, &sums, &partials, &counters);
  if ((long long)B * N <= kSmallCount) {
    bn_bwd_reduce_small_kernel<<<d3d_ceil_div(C, 8), 256, 0, st>>>(dy, x, y, gamma, beta, save_mean, save_invstd, B, C, N,
                                                                   relu, dgamma, dbeta, sums);
  } else {
    const int S = stat_splits(C);
    bn_bwd_reduce_kernel<<<dim3(C, S), kStatThreads, 0, st>>>(dy, x, y, gamma, beta, save_mean, save_invstd, B, C, N, S,
                                                              relu, dgamma, dbeta, sums, partials, counters);
  }
  const float inv_count = 1.0f / ((float)B * (float)N);
  const long long rows = (long long)B * C;
  if (vec_ok(N, dy, x, y, dx) && vec_ok(N, dres, nullptr, nullptr, nullptr)) {
    bn_bwd_apply_kernel<true><<<d3d_ceil_div(rows * (N / 4), 256), 256, 0, st>>>(dy, x, y, gamma, beta, save_mean, save_invstd,
                                                                                sums, C, N, rows, inv_count, relu, training,
                                                                                dx, dres);
  } else {
    bn_bwd_apply_kernel<false><<<d3d_ceil_div(rows * N, 256), 256, 0, st>>>(dy, x, y, gamma, beta, save_mean, save_invstd, sums,
                                                                           C, N, rows, inv_count, relu, training, dx, dres);
  }
  d3d_note_launches(2);
  return d3d_launch_status();
}

}  // extern "C"

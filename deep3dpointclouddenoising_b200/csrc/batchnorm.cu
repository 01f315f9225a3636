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
  __syncthreads();  // every partial of this block is written ...
  if (threadIdx.x == 0) {
    __threadfence();  // ... and (fences are cumulative) visible device-wide before the ticket is taken
    const unsigned ticket = atomicAdd(&counters[c], 1u);
    is_last = (ticket == (unsigned)S - 1u);
    if (is_last) {
      counters[c] = 0u;
      __threadfence();
    }
  }
  __syncthreads();
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
    for (int k = 0; k < S; ++k) { a1 += __ldcg(&partials[((size_t)c * S + k) * 2]); a2 += __ldcg(&partials[((size_t)c * S + k) * 2 + 1]); }
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
    for (int k = 0; k < S; ++k) { a1 += __ldcg(&partials[((size_t)c * S + k) * 2]); a2 += __ldcg(&partials[((size_t)c * S + k) * 2 + 1]); }
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

// ---------------------------------------------------------------------------------------------------------------
// Channel-last variants: the activation is (R, C) row-major with R = B*N rows — the layout the local aggregation
// kernels gather from and write.  With the 1x1 convolutions run as (R, Cin) x (Cin, Cout) GEMMs the whole network
// stays in this layout and no transposition kernel is needed between the convolutions and the aggregations.
// A thread owns one float4 of channels; a block covers up to kClGroups float4 columns x (256 / groups) rows per
// iteration.
constexpr int kClThreads = 256;
constexpr int kClGroups = 8;  // float4 channel groups per block at most: 128-byte row segments.  Narrow slabs keep
                              // the number of blocks per channel — hence the partials the finaliser reads — small

struct ClGeom {
  int c4, groups, rows_per_iter, gx;
};
__host__ __device__ inline ClGeom cl_geom(int C) {
  ClGeom g;
  g.c4 = C / 4;
  g.groups = g.c4 < kClGroups ? g.c4 : kClGroups;
  for (int d = kClGroups; d >= 4; --d)  // prefer a width that divides the channel count: no idle threads
    if (g.c4 % d == 0) { g.groups = d; break; }
  g.rows_per_iter = kClThreads / g.groups;
  g.gx = (g.c4 + g.groups - 1) / g.groups;
  return g;
}

// Tree reduction over the rows of the block: sh[0..7][row * groups + cx] -> row 0 (value-major: conflict-free).
// Fixed shape, so deterministic.
__device__ __forceinline__ void cl_tree(double (*sh)[kClThreads], int cx, int ry, bool active, const ClGeom g) {
  for (int n = g.rows_per_iter; n > 1;) {
    const int h = (n + 1) >> 1;
    if (active && ry + h < n) {
#pragma unroll
      for (int j = 0; j < 8; ++j) sh[j][threadIdx.x] += sh[j][threadIdx.x + h * g.groups];
    }
    __syncthreads();
    n = h;
  }
}

// Sums the per-thread float4 pairs over the rows of the block, stores the block's fp64 partials; in the LAST block of
// this channel slab all threads then add the S partials (thread (cx, ry) takes k = ry, ry + rows, ...; fixed order, so
// the result is deterministic) and the row-0 threads return true with the totals in tot1 / tot2.
__device__ __forceinline__ bool cl_combine(const float4 s1, const float4 s2, int cx, int ry, int group, bool active,
                                           const ClGeom g, int C, int S, double* __restrict__ partials,
                                           unsigned* __restrict__ counters, double tot1[4], double tot2[4]) {
  __shared__ double sh[8][kClThreads];
  const float f[8] = {s1.x, s1.y, s1.z, s1.w, s2.x, s2.y, s2.z, s2.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) sh[j][threadIdx.x] = (double)f[j];
  __syncthreads();
  cl_tree(sh, cx, ry, active, g);
  const bool owner = active && ry == 0;
  if (S == 1) {  // one block per slab: nothing to combine across blocks
    if (owner) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { tot1[j] = sh[j][threadIdx.x]; tot2[j] = sh[4 + j][threadIdx.x]; }
    }
    return owner;
  }
  if (owner) {
    // partials: [slab][S][column in slab][8]  -> the finaliser reads them coalesced
    double* dst = partials + (((size_t)blockIdx.x * S + blockIdx.y) * g.groups + cx) * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[j] = sh[j][threadIdx.x];
  }
  if (!last_block_of_channel(counters, blockIdx.x, S)) return false;
  double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (active) {
    for (int k = ry; k < S; k += g.rows_per_iter) {  // a handful of iterations: S <= ~150 per slab
      const double* src = partials + (((size_t)blockIdx.x * S + k) * g.groups + cx) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += __ldcg(src + j);  // L2: written by other SMs in this launch
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sh[j][threadIdx.x] = a[j];
  __syncthreads();
  cl_tree(sh, cx, ry, active, g);
  if (!owner) return false;
#pragma unroll
  for (int j = 0; j < 4; ++j) { tot1[j] = sh[j][threadIdx.x]; tot2[j] = sh[4 + j][threadIdx.x]; }
  return true;
}

__global__ void __launch_bounds__(kClThreads)
bn_stats_cl_kernel(const float* __restrict__ x, long long R, int C, int S, float eps, float momentum,
                   float* __restrict__ running_mean, float* __restrict__ running_var, long long* __restrict__ num_batches,
                   float* __restrict__ save_mean, float* __restrict__ save_invstd, double* __restrict__ partials,
                   unsigned* __restrict__ counters) {
  const ClGeom g = cl_geom(C);
  const int cx = threadIdx.x % g.groups, ry = threadIdx.x / g.groups;
  const int group = blockIdx.x * g.groups + cx;
  const bool active = ry < g.rows_per_iter && group < g.c4;
  float4 s1 = make_float4(0, 0, 0, 0), s2 = s1, shift = s1;
  if (active) {
    shift = __ldg(reinterpret_cast<const float4*>(x) + group);  // row 0
    // loads are issued in batches of 8 before any is consumed (a plain unrolled loop keeps its exit test between
    // consecutive loads, which leaves one load in flight per thread)
    const long long r0 = (long long)blockIdx.y * g.rows_per_iter + ry;
    const long long step = (long long)S * g.rows_per_iter;
    const int iters = r0 < R ? (int)((R - r0 + step - 1) / step) : 0;
    const float4* p = reinterpret_cast<const float4*>(x) + r0 * g.c4 + group;
    const long long pstep = step * g.c4;
    auto acc = [&](const float4 v) {
      const float a0 = v.x - shift.x, a1 = v.y - shift.y, a2 = v.z - shift.z, a3 = v.w - shift.w;
      s1.x += a0; s1.y += a1; s1.z += a2; s1.w += a3;
      s2.x += a0 * a0; s2.y += a1 * a1; s2.z += a2 * a2; s2.w += a3 * a3;
    };
    for (int it = 0; it < iters; it += 8, p += 8 * pstep) {  // the tail batch is predicated, not serialised
      float4 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (it + j < iters) v[j] = __ldg(p + j * pstep);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (it + j < iters) acc(v[j]);
    }
  }
  double t1[4], t2[4];
  if (!cl_combine(s1, s2, cx, ry, group, active, g, C, S, partials, counters, t1, t2)) return;
  if (group == 0 && num_batches) *num_batches += 1;  // BatchNorm1d.num_batches_tracked: one thread per launch gets here
  const float sh[4] = {shift.x, shift.y, shift.z, shift.w};
  const double n = (double)R;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = group * 4 + j;
    const double m = t1[j] / n;
    double var = t2[j] / n - m * m;
    if (var < 0) var = 0;
    const float mean = (float)(m + (double)sh[j]);
    save_mean[c] = mean;
    save_invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) {
      const double unbiased = n > 1 ? var * n / (n - 1) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
  }
}

__device__ __forceinline__ void cl_scale_offset(const float* __restrict__ gamma, const float* __restrict__ beta,
                                                const float* __restrict__ mean, const float* __restrict__ invstd,
                                                int group, float4& m, float4& is, float4& scale, float4& offset) {
  m = __ldg(reinterpret_cast<const float4*>(mean) + group);
  is = __ldg(reinterpret_cast<const float4*>(invstd) + group);
  const float4 gm = gamma ? __ldg(reinterpret_cast<const float4*>(gamma) + group) : make_float4(1, 1, 1, 1);
  const float4 bt = beta ? __ldg(reinterpret_cast<const float4*>(beta) + group) : make_float4(0, 0, 0, 0);
  scale = make_float4(is.x * gm.x, is.y * gm.y, is.z * gm.z, is.w * gm.w);
  offset = make_float4(bt.x - m.x * scale.x, bt.y - m.y * scale.y, bt.z - m.z * scale.z, bt.w - m.w * scale.w);
}

// The elementwise passes give every thread kApplyRows consecutive rows of ONE float4 channel group: the per-channel
// parameters (4 to 6 float4 loads) are fetched once per thread instead of once per 16 payload bytes — those loads hit L1
// but still cost LSU write-back passes — and kApplyRows independent payload loads are in flight per thread.
constexpr int kApplyRows = 4;

__global__ void __launch_bounds__(256)
bn_apply_cl_kernel(const float* __restrict__ x, const float* __restrict__ residual, const float* __restrict__ gamma,
                   const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ invstd,
                   int c4, long long R, int relu, float* __restrict__ y) {
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long rb = t / c4;
  const int group = (int)(t - rb * c4);
  const long long row0 = rb * kApplyRows;
  if (row0 >= R) return;
  float4 m, is, scale, offset;
  cl_scale_offset(gamma, beta, mean, invstd, group, m, is, scale, offset);
  float4 v[kApplyRows], r[kApplyRows];
#pragma unroll
  for (int k = 0; k < kApplyRows; ++k)
    if (row0 + k < R) {
      const long long e = (row0 + k) * c4 + group;
      v[k] = __ldg(reinterpret_cast<const float4*>(x) + e);
      if (residual) r[k] = __ldg(reinterpret_cast<const float4*>(residual) + e);
    }
#pragma unroll
  for (int k = 0; k < kApplyRows; ++k)
    if (row0 + k < R) {
      float4 o = v[k];
      o.x = o.x * scale.x + offset.x; o.y = o.y * scale.y + offset.y;
      o.z = o.z * scale.z + offset.z; o.w = o.w * scale.w + offset.w;
      if (residual) { o.x += r[k].x; o.y += r[k].y; o.z += r[k].z; o.w += r[k].w; }
      if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
      reinterpret_cast<float4*>(y)[(row0 + k) * c4 + group] = o;
    }
}

// dy masked by the ReLU, same expressions as the forward pass so the mask is bit-identical
__device__ __forceinline__ float4 cl_masked_dy(float4 d, const float4 xv, const float* __restrict__ y, long long e, int relu,
                                               const float4 scale, const float4 offset) {
  if (relu == 2) {
    const float4 yv = __ldg(reinterpret_cast<const float4*>(y) + e);
    d.x = yv.x > 0.f ? d.x : 0.f; d.y = yv.y > 0.f ? d.y : 0.f; d.z = yv.z > 0.f ? d.z : 0.f; d.w = yv.w > 0.f ? d.w : 0.f;
  } else if (relu == 1) {
    d.x = xv.x * scale.x + offset.x > 0.f ? d.x : 0.f; d.y = xv.y * scale.y + offset.y > 0.f ? d.y : 0.f;
    d.z = xv.z * scale.z + offset.z > 0.f ? d.z : 0.f; d.w = xv.w * scale.w + offset.w > 0.f ? d.w : 0.f;
  }
  return d;
}

__global__ void __launch_bounds__(kClThreads)
bn_bwd_reduce_cl_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ y,
                        const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean,
                        const float* __restrict__ invstd, long long R, int C, int S, int relu, int accumulate,
                        float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ sums,
                        double* __restrict__ partials, unsigned* __restrict__ counters) {
  const ClGeom g = cl_geom(C);
  const int cx = threadIdx.x % g.groups, ry = threadIdx.x / g.groups;
  const int group = blockIdx.x * g.groups + cx;
  const bool active = ry < g.rows_per_iter && group < g.c4;
  float4 s1 = make_float4(0, 0, 0, 0), s2 = s1, m = s1, is = s1, scale = s1, offset = s1;
  if (active) {
    cl_scale_offset(gamma, beta, mean, invstd, group, m, is, scale, offset);
    const long long r0 = (long long)blockIdx.y * g.rows_per_iter + ry;
    const long long step = (long long)S * g.rows_per_iter;
    const int iters = r0 < R ? (int)((R - r0 + step - 1) / step) : 0;
    long long e = r0 * g.c4 + group;
    const long long estep = step * g.c4;
    const float4* dy4 = reinterpret_cast<const float4*>(dy);
    const float4* x4 = reinterpret_cast<const float4*>(x);
    auto acc = [&](const float4 dv, const float4 xv, long long at) {
      const float4 d = cl_masked_dy(dv, xv, y, at, relu, scale, offset);
      s1.x += d.x; s1.y += d.y; s1.z += d.z; s1.w += d.w;
      s2.x += d.x * (xv.x - m.x); s2.y += d.y * (xv.y - m.y); s2.z += d.z * (xv.z - m.z); s2.w += d.w * (xv.w - m.w);
    };
    for (int it = 0; it < iters; it += 4, e += 4 * estep) {  // 8 independent loads in flight; predicated tail
      float4 dv[4], xv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (it + j < iters) {
          dv[j] = __ldg(dy4 + e + j * estep);
          xv[j] = __ldg(x4 + e + j * estep);
        }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (it + j < iters) acc(dv[j], xv[j], e + j * estep);
    }
  }
  double t1[4], t2[4];
  if (!cl_combine(s1, s2, cx, ry, group, active, g, C, S, partials, counters, t1, t2)) return;
  const float isv[4] = {is.x, is.y, is.z, is.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = group * 4 + j;
    const double a2 = t2[j] * (double)isv[j];
    // accumulate != 0: dgamma / dbeta ARE the parameters' gradient buffers (+=, one thread per channel: no race)
    if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + (float)t1[j];
    if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + (float)a2;
    sums[2 * c] = (float)t1[j];
    sums[2 * c + 1] = (float)a2;
  }
}


// L2 residency hints for the two-kernel backward: the reduction pass asks the L2 to KEEP what it reads (evict_last), the
// apply pass reads the same rows a few microseconds later and marks them evict_first (their last use).
__device__ __forceinline__ unsigned long long l2_policy_keep() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long l2_policy_drop() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float4 ldg_hint(const float4* ptr, unsigned long long policy) {
  float4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(ptr), "l"(policy));
  return v;
}

__global__ void __launch_bounds__(256)
bn_bwd_apply_cl_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ y,
                       const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean,
                       const float* __restrict__ invstd, const float* __restrict__ sums, int c4, long long R,
                       float inv_count, int relu, int use_batch_stats, float* __restrict__ dx, float* __restrict__ dres) {
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long rb = t / c4;
  const int group = (int)(t - rb * c4);
  const long long row0 = rb * kApplyRows;
  if (row0 >= R) return;
  float4 m, is, g, offset;
  cl_scale_offset(gamma, beta, mean, invstd, group, m, is, g, offset);
  float4 k1 = make_float4(0, 0, 0, 0), k2 = k1;
  if (use_batch_stats) {
    const float4 sa = __ldg(reinterpret_cast<const float4*>(sums) + 2 * group);      // (s1, s2) of channels 0, 1
    const float4 sb = __ldg(reinterpret_cast<const float4*>(sums) + 2 * group + 1);  // channels 2, 3
    k1 = make_float4(sa.x * inv_count, sa.z * inv_count, sb.x * inv_count, sb.z * inv_count);
    k2 = make_float4(sa.y * inv_count * is.x, sa.w * inv_count * is.y, sb.y * inv_count * is.z, sb.w * inv_count * is.w);
  }
  float4 xv[kApplyRows], dv[kApplyRows];
  const unsigned long long drop = l2_policy_drop();
#pragma unroll
  for (int k = 0; k < kApplyRows; ++k)
    if (row0 + k < R) {
      const long long e = (row0 + k) * c4 + group;
      xv[k] = ldg_hint(reinterpret_cast<const float4*>(x) + e, drop);
      dv[k] = ldg_hint(reinterpret_cast<const float4*>(dy) + e, drop);
    }
#pragma unroll
  for (int k = 0; k < kApplyRows; ++k)
    if (row0 + k < R) {
      const long long e = (row0 + k) * c4 + group;
      const float4 d = cl_masked_dy(dv[k], xv[k], y, e, relu, g, offset);
      if (dres) reinterpret_cast<float4*>(dres)[e] = d;
      float4 o;
      o.x = g.x * (d.x - k1.x - (xv[k].x - m.x) * k2.x); o.y = g.y * (d.y - k1.y - (xv[k].y - m.y) * k2.y);
      o.z = g.z * (d.z - k1.z - (xv[k].z - m.z) * k2.z); o.w = g.w * (d.w - k1.w - (xv[k].w - m.w) * k2.w);
      reinterpret_cast<float4*>(dx)[e] = o;
    }
}

// ---- backward reduction, second generation: row-range partials + a parallel finaliser (no tickets, no serial tail) ----
// A block streams a contiguous range of rows; a thread keeps ONE float4 channel group (block size = a multiple of the
// group count, so consecutive threads read consecutive 16-byte pieces of the row-major matrix — fully coalesced) and
// accumulates its rows in registers; the row lanes of the block are combined in shared memory and the block writes one
// fp32 partial per channel.  bn_bwd_finalize_kernel adds the <= 592 partials per channel in fp64 in a fixed order.
constexpr int kRedBlocks = 148 * 4;

struct RedGeom {
  int c4, lanes, threads;
};
__host__ __device__ inline RedGeom red_geom(int C) {
  RedGeom g;
  g.c4 = C / 4;
  g.lanes = 256 / g.c4 > 0 ? 256 / g.c4 : 1;
  g.threads = g.c4 * g.lanes;
  return g;
}

__global__ void __launch_bounds__(640)
bn_bwd_partial_cl_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ y,
                         const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean,
                         const float* __restrict__ invstd, long long R, int c4, int lanes, long long rows_per_block, int relu,
                         float* __restrict__ partial /* (gridDim.x, 2, 4 * c4) */) {
  extern __shared__ float4 red_sh[];  // [2][lanes][c4]
  const int group = threadIdx.x % c4, rl = threadIdx.x / c4;
  const long long row_beg = (long long)blockIdx.x * rows_per_block;
  const long long row_end = row_beg + rows_per_block < R ? row_beg + rows_per_block : R;
  float4 m, is, scale, offset;
  cl_scale_offset(gamma, beta, mean, invstd, group, m, is, scale, offset);
  float4 s1 = make_float4(0, 0, 0, 0), s2 = s1;
  const float4* dy4 = reinterpret_cast<const float4*>(dy);
  const float4* x4 = reinterpret_cast<const float4*>(x);
  const unsigned long long keep = l2_policy_keep();
  for (long long r = row_beg + rl; r < row_end; r += 4LL * lanes) {  // 8 independent loads in flight; predicated tail
    float4 dv[4], xv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (r + (long long)j * lanes < row_end) {
        const long long e = (r + (long long)j * lanes) * c4 + group;
        dv[j] = ldg_hint(dy4 + e, keep);
        xv[j] = ldg_hint(x4 + e, keep);
      }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (r + (long long)j * lanes < row_end) {
        const long long e = (r + (long long)j * lanes) * c4 + group;
        const float4 d = cl_masked_dy(dv[j], xv[j], y, e, relu, scale, offset);
        s1.x += d.x; s1.y += d.y; s1.z += d.z; s1.w += d.w;
        s2.x += d.x * (xv[j].x - m.x); s2.y += d.y * (xv[j].y - m.y); s2.z += d.z * (xv[j].z - m.z); s2.w += d.w * (xv[j].w - m.w);
      }
  }
  red_sh[rl * c4 + group] = s1;
  red_sh[(lanes + rl) * c4 + group] = s2;
  __syncthreads();
  if (rl == 0) {
    for (int k = 1; k < lanes; ++k) {  // fixed order: deterministic
      const float4 a = red_sh[k * c4 + group], b = red_sh[(lanes + k) * c4 + group];
      s1.x += a.x; s1.y += a.y; s1.z += a.z; s1.w += a.w;
      s2.x += b.x; s2.y += b.y; s2.z += b.z; s2.w += b.w;
    }
    float4* dst = reinterpret_cast<float4*>(partial) + (size_t)blockIdx.x * 2 * c4;
    dst[group] = s1;
    dst[c4 + group] = s2;
  }
}

constexpr int kRedFinCh = 8, kRedFinSlices = 32;

__global__ void __launch_bounds__(kRedFinCh * kRedFinSlices)
bn_bwd_finalize_kernel(const float* __restrict__ partial, int nblk, int C, const float* __restrict__ invstd, int accumulate,
                       float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ sums) {
  __shared__ double red[2][kRedFinSlices][kRedFinCh];
  const int cx = threadIdx.x % kRedFinCh, ty = threadIdx.x / kRedFinCh;
  const int c = min(blockIdx.x * kRedFinCh + cx, C - 1);
  double a1 = 0.0, a2 = 0.0;
  for (int t = ty; t < nblk; t += kRedFinSlices) {
    a1 += (double)partial[(size_t)t * 2 * C + c];
    a2 += (double)partial[(size_t)t * 2 * C + C + c];
  }
  red[0][ty][cx] = a1;
  red[1][ty][cx] = a2;
  __syncthreads();
  if (ty != 0 || blockIdx.x * kRedFinCh + cx >= C) return;
  double t1 = 0.0, t2 = 0.0;
#pragma unroll
  for (int k = 0; k < kRedFinSlices; ++k) { t1 += red[0][k][cx]; t2 += red[1][k][cx]; }
  const double g2 = t2 * (double)invstd[c];
  // accumulate != 0: dgamma / dbeta ARE the parameters' gradient buffers (+=, one thread per channel: no race)
  if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + (float)t1;
  if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + (float)g2;
  sums[2 * c] = (float)t1;
  sums[2 * c + 1] = (float)g2;
}

constexpr int kClBlocks = 148 * 4;  // four 8-warp blocks per SM, 8 independent 16-byte loads in flight per thread;
                                    // more splits only lengthen the finaliser's chain of L2 round trips

int cl_splits(const ClGeom g, long long R) {
  long long S = (kClBlocks + g.gx - 1) / g.gx;
  const long long max_s = (R + 4LL * g.rows_per_iter - 1) / (4LL * g.rows_per_iter);  // >= 4 rows per thread
  if (S > max_s) S = max_s;
  if (R * g.groups * 16 <= 300 * 1024) S = 1;  // a slab a single block streams in ~2 us: skip partials and ticket
  if (S < 1) S = 1;
  return (int)S;
}

bool aligned16(const void* p) { return p == nullptr || ((uintptr_t)p & 15) == 0; }

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
static size_t partial_doubles(int C) {
  // channel-major kernels: C * 64 splits * 2; channel-last kernels: slabs * S * groups * 8 with slabs * S <= 148 * 4 + slabs
  const size_t cm = (size_t)C * 64 * 2;
  const size_t cl = (size_t)(148 * 4 + (C + 15) / 16) * kClGroups * 8;
  return cm > cl ? cm : cl;
}

size_t d3d_bn_act_workspace_bytes(int C) {
  if (C <= 0) return 0;
  return align256((size_t)C * 2 * sizeof(float)) + align256(partial_doubles(C) * sizeof(double)) +
         align256((size_t)C * sizeof(unsigned)) + align256((size_t)kRedBlocks * 2 * C * sizeof(float));
}

static float* row_partials(void* ws, int C) {  // after [sums][partials][counters]: (kRedBlocks, 2, C) fp32
  return (float*)((unsigned char*)ws + align256((size_t)C * 2 * sizeof(float)) + align256(partial_doubles(C) * sizeof(double)) +
                  align256((size_t)C * sizeof(unsigned)));
}

static void carve(void* ws, int C, float** sums, double** partials, unsigned** counters) {
  unsigned char* p = (unsigned char*)ws;
  *sums = (float*)p; p += align256((size_t)C * 2 * sizeof(float));
  *partials = (double*)p; p += align256(partial_doubles(C) * sizeof(double));
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

/* Channel-last variants: x, residual, y, dy, dx, dres are (R, C) row-major with R = B * N rows.  C % 4 == 0 and
 * 16-byte aligned pointers are required (D3D_ERR_ARG otherwise: the caller falls back to the channel-major entry). */
int d3d_bn_act_cl_fwd(const float* x, const float* residual, const float* gamma, const float* beta, float* running_mean,
                      float* running_var, long long* num_batches_tracked, long long R, int C, float eps, float momentum,
                      int training, int relu, float* y, float* save_mean, float* save_invstd, void* ws, size_t ws_bytes,
                      void* stream) {
  D3D_REQUIRE(x && y && save_mean && save_invstd);
  D3D_REQUIRE(R > 0 && C > 0 && C % 4 == 0);
  D3D_REQUIRE(aligned16(x) && aligned16(residual) && aligned16(y) && aligned16(gamma) && aligned16(beta) &&
              aligned16(save_mean) && aligned16(save_invstd));
  D3D_REQUIRE(training || (running_mean && running_var));
  cudaStream_t st = (cudaStream_t)stream;
  const ClGeom g = cl_geom(C);
  if (training) {
    if (!ws || ws_bytes < d3d_bn_act_workspace_bytes(C)) return D3D_ERR_WORKSPACE;
    float* sums; double* partials; unsigned* counters;
    carve(ws, C, &sums, &partials, &counters);
    const int S = cl_splits(g, R);
    bn_stats_cl_kernel<<<dim3(g.gx, S), kClThreads, 0, st>>>(x, R, C, S, eps, momentum, running_mean, running_var,
                                                             num_batches_tracked, save_mean, save_invstd, partials, counters);
  } else {
    bn_eval_stats_kernel<<<d3d_ceil_div(C, 256), 256, 0, st>>>(running_mean, running_var, C, eps, save_mean, save_invstd);
  }
  const long long apply_threads = ((R + kApplyRows - 1) / kApplyRows) * g.c4;
  bn_apply_cl_kernel<<<(unsigned)d3d_ceil_div(apply_threads, 256), 256, 0, st>>>(x, residual, gamma, beta, save_mean, save_invstd,
                                                                                g.c4, R, relu, y);
  d3d_note_launches(2);
  return d3d_launch_status();
}

/* Apply pass alone with statistics that already exist (d3d_bn_finalize from the GEMM epilogue's tile partials):
 * y = act(gamma * (x - mean) * invstd + beta [+ residual]). */
int d3d_bn_apply_cl(const float* x, const float* residual, const float* gamma, const float* beta, const float* save_mean,
                    const float* save_invstd, long long R, int C, int relu, float* y, void* stream) {
  D3D_REQUIRE(x && y && save_mean && save_invstd);
  D3D_REQUIRE(R > 0 && C > 0 && C % 4 == 0);
  D3D_REQUIRE(aligned16(x) && aligned16(residual) && aligned16(y) && aligned16(gamma) && aligned16(beta) &&
              aligned16(save_mean) && aligned16(save_invstd));
  const ClGeom g = cl_geom(C);
  const long long apply_threads = ((R + kApplyRows - 1) / kApplyRows) * g.c4;
  bn_apply_cl_kernel<<<(unsigned)d3d_ceil_div(apply_threads, 256), 256, 0, (cudaStream_t)stream>>>(
      x, residual, gamma, beta, save_mean, save_invstd, g.c4, R, relu, y);
  d3d_note_launches(1);
  return d3d_launch_status();
}

int d3d_bn_act_cl_bwd(const float* dy, const float* x, const float* y, const float* gamma, const float* beta,
                      const float* save_mean, const float* save_invstd, long long R, int C, int training, int relu, float* dx,
                      float* dres, float* dgamma, float* dbeta, int accumulate_param_grads, void* ws, size_t ws_bytes,
                      void* stream) {
  D3D_REQUIRE(dy && x && save_mean && save_invstd && dx);
  D3D_REQUIRE(R > 0 && C > 0 && C % 4 == 0 && relu >= 0 && relu <= 2);
  D3D_REQUIRE(relu != 2 || y);
  D3D_REQUIRE(aligned16(dy) && aligned16(x) && aligned16(y) && aligned16(dx) && aligned16(dres) && aligned16(gamma) &&
              aligned16(beta) && aligned16(save_mean) && aligned16(save_invstd));
  if (!ws || ws_bytes < d3d_bn_act_workspace_bytes(C)) return D3D_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  float* sums; double* partials; unsigned* counters;
  carve(ws, C, &sums, &partials, &counters);
  const ClGeom g = cl_geom(C);
  const RedGeom rg = red_geom(C);
  if (rg.threads <= 640) {
    long long nblk = (R + 4LL * rg.lanes - 1) / (4LL * rg.lanes);  // at least four rows per thread
    if (nblk > kRedBlocks) nblk = kRedBlocks;
    const long long rows_per_block = (R + nblk - 1) / nblk;
    nblk = (R + rows_per_block - 1) / rows_per_block;
    float* part = row_partials(ws, C);
    bn_bwd_partial_cl_kernel<<<(unsigned)nblk, rg.threads, (size_t)rg.threads * 32, st>>>(dy, x, y, gamma, beta, save_mean,
                                                                                       save_invstd, R, rg.c4, rg.lanes,
                                                                                       rows_per_block, relu, part);
    bn_bwd_finalize_kernel<<<d3d_ceil_div(C, kRedFinCh), kRedFinCh * kRedFinSlices, 0, st>>>(part, (int)nblk, C, save_invstd,
                                                                                          accumulate_param_grads, dgamma,
                                                                                          dbeta, sums);
    d3d_note_launches(1);
  } else {
    const int S = cl_splits(g, R);
    bn_bwd_reduce_cl_kernel<<<dim3(g.gx, S), kClThreads, 0, st>>>(dy, x, y, gamma, beta, save_mean, save_invstd, R, C, S, relu,
                                                                  accumulate_param_grads, dgamma, dbeta, sums, partials,
                                                                  counters);
  }
  const long long apply_threads = ((R + kApplyRows - 1) / kApplyRows) * g.c4;
  bn_bwd_apply_cl_kernel<<<(unsigned)d3d_ceil_div(apply_threads, 256), 256, 0, st>>>(dy, x, y, gamma, beta, save_mean, save_invstd,
                                                                                    sums, g.c4, R, 1.0f / (float)R, relu, training,
                                                                                    dx, dres);
  d3d_note_launches(2);
  return d3d_launch_status();
}

}  // extern "C"

// PseudoGrid (depthwise KPConv-style) aggregation, fp32 CUDA-core path, forward and backward.
//
//   ref: u_net_arch/models/local_aggregation_operators.py:467-503
//     w[b,j,k,m]  = influence(|| (S[idx[b,j,m]] - Q[b,j]) - K[k] ||) * fm[b,j,m]
//     out[b,j,c]  = sum_k W[k,c] * sum_m w[b,j,k,m] * F[b, idx[b,j,m], c]
// The reference materialises (B,M,ns,K,3) twice plus the gathered features and calls a batched
// [K x ns].[ns x C] SGEMM per point.  Here the sums are re-associated as
//     out[b,j,c]  = sum_m F[b,idx,c] * E[m,c],    E[m,c] = sum_k w[j,k,m] * W[k,c]
// so a lane keeps its W[:, c..c+3] column block in registers, streams the neighbours' rows once
// (coalesced float4 loads, channel-last) and needs K FMAs per gathered element.  (pseudogrid_tc.cu
// moves the E = w.W product to tcgen05 tensor cores.)
// Backward:
//     dF[b,i,c]   = sum over inverse-map entries (j,m) of i:  g[b,j,c] * E[(j,m),c]     (segmented, ordered)
//     dW[k,c]     = sum_{b,j,m} w[b,j,k,m] * F[b,idx,c] * g[b,j,c]    (per-warp registers -> per-block
//                   partials -> ordered final reduction: deterministic two-pass, no atomics)
#include "common.cuh"

namespace {

constexpr int kWarps = 8;
constexpr int kK = 16;  // kernel points padded to 16 (weights of the padding are zero)

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float influence_weight(float dx, float dy, float dz, float extent, int influence) {
  const float sq = dx * dx + dy * dy + dz * dz;
  if (influence == D3D_KP_LINEAR) return fmaxf(1.0f - sqrtf(sq) / extent, 0.0f);  // :480
  if (influence == D3D_KP_GAUSSIAN) {
    const float sigma = extent * 0.3f;  // :484, models/utlis.py:287-294
    return expf(-sq / (2.0f * sigma * sigma + 1e-9f));
  }
  return 1.0f;  // constant (:476)
}

// per-warp staging of one query's slots: sidx[m], sw[m][0..16)
__device__ __forceinline__ int stage_query(const float* __restrict__ query_xyz, const float* __restrict__ support_xyz,
                                           const int* __restrict__ idx, const int* __restrict__ nvalid,
                                           const int* __restrict__ query_mask, const float* __restrict__ kp_smem,
                                           int b, int j, int M, int N, int nsample, int K, float extent, int influence,
                                           int lane, float* sw, int* sidx) {
  const size_t qrow = (size_t)b * M + j;
  const int n_eff = query_mask[qrow] != 0 ? nvalid[qrow] : nsample;
  const float qx = query_xyz[qrow * 3], qy = query_xyz[qrow * 3 + 1], qz = query_xyz[qrow * 3 + 2];
  const int* irow = idx + qrow * nsample;
  for (int m = lane; m < n_eff; m += 32) sidx[m] = d3d_clamp_index(irow[m], N);
  __syncwarp();
  for (int t = lane; t < n_eff * kK; t += 32) {
    const int m = t >> 4, k = t & 15;
    float w = 0.0f;
    if (k < K) {
      const float* s = support_xyz + ((size_t)b * N + sidx[m]) * 3;
      w = influence_weight((s[0] - qx) - kp_smem[3 * k], (s[1] - qy) - kp_smem[3 * k + 1], (s[2] - qz) - kp_smem[3 * k + 2],
                           extent, influence);
    }
    sw[t] = w;
  }
  __syncwarp();
  return n_eff;
}

__global__ void __launch_bounds__(kWarps * 32)
pseudogrid_fwd_kernel(const float* __restrict__ feat, const float* __restrict__ query_xyz,
                      const float* __restrict__ support_xyz, const int* __restrict__ idx,
                      const int* __restrict__ nvalid, const int* __restrict__ query_mask,
                      const float* __restrict__ kpoints, const float* __restrict__ weights, int M, int N, int C,
                      int nsample, int K, float extent, int influence, float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float kp[kK * 3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < K * 3) kp[threadIdx.x] = kpoints[threadIdx.x];
  __syncthreads();
  const int b = blockIdx.y;
  const int j = blockIdx.x * kWarps + warp;
  if (j >= M) return;
  float* sw = reinterpret_cast<float*>(smem_raw) + (size_t)warp * nsample * kK;
  int* sidx = reinterpret_cast<int*>(reinterpret_cast<float*>(smem_raw) + (size_t)kWarps * nsample * kK) + (size_t)warp * nsample;
  const int n_eff = stage_query(query_xyz, support_xyz, idx, nvalid, query_mask, kp, b, j, M, N, nsample, K, extent,
                                influence, lane, sw, sidx);
  const int cv = C >> 2;
  const float* fb = feat + (size_t)b * N * C;
  float* orow = out + ((size_t)b * M + j) * C;
  for (int q = lane; q < cv; q += 32) {
    float4 W[kK];
#pragma unroll
    for (int k = 0; k < kK; ++k) W[k] = k < K ? ld4(weights + (size_t)k * C + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
    for (int m = 0; m < n_eff; ++m) {
      const float4 x = ld4(fb + (size_t)sidx[m] * C + 4 * q);
      const float4* wm = reinterpret_cast<const float4*>(sw + m * kK);
      float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        const float4 w = wm[k4];
        const float ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float4 Wk = W[4 * k4 + r];
          e.x += ws[r] * Wk.x; e.y += ws[r] * Wk.y; e.z += ws[r] * Wk.z; e.w += ws[r] * Wk.w;
        }
      }
      acc.x += x.x * e.x; acc.y += x.y * e.y; acc.z += x.z * e.z; acc.w += x.w * e.w;
    }
    *reinterpret_cast<float4*>(orow + 4 * q) = acc;
  }
}

// dF: warp per support point over its inverse-map segment
__global__ void __launch_bounds__(kWarps * 32)
pseudogrid_bwd_feat_kernel(const float* __restrict__ grad_out, const float* __restrict__ query_xyz,
                           const float* __restrict__ support_xyz, const int* __restrict__ rowptr,
                           const int* __restrict__ entries, const int* __restrict__ nvalid,
                           const int* __restrict__ query_mask, const float* __restrict__ kpoints,
                           const float* __restrict__ weights, int M, int N, int C, int nsample, int K, float extent,
                           int influence, float* __restrict__ grad_feat) {
  __shared__ float kp[kK * 3];
  __shared__ __align__(16) float stage_w[kWarps][32][kK];
  __shared__ int stage_j[kWarps][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < K * 3) kp[threadIdx.x] = kpoints[threadIdx.x];
  __syncthreads();
  const int b = blockIdx.y;
  const int i = blockIdx.x * kWarps + warp;
  if (i >= N) return;
  const size_t srow = (size_t)b * N + i;
  const int beg = rowptr[srow], end = rowptr[srow + 1];
  const float sx = support_xyz[srow * 3], sy = support_xyz[srow * 3 + 1], sz = support_xyz[srow * 3 + 2];
  const int cv = C >> 2;
  const float* gb = grad_out + (size_t)b * M * C;
  float* orow = grad_feat + srow * C;
  for (int q0 = 0; q0 < cv; q0 += 32) {
    const int q = q0 + lane;
    const bool active = q < cv;
    float4 W[kK];
#pragma unroll
    for (int k = 0; k < kK; ++k)
      W[k] = (active && k < K) ? ld4(weights + (size_t)k * C + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e0 = beg; e0 < end; e0 += 32) {
      const int n_here = min(32, end - e0);
      __syncwarp();
      if (lane < n_here) {
        const int packed = entries[e0 + lane];
        int j = packed >> 8;
        const int m = packed & 255;
        const size_t qrow = (size_t)b * M + j;
        const int n_eff = query_mask[qrow] != 0 ? nvalid[qrow] : nsample;
        if (m < n_eff) {
          const float dx = sx - query_xyz[qrow * 3], dy = sy - query_xyz[qrow * 3 + 1], dz = sz - query_xyz[qrow * 3 + 2];
#pragma unroll
          for (int k = 0; k < kK; ++k)
            stage_w[warp][lane][k] =
                k < K ? influence_weight(dx - kp[3 * k], dy - kp[3 * k + 1], dz - kp[3 * k + 2], extent, influence) : 0.0f;
        } else {
          j = -1;
        }
        stage_j[warp][lane] = j;
      }
      __syncwarp();
      for (int t = 0; t < n_here; ++t) {
        const int j = stage_j[warp][t];
        if (j < 0 || !active) continue;
        const float4 g = ld4(gb + (size_t)j * C + 4 * q);
        const float4* wm = reinterpret_cast<const float4*>(stage_w[warp][t]);
        float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const float4 w = wm[k4];
          const float ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const float4 Wk = W[4 * k4 + r];
            e.x += ws[r] * Wk.x; e.y += ws[r] * Wk.y; e.z += ws[r] * Wk.z; e.w += ws[r] * Wk.w;
          }
        }
        acc.x += g.x * e.x; acc.y += g.y * e.y; acc.z += g.z * e.z; acc.w += g.w * e.w;
      }
    }
    if (active) *reinterpret_cast<float4*>(orow + 4 * q) = acc;
  }
}

// dW partials: grid (nblk, ceil(cv/32)); every warp walks queries  qi = warp_global, warp_global + stride, ...
__global__ void __launch_bounds__(kWarps * 32)
pseudogrid_bwd_weight_kernel(const float* __restrict__ grad_out, const float* __restrict__ feat,
                             const float* __restrict__ query_xyz, const float* __restrict__ support_xyz,
                             const int* __restrict__ idx, const int* __restrict__ nvalid,
                             const int* __restrict__ query_mask, const float* __restrict__ kpoints, int B, int M,
                             int N, int C, int nsample, int K, float extent, int influence,
                             float* __restrict__ partial /* (gridDim.x, kK, C) */) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float kp[kK * 3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < K * 3) kp[threadIdx.x] = kpoints[threadIdx.x];
  __syncthreads();
  float* sw = reinterpret_cast<float*>(smem_raw) + (size_t)warp * nsample * kK;
  int* sidx = reinterpret_cast<int*>(reinterpret_cast<float*>(smem_raw) + (size_t)kWarps * nsample * kK) + (size_t)warp * nsample;
  const int cv = C >> 2;
  const int q = blockIdx.y * 32 + lane;
  const bool active = q < cv;
  float4 acc[kK];
#pragma unroll
  for (int k = 0; k < kK; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long total = (long long)B * M;
  for (long long qi = (long long)blockIdx.x * kWarps + warp; qi < total; qi += (long long)gridDim.x * kWarps) {
    const int b = (int)(qi / M), j = (int)(qi - (long long)b * M);
    const int n_eff = stage_query(query_xyz, support_xyz, idx, nvalid, query_mask, kp, b, j, M, N, nsample, K, extent,
                                  influence, lane, sw, sidx);
    if (active) {
      const float4 g = ld4(grad_out + (size_t)qi * C + 4 * q);
      const float* fb = feat + (size_t)b * N * C + 4 * q;
#pragma unroll 2
      for (int m = 0; m < n_eff; ++m) {
        const float4 x = ld4(fb + (size_t)sidx[m] * C);
        const float4 xg = make_float4(x.x * g.x, x.y * g.y, x.z * g.z, x.w * g.w);
        const float4* wm = reinterpret_cast<const float4*>(sw + m * kK);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const float4 w = wm[k4];
          const float ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            float4& a = acc[4 * k4 + r];
            a.x += ws[r] * xg.x; a.y += ws[r] * xg.y; a.z += ws[r] * xg.z; a.w += ws[r] * xg.w;
          }
        }
      }
    }
    __syncwarp();  // staging buffers are reused by the next query
  }
  // block reduction in fixed warp order through shared memory (reuse the staging area)
  __syncthreads();
  float4* red = reinterpret_cast<float4*>(smem_raw);  // [kWarps][kK][32] float4 = 64 KB
  // two rounds of 8 kernel points keep the buffer at 32 KB
  for (int half = 0; half < 2; ++half) {
#pragma unroll
    for (int k = 0; k < 8; ++k) red[((size_t)warp * 8 + k) * 32 + lane] = acc[half * 8 + k];
    __syncthreads();
    if (warp == 0 && active) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float4 s = red[(size_t)k * 32 + lane];
        for (int w = 1; w < kWarps; ++w) {
          const float4 t = red[((size_t)w * 8 + k) * 32 + lane];
          s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
        }
        *reinterpret_cast<float4*>(partial + ((size_t)blockIdx.x * kK + half * 8 + k) * C + 4 * q) = s;
      }
    }
    __syncthreads();
  }
}

__global__ void reduce_partials_kernel(const float* __restrict__ partial, int nblk, int K, int C,
                                       float* __restrict__ grad_weights) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= K * C) return;
  const int k = t / C, c = t - k * C;
  float s = 0.0f;
  for (int blk = 0; blk < nblk; ++blk) s += partial[((size_t)blk * kK + k) * C + c];
  grad_weights[t] = s;
}

int weight_blocks(int B, int M) {
  const long long q = (long long)B * M;
  long long blk = (q + kWarps - 1) / kWarps;
  if (blk > 148 * 8) blk = 148 * 8;
  return (int)(blk < 1 ? 1 : blk);
}

size_t stage_smem(int nsample) { return (size_t)kWarps * nsample * (kK * sizeof(float) + sizeof(int)); }

bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace

int d3d_pseudogrid_fwd_tc(const float* feat_cl, const float* query_xyz, const float* support_xyz, const int* idx,
                          const int* nvalid, const int* query_mask, const float* kpoints, const float* weights, int B,
                          int M, int N, int C, int nsample, int K, float extent, int influence, float* out_cl,
                          cudaStream_t st);

extern "C" {

int d3d_pseudogrid_fwd(const float* feat_cl, const float* query_xyz, const float* support_xyz, const int* idx,
                       const int* nvalid, const int* query_mask, const float* kpoints, const float* weights, int B,
                       int M, int N, int C, int nsample, int K, float extent, int influence, int precision,
                       float* out_cl, void* stream) {
  D3D_REQUIRE(feat_cl && query_xyz && support_xyz && idx && nvalid && query_mask && kpoints && weights && out_cl);
  D3D_REQUIRE(B >= 0 && M >= 0 && N > 0 && C > 0 && nsample > 0 && nsample <= D3D_MAX_NSAMPLE);
  D3D_REQUIRE(K > 0 && K <= kK && extent > 0.f && influence >= 0 && influence <= 2 && (precision == 0 || precision == 1));
  if (C % 4 != 0 || !aligned16(feat_cl) || !aligned16(out_cl) || !aligned16(weights)) return D3D_ERR_UNSUPPORTED;
  if (B == 0 || M == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == 1)
    return d3d_pseudogrid_fwd_tc(feat_cl, query_xyz, support_xyz, idx, nvalid, query_mask, kpoints, weights, B, M, N, C,
                                 nsample, K, extent, influence, out_cl, st);
  const size_t smem = stage_smem(nsample);
  cudaError_t e = cudaFuncSetAttribute(pseudogrid_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  dim3 grid(d3d_ceil_div(M, kWarps), B);
  pseudogrid_fwd_kernel<<<grid, kWarps * 32, smem, st>>>(feat_cl, query_xyz, support_xyz, idx, nvalid, query_mask,
                                                         kpoints, weights, M, N, C, nsample, K, extent, influence, out_cl);
  d3d_note_launches(1);
  return d3d_launch_status();
}

size_t d3d_pseudogrid_bwd_workspace_bytes(int B, int M, int C, int K) {
  (void)K;
  if (B <= 0 || M <= 0 || C <= 0) return 0;
  return (size_t)weight_blocks(B, M) * kK * C * sizeof(float);
}

int d3d_pseudogrid_bwd(const float* grad_out_cl, const float* feat_cl, const float* query_xyz,
                       const float* support_xyz, const int* idx, const int* rowptr, const int* entries,
                       const int* nvalid, const int* query_mask, const float* kpoints, const float* weights, int B,
                       int M, int N, int C, int nsample, int K, float extent, int influence, float* grad_feat_cl,
                       float* grad_weights, void* ws, size_t ws_bytes, void* stream) {
  D3D_REQUIRE(grad_out_cl && feat_cl && query_xyz && support_xyz && idx && rowptr && entries && nvalid && query_mask);
  D3D_REQUIRE(kpoints && weights && (grad_feat_cl || grad_weights));
  D3D_REQUIRE(B >= 0 && M >= 0 && N > 0 && C > 0 && nsample > 0 && nsample <= D3D_MAX_NSAMPLE);
  D3D_REQUIRE(K > 0 && K <= kK && extent > 0.f && influence >= 0 && influence <= 2);
  if (C % 4 != 0 || !aligned16(grad_out_cl) || !aligned16(feat_cl) || !aligned16(weights)) return D3D_ERR_UNSUPPORTED;
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (grad_feat_cl) {
    if (!aligned16(grad_feat_cl)) return D3D_ERR_UNSUPPORTED;
    dim3 grid(d3d_ceil_div(N, kWarps), B);
    pseudogrid_bwd_feat_kernel<<<grid, kWarps * 32, 0, st>>>(grad_out_cl, query_xyz, support_xyz, rowptr, entries, nvalid,
                                                             query_mask, kpoints, weights, M, N, C, nsample, K, extent,
                                                             influence, grad_feat_cl);
    d3d_note_launches(1);
  }
  if (grad_weights) {
    if (M == 0) return (int)cudaMemsetAsync(grad_weights, 0, (size_t)K * C * sizeof(float), st);
    if (!ws || ws_bytes < d3d_pseudogrid_bwd_workspace_bytes(B, M, C, K)) return D3D_ERR_WORKSPACE;
    const int nblk = weight_blocks(B, M);
    size_t smem = stage_smem(nsample);
    const size_t red_bytes = (size_t)kWarps * 8 * 32 * sizeof(float4);
    if (smem < red_bytes) smem = red_bytes;
    cudaError_t e = cudaFuncSetAttribute(pseudogrid_bwd_weight_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    dim3 grid(nblk, d3d_ceil_div(C / 4, 32));
    pseudogrid_bwd_weight_kernel<<<grid, kWarps * 32, smem, st>>>(grad_out_cl, feat_cl, query_xyz, support_xyz, idx,
                                                                  nvalid, query_mask, kpoints, B, M, N, C, nsample, K,
                                                                  extent, influence, (float*)ws);
    reduce_partials_kernel<<<d3d_ceil_div((long long)K * C, 256), 256, 0, st>>>((const float*)ws, nblk, K, C, grad_weights);
    d3d_note_launches(2);
  }
  return d3d_launch_status();
}

}  // extern "C"

// Fused local aggregation on channel-last features: PosPool ('xyz' embedding, sum/avg), the gather+max
// of MaskedMaxPool and the row gather of nearest upsampling — forward and backward.
//
//   ref: u_net_arch/models/local_aggregation_operators.py:140-147,165-183   (PosPool)
//   ref: u_net_arch/pt_custom_ops/pt_utils.py:122-148 (MaskedQueryAndGroup), :199-205 (max pool),
//        :158-180,222-226 (nearest upsample)
//
// The reference materialises the (B, C, M, nsample) gather (1.96 GB at the first level) and then
// runs ~6 eager elementwise/reduction passes over it.  Here nothing is materialised:
//   forward : one warp per query.  The lanes first build the per-slot weights (relative position /
//             radius, mask) in shared memory, then stream the neighbours' feature rows as coalesced
//             float4 loads (a row of C floats is contiguous in the channel-last layout) and
//             accumulate in registers;
//   backward: one warp per SUPPORT point walks that point's segment of the inverse map
//             (inverse_map.cu) and accumulates the gradient rows of the queries that gathered it —
//             a segmented reduction in a fixed order: no float atomics, bit-reproducible.
// Both directions move   B*M*nsample*C*4 bytes  through L2 and only  (features + output + idx)  through HBM.
#include "common.cuh"

namespace {

constexpr int kWarps = 8;
constexpr int kMaxNV = 9;  // float4 vectors per lane: up to 9*32*4 = 1152 channels per pass

enum { kModePosPool = 0, kModeMax = 1, kModePlain = 2 };

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// PosPool weight of channel c is (relative position)[c mod 3].  A float4 of channels starting at 4q needs the
// components in the order r, r+1, r+2, r with r = (c_begin + 4q) mod 3 = (c_begin + q) mod 3: the staging code
// writes the three rotations of every slot's weight once, a lane reads the one it needs with a single LDS.128.
__device__ __forceinline__ void store_rotations(float4* dst, float x, float y, float z) {
  dst[0] = make_float4(x, y, z, x);
  dst[1] = make_float4(y, z, x, y);
  dst[2] = make_float4(z, x, y, z);
}

// ------------------------------------------------------------------------------------------------
// forward: warp per query.  Inner loop per (slot, float4 of channels): LDS.128 weight, LDG.128 row, 4 FFMA.
// ------------------------------------------------------------------------------------------------
template <int NV, int MODE>
__global__ void __launch_bounds__(kWarps * 32)
aggregate_fwd_kernel(const float* __restrict__ feat, const float* __restrict__ query_xyz,
                     const float* __restrict__ support_xyz, const int* __restrict__ idx,
                     const int* __restrict__ nvalid, const int* __restrict__ query_mask, int M, int N, int C,
                     int c_begin, int nsample, float inv_radius, int reduction, float* __restrict__ out,
                     uint8_t* __restrict__ argslot) {
  extern __shared__ __align__(16) unsigned char slot_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int j = blockIdx.x * kWarps + warp;
  if (j >= M) return;  // warp-uniform; no block-level barrier below
  // per warp: nsample row offsets (+ 3 weight rotations per slot for PosPool)
  const size_t per_warp = (size_t)nsample * (MODE == kModePosPool ? 3 * sizeof(float4) : 0) + (((size_t)nsample * 4 + 15) & ~(size_t)15);
  float4* sw = reinterpret_cast<float4*>(slot_smem + warp * per_warp);
  int* soff = reinterpret_cast<int*>(slot_smem + warp * per_warp + (MODE == kModePosPool ? (size_t)nsample * 3 * sizeof(float4) : 0));
  const size_t qrow = (size_t)b * M + j;
  const int* irow = idx + qrow * nsample;

  int n_eff = nsample;  // slots that take part
  if (MODE == kModePosPool) {
    const float qx = query_xyz[qrow * 3], qy = query_xyz[qrow * 3 + 1], qz = query_xyz[qrow * 3 + 2];
    // feature_mask = idx_mask + (1 - query_mask): a padded query uses all nsample slots (:171)
    n_eff = query_mask[qrow] != 0 ? nvalid[qrow] : nsample;
    for (int m = lane; m < n_eff; m += 32) {
      const int i = d3d_clamp_index(irow[m], N);
      const float* s = support_xyz + ((size_t)b * N + i) * 3;
      // pt_utils.py:131-133 (the CUDA reference divides by multiplying with 1/radius)
      store_rotations(sw + 3 * m, (s[0] - qx) * inv_radius, (s[1] - qy) * inv_radius, (s[2] - qz) * inv_radius);
      soff[m] = i * C;
    }
  } else {
    for (int m = lane; m < nsample; m += 32) soff[m] = d3d_clamp_index(irow[m], N) * C;
  }
  __syncwarp();

  const int cv = (C - c_begin) >> 2;  // vectors left in this pass
  float4 acc[NV];
  int arg[NV][4];
  int qv[NV], rot[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const float init = MODE == kModeMax ? -INFINITY : 0.0f;
    acc[v] = make_float4(init, init, init, init);
    arg[v][0] = arg[v][1] = arg[v][2] = arg[v][3] = 0;
    qv[v] = min(lane + 32 * v, cv - 1);  // a lane beyond the last vector re-reads the last one; its result is dropped
    rot[v] = (c_begin + qv[v]) % 3;
  }
  const float* fb = feat + (size_t)b * N * C + c_begin;
#pragma unroll 4
  for (int m = 0; m < n_eff; ++m) {
    const float* row = fb + soff[m];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const float4 x = ld4(row + 4 * qv[v]);
      if (MODE == kModePosPool) {
        const float4 w = sw[3 * m + rot[v]];
        acc[v].x += x.x * w.x; acc[v].y += x.y * w.y; acc[v].z += x.z * w.z; acc[v].w += x.w * w.w;
      } else {  // max over all slots, first maximum wins (F.max_pool2d, pt_utils.py:202-205)
        if (x.x > acc[v].x) { acc[v].x = x.x; arg[v][0] = m; }
        if (x.y > acc[v].y) { acc[v].y = x.y; arg[v][1] = m; }
        if (x.z > acc[v].z) { acc[v].z = x.z; arg[v][2] = m; }
        if (x.w > acc[v].w) { acc[v].w = x.w; arg[v][3] = m; }
      }
    }
  }
  float* orow = out + qrow * C + c_begin;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int q = lane + 32 * v;
    if (q < cv) {
      float4 r = acc[v];
      if (MODE == kModePosPool && reduction == D3D_REDUCE_AVG) {
        const float den = (float)n_eff;  // out_features /= neighborhood_num (:175-176)
        r.x /= den; r.y /= den; r.z /= den; r.w /= den;
      }
      *reinterpret_cast<float4*>(orow + 4 * q) = r;
      if (MODE == kModeMax) {
        uchar4 a;
        a.x = (unsigned char)arg[v][0]; a.y = (unsigned char)arg[v][1]; a.z = (unsigned char)arg[v][2]; a.w = (unsigned char)arg[v][3];
        *reinterpret_cast<uchar4*>(argslot + qrow * C + c_begin + 4 * q) = a;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward: warp per support point, segment of the inverse map.  Segment lengths vary a lot (mean nsample, max
// several hundred), so blocks are only kBwdWarps = 2 warps: a short segment does not hold a slot hostage.
// ------------------------------------------------------------------------------------------------
constexpr int kBwdWarps = 2;

template <int NV, int MODE>
__global__ void __launch_bounds__(kBwdWarps * 32)
aggregate_bwd_kernel(const float* __restrict__ grad_out, const float* __restrict__ query_xyz,
                     const float* __restrict__ support_xyz, const int* __restrict__ rowptr,
                     const int* __restrict__ entries, const int* __restrict__ nvalid,
                     const int* __restrict__ query_mask, const uint8_t* __restrict__ argslot, int M, int N, int C,
                     int c_begin, int nsample, float inv_radius, int reduction, float* __restrict__ grad_feat) {
  __shared__ __align__(16) float4 stage_w[kBwdWarps][32 * 3];
  __shared__ int stage_off[kBwdWarps][32];
  __shared__ int stage_k[kBwdWarps][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int i = blockIdx.x * kBwdWarps + warp;
  if (i >= N) return;
  const size_t srow = (size_t)b * N + i;
  const int beg = rowptr[srow], end = rowptr[srow + 1];
  float sx = 0.f, sy = 0.f, sz = 0.f;
  if (MODE == kModePosPool) { sx = support_xyz[srow * 3]; sy = support_xyz[srow * 3 + 1]; sz = support_xyz[srow * 3 + 2]; }

  const int cv = (C - c_begin) >> 2;
  float4 acc[NV];
  int qv[NV], rot[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    qv[v] = min(lane + 32 * v, cv - 1);
    rot[v] = (c_begin + qv[v]) % 3;
  }
  const float* gb = grad_out + (size_t)b * M * C + c_begin;
  const uint8_t* ab = MODE == kModeMax ? argslot + (size_t)b * M * C + c_begin : nullptr;

  for (int e0 = beg; e0 < end; e0 += 32) {
    const int n_here = min(32, end - e0);
    __syncwarp();
    if (lane < n_here) {
      const int packed = entries[e0 + lane];
      const int j = packed >> 8, k = packed & 255;
      if (MODE == kModePosPool) {
        const size_t qrow = (size_t)b * M + j;
        const int n_eff = query_mask[qrow] != 0 ? nvalid[qrow] : nsample;
        float wx = 0.f, wy = 0.f, wz = 0.f;
        if (k < n_eff) {  // a masked slot keeps weight 0 (its row is still read: no branch in the inner loop)
          float scale = inv_radius;
          if (reduction == D3D_REDUCE_AVG) scale /= (float)n_eff;
          wx = (sx - query_xyz[qrow * 3]) * scale;
          wy = (sy - query_xyz[qrow * 3 + 1]) * scale;
          wz = (sz - query_xyz[qrow * 3 + 2]) * scale;
        }
        store_rotations(&stage_w[warp][3 * lane], wx, wy, wz);
      } else if (MODE == kModeMax) {
        stage_k[warp][lane] = k;
      }
      stage_off[warp][lane] = j * C;
    }
    __syncwarp();
    // rows in flight per warp: 8 for one float4 per lane (C <= 128: measured 352 -> 320 us at C = 72), else 4
#pragma unroll(NV == 1 ? 8 : 4)
    for (int t = 0; t < n_here; ++t) {
      const float* row = gb + stage_off[warp][t];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const float4 g = ld4(row + 4 * qv[v]);
        if (MODE == kModePosPool) {
          const float4 w = stage_w[warp][3 * t + rot[v]];
          acc[v].x += g.x * w.x; acc[v].y += g.y * w.y; acc[v].z += g.z * w.z; acc[v].w += g.w * w.w;
        } else if (MODE == kModeMax) {
          const uchar4 a = __ldg(reinterpret_cast<const uchar4*>(ab + stage_off[warp][t] + 4 * qv[v]));
          const int k = stage_k[warp][t];
          acc[v].x += a.x == k ? g.x : 0.f; acc[v].y += a.y == k ? g.y : 0.f;
          acc[v].z += a.z == k ? g.z : 0.f; acc[v].w += a.w == k ? g.w : 0.f;
        } else {
          acc[v].x += g.x; acc[v].y += g.y; acc[v].z += g.z; acc[v].w += g.w;
        }
      }
    }
  }
  float* orow = grad_feat + srow * C + c_begin;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int q = lane + 32 * v;
    if (q < cv) *reinterpret_cast<float4*>(orow + 4 * q) = acc[v];
  }
}

// nearest upsampling forward: out[b, j, :] = feat[b, idx[b, j], :]
__global__ void __launch_bounds__(256)
nearest_gather_fwd_kernel(const float* __restrict__ feat, const int* __restrict__ idx, int M, int N, int CV,
                          float* __restrict__ out) {
  const int b = blockIdx.y;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)M * CV) return;
  const int j = (int)(t / CV), q = (int)(t - (long long)j * CV);
  const int i = d3d_clamp_index(idx[(size_t)b * M + j], N);
  const float4 v = ld4(feat + ((size_t)b * N + i) * CV * 4 + 4 * q);
  *reinterpret_cast<float4*>(out + ((size_t)b * M + j) * CV * 4 + 4 * q) = v;
}

int pick_nv(int cv) {
  const int need = (cv + 31) / 32;
  if (need <= 1) return 1;
  if (need <= 2) return 2;
  if (need <= 3) return 3;
  if (need <= 5) return 5;
  return kMaxNV;
}

template <int MODE>
int launch_fwd(const float* feat, const float* q, const float* s, const int* idx, const int* nvalid, const int* qm,
               int B, int M, int N, int C, int ns, float inv_r, int reduction, float* out, uint8_t* arg,
               cudaStream_t st) {
  const size_t per_warp = (size_t)ns * (MODE == kModePosPool ? 3 * sizeof(float4) : 0) + (((size_t)ns * 4 + 15) & ~(size_t)15);
  const size_t smem = (size_t)kWarps * per_warp;
  dim3 grid(d3d_ceil_div(M, kWarps), B);
  for (int c0 = 0; c0 < C; c0 += kMaxNV * 128) {
    const int cv = (C - c0) / 4;
#define D3D_FWD(NVV)                                                                                           \
  aggregate_fwd_kernel<NVV, MODE><<<grid, kWarps * 32, smem, st>>>(feat, q, s, idx, nvalid, qm, M, N, C, c0, ns, \
                                                                   inv_r, reduction, out, arg)
    switch (pick_nv(cv)) {
      case 1: D3D_FWD(1); break;
      case 2: D3D_FWD(2); break;
      case 3: D3D_FWD(3); break;
      case 5: D3D_FWD(5); break;
      default: D3D_FWD(9); break;
    }
#undef D3D_FWD
    d3d_note_launches(1);
  }
  return d3d_launch_status();
}

template <int MODE>
int launch_bwd(const float* gout, const float* q, const float* s, const int* rowptr, const int* entries,
               const int* nvalid, const int* qm, const uint8_t* arg, int B, int M, int N, int C, int ns, float inv_r,
               int reduction, float* gfeat, cudaStream_t st) {
  dim3 grid(d3d_ceil_div(N, kBwdWarps), B);
  for (int c0 = 0; c0 < C; c0 += kMaxNV * 128) {
    const int cv = (C - c0) / 4;
#define D3D_BWD(NVV)                                                                                              \
  aggregate_bwd_kernel<NVV, MODE><<<grid, kBwdWarps * 32, 0, st>>>(gout, q, s, rowptr, entries, nvalid, qm, arg, M, N, C, \
                                                                c0, ns, inv_r, reduction, gfeat)
    switch (pick_nv(cv)) {
      case 1: D3D_BWD(1); break;
      case 2: D3D_BWD(2); break;
      case 3: D3D_BWD(3); break;
      case 5: D3D_BWD(5); break;
      default: D3D_BWD(9); break;
    }
#undef D3D_BWD
    d3d_note_launches(1);
  }
  return d3d_launch_status();
}

bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace

extern "C" {

int d3d_pospool_fwd(const float* feat_cl, const float* query_xyz, const float* support_xyz, const int* idx,
                    const int* nvalid, const int* query_mask, int B, int M, int N, int C, int nsample, float radius,
                    int reduction, float* out_cl, void* stream) {
  D3D_REQUIRE(feat_cl && query_xyz && support_xyz && idx && nvalid && query_mask && out_cl);
  D3D_REQUIRE(B >= 0 && M >= 0 && N > 0 && C > 0 && nsample > 0 && nsample <= D3D_MAX_NSAMPLE && radius > 0.f);
  D3D_REQUIRE(reduction == D3D_REDUCE_SUM || reduction == D3D_REDUCE_AVG);
  if (C % 4 != 0 || !aligned16(feat_cl) || !aligned16(out_cl)) return D3D_ERR_UNSUPPORTED;
  if (B == 0 || M == 0) return 0;
  return launch_fwd<kModePosPool>(feat_cl, query_xyz, support_xyz, idx, nvalid, query_mask, B, M, N, C, nsample,
                                  1.0f / radius, reduction, out_cl, nullptr, (cudaStream_t)stream);
}

int d3d_pospool_bwd(const float* grad_out_cl, const float* query_xyz, const float* support_xyz, const int* rowptr,
                    const int* entries, const int* nvalid, const int* query_mask, int B, int M, int N, int C,
                    int nsample, float radius, int reduction, float* grad_feat_cl, void* stream) {
  D3D_REQUIRE(grad_out_cl && query_xyz && support_xyz && rowptr && entries && nvalid && query_mask && grad_feat_cl);
  D3D_REQUIRE(B >= 0 && M >= 0 && N > 0 && C > 0 && nsample > 0 && nsample <= D3D_MAX_NSAMPLE && radius > 0.f);
  D3D_REQUIRE(reduction == D3D_REDUCE_SUM || reduction == D3D_REDUCE_AVG);
  if (C % 4 != 0 || !aligned16(grad_out_cl) || !aligned16(grad_feat_cl)) return D3D_ERR_UNSUPPORTED;
  if (B == 0) return 0;
  return launch_bwd<kModePosPool>(grad_out_cl, query_xyz, support_xyz, rowptr, entries, nvalid, query_mask, nullptr,
                                  B, M, N, C, nsample, 1.0f / radius, reduction, grad_feat_cl, (cudaStream_t)stream);
}

int d3d_gather_max_fwd(const float* feat_cl, const int* idx, int B, int M, int N, int C, int nsample, float* out_cl,
                       uint8_t* argslot, void* stream) {
  D3D_REQUIRE(feat_cl && idx && out_cl && argslot);
  D3D_REQUIRE(B >= 0 && M >= 0 && N > 0 && C > 0 && nsample > 0 && nsample <= D3D_MAX_NSAMPLE);
  if (C % 4 != 0 || !aligned16(feat_cl) || !aligned16(out_cl)) return D3D_ERR_UNSUPPORTED;
  if (B == 0 || M == 0) return 0;
  return launch_fwd<kModeMax>(feat_cl, nullptr, nullptr, idx, nullptr, nullptr, B, M, N, C, nsample, 0.f, 0, out_cl,
                              argslot, (cudaStream_t)stream);
}

int d3d_gather_max_bwd(const float* grad_out_cl, const uint8_t* argslot, const int* rowptr, const int* entries, int B,
                       int M, int N, int C, float* grad_feat_cl, void* stream) {
  D3D_REQUIRE(grad_out_cl && argslot && rowptr && entries && grad_feat_cl);
  D3D_REQUIRE(B >= 0 && M >= 0 && N > 0 && C > 0);
  if (C % 4 != 0 || !aligned16(grad_out_cl) || !aligned16(grad_feat_cl)) return D3D_ERR_UNSUPPORTED;
  if (B == 0) return 0;
  return launch_bwd<kModeMax>(grad_out_cl, nullptr, nullptr, rowptr, entries, nullptr, nullptr, argslot, B, M, N, C, 0,
                              0.f, 0, grad_feat_cl, (cudaStream_t)stream);
}

int d3d_nearest_gather_fwd(const float* feat_cl, const int* idx, int B, int M, int N, int C, float* out_cl,
                           void* stream) {
  D3D_REQUIRE(feat_cl && idx && out_cl);
  D3D_REQUIRE(B >= 0 && M >= 0 && N > 0 && C > 0);
  if (C % 4 != 0 || !aligned16(feat_cl) || !aligned16(out_cl)) return D3D_ERR_UNSUPPORTED;
  if (B == 0 || M == 0) return 0;
  dim3 grid(d3d_ceil_div((long long)M * (C / 4), 256), B);
  nearest_gather_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(feat_cl, idx, M, N, C / 4, out_cl);
  d3d_note_launches(1);
  return d3d_launch_status();
}

int d3d_nearest_gather_bwd(const float* grad_out_cl, const int* rowptr, const int* entries, int B, int M, int N, int C,
                           float* grad_feat_cl, void* stream) {
  D3D_REQUIRE(grad_out_cl && rowptr && entries && grad_feat_cl);
  D3D_REQUIRE(B >= 0 && M >= 0 && N > 0 && C > 0);
  if (C % 4 != 0 || !aligned16(grad_out_cl) || !aligned16(grad_feat_cl)) return D3D_ERR_UNSUPPORTED;
  if (B == 0) return 0;
  return launch_bwd<kModePlain>(grad_out_cl, nullptr, nullptr, rowptr, entries, nullptr, nullptr, nullptr, B, M, N, C,
                                0, 0.f, 0, grad_feat_cl, (cudaStream_t)stream);
}

}  // extern "C"

// PseudoGrid (depthwise KPConv-style) aggregation kernels: forward, gradient w.r.t. features, gradient w.r.t.
// kernel weights — each with the kernel-point contraction either on CUDA cores (fp32) or on the 5th-generation
// tensor cores (tcgen05 + TMEM, bf16 operands, fp32 accumulate).
//
//   ref: u_net_arch/models/local_aggregation_operators.py:467-503
//   w[j,m,k]    = influence(|| (S[idx[j,m]] - Q[j]) - K[k] ||) * fm[j,m]
//   out[j,c]    = sum_m F[idx[j,m], c] * E[c,(j,m)],          E[c,(j,m)] = sum_k W[k,c] * w[j,m,k]
//   dF[i,c]     = sum_{(j,m) in inverse map of i} g[j,c] * E[c,(j,m)]
//   dW[k,c]     = sum_{j,m} w[j,m,k] * F[idx[j,m], c] * g[j,c]
//
// One thread owns ONE channel: the 32 lanes of a warp read 32 consecutive channels of the same gathered row (one
// coalesced 128-byte request per neighbour, many requests in flight per thread), and every reduction over
// neighbours happens in that thread's registers — no cross-thread reduction anywhere.
//
// E is a dense GEMM  E^T[C x items] = W^T[C x 16] . w^T[16 x items]  (K = 15 kernel points padded to 16 = exactly one
// tcgen05.mma K-step for bf16).  Tensor-core mapping: MMA M (TMEM lanes) = 128 channels (A operand = W^T tile,
// staged once per CTA), MMA N (TMEM columns) = up to 64.. 256 items (B operand = influence weights), accumulator
// lane = channel, column = item — so after tcgen05.ld a thread holds E for its own channel and all items.
// Operand layout: K-major, no swizzle (8-row x 16-byte core matrices, K chunks LBO = 128 B apart, row groups
// SBO = 256 B apart).  Features/gradients stay fp32; only W and the influence weights are rounded to bf16.
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

constexpr int kThreads = 128;     // 4 warps = the 4 TMEM lane quarters; 128 channels per channel tile
constexpr int kRowsPerCta = 8;    // rows (queries / supports) processed one after the other by a CTA
constexpr int kMaxCTiles = 9;     // 9 x 128 = 1152 channels
constexpr int kTileBytes = 128 * 32;
constexpr unsigned kLbo = 128, kSbo = 256;
constexpr int kK = 16;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ unsigned operand_offset(int row, int k) {
  return (unsigned)((row >> 3) * kSbo + (k >> 3) * kLbo + (row & 7) * 16 + (k & 7) * 2);
}

__device__ __forceinline__ unsigned long long make_smem_desc(unsigned addr) {
  return (unsigned long long)((addr >> 4) & 0x3fffu) | ((unsigned long long)(kLbo >> 4) << 16) |
         ((unsigned long long)(kSbo >> 4) << 32) | (1ull << 46);  // version 1 (sm_100), layout type 0 = no swizzle
}

// coef = 1/extent (linear) or -1/(2 sigma^2 + 1e-9), sigma = 0.3 extent (gaussian); multiplying by the reciprocal
// instead of dividing moves the result by <= 1 ulp of the quotient, far inside the stated fp32 tolerance
template <int kInfluence>
__device__ __forceinline__ float influence_weight(float dx, float dy, float dz, float coef) {
  const float sq = dx * dx + dy * dy + dz * dz;
  if (kInfluence == D3D_KP_LINEAR) {  // :480   (sqrt.approx: <= 1 ulp, no slow-path call)
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(sq));
    return fmaxf(1.0f - r * coef, 0.0f);
  }
  if (kInfluence == D3D_KP_GAUSSIAN) return __expf(sq * coef);                   // :484, models/utlis.py:287-294
  return 1.0f;                                                                   // constant (:476)
}

__host__ __device__ inline float influence_coef(float extent, int influence) {
  if (influence == D3D_KP_GAUSSIAN) {
    const float sigma = extent * 0.3f;
    return -1.0f / (2.0f * sigma * sigma + 1e-9f);
  }
  return 1.0f / extent;
}

// stage 2 of a piece: influence weight of every (item, kernel point); srow < 0 marks a masked item.
// thread = (item, pair of adjacent kernel points): the relative position is read once per two weights and the two
// bf16 go out as one 32-bit store (adjacent k are adjacent in the K-major operand layout).
template <int kInfluence, bool kTensorCore>
__device__ __forceinline__ void stage_weights(int tid, int pmax, int K, const int* srow, const float* srel,
                                              const float* kp, float coef, unsigned char* b_tile, float* w_f32) {
  const int k0 = (tid & 7) * 2;  // 128 threads: a thread's kernel-point pair never changes
  const float ax = kp[3 * k0], ay = kp[3 * k0 + 1], az = kp[3 * k0 + 2];
  const float bx = kp[3 * k0 + 3], by = kp[3 * k0 + 4], bz = kp[3 * k0 + 5];
  for (int t = tid; t < pmax * 8; t += 128) {
    const int p = t >> 3;
    float w0 = 0.0f, w1 = 0.0f;
    if (srow[p] >= 0) {
      const float dx = srel[3 * p], dy = srel[3 * p + 1], dz = srel[3 * p + 2];
      if (k0 < K) w0 = influence_weight<kInfluence>(dx - ax, dy - ay, dz - az, coef);
      if (k0 + 1 < K) w1 = influence_weight<kInfluence>(dx - bx, dy - by, dz - bz, coef);
    }
    if (kTensorCore)
      *reinterpret_cast<__nv_bfloat162*>(b_tile + operand_offset(p, k0)) = __floats2bfloat162_rn(w0, w1);
    else
      *reinterpret_cast<float2*>(w_f32 + p * 16 + k0) = make_float2(w0, w1);
  }
}

// Gathered element at an unsigned BYTE offset from the thread's base pointer: the address is one 64-bit add.  (An int
// element offset costs a sign extension, a shift and the add — 5 SASS instructions per load in the first version,
// a third of this instruction-issue-bound kernel.)
__device__ __forceinline__ float gather_at(const char* base, int byte_off) {
  return __ldg(reinterpret_cast<const float*>(base + (unsigned)byte_off));
}

__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}

struct Args {
  const float* src;          // rows that are gathered: features (forward) or grad_out (backward), channel-last
  const float* query_xyz;    // (B, M, 3)
  const float* support_xyz;  // (B, N, 3)
  const int* idx;            // (B, M, ns)          forward
  const int* rowptr;         // inverse map         backward
  const int* entries;
  const int* nvalid;         // (B, M)
  const int* query_mask;     // (B, M)
  const float* kpoints;      // (K, 3)
  const float* weights;      // (K, C)
  float* out;                // (B, rows, C)
  int M, N, C, nsample, K, influence;
  float extent;
  int n_items_max;  // items staged per piece (multiple of 16, <= 256)
  int tmem_cols;
  int rows_per_cta;
};

// kBackward: rows are support points and items are inverse-map entries; else rows are queries, items are slots.
// kTensorCore: E from tcgen05.mma into TMEM; else 16 fp32 FMAs per (item, channel) against W in registers.
// ncu: the tensor-core variant is instruction-issue bound (73 % of the issue slots; 40 % of the instructions are the
// gather epilogue, 26 % the influence weights), with barrier waits between the per-row phases as the main stall — so
// it runs 8 resident CTAs per SM (<= 64 registers, 64 TMEM columns each) and keeps the per-gather address arithmetic
// at one 64-bit add.  Prefetching the next row's indices was measured and changed nothing (not latency-bound).
template <bool kBackward, bool kTensorCore>
__global__ void __launch_bounds__(kThreads, kTensorCore ? 8 : 6)
pseudogrid_rows_kernel(const Args a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ unsigned tmem_base_slot;
  __shared__ float kp[52];

  const int tid = threadIdx.x, warp = tid >> 5;
  const int b = blockIdx.y;
  const int C = a.C, K = a.K;
  const int n_ctile = (C + 127) >> 7;
  const int n_rows = kBackward ? a.N : a.M;
  const int pmax = a.n_items_max;
  // shared-memory carve-up
  unsigned char* a_tiles = smem;                                                      // tensor core: n_ctile x 4 KB
  unsigned char* b_tile = a_tiles + (kTensorCore ? (size_t)n_ctile * kTileBytes : 0);  // tc: pmax x 32 B bf16
  float* w_f32 = reinterpret_cast<float*>(b_tile);                                    // fp32: pmax x 16 floats
  int* srow = reinterpret_cast<int*>(b_tile + (kTensorCore ? (size_t)pmax * 32 : (size_t)pmax * 64));
  float* srel = reinterpret_cast<float*>(srow + pmax);  // pmax x 3

  unsigned tmem_base = 0, bar = 0, phase = 0, idesc = 0;
  unsigned long long b_desc = 0;
  if (kTensorCore) {
    if (warp == 0) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                   "r"((unsigned)a.tmem_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // A tiles (W^T, K-major): a thread owns one channel = one 32-byte operand row per tile: 16 coalesced loads, two
    // 16-byte stores (k 0..7 and, one K chunk = LBO further, k 8..15)
    for (int ct = 0; ct < n_ctile; ++ct) {
      const int c = ct * 128 + tid;
      __nv_bfloat162 h[8];
#pragma unroll
      for (int k2 = 0; k2 < 8; ++k2) {
        const float w0 = (c < C && 2 * k2 < K) ? __ldg(a.weights + (size_t)(2 * k2) * C + c) : 0.0f;
        const float w1 = (c < C && 2 * k2 + 1 < K) ? __ldg(a.weights + (size_t)(2 * k2 + 1) * C + c) : 0.0f;
        h[k2] = __floats2bfloat162_rn(w0, w1);
      }
      unsigned char* dst = a_tiles + (size_t)ct * kTileBytes + (tid >> 3) * kSbo + (tid & 7) * 16;
      *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(&h[0]);
      *reinterpret_cast<uint4*>(dst + kLbo) = *reinterpret_cast<const uint4*>(&h[4]);
    }
  }
  if (tid < 52) kp[tid] = tid < K * 3 ? a.kpoints[tid] : 0.0f;
  if (kTensorCore) asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (kTensorCore) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tmem_base = tmem_base_slot;
    bar = smem_u32(&mbar);
    // instruction descriptor: D = f32, A = B = bf16, both K-major, N = pmax, M = 128
    idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(pmax >> 3) << 17) | ((128u >> 4) << 24);
    b_desc = make_smem_desc(smem_u32(b_tile));
  }
  const float* src_b = a.src + (size_t)b * (kBackward ? a.M : a.N) * C;
  const float coef = influence_coef(a.extent, a.influence);

  for (int ri = 0; ri < a.rows_per_cta; ++ri) {
    const int row = blockIdx.x * a.rows_per_cta + ri;
    if (row >= n_rows) break;  // block-uniform
    const size_t grow = (size_t)b * n_rows + row;
    int item_beg = 0, n_items = 0;
    float rx, ry, rz;  // the row's own coordinates
    if (kBackward) {
      item_beg = a.rowptr[grow];
      n_items = a.rowptr[grow + 1] - item_beg;
      rx = a.support_xyz[grow * 3]; ry = a.support_xyz[grow * 3 + 1]; rz = a.support_xyz[grow * 3 + 2];
    } else {
      n_items = a.query_mask[grow] != 0 ? a.nvalid[grow] : a.nsample;  // fm: padded query uses every slot (:490)
      rx = a.query_xyz[grow * 3]; ry = a.query_xyz[grow * 3 + 1]; rz = a.query_xyz[grow * 3 + 2];
    }
    float acc[kMaxCTiles];
#pragma unroll
    for (int ct = 0; ct < kMaxCTiles; ++ct) acc[ct] = 0.0f;

    for (int p0 = 0; p0 < n_items; p0 += pmax) {
      const int np = min(pmax, n_items - p0);
      // ---- stage 1: which row each item gathers, and its relative position
      for (int p = tid; p < pmax; p += kThreads) {
        int g = -1;
        float dx = 0.f, dy = 0.f, dz = 0.f;
        if (p < np) {
          if (kBackward) {
            const int e = a.entries[item_beg + p0 + p];
            const int j = e >> 8, m = e & 255;
            const size_t qrow = (size_t)b * a.M + j;
            const int n_eff = a.query_mask[qrow] != 0 ? a.nvalid[qrow] : a.nsample;
            if (m < n_eff) {
              g = j;
              dx = rx - a.query_xyz[qrow * 3]; dy = ry - a.query_xyz[qrow * 3 + 1]; dz = rz - a.query_xyz[qrow * 3 + 2];
            }
          } else {
            g = d3d_clamp_index(a.idx[grow * a.nsample + p0 + p], a.N);
            const float* s = a.support_xyz + ((size_t)b * a.N + g) * 3;
            dx = s[0] - rx; dy = s[1] - ry; dz = s[2] - rz;
          }
        }
        srow[p] = g;
        srel[3 * p] = dx; srel[3 * p + 1] = dy; srel[3 * p + 2] = dz;
      }
      __syncthreads();
      // ---- stage 2: influence weights of (item, kernel point)
      if (a.influence == D3D_KP_LINEAR) stage_weights<D3D_KP_LINEAR, kTensorCore>(tid, pmax, K, srow, srel, kp, coef, b_tile, w_f32);
      else if (a.influence == D3D_KP_GAUSSIAN) stage_weights<D3D_KP_GAUSSIAN, kTensorCore>(tid, pmax, K, srow, srel, kp, coef, b_tile, w_f32);
      else stage_weights<D3D_KP_CONSTANT, kTensorCore>(tid, pmax, K, srow, srel, kp, coef, b_tile, w_f32);
      __syncthreads();
      // gather offsets replace the row numbers (masked items read row 0 with weight 0)
      for (int p = tid; p < pmax; p += kThreads) srow[p] = srow[p] >= 0 ? srow[p] * C * 4 : 0;
      if (kTensorCore) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      }
      __syncthreads();

#pragma unroll
      for (int ct = 0; ct < kMaxCTiles; ++ct) {
        if (ct >= n_ctile) break;
        const int c = ct * 128 + tid;
        const bool active = c < C;
        const char* sc = reinterpret_cast<const char*>(src_b + (active ? c : 0));
        float sum = 0.0f;
        if (kTensorCore) {
          if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const unsigned long long a_desc = make_smem_desc(smem_u32(a_tiles + (size_t)ct * kTileBytes));
            asm volatile(
                "{\n\t"
                ".reg .pred p;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
                "}\n" ::"r"(tmem_base), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(0u) : "memory");
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
          }
          mbar_wait(bar, phase);
          phase ^= 1u;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (ct * 128 + warp * 32 < C) {  // warp-uniform: a warp whose 32 channels are all padding skips the epilogue
            for (int col0 = 0; col0 < np; col0 += 16) {
              unsigned r[16];
              const unsigned taddr = tmem_base + ((unsigned)(warp * 32) << 16) + (unsigned)col0;
              asm volatile(
                  "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                  "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                  : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                    "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                  : "r"(taddr));
              // gathers: unconditional (a padding lane re-reads channel 0, a masked / padding item reads row 0 with
              // E = 0), all 16 issued before the first use; offsets come as 4 x int4 broadcast reads
              float x[16];
#pragma unroll
              for (int i4 = 0; i4 < 4; ++i4) {
                const int4 o = *reinterpret_cast<const int4*>(srow + col0 + 4 * i4);
                x[4 * i4] = gather_at(sc, o.x); x[4 * i4 + 1] = gather_at(sc, o.y);
                x[4 * i4 + 2] = gather_at(sc, o.z); x[4 * i4 + 3] = gather_at(sc, o.w);
              }
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
              for (int i = 0; i < 16; ++i) sum += x[i] * __uint_as_float(r[i]);
            }
          }
          // every warp is done reading TMEM before the next MMA overwrites it
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncthreads();
        } else if (ct * 128 + warp * 32 < C) {  // warp-uniform
          float W[kK];
#pragma unroll
          for (int k = 0; k < kK; ++k) W[k] = (active && k < K) ? __ldg(a.weights + (size_t)k * C + c) : 0.0f;
          for (int q0 = 0; q0 < np; q0 += 8) {  // q0 + 8 <= pmax: offset 0 / weight 0 beyond np
            float x[8];
            const int4 o0 = *reinterpret_cast<const int4*>(srow + q0), o1 = *reinterpret_cast<const int4*>(srow + q0 + 4);
            x[0] = gather_at(sc, o0.x); x[1] = gather_at(sc, o0.y); x[2] = gather_at(sc, o0.z); x[3] = gather_at(sc, o0.w);
            x[4] = gather_at(sc, o1.x); x[5] = gather_at(sc, o1.y); x[6] = gather_at(sc, o1.z); x[7] = gather_at(sc, o1.w);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4* wp = reinterpret_cast<const float4*>(w_f32 + (q0 + i) * kK);  // broadcast reads
              float e = 0.0f;
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) {
                const float4 w = wp[k4];
                e += w.x * W[4 * k4] + w.y * W[4 * k4 + 1] + w.z * W[4 * k4 + 2] + w.w * W[4 * k4 + 3];
              }
              sum += x[i] * e;
            }
          }
        }
        acc[ct] += sum;
      }
      if (!kTensorCore) __syncthreads();  // the staging buffers are rewritten by the next piece / row
    }
#pragma unroll
    for (int ct = 0; ct < kMaxCTiles; ++ct) {
      const int c = ct * 128 + tid;
      if (ct < n_ctile && c < C) a.out[grow * C + c] = acc[ct];
    }
  }

  if (kTensorCore) {
    __syncthreads();
    if (warp == 0)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((unsigned)a.tmem_cols) : "memory");
  }
}

// dW partials (fp32, CUDA cores): grid (channel tiles, nblk); thread = channel, 16 accumulators in registers.
__global__ void __launch_bounds__(kThreads)
pseudogrid_weight_grad_kernel(const float* __restrict__ grad_out, const float* __restrict__ feat,
                              const float* __restrict__ query_xyz, const float* __restrict__ support_xyz,
                              const int* __restrict__ idx, const int* __restrict__ nvalid,
                              const int* __restrict__ query_mask, const float* __restrict__ kpoints, int B, int M, int N,
                              int C, int nsample, int K, float extent, int influence,
                              float* __restrict__ partial /* (gridDim.y, 16, C) */) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ float kp[52];
  const int pm = (nsample + 7) & ~7;
  float* w_f32 = reinterpret_cast<float*>(smem);        // pm x 16
  int* srow = reinterpret_cast<int*>(w_f32 + pm * kK);  // pm
  float* srel = reinterpret_cast<float*>(srow + pm);    // pm x 3
  const int tid = threadIdx.x;
  if (tid < 52) kp[tid] = tid < K * 3 ? kpoints[tid] : 0.0f;
  const int c = blockIdx.x * 128 + tid;
  const bool active = c < C;
  float acc[kK];
#pragma unroll
  for (int k = 0; k < kK; ++k) acc[k] = 0.0f;
  const long long total = (long long)B * M;
  const float coef = influence_coef(extent, influence);
  const int pmax = (nsample + 7) & ~7;
  for (long long qi = blockIdx.y; qi < total; qi += gridDim.y) {
    const int b = (int)(qi / M);
    const int n_eff = query_mask[qi] != 0 ? nvalid[qi] : nsample;
    const float qx = query_xyz[qi * 3], qy = query_xyz[qi * 3 + 1], qz = query_xyz[qi * 3 + 2];
    __syncthreads();  // previous query fully consumed
    for (int p = tid; p < pmax; p += kThreads) {
      int g = -1;
      float dx = 0.f, dy = 0.f, dz = 0.f;
      if (p < n_eff) {
        g = d3d_clamp_index(idx[qi * nsample + p], N);
        const float* s = support_xyz + ((size_t)b * N + g) * 3;
        dx = s[0] - qx; dy = s[1] - qy; dz = s[2] - qz;
      }
      srow[p] = g;
      srel[3 * p] = dx; srel[3 * p + 1] = dy; srel[3 * p + 2] = dz;
    }
    __syncthreads();
    if (influence == D3D_KP_LINEAR) stage_weights<D3D_KP_LINEAR, false>(tid, pmax, K, srow, srel, kp, coef, nullptr, w_f32);
    else if (influence == D3D_KP_GAUSSIAN) stage_weights<D3D_KP_GAUSSIAN, false>(tid, pmax, K, srow, srel, kp, coef, nullptr, w_f32);
    else stage_weights<D3D_KP_CONSTANT, false>(tid, pmax, K, srow, srel, kp, coef, nullptr, w_f32);
    __syncthreads();
    for (int p = tid; p < pmax; p += kThreads) srow[p] = srow[p] >= 0 ? srow[p] * C * 4 : 0;
    __syncthreads();
    if (blockIdx.x * 128 + (tid & ~31) < C) {  // warp-uniform: skip warps made of padding channels only
      const int cc = active ? c : 0;
      const float g = active ? __ldg(grad_out + (size_t)qi * C + cc) : 0.0f;
      const char* fc = reinterpret_cast<const char*>(feat + (size_t)b * N * C + cc);
      for (int q0 = 0; q0 < n_eff; q0 += 8) {  // masked tail: weight 0, offset 0
        float x[8];
        const int4 o0 = *reinterpret_cast<const int4*>(srow + q0), o1 = *reinterpret_cast<const int4*>(srow + q0 + 4);
        x[0] = gather_at(fc, o0.x) * g; x[1] = gather_at(fc, o0.y) * g; x[2] = gather_at(fc, o0.z) * g; x[3] = gather_at(fc, o0.w) * g;
        x[4] = gather_at(fc, o1.x) * g; x[5] = gather_at(fc, o1.y) * g; x[6] = gather_at(fc, o1.z) * g; x[7] = gather_at(fc, o1.w) * g;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4* wp = reinterpret_cast<const float4*>(w_f32 + (q0 + i) * kK);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const float4 w = wp[k4];
            acc[4 * k4] += w.x * x[i]; acc[4 * k4 + 1] += w.y * x[i]; acc[4 * k4 + 2] += w.z * x[i]; acc[4 * k4 + 3] += w.w * x[i];
          }
        }
      }
    }
  }
  if (active) {
#pragma unroll
    for (int k = 0; k < kK; ++k) partial[((size_t)blockIdx.y * kK + k) * C + c] = acc[k];
  }
}

// dW partials on the tensor cores (bf16 mode).  dW^T[c, k] = sum_items P[c, item] * w[item, k] with
// P[c, item] = F[idx_item, c] * g[row, c]:  MMA M = 128 channels (TMEM lanes), N = 16 kernel points, K = items
// (16 per K-step, up to 64 per row), accumulated in ONE TMEM tile over every row the CTA visits.  A thread owns a
// channel: it gathers its 16 features per K-step, multiplies by g, rounds to bf16 and writes them as two 16-byte
// stores into the K-major A tile (its own row of the tile); the B tile holds the influence weights transposed.
// kDwBufs staging buffers (see below).
constexpr int kDwItems = 64;                         // items staged per buffer (4 K-steps)
constexpr int kDwABytes = (kDwItems / 16) * 4096;    // 4 x [128 x 16] bf16
constexpr int kDwBBytes = (kDwItems / 16) * 512;     // 4 x [16 x 16] bf16
constexpr int kDwBufBytes = kDwABytes + kDwBBytes;
// Staging buffers per CTA.  One: 23 KB of shared memory per CTA, 8 resident CTAs per SM hide each other's MMA waits
// and row latencies (measured faster than two buffers at 5 CTAs per SM: the kernel is bound by the per-row chain).
constexpr int kDwBufs = 1;

__global__ void __launch_bounds__(kThreads, 8)
pseudogrid_weight_grad_tc_kernel(const float* __restrict__ grad_out, const float* __restrict__ feat,
                                 const float* __restrict__ query_xyz, const float* __restrict__ support_xyz,
                                 const int* __restrict__ idx, const int* __restrict__ nvalid,
                                 const int* __restrict__ query_mask, const float* __restrict__ kpoints, int B, int M,
                                 int N, int C, int nsample, int K, float extent, int influence,
                                 float* __restrict__ partial /* (gridDim.y, 16, C) */) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long mbar[2];
  __shared__ unsigned tmem_base_slot;
  __shared__ float kp[52];
  unsigned char* bufs = smem;                                           // 2 x (A | B)
  int* srow = reinterpret_cast<int*>(smem + kDwBufs * kDwBufBytes);     // kDwItems offsets
  float* srel = reinterpret_cast<float*>(srow + kDwItems);              // kDwItems x 3
  float* w_f32 = srel + 3 * kDwItems;                                   // kDwItems x 16 (fp32 staging of the weights)
  const int tid = threadIdx.x, warp = tid >> 5;
  const int c = blockIdx.x * 128 + tid;
  const bool active = c < C;
  const bool warp_active = blockIdx.x * 128 + warp * 32 < C;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[0])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[1])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 52) kp[tid] = tid < K * 3 ? kpoints[tid] : 0.0f;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem_base = tmem_base_slot;
  // D = f32, A = B = bf16, K-major, N = 16, M = 128
  const unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
  const float coef = influence_coef(extent, influence);
  unsigned phase[2] = {0u, 0u};
  int pending[2] = {0, 0};  // an MMA group that reads this buffer has been committed and not yet waited for
  int n_issued = 0;         // MMAs issued so far (the first one overwrites the accumulator)
  int buf = 0;

  const long long total = (long long)B * M;
  for (long long qi = blockIdx.y; qi < total; qi += gridDim.y) {
    const int b = (int)(qi / M);
    const int n_eff = query_mask[qi] != 0 ? nvalid[qi] : nsample;
    const float qx = query_xyz[qi * 3], qy = query_xyz[qi * 3 + 1], qz = query_xyz[qi * 3 + 2];
    const float g = active ? __ldg(grad_out + (size_t)qi * C + c) : 0.0f;
    const char* fc = reinterpret_cast<const char*>(feat + (size_t)b * N * C + (active ? c : 0));
    for (int p0 = 0; p0 < n_eff; p0 += kDwItems) {
      const int np = min(kDwItems, n_eff - p0);
      const int ksteps = (np + 15) >> 4;
      unsigned char* a_tile = bufs + (size_t)buf * kDwBufBytes;
      unsigned char* b_tile = a_tile + kDwABytes;
      // the tensor core must be done with this buffer (the MMA group committed two pieces ago)
      if (pending[buf]) {
        mbar_wait(smem_u32(&mbar[buf]), phase[buf]);
        phase[buf] ^= 1u;
        pending[buf] = 0;
      }
      __syncthreads();  // srow / srel / w_f32 of the previous piece fully consumed
      if (tid < kDwItems) {
        int gi = -1;
        float dx = 0.f, dy = 0.f, dz = 0.f;
        if (tid < np) {
          gi = d3d_clamp_index(idx[qi * nsample + p0 + tid], N);
          const float* sp = support_xyz + ((size_t)b * N + gi) * 3;
          dx = sp[0] - qx; dy = sp[1] - qy; dz = sp[2] - qz;
        }
        srow[tid] = gi;
        srel[3 * tid] = dx; srel[3 * tid + 1] = dy; srel[3 * tid + 2] = dz;
      }
      __syncthreads();
      if (influence == D3D_KP_LINEAR) stage_weights<D3D_KP_LINEAR, false>(tid, ksteps * 16, K, srow, srel, kp, coef, nullptr, w_f32);
      else if (influence == D3D_KP_GAUSSIAN) stage_weights<D3D_KP_GAUSSIAN, false>(tid, ksteps * 16, K, srow, srel, kp, coef, nullptr, w_f32);
      else stage_weights<D3D_KP_CONSTANT, false>(tid, ksteps * 16, K, srow, srel, kp, coef, nullptr, w_f32);
      __syncthreads();
      // B tiles: w^T, row = kernel point, column = item of the K-step;  gather offsets replace the row numbers
      for (int t = tid; t < ksteps * 256; t += kThreads) {
        const int p = t >> 4, k = t & 15;
        *reinterpret_cast<__nv_bfloat16*>(b_tile + (p >> 4) * 512 + operand_offset(k, p & 15)) = __float2bfloat16_rn(w_f32[t]);
      }
      if (tid < kDwItems) srow[tid] = srow[tid] >= 0 ? srow[tid] * C * 4 : 0;
      __syncthreads();
      // A tiles: this thread's row (channel) of every K-step: 16 products -> 16 bf16 -> two 16-byte stores
      if (warp_active) {
        for (int s = 0; s < ksteps; ++s) {
          float x[16];
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            const int4 o = *reinterpret_cast<const int4*>(srow + s * 16 + 4 * i4);
            x[4 * i4] = gather_at(fc, o.x); x[4 * i4 + 1] = gather_at(fc, o.y); x[4 * i4 + 2] = gather_at(fc, o.z); x[4 * i4 + 3] = gather_at(fc, o.w);
          }
          __nv_bfloat162 h[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) h[i] = __floats2bfloat162_rn(x[2 * i] * g, x[2 * i + 1] * g);
          unsigned char* dst = a_tile + s * 4096 + (tid >> 3) * kSbo + (tid & 7) * 16;
          *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(&h[0]);         // items 0..7  of the K-step
          *reinterpret_cast<uint4*>(dst + kLbo) = *reinterpret_cast<const uint4*>(&h[4]);  // items 8..15
        }
      } else {
        for (int s = 0; s < ksteps; ++s) {  // padding channels: zero rows (the tile is read in full by the MMA)
          unsigned char* dst = a_tile + s * 4096 + (tid >> 3) * kSbo + (tid & 7) * 16;
          *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
          *reinterpret_cast<uint4*>(dst + kLbo) = make_uint4(0u, 0u, 0u, 0u);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int s = 0; s < ksteps; ++s) {
          const unsigned long long a_desc = make_smem_desc(smem_u32(a_tile + s * 4096));
          const unsigned long long b_desc = make_smem_desc(smem_u32(b_tile + s * 512));
          const unsigned accumulate = (n_issued + s) > 0 ? 1u : 0u;
          asm volatile(
              "{\n\t"
              ".reg .pred p;\n\t"
              "setp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
              "}\n" ::"r"(tmem_base), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar[buf])) : "memory");
      }
      n_issued += ksteps;
      pending[buf] = 1;
      buf = (buf + 1) % kDwBufs;
    }
  }
  // drain: both buffers' MMA groups complete (in issue order), then read the accumulator
  for (int k = 0; k < kDwBufs; ++k) {
    if (pending[buf]) {
      mbar_wait(smem_u32(&mbar[buf]), phase[buf]);
      phase[buf] ^= 1u;
      pending[buf] = 0;
    }
    buf = (buf + 1) % kDwBufs;
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  unsigned r[16];
  if (n_issued > 0) {
    const unsigned taddr = tmem_base + ((unsigned)(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  } else {
#pragma unroll
    for (int k = 0; k < 16; ++k) r[k] = 0u;
  }
  if (active) {
#pragma unroll
    for (int k = 0; k < kK; ++k) partial[((size_t)blockIdx.y * kK + k) * C + c] = __uint_as_float(r[k]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(32u) : "memory");
}

__global__ void reduce_partials_kernel(const float* __restrict__ partial, int nblk, int K, int C,
                                       float* __restrict__ grad_weights) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= K * C) return;
  const int k = t / C, c = t - k * C;
  float s = 0.0f;
  for (int blk = 0; blk < nblk; ++blk) s += partial[((size_t)blk * kK + k) * C + c];  // fixed order: deterministic
  grad_weights[t] = s;
}

int weight_blocks(int B, int M, int C) {
  const long long q = (long long)B * M;
  const int n_ctile = (C + 127) / 128;
  long long blk = (148 * 12) / n_ctile;
  if (blk > q) blk = q;
  return (int)(blk < 1 ? 1 : blk);
}

template <bool kBackward, bool kTensorCore>
int launch_rows(Args a, int B, cudaStream_t st) {
  const int n_ctile = (a.C + 127) / 128;
  if (n_ctile > kMaxCTiles) return D3D_ERR_UNSUPPORTED;
  // items per piece: the whole slot list in the forward pass, 64 inverse-map entries in the backward pass
  a.n_items_max = kBackward ? 64 : ((a.nsample + 15) & ~15);
  if (a.n_items_max > 256) return D3D_ERR_UNSUPPORTED;
  a.tmem_cols = 32;
  while (a.tmem_cols < ((a.n_items_max + 31) & ~31)) a.tmem_cols <<= 1;
  const size_t smem = (kTensorCore ? (size_t)n_ctile * kTileBytes + (size_t)a.n_items_max * 32 : (size_t)a.n_items_max * 64) +
                      (size_t)a.n_items_max * 16 + 16;
  auto kernel = pseudogrid_rows_kernel<kBackward, kTensorCore>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const int n_rows = kBackward ? a.N : a.M;
  const char* env_rows = getenv("D3D_PG_ROWS");
  // inverse-map segments vary a lot in length: fewer rows per CTA balance the backward pass better (measured 0.86 ->
  // 0.83 ms at level 0), the forward pass prefers to amortise the per-CTA weight staging over 8 rows
  a.rows_per_cta = env_rows ? atoi(env_rows) : (kBackward ? kRowsPerCta / 2 : kRowsPerCta);
  dim3 grid(d3d_ceil_div(n_rows, a.rows_per_cta), B);
  kernel<<<grid, kThreads, smem, st>>>(a);
  d3d_note_launches(1);
  return d3d_launch_status();
}

bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace

extern "C" {

int d3d_pseudogrid_fwd(const float* feat_cl, const float* query_xyz, const float* support_xyz, const int* idx,
                       const int* nvalid, const int* query_mask, const float* kpoints, const float* weights, int B,
                       int M, int N, int C, int nsample, int K, float extent, int influence, int precision,
                       float* out_cl, void* stream) {
  D3D_REQUIRE(feat_cl && query_xyz && support_xyz && idx && nvalid && query_mask && kpoints && weights && out_cl);
  D3D_REQUIRE(B >= 0 && M >= 0 && N > 0 && C > 0 && nsample > 0 && nsample <= D3D_MAX_NSAMPLE);
  D3D_REQUIRE(K > 0 && K <= kK && extent > 0.f && influence >= 0 && influence <= 2 && (precision == 0 || precision == 1));
  if (B == 0 || M == 0) return 0;
  Args a{};
  a.src = feat_cl; a.query_xyz = query_xyz; a.support_xyz = support_xyz; a.idx = idx; a.nvalid = nvalid;
  a.query_mask = query_mask; a.kpoints = kpoints; a.weights = weights; a.out = out_cl;
  a.M = M; a.N = N; a.C = C; a.nsample = nsample; a.K = K; a.influence = influence; a.extent = extent;
  return precision == 1 ? launch_rows<false, true>(a, B, (cudaStream_t)stream)
                        : launch_rows<false, false>(a, B, (cudaStream_t)stream);
}

size_t d3d_pseudogrid_bwd_workspace_bytes(int B, int M, int C, int K) {
  (void)K;
  if (B <= 0 || M <= 0 || C <= 0) return 0;
  return (size_t)weight_blocks(B, M, C) * kK * C * sizeof(float);
}

int d3d_pseudogrid_bwd(const float* grad_out_cl, const float* feat_cl, const float* query_xyz,
                       const float* support_xyz, const int* idx, const int* rowptr, const int* entries,
                       const int* nvalid, const int* query_mask, const float* kpoints, const float* weights, int B,
                       int M, int N, int C, int nsample, int K, float extent, int influence, int precision,
                       float* grad_feat_cl, float* grad_weights, void* ws, size_t ws_bytes, void* stream) {
  D3D_REQUIRE(grad_out_cl && feat_cl && query_xyz && support_xyz && idx && rowptr && entries && nvalid && query_mask);
  D3D_REQUIRE(kpoints && weights && (grad_feat_cl || grad_weights));
  D3D_REQUIRE(B >= 0 && M >= 0 && N > 0 && C > 0 && nsample > 0 && nsample <= D3D_MAX_NSAMPLE);
  D3D_REQUIRE(K > 0 && K <= kK && extent > 0.f && influence >= 0 && influence <= 2 && (precision == 0 || precision == 1));
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (grad_feat_cl) {
    Args a{};
    a.src = grad_out_cl; a.query_xyz = query_xyz; a.support_xyz = support_xyz; a.rowptr = rowptr; a.entries = entries;
    a.nvalid = nvalid; a.query_mask = query_mask; a.kpoints = kpoints; a.weights = weights; a.out = grad_feat_cl;
    a.M = M; a.N = N; a.C = C; a.nsample = nsample; a.K = K; a.influence = influence; a.extent = extent;
    const int rc = precision == 1 ? launch_rows<true, true>(a, B, st) : launch_rows<true, false>(a, B, st);
    if (rc != 0) return rc;
  }
  if (grad_weights) {
    if (M == 0) return (int)cudaMemsetAsync(grad_weights, 0, (size_t)K * C * sizeof(float), st);
    if (!ws || ws_bytes < d3d_pseudogrid_bwd_workspace_bytes(B, M, C, K)) return D3D_ERR_WORKSPACE;
    int nblk = weight_blocks(B, M, C);
    const size_t smem = (size_t)((nsample + 7) & ~7) * (kK * sizeof(float) + sizeof(int) + 3 * sizeof(float)) + 16;
    cudaError_t e = cudaFuncSetAttribute(pseudogrid_weight_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    dim3 grid((C + 127) / 128, nblk);
    if (precision == 1) {
      const size_t smem_tc = kDwBufs * kDwBufBytes + (size_t)kDwItems * (sizeof(int) + 3 * sizeof(float) + kK * sizeof(float)) + 16;
      // exactly one wave: every CTA walks its share of the rows start to finish, so a partial second wave would run
      // at a fraction of the machine (1776 CTAs at 5 resident per SM took 2.4 waves' worth of 3)
      const char* env_occ = getenv("D3D_PG_DW_CTAS_PER_SM");
      int per_sm = env_occ ? atoi(env_occ) : (int)((227 * 1024) / (smem_tc + 1024));
      if (per_sm > 8) per_sm = 8;  // __launch_bounds__(kThreads, 8)
      const int one_wave = (148 * (per_sm < 1 ? 1 : per_sm)) / (int)grid.x;
      if (one_wave >= 1 && one_wave < nblk) nblk = one_wave;
      grid.y = nblk;
      e = cudaFuncSetAttribute(pseudogrid_weight_grad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tc);
      if (e != cudaSuccess) return (int)e;
      pseudogrid_weight_grad_tc_kernel<<<grid, kThreads, smem_tc, st>>>(grad_out_cl, feat_cl, query_xyz, support_xyz, idx,
                                                                       nvalid, query_mask, kpoints, B, M, N, C, nsample, K,
                                                                       extent, influence, (float*)ws);
    } else {
      pseudogrid_weight_grad_kernel<<<grid, kThreads, smem, st>>>(grad_out_cl, feat_cl, query_xyz, support_xyz, idx, nvalid,
                                                                  query_mask, kpoints, B, M, N, C, nsample, K, extent,
                                                                  influence, (float*)ws);
    }
    reduce_partials_kernel<<<d3d_ceil_div((long long)K * C, 256), 256, 0, st>>>((const float*)ws, nblk, K, C, grad_weights);
    d3d_note_launches(2);
  }
  return d3d_launch_status();
}

}  // extern "C"

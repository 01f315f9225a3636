// PseudoGrid forward with the kernel-point contraction on the 5th-generation tensor cores (tcgen05 + TMEM).
//
//   ref: u_net_arch/models/local_aggregation_operators.py:467-503
//   out[b,j,c] = sum_m F[b, idx[b,j,m], c] * E[c, m],      E[c, m] = sum_k W[k, c] * w[b,j,k,m]
//
// E is a real dense GEMM: per query  E^T[C x ns] = W^T[C x K] . w[K x ns], K = 15 padded to 16 — exactly one
// tcgen05.mma K-step for bf16.  Mapping chosen so that the epilogue needs NO cross-thread reduction:
//   MMA M (TMEM lanes)   = 128 channels  (A operand = W^T tile, bf16, staged once per CTA)
//   MMA N (TMEM columns) = the ns neighbour slots of one query (B operand = influence weights, bf16)
//   accumulator in TMEM: lane = channel, column = slot, fp32
// After tcgen05.ld a thread owns ONE channel and all slots of the query: it streams the neighbours' features
// (the 32 lanes of a warp read 32 consecutive channels of the same row: one coalesced 128-byte request per slot,
// up to 32 requests in flight per thread) and accumulates sum_m F * E in a register.  The CUDA-core version
// (pseudogrid.cu) spends 16 FMAs per gathered element on E; here that is one MMA per (query, 128 channels) and the
// kernel is back on the gather-bandwidth roofline.  Features stay fp32; only W and the influence weights are
// rounded to bf16 (fp32 accumulate) — tolerance stated in tests/test_gpu_aggregation.py.
//
// Operand layout in shared memory: K-major, no swizzle (UMMA "interleave" canonical layout): 8-row x 16-byte core
// matrices; the two 8-element K chunks of a row group are LBO = 128 B apart, row groups SBO = 256 B apart.
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

constexpr int kThreads = 128;       // 4 warps = the 4 TMEM lane quarters
constexpr int kQueriesPerCta = 8;   // processed one after the other; the W^T tiles are staged once
constexpr int kTileBytes = 128 * 32;  // one 128-row x 16-element bf16 operand tile
constexpr unsigned kLbo = 128, kSbo = 256;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// byte offset of element (row, k) inside a K-major no-swizzle operand tile
__device__ __forceinline__ unsigned operand_offset(int row, int k) {
  return (unsigned)((row >> 3) * kSbo + (k >> 3) * kLbo + (row & 7) * 16 + (k & 7) * 2);
}

__device__ __forceinline__ unsigned long long make_smem_desc(unsigned addr) {
  return (unsigned long long)((addr >> 4) & 0x3fffu) | ((unsigned long long)(kLbo >> 4) << 16) |
         ((unsigned long long)(kSbo >> 4) << 32) | (1ull << 46);  // version 1 (sm_100), layout type 0 = no swizzle
}

__device__ __forceinline__ float influence_weight(float dx, float dy, float dz, float extent, int influence) {
  const float sq = dx * dx + dy * dy + dz * dz;
  if (influence == D3D_KP_LINEAR) return fmaxf(1.0f - sqrtf(sq) / extent, 0.0f);
  if (influence == D3D_KP_GAUSSIAN) {
    const float sigma = extent * 0.3f;
    return expf(-sq / (2.0f * sigma * sigma + 1e-9f));
  }
  return 1.0f;
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

__global__ void __launch_bounds__(kThreads)
pseudogrid_fwd_tc_kernel(const float* __restrict__ feat, const float* __restrict__ query_xyz,
                         const float* __restrict__ support_xyz, const int* __restrict__ idx,
                         const int* __restrict__ nvalid, const int* __restrict__ query_mask,
                         const float* __restrict__ kpoints, const float* __restrict__ weights, int M, int N, int C,
                         int nsample, int K, float extent, int influence, int n_mma, int tmem_cols,
                         float* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ unsigned tmem_base_slot;
  __shared__ float kp[48];

  const int tid = threadIdx.x, warp = tid >> 5;
  const int b = blockIdx.y;
  const int n_ctile = (C + 127) >> 7;
  unsigned char* a_tiles = smem;                                  // n_ctile x 4 KB
  unsigned char* b_tile = smem + (size_t)n_ctile * kTileBytes;    // n_mma x 32 B
  int* sidx = reinterpret_cast<int*>(b_tile + (size_t)n_mma * 32);

  // ---- one-time setup: TMEM columns, mbarrier, kernel points, W^T tiles (bf16, zero padded)
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"((unsigned)tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < K * 3) kp[tid] = kpoints[tid];
  for (int e = tid; e < n_ctile * 128 * 16; e += kThreads) {
    const int ct = e >> 11, k = (e >> 7) & 15, cl = e & 127;  // consecutive threads -> consecutive channels
    const int c = ct * 128 + cl;
    const float w = (c < C && k < K) ? weights[(size_t)k * C + c] : 0.0f;
    *reinterpret_cast<__nv_bfloat16*>(a_tiles + (size_t)ct * kTileBytes + operand_offset(cl, k)) = __float2bfloat16_rn(w);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem_base = tmem_base_slot;
  const unsigned bar = smem_u32(&mbar);
  unsigned phase = 0;

  // instruction descriptor: D = f32, A = B = bf16, both K-major, N = n_mma, M = 128
  const unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(n_mma >> 3) << 17) | ((128u >> 4) << 24);
  const unsigned long long b_desc = make_smem_desc(smem_u32(b_tile));
  const float* fb = feat + (size_t)b * N * C;

  for (int qi = 0; qi < kQueriesPerCta; ++qi) {
    const int j = blockIdx.x * kQueriesPerCta + qi;
    if (j >= M) break;  // block-uniform
    const size_t qrow = (size_t)b * M + j;
    const int n_eff = query_mask[qrow] != 0 ? nvalid[qrow] : nsample;
    const float qx = query_xyz[qrow * 3], qy = query_xyz[qrow * 3 + 1], qz = query_xyz[qrow * 3 + 2];
    // ---- stage the B operand: influence weights of this query's slots
    for (int m = tid; m < nsample; m += kThreads) sidx[m] = d3d_clamp_index(idx[qrow * nsample + m], N);
    __syncthreads();
    for (int t = tid; t < n_mma * 16; t += kThreads) {
      const int p = t >> 4, k = t & 15;
      float w = 0.0f;
      if (p < n_eff && k < K) {
        const float* s = support_xyz + ((size_t)b * N + sidx[p]) * 3;
        w = influence_weight((s[0] - qx) - kp[3 * k], (s[1] - qy) - kp[3 * k + 1], (s[2] - qz) - kp[3 * k + 2], extent,
                             influence);
      }
      *reinterpret_cast<__nv_bfloat16*>(b_tile + operand_offset(p, k)) = __float2bfloat16_rn(w);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();

    for (int ct = 0; ct < n_ctile; ++ct) {
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

      // ---- epilogue: this thread = channel c (TMEM lane tid); columns = slots
      const int c = ct * 128 + tid;
      const bool active = c < C;
      const float* fc = fb + (active ? c : 0);
      float acc = 0.0f;
      for (int col0 = 0; col0 < n_eff; col0 += 32) {
        unsigned r[32];
        const unsigned taddr = tmem_base + ((unsigned)(warp * 32) << 16) + (unsigned)col0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (active) {
          float x[32];
#pragma unroll
          for (int i = 0; i < 32; ++i)  // all gathers of the chunk are issued before the first use
            x[i] = (col0 + i < n_eff) ? __ldg(fc + (size_t)sidx[col0 + i] * C) : 0.0f;
#pragma unroll
          for (int i = 0; i < 32; ++i) acc += x[i] * __uint_as_float(r[i]);
        }
      }
      if (active) out[qrow * C + c] = acc;
      // every warp is done reading TMEM before the next MMA overwrites it
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
    }
  }

  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((unsigned)tmem_cols) : "memory");
}

}  // namespace

int d3d_pseudogrid_fwd_tc(const float* feat_cl, const float* query_xyz, const float* support_xyz, const int* idx,
                          const int* nvalid, const int* query_mask, const float* kpoints, const float* weights, int B,
                          int M, int N, int C, int nsample, int K, float extent, int influence, float* out_cl,
                          cudaStream_t st) {
  const int n_mma = (nsample + 15) & ~15;  // MMA N: multiple of 16 for M = 128
  if (n_mma > 256) return D3D_ERR_UNSUPPORTED;
  int tmem_cols = 32;
  while (tmem_cols < ((n_mma + 31) & ~31)) tmem_cols <<= 1;
  const int n_ctile = (C + 127) / 128;
  const size_t smem = (size_t)n_ctile * kTileBytes + (size_t)n_mma * 32 + (size_t)nsample * sizeof(int) + 16;
  if (smem > 200 * 1024) return D3D_ERR_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(pseudogrid_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  dim3 grid(d3d_ceil_div(M, kQueriesPerCta), B);
  pseudogrid_fwd_tc_kernel<<<grid, kThreads, smem, st>>>(feat_cl, query_xyz, support_xyz, idx, nvalid, query_mask, kpoints,
                                                         weights, M, N, C, nsample, K, extent, influence, n_mma, tmem_cols,
                                                         out_cl);
  d3d_note_launches(1);
  return d3d_launch_status();
}

// The 1x1 convolutions of the U-Net as row-major GEMMs on the 5th-generation tensor cores: TMA-fed, tcgen05.mma
// kind::tf32, accumulator in TMEM, BatchNorm statistics in the epilogue.
//
//   ref: u_net_arch/models/backbones/resnet.py:32-45,58-66; models/local_aggregation_operators.py:121-123;
//        models/heads/multi_dimensional_head.py:40-59  (nn.Conv1d(kernel_size=1, bias=False) -> BatchNorm1d -> ReLU;
//        cuDNN runs them in TF32 by default, torch.backends.cudnn.allow_tf32)
//
//   C[M x N] = A[M x K] . B[N x K]^T        A = activation rows (channel-last), B = conv weight (Cout x Cin), fp32 in / out
//
// * operands travel global -> shared with TMA (cp.async.bulk.tensor.2d, 128-byte swizzle: a [rows x 32 floats] box IS the
//   canonical K-major SWIZZLE_128B operand tile), completion on mbarriers; a 4-stage ring, one producer thread;
// * one thread issues tcgen05.mma (M = 128, N = column tile <= 144, K = 8 per instruction, 4 per stage), releases stages
//   with tcgen05.commit; the accumulator lives in TMEM;
// * four epilogue warps read TMEM (thread = output row), store the rows, and — for a convolution that feeds a BatchNorm —
//   reduce per-column (count, mean, M2) partials of the tile through shared memory: the BatchNorm statistics pass over
//   the convolution output (one full read of it) disappears, d3d_bn_finalize combines the tile partials in fp64
//   (Chan's parallel variance), d3d_bn_apply_cl applies them;
// * A may be the channel concatenation of TWO row tensors (the decoder's skip connections,
//   heads/multi_dimensional_head.py:36) without materialising it: the K loop walks both tensor maps.
#include <cuda.h>

#include "common.cuh"
#include "umma.cuh"

namespace {

using namespace umma;

constexpr int kBM = 128;           // rows per CTA = MMA M = TMEM lanes
constexpr int kBK = 32;            // floats per stage row = 128 bytes = one swizzle row
constexpr int kBNMax = 144;        // column tile (multiple of 16): 72 -> 80, 144, 288 = 2 x 144, ...
constexpr int kStages = 4;         // 4 x 34 KB ring + a 74 KB output staging tile: one persistent CTA per SM
constexpr int kThreads = 320;      // warp 0: TMA producer, warp 1: TMEM allocator + MMA issuer, warps 2-9: epilogue
constexpr int kEpiThreads = kThreads - 64;
constexpr int kMaxSlots = 8;       // row slots of the store loop (partial column sums per slot in shared memory)
constexpr unsigned kABytes = kBM * kBK * 4;            // 16 KB
constexpr unsigned kBBytes = kBNMax * kBK * 4;         // 18 KB
constexpr unsigned kStageBytes = kABytes + kBBytes;    // 34 KB, a multiple of 1024
constexpr int kCStride = kBNMax + 4;                   // fp32 staging of the output tile for the column statistics

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

// row-major fp32 matrix [rows x cols], leading dimension ld (floats); box = [box_rows x 32 floats], 128-byte swizzle;
// out-of-range elements read as zero (tails in M, N and K need no special code)
int make_map(CUtensorMap* m, const float* p, long long rows, long long cols, long long ld, int box_rows,
             CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return D3D_ERR_UNSUPPORTED;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : D3D_ERR_BAD_ARG;
}

__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}

// K-major operand tile in the 128-byte-swizzle canonical layout: rows 128 bytes apart, 8-row groups 1024 bytes apart
__device__ __forceinline__ unsigned long long sw128_desc(unsigned addr) {
  return (unsigned long long)((addr >> 4) & 0x3fffu) | (1ull << 16) | ((unsigned long long)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}

__device__ __forceinline__ void mma_tf32(unsigned tmem_d, unsigned long long a_desc, unsigned long long b_desc, unsigned idesc,
                                         unsigned accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

struct GemmArgs {
  float* C;
  float* stats;   // (row tiles, 2, N): per-tile column mean and M2 (sum of squared deviations), or null
  int M, N, K0, K1;  // A = [A0 (M x K0) | A1 (M x K1)]
  int bn;            // column tile: multiple of 16, <= kBNMax
  int accumulate;    // C += instead of C =
  int tiles_m, tiles_n;
  // inference epilogue (eval-mode BatchNorm folded into the convolution): C = act(acc + bias[col] + residual[row, col])
  const float* bias;      // (N) or null
  const float* residual;  // (M, N) row-major like C, or null
  int relu;
};

__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Persistent: one CTA per SM walks output tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...  The three roles run
// decoupled — the producer streams stages of whatever tile comes next, the MMA thread fills one of TWO TMEM accumulators
// while the epilogue warps drain the other — so loads, tensor-core work and stores of consecutive tiles overlap.
template <bool kAct>  // kAct: the inference epilogue (bias / residual / ReLU) — compiled out of the training kernel
__global__ void __launch_bounds__(kThreads, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                 const __grid_constant__ CUtensorMap map_b, const GemmArgs g) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // the dynamic shared-memory window is only guaranteed 16-byte aligned: align the stage ring to 1024 by hand
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* sC = reinterpret_cast<float*>(smem + (size_t)kStages * kStageBytes);  // fp32 staging of one output tile
  __shared__ __align__(8) unsigned long long bars[2 * kStages + 4];
  __shared__ unsigned tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nk0 = (g.K0 + kBK - 1) / kBK, nk1 = (g.K1 + kBK - 1) / kBK, nk = nk0 + nk1;
  const int n_tiles = g.tiles_m * g.tiles_n;
  const unsigned full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[kStages]);
  const unsigned tfull0 = smem_u32(&bars[2 * kStages]), tempty0 = smem_u32(&bars[2 * kStages + 2]);

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, kEpiThreads); }
    mbar_init_fence();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), 512u);  // two accumulators of up to 256 columns
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_base = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {  // ---- TMA producer ----
      const unsigned tx = kABytes + (unsigned)g.bn * kBK * 4;
      int it_all = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int m0 = (t / g.tiles_n) * kBM, n0 = (t % g.tiles_n) * g.bn;
        for (int it = 0; it < nk; ++it, ++it_all) {
          const int s = it_all % kStages;
          mbar_wait_short(empty0 + 8 * s, (unsigned)(((it_all / kStages) & 1) ^ 1));
          const unsigned sa = smem_u32(smem + (size_t)s * kStageBytes), sb = sa + kABytes;
          mbar_arrive_expect_tx(full0 + 8 * s, tx);
          if (it < nk0) tma_load_2d(sa, &map_a0, it * kBK, m0, full0 + 8 * s);
          else tma_load_2d(sa, &map_a1, (it - nk0) * kBK, m0, full0 + 8 * s);
          tma_load_2d(sb, &map_b, it < nk0 ? it * kBK : g.K0 + (it - nk0) * kBK, n0, full0 + 8 * s);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---- MMA issuer ----
      // D = f32, A = B = tf32, both K-major, N = bn, M = 128
      const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(g.bn >> 3) << 17) | ((128u >> 4) << 24);
      int it_all = 0, n_local = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++n_local) {
        const int acc = n_local & 1;
        mbar_wait_short(tempty0 + 8 * acc, (unsigned)(((n_local >> 1) & 1) ^ 1));  // the epilogue has drained this accumulator
        tc_fence_after();
        const unsigned d = tmem_base + (unsigned)acc * 256u;
        for (int it = 0; it < nk; ++it, ++it_all) {
          const int s = it_all % kStages;
          mbar_wait_short(full0 + 8 * s, (unsigned)((it_all / kStages) & 1));
          tc_fence_after();
          const unsigned sa = smem_u32(smem + (size_t)s * kStageBytes), sb = sa + kABytes;
#pragma unroll
          for (int kk = 0; kk < kBK / 8; ++kk)
            mma_tf32(d, sw128_desc(sa + kk * 32), sw128_desc(sb + kk * 32), idesc, (it > 0 || kk > 0) ? 1u : 0u);
          mma_commit(empty0 + 8 * s);  // the stage is free once these MMAs have read it
        }
        mma_commit(tfull0 + 8 * acc);
      }
    }
  } else {
    // ---- epilogue (8 warps): TMEM -> shared memory (thread = output row, TMEM lane = 32 * (warp % 4) + lane; the two
    //      warps of a lane quarter take alternate 16-column pieces), then the tile goes out as contiguous row segments —
    //      a thread keeps ONE float4 column and walks the rows, so the column statistics accumulate in its registers ----
    const int q = warp & 3, r = q * 32 + lane, half = (warp - 2) >> 2;
    const int te = tid - 64;
    float* sP = sC + (size_t)kBM * kCStride;  // [row slot][2][kBNMax] partial column sums
    int n_local = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++n_local) {
      const int acc = n_local & 1;
      const int tile_m = t / g.tiles_n, m0 = tile_m * kBM, n0 = (t % g.tiles_n) * g.bn;
      mbar_wait_short(tfull0 + 8 * acc, (unsigned)((n_local >> 1) & 1));
      tc_fence_after();
      for (int c16 = 16 * half; c16 < g.bn; c16 += 32) {
        unsigned v[16];
        tmem_ld16(tmem_base + (unsigned)acc * 256u + ((unsigned)(q * 32) << 16) + (unsigned)c16, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(sC + (size_t)r * kCStride + c16 + 4 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
      tc_fence_before();
      mbar_arrive(tempty0 + 8 * acc);  // accumulator read out: the MMA thread may start the tile after next in it
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");  // staging tile complete
      const int rows_valid = min(kBM, g.M - m0);
      const int nc4 = min(g.bn, g.N - n0) >> 2;  // N % 4 == 0
      const int rows_per_pass = min(kEpiThreads / nc4, kMaxSlots);
      const int slot = te / nc4, c4 = te - slot * nc4;
      float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
      if (slot < rows_per_pass) {
        const float4 shift = *reinterpret_cast<const float4*>(sC + 4 * c4);  // row 0 of the tile: common shift of the sums
        float* cbase = g.C + (long long)m0 * g.N + n0 + 4 * c4;
        const float4 bias4 = (kAct && g.bias) ? __ldg(reinterpret_cast<const float4*>(g.bias + n0 + 4 * c4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float* rbase = (kAct && g.residual) ? g.residual + (long long)m0 * g.N + n0 + 4 * c4 : nullptr;
        const float* sbase = sC + 4 * c4;
        for (int rr = slot; rr < rows_valid; rr += 4 * rows_per_pass) {
          float4 o[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (rr + u * rows_per_pass < rows_valid) o[u] = *reinterpret_cast<const float4*>(sbase + (size_t)(rr + u * rows_per_pass) * kCStride);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int row = rr + u * rows_per_pass;
            if (row < rows_valid) {
              float4 w = o[u];
              float* dst = cbase + (long long)row * g.N;
              if (g.accumulate) {
                const float4 old = *reinterpret_cast<const float4*>(dst);
                w.x += old.x; w.y += old.y; w.z += old.z; w.w += old.w;
              }
              if (kAct && g.bias) { w.x += bias4.x; w.y += bias4.y; w.z += bias4.z; w.w += bias4.w; }
              if (kAct && rbase) {
                const float4 rs = __ldg(reinterpret_cast<const float4*>(rbase + (long long)row * g.N));
                w.x += rs.x; w.y += rs.y; w.z += rs.z; w.w += rs.w;
              }
              if (kAct && g.relu) { w.x = fmaxf(w.x, 0.f); w.y = fmaxf(w.y, 0.f); w.z = fmaxf(w.z, 0.f); w.w = fmaxf(w.w, 0.f); }
              *reinterpret_cast<float4*>(dst) = w;
              const float dx = o[u].x - shift.x, dy = o[u].y - shift.y, dz = o[u].z - shift.z, dw = o[u].w - shift.w;
              s1.x += dx; s1.y += dy; s1.z += dz; s1.w += dw;
              s2.x += dx * dx; s2.y += dy * dy; s2.z += dz * dz; s2.w += dw * dw;
            }
          }
        }
        if (g.stats) {
          *reinterpret_cast<float4*>(sP + (size_t)slot * 2 * kBNMax + 4 * c4) = s1;
          *reinterpret_cast<float4*>(sP + (size_t)slot * 2 * kBNMax + kBNMax + 4 * c4) = s2;
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");  // staging tile consumed; partial sums in place
      if (g.stats && te < 4 * nc4) {
        float a1 = 0.f, a2 = 0.f;
        for (int p = 0; p < rows_per_pass; ++p) {  // fixed order: deterministic
          a1 += sP[(size_t)p * 2 * kBNMax + te];
          a2 += sP[(size_t)p * 2 * kBNMax + kBNMax + te];
        }
        const float inv = 1.0f / (float)rows_valid;
        float* dst = g.stats + (size_t)tile_m * 2 * g.N + n0 + te;
        dst[0] = sC[te] + a1 * inv;        // tile mean (sC row 0 = the shift; the next tile's staging starts after bar 2)
        dst[g.N] = a2 - a1 * a1 * inv;     // tile M2
      }
      asm volatile("bar.sync 2, %0;" ::"n"(kEpiThreads) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512u);
}

// BatchNorm statistics from the GEMM's tile partials (count, mean, M2 per 128-row tile), combined in fp64:
//   mean = sum n_t mean_t / n,   M2 = sum (M2_t + n_t (mean_t - mean)^2)      (Chan et al., pairwise form summed up)
// Block = 8 channels x 32 tile slices (many small blocks: the kernel is pure latency), four loads in flight per thread;
// the slices are combined in a fixed order (deterministic).
constexpr int kFinCh = 8, kFinSlices = 32;

__device__ __forceinline__ double fin_reduce(double (*red)[kFinCh], int cx, int ty, double v) {
  red[ty][cx] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < kFinSlices; ++k) s += red[k][cx];
  __syncthreads();
  return s;
}

__global__ void __launch_bounds__(kFinCh * kFinSlices)
bn_finalize_kernel(const float* __restrict__ stats, int tiles, long long R, int C, float eps, float momentum,
                   float* __restrict__ running_mean, float* __restrict__ running_var, long long* __restrict__ num_batches,
                   float* __restrict__ save_mean, float* __restrict__ save_invstd) {
  __shared__ double red[kFinSlices][kFinCh];
  const int cx = threadIdx.x % kFinCh, ty = threadIdx.x / kFinCh;
  const int c = min(blockIdx.x * kFinCh + cx, C - 1);  // surplus lanes of the last block repeat channel C - 1
  const double n = (double)R;
  const double last_n = (double)(R - (long long)(tiles - 1) * kBM);  // rows of the last tile; all others hold kBM
  const float* pm = stats + c;
  // pass 1: mean = sum n_t mean_t / n
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int t = ty;
  for (; t + 3 * kFinSlices < tiles - 1; t += 4 * kFinSlices) {
    const float v0 = pm[(size_t)t * 2 * C], v1 = pm[(size_t)(t + kFinSlices) * 2 * C], v2 = pm[(size_t)(t + 2 * kFinSlices) * 2 * C],
                v3 = pm[(size_t)(t + 3 * kFinSlices) * 2 * C];
    a0 += v0; a1 += v1; a2 += v2; a3 += v3;
  }
  double acc = (a0 + a1 + a2 + a3) * (double)kBM;
  for (; t < tiles; t += kFinSlices) acc += (t == tiles - 1 ? last_n : (double)kBM) * (double)pm[(size_t)t * 2 * C];
  const double mean = fin_reduce(red, cx, ty, acc) / n;
  // pass 2: M2 = sum (M2_t + n_t (mean_t - mean)^2)
  a0 = a1 = a2 = a3 = 0.0;
  t = ty;
  for (; t + 3 * kFinSlices < tiles - 1; t += 4 * kFinSlices) {
    const float* p0 = pm + (size_t)t * 2 * C;
    const float* p1 = pm + (size_t)(t + kFinSlices) * 2 * C;
    const float* p2 = pm + (size_t)(t + 2 * kFinSlices) * 2 * C;
    const float* p3 = pm + (size_t)(t + 3 * kFinSlices) * 2 * C;
    const float m0 = p0[0], q0 = p0[C], m1 = p1[0], q1 = p1[C], m2 = p2[0], q2 = p2[C], m3 = p3[0], q3 = p3[C];
    const double d0 = (double)m0 - mean, d1 = (double)m1 - mean, d2 = (double)m2 - mean, d3 = (double)m3 - mean;
    a0 += (double)q0 + (double)kBM * d0 * d0; a1 += (double)q1 + (double)kBM * d1 * d1;
    a2 += (double)q2 + (double)kBM * d2 * d2; a3 += (double)q3 + (double)kBM * d3 * d3;
  }
  acc = a0 + a1 + a2 + a3;
  for (; t < tiles; t += kFinSlices) {
    const double d = (double)pm[(size_t)t * 2 * C] - mean;
    acc += (double)pm[(size_t)t * 2 * C + C] + (t == tiles - 1 ? last_n : (double)kBM) * d * d;
  }
  const double m2 = fin_reduce(red, cx, ty, acc);
  if (ty != 0 || blockIdx.x * kFinCh + cx >= C) return;
  const double var = m2 / n > 0 ? m2 / n : 0.0;
  save_mean[c] = (float)mean;
  save_invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) {
    const double unbiased = n > 1 ? m2 / (n - 1) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
  if (c == 0 && num_batches) *num_batches += 1;
}

// ---------------------------------------------------------------------------------------------------------------
// Weight gradient  dW[Cout x Cin] (+)= dY[R x Cout]^T . X[R x Cin]   (the reduction runs over the R rows)
//
// Both operands are "MN-major" for the tensor core: the M index (Cout) and the N index (Cin) are the contiguous ones
// in memory, the K index (row) strides.  For 32-bit MN-major operands tcgen05 knows ONE shared-memory layout, the
// 128-byte swizzle with 32-byte atomicity (descriptor layout type 1; TMA: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): rows (K)
// 128 bytes apart, the four 32-byte chunks of a row permuted by (row mod 4), 4-row groups 512 bytes apart (SBO), the
// next 32 channels one TMA box further (LBO).  A box of [32 rows x 32 channels] is one column block of that tile.  M tile = 128 output channels = 4 boxes, N tile <= 160 input
// channels = 5 boxes, one stage = 32 rows.  The rows are split over the CTAs (split-K): every CTA accumulates its row
// range in TMEM and writes an fp32 partial; wgrad_reduce_kernel adds the partials in a fixed order (deterministic) and
// stores or accumulates into the gradient buffer.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kWBK = 32;                                  // rows per stage
constexpr int kWBoxBytes = kWBK * 128;                    // one [32 rows x 32 channels] box
constexpr int kWNMaxBoxes = 5;                            // N tile <= 160 input channels
constexpr unsigned kWStageBytes = (4 + kWNMaxBoxes) * kWBoxBytes;  // 36 KB
constexpr int kWStages = 3;                               // 3 x 36 KB: two CTAs per SM
constexpr int kWThreads = 192;                            // warp 0 producer, warp 1 MMA, warps 2-5 epilogue

int make_map_rows(CUtensorMap* m, const float* p, long long rows, long long cols) {  // box = [kWBK rows x 32 floats]
  return make_map(m, p, rows, cols, cols, kWBK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
}

// MN-major 32-bit operand, SWIZZLE_128B_BASE32B: LBO = byte distance between 32-channel column blocks, SBO = 512
// (4 rows of 128 bytes)
__device__ __forceinline__ unsigned long long sw128_mn_desc(unsigned addr, unsigned lbo) {
  return (unsigned long long)((addr >> 4) & 0x3fffu) | ((unsigned long long)((lbo >> 4) & 0x3fffu) << 16) |
         ((unsigned long long)(512 >> 4) << 32) | (1ull << 46) | (1ull << 61);
}

struct WgradArgs {
  float* partial;       // (splits, Mdim, Ndim)
  long long R, rows_per_split;
  int Cout, Cin, bn;    // here: Cout = the M-side channel count, Cin = the N-side one (the host may swap dY and X so that
                        // the operand with fewer 128-channel tiles sits on the M side); bn: N tile (multiple of 16, <= 160)
  int n_boxes;          // 32-channel boxes per N tile
  float* direct;        // one split only: the epilogue writes dW itself (no partials, no reduction kernel)
  int direct_transposed, direct_accumulate;
};

__global__ void __launch_bounds__(kWThreads, 2)
wgrad_tf32_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x, const WgradArgs g) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) unsigned long long bars[2 * kWStages + 1];
  __shared__ unsigned tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * g.bn, split = blockIdx.z;
  const long long r_beg = (long long)split * g.rows_per_split;
  const long long r_end = r_beg + g.rows_per_split < g.R ? r_beg + g.rows_per_split : g.R;
  const int nk = (int)((r_end - r_beg + kWBK - 1) / kWBK);
  const unsigned full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[kWStages]), tmem_full = smem_u32(&bars[2 * kWStages]);
  const unsigned tmem_cols = g.bn <= 32 ? 32u : (g.bn <= 64 ? 64u : (g.bn <= 128 ? 128u : 256u));

  if (tid == 0) {
    for (int s = 0; s < kWStages; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    mbar_init(tmem_full, 1);
    mbar_init_fence();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dy) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_base = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {  // ---- TMA producer: 4 boxes of dY (128 output channels) + n_boxes of X per stage ----
      const unsigned tx = (unsigned)(4 + g.n_boxes) * kWBoxBytes;
      for (int it = 0; it < nk; ++it) {
        const int s = it % kWStages;
        mbar_wait_short(empty0 + 8 * s, (unsigned)(((it / kWStages) & 1) ^ 1));
        const unsigned sa = smem_u32(smem + (size_t)s * kWStageBytes), sb = sa + 4 * kWBoxBytes;
        mbar_arrive_expect_tx(full0 + 8 * s, tx);
        // rows beyond this split's range are NOT read: the box is clipped by loading a second, zero... (ranges are
        // multiples of kWBK except the last split, whose tail rows are out of bounds of the tensor and read as zero)
        const int row = (int)(r_beg + (long long)it * kWBK);
#pragma unroll
        for (int b = 0; b < 4; ++b) tma_load_2d(sa + b * kWBoxBytes, &map_dy, m0 + 32 * b, row, full0 + 8 * s);
        for (int b = 0; b < g.n_boxes; ++b) tma_load_2d(sb + b * kWBoxBytes, &map_x, n0 + 32 * b, row, full0 + 8 * s);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---- MMA issuer ----
      // D = f32, A = B = tf32, both MN-major, N = bn, M = 128
      const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((unsigned)(g.bn >> 3) << 17) |
                             ((128u >> 4) << 24);
      for (int it = 0; it < nk; ++it) {
        const int s = it % kWStages;
        mbar_wait_short(full0 + 8 * s, (unsigned)((it / kWStages) & 1));
        tc_fence_after();
        const unsigned sa = smem_u32(smem + (size_t)s * kWStageBytes), sb = sa + 4 * kWBoxBytes;
#pragma unroll
        for (int kk = 0; kk < kWBK / 8; ++kk)
          mma_tf32(tmem_base, sw128_mn_desc(sa + kk * 1024, kWBoxBytes), sw128_mn_desc(sb + kk * 1024, kWBoxBytes), idesc,
                   (it > 0 || kk > 0) ? 1u : 0u);
        mma_commit(empty0 + 8 * s);
      }
      mma_commit(tmem_full);
    }
  } else {
    // ---- epilogue: thread = output channel (TMEM lane); partial[split][cout][cin] ----
    mbar_wait_short(tmem_full, 0);
    tc_fence_after();
    const int q = warp & 3, r = q * 32 + lane;
    const int cout = m0 + r;
    float* prow = g.partial + ((size_t)split * g.Cout + (cout < g.Cout ? cout : 0)) * g.Cin + n0;
    for (int c16 = 0; c16 < g.bn; c16 += 16) {
      unsigned v[16];
      if (nk > 0) {
        tmem_ld16(tmem_base + ((unsigned)(q * 32) << 16) + (unsigned)c16, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0u;
      }
      if (g.direct != nullptr) {
        if (cout < g.Cout) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int n = n0 + c16 + j;
            if (n < g.Cin) {
              float* dst = g.direct_transposed ? g.direct + (size_t)n * g.Cout + cout : g.direct + (size_t)cout * g.Cin + n;
              *dst = (g.direct_accumulate ? *dst : 0.f) + __uint_as_float(v[j]);
            }
          }
        }
      } else if (cout < g.Cout) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n0 + c16 + 4 * j < g.Cin)  // Cin % 4 == 0
            *reinterpret_cast<uint4*>(prow + c16 + 4 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// Sum of the split partials into dW, fixed order (deterministic).  Block = 64 consecutive elements x 4 split lanes: a
// lane adds every fourth partial (coalesced 256-byte rows), the four lane sums are combined through shared memory.
// transposed: the partials are (Cin x Cout) — dY and X were swapped in the GEMM — while dW is (Cout x Cin).
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int Cout, int Cin, int transposed, int accumulate,
                    float* __restrict__ dw) {
  __shared__ float sh[4][64];
  const long long n = (long long)Cout * Cin;
  const int e = threadIdx.x & 63, sl = threadIdx.x >> 6;
  const long long i = (long long)blockIdx.x * 64 + e;  // index into the partial's own layout
  float a0 = 0.f, a1 = 0.f;
  if (i < n) {
    int s = sl;
    for (; s + 4 < splits; s += 8) {
      a0 += __ldg(partial + (size_t)s * n + i);
      a1 += __ldg(partial + (size_t)(s + 4) * n + i);
    }
    if (s < splits) a0 += __ldg(partial + (size_t)s * n + i);
  }
  sh[sl][e] = a0 + a1;
  __syncthreads();
  if (sl != 0 || i >= n) return;
  const float v = (sh[0][e] + sh[1][e]) + (sh[2][e] + sh[3][e]);
  long long o = i;
  if (transposed) {  // i = ci * Cout + co
    const long long ci = i / Cout, co = i - ci * Cout;
    o = co * Cin + ci;
  }
  dw[o] = (accumulate ? dw[o] : 0.f) + v;
}

struct WgradPlan {
  bool swapped;   // X on the M side, dY on the N side (the partials are then Cin x Cout)
  int m_dim, n_dim, bn, splits;
};

WgradPlan wgrad_plan(long long R, int Cout, int Cin) {
  auto bn_of = [](int n) { const int n16 = (n + 15) & ~15; return n16 < 160 ? n16 : 160; };
  // floats read per row of the reduction: every N tile re-reads the M operand and vice versa
  auto traffic = [&](int m, int n) { return (long long)((n + bn_of(n) - 1) / bn_of(n)) * m + (long long)((m + kBM - 1) / kBM) * n; };
  WgradPlan p;
  p.swapped = traffic(Cin, Cout) < traffic(Cout, Cin);
  p.m_dim = p.swapped ? Cin : Cout;
  p.n_dim = p.swapped ? Cout : Cin;
  p.bn = bn_of(p.n_dim);
  const long long tiles = (long long)((p.m_dim + kBM - 1) / kBM) * ((p.n_dim + p.bn - 1) / p.bn);
  long long s = (148 * 2 + tiles - 1) / tiles;              // about two CTAs per SM
  const long long cap = (16LL << 20) / (4LL * Cout * Cin);  // the partials are re-read by the reduction: <= 16 MB of them
  if (s > cap) s = cap;
  const long long max_s = (R + 4 * kWBK - 1) / (4 * kWBK);  // at least four stages per split
  if (s > max_s) s = max_s;
  if (s < 1 || tiles >= 96) s = 1;  // enough tiles to fill the machine: one split, the epilogue writes dW directly
  p.splits = (int)s;
  return p;
}

bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace

extern "C" {

int d3d_gemm_tf32_act(const float* A0, const float* A1, const float* B, float* C, long long M, int N, int K0, int K1,
                      int accumulate, float* stats, const float* bias, const float* residual, int relu, void* stream);


int d3d_gemm_row_tiles(long long M) { return (int)((M + kBM - 1) / kBM); }

/* C[M x N] (+)= [A0 | A1] . B^T ;  A0 (M x K0), A1 (M x K1) or NULL, B (N x (K0 + K1)), all row-major fp32, TF32 tensor cores.
 * stats (d3d_gemm_row_tiles(M), 2, N) or NULL: per row tile the column mean and M2 of the tile of C (for d3d_bn_finalize).
 * Requires K0 % 4 == 0, K1 % 4 == 0, N % 4 == 0 and 16-byte aligned pointers (TMA). */
int d3d_gemm_tf32(const float* A0, const float* A1, const float* B, float* C, long long M, int N, int K0, int K1, int accumulate,
                  float* stats, void* stream) {
  return d3d_gemm_tf32_act(A0, A1, B, C, M, N, K0, K1, accumulate, stats, nullptr, nullptr, 0, stream);
}

/* The same GEMM with the inference epilogue C = act(A . B^T + bias[col] + residual[row, col]): a convolution whose
 * eval-mode BatchNorm has been folded into its weights (B scaled per output channel, bias = the BatchNorm shift), the
 * bottleneck's residual add and the ReLU ride on the store.  bias (N), residual (M, N) may be NULL. */
int d3d_gemm_tf32_act(const float* A0, const float* A1, const float* B, float* C, long long M, int N, int K0, int K1,
                      int accumulate, float* stats, const float* bias, const float* residual, int relu, void* stream) {
  D3D_REQUIRE(A0 && B && C && M > 0 && N > 0 && K0 > 0 && K1 >= 0 && (K1 == 0 || A1));
  if (K0 % 4 || K1 % 4 || N % 4 || !aligned16(A0) || !aligned16(A1) || !aligned16(B) || !aligned16(C)) return D3D_ERR_UNSUPPORTED;
  if (!aligned16(bias) || !aligned16(residual)) return D3D_ERR_UNSUPPORTED;
  if (stats && (accumulate || bias || residual || relu)) return D3D_ERR_BAD_ARG;
  GemmArgs g{};
  g.C = C; g.stats = stats; g.M = (int)M; g.N = N; g.K0 = K0; g.K1 = K1; g.accumulate = accumulate;
  g.bias = bias; g.residual = residual; g.relu = relu;
  if (M > 0x7fffffffLL) return D3D_ERR_UNSUPPORTED;
  const int n16 = (N + 15) & ~15;
  g.bn = n16 < kBNMax ? n16 : kBNMax;
  g.tiles_m = (int)((M + kBM - 1) / kBM);
  // few row tiles (the coarse levels): narrower column tiles put more SMs to work
  for (int cand : {96, 80, 48})
    if ((long long)g.tiles_m * ((N + g.bn - 1) / g.bn) < 120 && cand < g.bn) g.bn = cand;
  g.tiles_n = (N + g.bn - 1) / g.bn;
  CUtensorMap ma0, ma1, mb;
  int rc = make_map(&ma0, A0, M, K0, K0, kBM);
  if (rc == 0) rc = K1 > 0 ? make_map(&ma1, A1, M, K1, K1, kBM) : make_map(&ma1, A0, M, K0, K0, kBM);
  if (rc == 0) rc = make_map(&mb, B, N, (long long)K0 + K1, (long long)K0 + K1, g.bn);
  if (rc != 0) return rc;
  const size_t smem = (size_t)kStages * kStageBytes + (size_t)kBM * kCStride * 4 + (size_t)kMaxSlots * 2 * kBNMax * 4 + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tf32_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tf32_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr_done = true;
  }
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (n_sm <= 0) n_sm = 148;
  }
  const long long n_tiles = (long long)g.tiles_m * g.tiles_n;
  const unsigned grid = (unsigned)(n_tiles < n_sm ? n_tiles : n_sm);
  if (bias || residual || relu) gemm_tf32_kernel<true><<<grid, kThreads, smem, (cudaStream_t)stream>>>(ma0, ma1, mb, g);
  else gemm_tf32_kernel<false><<<grid, kThreads, smem, (cudaStream_t)stream>>>(ma0, ma1, mb, g);
  d3d_note_launches(1);
  return d3d_launch_status();
}

/* dW (Cout x Cin) = or += dY (R x Cout)^T . X (R x Cin), fp32 in / out, TF32 tensor cores, split over the rows with a
 * deterministic reduction.  Workspace: d3d_wgrad_workspace_bytes.  Requires Cout % 4 == 0, Cin % 4 == 0, aligned pointers. */
size_t d3d_wgrad_workspace_bytes(long long R, int Cout, int Cin) {
  if (R <= 0 || Cout <= 0 || Cin <= 0) return 0;
  return (size_t)wgrad_plan(R, Cout, Cin).splits * Cout * Cin * sizeof(float);
}

int d3d_wgrad_tf32(const float* dY, const float* X, float* dW, long long R, int Cout, int Cin, int accumulate, void* ws,
                   size_t ws_bytes, void* stream) {
  D3D_REQUIRE(dY && X && dW && R > 0 && Cout > 0 && Cin > 0);
  if (Cout % 4 || Cin % 4 || !aligned16(dY) || !aligned16(X) || !aligned16(dW) || R > 0x7fffffffLL) return D3D_ERR_UNSUPPORTED;
  if (!ws || ws_bytes < d3d_wgrad_workspace_bytes(R, Cout, Cin)) return D3D_ERR_WORKSPACE;
  const WgradPlan plan = wgrad_plan(R, Cout, Cin);
  WgradArgs g{};
  g.bn = plan.bn;
  g.n_boxes = (g.bn + 31) / 32;
  g.Cout = plan.m_dim; g.Cin = plan.n_dim; g.R = R; g.partial = (float*)ws;
  long long rps = (R + plan.splits - 1) / plan.splits;
  rps = (rps + kWBK - 1) / kWBK * kWBK;  // whole stages, so that a split never reads another split's rows
  g.rows_per_split = rps;
  const int used = (int)((R + rps - 1) / rps);
  if (used == 1) {
    g.direct = dW;
    g.direct_transposed = plan.swapped ? 1 : 0;
    g.direct_accumulate = accumulate;
  }
  CUtensorMap mm, mn;
  int rc = make_map_rows(&mm, plan.swapped ? X : dY, R, plan.m_dim);
  if (rc == 0) rc = make_map_rows(&mn, plan.swapped ? dY : X, R, plan.n_dim);
  if (rc != 0) return rc;
  const size_t smem = (size_t)kWStages * kWStageBytes + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr_done = true;
  }
  dim3 grid((unsigned)((plan.m_dim + kBM - 1) / kBM), (unsigned)((plan.n_dim + g.bn - 1) / g.bn), (unsigned)used);
  wgrad_tf32_kernel<<<grid, kWThreads, smem, (cudaStream_t)stream>>>(mm, mn, g);
  const long long n = (long long)Cout * Cin;
  if (used > 1)
    wgrad_reduce_kernel<<<d3d_ceil_div(n, 64), 256, 0, (cudaStream_t)stream>>>((const float*)ws, used, Cout, Cin,
                                                                              plan.swapped ? 1 : 0, accumulate, dW);
  d3d_note_launches(used > 1 ? 2 : 1);
  return d3d_launch_status();
}

/* mean / invstd (and the running statistics, momentum update like torch.nn.BatchNorm1d) of the R x C matrix whose tile
 * partials `stats` d3d_gemm_tf32 wrote. */
int d3d_bn_finalize(const float* stats, long long R, int C, float eps, float momentum, float* running_mean, float* running_var,
                    long long* num_batches_tracked, float* save_mean, float* save_invstd, void* stream) {
  D3D_REQUIRE(stats && save_mean && save_invstd && R > 0 && C > 0);
  bn_finalize_kernel<<<d3d_ceil_div(C, kFinCh), kFinCh * kFinSlices, 0, (cudaStream_t)stream>>>(stats, d3d_gemm_row_tiles(R), R, C, eps, momentum,
                                                                           running_mean, running_var, num_batches_tracked,
                                                                           save_mean, save_invstd);
  d3d_note_launches(1);
  return d3d_launch_status();
}

}  // extern "C"

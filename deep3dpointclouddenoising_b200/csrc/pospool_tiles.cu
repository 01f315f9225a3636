// PosPool ('xyz' embedding, sum / avg) as staged-tile kernels: bulk-asynchronous row staging + tcgen05 contraction.
//
//   ref: u_net_arch/models/local_aggregation_operators.py:140-147,165-183   (PosPool forward; autograd backward)
//   ref: u_net_arch/pt_custom_ops/_ext_src/src/group_points_gpu.cu:13-33,48-69 (the gather / scatter-add it replaces)
//
// The per-query gather kernel (aggregate.cu) reads every (query, slot) row into registers: 1.96 GB through the LSU
// write-back path at the first level, although 128 spatially adjacent queries only touch ~460 distinct rows.  Here a
// CTA owns a TILE of 128 queries that are adjacent in space (spatial_order.cu).  Per (neighbour list, order) pair the
// tile plan kernel computes once: the UNION of the support rows the tile gathers (shared-memory bitmap + popcount
// ranks) and the union rank of every list entry.  With A the 128 x U multiplicity matrix of the tile (0/1/2.. — exact
// in bf16; every row walks its list with a cursor: the ball query emits each query's winners in ascending support
// index, union ranks are monotone in the index, so a chunk consumes a contiguous run of the list) and
// w[u, c] = (S[u] - centre)[c mod 3] the union row's own coordinate relative to the tile centre — PosPool's weight
// (S[u] - Q[q])[c mod 3] is bilinear in the two positions —
//      forward  (pospool_fwd_pipelined_kernel):  out[q, c] = rho_q * (Y2[q, c] - (Q[q] - centre)[c mod 3] * Y1[q, c]),
//                                                Y1 = A . X,  Y2 = A . (w * X),  X = the staged feature rows
//      backward (pospool_scatter_bwd_kernel):    dF[u, c] += (S[u] - centre)[c mod 3] * D1[u, c] - D2[u, c],
//                                                D1 = A^T . G1,  D2 = A^T . G2,  G1 = rho * g,  G2 = rho * (Q - centre) * g
// with rho_q = 1 / (radius * neighbourhood size) (avg) or 1 / radius (sum).  Rows travel global -> shared with
// cp.async.bulk (no register write-back, completion counted on an mbarrier); the contractions run on tcgen05.mma with
// accumulators in TMEM.
// fp32 accuracy: A is exact; X, w * X, G1, G2 are split into three bf16 terms each (8 + 8 + 8 mantissa bits, an exact
// decomposition), products are exact, accumulation is fp32 in TMEM.  Centring on the tile keeps the cancellation in
// (Y2 - rc * Y1) at the scale of (tile extent + radius) / radius.
#include <cstdlib>

#include "common.cuh"
#include "umma.cuh"

namespace {

using namespace umma;

constexpr int kTQ = 128;          // queries per tile = TMEM lanes
constexpr int kCB = 72;           // channels per CTA (multiple of 24: 8-channel groups and the c mod 3 phase line up)
constexpr int kMaxPoints = 16384; // union indices and ranks are uint16; the plan kernel's bitmap lives in shared memory
constexpr int kMaxNs = 64;

struct TileArgs {
  const float* src;          // rows that are staged: features (B, N, C) forward, grad_out (B, M, C) backward
  float* out;                // (B, M, C) forward, (B, N, C) backward
  const float* query_xyz;    // (B, M, 3)
  const float* support_xyz;  // (B, N, 3)
  const int* by_support;     // (B, M, ns) winners in ascending support index, (distance rank << 16) | index
  const int* nvalid;         // (B, M)
  const int* query_mask;     // (B, M)
  const int* order;          // (B, M) processing order of the queries
  int M, N, C, nsample, reduction;
  float inv_radius;
  const void* plan;          // tile plan of (by_support, order)
  float* partial;            // ordered backward: per-tile partial rows [tile][stride][C] (null: float atomics)
  unsigned long long* timing;  // diagnostics (tools/tile_phases.py): 8 timestamps per CTA, or null
};

unsigned long long* g_timing = nullptr;

__device__ __forceinline__ unsigned long long now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define D3D_STAMP(k)                                                                                                  \
  if (a.timing && tid == 0)                                                                                           \
  a.timing[((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 + (k)] = now_ns()

__host__ __device__ inline unsigned align16(unsigned x) { return (x + 15u) & ~15u; }

__device__ __forceinline__ float rot3(float x, float y, float z, int r) { return r == 0 ? x : (r == 1 ? y : z); }

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// bf16 bit pattern of a small non-negative integer (exact up to 256)
__device__ __forceinline__ unsigned short bf16_of_count(int m) { return (unsigned short)(__float_as_uint((float)m) >> 16); }


// ---- tile plan: what the forward tile and the scatter-form backward tile need from the neighbour lists, computed ONCE per
// (list, processing order) instead of in every kernel's prologue (measured 9 of 44 us per forward CTA):
//   counts[tile]            U = size of the union of the support rows the tile's 128 queries gather
//   ranks [tile][128][ns]   union rank of every list entry (uint16, 0xffff = unused slot), rows in tile order
//   unions[tile][stride]    the union's support indices in ascending order (uint16: N <= 16384)
//   tilemask[n], rank_of[tile][n]   the inverse view for the ordered (atomic-free) backward: which tiles gather support
//                           row n and at which union rank
struct PlanView {
  const int* counts;
  const unsigned short* ranks;
  const unsigned short* unions;
  const unsigned* tilemask;        // [B][N][4]: bit t set = tile t of the cloud gathers support row n (tiles <= 128)
  const unsigned short* rank_of;   // [B][tiles][N]: union rank of support n inside tile t (valid where the bit is set)
  int stride;
};

__host__ __device__ inline size_t align256z(size_t x) { return (x + 255) & ~(size_t)255; }
__host__ __device__ inline size_t plan_counts_bytes(int B, int tiles) { return align256z((size_t)B * tiles * 4); }
__host__ __device__ inline int plan_stride(int N, int ns) { return (min(N, kTQ * ns) + 7) & ~7; }
__host__ __device__ inline size_t plan_ranks_bytes(int B, int tiles, int ns) { return align256z((size_t)B * tiles * kTQ * ns * 2); }
__host__ __device__ inline size_t plan_unions_bytes(int B, int tiles, int N, int ns) {
  return align256z((size_t)B * tiles * plan_stride(N, ns) * 2);
}
__host__ __device__ inline size_t plan_tilemask_bytes(int B, int N) { return align256z((size_t)B * N * 16); }
__host__ __device__ inline size_t plan_total_bytes(int B, int M, int N, int ns) {
  const int tiles = (M + kTQ - 1) / kTQ;
  return plan_counts_bytes(B, tiles) + plan_ranks_bytes(B, tiles, ns) + plan_unions_bytes(B, tiles, N, ns) +
         plan_tilemask_bytes(B, N) + align256z((size_t)B * tiles * N * 2);
}
__host__ __device__ inline PlanView plan_view(const void* plan, int B, int M, int N, int ns) {
  const int tiles = (M + kTQ - 1) / kTQ;
  const unsigned char* p = static_cast<const unsigned char*>(plan);
  PlanView v;
  v.counts = reinterpret_cast<const int*>(p);
  p += plan_counts_bytes(B, tiles);
  v.ranks = reinterpret_cast<const unsigned short*>(p);
  p += plan_ranks_bytes(B, tiles, ns);
  v.unions = reinterpret_cast<const unsigned short*>(p);
  p += plan_unions_bytes(B, tiles, N, ns);
  v.tilemask = reinterpret_cast<const unsigned*>(p);
  p += plan_tilemask_bytes(B, N);
  v.rank_of = reinterpret_cast<const unsigned short*>(p);
  v.stride = plan_stride(N, ns);
  return v;
}

constexpr int kPlanThreads = 256;

__global__ void __launch_bounds__(kPlanThreads)
tile_plan_kernel(const int* __restrict__ by_support, const int* __restrict__ nvalid, const int* __restrict__ query_mask,
                 const int* __restrict__ order_all, int B, int M, int N, int ns, void* plan) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x, b = blockIdx.y, tiles = gridDim.x;
  const int W = (N + 31) / 32;
  unsigned short* sEnt = reinterpret_cast<unsigned short*>(smem);
  unsigned* sBitmap = reinterpret_cast<unsigned*>(smem + align16((unsigned)(kTQ * ns * 2)));
  unsigned* sPrefix = sBitmap + ((W + 3) & ~3);
  int* sOwnerId = reinterpret_cast<int*>(sPrefix + ((W + 4) & ~3));
  int* sOwnerInfo = sOwnerId + kTQ;
  __shared__ unsigned sScan[kPlanThreads / 32 + 1];
  const int row0 = tile * kTQ, n_rows = min(kTQ, M - row0);
  const size_t qbase = (size_t)b * M;
  const int* order = order_all + qbase;
  if (tid < kTQ) {
    int own = 0, info = 0;
    if (tid < n_rows) {
      own = order[row0 + tid];
      const int nv = min(nvalid[qbase + own], ns);
      const bool padded = query_mask[qbase + own] == 0;
      info = ((padded && nv == 0) ? 1 : nv) | (nv << 8) | ((padded ? 1 : 0) << 16);
    }
    sOwnerId[tid] = own; sOwnerInfo[tid] = info;
  }
  for (int w = tid; w < W; w += kPlanThreads) sBitmap[w] = 0u;
  __syncthreads();
  for (int r = warp; r < kTQ; r += kPlanThreads / 32) {
    const int info = r < n_rows ? sOwnerInfo[r] : 0;
    const int n_ent = info & 255;
    const bool row0_only = ((info >> 16) & 1) && ((info >> 8) & 255) == 0;  // padded query without neighbours: row 0
    const int* lrow = by_support + (qbase + sOwnerId[r]) * ns;
    for (int k = lane; k < ns; k += 32) {
      unsigned id = 0xffffu;
      if (k < n_ent) {
        id = row0_only ? 0u : (unsigned)(lrow[k] & 0xffff);
        if (id >= (unsigned)N) id = 0u;
        atomicOr(&sBitmap[id >> 5], 1u << (id & 31));
      }
      sEnt[r * ns + k] = (unsigned short)id;
    }
  }
  __syncthreads();
  // exclusive prefix of the per-word popcounts (W <= 512: two words per thread)
  const int w0 = 2 * tid;
  const unsigned p0 = w0 < W ? __popc(sBitmap[w0]) : 0u, p1 = w0 + 1 < W ? __popc(sBitmap[w0 + 1]) : 0u;
  const unsigned sum = p0 + p1;
  unsigned incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned up = __shfl_up_sync(D3D_FULL_MASK, incl, o);
    if (lane >= o) incl += up;
  }
  if (lane == 31) sScan[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const unsigned v = lane < kPlanThreads / 32 ? sScan[lane] : 0u;
    unsigned vi = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned up = __shfl_up_sync(D3D_FULL_MASK, vi, o);
      if (lane >= o) vi += up;
    }
    if (lane < kPlanThreads / 32) sScan[lane] = vi - v;
    if (lane == 31) sScan[kPlanThreads / 32] = vi;
  }
  __syncthreads();
  const unsigned base = sScan[warp] + incl - sum;
  if (w0 < W) sPrefix[w0] = base;
  if (w0 + 1 < W) sPrefix[w0 + 1] = base + p0;
  __syncthreads();
  const PlanView pv = plan_view(plan, B, M, N, ns);
  const size_t t_lin = (size_t)b * tiles + tile;
  if (tid == 0) const_cast<int*>(pv.counts)[t_lin] = (int)sScan[kPlanThreads / 32];
  unsigned short* ranks = const_cast<unsigned short*>(pv.ranks) + t_lin * kTQ * ns;
  for (int e = tid; e < kTQ * ns; e += kPlanThreads) {
    const unsigned v = sEnt[e];
    ranks[e] = v == 0xffffu ? (unsigned short)0xffffu
                            : (unsigned short)(sPrefix[v >> 5] + __popc(sBitmap[v >> 5] & ((1u << (v & 31)) - 1u)));
  }
  unsigned short* uni = const_cast<unsigned short*>(pv.unions) + t_lin * pv.stride;
  unsigned* tmask = const_cast<unsigned*>(pv.tilemask) + (size_t)b * N * 4 + (tile >> 5);
  unsigned short* rank_of = const_cast<unsigned short*>(pv.rank_of) + t_lin * N;
  for (int w = tid; w < W; w += kPlanThreads) {
    unsigned bits = sBitmap[w];
    unsigned r = sPrefix[w];
    while (bits) {
      const int bit = __ffs(bits) - 1;
      bits &= bits - 1;
      const int n = w * 32 + bit;
      atomicOr(tmask + (size_t)n * 4, 1u << (tile & 31));  // integer OR: the result does not depend on the order
      rank_of[n] = (unsigned short)r;
      uni[r++] = (unsigned short)n;
    }
  }
}

// ================================================================================================================
// Backward in SCATTER form: the CTA owns 128 QUERIES (the forward tile) and distributes their gradient rows over the
// union of the supports they gathered.  With A the forward multiplicity matrix (128 queries x U union rows),
//      dF[u, c] += (S[u] - centre)[c mod 3] * D1[u, c] - D2[u, c],     D1 = A^T . G1,   D2 = A^T . G2,
//      G1[q, c] = rho_q * grad_out[q, c],   G2[q, c] = rho_q * (Q[q] - centre)[c mod 3] * grad_out[q, c]
// — the transpose of the forward contraction.  The 128 gradient rows are staged ONCE per tile (bulk copies) and split
// into bf16 planes once (the gather form above stages the 3x larger union of the gathering queries chunk by chunk);
// A^T is the same shared-memory image as A read with the MN-major descriptor.  The union is processed in blocks of
// 128 rows = one accumulator (TMEM lanes); three roles run concurrently on double-buffered A blocks and accumulators:
//   warps 0-3   build block m+1 of A (thread = query, cursor along its list as in the forward kernel),
//   warp 12     issues the 48 MMAs of block m,
//   warps 4-11  drain block m-1: TMEM -> registers -> red.global.add.v4.f32 into dF (zero-filled by the entry point).
// The sums of one dF row over the (on average 3.6) tiles that reference it arrive as float atomics, like the
// reference's own backward (group_points_gpu.cu:48-69): results are reproducible to rounding, not bit for bit.
constexpr int kBT = 416;
constexpr int kUB = 128;                       // union rows per block = MMA M
constexpr unsigned kAtGroup = (kUB / 8) * 128; // one 8-query group of a block: 16 chunks of 8 union ranks, 2 KB
constexpr unsigned kAtBytes = (kTQ / 8) * kAtGroup;  // 32 KB per block

struct ScatterLayout {
  unsigned a, planes, plane_bytes, lbo_b, row_bytes, ent, owner_id, owner_info, owner_xyz, owner_rho, scan, bars, total;
  int np;
};

__host__ __device__ inline ScatterLayout make_scatter_layout(int cbn, int ns) {
  ScatterLayout L;
  unsigned o = 0;
  L.a = o; o += 2 * kAtBytes;  // the gradient rows are staged here before the first A block is built
  L.lbo_b = (unsigned)((cbn + 7) / 8) * 128u;
  L.plane_bytes = (kTQ / 8) * L.lbo_b;
  L.planes = o; o += 6 * L.plane_bytes + 128;
  L.row_bytes = (unsigned)cbn * 4u;
  L.ent = o; o += align16((unsigned)(kTQ * ns * 2));
  L.owner_id = o; o += kTQ * 4;
  L.owner_info = o; o += kTQ * 4;
  L.owner_xyz = o; o += kTQ * 12;
  L.owner_rho = o; o += kTQ * 4;
  L.scan = o; o += 32 * 4;
  L.bars = o; o += 80;
  L.total = o;
  L.np = (cbn + 15) & ~15;
  return L;
}

__global__ void __launch_bounds__(kBT, 1)
pospool_scatter_bwd_kernel(const TileArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x, b = blockIdx.z;
  const int c0 = blockIdx.y * kCB, cbn = min(kCB, a.C - c0);
  const int ns = a.nsample;
  const ScatterLayout L = make_scatter_layout(cbn, ns);

  unsigned char* sA = smem + L.a;
  unsigned char* sPlanes = smem + L.planes;
  unsigned short* sEnt = reinterpret_cast<unsigned short*>(smem + L.ent);
  int* sOwnerId = reinterpret_cast<int*>(smem + L.owner_id);
  int* sOwnerInfo = reinterpret_cast<int*>(smem + L.owner_info);
  float* sOwnerXyz = reinterpret_cast<float*>(smem + L.owner_xyz);
  float* sOwnerRho = reinterpret_cast<float*>(smem + L.owner_rho);
  float* sCtr = reinterpret_cast<float*>(smem + L.scan) + 20;
  const unsigned bar0 = smem_u32(smem + L.bars);
  const unsigned bar_stage = bar0, tmem_slot = bar0 + 56, bar_plan = bar0 + 64;
  auto bar_a_full = [&](int buf) { return bar0 + 8 + 8 * buf; };      // 128 builder arrivals
  auto bar_mma = [&](int buf) { return bar0 + 24 + 8 * buf; };         // tcgen05.commit: A block free, accumulator full
  auto bar_acc_free = [&](int buf) { return bar0 + 40 + 8 * buf; };    // 256 drain arrivals

  const float* own_xyz = a.query_xyz + (size_t)b * a.M * 3;
  const float* src_xyz = a.support_xyz + (size_t)b * a.N * 3;
  const int* order = a.order + (size_t)b * a.M;
  const int row0 = tile * kTQ;
  const int n_rows = min(kTQ, a.M - row0);
  const size_t qbase = (size_t)b * a.M;

  D3D_STAMP(0);
  // ---- owners (queries), barriers, TMEM ---------------------------------------------------------------------------
  if (warp == 12) tmem_alloc(tmem_slot, 512u);
  const float* g_rows = a.src + (size_t)b * a.M * a.C + c0;
  if (tid >= kBT - 128) {
    // the tile's gradient rows start flying at once: these threads read the processing order themselves
    const int t = tid - (kBT - 128);
    if (t == 0) {
      mbar_init(bar_stage, 1);
      mbar_init_fence();
      mbar_arrive_expect_tx(bar_stage, (unsigned)n_rows * L.row_bytes);
    }
    named_bar_sync(1, 128);  // barrier initialised, expect_tx precedes every complete_tx
    if (t < n_rows) bulk_g2s(smem_u32(sA + (size_t)t * L.row_bytes), g_rows + (size_t)order[row0 + t] * a.C, L.row_bytes, bar_stage);
  }
  if (tid == 160) {
    mbar_init(bar_plan, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_a_full(i), kTQ);
      mbar_init(bar_mma(i), 1);
      mbar_init(bar_acc_free(i), 256);
    }
    mbar_init_fence();
  }
  if (tid < kTQ) {
    int own = -1, info = 0;
    float px = 0.f, py = 0.f, pz = 0.f, rho = 0.f;
    if (tid < n_rows) {
      own = order[row0 + tid];
      px = own_xyz[3 * (size_t)own]; py = own_xyz[3 * (size_t)own + 1]; pz = own_xyz[3 * (size_t)own + 2];
      const int nv = min(a.nvalid[qbase + own], ns);
      const bool padded = a.query_mask[qbase + own] == 0;
      const int neff = padded ? ns : nv;
      info = ((padded && nv == 0) ? 1 : nv) | (nv << 8) | ((padded ? 1 : 0) << 16);
      rho = a.reduction == D3D_REDUCE_AVG ? a.inv_radius / (float)neff : a.inv_radius;
    }
    sOwnerId[tid] = own; sOwnerInfo[tid] = info;
    sOwnerXyz[3 * tid] = px; sOwnerXyz[3 * tid + 1] = py; sOwnerXyz[3 * tid + 2] = pz;
    sOwnerRho[tid] = rho;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      px += __shfl_xor_sync(D3D_FULL_MASK, px, o);
      py += __shfl_xor_sync(D3D_FULL_MASK, py, o);
      pz += __shfl_xor_sync(D3D_FULL_MASK, pz, o);
    }
    if (lane == 0) { sCtr[3 * warp] = px; sCtr[3 * warp + 1] = py; sCtr[3 * warp + 2] = pz; }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_base = *reinterpret_cast<volatile unsigned*>(smem + L.bars + 56);
  const float inv_rows = 1.0f / (float)n_rows;
  const float ctr_x = (sCtr[0] + sCtr[3] + sCtr[6] + sCtr[9]) * inv_rows, ctr_y = (sCtr[1] + sCtr[4] + sCtr[7] + sCtr[10]) * inv_rows,
              ctr_z = (sCtr[2] + sCtr[5] + sCtr[8] + sCtr[11]) * inv_rows;

  D3D_STAMP(1);

  // ---- union ranks of the list entries and the union itself: from the tile plan (one bulk copy into sEnt) ----------
  const PlanView pv = plan_view(a.plan, (int)gridDim.z, a.M, a.N, ns);
  const size_t t_lin = (size_t)b * gridDim.x + tile;
  const unsigned short* tile_union = pv.unions + t_lin * pv.stride;
  const int U = pv.counts[t_lin];
  if (tid == 0) {
    const unsigned bytes = (unsigned)(kTQ * ns * 2);
    mbar_arrive_expect_tx(bar_plan, bytes);
    bulk_g2s(smem_u32(sEnt), pv.ranks + t_lin * kTQ * ns, bytes, bar_plan);
  }

  // ---- gradient rows -> 3 bf16 planes of G1 and 3 of G2 (MN-major B operand, K = query) ----------------------------
  const int n_groups = (cbn + 7) >> 3;
  mbar_wait_short(bar_stage, 0u);
  D3D_STAMP(2);
  for (int task = tid; task < kTQ * n_groups; task += kBT) {
    const int u = task & (kTQ - 1), g = task >> 7;
    unsigned char* dst = sPlanes + (u >> 3) * L.lbo_b + g * 128 + (u & 7) * 16;
    unsigned hx[3][4], hy[3][4];
    if (u >= n_rows) {
#pragma unroll
      for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int i = 0; i < 4; ++i) hx[p][i] = hy[p][i] = 0u;
    } else {
      const float* row = reinterpret_cast<const float*>(sA + (size_t)u * L.row_bytes) + 8 * g;
      const float4 xa = *reinterpret_cast<const float4*>(row);
      const float4 xb = (8 * g + 4 < cbn) ? *reinterpret_cast<const float4*>(row + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float sc = sOwnerRho[u];
      const float x[8] = {xa.x * sc, xa.y * sc, xa.z * sc, xa.w * sc, xb.x * sc, xb.y * sc, xb.z * sc, xb.w * sc};
      const float wx = sOwnerXyz[3 * u] - ctr_x, wy = sOwnerXyz[3 * u + 1] - ctr_y, wz = sOwnerXyz[3 * u + 2] - ctr_z;
      const int base = (c0 + 8 * g) % 3;
      const float w0 = rot3(wx, wy, wz, base), w1 = rot3(wy, wz, wx, base), w2 = rot3(wz, wx, wy, base);
      const float y[8] = {x[0] * w0, x[1] * w1, x[2] * w2, x[3] * w0, x[4] * w1, x[5] * w2, x[6] * w0, x[7] * w1};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        split3_bf16x2(x[2 * i], x[2 * i + 1], hx[0][i], hx[1][i], hx[2][i]);
        split3_bf16x2(y[2 * i], y[2 * i + 1], hy[0][i], hy[1][i], hy[2][i]);
      }
    }
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      *reinterpret_cast<uint4*>(dst + (size_t)p * L.plane_bytes) = make_uint4(hx[p][0], hx[p][1], hx[p][2], hx[p][3]);
      *reinterpret_cast<uint4*>(dst + (size_t)(3 + p) * L.plane_bytes) = make_uint4(hy[p][0], hy[p][1], hy[p][2], hy[p][3]);
    }
  }
  fence_async_smem();
  mbar_wait_short(bar_plan, 0u);
  __syncthreads();  // planes complete, ranks in place, the staging area (= the A blocks) is free
  D3D_STAMP(3);

  const int n_blocks = (U + kUB - 1) / kUB;
  const int ksteps = (n_rows + 15) >> 4;
  if (warp < 4) {
    // ---- A blocks: thread = query; element (q, u) of block m at (q / 8) * 2 KB + (u / 8) * 128 + (q % 8) * 16 + (u % 8) * 2
    const int info = tid < n_rows ? sOwnerInfo[tid] : 0;
    const int cur_end = info & 255;
    int cur = 0;
    const int* prow = a.by_support + (qbase + (tid < n_rows ? sOwnerId[tid] : 0)) * ns;
    const unsigned short* e = sEnt + tid * ns;
    for (int m = 0; m < n_blocks; ++m) {
      const int buf = m & 1;
      if (m >= 2) {
        mbar_wait_short(bar_mma(buf), (unsigned)(((m >> 1) - 1) & 1));
        tc_fence_after();
      }
      unsigned char* arow = sA + buf * kAtBytes + (tid >> 3) * kAtGroup + (tid & 7) * 16;
#pragma unroll
      for (int g = 0; g < kUB / 8; ++g) *reinterpret_cast<uint4*>(arow + g * 128) = make_uint4(0u, 0u, 0u, 0u);
      const int limit = (m + 1) * kUB;
      while (cur < cur_end) {  // four list entries per round: their loads are independent of each other
        int rr[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) rr[i] = cur + i < cur_end ? (int)e[cur + i] : 0x7fffffff;
        bool full = false;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = rr[i];
          if (r >= limit) { full = true; break; }
          int mult = 1;
          if (info >> 16) {  // padded query: slot k of the reference list repeats winner k % nvalid
            const int nv = (info >> 8) & 255;
            mult = nv > 0 ? (ns - 1 - ((prow[cur] >> 16) & 255)) / nv + 1 : ns;
          }
          *reinterpret_cast<unsigned short*>(arow + ((r & (kUB - 1)) >> 3) * 128 + (r & 7) * 2) = bf16_of_count(mult);
          ++cur;
        }
        if (full) break;
      }
      fence_async_smem();
      mbar_arrive(bar_a_full(buf));
    }
  } else if (warp == 12) {
    // ---- MMA issue: D1 += A^T . G1 planes, D2 += A^T . G2 planes; accumulator buffer = 2 * np columns -------------
    if (lane == 0) {
      const unsigned idesc = idesc_bf16(L.np, true, true);  // A^T MN-major (M = union rank), B MN-major
      const unsigned p_addr = smem_u32(sPlanes);
      for (int m = 0; m < n_blocks; ++m) {
        const int buf = m & 1;
        mbar_wait_short(bar_a_full(buf), (unsigned)((m >> 1) & 1));
        if (m >= 2) mbar_wait_short(bar_acc_free(buf), (unsigned)(((m >> 1) - 1) & 1));
        tc_fence_after();
        const unsigned a_addr = smem_u32(sA + buf * kAtBytes);
        const unsigned acc0 = tmem_base + (unsigned)(buf * 256);
        for (int ks = 0; ks < ksteps; ++ks) {
          const unsigned long long a_desc = smem_desc(a_addr + ks * 2 * kAtGroup, kAtGroup, 128);
#pragma unroll
          for (int p = 0; p < 6; ++p) {
            const unsigned long long b_desc = smem_desc(p_addr + p * L.plane_bytes + ks * 2 * L.lbo_b, L.lbo_b, 128);
            mma_bf16(acc0 + (p < 3 ? 0u : (unsigned)L.np), a_desc, b_desc, idesc, (ks == 0 && (p == 0 || p == 3)) ? 0u : 1u);
          }
        }
        mma_commit(bar_mma(buf));
      }
    }
  } else {
    // ---- drain: thread = union row of the block (TMEM lane); the two warp groups alternate over the 16-column pieces ----
    const int lq = warp & 3, t = lq * 32 + lane, grp = (warp - 4) >> 2;
    float* out_rows = a.out + (size_t)b * a.N * a.C + c0;
    for (int m = 0; m < n_blocks; ++m) {
      const int buf = m & 1;
      const int r = m * kUB + t;
      int src = -1;
      float sx = 0.f, sy = 0.f, sz = 0.f;
      if (r < U) {
        src = tile_union[r];
        sx = src_xyz[3 * (size_t)src] - ctr_x; sy = src_xyz[3 * (size_t)src + 1] - ctr_y; sz = src_xyz[3 * (size_t)src + 2] - ctr_z;
      }
      mbar_wait_short(bar_mma(buf), (unsigned)((m >> 1) & 1));
      tc_fence_after();
      float* orow = out_rows + (size_t)(src >= 0 ? src : 0) * a.C;
      for (int ch = grp; ch * 16 < cbn; ch += 2) {
        unsigned d1[16], d2[16];
        const unsigned taddr = tmem_base + ((unsigned)(lq * 32) << 16) + (unsigned)(buf * 256 + ch * 16);
        tmem_ld16(taddr, d1);
        tmem_ld16(taddr + (unsigned)L.np, d2);
        tmem_ld_wait();
        const int base = (c0 + 16 * ch) % 3;
        const float r0 = rot3(sx, sy, sz, base), r1 = rot3(sy, sz, sx, base), r2 = rot3(sz, sx, sy, base);
        float o[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float rc = (i % 3 == 0) ? r0 : ((i % 3 == 1) ? r1 : r2);
          o[i] = rc * __uint_as_float(d1[i]) - __uint_as_float(d2[i]);
        }
        if (src >= 0) {
          if (a.partial) {  // ordered form: the tile's own row of partial sums; pospool_scatter_reduce_kernel adds them up
            float* prow_out = a.partial + ((t_lin * pv.stride + (size_t)r) * a.C + c0);
#pragma unroll
            for (int v = 0; v < 4; ++v)
              if (16 * ch + 4 * v < cbn)
                *reinterpret_cast<float4*>(prow_out + 16 * ch + 4 * v) = make_float4(o[4 * v], o[4 * v + 1], o[4 * v + 2], o[4 * v + 3]);
          } else {
#pragma unroll
            for (int v = 0; v < 4; ++v)
              if (16 * ch + 4 * v < cbn) red_add_v4(orow + 16 * ch + 4 * v, o[4 * v], o[4 * v + 1], o[4 * v + 2], o[4 * v + 3]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_acc_free(buf));
    }
  }
  tc_fence_before();
  __syncthreads();
  D3D_STAMP(4);
  D3D_STAMP(5);
  if (a.timing && tid == 0) a.timing[((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 + 6] = (unsigned long long)U;
  if (warp == 12) tmem_dealloc(tmem_base, 512u);
}

// Ordered form, second kernel: a warp owns one support row and adds the partial rows of the tiles that gathered it in
// ascending tile order (tilemask / rank_of of the plan) — a fixed-order segmented reduction, no float atomics; rows no
// tile gathered come out as zeros (no separate zero fill).
__global__ void __launch_bounds__(256)
pospool_scatter_reduce_kernel(const float* __restrict__ partial, const void* plan, int B, int M, int N, int C, int ns,
                              float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= (long long)B * N) return;
  const int b = (int)(row / N), n = (int)(row - (long long)b * N);
  const PlanView pv = plan_view(plan, B, M, N, ns);
  const int tiles = (M + kTQ - 1) / kTQ;
  const uint4 mask = __ldg(reinterpret_cast<const uint4*>(pv.tilemask) + row);
  const unsigned words[4] = {mask.x, mask.y, mask.z, mask.w};
  // lane l looks up the rank of tile (32 w + l) for every word w with its bit set: at most four independent loads
  unsigned rk[4];
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    rk[w] = 0u;
    const int t = 32 * w + lane;
    if ((words[w] >> lane) & 1u) rk[w] = __ldg(pv.rank_of + ((size_t)b * tiles + t) * N + n);
  }
  const int groups = C >> 2;
  float4* orow = reinterpret_cast<float4*>(out + (size_t)row * C);
  for (int g0 = 0; g0 < groups; g0 += 32) {
    const int g = g0 + lane;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      unsigned bits = words[w];
      while (bits) {
        const int l = __ffs(bits) - 1;
        bits &= bits - 1;
        const unsigned r = __shfl_sync(D3D_FULL_MASK, rk[w], l);
        const size_t t_lin = (size_t)b * tiles + 32 * w + l;
        if (g < groups) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(partial + (t_lin * pv.stride + r) * C) + g);
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
      }
    }
    if (g < groups) orow[g] = acc;
  }
}

int launch_scatter_bwd(const TileArgs& a, int B, cudaStream_t st) {
  if (a.M > kMaxPoints || a.N > kMaxPoints || a.nsample > kMaxNs || a.C % 4 != 0) return D3D_ERR_UNSUPPORTED;
  const int cbn_max = a.C < kCB ? a.C : kCB;
  const ScatterLayout L = make_scatter_layout(cbn_max, a.nsample);
  if ((unsigned)kTQ * L.row_bytes > 2 * kAtBytes) return D3D_ERR_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(pospool_scatter_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
  if (e != cudaSuccess) return (int)e;
  if (!a.partial) {  // atomic form: the kernel adds into the output
    e = cudaMemsetAsync(a.out, 0, (size_t)B * a.N * a.C * sizeof(float), st);
    if (e != cudaSuccess) return (int)e;
  }
  dim3 grid(d3d_ceil_div(a.M, kTQ), d3d_ceil_div(a.C, kCB), B);
  TileArgs t = a;
  t.timing = g_timing;
  pospool_scatter_bwd_kernel<<<grid, kBT, L.total, st>>>(t);
  d3d_note_launches(1);
  if (a.partial) {
    const long long rows = (long long)B * a.N;
    pospool_scatter_reduce_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(a.partial, a.plan, B, a.M, a.N, a.C, a.nsample, a.out);
    d3d_note_launches(1);
  }
  return d3d_launch_status();
}

// ================================================================================================================
// Forward tile, pipelined: the same contraction as pospool_tiles_kernel<false>, with the chunk loop turned into a
// producer / consumer pipeline over double-buffered stages (32 union rows per chunk) — no CTA-wide barrier inside the loop:
//   warp 4      loader: picks the chunk's rows from the tile plan and issues their bulk copies two chunks ahead,
//   warps 0-3   build the chunk of A (thread = query row), then join
//   warps 6-11  in converting the staged rows into the six bf16 operand planes (one (row, 8 channels) task per thread),
//   warp 5      issues the chunk's 12 MMAs as soon as A and the planes of the chunk are complete.
// Measured on the gather form above: 3.9 us per 64-row chunk, almost all of it exposed latency (copy -> convert ->
// CTA barrier -> MMA -> next copy in sequence); here those overlap across chunks.
constexpr int kFT = 384;
constexpr int kFC = 32;                          // union rows per chunk (2 K-steps)
constexpr int kFConv = 320;                      // converting threads: warps 0-3 and 6-11
constexpr unsigned kFASbo = (kFC / 8) * 128;     // A chunk: 8-row groups 512 B apart
constexpr unsigned kFABytes = (kTQ / 8) * kFASbo;  // 8 KB

struct FwdLayout {
  unsigned a, planes, plane_bytes, planes_stride, lbo_b, stage, stage_stride, row_bytes, ent, owner_id, owner_info, owner_xyz,
      owner_rho, src_id, src_w, scan, bars, total;
  int np, tmem_cols;
};

__host__ __device__ inline FwdLayout make_fwd_layout(int cbn, int ns) {
  FwdLayout L;
  unsigned o = 0;
  L.a = o; o += 2 * kFABytes;
  L.lbo_b = (unsigned)((cbn + 7) / 8) * 128u;
  L.plane_bytes = (kFC / 8) * L.lbo_b;
  L.planes_stride = 6 * L.plane_bytes + 128;  // +128: the MMA reads N rounded up to 16 channels
  L.planes = o; o += 2 * L.planes_stride;
  L.row_bytes = (unsigned)cbn * 4u;
  L.stage_stride = (kFC * L.row_bytes + 127u) & ~127u;
  L.stage = o; o += 2 * L.stage_stride;
  L.ent = o; o += align16((unsigned)(kTQ * ns * 2));
  L.owner_id = o; o += kTQ * 4;
  L.owner_info = o; o += kTQ * 4;
  L.owner_xyz = o; o += kTQ * 12;
  L.owner_rho = o; o += kTQ * 4;
  L.src_id = o; o += 2 * kFC * 4;
  L.src_w = o; o += 2 * kFC * 12;
  L.scan = o; o += 32 * 4;
  L.bars = o; o += 96;
  L.total = o;
  L.np = (cbn + 15) & ~15;
  L.tmem_cols = 32;
  while (L.tmem_cols < 2 * L.np) L.tmem_cols <<= 1;
  return L;
}

__global__ void __launch_bounds__(kFT, 2)
pospool_fwd_pipelined_kernel(const TileArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x, b = blockIdx.z;
  const int c0 = blockIdx.y * kCB, cbn = min(kCB, a.C - c0);
  const int ns = a.nsample;
  const FwdLayout L = make_fwd_layout(cbn, ns);

  unsigned short* sEnt = reinterpret_cast<unsigned short*>(smem + L.ent);
  int* sOwnerId = reinterpret_cast<int*>(smem + L.owner_id);
  int* sOwnerInfo = reinterpret_cast<int*>(smem + L.owner_info);
  float* sOwnerXyz = reinterpret_cast<float*>(smem + L.owner_xyz);
  float* sOwnerRho = reinterpret_cast<float*>(smem + L.owner_rho);
  int* sSrcId = reinterpret_cast<int*>(smem + L.src_id);    // [2][kFC]
  float* sSrcW = reinterpret_cast<float*>(smem + L.src_w);  // [2][kFC][3]
  float* sCtr = reinterpret_cast<float*>(smem + L.scan) + 20;
  const unsigned bar0 = smem_u32(smem + L.bars);
  auto bar_stage_full = [&](int buf) { return bar0 + 8 * buf; };        // loader: expect_tx + the bulk copies
  auto bar_stage_free = [&](int buf) { return bar0 + 16 + 8 * buf; };   // kFConv arrivals: the staged rows were read
  auto bar_ops_full = [&](int buf) { return bar0 + 32 + 8 * buf; };     // kFConv arrivals: A chunk and planes written
  auto bar_mma = [&](int buf) { return bar0 + 48 + 8 * buf; };          // tcgen05.commit: A chunk and planes free again
  const unsigned bar_plan = bar0 + 64, tmem_slot = bar0 + 72;

  const float* own_xyz = a.query_xyz + (size_t)b * a.M * 3;
  const float* src_xyz = a.support_xyz + (size_t)b * a.N * 3;
  const int* order = a.order + (size_t)b * a.M;
  const int row0 = tile * kTQ;
  const int n_rows = min(kTQ, a.M - row0);
  const size_t qbase = (size_t)b * a.M;
  const PlanView pv = plan_view(a.plan, (int)gridDim.z, a.M, a.N, ns);
  const size_t t_lin = (size_t)b * gridDim.x + tile;
  const unsigned short* tile_union = pv.unions + t_lin * pv.stride;

  D3D_STAMP(0);
  if (warp == 5) tmem_alloc(tmem_slot, (unsigned)L.tmem_cols);
  if (tid == 128) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_stage_full(i), 1);
      mbar_init(bar_stage_free(i), kFConv);
      mbar_init(bar_ops_full(i), kFConv);
      mbar_init(bar_mma(i), 1);
    }
    mbar_init(bar_plan, 1);
    mbar_init_fence();
    const unsigned bytes = (unsigned)(kTQ * ns * 2);  // union ranks of the list entries: one bulk copy into sEnt
    mbar_arrive_expect_tx(bar_plan, bytes);
    bulk_g2s(smem_u32(sEnt), pv.ranks + t_lin * kTQ * ns, bytes, bar_plan);
  }
  if (tid < kTQ) {
    int own = -1, info = 0;
    float px = 0.f, py = 0.f, pz = 0.f, rho = 0.f;
    if (tid < n_rows) {
      own = order[row0 + tid];
      px = own_xyz[3 * (size_t)own]; py = own_xyz[3 * (size_t)own + 1]; pz = own_xyz[3 * (size_t)own + 2];
      const int nv = min(a.nvalid[qbase + own], ns);
      const bool padded = a.query_mask[qbase + own] == 0;
      const int neff = padded ? ns : nv;
      info = ((padded && nv == 0) ? 1 : nv) | (nv << 8) | ((padded ? 1 : 0) << 16);
      rho = a.reduction == D3D_REDUCE_AVG ? a.inv_radius / (float)neff : a.inv_radius;  // :175-176
    }
    sOwnerId[tid] = own; sOwnerInfo[tid] = info;
    sOwnerXyz[3 * tid] = px; sOwnerXyz[3 * tid + 1] = py; sOwnerXyz[3 * tid + 2] = pz;
    sOwnerRho[tid] = rho;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      px += __shfl_xor_sync(D3D_FULL_MASK, px, o);
      py += __shfl_xor_sync(D3D_FULL_MASK, py, o);
      pz += __shfl_xor_sync(D3D_FULL_MASK, pz, o);
    }
    if (lane == 0) { sCtr[3 * warp] = px; sCtr[3 * warp + 1] = py; sCtr[3 * warp + 2] = pz; }
  }
  const int U = pv.counts[t_lin];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  D3D_STAMP(1);
  const unsigned tmem_base = *reinterpret_cast<volatile unsigned*>(smem + L.bars + 72);
  const float inv_rows = 1.0f / (float)n_rows;
  const float ctr_x = (sCtr[0] + sCtr[3] + sCtr[6] + sCtr[9]) * inv_rows, ctr_y = (sCtr[1] + sCtr[4] + sCtr[7] + sCtr[10]) * inv_rows,
              ctr_z = (sCtr[2] + sCtr[5] + sCtr[8] + sCtr[11]) * inv_rows;
  const int n_chunks = (U + kFC - 1) / kFC;
  const int n_groups = (cbn + 7) >> 3;
  D3D_STAMP(2);
  D3D_STAMP(3);

  if (warp == 4) {
    // ---- loader: one row per lane.  The row numbers and coordinates of chunk j + 1 are fetched while chunk j is being
    //      issued and its staging buffer waited for: the two dependent global loads (union entry -> coordinates) are the
    //      longest latency of the loop and must not sit between two chunks
    const float* src_rows = a.src + (size_t)b * a.N * a.C + c0;
    int nsrc = -1;
    float nx = 0.f, ny = 0.f, nz = 0.f;
    if (lane < U) {
      nsrc = tile_union[lane];
      nx = src_xyz[3 * (size_t)nsrc]; ny = src_xyz[3 * (size_t)nsrc + 1]; nz = src_xyz[3 * (size_t)nsrc + 2];
    }
    for (int j = 0; j < n_chunks; ++j) {
      const int buf = j & 1;
      const int src = nsrc;
      const float sx = nx, sy = ny, sz = nz;
      nsrc = -1;
      const int rn = (j + 1) * kFC + lane;
      if (rn < U) {  // next chunk's row: in flight during the wait below
        nsrc = tile_union[rn];
        nx = src_xyz[3 * (size_t)nsrc]; ny = src_xyz[3 * (size_t)nsrc + 1]; nz = src_xyz[3 * (size_t)nsrc + 2];
      }
      if (j >= 2) mbar_wait(bar_stage_free(buf), (unsigned)(((j >> 1) - 1) & 1));
      if (src >= 0) {
        bulk_g2s(smem_u32(smem + L.stage + buf * L.stage_stride + (size_t)lane * L.row_bytes), src_rows + (size_t)src * a.C,
                 L.row_bytes, bar_stage_full(buf));
        float* w = sSrcW + (buf * kFC + lane) * 3;
        w[0] = sx - ctr_x; w[1] = sy - ctr_y; w[2] = sz - ctr_z;
      }
      sSrcId[buf * kFC + lane] = src;
      __syncwarp();
      if (lane == 0) mbar_arrive_expect_tx(bar_stage_full(buf), (unsigned)min(kFC, U - j * kFC) * L.row_bytes);
    }
  } else if (warp == 5) {
    // ---- MMA issue: Y1 += A . X planes, Y2 += A . (w X) planes -----------------------------------------------------
    if (lane == 0) {
      const unsigned idesc = idesc_bf16(L.np, false, true);  // A K-major, B MN-major
      for (int j = 0; j < n_chunks; ++j) {
        const int buf = j & 1;
        const int ksteps = (min(kFC, U - j * kFC) + 15) >> 4;
        mbar_wait_short(bar_ops_full(buf), (unsigned)((j >> 1) & 1));
        tc_fence_after();
        const unsigned a_addr = smem_u32(smem + L.a + buf * kFABytes), p_addr = smem_u32(smem + L.planes + buf * L.planes_stride);
        for (int ks = 0; ks < ksteps; ++ks) {
          const unsigned long long a_desc = smem_desc(a_addr + ks * 256, 128, kFASbo);
#pragma unroll
          for (int p = 0; p < 6; ++p) {
            const unsigned long long b_desc = smem_desc(p_addr + p * L.plane_bytes + ks * 2 * L.lbo_b, L.lbo_b, 128);
            const unsigned acc = (j == 0 && ks == 0 && (p == 0 || p == 3)) ? 0u : 1u;
            mma_bf16(tmem_base + (p < 3 ? 0u : (unsigned)L.np), a_desc, b_desc, idesc, acc);
          }
        }
        mma_commit(bar_mma(buf));
      }
    }
  } else {
    // ---- A chunks (warps 0-3) and operand planes (warps 0-3, 6-11) -------------------------------------------------
    const int cid = warp < 4 ? tid : tid - 64;  // 0..319
    int cur = 0, cur_end = 0, row_info = 0;
    const int* prow = nullptr;
    const unsigned short* e = sEnt + tid * ns;
    if (warp < 4 && tid < n_rows) {
      row_info = sOwnerInfo[tid];
      cur_end = row_info & 255;
      prow = a.by_support + (qbase + sOwnerId[tid]) * ns;
    }
    mbar_wait(bar_plan, 0u);  // the ranks have landed in sEnt
    const int cu = cid & (kFC - 1), cg = cid >> 5;  // this thread's conversion task: (row of the chunk, 8-channel group)
    for (int j = 0; j < n_chunks; ++j) {
      const int buf = j & 1;
      const int rows16 = ((min(kFC, U - j * kFC) + 15) >> 4) * 16;
      if (j >= 2) {
        mbar_wait(bar_mma(buf), (unsigned)(((j >> 1) - 1) & 1));
        tc_fence_after();
      }
      if (warp < 4) {
        unsigned char* arow = smem + L.a + buf * kFABytes + (tid >> 3) * kFASbo + (tid & 7) * 16;
#pragma unroll
        for (int g = 0; g < kFC / 8; ++g) *reinterpret_cast<uint4*>(arow + g * 128) = make_uint4(0u, 0u, 0u, 0u);
        const int limit = (j + 1) * kFC;
        while (cur < cur_end) {  // two list entries per round (a 32-row chunk takes ~3.6 entries of a row)
          const int ra = e[cur], rb = cur + 1 < cur_end ? (int)e[cur + 1] : 0x7fffffff;
          if (ra >= limit) break;
          int m = 1;
          if (row_info >> 16) {  // padded query: slot k of the reference list repeats winner k % nvalid (cyclic padding)
            const int nv = (row_info >> 8) & 255;
            m = nv > 0 ? (ns - 1 - ((prow[cur] >> 16) & 255)) / nv + 1 : ns;
          }
          *reinterpret_cast<unsigned short*>(arow + ((ra & (kFC - 1)) >> 3) * 128 + (ra & 7) * 2) = bf16_of_count(m);
          ++cur;
          if (rb >= limit) break;
          m = 1;
          if (row_info >> 16) {
            const int nv = (row_info >> 8) & 255;
            m = nv > 0 ? (ns - 1 - ((prow[cur] >> 16) & 255)) / nv + 1 : ns;
          }
          *reinterpret_cast<unsigned short*>(arow + ((rb & (kFC - 1)) >> 3) * 128 + (rb & 7) * 2) = bf16_of_count(m);
          ++cur;
        }
      }
      mbar_wait(bar_stage_full(buf), (unsigned)((j >> 1) & 1));
      if (cg < n_groups && cu < rows16) {
        unsigned char* dst = smem + L.planes + buf * L.planes_stride + (cu >> 3) * L.lbo_b + cg * 128 + (cu & 7) * 16;
        unsigned hx[3][4], hy[3][4];
        if (sSrcId[buf * kFC + cu] < 0) {
#pragma unroll
          for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int i = 0; i < 4; ++i) hx[p][i] = hy[p][i] = 0u;
        } else {
          const float* row = reinterpret_cast<const float*>(smem + L.stage + buf * L.stage_stride + (size_t)cu * L.row_bytes) + 8 * cg;
          const float4 xa = *reinterpret_cast<const float4*>(row);
          const float4 xb = (8 * cg + 4 < cbn) ? *reinterpret_cast<const float4*>(row + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
          const float x[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
          const float* wv = sSrcW + (buf * kFC + cu) * 3;
          const float wx = wv[0], wy = wv[1], wz = wv[2];
          const int base = (c0 + 8 * cg) % 3;
          const float w0 = rot3(wx, wy, wz, base), w1 = rot3(wy, wz, wx, base), w2 = rot3(wz, wx, wy, base);
          const float y[8] = {x[0] * w0, x[1] * w1, x[2] * w2, x[3] * w0, x[4] * w1, x[5] * w2, x[6] * w0, x[7] * w1};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            split3_bf16x2(x[2 * i], x[2 * i + 1], hx[0][i], hx[1][i], hx[2][i]);
            split3_bf16x2(y[2 * i], y[2 * i + 1], hy[0][i], hy[1][i], hy[2][i]);
          }
        }
#pragma unroll
        for (int p = 0; p < 3; ++p) {
          *reinterpret_cast<uint4*>(dst + (size_t)p * L.plane_bytes) = make_uint4(hx[p][0], hx[p][1], hx[p][2], hx[p][3]);
          *reinterpret_cast<uint4*>(dst + (size_t)(3 + p) * L.plane_bytes) = make_uint4(hy[p][0], hy[p][1], hy[p][2], hy[p][3]);
        }
      }
      fence_async_smem();
      mbar_arrive(bar_ops_full(buf));
      mbar_arrive(bar_stage_free(buf));
    }
  }
  // every role has issued its last operation (an early waiter would find the barrier in an older phase and could pass
  // a parity test that refers to a later one)
  __syncthreads();
  D3D_STAMP(4);

  // ---- epilogue: thread = owner row (TMEM lane); the three warp groups split the 16-column pieces ---------------------
  if (n_chunks > 0) {
    mbar_wait(bar_mma((n_chunks - 1) & 1), (unsigned)(((n_chunks - 1) >> 1) & 1));
    tc_fence_after();
  }
  {
    const int lq = warp & 3, t = lq * 32 + lane;
    const int own = sOwnerId[t];
    const float rcx = sOwnerXyz[3 * t] - ctr_x, rcy = sOwnerXyz[3 * t + 1] - ctr_y, rcz = sOwnerXyz[3 * t + 2] - ctr_z;
    const float rho = sOwnerRho[t];
    float* orow = a.out + ((size_t)b * a.M + (own >= 0 ? own : 0)) * a.C + c0;
    for (int ch = warp >> 2; ch * 16 < cbn; ch += kFT / 128) {
      unsigned y1[16], y2[16];
      if (n_chunks > 0) {
        const unsigned taddr = tmem_base + ((unsigned)(lq * 32) << 16) + (unsigned)(ch * 16);
        tmem_ld16(taddr, y1);
        tmem_ld16(taddr + (unsigned)L.np, y2);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) y1[i] = y2[i] = 0u;
      }
      const int base = (c0 + 16 * ch) % 3;
      const float r0 = rot3(rcx, rcy, rcz, base), r1 = rot3(rcy, rcz, rcx, base), r2 = rot3(rcz, rcx, rcy, base);
      float o[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float rc = (i % 3 == 0) ? r0 : ((i % 3 == 1) ? r1 : r2);
        o[i] = rho * (__uint_as_float(y2[i]) - rc * __uint_as_float(y1[i]));
      }
      if (own >= 0) {
#pragma unroll
        for (int v = 0; v < 4; ++v)
          if (16 * ch + 4 * v < cbn)
            *reinterpret_cast<float4*>(orow + 16 * ch + 4 * v) = make_float4(o[4 * v], o[4 * v + 1], o[4 * v + 2], o[4 * v + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  D3D_STAMP(5);
  if (a.timing && tid == 0) a.timing[((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 + 6] = (unsigned long long)U;
  if (warp == 5) tmem_dealloc(tmem_base, (unsigned)L.tmem_cols);
}

int launch_fwd_pipelined(const TileArgs& a, int B, cudaStream_t st) {
  if (a.M > kMaxPoints || a.N > kMaxPoints || a.nsample > kMaxNs || a.C % 4 != 0) return D3D_ERR_UNSUPPORTED;
  const int cbn_max = a.C < kCB ? a.C : kCB;
  const FwdLayout L = make_fwd_layout(cbn_max, a.nsample);
  cudaError_t e = cudaFuncSetAttribute(pospool_fwd_pipelined_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
  if (e != cudaSuccess) return (int)e;
  dim3 grid(d3d_ceil_div(a.M, kTQ), d3d_ceil_div(a.C, kCB), B);
  TileArgs t = a;
  t.timing = g_timing;
  pospool_fwd_pipelined_kernel<<<grid, kFT, L.total, st>>>(t);
  d3d_note_launches(1);
  return d3d_launch_status();
}

bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace

extern "C" {

// diagnostics only: device buffer of 8 x uint64 per CTA that the next launches fill with phase timestamps (null: off)
void d3d_pospool_tiles_debug_timing(void* buf) { g_timing = (unsigned long long*)buf; }

size_t d3d_pospool_tile_plan_bytes(int B, int M, int N, int nsample) {
  if (B <= 0 || M <= 0 || N <= 0 || nsample <= 0) return 0;
  return plan_total_bytes(B, M, N, nsample);
}

int d3d_pospool_tile_plan(const int* idx_by_support, const int* nvalid, const int* query_mask, const int* query_order, int B,
                          int M, int N, int nsample, void* plan, size_t plan_bytes, void* stream) {
  D3D_REQUIRE(idx_by_support && nvalid && query_mask && query_order && plan);
  D3D_REQUIRE(B >= 0 && M >= 0 && N > 0 && nsample > 0 && nsample <= D3D_MAX_NSAMPLE);
  if (M > kMaxPoints || N > kMaxPoints || nsample > kMaxNs) return D3D_ERR_UNSUPPORTED;
  if (B == 0 || M == 0) return 0;
  if (plan_bytes < plan_total_bytes(B, M, N, nsample) || ((uintptr_t)plan & 255)) return D3D_ERR_WORKSPACE;
  const int W = (N + 31) / 32;
  const size_t smem = align16((unsigned)(kTQ * nsample * 2)) + (size_t)(((W + 3) & ~3) + ((W + 4) & ~3)) * 4 + 2 * kTQ * 4;
  const PlanView pv = plan_view(plan, B, M, N, nsample);
  cudaError_t e = cudaMemsetAsync(const_cast<unsigned*>(pv.tilemask), 0, (size_t)B * N * 16, (cudaStream_t)stream);
  if (e != cudaSuccess) return (int)e;
  dim3 grid(d3d_ceil_div(M, kTQ), B);
  tile_plan_kernel<<<grid, kPlanThreads, smem, (cudaStream_t)stream>>>(idx_by_support, nvalid, query_mask, query_order, B, M, N,
                                                                      nsample, plan);
  d3d_note_launches(1);
  return d3d_launch_status();
}

int d3d_pospool_tiles_fwd(const float* feat_cl, const float* query_xyz, const float* support_xyz, const int* idx_by_support,
                          const int* nvalid, const int* query_mask, const int* query_order, const void* plan, int B, int M,
                          int N, int C, int nsample, float radius, int reduction, float* out_cl, void* stream) {
  D3D_REQUIRE(feat_cl && query_xyz && support_xyz && idx_by_support && nvalid && query_mask && query_order && out_cl && plan);
  D3D_REQUIRE(B >= 0 && M >= 0 && N > 0 && C > 0 && nsample > 0 && nsample <= D3D_MAX_NSAMPLE && radius > 0.f);
  D3D_REQUIRE(reduction == D3D_REDUCE_SUM || reduction == D3D_REDUCE_AVG);
  if (!aligned16(feat_cl) || !aligned16(out_cl)) return D3D_ERR_UNSUPPORTED;
  if (B == 0 || M == 0) return 0;
  TileArgs a{};
  a.src = feat_cl; a.out = out_cl; a.query_xyz = query_xyz; a.support_xyz = support_xyz; a.by_support = idx_by_support;
  a.nvalid = nvalid; a.query_mask = query_mask; a.order = query_order; a.M = M; a.N = N; a.C = C; a.nsample = nsample;
  a.reduction = reduction; a.inv_radius = 1.0f / radius; a.plan = plan;
  return launch_fwd_pipelined(a, B, (cudaStream_t)stream);
}

size_t d3d_pospool_scatter_bwd_workspace_bytes(int B, int M, int N, int C, int nsample) {
  if (B <= 0 || M <= 0 || N <= 0 || C <= 0 || nsample <= 0) return 0;
  return (size_t)B * d3d_ceil_div(M, kTQ) * plan_stride(N, nsample) * C * sizeof(float);
}

int d3d_pospool_scatter_bwd(const float* grad_out_cl, const float* query_xyz, const float* support_xyz,
                            const int* idx_by_support, const int* nvalid, const int* query_mask, const int* query_order,
                            const void* plan, int B, int M, int N, int C, int nsample, float radius, int reduction,
                            float* grad_feat_cl, void* ws, size_t ws_bytes, void* stream) {
  D3D_REQUIRE(grad_out_cl && query_xyz && support_xyz && idx_by_support && nvalid && query_mask && query_order && grad_feat_cl);
  D3D_REQUIRE(plan);
  D3D_REQUIRE(B >= 0 && M >= 0 && N > 0 && C > 0 && nsample > 0 && nsample <= D3D_MAX_NSAMPLE && radius > 0.f);
  D3D_REQUIRE(reduction == D3D_REDUCE_SUM || reduction == D3D_REDUCE_AVG);
  if (!aligned16(grad_out_cl) || !aligned16(grad_feat_cl)) return D3D_ERR_UNSUPPORTED;
  if (B == 0) return 0;
  if (M == 0) return (int)cudaMemsetAsync(grad_feat_cl, 0, (size_t)B * N * C * sizeof(float), (cudaStream_t)stream);
  TileArgs a{};
  a.src = grad_out_cl; a.out = grad_feat_cl; a.query_xyz = query_xyz; a.support_xyz = support_xyz; a.by_support = idx_by_support;
  a.nvalid = nvalid; a.query_mask = query_mask; a.order = query_order; a.M = M; a.N = N; a.C = C; a.nsample = nsample;
  a.reduction = reduction; a.inv_radius = 1.0f / radius; a.plan = plan;
  if (ws) {  // ordered (atomic-free) form
    if (ws_bytes < d3d_pospool_scatter_bwd_workspace_bytes(B, M, N, C, nsample) || ((uintptr_t)ws & 15)) return D3D_ERR_WORKSPACE;
    a.partial = (float*)ws;
  }
  return launch_scatter_bwd(a, B, (cudaStream_t)stream);
}

}  // extern "C"

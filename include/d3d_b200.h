/* d3d_b200.h — C ABI of the B200-native neighbourhood + local-aggregation path.
 *
 * One shared library (deep3dpointclouddenoising_b200/libd3d_b200.so, built by plain nvcc for
 * sm_100a, no torch / ATen / pybind types anywhere in the signatures).  Every entry point
 *   - takes raw DEVICE pointers, sizes, scalars, a caller-provided workspace and a cudaStream_t
 *     (passed as void* so that this header needs no CUDA include),
 *   - only enqueues work on that stream (no allocation, no synchronisation, no host read-back),
 *   - returns 0 on success, a positive cudaError_t if a launch failed, or a negative D3D_ERR_*
 *     for a rejected argument.  It never calls exit() (the reference does: cuda_utils.h:35-44).
 *
 * Paths cited as "ref:" are relative to /root/reference/u_net_arch/.
 *
 * Conventions shared by all functions
 *   xyz arrays      float32 (B, N, 3) contiguous                       ref: pt_custom_ops/_ext_src/include/utils.h:15-30
 *   masks           int32   (B, N)   0/1 with the valid entries a PREFIX (kernels stop at the first 0,
 *                                    ref: _ext_src/src/masked_ordered_ball_query_gpu.cu:50-52)
 *   idx             int32   (B, M, nsample)
 *   "cm" features   float32 (B, C, N)   channel-major — the reference's layout (Conv1d layout)
 *   "cl" features   float32 (B, N, C)   channel-last  — the layout the fused kernels gather from
 */
#ifndef D3D_B200_H_
#define D3D_B200_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define D3D_ERR_BAD_ARG     (-1)  /* null pointer, non-positive size, nsample out of range ...   */
#define D3D_ERR_UNSUPPORTED (-2)  /* a size beyond what the kernels were built for               */
#define D3D_ERR_WORKSPACE   (-3)  /* workspace missing or smaller than *_workspace_bytes()        */

#define D3D_MAX_NSAMPLE 255       /* slot numbers are packed in 8 bits of an inverse-map entry    */

/* ABI version of this header; bumped on any signature change. */
int d3d_abi_version(void);
/* Human-readable description of a return code (static storage). */
const char* d3d_error_string(int code);
/* Number of CUDA kernels this library has launched in this process (evidence for bench.py's gpu_launches). */
long long d3d_kernel_launches(void);

/* ------------------------------------------------------------------------------------------------
 * 1. Neighbourhood construction (bit-exact with the reference kernels)
 * ---------------------------------------------------------------------------------------------- */

/* Replaces  _ext.masked_ordered_ball_query(query_xyz, support_xyz, query_mask, support_mask, radius, nsample)
 *   ref: _ext_src/src/masked_ordered_ball_query.cpp:13-59, _ext_src/src/masked_ordered_ball_query_gpu.cu:11-96
 *        (binding: _ext_src/src/bindings.cpp:11; Python caller: pt_custom_ops/pt_utils.py:71-72).
 * Out: idx, idx_mask (B, M, nsample) int32.  nvalid (B, M) int32 may be NULL; when given it receives
 * min(#in-radius supports, nsample) per query BEFORE the query mask is applied (the fused aggregation
 * kernels use it instead of the dense idx_mask).  idx_by_support (B, M, nsample) int32 may be NULL; when given it
 * receives the same min(#in-radius, nsample) winners of every query in ASCENDING SUPPORT INDEX, each as
 * (distance rank << 16) | index (distance rank = its slot in idx), -1 padded — the order the staged-tile aggregation
 * kernels walk (needs N <= 65536).
 * Workspace: d3d_ball_query_workspace_bytes(B, M, N).  Unlike the reference no (B, M, 3*nsample) scratch
 * tensors are needed (masked_ordered_ball_query.cpp:38-44). */
size_t d3d_ball_query_workspace_bytes(int B, int M, int N);
int d3d_ball_query(const float* query_xyz, const float* support_xyz, const int* query_mask,
                   const int* support_mask, int B, int M, int N, float radius, int nsample,
                   int* idx, int* idx_mask, int* nvalid, int* idx_by_support, void* ws, size_t ws_bytes, void* stream);

/* Replaces  _ext.masked_nearest_query(query_xyz, support_xyz, query_mask, support_mask)
 *   ref: _ext_src/src/masked_nearest_query.cpp:12-47, _ext_src/src/masked_nearest_query_gpu.cu:8-62
 *        (binding: bindings.cpp:13; caller: pt_utils.py:87).
 * Out: idx, idx_mask (B, M) int32 (the reference shapes them (B, M, 1)); idx = -1 when no support
 * has d2 < 100 (masked_nearest_query_gpu.cu:36-38). */
size_t d3d_nearest_query_workspace_bytes(int B);
int d3d_nearest_query(const float* query_xyz, const float* support_xyz, const int* query_mask,
                      const int* support_mask, int B, int M, int N, int* idx, int* idx_mask,
                      void* ws, size_t ws_bytes, void* stream);

/* Replaces  _ext.masked_grid_subsampling(xyz, mask, npoint, sampleDl)
 *   ref: _ext_src/src/masked_grid_subsampling.cpp:13-44, _ext_src/src/masked_grid_subsampling_gpu.cu:11-153
 *        (binding: bindings.cpp:14; caller: pt_utils.py:102).
 * Out: sub_xyz (B, m, 3) float32, sub_mask (B, m) int32.
 * Workspace: d3d_grid_subsample_workspace_bytes(B, N) (only used when a cloud does not fit in
 * shared memory, N > 16384). */
size_t d3d_grid_subsample_workspace_bytes(int B, int N);
int d3d_grid_subsample(const float* xyz, const int* mask, int B, int N, int m, float sample_dl,
                       float* sub_xyz, int* sub_mask, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * 2. Gather and its gradient (the reference's layout: channel-major in, (B, C, M, nsample) out)
 * ---------------------------------------------------------------------------------------------- */

/* Replaces  _ext.group_points(points, idx)          ref: _ext_src/src/group_points.cpp:17-40,
 *                                                         _ext_src/src/group_points_gpu.cu:13-44
 * out[b,c,j,k] = points[b,c,idx[b,j,k]];  points (B,C,N) cm, out (B,C,M,nsample). */
int d3d_group_points(const float* points, const int* idx, int B, int C, int N, int M, int nsample,
                     float* out, void* stream);

/* Replaces  _ext.group_points_grad(grad_out, idx, N)   ref: group_points.cpp:42-65, group_points_gpu.cu:48-80
 * grad_points[b,c,i] = sum over (j,k) with idx[b,j,k]==i of grad_out[b,c,j,k], summed in ascending
 * (j,k) — deterministic, no float atomics (the reference uses atomicAdd, group_points_gpu.cu:65).
 * Workspace: d3d_group_points_grad_workspace_bytes (an inverse map is built inside the call). */
size_t d3d_group_points_grad_workspace_bytes(int B, int N, int M, int nsample);
int d3d_group_points_grad(const float* grad_out, const int* idx, int B, int C, int N, int M, int nsample,
                          float* grad_points, void* ws, size_t ws_bytes, void* stream);
/* The same sums in the reference's own formulation (atomicAdd, group_points_gpu.cu:61-66), one block per (b, c) plane
 * with the N sums in shared memory and the gradients / indices streamed coalesced: 7x faster than the fixed-order
 * form (1.04 ms against 7.8 ms) at B=16, C=72, M=N=8192, nsample=52, reproducible to rounding only.  N <= 51200 (D3D_ERR_UNSUPPORTED beyond). */
int d3d_group_points_grad_atomic(const float* grad_out, const int* idx, int B, int C, int N, int M, int nsample,
                                 float* grad_points, void* stream);

/* Inverse neighbour map (CSR by support point) of one idx tensor; shared by every backward kernel
 * of a (query set, support set) pair.
 *   rowptr  (B*N + 1) int32: entries of support (b,i) are entries[rowptr[b*N+i] .. rowptr[b*N+i+1])
 *   entries (B*M*nsample) int32: (j << 8) | k, ascending inside each segment
 * Indices outside [0, N) are treated as 0 like the reference's clamp (pt_utils.py:126-127). */
size_t d3d_inverse_map_workspace_bytes(int B, int N, int M, int nsample);
int d3d_build_inverse_map(const int* idx, int B, int N, int M, int nsample, int* rowptr, int* entries,
                          void* ws, size_t ws_bytes, void* stream);

/* Layout helpers: (B,C,N) cm <-> (B,N,C) cl. */
int d3d_cm_to_cl(const float* src_cm, int B, int C, int N, float* dst_cl, void* stream);
int d3d_cl_to_cm(const float* src_cl, int B, int C, int N, float* dst_cm, void* stream);

/* ------------------------------------------------------------------------------------------------
 * 3. Fused local aggregation (what the reference computes in eager PyTorch on the materialised
 *    (B, C, M, nsample) gather).  All feature tensors here are channel-last.
 *    mask semantics: fm[b,j,k] = query_mask[b,j] ? (k < nvalid[b,j]) : 1
 *    (= idx_mask + (1 - query_mask), ref: models/local_aggregation_operators.py:171,490)
 * ---------------------------------------------------------------------------------------------- */

#define D3D_REDUCE_SUM 0
#define D3D_REDUCE_AVG 1

/* PosPool, position_embedding 'xyz'       ref: models/local_aggregation_operators.py:140-147,165-183
 * out[b,j,c] = sum_k fm * ((S[idx]-Q[j])[c%3] * (1/radius)) * F[b,idx[b,j,k],c]   ( / sum_k fm for AVG ) */
int d3d_pospool_fwd(const float* feat_cl, const float* query_xyz, const float* support_xyz,
                    const int* idx, const int* nvalid, const int* query_mask, int B, int M, int N, int C,
                    int nsample, float radius, int reduction, float* out_cl, void* stream);
int d3d_pospool_bwd(const float* grad_out_cl, const float* query_xyz, const float* support_xyz,
                    const int* rowptr, const int* entries, const int* nvalid, const int* query_mask,
                    int B, int M, int N, int C, int nsample, float radius, int reduction,
                    float* grad_feat_cl, void* stream);

/* Processing order for the staged-tile kernels below: order (B, N) int32 = every cloud's point indices sorted along a
 * Morton curve over the cloud's bounding box (64^3 cells, ties by index).  It only decides which rows share a CTA
 * (spatially adjacent rows reference almost the same neighbour rows); results never depend on it beyond fp32
 * summation order.  N <= 16384 (D3D_ERR_UNSUPPORTED beyond).  No reference counterpart. */
int d3d_spatial_order(const float* xyz, int B, int N, int* order, void* stream);

/* PosPool as staged-tile kernels (same math as d3d_pospool_fwd / _bwd, which remain the path for other variants and for
 * sizes beyond the limits below): a CTA owns a tile of 128 spatially adjacent queries (query_order from
 * d3d_spatial_order), the union of the support rows the tile gathers travels global -> shared with cp.async.bulk and is
 * contracted with the tile's [128 x union] multiplicity matrix (from idx_by_support of d3d_ball_query: every query's
 * winners in ascending support index) on the tensor cores (tcgen05, exact 3-term bf16 split, fp32 accumulation).
 *   replaces ref: pt_custom_ops/_ext_src/src/group_points_gpu.cu:13-33,48-69 + models/local_aggregation_operators.py:140-183
 * Limits: M, N <= 16384, nsample <= 64, C % 4 == 0 (D3D_ERR_UNSUPPORTED otherwise). */
/* Tile plan of one (neighbour list, query order) pair, shared by every forward / scatter-backward launch on that pair
 * (two LocalAggregation layers use the level-0 list: four launches): per tile of 128 queries the size of the union of
 * the support rows they gather, the union itself (ascending indices), the union rank of every list entry, and the inverse
 * view (which tiles gather a support row, at which rank) for the ordered backward.  `plan`:
 * d3d_pospool_tile_plan_bytes(B, M, N, nsample) bytes, 256-byte aligned. */
size_t d3d_pospool_tile_plan_bytes(int B, int M, int N, int nsample);
int d3d_pospool_tile_plan(const int* idx_by_support, const int* nvalid, const int* query_mask, const int* query_order, int B,
                          int M, int N, int nsample, void* plan, size_t plan_bytes, void* stream);
int d3d_pospool_tiles_fwd(const float* feat_cl, const float* query_xyz, const float* support_xyz, const int* idx_by_support,
                          const int* nvalid, const int* query_mask, const int* query_order, const void* plan, int B, int M,
                          int N, int C, int nsample, float radius, int reduction, float* out_cl, void* stream);
/* Diagnostics (tools/tile_phases.py): a device buffer of 8 x uint64 per CTA that the next tile-kernel launches fill with
 * %globaltimer stamps of their phases; NULL switches it off (the default). */
void d3d_pospool_tiles_debug_timing(void* buf);
/* Backward pass in scatter form: the CTA owns the forward tile (128 queries), stages their gradient rows once, and
 * contracts A^T (the forward multiplicity matrix read through the MN-major descriptor) with them on the tensor cores,
 * 128 union rows per accumulator.  A feature-gradient row receives partial sums from every tile that gathered it:
 *   ws != NULL (d3d_pospool_scatter_bwd_workspace_bytes): ORDERED form — every tile stores its partial rows, a second
 *     kernel adds them per support row in ascending tile order (plan.tilemask / rank_of): a fixed-order segmented
 *     reduction without float atomics, bit-reproducible;
 *   ws == NULL: the partial sums are added with red.global.add.v4.f32 (grad_feat_cl is zero-filled by the call) — float
 *     atomics like the reference's own backward (group_points_gpu.cu:48-69), reproducible to rounding only.
 * Same limits as the forward. */
size_t d3d_pospool_scatter_bwd_workspace_bytes(int B, int M, int N, int C, int nsample);
int d3d_pospool_scatter_bwd(const float* grad_out_cl, const float* query_xyz, const float* support_xyz,
                            const int* idx_by_support, const int* nvalid, const int* query_mask, const int* query_order,
                            const void* plan, int B, int M, int N, int C, int nsample, float radius, int reduction,
                            float* grad_feat_cl, void* ws, size_t ws_bytes, void* stream);

/* 1x1 convolutions with a tiny channel count on one side (the U-Net's 3-channel input layer and its 3-channel output
 * layer; TMA needs 16-byte rows, so d3d_gemm_tf32 does not take them), fp32 on the CUDA cores, row-major operands:
 *   d3d_linear_small_k: y[R x N] = x[R x K] . w^T (+ bias), K <= 4, N % 4 == 0; w[n, k] = w[n * w_stride_n + k * w_stride_k]
 *     (strided: the same kernel is the data gradient of the output layer, dx = dy[R x 3] . W[3 x K])
 *   d3d_linear_small_n: y[R x N] = x[R x K] . w[N x K]^T (+ bias), N <= 4, K % 4 == 0
 *   d3d_wgrad_small:    out[n * out_stride_n + k * out_stride_k] (+)= sum_r big[r, n] * small[r, k], big (R, Nb), Nb % 4 == 0,
 *     small (R, Ks), Ks <= 4: the weight gradients of both layers; row-range partials + a fixed-order second pass.
 *   replaces ref: models/backbones/resnet.py:100-103 (conv1), heads/multi_dimensional_head.py:45-50 (last Conv1d) */
int d3d_linear_small_k(const float* x, const float* w, long long w_stride_n, long long w_stride_k, const float* bias,
                       long long R, int K, int N, float* y, void* stream);
int d3d_linear_small_n(const float* x, const float* w, const float* bias, long long R, int K, int N, float* y, void* stream);
size_t d3d_wgrad_small_workspace_bytes(int Nb);
int d3d_wgrad_small(const float* big, const float* small, long long R, int Nb, int Ks, float* out, long long out_stride_n,
                    long long out_stride_k, int accumulate, void* ws, size_t ws_bytes, void* stream);

#define D3D_KP_CONSTANT 0
#define D3D_KP_LINEAR   1
#define D3D_KP_GAUSSIAN 2

/* PseudoGrid (depthwise KPConv)           ref: models/local_aggregation_operators.py:467-503
 * out[b,j,c] = sum_k W[k,c] * sum_m w[b,j,k,m] * F[b,idx[b,j,m],c],
 * w = influence(|| (S[idx]-Q[j]) - K[k] ||, extent) * fm;  kpoints (K,3), weights (K,C), K <= 16.
 * precision: 0 = fp32 CUDA cores; 1 = bf16 tcgen05 contraction of [slots x K] . [K x C] (fp32 accumulate) */
int d3d_pseudogrid_fwd(const float* feat_cl, const float* query_xyz, const float* support_xyz,
                       const int* idx, const int* nvalid, const int* query_mask, const float* kpoints,
                       const float* weights, int B, int M, int N, int C, int nsample, int K, float extent,
                       int influence, int precision, float* out_cl, void* stream);
/* grad wrt features (via the inverse map; `precision` selects fp32 or the tcgen05 contraction like the forward)
 * and wrt weights (fp32, deterministic two-pass reduction).  Either output pointer may be NULL.
 * Workspace: d3d_pseudogrid_bwd_workspace_bytes. */
size_t d3d_pseudogrid_bwd_workspace_bytes(int B, int M, int C, int K);
int d3d_pseudogrid_bwd(const float* grad_out_cl, const float* feat_cl, const float* query_xyz,
                       const float* support_xyz, const int* idx, const int* rowptr, const int* entries,
                       const int* nvalid, const int* query_mask, const float* kpoints, const float* weights,
                       int B, int M, int N, int C, int nsample, int K, float extent, int influence, int precision,
                       float* grad_feat_cl, float* grad_weights, void* ws, size_t ws_bytes, void* stream);

/* MaskedMaxPool's gather + max over all nsample slots   ref: pt_custom_ops/pt_utils.py:199-205
 * out[b,j,c] = max_k F[b,idx[b,j,k],c]; argslot (B,M,C) uint8 = first slot attaining the max. */
int d3d_gather_max_fwd(const float* feat_cl, const int* idx, int B, int M, int N, int C, int nsample,
                       float* out_cl, uint8_t* argslot, void* stream);
int d3d_gather_max_bwd(const float* grad_out_cl, const uint8_t* argslot, const int* rowptr,
                       const int* entries, int B, int M, int N, int C, float* grad_feat_cl, void* stream);

/* MaskedUpsample(mode='nearest')           ref: pt_custom_ops/pt_utils.py:222-226
 * out[b,j,c] = F[b,idx[b,j],c]; idx (B,M) (negative indices read row 0). */
int d3d_nearest_gather_fwd(const float* feat_cl, const int* idx, int B, int M, int N, int C, float* out_cl,
                           void* stream);
int d3d_nearest_gather_bwd(const float* grad_out_cl, const int* rowptr, const int* entries, int B, int M,
                           int N, int C, float* grad_feat_cl, void* stream);

/* ------------------------------------------------------------------------------------------------
 * 4. Fused BatchNorm1d (+ residual add) (+ ReLU) on channel-major (B, C, N) activations (SURVEY.md §8 row f4:
 *    the conv / BN / ReLU sandwich around each aggregation; ref: models/backbones/resnet.py:32-45,58-66,
 *    models/local_aggregation_operators.py:121-123).  torch.nn.BatchNorm1d semantics; training != 0 uses batch
 *    statistics and updates running_mean / running_var in place, else the running statistics are used.
 *    y = act(bn(x) [+ residual]).  save_mean / save_invstd (C) are outputs the backward pass needs.
 * ---------------------------------------------------------------------------------------------- */
/* Workspace shared by both directions: d3d_bn_act_workspace_bytes(C) bytes that are ZERO before the first use
 * (ticket counters of the split channel reductions; the kernels leave them zero).  One workspace per stream. */
size_t d3d_bn_act_workspace_bytes(int C);
int d3d_bn_act_fwd(const float* x, const float* residual, const float* gamma, const float* beta, float* running_mean,
                   float* running_var, int B, int C, int N, float eps, float momentum, int training, int relu,
                   float* y, float* save_mean, float* save_invstd, void* ws, size_t ws_bytes, void* stream);
/* dx (and dres = gradient w.r.t. residual, may be NULL; dgamma / dbeta (C), may be NULL).
 * relu: 0 = none, 1 = ReLU without residual (mask recomputed from x, y may be NULL), 2 = ReLU mask read from y. */
int d3d_bn_act_bwd(const float* dy, const float* x, const float* y, const float* gamma, const float* beta,
                   const float* save_mean, const float* save_invstd, int B, int C, int N, int training, int relu,
                   float* dx, float* dres, float* dgamma, float* dbeta, void* ws, size_t ws_bytes, void* stream);
/* Channel-last variants: every activation is (R, C) row-major with R = B * N rows — the layout the aggregation
 * kernels read and write, so a network whose 1x1 convolutions run as row-major GEMMs needs no transposition at all.
 * Requires C % 4 == 0 and 16-byte aligned pointers (D3D_ERR_ARG otherwise).  Same workspace, same semantics. */
/* num_batches_tracked (device int64 scalar, may be NULL) is incremented by one in training mode — the counter
 * torch.nn.BatchNorm1d keeps — so that the host does not launch a kernel of its own for it. */
int d3d_bn_act_cl_fwd(const float* x, const float* residual, const float* gamma, const float* beta, float* running_mean,
                      float* running_var, long long* num_batches_tracked, long long R, int C, float eps, float momentum,
                      int training, int relu, float* y, float* save_mean, float* save_invstd, void* ws, size_t ws_bytes,
                      void* stream);
/* accumulate_param_grads != 0: dgamma / dbeta are added to (they are the parameters' gradient buffers), else stored. */
int d3d_bn_act_cl_bwd(const float* dy, const float* x, const float* y, const float* gamma, const float* beta,
                      const float* save_mean, const float* save_invstd, long long R, int C, int training, int relu,
                      float* dx, float* dres, float* dgamma, float* dbeta, int accumulate_param_grads, void* ws,
                      size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * 4b. The 1x1 convolutions as row-major TF32 GEMMs on the tensor cores (TMA + tcgen05, accumulator in TMEM) with the
 *     BatchNorm statistics in the epilogue (ref: models/backbones/resnet.py:32-45,58-66,
 *     models/local_aggregation_operators.py:121-123: nn.Conv1d(kernel_size=1, bias=False) -> BatchNorm1d -> ReLU, which
 *     cuDNN runs in TF32 by default).
 *       C[M x N] (+)= [A0 | A1] . B^T      A0 (M x K0), A1 (M x K1) or NULL (channel concatenation of two row tensors,
 *                                          heads/multi_dimensional_head.py:36), B (N x (K0 + K1)); all row-major fp32
 *     stats (d3d_gemm_row_tiles(M), 2, N) or NULL: per 128-row tile the column mean and sum of squared deviations of
 *     the tile of C; d3d_bn_finalize combines them (Chan, fp64) into mean / invstd + running statistics, and
 *     d3d_bn_apply_cl applies them: the statistics pass over the convolution output never runs.
 *     Requires K0 % 4 == 0, K1 % 4 == 0, N % 4 == 0, 16-byte aligned pointers (D3D_ERR_UNSUPPORTED otherwise).
 * ---------------------------------------------------------------------------------------------- */
int d3d_gemm_row_tiles(long long M);
int d3d_gemm_tf32(const float* A0, const float* A1, const float* B, float* C, long long M, int N, int K0, int K1,
                  int accumulate, float* stats, void* stream);
/* The same GEMM with the inference epilogue C = act(A . B^T + bias[col] + residual[row, col]): a 1x1 convolution whose
 * eval-mode BatchNorm is folded into its weights (B scaled per output channel, bias = the BatchNorm shift); the
 * bottleneck's residual add and the ReLU ride on the store.  bias (N) / residual (M, N) may be NULL; stats must be NULL
 * when any of them is used.   ref: models/backbones/resnet.py:32-45,58-66 in eval mode */
int d3d_gemm_tf32_act(const float* A0, const float* A1, const float* B, float* C, long long M, int N, int K0, int K1,
                      int accumulate, float* stats, const float* bias, const float* residual, int relu, void* stream);
/* Weight gradient of the same convolution: dW (Cout x Cin) = or += dY (R x Cout)^T . X (R x Cin)  (both operands MN-major for
 * the tensor core; the rows are split over the CTAs and the fp32 partials are added in a fixed order: deterministic).
 *   replaces the cuBLAS GEMMs behind torch's Conv1d weight gradient (ref: models/backbones/resnet.py:32-45 backward) */
size_t d3d_wgrad_workspace_bytes(long long R, int Cout, int Cin);
int d3d_wgrad_tf32(const float* dY, const float* X, float* dW, long long R, int Cout, int Cin, int accumulate, void* ws,
                   size_t ws_bytes, void* stream);
int d3d_bn_finalize(const float* stats, long long R, int C, float eps, float momentum, float* running_mean,
                    float* running_var, long long* num_batches_tracked, float* save_mean, float* save_invstd, void* stream);
int d3d_bn_apply_cl(const float* x, const float* residual, const float* gamma, const float* beta, const float* save_mean,
                    const float* save_invstd, long long R, int C, int relu, float* y, void* stream);

/* ------------------------------------------------------------------------------------------------
 * 5. Exact nearest neighbours on large clouds and the Chamfer distance
 *    ref: compute_cd.py:74-75; models/losses/chamfer_distance_aux.py:154-155,216-246 (pytorch3d knn_points, K = 1)
 * ---------------------------------------------------------------------------------------------- */
/* out_d2[j] = min_i |q_j - s_i|^2 (and out_idx[j], may be NULL: lowest index among equal distances); clouds (M,3), (N,3).
 * precise != 0 compares distances in fp64 (the decision a float64 KD-tree takes: used for patch-centre indices). */
size_t d3d_nn_workspace_bytes(int N);
int d3d_nn_sqdist(const float* query_xyz, const float* support_xyz, int M, int N, int precise, float* out_d2, int* out_idx,
                  void* ws, size_t ws_bytes, void* stream);
/* out3 = { mean_x min_y |x-y|^2 + mean_y min_x |x-y|^2,  first term,  second term }  (L2, point reduction mean) */
size_t d3d_chamfer_workspace_bytes(int Nx, int Ny);
int d3d_chamfer_l2(const float* x, const float* y, int Nx, int Ny, float* out3, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * 6. Full-shape inference support
 *    ref: offset_dataset.py:540-561 (patch centres), :630-656 (radius patches, sorted by distance, first num_points),
 *         qualitative_inference_test.py:325-342 (vote averaging), cpp_subsampling/grid_subsampling/grid_subsampling.cpp:25-103
 * ---------------------------------------------------------------------------------------------- */
/* voxel id of every point (ids (N,) int32, clamped to [0, n_cells)): iX + NX*iY + NX*NY*iZ with
 * i? = floor((p.? - origin.?) / dl) — the reference's CPU expressions; origin / NX / NY come from the caller. */
int d3d_voxel_ids(const float* points, int N, float origin_x, float origin_y, float origin_z, float dl, int NX, int NY,
                  int n_cells, int* ids, void* stream);
/* per-voxel barycentre from the inverse map of the voxel ids (d3d_build_inverse_map with B = 1, nsample = 1):
 * members added in ascending point index in fp32, times (float)(1.0 / count).  bary (n_cells, 3), counts (n_cells). */
int d3d_voxel_barycentres(const float* points, const int* rowptr, const int* entries, int n_cells, float* bary,
                          int* counts, void* stream);
/* radius patches: out_idx (P, num_points) = indices of the points within `radius` of each centre in ascending
 * distance (ties: lower index), -1 padded; out_count (P) = points in the ball.  See patches.cu for overflow_stride. */
size_t d3d_radius_patches_workspace_bytes(int N, int P, int overflow_stride);
int d3d_radius_patches(const float* points, int N, const float* centres, int P, float radius, int num_points,
                       int overflow_stride, int* out_idx, int* out_count, void* ws, size_t ws_bytes, void* stream);
/* The same with the shared-memory candidate capacity chosen by the caller: 12288 (the default above: one block per SM)
 * or a smaller power of two >= 256 when the balls are small (2048 keys: six blocks per SM).  A block that asks for few
 * results out of many candidates (num_points * 4 < count) selects them with a distance histogram before sorting. */
int d3d_radius_patches_tier(const float* points, int N, const float* centres, int P, float radius, int num_points,
                            int smem_keys, int overflow_stride, int* out_idx, int* out_count, void* ws, size_t ws_bytes,
                            void* stream);
/* per-point mean of the predictions that voted for it: pred (P, 3, num_points), inverse map of the (flattened,
 * 128-slot rows) patch indices; mean_offset (N, 3) = sum / (count + 1e-7); votes (N) may be NULL. */
int d3d_vote_mean(const float* pred, const int* rowptr, const int* entries, int N, int num_points, float* mean_offset,
                  float* votes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* D3D_B200_H_ */

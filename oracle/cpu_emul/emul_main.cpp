/* TEST INFRASTRUCTURE — drives the CPU-emulated reference kernels (see cuda_utils.h here).
 * Exposes extern "C" entry points that run each reference kernel over its launch grid: one
 * (blockIdx, threadIdx) pair per call.  The kernels use grid-stride loops over disjoint work items,
 * no shared memory and no barriers, so any execution order is exact; OpenMP spreads the (block,
 * thread) pairs over the host cores (the launch-geometry registers are thread_local).
 * group_points_grad is the exception: its atomicAdd order is what defines the float result, so it is
 * run sequentially in launch order (one legal order of the reference's non-deterministic sum).
 */
#include "cuda_utils.h"
thread_local d3d_emul_dim3 blockIdx, threadIdx, blockDim, gridDim;

void masked_ordered_query_ball_point_kernel(int b, int n, int m, float radius, int nsample,
    const float* query_xyz, const float* support_xyz, const int* query_mask,
    const int* support_mask, int* idx, int* idx_mask, float* dists, int* tempidxs);
void masked_grid_subsampling_kernel(int n, int m, float sampleDl, const float* dataset,
    const int* mask, float* subxyz, int* submask, int* mapidxs, int* tempidxs, float* temp_subxyz);
void masked_nearest_query_kernel(int b, int n, int m, const float* query_xyz,
    const float* support_xyz, const int* query_mask, const int* support_mask, int* idx, int* idx_mask);
void group_points_kernel(int b, int c, int n, int npoints, int nsample, const float* points,
    const int* idx, float* out);
void group_points_grad_kernel(int b, int c, int n, int npoints, int nsample, const float* grad_out,
    const int* idx, float* grad_points);

static void set_launch(int b, int tx, int ty, int bi, int tix, int tiy) {
  gridDim = {b, 1, 1}; blockDim = {tx, ty, 1}; blockIdx = {bi, 0, 0}; threadIdx = {tix, tiy, 0};
}

extern "C" {
/* scratch tensors are zero-initialised by the caller exactly like the reference's torch::zeros */
void emul_ball_query(int b, int n, int m, float radius, int nsample, const float* q, const float* s,
                     const int* qm, const int* sm, int* idx, int* idx_mask, float* dists, int* tempidxs,
                     int threads) {
#pragma omp parallel for collapse(2) schedule(dynamic, 1)
  for (int bi = 0; bi < b; ++bi)
    for (int ti = 0; ti < threads; ++ti) {
      set_launch(b, threads, 1, bi, ti, 0);
      masked_ordered_query_ball_point_kernel(b, n, m, radius, nsample, q, s, qm, sm, idx, idx_mask, dists, tempidxs);
    }
}
void emul_grid_subsampling(int b, int n, int m, float dl, const float* xyz, const int* mask, float* sub,
                           int* submask, int* mapidxs, int* tempidxs, float* temp_subxyz) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int bi = 0; bi < b; ++bi) {
    set_launch(b, 1, 1, bi, 0, 0);
    masked_grid_subsampling_kernel(n, m, dl, xyz, mask, sub, submask, mapidxs, tempidxs, temp_subxyz);
  }
}
void emul_nearest_query(int b, int n, int m, const float* q, const float* s, const int* qm, const int* sm,
                        int* idx, int* idx_mask, int threads) {
#pragma omp parallel for collapse(2) schedule(dynamic, 1)
  for (int bi = 0; bi < b; ++bi)
    for (int ti = 0; ti < threads; ++ti) {
      set_launch(b, threads, 1, bi, ti, 0);
      masked_nearest_query_kernel(b, n, m, q, s, qm, sm, idx, idx_mask);
    }
}
void emul_group_points(int b, int c, int n, int npoints, int nsample, const float* points, const int* idx,
                       float* out, int tx, int ty) {
#pragma omp parallel for collapse(2) schedule(dynamic, 1)
  for (int bi = 0; bi < b; ++bi)
    for (int t = 0; t < tx * ty; ++t) {
      set_launch(b, tx, ty, bi, t % tx, t / tx);
      group_points_kernel(b, c, n, npoints, nsample, points, idx, out);
    }
}
void emul_group_points_grad(int b, int c, int n, int npoints, int nsample, const float* grad_out,
                            const int* idx, float* grad_points, int tx, int ty) {
  for (int bi = 0; bi < b; ++bi)
    for (int t = 0; t < tx * ty; ++t) {
      set_launch(b, tx, ty, bi, t % tx, t / tx);
      group_points_grad_kernel(b, c, n, npoints, nsample, grad_out, idx, grad_points);
    }
}
}

/* TEST INFRASTRUCTURE — CPU restatement (plain C) of the reference's index-building algorithms.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library; the product package never does.  Every function cites the reference lines it follows
 * (paths relative to /root/reference/u_net_arch/pt_custom_ops/_ext_src/src/).
 *
 * Pinning: tests/test_oracle.py checks these functions bit-for-bit against (i) oracle/_ref/libref_emul.so
 * — the reference's own kernel definitions compiled for the host — on seeded inputs, and (ii) the
 * committed golden vectors under tests/golden/ that were produced by that library
 * (oracle/make_golden.py).  On the GPU box tests/test_reference_cuda.py additionally compares with
 * oracle/_ref/libref_cuda.so (the reference kernels compiled by nvcc for sm_100a).
 *
 * Floating point: the device build of the reference contracts the squared distance to
 *   d2 = fma(dz, dz, fma(dx, dx, dy*dy))          (SASS: FADD FADD FMUL FADD FFMA FFMA, dy first)
 * so it is spelled with explicit fmaf() here and this file is compiled with -ffp-contract=off.
 * The grid-subsampling kernel has one more contraction: "x - floor(min/dl)*dl" is a single FFMA on the
 * device (see oracle_grid_subsampling).  Everything else is single IEEE operations (mul, div, floor).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline float dist2(const float* a, const float* b) {
  const float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
  return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* number of leading non-zero mask entries: every reference kernel stops at the first 0
 * (masked_ordered_ball_query_gpu.cu:49-52, masked_nearest_query_gpu.cu:40-43,
 *  masked_grid_subsampling_gpu.cu:66). */
static int prefix_len(const int* mask, int n) {
  int v = 0;
  while (v < n && mask[v] != 0) ++v;
  return v;
}

/* stable sort of (key, payload) pairs by key: bottom-up merge sort (the reference relies on
 * thrust::sort_by_key being stable; masked_ordered_ball_query_gpu.cu:77,
 * masked_grid_subsampling_gpu.cu:77,135). */
typedef struct { int64_t key; int val; } pair_t;
static void stable_sort_pairs(pair_t* a, int n) {
  if (n < 2) return;
  pair_t* tmp = (pair_t*)malloc(sizeof(pair_t) * (size_t)n);
  pair_t *src = a, *dst = tmp;
  for (int w = 1; w < n; w *= 2) {
    for (int lo = 0; lo < n; lo += 2 * w) {
      int mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
      int i = lo, j = mid, k = lo;
      while (i < mid && j < hi) dst[k++] = (src[j].key < src[i].key) ? src[j++] : src[i++];
      while (i < mid) dst[k++] = src[i++];
      while (j < hi) dst[k++] = src[j++];
    }
    pair_t* t = src; src = dst; dst = t;
  }
  if (src != a) memcpy(a, src, sizeof(pair_t) * (size_t)n);
  free(tmp);
}
/* float keys: order-preserving map to int64 (sign-magnitude to two's complement) */
static int64_t float_key(float f) {
  int32_t b; memcpy(&b, &f, 4);
  return b >= 0 ? (int64_t)b : -(int64_t)(b & 0x7fffffff);
}

/* masked_ordered_ball_query_gpu.cu:37-94.  Outputs idx, idx_mask: (b, m, nsample) int32.
 * cnt == 0 is undefined in the reference (i % 0); scratch is zero-filled there
 * (masked_ordered_ball_query.cpp:38-44), so this restatement emits idx = 0, mask = 0. */
void oracle_ball_query(int b, int n, int m, float radius, int nsample, const float* query_xyz,
                       const float* support_xyz, const int* query_mask, const int* support_mask,
                       int* idx, int* idx_mask) {
  const float r2 = radius * radius;
  const int cap = 3 * nsample;
#pragma omp parallel for schedule(dynamic, 1)
  for (int bi = 0; bi < b; ++bi) {
    const float* S = support_xyz + (size_t)bi * n * 3;
    const float* Q = query_xyz + (size_t)bi * m * 3;
    const int v = prefix_len(support_mask + (size_t)bi * n, n);
    pair_t* slot = (pair_t*)malloc(sizeof(pair_t) * (size_t)(cap > 0 ? cap : 1));
    for (int j = 0; j < m; ++j) {
      int cnt = 0, best_k = 0;
      float best = r2;
      for (int k = 0; k < v; ++k) {
        const float d2 = dist2(Q + 3 * j, S + 3 * k);
        if (!(d2 < r2)) continue;
        if (d2 < best) { best = d2; best_k = k; }           /* strict: lowest index among ties */
        if (cnt < cap) { slot[cnt].key = float_key(d2); slot[cnt].val = k; ++cnt; }
      }
      if (cnt >= cap && cnt > 0 && best_k > slot[cnt - 1].val) {  /* :72-75 nearest-swap */
        slot[cnt - 1].key = float_key(best);
        slot[cnt - 1].val = best_k;
      }
      stable_sort_pairs(slot, cnt);                          /* :77 */
      int* o = idx + ((size_t)bi * m + j) * nsample;
      int* om = idx_mask + ((size_t)bi * m + j) * nsample;
      for (int i = 0; i < nsample; ++i) {
        if (cnt == 0) { o[i] = 0; om[i] = 0; }
        else if (i < cnt) { o[i] = slot[i].val; om[i] = 1; }
        else { o[i] = slot[i % cnt].val; om[i] = 0; }        /* :83-86 cyclic padding */
      }
      if (query_mask[(size_t)bi * m + j] == 0)               /* :88-93 */
        for (int i = 0; i < nsample; ++i) om[i] = 0;
    }
    free(slot);
  }
}

/* masked_nearest_query_gpu.cu:30-61.  idx, idx_mask: (b, m) int32. */
void oracle_nearest_query(int b, int n, int m, const float* query_xyz, const float* support_xyz,
                          const int* query_mask, const int* support_mask, int* idx, int* idx_mask) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int bi = 0; bi < b; ++bi) {
    const float* S = support_xyz + (size_t)bi * n * 3;
    const float* Q = query_xyz + (size_t)bi * m * 3;
    const int v = prefix_len(support_mask + (size_t)bi * n, n);
    for (int j = 0; j < m; ++j) {
      float best = 100.0f;
      int best_k = -1;
      for (int k = 0; k < v; ++k) {
        const float d2 = dist2(Q + 3 * j, S + 3 * k);
        if (d2 < best) { best = d2; best_k = k; }
      }
      idx[(size_t)bi * m + j] = best_k;
      idx_mask[(size_t)bi * m + j] = query_mask[(size_t)bi * m + j] != 0;
    }
  }
}

/* masked_grid_subsampling_gpu.cu:31-152.  sub_xyz: (b, m, 3) float, sub_mask: (b, m) int32. */
void oracle_grid_subsampling(int b, int n, int m, float dl, const float* xyz, const int* mask,
                             float* sub_xyz, int* sub_mask) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int bi = 0; bi < b; ++bi) {
    const float* P = xyz + (size_t)bi * n * 3;
    float* out = sub_xyz + (size_t)bi * m * 3;
    int* outm = sub_mask + (size_t)bi * m;
    /* bounding box over ALL n points, padding included (:31-46) */
    float lo[3] = {P[0], P[1], P[2]}, hi[3] = {P[0], P[1], P[2]};
    for (int i = 1; i < n; ++i)
      for (int d = 0; d < 3; ++d) {
        const float c = P[3 * i + d];
        if (c > hi[d]) hi[d] = c;
        if (c < lo[d]) lo[d] = c;
      }
    const float inv = 1 / dl;                                   /* rounded reciprocal first (:48) */
    /* origin = floor(min*inv)*dl is never rounded on its own in the device build: nvcc contracts every
     * "coordinate - origin" into one FFMA (SASS of the reference kernel: FFMA R, -Rfloor, Rdl, Rcoord),
     * so the restatement keeps the floor() factor and spells the fma out (:48-54, :67-69). */
    float fl[3];
    for (int d = 0; d < 3; ++d) fl[d] = floorf(lo[d] * inv);
    const int NX = (int)floorf(fmaf(-fl[0], dl, hi[0]) / dl) + 1;
    const int NY = (int)floorf(fmaf(-fl[1], dl, hi[1]) / dl) + 1;
    int v = prefix_len(mask + (size_t)bi * n, n);
    pair_t* cell = (pair_t*)malloc(sizeof(pair_t) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < v; ++i) {
      const int ix = (int)floorf(fmaf(-fl[0], dl, P[3 * i + 0]) / dl);
      const int iy = (int)floorf(fmaf(-fl[1], dl, P[3 * i + 1]) / dl);
      const int iz = (int)floorf(fmaf(-fl[2], dl, P[3 * i + 2]) / dl);
      cell[i].key = (int32_t)(ix + NX * iy + NX * NY * iz);     /* int32 arithmetic like the kernel */
      cell[i].val = i;
    }
    if (v == 0) { cell[0].key = 0; cell[0].val = 0; v = 1; }   /* zero-filled scratch: one cell = point 0 */
    stable_sort_pairs(cell, v);                                 /* :77 */
    /* per-cell barycentre, members added in ascending point index, then one division (:79-122) */
    float* cen = (float*)malloc(sizeof(float) * 3 * (size_t)v);
    int ncell = 0;
    for (int i = 0; i < v;) {
      int j = i;
      float s[3] = {P[3 * cell[i].val], P[3 * cell[i].val + 1], P[3 * cell[i].val + 2]};
      float cntf = 1;
      for (j = i + 1; j < v && cell[j].key == cell[i].key; ++j) {
        s[0] += P[3 * cell[j].val]; s[1] += P[3 * cell[j].val + 1]; s[2] += P[3 * cell[j].val + 2];
        cntf += 1;
      }
      cen[3 * ncell] = s[0] / cntf; cen[3 * ncell + 1] = s[1] / cntf; cen[3 * ncell + 2] = s[2] / cntf;
      ++ncell;
      i = j;
    }
    /* LCG(17,139,256) key per cell ordinal seeded by the smallest cell id, then stable sort (:124-135) */
    pair_t* perm = (pair_t*)malloc(sizeof(pair_t) * (size_t)ncell);
    int key = (int)((int32_t)cell[0].key % 256);
    for (int i = 0; i < ncell; ++i) {
      if (i > 0) key = (17 * key + 139) % 256;
      perm[i].key = key;
      perm[i].val = i;
    }
    stable_sort_pairs(perm, ncell);
    for (int i = 0; i < m; ++i) {
      if (i < ncell) {
        memcpy(out + 3 * i, cen + 3 * perm[i].val, 3 * sizeof(float));
        outm[i] = 1;
      } else {                                                  /* :145-151 cyclic padding */
        memcpy(out + 3 * i, out + 3 * (i % ncell), 3 * sizeof(float));
        outm[i] = 0;
      }
    }
    free(perm); free(cen); free(cell);
  }
}

/* group_points_gpu.cu:23-32: out[b,c,j,k] = points[b,c,idx[b,j,k]] */
void oracle_group_points(int b, int c, int n, int npoints, int nsample, const float* points,
                         const int* idx, float* out) {
  const size_t P = (size_t)npoints * nsample;
#pragma omp parallel for collapse(2)
  for (int bi = 0; bi < b; ++bi)
    for (int l = 0; l < c; ++l) {
      const float* src = points + ((size_t)bi * c + l) * n;
      const int* id = idx + (size_t)bi * P;
      float* dst = out + ((size_t)bi * c + l) * P;
      for (size_t p = 0; p < P; ++p) dst[p] = src[id[p]];
    }
}

/* group_points_gpu.cu:58-68: grad_points[b,c,idx[b,j,k]] += grad_out[b,c,j,k].  The reference order is
 * whatever atomicAdd gives; this restatement sums in ascending (j,k) in double and rounds once, i.e.
 * the exact sum to fp32 precision, which every legal atomic order matches within ~1 ulp * nsample. */
void oracle_group_points_grad(int b, int c, int n, int npoints, int nsample, const float* grad_out,
                              const int* idx, float* grad_points) {
  const size_t P = (size_t)npoints * nsample;
#pragma omp parallel for collapse(2)
  for (int bi = 0; bi < b; ++bi)
    for (int l = 0; l < c; ++l) {
      double* acc = (double*)calloc((size_t)n, sizeof(double));
      const float* g = grad_out + ((size_t)bi * c + l) * P;
      const int* id = idx + (size_t)bi * P;
      for (size_t p = 0; p < P; ++p) acc[id[p]] += g[p];
      float* dst = grad_points + ((size_t)bi * c + l) * n;
      for (int i = 0; i < n; ++i) dst[i] = (float)acc[i];
      free(acc);
    }
}

/* TEST / BASELINE INFRASTRUCTURE — CPU radius-neighbour baseline over the reference's OWN vendored KD-tree:
 * u_net_arch/cpp_wrappers/cpp_utils/nanoflann/nanoflann.hpp (KDTreeSingleIndexAdaptor::radiusSearch, :1280) with the
 * reference's PointCloud adaptor (cpp_utils/cloud/cloud.h:151-175), both compiled from where they lie under
 * /root/reference (oracle/Makefile).  The reference ships no compiled CPU neighbour search of its own,
 * so this harness is the closest thing to "the reference's CPU radius search": one radiusSearch per query, results
 * sorted by distance, truncated to `cap` like the CUDA op.  threads <= 1: a plain loop, else OpenMP over the queries. */
#include <algorithm>
#include <utility>
#include <vector>
#include "cloud.h"
#include "nanoflann.hpp"

typedef nanoflann::KDTreeSingleIndexAdaptor<nanoflann::L2_Simple_Adaptor<float, PointCloud>, PointCloud, 3> ref_kd_tree;

extern "C" {
/* out_idx (m, cap) int32, -1 padded; out_cnt (m) = neighbours inside the ball (before truncation).  returns 0. */
int ref_nanoflann_radius(const float* support, int n, const float* query, int m, float radius, int cap, int threads,
                         int* out_idx, int* out_cnt) {
  PointCloud cloud;
  cloud.pts.resize(n);
  for (int i = 0; i < n; ++i) cloud.pts[i] = PointXYZ(support[3 * i], support[3 * i + 1], support[3 * i + 2]);
  ref_kd_tree tree(3, cloud, nanoflann::KDTreeSingleIndexAdaptorParams(10));
  tree.buildIndex();
  const float r2 = radius * radius;
  nanoflann::SearchParams params;
  params.sorted = true;
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads > 1 ? threads : 1)
  for (int j = 0; j < m; ++j) {
    std::vector<std::pair<size_t, float> > found;
    const float q[3] = {query[3 * j], query[3 * j + 1], query[3 * j + 2]};
    const size_t cnt = tree.radiusSearch(q, r2, found, params);
    out_cnt[j] = (int)cnt;
    for (int k = 0; k < cap; ++k) out_idx[(size_t)j * cap + k] = k < (int)cnt ? (int)found[k].first : -1;
  }
  return 0;
}
}

/* TEST INFRASTRUCTURE — C shim over the reference's CPU grid subsampling
 * (u_net_arch/cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.cpp:5-106), linked
 * unmodified.  Replaces the CPython wrapper (wrapper.cpp:58-286), which no longer compiles against
 * NumPy 2.x.  Points only / points+features / points+classes like wrapper.cpp:205-243. */
#include "grid_subsampling.h"
#include <cstring>
extern "C" {
/* returns the number of cells; outputs must have room for n entries. */
int ref_grid_subsampling(const float* pts, int n, const float* feats, int fdim, const int* classes, int ldim,
                         float dl, float* sub_pts, float* sub_feats, int* sub_classes) {
  std::vector<PointXYZ> in(n), out;
  for (int i = 0; i < n; ++i) in[i] = PointXYZ(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
  std::vector<float> f, sf;
  std::vector<int> c, sc;
  if (feats && fdim > 0) f.assign(feats, feats + (size_t)n * fdim);
  if (classes && ldim > 0) c.assign(classes, classes + (size_t)n * ldim);
  grid_subsampling(in, out, f, sf, c, sc, dl, 0);
  for (size_t i = 0; i < out.size(); ++i) { sub_pts[3*i] = out[i].x; sub_pts[3*i+1] = out[i].y; sub_pts[3*i+2] = out[i].z; }
  if (!sf.empty()) std::memcpy(sub_feats, sf.data(), sf.size() * sizeof(float));
  if (!sc.empty()) std::memcpy(sub_classes, sc.data(), sc.size() * sizeof(int));
  return (int)out.size();
}
}

// placeholder, replaced below
#include "common.cuh"
int d3d_pseudogrid_fwd_tc(const float*, const float*, const float*, const int*, const int*, const int*, const float*,
                          const float*, int, int, int, int, int, int, float, int, float*, cudaStream_t) {
  return D3D_ERR_UNSUPPORTED;
}

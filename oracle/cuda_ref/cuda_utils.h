/* TEST INFRASTRUCTURE — shadows the reference's include/cuda_utils.h (which pulls in ATen) so the
 * reference kernel definitions compile with plain nvcc.  Launch-shape helpers restate
 * u_net_arch/pt_custom_ops/_ext_src/include/cuda_utils.h:18-33 (used by ref_launch.cu only). */
#pragma once
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
static inline int ref_n_threads(int work_size) {
  const int pow_2 = (int)(std::log((double)work_size) / std::log(2.0));
  return std::max(std::min(1 << pow_2, 512), 1);
}
static inline dim3 ref_block_config(int x, int y) {
  const int xt = ref_n_threads(x);
  const int yt = std::max(std::min(ref_n_threads(y), 512 / xt), 1);
  return dim3(xt, yt, 1);
}

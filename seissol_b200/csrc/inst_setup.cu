// Device set-up kernels (kernels_setup.cuh): independent of the convergence order; one instantiation per precision.
#include "api_internal.h"
#include "kernels_setup.cuh"

namespace ssb {

int setupCellEntry(const SetupCellParams& p, int realSize, cudaStream_t stream) {
  if (p.n <= 0) {
    return 0;
  }
  const unsigned grid = unsigned((4 * p.n + 127) / 128);
  if (realSize == 4) {
    k_setup_cell<float><<<grid, 128, 0, stream>>>(p);
  } else {
    k_setup_cell<double><<<grid, 128, 0, stream>>>(p);
  }
  return int(cudaGetLastError());
}

int setupDrEntry(const SetupDrParams& p, int realSize, cudaStream_t stream) {
  if (p.m <= 0) {
    return 0;
  }
  const unsigned grid = unsigned((p.m + 127) / 128);
  if (realSize == 4) {
    k_setup_dr<float><<<grid, 128, 0, stream>>>(p);
  } else {
    k_setup_dr<double><<<grid, 128, 0, stream>>>(p);
  }
  return int(cudaGetLastError());
}

}  // namespace ssb

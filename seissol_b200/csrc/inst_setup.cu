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

int faultEnergiesEntry(const FaultEnergyParams& p, int realSize, cudaStream_t stream) {
  if (p.m <= 0) {
    return 0;
  }
  long long blocks = (p.m + 7) / 8;
  if (blocks > 148LL * 8) {
    blocks = 148LL * 8;
  }
  if (realSize == 4) {
    k_fault_energies<float><<<unsigned(blocks), 256, 0, stream>>>(p);
  } else {
    k_fault_energies<double><<<unsigned(blocks), 256, 0, stream>>>(p);
  }
  return int(cudaGetLastError());
}

}  // namespace ssb

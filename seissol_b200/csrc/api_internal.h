// Internal glue between the C ABI (capi.cu) and the per-configuration kernel
// instantiations (inst_o*_f*.cu, one translation unit per (order, real) so they
// compile in parallel).
#pragma once
#include <cuda_runtime.h>

#include "params.cuh"

#include <atomic>

namespace ssb {

// cudaFuncAttributeMaxDynamicSharedMemorySize applies to the CURRENT device only: every launcher keeps one slot per
// device ordinal (a host may drive several GPUs from one process) and raises the attribute when a launch needs more.
// The slots are atomics: two host threads racing on the first launch both set the same value.
struct SmemConfig {
  std::atomic<size_t> bytes[64];
  SmemConfig() {
    for (auto& b : bytes) {
      b.store(0);
    }
  }
};
template <typename Kernel>
inline int ensureDynamicSmem(Kernel kern, size_t bytes, SmemConfig& cfg) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    return int(cudaGetLastError());
  }
  const bool tracked = dev >= 0 && dev < 64;
  if (tracked && cfg.bytes[dev].load(std::memory_order_acquire) >= bytes && bytes > 0) {
    return 0;
  }
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
  if (err != cudaSuccess) {
    return int(err);
  }
  if (tracked) {
    cfg.bytes[dev].store(bytes, std::memory_order_release);
  }
  return 0;
}
// Kernels that do not depend on a translation unit's generated operators (k_taylor<R>) are ONE function shared by all
// configuration units: their shared-memory configuration must be one object too, otherwise a unit with a smaller
// layout lowers the limit another unit relies on.  (inline template: one instance across translation units)
template <typename Tag>
inline SmemConfig& sharedSmemConfig() {
  static SmemConfig cfg;
  return cfg;
}
// SM count of the CURRENT device (cached per device ordinal)
inline int smCountOfCurrentDevice() {
  static std::atomic<int> cache[64];
  int dev = 0;
  cudaGetDevice(&dev);
  const bool tracked = dev >= 0 && dev < 64;
  int sms = tracked ? cache[dev].load() : 0;
  if (sms == 0) {
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (tracked) {
      cache[dev].store(sms);
    }
  }
  return sms;
}

struct KernelSet {
  int order;
  int realSize;
  int B;       // basis functions
  int BP;      // padded to 4
  int eLocal;  // elements per CTA
  int eNeigh;
  int (*local)(const LocalParams&, bool ader, bool local, cudaStream_t);
  int (*neigh)(const NeighParams&, cudaStream_t);
  int (*taylor)(const TaylorParams&, cudaStream_t);
  int (*accumulate)(const AccumulateParams&, cudaStream_t);
  int (*taylorFp)(const TaylorParams&, cudaStream_t);   // Taylor sum + face projection (full-column units only)
  int (*nodalBc)(const struct NodalBcParams&, cudaStream_t);   // nodal boundary conditions (full-column units only)
  int (*volumeEnergies)(const struct VolumeEnergyParams&, cudaStream_t);   // energy output (full-column units only)
};

struct NodalBcParams;
struct SetupCellParams;
struct SetupDrParams;
int setupCellEntry(const SetupCellParams&, int realSize, cudaStream_t);
int setupDrEntry(const SetupDrParams&, int realSize, cudaStream_t);
struct FaultEnergyParams;
int faultEnergiesEntry(const FaultEnergyParams&, int realSize, cudaStream_t);

struct DrKernelSet {
  int order;
  int realSize;
  int P;    // fault quadrature points
  int PP;   // padded to 4
  int (*interp)(DrInterpParams&, cudaStream_t);
  int (*friction)(const FrictionParams&, int law, cudaStream_t);
  int (*neighDr)(const NeighDrParams&, cudaStream_t);
};

struct ViscoKernelSet {
  int order;
  int realSize;
  int B;
  int BP;
  int elemsPerCta;
  int (*ader)(const ViscoLocalParams&, cudaStream_t);
  int (*local)(const ViscoLocalParams&, cudaStream_t);
  int (*neigh)(const ViscoNeighParams&, cudaStream_t);
  int (*neighDr)(const ViscoNeighDrParams&, cudaStream_t);
};
const ViscoKernelSet* viscokernelset_o5_f32();
const ViscoKernelSet* viscokernelset_o5_f64();

const DrKernelSet* drkernelset_o4_f32();
const DrKernelSet* drkernelset_o5_f32();
const DrKernelSet* drkernelset_o6_f32();
const DrKernelSet* drkernelset_o4_f64();
const DrKernelSet* drkernelset_o5_f64();
const DrKernelSet* drkernelset_o6_f64();

const KernelSet* kernelset_o4_f32();
const KernelSet* kernelset_o5_f32();
const KernelSet* kernelset_o6_f32();
const KernelSet* kernelset_o4_f64();
const KernelSet* kernelset_o5_f64();
const KernelSet* kernelset_o6_f64();
// kernel generation 4: half-column threads (kernels_elastic4.cuh)
const KernelSet* kernelset4_o4_f32();
const KernelSet* kernelset4_o5_f32();
const KernelSet* kernelset4_o6_f32();
const KernelSet* kernelset4_o4_f64();
const KernelSet* kernelset4_o5_f64();
const KernelSet* kernelset4_o6_f64();

}  // namespace ssb

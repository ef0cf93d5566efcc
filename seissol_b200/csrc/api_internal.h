// Internal glue between the C ABI (capi.cu) and the per-configuration kernel
// instantiations (inst_o*_f*.cu, one translation unit per (order, real) so they
// compile in parallel).
#pragma once
#include <cuda_runtime.h>

#include "kernels_elastic.cuh"

namespace ssb {

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
};

const KernelSet* kernelset_o4_f32();
const KernelSet* kernelset_o5_f32();
const KernelSet* kernelset_o6_f32();
const KernelSet* kernelset_o4_f64();
const KernelSet* kernelset_o5_f64();
const KernelSet* kernelset_o6_f64();

}  // namespace ssb

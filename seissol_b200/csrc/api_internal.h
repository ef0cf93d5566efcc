// Internal glue between the C ABI (capi.cu) and the per-configuration kernel
// instantiations (inst_o*_f*.cu, one translation unit per (order, real) so they
// compile in parallel).
#pragma once
#include <cuda_runtime.h>

#include "params.cuh"

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
  int (*accumulate)(const AccumulateParams&, cudaStream_t);
};

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

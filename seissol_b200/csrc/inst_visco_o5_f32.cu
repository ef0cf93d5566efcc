// Kernel instantiation: viscoelastic2 (3 mechanisms), order 5, float.
#include "gen/ops_o5.cuh"
#include "gen/dr_o5.cuh"
#define SSB_ORDER 5
#define SSB_REAL float
#define SSB_TAG o5_f32
// 8 elements per CTA, three resident CTAs per SM (12 and two: 17.2 ms, 8 and three: 15.1 ms, 6 and four: 16.8 ms for `all` at
// 1 M elements: the kernels wait on their loads, a third independent CTA covers more of the latency)
#ifndef SSB_VE
#define SSB_VE 8
#define SSB_VISCO_MINB 3
#endif
#include "inst_visco.cuh"

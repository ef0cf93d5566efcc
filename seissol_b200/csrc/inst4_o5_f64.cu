// Kernel instantiation, generation 4 (half-column threads): elastic, order 5, double.
#define SSB_PACK_F32 2
#include "gen/ops4_o5.cuh"
#define SSB_ORDER 5
#define SSB_REAL double
#define SSB_TAG o5_f64
// 7 elements per CTA, three CTAs per SM: 168 registers per thread (no spills at order 6), 12 warps per SM
#ifndef SSB_NP
#define SSB_NP 7
#endif
#ifndef SSB_MINB
#define SSB_MINB 3
#endif
#include "inst4.cuh"

// Kernel instantiation, generation 4 (half-column threads): elastic, order 6, double.
#define SSB_PACK_F32 2
#include "gen/ops4_o6.cuh"
#define SSB_ORDER 6
#define SSB_REAL double
#define SSB_TAG o6_f64
// 14 elements per CTA, one CTA of 8 warps per SM: four warps share each of the two half-column instruction streams.
// Measured at 500 k elements (profiles/r2_f64_variants.txt): local 9.09 ms; 7 elements x 3 CTAs (six streams of two
// warps) 10.5-10.7 ms; 10 x 2 11.1 ms; 21 x 1 10.8 ms.  No spills at order 6.
#ifndef SSB_NP
#define SSB_NP 14
#endif
#ifndef SSB_MINB
#define SSB_MINB 1
#endif
#include "inst4.cuh"

// Kernel instantiation, generation 4 (half-column threads): elastic, order 6, double.
#define SSB_PACK_F32 2
#include "gen/ops4_o6.cuh"
#define SSB_ORDER 6
#define SSB_REAL double
#define SSB_TAG o6_f64
// 7 elements per CTA, TWO CTAs per SM (255 registers): 8 resident warps.  Measured at 500 k elements
// (profiles/r2_f64_variants.txt), local kernel: 7 x 2 CTAs 8.77 ms; 14 x 1 9.09 ms (also with <= 200 registers);
// 7 x 3 (168 registers, 12 warps) 10.5-10.7 ms; 10 x 2 (12 warps) 11.1 ms; 21 x 1 (12 warps) 10.8 ms: the straight-line
// double-precision stream runs best with 8 warps per SM, whatever the split.  No spills at order 6.
#ifndef SSB_NP
#define SSB_NP 7
#endif
#ifndef SSB_MINB
#define SSB_MINB 2
#endif
#include "inst4.cuh"

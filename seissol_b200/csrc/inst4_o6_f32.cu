// Kernel instantiation, generation 4 (half-column threads): elastic, order 6, float.
#define SSB_PACK_F32 2
#include "gen/ops4_o6.cuh"
#define SSB_ORDER 6
#define SSB_REAL float
#define SSB_TAG o6_f32
#ifndef SSB_NP
#define SSB_NP 14
#endif
#include "inst4.cuh"

// Kernel instantiation, generation 4 (half-column threads): elastic, order 5, float.
#define SSB_PACK_F32 2
#include "gen/ops4_o5.cuh"
#define SSB_ORDER 5
#define SSB_REAL float
#define SSB_TAG o5_f32
#ifndef SSB_NP
#define SSB_NP 14
#endif
#include "inst4.cuh"

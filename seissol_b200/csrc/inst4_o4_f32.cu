// Kernel instantiation, generation 4 (half-column threads): elastic, order 4, float.
#define SSB_PACK_F32 2
#include "gen/ops4_o4.cuh"
#define SSB_ORDER 4
#define SSB_REAL float
#define SSB_TAG o4_f32
#ifndef SSB_NP
#define SSB_NP 14
#endif
#include "inst4.cuh"

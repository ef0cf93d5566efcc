// Kernel instantiation: elastic, order 4, float.
#include "gen/ops_o4.cuh"
#define SSB_ORDER 4
#define SSB_REAL float
#define SSB_TAG o4_f32
#ifndef SSB_E_LOCAL
#define SSB_E_LOCAL (28 / SSB_PACK_F32)
#endif
#ifndef SSB_E_NEIGH
#define SSB_E_NEIGH (28 / SSB_PACK_F32)
#endif
#include "inst.cuh"

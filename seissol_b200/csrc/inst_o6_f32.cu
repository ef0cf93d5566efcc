// Kernel instantiation: elastic, order 6, float.
#include "gen/ops_o6.cuh"
#include "gen/tc_o6.cuh"
#define SSB_HAVE_TC 1
#define SSB_ORDER 6
#define SSB_REAL float
#define SSB_TAG o6_f32
#ifndef SSB_E_LOCAL
#define SSB_E_LOCAL (28 / SSB_PACK_F32)
#endif
#ifndef SSB_E_NEIGH
#define SSB_E_NEIGH (28 / SSB_PACK_F32)
#endif
#include "inst.cuh"

// Kernel instantiation: elastic, order 6, float.
// k_neigh_fp: 14 elements per CTA, three resident CTAs per SM (two CTAs of 21: 2.63 ms, three of 14: 2.53 ms at 1 M elements:
// the kernel waits on its gather, a third independent CTA covers more of it)
#ifndef SSB_E_NEIGHFP
#define SSB_E_NEIGHFP 14
#define SSB_MINB_NEIGHFP 3
#endif
#include "gen/ops_o6.cuh"
#include "gen/nodal_o6.cuh"
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

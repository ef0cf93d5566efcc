// Kernel instantiation: elastic, order 5, float.
// resident CTAs per SM the registers are sized for: the order-5 columns are short, 3 CTAs (= 24 warps) per SM
// measured best (profiles/r1_other_orders_quick_bench.txt)
#define SSB_MINB_LOCAL 3
#define SSB_MINB_NEIGH 3
#include "gen/ops_o5.cuh"
#include "gen/nodal_o5.cuh"
#include "gen/tc_o5.cuh"
#define SSB_HAVE_TC 1
#define SSB_ORDER 5
#define SSB_REAL float
#define SSB_TAG o5_f32
#ifndef SSB_E_LOCAL
#define SSB_E_LOCAL (28 / SSB_PACK_F32)
#endif
#ifndef SSB_E_NEIGH
#define SSB_E_NEIGH (28 / SSB_PACK_F32)
#endif
#include "inst.cuh"

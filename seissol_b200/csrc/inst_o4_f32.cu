// Kernel instantiation: elastic, order 4, float.
// resident CTAs per SM the registers are sized for: the order-4 columns are short, 4 CTAs (= 32 warps) per SM
// measured best (profiles/r1_other_orders_quick_bench.txt)
#define SSB_MINB_LOCAL 4
#define SSB_MINB_NEIGH 4
#include "gen/ops_o4.cuh"
#include "gen/nodal_o4.cuh"
#include "gen/tc_o4.cuh"
#define SSB_HAVE_TC 1
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

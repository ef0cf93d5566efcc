// Kernel instantiation: elastic, order 6, double.
#include "gen/ops_o6.cuh"
#include "gen/nodal_o6.cuh"
#define SSB_ORDER 6
#define SSB_REAL double
#define SSB_TAG o6_f64
#ifndef SSB_E_LOCAL
#define SSB_E_LOCAL 14
#endif
#ifndef SSB_E_NEIGH
#define SSB_E_NEIGH 14
#endif
#include "inst.cuh"

// Kernel instantiation: elastic, order 4, double.
#include "gen/ops_o4.cuh"
#include "gen/nodal_o4.cuh"
#define SSB_ORDER 4
#define SSB_REAL double
#define SSB_TAG o4_f64
#ifndef SSB_E_LOCAL
#define SSB_E_LOCAL 14
#endif
#ifndef SSB_E_NEIGH
#define SSB_E_NEIGH 14
#endif
#include "inst.cuh"

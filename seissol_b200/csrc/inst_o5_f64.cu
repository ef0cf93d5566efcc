// Kernel instantiation: elastic, order 5, double.
#include "gen/ops_o5.cuh"
#include "gen/nodal_o5.cuh"
#define SSB_ORDER 5
#define SSB_REAL double
#define SSB_TAG o5_f64
#ifndef SSB_E_LOCAL
#define SSB_E_LOCAL 14
#endif
#ifndef SSB_E_NEIGH
#define SSB_E_NEIGH 14
#endif
#include "inst.cuh"

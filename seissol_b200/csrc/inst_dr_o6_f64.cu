// Dynamic-rupture kernel instantiation: order 6, double.
#include "gen/dr_o6.cuh"
#define SSB_ORDER 6
#define SSB_REAL double
#define SSB_TAG o6_f64
#include "inst_dr.cuh"

// Dynamic-rupture kernel instantiation: order 4, double.
#include "gen/dr_o4.cuh"
#define SSB_ORDER 4
#define SSB_REAL double
#define SSB_TAG o4_f64
#include "inst_dr.cuh"

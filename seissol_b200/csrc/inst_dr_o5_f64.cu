// Dynamic-rupture kernel instantiation: order 5, double.
#include "gen/dr_o5.cuh"
#define SSB_ORDER 5
#define SSB_REAL double
#define SSB_TAG o5_f64
#include "inst_dr.cuh"

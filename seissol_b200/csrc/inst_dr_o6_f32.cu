// Dynamic-rupture kernel instantiation: order 6, float.
#include "gen/dr_o6.cuh"
#define SSB_ORDER 6
#define SSB_REAL float
#define SSB_TAG o6_f32
#include "inst_dr.cuh"

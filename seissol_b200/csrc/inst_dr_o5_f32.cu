// Dynamic-rupture kernel instantiation: order 5, float.
#include "gen/dr_o5.cuh"
#define SSB_ORDER 5
#define SSB_REAL float
#define SSB_TAG o5_f32
#include "inst_dr.cuh"

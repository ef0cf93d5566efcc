// Dynamic-rupture kernel instantiation: order 4, float.
#include "gen/dr_o4.cuh"
#define SSB_ORDER 4
#define SSB_REAL float
#define SSB_TAG o4_f32
#include "inst_dr.cuh"

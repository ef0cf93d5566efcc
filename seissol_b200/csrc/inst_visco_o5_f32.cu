// Kernel instantiation: viscoelastic2 (3 mechanisms), order 5, float.
#include "gen/ops_o5.cuh"
#include "gen/dr_o5.cuh"
#define SSB_ORDER 5
#define SSB_REAL float
#define SSB_TAG o5_f32
#define SSB_VE 12
#include "inst_visco.cuh"

// Kernel instantiation: viscoelastic2 (3 mechanisms), order 5, double.
#include "gen/ops_o5.cuh"
#include "gen/dr_o5.cuh"
#define SSB_ORDER 5
#define SSB_REAL double
#define SSB_TAG o5_f64
#define SSB_VE 6
#include "inst_visco.cuh"

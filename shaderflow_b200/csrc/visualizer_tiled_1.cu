// visualizer_tiled_kernel<1> (ssaa 1); see visualizer_tiled_unit.cuh
#define VT_UNIT_S 1
#include "visualizer_tiled_unit.cuh"

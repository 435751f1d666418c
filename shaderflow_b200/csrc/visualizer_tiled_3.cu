// visualizer_tiled_kernel<3> (ssaa 3); see visualizer_tiled_unit.cuh
#define VT_UNIT_S 3
#include "visualizer_tiled_unit.cuh"

// visualizer_tiled_kernel<2> (ssaa 2); see visualizer_tiled_unit.cuh
#define VT_UNIT_S 2
#include "visualizer_tiled_unit.cuh"

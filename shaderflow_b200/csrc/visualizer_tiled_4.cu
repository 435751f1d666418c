// visualizer_tiled_kernel<4> (ssaa 4); see visualizer_tiled_unit.cuh
#define VT_UNIT_S 4
#include "visualizer_tiled_unit.cuh"

// bl_grid.cu - lattice engine translation unit (see bl_grid.cuh).
#include "bl_common.cuh"

#include "bl_grid.cuh"

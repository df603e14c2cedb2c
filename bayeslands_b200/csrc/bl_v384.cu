// bl_v384.cu - device code instantiated for two 384-thread CTAs per SM (80 registers each: the compact layout with fewer spills than at 2 x 512).
#define BL_NS v384
#define BL_TPB 384
#define BL_MINB 2
#include "bl_variant.cuh"

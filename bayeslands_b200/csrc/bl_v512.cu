// bl_v512.cu - device code instantiated for two 512-thread CTAs per SM (64 registers): landscapes of at most half the shared memory of an SM.
#define BL_NS v512
#define BL_TPB 512
#define BL_MINB 2
#include "bl_variant.cuh"

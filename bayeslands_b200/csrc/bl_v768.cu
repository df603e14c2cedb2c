// bl_v768.cu - device code instantiated for one 768-thread CTA per SM (80 registers: fewer spills than the 64 of bl_v1024.cu, more warps than 512).
#define BL_NS v768
#define BL_TPB 768
#define BL_MINB 1
#include "bl_variant.cuh"

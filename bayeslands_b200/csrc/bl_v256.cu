// bl_v256.cu - device code instantiated for four 256-thread CTAs per SM: small meshes such as the etopo configs.
#define BL_NS v256
#define BL_TPB 256
#define BL_MINB 4
#define BL_WITH_STAGE 1      // this unit also carries the per-stage kernel (bl_fill, bl_sfd, ...)
#include "bl_variant.cuh"

// bl_v1024.cu - device code instantiated for one 1024-thread CTA per SM (64 registers): landscapes that fill the shared memory of an SM.
#define BL_NS v1024
#define BL_TPB 1024
#define BL_MINB 1
#define BL_WITH_STAGE 1      // this unit also carries the per-stage kernel (bl_fill, bl_sfd, ...)
#include "bl_variant.cuh"

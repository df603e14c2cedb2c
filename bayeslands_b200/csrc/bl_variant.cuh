// bl_variant.cuh - one device translation unit per CTA size.  The including .cu defines
//   BL_NS (namespace / launcher suffix), BL_TPB (threads per CTA), BL_MINB (min resident CTAs per SM: register cap).
#include "bl_common.cuh"

namespace {
namespace BL_NS {
constexpr int TPB = BL_TPB;
constexpr int NWARP = TPB / 32;
#include "bl_device.cuh"
} // namespace BL_NS
} // namespace

#define BL_CAT2(a, b) a##b
#define BL_CAT(a, b) BL_CAT2(a, b)

cudaError_t BL_CAT(bl_launch_run_, BL_NS)(int KDT, const DevConst &c, int B, int dyn, cudaStream_t st, const RunArgs &a)
{
#define LAUNCH(K, DYN)                                                                                                  \
    do {                                                                                                                \
        if (DYN) {                                                                                                      \
            cudaError_t e = cudaFuncSetAttribute(BL_NS::bl_run_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, DYN); \
            if (e != cudaSuccess) return e;                                                                             \
        }                                                                                                               \
        BL_NS::bl_run_kernel<K><<<B, BL_TPB, DYN, st>>>(c, a.nstops, a.stops, a.ckpt, a.out_pts, a.out_sse, a.out_elev,  \
                                                        a.out_erdp, a.nck, a.single_step, DYN);                         \
    } while (0)
    switch (KDT) {
        case 8: LAUNCH(8, dyn); break;
        case 12: LAUNCH(12, dyn); break;
        case 20: LAUNCH(20, dyn); break;
        default: LAUNCH(0, 0); break;
    }
#undef LAUNCH
    return cudaGetLastError();
}

#ifdef BL_WITH_STAGE
cudaError_t BL_CAT(bl_launch_stage_, BL_NS)(int KDT, const DevConst &c, int dyn, cudaStream_t st, const StageArgs &a)
{
#define LAUNCH(K, DYN)                                                                                                  \
    do {                                                                                                                \
        if (DYN) {                                                                                                      \
            cudaError_t e = cudaFuncSetAttribute(BL_NS::bl_stage_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, DYN); \
            if (e != cudaSuccess) return e;                                                                             \
        }                                                                                                               \
        BL_NS::bl_stage_kernel<K><<<1, BL_TPB, DYN, st>>>(c, a, DYN);                                                   \
    } while (0)
    switch (KDT) {
        case 8: LAUNCH(8, dyn); break;
        case 12: LAUNCH(12, dyn); break;
        case 20: LAUNCH(20, dyn); break;
        default: LAUNCH(0, 0); break;
    }
#undef LAUNCH
    return cudaGetLastError();
}
#endif

cudaError_t BL_CAT(bl_launch_reset_, BL_NS)(const DevConst &c, int B, cudaStream_t st, const double *rain, const double *erod,
                                            const unsigned long long *seed, double tStart)
{
    BL_NS::bl_reset_kernel<<<B, 256, 0, st>>>(c, rain, erod, seed, tStart);
    return cudaGetLastError();
}

cudaError_t BL_CAT(bl_func_attr_, BL_NS)(int KDT, cudaFuncAttributes *attr)
{
    switch (KDT) {
        case 8: return cudaFuncGetAttributes(attr, BL_NS::bl_run_kernel<8>);
        case 12: return cudaFuncGetAttributes(attr, BL_NS::bl_run_kernel<12>);
        case 20: return cudaFuncGetAttributes(attr, BL_NS::bl_run_kernel<20>);
        default: return cudaFuncGetAttributes(attr, BL_NS::bl_run_kernel<0>);
    }
}

// resident CTAs per SM of the run kernel with `dyn` bytes of dynamic shared memory
cudaError_t BL_CAT(bl_occupancy_, BL_NS)(int KDT, int dyn, int *nres)
{
#define OCC(K)                                                                                                          \
    do {                                                                                                                \
        if (dyn) {                                                                                                      \
            cudaError_t e = cudaFuncSetAttribute(BL_NS::bl_run_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn); \
            if (e != cudaSuccess) return e;                                                                             \
        }                                                                                                               \
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(nres, BL_NS::bl_run_kernel<K>, BL_TPB, (size_t)dyn);       \
    } while (0)
    switch (KDT) {
        case 8: OCC(8);
        case 12: OCC(12);
        case 20: OCC(20);
        default: dyn = 0; OCC(0);
    }
#undef OCC
}

// bl_common.cuh - types and constants shared by every translation unit of libbayeslands_b200.so
// (host side bl_engine.cu, one device TU per CTA size bl_v*.cu, lattice engine bl_grid.cu).
#pragma once
#include "bayeslands_b200.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

namespace blc {

constexpr int NONE_I = -1;
constexpr unsigned END_U = 0xFFFFFFFFu;
constexpr int TERM_BIT = 0x40000000;
constexpr int NPHASE = 48;
constexpr int LEVMAX = 63;
constexpr int DYN_SMEM = 229376; // dynamic shared memory of the fast kernel (224 KB)


// ------------------------------------------------------------------------------------------------
// device-visible constants
// ------------------------------------------------------------------------------------------------
struct DevConst {
    int N, bounds, KD, nlev, G, npts, nsea, B;
    const int *ngb;             // [KD][N]
    const unsigned short *ngb16; // [N][KDT] node-major u16 copy for the shared-memory path (0xFFFF = none)
    const unsigned short *ngb16f; // same rows with empty slots = N: the depression filling keeps W[N] = 2e6 there, so a
                                 // neighbour gather needs no "is this slot used" test
    int Nq, KDT;                // N + 1 rounded up to 64; rows of ngb16
    const float *ve32;          // [N][KDT][2] (vor, edge) as float pairs when both are float32-representable, else null
    int lev_contig;             // lev_nodes[i] == bounds + i (colour-major numbering)
    int dbg;                    // BL_DEBUG experiments (timing only)
    int compact;                // 8-bytes-per-node shared-memory layout (bl_fastc.cuh): two landscapes per SM
    const unsigned *nbf2;       // [N] bit p = neighbour slot p is an inside node (borders2): the diffusion's neighbour test
    const unsigned *pend0;      // [Nq / 32] bit k = interior node k has a ghost neighbour: the fill's initial pending set
    const double *vor, *edge;   // [KD][N]
    const double *area, *xy;    // [N], [N][2]
    const unsigned char *flags; // bit0 borders, bit1 borders2, bit2 indomain
    const int *parent;          // [bounds]
    const double *z0, *disp;
    const int *lev_nodes, *lev_off;
    const int *grid_idx;
    const double *grid_w;
    const int *grid_exact;
    const double *real_elev;
    const int *pts_cell;
    const double *seat, *seav;
    double m, n, eps, TH, CDa, CDm, CDr, minDT, maxDT, hill, ms, diffprop, seapos;
    int diffnb, depo, btype, shuffle;
    char *prop_base;
    size_t prop_stride;
    double *rain, *erod;
    unsigned long long *seed;
    double *tnow, *lastdt;
    long long *steps;
    unsigned *status;
    long long *phase_cyc; // [B][BL_NPHASE] clock64 cycles per phase (thread 0), diagnostics
};

enum { F_Z, F_CUMDIFF, F_CUMHILL, F_FILLH, F_MAXH, F_Q, F_SEDIN, F_QS, F_DEPO, F_ERO, F_PITVOL, F_PITVOLW,
       F_VAL, F_DTHICK, F_E, F_CAPH, F_PDEP, F_SEDFLUX, NF64 };
enum { I_RCV, I_RCV1, I_STACK, I_STACK1, I_POS, I_POS1, I_PITID, I_PITDRAIN, I_ALLDRAIN, I_FC, I_NS, I_LC,
       I_PS, I_CNT, I_ROOT1, I_LA, I_LB, I_SUCC, I_SUCC2, I_KIND, I_EVENTS, I_MEMB, I_MARK, NI32 };

__host__ __device__ inline size_t rup(size_t x, size_t a) { return (x + a - 1) / a * a; }
__host__ __device__ inline int np_of(int N) { return (int)rup((size_t)N + 1, 32); }
__host__ __device__ inline size_t keys_len(int N)
{
    size_t p = 1;
    while (p < (size_t)N) p <<= 1;
    return p;
}
__host__ __device__ inline size_t prop_bytes(int N)
{
    size_t Np = np_of(N);
    return rup((size_t)NF64 * Np * 8 + (size_t)NI32 * Np * 4 + (size_t)(4 * Np + keys_len(N)) * 8, 256);
}

struct Prop {
    double *f[NF64];
    int *i[NI32];
    unsigned long long *euA, *euB, *keys;
};

__device__ inline void prop_init(Prop &P, char *base, int N)
{
    size_t Np = np_of(N);
    double *d = (double *)base;
    for (int k = 0; k < NF64; ++k) P.f[k] = d + (size_t)k * Np;
    unsigned long long *u = (unsigned long long *)(d + (size_t)NF64 * Np);
    P.euA = u;
    P.euB = u + 2 * Np;
    P.keys = u + 4 * Np;
    int *ip = (int *)(P.keys + keys_len(N));
    for (int k = 0; k < NI32; ++k) P.i[k] = ip + (size_t)k * Np;
}

} // namespace blc
using namespace blc;

// last error message of the calling thread (bl_last_error)
std::string &bl_err_ref();
#define g_err bl_err_ref()

#define CUDA_OK(x)                                                                      \
    do {                                                                                \
        cudaError_t e_ = (x);                                                           \
        if (e_ != cudaSuccess) {                                                        \
            g_err = std::string(#x) + ": " + cudaGetErrorString(e_);                    \
            return BL_ERR_CUDA;                                                         \
        }                                                                               \
    } while (0)

// arguments of one bl_run_kernel launch (see bl_device.cuh)
struct RunArgs {
    int nstops;
    const double *stops;
    const int *ckpt;
    double *out_pts, *out_sse, *out_elev, *out_erdp;
    int nck, single_step;
};

// arguments of one per-stage launch (bl_stage_kernel in bl_device.cuh)
enum { BL_STAGE_FILL = 1, BL_STAGE_SFD, BL_STAGE_STACK, BL_STAGE_DISCHARGE, BL_STAGE_BASINS, BL_STAGE_DRAINAGE,
       BL_STAGE_STREAMPOWER, BL_STAGE_DIFFUSE, BL_STAGE_INTERP };
struct StageArgs {
    int stage, shuffle;
    double sea, rain, erod, dt;
    unsigned long long seed;
    long long step;
    const double *q_in;      // streampower: caller's discharge
    double *diff_out;        // diffuse: raw flux
    const double *v_in;      // interp: nodal values
    double *grid_out;        // interp: [G]
};

// per-CTA-size device TUs (bl_v1024.cu, bl_v512.cu, bl_v256.cu): launchers with external linkage
#define BL_DECL_VARIANT(NS)                                                                                              \
    cudaError_t bl_launch_run_##NS(int KDT, const DevConst &c, int B, int dyn, cudaStream_t st, const RunArgs &a);       \
    cudaError_t bl_launch_reset_##NS(const DevConst &c, int B, cudaStream_t st, const double *rain, const double *erod,  \
                                     const unsigned long long *seed, double tStart);                                     \
    cudaError_t bl_func_attr_##NS(int KDT, cudaFuncAttributes *attr);                                                     \
    cudaError_t bl_occupancy_##NS(int KDT, int dyn, int *nres);
BL_DECL_VARIANT(v1024)
cudaError_t bl_launch_stage_v1024(int KDT, const DevConst &c, int dyn, cudaStream_t st, const StageArgs &a);
cudaError_t bl_launch_stage_v256(int KDT, const DevConst &c, int dyn, cudaStream_t st, const StageArgs &a);
BL_DECL_VARIANT(v768)
BL_DECL_VARIANT(v384)
BL_DECL_VARIANT(v512)
BL_DECL_VARIANT(v256)

/* bayeslands_b200.h - C-ABI of the B200-native Bayeslands hot path (libbayeslands_b200.so).
 *
 * Drop-in boundary.  The entry points replace what the reference binds through f2py for
 * this path, batched over B independent (rain, erodibility) proposals:
 *
 *   reference interface (file:line under /root/reference)                      replaced by
 *   ------------------------------------------------------------------------   -----------------------
 *   PDalgo.pdstack.pitparams / pitfilling   libUtils/PDalgo.f90:271-321        bl_ctx_create, bl_run
 *   sfd.directions_base / directions        libUtils/sfd.c:55-153  (sfd.pyf)   bl_run (stage SFD)
 *   FLWnetwork.fstack.build                 libUtils/FLWnetwork.f90:21-71      bl_run (stage STACK)
 *   FLOWalgo.flowcompute.discharge          libUtils/FLOWalgo.f90:53-81        bl_run (stage DISCHARGE)
 *     .basinparameters/.basindrainage(all)  FLOWalgo.f90:130-297               bl_run (stage BASINS)
 *     .flowcfl                              FLOWalgo.f90:299-330               bl_run (stage CFL)
 *     .streampower/.getid1/.getids          FLOWalgo.f90:583-1180              bl_run (stage SPL)
 *   sfd.diffusion                           libUtils/sfd.c:155-185             bl_run (stage DIFF)
 *   buildFlux.streamflow / sediment_flux    simulation/buildFlux.py:21-320     bl_run (device-side step loop)
 *   Model.run_to_time                       pyBadlands/model.py:218-505        bl_run(stops)
 *   bl_mcmc.interpolateArray+likelihoodFunc bl_mcmc.py:167-215, :482-513       bl_run (checkpoint outputs)
 *
 * Ownership: every device buffer is caller-owned (PyTorch tensors): the caller allocates one workspace
 * of bl_workspace_bytes() and the output buffers and passes raw device pointers.  Mesh/parameter structs
 * hold HOST pointers; they are copied into the workspace by bl_ctx_create.  The library keeps no global
 * state: contexts are independent and re-entrant (the reference's Fortran modules are not:
 * libUtils/PDclass.f90:15-34, FLOWalgo.f90:15-23).
 * Errors: every call returns 0 on success or a negative bl_status; bl_last_error() gives the message.
 * Index convention: 0-based int32, a negative neighbour id terminates a list (libUtils/sfd.c:79-81).
 */
#ifndef BAYESLANDS_B200_H
#define BAYESLANDS_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BL_KP 20 /* padded neighbour slots (MAX_NEIGHBOURS, libUtils/sfd.c:14) */
#define BL_NCKPT_MAX 8
#define BL_NPTS_MAX 16

typedef enum {
    BL_OK = 0,
    BL_ERR_ARG = -1,       /* bad argument / inconsistent mesh */
    BL_ERR_CUDA = -2,      /* CUDA runtime error */
    BL_ERR_SPACE = -3,     /* workspace too small */
    BL_ERR_UNSUPPORTED = -4
} bl_status;

/* per-proposal status bits written by the device loop (bl_get_status) */
#define BL_ST_CASCADE_GUARD 1u   /* pit cascade hit the iteration guard (reference would spin) */
#define BL_ST_DRAIN_CYCLE   4u   /* pit-drainage graph had a cycle: sequential fallback was used */
#define BL_ST_NAN           8u   /* non-finite elevation / time step produced (set at every checkpoint and at the end of a run) */

typedef struct {
    int32_t N;                 /* TIN nodes */
    int32_t bounds;            /* ghost nodes are ids 0..bounds-1 (recGrid.boundsPt) */
    const int32_t *ngb;        /* [N][BL_KP] FVmesh.neighbours (negative = end) */
    const double *edge;        /* [N][BL_KP] FVmesh.edge_length */
    const double *vor;         /* [N][BL_KP] FVmesh.vor_edges */
    const double *area;        /* [N] FVmesh.control_volumes */
    const double *xy;          /* [N][2] FVmesh.node_coords */
    const int32_t *borders;    /* [N] flow.borders  (simulation/buildFlux.py:139-141) */
    const int32_t *borders2;   /* [N] flow.borders2 (buildFlux.py:149-151) */
    const int32_t *indomain;   /* [N] flow.domain.contains_points (buildFlux.py:135-138) */
    const int32_t *parent;     /* [bounds] flow.parentIDs */
    const double *z0;          /* [N] initial elevation (model.elevation after load_xml) */
    const double *disp;        /* [N] vertical displacement rate m/a, or NULL (model.disp) */
    /* TIN -> regular grid table of bl_mcmc.interpolateArray (k = 3 nearest, weights 1/d) */
    int32_t G;                 /* ny*nx output cells (0: no grid outputs) */
    const int32_t *grid_idx;   /* [G][3] */
    const double *grid_w;      /* [G][3] */
    const int32_t *grid_exact; /* [G] 1 where the nearest node is at distance 0 */
} bl_mesh;

typedef struct {
    double spl_m, spl_n;       /* FLOWalgo.eroparams (m, n) */
    double eps, fill_th;       /* PDalgo.pitparams (epsilon, fillTH) */
    double CDa, CDm, CDr;      /* creep coefficients */
    double minDT, maxDT;       /* input.minDT / maxDT */
    double hill_cfl, ms_cfl;   /* hillslope.CFL / CFLms (diffLinear.py:68, :142-144) */
    double diffprop;           /* PDalgo.pitparams pyProp */
    int32_t diffnb;            /* PDalgo.pitparams pyDiff */
    int32_t depo;              /* flow.depo */
    int32_t btype;             /* 0 fixed, 1 slope, 2 flat, 3 wall (buildFlux.py:292-302) */
    int32_t shuffle_mode;      /* 0: bases ascending; 1: counter-hash shuffle (flowNetwork.py:316) */
    double seapos;             /* force.sea0 */
    int32_t nsea;              /* sea-level curve length (0: constant seapos) */
    const double *seat;        /* [nsea] host */
    const double *seav;        /* [nsea] host */
    /* likelihood inputs (bl_mcmc.likelihoodFunc) */
    const double *real_elev;   /* [G] observed final topography, host, or NULL */
    int32_t npts;              /* erdp points (<= BL_NPTS_MAX) */
    const int32_t *pts_cell;   /* [npts] flat grid cell index row*nx+col (bl_mcmc.py:156-157) */
    /* engine options (all 0 = defaults).  Nothing in the library reads environment variables. */
    int32_t opt_tpb;           /* threads per CTA: 0 automatic, or 256 / 512 / 768 / 1024 */
    int32_t opt_flags;         /* BL_OPT_* bits */
    int32_t opt_grid_group;    /* lattice engine: proposals per fill group (0 = 64) */
    int32_t opt_debug;         /* timing experiments only */
} bl_params;

#define BL_OPT_FORCE_GENERIC 1   /* use the L2-resident device path even if the mesh fits shared memory */
#define BL_OPT_NO_L2_PERSIST 2   /* do not pin the mesh constants in L2 */
#define BL_OPT_COMPACT       8   /* use the 8-bytes-per-node shared-memory layout wherever it fits (it is chosen automatically when it lets two landscapes share an SM) */
#define BL_OPT_NO_COMPACT   16   /* never use it */

typedef struct bl_ctx bl_ctx;

const char *bl_last_error(void);
int bl_version(void);

/* Bytes of device workspace bl_ctx_create needs for B proposals on this mesh. */
size_t bl_workspace_bytes(const bl_mesh *mesh, const bl_params *par, int B);

/* Build a context: validates the mesh (symmetric adjacency, ghost ids), derives the Gauss-Seidel level
 * schedule of fillPD, uploads constants into `workspace` (a DEVICE pointer the caller owns). */
int bl_ctx_create(const bl_mesh *mesh, const bl_params *par, int B, int device, void *workspace,
                  size_t workspace_bytes, bl_ctx **out);
int bl_ctx_destroy(bl_ctx *ctx);

/* Model() + load_xml() + blackBox parameter pokes for all B proposals: elevation <- z0, cumdiff <- 0,
 * tNow <- tStart.  rain/erod/seed are DEVICE pointers [B] (bl_mcmc.py:132-136; seeds replace
 * model.py:68-74). */
int bl_reset(bl_ctx *ctx, const double *rain_dev, const double *erod_dev, const uint64_t *seed_dev,
             double tStart, void *stream);

/* Model.run_to_time for every proposal: step each proposal through the stop times (HOST array, ascending).
 * ckpt[i] >= 0 marks stops[i] as checkpoint number ckpt[i] (the 5 sim_interval times of blackBox): the
 * TIN->grid interpolation of elevation / cumdiff is evaluated there.  Outputs are DEVICE pointers, any may
 * be NULL:
 *   out_pts  [B][BL_NCKPT_MAX][BL_NPTS_MAX]  cumdiff grid value at the erdp points, per checkpoint
 *   out_sse  [B]                             sum (pred_elev - real_elev)^2 at the LAST checkpoint marked
 *   out_elev [B][nck][G], out_erdp [B][nck][G]  full grids (nck = number of marked stops), optional  */
int bl_run(bl_ctx *ctx, int nstops, const double *stops, const int32_t *ckpt, double *out_pts,
           double *out_sse, double *out_elev, double *out_erdp, int nck, void *stream);

/* One time step for every proposal (buildFlux.streamflow + sediment_flux with this tStop); leaves every
 * per-stage product in the workspace for bl_get_field - the per-stage parity entry point. */
int bl_step(bl_ctx *ctx, double tStop, void *stream);

typedef enum {
    BL_F_Z = 0, BL_F_CUMDIFF, BL_F_CUMHILL, BL_F_FILLH, BL_F_MAXH, BL_F_Q, BL_F_PITVOL, BL_F_ERO,
    BL_F_DEPO, BL_F_SEDFLUX,                                    /* double [N] */
    BL_F_RCV = 32, BL_F_RCV1, BL_F_STACK, BL_F_STACK1, BL_F_PITID, BL_F_PITDRAIN, BL_F_ALLDRAIN /* int32 [N] */
} bl_field;

/* Copy one per-proposal array to HOST memory (synchronises the stream). */
int bl_get_field(bl_ctx *ctx, int field, int proposal, void *dst_host, void *stream);
/* tNow[B], steps[B] (int64), status[B] (uint32), last dt[B] to HOST; any pointer may be NULL. */
int bl_get_scalars(bl_ctx *ctx, double *tnow, int64_t *steps, uint32_t *status, double *last_dt,
                   void *stream);
/* diagnostics: SM clock cycles spent per phase of the time step, [B][BL_NPHASE] int64 to HOST.
 * Phase order: fill, sfd, stack1, basins, stack, drainage, discharge, cfl, streampower, newdt(+rerun),
 * erodep, apply; slots 12-15 split streampower on the shared-memory path (event compaction, parallel replay,
 * serial replay, pre-pass + routing); slots 16-19 split erodep (first pass + land-pit compaction, small basins,
 * large basins, tail), 20 / 21 count large basins / land pits; 22-31 split discharge, the stream-power pre-pass and the
 * stack ordering of the compact layout; 32-37 are fill diagnostics of the compact layout (opt_debug bit 2: rounds, evaluations,
 * cycles of warp 0 in scan / evaluation / barrier); the rest are reserved. */
#define BL_NPHASE 48
int bl_get_phase_cycles(bl_ctx *ctx, int64_t *dst_host, void *stream);
/* number of kernel launches issued by this context so far (for bench.py's gpu_launches) */
int64_t bl_launch_count(const bl_ctx *ctx);
/* launch configuration the context selected: info[0] threads per CTA, info[1] device path (0 generic / L2-resident,
 * 1 shared memory 17 B per node, 2 compact shared memory 8 B per node), info[2] dynamic shared memory per CTA in bytes,
 * info[3] CTAs of this kernel resident per SM (occupancy query). */
int bl_ctx_info(const bl_ctx *ctx, int32_t info[4]);

/* ---- per-stage entry points on caller-owned DEVICE arrays ------------------------------------------
 * One routine of the time step per call, in the argument order of the f2py signature it replaces, executed by the SAME
 * device function the fused step (bl_run) uses - the shared-memory path or the generic path, whichever the context
 * selected.  Every array is a device pointer of length N (mesh constants come from the context, as the reference keeps
 * them in module state: libUtils/PDclass.f90:15-34).  Proposal slot 0 of the context is the scratch landscape: its
 * state is overwritten (use a context of its own, or bl_reset afterwards).  Calls are asynchronous on `stream`.
 *
 *   reference call site (flow/flowNetwork.py, surface/elevationTIN.py)                          entry point
 *   fillH = pdstack.pitfilling(elev, allFill = 0, sealevel)          elevationTIN.py:304         bl_fill
 *   base, rcv, maxh, maxdep = sfd.directions(fillH, elev, ngb, edges, dist, gids)  :310         bl_sfd(fillH, elev, ...)
 *   base1, rcv1 = sfd.directions_base(elev, ngb, edges, dist, gids)  flowNetwork.py:298         bl_sfd(NULL, elev, ...)
 *   donors, stack = fstack.build(base, receivers, delta)             :393 / :412                bl_stack
 *   Q, lay = flowcompute.discharge(stack, receivers, elev, area * rain)  :440                   bl_discharge
 *   pitID, pitVolume = flowcompute.basinparameters(stack1, rcv1, elev, fillH, area)  :549       bl_basins
 *   pitDrain = flowcompute.basindrainage(order, pitID, rcv, pIDs, fillH, sea)  :566             bl_drainage
 *   allDrain = flowcompute.basindrainageall(order, pitID, rcv, pIDs)  :568                      bl_drainage
 *   cdepo, cero, sedload = flowcompute.streampower(stack, rcv, pitID, pitVolume, pitDrain, xy, area, maxh, maxdep,
 *                          Q, fillH, elev, rivqs, K, actlay, perc_dep, slp_cr, sea, dt, borders)  :656   bl_streampower
 *   diff = sfd.diffusion(elev, borders2, ngb, edges, dist, gids)     :148                        bl_diffuse
 *   zreg = interpolateArray(coords, z)                               bl_mcmc.py:167-215          bl_interp
 */
int bl_fill(bl_ctx *ctx, const double *elev, double sealevel, double *fillH, void *stream);
/* fillH != NULL: receivers on fillH, maxh / maxdep from elev (sfd.c:55-111); fillH == NULL: receivers on elev (sfd.c:113-153).
 * base[i] = i for self-receivers, -1 elsewhere; base / maxh / maxdep may be NULL.  (A non-NULL base synchronises.) */
int bl_sfd(bl_ctx *ctx, const double *fillH, const double *elev, int32_t *base, int32_t *rcv, double *maxh, double *maxdep,
           void *stream);
/* Pre-order DFS stack of the receiver forest (FLWnetwork.f90:21-71, FLOWstack.f90:24-40).  The base order is the
 * reference's: ascending ids (shuffle = 0, flowNetwork.py:302-303) or shuffled (shuffle = 1, :316, counter-hash of
 * (seed, step): DESIGN.md 5).  base_out [N] receives the ordered bases, nbase_out [1] their number; both may be NULL. */
int bl_stack(bl_ctx *ctx, const int32_t *rcv, int shuffle, uint64_t seed, int64_t step, int32_t *stack, int32_t *base_out,
             int32_t *nbase_out, void *stream);
/* Q = area * rain (0 on ghost nodes) accumulated down the receiver tree in reverse stack order (FLOWalgo.f90:53-81). */
int bl_discharge(bl_ctx *ctx, const int32_t *rcv, double rain, double *Q, void *stream);
int bl_basins(bl_ctx *ctx, const int32_t *rcv1, const double *elev, const double *fillH, int32_t *pitID, double *pitVol,
              void *stream);
/* Both drainage maps of compute_parameters_depression (flowNetwork.py:558-577); the pit order is derived inside. */
int bl_drainage(bl_ctx *ctx, const int32_t *pitID, const double *pitVol, const int32_t *rcv, const double *fillH,
                double sealevel, int32_t *pitDrain, int32_t *allDrain, void *stream);
/* Single rock, detachment limited (the configuration every BASELINE config reaches).  The stack the cascade follows
 * is built from rcv with the base order (shuffle, seed, step) as in bl_stack.  Outputs: cdepo, cero in m^3. */
int bl_streampower(bl_ctx *ctx, const int32_t *rcv, const int32_t *pitID, const double *pitVol, const int32_t *pitDrain,
                   const double *maxh, const double *Q, const double *fillH, const double *elev, double erod,
                   double sealevel, double dt, int shuffle, uint64_t seed, int64_t step, double *depo, double *ero,
                   void *stream);
int bl_diffuse(bl_ctx *ctx, const double *elev, double *diff, void *stream);
/* v [N] nodal values -> grid [G] on the context's k = 3 table. */
int bl_interp(bl_ctx *ctx, const double *v, double *grid, void *stream);

/* ---- lattice engine: regular D8 lattice too large for one CTA (BASELINE config 5) -------------
 * Same reference path as above for the purely erosive configuration (flow.depo = 0, constant sea level): every
 * phase of the time step is a grid-wide kernel batched over all proposals.  Nodes are the (ny+2) x (nx+2) lattice
 * including the ghost ring, stored row-major (x fastest); mesh ids are those of mesh.build_regular(order="colour")
 * (ghost ring in raster2TIN order, interior nodes by 2x2 parity class), which define fillPD's sweep order
 * (PDalgo.f90:49-81) and the donor summation order of the discharge (FLOWalgo.f90:53-81). */
typedef struct {
    int32_t nx, ny;            /* interior lattice size, both even */
    double dist[8];            /* |xy(d) - xy(r)| per D8 slot E,NE,N,NW,W,SW,S,SE (FLOWalgo.f90:299-330, :640-650) */
    double vor[8], edge[8];    /* FVmesh.vor_edges / edge_length per slot for interior nodes */
    double area;               /* FVmesh.control_volumes of interior nodes */
    const double *z0;          /* [(ny+2)*(nx+2)] initial elevation, lattice order, HOST */
    const uint8_t *flags;      /* [(ny+2)*(nx+2)] bit0 flow.borders, bit1 flow.borders2 (buildFlux.py:139-151) */
    const double *disp;        /* lattice order displacement rate (m/a) or NULL */
    const int32_t *parent;     /* [2*(nx+2)+2*ny] lattice cell of each ghost node's parent (flow.parentIDs) or NULL */
    const double *real_elev;   /* lattice order observed topography or NULL */
} bl_lattice;

typedef struct bl_grid bl_grid;

size_t bl_grid_workspace_bytes(const bl_lattice *lat, int B, int debug);
/* debug != 0 keeps a copy of fillH per step for bl_grid_get_field (8 more bytes per node per proposal). */
int bl_grid_create(const bl_lattice *lat, const bl_params *par, int B, int device, void *workspace,
                   size_t workspace_bytes, int debug, bl_grid **out);
int bl_grid_destroy(bl_grid *ctx);
int bl_grid_reset(bl_grid *ctx, const double *rain_dev, const double *erod_dev, double tStart, void *stream);
/* as bl_run; grids are [B][nck][(ny+2)*(nx+2)] in lattice order */
int bl_grid_run(bl_grid *ctx, int nstops, const double *stops, const int32_t *ckpt, double *out_pts,
                double *out_sse, double *out_elev, double *out_erdp, int nck, void *stream);
int bl_grid_step(bl_grid *ctx, double tStop, void *stream);
/* BL_F_Z / CUMDIFF / CUMHILL / Q / FILLH (debug contexts) as double, BL_F_RCV as int32 lattice cell, lattice order */
int bl_grid_get_field(bl_grid *ctx, int field, int proposal, void *dst_host, void *stream);
int bl_grid_get_scalars(bl_grid *ctx, double *tnow, int64_t *steps, int64_t *fill_sweeps, double *last_dt,
                        void *stream);
int64_t bl_grid_launch_count(const bl_grid *ctx);

#ifdef __cplusplus
}
#endif
#endif

/* bl_oracle.h - CPU ORACLE for the Bayeslands hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's algorithm for the path named in BASELINE.json:north_star
 * (pyBadlands forward model + likelihood).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; the product path (bayeslands_b200/) never does.
 *
 * PARITY STATUS: the reference ships no unit tests and cannot be executed here (Python 2 + gfortran +
 * Triangle).  Pinned against the reference itself: sfd.c routines (compiled from
 * /root/reference into oracle/_ref, bit-exact; golden vectors in tests/golden) and the shipped end-to-end data
 * (likelihood formula, likelihood-surface argmax, MSE at truth, per-cell deviation of the shipped sweeps).
 * Pinned by hand-derived fixtures (tests/micro.py, derivations in tests/golden/README.md): fillPD incl. the capped
 * early stop and its order dependence, fstack.build, basinparameters, basindrainage(all) incl. a de-dup hit, discharge,
 * streampower incl. the two-pit overflow cascade and its marine branch (FLOWalgo.f90:979-991), the land-pit proportional fill,
 * the CFL time step incl. n != 1, getid1's newdt re-run, marine_distribution (every exit of the walk), diffmarine and its
 * sub-step loop.  No routine of the path is left "parity unpinned"; the fixtures are 1-D chains, so neighbour-order effects on
 * a real TIN are pinned only where the reference's own sfd.c covers them (receivers, hillslope diffusion).
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 */
#ifndef BL_ORACLE_H
#define BL_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OR_KP 20 /* MAX_NEIGHBOURS, pyBadlands/libUtils/sfd.c:14 */

typedef struct {
    int N;               /* TIN nodes */
    int bounds;          /* ghost nodes = ids 0..bounds-1 (recGrid.boundsPt) */
    const int *ngb;      /* [N][20] neighbours, negative terminates */
    const double *edge;  /* [N][20] Delaunay edge length  (FVmesh.edge_length) */
    const double *vor;   /* [N][20] Voronoi edge length   (FVmesh.vor_edges) */
    const double *area;  /* [N]     Voronoi cell area     (FVmesh.control_volumes) */
    const double *xy;    /* [N][2] */
    const int *borders;  /* [N] 1 on insideIDs   (buildFlux.py:139-141) */
    const int *borders2; /* [N] 1 on insideIDs2  (buildFlux.py:149-151) */
    const int *indomain; /* [N] 1 inside the [min-1,max+1] box (flow.domain, buildFlux.py:135-138) */
    const int *parent;   /* [bounds] parentIDs */
} or_mesh;

typedef struct {
    double spl_m, spl_n;        /* FLOWalgo.f90:41-42 */
    double eps, fill_th;        /* PDalgo.f90:288-289 */
    double CDa, CDm, CDr;       /* diffLinear */
    double minDT, maxDT;        /* xmlParser.py:328-338 */
    double hill_cfl, ms_cfl;    /* diffLinear.py:68, :142-144 */
    double diffprop;            /* PDalgo.f90:285 */
    int diffnb;                 /* PDalgo.f90:284 */
    int depo;                   /* flow.depo */
    int btype;                  /* 0 fixed, 1 slope, 2 flat, 3 wall */
    int pow_mode;               /* 0: libm pow (reference), 1: sqrt/identity fast paths for m=0.5, n=1 */
    int shuffle_mode;           /* 0: bases ascending, 1: hash-key shuffle (see or_shuffle_base) */
    int fill_mode;              /* 0: literal fillPD (sweep order + early stop), 1: run to the fixed point */
    uint64_t seed;
    double seapos;              /* forceSim.sea0 */
    int nsea;                   /* 0: constant sea level */
    const double *seat, *seav;  /* sea-level curve (forceSim.py:212-224) */
    double tStart, tEnd;
    int apply_disp;             /* vertical displacement rate active over the whole run */
    const double *disp;         /* [N] m/a */
} or_params;

/* per-proposal model state (Model: pyBadlands/model.py:18-44; flowNetwork: flow/flowNetwork.py:40-115) */
typedef struct {
    int N;
    double tNow;
    double rain, erod;
    long step;              /* time steps taken */
    long fill_sweeps;       /* diagnostic */
    long rerun;             /* streampower re-runs (newdt < dt) */
    double *z, *cumdiff, *cumhill;
    /* per-step products */
    double *fillH, *maxh, *maxdep, *Q, *lay, *pitVol;
    int *rcv, *rcv1, *stack, *stack1, *donors, *donors1, *pitID, *pitDrain, *allDrain;
    int nbase, nbase1;
    int *base, *base1;
    double *ero, *depo, *sedflux; /* cero, cdepo (m3), ed (m) */
    double last_dt, last_cfl;
    /* scratch */
    double *w0, *w1, *w2, *w3, *w4, *w5;
    int *i0, *i1, *i2, *i3, *i4;
} or_state;

/* ---- native kernels, one per reference routine ---- */
void or_directions_base(const or_mesh *m, const double *z, int *base, int *rcv);
void or_directions(const or_mesh *m, const double *fillH, const double *z, int *base, int *rcv,
                   double *maxh, double *maxdep);
void or_diffusion(const or_mesh *m, const double *z, double *diff);
long or_pitfilling(const or_mesh *m, const or_params *p, const double *z, double sea, double *fillH,
                   int *work1, int *work2);
void or_shuffle_base(int *base, int nb, uint64_t seed, uint64_t step, uint64_t *keys);
void or_stack_build(int N, const int *base, int nb, const int *rcv, int *delta, int *donors,
                    int *stack, int *tmp);
void or_discharge(int n, const int *stack, const int *rcv, const double *z, const double *q0, int N,
                  double *Q, double *lay);
void or_basinparameters(int n, const int *stack1, const int *rcv1, const double *z, const double *fillH,
                        const double *area, int N, int *pitID, double *pitVol);
void or_basindrainage(int np, const int *orderPits, const int *pitID, const int *rcv, const int *pIDs,
                      const double *fillH, double sea, int N, int *drain, int *chain);
void or_basindrainageall(int np, const int *orderPits, const int *pitID, const int *rcv, const int *pIDs,
                         int N, int *drain, int *chain);
double or_flowcfl(const or_mesh *m, const or_params *p, const int *rcv, const double *elev,
                  const double *Q, double erod);
int or_streampower(const or_mesh *m, const or_params *p, int n, const int *stack, const int *rcv,
                   const int *pitID, const double *pitVol1, const int *pitDrain, const double *maxh,
                   const double *maxdep, const double *Q, const double *fillH, const double *z, double erod,
                   double sea, double dt, double *depo, double *ero, double *sedfl, double *pitVol);
void or_diffmarine(const or_mesh *m, const double *z, const double *depoH, const double *coeff, double sea,
                   double maxth, double tstep, double *diff, double *mindt);
void or_marine_distribution(const or_mesh *m, const or_params *p, const double *elevation, const double *seavol,
                            double sea, const int *depIDs, int nIDs, double *diffsed, double *elev,
                            double *seadep);
double or_pairwise_sum(const double *a, long n);

/* ---- orchestration ---- */
or_state *or_state_create(int N);
void or_state_destroy(or_state *s);
void or_state_reset(or_state *s, const double *z0, double rain, double erod, double tStart);
double or_sea(const or_params *p, double t);
void or_streamflow(const or_mesh *m, const or_params *p, or_state *s);               /* buildFlux.py:21-100 */
int or_sediment_flux(const or_mesh *m, const or_params *p, or_state *s, double tStop); /* buildFlux.py:102-320 */
int or_run_to(const or_mesh *m, const or_params *p, or_state *s, int nstops, const double *stops);

/* bl_mcmc.interpolateArray (bl_mcmc.py:198-213) given the mesh-constant k=3 table */
void or_interp(int G, const int *idx, const double *w, const int *exact, const double *v, double *out);
/* sum of squared differences (bl_mcmc.py:488) with numpy's pairwise summation order */
double or_sse(const double *a, const double *b, long n, double *tmp);

#ifdef __cplusplus
}
#endif
#endif

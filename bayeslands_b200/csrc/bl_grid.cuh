// bl_grid.cuh - lattice engine: the hot path on a regular D8 lattice too large for one CTA (BASELINE config 5:
// 2048 x 2048 nodes x 256 proposals), included at the end of bl_engine.cu.
//
// Where bl_run_kernel gives one CTA a whole landscape, here every phase of the time step is a grid-wide kernel
// batched over all proposals (blockIdx.z / task index = proposal), state is structure-of-arrays [B][LY][LXP] in HBM
// and the streaming phases are bandwidth bound.  Scope: the purely erosive configuration (flow.depo = 0, constant
// sea level): with no deposition the pit cascade, basin volumes and both stack orders cannot influence the state
// (FLOWalgo.f90:875-892 writes erosion from node-local values only; getid1's overfill test FLOWalgo.f90:1047-1086
// needs eroded volume > pit volume > 0, impossible for erosion <= 0), so the step reduces to
//   fill (PDalgo.f90:19-91) -> receivers on the filled surface (sfd.c:55-111) -> discharge (FLOWalgo.f90:53-81)
//   -> flowcfl (FLOWalgo.f90:299-330) + dt logic (buildFlux.py:154-177, flowNetwork.py:687-698)
//   -> detachment-limited erosion (FLOWalgo.f90:635-714, :875-892) -> hillslope diffusion (sfd.c:155-185,
//   diffLinear.py:150-182, buildFlux.py:252-272) -> boundary rule + uplift (buildFlux.py:292-313).
// Order-dependent semantics kept bit-exact:
//   * fillPD's ascending-id Gauss-Seidel sweep: mesh.build_regular(order="colour") numbers interior nodes by the
//     2x2 parity class, so a sweep is 4 race-free colour passes; pending nodes live in a two-level bitmap per
//     (proposal, colour) and a pass touches only the front; the reference's early stop (`change`) is per proposal;
//   * discharge: Q[r] = (q0[r] + Q[d1]) + Q[d2] ... with donors in DESCENDING mesh id (what the reverse walk of the
//     pre-order stack does) - evaluated by a wait-free traversal: the thread that completes the last donor of r
//     (atomic arrival counter) sums r and walks on downstream; nobody ever spins;
//   * SFD ties: first lowest neighbour in slot order E,NE,N,NW,W,SW,S,SE (the builder's neighbour order).

#include <cooperative_groups.h>

namespace grid {

constexpr int CHS = 32;   // ints between the `change` flags of two proposals (one line each: they are stored to by every warp)

struct GConst {
    int nx, ny, LX, LY, LXP, B, hx, ncol, W0, W1, nb, btype, npts, has_parent;
    size_t NLP;                 // LY*LXP rounded up: length of one per-proposal array
    double m, n, eps, TH, CDa, CDm, minDT, maxDT, hill, sea, area, boff;
    double dist[8], vor[8], edge[8];
    const unsigned char *flags; // [NLP] bit0 borders, bit1 borders2
    const double *z0, *disp, *real;
    const int *parentL, *ghostL; // [nb] lattice index of ghost g / of its parent
    const int *pts_cell;
    double *z, *W, *Q, *cumdiff, *cumhill, *fdbg; // [B][NLP]
    unsigned char *rcv, *don;   // [B][NLP] receiver slot (8 = self), donor mask
    unsigned *cnt;              // [B][NLP/4] arrival counters, one byte per node
    unsigned *l0;               // [B][4][W0p] pending bitmaps
    unsigned *wl;               // [4][wlcap] rings of non-empty bitmap words (entry = p*W0 + word)
    void *wlctl;                // [4] ring head / tail / snapshot
    unsigned long long wlcap;   // power of two >= B*W0
    int W0p, W1p;
    double *rain, *erod, *tnow, *dt, *lastdt, *partial;
    unsigned long long *cflbits;
    long long *steps, *sweeps;
    int *running, *fill_active, *change, *si, *ckflag;
    int *counters;              // [0] running proposals, [1] proposals still sweeping, [2] sweeps this step
    const double *stops;
    const int *ckpt;
    int nstops, npart;
};

__constant__ int c_di[8] = {1, 1, 0, -1, -1, -1, 0, 1};
__constant__ int c_dj[8] = {0, 1, 1, 1, 0, -1, -1, -1};

__device__ __forceinline__ double powm(double x, double e)
{
    if (e == 0.5) return sqrt(x);
    if (e == 1.0) return x;
    if (e == 0.0) return 1.0;
    return pow(x, e);
}
__device__ __forceinline__ double floorish(double x) { return (x > 1.) ? round(x - 0.5) : x; }
__device__ __forceinline__ bool is_ghost(const GConst &g, int I, int J) { return I == 0 || J == 0 || I == g.LX - 1 || J == g.LY - 1; }

// mesh id of lattice node (I, J): ghost ring in raster2TIN order, interior colour-major (mesh.build_regular)
__device__ __forceinline__ int mesh_id(const GConst &g, int I, int J)
{
    if (J == 0) return I;
    if (J == g.LY - 1) return g.LX + g.ny + I;
    if (I == 0) return g.LX + (J - 1);
    if (I == g.LX - 1) return 2 * g.LX + g.ny + (J - 1);
    const int i = I - 1, j = J - 1;
    return g.nb + ((j & 1) * 2 + (i & 1)) * g.ncol + (j >> 1) * g.hx + (i >> 1);
}

__global__ void k_reset(GConst g, const double *rain, const double *erod, double tStart)
{
    const int p = blockIdx.y;
    const size_t L = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (L < g.NLP) {
        const size_t o = (size_t)p * g.NLP + L;
        g.z[o] = g.z0[L];
        g.cumdiff[o] = 0.;
        g.cumhill[o] = 0.;
        g.W[o] = 0.;
        g.Q[o] = 0.;
        g.rcv[o] = 8;
        g.don[o] = 0;
        if (g.fdbg) g.fdbg[o] = 0.;
    }
    if (L == 0) {
        g.rain[p] = rain[p]; g.erod[p] = erod[p]; g.tnow[p] = tStart; g.dt[p] = 0.; g.lastdt[p] = 0.;
        g.steps[p] = 0; g.sweeps[p] = 0; g.running[p] = 0; g.fill_active[p] = 0; g.change[p * CHS] = 0; g.si[p] = 0; g.ckflag[p] = 0;
    }
}

// Model.run_to_time bookkeeping (model.py:266-490): advance the clock, pass stop times, flag checkpoints
__global__ void k_advance(GConst g, int do_step, int restart)
{
    __shared__ int s_run;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (int p = threadIdx.x; p < g.B; p += blockDim.x) {
        if (restart) { g.running[p] = 1; g.si[p] = 0; }
        int ck = 0;
        if (g.running[p]) {
            if (do_step) { g.tnow[p] += g.dt[p]; g.lastdt[p] = g.dt[p]; g.steps[p] += 1; }
            int si = g.si[p];
            while (si < g.nstops && !(g.tnow[p] < g.stops[si])) {
                if (g.ckpt[si] >= 0) ck |= 1 << g.ckpt[si];
                ++si;
            }
            g.si[p] = si;
            if (si >= g.nstops) g.running[p] = 0; else atomicAdd(&s_run, 1);
        }
        g.ckflag[p] = ck;
    }
    __syncthreads();
    if (threadIdx.x == 0) g.counters[0] = s_run;
}

// ---- depression filling ----------------------------------------------------------------------------------
// Pending nodes: one bit per node in a per-(proposal, colour) bitmap, and per colour a ring of the bitmap WORDS that
// hold at least one bit (a word enters the ring when its first bit is set, leaves it when the colour's pass consumes
// it), so a pass is one warp per non-empty word - no scanning, work proportional to the front.
struct WlCtl { unsigned long long head, tail, snap, pad[13]; };   // one 128-byte line per colour: tail is an atomic hot spot

// proposals [p0, p0 + gridDim.y) are filled together (the host walks over groups: see grid_one_step)
__global__ void k_fill_init(GConst g, int p0)
{
    const int p = p0 + blockIdx.y;
    const size_t L = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (L < (size_t)4 * g.W0) {                                      // every word of every colour starts in its ring
        const int c = (int)(L / g.W0), w = (int)(L % g.W0);
        unsigned m = 0xFFFFFFFFu;
        if (w * 32 + 32 > g.ncol) m = (1u << (g.ncol - w * 32)) - 1u;
        g.l0[((size_t)p * 4 + c) * g.W0p + w] = m;
        g.wl[(size_t)c * g.wlcap + (size_t)blockIdx.y * g.W0 + w] = (unsigned)((size_t)p * g.W0 + w);
    }
    if (blockIdx.y == 0 && L < 4) {
        WlCtl *ctl = (WlCtl *)g.wlctl + L;
        ctl->head = 0; ctl->snap = 0; ctl->tail = (unsigned long long)gridDim.y * g.W0;
    }
    if (!g.running[p]) { if (L == 0) g.fill_active[p] = 0; return; }
    if (L < (size_t)g.LY * g.LXP) {
        const int I = (int)(L % g.LXP), J = (int)(L / g.LXP);
        if (I < g.LX) {
            const size_t o = (size_t)p * g.NLP + L;
            g.W[o] = is_ghost(g, I, J) ? g.z[o] : 1.e6;                    // PDalgo.f90:19-35
        }
    }
    if (L == 0) {
        g.fill_active[p] = 1; g.change[p * CHS] = 0;
        g.cflbits[p] = (unsigned long long)__double_as_longlong(1.e6);
    }
}

// one colour pass of the Gauss-Seidel sweep (PDalgo.f90:37-91): one warp per pending bitmap word.  Nodes of one
// colour are never adjacent, so the pass is race free and equals the sequential sweep over that id range.
// (W, the pending words, the rings and their control words are written by other SMs between passes of the persistent
// loop below: they are read through L2 - __ldcg - so that no stale L1 line can be seen.)
__device__ __forceinline__ void fill_pass_body(const GConst &g, int c)
{
    WlCtl *ctl = (WlCtl *)g.wlctl;
    const unsigned long long h = __ldcg(&ctl[c].head), t = __ldcg(&ctl[c].tail);   // nobody appends to colour c during its own pass
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        ctl[c].snap = t;
        const int cp = (c + 3) & 3;
        ctl[cp].head = __ldcg(&ctl[cp].snap);                       // retire what the previous pass consumed
    }
    const int lane = threadIdx.x & 31;
    const unsigned long long warp = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    const int hx = g.hx, LXP = g.LXP;
    const double eps = g.eps, TH = g.TH, sea = g.sea;
    const unsigned *ring = g.wl + (size_t)c * g.wlcap;
    // ring appends are staged per block: one global tail reservation per colour per block instead of one per word
    // (same-address atomics retire at ~0.7 ns each in L2 - a quarter of a pass at 32 proposals)
    constexpr unsigned SCAP = 256;
    __shared__ unsigned s_buf[4][SCAP];
    __shared__ unsigned s_cnt[4];
    __shared__ unsigned long long s_base[4];
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0u;
    __syncthreads();
    auto push = [&](int c2, unsigned ent) {
        const unsigned pos = atomicAdd(&s_cnt[c2], 1u);
        if (pos < SCAP) s_buf[c2][pos] = ent;
        else {
            const unsigned long long gp = atomicAdd(&ctl[c2].tail, 1ull);
            g.wl[(size_t)c2 * g.wlcap + (gp & (g.wlcap - 1))] = ent;
        }
    };
    // software pipeline over the warp's tasks: the next ring entry is requested before this task's gathers and the next
    // pending word (read and cleared: nobody else touches colour c's words during its pass) before this task's wake-up
    // atomics, so a task costs three dependent memory round trips instead of five
    auto take_word = [&](unsigned ent, int &p, int &w) -> unsigned {
        p = (int)(ent / (unsigned)g.W0); w = (int)(ent % (unsigned)g.W0);
        unsigned word = 0u;
        if (__ldcg(g.fill_active + p)) {
            unsigned *L0 = g.l0 + ((size_t)p * 4 + c) * g.W0p + w;
            if (lane == 0) { word = __ldcg(L0); if (word) *L0 = 0u; }
            word = __shfl_sync(0xffffffffu, word, 0);
        }
        return word;
    };
    unsigned long long e = h + warp;
    int pN = 0, wN = 0;
    unsigned wordN = 0u;
    bool more = e < t;
    if (more) wordN = take_word(__ldcg(ring + (e & (g.wlcap - 1))), pN, wN);
    while (more) {
        const int p = pN, w = wN;
        const unsigned word = wordN;
        e += nwarps;
        more = e < t;
        const unsigned ent2 = more ? __ldcg(ring + (e & (g.wlcap - 1))) : 0u;
        int loc = 0, changed = 0;
        double av[8], nwv = 0.;                     // neighbour values in slot order E,NE,N,NW,W,SW,S,SE; the new W
#pragma unroll
        for (int o = 0; o < 8; ++o) av[o] = 0.;
        if ((word >> lane) & 1u) {
            double *W = g.W + (size_t)p * g.NLP;
            const double *z = g.z + (size_t)p * g.NLP;
            const int q = w * 32 + lane;
            const int qy = q / hx, qx = q - qy * hx;
            const int i = 2 * qx + (c & 1), j = 2 * qy + (c >> 1);
            const size_t L = (size_t)(j + 1) * LXP + (i + 1);
            // all ten loads go out together (a pending node almost always still has W > z)
            const double *r0 = W + L - LXP, *r2 = W + L + LXP;
            const double wk = __ldcg(W + L), zk = z[L];
            av[0] = __ldcg(W + L + 1); av[1] = __ldcg(r2 + 1); av[2] = __ldcg(r2); av[3] = __ldcg(r2 - 1);
            av[4] = __ldcg(W + L - 1); av[5] = __ldcg(r0 - 1); av[6] = __ldcg(r0); av[7] = __ldcg(r0 + 1);
            if (wk > zk) {
                double hmin = fmin(fmin(fmin(av[0], av[1]), fmin(av[2], av[3])), fmin(fmin(av[4], av[5]), fmin(av[6], av[7])));
                hmin = fmin(hmin, 2.e6);
                const double he = hmin + eps;
                double nw = wk;
                if (zk >= he) nw = zk;
                else if (wk > he) {
                    nw = he;
                    if (zk >= sea) { if (nw - zk > TH) nw = zk + TH; else loc = 1; }
                    else           { if (nw - sea > TH) nw = sea + TH; else loc = 1; }
                }
                if (nw != wk) { W[L] = nw; changed = 1; nwv = nw; }
            }
        }
        if (more) wordN = take_word(ent2, pN, wN);
        // wake the neighbours a change can move (they are of the other three colours): re-evaluating j alters W(j) only
        // if min_nb W + eps < W(j), and the neighbour whose drop makes that true is the one that wakes j
        const double lim = nwv + eps;
        const unsigned M = __ballot_sync(0xffffffffu, changed);
        if (M) {
            if ((hx & 31) == 0) {
                // the word's 32 nodes share a lattice row and neighbour slot o of all of them lies in one or two
                // words of one colour: lanes 0..7 each post one slot with at most two atomics
                unsigned Ms = 0u;
#pragma unroll
                for (int o = 0; o < 8; ++o) {
                    const unsigned mo = __ballot_sync(0xffffffffu, changed && av[o] > lim);
                    if (lane == o) Ms = mo;
                }
                if (lane < 8) {
                    const int cx = c & 1, qx0 = (w * 32) % hx, jw = 2 * ((w * 32) / hx) + (c >> 1);
                    const int j2 = jw + c_dj[lane], tx = cx + c_di[lane];
                    if (j2 >= 0 && j2 < g.ny) {
                        const int sx = (tx + 2) / 2 - 1;                           // floor(tx / 2), tx in -1..2
                        const int c2 = (j2 & 1) * 2 + (tx & 1);
                        unsigned M2 = Ms;
                        if (sx < 0 && qx0 == 0) M2 &= ~1u;                         // i2 = -1
                        if (sx > 0 && qx0 + 32 == hx) M2 &= 0x7FFFFFFFu;           // i2 = nx
                        const int w2 = ((j2 >> 1) * hx + qx0) >> 5;
                        unsigned *B2 = g.l0 + ((size_t)p * 4 + c2) * g.W0p;
                        unsigned b0 = M2, b1 = 0u; int w1 = w2;
                        if (sx > 0) { b0 = M2 << 1; b1 = M2 >> 31; w1 = w2 + 1; }
                        else if (sx < 0) { b0 = M2 >> 1; b1 = (M2 & 1u) << 31; w1 = w2 - 1; }
                        // both bitmap atomics first, then one ring reservation for the words that were empty
                        const unsigned o0 = b0 ? atomicOr(B2 + w2, b0) : 1u;
                        const unsigned o1 = b1 ? atomicOr(B2 + w1, b1) : 1u;
                        const int n0 = (o0 == 0u), n1 = (o1 == 0u);
                        if (n0) push(c2, (unsigned)((size_t)p * g.W0 + w2));
                        if (n1) push(c2, (unsigned)((size_t)p * g.W0 + w1));
                    }
                }
            } else if (changed) {
                const int q = w * 32 + lane;
                const int qy = q / hx, qx = q - qy * hx;
                const int i = 2 * qx + (c & 1), j = 2 * qy + (c >> 1);
#pragma unroll
                for (int o = 0; o < 8; ++o) {
                    const int i2 = i + c_di[o], j2 = j + c_dj[o];
                    if (i2 < 0 || j2 < 0 || i2 >= g.nx || j2 >= g.ny || !(av[o] > lim)) continue;
                    const int c2 = (j2 & 1) * 2 + (i2 & 1);
                    const int q2 = (j2 >> 1) * hx + (i2 >> 1);
                    const unsigned old = atomicOr(g.l0 + ((size_t)p * 4 + c2) * g.W0p + (q2 >> 5), 1u << (q2 & 31));
                    if (old == 0u) push(c2, (unsigned)((size_t)p * g.W0 + (q2 >> 5)));
                }
            }
        }
        if (__any_sync(0xffffffffu, loc) && lane == 0 && __ldcg(g.change + p * CHS) == 0) g.change[p * CHS] = 1;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        const unsigned n = min(s_cnt[threadIdx.x], SCAP);
        s_base[threadIdx.x] = n ? atomicAdd(&ctl[threadIdx.x].tail, (unsigned long long)n) : 0ull;
    }
    __syncthreads();
#pragma unroll
    for (int c2 = 0; c2 < 4; ++c2) {
        const unsigned n = min(s_cnt[c2], SCAP);
        for (unsigned i = threadIdx.x; i < n; i += blockDim.x)
            g.wl[(size_t)c2 * g.wlcap + ((s_base[c2] + i) & (g.wlcap - 1))] = s_buf[c2][i];
    }
    __syncthreads();
}

// end of a sweep: the reference's `do while (change)` per proposal (one block)
__device__ __forceinline__ void fill_end_body(const GConst &g)
{
    __shared__ int s_act;
    if (threadIdx.x == 0) s_act = 0;
    __syncthreads();
    if (threadIdx.x == 0 && __ldcg(g.counters + 1) > 0) g.counters[2] += 1;      // sweeps this step (max over proposals)
    for (int p = threadIdx.x; p < g.B; p += blockDim.x) {
        if (__ldcg(g.fill_active + p)) {
            g.sweeps[p] += 1;
            if (__ldcg(g.change + p * CHS) == 0) g.fill_active[p] = 0; else atomicAdd(&s_act, 1);
            g.change[p * CHS] = 0;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) g.counters[1] = s_act;
}

// The whole `do while (change)` loop of a group of proposals in ONE cooperative launch: four colour passes and the
// end-of-sweep bookkeeping per sweep, separated by grid-wide barriers; the loop ends on the device when no proposal of
// the group raised `change` (no host round trip, no launch per pass).
__global__ void __launch_bounds__(256, 8) k_fill_loop(GConst g)
{
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    long long guard = 0;
    while (__ldcg(g.counters + 1) > 0) {
        for (int c = 0; c < 4; ++c) {
            fill_pass_body(g, c);
            grid.sync();
        }
        if (blockIdx.x == 0) fill_end_body(g);
        grid.sync();
        if (++guard > 50000000) break;
    }
}
__global__ void k_fill_begin(GConst g)
{
    __shared__ int s_act;
    if (threadIdx.x == 0) s_act = 0;
    __syncthreads();
    for (int p = threadIdx.x; p < g.B; p += blockDim.x) if (g.fill_active[p]) atomicAdd(&s_act, 1);
    __syncthreads();
    if (threadIdx.x == 0) { g.counters[1] = s_act; g.counters[2] = 0; }
}

// ---- receivers on the filled surface (sfd.c:55-111) ---------------------------------------------------------
// Each warp owns 32 columns and walks SROWS rows upwards with a 3 x 3 register window: one coalesced load per node, the
// left / right columns come from the neighbouring lanes (the two edge lanes fetch theirs).
constexpr int SROWS = 16;
template <typename T>
__device__ __forceinline__ void row3(const T *A, const GConst &g, int I, int J, int lane, T fill, T &l, T &c, T &r)
{
    const bool jok = (J >= 0 && J < g.LY);
    const T v = (jok && I < g.LX) ? A[(size_t)J * g.LXP + I] : fill;
    l = __shfl_up_sync(0xffffffffu, v, 1);
    r = __shfl_down_sync(0xffffffffu, v, 1);
    if (lane == 0) l = (jok && I >= 1 && I - 1 < g.LX) ? A[(size_t)J * g.LXP + I - 1] : fill;
    if (lane == 31) r = (jok && I + 1 < g.LX) ? A[(size_t)J * g.LXP + I + 1] : fill;
    c = v;
}
__global__ void __launch_bounds__(256) k_sfd(GConst g)
{
    const int p = blockIdx.z;
    if (!g.running[p]) return;
    const int lane = threadIdx.x, I = (blockIdx.x * 8 + threadIdx.y) * 32 + lane, J0 = blockIdx.y * SROWS;
    if (I - lane >= g.LX) return;                                   // the whole warp is outside the lattice
    const double *__restrict__ W = g.W + (size_t)p * g.NLP;
    unsigned char *__restrict__ rcvp = g.rcv + (size_t)p * g.NLP;
    double *__restrict__ fdbg = g.fdbg ? g.fdbg + (size_t)p * g.NLP : nullptr;
    const double INF = 1.e300;                                      // a missing neighbour is never lower
    double pl, pc, pr, cl, cc, cr, nl, nc, nr;
    row3(W, g, I, J0 - 1, lane, INF, pl, pc, pr);
    row3(W, g, I, J0, lane, INF, cl, cc, cr);
#pragma unroll 4
    for (int J = J0; J < J0 + SROWS; ++J) {
        row3(W, g, I, J + 1, lane, INF, nl, nc, nr);
        if (I < g.LX && J < g.LY) {
            const size_t o = (size_t)J * g.LXP + I;
            if (fdbg) fdbg[o] = cc;
            double wl = cc;
            int low = 8;                                            // slot order E,NE,N,NW,W,SW,S,SE, strict <, first wins
            if (cr < wl) { low = 0; wl = cr; }
            if (nr < wl) { low = 1; wl = nr; }
            if (nc < wl) { low = 2; wl = nc; }
            if (nl < wl) { low = 3; wl = nl; }
            if (cl < wl) { low = 4; wl = cl; }
            if (pl < wl) { low = 5; wl = pl; }
            if (pc < wl) { low = 6; wl = pc; }
            if (pr < wl) { low = 7; wl = pr; }
            rcvp[o] = (unsigned char)low;
        }
        pl = cl; pc = cc; pr = cr; cl = nl; cc = nc; cr = nr;
    }
}

// donor mask (which neighbours drain into me) + arrival counters
__global__ void __launch_bounds__(256) k_donors(GConst g)
{
    const int p = blockIdx.z;
    if (!g.running[p]) return;
    const int lane = threadIdx.x, I = (blockIdx.x * 8 + threadIdx.y) * 32 + lane, J0 = blockIdx.y * SROWS;
    if (I - lane >= g.LXP) return;
    const int NONE = 9;                                             // a missing neighbour drains nowhere
    const unsigned char *R = g.rcv + (size_t)p * g.NLP;
    int pl, pc, pr, cl, cc, cr, nl, nc, nr;
    auto row = [&](int J, int &l, int &c, int &r) {
        const bool jok = (J >= 0 && J < g.LY);
        const int v = (jok && I < g.LX) ? (int)R[(size_t)J * g.LXP + I] : NONE;
        l = __shfl_up_sync(0xffffffffu, v, 1);
        r = __shfl_down_sync(0xffffffffu, v, 1);
        if (lane == 0) l = (jok && I >= 1 && I - 1 < g.LX) ? (int)R[(size_t)J * g.LXP + I - 1] : NONE;
        if (lane == 31) r = (jok && I + 1 < g.LX) ? (int)R[(size_t)J * g.LXP + I + 1] : NONE;
        c = v;
    };
    row(J0 - 1, pl, pc, pr);
    row(J0, cl, cc, cr);
#pragma unroll 4
    for (int J = J0; J < J0 + SROWS; ++J) {
        row(J + 1, nl, nc, nr);
        if (I < g.LXP && J < g.LY) {
            const size_t o = (size_t)p * g.NLP + (size_t)J * g.LXP + I;
            if ((I & 3) == 0) g.cnt[o >> 2] = 0u;
            if (I < g.LX) {
                // neighbour in slot s drains into me when its receiver slot is the opposite one, (s + 4) & 7
                unsigned m = 0u;
                if (cr == 4) m |= 1u;        // E
                if (nr == 5) m |= 2u;        // NE
                if (nc == 6) m |= 4u;        // N
                if (nl == 7) m |= 8u;        // NW
                if (cl == 0) m |= 16u;       // W
                if (pl == 1) m |= 32u;       // SW
                if (pc == 2) m |= 64u;       // S
                if (pr == 3) m |= 128u;      // SE
                g.don[o] = (unsigned char)m;
            }
        }
        pl = cl; pc = cc; pr = cr; cl = nl; cc = nc; cr = nr;
    }
}

// discharge (FLOWalgo.f90:53-81; compute_flow flowNetwork.py:418-443): wait-free tree accumulation
__global__ void __launch_bounds__(256) k_discharge(GConst g)
{
    const int p = blockIdx.z;
    if (!g.running[p]) return;
    int I = blockIdx.x * 32 + threadIdx.x, J = blockIdx.y * 8 + threadIdx.y;
    if (I >= g.LX || J >= g.LY) return;
    const size_t base = (size_t)p * g.NLP;
    const unsigned char *rcv = g.rcv + base, *don = g.don + base;
    double *Q = g.Q + base;
    unsigned *cnt = g.cnt + (base >> 2);
    size_t L = (size_t)J * g.LXP + I;
    if (don[L] != 0) return;                                  // only leaves start a walk
    const double q0i = g.area * g.rain[p];                    // rain is 0 on ghost nodes (model.py:278-279)
    double acc = is_ghost(g, I, J) ? 0. : q0i;
    int s = rcv[L];
    for (;;) {
        Q[L] = acc;
        if (s == 8) break;
        const int Ir = I + c_di[s], Jr = J + c_dj[s];
        const size_t R = (size_t)Jr * g.LXP + Ir;
        unsigned mask = don[R];
        const int sr = rcv[R];                                    // requested together with the mask: one round trip per hop
        const double qr = is_ghost(g, Ir, Jr) ? 0. : q0i;
        if ((mask & (mask - 1u)) == 0u) {
            // R's only donor: no rendezvous needed, carry the value on in a register (most of a river)
            acc = qr + acc;
        } else {
            __threadfence();
            const unsigned sh = 8u * (unsigned)(R & 3);
            const unsigned old = atomicAdd(cnt + (R >> 2), 1u << sh);
            if ((int)((old >> sh) & 0xFFu) + 1 != __popc(mask)) break;   // somebody else will finish R
            __threadfence();
            acc = qr;
            while (mask) {                                        // donors in descending mesh id
                int best = -1, bid = -1;
                for (unsigned mm = mask; mm; mm &= mm - 1) {
                    const int o = __ffs(mm) - 1;
                    const int id = mesh_id(g, Ir + c_di[o], Jr + c_dj[o]);
                    if (id > bid) { bid = id; best = o; }
                }
                acc = acc + __ldcg(Q + (size_t)(Jr + c_dj[best]) * g.LXP + (Ir + c_di[best]));
                mask &= ~(1u << best);
            }
        }
        I = Ir; J = Jr; L = R; s = sr;
    }
}

// flowcfl (FLOWalgo.f90:299-330) on the filled surface (buildFlux.py:169)
__global__ void __launch_bounds__(256) k_cfl(GConst g)
{
    const int p = blockIdx.z;
    if (!g.running[p]) return;
    const int I = blockIdx.x * 32 + threadIdx.x, J = blockIdx.y * 8 + threadIdx.y;
    double tmp = 1.e6;
    if (I < g.LX && J < g.LY) {
        const size_t base = (size_t)p * g.NLP, L = (size_t)J * g.LXP + I;
        const int s = g.rcv[base + L];
        if (s != 8) {
            const double *W = g.W + base;
            const double dz = W[L] - W[(size_t)(J + c_dj[s]) * g.LXP + (I + c_di[s])];
            const double q = g.Q[base + L];
            if (dz > 0. && q > 0.) {
                const double dist = g.dist[s];
                tmp = dist / (g.erod[p] * powm(q, g.m) * powm(dz / dist, g.n - 1.));
            }
        }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) tmp = fmin(tmp, __shfl_xor_sync(0xffffffffu, tmp, o));
    __shared__ double s_min[8];
    if (threadIdx.x == 0) s_min[threadIdx.y] = tmp;
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        for (int w = 1; w < 8; ++w) tmp = fmin(tmp, s_min[w]);
        if (tmp < 1.e6) atomicMin(g.cflbits + p, (unsigned long long)__double_as_longlong(tmp));   // positive doubles order as integers
    }
}

// buildFlux.py:154-177 + flowNetwork.py:687-698 (newdt = py2-rounded dt; the re-run uses it)
__global__ void k_dt(GConst g)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= g.B || !g.running[p]) return;
    const double flowcfl = __longlong_as_double((long long)g.cflbits[p]);
    double cfl = fmin(flowcfl, g.hill);
    cfl = floorish(cfl);
    cfl = fmin(cfl, g.stops[g.si[p]] - g.tnow[p]);
    cfl = fmax(g.minDT, cfl);
    cfl = fmin(g.maxDT, cfl);
    double newdt = floorish(cfl);
    newdt = fmax(g.minDT, newdt);
    g.dt[p] = newdt;
}

// detachment-limited erosion (FLOWalgo.f90:635-714, :875-892; flowNetwork.py:706-712, buildFlux.py:193-195).
// Writes the eroded surface over W (fillH is dead after this kernel: only the node's own value is read).
__global__ void __launch_bounds__(256) k_erode(GConst g)
{
    const int p = blockIdx.z;
    if (!g.running[p]) return;
    const int I = blockIdx.x * 32 + threadIdx.x, J = blockIdx.y * 8 + threadIdx.y;
    if (I >= g.LX || J >= g.LY) return;
    const size_t base = (size_t)p * g.NLP, L = (size_t)J * g.LXP + I;
    const double *z = g.z + base;
    const int s = g.rcv[base + L];
    const double zd = z[L];
    double sed = 0.;
    if ((g.flags[L] & 1) && s != 8) {
        const double zr = z[(size_t)(J + c_dj[s]) * g.LXP + (I + c_di[s])];
        const double sea = g.sea, dt = g.dt[p];
        double dh = 0.95 * (zd - zr);
        if (zd > sea && zr < sea) dh = 0.99 * (zd - sea);
        if (dh < 0.001) dh = 0.;
        const double fh = g.W[base + L];
        const double waterH = fh - zd;
        if (dh > 0. && waterH == 0. && fh >= sea) {
            const double dist = g.dist[s];
            const double slp = dh / dist;
            if (dist > 0.) {
                double SPL = -g.erod[p] * 1. * 1. * powm(g.Q[base + L], g.m) * powm(slp, g.n);
                double totspl = 0. + SPL;
                if (-totspl * dt > dh) {
                    const double frac = dh / (-totspl * dt);
                    SPL = SPL * frac;
                    totspl = -dh;
                }
                if (totspl < 0.) {
                    const double erodep = SPL * dt * g.area;
                    const double cero = 0. + erodep;
                    sed = cero / g.area + 0.;
                }
            }
        }
    }
    g.W[base + L] = zd + sed;
    g.cumdiff[base + L] += sed;
}

// hillslope diffusion (sfd.c:155-185; diffLinear.py:150-182; buildFlux.py:252-272) on the eroded surface held in W;
// FUSE: no boundary rule to apply (btype fixed) -> uplift (buildFlux.py:312-313) in the same pass
template <int FUSE>
__global__ void __launch_bounds__(256) k_diffuse(GConst g)
{
    const int p = blockIdx.z;
    if (!g.running[p]) return;
    const int I = blockIdx.x * 32 + threadIdx.x, J = blockIdx.y * 8 + threadIdx.y;
    if (I >= g.LX || J >= g.LY) return;
    const size_t base = (size_t)p * g.NLP, L = (size_t)J * g.LXP + I;
    const double *zs = g.W + base;
    const unsigned char fl = g.flags[L];
    const double zg = zs[L], dt = g.dt[p];
    double acc = 0.;
    if (fl & 2) {
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const double ed = g.vor[s], di = g.edge[s];
            if (ed == 0.) continue;                      // diagonal slots carry no Voronoi edge: the term is +-0
            const size_t L2 = (size_t)(J + c_dj[s]) * g.LXP + (I + c_di[s]);
            const double zj = zs[L2];
            if (g.flags[L2] & 2) { if (di > 0.) acc += ed * (zj - zg) / di; }
            else if (zj < zg && di > 0.) acc += ed * (zj - zg) / di;
        }
    }
    double znew = zg;
    if (fl & 1) {
        const double ac = (g.area > 0.) ? 1. / g.area : 0.;
        double coeff = ac * ((zg >= g.sea) ? g.CDa : g.CDm);
        if (!(fl & 2)) { coeff = 0.; acc = 0.; }
        const double flux = coeff * acc * dt;
        znew = zg + flux;
        g.cumdiff[base + L] += flux;
        g.cumhill[base + L] += flux;
    }
    if (FUSE && g.disp) znew += g.disp[L] * dt;
    g.z[base + L] = znew;
}

// buildFlux.py:292-302 ghost nodes follow their parent; then uplift on every node (:312-313)
__global__ void k_boundary(GConst g)
{
    const int p = blockIdx.y;
    if (!g.running[p]) return;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= g.nb) return;
    double *z = g.z + (size_t)p * g.NLP;
    z[g.ghostL[k]] = z[g.parentL[k]] + g.boff;
}
__global__ void k_uplift(GConst g)
{
    const int p = blockIdx.y;
    if (!g.running[p]) return;
    const size_t L = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (L >= (size_t)g.LY * g.LXP || (int)(L % g.LXP) >= g.LX) return;
    g.z[(size_t)p * g.NLP + L] += g.disp[L] * g.dt[p];
}

// checkpoints: lattice cell = node (the k=3 IDW table of interpolateArray degenerates to exact hits), SSE against the
// observed grid in a fixed reduction order (bl_mcmc.py:482-513), optional grid copies
__global__ void __launch_bounds__(256) k_ckpt(GConst g, double *out_elev, double *out_erdp, int nck)
{
    const int p = blockIdx.y;
    const int ck = g.ckflag[p];
    if (!ck) return;
    const size_t Gn = (size_t)g.LX * g.LY;
    double acc = 0.;
    for (size_t cell = (size_t)blockIdx.x * blockDim.x + threadIdx.x; cell < Gn; cell += (size_t)gridDim.x * blockDim.x) {
        const int J = (int)(cell / g.LX), I = (int)(cell % g.LX);
        const size_t o = (size_t)p * g.NLP + (size_t)J * g.LXP + I;
        const double zv = g.z[o];
        if (g.real) { const double d = zv - g.real[(size_t)J * g.LXP + I]; acc += d * d; }
        if (out_elev || out_erdp) {
            const double cd = g.cumdiff[o];
            for (int k = 0; k < nck; ++k) if ((ck >> k) & 1) {
                if (out_elev) out_elev[((size_t)p * nck + k) * Gn + cell] = zv;
                if (out_erdp) out_erdp[((size_t)p * nck + k) * Gn + cell] = cd;
            }
        }
    }
    __shared__ double s_acc[256];
    s_acc[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o >= 1; o >>= 1) { if ((int)threadIdx.x < o) s_acc[threadIdx.x] += s_acc[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) g.partial[(size_t)p * g.npart + blockIdx.x] = s_acc[0];
}
__global__ void k_ckpt_final(GConst g, double *out_sse, double *out_pts)
{
    const int p = blockIdx.x;
    const int ck = g.ckflag[p];
    if (!ck) return;
    if (threadIdx.x == 0 && out_sse && g.real) {
        double a = 0.;
        for (int i = 0; i < g.npart; ++i) a += g.partial[(size_t)p * g.npart + i];
        out_sse[p] = a;
    }
    if (out_pts && (int)threadIdx.x < g.npts) {
        const int cell = g.pts_cell[threadIdx.x];
        const size_t o = (size_t)p * g.NLP + (size_t)(cell / g.LX) * g.LXP + (cell % g.LX);
        for (int k = 0; k < BL_NCKPT_MAX; ++k) if ((ck >> k) & 1) out_pts[((size_t)p * BL_NCKPT_MAX + k) * BL_NPTS_MAX + threadIdx.x] = g.cumdiff[o];
    }
}

} // namespace grid

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
struct bl_grid {
    grid::GConst g;
    int device, B, debug;
    int group;                 // proposals per fill group (bl_params.opt_grid_group, default 64)
    int coop_blocks;           // co-resident blocks of the cooperative fill loop (SMs x occupancy)
    long long launches;
    long long last_sweeps;     // sweeps the previous step's fill needed (first chunk size)
    int *h_counters;           // pinned [4]
    double *d_stops;
    int *d_ckpt;
};

namespace {
constexpr int GRID_STOPS_CAP = 4096;
constexpr int GRID_NPART = 64;

struct GLayout {
    size_t flags, z0, disp, real, parentL, ghostL, pts, stops, ckpt, scal, z, W, Q, cumdiff, cumhill, fdbg, rcv, don, cnt, l0, wl, wlctl, wlcap,
        partial, total;
    size_t NLP;
    int LX, LY, LXP, ncol, W0, W1, W0p, W1p, nb;
};

bool grid_layout(const bl_lattice *lt, int B, int debug, GLayout &L)
{
    if (!lt || B <= 0 || lt->nx < 2 || lt->ny < 2 || (lt->nx & 1) || (lt->ny & 1)) return false;
    L.LX = lt->nx + 2; L.LY = lt->ny + 2;
    L.LXP = (int)rup((size_t)L.LX, 16);
    L.NLP = rup((size_t)L.LY * L.LXP, 64);
    L.ncol = (lt->nx / 2) * (lt->ny / 2);
    L.W0 = (L.ncol + 31) / 32; L.W1 = (L.W0 + 31) / 32;
    L.W0p = (int)rup((size_t)L.W1 * 32, 32); L.W1p = (int)rup((size_t)L.W1, 32);
    L.nb = 2 * L.LX + 2 * lt->ny;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = rup(o + bytes, 256); return r; };
    L.flags = take(L.NLP);
    L.z0 = take(L.NLP * 8);
    L.disp = take(L.NLP * 8);
    L.real = take(L.NLP * 8);
    L.parentL = take((size_t)L.nb * 4);
    L.ghostL = take((size_t)L.nb * 4);
    L.pts = take(BL_NPTS_MAX * 4);
    L.stops = take((size_t)GRID_STOPS_CAP * 8);
    L.ckpt = take((size_t)GRID_STOPS_CAP * 4);
    L.scal = take((size_t)B * 256 + 256);
    L.partial = take((size_t)B * GRID_NPART * 8);
    L.z = take((size_t)B * L.NLP * 8);
    L.W = take((size_t)B * L.NLP * 8);
    L.Q = take((size_t)B * L.NLP * 8);
    L.cumdiff = take((size_t)B * L.NLP * 8);
    L.cumhill = take((size_t)B * L.NLP * 8);
    L.fdbg = take(debug ? (size_t)B * L.NLP * 8 : 8);
    L.rcv = take((size_t)B * L.NLP);
    L.don = take((size_t)B * L.NLP);
    L.cnt = take((size_t)B * L.NLP);
    L.l0 = take((size_t)B * 4 * L.W0p * 4);
    L.wlcap = 1; while (L.wlcap < (size_t)B * L.W0) L.wlcap <<= 1;
    L.wl = take((size_t)4 * L.wlcap * 4);
    L.wlctl = take(4 * 128);
    L.total = o;
    return true;
}
} // namespace

extern "C" size_t bl_grid_workspace_bytes(const bl_lattice *lt, int B, int debug)
{
    GLayout L;
    return grid_layout(lt, B, debug, L) ? L.total : 0;
}

extern "C" int bl_grid_create(const bl_lattice *lt, const bl_params *p, int B, int device, void *workspace,
                              size_t workspace_bytes, int debug, bl_grid **out)
{
    if (!lt || !p || !out || !workspace || !lt->z0 || !lt->flags) { g_err = "bl_grid_create: null argument"; return BL_ERR_ARG; }
    GLayout L;
    if (!grid_layout(lt, B, debug, L)) { g_err = "bl_grid_create: nx, ny must be even and >= 2, B > 0"; return BL_ERR_ARG; }
    if (p->depo != 0 || p->nsea > 0 || p->CDr > 0.) {
        g_err = "bl_grid_create: the lattice engine covers the erosive configuration only (depo = 0, constant sea level, CDr = 0)";
        return BL_ERR_UNSUPPORTED;
    }
    if (p->npts > BL_NPTS_MAX) { g_err = "bl_grid_create: too many erdp points"; return BL_ERR_ARG; }
    if (p->btype != 0 && !lt->parent) { g_err = "bl_grid_create: parent ids are needed for this boundary type"; return BL_ERR_ARG; }
    if (workspace_bytes < L.total) { g_err = "bl_grid_create: workspace too small"; return BL_ERR_SPACE; }
    CUDA_OK(cudaSetDevice(device));
    char *ws = (char *)workspace;
    const size_t Gn = (size_t)L.LX * L.LY;
    // pitch the lattice rows
    std::vector<unsigned char> flags(L.NLP, 0);
    std::vector<double> z0(L.NLP, 0.), disp(L.NLP, 0.), real(L.NLP, 0.);
    for (int J = 0; J < L.LY; ++J)
        for (int I = 0; I < L.LX; ++I) {
            const size_t s = (size_t)J * L.LX + I, d = (size_t)J * L.LXP + I;
            flags[d] = lt->flags[s]; z0[d] = lt->z0[s];
            if (lt->disp) disp[d] = lt->disp[s];
            if (lt->real_elev) real[d] = lt->real_elev[s];
        }
    (void)Gn;
    std::vector<int> ghostL(L.nb), parentL(L.nb, 0);
    {
        int k = 0;
        for (int I = 0; I < L.LX; ++I) ghostL[k++] = I;                                    // south row
        for (int J = 1; J <= lt->ny; ++J) ghostL[k++] = J * L.LXP;                          // x = -dx column
        for (int I = 0; I < L.LX; ++I) ghostL[k++] = (L.LY - 1) * L.LXP + I;                // north row
        for (int J = 1; J <= lt->ny; ++J) ghostL[k++] = J * L.LXP + (L.LX - 1);             // x = nx*dx column
    }
    if (lt->parent)
        for (int k = 0; k < L.nb; ++k) {
            const int s = lt->parent[k];
            if (s < 0 || (size_t)s >= (size_t)L.LX * L.LY) { g_err = "bl_grid_create: parent index out of range"; return BL_ERR_ARG; }
            parentL[k] = (s / L.LX) * L.LXP + (s % L.LX);
        }
    std::vector<int> pts(BL_NPTS_MAX, 0);
    for (int i = 0; i < p->npts; ++i) pts[i] = p->pts_cell[i];
#define UP(off, src, bytes) CUDA_OK(cudaMemcpy(ws + (off), (src), (bytes), cudaMemcpyHostToDevice))
    UP(L.flags, flags.data(), L.NLP);
    UP(L.z0, z0.data(), L.NLP * 8);
    UP(L.disp, disp.data(), L.NLP * 8);
    UP(L.real, real.data(), L.NLP * 8);
    UP(L.parentL, parentL.data(), (size_t)L.nb * 4);
    UP(L.ghostL, ghostL.data(), (size_t)L.nb * 4);
    UP(L.pts, pts.data(), BL_NPTS_MAX * 4);
#undef UP
    CUDA_OK(cudaMemset(ws + L.scal, 0, (size_t)B * 256 + 256));
    bl_grid *ctx = new bl_grid();
    grid::GConst &g = ctx->g;
    memset(&g, 0, sizeof(g));
    g.nx = lt->nx; g.ny = lt->ny; g.LX = L.LX; g.LY = L.LY; g.LXP = L.LXP; g.B = B; g.hx = lt->nx / 2; g.ncol = L.ncol;
    g.W0 = L.W0; g.W1 = L.W1; g.W0p = L.W0p; g.W1p = L.W1p; g.nb = L.nb; g.btype = p->btype; g.npts = p->npts;
    g.has_parent = lt->parent ? 1 : 0;
    g.NLP = L.NLP;
    g.m = p->spl_m; g.n = p->spl_n; g.eps = p->eps; g.TH = p->fill_th; g.CDa = p->CDa; g.CDm = p->CDm;
    g.minDT = p->minDT; g.maxDT = p->maxDT; g.hill = p->hill_cfl; g.sea = p->seapos; g.area = lt->area;
    g.boff = (p->btype == 1) ? -0.1 : (p->btype == 3 ? 100. : 0.);
    for (int s = 0; s < 8; ++s) { g.dist[s] = lt->dist[s]; g.vor[s] = lt->vor[s]; g.edge[s] = lt->edge[s]; }
    g.flags = (const unsigned char *)(ws + L.flags);
    g.z0 = (const double *)(ws + L.z0);
    g.disp = lt->disp ? (const double *)(ws + L.disp) : nullptr;
    g.real = lt->real_elev ? (const double *)(ws + L.real) : nullptr;
    g.parentL = (const int *)(ws + L.parentL);
    g.ghostL = (const int *)(ws + L.ghostL);
    g.pts_cell = (const int *)(ws + L.pts);
    g.z = (double *)(ws + L.z); g.W = (double *)(ws + L.W); g.Q = (double *)(ws + L.Q);
    g.cumdiff = (double *)(ws + L.cumdiff); g.cumhill = (double *)(ws + L.cumhill);
    g.fdbg = debug ? (double *)(ws + L.fdbg) : nullptr;
    g.rcv = (unsigned char *)(ws + L.rcv); g.don = (unsigned char *)(ws + L.don); g.cnt = (unsigned *)(ws + L.cnt);
    g.l0 = (unsigned *)(ws + L.l0); g.wl = (unsigned *)(ws + L.wl); g.wlctl = (void *)(ws + L.wlctl); g.wlcap = L.wlcap;
    {
        char *s = ws + L.scal;
        const size_t Bs = (size_t)B;
        g.rain = (double *)s; g.erod = g.rain + Bs; g.tnow = g.erod + Bs; g.dt = g.tnow + Bs; g.lastdt = g.dt + Bs;
        g.cflbits = (unsigned long long *)(g.lastdt + Bs);
        g.steps = (long long *)(g.cflbits + Bs); g.sweeps = g.steps + Bs;
        g.running = (int *)(g.sweeps + Bs); g.fill_active = g.running + Bs;
        g.si = g.fill_active + Bs; g.ckflag = g.si + Bs; g.counters = g.ckflag + Bs;
        g.change = (int *)(s + rup(Bs * 84 + 64, 128));          // [B][CHS]
    }
    g.partial = (double *)(ws + L.partial);
    g.npart = GRID_NPART;
    ctx->d_stops = (double *)(ws + L.stops);
    ctx->d_ckpt = (int *)(ws + L.ckpt);
    g.stops = ctx->d_stops; g.ckpt = ctx->d_ckpt; g.nstops = 0;
    ctx->device = device; ctx->B = B; ctx->debug = debug; ctx->launches = 0; ctx->last_sweeps = 0;
    ctx->group = (p->opt_grid_group > 0) ? p->opt_grid_group : 64;
    {
        int nsm = 148, occ = 1, coop = 0;
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
        if (!coop) { g_err = "bl_grid_create: the device does not support cooperative launches"; delete ctx; return BL_ERR_UNSUPPORTED; }
        CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, grid::k_fill_loop, 256, 0));
        const int want = (p->opt_debug >> 4) & 15;          // experiment knob: blocks per SM of the fill loop
        ctx->coop_blocks = nsm * std::max(1, std::min(occ, want ? want : 8));
    }
    CUDA_OK(cudaMallocHost((void **)&ctx->h_counters, 4 * sizeof(int)));
    *out = ctx;
    return BL_OK;
}

extern "C" int bl_grid_destroy(bl_grid *ctx)
{
    if (ctx) { if (ctx->h_counters) cudaFreeHost(ctx->h_counters); delete ctx; }
    return BL_OK;
}

extern "C" int bl_grid_reset(bl_grid *ctx, const double *rain_dev, const double *erod_dev, double tStart, void *stream)
{
    if (!ctx || !rain_dev || !erod_dev) { g_err = "bl_grid_reset: null argument"; return BL_ERR_ARG; }
    CUDA_OK(cudaSetDevice(ctx->device));
    const grid::GConst &g = ctx->g;
    dim3 gr((unsigned)((g.NLP + 255) / 256), (unsigned)ctx->B);
    grid::k_reset<<<gr, 256, 0, (cudaStream_t)stream>>>(g, rain_dev, erod_dev, tStart);
    ctx->launches++;
    ctx->last_sweeps = 0;
    CUDA_OK(cudaGetLastError());
    return BL_OK;
}

namespace {
// one time step for every running proposal; returns the number of proposals still running afterwards (via h_counters)
int grid_one_step(bl_grid *ctx, double *out_pts, double *out_sse, double *out_elev, double *out_erdp, int nck, cudaStream_t st)
{
    const grid::GConst &g = ctx->g;
    const int B = ctx->B;
    const dim3 b2(32, 8), g2((unsigned)((g.LX + 31) / 32), (unsigned)((g.LY + 7) / 8), (unsigned)B);
    const dim3 g2p((unsigned)((g.LXP + 31) / 32), (unsigned)((g.LY + 7) / 8), (unsigned)B);
    const dim3 g1((unsigned)((std::max((size_t)g.LY * g.LXP, (size_t)4 * g.W0) + 255) / 256), (unsigned)B);
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device);
    // Depression filling, a group of proposals at a time: the pending regions of ~64 landscapes (~100 MB of rows) stay in
    // L2 from pass to pass, those of 256 do not (measured: 5.2 ms per proposal-step at B = 256 in one group, 2.75 at 64)
    const int group = ctx->group;
    for (int p0 = 0; p0 < B; p0 += group) {
        const int nB = std::min(group, B - p0);
        const long long ntask = (long long)nB * g.W0;
        const int pass_blocks = (int)std::max<long long>(1, std::min<long long>((ntask + 7) / 8, (long long)nsm * 8));
        grid::k_fill_init<<<dim3(g1.x, (unsigned)nB), 256, 0, st>>>(g, p0);
        grid::k_fill_begin<<<1, 256, 0, st>>>(g);
        ctx->launches += 2;
        {
            grid::GConst garg = g;
            void *args[] = {(void *)&garg};
            CUDA_OK(cudaLaunchCooperativeKernel((const void *)grid::k_fill_loop, dim3((unsigned)std::min<long long>(pass_blocks, ctx->coop_blocks)),
                                                dim3(256), args, 0, st));
            ctx->launches += 1;
        }
    }
    grid::k_sfd<<<dim3((unsigned)((g.LX + 255) / 256), (unsigned)((g.LY + grid::SROWS - 1) / grid::SROWS), (unsigned)B), b2, 0, st>>>(g);
    grid::k_donors<<<dim3((unsigned)((g.LXP + 255) / 256), (unsigned)((g.LY + grid::SROWS - 1) / grid::SROWS), (unsigned)B), b2, 0, st>>>(g);
    grid::k_discharge<<<g2, b2, 0, st>>>(g);
    grid::k_cfl<<<g2, b2, 0, st>>>(g);
    grid::k_dt<<<(B + 127) / 128, 128, 0, st>>>(g);
    grid::k_erode<<<g2, b2, 0, st>>>(g);
    ctx->launches += 6;
    if (g.btype == 0 || !g.has_parent) {
        grid::k_diffuse<1><<<g2, b2, 0, st>>>(g);
        ctx->launches += 1;
    } else {
        grid::k_diffuse<0><<<g2, b2, 0, st>>>(g);
        grid::k_boundary<<<dim3((unsigned)((g.nb + 127) / 128), (unsigned)B), 128, 0, st>>>(g);
        ctx->launches += 2;
        if (g.disp) {
            grid::k_uplift<<<dim3((unsigned)(((size_t)g.LY * g.LXP + 255) / 256), (unsigned)B), 256, 0, st>>>(g);
            ctx->launches += 1;
        }
    }
    grid::k_advance<<<1, 256, 0, st>>>(g, 1, 0);
    grid::k_ckpt<<<dim3(GRID_NPART, (unsigned)B), 256, 0, st>>>(g, out_elev, out_erdp, nck);
    grid::k_ckpt_final<<<B, 32, 0, st>>>(g, out_sse, out_pts);
    ctx->launches += 3;
    CUDA_OK(cudaGetLastError());
    return BL_OK;
}

int grid_run_common(bl_grid *ctx, int nstops, const double *stops, const int32_t *ckpt, double *out_pts, double *out_sse,
                    double *out_elev, double *out_erdp, int nck, int single, cudaStream_t st)
{
    if (!ctx || nstops <= 0 || !stops) { g_err = "bl_grid_run: bad arguments"; return BL_ERR_ARG; }
    if (nstops > GRID_STOPS_CAP) { g_err = "bl_grid_run: too many stops"; return BL_ERR_ARG; }
    CUDA_OK(cudaSetDevice(ctx->device));
    std::vector<int> ck(nstops, -1);
    if (ckpt)
        for (int i = 0; i < nstops; ++i) {
            ck[i] = ckpt[i];
            if (ck[i] >= BL_NCKPT_MAX || ((out_elev || out_erdp) && ck[i] >= nck)) {
                g_err = "bl_grid_run: checkpoint index too large"; return BL_ERR_ARG;
            }
        }
    CUDA_OK(cudaMemcpyAsync(ctx->d_stops, stops, (size_t)nstops * 8, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(ctx->d_ckpt, ck.data(), (size_t)nstops * 4, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaStreamSynchronize(st));
    ctx->g.nstops = nstops;
    const grid::GConst &g = ctx->g;
    // stops already reached count as checkpoints without a step (run_to_time(t <= tNow))
    grid::k_advance<<<1, 256, 0, st>>>(g, 0, 1);
    grid::k_ckpt<<<dim3(GRID_NPART, (unsigned)ctx->B), 256, 0, st>>>(g, out_elev, out_erdp, nck);
    grid::k_ckpt_final<<<ctx->B, 32, 0, st>>>(g, out_sse, out_pts);
    ctx->launches += 3;
    CUDA_OK(cudaMemcpyAsync(ctx->h_counters, g.counters, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    long long guard = 0;
    while (ctx->h_counters[0] > 0) {
        int rc = grid_one_step(ctx, out_pts, out_sse, out_elev, out_erdp, nck, st);
        if (rc != BL_OK) return rc;
        CUDA_OK(cudaMemcpyAsync(ctx->h_counters, g.counters, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaStreamSynchronize(st));
        if (single) break;
        if (++guard > 100000000) { g_err = "bl_grid_run: step guard"; return BL_ERR_CUDA; }
    }
    return BL_OK;
}
} // namespace

extern "C" int bl_grid_run(bl_grid *ctx, int nstops, const double *stops, const int32_t *ckpt, double *out_pts,
                           double *out_sse, double *out_elev, double *out_erdp, int nck, void *stream)
{
    return grid_run_common(ctx, nstops, stops, ckpt, out_pts, out_sse, out_elev, out_erdp, nck, 0, (cudaStream_t)stream);
}

extern "C" int bl_grid_step(bl_grid *ctx, double tStop, void *stream)
{
    return grid_run_common(ctx, 1, &tStop, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 1, (cudaStream_t)stream);
}

extern "C" int bl_grid_get_field(bl_grid *ctx, int field, int proposal, void *dst, void *stream)
{
    if (!ctx || !dst || proposal < 0 || proposal >= ctx->B) { g_err = "bl_grid_get_field: bad arguments"; return BL_ERR_ARG; }
    CUDA_OK(cudaSetDevice(ctx->device));
    const grid::GConst &g = ctx->g;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t base = (size_t)proposal * g.NLP;
    const double *fsrc = nullptr;
    switch (field) {
        case BL_F_Z: fsrc = g.z; break;
        case BL_F_CUMDIFF: fsrc = g.cumdiff; break;
        case BL_F_CUMHILL: fsrc = g.cumhill; break;
        case BL_F_FILLH:
            if (!g.fdbg) { g_err = "bl_grid_get_field: fillH is kept only by a debug context"; return BL_ERR_UNSUPPORTED; }
            fsrc = g.fdbg; break;
        case BL_F_Q: fsrc = g.Q; break;
        case BL_F_RCV: break;
        default: g_err = "bl_grid_get_field: field not kept by the lattice engine"; return BL_ERR_UNSUPPORTED;
    }
    if (fsrc) {   // strip the row pitch
        CUDA_OK(cudaMemcpy2DAsync(dst, (size_t)g.LX * 8, fsrc + base, (size_t)g.LXP * 8, (size_t)g.LX * 8, g.LY,
                                  cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaStreamSynchronize(st));
        return BL_OK;
    }
    std::vector<unsigned char> code((size_t)g.LX * g.LY);
    CUDA_OK(cudaMemcpy2DAsync(code.data(), (size_t)g.LX, g.rcv + base, (size_t)g.LXP, (size_t)g.LX, g.LY, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    static const int di[8] = {1, 1, 0, -1, -1, -1, 0, 1}, dj[8] = {0, 1, 1, 1, 0, -1, -1, -1};
    int32_t *out = (int32_t *)dst;                       // receiver as a lattice cell index J*LX + I
    for (int J = 0; J < g.LY; ++J)
        for (int I = 0; I < g.LX; ++I) {
            const int s = code[(size_t)J * g.LX + I];
            out[(size_t)J * g.LX + I] = (s == 8) ? J * g.LX + I : (J + dj[s]) * g.LX + (I + di[s]);
        }
    return BL_OK;
}

extern "C" int bl_grid_get_scalars(bl_grid *ctx, double *tnow, int64_t *steps, int64_t *sweeps, double *last_dt, void *stream)
{
    if (!ctx) { g_err = "bl_grid_get_scalars: null ctx"; return BL_ERR_ARG; }
    CUDA_OK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t B = ctx->B;
    if (tnow) CUDA_OK(cudaMemcpyAsync(tnow, ctx->g.tnow, B * 8, cudaMemcpyDeviceToHost, st));
    if (steps) CUDA_OK(cudaMemcpyAsync(steps, ctx->g.steps, B * 8, cudaMemcpyDeviceToHost, st));
    if (sweeps) CUDA_OK(cudaMemcpyAsync(sweeps, ctx->g.sweeps, B * 8, cudaMemcpyDeviceToHost, st));
    if (last_dt) CUDA_OK(cudaMemcpyAsync(last_dt, ctx->g.lastdt, B * 8, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    return BL_OK;
}

extern "C" int64_t bl_grid_launch_count(const bl_grid *ctx) { return ctx ? ctx->launches : 0; }

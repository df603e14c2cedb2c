// bl_device.cuh - device code of the engine.  Included by bl_engine.cu once per CTA size (inside a namespace
// that defines TPB / NWARP / BL_MINB): 1024 threads for landscapes that fill an SM's shared memory, 256 threads
// (several CTAs per SM) for small meshes such as the etopo configs, 512 threads (128 registers, no spills; BL_TPB=512)
// as a measured alternative.  No include guard on purpose.
// ------------------------------------------------------------------------------------------------
// block primitives
// ------------------------------------------------------------------------------------------------
struct Smem {
    DevConst c;                // kernel-parameter copy: device functions take it by reference, and a
                               // reference to a by-value kernel parameter would live in (L2-latency) local memory
    Prop P;
    int lev_off[LEVMAX + 1];   // copy of c.lev_off: every CTA reads it every pass (one hot L2 line otherwise)
    int fl_head[LEVMAX + 1], fl_tail[LEVMAX + 1], fl_base[LEVMAX + 1], fl_mask[LEVMAX + 1];   // per-level worklist rings of the fill
    unsigned char *dsm;        // dynamic shared memory of the fast kernel (nullptr on the generic path)
    unsigned short boff[2048]; // depth-bucket offsets of the tree passes (fast path)
    int dsm_bytes;
    int scan[NWARP + 1];
    double red[NWARP];
    int ibuf[8];
    double dbuf[8];
    // Scalars of the proposal that live across the whole run.  They are kept here and re-read where needed instead of
    // being carried in registers: the kernel is capped at 64 registers per thread and this state alone held ~30 of them
    // through every phase (seen as spills and as S2R re-materialisation inside the inner loops).
    double tNow, last_dt, rain, erod, sea, tStop, newdt;
    long long step, tick0;
    unsigned long long seed;
    int nbase1, nbase, any_marine;
    double *diff_out;          // per-stage entry point bl_diffuse: raw diffusion flux of sfd.diffusion (null in a run)
    unsigned long long mbar;   // mbarrier of the TMA bulk copies global -> shared memory (bl_fastc.cuh: tma_to_smem)
    unsigned tma_phase;        // its phase parity
};

// diagnostics: thread 0 adds the cycles since the previous tick to slot i of this CTA's phase_cyc row (slots >= 16)
struct SubTick { long long *p, t; };
__device__ inline long long sub_clock() { return clock64(); }
__device__ inline SubTick sub_tick_begin(const DevConst &c) { SubTick s; s.p = c.phase_cyc + (size_t)blockIdx.x * NPHASE; s.t = sub_clock(); return s; }
// (__syncwarp: lane 0 must be back with its warp before the caller's next warp-synchronous instruction)
// (counters are bumped with a result-less atomic: a plain += would put a global load on thread 0's path at every tick)
__device__ inline void tick_add(long long *p, long long v) { atomicAdd(reinterpret_cast<unsigned long long *>(p), (unsigned long long)v); }
__device__ inline void sub_tick(SubTick &s, int i) { if (threadIdx.x == 0) { const long long t1 = sub_clock(); tick_add(s.p + i, t1 - s.t); s.t = t1; } __syncwarp(); }
__device__ inline void sub_count(SubTick &s, int i, int n) { if (threadIdx.x == 0) tick_add(s.p + i, n); __syncwarp(); }

__device__ inline int blk_exscan(int v, int &total, Smem &S)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) S.scan[w] = inc;
    __syncthreads();
    if (w == 0) {
        int x = (lane < NWARP) ? S.scan[lane] : 0;
        int xi = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, xi, o);
            if (lane >= o) xi += t;
        }
        if (lane < NWARP) S.scan[lane] = xi - x;
        if (lane == 31) S.scan[NWARP] = xi;
    }
    __syncthreads();
    int res = inc - v + S.scan[w];
    total = S.scan[NWARP];
    __syncthreads();
    return res;
}

__device__ inline double blk_min(double v, Smem &S)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) S.red[w] = v;
    __syncthreads();
    double r = (lane < NWARP) ? S.red[lane] : 1.e300;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r = fmin(r, __shfl_xor_sync(0xffffffffu, r, o));
    __syncthreads();
    return r;
}

__device__ inline double blk_sum(double v, Smem &S)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) S.red[w] = v;
    __syncthreads();
    double r = (lane < NWARP) ? S.red[lane] : 0.;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    __syncthreads();
    return r;
}

// ordered compaction: out[] receives, in ascending index order, every i in [0,n) with pred(i) != 0.
template <class Pred>
__device__ inline int blk_compact(int n, int *out, Pred pred, Smem &S)
{
    // warp w owns the contiguous range [w * per, (w + 1) * per): lanes read it 32 consecutive elements at a time
    // (coalesced, independent loads), ballots give the ranks; one block scan over the warp totals orders the warps
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int per = (((n + NWARP - 1) / NWARP) + 31) & ~31;
    const int lo = w * per, nj = per >> 5;
    const bool keep = nj <= 64;                  // predicate bits of this lane's elements, one per 32-element group
    unsigned long long mask = 0ull;
    int wtot = 0;
#pragma unroll 4
    for (int j = 0; j < nj; ++j) {
        const int i = lo + 32 * j + lane;
        const bool pr = (i < n) && pred(i);
        if (pr && keep) mask |= 1ull << j;
        wtot += __popc(__ballot_sync(0xffffffffu, pr));
    }
    int total;
    int off = blk_exscan(lane == 0 ? wtot : 0, total, S);
    off = __shfl_sync(0xffffffffu, off, 0);
    if (wtot) {
        for (int j = 0; j < nj; ++j) {
            const int i = lo + 32 * j + lane;
            const bool pr = keep ? ((mask >> j) & 1ull) : ((i < n) && pred(i));
            const unsigned b = __ballot_sync(0xffffffffu, pr);
            if (pr) out[off + __popc(b & ((1u << lane) - 1u))] = i;
            off += __popc(b);
        }
    }
    __syncthreads();
    return total;
}

// request a line into L2 (streaming passes over per-proposal arrays: 296 landscapes do not fit the L2, the data is in DRAM)
__device__ __forceinline__ void pf_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ inline double ldv(const double *p) { return __ldcg(p); }
__device__ inline void stv(double *p, double v) { __stcg(p, v); }

__device__ inline double powx(double x, double e)
{
    if (e == 0.5) return sqrt(x);
    if (e == 1.0) return x;
    if (e == 0.0) return 1.0;
    return pow(x, e);
}

__device__ inline unsigned long long mix64(unsigned long long x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// forceSim.getSea (forcing/forceSim.py:226-245) with scipy interp1d arithmetic
__device__ inline double sea_level(const DevConst &c, double t)
{
    if (c.nsea <= 0) return c.seapos;
    const double *x = c.seat, *y = c.seav;
    const int n = c.nsea;
    if (t < x[0]) t = x[0];
    if (t > x[n - 1]) t = x[n - 1];
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (x[mid] < t) lo = mid + 1; else hi = mid;
    }
    int idx = lo;
    if (idx < 1) idx = 1;
    if (idx > n - 1) idx = n - 1;
    double slope = (y[idx] - y[idx - 1]) / (x[idx] - x[idx - 1]);
    return slope * (t - x[idx - 1]) + y[idx - 1];
}

// ------------------------------------------------------------------------------------------------
// phase: depression filling.  PDalgo.f90:19-35 (initialisePD) + :37-91 (fillPD), literal Gauss-Seidel
// semantics: the sweep visits nodes in ascending id and sees new values of lower-id neighbours.  Nodes
// are grouped by the level of that dependency DAG (host, bl_ctx_create): within a level no two nodes are
// adjacent, so a level is one parallel pass and the sweep result equals the sequential one bit for bit,
// including the reference's early stop (`change` is not raised by dry-outs or capped writes, :58-75).
// ------------------------------------------------------------------------------------------------
__device__ void phase_fill(const DevConst &c, Smem &S, double sea)
{
    const int N = c.N, tid = threadIdx.x;
    double *W = S.P.f[F_FILLH];
    const double *z = S.P.f[F_Z];
    for (int i = tid; i < N; i += TPB) W[i] = (i < c.bounds) ? z[i] : 1.e6;
    __syncthreads();
    const double eps = c.eps, TH = c.TH;
    int change = 1;
    while (change) {
        int loc = 0;
        for (int l = 0; l < c.nlev; ++l) {
            const int e = (c.nlev <= LEVMAX) ? S.lev_off[l + 1] : c.lev_off[l + 1];
            const int b0 = (c.nlev <= LEVMAX) ? S.lev_off[l] : c.lev_off[l];
            for (int idx = b0 + tid; idx < e; idx += TPB) {
                const int k = c.lev_nodes[idx];
                const double w = W[k], zk = z[k];
                if (w > zk) {
                    double hmin = 2.e6;
                    for (int p = 0; p < c.KD; ++p) {
                        int j = c.ngb[(size_t)p * N + k];
                        if (j < 0) break;
                        hmin = fmin(hmin, W[j]);
                    }
                    const double he = hmin + eps;
                    if (zk >= he) W[k] = zk;
                    else if (w > he) {
                        double nw = he;
                        if (zk >= sea) { if (nw - zk > TH) nw = zk + TH; else loc = 1; }
                        else           { if (nw - sea > TH) nw = sea + TH; else loc = 1; }
                        W[k] = nw;
                    }
                }
            }
            __syncthreads();
        }
        change = __syncthreads_or(loc);
    }
}

// phase: receivers.  sfd.c:113-153 (directions_base) fused with sfd.c:55-111 (directions)
__device__ void phase_sfd(const DevConst &c, Smem &S)
{
    const int N = c.N;
    const double *z = S.P.f[F_Z], *W = S.P.f[F_FILLH];
    int *rcv = S.P.i[I_RCV], *rcv1 = S.P.i[I_RCV1];
    double *maxh = S.P.f[F_MAXH];
    for (int g = threadIdx.x; g < N; g += TPB) {
        const double zg = z[g], wg = W[g];
        int low1 = g, low = g;
        double zl = zg, wl = wg, dH = 1.e6;
        for (int p = 0; p < c.KD; ++p) {
            int j = c.ngb[(size_t)p * N + g];
            if (j < 0) break;
            const double zj = z[j], wj = W[j];
            if (zj < zl) { low1 = j; zl = zj; }
            if (wj < wl) { low = j; wl = wj; }
            const double dh = zj - zg;
            if (dh >= 0. && dh < dH) dH = dh;
        }
        rcv1[g] = low1;
        rcv[g] = low;
        if (dH > 9.99e5) dH = 0.;
        maxh[g] = dH;
    }
    __syncthreads();
}

// donor links of a receiver tree.  FLWnetwork.f90:52-55 builds the donor CSR in ascending id; the same
// order is kept here as first-child / next-sibling (and last-child / previous-sibling) links found by
// scanning the neighbour list, since a node's donors are always among its mesh neighbours.
__device__ void tree_links(const DevConst &c, const int *rcv, int *fc, int *ns, int *lc, int *ps, int *cnt)
{
    const int N = c.N;
    for (int v = threadIdx.x; v < N; v += TPB) {
        const int r = rcv[v];
        int f = NONE_I, l = NONE_I, n = 0;
        for (int p = 0; p < c.KD; ++p) {
            int w = c.ngb[(size_t)p * N + v];
            if (w < 0) break;
            if (rcv[w] == v) { ++n; if (f == NONE_I || w < f) f = w; if (l == NONE_I || w > l) l = w; }
        }
        fc[v] = f;
        if (lc) { lc[v] = l; cnt[v] = n; }
        int nx = NONE_I, pv = NONE_I;
        if (r != v)
            for (int p = 0; p < c.KD; ++p) {
                int w = c.ngb[(size_t)p * N + r];
                if (w < 0) break;
                if (w != v && rcv[w] == r) {
                    if (w > v && (nx == NONE_I || w < nx)) nx = w;
                    if (w < v && (pv == NONE_I || w > pv)) pv = w;
                }
            }
        ns[v] = nx;
        if (ps) ps[v] = pv;
    }
    __syncthreads();
}

// block bitonic sort of n2 (power of two) u64 keys, ascending
__device__ void blk_bitonic(unsigned long long *K, int n2)
{
    for (int k = 2; k <= n2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n2; i += TPB) {
                int ixj = i ^ j;
                if (ixj > i) {
                    unsigned long long a = K[i], b = K[ixj];
                    bool up = ((i & k) == 0);
                    if ((a > b) == up) { K[i] = b; K[ixj] = a; }
                }
            }
            __syncthreads();
        }
}

// phase: stack order.  fstack.build (FLWnetwork.f90:21-71) + addtostack (FLOWstack.f90:24-40): pre-order
// DFS from each base in base order, donors ascending.  Rebuilt as an Euler tour of the forest
// (enter(v) -> first child | exit(v); exit(v) -> next sibling | exit(parent) | next root) ranked by
// pointer jumping: position(v) = N - #enter events from enter(v) to the end of the tour.
// `shuffle`: bases ordered by counter-hash keys (numpy.random.shuffle stand-in, flowNetwork.py:316).
__device__ void phase_stack(const DevConst &c, Smem &S, const int *rcv, const int *fc, int *ns, int *stack,
                            int *pos, int *baselist, int &nbase, bool shuffle, unsigned long long seedstep)
{
    const int N = c.N, tid = threadIdx.x;
    nbase = blk_compact(N, baselist, [&](int i) { return rcv[i] == i; }, S);
    if (shuffle && nbase > 1) {
        int n2 = 1;
        while (n2 < nbase) n2 <<= 1;
        unsigned long long *K = S.P.keys;
        for (int i = tid; i < n2; i += TPB)
            K[i] = (i < nbase) ? (((mix64(seedstep ^ (unsigned long long)(unsigned)baselist[i]) >> 32) << 32) |
                                  (unsigned)baselist[i])
                               : ~0ull;
        __syncthreads();
        blk_bitonic(K, n2);
        for (int i = tid; i < nbase; i += TPB) baselist[i] = (int)(unsigned)(K[i] & 0xFFFFFFFFull);
        __syncthreads();
    }
    // chain the roots through ns[] (roots have no sibling)
    for (int i = tid; i < nbase; i += TPB) ns[baselist[i]] = (i + 1 < nbase) ? baselist[i + 1] : NONE_I;
    __syncthreads();
    unsigned long long *A = S.P.euA, *Bf = S.P.euB;
    for (int v = tid; v < N; v += TPB) {
        unsigned nxE = (fc[v] != NONE_I) ? (unsigned)(2 * fc[v]) : (unsigned)(2 * v + 1);
        unsigned nxX;
        if (ns[v] != NONE_I) nxX = (unsigned)(2 * ns[v]);
        else if (rcv[v] == v) nxX = END_U;
        else nxX = (unsigned)(2 * rcv[v] + 1);
        A[2 * v] = ((unsigned long long)nxE << 32) | 1ull;
        A[2 * v + 1] = ((unsigned long long)nxX << 32) | 0ull;
    }
    __syncthreads();
    const int M = 2 * N;
    int active = 1;
    while (active) {
        int loc = 0;
        for (int e = tid; e < M; e += TPB) {
            unsigned long long w = A[e];
            unsigned nx = (unsigned)(w >> 32);
            if (nx != END_U) {
                unsigned long long w2 = A[nx];
                unsigned rk = (unsigned)w + (unsigned)w2;
                w = (w2 & 0xFFFFFFFF00000000ull) | rk;
                loc = 1;
            }
            Bf[e] = w;
        }
        active = __syncthreads_or(loc);
        unsigned long long *t = A; A = Bf; Bf = t;
    }
    for (int v = tid; v < N; v += TPB) {
        int p = N - (int)(unsigned)A[2 * v];
        pos[v] = p;
        stack[p] = v;
    }
    __syncthreads();
}

// phase: pit ids and volumes.  FLOWalgo.f90:130-165 (basinparameters).  pitID(v) = root of v's tree on the
// real surface when v is under water (or is the root); pitVolume(root) = -1 + sum in stack order of
// (fillH - z) * area.  The sum is kept sequential per basin (one thread per basin walking its contiguous
// stack segment, dry nodes contributing +0.0) so that the fp64 result is bit-identical.
__device__ void phase_basins(const DevConst &c, Smem &S, const int *base1, int nbase1)
{
    const int N = c.N, tid = threadIdx.x;
    const double *z = S.P.f[F_Z], *W = S.P.f[F_FILLH];
    const int *rcv1 = S.P.i[I_RCV1], *pos1 = S.P.i[I_POS1], *ns1 = S.P.i[I_NS];
    int *root1 = S.P.i[I_ROOT1], *pitID = S.P.i[I_PITID];
    double *val = S.P.f[F_VAL], *pitVol = S.P.f[F_PITVOL];
    for (int v = tid; v < N; v += TPB) root1[v] = rcv1[v];
    __syncthreads();
    int ch = 1;
    while (ch) {
        int loc = 0;
        for (int v = tid; v < N; v += TPB) {
            int r = root1[v], rr = root1[r];
            if (rr != r) { root1[v] = rr; loc = 1; }
        }
        ch = __syncthreads_or(loc);
    }
    for (int v = tid; v < N; v += TPB) {
        const bool wet = W[v] > z[v];
        pitID[v] = (wet || rcv1[v] == v) ? root1[v] : -1;
        val[pos1[v]] = wet ? (W[v] - z[v]) * c.area[v] : 0.0;
        pitVol[v] = -1.0;
    }
    __syncthreads();
    for (int i = tid; i < nbase1; i += TPB) {
        const int root = base1[i];
        const int lo = pos1[root];
        const int nr = ns1[root];
        const int hi = (nr != NONE_I) ? pos1[nr] : N;
        double s = -1.0;
        for (int k = lo; k < hi; ++k) s = s + val[k];
        pitVol[root] = s;
    }
    __syncthreads();
}

// one pit's walk along the filled-surface receivers (inner loop of FLOWalgo.f90:190-234 / :260-293 with an
// empty chainDrain): returns the node it drains to, TERM_BIT set when that node is a terminal
// (self-receiver / marine node / itself when under the sea) rather than another pit.
__device__ inline int pit_walk(const Smem &S, int root, bool use_sea, double sea)
{
    const int *rcv = S.P.i[I_RCV], *pitID = S.P.i[I_PITID];
    const double *W = S.P.f[F_FILLH];
    if (use_sea && W[root] < sea) return root | TERM_BIT;
    int donor = root;
    for (;;) {
        const int r = rcv[donor];
        if (r == donor) return r | TERM_BIT;
        if (use_sea && W[r] < sea) return r | TERM_BIT;
        const int pid = pitID[r];
        if (pid == -1 || pid == root) donor = r;
        else if (use_sea && W[pid] < sea) donor = r;
        else return pid;
    }
}

// literal sequential FLOWalgo.f90:167-238 / :240-297 (used only when the drainage graph has a cycle)
__device__ void drain_sequential(const DevConst &c, Smem &S, const int *pids, int np, const int *order,
                                 int *drain, int *chain, bool use_sea, double sea)
{
    const int *rcv = S.P.i[I_RCV], *pitID = S.P.i[I_PITID];
    const double *W = S.P.f[F_FILLH];
    for (int k = 0; k < np; ++k) chain[k] = -1;
    for (int n = 0; n < np; ++n) {
        int nID = pids[order[n]];
        int donor = nID;
        bool exist = drain[nID] > -1;
        if (use_sea && W[nID] < sea) { exist = true; drain[nID] = nID; }
        while (!exist) {
            const int r = rcv[donor];
            if (r == donor || (use_sea && W[r] < sea)) {
                drain[nID] = r;
                for (int k = 0; k < np; ++k) chain[k] = -1;
                exist = true;
            } else if (pitID[r] == -1 || pitID[r] == nID) donor = r;
            else if (use_sea && W[pitID[r]] < sea) donor = r;
            else {
                int q = 0, nd = pitID[r];
                bool newpit = true;
                while (q < np && chain[q] >= 0) { if (chain[q] == nd) newpit = false; ++q; }
                if (newpit && q < np) { chain[q] = nd; drain[nID] = nd; donor = nd; nID = nd; }
                else donor = r;
            }
        }
    }
}

// phase: pit drainage.  flowNetwork.compute_parameters_depression (flowNetwork.py:558-577) ->
// basindrainage / basindrainageall.  Each basin root walks independently (parallel); when the resulting
// pit -> pit graph is acyclic the sequential chain logic of the reference assigns exactly these targets to
// every pit reachable from a pit with volume >= 0, in any processing order.  A cycle (the case chainDrain
// exists for) falls back to the literal sequential routine.
__device__ void phase_drainage(const DevConst &c, Smem &S, const int *base1, int nbase1, double sea)
{
    const int N = c.N, tid = threadIdx.x;
    int *succ = S.P.i[I_SUCC], *succ2 = S.P.i[I_SUCC2], *drain = S.P.i[I_PITDRAIN], *alld = S.P.i[I_ALLDRAIN];
    int *mark = S.P.i[I_MARK];
    const double *pitVol = S.P.f[F_PITVOL];
    const bool marine = (c.nsea > 0) || true;
    for (int v = tid; v < N; v += TPB) { drain[v] = -1; alld[v] = -1; mark[v] = 0; }
    __syncthreads();
    for (int i = tid; i < nbase1; i += TPB) {
        const int root = base1[i];
        succ[root] = pit_walk(S, root, true, sea);
        succ2[root] = marine ? pit_walk(S, root, false, sea) : succ[root];
    }
    __syncthreads();
    // cycle test from every pit with volume >= 0
    int cyc = 0;
    for (int i = tid; i < nbase1; i += TPB) {
        const int p = base1[i];
        if (pitVol[p] >= 0.) {
            int q = p, steps = 0;
            while (!(succ[q] & TERM_BIT)) { q = succ[q]; if (++steps > nbase1) { cyc |= 1; break; } }
            q = p; steps = 0;
            while (!(succ2[q] & TERM_BIT)) { q = succ2[q]; if (++steps > nbase1) { cyc |= 2; break; } }
        }
    }
    cyc = __syncthreads_or(cyc);
    if (!cyc) {
        for (int i = tid; i < nbase1; i += TPB) {
            const int p = base1[i];
            if (pitVol[p] >= 0.) {
                int q = p;
                for (;;) {
                    if (atomicOr(&mark[q], 1) & 1) break;
                    const int s = succ[q];
                    drain[q] = s & ~TERM_BIT;
                    if (s & TERM_BIT) break;
                    q = s;
                }
                q = p;
                for (;;) {
                    if (atomicOr(&mark[q], 2) & 2) break;
                    const int s = succ2[q];
                    alld[q] = s & ~TERM_BIT;
                    if (s & TERM_BIT) break;
                    q = s;
                }
            }
        }
        __syncthreads();
        return;
    }
    // fallback: literal order = argsort(fillH[pIDs])[::-1] with ties taken as stable-ascending reversed
    if (tid == 0) atomicOr(&c.status[blockIdx.x], BL_ST_DRAIN_CYCLE);
    int *pids = S.P.i[I_LB];
    const int np = blk_compact(N, pids, [&](int i) { return pitVol[i] >= 0.; }, S);
    int *order = S.P.i[I_EVENTS], *chain = S.P.i[I_MEMB], *chain2 = S.P.i[I_KIND];
    const double *W = S.P.f[F_FILLH];
    for (int i = tid; i < np; i += TPB) {
        const double ki = W[pids[i]];
        int rank = 0;
        for (int j = 0; j < np; ++j) {
            const double kj = W[pids[j]];
            if (kj > ki || (kj == ki && j > i)) ++rank;
        }
        order[rank] = i;
    }
    __syncthreads();
    if (tid == 0) drain_sequential(c, S, pids, np, order, drain, chain, true, sea);
    if (tid == 32 % TPB) drain_sequential(c, S, pids, np, order, alld, chain2, false, sea);
    __syncthreads();
}

// phase: discharge.  FLOWalgo.f90:53-81 walks the stack backwards adding each donor into its receiver, so
// a receiver sums  own + Q[donor_k] + ... + Q[donor_1]  with donors in DESCENDING id, each donor's subtree
// already complete.  Done here as a dataflow pass: a node is evaluated by the thread that delivers its last
// donor (atomic counter), which then sums the donors in that fixed order -> bit-identical fp64 results.
__device__ void phase_discharge(const DevConst &c, Smem &S, double rain)
{
    const int N = c.N, tid = threadIdx.x;
    const int *rcv = S.P.i[I_RCV], *lc = S.P.i[I_LC], *ps = S.P.i[I_PS];
    int *cnt = S.P.i[I_MARK];
    const int *cnt0 = S.P.i[I_CNT];
    double *Q = S.P.f[F_Q];
    for (int v = tid; v < N; v += TPB) cnt[v] = cnt0[v];
    __syncthreads();
    for (int v = tid; v < N; v += TPB) {
        if (cnt0[v] != 0) continue;
        int cur = v;
        for (;;) {
            double acc = c.area[cur] * ((cur >= c.bounds) ? rain : 0.);
            for (int d = lc[cur]; d != NONE_I; d = ps[d]) acc = acc + ldv(&Q[d]);
            stv(&Q[cur], acc);
            const int r = rcv[cur];
            if (r == cur) break;
            __threadfence_block();
            if (atomicSub(&cnt[r], 1) != 1) break;
            cur = r;
        }
    }
    __syncthreads();
}

// phase: flow CFL.  FLOWalgo.f90:299-330 evaluated on the FILLED surface (buildFlux.py:169)
__device__ double phase_cfl(const DevConst &c, Smem &S, double erod)
{
    const int N = c.N;
    const int *rcv = S.P.i[I_RCV];
    const double *W = S.P.f[F_FILLH], *Q = S.P.f[F_Q];
    double cfl = 1.e6;
    for (int d = threadIdx.x; d < N; d += TPB) {
        const int r = rcv[d];
        const double dz = W[d] - W[r];
        if (d != r && dz > 0. && Q[d] > 0.) {
            const double dx = c.xy[2 * d] - c.xy[2 * r], dy = c.xy[2 * d + 1] - c.xy[2 * r + 1];
            const double dist = sqrt(dx * dx + dy * dy);
            const double tmp = dist / (erod * powx(Q[d], c.m) * powx(dz / dist, c.n - 1.));
            cfl = fmin(tmp, cfl);
        }
    }
    return blk_min(cfl, S);
}

__device__ inline double py2_floorish(double x) { return (x > 1.) ? round(x - 0.5) : x; }

__device__ void phase_streampower_serial(const DevConst &c, Smem &S, double erod, double sea, double dt);

// phase: stream power law.  FLOWalgo.f90:583-1044, single rock, detachment-limited.
//  (1) per-node pre-pass: incision rate and the deposition rule that applies (parallel);
//  (2) sediment routing  sedFluxes(recvr) += Qs  as the same ordered dataflow pass as the discharge;
//  (3) the pit cascade (:967-1034) replayed sequentially over the lake-entry events in reverse stack order.
__device__ void phase_streampower(const DevConst &c, Smem &S, double erod, double sea, double dt)
{
    const int N = c.N, tid = threadIdx.x;
    const int *rcv = S.P.i[I_RCV], *lc = S.P.i[I_LC], *ps = S.P.i[I_PS], *pitID = S.P.i[I_PITID];
    const int *cnt0 = S.P.i[I_CNT], *stack = S.P.i[I_STACK], *pitDrain = S.P.i[I_PITDRAIN];
    int *cnt = S.P.i[I_MARK], *kind = S.P.i[I_KIND], *events = S.P.i[I_EVENTS];
    const double *z = S.P.f[F_Z], *W = S.P.f[F_FILLH], *Q = S.P.f[F_Q], *maxhA = S.P.f[F_MAXH];
    const double *pitVol1 = S.P.f[F_PITVOL];
    double *E = S.P.f[F_E], *caph = S.P.f[F_CAPH], *sedin = S.P.f[F_SEDIN], *qs = S.P.f[F_QS];
    double *depo = S.P.f[F_DEPO], *ero = S.P.f[F_ERO], *pdep = S.P.f[F_PDEP], *pitVol = S.P.f[F_PITVOLW];
    // (1)
    for (int d = tid; d < N; d += TPB) {
        const int r = rcv[d];
        const double zd = z[d], zr = z[r], fh = W[d], area = c.area[d];
        double dh = 0.95 * (zd - zr);
        if (zd > sea && zr < sea) dh = 0.99 * (zd - sea);
        if (dh < 0.001) dh = 0.;
        const double waterH = fh - zd;
        const double dx = c.xy[2 * d] - c.xy[2 * r], dy = c.xy[2 * d + 1] - c.xy[2 * r + 1];
        const double dist = sqrt(dx * dx + dy * dy);
        double SPL = 0., totspl = 0.;
        if (r != d && dh > 0. && waterH == 0. && fh >= sea && dist > 0.) {
            const double slp = dh / dist;
            SPL = -erod * 1. * 1. * powx(Q[d], c.m) * powx(slp, c.n);
            totspl = totspl + SPL;
            if (-totspl * dt > dh) {
                const double frac = dh / (-totspl * dt);
                SPL = SPL * frac;
                totspl = -dh;
            }
        }
        double maxh = maxhA[d];
        if (waterH > 0.) maxh = waterH;
        else if (zd < sea) maxh = sea - zd;
        maxh = 0.95 * maxh;
        int k;
        double e = 0.;
        if (totspl < 0.) { k = 0; e = SPL * dt * area; }
        else if (totspl >= 0. && area > 0.) {
            if (waterH > 0. && fh > sea) {
                k = 1;
                // first cascade iteration is static when the node's pit lies under the sea (:979-991): the load
                // settles on the node if it is under water itself, else it is passed on to the receiver
                const int pp = pitID[d];
                if (pp >= 0 && c.area[pp] > 0. && W[pp] < sea) k = (zd < sea) ? 2 : 4;
            }
            else if (zd <= sea) k = 2;
            else if (maxh > 0. && waterH == 0. && d != r && zd > sea) k = 3;
            else if (d == r) k = 2;
            else k = 4;
        } else k = 5;
        kind[d] = k;
        E[d] = e;
        caph[d] = maxh;
        cnt[d] = cnt0[d];
        depo[d] = 0.;
        ero[d] = 0.;
        pdep[d] = 0.;
        pitVol[d] = pitVol1[d];
    }
    __syncthreads();
    // (2)
    for (int v = tid; v < N; v += TPB) {
        if (cnt0[v] != 0) continue;
        int cur = v;
        for (;;) {
            double acc = 0. * dt;
            for (int d = lc[cur]; d != NONE_I; d = ps[d]) acc = acc + ldv(&qs[d]);
            sedin[cur] = acc;
            const int k = kind[cur];
            const double area = c.area[cur];
            double Qs = 0., erodep = 0.;
            if (k == 0) { erodep = E[cur]; Qs = -erodep + acc; }
            else if (k == 1) { pdep[cur] = acc; }
            else if (k == 2) { erodep = acc; }
            else if (k == 3) {
                const double totflx = 0. + acc, maxh = caph[cur];
                if (totflx / area < maxh) erodep = acc;
                else { const double frac = acc / totflx; erodep = frac * maxh * area; Qs = acc - erodep; }
            } else if (k == 4) Qs = acc;
            if (!(k == 1 && acc != 0.)) {
                if (erodep < 0.) ero[cur] = ero[cur] + erodep;
                else depo[cur] = depo[cur] + erodep;
            }
            stv(&qs[cur], Qs);
            const int r = rcv[cur];
            if (r == cur) break;
            __threadfence_block();
            if (atomicSub(&cnt[r], 1) != 1) break;
            cur = r;
        }
    }
    __syncthreads();
    // (3) lake-entry events in reverse stack order
    const int nev = blk_compact(N, events, [&](int i) { const int v = stack[N - 1 - i]; return kind[v] == 1 && pdep[v] > 0.; }, S);
    if (tid == 0) S.ibuf[0] = 0;
    __syncthreads();
    if (tid == 0 && nev > 0) {
        unsigned st = 0;
        for (int e = 0; e < nev; ++e) {
            const int donor = stack[N - 1 - events[e]];
            const int recvr = rcv[donor];
            double pitDep = pdep[donor];
            double totdist = 0. + pitDep;
            if (!(pitID[donor] >= 0 && c.area[pitID[donor]] > 0.)) continue;
            int tmpID = pitID[donor], nID = tmpID;
            long guard = 0;
            while (totdist > 0.) {
                if (++guard > 1000000 || tmpID < 0) { st |= BL_ST_CASCADE_GUARD; break; }
                const double tmpdist = 0. + depo[tmpID];
                if (W[tmpID] < sea) {
                    // an overflow reaching the sea is handed back to the routing (sedFluxes(recvr) += pitDep):
                    // redo the whole pass with the literal serial loop
                    if (z[donor] < sea) depo[donor] = depo[donor] + pitDep;
                    else { S.ibuf[0] = 1; e = nev; }
                    totdist = 0.;
                    nID = recvr;
                } else if (tmpdist + totdist <= pitVol[tmpID]) {
                    depo[tmpID] = depo[tmpID] + pitDep; totdist = 0.; nID = tmpID;
                } else if (pitDrain[tmpID] == tmpID) {
                    depo[tmpID] = depo[tmpID] + pitDep; totdist = 0.; nID = tmpID;
                } else {
                    if (!(c.flags[tmpID] & 1)) { totdist = 0.; nID = tmpID; }
                    else if (tmpdist == pitVol[tmpID]) nID = tmpID;
                    else {
                        const double frac = pitDep / totdist;
                        depo[tmpID] = depo[tmpID] + (pitVol[tmpID] - tmpdist) * frac;
                        pitDep = (totdist - (pitVol[tmpID] - tmpdist)) * frac;
                        const double newdist = 0. + pitDep;
                        const double totflx = 0. + depo[tmpID];
                        totdist = newdist;
                        pitVol[tmpID] = totflx;
                        nID = tmpID;
                    }
                }
                tmpID = pitDrain[nID];
            }
        }
        if (st) atomicOr(&c.status[blockIdx.x], st);
    }
    __syncthreads();
    if (S.ibuf[0]) phase_streampower_serial(c, S, erod, sea, dt);
}

// numpy pairwise summation order (numpy/_core/src/umath/loops_utils.h.src) for flowNetwork.py:759
__device__ double pairwise_sum(const double *a, int n)
{
    // iterative post-order evaluation of the recursion n -> (n2 = n/2 - (n/2)%8, n - n2), leaves <= 128
    double vstk[40];
    int ostk[40], nstk[40], sstk[40];
    int sp = 0, vp = 0;
    ostk[0] = 0; nstk[0] = n; sstk[0] = 0; sp = 1;
    while (sp > 0) {
        const int o = ostk[sp - 1], m = nstk[sp - 1], st = sstk[sp - 1];
        if (m <= 128) {
            double res;
            if (m < 8) {
                res = 0.;
                for (int i = 0; i < m; ++i) res += a[o + i];
            } else {
                double r[8];
                int i;
                for (int k = 0; k < 8; ++k) r[k] = a[o + k];
                for (i = 8; i < m - (m % 8); i += 8)
                    for (int k = 0; k < 8; ++k) r[k] += a[o + i + k];
                res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
                for (; i < m; ++i) res += a[o + i];
            }
            vstk[vp++] = res;
            --sp;
        } else if (st == 0) {
            int n2 = m / 2;
            n2 -= n2 % 8;
            sstk[sp - 1] = 1;
            ostk[sp] = o; nstk[sp] = n2; sstk[sp] = 0; ++sp;
        } else if (st == 1) {
            int n2 = m / 2;
            n2 -= n2 % 8;
            sstk[sp - 1] = 2;
            ostk[sp] = o + n2; nstk[sp] = m - n2; sstk[sp] = 0; ++sp;
        } else {
            const double b = vstk[--vp], a0 = vstk[--vp];
            vstk[vp++] = a0 + b;
            --sp;
        }
    }
    return vstk[0];
}

// ordered compaction of a bitmap: out[] receives the indices of the set bits of bm[0 .. nbits) in ascending order
__device__ inline int blk_compact_bits(const unsigned *bm, int nbits, int *out, Smem &S)
{
    const int nw = (nbits + 31) >> 5;
    const int wpt = (nw + TPB - 1) / TPB;
    const int w0 = min(nw, (int)threadIdx.x * wpt), w1 = min(nw, w0 + wpt);
    int cnt = 0;
    for (int w = w0; w < w1; ++w) cnt += __popc(bm[w]);
    int total;
    int off = blk_exscan(cnt, total, S);
    for (int w = w0; w < w1; ++w) {
        unsigned bits = bm[w];
        while (bits) { const int b = __ffs(bits) - 1; bits &= bits - 1; out[off++] = w * 32 + b; }
    }
    __syncthreads();
    return total;
}

// numpy's pairwise summation (see pairwise_sum) evaluated by the whole block: thread 0 lists the <= 128-element
// leaf blocks of the recursion, one thread sums each leaf with numpy's 8-accumulator loop, thread 0 folds the
// leaf sums back up the same recursion tree.  Bit-identical to the serial routine; result returned to all threads.
__device__ inline double pw_leaf(const double *a, int m)
{
    double res;
    if (m < 8) {
        res = 0.;
        for (int i = 0; i < m; ++i) res += a[i];
    } else {
        double r[8];
        int i;
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = a[k];
        for (i = 8; i < m - (m % 8); i += 8) {
#pragma unroll
            for (int k = 0; k < 8; ++k) r[k] += a[i + k];
        }
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < m; ++i) res += a[i];
    }
    return res;
}
// `stk`: 160 ints + 40 doubles of scratch for thread 0's recursion stacks (shared memory when the caller has some: as
// local arrays they are dynamically indexed, i.e. live in local memory, and the two serial walks of thread 0 dominate
// the routine); nullptr = local arrays.
__device__ double blk_pairwise_sum(const double *a, int n, Smem &S, int *loff, int *llen, double *lsum, int *stk = nullptr)
{
    if (stk) {
        int *ostk = stk, *nstk = stk + 40, *sstk = stk + 80;
        double *vstk = reinterpret_cast<double *>(stk + 120);
        if (threadIdx.x == 0) {
            int sp = 0, nl = 0;
            ostk[0] = 0; nstk[0] = n; sp = 1;
            while (sp > 0) {                       // leaves left to right
                const int o = ostk[sp - 1], m = nstk[sp - 1];
                --sp;
                if (m <= 128) { loff[nl] = o; llen[nl] = m; ++nl; }
                else {
                    int n2 = m / 2;
                    n2 -= n2 % 8;
                    ostk[sp] = o + n2; nstk[sp] = m - n2; ++sp;
                    ostk[sp] = o; nstk[sp] = n2; ++sp;
                }
            }
            S.ibuf[4] = nl;
        }
        __syncthreads();
        const int nl = S.ibuf[4];
        for (int t = threadIdx.x; t < nl; t += TPB) lsum[t] = pw_leaf(a + loff[t], llen[t]);
        __syncthreads();
        if (threadIdx.x == 0) {
            int sp = 0, vp = 0, li = 0;
            nstk[0] = n; sstk[0] = 0; sp = 1;
            while (sp > 0) {
                const int m = nstk[sp - 1], st = sstk[sp - 1];
                if (m <= 128) { vstk[vp++] = lsum[li++]; --sp; }
                else if (st == 0) { int n2 = m / 2; n2 -= n2 % 8; sstk[sp - 1] = 1; nstk[sp] = n2; sstk[sp] = 0; ++sp; }
                else if (st == 1) { int n2 = m / 2; n2 -= n2 % 8; sstk[sp - 1] = 2; nstk[sp] = m - n2; sstk[sp] = 0; ++sp; }
                else { const double b = vstk[--vp], a0 = vstk[--vp]; vstk[vp++] = a0 + b; --sp; }
            }
            S.dbuf[1] = vstk[0];
        }
        __syncthreads();
        return S.dbuf[1];
    }
    if (threadIdx.x == 0) {
        int ostk[40], nstk[40], sp = 0, nl = 0;
        ostk[0] = 0; nstk[0] = n; sp = 1;
        while (sp > 0) {                       // leaves left to right
            const int o = ostk[sp - 1], m = nstk[sp - 1];
            --sp;
            if (m <= 128) { loff[nl] = o; llen[nl] = m; ++nl; }
            else {
                int n2 = m / 2;
                n2 -= n2 % 8;
                ostk[sp] = o + n2; nstk[sp] = m - n2; ++sp;     // right pushed first, left popped first
                ostk[sp] = o; nstk[sp] = n2; ++sp;
            }
        }
        S.ibuf[4] = nl;
    }
    __syncthreads();
    const int nl = S.ibuf[4];
    for (int t = threadIdx.x; t < nl; t += TPB) lsum[t] = pw_leaf(a + loff[t], llen[t]);
    __syncthreads();
    if (threadIdx.x == 0) {
        // post-order fold: same recursion, consuming the leaf sums in order
        double vstk[40];
        int nstk[40], sstk[40], sp = 0, vp = 0, li = 0;
        nstk[0] = n; sstk[0] = 0; sp = 1;
        while (sp > 0) {
            const int m = nstk[sp - 1], st = sstk[sp - 1];
            if (m <= 128) { vstk[vp++] = lsum[li++]; --sp; }
            else if (st == 0) { int n2 = m / 2; n2 -= n2 % 8; sstk[sp - 1] = 1; nstk[sp] = n2; sstk[sp] = 0; ++sp; }
            else if (st == 1) { int n2 = m / 2; n2 -= n2 % 8; sstk[sp - 1] = 2; nstk[sp] = m - n2; sstk[sp] = 0; ++sp; }
            else { const double b = vstk[--vp], a0 = vstk[--vp]; vstk[vp++] = a0 + b; --sp; }
        }
        S.dbuf[1] = vstk[0];
    }
    __syncthreads();
    return S.dbuf[1];
}

// PDalgo.f90:136-269 marine_distribution (single rock): a strictly sequential "drop the sediment, spill what does
// not fit to the lowest neighbour" walk over the marine deposit nodes, diffnb passes, elevation updated in
// place.  Kept literal on one thread per proposal (parallelism is across proposals); the per-pass reset of the
// sediment array is done by the whole block.  Adds the resulting thickness to dth[].
__device__ __noinline__ void phase_marine_distribution(const DevConst &c, Smem &S, double sea)
{
    __builtin_assume(__isShared(&c)); __builtin_assume(__isShared(&S));
    const int N = c.N, tid = threadIdx.x;
    const double *z = S.P.f[F_Z];
    double *dep = S.P.f[F_SEDIN], *dth = S.P.f[F_DTHICK];
    double *elev = S.P.f[F_VAL], *seadep = S.P.f[F_E];
    int *seaid = S.P.i[I_MEMB];
    const int *isl = S.P.i[I_KIND];
    const int nsea = blk_compact(N, seaid, [&](int i) { return isl[i] == 2; }, S);
    // small meshes (the etopo configs): stage the walk's arrays and the neighbour ids in shared memory
    const unsigned short *nb16 = nullptr;
    const double *areaA = c.area;
    const unsigned char *flagA = c.flags;
    const int KDT = c.KDT;
    if (S.dsm && (size_t)c.Nq * (17 + 2 * KDT) + 64 <= (size_t)S.dsm_bytes) {
        elev = (double *)S.dsm;
        seadep = elev + c.Nq;
        unsigned short *nb = (unsigned short *)(seadep + c.Nq);
        unsigned char *fl = (unsigned char *)(nb + (size_t)c.Nq * KDT);
        for (int i = tid; i < N * KDT; i += TPB) nb[i] = c.ngb16[i];
        for (int i = tid; i < N; i += TPB) fl[i] = c.flags[i];
        nb16 = nb; flagA = fl;
    }
    for (int k = tid; k < N; k += TPB) elev[k] = z[k];
    __syncthreads();
    const double diff_res = 1.e-2;
    const int max_it_cyc = 500000;
    // The walk is sequential, but each step scans one node's neighbours twice: warp 0 runs it with one lane per neighbour
    // slot.  The two scans - highest neighbour, first lowest neighbour - are warp reductions (redux.sync) on order-preserving
    // 64-bit keys of the elevations, high word first; all lanes carry the same scalars, so every branch is uniform; lane 0
    // writes, once per iteration, between two warp barriers (readers of the old values before, readers of the new after).
    const int lane = tid & 31;
    const int KDv = nb16 ? KDT : c.KD;
    auto nbr = [&](int id, int q) -> int {
        if (q >= KDv) return -1;
        if (nb16) { const int j = nb16[id * KDT + q]; return (j == 0xFFFF) ? -1 : j; }
        return c.ngb[(size_t)q * N + id];
    };
    auto key_of = [](double x) -> unsigned long long {      // x < y  <=>  key(x) < key(y)   (-0 counts as +0)
        unsigned long long u = (unsigned long long)__double_as_longlong(x);
        if (u == 0x8000000000000000ull) u = 0ull;
        return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
    };
    auto val_of = [](unsigned long long k) -> double {
        const unsigned long long u = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
        return __longlong_as_double((long long)u);
    };
    for (int mm = 0; mm < c.diffnb; ++mm) {
        for (int k = tid; k < N; k += TPB) seadep[k] = ((isl[k] == 2) ? dep[k] : 0.) / (double)(float)c.diffnb;
        __syncthreads();
        if (tid < 32) {
            for (int k = 0; k < nsea; ++k) {
                int id = seaid[k];
                int it = 0;
                for (;;) {
                    // ---- reads of this iteration (every lane) ----
                    const bool inside = (flagA[id] & 1);
                    const double area = areaA[id];
                    const double eid = elev[id], sd = seadep[id];
                    const int j = nbr(id, lane);
                    const double ej = (j >= 0) ? elev[j] : -1.e8;
                    // highest neighbour (floor -1e8, as the reference's initial value)
                    const unsigned long long kj = key_of(ej);
                    const unsigned khi = (unsigned)(kj >> 32), klo = (unsigned)kj;
                    const unsigned mh = __reduce_max_sync(0xffffffffu, khi);
                    const unsigned ml = __reduce_max_sync(0xffffffffu, (khi == mh) ? klo : 0u);
                    const double maxnb = val_of(((unsigned long long)mh << 32) | ml);
                    // first lowest neighbour (PDalgo.f90:197-209: strictly lower than the running minimum, first slot wins ties)
                    const unsigned nh = __reduce_min_sync(0xffffffffu, (j >= 0) ? khi : 0xffffffffu);
                    const unsigned nl = __reduce_min_sync(0xffffffffu, (j >= 0 && khi == nh) ? klo : 0xffffffffu);
                    const unsigned best = __ballot_sync(0xffffffffu, j >= 0 && khi == nh && klo == nl);
                    const int bq = best ? (__ffs(best) - 1) : 0;
                    const int bj = __shfl_sync(0xffffffffu, j, bq);
                    const double bv = best ? val_of(((unsigned long long)nh << 32) | nl) : 1.e300;
                    double maxz = maxnb;
                    if (maxz > sea) maxz = sea;
                    if (maxz < eid) maxz = eid;
                    const double vol = fmax(0., c.diffprop * (maxz - eid) * area);
                    // ---- decide (uniform) and collect the writes: up to three (address, value) pairs ----
                    double *p0 = nullptr, *p1 = nullptr, *p2 = nullptr;
                    double v0 = 0., v1 = 0., v2 = 0.;
                    bool done = false;
                    int nextid = id;
                    if (!inside) { p0 = &seadep[id]; v0 = 0.; done = true; }
                    else if (it > max_it_cyc || sd / area < diff_res) { p0 = &elev[id]; v0 = eid + sd / area; p1 = &seadep[id]; v1 = 0.; done = true; }
                    else {
                        it = it + 1;
                        if (sd < vol) { p0 = &elev[id]; v0 = eid + sd / area; p1 = &seadep[id]; v1 = 0.; done = true; }
                        else {
                            const double sd2 = sd - vol, e2 = eid + vol / area;
                            if (sd2 > 0.) {
                                int nid = (best && bv < e2) ? bj : -1;
                                if (nid == -1) {
                                    const double dh = maxnb - e2 + 0.1;
                                    if (sd2 > dh * area) {
                                        const double e3 = e2 + dh, sd3 = sd2 - dh * area;
                                        p0 = &elev[id]; v0 = e3;
                                        nid = seaid[k];
                                        it = 0;
                                        if (nid == id) { p1 = &seadep[nid]; v1 = sd3; }
                                        else { const double t = seadep[nid]; p1 = &seadep[nid]; v1 = t + sd3; p2 = &seadep[id]; v2 = 0.; }
                                        nextid = nid;
                                    } else { p0 = &elev[id]; v0 = e2 + sd2 / area; p1 = &seadep[id]; v1 = 0.; done = true; }
                                } else {
                                    const double t = seadep[nid];
                                    p0 = &elev[id]; v0 = e2;
                                    p1 = &seadep[nid]; v1 = t + sd2;
                                    p2 = &seadep[id]; v2 = 0.;
                                    nextid = nid;
                                }
                            } else { p0 = &elev[id]; v0 = e2; p1 = &seadep[id]; v1 = 0.; done = true; }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) {
                        if (p0) *p0 = v0;
                        if (p1) *p1 = v1;
                        if (p2) *p2 = v2;
                    }
                    __syncwarp();
                    if (done) break;
                    id = nextid;
                }
            }
        }
        __syncthreads();
    }
    for (int k = tid; k < N; k += TPB) {
        dth[k] += elev[k] - z[k];
        if (isl[k] == 2) dep[k] = 0.;
    }
    __syncthreads();
}

// FLOWalgo.f90:332-388 diffmarine + the sub-stepping loop of buildFlux.py:198-247 (river-borne marine sediment
// diffusion, criver > 0).  z / cumdiff are updated in place; `sumdep` starts as the step's deposition thickness.
__device__ void phase_marine_diffusion(const DevConst &c, Smem &S, double sea, double timestep)
{
    const int N = c.N, tid = threadIdx.x;
    double *z = S.P.f[F_Z], *cumdiff = S.P.f[F_CUMDIFF];
    double *sumdep = S.P.f[F_DTHICK], *coeff = S.P.f[F_E], *dm = S.P.f[F_CAPH];
    const double maxth = 0.1;
    for (int k = tid; k < N; k += TPB) {                       // diffLinear.sedfluxmarine (diffLinear.py:184-213)
        const double area = c.area[k];
        const double ac = (area > 0.) ? 1. / area : 0.;
        coeff[k] = ac * ((z[k] >= sea) ? 0. : c.CDr);
    }
    __syncthreads();
    double diffstep = timestep;
    int it = 0;
    while (diffstep > 0. && it < 1000) {
        double maxstep = fmin(c.ms, diffstep);
        double mind = maxstep;
        for (int g = tid; g < N; g += TPB) {
            double acc = 0.;
            if ((c.flags[g] & 1) && z[g] < sea) {
                const double zg = z[g], dg = sumdep[g], cg = coeff[g];
                for (int q = 0; q < c.KD; ++q) {
                    const int j = c.ngb[(size_t)q * N + g];
                    if (j < 0) break;
                    const double zj = z[j];
                    if (c.flags[j] & 1) {
                        const double flx = c.vor[(size_t)q * N + g] * (zj - zg) / c.edge[(size_t)q * N + g];
                        if (dg > maxth && zg > zj) acc = acc + cg * flx;
                        else if (sumdep[j] > maxth && zg < zj && zj < sea) acc = acc + cg * flx;
                    } else {
                        if (dg > maxth && zg > zj) {
                            const double flx = c.vor[(size_t)q * N + g] * (zj - zg) / c.edge[(size_t)q * N + g];
                            acc = acc + cg * flx;
                        }
                    }
                }
                if (acc < 0. && acc * maxstep < -dg) mind = fmin(-dg / acc, mind);
            }
            dm[g] = (c.flags[g] & 1) ? acc : 0.;
        }
        mind = blk_min(mind, S);
        maxstep = fmin(mind, maxstep);
        diffstep -= maxstep;
        for (int k = tid; k < N; k += TPB) {
            const double d = dm[k] * maxstep;
            sumdep[k] += d; z[k] += d; cumdiff[k] += d;
        }
        __syncthreads();
        ++it;
    }
}

// FLOWalgo.f90:583-1044, literal reverse-stack loop on one thread.  Used only when the pit cascade hands sediment
// back to the routing (an overflowing lake draining into the sea: sedFluxes(recvr) += pitDep, :986-989), which
// couples the cascade state to the ordered routing and leaves no parallel shortcut.
__device__ void streampower_serial(const DevConst &c, Smem &S, const int *rcv, double erod, double sea, double dt,
                                   unsigned &st)
{
    const int N = c.N;
    const int *stack = S.P.i[I_STACK], *pitID = S.P.i[I_PITID], *pitDrain = S.P.i[I_PITDRAIN];
    const double *z = S.P.f[F_Z], *W = S.P.f[F_FILLH], *Q = S.P.f[F_Q], *maxhA = S.P.f[F_MAXH];
    double *depo = S.P.f[F_DEPO], *ero = S.P.f[F_ERO], *sedfl = S.P.f[F_SEDIN], *pitVol = S.P.f[F_PITVOLW];
    for (int i = N - 1; i >= 0; --i) {
        const int donor = stack[i], recvr = rcv[donor];
        const double zd = z[donor], zr = z[recvr], fh = W[donor], area = c.area[donor];
        double SPL = 0., totspl = 0.;
        double dh = 0.95 * (zd - zr);
        if (zd > sea && zr < sea) dh = 0.99 * (zd - sea);
        if (dh < 0.001) dh = 0.;
        const double waterH = fh - zd;
        double totdist = 0.;
        if (recvr != donor && dh > 0. && waterH == 0. && fh >= sea) {
            const double dx = c.xy[2 * donor] - c.xy[2 * recvr], dy = c.xy[2 * donor + 1] - c.xy[2 * recvr + 1];
            const double dist = sqrt(dx * dx + dy * dy);
            if (dist > 0.) {
                const double slp = dh / dist;
                SPL = -erod * 1. * 1. * powx(Q[donor], c.m) * powx(slp, c.n);
                totspl = totspl + SPL;
                if (-totspl * dt > dh) { const double frac = dh / (-totspl * dt); SPL = SPL * frac; totspl = -dh; }
            }
        }
        double maxh = maxhA[donor];
        if (waterH > 0.) maxh = waterH;
        else if (zd < sea) maxh = sea - zd;
        maxh = 0.95 * maxh;
        double Qs = 0., erodep = 0., pitDep = 0.;
        if (totspl < 0.) { erodep = SPL * dt * area; Qs = -erodep + sedfl[donor]; }
        else if (totspl >= 0. && area > 0.) {
            if (waterH > 0. && fh > sea) { pitDep = sedfl[donor]; totdist = totdist + pitDep; }
            else if (zd <= sea) erodep = sedfl[donor];
            else if (maxh > 0. && waterH == 0. && donor != recvr && zd > sea) {
                const double totflx = 0. + sedfl[donor];
                if (totflx / area < maxh) erodep = sedfl[donor];
                else { const double frac = sedfl[donor] / totflx; erodep = frac * maxh * area; Qs = sedfl[donor] - erodep; }
            } else if (donor == recvr) erodep = sedfl[donor];
            else Qs = sedfl[donor];
        }
        if (pitDep == 0.) {
            sedfl[recvr] = sedfl[recvr] + Qs;
            if (erodep < 0.) ero[donor] = ero[donor] + erodep;
            else depo[donor] = depo[donor] + erodep;
        } else if (pitDep > 0. && pitID[donor] >= 0 && c.area[pitID[donor]] > 0.) {
            int tmpID = pitID[donor], nID = tmpID;
            long guard = 0;
            while (totdist > 0.) {
                if (++guard > 1000000 || tmpID < 0) { st |= BL_ST_CASCADE_GUARD; break; }
                const double tmpdist = 0. + depo[tmpID];
                if (W[tmpID] < sea) {
                    if (zd < sea) depo[donor] = depo[donor] + pitDep;
                    else sedfl[recvr] = sedfl[recvr] + pitDep;
                    totdist = 0.; nID = recvr;
                } else if (tmpdist + totdist <= pitVol[tmpID]) { depo[tmpID] = depo[tmpID] + pitDep; totdist = 0.; nID = tmpID; }
                else if (pitDrain[tmpID] == tmpID) { depo[tmpID] = depo[tmpID] + pitDep; totdist = 0.; nID = tmpID; }
                else {
                    if (!(c.flags[tmpID] & 1)) { totdist = 0.; nID = tmpID; }
                    else if (tmpdist == pitVol[tmpID]) nID = tmpID;
                    else {
                        const double frac = pitDep / totdist;
                        depo[tmpID] = depo[tmpID] + (pitVol[tmpID] - tmpdist) * frac;
                        pitDep = (totdist - (pitVol[tmpID] - tmpdist)) * frac;
                        const double newdist = 0. + pitDep;
                        const double totflx = 0. + depo[tmpID];
                        totdist = newdist;
                        pitVol[tmpID] = totflx;
                        nID = tmpID;
                    }
                }
                tmpID = pitDrain[nID];
            }
        }
    }
}

// block wrapper: reset the outputs, then run the literal loop on thread 0
__device__ void phase_streampower_serial(const DevConst &c, Smem &S, double erod, double sea, double dt)
{
    const int N = c.N, tid = threadIdx.x;
    double *depo = S.P.f[F_DEPO], *ero = S.P.f[F_ERO], *sedfl = S.P.f[F_SEDIN], *pitVol = S.P.f[F_PITVOLW];
    const double *pitVol1 = S.P.f[F_PITVOL];
    for (int k = tid; k < N; k += TPB) { depo[k] = 0.; ero[k] = 0.; sedfl[k] = 0. * dt; pitVol[k] = pitVol1[k]; }
    __syncthreads();
    if (tid == 0) {
        unsigned st = 0;
        streampower_serial(c, S, S.P.i[I_RCV], erod, sea, dt, st);
        if (st) atomicOr(&c.status[blockIdx.x], st);
    }
    __syncthreads();
}

// phase: erosion / deposition thickness.  flowNetwork.compute_sedflux (flowNetwork.py:721-793) with getids
// (FLOWalgo.f90:1089-1180) fused.  Result sedflux[] (= ed) in metres.
__device__ __noinline__ void phase_erodep(const DevConst &c, Smem &S, double sea)
{
    __builtin_assume(__isShared(&c)); __builtin_assume(__isShared(&S));
    const int N = c.N, tid = threadIdx.x;
    // distinct per-proposal arrays: __restrict__ lets the streaming loops below keep several nodes' loads in flight
    const double *__restrict__ z = S.P.f[F_Z], *__restrict__ W = S.P.f[F_FILLH], *__restrict__ pitVol = S.P.f[F_PITVOL];
    const double *__restrict__ cdepo = S.P.f[F_DEPO], *__restrict__ cero = S.P.f[F_ERO];
    const int *pitID = S.P.i[I_PITID], *pos1 = S.P.i[I_POS1], *ns1 = S.P.i[I_NS], *stack1 = S.P.i[I_STACK1];
    double *__restrict__ dep = S.P.f[F_SEDIN], *__restrict__ dth = S.P.f[F_DTHICK], *tmpv = S.P.f[F_VAL];
    double *__restrict__ sedflux = S.P.f[F_SEDFLUX];
    int *__restrict__ kindA = S.P.i[I_KIND];
    const unsigned char *__restrict__ flagsA = c.flags;
    const double *__restrict__ areaA = c.area;
    int *land = S.P.i[I_LB], *memb = S.P.i[I_MEMB];
    int marine = 0, left = 0;
    SubTick st = sub_tick_begin(c);
    if (tid == 0) S.ibuf[6] = 0;             // number of land pits found by the first pass
    __syncthreads();
    // every load of a node is issued before the first branch on it, so the unrolled iterations keep one batch of
    // independent loads in flight (a load under a data-dependent branch costs a memory round trip per branch)
    if (c.depo != 0) {
#pragma unroll 4
        for (int k = tid; k < N; k += TPB) {
            const unsigned char fl = flagsA[k];
            const double cdk = cdepo[k], zk = z[k], fh = W[k], vol = pitVol[k], ar = areaA[k];
            if (k + 4 * TPB < N) { pf_l2(cdepo + k + 4 * TPB); pf_l2(z + k + 4 * TPB); pf_l2(W + k + 4 * TPB); pf_l2(pitVol + k + 4 * TPB); }
            double d = (fl & 1) ? cdk : 0.;
            double th = 0.;
            int island = 0;
            int in_ = 0;
            double sumdep = 0.;
            bool plain = false;
            if (zk > sea && fh == zk) { in_ = 1; sumdep = sumdep + d; plain = sumdep > 0.; }
            if (zk > sea && fh > sea && vol > 0.) {
                if (sumdep == 0. && in_ == 0) {
                    if (d / vol > 1.e-8) island = 1;   // perc = d / d = 1, sumperc = 1
                    else d = 0.;
                }
            }
            if (zk <= sea) { sumdep = sumdep + d; if (sumdep > 0.) { marine = 1; island = 2; } }
            if (plain) { th = d / ar; d = 0.; }
            dep[k] = d;
            dth[k] = th;
            kindA[k] = island;
            if (island == 1) land[atomicAdd(&S.ibuf[6], 1)] = k;     // a handful per step; their order does not matter:
                                                                     // every land pit scales its own members only
            left |= (d != 0. && island != 1);          // land pits hand their load on below: what else is left over?
        }
    } else {
#pragma unroll 4
        for (int k = tid; k < N; k += TPB) {
            const unsigned char fl = flagsA[k];
            const double cdk = cdepo[k];
            dep[k] = (fl & 1) ? cdk : 0.;
            dth[k] = 0.;
            kindA[k] = 0;
        }
    }
    {
        const int both = __syncthreads_or(marine | (left << 1));
        marine = both & 1;
        left = (both >> 1) & 1;
    }
    int any = 0;
    if (c.depo != 0) {
        const int nland = S.ibuf[6];         // (the barrier of the reduction above orders the appends)
        // land pits (flowNetwork.py:750-763): members = where(pitID == p) in ascending id; the basin's nodes
        // are a contiguous segment of the real-surface stack.  Small basins: one thread gathers + sorts its
        // members; large basins: the block compacts them in id order.
        constexpr int SMALL = 96;
        sub_tick(st, 16);
        sub_count(st, 21, nland);
        if (S.dsm) {
            // one WARP per small land pit, on scratch in the (idle) shared memory: the stack segment is read 32 positions at
            // a time (two dependent global loads per position, now in parallel), the members are ordered by rank counting,
            // lane 0 evaluates numpy's pairwise sum on the shared-memory terms.  Same members, same order, same arithmetic.
            const int lane = tid & 31, warp = tid >> 5;
            int *ids = reinterpret_cast<int *>(S.dsm + (size_t)warp * 1536), *srt = ids + SMALL;
            double *tvs = reinterpret_cast<double *>(S.dsm + (size_t)warp * 1536 + 768);
            const unsigned ltm = (1u << lane) - 1u;
            for (int q = warp; q < nland; q += NWARP) {
                const int pit = land[q];
                const int lo = pos1[pit];
                const int nr = ns1[pit];
                const int hi = (nr != NONE_I) ? pos1[nr] : N;
                if (hi - lo > SMALL) continue;
                int nm = 0;
                for (int i0 = lo; i0 < hi; i0 += 32) {
                    const int i = i0 + lane;
                    const int v = (i < hi) ? stack1[i] : -1;
                    const bool isM = (v >= 0) && (pitID[v] == pit);
                    const unsigned mk = __ballot_sync(0xffffffffu, isM);
                    if (isM) ids[nm + __popc(mk & ltm)] = v;
                    nm += __popc(mk);
                }
                __syncwarp();
                for (int e = lane; e < nm; e += 32) {
                    const int v = ids[e];
                    int r = 0;
                    for (int j = 0; j < nm; ++j) r += (ids[j] < v);
                    srt[r] = v;
                }
                __syncwarp();
                for (int e = lane; e < nm; e += 32) {
                    const int v = srt[e];
                    const double dd = (W[v] - z[v]) * 1.0;
                    tvs[e] = dd * c.area[v];
                }
                __syncwarp();
                double tmpd = 0.;
                if (lane == 0) tmpd = pairwise_sum(tvs, nm);
                tmpd = __shfl_sync(0xffffffffu, tmpd, 0);
                const double dfrac = (0. + dep[pit]) / tmpd;
                for (int e = lane; e < nm; e += 32) {
                    const int v = srt[e];
                    const double dd = (W[v] - z[v]) * 1.0;
                    dth[v] = dd * dfrac;
                }
                __syncwarp();
            }
        } else
        for (int q = tid; q < nland; q += TPB) {
            const int pit = land[q];
            const int lo = pos1[pit];
            const int nr = ns1[pit];
            const int hi = (nr != NONE_I) ? pos1[nr] : N;
            if (hi - lo > SMALL) continue;
            int nm = 0;
            int *mb = memb + lo;
            for (int i = lo; i < hi; ++i) {
                const int v = stack1[i];
                if (pitID[v] == pit) {
                    int j = nm++;
                    while (j > 0 && mb[j - 1] > v) { mb[j] = mb[j - 1]; --j; }
                    mb[j] = v;
                }
            }
            double *tv = tmpv + lo;
            for (int i = 0; i < nm; ++i) {
                const int v = mb[i];
                const double dd = (W[v] - z[v]) * 1.0;
                dth[v] = dd;
                tv[i] = dd * c.area[v];
            }
            const double tmpd = pairwise_sum(tv, nm);
            const double dfrac = (0. + dep[pit]) / tmpd;
            for (int i = 0; i < nm; ++i) dth[mb[i]] *= dfrac;
        }
        __syncthreads();
        sub_tick(st, 17);
        for (int q = 0; q < nland; ++q) {
            const int pit = land[q];
            const int lo = pos1[pit];
            const int nr = ns1[pit];
            const int hi = (nr != NONE_I) ? pos1[nr] : N;
            if (hi - lo <= SMALL) continue;
            sub_count(st, 20, 1);
            // members in ascending id: mark them (they all lie in the basin's stack segment), compact the bitmap
            // (shared memory is idle in this phase: the bitmap sits behind the first 8 bytes per node, the terms of the
            // pairwise sum in front; without dynamic shared memory both stay in global scratch)
            unsigned *bm = S.dsm ? (unsigned *)(S.dsm + (size_t)8 * c.Nq) : (unsigned *)S.P.i[I_MARK];
            double *tv = S.dsm ? (double *)S.dsm : tmpv;
            for (int w = tid; w < (N + 31) / 32; w += TPB) bm[w] = 0u;
            __syncthreads();
            for (int i = lo + tid; i < hi; i += TPB) {
                const int v = stack1[i];
                if (pitID[v] == pit) atomicOr(&bm[v >> 5], 1u << (v & 31));
            }
            __syncthreads();
            const int nm = blk_compact_bits(bm, N, memb, S);
            for (int i = tid; i < nm; i += TPB) {
                const int v = memb[i];
                const double dd = (W[v] - z[v]) * 1.0;
                dth[v] = dd;
                tv[i] = dd * c.area[v];
            }
            __syncthreads();
            double tmpd;
            if (S.dsm && nm <= 11000 && (size_t)8 * c.Nq + c.Nq / 8 + 4160 <= (size_t)S.dsm_bytes) {   // <= 200 leaves: offsets, lengths, sums, thread 0's stacks behind the bitmap
                int *sc = reinterpret_cast<int *>(S.dsm + (size_t)8 * c.Nq + c.Nq / 8);
                tmpd = blk_pairwise_sum(tv, nm, S, sc, sc + 200, reinterpret_cast<double *>(sc + 400), sc + 800);
            } else tmpd = blk_pairwise_sum(tv, nm, S, S.P.i[I_EVENTS], S.P.i[I_SUCC], S.P.f[F_CAPH]);
            if (tid == 0) S.dbuf[0] = (0. + dep[pit]) / tmpd;
            __syncthreads();
            const double dfrac = S.dbuf[0];
            for (int i = tid; i < nm; i += TPB) dth[memb[i]] *= dfrac;
            __syncthreads();
        }
        sub_tick(st, 18);
        for (int q = tid; q < nland; q += TPB) dep[land[q]] = 0.;
        __syncthreads();
        // forced deposition of anything left (flowNetwork.py:782-785).  Without marine deposition the only writes to
        // dep since the first pass are the zeroed land pits, so "anything left" is already known.
        any = left;
        if (marine) {
            phase_marine_distribution(c, S, sea);     // flowNetwork.py:769-776
            any = 0;
#pragma unroll 4
            for (int k = tid; k < N; k += TPB) any |= (dep[k] != 0.);
            any = __syncthreads_or(any);
        }
    }
    if (any) {
#pragma unroll 4
        for (int k = tid; k < N; k += TPB) {
            const unsigned char fl = flagsA[k];
            const double ce = cero[k], ar = areaA[k], dp = dep[k];
            double th = dth[k];
            if (k + 4 * TPB < N) { pf_l2(cero + k + 4 * TPB); pf_l2(dep + k + 4 * TPB); pf_l2(dth + k + 4 * TPB); }
            double sf = 0.;
            if (fl & 1) {
                th += dp / ar;
                dth[k] = th;
                const double erosion = ce / ar;
                sf = erosion + th;
            }
            sedflux[k] = sf;
        }
    } else {
#pragma unroll 4
        for (int k = tid; k < N; k += TPB) {
            const unsigned char fl = flagsA[k];
            const double ce = cero[k], ar = areaA[k], th = dth[k];
            if (k + 4 * TPB < N) { pf_l2(cero + k + 4 * TPB); pf_l2(dth + k + 4 * TPB); }
            const double erosion = ce / ar;
            sedflux[k] = (fl & 1) ? erosion + th : 0.;
        }
    }
    __syncthreads();
    sub_tick(st, 19);
}

// phase: apply ed, hillslope diffusion, boundary, tectonics.  buildFlux.py:193-195, :252-272, :292-315;
// sfd.c:155-185 (diffusion); diffLinear.py:150-182 (sedflux coefficients)
__device__ void phase_apply(const DevConst &c, Smem &S, double sea, double timestep)
{
    const int N = c.N, tid = threadIdx.x;
    double *z = S.P.f[F_Z], *cumdiff = S.P.f[F_CUMDIFF], *cumhill = S.P.f[F_CUMHILL];
    const double *sedflux = S.P.f[F_SEDFLUX];
    double *cd = S.P.f[F_VAL];
    for (int k = tid; k < N; k += TPB) { z[k] += sedflux[k]; cumdiff[k] += sedflux[k]; }
    __syncthreads();
    if (c.CDr > 0.) phase_marine_diffusion(c, S, sea, timestep);
    for (int g = tid; g < N; g += TPB) {
        double acc = 0.;
        const unsigned char fg = c.flags[g];
        if (fg & 2) {
            const double zg = z[g];
            for (int p = 0; p < c.KD; ++p) {
                const int j = c.ngb[(size_t)p * N + g];
                if (j < 0) break;
                const double di = c.edge[(size_t)p * N + g], ed = c.vor[(size_t)p * N + g], zj = z[j];
                const bool bj = c.flags[j] & 2;
                if (bj && di > 0.) acc += ed * (zj - zg) / di;
                if (!bj) { if (zj < zg && di > 0.) acc += ed * (zj - zg) / di; }
            }
        }
        const double area = c.area[g];
        const double ac = (area > 0.) ? 1. / area : 0.;
        double coeff = ac * ((z[g] >= sea) ? c.CDa : c.CDm);
        if (!(fg & 2)) { coeff = 0.; acc = 0.; }
        if (S.diff_out) S.diff_out[g] = acc;
        cd[g] = coeff * acc * timestep;
    }
    __syncthreads();
    for (int k = tid; k < N; k += TPB)
        if (c.flags[k] & 1) { const double d = cd[k]; z[k] += d; cumdiff[k] += d; cumhill[k] += d; }
    __syncthreads();
    if (c.btype != 0) {
        const double add = (c.btype == 1) ? -0.1 : (c.btype == 2 ? 0. : 100.);
        for (int k = tid; k < c.bounds; k += TPB) {
            const double zp = z[c.parent[k]];
            cd[k] = (c.btype == 2) ? zp : zp + add;
        }
        __syncthreads();
        for (int k = tid; k < c.bounds; k += TPB) z[k] = cd[k];
        __syncthreads();
    }
    if (c.disp) {
        for (int k = tid; k < N; k += TPB) z[k] += c.disp[k] * timestep;
        __syncthreads();
    }
}

// one time step = buildFlux.streamflow (buildFlux.py:21-100) + sediment_flux (:102-320).  Proposal scalars (tNow, step,
// rain, erodibility, seed, tStop) are read from / written to S (see Smem).
#define BL_TICK(i) do { if (threadIdx.x == 0) { long long *pc_ = S.c.phase_cyc + (size_t)blockIdx.x * NPHASE; \
                                               const long long t1_ = clock64(); tick_add(pc_ + (i), t1_ - S.tick0); S.tick0 = t1_; } } while (0)
__device__ void time_step(const DevConst &c, Smem &S)
{
    const int tid = threadIdx.x;
    if (tid == 0) { S.sea = sea_level(c, S.tNow); S.tick0 = clock64(); }
    __syncthreads();
    phase_fill(c, S, S.sea);
    BL_TICK(0);
    phase_sfd(c, S);
    BL_TICK(1);
    // real-surface stack (bases ascending), then filled-surface stack (bases shuffled)
    int nbase1, nbase;
    tree_links(c, S.P.i[I_RCV1], S.P.i[I_FC], S.P.i[I_NS], nullptr, nullptr, nullptr);
    phase_stack(c, S, S.P.i[I_RCV1], S.P.i[I_FC], S.P.i[I_NS], S.P.i[I_STACK1], S.P.i[I_POS1], S.P.i[I_LA], nbase1,
                false, 0ull);
    if (tid == 0) S.nbase1 = nbase1;     // read after the barriers inside phase_basins
    BL_TICK(2);
    phase_basins(c, S, S.P.i[I_LA], nbase1);
    BL_TICK(3);
    // the filled tree keeps I_NS for the real tree (basin segments are needed later); its links live in I_LC / I_PS
    // (+ I_SUCC as next sibling)
    tree_links(c, S.P.i[I_RCV], S.P.i[I_FC], S.P.i[I_SUCC], S.P.i[I_LC], S.P.i[I_PS], S.P.i[I_CNT]);
    phase_stack(c, S, S.P.i[I_RCV], S.P.i[I_FC], S.P.i[I_SUCC], S.P.i[I_STACK], S.P.i[I_POS], S.P.i[I_LB], nbase,
                c.shuffle != 0, mix64(S.seed + (unsigned long long)S.step));
    BL_TICK(4);
    phase_drainage(c, S, S.P.i[I_LA], S.nbase1, S.sea);
    BL_TICK(5);
    phase_discharge(c, S, S.rain);
    BL_TICK(6);
    // CFL and time step (buildFlux.py:169-177)
    const double flowcfl = phase_cfl(c, S, S.erod);
    double cfl = fmin(flowcfl, c.hill);
    cfl = py2_floorish(cfl);
    cfl = fmin(cfl, S.tStop - S.tNow);
    cfl = fmax(c.minDT, cfl);
    cfl = fmin(c.maxDT, cfl);
    // flowNetwork.compute_sedflux (flowNetwork.py:579-719)
    double newdt = cfl;
    BL_TICK(7);
    phase_streampower(c, S, S.erod, S.sea, newdt);
    BL_TICK(8);
    {
        const int N = c.N;
        const double *cdepo = S.P.f[F_DEPO], *cero = S.P.f[F_ERO], *pitVol = S.P.f[F_PITVOL];
        const int *pitID = S.P.i[I_PITID], *alld = S.P.i[I_ALLDRAIN];
        int nb = 0;
        double minperc = 1.e300;
        for (int k = tid; k < N; k += TPB) {
            const double vc = (c.depo == 0) ? cero[k] : cdepo[k] + cero[k];
            const double sumvol = 0. + vc;
            const bool in1 = (sumvol > pitVol[k] && pitVol[k] > 0.);
            const bool in2 = (pitID[k] >= 0 && alld[k] == pitID[k]);
            if (in1 && in2 && (c.flags[k] & 4)) { nb = 1; minperc = fmin(minperc, pitVol[k] / sumvol); }
        }
        nb = __syncthreads_or(nb);
        if (nb) { minperc = blk_min(minperc, S); newdt = cfl * minperc; }
        newdt = py2_floorish(newdt);
        newdt = fmax(c.minDT, newdt);
        if (newdt < cfl) phase_streampower(c, S, S.erod, S.sea, newdt);
    }
    BL_TICK(9);
    phase_erodep(c, S, S.sea);
    BL_TICK(10);
    phase_apply(c, S, S.sea, newdt);
    BL_TICK(11);
    if (tid == 0) { S.tNow += newdt; S.last_dt = newdt; S.step += 1; }
    __syncthreads();
}

#include "bl_fast.cuh"
#include "bl_fastc.cuh"

// bl_mcmc.interpolateArray (bl_mcmc.py:198-213) on the mesh-constant k=3 table
__device__ inline double interp_cell(const DevConst &c, const double *v, int g)
{
    const int *ix = c.grid_idx + 3 * (size_t)g;
    if (c.grid_exact[g]) return v[ix[0]];
    const double *w = c.grid_w + 3 * (size_t)g;
    const double num = ((0. + v[ix[0]] * w[0]) + v[ix[1]] * w[1]) + v[ix[2]] * w[2];
    const double den = ((0. + w[0]) + w[1]) + w[2];
    return num / den;
}

template <int KDT>
__global__ void __launch_bounds__(TPB, BL_MINB)
bl_run_kernel(const DevConst cpar, int nstops, const double *__restrict__ stops, const int *__restrict__ ckpt,
              double *out_pts, double *out_sse, double *out_elev, double *out_erdp, int nck, int single_step,
              int dyn_bytes)
{
    __shared__ Smem S;
    extern __shared__ __align__(16) unsigned char dsm[];
    const int b = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) {
        S.c = cpar;
        prop_init(S.P, cpar.prop_base + (size_t)b * cpar.prop_stride, cpar.N);
        S.dsm = (KDT > 0) ? dsm : nullptr;
        S.dsm_bytes = (KDT > 0) ? dyn_bytes : 0;
        S.diff_out = nullptr;
        S.tma_phase = 0u;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&S.mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const DevConst &c = S.c;
    for (int i = tid; i <= c.nlev && i <= LEVMAX; i += TPB) S.lev_off[i] = c.lev_off[i];
    if (tid < 8) S.ibuf[tid] = 0;
    __syncthreads();
    if (tid == 0) {
        S.tNow = c.tnow[b]; S.step = c.steps[b]; S.last_dt = c.lastdt[b];
        S.rain = c.rain[b]; S.erod = c.erod[b]; S.seed = c.seed ? c.seed[b] : 0ull;
    }
    __syncthreads();
    int ck_seen = 0;
    for (int s = 0; s < nstops; ++s) {
        __syncthreads();                    // every thread is done with the previous stop's loop condition
        if (tid == 0) S.tStop = stops[s];
        __syncthreads();
        while (S.tNow < S.tStop) {
            if (KDT > 0) {
                if (c.compact) fastc::time_step<(KDT > 0 ? KDT : 8)>(c, S, dsm);
                else fast::time_step<(KDT > 0 ? KDT : 8)>(c, S, dsm);
            }
            else time_step(c, S);
            if (single_step) break;
        }
        const int ck = ckpt ? ckpt[s] : -1;
        if (ck >= 0 && c.G > 0) {
            const double *z = S.P.f[F_Z], *cum = S.P.f[F_CUMDIFF];
            if (out_pts)
                for (int i = tid; i < c.npts; i += TPB)
                    out_pts[((size_t)b * BL_NCKPT_MAX + ck) * BL_NPTS_MAX + i] = interp_cell(c, cum, c.pts_cell[i]);
            if (out_elev || out_erdp || (out_sse && c.real_elev)) {
                double sse = 0.;
                for (int g = tid; g < c.G; g += TPB) {
                    const double e = interp_cell(c, z, g);
                    if (out_elev) out_elev[((size_t)b * nck + ck_seen) * c.G + g] = e;
                    if (out_erdp) out_erdp[((size_t)b * nck + ck_seen) * c.G + g] = interp_cell(c, cum, g);
                    if (c.real_elev) { const double d = e - c.real_elev[g]; sse += d * d; }
                }
                if (out_sse && c.real_elev) {
                    sse = blk_sum(sse, S);
                    if (tid == 0) {
                        out_sse[b] = sse;
                        if (!isfinite(sse)) atomicOr(&c.status[b], BL_ST_NAN);
                    }
                }
            }
            ++ck_seen;
            __syncthreads();
        }
    }
    {   // NaN / Inf guard: a non-finite elevation or clock marks the evaluation as failed
        const double *z = S.P.f[F_Z];
        int bad = !isfinite(S.tNow);
        for (int k = tid; k < c.N; k += TPB) bad |= !isfinite(z[k]);
        bad = __syncthreads_or(bad);
        if (tid == 0 && bad) atomicOr(&c.status[b], BL_ST_NAN);
    }
    if (tid == 0) { c.tnow[b] = S.tNow; c.steps[b] = S.step; c.lastdt[b] = S.last_dt; }
}

// ------------------------------------------------------------------------------------------------
// Per-stage entry points (bl_fill, bl_sfd, bl_stack, ... in bayeslands_b200.h): ONE routine of the time
// step on proposal slot 0, with the same device functions the fused step uses (shared-memory path when the context
// selected it, generic path otherwise).  The host copies the caller's arrays into / out of the slot.
// ------------------------------------------------------------------------------------------------
#ifdef BL_WITH_STAGE
template <int KDT>
__global__ void __launch_bounds__(TPB, BL_MINB)
bl_stage_kernel(const DevConst cpar, const StageArgs a, int dyn_bytes)
{
    __shared__ Smem S;
    extern __shared__ __align__(16) unsigned char dsm[];
    const int tid = threadIdx.x;
    if (tid == 0) {
        S.c = cpar;
        prop_init(S.P, cpar.prop_base, cpar.N);
        S.dsm = (KDT > 0) ? dsm : nullptr;
        S.dsm_bytes = (KDT > 0) ? dyn_bytes : 0;
        S.diff_out = nullptr;
        S.sea = a.sea; S.rain = a.rain; S.erod = a.erod; S.tNow = 0.; S.tStop = 0.; S.step = a.step; S.seed = a.seed;
        S.tick0 = clock64();
    }
    __syncthreads();
    const DevConst &c = S.c;
    for (int i = tid; i <= c.nlev && i <= LEVMAX; i += TPB) S.lev_off[i] = c.lev_off[i];
    if (tid < 8) S.ibuf[tid] = 0;
    __syncthreads();
    const int N = c.N;
    constexpr int K = (KDT > 0) ? KDT : 8;
    const double *z = S.P.f[F_Z], *W = S.P.f[F_FILLH];
    // helpers: "some node lies under the sea" (S.ibuf[1]) and the real-surface base list = {i : pitID[i] == i}
    auto any_marine = [&]() { int m = 0; for (int k = tid; k < N; k += TPB) m |= (W[k] < a.sea); return __syncthreads_or(m); };
    switch (a.stage) {
        case BL_STAGE_FILL: {                      // pdstack.pitfilling(elev, 0, sealevel) -> fillH (+ receivers on the fast path)
            if (KDT > 0) { int am; fast::fill_sfd<K>(c, S, dsm, a.sea, am); }
            else phase_fill(c, S, a.sea);
            break;
        }
        case BL_STAGE_SFD: {                       // sfd.directions(fillH, elev) + sfd.directions_base(elev) -> rcv, rcv1, maxh
            if (KDT > 0) {
                double *Ws = (double *)dsm, *zs = Ws + c.Nq;
                for (int i = tid; i < N; i += TPB) { Ws[i] = W[i]; zs[i] = z[i]; }
                __syncthreads();
                int am;
                fast::fill_receivers<K>(c, S, dsm, a.sea, am);
            } else phase_sfd(c, S);
            __syncthreads();
            // pyMaxDep (sfd.c:88-90, :109): largest rise to a neighbour; not used by the path (slp_cr = 0), stage output only
            double *maxdep = S.P.f[F_CAPH];
            for (int g = tid; g < N; g += TPB) {
                double dd = 0.;
                for (int p = 0; p < c.KD; ++p) {
                    const int j = c.ngb[(size_t)p * N + g];
                    if (j < 0) break;
                    const double dh = z[j] - z[g];
                    if (dh > dd) dd = dh;
                }
                maxdep[g] = dd;
            }
            break;
        }
        case BL_STAGE_STACK: {                     // fstack.build(base, receivers, delta) -> stack (bases ascending or shuffled)
            int nbase;
            const unsigned long long ss = mix64(a.seed + (unsigned long long)a.step);
            if (KDT > 0) fast::links_stack<K>(c, S, dsm, S.P.i[I_RCV], S.P.i[I_STACK], S.P.i[I_POS], S.P.i[I_LB], nbase, true, a.shuffle != 0, ss);
            else {
                tree_links(c, S.P.i[I_RCV], S.P.i[I_FC], S.P.i[I_SUCC], S.P.i[I_LC], S.P.i[I_PS], S.P.i[I_CNT]);
                phase_stack(c, S, S.P.i[I_RCV], S.P.i[I_FC], S.P.i[I_SUCC], S.P.i[I_STACK], S.P.i[I_POS], S.P.i[I_LB], nbase, a.shuffle != 0, ss);
            }
            if (tid == 0) S.P.i[I_MARK][0] = nbase;
            break;
        }
        case BL_STAGE_DISCHARGE: {                 // flowcompute.discharge(stack, receivers, elev, area * rain) -> Q
            if (KDT > 0) {
                int nbase;
                fast::links_stack<K>(c, S, dsm, S.P.i[I_RCV], S.P.i[I_STACK], S.P.i[I_POS], S.P.i[I_LB], nbase, true, false, 0ull);
                fast::discharge(c, S, dsm, a.rain);
            } else {
                tree_links(c, S.P.i[I_RCV], S.P.i[I_FC], S.P.i[I_SUCC], S.P.i[I_LC], S.P.i[I_PS], S.P.i[I_CNT]);
                phase_discharge(c, S, a.rain);
            }
            break;
        }
        case BL_STAGE_BASINS: {                    // flowcompute.basinparameters(stack1, receivers1, elev, fillH, area) -> pitID, pitVolume
            int nbase1;
            if (KDT > 0) {
                fast::links_stack<K>(c, S, dsm, S.P.i[I_RCV1], S.P.i[I_STACK1], S.P.i[I_POS1], S.P.i[I_LA], nbase1, false, false, 0ull);
                fast::basins(c, S, dsm, S.P.i[I_LA], nbase1);
            } else {
                tree_links(c, S.P.i[I_RCV1], S.P.i[I_FC], S.P.i[I_NS], nullptr, nullptr, nullptr);
                phase_stack(c, S, S.P.i[I_RCV1], S.P.i[I_FC], S.P.i[I_NS], S.P.i[I_STACK1], S.P.i[I_POS1], S.P.i[I_LA], nbase1, false, 0ull);
                phase_basins(c, S, S.P.i[I_LA], nbase1);
            }
            break;
        }
        case BL_STAGE_DRAINAGE: {                  // flowcompute.basindrainage / basindrainageall -> pitDrain, allDrain
            const int *pitID = S.P.i[I_PITID];
            const int nbase1 = blk_compact(N, S.P.i[I_LA], [&](int i) { return pitID[i] == i; }, S);
            const int am = any_marine();
            if (KDT > 0) fast::drainage(c, S, dsm, S.P.i[I_LA], nbase1, a.sea, am);
            else phase_drainage(c, S, S.P.i[I_LA], nbase1, a.sea);
            break;
        }
        case BL_STAGE_STREAMPOWER: {               // flowcompute.streampower(stack, receivers, pitID, pitVolume, pitDrain, ...) -> cdepo, cero
            const int *pitID = S.P.i[I_PITID];
            const int nbase1 = blk_compact(N, S.P.i[I_LA], [&](int i) { return pitID[i] == i; }, S);
            const int am = any_marine();
            if (tid == 0) S.ibuf[1] = am;
            int nbase;
            const unsigned long long ss = mix64(a.seed + (unsigned long long)a.step);
            if (KDT > 0) {
                fast::links_stack<K>(c, S, dsm, S.P.i[I_RCV], S.P.i[I_STACK], S.P.i[I_POS], S.P.i[I_LB], nbase, true, a.shuffle != 0, ss);
                fast::discharge(c, S, dsm, 0.);     // builds the depth buckets of the routing; its Q is replaced by the caller's
                __syncthreads();
                for (int k = tid; k < N; k += TPB) S.P.f[F_Q][k] = a.q_in[k];
                __syncthreads();
                fast::streampower(c, S, dsm, a.erod, a.sea, a.dt, false, S.P.i[I_LA], nbase1);
            } else {
                tree_links(c, S.P.i[I_RCV], S.P.i[I_FC], S.P.i[I_SUCC], S.P.i[I_LC], S.P.i[I_PS], S.P.i[I_CNT]);
                phase_stack(c, S, S.P.i[I_RCV], S.P.i[I_FC], S.P.i[I_SUCC], S.P.i[I_STACK], S.P.i[I_POS], S.P.i[I_LB], nbase, a.shuffle != 0, ss);
                for (int k = tid; k < N; k += TPB) S.P.f[F_Q][k] = a.q_in[k];
                __syncthreads();
                phase_streampower(c, S, a.erod, a.sea, a.dt);
            }
            break;
        }
        case BL_STAGE_DIFFUSE: {                   // sfd.diffusion(elev, borders2, ...) -> raw flux (hook in apply / phase_apply)
            for (int k = tid; k < N; k += TPB) S.P.f[F_SEDFLUX][k] = 0.;
            if (tid == 0) S.diff_out = a.diff_out;
            __syncthreads();
            if (KDT > 0) fast::apply<K>(c, S, dsm, a.sea, 0.);
            else phase_apply(c, S, a.sea, 0.);
            break;
        }
        case BL_STAGE_INTERP: {                    // bl_mcmc.interpolateArray on the mesh-constant k = 3 table
            for (int g = tid; g < c.G; g += TPB) a.grid_out[g] = interp_cell(c, a.v_in, g);
            break;
        }
        default: break;
    }
}
#endif

__global__ void bl_reset_kernel(const DevConst c, const double *rain, const double *erod,
                                const unsigned long long *seed, double tStart)
{
    __shared__ Prop P;
    const int b = blockIdx.x;
    if (threadIdx.x == 0) prop_init(P, c.prop_base + (size_t)b * c.prop_stride, c.N);
    __syncthreads();
    for (int k = threadIdx.x; k < c.N; k += blockDim.x) {
        P.f[F_Z][k] = c.z0[k];
        P.f[F_CUMDIFF][k] = 0.;
        P.f[F_CUMHILL][k] = 0.;
    }
    if (threadIdx.x == 0) {
        c.rain[b] = rain[b];
        c.erod[b] = erod[b];
        c.seed[b] = seed ? seed[b] : 0ull;
        c.tnow[b] = tStart;
        c.steps[b] = 0;
        c.status[b] = 0u;
        c.lastdt[b] = 0.;
        for (int i = 0; i < NPHASE; ++i) c.phase_cyc[(size_t)b * NPHASE + i] = 0;
    }
}


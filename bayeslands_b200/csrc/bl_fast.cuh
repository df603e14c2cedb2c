// bl_fast.cuh - shared-memory resident time step (included by bl_engine.cu).
//
// Same algorithms and the same fp64 operation order as the generic phases in bl_engine.cu (so results stay
// bit-identical to the CPU oracle), but every iterative / pointer-chasing phase runs on arrays held in the
// CTA's 227 KB of shared memory instead of L2:
//   fill + receivers : W (fillH) and z as fp64 + per-node dirty flags        2*8N + N  bytes
//   stack order      : Euler tour as packed (next:16 | rank:16) words          8N + 2N
//   basins           : root ids (u16) + stack-ordered volume terms (fp64)      2N + 8N
//   discharge / SPL  : node values fp64 + donor links (u16) + counters (u8)    8N + 6N + 3N
// Neighbour ids come from a u16 slot-major table (coalesced).  Requires N < 32768 and 17*Nq <= 225 KB.


// the diffusion pass of the compact layout (bl_fastc.cuh) needs 8 bytes of shared memory per node only and is the faster
// one wherever there is no river-borne marine diffusion: fast::time_step uses it too
namespace fastc {
template <int KDT>
__device__ __noinline__ void apply_c(const DevConst &c, Smem &S, unsigned char *dsm, double sea, double timestep);
}

namespace fast {

// The phases are separate (__noinline__) functions so that each gets its own register allocation under the kernel's
// 64-register cap; BL_ASSUME_SHARED tells the compiler that Smem / DevConst / the dynamic buffer live in shared memory,
// so that their accesses stay LDS / STS with 32-bit addresses although the pointers arrive as generic ones.
#define BL_ASSUME_SHARED(c, S, dsm) do { __builtin_assume(__isShared(&(c))); __builtin_assume(__isShared(&(S))); \
                                         __builtin_assume(__isShared(dsm)); } while (0)

constexpr unsigned short NONE16 = 0xFFFFu;
constexpr unsigned END16 = 0xFFFFu;

// Neighbour ids of node k: node-major u16 table [N][KDT], fetched as KDT/4 8-byte words (one or two coalesced
// transactions per node) and kept packed in registers - an indexable u16 array would live in local memory, which
// with the L1 carved out for shared memory costs an L2 round trip per access.
template <int KDT>
struct Ids {
    unsigned w[KDT / 2];
    __device__ __forceinline__ unsigned get(int p) const { return (w[p >> 1] >> (16 * (p & 1))) & 0xFFFFu; }
};
template <int KDT>
__device__ __forceinline__ void load_ids(const DevConst &c, int k, Ids<KDT> &ids)
{
    const uint2 *src = reinterpret_cast<const uint2 *>(c.ngb16 + (size_t)k * KDT);
#pragma unroll
    for (int q = 0; q < KDT / 4; ++q) {
        const uint2 v = __ldg(src + q);
        ids.w[2 * q] = v.x;
        ids.w[2 * q + 1] = v.y;
    }
}
template <int KDT>
__device__ __forceinline__ void none_ids(Ids<KDT> &ids)
{
#pragma unroll
    for (int q = 0; q < KDT / 2; ++q) ids.w[q] = 0xFFFFFFFFu;
}

// sfd.c:55-153 on W / z held in shared memory ([Nq] each at the start of dsm): receivers on both surfaces
template <int KDT>
__device__ __noinline__ void fill_receivers(const DevConst &c, Smem &S, unsigned char *dsm, double sea, int &any_marine)
{
    BL_ASSUME_SHARED(c, S, dsm);
    const int N = c.N, Nq = c.Nq, tid = threadIdx.x;
    double *W = (double *)dsm, *zs = W + Nq;
    // receivers on both surfaces
    int *rcv = S.P.i[I_RCV], *rcv1 = S.P.i[I_RCV1];
    double *maxh = S.P.f[F_MAXH], *fillH = S.P.f[F_FILLH];
    int mar = 0;
#pragma unroll 2
    for (int g = tid; g < N; g += TPB) {
        Ids<KDT> ids;
        load_ids<KDT>(c, g, ids);
        const double zg = zs[g], wg = W[g];
        int low1 = g, low = g;
        double zl = zg, wl = wg, dH = 1.e6;
#pragma unroll
        for (int p = 0; p < KDT; ++p) {
            if (ids.get(p) == NONE16) continue;
            const int j = ids.get(p);
            const double zj = zs[j], wj = W[j];
            if (zj < zl) { low1 = j; zl = zj; }
            if (wj < wl) { low = j; wl = wj; }
            const double dh = zj - zg;
            if (dh >= 0. && dh < dH) dH = dh;
        }
        rcv1[g] = low1;
        rcv[g] = low;
        if (dH > 9.99e5) dH = 0.;
        maxh[g] = dH;
        fillH[g] = wg;
        mar |= (wg < sea);
    }
    any_marine = __syncthreads_or(mar);
}

// ---- shared-memory accessors with explicit 32-bit addresses ------------------------------------------------
// The latency-critical loops address shared memory through 32-bit window offsets (one __cvta_generic_to_shared per
// buffer, plain integer arithmetic afterwards) and predicated stores: measured in situ, a dependent instruction costs
// ~8 cycles and a branch / re-convergence pair ~40, so what matters in a round with a few hundred active lanes is the
// number of instructions on the dependent chain, not the amount of work.
__device__ __forceinline__ unsigned s32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double lds_f64(unsigned a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint2 lds_v2(unsigned a)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ unsigned lds_u8(unsigned a)
{
    unsigned v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f64(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sts_u8_if(bool p, unsigned a, unsigned v)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %0, 0;\n\t@q st.shared.u8 [%1], %2;\n\t}" ::"r"((unsigned)p), "r"(a), "r"(v) : "memory");
}

// ---- depression filling, per-warp worklists (levels contiguous in node id) --------------------------------
// Same literal Gauss-Seidel sweep as below, for meshes whose level-l nodes are the id range
// [bounds + lev_off[l], bounds + lev_off[l+1]) (mesh.py numbers them so).  A round (one level of one sweep) evaluates a
// thin front - a few hundred pending nodes out of ~3 000 - so its cost is the LENGTH of the dependent instruction chain
// between two block barriers, not the amount of work.  Hence:
//  * the pending set is one BYTE per node: a changed node wakes its neighbours with plain stores (no atomics; concurrent
//    wakers store the same value), and nodes of other levels are the only flags a round writes besides clearing its own;
//  * no block-wide compaction: half of the warps sweep, each owning 8 groups of 32 ids out of every 256 * NWF (dealt
//    round-robin, so a front confined to one region of the landscape still spreads over all of them); a lane reads the 8
//    flags of 8 consecutive ids as one 64-bit word, eight ballots pack the warp's pending ids into a byte list of its
//    own, lanes evaluate them 32 at a time: straight-line code, no exchange between warps, ONE block barrier per round;
//  * the evaluation issues every load it may need up front (neighbour ids, then all neighbour values through safe
//    addresses), selects afterwards and stores under predicates: nothing sits behind a data-dependent branch.
// W and z both stay in shared memory; the only global access in the sweep is the node's packed neighbour ids.
template <int KDT>
__device__ __forceinline__ void fill_eval(const unsigned short *__restrict__ ngb16f, int k, unsigned aW, unsigned aZ,
                                          unsigned aP, double sea, double eps, double TH, int &loc)
{
    Ids<KDT> ids;        // empty slots hold the dummy node N, whose W is 2e6 (DevConst::ngb16f): no slot tests below
    {
        const uint2 *src = reinterpret_cast<const uint2 *>(ngb16f + (size_t)k * KDT);   // the longest latency: issued first
#pragma unroll
        for (int q = 0; q < KDT / 4; ++q) { const uint2 v = __ldg(src + q); ids.w[2 * q] = v.x; ids.w[2 * q + 1] = v.y; }
    }
    const unsigned ak = aW + 8u * (unsigned)k;
    const double wk = lds_f64(ak), zk = lds_f64(aZ + 8u * (unsigned)k);
    // min over the neighbours (exact: any association gives the same bits); four loads in flight at a time keep the
    // evaluation within the register budget (a spill costs an L2 round trip: the L1 is carved out for shared memory)
    double hmin = 2.e6;
    double wj[KDT];
#pragma unroll
    for (int p = 0; p < KDT; ++p) wj[p] = lds_f64(aW + 8u * ids.get(p));
#pragma unroll
    for (int p0 = 0; p0 < KDT; p0 += 4) hmin = fmin(hmin, fmin(fmin(wj[p0], wj[p0 + 1]), fmin(wj[p0 + 2], wj[p0 + 3])));
    // PDalgo.f90:51-79 without branches (same arithmetic, same comparisons):
    //   if (W > z) { if (z >= he) W = z; else if (W > he) { W = he; if capped W = floor + TH else change = true } }
    const double he = hmin + eps;
    const bool wet = wk > zk, dry = zk >= he, lower = wk > he;
    const double floor_ = (zk >= sea) ? zk : sea;
    const bool capped = (he - floor_) > TH;
    const double nlow = capped ? floor_ + TH : he;
    const double nw = wet ? (dry ? zk : (lower ? nlow : wk)) : wk;
    if (wet && !dry && lower && !capped) loc = 1;
    if (nw != wk) {
        sts_f64(ak, nw);
        // wake only the neighbours this change can move: re-evaluating j alters W(j) only if
        // min_nb W + eps < W(j), and the neighbour that brings the minimum down below W(j) - eps is the
        // one that wakes j (other levels do not change during this pass).  Flags of ghost nodes
        // and of the dummy node may get set too: no level ever reads them.
        const double lim = nw + eps;
        if (KDT > 8) {      // wide rows: re-read the neighbour values instead of holding 2 * KDT registers across the min
#pragma unroll
            for (int p = 0; p < KDT; ++p) wj[p] = lds_f64(aW + 8u * ids.get(p));
        }
#pragma unroll
        for (int p = 0; p < KDT; ++p) sts_u8_if(wj[p] > lim, aP + ids.get(p), 1u);
    }
}

template <int KDT>
__device__ __noinline__ void fill_bitmap(const DevConst &c, Smem &S, unsigned char *dsm, double sea)
{
    BL_ASSUME_SHARED(c, S, dsm);
    const int N = c.N, Nq = c.Nq, tid = threadIdx.x, nb = c.bounds;
    const int lane = tid & 31, warp = tid >> 5;
    // Half of the warps sweep, 8 flags per lane: a thin front then gives each sweeping warp ~20 pending ids, i.e. one
    // evaluation pass with most lanes busy.
    constexpr int NWF = (NWARP >= 2) ? NWARP / 2 : 1;
    {
        double *W = (double *)dsm, *zs = W + Nq;
        unsigned char *pend = (unsigned char *)(zs + Nq);                      // [Nq] pending flags (8-byte aligned)
        const double *z = S.P.f[F_Z];
        for (int i = tid; i < N; i += TPB) { const double zi = z[i]; zs[i] = zi; W[i] = (i < nb) ? zi : 1.e6; }
        // initially pending: the interior nodes next to a ghost node (see fastc::fill_c; opt_debug bit 128: all of them)
        for (int i = tid; i < Nq; i += TPB)
            pend[i] = (i >= nb && i < N && ((c.dbg & 128) || ((c.pend0[i >> 5] >> (i & 31)) & 1u))) ? 1 : 0;
        if (tid == 0) W[N] = 2.e6;                                              // the dummy neighbour of empty slots
    }
    __syncthreads();
    const unsigned aW = s32(dsm), aZ = aW + 8u * (unsigned)Nq, aP = aZ + 8u * (unsigned)Nq;
    const unsigned aL = aW + 17u * (unsigned)Nq + (unsigned)warp * 256u;       // [NWF][256] pending slots of this warp
    const unsigned short *ngb16f = c.ngb16f;
    const double eps = c.eps, TH = c.TH;
    const int nlev = c.nlev;
    const unsigned lt = (1u << lane) - 1u;
    // slot s = 8 * lane + byte of this warp's 256 ids of a batch: id = base + ((lane >> 2) * NWF + warp) * 32 + (lane & 3) * 8 + byte
    const int lane_off = ((lane >> 2) * NWF + warp) * 32 + (lane & 3) * 8;
#ifdef BL_FILL_TRACE
    // trace build only (tools/fill_trace.py): clocks of warps 0 and 5 along each round (whole-warp reads, lane 0 stores)
    unsigned long long *trc = S.P.keys + 8 + (warp == 5 ? 8000 : 0);
    int tri = 0;
    const bool trw = (warp == 0 || warp == 5) && lane == 0;
#define TRC(x) do { const long long v_ = (long long)(x); if (trw && tri < 7990) trc[tri] = (unsigned long long)v_; ++tri; } while (0)
#else
#define TRC(x) do { } while (0)
#endif
    int change = 1;
    while (change) {
        int loc = 0;
        for (int l = 0; l < nlev; ++l) {
            TRC(clock64());
            if (warp < NWF) {
                const int n0 = nb + S.lev_off[l], n1 = nb + S.lev_off[l + 1];     // node id range of the level
#ifdef BL_FILL_TRACE
                int offsum = 0;
#endif
                for (int base = n0 & ~7; base < n1; base += 256 * NWF) {
                    const int k0 = base + lane_off;
                    uint2 w2 = lds_v2(aP + (unsigned)min(k0, Nq - 8));
                    if (k0 >= n1) w2 = make_uint2(0u, 0u);
                    // ids outside [n0, n1) in the boundary words belong to other levels: not ours to take
                    bool pb[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        pb[q] = (((q < 4 ? w2.x : w2.y) >> (8 * (q & 3))) & 0xFFu) != 0u && (unsigned)(k0 + q - n0) < (unsigned)(n1 - n0);
                    // the id rows of these 8 nodes are one 128-byte line: requested now, needed by the evaluation below
                    if ((w2.x | w2.y) != 0u) asm volatile("prefetch.global.L1 [%0];" ::"l"(ngb16f + (size_t)k0 * KDT));
                    unsigned mb[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) mb[q] = __ballot_sync(0xffffffffu, pb[q]);
                    int off = 0;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        sts_u8_if(pb[q], aP + (unsigned)(k0 + q), 0u);
                        sts_u8_if(pb[q], aL + (unsigned)(off + __popc(mb[q] & lt)), (unsigned)(8 * lane + q));
                        off += __popc(mb[q]);
                    }
                    __syncwarp();
                    TRC(clock64());
                    for (int i = lane; i < off; i += 32) {
                        const int s = (int)lds_u8(aL + (unsigned)i);
                        const int ln = s >> 3;
                        fill_eval<KDT>(ngb16f, base + ((ln >> 2) * NWF + warp) * 32 + (ln & 3) * 8 + (s & 7), aW, aZ, aP, sea, eps, TH,
                                       loc);
                    }
                    __syncwarp();
#ifdef BL_FILL_TRACE
                    offsum += off;
#endif
                }
#ifdef BL_FILL_TRACE
                TRC(clock64());
                TRC(offsum);
#endif
            }
            if (l + 1 < nlev) __syncthreads();
#ifdef BL_FILL_TRACE
            else __syncthreads();
#endif
            TRC(clock64());
        }
        change = __syncthreads_or(loc);
    }
#ifdef BL_FILL_TRACE
    if (trw) trc[-8 + (warp == 5 ? 1 : 0)] = (unsigned long long)tri;
#endif
#undef TRC
}

// ---- depression filling + receivers ---------------------------------------------------------------
// PDalgo.f90:19-91 with the literal sweep semantics of phase_fill.  A node is re-evaluated only when one of
// its neighbours changed since its last evaluation (re-evaluating it otherwise is a no-op: after an
// evaluation W(k) <= min_nb W + eps, or W(k) = z(k) for good), so the values and the `change` flag of every
// sweep - and hence the reference's early stop - are unchanged.  Pending nodes are kept in one dense
// worklist per Gauss-Seidel level (a changed node appends its not-yet-pending neighbours to the list of THEIR
// level, which is consumed at that level's next turn - later in this sweep or in the next one), so a level
// pass costs a few full warps instead of every warp running the predicated evaluation code.
// sfd.c:55-153 follows on the same arrays.
template <int KDT>
__device__ void fill_sfd(const DevConst &c, Smem &S, unsigned char *dsm, double sea, int &any_marine)
{
    const int N = c.N, Nq = c.Nq, tid = threadIdx.x, nb = c.bounds;
    if (c.lev_contig) { fill_bitmap<KDT>(c, S, dsm, sea); fill_receivers<KDT>(c, S, dsm, sea, any_marine); return; }
    double *W = (double *)dsm;                                          // [Nq]
    unsigned short *pend = (unsigned short *)(dsm + (size_t)8 * Nq);    // [<2Nq] one power-of-two ring per level
    unsigned char *lev8 = dsm + (size_t)12 * Nq + 128;                  // [Nq] level of each node
    unsigned *bits = (unsigned *)(dsm + (size_t)13 * Nq + 256);         // [Nq/32] "is pending"
    const double *z = S.P.f[F_Z];
#pragma unroll 4
    for (int i = tid; i < N; i += TPB) W[i] = (i < nb) ? z[i] : 1.e6;
    for (int i = tid; i < Nq / 32; i += TPB) bits[i] = 0u;
    {   // ring i of level l lives at pend[fl_base[l] + (i & fl_mask[l])], capacity = level size rounded up to 2^k
        int base = 0;
        for (int l = 0; l < c.nlev; ++l) {
            const int b0 = S.lev_off[l], e = S.lev_off[l + 1];
            int cap = 1;
            while (cap < e - b0) cap <<= 1;
            for (int idx = b0 + tid; idx < e; idx += TPB) {
                const int k = c.lev_contig ? (nb + idx) : c.lev_nodes[idx];
                lev8[k] = (unsigned char)l;
                pend[base + (idx - b0)] = (unsigned short)k;
            }
            if (tid == 0) { S.fl_head[l] = 0; S.fl_tail[l] = e - b0; S.fl_base[l] = base; S.fl_mask[l] = cap - 1; }
            base += cap;
        }
    }
    __syncthreads();
    for (int i = nb + tid; i < N; i += TPB) atomicOr(&bits[i >> 5], 1u << (i & 31));
    __syncthreads();
    const double eps = c.eps, TH = c.TH;
    long long *pcd = c.phase_cyc + (size_t)blockIdx.x * NPHASE;
    int nevald = 0, nsweepd = 0;
    int change = 1;
    while (change) {
        int loc = 0;
        ++nsweepd;
        for (int l = 0; l < c.nlev; ++l) {
            const int head = S.fl_head[l], n = S.fl_tail[l] - head;
            if (n == 0) continue;                               // uniform: nothing pending at this level
            const int b0 = S.fl_base[l], msk = S.fl_mask[l];
            for (int i = tid; i < n; i += TPB) {
                const int k = pend[b0 + ((head + i) & msk)];
                atomicAnd(&bits[k >> 5], ~(1u << (k & 31)));
                const double w = W[k];
                Ids<KDT> ids;
                load_ids<KDT>(c, k, ids);
                const double zk = __ldg(z + k);
                ++nevald;
                if (w > zk) {
                    double hm[KDT];
#pragma unroll
                    for (int p = 0; p < KDT; ++p) hm[p] = (ids.get(p) != NONE16) ? W[ids.get(p)] : 2.e6;
                    double hmin;
                    if (KDT == 8) {      // min is exact: any association gives the same bits
                        hmin = fmin(fmin(fmin(hm[0], hm[1]), fmin(hm[2], hm[3])), fmin(fmin(hm[4], hm[5]), fmin(hm[6], hm[7])));
                    } else {
                        hmin = 2.e6;
#pragma unroll
                        for (int p = 0; p < KDT; ++p) hmin = fmin(hmin, hm[p]);
                    }
                    hmin = fmin(hmin, 2.e6);
                    const double he = hmin + eps;
                    double nw = w;
                    if (zk >= he) nw = zk;
                    else if (w > he) {
                        nw = he;
                        if (zk >= sea) { if (nw - zk > TH) nw = zk + TH; else loc = 1; }
                        else           { if (nw - sea > TH) nw = sea + TH; else loc = 1; }
                    }
                    if (nw != w) {
                        W[k] = nw;
                        unsigned fresh = 0;
#pragma unroll
                        for (int p = 0; p < KDT; ++p) {
                            const unsigned j = ids.get(p);
                            if (j == NONE16 || (int)j < nb) continue;
                            const unsigned bit = 1u << (j & 31);
                            if (!(atomicOr(&bits[j >> 5], bit) & bit)) fresh |= 1u << p;
                        }
#pragma unroll
                        for (int p = 0; p < KDT; ++p) {
                            if (!(fresh & (1u << p))) continue;
                            const unsigned j = ids.get(p);
                            const int lj = lev8[j];
                            const int t = atomicAdd(&S.fl_tail[lj], 1);
                            pend[S.fl_base[lj] + (t & S.fl_mask[lj])] = (unsigned short)j;
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(z + j));      // z(j) is read at j's next turn
                        }
                    }
                }
            }
            __syncthreads();
            if (tid == 0) S.fl_head[l] = head + n;   // consumed; nobody reads head[l] before its next turn
        }
        change = __syncthreads_or(loc);
    }
    (void)pcd; (void)nsweepd; (void)nevald;
    // stage z next to W for the receiver search
    double *zs = W + Nq;
    __syncthreads();
#pragma unroll 4
    for (int i = tid; i < N; i += TPB) zs[i] = z[i];
    __syncthreads();
    fill_receivers<KDT>(c, S, dsm, sea, any_marine);
}

// ---- stack order ----------------------------------------------------------------------------------
// Same Euler-tour ranking as phase_stack, with the tour held in shared memory as (next:16 | rank:16) words
// updated in place: a word is read and written as one 32-bit access, so any interleaving keeps the
// invariant "rank = number of enter events from this element up to (excluding) next".
template <int KDT>
__device__ __noinline__ void links_stack(const DevConst &c, Smem &S, unsigned char *dsm, const int *rcvg, int *stack, int *pos,
                            int *baselist, int &nbase, bool filled, bool shuffle, unsigned long long seedstep)
{
    BL_ASSUME_SHARED(c, S, dsm);
    const int N = c.N, Nq = c.Nq, tid = threadIdx.x;
    unsigned *eu = (unsigned *)dsm;                                  // [2*Nq]
    unsigned short *r16 = (unsigned short *)(dsm + (size_t)8 * Nq);  // [Nq]
    unsigned long long *K = (unsigned long long *)(dsm + (size_t)10 * Nq + 64 - ((size_t)10 * Nq) % 8);
    const int kcap = (int)(((size_t)17 * Nq - (size_t)10 * Nq - 128) / 8);
#pragma unroll 4
    for (int v = tid; v < N; v += TPB) r16[v] = (unsigned short)rcvg[v];
    __syncthreads();
    unsigned short *lc16 = (unsigned short *)S.P.i[I_LC], *ps16 = (unsigned short *)S.P.i[I_PS];
    unsigned char *cnt8 = (unsigned char *)S.P.i[I_CNT];
#pragma unroll 2
    for (int v = tid; v < N; v += TPB) {
        // both id rows are fetched up front (the receiver's row depends on shared memory only): two independent
        // global loads per node in flight instead of one behind the other
        Ids<KDT> ids, jds;
        const int r = r16[v];
        load_ids<KDT>(c, v, ids);
        load_ids<KDT>(c, r, jds);
        int f = NONE16, l = -1, n = 0;
#pragma unroll
        for (int p = 0; p < KDT; ++p) {
            if (ids.get(p) == NONE16) continue;
            const int w = ids.get(p);
            if (r16[w] == v) { ++n; if (w < f) f = w; if (w > l) l = w; }
        }
        int nx = NONE16, pv = -1;
        if (r != v) {
#pragma unroll
            for (int p = 0; p < KDT; ++p) {
                if (jds.get(p) == NONE16) continue;
                const int w = jds.get(p);
                if (w != v && r16[w] == r) {
                    if (w > v && w < nx) nx = w;
                    if (w < v && w > pv) pv = w;
                }
            }
        }
        eu[2 * v] = ((unsigned)((f != NONE16) ? 2 * f : 2 * v + 1) << 16) | 1u;
        if (r != v) eu[2 * v + 1] = ((unsigned)((nx != NONE16) ? 2 * nx : 2 * r + 1) << 16);
        if (filled) {
            lc16[v] = (l >= 0) ? (unsigned short)l : NONE16;
            ps16[v] = (pv >= 0) ? (unsigned short)pv : NONE16;
            cnt8[v] = (unsigned char)n;
        }
    }
    nbase = blk_compact(N, baselist, [&](int i) { return r16[i] == i; }, S);
    if (tid == 0) { if (filled) S.nbase = nbase; else S.nbase1 = nbase; }     // read by the later phases (barriers follow)
    if (shuffle && nbase > 1) {
        int n2 = 1;
        while (n2 < nbase) n2 <<= 1;
        unsigned long long *KK = (n2 <= kcap) ? K : S.P.keys;
        for (int i = tid; i < n2; i += TPB)
            KK[i] = (i < nbase) ? (((mix64(seedstep ^ (unsigned long long)(unsigned)baselist[i]) >> 32) << 32) |
                                   (unsigned)baselist[i])
                                : ~0ull;
        __syncthreads();
        blk_bitonic(KK, n2);
        for (int i = tid; i < nbase; i += TPB) baselist[i] = (int)(unsigned)(KK[i] & 0xFFFFFFFFull);
        __syncthreads();
    }
    int *nsroot = S.P.i[I_NS];
    for (int i = tid; i < nbase; i += TPB) {
        const int root = baselist[i];
        const int nxt = (i + 1 < nbase) ? baselist[i + 1] : NONE_I;
        eu[2 * root + 1] = ((nxt != NONE_I) ? (unsigned)(2 * nxt) : END16) << 16;
        if (!filled) nsroot[root] = nxt;
    }
    __syncthreads();
    const int M = 2 * N;
    int active = 1;
    while (active) {
        int loc = 0;
        // four elements per thread in flight: the loads of a group are issued before its stores (any interleaving
        // of word updates is valid, see above), which hides the shared-memory latency of the dependent lookup
        for (int e0 = tid; e0 < M; e0 += 4 * TPB) {
            unsigned w[4], nx[4], w2[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int e = e0 + u * TPB; w[u] = (e < M) ? eu[e] : 0xFFFF0000u; nx[u] = w[u] >> 16; }
#pragma unroll
            for (int u = 0; u < 4; ++u) w2[u] = (nx[u] != END16) ? eu[nx[u]] : 0u;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (nx[u] != END16) { eu[e0 + u * TPB] = (w2[u] & 0xFFFF0000u) | ((w[u] + w2[u]) & 0xFFFFu); loc = 1; }
        }
        active = __syncthreads_or(loc);
    }
    for (int v = tid; v < N; v += TPB) {
        const int p = N - (int)(eu[2 * v] & 0xFFFFu);
        pos[v] = p;
        stack[p] = v;
    }
    __syncthreads();
}

// ---- pit ids and volumes (FLOWalgo.f90:130-165) ------------------------------------------------------
__device__ __noinline__ void basins(const DevConst &c, Smem &S, unsigned char *dsm, const int *base1, int nbase1)
{
    BL_ASSUME_SHARED(c, S, dsm);
    const int N = c.N, Nq = c.Nq, tid = threadIdx.x;
    unsigned short *root16 = (unsigned short *)dsm;          // [Nq]
    double *val = (double *)(dsm + (size_t)8 * Nq);          // [Nq]
    const double *z = S.P.f[F_Z], *W = S.P.f[F_FILLH];
    const int *rcv1 = S.P.i[I_RCV1], *pos1 = S.P.i[I_POS1], *ns1 = S.P.i[I_NS];
    int *pitID = S.P.i[I_PITID];
    double *pitVol = S.P.f[F_PITVOL];
#pragma unroll 4
    for (int v = tid; v < N; v += TPB) root16[v] = (unsigned short)rcv1[v];
    __syncthreads();
    int ch = 1;
    while (ch) {
        int loc = 0;
        for (int v = tid; v < N; v += TPB) {
            const unsigned short r = root16[v], rr = root16[r];
            if (rr != r) { root16[v] = rr; loc = 1; }
        }
        ch = __syncthreads_or(loc);
    }
#pragma unroll 2
    for (int v = tid; v < N; v += TPB) {
        const double wv = W[v], zv = z[v];
        if (v + 2 * TPB < N) { pf_l2(W + v + 2 * TPB); pf_l2(z + v + 2 * TPB); pf_l2(pos1 + v + 2 * TPB); }
        const bool wet = wv > zv;
        const int root = root16[v];
        pitID[v] = (wet || root == v) ? root : -1;
        val[pos1[v]] = wet ? (wv - zv) * c.area[v] : 0.0;
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

// ---- pit drainage (FLOWalgo.f90:167-297), see phase_drainage -------------------------------------------
__device__ inline int pit_walk16(const unsigned short *r16, const unsigned short *pid16, const double *W, int root,
                                 bool use_sea, double sea)
{
    if (use_sea && W[root] < sea) return root | TERM_BIT;
    int donor = root;
    for (;;) {
        const int r = r16[donor];
        if (r == donor) return r | TERM_BIT;
        if (use_sea && W[r] < sea) return r | TERM_BIT;
        const int pid = pid16[r];
        if (pid == NONE16 || pid == root) donor = r;
        else if (use_sea && W[pid] < sea) donor = r;
        else return pid;
    }
}

// `staged`: r16 / pid16 are in shared memory already (compact layout: bulk copies), only the outputs are reset here
__device__ __forceinline__ void drainage_body(const DevConst &c, Smem &S, unsigned char *dsm, const int *base1, int nbase1, double sea,
                                              int any_marine, bool staged)
{
    BL_ASSUME_SHARED(c, S, dsm);
    const int N = c.N, Nq = c.Nq, tid = threadIdx.x;
    unsigned short *r16 = (unsigned short *)dsm, *pid16 = r16 + Nq;
    const int *rcv = S.P.i[I_RCV], *pitID = S.P.i[I_PITID];
    const double *W = S.P.f[F_FILLH], *pitVol = S.P.f[F_PITVOL];
    int *succ = S.P.i[I_SUCC], *succ2 = S.P.i[I_SUCC2], *drain = S.P.i[I_PITDRAIN], *alld = S.P.i[I_ALLDRAIN];
    int *mark = S.P.i[I_MARK];
    {
        const int *__restrict__ rcvA = rcv, *__restrict__ pitA = pitID;
        int *__restrict__ drainA = drain, *__restrict__ alldA = alld, *__restrict__ markA = mark;
        if (staged) {
#pragma unroll 4
            for (int v = tid; v < N; v += TPB) { drainA[v] = -1; alldA[v] = -1; markA[v] = 0; }
        } else {
#pragma unroll 4
            for (int v = tid; v < N; v += TPB) {
                r16[v] = (unsigned short)rcvA[v];
                const int p = pitA[v];
                pid16[v] = (p < 0) ? NONE16 : (unsigned short)p;
                drainA[v] = -1; alldA[v] = -1; markA[v] = 0;
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < nbase1; i += TPB) {
        const int root = base1[i];
        const int s2 = pit_walk16(r16, pid16, W, root, false, sea);
        succ2[root] = s2;
        succ[root] = any_marine ? pit_walk16(r16, pid16, W, root, true, sea) : s2;
    }
    __syncthreads();
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
    __syncthreads();
    phase_drainage(c, S, base1, nbase1, sea);   // generic path handles the cyclic case literally
}

__device__ __noinline__ void drainage(const DevConst &c, Smem &S, unsigned char *dsm, const int *base1, int nbase1, double sea,
                                      int any_marine)
{
    drainage_body(c, S, dsm, base1, nbase1, sea, any_marine, false);
}

// ---- shared-memory layout of the discharge / stream-power scans -------------------------------------------
struct ScanSm {
    double *vals;
    unsigned short *lc, *ps, *r16;
    unsigned short *order;    // nodes sorted by depth in the receiver tree (deepest bucket last)
    unsigned char *kind;
    unsigned *evbits;         // lake-entry events by reverse stack position
};
__device__ inline ScanSm scan_layout(unsigned char *dsm, int Nq)
{
    ScanSm s;
    s.vals = (double *)dsm;
    s.lc = (unsigned short *)(dsm + (size_t)8 * Nq);
    s.ps = s.lc + Nq;
    s.r16 = s.ps + Nq;
    s.order = (unsigned short *)(dsm + (size_t)14 * Nq);
    s.kind = dsm + (size_t)16 * Nq;
    s.evbits = (unsigned *)(dsm + (size_t)17 * Nq);
    return s;
}
constexpr int DEPMAX = 2040;   // bucket offsets kept in static shared memory up to this tree depth

// FLOWalgo.f90:53-81 (see phase_discharge); leaves Q in sm.vals and in global F_Q.
// Ordered accumulation without atomics: nodes are bucketed by their depth in the receiver tree (pointer jumping
// + counting sort); buckets are processed from the deepest to the roots with a barrier in between, so when a
// node is evaluated all its donors (depth + 1) are final and it sums them in the fixed descending-id order.
// The same buckets serve the sediment routing of streampower().
__device__ __noinline__ void discharge(const DevConst &c, Smem &S, unsigned char *dsm, double rain)
{
    BL_ASSUME_SHARED(c, S, dsm);
    const int N = c.N, Nq = c.Nq, tid = threadIdx.x;
    ScanSm sm = scan_layout(dsm, Nq);
    const unsigned short *lc16 = (const unsigned short *)S.P.i[I_LC], *ps16 = (const unsigned short *)S.P.i[I_PS];
    const int *rcv = S.P.i[I_RCV];
#pragma unroll 4
    for (int v = tid; v < N; v += TPB) { sm.lc[v] = lc16[v]; sm.ps[v] = ps16[v]; sm.r16[v] = (unsigned short)rcv[v]; }
    __syncthreads();
    // depth(v) = hops to the root: (jump:16 | dist:16) words updated in place, as in links_stack
    unsigned *jd = (unsigned *)sm.vals;                       // [Nq] scratch in the (not yet used) value array
    int *hist = (int *)(jd + Nq);                             // [Nq] scratch
    for (int v = tid; v < N; v += TPB) { const unsigned r = sm.r16[v]; jd[v] = (r << 16) | (r == (unsigned)v ? 0u : 1u); hist[v] = 0; }
    __syncthreads();
    int act = 1;
    while (act) {
        int loc = 0;
        for (int v = tid; v < N; v += TPB) {
            const unsigned w = jd[v], j = w >> 16;
            const unsigned wj = jd[j];
            if ((wj >> 16) != j) { jd[v] = (wj & 0xFFFF0000u) | ((w + wj) & 0xFFFFu); loc = 1; }
        }
        act = __syncthreads_or(loc);
    }
    int md = 0;
    for (int v = tid; v < N; v += TPB) { const int d = jd[v] & 0xFFFFu; md = max(md, d); atomicAdd(&hist[d], 1); }
    md = __reduce_max_sync(0xffffffffu, md);
    if ((tid & 31) == 0) atomicMax(&S.ibuf[2], md);
    __syncthreads();
    const int maxd = S.ibuf[2];
    __syncthreads();
    if (tid == 0) S.ibuf[2] = 0;
    // exclusive scan of the histogram -> bucket offsets (static shared memory, or global scratch for very deep trees)
    int *boffg = S.P.i[I_EVENTS];
    const bool deep = (maxd + 2 > DEPMAX);
    {
        int carry = 0;
        for (int base = 0; base <= maxd; base += TPB) {
            const int d = base + tid;
            const int h = (d <= maxd) ? hist[d] : 0;
            int tot;
            const int ex = blk_exscan(h, tot, S) + carry;
            if (d <= maxd) { if (deep) boffg[d] = ex; else S.boff[d] = (unsigned short)ex; hist[d] = ex; }
            carry += tot;
        }
        if (tid == 0) { if (deep) boffg[maxd + 1] = N; else S.boff[maxd + 1] = (unsigned short)N; }
    }
    __syncthreads();
    for (int v = tid; v < N; v += TPB) { const int d = jd[v] & 0xFFFFu; sm.order[atomicAdd(&hist[d], 1)] = (unsigned short)v; }
    __syncthreads();
    if (tid == 0) S.ibuf[3] = maxd;
#pragma unroll 4
    for (int v = tid; v < N; v += TPB) sm.vals[v] = c.area[v] * ((v >= c.bounds) ? rain : 0.);
    __syncthreads();
    for (int d = maxd; d >= 0; --d) {
        const int b0 = deep ? boffg[d] : S.boff[d], b1 = deep ? boffg[d + 1] : S.boff[d + 1];
        for (int i = b0 + tid; i < b1; i += TPB) {
            const int v = sm.order[i];
            unsigned dn = sm.lc[v];
            if (dn == NONE16) continue;
            double acc = sm.vals[v];
            for (; dn != NONE16; dn = sm.ps[dn]) acc = acc + sm.vals[dn];
            sm.vals[v] = acc;
        }
        __syncthreads();
    }
    __syncthreads();
    double *Q = S.P.f[F_Q];
#pragma unroll 4
    for (int v = tid; v < N; v += TPB) Q[v] = sm.vals[v];
}

__device__ __noinline__ double flow_cfl(const DevConst &c, Smem &S, unsigned char *dsm, double erod)
{
    BL_ASSUME_SHARED(c, S, dsm);
    const int N = c.N, Nq = c.Nq;
    ScanSm sm = scan_layout(dsm, Nq);
    const double *W = S.P.f[F_FILLH];
    double cfl = 1.e6;
#pragma unroll 2
    for (int d = threadIdx.x; d < N; d += TPB) {
        const int r = sm.r16[d];
        const double q = sm.vals[d];
        if (d == r || !(q > 0.)) continue;
        const double dz = W[d] - W[r];
        if (dz > 0.) {
            const double dx = c.xy[2 * d] - c.xy[2 * r], dy = c.xy[2 * d + 1] - c.xy[2 * r + 1];
            const double dist = sqrt(dx * dx + dy * dy);
            const double tmp = dist / (erod * powx(q, c.m) * powx(dz / dist, c.n - 1.));
            cfl = fmin(tmp, cfl);
        }
    }
    return blk_min(cfl, S);
}

// FLOWalgo.f90:583-1044 (see phase_streampower).  `first`: Q is still in sm.vals (straight after discharge).
__device__ __noinline__ void streampower(const DevConst &c, Smem &S, unsigned char *dsm, double erod, double sea, double dt,
                                         bool first, const int *base1, int nbase1)
{
    BL_ASSUME_SHARED(c, S, dsm);
    const int N = c.N, Nq = c.Nq, tid = threadIdx.x;
    ScanSm sm = scan_layout(dsm, Nq);
    const int *pitID = S.P.i[I_PITID], *stack = S.P.i[I_STACK], *pitDrain = S.P.i[I_PITDRAIN];
    const double *z = S.P.f[F_Z], *W = S.P.f[F_FILLH], *Qg = S.P.f[F_Q], *maxhA = S.P.f[F_MAXH];
    const double *pitVol1 = S.P.f[F_PITVOL];
    double *depo = S.P.f[F_DEPO], *ero = S.P.f[F_ERO], *pdep = S.P.f[F_PDEP], *pitVol = S.P.f[F_PITVOLW];
    long long *pcs = c.phase_cyc + (size_t)blockIdx.x * NPHASE;
    long long ts0 = clock64();
#define STICK(i) do { if (tid == 0) { long long t1_ = clock64(); tick_add(pcs + (i), t1_ - ts0); ts0 = t1_; } } while (0)
    // (1) per-node pre-pass
#pragma unroll 2
    for (int d = tid; d < N; d += TPB) {
        const int r = sm.r16[d];
        const double zd = z[d], zr = z[r], fh = W[d], area = c.area[d];
        if (d + 2 * TPB < N) { pf_l2(z + d + 2 * TPB); pf_l2(W + d + 2 * TPB); pf_l2(maxhA + d + 2 * TPB); pf_l2(pitID + d + 2 * TPB); }
        const double q = first ? sm.vals[d] : Qg[d];
        double dh = 0.95 * (zd - zr);
        if (zd > sea && zr < sea) dh = 0.99 * (zd - sea);
        if (dh < 0.001) dh = 0.;
        const double waterH = fh - zd;
        double SPL = 0., totspl = 0.;
        if (r != d && dh > 0. && waterH == 0. && fh >= sea) {
            const double dx = c.xy[2 * d] - c.xy[2 * r], dy = c.xy[2 * d + 1] - c.xy[2 * r + 1];
            const double dist = sqrt(dx * dx + dy * dy);
            if (dist > 0.) {
                const double slp = dh / dist;
                SPL = -erod * 1. * 1. * powx(q, c.m) * powx(slp, c.n);
                totspl = totspl + SPL;
                if (-totspl * dt > dh) {
                    const double frac = dh / (-totspl * dt);
                    SPL = SPL * frac;
                    totspl = -dh;
                }
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
                const int pp = pitID[d];          // pit under the sea: static first cascade step (:979-991)
                if (pp >= 0 && c.area[pp] > 0. && W[pp] < sea) k = (zd < sea) ? 2 : 4;
            }
            else if (zd <= sea) k = 2;
            else if (maxh > 0. && waterH == 0. && d != r && zd > sea) { k = 3; e = maxh; }
            else if (d == r) k = 2;
            else k = 4;
        } else k = 5;
        sm.kind[d] = (unsigned char)k;
        sm.vals[d] = e;
        depo[d] = 0.;
        ero[d] = 0.;
    }
    for (int v = tid; v < Nq / 32; v += TPB) sm.evbits[v] = 0u;
    const int *posg = S.P.i[I_POS];
    __syncthreads();
    // (2) ordered sediment routing
    {
        // (2) ordered sediment routing over the depth buckets built by discharge()
        const int maxd = S.ibuf[3];
        const int *boffg = S.P.i[I_EVENTS];
        const bool deep = (maxd + 2 > DEPMAX);
        double *vals = sm.vals;
        for (int d = maxd; d >= 0; --d) {
            const int b0 = deep ? boffg[d] : S.boff[d], b1 = deep ? boffg[d + 1] : S.boff[d + 1];
            for (int i = b0 + tid; i < b1; i += TPB) {
                const int cur = sm.order[i];
                double acc = 0. * dt;
                for (unsigned dn = sm.lc[cur]; dn != NONE16; dn = sm.ps[dn]) acc = acc + vals[dn];
                const int k = sm.kind[cur];
                const double own = vals[cur];
                double Qs = 0., erodep = 0.;
                if (k == 0) { erodep = own; Qs = -erodep + acc; }
                else if (k == 1) {
                    if (acc != 0.) {
                        pdep[cur] = acc;
                        if (acc > 0.) { const int ri = N - 1 - posg[cur]; atomicOr(&sm.evbits[ri >> 5], 1u << (ri & 31)); }
                    }
                }
                else if (k == 2) { erodep = acc; }
                else if (k == 3) {
                    const double area = c.area[cur];
                    const double totflx = 0. + acc, maxh = own;
                    if (totflx / area < maxh) erodep = acc;
                    else { const double frac = acc / totflx; erodep = frac * maxh * area; Qs = acc - erodep; }
                } else if (k == 4) Qs = acc;
                if (!(k == 1 && acc != 0.)) {
                    if (erodep < 0.) ero[cur] = 0. + erodep;
                    else if (erodep != 0.) depo[cur] = 0. + erodep;
                }
                vals[cur] = Qs;
            }
            __syncthreads();
        }
    }
    __syncthreads();
    STICK(15);  // prepass + routing
    // (3) lake-entry events in reverse stack order: compact the bitmap, stage (pit, load, donor) per event in the
    // shared-memory regions the routing no longer needs (vals -> loads, lc -> pits, ps -> donors)
    double *ev_dep = sm.vals;
    unsigned short *ev_pit = sm.lc, *ev_donor = sm.ps;
    int nev;
    {
        const int nw = (N + 31) >> 5;
        int wcount[4], tot_mine = 0;         // up to 4 bitmap words per thread (N <= 32768 -> 1024 words)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int wi = tid * 4 + u;
            wcount[u] = (wi < nw && tid * 4 < TPB * 4) ? __popc(sm.evbits[wi]) : 0;
            tot_mine += wcount[u];
        }
        const int off0 = blk_exscan(tot_mine, nev, S);
        unsigned words[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int wi = tid * 4 + u; words[u] = (wi < nw) ? sm.evbits[wi] : 0u; }
        __syncthreads();                     // every thread has its words: vals / lc / ps may be overwritten now
        int off = off0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            unsigned wbits = words[u];
            while (wbits) {
                const int bit = __ffs(wbits) - 1;
                wbits &= wbits - 1;
                ev_donor[off++] = (unsigned short)((tid * 4 + u) * 32 + bit);     // reverse stack index for now
            }
        }
    }
    // the routing is done with sm.kind: it now flags the pits that receive at least one event (most do not)
    unsigned char *has_ev = sm.kind;
    for (int v = tid; v < Nq / 4; v += TPB) reinterpret_cast<unsigned *>(has_ev)[v] = 0u;
    __syncthreads();
    for (int e = tid; e < nev; e += TPB) {       // one event per thread: the dependent global lookups run in parallel
        const int donor = stack[N - 1 - ev_donor[e]];
        const int pit = pitID[donor];
        const bool ok = (pit >= 0 && c.area[pit] > 0.);
        if (ok) has_ev[pit] = 1;
        ev_pit[e] = ok ? (unsigned short)pit : NONE16;
        ev_donor[e] = (unsigned short)donor;
        ev_dep[e] = pdep[donor];
    }
    __syncthreads();
    STICK(12);  // event compaction
    if (nev == 0) return;
    // parallel replay: when no event overflows its own pit (the usual case) events of different pits are
    // independent and events of one pit are applied in order by one thread -> same sums as the serial loop
    int over = 0;
    if (S.ibuf[1])      // some node lies under the sea: an event whose pit is submerged needs the serial path
        for (int e = tid; e < nev; e += TPB) { const int pit = ev_pit[e]; if (pit != NONE16 && W[pit] < sea) over = 1; }
    // one WARP per pit that receives events: the lanes scan the event list 32 entries at a time and the matching events
    // are applied in order (same sums as the serial loop: events of other pits do not touch this pit's deposit)
    int *plist = S.P.i[I_MEMB];
    const int npl = blk_compact(nbase1, plist, [&](int i) { return has_ev[base1[i]] != 0; }, S);
    {
        const int lane = tid & 31, warp = tid >> 5;
        for (int ip = warp; ip < npl; ip += NWARP) {
            const int pit = base1[plist[ip]];
            const double vol = pitVol1[pit];
            const bool self = (pitDrain[pit] == pit);
            double dp = depo[pit];
            bool ovf = false;
            for (int e0 = 0; e0 < nev && !ovf; e0 += 32) {
                const int e = e0 + lane;
                unsigned m = __ballot_sync(0xffffffffu, e < nev && ev_pit[e] == pit);
                while (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    const double pitDep = ev_dep[e0 + b];
                    const double totdist = 0. + pitDep, tmpdist = 0. + dp;
                    if (tmpdist + totdist <= vol || self) dp = dp + pitDep;
                    else { ovf = true; break; }
                }
            }
            if (ovf) over = 1;
            if (lane == 0) S.P.f[F_VAL][pit] = dp;   // staged, committed below when no event overflowed
        }
    }
    over = __syncthreads_or(over);
    STICK(13);  // parallel replay
    if (!over) {
        for (int i = tid; i < nbase1; i += TPB) { const int pit = base1[i]; if (has_ev[pit]) depo[pit] = S.P.f[F_VAL][pit]; }
        __syncthreads();
        return;
    }
    // serial replay, literal (FLOWalgo.f90:967-1034)
    for (int k = tid; k < N; k += TPB) pitVol[k] = pitVol1[k];
    if (tid == 0) S.ibuf[0] = 0;
    __syncthreads();
    if (tid == 0) {
        unsigned st = 0;
        for (int e = 0; e < nev; ++e) {
            if (ev_pit[e] == NONE16) continue;
            const int donor = ev_donor[e];
            const int recvr = sm.r16[donor];
            double pitDep = ev_dep[e];
            double totdist = 0. + pitDep;
            int tmpID = ev_pit[e], nID = tmpID;
            long guard = 0;
            while (totdist > 0.) {
                if (++guard > 1000000 || tmpID < 0) { st |= BL_ST_CASCADE_GUARD; break; }
                const double tmpdist = 0. + depo[tmpID];
                const double pv = pitVol[tmpID];
                if (W[tmpID] < sea) {
                    if (z[donor] < sea) depo[donor] = depo[donor] + pitDep;
                    else { S.ibuf[0] = 1; e = nev; }     // handed back to the routing: literal serial pass below
                    totdist = 0.;
                    nID = recvr;
                } else if (tmpdist + totdist <= pv) {
                    depo[tmpID] = depo[tmpID] + pitDep; totdist = 0.; nID = tmpID;
                } else if (pitDrain[tmpID] == tmpID) {
                    depo[tmpID] = depo[tmpID] + pitDep; totdist = 0.; nID = tmpID;
                } else {
                    if (!(c.flags[tmpID] & 1)) { totdist = 0.; nID = tmpID; }
                    else if (tmpdist == pv) nID = tmpID;
                    else {
                        const double frac = pitDep / totdist;
                        depo[tmpID] = depo[tmpID] + (pv - tmpdist) * frac;
                        pitDep = (totdist - (pv - tmpdist)) * frac;
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
    STICK(14);  // serial replay
    if (S.ibuf[0]) phase_streampower_serial(c, S, erod, sea, dt);
#undef STICK
}

// FLOWalgo.f90:332-388 diffmarine + buildFlux.py:198-247 (see phase_marine_diffusion): the up to 1000 sub-steps run on
// z / fresh-deposit thickness / rate / coefficient held in shared memory (33 bytes per node); the sub-step count is
// data dependent (mindt guard), so the loop stays a block-wide sequence of gather pass + min + update.
template <int KDT>
__device__ __noinline__ void marine_diffusion(const DevConst &c, Smem &S, unsigned char *dsm, double sea, double timestep)
{
    BL_ASSUME_SHARED(c, S, dsm);
    const int N = c.N, Nq = c.Nq, tid = threadIdx.x;
    double *zs = (double *)dsm, *sd = zs + Nq, *dmv = sd + Nq, *cf = dmv + Nq;
    unsigned char *fls = (unsigned char *)(cf + Nq);
    double *z = S.P.f[F_Z], *cumdiff = S.P.f[F_CUMDIFF], *sumdep = S.P.f[F_DTHICK];
    const double maxth = 0.1;
    for (int k = tid; k < N; k += TPB) {                       // diffLinear.sedfluxmarine (diffLinear.py:184-213)
        const double area = c.area[k], zk = z[k];
        const double ac = (area > 0.) ? 1. / area : 0.;
        zs[k] = zk; sd[k] = sumdep[k]; fls[k] = c.flags[k];
        cf[k] = ac * ((zk >= sea) ? 0. : c.CDr);
    }
    __syncthreads();
    double diffstep = timestep;
    int it = 0;
    while (diffstep > 0. && it < 1000) {
        double maxstep = fmin(c.ms, diffstep);
        double mind = maxstep;
        for (int g = tid; g < N; g += TPB) {
            double acc = 0.;
            const unsigned char fg = fls[g];
            const double zg = zs[g];
            if ((fg & 1) && zg < sea) {
                const double dg = sd[g], cg = cf[g];
                Ids<KDT> ids;
                load_ids<KDT>(c, g, ids);
#pragma unroll
                for (int q = 0; q < KDT; ++q) {
                    const unsigned j = ids.get(q);
                    if (j == NONE16) continue;
                    const double zj = zs[j];
                    bool take;
                    if (fls[j] & 1) take = (dg > maxth && zg > zj) || (sd[j] > maxth && zg < zj && zj < sea);
                    else take = (dg > maxth && zg > zj);
                    if (take) {
                        double vo, ed;
                        if (c.ve32) { const float2 ve = *reinterpret_cast<const float2 *>(c.ve32 + ((size_t)g * KDT + q) * 2); vo = ve.x; ed = ve.y; }
                        else { vo = c.vor[(size_t)q * N + g]; ed = c.edge[(size_t)q * N + g]; }
                        const double flx = vo * (zj - zg) / ed;
                        acc = acc + cg * flx;
                    }
                }
                if (acc < 0. && acc * maxstep < -dg) mind = fmin(-dg / acc, mind);
            }
            dmv[g] = (fg & 1) ? acc : 0.;
        }
        mind = blk_min(mind, S);
        maxstep = fmin(mind, maxstep);
        diffstep -= maxstep;
        for (int k = tid; k < N; k += TPB) {
            const double d = dmv[k] * maxstep;
            sd[k] += d; zs[k] += d;
            if (d != 0.) cumdiff[k] += d;              // x + 0 = x bit for bit (cumdiff is never -0)
        }
        __syncthreads();
        ++it;
    }
    for (int k = tid; k < N; k += TPB) { z[k] = zs[k]; sumdep[k] = sd[k]; }
    __syncthreads();
}

// buildFlux.py:193-195, :252-272, :292-315 (see phase_apply) with z staged in shared memory for the gather
template <int KDT>
__device__ __noinline__ void apply(const DevConst &c, Smem &S, unsigned char *dsm, double sea, double timestep)
{
    BL_ASSUME_SHARED(c, S, dsm);
    const int N = c.N, Nq = c.Nq, tid = threadIdx.x;
    double *__restrict__ zs = (double *)dsm;
    double *__restrict__ cd = zs + Nq;
    unsigned char *__restrict__ fls = (unsigned char *)(cd + Nq);          // node flags staged for the neighbour test
    double *__restrict__ z = S.P.f[F_Z];
    double *__restrict__ cumdiff = S.P.f[F_CUMDIFF];
    double *__restrict__ cumhill = S.P.f[F_CUMHILL];
    const double *__restrict__ sedflux = S.P.f[F_SEDFLUX];
    const unsigned char *__restrict__ gflags = c.flags;
    if (c.CDr > 0.) {            // river-borne marine sediment diffusion works on the global arrays
        for (int k = tid; k < N; k += TPB) { const double sf = sedflux[k]; z[k] += sf; cumdiff[k] += sf; }
        __syncthreads();
        // 33 bytes of shared memory per node when the CTA has them, else the global-memory form (same arithmetic)
        if ((size_t)33 * Nq + 64 <= (size_t)S.dsm_bytes) marine_diffusion<KDT>(c, S, dsm, sea, timestep);
        else phase_marine_diffusion(c, S, sea, timestep);
        for (int k = tid; k < N; k += TPB) { zs[k] = z[k]; fls[k] = c.flags[k]; }
    } else {
#pragma unroll 4
        for (int k = tid; k < N; k += TPB) {
            const double sf = sedflux[k];
            zs[k] = z[k] + sf;
            cumdiff[k] += sf;
            fls[k] = gflags[k];
        }
    }
    __syncthreads();
    for (int g = tid; g < N; g += TPB) {
        double acc = 0.;
        const unsigned char fg = fls[g];
        const double zg = zs[g];
        const double c0 = cumdiff[g], h0 = cumhill[g];
        if (fg & 2) {
            Ids<KDT> ids;
            load_ids<KDT>(c, g, ids);
            if (c.ve32) {
                // Voronoi / Delaunay edge lengths are float32-rounded by construction (FVmethod.py:152,236,273):
                // fetched as packed float pairs (half the bytes, 128-bit loads) and widened exactly
                const float4 *ve = reinterpret_cast<const float4 *>(c.ve32 + (size_t)g * KDT * 2);
#pragma unroll
                for (int q = 0; q < KDT / 2; ++q) {
                    const float4 v4 = __ldg(ve + q);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int p = 2 * q + h;
                        if (ids.get(p) == NONE16) continue;
                        const int j = ids.get(p);
                        const double ed = (double)(h ? v4.z : v4.x), di = (double)(h ? v4.w : v4.y), zj = zs[j];
                        const bool bj = fls[j] & 2;
                        if (bj && di > 0.) acc += ed * (zj - zg) / di;
                        if (!bj) { if (zj < zg && di > 0.) acc += ed * (zj - zg) / di; }
                    }
                }
            } else {
#pragma unroll
                for (int p = 0; p < KDT; ++p) {
                    if (ids.get(p) == NONE16) continue;
                    const int j = ids.get(p);
                    const double di = c.edge[(size_t)p * N + g], ed = c.vor[(size_t)p * N + g], zj = zs[j];
                    const bool bj = fls[j] & 2;
                    if (bj && di > 0.) acc += ed * (zj - zg) / di;
                    if (!bj) { if (zj < zg && di > 0.) acc += ed * (zj - zg) / di; }
                }
            }
        }
        const double area = c.area[g];
        const double ac = (area > 0.) ? 1. / area : 0.;
        double coeff = ac * ((zg >= sea) ? c.CDa : c.CDm);
        if (!(fg & 2)) { coeff = 0.; acc = 0.; }
        if (S.diff_out) S.diff_out[g] = acc;
        const double d = coeff * acc * timestep;
        cd[g] = d;
        // the cumulative arrays take the diffusion term here (same thread that added sedflux to cumdiff[g] above)
        if (fg & 1) { cumdiff[g] = c0 + d; cumhill[g] = h0 + d; }
    }
    __syncthreads();
#pragma unroll 4
    for (int k = tid; k < N; k += TPB)
        if (fls[k] & 1) zs[k] += cd[k];
    __syncthreads();
    if (c.btype != 0) {
        const double add = (c.btype == 1) ? -0.1 : (c.btype == 2 ? 0. : 100.);
        for (int k = tid; k < c.bounds; k += TPB) {
            const double zp = zs[c.parent[k]];
            cd[k] = (c.btype == 2) ? zp : zp + add;
        }
        __syncthreads();
        for (int k = tid; k < c.bounds; k += TPB) zs[k] = cd[k];
        __syncthreads();
    }
#pragma unroll 4
    for (int k = tid; k < N; k += TPB) {
        double zn = zs[k];
        if (c.disp) zn += c.disp[k] * timestep;
        z[k] = zn;
    }
    __syncthreads();
}

template <int KDT>
__device__ void time_step(const DevConst &c, Smem &S, unsigned char *dsm)
{
    const int tid = threadIdx.x;
    if (tid == 0) { S.sea = sea_level(c, S.tNow); S.tick0 = clock64(); }
    __syncthreads();
    {
        int any_marine;
        fill_sfd<KDT>(c, S, dsm, S.sea, any_marine);
        if (tid == 0) { S.ibuf[1] = any_marine; S.any_marine = any_marine; }
    }
    BL_TICK(0);
    {
        int nbase1;
        links_stack<KDT>(c, S, dsm, S.P.i[I_RCV1], S.P.i[I_STACK1], S.P.i[I_POS1], S.P.i[I_LA], nbase1, false, false, 0ull);
    }
    BL_TICK(2);
    basins(c, S, dsm, S.P.i[I_LA], S.nbase1);
    BL_TICK(3);
    {
        int nbase;
        links_stack<KDT>(c, S, dsm, S.P.i[I_RCV], S.P.i[I_STACK], S.P.i[I_POS], S.P.i[I_LB], nbase, true, c.shuffle != 0,
                         mix64(S.seed + (unsigned long long)S.step));
    }
    BL_TICK(4);
    drainage(c, S, dsm, S.P.i[I_LA], S.nbase1, S.sea, S.any_marine);
    BL_TICK(5);
    discharge(c, S, dsm, S.rain);
    BL_TICK(6);
    {
        const double flowcfl = flow_cfl(c, S, dsm, S.erod);
        double cfl = fmin(flowcfl, c.hill);
        cfl = py2_floorish(cfl);
        cfl = fmin(cfl, S.tStop - S.tNow);
        cfl = fmax(c.minDT, cfl);
        cfl = fmin(c.maxDT, cfl);
        if (tid == 0) { S.dbuf[2] = cfl; S.newdt = cfl; }
        __syncthreads();
    }
    BL_TICK(7);
    streampower(c, S, dsm, S.erod, S.sea, S.newdt, true, S.P.i[I_LA], S.nbase1);
    BL_TICK(8);
    {
        // getid1 (FLOWalgo.f90:1047-1086): pitVol > 0 only at basin roots, so only those can qualify
        const double *cdepo = S.P.f[F_DEPO], *cero = S.P.f[F_ERO], *pitVol = S.P.f[F_PITVOL];
        const int *pitID = S.P.i[I_PITID], *alld = S.P.i[I_ALLDRAIN], *base1 = S.P.i[I_LA];
        const int nbase1 = S.nbase1;
        const double cfl = S.dbuf[2];
        int nb = 0;
        double minperc = 1.e300;
        for (int i = tid; i < nbase1; i += TPB) {
            const int k = base1[i];
            const double vc = (c.depo == 0) ? cero[k] : cdepo[k] + cero[k];
            const double sumvol = 0. + vc;
            const bool in1 = (sumvol > pitVol[k] && pitVol[k] > 0.);
            const bool in2 = (pitID[k] >= 0 && alld[k] == pitID[k]);
            if (in1 && in2 && (c.flags[k] & 4)) { nb = 1; minperc = fmin(minperc, pitVol[k] / sumvol); }
        }
        nb = __syncthreads_or(nb);
        double newdt = cfl;
        if (nb) { minperc = blk_min(minperc, S); newdt = cfl * minperc; }
        newdt = py2_floorish(newdt);
        newdt = fmax(c.minDT, newdt);
        if (tid == 0) S.newdt = newdt;
        __syncthreads();
        if (newdt < cfl) streampower(c, S, dsm, S.erod, S.sea, S.newdt, false, S.P.i[I_LA], S.nbase1);
    }
    BL_TICK(9);
    phase_erodep(c, S, S.sea);
    BL_TICK(10);
    if (c.CDr > 0.) apply<KDT>(c, S, dsm, S.sea, S.newdt);
    else fastc::apply_c<KDT>(c, S, dsm, S.sea, S.newdt);
    BL_TICK(11);
    if (tid == 0) { S.tNow += S.newdt; S.last_dt = S.newdt; S.step += 1; }
    __syncthreads();
}

} // namespace fast

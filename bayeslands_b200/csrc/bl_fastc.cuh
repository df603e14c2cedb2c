// bl_fastc.cuh - "compact" shared-memory time step (included by bl_device.cuh after bl_fast.cuh).
//
// bl_fast.cuh keeps 17 bytes of shared memory per node, so a 12.7 k-node landscape owns an SM: one CTA, and nothing to
// hide its barrier waits (38 % of the warp samples in round 1).  The phases below compute the same values in the same
// fp64 operation order with at most 8 bytes per node (+ bitmaps and 4 KB of lists) in shared memory, so that TWO 512-thread
// CTAs - two landscapes - are resident per SM and one runs while the other waits:
//   fill            W fp64 in shared memory, pending set as one BIT per node (woken with red.shared.or), z of the evaluated
//                   node and its id row from L2; every warp sweeps, one batch of 8 bitmap words per warp and level; the
//                   initial pending set is the ring of nodes next to a ghost node (the rest is woken as the front arrives)
//   receivers       two passes over the same 8N bytes (W, then z brought in by a TMA bulk copy); they also leave W(d) - W(rcv),
//                   the receiver distance, z(rcv) and 16-bit copies of both receiver arrays for the later passes
//   stack order     receivers (2N, bulk copy) for the link construction (donor slots as a bit mask per node, two passes); the
//                   Euler tour as 2N u16 successors (4N, written to global scratch, bulk copy back), ranked in O(N) work:
//                   dense splitters, one walk per sublist, pointer jumping over the splitters only, a dense finishing pass;
//                   the same ranking yields every node's depth in the receiver tree
//   basins          roots from the stack segments (marks propagated over stack positions, 2N), then the stack-ordered
//                   volume terms (8N) in the same bytes; long segments summed by a warp that skips the zero terms
//   drainage        fast::drainage with its two u16 tables staged by bulk copies
//   discharge / stream power
//                   node values fp64 indexed by BREADTH-FIRST POSITION: the depth-first stack is stably sorted by depth,
//                   which puts the donors of a node next to each other (ascending id), so a node sums the contiguous
//                   run vals[fc .. fc+cnt) backwards - no donor links, no bucket order, no receiver ids in shared memory.
//                   One 16-byte global record per position (first child, donor count, rule, node id, fp64 payload) is
//                   fetched one item ahead; the flow CFL condition rides on the pass that writes Q by node id
//   diffusion       z fp64 (8N); the "neighbour is an inside node" test comes from a per-node bit mask (mesh constant);
//                   the ghost update is folded into the same pass
// Streaming passes over per-proposal arrays issue every load before the first branch, launder the loop index (so that the
// compiler keeps no per-array running address under the 64-register cap) and request the next batch's lines into L2.
// Selected by bl_ctx_create (DevConst::compact) for non-marine configurations on colour-major meshes when there are more
// landscapes than SMs; parity: tests/test_gpu_compact.py.

namespace fastc {

using namespace fast;

__device__ __forceinline__ unsigned lds_u32(unsigned a)
{
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void red_or_if(bool p, unsigned a, unsigned v)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %0, 0;\n\t@q red.shared.or.b32 [%1], %2;\n\t}" ::"r"((unsigned)p), "r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void red_and(unsigned a, unsigned v) { asm volatile("red.shared.and.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// ---- TMA bulk copy global -> shared memory (cp.async.bulk, completion on an mbarrier) -------------------------------
// Whole arrays that a phase needs in shared memory unchanged (the elevations for the receiver search, the tour successors
// for the list ranking) are fetched by the copy engine: one thread posts the transfers, nobody spends registers or issue
// slots on a load / store loop, and the block waits on the mbarrier's transaction count.  Called by every thread of the
// CTA from uniform code, AFTER a __syncthreads() that retired the previous users of the destination bytes; `bytes` is a
// multiple of 16, both addresses are 16-byte aligned.
__device__ __forceinline__ void tma_to_smem(Smem &S, void *dst_smem, const void *src_gmem, unsigned bytes)
{
    const unsigned mbar = (unsigned)__cvta_generic_to_shared(&S.mbar);
    const unsigned parity = S.tma_phase;
    if (threadIdx.x == 0) {
        // generic-proxy accesses of the destination (ordered by the caller's barrier) before the async-proxy writes
        asm volatile("fence.proxy.async;" ::: "memory");    // (all state spaces: the source was written by this CTA too)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
        const unsigned dst = (unsigned)__cvta_generic_to_shared(dst_smem);
        const char *src = (const char *)src_gmem;
        for (unsigned off = 0; off < bytes; off += 32768u) {
            const unsigned n = min(32768u, bytes - off);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst + off), "l"(src + off), "r"(n), "r"(mbar) : "memory");
        }
    }
    asm volatile("{\n\t.reg .pred p;\n\tTMA_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra TMA_DONE;\n\tbra TMA_WAIT;\n\tTMA_DONE:\n\t}"
                 ::"r"(mbar), "r"(parity) : "memory");
    __syncthreads();
    if (threadIdx.x == 0) S.tma_phase = parity ^ 1u;
    __syncthreads();                                     // the next call (possibly the very next statement) reads the new parity
}

// dynamic shared memory of the compact layout: 8 bytes per node, one bit per node, the fill's per-warp slot lists
__device__ __forceinline__ unsigned char *tail_of(unsigned char *dsm, int Nq) { return dsm + (size_t)8 * Nq + Nq / 8; }

// ---- depression filling (PDalgo.f90:19-91), literal Gauss-Seidel sweep by level, see fast::fill_bitmap -------------
template <int KDT, bool PF>
__device__ __forceinline__ void fill_eval_c(const unsigned short *__restrict__ ngb16f, const double *__restrict__ zg, int k,
                                            unsigned aW, unsigned aB, double sea, double eps, double TH, int &loc)
{
    Ids<KDT> ids;        // empty slots hold the dummy node N, whose W is 2e6 (DevConst::ngb16f)
    {
        const uint2 *src = reinterpret_cast<const uint2 *>(ngb16f + (size_t)k * KDT);
#pragma unroll
        for (int q = 0; q < KDT / 4; ++q) { const uint2 v = __ldg(src + q); ids.w[2 * q] = v.x; ids.w[2 * q + 1] = v.y; }
    }
    const double zk = zg[k];                     // constant during the fill; requested together with the id row
    const unsigned ak = aW + 8u * (unsigned)k;
    const double wk = lds_f64(ak);
    double hmin = 2.e6;
    double wj[KDT];
#pragma unroll
    for (int p = 0; p < KDT; ++p) wj[p] = lds_f64(aW + 8u * ids.get(p));
#pragma unroll
    for (int p0 = 0; p0 < KDT; p0 += 4) hmin = fmin(hmin, fmin(fmin(wj[p0], wj[p0 + 1]), fmin(wj[p0 + 2], wj[p0 + 3])));
    const double he = hmin + eps;
    const bool wet = wk > zk, dry = zk >= he, lower = wk > he;
    const double floor_ = (zk >= sea) ? zk : sea;
    const bool capped = (he - floor_) > TH;
    const double nlow = capped ? floor_ + TH : he;
    const double nw = wet ? (dry ? zk : (lower ? nlow : wk)) : wk;
    if (wet && !dry && lower && !capped) loc = 1;
    if (nw != wk) {
        sts_f64(ak, nw);
        const double lim = nw + eps;             // wake only the neighbours this change can move (fast::fill_eval)
        if (KDT > 8) {
#pragma unroll
            for (int p = 0; p < KDT; ++p) wj[p] = lds_f64(aW + 8u * ids.get(p));
        }
#pragma unroll
        for (int p = 0; p < KDT; ++p) {
            const unsigned j = ids.get(p);
            const bool wake = wj[p] > lim;
            red_or_if(wake, aB + 4u * (j >> 5), 1u << (j & 31));
            if (PF && wake) {          // the woken node is evaluated a round or more from now: its z and id row on the way
                asm volatile("prefetch.global.L1 [%0];" ::"l"(zg + j));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(ngb16f + (size_t)j * KDT));
            }
        }
    }
}

// bits of the interior nodes [nb, N) within the bitmap word that starts at node lo
__device__ __forceinline__ unsigned interior_mask(int lo, int nb, int N)
{
    unsigned m = 0u;
    if (lo + 32 > nb && lo < N) {
        m = 0xffffffffu;
        if (lo < nb) m &= 0xffffffffu << (nb - lo);
        if (lo + 32 > N) m &= 0xffffffffu >> (lo + 32 - N);
    }
    return m;
}

template <int KDT>
__device__ __noinline__ void fill_c(const DevConst &c, Smem &S, unsigned char *dsm, double sea)
{
    BL_ASSUME_SHARED(c, S, dsm);
    const int N = c.N, Nq = c.Nq, tid = threadIdx.x, nb = c.bounds;
    const int lane = tid & 31, warp = tid >> 5;
#ifndef BL_FILLC_ALL
#define BL_FILLC_ALL 1
#endif
#ifndef BL_FILLC_PF
#define BL_FILLC_PF 0
#endif
    // sweeping warps: with BL_FILLC_ALL every warp of a (<= 512-thread) CTA sweeps, so a thin front is spread over twice as
    // many warps; a batch is 8 bitmap words per warp (one batch covers a level of up to NWF * 256 nodes); the slot lists take
    // NWF * 256 bytes (compact_smem_need in bl_engine.cu)
    constexpr bool ALLW = BL_FILLC_ALL && (NWARP <= 16);
    constexpr int NWF = ALLW ? NWARP : ((NWARP >= 2) ? NWARP / 2 : 1);
    constexpr int BW = 8;
    const double *__restrict__ zg = S.P.f[F_Z];
    {
        double *W = (double *)dsm;
        unsigned *bits = (unsigned *)(dsm + (size_t)8 * Nq);
#pragma unroll 4
        for (int i = tid; i < N; i += TPB) W[i] = (i < nb) ? zg[i] : 1.e6;
        if (tid == 0) W[N] = 2.e6;                          // the dummy neighbour of empty slots
        // initially pending: the interior nodes next to a ghost node (DevConst::pend0).  Everything else still has all
        // its neighbours at 1e6, where an evaluation changes nothing and raises no flag; it is woken like any other
        // node when a neighbour drops, so the sweeps see the same effective evaluations as the literal loop.
        for (int w = tid; w < Nq / 32; w += TPB) bits[w] = ((c.dbg & 128) ? 0xffffffffu : c.pend0[w]) & interior_mask(w * 32, nb, N);
    }
    __syncthreads();
    const unsigned aW = s32(dsm), aB = aW + 8u * (unsigned)Nq;
    const unsigned aL = aB + (unsigned)(Nq / 8) + (unsigned)warp * (unsigned)(BW * 32);   // [NWF][BW * 32] slot list of this warp
    const unsigned short *ngb16f = c.ngb16f;
    const double eps = c.eps, TH = c.TH;
    const int nlev = c.nlev;
    const unsigned lt = (1u << lane) - 1u;
    const bool diag = (c.dbg & 4) && tid == 0;          // diagnostics: cycles of warp 0 per round (scan / evaluation / barrier)
    long long *pcd = c.phase_cyc + (size_t)blockIdx.x * NPHASE;
    long long tq = diag ? clock64() : 0;
#define FDIAG(i) do { if (diag) { const long long t_ = clock64(); tick_add(pcd + (i), t_ - tq); tq = t_; } } while (0)
    int change = 1;
    while (change) {
        int loc = 0;
        for (int l = 0; l < nlev; ++l) {
            const int n0 = nb + S.lev_off[l], n1 = nb + S.lev_off[l + 1];     // node id range of the level
            if (diag) tick_add(pcd + 32, 1);
            if (warp < NWF && n1 > n0) {
                const int whi = (n1 - 1) >> 5;
                // a batch = BW bitmap words of this warp (word w0 + NWF * t belongs to lane t): BW * 32 candidate ids
                for (int w0 = (n0 >> 5) + warp; w0 <= whi; w0 += BW * NWF) {
                    unsigned mine = 0u;
                    const int wi = w0 + NWF * lane;
                    if (lane < BW && wi <= whi) {
                        const int lo = wi * 32;
                        unsigned m = 0xffffffffu;            // bits of other levels in the boundary words are not ours
                        if (lo < n0) m &= 0xffffffffu << (n0 - lo);
                        if (lo + 32 > n1) m &= 0xffffffffu >> (lo + 32 - n1);
                        const unsigned a = aB + 4u * (unsigned)wi;
                        mine = lds_u32(a) & m;               // nobody sets a bit of this level during its own pass
                        if (mine) {
                            red_and(a, ~mine);
#if BL_FILLC_PF
                            // the elevations of these 32 nodes (two lines): requested now, needed by the evaluation below
                            // (the id rows - four more lines per word - are not worth their instructions: measured)
                            const char *pz = (const char *)(zg + lo);
                            asm volatile("prefetch.global.L1 [%0];" ::"l"(pz));
                            asm volatile("prefetch.global.L1 [%0];" ::"l"(pz + 128));
#endif
                        }
                    }
                    if (__ballot_sync(0xffffffffu, mine != 0u) == 0u) continue;     // nothing pending in this warp's words
                    int off = 0;
#pragma unroll
                    for (int t = 0; t < BW; ++t) {
                        const unsigned wb = __shfl_sync(0xffffffffu, mine, t);
                        if (wb) {
                            sts_u8_if((wb >> lane) & 1u, aL + (unsigned)(off + __popc(wb & lt)), (unsigned)(t * 32 + lane));
                            off += __popc(wb);
                        }
                    }
                    __syncwarp();
                    FDIAG(34);
                    if (diag) tick_add(pcd + 33, off);
                    for (int i = lane; i < off; i += 32) {
                        const int s = (int)lds_u8(aL + (unsigned)i);
                        fill_eval_c<KDT, false>(ngb16f, zg, (w0 + NWF * (s >> 5)) * 32 + (s & 31), aW, aB, sea, eps, TH, loc);
                    }
                    __syncwarp();
                    FDIAG(35);
                }
            }
            if (l + 1 < nlev) __syncthreads();
            FDIAG(36);
        }
        change = __syncthreads_or(loc);
        FDIAG(36);
    }
#undef FDIAG
}

// sfd.c:55-153: receivers on the filled surface (W is still in shared memory), then on the real one (z staged in the
// same bytes).  Same slot order and strict comparisons as fast::fill_receivers.
template <int KDT>
__device__ __noinline__ void receivers_c(const DevConst &c, Smem &S, unsigned char *dsm, double sea, int &any_marine)
{
    BL_ASSUME_SHARED(c, S, dsm);
    const int N = c.N, tid = threadIdx.x;
    double *A = (double *)dsm;
    int *__restrict__ rcv = S.P.i[I_RCV], *__restrict__ rcv1 = S.P.i[I_RCV1];
    double *__restrict__ maxh = S.P.f[F_MAXH], *__restrict__ fillH = S.P.f[F_FILLH];
    const double *__restrict__ zg = S.P.f[F_Z];
    // by-products for the CFL and stream-power passes (same expressions as there, evaluated once per step while the
    // surfaces are in shared memory): W(d) - W(rcv), the distance to the receiver, z(rcv)
    double *__restrict__ dzw = S.P.f[F_CAPH], *__restrict__ distr = S.P.f[F_E], *__restrict__ zrA = S.P.f[F_QS];
    // 16-bit copies of both receiver arrays: links_stack_c brings them into shared memory with one bulk copy each
    unsigned short *__restrict__ rcv16w = (unsigned short *)S.P.i[I_ROOT1], *__restrict__ rcv116w = rcv16w + np_of(N);
    const double2 *__restrict__ xy2 = reinterpret_cast<const double2 *>(c.xy);
    int mar = 0;
#pragma unroll 2
    for (int g = tid; g < N; g += TPB) {
        Ids<KDT> ids;
        load_ids<KDT>(c, g, ids);
        const double2 xg = xy2[g];
        const double wg = A[g];
        int low = g;
        double wl = wg;
#pragma unroll
        for (int p = 0; p < KDT; ++p) {
            if (ids.get(p) == NONE16) continue;
            const int j = ids.get(p);
            const double wj = A[j];
            if (wj < wl) { low = j; wl = wj; }
        }
        const double2 xr = xy2[low];
        rcv[g] = low;
        rcv16w[g] = (unsigned short)low;
        fillH[g] = wg;
        dzw[g] = wg - wl;
        const double dx = xg.x - xr.x, dy = xg.y - xr.y;
        distr[g] = sqrt(dx * dx + dy * dy);
        mar |= (wg < sea);
    }
    __syncthreads();
    {   // z replaces W in the same bytes: one bulk copy (the odd last element, if any, by hand)
        const unsigned bytes = ((unsigned)N * 8u) & ~15u;
        if (tid == 0 && (N & 1)) A[N - 1] = zg[N - 1];
        tma_to_smem(S, A, zg, bytes);
    }
#pragma unroll 2
    for (int g = tid; g < N; g += TPB) {
        Ids<KDT> ids;
        load_ids<KDT>(c, g, ids);
        const int r = rcv[g];
        const double z0 = A[g];
        int low1 = g;
        double zl = z0, dH = 1.e6;
#pragma unroll
        for (int p = 0; p < KDT; ++p) {
            if (ids.get(p) == NONE16) continue;
            const int j = ids.get(p);
            const double zj = A[j];
            if (zj < zl) { low1 = j; zl = zj; }
            const double dh = zj - z0;
            if (dh >= 0. && dh < dH) dH = dh;
        }
        rcv1[g] = low1;
        rcv116w[g] = (unsigned short)low1;
        if (dH > 9.99e5) dH = 0.;
        maxh[g] = dH;
        zrA[g] = A[r];
    }
    any_marine = __syncthreads_or(mar);
}

// ---- stack order (FLWnetwork.f90:21-71, FLOWstack.f90:24-40) ------------------------------------------------------------
// Same Euler tour as fast::links_stack (enter -> first child, exit -> next sibling / parent's exit / next root), ranked in
// O(2N) work instead of ~15 pointer-jumping rounds over all 2N elements (Helman-JaJa list ranking):
//  * an element's weight is its parity (enter events are the even elements), so the tour is 2N u16 successors (4N bytes);
//    they are written to global scratch while the receivers (2N bytes) sit in shared memory for the link construction
//    and come back with one coalesced copy;
//  * every element whose index is a multiple of 2^sh, plus the head of the tour, is a splitter; a thread walks from its
//    splitter to the next one counting enter events and elements (a few steps: the splitters are dense because the
//    other half of the shared memory holds their table), the splitters are ranked by pointer jumping on one 64-bit word
//    each (successor | enter events to the end | elements to the end), a second walk hands every enter event its rank;
//  * counting all elements as well gives the tour index, and with it the depth of a node in the receiver tree:
//        depth(v) = 2 * (#enter events before enter(v)) - (#elements before enter(v))
//    which discharge_c needs for its breadth-first layout (filled surface only).
template <int KDT>
__device__ __noinline__ void links_stack_c(const DevConst &c, Smem &S, unsigned char *dsm, const int *rcvg, int *stack, int *pos,
                                           int *baselist, int &nbase, bool filled, bool shuffle, unsigned long long seedstep)
{
    BL_ASSUME_SHARED(c, S, dsm);
    const int N = c.N, Nq = c.Nq, tid = threadIdx.x;
    unsigned short *r16 = (unsigned short *)dsm;                                  // [Nq]
    unsigned long long *K = (unsigned long long *)(dsm + (size_t)2 * Nq + 64);    // shuffle keys behind the receivers
    const int kcap = (int)(((size_t)6 * Nq - 128) / 8);
    unsigned short *nxg = (unsigned short *)S.P.euA;                              // [2N] successors, global scratch
    SubTick stl = sub_tick_begin(c);
    {   // receivers as u16 (written by receivers_c next to the int arrays): one bulk copy, the last < 8 entries by hand
        const unsigned short *r16g = (const unsigned short *)S.P.i[I_ROOT1] + (filled ? 0 : np_of(N));
        const unsigned bytes = ((unsigned)N * 2u) & ~15u;
        if (tid < 8) { const int v = (int)(bytes / 2u) + tid; if (v < N) r16[v] = r16g[v]; }
        tma_to_smem(S, r16, r16g, bytes);
    }
    (void)rcvg;
    unsigned char *cnt8 = (unsigned char *)S.P.i[I_CNT];
    unsigned *rootbits = (unsigned *)(dsm + (size_t)8 * Nq);     // one bit per node: base levels (self-receivers)
    for (int w = tid; w < Nq / 32; w += TPB) rootbits[w] = 0u;
    __syncthreads();
    // pass 1: the donors of v are the neighbours whose receiver is v: their slots as a bit mask (kept in shared memory for
    // pass 2), their number, the first (lowest-id) one = the successor of enter(v)
    unsigned char *dmask = dsm + (size_t)2 * Nq;                 // [Nq] donor slots of every node (KDT <= 8 bits used; wider rows: see below)
#pragma unroll 2
    for (int v = tid; v < N; v += TPB) {
        Ids<KDT> ids;
        load_ids<KDT>(c, v, ids);
        int f = NONE16, n = 0;
        unsigned dm = 0u;
#pragma unroll
        for (int p = 0; p < KDT; ++p) {
            if (ids.get(p) == NONE16) continue;
            const int w = ids.get(p);
            if (r16[w] == v) { ++n; if (w < f) f = w; dm |= 1u << p; }
        }
        if (KDT <= 8) dmask[v] = (unsigned char)dm;
        nxg[2 * v] = (unsigned short)((f != NONE16) ? 2 * f : 2 * v + 1);
        if (r16[v] == v) atomicOr(&rootbits[v >> 5], 1u << (v & 31));        // a root's exit is linked to the next root below
        if (filled) cnt8[v] = (unsigned char)n;              // donor counts: the breadth-first layout of discharge_c
    }
    __syncthreads();
    // pass 2: the successor of exit(v) is the next donor of v's receiver after v (read off the receiver's mask), else the
    // receiver's exit
#pragma unroll 4
    for (int v = tid; v < N; v += TPB) {
        const int r = r16[v];
        if (r == v) continue;
        Ids<KDT> jds;
        load_ids<KDT>(c, r, jds);
        int nx = NONE16;
        if (KDT <= 8) {
            const unsigned dm = dmask[r];
#pragma unroll
            for (int p = 0; p < KDT; ++p) {
                const int w = jds.get(p);
                if (((dm >> p) & 1u) && w > v && w < nx) nx = w;
            }
        } else {
#pragma unroll
            for (int p = 0; p < KDT; ++p) {
                if (jds.get(p) == NONE16) continue;
                const int w = jds.get(p);
                if (w > v && w < nx && r16[w] == r) nx = w;
            }
        }
        nxg[2 * v + 1] = (unsigned short)((nx != NONE16) ? 2 * nx : 2 * r + 1);
    }
    __syncthreads();
    sub_tick(stl, 28);      // receivers to shared memory + link construction
    nbase = blk_compact_bits(rootbits, N, baselist, S);       // base levels in ascending id
    if (tid == 0) { if (filled) S.nbase = nbase; else S.nbase1 = nbase; }
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
        nxg[2 * root + 1] = (unsigned short)((nxt != NONE_I) ? (unsigned)(2 * nxt) : END16);
        if (!filled) nsroot[root] = nxt;
    }
    __syncthreads();
    unsigned short *nx16 = (unsigned short *)dsm;                    // [2N] successor of every tour element
    const int M = 2 * N;
    {   // 2N u16 successors = 4N bytes, bulk copy (the tail of less than 16 bytes by hand)
        const unsigned bytes = ((unsigned)N * 4u) & ~15u;
        const unsigned *src = reinterpret_cast<const unsigned *>(nxg);
        unsigned *dst = reinterpret_cast<unsigned *>(nx16);
        if (tid < 4) { const int e = (int)(bytes / 4u) + tid; if (e < N) dst[e] = src[e]; }
        tma_to_smem(S, nx16, nxg, bytes);
    }
    sub_tick(stl, 29);      // base levels, shuffle, tour back into shared memory
    unsigned long long *T = (unsigned long long *)(dsm + (size_t)4 * Nq);   // [<= Nq / 2] splitter table
    const int sh = 3 + ((c.dbg >> 5) & 3);
    const int S_reg = ((M - 1) >> sh) + 1;                           // (M >> 3) + 2 <= Nq / 2 always
    const unsigned smask = (1u << sh) - 1u;
    const unsigned head = 2u * (unsigned)baselist[0];
    const bool head_extra = (head & smask) != 0u;
    const int S_tot = S_reg + (head_extra ? 1 : 0);
    // Both walks are ONE loop per lane over all its sublists (a lane that finishes a sublist starts its next one in the
    // same iteration), so the lanes of a warp stay busy until their whole share is done instead of idling at the end of
    // every sublist behind the longest one of the 32.
    // The first walk also leaves, for every enter event, (its sublist : 12 | enter events before it in the sublist : 10 |
    // elements before it : 10) in global scratch, so that the ranks can be finished by a dense pass over the nodes
    // instead of a second divergent walk; a sublist of 1024 elements or more (never seen: the splitters are 8 apart in
    // index space) raises S.ibuf[7] and the second walk below takes over.
    unsigned *infog = (unsigned *)S.P.euB;                           // [N]
    if (tid == 0) S.ibuf[7] = 0;
    __syncthreads();
    {
        int sp = tid;
        unsigned e = 0u, ce = 0u, ca = 0u;
        bool run = sp < S_tot;
        bool ovf = false;
        if (run) e = (sp < S_reg) ? ((unsigned)sp << sh) : head;
        while (run) {
            const unsigned nx = nx16[e];
            if (!(e & 1u)) infog[e >> 1] = ((unsigned)sp << 20) | (ce << 10) | ca;
            ovf |= (ca >= 1023u);
            ce += (~e) & 1u;
            ++ca;
            const bool last = (nx == END16), split = !last && ((nx & smask) == 0u || nx == head);
            if (last || split) {
                const unsigned t = last ? END16 : ((head_extra && nx == head) ? (unsigned)S_reg : (nx >> sh));
                T[sp] = (unsigned long long)t | ((unsigned long long)ce << 16) | ((unsigned long long)ca << 32);
                sp += TPB;
                run = sp < S_tot;
                ce = 0u; ca = 0u;
                e = (sp < S_reg) ? ((unsigned)sp << sh) : head;
            } else e = nx;
        }
        if (ovf) S.ibuf[7] = 1;
    }
    __syncthreads();
    int active = 1;
    while (active) {            // suffix sums over the splitter chain; a word is read and written as one 64-bit access
        int loc = 0;
        for (int sp = tid; sp < S_tot; sp += TPB) {
            const unsigned long long a = T[sp];
            const unsigned na = (unsigned)(a & 0xFFFFull);
            if (na != END16) {
                const unsigned long long a2 = T[na];
                T[sp] = (a2 & 0xFFFFull) | ((((a >> 16) + (a2 >> 16)) & 0xFFFFull) << 16) | (((a >> 32) + (a2 >> 32)) << 32);
                loc = 1;
            }
        }
        active = __syncthreads_or(loc);
    }
    sub_tick(stl, 30);      // first walk + splitter ranking
    unsigned short *depg = (unsigned short *)S.P.i[I_FC];            // depth by stack position (filled surface: discharge_c)
    int mdep = 0;
    if (!S.ibuf[7] && S_tot <= 4096 && !(c.dbg & 256)) {
        // dense finish: rank of enter(v) = (enter events from its sublist's splitter to the end) - (those before it).
        // The permutation (stack[p] = v, depth by position) is scattered in SHARED memory - the successors are no longer
        // needed, the splitter table stays where it is - and leaves with coalesced stores.
        unsigned short *stk_s = (unsigned short *)dsm;                        // [N] node at stack position p   (over nx16)
        unsigned short *dep_s = (unsigned short *)(dsm + (size_t)2 * Nq);     // [N] depth at stack position p  (over nx16)
#pragma unroll 4
        for (int v = tid; v < N; v += TPB) {
            const unsigned info = infog[v];
            const unsigned long long a = T[info >> 20];
            const int re = (int)((a >> 16) & 0xFFFFull) - (int)((info >> 10) & 0x3FFu);
            const int ra = (int)(a >> 32) - (int)(info & 0x3FFu);
            const int p = N - re;
            pos[v] = p;
            stk_s[p] = (unsigned short)v;
            if (filled) { const int dep = 2 * p - (M - ra); dep_s[p] = (unsigned short)dep; mdep = max(mdep, dep); }
        }
        __syncthreads();
#pragma unroll 4
        for (int p = tid; p < N; p += TPB) {
            stack[p] = stk_s[p];
            if (filled) depg[p] = dep_s[p];
        }
    } else {
        int sp = tid;
        unsigned e = 0u, re = 0u, ra = 0u;
        bool run = sp < S_tot;
        if (run) {
            e = (sp < S_reg) ? ((unsigned)sp << sh) : head;
            const unsigned long long a = T[sp];
            re = (unsigned)((a >> 16) & 0xFFFFull); ra = (unsigned)(a >> 32);
        }
        while (run) {
            const unsigned nx = nx16[e];
            if (!(e & 1u)) {
                const int v = (int)(e >> 1), p = N - (int)re;
                pos[v] = p;
                stack[p] = v;
                if (filled) { const int dep = 2 * p - (M - (int)ra); depg[p] = (unsigned short)dep; mdep = max(mdep, dep); }
                --re;
            }
            --ra;
            if (nx == END16 || (nx & smask) == 0u || nx == head) {
                sp += TPB;
                run = sp < S_tot;
                if (run) {
                    e = (sp < S_reg) ? ((unsigned)sp << sh) : head;
                    const unsigned long long a = T[sp];
                    re = (unsigned)((a >> 16) & 0xFFFFull); ra = (unsigned)(a >> 32);
                }
            } else e = nx;
        }
    }
    if (filled) {           // depth of the receiver forest, left in S.ibuf[2] for discharge_c
        mdep = __reduce_max_sync(0xffffffffu, mdep);
        if ((tid & 31) == 0) atomicMax(&S.ibuf[2], mdep);
    }
    __syncthreads();
    sub_tick(stl, 31);      // second walk
}

// ---- pit ids and volumes (FLOWalgo.f90:130-165), see fast::basins ---------------------------------------------------
// Root ids (2N) and the stack-ordered volume terms (8N) share the same bytes.  A basin's volume is the sequential sum
// over its stack segment; a dry node contributes +0.0, which never changes the running sum (it starts at -1 and only
// non-negative terms are added, so it is never -0), hence long segments are read 32 positions at a time by one warp
// and only the non-zero terms are added, in order.
__device__ __noinline__ void basins_c(const DevConst &c, Smem &S, unsigned char *dsm, const int *base1, int nbase1)
{
    BL_ASSUME_SHARED(c, S, dsm);
    const int N = c.N, Nq = c.Nq, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned short *root16 = (unsigned short *)dsm;          // [Nq]
    double *val = (double *)dsm;                             // [Nq], after the roots are consumed
    unsigned short *queue = (unsigned short *)tail_of(dsm, Nq);   // (lo, hi, root) of the long segments
    constexpr int QCAP = (NWARP / 2) * 256 / 6;
    const double *__restrict__ z = S.P.f[F_Z], *__restrict__ W = S.P.f[F_FILLH];
    const int *__restrict__ pos1 = S.P.i[I_POS1], *__restrict__ ns1 = S.P.i[I_NS];
    int *__restrict__ pitID = S.P.i[I_PITID];
    double *__restrict__ pitVol = S.P.f[F_PITVOL];
    unsigned short *__restrict__ pid16w = (unsigned short *)S.P.i[I_FC] + np_of(N);
    // The root of a node is the base level whose stack segment holds it (the real-surface stack lists tree after tree in
    // base order): every base marks the position where its segment starts, the marks are propagated to the right
    // (thread-local pass over a contiguous chunk, block scan of the chunks' last marks, second pass), and a node looks
    // its root up at its own stack position.  O(N) work instead of ~12 pointer-jumping rounds over the receivers.
    unsigned short *rootpos = root16;                        // [N] by stack position
    for (int p = tid; p < N; p += TPB) rootpos[p] = NONE16;
    if (tid == 0) S.ibuf[4] = 0;
    __syncthreads();
    for (int i = tid; i < nbase1; i += TPB) { const int root = base1[i]; rootpos[pos1[root]] = (unsigned short)root; }
    __syncthreads();
    {
        const int C = (N + TPB - 1) / TPB;
        const int p0 = min(N, tid * C), p1 = min(N, p0 + C);
        unsigned last = NONE16;
        for (int p = p0; p < p1; ++p) { const unsigned m = rootpos[p]; if (m != NONE16) last = m; }
        // inclusive scan of "latest mark" over the threads (warp shuffles, then the warps' totals through S.scan)
        unsigned inc = last;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o && inc == NONE16) inc = t;
        }
        if (lane == 31) S.scan[warp] = (int)inc;
        __syncthreads();
        unsigned carry = NONE16;                             // latest mark of all earlier warps
        for (int w = 0; w < warp; ++w) { const unsigned t = (unsigned)S.scan[w]; if (t != NONE16) carry = t; }
        unsigned prev = __shfl_up_sync(0xffffffffu, inc, 1); // latest mark up to the previous lane
        if (lane == 0) prev = NONE16;
        unsigned cur = (prev != NONE16) ? prev : carry;
        for (int p = p0; p < p1; ++p) { const unsigned m = rootpos[p]; if (m != NONE16) cur = m; else rootpos[p] = (unsigned short)cur; }
    }
    __syncthreads();
#pragma unroll 4
    for (int v = tid; v < N; v += TPB) {
        const bool wet = W[v] > z[v];
        if (v + 2 * TPB < N) { pf_l2(W + v + 2 * TPB); pf_l2(z + v + 2 * TPB); pf_l2(pos1 + v + 2 * TPB); }
        const int root = rootpos[pos1[v]];
        const bool in = wet || root == v;
        pitID[v] = in ? root : -1;
        pid16w[v] = in ? (unsigned short)root : NONE16;      // 16-bit copy: drainage_c stages it with a bulk copy
        pitVol[v] = -1.0;
    }
    __syncthreads();
#pragma unroll 4
    for (int v = tid; v < N; v += TPB) {
        const double wv = W[v], zv = z[v];
        if (v + 2 * TPB < N) { pf_l2(W + v + 2 * TPB); pf_l2(z + v + 2 * TPB); pf_l2(pos1 + v + 2 * TPB); }
        val[pos1[v]] = (wv > zv) ? (wv - zv) * c.area[v] : 0.0;
    }
    __syncthreads();
    for (int i = tid; i < nbase1; i += TPB) {
        const int root = base1[i];
        const int lo = pos1[root];
        const int nr = ns1[root];
        const int hi = (nr != NONE_I) ? pos1[nr] : N;
        int slot = QCAP;
        if (hi - lo > 24) slot = atomicAdd(&S.ibuf[4], 1);
        if (slot < QCAP) { queue[3 * slot] = (unsigned short)lo; queue[3 * slot + 1] = (unsigned short)hi; queue[3 * slot + 2] = (unsigned short)root; }
        else {
            double s = -1.0;
            for (int k = lo; k < hi; ++k) s = s + val[k];
            pitVol[root] = s;
        }
    }
    __syncthreads();
    const int nq = min(S.ibuf[4], QCAP);
    for (int i = warp; i < nq; i += NWARP) {
        const int lo = queue[3 * i], hi = queue[3 * i + 1], root = queue[3 * i + 2];
        double s = -1.0;
        for (int k0 = lo; k0 < hi; k0 += 32) {
            const int k = k0 + lane;
            const double v = (k < hi) ? val[k] : 0.0;
            unsigned m = __ballot_sync(0xffffffffu, v != 0.0);
            if (__popc(m) > 6) {            // wet stretch: all 32 terms in order (the zeros among them change nothing)
#pragma unroll
                for (int b = 0; b < 32; ++b) s = s + __shfl_sync(0xffffffffu, v, b);
            } else {
                while (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    s = s + __shfl_sync(0xffffffffu, v, b);
                }
            }
        }
        if (lane == 0) pitVol[root] = s;
    }
    __syncthreads();
}

// ---- pit drainage (FLOWalgo.f90:167-297): fast::drainage with its two shared-memory tables staged by bulk copies ----------
__device__ __noinline__ void drainage_c(const DevConst &c, Smem &S, unsigned char *dsm, const int *base1, int nbase1, double sea,
                                        int any_marine)
{
    BL_ASSUME_SHARED(c, S, dsm);
    const int N = c.N, Nq = c.Nq, tid = threadIdx.x;
    unsigned short *r16 = (unsigned short *)dsm, *pid16 = r16 + Nq;
    const unsigned short *r16g = (const unsigned short *)S.P.i[I_ROOT1];                 // filled-surface receivers (receivers_c)
    const unsigned short *pid16g = (const unsigned short *)S.P.i[I_FC] + np_of(N);       // pit ids (basins_c)
    const unsigned bytes = ((unsigned)N * 2u) & ~15u;
    if (tid < 8) { const int v = (int)(bytes / 2u) + tid; if (v < N) { r16[v] = r16g[v]; pid16[v] = pid16g[v]; } }
    tma_to_smem(S, r16, r16g, bytes);
    tma_to_smem(S, pid16, pid16g, bytes);
    drainage_body(c, S, dsm, base1, nbase1, sea, any_marine, true);
}

// ---- discharge (FLOWalgo.f90:53-81) on the breadth-first layout ---------------------------------------------------------
// One 16-byte record per breadth-first position q (global scratch, read as one 128-bit load, one level ahead):
//   x = first-child position : 16 | donor count : 8 | deposition rule : 8 (the rule is written by the stream-power pre-pass)
//   y = node at position q
//   z, w = fp64 payload: the node's own discharge (area * rain) for discharge_c; the pre-pass overwrites it with the
//          control volume where the routing needs it (rule 3)
// posb[v] = position of node v (u16, I_PS).  S.boff = bucket offsets.
struct Bfs {
    uint4 *rec;
    unsigned short *posb;
};
__device__ __forceinline__ Bfs bfs_arrays(Smem &S, int N)
{
    Bfs b;
    b.rec = reinterpret_cast<uint4 *>(S.P.euA);          // [Np] 16 bytes each; the tour successors lived here during links_stack_c
    b.posb = reinterpret_cast<unsigned short *>(S.P.i[I_PS]) + np_of(N);
    return b;
}

// S.ibuf[5] = 1 when the receiver tree is too deep for the shared-memory tables: generic phases take over for this step
__device__ __noinline__ void discharge_c(const DevConst &c, Smem &S, unsigned char *dsm, double rain, double erod)
{
    BL_ASSUME_SHARED(c, S, dsm);
    const int N = c.N, Nq = c.Nq, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int *__restrict__ rcv = S.P.i[I_RCV], *__restrict__ stack = S.P.i[I_STACK];
    const unsigned char *__restrict__ cnt8 = (const unsigned char *)S.P.i[I_CNT];
    const Bfs B = bfs_arrays(S, N);
    unsigned *jd = (unsigned *)dsm;                                              // [Nq] (jump:16 | dist:16)
    unsigned short *dpos = (unsigned short *)(dsm + (size_t)4 * Nq), *qpos = dpos + Nq;   // depth / BFS position by stack position
    unsigned short *cntW = (unsigned short *)dsm;                                // [maxd + 1][NWARP] after jd
    double *vals = (double *)dsm;                                                // [Nq] by BFS position, after all of them
    const unsigned lt = (1u << lane) - 1u;
    SubTick stc = sub_tick_begin(c);
    int md = 0;
    if (c.dbg & 16) {            // reference form: depths by pointer jumping over the receivers
        for (int v = tid; v < N; v += TPB) { const unsigned r = (unsigned)rcv[v]; jd[v] = (r << 16) | (r == (unsigned)v ? 0u : 1u); }
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
        for (int p = tid; p < N; p += TPB) { const int d = (int)(jd[stack[p]] & 0xFFFFu); dpos[p] = (unsigned short)d; md = max(md, d); }
    } else {                     // links_stack_c left the depth of every stack position behind, and their maximum in S.ibuf[2]
        const unsigned short *depg = (const unsigned short *)S.P.i[I_FC];
        const unsigned bytes = ((unsigned)N * 2u) & ~15u;
        if (tid < 8) { const int p = (int)(bytes / 2u) + tid; if (p < N) dpos[p] = depg[p]; }
        tma_to_smem(S, dpos, depg, bytes);
    }
    md = __reduce_max_sync(0xffffffffu, md);
    if (lane == 0) atomicMax(&S.ibuf[2], md);
    __syncthreads();
    const int maxd = S.ibuf[2];
    const int dlim = (c.dbg & 8) ? 12 : min(DEPMAX - 2, (4 * Nq) / (2 * NWARP) - 1);
    const bool deep = maxd > dlim;
    __syncthreads();
    if (tid == 0) { S.ibuf[2] = 0; S.ibuf[3] = maxd; S.ibuf[5] = deep ? 1 : 0; }
    if (deep) {
        __syncthreads();
        tree_links(c, S.P.i[I_RCV], S.P.i[I_FC], S.P.i[I_SUCC], S.P.i[I_LC], S.P.i[I_PS], S.P.i[I_CNT]);
        phase_discharge(c, S, rain);
        return;
    }
    sub_tick(stc, 22);      // depths
    // stable counting sort of the stack positions by depth: warp w owns the positions [w * L, (w + 1) * L)
    const int nflat = (maxd + 1) * NWARP;
    for (int i = tid; i < nflat; i += TPB) cntW[i] = 0;
    __syncthreads();
    const int L = (((N + NWARP - 1) / NWARP) + 31) & ~31;
    const int p0 = warp * L, p1 = min(N, p0 + L);
    for (int pb = p0; pb < p1; pb += 32) {
        const int p = pb + lane;
        const unsigned d = (p < p1) ? dpos[p] : 0xFFFFu;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        if (p < p1 && (peers & lt) == 0u) cntW[d * NWARP + warp] += (unsigned short)__popc(peers);
        __syncwarp();
    }
    __syncthreads();
    {
        const int R = (nflat + TPB - 1) / TPB;
        const int i0 = min(nflat, tid * R), i1 = min(nflat, i0 + R);
        int sum = 0;
        for (int i = i0; i < i1; ++i) sum += cntW[i];
        int tot;
        int ex = blk_exscan(sum, tot, S);
        for (int i = i0; i < i1; ++i) {
            const int cv = cntW[i];
            cntW[i] = (unsigned short)ex;
            if (i % NWARP == 0) S.boff[i / NWARP] = (unsigned short)ex;
            ex += cv;
        }
        if (tid == 0) S.boff[maxd + 1] = (unsigned short)N;
    }
    __syncthreads();
    for (int pb = p0; pb < p1; pb += 32) {
        const int p = pb + lane;
        const bool valid = p < p1;
        const unsigned d = valid ? dpos[p] : 0xFFFFu;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const int base = valid ? cntW[d * NWARP + warp] : 0;
        __syncwarp();
        if (valid && (peers & lt) == 0u) cntW[d * NWARP + warp] = (unsigned short)(base + __popc(peers));
        __syncwarp();
        if (valid) qpos[p] = (unsigned short)(base + __popc(peers & lt));
    }
    __syncthreads();
    sub_tick(stc, 23);      // stable sort by depth
    // in the stack a node's first (lowest-id) donor follows it directly; the donors of one node are contiguous in q
    // (by node id: the stack position, the donor count and the control volume of a node are coalesced, independent loads;
    // the positions come from shared memory; the initial value travels through global scratch because the position
    // tables still occupy the bytes of vals[])
    {
        const int *__restrict__ posA = S.P.i[I_POS];
        const double *__restrict__ areaA = c.area;
        uint4 *__restrict__ recw = B.rec;
        unsigned short *__restrict__ posw = B.posb;
        const int nbn = c.bounds;
#pragma unroll 4
        for (int v = tid; v < N; v += TPB) {
            const int p = posA[v];
            const unsigned n = cnt8[v];
            const double a = areaA[v];
            if (v + 4 * TPB < N) { pf_l2(posA + v + 4 * TPB); pf_l2(cnt8 + v + 4 * TPB); }
            const unsigned q = qpos[p];
            const unsigned f = (p + 1 < N) ? qpos[p + 1] : 0u;
            const double q0 = a * ((v >= nbn) ? rain : 0.);
            recw[q] = make_uint4(f | (n << 16), (unsigned)v, (unsigned)__double2loint(q0), (unsigned)__double2hiint(q0));
            posw[v] = (unsigned short)q;
        }
    }
    __syncthreads();
    sub_tick(stc, 24);      // records + initial values
    {
        const uint4 *__restrict__ recA = B.rec;
        const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
        int b0 = S.boff[maxd], b1 = S.boff[maxd + 1];
        uint4 rcur = (b0 + tid < b1) ? recA[b0 + tid] : zero4;
        for (int d = maxd; d >= 0; --d) {
            const int nb0 = (d > 0) ? S.boff[d - 1] : 0, nb1 = (d > 0) ? S.boff[d] : 0;
            const int niter = (b1 - b0 + TPB - 1) / TPB;          // uniform
            for (int it = 0; it < niter; ++it) {
                const int q = b0 + tid + it * TPB;
                // the record of this thread's next position (same bucket or the next one) is requested first
                const int nq = (it + 1 < niter) ? q + TPB : nb0 + tid;
                const int nlim = (it + 1 < niter) ? b1 : nb1;
                const uint4 rn = (nq < nlim) ? recA[nq] : zero4;
                if (q < b1) {
                    const int n = (int)((rcur.x >> 16) & 0xFFu), f = (int)(rcur.x & 0xFFFFu);
                    double acc = __hiloint2double((int)rcur.w, (int)rcur.z);           // area * rain of the node itself
                    for (int j = f + n - 1; j >= f; --j) acc = acc + vals[j];       // donors in descending id
                    vals[q] = acc;
                }
                rcur = rn;
            }
            __syncthreads();
            b0 = nb0; b1 = nb1;
        }
    }
    sub_tick(stc, 25);      // level loop
    // Q by node id (coalesced loads and stores, the gather is in shared memory) and, in the same pass, the flow CFL
    // condition (FLOWalgo.f90:299-330, same arithmetic as fast::flow_cfl; a base level has dz = 0 and is skipped like
    // d == r; STD exponents m = 0.5, n = 1 take sqrt and the identity): its minimum is left in S.dbuf[3]
    double *__restrict__ Q = S.P.f[F_Q];
    const unsigned short *__restrict__ posr = B.posb;
    const double *__restrict__ dzw = S.P.f[F_CAPH], *__restrict__ distr = S.P.f[F_E];
    const bool stdexp = (c.m == 0.5 && c.n == 1.0);
    const double em = c.m, en1 = c.n - 1.;
    double cfl = 1.e6;
#pragma unroll 4
    for (int v = tid; v < N; v += TPB) {
        const double q = vals[posr[v]], dz = dzw[v], dist = distr[v];
        if (v + 4 * TPB < N) { pf_l2(posr + v + 4 * TPB); pf_l2(dzw + v + 4 * TPB); pf_l2(distr + v + 4 * TPB); }
        Q[v] = q;
        if (q > 0. && dz > 0.) {
            const double tmp = stdexp ? dist / (erod * sqrt(q) * 1.0) : dist / (erod * powx(q, em) * powx(dz / dist, en1));
            cfl = fmin(tmp, cfl);
        }
    }
    cfl = blk_min(cfl, S);
    if (tid == 0) S.dbuf[3] = cfl;
    __syncthreads();
    sub_tick(stc, 26);      // Q by node id + CFL
}

// ---- stream power (FLOWalgo.f90:583-1044), see fast::streampower ---------------------------------------------------------
template <bool STD>
__device__ __noinline__ void sp_prepass_c(const DevConst &c, Smem &S, unsigned char *dsm, double erod, double sea, double dt)
{
    BL_ASSUME_SHARED(c, S, dsm);
    const int N = c.N, Nq = c.Nq, tid = threadIdx.x;
    const Bfs B = bfs_arrays(S, N);
    double *vals = (double *)dsm;
    unsigned *evbits = (unsigned *)(dsm + (size_t)8 * Nq);
    unsigned char *kindg;
    // The per-proposal arrays are addressed as base + field * Np + node: ONE base pointer per element type lives in
    // registers instead of a dozen array pointers (under the 64-register cap those were spilled and re-read from local
    // memory in every iteration, a second memory round trip in front of each load).
    double *__restrict__ const F = S.P.f[0];
    int *__restrict__ const I = S.P.i[0];
    const unsigned Np = (unsigned)np_of(N);
#define FD(k, i) F[(unsigned)(k) * Np + (unsigned)(i)]
#define ID(k, i) I[(unsigned)(k) * Np + (unsigned)(i)]
    const unsigned short *posbA = reinterpret_cast<const unsigned short *>(I + (size_t)I_PS * Np) + Np;
    kindg = reinterpret_cast<unsigned char *>(S.P.euA);            // byte 3 of the 16-byte record of a position
    double *areaq = reinterpret_cast<double *>(S.P.euA);           // its second half
    const double *__restrict__ areaA = c.area;
    // every load of a node is issued before the first branch on it (a load behind a data-dependent branch costs its own
    // memory round trip: the profile of the first version showed ten separate long-scoreboard waits per node)
#pragma unroll 1
    for (int d_ = tid; d_ < N; d_ += TPB) {
        // (the loop index is laundered through an asm move: otherwise the compiler keeps one running address per array,
        // a dozen 64-bit induction variables that it then spills and re-reads from local memory in every iteration)
        int d;
        asm volatile("mov.s32 %0, %1;" : "=r"(d) : "r"(d_));
        const int r = ID(I_RCV, d);
        const unsigned qd = posbA[d];
        const double zd = FD(F_Z, d), zr = FD(F_QS, d), fh = FD(F_FILLH, d), area = areaA[d];
        const double q = FD(F_Q, d), dist = FD(F_E, d), mh0 = FD(F_MAXH, d);
        const int pp = ID(I_PITID, d);
        {   // the per-proposal arrays come from DRAM (296 landscapes do not fit the L2): the next iteration's lines are
            // requested into L2 now
            const int dn = d + TPB;
            if (dn < N) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(&FD(F_Z, dn)));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(&FD(F_QS, dn)));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(&FD(F_FILLH, dn)));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(&FD(F_Q, dn)));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(&FD(F_E, dn)));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(&FD(F_MAXH, dn)));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(&ID(I_RCV, dn)));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(&ID(I_PITID, dn)));
            }
        }
        double dh = 0.95 * (zd - zr);
        if (zd > sea && zr < sea) dh = 0.99 * (zd - sea);
        if (dh < 0.001) dh = 0.;
        const double waterH = fh - zd;
        double SPL = 0., totspl = 0.;
        if (r != d && dh > 0. && waterH == 0. && fh >= sea) {
            if (dist > 0.) {
                const double slp = dh / dist;
                SPL = STD ? -erod * 1. * 1. * sqrt(q) * slp : -erod * 1. * 1. * powx(q, c.m) * powx(slp, c.n);
                totspl = totspl + SPL;
                if (-totspl * dt > dh) {
                    const double frac = dh / (-totspl * dt);
                    SPL = SPL * frac;
                    totspl = -dh;
                }
            }
        }
        double maxh = mh0;
        if (waterH > 0.) maxh = waterH;
        else if (zd < sea) maxh = sea - zd;
        maxh = 0.95 * maxh;
        int k;
        double e = 0.;
        if (totspl < 0.) { k = 0; e = SPL * dt * area; }
        else if (totspl >= 0. && area > 0.) {
            if (waterH > 0. && fh > sea) {
                k = 1;                            // pit under the sea: static first cascade step (:979-991)
                if (pp >= 0 && areaA[pp] > 0. && FD(F_FILLH, pp) < sea) k = (zd < sea) ? 2 : 4;
            }
            else if (zd <= sea) k = 2;
            else if (maxh > 0. && waterH == 0. && d != r && zd > sea) { k = 3; e = maxh; }
            else if (d == r) k = 2;
            else k = 4;
        } else k = 5;
        kindg[16 * qd + 3] = (unsigned char)k;
        vals[qd] = e;
        if (k == 3) areaq[2 * qd + 1] = area;         // the routing divides by it (rule 3 only)
        FD(F_DEPO, d) = 0.;
        FD(F_ERO, d) = 0.;
    }
#undef FD
#undef ID
    for (int v = tid; v < Nq / 32; v += TPB) evbits[v] = 0u;
    __syncthreads();
}

// ordered sediment routing (FLOWalgo.f90:875-964) over the breadth-first buckets
__device__ __noinline__ void sp_route_c(const DevConst &c, Smem &S, unsigned char *dsm, double dt)
{
    BL_ASSUME_SHARED(c, S, dsm);
    const int N = c.N, Nq = c.Nq, tid = threadIdx.x;
    const Bfs B = bfs_arrays(S, N);
    double *vals = (double *)dsm;
    unsigned *evbits = (unsigned *)(dsm + (size_t)8 * Nq);
    double *depo = S.P.f[F_DEPO], *ero = S.P.f[F_ERO], *pdep = S.P.f[F_PDEP];
    const uint4 *__restrict__ recA = B.rec;
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    const int *posg = S.P.i[I_POS];
    const int maxd = S.ibuf[3];
    // A thread's items are the positions q = b0 + tid, + TPB, ... of every bucket, deepest bucket first.  The record, the
    // node id and the control volume of the NEXT item (same bucket or the next one) are requested before the current
    // item is evaluated, so no global load sits on the dependent chain of a level.
    int b0 = S.boff[maxd], b1 = S.boff[maxd + 1];
    uint4 rcur = (b0 + tid < b1) ? recA[b0 + tid] : zero4;
    for (int d = maxd; d >= 0; --d) {
        const int nb0 = (d > 0) ? S.boff[d - 1] : 0, nb1 = (d > 0) ? S.boff[d] : 0;
        const int niter = (b1 - b0 + TPB - 1) / TPB;          // uniform
        for (int it = 0; it < niter; ++it) {
            const int q = b0 + tid + it * TPB;
            const int nq = (it + 1 < niter) ? q + TPB : nb0 + tid;
            const int nlim = (it + 1 < niter) ? b1 : nb1;
            const uint4 rn = (nq < nlim) ? recA[nq] : zero4;
            if (q < b1) {
                const unsigned r = rcur.x;
                const int cur = (int)rcur.y;
                const int n = (int)((r >> 16) & 0xFFu), f = (int)(r & 0xFFFFu), k = (int)(r >> 24);
                double acc = 0. * dt;
                for (int j = f + n - 1; j >= f; --j) acc = acc + vals[j];
                const double own = vals[q];
                double Qs = 0., erodep = 0.;
                if (k == 0) { erodep = own; Qs = -erodep + acc; }
                else if (k == 1) {
                    if (acc != 0.) {
                        pdep[cur] = acc;
                        if (acc > 0.) { const int ri = N - 1 - posg[cur]; atomicOr(&evbits[ri >> 5], 1u << (ri & 31)); }
                    }
                }
                else if (k == 2) { erodep = acc; }
                else if (k == 3) {
                    const double area = __hiloint2double((int)rcur.w, (int)rcur.z);
                    const double totflx = 0. + acc, maxh = own;
                    if (totflx / area < maxh) erodep = acc;
                    else { const double frac = acc / totflx; erodep = frac * maxh * area; Qs = acc - erodep; }
                } else if (k == 4) Qs = acc;
                if (!(k == 1 && acc != 0.)) {
                    if (erodep < 0.) ero[cur] = 0. + erodep;
                    else if (erodep != 0.) depo[cur] = 0. + erodep;
                }
                vals[q] = Qs;
            }
            rcur = rn;
        }
        __syncthreads();
        b0 = nb0; b1 = nb1;
    }
    __syncthreads();
}

__device__ __noinline__ void streampower_c(const DevConst &c, Smem &S, unsigned char *dsm, double erod, double sea, double dt,
                                           const int *base1, int nbase1)
{
    BL_ASSUME_SHARED(c, S, dsm);
    const int N = c.N, Nq = c.Nq, tid = threadIdx.x;
    const int *pitID = S.P.i[I_PITID], *stack = S.P.i[I_STACK], *pitDrain = S.P.i[I_PITDRAIN], *rcv = S.P.i[I_RCV];
    const double *z = S.P.f[F_Z], *W = S.P.f[F_FILLH];
    const double *pitVol1 = S.P.f[F_PITVOL];
    double *depo = S.P.f[F_DEPO], *pdep = S.P.f[F_PDEP], *pitVol = S.P.f[F_PITVOLW];
    long long *pcs = c.phase_cyc + (size_t)blockIdx.x * NPHASE;
    long long ts0 = clock64();
#define STICK(i) do { if (tid == 0) { long long t1_ = clock64(); tick_add(pcs + (i), t1_ - ts0); ts0 = t1_; } } while (0)
    if (c.m == 0.5 && c.n == 1.0) sp_prepass_c<true>(c, S, dsm, erod, sea, dt);
    else sp_prepass_c<false>(c, S, dsm, erod, sea, dt);
    STICK(27);
    sp_route_c(c, S, dsm, dt);
    STICK(15);
    // lake-entry events in reverse stack order, staged in the bytes the routing no longer needs:
    // loads fp64 [cap] | pits u16 [cap] | donors u16 [cap] | "pit receives an event" bytes [Nq]
    unsigned *evbits = (unsigned *)(dsm + (size_t)8 * Nq);
    const int cap = (int)(((size_t)7 * Nq) / 12) & ~7;
    double *ev_dep = (double *)dsm;
    unsigned short *ev_pit = (unsigned short *)(dsm + (size_t)8 * cap), *ev_donor = ev_pit + cap;
    unsigned char *has_ev = dsm + (size_t)7 * Nq;
    int nev;
    {
        const int nw = (N + 31) >> 5;
        int tot_mine = 0;
        unsigned words[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int wi = tid * 4 + u;
            words[u] = (wi < nw) ? evbits[wi] : 0u;
            tot_mine += __popc(words[u]);
        }
        const int off0 = blk_exscan(tot_mine, nev, S);      // barriers inside: every thread has its words, vals is free
        if (nev <= cap) {
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
    }
    if (nev > cap) {          // more events than the staging area holds: the literal serial pass (never seen in practice)
        __syncthreads();
        phase_streampower_serial(c, S, erod, sea, dt);
        return;
    }
    for (int v = tid; v < Nq / 4; v += TPB) reinterpret_cast<unsigned *>(has_ev)[v] = 0u;
    __syncthreads();
    for (int e = tid; e < nev; e += TPB) {
        const int donor = stack[N - 1 - ev_donor[e]];
        const int pit = pitID[donor];
        const bool ok = (pit >= 0 && c.area[pit] > 0.);
        if (ok) has_ev[pit] = 1;
        ev_pit[e] = ok ? (unsigned short)pit : NONE16;
        ev_donor[e] = (unsigned short)donor;
        ev_dep[e] = pdep[donor];
    }
    __syncthreads();
    STICK(12);
    if (nev == 0) return;
    int over = 0, allaff = 0;      // allaff: an event's pit lies under the sea - the serial replay takes every event
    if (S.ibuf[1])
        for (int e = tid; e < nev; e += TPB) { const int pit = ev_pit[e]; if (pit != NONE16 && W[pit] < sea) allaff = 1; }
    // one WARP per pit that receives events: the lanes scan the event list 32 entries at a time and the matching events
    // are applied in order (same sums as the serial loop: events of other pits do not touch this pit's deposit)
    int *plist = S.P.i[I_EVENTS];
    const int npl = blk_compact(nbase1, plist, [&](int i) { return has_ev[base1[i]] != 0; }, S);
    {
        const int lane = tid & 31, warp = tid >> 5;
        const double *pitVolA = pitVol1;
        for (int ip = warp; ip < npl; ip += NWARP) {
            const int pit = base1[plist[ip]];
            const double vol = pitVolA[pit];
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
            if (ovf) { over = 1; if (lane == 0) has_ev[pit] = 3; }     // bit 1: this pit's events must be replayed in order
            if (lane == 0) S.P.f[F_VAL][pit] = dp;
        }
    }
    {
        const int both = __syncthreads_or(over | (allaff << 1));
        allaff = (both >> 1) & 1;
        over = (both & 1) | allaff;
    }
    STICK(13);
    if (!over) {
        for (int i = tid; i < nbase1; i += TPB) { const int pit = base1[i]; if (has_ev[pit]) depo[pit] = S.P.f[F_VAL][pit]; }
        __syncthreads();
        return;
    }
    // Some pit overflows.  A cascade only moves sediment DOWN the drainage chain (FLOWalgo.f90:1002 tmpID = pitDrain), so
    // the pits whose deposit can differ from the independent sums above are the overflowing pits and everything
    // downstream of them: those are marked, every other pit commits its sum, and the literal serial replay
    // (FLOWalgo.f90:967-1034) runs over the events of the marked pits only, in the original order.
    if (!allaff) {
        for (int ip = tid; ip < npl; ip += TPB) {
            int p = base1[plist[ip]];
            if (!(has_ev[p] & 2)) continue;
            for (int it = 0; it < nbase1; ++it) {
                const int q = pitDrain[p];
                if (q < 0 || q == p) break;
                has_ev[q] |= 2;
                p = q;
            }
        }
        __syncthreads();
        for (int i = tid; i < nbase1; i += TPB) { const int pit = base1[i]; if (has_ev[pit] == 1) depo[pit] = S.P.f[F_VAL][pit]; }
    }
    for (int k = tid; k < N; k += TPB) pitVol[k] = pitVol1[k];
    if (tid == 0) S.ibuf[0] = 0;
    __syncthreads();
    if (tid == 0) {
        unsigned st = 0;
        for (int e = 0; e < nev; ++e) {
            if (ev_pit[e] == NONE16) continue;
            if (!allaff && !(has_ev[ev_pit[e]] & 2)) continue;
            const int donor = ev_donor[e];
            const int recvr = rcv[donor];
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
                    else { S.ibuf[0] = 1; e = nev; }
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
    STICK(14);
    if (S.ibuf[0]) phase_streampower_serial(c, S, erod, sea, dt);
#undef STICK
}

// buildFlux.py:193-195, :252-272, :292-315 (see fast::apply): z fp64 in shared memory, the increments in global scratch
template <int KDT>
__device__ __noinline__ void apply_c(const DevConst &c, Smem &S, unsigned char *dsm, double sea, double timestep)
{
    BL_ASSUME_SHARED(c, S, dsm);
    const int N = c.N, tid = threadIdx.x;
    double *__restrict__ zs = (double *)dsm;
    double *__restrict__ cd = S.P.f[F_VAL];
    double *__restrict__ z = S.P.f[F_Z];
    double *__restrict__ cumdiff = S.P.f[F_CUMDIFF];
    double *__restrict__ cumhill = S.P.f[F_CUMHILL];
    const double *__restrict__ sedflux = S.P.f[F_SEDFLUX];
    const unsigned char *__restrict__ gflags = c.flags;
    const unsigned *__restrict__ nbin = c.nbf2;
#pragma unroll 4
    for (int k = tid; k < N; k += TPB) {
        if (k + 4 * TPB < N) { pf_l2(z + k + 4 * TPB); pf_l2(sedflux + k + 4 * TPB); }
        zs[k] = z[k] + sedflux[k];
    }
    __syncthreads();
    const double *__restrict__ dispA = c.disp;
    const int nbn = c.bounds;
    const double *__restrict__ areaA = c.area;
    const float *__restrict__ ve32 = c.ve32;
#pragma unroll 1
    for (int g_ = tid; g_ < N; g_ += TPB) {
        // every load of the node up front (rows included: a row fetched behind the flag test is a second round trip);
        // the index is laundered so that the compiler keeps no per-array running address (see sp_prepass_c)
        int g;
        asm volatile("mov.s32 %0, %1;" : "=r"(g) : "r"(g_));
        double acc = 0.;
        const unsigned char fg = gflags[g];
        const unsigned inm = nbin[g];            // bit p: neighbour slot p is an inside node (borders2)
        const double zg = zs[g];
        // the three per-proposal values are requested here and first used after the gather below (they come from DRAM)
        const double sf = sedflux[g], cdf0 = cumdiff[g], h0 = cumhill[g];
        if (g + TPB < N) { pf_l2(sedflux + g + TPB); pf_l2(cumdiff + g + TPB); pf_l2(cumhill + g + TPB); }
        const double dsp = dispA ? dispA[g] * timestep : 0.;
        const double area = areaA[g];
        Ids<KDT> ids;
        load_ids<KDT>(c, g, ids);
        if (ve32) {
            float4 v4[KDT / 2];
            const float4 *ve = reinterpret_cast<const float4 *>(ve32 + (size_t)g * KDT * 2);
#pragma unroll
            for (int q = 0; q < KDT / 2; ++q) v4[q] = __ldg(ve + q);
            if (fg & 2) {
#pragma unroll
                for (int q = 0; q < KDT / 2; ++q) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int p = 2 * q + h;
                        if (ids.get(p) == NONE16) continue;
                        const int j = ids.get(p);
                        const double ed = (double)(h ? v4[q].z : v4[q].x), di = (double)(h ? v4[q].w : v4[q].y), zj = zs[j];
                        const bool bj = (inm >> p) & 1u;
                        if (bj && di > 0.) acc += ed * (zj - zg) / di;
                        if (!bj) { if (zj < zg && di > 0.) acc += ed * (zj - zg) / di; }
                    }
                }
            }
        } else if (fg & 2) {
#pragma unroll
            for (int p = 0; p < KDT; ++p) {
                if (ids.get(p) == NONE16) continue;
                const int j = ids.get(p);
                const double di = c.edge[(size_t)p * N + g], ed = c.vor[(size_t)p * N + g], zj = zs[j];
                const bool bj = (inm >> p) & 1u;
                if (bj && di > 0.) acc += ed * (zj - zg) / di;
                if (!bj) { if (zj < zg && di > 0.) acc += ed * (zj - zg) / di; }
            }
        }
        const double ac = (area > 0.) ? 1. / area : 0.;
        double coeff = ac * ((zg >= sea) ? c.CDa : c.CDm);
        if (!(fg & 2)) { coeff = 0.; acc = 0.; }
        if (S.diff_out) S.diff_out[g] = acc;
        const double d = coeff * acc * timestep;
        // the node's new elevation is final here (its neighbours read the old one from shared memory); ghost nodes
        // take their parent's new elevation below, so the increments of the parents are kept
        cd[g] = d;
        const double c0 = cdf0 + sf;
        if (fg & 1) { cumdiff[g] = c0 + d; cumhill[g] = h0 + d; }
        else cumdiff[g] = c0;
        if (g >= nbn || c.btype == 0) {
            double zn = (fg & 1) ? zg + d : zg;
            if (dispA) zn += dsp;
            z[g] = zn;
        }
    }
    __syncthreads();
    if (c.btype != 0) {
        const double add = (c.btype == 1) ? -0.1 : (c.btype == 2 ? 0. : 100.);
        for (int k = tid; k < nbn; k += TPB) {
            const int pk = c.parent[k];
            double zp = zs[pk];
            if (gflags[pk] & 1) zp += cd[pk];
            double zn = (c.btype == 2) ? zp : zp + add;
            if (dispA) zn += dispA[k] * timestep;
            z[k] = zn;
        }
        __syncthreads();
    }
}

template <int KDT>
__device__ void time_step(const DevConst &c, Smem &S, unsigned char *dsm)
{
    const int tid = threadIdx.x;
    if (tid == 0) { S.sea = sea_level(c, S.tNow); S.tick0 = clock64(); }
    __syncthreads();
    {
        int any_marine;
        fill_c<KDT>(c, S, dsm, S.sea);
        receivers_c<KDT>(c, S, dsm, S.sea, any_marine);
        if (tid == 0) { S.ibuf[1] = any_marine; S.any_marine = any_marine; }
    }
    BL_TICK(0);
    {
        int nbase1;
        links_stack_c<KDT>(c, S, dsm, S.P.i[I_RCV1], S.P.i[I_STACK1], S.P.i[I_POS1], S.P.i[I_LA], nbase1, false, false, 0ull);
    }
    BL_TICK(2);
    basins_c(c, S, dsm, S.P.i[I_LA], S.nbase1);
    BL_TICK(3);
    {
        int nbase;
        links_stack_c<KDT>(c, S, dsm, S.P.i[I_RCV], S.P.i[I_STACK], S.P.i[I_POS], S.P.i[I_LB], nbase, true, c.shuffle != 0,
                           mix64(S.seed + (unsigned long long)S.step));
    }
    BL_TICK(4);
    drainage_c(c, S, dsm, S.P.i[I_LA], S.nbase1, S.sea, S.any_marine);
    BL_TICK(5);
    discharge_c(c, S, dsm, S.rain, S.erod);
    BL_TICK(6);
    const bool deep = S.ibuf[5] != 0;          // uniform: written before the barriers that end discharge_c
    {
        const double flowcfl = deep ? phase_cfl(c, S, S.erod) : S.dbuf[3];       // discharge_c computed it with its last pass
        double cfl = fmin(flowcfl, c.hill);
        cfl = py2_floorish(cfl);
        cfl = fmin(cfl, S.tStop - S.tNow);
        cfl = fmax(c.minDT, cfl);
        cfl = fmin(c.maxDT, cfl);
        if (tid == 0) { S.dbuf[2] = cfl; S.newdt = cfl; }
        __syncthreads();
    }
    BL_TICK(7);
    if (deep) phase_streampower(c, S, S.erod, S.sea, S.newdt);
    else streampower_c(c, S, dsm, S.erod, S.sea, S.newdt, S.P.i[I_LA], S.nbase1);
    BL_TICK(8);
    {
        // getid1 (FLOWalgo.f90:1047-1086), see fast::time_step
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
        if (newdt < cfl) {
            if (deep) phase_streampower(c, S, S.erod, S.sea, S.newdt);
            else streampower_c(c, S, dsm, S.erod, S.sea, S.newdt, S.P.i[I_LA], S.nbase1);
        }
    }
    BL_TICK(9);
    phase_erodep(c, S, S.sea);
    BL_TICK(10);
    apply_c<KDT>(c, S, dsm, S.sea, S.newdt);
    BL_TICK(11);
    if (tid == 0) { S.tNow += S.newdt; S.last_dt = S.newdt; S.step += 1; }
    __syncthreads();
}

} // namespace fastc

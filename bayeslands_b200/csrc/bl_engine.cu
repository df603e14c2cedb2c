// bl_engine.cu - B200 (sm_100a) engine of the Bayeslands hot path: one persistent CTA per proposal.
//
// Design: each (rain, erodibility) proposal is an independent landscape; one thread block of
// 1024 threads owns it for the whole Model.run_to_time loop, so the ragged, CFL-adaptive time stepping
//  needs no host round trip and no grid-wide barrier: every phase of a time step is a
// block-wide pass separated by __syncthreads().  Per-proposal arrays are contiguous in HBM/L2; mesh
// constants are slot-major ([slot][node]) so that a warp reading one neighbour slot of 32 consecutive
// nodes is coalesced.  All arithmetic is fp64 with FMA contraction disabled (-fmad=false) so that results
// are bit-identical to the CPU oracle compiled with -ffp-contract=off.
//
// Reference routines restated per phase are cited at each device function (paths under /root/reference).
#include "bl_common.cuh"

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct bl_ctx {
    DevConst c;
    int tpb;      // threads per CTA: 256 for small meshes (several CTAs per SM), 1024 otherwise
    int device;
    int B;
    size_t ws_bytes;
    char *ws;
    double *d_stops;
    int *d_ckpt;
    int stops_cap;
    long long launches;
    size_t const_bytes;   // mesh constants at the start of the workspace (kept L2-persistent)
    int l2_persist;       // bl_params.opt_flags & BL_OPT_NO_L2_PERSIST == 0
    int dyn_max;          // dynamic shared memory one CTA of the chosen kernel may use (opt-in limit - static)
    int dyn;              // dynamic shared memory of a launch
    int compact;          // compact shared-memory layout (two landscapes per SM)
    size_t need_full;     // dynamic shared memory of the 17-bytes-per-node layout (per-stage kernels)
    // stop times / checkpoint ids travel through context-owned pinned slots (no stream synchronisation per launch)
    static constexpr int NSLOT = 4;
    char *h_stage;        // [NSLOT][STOPS_CAP * 12] pinned
    cudaEvent_t slot_ev[NSLOT];
    int slot;
};

std::string &bl_err_ref()
{
    thread_local std::string e;
    return e;
}

extern "C" const char *bl_last_error(void) { return g_err.c_str(); }
extern "C" int bl_version(void) { return 100; }

namespace {
struct Layout {
    size_t off_pend0, off_nbf2, off_ve32, off_ngb16, off_ngb16f, off_ngb, off_vor, off_edge, off_area, off_xy, off_flags, off_parent, off_z0, off_disp, off_levn,
        off_levo, off_gidx, off_gw, off_gex, off_real, off_pts, off_seat, off_seav, off_rain, off_erod, off_seed,
        off_tnow, off_lastdt, off_steps, off_status, off_phase, off_stops, off_ckpt, off_prop, total;
    int KD, nlev;
};

int max_degree(const bl_mesh *m)
{
    int kd = 0;
    for (int i = 0; i < m->N; ++i) {
        int d = 0;
        while (d < BL_KP && m->ngb[(size_t)i * BL_KP + d] >= 0) ++d;
        kd = std::max(kd, d);
    }
    return kd;
}

constexpr int STOPS_CAP = 4096;

// dynamic shared memory the shared-memory device path (bl_fast.cuh) needs for a mesh of Nq (padded) nodes:
// 17 bytes per node for the scans (value fp64 + links + order + kind) + the event bitmap; the fill keeps W, z, the pending
// flags and the per-warp id rings; the marine phases of the etopo configs stage 33 bytes per node (marine_diffusion) or
// the neighbour table as well (marine_distribution) - those fall back to global memory when the CTA cannot have it.
size_t fast_smem_need(int Nq, int KDT, bool marine)
{
    size_t need = (size_t)17 * Nq + Nq / 8 + 256;
    need = std::max(need, (size_t)17 * Nq + 32 * 128 + 64);        // fill: W, z, pending flags, per-warp slot lists
    if (marine) need = std::max(need, std::max((size_t)Nq * (17 + 2 * KDT) + 256, (size_t)33 * Nq + 64));
    return need;
}

// the compact layout (bl_fastc.cuh): 8 bytes per node, one bit per node, the fill's slot lists (256 bytes per sweeping warp)
size_t compact_smem_need(int Nq, int tpb)
{
    const int nwf = (tpb <= 512) ? tpb / 32 : tpb / 64;      // sweeping warps of the fill (bl_fastc.cuh: fill_c), 256-byte slot lists
    return (size_t)8 * Nq + Nq / 8 + (size_t)nwf * 256 + 64;
}

cudaError_t func_attr(int tpb, int KDT, cudaFuncAttributes *fa)
{
    if (tpb == 256) return bl_func_attr_v256(KDT, fa);
    if (tpb == 512) return bl_func_attr_v512(KDT, fa);
    if (tpb == 768) return bl_func_attr_v768(KDT, fa);
    if (tpb == 384) return bl_func_attr_v384(KDT, fa);
    return bl_func_attr_v1024(KDT, fa);
}

Layout make_layout(const bl_mesh *m, const bl_params *p, int B, int KD, int nlev)
{
    Layout L;
    size_t o = 0;
    const size_t N = m->N;
    auto take = [&](size_t bytes) { size_t r = o; o = rup(o + bytes, 256); return r; };
    L.off_ngb = take((size_t)KD * N * 4);
    L.off_ngb16 = take((size_t)BL_KP * rup(N + 1, 64) * 2);
    L.off_ngb16f = take((size_t)BL_KP * rup(N + 1, 64) * 2);
    L.off_ve32 = take((size_t)BL_KP * rup(N + 1, 64) * 8);
    L.off_vor = take((size_t)KD * N * 8);
    L.off_edge = take((size_t)KD * N * 8);
    L.off_area = take(N * 8);
    L.off_xy = take(N * 16);
    L.off_flags = take(N);
    L.off_nbf2 = take(N * 4);
    L.off_pend0 = take(rup(N + 1, 64) / 8 + 64);
    L.off_parent = take((size_t)std::max(1, m->bounds) * 4);
    L.off_z0 = take(N * 8);
    L.off_disp = take(N * 8);
    L.off_levn = take(N * 4);
    L.off_levo = take((size_t)(nlev + 1) * 4);
    L.off_gidx = take((size_t)std::max(1, m->G) * 12);
    L.off_gw = take((size_t)std::max(1, m->G) * 24);
    L.off_gex = take((size_t)std::max(1, m->G) * 4);
    L.off_real = take((size_t)std::max(1, m->G) * 8);
    L.off_pts = take(BL_NPTS_MAX * 4);
    L.off_seat = take((size_t)std::max(1, p->nsea) * 8);
    L.off_seav = take((size_t)std::max(1, p->nsea) * 8);
    L.off_rain = take((size_t)B * 8);
    L.off_erod = take((size_t)B * 8);
    L.off_seed = take((size_t)B * 8);
    L.off_tnow = take((size_t)B * 8);
    L.off_lastdt = take((size_t)B * 8);
    L.off_steps = take((size_t)B * 8);
    L.off_status = take((size_t)B * 4);
    L.off_phase = take((size_t)B * NPHASE * 8);
    L.off_stops = take((size_t)STOPS_CAP * 8);
    L.off_ckpt = take((size_t)STOPS_CAP * 4);
    L.off_prop = take((size_t)B * prop_bytes(m->N));
    L.total = o;
    L.KD = KD;
    L.nlev = nlev;
    return L;
}

// Gauss-Seidel dependency levels of fillPD's ascending-id sweep (see phase_fill)
int gs_levels(const bl_mesh *m, std::vector<int> &level)
{
    const int N = m->N;
    level.assign(N, 0);
    int nlev = 0;
    for (int k = m->bounds; k < N; ++k) {
        int l = 0;
        for (int p = 0; p < BL_KP; ++p) {
            int j = m->ngb[(size_t)k * BL_KP + p];
            if (j < 0) break;
            if (j >= m->bounds && j < k) l = std::max(l, level[j]);
        }
        level[k] = l + 1;
        nlev = std::max(nlev, l + 1);
    }
    return nlev;
}
} // namespace

extern "C" size_t bl_workspace_bytes(const bl_mesh *mesh, const bl_params *par, int B)
{
    if (!mesh || !par || B <= 0 || mesh->N <= 0) return 0;
    std::vector<int> level;
    int nlev = gs_levels(mesh, level);
    return make_layout(mesh, par, B, std::max(1, max_degree(mesh)), std::max(1, nlev)).total;
}

extern "C" int bl_ctx_create(const bl_mesh *m, const bl_params *p, int B, int device, void *workspace,
                             size_t workspace_bytes, bl_ctx **out)
{
    if (!m || !p || !out || !workspace || B <= 0) { g_err = "bl_ctx_create: null argument"; return BL_ERR_ARG; }
    const int N = m->N;
    if (N <= 0 || m->bounds < 0 || m->bounds > N) { g_err = "bl_ctx_create: bad N/bounds"; return BL_ERR_ARG; }
    if (p->npts > BL_NPTS_MAX) { g_err = "bl_ctx_create: too many erdp points"; return BL_ERR_ARG; }
    // adjacency must be symmetric: donors are looked up among a node's neighbours
    for (int i = 0; i < N; ++i)
        for (int q = 0; q < BL_KP; ++q) {
            int j = m->ngb[(size_t)i * BL_KP + q];
            if (j < 0) break;
            if (j >= N || j == i) { g_err = "bl_ctx_create: neighbour id out of range"; return BL_ERR_ARG; }
            bool ok = false;
            for (int r = 0; r < BL_KP; ++r) {
                int k = m->ngb[(size_t)j * BL_KP + r];
                if (k < 0) break;
                if (k == i) { ok = true; break; }
            }
            if (!ok) { g_err = "bl_ctx_create: neighbour lists are not symmetric"; return BL_ERR_ARG; }
        }
    std::vector<int> level;
    const int nlev = std::max(1, gs_levels(m, level));
    const int KD = std::max(1, max_degree(m));
    Layout L = make_layout(m, p, B, KD, nlev);
    if (workspace_bytes < L.total) { g_err = "bl_ctx_create: workspace too small"; return BL_ERR_SPACE; }
    CUDA_OK(cudaSetDevice(device));
    char *ws = (char *)workspace;
    // host staging
    std::vector<int> ngbT((size_t)KD * N, -1);
    std::vector<double> vorT((size_t)KD * N, 0.), edgeT((size_t)KD * N, 0.);
    for (int i = 0; i < N; ++i)
        for (int q = 0; q < KD; ++q) {
            int j = m->ngb[(size_t)i * BL_KP + q];
            if (j < 0) break;
            ngbT[(size_t)q * N + i] = j;
            vorT[(size_t)q * N + i] = m->vor[(size_t)i * BL_KP + q];
            edgeT[(size_t)q * N + i] = m->edge[(size_t)i * BL_KP + q];
        }
    const int Nq = (int)rup((size_t)N + 1, 64);     // one spare slot: the fill's dummy neighbour (see DevConst::ngb16f)
    // CTA size and device path.  The shared-memory path needs fast_smem_need() bytes of dynamic shared memory; what a CTA
    // may use is the device's opt-in limit minus the kernel's static shared memory (queried, not assumed).
    const int kdt_want = (KD <= 8) ? 8 : (KD <= 12 ? 12 : 20);
    if (p->opt_tpb != 0 && p->opt_tpb != 256 && p->opt_tpb != 384 && p->opt_tpb != 512 && p->opt_tpb != 768 && p->opt_tpb != 1024) {
        g_err = "bl_ctx_create: opt_tpb must be 0, 256, 384, 512, 768 or 1024"; return BL_ERR_ARG;
    }
    int optin = 0, smem_sm = 0, smem_res = 0;
    CUDA_OK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    CUDA_OK(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device));
    CUDA_OK(cudaDeviceGetAttribute(&smem_res, cudaDevAttrReservedSharedMemoryPerBlock, device));
    int static_smem[1025] = {0};
    for (int t : {256, 384, 512, 768, 1024}) {
        cudaFuncAttributes fa;
        memset(&fa, 0, sizeof(fa));
        CUDA_OK(func_attr(t, kdt_want, &fa));
        static_smem[t] = (int)fa.sharedSizeBytes;
    }
    // does the mesh numbering put the Gauss-Seidel levels in contiguous id ranges (mesh.py does)?
    bool lev_contig_h = true;
    for (int k = m->bounds + 1; k < N; ++k) if (level[k] < level[k - 1]) { lev_contig_h = false; break; }
    const bool marine_cfg = (p->CDr > 0. || p->nsea > 0);
    const bool small_ids = (N < 32760 && nlev <= LEVMAX);
    const size_t need_full = fast_smem_need(Nq, kdt_want, false), need_c512 = compact_smem_need(Nq, 512);
    auto fits = [&](int t, size_t need, int nres) {        // nres CTAs of t threads resident on one SM with `need` bytes each
        return (size_t)nres * (static_smem[t] + rup(need, 128) + smem_res) <= (size_t)smem_sm &&
               static_smem[t] + rup(need, 128) <= (size_t)optin;
    };
    const bool compact_ok = small_ids && lev_contig_h && !marine_cfg && !(p->opt_flags & BL_OPT_NO_COMPACT);
    // CTA size and layout: small meshes 256 threads (several CTAs per SM); else two 512-thread CTAs per SM when two
    // landscapes fit - in the 17-bytes-per-node layout or in the compact one - else one 1024-thread CTA per SM
    // Sharing an SM only pays when there are more landscapes than SMs: a lone CTA finishes a time step sooner with 1024
    // threads and the 17-byte layout (single chains, speculative look-ahead batches of one wave).
    int nsm = 0;
    CUDA_OK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
    const bool share = (B > nsm) || (p->opt_flags & BL_OPT_COMPACT);
    int tpb, compact = 0;
    if (p->opt_tpb) tpb = p->opt_tpb;
    else if (N <= 4096) tpb = 256;
    else if (share && small_ids && fits(512, need_full, 2)) tpb = 512;
    else if (share && compact_ok && fits(512, need_c512, 2)) tpb = 512;
    else tpb = 1024;
    if (compact_ok) {
        if (p->opt_flags & BL_OPT_COMPACT) compact = fits(tpb, compact_smem_need(Nq, tpb), 1) ? 1 : 0;
        else if (tpb == 512 && !fits(512, need_full, 2) && fits(512, need_c512, 2)) compact = 1;
    }
    const size_t need_c = compact_smem_need(Nq, tpb);
    const int dyn_max = (optin - static_smem[tpb]) / 1024 * 1024;
    int KDT = 0;
    if (small_ids && (compact ? need_c : need_full) <= (size_t)dyn_max) KDT = kdt_want;
    if (p->opt_flags & BL_OPT_FORCE_GENERIC) { KDT = 0; compact = 0; }
    if (!KDT) compact = 0;
    const int KDrow = KDT ? KDT : BL_KP;
    std::vector<unsigned short> ngb16((size_t)BL_KP * Nq, 0xFFFFu);
    for (int i = 0; i < N; ++i)
        for (int q = 0; q < KD; ++q) {
            int j = m->ngb[(size_t)i * BL_KP + q];
            if (j < 0) break;
            ngb16[(size_t)i * KDrow + q] = (unsigned short)j;
        }
    std::vector<unsigned short> ngb16f((size_t)BL_KP * Nq, (unsigned short)N);
    if (KDT)
        for (int i = 0; i < N; ++i)
            for (int q = 0; q < KD; ++q) {
                int j = m->ngb[(size_t)i * BL_KP + q];
                if (j < 0) break;
                ngb16f[(size_t)i * KDrow + q] = (unsigned short)j;
            }
    std::vector<float> ve32((size_t)BL_KP * Nq * 2, 0.f);
    bool ve_ok = KDT > 0;
    for (int i = 0; i < N && ve_ok; ++i)
        for (int q = 0; q < KD; ++q) {
            if (m->ngb[(size_t)i * BL_KP + q] < 0) break;
            const double v = m->vor[(size_t)i * BL_KP + q], e = m->edge[(size_t)i * BL_KP + q];
            if ((double)(float)v != v || (double)(float)e != e) { ve_ok = false; break; }
            ve32[((size_t)i * KDrow + q) * 2] = (float)v;
            ve32[((size_t)i * KDrow + q) * 2 + 1] = (float)e;
        }
    std::vector<unsigned char> flags(N);
    for (int i = 0; i < N; ++i)
        flags[i] = (unsigned char)((m->borders[i] ? 1 : 0) | (m->borders2[i] ? 2 : 0) | (m->indomain[i] ? 4 : 0));
    std::vector<unsigned> nbf2(N, 0u);
    for (int i = 0; i < N; ++i)
        for (int q = 0; q < KD; ++q) {
            int j = m->ngb[(size_t)i * BL_KP + q];
            if (j < 0) break;
            if (m->borders2[j]) nbf2[i] |= 1u << q;
        }
    // the fill starts from W = 1e6 on every interior node: evaluating a node is a no-op until one of its neighbours has
    // dropped, so only the interior nodes next to a ghost node (W = z from the start) are pending initially
    std::vector<unsigned> pend0((size_t)rup((size_t)N + 1, 64) / 32 + 16, 0u);
    for (int i = m->bounds; i < N; ++i)
        for (int q = 0; q < KD; ++q) {
            int j = m->ngb[(size_t)i * BL_KP + q];
            if (j < 0) break;
            if (j < m->bounds) { pend0[i >> 5] |= 1u << (i & 31); break; }
        }
    std::vector<int> levn, levo(nlev + 1, 0);
    {
        std::vector<int> cntl(nlev + 1, 0);
        for (int k = m->bounds; k < N; ++k) cntl[level[k]]++;
        for (int l = 1; l <= nlev; ++l) levo[l] = levo[l - 1] + cntl[l];
        levn.assign(std::max(1, N - m->bounds), 0);
        std::vector<int> cur(levo.begin(), levo.end());
        for (int k = m->bounds; k < N; ++k) levn[cur[level[k] - 1]++] = k;
    }
#define UP(off, src, bytes) CUDA_OK(cudaMemcpy(ws + (off), (src), (bytes), cudaMemcpyHostToDevice))
    UP(L.off_ngb, ngbT.data(), ngbT.size() * 4);
    UP(L.off_ngb16, ngb16.data(), ngb16.size() * 2);
    UP(L.off_ngb16f, ngb16f.data(), ngb16f.size() * 2);
    UP(L.off_ve32, ve32.data(), ve32.size() * 4);
    UP(L.off_vor, vorT.data(), vorT.size() * 8);
    UP(L.off_edge, edgeT.data(), edgeT.size() * 8);
    UP(L.off_area, m->area, (size_t)N * 8);
    UP(L.off_xy, m->xy, (size_t)N * 16);
    UP(L.off_flags, flags.data(), (size_t)N);
    UP(L.off_nbf2, nbf2.data(), (size_t)N * 4);
    UP(L.off_pend0, pend0.data(), (size_t)rup((size_t)N + 1, 64) / 8);
    if (m->bounds > 0) UP(L.off_parent, m->parent, (size_t)m->bounds * 4);
    UP(L.off_z0, m->z0, (size_t)N * 8);
    if (m->disp) UP(L.off_disp, m->disp, (size_t)N * 8);
    UP(L.off_levn, levn.data(), levn.size() * 4);
    UP(L.off_levo, levo.data(), levo.size() * 4);
    if (m->G > 0) {
        UP(L.off_gidx, m->grid_idx, (size_t)m->G * 12);
        UP(L.off_gw, m->grid_w, (size_t)m->G * 24);
        UP(L.off_gex, m->grid_exact, (size_t)m->G * 4);
        if (p->real_elev) UP(L.off_real, p->real_elev, (size_t)m->G * 8);
    }
    if (p->npts > 0) UP(L.off_pts, p->pts_cell, (size_t)p->npts * 4);
    if (p->nsea > 0) { UP(L.off_seat, p->seat, (size_t)p->nsea * 8); UP(L.off_seav, p->seav, (size_t)p->nsea * 8); }
#undef UP
    bl_ctx *ctx = new bl_ctx();
    DevConst &c = ctx->c;
    memset(&c, 0, sizeof(c));
    c.N = N; c.bounds = m->bounds; c.KD = KD; c.nlev = nlev; c.G = m->G; c.npts = p->npts; c.nsea = p->nsea; c.B = B;
    c.ngb = (const int *)(ws + L.off_ngb);
    c.ngb16 = (const unsigned short *)(ws + L.off_ngb16);
    c.ngb16f = (const unsigned short *)(ws + L.off_ngb16f);
    c.Nq = Nq; c.KDT = KDT;
    c.ve32 = ve_ok ? (const float *)(ws + L.off_ve32) : nullptr;
    c.lev_contig = 1;
    c.dbg = p->opt_debug;
    for (size_t i = 0; i < levn.size() && (int)i < N - m->bounds; ++i)
        if (levn[i] != m->bounds + (int)i) { c.lev_contig = 0; break; }
    c.vor = (const double *)(ws + L.off_vor);
    c.edge = (const double *)(ws + L.off_edge);
    c.area = (const double *)(ws + L.off_area);
    c.xy = (const double *)(ws + L.off_xy);
    c.flags = (const unsigned char *)(ws + L.off_flags);
    c.nbf2 = (const unsigned *)(ws + L.off_nbf2);
    c.pend0 = (const unsigned *)(ws + L.off_pend0);
    c.compact = compact;
    c.parent = (const int *)(ws + L.off_parent);
    c.z0 = (const double *)(ws + L.off_z0);
    c.disp = m->disp ? (const double *)(ws + L.off_disp) : nullptr;
    c.lev_nodes = (const int *)(ws + L.off_levn);
    c.lev_off = (const int *)(ws + L.off_levo);
    c.grid_idx = (const int *)(ws + L.off_gidx);
    c.grid_w = (const double *)(ws + L.off_gw);
    c.grid_exact = (const int *)(ws + L.off_gex);
    c.real_elev = (m->G > 0 && p->real_elev) ? (const double *)(ws + L.off_real) : nullptr;
    c.pts_cell = (const int *)(ws + L.off_pts);
    c.seat = (const double *)(ws + L.off_seat);
    c.seav = (const double *)(ws + L.off_seav);
    c.m = p->spl_m; c.n = p->spl_n; c.eps = p->eps; c.TH = p->fill_th; c.CDa = p->CDa; c.CDm = p->CDm; c.CDr = p->CDr;
    c.minDT = p->minDT; c.maxDT = p->maxDT; c.hill = p->hill_cfl; c.ms = p->ms_cfl; c.diffprop = p->diffprop;
    c.seapos = p->seapos; c.diffnb = p->diffnb; c.depo = p->depo; c.btype = p->btype; c.shuffle = p->shuffle_mode;
    c.prop_base = ws + L.off_prop;
    c.prop_stride = prop_bytes(N);
    c.rain = (double *)(ws + L.off_rain);
    c.erod = (double *)(ws + L.off_erod);
    c.seed = (unsigned long long *)(ws + L.off_seed);
    c.tnow = (double *)(ws + L.off_tnow);
    c.lastdt = (double *)(ws + L.off_lastdt);
    c.steps = (long long *)(ws + L.off_steps);
    c.status = (unsigned *)(ws + L.off_status);
    c.phase_cyc = (long long *)(ws + L.off_phase);
    ctx->tpb = (KDT == 0 && p->opt_tpb == 0) ? 1024 : tpb;
    ctx->compact = compact;
    ctx->need_full = rup(need_full, 1024);
    ctx->l2_persist = (p->opt_flags & BL_OPT_NO_L2_PERSIST) ? 0 : 1;
    ctx->dyn_max = dyn_max;
    {   // dynamic shared memory of a launch: what the phases of this configuration need; the marine phases (etopo configs)
        // stage more when it fits and fall back to their global-memory form otherwise (bl_device.cuh: dsm_bytes tests)
        const bool marine = (p->CDr > 0. || p->nsea > 0);
        size_t need = compact ? compact_smem_need(Nq, tpb) : fast_smem_need(Nq, KDT, false);
        if (marine) need = std::max(need, std::min(fast_smem_need(Nq, KDT, true), (size_t)dyn_max));
        ctx->dyn = KDT ? (int)std::min((size_t)dyn_max, rup(need, compact ? 128 : 1024)) : 0;
        if (KDT && (p->opt_debug & 1)) ctx->dyn = dyn_max;     // experiment: one CTA per SM whatever the mesh size
    }
    CUDA_OK(cudaMallocHost((void **)&ctx->h_stage, (size_t)bl_ctx::NSLOT * STOPS_CAP * 12));
    for (int i = 0; i < bl_ctx::NSLOT; ++i) CUDA_OK(cudaEventCreateWithFlags(&ctx->slot_ev[i], cudaEventDisableTiming));
    ctx->slot = 0;
    ctx->device = device; ctx->B = B; ctx->ws = ws; ctx->ws_bytes = L.total;
    ctx->d_stops = (double *)(ws + L.off_stops);
    ctx->d_ckpt = (int *)(ws + L.off_ckpt);
    ctx->stops_cap = STOPS_CAP;
    ctx->launches = 0;
    ctx->const_bytes = L.off_rain;
    {   // the mesh constants are read by every CTA all the time while ~3.5 MB of per-proposal scratch per CTA streams
        // through L2: pin them (persisting L2 carve-out + access-policy window set per launch)
        int maxp = 0;
        cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, device);
        if (maxp > 0) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min((size_t)maxp, rup(ctx->const_bytes, 1 << 20)));
        cudaGetLastError();
    }
    CUDA_OK(cudaMemset(ws + L.off_rain, 0, L.off_stops - L.off_rain));
    *out = ctx;
    return BL_OK;
}

extern "C" int bl_ctx_destroy(bl_ctx *ctx)
{
    if (ctx) {
        if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
        for (int i = 0; i < bl_ctx::NSLOT; ++i) if (ctx->slot_ev[i]) cudaEventDestroy(ctx->slot_ev[i]);
    }
    delete ctx;
    return BL_OK;
}

extern "C" int bl_reset(bl_ctx *ctx, const double *rain_dev, const double *erod_dev, const uint64_t *seed_dev,
                        double tStart, void *stream)
{
    if (!ctx || !rain_dev || !erod_dev) { g_err = "bl_reset: null argument"; return BL_ERR_ARG; }
    CUDA_OK(cudaSetDevice(ctx->device));
    CUDA_OK(bl_launch_reset_v256(ctx->c, ctx->B, (cudaStream_t)stream, rain_dev, erod_dev,
                                 (const unsigned long long *)seed_dev, tStart));
    ctx->launches++;
    return BL_OK;
}

static int run_common(bl_ctx *ctx, int nstops, const double *stops, const int32_t *ckpt, double *out_pts,
                      double *out_sse, double *out_elev, double *out_erdp, int nck, int single, cudaStream_t st)
{
    if (!ctx || nstops <= 0 || !stops) { g_err = "bl_run: bad arguments"; return BL_ERR_ARG; }
    if (nstops > ctx->stops_cap) { g_err = "bl_run: too many stops"; return BL_ERR_ARG; }
    if (nck < 0 || nck > BL_NCKPT_MAX) { g_err = "bl_run: nck out of range"; return BL_ERR_ARG; }
    CUDA_OK(cudaSetDevice(ctx->device));
    // stage stops / checkpoint ids in the next pinned slot (waits only if that slot's previous copy is still in flight)
    const int slot = ctx->slot;
    ctx->slot = (slot + 1) % bl_ctx::NSLOT;
    CUDA_OK(cudaEventSynchronize(ctx->slot_ev[slot]));
    double *h_stops = (double *)(ctx->h_stage + (size_t)slot * STOPS_CAP * 12);
    int *h_ck = (int *)(h_stops + STOPS_CAP);
    int marked = 0;
    for (int i = 0; i < nstops; ++i) {
        h_stops[i] = stops[i];
        const int k = ckpt ? ckpt[i] : -1;
        if (k >= BL_NCKPT_MAX) { g_err = "bl_run: checkpoint index too large"; return BL_ERR_ARG; }
        if (k >= 0) {
            // grids are written at [(b * nck + <number of checkpoints seen>) * G]: more marked stops than nck, or an id
            // beyond nck, would write out of bounds (mirrors bl_grid_run)
            if ((out_elev || out_erdp) && (k >= nck || marked >= nck)) {
                g_err = "bl_run: more checkpoints than nck grids"; return BL_ERR_ARG;
            }
            ++marked;
        }
        h_ck[i] = k;
    }
    CUDA_OK(cudaMemcpyAsync(ctx->d_stops, h_stops, (size_t)nstops * 8, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(ctx->d_ckpt, h_ck, (size_t)nstops * 4, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaEventRecord(ctx->slot_ev[slot], st));
    if (ctx->l2_persist) {
        cudaStreamAttrValue av;
        memset(&av, 0, sizeof(av));
        av.accessPolicyWindow.base_ptr = ctx->ws;
        av.accessPolicyWindow.num_bytes = ctx->const_bytes;
        av.accessPolicyWindow.hitRatio = 1.0f;
        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av);
        cudaGetLastError();
    }
    RunArgs a;
    a.nstops = nstops; a.stops = ctx->d_stops; a.ckpt = ctx->d_ckpt;
    a.out_pts = out_pts; a.out_sse = out_sse; a.out_elev = out_elev; a.out_erdp = out_erdp;
    a.nck = nck; a.single_step = single;
    if (ctx->tpb == 256) CUDA_OK(bl_launch_run_v256(ctx->c.KDT, ctx->c, ctx->B, ctx->dyn, st, a));
    else if (ctx->tpb == 512) CUDA_OK(bl_launch_run_v512(ctx->c.KDT, ctx->c, ctx->B, ctx->dyn, st, a));
    else if (ctx->tpb == 768) CUDA_OK(bl_launch_run_v768(ctx->c.KDT, ctx->c, ctx->B, ctx->dyn, st, a));
    else if (ctx->tpb == 384) CUDA_OK(bl_launch_run_v384(ctx->c.KDT, ctx->c, ctx->B, ctx->dyn, st, a));
    else CUDA_OK(bl_launch_run_v1024(ctx->c.KDT, ctx->c, ctx->B, ctx->dyn, st, a));
    ctx->launches++;
    return BL_OK;
}

extern "C" int bl_run(bl_ctx *ctx, int nstops, const double *stops, const int32_t *ckpt, double *out_pts,
                      double *out_sse, double *out_elev, double *out_erdp, int nck, void *stream)
{
    return run_common(ctx, nstops, stops, ckpt, out_pts, out_sse, out_elev, out_erdp, nck, 0, (cudaStream_t)stream);
}

extern "C" int bl_step(bl_ctx *ctx, double tStop, void *stream)
{
    return run_common(ctx, 1, &tStop, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 1, (cudaStream_t)stream);
}

extern "C" int bl_get_field(bl_ctx *ctx, int field, int proposal, void *dst, void *stream)
{
    if (!ctx || !dst || proposal < 0 || proposal >= ctx->B) { g_err = "bl_get_field: bad arguments"; return BL_ERR_ARG; }
    CUDA_OK(cudaSetDevice(ctx->device));
    const int N = ctx->c.N;
    const size_t Np = np_of(N);
    char *base = ctx->c.prop_base + (size_t)proposal * ctx->c.prop_stride;
    double *d = (double *)base;
    unsigned long long *u = (unsigned long long *)(d + (size_t)NF64 * Np);
    int *ip = (int *)(u + 4 * Np + keys_len(N));
    const void *src = nullptr;
    size_t bytes = 0;
    int fi = -1, ii = -1;
    switch (field) {
        case BL_F_Z: fi = F_Z; break;
        case BL_F_CUMDIFF: fi = F_CUMDIFF; break;
        case BL_F_CUMHILL: fi = F_CUMHILL; break;
        case BL_F_FILLH: fi = F_FILLH; break;
        case BL_F_MAXH: fi = F_MAXH; break;
        case BL_F_Q: fi = F_Q; break;
        case BL_F_PITVOL: fi = F_PITVOL; break;
        case BL_F_ERO: fi = F_ERO; break;
        case BL_F_DEPO: fi = F_DEPO; break;
        case BL_F_SEDFLUX: fi = F_SEDFLUX; break;
        case BL_F_RCV: ii = I_RCV; break;
        case BL_F_RCV1: ii = I_RCV1; break;
        case BL_F_STACK: ii = I_STACK; break;
        case BL_F_STACK1: ii = I_STACK1; break;
        case BL_F_PITID: ii = I_PITID; break;
        case BL_F_PITDRAIN: ii = I_PITDRAIN; break;
        case BL_F_ALLDRAIN: ii = I_ALLDRAIN; break;
        case 100: src = u + 4 * Np; bytes = (size_t)N * 8; break;     // diagnostics: the key scratch (trace builds)
        default: g_err = "bl_get_field: unknown field"; return BL_ERR_ARG;
    }
    if (src) { }
    else if (fi >= 0) { src = d + (size_t)fi * Np; bytes = (size_t)N * 8; }
    else { src = ip + (size_t)ii * Np; bytes = (size_t)N * 4; }
    CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
    return BL_OK;
}

extern "C" int bl_get_scalars(bl_ctx *ctx, double *tnow, int64_t *steps, uint32_t *status, double *last_dt,
                              void *stream)
{
    if (!ctx) { g_err = "bl_get_scalars: null ctx"; return BL_ERR_ARG; }
    CUDA_OK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t B = ctx->B;
    if (tnow) CUDA_OK(cudaMemcpyAsync(tnow, ctx->c.tnow, B * 8, cudaMemcpyDeviceToHost, st));
    if (steps) CUDA_OK(cudaMemcpyAsync(steps, ctx->c.steps, B * 8, cudaMemcpyDeviceToHost, st));
    if (status) CUDA_OK(cudaMemcpyAsync(status, ctx->c.status, B * 4, cudaMemcpyDeviceToHost, st));
    if (last_dt) CUDA_OK(cudaMemcpyAsync(last_dt, ctx->c.lastdt, B * 8, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    return BL_OK;
}

extern "C" int bl_get_phase_cycles(bl_ctx *ctx, int64_t *dst, void *stream)
{
    if (!ctx || !dst) { g_err = "bl_get_phase_cycles: null argument"; return BL_ERR_ARG; }
    CUDA_OK(cudaSetDevice(ctx->device));
    CUDA_OK(cudaMemcpyAsync(dst, ctx->c.phase_cyc, (size_t)ctx->B * NPHASE * 8, cudaMemcpyDeviceToHost,
                            (cudaStream_t)stream));
    CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
    return BL_OK;
}


// ------------------------------------------------------------------------------------------------
// per-stage entry points on caller-owned device arrays: proposal slot 0 is the scratch landscape
// ------------------------------------------------------------------------------------------------
namespace {
struct Slot0 {
    double *f[NF64];
    int *i[NI32];
};
Slot0 slot0(const bl_ctx *ctx)
{
    Slot0 s;
    const size_t Np = np_of(ctx->c.N);
    double *d = (double *)ctx->c.prop_base;
    for (int k = 0; k < NF64; ++k) s.f[k] = d + (size_t)k * Np;
    unsigned long long *u = (unsigned long long *)(d + (size_t)NF64 * Np);
    int *ip = (int *)(u + 4 * Np + keys_len(ctx->c.N));
    for (int k = 0; k < NI32; ++k) s.i[k] = ip + (size_t)k * Np;
    return s;
}
int stage_launch(bl_ctx *ctx, const StageArgs &a, cudaStream_t st)
{
    // the 256-thread unit serves the contexts that run 256-thread CTAs, the 1024-thread unit all others; the stage
    // kernels use the 17-bytes-per-node layout (or the generic path when that does not fit the unit)
    const bool small = (ctx->tpb == 256);
    int KDT = ctx->c.KDT, dyn = 0;
    if (KDT) {
        cudaFuncAttributes fa;
        memset(&fa, 0, sizeof(fa));
        int optin = 0;
        CUDA_OK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
        CUDA_OK(small ? bl_func_attr_v256(KDT, &fa) : bl_func_attr_v1024(KDT, &fa));
        const int avail = (optin - (int)fa.sharedSizeBytes) / 1024 * 1024;
        if (ctx->need_full > (size_t)avail) KDT = 0;
        else dyn = (int)std::min((size_t)avail, std::max(ctx->need_full, (size_t)(ctx->compact ? 0 : ctx->dyn)));
    }
    if (small) CUDA_OK(bl_launch_stage_v256(KDT, ctx->c, dyn, st, a));
    else CUDA_OK(bl_launch_stage_v1024(KDT, ctx->c, dyn, st, a));
    ctx->launches++;
    return BL_OK;
}
StageArgs stage_args(int stage)
{
    StageArgs a;
    memset(&a, 0, sizeof(a));
    a.stage = stage;
    a.sea = -1.e6;
    return a;
}
#define D2D(dst, src, bytes) CUDA_OK(cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyDeviceToDevice, st))
#define STAGE_PROLOGUE(name)                                                                 \
    if (!ctx) { g_err = name ": null context"; return BL_ERR_ARG; }                          \
    CUDA_OK(cudaSetDevice(ctx->device));                                                     \
    cudaStream_t st = (cudaStream_t)stream;                                                  \
    const size_t N = (size_t)ctx->c.N;                                                       \
    Slot0 s = slot0(ctx);                                                                    \
    (void)N; (void)s
} // namespace

extern "C" int bl_fill(bl_ctx *ctx, const double *elev, double sealevel, double *fillH, void *stream)
{
    STAGE_PROLOGUE("bl_fill");
    if (!elev || !fillH) { g_err = "bl_fill: null argument"; return BL_ERR_ARG; }
    D2D(s.f[F_Z], elev, N * 8);
    StageArgs a = stage_args(BL_STAGE_FILL);
    a.sea = sealevel;
    if (int rc = stage_launch(ctx, a, st)) return rc;
    D2D(fillH, s.f[F_FILLH], N * 8);
    return BL_OK;
}

extern "C" int bl_sfd(bl_ctx *ctx, const double *fillH, const double *elev, int32_t *base, int32_t *rcv, double *maxh,
                      double *maxdep, void *stream)
{
    STAGE_PROLOGUE("bl_sfd");
    if (!elev || !rcv) { g_err = "bl_sfd: null argument"; return BL_ERR_ARG; }
    D2D(s.f[F_Z], elev, N * 8);
    D2D(s.f[F_FILLH], fillH ? fillH : elev, N * 8);
    StageArgs a = stage_args(BL_STAGE_SFD);
    if (int rc = stage_launch(ctx, a, st)) return rc;
    // receivers on the first array given: fillH (sfd.directions) or, without it, elev (sfd.directions_base)
    D2D(rcv, fillH ? s.i[I_RCV] : s.i[I_RCV1], N * 4);
    if (maxh) D2D(maxh, s.f[F_MAXH], N * 8);
    if (maxdep) D2D(maxdep, s.f[F_CAPH], N * 8);
    if (base) {
        // pyBase: the self-receivers, -1 elsewhere (sfd.c:99-101)
        std::vector<int> h(N);
        CUDA_OK(cudaMemcpyAsync(h.data(), fillH ? s.i[I_RCV] : s.i[I_RCV1], N * 4, cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaStreamSynchronize(st));
        for (size_t i = 0; i < N; ++i) h[i] = (h[i] == (int)i) ? (int)i : -1;
        CUDA_OK(cudaMemcpyAsync(base, h.data(), N * 4, cudaMemcpyHostToDevice, st));
        CUDA_OK(cudaStreamSynchronize(st));
    }
    return BL_OK;
}

extern "C" int bl_stack(bl_ctx *ctx, const int32_t *rcv, int shuffle, uint64_t seed, int64_t step, int32_t *stack,
                        int32_t *base_out, int32_t *nbase_out, void *stream)
{
    STAGE_PROLOGUE("bl_stack");
    if (!rcv || !stack) { g_err = "bl_stack: null argument"; return BL_ERR_ARG; }
    D2D(s.i[I_RCV], rcv, N * 4);
    StageArgs a = stage_args(BL_STAGE_STACK);
    a.shuffle = shuffle; a.seed = seed; a.step = step;
    if (int rc = stage_launch(ctx, a, st)) return rc;
    D2D(stack, s.i[I_STACK], N * 4);
    if (base_out) D2D(base_out, s.i[I_LB], N * 4);
    if (nbase_out) D2D(nbase_out, s.i[I_MARK], 4);
    return BL_OK;
}

extern "C" int bl_discharge(bl_ctx *ctx, const int32_t *rcv, double rain, double *Q, void *stream)
{
    STAGE_PROLOGUE("bl_discharge");
    if (!rcv || !Q) { g_err = "bl_discharge: null argument"; return BL_ERR_ARG; }
    D2D(s.i[I_RCV], rcv, N * 4);
    StageArgs a = stage_args(BL_STAGE_DISCHARGE);
    a.rain = rain;
    if (int rc = stage_launch(ctx, a, st)) return rc;
    D2D(Q, s.f[F_Q], N * 8);
    return BL_OK;
}

extern "C" int bl_basins(bl_ctx *ctx, const int32_t *rcv1, const double *elev, const double *fillH, int32_t *pitID,
                         double *pitVol, void *stream)
{
    STAGE_PROLOGUE("bl_basins");
    if (!rcv1 || !elev || !fillH || !pitID || !pitVol) { g_err = "bl_basins: null argument"; return BL_ERR_ARG; }
    D2D(s.i[I_RCV1], rcv1, N * 4);
    D2D(s.f[F_Z], elev, N * 8);
    D2D(s.f[F_FILLH], fillH, N * 8);
    StageArgs a = stage_args(BL_STAGE_BASINS);
    if (int rc = stage_launch(ctx, a, st)) return rc;
    D2D(pitID, s.i[I_PITID], N * 4);
    D2D(pitVol, s.f[F_PITVOL], N * 8);
    return BL_OK;
}

extern "C" int bl_drainage(bl_ctx *ctx, const int32_t *pitID, const double *pitVol, const int32_t *rcv, const double *fillH,
                           double sealevel, int32_t *pitDrain, int32_t *allDrain, void *stream)
{
    STAGE_PROLOGUE("bl_drainage");
    if (!pitID || !pitVol || !rcv || !fillH || !pitDrain || !allDrain) { g_err = "bl_drainage: null argument"; return BL_ERR_ARG; }
    D2D(s.i[I_PITID], pitID, N * 4);
    D2D(s.f[F_PITVOL], pitVol, N * 8);
    D2D(s.i[I_RCV], rcv, N * 4);
    D2D(s.f[F_FILLH], fillH, N * 8);
    StageArgs a = stage_args(BL_STAGE_DRAINAGE);
    a.sea = sealevel;
    if (int rc = stage_launch(ctx, a, st)) return rc;
    D2D(pitDrain, s.i[I_PITDRAIN], N * 4);
    D2D(allDrain, s.i[I_ALLDRAIN], N * 4);
    return BL_OK;
}

extern "C" int bl_streampower(bl_ctx *ctx, const int32_t *rcv, const int32_t *pitID, const double *pitVol,
                              const int32_t *pitDrain, const double *maxh, const double *Q, const double *fillH,
                              const double *elev, double erod, double sealevel, double dt, int shuffle, uint64_t seed,
                              int64_t step, double *depo, double *ero, void *stream)
{
    STAGE_PROLOGUE("bl_streampower");
    if (!rcv || !pitID || !pitVol || !pitDrain || !maxh || !Q || !fillH || !elev || !depo || !ero) {
        g_err = "bl_streampower: null argument"; return BL_ERR_ARG;
    }
    D2D(s.i[I_RCV], rcv, N * 4);
    D2D(s.i[I_PITID], pitID, N * 4);
    D2D(s.f[F_PITVOL], pitVol, N * 8);
    D2D(s.i[I_PITDRAIN], pitDrain, N * 4);
    D2D(s.f[F_MAXH], maxh, N * 8);
    D2D(s.f[F_FILLH], fillH, N * 8);
    D2D(s.f[F_Z], elev, N * 8);
    StageArgs a = stage_args(BL_STAGE_STREAMPOWER);
    a.erod = erod; a.sea = sealevel; a.dt = dt; a.shuffle = shuffle; a.seed = seed; a.step = step; a.q_in = Q;
    if (int rc = stage_launch(ctx, a, st)) return rc;
    D2D(depo, s.f[F_DEPO], N * 8);
    D2D(ero, s.f[F_ERO], N * 8);
    return BL_OK;
}

extern "C" int bl_diffuse(bl_ctx *ctx, const double *elev, double *diff, void *stream)
{
    STAGE_PROLOGUE("bl_diffuse");
    if (!elev || !diff) { g_err = "bl_diffuse: null argument"; return BL_ERR_ARG; }
    D2D(s.f[F_Z], elev, N * 8);
    CUDA_OK(cudaMemsetAsync(s.f[F_CUMDIFF], 0, N * 8, st));
    CUDA_OK(cudaMemsetAsync(s.f[F_CUMHILL], 0, N * 8, st));
    StageArgs a = stage_args(BL_STAGE_DIFFUSE);
    a.diff_out = diff;
    return stage_launch(ctx, a, st);
}

extern "C" int bl_interp(bl_ctx *ctx, const double *v, double *grid, void *stream)
{
    STAGE_PROLOGUE("bl_interp");
    if (!v || !grid || ctx->c.G <= 0) { g_err = "bl_interp: null argument or no grid table"; return BL_ERR_ARG; }
    StageArgs a = stage_args(BL_STAGE_INTERP);
    a.v_in = v; a.grid_out = grid;
    return stage_launch(ctx, a, st);
}

extern "C" int64_t bl_launch_count(const bl_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int bl_ctx_info(const bl_ctx *ctx, int32_t info[4])
{
    if (!ctx || !info) { g_err = "bl_ctx_info: null argument"; return BL_ERR_ARG; }
    CUDA_OK(cudaSetDevice(ctx->device));
    info[0] = ctx->tpb;
    info[1] = ctx->c.KDT ? (ctx->compact ? 2 : 1) : 0;
    info[2] = ctx->dyn;
    int nres = 0;
    if (ctx->tpb == 256) CUDA_OK(bl_occupancy_v256(ctx->c.KDT, ctx->dyn, &nres));
    else if (ctx->tpb == 512) CUDA_OK(bl_occupancy_v512(ctx->c.KDT, ctx->dyn, &nres));
    else if (ctx->tpb == 768) CUDA_OK(bl_occupancy_v768(ctx->c.KDT, ctx->dyn, &nres));
    else if (ctx->tpb == 384) CUDA_OK(bl_occupancy_v384(ctx->c.KDT, ctx->dyn, &nres));
    else CUDA_OK(bl_occupancy_v1024(ctx->c.KDT, ctx->dyn, &nres));
    info[3] = nres;
    return BL_OK;
}


/* bl_oracle.c - CPU ORACLE (test infrastructure, see bl_oracle.h).  Pinned: the sfd.c routines (against
 * oracle/_ref/libsfd_ref.so and tests/golden) and, by the hand-derived fixtures of tests/micro.py, fillPD, fstack.build,
 * basinparameters, basindrainage(all), discharge, streampower + pit cascade, the land-pit fill, marine_distribution (F4a / F4b)
 * and diffmarine with its sub-step loop (F5), getid1 + the re-run of the step (F6), the marine branch of the pit cascade (F7),
 * flowcfl with n != 1 (F8), the eps-puddle on a boundary tree (F9, oracle level).  No routine of the path is left "parity unpinned"; the fixtures are 1-D chains, so neighbour-order
 * effects on a real TIN are pinned only where the reference's own sfd.c covers them (receivers, diffusion).
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (no -ffast-math), mirroring
 * pyBadlands/libUtils/Makefile:47 (f2py modules: -O3 -ftree-vectorize, no fast-math, no FMA).
 *
 * Restatement conventions: 0-based ids everywhere (the Fortran adds/subtracts 1 at its boundary);
 * one rock type (pyRockNb = 1), incisiontype = 0, bedslptype = 0, slp_cr = 0, no rivers - the only
 * combination the five BASELINE configs reach.
 */
#include "bl_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* diagnostics (not part of the algorithm): pit-cascade event / iteration counters */
long or_diag_cascade_events = 0, or_diag_cascade_iters = 0, or_diag_marine_iters = 0, or_diag_marine_calls = 0,
     or_diag_marine_nodes = 0, or_diag_msub = 0, or_diag_cascade_marine_pass = 0;

/* ------------------------------------------------------------------------------------------ */
/* libUtils/sfd.c                                                                              */
/* ------------------------------------------------------------------------------------------ */

/* sfd.c:113-153 directions_base: receiver = lowest neighbour on the real surface (strict <). */
void or_directions_base(const or_mesh *m, const double *z, int *base, int *rcv)
{
    for (int g = 0; g < m->N; ++g) {
        const int *nb = m->ngb + (size_t)g * OR_KP;
        int low = g;
        for (int p = 0; p < OR_KP; ++p) {
            int j = nb[p];
            if (j < 0) break;
            if (z[j] < z[low]) low = j;
        }
        rcv[g] = low;
        base[g] = (low == g) ? g : -1;
    }
}

/* sfd.c:55-111 directions: receiver on the filled surface; maxh = smallest non-negative rise to a
 * neighbour on the real surface (0 if none), maxdep = largest rise. */
void or_directions(const or_mesh *m, const double *fillH, const double *z, int *base, int *rcv,
                   double *maxh, double *maxdep)
{
    for (int g = 0; g < m->N; ++g) {
        const int *nb = m->ngb + (size_t)g * OR_KP;
        int low = g;
        double dH = 1.e6, dD = 0.;
        for (int p = 0; p < OR_KP; ++p) {
            int j = nb[p];
            if (j < 0) break;
            if (fillH[j] < fillH[low]) low = j;
            double dh = z[j] - z[g];
            if (dh >= 0. && dh < dH) dH = dh;
            if (dh > dD) dD = dh;
        }
        rcv[g] = low;
        base[g] = (low == g) ? g : -1;
        if (dH > 9.99e5) dH = 0.;
        maxh[g] = dH;
        maxdep[g] = dD;
    }
}

/* sfd.c:155-185 diffusion: FV flux sum; towards a non-border2 neighbour only downhill. */
void or_diffusion(const or_mesh *m, const double *z, double *diff)
{
    const int *bord = m->borders2;
    for (int g = 0; g < m->N; ++g) {
        double acc = 0.;
        if (bord[g] > 0) {
            const int *nb = m->ngb + (size_t)g * OR_KP;
            const double *ed = m->vor + (size_t)g * OR_KP;
            const double *di = m->edge + (size_t)g * OR_KP;
            for (int p = 0; p < OR_KP; ++p) {
                int j = nb[p];
                if (j < 0) break;
                if (bord[j] > 0 && di[p] > 0.) acc += ed[p] * (z[j] - z[g]) / di[p];
                if (bord[j] < 1) {
                    if (z[j] < z[g] && di[p] > 0.) acc += ed[p] * (z[j] - z[g]) / di[p];
                }
            }
        }
        diff[g] = acc;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* libUtils/PDalgo.f90: pitfilling -> initialisePD (:19-35) + fillPD (:37-91)                  */
/* ------------------------------------------------------------------------------------------ */
long or_pitfilling(const or_mesh *m, const or_params *p, const double *z, double sea, double *W,
                   int *d1, int *d2)
{
    const int N = m->N, bds = m->bounds;
    const double eps = p->eps, TH = p->fill_th;
    int s1 = 0, s2 = 0;
    long sweeps = 0;
    for (int k = 0; k < bds; ++k) W[k] = z[k];
    for (int k = bds; k < N; ++k) { W[k] = 1.e6; d1[s1++] = k; }
    int change = 1;
    while (change) {
        change = 0;
        ++sweeps;
        for (int n = 0; n < s1; ++n) {
            int k = d1[n];
            if (W[k] > z[k]) {
                double hmin = 2.e6;
                const int *nb = m->ngb + (size_t)k * OR_KP;
                for (int q = 0; q < OR_KP; ++q) {
                    if (nb[q] < 0) break;
                    hmin = fmin(hmin, W[nb[q]]);
                }
                if (z[k] >= hmin + eps) {
                    W[k] = z[k];
                    if (p->fill_mode == 1) change = 1;
                } else {
                    if (W[k] > hmin + eps) {
                        W[k] = hmin + eps;
                        if (z[k] >= sea) {
                            if (W[k] - z[k] > TH) { W[k] = z[k] + TH; if (p->fill_mode == 1) change = 1; }
                            else change = 1;
                        } else {
                            if (W[k] - sea > TH) { W[k] = sea + TH; if (p->fill_mode == 1) change = 1; }
                            else change = 1;
                        }
                    }
                    d2[s2++] = k;
                }
            }
        }
        if (s2 > 0) { s1 = s2; memcpy(d1, d2, (size_t)s2 * sizeof(int)); s2 = 0; }
    }
    return sweeps;
}

/* ------------------------------------------------------------------------------------------ */
/* numpy.random.shuffle(self.base) (flowNetwork.py:316).  The reference seeds numpy from OS entropy on
 * every load_xml (model.py:68-74), so no fixed stream exists to match.  Protocol here:
 * a uniformly random permutation defined by counter-based keys, key = hash(seed, step, node id),
 * bases ordered by (key32, id) ascending.  The GPU path implements the same definition. */
static inline uint64_t mix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
static int cmp_u64(const void *a, const void *b)
{
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return (x > y) - (x < y);
}
void or_shuffle_base(int *base, int nb, uint64_t seed, uint64_t step, uint64_t *keys)
{
    uint64_t s = mix64(seed + step);
    for (int i = 0; i < nb; ++i)
        keys[i] = ((mix64(s ^ (uint64_t)(uint32_t)base[i]) >> 32) << 32) | (uint32_t)base[i];
    qsort(keys, (size_t)nb, sizeof(uint64_t), cmp_u64);
    for (int i = 0; i < nb; ++i) base[i] = (int)(uint32_t)(keys[i] & 0xFFFFFFFFull);
}

/* ------------------------------------------------------------------------------------------ */
/* flowNetwork._donors_number_array/_delta_array (flowNetwork.py:330-378) + fstack.build
 * (FLWnetwork.f90:21-71) + flwstack.addtostack (FLOWstack.f90:24-40): donor lists in ascending id,
 * pre-order depth-first walk from each base in the given base order. */
void or_stack_build(int N, const int *base, int nb, const int *rcv, int *delta, int *donors,
                    int *stack, int *tmp)
{
    /* delta[i] = N - sum_{j>=i} ndon[j] = exclusive prefix sum of donor counts */
    memset(delta, 0, (size_t)(N + 1) * sizeof(int));
    for (int k = 0; k < N; ++k) delta[rcv[k] + 1]++;
    for (int k = 0; k < N; ++k) delta[k + 1] += delta[k];
    int *fillc = tmp; /* [N] */
    memset(fillc, 0, (size_t)N * sizeof(int));
    for (int k = 0; k < N; ++k) { int r = rcv[k]; donors[delta[r] + fillc[r]++] = k; }
    /* iterative form of the recursion: cursor per node */
    int *cur = tmp;            /* reuse: next donor slot to visit */
    int *path = tmp + N;       /* [N] explicit recursion stack */
    for (int k = 0; k < N; ++k) cur[k] = delta[k];
    for (int k = 0; k < N; ++k) stack[k] = -1;
    int j = 0;
    for (int p = 0; p < nb; ++p) {
        int root = base[p];
        stack[j++] = root;
        int depth = 0;
        path[depth++] = root;
        while (depth > 0) {
            int v = path[depth - 1];
            if (cur[v] < delta[v + 1]) {
                int d = donors[cur[v]++];
                if (d == v) continue; /* allocs(d) == base: only a base's self-loop is ever skipped */
                stack[j++] = d;
                path[depth++] = d;
            } else --depth;
        }
    }
}

/* FLOWalgo.f90:53-81 discharge */
void or_discharge(int n, const int *stack, const int *rcv, const double *z, const double *q0, int N,
                  double *Q, double *lay)
{
    memcpy(Q, q0, (size_t)N * sizeof(double));
    memset(lay, 0, (size_t)N * sizeof(double));
    for (int i = n - 1; i >= 0; --i) {
        int d = stack[i], r = rcv[d];
        if (d != r) Q[r] = Q[r] + Q[d];
        lay[d] = z[d] - z[r];
    }
}

/* FLOWalgo.f90:130-165 basinparameters */
void or_basinparameters(int n, const int *stack1, const int *rcv1, const double *z, const double *fillH,
                        const double *area, int N, int *pitID, double *pitVol)
{
    for (int k = 0; k < N; ++k) { pitID[k] = -1; pitVol[k] = -1.; }
    int pit = -1;
    for (int i = 0; i < n; ++i) {
        int d = stack1[i], r = rcv1[d];
        if (d == r) { pit = d; pitID[d] = pit; }
        if (fillH[d] > z[d] && pit > -1) {
            pitID[d] = pit;
            pitVol[pit] = pitVol[pit] + (fillH[d] - z[d]) * area[d];
        }
    }
}

/* FLOWalgo.f90:167-238 basindrainage.  `chain` is chainDrain[np]. */
void or_basindrainage(int np, const int *orderPits, const int *pitID, const int *rcv, const int *pIDs,
                      const double *fillH, double sea, int N, int *drain, int *chain)
{
    for (int k = 0; k < N; ++k) drain[k] = -1;
    for (int k = 0; k < np; ++k) chain[k] = -1;
    for (int n = 0; n < np; ++n) {
        int nID = pIDs[orderPits[n]];
        int donor = nID;
        int exist = 0;
        if (drain[nID] > -1) exist = 1;
        if (fillH[nID] < sea) { exist = 1; drain[nID] = nID; }
        while (!exist) {
            int r = rcv[donor];
            if (r == donor) {
                drain[nID] = r;
                for (int k = 0; k < np; ++k) chain[k] = -1;
                exist = 1;
            } else if (fillH[r] < sea) {
                drain[nID] = r;
                for (int k = 0; k < np; ++k) chain[k] = -1;
                exist = 1;
            } else if (pitID[r] == -1 || pitID[r] == nID) {
                donor = r;
            } else if (fillH[pitID[r]] < sea) {
                donor = r;
            } else {
                int q = 0, newpit = 1, nd = pitID[r];
                while (q < np && chain[q] >= 0) { if (chain[q] == nd) newpit = 0; ++q; }
                if (newpit && q < np) {
                    chain[q] = nd;
                    drain[nID] = nd;
                    donor = nd;
                    nID = nd;
                } else donor = r;
            }
        }
    }
}

/* FLOWalgo.f90:240-297 basindrainageall */
void or_basindrainageall(int np, const int *orderPits, const int *pitID, const int *rcv, const int *pIDs,
                         int N, int *drain, int *chain)
{
    for (int k = 0; k < N; ++k) drain[k] = -1;
    for (int k = 0; k < np; ++k) chain[k] = -1;
    for (int n = 0; n < np; ++n) {
        int nID = pIDs[orderPits[n]];
        int donor = nID;
        int exist = 0;
        if (drain[nID] > -1) exist = 1;
        while (!exist) {
            int r = rcv[donor];
            if (r == donor) {
                drain[nID] = r;
                for (int k = 0; k < np; ++k) chain[k] = -1;
                exist = 1;
            } else if (pitID[r] == -1 || pitID[r] == nID) {
                donor = r;
            } else {
                int q = 0, newpit = 1, nd = pitID[r];
                while (q < np && chain[q] >= 0) { if (chain[q] == nd) newpit = 0; ++q; }
                if (newpit && q < np) {
                    chain[q] = nd;
                    drain[nID] = nd;
                    donor = nd;
                    nID = nd;
                } else donor = r;
            }
        }
    }
}

/* x**m and x**n as the Fortran evaluates them (libm pow), or the GPU-matching fast paths */
static inline double powm(const or_params *p, double x, double e)
{
    if (p->pow_mode == 1) {
        if (e == 0.5) return sqrt(x);
        if (e == 1.0) return x;
        if (e == 0.0) return 1.0;
    }
    return pow(x, e);
}

/* FLOWalgo.f90:299-330 flowcfl over all nodes (inGIDs = every node at one rank) */
double or_flowcfl(const or_mesh *m, const or_params *p, const int *rcv, const double *elev,
                  const double *Q, double erod)
{
    double cfl = 1.e6;
    for (int d = 0; d < m->N; ++d) {
        int r = rcv[d];
        double dz = elev[d] - elev[r];
        if (d != r && dz > 0. && Q[d] > 0.) {
            double dx = m->xy[2 * d] - m->xy[2 * r], dy = m->xy[2 * d + 1] - m->xy[2 * r + 1];
            double dist = sqrt(dx * dx + dy * dy);
            double tmp = dist / (erod * powm(p, Q[d], p->spl_m) * powm(p, dz / dist, p->spl_n - 1.));
            cfl = fmin(tmp, cfl);
        }
    }
    return cfl;
}

/* FLOWalgo.f90:583-1044 streampower, single rock, detachment-limited.  Returns 0, or 1 if the pit
 * cascade hit the iteration guard (the reference would spin forever there). */
int or_streampower(const or_mesh *m, const or_params *p, int n, const int *stack, const int *rcv,
                   const int *pitID, const double *pitVol1, const int *pitDrain, const double *maxhA,
                   const double *maxdep, const double *Q, const double *fillH, const double *z, double erod,
                   double sea, double dt, double *depo, double *ero, double *sedfl, double *pitVol)
{
    const int N = m->N;
    int status = 0;
    (void)maxdep;
    for (int k = 0; k < N; ++k) { depo[k] = 0.; ero[k] = 0.; sedfl[k] = 0. * dt; pitVol[k] = pitVol1[k]; }
    for (int i = n - 1; i >= 0; --i) {
        int donor = stack[i], recvr = rcv[donor];
        double SPL = 0., totspl = 0.;
        double dh = 0.95 * (z[donor] - z[recvr]);
        if (z[donor] > sea && z[recvr] < sea) dh = 0.99 * (z[donor] - sea);
        if (dh < 0.001) dh = 0.;
        double waterH = fillH[donor] - z[donor];
        double dx = m->xy[2 * donor] - m->xy[2 * recvr], dy = m->xy[2 * donor + 1] - m->xy[2 * recvr + 1];
        double dist = sqrt(dx * dx + dy * dy);
        double totdist = 0.;
        if (recvr != donor && dh > 0.) {
            if (waterH == 0. && fillH[donor] >= sea) {
                double slp = dh / dist;
                if (dist > 0.) {
                    /* :700-714 (frck = 1, bedfrac = 1) */
                    SPL = -erod * 1. * 1. * powm(p, Q[donor], p->spl_m) * powm(p, slp, p->spl_n);
                    totspl = totspl + SPL;
                    if (-totspl * dt > dh) {
                        double frac = dh / (-totspl * dt);
                        SPL = SPL * frac;
                        totspl = -dh;
                    }
                }
            }
        }
        /* :858-868 */
        double maxh = maxhA[donor];
        if (waterH > 0.) maxh = waterH;
        else if (z[donor] < sea) maxh = sea - z[donor];
        maxh = 0.95 * maxh;

        double Qs = 0., erodep = 0., pitDep = 0.;
        if (totspl < 0.) {                                   /* :875-892 erosion */
            erodep = SPL * dt * m->area[donor];
            Qs = -erodep + sedfl[donor];
        } else if (totspl >= 0. && m->area[donor] > 0.) {     /* :895-952 deposition */
            if (waterH > 0. && fillH[donor] > sea) {
                pitDep = sedfl[donor];
                totdist = totdist + pitDep;
            } else if (z[donor] <= sea) {
                erodep = sedfl[donor];
            } else if (maxh > 0. && waterH == 0. && donor != recvr && z[donor] > sea) {
                double totflx = 0. + sedfl[donor];
                if (totflx / m->area[donor] < maxh) {
                    erodep = sedfl[donor];
                } else {
                    double frac = sedfl[donor] / totflx;
                    erodep = frac * maxh * m->area[donor];
                    Qs = sedfl[donor] - erodep;
                }
            } else if (donor == recvr && m->area[donor] > 0.) {
                erodep = sedfl[donor];
            } else {
                Qs = sedfl[donor];
            }
        }
        if (pitDep == 0.) {                                   /* :955-964 */
            sedfl[recvr] = sedfl[recvr] + Qs;
            if (erodep < 0.) ero[donor] = ero[donor] + erodep;
            else depo[donor] = depo[donor] + erodep;
        } else if (pitDep > 0. && pitID[donor] >= 0 && m->area[pitID[donor]] > 0.) { /* :967-1034 */
            int tmpID = pitID[donor], nID = tmpID;
            long guard = 0;
            ++or_diag_cascade_events;
            while (totdist > 0.) {
                if (++guard > 1000000 || tmpID < 0) { status = 1; break; }
                ++or_diag_cascade_iters;
                double tmpdist = 0. + depo[tmpID];
                if (fillH[tmpID] < sea) {
                    if (z[donor] < sea) depo[donor] = depo[donor] + pitDep;
                    else { sedfl[recvr] = sedfl[recvr] + pitDep; if (tmpID != pitID[donor]) ++or_diag_cascade_marine_pass; }
                    totdist = 0.;
                    nID = recvr;
                } else if (tmpdist + totdist <= pitVol[tmpID]) {
                    depo[tmpID] = depo[tmpID] + pitDep;
                    totdist = 0.;
                    nID = tmpID;
                } else if (pitDrain[tmpID] == tmpID) {
                    depo[tmpID] = depo[tmpID] + pitDep;
                    totdist = 0.;
                    nID = tmpID;
                } else {
                    if (m->borders[tmpID] == 0) {
                        totdist = 0.;
                        nID = tmpID;
                    } else if (tmpdist == pitVol[tmpID]) {
                        nID = tmpID;
                    } else {
                        double frac = pitDep / totdist;
                        depo[tmpID] = depo[tmpID] + (pitVol[tmpID] - tmpdist) * frac;
                        pitDep = (totdist - (pitVol[tmpID] - tmpdist)) * frac;
                        double newdist = 0. + pitDep;
                        double totflx = 0. + depo[tmpID];
                        totdist = newdist;
                        pitVol[tmpID] = totflx;
                        nID = tmpID;
                    }
                }
                tmpID = pitDrain[nID];
            }
        }
    }
    return status;
}

/* FLOWalgo.f90:332-388 diffmarine */
void or_diffmarine(const or_mesh *m, const double *z, const double *depoH, const double *coeff, double sea,
                   double maxth, double tstep, double *diff, double *mindt)
{
    const int *bord = m->borders;
    *mindt = tstep;
    for (int g = 0; g < m->N; ++g) {
        double acc = 0.;
        if (bord[g] > 0 && z[g] < sea) {
            const int *nb = m->ngb + (size_t)g * OR_KP;
            const double *ed = m->vor + (size_t)g * OR_KP;
            const double *di = m->edge + (size_t)g * OR_KP;
            for (int q = 0; q < OR_KP; ++q) {
                int j = nb[q];
                if (j < 0) break;
                if (bord[j] > 0) {
                    double flx = ed[q] * (z[j] - z[g]) / di[q];
                    if (depoH[g] > maxth && z[g] > z[j]) acc = acc + coeff[g] * flx;
                    else if (depoH[j] > maxth && z[g] < z[j] && z[j] < sea) acc = acc + coeff[g] * flx;
                } else if (bord[j] < 1) {
                    if (depoH[g] > maxth && z[g] > z[j]) {
                        double flx = ed[q] * (z[j] - z[g]) / di[q];
                        acc = acc + coeff[g] * flx;
                    }
                }
            }
            if (acc < 0. && acc * tstep < -depoH[g]) *mindt = fmin(-depoH[g] / acc, *mindt);
        }
        diff[g] = acc;
    }
}

/* PDalgo.f90:136-269 marine_distribution (single rock).  elev/seadep are [N] work arrays. */
void or_marine_distribution(const or_mesh *m, const or_params *p, const double *elevation, const double *seavol,
                            double sea, const int *depIDs, int nIDs, double *diffsed, double *elev,
                            double *seadep)
{
    const int N = m->N;
    const double diff_res = 1.e-2;
    const int max_it_cyc = 500000;
    memcpy(elev, elevation, (size_t)N * sizeof(double));
    ++or_diag_marine_calls; or_diag_marine_nodes += nIDs;
    for (int mm = 0; mm < p->diffnb; ++mm) {
        for (int k = 0; k < N; ++k) seadep[k] = seavol[k] / (double)(float)p->diffnb;
        for (int k = 0; k < nIDs; ++k) {
            int id = depIDs[k];
            int it = 0;
            for (;;) {
                ++or_diag_marine_iters;
                if (m->borders[id] < 1) { seadep[id] = 0.; break; }
                const int *nb = m->ngb + (size_t)id * OR_KP;
                double maxz = -1.e8;
                for (int q = 0; q < OR_KP; ++q) {
                    if (nb[q] < 0) break;
                    if (maxz < elev[nb[q]]) maxz = elev[nb[q]];
                }
                if (maxz > sea) maxz = sea;
                if (maxz < elev[id]) maxz = elev[id];
                double vol = fmax(0., p->diffprop * (maxz - elev[id]) * m->area[id]);
                if (it > max_it_cyc) { elev[id] = elev[id] + seadep[id] / m->area[id]; seadep[id] = 0.; break; }
                if (seadep[id] / m->area[id] < diff_res) { elev[id] = elev[id] + seadep[id] / m->area[id]; seadep[id] = 0.; break; }
                it = it + 1;
                if (seadep[id] < vol) { elev[id] = elev[id] + seadep[id] / m->area[id]; seadep[id] = 0.; break; }
                else { seadep[id] = seadep[id] - vol; elev[id] = elev[id] + vol / m->area[id]; }
                if (seadep[id] > 0.) {
                    double minz = elev[id];
                    int nid = -1;
                    maxz = -1.e8;
                    for (int q = 0; q < OR_KP; ++q) {
                        if (nb[q] < 0) break;
                        if (minz > elev[nb[q]]) { nid = nb[q]; minz = elev[nid]; }
                        if (maxz < elev[nb[q]]) maxz = elev[nb[q]];
                    }
                    if (nid == -1) {
                        double dh = maxz - elev[id] + 0.1;
                        if (seadep[id] > dh * m->area[id]) {
                            elev[id] = elev[id] + dh;
                            seadep[id] = seadep[id] - dh * m->area[id];
                            nid = depIDs[k];
                            it = 0;
                            if (nid == id) seadep[nid] = seadep[id];
                            else { seadep[nid] = seadep[nid] + seadep[id]; seadep[id] = 0.; }
                            id = depIDs[k];
                        } else { elev[id] = elev[id] + seadep[id] / m->area[id]; seadep[id] = 0.; break; }
                    } else {
                        if (minz > elev[id]) {
                            /* unreachable (nid != -1 implies minz < elev[id]); kept for the record :232-247 */
                            double dh = (minz - elev[id]) + 0.1;
                            if (seadep[id] > dh * m->area[id]) { elev[id] = elev[id] + dh; seadep[id] = seadep[id] - dh * m->area[id]; }
                            else { elev[id] = elev[id] + seadep[id] / m->area[id]; seadep[id] = 0.; break; }
                            nid = depIDs[k];
                            seadep[nid] = seadep[nid] + seadep[id];
                            seadep[id] = 0.;
                            id = nid;
                            it = 0;
                        } else {
                            seadep[nid] = seadep[nid] + seadep[id];
                            seadep[id] = 0.;
                            id = nid;
                        }
                    }
                } else { seadep[id] = 0.; break; }
            }
        }
    }
    for (int k = 0; k < N; ++k) diffsed[k] = elev[k] - elevation[k];
}

/* numpy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src, pairwise_sum_DOUBLE) - the
 * order np.sum uses on a contiguous float64 array; needed for flowNetwork.py:759 and bl_mcmc.py:488 */
double or_pairwise_sum(const double *a, long n)
{
    if (n < 8) {
        double res = 0.;
        for (long i = 0; i < n; ++i) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        long i;
        for (int k = 0; k < 8; ++k) r[k] = a[k];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; ++k) r[k] += a[i + k];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    } else {
        long n2 = n / 2;
        n2 -= n2 % 8;
        return or_pairwise_sum(a, n2) + or_pairwise_sum(a + n2, n - n2);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* state                                                                                       */
/* ------------------------------------------------------------------------------------------ */
static double *dalloc(int n) { return (double *)calloc((size_t)n, sizeof(double)); }
static int *ialloc(int n) { return (int *)calloc((size_t)n, sizeof(int)); }

or_state *or_state_create(int N)
{
    or_state *s = (or_state *)calloc(1, sizeof(or_state));
    s->N = N;
    s->z = dalloc(N); s->cumdiff = dalloc(N); s->cumhill = dalloc(N);
    s->fillH = dalloc(N); s->maxh = dalloc(N); s->maxdep = dalloc(N); s->Q = dalloc(N); s->lay = dalloc(N);
    s->pitVol = dalloc(N);
    s->rcv = ialloc(N); s->rcv1 = ialloc(N); s->stack = ialloc(N); s->stack1 = ialloc(N);
    s->donors = ialloc(N); s->donors1 = ialloc(N); s->pitID = ialloc(N); s->pitDrain = ialloc(N);
    s->allDrain = ialloc(N); s->base = ialloc(N); s->base1 = ialloc(N);
    s->ero = dalloc(N); s->depo = dalloc(N); s->sedflux = dalloc(N);
    s->w0 = dalloc(2 * N); s->w1 = dalloc(N); s->w2 = dalloc(N); s->w3 = dalloc(N); s->w4 = dalloc(N); s->w5 = dalloc(N);
    s->i0 = ialloc(2 * N + 2); s->i1 = ialloc(2 * N + 2); s->i2 = ialloc(N + 1); s->i3 = ialloc(N + 1); s->i4 = ialloc(N + 1);
    return s;
}
void or_state_destroy(or_state *s)
{
    if (!s) return;
    free(s->z); free(s->cumdiff); free(s->cumhill); free(s->fillH); free(s->maxh); free(s->maxdep);
    free(s->Q); free(s->lay); free(s->pitVol); free(s->rcv); free(s->rcv1); free(s->stack); free(s->stack1);
    free(s->donors); free(s->donors1); free(s->pitID); free(s->pitDrain); free(s->allDrain); free(s->base);
    free(s->base1); free(s->ero); free(s->depo); free(s->sedflux);
    free(s->w0); free(s->w1); free(s->w2); free(s->w3); free(s->w4); free(s->w5);
    free(s->i0); free(s->i1); free(s->i2); free(s->i3); free(s->i4);
    free(s);
}
void or_state_reset(or_state *s, const double *z0, double rain, double erod, double tStart)
{
    memcpy(s->z, z0, (size_t)s->N * sizeof(double));
    memset(s->cumdiff, 0, (size_t)s->N * sizeof(double));
    memset(s->cumhill, 0, (size_t)s->N * sizeof(double));
    s->tNow = tStart; s->rain = rain; s->erod = erod; s->step = 0; s->fill_sweeps = 0; s->rerun = 0;
    s->last_dt = 0.; s->last_cfl = 0.;
}

/* forceSim.getSea (forceSim.py:226-245) with scipy interp1d(kind='linear') arithmetic:
 * slope = (y_hi - y_lo)/(x_hi - x_lo); y = slope*(x - x_lo) + y_lo, index by searchsorted(side='left') */
double or_sea(const or_params *p, double t)
{
    if (p->nsea <= 0) return p->seapos;
    const double *x = p->seat, *y = p->seav;
    int n = p->nsea;
    if (t < x[0]) t = x[0];
    if (t > x[n - 1]) t = x[n - 1];
    int lo = 0, hi = n; /* first index with x[idx] >= t */
    while (lo < hi) { int mid = (lo + hi) / 2; if (x[mid] < t) lo = mid + 1; else hi = mid; }
    int idx = lo;
    if (idx < 1) idx = 1;
    if (idx > n - 1) idx = n - 1;
    double slope = (y[idx] - y[idx - 1]) / (x[idx] - x[idx - 1]);
    return slope * (t - x[idx - 1]) + y[idx - 1];
}

/* stable ascending argsort of keys[ids] (insertion; np is small), then reversed:
 * numpy.argsort(fillH[pIDs])[::-1] (flowNetwork.py:563).  Tie order in numpy's introsort is
 * unspecified; we define ties as stable-ascending-then-reversed. */
typedef struct { double key; int idx; } kv_t;
static int cmp_kv(const void *a, const void *b)
{
    const kv_t *x = (const kv_t *)a, *y = (const kv_t *)b;
    if (x->key < y->key) return -1;
    if (x->key > y->key) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx);
}
static void argsort_desc(const double *key, int np, int *order)
{
    kv_t *kv = (kv_t *)malloc(sizeof(kv_t) * (size_t)np);
    for (int i = 0; i < np; ++i) { kv[i].key = key[i]; kv[i].idx = i; }
    qsort(kv, (size_t)np, sizeof(kv_t), cmp_kv);
    for (int i = 0; i < np; ++i) order[i] = kv[np - 1 - i].idx;
    free(kv);
}

/* buildFlux.streamflow (buildFlux.py:21-100) */
void or_streamflow(const or_mesh *m, const or_params *p, or_state *s)
{
    const int N = m->N;
    double sea = or_sea(p, s->tNow);
    s->fill_sweeps += or_pitfilling(m, p, s->z, sea, s->fillH, s->i0, s->i1);
    /* flowNetwork.SFD_receivers (flowNetwork.py:269-328) */
    int *b1 = s->i2, *b = s->i3;
    or_directions_base(m, s->z, b1, s->rcv1);
    or_directions(m, s->fillH, s->z, b, s->rcv, s->maxh, s->maxdep);
    s->nbase1 = 0; s->nbase = 0;
    for (int k = 0; k < N; ++k) { if (b1[k] >= 0) s->base1[s->nbase1++] = b1[k]; }
    for (int k = 0; k < N; ++k) { if (b[k] >= 0) s->base[s->nbase++] = b[k]; }
    if (p->shuffle_mode == 1) or_shuffle_base(s->base, s->nbase, p->seed, (uint64_t)s->step, (uint64_t *)s->w0);
    /* ordered_node_array_filled / _elev (flowNetwork.py:380-416) */
    or_stack_build(N, s->base, s->nbase, s->rcv, s->i4, s->donors, s->stack, s->i0);
    or_stack_build(N, s->base1, s->nbase1, s->rcv1, s->i4, s->donors1, s->stack1, s->i0);
    /* compute_parameters_depression (flowNetwork.py:523-577) */
    or_basinparameters(N, s->stack1, s->rcv1, s->z, s->fillH, m->area, N, s->pitID, s->pitVol);
    int *pIDs = s->i0, np_ = 0;
    for (int k = 0; k < N; ++k) if (s->pitVol[k] >= 0.) pIDs[np_++] = k;
    if (np_ > 0) {
        double *key = s->w0;
        int *order = s->i1;
        for (int i = 0; i < np_; ++i) key[i] = s->fillH[pIDs[i]];
        argsort_desc(key, np_, order);
        or_basindrainage(np_, order, s->pitID, s->rcv, pIDs, s->fillH, sea, N, s->pitDrain, s->i2);
        or_basindrainageall(np_, order, s->pitID, s->rcv, pIDs, N, s->allDrain, s->i2);
    } else {
        for (int k = 0; k < N; ++k) { s->pitDrain[k] = -1; s->allDrain[k] = -1; }
    }
    /* compute_flow (flowNetwork.py:418-443): rain is 0 on ghost nodes (model.py:278-279) */
    double *q0 = s->w0;
    for (int k = 0; k < N; ++k) q0[k] = m->area[k] * ((k >= m->bounds) ? s->rain : 0.);
    or_discharge(N, s->stack, s->rcv, s->z, q0, N, s->Q, s->lay);
}

/* python-2 round(x - 0.5) applied when x > 1 (buildFlux.py:171-172, flowNetwork.py:687-688):
 * C99 round() is half-away-from-zero like py2's. */
static inline double py2_floorish(double x) { return (x > 1.) ? round(x - 0.5) : x; }

/* flowNetwork.compute_sedflux (flowNetwork.py:579-793), single rock */
static int compute_sedflux(const or_mesh *m, const or_params *p, or_state *s, double dt, double sea,
                           double *newdt_out, double *erosion, double *deposition)
{
    const int N = m->N;
    int status = 0;
    double newdt = dt;
    double *cdepo = s->depo, *cero = s->ero, *sedl = s->w1, *pv = s->w2;
    status |= or_streampower(m, p, N, s->stack, s->rcv, s->pitID, s->pitVol, s->pitDrain, s->maxh, s->maxdep,
                             s->Q, s->fillH, s->z, s->erod, sea, newdt, cdepo, cero, sedl, pv);
    /* getid1 (FLOWalgo.f90:1047-1086) + overfilled closed basins (flowNetwork.py:671-685) */
    {
        int nb1 = 0, nb2 = 0;
        double minperc = 0.;
        int have = 0;
        for (int k = 0; k < N; ++k) {
            double vc = (p->depo == 0) ? cero[k] : cdepo[k] + cero[k];
            double sumvol = 0. + vc;
            int in1 = (sumvol > s->pitVol[k] && s->pitVol[k] > 0.);
            int in2 = (s->pitID[k] >= 0 && s->allDrain[k] == s->pitID[k]);
            nb1 += in1; nb2 += in2;
            if (in1 && in2 && m->indomain[k]) {
                double perc = s->pitVol[k] / sumvol;
                if (!have || perc < minperc) { minperc = perc; have = 1; }
            }
        }
        if (nb1 > 0 && nb2 > 0 && have) newdt = dt * minperc;
    }
    newdt = py2_floorish(newdt);
    newdt = fmax(p->minDT, newdt);
    if (newdt < dt) {
        s->rerun++;
        status |= or_streampower(m, p, N, s->stack, s->rcv, s->pitID, s->pitVol, s->pitDrain, s->maxh, s->maxdep,
                                 s->Q, s->fillH, s->z, s->erod, sea, newdt, cdepo, cero, sedl, pv);
    }
    for (int k = 0; k < N; ++k) { erosion[k] = 0.; deposition[k] = 0.; }
    for (int k = 0; k < N; ++k) if (m->borders[k]) erosion[k] = cero[k] / m->area[k];
    if (p->depo != 0) {
        double *dep = s->w3; /* depo restricted to insideIDs, then ndepo */
        for (int k = 0; k < N; ++k) dep[k] = m->borders[k] ? cdepo[k] : 0.;
        /* getids (FLOWalgo.f90:1089-1180) fused with the consumers (flowNetwork.py:741-776) */
        int *landid = s->i0, nland = 0, *seaid = s->i1, nsea = 0;
        double *perc = s->w4;
        for (int k = 0; k < N; ++k) {
            int in_ = 0;
            double sumdep = 0., sumperc = 0.;
            perc[k] = 0.;
            int plain = 0;
            if (s->z[k] > sea && s->fillH[k] == s->z[k]) {
                in_ = 1;
                sumdep = sumdep + dep[k];
                if (sumdep > 0.) plain = 1;
            }
            if (s->z[k] > sea && s->fillH[k] > sea && s->pitVol[k] > 0.) {
                if (sumdep == 0. && in_ == 0) {
                    double nvol = 0.;
                    if (dep[k] / s->pitVol[k] > 1.e-8) nvol = nvol + dep[k];
                    if (dep[k] / s->pitVol[k] > 1.e-8) { perc[k] = dep[k] / nvol; sumperc = sumperc + perc[k]; }
                    else { dep[k] = 0.; perc[k] = 0.; }
                }
                if (sumperc > 0.) {
                    landid[nland++] = k;
                    if (fabs(sumperc - 1.) > 1.e-8) perc[k] = perc[k] / sumperc;
                }
            }
            if (s->z[k] <= sea) {
                sumdep = sumdep + dep[k];
                if (sumdep > 0.) seaid[nsea++] = k;
            }
            if (plain) { deposition[k] = dep[k] / m->area[k]; dep[k] = 0.; } /* :741-744 */
        }
        if (nland > 0) {                                                    /* :750-763 */
            double *tmpv = s->w5;
            int *members = s->i2;
            for (int q = 0; q < nland; ++q) {
                int pit = landid[q], nm = 0;
                for (int k = 0; k < N; ++k) if (s->pitID[k] == pit) members[nm++] = k;
                for (int i = 0; i < nm; ++i) {
                    int k = members[i];
                    deposition[k] = (s->fillH[k] - s->z[k]) * perc[pit];
                    tmpv[i] = deposition[k] * m->area[k];
                }
                double tmpd = or_pairwise_sum(tmpv, nm);
                double dfrac = (0. + dep[pit]) / tmpd;
                for (int i = 0; i < nm; ++i) deposition[members[i]] *= dfrac;
            }
            for (int q = 0; q < nland; ++q) dep[landid[q]] = 0.;
        }
        if (nsea > 0) {                                                     /* :769-776 */
            double *seavol = s->w5, *seadep = s->w1;
            for (int k = 0; k < N; ++k) seavol[k] = 0.;
            for (int q = 0; q < nsea; ++q) seavol[seaid[q]] = dep[seaid[q]];
            or_marine_distribution(m, p, s->z, seavol, sea, seaid, nsea, seadep, s->w2, s->w0);
            for (int k = 0; k < N; ++k) deposition[k] += seadep[k];
            for (int q = 0; q < nsea; ++q) dep[seaid[q]] = 0.;
        }
        int any = 0;                                                        /* :782-785 */
        for (int k = 0; k < N; ++k) if (dep[k] != 0.) { any = 1; break; }
        if (any) for (int k = 0; k < N; ++k) if (m->borders[k]) deposition[k] += dep[k] / m->area[k];
    }
    for (int k = 0; k < N; ++k) s->sedflux[k] = m->borders[k] ? erosion[k] + deposition[k] : 0.;
    *newdt_out = newdt;
    return status;
}

/* buildFlux.sediment_flux (buildFlux.py:102-320) */
int or_sediment_flux(const or_mesh *m, const or_params *p, or_state *s, double tStop)
{
    const int N = m->N;
    int status = 0;
    double sea = or_sea(p, s->tNow);
    /* :169-177 */
    double flowcfl = or_flowcfl(m, p, s->rcv, s->fillH, s->Q, s->erod);
    double cfl = fmin(flowcfl, p->hill_cfl);
    s->last_cfl = cfl;
    cfl = py2_floorish(cfl);
    cfl = fmin(cfl, tStop - s->tNow);
    cfl = fmax(p->minDT, cfl);
    cfl = fmin(p->maxDT, cfl);
    double timestep;
    double *erosion = (double *)malloc(sizeof(double) * (size_t)N), *deposition = (double *)malloc(sizeof(double) * (size_t)N);
    status |= compute_sedflux(m, p, s, cfl, sea, &timestep, erosion, deposition);
    for (int k = 0; k < N; ++k) { s->z[k] += s->sedflux[k]; s->cumdiff[k] += s->sedflux[k]; } /* :193-195 */
    if (p->CDr > 0.) {                                                        /* :198-247 */
        int it = 0;
        double *sumdep = deposition; /* one rock: np.sum(deposition, axis=1) */
        double maxth = 0.1, diffstep = timestep;
        double *coeff = s->w1, *dm = s->w2;
        for (int k = 0; k < N; ++k) {                                         /* diffLinear.sedfluxmarine :184-213 */
            double ac = (m->area[k] > 0.) ? 1. / m->area[k] : 0.;
            coeff[k] = ac * ((s->z[k] >= sea) ? 0. : p->CDr);
        }
        while (diffstep > 0. && it < 1000) {
            double maxstep = fmin(p->ms_cfl, diffstep), mindt;
            or_diffmarine(m, s->z, sumdep, coeff, sea, maxth, maxstep, dm, &mindt);
            for (int k = 0; k < N; ++k) if (!m->borders[k]) dm[k] = 0.;
            maxstep = fmin(mindt, maxstep);
            diffstep -= maxstep;
            ++or_diag_msub;
            for (int k = 0; k < N; ++k) {
                double d = dm[k] * maxstep;
                sumdep[k] += d; s->z[k] += d; s->cumdiff[k] += d;
            }
            ++it;
        }
    }
    /* hillslope (:252-272; diffLinear.sedflux :150-182; sfd.diffusion) */
    {
        double *flux = s->w1;
        or_diffusion(m, s->z, flux);
        for (int k = 0; k < N; ++k) {
            double ac = (m->area[k] > 0.) ? 1. / m->area[k] : 0.;
            double coeff = ac * ((s->z[k] >= sea) ? p->CDa : p->CDm);
            if (!m->borders2[k]) { coeff = 0.; flux[k] = 0.; }
            flux[k] = coeff * flux[k] * timestep;
        }
        for (int k = 0; k < N; ++k) if (m->borders[k]) { s->z[k] += flux[k]; s->cumdiff[k] += flux[k]; s->cumhill[k] += flux[k]; }
    }
    /* boundary (:292-302) */
    if (p->btype == 1) for (int k = 0; k < m->bounds; ++k) s->z[k] = s->z[m->parent[k]] - 0.1;
    else if (p->btype == 2) for (int k = 0; k < m->bounds; ++k) s->z[k] = s->z[m->parent[k]];
    else if (p->btype == 3) for (int k = 0; k < m->bounds; ++k) s->z[k] = s->z[m->parent[k]] + 100.;
    if (p->apply_disp) for (int k = 0; k < N; ++k) s->z[k] += p->disp[k] * timestep;   /* :312-313 */
    s->tNow += timestep;
    s->last_dt = timestep;
    s->step++;
    free(erosion); free(deposition);
    return status;
}

/* Model.run_to_time main loop (model.py:266-490) for a precomputed list of stop times */
int or_run_to(const or_mesh *m, const or_params *p, or_state *s, int nstops, const double *stops)
{
    int status = 0;
    for (int i = 0; i < nstops; ++i) {
        while (s->tNow < stops[i]) {
            or_streamflow(m, p, s);
            status |= or_sediment_flux(m, p, s, stops[i]);
        }
    }
    return status;
}

/* bl_mcmc.interpolateArray (bl_mcmc.py:198-213): np.average(vals, weights=1/d, axis=1) */
void or_interp(int G, const int *idx, const double *w, const int *exact, const double *v, double *out)
{
    for (int g = 0; g < G; ++g) {
        const int *ix = idx + 3 * g;
        const double *ww = w + 3 * g;
        if (exact[g]) { out[g] = v[ix[0]]; continue; }
        double num = ((0. + v[ix[0]] * ww[0]) + v[ix[1]] * ww[1]) + v[ix[2]] * ww[2];
        double den = ((0. + ww[0]) + ww[1]) + ww[2];
        out[g] = num / den;
    }
}

double or_sse(const double *a, const double *b, long n, double *tmp)
{
    for (long i = 0; i < n; ++i) { double d = a[i] - b[i]; tmp[i] = d * d; }
    return or_pairwise_sum(tmp, n);
}

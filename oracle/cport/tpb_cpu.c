/* tpb_cpu.c - CPU restatement (C99 + OpenMP) of the thermalporous hot path.
 *
 * TEST INFRASTRUCTURE / CPU BASELINE - never shipped, never linked into libtpb200.so.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs
 * may load the library built from this file (oracle/_build/libtpb_cpu.so).
 *
 * What it restates (reference = tlroy/thermalporous, all un-vendored maths lives in
 * Firedrake/PETSc/hypre):
 *   - residual + Jacobian of the DG0/TPFA forms: singlephase.py:120-127,226-235;
 *     twophase.py:162-178,333-354; sources singlephase.py:151-165, twophase.py:388-411;
 *     properties physicalparameters.py:37-98; Peaceman rates wellcase.py:171-235
 *   - PETSc MatMult on the block-stencil Jacobian
 *   - CPRStage1PC / CPTRStage1PC decoupling + restriction (preconditioners.py:680-903,1442-1567)
 *   - ConvDiffSchur(TwoPhases)PC operator (preconditioners.py:63-108,225-276)
 *   - PCFIELDSPLIT schur FULL, PCCOMPOSITE multiplicative (singlephase.py:309-351, twophase.py:531-550)
 *   - a geometric-aggregation multigrid V-cycle in the role of hypre BoomerAMG (one V-cycle,
 *     pc_hypre_boomeramg_max_iter 1) and block ILU(0) in red-black ordering in the role of
 *     PETSc bjacobi+ilu(0)  (these two are OUR algorithms - hypre/PETSc are not restated -
 *     so iteration counts differ from the reference; converged fields do not)
 *   - right-preconditioned GMRES / FGMRES with classical Gram-Schmidt (PETSc KSPGMRES defaults)
 *   - SNES newtonls with PETSc's default convergence tests
 *
 * PARITY STATUS: assembly is pinned by tests/golden (vectors made by the reference's own form
 * code, tests/golden/make_golden.py) through the NumPy oracle and directly; the solver stack has
 * no reference vectors (hypre/PETSc absent) - it is pinned on converged fields of the golden time
 * loops (tests/golden/l*.npz) only.
 *
 * Layouts are those of include/tpb200.h (only its struct/enum definitions are used here).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "tpb200.h"

#define MAXF 3
#define NSMAX 7
#define ND 6
#define MAXLEV 40

typedef struct {
    int dim, nx, ny, nz, ns;
    long n;
    double h[3], area[3], vol;
} cgrid;

typedef struct {
    double ko, kw, kr, cw, co, cr, rho_r, T_inj, T_prod, U, g, Wp, Wo;
    double rho_ref, mu_o_pref, mu_o_exp;
} cparams;

/* scalar 5|7-point stencil operator on a structured grid: a[s*n + cell] */
typedef struct {
    int nx, ny, nz;
    long n;
    int cx, cy, cz; /* coarsening factors towards the next level */
    double* a;
    double *x, *b, *r;
    double* xt;          /* line-smoothed levels: the iterate between pre- and post-smoothing (smoothing is out of place) */
    double *xs0, *xs1;   /* ... and two work vectors for the groups of sweeps in between */
    int line;            /* smoothed by zebra z-line Gauss-Seidel (tpb_solver_opts.mg_smoother) */
    double *lid, *llf, *lcp; /* Thomas factors of every column: 1/pivot, lower * 1/pivot, upper * 1/pivot */
} mglevel;

typedef struct {
    int nlev;
    mglevel lev[MAXLEV];
    int last_sweeps; /* > 0: the last level is diagonally dominant (mg_dd_stop) and gets this many sweeps */
} mghier;

/* sweeps for a 1e-3 contraction on a level whose rows have sum|off-diag| <= rho |diag| (csrc/tpb_pc.cu dd_sweeps) */
static int dd_sweeps(double rho, int cap) {
    int k = 1;
    if (rho > 0.0 && rho < 1.0) k = (int)ceil(log(1e-3) / log(rho));
    if (rho >= 1.0) k = cap;   /* (only reachable with mg_dd_stop >= 1) */
    if (k < 1) k = 1;
    if (cap > 0 && k > cap) k = cap;
    return k;
}

typedef struct tpc_handle_s {
    cgrid g;
    cparams P;
    int nphase, nf;
    double* fld[5];
    int nsrc;
    tpb_source* src;
    tpb_solver_opts opts;
    /* pc state */
    const double* J;
    int pc_ready;
    double* w[MAXF];      /* decoupling weights: r_a = x_a - w_a * x_s (CPR: w[1..nf-1] on field 0) */
    double* App;          /* scalar stencil for the pressure block (level-0 operator of mg_p) */
    double* A00;          /* CPTR / fieldsplit: 2x2 block stencil over (p,T): [s][a][b][cell] */
    double* AT;           /* scalar stencil for the temperature Schur operator */
    mghier mg_p, mg_T;
    double* Dinv;         /* nf*nf*n inverted (modified) diagonal blocks of stage 2 */
    double *t0, *t1, *t2, *t3; /* work vectors nf*n */
    /* krylov */
    double *V, *Z;
    int kcap;
    long nlaunch;
} tpc_handle_s;

/* --------------------------------------------------------------------------------------- */
/* properties, physicalparameters.py:37-98                                                  */
/* --------------------------------------------------------------------------------------- */
static inline void oil_rho_d(const cparams* P, double p, double T, double* r, double* r_p, double* r_T) {
    *r = P->rho_ref * exp(5.5e-5 * (p * 10.0 - 1.01325)) * exp(-2.5e-4 * (T - (15.5556 + 273.15)));
    *r_p = 5.5e-4 * (*r);
    *r_T = -2.5e-4 * (*r);
}
static inline void oil_imu_d(const cparams* P, double T, double* im, double* im_T) {
    double Tf = 1.8 * (T - 273.15) + 32.0;
    double mu = P->mu_o_pref * pow(Tf, P->mu_o_exp);
    *im = 1.0 / mu;
    *im_T = -(*im) * P->mu_o_exp * 1.8 / Tf;
}
static inline void water_rho_d(double p, double T, double* r, double* r_p, double* r_T) {
    const double E0 = 999.83952, E1 = 16.955176, E2 = -7.987e-3, E3 = -46.170461e-6, E4 = 105.56302e-9,
                 E5 = -280.54353e-12, E6 = 16.87985e-3, E7 = 10.2, Cw = 3.98854e-4;
    double Tc = T - 272.15;
    double poly = E0 + E1 * Tc + E2 * Tc * Tc + E3 * Tc * Tc * Tc + E4 * Tc * Tc * Tc * Tc + E5 * Tc * Tc * Tc * Tc * Tc;
    double dpoly = E1 + 2 * E2 * Tc + 3 * E3 * Tc * Tc + 4 * E4 * Tc * Tc * Tc + 5 * E5 * Tc * Tc * Tc * Tc;
    double den = 1.0 + E6 * Tc;
    double ex = exp(Cw * (p - E7));
    *r = poly * ex / den;
    *r_p = Cw * (*r);
    *r_T = dpoly * ex / den - poly * ex * E6 / (den * den);
}
static inline void water_imu_d(double T, double* im, double* im_T) {
    const double Aw = 2.1850, Bw = 0.04012, Cw = 5.1547e-6;
    double Tf = 1.8 * (T - 272.15) + 32.0;
    *im = (-1.0 + Bw * Tf + Cw * Tf * Tf) / (1e-3 * Aw);
    *im_T = (Bw + 2.0 * Cw * Tf) * 1.8 / (1e-3 * Aw);
}

/* --------------------------------------------------------------------------------------- */
/* forward-mode duals over the 2*nf unknowns of a face                                      */
/* --------------------------------------------------------------------------------------- */
typedef struct {
    double v, d[ND];
} dual;
static inline dual dc(double a) {
    dual r;
    r.v = a;
    for (int i = 0; i < ND; i++) r.d[i] = 0.0;
    return r;
}
static inline dual dadd(dual a, dual b) {
    dual r;
    r.v = a.v + b.v;
    for (int i = 0; i < ND; i++) r.d[i] = a.d[i] + b.d[i];
    return r;
}
static inline dual dsub(dual a, dual b) {
    dual r;
    r.v = a.v - b.v;
    for (int i = 0; i < ND; i++) r.d[i] = a.d[i] - b.d[i];
    return r;
}
static inline dual dmul(dual a, dual b) {
    dual r;
    r.v = a.v * b.v;
    for (int i = 0; i < ND; i++) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
    return r;
}
static inline dual dscale(double s, dual a) {
    dual r;
    r.v = s * a.v;
    for (int i = 0; i < ND; i++) r.d[i] = s * a.d[i];
    return r;
}
static inline dual ddiv(dual a, dual b) {
    dual r;
    r.v = a.v / b.v;
    for (int i = 0; i < ND; i++) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v;
    return r;
}

/* per-cell properties with partials w.r.t. the cell's own (p, T, S) */
typedef struct {
    double p, T, S;
    double ro[3], rw[3];   /* value, d/dp, d/dT */
    double lo[4], lw[4];   /* value, d/dp, d/dT, d/dS : k_r rho / mu */
    double kT[2];          /* value, d/dS */
} cprops;

static inline void cell_props(const tpc_handle_s* h, const double* u, long c, cprops* q) {
    const cparams* P = &h->P;
    long n = h->g.n;
    double p = u[c], T = u[n + c];
    double imo, imo_T;
    q->p = p;
    q->T = T;
    oil_rho_d(P, p, T, &q->ro[0], &q->ro[1], &q->ro[2]);
    oil_imu_d(P, T, &imo, &imo_T);
    if (h->nf == 3) {
        double S = u[2 * n + c], imw, imw_T, phi = h->fld[TPB_PHI][c];
        q->S = S;
        water_rho_d(p, T, &q->rw[0], &q->rw[1], &q->rw[2]);
        water_imu_d(T, &imw, &imw_T);
        q->lo[0] = S * q->ro[0] * imo;
        q->lo[1] = S * q->ro[1] * imo;
        q->lo[2] = S * (q->ro[2] * imo + q->ro[0] * imo_T);
        q->lo[3] = q->ro[0] * imo;
        q->lw[0] = (1.0 - S) * q->rw[0] * imw;
        q->lw[1] = (1.0 - S) * q->rw[1] * imw;
        q->lw[2] = (1.0 - S) * (q->rw[2] * imw + q->rw[0] * imw_T);
        q->lw[3] = -q->rw[0] * imw;
        q->kT[0] = phi * (S * P->ko + (1.0 - S) * P->kw) + (1.0 - phi) * P->kr; /* twophase.py:135,311 */
        q->kT[1] = phi * (P->ko - P->kw);
    } else {
        q->S = 0.0;
        q->rw[0] = q->rw[1] = q->rw[2] = 0.0;
        q->lo[0] = q->ro[0] * imo;
        q->lo[1] = q->ro[1] * imo;
        q->lo[2] = q->ro[2] * imo + q->ro[0] * imo_T;
        q->lo[3] = 0.0;
        q->lw[0] = q->lw[1] = q->lw[2] = q->lw[3] = 0.0;
        q->kT[0] = h->fld[TPB_KT][c];
        q->kT[1] = 0.0;
    }
}

static inline dual mk3(const double* v4, int off, int nf) { /* v, d/dp, d/dT, d/dS */
    dual r = dc(v4[0]);
    r.d[off] = v4[1];
    r.d[off + 1] = v4[2];
    if (nf == 3) r.d[off + 2] = v4[3];
    return r;
}
static inline dual mk2(const double* v3, int off) { /* v, d/dp, d/dT */
    dual r = dc(v3[0]);
    r.d[off] = v3[1];
    r.d[off + 1] = v3[2];
    return r;
}
static inline dual mkvar(double v, int slot) {
    dual r = dc(v);
    r.d[slot] = 1.0;
    return r;
}
static inline double harm(double a, double b) {
    double s = 0.5 * (a + b);
    return s > 0.0 ? a * b / s : 0.0;
}

/* fluxes through one face, '+' = lower-index cell (slots [0,nf)), '-' = higher (slots [nf,2nf)).
 * f[r] is added to the '+' row and subtracted from the '-' row. */
static void face_flux(const tpc_handle_s* h, const cprops* pl, const cprops* mi, double Kf, double area,
                      double ih, double grav, dual* f) {
    const cparams* P = &h->P;
    int nf = h->nf;
    dual pp = mkvar(pl->p, 0), pm = mkvar(mi->p, nf);
    dual Tp = mkvar(pl->T, 1), Tm = mkvar(mi->T, nf + 1);
    dual dp = dscale(ih, dsub(pp, pm));
    dual dT = dscale(ih, dsub(Tp, Tm));
    double aK = area * Kf;
    dual rop = mk2(pl->ro, 0), rom = mk2(mi->ro, nf);
    dual flo = dsub(dp, dscale(0.5 * grav, dadd(rop, rom)));
    int upo = flo.v > 0.0;
    dual lamo = upo ? mk3(pl->lo, 0, nf) : mk3(mi->lo, nf, nf);
    dual fo = dscale(aK, dmul(lamo, flo));
    if (nf == 2) {
        dual kTf = dc(harm(pl->kT[0], mi->kT[0]));
        f[0] = fo;
        f[1] = dadd(dscale(P->co, dmul(upo ? Tp : Tm, fo)), dscale(area, dmul(kTf, dT)));
    } else {
        dual rwp = mk2(pl->rw, 0), rwm = mk2(mi->rw, nf);
        dual flw = dsub(dp, dscale(0.5 * grav, dadd(rwp, rwm)));
        int upw = flw.v > 0.0;
        dual lamw = upw ? mk3(pl->lw, 0, nf) : mk3(mi->lw, nf, nf);
        dual fw = dscale(aK, dmul(lamw, flw));
        dual few = dscale(P->cw, dmul(upw ? Tp : Tm, fw));
        dual feo = dscale(P->co, dmul(upo ? Tp : Tm, fo));
        dual kp = dc(pl->kT[0]), km = dc(mi->kT[0]);
        kp.d[2] = pl->kT[1];
        km.d[nf + 2] = mi->kT[1];
        dual ksum = dscale(0.5, dadd(kp, km));
        dual kTf = ksum.v > 0.0 ? ddiv(dmul(kp, km), ksum) : dc(0.0);
        f[0] = dscale(P->Wp, dadd(dscale(P->cw, fw), dscale(P->co, fo)));
        f[1] = dadd(dadd(few, feo), dscale(area, dmul(kTf, dT)));
        f[2] = dscale(P->Wo, fo);
    }
}

static const int OPP[NSMAX] = {0, 2, 1, 4, 3, 6, 5};

/* neighbour of (i,j,k) through slot s on an (nx,ny,nz) grid, -1 if outside */
static inline long nbr(int nx, int ny, int nz, int i, int j, int k, int s) {
    switch (s) {
        case 0: return i + (long)nx * (j + (long)ny * k);
        case 1: return i > 0 ? (i - 1) + (long)nx * (j + (long)ny * k) : -1;
        case 2: return i < nx - 1 ? (i + 1) + (long)nx * (j + (long)ny * k) : -1;
        case 3: return j > 0 ? i + (long)nx * (j - 1 + (long)ny * k) : -1;
        case 4: return j < ny - 1 ? i + (long)nx * (j + 1 + (long)ny * k) : -1;
        case 5: return k > 0 ? i + (long)nx * (j + (long)ny * (k - 1)) : -1;
        default: return k < nz - 1 ? i + (long)nx * (j + (long)ny * (k + 1)) : -1;
    }
}

static double peaceman_wi(double Kx, double Ky) { /* wellcase.py:180-192 */
    const double hh = 5.0, rw = 0.1, Dx = 5.0, Dy = 5.0;
    double ro = 0.28 * sqrt(sqrt(Ky / Kx) * Dx * Dx + sqrt(Kx / Ky) * Dy * Dy) / (pow(Ky / Kx, 0.25) + pow(Kx / Ky, 0.25));
    return 2.0 * 3.141592653589793 * hh * sqrt(Kx * Ky) / log(ro / rw);
}

/* rate and its partials w.r.t. (p,T,S) given 1/mu and its partials; wellcase.py:191-199 */
static void well_rate(const tpb_source* s, double wi, const double* imu /*v,p,T,S*/, double p, double* q /*v,p,T,S*/) {
    q[0] = s->max_rate;
    q[1] = q[2] = q[3] = 0.0;
    if (s->const_rate) return;
    double d = s->bhp - p;
    int active = s->max_rate < 0.0 ? !(d >= 0.0) : !(d <= 0.0);
    double dd = active ? d : 0.0, dd_p = active ? -1.0 : 0.0;
    double rate = wi * imu[0] * dd;
    if (fabs(rate) - fabs(s->max_rate) >= 0.0) return;
    q[0] = rate;
    q[1] = wi * (imu[1] * dd + imu[0] * dd_p);
    q[2] = wi * imu[2] * dd;
    q[3] = wi * imu[3] * dd;
}

/* source contributions of one entry at its cell: acc[r] (v, d/dp, d/dT, d/dS) to be ADDED to F / J diag */
static void source_terms(const tpc_handle_s* h, const tpb_source* s, const double* u, double acc[MAXF][4]) {
    const cparams* P = &h->P;
    long n = h->g.n, c = s->cell;
    int nf = h->nf;
    double p = u[c], T = u[n + c], w = s->weight;
    for (int r = 0; r < MAXF; r++)
        for (int k = 0; k < 4; k++) acc[r][k] = 0.0;
    if (s->kind == TPB_HEATER) { /* F -= delta*U*(T_inj - T) */
        acc[1][0] = -w * P->U * (P->T_inj - T);
        acc[1][2] = w * P->U;
        return;
    }
    double wi = s->const_rate ? 0.0 : peaceman_wi(h->fld[TPB_KX][c], h->fld[TPB_KY][c]);
    double ro[3], imo, imo_T;
    oil_rho_d(P, p, T, &ro[0], &ro[1], &ro[2]);
    oil_imu_d(P, T, &imo, &imo_T);
    if (nf == 2) {
        double imu[4] = {imo, 0.0, imo_T, 0.0}, q[4];
        well_rate(s, wi, imu, p, q);
        if (s->kind == TPB_PROD) { /* singlephase.py:151-156 */
            /* m = w rho q ; rows: -m, -c_v m T */
            double m = w * ro[0] * q[0], m_p = w * (ro[1] * q[0] + ro[0] * q[1]), m_T = w * (ro[2] * q[0] + ro[0] * q[2]);
            acc[0][0] = -m;
            acc[0][1] = -m_p;
            acc[0][2] = -m_T;
            acc[1][0] = -P->co * m * T;
            acc[1][1] = -P->co * m_p * T;
            acc[1][2] = -P->co * (m_T * T + m);
        } else { /* :157-162, rho at T_inj */
            double ri[3];
            oil_rho_d(P, p, P->T_inj, &ri[0], &ri[1], &ri[2]);
            double m = w * ri[0] * q[0], m_p = w * (ri[1] * q[0] + ri[0] * q[1]), m_T = w * ri[0] * q[2];
            acc[0][0] = -m;
            acc[0][1] = -m_p;
            acc[0][2] = -m_T;
            acc[1][0] = -P->co * P->T_inj * m;
            acc[1][1] = -P->co * P->T_inj * m_p;
            acc[1][2] = -P->co * P->T_inj * m_T;
        }
        return;
    }
    double S = u[2 * n + c], rw[3], imw, imw_T;
    water_rho_d(p, T, &rw[0], &rw[1], &rw[2]);
    water_imu_d(T, &imw, &imw_T);
    if (s->kind == TPB_PROD) {
        /* mu = 1/(S/mu_o + (1-S)/mu_w); q_w = (1-S)/mu_w mu q; q_o = S/mu_o mu q   wellcase.py:204-235 */
        dual Sd = dc(S), pd = dc(p), Td = dc(T);
        (void)pd;
        Sd.d[2] = 1.0;
        Td.d[1] = 1.0;
        dual imod = dc(imo), imwd = dc(imw);
        imod.d[1] = imo_T;
        imwd.d[1] = imw_T;
        dual one_m_S = dsub(dc(1.0), Sd);
        dual mob_o = dmul(Sd, imod), mob_w = dmul(one_m_S, imwd);
        dual imu_d = dadd(mob_o, mob_w);
        double imu[4] = {imu_d.v, imu_d.d[0], imu_d.d[1], imu_d.d[2]}, q[4];
        well_rate(s, wi, imu, p, q);
        dual qd = dc(q[0]);
        qd.d[0] = q[1];
        qd.d[1] = q[2];
        qd.d[2] = q[3];
        dual mu = ddiv(dc(1.0), imu_d);
        dual qw = dmul(dmul(mob_w, mu), qd), qo = dmul(dmul(mob_o, mu), qd);
        dual rod = dc(ro[0]), rwd = dc(rw[0]);
        rod.d[0] = ro[1];
        rod.d[1] = ro[2];
        rwd.d[0] = rw[1];
        rwd.d[1] = rw[2];
        dual mw = dmul(rwd, qw), mo = dmul(rod, qo);
        dual hsum = dadd(dscale(P->cw, mw), dscale(P->co, mo));
        dual a0 = dscale(-P->Wp * w, hsum);           /* twophase.py:396 */
        dual a2 = dscale(-P->Wo * w, mo);
        dual a1 = dscale(-w, dmul(hsum, Td));         /* :399 */
        dual* a[3] = {&a0, &a1, &a2};
        for (int r = 0; r < 3; r++) {
            acc[r][0] = a[r]->v;
            acc[r][1] = a[r]->d[0];
            acc[r][2] = a[r]->d[1];
            acc[r][3] = a[r]->d[2];
        }
    } else { /* water injector, twophase.py:400-408 */
        double imu[4] = {imw, 0.0, imw_T, 0.0}, q[4], ri[3];
        well_rate(s, wi, imu, p, q);
        water_rho_d(p, P->T_inj, &ri[0], &ri[1], &ri[2]);
        double m = w * ri[0] * q[0], m_p = w * (ri[1] * q[0] + ri[0] * q[1]), m_T = w * ri[0] * q[2];
        acc[0][0] = -P->Wp * P->cw * m;
        acc[0][1] = -P->Wp * P->cw * m_p;
        acc[0][2] = -P->Wp * P->cw * m_T;
        acc[1][0] = -P->cw * P->T_inj * m;
        acc[1][1] = -P->cw * P->T_inj * m_p;
        acc[1][2] = -P->cw * P->T_inj * m_T;
    }
}

/* --------------------------------------------------------------------------------------- */
/* assembly                                                                                 */
/* --------------------------------------------------------------------------------------- */
static void assemble(tpc_handle_s* h, const double* u, const double* uo, double dt, double* F, double* J) {
    const cgrid* g = &h->g;
    const cparams* P = &h->P;
    const int nf = h->nf, ns = g->ns, nx = g->nx, ny = g->ny, nz = g->nz;
    const long n = g->n;
    cprops* cp = (cprops*)malloc(sizeof(cprops) * n);
#pragma omp parallel for schedule(static)
    for (long c = 0; c < n; c++) cell_props(h, u, c, &cp[c]);
    const double* Kax[3] = {h->fld[TPB_KX], h->fld[TPB_KY], g->dim == 3 ? h->fld[TPB_KZ] : h->fld[TPB_KY]};
    const double w = g->vol / dt;
#pragma omp parallel for schedule(static)
    for (long c = 0; c < n; c++) {
        int i = (int)(c % nx), j = (int)((c / nx) % ny), k = (int)(c / ((long)nx * ny));
        const cprops* me = &cp[c];
        double R[MAXF] = {0, 0, 0}, D[MAXF][MAXF] = {{0}};
        double phi = h->fld[TPB_PHI][c];
        double po = uo[c], To = uo[n + c];
        double ro_o, t1, t2;
        oil_rho_d(P, po, To, &ro_o, &t1, &t2);
        if (nf == 2) { /* singlephase.py:120,123 */
            R[0] = w * phi * (me->ro[0] - ro_o);
            D[0][0] = w * phi * me->ro[1];
            D[0][1] = w * phi * me->ro[2];
            double rk = w * (1.0 - phi) * P->rho_r * P->cr;
            R[1] = w * phi * P->co * (me->ro[0] * me->T - ro_o * To) + rk * (me->T - To);
            D[1][0] = w * phi * P->co * me->ro[1] * me->T;
            D[1][1] = w * phi * P->co * (me->ro[2] * me->T + me->ro[0]) + rk;
        } else { /* twophase.py:333,337,344,349 */
            double So = uo[2 * n + c], rw_o, S = me->S, T = me->T;
            water_rho_d(po, To, &rw_o, &t1, &t2);
            double aw = w * phi * (me->rw[0] * (1.0 - S) - rw_o * (1.0 - So));
            double aw_p = w * phi * me->rw[1] * (1.0 - S), aw_T = w * phi * me->rw[2] * (1.0 - S), aw_S = -w * phi * me->rw[0];
            double ao = w * phi * (me->ro[0] * S - ro_o * So);
            double ao_p = w * phi * me->ro[1] * S, ao_T = w * phi * me->ro[2] * S, ao_S = w * phi * me->ro[0];
            double rk = w * (1.0 - phi) * P->rho_r * P->cr;
            R[0] = P->Wp * (P->cw * aw + P->co * ao);
            D[0][0] = P->Wp * (P->cw * aw_p + P->co * ao_p);
            D[0][1] = P->Wp * (P->cw * aw_T + P->co * ao_T);
            D[0][2] = P->Wp * (P->cw * aw_S + P->co * ao_S);
            R[2] = P->Wo * ao;
            D[2][0] = P->Wo * ao_p;
            D[2][1] = P->Wo * ao_T;
            D[2][2] = P->Wo * ao_S;
            R[1] = w * phi * P->cw * (me->rw[0] * (1.0 - S) * T - rw_o * (1.0 - So) * To) +
                   w * phi * P->co * (me->ro[0] * S * T - ro_o * So * To) + rk * (T - To);
            D[1][0] = P->cw * aw_p * T + P->co * ao_p * T;
            D[1][1] = P->cw * (aw_T * T + w * phi * me->rw[0] * (1.0 - S)) + P->co * (ao_T * T + w * phi * me->ro[0] * S) + rk;
            D[1][2] = P->cw * aw_S * T + P->co * ao_S * T;
        }
        for (int s = 1; s < ns; s++) {
            long nb = nbr(nx, ny, nz, i, j, k, s);
            double O[MAXF][MAXF] = {{0}};
            if (nb >= 0) {
                int axis = (s - 1) >> 1, hi = (s - 1) & 1;
                double Kf = harm(Kax[axis][c], Kax[axis][nb]);
                double grav = axis == 2 ? P->g : 0.0;
                dual f[MAXF];
                if (hi) {
                    face_flux(h, me, &cp[nb], Kf, g->area[axis], 1.0 / g->h[axis], grav, f);
                    for (int r = 0; r < nf; r++) {
                        R[r] += f[r].v;
                        for (int q = 0; q < nf; q++) {
                            D[r][q] += f[r].d[q];
                            O[r][q] = f[r].d[nf + q];
                        }
                    }
                } else {
                    face_flux(h, &cp[nb], me, Kf, g->area[axis], 1.0 / g->h[axis], grav, f);
                    for (int r = 0; r < nf; r++) {
                        R[r] -= f[r].v;
                        for (int q = 0; q < nf; q++) {
                            D[r][q] -= f[r].d[nf + q];
                            O[r][q] = -f[r].d[q];
                        }
                    }
                }
            }
            if (J)
                for (int r = 0; r < nf; r++)
                    for (int q = 0; q < nf; q++) J[((long)(s * nf + r) * nf + q) * n + c] = O[r][q];
        }
        for (int r = 0; r < nf; r++) F[(long)r * n + c] = R[r];
        if (J)
            for (int r = 0; r < nf; r++)
                for (int q = 0; q < nf; q++) J[((long)r * nf + q) * n + c] = D[r][q];
    }
    free(cp);
    for (int e = 0; e < h->nsrc; e++) {
        double acc[MAXF][4];
        long c = h->src[e].cell;
        source_terms(h, &h->src[e], u, acc);
        for (int r = 0; r < nf; r++) {
            F[(long)r * n + c] += acc[r][0];
            if (J)
                for (int q = 0; q < nf; q++) J[((long)r * nf + q) * n + c] += acc[r][1 + q];
        }
    }
}

/* --------------------------------------------------------------------------------------- */
/* block-stencil SpMV (PETSc MatMult)                                                       */
/* --------------------------------------------------------------------------------------- */
static void spmv(const tpc_handle_s* h, const double* J, const double* x, double* y) {
    const cgrid* g = &h->g;
    const int nf = h->nf, ns = g->ns, nx = g->nx, ny = g->ny, nz = g->nz;
    const long n = g->n;
#pragma omp parallel for schedule(static)
    for (long c = 0; c < n; c++) {
        int i = (int)(c % nx), j = (int)((c / nx) % ny), k = (int)(c / ((long)nx * ny));
        double acc[MAXF] = {0, 0, 0};
        for (int s = 0; s < ns; s++) {
            long nb = nbr(nx, ny, nz, i, j, k, s);
            if (nb < 0) continue;
            for (int r = 0; r < nf; r++)
                for (int q = 0; q < nf; q++) acc[r] += J[((long)(s * nf + r) * nf + q) * n + c] * x[(long)q * n + nb];
        }
        for (int r = 0; r < nf; r++) y[(long)r * n + c] = acc[r];
    }
}

/* --------------------------------------------------------------------------------------- */
/* scalar-stencil multigrid (role of BoomerAMG)                                             */
/* --------------------------------------------------------------------------------------- */
static void mg_free(mghier* m) {
    for (int l = 0; l < m->nlev; l++) {
        if (l > 0) free(m->lev[l].a);
        free(m->lev[l].x);
        free(m->lev[l].b);
        free(m->lev[l].r);
        free(m->lev[l].xt);
        free(m->lev[l].xs0);
        free(m->lev[l].xs1);
        m->lev[l].xt = m->lev[l].xs0 = m->lev[l].xs1 = NULL;
        free(m->lev[l].lid);
        free(m->lev[l].llf);
        free(m->lev[l].lcp);
        m->lev[l].lid = m->lev[l].llf = m->lev[l].lcp = NULL;
    }
    m->nlev = 0;
}

/* Galerkin coarse operator for piecewise-constant aggregation: stays a 5|7-point stencil */
static void mg_coarsen_op(const mglevel* f, mglevel* c, int ns, double sc) {
    const int cx = f->cx, cy = f->cy, cz = f->cz;
    memset(c->a, 0, sizeof(double) * ns * c->n);
#pragma omp parallel for schedule(static)
    for (long C = 0; C < c->n; C++) {
        int I = (int)(C % c->nx), Jc = (int)((C / c->nx) % c->ny), Kc = (int)(C / ((long)c->nx * c->ny));
        double acc[NSMAX] = {0};
        for (int dk = 0; dk < cz; dk++)
            for (int dj = 0; dj < cy; dj++)
                for (int di = 0; di < cx; di++) {
                    int i = I * cx + di, j = Jc * cy + dj, k = Kc * cz + dk;
                    if (i >= f->nx || j >= f->ny || k >= f->nz) continue;
                    long fc = i + (long)f->nx * (j + (long)f->ny * k);
                    acc[0] += f->a[fc];
                    for (int s = 1; s < ns; s++) {
                        int axis = (s - 1) >> 1, hi = (s - 1) & 1;
                        int pos = axis == 0 ? i : (axis == 1 ? j : k);
                        int cf = axis == 0 ? cx : (axis == 1 ? cy : cz);
                        int npos = pos + (hi ? 1 : -1);
                        int same = (npos >= 0) && (npos / cf == pos / cf);
                        acc[same ? 0 : s] += f->a[(long)s * f->n + fc];
                    }
                }
        /* tpb_solver_opts.mg_coarse_scale: couplings along a coarsened axis are scaled, row sums kept */
        if (sc != 1.0)
            for (int s = 1; s < ns; s++) {
                int axis = (s - 1) >> 1;
                int cf = axis == 0 ? cx : (axis == 1 ? cy : cz);
                if (cf == 2) {
                    double nv = sc * acc[s];
                    acc[0] += acc[s] - nv;
                    acc[s] = nv;
                }
            }
        for (int s = 0; s < ns; s++) c->a[(long)s * c->n + C] = acc[s];
    }
}

/* LU factors of every column's tridiagonal (diag, z-, z+): id = 1/pivot, lf = lower * id, cp = upper * id; the
 * couplings through the bottom and the top of the box (k = 0 lower, k = nz-1 upper: slab faces on multi-rank runs)
 * are not part of the column.  A zero pivot gives id = 0 (the row's x stays 0, as the point smoother does). */
static void mg_line_factors(mglevel* L) {
    const int nz = L->nz;
    const long n = L->n, np = (long)L->nx * L->ny;
    L->lid = (double*)malloc(sizeof(double) * n);
    L->llf = (double*)malloc(sizeof(double) * n);
    L->lcp = (double*)malloc(sizeof(double) * n);
#pragma omp parallel for schedule(static)
    for (long q = 0; q < np; q++) {
        double cprev = 0.0;
        for (int k = 0; k < nz; k++) {
            long c = q + np * k;
            double lo = k > 0 ? L->a[5 * n + c] : 0.0, up = k < nz - 1 ? L->a[6 * n + c] : 0.0;
            double den = L->a[c] - lo * cprev;
            double id = den != 0.0 ? 1.0 / den : 0.0;
            L->lid[c] = id;
            L->llf[c] = lo * id;
            cprev = up * id;
            L->lcp[c] = cprev;
        }
    }
}

static void mg_setup(const tpc_handle_s* h, mghier* m, double* a0) {
    const tpb_solver_opts* o = &h->opts;
    const int ns = h->g.ns;
    mg_free(m);
    /* row repair (csrc/tpb_pc.cu row_repair_kernel): a diagonal far below the sum of the row's couplings is raised
     * to that sum - rows of cells whose Newton iterate left the physical range (S_o outside [0,1]) would
     * otherwise make Gauss-Seidel diverge */
    {
        const long n0 = h->g.n;
#pragma omp parallel for schedule(static)
        for (long c = 0; c < n0; c++) {
            double sum = 0.0;
            for (int s = 1; s < ns; s++) sum += fabs(a0[(long)s * n0 + c]);
            if (a0[c] < 0.8 * sum) a0[c] = sum;
        }
    }
    mglevel* L = &m->lev[0];
    L->nx = h->g.nx;
    L->ny = h->g.ny;
    L->nz = h->g.nz;
    L->n = h->g.n;
    L->a = a0;
    m->last_sweeps = 0;
    /* zebra z-line smoothing (tpb_solver_opts.mg_smoother) on every level of a 3-D hierarchy; z is then not coarsened */
    const int line = (o->mg_smoother == TPB_MG_ZLINE && ns == 7 && h->g.nz > 1);
    int l = 0;
    for (;;) {
        L = &m->lev[l];
        L->x = (double*)calloc(L->n, sizeof(double));
        L->b = (double*)calloc(L->n, sizeof(double));
        L->r = (double*)calloc(L->n, sizeof(double));
        L->cx = L->cy = L->cz = 1;
        L->line = line;
        L->xt = line ? (double*)calloc(L->n, sizeof(double)) : NULL;
        L->xs0 = line ? (double*)calloc(L->n, sizeof(double)) : NULL;
        L->xs1 = line ? (double*)calloc(L->n, sizeof(double)) : NULL;
        L->lid = L->llf = L->lcp = NULL;
        if (line) mg_line_factors(L);
        if (line && L->nx == 1 && L->ny == 1) { /* a single column: the line solve is exact */
            m->last_sweeps = 1;
            break;
        }
        if (L->n <= o->mg_min_cells || L->n <= 1 || l == MAXLEV - 1) break;
        if (o->mg_dd_stop > 0.0) {
            /* strongest row of the level: max over cells of sum|off-diag| / |diag| */
            double rho = 0.0;
#pragma omp parallel for reduction(max : rho) schedule(static)
            for (long c = 0; c < L->n; c++) {
                double sum = 0.0;
                for (int s = 1; s < ns; s++) sum += fabs(L->a[(long)s * L->n + c]);
                const double d = fabs(L->a[c]);
                const double r = d > 0.0 ? sum / d : (sum > 0.0 ? 1e300 : 0.0);
                if (r > rho) rho = r;
            }
            if (rho <= o->mg_dd_stop) {
                m->last_sweeps = dd_sweeps(rho, o->mg_coarse_sweeps);
                break;
            }
        }
        /* mean coupling per axis decides which axes are coarsened (semi-coarsening) */
        double m_ax[3] = {0, 0, 0};
        int dims[3] = {L->nx, L->ny, L->nz};
        for (int ax = 0; ax < (ns - 1) / 2; ax++) {
            double sum = 0.0;
            const double* a1 = L->a + (long)(2 * ax + 1) * L->n;
            const double* a2 = L->a + (long)(2 * ax + 2) * L->n;
#pragma omp parallel for reduction(+ : sum) schedule(static)
            for (long c = 0; c < L->n; c++) sum += fabs(a1[c]) + fabs(a2[c]);
            m_ax[ax] = sum;
        }
        if (line) dims[2] = 1; /* z-line smoothing: z is never coarsened */
        double mmax = 0.0;
        for (int ax = 0; ax < 3; ax++)
            if (dims[ax] > 1 && m_ax[ax] > mmax) mmax = m_ax[ax];
        int cf[3] = {1, 1, 1}, any = 0;
        for (int ax = 0; ax < 3; ax++)
            if (dims[ax] > 1 && (m_ax[ax] >= o->mg_semi_theta * mmax || L->n <= o->mg_full_below)) {
                cf[ax] = 2;
                any = 1;
            }
        if (!any) { /* all couplings zero: coarsen every axis that can be */
            for (int ax = 0; ax < 3; ax++)
                if (dims[ax] > 1) cf[ax] = 2, any = 1;
        }
        if (!any) break;
        L->cx = cf[0];
        L->cy = cf[1];
        L->cz = cf[2];
        mglevel* Cc = &m->lev[l + 1];
        Cc->nx = (L->nx + cf[0] - 1) / cf[0];
        Cc->ny = (L->ny + cf[1] - 1) / cf[1];
        Cc->nz = (L->nz + cf[2] - 1) / cf[2];
        Cc->n = (long)Cc->nx * Cc->ny * Cc->nz;
        Cc->a = (double*)malloc(sizeof(double) * ns * Cc->n);
        mg_coarsen_op(L, Cc, ns, o->mg_coarse_scale > 0.0 ? o->mg_coarse_scale : 1.0);
        l++;
    }
    m->nlev = l + 1;
}

/* one red-black Gauss-Seidel sweep; colour = (i+j+k)&1, red (0) first.  zero_guess: x is
 * treated as 0 on entry (the red pass then needs no neighbour reads) */
static void mg_rbgs(const mglevel* L, int ns, const double* b, double* x, int zero_guess) {
    const int nx = L->nx, ny = L->ny, nz = L->nz;
    const long n = L->n;
    for (int col = 0; col < 2; col++) {
#pragma omp parallel for schedule(static)
        for (long c = 0; c < n; c++) {
            int i = (int)(c % nx), j = (int)((c / nx) % ny), k = (int)(c / ((long)nx * ny));
            if (((i + j + k) & 1) != col) continue;
            double acc = b[c];
            if (!(zero_guess && col == 0)) {
                for (int s = 1; s < ns; s++) {
                    long nb = nbr(nx, ny, nz, i, j, k, s);
                    if (nb >= 0) acc -= L->a[(long)s * n + c] * x[nb];
                }
            }
            double d = L->a[c];
            x[c] = d != 0.0 ? acc / d : 0.0;
        }
    }
}

/* Tile shape of the line smoother (csrc/tpb_pc.cu line_tile_shape): as many columns as give at most 2048 cells */
static void line_tile_shape(int nz, int* tx, int* ty) {
    static const int menu[3][2] = {{8, 3}, {4, 2}, {1, 1}};
    const int cols_max = nz > 0 ? 2048 / nz : 1;
    for (int q = 0; q < 3; q++)
        if (menu[q][0] * menu[q][1] <= cols_max || q == 2) {
            *tx = menu[q][0];
            *ty = menu[q][1];
            return;
        }
}

/* Hybrid zebra z-line Gauss-Seidel (csrc/tpb_pc.cu line_smooth_kernel): the xy-plane is cut into tiles of tx x ty
 * columns aligned at multiples of the tile shape; every tile does `nsweeps` zebra sweeps on its own (columns coloured
 * by the global (i+j)&1, colour 0 first, every column solved exactly with the factors of mg_line_factors:
 *   d_k = rhs_k * id_k - lf_k * d_{k-1} ;  x_k = d_k - cp_k * x_{k+1},
 *   rhs_k = b_k - sum over the four lateral neighbours) while the columns outside the tile keep their values of
 * entry - Gauss-Seidel inside a tile, block Jacobi between tiles, what hypre's hybrid smoother is between processes.
 * On the GPU a tile lives in one thread block's shared memory for all its sweeps.  Out of place: xin (NULL = zero
 * guess) is read, xout written. */
static void mg_line_smooth1(const mglevel* L, const double* b, const double* xin, double* xout, int nsweeps) {
    const int nx = L->nx, ny = L->ny, nz = L->nz;
    const long n = L->n, np = (long)nx * ny;
    int tx, ty;
    line_tile_shape(nz, &tx, &ty);
    const int ntx = (nx + tx - 1) / tx, nty = (ny + ty - 1) / ty;
    const int hx = tx + 2, hy = ty + 2;
#pragma omp parallel
    {
        double* xl = (double*)malloc(sizeof(double) * hx * hy * nz); /* tile + halo ring, [jj][ii][k] */
#pragma omp for schedule(static)
        for (int tile = 0; tile < ntx * nty; tile++) {
            const int i0 = (tile % ntx) * tx, j0 = (tile / ntx) * ty;
            for (int jj = 0; jj < hy; jj++)
                for (int ii = 0; ii < hx; ii++) {
                    const int i = i0 + ii - 1, j = j0 + jj - 1;
                    double* col = xl + ((long)jj * hx + ii) * nz;
                    const int in = xin && i >= 0 && i < nx && j >= 0 && j < ny;
                    for (int k = 0; k < nz; k++) col[k] = in ? xin[i + (long)nx * j + np * k] : 0.0;
                }
            for (int sw = 0; sw < nsweeps; sw++)
                for (int colr = 0; colr < 2; colr++)
                    for (int jj = 1; jj <= ty; jj++)
                        for (int ii = 1; ii <= tx; ii++) {
                            const int i = i0 + ii - 1, j = j0 + jj - 1;
                            if (i >= nx || j >= ny || ((i + j) & 1) != colr) continue;
                            double* me = xl + ((long)jj * hx + ii) * nz;
                            const double* xm = me - nz, *xp = me + nz, *ym = me - (long)hx * nz, *yp = me + (long)hx * nz;
                            const long q = i + (long)nx * j;
                            double dprev = 0.0;
                            for (int k = 0; k < nz; k++) {
                                const long c = q + np * k;
                                double rhs = b[c];
                                if (i > 0) rhs -= L->a[1 * n + c] * xm[k];
                                if (i < nx - 1) rhs -= L->a[2 * n + c] * xp[k];
                                if (j > 0) rhs -= L->a[3 * n + c] * ym[k];
                                if (j < ny - 1) rhs -= L->a[4 * n + c] * yp[k];
                                dprev = rhs * L->lid[c] - L->llf[c] * dprev;
                                me[k] = dprev;
                            }
                            double xn = 0.0;
                            for (int k = nz - 1; k >= 0; k--) {
                                xn = me[k] - L->lcp[q + np * k] * xn;
                                me[k] = xn;
                            }
                        }
            for (int jj = 1; jj <= ty; jj++)
                for (int ii = 1; ii <= tx; ii++) {
                    const int i = i0 + ii - 1, j = j0 + jj - 1;
                    if (i >= nx || j >= ny) continue;
                    const double* me = xl + ((long)jj * hx + ii) * nz;
                    for (int k = 0; k < nz; k++) xout[i + (long)nx * j + np * k] = me[k];
                }
        }
        free(xl);
    }
}

/* `nsweeps` sweeps in groups of `group` (tpb_solver_opts.mg_tile_sweeps; <= 0: all in one group): a tile exchanges
 * its rim with the neighbouring tiles between groups.  xin (NULL = zero guess) is only read, the last group writes
 * xout, the groups before it alternate between xout's partner buffers (o0, o1 are two work vectors distinct from xin
 * and xout). */
static void mg_line_smooth(const mglevel* L, const double* b, const double* xin, double* xout, int nsweeps, int group,
                           double* o0, double* o1) {
    if (group <= 0 || group > nsweeps) group = nsweeps;
    const int ncalls = (nsweeps + group - 1) / group;
    const double* in = xin;
    int left = nsweeps;
    for (int j = 0; j < ncalls; j++) {
        double* out = (j == ncalls - 1) ? xout : (((ncalls - 1 - j) & 1) ? o0 : o1);
        const int sw = left < group ? left : group;
        mg_line_smooth1(L, b, in, out, sw);
        left -= sw;
        in = out;
    }
}

static void mg_residual(const mglevel* L, int ns, const double* b, const double* x, double* r) {
    const int nx = L->nx, ny = L->ny, nz = L->nz;
    const long n = L->n;
#pragma omp parallel for schedule(static)
    for (long c = 0; c < n; c++) {
        int i = (int)(c % nx), j = (int)((c / nx) % ny), k = (int)(c / ((long)nx * ny));
        double acc = b[c] - L->a[c] * x[c];
        for (int s = 1; s < ns; s++) {
            long nb = nbr(nx, ny, nz, i, j, k, s);
            if (nb >= 0) acc -= L->a[(long)s * n + c] * x[nb];
        }
        r[c] = acc;
    }
}

static void mg_restrict(const mglevel* f, const mglevel* c, const double* r, double* bc) {
#pragma omp parallel for schedule(static)
    for (long C = 0; C < c->n; C++) {
        int I = (int)(C % c->nx), Jc = (int)((C / c->nx) % c->ny), Kc = (int)(C / ((long)c->nx * c->ny));
        double acc = 0.0;
        for (int dk = 0; dk < f->cz; dk++)
            for (int dj = 0; dj < f->cy; dj++)
                for (int di = 0; di < f->cx; di++) {
                    int i = I * f->cx + di, j = Jc * f->cy + dj, k = Kc * f->cz + dk;
                    if (i >= f->nx || j >= f->ny || k >= f->nz) continue;
                    /* point smoother: the pre-smoothing sweep ended on colour 1, those rows were just solved, their
                     * residual is zero and is not summed (csrc/tpb_pc.cu restrict_cell).  Line-smoothed levels sum
                     * every row: the hybrid smoother's tiles freeze their surroundings, so no row is exactly solved */
                    if (!f->line && ((i + j + k) & 1)) continue;
                    acc += r[i + (long)f->nx * (j + (long)f->ny * k)];
                }
        bc[C] = acc;
    }
}

static void mg_prolong_add(const mglevel* f, const mglevel* c, const double* xc, double* x, double omega) {
#pragma omp parallel for schedule(static)
    for (long fc = 0; fc < f->n; fc++) {
        int i = (int)(fc % f->nx), j = (int)((fc / f->nx) % f->ny), k = (int)(fc / ((long)f->nx * f->ny));
        long C = (i / f->cx) + (long)c->nx * ((j / f->cy) + (long)c->ny * (k / f->cz));
        x[fc] += omega * xc[C];
    }
}

static void mg_vcycle_level(const tpc_handle_s* h, mghier* m, int l) {
    const tpb_solver_opts* o = &h->opts;
    const int ns = h->g.ns;
    mglevel* L = &m->lev[l];
    if (l == m->nlev - 1) {
        int sweeps = m->last_sweeps > 0 ? m->last_sweeps : (o->mg_coarse_sweeps > 0 ? o->mg_coarse_sweeps : 1);
        if (L->line)
            mg_line_smooth(L, L->b, NULL, L->x, sweeps, o->mg_tile_sweeps, L->xs0, L->xs1);
        else
            for (int s = 0; s < sweeps; s++) mg_rbgs(L, ns, L->b, L->x, s == 0);
        return;
    }
    int pre = o->mg_pre > 0 ? o->mg_pre : 1;
    const int post = o->mg_post;
    double* cur = L->x; /* the iterate after pre-smoothing */
    if (L->line) {
        cur = L->xt;
        mg_line_smooth(L, L->b, NULL, cur, pre, o->mg_tile_sweeps, L->xs0, L->xs1);
    } else
        for (int s = 0; s < pre; s++) mg_rbgs(L, ns, L->b, L->x, s == 0);
    mg_residual(L, ns, L->b, cur, L->r);
    mglevel* Cc = &m->lev[l + 1];
    mg_restrict(L, Cc, L->r, Cc->b);
    mg_vcycle_level(h, m, l + 1);
    mg_prolong_add(L, Cc, Cc->x, cur, o->mg_overcorrection);
    if (L->line) {
        if (post > 0)
            mg_line_smooth(L, L->b, cur, L->x, post, o->mg_tile_sweeps, L->xs0, L->xs1);
        else
            memcpy(L->x, cur, sizeof(double) * L->n);
    } else
        for (int s = 0; s < post; s++) mg_rbgs(L, ns, L->b, L->x, 0);
}

/* y = V(b): mg_cycles V-cycles from a zero initial guess */
static void mg_apply(const tpc_handle_s* h, mghier* m, const double* b, double* y) {
    mglevel* L = &m->lev[0];
    const long n = L->n;
    const int ns = h->g.ns;
    int cycles = h->opts.mg_cycles > 0 ? h->opts.mg_cycles : 1;
    memcpy(L->b, b, sizeof(double) * n);
    mg_vcycle_level(h, m, 0);
    memcpy(y, L->x, sizeof(double) * n);
    for (int cyc = 1; cyc < cycles; cyc++) {
        mg_residual(L, ns, b, y, L->b);
        mg_vcycle_level(h, m, 0);
#pragma omp parallel for schedule(static)
        for (long c = 0; c < n; c++) y[c] += L->x[c];
    }
}

/* --------------------------------------------------------------------------------------- */
/* ConvDiff temperature operator (preconditioners.py:63-108, 225-276)                       */
/* --------------------------------------------------------------------------------------- */
static void assemble_convdiff(const tpc_handle_s* h, const double* u, double dt, double* A) {
    const cgrid* g = &h->g;
    const cparams* P = &h->P;
    const int nf = h->nf, ns = g->ns, nx = g->nx, ny = g->ny, nz = g->nz;
    const long n = g->n;
    cprops* cp = (cprops*)malloc(sizeof(cprops) * n);
#pragma omp parallel for schedule(static)
    for (long c = 0; c < n; c++) cell_props(h, u, c, &cp[c]);
    const double* Kax[3] = {h->fld[TPB_KX], h->fld[TPB_KY], g->dim == 3 ? h->fld[TPB_KZ] : h->fld[TPB_KY]};
#pragma omp parallel for schedule(static)
    for (long c = 0; c < n; c++) {
        int i = (int)(c % nx), j = (int)((c / nx) % ny), k = (int)(c / ((long)nx * ny));
        const cprops* me = &cp[c];
        double phi = h->fld[TPB_PHI][c];
        double diag;
        if (nf == 2)
            diag = g->vol / dt * (phi * P->co * me->ro[0] + (1.0 - phi) * P->rho_r * P->cr);
        else
            diag = g->vol / dt * (phi * P->co * me->S * me->ro[0] + phi * P->cw * (1.0 - me->S) * me->rw[0] +
                                  (1.0 - phi) * P->rho_r * P->cr);
        for (int s = 1; s < ns; s++) {
            long nb = nbr(nx, ny, nz, i, j, k, s);
            double off = 0.0;
            if (nb >= 0) {
                int axis = (s - 1) >> 1, hi = (s - 1) & 1;
                const cprops* pl = hi ? me : &cp[nb];
                const cprops* mi = hi ? &cp[nb] : me;
                double Kf = harm(Kax[axis][c], Kax[axis][nb]);
                double grav = axis == 2 ? P->g : 0.0;
                double ih = 1.0 / g->h[axis], area = g->area[axis];
                double dp = ih * (pl->p - mi->p);
                double sgn = hi ? 1.0 : -1.0; /* row sign of jump(r) for this cell */
                /* oil (the only phase of the single-phase model) */
                double flo = dp - 0.5 * grav * (pl->ro[0] + mi->ro[0]);
                int upo = flo > 0.0;
                double co = area * Kf * P->co * (upo ? pl->lo[0] : mi->lo[0]) * flo;
                int up_is_me = hi ? upo : !upo;
                if (up_is_me) diag += sgn * co; else off += sgn * co;
                if (nf == 3) {
                    double flw = dp - 0.5 * grav * (pl->rw[0] + mi->rw[0]);
                    int upw = flw > 0.0;
                    double cw = area * Kf * P->cw * (upw ? pl->lw[0] : mi->lw[0]) * flw;
                    int upw_me = hi ? upw : !upw;
                    if (upw_me) diag += sgn * cw; else off += sgn * cw;
                }
                double d = area * harm(pl->kT[0], mi->kT[0]) * ih;
                diag += d;
                off -= d;
            }
            A[(long)s * n + c] = off;
        }
        A[c] = diag;
    }
    free(cp);
    for (int e = 0; e < h->nsrc; e++) {
        const tpb_source* s = &h->src[e];
        long c = s->cell;
        if (s->kind == TPB_HEATER) {
            A[c] += s->weight * P->U;
        } else if (s->kind == TPB_PROD) {
            /* a -= rho q c_v T : the frozen-coefficient energy sink of producers */
            double acc[MAXF][4];
            double T = u[n + c];
            source_terms(h, s, u, acc);
            /* acc[1][0] = -w (sum rho q c) T  => coefficient of T */
            if (T != 0.0) A[c] += acc[1][0] / T;
        }
    }
}

/* --------------------------------------------------------------------------------------- */
/* stage 1 set-up: block extraction + decoupling                                            */
/* --------------------------------------------------------------------------------------- */
static inline double Jat(const double* J, long n, int nf, int s, int r, int q, long c) {
    return J[((long)(s * nf + r) * nf + q) * n + c];
}

/* column sum of block (r,q) for column cell c: entries of every row that points at c */
static double colsum(const tpc_handle_s* h, const double* J, int r, int q, long c) {
    const cgrid* g = &h->g;
    int i = (int)(c % g->nx), j = (int)((c / g->nx) % g->ny), k = (int)(c / ((long)g->nx * g->ny));
    double sum = Jat(J, g->n, h->nf, 0, r, q, c);
    for (int s = 1; s < g->ns; s++) {
        long nb = nbr(g->nx, g->ny, g->nz, i, j, k, s);
        if (nb >= 0) sum += Jat(J, g->n, h->nf, OPP[s], r, q, nb);
    }
    return sum;
}

static void inv_small(int m, const double* A, double* Ai) { /* m = 1..3, row-major */
    if (m == 1) {
        Ai[0] = 1.0 / A[0];
    } else if (m == 2) {
        double det = A[0] * A[3] - A[1] * A[2], id = 1.0 / det;
        Ai[0] = A[3] * id;
        Ai[1] = -A[1] * id;
        Ai[2] = -A[2] * id;
        Ai[3] = A[0] * id;
    } else {
        double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
        double det = A[0] * c00 + A[1] * c01 + A[2] * c02, id = 1.0 / det;
        Ai[0] = c00 * id;
        Ai[1] = (A[2] * A[7] - A[1] * A[8]) * id;
        Ai[2] = (A[1] * A[5] - A[2] * A[4]) * id;
        Ai[3] = c01 * id;
        Ai[4] = (A[0] * A[8] - A[2] * A[6]) * id;
        Ai[5] = (A[2] * A[3] - A[0] * A[5]) * id;
        Ai[6] = c02 * id;
        Ai[7] = (A[1] * A[6] - A[0] * A[7]) * id;
        Ai[8] = (A[0] * A[4] - A[1] * A[3]) * id;
    }
}

static void stage1_setup(tpc_handle_s* h, const double* J, const double* u, double dt) {
    const cgrid* g = &h->g;
    const tpb_solver_opts* o = &h->opts;
    const int nf = h->nf, ns = g->ns, L = nf - 1;
    const long n = g->n;
    const int dec = o->decoup;
    if (o->stage1 == TPB_S1_NONE) return;
    for (int f = 0; f < MAXF; f++)
        if (!h->w[f]) h->w[f] = (double*)calloc(n, sizeof(double));
    if (!h->App) h->App = (double*)malloc(sizeof(double) * NSMAX * n);
    if (o->stage1 == TPB_S1_CPR) {
        /* r_p = x_p - sum_f w[f] x_f ; Atilde_pp = A_pp - sum_f w[f] A_fp */
#pragma omp parallel for schedule(static)
        for (long c = 0; c < n; c++) {
            double wf[MAXF] = {0, 0, 0};
            if (dec == TPB_DECOUP_QI) {
                wf[L] = Jat(J, n, nf, 0, 0, L, c) / Jat(J, n, nf, 0, L, L, c);          /* :785-808 */
            } else if (dec == TPB_DECOUP_TI) {
                wf[L] = colsum(h, J, 0, L, c) / colsum(h, J, L, L, c);                   /* :684-711 */
            } else if (dec == TPB_DECOUP_QI_TEMP || dec == TPB_DECOUP_TI_TEMP) {
                int ti = dec == TPB_DECOUP_TI_TEMP;                                     /* :714-783, 810-873 */
                double B[4], Bi[4], pT, pS;
                if (ti) {
                    B[0] = colsum(h, J, 1, 1, c); B[1] = colsum(h, J, 1, 2, c);
                    B[2] = colsum(h, J, 2, 1, c); B[3] = colsum(h, J, 2, 2, c);
                    pT = colsum(h, J, 0, 1, c); pS = colsum(h, J, 0, 2, c);
                } else {
                    B[0] = Jat(J, n, nf, 0, 1, 1, c); B[1] = Jat(J, n, nf, 0, 1, 2, c);
                    B[2] = Jat(J, n, nf, 0, 2, 1, c); B[3] = Jat(J, n, nf, 0, 2, 2, c);
                    pT = Jat(J, n, nf, 0, 0, 1, c); pS = Jat(J, n, nf, 0, 0, 2, c);
                }
                inv_small(2, B, Bi);
                wf[1] = pT * Bi[0] + pS * Bi[2];
                wf[2] = pT * Bi[1] + pS * Bi[3];
            }
            for (int f = 1; f < nf; f++) h->w[f][c] = wf[f];
            for (int s = 0; s < ns; s++) {
                double v = Jat(J, n, nf, s, 0, 0, c);
                for (int f = 1; f < nf; f++) v -= wf[f] * Jat(J, n, nf, s, f, 0, c);
                h->App[(long)s * n + c] = v;
            }
        }
        mg_setup(h, &h->mg_p, h->App);
        return;
    }
    /* CPTR (two-phase, secondary = S) and single-phase FIELDSPLIT (no secondary): 2x2 primary block */
    if (!h->A00) h->A00 = (double*)malloc(sizeof(double) * NSMAX * 4 * n);
    if (!h->AT) h->AT = (double*)malloc(sizeof(double) * NSMAX * n);
    const int has_s = (o->stage1 == TPB_S1_CPTR);
#pragma omp parallel for schedule(static)
    for (long c = 0; c < n; c++) {
        double wa[2] = {0, 0};
        if (has_s && dec == TPB_DECOUP_QI) {                                            /* :1505-1543 */
            double dss = Jat(J, n, nf, 0, 2, 2, c);
            wa[0] = Jat(J, n, nf, 0, 0, 2, c) / dss;
            wa[1] = Jat(J, n, nf, 0, 1, 2, c) / dss;
        } else if (has_s && dec == TPB_DECOUP_TI) {                                     /* :1445-1503 */
            double dss = colsum(h, J, 2, 2, c);
            wa[0] = colsum(h, J, 0, 2, c) / dss;
            wa[1] = colsum(h, J, 1, 2, c) / dss;
        }
        h->w[0][c] = wa[0];
        h->w[1][c] = wa[1];
        for (int s = 0; s < ns; s++)
            for (int a = 0; a < 2; a++)
                for (int b = 0; b < 2; b++) {
                    double v = Jat(J, n, nf, s, a, b, c);
                    if (has_s) v -= wa[a] * Jat(J, n, nf, s, 2, b, c);
                    h->A00[((long)(s * 2 + a) * 2 + b) * n + c] = v;
                    if (a == 0 && b == 0) h->App[(long)s * n + c] = v;
                    if (a == 1 && b == 1 && o->schur_pre != TPB_SCHUR_CONVDIFF) h->AT[(long)s * n + c] = v;
                }
    }
    if (o->schur_pre == TPB_SCHUR_CONVDIFF) assemble_convdiff(h, u, dt, h->AT);
    if (o->schur_pre == TPB_SCHUR_SELFP) {
        /* AT = A11 - A10 diag(A00)^-1 A01 collapsed onto the stencil (csrc/tpb_pc.cu selfp_kernel) */
#pragma omp parallel for schedule(static)
        for (long c = 0; c < n; c++) {
            int i = (int)(c % g->nx), j = (int)((c / g->nx) % g->ny), k = (int)(c / ((long)g->nx * g->ny));
            double sub[NSMAX] = {0, 0, 0, 0, 0, 0, 0};
            for (int s1 = 0; s1 < ns; s1++) {
                long nb = nbr(g->nx, g->ny, g->nz, i, j, k, s1);
                if (nb < 0) continue;
                const double d = h->A00[nb];
                const double f = d != 0.0 ? h->A00[((long)(s1 * 2 + 1) * 2 + 0) * n + c] / d : 0.0;
                int i2 = (int)(nb % g->nx), j2 = (int)((nb / g->nx) % g->ny), k2 = (int)(nb / ((long)g->nx * g->ny));
                for (int s2 = 0; s2 < ns; s2++) {
                    if (nbr(g->nx, g->ny, g->nz, i2, j2, k2, s2) < 0) continue;
                    const double v = f * h->A00[((long)(s2 * 2 + 0) * 2 + 1) * n + nb];
                    const int slot = s1 == 0 ? s2 : (s2 == 0 ? s1 : 0);
                    sub[slot] += v;
                }
            }
            for (int s = 0; s < ns; s++) h->AT[(long)s * n + c] -= sub[s];
        }
    }
    mg_setup(h, &h->mg_p, h->App);
    mg_setup(h, &h->mg_T, h->AT);
}

/* y_a = sum_s A00[s][a][b] x_b[nb] for one (a,b) coupling block */
static void a00_mult(const tpc_handle_s* h, int a, int b, const double* x, double* y) {
    const cgrid* g = &h->g;
    const int ns = g->ns, nx = g->nx, ny = g->ny, nz = g->nz;
    const long n = g->n;
#pragma omp parallel for schedule(static)
    for (long c = 0; c < n; c++) {
        int i = (int)(c % nx), j = (int)((c / nx) % ny), k = (int)(c / ((long)nx * ny));
        double acc = 0.0;
        for (int s = 0; s < ns; s++) {
            long nb = nbr(nx, ny, nz, i, j, k, s);
            if (nb >= 0) acc += h->A00[((long)(s * 2 + a) * 2 + b) * n + c] * x[nb];
        }
        y[c] = acc;
    }
}

static void stage1_apply(tpc_handle_s* h, const double* x, double* y) {
    const tpb_solver_opts* o = &h->opts;
    const int nf = h->nf;
    const long n = h->g.n;
    double* rp = h->t2;
    memset(y, 0, sizeof(double) * nf * n);
    if (o->stage1 == TPB_S1_CPR) {
#pragma omp parallel for schedule(static)
        for (long c = 0; c < n; c++) {
            double v = x[c];
            for (int f = 1; f < nf; f++) v -= h->w[f][c] * x[(long)f * n + c];
            rp[c] = v;
        }
        mg_apply(h, &h->mg_p, rp, y);
        return;
    }
    /* primary (p,T); PCFIELDSPLIT schur FULL (or additive for SCHUR_DIAG) */
    double* rT = h->t2 + n;
    double* tmp = h->t3;
    const int has_s = (o->stage1 == TPB_S1_CPTR);
#pragma omp parallel for schedule(static)
    for (long c = 0; c < n; c++) {
        double xs = has_s ? x[2 * n + c] : 0.0;
        rp[c] = x[c] - h->w[0][c] * xs;
        rT[c] = x[n + c] - h->w[1][c] * xs;
    }
    double* yp = y;
    double* yT = y + n;
    mg_apply(h, &h->mg_p, rp, yp);
    if (o->schur_pre == TPB_SCHUR_DIAG) {
        mg_apply(h, &h->mg_T, rT, yT);
        return;
    }
    a00_mult(h, 1, 0, yp, tmp);
#pragma omp parallel for schedule(static)
    for (long c = 0; c < n; c++) rT[c] -= tmp[c];
    mg_apply(h, &h->mg_T, rT, yT);
    a00_mult(h, 0, 1, yT, tmp);
#pragma omp parallel for schedule(static)
    for (long c = 0; c < n; c++) rp[c] -= tmp[c];
    mg_apply(h, &h->mg_p, rp, yp);
}

/* --------------------------------------------------------------------------------------- */
/* stage 2: block ILU(0) in red-black ordering (role of PETSc bjacobi + ilu(0))             */
/* --------------------------------------------------------------------------------------- */
static void stage2_setup(tpc_handle_s* h, const double* J) {
    const cgrid* g = &h->g;
    const int nf = h->nf, ns = g->ns, nx = g->nx, ny = g->ny, nz = g->nz, bb = nf * nf;
    const long n = g->n;
    if (h->opts.stage2 == TPB_S2_NONE) return;
    if (!h->Dinv) h->Dinv = (double*)malloc(sizeof(double) * bb * n);
    for (int col = 0; col < 2; col++) {
#pragma omp parallel for schedule(static)
        for (long c = 0; c < n; c++) {
            int i = (int)(c % nx), j = (int)((c / nx) % ny), k = (int)(c / ((long)nx * ny));
            if (((i + j + k) & 1) != col) continue;
            double D[9], Di[9];
            for (int r = 0; r < nf; r++)
                for (int q = 0; q < nf; q++) D[r * nf + q] = Jat(J, n, nf, 0, r, q, c);
            if (col == 1 && h->opts.stage2 == TPB_S2_ILU0) {
                for (int s = 1; s < ns; s++) {
                    long nb = nbr(nx, ny, nz, i, j, k, s);
                    if (nb < 0) continue;
                    /* D -= A_{c,nb} Dinv_nb A_{nb,c} */
                    double T1[9];
                    for (int r = 0; r < nf; r++)
                        for (int q = 0; q < nf; q++) {
                            double acc = 0.0;
                            for (int m = 0; m < nf; m++) acc += Jat(J, n, nf, s, r, m, c) * h->Dinv[(long)(m * nf + q) * n + nb];
                            T1[r * nf + q] = acc;
                        }
                    for (int r = 0; r < nf; r++)
                        for (int q = 0; q < nf; q++) {
                            double acc = 0.0;
                            for (int m = 0; m < nf; m++) acc += T1[r * nf + m] * Jat(J, n, nf, OPP[s], m, q, nb);
                            D[r * nf + q] -= acc;
                        }
                }
            }
            inv_small(nf, D, Di);
            for (int e = 0; e < bb; e++) h->Dinv[(long)e * n + c] = Di[e];
        }
    }
}

/* half sweep over one colour: out[c] = Dinv_c * (rhs[c] - sum_nb A_{c,nb} v[nb]) or with mode:
 * mode 0: z = Dinv r                       (red, forward)
 * mode 1: z = Dinv (r - sum A z[nb])       (black, forward)
 * mode 2: z = z - Dinv sum A z[nb]         (red, backward) */
/* the triangular solves read fp32-rounded coefficients (the CUDA path stores the factor in fp32), fp64 arithmetic */
static void stage2_half(const tpc_handle_s* h, const double* J, const double* r, double* z, int col, int mode) {
    const cgrid* g = &h->g;
    const int nf = h->nf, ns = g->ns, nx = g->nx, ny = g->ny, nz = g->nz;
    const long n = g->n;
#pragma omp parallel for schedule(static)
    for (long c = 0; c < n; c++) {
        int i = (int)(c % nx), j = (int)((c / nx) % ny), k = (int)(c / ((long)nx * ny));
        if (((i + j + k) & 1) != col) continue;
        double t[MAXF] = {0, 0, 0};
        if (mode != 0) {
            for (int s = 1; s < ns; s++) {
                long nb = nbr(nx, ny, nz, i, j, k, s);
                if (nb < 0) continue;
                for (int a = 0; a < nf; a++)
                    for (int q = 0; q < nf; q++) t[a] += (double)(float)Jat(J, n, nf, s, a, q, c) * z[(long)q * n + nb];
            }
        }
        double v[MAXF];
        for (int a = 0; a < nf; a++) v[a] = mode == 2 ? t[a] : r[(long)a * n + c] - t[a];
        for (int a = 0; a < nf; a++) {
            double acc = 0.0;
            for (int q = 0; q < nf; q++) acc += (double)(float)h->Dinv[(long)(a * nf + q) * n + c] * v[q];
            if (mode == 2)
                z[(long)a * n + c] -= acc;
            else
                z[(long)a * n + c] = acc;
        }
    }
}

static void stage2_apply(tpc_handle_s* h, const double* r, double* z) {
    const long n = h->g.n;
    const int nf = h->nf;
    if (h->opts.stage2 == TPB_S2_BJACOBI) {
#pragma omp parallel for schedule(static)
        for (long c = 0; c < n; c++)
            for (int a = 0; a < nf; a++) {
                double acc = 0.0;
                for (int q = 0; q < nf; q++) acc += h->Dinv[(long)(a * nf + q) * n + c] * r[(long)q * n + c];
                z[(long)a * n + c] = acc;
            }
        return;
    }
    stage2_half(h, h->J, r, z, 0, 0);
    stage2_half(h, h->J, r, z, 1, 1);
    stage2_half(h, h->J, r, z, 0, 2);
}

/* PCCOMPOSITE multiplicative: y = B1 x ; y += B2 (x - J y) */
static void pc_setup(tpc_handle_s* h, const double* J, const double* u, double dt) {
    const long nd = (long)h->nf * h->g.n;
    h->J = J;
    if (!h->t0) {
        h->t0 = (double*)malloc(sizeof(double) * nd);
        h->t1 = (double*)malloc(sizeof(double) * nd);
        h->t2 = (double*)malloc(sizeof(double) * nd);
        h->t3 = (double*)malloc(sizeof(double) * nd);
    }
    stage1_setup(h, J, u, dt);
    stage2_setup(h, J);
    h->pc_ready = 1;
}

static void pc_apply(tpc_handle_s* h, const double* x, double* y) {
    const tpb_solver_opts* o = &h->opts;
    const long nd = (long)h->nf * h->g.n;
    if (o->stage1 == TPB_S1_NONE && o->stage2 == TPB_S2_NONE) {
        memcpy(y, x, sizeof(double) * nd);
        return;
    }
    if (o->stage1 == TPB_S1_NONE) {
        stage2_apply(h, x, y);
        return;
    }
    stage1_apply(h, x, y);
    if (o->stage2 == TPB_S2_NONE || o->stage1 == TPB_S1_FIELDSPLIT) return;
    spmv(h, h->J, y, h->t0);
#pragma omp parallel for schedule(static)
    for (long q = 0; q < nd; q++) h->t0[q] = x[q] - h->t0[q];
    stage2_apply(h, h->t0, h->t1);
#pragma omp parallel for schedule(static)
    for (long q = 0; q < nd; q++) y[q] += h->t1[q];
}

/* --------------------------------------------------------------------------------------- */
/* (F)GMRES, right preconditioning, classical Gram-Schmidt (PETSc KSPGMRES defaults)        */
/* --------------------------------------------------------------------------------------- */
static double vdot(long n, const double* a, const double* b) {
    double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (long q = 0; q < n; q++) s += a[q] * b[q];
    return s;
}

static void ksp_solve(tpc_handle_s* h, const double* J, const double* b, double* x, int* its_out, int* reason_out,
                      double* rnorm_out) {
    const tpb_solver_opts* o = &h->opts;
    const long nd = (long)h->nf * h->g.n;
    const int m = o->ksp_restart, flex = (o->ksp_type == TPB_KSP_FGMRES);
    int its = 0, reason = 0;
    if (h->kcap < m + 1) {
        free(h->V);
        free(h->Z);
        h->V = (double*)malloc(sizeof(double) * nd * (m + 1));
        h->Z = flex ? (double*)malloc(sizeof(double) * nd * m) : NULL;
        h->kcap = m + 1;
    } else if (flex && !h->Z) {
        h->Z = (double*)malloc(sizeof(double) * nd * m);
    }
    double* H = (double*)calloc((size_t)(m + 1) * m, sizeof(double));
    double *cs = (double*)calloc(m, sizeof(double)), *sn = (double*)calloc(m, sizeof(double));
    double* gv = (double*)calloc(m + 1, sizeof(double));
    double* yv = (double*)calloc(m, sizeof(double));
    double* wv = (double*)malloc(sizeof(double) * nd);
    double* zt = (double*)malloc(sizeof(double) * nd);
    memset(x, 0, sizeof(double) * nd);
    double bnorm = sqrt(vdot(nd, b, b));
    double tol = fmax(o->ksp_rtol * bnorm, o->ksp_atol);
    double rnorm = bnorm;
    if (!(bnorm == bnorm) || isinf(bnorm)) {
        reason = -9;
    } else if (rnorm <= tol) {
        reason = bnorm <= o->ksp_atol ? 3 : 2;
    }
    int first = 1;
    while (reason == 0) {
        double* V = h->V;
        /* r = b - J x */
        if (first) {
            memcpy(V, b, sizeof(double) * nd);
        } else {
            spmv(h, J, x, V);
#pragma omp parallel for schedule(static)
            for (long q = 0; q < nd; q++) V[q] = b[q] - V[q];
        }
        first = 0;
        double beta = sqrt(vdot(nd, V, V));
        rnorm = beta;
        if (rnorm <= tol) {
            reason = 2;
            break;
        }
#pragma omp parallel for schedule(static)
        for (long q = 0; q < nd; q++) V[q] /= beta;
        memset(gv, 0, sizeof(double) * (m + 1));
        gv[0] = beta;
        int k = 0;
        for (; k < m && reason == 0; k++) {
            double* vk = V + (long)k * nd;
            double* zk = flex ? h->Z + (long)k * nd : zt;
            pc_apply(h, vk, zk);
            spmv(h, J, zk, wv);
            double* Hk = H + (size_t)k * (m + 1);
            for (int jv = 0; jv <= k; jv++) Hk[jv] = vdot(nd, wv, V + (long)jv * nd);
            for (int jv = 0; jv <= k; jv++) {
                const double hj = Hk[jv];
                const double* vj = V + (long)jv * nd;
#pragma omp parallel for schedule(static)
                for (long q = 0; q < nd; q++) wv[q] -= hj * vj[q];
            }
            /* second Gram-Schmidt pass once the residual is below 1e-3 ||b|| (csrc/tpb_solver.cu REFINE_AT):
             * unrefined CGS loses orthogonality like eps (||b||/||r||)^2 and stalls short of rtol = 1e-8 */
            double hn = -1.0;
            if (rnorm <= 1e-3 * bnorm) {
                double* cc = yv; /* scratch: yv is only used after the cycle */
                for (int jv = 0; jv <= k; jv++) cc[jv] = vdot(nd, wv, V + (long)jv * nd);
                const double ww = vdot(nd, wv, wv);
                double c2 = 0.0;
                for (int jv = 0; jv <= k; jv++) {
                    c2 += cc[jv] * cc[jv];
                    Hk[jv] += cc[jv];
                }
                for (int jv = 0; jv <= k; jv++) {
                    const double hj = cc[jv];
                    const double* vj = V + (long)jv * nd;
#pragma omp parallel for schedule(static)
                    for (long q = 0; q < nd; q++) wv[q] -= hj * vj[q];
                }
                if (ww - c2 > 0.25 * ww) hn = sqrt(ww - c2);
            }
            if (hn < 0.0) hn = sqrt(vdot(nd, wv, wv));
            Hk[k + 1] = hn;
            double* vn = V + (long)(k + 1) * nd;
            if (hn > 0.0) {
#pragma omp parallel for schedule(static)
                for (long q = 0; q < nd; q++) vn[q] = wv[q] / hn;
            }
            for (int jv = 0; jv < k; jv++) {
                double t = cs[jv] * Hk[jv] + sn[jv] * Hk[jv + 1];
                Hk[jv + 1] = -sn[jv] * Hk[jv] + cs[jv] * Hk[jv + 1];
                Hk[jv] = t;
            }
            double den = sqrt(Hk[k] * Hk[k] + Hk[k + 1] * Hk[k + 1]);
            if (den == 0.0) {
                reason = -5;
                break;
            }
            cs[k] = Hk[k] / den;
            sn[k] = Hk[k + 1] / den;
            Hk[k] = den;
            Hk[k + 1] = 0.0;
            gv[k + 1] = -sn[k] * gv[k];
            gv[k] = cs[k] * gv[k];
            rnorm = fabs(gv[k + 1]);
            its++;
            if (o->verbose >= 2) fprintf(stderr, "    [cpu] ksp %3d  %.6e  (h %.3e)\n", its, rnorm, hn);
            if (!(rnorm == rnorm)) reason = -9;
            else if (rnorm <= tol) reason = 2;
            else if (its >= o->ksp_max_it) reason = -3;
            else if (hn == 0.0) reason = -5;
        }
        const int kk = k; /* completed Arnoldi steps of this cycle */
        /* solve H y = g */
        for (int r = kk - 1; r >= 0; r--) {
            double acc = gv[r];
            for (int c2 = r + 1; c2 < kk; c2++) acc -= H[(size_t)c2 * (m + 1) + r] * yv[c2];
            yv[r] = acc / H[(size_t)r * (m + 1) + r];
        }
        if (flex) {
            for (int jv = 0; jv < kk; jv++) {
                const double yj = yv[jv];
                const double* zj = h->Z + (long)jv * nd;
#pragma omp parallel for schedule(static)
                for (long q = 0; q < nd; q++) x[q] += yj * zj[q];
            }
        } else {
            memset(wv, 0, sizeof(double) * nd);
            for (int jv = 0; jv < kk; jv++) {
                const double yj = yv[jv];
                const double* vj = V + (long)jv * nd;
#pragma omp parallel for schedule(static)
                for (long q = 0; q < nd; q++) wv[q] += yj * vj[q];
            }
            pc_apply(h, wv, zt);
#pragma omp parallel for schedule(static)
            for (long q = 0; q < nd; q++) x[q] += zt[q];
        }
    }
    free(H);
    free(cs);
    free(sn);
    free(gv);
    free(yv);
    free(wv);
    free(zt);
    *its_out = its;
    *reason_out = reason;
    *rnorm_out = rnorm;
}

/* --------------------------------------------------------------------------------------- */
/* SNES newtonls (thermalmodel.py:165)                                                      */
/* --------------------------------------------------------------------------------------- */
static double now_ms(void) {
#ifdef _OPENMP
    return omp_get_wtime() * 1e3;
#else
    return 0.0;
#endif
}

static void newton(tpc_handle_s* h, double* u, const double* uo, double dt, tpb_stats* st) {
    const tpb_solver_opts* o = &h->opts;
    const long nd = (long)h->nf * h->g.n;
    double* F = (double*)malloc(sizeof(double) * nd);
    double* Jm = (double*)malloc(sizeof(double) * h->g.ns * h->nf * h->nf * h->g.n);
    double* du = (double*)malloc(sizeof(double) * nd);
    double* ut = (double*)malloc(sizeof(double) * nd);
    double* Ft = (double*)malloc(sizeof(double) * nd);
    memset(st, 0, sizeof(*st));
    double t_begin = now_ms(), t0;
    t0 = now_ms();
    assemble(h, u, uo, dt, F, Jm);
    st->t_assemble_ms += now_ms() - t0;
    st->nfev = 1;
    double fnorm = sqrt(vdot(nd, F, F));
    st->fnorm0 = fnorm;
    int reason = 0, have_J = 1;
    if (!(fnorm == fnorm) || isinf(fnorm)) reason = -4;
    else if (fnorm < o->snes_atol) reason = 2;
    while (reason == 0) {
        if (st->nits >= o->snes_max_it) {
            reason = -5;
            break;
        }
        if (!have_J) {
            t0 = now_ms();
            assemble(h, u, uo, dt, F, Jm);
            st->t_assemble_ms += now_ms() - t0;
        }
        have_J = 0;
        t0 = now_ms();
        pc_setup(h, Jm, u, dt);
        st->t_pcsetup_ms += now_ms() - t0;
        int its, kr;
        double rn;
        t0 = now_ms();
        ksp_solve(h, Jm, F, du, &its, &kr, &rn);
        st->t_ksp_ms += now_ms() - t0;
        st->lits += its;
        if (o->verbose) fprintf(stderr, "  [cpu] newton %d |F| %.6e  ksp its %d reason %d\n", st->nits, fnorm, its, kr);
        if (kr < 0) {
            reason = -3;
            break;
        }
        double lambda = 1.0, fnew = 0.0;
        int ok = 0;
        for (int ls = 0; ls < (o->linesearch ? 12 : 1); ls++) {
#pragma omp parallel for schedule(static)
            for (long q = 0; q < nd; q++) ut[q] = u[q] - lambda * du[q];
            t0 = now_ms();
            assemble(h, ut, uo, dt, Ft, NULL);
            st->t_assemble_ms += now_ms() - t0;
            st->nfev++;
            fnew = sqrt(vdot(nd, Ft, Ft));
            if (!o->linesearch || (fnew == fnew && fnew <= (1.0 - 1e-4 * lambda) * fnorm)) {
                ok = 1;
                break;
            }
            lambda *= 0.5;
        }
        if (!ok) {
            reason = -6;
            break;
        }
        double unorm, dnorm = lambda * sqrt(vdot(nd, du, du));
        memcpy(u, ut, sizeof(double) * nd);
        memcpy(F, Ft, sizeof(double) * nd);
        unorm = sqrt(vdot(nd, u, u));
        fnorm = fnew;
        st->nits++;
        if (!(fnorm == fnorm) || isinf(fnorm)) reason = -4;
        else if (fnorm < o->snes_atol) reason = 2;
        else if (fnorm <= o->snes_rtol * st->fnorm0) reason = 3;
        else if (dnorm < o->snes_stol * unorm) reason = 4;
    }
    st->fnorm = fnorm;
    st->reason = reason;
    st->t_total_ms = now_ms() - t_begin;
    free(F);
    free(Jm);
    free(du);
    free(ut);
    free(Ft);
}

/* --------------------------------------------------------------------------------------- */
/* C API (mirrors include/tpb200.h with host pointers; prefix tpc_)                         */
/* --------------------------------------------------------------------------------------- */
tpc_handle_s* tpc_create(const tpb_grid* grid, int nphase, const tpb_params* prm) {
    tpc_handle_s* h = (tpc_handle_s*)calloc(1, sizeof(*h));
    cgrid* g = &h->g;
    g->dim = grid->dim;
    g->nx = grid->nx;
    g->ny = grid->ny;
    g->nz = grid->nz;
    g->ns = grid->dim == 3 ? 7 : 5;
    g->n = (long)g->nx * g->ny * g->nz;
    g->h[0] = grid->dx;
    g->h[1] = grid->dy;
    g->h[2] = grid->dim == 3 ? grid->dz : 1.0;
    if (grid->dim == 3) {
        g->area[0] = grid->dy * grid->dz;
        g->area[1] = grid->dx * grid->dz;
        g->area[2] = grid->dx * grid->dy;
        g->vol = grid->dx * grid->dy * grid->dz;
    } else {
        g->area[0] = grid->dy;
        g->area[1] = grid->dx;
        g->area[2] = 0.0;
        g->vol = grid->dx * grid->dy;
    }
    h->nphase = nphase;
    h->nf = nphase == 1 ? 2 : 3;
    cparams* P = &h->P;
    P->ko = prm->ko;
    P->kw = prm->kw;
    P->kr = prm->kr;
    P->cw = prm->c_v_w;
    P->co = prm->c_v_o;
    P->cr = prm->c_r;
    P->rho_r = prm->rho_r;
    P->T_inj = prm->T_inj;
    P->T_prod = prm->T_prod;
    P->U = prm->U;
    P->g = (grid->dim == 3 && prm->gravity) ? prm->g : 0.0;
    P->Wp = nphase == 2 ? prm->T_prod : 1.0;
    P->Wo = nphase == 2 ? prm->T_prod * (prm->c_v_w * (1.0 - prm->S_o) + prm->c_v_o * prm->S_o) : 1.0;
    P->rho_ref = 141.5 / (prm->API + 131.5) * 999.0;
    P->mu_o_pref = 1e-3 * pow(10.0, -0.8021 * prm->API + 23.8765);
    P->mu_o_exp = 0.31458 * prm->API - 9.21592;
    for (int f = 0; f < 5; f++) h->fld[f] = (double*)calloc(g->n, sizeof(double));
    return h;
}

void tpc_destroy(tpc_handle_s* h) {
    if (!h) return;
    for (int f = 0; f < 5; f++) free(h->fld[f]);
    free(h->src);
    for (int f = 0; f < MAXF; f++) free(h->w[f]);
    mg_free(&h->mg_p);
    mg_free(&h->mg_T);
    free(h->App);
    free(h->A00);
    free(h->AT);
    free(h->Dinv);
    free(h->t0);
    free(h->t1);
    free(h->t2);
    free(h->t3);
    free(h->V);
    free(h->Z);
    free(h);
}

int tpc_set_field(tpc_handle_s* h, int field, const double* data) {
    memcpy(h->fld[field], data, sizeof(double) * h->g.n);
    return 0;
}
int tpc_set_sources(tpc_handle_s* h, int n, const tpb_source* src) {
    free(h->src);
    h->src = (tpb_source*)malloc(sizeof(tpb_source) * (n > 0 ? n : 1));
    if (n > 0) memcpy(h->src, src, sizeof(tpb_source) * n);
    h->nsrc = n;
    return 0;
}
int tpc_set_solver_opts(tpc_handle_s* h, const tpb_solver_opts* o) {
    h->opts = *o;
    h->pc_ready = 0;
    return 0;
}
int tpc_assemble(tpc_handle_s* h, const double* u, const double* uo, double dt, double* F, double* J) {
    assemble(h, u, uo, dt, F, J);
    return 0;
}
int tpc_spmv(tpc_handle_s* h, const double* J, const double* x, double* y) {
    spmv(h, J, x, y);
    return 0;
}
int tpc_pc_setup(tpc_handle_s* h, const double* J, const double* u, double dt) {
    pc_setup(h, J, u, dt);
    return 0;
}
int tpc_pc_apply(tpc_handle_s* h, const double* x, double* y) {
    pc_apply(h, x, y);
    return 0;
}
int tpc_ksp_solve(tpc_handle_s* h, const double* J, const double* b, double* x, int* its, int* reason, double* rnorm) {
    ksp_solve(h, J, b, x, its, reason, rnorm);
    return 0;
}
int tpc_newton_solve(tpc_handle_s* h, double* u, const double* uo, double dt, tpb_stats* st) {
    newton(h, u, uo, dt, st);
    return 0;
}
/* introspection for component-level parity tests */
int tpc_mg_nlevels(tpc_handle_s* h, int which) { return which == 0 ? h->mg_p.nlev : h->mg_T.nlev; }
int tpc_mg_level_dims(tpc_handle_s* h, int which, int l, int* dims6) {
    mghier* m = which == 0 ? &h->mg_p : &h->mg_T;
    if (l < 0 || l >= m->nlev) return -1;
    dims6[0] = m->lev[l].nx;
    dims6[1] = m->lev[l].ny;
    dims6[2] = m->lev[l].nz;
    dims6[3] = m->lev[l].cx;
    dims6[4] = m->lev[l].cy;
    dims6[5] = m->lev[l].cz;
    return 0;
}
int tpc_mg_level_op(tpc_handle_s* h, int which, int l, double* out) {
    mghier* m = which == 0 ? &h->mg_p : &h->mg_T;
    if (l < 0 || l >= m->nlev) return -1;
    memcpy(out, m->lev[l].a, sizeof(double) * h->g.ns * m->lev[l].n);
    return 0;
}
int tpc_mg_apply(tpc_handle_s* h, int which, const double* b, double* y) {
    mg_apply(h, which == 0 ? &h->mg_p : &h->mg_T, b, y);
    return 0;
}
int tpc_stage2_apply(tpc_handle_s* h, const double* r, double* z) {
    stage2_apply(h, r, z);
    return 0;
}
int tpc_get_weights(tpc_handle_s* h, int f, double* out) {
    if (!h->w[f]) return -1;
    memcpy(out, h->w[f], sizeof(double) * h->g.n);
    return 0;
}
/* launchers such as torchrun export OMP_NUM_THREADS=1 to every rank; the CPU baseline asks for the cores back */
void tpc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int tpc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Host memory bandwidth of the box the baseline runs on (STREAM triad a = b + s*c over three arrays of n doubles, best of
 * reps, all threads; 24 bytes per element counted, write-allocate traffic not): the denominator bench.py reports the
 * port's own assembly / SpMV GB/s against. */
double tpc_stream_triad_gbs(long n, int reps) {
    double* a = (double*)malloc(sizeof(double) * n);
    double* b = (double*)malloc(sizeof(double) * n);
    double* c = (double*)malloc(sizeof(double) * n);
    if (!a || !b || !c) {
        free(a);
        free(b);
        free(c);
        return 0.0;
    }
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; i++) {
        a[i] = 0.0;
        b[i] = 1.0 + (double)(i & 7);
        c[i] = 0.5;
    }
    double best = 0.0;
    for (int r = 0; r < reps; r++) {
        const double t0 = now_ms();
#pragma omp parallel for schedule(static)
        for (long i = 0; i < n; i++) a[i] = b[i] + 3.0 * c[i];
        const double ms = now_ms() - t0;
        const double gbs = 24.0 * (double)n / (ms * 1e6);
        if (gbs > best) best = gbs;
    }
    volatile double sink = a[n / 2];
    (void)sink;
    free(a);
    free(b);
    free(c);
    return best;
}

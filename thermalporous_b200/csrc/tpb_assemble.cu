// tpb_assemble.cu - K1/K2: fused residual + Jacobian assembly of the DG0/TPFA forms.
//
// Stands in for Firedrake's assemble(F) / assemble(J) of
//   singlephase.py:120-127 (2-D), :226-235 (3-D); twophase.py:162-178 (2-D), :333-354 (3-D)
// and the source terms singlephase.py:151-165, twophase.py:388-411.
//
// Two launches per assembly:
//   props_kernel     the transcendental part of the property laws (physicalparameters.py:37-90: two exp, one
//                    pow per cell) evaluated ONCE per cell (owned + ghost planes) into a 16|40 B/cell scratch:
//                    rho_o, rho_o/mu_o [, rho_w, d rho_w/dT, rho_w/mu_w]
//   assemble_kernel  one thread per cell, x fastest => every load/store of a field or of a Jacobian slot is a
//                    coalesced 256 B row per warp; the 4|6 facet fluxes of the cell are evaluated from the row
//                    side (no atomics, no colouring) with hand-derived partials w.r.t. the 2*nf unknowns of
//                    the two cells sharing the face, upwind conditionals frozen - the exact derivative UFL's
//                    derivative() gives the reference (thermalmodel.py:36).  The Jacobian is written once, in
//                    block-stencil layout J[s][r][c][cell] (no column indices), with streaming stores.
// Compulsory traffic (3-D two-phase): 80 B in + 24 B F + 504 B J = 608 B/cell; the scratch adds 40 B written
// and 40 B read (neighbour re-reads hit L1/L2).
#include <stdlib.h>

#include "tpb_internal.cuh"

namespace {

__device__ __forceinline__ double harm(double a, double b) {
    // conditional(gt(avg(K),0), K('+')*K('-')/avg(K), 0)   singlephase.py:98
    double s = 0.5 * (a + b);
    return s > 0.0 ? a * b / s : 0.0;
}

template <int NF>
struct Fields {
    GField u[NF];
    GField phi, K[3], kT;
};

__device__ __forceinline__ double gl(const GField& f, long long c, long long n, int np) {
    return c < 0 ? f.lo[c + np] : (c >= n ? f.hi[c - n] : f.v[c]);
}

// scratch layout: q*(n + 2 np) + (np + cell), cell in [-np, n + np)
//   q: 0 rho_o, 1 m_o = rho_o/mu_o, 2 d m_o/dT [, 3 rho_w, 4 d rho_w/dT, 5 m_w = rho_w/mu_w, 6 d m_w/dT]
constexpr int NSCR1 = 3, NSCR2 = 7;

template <int NF>
__global__ void __launch_bounds__(256) props_kernel(Fields<NF> fl, long long n, int np, int has_lo, int has_hi,
                                                    DevParams P, double* __restrict__ scr, int late_wait) {
    // PDL chain of an assembly (launch_t): [sources_kernel ->] props_kernel -> assemble_kernel on one stream.
    // late_wait: the predecessor is sources_kernel, which this kernel does not depend on - it only has to finish
    // before assemble_kernel reads its output, so the wait moves to the end (every thread still waits before it
    // exits, tpb_internal.cuh) and the two run side by side.  Otherwise the predecessor may have produced u.
    pdl_launch_dependents();
    if (!late_wait) pdl_wait();
    const long long ne = n + 2LL * np;
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long c = e - np;
    if (e >= ne || (c < 0 && !has_lo) || (c >= n && !has_hi)) {
        if (late_wait) pdl_wait();
        return;
    }
    double p = gl(fl.u[0], c, n, np), T = gl(fl.u[1], c, n, np);
    double ro, ro_p, ro_T, imo, imo_T;
    oil_rho_d(P, p, T, ro, ro_p, ro_T);
    oil_imu_d(P, T, imo, imo_T);
    scr[e] = ro;
    scr[ne + e] = ro * imo;
    scr[2 * ne + e] = ro_T * imo + ro * imo_T;
    if (NF == 3) {
        double rw, rw_p, rw_T, imw, imw_T;
        water_rho_d(p, T, rw, rw_p, rw_T);
        water_imu_d(T, imw, imw_T);
        scr[3 * ne + e] = rw;
        scr[4 * ne + e] = rw_T;
        scr[5 * ne + e] = rw * imw;
        scr[6 * ne + e] = rw_T * imw + rw * imw_T;
    }
    if (late_wait) pdl_wait();
}

// static face transmissibilities area * harmonic K of the face between a cell and its +axis neighbour
// (K_facet of singlephase.py:98-101,207-211); 0 on the domain boundary.  tlo: the slab's bottom faces.
template <int DIM>
__global__ void __launch_bounds__(256) trans_kernel(GField Kx, GField Ky, GField Kz, Geom g, double* __restrict__ tx,
                                                    double* __restrict__ ty, double* __restrict__ tz,
                                                    double* __restrict__ tlo) {
    const long long n = g.n;
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const int nx = g.nx, ny = g.ny, np = g.np;
    int i, j, k;
    tpb_ijk(c, nx, ny, i, j, k);
    tx[c] = i < nx - 1 ? g.area[0] * harm(Kx.v[c], Kx.v[c + 1]) : 0.0;
    if (DIM == 2) {
        bool ex = j < ny - 1 || g.has_hi;
        ty[c] = ex ? g.area[1] * harm(Ky.v[c], gl(Ky, c + nx, n, np)) : 0.0;
        if (j == 0 && g.has_lo) tlo[c] = g.area[1] * harm(Ky.lo[c], Ky.v[c]);
    } else {
        ty[c] = j < ny - 1 ? g.area[1] * harm(Ky.v[c], Ky.v[c + nx]) : 0.0;
        bool ex = k < g.nz - 1 || g.has_hi;
        tz[c] = ex ? g.area[2] * harm(Kz.v[c], gl(Kz, c + np, n, np)) : 0.0;
        if (k == 0 && g.has_lo) tlo[c] = g.area[2] * harm(Kz.lo[c], Kz.v[c]);
    }
}

// everything the flux through a face needs from one of its two cells
template <int NF>
struct Side {
    double p, T, S;
    double ro, mo, moT;
    double rw, rwT, mw, mwT;
    double kT, kT_S;
};

template <int NF, bool HALO>
__device__ __forceinline__ Side<NF> load_side(const DevParams& P, const Fields<NF>& fl, const double* __restrict__ scr,
                                              long long n, int np, long long ne, long long c) {
    Side<NF> q;
    const long long e = c + np;
    q.p = HALO ? gl(fl.u[0], c, n, np) : fl.u[0].v[c];
    q.T = HALO ? gl(fl.u[1], c, n, np) : fl.u[1].v[c];
    q.ro = scr[e];
    q.mo = scr[ne + e];
    q.moT = scr[2 * ne + e];
    if (NF == 3) {
        q.S = HALO ? gl(fl.u[NF - 1], c, n, np) : fl.u[NF - 1].v[c];
        q.rw = scr[3 * ne + e];
        q.rwT = scr[4 * ne + e];
        q.mw = scr[5 * ne + e];
        q.mwT = scr[6 * ne + e];
        const double phi = HALO ? gl(fl.phi, c, n, np) : fl.phi.v[c];
        q.kT = phi * (P.ko * q.S + P.kw * (1.0 - q.S)) + (1.0 - phi) * P.kr;   // twophase.py:135,311
        q.kT_S = phi * (P.ko - P.kw);
    } else {
        q.S = 1.0;
        q.rw = q.rwT = q.mw = q.mwT = 0.0;
        q.kT = HALO ? gl(fl.kT, c, n, np) : fl.kT.v[c];
        q.kT_S = 0.0;
    }
    return q;
}

// Flux through one face; `pl` is the '+' cell (lower index), `mi` the '-' cell.  f[r] is added to the '+' row
// and subtracted from the '-' row (jump(test) = test+ - test-).  dP[r][c] / dM[r][c] are the partials of f[r]
// w.r.t. unknown c of the '+' / '-' cell.  aK = area * K_facet, Ak = area / Delta_h, gh = g/2 on horizontal facets.
template <int NF, bool JAC>
__device__ __forceinline__ void face_flux(const DevParams& P, const Side<NF>& pl, const Side<NF>& mi, double aK,
                                          double Ak, double ih, double gh, double* f, double (*dP)[NF],
                                          double (*dM)[NF]) {
    constexpr double CO_P = 5.5e-4, CO_T = -2.5e-4, CW_P = 3.98854e-4;   // physicalparameters.py:41-46,77-82
    const double dp = ih * (pl.p - mi.p);                       // jump(p)/Delta_h
    const double dTj = pl.T - mi.T;
    // ---- oil phase (the only phase of the single-phase model): singlephase.py:215, twophase.py:318
    const double flo = dp - gh * (pl.ro + mi.ro);
    const bool upo = flo > 0.0;
    const Side<NF>& uo = upo ? pl : mi;
    const double So = NF == 3 ? uo.S : 1.0;                     // rel_perm_o = S_o (:92-94)
    const double lamo = So * uo.mo;
    const double Tuo = uo.T;
    const double Fo = aK * lamo * flo;                          // a_flow (:121,227-228 | :338-339)
    double Fo_P[3], Fo_M[3];
    if (JAC) {
        // partials of the upwinded mobility w.r.t. the upwind cell's (p, T, S)
        const double a_p = aK * flo * CO_P * lamo, a_T = aK * flo * So * uo.moT, a_S = aK * flo * uo.mo;
        const double al = aK * lamo;
        Fo_P[0] = al * (ih - gh * CO_P * pl.ro) + (upo ? a_p : 0.0);
        Fo_P[1] = al * (-gh * CO_T * pl.ro) + (upo ? a_T : 0.0);
        Fo_P[2] = upo ? a_S : 0.0;
        Fo_M[0] = al * (-ih - gh * CO_P * mi.ro) + (upo ? 0.0 : a_p);
        Fo_M[1] = al * (-gh * CO_T * mi.ro) + (upo ? 0.0 : a_T);
        Fo_M[2] = upo ? 0.0 : a_S;
    }
    if (NF == 2) {
        const double kTf = harm(pl.kT, mi.kT);
        const double ce = P.c_v_o * Tuo;
        f[0] = Fo;
        f[1] = ce * Fo + Ak * kTf * dTj;                        // a_advec + a_diff (:124-125,231-233)
        if (JAC) {
#pragma unroll
            for (int c = 0; c < NF; c++) {
                dP[0][c] = Fo_P[c];
                dM[0][c] = Fo_M[c];
                dP[1][c] = ce * Fo_P[c];
                dM[1][c] = ce * Fo_M[c];
            }
            dP[1][1] += (upo ? P.c_v_o * Fo : 0.0) + Ak * kTf;
            dM[1][1] += (upo ? 0.0 : P.c_v_o * Fo) - Ak * kTf;
        }
    } else {
        // ---- water phase: twophase.py:317,334-335
        const double flw = dp - gh * (pl.rw + mi.rw);
        const bool upw = flw > 0.0;
        const Side<NF>& uw = upw ? pl : mi;
        const double Sw = 1.0 - uw.S;                           // rel_perm_w = 1 - S_o (:96-98)
        const double lamw = Sw * uw.mw;
        const double Tuw = uw.T;
        const double Fw = aK * lamw * flw;
        // harmonic conductivity, state dependent (:315)
        const double ks = 0.5 * (pl.kT + mi.kT);
        const double iks = ks > 0.0 ? 1.0 / ks : 0.0;
        const double kTf = pl.kT * mi.kT * iks;
        const double cew = P.c_v_w * Tuw, ceo = P.c_v_o * Tuo;
        f[0] = P.Wp * (P.c_v_w * Fw + P.c_v_o * Fo);           // weighted-sum pressure equation (:343-346)
        f[1] = cew * Fw + ceo * Fo + Ak * kTf * dTj;            // :350-352
        f[2] = P.Wo * Fo;
        if (JAC) {
            double Fw_P[3], Fw_M[3];
            const double b_p = aK * flw * CW_P * lamw, b_T = aK * flw * Sw * uw.mwT, b_S = -aK * flw * uw.mw;
            const double bl = aK * lamw;
            Fw_P[0] = bl * (ih - gh * CW_P * pl.rw) + (upw ? b_p : 0.0);
            Fw_P[1] = bl * (-gh * pl.rwT) + (upw ? b_T : 0.0);
            Fw_P[2] = upw ? b_S : 0.0;
            Fw_M[0] = bl * (-ih - gh * CW_P * mi.rw) + (upw ? 0.0 : b_p);
            Fw_M[1] = bl * (-gh * mi.rwT) + (upw ? 0.0 : b_T);
            Fw_M[2] = upw ? 0.0 : b_S;
            const double wpw = P.Wp * P.c_v_w, wpo = P.Wp * P.c_v_o;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                dP[0][c] = wpw * Fw_P[c] + wpo * Fo_P[c];
                dM[0][c] = wpw * Fw_M[c] + wpo * Fo_M[c];
                dP[1][c] = cew * Fw_P[c] + ceo * Fo_P[c];
                dM[1][c] = cew * Fw_M[c] + ceo * Fo_M[c];
                dP[2][c] = P.Wo * Fo_P[c];
                dM[2][c] = P.Wo * Fo_M[c];
            }
            const double adv = P.c_v_w * Fw, ado = P.c_v_o * Fo, cond = Ak * kTf;
            dP[1][1] += (upw ? adv : 0.0) + (upo ? ado : 0.0) + cond;
            dM[1][1] += (upw ? 0.0 : adv) + (upo ? 0.0 : ado) - cond;
            // d kTf / d k+ = k-^2 / (2 ks^2), and symmetrically
            const double hk = 0.5 * iks * iks * Ak * dTj;
            dP[1][2] += hk * mi.kT * mi.kT * pl.kT_S;
            dM[1][2] += hk * pl.kT * pl.kT * mi.kT_S;
        }
    }
}

struct Trans {
    const double* t[3];
    const double* lo;
};

// what a cell's row needs from global memory besides the Sides: face transmissibilities, old state, source slot
template <int NF, int DIM>
struct CellIn {
    double aK[2 * DIM + 1];   // area * K_facet per stencil slot (0 where the face does not exist)
    bool ex[2 * DIM + 1];
    double uo[NF];
    int si;
};

template <int NF, int DIM>
__device__ __forceinline__ CellIn<NF, DIM> load_cell_in(const Geom& g, const Trans& tr, const double* __restrict__ u_old,
                                                        const int* __restrict__ src_index, long long cell, int i, int j,
                                                        int k) {
    CellIn<NF, DIM> in;
    const long long n = g.n;
    const int nx = g.nx, ny = g.ny, np = g.np;
    in.aK[0] = 0.0;
    in.ex[0] = true;
#pragma unroll
    for (int s = 1; s < 2 * DIM + 1; s++) {
        const int axis = (s - 1) >> 1;
        const bool hi_side = ((s - 1) & 1) != 0;   // neighbour has the higher index => this cell is '+'
        bool exists;
        long long nb;
        double aK;                                 // area * K_facet
        if (axis == 0) {
            exists = hi_side ? (i < nx - 1) : (i > 0);
            nb = cell + (hi_side ? 1 : -1);
            aK = exists ? tr.t[0][hi_side ? cell : nb] : 0.0;
        } else if (axis == 1) {
            if (DIM == 2) {
                exists = hi_side ? (j < ny - 1 || g.has_hi) : (j > 0 || g.has_lo);
                nb = cell + (hi_side ? nx : -nx);
                aK = !exists ? 0.0 : (hi_side ? tr.t[1][cell] : (nb >= 0 ? tr.t[1][nb] : tr.lo[cell]));
            } else {
                exists = hi_side ? (j < ny - 1) : (j > 0);
                nb = cell + (hi_side ? nx : -nx);
                aK = exists ? tr.t[1][hi_side ? cell : nb] : 0.0;
            }
        } else {
            exists = hi_side ? (k < g.nz - 1 || g.has_hi) : (k > 0 || g.has_lo);
            nb = cell + (hi_side ? (long long)np : -(long long)np);
            aK = !exists ? 0.0 : (hi_side ? tr.t[2][cell] : (nb >= 0 ? tr.t[2][nb] : tr.lo[cell]));
        }
        in.aK[s] = aK;
        in.ex[s] = exists;
    }
#pragma unroll
    for (int f = 0; f < NF; f++) in.uo[f] = u_old[(long long)f * n + cell];
    in.si = src_index ? src_index[cell] : -1;
    return in;
}

// One cell's row: accumulation, the 2*DIM facet integrals seen from the row side, source cells, stores.  `nbr(s, nb)`
// returns the Side of the neighbour through stencil slot s (cell index nb).
template <int NF, int DIM, bool JAC, class NbrF>
__device__ __forceinline__ void assemble_cell(const DevParams& P, const Geom& g, const CellIn<NF, DIM>& in,
                                              const double* __restrict__ src_acc, double idt, long long cell,
                                              const Side<NF>& me, double phi, NbrF&& nbr, double* __restrict__ F,
                                              double* __restrict__ J) {
    const long long n = g.n;
    const int nx = g.nx, np = g.np;
    double R[NF];
    double D[NF][NF];

    // ---- accumulation (cell integrals) ------------------------------------------------------
    {
        constexpr double CO_P = 5.5e-4, CO_T = -2.5e-4, CW_P = 3.98854e-4;
        const double po = in.uo[0], To = in.uo[1];
        const double w = g.vol * idt;
        const double ro_old = oil_rho_v(P, po, To);
        const double rk = w * (1.0 - phi) * P.rho_r * P.c_r;
        const double wp = w * phi;
        if (NF == 2) {
            R[0] = wp * (me.ro - ro_old);                                              // singlephase.py:120
            R[1] = wp * P.c_v_o * (me.ro * me.T - ro_old * To) + rk * (me.T - To);      // :123
            if (JAC) {
                D[0][0] = wp * CO_P * me.ro;
                D[0][1] = wp * CO_T * me.ro;
                D[1][0] = wp * P.c_v_o * CO_P * me.ro * me.T;
                D[1][1] = wp * P.c_v_o * (CO_T * me.ro * me.T + me.ro) + rk;
            }
        } else {
            const double So = in.uo[NF - 1];
            const double rw_old = water_rho_v(po, To);
            const double S = me.S, Sw = 1.0 - me.S, T = me.T;
            const double aw = wp * (me.rw * Sw - rw_old * (1.0 - So));                  // twophase.py:333
            const double ao = wp * (me.ro * S - ro_old * So);                           // :337
            R[0] = P.Wp * (P.c_v_w * aw + P.c_v_o * ao);                                // :344
            R[1] = wp * P.c_v_w * (me.rw * Sw * T - rw_old * (1.0 - So) * To) +
                   wp * P.c_v_o * (me.ro * S * T - ro_old * So * To) + rk * (T - To);   // :349
            R[NF - 1] = P.Wo * ao;
            if (JAC) {
                const double aw_p = wp * CW_P * me.rw * Sw, aw_T = wp * me.rwT * Sw, aw_S = -wp * me.rw;
                const double ao_p = wp * CO_P * me.ro * S, ao_T = wp * CO_T * me.ro * S, ao_S = wp * me.ro;
                const double h_p = P.c_v_w * aw_p + P.c_v_o * ao_p, h_T = P.c_v_w * aw_T + P.c_v_o * ao_T,
                             h_S = P.c_v_w * aw_S + P.c_v_o * ao_S;
                D[0][0] = P.Wp * h_p;
                D[0][1] = P.Wp * h_T;
                D[0][NF - 1] = P.Wp * h_S;
                D[1][0] = h_p * T;
                D[1][1] = h_T * T + wp * (P.c_v_w * me.rw * Sw + P.c_v_o * me.ro * S) + rk;
                D[1][NF - 1] = h_S * T;
                D[NF - 1][0] = P.Wo * ao_p;
                D[NF - 1][1] = P.Wo * ao_T;
                D[NF - 1][NF - 1] = P.Wo * ao_S;
            }
        }
    }

    // ---- facet integrals ------------------------------------------------------------------------
#pragma unroll
    for (int s = 1; s < 2 * DIM + 1; s++) {
        const int axis = (s - 1) >> 1;
        const bool hi_side = ((s - 1) & 1) != 0;   // neighbour has the higher index => this cell is '+'
        const bool exists = in.ex[s];
        const double aK = in.aK[s];
        const long long nb = cell + (axis == 0 ? (hi_side ? 1 : -1)
                                               : (axis == DIM - 1 ? (hi_side ? (long long)np : -(long long)np)
                                                                  : (hi_side ? (long long)nx : -(long long)nx)));
        double O[NF][NF];
#pragma unroll
        for (int r = 0; r < NF; r++)
#pragma unroll
            for (int c = 0; c < NF; c++) O[r][c] = 0.0;

        if (exists) {
            const Side<NF> ot = nbr(s, nb);
            const double gh = (axis == 2) ? 0.5 * P.g : 0.0;
            const double ih = g.ih[axis], Ak = g.area[axis] * ih;
            double f[NF], dP[NF][NF], dM[NF][NF];
            if (hi_side) {
                face_flux<NF, JAC>(P, me, ot, aK, Ak, ih, gh, f, dP, dM);
#pragma unroll
                for (int r = 0; r < NF; r++) {
                    R[r] += f[r];
                    if (JAC) {
#pragma unroll
                        for (int c = 0; c < NF; c++) {
                            D[r][c] += dP[r][c];
                            O[r][c] = dM[r][c];
                        }
                    }
                }
            } else {
                face_flux<NF, JAC>(P, ot, me, aK, Ak, ih, gh, f, dP, dM);
#pragma unroll
                for (int r = 0; r < NF; r++) {
                    R[r] -= f[r];
                    if (JAC) {
#pragma unroll
                        for (int c = 0; c < NF; c++) {
                            D[r][c] -= dM[r][c];
                            O[r][c] = -dP[r][c];
                        }
                    }
                }
            }
        }
        if (JAC) {
            double* jp = J + (long long)(s * NF * NF) * n + cell;
#pragma unroll
            for (int r = 0; r < NF; r++)
#pragma unroll
                for (int c = 0; c < NF; c++) {
                    __stcs(jp, O[r][c]);
                    jp += n;
                }
        }
    }

    {   // wells / heaters in this cell (sources_kernel)
        const int si = in.si;
        if (si >= 0) {
            const double* o = src_acc + (long long)si * (NF + NF * NF);
#pragma unroll
            for (int r = 0; r < NF; r++) {
                R[r] += o[r];
                if (JAC) {
#pragma unroll
                    for (int c = 0; c < NF; c++) D[r][c] += o[NF + r * NF + c];
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < NF; r++) F[(long long)r * n + cell] = R[r];
    if (JAC) {
        double* jp = J + cell;
#pragma unroll
        for (int r = 0; r < NF; r++)
#pragma unroll
            for (int c = 0; c < NF; c++) {
                *jp = D[r][c];
                jp += n;
            }
    }
}

// K1/K2: thread per cell, neighbour values straight from global memory (L1/L2 serve the re-reads).  A plane-marching
// variant with cp.async-staged shared-memory tiles (32x8 .. 64x2 threads, three rotating plane buffers, own column's
// lower plane in registers) produced the same bits and was measured 1.4-2.2x SLOWER on B200 at 60x220x85 (0.224-0.349
// ms against 0.156 ms): the per-plane block barrier puts a block's warps in lock-step and the staging costs 15
// cp.async per cell, more than the L1/L2 re-reads it saves.  It was removed; DESIGN.md section 4 keeps the numbers.
template <int NF, int DIM, bool JAC, bool HALO, int MINB>
__global__ void __launch_bounds__(128, MINB) assemble_kernel(Fields<NF> fl, const double* __restrict__ u_old,
                                                             const double* __restrict__ scr, Trans tr,
                                                             const int* __restrict__ src_index,
                                                             const double* __restrict__ src_acc, double idt, Geom g,
                                                             DevParams P, double* __restrict__ F,
                                                             double* __restrict__ J) {
    const long long n = g.n;
    long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    // scr comes from props_kernel and src_acc from sources_kernel, the two predecessors (PDL chain, see
    // props_kernel); u, u_old and the static fields were complete before either started
    pdl_wait();
    if (cell >= n) return;
    const int nx = g.nx, ny = g.ny, np = g.np;
    const long long ne = n + 2LL * np;
    int i, j, k;
    tpb_ijk(cell, nx, ny, i, j, k);
    const Side<NF> me = load_side<NF, false>(P, fl, scr, n, np, ne, cell);
    const double phi = fl.phi.v[cell];
    auto nbr = [&](int s, long long nb) {
        const bool slab_axis = ((s - 1) >> 1) == DIM - 1;
        return (HALO && slab_axis) ? load_side<NF, true>(P, fl, scr, n, np, ne, nb)
                                   : load_side<NF, false>(P, fl, scr, n, np, ne, nb);
    };
    const CellIn<NF, DIM> in = load_cell_in<NF, DIM>(g, tr, u_old, src_index, cell, i, j, k);
    assemble_cell<NF, DIM, JAC>(P, g, in, src_acc, idt, cell, me, phi, nbr, F, J);
}

// ---- well / heater source terms ---------------------------------------------------------------
// wellcase.py:180-192: Peaceman well index with h = 5, rw = 0.1, Dx = Dy = 5 hard-wired
__device__ __forceinline__ double peaceman_wi(double Kx, double Ky) {
    const double hh = 5.0, rw = 0.1, Dx = 5.0, Dy = 5.0;
    double a = Ky / Kx, b = Kx / Ky;
    double ro = 0.28 * sqrt(sqrt(a) * Dx * Dx + sqrt(b) * Dy * Dy) / (sqrt(sqrt(a)) + sqrt(sqrt(b)));
    double Ke = sqrt(Kx * Ky);
    return 2.0 * 3.141592653589793 * hh * Ke / log(ro / rw);
}

// rate = conditional(|WI/mu*dd| >= |max_rate|, max_rate, WI/mu*dd)   wellcase.py:191-199
template <int NF>
__device__ __forceinline__ Dual<NF> well_rate(const tpb_source& s, double wi, const Dual<NF>& imu, const Dual<NF>& p) {
    if (s.const_rate) return dconst<NF>(s.max_rate);
    Dual<NF> d = s.bhp - p;
    Dual<NF> dd;
    if (s.max_rate < 0.0)
        dd = (d.v >= 0.0) ? dconst<NF>(0.0) : d;
    else
        dd = (d.v <= 0.0) ? dconst<NF>(0.0) : d;
    Dual<NF> rate = wi * (imu * dd);
    if (fabs(rate.v) - fabs(s.max_rate) >= 0.0) return dconst<NF>(s.max_rate);
    return rate;
}

// One thread per source cell; runs on the handle's side stream concurrently with props_kernel and leaves the
// residual / diagonal-block contributions of the cell's wells and heaters in acc_out[t][NF + NF*NF], which
// assemble_kernel adds to the row it is writing anyway (src_index[cell] = t, or -1).
template <int NF, bool JAC>
__global__ void sources_kernel(int ncells, const int64_t* __restrict__ cells, const int* __restrict__ off,
                               const tpb_source* __restrict__ ent, const double* __restrict__ u,
                               const double* __restrict__ Kx, const double* __restrict__ Ky, long long n, DevParams P,
                               double* __restrict__ acc_out) {
    pdl_launch_dependents();   // props_kernel may run beside this kernel (it waits for it at its end)
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncells) return;
    long long cell = cells[t];
    double uc[NF];
#pragma unroll
    for (int f = 0; f < NF; f++) uc[f] = u[(long long)f * n + cell];
    Dual<NF> p = dvar<NF>(uc[0], 0), T = dvar<NF>(uc[1], 1);
    Dual<NF> acc[NF];
#pragma unroll
    for (int r = 0; r < NF; r++) acc[r] = dconst<NF>(0.0);

    double ro, ro_p, ro_T, imo, imo_T;
    oil_rho_d(P, uc[0], uc[1], ro, ro_p, ro_T);
    oil_imu_d(P, uc[1], imo, imo_T);
    Dual<NF> rho_o = dconst<NF>(ro);
    rho_o.d[0] = ro_p;
    rho_o.d[1] = ro_T;
    Dual<NF> im_o = dconst<NF>(imo);
    im_o.d[1] = imo_T;
    double wi = 0.0;
    bool wi_done = false;

    for (int e = off[t]; e < off[t + 1]; e++) {
        tpb_source s = ent[e];
        double w = s.weight;
        if (s.kind == TPB_HEATER) {
            // F -= delta*U*(T_inj - T)*r*dx   singlephase.py:163-165, twophase.py:409-411
            acc[1] = acc[1] - (w * P.U) * (P.T_inj - T);
            continue;
        }
        if (!wi_done && !s.const_rate) {
            wi = peaceman_wi(Kx[cell], Ky[cell]);
            wi_done = true;
        }
        if constexpr (NF == 2) {
            Dual<NF> q = well_rate<NF>(s, wi, im_o, p);   // flow_rate(p, T, well): oil viscosity at the cell
            if (s.kind == TPB_PROD) {
                Dual<NF> m = w * (rho_o * q);             // singlephase.py:151-156
                acc[0] = acc[0] - m;
                acc[1] = acc[1] - P.c_v_o * (m * T);
            } else {
                double ri, ri_p, ri_T;
                oil_rho_d(P, uc[0], P.T_inj, ri, ri_p, ri_T);   // rhow = oil_rho(p, T_inj)  :131
                Dual<NF> rinj = dconst<NF>(ri);
                rinj.d[0] = ri_p;
                Dual<NF> m = w * (rinj * q);              // :157-162
                acc[0] = acc[0] - m;
                acc[1] = acc[1] - (P.c_v_o * P.T_inj) * m;
            }
        } else {
            Dual<NF> S = dvar<NF>(uc[2], 2);
            double rw, rw_p, rw_T, imw, imw_T;
            water_rho_d(uc[0], uc[1], rw, rw_p, rw_T);
            water_imu_d(uc[1], imw, imw_T);
            Dual<NF> im_w = dconst<NF>(imw);
            im_w.d[1] = imw_T;
            if (s.kind == TPB_PROD) {
                Dual<NF> rho_w = dconst<NF>(rw);
                rho_w.d[0] = rw_p;
                rho_w.d[1] = rw_T;
                // mu = 1/(S/mu_o + (1-S)/mu_w)   wellcase.py:212
                Dual<NF> mob_o = S * im_o, mob_w = (1.0 - S) * im_w;
                Dual<NF> imu = mob_o + mob_w;
                Dual<NF> q = well_rate<NF>(s, wi, imu, p);
                Dual<NF> mu = 1.0 / imu;
                Dual<NF> qw = mob_w * mu * q;             // :233
                Dual<NF> qo = mob_o * mu * q;             // :234
                Dual<NF> mw = rho_w * qw, mo = rho_o * qo;
                acc[0] = acc[0] - (P.Wp * w) * (P.c_v_w * mw + P.c_v_o * mo);   // twophase.py:396
                acc[2] = acc[2] - (P.Wo * w) * mo;
                acc[1] = acc[1] - w * ((P.c_v_w * mw + P.c_v_o * mo) * T);      // :399
            } else {
                Dual<NF> q = well_rate<NF>(s, wi, im_w, p);   // flow_rate(..., phase='water')  :401
                double ri, ri_p, ri_T;
                water_rho_d(uc[0], P.T_inj, ri, ri_p, ri_T);  // rhow = water_rho(p_w, T_inj)   :359
                Dual<NF> rinj = dconst<NF>(ri);
                rinj.d[0] = ri_p;
                Dual<NF> m = w * (rinj * q);
                acc[0] = acc[0] - (P.Wp * P.c_v_w) * m;       // :405
                acc[1] = acc[1] - (P.c_v_w * P.T_inj) * m;    // :408
            }
        }
    }
    double* o = acc_out + (long long)t * (NF + NF * NF);
#pragma unroll
    for (int r = 0; r < NF; r++) {
        o[r] = acc[r].v;
        if (JAC) {
#pragma unroll
            for (int c = 0; c < NF; c++) o[NF + r * NF + c] = acc[r].d[c];
        }
    }
}

template <int NF, int DIM, bool HALO>
void launch_k(tpb_handle_s* h, const Fields<NF>& fl, const double* u_old, double dt, double* F, double* J) {
    const long long n = h->g.n;
    Trans tr;
    tr.t[0] = h->trans[0];
    tr.t[1] = h->trans[1];
    tr.t[2] = h->trans[2];
    tr.lo = h->trans_lo;
    const int threads = 128;
    const unsigned blocks = (unsigned)((n + threads - 1) / threads);
    static int variant = getenv("TPB_ASM_MINB") ? atoi(getenv("TPB_ASM_MINB")) : 4;
#define TPB_ASM(JAC, MINB) \
    launch_pdl(assemble_kernel<NF, DIM, JAC, HALO, MINB>, blocks, threads, h->stream, fl, u_old, (const double*)h->scr, tr, (const int*)(h->nsrc_cells > 0 ? h->src_index : nullptr), (const double*)h->src_acc, 1.0 / dt, h->g, h->dp, F, J)
    if (J) {
        if (variant == 3)
            TPB_ASM(true, 3);
        else if (variant == 5)
            TPB_ASM(true, 5);
        else if (variant == 6)
            TPB_ASM(true, 6);
        else
            TPB_ASM(true, 4);
    } else {
        TPB_ASM(false, 6);
    }
#undef TPB_ASM
}

template <int NF, int DIM>
void launch_t(tpb_handle_s* h, const double* u, const double* u_old, double dt, double* F, double* J) {
    Fields<NF> fl;
    const long long n = h->g.n;
    const int np = h->g.np;
    for (int f = 0; f < NF; f++) {
        fl.u[f].v = u + (size_t)f * n;
        fl.u[f].lo = h->u_lo + (size_t)f * np;
        fl.u[f].hi = h->u_hi + (size_t)f * np;
    }
    auto gf = [&](int id) {
        GField x;
        x.v = h->fld[id];
        x.lo = h->fld_lo[id];
        x.hi = h->fld_hi[id];
        return x;
    };
    fl.phi = gf(TPB_PHI);
    fl.K[0] = gf(TPB_KX);
    fl.K[1] = gf(TPB_KY);
    fl.K[2] = gf(DIM == 3 ? TPB_KZ : TPB_KY);
    fl.kT = gf(TPB_KT);
    const long long ne = n + 2LL * np;
    if (!h->scr) h->scr = tpb_dalloc<double>((size_t)(NF == 3 ? NSCR2 : NSCR1) * ne);
    if (h->trans_dirty) {
        for (int a = 0; a < 3; a++)
            if (!h->trans[a]) h->trans[a] = tpb_dalloc<double>(n);
        if (!h->trans_lo) h->trans_lo = tpb_dalloc<double>(np);
        trans_kernel<DIM><<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(fl.K[0], fl.K[1], fl.K[2], h->g, h->trans[0],
                                                                          h->trans[1], h->trans[2], h->trans_lo);
        h->launches++;
        h->trans_dirty = false;
    }
    // One stream, three kernels chained by programmatic dependent launch: the few source cells (a normal launch:
    // everything before it is complete when it starts), the property pre-pass running beside them, then the
    // flux kernel, whose blocks start while the pre-pass drains and wait for both before they read the scratch.
    const bool pdl = tpb_pdl_enabled();
    if (h->nsrc_cells > 0) {
        const unsigned sb = (unsigned)((h->nsrc_cells + 127) / 128);
        if (J)
            sources_kernel<NF, true><<<sb, 128, 0, h->stream>>>(h->nsrc_cells, h->src_cell, h->src_off, h->src_ent, u,
                                                               h->fld[TPB_KX], h->fld[TPB_KY], n, h->dp, h->src_acc);
        else
            sources_kernel<NF, false><<<sb, 128, 0, h->stream>>>(h->nsrc_cells, h->src_cell, h->src_off, h->src_ent, u,
                                                                h->fld[TPB_KX], h->fld[TPB_KY], n, h->dp, h->src_acc);
        h->launches++;
    }
    launch_pdl(props_kernel<NF>, (unsigned)((ne + 255) / 256), 256, h->stream, fl, n, np, h->g.has_lo, h->g.has_hi, h->dp,
               h->scr, (pdl && h->nsrc_cells > 0) ? 1 : 0);
    h->launches++;
    if (h->g.has_lo || h->g.has_hi)
        launch_k<NF, DIM, true>(h, fl, u_old, dt, F, J);
    else
        launch_k<NF, DIM, false>(h, fl, u_old, dt, F, J);
    h->launches++;
    TPB_CUDA(cudaGetLastError());
}

}  // namespace

void tpb_launch_assemble(tpb_handle_s* h, const double* u, const double* u_old, double dt, double* F, double* J) {
    if (h->nf == 2) {
        if (h->g.dim == 2)
            launch_t<2, 2>(h, u, u_old, dt, F, J);
        else
            launch_t<2, 3>(h, u, u_old, dt, F, J);
    } else {
        if (h->g.dim == 2)
            launch_t<3, 2>(h, u, u_old, dt, F, J);
        else
            launch_t<3, 3>(h, u, u_old, dt, F, J);
    }
}

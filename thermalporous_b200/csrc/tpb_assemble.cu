// tpb_assemble.cu - K1/K2: fused residual + Jacobian assembly of the DG0/TPFA forms.
//
// Stands in for Firedrake's assemble(F) / assemble(J) of
//   singlephase.py:120-127 (2-D), :226-235 (3-D); twophase.py:162-178 (2-D), :333-354 (3-D)
// and the source terms singlephase.py:151-165, twophase.py:388-411.
//
// One thread per cell, x fastest => every load/store of a field or of a Jacobian slot is a
// fully coalesced 256 B row per warp.  The Jacobian is written in block-stencil layout
// J[s][r][c][cell] (no column indices).  Facet fluxes are evaluated from the row cell's side
// for each of its 4|6 faces (no atomics, no colouring); the derivative blocks come from
// forward-mode duals over the 2*nf unknowns of the two cells sharing the face, so the
// Jacobian is the exact derivative with the upwind conditionals frozen - what UFL's
// derivative() gives the reference (thermalmodel.py:36).
#include "tpb_internal.cuh"

namespace {

// per-cell quantities with partials w.r.t. the cell's own unknowns
template <int NF>
struct CellProps {
    Dual<NF> p, T;
    Dual<NF> S;        // two-phase only
    Dual<NF> rho_o;    // oil density
    Dual<NF> rho_w;    // two-phase only
    Dual<NF> lam_o;    // k_ro rho_o / mu_o   (single-phase: rho_o / mu_o)
    Dual<NF> lam_w;    // k_rw rho_w / mu_w
    Dual<NF> kT;       // conductivity (constant partials for single-phase)
};

template <int NF>
__device__ __forceinline__ CellProps<NF> cell_props(const DevParams& P, const double* u, double phi, double kT_static) {
    CellProps<NF> c;
    c.p = dvar<NF>(u[0], 0);
    c.T = dvar<NF>(u[1], 1);
    double ro, ro_p, ro_T, imo, imo_T;
    oil_rho_d(P, u[0], u[1], ro, ro_p, ro_T);
    oil_imu_d(P, u[1], imo, imo_T);
    c.rho_o = dconst<NF>(ro);
    c.rho_o.d[0] = ro_p;
    c.rho_o.d[1] = ro_T;
    Dual<NF> im_o = dconst<NF>(imo);
    im_o.d[1] = imo_T;
    if constexpr (NF == 3) {
        c.S = dvar<NF>(u[2], 2);
        double rw, rw_p, rw_T, imw, imw_T;
        water_rho_d(u[0], u[1], rw, rw_p, rw_T);
        water_imu_d(u[1], imw, imw_T);
        c.rho_w = dconst<NF>(rw);
        c.rho_w.d[0] = rw_p;
        c.rho_w.d[1] = rw_T;
        Dual<NF> im_w = dconst<NF>(imw);
        im_w.d[1] = imw_T;
        c.lam_o = c.S * c.rho_o * im_o;            // rel_perm_o = S_o  (physicalparameters.py:92-94)
        c.lam_w = (1.0 - c.S) * c.rho_w * im_w;    // rel_perm_w = 1 - S_o (:96-98)
        // kT = phi*(S ko + (1-S) kw) + (1-phi) kr   (twophase.py:135,311)
        c.kT = phi * (P.ko * c.S + P.kw * (1.0 - c.S)) + (1.0 - phi) * P.kr;
    } else {
        c.S = dconst<NF>(0.0);
        c.rho_w = dconst<NF>(0.0);
        c.lam_w = dconst<NF>(0.0);
        c.lam_o = c.rho_o * im_o;
        c.kT = dconst<NF>(kT_static);
    }
    return c;
}

template <int NF, int OFF>
__device__ __forceinline__ void embed_props(const CellProps<NF>& a, CellProps<2 * NF>& b) {
    b.p = dembed<2 * NF, OFF, NF>(a.p);
    b.T = dembed<2 * NF, OFF, NF>(a.T);
    b.S = dembed<2 * NF, OFF, NF>(a.S);
    b.rho_o = dembed<2 * NF, OFF, NF>(a.rho_o);
    b.rho_w = dembed<2 * NF, OFF, NF>(a.rho_w);
    b.lam_o = dembed<2 * NF, OFF, NF>(a.lam_o);
    b.lam_w = dembed<2 * NF, OFF, NF>(a.lam_w);
    b.kT = dembed<2 * NF, OFF, NF>(a.kT);
}

__device__ __forceinline__ double harm(double a, double b) {
    // conditional(gt(avg(K),0), K('+')*K('-')/avg(K), 0)   singlephase.py:98
    double s = 0.5 * (a + b);
    return s > 0.0 ? a * b / s : 0.0;
}

// Fluxes through one face; `pl` is the '+' cell (lower index), `mi` the '-' cell.
// f[r] is added to the '+' row and subtracted from the '-' row (jump(test) = test+ - test-).
// partial slots: [0, NF) = '+' unknowns, [NF, 2NF) = '-' unknowns.
template <int NF>
__device__ __forceinline__ void face_flux(const DevParams& P, const CellProps<2 * NF>& pl,
                                          const CellProps<2 * NF>& mi, double Kf, double area, double ih, double grav,
                                          Dual<2 * NF>* f) {
    constexpr int NV = 2 * NF;
    Dual<NV> dp = ih * (pl.p - mi.p);                    // jump(p)/Delta_h
    Dual<NV> dT = ih * (pl.T - mi.T);
    double aK = area * Kf;
    if constexpr (NF == 2) {
        // singlephase.py:215 z_flow = jump(p)/Delta_h - g*avg(rho_o); lateral: jump(p)/Delta_h
        Dual<NV> fl = dp - (0.5 * grav) * (pl.rho_o + mi.rho_o);
        bool up = fl.v > 0.0;
        Dual<NV> lam = up ? pl.lam_o : mi.lam_o;
        Dual<NV> Tup = up ? pl.T : mi.T;
        Dual<NV> fm = aK * (lam * fl);                   // a_flow (:121,227-228)
        Dual<NV> kTf = dconst<NV>(harm(pl.kT.v, mi.kT.v));
        f[0] = fm;
        f[1] = P.c_v_o * (Tup * fm) + area * (kTf * dT); // a_advec + a_diff (:124-125,231-233)
    } else {
        // twophase.py:317-318
        Dual<NV> fl_w = dp - (0.5 * grav) * (pl.rho_w + mi.rho_w);
        Dual<NV> fl_o = dp - (0.5 * grav) * (pl.rho_o + mi.rho_o);
        bool upw = fl_w.v > 0.0, upo = fl_o.v > 0.0;
        Dual<NV> fw = aK * ((upw ? pl.lam_w : mi.lam_w) * fl_w);   // :334-335
        Dual<NV> fo = aK * ((upo ? pl.lam_o : mi.lam_o) * fl_o);   // :338-339
        Dual<NV> few = P.c_v_w * ((upw ? pl.T : mi.T) * fw);       // :350-351
        Dual<NV> feo = P.c_v_o * ((upo ? pl.T : mi.T) * fo);
        // harmonic conductivity, state dependent (:315)
        Dual<NV> ksum = 0.5 * (pl.kT + mi.kT);
        Dual<NV> kTf = ksum.v > 0.0 ? (pl.kT * mi.kT) / ksum : dconst<NV>(0.0);
        f[0] = P.Wp * (P.c_v_w * fw + P.c_v_o * fo);     // weighted-sum pressure equation (:343-346)
        f[1] = few + feo + area * (kTf * dT);            // :352
        f[2] = P.Wo * fo;
    }
}

template <int NF>
struct Fields {
    GField u[NF];
    GField phi, K[3], kT;
};

template <int NF>
__device__ __forceinline__ double gl(const GField& f, long long c, long long n, int np) {
    return c < 0 ? f.lo[c + np] : (c >= n ? f.hi[c - n] : f.v[c]);
}

template <int NF, int DIM, bool JAC>
__global__ void __launch_bounds__(128) assemble_kernel(Fields<NF> fl, const double* __restrict__ u_old, double idt,
                                                       Geom g, DevParams P, double* __restrict__ F,
                                                       double* __restrict__ J) {
    const long long n = g.n;
    long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= n) return;
    const int nx = g.nx, ny = g.ny, np = g.np;
    int i = (int)(cell % nx);
    long long t = cell / nx;
    int j = (int)(t % ny);
    int k = (int)(t / ny);

    double uc[NF];
#pragma unroll
    for (int f = 0; f < NF; f++) uc[f] = fl.u[f].v[cell];
    double phi = fl.phi.v[cell];
    double kTs = (NF == 2) ? fl.kT.v[cell] : 0.0;
    CellProps<NF> me = cell_props<NF>(P, uc, phi, kTs);

    double R[NF];
    double D[NF][NF];
#pragma unroll
    for (int r = 0; r < NF; r++) {
        R[r] = 0.0;
#pragma unroll
        for (int c = 0; c < NF; c++) D[r][c] = 0.0;
    }

    // ---- accumulation (cell integrals) ------------------------------------------------------
    {
        double po = u_old[cell], To = u_old[n + cell];
        double w = g.vol * idt;
        if constexpr (NF == 2) {
            double ro_old = oil_rho_v(P, po, To);
            Dual<NF> am = (w * phi) * (me.rho_o - ro_old);                       // singlephase.py:120
            Dual<NF> ae = (w * phi * P.c_v_o) * (me.rho_o * me.T - ro_old * To)  // :123
                          + (w * (1.0 - phi) * P.rho_r * P.c_r) * (me.T - To);
            R[0] += am.v;
            R[1] += ae.v;
#pragma unroll
            for (int c = 0; c < NF; c++) {
                D[0][c] += am.d[c];
                D[1][c] += ae.d[c];
            }
        } else {
            double So = u_old[2 * n + cell];
            double ro_old = oil_rho_v(P, po, To), rw_old = water_rho_v(po, To);
            Dual<NF> Sw = 1.0 - me.S;
            Dual<NF> aw = (w * phi) * (me.rho_w * Sw - rw_old * (1.0 - So));     // twophase.py:333
            Dual<NF> ao = (w * phi) * (me.rho_o * me.S - ro_old * So);          // :337
            Dual<NF> ae = (w * phi * P.c_v_w) * (me.rho_w * Sw * me.T - rw_old * (1.0 - So) * To) +
                          (w * phi * P.c_v_o) * (me.rho_o * me.S * me.T - ro_old * So * To) +
                          (w * (1.0 - phi) * P.rho_r * P.c_r) * (me.T - To);    // :349
            Dual<NF> ap = P.Wp * (P.c_v_w * aw + P.c_v_o * ao);                 // :344
            Dual<NF> as = P.Wo * ao;
            R[0] += ap.v;
            R[1] += ae.v;
            R[2] += as.v;
#pragma unroll
            for (int c = 0; c < NF; c++) {
                D[0][c] += ap.d[c];
                D[1][c] += ae.d[c];
                D[2][c] += as.d[c];
            }
        }
    }

    // ---- facet integrals ------------------------------------------------------------------------
    constexpr int NV = 2 * NF;
    CellProps<NV> me_pl, me_mi;
    embed_props<NF, 0>(me, me_pl);
    embed_props<NF, NF>(me, me_mi);

#pragma unroll
    for (int s = 1; s < 2 * DIM + 1; s++) {
        const int axis = (s - 1) >> 1;
        const bool hi_side = ((s - 1) & 1) != 0;   // neighbour has the higher index => this cell is '+'
        bool exists;
        long long nb;
        if (axis == 0) {
            exists = hi_side ? (i < nx - 1) : (i > 0);
            nb = cell + (hi_side ? 1 : -1);
        } else if (axis == 1) {
            if (DIM == 2)
                exists = hi_side ? (j < ny - 1 || g.has_hi) : (j > 0 || g.has_lo);
            else
                exists = hi_side ? (j < ny - 1) : (j > 0);
            nb = cell + (hi_side ? nx : -nx);
        } else {
            exists = hi_side ? (k < g.nz - 1 || g.has_hi) : (k > 0 || g.has_lo);
            nb = cell + (hi_side ? (long long)nx * ny : -(long long)nx * ny);
        }
        double Oblk[NF][NF];
#pragma unroll
        for (int r = 0; r < NF; r++)
#pragma unroll
            for (int c = 0; c < NF; c++) Oblk[r][c] = 0.0;

        if (exists) {
            double un[NF];
#pragma unroll
            for (int f = 0; f < NF; f++) un[f] = gl<NF>(fl.u[f], nb, n, np);
            double phin = gl<NF>(fl.phi, nb, n, np);
            double kTn = (NF == 2) ? gl<NF>(fl.kT, nb, n, np) : 0.0;
            double Kn = gl<NF>(fl.K[axis], nb, n, np);
            double Kf = harm(fl.K[axis].v[cell], Kn);
            CellProps<NF> other = cell_props<NF>(P, un, phin, kTn);
            double grav = (axis == 2) ? P.g : 0.0;
            Dual<NV> f[NF];
            if (hi_side) {
                CellProps<NV> ot;
                embed_props<NF, NF>(other, ot);
                face_flux<NF>(P, me_pl, ot, Kf, g.area[axis], 1.0 / g.h[axis], grav, f);
#pragma unroll
                for (int r = 0; r < NF; r++) {
                    R[r] += f[r].v;
#pragma unroll
                    for (int c = 0; c < NF; c++) {
                        D[r][c] += f[r].d[c];
                        Oblk[r][c] = f[r].d[NF + c];
                    }
                }
            } else {
                CellProps<NV> ot;
                embed_props<NF, 0>(other, ot);
                face_flux<NF>(P, ot, me_mi, Kf, g.area[axis], 1.0 / g.h[axis], grav, f);
#pragma unroll
                for (int r = 0; r < NF; r++) {
                    R[r] -= f[r].v;
#pragma unroll
                    for (int c = 0; c < NF; c++) {
                        D[r][c] -= f[r].d[NF + c];
                        Oblk[r][c] = -f[r].d[c];
                    }
                }
            }
        }
        if (JAC) {
#pragma unroll
            for (int r = 0; r < NF; r++)
#pragma unroll
                for (int c = 0; c < NF; c++) J[((long long)(s * NF + r) * NF + c) * n + cell] = Oblk[r][c];
        }
    }

#pragma unroll
    for (int r = 0; r < NF; r++) F[(long long)r * n + cell] = R[r];
    if (JAC) {
#pragma unroll
        for (int r = 0; r < NF; r++)
#pragma unroll
            for (int c = 0; c < NF; c++) J[((long long)r * NF + c) * n + cell] = D[r][c];
    }
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

template <int NF, bool JAC>
__global__ void sources_kernel(int ncells, const int64_t* __restrict__ cells, const int* __restrict__ off,
                               const tpb_source* __restrict__ ent, const double* __restrict__ u,
                               const double* __restrict__ Kx, const double* __restrict__ Ky, long long n, DevParams P,
                               double* __restrict__ F, double* __restrict__ J) {
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
#pragma unroll
    for (int r = 0; r < NF; r++) {
        F[(long long)r * n + cell] += acc[r].v;
        if (JAC) {
#pragma unroll
            for (int c = 0; c < NF; c++) J[((long long)r * NF + c) * n + cell] += acc[r].d[c];
        }
    }
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
    const int threads = 128;
    const unsigned blocks = (unsigned)((n + threads - 1) / threads);
    if (J)
        assemble_kernel<NF, DIM, true><<<blocks, threads, 0, h->stream>>>(fl, u_old, 1.0 / dt, h->g, h->dp, F, J);
    else
        assemble_kernel<NF, DIM, false><<<blocks, threads, 0, h->stream>>>(fl, u_old, 1.0 / dt, h->g, h->dp, F, J);
    h->launches++;
    if (h->nsrc_cells > 0) {
        const unsigned sb = (unsigned)((h->nsrc_cells + 127) / 128);
        if (J)
            sources_kernel<NF, true><<<sb, 128, 0, h->stream>>>(h->nsrc_cells, h->src_cell, h->src_off, h->src_ent, u,
                                                               h->fld[TPB_KX], h->fld[TPB_KY], n, h->dp, F, J);
        else
            sources_kernel<NF, false><<<sb, 128, 0, h->stream>>>(h->nsrc_cells, h->src_cell, h->src_off, h->src_ent, u,
                                                                h->fld[TPB_KX], h->fld[TPB_KY], n, h->dp, F, J);
        h->launches++;
    }
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

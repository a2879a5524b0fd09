// tpb_pc.cu - K3-K8: the two-stage preconditioner.
//
// Stage 1 (reference: CPRStage1PC preconditioners.py:335-906, CPTRStage1PC :1243-1571, PCFIELDSPLIT
// schur FULL singlephase.py:309-319 / twophase.py:536-545, ConvDiffSchur(TwoPhases)PC :11-333):
//   K3  block extraction straight from the block-stencil Jacobian (no re-assembly)
//   K5  QI / TI / QI_temp / TI_temp decoupling   Atilde_pp = A_pp - D_ps D_ss^-1 A_sp
//   K6  restriction r_p = x_p - D_ps D_ss^-1 x_s, prolongation y = [y_p; 0]
//   K4  ConvDiff temperature operator with coefficients frozen at the Newton state
//   K7  scalar-stencil multigrid V-cycle in the role of hypre BoomerAMG (one V-cycle per apply):
//       piecewise-constant aggregation with per-level semi-coarsening (an axis is coarsened only
//       where its mean coupling is strong), Galerkin coarse operators (stay 5|7-point) with the
//       couplings along coarsened axes scaled by mg_coarse_scale (cell-centred multigrid: the plain
//       Galerkin operator of constant transfers is 2x too stiff per coarsened axis); smoothing by
//       zebra z-line Gauss-Seidel in 3-D (every column solved exactly from factors computed at
//       set-up, z never coarsened - thin layers, Dz << Dx) or red-black point Gauss-Seidel (2-D,
//       mg_smoother 0); the coarsest levels (<= TAIL_CELLS cells) run inside one CTA.
// Stage 2 (reference: PETSc bjacobi + ilu(0), singlephase.py:348-349, twophase.py:547-548):
//   K8  block ILU(0) of the nf x nf block stencil in red-black ordering: for a 5|7-point stencil
//       ILU(0) only modifies the diagonal blocks, D_b = A_bb - sum_r A_br A_rr^-1 A_rb, and both
//       triangular solves are two fully parallel colour sweeps.
// PCCOMPOSITE multiplicative: y = B1 x ; y += B2 (x - J y).
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "tpb_internal.cuh"

namespace {

constexpr int MAXLEV = 40;

struct MgLevel {
    int nx = 0, ny = 0, nz = 0;
    long long n = 0;
    long long cap = 0;    // allocated cells (levels are re-shaped in place while they fit)
    int cx = 1, cy = 1, cz = 1;
    double* a = nullptr;  // ns * n
    bool own_a = false;
    double* x = nullptr;  // the vectors the cycle works on: the level's own storage, or (gather level of a
    double* b = nullptr;  // multi-rank hierarchy) this rank's section of the gathered level
    double* x_own = nullptr;
    double* b_own = nullptr;
    bool line = false;    // smoothed by the hybrid zebra z-line Gauss-Seidel
    double* fac = nullptr;   // 6 * cap: Thomas factors of every column (1/pivot | lower/pivot | upper/pivot), then the
    long long fac_cap = 0;   // iterate between pre- and post-smoothing and two work vectors (xt, xs0, xs1 below)
    double* xt() const { return fac + 3 * fac_cap; }
    double* xs0() const { return fac + 4 * fac_cap; }
    double* xs1() const { return fac + 5 * fac_cap; }
};

// Multi-rank slabs (K7 over NCCL): every rank coarsens its own slab with globally agreed coarsening factors
// until the levels of all ranks together hold <= GATHER_CELLS cells.  That level is all-gathered (operator at
// set-up, right-hand side once per V-cycle) - slabs are cut along the slowest axis, so the concatenation of the
// ranks' arrays IS the global level, and the couplings across slab faces that the Galerkin sums carried down in
// the boundary slots become interior couplings there.  Every rank then runs the remaining levels (`glob`)
// redundantly on identical data (bitwise identical results, no second exchange) and prolongs from its section.
// The levels above the gather level smooth with the couplings across slab faces dropped (block-Jacobi between
// slabs, Gauss-Seidel inside - what hypre's hybrid smoother does between processes).
struct MgHier {
    int nlev = 0;
    MgLevel lev[MAXLEV];
    MgHier* glob = nullptr;
    double* ga = nullptr;          // gathered operator of the gather level (ns * gn)
    long long ga_cap = 0;
    std::vector<long long> gcnt, goff;   // cells of every rank on the gather level and their offsets
    int last_sweeps = 0;           // > 0: the last level is diagonally dominant (mg_dd_stop) and gets this many sweeps
    bool skip_glob = false;        // multi-rank hierarchy that stopped on diagonal dominance: no gather level this set-up
};

// mg_dd_stop on multi-rank slabs (TPB_MG_DD_DIST=0 switches it off): the row ratio is max-reduced over the ranks, so
// all of them stop at the same level, and a hierarchy that stopped this way has no gather level - its last level is
// smoothed slab by slab like the levels above it
inline bool dd_on_slabs() {
    static const bool v = !(getenv("TPB_MG_DD_DIST") && atoi(getenv("TPB_MG_DD_DIST")) == 0);
    return v;
}

// sweeps for a 1e-3 contraction on a level whose rows have sum|off-diag| <= rho |diag| (same rule in oracle/cport)
inline int dd_sweeps(double rho, int cap) {
    int k = 1;
    if (rho > 0.0 && rho < 1.0) k = (int)ceil(log(1e-3) / log(rho));
    if (rho >= 1.0) k = cap;   // (only reachable with mg_dd_stop >= 1)
    if (k < 1) k = 1;
    if (cap > 0 && k > cap) k = cap;
    return k;
}

}  // namespace

struct PcState {
    const double* J = nullptr;
    double* w[TPB_MAXF] = {nullptr, nullptr, nullptr};
    double* App = nullptr;
    double* A00 = nullptr;
    double* AT = nullptr;
    MgHier mg_p, mg_T;
    double* Dinv = nullptr;
    float *Dc = nullptr, *Lc = nullptr;   // colour-separated fp32 copies of the ILU factor (tpb_pc.cu: ilu_setup_kernel)
    double *t0 = nullptr, *t1 = nullptr, *t2 = nullptr, *t3 = nullptr;
    double* strength = nullptr;  // 3 doubles (device)
    bool ready = false;
    // the whole PC apply (~100 small dependent kernels) replayed from CUDA graphs on fixed in/out buffers: one
    // graph on a single slab; on slabs with neighbours a short program of graphs with the NCCL steps (all-gather
    // of each V-cycle's gather level, halo exchange of the residual SpMV) between them
    struct Item {
        cudaGraphExec_t g = nullptr;
        int64_t nodes = 0;
        std::function<void()> op;
    };
    std::vector<Item> prog;
    std::vector<cudaGraphExec_t> spare;   // executables of the previous set-up, re-used through cudaGraphExecUpdate
    size_t spare_next = 0;
    bool capturing = false;
    int64_t cap_l0 = 0;
    double *gx = nullptr, *gy = nullptr;
    bool graph_ok = false;
    std::vector<long long> graph_sig;   // what the captured application depends on (level shapes, buffers, sweep counts)
    // programs of earlier set-ups, by signature: the coarsening schedule flips between a few variants from one Newton
    // iteration to the next, and an executable graph stays valid as long as its signature comes back
    struct Cached {
        std::vector<long long> sig;
        std::vector<Item> prog;
    };
    std::vector<Cached> cache;
};

namespace {

__host__ __device__ __forceinline__ int opp_slot(int s) { return s == 0 ? 0 : (((s - 1) ^ 1) + 1); }

// neighbour of cell (i,j,k) through stencil slot s on an (nx,ny,nz) box; -1 when outside
__device__ __forceinline__ long long nbr_cell(int nx, int ny, int nz, int i, int j, int k, long long c, int s) {
    switch (s) {
        case 0: return c;
        case 1: return i > 0 ? c - 1 : -1;
        case 2: return i < nx - 1 ? c + 1 : -1;
        case 3: return j > 0 ? c - nx : -1;
        case 4: return j < ny - 1 ? c + nx : -1;
        case 5: return k > 0 ? c - (long long)nx * ny : -1;
        default: return k < nz - 1 ? c + (long long)nx * ny : -1;
    }
}

__device__ __forceinline__ void inv_block(int m, const double* A, double* Ai) {
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

#define JAT(s, r, q, c) J[((long long)((s) * NF + (r)) * NF + (q)) * n + (c)]

// column sum of block (r,q) over every row that points at column cell c (TI decoupling:
// transpose + getRowSum, preconditioners.py:693-703)
template <int NF, int DIM>
__device__ __forceinline__ double colsum(const double* __restrict__ J, long long n, int nx, int ny, int nz, int i,
                                         int j, int k, long long c, int r, int q) {
    double sum = JAT(0, r, q, c);
#pragma unroll
    for (int s = 1; s < 2 * DIM + 1; s++) {
        long long nb = nbr_cell(nx, ny, nz, i, j, k, c, s);
        if (nb >= 0) sum += JAT(opp_slot(s), r, q, nb);
    }
    return sum;
}

// ---- K3 + K5 for CPR: weights w[f] and the decoupled pressure stencil ------------------------
template <int NF, int DIM>
__global__ void __launch_bounds__(128) cpr_setup_kernel(const double* __restrict__ J, Geom g, int dec, double* w1,
                                                        double* w2, double* __restrict__ App) {
    const long long n = g.n;
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const int nx = g.nx, ny = g.ny, nz = g.nz;
    int i, j, k;
    tpb_ijk(c, nx, ny, i, j, k);
    constexpr int L = NF - 1;
    double wf[TPB_MAXF] = {0.0, 0.0, 0.0};
    if (dec == TPB_DECOUP_QI) {
        wf[L] = JAT(0, 0, L, c) / JAT(0, L, L, c);
    } else if (dec == TPB_DECOUP_TI) {
        wf[L] = colsum<NF, DIM>(J, n, nx, ny, nz, i, j, k, c, 0, L) / colsum<NF, DIM>(J, n, nx, ny, nz, i, j, k, c, L, L);
    } else if (NF == 3 && (dec == TPB_DECOUP_QI_TEMP || dec == TPB_DECOUP_TI_TEMP)) {
        double B[4], Bi[4], pT, pS;
        if (dec == TPB_DECOUP_TI_TEMP) {
            B[0] = colsum<NF, DIM>(J, n, nx, ny, nz, i, j, k, c, 1, 1);
            B[1] = colsum<NF, DIM>(J, n, nx, ny, nz, i, j, k, c, 1, NF - 1);
            B[2] = colsum<NF, DIM>(J, n, nx, ny, nz, i, j, k, c, NF - 1, 1);
            B[3] = colsum<NF, DIM>(J, n, nx, ny, nz, i, j, k, c, NF - 1, NF - 1);
            pT = colsum<NF, DIM>(J, n, nx, ny, nz, i, j, k, c, 0, 1);
            pS = colsum<NF, DIM>(J, n, nx, ny, nz, i, j, k, c, 0, NF - 1);
        } else {
            B[0] = JAT(0, 1, 1, c);
            B[1] = JAT(0, 1, NF - 1, c);
            B[2] = JAT(0, NF - 1, 1, c);
            B[3] = JAT(0, NF - 1, NF - 1, c);
            pT = JAT(0, 0, 1, c);
            pS = JAT(0, 0, NF - 1, c);
        }
        inv_block(2, B, Bi);
        wf[1] = pT * Bi[0] + pS * Bi[2];
        wf[NF - 1] = pT * Bi[1] + pS * Bi[3];
    }
    w1[c] = wf[1];
    if (NF == 3) w2[c] = wf[2];
#pragma unroll
    for (int s = 0; s < 2 * DIM + 1; s++) {
        double v = JAT(s, 0, 0, c);
#pragma unroll
        for (int f = 1; f < NF; f++) v -= wf[f] * JAT(s, f, 0, c);
        App[(long long)s * n + c] = v;
    }
}

// ---- K3 + K5 for CPTR / single-phase field-split: 2x2 primary block --------------------------
template <int NF, int DIM>
__global__ void __launch_bounds__(128) cptr_setup_kernel(const double* __restrict__ J, Geom g, int dec, int a11,
                                                         double* w0, double* w1, double* __restrict__ A00,
                                                         double* __restrict__ App, double* __restrict__ AT) {
    const long long n = g.n;
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const int nx = g.nx, ny = g.ny, nz = g.nz;
    int i, j, k;
    tpb_ijk(c, nx, ny, i, j, k);
    double wa[2] = {0.0, 0.0};
    if (NF == 3 && dec == TPB_DECOUP_QI) {
        double dss = JAT(0, NF - 1, NF - 1, c);
        wa[0] = JAT(0, 0, NF - 1, c) / dss;
        wa[1] = JAT(0, 1, NF - 1, c) / dss;
    } else if (NF == 3 && dec == TPB_DECOUP_TI) {
        double dss = colsum<NF, DIM>(J, n, nx, ny, nz, i, j, k, c, NF - 1, NF - 1);
        wa[0] = colsum<NF, DIM>(J, n, nx, ny, nz, i, j, k, c, 0, NF - 1) / dss;
        wa[1] = colsum<NF, DIM>(J, n, nx, ny, nz, i, j, k, c, 1, NF - 1) / dss;
    }
    w0[c] = wa[0];
    w1[c] = wa[1];
#pragma unroll
    for (int s = 0; s < 2 * DIM + 1; s++)
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int b = 0; b < 2; b++) {
                double v = JAT(s, a, b, c);
                if (NF == 3) v -= wa[a] * JAT(s, NF - 1, b, c);
                A00[((long long)(s * 2 + a) * 2 + b) * n + c] = v;
                if (a == 0 && b == 0) App[(long long)s * n + c] = v;
                if (a == 1 && b == 1 && a11) AT[(long long)s * n + c] = v;
            }
}

// ---- K4: ConvDiff temperature operator (preconditioners.py:63-108, 225-276) ------------------
struct CdCell {
    double p, T, ro, rw, lo, lw, kT;
};
// inputs of the ConvDiff operator with the slab's ghost planes (state ghosts are the ones the assembly of the
// same Newton state exchanged)
struct CdIn {
    GField u[TPB_MAXF];
    GField phi, K[3], kT;
};
__device__ __forceinline__ double cd_gl(const GField& f, long long c, long long n, int np) {
    return c < 0 ? f.lo[c + np] : (c >= n ? f.hi[c - n] : f.v[c]);
}
template <int NF>
__device__ __forceinline__ CdCell cd_props(const DevParams& P, const CdIn& in, long long n, int np, long long c) {
    CdCell q;
    q.p = cd_gl(in.u[0], c, n, np);
    q.T = cd_gl(in.u[1], c, n, np);
    double a, b, imo, imo_T;
    oil_rho_d(P, q.p, q.T, q.ro, a, b);
    oil_imu_d(P, q.T, imo, imo_T);
    if (NF == 3) {
        double S = cd_gl(in.u[2], c, n, np), imw, imw_T;
        double phi = cd_gl(in.phi, c, n, np);
        water_rho_d(q.p, q.T, q.rw, a, b);
        water_imu_d(q.T, imw, imw_T);
        q.lo = S * q.ro * imo;
        q.lw = (1.0 - S) * q.rw * imw;
        q.kT = phi * (S * P.ko + (1.0 - S) * P.kw) + (1.0 - phi) * P.kr;
    } else {
        q.rw = 0.0;
        q.lw = 0.0;
        q.lo = q.ro * imo;
        q.kT = cd_gl(in.kT, c, n, np);
    }
    return q;
}
__device__ __forceinline__ double harm_d(double a, double b) {
    double s = 0.5 * (a + b);
    return s > 0.0 ? a * b / s : 0.0;
}

template <int NF, int DIM>
__global__ void __launch_bounds__(128) convdiff_kernel(CdIn in, double idt, Geom g, DevParams P, double* __restrict__ A) {
    const long long n = g.n;
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const int nx = g.nx, ny = g.ny, nz = g.nz, np = g.np;
    int i, j, k;
    tpb_ijk(c, nx, ny, i, j, k);
    double phi = in.phi.v[c];
    CdCell me = cd_props<NF>(P, in, n, np, c);
    double diag;
    if (NF == 2)
        diag = g.vol * idt * (phi * P.c_v_o * me.ro + (1.0 - phi) * P.rho_r * P.c_r);
    else {
        double S = in.u[2].v[c];
        diag = g.vol * idt * (phi * P.c_v_o * S * me.ro + phi * P.c_v_w * (1.0 - S) * me.rw + (1.0 - phi) * P.rho_r * P.c_r);
    }
#pragma unroll
    for (int s = 1; s < 2 * DIM + 1; s++) {
        const int axis = (s - 1) >> 1;
        const bool hi = ((s - 1) & 1) != 0;
        // neighbour through slot s; across a slab face it lives in a ghost plane (index < 0 or >= n)
        bool ex;
        long long nb;
        if (axis == 0) {
            ex = hi ? (i < nx - 1) : (i > 0);
            nb = c + (hi ? 1 : -1);
        } else if (axis == 1) {
            ex = hi ? (j < ny - 1 || (DIM == 2 && g.has_hi)) : (j > 0 || (DIM == 2 && g.has_lo));
            nb = c + (hi ? nx : -nx);
        } else {
            ex = hi ? (k < nz - 1 || g.has_hi) : (k > 0 || g.has_lo);
            nb = c + (hi ? (long long)np : -(long long)np);
        }
        double off = 0.0;
        if (ex) {
            const GField& Kax = in.K[axis];
            CdCell ot = cd_props<NF>(P, in, n, np, nb);
            const CdCell& pl = hi ? me : ot;
            const CdCell& mi = hi ? ot : me;
            double Kf = harm_d(Kax.v[c], cd_gl(Kax, nb, n, np));
            double grav = axis == 2 ? P.g : 0.0;
            double ih = 1.0 / g.h[axis], area = g.area[axis];
            double dp = ih * (pl.p - mi.p);
            double sgn = hi ? 1.0 : -1.0;
            double flo = dp - 0.5 * grav * (pl.ro + mi.ro);
            bool upo = flo > 0.0;
            double co = area * Kf * P.c_v_o * (upo ? pl.lo : mi.lo) * flo;
            if (hi ? upo : !upo) diag += sgn * co; else off += sgn * co;
            if (NF == 3) {
                double flw = dp - 0.5 * grav * (pl.rw + mi.rw);
                bool upw = flw > 0.0;
                double cw = area * Kf * P.c_v_w * (upw ? pl.lw : mi.lw) * flw;
                if (hi ? upw : !upw) diag += sgn * cw; else off += sgn * cw;
            }
            double d = area * harm_d(pl.kT, mi.kT) * ih;
            diag += d;
            off -= d;
        }
        A[(long long)s * n + c] = off;
    }
    A[c] = diag;
}

// producers and heaters add to the diagonal of the ConvDiff operator (preconditioners.py:91-108, 257-276);
// the coefficient of T in the producers' energy sink is -w (sum rho q c_v) = (energy source term)/T
__device__ __forceinline__ double peaceman_wi_d(double Kx, double Ky) {
    const double hh = 5.0, rw = 0.1, Dx = 5.0, Dy = 5.0;
    double a = Ky / Kx, b = Kx / Ky;
    double ro = 0.28 * sqrt(sqrt(a) * Dx * Dx + sqrt(b) * Dy * Dy) / (sqrt(sqrt(a)) + sqrt(sqrt(b)));
    return 2.0 * 3.141592653589793 * hh * sqrt(Kx * Ky) / log(ro / rw);
}
__device__ __forceinline__ double rate_value(const tpb_source& s, double wi, double imu, double p) {
    if (s.const_rate) return s.max_rate;
    double d = s.bhp - p;
    double dd = s.max_rate < 0.0 ? (d >= 0.0 ? 0.0 : d) : (d <= 0.0 ? 0.0 : d);
    double rate = wi * imu * dd;
    return (fabs(rate) - fabs(s.max_rate) >= 0.0) ? s.max_rate : rate;
}
template <int NF>
__global__ void convdiff_sources_kernel(int ncells, const int64_t* __restrict__ cells, const int* __restrict__ off,
                                        const tpb_source* __restrict__ ent, const double* __restrict__ u,
                                        const double* __restrict__ Kx, const double* __restrict__ Ky, long long n,
                                        DevParams P, double* __restrict__ A) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncells) return;
    long long c = cells[t];
    double p = u[c], T = u[n + c];
    double add = 0.0;
    for (int e = off[t]; e < off[t + 1]; e++) {
        tpb_source s = ent[e];
        if (s.kind == TPB_HEATER) {
            add += s.weight * P.U;
        } else if (s.kind == TPB_PROD) {
            double wi = s.const_rate ? 0.0 : peaceman_wi_d(Kx[c], Ky[c]);
            double ro, a, b, imo, imo_T;
            oil_rho_d(P, p, T, ro, a, b);
            oil_imu_d(P, T, imo, imo_T);
            if (NF == 2) {
                double q = rate_value(s, wi, imo, p);
                add -= s.weight * ro * q * P.c_v_o;
            } else {
                double S = u[2 * n + c], rw, imw, imw_T;
                water_rho_d(p, T, rw, a, b);
                water_imu_d(T, imw, imw_T);
                double mob_o = S * imo, mob_w = (1.0 - S) * imw, imu = mob_o + mob_w;
                double q = rate_value(s, wi, imu, p);
                double mu = 1.0 / imu;
                double qw = mob_w * mu * q, qo = mob_o * mu * q;
                add -= s.weight * (rw * qw * P.c_v_w + ro * qo * P.c_v_o);
            }
        }
    }
    A[c] += add;
}

// ---- K7 multigrid kernels ---------------------------------------------------------------------
struct LevGeom {
    int nx, ny, nz, cx, cy, cz;   // cx,cy,cz in {1,2}: aggregate index = fine index >> (c - 1)
    long long n;
    int line;                     // z-line smoothed level: colours are (i+j)&1 columns instead of (i+j+k)&1 cells
};

// clamped neighbour for branch-free stencil sweeps: index of the neighbour through slot s, or the cell itself
// when the neighbour is outside (the caller discards that product with a select, so loads stay unconditional
// and all of a thread's loads can be in flight together)
__device__ __forceinline__ long long nbr_clamped(int nx, int ny, int nz, int i, int j, int k, long long c, int s,
                                                 bool& exists) {
    switch (s) {
        case 1: exists = i > 0; return exists ? c - 1 : c;
        case 2: exists = i < nx - 1; return exists ? c + 1 : c;
        case 3: exists = j > 0; return exists ? c - nx : c;
        case 4: exists = j < ny - 1; return exists ? c + nx : c;
        case 5: exists = k > 0; return exists ? c - (long long)nx * ny : c;
        case 6: exists = k < nz - 1; return exists ? c + (long long)nx * ny : c;
        default: exists = true; return c;
    }
}

// sum over cells of |a[2ax+1]| + |a[2ax+2]| for the three axes -> out[0..2] (atomics; a few thousand adds), and the
// strongest row of the level, max over cells of sum|off-diagonals| / |diagonal| -> out[3] (mg_dd_stop; summed in
// slot order and divided exactly as oracle/cport does, a maximum has no order: the two agree to the bit)
template <int NS>
__global__ void __launch_bounds__(256) strength_kernel(const double* __restrict__ a, long long n, double* out) {
    double acc[3] = {0.0, 0.0, 0.0};
    double rho = 0.0;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
        double v[NS];
        double sum = 0.0;
#pragma unroll
        for (int s = 1; s < NS; s++) {
            v[s] = fabs(a[(long long)s * n + c]);
            sum += v[s];
        }
#pragma unroll
        for (int ax = 0; ax < (NS - 1) / 2; ax++) acc[ax] += v[2 * ax + 1] + v[2 * ax + 2];
        const double d = fabs(a[c]);
        const double r = d > 0.0 ? sum / d : (sum > 0.0 ? 1e300 : 0.0);
        rho = fmax(rho, r);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rho = fmax(rho, __shfl_down_sync(0xffffffffu, rho, o));
    // non-negative doubles order like their bit patterns
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned long long*>(out + 3), (unsigned long long)__double_as_longlong(rho));
    __shared__ double sm[3][8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int ax = 0; ax < 3; ax++) {
        double v = acc[ax];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) sm[ax][wid] = v;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double v = 0.0;
        for (int w = 0; w < 8; w++) v += sm[threadIdx.x][w];
        atomicAdd(&out[threadIdx.x], v);
    }
}

// Row repair of a scalar operator before it is handed to the multigrid.  Newton iterates may leave the physical
// range for a few cells (S_o < 0 at an injector, S_o > 1: basic line search, thermalmodel.py:165), which gives
// the pressure row of those cells negative mobilities: positive off-diagonals and a diagonal that is negative or
// far from dominant.  Gauss-Seidel amplifies on such rows and the V-cycle stops being a contraction (measured: a
// 200-iteration FGMRES stall with 5 such rows out of 4.5 M).  A diagonal below 0.8 x the sum of the row's
// |couplings| is raised to that sum; every other row is left bit-for-bit alone, and only the preconditioner's copy
// of the operator is touched.
template <int NS>
__global__ void __launch_bounds__(256) row_repair_kernel(double* __restrict__ a, long long n) {
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    double sum = 0.0;
#pragma unroll
    for (int s = 1; s < NS; s++) sum += fabs(a[(long long)s * n + c]);
    if (a[c] < 0.8 * sum) a[c] = sum;
}

// Galerkin coarse operator for piecewise-constant aggregates (stays a 5|7-point stencil)
template <int NS>
__global__ void __launch_bounds__(128) coarsen_op_kernel(const double* __restrict__ af, LevGeom f, LevGeom cg,
                                                         double sc, double* __restrict__ ac) {
    long long C = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (C >= cg.n) return;
    int I, Jc, Kc;
    tpb_ijk(C, cg.nx, cg.ny, I, Jc, Kc);
    double acc[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) acc[s] = 0.0;
    for (int dk = 0; dk < f.cz; dk++)
        for (int dj = 0; dj < f.cy; dj++)
            for (int di = 0; di < f.cx; di++) {
                int i = I * f.cx + di, j = Jc * f.cy + dj, k = Kc * f.cz + dk;
                if (i >= f.nx || j >= f.ny || k >= f.nz) continue;
                long long fc = i + (long long)f.nx * (j + (long long)f.ny * k);
                acc[0] += af[fc];
#pragma unroll
                for (int s = 1; s < NS; s++) {
                    const int axis = (s - 1) >> 1;
                    const bool hi = ((s - 1) & 1) != 0;
                    int pos = axis == 0 ? i : (axis == 1 ? j : k);
                    int cf = axis == 0 ? f.cx : (axis == 1 ? f.cy : f.cz);
                    int npos = pos + (hi ? 1 : -1);
                    bool same = (npos >= 0) && (npos / cf == pos / cf);
                    double v = af[(long long)s * f.n + fc];
                    if (same) acc[0] += v; else acc[s] += v;
                }
            }
    // tpb_solver_opts.mg_coarse_scale: couplings along a coarsened axis are scaled, the row sum is kept
    if (sc != 1.0) {
#pragma unroll
        for (int s = 1; s < NS; s++) {
            const int axis = (s - 1) >> 1;
            const int cf = axis == 0 ? f.cx : (axis == 1 ? f.cy : f.cz);
            if (cf == 2) {
                const double nv = sc * acc[s];
                acc[0] += acc[s] - nv;
                acc[s] = nv;
            }
        }
    }
#pragma unroll
    for (int s = 0; s < NS; s++) ac[(long long)s * cg.n + C] = acc[s];
}

// ---- zebra z-line Gauss-Seidel (3-D levels) ------------------------------------------------------
// LU factors of every column's tridiagonal (diag, z-, z+), one thread per column marching up: id = 1/pivot,
// lf = lower * id, cp = upper * id.  The couplings through the bottom and the top of the box (slab faces on
// multi-rank runs) are not part of the column.  Same arithmetic as oracle/cport mg_line_factors.
__global__ void __launch_bounds__(128) line_factor_kernel(const double* __restrict__ a, LevGeom g, double* __restrict__ fac) {
    const long long np = (long long)g.nx * g.ny, n = g.n;
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= np) return;
    double cprev = 0.0;
#pragma unroll 4
    for (int k = 0; k < g.nz; k++) {
        const long long c = q + np * k;
        const double lo = k > 0 ? a[5 * n + c] : 0.0, up = k < g.nz - 1 ? a[6 * n + c] : 0.0;
        const double den = a[c] - lo * cprev;
        const double id = den != 0.0 ? 1.0 / den : 0.0;
        fac[c] = id;
        fac[n + c] = lo * id;
        cprev = up * id;
        fac[2 * n + c] = cprev;
    }
}

// Hybrid zebra z-line smoother.  The xy-plane is cut into tiles of tx x ty columns (8 x 3 while a tile holds at most
// 2048 cells, i.e. nz <= 85); ONE thread block keeps a tile - its x values with a one-column rim, and the right-hand
// sides, couplings and Thomas factors of its columns - in shared memory and does `nsw` complete zebra sweeps on it
// (columns coloured by the global (i+j)&1, colour 0 first) while the rim stays frozen at its values of entry:
// Gauss-Seidel inside a tile, block Jacobi between tiles, what hypre's hybrid smoother is between processes.
// A smoothing step is one launch per group of mg_tile_sweeps sweeps (default: one launch) instead of two launches per
// sweep, the operator is read once per launch, and everything between the first load and the last store runs out of
// shared memory:
//   load    couplings a1..a4 (zeroed where there is no neighbour), 1/pivot, lf, cp and b of the tile's cells; x of tile
//           + rim (PROLONG: x + omega * xc[aggregate], the coarse correction never exists in memory; xin == nullptr:
//           zero guess, no loads);
//   pass    (2 per sweep) the cells of the pass's colour: e = (b - sum a_s x[lateral]) / pivot; then one warp per
//           column of the colour: d_k = e_k - lf_k d_{k-1} up, x_k = d_k - cp_k x_{k+1} down, each as ONE warp scan
//           over affine maps (a lane composes its 3 consecutive planes first);
//   store   x of the tile -> xout (out of place: other tiles read xin concurrently).
// Same arithmetic as oracle/cport mg_line_smooth1 up to the association order of the scans.
// Measured (tools/bench_line.cu, B200): a block lives ~17 k cycles for one sweep, ~24 k for two, issue-bound (1.1 k
// instructions per warp at 35 % issue utilisation), not memory-bound; 60x220x85: 57 / 71 us per launch.
constexpr int LS_THREADS = 1024;
constexpr int LS_CPT = 2;       // a tile holds at most LS_CPT * LS_THREADS cells (bounds its shared memory)
__host__ __device__ inline void line_tile_shape(int nz, int& tx, int& ty) {
    const int menu[3][2] = {{8, 3}, {4, 2}, {1, 1}};
    const int cols_max = nz > 0 ? (LS_CPT * LS_THREADS) / nz : 1;
    for (int q = 0; q < 3; q++)
        if (menu[q][0] * menu[q][1] <= cols_max || q == 2) {
            tx = menu[q][0];
            ty = menu[q][1];
            return;
        }
}

// inclusive warp scan of the affine maps v -> M v + C (composition order: lower lanes first)
__device__ __forceinline__ void affine_scan(double& M, double& C, int lane) {
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const double Mp = __shfl_up_sync(0xffffffffu, M, off);
        const double Cp = __shfl_up_sync(0xffffffffu, C, off);
        if (lane >= off) {
            C = fma(M, Cp, C);
            M = M * Mp;
        }
    }
}

// one column by one warp: ev holds e on entry and d in between; x -> xcol.  MM consecutive levels per lane; the
// arrays are walked with strides `st` / `xst` doubles (odd: free of bank conflicts).
template <int MM>
__device__ __forceinline__ void line_column_solve(double* ev, const double* lv, const double* cv, double* xcol, int st,
                                                  int xst, int nz, int lane) {
    double m[MM], c[MM];
    double Mc = 1.0, Cc = 0.0;
#pragma unroll
    for (int u = 0; u < MM; u++) {                       // up: d_k = e_k - lf_k d_{k-1}
        const int k = lane * MM + u;
        m[u] = k < nz ? -lv[k * st] : 1.0;
        c[u] = k < nz ? ev[k * st] : 0.0;
        Cc = fma(m[u], Cc, c[u]);
        Mc = Mc * m[u];
    }
    affine_scan(Mc, Cc, lane);
    double d = __shfl_up_sync(0xffffffffu, Cc, 1);
    if (lane == 0) d = 0.0;
#pragma unroll
    for (int u = 0; u < MM; u++) {
        const int k = lane * MM + u;
        d = fma(m[u], d, c[u]);
        if (k < nz) ev[k * st] = d;
    }
    __syncwarp();
    Mc = 1.0;
    Cc = 0.0;
#pragma unroll
    for (int u = 0; u < MM; u++) {                       // down: x_k = d_k - cp_k x_{k+1}, walked as r = nz-1-k
        const int k = nz - 1 - (lane * MM + u);
        m[u] = k >= 0 ? -cv[k * st] : 1.0;
        c[u] = k >= 0 ? ev[k * st] : 0.0;
        Cc = fma(m[u], Cc, c[u]);
        Mc = Mc * m[u];
    }
    affine_scan(Mc, Cc, lane);
    double xv = __shfl_up_sync(0xffffffffu, Cc, 1);
    if (lane == 0) xv = 0.0;
#pragma unroll
    for (int u = 0; u < MM; u++) {
        const int k = nz - 1 - (lane * MM + u);
        xv = fma(m[u], xv, c[u]);
        if (k >= 0) xcol[k * xst] = xv;
    }
}
// any nz: 32 levels per round, the carry broadcast from the last lane (xcol has its own stride)
__device__ __forceinline__ void line_column_solve_rounds(double* ev, const double* lv, const double* cv, double* xcol,
                                                         int st, int xst, int nz, int lane) {
    double carry = 0.0;
    for (int k0 = 0; k0 < nz; k0 += 32) {
        const int k = k0 + lane;
        double M = k < nz ? -lv[k * st] : 1.0, C = k < nz ? ev[k * st] : 0.0;
        affine_scan(M, C, lane);
        const double d = fma(M, carry, C);
        if (k < nz) ev[k * st] = d;
        carry = __shfl_sync(0xffffffffu, d, 31);
    }
    __syncwarp();
    carry = 0.0;
    for (int r0 = 0; r0 < nz; r0 += 32) {
        const int k = nz - 1 - (r0 + lane);
        double M = k >= 0 ? -cv[k * st] : 1.0, C = k >= 0 ? ev[k * st] : 0.0;
        affine_scan(M, C, lane);
        const double xk = fma(M, carry, C);
        if (k >= 0) xcol[k * xst] = xk;
        carry = __shfl_sync(0xffffffffu, xk, 31);
    }
}

#ifdef TPB_LINE_TIMING   // tools/bench_line.cu: phase time stamps of block 0
__device__ long long g_line_clk[8];
#define LINE_STAMP(q) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_line_clk[q] = clock64(); } while (0)
#else
#define LINE_STAMP(q) do { } while (0)
#endif
// Shared layout: plane-major, [k][column] with the column count padded to an odd number - a warp walking the columns of
// one plane (load, right-hand-side and store phases) and a warp walking the planes of one column (solve phase) are
// both free of bank conflicts.
struct LineSmem {
    double *xs;                                    // [nz][HP]  iterate of tile + rim, HP = (TX+2)*(TY+2) | 1
    double *e, *lf, *cp, *a1, *a2, *a3, *a4, *id, *bb;   // [nz][CP] each, CP = TX*TY | 1
};
__host__ __device__ __forceinline__ int line_cp(int tx, int ty) { return (tx * ty) | 1; }
__host__ __device__ __forceinline__ int line_hp(int tx, int ty) { return ((tx + 2) * (ty + 2)) | 1; }
__host__ __device__ inline size_t line_smem_doubles(int nz, int tx, int ty) {
    return (size_t)nz * (line_hp(tx, ty) + 9 * line_cp(tx, ty));
}
__device__ __forceinline__ LineSmem line_smem_of(double* sm, int nz, int tx, int ty) {
    const size_t halo = (size_t)nz * line_hp(tx, ty), t = (size_t)nz * line_cp(tx, ty);
    double* q = sm + halo;
    return LineSmem{sm, q, q + t, q + 2 * t, q + 3 * t, q + 4 * t, q + 5 * t, q + 6 * t, q + 7 * t, q + 8 * t};
}

// All sweeps of one TX x TY tile by one thread block (tid, nth: thread index and count inside the block, nth a
// multiple of 32).  Nothing is kept in registers between the phases - couplings, factors and right-hand sides of the
// tile live in shared memory next to the iterate - so the kernel runs without spills at 1024 threads per block, and
// every phase is a short grid-stride loop over the tile's cells (the right-hand-side phase over the cells of the
// pass's colour only: all lanes busy).  TX is a power of two and both are compile-time constants, so decoding a cell
// costs shifts and a constant division; the rim of the tile is loaded by the threads of the tile's edge columns.
// 32-bit cell indices (a slab holds fewer than 2^31 cells).  Cells of the tile that lie outside the box get zero
// couplings and a zero 1/pivot: they stay 0 and need no tests in the passes.
template <bool PROLONG, int TX, int TY>
__device__ __forceinline__ void line_tile_smooth(const double* __restrict__ a, const double* __restrict__ fac,
                                                 const double* b, const double* xin, double* xout, const LevGeom& g,
                                                 const double* xc, int cnx, int cny, double omega, int nsw, int i0, int j0,
                                                 double* sm, int tid, int nth, bool wait_first) {
    constexpr int HX = TX + 2, CP = (TX * TY) | 1, HP = ((TX + 2) * (TY + 2)) | 1;
    constexpr int TXH = TX > 1 ? TX / 2 : 1;   // columns of one colour per tile row
    const int nz = g.nz, nx = g.nx, ny = g.ny;
    const int np = nx * ny;
    const long long n = g.n;
    const LineSmem S = line_smem_of(sm, nz, TX, TY);
    const int cells = TX * TY * nz;
    LINE_STAMP(0);
    // ---- set-up products of the tile (may be loaded before the predecessor kernel is done)
#pragma unroll 2
    for (int idx = tid; idx < cells; idx += nth) {
        const int ii = idx & (TX - 1), r = idx / TX, jj = r % TY, k = r / TY;
        const int i = i0 + ii, j = j0 + jj;
        const bool in = i < nx && j < ny;
        const int c = i + nx * j + np * k;
        const int o = k * CP + jj * TX + ii;
        S.a1[o] = (in && i > 0) ? a[n + c] : 0.0;
        S.a2[o] = (in && i < nx - 1) ? a[2 * n + c] : 0.0;
        S.a3[o] = (in && j > 0) ? a[3 * n + c] : 0.0;
        S.a4[o] = (in && j < ny - 1) ? a[4 * n + c] : 0.0;
        S.id[o] = in ? fac[c] : 0.0;
        S.lf[o] = in ? fac[n + c] : 0.0;
        S.cp[o] = in ? fac[2 * n + c] : 0.0;
    }
    LINE_STAMP(1);
    if (wait_first) pdl_wait();
    // ---- right-hand sides; x of the tile (own cells) and of its rim (edge columns' threads)
#pragma unroll 2
    for (int idx = tid; idx < cells; idx += nth) {
        const int ii = idx & (TX - 1), r = idx / TX, jj = r % TY, k = r / TY;
        const int i = i0 + ii, j = j0 + jj;
        const bool in = i < nx && j < ny;
        const int c = i + nx * j + np * k;
        const int so = k * HP + (jj + 1) * HX + (ii + 1);
        S.bb[k * CP + jj * TX + ii] = in ? b[c] : 0.0;
        auto xval = [&](int ci, int cj, int cc) -> double {   // input iterate at cell (ci, cj, k), 0 outside the box
            if (xin == nullptr || ci < 0 || ci >= nx || cj < 0 || cj >= ny) return 0.0;
            double v = xin[cc];
            if (PROLONG) v += omega * xc[(ci >> (g.cx - 1)) + cnx * ((cj >> (g.cy - 1)) + cny * k)];
            return v;
        };
        S.xs[so] = xval(i, j, c);
        if (ii == 0) S.xs[so - 1] = xval(i - 1, j, c - 1);
        if (ii == TX - 1) S.xs[so + 1] = xval(i + 1, j, c + 1);
        if (jj == 0) S.xs[so - HX] = xval(i, j - 1, c - nx);
        if (jj == TY - 1) S.xs[so + HX] = xval(i, j + 1, c + nx);
    }
    __syncthreads();
    LINE_STAMP(2);
    const int lane = tid & 31, wid = tid >> 5, nw = nth >> 5;
    const int mm = (nz + 31) / 32;
    const int par0 = (i0 + j0) & 1;
    for (int sw = 0; sw < nsw; sw++) {
        for (int col = 0; col < 2; col++) {
            // cells of this colour: ih-th column of the colour in tile row jj
            for (int idx = tid; idx < TXH * TY * nz; idx += nth) {
                const int ih = idx & (TXH - 1), r = idx / TXH, jj = r % TY, k = r / TY;
                const int ii = TX > 1 ? 2 * ih + ((par0 + jj + col) & 1) : 0;
                if (TX == 1 && ((par0 + jj) & 1) != col) continue;
                const int o = k * CP + jj * TX + ii;
                const double* xs = S.xs + k * HP + (jj + 1) * HX + (ii + 1);
                double rhs = S.bb[o];
                rhs = fma(-S.a1[o], xs[-1], rhs);
                rhs = fma(-S.a2[o], xs[1], rhs);
                rhs = fma(-S.a3[o], xs[-HX], rhs);
                rhs = fma(-S.a4[o], xs[HX], rhs);
                S.e[o] = rhs * S.id[o];
            }
            __syncthreads();
            if (sw == 0 && col == 0) LINE_STAMP(3);
            for (int q = wid; q < TX * TY; q += nw) {
                const int ii = q & (TX - 1), jj = q / TX;
                if (((par0 + ii + jj) & 1) != col) continue;
                double* ev = S.e + q;
                const double* lv = S.lf + q;
                const double* cv = S.cp + q;
                double* xcol = S.xs + (jj + 1) * HX + (ii + 1);
                if (mm <= 3)
                    line_column_solve<3>(ev, lv, cv, xcol, CP, HP, nz, lane);
                else if (mm <= 7)
                    line_column_solve<7>(ev, lv, cv, xcol, CP, HP, nz, lane);
                else
                    line_column_solve_rounds(ev, lv, cv, xcol, CP, HP, nz, lane);
            }
            __syncthreads();
            if (sw == 0 && col == 0) LINE_STAMP(4);
        }
    }
    LINE_STAMP(5);
#pragma unroll 2
    for (int idx = tid; idx < cells; idx += nth) {
        const int ii = idx & (TX - 1), r = idx / TX, jj = r % TY, k = r / TY;
        const int i = i0 + ii, j = j0 + jj;
        if (i < nx && j < ny) xout[i + nx * j + np * k] = S.xs[k * HP + (jj + 1) * HX + (ii + 1)];
    }
    LINE_STAMP(6);
}

// run-time tile shape -> compile-time instantiation (the shapes of line_tile_shape)
template <bool PROLONG>
__device__ __forceinline__ void line_tile_smooth_any(int tx, int ty, const double* __restrict__ a,
                                                     const double* __restrict__ fac, const double* b, const double* xin,
                                                     double* xout, const LevGeom& g, const double* xc, int cnx, int cny,
                                                     double omega, int nsw, int i0, int j0, double* sm, int tid, int nth,
                                                     bool wait_first) {
    if (tx == 8 && ty == 3)
        line_tile_smooth<PROLONG, 8, 3>(a, fac, b, xin, xout, g, xc, cnx, cny, omega, nsw, i0, j0, sm, tid, nth, wait_first);
    else if (tx == 4 && ty == 2)
        line_tile_smooth<PROLONG, 4, 2>(a, fac, b, xin, xout, g, xc, cnx, cny, omega, nsw, i0, j0, sm, tid, nth, wait_first);
    else
        line_tile_smooth<PROLONG, 1, 1>(a, fac, b, xin, xout, g, xc, cnx, cny, omega, nsw, i0, j0, sm, tid, nth, wait_first);
}

template <bool PROLONG>
__global__ void __launch_bounds__(LS_THREADS) line_smooth_kernel(const double* __restrict__ a, const double* __restrict__ fac,
                                                                 const double* b, const double* xin, double* xout,
                                                                 LevGeom g, const double* xc, int cnx, int cny, double omega,
                                                                 int nsw, int tx, int ty) {
    extern __shared__ double ls_sm[];
    pdl_launch_dependents();
    const int ntx = (g.nx + tx - 1) / tx;
    const int i0 = ((int)blockIdx.x % ntx) * tx, j0 = ((int)blockIdx.x / ntx) * ty;
    line_tile_smooth_any<PROLONG>(tx, ty, a, fac, b, xin, xout, g, xc, cnx, cny, omega, nsw, i0, j0, ls_sm, (int)threadIdx.x,
                                  (int)blockDim.x, true);
}

// one colour of a red-black Gauss-Seidel sweep: colour = (i+j+k)&1.  Threads walk the cells of the
// colour only (i = 2*ih + parity of the row).  zero_guess: x == 0 on entry, no neighbour reads.
// PROLONG: the coarse correction is folded into this pass - a Gauss-Seidel update never reads the cell's own
// old value, so after a correction x += omega P xc only the OTHER colour's corrected values matter, and they
// are formed on the fly as x[nb] + omega xc[aggregate(nb)] (saves the prolongation pass over x entirely).
template <int NS, bool PROLONG>
__device__ __forceinline__ void rbgs_cell(const double* __restrict__ a, const double* b, double* x,
                                          const LevGeom& g, int i, int j, int k, bool zero_guess,
                                          const double* xc, int cnx, int cny, double omega) {
    long long c = i + (long long)g.nx * (j + (long long)g.ny * k);
    double acc = b[c];
    if (!zero_guess) {
        double t[NS];
#pragma unroll
        for (int s = 1; s < NS; s++) {
            bool ex;
            long long nb = nbr_clamped(g.nx, g.ny, g.nz, i, j, k, c, s, ex);
            double xv = x[nb];
            if (PROLONG) {
                const int axis = (s - 1) >> 1, d = ((s - 1) & 1) ? 1 : -1;
                const int ii = ex ? i + (axis == 0 ? d : 0) : i, jj = ex ? j + (axis == 1 ? d : 0) : j,
                          kk = ex ? k + (axis == 2 ? d : 0) : k;
                xv += omega * xc[(ii >> (g.cx - 1)) + (long long)cnx * ((jj >> (g.cy - 1)) + (long long)cny * (kk >> (g.cz - 1)))];
            }
            double p = a[(long long)s * g.n + c] * xv;
            t[s] = ex ? p : 0.0;
        }
#pragma unroll
        for (int s = 1; s < NS; s++) acc -= t[s];
    }
    double d0 = a[c];
    x[c] = d0 != 0.0 ? acc / d0 : 0.0;
}

template <int NS, bool PROLONG>
__global__ void __launch_bounds__(256) rbgs_kernel(const double* __restrict__ a, const double* b,
                                                   double* x, LevGeom g, int col, int zero_guess,
                                                   const double* xc, int cnx, int cny, double omega) {
    const int nxh = (g.nx + 1) >> 1;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool active = t < (long long)g.ny * g.nz * nxh;
    int i = 0, j = 0, k = 0;
    if (active) {
        int ih;
        tpb_ijk(t, nxh, g.ny, ih, j, k);
        i = 2 * ih + ((col + j + k) & 1);
        active = i < g.nx;
    }
    pdl_launch_dependents();
    const long long c = i + (long long)g.nx * (j + (long long)g.ny * k);
    double aa[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) aa[s] = 0.0;
    if (active) {
        aa[0] = a[c];
        if (!zero_guess) {
#pragma unroll
            for (int s = 1; s < NS; s++) aa[s] = a[(long long)s * g.n + c];
        }
    }
    pdl_wait();
    if (!active) return;
    // from here on the arithmetic is rbgs_cell's, in the same order
    double acc = b[c];
    if (!zero_guess) {
        double tt[NS];
#pragma unroll
        for (int s = 1; s < NS; s++) {
            bool ex;
            long long nb = nbr_clamped(g.nx, g.ny, g.nz, i, j, k, c, s, ex);
            double xv = x[nb];
            if (PROLONG) {
                const int axis = (s - 1) >> 1, d = ((s - 1) & 1) ? 1 : -1;
                const int ii = ex ? i + (axis == 0 ? d : 0) : i, jj = ex ? j + (axis == 1 ? d : 0) : j,
                          kk = ex ? k + (axis == 2 ? d : 0) : k;
                xv += omega * xc[(ii >> (g.cx - 1)) + (long long)cnx * ((jj >> (g.cy - 1)) + (long long)cny * (kk >> (g.cz - 1)))];
            }
            double p = aa[s] * xv;
            tt[s] = ex ? p : 0.0;
        }
#pragma unroll
        for (int s = 1; s < NS; s++) acc -= tt[s];
    }
    x[c] = aa[0] != 0.0 ? acc / aa[0] : 0.0;
}

// bc[C] = sum over the aggregate of (b - A x)  (residual + restriction fused).  The aggregate is walked as a
// fully unrolled 2x2x2 box; the summation order (k, j, i; within a cell diag, x-, x+, ...) is the one the CPU
// restatement uses.
// Point-smoothed levels: only the cells of colour 0 contribute - the restriction always follows a pre-smoothing sweep
// whose last pass updated colour 1, and a Gauss-Seidel update leaves a zero residual in the rows it solved (half of the
// loads).  Line-smoothed levels sum every row: the hybrid smoother's tiles freeze their rims, no row is exactly solved.
template <int NS>
__device__ __forceinline__ double restrict_cell(const double* __restrict__ a, const double* b, const double* x,
                                                const LevGeom& f, int I, int Jc, int Kc) {
    double r[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const int di = q & 1, dj = (q >> 1) & 1, dk = q >> 2;
        const int i = I * f.cx + di, j = Jc * f.cy + dj, k = Kc * f.cz + dk;
        const bool ok = di < f.cx && dj < f.cy && dk < f.cz && i < f.nx && j < f.ny && k < f.nz &&
                        (f.line || ((i + j + k) & 1) == 0);
        r[q] = 0.0;
        if (ok) {
            long long c = i + (long long)f.nx * (j + (long long)f.ny * k);
            double acc = b[c] - a[c] * x[c];
#pragma unroll
            for (int s = 1; s < NS; s++) {
                bool ex;
                long long nb = nbr_clamped(f.nx, f.ny, f.nz, i, j, k, c, s, ex);
                double p = a[(long long)s * f.n + c] * x[nb];
                acc -= ex ? p : 0.0;
            }
            r[q] = acc;
        }
    }
    double sum = 0.0;
#pragma unroll
    for (int q = 0; q < 8; q++) sum += r[q];
    return sum;
}

// One thread per FINE cell of the aggregate box (8 lanes per coarse cell): each thread has one row's 2*NS loads in
// flight instead of eight rows' one after the other (measured: 13 us -> latency of one row on the small levels);
// lane 0 of the group then adds the eight residuals in the fixed order q = 0..7 of restrict_cell.
template <int NS>
__global__ void __launch_bounds__(256) restrict_kernel(const double* __restrict__ a, const double* b,
                                                       const double* x, LevGeom f, LevGeom cg,
                                                       double* __restrict__ bc) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long C = t >> 3;
    const int q = (int)(t & 7);
    bool ok = false;
    int i = 0, j = 0, k = 0;
    if (C < cg.n) {
        int I, Jc, Kc;
        tpb_ijk(C, cg.nx, cg.ny, I, Jc, Kc);
        const int di = q & 1, dj = (q >> 1) & 1, dk = q >> 2;
        i = I * f.cx + di, j = Jc * f.cy + dj, k = Kc * f.cz + dk;
        ok = di < f.cx && dj < f.cy && dk < f.cz && i < f.nx && j < f.ny && k < f.nz && (f.line || ((i + j + k) & 1) == 0);
    }
    pdl_launch_dependents();
    const long long c = ok ? i + (long long)f.nx * (j + (long long)f.ny * k) : 0;
    double aa[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) aa[s] = ok ? a[(long long)s * f.n + c] : 0.0;   // the operator does not change during a solve
    pdl_wait();
    double r = 0.0;
    if (ok) {
        double acc = b[c] - aa[0] * x[c];
#pragma unroll
        for (int s = 1; s < NS; s++) {
            bool ex;
            long long nb = nbr_clamped(f.nx, f.ny, f.nz, i, j, k, c, s, ex);
            double p = aa[s] * x[nb];
            acc -= ex ? p : 0.0;
        }
        r = acc;
    }
    // whole warps reach this point (the grid is padded to full blocks)
    const int base = (threadIdx.x & 31) & ~7;
    double sum = 0.0;
#pragma unroll
    for (int m = 0; m < 8; m++) sum += __shfl_sync(0xffffffffu, r, base + m);
    if (q == 0 && C < cg.n) bc[C] = sum;
}

// Line-smoothed levels (z never coarsened, every row contributes): one thread per coarse cell walks its <= 2 x 2 fine
// cells; the two x-neighbours share their sectors inside the thread, so the big levels run at memory speed with a
// quarter of the threads of the 8-lane kernel above.  Summation order as restrict_cell (j, then i).
__global__ void __launch_bounds__(256) restrict_line_kernel(const double* __restrict__ a, const double* b, const double* x,
                                                            LevGeom f, LevGeom cg, double* __restrict__ bc) {
    const long long C = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    pdl_launch_dependents();
    int I = 0, Jc = 0, k = 0;
    const bool on = C < cg.n;
    if (on) tpb_ijk(C, cg.nx, cg.ny, I, Jc, k);
    double aa[4][7];
    bool ok[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int di = q & 1, dj = q >> 1;
        const int i = I * f.cx + di, j = Jc * f.cy + dj;
        ok[q] = on && di < f.cx && dj < f.cy && i < f.nx && j < f.ny;
        const long long c = ok[q] ? i + (long long)f.nx * (j + (long long)f.ny * k) : 0;
#pragma unroll
        for (int s = 0; s < 7; s++) aa[q][s] = ok[q] ? a[(long long)s * f.n + c] : 0.0;   // set-up data: before the wait
    }
    pdl_wait();
    double sum = 0.0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        if (!ok[q]) continue;
        const int di = q & 1, dj = q >> 1;
        const int i = I * f.cx + di, j = Jc * f.cy + dj;
        const long long c = i + (long long)f.nx * (j + (long long)f.ny * k);
        double acc = b[c] - aa[q][0] * x[c];
#pragma unroll
        for (int s = 1; s < 7; s++) {
            bool ex;
            const long long nb = nbr_clamped(f.nx, f.ny, f.nz, i, j, k, c, s, ex);
            const double p = aa[q][s] * x[nb];
            acc -= ex ? p : 0.0;
        }
        sum += acc;
    }
    if (on) bc[C] = sum;
}

__global__ void __launch_bounds__(256) prolong_add_kernel(const double* __restrict__ xc, LevGeom f, LevGeom cg,
                                                          double omega, const double* xin, double* x) {
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= f.n) return;
    int i, j, k;
    tpb_ijk(c, f.nx, f.ny, i, j, k);
    long long C = (i >> (f.cx - 1)) + (long long)cg.nx * ((j >> (f.cy - 1)) + (long long)cg.ny * (k >> (f.cz - 1)));
    x[c] = xin[c] + omega * xc[C];   // (xin == x on point-smoothed levels, the iterate after pre-smoothing otherwise)
}

// residual only (used between V-cycles when mg_cycles > 1): r = b - A x
template <int NS>
__global__ void __launch_bounds__(256) residual_kernel(const double* __restrict__ a, const double* __restrict__ b,
                                                       const double* __restrict__ x, LevGeom g, double* __restrict__ r) {
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= g.n) return;
    int i, j, k;
    tpb_ijk(c, g.nx, g.ny, i, j, k);
    double acc = b[c] - a[c] * x[c];
#pragma unroll
    for (int s = 1; s < NS; s++) {
        long long nb = nbr_cell(g.nx, g.ny, g.nz, i, j, k, c, s);
        if (nb >= 0) acc -= a[(long long)s * g.n + c] * x[nb];
    }
    r[c] = acc;
}

// ---- the coarse tail: every level from `l0` down runs inside ONE CTA (levels of <= TAIL_CELLS cells) ---
constexpr int TAIL_CELLS = 4096;
constexpr int TAIL_THREADS = 1024;
struct TailLevel {
    LevGeom g;
    const double* a;
    double* x;
    double* b;
    const double* fac;      // line-smoothed levels: Thomas factors,
    double *xt, *xs0, *xs1; // the iterate between pre- and post-smoothing and two work vectors (smoothing is out of place)
    int tx, ty;             // tile shape
};
struct TailArgs {
    int nlev;
    TailLevel lev[MAXLEV];
    int pre, post, coarse_sweeps, tile_sweeps;
    double omega;
};

// The sub-V-cycle over a run of levels inside one kernel; the barrier between passes is a template parameter.
// Measured on B200 (60x220x85, r1): sharing the levels of <= 160 k cells among a co-resident cooperative grid
// (grid.sync) was no faster than one launch per pass replayed from a CUDA graph, and one 8-CTA thread-block cluster
// (cluster.sync) was 12 % slower (too little memory-level parallelism); r2 tried a counter barrier over a co-resident
// grid for the line smoother's passes: no faster either (a grid-wide barrier costs what a dependent launch costs).
// So only the single-CTA tail (levels <= TAIL_CELLS, __syncthreads) is kept.  Level vectors b, x are written and
// re-read inside the same kernel, so they are plain pointers here (no __restrict__/read-only path).
struct BlockBarrier {
    __device__ __forceinline__ void operator()() const { __syncthreads(); }
};

extern __shared__ double tail_sm[];

// `nsw` sweeps of the hybrid line smoother on a whole level by one block: groups of tile_sweeps sweeps, the tiles of a
// group one after the other (out of place, so the order does not matter); same buffer rotation as oracle/cport
// mg_line_smooth.  xin == nullptr: zero guess; PROLONG: xin + omega * P xc is what the first group reads.
template <bool PROLONG, class Barrier>
__device__ __forceinline__ void cyc_line_smooth(const TailLevel& L, const double* xin, double* xout, int nsw, int group,
                                                const double* xc, int cnx, int cny, double omega, Barrier& bar) {
    const LevGeom& g = L.g;
    if (group <= 0 || group > nsw) group = nsw;
    const int ncalls = (nsw + group - 1) / group;
    const int ntx = (g.nx + L.tx - 1) / L.tx, nty = (g.ny + L.ty - 1) / L.ty;
    const double* in = xin;
    int left = nsw;
    for (int c = 0; c < ncalls; c++) {
        double* out = (c == ncalls - 1) ? xout : (((ncalls - 1 - c) & 1) ? L.xs0 : L.xs1);
        const int sw = left < group ? left : group;
        for (int t = 0; t < ntx * nty; t++) {
            const int i0 = (t % ntx) * L.tx, j0 = (t / ntx) * L.ty;
            if (PROLONG && c == 0)
                line_tile_smooth_any<true>(L.tx, L.ty, L.a, L.fac, L.b, in, out, g, xc, cnx, cny, omega, sw, i0, j0, tail_sm,
                                           (int)threadIdx.x, (int)blockDim.x, false);
            else
                line_tile_smooth_any<false>(L.tx, L.ty, L.a, L.fac, L.b, in, out, g, nullptr, 0, 0, 0.0, sw, i0, j0, tail_sm,
                                            (int)threadIdx.x, (int)blockDim.x, false);
            __syncthreads();   // the tile's shared memory is reused by the next one
        }
        bar();
        left -= sw;
        in = out;
    }
}

template <int NS, bool PROLONG, class Barrier>
__device__ __forceinline__ void cyc_sweep(const TailLevel& L, bool zero_guess, const double* xc, int cnx, int cny,
                                          double omega, long long tid, long long nth, Barrier& bar) {
    const LevGeom& g = L.g;
    const int nxh = (g.nx + 1) >> 1;
    const long long total = (long long)g.ny * g.nz * nxh;
    for (int col = 0; col < 2; col++) {
        for (long long t = tid; t < total; t += nth) {
            int ih, j, k;
            tpb_ijk(t, nxh, g.ny, ih, j, k);
            int i = 2 * ih + ((col + j + k) & 1);
            if (i < g.nx) {
                if (PROLONG && col == 0)
                    rbgs_cell<NS, true>(L.a, L.b, L.x, g, i, j, k, false, xc, cnx, cny, omega);
                else
                    rbgs_cell<NS, false>(L.a, L.b, L.x, g, i, j, k, zero_guess && col == 0, nullptr, 0, 0, 0.0);
            }
        }
        bar();
    }
}

// levels l0 .. l1-1: pre-smooth, restrict the residual into level l+1
template <int NS, class Barrier>
__device__ __forceinline__ void cyc_down(const TailArgs& A, int l0, int l1, long long tid, long long nth, Barrier& bar) {
    for (int l = l0; l < l1; l++) {
        const TailLevel& L = A.lev[l];
        const bool line = NS == 7 && L.g.line;
        if (line)
            cyc_line_smooth<false>(L, nullptr, L.xt, A.pre, A.tile_sweeps, nullptr, 0, 0, 0.0, bar);
        else
            for (int s = 0; s < A.pre; s++) cyc_sweep<NS, false>(L, s == 0, nullptr, 0, 0, 0.0, tid, nth, bar);
        const double* cur = line ? L.xt : L.x;
        const TailLevel& Cc = A.lev[l + 1];
        for (long long C = tid; C < Cc.g.n; C += nth) {
            int I, Jc, Kc;
            tpb_ijk(C, Cc.g.nx, Cc.g.ny, I, Jc, Kc);
            Cc.b[C] = restrict_cell<NS>(L.a, L.b, cur, L.g, I, Jc, Kc);
        }
        bar();
    }
}

template <int NS, class Barrier>
__device__ __forceinline__ void cyc_coarsest(const TailArgs& A, long long tid, long long nth, Barrier& bar) {
    const TailLevel& L = A.lev[A.nlev - 1];
    if (NS == 7 && L.g.line)
        cyc_line_smooth<false>(L, nullptr, L.x, A.coarse_sweeps, A.tile_sweeps, nullptr, 0, 0, 0.0, bar);
    else
        for (int s = 0; s < A.coarse_sweeps; s++) cyc_sweep<NS, false>(L, s == 0, nullptr, 0, 0, 0.0, tid, nth, bar);
}

// levels l1-1 down to l0: coarse correction (folded into the first post-smoothing sweep) + post-smoothing
template <int NS, class Barrier>
__device__ __forceinline__ void cyc_up(const TailArgs& A, int l1, int l0, long long tid, long long nth, Barrier& bar) {
    for (int l = l1 - 1; l >= l0; l--) {
        const TailLevel& L = A.lev[l];
        const TailLevel& Cc = A.lev[l + 1];
        const LevGeom& f = L.g;
        const bool line = NS == 7 && f.line;
        if (A.post > 0 && line) {
            cyc_line_smooth<true>(L, L.xt, L.x, A.post, A.tile_sweeps, Cc.x, Cc.g.nx, Cc.g.ny, A.omega, bar);
        } else if (A.post > 0) {
            cyc_sweep<NS, true>(L, false, Cc.x, Cc.g.nx, Cc.g.ny, A.omega, tid, nth, bar);
            for (int s = 1; s < A.post; s++) cyc_sweep<NS, false>(L, false, nullptr, 0, 0, 0.0, tid, nth, bar);
        } else {
            const double* cur = line ? L.xt : L.x;
            for (long long c = tid; c < f.n; c += nth) {
                int i, j, k;
                tpb_ijk(c, f.nx, f.ny, i, j, k);
                long long C = (i >> (f.cx - 1)) + (long long)Cc.g.nx * ((j >> (f.cy - 1)) + (long long)Cc.g.ny * (k >> (f.cz - 1)));
                L.x[c] = cur[c] + A.omega * Cc.x[C];
            }
            bar();
        }
    }
}

template <int NS>
__global__ void __launch_bounds__(TAIL_THREADS) tail_kernel(TailArgs A) {
    pdl_launch_dependents();
    pdl_wait();
    BlockBarrier bar;
    const long long tid = threadIdx.x, nth = blockDim.x;
    cyc_down<NS>(A, 0, A.nlev - 1, tid, nth, bar);
    cyc_coarsest<NS>(A, tid, nth, bar);
    cyc_up<NS>(A, A.nlev - 1, 0, tid, nth, bar);
}

// multi-rank hierarchies: the single-CTA zone above the gather level is cut in two around the all-gather
template <int NS>
__global__ void __launch_bounds__(TAIL_THREADS) tail_down_kernel(TailArgs A) {
    pdl_launch_dependents();
    pdl_wait();
    BlockBarrier bar;
    cyc_down<NS>(A, 0, A.nlev - 1, (long long)threadIdx.x, (long long)blockDim.x, bar);
}
template <int NS>
__global__ void __launch_bounds__(TAIL_THREADS) tail_up_kernel(TailArgs A) {
    pdl_launch_dependents();
    pdl_wait();
    BlockBarrier bar;
    cyc_up<NS>(A, A.nlev - 1, 0, (long long)threadIdx.x, (long long)blockDim.x, bar);
}

// ... or, with the peer-memory mailboxes up and the whole gathered hierarchy inside the single-CTA zone, not cut at
// all: ONE block restricts down the slab's last levels, all-gathers the gather level's right-hand side through the
// mailboxes (p2p_gather_block: stores into the peers' HBM, flags, copy-out), runs the gathered levels (redundantly on
// every rank, identical bits) and prolongs/post-smooths back up - four dependent single-CTA launches become one.
template <int NS>
__global__ void __launch_bounds__(TAIL_THREADS) tail_dist_kernel(TailArgs A, TailArgs G, P2PView v, int hier, double* gbuf,
                                                                 long long my_off, long long my_cnt, long long total) {
    pdl_launch_dependents();
    pdl_wait();
    BlockBarrier bar;
    const long long tid = threadIdx.x, nth = blockDim.x;
    cyc_down<NS>(A, 0, A.nlev - 1, tid, nth, bar);
    p2p_gather_block(v, hier, gbuf, my_off, my_cnt, total);
    cyc_down<NS>(G, 0, G.nlev - 1, tid, nth, bar);
    cyc_coarsest<NS>(G, tid, nth, bar);
    cyc_up<NS>(G, G.nlev - 1, 0, tid, nth, bar);
    cyc_up<NS>(A, A.nlev - 1, 0, tid, nth, bar);
}

// ---- K6 / coupling kernels ----------------------------------------------------------------------
// CPR restriction: rp = x_p - sum_f w_f x_f
template <int NF>
__global__ void __launch_bounds__(256) cpr_restrict_kernel(const double* __restrict__ x, const double* __restrict__ w1,
                                                           const double* __restrict__ w2, long long n,
                                                           double* __restrict__ rp) {
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    pdl_launch_dependents();
    const bool in = c < n;
    const double a1 = in ? w1[c] : 0.0, a2 = (in && NF == 3) ? w2[c] : 0.0;   // set-up data: before the wait
    pdl_wait();
    if (!in) return;
    double v = x[c] - a1 * x[n + c];
    if (NF == 3) v -= a2 * x[2 * n + c];
    rp[c] = v;
}
// CPTR restriction: r_a = x_a - w_a x_S, a in (p, T)
template <int NF>
__global__ void __launch_bounds__(256) cptr_restrict_kernel(const double* __restrict__ x, const double* __restrict__ w0,
                                                            const double* __restrict__ w1, long long n,
                                                            double* __restrict__ rp, double* __restrict__ rT) {
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    pdl_launch_dependents();
    const bool in = c < n;
    const double a0 = in ? w0[c] : 0.0, a1 = in ? w1[c] : 0.0;   // set-up data: before the wait
    pdl_wait();
    if (!in) return;
    double xs = NF == 3 ? x[2 * n + c] : 0.0;
    rp[c] = x[c] - a0 * xs;
    rT[c] = x[n + c] - a1 * xs;
}
// r -= A00[.,a,b] x  (one coupling block of the 2x2 primary system)
template <int DIM>
__global__ void __launch_bounds__(256) a00_sub_kernel(const double* __restrict__ A00, int a, int b,
                                                      const double* x, Geom g, double* r) {
    const long long n = g.n;
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    pdl_launch_dependents();
    const bool in = c < n;
    int i = 0, j = 0, k = 0;
    double co[2 * DIM + 1];
    if (in) tpb_ijk(c, g.nx, g.ny, i, j, k);
#pragma unroll
    for (int s = 0; s < 2 * DIM + 1; s++) co[s] = in ? A00[((long long)(s * 2 + a) * 2 + b) * n + c] : 0.0;
    pdl_wait();
    if (!in) return;
    double acc = 0.0;
#pragma unroll
    for (int s = 0; s < 2 * DIM + 1; s++) {
        long long nb = nbr_cell(g.nx, g.ny, g.nz, i, j, k, c, s);
        if (nb >= 0) acc += co[s] * x[nb];
    }
    r[c] -= acc;
}

// selfp Schur approximation (PETSc pc_fieldsplit_schur_precondition selfp, singlephase.py:322-329):
// AT = A11 - A10 diag(A00)^-1 A01, kept on the 5|7-point stencil: a product A10[s1] * A01[s2] lands on slot s2 when
// s1 is the diagonal, on slot s1 when s2 is, on the diagonal when s2 leads back to the row's cell, and is lumped
// into the diagonal otherwise (two steps away: outside the stencil).  AT holds A11 on entry.
template <int DIM>
__global__ void __launch_bounds__(128) selfp_kernel(const double* __restrict__ A00, Geom g, double* __restrict__ AT) {
    const long long n = g.n;
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    constexpr int NS = 2 * DIM + 1;
    const int nx = g.nx, ny = g.ny, nz = g.nz;
    int i, j, k;
    tpb_ijk(c, nx, ny, i, j, k);
    double sub[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) sub[s] = 0.0;
#pragma unroll
    for (int s1 = 0; s1 < NS; s1++) {
        long long nb = nbr_cell(nx, ny, nz, i, j, k, c, s1);
        if (nb < 0) continue;
        const double d = A00[nb];   // (s=0, a=0, b=0) of the neighbour
        const double f = d != 0.0 ? A00[((long long)(s1 * 2 + 1) * 2 + 0) * n + c] / d : 0.0;
        int i2, j2, k2;
        tpb_ijk(nb, nx, ny, i2, j2, k2);
#pragma unroll
        for (int s2 = 0; s2 < NS; s2++) {
            if (nbr_cell(nx, ny, nz, i2, j2, k2, nb, s2) < 0) continue;
            const double v = f * A00[((long long)(s2 * 2 + 0) * 2 + 1) * n + nb];
            const int slot = s1 == 0 ? s2 : (s2 == 0 ? s1 : 0);
            sub[slot] += v;
        }
    }
#pragma unroll
    for (int s = 0; s < NS; s++) AT[(long long)s * n + c] -= sub[s];
}

// ---- K8: red-black block ILU(0) --------------------------------------------------------------------
template <int NF, int DIM>
__global__ void __launch_bounds__(128) ilu_setup_kernel(const double* __restrict__ J, Geom g, int col, int ilu,
                                                        double* Dinv, float* __restrict__ Dc, float* __restrict__ Lc) {
    const long long n = g.n;
    const int nx = g.nx, ny = g.ny, nz = g.nz;
    const int nxh = (nx + 1) >> 1;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)ny * nz * nxh) return;
    int ih, j, k;
    tpb_ijk(t, nxh, ny, ih, j, k);
    int i = 2 * ih + ((col + j + k) & 1);
    if (i >= nx) return;
    long long c = i + (long long)nx * (j + (long long)ny * k);
    double D[NF * NF], Di[NF * NF];
#pragma unroll
    for (int r = 0; r < NF; r++)
#pragma unroll
        for (int q = 0; q < NF; q++) D[r * NF + q] = JAT(0, r, q, c);
    if (col == 1 && ilu) {
#pragma unroll
        for (int s = 1; s < 2 * DIM + 1; s++) {
            long long nb = nbr_cell(nx, ny, nz, i, j, k, c, s);
            if (nb < 0) continue;
            double T1[NF * NF];
#pragma unroll
            for (int r = 0; r < NF; r++)
#pragma unroll
                for (int q = 0; q < NF; q++) {
                    double acc = 0.0;
#pragma unroll
                    for (int m = 0; m < NF; m++) acc += JAT(s, r, m, c) * Dinv[(long long)(m * NF + q) * n + nb];
                    T1[r * NF + q] = acc;
                }
#pragma unroll
            for (int r = 0; r < NF; r++)
#pragma unroll
                for (int q = 0; q < NF; q++) {
                    double acc = 0.0;
#pragma unroll
                    for (int m = 0; m < NF; m++) acc += T1[r * NF + m] * JAT(opp_slot(s), m, q, nb);
                    D[r * NF + q] -= acc;
                }
        }
    }
    inv_block(NF, D, Di);
#pragma unroll
    for (int e = 0; e < NF * NF; e++) Dinv[(long long)e * n + c] = Di[e];
    if (Lc) {
        // what the triangular solves read: colour-separated (cells of one colour contiguous, index t) fp32 copies of
        // the inverted diagonal blocks and of the off-diagonal blocks - a colour pass then uses every byte of
        // every sector it touches and moves half the bytes
        const long long nth = (long long)ny * nz * nxh;
#pragma unroll
        for (int e = 0; e < NF * NF; e++) Dc[((long long)col * NF * NF + e) * nth + t] = (float)Di[e];
#pragma unroll
        for (int s = 1; s < 2 * DIM + 1; s++)
#pragma unroll
            for (int e = 0; e < NF * NF; e++)
                Lc[(((long long)col * 2 * DIM + (s - 1)) * NF * NF + e) * nth + t] = (float)J[((long long)s * NF * NF + e) * n + c];
    }
}

// mode 0: z = Dinv r (red, forward); 1: z = Dinv (r - sum A z[nb]) (black, forward);
// mode 2: z -= Dinv sum A z[nb] (red, backward).  Coefficients come from the colour-separated fp32 copies
// (the factor is a preconditioner: fp32 storage, fp64 arithmetic); vectors stay fp64 in natural order.
template <int NF, int DIM>
__global__ void __launch_bounds__(128) ilu_half_kernel(const float* __restrict__ Lc, const float* __restrict__ Dc,
                                                       const double* __restrict__ r, double* z, Geom g, int col,
                                                       int mode) {
    const long long n = g.n;
    const int nx = g.nx, ny = g.ny, nz = g.nz;
    const int nxh = (nx + 1) >> 1;
    const long long nth = (long long)ny * nz * nxh;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    pdl_launch_dependents();
    bool active = t < nth;
    int ih = 0, j = 0, k = 0, i = 0;
    if (active) {
        tpb_ijk(t, nxh, ny, ih, j, k);
        i = 2 * ih + ((col + j + k) & 1);
        active = i < nx;
    }
    // the inverted diagonal block is set-up data: load it before waiting for the predecessor
    float dd[NF * NF];
#pragma unroll
    for (int e = 0; e < NF * NF; e++) dd[e] = active ? Dc[((long long)col * NF * NF + e) * nth + t] : 0.f;
    pdl_wait();
    if (!active) return;
    long long c = i + (long long)nx * (j + (long long)ny * k);
    double tt[NF];
#pragma unroll
    for (int a = 0; a < NF; a++) tt[a] = 0.0;
    if (mode != 0) {
#pragma unroll
        for (int s = 1; s < 2 * DIM + 1; s++) {
            bool ex;
            long long nb = nbr_clamped(nx, ny, nz, i, j, k, c, s, ex);
            double zn[NF];
#pragma unroll
            for (int q = 0; q < NF; q++) zn[q] = z[(long long)q * n + nb];
            const float* L = Lc + (((long long)col * 2 * DIM + (s - 1)) * NF * NF) * nth + t;
#pragma unroll
            for (int a = 0; a < NF; a++) {
                double p = 0.0;
#pragma unroll
                for (int q = 0; q < NF; q++) p += (double)L[(long long)(a * NF + q) * nth] * zn[q];
                tt[a] += ex ? p : 0.0;
            }
        }
    }
    double v[NF];
#pragma unroll
    for (int a = 0; a < NF; a++) v[a] = mode == 2 ? tt[a] : r[(long long)a * n + c] - tt[a];
#pragma unroll
    for (int a = 0; a < NF; a++) {
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < NF; q++) acc += (double)dd[a * NF + q] * v[q];
        if (mode == 2)
            z[(long long)a * n + c] -= acc;
        else
            z[(long long)a * n + c] = acc;
    }
}

template <int NF>
__global__ void __launch_bounds__(256) bjacobi_kernel(const double* __restrict__ Dinv, const double* __restrict__ r,
                                                      double* __restrict__ z, long long n) {
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    double v[NF];
#pragma unroll
    for (int q = 0; q < NF; q++) v[q] = r[(long long)q * n + c];
#pragma unroll
    for (int a = 0; a < NF; a++) {
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < NF; q++) acc += Dinv[(long long)(a * NF + q) * n + c] * v[q];
        z[(long long)a * n + c] = acc;
    }
}

// =================================================================================================
// host side
// =================================================================================================
inline unsigned nblk(long long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

inline LevGeom lg(const MgLevel& L) { return LevGeom{L.nx, L.ny, L.nz, L.cx, L.cy, L.cz, L.n, L.line ? 1 : 0}; }

void mg_free_levels(MgHier& m) {
    for (int l = 0; l < m.nlev; l++) {
        if (m.lev[l].own_a) tpb_dfree(m.lev[l].a);
        tpb_dfree(m.lev[l].x_own);
        tpb_dfree(m.lev[l].b_own);
        tpb_dfree(m.lev[l].fac);
        m.lev[l] = MgLevel();
    }
    m.nlev = 0;
}

void mg_free(MgHier& m) {
    mg_free_levels(m);
    if (m.glob) {
        mg_free_levels(*m.glob);
        delete m.glob;
        m.glob = nullptr;
    }
    tpb_dfree(m.ga);
    m.ga = nullptr;
    m.ga_cap = 0;
}

// cells of all ranks on the gather level of a multi-rank hierarchy (TPB_MG_GATHER overrides); levels of at most
// TAIL_CELLS cells run inside one CTA, so with the default the whole global part of a V-cycle is one kernel
long long gather_cells() {
    static const long long v = getenv("TPB_MG_GATHER") ? atoll(getenv("TPB_MG_GATHER")) : 4096;   // = TAIL_CELLS
    return v;
}

// Builds levels 1.. of `m` from the level-0 operator a0 on an (nx, ny, nz) box.  dist: the box is this rank's
// slab; coarsening factors are agreed between the ranks (all-reduced coupling sums, the largest slab decides
// whether the slab axis can still be halved) and coarsening stops at the gather level.  planes: owned planes of
// every rank along the slab axis, updated to the last level built.
// dynamic shared memory of a hybrid line-smoother block on a level with nz planes
inline size_t line_smem(int nz) {
    int tx, ty;
    line_tile_shape(nz, tx, ty);
    return line_smem_doubles(nz, tx, ty) * sizeof(double);
}
constexpr size_t LINE_SMEM_BUDGET = 200 * 1024;

template <int NS>
void mg_coarsen_t(tpb_handle_s* h, MgHier& m, double* a0, int nx0, int ny0, int nz0, bool dist, std::vector<int>& planes) {
    PcState* pc = h->pc;
    const tpb_solver_opts& o = h->opts;
    // z-line smoothing on every level of a 3-D hierarchy (z is then never coarsened); falls back to the point smoother
    // when a one-column tile is more than a block's threads (LS_CPT cells each) or shared memory can hold
    const bool line = NS == 7 && o.mg_smoother == TPB_MG_ZLINE && nz0 > 1 && nz0 <= LS_CPT * LS_THREADS &&
                      line_smem(nz0) <= LINE_SMEM_BUDGET;
    // geometry of level 0 never changes between set-ups, and the coarsening schedule is recomputed from
    // the operator each time (as hypre's set-up is, preconditioners.py:878), so levels are re-allocated
    // only when their shape changes
    MgHier old;
    old.nlev = m.nlev;
    for (int l = 0; l < m.nlev; l++) old.lev[l] = m.lev[l];
    MgHier nw;
    auto take = [&](int l, int nx, int ny, int nz) -> MgLevel {
        MgLevel L;
        const long long n = (long long)nx * ny * nz;
        if (l < old.nlev && old.lev[l].cap >= n) {
            L = old.lev[l];
            old.lev[l] = MgLevel();
        } else {
            L.cap = n + n / 8 + 64;   // head-room: the coarsening schedule shifts a little between set-ups
            L.x_own = tpb_dalloc<double>(L.cap);
            L.b_own = tpb_dalloc<double>(L.cap);
            if (l > 0) {
                L.a = tpb_dalloc<double>((size_t)NS * L.cap);
                L.own_a = true;
            }
        }
        L.line = line;
        if (line && L.fac_cap < L.cap) {
            tpb_dfree(L.fac);
            L.fac = tpb_dalloc<double>((size_t)6 * L.cap);
            L.fac_cap = L.cap;
        }
        L.x = L.x_own;
        L.b = L.b_own;
        L.nx = nx;
        L.ny = ny;
        L.nz = nz;
        L.n = n;
        return L;
    };
    const int sax = h->g.dim - 1;   // slab axis
    int l = 0;
    bool side_used = false;
    nw.lev[0] = take(0, nx0, ny0, nz0);
    if (nw.lev[0].own_a) tpb_dfree(nw.lev[0].a);
    nw.lev[0].a = a0;
    nw.lev[0].own_a = false;
    for (;;) {
        MgLevel& L = nw.lev[l];
        L.cx = L.cy = L.cz = 1;
        if (line) {
            // the column factorisations (one dependent chain of nz divisions per column: ~40 us whatever the level) run
            // on the side stream beside the coarsening of the next levels; joined below
            TPB_CUDA(cudaEventRecord(h->ev_fork, h->stream));
            TPB_CUDA(cudaStreamWaitEvent(h->stream2, h->ev_fork, 0));
            line_factor_kernel<<<nblk((long long)L.nx * L.ny, 128), 128, 0, h->stream2>>>(L.a, lg(L), L.fac);
            h->launches++;
            side_used = true;
            if (L.nx == 1 && L.ny == 1) {   // a single column: the line solve is exact
                nw.last_sweeps = 1;
                break;
            }
        }
        int dims[3] = {L.nx, L.ny, L.nz};
        long long nglob = L.n;
        if (dist) {
            long long tot = 0;
            int mx = 0;
            for (int p : planes) {
                tot += p;
                mx = std::max(mx, p);
            }
            nglob = L.n / dims[sax] * tot;
            dims[sax] = mx;
        }
        // Line-smoothed hierarchies on slabs do not gather (TPB_MG_GATHER_LINE=1 brings it back): z is never coarsened,
        // so the gathered levels would be N x nz planes tall and walked by one CTA - measured on the stacked bench
        // workload: identical Krylov counts with and without, 5 % (N=2) to 35 % (N=4) more time per iteration with.
        // Every slab then ends on its own exactly solved column, the slabs stay coupled through the Krylov method only.
        static const bool gather_line = getenv("TPB_MG_GATHER_LINE") && atoi(getenv("TPB_MG_GATHER_LINE")) != 0;
        const bool gather = dist && (!line || gather_line);
        const long long stop_at = gather ? std::max<long long>(gather_cells(), o.mg_min_cells) : o.mg_min_cells;
        if (nglob <= stop_at || nglob <= 1 || l == MAXLEV - 1) break;
        TPB_CUDA(cudaMemsetAsync(pc->strength, 0, 4 * sizeof(double), h->stream));
        unsigned blocks = std::min<unsigned>(nblk(L.n, 256), 1184u);
        strength_kernel<NS><<<blocks, 256, 0, h->stream>>>(L.a, L.n, pc->strength);
        h->launches++;
        if (dist) tpb_allreduce_sum(h, pc->strength, 3);
        if (dist && dd_on_slabs()) tpb_allreduce_max(h, pc->strength + 3, 1);
        double m_ax[4];
        TPB_CUDA(cudaMemcpyAsync(m_ax, pc->strength, 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        TPB_CUDA(cudaStreamSynchronize(h->stream));
        // a diagonally dominant level is the last one (tpb_solver_opts.mg_dd_stop).  Single-rank handles only in r1:
        // on slabs the decision would have to be agreed between the ranks and a stopped hierarchy skip its gather level
        if (o.mg_dd_stop > 0.0 && (tpb_comm_size(h) == 1 || dd_on_slabs()) && m_ax[3] <= o.mg_dd_stop) {
            nw.last_sweeps = dd_sweeps(m_ax[3], o.mg_coarse_sweeps);
            break;
        }
        if (line) dims[2] = 1;   // z-line smoothing: z is never coarsened
        double mmax = 0.0;
        for (int ax = 0; ax < 3; ax++)
            if (dims[ax] > 1 && m_ax[ax] > mmax) mmax = m_ax[ax];
        int cf[3] = {1, 1, 1};
        bool any = false;
        for (int ax = 0; ax < 3; ax++)
            if (dims[ax] > 1 && (m_ax[ax] >= o.mg_semi_theta * mmax || nglob <= o.mg_full_below)) {
                cf[ax] = 2;
                any = true;
            }
        if (!any)
            for (int ax = 0; ax < 3; ax++)
                if (dims[ax] > 1) {
                    cf[ax] = 2;
                    any = true;
                }
        if (!any) break;
        L.cx = cf[0];
        L.cy = cf[1];
        L.cz = cf[2];
        nw.lev[l + 1] = take(l + 1, (L.nx + cf[0] - 1) / cf[0], (L.ny + cf[1] - 1) / cf[1], (L.nz + cf[2] - 1) / cf[2]);
        if (dist)
            for (int& p : planes) p = (p + cf[sax] - 1) / cf[sax];
        MgLevel& Cc = nw.lev[l + 1];
        coarsen_op_kernel<NS><<<nblk(Cc.n, 128), 128, 0, h->stream>>>(L.a, lg(L), lg(Cc),
                                                                     o.mg_coarse_scale > 0.0 ? o.mg_coarse_scale : 1.0, Cc.a);
        h->launches++;
        l++;
    }
    nw.nlev = l + 1;
    if (side_used) {
        TPB_CUDA(cudaEventRecord(h->ev_join, h->stream2));
        TPB_CUDA(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    }
    mg_free_levels(old);
    m.nlev = nw.nlev;
    m.last_sweeps = nw.last_sweeps;
    for (int q = 0; q < MAXLEV; q++) m.lev[q] = q < nw.nlev ? nw.lev[q] : MgLevel();
    TPB_CUDA(cudaGetLastError());
}

template <int NS>
void mg_setup_t(tpb_handle_s* h, MgHier& m, double* a0) {
    const int nranks = tpb_comm_size(h);
    const bool dist = nranks > 1;
    std::vector<int> planes;
    if (dist) planes = tpb_comm_planes(h);
    row_repair_kernel<NS><<<nblk(h->g.n, 256), 256, 0, h->stream>>>(a0, h->g.n);
    h->launches++;
    mg_coarsen_t<NS>(h, m, a0, h->g.nx, h->g.ny, h->g.nz, dist, planes);
    m.skip_glob = dist && m.last_sweeps > 0;   // (only with dd_on_slabs(): otherwise last_sweeps stays 0 on slabs)
    if (!dist || m.skip_glob) return;
    // ---- gather level: concatenate the ranks' last levels along the slab axis ----------------------
    const int rank = tpb_comm_rank(h);
    MgLevel& L = m.lev[m.nlev - 1];
    const int sax = h->g.dim - 1;
    const int ldims[3] = {L.nx, L.ny, L.nz};
    const long long npl = L.n / ldims[sax];   // cells per plane of the slab axis on this level
    m.gcnt.assign(nranks, 0);
    m.goff.assign(nranks, 0);
    long long gn = 0, gplanes = 0;
    for (int r = 0; r < nranks; r++) {
        m.goff[r] = gn;
        m.gcnt[r] = npl * planes[r];
        gn += m.gcnt[r];
        gplanes += planes[r];
    }
    TPB_REQUIRE(m.gcnt[rank] == L.n, TPB_ERR_STATE, "multigrid gather level: slab sizes disagree between ranks");
    if (m.ga_cap < gn) {
        tpb_dfree(m.ga);
        m.ga_cap = gn + gn / 8 + 64;
        m.ga = tpb_dalloc<double>((size_t)NS * m.ga_cap);
    }
    for (int s = 0; s < NS; s++)
        TPB_CUDA(cudaMemcpyAsync(m.ga + (long long)s * gn + m.goff[rank], L.a + (long long)s * L.n, L.n * sizeof(double),
                                 cudaMemcpyDeviceToDevice, h->stream));
    tpb_allgatherv(h, m.ga, m.goff.data(), m.gcnt.data(), NS, gn);
    if (!m.glob) m.glob = new MgHier();
    int gd[3] = {L.nx, L.ny, L.nz};
    gd[sax] = (int)gplanes;
    std::vector<int> none;
    mg_coarsen_t<NS>(h, *m.glob, m.ga, gd[0], gd[1], gd[2], false, none);
    // the gather level works directly on this rank's section of the gathered vectors
    L.x = m.glob->lev[0].x + m.goff[rank];
    L.b = m.glob->lev[0].b + m.goff[rank];
}

void mg_setup(tpb_handle_s* h, MgHier& m, double* a0) {
    if (h->ns == 7)
        mg_setup_t<7>(h, m, a0);
    else
        mg_setup_t<5>(h, m, a0);
}

// kernels with more than 48 KB of dynamic shared memory opt in once per process
template <typename K>
void want_smem(K kernel, size_t bytes) {
    static std::vector<std::pair<const void*, size_t>> done;
    for (auto& d : done)
        if (d.first == (const void*)kernel && d.second >= bytes) return;
    TPB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(bytes, 48 * 1024)));
    done.push_back({(const void*)kernel, bytes});
}

// `nsw` sweeps of the hybrid line smoother on a level: one launch per group of mg_tile_sweeps sweeps, out of place
// (xin == nullptr: zero guess; coarse != nullptr: the first group reads xin + omega P coarse->x); buffer rotation as in
// oracle/cport mg_line_smooth
void mg_line_smooth(tpb_handle_s* h, const MgLevel& L, const double* xin, double* xout, int nsw,
                    const MgLevel* coarse = nullptr, double omega = 0.0) {
    LevGeom g = lg(L);
    int group = h->opts.mg_tile_sweeps;
    if (group <= 0 || group > nsw) group = nsw;
    const int ncalls = (nsw + group - 1) / group;
    int tx, ty;
    line_tile_shape(L.nz, tx, ty);
    const unsigned grid = (unsigned)(((L.nx + tx - 1) / tx) * ((L.ny + ty - 1) / ty));
    const size_t smem = line_smem(L.nz);
    const double* in = xin;
    int left = nsw;
    for (int c = 0; c < ncalls; c++) {
        double* out = (c == ncalls - 1) ? xout : (((ncalls - 1 - c) & 1) ? L.xs0() : L.xs1());
        const int sw = std::min(left, group);
        if (coarse && c == 0) {
            want_smem(line_smooth_kernel<true>, smem);
            launch_pdl_smem(line_smooth_kernel<true>, grid, LS_THREADS, smem, h->stream, L.a, L.fac, L.b, in, out, g, coarse->x,
                            coarse->nx, coarse->ny, omega, sw, tx, ty);
        } else {
            want_smem(line_smooth_kernel<false>, smem);
            launch_pdl_smem(line_smooth_kernel<false>, grid, LS_THREADS, smem, h->stream, L.a, L.fac, L.b, in, out, g, nullptr,
                            0, 0, 0.0, sw, tx, ty);
        }
        h->launches++;
        left -= sw;
        in = out;
    }
}

// one red-black point Gauss-Seidel sweep = two colour passes
template <int NS>
void mg_rbgs(tpb_handle_s* h, const MgLevel& L, bool zero_guess, const MgLevel* coarse = nullptr, double omega = 0.0) {
    LevGeom g = lg(L);
    long long threads = (long long)L.ny * L.nz * ((L.nx + 1) >> 1);
    for (int col = 0; col < 2; col++) {
        if (coarse && col == 0)
            launch_pdl(rbgs_kernel<NS, true>, nblk(threads, 256), 256, h->stream, L.a, L.b, L.x, g, col, 0, coarse->x,
                       coarse->nx, coarse->ny, omega);
        else
            launch_pdl(rbgs_kernel<NS, false>, nblk(threads, 256), 256, h->stream, L.a, L.b, L.x, g, col,
                       (zero_guess && col == 0) ? 1 : 0, nullptr, 0, 0, 0.0);
        h->launches++;
    }
}

template <int NS>
void mg_vcycle_t(tpb_handle_s* h, MgHier& m) {
    const tpb_solver_opts& o = h->opts;
    const int pre = o.mg_pre > 0 ? o.mg_pre : 1;
    const int coarse_opt = o.mg_coarse_sweeps > 0 ? o.mg_coarse_sweeps : 1;
    const int coarse = m.last_sweeps > 0 ? m.last_sweeps : coarse_opt;
    const bool dist = m.glob != nullptr && !m.skip_glob;   // the last level is the gather level: smoothed as level 0 of m.glob
    const int last = m.nlev - 1;
    // level zones: [0, lcoop) one kernel per smoothing pass, [ltail, last] inside one CTA (levels <= TAIL_CELLS)
    // (line-smoothed levels only while they are a single tile: the one CTA walks a level's tiles one after the other,
    // a launch of its own spreads them over the SMs)
    static const int tail_tiles = getenv("TPB_MG_TAIL_TILES") ? atoi(getenv("TPB_MG_TAIL_TILES")) : 1;
    auto in_tail = [&](const MgLevel& L) {
        if (L.n > TAIL_CELLS) return false;
        if (!L.line) return true;
        int tx, ty;
        line_tile_shape(L.nz, tx, ty);
        return ((L.nx + tx - 1) / tx) * ((L.ny + ty - 1) / ty) <= tail_tiles;
    };
    int ltail = m.nlev;
    while (ltail > 0 && in_tail(m.lev[ltail - 1])) ltail--;
    const int lcoop = std::min(ltail, dist ? last : m.nlev);
    size_t tail_smem = 0;   // dynamic shared memory the single-CTA kernels of this cycle need (line tiles)
    auto tail_args_of = [&](MgHier& mm, int l0) {
        TailArgs A;
        A.nlev = mm.nlev - l0;
        for (int l = l0; l < mm.nlev; l++) {
            TailLevel& T = A.lev[l - l0];
            const MgLevel& L = mm.lev[l];
            T.g = lg(L);
            T.a = L.a;
            T.x = L.x;
            T.b = L.b;
            T.fac = L.fac;
            T.xt = T.xs0 = T.xs1 = nullptr;
            T.tx = T.ty = 1;
            if (L.line) {
                T.xt = L.xt();
                T.xs0 = L.xs0();
                T.xs1 = L.xs1();
                line_tile_shape(L.nz, T.tx, T.ty);
                tail_smem = std::max(tail_smem, line_smem(L.nz));
            }
        }
        A.pre = pre;
        A.post = o.mg_post;
        A.coarse_sweeps = mm.last_sweeps > 0 ? mm.last_sweeps : coarse_opt;
        A.tile_sweeps = o.mg_tile_sweeps;
        A.omega = o.mg_overcorrection;
        return A;
    };
    auto tail_args = [&](int l0) { return tail_args_of(m, l0); };
    for (int l = 0; l < lcoop; l++) {
        MgLevel& L = m.lev[l];
        if (!dist && l == last) {
            if (L.line)
                mg_line_smooth(h, L, nullptr, L.x, coarse);
            else
                for (int s = 0; s < coarse; s++) mg_rbgs<NS>(h, L, s == 0);
            break;
        }
        if (L.line)
            mg_line_smooth(h, L, nullptr, L.xt(), pre);
        else
            for (int s = 0; s < pre; s++) mg_rbgs<NS>(h, L, s == 0);
        MgLevel& Cc = m.lev[l + 1];
        if (NS == 7 && L.line && L.cz == 1)
            launch_pdl(restrict_line_kernel, nblk(Cc.n, 256), 256, h->stream, L.a, L.b, L.xt(), lg(L), lg(Cc), Cc.b);
        else
            launch_pdl(restrict_kernel<NS>, nblk(Cc.n * 8, 256), 256, h->stream, L.a, L.b, L.line ? L.xt() : L.x, lg(L), lg(Cc), Cc.b);
        h->launches++;
    }
    if (!dist) {
        if (ltail < m.nlev) {
            const TailArgs A = tail_args(ltail);
            want_smem(tail_kernel<NS>, tail_smem);
            launch_pdl_smem(tail_kernel<NS>, 1, TAIL_THREADS, tail_smem, h->stream, A);
            h->launches++;
        }
    } else if (P2PView pv; m.glob->lev[0].n <= TAIL_CELLS && tpb_p2p_view(h, 4, &pv) &&
                           m.goff.back() + m.gcnt.back() <= pv.mg_cap) {
        // slab tail + gather + gathered hierarchy + way back up in one single-CTA kernel (tail_dist_kernel)
        const int rank = tpb_comm_rank(h);
        const TailArgs A = tail_args(std::min(ltail, last)), G = tail_args_of(*m.glob, 0);
        want_smem(tail_dist_kernel<NS>, tail_smem);
        launch_pdl_smem(tail_dist_kernel<NS>, 1, TAIL_THREADS, tail_smem, h->stream, A, G, pv, &m == &h->pc->mg_T ? 1 : 0,
                        m.glob->lev[0].b, m.goff[rank], m.gcnt[rank], m.goff.back() + m.gcnt.back());
        h->launches++;
    } else {
        if (ltail < last) {
            const TailArgs A = tail_args(ltail);
            want_smem(tail_down_kernel<NS>, tail_smem);
            launch_pdl_smem(tail_down_kernel<NS>, 1, TAIL_THREADS, tail_smem, h->stream, A);
            h->launches++;
        }
        MgHier* mp = &m;
        const int rank = tpb_comm_rank(h);
        const long long gn = m.goff.back() + m.gcnt.back();
        // peer-memory gather kernel (stays inside the graph capture), else NCCL between two graphs
        if (!tpb_p2p_gather(h, &m == &h->pc->mg_T ? 1 : 0, m.glob->lev[0].b, m.goff[rank], m.gcnt[rank], gn))
            tpb_comm_op(h, [h, mp]() {
                tpb_allgatherv(h, mp->glob->lev[0].b, mp->goff.data(), mp->gcnt.data(), 1, 0);
            });
        mg_vcycle_t<NS>(h, *m.glob);
        if (ltail < last) {
            const TailArgs A = tail_args(ltail);
            want_smem(tail_up_kernel<NS>, tail_smem);
            launch_pdl_smem(tail_up_kernel<NS>, 1, TAIL_THREADS, tail_smem, h->stream, A);
            h->launches++;
        }
    }
    for (int l = std::min(lcoop, last) - 1; l >= 0; l--) {
        MgLevel& L = m.lev[l];
        MgLevel& Cc = m.lev[l + 1];
        if (o.mg_post > 0) {
            // coarse correction folded into the first post-smoothing pass (see rbgs_cell / line_tile_smooth)
            if (L.line) {
                mg_line_smooth(h, L, L.xt(), L.x, o.mg_post, &Cc, o.mg_overcorrection);
            } else {
                mg_rbgs<NS>(h, L, false, &Cc, o.mg_overcorrection);
                for (int s = 1; s < o.mg_post; s++) mg_rbgs<NS>(h, L, false);
            }
        } else {
            prolong_add_kernel<<<nblk(L.n, 256), 256, 0, h->stream>>>(Cc.x, lg(L), lg(Cc), o.mg_overcorrection,
                                                                    L.line ? L.xt() : L.x, L.x);
            h->launches++;
        }
    }
}

// y = V(b): mg_cycles V-cycles from a zero guess.  b and y are level-0 sized device vectors.
template <int NS>
void mg_apply_t(tpb_handle_s* h, MgHier& m, const double* b, double* y) {
    MgLevel& L = m.lev[0];
    const int cycles = h->opts.mg_cycles > 0 ? h->opts.mg_cycles : 1;
    if (m.glob && !m.skip_glob && m.nlev == 1) {
        // level 0 is itself the gather level (small grids on several ranks): its vectors are sections of the
        // gathered level, so copy in and out
        tpb_copy(h, (size_t)L.n, b, L.b);
        mg_vcycle_t<NS>(h, m);
        tpb_copy(h, (size_t)L.n, L.x, y);
    } else {
        double* keep_b = L.b;
        double* keep_x = L.x;
        // run level 0 directly on the caller's vectors (no copies)
        L.b = const_cast<double*>(b);
        L.x = y;
        mg_vcycle_t<NS>(h, m);
        L.b = keep_b;
        L.x = keep_x;
    }
    for (int cyc = 1; cyc < cycles; cyc++) {
        residual_kernel<NS><<<nblk(L.n, 256), 256, 0, h->stream>>>(L.a, b, y, lg(L), L.b);
        h->launches++;
        mg_vcycle_t<NS>(h, m);
        tpb_axpy(h, (size_t)L.n, 1.0, L.x, y);
    }
}

void mg_apply(tpb_handle_s* h, MgHier& m, const double* b, double* y) {
    if (h->ns == 7)
        mg_apply_t<7>(h, m, b, y);
    else
        mg_apply_t<5>(h, m, b, y);
}

template <int NF, int DIM>
void stage1_setup_t(tpb_handle_s* h, const double* J, const double* u, double dt) {
    PcState* pc = h->pc;
    const tpb_solver_opts& o = h->opts;
    const long long n = h->g.n;
    constexpr int NS = 2 * DIM + 1;
    if (o.stage1 == TPB_S1_NONE) return;
    for (int f = 0; f < TPB_MAXF; f++)
        if (!pc->w[f]) pc->w[f] = tpb_dalloc<double>(n);
    if (!pc->App) pc->App = tpb_dalloc<double>((size_t)NS * n);
    if (o.stage1 == TPB_S1_CPR) {
        cpr_setup_kernel<NF, DIM><<<nblk(n, 128), 128, 0, h->stream>>>(J, h->g, o.decoup, pc->w[1], pc->w[2], pc->App);
        h->launches++;
        mg_setup(h, pc->mg_p, pc->App);
        return;
    }
    if (!pc->A00) pc->A00 = tpb_dalloc<double>((size_t)NS * 4 * n);
    if (!pc->AT) pc->AT = tpb_dalloc<double>((size_t)NS * n);
    const int a11 = o.schur_pre != TPB_SCHUR_CONVDIFF;
    const int dec = o.stage1 == TPB_S1_CPTR ? o.decoup : TPB_DECOUP_NO;
    cptr_setup_kernel<NF, DIM><<<nblk(n, 128), 128, 0, h->stream>>>(J, h->g, dec, a11, pc->w[0], pc->w[1], pc->A00,
                                                                   pc->App, pc->AT);
    h->launches++;
    if (!a11) {
        CdIn in;
        const int np = h->g.np;
        for (int f = 0; f < TPB_MAXF; f++) {
            const int ff = f < NF ? f : 0;
            in.u[f] = GField{u + (size_t)ff * n, h->u_lo + (size_t)ff * np, h->u_hi + (size_t)ff * np};
        }
        auto gf = [&](int id) { return GField{h->fld[id], h->fld_lo[id], h->fld_hi[id]}; };
        in.phi = gf(TPB_PHI);
        in.K[0] = gf(TPB_KX);
        in.K[1] = gf(TPB_KY);
        in.K[2] = gf(DIM == 3 ? TPB_KZ : TPB_KY);
        in.kT = gf(TPB_KT);
        convdiff_kernel<NF, DIM><<<nblk(n, 128), 128, 0, h->stream>>>(in, 1.0 / dt, h->g, h->dp, pc->AT);
        h->launches++;
        if (h->nsrc_cells > 0) {
            convdiff_sources_kernel<NF><<<nblk(h->nsrc_cells, 128), 128, 0, h->stream>>>(
                h->nsrc_cells, h->src_cell, h->src_off, h->src_ent, u, h->fld[TPB_KX], h->fld[TPB_KY], n, h->dp, pc->AT);
            h->launches++;
        }
    }
    if (o.schur_pre == TPB_SCHUR_SELFP) {
        selfp_kernel<DIM><<<nblk(n, 128), 128, 0, h->stream>>>(pc->A00, h->g, pc->AT);
        h->launches++;
    }
    mg_setup(h, pc->mg_p, pc->App);
    mg_setup(h, pc->mg_T, pc->AT);
}

template <int NF, int DIM>
void stage2_setup_t(tpb_handle_s* h, const double* J) {
    PcState* pc = h->pc;
    const long long n = h->g.n;
    if (h->opts.stage2 == TPB_S2_NONE) return;
    if (!pc->Dinv) pc->Dinv = tpb_dalloc<double>((size_t)NF * NF * n);
    long long threads = (long long)h->g.ny * h->g.nz * ((h->g.nx + 1) >> 1);
    const bool ilu = h->opts.stage2 == TPB_S2_ILU0;
    if (ilu && !pc->Lc) {
        pc->Dc = tpb_dalloc<float>((size_t)2 * NF * NF * threads);
        pc->Lc = tpb_dalloc<float>((size_t)2 * 2 * DIM * NF * NF * threads);
    }
    for (int col = 0; col < 2; col++) {
        ilu_setup_kernel<NF, DIM><<<nblk(threads, 128), 128, 0, h->stream>>>(J, h->g, col, ilu ? 1 : 0, pc->Dinv,
                                                                            ilu ? pc->Dc : nullptr, ilu ? pc->Lc : nullptr);
        h->launches++;
    }
}

template <int NF, int DIM>
void stage2_apply_t(tpb_handle_s* h, const double* r, double* z) {
    PcState* pc = h->pc;
    const long long n = h->g.n;
    if (h->opts.stage2 == TPB_S2_BJACOBI) {
        bjacobi_kernel<NF><<<nblk(n, 256), 256, 0, h->stream>>>(pc->Dinv, r, z, n);
        h->launches++;
        return;
    }
    long long threads = (long long)h->g.ny * h->g.nz * ((h->g.nx + 1) >> 1);
    const int seq[3][2] = {{0, 0}, {1, 1}, {0, 2}};
    for (int q = 0; q < 3; q++) {
        launch_pdl(ilu_half_kernel<NF, DIM>, nblk(threads, 128), 128, h->stream, pc->Lc, pc->Dc, r, z, h->g, seq[q][0], seq[q][1]);
        h->launches++;
    }
}

template <int NF, int DIM>
void stage1_apply_t(tpb_handle_s* h, const double* x, double* y) {
    PcState* pc = h->pc;
    const tpb_solver_opts& o = h->opts;
    const long long n = h->g.n;
    double* rp = pc->t2;
    if (o.stage1 == TPB_S1_CPR) {
        tpb_zero(h, (size_t)(NF - 1) * n, y + n);
        launch_pdl(cpr_restrict_kernel<NF>, nblk(n, 256), 256, h->stream, x, pc->w[1], pc->w[2], n, rp);
        h->launches++;
        mg_apply(h, pc->mg_p, rp, y);
        return;
    }
    double* rT = pc->t2 + n;
    if (NF == 3) tpb_zero(h, (size_t)n, y + 2 * n);
    launch_pdl(cptr_restrict_kernel<NF>, nblk(n, 256), 256, h->stream, x, pc->w[0], pc->w[1], n, rp, rT);
    h->launches++;
    double* yp = y;
    double* yT = y + n;
    mg_apply(h, pc->mg_p, rp, yp);
    if (o.schur_pre == TPB_SCHUR_DIAG) {
        mg_apply(h, pc->mg_T, rT, yT);
        return;
    }
    launch_pdl(a00_sub_kernel<DIM>, nblk(n, 256), 256, h->stream, pc->A00, 1, 0, yp, h->g, rT);
    h->launches++;
    mg_apply(h, pc->mg_T, rT, yT);
    launch_pdl(a00_sub_kernel<DIM>, nblk(n, 256), 256, h->stream, pc->A00, 0, 1, yT, h->g, rp);
    h->launches++;
    mg_apply(h, pc->mg_p, rp, yp);
}

template <int NF, int DIM>
void pc_setup_t(tpb_handle_s* h, const double* J, const double* u, double dt) {
    stage1_setup_t<NF, DIM>(h, J, u, dt);
    stage2_setup_t<NF, DIM>(h, J);
}

template <int NF, int DIM>
void pc_apply_t(tpb_handle_s* h, const double* x, double* y) {
    PcState* pc = h->pc;
    const tpb_solver_opts& o = h->opts;
    const size_t nd = (size_t)NF * h->g.n;
    if (o.stage1 == TPB_S1_NONE && o.stage2 == TPB_S2_NONE) {
        tpb_copy(h, nd, x, y);
        return;
    }
    if (o.stage1 == TPB_S1_NONE) {
        stage2_apply_t<NF, DIM>(h, x, y);
        return;
    }
    stage1_apply_t<NF, DIM>(h, x, y);
    if (o.stage2 == TPB_S2_NONE || o.stage1 == TPB_S1_FIELDSPLIT) return;
    tpb_launch_spmv(h, pc->J, y, pc->t0);   // on slabs with neighbours: NCCL halo exchange of y first
    tpb_axpby(h, nd, 1.0, x, -1.0, pc->t0);  // t0 = x - J y
    stage2_apply_t<NF, DIM>(h, pc->t0, pc->t1);
    tpb_axpy(h, nd, 1.0, pc->t1, y);
}

#define DISPATCH(fn, ...)                         \
    do {                                          \
        if (h->nf == 2) {                         \
            if (h->g.dim == 2)                    \
                fn<2, 2>(__VA_ARGS__);            \
            else                                  \
                fn<2, 3>(__VA_ARGS__);            \
        } else {                                  \
            if (h->g.dim == 2)                    \
                fn<3, 2>(__VA_ARGS__);            \
            else                                  \
                fn<3, 3>(__VA_ARGS__);            \
        }                                         \
    } while (0)

}  // namespace

void tpb_pc_free(tpb_handle_s* h) {
    if (!h->pc) return;
    PcState* pc = h->pc;
    for (int f = 0; f < TPB_MAXF; f++) tpb_dfree(pc->w[f]);
    mg_free(pc->mg_p);
    mg_free(pc->mg_T);
    tpb_dfree(pc->App);
    tpb_dfree(pc->A00);
    tpb_dfree(pc->AT);
    tpb_dfree(pc->Dinv);
    tpb_dfree(pc->Dc);
    tpb_dfree(pc->Lc);
    tpb_dfree(pc->t0);
    tpb_dfree(pc->t1);
    tpb_dfree(pc->t2);
    tpb_dfree(pc->t3);
    tpb_dfree(pc->strength);
    for (auto& it : pc->prog)
        if (it.g) cudaGraphExecDestroy(it.g);
    for (auto& c : pc->cache)
        for (auto& it : c.prog)
            if (it.g) cudaGraphExecDestroy(it.g);
    for (auto g : pc->spare)
        if (g) cudaGraphExecDestroy(g);
    tpb_dfree(pc->gx);
    tpb_dfree(pc->gy);
    delete pc;
    h->pc = nullptr;
}

namespace {

// TPB_GRAPH_NCCL=1: leave the NCCL calls inside the capture (one graph per application) instead of cutting the
// capture around them
bool capture_nccl() {
    static const bool v = getenv("TPB_GRAPH_NCCL") && atoi(getenv("TPB_GRAPH_NCCL")) != 0;
    return v;
}

void seg_begin(tpb_handle_s* h) {
    TPB_CUDA(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    h->pc->capturing = true;
    h->pc->cap_l0 = h->launches;
}

// close the running capture and append it to the program (dropped when it recorded nothing)
void seg_end(tpb_handle_s* h) {
    PcState* pc = h->pc;
    cudaGraph_t graph = nullptr;
    pc->capturing = false;
    TPB_CUDA(cudaStreamEndCapture(h->stream, &graph));
    const int64_t nodes = h->launches - pc->cap_l0;
    h->launches = pc->cap_l0;
    size_t nn = 0;
    TPB_CUDA(cudaGraphGetNodes(graph, nullptr, &nn));
    if (nn == 0) {
        cudaGraphDestroy(graph);
        return;
    }
    cudaGraphExec_t gexec = nullptr;
    if (pc->spare_next < pc->spare.size()) {
        gexec = pc->spare[pc->spare_next];
        pc->spare[pc->spare_next++] = nullptr;
        cudaGraphExecUpdateResultInfo info;
        if (cudaGraphExecUpdate(gexec, graph, &info) != cudaSuccess) {
            cudaGetLastError();
            cudaGraphExecDestroy(gexec);
            gexec = nullptr;
        }
    }
    if (!gexec) TPB_CUDA(cudaGraphInstantiate(&gexec, graph, 0));
    cudaGraphDestroy(graph);
    PcState::Item it;
    it.g = gexec;
    it.nodes = nodes;
    pc->prog.push_back(it);
}

}  // namespace

void tpb_comm_op(tpb_handle_s* h, const std::function<void()>& op) {
    PcState* pc = h->pc;
    if (!pc || !pc->capturing || capture_nccl()) {
        op();
        return;
    }
    seg_end(h);
    PcState::Item it;
    it.op = op;
    pc->prog.push_back(it);
    seg_begin(h);
}

void tpb_pc_setup_impl(tpb_handle_s* h, const double* J, const double* u, double dt) {
    if (!h->pc) {
        h->pc = new PcState();
        const size_t nd = (size_t)h->nf * h->g.n;
        h->pc->t0 = tpb_dalloc<double>(nd);
        h->pc->t1 = tpb_dalloc<double>(nd);
        h->pc->t2 = tpb_dalloc<double>(nd);
        h->pc->t3 = tpb_dalloc<double>(nd);
        h->pc->strength = tpb_dalloc<double>(4);
    }
    h->pc->J = J;
    DISPATCH(pc_setup_t, h, J, u, dt);
    TPB_CUDA(cudaGetLastError());
    h->pc->ready = true;
    // Capture one application into CUDA graphs: the Krylov loop applies the PC 20-30 times per set-up and
    // replaying removes the host-side launch cost of its ~100 small kernels.  TPB_GRAPH=0 disables it.
    PcState* pc = h->pc;
    pc->graph_ok = false;
    static const bool want = !(getenv("TPB_GRAPH") && atoi(getenv("TPB_GRAPH")) == 0);
    const tpb_solver_opts& o = h->opts;
    const bool any = o.stage1 != TPB_S1_NONE || o.stage2 != TPB_S2_NONE;
    // The captured application only depends on pointers, level shapes and sweep counts - not on the operator values.
    // When none of them changed since the last set-up (the usual case from one Newton iteration to the next: the
    // coarsening schedule rarely moves), the executable graphs are kept as they are.
    std::vector<long long> sig;
    auto sig_hier = [&](const MgHier& m, auto&& self) -> void {
        sig.push_back(m.nlev);
        sig.push_back(m.last_sweeps);
        sig.push_back(m.skip_glob ? 1 : 0);
        for (int l = 0; l < m.nlev; l++) {
            const MgLevel& L = m.lev[l];
            for (long long v : {(long long)L.nx, (long long)L.ny, (long long)L.nz, (long long)L.cx, (long long)L.cy,
                                (long long)L.cz, (long long)L.line, (long long)(intptr_t)L.a, (long long)(intptr_t)L.x,
                                (long long)(intptr_t)L.b, (long long)(intptr_t)L.fac, L.fac_cap})
                sig.push_back(v);
        }
        for (long long v : m.gcnt) sig.push_back(v);
        sig.push_back(m.glob ? 1 : 0);
        if (m.glob) self(*m.glob, self);
    };
    sig_hier(pc->mg_p, sig_hier);
    sig_hier(pc->mg_T, sig_hier);
    sig.push_back((long long)(intptr_t)J);
    {
        const unsigned char* ob = reinterpret_cast<const unsigned char*>(&h->opts);
        long long hsh = 1469598103934665603LL;
        for (size_t q = 0; q < sizeof(h->opts); q++) hsh = (hsh ^ ob[q]) * 1099511628211LL;
        sig.push_back(hsh);
    }
    static const bool keep_graph = !(getenv("TPB_GRAPH_KEEP") && atoi(getenv("TPB_GRAPH_KEEP")) == 0);
    if (want && any && keep_graph && !pc->prog.empty() && sig == pc->graph_sig) {
        pc->graph_ok = true;
        return;
    }
    if (want && any && keep_graph) {
        // park the current program under its signature, take the one of the new signature if it was seen before
        constexpr size_t CACHE_MAX = 6;
        if (!pc->prog.empty()) {
            pc->cache.insert(pc->cache.begin(), PcState::Cached{pc->graph_sig, std::move(pc->prog)});
            pc->prog.clear();
            while (pc->cache.size() > CACHE_MAX) {
                for (auto& it : pc->cache.back().prog)
                    if (it.g) cudaGraphExecDestroy(it.g);
                pc->cache.pop_back();
            }
        }
        for (size_t q = 0; q < pc->cache.size(); q++)
            if (pc->cache[q].sig == sig) {
                pc->prog = std::move(pc->cache[q].prog);
                pc->cache.erase(pc->cache.begin() + q);
                pc->graph_sig = sig;
                pc->graph_ok = true;
                return;
            }
    }
    pc->graph_sig = sig;
    if (want && any) {
        const size_t nd = (size_t)h->nf * h->g.n;
        if (!pc->gx) pc->gx = tpb_dalloc<double>(nd);
        if (!pc->gy) pc->gy = tpb_dalloc<double>(nd);
        for (auto g : pc->spare)
            if (g) cudaGraphExecDestroy(g);
        pc->spare.clear();
        for (auto& it : pc->prog)
            if (it.g) pc->spare.push_back(it.g);
        pc->spare_next = 0;
        pc->prog.clear();
        if (capture_nccl() && tpb_comm_size(h) > 1) {
            // NCCL connects its channels lazily: run the application once outside any capture first
            DISPATCH(pc_apply_t, h, pc->gx, pc->gy);
            TPB_CUDA(cudaStreamSynchronize(h->stream));
        }
        seg_begin(h);
        try {
            DISPATCH(pc_apply_t, h, pc->gx, pc->gy);
        } catch (...) {
            cudaGraph_t graph = nullptr;
            pc->capturing = false;
            cudaStreamEndCapture(h->stream, &graph);
            if (graph) cudaGraphDestroy(graph);
            throw;
        }
        seg_end(h);
        for (auto g : pc->spare)
            if (g) cudaGraphExecDestroy(g);
        pc->spare.clear();
        pc->graph_ok = true;
    }
}

void tpb_pc_apply_impl(tpb_handle_s* h, const double* x, double* y) {
    TPB_REQUIRE(h->pc && h->pc->ready, TPB_ERR_STATE, "tpb_pc_apply before tpb_pc_setup");
    PcState* pc = h->pc;
    if (pc->graph_ok) {
        const size_t nd = (size_t)h->nf * h->g.n;
        TPB_CUDA(cudaMemcpyAsync(pc->gx, x, nd * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        for (auto& it : pc->prog) {
            if (it.g) {
                TPB_CUDA(cudaGraphLaunch(it.g, h->stream));
                h->launches += it.nodes;
            } else {
                it.op();
            }
        }
        TPB_CUDA(cudaMemcpyAsync(y, pc->gy, nd * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        return;
    }
    DISPATCH(pc_apply_t, h, x, y);
    TPB_CUDA(cudaGetLastError());
}

// ---- introspection for the component parity tests (tests/ compare against oracle/cport) ------------
int tpb_pc_mg_nlevels_impl(tpb_handle_s* h, int which) {
    if (!h->pc) return 0;
    return which == 0 ? h->pc->mg_p.nlev : h->pc->mg_T.nlev;
}
int tpb_pc_mg_level_impl(tpb_handle_s* h, int which, int l, int* dims6, double* op_out) {
    if (!h->pc) return -1;
    MgHier& m = which == 0 ? h->pc->mg_p : h->pc->mg_T;
    if (l < 0 || l >= m.nlev) return -1;
    const MgLevel& L = m.lev[l];
    if (dims6) {
        dims6[0] = L.nx;
        dims6[1] = L.ny;
        dims6[2] = L.nz;
        dims6[3] = L.cx;
        dims6[4] = L.cy;
        dims6[5] = L.cz;
    }
    if (op_out) {
        TPB_CUDA(cudaMemcpyAsync(op_out, L.a, (size_t)h->ns * L.n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        TPB_CUDA(cudaStreamSynchronize(h->stream));
    }
    return 0;
}
void tpb_pc_mg_apply_impl(tpb_handle_s* h, int which, const double* b, double* y) {
    TPB_REQUIRE(h->pc && h->pc->ready, TPB_ERR_STATE, "multigrid not set up");
    MgHier& m = which == 0 ? h->pc->mg_p : h->pc->mg_T;
    TPB_REQUIRE(m.nlev > 0, TPB_ERR_STATE, "this multigrid hierarchy is not part of the selected PC");
    mg_apply(h, m, b, y);
}
void tpb_pc_stage2_apply_impl(tpb_handle_s* h, const double* r, double* z) {
    TPB_REQUIRE(h->pc && h->pc->ready && h->pc->Dinv, TPB_ERR_STATE, "stage 2 not set up");
    if (h->nf == 2) {
        if (h->g.dim == 2) stage2_apply_t<2, 2>(h, r, z); else stage2_apply_t<2, 3>(h, r, z);
    } else {
        if (h->g.dim == 2) stage2_apply_t<3, 2>(h, r, z); else stage2_apply_t<3, 3>(h, r, z);
    }
}
const double* tpb_pc_weights_impl(tpb_handle_s* h, int f) { return h->pc ? h->pc->w[f] : nullptr; }

// one smoothing launch of the fine-level pressure smoother on its own (bench.py roofline of the dominant kernel):
// a group of mg_tile_sweeps sweeps of the hybrid line smoother, or one colour pass of the point smoother; returns the
// number of cells it updates
long long tpb_pc_rbgs_pass_impl(tpb_handle_s* h, int col) {
    TPB_REQUIRE(h->pc && h->pc->ready && h->pc->mg_p.nlev > 0, TPB_ERR_STATE, "pressure multigrid not set up");
    const MgLevel& L = h->pc->mg_p.lev[0];
    LevGeom g = lg(L);
    if (L.line) {
        const int group = h->opts.mg_tile_sweeps > 0 ? h->opts.mg_tile_sweeps : std::max(1, h->opts.mg_pre);
        int tx, ty;
        line_tile_shape(L.nz, tx, ty);
        const unsigned grid = (unsigned)(((L.nx + tx - 1) / tx) * ((L.ny + ty - 1) / ty));
        const size_t smem = line_smem(L.nz);
        want_smem(line_smooth_kernel<false>, smem);
        launch_pdl_smem(line_smooth_kernel<false>, grid, LS_THREADS, smem, h->stream, L.a, L.fac, L.b, L.xs0(), L.xs1(), g,
                        nullptr, 0, 0, 0.0, group, tx, ty);
        h->launches++;
        return L.n;
    }
    long long threads = (long long)L.ny * L.nz * ((L.nx + 1) >> 1);
    if (h->ns == 7)
        rbgs_kernel<7, false><<<nblk(threads, 256), 256, 0, h->stream>>>(L.a, L.b, L.x, g, col, 0, nullptr, 0, 0, 0.0);
    else
        rbgs_kernel<5, false><<<nblk(threads, 256), 256, 0, h->stream>>>(L.a, L.b, L.x, g, col, 0, nullptr, 0, 0, 0.0);
    h->launches++;
    return threads;
}

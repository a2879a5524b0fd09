"""CPU oracle for the thermalporous hot path (TEST INFRASTRUCTURE - never shipped).

A NumPy fp64 restatement of the discrete equations the reference hands to
Firedrake: the DG0 / two-point-flux residual of single- and two-phase
non-isothermal flow and its exact Jacobian.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
leg may import this module; the product path (thermalporous_b200) never does.

PARITY STATUS: the reference (Firedrake/PETSc, un-vendored, unpinned) cannot be
run in this image, and its repository holds no golden vectors.  The oracle is
therefore pinned against the reference's OWN form-building code
(`singlephase.py:60-273`, `twophase.py:67-411`, `wellcase.py`, `heatercase.py`,
`sourceterms.py`, `physicalparameters.py`) executed through a minimal DG0
evaluator (`tests/golden/fd_shim`, script `tests/golden/make_golden.py`);
the resulting vectors live in `tests/golden/*.npz`.  What stays unpinned is
Firedrake's own facet orientation on extruded meshes ('+' = lower cell), which
SURVEY.md appendix item 5 argues from the hydrostatic-column identity.

The residual is written face-array-wise (vectorised over whole families of
faces) and the Jacobian is obtained by complex-step differentiation of that
residual with a 7-colour distance-2 colouring of the stencil, so it is the
exact derivative with the upwind / rate-cap conditionals frozen - the same thing
UFL's `derivative` produces - and shares no hand-derived formula with the CUDA
kernels it checks.

Layouts (shared with the C-ABI, include/tpb200.h):
  cell index      c = i + nx*(j + ny*k)
  state           u[f, c], f in (p, T) or (p, T, S_o)
  Jacobian        J[s, r, c_, cell], s in (diag, x-, x+, y-, y+, z-, z+)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

E = 2.718281828459045  # UFL `e`; generated code evaluates pow(e, x)

PROD, INJ, HEATER = 0, 1, 2


@dataclass
class Params:
    """physicalparameters.py:9-35 (class attributes of PhysicalParameters)."""
    ko: float = 0.15
    kw: float = 0.6005638
    kr: float = 1.7295772056
    c_v_w: float = 4181.3
    c_v_o: float = 2093.4
    c_r: float = 920.0
    rho_r: float = 2650.0
    p_inj: float = 6.895e7 * 1e-6
    p_prod: float = 2.7579e7 * 1e-6
    T_inj: float = 422.039
    T_prod: float = 288.706
    API: float = 10.0
    p_ref: float = 4.1369e7 * 1e-6
    g: float = 9.80665 * 1e-6
    S_o: float = 1.0
    U: float = 5.44409e6
    rate: float = 1.8e-3
    well_radius: float = 0.1


# ---------------------------------------------------------------------------
# properties, physicalparameters.py:37-98
# ---------------------------------------------------------------------------
def oil_rho(prm, p, T):
    """physicalparameters.py:37-46."""
    SG = 141.5 / (prm.API + 131.5)
    rho_ref = SG * 999.0
    c = 5.5e-5
    p0 = 1.01325
    e1 = 2.5e-4
    T0 = 15.5556 + 273.15
    pbar = p * 1e1
    return rho_ref * np.power(E, c * (pbar - p0)) * np.power(E, -e1 * (T - T0))


def oil_mu(prm, T):
    """physicalparameters.py:48-57."""
    A1, A2, A3, A4 = -0.8021, 23.8765, 0.31458, -9.21592
    Tf = 1.8 * (T - 273.15) + 32.0
    return 1e-3 * (10.0 ** (A1 * prm.API + A2) * np.power(Tf, A3 * prm.API + A4))


def water_rho(prm, p, T):
    """physicalparameters.py:69-82 (note Tc = T - 272.15)."""
    E_0, E_1, E_2 = 999.83952, 16.955176, -7.987e-3
    E_3, E_4, E_5 = -46.170461e-6, 105.56302e-9, -280.54353e-12
    E_6, E_7 = 16.87985e-3, 10.2
    Cw = 3.98854e-4
    Tc = T - 272.15
    poly = E_0 + E_1 * Tc + E_2 * Tc ** 2 + E_3 * Tc ** 3 + E_4 * Tc ** 4 + E_5 * Tc ** 5
    return poly * np.power(E, Cw * (p - E_7)) / (1 + E_6 * Tc)


def water_mu(prm, T):
    """physicalparameters.py:84-90 (note Tf uses 272.15)."""
    Aw, Bw, Cw = 2.1850, 0.04012, 5.1547e-6
    Tf = 1.8 * (T - 272.15) + 32
    return 1e-3 * Aw / (-1 + Bw * Tf + Cw * Tf ** 2)


# ---------------------------------------------------------------------------
@dataclass
class Grid:
    nx: int
    ny: int
    nz: int = 1
    dx: float = 1.0
    dy: float = 1.0
    dz: float = 1.0
    dim: int = 3

    @property
    def n(self):
        return self.nx * self.ny * self.nz

    @property
    def shape(self):
        return (self.nz, self.ny, self.nx)

    @property
    def vol(self):
        return self.dx * self.dy * (self.dz if self.dim == 3 else 1.0)

    def face_area(self, axis):
        """axis 0:x 1:y 2:z.  2-D facets are edges (rectanglegeo.py:28-34)."""
        if self.dim == 2:
            return (self.dy, self.dx)[axis]
        return (self.dy * self.dz, self.dx * self.dz, self.dx * self.dy)[axis]

    def h(self, axis):
        return (self.dx, self.dy, self.dz)[axis]


@dataclass
class Source:
    """One (cell, kind) source entry; weight = V_cell * delta(cell)."""
    cell: int
    kind: int
    weight: float = 1.0
    bhp: float = 0.0
    max_rate: float = 0.0
    const_rate: bool = False


@dataclass
class Problem:
    grid: Grid
    nphase: int
    prm: Params
    phi: np.ndarray
    Kx: np.ndarray
    Ky: np.ndarray
    Kz: np.ndarray | None = None
    kT: np.ndarray | None = None          # single-phase only (geo.kT)
    sources: list = field(default_factory=list)
    gravity: bool = True                  # 3-D forms always carry g (twophase.py:317)

    @property
    def nf(self):
        return 2 if self.nphase == 1 else 3


def _harm(a, b):
    """conditional(gt(avg(K),0), K('+')*K('-')/avg(K), 0)  (singlephase.py:98)."""
    s = 0.5 * (a + b)
    with np.errstate(divide="ignore", invalid="ignore"):
        out = np.where(np.real(s) > 0.0, a * b / np.where(np.real(s) > 0.0, s, 1.0), 0.0)
    return out


def _sl(axis, lo):
    """slices picking the '+'(lower-index) / '-'(higher-index) cell of each face."""
    ax = 2 - axis  # array axis in (nz, ny, nx)
    s = [slice(None)] * 3
    s[ax] = slice(None, -1) if lo else slice(1, None)
    return tuple(s)


def _peaceman_wi(Kx, Ky):
    """wellcase.py:180-192: h=5, rw=0.1, Dx=Dy=5 hard-wired."""
    h, rw, Dx, Dy = 5.0, 0.1, 5.0, 5.0
    ro = 0.28 * ((Ky / Kx) ** 0.5 * Dx ** 2 + (Kx / Ky) ** 0.5 * Dy ** 2) ** 0.5 / (
        (Ky / Kx) ** 0.25 + (Kx / Ky) ** 0.25)
    Ke = (Kx * Ky) ** 0.5
    return 2 * math.pi * h * Ke / math.log(ro / rw)


def _cabs(z):
    """|z| that stays analytic for the complex step (sign taken from the real part)."""
    return np.where(np.real(z) < 0.0, -z, z)


def _rate(src, wi_over_mu, p):
    """wellcase.py:191-199 draw-down clipping and rate cap."""
    if src.const_rate:
        return src.max_rate + 0.0 * p
    d = src.bhp - p
    if src.max_rate < 0.0:
        dd = np.where(np.real(d) >= 0.0, 0.0, d)
    else:
        dd = np.where(np.real(d) <= 0.0, 0.0, d)
    rate = wi_over_mu * dd
    return np.where(np.real(_cabs(rate)) - abs(src.max_rate) >= 0.0, src.max_rate, rate)


def residual(pb: Problem, u, u_old, dt):
    """F(u; u_old, dt) as the (nf, N) array Firedrake's assemble(F) would hold.

    Single-phase: singlephase.py:120-127 (2-D), :226-235 (3-D).
    Two-phase:    twophase.py:162-178 (2-D), :333-354 (3-D).
    Sources:      singlephase.py:151-165, twophase.py:388-411.
    """
    g, prm = pb.grid, pb.prm
    nf = pb.nf
    shp = g.shape
    u = np.asarray(u).reshape(nf, *shp)
    uo = np.asarray(u_old).reshape(nf, *shp)
    dtype = np.result_type(u.dtype, uo.dtype, np.float64)
    F = np.zeros((nf, *shp), dtype=dtype)
    phi = pb.phi.reshape(shp)
    V = g.vol
    p, T = u[0], u[1]
    p_, T_ = uo[0], uo[1]
    K = [pb.Kx.reshape(shp), pb.Ky.reshape(shp)]
    if g.dim == 3:
        K.append(pb.Kz.reshape(shp))

    if pb.nphase == 1:
        c_v = prm.c_v_o
        rho = oil_rho(prm, p, T)
        rho_ = oil_rho(prm, p_, T_)
        mob = rho / oil_mu(prm, T)                      # rho_o/mu_o
        kT = pb.kT.reshape(shp)
        F[0] += V * phi * (rho - rho_) / dt
        F[1] += V * (phi * c_v * (rho * T - rho_ * T_) + (1 - phi) * prm.rho_r * prm.c_r * (T - T_)) / dt
        for ax in range(g.dim):
            if shp[2 - ax] < 2:
                continue
            P, M = _sl(ax, True), _sl(ax, False)
            A, h = g.face_area(ax), g.h(ax)
            Kf = _harm(K[ax][P], K[ax][M])
            flow = (p[P] - p[M]) / h
            if ax == 2 and pb.gravity:
                flow = flow - prm.g * 0.5 * (rho[P] + rho[M])
            up = np.real(flow) > 0.0
            lam = np.where(up, mob[P], mob[M])
            Tlam = np.where(up, T[P] * mob[P], T[M] * mob[M])
            fm = A * Kf * lam * flow
            fe = A * Kf * Tlam * c_v * flow + A * _harm(kT[P], kT[M]) * (T[P] - T[M]) / h
            F[0][P] += fm
            F[0][M] -= fm
            F[1][P] += fe
            F[1][M] -= fe
    else:
        S, S_ = u[2], uo[2]
        cw, co = prm.c_v_w, prm.c_v_o
        Wp = prm.T_prod
        Wo = prm.T_prod * (cw * (1 - prm.S_o) + co * prm.S_o)
        ro, rw = oil_rho(prm, p, T), water_rho(prm, p, T)
        ro_, rw_ = oil_rho(prm, p_, T_), water_rho(prm, p_, T_)
        lo = S * ro / oil_mu(prm, T)
        lw = (1.0 - S) * rw / water_mu(prm, T)
        kT = phi * (S * prm.ko + (1 - S) * prm.kw) + (1 - phi) * prm.kr
        acc_w = phi * (rw * (1.0 - S) - rw_ * (1.0 - S_)) / dt
        acc_o = phi * (ro * S - ro_ * S_) / dt
        F[0] += V * Wp * (cw * acc_w + co * acc_o)
        F[2] += V * Wo * acc_o
        F[1] += V * (phi * cw * (rw * (1.0 - S) * T - rw_ * (1.0 - S_) * T_) / dt
                     + phi * co * (ro * S * T - ro_ * S_ * T_) / dt
                     + (1 - phi) * prm.rho_r * prm.c_r * (T - T_) / dt)
        for ax in range(g.dim):
            if shp[2 - ax] < 2:
                continue
            P, M = _sl(ax, True), _sl(ax, False)
            A, h = g.face_area(ax), g.h(ax)
            Kf = _harm(K[ax][P], K[ax][M])
            dp = (p[P] - p[M]) / h
            if ax == 2 and pb.gravity:
                fl_w = dp - prm.g * 0.5 * (rw[P] + rw[M])
                fl_o = dp - prm.g * 0.5 * (ro[P] + ro[M])
            else:
                fl_w = dp
                fl_o = dp
            upw = np.real(fl_w) > 0.0
            upo = np.real(fl_o) > 0.0
            fw = A * Kf * np.where(upw, lw[P], lw[M]) * fl_w
            fo = A * Kf * np.where(upo, lo[P], lo[M]) * fl_o
            few = A * Kf * np.where(upw, T[P] * lw[P], T[M] * lw[M]) * cw * fl_w
            feo = A * Kf * np.where(upo, T[P] * lo[P], T[M] * lo[M]) * co * fl_o
            fd = A * _harm(kT[P], kT[M]) * (T[P] - T[M]) / h
            fp = Wp * (cw * fw + co * fo)
            fs = Wo * fo
            fe = few + feo + fd
            F[0][P] += fp
            F[0][M] -= fp
            F[1][P] += fe
            F[1][M] -= fe
            F[2][P] += fs
            F[2][M] -= fs

    F = F.reshape(nf, -1)
    uf = u.reshape(nf, -1)
    Kx, Ky = pb.Kx.ravel(), pb.Ky.ravel()
    for s in pb.sources:
        c, w = s.cell, s.weight
        pc, Tc = uf[0, c], uf[1, c]
        if s.kind == HEATER:
            F[1, c] -= w * prm.U * (prm.T_inj - Tc)
            continue
        wi = _peaceman_wi(Kx[c], Ky[c])
        if pb.nphase == 1:
            q = _rate(s, wi / oil_mu(prm, Tc), pc)
            if s.kind == PROD:
                r = oil_rho(prm, pc, Tc)
                F[0, c] -= w * r * q
                F[1, c] -= w * r * q * prm.c_v_o * Tc
            else:
                r = oil_rho(prm, pc, prm.T_inj)
                F[0, c] -= w * r * q
                F[1, c] -= w * r * q * prm.c_v_o * prm.T_inj
        else:
            Sc = uf[2, c]
            cw, co = prm.c_v_w, prm.c_v_o
            Wp = prm.T_prod
            Wo = prm.T_prod * (cw * (1 - prm.S_o) + co * prm.S_o)
            if s.kind == PROD:
                muo, muw = oil_mu(prm, Tc), water_mu(prm, Tc)
                mu = 1.0 / (Sc / muo + (1.0 - Sc) / muw)
                q = _rate(s, wi / mu, pc)
                qw = (1 - Sc) / muw * mu * q
                qo = Sc / muo * mu * q
                r_o, r_w = oil_rho(prm, pc, Tc), water_rho(prm, pc, Tc)
                F[0, c] -= Wp * w * (cw * r_w * qw + co * r_o * qo)
                F[2, c] -= Wo * w * r_o * qo
                F[1, c] -= w * (r_w * qw * cw + r_o * qo * co) * Tc
            else:
                q = _rate(s, wi / water_mu(prm, Tc), pc)
                r = water_rho(prm, pc, prm.T_inj)
                F[0, c] -= Wp * cw * r * q * w
                F[1, c] -= r * q * cw * prm.T_inj * w
    return F


NSTENCIL = {2: 5, 3: 7}


def stencil_offsets(g: Grid):
    """(di, dj, dk) of each stencil slot: diag, x-, x+, y-, y+[, z-, z+]."""
    offs = [(0, 0, 0), (-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0)]
    if g.dim == 3:
        offs += [(0, 0, -1), (0, 0, 1)]
    return offs


def jacobian(pb: Problem, u, u_old, dt, h=1e-30):
    """dF/du by complex step with a distance-2 colouring -> J[s, r, c, cell].

    Equals UFL's derivative(F, u) (thermalmodel.py:36): every conditional keeps
    the branch it has at `u` because branches are chosen on real parts.
    """
    g = pb.grid
    nf = pb.nf
    ns = NSTENCIL[g.dim]
    N = g.n
    u = np.asarray(u, dtype=np.float64).reshape(nf, N)
    kk, jj, ii = np.meshgrid(np.arange(g.nz), np.arange(g.ny), np.arange(g.nx), indexing="ij")
    colour = ((ii + 2 * jj + 3 * kk) % 7).ravel()
    offs = stencil_offsets(g)
    J = np.zeros((ns, nf, nf, N))
    ii, jj, kk = ii.ravel(), jj.ravel(), kk.ravel()
    for col in range(7):
        mask = colour == col
        if not mask.any():
            continue
        for c in range(nf):
            up = u.astype(np.complex128)
            up[c, mask] += 1j * h
            dF = np.imag(residual(pb, up, u_old, dt)) / h      # (nf, N)
            # row cell `cell`, slot s looks at column cell = cell + off; that column
            # must be a perturbed one
            for s, (di, dj, dk) in enumerate(offs):
                ni, nj, nk = ii + di, jj + dj, kk + dk
                ok = (ni >= 0) & (ni < g.nx) & (nj >= 0) & (nj < g.ny) & (nk >= 0) & (nk < g.nz)
                nb = np.where(ok, ni + g.nx * (nj + g.ny * nk), 0)
                sel = ok & mask[nb]
                J[s, :, c, sel] = dF[:, sel].T
    return J


def to_csr(J, g: Grid, ordering="field"):
    """Block-stencil J -> scipy CSR.  ordering 'field': row = f*N + cell (Firedrake's
    mixed-space `aij` layout, preconditioners.py:356-361); 'cell': row = cell*nf + f."""
    import scipy.sparse as sp
    ns, nf, _, N = J.shape
    offs = stencil_offsets(g)
    cells = np.arange(N)
    ii = cells % g.nx
    jj = (cells // g.nx) % g.ny
    kk = cells // (g.nx * g.ny)
    rows, cols, vals = [], [], []
    for s, (di, dj, dk) in enumerate(offs):
        ni, nj, nk = ii + di, jj + dj, kk + dk
        ok = (ni >= 0) & (ni < g.nx) & (nj >= 0) & (nj < g.ny) & (nk >= 0) & (nk < g.nz)
        nb = (ni + g.nx * (nj + g.ny * nk))[ok]
        me = cells[ok]
        for r in range(nf):
            for c in range(nf):
                if ordering == "field":
                    rows.append(r * N + me)
                    cols.append(c * N + nb)
                else:
                    rows.append(me * nf + r)
                    cols.append(nb * nf + c)
                vals.append(J[s, r, c, ok])
    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                      shape=(nf * N, nf * N)).tocsr()
    return A


def spmv(J, g: Grid, x):
    """y = J x on the block-stencil layout (reference: PETSc MatMult on the aij Jacobian)."""
    ns, nf, _, N = J.shape
    x = np.asarray(x).reshape(nf, *g.shape)
    y = np.zeros((nf, N))
    for s, (di, dj, dk) in enumerate(stencil_offsets(g)):
        xs = np.zeros_like(x)
        src = [slice(None)] * 3
        dst = [slice(None)] * 3
        for ax, d in ((2, di), (1, dj), (0, dk)):
            if d == 1:
                src[ax], dst[ax] = slice(1, None), slice(None, -1)
            elif d == -1:
                src[ax], dst[ax] = slice(None, -1), slice(1, None)
        xs[(slice(None), *dst)] = x[(slice(None), *src)]
        xs = xs.reshape(nf, N)
        for r in range(nf):
            for c in range(nf):
                y[r] += J[s, r, c] * xs[c]
    return y


def newton_solve(pb: Problem, u0, u_old, dt, rtol=1e-12, max_it=25, verbose=False):
    """Newton with a sparse direct solve (reference: SNES newtonls driven by
    thermalmodel.py:165; here converged far below the reference's rtol 1e-8 so
    converged fields can be compared at 1e-8).  Returns (u, nits, converged)."""
    import scipy.sparse.linalg as spla
    g = pb.grid
    nf = pb.nf
    u = np.array(u0, dtype=np.float64).reshape(nf, g.n).copy()
    F = residual(pb, u, u_old, dt)
    f0 = np.linalg.norm(F)
    for it in range(max_it):
        fn = np.linalg.norm(F)
        if verbose:
            print(f"  oracle newton {it}: |F| = {fn:.6e}")
        if fn <= rtol * f0 or fn < 1e-300:
            return u, it, True
        A = to_csr(jacobian(pb, u, u_old, dt), g, "cell")
        d = spla.spsolve(A.tocsc(), -F.T.ravel())
        du = d.reshape(g.n, nf).T
        u = u + du
        F = residual(pb, u, u_old, dt)
        if np.linalg.norm(du) <= 1e-14 * np.linalg.norm(u):
            return u, it + 1, True
    return u, max_it, np.linalg.norm(F) <= rtol * f0

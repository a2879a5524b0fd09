"""Well / heater source cases - host-side mirror of wellcase.py, heatercase.py,
wellheatercase.py and sourceterms.py of the reference.

Same constructor arguments, presets and attribute names (prod_wells / inj_wells / heaters lists of
dicts with 'name', 'bhp', 'location', 'delta', 'max_rate'; deltas_prod / deltas_inj / deltas_heaters
summed fields for SourceTerms).  A 'delta' here is a NumPy array over the cells (the reference's
DG0 Function values); `source_entries()` flattens a case into the (cell, kind, weight, bhp,
max_rate, const_rate) records libtpb200 takes (include/tpb200.h: tpb_source), weight = V_cell*delta.

Documented deviation: when no cell centre lies inside the 0.1 m bump the
reference falls back to the nearest cell with ties broken by Firedrake's dof numbering
(utils.py:7-25); here ties go to the lowest cell index.
"""
from __future__ import annotations

import numpy as np

PROD, INJ, HEATER = 0, 1, 2


class _DeltaMixin:
    _height = 1.0   # wellcase.py:146, heatercase.py:99;  sourceterms.py:116 uses 0.1

    def _radius(self):
        return getattr(self.params, "well_radius", 0.1)

    def _centres(self):
        if getattr(self, "_cc", None) is None:
            self._cc = self.geo.cell_centres()
        return self._cc

    def well_delta(self, w):
        """wellcase.py:157-169: 1/V at the cell whose centre is closest to w."""
        cc = self._centres()
        d = np.linalg.norm(cc - np.asarray(w, dtype=np.float64)[None, :self.geo.dim], axis=1)
        delta = np.zeros(self.geo.ncell)
        delta[int(np.argmin(d))] = 1.0 / self.geo.cell_volume
        return delta

    def _bump(self, w, radius):
        cc = self._centres()
        r2 = (cc[:, 0] - w[0]) ** 2 + (cc[:, 1] - w[1]) ** 2
        inside = r2 < radius ** 2
        if self.geo.dim == 3:
            inside &= np.abs(cc[:, 2] - w[2]) < self._height
        delta = np.zeros(self.geo.ncell)
        if inside.any():
            delta[inside] = np.exp(-(1.0 / (-r2[inside] + radius ** 2)))
        return delta

    def well_circle(self, w, radius=None):
        """wellcase.py:110-123 / :141-155: C-infinity bump at the cell centres, normalised to
        integrate to 1; nearest-cell delta when it integrates to 0."""
        delta = self._bump(w, self._radius() if radius is None else radius)
        normalise = delta.sum() * self.geo.cell_volume
        if normalise == 0:
            return self.well_delta(w)
        return delta / normalise

    well_circle3D = well_circle


_PRESETS_2D = {
    "default": lambda s: ([[0.2 * s.Length, s.Length_y / 2]], [[0.8 * s.Length, s.Length_y / 2]]),
    "SPE10_60x120": lambda s: ([[140.0, 210.0]], [[265.0, 260.0]]),                       # wellcase.py:30-36
    "test0": lambda s: ([[2., s.Length_y / 4.], [2., s.Length_y / 2.], [2., 3. * s.Length_y / 4]],
                        [[s.Length - 2., s.Length_y / 4.], [s.Length - 2., s.Length_y / 2.],
                         [s.Length - 2., 3. * s.Length_y / 4]]),                           # :37-41
    "test": lambda s: ([[10., f * s.Length_y] for f in (0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9)],
                       [[s.Length - 10., f * s.Length_y] for f in (0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9)]),
    "SPE10_40x40": lambda s: ([[2 * s.geo.Dx, s.Length_y / 4.], [2 * s.geo.Dx, s.Length_y / 2.],
                               [2 * s.geo.Dx, 3. * s.Length_y / 4]],
                              [[s.Length - 2 * s.geo.Dx, s.Length_y / 4.], [s.Length - 2 * s.geo.Dx, s.Length_y / 2.],
                               [s.Length - 2 * s.geo.Dx, 3. * s.Length_y / 4]]),
}


def _large3d(s, zf):
    """wellcase.py:58-64: 3 rows of 7 points; the last point of the third row repeats (7Lx/8, Ly/4)."""
    Lx, Ly, Lz = s.Length, s.Length_y, s.Length_z
    xs = [Lx / 8, Lx / 4, 3 * Lx / 8, Lx / 2, 5 * Lx / 8, 3 * Lx / 4, 7 * Lx / 8]
    pts = [[x, Ly / 2, Lz * zf] for x in xs] + [[x, Ly / 4, Lz * zf] for x in xs] + \
          [[x, 3 * Ly / 4, Lz * zf] for x in xs[:-1]] + [[7 * Lx / 8, Ly / 4, Lz * zf]]
    return pts


_PRESETS_3D = {
    "default": lambda s: ([[s.Length / 2, s.Length_y / 2, s.Length_z * 0.2]],
                          [[s.Length / 2, s.Length_y / 2, s.Length_z * 0.8]]),             # wellcase.py:55-57
    "large": lambda s: (_large3d(s, 0.2), _large3d(s, 0.8)),
}


class _CaseBase(_DeltaMixin):
    def _init_geo(self, params, geo):
        self.Length = geo.Length
        self.Length_y = geo.Length_y
        if geo.dim == 3:
            self.Length_z = geo.Length_z
        self.geo = geo
        self.params = params
        self._cc = None

    def _preset(self, well_case, prod_points, inj_points):
        if well_case is not None:
            table = _PRESETS_2D if self.geo.dim == 2 else _PRESETS_3D
            if well_case not in table:
                raise KeyError("unknown well_case %r for a %d-D geo" % (well_case, self.geo.dim))
            prod_points, inj_points = table[well_case](self)
        return list(prod_points or []), list(inj_points or [])


class WellCase(_CaseBase):
    """wellcase.py:6-108."""

    def __init__(self, params, geo, well_case=None, prod_points=None, inj_points=None, constant_rate=False):
        self.name = "Wells"
        self._init_geo(params, geo)
        self.wellfunc = "circle"
        self.constant_rate = bool(constant_rate)
        prod_points, inj_points = self._preset(well_case, prod_points, inj_points)
        self.init_wells(prod_points, inj_points, self.wellfunc)

    def init_wells(self, prod_points, inj_points, wellfunc):
        self.prod_wells, self.inj_wells = [], []
        self.prodcount = self.injcount = 0
        for point in prod_points:
            self.prod_wells.append(self.make_well(point, wellfunc, "prod"))
        for point in inj_points:
            self.inj_wells.append(self.make_well(point, wellfunc, "inj"))

    def make_well(self, w, wellfunc, welltype):
        rate, p_inj, p_prod = self.params.rate, self.params.p_inj, self.params.p_prod
        delta = self.well_delta(w) if wellfunc == "delta" else self.well_circle(w)
        if welltype == "prod":
            name, bhp, max_rate = "prod" + str(self.prodcount), p_prod, -rate      # wellcase.py:97-101
            self.prodcount += 1
        else:
            name, bhp, max_rate = "inj" + str(self.injcount), p_inj, rate
            self.injcount += 1
        return {"name": name, "bhp": bhp, "location": w, "delta": delta, "max_rate": max_rate, "rate": 0.0}


class HeaterCase(_CaseBase):
    """heatercase.py:6-118."""

    def __init__(self, params, geo, well_case=None, heater_points=()):
        self.name = "Heaters"
        self._init_geo(params, geo)
        self.constant_rate = False
        heater_points = list(heater_points)
        if well_case is not None:
            # heatercase.py:18-51: the well presets, heaters at producers + injectors
            # exactly the names heatercase.py handles: 2-D SPE10_60x120 / test0 / test ("default" sets only the unused
            # prod/inj points there), 3-D default / multiple; any other name leaves heater_points untouched
            if geo.dim == 2:
                table = {k: _PRESETS_2D[k] for k in ("SPE10_60x120", "test0", "test")}
            else:
                table = {"default": _PRESETS_3D["default"], "multiple": lambda s: (
                    [[s.Length / 4, s.Length_y / 2, s.Length_z * 0.2], [s.Length / 2, s.Length_y / 2, s.Length_z * 0.2],
                     [3 * s.Length / 4, s.Length_y / 2, s.Length_z * 0.2]],
                    [[s.Length / 4, s.Length_y / 2, s.Length_z * 0.8], [s.Length / 2, s.Length_y / 2, s.Length_z * 0.8],
                     [3 * s.Length / 4, s.Length_y / 2, s.Length_z * 0.8]])}
            if well_case in table:
                pp, ip = table[well_case](self)
                heater_points = pp + ip
        self.init_heaters(heater_points, "circle")

    def _radius(self):
        return 0.1     # heatercase.py:84,98: hard-wired, not params.well_radius

    def init_heaters(self, heater_points, wellfunc):
        self.heaters = []
        self.heatercount = 0
        for point in heater_points:
            self.heaters.append(self.make_heater(point, wellfunc))

    def make_heater(self, w, wellfunc):
        delta = self.well_delta(w) if wellfunc == "delta" else self.well_circle(w)
        name = "heater" + str(self.heatercount)
        self.heatercount += 1
        return {"name": name, "location": w, "delta": delta}


class WellHeaterCase(WellCase, HeaterCase):
    """wellheatercase.py:6-12: heaters at the injector and producer points (in that order); the
    delta functions are WellCase's (method resolution order), i.e. radius = params.well_radius."""

    def __init__(self, params, geo, well_case=None, prod_points=(), inj_points=(), constant_rate=False):
        WellCase.__init__(self, params, geo, well_case=well_case, prod_points=prod_points, inj_points=inj_points,
                          constant_rate=constant_rate)
        const = self.constant_rate
        HeaterCase.__init__(self, params, geo, well_case=well_case, heater_points=list(inj_points) + list(prod_points))
        self.constant_rate = const
        self.name = "Wells and Heaters"

    def _radius(self):
        return getattr(self.params, "well_radius", 0.1)


class SourceTerms(_CaseBase):
    """sourceterms.py:6-153: one summed delta field per kind (the form for many wells,
    thermalmodel.py:247)."""
    _height = 0.1   # sourceterms.py:116

    def __init__(self, params, geo, well_case=None, prod_points=(), inj_points=(), heater_points=(),
                 constant_rate=False):
        self.name = "Sources"
        self._init_geo(params, geo)
        if not hasattr(params, "prod_rate"):
            params.prod_rate = params.rate      # sourceterms.py:18-25
        if not hasattr(params, "inj_rate"):
            params.inj_rate = params.rate
        self.constant_rate = bool(constant_rate)
        prod_points, inj_points = self._preset(well_case, prod_points, inj_points)
        self.init_deltas(prod_points, inj_points, list(heater_points), "circle")

    def init_deltas(self, prod_points, inj_points, heater_points, well_func):
        deltas = self.make_deltas if well_func == "delta" else self.make_circles
        self.deltas_prod = deltas(prod_points)
        self.deltas_inj = deltas(inj_points)
        self.deltas_heaters = deltas(heater_points)

    def make_circles(self, ws):
        out = np.zeros(self.geo.ncell)
        for w in ws:
            out += self.well_circle(w)
        return out

    def make_deltas(self, ws):
        """sourceterms.py:126-139: coincident points do NOT accumulate here (vec[node] = 1.0)."""
        out = np.zeros(self.geo.ncell)
        for w in ws:
            out[np.nonzero(self.well_delta(w))[0]] = 1.0 / self.geo.cell_volume
        return out


class MixingCase:
    """mixingcase.py:3-47: no sources, a non-uniform initial condition (`init_IC`) instead - cold/hot halves in y
    ("coldandhot") or z ("coldonhot", 3-D only), or oil over water in the upper/lower half ("heavyonlight",
    two-phase only; sets API = 40).  `init_IC` returns the (nf, ncell) field array in the C-ABI cell order."""

    def __init__(self, params, geo, mixing_case="coldandhot"):
        self.name = "Mixing"
        self.geo, self.params, self.mixing_case = geo, params, mixing_case
        if mixing_case not in ("coldonhot", "coldandhot", "heavyonlight"):
            raise SystemExit("Error: Undefined mixing_case: %s" % mixing_case)                  # mixingcase.py:11-12
        if mixing_case == "coldonhot" and geo.dim == 2:
            print("Warning: coldonhot only implemented for 3D cases. Switching to coldandhot")  # :13-15
            self.mixing_case = "coldandhot"
        if mixing_case == "heavyonlight":
            self.params.API = 40                                                               # :16-18

    def init_IC(self, phases="Single phase"):
        geo, prm = self.geo, self.params
        cc = geo.cell_centres()
        n = geo.ncell
        p = np.full(n, float(prm.p_ref))
        T = np.full(n, float(prm.T_prod))
        S = None
        if self.mixing_case == "coldandhot":                                                   # :27-30
            T = np.where(cc[:, 1] > geo.Length_y / 2.0, prm.T_prod, prm.T_inj).astype(np.float64)
        elif self.mixing_case == "coldonhot":                                                  # :31-34
            T = np.where(cc[:, 2] > geo.Length_z / 2.0, prm.T_prod, prm.T_inj).astype(np.float64)
        else:                                                                                  # :35-43
            if phases == "Single phase":
                raise SystemExit("Error: heavyonlight case not defined for Single Phase")
            ax, half = (1, geo.Length_y / 2.0) if geo.dim == 2 else (2, geo.Length_z / 2.0)
            S = np.where(cc[:, ax] > half, 0.0, 1.0)
        if phases == "Two-phase":
            if S is None:
                S = np.full(n, float(prm.S_o))                                                 # :44-46
            return np.stack([p, T, S])
        return np.stack([p, T])


def source_entries(case, params, geo):
    """Flatten a case into libtpb200 source records (cell, kind, weight, bhp, max_rate, const_rate)."""
    V = geo.cell_volume
    const = bool(getattr(case, "constant_rate", False))
    out = []

    def add(delta, kind, bhp, max_rate):
        for c in np.nonzero(delta)[0]:
            out.append((int(c), kind, float(delta[c] * V), float(bhp), float(max_rate), const))

    if getattr(case, "name", "").startswith("Sources"):
        add(case.deltas_prod, PROD, params.p_prod, -params.prod_rate)     # sourceterms.py:185-186
        add(case.deltas_inj, INJ, params.p_inj, params.inj_rate)          # :159-160
        add(case.deltas_heaters, HEATER, 0.0, 0.0)
    for w in getattr(case, "prod_wells", []) or []:
        add(w["delta"], PROD, w["bhp"], w["max_rate"])
    for w in getattr(case, "inj_wells", []) or []:
        add(w["delta"], INJ, w["bhp"], w["max_rate"])
    for h in getattr(case, "heaters", []) or []:
        add(h["delta"], HEATER, 0.0, 0.0)
    return out


def source_rates(entries, u_at, Kx_at, Ky_at, params, nphase):
    """The volumetric well totals the reference's time loop evaluates after every step (thermalmodel.py:231-270:
    `assemble(delta * rate * dx)` summed over the wells, or over the global deltas of a SourceTerms case).

    entries: records as returned by source_entries() (cell, kind, weight = delta * cell volume, bhp, max_rate,
    const_rate); u_at[f][e], Kx_at[e], Ky_at[e]: state and horizontal permeabilities at each record's cell.
    The rate laws are wellcase.py:171-266 / sourceterms.py:155-269 (Peaceman index with h = 5, rw = 0.1, Dx = Dy = 5
    hard-wired; the side of the draw-down conditional follows the sign of max_rate; rates at or beyond max_rate are
    capped; oil viscosity for every single-phase well, mixture mobility for two-phase producers, water for two-phase
    injectors; twophase.py:388-408, singlephase.py:151-162).  Returns {"inj", "prod"[, "oil", "water"]} with None for
    an absent well type; heaters carry no rate."""
    out = {"inj": None, "prod": None}
    if nphase == 2:
        out.update(oil=None, water=None)
    ent = [(k, e) for k, e in enumerate(entries) if e[1] in (PROD, INJ)]
    if not ent:
        return out
    idx = np.array([k for k, _ in ent])
    kind = np.array([e[1] for _, e in ent])
    w = np.array([e[2] for _, e in ent], dtype=float)
    bhp = np.array([e[3] for _, e in ent], dtype=float)
    qmax = np.array([e[4] for _, e in ent], dtype=float)
    const = np.array([bool(e[5]) for _, e in ent])
    u_at = np.asarray(u_at, dtype=float)
    p, T = u_at[0][idx], u_at[1][idx]
    imu_o = 1.0 / params.oil_mu(T)
    if nphase == 2:
        S = u_at[2][idx]
        imu_w = 1.0 / params.water_mu(T)
        mu_mix = 1.0 / (S * imu_o + (1.0 - S) * imu_w)                   # wellcase.py:212
        mu = np.where(kind == PROD, mu_mix, 1.0 / imu_w)                 # injectors: phase='water' (twophase.py:401)
    else:
        mu = 1.0 / imu_o
    Kx, Ky = np.asarray(Kx_at, dtype=float)[idx], np.asarray(Ky_at, dtype=float)[idx]
    h, rw, Dx, Dy = 5.0, 0.1, 5.0, 5.0                                   # wellcase.py:182-189
    with np.errstate(divide="ignore", invalid="ignore"):
        ro = 0.28 * ((Ky / Kx) ** 0.5 * Dx ** 2 + (Kx / Ky) ** 0.5 * Dy ** 2) ** 0.5 / ((Ky / Kx) ** 0.25 + (Kx / Ky) ** 0.25)
        factor = 2.0 * np.pi * h * (Kx * Ky) ** 0.5 / np.log(ro / rw) / mu
    d = bhp - p
    dd = np.where(qmax < 0.0, np.where(d >= 0.0, 0.0, d), np.where(d <= 0.0, 0.0, d))   # :193-196
    rate = factor * dd
    rate = np.where(np.abs(rate) - np.abs(qmax) >= 0.0, qmax, rate)     # :198
    rate = np.where(const, qmax, rate)                                   # flow_rate_constant (:200-201, 236-266)
    for name, k in (("inj", INJ), ("prod", PROD)):
        m = kind == k
        if m.any():
            out[name] = float(np.sum(w[m] * rate[m]))
    if nphase == 2 and (kind == PROD).any():
        m = kind == PROD
        out["water"] = float(np.sum(w[m] * ((1.0 - S[m]) * imu_w[m] * mu[m] * rate[m])))   # :233
        out["oil"] = float(np.sum(w[m] * (S[m] * imu_o[m] * mu[m] * rate[m])))             # :234
    return out

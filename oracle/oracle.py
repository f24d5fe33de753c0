"""ctypes/numpy front end of the CPU oracle (oracle/pyro_oracle.c).  TEST INFRASTRUCTURE ONLY:
importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs -- never from pyro2_b200/.

Arrays cross this interface in the reference's own layout ``[i, j, n]`` (x slowest, variable
fastest -- pyro/mesh/patch.py:450-452) and are transposed to the oracle's SoA planes here.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

BC_CODES = {"outflow": 0, "neumann": 0, "reflect-even": 1, "reflect-odd": 2, "dirichlet": 2,
            "periodic": 3, "hse": 4, "ambient": 4, "ramp": 4, "none": 4}   # 4: leave that side alone (user BCs are filled separately)


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("pyro_oracle.c", "incomp_oracle.c")]
    if force or not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(src) for src in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


class CompParams(C.Structure):
    _fields_ = [("gamma", C.c_double), ("z0", C.c_double), ("z1", C.c_double),
                ("delta", C.c_double), ("cvisc", C.c_double), ("limiter", C.c_int),
                ("use_flattening", C.c_int), ("no_avisc_xhi", C.c_int), ("no_avisc_yhi", C.c_int),
                ("grav", C.c_double), ("src_bc", C.c_int * 16),
                ("riemann", C.c_int), ("xl_solid", C.c_int), ("yl_solid", C.c_int),
                ("heat_rate", C.c_double), ("heat_profile", C.c_void_p),
                ("do_sponge", C.c_int), ("sponge_rho_begin", C.c_double), ("sponge_rho_full", C.c_double),
                ("sponge_timescale", C.c_double), ("geom", C.c_void_p)]


class Geom(C.Structure):
    _fields_ = [("xmin", C.c_double), ("ymin", C.c_double)] + \
               [(n, C.c_void_p) for n in ("Ly", "Ax", "Ay", "V", "dlogAx", "dlogAy", "x2d")]


class LmParams(C.Structure):
    _fields_ = [("n", C.c_int), ("ng", C.c_int), ("xmin", C.c_double), ("xmax", C.c_double), ("ymin", C.c_double),
                ("ymax", C.c_double), ("grav", C.c_double), ("gamma", C.c_double), ("limiter", C.c_int),
                ("proj_type", C.c_int), ("bc_dens", C.c_int * 4), ("bc_xvel", C.c_int * 4), ("bc_yvel", C.c_int * 4),
                ("bc_phi", C.c_int * 4)]


_STAGE_NAMES = ["q", "xi", "ldx", "ldy", "Uxl_hat", "Uxr_hat", "Uyl_hat", "Uyr_hat", "Fx_t", "Fy_t",
                "Uxl", "Uxr", "Uyl", "Uyr", "Fx", "Fy"]


class CompStages(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _STAGE_NAMES]


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp = C.c_void_p
        for name in ("orc_fill_ghost_f64", "orc_fill_ghost_i64"):
            f = getattr(L, name)
            f.argtypes = [dp] + [C.c_int] * 7 + [dp] * 4 + [C.c_double] * 2
            f.restype = None
        L.orc_cfl_dt.argtypes = [dp, C.c_int, C.c_int, C.c_int] + [C.c_double] * 4
        L.orc_cfl_dt.restype = C.c_double
        L.orc_compressible_step.argtypes = [dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                            C.c_double, C.POINTER(CompParams), C.POINTER(CompStages)]
        L.orc_compressible_step.restype = C.c_int
        L.orc_fill_hse.argtypes = [dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int]
        L.orc_fill_hse.restype = None
        L.orc_mg_create.argtypes = [C.c_int, C.POINTER(C.c_int * 4)] + [C.c_double] * 6 + [C.c_int] * 2
        L.orc_mg_create.restype = C.c_void_p
        L.orc_mg_destroy.argtypes = [C.c_void_p]
        L.orc_mg_nlevels.argtypes = [C.c_void_p]
        L.orc_mg_nlevels.restype = C.c_int
        L.orc_mg_plane.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_mg_plane.restype = C.POINTER(C.c_double)
        L.orc_mg_set_bc_values.argtypes = [C.c_void_p, C.c_int, dp]
        for name in ("orc_mg_residual", "orc_mg_restrict", "orc_mg_prolong_correct", "orc_mg_vcycle"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_int]
            getattr(L, name).restype = None
        L.orc_mg_smooth.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_mg_smooth.restype = None
        L.orc_mg_set_coeffs.argtypes = [C.c_void_p, dp, C.POINTER(C.c_int * 4)]
        L.orc_mg_set_coeffs.restype = None
        L.orc_mg_coef_plane.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_mg_coef_plane.restype = C.POINTER(C.c_double)
        L.orc_incomp_mac_vels.argtypes = [dp] * 4 + [C.c_int] * 3 + [C.c_double] * 3 + [C.c_int, dp, dp]
        L.orc_incomp_mac_vels.restype = None
        L.orc_incomp_states.argtypes = [dp] * 4 + [C.c_int] * 3 + [C.c_double] * 3 + [C.c_int] + [dp] * 6
        L.orc_incomp_states.restype = None
        L.orc_burgers_evolve.argtypes = [dp, dp] + [C.c_int] * 3 + [C.c_double] * 3 + [C.c_int]
        L.orc_burgers_evolve.restype = None
        L.orc_advection_evolve.argtypes = [dp] + [C.c_int] * 3 + [C.c_double] * 5 + [C.c_int]
        L.orc_advection_evolve.restype = None
        L.orc_diffusion_evolve.argtypes = [dp, C.c_int] + [C.c_double] * 6 + [dp]
        L.orc_diffusion_evolve.restype = C.c_int
        L.orc_lm_evolve.argtypes = [dp, dp, C.POINTER(LmParams), C.c_double, dp]
        L.orc_lm_evolve.restype = None
        L.orc_lm_initial_projection.argtypes = [dp, dp, C.POINTER(LmParams)]
        L.orc_lm_initial_projection.restype = C.c_int
        L.orc_lm_timestep.argtypes = [dp, dp, C.POINTER(LmParams), C.c_double]
        L.orc_lm_timestep.restype = C.c_double
        L.orc_incomp_evolve.argtypes = [dp, C.c_int, C.c_int] + [C.c_double] * 5 + [C.c_int, C.c_int, dp, dp, dp, dp]
        L.orc_incomp_evolve.restype = None
        L.orc_norm.argtypes = [dp, C.c_int, C.c_double, C.c_double]
        L.orc_norm.restype = C.c_double
        L.orc_mg_solve.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, dp, dp]
        L.orc_mg_solve.restype = C.c_int
        _LIB = L
    return _LIB


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _bc4(bc):
    """bc: 4 names (xlb, xrb, ylb, yrb) or an object with those attributes"""
    if hasattr(bc, "xlb"):
        bc = (bc.xlb, bc.xrb, bc.ylb, bc.yrb)
    return [BC_CODES[b] for b in bc]


def fill_ghost(a, ng, bc, values=(None, None, None, None), dx=1.0, dy=1.0):
    """in-place ghost fill of one 2-d plane ``a[qx, qy]`` (float64 or int64);
    pyro/mesh/array_indexer.py:150-274"""
    assert a.ndim == 2 and a.flags.c_contiguous
    qx, qy = a.shape
    codes = _bc4(bc)
    vals = [None if v is None else np.ascontiguousarray(v, dtype=a.dtype) for v in values]
    if a.dtype == np.float64:
        f = lib().orc_fill_ghost_f64
    elif a.dtype == np.int64:
        f = lib().orc_fill_ghost_i64
    else:
        raise TypeError(a.dtype)
    f(_ptr(a), qx - 2 * ng, qy - 2 * ng, ng, *codes, *[_ptr(v) for v in vals], dx, dy)
    return a


def to_planes(U_ijn):
    return np.ascontiguousarray(np.moveaxis(np.asarray(U_ijn, dtype=np.float64), 2, 0))


def from_planes(P):
    return np.ascontiguousarray(np.moveaxis(P, 0, 2))


def cfl_dt(U_ijn, ng, dx, dy, gamma, cfl):
    P = to_planes(U_ijn)
    _, qx, qy = P.shape
    return lib().orc_cfl_dt(_ptr(P), qx - 2 * ng, qy - 2 * ng, ng, dx, dy, gamma, cfl)


def comp_params(gamma=1.4, z0=0.75, z1=0.85, delta=0.33, cvisc=0.1, limiter=2, use_flattening=1,
                no_avisc_xhi=1, no_avisc_yhi=1, grav=0.0, src_bcs=None, riemann="HLLC", xl_solid=0, yl_solid=0,
                heat_rate=0.0, heat_profile=None, sponge=None, geom=None):
    """geom: a spherical_geometry() dict for SphericalPolar grids (CGF only).  src_bcs: BC names (xlb, xrb, ylb, yrb) of the four source arrays in variable order dens, ener, xmom,
    ymom (only needed with sources); "hse" and "ambient" copy like outflow for them.  heat_profile: (qx, qy)
    array P, S_ener = dens * heat_rate * P; sponge: (rho_begin, rho_full, timescale) or None"""
    codes = (C.c_int * 16)()
    if src_bcs is not None:
        flat = [BC_CODES["outflow" if b in ("hse", "ambient", "ramp") else b] for bc in src_bcs for b in bc]
        codes = (C.c_int * 16)(*flat)
    hp = None if heat_profile is None else np.ascontiguousarray(heat_profile, dtype=np.float64)
    sp = sponge or (0.0, 0.0, 1.0)
    prm = CompParams(gamma, z0, z1, delta, cvisc, limiter, use_flattening, no_avisc_xhi, no_avisc_yhi, grav, codes,
                     {"HLLC": 0, "CGF": 1, "HLLC_lm": 2}[riemann], xl_solid, yl_solid, heat_rate,
                     None if hp is None else hp.ctypes.data, int(sponge is not None), sp[0], sp[1], sp[2], None)
    prm._keepalive = hp
    if geom is not None:
        arrs = {k: np.ascontiguousarray(geom[k], dtype=np.float64) for k in ("Ly", "Ax", "Ay", "V", "dlogAx", "dlogAy", "x2d")}
        gs = Geom(geom["xmin"], geom["ymin"], *[arrs[k].ctypes.data for k in ("Ly", "Ax", "Ay", "V", "dlogAx", "dlogAy", "x2d")])
        prm._geom = (gs, arrs)
        prm.geom = C.addressof(gs)
    return prm


def spherical_geometry(nx, ny, ng, xmin, xmax, ymin, ymax):
    """the arrays of a SphericalPolar grid (pyro/mesh/patch.py:242-312; x = r, y = theta), with the reference's numpy
    expressions: side lengths Lx = dx, Ly = r dtheta, face areas Ax, Ay (on the low faces), cell volumes V and the
    logarithmic area derivatives that appear as geometric sources in the characteristic tracing"""
    dx, dy = (xmax - xmin) / nx, (ymax - ymin) / ny
    xl = (np.arange(nx + 2 * ng) - ng) * dx + xmin
    xr = (np.arange(nx + 2 * ng) + 1.0 - ng) * dx + xmin
    x = 0.5 * (xl + xr)
    yl = (np.arange(ny + 2 * ng) - ng) * dy + ymin
    yr = (np.arange(ny + 2 * ng) + 1.0 - ng) * dy + ymin
    y = 0.5 * (yl + yr)
    x2d, y2d = np.meshgrid(x, y, indexing="ij")
    xl2d, yl2d = np.meshgrid(xl, yl, indexing="ij")
    xr2d, yr2d = np.meshgrid(xr, yr, indexing="ij")
    return {"xmin": xmin, "ymin": ymin, "dx": dx, "dy": dy, "x2d": x2d, "y2d": y2d,
            "Lx": np.full(x2d.shape, dx), "Ly": x2d * dy,
            "Ax": np.abs(-2.0 * np.pi * xl2d ** 2 * (np.cos(yr2d) - np.cos(yl2d))),
            "Ay": np.abs(np.pi * np.sin(yl2d) * (xr2d ** 2 - xl2d ** 2)),
            "dlogAx": 2.0 / x2d, "dlogAy": 1.0 / (np.tan(y2d) * x2d),
            "V": np.abs(-2.0 * np.pi / 3.0 * (np.cos(yr2d) - np.cos(yl2d)) * (xr2d - xl2d) *
                        (xr2d ** 2 + xl2d ** 2 + xr2d * xl2d))}


def cfl_dt_spherical(U_ijn, gamma, cfl, geom):
    """method_compute_timestep with the cell side lengths of a curvilinear grid (compressible/simulation.py:267-288):
    cfl * min(Lx / (|u| + cs), Ly / (|v| + cs)) over the whole array"""
    P = to_planes(U_ijn)
    rho = P[0]
    u, v = P[2] / rho, P[3] / rho
    e = (P[1] - 0.5 * rho * (u * u + v * v)) / rho
    cs = np.sqrt(gamma * (rho * e * (gamma - 1.0)) / rho)
    return cfl * float(min((geom["Lx"] / (np.abs(u) + cs)).min(), (geom["Ly"] / (np.abs(v) + cs)).min()))


def fill_hse(P, ng, dy, grav, gamma, var, side):
    """the "hse" boundary of compressible/BC.py for plane `var` (0 dens, 1 ener, 2 xmom, 3 ymom) of the SoA
    state P[n, i, j] on side "ylb" / "yrb", in place"""
    assert P.flags.c_contiguous and P.dtype == np.float64 and P.shape[0] == 4
    lib().orc_fill_hse(_ptr(P), P.shape[1] - 2 * ng, P.shape[2] - 2 * ng, ng, dy, grav, gamma, var,
                       {"ylb": 0, "yrb": 1}[side])


def ramp_inflow(var, gamma, post=True):
    """conserved post- / pre-shock value of plane `var` (0 dens, 1 ener, 2 xmom, 3 ymom) for the double Mach
    reflection boundaries (compressible/BC.py:259-296: inflow_post_bc / inflow_pre_bc, constants in the source)"""
    r, u, v, p = (8.0, 7.1447096, -4.125, 116.5) if post else (1.4, 0.0, 0.0, 1.0)
    return [r, p / (gamma - 1.0) + 0.5 * r * (u * u + v * v), r * u, r * v][var]


def fill_ramp(a, var, side, ng, x, y, dx, dy, t, gamma):
    """the "ramp" boundary of compressible/BC.py:183-256 for one ghost-padded plane a[qx, qy] of variable `var`
    on side "xlb" / "ylb" / "yrb" at time t, in place.  Plain numpy loops over the ghost cells, in the reference's
    order: the upper boundary is a 2 x 2 supersampled average of the post- and pre-shock states about the
    position of the Mach-10 shock, summed as ((q1 + q2) + q3) + q4."""
    import math
    qx, qy = a.shape
    nx, ny = qx - 2 * ng, qy - 2 * ng
    post, pre = ramp_inflow(var, gamma, True), ramp_inflow(var, gamma, False)
    if side == "xlb":                        # post-shock inflow
        a[:ng, :] = post
    elif side == "ylb":                      # inflow left of the ramp's foot at x = 1/6, reflecting wall right of it
        left = x < 1.0 / 6.0
        right = ~left
        for jj in range(ng):
            j = ng - 1 - jj
            a[left, j] = post
            a[right, j] = (-1.0 if var == 3 else 1.0) * a[right, ng + jj]
    elif side == "yrb":                      # the shock's intersection with each ghost row moves with time
        for j in range(ng + ny, qy):
            fronts = [1.0 / 6.0 + (y[j] + s * 0.5 * dy * math.sqrt(3)) / math.tan(math.pi / 3.0)
                      + (10.0 / math.sin(math.pi / 3.0)) * t for s in (-1.0, 1.0)]
            for i in range(qx):
                acc = 0.0
                for sf in fronts:
                    for cx in (x[i] - 0.5 * dx * math.sqrt(3), x[i] + 0.5 * dx * math.sqrt(3)):
                        acc = acc + 0.25 * (post if cx < sf else pre)
                a[i, j] = acc
    else:
        raise ValueError(side)


def compressible_step(U_ijn, ng, dx, dy, dt, params=None, stages=False, planes=False):
    """one evolve() on a ghost-filled state; returns the new state (and the per-stage arrays).
    ``planes=True``: U is already SoA ``[n, i, j]`` and is updated in place (no transposes)."""
    params = params or comp_params()
    P = U_ijn if planes else to_planes(U_ijn)
    assert P.flags.c_contiguous and P.dtype == np.float64
    _, qx, qy = P.shape
    st = CompStages()
    out = {}
    if stages:
        for name in _STAGE_NAMES:
            nvar = 1 if name == "xi" else 4
            out[name] = np.zeros((nvar, qx, qy))
            setattr(st, name, out[name].ctypes.data)
    rc = lib().orc_compressible_step(_ptr(P), qx - 2 * ng, qy - 2 * ng, ng, dx, dy, dt,
                                     C.byref(params), C.byref(st) if stages else None)
    if rc:
        raise AssertionError("invalid state (rho <= 0 or e <= 0)")
    res = P if planes else from_planes(P)
    if stages:
        return res, {k: (v[0] if k == "xi" else from_planes(v)) for k, v in out.items()}
    return res


class MG:
    """oracle multigrid hierarchy; mirrors the call surface the tests need from
    pyro/multigrid/MG.py (init_zeros/init_solution/init_RHS/solve/v_cycle/get_solution)"""

    def __init__(self, nx, bc=("dirichlet",) * 4, alpha=0.0, beta=-1.0, xmin=0.0, xmax=1.0,
                 ymin=0.0, ymax=1.0, nsmooth=10, nsmooth_bottom=50):
        self.nx = nx
        codes = (C.c_int * 4)(*_bc4(bc))
        self._h = lib().orc_mg_create(nx, C.byref(codes), alpha, beta, xmin, xmax, ymin, ymax,
                                      nsmooth, nsmooth_bottom)
        self.nlevels = lib().orc_mg_nlevels(self._h)
        assert 2 ** self.nlevels == nx, "nx must be a power of two"
        self.dx = (xmax - xmin) / nx
        self.dy = (ymax - ymin) / nx
        self.source_norm = 0.0
        self.max_cycles = 100
        self.num_cycles = 0
        self.residuals = []
        self.relative_errors = []

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_mg_destroy(self._h)
            self._h = None

    def plane(self, level, which):
        n = 2 ** (level + 1) + 2
        idx = {"v": 0, "f": 1, "r": 2}[which]
        return np.ctypeslib.as_array(lib().orc_mg_plane(self._h, level, idx), shape=(n, n))

    def set_bc_values(self, side, vals):
        v = None if vals is None else np.ascontiguousarray(vals, dtype=np.float64)
        lib().orc_mg_set_bc_values(self._h, {"xl": 0, "xr": 1, "yl": 2, "yr": 3}[side], _ptr(v))

    def set_coeffs(self, coeffs, coeffs_bc):
        """variable-coefficient mode (pyro/multigrid/variable_coeff_MG.py): div(eta grad phi) = f with
        eta given at the finest level's cell centres, (n+2)^2, and its boundary types"""
        c = np.ascontiguousarray(coeffs, dtype=np.float64)
        codes = (C.c_int * 4)(*_bc4(coeffs_bc))
        lib().orc_mg_set_coeffs(self._h, _ptr(c), C.byref(codes))

    def coef_plane(self, level, which):
        n = 2 ** (level + 1) + 2
        idx = {"c": 0, "ex": 1, "ey": 2}[which]
        return np.ctypeslib.as_array(lib().orc_mg_coef_plane(self._h, level, idx), shape=(n, n))

    def init_zeros(self):
        self.plane(self.nlevels - 1, "v")[:] = 0.0

    def init_solution(self, data):
        self.plane(self.nlevels - 1, "v")[:] = data

    def init_RHS(self, data):
        f = self.plane(self.nlevels - 1, "f")
        f[:] = data
        self.source_norm = self.norm(f)

    def norm(self, a):
        a = np.ascontiguousarray(a)
        return lib().orc_norm(_ptr(a), a.shape[0] - 2, self.dx, self.dy)

    def smooth(self, level, n):
        lib().orc_mg_smooth(self._h, level, n)

    def residual(self, level):
        lib().orc_mg_residual(self._h, level)

    def restrict(self, level):
        lib().orc_mg_restrict(self._h, level)

    def prolong_correct(self, level):
        lib().orc_mg_prolong_correct(self._h, level)

    def v_cycle(self, level=None):
        lib().orc_mg_vcycle(self._h, self.nlevels - 1 if level is None else level)

    def solve(self, rtol=1.e-11):
        res = np.zeros(self.max_cycles)
        rel = np.zeros(self.max_cycles)
        self.num_cycles = lib().orc_mg_solve(self._h, rtol, self.source_norm, self.max_cycles,
                                             _ptr(res), _ptr(rel))
        self.residuals = list(res[:self.num_cycles])
        self.relative_errors = list(rel[:self.num_cycles])
        self.residual_error = self.residuals[-1] if self.residuals else 1.e33
        self.relative_error = self.relative_errors[-1] if self.relative_errors else 1.e33

    def get_solution(self):
        return self.plane(self.nlevels - 1, "v").copy()


# ---- Burgers / incompressible (oracle/incomp_oracle.c) ------------------------------------------------
def _c(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    assert a.ndim == 2
    return a


def incomp_mac_vels(u, v, gradp_x, gradp_y, ng, dx, dy, dt, limiter):
    """incomp_interface.mac_vels on ghost-filled (qx, qy) arrays -> (u_MAC, v_MAC)"""
    u, v, gx, gy = _c(u), _c(v), _c(gradp_x), _c(gradp_y)
    um, vm = np.zeros_like(u), np.zeros_like(u)
    lib().orc_incomp_mac_vels(_ptr(u), _ptr(v), _ptr(gx), _ptr(gy), u.shape[0] - 2 * ng, u.shape[1] - 2 * ng, ng,
                              dx, dy, dt, limiter, _ptr(um), _ptr(vm))
    return um, vm


def incomp_states(u, v, gradp_x, gradp_y, ng, dx, dy, dt, limiter, u_mac, v_mac):
    """incomp_interface.states -> (u_xint, v_xint, u_yint, v_yint)"""
    u, v, gx, gy, um, vm = _c(u), _c(v), _c(gradp_x), _c(gradp_y), _c(u_mac), _c(v_mac)
    out = [np.zeros_like(u) for _ in range(4)]
    lib().orc_incomp_states(_ptr(u), _ptr(v), _ptr(gx), _ptr(gy), u.shape[0] - 2 * ng, u.shape[1] - 2 * ng, ng,
                            dx, dy, dt, limiter, _ptr(um), _ptr(vm), *[_ptr(o) for o in out])
    return out


def burgers_evolve(u, v, ng, dx, dy, dt, limiter):
    """burgers Simulation.evolve: returns the updated (u, v) (valid cells updated, ghosts untouched)"""
    u, v = _c(u).copy(), _c(v).copy()
    lib().orc_burgers_evolve(_ptr(u), _ptr(v), u.shape[0] - 2 * ng, u.shape[1] - 2 * ng, ng, dx, dy, dt, limiter)
    return u, v


def incomp_evolve(planes, ng, dt, limiter=2, proj_type=2, vel_bc=(("periodic",) * 4, ("periodic",) * 4),
                  phi_bc=("periodic",) * 4, xmin=0.0, xmax=1.0, ymin=0.0, ymax=1.0, dump=False):
    """incompressible Simulation.evolve on SoA planes [x-velocity, y-velocity, phi-MAC, phi, gradp_x, gradp_y]
    of shape (6, n + 2 ng, n + 2 ng), updated in place; returns (cycles_MAC, cycles_final[, stage arrays])"""
    assert planes.flags.c_contiguous and planes.dtype == np.float64 and planes.shape[0] == 6
    n = planes.shape[1] - 2 * ng
    vb = np.array(_bc4(vel_bc[0]) + _bc4(vel_bc[1]), dtype=np.int32)
    pb = np.array(_bc4(phi_bc), dtype=np.int32)
    cyc = np.zeros(2, dtype=np.int32)
    d = np.zeros((6,) + planes.shape[1:]) if dump else None
    lib().orc_incomp_evolve(_ptr(planes), n, ng, xmin, xmax, ymin, ymax, dt, limiter, proj_type, _ptr(vb), _ptr(pb),
                            _ptr(cyc), _ptr(d))
    return (int(cyc[0]), int(cyc[1]), d) if dump else (int(cyc[0]), int(cyc[1]))


def advection_evolve(a, ng, dx, dy, dt, u, v, limiter):
    """advection Simulation.evolve for one ghost-filled scalar plane; returns the updated plane"""
    a = _c(a).copy()
    lib().orc_advection_evolve(_ptr(a), a.shape[0] - 2 * ng, a.shape[1] - 2 * ng, ng, dx, dy, dt, u, v, limiter)
    return a


def diffusion_evolve(phi, dt, k, bc, xmin=0.0, xmax=1.0, ymin=0.0, ymax=1.0):
    """diffusion Simulation.evolve on an (n+2, n+2) plane (ng = 1), in place; returns the V-cycle count"""
    assert phi.flags.c_contiguous and phi.dtype == np.float64 and phi.shape[0] == phi.shape[1]
    codes = np.array(_bc4(bc), dtype=np.int32)
    return lib().orc_diffusion_evolve(_ptr(phi), phi.shape[0] - 2, xmin, xmax, ymin, ymax, dt, k, _ptr(codes))


# ---- low Mach number atmosphere (oracle/lm_oracle.c) ----------------------------------------------------
LM_VARS = ["density", "x-velocity", "y-velocity", "eint", "phi-MAC", "phi", "gradp_x", "gradp_y"]


def lm_params(n, ng=4, grav=-2.0, gamma=1.4, limiter=2, proj_type=2, bc_dens=("periodic", "periodic", "reflect-even", "outflow"),
              bc_xvel=None, bc_yvel=None, bc_phi=("periodic", "periodic", "neumann", "dirichlet"),
              xmin=0.0, xmax=1.0, ymin=0.0, ymax=1.0):
    """BC names per variable class as CellCenterData2d resolves them (reflect -> reflect-even / reflect-odd)"""
    bc_xvel = bc_xvel or bc_dens
    bc_yvel = bc_yvel or tuple("reflect-odd" if b == "reflect-even" and k >= 2 else b for k, b in enumerate(bc_dens))
    mk = lambda bc: (C.c_int * 4)(*_bc4(bc))
    return LmParams(n, ng, xmin, xmax, ymin, ymax, grav, gamma, limiter, proj_type, mk(bc_dens), mk(bc_xvel), mk(bc_yvel),
                    mk(bc_phi))


def lm_evolve(S, base, prm, dt):
    """lm_atm Simulation.evolve on the 8 SoA state planes S[n, i, j] (ghost cells filled) and the base-state
    arrays base[4, qy] = rho0, p0, beta0, beta0-edges; in place; returns the two V-cycle counts"""
    assert S.flags.c_contiguous and S.shape[0] == 8 and base.flags.c_contiguous and base.shape[0] == 4
    cyc = np.zeros(2, dtype=np.int32)
    lib().orc_lm_evolve(_ptr(S), _ptr(base), C.byref(prm), dt, _ptr(cyc))
    return int(cyc[0]), int(cyc[1])


def lm_initial_projection(S, base, prm):
    return lib().orc_lm_initial_projection(_ptr(S), _ptr(base), C.byref(prm))


def lm_timestep(S, base, prm, cfl):
    return lib().orc_lm_timestep(_ptr(S), _ptr(base), C.byref(prm), cfl)

"""Shared helpers of the emulator-based CPU tests: build / load the host-compiled kernel libraries
(tests/emu/) and drive them through the product's own ctypes signatures.  TEST INFRASTRUCTURE ONLY."""
import contextlib
import ctypes as C
import fcntl
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "emu")
CSRC = os.path.join(os.path.dirname(HERE), "pyro2_b200", "csrc")
BC = {"outflow": 0, "neumann": 0, "reflect-even": 1, "reflect-odd": 2, "dirichlet": 2, "periodic": 3}


_LIBS = {}


@contextlib.contextmanager
def _build_lock():
    """several test processes (multi-rank gloo tests, xdist workers) may find the same library stale at the same
    time: one builds, the others wait and then find it fresh"""
    with open(os.path.join(EMU_DIR, ".build.lock"), "w") as fh:
        fcntl.flock(fh, fcntl.LOCK_EX)
        try:
            yield
        finally:
            fcntl.flock(fh, fcntl.LOCK_UN)


def _build(so, deps, cmd, cwd=None):
    with _build_lock():
        if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            tmp = so + f".{os.getpid()}.tmp"
            subprocess.check_call([c if c != so else tmp for c in cmd], cwd=cwd)
            os.replace(tmp, so)


def _load(kind, sources, prefix, extra=()):
    if kind in _LIBS:
        return _LIBS[kind]
    so = os.path.join(EMU_DIR, f"lib{kind}_emu.so")
    deps = [os.path.join(EMU_DIR, f) for f in (f"{kind}_emu.cpp", "cuda_emu.h", "cuda_emu_runtime.inc")] + \
           [os.path.join(CSRC, f) for f in sources + ["hydro_core.cuh", "common.cuh"]]
    _build(so, deps, ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
                      "-x", "c++", "-DP2B_EMU_HEADER=\"../../tests/emu/cuda_emu.h\"", *extra,
                      "-o", so, os.path.join(EMU_DIR, f"{kind}_emu.cpp")], cwd=EMU_DIR)
    lib = C.CDLL(so)
    from pyro2_b200 import _lib
    for name, (res, args) in _lib.SIGNATURES.items():
        if name.startswith(prefix) or name == "p2b_last_error":
            f = getattr(lib, name)
            f.restype, f.argtypes = res, args
    _LIBS[kind] = lib
    return lib


def load_mg_emu():
    """pyro2_b200/csrc/mg.cu compiled for the host: the p2b_mg_* ABI over numpy memory"""
    return _load("mg", ["mg.cu", "mg_kernels.cuh"], ("p2b_mg_", "p2b_shared_"), ["-DMG_COARSE_THREADS=128"])


def load_bc_emu():
    """pyro2_b200/csrc/bc_user.cu compiled for the host: p2b_fill_hse_f64 over numpy memory"""
    return _load("bc", ["bc_user.cu", "bc_user_kernels.cuh"], ("p2b_fill_hse", "p2b_fill_ambient"))


def load_ghost_emu():
    """pyro2_b200/csrc/ghost_cfl.cu compiled for the host: ghost fill and CFL wave speeds over numpy memory"""
    return _load("ghost", ["ghost_cfl.cu", "slab_comm.cu", "peer_comm.cuh"], ("p2b_fill_ghost", "p2b_cfl_wavemax", "p2b_device_sms", "p2b_slab_"))


def load_sweep_emu():
    """the fused sweep's task source (sweep_task.cuh) under its own warp emulator (tests/emu/sweep_emu.cpp: 32 host
    threads in lock step stand in for the lanes, memcpy for the TMA bulk copies)"""
    if "sweep" in _LIBS:
        return _LIBS["sweep"]
    so = os.path.join(EMU_DIR, "libsweep_emu.so")
    src = os.path.join(EMU_DIR, "sweep_emu.cpp")
    hdrs = [os.path.join(CSRC, h) for h in ("sweep_task.cuh", "sweep_args.cuh", "hydro_core.cuh")] + \
           [os.path.join(os.path.dirname(CSRC), "..", "include", "pyro2b200.h")]
    _build(so, [src] + hdrs, ["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                              "-Wno-unknown-pragmas", "-pthread", src, "-o", so])
    lib = C.CDLL(so)
    lib.emu_compressible_sweep.argtypes = ([C.c_void_p, C.c_void_p] + [C.c_int] * 4 + [C.c_longlong] +
                                           [C.c_double] * 8 + [C.c_int] * 5 + [C.c_void_p] * 2 +
                                           [C.c_double, C.c_int, C.c_int] + [C.c_int] * 3 +
                                           [C.c_void_p, C.c_double, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int] +
                                           [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int])
    from pyro2_b200 import _lib
    lib.emu_compressible_sweep_abi.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(_lib.Grid), C.POINTER(_lib.CompParams),
                                               C.c_double, C.c_void_p, C.c_int, C.POINTER(C.c_char_p)]
    lib.emu_sweep_decomposition.argtypes = [C.POINTER(_lib.Grid), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.emu_sweep_decomposition.restype = None
    lib.emu_test_fastmath.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    _LIBS["sweep"] = lib
    return lib


def load_lm_emu():
    """pyro2_b200/csrc/lm.cu compiled for the host: the p2b_lm_* ABI over numpy memory"""
    return _load("lm", ["lm.cu", "lm_kernels.cuh", "flow_kernels.cuh"], "p2b_lm_", ["-Wno-unused-function"])


def load_flow_emu():
    """pyro2_b200/csrc/flow.cu compiled for the host: the p2b_flow_* ABI over numpy memory"""
    return _load("flow", ["flow.cu", "flow_kernels.cuh"], "p2b_flow_")


class EmuMG:
    """solve() of pyro2_b200/multigrid/MG.py re-stated over the emulated library (host memory)"""

    def __init__(self, lib, nx, bc=("dirichlet",) * 4, alpha=0.0, beta=-1.0, blocking=True):
        self.lib, self.nx = lib, nx
        codes = (C.c_int * 4)(*[BC[b] for b in bc])
        self.h = lib.p2b_mg_create(nx, codes, alpha, beta, 0.0, 1.0, 0.0, 1.0, 10, 50)
        assert self.h, lib.p2b_last_error()
        self.nlevels = lib.p2b_mg_nlevels(self.h)
        nbytes = lib.p2b_mg_workspace_bytes(self.h)
        self.ws = np.zeros(nbytes // 8 + 2)
        off = (-self.ws.ctypes.data // 8) % 2          # 16-byte alignment
        self.base = self.ws[off:]
        self.ck(lib.p2b_mg_bind(self.h, self.base.ctypes.data, nbytes))
        if not blocking:
            self.ck(lib.p2b_mg_set_blocking(self.h, 0))
        self.out = np.zeros(2)
        self.keep = []

    def ck(self, rc):
        assert rc == 0, self.lib.p2b_last_error().decode()

    def close(self):
        self.lib.p2b_mg_destroy(self.h)

    def plane(self, level, which):
        n = 2 << level
        pitch = self.lib.p2b_mg_level_pitch(self.h, level)
        ptr = self.lib.p2b_mg_level_ptr(self.h, level, {"v": 0, "f": 1, "r": 2, "w": 3}[which])
        off = (ptr - self.base.ctypes.data) // 8
        return np.lib.stride_tricks.as_strided(self.base[off:], (n + 2, n + 2), (pitch * 8, 8))

    def set_bc_values(self, xl, xr, yl, yr):
        vals = [None if v is None else np.ascontiguousarray(v, dtype=np.float64) for v in (xl, xr, yl, yr)]
        self.keep = vals
        self.ck(self.lib.p2b_mg_set_bc_values(self.h, *[None if v is None else v.ctypes.data for v in vals]))

    def set_coeffs(self, coeffs, coeffs_bc):
        nbytes = self.lib.p2b_mg_coeff_workspace_bytes(self.h)
        self.cws = np.zeros(nbytes // 8 + 2)
        off = (-self.cws.ctypes.data // 8) % 2
        self.cbase = self.cws[off:]
        c = np.ascontiguousarray(coeffs, dtype=np.float64)
        codes = (C.c_int * 4)(*[BC[b] for b in coeffs_bc])
        self.ck(self.lib.p2b_mg_set_coeffs(self.h, self.cbase.ctypes.data, nbytes, c.ctypes.data, c.shape[1], codes, None))

    def coef_plane(self, level, which):
        n = 2 << level
        pitch = self.lib.p2b_mg_level_pitch(self.h, level)
        ptr = self.lib.p2b_mg_coeff_ptr(self.h, level, {"c": 0, "ex": 1, "ey": 2}[which])
        off = (ptr - self.cbase.ctypes.data) // 8
        return np.lib.stride_tricks.as_strided(self.cbase[off:], (n + 2, n + 2), (pitch * 8, 8))

    def sumsq(self, level, which):
        self.ck(self.lib.p2b_mg_norm2(self.h, level, {"v": 0, "f": 1, "r": 2}[which], self.out.ctypes.data, None))
        return float(self.out[0])

    def solve(self, f, rtol=1e-11, max_cycles=100, v0=None):
        fine = self.nlevels - 1
        n = self.nx
        h2 = (1.0 / n) ** 2
        self.plane(fine, "v")[:] = 0.0 if v0 is None else v0
        self.plane(fine, "f")[:] = f
        self.source_norm = math.sqrt(h2 * self.sumsq(fine, "f"))
        pitch = self.lib.p2b_mg_level_pitch(self.h, fine)
        old_phi = np.zeros((n + 2, pitch))
        old_phi[:, :n + 2] = self.plane(fine, "v")
        cycle, resid = 1, 1e33
        while resid > rtol and cycle <= max_cycles:
            self.ck(self.lib.p2b_mg_zero_coarse(self.h, None))
            self.ck(self.lib.p2b_mg_vcycle(self.h, None))
            self.ck(self.lib.p2b_mg_cycle_diagnostics(self.h, old_phi.ctypes.data, self.out.ctypes.data, None))
            rnorm = math.sqrt(h2 * self.out[1])
            resid = rnorm / self.source_norm if self.source_norm != 0.0 else rnorm
            cycle += 1
        self.ck(self.lib.p2b_mg_fill_bc(self.h, fine, None))
        self.num_cycles = cycle - 1
        self.residual_error = resid
        return self.plane(fine, "v").copy()




class EmuFlow:
    """the p2b_flow_* stage calls over numpy planes of shape (qx, qy), plus the orchestration of
    incompressible Simulation.evolve (simulation.py:159-404) with EmuMG doing the two projections --
    the same sequence pyro2_b200/incompressible/simulation.py issues on the device"""

    PLANES = ["u_xl", "u_xr", "u_yl", "u_yr", "v_xl", "v_xr", "v_yl", "v_yr", "uhat", "vhat",
              "u_xint", "v_xint", "u_yint", "v_yint", "u_MAC", "v_MAC"]

    def __init__(self, lib, nx, ny, ng, dx, dy):
        from pyro2_b200 import _lib
        self.lib, self.nx, self.ny, self.ng = lib, nx, ny, ng
        self.qx, self.qy = nx + 2 * ng, ny + 2 * ng
        self.grid = _lib.Grid(nx, ny, ng, self.qy, self.qx * self.qy, dx, dy)
        self.h = lib.p2b_flow_create(C.byref(self.grid))
        assert self.h, lib.p2b_last_error()
        nbytes = lib.p2b_flow_workspace_bytes(self.h)
        self.ws = np.zeros(nbytes // 8)
        self.ck(lib.p2b_flow_bind(self.h, self.ws.ctypes.data, nbytes))

    def ck(self, rc):
        assert rc == 0, self.lib.p2b_last_error().decode()

    def close(self):
        self.lib.p2b_flow_destroy(self.h)

    def plane(self, name):
        ptr = self.lib.p2b_flow_plane(self.h, self.PLANES.index(name))
        off = (ptr - self.ws.ctypes.data) // 8
        return self.ws[off:off + self.qx * self.qy].reshape(self.qx, self.qy)

    def interface_states(self, u, v, gpx, gpy, dt, limiter):
        self.ck(self.lib.p2b_flow_interface_states(self.h, u.ctypes.data, v.ctypes.data,
                                                   None if gpx is None else gpx.ctypes.data,
                                                   None if gpy is None else gpy.ctypes.data, dt, limiter, None))

    def mac_vels(self):
        self.ck(self.lib.p2b_flow_mac_vels(self.h, None))

    def maxabs(self, u, v):
        out = np.zeros(2, dtype=np.uint64)
        self.ck(self.lib.p2b_flow_maxabs(self.h, u.ctypes.data, v.ctypes.data, out.ctypes.data, None))
        return out.view(np.float64)

    def burgers_evolve(self, u, v, dt, limiter):
        self.interface_states(u, v, None, None, dt, limiter)
        self.mac_vels()
        self.ck(self.lib.p2b_flow_burgers_update(self.h, u.ctypes.data, v.ctypes.data, dt, None))

    def incomp_evolve(self, mg_lib, P, dt, limiter, proj_type, bc, fill):
        """P: (6, qx, qy) planes x-velocity, y-velocity, phi-MAC, phi, gradp_x, gradp_y; fill(plane) = ghost fill"""
        L, h, n, ng = self.lib, self.h, self.nx, self.ng
        u, v, phi_mac, phi, gpx, gpy = (P[k] for k in range(6))
        ptr = lambda a: a.ctypes.data
        self.interface_states(u, v, gpx, gpy, dt, limiter)
        self.mac_vels()
        div = np.zeros((n + 2, n + 2))
        self.ck(L.p2b_flow_mac_divergence(h, ptr(div), n + 2, None))
        mg = EmuMG(mg_lib, n, bc, 0.0, -1.0)
        sol = mg.solve(div, rtol=1.e-12)
        mg.close()
        phi_mac[ng - 1:ng + n + 1, ng - 1:ng + n + 1] = sol
        self.ck(L.p2b_flow_mac_project(h, ptr(phi_mac), None))
        self.ck(L.p2b_flow_upwind_states(h, None))
        self.ck(L.p2b_flow_advect_update(h, ptr(u), ptr(v), ptr(gpx), ptr(gpy), dt, proj_type, None))
        fill(u); fill(v)
        div = np.zeros((n + 2, n + 2))
        self.ck(L.p2b_flow_cc_divergence(h, ptr(u), ptr(v), ptr(div), n + 2, dt, 1, None))
        mg = EmuMG(mg_lib, n, bc, 0.0, -1.0)
        sol = mg.solve(div, rtol=1.e-12, v0=phi[ng - 1:ng + n + 1, ng - 1:ng + n + 1].copy())
        mg.close()
        phi[:] = 0.0
        phi[ng - 1:ng + n + 1, ng - 1:ng + n + 1] = sol
        self.ck(L.p2b_flow_project(h, ptr(phi), ptr(u), ptr(v), ptr(gpx), ptr(gpy), dt, proj_type, None))
        fill(u); fill(v)


class EmuLm:
    """the p2b_lm_* stage calls over numpy planes plus the orchestration of lm_atm Simulation.evolve / preevolve /
    method_compute_timestep (lm_atm/simulation.py:138-618) with EmuMG doing the variable-coefficient projections --
    the same sequence pyro2_b200/lm_atm/simulation.py issues on the device"""

    COEFF, SOURCE = 16, 17

    def __init__(self, lib, mg_lib, n, ng, base, fills, phi_bc, grav=-2.0, gamma=1.4, limiter=2, proj_type=2):
        """base: (4, qy) rho0, p0, beta0, beta0-edges; fills: name -> BC names for oracle.fill_ghost"""
        from pyro2_b200 import _lib
        self.lib, self.mg_lib, self.n, self.ng = lib, mg_lib, n, ng
        self.q = n + 2 * ng
        self.base = np.ascontiguousarray(base, dtype=np.float64)
        self.fills, self.phi_bc = fills, phi_bc
        self.grav, self.gamma, self.limiter, self.proj_type = grav, gamma, limiter, proj_type
        self.grid = _lib.Grid(n, n, ng, self.q, self.q * self.q, 1.0 / n, 1.0 / n)
        self.h = lib.p2b_lm_create(C.byref(self.grid), self.base.ctypes.data)
        assert self.h, lib.p2b_last_error()
        nbytes = lib.p2b_lm_workspace_bytes(self.h)
        self.ws = np.zeros(nbytes // 8)
        self.ck(lib.p2b_lm_bind(self.h, self.ws.ctypes.data, nbytes))

    def ck(self, rc):
        assert rc == 0, self.lib.p2b_last_error().decode()

    def close(self):
        self.lib.p2b_lm_destroy(self.h)

    def plane(self, idx):
        ptr = self.lib.p2b_lm_plane(self.h, idx)
        off = (ptr - self.ws.ctypes.data) // 8
        return self.ws[off:off + self.q * self.q].reshape(self.q, self.q)

    def _fill(self, a, name):
        import oracle
        oracle.fill_ghost(a, self.ng, self.fills[name])

    def _solve(self, coeff_plane, div, rtol, v0=None):
        n, ng = self.n, self.ng
        mg = EmuMG(self.mg_lib, n, self.phi_bc, 0.0, 0.0)
        mg.set_coeffs(np.ascontiguousarray(coeff_plane[ng - 1:ng + n + 1, ng - 1:ng + n + 1]), self.fills["density"])
        sol = mg.solve(div, rtol=rtol, v0=v0)
        cycles = mg.num_cycles
        mg.close()
        return sol, cycles

    def timestep(self, S, cfl):
        out = np.zeros(5, dtype=np.uint64)
        self.ck(self.lib.p2b_lm_reduce(self.h, S[0].ctypes.data, S[1].ctypes.data, S[2].ctypes.data, self.grav,
                                       out.ctypes.data, None))
        uall, vall, uval, vval, fb = out.view(np.float64)
        dx = 1.0 / self.n
        xtmp = ytmp = 1.e33
        if not uall == 0:
            xtmp = dx / uval
        if not vall == 0:
            ytmp = dx / vval
        dt = cfl * min(xtmp, ytmp)
        with np.errstate(divide="ignore"):
            dt_buoy = np.sqrt(2.0 * dx / fb)
        return float(min(dt, dt_buoy))

    def initial_projection(self, S):
        L, h, n, ng = self.lib, self.h, self.n, self.ng
        rho, u, v, phi = S[0], S[1], S[2], S[5]
        p = lambda a: a.ctypes.data
        self._fill(rho, "density"); self._fill(u, "x-velocity"); self._fill(v, "y-velocity")
        self.ck(L.p2b_lm_coeff(h, p(rho), None, 1.0, 1, 0, None))
        div = np.zeros((n + 2, n + 2))
        self.ck(L.p2b_lm_cc_divergence(h, p(u), p(v), p(div), n + 2, 1.0, 0, None))
        sol, _ = self._solve(self.plane(self.COEFF), div, 1.e-10)
        phi[:] = 0.0
        phi[ng - 1:ng + n + 1, ng - 1:ng + n + 1] = sol
        self.ck(L.p2b_lm_project(h, p(rho), p(phi), p(u), p(v), None, None, 1.0, 0, None))
        self._fill(u, "x-velocity"); self._fill(v, "y-velocity")

    def evolve(self, S, dt):
        L, h, n, ng = self.lib, self.h, self.n, self.ng
        rho, u, v, eint, phi_mac, phi, gpx, gpy = (S[k] for k in range(8))
        p = lambda a: a.ctypes.data
        coeff, source = self.plane(self.COEFF), self.plane(self.SOURCE)
        # MAC velocities
        self.ck(L.p2b_lm_coeff(h, p(rho), None, 1.0, 0, 0, None)); self._fill(coeff, "density")
        self.ck(L.p2b_lm_source(h, p(rho), None, self.grav, None)); self._fill(source, "y-velocity")
        self.ck(L.p2b_lm_interface_states(h, p(u), p(v), p(gpx), p(gpy), dt, self.limiter, None))
        self.ck(L.p2b_lm_mac_vels(h, None))
        # MAC projection
        self.ck(L.p2b_lm_coeff(h, p(rho), None, 1.0, 1, 1, None))
        div = np.zeros((n + 2, n + 2))
        self.ck(L.p2b_lm_mac_divergence(h, p(div), n + 2, None))
        sol, c0 = self._solve(coeff, div, 1.e-12)
        phi_mac[:] = 0.0
        phi_mac[ng - 1:ng + n + 1, ng - 1:ng + n + 1] = sol
        self.ck(L.p2b_lm_coeff(h, p(rho), None, 1.0, 0, 0, None)); self._fill(coeff, "density")
        self.ck(L.p2b_lm_mac_project(h, p(phi_mac), None))
        # density
        self.ck(L.p2b_lm_density_update(h, p(rho), p(eint), dt, self.limiter, self.gamma, None))
        self._fill(rho, "density")
        # velocities
        rho_old = self.plane(24)
        self.ck(L.p2b_lm_coeff(h, p(rho), p(rho_old), 2.0, 0, 0, None)); self._fill(coeff, "density")
        self.ck(L.p2b_lm_interface_states(h, p(u), p(v), p(gpx), p(gpy), dt, self.limiter, None))
        self.ck(L.p2b_lm_upwind_states(h, None))
        self.ck(L.p2b_lm_advect_update(h, p(u), p(v), p(gpx), p(gpy), dt, self.proj_type, None))
        self.ck(L.p2b_lm_source(h, p(rho), p(rho_old), self.grav, None)); self._fill(source, "y-velocity")
        self.ck(L.p2b_lm_add_source(h, p(v), dt, None))
        self._fill(u, "x-velocity"); self._fill(v, "y-velocity")
        # final projection
        self.ck(L.p2b_lm_coeff(h, p(rho), None, 1.0, 1, 0, None))
        div = np.zeros((n + 2, n + 2))
        self.ck(L.p2b_lm_cc_divergence(h, p(u), p(v), p(div), n + 2, dt, 1, None))
        sol, c1 = self._solve(coeff, div, 1.e-12, v0=phi[ng - 1:ng + n + 1, ng - 1:ng + n + 1].copy())
        phi[:] = 0.0
        phi[ng - 1:ng + n + 1, ng - 1:ng + n + 1] = sol
        self.ck(L.p2b_lm_project(h, p(rho), p(phi), p(u), p(v), p(gpx), p(gpy), dt, self.proj_type, None))
        self._fill(u, "x-velocity"); self._fill(v, "y-velocity")
        self._fill(gpx, "gradp_x"); self._fill(gpy, "gradp_y")
        return c0, c1

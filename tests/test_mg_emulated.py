"""CPU: the multigrid CUDA kernels AND their host orchestration (pyro2_b200/csrc/mg.cu, unchanged)
compiled for the host through tests/emu/cuda_emu.h and driven through the same p2b_mg_* C ABI over
numpy memory, compared bit-for-bit with the oracle.  One fiber per CUDA thread for the kernels that
synchronise (temporally blocked smoother, fused coarse V-cycle, reductions).  The emulator is test
infrastructure: the product only loads the nvcc-built library and has no CPU path."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

import oracle
from golden_util import load_mg, load_mgvc

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "emu")
CSRC = os.path.join(os.path.dirname(HERE), "pyro2_b200", "csrc")
BC = {"outflow": 0, "neumann": 0, "reflect-even": 1, "reflect-odd": 2, "dirichlet": 2, "periodic": 3}


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(EMU_DIR, "libmg_emu.so")
    deps = [os.path.join(EMU_DIR, f) for f in ("mg_emu.cpp", "cuda_emu.h")] + \
           [os.path.join(CSRC, f) for f in ("mg.cu", "mg_kernels.cuh", "hydro_core.cuh", "common.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
                               "-x", "c++", "-DP2B_EMU_HEADER=\"../../tests/emu/cuda_emu.h\"",
                               "-DMG_COARSE_THREADS=128", "-o", so, os.path.join(EMU_DIR, "mg_emu.cpp")],
                              cwd=EMU_DIR)
    lib = C.CDLL(so)
    from pyro2_b200 import _lib
    for name, (res, args) in _lib.SIGNATURES.items():
        if name.startswith("p2b_mg_") or name == "p2b_last_error":
            f = getattr(lib, name)
            f.restype, f.argtypes = res, args
    return lib


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

    def solve(self, f, rtol=1e-11, max_cycles=100):
        fine = self.nlevels - 1
        n = self.nx
        h2 = (1.0 / n) ** 2
        self.plane(fine, "v")[:] = 0.0
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


@pytest.mark.parametrize("name,blocking", [("poisson_dirichlet_64", True), ("poisson_periodic_64", True),
                                           ("helmholtz_neumann_64", True), ("poisson_mixed_128", True),
                                           ("poisson_mixed_128", False)])
def test_emulated_constant_coefficient_solve_matches_reference(emu, name, blocking):
    """fixtures produced by the reference; 128^2 goes through the temporally blocked smoother (EDGE
    path on every CTA) and the fused coarse V-cycle kernel"""
    z = load_mg(name)
    m = EmuMG(emu, int(z["nx"]), tuple(str(b) for b in z["bc"]), float(z["alpha"]), float(z["beta"]), blocking)
    v = m.solve(z["f"], rtol=float(z["rtol"]))
    assert m.num_cycles == int(z["num_cycles"])
    assert np.array_equal(v, z["v"])
    m.close()


def test_emulated_blocked_smoother_interior_path_256(emu):
    """256^2: the first size with CTAs whose whole region is interior (EDGE = false instruction stream)"""
    n = 256
    rng = np.random.default_rng(3)
    m = EmuMG(emu, n, ("dirichlet", "neumann", "periodic", "periodic"), 0.3, 0.7)
    o = oracle.MG(n, bc=("dirichlet", "neumann", "periodic", "periodic"), alpha=0.3, beta=0.7)
    f = rng.standard_normal((n + 2, n + 2))
    v0 = rng.standard_normal((n + 2, n + 2))
    fine = m.nlevels - 1
    m.plane(fine, "v")[:] = v0
    m.plane(fine, "f")[:] = f
    o.plane(fine, "v")[:] = v0
    o.plane(fine, "f")[:] = f
    m.ck(emu.p2b_mg_smooth(m.h, fine, 7, None))      # passes of 5 + 2 iterations, result copied back from w
    o.smooth(fine, 7)
    assert np.array_equal(m.plane(fine, "v"), o.plane(fine, "v"))
    m.close()


@pytest.mark.parametrize("name", ["dirichlet_64", "periodic_64", "constant_32", "dirichlet_128"])
def test_emulated_variable_coefficient_solve_matches_reference(emu, name):
    """VarCoeffCCMG2d fixtures produced by the reference (its mg_test_vc_* setups)"""
    z = load_mgvc(name)
    n = int(z["nx"])
    m = EmuMG(emu, n, tuple(str(b) for b in z["bc"]), 0.0, 0.0)
    m.set_coeffs(z["coeffs"], tuple(str(b) for b in z["coeffs_bc"]))
    assert np.array_equal(m.coef_plane(2, "ex"), z["ex_coarse"])
    assert np.array_equal(m.coef_plane(2, "ey"), z["ey_coarse"])
    v = m.solve(z["f"], rtol=float(z["rtol"]))
    assert m.num_cycles == int(z["num_cycles"])
    assert np.array_equal(v, z["v"])
    assert np.array_equal(m.plane(m.nlevels - 1, "r")[1:n + 1, 1:n + 1], z["r"][1:n + 1, 1:n + 1])
    m.close()


def test_emulated_variable_coefficient_hierarchy_matches_oracle(emu):
    """every level's eta, eta_x, eta_y for random coefficients and mixed coefficient BCs"""
    n = 64
    rng = np.random.default_rng(11)
    coeffs = 0.5 + rng.random((n + 2, n + 2))
    cbc = ("neumann", "reflect-even", "periodic", "periodic")
    m = EmuMG(emu, n, ("dirichlet", "neumann", "periodic", "periodic"), 0.0, 0.0)
    m.set_coeffs(coeffs, cbc)
    o = oracle.MG(n, bc=("dirichlet", "neumann", "periodic", "periodic"), alpha=0.0, beta=0.0)
    o.set_coeffs(coeffs, cbc)
    for lev in range(o.nlevels):
        for which in ("c", "ex", "ey"):
            assert np.array_equal(m.coef_plane(lev, which), o.coef_plane(lev, which)), (lev, which)
    m.close()


@pytest.mark.parametrize("n,bc,cbc", [
    (128, ("periodic",) * 4, ("neumann", "neumann", "reflect-even", "neumann")),   # seam cells: eta(n+1) != eta(1)
    (128, ("dirichlet", "neumann", "periodic", "periodic"), ("neumann", "neumann", "periodic", "periodic")),
    (256, ("dirichlet", "neumann", "neumann", "dirichlet"), ("neumann",) * 4),    # has interior-path CTAs
    (256, ("periodic",) * 4, ("periodic",) * 4),
])
def test_emulated_variable_coefficient_blocked_smoother(emu, n, bc, cbc):
    """the temporally blocked smoother with the coefficient tiles in shared memory vs the oracle's
    plain red-black sweeps, and vs the per-colour kernels of the same library"""
    rng = np.random.default_rng(n)
    coeffs = 0.5 + rng.random((n + 2, n + 2))
    f = rng.standard_normal((n + 2, n + 2))
    v0 = rng.standard_normal((n + 2, n + 2))
    o = oracle.MG(n, bc=bc, alpha=0.0, beta=0.0)
    o.set_coeffs(coeffs, cbc)
    fine = o.nlevels - 1
    o.plane(fine, "v")[:] = v0
    o.plane(fine, "f")[:] = f
    o.smooth(fine, 7)
    for blocking in (True, False):
        m = EmuMG(emu, n, bc, 0.0, 0.0, blocking)
        m.set_coeffs(coeffs, cbc)
        m.plane(fine, "v")[:] = v0
        m.plane(fine, "f")[:] = f
        m.ck(emu.p2b_mg_smooth(m.h, fine, 7, None))
        assert np.array_equal(m.plane(fine, "v"), o.plane(fine, "v")), blocking
        m.close()

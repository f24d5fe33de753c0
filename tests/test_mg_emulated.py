"""CPU: the multigrid CUDA kernels AND their host orchestration (pyro2_b200/csrc/mg.cu, unchanged)
compiled for the host through tests/emu/cuda_emu.h and driven through the same p2b_mg_* C ABI over
numpy memory, compared bit-for-bit with the oracle.  One fiber per CUDA thread for the kernels that
synchronise (temporally blocked smoother, fused coarse V-cycle, reductions).  The emulator is test
infrastructure: the product only loads the nvcc-built library and has no CPU path."""
import numpy as np
import pytest

import oracle
from emu_util import EmuMG, load_mg_emu
from golden_util import load_mg, load_mgvc



@pytest.fixture(scope="module")
def emu():
    return load_mg_emu()


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


@pytest.mark.parametrize("n,bc", [(32, ("neumann",) * 4), (64, ("periodic", "periodic", "dirichlet", "neumann"))])
def test_emulated_diffusion_step_matches_oracle(emu, n, bc):
    """the diffusion solver's step (diffusion/simulation.py:62-104): Crank-Nicolson right-hand side kernel,
    a hierarchy whose beta follows dt (p2b_mg_set_operator), solve to 1e-10"""
    rng = np.random.default_rng(n)
    phi = 1.0 + rng.random((n + 2, n + 2))
    ref = phi.copy()
    m = EmuMG(emu, n, bc, 1.0, 0.123)
    fine = m.nlevels - 1
    for dt in (0.4 / n ** 2, 1.7 / n ** 2):
        k = 1.3
        oracle.fill_ghost(phi, 1, bc)
        m.ck(emu.p2b_mg_set_operator(m.h, 1.0, 0.5 * dt * k))
        m.ck(emu.p2b_mg_cn_rhs(m.h, phi.ctypes.data, n + 2, 0.5 * dt * k, None))
        f = m.plane(fine, "f").copy()
        sol = m.solve(f, rtol=1e-10)
        phi[1:-1, 1:-1] = sol[1:-1, 1:-1]
        cyc = oracle.diffusion_evolve(ref, dt, k, bc)
        assert cyc == m.num_cycles
        assert np.array_equal(phi[1:-1, 1:-1], ref[1:-1, 1:-1])
    m.close()


@pytest.mark.parametrize("bc", [("dirichlet", "neumann", "periodic", "periodic"), ("periodic", "periodic", "dirichlet", "dirichlet")])
def test_emulated_blocked_smoother_inhomogeneous_values_with_periodic_other_direction(emu, bc):
    """n = 128 (the temporally blocked smoother) with inhomogeneous Dirichlet values on one pair of sides and periodic
    boundaries on the other: halo cells that are periodic images must index the boundary values with the wrapped
    row / column.  Found by scripts/fuzz_mg_emulated.py; the fixed-case tests had inhomogeneous values only at n = 64."""
    n = 128
    rng = np.random.default_rng(7)
    o = oracle.MG(n, bc=bc)
    m = EmuMG(emu, n, bc, 0.0, -1.0, True)
    vals = [rng.standard_normal(n + 2) if b == "dirichlet" else None for b in bc]
    for side, v in zip(("xl", "xr", "yl", "yr"), vals):
        if v is not None:
            o.set_bc_values(side, v)
    m.set_bc_values(*vals)
    fine = o.nlevels - 1
    f = rng.standard_normal((n + 2, n + 2))
    o.plane(fine, "f")[:] = f
    m.plane(fine, "f")[:] = f
    o.smooth(fine, 7)
    m.ck(emu.p2b_mg_smooth(m.h, fine, 7, None))
    assert np.array_equal(m.plane(fine, "v")[1:-1, 1:-1], o.plane(fine, "v")[1:-1, 1:-1])
    m.close()


@pytest.mark.parametrize("kind,size,n,split", [("dirichlet", 2, 128, 64), ("xper_inhom", 2, 128, 32)])
def test_slabs_as_threads_on_the_emulated_device(kind, size, n, split):
    """ranks as host threads of one process (parallel.LocalSlabGroup): the hardware test of the same name minus the
    hardware -- the peer-memory protocol of the decomposed V-cycle over plain pointers"""
    import emu_device
    from test_gpu_mg import slab_case, slabs_in_one_process
    kw, rhs = slab_case(kind)
    with emu_device.emulated_device():
        full, cyc, one, cyc1 = slabs_in_one_process(size, n, split, kw, rhs)
    assert cyc == cyc1 and np.array_equal(full, one)

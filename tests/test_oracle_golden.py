"""CPU: the oracle (oracle/pyro_oracle.c) against the fixtures produced by running the unmodified
reference (tests/golden/make_golden.py) -- this is what pins the oracle.  Tolerances: the
compressible step matches the reference to ~1e-15 per step (np.dot / pow ordering), 1e-12 after
tens of steps; multigrid solutions are bit-identical; dt is bit-identical."""
import os

import numpy as np
import pytest

import oracle
from conftest import rel_l2, state_errors
from golden_util import load_comp, load_flow, load_mg, load_mgvc
from oracle_runs import (run_advection, run_burgers, run_compressible, run_diffusion, run_incompressible,
                         run_lm_atm)


@pytest.mark.parametrize("name", ["sedov64", "quad64", "sod_x", "kh32", "acoustic64", "advect32", "gresho40",
                                  "bubble32", "rt16", "hse16", "rt16_reflect", "sedov32_cgf", "quad32_cgf_walls",
                                  "heating32", "plume32", "convection16", "rt2_48", "rt_multimode16", "ramp64", "gresho40_lm", "sedov32_lm", "sedov_sph32", "advect_sph32"])
def test_compressible_run_matches_reference(name):
    z, rp, inputs = load_comp(name)
    U, dts, ng = run_compressible(z, rp, fix_dt=inputs.get("driver.fix_dt", -1.0))
    ref = z["U"]
    v = (slice(ng, -ng), slice(ng, -ng))
    assert np.allclose(dts, z["dts"], rtol=1e-12, atol=0)
    assert dts[0] == z["dts"][0]          # first dt: pure CFL reduction, bit-exact
    assert max(state_errors(U[v], ref[v], rp["eos.gamma"])) < 1e-12


@pytest.mark.parametrize("name", ["poisson_dirichlet_64", "poisson_dirichlet_256", "poisson_periodic_64",
                                  "helmholtz_neumann_64", "poisson_mixed_128"])
def test_mg_solve_matches_reference(name):
    z = load_mg(name)
    o = oracle.MG(int(z["nx"]), bc=tuple(str(b) for b in z["bc"]), alpha=float(z["alpha"]), beta=float(z["beta"]))
    o.init_zeros()
    o.init_RHS(z["f"])
    o.solve(rtol=float(z["rtol"]))
    assert o.num_cycles == int(z["num_cycles"])
    assert np.array_equal(o.get_solution(), z["v"])
    assert abs(o.residual_error - float(z["residual_error"])) <= 1e-12 * float(z["residual_error"]) + 1e-25
    assert abs(o.source_norm - float(z["source_norm"])) <= 1e-14 * float(z["source_norm"])


def test_mg_inhomogeneous_dirichlet_matches_reference():
    z = load_mg("poisson_inhom_64")
    nx = int(z["nx"])
    o = oracle.MG(nx)
    c = (np.arange(nx + 2) - 0.5) / nx
    o.set_bc_values("xl", c ** 2); o.set_bc_values("xr", 1.0 + c)
    o.set_bc_values("yl", c); o.set_bc_values("yr", 1.0 + c ** 2)
    o.init_zeros()
    o.init_RHS(z["f"])
    o.solve(rtol=float(z["rtol"]))
    assert o.num_cycles == int(z["num_cycles"])
    assert np.array_equal(o.get_solution(), z["v"])


@pytest.mark.parametrize("name", ["dirichlet_64", "periodic_64", "constant_32", "dirichlet_128"])
def test_mg_variable_coeff_solve_matches_reference(name):
    """VarCoeffCCMG2d (variable_coeff_MG.py:24-213) run by the reference on its mg_test_vc_* setups"""
    z = load_mgvc(name)
    o = oracle.MG(int(z["nx"]), bc=tuple(str(b) for b in z["bc"]), alpha=0.0, beta=0.0)
    o.set_coeffs(z["coeffs"], tuple(str(b) for b in z["coeffs_bc"]))
    assert np.array_equal(o.coef_plane(2, "ex"), z["ex_coarse"])
    assert np.array_equal(o.coef_plane(2, "ey"), z["ey_coarse"])
    o.init_zeros()
    o.init_RHS(z["f"])
    o.solve(rtol=float(z["rtol"]))
    assert o.num_cycles == int(z["num_cycles"])
    assert np.array_equal(o.get_solution(), z["v"])
    n = int(z["nx"])
    assert np.array_equal(o.plane(o.nlevels - 1, "r")[1:n + 1, 1:n + 1], z["r"][1:n + 1, 1:n + 1])
    assert abs(o.residual_error - float(z["residual_error"])) <= 1e-12 * float(z["residual_error"]) + 1e-25


def test_mg_convergence_table():
    """pyro/multigrid/tests/mg_convergence.txt: L2 error vs the analytic solution, N = 16 .. 256"""
    kat = np.load(__import__("os").path.join(__import__("golden_util").GOLDEN, "ref_kats.npz"))
    for n, err in kat["mg_convergence"]:
        n = int(n)
        o = oracle.MG(n)
        x = (np.arange(n + 2) - 0.5) / n
        X, Y = np.meshgrid(x, x, indexing="ij")
        o.init_zeros()
        o.init_RHS(-2.0 * ((1.0 - 6.0 * X ** 2) * Y ** 2 * (1.0 - Y ** 2) + (1.0 - 6.0 * Y ** 2) * X ** 2 * (1.0 - X ** 2)))
        o.solve(rtol=1.e-11)
        e = o.get_solution() - (X ** 2 - X ** 4) * (Y ** 4 - Y ** 2)
        assert abs(o.norm(e) - err) <= 1e-5 * err     # table is printed with 6 significant digits


def test_ghost_fill_matches_reference_int_and_all_types():
    z = np.load(__import__("os").path.join(__import__("golden_util").GOLDEN, "mesh_bcs.npz"))
    for ng in (1, 4):
        for t in ("outflow", "periodic", "reflect-even", "reflect-odd"):
            a = z[f"base_ng{ng}"].astype(np.int64).copy()
            oracle.fill_ghost(a, ng, (t,) * 4)
            assert np.array_equal(a, z[f"{t}_ng{ng}"])
            b = z[f"base_ng{ng}"].astype(np.float64).copy()
            oracle.fill_ghost(b, ng, (t,) * 4)
            assert np.array_equal(b, z[f"{t}_ng{ng}"].astype(np.float64))


@pytest.mark.parametrize("fname", ["incomp_shear32.npz", "incomp_shear64.npz", "incomp_converge32.npz"])
def test_incompressible_run_matches_reference(fname):
    """Pyro("incompressible") fixtures: the oracle's evolve (explicit part + two multigrid projections)
    stepped with the recorded dts reproduces all six state planes bit for bit"""
    z, rp, _ = load_flow(fname)
    P = run_incompressible(z, rp)
    assert np.array_equal(P, z["P"])


@pytest.mark.parametrize("fname", ["burgers_test.npz", "burgers_converge32.npz", "burgers_tophat32.npz"])
def test_burgers_run_matches_reference(fname):
    z, rp, _ = load_flow(fname)
    u, v = run_burgers(z, rp)
    assert np.array_equal(u, z["P"][0]) and np.array_equal(v, z["P"][1])


@pytest.mark.parametrize("fname", ["advection_smooth64.npz", "advection_tophat32.npz"])
def test_advection_run_matches_reference(fname):
    """Pyro("advection") fixtures; smooth64 is BASELINE config 1 (81 steps to t = 1) with the known answers
    SURVEY.md quotes for the reference: sum 4.310466040637315e+03, min 0.9999998946441166, max 1.960068731417340"""
    z, rp, _ = load_flow(fname)
    ng = int(z["ng"])
    a = run_advection(z, rp)
    v = (slice(ng, -ng), slice(ng, -ng))
    assert np.array_equal(a[v], z["P"][0][v])
    if fname == "advection_smooth64.npz":
        kat = np.load(__import__("os").path.join(__import__("golden_util").GOLDEN, "ref_kats.npz"))
        assert len(z["dts"]) == 81
        assert float(np.sum(a[v])) == float(kat["advection_smooth_sum"]) == 4.310466040637315e+03
        assert a[v].min() == 0.9999998946441166 and abs(a[v].max() - 1.960068731417340) < 1e-15


@pytest.mark.parametrize("fname", ["diffusion_gaussian64.npz", "diffusion_gaussian32_mixed.npz"])
def test_diffusion_run_matches_reference(fname):
    """Pyro("diffusion") fixtures: one Crank-Nicolson multigrid solve per step"""
    z, rp, _ = load_flow(fname)
    phi = run_diffusion(z, rp)
    assert np.array_equal(phi[1:-1, 1:-1], z["P"][0][1:-1, 1:-1])


@pytest.mark.parametrize("fname", ["lm_bubble32.npz", "lm_bubble64_lim1.npz"])
def test_lm_atm_run_matches_reference(fname):
    """Pyro("lm_atm") fixtures (bubble): the oracle's evolve -- numba interface routines restated, two
    variable-coefficient multigrid projections -- reproduces all eight state planes bit for bit"""
    z, rp, _ = load_flow(fname)
    S = run_lm_atm(z, rp)
    assert np.array_equal(S, z["P"])


# ---- the regression files the reference itself stores (read with tests/h5lite.py, tests/golden/make_h5_golden.py) -------
def _refh5(name):
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"refh5_{name}.npz"))
    meta = dict(s.split("=", 1) for s in z["meta"])
    return z, meta


def mg_test_simple_rhs(n):
    """right-hand side of pyro/multigrid/examples/mg_test_simple.py on the (n + 2)^2 cell centres"""
    x = (np.arange(n + 2) - 0.5) / n
    X, Y = np.meshgrid(x, x, indexing="ij")
    return -2.0 * ((1.0 - 6.0 * X ** 2) * Y ** 2 * (1.0 - Y ** 2) + (1.0 - 6.0 * Y ** 2) * X ** 2 * (1.0 - X ** 2))


def test_h5_reader_against_the_reference_files():
    """tests/h5lite.py on the reference's own snapshots: groups, attributes (incl. variable-length strings), contiguous
    datasets; the committed fixtures hold exactly these bytes"""
    ref = "/root/reference/pyro"
    if not os.path.isdir(ref):
        pytest.skip("the reference tree is not on this box")
    import h5lite
    f = h5lite.File(os.path.join(ref, "compressible/tests/sod_x_0076.h5"))
    assert f.keys() == ["BC", "aux", "grid", "runtime parameters", "state"]
    assert f.attrs()["nsteps"] == 76 and f.attrs()["problem"] == "sod" and f.attrs("grid")["nx"] == 128
    assert f.attrs("runtime parameters")["compressible.riemann"] == "HLLC"
    assert f.attrs("state/y-momentum")["ylb"] == "reflect-odd"
    z, _ = _refh5("sod_x_0076")
    for name in ("density", "energy", "x-momentum", "y-momentum"):
        assert np.array_equal(f[f"state/{name}/data"], z[name.replace("-", "_")])
    g = h5lite.File(os.path.join(ref, "multigrid/tests/mg_poisson_dirichlet.h5"))
    zm, _ = _refh5("mg_poisson_dirichlet")
    assert np.array_equal(g["state/v/data"], zm["v"]) and g["state/v/data"].shape == (256, 256)


def test_oracle_reproduces_the_stored_sod_golden():
    """pyro/compressible/tests/sod_x_0076.h5 -- the file the reference's own regression test compares against"""
    z, rp, inputs = load_comp("sod_x")
    stored, meta = _refh5("sod_x_0076")
    assert int(meta["nsteps"]) == int(z["n"]) == 76 and float(meta["time"]) == pytest.approx(float(z["t"]), rel=1e-14)
    U, dts, ng = run_compressible(z, rp, fix_dt=inputs.get("driver.fix_dt", -1.0))
    v = (slice(ng, -ng), slice(ng, -ng))
    for k, name in enumerate(("density", "energy", "x_momentum", "y_momentum")):
        assert np.abs(U[v][..., k] - stored[name]).max() <= 2e-14, name       # another machine's libm / numba: round-off


def test_oracle_reproduces_the_stored_multigrid_golden():
    """pyro/multigrid/tests/mg_poisson_dirichlet.h5: the right-hand side bit for bit, then the solution bit for bit"""
    import hashlib
    stored, meta = _refh5("mg_poisson_dirichlet")
    n = int(meta["grid.nx"])
    f = mg_test_simple_rhs(n)
    assert hashlib.sha256(np.ascontiguousarray(f[1:-1, 1:-1]).tobytes()).hexdigest() == str(stored["f_sha256"])
    o = oracle.MG(n)
    o.init_zeros()
    o.init_RHS(f)
    o.solve(rtol=1.e-11)
    assert np.array_equal(o.get_solution()[1:-1, 1:-1], stored["v"])
    assert np.abs(o.plane(o.nlevels - 1, "r")[1:-1, 1:-1] - stored["r"]).max() <= 1e-18 + 1e-12 * np.abs(stored["r"]).max()


# ---- every other stored regression file of the paths built here (tests/golden/pin_stored_goldens.py; the full-resolution
#      table of that script's run is profiles/r2_oracle_vs_stored_goldens.txt) ---------------------------------------------
def _stored_fixture(case):
    import golden_util
    z = np.load(os.path.join(golden_util.GOLDEN, f"refh5_{case}.npz"))
    rp = {s.split("=", 1)[0]: golden_util._parse(s.split("=", 1)[1]) for s in z["rp"]}
    return z, rp, int(z["stride"])


@pytest.mark.parametrize("case", ["quad", "rt"])
def test_oracle_reproduces_the_stored_compressible_goldens(case):
    """pyro/compressible/tests/quad_unsplit_0606.h5 (256^2, 606 steps; every second cell kept in the fixture) and
    rt_0945.h5 (64 x 192, gravity, "hse" boundaries, 945 steps through the instability's growth).  The reference's own
    criterion is np.allclose(rtol=1e-12) (util/compare.py:59, i.e. plus atol 1e-8); the unmodified reference run in this
    container is itself 1e-13 .. 5e-13 away from these files.  Bar here: 2e-12 of each variable's maximum."""
    z, rp, stride = _stored_fixture(case)
    U, dts, ng = run_compressible(z, rp)
    assert len(dts) == int(z["n"]) and np.allclose(dts, z["dts"], rtol=1e-11, atol=0)
    v = (slice(ng, -ng, stride), slice(ng, -ng, stride))
    for k, name in enumerate(z["names"]):
        assert np.abs(U[v][..., k] - z["stored"][k]).max() <= 2e-12 * np.abs(z["stored"][k]).max(), name


@pytest.mark.parametrize("case,runner", [("advection", run_advection), ("burgers", run_burgers), ("diffusion", run_diffusion)])
def test_oracle_reproduces_the_stored_flow_goldens_bit_for_bit(case, runner):
    """pyro/advection/tests/smooth_0040.h5, burgers/tests/test_0051.h5 (128^2, 51 steps), diffusion/tests/gaussian_0164.h5
    (128^2, 164 Crank-Nicolson multigrid solves): the stored planes, bit for bit"""
    z, rp, stride = _stored_fixture(case)
    out = runner(z, rp)
    out = [out] if isinstance(out, np.ndarray) and out.ndim == 2 else list(out)
    ng = int(z["ng"])
    assert stride == 1 and len(out) == len(z["stored"])
    for a, s in zip(out, z["stored"]):
        assert np.array_equal(a[ng:-ng, ng:-ng], s)


def test_oracle_reproduces_the_stored_incompressible_golden():
    """pyro/incompressible/tests/shear_128_0216.h5: 216 steps, two multigrid projections each.  The oracle follows the
    reference run here bit for bit; that run is 6e-15 (velocities) .. 1.8e-12 (phi) away from the file made on another
    machine.  Bar: 2e-12 of each variable's maximum."""
    z, rp, stride = _stored_fixture("incomp")
    P = run_incompressible(z, rp)
    ng = int(z["ng"])
    for a, s, name in zip(P, z["stored"], z["names"]):
        assert np.abs(a[ng:-ng, ng:-ng] - s).max() <= 2e-12 * np.abs(s).max(), name


@pytest.mark.parametrize("case", ["mgvc_dirichlet", "mgvc_periodic"])
def test_oracle_reproduces_the_stored_variable_coefficient_goldens(case):
    """pyro/multigrid/tests/mg_vc_poisson_{dirichlet,periodic}.h5, 512^2 (pyro/test.py:143-151): coefficients and
    right-hand side from the examples' formulas (the stored right-hand side differs from this machine's by one ulp of
    cos/sin), 7 V-cycles, the solution within 1e-13 of the file (every fourth cell kept in the fixture; the reference run
    here: 7e-16 / 7e-15)"""
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"refh5_{case}.npz"))
    n, stride, pi = int(z["nx"]), int(z["stride"]), np.pi
    x = (np.arange(n + 2) - 0.5) / n
    X, Y = np.meshgrid(x, x, indexing="ij")
    o = oracle.MG(n, bc=tuple(str(b) for b in z["bc"]), alpha=0.0, beta=0.0)
    o.set_coeffs(2.0 + np.cos(2.0 * pi * X) * np.cos(2.0 * pi * Y), tuple(str(b) for b in z["coeffs_bc"]))
    o.init_zeros()
    o.init_RHS(-16.0 * pi ** 2 * (np.cos(2 * pi * X) * np.cos(2 * pi * Y) + 1) * np.sin(2 * pi * X) * np.sin(2 * pi * Y))
    o.solve(rtol=float(z["rtol"]))
    assert o.num_cycles == int(z["num_cycles"]) == 7
    assert np.abs(o.get_solution()[1:-1:stride, 1:-1:stride] - z["stored_v"]).max() <= 1e-13

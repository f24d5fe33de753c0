"""CPU: the oracle (oracle/pyro_oracle.c) against the fixtures produced by running the unmodified
reference (tests/golden/make_golden.py) -- this is what pins the oracle.  Tolerances: the
compressible step matches the reference to ~1e-15 per step (np.dot / pow ordering), 1e-12 after
tens of steps; multigrid solutions are bit-identical; dt is bit-identical."""
import os

import numpy as np
import pytest

import oracle
from conftest import rel_l2, state_errors
from golden_util import load_comp, load_flow, load_mg, load_mgvc, var_bcs


def _run_oracle(z, rp, nsteps=None, fix_dt=-1.0):
    """the driver loop of pyro_sim.py:241-256 over the oracle: fill_BC_all (variable by variable, the "hse"
    user boundary after the standard fill, like CellCenterData2d.fill_BC), compute_timestep, evolve"""
    ng = int(z["ng"])
    P = oracle.to_planes(z["U0"])
    nx, ny = rp["mesh.nx"], rp["mesh.ny"]
    dx = (rp["mesh.xmax"] - rp["mesh.xmin"]) / nx
    dy = (rp["mesh.ymax"] - rp["mesh.ymin"]) / ny
    grav = rp.get("compressible.grav", 0.0)
    gamma = rp["eos.gamma"]
    bcs = var_bcs(rp)
    geom = None
    if rp.get("mesh.grid_type", "Cartesian2d") == "SphericalPolar":
        geom = oracle.spherical_geometry(nx, ny, ng, rp["mesh.xmin"], rp["mesh.xmax"], rp["mesh.ymin"], rp["mesh.ymax"])
    xc = (np.arange(nx + 2 * ng) + 0.5 - ng) * dx + rp["mesh.xmin"]
    yc = (np.arange(ny + 2 * ng) + 0.5 - ng) * dy + rp["mesh.ymin"]
    prm = oracle.comp_params(gamma=gamma, z0=rp["compressible.z0"], z1=rp["compressible.z1"],
                             delta=rp["compressible.delta"], cvisc=rp["compressible.cvisc"],
                             limiter=rp["compressible.limiter"], use_flattening=rp["compressible.use_flattening"],
                             grav=grav, src_bcs=bcs, riemann=rp.get("compressible.riemann", "HLLC"),
                             xl_solid=int(rp["mesh.xlboundary"] == "reflect"), yl_solid=int(rp["mesh.ylboundary"] == "reflect"),
                             heat_rate=float(z["heat_rate"]) if "heat_rate" in z else 0.0,
                             heat_profile=z["heat_profile"] if "heat_profile" in z else None,
                             sponge=(rp["sponge.sponge_rho_begin"], rp["sponge.sponge_rho_full"], rp["sponge.sponge_timescale"])
                             if rp.get("sponge.do_sponge", 0) else None, geom=geom)
    ambient = None
    if "ambient" in z:        # compressible/BC.py:142-168: constant state above the top boundary
        ar, au, av, ap = (float(x) for x in z["ambient"])
        ambient = [ar, ap / (gamma - 1.0) + 0.5 * ar * (au ** 2 + av ** 2), ar * au, ar * av]
    small_dens = rp.get("compressible.small_dens", -1.e200)
    t, dt_old, dts = 0.0, None, []
    nsteps = len(z["dts"]) if nsteps is None else nsteps
    for n in range(nsteps):
        for k in range(4):
            oracle.fill_ghost(P[k], ng, bcs[k])
            for side in ("ylb", "yrb"):
                if bcs[k][2 + (side == "yrb")] == "hse":
                    oracle.fill_hse(P, ng, dy, grav, gamma, k, side)
                if bcs[k][2 + (side == "yrb")] == "ambient":
                    P[k][:, ng + ny:] = ambient[k]
            for s_, side in enumerate(("xlb", "xrb", "ylb", "yrb")):      # user boundaries after the standard ones, in this order
                if bcs[k][s_] == "ramp":
                    oracle.fill_ramp(P[k], k, side, ng, xc, yc, dx, dy, t, gamma)
        dt = oracle.cfl_dt(oracle.from_planes(P), ng, dx, dy, gamma, rp["driver.cfl"]) if geom is None else \
            oracle.cfl_dt_spherical(oracle.from_planes(P), gamma, rp["driver.cfl"], geom)
        # NullSimulation.compute_timestep (simulation_null.py:222-244)
        dt = rp["driver.init_tstep_factor"] * dt if n == 0 else min(rp["driver.max_dt_change"] * dt_old, dt)
        dt_old = dt
        if fix_dt > 0.0:
            dt = fix_dt
        if t + dt > rp["driver.tmax"]:
            dt = rp["driver.tmax"] - t
        P[0][ng:-ng, ng:-ng] = np.maximum(P[0][ng:-ng, ng:-ng], small_dens)     # clean_state (simulation.py:296, 452-456)
        oracle.compressible_step(P, ng, dx, dy, dt, prm, planes=True)
        t += dt
        dts.append(dt)
    U = oracle.from_planes(P)
    return U, np.array(dts), ng


@pytest.mark.parametrize("name", ["sedov64", "quad64", "sod_x", "kh32", "acoustic64", "advect32", "gresho40",
                                  "bubble32", "rt16", "hse16", "rt16_reflect", "sedov32_cgf", "quad32_cgf_walls",
                                  "heating32", "plume32", "convection16", "rt2_48", "rt_multimode16", "ramp64", "gresho40_lm", "sedov32_lm", "sedov_sph32", "advect_sph32"])
def test_compressible_run_matches_reference(name):
    z, rp, inputs = load_comp(name)
    U, dts, ng = _run_oracle(z, rp, fix_dt=inputs.get("driver.fix_dt", -1.0))
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
    ng = int(z["ng"])
    P = np.ascontiguousarray(z["P0"])
    bc = (rp["mesh.xlboundary"], rp["mesh.xrboundary"], rp["mesh.ylboundary"], rp["mesh.yrboundary"])
    assert bc == ("periodic",) * 4
    for dt in z["dts"]:
        for k in range(6):          # the driver's fill_BC_all before every step (pyro_sim.py:241-256)
            oracle.fill_ghost(P[k], ng, bc)
        oracle.incomp_evolve(P, ng, float(dt), limiter=rp["incompressible.limiter"], proj_type=rp["incompressible.proj_type"],
                             vel_bc=(bc, bc), phi_bc=bc, xmin=rp["mesh.xmin"], xmax=rp["mesh.xmax"],
                             ymin=rp["mesh.ymin"], ymax=rp["mesh.ymax"])
    assert np.array_equal(P, z["P"])


@pytest.mark.parametrize("fname", ["burgers_test.npz", "burgers_converge32.npz", "burgers_tophat32.npz"])
def test_burgers_run_matches_reference(fname):
    z, rp, _ = load_flow(fname)
    ng = int(z["ng"])
    u, v = z["P0"][0].copy(), z["P0"][1].copy()
    n = rp["mesh.nx"]
    dx = (rp["mesh.xmax"] - rp["mesh.xmin"]) / n
    bc = (rp["mesh.xlboundary"], rp["mesh.xrboundary"], rp["mesh.ylboundary"], rp["mesh.yrboundary"])
    for step, dt in enumerate(z["dts"]):
        oracle.fill_ghost(u, ng, bc)
        oracle.fill_ghost(v, ng, bc)
        # burgers/simulation.py:41-58 (then the driver's first-step factor and growth limit, both inactive here)
        raw = rp["driver.cfl"] * min(dx / max(np.abs(u).max(), 1.e-12), dx / max(np.abs(v).max(), 1.e-12))
        if rp["driver.fix_dt"] > 0:
            assert float(dt) == rp["driver.fix_dt"]
        elif step > 0 and z["t"] > 0:
            assert raw >= float(dt) * (1 - 1e-15)
        u, v = oracle.burgers_evolve(u, v, ng, dx, dx, float(dt), rp["advection.limiter"])
    assert np.array_equal(u, z["P"][0]) and np.array_equal(v, z["P"][1])


@pytest.mark.parametrize("fname", ["advection_smooth64.npz", "advection_tophat32.npz"])
def test_advection_run_matches_reference(fname):
    """Pyro("advection") fixtures; smooth64 is BASELINE config 1 (81 steps to t = 1) with the known answers
    SURVEY.md quotes for the reference: sum 4.310466040637315e+03, min 0.9999998946441166, max 1.960068731417340"""
    z, rp, _ = load_flow(fname)
    ng, n = int(z["ng"]), rp["mesh.nx"]
    a = z["P0"][0].copy()
    dx = (rp["mesh.xmax"] - rp["mesh.xmin"]) / n
    dy = (rp["mesh.ymax"] - rp["mesh.ymin"]) / rp["mesh.ny"]
    bc = (rp["mesh.xlboundary"], rp["mesh.xrboundary"], rp["mesh.ylboundary"], rp["mesh.yrboundary"])
    for dt in z["dts"]:
        oracle.fill_ghost(a, ng, bc)
        a = oracle.advection_evolve(a, ng, dx, dy, float(dt), rp["advection.u"], rp["advection.v"], rp["advection.limiter"])
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
    phi = np.ascontiguousarray(z["P0"][0])
    bc = (rp["mesh.xlboundary"], rp["mesh.xrboundary"], rp["mesh.ylboundary"], rp["mesh.yrboundary"])
    for dt in z["dts"]:
        oracle.diffusion_evolve(phi, float(dt), rp["diffusion.k"], bc, rp["mesh.xmin"], rp["mesh.xmax"],
                                rp["mesh.ymin"], rp["mesh.ymax"])
    assert np.array_equal(phi[1:-1, 1:-1], z["P"][0][1:-1, 1:-1])


def _lm_setup(z, rp):
    names = [str(n) for n in z["names"]]
    bc = (rp["mesh.xlboundary"], rp["mesh.xrboundary"], rp["mesh.ylboundary"], rp["mesh.yrboundary"])
    assert bc == ("periodic", "periodic", "reflect", "outflow")      # the setup the fixtures were generated with
    even = ("periodic", "periodic", "reflect-even", "outflow")
    odd_y = ("periodic", "periodic", "reflect-odd", "outflow")
    phi_bc = ("periodic", "periodic", "neumann", "dirichlet")
    fills = dict(zip(names, (even, even, odd_y, even, phi_bc, phi_bc, even, even)))
    prm = oracle.lm_params(rp["mesh.nx"], grav=rp["lm-atmosphere.grav"], gamma=rp["eos.gamma"],
                           limiter=rp["lm-atmosphere.limiter"], proj_type=rp["lm-atmosphere.proj_type"],
                           xmin=rp["mesh.xmin"], xmax=rp["mesh.xmax"], ymin=rp["mesh.ymin"], ymax=rp["mesh.ymax"])
    return names, fills, prm


@pytest.mark.parametrize("fname", ["lm_bubble32.npz", "lm_bubble64_lim1.npz"])
def test_lm_atm_run_matches_reference(fname):
    """Pyro("lm_atm") fixtures (bubble): the oracle's evolve -- numba interface routines restated, two
    variable-coefficient multigrid projections -- reproduces all eight state planes bit for bit"""
    z, rp, _ = load_flow(fname)
    ng = int(z["ng"])
    names, fills, prm = _lm_setup(z, rp)
    S = np.ascontiguousarray(z["P0"])
    base = np.ascontiguousarray(z["base"])
    for n, dt in enumerate(z["dts"]):
        for k, name in enumerate(names):
            oracle.fill_ghost(S[k], ng, fills[name])
        raw = oracle.lm_timestep(S, base, prm, rp["driver.cfl"])
        assert raw >= float(dt) * (1 - 1e-15)          # the driver only ever shrinks the method's dt
        oracle.lm_evolve(S, base, prm, float(dt))
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
    U, dts, ng = _run_oracle(z, rp, fix_dt=inputs.get("driver.fix_dt", -1.0))
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

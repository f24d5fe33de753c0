"""CPU, build container only: the oracle against the LIVE reference imported from /root/reference
(skipped on the GPU box, where the reference tree does not exist)."""
import numpy as np
import pytest

import oracle
import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")


def test_compressible_stages_bit_identical():
    p = ref_shim.make_sim("compressible", "sedov", {"mesh.nx": 40, "mesh.ny": 32, "sedov.r_init": 0.1,
                                                    "driver.tmax": 10.0})
    import pyro.compressible.unsplit_fluxes as flx
    sim = p.sim
    g = sim.cc_data.grid
    for _ in range(3):
        sim.cc_data.fill_BC_all()
        sim.compute_timestep()
        U0 = np.asarray(sim.cc_data.data).copy()
        raw_dt = sim.dt
        sim.method_compute_timestep()
        assert sim.dt == oracle.cfl_dt(U0, g.ng, g.dx, g.dy, 1.4, 0.8)
        sim.dt = raw_dt
        hat = [np.asarray(x).copy() for x in flx.interface_states(sim.cc_data, sim.rp, sim.ivars, sim.tc, sim.dt)]
        Unew, st = oracle.compressible_step(U0, g.ng, g.dx, g.dy, sim.dt, stages=True)
        for nm, a in zip(["Uxl_hat", "Uxr_hat", "Uyl_hat", "Uyr_hat"], hat):
            assert np.array_equal(st[nm], a), nm
        sim.evolve()
        v = (slice(g.ilo, g.ihi + 1), slice(g.jlo, g.jhi + 1))
        assert np.abs(Unew[v] - np.asarray(sim.cc_data.data)[v]).max() < 1e-14


def test_mg_bit_identical():
    ref_shim.load()
    import pyro.multigrid.MG as MG
    a = MG.CellCenterMG2d(32, 32, xl_BC_type="neumann", xr_BC_type="neumann", yl_BC_type="dirichlet",
                          yr_BC_type="dirichlet", alpha=0.5, beta=0.02)
    f = np.cos(3 * a.x2d) * a.y2d
    a.init_zeros(); a.init_RHS(f); a.solve(rtol=1e-12)
    o = oracle.MG(32, bc=("neumann", "neumann", "dirichlet", "dirichlet"), alpha=0.5, beta=0.02)
    o.init_zeros(); o.init_RHS(np.asarray(f)); o.solve(rtol=1e-12)
    assert o.num_cycles == a.num_cycles
    assert np.array_equal(o.get_solution(), np.asarray(a.get_solution()))


def test_mg_variable_coeff_bit_identical():
    """every level's edge coefficients and the solve of a random-coefficient problem"""
    ref_shim.load()
    import pyro.mesh.boundary as bnd
    import pyro.multigrid.variable_coeff_MG as VMG
    from pyro.mesh import patch
    rng = np.random.default_rng(5)
    nx = 32
    g = patch.Grid2d(nx, nx, ng=1)
    d = patch.CellCenterData2d(g)
    bc_c = bnd.BC(xlb="neumann", xrb="neumann", ylb="periodic", yrb="periodic")
    d.register_var("c", bc_c)
    d.create()
    c = d.get_var("c")
    c[:, :] = 0.5 + rng.random((nx + 2, nx + 2))
    a = VMG.VarCoeffCCMG2d(nx, nx, xl_BC_type="dirichlet", xr_BC_type="neumann", yl_BC_type="periodic",
                           yr_BC_type="periodic", coeffs=c, coeffs_bc=bc_c, verbose=0)
    f = np.sin(2 * np.pi * a.y2d) * (a.x2d + 0.3)
    a.init_zeros(); a.init_RHS(f); a.solve(rtol=1e-11)
    o = oracle.MG(nx, bc=("dirichlet", "neumann", "periodic", "periodic"), alpha=0.0, beta=0.0)
    o.set_coeffs(np.asarray(c), ("neumann", "neumann", "periodic", "periodic"))
    for lev in range(o.nlevels):
        assert np.array_equal(o.coef_plane(lev, "ex"), np.asarray(a.edge_coeffs[lev].x))
        assert np.array_equal(o.coef_plane(lev, "ey"), np.asarray(a.edge_coeffs[lev].y))
    o.init_zeros(); o.init_RHS(np.asarray(f)); o.solve(rtol=1e-11)
    assert o.num_cycles == a.num_cycles
    assert np.array_equal(o.get_solution(), np.asarray(a.get_solution()))


def test_incompressible_stages_bit_identical():
    """mac_vels / states (incomp_interface.py) and the full evolve against the live reference"""
    ref_shim.load()
    from pyro.incompressible import incomp_interface
    from pyro.mesh import reconstruction
    p = ref_shim.make_sim("incompressible", "shear", {"mesh.nx": 32, "mesh.ny": 32, "driver.max_steps": 100})
    sim = p.sim
    g = sim.cc_data.grid
    names = ["x-velocity", "y-velocity", "phi-MAC", "phi", "gradp_x", "gradp_y"]
    for _ in range(3):
        sim.cc_data.fill_BC_all()
        sim.compute_timestep()
        dt = sim.dt
        P = np.ascontiguousarray(np.stack([np.asarray(sim.cc_data.get_var(n)) for n in names]))
        u, v, gx, gy = (sim.cc_data.get_var(n) for n in ("x-velocity", "y-velocity", "gradp_x", "gradp_y"))
        ld = [reconstruction.limit(q, g, d, 2) for q, d in ((u, 1), (v, 1), (u, 2), (v, 2))]
        um, vm = incomp_interface.mac_vels(g, dt, u, v, ld[0], ld[1], ld[2], ld[3], gx, gy)
        oum, ovm = oracle.incomp_mac_vels(P[0], P[1], P[4], P[5], g.ng, g.dx, g.dy, dt, 2)
        assert np.array_equal(oum, np.asarray(um)) and np.array_equal(ovm, np.asarray(vm))
        import pyro.mesh.array_indexer as ai
        st = incomp_interface.states(g, dt, u, v, ld[0], ld[1], ld[2], ld[3], gx, gy,
                                     ai.ArrayIndexer(d=um, grid=g), ai.ArrayIndexer(d=vm, grid=g))
        ost = oracle.incomp_states(P[0], P[1], P[4], P[5], g.ng, g.dx, g.dy, dt, 2, oum, ovm)
        for a, b in zip(ost, st):
            assert np.array_equal(a, np.asarray(b))
        oracle.incomp_evolve(P, g.ng, dt)
        sim.evolve()
        for k, n in enumerate(names):
            assert np.array_equal(P[k], np.asarray(sim.cc_data.get_var(n))), n


def test_lm_atm_bit_identical():
    """lm_atm: timestep, preevolve (initial projection + throw-away step) and steps against the live reference"""
    ref_shim.load()
    import pyro.lm_atm.simulation as lms
    captured = {}
    orig_pre = lms.Simulation.preevolve

    def pre(self):
        captured["S"] = np.ascontiguousarray(np.stack([np.asarray(self.cc_data.get_var(n)) for n in oracle.LM_VARS]))
        orig_pre(self)
    lms.Simulation.preevolve = pre
    try:
        p = ref_shim.make_sim("lm_atm", "bubble", {"mesh.nx": 32, "mesh.ny": 32, "driver.max_steps": 100})
    finally:
        lms.Simulation.preevolve = orig_pre
    sim = p.sim
    g = sim.cc_data.grid
    base = np.ascontiguousarray(np.stack([sim.base[k].d for k in ("rho0", "p0", "beta0", "beta0-edges")]))
    prm = oracle.lm_params(g.nx)
    S = captured["S"].copy()
    oracle.lm_initial_projection(S, base, prm)
    saved = S.copy()
    oracle.lm_evolve(S, base, prm, oracle.lm_timestep(S, base, prm, 0.8))
    saved[6:8] = S[6:8]
    S = saved
    state = lambda: np.stack([np.asarray(sim.cc_data.get_var(n)) for n in oracle.LM_VARS])
    assert np.array_equal(S, state())
    bcn = {n: (sim.cc_data.BCs[n].xlb, sim.cc_data.BCs[n].xrb, sim.cc_data.BCs[n].ylb, sim.cc_data.BCs[n].yrb)
           for n in oracle.LM_VARS}
    for _ in range(2):
        sim.cc_data.fill_BC_all()
        for k, n in enumerate(oracle.LM_VARS):
            oracle.fill_ghost(S[k], g.ng, bcn[n])
        sim.method_compute_timestep()
        assert sim.dt == oracle.lm_timestep(S, base, prm, 0.8)
        sim.compute_timestep()
        oracle.lm_evolve(S, base, prm, sim.dt)
        sim.evolve()
        assert np.array_equal(S, state())


def test_fv2d_conversions_bit_identical():
    """mesh/fv.py: to_centers (both positivity settings) and the from_centers stencil against the live reference on
    random data (pyro/mesh/fv.py:18-39)"""
    ref_shim.load()
    import torch
    import pyro.mesh.boundary as rbnd
    import pyro.mesh.fv as rfv
    import pyro.mesh.patch as rpatch
    from pyro2_b200.mesh import boundary as bnd
    from pyro2_b200.mesh import fv, patch
    rng = np.random.default_rng(11)
    rg = rpatch.Grid2d(24, 24, ng=4)
    rd = rfv.FV2d(rg)
    rd.register_var("a", rbnd.BC(xlb="periodic", xrb="periodic", ylb="periodic", yrb="periodic"))
    rd.create()
    g = patch.Grid2d(24, 24, ng=4, device="cpu")
    d = fv.FV2d(g)
    d.register_var("a", bnd.BC(xlb="periodic", xrb="periodic", ylb="periodic", yrb="periodic"))
    d.create()
    data = rng.standard_normal((rg.qx, rg.qy))
    rd.get_var("a")[:, :] = data
    d.get_var("a")[:, :] = data
    for pos in (False, True):
        assert np.array_equal(np.asarray(rd.to_centers("a", is_positive=pos)), d.to_centers("a", is_positive=pos).numpy())
    rd.from_centers("a")
    ng = rg.ng
    filled = data.copy()                        # the reference's periodic ghost fill, restated
    filled[:ng] = filled[-2 * ng:-ng]; filled[-ng:] = filled[ng:2 * ng]
    filled[:, :ng] = filled[:, -2 * ng:-ng]; filled[:, -ng:] = filled[:, ng:2 * ng]
    d.get_var("a")[:, :] = filled
    d.fill_BC = lambda name: None               # the product's ghost fill is a device kernel (tested on the GPU)
    d.from_centers("a")
    assert np.array_equal(np.asarray(rd.get_var("a")), d.get_var("a").numpy())


def test_ramp_boundary_bit_identical():
    """oracle.fill_ramp against the reference's own fill_BC with its "ramp" user boundary (compressible/BC.py:183-256)
    on random data, at t = 0 and at a later time"""
    ref_shim.load()
    import pyro.compressible.BC as rBC
    import pyro.mesh.boundary as rbnd
    import pyro.mesh.patch as rpatch
    rbnd.define_bc("ramp", rBC.user, is_solid=False)
    names = ("density", "energy", "x-momentum", "y-momentum")
    rng = np.random.default_rng(8)
    g = rpatch.Cartesian2d(40, 10, ng=4, xmax=4.0, ymax=1.0)
    d = rpatch.CellCenterData2d(g)
    for n in names:
        d.register_var(n, rbnd.BC(xlb="ramp", xrb="outflow", ylb="ramp", yrb="ramp"))
    d.set_aux("gamma", 1.4)
    d.create()
    for t in (0.0, 0.0123):
        d.t = t
        mine = rng.standard_normal((4, g.qx, g.qy))
        for k, n in enumerate(names):
            d.get_var(n)[:, :] = mine[k]
            d.fill_BC(n)
            oracle.fill_ghost(mine[k], g.ng, ("ramp", "outflow", "ramp", "ramp"))
            for side in ("xlb", "ylb", "yrb"):
                oracle.fill_ramp(mine[k], k, side, g.ng, g.x, g.y, g.dx, g.dy, t, 1.4)
            assert np.array_equal(np.asarray(d.get_var(n)), mine[k]), (n, t)


@pytest.mark.parametrize("riemann", ["HLLC", "CGF", "HLLC_lm"])
def test_unphysical_interface_states_follow_the_reference(riemann):
    """unlimited slopes (limiter 0) at a 10x density / pressure jump give interface states with negative density; the
    reference's numba max(smallc, sqrt(negative)) is smallc (Python max semantics), so it carries on with finite
    numbers.  The oracle (and the kernel, tests/test_sweep_emulated.py) must do the same, not propagate NaN."""
    p = ref_shim.make_sim("compressible", "kh", {"mesh.nx": 8, "mesh.ny": 32, "compressible.limiter": 0,
                                                 "compressible.use_flattening": 0, "compressible.cvisc": 0.0,
                                                 "compressible.riemann": riemann, "driver.tmax": 10.0})
    sim = p.sim
    g = sim.cc_data.grid
    rng = np.random.default_rng(135)
    y = np.asarray(g.y)
    dens = np.broadcast_to(1.5 * np.exp(-y / 0.4)[None, :], (g.qx, g.qy)) * (1.0 + 0.1 * rng.standard_normal((g.qx, g.qy)))
    pres = 1.8 * dens * (1.0 + 0.05 * rng.standard_normal((g.qx, g.qy)))
    u, v = 0.3 * rng.standard_normal((g.qx, g.qy)), 0.3 * rng.standard_normal((g.qx, g.qy))
    for name, a in (("density", dens), ("x-momentum", dens * u), ("y-momentum", dens * v),
                    ("energy", pres / 0.4 + 0.5 * dens * (u * u + v * v))):
        sim.cc_data.get_var(name)[:, :] = a
    sim.cc_data.fill_BC_all()                       # periodic in y: the stratification wraps into a strong jump
    U0 = np.asarray(sim.cc_data.data).copy()
    sim.dt = 0.4 * oracle.cfl_dt(U0, g.ng, g.dx, g.dy, 1.4, 0.8)
    import pyro.compressible.unsplit_fluxes as flx
    hat = flx.interface_states(sim.cc_data, sim.rp, sim.ivars, sim.tc, sim.dt)
    assert min(float(np.asarray(h)[..., sim.ivars.idens].min()) for h in hat) < 0.0        # the regime in question
    prm = oracle.comp_params(limiter=0, use_flattening=0, cvisc=0.0, riemann=riemann)
    Unew = oracle.compressible_step(U0, g.ng, g.dx, g.dy, sim.dt, prm)
    sim.evolve()
    v_ = (slice(g.ilo, g.ihi + 1), slice(g.jlo, g.jhi + 1))
    ref = np.asarray(sim.cc_data.data)[v_]
    assert np.isfinite(ref).all() and np.isfinite(Unew[v_]).all()
    assert np.abs(Unew[v_] - ref).max() < 1e-12 * np.abs(ref).max()


@pytest.mark.parametrize("problem,extra", [
    ("sedov", {"mesh.xmin": 0.2, "mesh.xmax": 1.0, "mesh.ymin": 0.785, "mesh.ymax": 2.355, "mesh.xlboundary": "reflect-odd",
               "sedov.r_init": 0.3, "compressible.limiter": 2}),
    ("advect", {"mesh.xmin": 1.0, "mesh.xmax": 2.0, "mesh.ymin": 0.523, "mesh.ymax": 2.617, "mesh.xlboundary": "outflow",
                "compressible.limiter": 0, "compressible.grav": -0.7}),
    # problem heating and the sponge apply in any geometry (simulation.py:148-153, 425-441)
    ("heating", {"mesh.xmin": 0.5, "mesh.xmax": 1.5, "mesh.ymin": 0.6, "mesh.ymax": 2.4, "mesh.xlboundary": "reflect",
                 "compressible.limiter": 2, "compressible.grav": -0.5, "sponge.do_sponge": 1, "sponge.sponge_rho_begin": 1.2,
                 "sponge.sponge_rho_full": 0.8, "sponge.sponge_timescale": 0.01, "heating.e_rate": 2.0})])
def test_spherical_polar_step_matches_reference(problem, extra):
    """SphericalPolar geometry (mesh/patch.py:242-312 and the coord_type == 1 branches of the compressible solver):
    geometry arrays bit for bit, the CFL time step, and whole steps of the live reference against the oracle"""
    p = ref_shim.make_sim("compressible", problem, dict({
        "mesh.grid_type": "SphericalPolar", "mesh.nx": 32, "mesh.ny": 24, "mesh.xrboundary": "outflow",
        "mesh.ylboundary": "outflow", "mesh.yrboundary": "outflow", "compressible.riemann": "CGF", "driver.tmax": 10.0}, **extra))
    sim = p.sim
    g = sim.cc_data.grid
    assert g.coord_type == 1
    geom = oracle.spherical_geometry(g.nx, g.ny, g.ng, g.xmin, g.xmax, g.ymin, g.ymax)
    for k in ("Lx", "Ly", "Ax", "Ay", "V", "dlogAx", "dlogAy"):
        assert np.array_equal(geom[k], np.asarray(getattr(g, k))), k
    bcs = [tuple(getattr(sim.cc_data.BCs[n], s) for s in ("xlb", "xrb", "ylb", "yrb"))
           for n in ("density", "energy", "x-momentum", "y-momentum")]
    heat = {}
    if sim.problem_source is not None:
        # the profile P with S_ener = dens * e_rate * P: the reference's own source_terms() on unit density, e_rate = 1
        rate = sim.rp.get_param(f"{problem}.e_rate")
        sim.rp.set_param(f"{problem}.e_rate", 1.0)
        ones = g.scratch_array(nvar=4)
        ones[:, :, 0] = 1.0
        heat = {"heat_rate": rate, "heat_profile": np.asarray(sim.problem_source(g, ones, sim.ivars, sim.rp))[:, :, sim.ivars.iener].copy()}
        sim.rp.set_param(f"{problem}.e_rate", rate)
    sponge = None
    if sim.rp.get_param("sponge.do_sponge"):
        sponge = tuple(sim.rp.get_param(f"sponge.sponge_{k}") for k in ("rho_begin", "rho_full", "timescale"))
    prm = oracle.comp_params(limiter=extra["compressible.limiter"], riemann="CGF", grav=sim.rp.get_param("compressible.grav"),
                             src_bcs=bcs, geom=geom, xl_solid=int(sim.solid.xl), yl_solid=int(sim.solid.yl), sponge=sponge, **heat)
    v = (slice(g.ilo, g.ihi + 1), slice(g.jlo, g.jhi + 1))
    for _ in range(4):
        sim.cc_data.fill_BC_all()
        U0 = np.asarray(sim.cc_data.data).copy()
        sim.method_compute_timestep()
        assert sim.dt == oracle.cfl_dt_spherical(U0, 1.4, 0.8, geom)
        sim.dt *= 0.5
        Unew = oracle.compressible_step(U0, g.ng, g.dx, g.dy, sim.dt, prm)
        sim.evolve()
        ref = np.asarray(sim.cc_data.data)[v]
        assert np.abs(Unew[v] - ref).max() <= 1e-13 * np.abs(ref).max()

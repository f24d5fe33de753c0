"""The reference's driver loops (pyro_sim.py:241-256: fill_BC_all, compute_timestep, evolve) over the ORACLE, one per
solver -- shared by tests/test_oracle_golden.py (fixtures from running the reference here) and by the pins against the
regression files the reference itself stores (tests/golden/pin_stored_goldens.py).  Test infrastructure."""
import numpy as np

import oracle
from golden_util import var_bcs


def mesh_bc(rp):
    return (rp["mesh.xlboundary"], rp["mesh.xrboundary"], rp["mesh.ylboundary"], rp["mesh.yrboundary"])


def run_compressible(z, rp, nsteps=None, fix_dt=-1.0):
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


def run_incompressible(z, rp):
    """the oracle's evolve (explicit part + two multigrid projections) stepped with the recorded dts; all six planes"""
    ng = int(z["ng"])
    P = np.ascontiguousarray(z["P0"])
    bc = mesh_bc(rp)
    assert bc == ("periodic",) * 4
    for dt in z["dts"]:
        for k in range(6):          # the driver's fill_BC_all before every step (pyro_sim.py:241-256)
            oracle.fill_ghost(P[k], ng, bc)
        oracle.incomp_evolve(P, ng, float(dt), limiter=rp["incompressible.limiter"], proj_type=rp["incompressible.proj_type"],
                             vel_bc=(bc, bc), phi_bc=bc, xmin=rp["mesh.xmin"], xmax=rp["mesh.xmax"],
                             ymin=rp["mesh.ymin"], ymax=rp["mesh.ymax"])
    return P


def run_burgers(z, rp):
    ng = int(z["ng"])
    u, v = z["P0"][0].copy(), z["P0"][1].copy()
    n = rp["mesh.nx"]
    dx = (rp["mesh.xmax"] - rp["mesh.xmin"]) / n
    bc = mesh_bc(rp)
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
    return u, v


def run_advection(z, rp):
    ng, n = int(z["ng"]), rp["mesh.nx"]
    a = z["P0"][0].copy()
    dx = (rp["mesh.xmax"] - rp["mesh.xmin"]) / n
    dy = (rp["mesh.ymax"] - rp["mesh.ymin"]) / rp["mesh.ny"]
    bc = mesh_bc(rp)
    for dt in z["dts"]:
        oracle.fill_ghost(a, ng, bc)
        a = oracle.advection_evolve(a, ng, dx, dy, float(dt), rp["advection.u"], rp["advection.v"], rp["advection.limiter"])
    return a


def run_diffusion(z, rp):
    """one Crank-Nicolson multigrid solve per step"""
    phi = np.ascontiguousarray(z["P0"][0])
    for dt in z["dts"]:
        oracle.diffusion_evolve(phi, float(dt), rp["diffusion.k"], mesh_bc(rp), rp["mesh.xmin"], rp["mesh.xmax"],
                                rp["mesh.ymin"], rp["mesh.ymax"])
    return phi


def lm_setup(z, rp):
    names = [str(n) for n in z["names"]]
    bc = mesh_bc(rp)
    assert bc == ("periodic", "periodic", "reflect", "outflow")      # the setup the fixtures were generated with
    even = ("periodic", "periodic", "reflect-even", "outflow")
    odd_y = ("periodic", "periodic", "reflect-odd", "outflow")
    phi_bc = ("periodic", "periodic", "neumann", "dirichlet")
    fills = dict(zip(names, (even, even, odd_y, even, phi_bc, phi_bc, even, even)))
    prm = oracle.lm_params(rp["mesh.nx"], grav=rp["lm-atmosphere.grav"], gamma=rp["eos.gamma"],
                           limiter=rp["lm-atmosphere.limiter"], proj_type=rp["lm-atmosphere.proj_type"],
                           xmin=rp["mesh.xmin"], xmax=rp["mesh.xmax"], ymin=rp["mesh.ymin"], ymax=rp["mesh.ymax"])
    return names, fills, prm


def run_lm_atm(z, rp):
    """the oracle's evolve -- numba interface routines restated, two variable-coefficient projections; all eight planes"""
    ng = int(z["ng"])
    names, fills, prm = lm_setup(z, rp)
    S = np.ascontiguousarray(z["P0"])
    base = np.ascontiguousarray(z["base"])
    for dt in z["dts"]:
        for k, name in enumerate(names):
            oracle.fill_ghost(S[k], ng, fills[name])
        raw = oracle.lm_timestep(S, base, prm, rp["driver.cfl"])
        assert raw >= float(dt) * (1 - 1e-15)          # the driver only ever shrinks the method's dt
        oracle.lm_evolve(S, base, prm, float(dt))
    return S

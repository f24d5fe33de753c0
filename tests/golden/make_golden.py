"""Generate the golden fixtures in tests/golden/ by RUNNING THE UNMODIFIED REFERENCE
(/root/reference, python-hydro/pyro2) through its own public API.  Run in the build container:

    python tests/golden/make_golden.py

Outputs (small .npz files, committed):
  comp_<problem>.npz   Pyro("compressible") runs: parameters, initial state, per-step dt, final state
  mg_<case>.npz        CellCenterMG2d solves: rhs, solution, cycle count, residual / relative errors
  mgvc_<case>.npz      VarCoeffCCMG2d solves (the reference's mg_test_vc_* setups): coefficients, rhs, solution
  incomp_<problem>.npz Pyro("incompressible") runs: state after initialize + preevolve, per-step dt, final state
  burgers_test.npz     Pyro("burgers") run of its "test" problem
  mesh_bcs.npz         ghost fill of an integer array for every standard BC type (test_patch.py style)
  ref_kats.npz         known answers quoted from the reference's own unit tests / stored outputs
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
import ref_shim  # noqa: E402

ref_shim.load()

# optional filter: `python make_golden.py rt2 burgers_tophat` regenerates only the fixtures whose name contains a word
ONLY = sys.argv[1:]


def _wanted(name):
    return not ONLY or any(w in name for w in ONLY)


def comp_case(name, problem, params, nsteps):
    if not _wanted("comp_" + name):
        return
    p = ref_shim.make_sim("compressible", problem, dict(params, **{"driver.max_steps": 100000}))
    sim = p.sim
    g = sim.cc_data.grid
    U0 = np.asarray(sim.cc_data.data).copy()
    dts = []
    for _ in range(nsteps):
        if sim.finished():
            break
        p.single_step()
        dts.append(sim.dt)
    rp = sim.rp
    keys = ["eos.gamma", "compressible.limiter", "compressible.use_flattening", "compressible.cvisc",
            "compressible.z0", "compressible.z1", "compressible.delta", "driver.cfl", "driver.tmax",
            "driver.init_tstep_factor", "driver.max_dt_change",
            "mesh.xlboundary", "mesh.xrboundary", "mesh.ylboundary", "mesh.yrboundary",
            "mesh.nx", "mesh.ny", "mesh.xmin", "mesh.xmax", "mesh.ymin", "mesh.ymax", "compressible.grav", "compressible.riemann",
            "compressible.small_dens", "sponge.do_sponge", "sponge.sponge_rho_begin", "sponge.sponge_rho_full",
            "sponge.sponge_timescale", "mesh.grid_type"]
    extra = {}
    import importlib
    mod = importlib.import_module(f"pyro.compressible.problems.{problem}")
    if hasattr(mod, "source_terms"):
        # the heating profile P with S_ener = dens * e_rate * P, evaluated exactly as the reference does:
        # its own source_terms() on a state of unit density with e_rate temporarily set to 1
        e_rate = rp.get_param(f"{problem}.e_rate")
        rp.set_param(f"{problem}.e_rate", 1.0)
        ones = g.scratch_array(nvar=4)
        ones[:, :, 0] = 1.0
        extra["heat_profile"] = np.asarray(mod.source_terms(g, ones, sim.ivars, rp))[:, :, sim.ivars.iener].copy()
        extra["heat_rate"] = e_rate
        rp.set_param(f"{problem}.e_rate", e_rate)
    if sim.cc_data.get_aux("ambient_rho") is not None:
        extra["ambient"] = np.array([sim.cc_data.get_aux(k) for k in ("ambient_rho", "ambient_u", "ambient_v", "ambient_p")])
    np.savez_compressed(os.path.join(HERE, f"comp_{name}.npz"), **extra,
                        problem=problem, inputs=np.array([f"{k}={v}" for k, v in params.items()]),
                        rp=np.array([f"{k}={rp.get_param(k)}" for k in keys]),
                        ng=g.ng, U0=U0, U=np.asarray(sim.cc_data.data).copy(), dts=np.array(dts),
                        t=sim.cc_data.t, n=sim.n)
    print(name, "steps", sim.n, "t", sim.cc_data.t)


def mg_case(name, nx, bc, alpha, beta, rhs_kind, rtol, bcfuncs=None):
    if not _wanted("mg_" + name):
        return
    import pyro.multigrid.MG as MG
    kw = dict(xl_BC_type=bc[0], xr_BC_type=bc[1], yl_BC_type=bc[2], yr_BC_type=bc[3], alpha=alpha, beta=beta)
    if bcfuncs:
        kw.update(bcfuncs)
    a = MG.CellCenterMG2d(nx, nx, **kw)
    x, y = a.x2d, a.y2d
    if rhs_kind == "poly":
        f = -2.0 * ((1.0 - 6.0 * x ** 2) * y ** 2 * (1.0 - y ** 2) + (1.0 - 6.0 * y ** 2) * x ** 2 * (1.0 - x ** 2))
    elif rhs_kind == "periodic":
        f = np.sin(2 * np.pi * x) * np.cos(4 * np.pi * y)
    else:
        f = np.exp(-((x - 0.5) ** 2 + (y - 0.5) ** 2) / 0.02)
    a.init_zeros()
    a.init_RHS(f)
    # per-cycle history: re-implement the loop's prints by wrapping v_cycle
    hist = []
    a.solve(rtol=rtol)
    np.savez_compressed(os.path.join(HERE, f"mg_{name}.npz"), nx=nx, bc=np.array(bc), alpha=alpha, beta=beta,
                        rtol=rtol, f=np.asarray(f), v=np.asarray(a.get_solution()),
                        num_cycles=a.num_cycles, residual_error=a.residual_error,
                        relative_error=a.relative_error, source_norm=a.source_norm)
    print(name, "cycles", a.num_cycles, "resid", a.residual_error)


def mgvc_case(name, nx, phibc, cbc, kind, rtol=1.e-11):
    if not _wanted("mgvc_" + name):
        return
    """the setups of pyro/multigrid/examples/mg_test_vc_{dirichlet,periodic,constant}.py"""
    import pyro.mesh.boundary as bnd
    import pyro.multigrid.variable_coeff_MG as VMG
    from pyro.mesh import patch
    g = patch.Grid2d(nx, nx, ng=1)
    d = patch.CellCenterData2d(g)
    bc_c = bnd.BC(xlb=cbc, xrb=cbc, ylb=cbc, yrb=cbc)
    d.register_var("c", bc_c)
    d.create()
    c = d.get_var("c")
    x, y = g.x2d, g.y2d
    pi = np.pi
    if kind == "dirichlet":        # mg_test_vc_dirichlet.py:40-52
        c[:, :] = 2.0 + np.cos(2.0 * pi * x) * np.cos(2.0 * pi * y)
        rhs = -16.0 * pi ** 2 * (np.cos(2 * pi * x) * np.cos(2 * pi * y) + 1) * np.sin(2 * pi * x) * np.sin(2 * pi * y)
    elif kind == "periodic":       # mg_test_vc_periodic.py:41-54
        c[:, :] = 2.0 + np.cos(2.0 * pi * x) * np.cos(2.0 * pi * y)
        rhs = -16.0 * pi ** 2 * (np.cos(2 * pi * x) * np.cos(2 * pi * y) + 1) * np.sin(2 * pi * x) * np.sin(2 * pi * y)
    else:                          # mg_test_vc_constant.py:33-45: alpha = 1, the constant-coefficient problem
        c[:, :] = 1.0
        rhs = -2.0 * ((1.0 - 6.0 * x ** 2) * y ** 2 * (1.0 - y ** 2) + (1.0 - 6.0 * y ** 2) * x ** 2 * (1.0 - x ** 2))
    a = VMG.VarCoeffCCMG2d(nx, nx, xl_BC_type=phibc, yl_BC_type=phibc, xr_BC_type=phibc, yr_BC_type=phibc,
                           coeffs=c, coeffs_bc=bc_c, verbose=0)
    a.init_zeros()
    a.init_RHS(rhs)
    a.solve(rtol=rtol)
    r = a.grids[a.nlevels - 1].get_var("r")
    np.savez_compressed(os.path.join(HERE, f"mgvc_{name}.npz"), nx=nx, bc=np.array((phibc,) * 4),
                        coeffs_bc=np.array((cbc,) * 4), rtol=rtol, coeffs=np.asarray(c), f=np.asarray(rhs),
                        v=np.asarray(a.get_solution()), r=np.asarray(r), num_cycles=a.num_cycles,
                        residual_error=a.residual_error, relative_error=a.relative_error,
                        source_norm=a.source_norm,
                        ex_coarse=np.asarray(a.edge_coeffs[2].x), ey_coarse=np.asarray(a.edge_coeffs[2].y))
    print(name, "cycles", a.num_cycles, "resid", a.residual_error)


INCOMP_VARS = ["x-velocity", "y-velocity", "phi-MAC", "phi", "gradp_x", "gradp_y"]


def flow_case(fname, solver, problem, params, nsteps, names):
    """incompressible / burgers: planes [var, i, j] after initialize_problem (which includes the
    incompressible solver's preevolve), the dt of every step, the planes after nsteps"""
    if not _wanted(fname):
        return
    import importlib
    mod = importlib.import_module(f"pyro.{solver}.problems.{problem}")
    if not hasattr(mod, "PROBLEM_PARAMS"):
        # burgers converge / tophat omit the (empty) parameter table Pyro.initialize_problem reads; supply it here,
        # in memory, so that the stock setups can be run at all -- the physics is untouched
        mod.PROBLEM_PARAMS = {}
    p = ref_shim.make_sim(solver, problem, dict(params, **{"driver.max_steps": 100000}))
    sim = p.sim
    g = sim.cc_data.grid
    P0 = np.stack([np.asarray(sim.cc_data.get_var(n)) for n in names]).copy()
    dts = []
    for _ in range(nsteps):
        if sim.finished():
            break
        p.single_step()
        dts.append(sim.dt)
    rp = sim.rp
    keys = ["driver.cfl", "driver.tmax", "driver.init_tstep_factor", "driver.max_dt_change", "driver.fix_dt",
            "mesh.xlboundary", "mesh.xrboundary", "mesh.ylboundary", "mesh.yrboundary",
            "mesh.nx", "mesh.ny", "mesh.xmin", "mesh.xmax", "mesh.ymin", "mesh.ymax"]
    keys += {"incompressible": ["incompressible.limiter", "incompressible.proj_type"], "burgers": ["advection.limiter"],
             "advection": ["advection.limiter", "advection.u", "advection.v"], "diffusion": ["diffusion.k"],
             "lm_atm": ["lm-atmosphere.limiter", "lm-atmosphere.proj_type", "lm-atmosphere.grav", "eos.gamma"]}[solver]
    extra = {}
    if solver == "lm_atm":      # the 1-d base state the problem setup builds (rho0, p0) and the derived beta0 arrays
        extra["base"] = np.stack([sim.base[k].d for k in ("rho0", "p0", "beta0", "beta0-edges")])
    np.savez_compressed(os.path.join(HERE, fname), problem=problem,
                        inputs=np.array([f"{k}={v}" for k, v in params.items()]),
                        rp=np.array([f"{k}={rp.get_param(k)}" for k in keys]), names=np.array(names),
                        ng=g.ng, P0=P0, P=np.stack([np.asarray(sim.cc_data.get_var(n)) for n in names]),
                        dts=np.array(dts), t=sim.cc_data.t, n=sim.n, **extra)
    print(fname, "steps", sim.n, "t", sim.cc_data.t)


def mesh_bcs():
    if not _wanted("mesh_bcs"):
        return
    from pyro.mesh import boundary as bnd
    from pyro.mesh import patch
    out = {}
    rng = np.random.default_rng(7)
    for ng in (1, 4):
        g = patch.Grid2d(6, 5, ng=ng)
        base = rng.integers(-50, 50, size=(g.qx, g.qy))
        out[f"base_ng{ng}"] = base
        for t in ("outflow", "periodic", "reflect-even", "reflect-odd"):
            d = patch.CellCenterData2d(g, dtype=np.int_)
            d.register_var("a", bnd.BC(xlb=t, xrb=t, ylb=t, yrb=t))
            d.create()
            d.get_var("a")[:, :] = base
            d.fill_BC("a")
            out[f"{t}_ng{ng}"] = np.asarray(d.get_var("a")).copy()
    np.savez_compressed(os.path.join(HERE, "mesh_bcs.npz"), **out)
    print("mesh_bcs done")


def ref_kats():
    if not _wanted("ref_kats"):
        return
    """constants the reference's own tests assert (cited so the oracle is pinned to them too)"""
    conv = np.loadtxt(os.path.join(ref_shim.REF_ROOT, "pyro/multigrid/tests/mg_convergence.txt"))
    np.savez_compressed(os.path.join(HERE, "ref_kats.npz"),
                        mg_convergence=conv,                       # pyro/multigrid/tests/mg_convergence.txt
                        mg_gradient_row=np.array([0, 36, 60, 36, 12, -12, -36, -60, -36, 0.]),  # test_multigrid_comps.py:54-57
                        indexer_v=np.array([[16., 17., 18.], [23., 24., 25.]]),                  # test_array_indexer.py:22
                        indexer_ip1=np.array([[23., 24., 25.], [30., 31., 32.]]),
                        indexer_jpm1=np.array([[15., 16., 17.], [22., 23., 24.]]),
                        advection_smooth_sum=4.310466040637315e+03)
    print("kats done", conv.shape)


if __name__ == "__main__":
    comp_case("sedov64", "sedov", {"mesh.nx": 64, "mesh.ny": 64, "sedov.r_init": 0.05, "driver.tmax": 0.05}, 40)
    comp_case("quad64", "quad", {"mesh.nx": 64, "mesh.ny": 64, "driver.tmax": 0.3}, 40)
    comp_case("sod_x", "sod", {}, 1000)          # the reference's own regression setup: 128x10, limiter 1, 76 steps
    comp_case("kh32", "kh", {"mesh.nx": 32, "mesh.ny": 32, "driver.tmax": 0.2}, 25)
    comp_case("acoustic64", "acoustic_pulse", {"mesh.nx": 64, "mesh.ny": 64, "driver.fix_dt": 3.0e-3}, 20)
    comp_case("advect32", "advect", {"mesh.nx": 32, "mesh.ny": 32, "driver.fix_dt": 0.01}, 20)   # limiter 0
    comp_case("gresho40", "gresho", {}, 15)
    # the CGF Riemann solver (riemann_cgf + consFlux), with and without solid walls
    comp_case("sedov32_cgf", "sedov", {"mesh.nx": 32, "mesh.ny": 32, "sedov.r_init": 0.1, "compressible.riemann": "CGF"}, 30)
    comp_case("quad32_cgf_walls", "quad", {"mesh.nx": 32, "mesh.ny": 32, "compressible.riemann": "CGF",
                                           "mesh.xlboundary": "reflect", "mesh.xrboundary": "reflect",
                                           "mesh.ylboundary": "reflect", "mesh.yrboundary": "outflow"}, 30)
    # problem heating sources, the sponge, the "ambient" boundary, the density floor
    comp_case("heating32", "heating", {"mesh.nx": 32, "mesh.ny": 32}, 25)
    comp_case("plume32", "plume", {"mesh.nx": 32, "mesh.ny": 64, "mesh.ymax": 4.0}, 25)
    comp_case("convection16", "convection", {"mesh.nx": 16, "mesh.ny": 96}, 25)
    # gravity + the compressible solver's "hse" boundary (compressible/BC.py)
    comp_case("bubble32", "bubble", {"mesh.nx": 32, "mesh.ny": 64, "mesh.ymax": 4.0}, 25)
    comp_case("rt16", "rt", {"mesh.nx": 16, "mesh.ny": 48}, 25)
    comp_case("hse16", "hse", {"mesh.nx": 16, "mesh.ny": 48}, 20)
    comp_case("gresho40_lm", "gresho", {"compressible.riemann": "HLLC_lm"}, 15)
    comp_case("sedov32_lm", "sedov", {"mesh.nx": 32, "mesh.ny": 32, "sedov.r_init": 0.1, "compressible.riemann": "HLLC_lm"}, 30)
    SPH = {"mesh.grid_type": "SphericalPolar", "mesh.nx": 32, "mesh.ny": 32, "compressible.riemann": "CGF",
           "mesh.xrboundary": "outflow", "mesh.ylboundary": "outflow", "mesh.yrboundary": "outflow"}
    # the reference's inputs.sedov.spherical / inputs.advect.spherical.64 at 32 x 32 (xmin raised so that the coarser
    # grid's ghost cells keep r > 0).  The stock sedov inputs put "reflect-odd" on the inner boundary, which reflects
    # EVERY variable oddly: the ghost rows then hold negative densities, the CGF star states on that side have
    # c = smallc = 1e-10 and round-off is amplified until even two builds of the same arithmetic part ways within ten
    # steps -- not a usable fixture.  "reflect" (a solid inner wall) is the physical version of that setup.
    comp_case("sedov_sph32", "sedov", dict(SPH, **{"mesh.xmin": 0.2, "mesh.xmax": 1.0, "mesh.ymin": 0.785, "mesh.ymax": 2.355,
                                                    "mesh.xlboundary": "reflect", "sedov.r_init": 0.3, "driver.tmax": 0.1}), 25)
    comp_case("advect_sph32", "advect", dict(SPH, **{"mesh.xmin": 1.0, "mesh.xmax": 2.0, "mesh.ymin": 0.523, "mesh.ymax": 2.617,
                                                      "mesh.xlboundary": "outflow", "compressible.limiter": 0,
                                                      "driver.fix_dt": 0.005, "driver.init_tstep_factor": 1.0}), 20)
    comp_case("ramp64", "ramp", {"mesh.nx": 64, "mesh.ny": 16}, 30)
    comp_case("rt2_48", "rt2", {"mesh.nx": 48, "mesh.ny": 48, "rt2.sigma": 0.1}, 25)
    comp_case("rt_multimode16", "rt_multimode", {"mesh.nx": 16, "mesh.ny": 48}, 25)
    comp_case("rt16_reflect", "rt", {"mesh.nx": 16, "mesh.ny": 48, "mesh.xlboundary": "reflect", "mesh.xrboundary": "outflow",
                                     "mesh.ylboundary": "reflect", "mesh.yrboundary": "reflect"}, 20)
    mg_case("poisson_dirichlet_64", 64, ("dirichlet",) * 4, 0.0, -1.0, "poly", 1.e-11)
    mg_case("poisson_dirichlet_256", 256, ("dirichlet",) * 4, 0.0, -1.0, "poly", 1.e-11)
    mg_case("poisson_periodic_64", 64, ("periodic",) * 4, 0.0, -1.0, "periodic", 1.e-11)
    mg_case("helmholtz_neumann_64", 64, ("neumann",) * 4, 1.0, 0.01, "gauss", 1.e-12)
    mg_case("poisson_mixed_128", 128, ("dirichlet", "dirichlet", "neumann", "neumann"), 0.0, -1.0, "poly", 1.e-11)
    mg_case("poisson_inhom_64", 64, ("dirichlet",) * 4, 0.0, -1.0, "poly", 1.e-11,
            bcfuncs=dict(xl_BC=lambda y: y ** 2, xr_BC=lambda y: 1.0 + y, yl_BC=lambda x: x, yr_BC=lambda x: 1.0 + x ** 2))
    mgvc_case("dirichlet_64", 64, "dirichlet", "neumann", "dirichlet")
    mgvc_case("periodic_64", 64, "periodic", "periodic", "periodic")
    mgvc_case("constant_32", 32, "dirichlet", "neumann", "constant")
    mgvc_case("dirichlet_128", 128, "dirichlet", "neumann", "dirichlet")
    flow_case("incomp_shear32.npz", "incompressible", "shear", {"mesh.nx": 32, "mesh.ny": 32}, 12, INCOMP_VARS)
    flow_case("incomp_shear64.npz", "incompressible", "shear", {"mesh.nx": 64, "mesh.ny": 64}, 6, INCOMP_VARS)
    flow_case("incomp_converge32.npz", "incompressible", "converge",
              {"mesh.nx": 32, "mesh.ny": 32, "driver.cfl": 0.5, "driver.fix_dt": 5.e-3, "driver.init_tstep_factor": 1.0}, 10,
              INCOMP_VARS)
    flow_case("burgers_test.npz", "burgers", "test", {"mesh.nx": 64, "mesh.ny": 64}, 12, ["x-velocity", "y-velocity"])
    # stock inputs.converge.64 / inputs.tophat at 32 x 32 (particles are not part of the build)
    flow_case("burgers_converge32.npz", "burgers", "converge", {"mesh.nx": 32, "mesh.ny": 32, "particles.do_particles": 0}, 15,
              ["x-velocity", "y-velocity"])
    flow_case("burgers_tophat32.npz", "burgers", "tophat", {}, 15, ["x-velocity", "y-velocity"])
    # BASELINE config 1: advection smooth 64 x 64, 81 steps to t = 1 (sum(density) = 4.310466040637315e+03)
    flow_case("advection_smooth64.npz", "advection", "smooth", {"mesh.nx": 64, "mesh.ny": 64, "particles.do_particles": 0}, 1000, ["density"])
    flow_case("advection_tophat32.npz", "advection", "tophat", {"advection.u": -0.6, "advection.v": 1.0, "advection.limiter": 1}, 30, ["density"])
    flow_case("diffusion_gaussian64.npz", "diffusion", "gaussian", {"mesh.nx": 64, "mesh.ny": 64}, 12, ["phi"])
    flow_case("diffusion_gaussian32_mixed.npz", "diffusion", "gaussian",
              {"mesh.nx": 32, "mesh.ny": 32, "mesh.xlboundary": "periodic", "mesh.xrboundary": "periodic",
               "mesh.ylboundary": "dirichlet", "mesh.yrboundary": "neumann", "driver.cfl": 0.7, "driver.tmax": 1.0}, 10, ["phi"])
    LM_VARS = ["density", "x-velocity", "y-velocity", "eint", "phi-MAC", "phi", "gradp_x", "gradp_y"]
    flow_case("lm_bubble32.npz", "lm_atm", "bubble", {"mesh.nx": 32, "mesh.ny": 32}, 10, LM_VARS)
    flow_case("lm_bubble64_lim1.npz", "lm_atm", "bubble",
              {"mesh.nx": 64, "mesh.ny": 64, "lm-atmosphere.limiter": 1, "lm-atmosphere.proj_type": 1}, 6, LM_VARS)
    mesh_bcs()
    ref_kats()

"""Randomised end-to-end comparison: Pyro("compressible") runs on the emulated device (the product's Python layer over
the host-compiled kernels, tests/emu_device.py) against the oracle's driver loop (tests/oracle_runs.run_compressible)
started from the same initial state -- random problems, boundary overrides, Riemann solvers, limiters, gravity.  This
exercises the host-side plumbing (parameter structs, boundary hooks, source flags, time-step control).
Development tool (CPU only):

    python scripts/fuzz_pyro_compressible_emulated.py [ncases] [seed]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), ROOT]
import emu_device  # noqa: E402
import oracle  # noqa: E402
from conftest import state_errors  # noqa: E402
from oracle_runs import run_compressible as _run_oracle  # noqa: E402

PROBLEMS = {  # problem -> (base inputs, allowed y boundaries, allowed x boundaries)
    "sedov": ({"mesh.nx": 24, "mesh.ny": 20, "sedov.r_init": 0.12}, ["outflow", "reflect", "periodic"], ["outflow", "reflect", "periodic"]),
    "quad": ({"mesh.nx": 20, "mesh.ny": 24}, ["outflow", "reflect"], ["outflow", "reflect"]),
    "kh": ({"mesh.nx": 16, "mesh.ny": 32}, ["periodic", "reflect"], ["periodic"]),
    "bubble": ({"mesh.nx": 16, "mesh.ny": 32, "mesh.ymax": 4.0}, ["hse", "reflect"], ["outflow", "periodic", "reflect"]),
    "rt": ({"mesh.nx": 12, "mesh.ny": 36}, ["hse", "reflect"], ["periodic", "reflect"]),
    "hse": ({"mesh.nx": 10, "mesh.ny": 40}, ["hse", "reflect"], ["periodic"]),
    "heating": ({"mesh.nx": 20, "mesh.ny": 20}, ["outflow"], ["outflow", "periodic"]),
    "plume": ({"mesh.nx": 16, "mesh.ny": 32, "mesh.ymax": 4.0}, ["hse"], ["outflow", "reflect"]),
    "convection": ({"mesh.nx": 12, "mesh.ny": 72}, ["reflect+ambient"], ["periodic"]),
    # SphericalPolar grids (CGF only); sizes chosen per case below
    "sedov:sph": ({"mesh.grid_type": "SphericalPolar", "mesh.xmin": 0.4, "mesh.xmax": 1.2, "mesh.ymin": 0.785, "mesh.ymax": 2.355,
                   "sedov.r_init": 0.6}, ["outflow", "reflect"], ["reflect", "outflow"]),
    "advect:sph": ({"mesh.grid_type": "SphericalPolar", "mesh.xmin": 1.0, "mesh.xmax": 2.0, "mesh.ymin": 0.523, "mesh.ymax": 2.617,
                    "driver.fix_dt": -1.0}, ["outflow", "reflect"], ["outflow", "periodic"]),
}
KEYS = ["eos.gamma", "compressible.limiter", "compressible.use_flattening", "compressible.cvisc", "compressible.z0",
        "compressible.z1", "compressible.delta", "driver.cfl", "driver.tmax", "driver.init_tstep_factor",
        "driver.max_dt_change", "mesh.xlboundary", "mesh.xrboundary", "mesh.ylboundary", "mesh.yrboundary", "mesh.nx",
        "mesh.ny", "mesh.xmin", "mesh.xmax", "mesh.ymin", "mesh.ymax", "compressible.grav", "compressible.riemann",
        "compressible.small_dens", "sponge.do_sponge", "sponge.sponge_rho_begin", "sponge.sponge_rho_full",
        "sponge.sponge_timescale", "mesh.grid_type"]


def one_case(rng):
    from pyro2_b200.pyro_sim import Pyro
    problem = str(rng.choice(list(PROBLEMS)))
    base, ybcs, xbcs = PROBLEMS[problem]
    inputs = dict(base)
    spherical = problem.endswith(":sph")
    problem = problem.split(":")[0]
    if spherical:
        inputs.update({"mesh.nx": int(rng.choice([16, 20, 24])), "mesh.ny": int(rng.choice([16, 24, 31, 40]))})
    yb, xb = str(rng.choice(ybcs)), str(rng.choice(xbcs))
    if yb == "reflect+ambient":
        inputs.update({"mesh.ylboundary": "reflect", "mesh.yrboundary": "ambient"})
    else:
        inputs.update({"mesh.ylboundary": yb, "mesh.yrboundary": yb if yb == "periodic" or rng.integers(2) else "outflow"})
        if problem in ("bubble", "rt", "hse", "plume") and inputs["mesh.yrboundary"] == "outflow" and yb == "hse":
            inputs["mesh.yrboundary"] = "hse"
    inputs.update({"mesh.xlboundary": xb, "mesh.xrboundary": xb if xb == "periodic" or rng.integers(2) else "outflow"})
    inputs["compressible.riemann"] = "CGF" if spherical else str(rng.choice(["HLLC", "CGF", "HLLC_lm"]))
    if spherical:
        inputs["compressible.grav"] = float(rng.choice([0.0, -0.5]))
    inputs["compressible.limiter"] = int(rng.choice([1, 2, 2]))
    inputs["compressible.use_flattening"] = int(rng.integers(2))
    inputs["compressible.cvisc"] = float(rng.choice([0.1, 0.0]))
    inputs["driver.cfl"] = float(rng.choice([0.8, 0.5]))
    inputs.update({"driver.max_steps": 10 ** 6, "driver.tmax": 1e9, "driver.verbose": 0})
    nsteps = 3
    with emu_device.emulated_device():
        p = Pyro("compressible")
        p.initialize_problem(problem, inputs_dict=inputs)
        sim = p.sim
        g = sim.cc_data.grid
        U0 = sim.cc_data.data.numpy().copy()
        rp = {k: sim.rp.get_param(k) for k in KEYS}
        z = {"ng": g.ng, "U0": U0, "dts": np.zeros(nsteps)}
        if sim._heat_plane is not None:
            import importlib
            mod = importlib.import_module(f"pyro2_b200.compressible.problems.{problem}")
            rate, prof = mod.heating(g, sim.rp)
            z["heat_rate"], z["heat_profile"] = rate, np.ascontiguousarray(prof)
        if sim.cc_data.get_aux("ambient_rho") is not None:
            z["ambient"] = np.array([sim.cc_data.get_aux(k) for k in ("ambient_rho", "ambient_u", "ambient_v", "ambient_p")])
        dts = []
        for _ in range(nsteps):
            p.single_step()
            dts.append(sim.dt)
        sim.check_state()
        got = sim.cc_data.data.numpy().copy()
    ref, rdts, ng = _run_oracle(z, rp, nsteps=nsteps)
    v = (slice(ng, -ng), slice(ng, -ng))
    err = max(state_errors(got[v], ref[v], rp["eos.gamma"]))
    ok = np.allclose(dts, rdts, rtol=1e-12, atol=0) and err < 1e-11
    bcs = tuple(inputs[f"mesh.{s}boundary"] for s in ("xl", "xr", "yl", "yr"))
    return ok, dict(problem=problem, bc=bcs, riemann=inputs["compressible.riemann"], limiter=inputs["compressible.limiter"],
                    flat=inputs["compressible.use_flattening"], cvisc=inputs["compressible.cvisc"], err=float(f"{err:.2e}"),
                    dts=[float(f"{a / b - 1:.1e}") for a, b in zip(dts, rdts)])


if __name__ == "__main__":
    ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    bad = 0
    for c in range(ncases):
        try:
            ok, desc = one_case(rng)
        except (SystemExit, AssertionError, ValueError) as e:
            ok, desc = False, {"exception": repr(e)}
        if not ok:
            bad += 1
            print("FAIL", c, desc, flush=True)
    print(f"{ncases} cases, {bad} failed")
    sys.exit(1 if bad else 0)

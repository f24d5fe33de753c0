"""Randomised end-to-end comparison of Pyro("advection" / "burgers" / "diffusion") runs on the emulated device with the
oracle's step functions started from the same initial state: random grid shapes, boundary types, limiters, CFL numbers.
Development tool (CPU only):

    python scripts/fuzz_pyro_flow_emulated.py [ncases] [seed]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), ROOT]
import emu_device  # noqa: E402
import oracle  # noqa: E402


def pair(rng, kinds):
    a = str(rng.choice(kinds + ["periodic"]))
    return (a, a) if a == "periodic" else (a, str(rng.choice(kinds)))


def one_case(rng):
    from pyro2_b200.pyro_sim import Pyro
    solver = str(rng.choice(["advection", "burgers", "diffusion"]))
    inputs = {"driver.max_steps": 10 ** 6, "driver.tmax": 1e9, "driver.verbose": 0, "driver.cfl": float(rng.choice([0.8, 0.4]))}
    nsteps = 4
    if solver == "diffusion":
        n = int(rng.choice([16, 32, 64]))
        problem = "gaussian"
        bc = pair(rng, ["neumann", "dirichlet"]) + pair(rng, ["neumann", "dirichlet"])
        inputs.update({"mesh.nx": n, "mesh.ny": n, "diffusion.k": float(rng.choice([1.0, 0.3])), "driver.cfl": float(rng.choice([0.8, 2.0]))})
    else:
        nx, ny = int(rng.integers(8, 40)), int(rng.integers(8, 40))
        problem = str(rng.choice(["smooth", "tophat"] if solver == "advection" else ["test", "tophat", "converge"]))
        bc = pair(rng, ["outflow", "reflect-even"]) + pair(rng, ["outflow", "reflect-even"])
        inputs.update({"mesh.nx": nx, "mesh.ny": ny, "advection.limiter": int(rng.integers(3))})
        if solver == "advection":
            inputs.update({"advection.u": float(rng.standard_normal()), "advection.v": float(rng.standard_normal())})
        inputs["driver.fix_dt"] = -1.0
    inputs.update(dict(zip(("mesh.xlboundary", "mesh.xrboundary", "mesh.ylboundary", "mesh.yrboundary"), bc)))
    with emu_device.emulated_device():
        p = Pyro(solver)
        p.initialize_problem(problem, inputs_dict=inputs)
        sim = p.sim
        g = sim.cc_data.grid
        P0 = sim.cc_data.planes[:, :, :g.qy].numpy().copy()
        dts = []
        for _ in range(nsteps):
            p.single_step()
            dts.append(sim.dt)
        got = sim.cc_data.planes[:, :, :g.qy].numpy().copy()
    ng = g.ng
    v = (slice(ng, -ng), slice(ng, -ng))
    if solver == "diffusion":
        phi = np.ascontiguousarray(P0[0])
        for dt in dts:
            oracle.diffusion_evolve(phi, dt, inputs["diffusion.k"], bc)
        ok = np.array_equal(phi[v], got[0][v]) and all(dt == inputs["driver.cfl"] * min(g.dx, g.dy) ** 2 / inputs["diffusion.k"] for dt in dts)
    elif solver == "advection":
        a = P0[0].copy()
        for dt in dts:
            oracle.fill_ghost(a, ng, bc)
            a = oracle.advection_evolve(a, ng, g.dx, g.dy, dt, inputs["advection.u"], inputs["advection.v"], inputs["advection.limiter"])
        ok = np.array_equal(a[v], got[0][v])
    else:
        u, w = P0[0].copy(), P0[1].copy()
        ok = True
        for step, dt in enumerate(dts):
            oracle.fill_ghost(u, ng, bc)
            oracle.fill_ghost(w, ng, bc)
            raw = inputs["driver.cfl"] * min(g.dx / max(np.abs(u).max(), 1e-12), g.dy / max(np.abs(w).max(), 1e-12))
            # NullSimulation.compute_timestep: first-step factor, then growth limited to max_dt_change per step
            want = raw * sim.rp.get_param("driver.init_tstep_factor") if step == 0 else min(raw, sim.rp.get_param("driver.max_dt_change") * dts[step - 1])
            ok &= dt == want
            u, w = oracle.burgers_evolve(u, w, ng, g.dx, g.dy, dt, inputs["advection.limiter"])
        ok = bool(ok) and np.array_equal(u[v], got[0][v]) and np.array_equal(w[v], got[1][v])
    return ok, dict(solver=solver, problem=problem, bc=bc, nx=inputs["mesh.nx"], ny=inputs["mesh.ny"], dts=dts)


if __name__ == "__main__":
    ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    bad = 0
    for c in range(ncases):
        try:
            ok, desc = one_case(rng)
        except (SystemExit, AssertionError, ValueError, KeyError) as e:
            ok, desc = False, {"exception": repr(e)}
        if not ok:
            bad += 1
            print("FAIL", c, desc, flush=True)
    print(f"{ncases} cases, {bad} failed")
    sys.exit(1 if bad else 0)

"""Randomised check of decomposed Pyro("compressible") runs (x-slabs, halo exchange, all-reduced dt) against the
single-domain run on the emulated device over gloo: 2-4 ranks, random problems, x / y boundary types, Riemann solvers,
gravity.  Development tool (CPU only):

    python scripts/fuzz_compressible_slabs_gloo.py [ncases] [seed]
"""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def worker(rank, size, port, problem, inputs, nsteps, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), ROOT]
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        import emu_device
        from pyro2_b200.parallel import SlabDecomposition
        from pyro2_b200.pyro_sim import Pyro
        with emu_device.emulated_device():
            def run(**kw):
                p = Pyro("compressible")
                p.initialize_problem(problem, inputs_dict=inputs, **kw)
                dts = []
                for _ in range(nsteps):
                    p.single_step()
                    dts.append(p.sim.dt)
                p.sim.check_state()
                g = p.sim.cc_data.grid
                return p.sim.cc_data.planes[:, g.ilo:g.ihi + 1, g.jlo:g.jhi + 1].contiguous(), dts
            mine, dts = run(decomposition=SlabDecomposition())
            parts = [torch.empty_like(mine) for _ in range(size)] if rank == 0 else None
            dist.gather(mine, parts, dst=0)
            res = None
            if rank == 0:
                full = torch.cat(parts, dim=1).numpy()
                one, dts1 = run()
                one = one.numpy()
                res = (bool(np.array_equal(full, one)) and dts == dts1, float(np.abs(full - one).max()), dts == dts1)
        dist.barrier()
        if rank == 0:
            q.put(res)
    finally:
        dist.destroy_process_group()


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


PROBLEMS = {"sedov": ({"sedov.r_init": 0.15}, ["outflow", "reflect", "periodic"], ["outflow", "reflect", "periodic"]),
            "quad": ({}, ["outflow", "reflect"], ["outflow", "reflect"]),
            "kh": ({}, ["periodic"], ["periodic", "reflect"]),
            "rt": ({"mesh.ymax": 3.0}, ["periodic", "reflect"], ["reflect"]),
            "advect": ({}, ["periodic", "outflow"], ["periodic", "outflow"]),
            # gravity problems with the user-defined y boundaries (hooks run per slab after the halo exchange)
            "bubble": ({"mesh.ymax": 4.0}, ["outflow", "periodic", "reflect"], ["hse"]),
            "hse": ({}, ["periodic"], ["hse"]),
            "plume": ({"mesh.ymax": 4.0}, ["outflow", "reflect"], ["hse"]),
            "convection": ({}, ["periodic"], ["reflect+ambient"]),
            # the remaining problem setups on their stock boundaries (None: keep the problem's own) -- the gresho setup
            # centred its vortex on the SLAB until this list grew (gresho.py took the midpoint of the local x array)
            "gresho": ({}, ["periodic"], ["periodic"]),
            "acoustic_pulse": ({}, ["periodic", "outflow"], ["periodic", "outflow"]),
            "sod": ({}, ["outflow"], ["reflect", "outflow"]),
            "rt2": ({}, ["periodic"], None),
            "rt_multimode": ({}, ["periodic"], None),
            "heating": ({}, None, None),
            "ramp": ({"mesh.ny": 16, "nx_per_ny": 4}, None, None),
            # SphericalPolar grids (slabs along r; CGF only, no user boundaries)
            "advect@sph": ({"mesh.grid_type": "SphericalPolar", "mesh.xmin": 1.0, "mesh.xmax": 2.0, "mesh.ymin": 0.523, "mesh.ymax": 2.617,
                            "driver.fix_dt": -1.0}, ["reflect", "outflow"], ["outflow"]),
            "sedov@sph": ({"mesh.grid_type": "SphericalPolar", "mesh.xmin": 0.2, "mesh.xmax": 1.0, "mesh.ymin": 0.785, "mesh.ymax": 2.355,
                           "sedov.r_init": 0.3}, ["reflect", "outflow"], ["outflow", "reflect"])}

if __name__ == "__main__":
    ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    ctx = mp.get_context("spawn")
    bad = 0
    seen = {}
    for c in range(ncases):
        size = int(rng.choice([2, 3, 4]))
        problem = str(rng.choice(list(PROBLEMS)))
        base, xbcs, ybcs = PROBLEMS[problem]
        xb = str(rng.choice(xbcs)) if xbcs else None
        yb = str(rng.choice(ybcs)) if ybcs else None
        seen[problem] = seen.get(problem, 0) + 1
        inputs = dict(base)
        inputs.update({"mesh.nx": size * int(rng.integers(4, 9)), "mesh.ny": int(rng.choice([12, 31, 36]))})
        if xb:
            inputs.update({"mesh.xlboundary": xb, "mesh.xrboundary": xb if xb == "periodic" or rng.integers(2) else "outflow"})
        if yb:
            inputs.update({"mesh.ylboundary": yb, "mesh.yrboundary": yb if yb in ("periodic", "hse") or rng.integers(2) else "outflow"})
        if "nx_per_ny" in inputs:                       # the double Mach reflection's 4 : 1 domain
            inputs["mesh.ny"] = base["mesh.ny"]
            inputs["mesh.nx"] = size * (-(-inputs.pop("nx_per_ny") * inputs["mesh.ny"] // size))
        inputs.update({"compressible.riemann": str(rng.choice(["HLLC", "CGF", "HLLC_lm"])),
                       "compressible.limiter": int(rng.integers(3)), "compressible.cvisc": float(rng.choice([0.1, 0.0])),
                       "driver.max_steps": 10 ** 6, "driver.tmax": 1e9, "driver.verbose": 0})
        if yb == "reflect+ambient":
            inputs.update({"mesh.ylboundary": "reflect", "mesh.yrboundary": "ambient", "mesh.ny": 72})
        if yb == "hse":
            inputs["mesh.ny"] = 36
        if problem == "rt":
            inputs["compressible.grav"] = -1.0
        if problem.endswith("@sph"):
            problem = problem[:-4]
            # the ghost rows must keep r > 0 (the grid refuses otherwise, like the reference's): xmin > ng dx
            need = int(4 * (inputs["mesh.xmax"] - inputs["mesh.xmin"]) / inputs["mesh.xmin"]) + 1
            inputs["mesh.nx"] = size * max(inputs["mesh.nx"] // size, -(-need // size))
            inputs["compressible.riemann"] = "CGF"
            inputs["compressible.grav"] = float(rng.choice([0.0, -0.5]))
        if problem in ("sedov", "quad") and inputs["compressible.limiter"] == 0:
            inputs["compressible.limiter"] = 1          # unlimited slopes at the blast / the contacts go negative
                                                        # (in the single-domain run and in the reference too)
        q = ctx.Queue()
        port = free_port()
        procs = [ctx.Process(target=worker, args=(r, size, port, problem, inputs, 3, q)) for r in range(size)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(900)
        ok = all(p.exitcode == 0 for p in procs)
        res = q.get(timeout=5) if ok else None
        desc = dict(size=size, problem=problem, nx=inputs["mesh.nx"], ny=inputs["mesh.ny"],
                    bc=tuple(inputs.get(f"mesh.{s}boundary", "stock") for s in ("xl", "xr", "yl", "yr")), riemann=inputs["compressible.riemann"],
                    limiter=inputs["compressible.limiter"], res=res)
        if not ok or not res[0]:
            bad += 1
            print("FAIL", c, desc, flush=True)
        else:
            print("ok  ", c, desc, flush=True)
    print(f"{ncases} cases, {bad} failed;  problems: " + ", ".join(f"{k} {v}" for k, v in sorted(seen.items())))
    sys.exit(1 if bad else 0)

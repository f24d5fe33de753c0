"""Randomised check of decomposed Pyro("advection" / "burgers" / "diffusion" / "incompressible" / "lm_atm") runs (x-slabs,
halo rows, all-reduced dt, the projections / the Crank-Nicolson solve on the x-slab multigrid) against the single-domain run
on the emulated device over gloo: 2-4 ranks, random problems, grid shapes, limiters, boundary types, multigrid split
levels.  Every state plane and every dt must agree bit for bit.  Development tool (CPU only):

    python scripts/fuzz_flow_slabs_gloo.py [ncases] [seed]
"""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), ROOT]
from test_parallel_gloo import _emulated_flow_worker, _free_port  # noqa: E402


def draw(rng):
    size = int(rng.choice([2, 3, 4]))
    kind = str(rng.choice(["advection", "burgers", "diffusion", "incompressible", "lm_atm"], p=[0.3, 0.3, 0.2, 0.1, 0.1]))
    per = int(rng.integers(4, 13))                       # rows per slab (>= ng)
    ny = int(rng.choice([8, 17, 24, 32]))
    if kind == "advection":
        problem = str(rng.choice(["smooth", "tophat"]))
        inputs = {"mesh.nx": size * per, "mesh.ny": ny, "advection.u": float(rng.choice([1.0, -0.6, 0.3])),
                  "advection.v": float(rng.choice([1.0, -1.0, 0.0])), "advection.limiter": int(rng.integers(3))}
        xb = str(rng.choice(["periodic", "outflow"]))
        yb = str(rng.choice(["periodic", "outflow"]))
        inputs.update({"mesh.xlboundary": xb, "mesh.xrboundary": xb, "mesh.ylboundary": yb, "mesh.yrboundary": yb})
        nsteps = int(rng.integers(2, 7))
    elif kind == "burgers":
        problem = str(rng.choice(["test", "tophat", "converge"]))
        inputs = {"mesh.nx": size * per, "mesh.ny": ny, "advection.limiter": int(rng.integers(3))}
        xb = str(rng.choice(["periodic", "outflow"]))
        yb = str(rng.choice(["periodic", "outflow"]))
        inputs.update({"mesh.xlboundary": xb, "mesh.xrboundary": xb, "mesh.ylboundary": yb, "mesh.yrboundary": yb})
        nsteps = int(rng.integers(2, 7))
    elif kind == "diffusion":
        problem = "gaussian"
        n = 128                                          # the finest level must be split: >= 128 columns, >= 20 rows per
        size = int(rng.choice([2, 4]))                   # slab (the library refuses smaller decomposed hierarchies)
        xb = str(rng.choice(["neumann", "dirichlet", "periodic"]))
        yb = str(rng.choice(["neumann", "dirichlet", "periodic"]))
        inputs = {"mesh.nx": n, "mesh.ny": n, "mesh.xlboundary": xb, "mesh.xrboundary": xb, "mesh.ylboundary": yb,
                  "mesh.yrboundary": yb, "diffusion.mg_split_n": int(rng.choice([64, 128])),
                  "driver.cfl": float(rng.choice([0.7, 2.0]))}
        nsteps = int(rng.integers(1, 3))
    elif kind == "incompressible":
        problem = str(rng.choice(["shear", "converge"]))
        n, size = 128, int(rng.choice([2, 4]))
        inputs = {"mesh.nx": n, "mesh.ny": n, "incompressible.limiter": int(rng.integers(3)),
                  "incompressible.proj_type": int(rng.choice([1, 2])), "incompressible.mg_split_n": int(rng.choice([64, 128]))}
        nsteps = 0                                       # initialize_problem already runs a projection and one full step
    else:
        problem = "bubble"
        n = int(rng.choice([32, 64]))
        size = int(rng.choice([2, 4]))
        inputs = {"mesh.nx": n, "mesh.ny": n, "lm-atmosphere.limiter": int(rng.integers(1, 3)),
                  "lm-atmosphere.proj_type": int(rng.choice([1, 2]))}
        nsteps = 1
    return size, kind, problem, inputs, nsteps


if __name__ == "__main__":
    ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    ctx = mp.get_context("spawn")
    bad, seen = 0, {}
    for c in range(ncases):
        size, solver, problem, inputs, nsteps = draw(rng)
        seen[solver] = seen.get(solver, 0) + 1
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_emulated_flow_worker, args=(r, size, port, solver, problem, inputs, nsteps, q))
                 for r in range(size)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(1800)
        ok = all(p.exitcode == 0 for p in procs)
        res = dict(q.get(timeout=5) for _ in range(size)) if ok else {}
        ok = ok and res == {r: True for r in range(size)}
        bad += not ok
        print("ok  " if ok else "FAIL", c, dict(size=size, solver=solver, problem=problem, nsteps=nsteps, **inputs), flush=True)
    print(f"{ncases} cases, {bad} failed;  solvers: " + ", ".join(f"{k} {v}" for k, v in sorted(seen.items())))
    sys.exit(1 if bad else 0)

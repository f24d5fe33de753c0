"""Randomised check of Pyro.single_step_streamed() (host-resident state, row blocks host -> device -> host) against resident
steps on the emulated device: random grid shapes, block counts, problems, boundary types, Riemann solvers; every dt and the
final state bit for bit.  Found that the CGF solver's solid-wall rule (xl_solid) was applied at the low-x face of EVERY block
when the domain's -x boundary reflects, and that every block read the heating profile's FIRST rows.  Development tool (CPU only):

    python scripts/fuzz_streamed_emulated.py [ncases] [seed]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), ROOT]
import emu_device
real_empty = torch.empty
def empty(*a, **k):
    k.pop("pin_memory", None); return real_empty(*a, **k)
torch.empty = empty
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
with emu_device.emulated_device():
    from pyro2_b200.pyro_sim import Pyro
    for c in range(N):
        nx, ny = int(rng.integers(16, 90)), int(rng.integers(8, 40))
        nchunks = int(rng.integers(1, 24))
        problem = str(rng.choice(["sedov", "quad", "sod", "advect", "kh", "heating", "rt", "bubble", "gresho", "acoustic_pulse"]))
        xb = str(rng.choice(["outflow", "reflect"]))
        yb = str(rng.choice(["outflow", "reflect", "periodic"]))
        inputs = {"mesh.nx": nx, "mesh.ny": ny, "driver.max_steps": 10**6, "driver.tmax": 1e9, "driver.verbose": 0,
                  "mesh.xlboundary": xb, "mesh.xrboundary": str(rng.choice([xb, "outflow"])), "mesh.ylboundary": yb, "mesh.yrboundary": yb,
                  "compressible.riemann": str(rng.choice(["HLLC", "CGF", "HLLC_lm"])), "compressible.limiter": int(rng.integers(1, 3))}
        if rng.integers(4) == 0:                        # an active density floor (clean_state), applied block by block
            inputs["compressible.small_dens"] = {"sedov": 0.95, "quad": 0.6, "sod": 0.2}.get(problem, 0.99)
        if problem == "sedov":
            inputs["sedov.r_init"] = 0.2
        if problem in ("rt", "bubble"):                 # gravity; their stock hse boundaries are user hooks (refused here)
            inputs.update({"compressible.grav": -1.0, "mesh.ymax": 3.0})
            if yb == "periodic":
                inputs.update({"mesh.ylboundary": "reflect", "mesh.yrboundary": "reflect"})
        def make():
            p = Pyro("compressible"); p.initialize_problem(problem, inputs_dict=inputs); return p
        try:
            ref, p = make(), make()
            planes = p.sim.cc_data.planes
            bufs = [torch.empty(planes.shape, dtype=planes.dtype) for _ in range(2)]
            bufs[0].copy_(planes)
            nsteps = int(rng.integers(2, 6)); ok = True
            for step in range(nsteps):
                ref.single_step(); p.single_step_streamed(bufs[step % 2], bufs[(step + 1) % 2], nchunks=nchunks)
                ok = ok and p.sim.dt == ref.sim.dt
            g = p.sim.cc_data.grid
            host = bufs[nsteps % 2][:, g.ilo:g.ihi + 1, g.jlo:g.jhi + 1]
            want = ref.sim.cc_data.planes[:, g.ilo:g.ihi + 1, g.jlo:g.jhi + 1]
            ok = ok and torch.equal(host, want)
        except Exception as e:
            ok = False; print("EXC", repr(e)[:300])
        bad += not ok
        print("ok  " if ok else "FAIL", c, problem, nx, ny, nchunks, xb, yb, inputs["compressible.riemann"], flush=True)
print(N, "cases", bad, "failed")

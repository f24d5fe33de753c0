r"""Smooth convergence test (Minion 1996) on the periodic unit square; same setup as
pyro/incompressible/problems/converge.py:

    u = 1 - 2 cos(2 pi x) sin(2 pi y),   v = 1 + 2 sin(2 pi x) cos(2 pi y)

with the exact solution translating at unit speed in x and y."""
import math

import numpy as np

from ...util import msg

DEFAULT_INPUTS = "inputs.converge.64"

# stock run (the reference's inputs.converge.64)
INPUTS = {"driver.max_steps": 1000, "driver.tmax": 0.5, "driver.cfl": 0.5, "driver.init_tstep_factor": 1.0,
          "driver.fix_dt": 2.5e-3, "io.basename": "converge_64_", "io.n_out": 20, "mesh.nx": 64, "mesh.ny": 64,
          "mesh.xmax": 1.0, "mesh.ymax": 1.0, "mesh.xlboundary": "periodic", "mesh.xrboundary": "periodic",
          "mesh.ylboundary": "periodic", "mesh.yrboundary": "periodic"}

PROBLEM_PARAMS = {}


def init_data(my_data, rp):
    if rp.get_param("driver.verbose"):
        msg.bold("initializing the incompressible converge problem...")
    g = my_data.grid
    if g.xmin != 0 or g.xmax != 1 or g.ymin != 0 or g.ymax != 1:
        msg.fail("ERROR: domain should be a unit square")
    x = np.broadcast_to(g.x[:, None], (g.qx, g.qy))
    y = np.broadcast_to(g.y[None, :], (g.qx, g.qy))
    my_data.get_var("x-velocity")[:, :] = 1.0 - 2.0 * np.cos(2.0 * math.pi * x) * np.sin(2.0 * math.pi * y)
    my_data.get_var("y-velocity")[:, :] = 1.0 + 2.0 * np.sin(2.0 * math.pi * x) * np.cos(2.0 * math.pi * y)


def finalize():
    print("""
          Comparisons to the analytic solution: pyro's analysis/incomp_converge_error.py
          """)

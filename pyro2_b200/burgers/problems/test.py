"""A shock running diagonally through the domain: u = v = 3 below the line y = 1 - x, 1 above it
(same setup and parameters as pyro/burgers/problems/test.py)."""
import numpy as np

from ...util import msg

DEFAULT_INPUTS = "inputs.test"

# stock run (the reference's inputs.test)
INPUTS = {"driver.max_steps": 500, "driver.tmax": 0.1, "driver.max_dt_change": 1.e33, "driver.init_tstep_factor": 1.0,
          "driver.cfl": 0.8, "io.basename": "test_", "io.n_out": 10, "mesh.nx": 128, "mesh.ny": 128,
          "mesh.xmax": 1.0, "mesh.ymax": 1.0, "mesh.xlboundary": "outflow", "mesh.xrboundary": "outflow",
          "mesh.ylboundary": "outflow", "mesh.yrboundary": "outflow", "advection.limiter": 2}

PROBLEM_PARAMS = {}


def init_data(myd, rp):
    if rp.get_param("driver.verbose"):
        msg.bold("initializing the burgers test problem...")
    g = myd.grid
    above = g.y[None, :] > -1.0 * g.x[:, None] + 1.0
    vel = np.where(above, 1.0, 3.0)
    myd.get_var("x-velocity")[:, :] = vel
    myd.get_var("y-velocity")[:, :] = vel


def finalize():
    pass

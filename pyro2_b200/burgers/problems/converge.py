"""Smooth velocity field, each component a constant plus a Gaussian bump: for convergence tests of the
Burgers solver before the profile steepens (same setup as pyro/burgers/problems/converge.py)."""
import numpy as np

from ...util import msg

DEFAULT_INPUTS = "inputs.converge.64"

# stock run (the reference's inputs.converge.64, without its tracer particles: not part of this build)
INPUTS = {"driver.max_steps": 500, "driver.tmax": 1.0, "driver.max_dt_change": 1.e33, "driver.init_tstep_factor": 1.0,
          "driver.fix_dt": 0.005, "driver.cfl": 0.8, "io.basename": "converge.64_", "io.dt_out": 0.2,
          "mesh.nx": 64, "mesh.ny": 64, "mesh.xmax": 1.0, "mesh.ymax": 1.0,
          "mesh.xlboundary": "periodic", "mesh.xrboundary": "periodic",
          "mesh.ylboundary": "periodic", "mesh.yrboundary": "periodic", "advection.limiter": 0}

PROBLEM_PARAMS = {}


def init_data(my_data, rp):
    if rp.get_param("driver.verbose"):
        msg.bold("initializing the smooth burgers convergence problem...")
    g = my_data.grid
    xctr, yctr = 0.5 * (g.xmin + g.xmax), 0.5 * (g.ymin + g.ymax)
    x = np.broadcast_to(g.x[:, None], (g.qx, g.qy))
    y = np.broadcast_to(g.y[None, :], (g.qx, g.qy))
    A = 0.05
    bump = A + A * np.exp(-50.0 * ((x - xctr) ** 2 + (y - yctr) ** 2))
    my_data.get_var("x-velocity")[:, :] = bump
    my_data.get_var("y-velocity")[:, :] = bump


def finalize():
    pass

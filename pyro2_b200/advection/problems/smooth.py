"""Gaussian bump on a unit background (smooth: the limiters barely act, used for convergence testing);
same setup as pyro/advection/problems/smooth.py."""
import numpy as np

from ...util import msg

DEFAULT_INPUTS = "inputs.smooth"

# stock run (the reference's inputs.smooth, without its tracer particles)
INPUTS = {"driver.max_steps": 500, "driver.tmax": 1.0, "driver.max_dt_change": 1.e33, "driver.init_tstep_factor": 1.0,
          "driver.cfl": 0.8, "io.basename": "smooth_", "io.dt_out": 0.2, "mesh.nx": 32, "mesh.ny": 32,
          "mesh.xmax": 1.0, "mesh.ymax": 1.0, "mesh.xlboundary": "periodic", "mesh.xrboundary": "periodic",
          "mesh.ylboundary": "periodic", "mesh.yrboundary": "periodic",
          "advection.u": 1.0, "advection.v": 1.0, "advection.limiter": 2}

PROBLEM_PARAMS = {}


def init_data(my_data, rp):
    if rp.get_param("driver.verbose"):
        msg.bold("initializing the smooth advection problem...")
    g = my_data.grid
    xctr, yctr = 0.5 * (g.xmin + g.xmax), 0.5 * (g.ymin + g.ymax)
    x = np.broadcast_to(g.x[:, None], (g.qx, g.qy))
    y = np.broadcast_to(g.y[None, :], (g.qx, g.qy))
    my_data.get_var("density")[:, :] = 1.0 + np.exp(-60.0 * ((x - xctr) ** 2 + (y - yctr) ** 2))


def finalize():
    pass

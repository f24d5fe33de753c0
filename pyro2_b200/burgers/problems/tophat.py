"""A disc moving diagonally at unit speed through fluid at rest: drives a curved shock ahead of it and a
rarefaction behind (same setup as pyro/burgers/problems/tophat.py)."""
import numpy as np

from ...util import msg

DEFAULT_INPUTS = "inputs.tophat"

# stock run (the reference's inputs.tophat)
INPUTS = {"driver.max_steps": 500, "driver.tmax": 1.0, "driver.max_dt_change": 1.e33, "driver.init_tstep_factor": 1.0,
          "driver.cfl": 0.8, "io.basename": "tophat_", "io.n_out": 10, "mesh.nx": 32, "mesh.ny": 32,
          "mesh.xmax": 1.0, "mesh.ymax": 1.0, "mesh.xlboundary": "periodic", "mesh.xrboundary": "periodic",
          "mesh.ylboundary": "periodic", "mesh.yrboundary": "periodic", "advection.limiter": 2}

PROBLEM_PARAMS = {}


def init_data(myd, rp):
    if rp.get_param("driver.verbose"):
        msg.bold("initializing the tophat burgers problem...")
    g = myd.grid
    xctr, yctr = 0.5 * (g.xmin + g.xmax), 0.5 * (g.ymin + g.ymax)
    x = np.broadcast_to(g.x[:, None], (g.qx, g.qy))
    y = np.broadcast_to(g.y[None, :], (g.qx, g.qy))
    R = 0.1
    vel = np.where((x - xctr) ** 2 + (y - yctr) ** 2 < R ** 2, 1.0, 0.0)
    myd.get_var("x-velocity")[:, :] = vel
    myd.get_var("y-velocity")[:, :] = vel


def finalize():
    pass

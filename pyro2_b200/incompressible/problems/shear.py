r"""Doubly periodic shear layer (Martin & Colella 2000) on the unit square; same parameters as
pyro/incompressible/problems/shear.py:

    u = tanh(rho_s (y - 1/4))  for y <= 1/2,   tanh(rho_s (3/4 - y))  for y > 1/2
    v = delta_s sin(2 pi x)

The initial data are evaluated on the host with numpy, like the reference's, and uploaded."""
import math

import numpy as np

from ...util import msg

DEFAULT_INPUTS = "inputs.shear"

# stock run (the reference's inputs.shear)
INPUTS = {"driver.max_steps": 2000, "driver.tmax": 1.0, "driver.cfl": 0.8, "io.basename": "shear_128_", "io.n_out": 1,
          "mesh.nx": 128, "mesh.ny": 128, "mesh.xmax": 1.0, "mesh.ymax": 1.0,
          "mesh.xlboundary": "periodic", "mesh.xrboundary": "periodic",
          "mesh.ylboundary": "periodic", "mesh.yrboundary": "periodic", "shear.rho_s": 42.0, "shear.delta_s": 0.05}

PROBLEM_PARAMS = {"shear.rho_s": 42.0,     # inverse width of the shear layers
                  "shear.delta_s": 0.05}   # amplitude of the perturbation


def init_data(my_data, rp):
    if rp.get_param("driver.verbose"):
        msg.bold("initializing the incompressible shear problem...")
    rho_s = rp.get_param("shear.rho_s")
    delta_s = rp.get_param("shear.delta_s")
    g = my_data.grid
    if g.xmin != 0 or g.xmax != 1 or g.ymin != 0 or g.ymax != 1:
        msg.fail("ERROR: domain should be a unit square")
    y_half = 0.5 * (g.ymin + g.ymax)
    y = np.broadcast_to(g.y[None, :], (g.qx, g.qy))
    x = np.broadcast_to(g.x[:, None], (g.qx, g.qy))
    lower = y <= y_half
    u = np.where(lower, np.tanh(rho_s * (y - 0.25)), np.tanh(rho_s * (0.75 - y)))
    my_data.get_var("x-velocity")[:, :] = u
    my_data.get_var("y-velocity")[:, :] = delta_s * np.sin(2.0 * math.pi * x)
    if rp.get_param("driver.verbose"):
        print("extrema: ", u.min(), u.max())


def finalize():
    pass

"""Shock tube (Sod by default) along x or y; same parameters as
pyro/compressible/problems/sod.py."""
import numpy as np

from ...util import msg

DEFAULT_INPUTS = "inputs.sod.x"

# stock run: the 128 x 10 tube along x used by the reference's regression suite (inputs.sod.x)
INPUTS = {"driver.max_steps": 200, "driver.tmax": 0.2, "compressible.limiter": 1,
          "io.basename": "sod_x_", "io.dt_out": 0.05, "mesh.nx": 128, "mesh.ny": 10,
          "mesh.xmax": 1.0, "mesh.ymax": 0.05, "mesh.xlboundary": "outflow", "mesh.xrboundary": "outflow",
          "sod.direction": "x", "sod.dens_left": 1.0, "sod.dens_right": 0.125,
          "sod.u_left": 0.0, "sod.u_right": 0.0, "sod.p_left": 1.0, "sod.p_right": 0.1}

PROBLEM_PARAMS = {"sod.direction": "x", "sod.dens_left": 1.0, "sod.dens_right": 0.125,
                  "sod.u_left": 0.0, "sod.u_right": 0.0, "sod.p_left": 1.0, "sod.p_right": 0.1}


def init_data(my_data, rp):
    if rp.get_param("driver.verbose"):
        msg.bold("initializing the sod problem...")
    g = my_data.grid
    gamma = rp.get_param("eos.gamma")
    dl, dr = rp.get_param("sod.dens_left"), rp.get_param("sod.dens_right")
    ul, ur = rp.get_param("sod.u_left"), rp.get_param("sod.u_right")
    pl, pr = rp.get_param("sod.p_left"), rp.get_param("sod.p_right")
    if rp.get_param("sod.direction") == "x":
        ctr = 0.5 * (rp.get_param("mesh.xmin") + rp.get_param("mesh.xmax"))
        left = np.broadcast_to(g.x[:, None] <= ctr, (g.qx, g.qy))
        normal, transverse = "x-momentum", "y-momentum"
    else:
        ctr = 0.5 * (rp.get_param("mesh.ymin") + rp.get_param("mesh.ymax"))
        left = np.broadcast_to(g.y[None, :] <= ctr, (g.qx, g.qy))
        normal, transverse = "y-momentum", "x-momentum"
    dens = np.where(left, dl, dr)
    mom = np.where(left, dl * ul, dr * ur)
    ener = np.where(left, pl / (gamma - 1.0) + 0.5 * (dl * ul) * ul, pr / (gamma - 1.0) + 0.5 * (dr * ur) * ur)
    my_data.get_var("density")[:, :] = dens
    my_data.get_var(normal)[:, :] = mom
    my_data.get_var(transverse)[:, :] = 0.0
    my_data.get_var("energy")[:, :] = ener


def finalize():
    pass

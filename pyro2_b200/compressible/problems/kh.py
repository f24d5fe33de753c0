"""Kelvin-Helmholtz double shear layer (McNally et al. 2012 style) on a periodic box; same
parameters as pyro/compressible/problems/kh.py."""
import numpy as np

from ...util import msg

DEFAULT_INPUTS = "inputs.kh"

# stock run: 64^2 doubly periodic box (the reference's inputs.kh)
INPUTS = {"driver.max_steps": 5000, "driver.tmax": 2.0, "compressible.limiter": 2, "compressible.cvisc": 0.1,
          "io.basename": "kh_", "eos.gamma": 1.4, "mesh.nx": 64, "mesh.ny": 64, "mesh.xmax": 1.0, "mesh.ymax": 1.0,
          "mesh.xlboundary": "periodic", "mesh.xrboundary": "periodic",
          "mesh.ylboundary": "periodic", "mesh.yrboundary": "periodic",
          "kh.rho_1": 1, "kh.u_1": -0.5, "kh.rho_2": 2, "kh.u_2": 0.5}

PROBLEM_PARAMS = {"kh.rho_1": 1.0, "kh.u_1": -1.0, "kh.rho_2": 2.0, "kh.u_2": 1.0,
                  "kh.bulk_velocity": 0.0}


def init_data(my_data, rp):
    if rp.get_param("driver.verbose"):
        msg.bold("initializing the Kelvin-Helmholtz problem...")
    g = my_data.grid
    gamma = rp.get_param("eos.gamma")
    rho_1, u_1 = rp.get_param("kh.rho_1"), rp.get_param("kh.u_1")
    rho_2, u_2 = rp.get_param("kh.rho_2"), rp.get_param("kh.u_2")
    bulk = rp.get_param("kh.bulk_velocity")
    width, w0 = 0.025, 0.01
    vm, rhom = 0.5 * (u_1 - u_2), 0.5 * (rho_1 - rho_2)

    y = np.broadcast_to(g.y[None, :], (g.qx, g.qy))
    x = np.broadcast_to(g.x[:, None], (g.qx, g.qy))
    bands = [(y < 0.25, rho_1, u_1, -1.0, y - 0.25), ((y >= 0.25) & (y < 0.5), rho_2, u_2, 1.0, 0.25 - y),
             ((y >= 0.5) & (y < 0.75), rho_2, u_2, 1.0, y - 0.75), (y >= 0.75, rho_1, u_1, -1.0, 0.75 - y)]
    dens = np.ones((g.qx, g.qy))
    u = np.zeros((g.qx, g.qy))
    for mask, rho0, u0, sign, arg in bands:
        prof = np.exp(arg[mask] / width)
        dens[mask] = rho0 + sign * rhom * prof
        u[mask] = u0 + sign * vm * prof
    xmom = u * dens
    ymom = dens * (bulk + w0 * np.sin(4 * np.pi * x))
    p = 2.5
    ener = p / (gamma - 1.0) + 0.5 * (xmom ** 2 + ymom ** 2) / dens
    my_data.get_var("density")[:, :] = dens
    my_data.get_var("x-momentum")[:, :] = xmom
    my_data.get_var("y-momentum")[:, :] = ymom
    my_data.get_var("energy")[:, :] = ener


def finalize():
    pass

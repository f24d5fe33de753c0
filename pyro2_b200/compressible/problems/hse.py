"""Isothermal atmosphere in hydrostatic equilibrium (a test of the hse boundaries and the gravity source:
it should stay at rest); same parameters as pyro/compressible/problems/hse.py."""
import numpy as np

from ...util import msg

DEFAULT_INPUTS = "inputs.hse"

# stock run (the reference's inputs.hse)
INPUTS = {"driver.max_steps": 10000, "driver.tmax": 3.0, "io.basename": "hse_", "io.n_out": 100,
          "mesh.nx": 64, "mesh.ny": 192, "mesh.xmax": 1.0, "mesh.ymax": 3.0,
          "mesh.xlboundary": "periodic", "mesh.xrboundary": "periodic", "mesh.ylboundary": "hse", "mesh.yrboundary": "hse",
          "hse.dens0": 1.0, "hse.h": 1.0, "compressible.grav": -1.0, "compressible.limiter": 2}

PROBLEM_PARAMS = {"hse.dens0": 1.0, "hse.h": 1.0}


def stratify(g, dens_of_y, cs2, grav):
    """density rows from dens_of_y(y_j) and the pressure from a trapezoidal integration of dp/dy = rho g,
    valid rows only (ghost rows stay empty until the first boundary fill)"""
    dens = np.zeros((g.qx, g.qy))
    p = np.zeros((g.qx, g.qy))
    for j in range(g.jlo, g.jhi + 1):
        dens[:, j] = dens_of_y(g.y[j])
        if j == g.jlo:
            p[:, j] = dens[:, j] * cs2
        else:
            p[:, j] = p[:, j - 1] + 0.5 * g.dy * (dens[:, j] + dens[:, j - 1]) * grav
    return dens, p


def init_data(my_data, rp):
    if rp.get_param("driver.verbose"):
        msg.bold("initializing the HSE problem...")
    g = my_data.grid
    gamma = rp.get_param("eos.gamma")
    grav = rp.get_param("compressible.grav")
    dens0, H = rp.get_param("hse.dens0"), rp.get_param("hse.h")
    cs2 = H * abs(grav)                       # isothermal sound speed squared
    dens, p = stratify(g, lambda yj: dens0 * np.exp(-yj / H), cs2, grav)
    xmom = np.zeros((g.qx, g.qy))
    ymom = np.zeros((g.qx, g.qy))
    with np.errstate(invalid="ignore", divide="ignore"):
        ener = p / (gamma - 1.0) + 0.5 * (xmom ** 2 + ymom ** 2) / dens
    my_data.get_var("density")[:, :] = dens
    my_data.get_var("x-momentum")[:, :] = xmom
    my_data.get_var("y-momentum")[:, :] = ymom
    my_data.get_var("energy")[:, :] = ener


def finalize():
    pass

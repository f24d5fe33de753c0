"""Rayleigh-Taylor instability: heavy fluid over light in a constant gravitational field with a
single-mode velocity perturbation at the interface; same parameters as pyro/compressible/problems/rt.py.
Run with the hse boundaries in y."""
import numpy as np

from ...util import msg

DEFAULT_INPUTS = "inputs.rt"

# stock run (the reference's inputs.rt)
INPUTS = {"driver.max_steps": 10000, "driver.tmax": 3.0, "io.basename": "rt_", "io.n_out": 100,
          "mesh.nx": 64, "mesh.ny": 192, "mesh.xmax": 1.0, "mesh.ymax": 3.0,
          "mesh.xlboundary": "periodic", "mesh.xrboundary": "periodic", "mesh.ylboundary": "hse", "mesh.yrboundary": "hse",
          "rt.amp": 0.25, "compressible.grav": -1.0, "compressible.limiter": 2}

PROBLEM_PARAMS = {"rt.dens1": 1.0, "rt.dens2": 2.0, "rt.amp": 1.0, "rt.sigma": 0.1, "rt.p0": 10.0}


def init_data(my_data, rp):
    if rp.get_param("driver.verbose"):
        msg.bold("initializing the rt problem...")
    g = my_data.grid
    gamma = rp.get_param("eos.gamma")
    grav = rp.get_param("compressible.grav")
    dens1, dens2 = rp.get_param("rt.dens1"), rp.get_param("rt.dens2")
    p0, amp, sigma = rp.get_param("rt.p0"), rp.get_param("rt.amp"), rp.get_param("rt.sigma")
    ycenter = 0.5 * (g.ymin + g.ymax)
    # stratification of the valid rows; ghost rows stay empty until the first boundary fill
    dens = np.zeros((g.qx, g.qy))
    p = np.zeros((g.qx, g.qy))
    for j in range(g.jlo, g.jhi + 1):
        if g.y[j] < ycenter:
            dens[:, j] = dens1
            p[:, j] = p0 + dens1 * grav * g.y[j]
        else:
            dens[:, j] = dens2
            p[:, j] = p0 + dens1 * grav * ycenter + dens2 * grav * (g.y[j] - ycenter)
    x = np.broadcast_to(g.x[:, None], (g.qx, g.qy))
    y = np.broadcast_to(g.y[None, :], (g.qx, g.qy))
    L = g.xmax - g.xmin
    ymom = amp * 0.5 * (np.cos(2.0 * np.pi * x / L) + np.cos(2.0 * np.pi * (L - x) / L)) * \
        np.exp(-(y - ycenter) ** 2 / sigma ** 2)
    ymom = ymom * dens
    xmom = np.zeros((g.qx, g.qy))
    with np.errstate(invalid="ignore", divide="ignore"):     # 0/0 in the still-empty ghost rows, as in the reference
        ener = p / (gamma - 1.0) + 0.5 * (xmom ** 2 + ymom ** 2) / dens
    my_data.get_var("density")[:, :] = dens
    my_data.get_var("x-momentum")[:, :] = xmom
    my_data.get_var("y-momentum")[:, :] = ymom
    my_data.get_var("energy")[:, :] = ener


def finalize():
    pass

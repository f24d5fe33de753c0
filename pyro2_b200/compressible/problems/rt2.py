"""Rayleigh-Taylor instability with two perturbation wavelengths side by side -- short on the left third of
the domain, long on the rest -- to show how the growth rate depends on wavenumber; same parameters as
pyro/compressible/problems/rt2.py.  Run with the hse boundaries in y."""
import numpy as np

from ...util import msg

DEFAULT_INPUTS = "inputs.rt2"

# stock run (the reference's inputs.rt2)
INPUTS = {"driver.max_steps": 10000, "driver.tmax": 5.0, "io.basename": "rt_", "io.n_out": 10000, "io.dt_out": 0.025,
          "mesh.nx": 384, "mesh.ny": 192, "mesh.xmax": 6.0, "mesh.ymax": 3.0,
          "mesh.xlboundary": "periodic", "mesh.xrboundary": "periodic", "mesh.ylboundary": "hse", "mesh.yrboundary": "hse",
          "rt2.amp": 0.1, "rt2.sigma": 0.025, "compressible.grav": -1.0, "compressible.limiter": 2}

PROBLEM_PARAMS = {"rt2.dens1": 1.0, "rt2.dens2": 2.0, "rt2.amp": 1.0, "rt2.sigma": 0.1, "rt2.p0": 10.0}

# perturbation frequencies of the left / right part
F_LEFT, F_RIGHT = 18, 3


def stratified(g, dens1, dens2, p0, grav):
    """two constant-density layers in hydrostatic equilibrium, valid rows only (ghost rows stay empty until the
    first boundary fill); shared by the Rayleigh-Taylor setups"""
    ycenter = 0.5 * (g.ymin + g.ymax)
    dens = np.zeros((g.qx, g.qy))
    p = np.zeros((g.qx, g.qy))
    for j in range(g.jlo, g.jhi + 1):
        if g.y[j] < ycenter:
            dens[:, j] = dens1
            p[:, j] = p0 + dens1 * grav * g.y[j]
        else:
            dens[:, j] = dens2
            p[:, j] = p0 + dens1 * grav * ycenter + dens2 * grav * (g.y[j] - ycenter)
    return dens, p, ycenter


def store(my_data, dens, p, ymom, gamma):
    xmom = np.zeros_like(dens)
    with np.errstate(invalid="ignore", divide="ignore"):     # 0/0 in the still-empty ghost rows, as in the reference
        ener = p / (gamma - 1.0) + 0.5 * (xmom ** 2 + ymom ** 2) / dens
    my_data.get_var("density")[:, :] = dens
    my_data.get_var("x-momentum")[:, :] = xmom
    my_data.get_var("y-momentum")[:, :] = ymom
    my_data.get_var("energy")[:, :] = ener


def init_data(my_data, rp):
    if rp.get_param("driver.verbose"):
        msg.bold("initializing the rt problem...")
    g = my_data.grid
    amp, sigma = rp.get_param("rt2.amp"), rp.get_param("rt2.sigma")
    dens, p, ycenter = stratified(g, rp.get_param("rt2.dens1"), rp.get_param("rt2.dens2"), rp.get_param("rt2.p0"),
                                  rp.get_param("compressible.grav"))
    x = np.broadcast_to(g.x[:, None], (g.qx, g.qy))
    y = np.broadcast_to(g.y[None, :], (g.qx, g.qy))
    L = g.xmax - g.xmin
    freq = np.where(x < L / 3.0, float(F_LEFT), float(F_RIGHT))
    ymom = amp * np.sin(4.0 * np.pi * freq * x / L) * np.exp(-(y - ycenter) ** 2 / sigma ** 2)
    store(my_data, dens, p, ymom * dens, rp.get_param("eos.gamma"))


def finalize():
    pass

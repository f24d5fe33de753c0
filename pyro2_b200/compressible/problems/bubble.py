"""A hot, buoyant bubble in an isothermal, hydrostatic atmosphere; same parameters as
pyro/compressible/problems/bubble.py."""
import numpy as np

from ...util import msg
from .hse import stratify

DEFAULT_INPUTS = "inputs.bubble"

# stock run (the reference's inputs.bubble)
INPUTS = {"driver.max_steps": 1000, "driver.tmax": 100.0, "io.basename": "bubble_", "io.n_out": 100,
          "mesh.nx": 128, "mesh.ny": 256, "mesh.xmax": 4.0, "mesh.ymax": 8.0,
          "mesh.xlboundary": "outflow", "mesh.xrboundary": "outflow", "mesh.ylboundary": "hse", "mesh.yrboundary": "hse",
          "bubble.scale_height": 1.0, "bubble.dens_base": 1000.0, "bubble.x_pert": 2.0, "bubble.y_pert": 2.0,
          "bubble.r_pert": 0.25, "bubble.pert_amplitude_factor": 2.0, "compressible.grav": -2.0,
          "compressible.limiter": 2}

PROBLEM_PARAMS = {"bubble.dens_base": 10.0,              # density at the base of the atmosphere
                  "bubble.scale_height": 2.0,            # scale height of the isothermal atmosphere
                  "bubble.x_pert": 2.0, "bubble.y_pert": 2.0, "bubble.r_pert": 0.25,
                  "bubble.pert_amplitude_factor": 5.0,   # boost of the specific internal energy in the bubble
                  "bubble.dens_cutoff": 0.01}


def init_data(my_data, rp):
    if rp.get_param("driver.verbose"):
        msg.bold("initializing the bubble problem...")
    g = my_data.grid
    gamma = rp.get_param("eos.gamma")
    grav = rp.get_param("compressible.grav")
    scale_height, dens_base = rp.get_param("bubble.scale_height"), rp.get_param("bubble.dens_base")
    dens_cutoff = rp.get_param("bubble.dens_cutoff")
    x_pert, y_pert, r_pert = rp.get_param("bubble.x_pert"), rp.get_param("bubble.y_pert"), rp.get_param("bubble.r_pert")
    factor = rp.get_param("bubble.pert_amplitude_factor")
    cs2 = scale_height * abs(grav)
    dens, p = stratify(g, lambda yj: max(dens_base * np.exp(-yj / scale_height), dens_cutoff), cs2, grav)
    dens[:, :g.jlo] = dens_cutoff            # the reference starts from dens = dens_cutoff everywhere
    dens[:, g.jhi + 1:] = dens_cutoff
    xmom = np.zeros((g.qx, g.qy))
    ymom = np.zeros((g.qx, g.qy))
    ener = p / (gamma - 1.0) + 0.5 * (xmom ** 2 + ymom ** 2) / dens
    x = np.broadcast_to(g.x[:, None], (g.qx, g.qy))
    y = np.broadcast_to(g.y[None, :], (g.qx, g.qy))
    idx = np.sqrt((x - x_pert) ** 2 + (y - y_pert) ** 2) <= r_pert
    # raise the specific internal energy inside the bubble at constant pressure by lowering the density
    eint = (ener[idx] - 0.5 * (xmom[idx] ** 2 - ymom[idx] ** 2) / dens[idx]) / dens[idx]
    pres = dens[idx] * eint * (gamma - 1.0)
    eint = eint * factor
    dens[idx] = pres / (eint * (gamma - 1.0))
    ener[idx] = dens[idx] * eint + 0.5 * (xmom[idx] ** 2 + ymom[idx] ** 2) / dens[idx]
    my_data.get_var("density")[:, :] = dens
    my_data.get_var("x-momentum")[:, :] = xmom
    my_data.get_var("y-momentum")[:, :] = ymom
    my_data.get_var("energy")[:, :] = ener


def finalize():
    pass

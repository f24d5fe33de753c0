"""A heat source at the bottom of an adiabatically stratified atmosphere drives a buoyant plume; same
parameters as pyro/compressible/problems/plume.py (run with the hse boundaries in y)."""
import numpy as np

from ...util import msg

DEFAULT_INPUTS = "inputs.plume"

# stock run (the reference's inputs.plume)
INPUTS = {"driver.max_steps": 10000, "driver.tmax": 10.0, "io.basename": "plume_", "io.n_out": 100,
          "mesh.nx": 128, "mesh.ny": 256, "mesh.xmax": 4.0, "mesh.ymax": 8.0,
          "mesh.xlboundary": "outflow", "mesh.xrboundary": "outflow", "mesh.ylboundary": "hse", "mesh.yrboundary": "hse",
          "plume.scale_height": 3.0, "plume.dens_base": 1000.0, "plume.x_pert": 2.0, "plume.y_pert": 2.0,
          "plume.r_pert": 0.25, "plume.e_rate": 0.5, "compressible.grav": -2.0, "compressible.limiter": 2}

PROBLEM_PARAMS = {"plume.dens_base": 10.0,       # density at the base of the atmosphere
                  "plume.scale_height": 4.0,     # scale height of the atmosphere
                  "plume.x_pert": 2.0, "plume.y_pert": 2.0, "plume.r_pert": 0.25,
                  "plume.e_rate": 0.1, "plume.dens_cutoff": 0.01}


def adiabatic_rows(g, gamma, grav, scale_height, dens_base, dens_cutoff, pressure_of_row):
    """density of an adiabatic atmosphere row by row (valid rows; dens_cutoff elsewhere) and the pressure from
    pressure_of_row(j, dens, p) -- the two stratified problems differ only in how they integrate p"""
    dens = np.full((g.qx, g.qy), dens_cutoff)
    p = np.zeros((g.qx, g.qy))
    for j in range(g.jlo, g.jhi + 1):
        profile = 1.0 - (gamma - 1.0) / gamma * g.y[j] / scale_height
        if profile > 0.0:
            dens[:, j] = max(dens_base * profile ** (1.0 / (gamma - 1.0)), dens_cutoff)
        else:
            dens[:, j] = dens_cutoff
        p[:, j] = pressure_of_row(j, dens, p)
    return dens, p


def init_data(my_data, rp):
    if rp.get_param("driver.verbose"):
        msg.bold("initializing the plume problem...")
    g = my_data.grid
    gamma = rp.get_param("eos.gamma")
    grav = rp.get_param("compressible.grav")
    scale_height, dens_base = rp.get_param("plume.scale_height"), rp.get_param("plume.dens_base")
    pres_base = scale_height * dens_base * abs(grav)

    def hydrostatic(j, dens, p):
        if j == g.jlo:
            return pres_base
        return p[:, j - 1] + 0.5 * g.dy * (dens[:, j] + dens[:, j - 1]) * grav
    dens, p = adiabatic_rows(g, gamma, grav, scale_height, dens_base, rp.get_param("plume.dens_cutoff"), hydrostatic)
    xmom = np.zeros((g.qx, g.qy))
    ymom = np.zeros((g.qx, g.qy))
    my_data.get_var("density")[:, :] = dens
    my_data.get_var("x-momentum")[:, :] = xmom
    my_data.get_var("y-momentum")[:, :] = ymom
    my_data.get_var("energy")[:, :] = p / (gamma - 1.0) + 0.5 * (xmom ** 2 + ymom ** 2) / dens


def heating(myg, rp):
    """(e_rate, profile): S_ener = dens * e_rate * exp(-(dist / r_pert)**2) around (x_pert, y_pert)"""
    x = np.broadcast_to(myg.x[:, None], (myg.qx, myg.qy))
    y = np.broadcast_to(myg.y[None, :], (myg.qx, myg.qy))
    dist = np.sqrt((x - rp.get_param("plume.x_pert")) ** 2 + (y - rp.get_param("plume.y_pert")) ** 2)
    return rp.get_param("plume.e_rate"), np.exp(-(dist / rp.get_param("plume.r_pert")) ** 2)


def source_terms(myg, U, ivars, rp):
    import torch
    rate, prof = heating(myg, rp)
    S = myg.scratch_array(nvar=ivars.nvar)
    S[:, :, ivars.iener] = U[:, :, ivars.idens].t() * rate * torch.from_numpy(prof).to(U.device)
    return S


def finalize():
    pass

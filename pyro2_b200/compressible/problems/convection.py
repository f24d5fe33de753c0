"""A heated layer above the bottom of an adiabatic atmosphere drives convection; reflecting floor, "ambient"
top, sponge in the low-density region.  Same parameters (and the same seeded velocity perturbation) as
pyro/compressible/problems/convection.py."""
import numpy as np

from ...util import msg
from .plume import adiabatic_rows

DEFAULT_INPUTS = "inputs.convection"

# stock run (the reference's inputs.convection)
INPUTS = {"driver.max_steps": 100000, "driver.tmax": 25.0, "io.basename": "convection_", "io.n_out": 100000000,
          "io.dt_out": 0.5, "mesh.nx": 128, "mesh.ny": 384, "mesh.xmax": 4.0, "mesh.ymax": 12.0,
          "mesh.xlboundary": "periodic", "mesh.xrboundary": "periodic", "mesh.ylboundary": "reflect",
          "mesh.yrboundary": "ambient", "convection.scale_height": 2.0, "convection.dens_base": 1000.0,
          "convection.dens_cutoff": 1.e-3, "convection.e_rate": 0.5, "sponge.do_sponge": 1,
          "compressible.grav": -2.0, "compressible.limiter": 2, "compressible.small_dens": 1.e-4}

PROBLEM_PARAMS = {"convection.dens_base": 10.0,      # density at the base of the atmosphere
                  "convection.scale_height": 4.0,    # scale height of the atmosphere
                  "convection.y_height": 2.0,        # height of the heated layer
                  "convection.thickness": 0.25,      # ... and its thickness
                  "convection.e_rate": 0.1, "convection.dens_cutoff": 0.01}


def init_data(my_data, rp):
    if rp.get_param("driver.verbose"):
        msg.bold("initializing the convection problem...")
    g = my_data.grid
    gamma = rp.get_param("eos.gamma")
    grav = rp.get_param("compressible.grav")
    scale_height, dens_base = rp.get_param("convection.scale_height"), rp.get_param("convection.dens_base")
    dens_cutoff = rp.get_param("convection.dens_cutoff")
    pres_base = scale_height * dens_base * abs(grav)

    def adiabat(j, dens, p):
        if j == g.jlo:
            return pres_base
        if dens[0, j] <= dens_cutoff + 1.e-30:
            return p[:, j - 1]
        return pres_base * (dens[:, j] / dens_base) ** gamma
    dens, p = adiabatic_rows(g, gamma, grav, scale_height, dens_base, dens_cutoff, adiabat)

    # the state beyond the top boundary
    my_data.set_aux("ambient_rho", dens_cutoff)
    my_data.set_aux("ambient_u", 0.0)
    my_data.set_aux("ambient_v", 0.0)
    my_data.set_aux("ambient_p", p[g.ilo:g.ihi + 1, g.jlo:g.jhi + 1].min())

    ener = p / (gamma - 1.0)
    # seeded velocity perturbation with Mach number < 0.05 wherever the gas is dense enough
    rng = np.random.default_rng(12345)
    # (drawn for the global array, of which an x-slab takes its rows: a decomposed run then starts from the
    # single-domain state)
    vel_pert = 2.0 * rng.random(size=(g.nx_global + 2 * g.ng, g.qy, 2))[g.ioffset:g.ioffset + g.qx] - 1
    with np.errstate(invalid="ignore", divide="ignore"):
        cs = np.sqrt(gamma * p / dens)
    vel_pert[:, :, 0] *= 0.05 * cs
    vel_pert[:, :, 1] *= 0.05 * cs
    idx = dens > 2 * dens_cutoff
    xmom = np.zeros((g.qx, g.qy))
    ymom = np.zeros((g.qx, g.qy))
    xmom[idx] = dens[idx] * vel_pert[idx, 0]
    ymom[idx] = dens[idx] * vel_pert[idx, 1]
    ener = ener + 0.5 * (xmom ** 2 + ymom ** 2) / dens
    my_data.get_var("density")[:, :] = dens
    my_data.get_var("x-momentum")[:, :] = xmom
    my_data.get_var("y-momentum")[:, :] = ymom
    my_data.get_var("energy")[:, :] = ener


def heating(myg, rp):
    """(e_rate, profile): S_ener = dens * e_rate * exp(-(|y - y_height| / thickness)**2)"""
    y = np.broadcast_to(myg.y[None, :], (myg.qx, myg.qy))
    dist = np.abs(y - rp.get_param("convection.y_height"))
    return rp.get_param("convection.e_rate"), np.exp(-(dist / rp.get_param("convection.thickness")) ** 2)


def source_terms(myg, U, ivars, rp):
    import torch
    rate, prof = heating(myg, rp)
    S = myg.scratch_array(nvar=ivars.nvar)
    S[:, :, ivars.iener] = U[:, :, ivars.idens].t() * rate * torch.from_numpy(prof).to(U.device)
    return S


def finalize():
    pass

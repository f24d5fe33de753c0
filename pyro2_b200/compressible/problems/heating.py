"""A localised heat source in a uniform medium at rest (it drives a Sedov-like expansion); same parameters
as pyro/compressible/problems/heating.py.  The source is dens * e_rate * exp(-(dist / r_src)**2) in the energy."""
import numpy as np

from ...util import msg

DEFAULT_INPUTS = "inputs.heating"

# stock run (the reference's inputs.heating)
INPUTS = {"driver.max_steps": 5000, "driver.tmax": 1.0, "compressible.limiter": 2, "compressible.cvisc": 0.1,
          "io.basename": "heating_", "io.dt_out": 0.1, "eos.gamma": 1.4, "mesh.nx": 64, "mesh.ny": 64,
          "mesh.xmax": 1.0, "mesh.ymax": 1.0, "mesh.xlboundary": "outflow", "mesh.xrboundary": "outflow",
          "mesh.ylboundary": "outflow", "mesh.yrboundary": "outflow",
          "heating.rho_ambient": 1.0, "heating.p_ambient": 10.0, "heating.r_src": 0.05, "heating.e_rate": 0.1}

PROBLEM_PARAMS = {"heating.rho_ambient": 1.0,   # ambient density
                  "heating.p_ambient": 10.0,    # ambient pressure
                  "heating.r_src": 0.1,         # size of the heating source
                  "heating.e_rate": 0.1}        # energy generation rate (energy / mass / time)


def init_data(my_data, rp):
    if rp.get_param("driver.verbose"):
        msg.bold("initializing the heating problem...")
    gamma = rp.get_param("eos.gamma")
    my_data.get_var("density")[:, :] = rp.get_param("heating.rho_ambient")
    my_data.get_var("x-momentum")[:, :] = 0.0
    my_data.get_var("y-momentum")[:, :] = 0.0
    my_data.get_var("energy")[:, :] = rp.get_param("heating.p_ambient") / (gamma - 1.0)


def heating(myg, rp):
    """(e_rate, profile): the source is S_ener = dens * e_rate * profile (what source_terms evaluates)"""
    xctr, yctr = 0.5 * (myg.xmin + myg.xmax), 0.5 * (myg.ymin + myg.ymax)
    x = np.broadcast_to(myg.x[:, None], (myg.qx, myg.qy))
    y = np.broadcast_to(myg.y[None, :], (myg.qx, myg.qy))
    dist = np.sqrt((x - xctr) ** 2 + (y - yctr) ** 2)
    return rp.get_param("heating.e_rate"), np.exp(-(dist / rp.get_param("heating.r_src")) ** 2)


def source_terms(myg, U, ivars, rp):
    """the reference's problem-source interface: S[i, j, n] with the heating term in the energy"""
    import torch
    rate, prof = heating(myg, rp)
    S = myg.scratch_array(nvar=ivars.nvar)
    S[:, :, ivars.iener] = U[:, :, ivars.idens].t() * rate * torch.from_numpy(prof).to(U.device)
    return S


def finalize():
    pass

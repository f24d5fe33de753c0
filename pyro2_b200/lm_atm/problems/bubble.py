"""A buoyant bubble in an isothermal, hydrostatic atmosphere (the low Mach counterpart of the compressible
bubble problem); same parameters as pyro/lm_atm/problems/bubble.py.  Sets the 2-d state and the 1-d base
state rho0(y), p0(y) (horizontal averages, p0 then re-integrated in hydrostatic equilibrium)."""
import numpy as np

from ...util import msg

DEFAULT_INPUTS = "inputs.bubble"

# stock run (the reference's inputs.bubble)
INPUTS = {"driver.max_steps": 2000, "driver.tmax": 1.0, "driver.cfl": 0.8, "io.basename": "lm_bubble_128_", "io.n_out": 1,
          "mesh.nx": 128, "mesh.ny": 128, "mesh.xmax": 1.0, "mesh.ymax": 1.0,
          "mesh.xlboundary": "periodic", "mesh.xrboundary": "periodic", "mesh.ylboundary": "reflect", "mesh.yrboundary": "outflow",
          "bubble.x_pert": 0.5, "bubble.y_pert": 0.5, "bubble.r_pert": 0.05}

PROBLEM_PARAMS = {"bubble.dens_base": 10.0,              # density at the base of the atmosphere
                  "bubble.scale_height": 2.0,            # scale height of the isothermal atmosphere
                  "bubble.x_pert": 2.0, "bubble.y_pert": 2.0, "bubble.r_pert": 0.25,
                  "bubble.pert_amplitude_factor": 5.0, "bubble.dens_cutoff": 0.01}


def init_data(my_data, base, rp):
    if rp.get_param("driver.verbose"):
        msg.bold("initializing the bubble problem...")
    g = my_data.grid
    grav = rp.get_param("lm-atmosphere.grav")
    gamma = rp.get_param("eos.gamma")
    scale_height, dens_base = rp.get_param("bubble.scale_height"), rp.get_param("bubble.dens_base")
    dens_cutoff = rp.get_param("bubble.dens_cutoff")
    x_pert, y_pert, r_pert = rp.get_param("bubble.x_pert"), rp.get_param("bubble.y_pert"), rp.get_param("bubble.r_pert")
    factor = rp.get_param("bubble.pert_amplitude_factor")

    # The fields are built for the whole domain (an x-slab of a decomposed run takes its rows at the end): the base
    # state is the mean over ALL rows, summed in the single-domain order.
    qx, ng = g.nx_global + 2 * g.ng, g.ng
    gx = 0.5 * (((np.arange(qx) - ng) * g.dx + g.xmin) + ((np.arange(qx) + 1.0 - ng) * g.dx + g.xmin))   # Grid2d's x, all rows
    dens = np.full((qx, g.qy), dens_cutoff)
    for j in range(g.jlo, g.jhi + 1):
        dens[:, j] = max(dens_base * np.exp(-g.y[j] / scale_height), dens_cutoff)
    cs2 = scale_height * abs(grav)
    pres = cs2 * dens                                    # isothermal: p = cs^2 rho
    eint = pres / (gamma - 1.0) / dens
    x = np.broadcast_to(gx[:, None], (qx, g.qy))
    y = np.broadcast_to(g.y[None, :], (qx, g.qy))
    idx = np.sqrt((x - x_pert) ** 2 + (y - y_pert) ** 2) <= r_pert
    eint[idx] = eint[idx] * factor                       # hotter at constant pressure -> lighter
    dens[idx] = pres[idx] / (eint[idx] * (gamma - 1.0))
    rows = slice(g.ioffset, g.ioffset + g.qx)
    my_data.get_var("density")[:, :] = dens[rows]
    my_data.get_var("x-velocity")[:, :] = 0.0
    my_data.get_var("y-velocity")[:, :] = 0.0
    my_data.get_var("eint")[:, :] = eint[rows]

    base["rho0"].d[:] = np.mean(dens, axis=0)
    base["p0"].d[:] = np.mean(pres, axis=0)
    for j in range(g.jlo + 1, g.jhi):                    # p0 again, from hydrostatic equilibrium with rho0
        base["p0"].d[j] = base["p0"].d[j - 1] + 0.5 * g.dy * (base["rho0"].d[j] + base["rho0"].d[j - 1]) * grav


def finalize():
    pass

"""Gresho vortex in the low-Mach form of Miczek, Roepke & Edelmann (2014): a rotating velocity field in
balance with its pressure gradient; same parameters as pyro/compressible/problems/gresho.py."""
import numpy as np

from ...util import msg

DEFAULT_INPUTS = "inputs.gresho"

# stock run (the reference's inputs.gresho)
INPUTS = {"driver.max_steps": 50000, "driver.tmax": 1, "driver.cfl": 0.8, "io.basename": "lm_gresho_128_",
          "io.dt_out": 0.2, "mesh.nx": 40, "mesh.ny": 40, "mesh.xmax": 1.0, "mesh.ymax": 1.0,
          "mesh.xlboundary": "periodic", "mesh.xrboundary": "periodic",
          "mesh.ylboundary": "periodic", "mesh.yrboundary": "periodic",
          "gresho.r": 0.2, "gresho.rho0": 1.0, "gresho.mach": 0.1, "gresho.t_r": 1.0, "compressible.grav": 0}

PROBLEM_PARAMS = {"gresho.rho0": 1.0,    # ambient density
                  "gresho.r": 0.2,       # radius of the velocity peak
                  "gresho.mach": 0.1,    # peak Mach number
                  "gresho.t_r": 1.0}     # reference time scale


def init_data(my_data, rp):
    if rp.get_param("driver.verbose"):
        msg.bold("initializing the Gresho vortex problem...")
    g = my_data.grid
    gamma = rp.get_param("eos.gamma")
    rho0, mach = rp.get_param("gresho.rho0"), rp.get_param("gresho.mach")
    rr, t_r = rp.get_param("gresho.r"), rp.get_param("gresho.t_r")
    # the reference takes the midpoint of the cell-centre array, ghost cells included (gresho.py:37); on an x-slab of a
    # decomposed run that array covers the slab only, so the first and last centre of the WHOLE domain are rebuilt with
    # the grid's own expressions (mesh/patch.py) -- same bits as g.x[0], g.x[-1] of the single domain
    ends = np.array([0, g.nx_global + 2 * g.ng - 1])
    x_ends = 0.5 * (((ends - g.ng) * g.dx + g.xmin) + ((ends + 1.0 - g.ng) * g.dx + g.xmin))
    x_center = 0.5 * (x_ends[0] + x_ends[1])
    y_center = 0.5 * (g.y[0] + g.y[-1])
    q_r = 0.4 * np.pi * (g.xmax - g.xmin) / t_r
    # p0 from the requested Mach number at the velocity peak u_phi = 5 rr, where p = p0 + 12.5 rr^2
    p0 = rho0 * q_r ** 2 * (5 * rr) ** 2 / (gamma * mach ** 2) - 12.5 * rr ** 2
    x = np.broadcast_to(g.x[:, None], (g.qx, g.qy))
    y = np.broadcast_to(g.y[None, :], (g.qx, g.qy))
    rad = np.sqrt((x - x_center) ** 2 + (y - y_center) ** 2)
    u_phi = np.zeros((g.qx, g.qy))
    pres = np.full((g.qx, g.qy), p0)
    core = rad < rr
    u_phi[core] = 5.0 * rad[core]
    pres[core] = p0 + 12.5 * rad[core] ** 2
    ring = np.logical_and(rad >= rr, rad < 2.0 * rr)
    u_phi[ring] = 2.0 - 5.0 * rad[ring]
    pres[ring] = p0 + 12.5 * rad[ring] ** 2 + 4.0 * (1.0 - 5.0 * rad[ring] - np.log(rr) + np.log(rad[ring]))
    outer = rad >= 2.0 * rr
    pres[outer] = p0 + 12.5 * (2.0 * rr) ** 2 + 4.0 * (1.0 - 5.0 * (2.0 * rr) - np.log(rr) + np.log(2.0 * rr))
    dens = np.full((g.qx, g.qy), rho0)
    xmom = -dens * q_r * u_phi * (y - y_center) / rad
    ymom = dens * q_r * u_phi * (x - x_center) / rad
    my_data.get_var("density")[:, :] = dens
    my_data.get_var("x-momentum")[:, :] = xmom
    my_data.get_var("y-momentum")[:, :] = ymom
    my_data.get_var("energy")[:, :] = pres / (gamma - 1.0) + 0.5 * (xmom ** 2 + ymom ** 2) / dens
    if rp.get_param("driver.verbose"):
        cs = np.sqrt(gamma * pres / dens)
        print(f"peak Mach number = {np.abs(q_r * u_phi).max() / cs.max()}")


def finalize():
    pass

"""Acoustic pulse of McCorquodale & Colella (2011): uniform state plus a smooth density / pressure bump
that launches a low-Mach sound wave; same parameters as pyro/compressible/problems/acoustic_pulse.py."""
import numpy as np

from ...util import msg

DEFAULT_INPUTS = "inputs.acoustic_pulse"

# stock run (the reference's inputs.acoustic_pulse)
INPUTS = {"driver.max_steps": 5000, "driver.tmax": 0.24, "driver.fix_dt": 1.5e-3, "compressible.cvisc": 0.1,
          "io.basename": "acoustic_pulse_", "io.dt_out": 0.03, "eos.gamma": 1.4, "mesh.nx": 128, "mesh.ny": 128,
          "mesh.xmax": 1.0, "mesh.ymax": 1.0, "mesh.xlboundary": "periodic", "mesh.xrboundary": "periodic",
          "mesh.ylboundary": "periodic", "mesh.yrboundary": "periodic",
          "acoustic_pulse.rho0": 1.4, "acoustic_pulse.drho0": 0.14}

PROBLEM_PARAMS = {"acoustic_pulse.rho0": 1.4, "acoustic_pulse.drho0": 0.14}


def init_data(myd, rp):
    if rp.get_param("driver.verbose"):
        msg.bold("initializing the acoustic pulse problem...")
    g = myd.grid
    gamma = rp.get_param("eos.gamma")
    rho0, drho0 = rp.get_param("acoustic_pulse.rho0"), rp.get_param("acoustic_pulse.drho0")
    xctr = 0.5 * (rp.get_param("mesh.xmin") + rp.get_param("mesh.xmax"))
    yctr = 0.5 * (rp.get_param("mesh.ymin") + rp.get_param("mesh.ymax"))
    x = np.broadcast_to(g.x[:, None], (g.qx, g.qy))
    y = np.broadcast_to(g.y[None, :], (g.qx, g.qy))
    dist = np.sqrt((x - xctr) ** 2 + (y - yctr) ** 2)
    dens = np.full((g.qx, g.qy), rho0)
    inside = dist <= 0.5
    dens[inside] = rho0 + drho0 * np.exp(-16 * dist[inside] ** 2) * np.cos(np.pi * dist[inside]) ** 6
    p = (dens / rho0) ** gamma
    myd.get_var("density")[:, :] = dens
    myd.get_var("x-momentum")[:, :] = 0.0
    myd.get_var("y-momentum")[:, :] = 0.0
    myd.get_var("energy")[:, :] = p / (gamma - 1)


def finalize():
    pass

"""Smooth advection test: a Gaussian density bump carried diagonally through a periodic box at constant
pressure (Cartesian branch of pyro/compressible/problems/advect.py)."""
import numpy as np

from ...util import msg

DEFAULT_INPUTS = "inputs.advect.64"

# stock run (the reference's inputs.advect.64)
INPUTS = {"driver.max_steps": 500, "driver.tmax": 1.0, "driver.init_tstep_factor": 1.0, "driver.fix_dt": 0.005,
          "compressible.limiter": 0, "compressible.cvisc": 0.1, "io.basename": "advect_64_", "eos.gamma": 1.4,
          "mesh.nx": 64, "mesh.ny": 64, "mesh.xmax": 1.0, "mesh.ymax": 1.0,
          "mesh.xlboundary": "periodic", "mesh.xrboundary": "periodic",
          "mesh.ylboundary": "periodic", "mesh.yrboundary": "periodic"}

PROBLEM_PARAMS = {}


def init_data(my_data, rp):
    if rp.get_param("driver.verbose"):
        msg.bold("initializing the advect problem...")
    g = my_data.grid
    gamma = rp.get_param("eos.gamma")
    xctr = 0.5 * (rp.get_param("mesh.xmin") + rp.get_param("mesh.xmax"))
    yctr = 0.5 * (rp.get_param("mesh.ymin") + rp.get_param("mesh.ymax"))
    x = np.broadcast_to(g.x[:, None], (g.qx, g.qy))
    y = np.broadcast_to(g.y[None, :], (g.qx, g.qy))
    dens = 1.0 + np.exp(-60.0 * ((x - xctr) ** 2 + (y - yctr) ** 2))
    u = v = 1.0                      # diagonal flow
    if getattr(g, "coord_type", 0) == 1:
        # SphericalPolar (advect.py:57-72): the blob sits at mid radius on the bisector of the theta range (in the
        # Cartesian coordinates r sin(theta), r cos(theta)) and moves along theta
        xmin, xmax = rp.get_param("mesh.xmin"), rp.get_param("mesh.xmax")
        ymin, ymax = rp.get_param("mesh.ymin"), rp.get_param("mesh.ymax")
        xc = 0.5 * (xmin + xmax) * np.sin((ymin + ymax) * 0.25)
        yc = 0.5 * (xmin + xmax) * np.cos((ymin + ymax) * 0.25)
        xx, yy = x * np.sin(y), x * np.cos(y)
        dens = 1.0 + np.exp(-120.0 * ((xx - xc) ** 2 + (yy - yc) ** 2))
        u, v = 0.0, 1.0
    xmom, ymom = dens * u, dens * v
    p = 1.0
    my_data.get_var("density")[:, :] = dens
    my_data.get_var("x-momentum")[:, :] = xmom
    my_data.get_var("y-momentum")[:, :] = ymom
    my_data.get_var("energy")[:, :] = p / (gamma - 1.0) + 0.5 * (xmom ** 2 + ymom ** 2) / dens


def finalize():
    pass

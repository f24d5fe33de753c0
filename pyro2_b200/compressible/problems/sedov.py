"""Sedov blast wave: a fixed energy deposited in a small disc at the domain centre, ambient gas at
rest.  Same initial state, parameter names and defaults as pyro/compressible/problems/sedov.py
(:15-93, Cartesian branch); the state is assembled on the host with numpy and uploaded once."""
import math

import numpy as np

from ...util import msg

DEFAULT_INPUTS = "inputs.sedov"

# stock run: 128^2 outflow box, blast radius 0.01 (what the reference's inputs.sedov sets)
INPUTS = {"driver.max_steps": 5000, "driver.tmax": 0.1, "compressible.limiter": 2, "compressible.cvisc": 0.1,
          "io.basename": "sedov_unsplit_", "io.dt_out": 0.0125, "eos.gamma": 1.4,
          "mesh.nx": 128, "mesh.ny": 128, "mesh.xmax": 1.0, "mesh.ymax": 1.0,
          "mesh.xlboundary": "outflow", "mesh.xrboundary": "outflow",
          "mesh.ylboundary": "outflow", "mesh.yrboundary": "outflow", "sedov.r_init": 0.01}

PROBLEM_PARAMS = {"sedov.r_init": 0.1,   # radius of the initial energy deposit
                  "sedov.nsub": 4}       # sub-samples per direction in partially covered zones


def init_data(my_data, rp):
    if rp.get_param("driver.verbose"):
        msg.bold("initializing the sedov problem...")
    g = my_data.grid
    gamma = rp.get_param("eos.gamma")
    r_init = rp.get_param("sedov.r_init")
    nsub = rp.get_param("sedov.nsub")
    xctr = 0.5 * (rp.get_param("mesh.xmin") + rp.get_param("mesh.xmax"))
    yctr = 0.5 * (rp.get_param("mesh.ymin") + rp.get_param("mesh.ymax"))
    E_sedov = 1.0
    p_ambient = 1.e-5

    if getattr(g, "coord_type", 0) == 1:
        # SphericalPolar (sedov.py:84-93): a large energy inside r < r_init, nothing sub-sampled
        ener = np.full((g.qx, g.qy), 1.e-6 / (gamma - 1.0))
        ener[np.broadcast_to(g.x[:, None], (g.qx, g.qy)) < r_init] = 1.e6
        my_data.get_var("density")[:, :] = 1.0
        my_data.get_var("x-momentum")[:, :] = 0.0
        my_data.get_var("y-momentum")[:, :] = 0.0
        my_data.get_var("energy")[:, :] = ener
        return

    ener = np.full((g.qx, g.qy), p_ambient / (gamma - 1.0))

    # zones whose centre is within 2 r_init of the centre get an area-weighted pressure from
    # nsub x nsub sub-samples
    dist = np.sqrt((g.x[:, None] - xctr) ** 2 + (g.y[None, :] - yctr) ** 2)
    ii, jj = np.nonzero(dist < 2.0 * r_init)
    if len(ii) > 0:
        off = (np.arange(nsub) + 0.5)
        xsub = g.xl[ii][:, None] + (g.dx / nsub) * off[None, :]            # (ncell, nsub)
        ysub = g.yl[jj][:, None] + (g.dy / nsub) * off[None, :]
        d = np.sqrt((xsub[:, :, None] - xctr) ** 2 + (ysub[:, None, :] - yctr) ** 2)
        n_in = np.count_nonzero(d <= r_init, axis=(1, 2))
        p = n_in * (gamma - 1.0) * E_sedov / (math.pi * r_init * r_init) + (nsub * nsub - n_in) * 1.e-5
        p = p / (nsub * nsub)
        ener[ii, jj] = p / (gamma - 1.0)

    my_data.get_var("density")[:, :] = 1.0
    my_data.get_var("x-momentum")[:, :] = 0.0
    my_data.get_var("y-momentum")[:, :] = 0.0
    my_data.get_var("energy")[:, :] = ener


def finalize():
    print("\n          compare the radial profile with the exact cylindrical Sedov solution\n")

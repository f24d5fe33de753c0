"""A Gaussian profile, which stays Gaussian under constant-conductivity diffusion (peak dropping, width
growing): usable for verification.  Same parameters as pyro/diffusion/problems/gaussian.py."""
import numpy as np

from ...util import msg

DEFAULT_INPUTS = "inputs.gaussian"

# stock run (the reference's inputs.gaussian)
INPUTS = {"driver.max_steps": 500, "driver.tmax": 0.02, "driver.max_dt_change": 1.e33, "driver.init_tstep_factor": 1.0,
          "driver.cfl": 2.0, "io.basename": "gaussian_", "io.dt_out": 0.005, "mesh.nx": 128, "mesh.ny": 128,
          "mesh.xmax": 1.0, "mesh.ymax": 1.0, "mesh.xlboundary": "neumann", "mesh.xrboundary": "neumann",
          "mesh.ylboundary": "neumann", "mesh.yrboundary": "neumann", "diffusion.k": 1.0, "gaussian.t_0": 0.0001}

PROBLEM_PARAMS = {"gaussian.t_0": 0.001, "gaussian.phi_0": 1.0, "gaussian.phi_max": 2.0}


def phi_analytic(dist, t, t_0, k, phi_1, phi_2):
    """the exact solution at time t for the profile that is a Gaussian of amplitude phi_2 - phi_1 at t = 0"""
    return (phi_2 - phi_1) * (t_0 / (t + t_0)) * np.exp(-0.25 * dist ** 2 / (k * (t + t_0))) + phi_1


def init_data(my_data, rp):
    if rp.get_param("driver.verbose"):
        msg.bold("initializing the Gaussian diffusion problem...")
    g = my_data.grid
    xctr, yctr = 0.5 * (g.xmin + g.xmax), 0.5 * (g.ymin + g.ymax)
    k = rp.get_param("diffusion.k")
    t_0, phi_max, phi_0 = rp.get_param("gaussian.t_0"), rp.get_param("gaussian.phi_max"), rp.get_param("gaussian.phi_0")
    x = np.broadcast_to(g.x[:, None], (g.qx, g.qy))
    y = np.broadcast_to(g.y[None, :], (g.qx, g.qy))
    dist = np.sqrt((x - xctr) ** 2 + (y - yctr) ** 2)
    my_data.get_var("phi")[:, :] = phi_analytic(dist, 0.0, t_0, k, phi_0, phi_max)
    # kept for later analysis, as in the reference
    for key, val in (("k", k), ("t_0", t_0), ("phi_0", phi_0), ("phi_max", phi_max)):
        my_data.set_aux(key, val)


def finalize():
    print("""
          The solution can be compared to the analytic solution (phi_analytic in this module)
          """)

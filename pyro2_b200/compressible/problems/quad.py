"""Four-quadrant 2-d Riemann problem (Schulz-Rinne et al. 1993 configuration 3); same parameters
as pyro/compressible/problems/quad.py."""
import numpy as np

from ...util import msg

DEFAULT_INPUTS = "inputs.quad"

# stock run: 256^2, corner at (0.8, 0.8) (the reference's inputs.quad)
INPUTS = {"driver.max_steps": 1000, "driver.tmax": 0.8, "compressible.limiter": 2, "compressible.cvisc": 0.1,
          "io.basename": "quad_unsplit_", "io.dt_out": 0.1, "mesh.nx": 256, "mesh.ny": 256,
          "mesh.xmax": 1.0, "mesh.ymax": 1.0, "mesh.xlboundary": "outflow", "mesh.xrboundary": "outflow",
          "mesh.ylboundary": "outflow", "mesh.yrboundary": "outflow",
          "quadrant.rho1": 1.4, "quadrant.u1": 0.0, "quadrant.v1": 0.0, "quadrant.p1": 1.5,
          "quadrant.rho2": 0.532258064516129, "quadrant.u2": 1.206045378311055, "quadrant.v2": 0.0, "quadrant.p2": 0.3,
          "quadrant.rho3": 0.137992831541219, "quadrant.u3": 1.206045378311055, "quadrant.v3": 1.206045378311055,
          "quadrant.p3": 0.029032258064516,
          "quadrant.rho4": 0.532258064516129, "quadrant.u4": 0.0, "quadrant.v4": 1.206045378311055, "quadrant.p4": 0.3,
          "quadrant.cx": 0.8, "quadrant.cy": 0.8}

PROBLEM_PARAMS = {"quadrant.rho1": 1.5, "quadrant.u1": 0.0, "quadrant.v1": 0.0, "quadrant.p1": 1.5,
                  "quadrant.rho2": 0.532258064516129, "quadrant.u2": 1.206045378311055,
                  "quadrant.v2": 0.0, "quadrant.p2": 0.3,
                  "quadrant.rho3": 0.137992831541219, "quadrant.u3": 1.206045378311055,
                  "quadrant.v3": 1.206045378311055, "quadrant.p3": 0.029032258064516,
                  "quadrant.rho4": 0.532258064516129, "quadrant.u4": 0.0,
                  "quadrant.v4": 1.206045378311055, "quadrant.p4": 0.3,
                  "quadrant.cx": 0.5, "quadrant.cy": 0.5}


def init_data(my_data, rp):
    if rp.get_param("driver.verbose"):
        msg.bold("initializing the quadrant problem...")
    g = my_data.grid
    gamma = rp.get_param("eos.gamma")
    cx, cy = rp.get_param("quadrant.cx"), rp.get_param("quadrant.cy")
    right = np.broadcast_to(g.x[:, None] >= cx, (g.qx, g.qy))
    top = np.broadcast_to(g.y[None, :] >= cy, (g.qx, g.qy))
    masks = {1: right & top, 2: ~right & top, 3: ~right & ~top, 4: right & ~top}
    dens = np.zeros((g.qx, g.qy))
    xmom, ymom, ener = dens.copy(), dens.copy(), dens.copy()
    for k, m in masks.items():
        r, u, v, p = (rp.get_param(f"quadrant.{n}{k}") for n in ("rho", "u", "v", "p"))
        dens[m] = r
        xmom[m] = r * u
        ymom[m] = r * v
        ener[m] = p / (gamma - 1.0) + 0.5 * r * (u * u + v * v)
    my_data.get_var("density")[:, :] = dens
    my_data.get_var("x-momentum")[:, :] = xmom
    my_data.get_var("y-momentum")[:, :] = ymom
    my_data.get_var("energy")[:, :] = ener


def finalize():
    pass

"""Double Mach reflection (Woodward & Colella 1984): a Mach-10 shock, inclined at 60 degrees to the x axis,
meets a reflecting wall that starts at x = 1/6.  Same parameters as pyro/compressible/problems/ramp.py; run with
the "ramp" boundaries (left, bottom, top) and outflow on the right."""
import math

import numpy as np

from ...util import msg

DEFAULT_INPUTS = "inputs.ramp"

# stock run (the reference's inputs.ramp)
INPUTS = {"driver.max_steps": 12500, "driver.tmax": 0.25, "compressible.limiter": 2, "compressible.cvisc": 0.1,
          "io.basename": "double_mach_reflection_", "io.dt_out": 0.1,
          "mesh.nx": 1024, "mesh.ny": 256, "mesh.xmax": 4.0, "mesh.ymax": 1.0,
          "mesh.xlboundary": "ramp", "mesh.xrboundary": "outflow", "mesh.ylboundary": "ramp", "mesh.yrboundary": "ramp",
          "ramp.rhol": 8.0, "ramp.ul": 7.1447096, "ramp.vl": -4.125, "ramp.pl": 116.5,
          "ramp.rhor": 1.4, "ramp.ur": 0.0, "ramp.vr": 0.0, "ramp.pr": 1.0}

PROBLEM_PARAMS = {"ramp.rhol": 8.0, "ramp.ul": 7.1447096, "ramp.vl": -4.125, "ramp.pl": 116.5,    # post-shock state
                  "ramp.rhor": 1.4, "ramp.ur": 0.0, "ramp.vr": 0.0, "ramp.pr": 1.0}               # pre-shock state


def init_data(my_data, rp):
    if rp.get_param("driver.verbose"):
        msg.bold("initializing the double Mach reflection problem...")
    g = my_data.grid
    gamma = rp.get_param("eos.gamma")
    r_l, u_l, v_l, p_l = (rp.get_param(f"ramp.{k}") for k in ("rhol", "ul", "vl", "pl"))
    r_r, u_r, v_r, p_r = (rp.get_param(f"ramp.{k}") for k in ("rhor", "ur", "vr", "pr"))
    energy_l = p_l / (gamma - 1.0) + 0.5 * r_l * (u_l * u_l + v_l * v_l)
    energy_r = p_r / (gamma - 1.0) + 0.5 * r_r * (u_r * u_r + v_r * v_r)
    post = (0.25 * r_l, 0.25 * r_l * u_l, 0.25 * r_l * v_l, 0.25 * energy_l)
    pre = (0.25 * r_r, 0.25 * r_r * u_r, 0.25 * r_r * v_r, 0.25 * energy_r)
    # 2 x 2 supersampling of each valid cell about the initial shock y = tan(60 deg) (x - 1/6): two sample
    # heights, two front positions (through the cell's left / right sample abscissa); four quarter-weights added
    # in the order (low, left), (low, right), (high, left), (high, right)
    x = g.x[g.ilo:g.ihi + 1, None]
    y = g.y[None, g.jlo:g.jhi + 1]
    half_x, half_y = 0.5 * g.dx * math.sqrt(3), 0.5 * g.dy * math.sqrt(3)
    fronts = [math.tan(math.pi / 3.0) * (x + s * half_x - 1.0 / 6.0) for s in (-1.0, 1.0)]
    fields = [np.zeros((g.nx, g.ny)) for _ in range(4)]
    for cy in (y - half_y, y + half_y):
        for sf in fronts:
            behind = cy >= sf
            for k in range(4):
                fields[k] = fields[k] + np.where(behind, post[k], pre[k])
    # ghost cells: density 1.4, the rest empty, until the first boundary fill
    full = [np.full((g.qx, g.qy), 1.4)] + [np.zeros((g.qx, g.qy)) for _ in range(3)]
    for k in range(4):
        full[k][g.ilo:g.ihi + 1, g.jlo:g.jhi + 1] = fields[k]
    for name, a in zip(("density", "x-momentum", "y-momentum", "energy"), full):
        my_data.get_var(name)[:, :] = a


def finalize():
    pass

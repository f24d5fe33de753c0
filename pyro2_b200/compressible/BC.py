"""Compressible-specific boundary conditions: the mirror of pyro/compressible/BC.py (user :21-258).

"ambient" (upper y boundary, BC.py:142-168): the ghost rows hold the constant ambient state the problem
registered with set_aux (ambient_rho, ambient_u, ambient_v, ambient_p).

"hse" (lower / upper y boundary): density and momenta get a zero-gradient copy of the first interior
row, the energy is integrated outward in hydrostatic equilibrium at that row's density
(BC.py:45-148) -- one small CUDA kernel per variable and side (csrc/bc_user.cu), called from
CellCenterData2d.fill_BC after the standard fill exactly like the reference's ext_bcs hook, so the
ghost cells are bit-identical.

"ramp" (double Mach reflection, BC.py:183-256): post-shock inflow on the left, inflow / reflecting wall on the
bottom split at the ramp's foot x = 1/6, and on top the post- / pre-shock states about the position of the
Mach-10 shock at the current time, 2 x 2 supersampled.  The values are two constants per variable; the ghost strips
are written with device slice assignments, the time-dependent top rows (ng rows of qx numbers per variable) are
evaluated on the host and copied, once per fill."""
import math

import numpy as np
import torch

from .. import ops
from ..util import msg

_VARS = ("density", "energy", "x-momentum", "y-momentum")


def _ambient_value(variable, ccdata):
    rho, u, v, p = (ccdata.get_aux(k) for k in ("ambient_rho", "ambient_u", "ambient_v", "ambient_p"))
    if variable == "density":
        return rho
    if variable == "x-momentum":
        return rho * u
    if variable == "y-momentum":
        return rho * v
    ke = 0.5 * rho * (u ** 2 + v ** 2)
    return p / (ccdata.get_aux("gamma") - 1.0) + ke


def inflow_post_bc(var, g):
    """conserved post-shock inflow state of the double Mach reflection problem (BC.py:259-277)"""
    r_l, u_l, v_l, p_l = 8.0, 7.1447096, -4.125, 116.5
    return {"density": r_l, "x-momentum": r_l * u_l, "y-momentum": r_l * v_l,
            "energy": p_l / (g - 1.0) + 0.5 * r_l * (u_l * u_l + v_l * v_l)}.get(var, 0.0)


def inflow_pre_bc(var, g):
    """conserved pre-shock state (BC.py:280-296)"""
    r_r, u_r, v_r, p_r = 1.4, 0.0, 0.0, 1.0
    return {"density": r_r, "x-momentum": r_r * u_r, "y-momentum": r_r * v_r,
            "energy": p_r / (g - 1.0) + 0.5 * r_r * (u_r * u_r + v_r * v_r)}.get(var, 0.0)


def _ramp(bc_edge, variable, ccdata):
    g = ccdata.grid
    gamma = ccdata.get_aux("gamma")
    v = ccdata.get_var(variable).t()
    post, pre = inflow_post_bc(variable, gamma), inflow_pre_bc(variable, gamma)
    if bc_edge == "xlb":
        v[:g.ilo, :] = post
    elif bc_edge == "ylb":
        k = int(np.count_nonzero(g.x < 1.0 / 6.0))          # columns left of the ramp's foot (x is increasing)
        sign = -1.0 if variable == "y-momentum" else 1.0
        for jj in range(g.ng):
            j = g.jlo - 1 - jj
            v[:k, j] = post
            v[k:, j] = sign * v[k:, g.jlo + jj]
    elif bc_edge == "yrb":
        half = 0.5 * g.dx * math.sqrt(3)
        rows = np.zeros((g.qx, g.ng))
        for n, j in enumerate(range(g.jhi + 1, g.jhi + g.ng + 1)):
            # where the shock crosses this row now, through the lower and the upper sample point
            fronts = [1.0 / 6.0 + (g.y[j] + s * 0.5 * g.dy * math.sqrt(3)) / math.tan(math.pi / 3.0)
                      + (10.0 / math.sin(math.pi / 3.0)) * ccdata.t for s in (-1.0, 1.0)]
            acc = np.zeros(g.qx)
            for sf in fronts:
                for cx in (g.x - half, g.x + half):
                    acc = acc + np.where(cx < sf, 0.25 * post, 0.25 * pre)
            rows[:, n] = acc
        v[:, g.jhi + 1:g.jhi + g.ng + 1] = torch.from_numpy(rows).to(v.device)
    else:
        msg.fail("error: ramp BC not supported for xrb")
    ccdata.version += 1


def user(bc_name, bc_edge, variable, ccdata, ivars=None):   # pylint: disable=unused-argument
    if bc_name not in ("hse", "ambient", "ramp"):
        msg.fail(f"ERROR: bc type {bc_name} not supported")
    if variable not in _VARS or tuple(ccdata.names[:4]) != _VARS:
        raise NotImplementedError("variable not defined")
    g = ccdata.grid
    if bc_name == "ramp":
        _ramp(bc_edge, variable, ccdata)
        return
    if bc_name == "ambient":
        if bc_edge != "yrb":
            msg.fail("error: ambient BC not supported for xlb, xrb, or ylb")
        ops.fill_ambient(ccdata.planes, g.nx, g.ny, g.ng, _VARS.index(variable), 1, _ambient_value(variable, ccdata))
        ccdata.version += 1
        return
    if bc_edge not in ("ylb", "yrb"):
        msg.fail("error: hse BC not supported for xlb or xrb")
    ops.fill_hse(ccdata.planes, g.nx, g.ny, g.ng, g.dy, ccdata.get_aux("grav"), ccdata.get_aux("gamma"),
                 _VARS.index(variable), 0 if bc_edge == "ylb" else 1)
    ccdata.version += 1

"""Compressible-specific boundary conditions: the mirror of pyro/compressible/BC.py (user :21-258).

"ambient" (upper y boundary, BC.py:142-168): the ghost rows hold the constant ambient state the problem
registered with set_aux (ambient_rho, ambient_u, ambient_v, ambient_p).

"hse" (lower / upper y boundary): density and momenta get a zero-gradient copy of the first interior
row, the energy is integrated outward in hydrostatic equilibrium at that row's density
(BC.py:45-148) -- one small CUDA kernel per variable and side (csrc/bc_user.cu), called from
CellCenterData2d.fill_BC after the standard fill exactly like the reference's ext_bcs hook, so the
ghost cells are bit-identical.  "ramp" is not built."""
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


def user(bc_name, bc_edge, variable, ccdata, ivars=None):   # pylint: disable=unused-argument
    if bc_name not in ("hse", "ambient"):
        msg.fail(f"ERROR: the device build implements the hse and ambient boundaries (got {bc_name})")
    if variable not in _VARS or tuple(ccdata.names[:4]) != _VARS:
        raise NotImplementedError("variable not defined")
    g = ccdata.grid
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

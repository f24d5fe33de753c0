"""Compressible-specific boundary conditions: the mirror of pyro/compressible/BC.py (user :21-258).

"hse" (lower / upper y boundary): density and momenta get a zero-gradient copy of the first interior
row, the energy is integrated outward in hydrostatic equilibrium at that row's density
(BC.py:45-148) -- one small CUDA kernel per variable and side (csrc/bc_user.cu), called from
CellCenterData2d.fill_BC after the standard fill exactly like the reference's ext_bcs hook, so the
ghost cells are bit-identical.  "ambient" and "ramp" are not built."""
from .. import ops
from ..util import msg

_VARS = ("density", "energy", "x-momentum", "y-momentum")


def user(bc_name, bc_edge, variable, ccdata, ivars=None):   # pylint: disable=unused-argument
    if bc_name != "hse":
        msg.fail(f"ERROR: the device build implements the hse boundary only (got {bc_name})")
    if bc_edge not in ("ylb", "yrb"):
        msg.fail("error: hse BC not supported for xlb or xrb")
    if variable not in _VARS or tuple(ccdata.names[:4]) != _VARS:
        raise NotImplementedError("variable not defined")
    g = ccdata.grid
    ops.fill_hse(ccdata.planes, g.nx, g.ny, g.ng, g.dy, ccdata.get_aux("grav"), ccdata.get_aux("gamma"),
                 _VARS.index(variable), 0 if bc_edge == "ylb" else 1)
    ccdata.version += 1

"""Boundary-condition descriptors: the data-only mirror of pyro/mesh/boundary.py (BC :64-204,
define_bc :22, bc_is_solid :53).  The actual ghost fill is a CUDA kernel (csrc/ghost_cfl.cu)."""
from ..util import msg

# is this boundary type a solid wall? (boundary.py:10-17)
bc_solid = {"outflow": False, "periodic": False, "reflect": True, "reflect-even": True,
            "reflect-odd": True, "dirichlet": True, "neumann": False}

# user-registered ghost-fill callbacks, keyed by BC name (boundary.py:19, 22-33)
ext_bcs = {}


def define_bc(bc_type, function, is_solid=False):
    """register a solver-specific boundary type filled by a Python callback"""
    bc_solid[bc_type] = is_solid
    ext_bcs[bc_type] = function


class BCProp:
    """one flag per boundary (boundary.py:42-50)"""

    def __init__(self, xl_prop, xr_prop, yl_prop, yr_prop):
        self.xl, self.xr, self.yl, self.yr = xl_prop, xr_prop, yl_prop, yr_prop


def bc_is_solid(bc):
    return BCProp(int(bc_solid[bc.xlb]), int(bc_solid[bc.xrb]), int(bc_solid[bc.ylb]), int(bc_solid[bc.yrb]))


class BC:
    """boundary types of one variable on the four faces, plus optional inhomogeneous
    Dirichlet / Neumann boundary values evaluated on the physical edge (boundary.py:64-204)"""

    def __init__(self, *, xlb="outflow", xrb="outflow", ylb="outflow", yrb="outflow",
                 xl_func=None, xr_func=None, yl_func=None, yr_func=None, grid=None, odd_reflect_dir=""):
        def resolve(name, label, direction):
            if name not in bc_solid:
                msg.fail(f"ERROR: {label} = {name} invalid BC")
            if name == "reflect":
                return "reflect-odd" if odd_reflect_dir == direction else "reflect-even"
            return name

        self.xlb = resolve(xlb, "xlb", "x")
        self.xrb = resolve(xrb, "xrb", "x")
        self.ylb = resolve(ylb, "ylb", "y")
        self.yrb = resolve(yrb, "yrb", "y")

        if (xlb == "periodic") != (xrb == "periodic"):
            msg.fail("ERROR: both xlb and xrb must be periodic")
        if (ylb == "periodic") != (yrb == "periodic"):
            msg.fail("ERROR: both ylb and yrb must be periodic")

        self.xl_value = xl_func(grid.y) if xl_func is not None else None
        self.xr_value = xr_func(grid.y) if xr_func is not None else None
        self.yl_value = yl_func(grid.x) if yl_func is not None else None
        self.yr_value = yr_func(grid.x) if yr_func is not None else None

    def names(self):
        return (self.xlb, self.xrb, self.ylb, self.yrb)

    def __str__(self):
        return f"BCs: -x: {self.xlb}  +x: {self.xrb}  -y: {self.ylb}  +y: {self.yrb}"

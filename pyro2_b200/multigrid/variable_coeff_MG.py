r"""Variable-coefficient cell-centred multigrid, :math:`\nabla\cdot(\eta\nabla\phi) = f`, on the
B200 -- the interface of pyro/multigrid/variable_coeff_MG.py (VarCoeffCCMG2d :24-213).

    a = VarCoeffCCMG2d(nx, ny, xl_BC_type="dirichlet", ..., coeffs=c, coeffs_bc=bc_c)
    a.init_zeros(); a.init_RHS(f(a.x2d, a.y2d)); a.solve(rtol=1.e-11)

``coeffs`` is eta on the solution grid (an ArrayIndexer with one ghost cell, a tensor or an ndarray of
shape (nx+2, ny+2); only the valid cells are read) and ``coeffs_bc`` its boundary conditions.  As in the
reference, eta is restricted to every level and ghost-filled, averaged onto the edges of the finest
level and the edge values are restricted down the hierarchy once, in the constructor
(variable_coeff_MG.py:57-109) -- here by CUDA kernels behind ``p2b_mg_set_coeffs``.  smooth(),
_compute_residual(), v_cycle() and solve() are inherited: the library applies the variable-coefficient
stencils (:112-213) once coefficients are set, with the reference's operation order, so the solution is
bit-identical to the reference's.

Differences from the reference: the per-level eta is exposed as ``coeffs[level]`` (and
``edge_coeffs[level].x / .y``) rather than as an auxiliary variable "coeffs" of ``grids[level]``.
"""
import numpy as np
import torch

from ..mesh.array_indexer import ArrayIndexer
from . import MG
from .edge_coeffs import EdgeCoeffs


class VarCoeffCCMG2d(MG.CellCenterMG2d):
    def __init__(self, nx, ny, xmin=0.0, xmax=1.0, ymin=0.0, ymax=1.0,
                 xl_BC_type="dirichlet", xr_BC_type="dirichlet",
                 yl_BC_type="dirichlet", yr_BC_type="dirichlet",
                 nsmooth=10, nsmooth_bottom=50,
                 verbose=0,
                 coeffs=None, coeffs_bc=None,
                 true_function=None, vis=0, vis_title=""):
        if coeffs is None or coeffs_bc is None:
            raise ValueError("ERROR: the variable-coefficient solver needs coeffs and coeffs_bc")
        super().__init__(nx, ny, ng=1, xmin=xmin, xmax=xmax, ymin=ymin, ymax=ymax,
                         xl_BC_type=xl_BC_type, xr_BC_type=xr_BC_type,
                         yl_BC_type=yl_BC_type, yr_BC_type=yr_BC_type,
                         alpha=0.0, beta=0.0, nsmooth=nsmooth, nsmooth_bottom=nsmooth_bottom,
                         verbose=verbose, true_function=true_function, vis=vis, vis_title=vis_title)
        self.coeffs_bc = coeffs_bc
        self._nxy = (nx, ny)
        self.set_coeffs(coeffs)
        # coarsest first, like grids[] (variable_coeff_MG.py:44-46, 78-90)
        self.coeffs = []
        self.edge_coeffs = []
        for level in range(self.nlevels):
            lg = self.grids[level].grid
            self.coeffs.append(ArrayIndexer(self._h.coeff_plane(level, "c"), grid=lg))
            self.edge_coeffs.append(EdgeCoeffs(lg, self._h.coeff_plane(level, "x"), self._h.coeff_plane(level, "y")))

    def set_coeffs(self, coeffs):
        """(re)build every level's eta and edge coefficients from eta on the solution grid.  The constructor
        calls this once, as the reference does; calling it again on an existing solver replaces what would be a
        new VarCoeffCCMG2d with other coefficients (the lm_atm solver builds three per step) -- workspace, views
        and a captured V-cycle graph stay valid because the planes are rewritten in place."""
        nx, ny = self._nxy
        g = getattr(coeffs, "g", None)
        if g is not None and (g.nx != nx or g.ny != ny):
            raise IndexError("coefficient array not the same size as multigrid problem")
        if isinstance(coeffs, torch.Tensor):
            c = coeffs.as_subclass(torch.Tensor).to(device="cuda", dtype=torch.float64)
        else:
            c = torch.as_tensor(np.asarray(coeffs, dtype=np.float64)).cuda()
        if tuple(c.shape) != (nx + 2, ny + 2):
            raise IndexError("coefficient array not the same size as multigrid problem")
        if c.stride(1) != 1:
            c = c.contiguous()
        bc = self.coeffs_bc
        self._h.set_coeffs(c, (bc.xlb, bc.xrb, bc.ylb, bc.yrb))

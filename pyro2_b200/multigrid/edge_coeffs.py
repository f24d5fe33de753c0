"""Edge-centred coefficients of the variable-coefficient multigrid solver -- the container of
pyro/multigrid/edge_coeffs.py:1-54.  On the B200 the arrays are built by the library
(p2b_mg_set_coeffs: the finest level from the cell-centred eta, coarser levels by restriction); this
class only gives them the reference's names: ``x[i, j]`` is eta_{i-1/2, j} / dx**2 and ``y[i, j]`` is
eta_{i, j-1/2} / dy**2."""
from ..mesh.array_indexer import ArrayIndexer


class EdgeCoeffs:
    def __init__(self, g, x, y):
        self.grid = g
        self.x = ArrayIndexer(x, grid=g)
        self.y = ArrayIndexer(y, grid=g)

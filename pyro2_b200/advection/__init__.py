"""Linear advection (second-order unsplit Godunov, Colella 1990) on the B200 -- the interface of
pyro/advection; the reference's own plumbing test case (BASELINE config 1: smooth 64 x 64)."""
__all__ = ["simulation"]

from .simulation import Simulation   # noqa: F401

"""Incompressible flow (second-order approximate projection) on the B200 -- the interface of
pyro/incompressible.  One step = limited slopes and interface states -> MAC velocities -> MAC
projection (multigrid) -> upwinded interface states -> provisional velocity -> final projection
(multigrid); the explicit stages are the p2b_flow_* kernels (csrc/flow.cu), the projections the
p2b_mg_* V-cycles (csrc/mg.cu)."""
__all__ = ["simulation"]

from .simulation import Simulation   # noqa: F401

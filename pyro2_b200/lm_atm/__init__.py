"""Low Mach number atmospheric flow on the B200 -- the interface of pyro/lm_atm: a pseudo-incompressible
projection method on a stratified base state (divergence constraint D(beta0 U) = 0).  The explicit stages are
the p2b_lm_* kernels (csrc/lm.cu); both projections of a step are variable-coefficient multigrid solves
(multigrid/variable_coeff_MG.py), which makes this the third caller of the multigrid hot path."""
__all__ = ["simulation"]

from .simulation import Basestate, Simulation   # noqa: F401

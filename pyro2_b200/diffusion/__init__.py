"""Implicit (Crank-Nicolson) diffusion on the B200 -- the interface of pyro/diffusion: each step is one
constant-coefficient multigrid solve (the HP-2 path)."""
__all__ = ["simulation"]

from .simulation import Simulation   # noqa: F401

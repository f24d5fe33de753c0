"""Inviscid Burgers solver (unsplit CTU) on the B200 -- the interface of pyro/burgers; parent class of
the incompressible solver, whose explicit stages it shares (csrc/flow.cu)."""
__all__ = ["simulation"]

from .simulation import Simulation   # noqa: F401

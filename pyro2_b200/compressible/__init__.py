"""Compressible hydrodynamics (unsplit CTU Godunov, HLLC) on the B200: drop-in for the hot path of
pyro/compressible -- Simulation.evolve(), method_compute_timestep() and the ghost fill run as
hand-written sm_100a kernels (pyro2_b200/csrc)."""
__all__ = ["simulation"]

from .simulation import (Simulation, Variables, cons_to_prim,   # noqa: F401
                         prim_to_cons)

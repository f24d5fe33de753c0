"""Derived fields of the compressible state -- velocity, specific internal energy, pressure, sound
speed, Mach number, vorticity -- evaluated lazily with torch on the device (the reference computes
all of them eagerly in pyro/compressible/derives.py:6-69).  Off the hot path: the CFL reduction has
its own kernel."""
import torch

from . import eos


class _Fields:
    """lazy evaluation of the quantities the requested names need"""

    def __init__(self, myd):
        self.myd = myd
        self.rho = myd.get_var("density")
        self._cache = {}

    def _get(self, key, fn):
        if key not in self._cache:
            self._cache[key] = fn()
        return self._cache[key]

    u = property(lambda s: s._get("u", lambda: s.myd.get_var("x-momentum") / s.rho))
    v = property(lambda s: s._get("v", lambda: s.myd.get_var("y-momentum") / s.rho))
    e = property(lambda s: s._get("e", lambda: (s.myd.get_var("energy") - 0.5 * s.rho * (s.u * s.u + s.v * s.v)) / s.rho))
    p = property(lambda s: s._get("p", lambda: eos.pres(s.myd.get_aux("gamma"), s.rho, s.e)))
    cs = property(lambda s: s._get("cs", lambda: torch.sqrt(s.myd.get_aux("gamma") * s.p / s.rho)))

    def vorticity(self):
        g = self.myd.grid
        w = g.scratch_array()
        w.v()[:, :] = 0.5 * (self.v.ip(1) - self.v.ip(-1)) / g.dx - 0.5 * (self.u.jp(1) - self.u.jp(-1)) / g.dy
        return w


_RECIPES = {
    "velocity": lambda f: [f.u, f.v],
    "e": lambda f: [f.e], "eint": lambda f: [f.e],
    "p": lambda f: [f.p], "pressure": lambda f: [f.p],
    "primitive": lambda f: [f.rho, f.u, f.v, f.p],
    "soundspeed": lambda f: [f.cs],
    "machnumber": lambda f: [torch.sqrt(f.u ** 2 + f.v ** 2) / f.cs],
    "vorticity": lambda f: [f.vorticity()],
}


def derive_primitives(myd, varnames):
    """same contract as the reference: one name -> one array, several -> a list, unknown -> empty"""
    fields = _Fields(myd)
    out = []
    for name in ([varnames] if isinstance(varnames, str) else list(varnames)):
        if name in _RECIPES:
            out += _RECIPES[name](fields)
    if len(out) > 1:
        return out
    return out[0] if out else []

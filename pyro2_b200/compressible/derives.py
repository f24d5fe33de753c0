"""Derived fields of the compressible state (pyro/compressible/derives.py:6-69), evaluated with
torch on the device.  Off the hot path: the CFL reduction has its own kernel."""
import torch

from . import eos


def derive_primitives(myd, varnames):
    dens = myd.get_var("density")
    xmom = myd.get_var("x-momentum")
    ymom = myd.get_var("y-momentum")
    ener = myd.get_var("energy")
    u = xmom / dens
    v = ymom / dens
    e = (ener - 0.5 * dens * (u * u + v * v)) / dens
    gamma = myd.get_aux("gamma")
    p = eos.pres(gamma, dens, e)
    myg = myd.grid

    wanted = [varnames] if isinstance(varnames, str) else list(varnames)
    out = []
    for var in wanted:
        if var == "velocity":
            out += [u, v]
        elif var in ("e", "eint"):
            out.append(e)
        elif var in ("p", "pressure"):
            out.append(p)
        elif var == "primitive":
            out += [dens, u, v, p]
        elif var == "soundspeed":
            out.append(torch.sqrt(gamma * p / dens))
        elif var == "machnumber":
            out.append(torch.sqrt(u ** 2 + v ** 2) / torch.sqrt(gamma * p / dens))
        elif var == "vorticity":
            vort = myg.scratch_array()
            vort.v()[:, :] = 0.5 * (v.ip(1) - v.ip(-1)) / myg.dx - 0.5 * (u.jp(1) - u.jp(-1)) / myg.dy
            out.append(vort)
    if len(out) > 1:
        return out
    return out[0] if out else []

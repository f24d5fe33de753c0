"""helpers for reading the reference-generated fixtures in tests/golden/"""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _parse(v):
    for cast in (int, float):
        try:
            return cast(v)
        except ValueError:
            pass
    return v


def load_comp(name):
    z = np.load(os.path.join(GOLDEN, f"comp_{name}.npz"))
    rp = {s.split("=", 1)[0]: _parse(s.split("=", 1)[1]) for s in z["rp"]}
    inputs = {s.split("=", 1)[0]: _parse(s.split("=", 1)[1]) for s in z["inputs"]}
    return z, rp, inputs


def load_mg(name):
    return np.load(os.path.join(GOLDEN, f"mg_{name}.npz"))


def load_mgvc(name):
    return np.load(os.path.join(GOLDEN, f"mgvc_{name}.npz"))


def load_flow(fname):
    """incompressible / burgers fixtures: (npz, runtime parameters of the run, the inputs it was started with)"""
    z = np.load(os.path.join(GOLDEN, fname))
    rp = {s.split("=", 1)[0]: _parse(s.split("=", 1)[1]) for s in z["rp"]}
    inputs = {s.split("=", 1)[0]: _parse(s.split("=", 1)[1]) for s in z["inputs"]}
    return z, rp, inputs


def var_bcs(rp):
    """BC names per conserved variable (density, energy, x-momentum, y-momentum) following
    pyro/simulation_null.py:71-112: 'reflect' is even, except odd for the momentum normal to it"""
    sides = [rp["mesh.xlboundary"], rp["mesh.xrboundary"], rp["mesh.ylboundary"], rp["mesh.yrboundary"]]

    def res(odd_dir):
        out = []
        for k, s in enumerate(sides):
            d = "x" if k < 2 else "y"
            out.append(("reflect-odd" if d == odd_dir else "reflect-even") if s == "reflect" else s)
        return tuple(out)
    return [res(""), res(""), res("x"), res("y")]

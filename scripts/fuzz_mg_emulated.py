"""Randomised comparison of the multigrid library (mg.cu host-compiled under the CUDA emulator) with the oracle:
sizes 2..128, every boundary combination, constant and variable coefficients, Helmholtz terms, inhomogeneous
Dirichlet values, blocked and per-colour smoothers.  Development tool (CPU only):

    python scripts/fuzz_mg_emulated.py [ncases] [seed]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), ROOT]
import oracle  # noqa: E402
from emu_util import EmuMG, load_mg_emu  # noqa: E402


def sides(rng, kinds):
    out = []
    for _ in range(2):
        a = str(rng.choice(kinds + ["periodic"]))
        out += [a, a] if a == "periodic" else [a, str(rng.choice(kinds))]
    return tuple(out)


def one_case(lib, rng):
    n = int(rng.choice([2, 4, 8, 16, 32, 64, 128], p=[.1, .1, .15, .2, .2, .15, .1]))
    bc = sides(rng, ["dirichlet", "neumann"])
    vc = bool(rng.integers(3) == 0)
    alpha, beta = (0.0, 0.0) if vc else (float(rng.choice([0.0, 1.0, 2.5])), float(rng.choice([-1.0, 0.05, 3.0])))
    if not vc and alpha > 0.0 and beta < 0:
        beta = -beta                       # keep alpha - beta L definite (an indefinite one diverges to inf / NaN)
    blocking = bool(rng.integers(2))
    f = rng.standard_normal((n + 2, n + 2))
    if all(b in ("neumann", "periodic") for b in bc) and alpha == 0.0:
        f[1:-1, 1:-1] -= f[1:-1, 1:-1].mean()            # solvable singular problem
    v0 = rng.standard_normal((n + 2, n + 2)) * float(rng.integers(2))
    o = oracle.MG(n, bc=bc, alpha=alpha, beta=beta)
    m = EmuMG(lib, n, bc, alpha, beta, blocking)
    desc = dict(n=n, bc=bc, vc=vc, alpha=alpha, beta=beta, blocking=blocking)
    if vc:
        cbc = tuple("periodic" if b == "periodic" else str(rng.choice(["neumann", "reflect-even"])) for b in bc)
        coeffs = 0.5 + rng.random((n + 2, n + 2))
        o.set_coeffs(coeffs, cbc)
        m.set_coeffs(coeffs, cbc)
        desc["cbc"] = cbc
    vals = [None] * 4
    if not vc and rng.integers(3) == 0:
        for s in range(4):
            if bc[s] == "dirichlet" and rng.integers(2):
                vals[s] = rng.standard_normal(n + 2)
                o.set_bc_values(("xl", "xr", "yl", "yr")[s], vals[s])
        m.set_bc_values(*vals)
        desc["bcvals"] = [v is not None for v in vals]
    fine = o.nlevels - 1
    what = str(rng.choice(["smooth", "vcycle", "solve"]))
    desc["what"] = what
    o.plane(fine, "v")[:] = v0
    o.plane(fine, "f")[:] = f
    if what == "solve":
        o.init_RHS(f)
        o.solve(rtol=1e-10)
        got = m.solve(f, rtol=1e-10, v0=v0)
        ok = np.array_equal(got, o.get_solution()) and m.num_cycles == o.num_cycles
        desc["cycles"] = (m.num_cycles, o.num_cycles)
    else:
        m.plane(fine, "v")[:] = v0
        m.plane(fine, "f")[:] = f
        if what == "smooth":
            k = int(rng.choice([1, 3, 5, 7, 10]))
            o.smooth(fine, k)
            m.ck(lib.p2b_mg_smooth(m.h, fine, k, None))
        else:
            o.v_cycle()
            m.ck(lib.p2b_mg_zero_coarse(m.h, None))
            m.ck(lib.p2b_mg_vcycle(m.h, None))
        ok = np.array_equal(m.plane(fine, "v")[1:-1, 1:-1], o.plane(fine, "v")[1:-1, 1:-1])
    m.close()
    return ok, desc


if __name__ == "__main__":
    ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    lib = load_mg_emu()
    bad = 0
    for c in range(ncases):
        ok, desc = one_case(lib, rng)
        if not ok:
            bad += 1
            print("FAIL", c, desc, flush=True)
    print(f"{ncases} cases, {bad} failed")
    sys.exit(1 if bad else 0)

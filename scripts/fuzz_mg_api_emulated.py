"""Randomised check of the public multigrid classes (CellCenterMG2d, VarCoeffCCMG2d: the product's Python layer on
the emulated device) against the oracle: sizes, boundary types, inhomogeneous values, Helmholtz terms, non-zero initial
guesses, coefficient replacement in place, operator changes on an existing hierarchy, gradients of the solution.
Development tool (CPU only):

    python scripts/fuzz_mg_api_emulated.py [ncases] [seed]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), ROOT]
import emu_device  # noqa: E402
import oracle  # noqa: E402


def pair(rng, kinds):
    a = str(rng.choice(kinds + ["periodic"]))
    return (a, a) if a == "periodic" else (a, str(rng.choice(kinds)))


def one_case(rng):
    from pyro2_b200.mesh import boundary as bnd
    from pyro2_b200.mesh import patch
    from pyro2_b200.multigrid import MG
    from pyro2_b200.multigrid import variable_coeff_MG as VMG
    n = int(rng.choice([8, 16, 32, 64, 128], p=[.15, .25, .25, .2, .15]))
    bc = pair(rng, ["dirichlet", "neumann"]) + pair(rng, ["dirichlet", "neumann"])
    vc = bool(rng.integers(2))
    f = rng.standard_normal((n + 2, n + 2))
    desc = dict(n=n, bc=bc, vc=vc)
    if vc:
        cbc = tuple("periodic" if b == "periodic" else "neumann" for b in bc)
        if "dirichlet" not in bc:
            f[1:-1, 1:-1] -= f[1:-1, 1:-1].mean()
        c1, c2 = 0.5 + rng.random((n + 2, n + 2)), 0.5 + rng.random((n + 2, n + 2))
        g = patch.Grid2d(n, n, ng=1)
        d = patch.CellCenterData2d(g)
        bc_c = bnd.BC(xlb=cbc[0], xrb=cbc[1], ylb=cbc[2], yrb=cbc[3])
        d.register_var("c", bc_c)
        d.create()
        d.get_var("c")[:, :] = c1
        a = VMG.VarCoeffCCMG2d(n, n, xl_BC_type=bc[0], xr_BC_type=bc[1], yl_BC_type=bc[2], yr_BC_type=bc[3],
                               coeffs=d.get_var("c"), coeffs_bc=bc_c)
        ok = True
        for coeffs in (c1, c2):                     # the second pass replaces the coefficients in place
            if coeffs is c2:
                a.set_coeffs(c2)
            o = oracle.MG(n, bc=bc, alpha=0.0, beta=0.0)
            o.set_coeffs(coeffs, cbc)
            a.init_zeros(); a.init_RHS(f); a.solve(rtol=1e-10)
            o.init_zeros(); o.init_RHS(f); o.solve(rtol=1e-10)
            ok &= a.num_cycles == o.num_cycles and np.array_equal(a.get_solution().numpy(), o.get_solution())
        return bool(ok), desc
    alpha, beta = (0.0, -1.0) if rng.integers(2) else (float(rng.choice([1.0, 2.5])), float(rng.choice([0.05, 1e-3, 3.0])))
    if alpha == 0.0 and "dirichlet" not in bc:
        f[1:-1, 1:-1] -= f[1:-1, 1:-1].mean()
    kw, vals = {}, {}
    if rng.integers(2):
        for side, name, key in zip(bc, ("xl_BC", "xr_BC", "yl_BC", "yr_BC"), ("xl", "xr", "yl", "yr")):
            if side == "dirichlet" and rng.integers(2):
                c0, c1 = rng.standard_normal(2)
                kw[name] = (lambda a_, b_: (lambda s: a_ + b_ * np.cos(2.0 * s)))(c0, c1)
                x = (np.arange(n + 2) - 0.5) / n
                vals[key] = kw[name](x)
    a = MG.CellCenterMG2d(n, n, xl_BC_type=bc[0], xr_BC_type=bc[1], yl_BC_type=bc[2], yr_BC_type=bc[3],
                          alpha=alpha, beta=beta, **kw)
    o = oracle.MG(n, bc=bc, alpha=alpha, beta=beta)
    for key, v in vals.items():
        o.set_bc_values(key, v)
    v0 = rng.standard_normal((n + 2, n + 2)) * float(rng.integers(2))
    a.init_solution(v0); a.init_RHS(f); a.solve(rtol=1e-10)
    o.init_solution(v0); o.init_RHS(f); o.solve(rtol=1e-10)
    ok = a.num_cycles == o.num_cycles and np.array_equal(a.get_solution().numpy(), o.get_solution())
    desc.update(alpha=alpha, beta=beta, inhom=sorted(vals), cycles=(a.num_cycles, o.num_cycles))
    # gradient of the solution (MG.py:573-600): centred differences of the solution incl. its ghost cells
    gx, gy = a.get_solution_gradient()
    s = o.get_solution()
    ok &= np.array_equal(gx.numpy()[1:-1, 1:-1], 0.5 * (s[2:, 1:-1] - s[:-2, 1:-1]) / (1.0 / n))
    if alpha != 0.0 and not vals:
        # change the operator of the existing hierarchy (what the diffusion solver does every step)
        beta2 = beta * 0.37
        a.set_operator(alpha, beta2)
        o2 = oracle.MG(n, bc=bc, alpha=alpha, beta=beta2)
        a.init_zeros(); a.init_RHS(f); a.solve(rtol=1e-10)
        o2.init_zeros(); o2.init_RHS(f); o2.solve(rtol=1e-10)
        ok &= a.num_cycles == o2.num_cycles and np.array_equal(a.get_solution().numpy(), o2.get_solution())
    return bool(ok), desc


if __name__ == "__main__":
    ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    bad = 0
    with emu_device.emulated_device():
        for c in range(ncases):
            try:
                ok, desc = one_case(rng)
            except (AssertionError, ValueError, IndexError) as e:
                ok, desc = False, {"exception": repr(e)}
            if not ok:
                bad += 1
                print("FAIL", c, desc, flush=True)
    print(f"{ncases} cases, {bad} failed")
    sys.exit(1 if bad else 0)

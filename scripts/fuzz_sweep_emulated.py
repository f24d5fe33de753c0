"""Randomised comparison of the fused sweep's task source (host warp emulator) with the oracle: grid shapes down to a
single cell, every boundary combination the solver accepts, the three Riemann solvers, limiters, flattening, gravity,
heating, sponge, random segment lengths.  Development tool (CPU only):

    python scripts/fuzz_sweep_emulated.py [ncases] [seed]

Prints one line per failing case with the parameters to reproduce it; exit status 1 if any case failed."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), ROOT]
import oracle  # noqa: E402
from conftest import rel_l2  # noqa: E402
from emu_util import load_sweep_emu  # noqa: E402
from golden_util import var_bcs  # noqa: E402
from test_sweep_emulated import _emu_step  # noqa: E402

XBC = [("outflow", "outflow"), ("periodic", "periodic"), ("reflect", "reflect"), ("reflect", "outflow"), ("outflow", "reflect")]
YBC = XBC + [("hse", "hse"), ("reflect", "ambient"), ("hse", "outflow"), ("reflect", "hse")]


def one_case(lib, rng):
    ng, gamma = 4, 1.4
    nx = int(rng.choice([1, 2, 3, 5, 8, 13, 24, 31, 40, 67]))
    ny = int(rng.choice([1, 2, 4, 7, 16, 29, 30, 31, 45, 61, 64]))
    xbc, ybc = XBC[rng.integers(len(XBC))], YBC[rng.integers(len(YBC))]
    if "periodic" in xbc and nx < ng or "reflect" in xbc and nx < ng:
        xbc = ("outflow", "outflow")
    if ("periodic" in ybc or "reflect" in ybc) and ny < ng:
        ybc = ("outflow", "outflow")
    bc = xbc + ybc
    grav = float(rng.choice([0.0, 0.0, -1.5, 0.8])) if not ("hse" in bc) else float(rng.choice([-1.5, -0.4]))
    riemann = str(rng.choice(["HLLC", "CGF", "HLLC_lm"]))
    limiter, flat = int(rng.integers(3)), int(rng.integers(2))
    heat_on = bool(rng.integers(3) == 0)
    sponge = (0.6, 0.2, 1.e-2) if rng.integers(4) == 0 else None
    seglen = int(rng.choice([8, 9, 16, 31, 64, 200]))
    cvisc = float(rng.choice([0.1, 0.0, 0.3]))
    dx, dy = 1.0 / nx, float(rng.choice([1.0, 2.0])) / ny
    qx, qy = nx + 2 * ng, ny + 2 * ng
    x = (np.arange(qx) + 0.5 - ng) * dx
    y = (np.arange(qy) + 0.5 - ng) * dy
    dens = np.broadcast_to(1.5 * np.exp(-y / 0.9)[None, :], (qx, qy)) * (1.0 + 0.1 * rng.standard_normal((qx, qy)))
    jump = 1.0 + 4.0 * (rng.random() < 0.5) * ((x[:, None] + 0.7 * y[None, :]) < 0.6)         # sometimes a strong pressure jump
    pres = 1.8 * dens * jump * (1.0 + 0.05 * rng.standard_normal((qx, qy)))
    u, v = 0.3 * rng.standard_normal((qx, qy)), 0.3 * rng.standard_normal((qx, qy))
    P = np.stack([dens, pres / (gamma - 1.0) + 0.5 * dens * (u * u + v * v), dens * u, dens * v])
    prof = np.exp(-(np.sqrt((x[:, None] - 0.5) ** 2 + (y[None, :] - 0.7) ** 2) / 0.3) ** 2)
    bcs = var_bcs(dict(zip(("mesh.xlboundary", "mesh.xrboundary", "mesh.ylboundary", "mesh.yrboundary"), bc)))
    for k in range(4):
        oracle.fill_ghost(P[k], ng, bcs[k])
        for side in ("ylb", "yrb"):
            b = bcs[k][2 + (side == "yrb")]
            if b == "hse":
                oracle.fill_hse(P, ng, dy, grav, gamma, k, side)
            if b == "ambient":
                P[k][:, ng + ny:] = (0.05, 0.4, 0.0, 0.0)[k]
    U = oracle.from_planes(P)
    dt = float(rng.choice([0.5, 1.0])) * oracle.cfl_dt(U, ng, dx, dy, gamma, 0.8)
    prm = oracle.comp_params(gamma=gamma, cvisc=cvisc, limiter=limiter, use_flattening=flat, grav=grav, src_bcs=bcs,
                             riemann=riemann, xl_solid=int(bc[0] == "reflect"), yl_solid=int(bc[2] == "reflect"),
                             heat_rate=0.7 if heat_on else 0.0, heat_profile=prof if heat_on else None, sponge=sponge)
    heat = None
    if heat_on:
        heat = prof.copy()
        oracle.fill_ghost(heat, ng, tuple("outflow" if b in ("hse", "ambient") else b for b in bcs[1]))
    flips = (int(bc[2] == "reflect"), int(bc[3] == "reflect"))
    desc = dict(nx=nx, ny=ny, bc=bc, grav=grav, riemann=riemann, limiter=limiter, flat=flat, heat=heat_on, sponge=sponge,
                seglen=seglen, cvisc=cvisc)
    try:
        ref = oracle.compressible_step(U, ng, dx, dy, dt, prm)
    except AssertionError:
        return None, desc           # the random state went invalid in the oracle as well: not a parity question
    vv = (slice(ng, ng + nx), slice(ng, ng + ny))
    if not np.isfinite(ref[vv]).all():
        return None, desc           # e.g. hse ghost cells with a negative pressure (one row, large dy): NaN in the oracle too
    got, scratch = _emu_step(lib, U, ng, dx, dy, dt, prm, seglen, flips, heat=heat, src_copy_yhi=int(bc[3] == "ambient"))
    errs = [rel_l2(got[vv][..., n], ref[vv][..., n]) for n in range(4)]
    ok = not np.isnan(got[vv]).any() and max(errs) < 1e-12 and scratch[3] == 0
    return ok, dict(desc, errs=[float(f"{e:.2e}") for e in errs], status=int(scratch[3]))


def spherical_case(lib, rng):
    """the SPH instantiation: SphericalPolar geometry, CGF; physical boundary types only (the literal reflect-odd /
    reflect-even types put negative densities into the ghost cells, where the CGF star states are ill-conditioned)"""
    from pyro2_b200.mesh import patch
    ng, gamma = 4, 1.4
    nx = int(rng.choice([4, 5, 8, 13, 24, 31, 40]))
    ny = int(rng.choice([4, 7, 16, 29, 30, 31, 45, 61]))
    xb = [("outflow", "outflow"), ("periodic", "periodic"), ("reflect", "reflect"), ("reflect", "outflow"), ("outflow", "reflect")]
    bc = xb[rng.integers(len(xb))] + xb[rng.integers(len(xb))]
    xmin = float(rng.choice([0.5, 1.0, 3.0]))
    xmax = xmin + float(rng.choice([0.5, 1.0]))
    if xmin - ng * (xmax - xmin) / nx < 0.0:        # the ghost rows must keep r > 0
        xmin, xmax = 3.0, 3.0 + (xmax - xmin)
    ymin, ymax = float(rng.choice([0.3, 0.785])), float(rng.choice([2.0, 2.8]))
    g = patch.SphericalPolar(nx, ny, ng=ng, xmin=xmin, xmax=xmax, ymin=ymin, ymax=ymax, device="cpu")
    geom = oracle.spherical_geometry(nx, ny, ng, xmin, xmax, ymin, ymax)
    grav = float(rng.choice([0.0, -1.0, 0.6]))
    limiter, flat, cvisc = int(rng.integers(3)), int(rng.integers(2)), float(rng.choice([0.1, 0.0]))
    seglen = int(rng.choice([8, 9, 16, 31, 64]))
    qx, qy = nx + 2 * ng, ny + 2 * ng
    dens = (1.0 + 0.5 / geom["x2d"]) * (1.0 + 0.1 * rng.standard_normal((qx, qy)))
    pres = 1.5 * dens * (1.0 + 0.05 * rng.standard_normal((qx, qy)))
    u, v = 0.3 * rng.standard_normal((qx, qy)), 0.3 * rng.standard_normal((qx, qy))
    P = np.stack([dens, pres / (gamma - 1.0) + 0.5 * dens * (u * u + v * v), dens * u, dens * v])
    bcs = var_bcs(dict(zip(("mesh.xlboundary", "mesh.xrboundary", "mesh.ylboundary", "mesh.yrboundary"), bc)))
    for k in range(4):
        oracle.fill_ghost(P[k], ng, bcs[k])
    U = oracle.from_planes(P)
    dt = float(rng.choice([0.5, 1.0])) * oracle.cfl_dt_spherical(U, gamma, 0.8, geom)
    prm = oracle.comp_params(gamma=gamma, cvisc=cvisc, limiter=limiter, use_flattening=flat, grav=grav, src_bcs=bcs,
                             riemann="CGF", geom=geom, xl_solid=int(bc[0] == "reflect"), yl_solid=int(bc[2] == "reflect"))
    desc = dict(spherical=True, nx=nx, ny=ny, bc=bc, grav=grav, limiter=limiter, flat=flat, cvisc=cvisc, seglen=seglen,
                x=(xmin, xmax), y=(ymin, ymax))
    try:
        ref = oracle.compressible_step(U, ng, g.dx, g.dy, dt, prm)
    except AssertionError:
        return None, desc
    tables = patch.spherical_sweep_tables(g, (qy + 15) // 16 * 16, bc[0], bc[1])
    got, scratch = _emu_step(lib, U, ng, g.dx, g.dy, dt, prm, seglen, geometry=tables,
                             xflips=(int(bc[0] == "reflect"), int(bc[1] == "reflect")))
    vv = (slice(ng, ng + nx), slice(ng, ng + ny))
    err = float(np.abs(got[vv] - ref[vv]).max() / np.abs(ref[vv]).max())
    ok = bool(np.isfinite(got[vv]).all()) and err < 1e-12 and scratch[3] == 0
    return ok, dict(desc, err=float(f"{err:.2e}"))


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rng = np.random.default_rng(seed)
    lib = load_sweep_emu()
    bad = skipped = 0
    for c in range(n):
        ok, desc = spherical_case(lib, rng) if c % 4 == 3 else one_case(lib, rng)
        if ok is None:
            skipped += 1
        elif not ok:
            bad += 1
            print("FAIL", c, desc, flush=True)
    print(f"{n} cases, {skipped} skipped (invalid in the oracle too), {bad} failed")
    sys.exit(1 if bad else 0)

"""Randomised comparison of the ghost-fill, CFL and explicit-stage kernels (host-compiled under the CUDA emulator) with
the oracle: tiny and odd grid shapes, every boundary combination, all limiters.  Development tool (CPU only):

    python scripts/fuzz_ghost_flow_emulated.py [ncases] [seed]
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), ROOT]
import oracle  # noqa: E402
from emu_util import EmuFlow, load_flow_emu, load_ghost_emu  # noqa: E402
from pyro2_b200 import _lib  # noqa: E402

BCS = ["outflow", "periodic", "reflect-even", "reflect-odd", "dirichlet", "neumann"]


def ghost_case(lib, rng):
    ng = int(rng.integers(1, 5))
    nx, ny = int(rng.integers(1, 14)), int(rng.integers(1, 14))
    nvar = int(rng.integers(1, 4))
    bcs = []
    for _ in range(nvar):
        xb = str(rng.choice(BCS))
        yb = str(rng.choice(BCS))
        # periodic comes in pairs; a periodic / reflecting side needs at least ng valid cells (array_indexer.py reads them)
        xs = (xb, xb) if xb == "periodic" else (xb, str(rng.choice([b for b in BCS if b != "periodic"])))
        ys = (yb, yb) if yb == "periodic" else (yb, str(rng.choice([b for b in BCS if b != "periodic"])))
        bcs.append(xs + ys)
    if nx < ng and any(b in ("periodic", "reflect-even", "reflect-odd") for bc in bcs for b in bc[:2]):
        nx = ng
    if ny < ng and any(b in ("periodic", "reflect-even", "reflect-odd") for bc in bcs for b in bc[2:]):
        ny = ng
    qx, qy = nx + 2 * ng, ny + 2 * ng
    pitch = (qy + 15) // 16 * 16 if qy >= 16 else (qy + 1) // 2 * 2
    dtype = np.float64 if rng.integers(2) else np.int64
    P = np.zeros((nvar, qx, pitch), dtype=dtype)
    P[:, :, :qy] = rng.integers(-50, 50, (nvar, qx, qy)) if dtype == np.int64 else rng.standard_normal((nvar, qx, qy))
    ref = P.copy()
    for k in range(nvar):
        a = np.ascontiguousarray(ref[k, :, :qy])
        oracle.fill_ghost(a, ng, bcs[k])
        ref[k, :, :qy] = a
    g = _lib.Grid(nx, ny, ng, pitch, qx * pitch, 1.0, 1.0)
    arr = _lib.bc_array(bcs)
    f = lib.p2b_fill_ghost_f64 if dtype == np.float64 else lib.p2b_fill_ghost_i64
    rc = f(P.ctypes.data, C.byref(g), nvar, arr, None)
    ok = rc == 0 and np.array_equal(P[:, :, :qy], ref[:, :, :qy])
    return ok, dict(kind="ghost", nx=nx, ny=ny, ng=ng, bcs=bcs, dtype=dtype.__name__, rc=rc)


def cfl_case(lib, rng):
    """max(|u| + cs), max(|v| + cs) over the whole padded array -> the reference's dt, bit for bit"""
    ng = int(rng.integers(1, 5))
    nx, ny = int(rng.integers(1, 70)), int(rng.integers(1, 70))
    qx, qy = nx + 2 * ng, ny + 2 * ng
    pitch = (qy + 15) // 16 * 16 if qy >= 16 else (qy + 1) // 2 * 2
    gamma = float(rng.choice([1.4, 5.0 / 3.0]))
    dens = 0.1 + rng.random((qx, qy)) * float(rng.choice([1.0, 100.0]))
    pres = 0.01 + rng.random((qx, qy)) * float(rng.choice([1.0, 1000.0]))
    u, v = rng.standard_normal((qx, qy)) * 3.0, rng.standard_normal((qx, qy)) * 3.0
    P = np.zeros((4, qx, pitch))
    P[:, :, :qy] = np.stack([dens, pres / (gamma - 1.0) + 0.5 * dens * (u * u + v * v), dens * u, dens * v])
    P[0, :, qy:] = 1.0                      # padding columns are never read: make them valid but extreme
    P[1, :, qy:] = 1e9
    dx, dy = 1.0 / nx, float(rng.choice([1.0, 0.3])) / ny
    g = _lib.Grid(nx, ny, ng, pitch, qx * pitch, dx, dy)
    scratch = np.zeros(8, dtype=np.uint64)
    rc = lib.p2b_cfl_wavemax(P.ctypes.data, C.byref(g), gamma, scratch.ctypes.data, None)
    w = scratch[:2].view(np.float64)
    U = oracle.from_planes(np.ascontiguousarray(P[:, :, :qy]))
    ok = rc == 0 and 0.8 * min(dx / w[0], dy / w[1]) == oracle.cfl_dt(U, ng, dx, dy, gamma, 0.8)
    return ok, dict(kind="cfl", nx=nx, ny=ny, ng=ng, gamma=gamma)


def ghost_values_case(lib, rng):
    """inhomogeneous Dirichlet / Neumann values on some sides (array_indexer.py:165-183, 220-238)"""
    ng = int(rng.integers(1, 5))
    nx, ny = int(rng.integers(ng, 14)), int(rng.integers(ng, 14))
    kinds = ["dirichlet", "neumann", "outflow", "reflect-even"]
    xb, yb = str(rng.choice(kinds + ["periodic"])), str(rng.choice(kinds + ["periodic"]))
    bc = ((xb, xb) if xb == "periodic" else (xb, str(rng.choice(kinds)))) + ((yb, yb) if yb == "periodic" else (yb, str(rng.choice(kinds))))
    qx, qy = nx + 2 * ng, ny + 2 * ng
    pitch = (qy + 15) // 16 * 16 if qy >= 16 else (qy + 1) // 2 * 2
    vals = []
    for s_, b in enumerate(bc):
        length = qy if s_ < 2 else qx
        vals.append(rng.standard_normal(length) if b in ("dirichlet", "neumann") and rng.integers(3) else None)
    dx, dy = 1.0 / nx, 0.7 / ny
    a = np.zeros((qx, pitch))
    a[:, :qy] = rng.standard_normal((qx, qy))
    ref = np.ascontiguousarray(a[:, :qy])
    oracle.fill_ghost(ref, ng, bc, values=tuple(vals), dx=dx, dy=dy)
    g = _lib.Grid(nx, ny, ng, pitch, qx * pitch, dx, dy)
    codes = (C.c_int * 4)(*[_lib.BC_CODES[b] for b in bc])
    ptr = [None if (v is None or codes[s_] not in (0, 2)) else v.ctypes.data for s_, v in enumerate(vals)]
    rc = lib.p2b_fill_ghost_values_f64(a.ctypes.data, C.byref(g), codes, *ptr, None)
    ok = rc == 0 and np.array_equal(a[:, :qy], ref)
    return ok, dict(kind="ghost_values", nx=nx, ny=ny, ng=ng, bc=bc, given=[v is not None for v in vals], rc=rc)


def flow_case(lib, rng):
    ng = 4
    nx, ny = int(rng.integers(4, 40)), int(rng.integers(4, 40))
    limiter = int(rng.integers(3))
    dx, dy = 1.0 / nx, float(rng.choice([1.0, 0.5])) / ny
    qx, qy = nx + 2 * ng, ny + 2 * ng
    u, v = rng.standard_normal((qx, qy)), rng.standard_normal((qx, qy))
    dt = 0.4 * min(dx / np.abs(u).max(), dy / np.abs(v).max())
    f = EmuFlow(lib, nx, ny, ng, dx, dy)
    which = str(rng.choice(["burgers", "advection"]))
    if which == "burgers":
        ou, ov = oracle.burgers_evolve(u.copy(), v.copy(), ng, dx, dy, dt, limiter)
        gu, gv = u.copy(), v.copy()
        f.burgers_evolve(gu, gv, dt, limiter)
        v_ = (slice(ng, ng + nx), slice(ng, ng + ny))
        ok = np.array_equal(gu[v_], ou[v_]) and np.array_equal(gv[v_], ov[v_])
    else:
        a = rng.standard_normal((qx, qy))
        cu, cv = float(rng.standard_normal()), float(rng.standard_normal())
        ref = oracle.advection_evolve(a.copy(), ng, dx, dy, dt, cu, cv, limiter)
        got = a.copy()
        f.ck(lib.p2b_flow_advection_update(f.h, got.ctypes.data, cu, cv, dt, limiter, None))
        v_ = (slice(ng, ng + nx), slice(ng, ng + ny))
        ok = np.array_equal(got[v_], ref[v_])
    f.close()
    return ok, dict(kind=which, nx=nx, ny=ny, limiter=limiter)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    ghost, flow = load_ghost_emu(), load_flow_emu()
    bad = 0
    for c in range(n):
        ok, desc = (flow_case, cfl_case, ghost_case, ghost_values_case)[c % 4](flow if c % 4 == 0 else ghost, rng)
        if not ok:
            bad += 1
            print("FAIL", c, desc, flush=True)
    print(f"{n} cases, {bad} failed")
    sys.exit(1 if bad else 0)

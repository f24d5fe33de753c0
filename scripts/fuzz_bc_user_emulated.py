"""Randomised check of the compressible user boundaries (hse, ambient kernels in bc_user.cu, host-compiled) against the
oracle: grid shapes, gravity, both sides.  Development tool (CPU only):

    python scripts/fuzz_bc_user_emulated.py [ncases] [seed]
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), ROOT]
import oracle  # noqa: E402
from emu_util import load_bc_emu  # noqa: E402
from pyro2_b200 import _lib  # noqa: E402

if __name__ == "__main__":
    ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    lib = load_bc_emu()
    bad = 0
    for c in range(ncases):
        ng = 4
        nx, ny = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        qx, qy = nx + 2 * ng, ny + 2 * ng
        pitch = (qy + 15) // 16 * 16 if qy >= 16 else (qy + 1) // 2 * 2
        gamma, grav, dy = 1.4, float(rng.choice([-1.0, -9.8, 0.5])), float(rng.choice([0.1, 0.01, 1.0]))
        P = np.zeros((4, qx, pitch))
        dens = 0.5 + rng.random((qx, qy))
        u, v = rng.standard_normal((qx, qy)), rng.standard_normal((qx, qy))
        pres = 1.0 + rng.random((qx, qy))
        P[:, :, :qy] = np.stack([dens, pres / (gamma - 1.0) + 0.5 * dens * (u * u + v * v), dens * u, dens * v])
        ref = np.ascontiguousarray(P[:, :, :qy])
        g = _lib.Grid(nx, ny, ng, pitch, qx * pitch, 1.0, dy)
        ok = True
        for var in rng.permutation(4):
            for side in (0, 1):
                rc = lib.p2b_fill_hse_f64(P.ctypes.data, C.byref(g), grav, gamma, int(var), side, None)
                oracle.fill_hse(ref, ng, dy, grav, gamma, int(var), ("ylb", "yrb")[side])
                ok &= rc == 0
        ok &= np.array_equal(P[:, :, :qy], ref)
        val = float(rng.standard_normal())
        rc = lib.p2b_fill_ambient_f64(P.ctypes.data, C.byref(g), 2, 1, val, None)
        ref[2][:, ng + ny:] = val
        ok &= rc == 0 and np.array_equal(P[:, :, :qy], ref)
        if not ok:
            bad += 1
            print("FAIL", c, dict(nx=nx, ny=ny, grav=grav, dy=dy), flush=True)
    print(f"{ncases} cases, {bad} failed")
    sys.exit(1 if bad else 0)

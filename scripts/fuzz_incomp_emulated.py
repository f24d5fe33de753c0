"""Randomised comparison of the incompressible step (flow.cu stage kernels + multigrid projections, host-compiled under
the CUDA emulator) with the oracle: random periodic velocity fields and lagged pressure gradients on 16^2 / 32^2 grids,
every limiter and projection type.  Development tool (CPU only):

    python scripts/fuzz_incomp_emulated.py [ncases] [seed]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), ROOT]
import oracle  # noqa: E402
from emu_util import EmuFlow, load_flow_emu, load_mg_emu  # noqa: E402

if __name__ == "__main__":
    ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    flow, mg = load_flow_emu(), load_mg_emu()
    ng, bc = 4, ("periodic",) * 4
    fill = lambda a: oracle.fill_ghost(a, ng, bc)
    bad = 0
    for c in range(ncases):
        n = int(rng.choice([16, 32]))
        limiter, proj = int(rng.integers(3)), int(rng.integers(1, 3))
        x = (np.arange(n + 2 * ng) + 0.5 - ng) / n
        X, Y = np.meshgrid(x, x, indexing="ij")
        P = np.zeros((6, n + 2 * ng, n + 2 * ng))
        kx, ky = rng.integers(1, 4, 2)
        P[0] = np.sin(2 * np.pi * kx * X) * np.cos(2 * np.pi * ky * Y) + 0.3 * rng.standard_normal(X.shape)
        P[1] = -np.cos(2 * np.pi * kx * X) * np.sin(2 * np.pi * ky * Y) + 0.3 * rng.standard_normal(X.shape)
        P[3] = 0.1 * rng.standard_normal(X.shape)                     # phi (initial guess of the final projection)
        P[4], P[5] = 0.2 * rng.standard_normal(X.shape), 0.2 * rng.standard_normal(X.shape)
        Q = P.copy()
        dt = float(rng.choice([0.2, 0.8])) / n / max(np.abs(P[0]).max(), np.abs(P[1]).max())
        f = EmuFlow(flow, n, n, ng, 1.0 / n, 1.0 / n)
        for step in range(2):
            for k in range(6):
                fill(P[k])
                fill(Q[k])
            f.incomp_evolve(mg, P, dt, limiter, proj, bc, fill)
            oracle.incomp_evolve(Q, ng, dt, limiter=limiter, proj_type=proj)
        f.close()
        if not np.array_equal(P, Q):
            bad += 1
            print("FAIL", c, dict(n=n, limiter=limiter, proj=proj, maxdiff=np.abs(P - Q).max(axis=(1, 2)).tolist()), flush=True)
    print(f"{ncases} cases, {bad} failed")
    sys.exit(1 if bad else 0)

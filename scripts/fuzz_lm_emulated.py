"""Randomised comparison of the low Mach number atmosphere step (lm.cu stage kernels + the variable-coefficient
multigrid, host-compiled under the CUDA emulator) with the oracle: the bubble fixture's base state with randomly
perturbed fields, every limiter and projection type.  Development tool (CPU only):

    python scripts/fuzz_lm_emulated.py [ncases] [seed]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), ROOT]
import oracle  # noqa: E402
from emu_util import EmuLm, load_lm_emu, load_mg_emu  # noqa: E402
from golden_util import load_flow  # noqa: E402
from oracle_runs import lm_setup as _lm_setup  # noqa: E402

if __name__ == "__main__":
    ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    z, rp, _ = load_flow("lm_bubble32.npz")
    ng, n = int(z["ng"]), rp["mesh.nx"]
    names, fills, _ = _lm_setup(z, rp)
    base = np.ascontiguousarray(z["base"])
    bad = 0
    for c in range(ncases):
        limiter, proj = int(rng.integers(3)), int(rng.integers(1, 3))
        prm = oracle.lm_params(n, grav=rp["lm-atmosphere.grav"], gamma=rp["eos.gamma"], limiter=limiter, proj_type=proj,
                               xmin=rp["mesh.xmin"], xmax=rp["mesh.xmax"], ymin=rp["mesh.ymin"], ymax=rp["mesh.ymax"])
        S = np.ascontiguousarray(z["P0"]).copy()
        amp = float(rng.choice([0.01, 0.1, 0.4]))
        S[0] *= 1.0 + 0.2 * amp * rng.standard_normal(S[0].shape)          # density
        S[1] += amp * rng.standard_normal(S[1].shape)                      # velocities
        S[2] += amp * rng.standard_normal(S[2].shape)
        S[6] += 0.1 * amp * rng.standard_normal(S[6].shape)                # lagged pressure gradient
        S[7] += 0.1 * amp * rng.standard_normal(S[7].shape)
        for k, name in enumerate(names):
            oracle.fill_ghost(S[k], ng, fills[name])
        R = S.copy()
        e = EmuLm(load_lm_emu(), load_mg_emu(), n, ng, base, fills, fills["phi"], grav=rp["lm-atmosphere.grav"],
                  gamma=rp["eos.gamma"], limiter=limiter, proj_type=proj)
        dt_e, dt_o = e.timestep(S, 0.8), oracle.lm_timestep(R, base, prm, 0.8)
        ok = dt_e == dt_o
        if ok:
            cyc_e = e.evolve(S, dt_e)
            cyc_o = oracle.lm_evolve(R, base, prm, dt_o)
            ok = cyc_e == cyc_o and np.array_equal(S, R)
        e.close()
        if not ok:
            bad += 1
            print("FAIL", c, dict(limiter=limiter, proj=proj, amp=amp, dt=(dt_e, dt_o)), flush=True)
    print(f"{ncases} cases, {bad} failed")
    sys.exit(1 if bad else 0)

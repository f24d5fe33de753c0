"""CPU: the low Mach number atmosphere solver's stage kernels and host entry points (pyro2_b200/csrc/lm.cu,
unchanged) compiled for the host through tests/emu/cuda_emu.h and strung together exactly as
pyro2_b200/lm_atm/simulation.py does on the device (emulated variable-coefficient multigrid for the projections),
compared bit-for-bit with the oracle and the reference-generated fixtures.  Test infrastructure only."""
import numpy as np
import pytest

import oracle
from emu_util import EmuLm, load_lm_emu, load_mg_emu
from golden_util import load_flow
from oracle_runs import lm_setup as _lm_setup


@pytest.mark.parametrize("fname,nsteps", [("lm_bubble32.npz", 3), ("lm_bubble64_lim1.npz", 2)])
def test_emulated_lm_atm_steps_match_oracle_and_reference(fname, nsteps):
    z, rp, _ = load_flow(fname)
    ng, n = int(z["ng"]), rp["mesh.nx"]
    names, fills, prm = _lm_setup(z, rp)
    base = np.ascontiguousarray(z["base"])
    e = EmuLm(load_lm_emu(), load_mg_emu(), n, ng, base, fills, fills["phi"], grav=rp["lm-atmosphere.grav"],
              gamma=rp["eos.gamma"], limiter=rp["lm-atmosphere.limiter"], proj_type=rp["lm-atmosphere.proj_type"])
    S = np.ascontiguousarray(z["P0"])
    R = S.copy()
    for dt in z["dts"][:nsteps]:
        for k, name in enumerate(names):
            oracle.fill_ghost(S[k], ng, fills[name])
            oracle.fill_ghost(R[k], ng, fills[name])
        assert e.timestep(S, rp["driver.cfl"]) == oracle.lm_timestep(R, base, prm, rp["driver.cfl"])
        cyc = e.evolve(S, float(dt))
        assert cyc == oracle.lm_evolve(R, base, prm, float(dt))
        assert np.array_equal(S, R)
    if nsteps >= len(z["dts"]):
        assert np.array_equal(S, z["P"])
    e.close()


def test_emulated_lm_atm_initial_projection_matches_oracle():
    """preevolve's first half on a perturbed state (the fixtures start after preevolve)"""
    z, rp, _ = load_flow("lm_bubble32.npz")
    ng, n = int(z["ng"]), rp["mesh.nx"]
    names, fills, prm = _lm_setup(z, rp)
    base = np.ascontiguousarray(z["base"])
    rng = np.random.default_rng(2)
    S = np.ascontiguousarray(z["P0"])
    S[1] += 0.05 * rng.standard_normal(S[1].shape)
    S[2] += 0.05 * rng.standard_normal(S[2].shape)
    R = S.copy()
    e = EmuLm(load_lm_emu(), load_mg_emu(), n, ng, base, fills, fills["phi"])
    e.initial_projection(S)
    oracle.lm_initial_projection(R, base, prm)
    assert np.array_equal(S, R)
    e.close()

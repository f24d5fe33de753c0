"""Further stock problems through the public Pyro API against reference-generated fixtures: the second and the
multi-mode Rayleigh-Taylor setups, double Mach reflection with its "ramp" boundaries (device slice assignments
plus host-evaluated top rows), the low-Mach HLLC solver (RIEMANN = 2 instantiations of the sweep), the Burgers
convergence and tophat problems.  Same kernels and code paths as
the rt16 / burgers_test cases in test_gpu_api.py / test_gpu_flow.py; only the initial conditions differ (those
are checked on the CPU in test_capi_and_host.py)."""
import numpy as np
import pytest

from conftest import state_errors
from golden_util import load_comp, load_flow

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["rt2_48", "rt_multimode16", "ramp64", "gresho40_lm", "sedov32_lm"])
def test_pyro_compressible_rt_variants_match_reference(name):
    from pyro2_b200.pyro_sim import Pyro
    z, rp, inputs = load_comp(name)
    p = Pyro("compressible")
    p.initialize_problem(str(z["problem"]), inputs_dict=dict(inputs, **{"driver.max_steps": 100000}))
    sim = p.sim
    g = sim.cc_data.grid
    v = (slice(g.ilo, g.ihi + 1), slice(g.jlo, g.jhi + 1))
    assert np.array_equal(sim.cc_data.data.numpy()[v], z["U0"][v])
    dts = []
    for _ in range(len(z["dts"])):
        p.single_step()
        dts.append(sim.dt)
    assert sim.n == int(z["n"]) and sim.cc_data.t == pytest.approx(float(z["t"]), rel=1e-12)
    assert np.allclose(np.array(dts), z["dts"], rtol=1e-12, atol=0)
    assert max(state_errors(sim.cc_data.data.numpy()[v], z["U"][v], rp["eos.gamma"])) < 1e-10


@pytest.mark.parametrize("fname", ["burgers_converge32.npz", "burgers_tophat32.npz"])
def test_pyro_burgers_problems_match_reference(fname):
    from pyro2_b200.pyro_sim import Pyro
    z, rp, inputs = load_flow(fname)
    p = Pyro("burgers")
    p.initialize_problem(str(z["problem"]), inputs_dict=dict(inputs, **{"driver.max_steps": 100000}))
    sim = p.sim
    g = sim.cc_data.grid
    state = lambda: sim.cc_data.planes[:, :, :g.qy].cpu().numpy()
    assert np.array_equal(state(), z["P0"])
    dts = []
    for _ in range(len(z["dts"])):
        p.single_step()
        dts.append(sim.dt)
    assert np.array_equal(np.array(dts), z["dts"])
    assert np.array_equal(state(), z["P"])

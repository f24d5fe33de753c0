"""GPU: the public API against the regression files the REFERENCE ITSELF STORES (tests/golden/refh5_*.npz, re-packed from
pyro's .h5 goldens by tests/golden/make_h5_golden.py with the pure-Python reader tests/h5lite.py):

  pyro/compressible/tests/sod_x_0076.h5         Pyro("compressible"), problem sod, 128 x 10, 76 steps   (1e-11 of scale)
  pyro/multigrid/tests/mg_poisson_dirichlet.h5  CellCenterMG2d(256, 256).solve(rtol=1e-11)             (solution bit for bit)
  pyro/advection/tests/smooth_0040.h5           Pyro("advection"), smooth, 32^2, 40 steps               (bit for bit)
  pyro/burgers/tests/test_0051.h5               Pyro("burgers"), test, 128^2, 51 steps                  (bit for bit)
  pyro/diffusion/tests/gaussian_0164.h5         Pyro("diffusion"), gaussian, 128^2, 164 multigrid solves (bit for bit)
  pyro/incompressible/tests/shear_128_0216.h5   Pyro("incompressible"), shear, 128^2, 216 steps         (2e-12 of each variable's
                                                maximum: the unmodified reference run on this image is itself 6e-13 away)

  pyro/compressible/tests/quad_unsplit_0606.h5  Pyro("compressible"), quad, 256^2, 606 steps            (1e-10 of each variable's
  pyro/compressible/tests/rt_0945.h5            Pyro("compressible"), rt, 64 x 192, gravity, hse, 945   maximum = north_star's bar; the
                                                steps through the instability's growth                 oracle: 1e-13 / 4e-13)

The flow runs start from the runtime parameters recorded INSIDE the stored file (tests/golden/pin_stored_goldens.py), set
up the problem with this build's own problem modules and compute their own time steps.

(The file sorts last on purpose: these cases were added after the round's last GPU run and are rehearsed on the emulated
device, tests/test_gpu_rehearsal.py.)"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_pyro_sod_matches_the_stored_reference_golden():
    from golden_util import load_comp
    from pyro2_b200.pyro_sim import Pyro
    stored = np.load(os.path.join(GOLDEN, "refh5_sod_x_0076.npz"))
    z, rp, inputs = load_comp("sod_x")
    p = Pyro("compressible")
    p.initialize_problem("sod", inputs_dict=dict(inputs, **{"driver.max_steps": 100000}))
    for _ in range(76):
        p.single_step()
    p.sim.check_state()
    g = p.sim.cc_data.grid
    assert p.sim.n == 76 and p.sim.cc_data.t == pytest.approx(0.2, rel=1e-12)
    U = p.sim.cc_data.data.numpy()[g.ilo:g.ihi + 1, g.jlo:g.jhi + 1]
    for k, name in enumerate(("density", "energy", "x_momentum", "y_momentum")):
        scale = max(np.abs(stored[name]).max(), 1.0)
        assert np.abs(U[..., k] - stored[name]).max() <= 1e-11 * scale, name      # emulated device: 1.4e-14; oracle: 2e-14


def test_multigrid_matches_the_stored_reference_golden():
    from pyro2_b200.multigrid import MG
    stored = np.load(os.path.join(GOLDEN, "refh5_mg_poisson_dirichlet.npz"))
    a = MG.CellCenterMG2d(256, 256)
    x, y = a.x2d.numpy(), a.y2d.numpy()          # host arithmetic for the right-hand side, as the reference's test
    a.init_zeros()
    a.init_RHS(-2.0 * ((1.0 - 6.0 * x ** 2) * y ** 2 * (1.0 - y ** 2) + (1.0 - 6.0 * y ** 2) * x ** 2 * (1.0 - x ** 2)))
    a.solve(rtol=1.e-11)
    assert np.array_equal(a.get_solution().numpy()[1:-1, 1:-1], stored["v"])


@pytest.mark.parametrize("case,solver,problem", [("advection", "advection", "smooth"), ("burgers", "burgers", "test"),
                                                 ("diffusion", "diffusion", "gaussian")])
def test_pyro_flow_run_matches_the_stored_reference_golden(case, solver, problem):
    from golden_util import _parse
    from pyro2_b200.pyro_sim import Pyro
    z = np.load(os.path.join(GOLDEN, f"refh5_{case}.npz"))
    inputs = {s.split("=", 1)[0]: _parse(s.split("=", 1)[1]) for s in z["inputs"]}
    inputs = {k: v for k, v in inputs.items() if not k.startswith("particles.")}       # not part of this build (and off)
    p = Pyro(solver)
    p.initialize_problem(problem, inputs_dict=dict(inputs, **{"driver.max_steps": 100000, "driver.verbose": 0}))
    while not p.sim.finished():
        p.single_step()
    assert p.sim.n == int(z["n"]) and p.sim.cc_data.t == pytest.approx(float(z["t"]), rel=1e-12)
    for name, stored in zip(z["names"], z["stored"]):
        assert np.array_equal(p.sim.cc_data.get_var(str(name)).v().numpy(), stored), name


def test_pyro_incompressible_run_matches_the_stored_reference_golden():
    from golden_util import _parse
    from pyro2_b200.pyro_sim import Pyro
    z = np.load(os.path.join(GOLDEN, "refh5_incomp.npz"))
    inputs = {s.split("=", 1)[0]: _parse(s.split("=", 1)[1]) for s in z["inputs"]}
    inputs = {k: v for k, v in inputs.items() if not k.startswith("particles.")}
    p = Pyro("incompressible")
    p.initialize_problem("shear", inputs_dict=dict(inputs, **{"driver.max_steps": 100000, "driver.verbose": 0}))
    dts = []
    while not p.sim.finished():
        p.single_step()
        dts.append(p.sim.dt)
    assert p.sim.n == int(z["n"]) == 216 and p.sim.cc_data.t == pytest.approx(float(z["t"]), rel=1e-12)
    assert dts == [float(d) for d in z["dts"]]                  # the reference's own time steps (run on this image), bit for bit
    for name, stored in zip(z["names"], z["stored"]):
        got = p.sim.cc_data.get_var(str(name)).v().numpy()
        assert np.abs(got - stored).max() <= 2e-12 * np.abs(stored).max(), name


@pytest.mark.parametrize("case,problem", [("rt", "rt"), ("quad", "quad")])      # the one case never rehearsed end to end runs last
def test_pyro_compressible_run_matches_the_stored_reference_golden(case, problem):
    """the reference's two long compressible regression runs from the parameters inside the stored files, with this build's
    problem setups and time-step control: same number of steps to tmax, the state within north_star's 1e-10 of each
    variable's maximum (every second cell of the 256^2 file is kept in the fixture)"""
    from golden_util import _parse
    from pyro2_b200.pyro_sim import Pyro
    z = np.load(os.path.join(GOLDEN, f"refh5_{case}.npz"))
    inputs = {s.split("=", 1)[0]: _parse(s.split("=", 1)[1]) for s in z["inputs"]}
    inputs = {k: v for k, v in inputs.items() if not k.startswith("particles.")}
    p = Pyro("compressible")
    p.initialize_problem(problem, inputs_dict=dict(inputs, **{"driver.max_steps": 100000, "driver.verbose": 0}))
    while not p.sim.finished():
        p.single_step()
    p.sim.check_state()
    assert p.sim.n == int(z["n"]) and p.sim.cc_data.t == pytest.approx(float(z["t"]), rel=1e-12)
    stride = int(z["stride"])
    for name, stored in zip(z["names"], z["stored"]):
        got = p.sim.cc_data.get_var(str(name)).v().numpy()[::stride, ::stride]
        assert np.abs(got - stored).max() <= 1e-10 * np.abs(stored).max(), name

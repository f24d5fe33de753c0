"""Further stock problems through the public Pyro API against reference-generated fixtures: the second and the
multi-mode Rayleigh-Taylor setups, double Mach reflection with its "ramp" boundaries (device slice assignments
plus host-evaluated top rows), the low-Mach HLLC solver (RIEMANN = 2 instantiations of the sweep), SphericalPolar
grids (the SPH instantiation), the Burgers convergence and tophat problems.  Same kernels and code paths as
the rt16 / burgers_test cases in test_gpu_api.py / test_gpu_flow.py; only the initial conditions differ (those
are checked on the CPU in test_capi_and_host.py)."""
import numpy as np
import pytest

from conftest import state_errors
from golden_util import load_comp, load_flow

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["rt2_48", "rt_multimode16", "ramp64", "gresho40_lm", "sedov32_lm", "sedov_sph32", "advect_sph32"])
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


@pytest.mark.parametrize("nx,bc", [(128, ("dirichlet",) * 4), (128, ("dirichlet", "neumann", "periodic", "periodic")),
                                   (256, ("periodic", "periodic", "dirichlet", "dirichlet"))])
def test_blocked_smoother_with_inhomogeneous_boundary_values(nx, bc):
    """inhomogeneous Dirichlet values on levels the temporally blocked smoother handles (n >= 128), also with the other
    direction periodic: a halo cell that is the periodic image of an interior cell must index the boundary values
    with that cell's row / column (found by scripts/fuzz_mg_emulated.py: the tests before it stopped at n = 64)"""
    import torch
    import oracle
    from pyro2_b200.mg_handle import MGHandle
    o = oracle.MG(nx, bc=bc)
    d = MGHandle(nx, bc, 0.0, -1.0, 0.0, 1.0, 0.0, 1.0, 10, 50)
    rng = np.random.default_rng(nx)
    vals = {k: (rng.standard_normal(nx + 2) if b == "dirichlet" else None) for k, b in zip(("xl", "xr", "yl", "yr"), bc)}
    for k, v in vals.items():
        if v is not None:
            o.set_bc_values(k, v)
    d.set_bc_values(**vals)
    L = o.nlevels - 1
    f = rng.standard_normal((nx + 2, nx + 2))
    o.init_zeros()
    o.init_RHS(f)
    d.plane(L, "f").copy_(torch.from_numpy(f))
    o.smooth(L, 7)
    d.smooth(L, 7)
    assert np.array_equal(d.plane(L, "v").cpu().numpy()[1:-1, 1:-1], o.plane(L, "v")[1:-1, 1:-1])
    for _ in range(2):
        for lev in range(L):
            o.plane(lev, "v")[:] = 0.0
        d.zero_coarse()
        o.v_cycle()
        d.vcycle()
        assert np.array_equal(d.plane(L, "v").cpu().numpy(), o.plane(L, "v"))

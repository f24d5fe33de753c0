"""GPU: the Burgers / incompressible solvers (next row 8f #2).  Stage kernels through the C ABI vs the
oracle on random data, and Pyro("incompressible") / Pyro("burgers") runs vs the reference-generated
fixtures.  Every array the explicit stages produce is built from individually rounded operations in the
reference's order and the multigrid projections are bit-identical, so whole runs must reproduce the
reference's state planes BIT FOR BIT (the tolerance north_star allows is 1e-10)."""
import numpy as np
import pytest

from golden_util import load_flow

pytestmark = pytest.mark.gpu


def _fields(nx, ny, ng, seed, count):
    import oracle
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(count):
        a = np.zeros((nx + 2 * ng, ny + 2 * ng))
        a[ng:-ng, ng:-ng] = rng.standard_normal((nx, ny))
        oracle.fill_ghost(a, ng, ("periodic",) * 4)
        out.append(a)
    return out


@pytest.mark.parametrize("limiter", [0, 1, 2])
@pytest.mark.parametrize("nx,ny", [(16, 16), (96, 200), (256, 256)])
def test_flow_stages_bit_exact(limiter, nx, ny):
    import torch
    import oracle
    from pyro2_b200 import ops
    from pyro2_b200.flow_handle import FlowHandle
    from pyro2_b200.mesh import patch
    ng, dt = 4, 0.011
    g = patch.Grid2d(nx, ny, ng=ng, xmax=1.0, ymax=0.7)
    u, v, gx, gy = _fields(nx, ny, ng, 7 * limiter + nx, 4)
    u[ng + 2:ng + 5, :] = 0.0
    v[:, ng + 1:ng + 3] = 0.0
    planes = ops.alloc_planes(4, g.qx, g.qy)
    for k, a in enumerate((u, v, gx, gy)):
        planes[k, :, :g.qy].copy_(torch.from_numpy(a))
    f = FlowHandle(planes, g)
    P = [planes[k, :, :g.qy] for k in range(4)]
    f.interface_states(P[0], P[1], P[2], P[3], dt, limiter)
    f.mac_vels()
    um, vm = oracle.incomp_mac_vels(u, v, gx, gy, ng, g.dx, g.dy, dt, limiter)
    assert np.array_equal(f.plane("u_MAC").cpu().numpy(), um)
    assert np.array_equal(f.plane("v_MAC").cpu().numpy(), vm)
    f.upwind_states()
    ref = oracle.incomp_states(u, v, gx, gy, ng, g.dx, g.dy, dt, limiter, um, vm)
    for name, r in zip(("u_xint", "v_xint", "u_yint", "v_yint"), ref):
        assert np.array_equal(f.plane(name).cpu().numpy(), r), name
    assert f.maxabs(P[0], P[1]) == (np.abs(u).max(), np.abs(v).max())
    # Burgers update from the same states (no pressure gradient)
    f.interface_states(P[0], P[1], None, None, dt, limiter)
    f.mac_vels()
    f.burgers_update(P[0], P[1], dt)
    ou, ov = oracle.burgers_evolve(u, v, ng, g.dx, g.dy, dt, limiter)
    assert np.array_equal(P[0].cpu().numpy(), ou) and np.array_equal(P[1].cpu().numpy(), ov)


def _run(solver, fname):
    from pyro2_b200.pyro_sim import Pyro
    z, rp, inputs = load_flow(fname)
    p = Pyro(solver)
    p.initialize_problem(str(z["problem"]), inputs_dict=dict(inputs, **{"driver.max_steps": 100000}))
    sim = p.sim
    g = sim.cc_data.grid
    assert list(sim.cc_data.names) == [str(n) for n in z["names"]] and g.ng == int(z["ng"])
    state = lambda: sim.cc_data.planes[:, :, :g.qy].cpu().numpy()
    return p, sim, z, state


@pytest.mark.parametrize("fname", ["incomp_shear32.npz", "incomp_shear64.npz", "incomp_converge32.npz"])
def test_pyro_incompressible_run_matches_reference(fname):
    p, sim, z, state = _run("incompressible", fname)
    # after initialize + preevolve (initial projection, lagged pressure gradient): bit for bit
    assert np.array_equal(state(), z["P0"])
    dts = []
    for _ in range(len(z["dts"])):
        p.single_step()
        dts.append(sim.dt)
    assert np.array_equal(np.array(dts), z["dts"])
    assert sim.n == int(z["n"]) and sim.cc_data.t == float(z["t"])
    assert np.array_equal(state(), z["P"])


def test_pyro_burgers_run_matches_reference():
    p, sim, z, state = _run("burgers", "burgers_test.npz")
    assert np.array_equal(state(), z["P0"])
    dts = []
    for _ in range(len(z["dts"])):
        p.single_step()
        dts.append(sim.dt)
    assert np.array_equal(np.array(dts), z["dts"])
    assert np.array_equal(state(), z["P"])


@pytest.mark.parametrize("fname", ["advection_smooth64.npz", "advection_tophat32.npz"])
def test_pyro_advection_run_matches_reference(fname):
    """smooth64 = BASELINE config 1: 81 steps to t = 1, sum(density) = 4.310466040637315e+03 in the reference"""
    p, sim, z, state = _run("advection", fname)
    assert np.array_equal(state(), z["P0"])
    dts = []
    while not sim.finished() and len(dts) < len(z["dts"]):
        p.single_step()
        dts.append(sim.dt)
    assert np.array_equal(np.array(dts), z["dts"])
    g = sim.cc_data.grid
    v = (slice(g.ilo, g.ihi + 1), slice(g.jlo, g.jhi + 1))
    d = state()[0]
    assert np.array_equal(d[v], z["P"][0][v])
    if fname == "advection_smooth64.npz":
        assert sim.n == 81 and sim.cc_data.t == 1.0
        assert float(np.sum(d[v])) == 4.310466040637315e+03
        assert d[v].min() == 0.9999998946441166


@pytest.mark.parametrize("fname", ["diffusion_gaussian64.npz", "diffusion_gaussian32_mixed.npz"])
def test_pyro_diffusion_run_matches_reference(fname):
    """one Crank-Nicolson multigrid solve per step; the hierarchy is reused with beta following dt"""
    p, sim, z, state = _run("diffusion", fname)
    assert np.array_equal(state(), z["P0"])
    dts = []
    for _ in range(len(z["dts"])):
        p.single_step()
        dts.append(sim.dt)
    assert np.array_equal(np.array(dts), z["dts"])
    assert np.array_equal(state()[0][1:-1, 1:-1], z["P"][0][1:-1, 1:-1])


def test_diffusion_gaussian_follows_analytic_solution():
    from pyro2_b200.diffusion.problems.gaussian import phi_analytic
    from pyro2_b200.pyro_sim import Pyro
    p = Pyro("diffusion")
    p.initialize_problem("gaussian", inputs_dict={"mesh.nx": 128, "mesh.ny": 128, "driver.tmax": 0.002})
    p.run_sim()
    sim = p.sim
    g = sim.cc_data.grid
    x, y = np.meshgrid(g.x, g.y, indexing="ij")
    exact = phi_analytic(np.sqrt((x - 0.5) ** 2 + (y - 0.5) ** 2), sim.cc_data.t, 1.e-4, 1.0, 1.0, 2.0)
    phi = sim.cc_data.get_var("phi").numpy()
    v = (slice(1, -1), slice(1, -1))
    assert np.abs(phi[v] - exact[v]).max() < 5e-3 * (exact[v].max() - 1.0) + 2e-4


@pytest.mark.parametrize("fname", ["lm_bubble32.npz", "lm_bubble64_lim1.npz"])
def test_pyro_lm_atm_run_matches_reference(fname):
    """the low Mach number atmosphere solver (third multigrid caller, variable-coefficient projections): state
    after initialize + preevolve, every dt and all eight planes after the run, bit for bit"""
    p, sim, z, state = _run("lm_atm", fname)
    assert np.array_equal(np.stack([sim.base[k].d for k in ("rho0", "p0", "beta0", "beta0-edges")]), z["base"])
    assert np.array_equal(state(), z["P0"])
    dts = []
    for _ in range(len(z["dts"])):
        p.single_step()
        dts.append(sim.dt)
    assert np.array_equal(np.array(dts), z["dts"])
    assert np.array_equal(state(), z["P"])


def test_lm_atm_bubble_rises():
    from pyro2_b200.pyro_sim import Pyro
    p = Pyro("lm_atm")
    p.initialize_problem("bubble", inputs_dict={"mesh.nx": 64, "mesh.ny": 64, "driver.max_steps": 12})
    p.run_sim()
    sim = p.sim
    v = sim.cc_data.get_var("y-velocity").v()
    assert sim.n == 12 and float(v.max()) > 1e-3          # the light bubble accelerates upward
    rho = sim.cc_data.get_var("density").v()
    assert float(rho.min()) > 0.0


def test_incompressible_projection_leaves_divergence_free_field():
    """after a step the cell-centred divergence of (u, v) is at the level the projection tolerance allows"""
    from pyro2_b200.pyro_sim import Pyro
    p = Pyro("incompressible")
    p.initialize_problem("shear", inputs_dict={"mesh.nx": 128, "mesh.ny": 128, "driver.max_steps": 3})
    p.run_sim()
    sim = p.sim
    g = sim.cc_data.grid
    u, v = sim.cc_data.get_var("x-velocity"), sim.cc_data.get_var("y-velocity")
    div = 0.5 * (u.ip(1) - u.ip(-1)) / g.dx + 0.5 * (v.jp(1) - v.jp(-1)) / g.dy
    assert sim.n == 3
    # approximate projection: the centred divergence is O(h^2) small, not round-off
    assert float(div.abs().max()) < 0.5
    assert float(u.v().abs().max()) < 1.01


def test_incompressible_rejects_unsupported_boundaries():
    from pyro2_b200.pyro_sim import Pyro
    p = Pyro("incompressible")
    with pytest.raises((SystemExit, RuntimeError)):
        p.initialize_problem("shear", inputs_dict={"mesh.nx": 32, "mesh.ny": 32, "mesh.xlboundary": "outflow",
                                                   "mesh.xrboundary": "outflow"})

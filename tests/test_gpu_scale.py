"""GPU parity at benchmark scale (SURVEY.md 8(d) gate): the public API against the CPU oracle on grids one or two
refinements below the 4096^2 benchmark -- as large as the oracle finishes in seconds on the GPU box's host cores.

    Sedov (pyro/compressible, HLLC, limiter 2, flattening, outflow) 1024^2 after 1, 10 and 100 driver steps and 2048^2
    after 1: every conserved variable within 1e-10 relative L2 of the oracle (north_star's tolerance; observed ~1e-15),
    the first dt bit-exact, later dts within 1e-12 of the oracle's (states differ at round-off), and the CFL reduction
    itself bit-exact on the device's own state.
    Multigrid 2048^2 Dirichlet (pyro/multigrid/examples/mg_test_simple.py:32-36): v bit-identical after each of the
    first 7 V-cycles, then a full solve with equal cycle count and bit-identical solution.
"""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,checkpoints", [(1024, (1, 10, 100)), (2048, (1,))])
def test_sedov_through_the_driver_at_scale(n, checkpoints):
    import oracle
    from pyro2_b200.pyro_sim import Pyro
    p = Pyro("compressible")
    p.initialize_problem("sedov", inputs_dict={"mesh.nx": n, "mesh.ny": n, "driver.max_steps": 10 ** 6, "driver.tmax": 1e9})
    sim = p.sim
    g = sim.cc_data.grid
    U = sim.cc_data.data.numpy().copy()
    prm = oracle.comp_params()
    bc = ("outflow",) * 4
    v = (slice(g.ilo, g.ihi + 1), slice(g.jlo, g.jhi + 1))
    dt_old = None
    for step in range(max(checkpoints)):
        for k in range(4):
            pl = np.ascontiguousarray(U[:, :, k])
            oracle.fill_ghost(pl, g.ng, bc)
            U[:, :, k] = pl
        # the driver's limits on the method's dt (pyro/simulation_null.py:222-244)
        dt_o = oracle.cfl_dt(U, g.ng, g.dx, g.dy, 1.4, 0.8)
        dt_o = 0.01 * dt_o if step == 0 else min(2.0 * dt_old, dt_o)
        dt_old = dt_o
        p.single_step()
        if step == 0:
            assert sim.dt == dt_o                                   # identical states: identical bits
        assert abs(sim.dt - dt_o) <= 1e-12 * dt_o
        U = oracle.compressible_step(U, g.ng, g.dx, g.dy, sim.dt, prm)
        if step + 1 in checkpoints:
            sim.check_state()
            got = sim.cc_data.data.numpy()
            for k in range(4):
                ref = U[v][..., k]
                scale = np.linalg.norm(ref.ravel()) if k < 2 else np.linalg.norm(U[v][..., 1].ravel())
                err = np.linalg.norm((got[v][..., k] - ref).ravel()) / scale
                assert err < 1e-10, (step + 1, k, err)
            # the fused wave-speed maxima of the sweep give, bit for bit, the CFL step of the device's own state
            sim.cc_data.fill_BC_all()
            sim.method_compute_timestep()
            dev_state = sim.cc_data.data.numpy()
            assert sim.dt == oracle.cfl_dt(np.ascontiguousarray(dev_state), g.ng, g.dx, g.dy, 1.4, 0.8)


def test_multigrid_2048_cycle_by_cycle(n=2048):
    import torch
    import oracle
    from pyro2_b200.multigrid import MG
    a = MG.CellCenterMG2d(n, n)
    x, y = a.x2d.t(), a.y2d.t()
    f = -2.0 * ((1.0 - 6.0 * x ** 2) * y ** 2 * (1.0 - y ** 2) + (1.0 - 6.0 * y ** 2) * x ** 2 * (1.0 - x ** 2))
    a.init_zeros()
    a.init_RHS(f)
    o = oracle.MG(n)
    o.init_zeros()
    o.init_RHS(f.cpu().numpy())
    assert abs(a.source_norm - o.source_norm) <= 1e-13 * o.source_norm
    fine = a.nlevels - 1
    for cycle in range(7):
        a._h.zero_coarse()
        a.v_cycle(fine)
        for lvl in range(fine):
            o.plane(lvl, "v")[:] = 0.0
        o.v_cycle()
        torch.cuda.synchronize()
        got = a.grids[fine].get_var("v").numpy()
        assert np.array_equal(got[1:-1, 1:-1], o.plane(fine, "v")[1:-1, 1:-1]), cycle + 1
    # and the whole solve (device-side stopping rule, cycles enqueued ahead): same count, same bits
    a.init_zeros()
    o.init_zeros()
    a.solve(rtol=1.e-11)
    o.solve(rtol=1.e-11)
    assert a.num_cycles == o.num_cycles
    assert np.array_equal(a.get_solution().numpy()[1:-1, 1:-1], o.get_solution()[1:-1, 1:-1])
    assert abs(a.residual_error - o.residual_error) <= 1e-10 * o.residual_error

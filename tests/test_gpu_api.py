"""GPU: the drop-in boundary.  The reference's public API (Pyro / Simulation.evolve /
CellCenterMG2d.solve / CellCenterData2d.fill_BC) driven exactly as a pyro user drives it, compared
with fixtures produced by the unmodified reference (tests/golden) and with the reference's own
unit-test assertions.

Tolerances: compressible state 1e-10 relative L2 per variable after tens of steps (north_star),
per-step dt 1e-12 relative (first dt bit-exact); multigrid solutions bit-identical, cycle counts
equal; ghost fill bit-exact (int and float)."""
import os

import numpy as np
import pytest

from conftest import rel_l2, state_errors
from golden_util import GOLDEN, load_comp, load_mg, load_mgvc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["sedov64", "quad64", "sod_x", "kh32", "acoustic64", "advect32", "gresho40",
                                  "bubble32", "rt16", "hse16", "rt16_reflect", "sedov32_cgf", "quad32_cgf_walls",
                                  "heating32", "plume32", "convection16"])
def test_pyro_compressible_run_matches_reference(name):
    from pyro2_b200.pyro_sim import Pyro
    z, rp, inputs = load_comp(name)
    p = Pyro("compressible")
    p.initialize_problem(str(z["problem"]), inputs_dict=dict(inputs, **{"driver.max_steps": 100000}))
    sim = p.sim
    g = sim.cc_data.grid
    assert g.ng == int(z["ng"])
    # same initial state as the reference's problem setup, bit for bit
    U0 = sim.cc_data.data.numpy()
    v = (slice(g.ilo, g.ihi + 1), slice(g.jlo, g.jhi + 1))
    assert np.array_equal(U0[v], z["U0"][v])
    dts = []
    for _ in range(len(z["dts"])):
        p.single_step()
        dts.append(sim.dt)
    sim.check_state()
    assert sim.n == int(z["n"]) and sim.cc_data.t == pytest.approx(float(z["t"]), rel=1e-12)
    assert dts[0] == z["dts"][0]
    assert np.allclose(dts, z["dts"], rtol=1e-11, atol=0)
    U = sim.cc_data.data.numpy()
    errs = state_errors(U[v], z["U"][v], rp["eos.gamma"])
    assert max(errs) < 1e-10, errs


def test_pyro_run_sim_and_accessors():
    from pyro2_b200.pyro_sim import Pyro
    p = Pyro("compressible")
    p.initialize_problem("sedov", inputs_dict={"mesh.nx": 32, "mesh.ny": 32, "sedov.r_init": 0.1,
                                                "driver.max_steps": 7})
    p.run_sim()
    assert p.sim.n == 7 and p.sim.finished()
    dens = p.get_var("density")
    assert dens.shape == (40, 40) and float(dens.v().min()) > 0
    assert p.get_grid().nx == 32 and "compressible" in repr(p)
    # total mass is conserved to round-off while nothing has reached the outflow boundary
    assert float(dens.v().sum()) == pytest.approx(32 * 32, rel=1e-13)


def test_compressible_unit_assertions():
    """pyro/compressible/tests/test_compressible.py:40-62: rho = 1, E = 2.5 -> p = 1, cs = sqrt(gamma);
    cons -> prim -> cons round trip is exact"""
    import torch
    from pyro2_b200.compressible import simulation as sim
    from pyro2_b200.pyro_sim import Pyro

    def init(my_data, rp):
        my_data.get_var("density")[:, :] = 1.0
        my_data.get_var("energy")[:, :] = 2.5
        my_data.get_var("x-momentum")[:, :] = 0.0
        my_data.get_var("y-momentum")[:, :] = 0.0

    p = Pyro("compressible")
    p.add_problem("test", init)
    p.initialize_problem("test", inputs_dict={"mesh.nx": 8, "mesh.ny": 8})
    s = p.sim
    gamma = s.cc_data.get_aux("gamma")
    q = sim.cons_to_prim(s.cc_data.data, gamma, s.ivars, s.cc_data.grid)
    assert float(q[:, :, s.ivars.ip].min()) == pytest.approx(1.0)
    U = sim.prim_to_cons(q, gamma, s.ivars, s.cc_data.grid)
    assert torch.equal(U.t(), s.cc_data.data.t())
    cs = s.cc_data.get_var("soundspeed")
    assert bool((cs.t() == np.sqrt(gamma)).all())
    # a uniform state at rest stays exactly uniform through the sweep
    for _ in range(3):
        p.single_step()
    assert float(s.cc_data.get_var("density").v().min()) == 1.0 == float(s.cc_data.get_var("density").v().max())
    assert float(s.cc_data.get_var("energy").v().min()) == 2.5


def test_unsupported_configurations_fail_loudly():
    from pyro2_b200.pyro_sim import Pyro
    for key, val in (("compressible.riemann", "Roe"), ("particles.do_particles", 1), ("mesh.ylboundary", "sliding-wall")):
        p = Pyro("compressible")
        with pytest.raises(SystemExit):
            p.initialize_problem("sedov", inputs_dict={key: val})
    with pytest.raises(SystemExit):
        Pyro("swe")          # a reference solver this build does not provide


@pytest.mark.parametrize("name", ["poisson_dirichlet_64", "poisson_dirichlet_256", "poisson_periodic_64",
                                  "helmholtz_neumann_64", "poisson_mixed_128"])
def test_mg_solve_matches_reference(name):
    from pyro2_b200.multigrid import MG
    z = load_mg(name)
    bc = [str(b) for b in z["bc"]]
    nx = int(z["nx"])
    a = MG.CellCenterMG2d(nx, nx, xl_BC_type=bc[0], xr_BC_type=bc[1], yl_BC_type=bc[2], yr_BC_type=bc[3],
                          alpha=float(z["alpha"]), beta=float(z["beta"]))
    a.init_zeros()
    a.init_RHS(z["f"])
    assert a.source_norm == pytest.approx(float(z["source_norm"]), rel=1e-13)
    a.solve(rtol=float(z["rtol"]))
    assert a.num_cycles == int(z["num_cycles"])
    assert np.array_equal(a.get_solution().numpy(), z["v"])          # bit-identical incl. ghost cells
    assert a.residual_error == pytest.approx(float(z["residual_error"]), rel=1e-9)
    assert a.relative_error == pytest.approx(float(z["relative_error"]), rel=1e-9)


def test_mg_inhomogeneous_dirichlet_matches_reference():
    from pyro2_b200.multigrid import MG
    z = load_mg("poisson_inhom_64")
    nx = int(z["nx"])
    a = MG.CellCenterMG2d(nx, nx, xl_BC=lambda y: y ** 2, xr_BC=lambda y: 1.0 + y,
                          yl_BC=lambda x: x, yr_BC=lambda x: 1.0 + x ** 2)
    a.init_zeros()
    a.init_RHS(z["f"])
    a.solve(rtol=float(z["rtol"]))
    assert a.num_cycles == int(z["num_cycles"])
    assert np.array_equal(a.get_solution().numpy(), z["v"])


def _vc_solver(nx, bc, cbc, coeffs):
    from pyro2_b200.mesh import boundary as bnd
    from pyro2_b200.mesh import patch
    from pyro2_b200.multigrid import variable_coeff_MG as VMG
    g = patch.Grid2d(nx, nx, ng=1)
    d = patch.CellCenterData2d(g)
    bc_c = bnd.BC(xlb=cbc[0], xrb=cbc[1], ylb=cbc[2], yrb=cbc[3])
    d.register_var("c", bc_c)
    d.create()
    c = d.get_var("c")
    c[:, :] = coeffs(g) if callable(coeffs) else coeffs
    return VMG.VarCoeffCCMG2d(nx, nx, xl_BC_type=bc[0], xr_BC_type=bc[1], yl_BC_type=bc[2], yr_BC_type=bc[3],
                              coeffs=c, coeffs_bc=bc_c)


@pytest.mark.parametrize("name", ["dirichlet_64", "periodic_64", "constant_32", "dirichlet_128"])
def test_mg_variable_coeff_solve_matches_reference(name):
    """VarCoeffCCMG2d on the reference's mg_test_vc_{dirichlet,periodic,constant}.py setups"""
    z = load_mgvc(name)
    nx = int(z["nx"])
    a = _vc_solver(nx, [str(b) for b in z["bc"]], [str(b) for b in z["coeffs_bc"]], z["coeffs"])
    assert np.array_equal(a.edge_coeffs[2].x.numpy(), z["ex_coarse"])
    assert np.array_equal(a.edge_coeffs[2].y.numpy(), z["ey_coarse"])
    a.init_zeros()
    a.init_RHS(z["f"])
    assert a.source_norm == pytest.approx(float(z["source_norm"]), rel=1e-13)
    a.solve(rtol=float(z["rtol"]))
    assert a.num_cycles == int(z["num_cycles"])
    assert np.array_equal(a.get_solution().numpy(), z["v"])
    assert a.residual_error == pytest.approx(float(z["residual_error"]), rel=1e-9)
    assert a.relative_error == pytest.approx(float(z["relative_error"]), rel=1e-9)


def test_mg_variable_coeff_converges_second_order():
    """mg_test_vc_dirichlet.py: alpha = 2 + cos(2 pi x) cos(2 pi y), exact phi = sin(2 pi x) sin(2 pi y)"""
    import torch
    pi = np.pi
    errs = []
    for nx in (64, 128, 256):
        a = _vc_solver(nx, ("dirichlet",) * 4, ("neumann",) * 4,
                       lambda g: 2.0 + torch.cos(2 * pi * g.x2d) * torch.cos(2 * pi * g.y2d))
        x, y = a.x2d, a.y2d
        a.init_zeros()
        a.init_RHS(-16.0 * pi ** 2 * (torch.cos(2 * pi * x) * torch.cos(2 * pi * y) + 1) *
                   torch.sin(2 * pi * x) * torch.sin(2 * pi * y))
        a.solve(rtol=1.e-11)
        e = a.get_solution() - torch.sin(2 * pi * x) * torch.sin(2 * pi * y)
        errs.append(e.norm())
    assert errs[0] / errs[1] == pytest.approx(4.0, rel=0.05)
    assert errs[1] / errs[2] == pytest.approx(4.0, rel=0.05)


def test_mg_variable_coeff_argument_errors():
    from pyro2_b200.mesh import boundary as bnd
    from pyro2_b200.multigrid import variable_coeff_MG as VMG
    with pytest.raises(ValueError):
        VMG.VarCoeffCCMG2d(32, 32)
    with pytest.raises(IndexError):
        VMG.VarCoeffCCMG2d(64, 64, coeffs=np.ones((34, 34)),
                           coeffs_bc=bnd.BC(xlb="neumann", xrb="neumann", ylb="neumann", yrb="neumann"))


def test_mg_gradient_known_answer():
    """pyro/multigrid/tests/test_multigrid_comps.py:39-57"""
    from pyro2_b200.multigrid import MG
    kat = np.load(os.path.join(GOLDEN, "ref_kats.npz"))
    a = MG.CellCenterMG2d(8, 8, ng=1, xmax=8, ymax=8)
    s = a.soln_grid.scratch_array()
    s.v()[:, :] = np.fromfunction(lambda i, j: i * (s.g.nx - i - 1) * j * (s.g.ny - j - 1), (s.g.nx, s.g.ny))
    a.init_solution(s)
    a.grids[a.nlevels - 1].fill_BC("v")
    gx, gy = a.get_solution_gradient()
    assert np.array_equal(gx.numpy()[:, gx.g.jc], kat["mg_gradient_row"])
    assert np.array_equal(gy.numpy()[gx.g.ic, :], kat["mg_gradient_row"])


def test_mg_subclass_hooks_are_used():
    """VarCoeffCCMG2d / GeneralMG2d override smooth and _compute_residual: the
    hooks must stay live.  A subclass that delegates to the stock hooks must reproduce the stock
    solve exactly while being called once per level visit."""
    from pyro2_b200.multigrid import MG
    z = load_mg("poisson_dirichlet_64")
    calls = {"smooth": 0, "resid": 0}

    class Sub(MG.CellCenterMG2d):
        def smooth(self, level, nsmooth):
            calls["smooth"] += 1
            super().smooth(level, nsmooth)

        def _compute_residual(self, level):
            calls["resid"] += 1
            super()._compute_residual(level)

    a = Sub(64, 64)
    a.init_zeros()
    a.init_RHS(z["f"])
    a.solve(rtol=float(z["rtol"]))
    assert a.num_cycles == int(z["num_cycles"])
    assert np.array_equal(a.get_solution().numpy(), z["v"])
    assert calls["smooth"] == a.num_cycles * (2 * (a.nlevels - 1) + 1)
    assert calls["resid"] == a.num_cycles * (a.nlevels - 1 + 1)


def test_mg_argument_errors():
    from pyro2_b200.multigrid import MG
    with pytest.raises(ValueError):
        MG.CellCenterMG2d(8, 16)
    with pytest.raises(ValueError):
        MG.CellCenterMG2d(8, 8, xmax=2.0)
    with pytest.raises(ValueError):
        MG.CellCenterMG2d(12, 12)
    a = MG.CellCenterMG2d(8, 8)
    with pytest.raises(SystemExit):
        a.solve()                      # RHS not initialised (MG.py:640-641)


def test_cellcenterdata_fill_bc_matches_reference_fixture():
    """integer-dtype ghost fill for every standard type (pyro/mesh/tests/test_patch.py:249-337 style)"""
    from pyro2_b200.mesh import boundary as bnd
    from pyro2_b200.mesh import patch
    z = np.load(os.path.join(GOLDEN, "mesh_bcs.npz"))
    for ng in (1, 4):
        g = patch.Grid2d(6, 5, ng=ng)
        for t in ("outflow", "periodic", "reflect-even", "reflect-odd"):
            for dtype in (np.int64, np.float64):
                d = patch.CellCenterData2d(g, dtype=dtype)
                d.register_var("a", bnd.BC(xlb=t, xrb=t, ylb=t, yrb=t))
                d.create()
                d.get_var("a")[:, :] = z[f"base_ng{ng}"].astype(dtype)
                d.fill_BC("a")
                assert np.array_equal(d.get_var("a").numpy(), z[f"{t}_ng{ng}"].astype(dtype))
                d.get_var("a")[:, :] = z[f"base_ng{ng}"].astype(dtype)
                d.fill_BC_all()
                assert np.array_equal(d.get_var("a").numpy(), z[f"{t}_ng{ng}"].astype(dtype))


def test_user_defined_bc_callback_runs_after_standard_fill():
    from pyro2_b200.mesh import boundary as bnd
    from pyro2_b200.mesh import patch
    seen = []

    def user(bc_name, bc_edge, variable, ccdata):
        seen.append((bc_name, bc_edge, variable))
        ccdata.get_var(variable)[:ccdata.grid.ilo, :] = -7.0

    bnd.define_bc("mine", user, is_solid=False)
    try:
        g = patch.Grid2d(4, 4, ng=2)
        d = patch.CellCenterData2d(g)
        d.register_var("a", bnd.BC(xlb="mine", xrb="outflow", ylb="outflow", yrb="outflow"))
        d.create()
        d.get_var("a")[:, :] = 1.0
        d.fill_BC("a")
        assert seen == [("mine", "xlb", "a")]
        a = d.get_var("a").numpy()
        assert (a[:2, :] == -7.0).all() and (a[2:, :] == 1.0).all()
    finally:
        del bnd.ext_bcs["mine"], bnd.bc_solid["mine"]


@pytest.mark.parametrize("nx,ny,nchunks,problem", [(256, 128, 7, "sedov"), (96, 200, 16, "sedov"), (128, 64, 1, "quad")])
def test_streamed_steps_are_bit_identical_to_resident_steps(nx, ny, nchunks, problem, nsteps=12):
    """Pyro.single_step_streamed(): the state lives in pinned host memory, blocks of rows travel host -> device -> host
    while the neighbouring blocks are swept.  Every dt and the final state must equal the resident run's bit for bit
    (blocks are x-slabs: interior block faces get the artificial viscosity, the global +x face does not)."""
    import torch
    from pyro2_b200.pyro_sim import Pyro
    inputs = {"mesh.nx": nx, "mesh.ny": ny, "driver.max_steps": 10 ** 6, "driver.tmax": 1e9}

    def make():
        p = Pyro("compressible")
        p.initialize_problem(problem, inputs_dict=inputs)
        return p
    ref, p = make(), make()
    planes = p.sim.cc_data.planes
    bufs = [torch.empty(planes.shape, dtype=planes.dtype, pin_memory=True) for _ in range(2)]
    bufs[0].copy_(planes)
    dts_ref, dts = [], []
    for step in range(nsteps):
        ref.single_step()
        dts_ref.append(ref.sim.dt)
        p.single_step_streamed(bufs[step % 2], bufs[(step + 1) % 2], nchunks=nchunks)
        dts.append(p.sim.dt)
    ref.sim.check_state()
    p.sim.check_state()
    torch.cuda.synchronize()
    g = p.sim.cc_data.grid
    assert dts == dts_ref
    host = bufs[nsteps % 2][:, g.ilo:g.ihi + 1, g.jlo:g.jhi + 1]
    want = ref.sim.cc_data.planes[:, g.ilo:g.ihi + 1, g.jlo:g.jhi + 1].cpu()
    assert torch.equal(host, want)
    # and a step fed from a buffer the previous step did not write takes the reduction path, same bits
    other = torch.empty(planes.shape, dtype=planes.dtype, pin_memory=True)
    other.copy_(bufs[nsteps % 2])
    ref.single_step()
    p.single_step_streamed(other, bufs[1], nchunks=nchunks)
    torch.cuda.synchronize()
    assert p.sim.dt == ref.sim.dt
    assert torch.equal(bufs[1][:, g.ilo:g.ihi + 1, g.jlo:g.jhi + 1], ref.sim.cc_data.planes[:, g.ilo:g.ihi + 1, g.jlo:g.jhi + 1].cpu())


def compressible_slabs_in_one_process(size, problem, nx, ny, nsteps):
    """a decomposed Pyro("compressible") run whose slabs live in ONE process (a host thread and a stream each; halo rows
    and the wave-speed reduction through the peer-memory transport of csrc/slab_comm.cu over plain pointers) against the
    single-domain run.  Shared by the GPU test below and the emulated-device CPU test."""
    import contextlib
    import threading

    import torch
    from pyro2_b200.parallel import LocalSlabGroup
    from pyro2_b200.pyro_sim import Pyro
    inputs = {"mesh.nx": nx, "mesh.ny": ny, "driver.max_steps": 10 ** 6, "driver.tmax": 1e9}
    # the single-domain run first: it also loads every kernel (a lazily loaded kernel must not meet a spinning one)
    s = Pyro("compressible")
    s.initialize_problem(problem, inputs_dict=inputs)
    dts1 = []
    for _ in range(nsteps):
        s.single_step()
        dts1.append(s.sim.dt)
    s.sim.check_state()
    g1 = s.sim.cc_data.grid
    one = s.sim.cc_data.planes[:, g1.ilo:g1.ihi + 1, g1.jlo:g1.jhi + 1].cpu().numpy().copy()
    group = LocalSlabGroup(size)
    out, errs = [None] * size, []

    def run(rank):
        try:
            stream = torch.cuda.Stream() if torch.cuda.is_available() else None
            with (torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext()):
                p = Pyro("compressible")
                p.initialize_problem(problem, inputs_dict=inputs, decomposition=group.member(rank))
                dts = []
                for _ in range(nsteps):
                    p.single_step()
                    dts.append(p.sim.dt)
                p.sim.check_state()
                p.sim.decomposition.check_peer()
                g = p.sim.cc_data.grid
                out[rank] = (p.sim.cc_data.planes[:, g.ilo:g.ihi + 1, g.jlo:g.jhi + 1].cpu().numpy().copy(), dts)
        except Exception as exc:   # pylint: disable=broad-except
            import traceback
            errs.append(traceback.format_exc() + repr(exc))
            group._barrier.abort()
    threads = [threading.Thread(target=run, args=(r,)) for r in range(size)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(600)
    assert not errs, errs
    assert all(o[1] == dts1 for o in out)
    return np.concatenate([o[0] for o in out], axis=1), one


@pytest.mark.parametrize("size,problem,nx,ny,nsteps", [(2, "sedov", 128, 64, 12), (2, "kh", 64, 48, 8), (4, "quad", 128, 32, 8)])
def test_compressible_slabs_sharing_one_gpu_are_bit_identical(size, problem, nx, ny, nsteps):
    """HP-1 on x-slabs with the peer-memory transport, on ONE device (the driver's GPU-test box has one): each slab a host
    thread + stream; state and every dt identical to the single-domain run (kh: periodic x, both neighbours the same slab)"""
    full, one = compressible_slabs_in_one_process(size, problem, nx, ny, nsteps)
    assert np.array_equal(full, one)

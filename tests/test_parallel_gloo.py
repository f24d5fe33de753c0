"""CPU, world_size 2 and 3 over gloo: the slab halo exchange (pyro2_b200/parallel.py) reproduces the
single-domain ghost rows, including the periodic wrap and the 2-rank case where both neighbours are
the same peer; allreduce_max_ gives the global wave-speed maxima."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, size, port, periodic, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        from pyro2_b200.parallel import SlabDecomposition
        ng, nxl, ny, nvar = 4, 6, 5, 3
        pitch = 16
        nxg = nxl * size
        rng = np.random.default_rng(123)
        G = rng.standard_normal((nvar, nxg, pitch))          # global valid rows
        d = SlabDecomposition()
        assert (d.rank, d.size) == (rank, size)
        assert d.local_nx(nxg) == nxl and d.ioffset(nxg) == rank * nxl
        planes = torch.full((nvar, nxl + 2 * ng, pitch), float("nan"), dtype=torch.float64)
        planes[:, ng:ng + nxl] = torch.from_numpy(G[:, rank * nxl:(rank + 1) * nxl])
        d.exchange(planes, nxl, ng, periodic=periodic)
        got = planes.numpy()
        ok = True
        lo_int, hi_int = d.interior_sides(periodic)
        if lo_int:
            src = (np.arange(rank * nxl - ng, rank * nxl)) % nxg
            ok &= np.array_equal(got[:, :ng], G[:, src])
        else:
            ok &= bool(np.isnan(got[:, :ng]).all())           # physical side: left for the BC fill
        if hi_int:
            src = (np.arange((rank + 1) * nxl, (rank + 1) * nxl + ng)) % nxg
            ok &= np.array_equal(got[:, ng + nxl:], G[:, src])
        else:
            ok &= bool(np.isnan(got[:, ng + nxl:]).all())
        ok &= np.array_equal(got[:, ng:ng + nxl], G[:, rank * nxl:(rank + 1) * nxl])
        w = torch.tensor([1.0 + rank, 10.0 - rank], dtype=torch.float64)
        d.allreduce_max_(w)
        ok &= w.tolist() == [float(size), 10.0]
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("size,periodic", [(2, False), (2, True), (3, True), (3, False)])
def test_slab_exchange(size, periodic):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, size, port, periodic, q)) for r in range(size)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(size))
    assert res == {r: True for r in range(size)}


def test_decomposition_rules():
    from pyro2_b200.parallel import SlabDecomposition
    d = SlabDecomposition(rank=1, size=4)
    assert d.neighbours(False) == (0, 2) and not d.is_first and not d.is_last
    assert SlabDecomposition(rank=0, size=4).neighbours(False) == (None, 1)
    assert SlabDecomposition(rank=3, size=4).neighbours(True) == (2, 0)
    assert SlabDecomposition(rank=0, size=1).neighbours(True) == (None, None)
    with pytest.raises(ValueError):
        d.local_nx(10)
    # slab grids reproduce the global coordinates bit for bit
    from pyro2_b200.mesh.patch import Cartesian2d
    g = Cartesian2d(24, 8, ng=4, xmax=3.0, device="cpu")
    s = Cartesian2d(6, 8, ng=4, xmax=3.0, device="cpu", nx_global=24, ioffset=12)
    assert s.dx == g.dx and np.array_equal(s.x[4:10], g.x[16:22]) and np.array_equal(s.xl, g.xl[12:26])
    # a slab must hold at least ng rows: its neighbour's halo comes from it alone
    from pyro2_b200.simulation_null import grid_setup

    class RP:
        def get_param(self, k):
            return {"mesh.nx": 8, "mesh.ny": 8}[k]
    with pytest.raises(ValueError):
        grid_setup(RP(), ng=4, decomposition=d)            # 8 rows on 4 slabs: 2 < ng


def _emulated_run_worker(rank, size, port, problem, nx, ny, nsteps, q):
    """the body of tests/multi_gpu_worker.py on the emulated device over gloo: a decomposed Pyro("compressible") run,
    slabs gathered on rank 0 and compared with the single-domain run bit for bit (state and every dt)"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        import emu_device
        from pyro2_b200.parallel import SlabDecomposition
        from pyro2_b200.pyro_sim import Pyro
        inputs = {"mesh.nx": nx, "mesh.ny": ny, "driver.max_steps": 10 ** 6, "driver.tmax": 1e9, "driver.verbose": 0}
        if problem == "sedov":
            inputs["sedov.r_init"] = 0.15
        with emu_device.emulated_device() as dev:
            p = Pyro("compressible")
            p.initialize_problem(problem, inputs_dict=inputs, decomposition=SlabDecomposition())
            dts = []
            for _ in range(nsteps):
                p.single_step()
                dts.append(p.sim.dt)
            p.sim.check_state()
            g = p.sim.cc_data.grid
            assert g.nx == nx // size and dev.calls["p2b_compressible_sweep"] == nsteps
            # the state's halo rows and the wave-speed reduction went through peer memory (csrc/slab_comm.cu), not gloo
            assert dev.calls.get("p2b_slab_exchange", 0) >= nsteps and dev.calls.get("p2b_slab_allreduce_max4", 0) >= nsteps - 1
            p.sim.decomposition.check_peer()
            mine = p.sim.cc_data.planes[:, g.ilo:g.ihi + 1, g.jlo:g.jhi + 1].contiguous()
            parts = [torch.empty_like(mine) for _ in range(size)] if rank == 0 else None
            dist.gather(mine, parts, dst=0)
            ok = True
            if rank == 0:
                full = torch.cat(parts, dim=1).numpy()
                s = Pyro("compressible")
                s.initialize_problem(problem, inputs_dict=inputs)
                dts1 = []
                for _ in range(nsteps):
                    s.single_step()
                    dts1.append(s.sim.dt)
                g1 = s.sim.cc_data.grid
                one = s.sim.cc_data.planes[:, g1.ilo:g1.ihi + 1, g1.jlo:g1.jhi + 1].numpy()
                ok = bool(np.array_equal(full, one)) and dts == dts1
        dist.barrier()
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("problem,nx,ny,nsteps,size", [("sedov", 32, 16, 4, 2), ("kh", 24, 16, 3, 2), ("quad", 24, 12, 3, 3)][::2] +
                         [pytest.param("kh", 24, 16, 3, 2, marks=pytest.mark.skipif(not os.environ.get("P2B_FULL_TESTS"), reason="long cases"))])
def test_decomposed_run_is_bit_identical_on_emulated_device(problem, nx, ny, nsteps, size):
    """the N > 1 product path end to end without a GPU: SlabDecomposition + halo exchange over gloo, the sweep / ghost
    fill / CFL kernels through the host-compiled libraries (tests/emu_device.py).  periodic (kh) and physical
    (sedov, quad) x boundaries; the all-reduced time step must agree on every rank and with the single-domain run"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_emulated_run_worker, args=(r, size, port, problem, nx, ny, nsteps, q)) for r in range(size)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(size))
    assert res == {r: True for r in range(size)}


def _emulated_mg_worker(rank, size, port, kind, n, split, q):
    """the body of tests/multi_gpu_mg_worker.py on the emulated device over gloo: x-slab multigrid (deep-halo exchange
    per blocked-smoother pass, replicated coarse levels) against the single-domain solve, bit for bit"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [here, os.path.join(os.path.dirname(here), "oracle")]
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        import emu_device
        from multi_gpu_mg_worker import rhs
        from pyro2_b200.multigrid import MG
        from pyro2_b200.parallel import SlabDecomposition
        bc = {"dirichlet": ("dirichlet",) * 4, "periodic": ("periodic",) * 4,
              "mixed": ("neumann", "dirichlet", "dirichlet", "neumann"),
              "xper_inhom": ("periodic", "periodic", "dirichlet", "neumann")}[kind]
        kw = dict(xl_BC_type=bc[0], xr_BC_type=bc[1], yl_BC_type=bc[2], yr_BC_type=bc[3])
        if kind == "mixed":
            kw.update(alpha=1.0, beta=0.05)
        if kind == "xper_inhom":
            # inhomogeneous values along the y sides, indexed by the GLOBAL row: the halo rows a slab receives across the
            # periodic x boundary must wrap that index
            kw.update(yl_BC=lambda s: 0.3 + np.sin(2.0 * np.pi * s), yr_BC=lambda s: np.cos(4.0 * np.pi * s))
        with emu_device.emulated_device():
            a = MG.CellCenterMG2d(n, n, decomposition=SlabDecomposition(), split_n=split, **kw)
            a.init_zeros()
            a.init_RHS(rhs(kind, a.x2d.t(), a.y2d.t()))
            a.solve(rtol=1.e-11)
            g = a.soln_grid
            mine = a.get_solution().t()[g.ilo:g.ihi + 1, g.jlo:g.jhi + 1].contiguous()
            parts = [torch.empty_like(mine) for _ in range(size)] if rank == 0 else None
            dist.gather(mine, parts, dst=0)
            ok = True
            if rank == 0:
                full = torch.cat(parts, dim=0).numpy()
                b = MG.CellCenterMG2d(n, n, **kw)
                b.init_zeros()
                b.init_RHS(rhs(kind, b.x2d.t(), b.y2d.t()))
                b.solve(rtol=1.e-11)
                one = b.get_solution().numpy()[1:-1, 1:-1]
                ok = bool(np.array_equal(full, one)) and a.num_cycles == b.num_cycles
        dist.barrier()
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


# the periodic and Neumann / Helmholtz hierarchies are also exercised by the decomposed incompressible and diffusion
# runs below; their stand-alone cases run with P2B_FULL_TESTS=1
_FULL = pytest.mark.skipif(not os.environ.get("P2B_FULL_TESTS"), reason="set P2B_FULL_TESTS=1 for the long cases")


@pytest.mark.parametrize("kind,n,split,size", [("xper_inhom", 128, 32, 2), pytest.param("dirichlet", 128, 32, 2, marks=_FULL),
                                               pytest.param("periodic", 128, 64, 2, marks=_FULL),
                                               pytest.param("mixed", 128, 64, 2, marks=_FULL)])
def test_decomposed_multigrid_is_bit_identical_on_emulated_device(kind, n, split, size):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_emulated_mg_worker, args=(r, size, port, kind, n, split, q)) for r in range(size)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(size))
    assert res == {r: True for r in range(size)}


def _emulated_flow_worker(rank, size, port, solver, problem, inputs, nsteps, q):
    """a decomposed Pyro(solver) run of the explicit-stage solvers (advection, burgers) or the incompressible
    solver (explicit stages + two x-slab multigrid projections per step) on the emulated device over gloo; every
    state plane gathered on rank 0 and compared with the single-domain run bit for bit, and every dt"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [here, os.path.join(os.path.dirname(here), "oracle")]
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        import emu_device
        from pyro2_b200.parallel import SlabDecomposition
        from pyro2_b200.pyro_sim import Pyro
        inputs = dict(inputs, **{"driver.max_steps": 10 ** 6, "driver.tmax": 1e9, "driver.verbose": 0})
        with emu_device.emulated_device():
            p = Pyro(solver)
            p.initialize_problem(problem, inputs_dict=inputs, decomposition=SlabDecomposition())
            dts = []
            for _ in range(nsteps):
                p.single_step()
                dts.append(p.sim.dt)
            g = p.sim.cc_data.grid
            assert g.nx == inputs["mesh.nx"] // size
            mine = p.sim.cc_data.planes[:, g.ilo:g.ihi + 1, g.jlo:g.jhi + 1].contiguous()
            parts = [torch.empty_like(mine) for _ in range(size)] if rank == 0 else None
            dist.gather(mine, parts, dst=0)
            ok = True
            if rank == 0:
                full = torch.cat(parts, dim=1).numpy()
                s = Pyro(solver)
                s.initialize_problem(problem, inputs_dict=inputs)
                dts1 = []
                for _ in range(nsteps):
                    s.single_step()
                    dts1.append(s.sim.dt)
                g1 = s.sim.cc_data.grid
                one = s.sim.cc_data.planes[:, g1.ilo:g1.ihi + 1, g1.jlo:g1.jhi + 1].numpy()
                ok = bool(np.array_equal(full, one)) and dts == dts1
                if not ok:
                    print("MISMATCH", solver, problem, np.abs(full - one).max(axis=(1, 2)), dts, dts1, flush=True)
        dist.barrier()
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("solver,problem,inputs,nsteps,size", [
    pytest.param("advection", "smooth", {"mesh.nx": 32, "mesh.ny": 32}, 5, 2, marks=_FULL),
    ("advection", "tophat", {"mesh.nx": 36, "mesh.ny": 24, "advection.u": -0.6, "advection.limiter": 1}, 5, 3),
    ("compressible", "advect", {"mesh.grid_type": "SphericalPolar", "mesh.nx": 24, "mesh.ny": 20, "mesh.xmin": 1.0, "mesh.xmax": 2.0,
                                "mesh.ymin": 0.523, "mesh.ymax": 2.617, "mesh.xlboundary": "reflect", "mesh.xrboundary": "outflow",
                                "mesh.ylboundary": "outflow", "mesh.yrboundary": "outflow", "compressible.riemann": "CGF",
                                "compressible.grav": -0.5, "driver.fix_dt": -1.0}, 3, 3),      # SphericalPolar on slabs
    # user-defined y boundaries: the hse hooks run per slab after all halo rows have arrived
    ("compressible", "bubble", {"mesh.nx": 18, "mesh.ny": 36, "mesh.ymax": 4.0, "mesh.xlboundary": "outflow", "mesh.xrboundary": "outflow",
                                "mesh.ylboundary": "hse", "mesh.yrboundary": "hse"}, 3, 3),
    # periodic x AND a user hook that reads other variables: across the periodic seam the halo rows stand for the single
    # domain's x GHOST rows, which the hse energy sees as the previous fill left them (regression: differed by 2e-21)
    ("compressible", "bubble", {"mesh.nx": 24, "mesh.ny": 36, "mesh.ymax": 4.0, "mesh.xlboundary": "periodic", "mesh.xrboundary": "periodic",
                                "mesh.ylboundary": "hse", "mesh.yrboundary": "hse"}, 6, 3),
    # a setup that places its feature relative to the DOMAIN (regression: the vortex was centred on each slab)
    ("compressible", "gresho", {"mesh.nx": 20, "mesh.ny": 20}, 3, 2),
    pytest.param("compressible", "convection", {"mesh.nx": 16, "mesh.ny": 72}, 3, 4, marks=_FULL),    # ambient top, heating, sponge
    # lm_atm: explicit stages on slabs, the two variable-coefficient projections replicated after all-gathers
    ("lm_atm", "bubble", {"mesh.nx": 32, "mesh.ny": 32}, 1, 2),
    ("burgers", "test", {"mesh.nx": 32, "mesh.ny": 32}, 5, 2),                  # outflow x sides
    pytest.param("burgers", "tophat", {"mesh.nx": 48, "mesh.ny": 32}, 4, 3, marks=_FULL),                # periodic
    ("diffusion", "gaussian", {"mesh.nx": 128, "mesh.ny": 128, "diffusion.mg_split_n": 64}, 2, 2),
    # (initialize_problem already runs the initial projection and one full step in preevolve: no further steps needed)
    ("incompressible", "shear", {"mesh.nx": 128, "mesh.ny": 128, "incompressible.mg_split_n": 64}, 0, 2),
    pytest.param("incompressible", "converge", {"mesh.nx": 128, "mesh.ny": 128, "incompressible.mg_split_n": 32}, 1, 4, marks=_FULL)])
def test_decomposed_flow_solvers_are_bit_identical_on_emulated_device(solver, problem, inputs, nsteps, size):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_emulated_flow_worker, args=(r, size, port, solver, problem, inputs, nsteps, q))
             for r in range(size)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(900)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(size))
    assert res == {r: True for r in range(size)}

"""GPU parity of the multigrid kernels vs the CPU oracle through the C ABI.  The device arithmetic
is unfused and ordered like the reference, so v / f / r planes must be BIT-IDENTICAL; only the
norms (reductions) are compared to 1e-13."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BC_SETS = [("dirichlet",) * 4, ("neumann",) * 4, ("periodic",) * 4,
           ("dirichlet", "dirichlet", "neumann", "neumann"), ("periodic", "periodic", "dirichlet", "neumann")]


def _rhs(n):
    x = (np.arange(n + 2) - 0.5) / n
    X, Y = np.meshgrid(x, x, indexing="ij")
    return -2.0 * ((1.0 - 6.0 * X ** 2) * Y ** 2 * (1.0 - Y ** 2) + (1.0 - 6.0 * Y ** 2) * X ** 2 * (1.0 - X ** 2))


def _pair(nx, bc, alpha=0.0, beta=-1.0, seed=0):
    import torch
    import oracle
    from pyro2_b200.mg_handle import MGHandle
    o = oracle.MG(nx, bc=bc, alpha=alpha, beta=beta)
    d = MGHandle(nx, bc, alpha, beta, 0.0, 1.0, 0.0, 1.0, 10, 50)
    rng = np.random.default_rng(seed)
    L = o.nlevels - 1
    for which in ("v", "f"):
        a = rng.standard_normal((nx + 2, nx + 2))
        o.plane(L, which)[:] = a
        d.plane(L, which).copy_(torch.from_numpy(a))
    return o, d


def _same(o, d, level, which):
    return np.array_equal(d.plane(level, which).cpu().numpy(), o.plane(level, which))


@pytest.mark.parametrize("bc", BC_SETS)
@pytest.mark.parametrize("nx", [2, 4, 16, 64, 128])
def test_smooth_residual_bit_exact(bc, nx):
    o, d = _pair(nx, bc)
    L = o.nlevels - 1
    o.smooth(L, 3); d.smooth(L, 3)
    assert _same(o, d, L, "v")
    o.residual(L); d.residual(L)
    assert np.array_equal(d.plane(L, "r").cpu().numpy()[1:-1, 1:-1], o.plane(L, "r")[1:-1, 1:-1])


@pytest.mark.parametrize("bc", BC_SETS)
@pytest.mark.parametrize("nx", [4, 64, 256])
def test_restrict_prolong_bit_exact(bc, nx):
    import torch
    o, d = _pair(nx, bc, seed=3)
    L = o.nlevels - 1
    o.residual(L); d.residual(L)
    o.restrict(L); d.restrict(L)
    assert np.array_equal(d.plane(L - 1, "f").cpu().numpy()[1:-1, 1:-1], o.plane(L - 1, "f")[1:-1, 1:-1])
    o.smooth(L - 1, 2); d.smooth(L - 1, 2)
    assert _same(o, d, L - 1, "v")
    o.prolong_correct(L); d.prolong_correct(L)
    assert _same(o, d, L, "v")


@pytest.mark.parametrize("bc,alpha,beta", [(("dirichlet",) * 4, 0.0, -1.0), (("periodic",) * 4, 0.0, -1.0),
                                            (("neumann",) * 4, 1.0, 0.01)])
@pytest.mark.parametrize("nx", [8, 128, 512])
def test_vcycle_bit_exact(bc, alpha, beta, nx):
    import torch
    import oracle
    from pyro2_b200.mg_handle import MGHandle
    o = oracle.MG(nx, bc=bc, alpha=alpha, beta=beta)
    d = MGHandle(nx, bc, alpha, beta, 0.0, 1.0, 0.0, 1.0, 10, 50)
    L = o.nlevels - 1
    f = _rhs(nx) if bc[0] != "periodic" else np.sin(2 * np.pi * np.linspace(0, 1, nx + 2))[:, None] * np.ones(nx + 2)[None, :]
    o.init_zeros(); o.init_RHS(f)
    d.plane(L, "f").copy_(torch.from_numpy(f))
    for cyc in range(3):
        # solve() zeroes the coarse v before each cycle (MG.py:658-659)
        for l in range(L):
            o.plane(l, "v")[:] = 0.0
        d.zero_coarse()
        o.v_cycle(); d.vcycle()
        assert _same(o, d, L, "v"), cyc
    o.residual(L)
    old = torch.zeros((nx + 2) * d.plane(L, "v").stride(0), dtype=torch.float64, device="cuda")
    relsq, rsq = d.cycle_diagnostics(old)
    rn = np.sqrt(rsq / nx / nx)
    assert abs(rn - o.norm(o.plane(L, "r"))) <= 1e-13 * max(rn, 1e-300)


def test_inhomogeneous_dirichlet_bit_exact():
    import torch
    import oracle
    from pyro2_b200.mg_handle import MGHandle
    nx = 64
    o = oracle.MG(nx)
    d = MGHandle(nx, ("dirichlet",) * 4, 0.0, -1.0, 0.0, 1.0, 0.0, 1.0, 10, 50)
    c = (np.arange(nx + 2) - 0.5) / nx
    vals = {"xl": c ** 2, "xr": 1.0 + c, "yl": c, "yr": 1.0 + c ** 2}
    for k, v in vals.items():
        o.set_bc_values(k, v)
    d.set_bc_values(**vals)
    L = o.nlevels - 1
    f = _rhs(nx)
    o.init_zeros(); o.init_RHS(f)
    d.plane(L, "f").copy_(torch.from_numpy(f))
    for cyc in range(2):
        for l in range(L):
            o.plane(l, "v")[:] = 0.0
        d.zero_coarse()
        o.v_cycle(); d.vcycle()
        assert _same(o, d, L, "v")


def test_unreachable_tolerance_runs_max_cycles_like_the_reference():
    """SURVEY.md 7 "solve() tolerance floor": when rtol is below the fp64 residual floor the reference
    loops to max_cycles = 100 (MG.py:190,653); the drop-in must reproduce that control flow -- same
    cycle count, same bits after 100 cycles (periodic, singular Poisson as in the incompressible
    projection)."""
    import torch
    import oracle
    from pyro2_b200.multigrid import MG
    nx = 128
    bc = ("periodic",) * 4
    a = MG.CellCenterMG2d(nx, nx, xl_BC_type="periodic", xr_BC_type="periodic", yl_BC_type="periodic",
                          yr_BC_type="periodic")
    x, y = a.x2d.numpy(), a.y2d.numpy()
    f = np.sin(2 * np.pi * x) * np.cos(6 * np.pi * y) + 0.3 * np.cos(4 * np.pi * x)
    a.init_zeros()
    a.init_RHS(f)
    a.solve(rtol=1.e-30)
    o = oracle.MG(nx, bc=bc)
    o.init_zeros()
    o.init_RHS(f)
    o.solve(rtol=1.e-30)
    assert a.num_cycles == o.num_cycles == 100
    assert np.array_equal(a.get_solution().numpy(), o.get_solution())
    assert a.residual_error == pytest.approx(o.residual_error, rel=1e-6)


@pytest.mark.parametrize("nsmooth", [1, 5, 7, 12])
def test_blocked_and_per_colour_smoothers_agree_bitwise(nsmooth):
    """the temporally blocked smoother (odd / even pass counts, partial last pass) against the
    one-launch-per-colour kernels and the oracle"""
    o, d = _pair(256, ("dirichlet", "neumann", "periodic", "periodic"), alpha=0.3, beta=0.02, seed=11)
    o2, e = _pair(256, ("dirichlet", "neumann", "periodic", "periodic"), alpha=0.3, beta=0.02, seed=11)
    L = o.nlevels - 1
    e.set_blocking(False)
    o.smooth(L, nsmooth); d.smooth(L, nsmooth); e.smooth(L, nsmooth)
    assert _same(o, d, L, "v")
    assert np.array_equal(d.plane(L, "v").cpu().numpy(), e.plane(L, "v").cpu().numpy())


# ---- variable coefficients (VarCoeffCCMG2d) ------------------------------------------------------------
def _vc_pair(nx, bc, cbc, seed=0):
    import torch
    o, d = _pair(nx, bc, 0.0, 0.0, seed=seed)
    rng = np.random.default_rng(seed + 100)
    coeffs = 0.5 + rng.random((nx + 2, nx + 2))
    o.set_coeffs(coeffs, cbc)
    d.set_coeffs(torch.from_numpy(coeffs).cuda(), cbc)
    return o, d


VC_SETS = [(("dirichlet",) * 4, ("neumann",) * 4), (("periodic",) * 4, ("periodic",) * 4),
           (("dirichlet", "neumann", "periodic", "periodic"), ("neumann", "reflect-even", "periodic", "periodic"))]


@pytest.mark.parametrize("bc,cbc", VC_SETS)
@pytest.mark.parametrize("nx", [4, 64, 256])
def test_vc_edge_coefficients_bit_exact(bc, cbc, nx):
    o, d = _vc_pair(nx, bc, cbc)
    for lev in range(o.nlevels):
        for a, b in (("c", "c"), ("ex", "x"), ("ey", "y")):
            assert np.array_equal(d.coeff_plane(lev, b).cpu().numpy(), o.coef_plane(lev, a)), (lev, a)


@pytest.mark.parametrize("bc,cbc", VC_SETS)
@pytest.mark.parametrize("nx", [2, 16, 64, 128, 512])
def test_vc_smooth_residual_vcycle_bit_exact(bc, cbc, nx):
    o, d = _vc_pair(nx, bc, cbc, seed=nx)
    L = o.nlevels - 1
    o.smooth(L, 7); d.smooth(L, 7)
    assert _same(o, d, L, "v")
    o.residual(L); d.residual(L)
    assert np.array_equal(d.plane(L, "r").cpu().numpy()[1:-1, 1:-1], o.plane(L, "r")[1:-1, 1:-1])
    o.v_cycle(); d.vcycle()
    assert _same(o, d, L, "v")
    for lev in range(L):
        assert _same(o, d, lev, "v"), lev


def test_vc_cycle_diagnostics():
    import torch
    nx = 256
    o, d = _vc_pair(nx, ("dirichlet",) * 4, ("neumann",) * 4, seed=9)
    L = o.nlevels - 1
    old = d.plane(L, "v").clone()
    pitch = d.info(L)["pitch"]
    old_phi = torch.zeros((nx + 2, pitch), dtype=torch.float64, device="cuda")
    old_phi[:, :nx + 2] = old
    o.v_cycle(); d.vcycle()
    relsq, rsq = d.cycle_diagnostics(old_phi)
    o.residual(L)
    r = o.plane(L, "r")[1:-1, 1:-1]
    assert rsq == pytest.approx(float((r ** 2).sum()), rel=1e-12)
    assert np.array_equal(d.plane(L, "r").cpu().numpy()[1:-1, 1:-1], r)
    v = o.plane(L, "v")[1:-1, 1:-1]
    ref = (((v - old.cpu().numpy()[1:-1, 1:-1]) / (v + 1e-16)) ** 2).sum()
    assert relsq == pytest.approx(float(ref), rel=1e-12)
    assert np.array_equal(old_phi[:, :nx + 2].cpu().numpy()[1:-1, 1:-1], v)


def slabs_in_one_process(size, n, split, kw, rhs, use_graph=False, rtol=1.e-11):
    """solve on `size` x-slabs that live in ONE process (a host thread and a stream each; the slabs reach each other's
    workspaces through plain pointers) and on a single domain; returns (stitched solution, cycles, single, cycles).
    Exercises the peer-memory protocol of the decomposed V-cycle (csrc/mg_kernels.cuh) where only one device exists;
    shared by the GPU test below and by the emulated-device CPU test."""
    import threading

    import torch
    from pyro2_b200.multigrid import MG
    from pyro2_b200.parallel import LocalSlabGroup
    group = LocalSlabGroup(size)
    out, errs = [None] * size, []

    def run(rank):
        try:
            stream = torch.cuda.Stream() if torch.cuda.is_available() else None
            ctx = torch.cuda.stream(stream) if stream is not None else __import__("contextlib").nullcontext()
            with ctx:
                a = MG.CellCenterMG2d(n, n, decomposition=group.member(rank), split_n=split, **kw)
                a.use_graph = use_graph
                a.init_zeros()
                a.init_RHS(rhs(a.x2d.t(), a.y2d.t()))
                a.solve(rtol=rtol)
                g = a.soln_grid
                out[rank] = (a.get_solution().t()[g.ilo:g.ihi + 1, g.jlo:g.jhi + 1].cpu().numpy().copy(), a.num_cycles,
                             a.residual_error)
        except Exception as exc:   # pylint: disable=broad-except
            errs.append(repr(exc))
            group._barrier.abort()
    threads = [threading.Thread(target=run, args=(r,)) for r in range(size)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(600)
    assert not errs, errs
    b = MG.CellCenterMG2d(n, n, **kw)
    b.init_zeros()
    b.init_RHS(rhs(b.x2d.t(), b.y2d.t()))
    b.solve(rtol=rtol)
    assert len({o[1] for o in out}) == 1 and len({o[2] for o in out}) == 1      # every rank stopped at the same cycle
    return np.concatenate([o[0] for o in out], axis=0), out[0][1], b.get_solution().cpu().numpy()[1:-1, 1:-1], b.num_cycles


SLAB_CASES = [("dirichlet", 2, 256, 64), ("periodic", 2, 256, 128), ("mixed", 4, 512, 128), ("xper_inhom", 2, 256, 64)]


def slab_case(kind):
    import torch
    bc = {"dirichlet": ("dirichlet",) * 4, "periodic": ("periodic",) * 4,
          "mixed": ("neumann", "dirichlet", "dirichlet", "neumann"),
          "xper_inhom": ("periodic", "periodic", "dirichlet", "neumann")}[kind]
    kw = dict(xl_BC_type=bc[0], xr_BC_type=bc[1], yl_BC_type=bc[2], yr_BC_type=bc[3])
    if kind == "mixed":
        kw.update(alpha=1.0, beta=0.05)
    if kind == "xper_inhom":
        kw.update(yl_BC=lambda s: 0.3 + np.sin(2.0 * np.pi * s), yr_BC=lambda s: np.cos(4.0 * np.pi * s))

    def rhs(x, y):
        if kind == "periodic":
            return torch.sin(2 * np.pi * x) * torch.cos(4 * np.pi * y)
        return -2.0 * ((1.0 - 6.0 * x ** 2) * y ** 2 * (1.0 - y ** 2) + (1.0 - 6.0 * y ** 2) * x ** 2 * (1.0 - x ** 2))
    return kw, rhs


@pytest.mark.parametrize("kind,size,n,split", SLAB_CASES)
def test_slabs_sharing_one_gpu_are_bit_identical(kind, size, n, split):
    """the decomposed V-cycle's peer-memory protocol on ONE device: each slab a host thread + stream, halo rows pushed
    from the kernels' epilogues, flags, device-side all-reduce -- identical solution bits and cycle count"""
    kw, rhs = slab_case(kind)
    full, cyc, one, cyc1 = slabs_in_one_process(size, n, split, kw, rhs)
    assert cyc == cyc1 and np.array_equal(full, one)

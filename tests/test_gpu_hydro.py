"""GPU parity: ghost fill, CFL dt and the fused compressible sweep vs the CPU oracle, called through
the C ABI (pyro2_b200.ops -> libpyro2b200.so).

Tolerances: ghost fill and dt are bit-exact; the sweep differs from the reference's unfused
arithmetic only by FMA contraction / shared reciprocals, tolerance 1e-12 relative L2 per variable
per step (north_star allows 1e-10).
"""
import numpy as np
import pytest

from conftest import make_state, rel_l2

pytestmark = pytest.mark.gpu

BCS = ["outflow", "reflect-even", "reflect-odd", "periodic"]


def _to_device(U_ijn):
    import torch
    from pyro2_b200 import ops
    qx, qy, nvar = U_ijn.shape
    P = ops.alloc_planes(nvar, qx, qy, dtype=torch.float64 if U_ijn.dtype == np.float64 else torch.int64)
    P[:, :, :qy] = torch.from_numpy(np.ascontiguousarray(np.moveaxis(U_ijn, 2, 0))).cuda()
    return P


def _to_host(P, qy):
    return np.ascontiguousarray(np.moveaxis(P[:, :, :qy].cpu().numpy(), 0, 2))


@pytest.mark.parametrize("dtype", [np.float64, np.int64])
@pytest.mark.parametrize("xbc", BCS)
@pytest.mark.parametrize("ybc", BCS)
def test_fill_ghost_bit_exact(dtype, xbc, ybc):
    import oracle
    from pyro2_b200 import ops
    rng = np.random.default_rng(1)
    nx, ny, ng = 9, 12, 4
    if dtype == np.int64:
        a = rng.integers(-1000, 1000, size=(nx + 2 * ng, ny + 2 * ng, 2)).astype(np.int64)
    else:
        a = rng.standard_normal((nx + 2 * ng, ny + 2 * ng, 2))
    bc = (xbc, xbc, ybc, ybc)
    P = _to_device(a)
    ops.fill_ghost(P, nx, ny, ng, [bc, bc])
    got = _to_host(P, ny + 2 * ng)
    for n in range(2):
        ref = np.ascontiguousarray(a[:, :, n])
        oracle.fill_ghost(ref, ng, bc)
        assert np.array_equal(got[:, :, n], ref)


def test_fill_ghost_mixed_per_variable():
    import oracle
    from pyro2_b200 import ops
    rng = np.random.default_rng(2)
    nx, ny, ng = 16, 8, 4
    a = rng.standard_normal((nx + 2 * ng, ny + 2 * ng, 4))
    bcs = [("outflow", "reflect-even", "reflect-even", "outflow"), ("reflect-odd", "reflect-odd", "periodic", "periodic"),
           ("periodic", "periodic", "reflect-odd", "outflow"), ("outflow", "outflow", "outflow", "outflow")]
    P = _to_device(a)
    ops.fill_ghost(P, nx, ny, ng, bcs)
    got = _to_host(P, ny + 2 * ng)
    for n in range(4):
        ref = np.ascontiguousarray(a[:, :, n])
        oracle.fill_ghost(ref, ng, bcs[n])
        assert np.array_equal(got[:, :, n], ref)


@pytest.mark.parametrize("kind,nx,ny", [("smooth", 64, 48), ("shock", 100, 37), ("sedov", 128, 128)])
def test_cfl_dt_bit_exact(kind, nx, ny):
    import oracle
    from pyro2_b200 import ops
    ng = 4
    U = make_state(nx, ny, ng, kind)
    dx, dy = 1.0 / nx, 1.0 / ny
    P = _to_device(U)
    wx, wy = ops.cfl_wavemax(P, nx, ny, ng, 1.4, ops.new_scratch())
    dt = 0.8 * min(dx / wx, dy / wy)
    assert dt == oracle.cfl_dt(U, ng, dx, dy, 1.4, 0.8)


def _filled(U, ng, bc=("outflow",) * 4):
    import oracle
    U = U.copy()
    for n in range(U.shape[2]):
        pl = np.ascontiguousarray(U[:, :, n])
        oracle.fill_ghost(pl, ng, bc)
        U[:, :, n] = pl
    return U


SWEEP_CASES = [
    ("smooth", 64, 64, 2, 1), ("shock", 64, 70, 2, 1), ("shock", 37, 61, 1, 1), ("shock", 20, 20, 0, 0),
    ("sedov", 96, 96, 2, 1), ("shock", 200, 31, 2, 1), ("shock", 31, 200, 2, 0), ("smooth", 256, 256, 2, 1),
]


@pytest.mark.parametrize("kind,nx,ny,limiter,flat", SWEEP_CASES)
def test_sweep_one_step(kind, nx, ny, limiter, flat):
    import oracle
    from pyro2_b200 import ops
    ng = 4
    U = _filled(make_state(nx, ny, ng, kind), ng)
    dx, dy = 1.0 / nx, 1.0 / ny
    dt = 0.5 * oracle.cfl_dt(U, ng, dx, dy, 1.4, 0.8)
    ref = oracle.compressible_step(U, ng, dx, dy, dt, oracle.comp_params(limiter=limiter, use_flattening=flat))
    Pin = _to_device(U)
    Pout = Pin.clone()
    scratch = ops.new_scratch()
    ops.compressible_sweep(Pin, Pout, nx, ny, ng, dx, dy, dt, ops.comp_params(limiter=limiter, use_flattening=flat), scratch)
    got = _to_host(Pout, ny + 2 * ng)
    v = (slice(ng, ng + nx), slice(ng, ng + ny))
    assert not np.isnan(got[v]).any()
    for n in range(4):
        assert rel_l2(got[v][..., n], ref[v][..., n]) < 1e-12, (n, rel_l2(got[v][..., n], ref[v][..., n]))
    # ghost cells of the output buffer are not written by the sweep
    assert np.array_equal(got[:ng], U[:ng]) and np.array_equal(got[:, :ng], U[:, :ng])
    # the fused wave-speed maxima give the next step's dt bit-exactly (valid region of the new state)
    sc = scratch.cpu().numpy()
    assert sc[3] == 0
    w = sc[:2].view(np.float64)
    new_valid = np.ascontiguousarray(got[v])
    assert 0.8 * min(dx / w[0], dy / w[1]) == oracle.cfl_dt(new_valid, 0, dx, dy, 1.4, 0.8)


@pytest.mark.parametrize("kind,nx,ny,nsteps", [("sedov", 128, 128, 40), ("shock", 96, 80, 25)])
def test_sweep_multi_step(kind, nx, ny, nsteps):
    """fill_BC -> dt -> evolve loop (pyro_sim.py:241-256) for many steps vs the oracle"""
    import oracle
    from pyro2_b200 import ops
    ng = 4
    bc = ("outflow",) * 4
    U = make_state(nx, ny, ng, kind)
    dx, dy = 1.0 / nx, 1.0 / ny
    prm_o, prm_d = oracle.comp_params(), ops.comp_params()
    A = _to_device(U)
    B = A.clone()
    scratch = ops.new_scratch()
    Uo = U.copy()
    worst = 0.0
    v = (slice(ng, ng + nx), slice(ng, ng + ny))
    for step in range(nsteps):
        Uo = _filled(Uo, ng, bc)
        dt_o = oracle.cfl_dt(Uo, ng, dx, dy, 1.4, 0.8)
        ops.fill_ghost(A, nx, ny, ng, [bc] * 4)
        wx, wy = ops.cfl_wavemax(A, nx, ny, ng, 1.4, scratch)
        dt_d = 0.8 * min(dx / wx, dy / wy)
        assert abs(dt_d - dt_o) <= 1e-12 * dt_o
        dt = dt_o * (0.01 if step == 0 else 1.0)
        Uo = oracle.compressible_step(Uo, ng, dx, dy, dt, prm_o)
        ops.compressible_sweep(A, B, nx, ny, ng, dx, dy, dt, prm_d, scratch)
        A, B = B, A
        got = _to_host(A, ny + 2 * ng)
        worst = max(worst, max(rel_l2(got[v][..., n], Uo[v][..., n]) for n in (0, 1)))
    assert worst < 1e-10, worst


def test_sweep_invalid_state_flag():
    from pyro2_b200 import ops
    ng, nx, ny = 4, 32, 32
    U = _filled(make_state(nx, ny, ng, "smooth"), ng)
    U[ng + 5, ng + 7, 0] = -1.0   # negative density in a valid cell
    Pin = _to_device(U)
    Pout = Pin.clone()
    scratch = ops.new_scratch()
    ops.compressible_sweep(Pin, Pout, nx, ny, ng, 1.0 / nx, 1.0 / ny, 1e-4, ops.comp_params(), scratch)
    assert int(scratch[3]) != 0


def test_sweep_rejects_bad_arguments():
    from pyro2_b200 import ops
    Pin = ops.alloc_planes(4, 16, 16)
    with pytest.raises(ValueError):
        ops.compressible_sweep(Pin, Pin, 8, 8, 4, 1.0, 1.0, 1e-3, ops.comp_params(), ops.new_scratch())
    with pytest.raises(ValueError):
        ops.compressible_sweep(Pin, Pin.clone(), 10, 10, 3, 1.0, 1.0, 1e-3, ops.comp_params(), ops.new_scratch())

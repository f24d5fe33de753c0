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


# ---- the sweep's branch-free fp64 helpers on the device (hydro_core.cuh: rcp / fdiv / fsqrt) -----------------------
def fastmath_cases():
    """operands for the helper probe: ordinary magnitudes (<= 2 ulp of IEEE expected) and the special operands whose
    results the kernels must not rely on (documented classes: NaN for 0 / denormal / inf operands)"""
    rng = np.random.default_rng(7)
    normal = np.concatenate([10.0 ** rng.uniform(-12, 12, 4000), rng.uniform(0.5, 2.0, 2000),
                             10.0 ** rng.uniform(-290, -250, 200), 10.0 ** rng.uniform(250, 290, 200),
                             [1.0, 2.0, 0.5, 1.4, 3.0, 1e-10, 1e-5, 1.0 - 2.0 ** -53, 1.0 + 2.0 ** -52]])
    special = np.array([0.0, -0.0, 5e-324, 1e-310, -1e-310, np.inf, -np.inf, np.nan])
    return normal, special


def ulps(got, ref):
    return np.abs(got - ref) / np.spacing(np.abs(ref))


def check_fastmath(probe):
    """probe(op, a, b) -> out (numpy).  Shared by the device test and the emulator's CPU test."""
    normal, special = fastmath_cases()
    rng = np.random.default_rng(8)
    sign = np.where(rng.random(normal.size) < 0.5, -1.0, 1.0)
    with np.errstate(all="ignore"):
        # ordinary operands: within 2 ulp of the correctly rounded result
        assert ulps(probe(0, sign * normal, None), 1.0 / (sign * normal)).max() <= 2.0
        num = rng.permutation(normal)[: normal.size]
        mid = (np.abs(np.log10(normal)) < 200) & (np.abs(np.log10(num)) < 200)     # keep the quotient representable
        assert ulps(probe(1, (sign * num)[mid], normal[mid]), (sign * num)[mid] / normal[mid]).max() <= 2.0
        assert ulps(probe(2, normal, None), np.sqrt(normal)).max() <= 2.0
        # special operands: no IEEE special-case handling -- 0 and denormals (flushed) give NaN through 0 * inf, inf
        # gives NaN, NaN propagates (never a finite value); negative operands of fsqrt give NaN.  Callers must guard (hllc_lm's chi does).
        assert not np.isfinite(probe(0, special, None)).any()
        assert not np.isfinite(probe(1, np.ones_like(special), special)).any()
        assert not np.isfinite(probe(2, special, None)).any()
        assert np.isnan(probe(2, np.zeros(1), None)).all()          # the one that bit: fsqrt(0) = 0 * inf
        assert np.isnan(probe(2, -normal[:64], None)).all()
        # 0 / x is an exact zero
        assert (probe(1, np.zeros(64), normal[:64]) == 0.0).all()


def hllc_lm_face_cases():
    """(left, right) conserved states (rho, E, mn, mt) of faces in a gas at rest with a pressure jump (Sedov's and
    Sod's initial data), plus moving ones; expected normal-momentum fluxes from a direct restatement of
    riemann.py:864-1019 in numpy"""
    rng = np.random.default_rng(9)
    n = 256
    rho_l, rho_r = rng.uniform(0.5, 2, n), rng.uniform(0.5, 2, n)
    p_l, p_r = 10.0 ** rng.uniform(-5, 2, n), 10.0 ** rng.uniform(-5, 2, n)
    u_l, u_r, t_l, t_r = (rng.uniform(-1, 1, n) for _ in range(4))
    rest = np.arange(n) < n // 2
    for a in (u_l, u_r, t_l, t_r):
        a[rest] = 0.0
    g = 1.4
    L = np.stack([rho_l, p_l / (g - 1) + 0.5 * rho_l * (u_l ** 2 + t_l ** 2), rho_l * u_l, rho_l * t_l], axis=1)
    R = np.stack([rho_r, p_r / (g - 1) + 0.5 * rho_r * (u_r ** 2 + t_r ** 2), rho_r * u_r, rho_r * t_r], axis=1)
    return np.ascontiguousarray(L), np.ascontiguousarray(R), rest


def hllc_lm_mn_flux_numpy(L, R, gamma=1.4):
    """riemann_hllc_lowspeed (riemann.py:864-1019), the normal-momentum flux only, vectorised over faces; pressure jumps
    strong enough for the two-rarefaction / two-shock refinement of estimate_wave_speed are excluded by the caller"""
    out = np.zeros(len(L))
    for k, (l, r) in enumerate(zip(L, R)):
        rho_l, E_l, mn_l, mt_l = l
        rho_r, E_r, mn_r, mt_r = r
        un_l, ut_l, un_r, ut_r = mn_l / rho_l, mt_l / rho_l, mn_r / rho_r, mt_r / rho_r
        p_l = max((E_l - 0.5 * rho_l * (un_l ** 2 + ut_l ** 2)) * (gamma - 1), 1e-10)
        p_r = max((E_r - 0.5 * rho_r * (un_r ** 2 + ut_r ** 2)) * (gamma - 1), 1e-10)
        c_l, c_r = max(1e-10, np.sqrt(gamma * p_l / rho_l)), max(1e-10, np.sqrt(gamma * p_r / rho_r))
        pstar = 0.5 * (p_l + p_r) + 0.5 * (un_l - un_r) * 0.5 * (rho_l + rho_r) * 0.5 * (c_l + c_r)
        assert not (max(p_l, p_r) > 2 * min(p_l, p_r) and (pstar < min(p_l, p_r) or pstar > max(p_l, p_r)))
        S_l = un_l - c_l * (np.sqrt(1 + (gamma + 1) / (2 * gamma) * (pstar / p_l - 1)) if pstar > p_l else 1.0)
        S_r = un_r + c_r * (np.sqrt(1 + (gamma + 1) / (2 / gamma) * (pstar / p_r - 1)) if pstar > p_r else 1.0)
        al, ar = rho_l * (S_l - un_l), rho_r * (S_r - un_r)
        S_c = (p_r - p_l + al * un_l - ar * un_r) / (al - ar)
        chi = min(1.0, np.sqrt(max(un_l ** 2 + ut_l ** 2, un_r ** 2 + ut_r ** 2)) / max(c_l, c_r))
        phi = chi * (2 - chi)
        pstar_lr = 0.5 * (p_l + p_r) + 0.5 * phi * (al * (S_c - un_l) + ar * (S_c - un_r))
        if S_r <= 0:
            out[k] = mn_r * un_r + p_r
        elif S_c <= 0 < S_r:
            out[k] = (S_c * (S_r * mn_r - (mn_r * un_r + p_r)) + S_r * pstar_lr) / (S_r - S_c)
        elif S_l < 0 < S_c:
            out[k] = (S_c * (S_l * mn_l - (mn_l * un_l + p_l)) + S_l * pstar_lr) / (S_l - S_c)
        else:
            out[k] = mn_l * un_l + p_l
    return out


def check_hllc_lm_at_rest(probe):
    L, R, rest = hllc_lm_face_cases()
    keep = []
    for k in range(len(L)):
        try:
            hllc_lm_mn_flux_numpy(L[k:k + 1], R[k:k + 1])
            keep.append(k)
        except AssertionError:
            pass
    keep = np.array(keep)
    assert rest[keep].sum() > 20 and (~rest[keep]).sum() > 20
    L, R = L[keep], R[keep]
    got = probe(3, L, R)
    ref = hllc_lm_mn_flux_numpy(L, R)
    assert np.isfinite(got).all()
    assert np.abs(got - ref).max() <= 1e-13 * np.abs(ref).max()


def _device_probe(op, a, b):
    import ctypes as C
    import torch
    from pyro2_b200 import _lib
    ta = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()
    tb = None if b is None else torch.from_numpy(np.ascontiguousarray(b, dtype=np.float64)).cuda()
    n = len(a)
    out = torch.empty(n, dtype=torch.float64, device="cuda")
    _lib.check(_lib.lib().p2b_test_fastmath(op, ta.data_ptr(), None if tb is None else tb.data_ptr(), out.data_ptr(),
                                            n, _lib.stream_ptr()))
    return out.cpu().numpy()


def test_fastmath_helpers_on_device():
    check_fastmath(_device_probe)


def test_hllc_lm_gas_at_rest_on_device():
    """the chi blend at Mach 0 (riemann.py:989-998): fsqrt(0) is NaN on the device, the solver must not see it"""
    check_hllc_lm_at_rest(_device_probe)


@pytest.mark.parametrize("solver", ["CGF", "HLLC_lm"])
def test_sweep_other_riemann_solvers_gas_at_rest(solver):
    """one step from Sedov's initial data (u = v = 0 exactly, pressure jump) with the other Riemann solvers"""
    import oracle
    from pyro2_b200 import ops
    ng, nx, ny = 4, 96, 80
    U = _filled(make_state(nx, ny, ng, "sedov"), ng)
    dx, dy = 1.0 / nx, 1.0 / ny
    dt = 0.5 * oracle.cfl_dt(U, ng, dx, dy, 1.4, 0.8)
    ref = oracle.compressible_step(U, ng, dx, dy, dt, oracle.comp_params(riemann=solver))
    Pin = _to_device(U)
    Pout = Pin.clone()
    ops.compressible_sweep(Pin, Pout, nx, ny, ng, dx, dy, dt, ops.comp_params(riemann=solver), ops.new_scratch())
    got = _to_host(Pout, ny + 2 * ng)
    v = (slice(ng, ng + nx), slice(ng, ng + ny))
    for n in range(4):
        scale = np.linalg.norm(ref[v][..., n].ravel()) if n < 2 else np.linalg.norm(ref[v][..., 1].ravel())
        assert np.linalg.norm((got[v][..., n] - ref[v][..., n]).ravel()) < 1e-12 * scale


def test_shared_divisor_quotients_are_ieee_on_device():
    """cons_to_prim / cfl_speeds divide by the density through one refined reciprocal (hydro_core.cuh: shared_div / div_by);
    the quotients must be the IEEE ones bit for bit -- zero numerators of either sign (gas at rest, reflected ghosts) and
    negative or extreme divisors (library fallback) included"""
    rng = np.random.default_rng(11)
    n = 20000
    a = np.concatenate([rng.standard_normal(n) * 10.0 ** rng.uniform(-8, 8, n), np.zeros(64), -np.zeros(64),
                        10.0 ** rng.uniform(-280, 280, 256), [1.0, -1.0, 3.0, 1e-300, 1e300]])
    b = np.concatenate([10.0 ** rng.uniform(-6, 6, n) * np.where(rng.random(n) < 0.1, -1.0, 1.0),
                        10.0 ** rng.uniform(-3, 3, 64), -(10.0 ** rng.uniform(-3, 3, 64)),
                        10.0 ** rng.uniform(-3, 3, 256), [3.0, 3.0, 1e-250, 1e-300, 1e250]])
    with np.errstate(all="ignore"):
        want = a / b
    sane = (np.abs(a) == 0) | ((np.abs(a) > 2.0 ** -900) & (np.abs(a) < 2.0 ** 900))
    got = _device_probe(4, a, b)
    lib = _device_probe(5, a, b)
    assert np.array_equal(lib.view(np.int64), want.view(np.int64))                       # __ddiv_rn is IEEE (sanity)
    assert np.array_equal(got[sane].view(np.int64), want[sane].view(np.int64))           # bits, signed zeros included
    assert (np.signbit(got[n:n + 128]) == np.signbit(want[n:n + 128])).all()

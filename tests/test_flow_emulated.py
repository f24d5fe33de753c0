"""CPU: the Burgers / incompressible stage kernels and their host entry points (pyro2_b200/csrc/flow.cu,
unchanged) compiled for the host through tests/emu/cuda_emu.h, driven through the p2b_flow_* C ABI over
numpy memory and compared bit-for-bit with the oracle and the reference-generated fixtures.  The
emulator is test infrastructure: the product only loads the nvcc-built library."""
import numpy as np
import pytest

import oracle
from emu_util import EmuFlow, load_flow_emu, load_mg_emu
from golden_util import load_flow


@pytest.fixture(scope="module")
def flow():
    return load_flow_emu()


def _periodic_fields(nx, ny, ng, seed, count):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(count):
        a = np.zeros((nx + 2 * ng, ny + 2 * ng))
        a[ng:-ng, ng:-ng] = rng.standard_normal((nx, ny))
        oracle.fill_ghost(a, ng, ("periodic",) * 4)
        out.append(a)
    return out


@pytest.mark.parametrize("limiter", [0, 1, 2])
@pytest.mark.parametrize("nx,ny", [(16, 16), (24, 40)])
def test_emulated_mac_vels_and_states_match_oracle(flow, limiter, nx, ny):
    ng, dx, dy, dt = 4, 1.0 / nx, 0.7 / ny, 0.013
    u, v, gx, gy = _periodic_fields(nx, ny, ng, 10 * limiter + nx, 4)
    # include exact zeros and sign changes: the Riemann / upwind selections have == 0 branches
    u[ng + 2:ng + 5, :] = 0.0
    v[:, ng + 1:ng + 3] = 0.0
    f = EmuFlow(flow, nx, ny, ng, dx, dy)
    f.interface_states(u, v, gx, gy, dt, limiter)
    f.mac_vels()
    um, vm = oracle.incomp_mac_vels(u, v, gx, gy, ng, dx, dy, dt, limiter)
    assert np.array_equal(f.plane("u_MAC"), um) and np.array_equal(f.plane("v_MAC"), vm)
    f.ck(flow.p2b_flow_upwind_states(f.h, None))
    ref = oracle.incomp_states(u, v, gx, gy, ng, dx, dy, dt, limiter, um, vm)
    for name, r in zip(("u_xint", "v_xint", "u_yint", "v_yint"), ref):
        assert np.array_equal(f.plane(name), r), name
    m = f.maxabs(u, v)
    assert m[0] == np.abs(u).max() and m[1] == np.abs(v).max()
    f.close()


@pytest.mark.parametrize("fname", ["incomp_shear32.npz", "incomp_converge32.npz"])
def test_emulated_incompressible_run_matches_reference(flow, fname):
    """all six state planes after the recorded steps of the reference's Pyro("incompressible") run:
    emulated stage kernels + emulated multigrid projections"""
    z, rp, _ = load_flow(fname)
    ng, n = int(z["ng"]), rp["mesh.nx"]
    bc = (rp["mesh.xlboundary"], rp["mesh.xrboundary"], rp["mesh.ylboundary"], rp["mesh.yrboundary"])
    P = np.ascontiguousarray(z["P0"])
    f = EmuFlow(flow, n, n, ng, 1.0 / n, 1.0 / n)
    mg = load_mg_emu()
    fill = lambda a: oracle.fill_ghost(a, ng, bc)
    for dt in z["dts"][:6]:
        for k in range(6):
            fill(P[k])
        f.incomp_evolve(mg, P, float(dt), rp["incompressible.limiter"], rp["incompressible.proj_type"], bc, fill)
    Q = np.ascontiguousarray(z["P0"])
    for dt in z["dts"][:6]:
        for k in range(6):
            fill(Q[k])
        oracle.incomp_evolve(Q, ng, float(dt), limiter=rp["incompressible.limiter"], proj_type=rp["incompressible.proj_type"])
    assert np.array_equal(P, Q)
    if len(z["dts"]) <= 6:
        assert np.array_equal(P, z["P"])
    f.close()


def test_emulated_incompressible_proj_type_1(flow):
    n, ng = 32, 4
    z, rp, _ = load_flow("incomp_shear32.npz")
    bc = ("periodic",) * 4
    P = np.ascontiguousarray(z["P0"])
    Q = P.copy()
    f = EmuFlow(flow, n, n, ng, 1.0 / n, 1.0 / n)
    fill = lambda a: oracle.fill_ghost(a, ng, bc)
    for dt in (1e-3, 2e-3):
        for k in range(6):
            fill(P[k]); fill(Q[k])
        f.incomp_evolve(load_mg_emu(), P, dt, 1, 1, bc, fill)
        oracle.incomp_evolve(Q, ng, dt, limiter=1, proj_type=1)
    assert np.array_equal(P, Q)
    f.close()


@pytest.mark.parametrize("fname", ["burgers_test.npz", "burgers_converge32.npz", "burgers_tophat32.npz"])
def test_emulated_burgers_run_matches_reference(flow, fname):
    z, rp, _ = load_flow(fname)
    ng, n = int(z["ng"]), rp["mesh.nx"]
    bc = (rp["mesh.xlboundary"], rp["mesh.xrboundary"], rp["mesh.ylboundary"], rp["mesh.yrboundary"])
    u, v = z["P0"][0].copy(), z["P0"][1].copy()
    f = EmuFlow(flow, n, n, ng, 1.0 / n, 1.0 / n)
    for dt in z["dts"]:
        oracle.fill_ghost(u, ng, bc)
        oracle.fill_ghost(v, ng, bc)
        f.burgers_evolve(u, v, float(dt), rp["advection.limiter"])
    assert np.array_equal(u, z["P"][0]) and np.array_equal(v, z["P"][1])
    f.close()


@pytest.mark.parametrize("nx,ny", [(16, 24), (40, 33)])
def test_emulated_hse_boundary_matches_oracle(nx, ny):
    """the compressible "hse" user boundary (compressible/BC.py): kernel vs the oracle, variable by variable"""
    import ctypes as C
    from emu_util import load_bc_emu
    from pyro2_b200 import _lib
    lib = load_bc_emu()
    ng, gamma, grav, dy = 4, 1.4, -1.7, 0.031
    rng = np.random.default_rng(nx)
    P = 1.0 + rng.random((4, nx + 2 * ng, ny + 2 * ng))
    P[1] += 3.0
    Q = P.copy()
    g = _lib.Grid(nx, ny, ng, ny + 2 * ng, (nx + 2 * ng) * (ny + 2 * ng), 1.0, dy)
    for var in (0, 1, 2, 3, 1):
        for side, name in ((0, "ylb"), (1, "yrb")):
            assert lib.p2b_fill_hse_f64(P.ctypes.data, C.byref(g), grav, gamma, var, side, None) == 0
            oracle.fill_hse(Q, ng, dy, grav, gamma, var, name)
            assert np.array_equal(P, Q), (var, name)


@pytest.mark.parametrize("u,v,limiter,nx,ny", [(1.0, 1.0, 2, 32, 32), (-0.7, 0.4, 1, 24, 40), (0.5, -1.2, 0, 16, 16),
                                                (0.0, 1.0, 2, 16, 24)])
def test_emulated_advection_matches_oracle(flow, u, v, limiter, nx, ny):
    import ctypes as C
    ng, dx, dy = 4, 1.0 / nx, 1.0 / ny
    dt = 0.8 * min(dx / max(abs(u), 1e-12), dy / max(abs(v), 1e-12))
    (a,) = _periodic_fields(nx, ny, ng, nx + limiter, 1)
    ref = a.copy()
    f = EmuFlow(flow, nx, ny, ng, dx, dy)
    for _ in range(3):
        oracle.fill_ghost(a, ng, ("periodic",) * 4)
        oracle.fill_ghost(ref, ng, ("periodic",) * 4)
        f.ck(flow.p2b_flow_advection_update(f.h, a.ctypes.data, u, v, dt, limiter, None))
        ref = oracle.advection_evolve(ref, ng, dx, dy, dt, u, v, limiter)
        assert np.array_equal(a, ref)
    f.close()


def test_emulated_ambient_boundary():
    import ctypes as C
    from emu_util import load_bc_emu
    from pyro2_b200 import _lib
    lib = load_bc_emu()
    nx, ny, ng = 12, 20, 4
    rng = np.random.default_rng(0)
    P = rng.random((4, nx + 2 * ng, ny + 2 * ng))
    Q = P.copy()
    g = _lib.Grid(nx, ny, ng, ny + 2 * ng, (nx + 2 * ng) * (ny + 2 * ng), 1.0, 1.0)
    assert lib.p2b_fill_ambient_f64(P.ctypes.data, C.byref(g), 2, 1, 0.375, None) == 0
    Q[2][:, ng + ny:] = 0.375
    assert np.array_equal(P, Q)

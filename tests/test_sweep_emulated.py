"""CPU: the fused sweep kernel's source (pyro2_b200/csrc/sweep_task.cuh) compiled for a host-side
warp emulator (tests/emu/sweep_emu.cpp: 32 threads in lock step stand in for the lanes, memcpy for
the TMA bulk copies) and compared with the oracle.  This checks the kernel's strip / segment
indexing, halo handling and row pipeline on the GPU-less build box.  The emulator is test
infrastructure; nothing in pyro2_b200/ can load it."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle
from conftest import make_state, rel_l2

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "emu")


@pytest.fixture(scope="module")
def emu():
    from emu_util import load_sweep_emu
    return load_sweep_emu()


def _emu_step(lib, U, ng, dx, dy, dt, prm, seglen, flips=(0, 0), heat=None, src_copy_yhi=0, geometry=None, xflips=(0, 0)):
    """heat: the heating profile with its ghost cells filled like a scalar (what the product passes); geometry: the
    (geo_i, geo_j) tables of a SphericalPolar grid (rows of length qx / pitch)"""
    P = oracle.to_planes(U)
    _, qx, qy = P.shape
    pitch = (qy + 15) // 16 * 16
    Pin = np.zeros((4, qx, pitch))
    Pin[:, :, :qy] = P
    Pout = Pin.copy()
    heat_p = None
    if heat is not None:
        heat_p = np.zeros((qx, pitch))
        heat_p[:, :qy] = heat
    scratch = np.zeros(8, dtype=np.uint64)
    lib.emu_compressible_sweep(Pin.ctypes.data, Pout.ctypes.data, qx - 2 * ng, qy - 2 * ng, ng, pitch, qx * pitch,
                               dx, dy, dt, prm.gamma, prm.z0, prm.z1, prm.delta, prm.cvisc, prm.limiter,
                               prm.use_flattening, prm.no_avisc_xhi, prm.no_avisc_yhi, seglen, scratch.ctypes.data, None,
                               prm.grav, flips[0], flips[1], prm.riemann, prm.xl_solid, prm.yl_solid,
                               None if heat is None else heat_p.ctypes.data, prm.heat_rate, prm.do_sponge,
                               prm.sponge_rho_begin, prm.sponge_rho_full, prm.sponge_timescale, src_copy_yhi,
                               None if geometry is None else geometry[0].ctypes.data,
                               None if geometry is None else geometry[1].ctypes.data,
                               0 if geometry is None else geometry[0].shape[1], 0 if geometry is None else geometry[1].shape[1],
                               xflips[0], xflips[1])
    return oracle.from_planes(np.ascontiguousarray(Pout[:, :, :qy])), scratch


@pytest.mark.parametrize("kind,nx,ny,limiter,flat,seglen", [
    ("smooth", 24, 20, 2, 1, 8), ("shock", 33, 37, 2, 1, 11), ("shock", 16, 64, 1, 1, 16),
    ("shock", 20, 12, 0, 0, 32), ("sedov", 32, 32, 2, 1, 5)])
def test_emulated_sweep_matches_oracle(emu, kind, nx, ny, limiter, flat, seglen):
    ng = 4
    U = make_state(nx, ny, ng, kind)
    for n in range(4):
        pl = np.ascontiguousarray(U[:, :, n])
        oracle.fill_ghost(pl, ng, ("outflow",) * 4)
        U[:, :, n] = pl
    dx, dy = 1.0 / nx, 1.0 / ny
    dt = 0.5 * oracle.cfl_dt(U, ng, dx, dy, 1.4, 0.8)
    prm = oracle.comp_params(limiter=limiter, use_flattening=flat)
    got, scratch = _emu_step(emu, U, ng, dx, dy, dt, prm, seglen)
    ref = oracle.compressible_step(U, ng, dx, dy, dt, prm)
    v = (slice(ng, ng + nx), slice(ng, ng + ny))
    assert not np.isnan(got[v]).any()
    for n in range(4):
        assert rel_l2(got[v][..., n], ref[v][..., n]) < 1e-13
    assert scratch[3] == 0
    w = scratch[:2].view(np.float64)
    assert 0.8 * min(dx / w[0], dy / w[1]) == oracle.cfl_dt(np.ascontiguousarray(got[v]), 0, dx, dy, 1.4, 0.8)


@pytest.mark.parametrize("bc,nx,ny,seglen", [
    (("periodic", "periodic", "hse", "hse"), 24, 40, 8),
    (("outflow", "outflow", "hse", "hse"), 33, 37, 11),
    (("reflect", "outflow", "reflect", "reflect"), 16, 48, 16),     # ghost-cell sources change sign
    (("periodic", "periodic", "reflect", "outflow"), 20, 36, 32)])
def test_emulated_sweep_with_gravity_matches_oracle(emu, bc, nx, ny, seglen):
    """the GRAV instantiation of the sweep (source terms on the interface states + predictor-corrector
    source update) on a stratified atmosphere with a perturbation, ghost cells filled like the driver does"""
    from golden_util import var_bcs
    ng, gamma, grav = 4, 1.4, -1.5
    dx, dy = 1.0 / nx, 2.0 / ny
    rng = np.random.default_rng(nx)
    y = (np.arange(ny + 2 * ng) + 0.5 - ng) * dy
    dens = np.broadcast_to(2.0 * np.exp(-y / 1.3)[None, :], (nx + 2 * ng, ny + 2 * ng)).copy()
    dens *= 1.0 + 0.05 * rng.standard_normal(dens.shape)
    pres = 1.3 * abs(grav) * dens * (1.0 + 0.02 * rng.standard_normal(dens.shape))
    u, v = 0.1 * rng.standard_normal(dens.shape), 0.1 * rng.standard_normal(dens.shape)
    P = np.stack([dens, pres / (gamma - 1.0) + 0.5 * dens * (u * u + v * v), dens * u, dens * v])
    rp = dict(zip(("mesh.xlboundary", "mesh.xrboundary", "mesh.ylboundary", "mesh.yrboundary"), bc))
    bcs = var_bcs(rp)
    for k in range(4):
        oracle.fill_ghost(P[k], ng, bcs[k])
        for side in ("ylb", "yrb"):
            if bcs[k][2 + (side == "yrb")] == "hse":
                oracle.fill_hse(P, ng, dy, grav, gamma, k, side)
    U = oracle.from_planes(P)
    dt = 0.5 * oracle.cfl_dt(U, ng, dx, dy, gamma, 0.8)
    prm = oracle.comp_params(grav=grav, src_bcs=bcs)
    flips = (int(bc[2] == "reflect"), int(bc[3] == "reflect"))
    got, scratch = _emu_step(emu, U, ng, dx, dy, dt, prm, seglen, flips)
    ref = oracle.compressible_step(U, ng, dx, dy, dt, prm)
    v_ = (slice(ng, ng + nx), slice(ng, ng + ny))
    assert not np.isnan(got[v_]).any() and scratch[3] == 0
    for n in range(4):
        assert rel_l2(got[v_][..., n], ref[v_][..., n]) < 1e-13
    # and gravity really did something
    ref0 = oracle.compressible_step(U, ng, dx, dy, dt, oracle.comp_params())
    assert rel_l2(ref[v_][..., 3], ref0[v_][..., 3]) > 1e-4


@pytest.mark.parametrize("kind,bc,nx,ny,grav,seglen", [
    ("shock", ("outflow",) * 4, 33, 37, 0.0, 11), ("sedov", ("outflow",) * 4, 32, 32, 0.0, 5),
    ("shock", ("reflect", "reflect", "reflect", "outflow"), 24, 40, 0.0, 8),        # solid -x / -y walls
    ("smooth", ("reflect", "outflow", "reflect", "reflect"), 16, 48, -1.2, 16)])    # walls + gravity
@pytest.mark.parametrize("solver", ["CGF", "HLLC_lm"])
def test_emulated_sweep_with_cgf_matches_oracle(emu, kind, bc, nx, ny, grav, seglen, solver):
    """the RIEMANN = 1 / 2 instantiations: riemann_cgf + consFlux, or the low-Mach HLLC, at all four Riemann problems
    of a cell"""
    from golden_util import var_bcs
    ng = 4
    U = make_state(nx, ny, ng, kind)
    rp = dict(zip(("mesh.xlboundary", "mesh.xrboundary", "mesh.ylboundary", "mesh.yrboundary"), bc))
    bcs = var_bcs(rp)
    for n in range(4):
        pl = np.ascontiguousarray(U[:, :, n])
        oracle.fill_ghost(pl, ng, bcs[n])
        U[:, :, n] = pl
    dx, dy = 1.0 / nx, 1.0 / ny
    dt = 0.5 * oracle.cfl_dt(U, ng, dx, dy, 1.4, 0.8)
    prm = oracle.comp_params(riemann=solver, xl_solid=int(bc[0] == "reflect"), yl_solid=int(bc[2] == "reflect"),
                             grav=grav, src_bcs=bcs)
    flips = (int(bc[2] == "reflect"), int(bc[3] == "reflect"))
    got, scratch = _emu_step(emu, U, ng, dx, dy, dt, prm, seglen, flips)
    ref = oracle.compressible_step(U, ng, dx, dy, dt, prm)
    v = (slice(ng, ng + nx), slice(ng, ng + ny))
    assert not np.isnan(got[v]).any() and scratch[3] == 0
    for n in range(4):
        assert rel_l2(got[v][..., n], ref[v][..., n]) < 1e-13
    hllc = oracle.compressible_step(U, ng, dx, dy, dt, oracle.comp_params(grav=grav, src_bcs=bcs))
    assert rel_l2(ref[v][..., 0], hllc[v][..., 0]) > (1e-6 if solver == "CGF" else 1e-9)        # and it is a different solver


@pytest.mark.parametrize("bc,nx,ny,grav,sponge,seglen", [
    (("outflow",) * 4, 24, 24, 0.0, None, 8),                                   # the heating problem: no gravity
    (("outflow", "outflow", "hse", "hse"), 20, 40, -2.0, None, 11),             # plume
    (("periodic", "periodic", "reflect", "outflow"), 16, 48, -2.0, (0.6, 0.2, 1.e-2), 16),   # convection-like
    (("periodic", "periodic", "reflect", "ambient"), 16, 48, -2.0, (0.6, 0.2, 1.e-2), 16)])  # convection: ambient top
def test_emulated_sweep_with_heating_and_sponge_matches_oracle(emu, bc, nx, ny, grav, sponge, seglen):
    from golden_util import var_bcs
    ng, gamma = 4, 1.4
    dx, dy = 1.0 / nx, 2.0 / ny
    rng = np.random.default_rng(ny)
    qx, qy = nx + 2 * ng, ny + 2 * ng
    y = (np.arange(qy) + 0.5 - ng) * dy
    x = (np.arange(qx) + 0.5 - ng) * dx
    dens = np.broadcast_to(1.5 * np.exp(-y / 0.9)[None, :], (qx, qy)) * (1.0 + 0.05 * rng.standard_normal((qx, qy)))
    pres = 1.8 * dens * (1.0 + 0.02 * rng.standard_normal((qx, qy)))
    u, v = 0.1 * rng.standard_normal((qx, qy)), 0.1 * rng.standard_normal((qx, qy))
    P = np.stack([dens, pres / (gamma - 1.0) + 0.5 * dens * (u * u + v * v), dens * u, dens * v])
    prof = np.exp(-(np.sqrt((x[:, None] - 0.5) ** 2 + (y[None, :] - 0.7) ** 2) / 0.3) ** 2)
    rp = dict(zip(("mesh.xlboundary", "mesh.xrboundary", "mesh.ylboundary", "mesh.yrboundary"), bc))
    bcs = var_bcs(rp)
    for k in range(4):
        oracle.fill_ghost(P[k], ng, bcs[k])
        for side in ("ylb", "yrb"):
            if bcs[k][2 + (side == "yrb")] == "hse":
                oracle.fill_hse(P, ng, dy, grav, gamma, k, side)
            if bcs[k][2 + (side == "yrb")] == "ambient":          # constant state beyond the top boundary
                P[k][:, ng + ny:] = (0.05, 0.4, 0.0, 0.0)[k]
    U = oracle.from_planes(P)
    dt = 0.5 * oracle.cfl_dt(U, ng, dx, dy, gamma, 0.8)
    prm = oracle.comp_params(grav=grav, src_bcs=bcs, heat_rate=0.7, heat_profile=prof, sponge=sponge)
    # what the product hands the kernel: the profile ghost-filled like the (even) energy-source array
    heat = prof.copy()
    oracle.fill_ghost(heat, ng, tuple("outflow" if b in ("hse", "ambient") else b for b in bcs[1]))
    flips = (int(bc[2] == "reflect"), int(bc[3] == "reflect"))
    got, scratch = _emu_step(emu, U, ng, dx, dy, dt, prm, seglen, flips, heat=heat, src_copy_yhi=int(bc[3] == "ambient"))
    ref = oracle.compressible_step(U, ng, dx, dy, dt, prm)
    v_ = (slice(ng, ng + nx), slice(ng, ng + ny))
    assert not np.isnan(got[v_]).any() and scratch[3] == 0
    for n in range(4):
        assert rel_l2(got[v_][..., n], ref[v_][..., n]) < 1e-13
    plain = oracle.compressible_step(U, ng, dx, dy, dt, oracle.comp_params(grav=grav, src_bcs=bcs))
    assert rel_l2(ref[v_][..., 1], plain[v_][..., 1]) > 1e-6


@pytest.mark.parametrize("solver", ["HLLC", "CGF", "HLLC_lm"])
def test_emulated_sweep_with_unphysical_interface_states_matches_oracle(emu, solver):
    """limiter 0 at a strong jump (a stratification wrapped around by periodic y boundaries): interface states with
    negative density reach the Riemann solvers.  The reference carries on with c = max(smallc, sqrt(negative)) =
    smallc (Python max semantics under numba; pinned in test_oracle_vs_reference.py); dmax / dmin in hydro_core.cuh
    have the same semantics, so the kernel must give the oracle's finite numbers rather than NaN.  (Found by
    scripts/fuzz_sweep_emulated.py.)"""
    from golden_util import var_bcs
    ng, nx, ny, gamma = 4, 6, 31, 1.4
    rng = np.random.default_rng(135)
    qx, qy = nx + 2 * ng, ny + 2 * ng
    y = (np.arange(qy) + 0.5 - ng) / ny
    dens = np.broadcast_to(1.5 * np.exp(-y / 0.4)[None, :], (qx, qy)) * (1.0 + 0.1 * rng.standard_normal((qx, qy)))
    pres = 1.8 * dens * (1.0 + 0.05 * rng.standard_normal((qx, qy)))
    u, v = 0.3 * rng.standard_normal((qx, qy)), 0.3 * rng.standard_normal((qx, qy))
    P = np.stack([dens, pres / (gamma - 1.0) + 0.5 * dens * (u * u + v * v), dens * u, dens * v])
    bcs = var_bcs({"mesh.xlboundary": "outflow", "mesh.xrboundary": "outflow", "mesh.ylboundary": "periodic",
                   "mesh.yrboundary": "periodic"})
    for k in range(4):
        oracle.fill_ghost(P[k], ng, bcs[k])
    U = oracle.from_planes(P)
    dx, dy = 1.0 / nx, 1.0 / ny
    dt = 0.4 * oracle.cfl_dt(U, ng, dx, dy, gamma, 0.8)
    prm = oracle.comp_params(limiter=0, use_flattening=0, cvisc=0.0, riemann=solver)
    _, st = oracle.compressible_step(U, ng, dx, dy, dt, prm, stages=True)
    assert min(st[k][..., 0].min() for k in ("Uxl_hat", "Uxr_hat", "Uyl_hat", "Uyr_hat")) < 0.0     # the regime in question
    ref = oracle.compressible_step(U, ng, dx, dy, dt, prm)
    got, scratch = _emu_step(emu, U, ng, dx, dy, dt, prm, 8)
    v_ = (slice(ng, ng + nx), slice(ng, ng + ny))
    assert np.isfinite(ref[v_]).all() and np.isfinite(got[v_]).all()
    for n in range(4):
        assert rel_l2(got[v_][..., n], ref[v_][..., n]) < 1e-12


def test_emulated_sweep_spherical_polar_with_heating_and_sponge(emu):
    """problem heating and the sponge on a SphericalPolar grid (the reference applies both in any geometry)"""
    from golden_util import var_bcs
    from pyro2_b200.mesh import patch
    ng, gamma, nx, ny = 4, 1.4, 16, 33
    xmin, xmax, ymin, ymax = 0.6, 1.4, 0.6, 2.4
    g = patch.SphericalPolar(nx, ny, ng=ng, xmin=xmin, xmax=xmax, ymin=ymin, ymax=ymax, device="cpu")
    geom = oracle.spherical_geometry(nx, ny, ng, xmin, xmax, ymin, ymax)
    rng = np.random.default_rng(3)
    qx, qy = nx + 2 * ng, ny + 2 * ng
    r = geom["x2d"]
    dens = np.exp(-(r - xmin) / 0.3) * (1.0 + 0.1 * rng.standard_normal((qx, qy)))      # falls through the sponge range
    pres = 1.5 * dens * (1.0 + 0.05 * rng.standard_normal((qx, qy)))
    u, v = 0.2 * rng.standard_normal((qx, qy)), 0.2 * rng.standard_normal((qx, qy))
    P = np.stack([dens, pres / (gamma - 1.0) + 0.5 * dens * (u * u + v * v), dens * u, dens * v])
    bc = ("reflect", "outflow", "outflow", "outflow")
    bcs = var_bcs(dict(zip(("mesh.xlboundary", "mesh.xrboundary", "mesh.ylboundary", "mesh.yrboundary"), bc)))
    for k in range(4):
        oracle.fill_ghost(P[k], ng, bcs[k])
    U = oracle.from_planes(P)
    dt = 0.5 * oracle.cfl_dt_spherical(U, gamma, 0.8, geom)
    prof = np.exp(-((r - 1.0) ** 2 + (geom["y2d"] - 1.5) ** 2) / 0.05)
    prm = oracle.comp_params(riemann="CGF", grav=-0.8, src_bcs=bcs, geom=geom, xl_solid=1, heat_rate=0.7, heat_profile=prof,
                             sponge=(0.5, 0.15, 1.e-2))
    ref = oracle.compressible_step(U, ng, g.dx, g.dy, dt, prm)
    heat = prof.copy()
    oracle.fill_ghost(heat, ng, bcs[1])
    tables = patch.spherical_sweep_tables(g, (qy + 15) // 16 * 16, bc[0], bc[1])
    got, scratch = _emu_step(emu, U, ng, g.dx, g.dy, dt, prm, 8, heat=heat, geometry=tables, xflips=(1, 0))
    v_ = (slice(ng, ng + nx), slice(ng, ng + ny))
    assert np.isfinite(got[v_]).all() and scratch[3] == 0
    assert np.abs(got[v_] - ref[v_]).max() < 2e-14 * np.abs(ref[v_]).max()
    plain = oracle.compressible_step(U, ng, g.dx, g.dy, dt, oracle.comp_params(riemann="CGF", grav=-0.8, src_bcs=bcs, geom=geom, xl_solid=1))
    assert rel_l2(ref[v_][..., 1], plain[v_][..., 1]) > 1e-6 and rel_l2(ref[v_][..., 2], plain[v_][..., 2]) > 1e-6


@pytest.mark.parametrize("xbc,ybc,nx,ny,grav,limiter,seglen", [
    (("reflect-odd", "outflow"), ("outflow", "outflow"), 20, 37, 0.0, 2, 8),        # the reference's sedov.spherical setup
    (("outflow", "outflow"), ("outflow", "outflow"), 24, 31, -0.7, 0, 11),          # advect.spherical + radial gravity
    (("reflect", "outflow"), ("reflect", "reflect"), 16, 45, -1.2, 1, 16),          # solid inner wall, reflecting cone walls
    (("periodic", "periodic"), ("outflow", "reflect"), 12, 30, 0.5, 2, 32)])
def test_emulated_sweep_spherical_polar_matches_oracle(emu, xbc, ybc, nx, ny, grav, limiter, seglen):
    """the SPH instantiation (SphericalPolar geometry: geometric sources in the tracing, radial gravity and centrifugal
    terms, CGF interface pressures, area / volume weighted corrections and update, spherical viscosity) against the
    oracle, which is pinned to the live reference (test_oracle_vs_reference.py)"""
    from golden_util import var_bcs
    from pyro2_b200.mesh import patch
    ng, gamma = 4, 1.4
    xmin, xmax, ymin, ymax = 0.6, 1.4, 0.6, 2.4
    g = patch.SphericalPolar(nx, ny, ng=ng, xmin=xmin, xmax=xmax, ymin=ymin, ymax=ymax, device="cpu")
    geom = oracle.spherical_geometry(nx, ny, ng, xmin, xmax, ymin, ymax)
    for k in ("Ly", "Ax", "Ay", "V", "dlogAx", "dlogAy"):
        assert np.array_equal(getattr(g, k).numpy(), geom[k]), k            # the product's grid class, bit for bit
    rng = np.random.default_rng(nx * ny)
    qx, qy = nx + 2 * ng, ny + 2 * ng
    r = geom["x2d"]
    dens = (1.0 + 0.5 / r) * (1.0 + 0.1 * rng.standard_normal((qx, qy)))
    pres = 1.5 * dens * (1.0 + 0.05 * rng.standard_normal((qx, qy)))
    u, v = 0.3 * rng.standard_normal((qx, qy)), 0.3 * rng.standard_normal((qx, qy))
    if xbc[0] == "reflect-odd":
        # the literal type reflects EVERY variable oddly (the reference's inputs.sedov.spherical does that): the ghost
        # rows then hold negative densities, and a CGF star state on that side has c = smallc = 1e-10 -- finite in
        # the reference, but ill-conditioned (differences of 1e20-sized terms), so only a face whose upwind side is the
        # physical one is comparable.
        # A purely radial inflow (no theta dependence, v = 0: the y-faces between ghost cells see equal states, the wall
        # face is upwinded from the interior) keeps every face comparable -- the situation behind a Sedov blast.
        prof = rng.standard_normal(qx)[:, None]
        dens = np.broadcast_to((1.0 + 0.5 / r[:, :1]) * (1.0 + 0.1 * prof), (qx, qy)).copy()
        pres = 1.5 * dens
        u, v = np.broadcast_to(-0.2 - 0.1 * np.abs(prof), (qx, qy)).copy(), np.zeros((qx, qy))
    P = np.stack([dens, pres / (gamma - 1.0) + 0.5 * dens * (u * u + v * v), dens * u, dens * v])
    bc = xbc + ybc
    # "reflect" picks the parity per variable, the literal reflect-odd / reflect-even apply to every variable alike
    bcs = var_bcs(dict(zip(("mesh.xlboundary", "mesh.xrboundary", "mesh.ylboundary", "mesh.yrboundary"), bc)))
    for k in range(4):
        oracle.fill_ghost(P[k], ng, bcs[k])
    U = oracle.from_planes(P)
    dt = 0.5 * oracle.cfl_dt_spherical(U, gamma, 0.8, geom)
    prm = oracle.comp_params(limiter=limiter, riemann="CGF", grav=grav, src_bcs=bcs, geom=geom,
                             xl_solid=int(bc[0] == "reflect"), yl_solid=int(bc[2] == "reflect"))
    ref = oracle.compressible_step(U, ng, g.dx, g.dy, dt, prm)
    pitch = (qy + 15) // 16 * 16
    tables = patch.spherical_sweep_tables(g, pitch, bc[0], bc[1])
    got, scratch = _emu_step(emu, U, ng, g.dx, g.dy, dt, prm, seglen, geometry=tables,
                             xflips=(int(bc[0] == "reflect"), int(bc[1] == "reflect")))
    v_ = (slice(ng, ng + nx), slice(ng, ng + ny))
    assert np.isfinite(got[v_]).all() and scratch[3] == 0
    # (the theta momentum of the radial-inflow case is itself only 1e-5: judge every variable on the state's scale)
    assert np.abs(got[v_] - ref[v_]).max() < 2e-14 * np.abs(ref[v_]).max()
    for n in range(4):
        assert rel_l2(got[v_][..., n], ref[v_][..., n]) < 1e-10, n
    cart = oracle.compressible_step(U, ng, g.dx, g.dy, dt, oracle.comp_params(limiter=limiter, riemann="CGF"))
    assert rel_l2(ref[v_][..., 0], cart[v_][..., 0]) > 1e-4                 # the geometry matters


# ---- the device's branch-free fp64 helpers as the emulator restates them ------------------------------------------------
def _emu_probe(lib):
    def probe(op, a, b):
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = None if b is None else np.ascontiguousarray(b, dtype=np.float64)
        out = np.empty(len(a))
        lib.emu_test_fastmath(op, a.ctypes.data, None if b is None else b.ctypes.data, out.ctypes.data, len(a))
        return out
    return probe


def test_emulated_fastmath_helpers_have_the_device_special_cases(emu):
    """round 1's HLLC_lm defect (fsqrt(0) = NaN on the device, sqrt(0) = 0 under the emulator) was invisible here: the
    host build now runs the device's own Newton sequences on an emulated MUFU seed, special operands included.  The same
    checks run against the device in tests/test_gpu_hydro.py."""
    from test_gpu_hydro import check_fastmath, check_hllc_lm_at_rest
    check_fastmath(_emu_probe(emu))
    check_hllc_lm_at_rest(_emu_probe(emu))

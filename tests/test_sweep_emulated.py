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
    so = os.path.join(EMU_DIR, "libsweep_emu.so")
    src = os.path.join(EMU_DIR, "sweep_emu.cpp")
    hdrs = [os.path.join(HERE, "..", "pyro2_b200", "csrc", h) for h in ("sweep_task.cuh", "hydro_core.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(p) > os.path.getmtime(so) for p in [src] + hdrs):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                               "-Wno-unknown-pragmas", "-pthread", src, "-o", so])
    lib = C.CDLL(so)
    lib.emu_compressible_sweep.argtypes = ([C.c_void_p, C.c_void_p] + [C.c_int] * 4 + [C.c_longlong] +
                                           [C.c_double] * 8 + [C.c_int] * 5 + [C.c_void_p] * 2)
    return lib


def _emu_step(lib, U, ng, dx, dy, dt, prm, seglen):
    P = oracle.to_planes(U)
    _, qx, qy = P.shape
    pitch = (qy + 15) // 16 * 16
    Pin = np.zeros((4, qx, pitch))
    Pin[:, :, :qy] = P
    Pout = Pin.copy()
    scratch = np.zeros(8, dtype=np.uint64)
    lib.emu_compressible_sweep(Pin.ctypes.data, Pout.ctypes.data, qx - 2 * ng, qy - 2 * ng, ng, pitch, qx * pitch,
                               dx, dy, dt, prm.gamma, prm.z0, prm.z1, prm.delta, prm.cvisc, prm.limiter,
                               prm.use_flattening, prm.no_avisc_xhi, prm.no_avisc_yhi, seglen, scratch.ctypes.data, None)
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

"""pytest configuration: `gpu` marker (tests that need a real B200) and shared helpers.

CPU tests (`-m "not gpu"`) cover the oracle vs golden fixtures / live reference, the host logic,
the emulated sweep kernel and the C-ABI export list.  GPU tests (`-m gpu`) are the parity tests
proper: CUDA path vs oracle through the C ABI.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    n = np.linalg.norm(b.ravel())
    return np.linalg.norm((a - b).ravel()) / (n if n > 0 else 1.0)


def state_errors(U, ref, gamma=1.4):
    """per-variable relative L2 errors of a conserved state [i, j, (dens, ener, xmom, ymom)] against `ref`.
    A momentum component that is pure round-off noise in the reference (a static atmosphere: |m| ~ 1e-16) has
    no meaningful relative error of its own; it is then measured against the acoustic momentum scale rho*cs."""
    U = np.asarray(U, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    dens, ener, xmom, ymom = (ref[..., n] for n in range(4))
    p = (ener - 0.5 * (xmom ** 2 + ymom ** 2) / dens) * (gamma - 1.0)
    acoustic = np.linalg.norm((dens * np.sqrt(np.abs(gamma * p / dens))).ravel())
    errs = []
    for n in range(4):
        scale = np.linalg.norm(ref[..., n].ravel())
        if n >= 2 and scale < 1e-6 * acoustic:
            scale = acoustic
        errs.append(np.linalg.norm((U[..., n] - ref[..., n]).ravel()) / (scale if scale > 0 else 1.0))
    return errs


def make_state(nx, ny, ng, kind, seed=0, gamma=1.4):
    """synthetic conserved state [i, j, n] with ghosts (not yet filled consistently)"""
    rng = np.random.default_rng(seed)
    qx, qy = nx + 2 * ng, ny + 2 * ng
    x = (np.arange(qx) + 0.5 - ng) / nx
    y = (np.arange(qy) + 0.5 - ng) / ny
    X, Y = np.meshgrid(x, y, indexing="ij")
    if kind == "smooth":
        rho = 1 + 0.3 * np.sin(2 * np.pi * X) * np.cos(2 * np.pi * Y)
        u = 0.5 * np.cos(2 * np.pi * Y)
        v = -0.3 * np.sin(2 * np.pi * X)
        p = 1 + 0.2 * np.cos(2 * np.pi * (X + Y))
    elif kind == "shock":
        r = np.sqrt((X - 0.4) ** 2 + (Y - 0.55) ** 2)
        rho = np.where(r < 0.2, 2.0, 1.0) + 0.05 * rng.standard_normal(X.shape)
        u = 0.8 * np.where(X < 0.5, 1, -1) + 0.05 * rng.standard_normal(X.shape)
        v = 0.3 * np.where(Y < 0.6, 1.0, -0.7) + 0.05 * rng.standard_normal(X.shape)
        p = np.where(r < 0.2, 10.0, 0.1)
    elif kind == "sedov":
        r = np.sqrt((X - 0.5) ** 2 + (Y - 0.5) ** 2)
        rho = np.ones_like(X)
        u = np.zeros_like(X)
        v = np.zeros_like(X)
        p = np.where(r < 0.08, 50.0, 1.e-5)
    else:
        raise ValueError(kind)
    U = np.zeros((qx, qy, 4))
    U[..., 0] = rho
    U[..., 2] = rho * u
    U[..., 3] = rho * v
    U[..., 1] = p / (gamma - 1.0) + 0.5 * rho * (u * u + v * v)
    return U

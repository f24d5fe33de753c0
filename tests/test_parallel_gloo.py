"""CPU, world_size 2 and 3 over gloo: the slab halo exchange (pyro2_b200/parallel.py) reproduces the
single-domain ghost rows, including the periodic wrap and the 2-rank case where both neighbours are
the same peer; allreduce_max_ gives the global wave-speed maxima."""
import os
import socket

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

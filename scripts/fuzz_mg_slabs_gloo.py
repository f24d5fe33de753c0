"""Randomised check of the x-slab multigrid (deep-halo exchange per blocked-smoother pass, replicated coarse levels)
against the single-domain solver, through the product's CellCenterMG2d on the emulated device over gloo: 2 or 4 ranks,
n in {128, 256}, split levels, every boundary combination, Helmholtz terms, inhomogeneous Dirichlet values.
Development tool (CPU only):

    python scripts/fuzz_mg_slabs_gloo.py [ncases] [seed]
"""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def worker(rank, size, port, case, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), ROOT]
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        import emu_device
        from pyro2_b200.multigrid import MG
        from pyro2_b200.parallel import SlabDecomposition
        n, bc, alpha, beta, split, seed, inhom = case
        kw = dict(xl_BC_type=bc[0], xr_BC_type=bc[1], yl_BC_type=bc[2], yr_BC_type=bc[3], alpha=alpha, beta=beta)
        rng = np.random.default_rng(seed)
        coef = rng.standard_normal(8)
        if inhom:
            for side, name in zip(bc, ("xl_BC", "xr_BC", "yl_BC", "yr_BC")):
                if side == "dirichlet":
                    a, b = rng.standard_normal(2)
                    kw[name] = (lambda a_, b_: (lambda s: a_ + b_ * np.sin(3.0 * s)))(a, b)

        def rhs(x, y):
            f = coef[0] * torch.sin(2 * np.pi * x) * torch.cos(4 * np.pi * y) + coef[1] * torch.cos(6 * np.pi * x) * torch.sin(2 * np.pi * y)
            return f + (coef[2] * (x - 0.5) * (y - 0.3) if "dirichlet" in bc or alpha != 0.0 else 0.0)
        with emu_device.emulated_device():
            a = MG.CellCenterMG2d(n, n, decomposition=SlabDecomposition(), split_n=split, **kw)
            a.init_zeros()
            a.init_RHS(rhs(a.x2d.t(), a.y2d.t()))
            a.solve(rtol=1.e-10)
            g = a.soln_grid
            mine = a.get_solution().t()[g.ilo:g.ihi + 1, g.jlo:g.jhi + 1].contiguous()
            parts = [torch.empty_like(mine) for _ in range(size)] if rank == 0 else None
            dist.gather(mine, parts, dst=0)
            res = None
            if rank == 0:
                full = torch.cat(parts, dim=0).numpy()
                b = MG.CellCenterMG2d(n, n, **kw)
                b.init_zeros()
                b.init_RHS(rhs(b.x2d.t(), b.y2d.t()))
                b.solve(rtol=1.e-10)
                one = b.get_solution().numpy()[1:-1, 1:-1]
                res = (bool(np.array_equal(full, one)) and a.num_cycles == b.num_cycles, a.num_cycles, b.num_cycles,
                       float(np.abs(full - one).max()))
        dist.barrier()
        if rank == 0:
            q.put(res)
    finally:
        dist.destroy_process_group()


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


if __name__ == "__main__":
    ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    ctx = mp.get_context("spawn")
    bad = 0
    for c in range(ncases):
        size = int(rng.choice([2, 4]))
        n = int(rng.choice([128, 256], p=[0.7, 0.3]))
        split = int(rng.choice([s for s in (32, 64, 128) if s <= n and s // size >= 8]))

        def pair():
            a = str(rng.choice(["dirichlet", "neumann", "periodic"]))
            return (a, a) if a == "periodic" else (a, str(rng.choice(["dirichlet", "neumann"])))
        bc = pair() + pair()
        alpha, beta = (0.0, -1.0) if rng.integers(2) else (float(rng.choice([1.0, 2.5])), float(rng.choice([0.05, 1e-4, 3.0])))
        if alpha == 0.0 and "dirichlet" not in bc:
            alpha, beta = 1.0, 0.05                 # keep the problem non-singular
        case = (n, bc, alpha, beta, split, int(rng.integers(1 << 30)), bool(rng.integers(2)))
        q = ctx.Queue()
        port = free_port()
        procs = [ctx.Process(target=worker, args=(r, size, port, case, q)) for r in range(size)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(900)
        ok = all(p.exitcode == 0 for p in procs)
        res = q.get(timeout=5) if ok else None
        if not ok or not res[0]:
            bad += 1
            print("FAIL", c, dict(size=size, n=n, bc=bc, alpha=alpha, beta=beta, split=split, inhom=case[6], res=res), flush=True)
        else:
            print("ok  ", c, dict(size=size, n=n, bc=bc, alpha=alpha, beta=beta, split=split, inhom=case[6], cycles=res[1]), flush=True)
    print(f"{ncases} cases, {bad} failed")
    sys.exit(1 if bad else 0)

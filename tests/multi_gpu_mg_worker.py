"""torchrun worker used by test_gpu_multi.py: solves the same elliptic problem with the x-slab
decomposed multigrid on WORLD_SIZE GPUs and on one GPU, and compares the gathered solution
BIT FOR BIT (plus cycle counts)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def rhs(kind, x, y):
    if kind == "periodic":
        return torch.sin(2 * np.pi * x) * torch.cos(4 * np.pi * y)
    return -2.0 * ((1.0 - 6.0 * x ** 2) * y ** 2 * (1.0 - y ** 2) + (1.0 - 6.0 * y ** 2) * x ** 2 * (1.0 - x ** 2))


def main():
    kind, n, split = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
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

    a = MG.CellCenterMG2d(n, n, decomposition=SlabDecomposition(), split_n=split, **kw)
    a.init_zeros()
    a.init_RHS(rhs(kind, a.x2d.t(), a.y2d.t()))
    a.solve(rtol=1.e-11)
    g = a.soln_grid
    mine = a.get_solution().t()[g.ilo:g.ihi + 1, g.jlo:g.jhi + 1].contiguous()
    parts = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, parts, dst=0)
    ok = True
    if rank == 0:
        full = torch.cat(parts, dim=0).cpu().numpy()
        b = MG.CellCenterMG2d(n, n, **kw)
        b.init_zeros()
        b.init_RHS(rhs(kind, b.x2d.t(), b.y2d.t()))
        b.solve(rtol=1.e-11)
        one = b.get_solution().numpy()[1:-1, 1:-1]
        same = np.array_equal(full, one)
        print(f"MULTI_GPU_MG world={world} kind={kind} n={n} bit_identical={same} cycles={a.num_cycles}/{b.num_cycles} "
              f"resid={a.residual_error:.3e}/{b.residual_error:.3e} maxabs={np.abs(full - one).max():.3e}", flush=True)
        ok = same and a.num_cycles == b.num_cycles
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

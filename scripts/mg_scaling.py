"""V-cycle time of the decomposed multigrid at 4096^2 (torchrun) -- development aid"""
import os, sys
sys.path.insert(0, ".")
import torch, torch.distributed as dist
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
from pyro2_b200.multigrid import MG
from pyro2_b200.parallel import SlabDecomposition
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
split = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
a = MG.CellCenterMG2d(n, n, decomposition=SlabDecomposition() if world > 1 else None, split_n=split)
x, y = a.x2d.t(), a.y2d.t()
a.init_zeros()
a.init_RHS(-2.0 * ((1.0 - 6.0 * x ** 2) * y ** 2 * (1.0 - y ** 2) + (1.0 - 6.0 * y ** 2) * x ** 2 * (1.0 - x ** 2)))
a.max_cycles = 3; a.solve(rtol=0.0)
a.max_cycles = 10
if world > 1: dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); a.solve(rtol=0.0); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.num_cycles
if world > 1:
    t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t)
if rank == 0: print(f"MG_SCALING world={world} n={n} split={split} ms_per_cycle={ms:.3f} resid={a.residual_error:.3e} graph={a._graph is not None} err={a._graph_error}", flush=True)
if world > 1: dist.destroy_process_group()

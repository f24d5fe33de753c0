"""torchrun worker used by test_gpu_multi.py: runs a decomposed compressible problem on WORLD_SIZE
GPUs, gathers the slabs on rank 0 and compares them with the single-domain run BIT FOR BIT."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    problem, nx, ny, nsteps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    from pyro2_b200.parallel import SlabDecomposition
    from pyro2_b200.pyro_sim import Pyro
    inputs = {"mesh.nx": nx, "mesh.ny": ny, "driver.max_steps": 10 ** 6, "driver.tmax": 1e9}
    if problem == "sedov":
        inputs["sedov.r_init"] = 0.1

    p = Pyro("compressible")
    p.initialize_problem(problem, inputs_dict=inputs, decomposition=SlabDecomposition())
    dts = []
    for _ in range(nsteps):
        p.single_step()
        dts.append(p.sim.dt)
    p.sim.check_state()
    g = p.sim.cc_data.grid
    mine = p.sim.cc_data.planes[:, g.ilo:g.ihi + 1, g.jlo:g.jhi + 1].contiguous()
    parts = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, parts, dst=0)
    ok = True
    if rank == 0:
        full = torch.cat(parts, dim=1).cpu().numpy()
        q = Pyro("compressible")
        q.initialize_problem(problem, inputs_dict=inputs)
        dts1 = []
        for _ in range(nsteps):
            q.single_step()
            dts1.append(q.sim.dt)
        g1 = q.sim.cc_data.grid
        one = q.sim.cc_data.planes[:, g1.ilo:g1.ihi + 1, g1.jlo:g1.jhi + 1].cpu().numpy()
        same = np.array_equal(full, one)
        print(f"MULTI_GPU_PARITY world={world} problem={problem} bit_identical={same} "
              f"dt_identical={dts == dts1} maxabs={np.abs(full - one).max():.3e}", flush=True)
        ok = same and dts == dts1
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

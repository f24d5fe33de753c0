"""torchrun worker used by test_gpu_multi.py: runs a decomposed advection / burgers / incompressible problem on
WORLD_SIZE GPUs, gathers the slabs on rank 0 and compares every state plane and every dt with the single-domain
run BIT FOR BIT.  (tests/test_parallel_gloo.py runs the same comparison over gloo on the emulated device.)

    torchrun ... multi_gpu_flow_worker.py <solver> <problem> <nx> <ny> <nsteps> [key=value ...]
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _parse(v):
    for cast in (int, float):
        try:
            return cast(v)
        except ValueError:
            pass
    return v


def main():
    solver, problem, nx, ny, nsteps = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    extra = dict(a.split("=", 1) for a in sys.argv[6:])
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    from pyro2_b200.parallel import SlabDecomposition
    from pyro2_b200.pyro_sim import Pyro
    inputs = {"mesh.nx": nx, "mesh.ny": ny, "driver.max_steps": 10 ** 6, "driver.tmax": 1e9, "driver.verbose": 0}
    inputs.update({k: _parse(v) for k, v in extra.items()})

    def run(**kw):
        p = Pyro(solver)
        p.initialize_problem(problem, inputs_dict=inputs, **kw)
        dts = []
        for _ in range(nsteps):
            p.single_step()
            dts.append(p.sim.dt)
        g = p.sim.cc_data.grid
        return p.sim.cc_data.planes[:, g.ilo:g.ihi + 1, g.jlo:g.jhi + 1].contiguous(), dts

    mine, dts = run(decomposition=SlabDecomposition())
    parts = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, parts, dst=0)
    ok = True
    if rank == 0:
        full = torch.cat(parts, dim=1).cpu().numpy()
        one, dts1 = run()
        one = one.cpu().numpy()
        same = np.array_equal(full, one)
        print(f"MULTI_GPU_FLOW world={world} solver={solver} problem={problem} bit_identical={same} "
              f"dt_identical={dts == dts1} maxabs={np.abs(full - one).max():.3e}", flush=True)
        ok = same and dts == dts1
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

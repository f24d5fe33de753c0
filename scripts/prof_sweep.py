"""a few sweeps at 4096^2 (ncu target)"""
import sys
sys.path.insert(0, ".")
import torch
from pyro2_b200.pyro_sim import Pyro
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
p = Pyro("compressible")
p.initialize_problem("sedov", inputs_dict={"mesh.nx": n, "mesh.ny": n, "driver.max_steps": 10**9, "driver.tmax": 1e9})
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 5):
    p.single_step()
torch.cuda.synchronize()
print("done", p.sim.n, p.sim.dt)

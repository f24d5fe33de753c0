"""runs the fused coarse sub-V-cycle (64^2 hierarchy) a few times: target for ncu (development aid)"""
import sys
sys.path.insert(0, ".")
import torch
from pyro2_b200.mg_handle import MGHandle
d = MGHandle(64, ("dirichlet",) * 4, 0.0, -1.0, 0.0, 1.0, 0.0, 1.0, 10, 50)
d.plane(d.nlevels - 1, "f").fill_(1.0)
for _ in range(3):
    d.zero_coarse(); d.vcycle()
torch.cuda.synchronize()
print("ok")

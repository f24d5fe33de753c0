"""V-cycles at n^2 (ncu target)"""
import sys
sys.path.insert(0, ".")
import torch
from pyro2_b200.multigrid import MG
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
a = MG.CellCenterMG2d(n, n)
x, y = a.x2d.t(), a.y2d.t()
a.init_zeros()
a.init_RHS(-2.0 * ((1.0 - 6.0 * x ** 2) * y ** 2 * (1.0 - y ** 2) + (1.0 - 6.0 * y ** 2) * x ** 2 * (1.0 - x ** 2)))
a.max_cycles = int(sys.argv[2]) if len(sys.argv) > 2 else 2
a.solve(rtol=0.0)
torch.cuda.synchronize()
print("done", a.num_cycles, a.residual_error)

import sys, os
sys.path.insert(0, ".")
import torch
from pyro2_b200 import ops
n = 4096; ng = 4; q = n + 8
A = ops.alloc_planes(4, q, q)
x = (torch.arange(q, device="cuda", dtype=torch.float64) + 0.5 - ng)/n
X, Y = torch.meshgrid(x, x, indexing="ij")
r = torch.sqrt((X-0.5)**2 + (Y-0.5)**2)
A[0, :, :q] = 1.0; A[1, :, :q] = torch.where(r < 0.05, 50.0, 1e-5)/0.4
B = A.clone(); sc = ops.new_scratch(); prm = ops.comp_params(); bc = [("outflow",)*4]*4
ops.fill_ghost(A, n, n, ng, bc)
w = ops.cfl_wavemax(A, n, n, ng, 1.4, sc); dt = 0.8*min(1/n/w[0], 1/n/w[1])*0.01
for _ in range(30):
    ops.fill_ghost(A, n, n, ng, bc); ops.compressible_sweep(A, B, n, n, ng, 1/n, 1/n, dt, prm, sc); A, B = B, A
    w = sc[:2].view(torch.float64).tolist(); dt = min(2*dt, 0.8*min(1/n/w[0], 1/n/w[1]))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.compressible_sweep(A, B, n, n, ng, 1/n, 1/n, dt, prm, sc)
e1.record(); torch.cuda.synchronize()
print(os.environ.get("P2B_SO", "default"), "sweep kernel %.3f ms" % (e0.elapsed_time(e1)/10), ops.sweep_info())

"""first-light timing of the two hot paths (development aid, not the bench)"""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from pyro2_b200 import ops
from pyro2_b200.mg_handle import MGHandle

def time_sweep(n, steps=10):
    ng = 4
    q = n + 2*ng
    A = ops.alloc_planes(4, q, q)
    x = (torch.arange(q, device="cuda", dtype=torch.float64) + 0.5 - ng)/n
    X, Y = torch.meshgrid(x, x, indexing="ij")
    r = torch.sqrt((X-0.5)**2 + (Y-0.5)**2)
    A[0, :, :q] = 1.0
    A[1, :, :q] = torch.where(r < 0.05, 50.0, 1e-5)/0.4
    B = A.clone()
    scratch = ops.new_scratch()
    prm = ops.comp_params()
    bc = [("outflow",)*4]*4
    dx = 1.0/n
    def step(A, B, first=False):
        ops.fill_ghost(A, n, n, ng, bc)
        w = scratch[:2].view(torch.float64).tolist() if not first else ops.cfl_wavemax(A, n, n, ng, 1.4, scratch)
        dt = 0.8*min(dx/w[0], dx/w[1])*(0.01 if first else 1.0)
        ops.compressible_sweep(A, B, n, n, ng, dx, dx, dt, prm, scratch)
        return B, A
    A, B = step(A, B, True)
    for _ in range(3): A, B = step(A, B)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): A, B = step(A, B)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/steps
    print(f"sweep {n}^2: {ms:.3f} ms/step  {n*n/ms*1e3:.3e} cell-updates/s  hbm-algorithmic {n*n*64/ms*1e3/1e9:.1f} GB/s", ops.sweep_info(), "status", int(scratch[3]), flush=True)

def time_mg(n, cycles=5):
    d = MGHandle(n, ("dirichlet",)*4, 0.0, -1.0, 0.0, 1.0, 0.0, 1.0, 10, 50)
    L = d.nlevels - 1
    x = (torch.arange(n+2, device="cuda", dtype=torch.float64) - 0.5)/n
    X, Y = torch.meshgrid(x, x, indexing="ij")
    d.plane(L, "f").copy_(-2.0*((1-6*X**2)*Y**2*(1-Y**2) + (1-6*Y**2)*X**2*(1-X**2)))
    src = np.sqrt(d.sumsq(L, "f")/n/n)
    old = torch.zeros((n+2)*d.plane(L, "v").stride(0), dtype=torch.float64, device="cuda")
    d.zero_coarse(); d.vcycle(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(cycles):
        d.zero_coarse(); d.vcycle()
        rel, rs = d.cycle_diagnostics(old)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/cycles
    print(f"mg {n}^2: {ms:.3f} ms/V-cycle  {1e3/ms:.1f} V-cycles/s  resid {np.sqrt(rs/n/n)/src:.3e}", flush=True)

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    for n in (1024, 4096): time_sweep(n)
    for n in (1024, 4096): time_mg(n)

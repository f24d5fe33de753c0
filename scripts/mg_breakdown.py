"""event-timed breakdown of one 4096^2 V-cycle (development aid)"""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from pyro2_b200.mg_handle import MGHandle
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = MGHandle(n, ("dirichlet",)*4, 0.0, -1.0, 0.0, 1.0, 0.0, 1.0, 10, 50)
L = d.nlevels - 1
x = (torch.arange(n+2, device="cuda", dtype=torch.float64) - 0.5)/n
X, Y = torch.meshgrid(x, x, indexing="ij")
d.plane(L, "f").copy_(-2.0*((1-6*X**2)*Y**2*(1-Y**2) + (1-6*Y**2)*X**2*(1-X**2)))
old = torch.zeros((n+2)*d.plane(L, "v").stride(0), dtype=torch.float64, device="cuda")
def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps
print("vcycle (no sync between)      %.3f ms" % timeit(lambda: (d.zero_coarse(), d.vcycle())))
print("diagnostics                   %.3f ms" % timeit(lambda: d.cycle_diagnostics(old)))
for l in range(L, 4, -1):
    print("level %2d n=%5d smooth(10)  %.3f ms   residual %.3f  restrict %.3f  prolong %.3f" % (l, 2**(l+1), timeit(lambda: d.smooth(l, 10)), timeit(lambda: d.residual(l)), timeit(lambda: d.restrict(l)), timeit(lambda: d.prolong_correct(l))))
import ctypes as C
from pyro2_b200 import _lib
t0 = time.perf_counter()
for _ in range(20): d.zero_coarse(); d.vcycle()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("host time to enqueue one vcycle %.3f ms; drain %.3f ms" % ((t1-t0)/20*1e3, (t2-t1)*1e3))
d.set_blocking(False)
print("vcycle, one launch per colour   %.3f ms" % timeit(lambda: (d.zero_coarse(), d.vcycle()), 5))

"""event-timed per-kernel breakdown of the V-cycle at every level (development aid; B200)"""
import sys
sys.path.insert(0, ".")
import torch
from pyro2_b200.mg_handle import MGHandle

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = MGHandle(n, ("dirichlet",) * 4, 0.0, -1.0, 0.0, 1.0, 0.0, 1.0, 10, 50)
L = d.nlevels - 1
x = (torch.arange(n + 2, device="cuda", dtype=torch.float64) - 0.5) / n
X, Y = torch.meshgrid(x, x, indexing="ij")
d.plane(L, "f").copy_(-2.0 * ((1 - 6 * X ** 2) * Y ** 2 * (1 - Y ** 2) + (1 - 6 * Y ** 2) * X ** 2 * (1 - X ** 2)))
old = torch.zeros((n + 2) * d.plane(L, "v").stride(0), dtype=torch.float64, device="cuda")


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3     # us


def graphed(fn, inner=10):
    """time fn inside a CUDA graph holding `inner` copies (no launch gaps)"""
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(inner):
            fn()
    return timeit(g.replay, 10) / inner


print("whole cycle (zero_coarse + vcycle), eager: %.1f us; graphed: %.1f us" % (
    timeit(lambda: (d.zero_coarse(), d.vcycle()), 10), graphed(lambda: (d.zero_coarse(), d.vcycle()), 2)))
print("diagnostics graphed %.1f us" % graphed(lambda: d.cycle_diagnostics_enqueue(old), 4))
print("zero_coarse graphed %.1f us" % graphed(lambda: d.zero_coarse(), 4))
for l in range(L, 5, -1):
    nn = 2 ** (l + 1)
    print("level %2d n=%5d  tb pass(5) %.1f us  | graphed: pass(5) %.1f  pass(1) %.1f  residual %.1f  restrict %.1f  prolong %.1f" % (
        l, nn, timeit(lambda: d.tb_pass(l, "v", "w", 5)), graphed(lambda: d.tb_pass(l, "v", "w", 5)),
        graphed(lambda: d.tb_pass(l, "v", "w", 1)),
        graphed(lambda: d.residual(l)), graphed(lambda: d.restrict(l)), graphed(lambda: d.prolong_correct(l))))
for l in range(5, -1, -1):
    print("coarse sub-cycle from level %d (n=%d): eager %.1f us, graphed %.1f us" % (
        l, 2 ** (l + 1), timeit(lambda: d.vcycle_level(l)), graphed(lambda: d.vcycle_level(l), 4)))

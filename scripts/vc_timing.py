"""event-timed V-cycle of the variable-coefficient solver next to the constant-coefficient one
(development aid; writes gpurun_out/vc_timing.json)"""
import json, os, sys
sys.path.insert(0, ".")
import torch
from pyro2_b200.mg_handle import MGHandle

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
pi = 3.141592653589793
x = (torch.arange(n + 2, device="cuda", dtype=torch.float64) - 0.5) / n
X, Y = torch.meshgrid(x, x, indexing="ij")


def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {"n": n}
for kind in ("constant", "variable"):
    d = MGHandle(n, ("dirichlet",) * 4, 0.0, -1.0 if kind == "constant" else 0.0, 0.0, 1.0, 0.0, 1.0, 10, 50)
    L = d.nlevels - 1
    if kind == "variable":
        d.set_coeffs((2.0 + torch.cos(2 * pi * X) * torch.cos(2 * pi * Y)).contiguous(), ("neumann",) * 4)
    d.plane(L, "f").copy_(-16.0 * pi ** 2 * (torch.cos(2 * pi * X) * torch.cos(2 * pi * Y) + 1) * torch.sin(2 * pi * X) * torch.sin(2 * pi * Y))
    old = torch.zeros((n + 2) * d.plane(L, "v").stride(0), dtype=torch.float64, device="cuda")
    r = {"vcycle_ms": timeit(lambda: (d.zero_coarse(), d.vcycle())),
         "diagnostics_ms": timeit(lambda: d.cycle_diagnostics(old)),
         "smooth10_finest_ms": timeit(lambda: d.smooth(L, 10)),
         "residual_finest_ms": timeit(lambda: d.residual(L))}
    d.set_blocking(False)
    r["vcycle_per_colour_ms"] = timeit(lambda: (d.zero_coarse(), d.vcycle()), 3)
    r["smooth10_finest_per_colour_ms"] = timeit(lambda: d.smooth(L, 10), 3)
    out[kind] = r
    print(kind, r)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/vc_timing.json", "w"), indent=1)

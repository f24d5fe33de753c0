"""time per step of the incompressible shear problem and its split into explicit stages and multigrid
projections (development aid; writes gpurun_out/incomp_timing.json)"""
import json, os, sys, time
sys.path.insert(0, ".")
import torch
from pyro2_b200.pyro_sim import Pyro

out = {}
for n in [int(a) for a in sys.argv[1:]] or [1024, 4096]:
    p = Pyro("incompressible")
    t0 = time.perf_counter()
    p.initialize_problem("shear", inputs_dict={"mesh.nx": n, "mesh.ny": n, "driver.max_steps": 1000})
    torch.cuda.synchronize()
    t_init = time.perf_counter() - t0
    sim = p.sim
    for _ in range(2):
        p.single_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    K = 5
    for _ in range(K):
        p.single_step()
    torch.cuda.synchronize()
    step = (time.perf_counter() - t0) / K
    mg = next(iter(sim._mg.values()))[0]
    # the explicit stages alone, event-timed
    P = sim._planes()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f = sim._flow
    e0.record()
    for _ in range(10):
        f.interface_states(P["x-velocity"], P["y-velocity"], P["gradp_x"], P["gradp_y"], sim.dt, 2)
        f.mac_vels(); f.upwind_states()
    e1.record(); torch.cuda.synchronize()
    out[n] = {"init_s": t_init, "step_ms": step * 1e3, "zones_per_s": n * n / step, "last_solve_cycles": mg.num_cycles,
              "explicit_states_ms": e0.elapsed_time(e1) / 10, "dt": sim.dt, "t": sim.cc_data.t}
    print(n, out[n])
    del p, sim, mg, f, P
    torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/incomp_timing.json", "w"), indent=1)

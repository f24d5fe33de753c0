#!/usr/bin/env python
"""bench.py -- headline benchmark of the two hot paths on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--nx 4096] [--skip-cpu]

Workload (BASELINE.json): compressible Sedov HLLC 4096^2 fp64 -- metric ``cell-updates/s`` -- with
the multigrid constant-coefficient Poisson 4096^2 V-cycle rate reported beside it in ``"mg"``.
A "step" is one pass of the driver loop over the whole grid through the public API
(``Pyro.single_step()``: fill_BC_all -> compute_timestep -> evolve, pyro/pyro_sim.py:241-256).

One JSON line on stdout (rank 0).  ``value``: state resident in HBM.  ``e2e``: the same step with
the state pushed from pinned host memory before and pulled back after every step.  ``roofline``:
algorithmic bytes (64 B per cell update) / CUDA-event time of the sweep kernel alone,
against the measured copy bandwidth in MEASURED_PEAKS.json.  ``cpu_baseline``: the oracle port
(oracle/pyro_oracle.c, OpenMP) on this box's host cores on a bounded sample.

N > 1 (torchrun, one rank per GPU): the domain is split into x-slabs with a 4-row halo exchanged
over NCCL each step (weak scaling: every rank owns an nx x ny block).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nx", type=int, default=4096, help="zones per side (per GPU)")
    ap.add_argument("--mg-cycles", type=int, default=10)
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--skip-mg", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-incomp", action="store_true", help="skip the incompressible-shear leg (N = 1 only)")
    ap.add_argument("--incomp-nx", type=int, default=2048)
    ap.add_argument("--incomp-multi", action="store_true",
                    help="N > 1: also run the incompressible leg, on x-slabs of the same global problem (strong scaling; "
                         "off by default until the decomposed flow solvers have been run over NCCL)")
    return ap.parse_args()


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except OSError:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 20 ms; the median is taken over the samples
    that arrived between mark_start() and mark_end() (the timed region)"""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.proc = None
        self.lines = []
        self.index = index
        self.t0 = self.t1 = None

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0 = self.t0 or 0.0
        t1 = (self.t1 or time.time()) + 0.03
        power = []
        for stamp, ln in self.lines:
            p = [s.strip() for s in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                clk, cmax, pw = float(p[0]), float(p[1]), float(p[2])
            except ValueError:
                continue
            mx.append(cmax)
            if stamp < t0 or stamp > t1:
                continue          # outside the timed region
            sm.append(clk)
            power.append(pw)
            for nm, val in zip(names, p[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "power_w_max": max(power) if power else None, "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------
# CPU legs (oracle port).  The only places bench.py touches oracle/.
# --------------------------------------------------------------------------------------------
def sedov_planes_numpy(n, ng=4, gamma=1.4):
    import numpy as np
    q = n + 2 * ng
    x = (np.arange(q) + 0.5 - ng) / n
    r2 = (x[:, None] - 0.5) ** 2 + (x[None, :] - 0.5) ** 2
    P = np.zeros((4, q, q))
    P[0] = 1.0
    P[1] = np.where(r2 < 0.01 ** 2 + (1.0 / n) ** 2, 1.0 / (np.pi * 0.01 ** 2), 1.e-5) / (gamma - 1.0)
    return P


def cpu_compressible(n, steps):
    """oracle port of the compressible step on the host cores: cell-updates/s"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import oracle
    ng = 4
    P = sedov_planes_numpy(n)
    dx = 1.0 / n
    bc = ("outflow",) * 4
    prm = oracle.comp_params()

    def step(first):
        for k in range(4):
            oracle.fill_ghost(P[k], ng, bc)
        dt = oracle.lib().orc_cfl_dt(P.ctypes.data, n, n, ng, dx, dx, 1.4, 0.8) * (0.01 if first else 1.0)
        oracle.compressible_step(P, ng, dx, dx, dt, prm, planes=True)
    step(True)                     # warm-up (page faults, OpenMP pool)
    t0 = time.perf_counter()
    for _ in range(steps):
        step(False)
    dt = time.perf_counter() - t0
    assert np.isfinite(P).all()
    return n * n * steps / dt, dt


def cpu_mg(n, cycles):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import oracle
    o = oracle.MG(n)
    x = (np.arange(n + 2) - 0.5) / n
    X, Y = np.meshgrid(x, x, indexing="ij")
    o.init_zeros()
    o.init_RHS(-2.0 * ((1.0 - 6.0 * X ** 2) * Y ** 2 * (1.0 - Y ** 2) + (1.0 - 6.0 * Y ** 2) * X ** 2 * (1.0 - X ** 2)))
    o.v_cycle()
    t0 = time.perf_counter()
    for _ in range(cycles):
        o.v_cycle()
    dt = time.perf_counter() - t0
    return cycles / dt, dt


def cpu_incompressible(n):
    """one step of the incompressible shear problem through the oracle (explicit stages + both projections)"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import math
    import numpy as np
    import oracle
    ng = 4
    c = (np.arange(n + 2 * ng) + 0.5 - ng) / n
    X, Y = np.meshgrid(c, c, indexing="ij")
    P = np.zeros((6, n + 2 * ng, n + 2 * ng))
    P[0] = np.where(Y <= 0.5, np.tanh(42.0 * (Y - 0.25)), np.tanh(42.0 * (0.75 - Y)))
    P[1] = 0.05 * np.sin(2.0 * math.pi * X)
    dt = 0.8 * min(1.0 / n / np.abs(P[0]).max(), 1.0 / n / np.abs(P[1]).max())
    for k in range(6):
        oracle.fill_ghost(P[k], ng, ("periodic",) * 4)
    t0 = time.perf_counter()
    cyc = oracle.incomp_evolve(P, ng, dt)
    return time.perf_counter() - t0, cyc


def host_threads():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path.  pyro2 is Python and cannot
    travel to the GPU box, so this is the oracle port (oracle/pyro_oracle.c, bit-identical to the
    reference per stage, OpenMP over all host threads) on a bounded sample of the same workload."""
    if rank != 0:
        return
    n = min(args.nx, 2048)
    steps = max(1, min(args.steps, 2))
    rate, secs = cpu_compressible(n, steps)
    line = {
        "impl": "reference", "metric": "cell-updates/s", "value": rate, "unit": "cell-updates/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": 1, "ms_per_step": secs / steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"compressible Sedov HLLC {args.nx}^2 fp64 (CPU sample: {n}^2 zones)",
                   "note": "oracle port of the reference path (pyro2 itself is Python; not present on this box)"},
        "cpu_baseline": {"value": rate, "unit": "cell-updates/s", "cores": host_threads(), "kind": "port",
                         "sample": f"{steps} step(s) of the {n}^2 Sedov state after 1 warm-up step"},
        "e2e": {"value": rate, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_RESULT_FD = None


def guard_stdout():
    """stdout carries exactly one JSON line.  Libraries loaded below (NCCL's version banner, nvcc
    during a first build) write to file descriptor 1 behind Python's back, so point fd 1 at stderr for
    the whole run and keep a private duplicate of the real stdout for the result line."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


# --------------------------------------------------------------------------------------------
def main():
    args = parse()
    guard_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    # keep stdout to the single JSON line: at NCCL_DEBUG=VERSION or WARN (from the environment or an
    # nccl.conf) NCCL printf()s its version banner to stdout; an unrecognised level silences it and,
    # being an environment variable, takes precedence over a conf file
    os.environ["NCCL_DEBUG"] = "NONE"
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; pyro2_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import __graft_entry__
    __graft_entry__.build()
    from pyro2_b200 import ops
    from pyro2_b200.parallel import SlabDecomposition
    from pyro2_b200.pyro_sim import Pyro

    n = args.nx
    K, W = args.steps, max(args.warmup, 3)
    peaks, peak_kind = measured_peaks()

    # ---- the compressible Sedov problem through the public API --------------------------------
    p = Pyro("compressible")
    inputs = {"mesh.nx": n * world, "mesh.ny": n, "mesh.xmax": float(world), "driver.max_steps": 10 ** 9,
              "driver.tmax": 1.e9}
    slab = SlabDecomposition(rank, world) if world > 1 else None
    p.initialize_problem("sedov", inputs_dict=inputs, **({"decomposition": slab} if slab else {}))
    sim = p.sim

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(W):
        p.single_step()
    sim.check_state()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.mark_start()
    e0.record()
    for _ in range(K):
        p.single_step()
    e1.record()
    barrier()
    sampler.mark_end()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    sim.check_state()
    if world > 1:
        tms = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms)
    value = n * n * world * K / (ms * 1e-3)
    launches_per_step = 3     # fill_x_kernel, fill_y_kernel, sweep_kernel (the halo exchange is NCCL's)

    # ---- the sweep kernel alone (roofline numerator's denominator) ----------------------------
    g = sim.cc_data.grid
    prm = sim._comp_params()
    A, B = sim.cc_data.planes, sim._alt_planes
    ks, ke = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms = 0.0
    for _ in range(K):
        sim.cc_data.fill_BC_all()
        ks.record()
        ops.compressible_sweep(A, B, g.nx, g.ny, g.ng, g.dx, g.dy, float(sim.dt), prm, sim._scratch)
        ke.record()
        ke.synchronize()
        kms += ks.elapsed_time(ke)
        A, B = B, A
    kms /= K
    alg_bytes = 64.0 * g.nx * g.ny
    achieved = alg_bytes / (kms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "sweep_traffic.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except OSError:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": f"of {peak_kind}",
                "kernel": "pyro::sweep_kernel", "kernel_ms": kms, "algorithmic_bytes_per_launch": alg_bytes,
                "note": "the fused sweep is FP64-pipe bound (~1.3 k DP instructions per cell update, DESIGN.md); "
                        "this is its algorithmic-byte rate against the HBM copy peak"}

    # ---- e2e: host buffers, H2D before and D2H after every step --------------------------------
    e2e = None
    if not args.skip_e2e:
        planes = sim.cc_data.planes
        host_in = torch.empty(planes.shape, dtype=planes.dtype, pin_memory=True)
        host_out = torch.empty(planes.shape, dtype=planes.dtype, pin_memory=True)
        host_in.copy_(planes)
        nbytes = planes.numel() * 8
        ke2 = min(K, 5)
        barrier()
        e0.record()
        for _ in range(ke2):
            sim.cc_data.planes.copy_(host_in, non_blocking=True)
            sim.cc_data.version += 1          # the state was replaced: forces the stand-alone CFL kernel
            p.single_step()
            host_out.copy_(sim.cc_data.planes, non_blocking=True)
        e1.record()
        barrier()
        ems = e0.elapsed_time(e1)
        if world > 1:
            tms = torch.tensor([ems], device="cuda", dtype=torch.float64)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ems = float(tms)
        e2e = {"value": n * n * world * ke2 / (ems * 1e-3), "unit": "cell-updates/s",
               "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes, "steps": ke2,
               "note": "Pyro.single_step() with the full state copied from/to pinned host memory every step"}
        del host_in, host_out

    # ---- multigrid V-cycles: the SAME 4096^2 problem on all N GPUs (strong scaling) -----------------
    mg = None
    if not args.skip_mg:
        del p, sim, A, B
        torch.cuda.empty_cache()
        from pyro2_b200.multigrid import MG
        a = MG.CellCenterMG2d(n, n, decomposition=slab, split_n=1024)
        x = a.x2d.t()
        y = a.y2d.t()
        a.init_zeros()
        a.init_RHS(-2.0 * ((1.0 - 6.0 * x ** 2) * y ** 2 * (1.0 - y ** 2) + (1.0 - 6.0 * y ** 2) * x ** 2 * (1.0 - x ** 2)))
        a.max_cycles = 3
        a.solve(rtol=0.0)                       # warm-up cycles (also captures the CUDA graph at N = 1)
        a.max_cycles = args.mg_cycles
        barrier()
        e0.record()
        a.solve(rtol=0.0)                       # exactly mg_cycles V-cycles through the public API
        e1.record()
        barrier()
        mms = e0.elapsed_time(e1)
        if world > 1:
            tms = torch.tensor([mms], device="cuda", dtype=torch.float64)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            mms = float(tms)
        mms /= a.num_cycles
        mg_bytes = 776.0 * n * n                # one-pass-per-operator model
        mg = {"metric": "V-cycles/s", "value": 1e3 / mms, "unit": "V-cycles/s", "ms_per_cycle": mms,
              "cycles": a.num_cycles, "residual_error": a.residual_error, "scaling": "strong", "n_gpus": world,
              "config": {"workload": f"multigrid constant-coefficient Poisson {n}^2 fp64 (global), dirichlet, nsmooth 10/50",
                         "parallelism": (f"x-slabs x{world} on levels >= 1024^2, coarser levels replicated" if world > 1
                                         else "single GPU, cycle replayed as a CUDA graph")},
              "roofline": {"bound": "hbm", "achieved": mg_bytes / (mms * 1e-3) / 1e9 / world, "peak": peaks["hbm_gbs"],
                           "unit": "GB/s per GPU", "frac": mg_bytes / (mms * 1e-3) / 1e9 / world / peaks["hbm_gbs"],
                           "traffic": None, "algorithmic_bytes_per_cycle": mg_bytes,
                           "note": "776 B per finest cell per V-cycle is the one-pass-per-operator model; the temporally "
                                   "blocked smoother moves ~40% of it"}}
        del a

    # ---- incompressible shear 2048^2 (BASELINE config 4): explicit stages + two multigrid projections ----
    incomp = None
    if (world == 1 or args.incomp_multi) and not args.skip_incomp:
        torch.cuda.empty_cache()
        ni = args.incomp_nx
        pi = Pyro("incompressible")
        pi.initialize_problem("shear", inputs_dict={"mesh.nx": ni, "mesh.ny": ni, "driver.max_steps": 10 ** 9,
                                                    "driver.tmax": 1.e9}, **({"decomposition": slab} if slab else {}))
        isim = pi.sim
        pi.single_step()
        barrier()
        ki = 3
        cyc = 0
        e0.record()
        for _ in range(ki):
            pi.single_step()
        e1.record()
        barrier()
        ims = e0.elapsed_time(e1) / ki
        if world > 1:
            tms = torch.tensor([ims], device="cuda", dtype=torch.float64)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ims = float(tms)
        solver = next(iter(isim._mg.values()))[0]
        incomp = {"metric": "zone-updates/s", "value": ni * ni / (ims * 1e-3), "unit": "zone-updates/s", "ms_per_step": ims,
                  "steps": ki, "n_gpus": world, "scaling": "strong",
                  "config": {"workload": f"incompressible shear {ni}^2 fp64 (global), periodic, limiter 2, proj_type 2",
                                          "note": "each step = p2b_flow_* explicit stages + 2 multigrid projections at rtol 1e-12; "
                                                  "at this size the reference's own stopping rule runs both to max_cycles = 100"},
                  "v_cycles_last_solve": solver.num_cycles, "gpu_launches_explicit": 14}
        if rank == 0 and not args.skip_cpu:
            nc = min(ni, 1024)
            secs, cyc = cpu_incompressible(nc)
            incomp["cpu_baseline"] = {"value": nc * nc / secs, "unit": "zone-updates/s", "cores": host_threads(), "kind": "port",
                                      "sample": f"1 step of the {nc}^2 shear problem ({cyc[0]} + {cyc[1]} V-cycles, {secs:.1f} s)"}
        del pi, isim, solver

    # ---- CPU baseline (rank 0, N = 1) ------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        ncpu = min(n, 2048)
        rate, secs = cpu_compressible(ncpu, 2)
        cpu = {"value": rate, "unit": "cell-updates/s", "cores": host_threads(), "kind": "port",
               "sample": f"2 steps of the {ncpu}^2 Sedov state after 1 warm-up step ({secs:.1f} s)"}
        if mg is not None:
            mrate, msecs = cpu_mg(ncpu, 3)
            mg["cpu_baseline"] = {"value": mrate, "unit": "V-cycles/s", "cores": host_threads(), "kind": "port",
                                  "sample": f"3 V-cycles at {ncpu}^2 ({msecs:.1f} s)"}

    if rank == 0:
        line = {
            "metric": "cell-updates/s", "value": value, "unit": "cell-updates/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"compressible Sedov HLLC {n}^2 fp64 per GPU (limiter 2, flattening, cvisc 0.1, outflow)",
                       "global_zones": [n * world, n], "parallelism": f"x-slabs x{world}" if world > 1 else "single GPU",
                       "l2": "state (2 x 539 MB at 4096^2) is larger than L2; no flush needed",
                       "sweep": ops.sweep_info()},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * K,
            "roofline": roofline, "cpu_baseline": cpu, "mg": mg, "incompressible": incomp,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- headline benchmark of the two hot paths on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--nx 4096] [--skip-...]

Workload (BASELINE.json): compressible Sedov HLLC 4096^2 fp64 -- metric ``cell-updates/s`` -- with the multigrid
constant-coefficient Poisson 4096^2 V-cycle rate in ``"mg"``, the incompressible shear 2048^2 step in
``"incompressible"`` (config 4), the 16384^2 Sedov on 8 slabs in ``"sedov_16384"`` (config 5, N = 8) and, at N > 1, a
decomposed-vs-single-domain bit comparison in ``"parity"``.  A "step" is one pass of the driver loop over the whole grid
through the public API (``Pyro.single_step()``: fill_BC_all -> compute_timestep -> evolve, pyro/pyro_sim.py:241-256).

One JSON line on stdout (rank 0).  ``value``: state resident in HBM.  ``e2e``: the same step with the state living in
pinned HOST memory (``Pyro.single_step_streamed``: row blocks travel host -> device -> host while their neighbours are
swept), every byte of both copies inside the timed region.  ``roofline``: algorithmic bytes (64 B per cell update,
DESIGN.md) / CUDA-event time of the sweep kernel alone, against the measured copy bandwidth in MEASURED_PEAKS.json.
``cpu_baseline``: the oracle port (oracle/pyro_oracle.c, OpenMP, threads pinned) on this box's host cores on a bounded
sample.  ``--impl reference``: the same port on the SAME configuration, steps and warm-up as the GPU arm.

N > 1 (torchrun, one rank per GPU): x-slabs.  Compressible: 4-row halo over NCCL each step (weak scaling: every rank an
nx x ny block).  Multigrid: the same 4096^2 problem on all N GPUs (strong scaling), halo rows pushed through peer memory
from the kernels' epilogues, no NCCL inside a cycle.
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
    ap.add_argument("--mg-cycles", type=int, default=20)
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--skip-mg", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-incomp", action="store_true", help="skip the incompressible-shear leg")
    ap.add_argument("--skip-parity", action="store_true", help="N > 1: skip the decomposed-vs-single-domain comparison")
    ap.add_argument("--skip-config5", action="store_true", help="N = 8: skip the 16384^2 Sedov leg")
    ap.add_argument("--config5", action="store_true", help="run the 16384^2 Sedov leg at this N > 1 too (rehearsal of the N = 8 leg)")
    ap.add_argument("--incomp-nx", type=int, default=2048)
    ap.add_argument("--ref-budget", type=float, default=200.0,
                    help="--impl reference: seconds of CPU time after which the number of timed steps is cut (>= 3 kept)")
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
def pin_host_threads():
    """OpenMP placement for the CPU legs: one thread per core, neighbours close, fixed for the run (unpinned 128-thread
    runs of the memory-bound port varied x3.6 between repeats in round 1).  Must run before any OpenMP runtime loads."""
    os.environ.setdefault("OMP_PROC_BIND", "close")
    os.environ.setdefault("OMP_PLACES", "cores")
    os.environ.setdefault("OMP_DYNAMIC", "false")


def sedov_planes_numpy(n, ng=4, gamma=1.4):
    import numpy as np
    q = n + 2 * ng
    x = (np.arange(q) + 0.5 - ng) / n
    r2 = (x[:, None] - 0.5) ** 2 + (x[None, :] - 0.5) ** 2
    P = np.zeros((4, q, q))
    P[0] = 1.0
    P[1] = np.where(r2 < 0.01 ** 2 + (1.0 / n) ** 2, 1.0 / (np.pi * 0.01 ** 2), 1.e-5) / (gamma - 1.0)
    return P


def cpu_compressible(n, steps, warmup=1, budget=None):
    """oracle port of the compressible step on the host cores.  Returns (cell-updates/s from the MEDIAN step time,
    list of step times, steps actually timed): every step is timed on its own, `budget` seconds bound the timed part
    (at least 3 steps are kept)."""
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
    for w in range(max(1, warmup)):
        step(w == 0)                   # warm-up (page faults, OpenMP pool)
    times = []
    for k in range(steps):
        t0 = time.perf_counter()
        step(False)
        times.append(time.perf_counter() - t0)
        if budget is not None and k + 1 >= 3 and sum(times) + statistics.median(times) > budget:
            break
    assert np.isfinite(P).all()
    return n * n / statistics.median(times), times, len(times)


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


_HOST_THREADS = None


def host_threads():
    """threads the CPU legs can use -- taken once, BEFORE an OpenMP runtime binds the calling thread to its first place
    (with OMP_PROC_BIND the main thread's affinity mask shrinks to one core and would be reported as `cores: 1`)"""
    global _HOST_THREADS
    if _HOST_THREADS is None:
        n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        env = os.environ.get("OMP_NUM_THREADS", "")
        _HOST_THREADS = min(n, int(env)) if env.isdigit() and int(env) > 0 else n
    return _HOST_THREADS


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path on the SAME configuration as the GPU arm
    (args.nx^2 zones, args.steps timed steps after args.warmup warm-up steps).  pyro2 is Python and cannot travel to the
    GPU box, so this is the oracle port (oracle/pyro_oracle.c, bit-identical to the reference per stage, OpenMP over all
    host threads, pinned).  The value is taken from the MEDIAN step time; --ref-budget bounds the timed part."""
    if rank != 0:
        return
    n, K, W = args.nx, args.steps, max(args.warmup, 1)
    rate, times, done = cpu_compressible(n, K, warmup=W, budget=args.ref_budget)
    med = statistics.median(times)
    line = {
        "impl": "reference", "metric": "cell-updates/s", "value": rate, "unit": "cell-updates/s",
        "n_gpus": args.gpus, "steps": done, "warmup": W, "ms_per_step": med * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"compressible Sedov HLLC {n}^2 fp64 per GPU (limiter 2, flattening, cvisc 0.1, outflow)",
                   "global_zones": [n, n],
                   "note": "oracle port of the reference path on the host cores (pyro2 itself is Python and is not on this "
                           "box); same zones per step as the GPU arm; value from the median step time",
                   "step_seconds": {"min": min(times), "median": med, "max": max(times)},
                   "steps_requested": K, "omp": {k: os.environ.get(k) for k in ("OMP_PROC_BIND", "OMP_PLACES", "OMP_NUM_THREADS")}},
        "cpu_baseline": {"value": rate, "unit": "cell-updates/s", "cores": host_threads(), "kind": "port",
                         "sample": f"{done} step(s) of the {n}^2 Sedov state after {W} warm-up step(s), median step time"},
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


def nccl_logging(rank):
    """NCCL's own record of the communicator (version, ranks, transports) goes to a per-rank FILE -- INFO on stdout would
    break the one-line contract, and silencing it hides what the run used.  Returns the file name pattern."""
    if "NCCL_DEBUG" in os.environ and os.environ["NCCL_DEBUG"].upper() not in ("", "VERSION", "WARN"):
        return os.environ.get("NCCL_DEBUG_FILE")
    logdir = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(logdir, exist_ok=True)
    except OSError:
        logdir = "/tmp"
    os.environ["NCCL_DEBUG"] = "INFO"
    os.environ["NCCL_DEBUG_SUBSYS"] = "INIT"
    os.environ["NCCL_DEBUG_FILE"] = os.path.join(logdir, "nccl_bench_rank%s.log" % rank)
    return os.environ["NCCL_DEBUG_FILE"]


def nccl_summary(path):
    """what NCCL's log says about the communicator: version, nranks, NVLS / P2P use (rank 0's file)"""
    out = {"log": os.path.relpath(path, ROOT) if path else None}
    try:
        with open(path) as f:
            text = f.read()
    except (OSError, TypeError):
        return out
    import re
    m = re.search(r"NCCL version ([0-9.+a-z]+)", text)
    if m:
        out["version"] = m.group(1)
    ranks = [int(x) for x in re.findall(r"nranks (\d+)", text)]
    if ranks:
        out["nranks"] = max(ranks)
    out["nvls"] = "NVLS" in text
    for ln in text.splitlines():
        if "nranks" in ln and "comm" in ln:
            sys.stderr.write(ln.strip() + "\n")        # the rank line, where a log scraper looks for it
            break
    return out


class LaunchCounter:
    """counts the kernels this repository launches through pyro2_b200.ops: fill_ghost = 2 (x faces, y faces), the sweep 1,
    the stand-alone CFL reduction 1"""
    PER_CALL = {"fill_ghost": 2, "compressible_sweep": 1, "cfl_wavemax": 1}

    def __init__(self, ops):
        self.ops, self.n, self._orig = ops, 0, {}

    def __enter__(self):
        for name, k in self.PER_CALL.items():
            orig = getattr(self.ops, name)
            self._orig[name] = orig

            def wrapped(*a, _o=orig, _k=k, **kw):
                self.n += _k
                return _o(*a, **kw)
            setattr(self.ops, name, wrapped)
        return self

    def __exit__(self, *exc):
        for name, orig in self._orig.items():
            setattr(self.ops, name, orig)


# --------------------------------------------------------------------------------------------
def main():
    args = parse()
    guard_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # OpenMP pinning is for the CPU legs only (reference arm; cpu_baseline at N = 1).  Under torchrun every rank would bind
    # its threads -- the main thread included -- to the SAME first cores: eight GPU ranks sharing one core made every
    # launch-bound leg 10x slower in an N = 8 run of this round.
    if args.impl == "reference" and world > 1 and os.environ.get("OMP_NUM_THREADS") == "1":
        # torchrun exports OMP_NUM_THREADS=1 to its workers when the variable is unset; the reference arm is ONE process
        # (rank 0) and is meant to use every host thread it can, exactly as at N = 1
        os.environ.pop("OMP_NUM_THREADS")
    if args.impl == "reference" or world == 1:
        pin_host_threads()
    host_threads()
    if args.impl == "reference":
        run_reference(args, rank)
        return

    nccl_file = nccl_logging(rank) if world > 1 else None
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; pyro2_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import __graft_entry__
    __graft_entry__.build()
    from pyro2_b200 import _lib, ops
    import pyro2_b200.mesh.patch as patch_mod
    import pyro2_b200.compressible.simulation as comp_mod
    from pyro2_b200.parallel import SlabDecomposition
    from pyro2_b200.pyro_sim import Pyro

    n = args.nx
    K, W = args.steps, max(args.warmup, 3)
    peaks, peak_kind = measured_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t)
        return ms

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    # ---- the compressible Sedov problem through the public API --------------------------------
    p = Pyro("compressible")
    inputs = {"mesh.nx": n * world, "mesh.ny": n, "mesh.xmax": float(world), "driver.max_steps": 10 ** 9,
              "driver.tmax": 1.e9}
    slab = SlabDecomposition(rank, world) if world > 1 else None
    p.initialize_problem("sedov", inputs_dict=inputs, **({"decomposition": slab} if slab else {}))
    sim = p.sim

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(W):
        p.single_step()
    sim.check_state()
    barrier()
    counters = [LaunchCounter(m) for m in (ops,)]
    # the solver modules call ops.<name> through the module object, so patching the module's attributes counts them
    with counters[0]:
        sampler.mark_start()
        e0.record()
        for _ in range(K):
            p.single_step()
        e1.record()
        barrier()
        sampler.mark_end()
    ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    sim.check_state()
    value = n * n * world * K / (ms * 1e-3)
    gpu_launches = counters[0].n
    del patch_mod, comp_mod

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
    traffic, traffic_note = None, "no ncu capture of this build on record (profiles/sweep_traffic.json)"
    try:
        with open(os.path.join(ROOT, "profiles", "sweep_traffic.json")) as f:
            rec = json.load(f)
        if rec.get("source_hash") == _lib.built_hash():
            traffic = rec.get("dram_bytes_per_launch")
            traffic_note = f"ncu dram__bytes_read+write of this build ({rec.get('captured', '?')})"
        else:
            traffic_note = "profiles/sweep_traffic.json was captured on another build (source hash differs): not reported"
    except OSError:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "traffic_note": traffic_note,
                "peak_source": f"of {peak_kind}", "kernel": "pyro::sweep_kernel", "kernel_ms": kms,
                "algorithmic_bytes_per_launch": alg_bytes, "library_source_hash": _lib.built_hash(),
                "fp64": {"dp_instructions_per_cell": 1010, "pipe_rate_measured_dfma_per_clk_sm": 59.0,
                         "floor_ms": 1010 * g.nx * g.ny / (59.0 * 148 * 1.965e9) * 1e3,
                         "frac_of_fp64_floor": (1010 * g.nx * g.ny / (59.0 * 148 * 1.965e9) * 1e3) / kms,
                         "note": "the fused sweep is FP64-pipe bound (~1.0 k DP instructions per cell update; measured issue "
                                 "rate 59 DFMA/clk/SM, profiles/r2_ubench_fp64_latency.txt): this is the roofline that binds"},
                "note": "algorithmic-byte rate against the HBM copy peak, as the contract asks; see fp64 for the binding one"}

    # ---- e2e: the state lives in pinned host memory; every step copies it in and out -------------------
    e2e = None
    if not args.skip_e2e:
        planes = sim.cc_data.planes
        nbytes = 4 * g.nx * planes.stride(1) * 8            # valid rows of the four planes, each way
        ke2 = min(K, 10)
        if world == 1:
            bufs = [torch.empty(planes.shape, dtype=planes.dtype, pin_memory=True) for _ in range(2)]
            bufs[0].copy_(planes)
            torch.cuda.synchronize()
            p.single_step_streamed(bufs[0], bufs[1])            # warm-up (reduction path: the buffer is new)
            p.single_step_streamed(bufs[1], bufs[0])
            barrier()
            e0.record()
            for k in range(ke2):
                p.single_step_streamed(bufs[k % 2], bufs[(k + 1) % 2])
            e1.record()
            barrier()
            ems = e0.elapsed_time(e1)
            note = ("Pyro.single_step_streamed(): state in pinned host memory, 16 row blocks host -> device -> host on "
                    "three streams, overlapping the sweep; the result of step k is the input of step k + 1")
            del bufs
        else:
            # decomposed runs: whole-slab copies around single_step() (the streamed step is single-GPU)
            host_in = torch.empty(planes.shape, dtype=planes.dtype, pin_memory=True)
            host_out = torch.empty(planes.shape, dtype=planes.dtype, pin_memory=True)
            host_in.copy_(planes)
            nbytes = planes.numel() * 8
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
            note = "Pyro.single_step() with the rank's slab copied from/to pinned host memory every step"
            del host_in, host_out
        ems = max_over_ranks(ems)
        sim.check_state()
        e2e = {"value": n * n * world * ke2 / (ems * 1e-3), "unit": "cell-updates/s", "ms_per_step": ems / ke2,
               "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes, "steps": ke2, "note": note}

    # ---- multigrid V-cycles: the SAME 4096^2 problem on all N GPUs (strong scaling) -----------------
    mg = None
    if not args.skip_mg:
        del p, sim, A, B
        torch.cuda.empty_cache()
        from pyro2_b200.multigrid import MG
        a = MG.CellCenterMG2d(n, n, decomposition=SlabDecomposition(rank, world) if world > 1 else None, split_n=1024)
        x = a.x2d.t()
        y = a.y2d.t()
        a.init_zeros()
        a.init_RHS(-2.0 * ((1.0 - 6.0 * x ** 2) * y ** 2 * (1.0 - y ** 2) + (1.0 - 6.0 * y ** 2) * x ** 2 * (1.0 - x ** 2)))
        a.max_cycles = 3
        a.solve(rtol=0.0)                       # warm-up cycles (also captures the CUDA graph)
        a.max_cycles = args.mg_cycles
        barrier()
        e0.record()
        a.solve(rtol=0.0)                       # exactly mg_cycles V-cycles through the public API
        e1.record()
        barrier()
        mms = max_over_ranks(e0.elapsed_time(e1)) / a.num_cycles
        cycles_run, resid = a.num_cycles, a.residual_error

        # where a cycle's time goes, from the record (each part event-timed on its own, 10 repeats, max over ranks)
        h = a._h
        fine = a.nlevels - 1
        split = h.info(fine)["split_level"] if world > 1 else 0

        def timed(fn, reps=10):
            fn()
            barrier()
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            barrier()
            return max_over_ranks(e0.elapsed_time(e1)) / reps
        old_phi = a._old_phi
        parts = {"cycle_ms": mms,
                 "vcycle_only_ms": timed(lambda: (h.zero_coarse(), h.vcycle())),
                 "diagnostics_ms": timed(lambda: h.cycle_diagnostics_enqueue(old_phi)),
                 "coarse_fused_le_64_ms": timed(lambda: h.vcycle_level(min(5, fine)))}
        if world > 1:
            parts["replicated_levels_ms"] = timed(lambda: h.vcycle_level(split - 1))
            parts["split_levels_incl_halo_waits_ms"] = parts["vcycle_only_ms"] - parts["replicated_levels_ms"]
            parts["first_split_level_n"] = 2 ** (split + 1)
            parts["note"] = ("split levels: x-slabs, halo rows pushed through peer memory from the kernels' epilogues (no "
                             "separate exchange launches: their cost is the flag waits inside the split-level kernels); "
                             "replicated levels: every rank runs the identical sub-cycle")
        h.set_stop(False)
        mg_bytes = 776.0 * n * n                # one-pass-per-operator model
        blocked_bytes = (4 * 42.0 + 24 + 10 + 18 + 40) * n * n * 4.0 / 3.0     # 4 blocked passes + residual, restrict, prolong, bookkeeping; x4/3 for the coarser levels
        mg = {"metric": "V-cycles/s", "value": 1e3 / mms, "unit": "V-cycles/s", "ms_per_cycle": mms,
              "cycles": cycles_run, "residual_error": resid, "scaling": "strong", "n_gpus": world,
              "config": {"workload": f"multigrid constant-coefficient Poisson {n}^2 fp64 (global), dirichlet, nsmooth 10/50",
                         "parallelism": (f"x-slabs x{world} on levels >= {2 ** (split + 1)}^2 (peer-memory halo rows, no NCCL in "
                                         "the cycle), coarser levels replicated; cycle replayed as a CUDA graph per rank"
                                         if world > 1 else "single GPU, cycle replayed as a CUDA graph"),
                         "solve": "CellCenterMG2d.solve(rtol=0) with the stopping rule on the device, 2 cycles enqueued per read-back"},
              "breakdown": parts,
              "roofline": {"bound": "hbm", "achieved": mg_bytes / (mms * 1e-3) / 1e9 / world, "peak": peaks["hbm_gbs"],
                           "unit": "GB/s per GPU", "frac": mg_bytes / (mms * 1e-3) / 1e9 / world / peaks["hbm_gbs"],
                           "traffic": None, "algorithmic_bytes_per_cycle": mg_bytes,
                           "blocked_model_bytes_per_cycle": blocked_bytes,
                           "frac_blocked_model": blocked_bytes / (mms * 1e-3) / 1e9 / world / peaks["hbm_gbs"],
                           "note": "776 B per finest cell per V-cycle is the one-pass-per-operator model (SURVEY 8d); the "
                                   "temporally blocked smoother needs ~42 B per cell per 5 iterations: blocked_model is what "
                                   "this implementation has to move, frac_blocked_model its fraction of the copy peak"}}
        del a, h, old_phi

    # ---- incompressible shear 2048^2 (BASELINE config 4): explicit stages + two multigrid projections ----
    incomp = None
    if not args.skip_incomp:
        torch.cuda.empty_cache()
        ni = args.incomp_nx
        pi = Pyro("incompressible")
        pi.initialize_problem("shear", inputs_dict={"mesh.nx": ni, "mesh.ny": ni, "driver.max_steps": 10 ** 9,
                                                    "driver.tmax": 1.e9},
                              **({"decomposition": SlabDecomposition(rank, world)} if world > 1 else {}))
        isim = pi.sim
        pi.single_step()
        barrier()
        ki = 3
        e0.record()
        for _ in range(ki):
            pi.single_step()
        e1.record()
        barrier()
        ims = max_over_ranks(e0.elapsed_time(e1) / ki)
        solver = next(iter(isim._mg.values()))[0]
        incomp = {"metric": "zone-updates/s", "value": ni * ni / (ims * 1e-3), "unit": "zone-updates/s", "ms_per_step": ims,
                  "steps": ki, "n_gpus": world, "scaling": "strong",
                  "config": {"workload": f"incompressible shear {ni}^2 fp64 (global), periodic, limiter 2, proj_type 2",
                             "parallelism": f"x-slabs x{world}: NCCL halo rows for the explicit stages, peer-memory multigrid" if world > 1 else "single GPU",
                             "note": "each step = p2b_flow_* explicit stages + 2 multigrid projections at rtol 1e-12; "
                                     "at this size the reference's own stopping rule runs both to max_cycles = 100"},
                  "v_cycles_last_solve": solver.num_cycles, "gpu_launches_explicit": 14}
        if rank == 0 and world == 1 and not args.skip_cpu:
            nc = min(ni, 1024)
            secs, cyc = cpu_incompressible(nc)
            incomp["cpu_baseline"] = {"value": nc * nc / secs, "unit": "zone-updates/s", "cores": host_threads(), "kind": "port",
                                      "sample": f"1 step of the {nc}^2 shear problem ({cyc[0]} + {cyc[1]} V-cycles, {secs:.1f} s)"}
        del pi, isim, solver

    # ---- BASELINE config 5: Sedov 16384^2 on 8 x-slabs of 2048 x 16384 ---------------------------------
    config5 = None
    if (world == 8 or (args.config5 and world > 1)) and not args.skip_config5:
        try:
            torch.cuda.empty_cache()
            N5 = 16384
            p5 = Pyro("compressible")
            p5.initialize_problem("sedov", inputs_dict={"mesh.nx": N5, "mesh.ny": N5, "driver.max_steps": 10 ** 9, "driver.tmax": 1.e9},
                                  decomposition=SlabDecomposition(rank, world))
            for _ in range(3):
                p5.single_step()
            p5.sim.check_state()
            k5 = 10
            barrier()
            e0.record()
            for _ in range(k5):
                p5.single_step()
            e1.record()
            barrier()
            m5 = max_over_ranks(e0.elapsed_time(e1))
            p5.sim.check_state()
            config5 = {"metric": "cell-updates/s", "value": N5 * N5 * k5 / (m5 * 1e-3), "unit": "cell-updates/s",
                       "ms_per_step": m5 / k5, "steps": k5, "n_gpus": world,
                       "config": {"workload": "compressible Sedov 16384^2 fp64 (global), HLLC, outflow",
                                  "parallelism": f"{world} x-slabs of {N5 // world} x 16384 zones, 4-row halo rows and the dt reduction "
                                                 "through peer memory each step"}}
            del p5
        except Exception as exc:   # pylint: disable=broad-except
            config5 = {"error": repr(exc)[:500]}

    # ---- N > 1: decomposed vs single-domain, bit for bit (the driver's GPU-test box has one GPU) --------------
    parity = None
    if world > 1 and not args.skip_parity:
        try:
            parity = multi_gpu_parity(rank, world, SlabDecomposition(rank, world), dist, torch)
        except Exception as exc:   # pylint: disable=broad-except
            parity = {"error": repr(exc)[:500]}

    # ---- CPU baseline (rank 0, N = 1) ------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        ncpu = min(n, 2048)
        rate, times, done = cpu_compressible(ncpu, 3, warmup=1)
        cpu = {"value": rate, "unit": "cell-updates/s", "cores": host_threads(), "kind": "port",
               "sample": f"{done} steps of the {ncpu}^2 Sedov state after 1 warm-up step (median of {['%.2f' % t for t in times]} s)"}
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
            "clocks": clocks, "e2e": e2e, "gpu_launches": gpu_launches,
            "roofline": roofline, "cpu_baseline": cpu, "mg": mg, "incompressible": incomp,
            "sedov_16384": config5, "parity": parity,
            "nccl": nccl_summary(nccl_file) if world > 1 else None,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def multi_gpu_parity(rank, world, slab, dist, torch):
    """decomposed run vs single-domain run in the same process group: Sedov 1024^2 for 30 steps (state and every dt) and a
    1024^2 multigrid solve (solution and cycle count); rank 0 runs the single-domain versions and compares bits"""
    from pyro2_b200.multigrid import MG
    from pyro2_b200.pyro_sim import Pyro
    out = {}
    N, steps = 1024, 30
    inputs = {"mesh.nx": N, "mesh.ny": N, "driver.max_steps": 10 ** 9, "driver.tmax": 1.e9}
    p = Pyro("compressible")
    p.initialize_problem("sedov", inputs_dict=inputs, decomposition=slab)
    dts = []
    for _ in range(steps):
        p.single_step()
        dts.append(p.sim.dt)
    p.sim.check_state()
    g = p.sim.cc_data.grid
    mine = p.sim.cc_data.planes[:, g.ilo:g.ihi + 1, g.jlo:g.jhi + 1].contiguous()
    parts = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, parts, dst=0)
    if rank == 0:
        s = Pyro("compressible")
        s.initialize_problem("sedov", inputs_dict=inputs)
        dts1 = []
        for _ in range(steps):
            s.single_step()
            dts1.append(s.sim.dt)
        g1 = s.sim.cc_data.grid
        one = s.sim.cc_data.planes[:, g1.ilo:g1.ihi + 1, g1.jlo:g1.jhi + 1]
        out["sedov_1024"] = {"steps": steps, "slabs": world, "bit_identical": bool(torch.equal(torch.cat(parts, dim=1), one)),
                             "dt_identical": dts == dts1}
        del s, one
    del p, mine, parts
    a = MG.CellCenterMG2d(N, N, decomposition=slab, split_n=256)
    x, y = a.x2d.t(), a.y2d.t()
    f = -2.0 * ((1.0 - 6.0 * x ** 2) * y ** 2 * (1.0 - y ** 2) + (1.0 - 6.0 * y ** 2) * x ** 2 * (1.0 - x ** 2))
    a.init_zeros()
    a.init_RHS(f)
    a.solve(rtol=1.e-11)
    gg = a.soln_grid
    mine = a.get_solution().t()[gg.ilo:gg.ihi + 1, gg.jlo:gg.jhi + 1].contiguous()
    parts = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, parts, dst=0)
    if rank == 0:
        b = MG.CellCenterMG2d(N, N)
        xb, yb = b.x2d.t(), b.y2d.t()
        b.init_zeros()
        b.init_RHS(-2.0 * ((1.0 - 6.0 * xb ** 2) * yb ** 2 * (1.0 - yb ** 2) + (1.0 - 6.0 * yb ** 2) * xb ** 2 * (1.0 - xb ** 2)))
        b.solve(rtol=1.e-11)
        one = b.get_solution().t()[1:-1, 1:-1]
        out["mg_1024"] = {"slabs": world, "split_levels_from": 256, "bit_identical": bool(torch.equal(torch.cat(parts, dim=0), one)),
                          "cycles": [a.num_cycles, b.num_cycles], "residual_error": [a.residual_error, b.residual_error]}
        del b
    dist.barrier()
    return out if rank == 0 else None


if __name__ == "__main__":
    main()

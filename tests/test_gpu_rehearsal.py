"""CPU: a fast subset of the `-m gpu` tests, their bodies run unchanged on the emulated device (tests/emu_device.py:
the product's Python layer over the host-compiled kernel libraries).  One case per GPU test function that finishes
in seconds here; `python tests/rehearse.py test_gpu_api test_gpu_flow test_gpu_hydro test_gpu_mg test_gpu_zz_problems`
rehearses every case (about 40 minutes on 8 cores; all of them passed at the end of round 1).

What this buys: a GPU test that breaks because of host-side logic (a stale expectation, parameter plumbing, a
boundary hook, a problem setup) or kernel indexing fails here first, where there is no GPU budget to lose."""
import pytest

import rehearse

# (module, test function, substring selecting the parametrised case or None)
SUBSET = [
    ("test_gpu_api", "test_pyro_compressible_run_matches_reference", "[rt16]"),            # gravity + hse boundaries
    ("test_gpu_api", "test_pyro_run_sim_and_accessors", None),
    ("test_gpu_api", "test_compressible_unit_assertions", None),
    ("test_gpu_api", "test_unsupported_configurations_fail_loudly", None),
    ("test_gpu_api", "test_mg_solve_matches_reference", "helmholtz_neumann_64"),
    ("test_gpu_api", "test_mg_inhomogeneous_dirichlet_matches_reference", None),
    ("test_gpu_api", "test_mg_variable_coeff_solve_matches_reference", "constant_32"),
    ("test_gpu_api", "test_mg_variable_coeff_argument_errors", None),
    ("test_gpu_api", "test_mg_gradient_known_answer", None),
    ("test_gpu_api", "test_mg_subclass_hooks_are_used", None),
    ("test_gpu_api", "test_mg_argument_errors", None),
    ("test_gpu_api", "test_cellcenterdata_fill_bc_matches_reference_fixture", None),
    ("test_gpu_api", "test_user_defined_bc_callback_runs_after_standard_fill", None),
    ("test_gpu_flow", "test_flow_stages_bit_exact", "[16-16-2]"),
    ("test_gpu_flow", "test_pyro_burgers_run_matches_reference", None),
    ("test_gpu_flow", "test_pyro_advection_run_matches_reference", "tophat32"),
    ("test_gpu_flow", "test_pyro_diffusion_run_matches_reference", "gaussian32_mixed"),
    ("test_gpu_flow", "test_incompressible_rejects_unsupported_boundaries", None),
    ("test_gpu_hydro", "test_fill_ghost_mixed_per_variable", None),
    ("test_gpu_hydro", "test_cfl_dt_bit_exact", "shock-100-37"),
    ("test_gpu_hydro", "test_sweep_one_step", "shock-20-20-0-0"),
    ("test_gpu_hydro", "test_sweep_invalid_state_flag", None),
    ("test_gpu_hydro", "test_sweep_rejects_bad_arguments", None),
    ("test_gpu_mg", "test_smooth_residual_bit_exact", "[16-"),
    ("test_gpu_mg", "test_restrict_prolong_bit_exact", "[4-"),
    ("test_gpu_mg", "test_vcycle_bit_exact", "[8-"),
    ("test_gpu_zz_problems", "test_pyro_burgers_problems_match_reference", "converge32"),
    ("test_gpu_zz_problems", "test_pyro_burgers_problems_match_reference", "tophat32"),
]


def _find(module, name, pick):
    for ident, fn, kw in rehearse.cases(module):
        if ident.split("[")[0] == name and (pick is None or pick in ident):
            return fn, kw
    raise LookupError(f"{module}::{name} {pick}: no such case (was the GPU test renamed?)")


@pytest.mark.parametrize("module,name,pick", SUBSET, ids=[f"{m}::{n}{p or ''}" for m, n, p in SUBSET])
def test_gpu_test_body_on_emulated_device(module, name, pick):
    fn, kw = _find(module, name, pick)
    calls = rehearse.run_case(fn, kw)
    assert isinstance(calls, dict)


def test_every_gpu_test_function_is_enumerated():
    """the enumeration sees the parametrisation of the GPU test modules (a rename or a new fixture argument that the
    rehearsal cannot supply shows up here)"""
    import inspect
    for module in ("test_gpu_api", "test_gpu_flow", "test_gpu_hydro", "test_gpu_mg", "test_gpu_zz_problems"):
        cs = rehearse.cases(module)
        assert cs
        for ident, fn, kw in cs:
            required = {n for n, prm in inspect.signature(fn).parameters.items() if prm.default is inspect.Parameter.empty}
            assert required == set(kw), ident            # (arguments with defaults are knobs of the rehearsal, not fixtures)


def test_smoke_on_emulated_device():
    """__graft_entry__.smoke() -- the first thing the driver runs on the GPU box -- with the product on the emulated
    device (CUDA availability patched for the duration): Pyro("compressible") and CellCenterMG2d against the oracle"""
    from unittest import mock

    import torch

    import __graft_entry__ as ge
    import emu_device
    with emu_device.emulated_device() as dev, mock.patch.object(torch.cuda, "is_available", lambda: True), \
            mock.patch.object(torch.cuda, "set_device", lambda d: None), mock.patch.object(ge, "build", lambda: None):
        ge.smoke()
    assert dev.calls["p2b_compressible_sweep"] == 3 and dev.calls["p2b_mg_vcycle"] >= 1


def test_bench_on_emulated_device(capfd, monkeypatch):
    """bench.py's main() at toy sizes with the product on the emulated device and the CUDA timing / memory calls it
    makes replaced by host stand-ins: every leg (resident value, sweep-only roofline timing, e2e with host buffers,
    multigrid, incompressible) runs and the one JSON line carries the keys of the contract.  The numbers mean nothing
    here; what is checked is that the host code of the bench still fits the product after a change."""
    import json
    import sys
    import time
    from unittest import mock

    import torch

    import __graft_entry__ as ge
    import bench
    import emu_device

    class Event:
        def __init__(self, enable_timing=False):
            self.t = 0.0

        def record(self, stream=None):
            self.t = time.perf_counter()

        def synchronize(self):
            pass

        def elapsed_time(self, other):
            return max((other.t - self.t) * 1e3, 1e-3)

    real_empty = torch.empty

    def empty(*a, **k):
        k.pop("pin_memory", None)
        return real_empty(*a, **k)

    real_tensor = torch.tensor

    def tensor(*a, **k):
        if str(k.get("device", "")).startswith("cuda"):
            k["device"] = "cpu"
        return real_tensor(*a, **k)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--nx", "32", "--steps", "2", "--warmup", "1", "--mg-cycles", "2",
                                      "--incomp-nx", "32", "--skip-cpu"])
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        monkeypatch.delenv(k, raising=False)
    with emu_device.emulated_device() as dev, mock.patch.object(torch.cuda, "is_available", lambda: True), \
            mock.patch.object(torch.cuda, "set_device", lambda d: None), mock.patch.object(ge, "build", lambda: None), \
            mock.patch.object(torch.cuda, "Event", Event), mock.patch.object(torch.cuda, "empty_cache", lambda: None), \
            mock.patch.object(torch, "empty", empty), mock.patch.object(torch, "tensor", tensor):
        bench.main()
    out = [line for line in capfd.readouterr().out.splitlines() if line.startswith("{")]
    assert len(out) == 1
    line = json.loads(out[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "mg", "incompressible"):
        assert key in line, key
    assert line["metric"] == "cell-updates/s" and line["n_gpus"] == 1 and line["steps"] == 2
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["roofline"]["bound"] == "hbm"
    assert dev.calls["p2b_compressible_sweep"] >= 2 + 3 and dev.calls["p2b_mg_vcycle"] >= 2



def test_streamed_step_on_emulated_device(monkeypatch):
    """Pyro.single_step_streamed() (host-resident state, row blocks host -> device -> host) against resident steps, bit for
    bit, on the emulated device: the block decomposition, per-block ghost fill and the reuse of the fused wave-speed maxima
    are host logic that runs here exactly as on the GPU (streams and events are no-ops: emulated kernels are synchronous)"""
    import torch

    import emu_device
    from test_gpu_api import test_streamed_steps_are_bit_identical_to_resident_steps as body
    real_empty = torch.empty

    def empty(*a, **k):
        k.pop("pin_memory", None)
        return real_empty(*a, **k)
    monkeypatch.setattr(torch, "empty", empty)
    with emu_device.emulated_device():
        body(64, 48, 4, "sedov", nsteps=5)
        body(48, 40, 1, "quad", nsteps=3)


def test_scale_tests_rehearsed_small():
    """the bodies of tests/test_gpu_scale.py at toy sizes on the emulated device (driver dt limits, per-cycle comparison,
    the solve with the device-side stopping rule and the emulated graph replay)"""
    import emu_device
    import test_gpu_scale as ts
    with emu_device.emulated_device():
        ts.test_sedov_through_the_driver_at_scale(64, (1, 6))
        ts.test_multigrid_2048_cycle_by_cycle(128)


def test_compressible_slabs_as_threads_on_the_emulated_device():
    """the hardware test of the same name minus the hardware: two compressible slabs as host threads of one process,
    halo rows and the wave-speed reduction through the emulated peer memory"""
    import numpy as np

    import emu_device
    from test_gpu_api import compressible_slabs_in_one_process
    with emu_device.emulated_device():
        full, one = compressible_slabs_in_one_process(2, "sedov", 32, 16, 4)
    assert np.array_equal(full, one)


def test_stored_reference_goldens_rehearsed():
    """tests/test_gpu_zzz_reference_h5.py on the emulated device: the product against the regression files the reference
    itself stores (Sod 128 x 10 after 76 steps; the 256^2 Dirichlet multigrid solve, bit for bit)"""
    import emu_device
    import test_gpu_zzz_reference_h5 as t
    import os
    if not os.environ.get("P2B_FULL_TESTS"):          # ~100 s of emulated Sod steps + ~100 s of emulated 256^2 V-cycles
        pytest.skip("set P2B_FULL_TESTS=1 (last full run: passed, see profiles/README.md)")
    with emu_device.emulated_device():
        t.test_pyro_sod_matches_the_stored_reference_golden()
        t.test_multigrid_matches_the_stored_reference_golden()



@pytest.mark.parametrize("case,solver,problem", [("advection", "advection", "smooth"), ("burgers", "burgers", "test"),
                                                 ("diffusion", "diffusion", "gaussian")])
def test_stored_flow_goldens_rehearsed(case, solver, problem):
    """tests/test_gpu_zzz_reference_h5.py's flow runs on the emulated device: this build's problem setups, time-step
    control and kernels reproduce the reference's stored advection / Burgers / diffusion files bit for bit"""
    import os
    import emu_device
    import test_gpu_zzz_reference_h5 as t
    if case == "diffusion" and not os.environ.get("P2B_FULL_TESTS"):          # 164 emulated 128^2 multigrid solves: ~9 min
        pytest.skip("set P2B_FULL_TESTS=1 (last full run: passed, see profiles/README.md)")
    with emu_device.emulated_device():
        t.test_pyro_flow_run_matches_the_stored_reference_golden(case, solver, problem)


def test_stored_incompressible_golden_rehearsed():
    """the 216-step shear run against pyro/incompressible/tests/shear_128_0216.h5 on the emulated device (~25 min)"""
    import os
    import emu_device
    import test_gpu_zzz_reference_h5 as t
    if not os.environ.get("P2B_FULL_TESTS"):
        pytest.skip("set P2B_FULL_TESTS=1 (last full run: passed, see profiles/README.md)")
    with emu_device.emulated_device():
        t.test_pyro_incompressible_run_matches_the_stored_reference_golden()


def test_streamed_step_block_parameters(monkeypatch):
    """regression (scripts/fuzz_streamed_emulated.py): with a reflecting -x boundary the CGF solver zeroes the normal
    velocity on the domain's -x face; the streamed step applied that on the low face of every row block"""
    import torch

    import emu_device
    real_empty = torch.empty
    monkeypatch.setattr(torch, "empty", lambda *a, **k: real_empty(*a, **{kk: v for kk, v in k.items() if kk != "pin_memory"}))
    inputs = {"mesh.nx": 62, "mesh.ny": 25, "driver.max_steps": 10 ** 6, "driver.tmax": 1e9, "driver.verbose": 0,
              "mesh.xlboundary": "reflect", "mesh.xrboundary": "outflow", "mesh.ylboundary": "outflow", "mesh.yrboundary": "outflow",
              "compressible.riemann": "CGF"}
    with emu_device.emulated_device():
        from pyro2_b200.pyro_sim import Pyro

        def make():
            p = Pyro("compressible")
            p.initialize_problem("kh", inputs_dict=inputs)
            return p
        ref, p = make(), make()
        planes = p.sim.cc_data.planes
        bufs = [torch.empty(planes.shape, dtype=planes.dtype, pin_memory=True) for _ in range(2)]
        bufs[0].copy_(planes)
        for step in range(2):
            ref.single_step()
            p.single_step_streamed(bufs[step % 2], bufs[(step + 1) % 2], nchunks=3)
            assert p.sim.dt == ref.sim.dt
        g = p.sim.cc_data.grid
        assert torch.equal(bufs[0][:, g.ilo:g.ihi + 1, g.jlo:g.jhi + 1], ref.sim.cc_data.planes[:, g.ilo:g.ihi + 1, g.jlo:g.jhi + 1])
        # second regression of the same fuzzer: every block read the FIRST rows of the heating profile plane
        inputs = {"mesh.nx": 48, "mesh.ny": 32, "driver.max_steps": 10 ** 6, "driver.tmax": 1e9, "driver.verbose": 0}

        def make_heating():
            q = Pyro("compressible")
            q.initialize_problem("heating", inputs_dict=inputs)
            return q
        ref, p = make_heating(), make_heating()
        planes = p.sim.cc_data.planes
        bufs = [torch.empty(planes.shape, dtype=planes.dtype, pin_memory=True) for _ in range(2)]
        bufs[0].copy_(planes)
        for step in range(2):
            ref.single_step()
            p.single_step_streamed(bufs[step % 2], bufs[(step + 1) % 2], nchunks=3)
            assert p.sim.dt == ref.sim.dt
        g = p.sim.cc_data.grid
        assert torch.equal(bufs[0][:, g.ilo:g.ihi + 1, g.jlo:g.jhi + 1], ref.sim.cc_data.planes[:, g.ilo:g.ihi + 1, g.jlo:g.jhi + 1])


def test_stored_rt_golden_rehearsed():
    """the 945-step Rayleigh-Taylor run (gravity, hse boundaries) against pyro/compressible/tests/rt_0945.h5 on the emulated
    device (~70 min; the 256^2 quad run would take ~4 h and is not rehearsed)"""
    import os
    import emu_device
    import test_gpu_zzz_reference_h5 as t
    if not os.environ.get("P2B_LONG_TESTS"):
        pytest.skip("set P2B_LONG_TESTS=1 (last run: passed, see profiles/README.md)")
    with emu_device.emulated_device():
        t.test_pyro_compressible_run_matches_the_stored_reference_golden("rt", "rt")

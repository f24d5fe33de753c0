"""Minimal step driver for the solvers of this package.

The reference's driver (pyro/pyro_sim.py) is out of scope (reused as-is by a pyro maintainer who
swaps the solver modules); what the tests, bench.py and smoke() need from it is only the calling order of one step
(pyro/pyro_sim.py:241-256: fill_BC_all -> compute_timestep -> evolve), the layering of runtime parameters and the
lookup of a problem module.  That is all this file provides, under the reference's names so call sites read the same:

    p = Pyro("compressible")
    p.initialize_problem("sedov", inputs_dict={"mesh.nx": 4096, "mesh.ny": 4096})
    p.run_sim()            # or p.single_step() in a loop
"""
import importlib
import os

from . import defaults
from .util import msg
from .util import profile_pyro as profile
from .util.runparams import RuntimeParameters

valid_solvers = ["advection", "burgers", "compressible", "diffusion", "incompressible", "lm_atm"]

# what a driver used as a library switches off (plots, chatter, snapshots) unless inputs_dict says otherwise
_QUIET = {"vis.dovis": 0, "driver.verbose": 0, "io.do_io": 0}


class _Problem:
    """initial conditions + their parameters: a problems/<name>.py module of a solver, or a user function"""

    def __init__(self, name, init, params, finalize=None, sources=None, stock_inputs=None):
        self.name, self.init, self.params = name, init, params
        self.finalize, self.sources, self.stock_inputs = finalize, sources, stock_inputs or {}

    @classmethod
    def from_module(cls, solver_name, name):
        mod = importlib.import_module(f"{__package__}.{solver_name}.problems.{name}")
        return cls(name, mod.init_data, mod.PROBLEM_PARAMS, mod.finalize, getattr(mod, "source_terms", None),
                   getattr(mod, "INPUTS", {}))


class Pyro:
    def __init__(self, solver_name):
        if solver_name not in valid_solvers:
            msg.fail(f"ERROR: {solver_name} is not a valid solver (this build provides {valid_solvers})")
        self.solver_name = solver_name
        self.solver = importlib.import_module(f"{__package__}.{solver_name}")
        self.rp = RuntimeParameters()
        for layer in (defaults.GLOBAL, defaults.SOLVER[solver_name]):
            self.rp.load_dict(layer)
        self.tc = profile.TimerCollection()
        self.custom_problems = {}
        self.sim = None

    is_initialized = property(lambda self: self.sim is not None)

    def add_problem(self, name, problem_func, *, problem_params=None):
        self.custom_problems[name] = _Problem(name, problem_func, problem_params or {})

    def initialize_problem(self, problem_name, *, inputs_file=None, inputs_dict=None, decomposition=None):
        """parameter layers, weakest first: package defaults, the problem's own parameters, its stock inputs (or an
        inputs file), the quiet-library switches, inputs_dict.  decomposition (extension): a
        parallel.SlabDecomposition -- this process then owns one x-slab of the mesh.nx x mesh.ny domain."""
        prob = self.custom_problems.get(problem_name) or _Problem.from_module(self.solver_name, problem_name)
        for key, value in prob.params.items():
            self.rp.set_param(key, value, no_new=False)
        if inputs_file is None:
            self.rp.load_dict(prob.stock_inputs, no_new=True)
        elif os.path.isfile(inputs_file):
            self.rp.load_params(inputs_file, no_new=1)
        else:
            msg.fail(f"ERROR: inputs file {inputs_file} does not exist")
        for key, value in {**_QUIET, **(inputs_dict or {})}.items():
            self.rp.set_param(key, value)
        self.problem_name = prob.name
        self.verbose = self.rp.get_param("driver.verbose")

        sim = self.solver.Simulation(self.solver_name, prob.name, prob.init, self.rp, problem_finalize_func=prob.finalize,
                                     problem_source_func=prob.sources, timers=self.tc)
        sim.decomposition = decomposition
        sim.initialize()
        sim.preevolve()
        sim.cc_data.t = 0.0
        self.sim = sim

    def _require_sim(self):
        if self.sim is None:
            msg.fail("ERROR: problem has not been initialized")
        return self.sim

    def single_step(self):
        sim = self._require_sim()
        sim.cc_data.fill_BC_all()
        sim.compute_timestep()
        sim.evolve()
        if self.verbose > 0:
            print(f"{sim.n:5d} {sim.cc_data.t:10.5f} {sim.dt:10.5f}")
        if sim.do_output():
            self._snapshot()

    def single_step_streamed(self, host_in, host_out, nchunks=16):
        """single_step() for a state that lives in pinned host memory (compressible solver): host_in -> device -> step
        -> host_out with the copies pipelined against the sweep (Simulation.step_streamed)"""
        sim = self._require_sim()
        sim.step_streamed(host_in, host_out, nchunks=nchunks)
        if self.verbose > 0:
            print(f"{sim.n:5d} {sim.cc_data.t:10.5f} {sim.dt:10.5f}")

    def _snapshot(self):
        self.sim.write(f"{self.rp.get_param('io.basename')}{self.sim.n:04d}")

    def run_sim(self):
        sim = self._require_sim()
        with_io = self.rp.get_param("io.do_io")
        if with_io:
            self._snapshot()
        while not sim.finished():
            self.single_step()
        sim.check_state()          # the last step's device-side status word (the reference asserts inside evolve)
        if with_io or self.rp.get_param("io.force_final_output"):
            self._snapshot()
        sim.finalize()
        return sim

    def get_var(self, v):
        return self._require_sim().cc_data.get_var(v)

    def get_grid(self):
        return self._require_sim().cc_data.grid

    def get_sim(self):
        return self.sim

    def __repr__(self):
        return f"Pyro('{self.solver_name}')"

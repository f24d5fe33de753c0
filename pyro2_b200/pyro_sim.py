"""The driver: mirror of pyro/pyro_sim.py (Pyro :34-322) for the solvers this package provides.

    p = Pyro("compressible")
    p.initialize_problem("sedov", inputs_dict={"mesh.nx": 4096, "mesh.ny": 4096})
    p.run_sim()            # or p.single_step() in a loop

single_step() keeps the reference's order: fill_BC_all -> compute_timestep -> evolve
(pyro_sim.py:241-256)."""
import importlib
import os

from . import defaults
from .util import msg
from .util import profile_pyro as profile
from .util.runparams import RuntimeParameters

valid_solvers = ["advection", "burgers", "compressible", "diffusion", "incompressible", "lm_atm"]


class Pyro:
    def __init__(self, solver_name, *, from_commandline=False):
        if from_commandline:
            msg.bold("pyro (B200 hot-path build) ...")
        if solver_name not in valid_solvers:
            msg.fail(f"ERROR: {solver_name} is not a valid solver (this build provides {valid_solvers})")
        self.from_commandline = from_commandline
        self.pyro_home = os.path.dirname(os.path.realpath(__file__)) + "/"
        self.solver = importlib.import_module(f"{__package__}.{solver_name}")
        self.solver_name = solver_name
        self.problem_name = None
        self.problem_func = None
        self.problem_source = None
        self.problem_params = None
        self.problem_finalize = None
        self.custom_problems = {}
        self.rp = RuntimeParameters()
        self.rp.load_dict(defaults.GLOBAL)
        self.rp.load_dict(defaults.SOLVER[self.solver_name])
        self.tc = profile.TimerCollection()
        self.is_initialized = False

    def add_problem(self, name, problem_func, *, problem_params=None):
        """register a custom initial-condition function (pyro_sim.py:91-106)"""
        self.custom_problems[name] = (problem_func, problem_params or {})

    def initialize_problem(self, problem_name, *, inputs_file=None, inputs_dict=None, decomposition=None):
        """decomposition (extension): a parallel.SlabDecomposition; this process then owns one
        x-slab of the mesh.nx x mesh.ny domain"""
        if problem_name in self.custom_problems:
            self.problem_name = problem_name
            self.problem_func, self.problem_params = self.custom_problems[problem_name]
            self.problem_finalize = None
            self.problem_source = None
        else:
            problem = importlib.import_module(f"{__package__}.{self.solver_name}.problems.{problem_name}")
            self.problem_name = problem_name
            self.problem_func = problem.init_data
            self.problem_params = problem.PROBLEM_PARAMS
            self.problem_finalize = problem.finalize
            self.problem_source = getattr(problem, "source_terms", None)
            stock = getattr(problem, "INPUTS", {}) if inputs_file is None else None

        for k, v in self.problem_params.items():
            self.rp.set_param(k, v, no_new=False)

        if problem_name not in self.custom_problems and stock is not None:
            # the problem's stock parameter set (the reference ships these as inputs.* files), applied
            # like an inputs file: after the problem's own parameters, existing keys only
            self.rp.load_dict(stock, no_new=True)

        if inputs_file is not None:
            if not os.path.isfile(inputs_file):
                msg.fail(f"ERROR: inputs file {inputs_file} does not exist")
            self.rp.load_params(inputs_file, no_new=1)

        if not self.from_commandline:
            self.rp.set_param("vis.dovis", 0)
            self.rp.set_param("driver.verbose", 0)
            self.rp.set_param("io.do_io", 0)

        if inputs_dict is not None:
            for k, v in inputs_dict.items():
                self.rp.set_param(k, v)

        self.verbose = self.rp.get_param("driver.verbose")
        self.dovis = self.rp.get_param("vis.dovis")

        self.sim = self.solver.Simulation(self.solver_name, self.problem_name, self.problem_func, self.rp,
                                          problem_finalize_func=self.problem_finalize,
                                          problem_source_func=self.problem_source, timers=self.tc)
        self.sim.decomposition = decomposition
        self.sim.initialize()
        self.sim.preevolve()
        self.sim.cc_data.t = 0.0
        self.is_initialized = True

    def run_sim(self):
        if not self.is_initialized:
            msg.fail("ERROR: problem has not been initialized")
        tm_main = self.tc.timer("main")
        tm_main.begin()
        basename = self.rp.get_param("io.basename")
        do_io = self.rp.get_param("io.do_io")
        if do_io:
            self.sim.write(f"{basename}{self.sim.n:04d}")
        while not self.sim.finished():
            self.single_step()
        if do_io or self.rp.get_param("io.force_final_output"):
            self.sim.write(f"{basename}{self.sim.n:04d}")
        tm_main.end()
        if self.verbose > 0:
            self.rp.print_unused_params()
            self.tc.report()
        self.sim.finalize()

    def single_step(self):
        if not self.is_initialized:
            msg.fail("ERROR: problem has not been initialized")
        self.sim.cc_data.fill_BC_all()
        self.sim.compute_timestep()
        self.sim.evolve()
        if self.verbose > 0:
            print("%5d %10.5f %10.5f" % (self.sim.n, self.sim.cc_data.t, self.sim.dt))
        if self.sim.do_output():
            basename = self.rp.get_param("io.basename")
            self.sim.write(f"{basename}{self.sim.n:04d}")

    def __repr__(self):
        return f"Pyro('{self.solver_name}')"

    def __str__(self):
        s = f"Solver = {self.solver_name}\n"
        if self.is_initialized:
            s += f"Problem = {self.sim.problem_name}\n"
            s += f"Simulation time = {self.sim.cc_data.t}\n"
            s += f"Simulation step number = {self.sim.n}\n"
        return s + "\nRuntime Parameters\n------------------\n" + str(self.rp)

    def get_var(self, v):
        if not self.is_initialized:
            msg.fail("ERROR: problem has not been initialized")
        return self.sim.cc_data.get_var(v)

    def get_grid(self):
        if not self.is_initialized:
            msg.fail("ERROR: problem has not been initialized")
        return self.sim.cc_data.grid

    def get_sim(self):
        return self.sim


class PyroBenchmark(Pyro):
    """Pyro plus the regression bookkeeping of the reference (pyro_sim.py:324-408): compare the final state with a stored
    snapshot <pyro_home>/<solver>/tests/<basename><nnnn>.h5, or store one.  Snapshots need h5py (util/io_pyro.py)."""

    def __init__(self, solver_name, *, comp_bench=False, reset_bench_on_fail=False, make_bench=False):
        super().__init__(solver_name)
        self.comp_bench = comp_bench
        self.reset_bench_on_fail = reset_bench_on_fail
        self.make_bench = make_bench

    def _bench_file(self):
        basename = self.rp.get_param("io.basename")
        return f"{self.pyro_home}{self.solver_name}/tests/{basename}{self.sim.n:04d}"

    def run_sim(self, rtol=1.e-12):
        """evolve to the end, then compare with / store the benchmark; returns the comparison result when
        comparing (0 = match), else the simulation object"""
        super().run_sim()
        result = 0
        if self.comp_bench:
            result = self.compare_to_benchmark(rtol)
        if self.make_bench or (result != 0 and self.reset_bench_on_fail):
            self.store_as_benchmark()
        if self.comp_bench:
            return result
        return self.sim

    def compare_to_benchmark(self, rtol):
        from .util import compare   # pylint: disable=import-outside-toplevel
        from .util import io_pyro as io   # pylint: disable=import-outside-toplevel
        compare_file = self._bench_file()
        msg.warning(f"comparing to: {compare_file} ")
        try:
            sim_bench = io.read(compare_file, device=self.sim.cc_data.grid.device)
        except OSError:
            msg.warning("ERROR opening compare file")
            return "ERROR opening compare file"
        result = compare.compare(self.sim.cc_data, sim_bench.cc_data, rtol)
        if result == 0:
            msg.success(f"results match benchmark to within relative tolerance of {rtol}\n")
        else:
            msg.warning("ERROR: " + compare.errors[result] + "\n")
        return result

    def store_as_benchmark(self):
        tests = f"{self.pyro_home}{self.solver_name}/tests/"
        if not os.path.isdir(tests):
            try:
                os.mkdir(tests)
            except (FileNotFoundError, PermissionError):
                msg.fail("ERROR: unable to create the solver's tests/ directory")
        bench_file = self._bench_file()
        msg.warning(f"storing new benchmark: {bench_file}\n")
        self.sim.write(bench_file)

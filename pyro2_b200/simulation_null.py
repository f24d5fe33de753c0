"""Solver-independent simulation scaffolding: mirror of pyro/simulation_null.py
(grid_setup :10-68, bc_setup :71-112, NullSimulation :115-300)."""
import numpy as np

from .mesh import boundary as bnd
from .mesh import patch
from .util import msg
from .util import profile_pyro as profile


def _param(rp, key, default):
    try:
        return rp.get_param(key)
    except KeyError:
        msg.warning(f"{key} not set, defaulting to {default}")
        return default


def grid_setup(rp, ng=1, decomposition=None):
    """decomposition (extension): build only this rank's x-slab of the mesh.nx x mesh.ny grid"""
    nx = rp.get_param("mesh.nx")
    ny = rp.get_param("mesh.ny")
    xmin = _param(rp, "mesh.xmin", 0.0)
    xmax = _param(rp, "mesh.xmax", 1.0)
    ymin = _param(rp, "mesh.ymin", 0.0)
    ymax = _param(rp, "mesh.ymax", 1.0)
    grid_type = _param(rp, "mesh.grid_type", "Cartesian2d")
    if grid_type == "SphericalPolar":
        # x = r, y = theta (simulation_null.py:46-68)
        if decomposition is not None and decomposition.size > 1:
            if decomposition.local_nx(nx) < ng:
                raise ValueError(f"mesh.nx = {nx} on {decomposition.size} slabs leaves fewer than ng = {ng} rows per slab")
            my_grid = patch.SphericalPolar(decomposition.local_nx(nx), ny, xmin=xmin, xmax=xmax, ymin=ymin, ymax=ymax,
                                           ng=ng, nx_global=nx, ioffset=decomposition.ioffset(nx))
        else:
            my_grid = patch.SphericalPolar(nx, ny, xmin=xmin, xmax=xmax, ymin=ymin, ymax=ymax, ng=ng)
        # the polar axis is a reflecting boundary
        if ymin <= 0.05:
            rp.set_param("mesh.ylboundary", "reflect")
            msg.warning("With SphericalPolar grid, mesh.ylboundary auto set to reflect when ymin ~ 0")
        if abs(np.pi - ymax) <= 0.05:
            rp.set_param("mesh.yrboundary", "reflect")
            msg.warning("With SphericalPolar grid, mesh.yrboundary auto set to reflect when ymax ~ pi")
        return my_grid
    if grid_type != "Cartesian2d":
        raise ValueError("Unsupported grid type!")
    if decomposition is not None and decomposition.size > 1:
        if decomposition.local_nx(nx) < ng:
            # a halo of ng rows must come from the adjacent slab alone
            raise ValueError(f"mesh.nx = {nx} on {decomposition.size} slabs leaves fewer than ng = {ng} rows per slab")
        return patch.Cartesian2d(decomposition.local_nx(nx), ny, xmin=xmin, xmax=xmax, ymin=ymin, ymax=ymax,
                                 ng=ng, nx_global=nx, ioffset=decomposition.ioffset(nx))
    return patch.Cartesian2d(nx, ny, xmin=xmin, xmax=xmax, ymin=ymin, ymax=ymax, ng=ng)


def bc_setup(rp):
    """BC objects for scalars, x-velocity-like and y-velocity-like variables (odd reflection in
    the normal direction)"""
    types = {k: _param(rp, f"mesh.{k}boundary", "periodic") for k in ("xl", "xr", "yl", "yr")}
    kw = {"xlb": types["xl"], "xrb": types["xr"], "ylb": types["yl"], "yrb": types["yr"]}
    return bnd.BC(**kw), bnd.BC(**kw, odd_reflect_dir="x"), bnd.BC(**kw, odd_reflect_dir="y")


def _optional(rp, key, default=None):
    """a driver parameter that may be absent (or rp itself may be a stub without parameters)"""
    try:
        return rp.get_param(key)
    except (AttributeError, KeyError):
        return default


class NullSimulation:
    """time-step bookkeeping shared by all solvers (constructor contract: simulation_null.py:117-191)"""

    def __init__(self, solver_name, problem_name, problem_func, rp, *,
                 problem_finalize_func=None, problem_source_func=None,
                 timers=None, data_class=patch.CellCenterData2d):
        # identity of the run
        self.solver_name, self.problem_name = solver_name, problem_name
        self.problem_func = problem_func
        self.problem_finalize = problem_finalize_func
        self.problem_source = problem_source_func
        self.rp = rp
        self.data_class = data_class
        self.tc = timers if timers is not None else profile.TimerCollection()
        # step counters and limits
        self.n = 0
        self.dt = self.dt_old = -1.e33
        self.tmax = _optional(rp, "driver.tmax")
        self.max_steps = _optional(rp, "driver.max_steps")
        self.verbose = _optional(rp, "driver.verbose", 0)
        self.n_num_out = 0
        self.SMALL = 1.e-12
        # state filled in by initialize()
        self.cc_data = None
        self.particles = None
        self.cm = "viridis"
        self.decomposition = None   # set by Pyro.initialize_problem(decomposition=...)

    def __str__(self):
        return f"pyro Simulation:\n  solver: {self.solver_name}\n  problem: {self.problem_name}\n"

    def finished(self):
        return self.cc_data.t >= self.tmax or self.n >= self.max_steps

    def do_output(self):
        dt_out = self.rp.get_param("io.dt_out")
        n_out = self.rp.get_param("io.n_out")
        do_io = self.rp.get_param("io.do_io")
        is_time = self.cc_data.t >= (self.n_num_out + 1) * dt_out or self.n % n_out == 0
        if is_time and do_io == 1:
            self.n_num_out += 1
            return True
        return False

    def initialize(self):
        pass

    def method_compute_timestep(self):
        """the method-specific timestep code"""

    def compute_timestep(self):
        """driver-level limits on the method's dt (simulation_null.py:222-244)"""
        init_tstep_factor = self.rp.get_param("driver.init_tstep_factor")
        max_dt_change = self.rp.get_param("driver.max_dt_change")
        fix_dt = self.rp.get_param("driver.fix_dt")
        if fix_dt > 0.0:
            self.dt = fix_dt
            self.check_state()     # the method's timestep code, which reads the device status word, is skipped
        else:
            self.method_compute_timestep()
            if self.n == 0:
                self.dt = init_tstep_factor * self.dt
            else:
                self.dt = min(max_dt_change * self.dt_old, self.dt)
            self.dt_old = self.dt
        if self.cc_data.t + self.dt > self.tmax:
            self.dt = self.tmax - self.cc_data.t

    def check_state(self):
        """raise if the last evolve() left an invalid state (solvers whose kernels keep a device-side status word)"""

    def preevolve(self):
        pass

    def evolve(self):
        self.cc_data.t += self.dt
        self.n += 1

    def dovis(self):
        pass

    def finalize(self):
        if self.problem_finalize:
            self.problem_finalize()

    def write(self, filename):
        """HDF5 snapshot in the reference's layout (simulation_null.py:270-290); needs h5py"""
        import h5py   # pylint: disable=import-outside-toplevel
        decomp = getattr(self, "decomposition", None)
        if decomp is not None and decomp.size > 1:
            # every rank holds one x-slab: writing it under the domain's name and extent would race between ranks
            # and leave an inconsistent file.  No parallel writer exists; refuse instead of corrupting.
            msg.fail("ERROR: snapshots of a decomposed run are not supported (gather the slabs and write from one process)")
        self.check_state()
        if not filename.endswith(".h5"):
            filename += ".h5"
        with h5py.File(filename, "w") as f:
            f.attrs["solver"] = self.solver_name
            f.attrs["problem"] = self.problem_name
            f.attrs["time"] = self.cc_data.t
            f.attrs["nsteps"] = self.n
            self.cc_data.write_data(f)
            self.rp.write_params(f)
            self.write_extras(f)

    def write_extras(self, f):
        pass

    def read_extras(self, f):
        pass

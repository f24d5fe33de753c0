"""Reading back the HDF5 snapshots written by CellCenterData2d.write / Simulation.write -- the mirror of
pyro/util/io_pyro.py (read_bcs :13-24, read :27-148).  The file layout is the reference's, so files written by
either code are read by both; the data land in device memory (host -> device copy of the valid region).

Differences from the reference, all deliberate:
  * `device=` selects where the planes are allocated (default: the CUDA device, like every container here);
  * the step counter is restored from the file's "nsteps" attribute (the reference overwrites it with a
    variable name through a reused loop variable, io_pyro.py:79-82,124);
  * particle records are not part of this build: a file holding them is refused rather than read into
    something else."""
import importlib

import numpy as np
import torch

from ..mesh import boundary as bnd
from ..mesh.patch import Cartesian2d, CellCenterData2d, SphericalPolar

# solvers whose user boundary conditions live in another solver's BC module (io_pyro.py:69-72)
_BC_MODULE = {"compressible_fv4": "compressible", "compressible_rk": "compressible", "compressible_sdc": "compressible"}


def _scalar(v):
    """h5py hands back numpy scalars / bytes; the containers want Python values"""
    if isinstance(v, bytes):
        return v.decode()
    if isinstance(v, np.generic):
        return v.item()
    return v


def read_bcs(f):
    """the custom boundary-condition record {name: is_solid}, or None if the file has none"""
    try:
        gb = f["BC"]
    except KeyError:
        return None
    return {name: bool(np.asarray(gb[name])[()]) for name in gb}


def read(filename, device=None):
    """recreate the Simulation (or, for a bare patch file, the CellCenterData2d) stored in an HDF5 file"""
    import h5py   # pylint: disable=import-outside-toplevel
    if not filename.endswith(".h5"):
        filename += ".h5"
    with h5py.File(filename, "r") as f:
        try:
            solver_name = _scalar(f.attrs["solver"])
            problem_name = _scalar(f.attrs["problem"])
            t = float(f.attrs["time"])
            nsteps = int(f.attrs["nsteps"])
        except KeyError:
            solver_name = None          # a patch written on its own
        grid = f["grid"].attrs
        try:
            coord_type = int(grid["coord_type"])
        except KeyError:
            coord_type = 0
        grid_class = SphericalPolar if coord_type == 1 else Cartesian2d
        if "particles" in f:
            raise NotImplementedError("particle records are not part of the B200 build")
        myg = grid_class(int(grid["nx"]), int(grid["ny"]), ng=int(grid["ng"]),
                          xmin=float(grid["xmin"]), xmax=float(grid["xmax"]),
                          ymin=float(grid["ymin"]), ymax=float(grid["ymax"]), device=device)
        # custom boundary types must exist before variables carrying them are registered
        custom_bcs = read_bcs(f)
        if custom_bcs is not None:
            bcmod = importlib.import_module(f"pyro2_b200.{_BC_MODULE.get(solver_name, solver_name)}.BC")
            for name, is_solid in custom_bcs.items():
                bnd.define_bc(name, bcmod.user, is_solid=is_solid)
        gs = f["state"]
        names = list(gs)
        myd = CellCenterData2d(myg)
        for name in names:
            a = gs[name].attrs
            myd.register_var(name, bnd.BC(xlb=_scalar(a["xlb"]), xrb=_scalar(a["xrb"]),
                                          ylb=_scalar(a["ylb"]), yrb=_scalar(a["yrb"])))
        myd.create()
        for k in f["aux"].attrs:
            myd.set_aux(k, _scalar(f["aux"].attrs[k]))
        for name in names:
            host = np.ascontiguousarray(np.asarray(gs[name]["data"]), dtype=np.float64)
            myd.get_var(name).v()[:, :] = torch.from_numpy(host).to(myg.device)
        if solver_name is None:
            return myd
        solver = importlib.import_module(f"pyro2_b200.{solver_name}")
        sim = solver.Simulation(solver_name, problem_name, None, None)
        sim.n = nsteps
        sim.cc_data = myd
        sim.cc_data.t = t
        sim.particles = None
        sim.read_extras(f)
        # derived variables: the nearest derives module up the class hierarchy (io_pyro.py:131-141)
        for mod in [cls.__module__ for cls in type(sim).__mro__ if cls is not object]:
            try:
                derives = importlib.import_module(mod.replace("simulation", "derives"))
            except ModuleNotFoundError:
                continue
            if hasattr(derives, "derive_primitives"):
                sim.cc_data.add_derived(derives.derive_primitives)
                break
    return sim

"""Import shim for the UNMODIFIED reference (python-hydro/pyro2) from /root/reference.

TEST INFRASTRUCTURE ONLY.  Used by tests/golden/make_golden.py (fixture generation) and by
CPU tests that cross-check the C oracle when /root/reference is present (it is not on the
GPU box).  Nothing in the product path (pyro2_b200/) may import this file.

Recipe follows SURVEY.md section 8(c) "Recipe A": the reference imports h5py / matplotlib /
a setuptools_scm generated pyro._version at module import time
(pyro/mesh/patch.py:34, pyro/simulation_null.py:1, pyro/pyro_sim.py:8, pyro/multigrid/MG.py:68-69,
pyro/__init__.py:5); none of them is needed for the arithmetic, so they are stubbed.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("PYRO_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "pyro"))


class _Anything:
    """object that swallows every attribute access / call (matplotlib stand-in)"""

    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()

    def __iter__(self):
        return iter(())


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__getattr__ = lambda attr: _Anything()   # PEP 562 module getattr
    sys.modules[name] = m
    return m


def load():
    """make `import pyro` resolve to the read-only reference tree; returns the pyro package"""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/pyro_ref_numba_cache")
    os.environ.setdefault("NUMBA_NUM_THREADS", "1")
    if "pyro" in sys.modules and getattr(sys.modules["pyro"], "__file__", "").startswith(REF_ROOT):
        return sys.modules["pyro"]
    _stub("pyro._version", version="0.0.0+oracle")

    def _h5_fail(*a, **k):
        raise RuntimeError("h5py is stubbed in the oracle shim; run with io.do_io=0")
    _stub("h5py", File=_h5_fail)
    _stub("matplotlib")
    _stub("matplotlib.pyplot")
    _stub("matplotlib.ticker")
    _stub("mpl_toolkits")
    _stub("mpl_toolkits.axes_grid1")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import pyro  # noqa: E402
    return pyro


def make_sim(solver, problem, params, workdir="/tmp/pyro_ref_work"):
    """Pyro(solver).initialize_problem(problem, inputs_dict=params) run from a writable cwd
    (the reference writes inputs.auto, pyro/pyro_sim.py:170)."""
    load()
    from pyro.pyro_sim import Pyro
    os.makedirs(workdir, exist_ok=True)
    cwd = os.getcwd()
    os.chdir(workdir)
    try:
        p = Pyro(solver)
        p.initialize_problem(problem, inputs_dict=params)
    finally:
        os.chdir(cwd)
    return p

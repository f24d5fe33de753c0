"""Pin the oracle against EVERY regression file the reference stores for the paths built here (pyro/test.py:91-157 runs
exactly these configurations and compares with these files, util/compare.py, rtol 1e-12).  Runs in the build container
only (it needs /root/reference for the problem setups and the files):

    python tests/golden/pin_stored_goldens.py [case ...] > profiles/r2_oracle_vs_stored_goldens.txt

For each file: (1) the unmodified reference is run HERE on the runtime parameters recorded inside the file, which gives
the initial state and the dt of every step; (2) the oracle is stepped from that state with those dts (the compressible
oracle computes its own dts); (3) three differences are printed per variable, max-abs over the interior:
reference-here vs stored, oracle vs stored, oracle vs reference-here.  (4) A travelling fixture
tests/golden/refh5_<case>.npz is written (initial state, dts, parameters, the stored answer) for
tests/test_oracle_golden.py::test_oracle_reproduces_the_stored_*.

The stored files are read with tests/h5lite.py (h5py is not in this image).  Test infrastructure.
"""
import os
import shutil
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
sys.path.insert(0, TESTS)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(os.path.dirname(TESTS), "oracle"))

import golden_util  # noqa: E402
import h5lite  # noqa: E402
import make_golden as mk  # noqa: E402   (imports the reference through oracle/ref_shim.py)
import oracle  # noqa: E402
import oracle_runs as runs  # noqa: E402

REF = "/root/reference/pyro/"
mk.ONLY = []                # this script has its own case filter
CONS = ["density", "energy", "x-momentum", "y-momentum"]
LM_VARS = ["density", "x-velocity", "y-velocity", "eint", "phi-MAC", "phi", "gradp_x", "gradp_y"]

#        case          stored file                                  solver            problem     variables        stride
RUNS = [("quad",       "compressible/tests/quad_unsplit_0606.h5",   "compressible",   "quad",     CONS,            2),
        ("rt",         "compressible/tests/rt_0945.h5",             "compressible",   "rt",       CONS,            1),
        ("advection",  "advection/tests/smooth_0040.h5",            "advection",      "smooth",   ["density"],     1),
        ("burgers",    "burgers/tests/test_0051.h5",                "burgers",        "test",     ["x-velocity", "y-velocity"], 1),
        ("diffusion",  "diffusion/tests/gaussian_0164.h5",          "diffusion",      "gaussian", ["phi"],         1),
        ("incomp",     "incompressible/tests/shear_128_0216.h5",    "incompressible", "shear",    mk.INCOMP_VARS,  1),
        ("lm_atm",     "lm_atm/tests/lm_bubble_128_0065.h5",        "lm_atm",         "bubble",   LM_VARS,         0)]   # 0: no fixture
MGVC = [("mgvc_dirichlet", "multigrid/tests/mg_vc_poisson_dirichlet.h5", "dirichlet", "neumann"),
        ("mgvc_periodic",  "multigrid/tests/mg_vc_poisson_periodic.h5",  "periodic",  "periodic")]
RUNNERS = {"advection": runs.run_advection, "burgers": runs.run_burgers, "diffusion": runs.run_diffusion,
           "incompressible": runs.run_incompressible, "lm_atm": runs.run_lm_atm}


def stored_state(path, names):
    f = h5lite.File(REF + path)
    rp = {k: v for k, v in f.attrs("runtime parameters").items() if not k.startswith(("vis.", "io."))} \
        if "runtime parameters" in f.keys() else {}
    return [f[f"state/{n}/data"] for n in names], f.attrs(""), rp


def line(path, name, ora, here, sto):
    sc = max(float(np.abs(sto).max()), 1e-300)
    print(f"  {path:44s} {name:11s} max|stored| {sc:9.3e}   here-stored {np.abs(here - sto).max():9.3e}   "
          f"oracle-stored {np.abs(ora - sto).max():9.3e}   oracle-here {np.abs(ora - here).max():9.3e}", flush=True)


def pin_run(case, path, solver, problem, names, stride, tmp):
    sto, attrs, rp_in = stored_state(path, names)
    t0 = time.time()
    with open(os.devnull, "w") as null:                # the reference prints a line per step
        out, sys.stdout = sys.stdout, null
        try:
            if solver == "compressible":
                mk.comp_case(case, problem, rp_in, 10 ** 6)
            else:
                mk.flow_case(case + ".npz", solver, problem, rp_in, 10 ** 6, names)
        finally:
            sys.stdout = out
    t_ref = time.time() - t0
    t0 = time.time()
    if solver == "compressible":
        z, rp, _ = golden_util.load_comp(case)
        U, dts, ng = runs.run_compressible(z, rp)
        v = (slice(ng, -ng), slice(ng, -ng))
        ora = [U[v][..., k] for k in range(4)]
        here = [z["U"][v][..., k] for k in range(4)]
        first = z["U0"]
        note = f"oracle dts vs reference-here: max rel {np.abs(dts / z['dts'] - 1).max():.1e}"
    else:
        z, rp, _ = golden_util.load_flow(case + ".npz")
        res = RUNNERS[solver](z, rp)
        res = [res] if isinstance(res, np.ndarray) and res.ndim == 2 else list(res)
        ng = int(z["ng"])
        v = (slice(ng, -ng), slice(ng, -ng))
        ora, here, first = [a[v] for a in res], [p[v] for p in z["P"]], z["P0"]
        note = "oracle stepped with the reference's dts"
    assert attrs["nsteps"] == int(z["n"]) and abs(attrs["time"] - float(z["t"])) <= 1e-12 * float(z["t"]), (attrs, int(z["n"]))
    print(f"{case}: {solver} {problem} {sto[0].shape[0]} x {sto[0].shape[1]}, {attrs['nsteps']} steps to t = {attrs['time']}; "
          f"reference here {t_ref:.0f} s, oracle {time.time() - t0:.0f} s; {note}")
    for n, o, h, s in zip(names, ora, here, sto):
        line(path, n, o, h, s)
    if stride:
        keep = {k: z[k] for k in z.files if k not in ("U", "P", "names")}
        keep["U0" if solver == "compressible" else "P0"] = first
        np.savez_compressed(os.path.join(HERE, f"refh5_{case}.npz"), source=path, stride=stride, names=np.array(names),
                            stored=np.stack([s[::stride, ::stride] for s in sto]), **keep)


def pin_mgvc(case, path, phibc, cbc, stride=4):
    """pyro/multigrid/examples/mg_test_vc_{dirichlet,periodic}.py at 512^2 (pyro/test.py:143-151)"""
    f = h5lite.File(REF + path)
    sto = {n: f[f"state/{n}/data"] for n in f.keys("state")}
    n = f.attrs("grid")["nx"]
    with open(os.devnull, "w") as null:
        out, sys.stdout = sys.stdout, null
        try:
            mk.mgvc_case(f"{case}_{n}", n, phibc, cbc, "dirichlet" if phibc == "dirichlet" else "periodic")
        finally:
            sys.stdout = out
    z = golden_util.load_mgvc(f"{case}_{n}")
    o = oracle.MG(n, bc=(phibc,) * 4, alpha=0.0, beta=0.0)
    o.set_coeffs(z["coeffs"], (cbc,) * 4)
    o.init_zeros()
    o.init_RHS(z["f"])
    o.solve(rtol=1.e-11)
    i = (slice(1, -1), slice(1, -1))
    print(f"{case}: VarCoeffCCMG2d {n} x {n}, {o.num_cycles} V-cycles (reference here: {int(z['num_cycles'])})")
    line(path, "coeffs", z["coeffs"][i], z["coeffs"][i], sto["coeffs"])
    line(path, "f", z["f"][i], z["f"][i], sto["f"])
    line(path, "v", o.get_solution()[i], z["v"][i], sto["v"])
    line(path, "r", o.plane(o.nlevels - 1, "r")[i], z["r"][i], sto["r"])
    np.savez_compressed(os.path.join(HERE, f"refh5_{case}.npz"), source=path, stride=stride, nx=n,
                        bc=np.array((phibc,) * 4), coeffs_bc=np.array((cbc,) * 4), rtol=1.e-11, num_cycles=o.num_cycles,
                        stored_v=sto["v"][::stride, ::stride])


if __name__ == "__main__":
    only = sys.argv[1:]
    tmp = tempfile.mkdtemp(prefix="pin_stored_")
    mk.HERE = golden_util.GOLDEN = tmp               # the intermediate full fixtures stay out of the repository
    print(__doc__.split("\n\n")[0].replace("\n", " "))
    print("max-abs over the interior; 'here' = the unmodified reference run in this container, 'stored' = the file\n")
    try:
        for case, path, solver, problem, names, stride in RUNS:
            if not only or case in only:
                pin_run(case, path, solver, problem, names, stride, tmp)
        for case, path, phibc, cbc in MGVC:
            if not only or case in only:
                pin_mgvc(case, path, phibc, cbc)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)

"""Fixtures from the regression files the REFERENCE ITSELF STORES (not from running it): pyro's `.h5` goldens read with
tests/h5lite.py (h5py is not in this image) and re-packed as small .npz files, so that the tests can compare against the
reference's own stored answers on a box without /root/reference.

    python tests/golden/make_h5_golden.py      # writes tests/golden/refh5_*.npz

  pyro/compressible/tests/sod_x_0076.h5        Sod along x, 128 x 10, limiter 1, HLLC, 76 steps to t = 0.2
  pyro/multigrid/tests/mg_poisson_dirichlet.h5 256^2 Dirichlet Poisson solve of examples/multigrid/mg_test_simple.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import h5lite  # noqa: E402

REF = "/root/reference/pyro"


def pack(path, out, digest_only=()):
    """digest_only: variables stored as the sha256 of their bytes instead of the array (a right-hand side the test
    recomputes from its formula and checks bit for bit against the digest)"""
    import hashlib
    f = h5lite.File(os.path.join(REF, path))
    arrays = {name.replace("-", "_"): f[f"state/{name}/data"] for name in f.keys("state")}
    for name in digest_only:
        arrays[name + "_sha256"] = np.array(hashlib.sha256(np.ascontiguousarray(arrays.pop(name)).tobytes()).hexdigest())
    meta = dict(f.attrs(""), **{"grid." + k: v for k, v in f.attrs("grid").items()})
    bcs = {name: [f.attrs(f"state/{name}")[s] for s in ("xlb", "xrb", "ylb", "yrb")] for name in f.keys("state")}
    rp = f.attrs("runtime parameters") if "runtime parameters" in f.keys("") else {}
    np.savez_compressed(os.path.join(HERE, out), source=path,
                        meta=np.array([f"{k}={v}" for k, v in sorted(meta.items())]),
                        rp=np.array([f"{k}={v}" for k, v in sorted(rp.items())]),
                        bcs=np.array([f"{k}={','.join(v)}" for k, v in sorted(bcs.items())]), **arrays)
    print(out, {k: getattr(a, "shape", a) for k, a in arrays.items()}, meta)


if __name__ == "__main__":
    pack("compressible/tests/sod_x_0076.h5", "refh5_sod_x_0076.npz")
    pack("multigrid/tests/mg_poisson_dirichlet.h5", "refh5_mg_poisson_dirichlet.npz", digest_only=("f",))

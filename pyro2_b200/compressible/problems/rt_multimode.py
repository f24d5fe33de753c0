"""Rayleigh-Taylor instability seeded with a sum of cosine modes of random phase and amplitude (fixed seed, so
the run is reproducible); same parameters and random sequence as pyro/compressible/problems/rt_multimode.py.
Run with the hse boundaries in y."""
import numpy as np

from ...util import msg
from .rt2 import store, stratified

DEFAULT_INPUTS = "inputs.rt_multimode"

# stock run (the reference's inputs.rt_multimode)
INPUTS = {"driver.max_steps": 10000, "driver.tmax": 3.0, "io.basename": "rt_", "io.n_out": 100,
          "mesh.nx": 64, "mesh.ny": 192, "mesh.xmax": 1.0, "mesh.ymax": 3.0,
          "mesh.xlboundary": "periodic", "mesh.xrboundary": "periodic", "mesh.ylboundary": "hse", "mesh.yrboundary": "hse",
          "rt_multimode.amp": 0.25, "rt_multimode.nmodes": 12, "compressible.grav": -1.0, "compressible.limiter": 2}

PROBLEM_PARAMS = {"rt_multimode.dens1": 1.0, "rt_multimode.dens2": 2.0, "rt_multimode.amp": 1.0,
                  "rt_multimode.sigma": 0.1, "rt_multimode.nmodes": 10, "rt_multimode.p0": 10.0}

SEED = 12345


def init_data(my_data, rp):
    if rp.get_param("driver.verbose"):
        msg.bold("initializing the rt problem...")
    rng = np.random.default_rng(SEED)
    g = my_data.grid
    amp, sigma = rp.get_param("rt_multimode.amp"), rp.get_param("rt_multimode.sigma")
    nmodes = rp.get_param("rt_multimode.nmodes")
    dens, p, ycenter = stratified(g, rp.get_param("rt_multimode.dens1"), rp.get_param("rt_multimode.dens2"),
                                  rp.get_param("rt_multimode.p0"), rp.get_param("compressible.grav"))
    x = np.broadcast_to(g.x[:, None], (g.qx, g.qy))
    y = np.broadcast_to(g.y[None, :], (g.qx, g.qy))
    L = g.xmax - g.xmin
    envelope = np.exp(-(y - ycenter) ** 2 / sigma ** 2)
    ymom = np.zeros((g.qx, g.qy))
    for k in range(1, nmodes + 1):
        # one phase, then one amplitude per mode: the order fixes the random sequence
        phase = rng.random() * 2 * np.pi
        mode_amp = amp * rng.random()
        ymom += mode_amp * np.cos(2.0 * np.pi * k * x / L + phase) * envelope
    ymom /= nmodes
    store(my_data, dens, p, ymom * dens, rp.get_param("eos.gamma"))


def finalize():
    pass

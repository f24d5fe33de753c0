"""Shock-speed check for the Burgers `test` problem from two of its snapshots -- the counterpart of
pyro/burgers/problems/verify.py.  The speed |u| is averaged along the anti-diagonals (the shock front is one of
them), the front is located in both snapshots and its displacement over the elapsed time is compared with the
Rankine-Hugoniot speed of the 3 -> 1 jump, |(u_l + u_r)/2| * sqrt(2) = sqrt(8).

    python -m pyro2_b200.burgers.problems.verify file1 file2
"""
import sys

import numpy as np

from ...util import io_pyro as io

SHOCK_SPEED = np.sqrt(2.0 * 2.0 + 2.0 * 2.0)


def _front(d):
    """abscissa of the first anti-diagonal (beyond x = 0.5) whose mean speed has dropped below 90 % of sqrt(8)"""
    g = d.grid
    u = d.get_var("x-velocity").v().cpu().numpy()
    v = d.get_var("y-velocity").v().cpu().numpy()
    speed = np.flipud(np.sqrt(u * u + v * v))
    xc = g.x[g.ilo:g.ihi]
    nx = len(xc)
    means = [np.diagonal(speed, n).mean() for n in range(-(nx - 1), nx)]
    # abscissae of the diagonals: cell centres interleaved with running midpoints, as the reference builds them
    x = [xc[0]]
    for c in xc[1:]:
        x.append(0.5 * (x[-1] + c))
        x.append(c)
    return next(x[n] for n, m in enumerate(means) if m < 0.9 * SHOCK_SPEED and x[n] > 0.5)


def shock_speed(d1, d2):
    dx = _front(d1) - _front(d2)
    dt = d2.t - d1.t
    return np.sqrt(2.0 * (dx / dt) * (dx / dt))


def verify(file1, file2, device=None):
    s1, s2 = io.read(file1, device=device), io.read(file2, device=device)
    sim = shock_speed(s1.cc_data, s2.cc_data)
    print(f"Theoretical shock speed is: {SHOCK_SPEED}")
    print(f"Shock speed from simulation is: {sim}")
    ok = bool(np.isclose(sim, SHOCK_SPEED))
    print("SUCCESS, shock speeds match" if ok else "ERROR, shock speeds don't match")
    return ok


if __name__ == "__main__":
    verify(sys.argv[1], sys.argv[2])

"""Low-level wrapper of the p2b_mg_* C ABI: owns the torch workspace the hierarchy lives in and
exposes each level's v / f / r planes as torch views.  multigrid/MG.py builds the reference's
CellCenterMG2d interface on top of this."""
import ctypes as C

import torch

from . import _lib
from .ops import require_cuda


class MGHandle:
    def __init__(self, nx, bc, alpha, beta, xmin, xmax, ymin, ymax, nsmooth, nsmooth_bottom):
        require_cuda()
        L = _lib.lib()
        codes = (C.c_int * 4)(*[_lib.BC_CODES[b] for b in bc])
        self._h = L.p2b_mg_create(nx, codes, alpha, beta, xmin, xmax, ymin, ymax, nsmooth, nsmooth_bottom)
        if not self._h:
            raise ValueError(L.p2b_last_error().decode())
        self.nlevels = L.p2b_mg_nlevels(self._h)
        nbytes = L.p2b_mg_workspace_bytes(self._h)
        self.workspace = torch.zeros(nbytes // 8, dtype=torch.float64, device="cuda")
        _lib.check(L.p2b_mg_bind(self._h, self.workspace.data_ptr(), nbytes))
        self._bc_vals = [None] * 4
        self._out = torch.zeros(2, dtype=torch.float64, device="cuda")
        self._planes = {}

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().p2b_mg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def plane(self, level, which):
        """(n+2, n+2) strided view of plane which in {'v','f','r'} on `level`"""
        key = (level, which)
        if key not in self._planes:
            idx = {"v": 0, "f": 1, "r": 2}[which]
            n = 2 ** (level + 1)
            ptr = _lib.lib().p2b_mg_level_ptr(self._h, level, idx)
            pitch = _lib.lib().p2b_mg_level_pitch(self._h, level)
            off = (ptr - self.workspace.data_ptr()) // 8
            self._planes[key] = self.workspace.as_strided((n + 2, n + 2), (pitch, 1), off)
        return self._planes[key]

    def set_bc_values(self, xl=None, xr=None, yl=None, yr=None):
        vals = []
        for v in (xl, xr, yl, yr):
            vals.append(None if v is None else torch.as_tensor(v, dtype=torch.float64).contiguous().cuda())
        self._bc_vals = vals   # keep alive
        _lib.check(_lib.lib().p2b_mg_set_bc_values(self._h, *[None if v is None else v.data_ptr() for v in vals]))

    def set_blocking(self, enable):
        _lib.check(_lib.lib().p2b_mg_set_blocking(self._h, int(bool(enable))))

    def _s(self):
        return _lib.stream_ptr()

    def smooth(self, level, nsmooth):
        _lib.check(_lib.lib().p2b_mg_smooth(self._h, level, nsmooth, self._s()))

    def residual(self, level):
        _lib.check(_lib.lib().p2b_mg_residual(self._h, level, self._s()))

    def restrict(self, level):
        _lib.check(_lib.lib().p2b_mg_restrict(self._h, level, self._s()))

    def prolong_correct(self, level):
        _lib.check(_lib.lib().p2b_mg_prolong_correct(self._h, level, self._s()))

    def fill_bc(self, level):
        _lib.check(_lib.lib().p2b_mg_fill_bc(self._h, level, self._s()))

    def zero_coarse(self):
        _lib.check(_lib.lib().p2b_mg_zero_coarse(self._h, self._s()))

    def vcycle(self):
        _lib.check(_lib.lib().p2b_mg_vcycle(self._h, self._s()))

    def sumsq(self, level, which):
        idx = {"v": 0, "f": 1, "r": 2}[which]
        _lib.check(_lib.lib().p2b_mg_norm2(self._h, level, idx, self._out.data_ptr(), self._s()))
        return float(self._out[0])

    def cycle_diagnostics(self, old_phi):
        """returns (sum rel-change^2, sum r^2); updates old_phi <- v and the r plane"""
        _lib.check(_lib.lib().p2b_mg_cycle_diagnostics(self._h, old_phi.data_ptr(), self._out.data_ptr(), self._s()))
        a, b = self._out.tolist()
        return a, b

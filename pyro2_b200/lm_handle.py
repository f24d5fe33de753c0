"""Low-level wrapper of the p2b_lm_* C ABI (csrc/lm.cu): the explicit stages of the low Mach number atmosphere
solver.  Owns the scratch planes (among them the reference's aux_data "coeff" and "source_y") and the device
copy of the 1-d base state; lm_atm/simulation.py passes the state planes in and issues the ghost fills and the
variable-coefficient multigrid projections between the calls."""
import ctypes as C

import torch

from . import _lib, ops


class LmHandle:
    U_MAC, V_MAC, COEFF, SOURCE, RHO_OLD = 14, 15, 16, 17, 24

    def __init__(self, planes, grid, basestate):
        """planes: the solver's (nvar, qx, pitch) state storage; basestate: (4, qy) CUDA float64 tensor holding
        rho0, p0, beta0, beta0-edges"""
        ops.require_cuda()
        L = _lib.lib()
        self.grid = grid
        self.pitch = planes.stride(1)
        assert basestate.is_cuda and basestate.dtype == torch.float64 and basestate.is_contiguous()
        assert tuple(basestate.shape) == (4, grid.qy)
        self.basestate = basestate          # kept alive: the handle stores its address
        self._g = ops.grid_struct(planes, grid.nx, grid.ny, grid.ng, grid.dx, grid.dy)
        self._h = L.p2b_lm_create(C.byref(self._g), basestate.data_ptr())
        if not self._h:
            raise ValueError(L.p2b_last_error().decode())
        nbytes = L.p2b_lm_workspace_bytes(self._h)
        self.workspace = torch.zeros(nbytes // 8, dtype=torch.float64, device=planes.device)
        _lib.check(L.p2b_lm_bind(self._h, self.workspace.data_ptr(), nbytes))
        self._scratch = torch.zeros(5, dtype=torch.int64, device=planes.device)

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().p2b_lm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _s(self):
        return _lib.stream_ptr()

    def _p(self, t):
        if t is None:
            return None
        assert t.is_cuda and t.dtype == torch.float64 and t.stride() == (self.pitch, 1), "state plane layout"
        return t.data_ptr()

    def plane(self, n):
        """(qx, pitch) view of scratch plane n (the layout ops.fill_ghost expects after unsqueeze(0))"""
        g = self.grid
        ptr = _lib.lib().p2b_lm_plane(self._h, n)
        off = (ptr - self.workspace.data_ptr()) // 8
        return self.workspace.as_strided((g.qx, self.pitch), (self.pitch, 1), off)

    def fill(self, n, bc_names):
        """ghost fill of scratch plane n with the given boundary types (aux_data.fill_BC in the reference)"""
        g = self.grid
        ops.fill_ghost(self.plane(n).unsqueeze(0), g.nx, g.ny, g.ng, [bc_names])

    def coeff(self, d1, d2, numer, squared, buf):
        _lib.check(_lib.lib().p2b_lm_coeff(self._h, self._p(d1), self._p(d2), numer, int(squared), buf, self._s()))

    def source(self, rho, rho_old, grav):
        _lib.check(_lib.lib().p2b_lm_source(self._h, self._p(rho), self._p(rho_old), grav, self._s()))

    def interface_states(self, u, v, gradp_x, gradp_y, dt, limiter):
        _lib.check(_lib.lib().p2b_lm_interface_states(self._h, self._p(u), self._p(v), self._p(gradp_x), self._p(gradp_y),
                                                      dt, limiter, self._s()))

    def mac_vels(self):
        _lib.check(_lib.lib().p2b_lm_mac_vels(self._h, self._s()))

    def mac_divergence(self, div):
        assert div.stride(1) == 1
        _lib.check(_lib.lib().p2b_lm_mac_divergence(self._h, div.data_ptr(), div.stride(0), self._s()))

    def mac_project(self, phi_mac):
        _lib.check(_lib.lib().p2b_lm_mac_project(self._h, self._p(phi_mac), self._s()))

    def density_update(self, rho, eint, dt, limiter, gamma):
        _lib.check(_lib.lib().p2b_lm_density_update(self._h, self._p(rho), self._p(eint), dt, limiter, gamma, self._s()))

    def upwind_states(self):
        _lib.check(_lib.lib().p2b_lm_upwind_states(self._h, self._s()))

    def advect_update(self, u, v, gradp_x, gradp_y, dt, proj_type):
        _lib.check(_lib.lib().p2b_lm_advect_update(self._h, self._p(u), self._p(v), self._p(gradp_x), self._p(gradp_y),
                                                   dt, proj_type, self._s()))

    def add_source(self, v, dt):
        _lib.check(_lib.lib().p2b_lm_add_source(self._h, self._p(v), dt, self._s()))

    def cc_divergence(self, u, v, div, dt=1.0, divide=False):
        assert div.stride(1) == 1
        _lib.check(_lib.lib().p2b_lm_cc_divergence(self._h, self._p(u), self._p(v), div.data_ptr(), div.stride(0), dt,
                                                   int(divide), self._s()))

    def project(self, rho, phi, u, v, gradp_x, gradp_y, dt, proj_type):
        _lib.check(_lib.lib().p2b_lm_project(self._h, self._p(rho), self._p(phi), self._p(u), self._p(v),
                                             self._p(gradp_x), self._p(gradp_y), dt, proj_type, self._s()))

    def reduce(self, rho, u, v, grav):
        """max|u|, max|v| over the whole arrays; max|u|, max|v|, max(|rho' g|/rho) over the valid cells"""
        self._scratch.zero_()
        _lib.check(_lib.lib().p2b_lm_reduce(self._h, self._p(rho), self._p(u), self._p(v), grav,
                                            self._scratch.data_ptr(), self._s()))
        return self._scratch.view(torch.float64).tolist()

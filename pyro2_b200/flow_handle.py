"""Low-level wrapper of the p2b_flow_* C ABI (csrc/flow.cu): the explicit stages of the Burgers and
incompressible solvers.  Owns the torch workspace that stands in for the reference's per-call
``grid.scratch_array()`` temporaries; the solver state planes are passed in by the Simulation classes
(burgers/simulation.py, incompressible/simulation.py)."""
import ctypes as C

import torch

from . import _lib, ops


class FlowHandle:
    PLANES = ["u_xl", "u_xr", "u_yl", "u_yr", "v_xl", "v_xr", "v_yl", "v_yr", "uhat", "vhat",
              "u_xint", "v_xint", "u_yint", "v_yint", "u_MAC", "v_MAC"]

    def __init__(self, planes, grid):
        """planes: the solver's (nvar, qx, pitch) state storage (fixes the pitch); grid: its Grid2d"""
        ops.require_cuda()
        L = _lib.lib()
        self.grid = grid
        self.pitch = planes.stride(1)
        self._g = ops.grid_struct(planes, grid.nx, grid.ny, grid.ng, grid.dx, grid.dy)
        self._h = L.p2b_flow_create(C.byref(self._g))
        if not self._h:
            raise ValueError(L.p2b_last_error().decode())
        nbytes = L.p2b_flow_workspace_bytes(self._h)
        self.workspace = torch.zeros(nbytes // 8, dtype=torch.float64, device=planes.device)
        _lib.check(L.p2b_flow_bind(self._h, self.workspace.data_ptr(), nbytes))
        self._scratch = torch.zeros(2, dtype=torch.int64, device=planes.device)

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().p2b_flow_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _s(self):
        return _lib.stream_ptr()

    def _check_plane(self, t):
        assert t.is_cuda and t.dtype == torch.float64 and t.stride() == (self.pitch, 1), "state plane layout"
        return t.data_ptr()

    def plane(self, name):
        """(qx, qy) view of a scratch plane (u_xl ... v_MAC)"""
        g = self.grid
        ptr = _lib.lib().p2b_flow_plane(self._h, self.PLANES.index(name))
        off = (ptr - self.workspace.data_ptr()) // 8
        return self.workspace.as_strided((g.qx, g.qy), (self.pitch, 1), off)

    def interface_states(self, u, v, gradp_x, gradp_y, dt, limiter):
        p = self._check_plane
        _lib.check(_lib.lib().p2b_flow_interface_states(self._h, p(u), p(v),
                                                        None if gradp_x is None else p(gradp_x),
                                                        None if gradp_y is None else p(gradp_y), dt, limiter, self._s()))

    def mac_vels(self):
        _lib.check(_lib.lib().p2b_flow_mac_vels(self._h, self._s()))

    def mac_divergence(self, div):
        """div: (nx+2, ny+2) float64 view (unit column stride) on the multigrid grid"""
        assert div.stride(1) == 1
        _lib.check(_lib.lib().p2b_flow_mac_divergence(self._h, div.data_ptr(), div.stride(0), self._s()))

    def mac_project(self, phi_mac):
        _lib.check(_lib.lib().p2b_flow_mac_project(self._h, self._check_plane(phi_mac), self._s()))

    def upwind_states(self):
        _lib.check(_lib.lib().p2b_flow_upwind_states(self._h, self._s()))

    def advect_update(self, u, v, gradp_x, gradp_y, dt, proj_type):
        p = self._check_plane
        _lib.check(_lib.lib().p2b_flow_advect_update(self._h, p(u), p(v), p(gradp_x), p(gradp_y), dt, proj_type, self._s()))

    def cc_divergence(self, u, v, div, dt=1.0, divide=False):
        assert div.stride(1) == 1
        p = self._check_plane
        _lib.check(_lib.lib().p2b_flow_cc_divergence(self._h, p(u), p(v), div.data_ptr(), div.stride(0), dt,
                                                     int(divide), self._s()))

    def project(self, phi, u, v, gradp_x, gradp_y, dt, proj_type):
        p = self._check_plane
        _lib.check(_lib.lib().p2b_flow_project(self._h, p(phi), p(u), p(v),
                                               None if gradp_x is None else p(gradp_x),
                                               None if gradp_y is None else p(gradp_y), dt, proj_type, self._s()))

    def burgers_update(self, u, v, dt):
        p = self._check_plane
        _lib.check(_lib.lib().p2b_flow_burgers_update(self._h, p(u), p(v), dt, self._s()))

    def advection_update(self, a, u, v, dt, limiter):
        _lib.check(_lib.lib().p2b_flow_advection_update(self._h, self._check_plane(a), u, v, dt, limiter, self._s()))

    def maxabs(self, u, v):
        """(max|u|, max|v|) over the full arrays including ghost cells, as python floats"""
        p = self._check_plane
        self._scratch.zero_()
        _lib.check(_lib.lib().p2b_flow_maxabs(self._h, p(u), p(v), self._scratch.data_ptr(), self._s()))
        a, b = self._scratch.view(torch.float64).tolist()
        return a, b

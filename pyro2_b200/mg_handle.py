"""Low-level wrapper of the p2b_mg_* C ABI: owns the torch workspace the hierarchy lives in and
exposes each level's v / f / r planes as torch views.  multigrid/MG.py builds the reference's
CellCenterMG2d interface on top of this."""
import ctypes as C

import torch

from . import _lib
from . import ops
from .ops import require_cuda


class MGHandle:
    def __init__(self, nx, bc, alpha, beta, xmin, xmax, ymin, ymax, nsmooth, nsmooth_bottom,
                 decomposition=None, split_n=1024):
        """decomposition: parallel.SlabDecomposition -- this process then holds x-slab `rank` of every
        level with >= split_n columns and a full copy of the coarser ones"""
        require_cuda()
        L = _lib.lib()
        codes = (C.c_int * 4)(*[_lib.BC_CODES[b] for b in bc])
        self.decomp = decomposition if (decomposition is not None and decomposition.size > 1) else None
        if self.decomp is None:
            self._h = L.p2b_mg_create(nx, codes, alpha, beta, xmin, xmax, ymin, ymax, nsmooth, nsmooth_bottom)
        else:
            self._h = L.p2b_mg_create_slab(nx, codes, alpha, beta, xmin, xmax, ymin, ymax, nsmooth,
                                           nsmooth_bottom, self.decomp.rank, self.decomp.size, split_n)
        if not self._h:
            raise ValueError(L.p2b_last_error().decode())
        self.nlevels = L.p2b_mg_nlevels(self._h)
        self.tb_halo = L.p2b_mg_tb_halo()
        self.tb_iters = L.p2b_mg_tb_iters()
        self.xperiodic = bc[0] == "periodic"
        self._info = {}
        nbytes = L.p2b_mg_workspace_bytes(self._h)
        self._shared_ptr = None
        self._peer_ptrs = None
        if self.decomp is None:
            self.workspace = torch.zeros(nbytes // 8, dtype=torch.float64, device="cuda")
        else:
            # the slabs talk through each other's workspaces (csrc/mg_kernels.cuh, "peer-memory communication"): the
            # allocation must be mappable by the other ranks, so the library owns it
            self._shared_ptr = L.p2b_shared_alloc(nbytes)
            if not self._shared_ptr:
                raise RuntimeError(L.p2b_last_error().decode())
            self.workspace = ops.tensor_from_pointer(self._shared_ptr, nbytes // 8)
        _lib.check(L.p2b_mg_bind(self._h, self.workspace.data_ptr(), nbytes))
        if self.decomp is not None:
            self._peer_ptrs = self.decomp.map_peer_workspaces(self._shared_ptr)
            arr = (C.c_void_p * len(self._peer_ptrs))(*self._peer_ptrs)
            _lib.check(L.p2b_mg_set_peers(self._h, arr))
        self._bc_vals = [None] * 4
        self._out = torch.zeros(2, dtype=torch.float64, device="cuda")
        self._planes = {}

    def close(self):
        if getattr(self, "_h", None):
            L = _lib.lib()
            L.p2b_mg_destroy(self._h)
            self._h = None
            if self._peer_ptrs is not None and type(self.decomp).__name__ == "SlabDecomposition":
                for r, p in enumerate(self._peer_ptrs):
                    if r != self.decomp.rank:
                        L.p2b_shared_close(C.c_void_p(p))
            # the shared allocation itself is released when the process ends: a neighbour may still have it mapped

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    _WHICH = {"v": 0, "f": 1, "r": 2, "w": 3}

    def info(self, level):
        """geometry of a level: owned rows ni, columns n, pitch, halo rows gx, global row offset ioff,
        slab flag, plane stride (elements), first slab level"""
        if level not in self._info:
            out = (C.c_longlong * 8)()
            _lib.check(_lib.lib().p2b_mg_level_info(self._h, level, out))
            keys = ("ni", "n", "pitch", "gx", "ioff", "slab", "plane_stride", "split_level")
            self._info[level] = dict(zip(keys, [int(x) for x in out]))
        return self._info[level]

    def _row0_offset(self, level, which):
        ptr = _lib.lib().p2b_mg_level_ptr(self._h, level, self._WHICH[which])
        return (ptr - self.workspace.data_ptr()) // 8

    def plane(self, level, which):
        """(ni+2, n+2) strided view (one ghost row / column all round) of plane which in
        {'v','f','r','w'} on `level`; ni = n unless the level is a slab"""
        key = (level, which)
        if key not in self._planes:
            g = self.info(level)
            self._planes[key] = self.workspace.as_strided((g["ni"] + 2, g["n"] + 2), (g["pitch"], 1),
                                                          self._row0_offset(level, which))
        return self._planes[key]

    def halo_rows(self, level, which, depth):
        """contiguous (ni + 2*depth, pitch) view: `depth` halo rows, the owned rows, `depth` halo rows"""
        g = self.info(level)
        assert depth <= g["gx"]
        off = self._row0_offset(level, which) + (1 - depth) * g["pitch"]
        return self.workspace.as_strided((g["ni"] + 2 * depth, g["pitch"]), (g["pitch"], 1), off)

    def exchange(self, level, which, depth):
        """halo exchange of `depth` rows with the neighbouring slabs through peer memory (no-op on replicated levels;
        collective)"""
        g = self.info(level)
        if self.decomp is None or not g["slab"]:
            return
        _lib.check(_lib.lib().p2b_mg_exchange(self._h, level, self._WHICH[which], depth, self._s()))

    def tb_pass(self, level, src, dst, niter):
        _lib.check(_lib.lib().p2b_mg_tb_pass(self._h, level, self._WHICH[src], self._WHICH[dst], niter, self._s()))

    def vcycle_level(self, level):
        _lib.check(_lib.lib().p2b_mg_vcycle_level(self._h, level, self._s()))

    def set_bc_values(self, xl=None, xr=None, yl=None, yr=None):
        vals = []
        for v in (xl, xr, yl, yr):
            vals.append(None if v is None else torch.as_tensor(v, dtype=torch.float64).contiguous().cuda())
        self._bc_vals = vals   # keep alive
        _lib.check(_lib.lib().p2b_mg_set_bc_values(self._h, *[None if v is None else v.data_ptr() for v in vals]))

    def set_coeffs(self, coeffs, coeffs_bc):
        """variable-coefficient mode: coeffs = eta on the finest level, an (n+2, n+2) CUDA float64 tensor
        (any row stride, unit column stride); coeffs_bc = its four boundary-type names"""
        L = _lib.lib()
        assert self.decomp is None, "variable coefficients: single-GPU hierarchies only"
        assert coeffs.is_cuda and coeffs.dtype == torch.float64 and coeffs.stride(1) == 1
        nbytes = L.p2b_mg_coeff_workspace_bytes(self._h)
        if getattr(self, "coeff_workspace", None) is None:      # kept across calls: views and graphs point into it
            self.coeff_workspace = torch.zeros(nbytes // 8, dtype=torch.float64, device="cuda")
        codes = (C.c_int * 4)(*[_lib.BC_CODES[b] for b in coeffs_bc])
        _lib.check(L.p2b_mg_set_coeffs(self._h, self.coeff_workspace.data_ptr(), nbytes, coeffs.data_ptr(),
                                       coeffs.stride(0), codes, self._s()))

    def coeff_plane(self, level, which):
        """(n+2, n+2) view of the level's eta ('c'), eta_x ('x') or eta_y ('y') plane"""
        g = self.info(level)
        ptr = _lib.lib().p2b_mg_coeff_ptr(self._h, level, {"c": 0, "x": 1, "y": 2}[which])
        if not ptr:
            raise ValueError("no coefficients set")
        off = (ptr - self.coeff_workspace.data_ptr()) // 8
        return self.coeff_workspace.as_strided((g["n"] + 2, g["n"] + 2), (g["pitch"], 1), off)

    def set_operator(self, alpha, beta):
        _lib.check(_lib.lib().p2b_mg_set_operator(self._h, alpha, beta))

    def cn_rhs(self, phi, coef):
        """finest-level f <- phi + coef * lap(phi) for a ghost-filled (n+2, n+2) CUDA plane phi"""
        assert phi.is_cuda and phi.dtype == torch.float64 and phi.stride(1) == 1
        _lib.check(_lib.lib().p2b_mg_cn_rhs(self._h, phi.data_ptr(), phi.stride(0), coef, self._s()))

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
        # on a slab level the library sums over the ranks (rank order, same bits everywhere)
        _lib.check(_lib.lib().p2b_mg_norm2(self._h, level, idx, self._out.data_ptr(), self._s()))
        return float(self._out[0])

    def cycle_diagnostics_enqueue(self, old_phi):
        """device part of cycle_diagnostics (capturable in a CUDA graph): results land in self._out"""
        _lib.check(_lib.lib().p2b_mg_cycle_diagnostics(self._h, old_phi.data_ptr(), self._out.data_ptr(), self._s()))

    def set_stop(self, enable, source_norm=0.0, rtol=0.0, max_cycles=0):
        """arm / disarm the device-side stopping rule of solve(); clears the stop word and the cycle count"""
        _lib.check(_lib.lib().p2b_mg_set_stop(self._h, int(enable), source_norm, rtol, max_cycles, self._s()))

    def result(self):
        """(relsq, rsq, residual_error, cycles) of the last counted cycle; synchronises.  Raises if a wait on another
        rank timed out."""
        out = (C.c_double * 4)()
        err = C.c_longlong(0)
        _lib.check(_lib.lib().p2b_mg_result(self._h, out, C.byref(err), self._s()))
        if err.value:
            w = self.control_words()
            raise RuntimeError(f"multigrid: a wait on another rank timed out (control word {err.value - 1}); rank "
                               f"{self.decomp.rank if self.decomp else 0}: epoch {w[0]} from_lo {w[3]} from_hi {w[4]} counters "
                               f"{w[5:8]} all-rank words {w[8:8 + (self.decomp.size if self.decomp else 1)]}")
        return tuple(out)

    def control_words(self):
        """the hierarchy's control words (csrc/mg_kernels.cuh, CW_*) as a list of ints -- diagnostics"""
        ptr = _lib.lib().p2b_mg_control_ptr(self._h)
        off = (ptr - self.workspace.data_ptr()) // 8
        return self.workspace[off:off + 24].view(torch.int64).tolist()

    def cycle_diagnostics(self, old_phi):
        """returns (sum rel-change^2, sum r^2); updates old_phi <- v and the r plane"""
        _lib.check(_lib.lib().p2b_mg_cycle_diagnostics(self._h, old_phi.data_ptr(), self._out.data_ptr(), self._s()))
        a, b = self._out.tolist()
        return a, b

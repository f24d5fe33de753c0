"""Device-resident ArrayIndexer.

Mirror of pyro/mesh/array_indexer.py:29-333: the reference's class is an ``np.ndarray`` subclass
carrying the grid (``.g``) and the rank (``.c``) and offering shifted / strided stencil views
``v / ip / jp / ip_jp``, ``lap``, ``norm``, ``copy``, ``fill_ghost``.  Here it is a
``torch.Tensor`` subclass over CUDA storage with the same methods and the same ``[i, j, n]``
indexing (x slowest, variable last).  Views alias the underlying planes exactly as numpy views do,
so problem setups that write ``dens[:, :] = ...`` or ``ener[i, j] = ...`` keep working.

``fill_ghost`` dispatches to the CUDA ghost-fill kernels (csrc/ghost_cfl.cu), bit-exact for
float64 and int64; there is no host fallback.
"""
import numbers

import numpy as np
import torch


def _buf_split(b):
    """int, (lo, hi) or (xlo, xhi, ylo, yhi) -> the four buffer widths (array_indexer.py:12-26)"""
    try:
        bxlo, bxhi, bylo, byhi = b
    except (ValueError, TypeError):
        try:
            blo, bhi = b
        except (ValueError, TypeError):
            blo = bhi = b
        bxlo = bylo = blo
        bxhi = byhi = bhi
    return bxlo, bxhi, bylo, byhi


def _as_tensor(value, like):
    if isinstance(value, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(value)).to(device=like.device, dtype=like.dtype)
    if isinstance(value, (list, tuple)):
        return torch.as_tensor(value, device=like.device, dtype=like.dtype)
    return value


class DeviceView(torch.Tensor):
    """what the stencil accessors return: a torch view that, like the numpy views of the reference,
    accepts numpy arrays / lists on assignment and converts to numpy on request"""

    def __setitem__(self, key, value):
        torch.Tensor.__setitem__(self.as_subclass(torch.Tensor), key, _as_tensor(value, self))

    def numpy(self):   # pylint: disable=arguments-differ
        """host copy (device -> host), mostly for tests and output"""
        return self.as_subclass(torch.Tensor).detach().cpu().numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a.astype(dtype) if dtype is not None else a


class ArrayIndexer(DeviceView):
    """a tensor that knows its grid; ``d`` may be a torch tensor (aliased, not copied) or array-like"""

    @staticmethod
    def __new__(cls, d, grid=None):
        if not isinstance(d, torch.Tensor):
            d = torch.as_tensor(np.asarray(d))
        obj = d.as_subclass(cls)
        obj.g = grid
        obj.c = d.dim()
        return obj

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        out = super().__torch_function__(func, types, args, kwargs or {})
        if isinstance(out, ArrayIndexer) and getattr(out, "g", None) is None:
            for a in args:
                g = getattr(a, "g", None) if isinstance(a, ArrayIndexer) else None
                if g is not None:
                    out.g = g
                    out.c = out.dim()
                    break
        return out

    # ---- plain-tensor escape hatch (the reference returns np.asarray views) --------------------
    def t(self):
        return self.as_subclass(torch.Tensor)

    # ---- stencil views (array_indexer.py:49-90) -------------------------------------------------
    def v(self, buf=0, n=0, s=1):
        return self.ip_jp(0, 0, buf=buf, n=n, s=s)

    def ip(self, shift, buf=0, n=0, s=1):
        return self.ip_jp(shift, 0, buf=buf, n=n, s=s)

    def jp(self, shift, buf=0, n=0, s=1):
        return self.ip_jp(0, shift, buf=buf, n=n, s=s)

    def ip_jp(self, ishift, jshift, buf=0, n=0, s=1):
        bxlo, bxhi, bylo, byhi = _buf_split(buf)
        g = self.g
        t = self.t()
        si = slice(g.ilo - bxlo + ishift, g.ihi + 1 + bxhi + ishift, s)
        sj = slice(g.jlo - bylo + jshift, g.jhi + 1 + byhi + jshift, s)
        if t.dim() == 2:
            return t[si, sj].as_subclass(DeviceView)
        return t[si, sj, n].as_subclass(DeviceView)

    def lap(self, n=0, buf=0):
        """5-point Laplacian (array_indexer.py:92-96)"""
        return (self.ip(-1, n=n, buf=buf) - 2 * self.v(n=n, buf=buf) + self.ip(1, n=n, buf=buf)) / self.g.dx ** 2 + \
               (self.jp(-1, n=n, buf=buf) - 2 * self.v(n=n, buf=buf) + self.jp(1, n=n, buf=buf)) / self.g.dy ** 2

    def norm(self, n=0):
        """sqrt(dx dy sum(valid^2)) (array_indexer.py:98-111)"""
        a = self.v(n=n)
        return float(torch.sqrt(self.g.dx * self.g.dy * torch.sum(a.to(torch.float64) ** 2)))

    def copy(self, order="C"):   # pylint: disable=unused-argument
        """a new array on the same grid (array_indexer.py:113-115)"""
        return ArrayIndexer(self.t().clone(), grid=self.g)

    def is_symmetric(self, nodal=False, tol=1.e-14, asymmetric=False):
        """left-right symmetry test (array_indexer.py:117-140)"""
        s = -1 if asymmetric else 1
        g = self.g
        t = self.t()
        if not nodal:
            L = t[g.ilo:g.ilo + g.nx // 2, g.jlo:g.jhi + 1]
            R = t[g.ilo + g.nx // 2:g.ihi + 1, g.jlo:g.jhi + 1]
        else:
            L = t[g.ilo:g.ilo + g.nx // 2 + 1, g.jlo:g.jhi + 1]
            R = t[g.ilo + g.nx // 2:g.ihi + 2, g.jlo:g.jhi + 1]
        return float((L - s * torch.flip(R, dims=(0,))).abs().max()) < tol

    def is_asymmetric(self, nodal=False, tol=1.e-14):
        return self.is_symmetric(nodal=nodal, tol=tol, asymmetric=True)

    # ---- ghost fill (array_indexer.py:150-274) --------------------------------------------------
    def fill_ghost(self, n=0, bc=None):
        """fill the ghost cells of component n according to the BC object, on the device"""
        from .. import ops
        g = self.g
        t = self.t()
        plane = t if t.dim() == 2 else t[:, :, n]
        if plane.stride(1) != 1:
            raise ValueError("fill_ghost needs y-contiguous storage (a view of a CellCenterData2d plane)")
        vals = (bc.xl_value, bc.xr_value, bc.yl_value, bc.yr_value)
        if all(v is None for v in vals):
            ops.fill_ghost(plane.unsqueeze(0), g.nx, g.ny, g.ng, [bc.names()])
        else:
            if plane.dtype != torch.float64:
                raise TypeError("inhomogeneous boundary values need float64 data")
            dev = [None if v is None else torch.as_tensor(np.asarray(v, dtype=np.float64)).cuda() for v in vals]
            ops.fill_ghost_values(plane, g.nx, g.ny, g.ng, bc.names(), g.dx, g.dy, *dev)

    def pretty_print(self, n=0, fmt=None, show_ghost=True):
        """print a small array with the ghost cells highlighted (array_indexer.py:276-333)"""
        a = self.numpy()
        if fmt is None:
            if issubclass(a.dtype.type, numbers.Integral):
                fmt = "%4d"
            elif a.dtype == np.float64:
                fmt = "%10.5g"
            else:
                raise ValueError("ERROR: dtype not supported")
        g = self.g
        ilo, ihi, jlo, jhi = (0, g.qx - 1, 0, g.qy - 1) if show_ghost else (g.ilo, g.ihi, g.jlo, g.jhi)
        for j in reversed(range(jlo, jhi + 1)):
            for i in range(ilo, ihi + 1):
                ghost = j < g.jlo or j > g.jhi or i < g.ilo or i > g.ihi
                val = a[i, j] if a.ndim == 2 else a[i, j, n]
                print(("\033[31m" + fmt % val + "\033[0m") if ghost else fmt % val, end="")
            print(" ")
        print("\n         ^ y\n         |\n         +---> x\n        ")

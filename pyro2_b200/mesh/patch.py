"""Grid and cell-centred data containers with device-resident storage.

Mirror of pyro/mesh/patch.py: Grid2d (:42-190), Cartesian2d (:192-239), CellCenterData2d (:315-795).
Same attribute names and call surface; differences that matter on a GPU:

* state lives in HBM as structure-of-arrays planes ``planes[n, i, j]`` (float64 by default) with a
  128-byte aligned row pitch; ``data`` is the permuted ``[i, j, n]`` view the reference exposes;
* the 2-d coordinate / geometry arrays (x2d, y2d, Lx, Ly, Ax, Ay, V, ...) are built lazily on first
  access -- the reference allocates 13 full-size arrays per grid up front (patch.py:137-147,
  :210-233); a Cartesian sweep only needs the scalars dx, dy;
* ``fill_BC`` / ``fill_BC_all`` run CUDA kernels; Python ``ext_bcs`` callbacks still work (they see
  ArrayIndexer views of device memory).

SphericalPolar and FaceCenterData2d are out of scope (no hot-path caller).
"""
import numpy as np
import torch

from .. import ops
from ..util import msg
from . import boundary as bnd
from .array_indexer import ArrayIndexer

_TORCH_DTYPES = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32,
                 np.dtype(np.int64): torch.int64, np.dtype(np.int32): torch.int32}


def _torch_dtype(dtype):
    if isinstance(dtype, torch.dtype):
        return dtype
    return _TORCH_DTYPES[np.dtype(dtype)]


def _default_device():
    ops.require_cuda()
    return torch.device("cuda")


class Grid2d:
    """the 2-d grid: index space, coordinates, scratch allocation (patch.py:42-190)"""

    def __init__(self, nx, ny, *, ng=1, xmin=0.0, xmax=1.0, ymin=0.0, ymax=1.0, device=None,
                 nx_global=None, ioffset=0):
        """nx_global / ioffset (extension): this grid is the x-slab [ioffset, ioffset + nx) of a global
        grid of nx_global zones spanning [xmin, xmax]; dx and the coordinates are then computed with
        the global formulas so they are bit-identical to the single-domain grid's."""
        self.nx, self.ny, self.ng = int(nx), int(ny), int(ng)
        self.nx_global = int(nx_global) if nx_global is not None else self.nx
        self.ioffset = int(ioffset)
        self.qx = int(2 * ng + nx)
        self.qy = int(2 * ng + ny)
        self.xmin, self.xmax, self.ymin, self.ymax = xmin, xmax, ymin, ymax
        self.ilo, self.ihi = self.ng, self.ng + self.nx - 1
        self.jlo, self.jhi = self.ng, self.ng + self.ny - 1
        self.ic = self.ilo + self.nx // 2 - 1
        self.jc = self.jlo + self.ny // 2 - 1
        self.device = torch.device(device) if device is not None else _default_device()

        # 1-d coordinates stay on the host (problem setups and BC callbacks index them)
        self.dx = (xmax - xmin) / self.nx_global
        self.xl = (np.arange(self.qx) + self.ioffset - ng) * self.dx + xmin
        self.xr = (np.arange(self.qx) + self.ioffset + 1.0 - ng) * self.dx + xmin
        self.x = 0.5 * (self.xl + self.xr)
        self.dy = (ymax - ymin) / ny
        self.yl = (np.arange(self.qy) - ng) * self.dy + ymin
        self.yr = (np.arange(self.qy) + 1.0 - ng) * self.dy + ymin
        self.y = 0.5 * (self.yl + self.yr)
        self._lazy = {}

    # 2-d coordinate arrays, built on first use (patch.py:137-147)
    def _mesh(self, name, xs, ys, which):
        if name not in self._lazy:
            X = torch.from_numpy(xs).to(self.device)
            Y = torch.from_numpy(ys).to(self.device)
            X2, Y2 = torch.meshgrid(X, Y, indexing="ij")
            self._lazy[name] = ArrayIndexer((X2 if which == 0 else Y2).contiguous(), grid=self)
        return self._lazy[name]

    x2d = property(lambda self: self._mesh("x2d", self.x, self.y, 0))
    y2d = property(lambda self: self._mesh("y2d", self.x, self.y, 1))
    xl2d = property(lambda self: self._mesh("xl2d", self.xl, self.yl, 0))
    yl2d = property(lambda self: self._mesh("yl2d", self.xl, self.yl, 1))
    xr2d = property(lambda self: self._mesh("xr2d", self.xr, self.yr, 0))
    yr2d = property(lambda self: self._mesh("yr2d", self.xr, self.yr, 1))

    def scratch_array(self, *, nvar=1, dtype=np.float64):
        """zeroed array with the grid's shape and ghost cells (patch.py:149-158); multi-variable
        scratch uses the same SoA storage as the state, exposed as [i, j, n]"""
        td = _torch_dtype(dtype)
        if self.device.type == "cuda":
            planes = ops.alloc_planes(max(nvar, 1), self.qx, self.qy, dtype=td, device=self.device)
        else:
            planes = torch.zeros((max(nvar, 1), self.qx, ops.row_pitch(self.qy)), dtype=td)
        if nvar == 1:
            return ArrayIndexer(planes[0, :, :self.qy], grid=self)
        return ArrayIndexer(planes.permute(1, 2, 0)[:, :self.qy, :], grid=self)

    def _like(self, nx, ny):
        return type(self)(nx, ny, ng=self.ng, xmin=self.xmin, xmax=self.xmax, ymin=self.ymin,
                          ymax=self.ymax, device=self.device)

    def coarse_like(self, N):
        return self._like(self.nx // N, self.ny // N)

    def fine_like(self, N):
        return self._like(self.nx * N, self.ny * N)

    def __str__(self):
        return f"2-d grid: nx = {self.nx}, ny = {self.ny}, ng = {self.ng}"

    def __eq__(self, other):
        return (self.nx == other.nx and self.ny == other.ny and self.ng == other.ng and
                self.xmin == other.xmin and self.xmax == other.xmax and
                self.ymin == other.ymin and self.ymax == other.ymax)

    __hash__ = None


class Cartesian2d(Grid2d):
    """Cartesian geometry (patch.py:192-239): coord_type 0, constant face lengths / areas / volume"""

    def __init__(self, nx, ny, *, ng=1, xmin=0.0, xmax=1.0, ymin=0.0, ymax=1.0, device=None,
                 nx_global=None, ioffset=0):
        super().__init__(nx, ny, ng=ng, xmin=xmin, xmax=xmax, ymin=ymin, ymax=ymax, device=device,
                         nx_global=nx_global, ioffset=ioffset)
        self.coord_type = 0

    def _full(self, name, value):
        if name not in self._lazy:
            self._lazy[name] = ArrayIndexer(torch.full((self.qx, self.qy), value, dtype=torch.float64,
                                                       device=self.device), grid=self)
        return self._lazy[name]

    Lx = property(lambda self: self._full("Lx", self.dx))
    Ly = property(lambda self: self._full("Ly", self.dy))
    Ax = property(lambda self: self.Ly)
    Ay = property(lambda self: self.Lx)
    dlogAx = property(lambda self: self._full("dlogA", 0.0))
    dlogAy = property(lambda self: self._full("dlogA", 0.0))
    V = property(lambda self: self._full("V", self.dx * self.dy))

    def __str__(self):
        return (f"Cartesian 2D Grid: xmin = {self.xmin}, xmax = {self.xmax}, ymin = {self.ymin}, "
                f"ymax = {self.ymax}, nx = {self.nx}, ny = {self.ny}, ng = {self.ng}")


def spherical_sweep_tables(grid, nj, xlb, xrb):
    """the separable geometry tables the SphericalPolar instantiation of the sweep reads (include/pyro2b200.h,
    p2b_comp_params.geo_i / geo_j), as host arrays: geo_i (9, qx), geo_j (7, nj) with nj >= qy + 1.  Every entry is the
    sub-expression of mesh/patch.py:272-305 (or of the viscosity's vertex coordinates, interface.py:333-341) that
    depends on one index only, evaluated with numpy like the reference, so that the kernel's products -- formed in the
    reference's order -- give its 2-d arrays bit for bit.  xlb / xrb: the x boundary types; they decide which row's
    radius the reference's ghost-filled source arrays carry in the ghost rows (row 1 of geo_i); None for a side that
    faces another x-slab"""
    g = grid
    xl, xr, x = g.xl, g.xr, g.x
    ximg = x.copy()
    for side, btype in (("lo", xlb), ("hi", xrb)):
        for k in range(g.ng):
            i = g.ilo - 1 - k if side == "lo" else g.ihi + 1 + k
            if btype is None:                       # the side faces another slab: the halo rows are real cells
                src = i
            elif btype == "periodic":
                src = i + g.nx if side == "lo" else i - g.nx
            elif str(btype).startswith("reflect"):
                src = g.ilo + k if side == "lo" else g.ihi - k
            else:                                   # zero-gradient copies (outflow and everything that fills like it)
                src = g.ilo if side == "lo" else g.ihi
            ximg[i] = x[src]
    idx = np.arange(g.qx) + g.ioffset                # global row index (a slab's rows start at ioffset)
    geo_i = np.stack([x, ximg, -2.0 * np.pi * xl ** 2, xr ** 2 - xl ** 2, xr - xl, xr ** 2 + xl ** 2 + xr * xl,
                      (idx + 0.5 - g.ng) * g.dx + g.xmin, (idx - 0.5 - g.ng) * g.dx + g.xmin, (idx - g.ng) * g.dx + g.xmin])
    jdx = np.arange(nj)
    yl = (jdx - g.ng) * g.dy + g.ymin
    yr = (jdx + 1.0 - g.ng) * g.dy + g.ymin
    y = 0.5 * (yl + yr)
    cd = np.cos(yr) - np.cos(yl)
    geo_j = np.stack([cd, -2.0 * np.pi / 3.0 * cd, np.pi * np.sin(yl), np.tan(y),
                      np.sin((jdx + 0.5 - g.ng) * g.dy + g.ymin), np.sin((jdx - 0.5 - g.ng) * g.dy + g.ymin),
                      np.sin((jdx - g.ng) * g.dy + g.ymin)])
    return np.ascontiguousarray(geo_i), np.ascontiguousarray(geo_j)


class SphericalPolar(Grid2d):
    """spherical polar geometry with azimuthal symmetry, x = r, y = theta (patch.py:242-312): coord_type 1; side
    lengths, face areas (on the low faces), cell volumes and logarithmic area derivatives as device arrays"""

    def __init__(self, nx, ny, *, ng=1, xmin=0.2, xmax=1.0, ymin=0.0, ymax=1.0, device=None,
                 nx_global=None, ioffset=0):
        super().__init__(nx, ny, ng=ng, xmin=xmin, xmax=xmax, ymin=ymin, ymax=ymax, device=device,
                         nx_global=nx_global, ioffset=ioffset)
        assert ymin >= 0.0 and ymax <= np.pi, "y or \u03b8 should be within [0, \u03c0]."
        assert xmin - ng * self.dx >= 0.0, \
            "xmin (r-direction), must be large enough so ghost cell doesn't have negative x."
        self.coord_type = 1

    def _geom(self, name):
        if name not in self._lazy:
            x2d, _ = np.meshgrid(self.x, self.y, indexing="ij")
            xl2d, yl2d = np.meshgrid(self.xl, self.yl, indexing="ij")
            xr2d, yr2d = np.meshgrid(self.xr, self.yr, indexing="ij")
            y2d = np.meshgrid(self.x, self.y, indexing="ij")[1]
            host = {"Lx": lambda: np.full((self.qx, self.qy), self.dx), "Ly": lambda: x2d * self.dy,
                    "Ax": lambda: np.abs(-2.0 * np.pi * xl2d ** 2 * (np.cos(yr2d) - np.cos(yl2d))),
                    "Ay": lambda: np.abs(np.pi * np.sin(yl2d) * (xr2d ** 2 - xl2d ** 2)),
                    "dlogAx": lambda: 2.0 / x2d, "dlogAy": lambda: 1.0 / (np.tan(y2d) * x2d),
                    "V": lambda: np.abs(-2.0 * np.pi / 3.0 * (np.cos(yr2d) - np.cos(yl2d)) * (xr2d - xl2d) *
                                        (xr2d ** 2 + xl2d ** 2 + xr2d * xl2d))}[name]()
            self._lazy[name] = ArrayIndexer(torch.from_numpy(np.ascontiguousarray(host)).to(self.device), grid=self)
        return self._lazy[name]

    Lx = property(lambda self: self._geom("Lx"))
    Ly = property(lambda self: self._geom("Ly"))
    Ax = property(lambda self: self._geom("Ax"))
    Ay = property(lambda self: self._geom("Ay"))
    dlogAx = property(lambda self: self._geom("dlogAx"))
    dlogAy = property(lambda self: self._geom("dlogAy"))
    V = property(lambda self: self._geom("V"))

    def __str__(self):
        return ("Spherical Polar 2D Grid: Define x : r, y : \u03b8. " +
                f"xmin (r) = {self.xmin}, xmax= {self.xmax}, ymin = {self.ymin}, ymax = {self.ymax}, "
                f"nx = {self.nx}, ny = {self.ny}, ng = {self.ng}")


class CellCenterData2d:
    """named cell-centred variables on a grid (patch.py:315-795): register_var / set_aux / create,
    then get_var, fill_BC, restrict, prolong, ..."""

    def __init__(self, grid, *, dtype=np.float64):
        self.grid = grid
        self.dtype = dtype
        self.planes = None          # (nvar, qx, pitch) SoA storage
        self.names = []
        self.vars = self.names      # backwards-compatibility alias kept by the reference
        self.nvar = 0
        self.ivars = []
        self.aux = {}
        self.derives = []
        self.BCs = {}
        self.t = -1.0
        self.initialized = 0
        self.version = 0            # bumped whenever user code may have modified the data
        self.decomposition = None   # parallel.SlabDecomposition when the grid is one x-slab of many

    def register_var(self, name, bc):
        if self.initialized == 1:
            msg.fail("ERROR: grid already initialized")
        self.names.append(name)
        self.nvar += 1
        self.BCs[name] = bc

    def set_aux(self, keyword, value):
        self.aux[keyword] = value

    def add_derived(self, func):
        self.derives.append(func)

    def add_ivars(self, ivars):
        self.ivars = ivars

    def create(self, planes=None):
        """allocate the storage (patch.py:441-454), or adopt existing SoA planes (multigrid levels)"""
        if self.initialized == 1:
            msg.fail("ERROR: grid already initialized")
        g = self.grid
        if planes is None:
            td = _torch_dtype(self.dtype)
            if g.device.type == "cuda":
                planes = ops.alloc_planes(self.nvar, g.qx, g.qy, dtype=td, device=g.device)
            else:
                planes = torch.zeros((self.nvar, g.qx, ops.row_pitch(g.qy)), dtype=td)
        assert planes.shape[0] == self.nvar and planes.shape[1] == g.qx and planes.stride(2) == 1
        self.planes = planes
        self.initialized = 1

    # the reference's [i, j, n] array; a fresh view of the current planes on every access
    @property
    def data(self):
        self.version += 1
        return ArrayIndexer(self.planes.permute(1, 2, 0)[:, :self.grid.qy, :], grid=self.grid)

    def __str__(self):
        if self.initialized == 0:
            return "CellCenterData2d object not yet initialized"
        g = self.grid
        s = f"cc data: nx = {g.nx}, ny = {g.ny}, ng = {g.ng}\n         nvars = {self.nvar}\n         variables:\n"
        for name in self.names:
            b = self.BCs[name]
            s += f"{name:>16s}: min: {float(self.min(name)):15.10f}    max: {float(self.max(name)):15.10f}\n"
            s += f"{' ':>16s}  BCs: -x: {b.xlb:12s} +x: {b.xrb:12s} -y: {b.ylb:12s} +y: {b.yrb:12s}\n"
        return s

    def get_var(self, name):
        """stored variable (aliasing view) or derived variable (patch.py:480-512)"""
        try:
            n = self.names.index(name)
        except ValueError:
            for f in self.derives:
                try:
                    var = f(self, name)
                except TypeError:
                    var = f(self, name, self.ivars, self.grid)
                if len(var) > 0:
                    return var
            raise KeyError(f"name {name} is not valid") from None
        return self.get_var_by_index(n)

    def get_var_by_index(self, n):
        self.version += 1
        return ArrayIndexer(self.planes[n, :, :self.grid.qy], grid=self.grid)

    def get_vars(self):
        return self.data

    def get_aux(self, keyword):
        return self.aux.get(keyword)

    def zero(self, name):
        self.planes[self.names.index(name)].zero_()
        self.version += 1

    def fill_BC_all(self):
        """all variables in one pair of launches (patch.py:575-580)"""
        g = self.grid
        bcs = [self.BCs[name] for name in self.names]
        if self.decomposition is not None and self.decomposition.size > 1:
            self._fill_BC_all_slab(bcs)
            return
        has_values = any(v is not None for b in bcs for v in (b.xl_value, b.xr_value, b.yl_value, b.yr_value))
        has_ext = any(t in bnd.ext_bcs for b in bcs for t in b.names())
        if has_values or has_ext or g.device.type != "cuda":
            for name in self.names:
                self.fill_BC(name)
            return
        ops.fill_ghost(self.planes, g.nx, g.ny, g.ng, [b.names() for b in bcs])

    def _fill_BC_all_slab(self, bcs, planes=None, first=0):
        """x-slab of a decomposed domain: neighbour rows first (they are the "x fill" of interior
        sides), then the physical x sides and the y sides over the full x range -- the same order as
        the single-domain fill (array_indexer.py:164-274), so corners come out identical.  planes: the planes of the
        variables bcs describes (default: all of them), the first of which is variable number `first`"""
        g = self.grid
        planes = self.planes if planes is None else planes
        periodic = bcs[0].xlb == "periodic"
        has_ext = any(t in bnd.ext_bcs for b in bcs for t in b.names())
        # The halo rows of an inter-slab boundary stand for ordinary interior cells of the single domain.  Those across the
        # PERIODIC SEAM (low side of the first slab, high side of the last) stand for its x ghost rows, and there the
        # reference's variable-by-variable fill matters when a user hook reads other variables: the hse energy takes the
        # momenta of the base row over the full x range, x ghost rows included, and the momenta are refilled only AFTER the
        # energy (patch.py:575-624, compressible/BC.py:62-76) -- so it sees their ghost rows as the previous fill left
        # them.  Reproduced here: the seam rows are kept as they were, and each variable's fresh rows are put in place
        # when its turn comes.  (Without this a decomposed periodic-x / hse-y run differs from the single-domain run in the
        # corner ghost energies: found by scripts/fuzz_compressible_slabs_gloo.py.)
        seam = []
        if has_ext and periodic and len(bcs) > 1:
            if self.decomposition.rank == 0:
                seam.append(slice(0, g.ng))
            if self.decomposition.rank == self.decomposition.size - 1:
                seam.append(slice(g.ng + g.nx, g.ng + g.nx + g.ng))
        kept = [planes[:, rows, :].clone() for rows in seam]
        self.decomposition.exchange(planes, g.nx, g.ng, periodic=periodic)
        fresh = [planes[:, rows, :].clone() for rows in seam]
        for rows, old in zip(seam, kept):
            planes[:, rows, :] = old
        lo_int, hi_int = self.decomposition.interior_sides(periodic)
        names = []
        for b in bcs:
            n = [None if t in bnd.ext_bcs else t for t in b.names()]     # user types: filled by their hooks below
            if lo_int:
                n[0] = None
            if hi_int:
                n[1] = None
            names.append(tuple(n))
        if not has_ext:
            ops.fill_ghost(planes, g.nx, g.ny, g.ng, names)
            return
        # user-defined boundaries: variable by variable like the single-domain fill (standard types, then the hooks in
        # the order xlb, xrb, ylb, yrb -- patch.py:582-624), after ALL halo rows have arrived: a hook may read other
        # variables (the hse energy uses the base row's density and momenta), and the halo rows stand for cells that
        # are ordinary interior cells of the single domain.  Hooks on an x side run on the physical sides only.
        # (fill_BC(name) exchanges that variable's planes alone: a hook that reads OTHER variables then sees their halo
        # rows as of the last full fill -- the solvers here only ever fill all variables together.)
        for k, b in enumerate(bcs):
            for rows, new in zip(seam, fresh):
                planes[k, rows, :] = new[k]
            ops.fill_ghost(planes[k:k + 1], g.nx, g.ny, g.ng, [names[k]])
            name = self.names[first + k]
            for side, btype, interior in zip(("xlb", "xrb", "ylb", "yrb"), b.names(), (lo_int, hi_int, False, False)):
                if btype in bnd.ext_bcs and not interior:
                    try:
                        bnd.ext_bcs[btype](btype, side, name, self, self.ivars)
                    except TypeError:
                        bnd.ext_bcs[btype](btype, side, name, self)

    def fill_BC(self, name):
        """one variable: standard types on the device, then any user-registered callbacks
        (patch.py:582-624)"""
        n = self.names.index(name)
        bc = self.BCs[name]
        if self.decomposition is not None and self.decomposition.size > 1:
            self._fill_BC_all_slab([bc], planes=self.planes[n:n + 1], first=n)
            return
        self.get_var_by_index(n).fill_ghost(bc=_StandardOnly(bc))
        for side, btype in zip(("xlb", "xrb", "ylb", "yrb"), bc.names()):
            if btype in bnd.ext_bcs:
                try:
                    bnd.ext_bcs[btype](btype, side, name, self, self.ivars)
                except TypeError:
                    bnd.ext_bcs[btype](btype, side, name, self)

    def min(self, name, *, ng=0):
        return self.get_var(name).v(buf=ng).min()

    def max(self, name, *, ng=0):
        return self.get_var(name).v(buf=ng).max()

    def restrict(self, varname, N=2):
        """average onto a grid coarser by N (patch.py:640-676)"""
        fdata = self.get_var(varname)
        cdata = self.grid.coarse_like(N).scratch_array()
        if N == 2:
            cdata.v()[:, :] = 0.25 * (fdata.v(s=2) + fdata.ip(1, s=2) + fdata.jp(1, s=2) + fdata.ip_jp(1, 1, s=2))
        elif N == 4:
            acc = 0
            for jj in range(4):
                for ii in range(4):
                    acc = acc + fdata.ip_jp(ii, jj, s=4)
            cdata.v()[:, :] = acc / 16.0
        else:
            raise ValueError("restriction is only allowed by 2 or 4")
        return cdata

    def prolong(self, varname):
        """linear reconstruction onto a grid finer by 2 with centred slopes (patch.py:678-736)"""
        cdata = self.get_var(varname)
        fdata = self.grid.fine_like(2).scratch_array()
        m_x = 0.5 * (cdata.ip(1) - cdata.ip(-1))
        m_y = 0.5 * (cdata.jp(1) - cdata.jp(-1))
        fdata.v(s=2)[:, :] = cdata.v() - 0.25 * m_x - 0.25 * m_y
        fdata.ip(1, s=2)[:, :] = cdata.v() + 0.25 * m_x - 0.25 * m_y
        fdata.jp(1, s=2)[:, :] = cdata.v() - 0.25 * m_x + 0.25 * m_y
        fdata.ip_jp(1, 1, s=2)[:, :] = cdata.v() + 0.25 * m_x + 0.25 * m_y
        return fdata

    def write(self, filename):
        """HDF5 output needs h5py (patch.py:738-748)"""
        import h5py   # pylint: disable=import-outside-toplevel
        if not filename.endswith(".h5"):
            filename += ".h5"
        with h5py.File(filename, "w") as f:
            self.write_data(f)

    def write_data(self, f):
        """same layout as the reference (patch.py:750-788): valid region only, device -> host copy"""
        gaux = f.create_group("aux")
        for k, v in self.aux.items():
            gaux.attrs[k] = v
        ggrid = f.create_group("grid")
        for k in ("nx", "ny", "ng", "xmin", "xmax", "ymin", "ymax"):
            ggrid.attrs[k] = getattr(self.grid, k)
        if hasattr(self.grid, "coord_type"):
            ggrid.attrs["coord_type"] = self.grid.coord_type
        gstate = f.create_group("state")
        for n, name in enumerate(self.names):
            gvar = gstate.create_group(name)
            gvar.create_dataset("data", data=self.get_var_by_index(n).v().cpu().numpy())
            for side, btype in zip(("xlb", "xrb", "ylb", "yrb"), self.BCs[name].names()):
                gvar.attrs[side] = btype

    def pretty_print(self, var, fmt=None):
        self.get_var(var).pretty_print(fmt=fmt)


class _StandardOnly:
    """view of a BC whose user-defined sides are skipped by the device kernel (they are filled by
    the registered Python callback afterwards, exactly as patch.py:604-624 orders it)"""

    def __init__(self, bc):
        self._bc = bc
        self.xl_value, self.xr_value = bc.xl_value, bc.xr_value
        self.yl_value, self.yr_value = bc.yl_value, bc.yr_value

    def names(self):
        return tuple(None if t in bnd.ext_bcs else t for t in self._bc.names())

r"""Cell-centred multigrid for the constant-coefficient Helmholtz equation
:math:`(\alpha - \beta L)\phi = f` on the B200 -- the interface of pyro/multigrid/MG.py
(CellCenterMG2d :77-778) with the V-cycle executed by CUDA kernels (csrc/mg.cu).

    a = MG.CellCenterMG2d(nx, ny, xl_BC_type="dirichlet", ..., verbose=0)
    a.init_zeros(); a.init_RHS(f(a.x2d, a.y2d)); a.solve(rtol=1.e-11)
    phi = a.get_solution()

Same constructor arguments, attributes (``grids``, ``soln_grid``, ``x2d``, ``num_cycles``,
``residual_error``, ``relative_error``, ``source_norm`` ...) and overridable hooks ``smooth``,
``_compute_residual``, ``v_cycle``.  The hierarchy lives in one torch allocation; each level's
``CellCenterData2d`` aliases its v / f / r planes.  The device arithmetic is unfused and ordered like
the reference, so solutions are bit-identical to the reference's; only the norms (reductions)
differ at round-off, which never changes the cycle count in practice.
"""
import math

import numpy as np
import torch

from ..mesh import boundary as bnd
from ..mesh import patch
from ..mesh.array_indexer import ArrayIndexer
from ..mg_handle import MGHandle
from ..util import msg


def _to_host_1d(v):
    if v is None:
        return None
    if isinstance(v, torch.Tensor):
        v = v.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(v, dtype=np.float64))


class CellCenterMG2d:
    """the multigrid hierarchy and solver (MG.py:77-295 for the constructor contract)"""

    def __init__(self, nx, ny, ng=1,
                 xmin=0.0, xmax=1.0, ymin=0.0, ymax=1.0,
                 xl_BC_type="dirichlet", xr_BC_type="dirichlet",
                 yl_BC_type="dirichlet", yr_BC_type="dirichlet",
                 xl_BC=None, xr_BC=None, yl_BC=None, yr_BC=None,
                 alpha=0.0, beta=-1.0,
                 nsmooth=10, nsmooth_bottom=50,
                 verbose=0,
                 aux_field=None, aux_bc=None,
                 true_function=None, vis=0, vis_title="",
                 decomposition=None, split_n=1024):
        """decomposition / split_n (extension, multi-GPU): a parallel.SlabDecomposition; every level
        with at least split_n columns is then split into x-slabs (this process owns one), coarser
        levels are replicated.  x2d / y2d / init_RHS / get_solution refer to the local slab."""
        if nx != ny:
            raise ValueError("ERROR: multigrid currently requires nx = ny")
        if (xmax - xmin) != (ymax - ymin):
            raise ValueError("ERROR: multigrid currently requires a square domain")
        if ng != 1:
            raise ValueError("ERROR: the device multigrid uses ng = 1 (as every caller in the reference does)")
        if aux_field is not None:
            raise NotImplementedError("aux fields (variable-coefficient subclasses) are not built yet")
        if nx < 2 or nx & (nx - 1):
            raise ValueError("ERROR: multigrid requires nx to be a power of 2")

        self.nx, self.ny, self.ng = nx, ny, ng
        self.xmin, self.xmax, self.ymin, self.ymax = xmin, xmax, ymin, ymax
        self.alpha, self.beta = alpha, beta
        self.nsmooth, self.nsmooth_bottom = nsmooth, nsmooth_bottom
        self.max_cycles = 100
        self.verbose = verbose
        if true_function is not None:
            self.true_function = true_function
        self.small = 1.e-16
        self.initialized_rhs = 0
        self.nlevels = int(math.log(self.nx) / math.log(2.0))
        if 2 ** self.nlevels != nx:   # float log of an exact power of two can round down
            self.nlevels = nx.bit_length() - 1

        bc_names = (xl_BC_type, xr_BC_type, yl_BC_type, yr_BC_type)
        self._h = MGHandle(nx, bc_names, alpha, beta, xmin, xmax, ymin, ymax, nsmooth, nsmooth_bottom,
                           decomposition=decomposition, split_n=split_n)
        assert self._h.nlevels == self.nlevels
        self._decomp = self._h.decomp
        self._split = self._h.info(self.nlevels - 1)["split_level"] if self._decomp is not None else 0

        # grids[0] is the coarsest (2x2), grids[nlevels-1] the finest (MG.py:207-257)
        self.grids = []
        bc = bnd.BC(xlb=xl_BC_type, xrb=xr_BC_type, ylb=yl_BC_type, yrb=yr_BC_type)
        n_t = 2
        for i in range(self.nlevels):
            info = self._h.info(i)
            if info["slab"]:
                my_grid = patch.Grid2d(info["ni"], n_t, ng=self.ng, xmin=xmin, xmax=xmax, ymin=ymin, ymax=ymax,
                                       nx_global=n_t, ioffset=info["ioff"])
            else:
                my_grid = patch.Grid2d(n_t, n_t, ng=self.ng, xmin=xmin, xmax=xmax, ymin=ymin, ymax=ymax)
            lev = patch.CellCenterData2d(my_grid, dtype=np.float64)
            if info["slab"]:
                lev.decomposition = self._decomp
            if i == self.nlevels - 1:
                # inhomogeneous boundary values apply to phi on the finest level only
                bc_p = bnd.BC(xlb=xl_BC_type, xrb=xr_BC_type, ylb=yl_BC_type, yrb=yr_BC_type,
                              xl_func=xl_BC, xr_func=xr_BC, yl_func=yl_BC, yr_func=yr_BC, grid=my_grid)
                lev.register_var("v", bc_p)
                # the library indexes the y-side values with GLOBAL row indices
                full = patch.Grid2d(n_t, n_t, ng=self.ng, xmin=xmin, xmax=xmax, ymin=ymin, ymax=ymax,
                                    device=my_grid.device) if info["slab"] else my_grid
                self._h.set_bc_values(_to_host_1d(bc_p.xl_value), _to_host_1d(bc_p.xr_value),
                                      _to_host_1d(yl_BC(full.x)) if yl_BC is not None else None,
                                      _to_host_1d(yr_BC(full.x)) if yr_BC is not None else None)
            else:
                lev.register_var("v", bc)
            lev.register_var("f", bc)
            lev.register_var("r", bc)
            v = self._h.plane(i, "v")
            pitch = v.stride(0)
            planes = self._h.workspace.as_strided((3, info["ni"] + 2, pitch), (info["plane_stride"], pitch, 1),
                                                  v.storage_offset())
            lev.create(planes=planes)
            self.grids.append(lev)
            if self.verbose:
                print(lev)
            n_t *= 2

        soln_grid = self.grids[self.nlevels - 1].grid
        self.ilo, self.ihi, self.jlo, self.jhi = soln_grid.ilo, soln_grid.ihi, soln_grid.jlo, soln_grid.jhi
        self.x, self.dx = soln_grid.x, soln_grid.dx
        self.y, self.dy = soln_grid.y, soln_grid.dy
        self.soln_grid = soln_grid

        self.source_norm = 0.0
        self.num_cycles = 0
        self.residual_error = 1.e33
        self.relative_error = 1.e33
        self.current_cycle = -1
        self.current_level = -1
        self.up_or_down = ""
        self.vis = 0
        self.vis_title = vis_title
        self.frame = 0
        self._old_phi = None
        self.use_graph = True       # replay the V-cycle as a CUDA graph after one eager cycle
        self.enqueue_ahead = 2      # cycles enqueued per read-back of the device-side stopping rule (solve())
        self._graph = None
        self._graph_error = None

    x2d = property(lambda self: self.soln_grid.x2d)
    y2d = property(lambda self: self.soln_grid.y2d)

    # ---- I/O of the finest level (MG.py:407-527) ------------------------------------------------
    def grid_info(self, level, indent=0):
        print(f"{indent * ' '}level: {level}, grid: {self.grids[level].grid.nx} x {self.grids[level].grid.ny}")

    def get_solution(self, grid=None):
        v = self.grids[self.nlevels - 1].get_var("v")
        if grid is None:
            return v.copy()
        myg = self.soln_grid
        assert grid.dx == myg.dx and grid.dy == myg.dy
        sol = grid.scratch_array()
        sol.v(buf=1)[:, :] = v.v(buf=1)
        return sol

    def get_solution_gradient(self, grid=None):
        myg = self.soln_grid
        og = myg if grid is None else grid
        assert og.dx == myg.dx and og.dy == myg.dy
        v = self.grids[self.nlevels - 1].get_var("v")
        gx, gy = og.scratch_array(), og.scratch_array()
        gx.v()[:, :] = 0.5 * (v.ip(1) - v.ip(-1)) / myg.dx
        gy.v()[:, :] = 0.5 * (v.jp(1) - v.jp(-1)) / myg.dy
        return gx, gy

    def get_solution_object(self):
        return self.grids[self.nlevels - 1]

    @staticmethod
    def _assign(dst, data):
        if isinstance(data, torch.Tensor):
            dst.t().copy_(data.as_subclass(torch.Tensor) if isinstance(data, ArrayIndexer) else data)
        else:
            dst[:, :] = np.asarray(data, dtype=np.float64)

    def init_solution(self, data):
        self._assign(self.grids[self.nlevels - 1].get_var("v"), data)

    def init_zeros(self):
        self.grids[self.nlevels - 1].get_var("v").t().zero_()

    def init_RHS(self, data):
        f = self.grids[self.nlevels - 1].get_var("f")
        self._assign(f, data)
        self.source_norm = self._norm(self.nlevels - 1, "f")
        if self._decomp is not None:
            self._h.exchange(self.nlevels - 1, "f", self._h.tb_halo)   # halo cells need f too
        if self.verbose:
            print("Source norm = ", self.source_norm)
        self.initialized_rhs = 1

    def set_operator(self, alpha, beta):
        """(extension) change alpha / beta of this hierarchy in place -- what constructing a new solver with
        other coefficients amounts to (the diffusion solver's beta follows dt).  A captured V-cycle graph holds
        the old coefficients as kernel arguments, so it is dropped."""
        if alpha != self.alpha or beta != self.beta:
            self.alpha, self.beta = alpha, beta
            self._h.set_operator(alpha, beta)
            self._graph = None

    def init_RHS_crank_nicolson(self, phi, coef):
        """(extension) init_RHS(phi + coef * lap(phi)) evaluated on the device from the ghost-filled plane phi
        (diffusion/simulation.py:87-93)"""
        self._h.cn_rhs(phi, coef)
        self.source_norm = self._norm(self.nlevels - 1, "f")
        if self._decomp is not None:
            self._h.exchange(self.nlevels - 1, "f", self._h.tb_halo)   # halo cells need f too (as in init_RHS)
        self.initialized_rhs = 1

    def _norm(self, level, which):
        g = self.grids[level].grid
        return math.sqrt(g.dx * g.dy * self._h.sumsq(level, which))

    # ---- overridable building blocks (MG.py:529-621) --------------------------------------------
    def _compute_residual(self, level):
        self._h.residual(level)

    def smooth(self, level, nsmooth):
        self._h.smooth(level, nsmooth)

    def _stock(self):
        cls = type(self)
        return cls.smooth is CellCenterMG2d.smooth and cls._compute_residual is CellCenterMG2d._compute_residual

    def v_cycle(self, level):
        """one V-cycle from `level` down and back (MG.py:699-778).  With the stock smoother and
        residual the whole hierarchy is traversed inside the library; subclasses that override the
        hooks get the reference's recursion with their hooks called per level."""
        if self._stock() and level == self.nlevels - 1 and not self.verbose:
            # (decomposed hierarchies too: slab levels push their halo rows from the kernels' epilogues,
            # csrc/mg_kernels.cuh "peer-memory communication"; nothing but kernels is enqueued)
            self.current_level = level
            self._h.vcycle()
            return
        if level > 0:
            self.current_level = level
            self.up_or_down = "down"
            if self.verbose:
                self._compute_residual(level)
                orig = self._norm(level, "r")
            self.smooth(level, self.nsmooth)
            self._compute_residual(level)
            if self.verbose:
                print(f"  level = {level:2}, nx = {self.grids[level].grid.nx:4}, residual change: "
                      f"{orig:11.6g} → {self._norm(level, 'r'):11.6g}")
            self._h.restrict(level)
            self.v_cycle(level - 1)
            self.current_level = level
            self.up_or_down = "up"
            self._h.prolong_correct(level)
            if self.verbose:
                self._compute_residual(level)
                orig = self._norm(level, "r")
            self.smooth(level, self.nsmooth)
            if self.verbose:
                self._compute_residual(level)
                print(f"  level = {level:2}, nx = {self.grids[level].grid.nx:4}, residual change: "
                      f"{orig:11.6g} → {self._norm(level, 'r'):11.6g}")
        else:
            if self.verbose:
                print("  bottom solve")
            self.current_level = level
            self.smooth(level, self.nsmooth_bottom)
            self._h.fill_bc(level)

    def solve(self, rtol=1.e-11):
        """V-cycles until ||r|| / ||f|| <= rtol or max_cycles (MG.py:623-697).

        The cycle body (zero the coarse v, V-cycle, bookkeeping) is ~60 kernel launches and nothing else -- also on a
        decomposed hierarchy, whose halo rows travel in the kernels' epilogues -- so after one eager cycle it is
        captured into a CUDA graph and replayed.  The stopping rule runs on the device (p2b_mg_set_stop): the host
        enqueues `enqueue_ahead` cycles at a time and reads the scalars back once per batch; once the rule has
        fired, the kernels of the cycles enqueued ahead return at once, so the solution and the cycle count are the
        reference's.  Verbose mode and subclasses that override the hooks stay eager, one read-back per cycle."""
        if not self.initialized_rhs:
            msg.fail("ERROR: RHS not initialized")
        if self.verbose:
            print("source norm = ", self.source_norm)
        fine = self.nlevels - 1
        g = self.soln_grid
        v = self.grids[fine].get_var("v").t()
        pitch = v.stride(0)
        if self._old_phi is None:       # persistent: the captured graph holds its address
            self._old_phi = torch.empty((g.qx, pitch), dtype=torch.float64, device=v.device)
        old_phi = self._old_phi
        old_phi[:, :g.qy].copy_(v)
        h = self._h
        if self._decomp is not None:
            h.exchange(fine, "v", h.tb_halo)        # the blocked smoother reads tb_halo rows of the neighbours' v

        residual_error = 1.e33
        relative_error = 1.e33
        fast = self._stock() and not self.verbose

        def body():
            h.zero_coarse()
            self.v_cycle(fine)
            h.cycle_diagnostics_enqueue(old_phi)

        if fast:
            h.set_stop(True, self.source_norm, rtol, self.max_cycles)
            use_graph = self.use_graph
            enq = 0
            while True:
                batch = 1 if enq == 0 else max(1, int(self.enqueue_ahead))
                for _ in range(batch):
                    self.current_cycle = enq + 1
                    if use_graph and self._graph is not None:
                        self._graph.replay()
                    elif use_graph and enq >= 1:
                        try:
                            torch.cuda.synchronize()
                            graph = torch.cuda.CUDAGraph()
                            with torch.cuda.graph(graph):
                                body()
                            self._graph = graph
                            graph.replay()
                        except Exception as exc:   # pylint: disable=broad-except
                            # capture is an optimisation only: fall back to eager launches
                            use_graph = self.use_graph = False
                            self._graph_error = repr(exc)
                            torch.cuda.synchronize()
                            body()
                    else:
                        body()
                    enq += 1
                relsq, rsq, residual_error, ncyc = h.result()
                if ncyc < enq or not residual_error > rtol or ncyc >= self.max_cycles:
                    break
            relative_error = math.sqrt(g.dx * g.dy * relsq)
            cycle = int(ncyc) + 1
            h.set_stop(False)
        else:
            h.set_stop(False)
            cycle = 1
            while residual_error > rtol and cycle <= self.max_cycles:
                self.current_cycle = cycle
                if self.verbose:
                    print(f"<<< beginning V-cycle (cycle {cycle}) >>>\n")
                h.zero_coarse()
                self.v_cycle(fine)
                # relative change, old_phi <- v, residual and its norm, all on the device
                if self._decomp is not None:
                    h.exchange(fine, "v", 1)
                if self._stock():
                    relsq, rsq = h.cycle_diagnostics(old_phi)
                else:
                    relsq, _ = h.cycle_diagnostics(old_phi)
                    self._compute_residual(fine)
                    rsq = h.sumsq(fine, "r")
                relative_error = math.sqrt(g.dx * g.dy * relsq)
                rnorm = math.sqrt(g.dx * g.dy * rsq)
                residual_error = rnorm / self.source_norm if self.source_norm != 0.0 else rnorm
                if self.verbose:
                    print(f"cycle {cycle}: relative err = {relative_error}, residual err = {residual_error}\n")
                cycle += 1

        self.num_cycles = cycle - 1
        self.relative_error = relative_error
        self.residual_error = residual_error
        if self._decomp is not None:
            h.exchange(fine, "v", 1)      # the slab's x "ghost" rows are the neighbours' rows
        else:
            h.fill_bc(fine)

"""Diffusion solver front end: pyro/diffusion/simulation.py (Simulation :12-104).

phi_t = k L phi advanced with Crank-Nicolson: (1 - dt k/2 L) phi^{n+1} = phi^n + dt k/2 L phi^n, solved by
CellCenterMG2d with alpha = 1, beta = dt k / 2 to rtol 1e-10.  The right-hand side is formed on the device
from the ghost-filled state (p2b_mg_cn_rhs); the reference constructs a new solver every step, here one
hierarchy is kept and its beta follows dt (p2b_mg_set_operator).  phi is bit-identical to the reference's
after every step."""
import numpy as np

from ..multigrid import MG
from ..simulation_null import NullSimulation, bc_setup, grid_setup
from ..util import msg


class Simulation(NullSimulation):
    def initialize(self):
        # decomposition (extension): this rank owns an x-slab; the solve runs on the x-slab multigrid
        my_grid = grid_setup(self.rp, ng=1, decomposition=self.decomposition)
        if my_grid.nx_global != my_grid.ny:
            msg.fail("need nx = ny for diffusion problems")
        n = int(np.log(my_grid.nx_global) / np.log(2.0))
        if 2 ** n != my_grid.nx_global and my_grid.nx_global & (my_grid.nx_global - 1):
            msg.fail("grid needs to be a power of 2")
        bc, _, _ = bc_setup(self.rp)
        for b in (bc.xlb, bc.xrb, bc.ylb, bc.yrb):
            if b not in ("periodic", "neumann", "dirichlet"):
                msg.fail("invalid BC")
        my_data = self.data_class(my_grid)
        my_data.register_var("phi", bc)
        my_data.create()
        my_data.decomposition = self.decomposition
        self.cc_data = my_data
        self._mg = None
        self.problem_func(self.cc_data, self.rp)

    def method_compute_timestep(self):
        """cfl times the explicit limit min(dx^2, dy^2) / k (diffusion/simulation.py:46-60); the implicit
        update does not need cfl < 1"""
        cfl = self.rp.get_param("driver.cfl")
        k = self.rp.get_param("diffusion.k")
        g = self.cc_data.grid
        self.dt = cfl * min(g.dx ** 2 / k, g.dy ** 2 / k)

    def evolve(self):
        self.cc_data.fill_BC_all()
        g = self.cc_data.grid
        phi = self.cc_data.planes[self.cc_data.names.index("phi")][:, :g.qy]
        k = self.rp.get_param("diffusion.k")
        b = self.cc_data.BCs["phi"]
        beta = 0.5 * self.dt * k
        if self._mg is None:
            split = {}
            if self.decomposition is not None and self.decomposition.size > 1:
                split = {"decomposition": self.decomposition, "split_n": self.rp.get_param("diffusion.mg_split_n")}
            self._mg = MG.CellCenterMG2d(g.nx_global, g.ny, xmin=g.xmin, xmax=g.xmax, ymin=g.ymin, ymax=g.ymax,
                                         xl_BC_type=b.xlb, xr_BC_type=b.xrb, yl_BC_type=b.ylb, yr_BC_type=b.yrb,
                                         alpha=1.0, beta=beta, verbose=0, **split)
            if self._mg.soln_grid.nx != g.nx:
                msg.fail("ERROR: the finest multigrid level must be split like the solver grid (lower diffusion.mg_split_n)")
        mg = self._mg
        mg.set_operator(1.0, beta)
        mg.init_RHS_crank_nicolson(phi, beta)
        mg.init_zeros()
        mg.solve(rtol=1.e-10)
        soln = mg.grids[mg.nlevels - 1].get_var("v").t()
        phi[g.ilo:g.ihi + 1, g.jlo:g.jhi + 1].copy_(soln[1:-1, 1:-1])
        self.cc_data.version += 1
        self.cc_data.t += self.dt
        self.n += 1

"""Incompressible solver front end: pyro/incompressible/simulation.py (Simulation :15-483).

What maps to what
  Simulation.initialize      simulation.py:17-65    grid (ng = 4), the six variables and their BCs
  Simulation.preevolve       simulation.py:67-157   initial projection, one throw-away step for gradp
  Simulation.evolve          simulation.py:159-404  the stages below, in the reference's order
     reconstruction.limit x4, incomp_interface.mac_vels   -> FlowHandle.interface_states / mac_vels
     MAC projection (CellCenterMG2d, rtol 1e-12)           -> multigrid.MG.CellCenterMG2d.solve
     incomp_interface.states                               -> FlowHandle.upwind_states
     advective update, fill_BC                             -> FlowHandle.advect_update, ops.fill_ghost
     final projection (initial guess phi, rtol 1e-12)      -> CellCenterMG2d.solve, FlowHandle.project
  method_compute_timestep    burgers/simulation.py:41-58 (inherited)

mac_vels and states build the same eight corrected interface states from the same inputs
(incomp_interface.py:38-62 and :105-129); they are computed once per step here and kept on the device.
The reference constructs a fresh CellCenterMG2d for every projection; a solver with the same boundary
types is reused instead (its coarse levels are re-zeroed by solve(), the finest level is overwritten
by init_zeros / init_solution / init_RHS), which keeps the captured V-cycle graph alive across steps.
Every state plane is bit-identical to the reference's after every step.
"""
from ..burgers.simulation import Simulation as burgers_simulation
from ..burgers.simulation import _no_particles
from ..mesh import boundary as bnd
from ..multigrid import MG
from ..simulation_null import bc_setup, grid_setup
from ..util import msg


class Simulation(burgers_simulation):
    def initialize(self, *, other_bc=False, aux_vars=()):
        # decomposition (extension): this rank owns an x-slab with ng = 4 halo rows; every fill_BC then exchanges the
        # halo rows first, the projections run on the x-slab multigrid (bit-identical to the single-domain solver)
        # and the time step is all-reduced -- the stage kernels themselves are unchanged
        my_grid = grid_setup(self.rp, ng=4, decomposition=self.decomposition)
        my_data = self.data_class(my_grid)
        if other_bc:
            self.define_other_bc()
        bc, bc_xodd, bc_yodd = bc_setup(self.rp)
        my_data.register_var("x-velocity", bc_xodd)
        my_data.register_var("y-velocity", bc_yodd)
        # phi (the projections' unknown) is periodic with the velocities, Neumann where they are
        # Dirichlet; the reference assumes all-periodic or all-Dirichlet boundaries (simulation.py:38-46)
        if bc.xlb == "periodic":
            phi_bc = bc
        elif bc.xlb == "dirichlet":
            phi_bc = bnd.BC(xlb="neumann", xrb="neumann", ylb="neumann", yrb="neumann")
        else:
            msg.fail("ERROR: the incompressible solver needs periodic or dirichlet boundaries")
        for name in ("phi-MAC", "phi", "gradp_x", "gradp_y"):
            my_data.register_var(name, phi_bc)
        for keyword, value in aux_vars:
            my_data.set_aux(keyword=keyword, value=value)
        my_data.create()
        my_data.decomposition = self.decomposition
        self.cc_data = my_data
        _no_particles(self.rp)
        self._make_flow()
        self._mg = {}
        self.in_preevolve = False
        self.problem_func(self.cc_data, self.rp)

    # ---- helpers ---------------------------------------------------------------------------------------
    def _planes(self):
        g = self.cc_data.grid
        names = self.cc_data.names
        return {n: self.cc_data.planes[names.index(n)][:, :g.qy] for n in names}

    def _fill_velocities(self):
        self.cc_data.fill_BC("x-velocity")
        self.cc_data.fill_BC("y-velocity")

    def _solver(self, bc_types):
        """CellCenterMG2d on the solver's domain with the given boundary types (cached)"""
        key = tuple(bc_types)
        if key not in self._mg:
            g = self.cc_data.grid
            split = {}
            if self.decomposition is not None and self.decomposition.size > 1:
                split = {"decomposition": self.decomposition, "split_n": self.rp.get_param("incompressible.mg_split_n")}
            mg = MG.CellCenterMG2d(g.nx_global, g.ny, xl_BC_type=key[0], xr_BC_type=key[1], yl_BC_type=key[2],
                                   yr_BC_type=key[3], xmin=g.xmin, xmax=g.xmax, ymin=g.ymin, ymax=g.ymax, verbose=0,
                                   **split)
            if mg.soln_grid.nx != g.nx:
                msg.fail("ERROR: the finest multigrid level must be split like the solver grid (lower incompressible.mg_split_n)")
            self._mg[key] = (mg, mg.soln_grid.scratch_array())
        return self._mg[key]

    def _buf1(self, plane):
        """the buf = 1 region of a solver-grid plane: the cells a multigrid-grid array covers"""
        g = self.cc_data.grid
        return plane[g.ilo - 1:g.ihi + 2, g.jlo - 1:g.jhi + 2]

    # ---- preevolve (simulation.py:67-157) ----------------------------------------------------------------
    def preevolve(self):
        self.in_preevolve = True
        P = self._planes()
        u, v, phi = P["x-velocity"], P["y-velocity"], P["phi"]
        self._fill_velocities()
        # initial projection: L phi = D U with periodic boundaries (hard-wired in the reference), U -= G phi
        mg, divU = self._solver(("periodic",) * 4)
        self._flow.cc_divergence(u, v, divU.t())
        mg.init_zeros()
        mg.init_RHS(divU)
        mg.solve(rtol=1.e-10)
        phi.zero_()
        self._buf1(phi).copy_(mg.grids[mg.nlevels - 1].get_var("v").t())
        self._flow.project(phi, u, v, None, None, 1.0, 0)
        self._fill_velocities()
        # one step from here only to obtain the lagged pressure gradient; everything else is rolled back
        saved = self.cc_data.planes.clone()
        self.method_compute_timestep()
        self.evolve()
        names = self.cc_data.names
        for name in ("gradp_x", "gradp_y"):
            saved[names.index(name)].copy_(self.cc_data.planes[names.index(name)])
        self.cc_data.planes.copy_(saved)
        self.cc_data.version += 1
        if self.verbose > 0:
            print("done with the pre-evolution")
        self.in_preevolve = False

    # ---- one step (simulation.py:159-404) --------------------------------------------------------------------
    def evolve(self, other_update_velocity=False, other_source_term=False):
        if other_update_velocity or other_source_term:
            raise NotImplementedError("viscous / source-term variants are not provided by the device build")
        P = self._planes()
        u, v, phi_MAC, phi = P["x-velocity"], P["y-velocity"], P["phi-MAC"], P["phi"]
        gradp_x, gradp_y = P["gradp_x"], P["gradp_y"]
        flow, dt = self._flow, self.dt
        limiter = self.rp.get_param("incompressible.limiter")
        proj_type = self.rp.get_param("incompressible.proj_type")
        b = self.cc_data.BCs["phi"]
        mg, divU = self._solver((b.xlb, b.xrb, b.ylb, b.yrb))
        soln = mg.grids[mg.nlevels - 1].get_var("v").t()

        if self.verbose > 0:
            print("  making MAC velocities")
        flow.interface_states(u, v, gradp_x, gradp_y, dt, limiter)
        flow.mac_vels()

        if self.verbose > 0:
            print("  MAC projection")
        flow.mac_divergence(divU.t())
        mg.init_zeros()
        mg.init_RHS(divU)
        mg.solve(rtol=1.e-12)
        self._buf1(phi_MAC).copy_(soln)
        flow.mac_project(phi_MAC)

        if self.verbose > 0:
            print("  making u, v edge states")
        flow.upwind_states()
        if self.verbose > 0:
            print("  doing provisional update of u, v")
        flow.advect_update(u, v, gradp_x, gradp_y, dt, proj_type)
        self._fill_velocities()

        if self.verbose > 0:
            print("  final projection")
        flow.cc_divergence(u, v, divU.t(), dt=dt, divide=True)
        mg.init_RHS(divU)
        mg.init_solution(self._buf1(phi))
        mg.solve(rtol=1.e-12)
        phi.zero_()
        self._buf1(phi).copy_(soln)
        flow.project(phi, u, v, gradp_x, gradp_y, dt, proj_type)
        self._fill_velocities()

        self.cc_data.version += 1
        if not self.in_preevolve:
            self.cc_data.t += self.dt
            self.n += 1

    def define_other_bc(self):
        """hook for subclasses with user-defined BCs (incompressible_viscous in the reference)"""

    def other_source_term(self):
        return None, None

    def do_other_update_velocity(self, U_MAC, U_INT):
        """hook for subclasses that change the velocity update"""

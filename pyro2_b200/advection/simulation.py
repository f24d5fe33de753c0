"""Advection solver front end: pyro/advection/simulation.py (Simulation :12-92) with
interface.linear_interface + advective_fluxes.unsplit_fluxes + the flux differencing executed by
p2b_flow_advection_update (csrc/flow.cu); a_t + u a_x + v a_y = 0 for constant (u, v)."""
from ..burgers.simulation import _no_particles
from ..flow_handle import FlowHandle
from ..simulation_null import NullSimulation, bc_setup, grid_setup


class Simulation(NullSimulation):
    def initialize(self):
        # decomposition (extension, as in the compressible solver): this rank owns an x-slab; the update's stencil
        # fits in the ng = 4 halo rows that fill_BC_all exchanges, and dt is analytic, so nothing else changes
        my_grid = grid_setup(self.rp, ng=4, decomposition=self.decomposition)
        my_data = self.data_class(my_grid)
        bc = bc_setup(self.rp)[0]
        my_data.register_var("density", bc)
        my_data.create()
        my_data.decomposition = self.decomposition
        self.cc_data = my_data
        _no_particles(self.rp)
        self._flow = FlowHandle(my_data.planes, my_grid)
        self.problem_func(self.cc_data, self.rp)

    def method_compute_timestep(self):
        """dt = cfl * min(dx / |u|, dy / |v|) (advection/simulation.py:41-54)"""
        cfl = self.rp.get_param("driver.cfl")
        u = self.rp.get_param("advection.u")
        v = self.rp.get_param("advection.v")
        xtmp = self.cc_data.grid.dx / max(abs(u), self.SMALL)
        ytmp = self.cc_data.grid.dy / max(abs(v), self.SMALL)
        self.dt = cfl * min(xtmp, ytmp)

    def evolve(self):
        g = self.cc_data.grid
        dens = self.cc_data.planes[self.cc_data.names.index("density")][:, :g.qy]
        self._flow.advection_update(dens, self.rp.get_param("advection.u"), self.rp.get_param("advection.v"),
                                    self.dt, self.rp.get_param("advection.limiter"))
        self.cc_data.version += 1
        self.cc_data.t += self.dt
        self.n += 1

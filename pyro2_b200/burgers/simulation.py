"""Burgers solver front end: pyro/burgers/simulation.py (Simulation :12-131) with the interface-state
construction, Riemann / upwinding, flux differencing and the CFL reduction executed by the p2b_flow_*
kernels.  u_t + u u_x + v u_y = 0, v_t + u v_x + v v_y = 0."""
import torch

from ..flow_handle import FlowHandle
from ..simulation_null import NullSimulation, bc_setup, grid_setup
from ..util import msg


def _no_particles(rp):
    try:
        if rp.get_param("particles.do_particles") == 1:
            msg.fail("ERROR: tracer particles are not supported by the device build")
    except KeyError:
        pass


class Simulation(NullSimulation):
    def _make_flow(self):
        self._flow = FlowHandle(self.cc_data.planes, self.cc_data.grid)

    def initialize(self):
        """grid, the two velocities, their BCs, problem initial conditions (burgers/simulation.py:14-39)"""
        # decomposition (extension, as in the compressible solver): this rank owns an x-slab with ng = 4 halo rows
        my_grid = grid_setup(self.rp, ng=4, decomposition=self.decomposition)
        my_data = self.data_class(my_grid)
        bc = bc_setup(self.rp)[0]
        my_data.register_var("x-velocity", bc)
        my_data.register_var("y-velocity", bc)
        my_data.create()
        my_data.decomposition = self.decomposition
        self.cc_data = my_data
        _no_particles(self.rp)
        self._make_flow()
        self.problem_func(self.cc_data, self.rp)

    def _velocities(self):
        names = self.cc_data.names
        return self.cc_data.planes[names.index("x-velocity")], self.cc_data.planes[names.index("y-velocity")]

    def method_compute_timestep(self):
        """dt = cfl * min(dx / max|u|, dy / max|v|), maxima over the whole arrays including ghost cells
        (burgers/simulation.py:41-58)"""
        cfl = self.rp.get_param("driver.cfl")
        g = self.cc_data.grid
        u, v = self._velocities()
        umax, vmax = self._flow.maxabs(u[:, :g.qy], v[:, :g.qy])
        if self.decomposition is not None and self.decomposition.size > 1:
            # the slabs with their halo rows and physical ghost cells cover the global array exactly
            w = torch.tensor([umax, vmax], dtype=torch.float64, device=u.device)
            umax, vmax = self.decomposition.allreduce_max_(w).tolist()
        xtmp = g.dx / max(umax, self.SMALL)
        ytmp = g.dy / max(vmax, self.SMALL)
        self.dt = cfl * min(xtmp, ytmp)

    def evolve(self):
        """one step (burgers/simulation.py:66-131)"""
        g = self.cc_data.grid
        u, v = (p[:, :g.qy] for p in self._velocities())
        limiter = self.rp.get_param("advection.limiter")
        self._flow.interface_states(u, v, None, None, self.dt, limiter)
        self._flow.mac_vels()
        self._flow.burgers_update(u, v, self.dt)
        self.cc_data.version += 1
        self.cc_data.t += self.dt
        self.n += 1

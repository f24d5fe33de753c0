"""lm_atm solver front end: pyro/lm_atm/simulation.py (Basestate :12-35, Simulation :37-691).

What maps to what
  Simulation.initialize               :51-133   grid (ng = 4), eight variables and their BCs, base state
                                                (rho0, p0 from the problem; beta0 = p0**(1/gamma) and its edge values)
  Simulation.method_compute_timestep  :138-178  -> LmHandle.reduce (advective and buoyancy limits)
  Simulation.preevolve                :180-284  initial projection, one throw-away step for the lagged gradp
  Simulation.evolve                   :286-618  the numba routines of LM_atm_interface.py and every array
                                                expression in between -> LmHandle stage calls; the MAC and the final
                                                projection -> VarCoeffCCMG2d.solve (rtol 1e-12)

The reference constructs a new VarCoeffCCMG2d for every projection; one solver per boundary signature is kept
and its coefficients are replaced in place (VarCoeffCCMG2d.set_coeffs).  The auxiliary arrays "coeff" and
"source_y" live in the handle's scratch planes and are ghost-filled where the reference calls aux_data.fill_BC.
All eight state planes are bit-identical to the reference's after every step.
"""
import numpy as np
import torch

from ..lm_handle import LmHandle
from ..mesh import boundary as bnd
from ..multigrid import variable_coeff_MG as vcMG
from ..simulation_null import NullSimulation, bc_setup, grid_setup
from ..util import msg


class Basestate:
    """a 1-d (vertical) base-state array with ghost cells, kept on the host (simulation.py:12-35)"""

    def __init__(self, ny, *, ng=0):
        self.ny, self.ng = ny, ng
        self.qy = ny + 2 * ng
        self.d = np.zeros(self.qy, dtype=np.float64)
        self.jlo, self.jhi = ng, ng + ny - 1

    def v(self, buf=0):
        return self.d[self.jlo - buf:self.jhi + 1 + buf]

    def v2d(self, buf=0):
        return self.d[np.newaxis, self.jlo - buf:self.jhi + 1 + buf]

    def v2dp(self, shift, buf=0):
        return self.d[np.newaxis, self.jlo + shift - buf:self.jhi + 1 + shift + buf]

    def jp(self, shift, buf=0):
        return self.d[self.jlo - buf + shift:self.jhi + 1 + buf + shift]


class Simulation(NullSimulation):
    def __init__(self, solver_name, problem_name, problem_func, rp, *, problem_finalize_func=None,
                 problem_source_func=None, timers=None):
        super().__init__(solver_name, problem_name, problem_func, rp, problem_finalize_func=problem_finalize_func,
                         problem_source_func=problem_source_func, timers=timers)
        self.base = {}
        self.in_preevolve = False

    def initialize(self):
        if self.decomposition is not None and self.decomposition.size > 1:
            msg.fail("ERROR: lm_atm runs on one GPU (the variable-coefficient multigrid is not decomposed)")
        myg = grid_setup(self.rp, ng=4)
        bc_dens, bc_xodd, bc_yodd = bc_setup(self.rp)
        my_data = self.data_class(myg)
        my_data.register_var("density", bc_dens)
        my_data.register_var("x-velocity", bc_xodd)
        my_data.register_var("y-velocity", bc_yodd)
        my_data.register_var("eint", bc_dens)       # not evolved: carried for output and comparisons
        # phi: periodic with the state, Neumann at walls, Dirichlet at outflow (simulation.py:75-92)
        bcs = []
        for b in (self.rp.get_param("mesh.xlboundary"), self.rp.get_param("mesh.xrboundary"),
                  self.rp.get_param("mesh.ylboundary"), self.rp.get_param("mesh.yrboundary")):
            kind = {"periodic": "periodic", "reflect": "neumann", "slipwall": "neumann", "outflow": "dirichlet"}.get(b)
            if kind is None:
                msg.fail(f"ERROR: lm_atm cannot derive a boundary condition for phi from {b}")
            bcs.append(kind)
        bc_phi = bnd.BC(xlb=bcs[0], xrb=bcs[1], ylb=bcs[2], yrb=bcs[3])
        my_data.register_var("phi-MAC", bc_phi)
        my_data.register_var("phi", bc_phi)
        my_data.register_var("gradp_x", bc_dens)
        my_data.register_var("gradp_y", bc_dens)
        my_data.create()
        self.cc_data = my_data
        self._bc_dens, self._bc_yodd, self._bc_phi = bc_dens, bc_yodd, bc_phi

        # the base state: rho0, p0 from the problem setup, beta0 = p0**(1/gamma) and its edge-centred values
        self.base["rho0"] = Basestate(myg.ny, ng=myg.ng)
        self.base["p0"] = Basestate(myg.ny, ng=myg.ng)
        self.problem_func(self.cc_data, self.base, self.rp)
        gamma = self.rp.get_param("eos.gamma")
        self.base["beta0"] = Basestate(myg.ny, ng=myg.ng)
        self.base["beta0"].d[:] = self.base["p0"].d ** (1.0 / gamma)
        edges = Basestate(myg.ny, ng=myg.ng)
        edges.jp(1)[:] = 0.5 * (self.base["beta0"].v() + self.base["beta0"].jp(1))
        edges.d[myg.jlo] = self.base["beta0"].d[myg.jlo]
        edges.d[myg.jhi + 1] = self.base["beta0"].d[myg.jhi]
        self.base["beta0-edges"] = edges
        dev = np.stack([self.base[k].d for k in ("rho0", "p0", "beta0", "beta0-edges")])
        self._base_dev = torch.from_numpy(np.ascontiguousarray(dev)).to(my_data.planes.device)
        self._lm = LmHandle(my_data.planes, myg, self._base_dev)
        self._mg = None
        self._divU = None

    def make_prime(self, a, a0):
        return a - torch.from_numpy(a0.v2d(buf=a0.ng)).to(a.device)

    # ---- helpers ----------------------------------------------------------------------------------------------
    def _planes(self):
        g = self.cc_data.grid
        return {n: self.cc_data.planes[k][:, :g.qy] for k, n in enumerate(self.cc_data.names)}

    def _buf1(self, plane):
        g = self.cc_data.grid
        return plane[g.ilo - 1:g.ihi + 2, g.jlo - 1:g.jhi + 2]

    def _solver(self):
        """the variable-coefficient solver on the solver's domain with phi's boundary types; its coefficients are
        replaced before every projection (the MAC and final projections share the BCs, simulation.py:75-96)"""
        g = self.cc_data.grid
        coeff = self._buf1(self._lm.plane(LmHandle.COEFF)[:, :g.qy])
        if self._mg is None:
            b = self._bc_phi
            self._mg = vcMG.VarCoeffCCMG2d(g.nx, g.ny, xl_BC_type=b.xlb, xr_BC_type=b.xrb, yl_BC_type=b.ylb,
                                           yr_BC_type=b.yrb, xmin=g.xmin, xmax=g.xmax, ymin=g.ymin, ymax=g.ymax,
                                           coeffs=coeff, coeffs_bc=self._bc_dens, verbose=0)
            self._divU = self._mg.soln_grid.scratch_array()
        else:
            self._mg.set_coeffs(coeff)
        return self._mg, self._divU

    def _fill_aux(self, plane, bc):
        self._lm.fill(plane, bc.names())

    # ---- timestep (simulation.py:138-178) ---------------------------------------------------------------------
    def method_compute_timestep(self):
        g = self.cc_data.grid
        cfl = self.rp.get_param("driver.cfl")
        P = self._planes()
        grav = self.rp.get_param("lm-atmosphere.grav")
        uall, vall, uval, vval, fbuoy = self._lm.reduce(P["density"], P["x-velocity"], P["y-velocity"], grav)
        xtmp = ytmp = 1.e33
        if not uall == 0:
            xtmp = g.dx / uval
        if not vall == 0:
            ytmp = g.dy / vval
        dt = cfl * min(xtmp, ytmp)
        # the buoyancy limit: F_buoy = max(|rho' g| / rho) over the valid cells
        with np.errstate(divide="ignore"):
            dt_buoy = float(np.sqrt(np.float64(2.0 * g.dx) / np.float64(fbuoy)))
        self.dt = min(dt, dt_buoy)
        if self.verbose > 0:
            print(f"timestep is {dt}")

    # ---- preevolve (simulation.py:180-284) -------------------------------------------------------------------
    def preevolve(self):
        self.in_preevolve = True
        lm = self._lm
        P = self._planes()
        rho, u, v, phi = P["density"], P["x-velocity"], P["y-velocity"], P["phi"]
        for name in ("density", "x-velocity", "y-velocity"):
            self.cc_data.fill_BC(name)
        # initial projection: L_coeff phi = D(beta0 U) with coeff = beta0^2 / rho, U -= (beta0 / rho) G phi
        lm.coeff(rho, None, 1.0, True, 0)
        mg, divU = self._solver()
        lm.cc_divergence(u, v, divU.t())
        mg.init_zeros()
        mg.init_RHS(divU)
        mg.solve(rtol=1.e-10)
        phi.zero_()
        self._buf1(phi).copy_(mg.grids[mg.nlevels - 1].get_var("v").t())
        lm.project(rho, phi, u, v, None, None, 1.0, 0)
        self.cc_data.fill_BC("x-velocity")
        self.cc_data.fill_BC("y-velocity")
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

    # ---- one step (simulation.py:286-618) --------------------------------------------------------------------
    def evolve(self):
        lm, dt = self._lm, self.dt
        P = self._planes()
        rho, u, v, eint = P["density"], P["x-velocity"], P["y-velocity"], P["eint"]
        phi_MAC, phi, gradp_x, gradp_y = P["phi-MAC"], P["phi"], P["gradp_x"], P["gradp_y"]
        limiter = self.rp.get_param("lm-atmosphere.limiter")
        proj_type = self.rp.get_param("lm-atmosphere.proj_type")
        grav = self.rp.get_param("lm-atmosphere.grav")
        gamma = self.rp.get_param("eos.gamma")
        g = self.cc_data.grid
        rho_old = lm.plane(LmHandle.RHO_OLD)[:, :g.qy]

        if self.verbose > 0:
            print("  making MAC velocities")
        lm.coeff(rho, None, 1.0, False, 0)                 # coeff = beta0 / rho
        self._fill_aux(LmHandle.COEFF, self._bc_dens)
        lm.source(rho, None, grav)                         # source = rho' g / rho
        self._fill_aux(LmHandle.SOURCE, self._bc_yodd)
        lm.interface_states(u, v, gradp_x, gradp_y, dt, limiter)
        lm.mac_vels()

        if self.verbose > 0:
            print("  MAC projection")
        lm.coeff(rho, None, 1.0, True, 1)                  # coeff.v(buf=1) = beta0^2 / rho
        mg, divU = self._solver()
        soln = mg.grids[mg.nlevels - 1].get_var("v").t()
        lm.mac_divergence(divU.t())
        mg.init_zeros()
        mg.init_RHS(divU)
        mg.solve(rtol=1.e-12)
        phi_MAC.zero_()
        self._buf1(phi_MAC).copy_(soln)
        lm.coeff(rho, None, 1.0, False, 0)
        self._fill_aux(LmHandle.COEFF, self._bc_dens)
        lm.mac_project(phi_MAC)

        # density: predict to the faces with the MAC velocities, conservative update, eint from the base pressure
        lm.density_update(rho, eint, dt, limiter, gamma)
        self.cc_data.fill_BC("density")

        if self.verbose > 0:
            print("  making u, v edge states")
        lm.coeff(rho, rho_old, 2.0, False, 0)              # coeff = beta0 * 2 / (rho + rho_old)
        self._fill_aux(LmHandle.COEFF, self._bc_dens)
        lm.interface_states(u, v, gradp_x, gradp_y, dt, limiter)
        lm.upwind_states()
        if self.verbose > 0:
            print("  doing provisional update of u, v")
        lm.advect_update(u, v, gradp_x, gradp_y, dt, proj_type)
        lm.source(rho, rho_old, grav)                      # time-centred buoyancy over the whole array
        self._fill_aux(LmHandle.SOURCE, self._bc_yodd)
        lm.add_source(v, dt)
        self.cc_data.fill_BC("x-velocity")
        self.cc_data.fill_BC("y-velocity")

        if self.verbose > 0:
            print("  final projection")
        lm.coeff(rho, None, 1.0, True, 0)
        mg, divU = self._solver()
        lm.cc_divergence(u, v, divU.t(), dt=dt, divide=True)
        mg.init_RHS(divU)
        mg.init_solution(self._buf1(phi))
        mg.solve(rtol=1.e-12)
        phi.zero_()
        self._buf1(phi).copy_(soln)
        lm.project(rho, phi, u, v, gradp_x, gradp_y, dt, proj_type)
        for name in ("x-velocity", "y-velocity", "gradp_x", "gradp_y"):
            self.cc_data.fill_BC(name)

        self.cc_data.version += 1
        if not self.in_preevolve:
            self.cc_data.t += self.dt
            self.n += 1

    def write_extras(self, f):
        """the 1-d base state goes into the snapshot with the 2-d fields (simulation.py:670-681)"""
        gb = f.create_group("base state")
        for name, state in self.base.items():
            gb.create_dataset(name, data=state.d)

    def read_extras(self, f):
        """restore the base state of a snapshot (simulation.py:683-691)"""
        gb = f["base state"]
        g = self.cc_data.grid
        for name in gb:
            self.base[name] = Basestate(g.ny, ng=g.ng)
            self.base[name].d[:] = np.asarray(gb[name])

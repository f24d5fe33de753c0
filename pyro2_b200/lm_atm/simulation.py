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


class _Plain:
    """a bare device array with the one ArrayIndexer method the divergence kernels' callers use"""

    def __init__(self, tensor):
        self._t = tensor

    def t(self):
        return self._t


class Simulation(NullSimulation):
    def __init__(self, solver_name, problem_name, problem_func, rp, *, problem_finalize_func=None,
                 problem_source_func=None, timers=None):
        super().__init__(solver_name, problem_name, problem_func, rp, problem_finalize_func=problem_finalize_func,
                         problem_source_func=problem_source_func, timers=timers)
        self.base = {}
        self.in_preevolve = False

    def initialize(self):
        # decomposition (extension): x-slabs for the explicit stages; the two variable-coefficient projections of a
        # step are solved REPLICATED -- coefficients, right-hand side and initial guess are all-gathered and every rank
        # runs the single-domain solver (the variable-coefficient hierarchy itself is not decomposed), then keeps its
        # slab of the solution.  Bit-identical to the single-domain run; periodic in x only.
        self._decomposed = self.decomposition is not None and self.decomposition.size > 1
        if self._decomposed and self.rp.get_param("mesh.xlboundary") != "periodic":
            msg.fail("ERROR: a decomposed lm_atm run needs periodic x boundaries")
        myg = grid_setup(self.rp, ng=4, decomposition=self.decomposition)
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
        my_data.decomposition = self.decomposition
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

    def _gather_rows(self, local):
        """slab runs: the (nx_global + 2, ny + 2) array of the whole domain from every rank's (nx + 2, ny + 2) buf-1
        array -- the owned rows of each slab, the low ghost row of the first and the high ghost row of the last"""
        import torch.distributed as dist   # pylint: disable=import-outside-toplevel
        d, g = self.decomposition, self.cc_data.grid
        mine = local.contiguous()
        parts = torch.empty((d.size,) + tuple(mine.shape), dtype=mine.dtype, device=mine.device)
        dist.all_gather_into_tensor(parts.view(-1), mine.reshape(-1), group=d.group)
        full = torch.empty((g.nx_global + 2, mine.shape[1]), dtype=mine.dtype, device=mine.device)
        full[1:-1] = parts[:, 1:-1].reshape(g.nx_global, mine.shape[1])
        full[0], full[-1] = parts[0, 0], parts[-1, -1]
        return full

    def _my_rows(self, full):
        """this slab's buf-1 rows of a whole-domain (nx_global + 2, ny + 2) array"""
        g = self.cc_data.grid
        return full[g.ioffset:g.ioffset + g.nx + 2]

    def _solver(self):
        """the variable-coefficient solver on the solver's domain with phi's boundary types; its coefficients are
        replaced before every projection (the MAC and final projections share the BCs, simulation.py:75-96).
        Returns (solver, the array the divergence is written into); slab runs: the solver covers the whole domain and
        the divergence array is this slab's part, handed to the solver by _set_rhs"""
        g = self.cc_data.grid
        coeff = self._buf1(self._lm.plane(LmHandle.COEFF)[:, :g.qy])
        if self._decomposed:
            coeff = self._gather_rows(coeff)
        if self._mg is None:
            b = self._bc_phi
            self._mg = vcMG.VarCoeffCCMG2d(g.nx_global, g.ny, xl_BC_type=b.xlb, xr_BC_type=b.xrb, yl_BC_type=b.ylb,
                                           yr_BC_type=b.yrb, xmin=g.xmin, xmax=g.xmax, ymin=g.ymin, ymax=g.ymax,
                                           coeffs=coeff, coeffs_bc=self._bc_dens, verbose=0)
            self._divU = self._mg.soln_grid.scratch_array()
            if self._decomposed:
                self._divU = _Plain(torch.zeros((g.nx + 2, self._divU.shape[1]), dtype=torch.float64,
                                                device=self.cc_data.planes.device))
        else:
            self._mg.set_coeffs(coeff)
        return self._mg, self._divU

    def _set_rhs(self, mg, divU):
        mg.init_RHS(self._gather_rows(divU.t()) if self._decomposed else divU)

    def _solution(self, mg):
        """the buf-1 part of the solution this rank's phi planes take"""
        soln = mg.grids[mg.nlevels - 1].get_var("v").t()
        return self._my_rows(soln) if self._decomposed else soln

    def _fill_aux(self, plane, bc):
        if self._decomposed:
            # periodic x: both sides face another slab -- their rows are exchanged, the y sides filled as usual
            g = self.cc_data.grid
            self.decomposition.exchange(self._lm.plane(plane).unsqueeze(0), g.nx, g.ng, periodic=True)
            names = bc.names()
            self._lm.fill(plane, (None, None, names[2], names[3]))
            return
        self._lm.fill(plane, bc.names())

    # ---- timestep (simulation.py:138-178) ---------------------------------------------------------------------
    def method_compute_timestep(self):
        g = self.cc_data.grid
        cfl = self.rp.get_param("driver.cfl")
        P = self._planes()
        grav = self.rp.get_param("lm-atmosphere.grav")
        red = self._lm.reduce(P["density"], P["x-velocity"], P["y-velocity"], grav)
        if self._decomposed:
            red = self.decomposition.allreduce_max_(torch.tensor(red, dtype=torch.float64,
                                                                 device=self.cc_data.planes.device)).tolist()
        uall, vall, uval, vval, fbuoy = red
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
        self._set_rhs(mg, divU)
        mg.solve(rtol=1.e-10)
        phi.zero_()
        self._buf1(phi).copy_(self._solution(mg))
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
        lm.mac_divergence(divU.t())
        mg.init_zeros()
        self._set_rhs(mg, divU)
        mg.solve(rtol=1.e-12)
        phi_MAC.zero_()
        self._buf1(phi_MAC).copy_(self._solution(mg))
        lm.coeff(rho, None, 1.0, False, 0)
        self._fill_aux(LmHandle.COEFF, self._bc_dens)
        lm.mac_project(phi_MAC)
        if self._decomposed:
            # The density prediction in a cell next to a slab boundary upwinds against the neighbour cell's state, which
            # is traced with the MAC velocities on THAT cell's faces: projected interior faces in the single-domain run.
            # The halo faces here were not projected (the projection covers the owned faces), so take the neighbour's.
            # Only across interior slab boundaries: at the domain's periodic seam the single-domain run itself uses the
            # unprojected ghost-face velocities.
            for k in (LmHandle.U_MAC, LmHandle.V_MAC):
                self.decomposition.exchange(lm.plane(k).unsqueeze(0), g.nx, g.ng, periodic=False)

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
        self._set_rhs(mg, divU)
        mg.init_solution(self._gather_rows(self._buf1(phi)) if self._decomposed else self._buf1(phi))
        mg.solve(rtol=1.e-12)
        phi.zero_()
        self._buf1(phi).copy_(self._solution(mg))
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

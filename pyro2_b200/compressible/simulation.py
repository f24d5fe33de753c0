"""Compressible solver front end: the reference's ``Simulation`` interface
(pyro/compressible/simulation.py) with the time step executed by one fused CUDA kernel.

What maps to what
  Simulation.initialize               simulation.py:193-265  (grid, variables, BCs, problem init)
  Simulation.method_compute_timestep  simulation.py:267-288  -> p2b_cfl_wavemax, or the maxima the
                                                               previous sweep already accumulated
  Simulation.evolve                   simulation.py:290-450  -> p2b_compressible_sweep
  Simulation.clean_state              simulation.py:452-456
  cons_to_prim / prim_to_cons         simulation.py:49-102   (torch, for users / tests / output)
  Variables                           simulation.py:12-46

Scope: Cartesian grid, HLLC or CGF, constant gravity, heating
sources dens * e_rate * profile, the sponge, the hse and ambient boundaries; no particles / ramp boundary.
Anything else raises instead of silently taking another path.
"""
import torch

from .. import ops
from ..mesh import boundary as bnd
from ..mesh import patch
from ..simulation_null import NullSimulation, bc_setup, grid_setup
from ..util import msg
from . import BC, derives, eos


class Variables:
    """integer keys of the conserved / primitive variables (simulation.py:12-46)"""

    def __init__(self, myd):
        self.nvar = len(myd.names)
        self.idens = myd.names.index("density")
        self.ixmom = myd.names.index("x-momentum")
        self.iymom = myd.names.index("y-momentum")
        self.iener = myd.names.index("energy")
        self.naux = self.nvar - 4
        self.irhox = 4 if self.naux > 0 else -1
        self.nq = 4 + self.naux
        self.irho, self.iu, self.iv, self.ip = 0, 1, 2, 3
        self.ix = 4 if self.naux > 0 else -1


def cons_to_prim(U, gamma, ivars, myg):
    """conserved [i, j, n] -> primitive (rho, u, v, p) (simulation.py:49-80), on the device"""
    q = myg.scratch_array(nvar=ivars.nq)
    rho = U[:, :, ivars.idens].t()
    nz = rho != 0.0
    safe = torch.where(nz, rho, torch.ones_like(rho))
    u = torch.where(nz, U[:, :, ivars.ixmom].t() / safe, torch.zeros_like(rho))
    v = torch.where(nz, U[:, :, ivars.iymom].t() / safe, torch.zeros_like(rho))
    e = torch.where(nz, (U[:, :, ivars.iener].t() - 0.5 * rho * (u ** 2 + v ** 2)) / safe, torch.zeros_like(rho))
    g = myg
    valid = (slice(g.ilo, g.ihi + 1), slice(g.jlo, g.jhi + 1))
    e_min, rho_min = float(e[valid].min()), float(rho[valid].min())
    assert e_min > 0.0 and rho_min > 0.0, f"invalid state, min(rho) = {rho_min}, min(e) = {e_min}"
    q[:, :, ivars.irho] = rho
    q[:, :, ivars.iu] = u
    q[:, :, ivars.iv] = v
    q[:, :, ivars.ip] = eos.pres(gamma, rho, e)
    return q


def prim_to_cons(q, gamma, ivars, myg):
    """primitive -> conserved (simulation.py:83-102)"""
    U = myg.scratch_array(nvar=ivars.nvar)
    rho, u, v, p = (q[:, :, k].t() for k in (ivars.irho, ivars.iu, ivars.iv, ivars.ip))
    U[:, :, ivars.idens] = rho
    U[:, :, ivars.ixmom] = u * rho
    U[:, :, ivars.iymom] = v * rho
    U[:, :, ivars.iener] = eos.rhoe(gamma, p) + 0.5 * rho * (u ** 2 + v ** 2)
    return U


class Simulation(NullSimulation):
    """unsplit CTU compressible hydrodynamics; same life cycle as the reference class"""

    def initialize(self, *, extra_vars=None, ng=4):
        if extra_vars:
            raise NotImplementedError("passively advected extra variables are not in the device sweep")
        rp = self.rp
        my_grid = grid_setup(rp, ng=ng, decomposition=self.decomposition)
        if ng < 4 or ng % 2:
            raise ValueError("the compressible sweep needs an even ng >= 4 (dependency radius, SURVEY.md 9.3; 16-byte "
                             "aligned row copies)")
        my_data = self.data_class(my_grid)

        riemann_method = rp.get_param("compressible.riemann")
        if riemann_method not in ("HLLC", "HLLC_lm", "CGF"):
            msg.fail("ERROR: Riemann solver undefined")
        self._spherical = getattr(my_grid, "coord_type", 0) == 1
        if self._spherical and riemann_method != "CGF":
            # the reference refuses HLLC (simulation.py:201-209): the update needs the interface pressure
            msg.fail("ERROR: the SphericalPolar geometry needs the CGF Riemann solver")
        # solver-specific boundary types (simulation.py:212-214)
        bnd.define_bc("hse", BC.user, is_solid=False)
        bnd.define_bc("ambient", BC.user, is_solid=False)
        bnd.define_bc("ramp", BC.user, is_solid=False)
        try:
            if rp.get_param("particles.do_particles") == 1:
                msg.fail("ERROR: particles are not supported")
        except KeyError:
            pass
        # problem source terms: the reference's heating / plume / convection problems all heat at the rate
        # dens * e_rate * profile(x, y); a problem describes that with heating(grid, rp) -> (e_rate, profile)
        self._heating = None
        if self.problem_source is not None:
            import sys
            self._heating = getattr(sys.modules.get(self.problem_func.__module__), "heating", None)
            if self._heating is None:
                msg.fail("ERROR: problem source terms need a heating(grid, rp) description for the device sweep")

        bc, bc_xodd, bc_yodd = bc_setup(rp)
        self.solid = bnd.bc_is_solid(bc)
        # the reference fills the ghost cells of its gravity-source arrays odd / even across a reflecting
        # y wall (simulation.py:248-253): the sweep flips the sign of the ghost-cell sources there
        self._src_flip = (int(bc.ylb in ("reflect", "reflect-even", "reflect-odd")),
                          int(bc.yrb in ("reflect", "reflect-even", "reflect-odd")))

        # registration order fixes the variable indices: dens 0, ener 1, xmom 2, ymom 3
        # (simulation.py:223-226) -- the kernels rely on it
        my_data.register_var("density", bc)
        my_data.register_var("energy", bc)
        my_data.register_var("x-momentum", bc_xodd)
        my_data.register_var("y-momentum", bc_yodd)
        my_data.set_aux("gamma", rp.get_param("eos.gamma"))
        my_data.set_aux("grav", rp.get_param("compressible.grav"))
        dec = self.decomposition
        from .. import _lib as _libmod
        on_device = my_grid.device.type == "cuda" or getattr(_libmod.lib(), "is_emulated", False)
        peer = dec is not None and dec.size > 1 and on_device and hasattr(dec, "shared_planes")
        if peer:
            # decomposed: both state buffers live in memory the neighbours can write -- halo rows and the wave-speed
            # reduction then go through peer memory (csrc/slab_comm.cu) instead of NCCL
            my_data.create(planes=dec.shared_planes(my_data.nvar, my_grid.qx, my_grid.qy, bc.xlb == "periodic"))
        else:
            my_data.create()
        my_data.decomposition = self.decomposition
        self.cc_data = my_data
        # artificial viscosity is left unset on the GLOBAL +x face only
        self._no_avisc_xhi = 1 if (self.decomposition is None or self.decomposition.is_last) else 0

        self.ivars = Variables(my_data)
        assert (self.ivars.idens, self.ivars.iener, self.ivars.ixmom, self.ivars.iymom) == (0, 1, 2, 3)
        self.cc_data.add_derived(derives.derive_primitives)

        # second state buffer (the sweep is out of place) and the device scratch words
        self._alt_planes = (dec.shared_planes(my_data.nvar, my_grid.qx, my_grid.qy, bc.xlb == "periodic") if peer
                            else torch.zeros_like(my_data.planes))
        self._scratch = ops.new_scratch()
        self._wave_version = None     # cc_data.version for which the cached wave speeds are valid
        self._pending_status = False

        self._geometry = None
        if self._spherical:
            if any(t in bnd.ext_bcs for t in bc.names()):
                msg.fail("ERROR: the hse / ambient / ramp boundaries are not built for SphericalPolar grids")
            # (one entry more than the padded row: the kernel also reads column j + 1 of the per-column tables)
            xl_type, xr_type = bc.xlb, bc.xrb
            lo_int = hi_int = False
            if self.decomposition is not None and self.decomposition.size > 1:
                lo_int, hi_int = self.decomposition.interior_sides(bc.xlb == "periodic")
                xl_type, xr_type = (None if lo_int else bc.xlb), (None if hi_int else bc.xrb)
            gi, gj = patch.spherical_sweep_tables(my_grid, my_data.planes.stride(1) + 2, xl_type, xr_type)
            dev = my_data.planes.device
            self._geometry = (torch.from_numpy(gi).to(dev), torch.from_numpy(gj).to(dev))
            # across a "reflect" x boundary the reference's source arrays change sign (their own BCs, odd for the
            # normal momentum's source, with the state's parities: every product of them is -1); the literal
            # reflect-even / reflect-odd types and everything else leave the sign alone
            self._src_flip_x = (int(rp.get_param("mesh.xlboundary") == "reflect" and not lo_int),
                                int(rp.get_param("mesh.xrboundary") == "reflect" and not hi_int))
            # ... and no flips in y: the parities of the spherical sources match those of the state there
            self._src_flip = (0, 0)

        self.problem_func(self.cc_data, self.rp)
        self._heat_rate, self._heat_plane = 0.0, None
        if self._heating is not None:
            import numpy as np
            rate, prof = self._heating(my_grid, rp)
            plane = ops.alloc_planes(1, my_grid.qx, my_grid.qy, device=my_data.planes.device)
            assert plane.stride(1) == my_data.planes.stride(1)
            plane[0, :, :my_grid.qy].copy_(torch.from_numpy(np.ascontiguousarray(prof, dtype=np.float64)))
            # the reference ghost-fills its energy-source array with the scalar BCs (user types copy like outflow):
            # filling the profile the same way gives the kernel the source of the cell each ghost cell mirrors
            names = ["outflow" if t in bnd.ext_bcs else t for t in bc.names()]
            if self.decomposition is not None and self.decomposition.size > 1:
                # the profile was evaluated on the slab's own coordinates: the rows facing another slab are right already
                lo_int, hi_int = self.decomposition.interior_sides(bc.xlb == "periodic")
                names = [None if lo_int else names[0], None if hi_int else names[1], names[2], names[3]]
            ops.fill_ghost(plane, my_grid.nx, my_grid.ny, my_grid.ng, [tuple(names)])
            self._heat_rate, self._heat_plane = float(rate), plane
        if self.verbose > 0:
            print(my_data)

    # ---- parameters of the sweep ------------------------------------------------------------
    def _comp_params(self):
        rp = self.rp
        return ops.comp_params(gamma=rp.get_param("eos.gamma"), z0=rp.get_param("compressible.z0"),
                               z1=rp.get_param("compressible.z1"), delta=rp.get_param("compressible.delta"),
                               cvisc=rp.get_param("compressible.cvisc"),
                               limiter=rp.get_param("compressible.limiter"),
                               use_flattening=rp.get_param("compressible.use_flattening"),
                               no_avisc_xhi=getattr(self, "_no_avisc_xhi", 1), no_avisc_yhi=1,
                               grav=rp.get_param("compressible.grav"),
                               src_flip_ylo=self._src_flip[0], src_flip_yhi=self._src_flip[1],
                               heat_rate=self._heat_rate, heat_profile=self._heat_plane,
                               src_copy_yhi=int(self.cc_data.BCs["density"].yrb == "ambient"),
                               sponge=(rp.get_param("sponge.sponge_rho_begin"), rp.get_param("sponge.sponge_rho_full"),
                                       rp.get_param("sponge.sponge_timescale")) if rp.get_param("sponge.do_sponge") else None,
                               geometry=self._geometry,
                               src_flip_xlo=self._src_flip_x[0] if self._geometry is not None else 0,
                               src_flip_xhi=self._src_flip_x[1] if self._geometry is not None else 0,
                               riemann=rp.get_param("compressible.riemann"),
                               xl_solid=int(self.solid.xl) if (self.decomposition is None or self.decomposition.is_first) else 0,
                               yl_solid=int(self.solid.yl))

    def _read_scratch(self):
        """one D2H copy: wave-speed maxima + status word of the last sweep"""
        if self.decomposition is not None and self.decomposition.size > 1:
            # one all-reduce for everything: positive doubles order like their bit patterns (the same
            # trick the kernel's atomicMax uses), the status word is a small integer
            self.decomposition.allreduce_max_(self._scratch[:4])
        words = self._scratch[:4].cpu()
        if self._pending_status:
            self._pending_status = False
            if int(words[3]) != 0:
                raise AssertionError("invalid state: rho <= 0 or e <= 0 in a valid zone "
                                     "(compressible/simulation.py:71)")
        return words[:2].view(torch.float64).tolist()

    def check_state(self):
        """raise now if the last evolve() saw an invalid state (otherwise raised by the next dt)"""
        if getattr(self, "_pending_status", False):
            self._read_scratch()

    def method_compute_timestep(self):
        """CFL timestep (simulation.py:267-288): cfl * min(dx/(|u|+cs), dy/(|v|+cs)), bit-identical.
        Uses the maxima the last sweep accumulated when nothing touched the data since; the ghost
        cells the reference includes cannot change the minimum for the standard boundary types
        (SURVEY.md 9.2-3)."""
        cfl = self.rp.get_param("driver.cfl")
        g = self.cc_data.grid
        if getattr(self, "_spherical", False):
            # min(Lx / (|u| + cs), Ly / (|v| + cs)) over the whole array (simulation.py:285-288): Ly = r dtheta shrinks
            # towards the inner ghost rows, so the ghost cells can set the step and the fused maxima do not apply
            if self._pending_status:
                self._read_scratch()
            u, v, cs = self.cc_data.get_var(["velocity", "soundspeed"])
            xtmp = g.Lx.t() / (u.t().abs() + cs.t())
            ytmp = g.Ly.t() / (v.t().abs() + cs.t())
            local = torch.minimum(xtmp.min(), ytmp.min())
            if self.decomposition is not None and self.decomposition.size > 1:
                local = -self.decomposition.allreduce_max_((-local).reshape(1))[0]     # min over the slabs
            self.dt = cfl * float(local)
            return
        standard = all(t not in bnd.ext_bcs for b in self.cc_data.BCs.values() for t in b.names())
        if self._wave_version is not None and self._wave_version == self.cc_data.version and standard:
            wx, wy = self._read_scratch()
        else:
            if self._pending_status:
                self._read_scratch()
            wx, wy = ops.cfl_wavemax(self.cc_data.planes, g.nx, g.ny, g.ng, self.rp.get_param("eos.gamma"),
                                     self._scratch)
            if self.decomposition is not None and self.decomposition.size > 1:
                # as four 64-bit words (positive doubles order like their bit patterns): the form the peer-memory
                # reduction of the sweep's scratch words takes, so this path needs no other transport
                w = torch.tensor([wx, wy, 0.0, 0.0], dtype=torch.float64, device=self.cc_data.planes.device).view(torch.int64)
                self.decomposition.allreduce_max_(w)
                wx, wy = w.view(torch.float64)[:2].tolist()
        self.dt = cfl * float(min(g.dx / wx, g.dy / wy))

    def evolve(self):
        """advance the state through dt with the fused sweep (simulation.py:290-450)"""
        tm_evolve = self.tc.timer("evolve")
        tm_evolve.begin()
        myd = self.cc_data
        g = myd.grid
        self.clean_state(None)
        ops.compressible_sweep(myd.planes, self._alt_planes, g.nx, g.ny, g.ng, g.dx, g.dy, float(self.dt),
                               self._comp_params(), self._scratch)
        # the new state lives in the other buffer; its ghost cells are filled by the next fill_BC
        myd.planes, self._alt_planes = self._alt_planes, myd.planes
        self._wave_version = myd.version
        self._pending_status = True
        myd.t += self.dt
        self.n += 1
        tm_evolve.end()

    # ---- host-resident state: one step with the copies pipelined against the sweep ------------------------------
    def step_streamed(self, host_in, host_out, nchunks=16):
        """One driver step (fill_BC_all -> compute_timestep -> evolve, pyro/pyro_sim.py:241-256) of a state that lives
        in HOST memory: host_in -> device, step, device -> host_out (both pinned, shaped and strided like
        cc_data.planes; the x ghost rows of host_in need not be current, host_out's valid rows are written).

        The fused sweep consumes rows in order and needs only 4 rows of context, so the grid is cut into nchunks row
        blocks: block c + 1 travels host -> device on one stream while block c is swept on the compute stream and block
        c - 1 travels device -> host on a third: PCIe runs in both directions at once and the step costs about one
        transfer of the state instead of two plus the sweep.  Every block is an x-slab of the domain exactly as in the
        multi-GPU decomposition (interior block faces get the artificial viscosity, the global +x face does not --
        SURVEY.md 9.2-13), so the result is bit-identical to single_step() on a resident state.

        The time step needs max(|u| + cs) over the WHOLE incoming state before the first block can be swept.  When
        host_in is the buffer the previous step_streamed() wrote, those maxima are the ones the previous sweep's
        epilogue accumulated (same data, same bits as the stand-alone reduction); otherwise the state is copied in one
        piece first and reduced on the device (no overlap for that step)."""
        myd = self.cc_data
        g = myd.grid
        ng, nx, ny = g.ng, g.nx, g.ny
        if (self.decomposition is not None and self.decomposition.size > 1) or getattr(self, "_spherical", False):
            raise NotImplementedError("step_streamed: single-GPU Cartesian grids")
        names = [myd.BCs[n].names() for n in myd.names]
        if any(t in bnd.ext_bcs or t == "periodic" for b in names for t in b[:2]) or any(t in bnd.ext_bcs for b in names for t in b):
            raise NotImplementedError("step_streamed: standard, non-periodic x boundaries")
        dev = myd.planes
        for h in (host_in, host_out):
            assert h.shape == dev.shape and h.stride() == dev.stride() and h.dtype == dev.dtype and h.is_pinned()
        if getattr(self, "_st_h2d", None) is None:
            self._st_h2d, self._st_d2h = torch.cuda.Stream(), torch.cuda.Stream()
            self._chunk_scratch = None
        nchunks = max(1, min(int(nchunks), nx // (4 * ng)))      # a block is a grid of its own: >= 4 ng rows
        if self._chunk_scratch is None or self._chunk_scratch.shape[0] != nchunks:
            self._chunk_scratch = torch.zeros((nchunks, 8), dtype=torch.int64, device=dev.device)
        bounds = [ng + (nx * c) // nchunks for c in range(nchunks + 1)]        # block c = rows bounds[c] .. bounds[c+1]-1
        cur = torch.cuda.current_stream()
        h2d_done = [torch.cuda.Event() for _ in range(nchunks)]
        swept = [torch.cuda.Event() for _ in range(nchunks)]

        known = getattr(self, "_streamed_out_ptr", None) == host_in.data_ptr() and self._wave_version == myd.version
        self._st_h2d.wait_stream(cur)              # the device buffer may still be read by earlier work
        with torch.cuda.stream(self._st_h2d):
            for c in range(nchunks):
                a, b = bounds[c], bounds[c + 1]
                for k in range(dev.shape[0]):       # rows a..b-1 of ONE plane are contiguous: a plain DMA each
                    dev[k, a:b].copy_(host_in[k, a:b], non_blocking=True)
                h2d_done[c].record()
        if not known:
            # the whole state first, then the ordinary ghost fill and CFL reduction
            cur.wait_stream(self._st_h2d)
            myd.version += 1
            myd.fill_BC_all()
        self.compute_timestep()
        prm = self._comp_params()
        xl_solid = prm.xl_solid
        out = self._alt_planes
        small = self.rp.get_param("compressible.small_dens")
        if not known:
            self.clean_state(None)                  # the state is resident already
        # (otherwise the density floor is applied block by block as the rows arrive: after the block's ghost fill, like
        # clean_state() after fill_BC_all() in the resident step)

        def block_bcs(c):
            # an x side of a block that faces another block is left alone (its rows arrive with that block)
            return [(b[0] if c == 0 else None, b[1] if c == nchunks - 1 else None, b[2], b[3]) for b in names]

        def fill_block(c):
            # the block's own rows (plus the global x ghost rows for the first / last block) as a grid of its own whose
            # outermost ng rows play the part of ghost rows: x sides that face another block are None (left alone),
            # the y fill runs over exactly these rows
            a, b = bounds[c], bounds[c + 1]
            cur.wait_event(h2d_done[c])
            lo = 0 if c == 0 else a
            hi = g.qx if c == nchunks - 1 else b
            ops.fill_ghost(dev[:, lo:hi], hi - lo - 2 * ng, ny, ng, block_bcs(c))
            if small > -1.e100:
                dev[0, a:b, g.jlo:g.jhi + 1].clamp_(min=small)

        if known:
            fill_block(0)
        for c in range(nchunks):
            a, b = bounds[c], bounds[c + 1]
            if known and c + 1 < nchunks:
                fill_block(c + 1)                   # block c reads 4 rows of block c + 1
            prm.no_avisc_xhi = 1 if c == nchunks - 1 else 0
            prm.xl_solid = xl_solid if c == 0 else 0           # the solid-wall rule of the CGF solver: the domain's -x face only
            if self._heat_plane is not None:                   # the heating profile is a plane like the state's: same rows
                prm.heat_profile = self._heat_plane[0, a - ng:].data_ptr()
            ops.compressible_sweep(dev[:, a - ng:b + ng], out[:, a - ng:b + ng], b - a, ny, ng, g.dx, g.dy, float(self.dt),
                                   prm, self._chunk_scratch[c])
            swept[c].record()
            with torch.cuda.stream(self._st_d2h):
                self._st_d2h.wait_event(swept[c])
                for k in range(out.shape[0]):
                    host_out[k, a:b].copy_(out[k, a:b], non_blocking=True)
        # maxima and status of the whole grid from the blocks' (positive doubles order like their bit patterns)
        self._scratch[:4] = self._chunk_scratch[:, :4].max(dim=0).values
        cur.wait_stream(self._st_d2h)               # the step is complete when its result is on the host
        myd.planes, self._alt_planes = self._alt_planes, myd.planes
        self._wave_version = myd.version
        self._pending_status = True
        self._streamed_out_ptr = host_out.data_ptr()
        myd.t += self.dt
        self.n += 1

    def clean_state(self, U):   # pylint: disable=unused-argument
        """density floor (simulation.py:452-456); a no-op for the default small_dens = -1e200"""
        small = self.rp.get_param("compressible.small_dens")
        if small > -1.e100:
            g = self.cc_data.grid
            d = self.cc_data.planes[0, g.ilo:g.ihi + 1, g.jlo:g.jhi + 1]
            d.clamp_(min=small)

    def dovis(self):
        """runtime visualisation is host-side matplotlib in the reference; not part of the device build"""

    def write_extras(self, f):
        gb = f.create_group("BC")
        gb.create_dataset("hse", data=False)
        gb.create_dataset("ambient", data=False)

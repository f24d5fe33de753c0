// sweep_task.cuh -- one warp's share of the fused compressible CTU sweep (HP-1).
//
// Replaces, in a single pass over the state, everything the reference does in
// pyro/compressible/simulation.py:290-450 (evolve); the default instantiation is Cartesian / HLLC / no sources, the
// template parameters of SweepTask add the source terms, the other Riemann solvers and SphericalPolar geometry:
//   cons_to_prim -> flatten/flatten_multid -> limit -> interface.states (x, y) -> prim_to_cons ->
//   transverse riemann_hllc (x, y) -> transverse correction -> final riemann_hllc (x, y) ->
//   artificial viscosity -> conservative update, plus the wave-speed maxima the next
//   method_compute_timestep() needs (simulation.py:267-288).
//
// Decomposition ("marching strips"): a warp owns a strip of 30 output columns (y, the contiguous
// axis; lanes 0 and 31 are the +-1 halo columns whose interface states the neighbours need) and
// marches along x over a segment of rows.  Everything that couples rows (x-direction stencils,
// the x Riemann problems, the row-lagged update) lives in the lane's own registers, carried from
// one iteration to the next; everything that couples columns goes through warp shuffles or the
// warp's shared-memory row ring.  No block-level barrier exists anywhere: warps are independent.
//
//   iteration i (i0-1 <= i <= i1) does
//     ready row i+3 (TMA bulk copy landed -> convert cons->prim in place in the ring)
//     xi(i), limited slopes(i), traced states of row i                     [hat states]
//     transverse F_x at face i-1/2 (needs row i-1 from regs)  -> dF_x of row i-1
//     corrected y-states of row i-1 -> final F_y of row i-1 (+ viscosity)
//     transverse F_y of row i (lane shuffles) -> dF_y of row i -> corrected x-states of row i
//     final F_x at face i-1/2 (+ viscosity)
//     conservative update of row i-1, store, wave-speed maxima
//
// HBM traffic per updated cell: one read of U (the 38/30 column overlap and the 8-row segment
// overlap are absorbed by L2) and one write: 64 B.  Dependency radius is exactly ng = 4
// (SURVEY.md 9.3), which is why rows i0-4 .. i1+3 and columns j0-4 .. j0+33 are staged.
//
// The code is written against a small "warp services" policy W so the identical source can be
// compiled for a host-side lane-per-thread emulator (tests/emu) -- test infrastructure only.
#pragma once
#include "hydro_core.cuh"

namespace pyro {

constexpr int SW_OUT = 30;     // output columns per warp
constexpr int SW_QW = 38;      // staged columns: 32 lanes + 3 each side
constexpr int SW_RING = 8;     // row ring depth (rows i-2 .. i+5)
constexpr int SW_XW = 34;      // exchange width: 32 lanes + 1 each side

struct SweepArgs {
    const double* Uin;         // 4 planes [n][i][j], ghosts filled
    double* Uout;              // 4 planes, valid region written
    long long plane_stride;    // elements between planes
    int pitch;                 // elements between rows (>= qy, even)
    int nx, ny, ng;            // ng >= 4
    double dx, dy, dt, gamma;
    double z0, z1, delta, cvisc;
    int limiter, use_flattening;
    int no_avisc_xhi, no_avisc_yhi;   // reference quirk 9.2-13 on the global +x / +y faces
    int nstrips, nsegs, seglen;
    unsigned long long* wavemax;      // [0] = bits of max(|u|+cs), [1] = bits of max(|v|+cs)
    int* status;                      // set to 1 when a valid cell has rho <= 0 or e <= 0
#ifdef SWEEP_DEBUG
    double* dbg;                      // host emulator only: [k][i][j] planes of intermediates
#endif
    // gravity (GRAV instantiation only; kept at the end so the default kernel's parameter layout is unchanged)
    double grav;                      // compressible.grav, acceleration along y
    int src_flip_ylo, src_flip_yhi;   // 1: that y boundary reflects -> the ghost-cell SOURCES change sign
    int xl_solid, yl_solid;           // CGF: the -x / -y boundary is a solid wall (boundary.bc_is_solid)
    // more sources of the GRAV ("with sources") instantiations
    const double* heat;               // heating profile P (one plane, same pitch, ghost-filled like a scalar) or NULL:
    double heat_rate;                 //   S_ener += dens * heat_rate * P  (problems heating / plume / convection)
    int do_sponge;                    // sponge (simulation.py:164-184, 425-441)
    double sponge_rho_begin, sponge_rho_full, sponge_timescale;
    int src_copy_yhi;                 // 1: the +y boundary is "ambient": its ghost cells hold a constant state, but the
                                      //    reference's source arrays are zero-gradient copies of row jhi there
    // SphericalPolar grids (SPH instantiation; x = r, y = theta): separable geometry tables built on the host with the
    // reference's numpy expressions (mesh/patch.py:242-312), so that products formed here in the reference's order
    // reproduce its Ax, Ay, V, Ly, dlogA arrays bit for bit.
    //   geo_i[k * geo_ni + i], k = 0 r_i (cell centre), 1 r of the row whose source the reference's source arrays hold
    //                          in row i (itself; the boundary image for ghost rows), 2 -2 pi rl^2, 3 rr^2 - rl^2,
    //                          4 rr - rl, 5 rr^2 + rl^2 + rr rl, 6..8 the viscosity's vertex radii (rr, rl, rc)
    //   geo_j[k * geo_nj + j], k = 0 cos(th_r) - cos(th_l), 1 -2 pi / 3 times that, 2 pi sin(th_l), 3 tan(th),
    //                          4..6 the viscosity's sines at / around the vertex (sint, sinb, sinc)
    const double* geo_i;
    const double* geo_j;
    int geo_ni, geo_nj;
    int src_flip_xlo, src_flip_xhi;   // 1: that x boundary is "reflect" -> the ghost-row SOURCES change sign
};

// One ring slot = one staged row: the four variables' 38 columns back to back -- the box of ONE 3-d TMA tensor copy
// (columns x 1 row x 4 planes, 1216 bytes) -- padded to the 128-byte alignment the tensor engine wants for its destination.
struct alignas(128) SweepSlot {
    double p[4][SW_QW];               // raw U on arrival, primitives after ready()
    double pad[8];
};
struct alignas(128) SweepSmem {
    SweepSlot q[SW_RING];
    double xch[5][SW_XW];             // [0] xi_y, [1..4] limit2_y of the 4 primitives
    unsigned long long mbar[SW_RING];
};

// ---------------------------------------------------------------------------------------------
// GRAV = true adds the constant-gravity source terms of the reference: half a time step of
// S = (0, rho g v... ) on the traced interface states (apply_source_terms, unsplit_fluxes.py:247-330) and the
// predictor-corrector source update after the flux divergence (simulation.py:398-423,
// get_external_sources :105-160).  The reference fills the ghost cells of the source ARRAYS with their own
// BCs (ymom_src odd, E_src even across a reflecting y wall), which is the source of the ghost STATE with
// the sign flipped there and identical to it for every other boundary type.
// RIEMANN: 0 = HLLC (riemann_hllc), 1 = CGF (riemann_cgf + consFlux), 2 = low-Mach HLLC (riemann_hllc_lowspeed);
// selected by compressible.riemann.
// SPH = true (with GRAV = true, RIEMANN = 1 -- the reference insists on CGF there): SphericalPolar geometry, i.e. the
// coord_type == 1 branches of interface.states, get_external_sources, riemann_flux / consFlux,
// apply_transverse_flux, the conservative update and the artificial viscosity.
template <class W, bool GRAV = false, int RIEMANN = 0, bool SPH = false>
struct SweepTask {
    W& w;
    const SweepArgs& A;
    SweepSmem& S;
    // mbarrier parities: the parity slot s has at the START of the running task (one bit per ring slot, warp-uniform,
    // updated once per task) and the first row the task stages; the parity of a row's wait follows from the two -- no
    // per-row read-modify-write of carried state (it used to be spilled: 1.5% of the executed instructions)
    unsigned phase_base;
    int row_base;

    HD SweepTask(W& w_, const SweepArgs& a, SweepSmem& s, unsigned ph) : w(w_), A(a), S(s), phase_base(ph), row_base(0) {}

    HD double& Q(int n, int r, int c) { return S.q[r & (SW_RING - 1)].p[n][c]; }

    // the Riemann problem at a face; `wall`: the face lies on a solid lower boundary (CGF only)
    HD Flux riemann(double rho_l, double E_l, double mn_l, double mt_l, double rho_r, double E_r, double mn_r,
                    double mt_r, const HllcPar& hp, bool wall)
    {
        if (RIEMANN == 1) return cgf(rho_l, E_l, mn_l, mt_l, rho_r, E_r, mn_r, mt_r, hp, wall);
        if (RIEMANN == 2) return hllc_lm(rho_l, E_l, mn_l, mt_l, rho_r, E_r, mn_r, mt_r, hp);
        return hllc(rho_l, E_l, mn_l, mt_l, rho_r, E_r, mn_r, mt_r, hp);
    }

    // SphericalPolar: CGF flux without the pressure in the normal momentum, and the interface pressure
    HD Flux riemann_p(double rho_l, double E_l, double mn_l, double mt_l, double rho_r, double E_r, double mn_r,
                      double mt_r, const HllcPar& hp, bool wall, double& pface)
    {
        return cgf_impl<true>(rho_l, E_l, mn_l, mt_l, rho_r, E_r, mn_r, mt_r, hp, wall, pface);
    }

    // wait for the bulk copy of row r, convert cons -> prim in place, publish to the warp
    HD void ready(int r, int col0, int jvalid_lo, int jvalid_hi, bool row_valid)
    {
        const int slot = r & (SW_RING - 1);
        // rows row_base, row_base + 1, ... use the slots in turn: this is wait number (r - row_base) / 8 on this slot
        w.load_wait(S.mbar[slot], ((phase_base >> slot) ^ ((unsigned)(r - row_base) >> 3)) & 1u);
        bool anybad = false;
#pragma unroll 1
        for (int c = w.lane(); c < SW_QW; c += 32) {
            Cons U;
            U.dens = S.q[slot].p[IDENS][c]; U.ener = S.q[slot].p[IENER][c];
            U.xmom = S.q[slot].p[IXMOM][c]; U.ymom = S.q[slot].p[IYMOM][c];
            bool bad;
            Prim p = cons_to_prim(U, A.gamma, &bad);
            S.q[slot].p[IRHO][c] = p.rho; S.q[slot].p[IU][c] = p.u;
            S.q[slot].p[IV][c] = p.v;     S.q[slot].p[IP][c] = p.p;
            int j = col0 + c;
            if (bad && row_valid && j >= jvalid_lo && j < jvalid_hi) anybad = true;
        }
        if (anybad) *A.status = 1;
        w.sync();
    }

    HD void issue(int r, int col0, int ncols)
    {
        const int slot = r & (SW_RING - 1);
        // row r, columns col0 .. col0 + 37 of the four planes -> the slot (columns past the end of the row: not read)
        w.load_issue(S.mbar[slot], &S.q[slot].p[0][0], A.Uin, A.plane_stride, A.pitch, r, col0, ncols);
    }

    HD double flat_x(int r, int c, const FlatPar& fp)
    {
        return flatten_1d(Q(IP, r - 2, c), Q(IP, r - 1, c), Q(IP, r + 1, c), Q(IP, r + 2, c),
                          Q(IU, r - 1, c), Q(IU, r + 1, c), fp);
    }

    HD void run(int strip, int seg)
    {
        const int lane = w.lane();
        const int ng = A.ng;
        const int i0 = ng + seg * A.seglen;
        const int i1 = (i0 + A.seglen < ng + A.nx) ? i0 + A.seglen : ng + A.nx;   // exclusive
        const int jout0 = ng + SW_OUT * strip;          // first output column
        const int col0 = jout0 - 4;                     // global column of ring column 0
        const int ncols = (A.pitch - col0 < SW_QW) ? A.pitch - col0 : SW_QW;
        const int j = jout0 - 1 + lane;                 // this lane's global column
        const int cc = lane + 3;                        // ... and its ring column
        const int jhi = ng + A.ny;                      // one past the last valid column
        const int ihi = ng + A.nx;
        const int rlast = i1 + 3;
        const bool out_lane = (lane >= 1 && lane <= SW_OUT && j < jhi);

        const double gamma = A.gamma;
        const double ginv1 = 1.0 / (gamma - 1.0);
        const HllcPar hp = hllc_par(gamma);
        FlatPar fp; fp.z0 = A.z0; fp.inv_dz = 1.0 / (A.z1 - A.z0); fp.delta = A.delta;
        const double dtdx = A.dt / A.dx, dtdy = A.dt / A.dy;
        const double hdtdx = 0.5 * dtdx, hdtdy = 0.5 * dtdy;
        const double dxinv = 1.0 / A.dx, dyinv = 1.0 / A.dy;
        const int lim = A.limiter;
        const bool flat = A.use_flattening != 0;
        // CGF zeroes the normal velocity on a solid -x / -y boundary face (x face index ng, y face index ng)
        const bool xwall = RIEMANN == 1 && A.xl_solid != 0;
        const bool ywall = RIEMANN == 1 && A.yl_solid != 0 && j == ng;

        // SphericalPolar: this lane's theta factors (and those of column j + 1, whose cell volume and low-face area
        // enter the corrections of the states this lane owns)
        const int jg = (j < A.geo_nj - 1) ? j : A.geo_nj - 2;
        const double g_cd = SPH ? A.geo_j[jg] : 0.0, g_kc = SPH ? A.geo_j[A.geo_nj + jg] : 0.0;
        const double g_kc1 = SPH ? A.geo_j[A.geo_nj + jg + 1] : 0.0;
        const double g_ps = SPH ? A.geo_j[2 * A.geo_nj + jg] : 0.0, g_ps1 = SPH ? A.geo_j[2 * A.geo_nj + jg + 1] : 0.0;
        const double g_tn = SPH ? A.geo_j[3 * A.geo_nj + jg] : 1.0;
        const double g_sint = SPH ? A.geo_j[4 * A.geo_nj + jg] : 0.0, g_sinb = SPH ? A.geo_j[5 * A.geo_nj + jg] : 0.0;
        const double g_sinc = SPH ? A.geo_j[6 * A.geo_nj + jg] : 0.0;
        const double* GI = A.geo_i;
        auto gi = [&](int k, int r) { return GI[k * A.geo_ni + (r < 0 ? 0 : (r < A.geo_ni ? r : A.geo_ni - 1))]; };
        auto area_x = [&](int r) { return fabs(gi(2, r) * g_cd); };                 // Ax(r, j): the low-r face
        auto area_y = [&](int r) { return fabs(g_ps * gi(3, r)); };                 // Ay(r, j): the low-theta face
        auto area_y1 = [&](int r) { return fabs(g_ps1 * gi(3, r)); };               // Ay(r, j + 1)
        auto vol = [&](int r) { return fabs(g_kc * gi(4, r) * gi(5, r)); };         // V(r, j)
        auto vol1 = [&](int r) { return fabs(g_kc1 * gi(4, r) * gi(5, r)); };       // V(r, j + 1)

        // raw U of the lane's own column is re-read from global (L2 hit: the bulk copy just
        // streamed it) so that the update adds the flux divergence to the unmodified state
        const long long jj = (j < A.pitch) ? j : A.pitch - 1;
        const double* Ucol = A.Uin + jj;
        double* Ocol = A.Uout + jj;

        row_base = i0 - 4;
        // ---- prologue: rows i0-4 .. i0+3 in flight, i0-4 .. i0+1 ready ---------------------
        for (int r = i0 - 4; r <= i0 + 3 && r <= rlast; ++r) issue(r, col0, ncols);
        for (int r = i0 - 4; r <= i0 + 1; ++r) ready(r, col0, ng, jhi, r >= ng && r < ihi);

        double xix_m1 = 1.0, xix_0 = 1.0;               // xi_x of rows i-1, i
        if (flat) { xix_m1 = flat_x(i0 - 2, cc, fp); xix_0 = flat_x(i0 - 1, cc, fp); }
        double l2x_m1[4], l2x_0[4];                     // limit2_x of rows i-1, i
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            l2x_m1[n] = slope2(Q(n, i0 - 3, cc), Q(n, i0 - 2, cc), Q(n, i0 - 1, cc), lim);
            l2x_0[n] = slope2(Q(n, i0 - 2, cc), Q(n, i0 - 1, cc), Q(n, i0, cc), lim);
        }
        w.sync();
        if (i0 + 4 <= rlast) issue(i0 + 4, col0, ncols);

        // ---- state carried from row i-1 to row i (all column-local) -------------------------
        Cons XPc = {0, 0, 0, 0};      // traced state on the +x face of cell (i-1)   [ref: U_xl[i]]
        Cons XPpc = {0, 0, 0, 0};     // ... after the transverse correction
        Cons YMc = {0, 0, 0, 0}, YPc = {0, 0, 0, 0};    // traced y-face states of cell (i-1)
        Cons FxT_prev = {0, 0, 0, 0};                   // transverse x-flux at face (i-1)-1/2
        Cons Fx_prev = {0, 0, 0, 0};                    // final x-flux at face (i-1)-1/2
        Cons U_prev = {0, 0, 0, 0};                     // raw U(i-1, j)
        double divU_prev = 0.0;                         // vertex divergence at (i-1 -1/2, j-1/2)
        double pxT_prev = 0.0, pxF_prev = 0.0;          // SPH: interface pressures at face (i-1)-1/2 (transverse, final)
        double wmax_x = 0.0, wmax_y = 0.0;

        for (int i = i0 - 1; i <= i1; ++i) {
            ready(i + 3, col0, ng, jhi, (i + 3) >= ng && (i + 3) < ihi);

            // raw state of this row (latency hidden behind the hat-state arithmetic)
            Cons Uc;
            {
                const double* p = Ucol + (long long)i * A.pitch;
                Uc.dens = p[0]; Uc.ener = p[A.plane_stride];
                Uc.xmom = p[2 * A.plane_stride]; Uc.ymom = p[3 * A.plane_stride];
            }

            // ---- B. y-direction helpers on the 34-column exchange row ------------------------
            // exchange column ee <-> global column jout0-2+ee <-> ring column ee+2; a lane's own
            // column is ee = lane+1.  Second trip: lanes 0 and 1 do the two outer columns 0 and 33.
            for (int pass = 0; pass < 2; ++pass) {
                if (pass == 1 && lane >= 2) break;
                const int ee = (pass == 0) ? lane + 1 : (lane == 0 ? 0 : SW_XW - 1);
                const int c = ee + 2;
                double xiy = 1.0;
                if (flat)
                    xiy = flatten_1d(Q(IP, i, c - 2), Q(IP, i, c - 1), Q(IP, i, c + 1), Q(IP, i, c + 2),
                                     Q(IV, i, c - 1), Q(IV, i, c + 1), fp);
                S.xch[0][ee] = xiy;
#pragma unroll
                for (int n = 0; n < 4; ++n)
                    S.xch[1 + n][ee] = slope2(Q(n, i, c - 1), Q(n, i, c), Q(n, i, c + 1), lim);
            }
            w.sync();

            // ---- C. xi(i, j) (flatten_multid, reconstruction.py:167-183) ----------------------
            double xix_p1 = flat ? flat_x(i + 1, cc, fp) : 1.0;
            double xi = 1.0;
            if (flat) {
                double px = (Q(IP, i + 1, cc) - Q(IP, i - 1, cc) > 0.0) ? xix_m1 : xix_p1;
                double py = (Q(IP, i, cc + 1) - Q(IP, i, cc - 1) > 0.0) ? S.xch[0][lane] : S.xch[0][lane + 2];
                xi = dmin(dmin(xix_0, px), dmin(S.xch[0][lane + 1], py));
            }

#ifdef SWEEP_DEBUG
            if (A.dbg && j < A.pitch) {
                long long np_ = (long long)(A.nx + 2 * A.ng) * A.pitch;
                A.dbg[0 * np_ + (long long)i * A.pitch + j] = xi;
                A.dbg[1 * np_ + (long long)i * A.pitch + j] = xix_0;
                A.dbg[2 * np_ + (long long)i * A.pitch + j] = S.xch[0][lane + 1];
            }
#endif
            // ---- D. limited slopes (unsplit_fluxes.py:192-197) -------------------------------
            Prim q0;
            q0.rho = Q(IRHO, i, cc); q0.u = Q(IU, i, cc); q0.v = Q(IV, i, cc); q0.p = Q(IP, i, cc);
            double ldx[4], ldy[4], l2x_p1[4];
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                double am = Q(n, i - 1, cc), a0 = Q(n, i, cc), ap = Q(n, i + 1, cc);
                l2x_p1[n] = slope2(a0, ap, Q(n, i + 2, cc), lim);
                double sx = (lim == 2) ? slope4(am, a0, ap, l2x_m1[n], l2x_p1[n]) : l2x_0[n];
                ldx[n] = xi * sx;
                double bm = Q(n, i, cc - 1), bp = Q(n, i, cc + 1);
                double sy = (lim == 2) ? slope4(bm, a0, bp, S.xch[1 + n][lane], S.xch[1 + n][lane + 2])
                                       : S.xch[1 + n][lane + 1];
                ldy[n] = xi * sy;
            }

            // ---- E. characteristic tracing, both directions -> conserved face states ---------
            Cons XM, XP, YM, YP;
            {
                TraceGeom g = trace_geom(q0, gamma);
                Prim m, p;
                trace_1d(q0.rho, q0.u, q0.v, q0.p, ldx[IRHO], ldx[IU], ldx[IV], ldx[IP], g, dtdx,
                         m.rho, m.u, m.v, m.p, p.rho, p.u, p.v, p.p);
                if (SPH) {
                    // geometric source of the divergence, dlogAx = 2 / r (interface.py:218-225)
                    const double rs = -0.5 * A.dt * (2.0 / gi(0, i)) * q0.rho * q0.u;
                    m.rho += rs; p.rho += rs; m.p += rs * g.cs2; p.p += rs * g.cs2;
                }
                XM = prim_to_cons(m, ginv1); XP = prim_to_cons(p, ginv1);
                // SPH: the cell's own length along theta, Ly = r dtheta
                trace_1d(q0.rho, q0.v, q0.u, q0.p, ldy[IRHO], ldy[IV], ldy[IU], ldy[IP], g,
                         SPH ? A.dt / (gi(0, i) * A.dy) : dtdy,
                         m.rho, m.v, m.u, m.p, p.rho, p.v, p.u, p.p);
                if (SPH) {
                    // dlogAy = cot(theta) / r (interface.py:227-234)
                    const double rs = -0.5 * A.dt * (1.0 / (g_tn * gi(0, i))) * q0.rho * q0.v;
                    m.rho += rs; p.rho += rs; m.p += rs * g.cs2; p.p += rs * g.cs2;
                }
                YM = prim_to_cons(m, ginv1); YP = prim_to_cons(p, ginv1);
            }

            if (SPH) {
                // S_xmom = rho g + ymom^2 / (rho r), S_ymom = -xmom ymom / rho, S_ener = xmom g (radial gravity plus the
                // geometric terms, simulation.py:117-124).  The reference fills the ghost cells of the source ARRAYS
                // with their own BCs: for a ghost row that is the source of the row's boundary image -- the ghost
                // state itself evaluated at the image's radius, sign flipped across a reflecting wall; ghost columns
                // need nothing (the radius does not change and the parities of S match those of the state)
                const bool flip = (i < ng && A.src_flip_xlo) || (i >= ihi && A.src_flip_xhi);
                const double r = gi(1, i);
                double sx = Uc.dens * A.grav + Uc.ymom * Uc.ymom / (Uc.dens * r);
                double sy = -Uc.xmom * Uc.ymom / Uc.dens, se = Uc.xmom * A.grav;
                if (flip) { sx = -sx; sy = -sy; se = -se; }
                // problem heating: the profile plane is ghost-filled like the (even) source array it feeds
                if (A.heat) se += Uc.dens * A.heat_rate * A.heat[(long long)i * A.pitch + jj];
                const double hx = 0.5 * A.dt * sx, hy = 0.5 * A.dt * sy, he = 0.5 * A.dt * se;
                XM.xmom += hx; XP.xmom += hx; YM.xmom += hx; YP.xmom += hx;
                XM.ymom += hy; XP.ymom += hy; YM.ymom += hy; YP.ymom += hy;
                XM.ener += he; XP.ener += he; YM.ener += he; YP.ener += he;
            } else if (GRAV) {
                // U_xl[i+1], U_xr[i], U_yl[j+1], U_yr[j] += 0.5 dt S(i, j); S_ymom = rho g, S_ener = (rho v) g
                const bool flip = (j < ng && A.src_flip_ylo) || (j >= jhi && A.src_flip_yhi);
                // the state the ghost-cell source is evaluated from: the ghost state itself, except above an
                // "ambient" boundary, where the reference's source arrays copy the last valid row's source
                double sd = Uc.dens, sm = Uc.ymom;
                if (A.src_copy_yhi && j >= jhi) {
                    const int cl = jhi - 1 - col0;
                    sd = Q(IRHO, i, cl);
                    sm = sd * Q(IV, i, cl);
                }
                double sy = sd * A.grav, se = sm * A.grav;
                if (flip) { sy = -sy; se = -se; }
                // problem heating: the profile plane is ghost-filled like the (even) source array it feeds
                if (A.heat) se += sd * A.heat_rate * A.heat[(long long)i * A.pitch + jj];
                const double hy = 0.5 * A.dt * sy, he = 0.5 * A.dt * se;
                XM.ymom += hy; XP.ymom += hy; YM.ymom += hy; YP.ymom += hy;
                XM.ener += he; XP.ener += he; YM.ener += he; YP.ener += he;
            }

            // ---- H. vertex divergence for the artificial viscosity ---------------------------
            double divU = SPH
                ? vertex_divU_sph(Q(IU, i, cc), Q(IU, i, cc - 1), Q(IU, i - 1, cc), Q(IU, i - 1, cc - 1),
                                  Q(IV, i, cc), Q(IV, i, cc - 1), Q(IV, i - 1, cc), Q(IV, i - 1, cc - 1),
                                  gi(6, i), gi(7, i), gi(8, i), A.dx, g_sint, g_sinb, g_sinc, A.dy)
                : vertex_divU(Q(IU, i, cc), Q(IU, i, cc - 1), Q(IU, i - 1, cc), Q(IU, i - 1, cc - 1),
                              Q(IV, i, cc), Q(IV, i, cc - 1), Q(IV, i - 1, cc), Q(IV, i - 1, cc - 1),
                              dxinv, dyinv);
            double divU_jp1 = w.down(divU);

            // The first one / two iterations of a segment run F..M on not-yet-meaningful carried state
            // (zeros -> NaNs); nothing of it is stored and every carried value is overwritten before
            // it is used for real.  Keeping the body branch-free lets the scheduler interleave the
            // independent Riemann problems (F with J, I with L).
            Cons Fy;
            double pyF = 0.0;                               // SPH: pressure of the final y-interface state of row i-1
            {
                // ---- F. transverse x-flux at face i-1/2; dF_x of cell (i-1) -------------------
                double pxT = 0.0;
                Flux f = SPH ? riemann_p(XPc.dens, XPc.ener, XPc.xmom, XPc.ymom, XM.dens, XM.ener, XM.xmom, XM.ymom, hp, xwall && i == ng, pxT)
                             : riemann(XPc.dens, XPc.ener, XPc.xmom, XPc.ymom, XM.dens, XM.ener, XM.xmom, XM.ymom, hp, xwall && i == ng);
                Cons FxT = {f.dens, f.ener, f.mn, f.mt};
                // ---- G. transverse correction of the y-face states of cell (i-1)
                //         (unsplit_fluxes.py:463-471: U_yl[i,j+1], U_yr[i,j] -= dt/2dx * dF_x)
                Cons YMp, YPp;
                if (SPH) {
                    // fluxes times face areas over the cell volume, then the radial pressure gradient
                    // (unsplit_fluxes.py:463-490).  The "+" state of cell (i-1, j) is U_yl[i-1, j+1], which the
                    // reference corrects with the volume of cell (i-1, j+1) -- the cell on the far side of the face
                    const double a0 = area_x(i - 1), a1 = area_x(i);
                    const double hv = (0.5 * A.dt) / vol(i - 1), hv1 = (0.5 * A.dt) / vol1(i - 1);
                    const double dd = FxT.dens * a1 - FxT_prev.dens * a0, de = FxT.ener * a1 - FxT_prev.ener * a0;
                    const double dmx = FxT.xmom * a1 - FxT_prev.xmom * a0, dmy = FxT.ymom * a1 - FxT_prev.ymom * a0;
                    const double gp = (0.5 * A.dt) * (pxT - pxT_prev) / A.dx;
                    YMp.dens = YMc.dens - hv * dd; YMp.ener = YMc.ener - hv * de;
                    YMp.xmom = YMc.xmom - hv * dmx - gp; YMp.ymom = YMc.ymom - hv * dmy;
                    YPp.dens = YPc.dens - hv1 * dd; YPp.ener = YPc.ener - hv1 * de;
                    YPp.xmom = YPc.xmom - hv1 * dmx - gp; YPp.ymom = YPc.ymom - hv1 * dmy;
                    pxT_prev = pxT;
                } else {
                YMp.dens = YMc.dens - hdtdx * (FxT.dens - FxT_prev.dens);
                YMp.ener = YMc.ener - hdtdx * (FxT.ener - FxT_prev.ener);
                YMp.xmom = YMc.xmom - hdtdx * (FxT.xmom - FxT_prev.xmom);
                YMp.ymom = YMc.ymom - hdtdx * (FxT.ymom - FxT_prev.ymom);
                YPp.dens = YPc.dens - hdtdx * (FxT.dens - FxT_prev.dens);
                YPp.ener = YPc.ener - hdtdx * (FxT.ener - FxT_prev.ener);
                YPp.xmom = YPc.xmom - hdtdx * (FxT.xmom - FxT_prev.xmom);
                YPp.ymom = YPc.ymom - hdtdx * (FxT.ymom - FxT_prev.ymom);
                }
                FxT_prev = FxT;

                {
                    // ---- I. final y-flux of row i-1 at face j-1/2 (left state from lane-1) ----
                    double ld = w.up(YPp.dens), le = w.up(YPp.ener), lx = w.up(YPp.xmom), ly = w.up(YPp.ymom);
                    Flux g = SPH ? riemann_p(ld, le, ly, lx, YMp.dens, YMp.ener, YMp.ymom, YMp.xmom, hp, ywall, pyF)
                                 : riemann(ld, le, ly, lx, YMp.dens, YMp.ener, YMp.ymom, YMp.xmom, hp, ywall);
                    Fy.dens = g.dens; Fy.ener = g.ener; Fy.ymom = g.mn; Fy.xmom = g.mt;
                    // viscosity (unsplit_fluxes.py:545-547); zero on the global +y face
                    double avy = avisc_coeff(divU_prev, divU, SPH ? gi(0, i - 1) * A.dy : A.dy, A.cvisc);
                    if (A.no_avisc_yhi && j == jhi) avy = 0.0;
                    double bd = w.up(U_prev.dens), be = w.up(U_prev.ener), bx = w.up(U_prev.xmom), by = w.up(U_prev.ymom);
                    Fy.dens += avy * (bd - U_prev.dens);
                    Fy.ener += avy * (be - U_prev.ener);
                    Fy.xmom += avy * (bx - U_prev.xmom);
                    Fy.ymom += avy * (by - U_prev.ymom);
                }
            }

            // ---- J. transverse y-flux of row i at face j-1/2; dF_y of cell (i, j) -------------
            Cons XMp, XPp;
            {
                double ld = w.up(YP.dens), le = w.up(YP.ener), lx = w.up(YP.xmom), ly = w.up(YP.ymom);
                double pyT = 0.0;
                Flux g = SPH ? riemann_p(ld, le, ly, lx, YM.dens, YM.ener, YM.ymom, YM.xmom, hp, ywall, pyT)
                             : riemann(ld, le, ly, lx, YM.dens, YM.ener, YM.ymom, YM.xmom, hp, ywall);
                if (SPH) { const double ay = area_y(i); g.dens *= ay; g.ener *= ay; g.mn *= ay; g.mt *= ay; }
                // g.mn is the y-momentum flux, g.mt the x-momentum flux
                double dd = w.down(g.dens) - g.dens, de = w.down(g.ener) - g.ener;
                double dmy = w.down(g.mn) - g.mn, dmx = w.down(g.mt) - g.mt;
                // ---- K. (unsplit_fluxes.py:453-461: U_xl[i+1,j], U_xr[i,j] -= dt/2dy * dF_y)
                if (SPH) {
                    // unsplit_fluxes.py:453-461, 478-484: the "+" state of cell (i, j) is U_xl[i+1, j], corrected with
                    // the volume and theta-length of cell (i+1, j)
                    const double dp = w.down(pyT) - pyT;
                    const double hv = (0.5 * A.dt) / vol(i), hv1 = (0.5 * A.dt) / vol(i + 1);
                    XMp.dens = XM.dens - hv * dd; XMp.ener = XM.ener - hv * de;
                    XMp.xmom = XM.xmom - hv * dmx; XMp.ymom = XM.ymom - hv * dmy - (0.5 * A.dt) * dp / (gi(0, i) * A.dy);
                    XPp.dens = XP.dens - hv1 * dd; XPp.ener = XP.ener - hv1 * de;
                    XPp.xmom = XP.xmom - hv1 * dmx; XPp.ymom = XP.ymom - hv1 * dmy - (0.5 * A.dt) * dp / (gi(0, i + 1) * A.dy);
                } else {
                XMp.dens = XM.dens - hdtdy * dd; XMp.ener = XM.ener - hdtdy * de;
                XMp.xmom = XM.xmom - hdtdy * dmx; XMp.ymom = XM.ymom - hdtdy * dmy;
                XPp.dens = XP.dens - hdtdy * dd; XPp.ener = XP.ener - hdtdy * de;
                XPp.xmom = XP.xmom - hdtdy * dmx; XPp.ymom = XP.ymom - hdtdy * dmy;
                }
            }

            {
                // ---- L. final x-flux at face i-1/2 ------------------------------------------
                double pxF = 0.0;
                Flux f = SPH ? riemann_p(XPpc.dens, XPpc.ener, XPpc.xmom, XPpc.ymom,
                                         XMp.dens, XMp.ener, XMp.xmom, XMp.ymom, hp, xwall && i == ng, pxF)
                             : riemann(XPpc.dens, XPpc.ener, XPpc.xmom, XPpc.ymom,
                                       XMp.dens, XMp.ener, XMp.xmom, XMp.ymom, hp, xwall && i == ng);
                Cons Fx = {f.dens, f.ener, f.mn, f.mt};
                double avx = avisc_coeff(divU, divU_jp1, A.dx, A.cvisc);
                if (A.no_avisc_xhi && i == ihi) avx = 0.0;
                Fx.dens += avx * (U_prev.dens - Uc.dens);
                Fx.ener += avx * (U_prev.ener - Uc.ener);
                Fx.xmom += avx * (U_prev.xmom - Uc.xmom);
                Fx.ymom += avx * (U_prev.ymom - Uc.ymom);

                {
                    // ---- M. conservative update of cell (i-1, j) (simulation.py:377-384) ------
                    double fd = w.down(Fy.dens), fe = w.down(Fy.ener), fx = w.down(Fy.xmom), fy = w.down(Fy.ymom);
                    Cons Un;
                    if (SPH) {
                        // simulation.py:377-396: area-weighted flux differences over the cell volume, then the pressure
                        // gradients along r and theta
                        const double a0 = area_x(i - 1), a1 = area_x(i), b0 = area_y(i - 1), b1 = area_y1(i - 1);
                        const double dtv = A.dt / vol(i - 1);
                        Un.dens = U_prev.dens + dtv * (Fx_prev.dens * a0 - Fx.dens * a1 + Fy.dens * b0 - fd * b1);
                        Un.ener = U_prev.ener + dtv * (Fx_prev.ener * a0 - Fx.ener * a1 + Fy.ener * b0 - fe * b1);
                        Un.xmom = U_prev.xmom + dtv * (Fx_prev.xmom * a0 - Fx.xmom * a1 + Fy.xmom * b0 - fx * b1);
                        Un.ymom = U_prev.ymom + dtv * (Fx_prev.ymom * a0 - Fx.ymom * a1 + Fy.ymom * b0 - fy * b1);
                        Un.xmom -= A.dt * (pxF - pxF_prev) / A.dx;
                        Un.ymom -= A.dt * (w.down(pyF) - pyF) / (gi(0, i - 1) * A.dy);
                        // sources, predictor-corrector with the SphericalPolar branches of get_external_sources
                        // (simulation.py:398-423, :117-124, :135-146)
                        const double r = gi(0, i - 1), g = A.grav;
                        const double hp = A.heat ? A.heat_rate * A.heat[(long long)(i - 1) * A.pitch + jj] : 0.0;
                        const double so_xg = U_prev.dens * g;
                        const double so_x = so_xg + U_prev.ymom * U_prev.ymom / (U_prev.dens * r);
                        const double so_y = -U_prev.xmom * U_prev.ymom / U_prev.dens, so_e = U_prev.xmom * g + U_prev.dens * hp;
                        Un.xmom += A.dt * so_x; Un.ymom += A.dt * so_y; Un.ener += A.dt * so_e;
                        const double sn_xg = Un.dens * g;
                        const double sn_e = (Un.xmom + 0.5 * A.dt * (sn_xg - so_xg)) * g + Un.dens * hp;
                        const double sn_x = sn_xg + Un.ymom * Un.ymom / (Un.dens * r);
                        const double sn_y = -Un.xmom * Un.ymom / Un.dens;
                        Un.xmom += 0.5 * A.dt * (sn_x - so_x);
                        Un.ymom += 0.5 * A.dt * (sn_y - so_y);
                        Un.ener += 0.5 * A.dt * (sn_e - so_e);
                        pxF_prev = pxF;
                    } else {
                    Un.dens = U_prev.dens + dtdx * (Fx_prev.dens - Fx.dens) + dtdy * (Fy.dens - fd);
                    Un.ener = U_prev.ener + dtdx * (Fx_prev.ener - Fx.ener) + dtdy * (Fy.ener - fe);
                    Un.xmom = U_prev.xmom + dtdx * (Fx_prev.xmom - Fx.xmom) + dtdy * (Fy.xmom - fx);
                    Un.ymom = U_prev.ymom + dtdx * (Fx_prev.ymom - Fx.ymom) + dtdy * (Fy.ymom - fy);
                    }
                    if (GRAV && !SPH) {
                        // U += dt S(U_old); S_new from the new density and a time-centred y-momentum;
                        // U += dt/2 (S_new - S_old)
                        const double hp = A.heat ? A.heat_rate * A.heat[(long long)(i - 1) * A.pitch + jj] : 0.0;
                        const double so_y = U_prev.dens * A.grav, so_e = U_prev.ymom * A.grav + U_prev.dens * hp;
                        Un.ymom += A.dt * so_y;
                        Un.ener += A.dt * so_e;
                        const double sn_y = Un.dens * A.grav;
                        const double corr = 0.5 * A.dt * (sn_y - so_y);
                        const double sn_e = (Un.ymom + corr) * A.grav + Un.dens * hp;
                        Un.ymom += corr;
                        Un.ener += 0.5 * A.dt * (sn_e - so_e);
                    }
                    if (GRAV) {
                        if (A.do_sponge) {
                            // implicit damping of the momenta in the low-density region, kinetic-energy change
                            // booked into the energy (simulation.py:425-441)
                            const double rb = A.sponge_rho_begin, rf = A.sponge_rho_full;
                            const double f = Un.dens > rb ? 0.0 : (Un.dens < rf ? 1.0
                                             : 0.5 * (1.0 - cos(3.14159265358979323846 * (Un.dens - rb) / (rf - rb))));
                            const double damp = 1.0 / (1.0 + A.dt * (f / A.sponge_timescale));
                            const double xo = Un.xmom, yo = Un.ymom;
                            Un.xmom = xo * damp; Un.ymom = yo * damp;
                            Un.ener += 0.5 * ((Un.xmom * Un.xmom + Un.ymom * Un.ymom) - (xo * xo + yo * yo)) / Un.dens;
                        }
                    }
                    if (out_lane && i > i0) {
                        double* o = Ocol + (long long)(i - 1) * A.pitch;
                        o[0] = Un.dens; o[A.plane_stride] = Un.ener;
                        o[2 * A.plane_stride] = Un.xmom; o[3 * A.plane_stride] = Un.ymom;
                        double ax, ay;
                        cfl_speeds(Un.dens, Un.ener, Un.xmom, Un.ymom, gamma, ax, ay);
                        wmax_x = dmax(wmax_x, ax); wmax_y = dmax(wmax_y, ay);
                    }
                }
                Fx_prev = Fx;
            }

            // ---- N. carry ---------------------------------------------------------------------
            XPc = XP; XPpc = XPp; YMc = YM; YPc = YP; U_prev = Uc; divU_prev = divU;
            xix_m1 = xix_0; xix_0 = xix_p1;
#pragma unroll
            for (int n = 0; n < 4; ++n) { l2x_m1[n] = l2x_0[n]; l2x_0[n] = l2x_p1[n]; }

            // ---- O. row i-2 is dead: recycle its slot for row i+6 -----------------------------
            w.sync();
            if (i + 6 <= rlast) issue(i + 6, col0, ncols);
        }

        // every staged row row_base .. rlast was waited on once: slot s completed one phase per row with r & 7 == s
#pragma unroll
        for (int sl = 0; sl < SW_RING; ++sl) {
            const int first = row_base + ((sl - row_base) & (SW_RING - 1));      // first staged row in slot sl
            const int uses = first > rlast ? 0 : (rlast - first) / SW_RING + 1;
            phase_base ^= (unsigned)(uses & 1) << sl;
        }
        // wave-speed maxima of this strip (positive doubles order like their bit patterns)
        wmax_x = w.reduce_max(wmax_x); wmax_y = w.reduce_max(wmax_y);
        if (lane == 0) {
            w.atomic_max_bits(&A.wavemax[0], wmax_x);
            w.atomic_max_bits(&A.wavemax[1], wmax_y);
        }
    }
};

}  // namespace pyro

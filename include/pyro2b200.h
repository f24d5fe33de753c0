/*
 * pyro2b200.h -- C ABI of libpyro2b200.so: the B200 (sm_100a) implementation of pyro2's two
 * data-parallel hot paths.  Plain C, plain pointers and sizes, no torch / C++ types.
 *
 * pyro2 is pure Python: it has no FFI or plugin registry to bind against, so each
 * entry point below names the reference *Python* interface it replaces (file:line relative to the
 * pyro2 tree).  The Python host side (pyro2_b200/) mirrors those interfaces signature for
 * signature and calls these functions through ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (torch allocations); the library never
 *     frees caller memory.  Exceptions: p2b_last_error() returns a host string.
 *   - every call is asynchronous on the given cudaStream_t (passed as void*).
 *   - return 0 on success, P2B_EINVAL (-1) for a bad argument, P2B_ECUDA (-2) for a CUDA error
 *     (text via p2b_last_error()).  Invalid-state detection (the reference's assert at
 *     compressible/simulation.py:71) is reported through a device-side status word, see below.
 *   - state layout: structure of arrays.  A "state" is nvar planes; plane n starts at
 *     base + n*plane_stride; element (i, j) of a plane is at i*pitch + j, x = i is the slow axis,
 *     y = j is contiguous (the reference indexes [i, j, n]; pyro/mesh/patch.py:450-452).
 *     pitch must be even (rows 16-byte aligned for the TMA bulk copies), base 16-byte aligned.
 *   - boundary-condition codes: pyro/mesh/boundary.py names mapped as
 *       outflow / neumann -> P2B_BC_OUTFLOW, reflect-even -> P2B_BC_REFLECT_EVEN,
 *       reflect-odd / dirichlet -> P2B_BC_REFLECT_ODD, periodic -> P2B_BC_PERIODIC,
 *       P2B_BC_NONE leaves that side untouched (interior slab boundary in a decomposed run).
 */
#ifndef PYRO2B200_H
#define PYRO2B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define P2B_OK 0
#define P2B_EINVAL (-1)
#define P2B_ECUDA (-2)

enum { P2B_BC_OUTFLOW = 0, P2B_BC_REFLECT_EVEN = 1, P2B_BC_REFLECT_ODD = 2, P2B_BC_PERIODIC = 3,
       P2B_BC_NONE = 4 };

typedef struct {
    int nx, ny, ng;          /* interior zones and ghost width (pyro/mesh/patch.py:62-117) */
    int pitch;               /* elements between consecutive i rows (>= ny + 2 ng) */
    long long plane_stride;  /* elements between consecutive variables */
    double dx, dy;
} p2b_grid;

/* compressible/_defaults + eos.gamma: the runtime parameters the sweep consumes */
typedef struct {
    double gamma;
    double z0, z1, delta;    /* flattening */
    double cvisc;
    int limiter;             /* 0 none, 1 MC 2nd order, 2 MC 4th order */
    int use_flattening;
    int no_avisc_xhi;        /* 1 on the global +x face: reference leaves avisco_x unset there */
    int no_avisc_yhi;        /*   (pyro/compressible/interface.py:366-367); 0 on interior slab faces */
    double grav;             /* compressible.grav: constant acceleration along y (0 = the gravity-free kernel);
                              * sources as in pyro/compressible/simulation.py:105-160, 398-423 and
                              * unsplit_fluxes.py:247-330 */
    int src_flip_ylo;        /* 1 when the -y / +y boundary reflects: the reference fills the ghost cells of  */
    int src_flip_yhi;        /*   its source arrays odd (ymom_src) / even (E_src) there (simulation.py:248-253) */
    int riemann;             /* compressible.riemann: 0 HLLC (riemann.py:682-860), 1 CGF (riemann.py:9-310 + consFlux),
                              * 2 HLLC_lm (riemann_hllc_lowspeed, riemann.py:864-1019) */
    int xl_solid, yl_solid;  /* CGF: the -x / -y boundary is a solid wall (boundary.bc_is_solid): zero normal
                              * velocity in the interface state on that face (riemann.py:283-292) */
    double heat_rate;        /* problem heating source S_ener = dens * heat_rate * heat_profile[i, j]           */
    const double* heat_profile; /* (source_terms of compressible/problems/heating.py, plume.py, convection.py): one
                              * device plane with the state's pitch, ghost cells filled like a scalar; NULL = none */
    int do_sponge;           /* sponge damping (simulation.py:164-184, 425-441): on/off, the density below which it   */
    double sponge_rho_begin, sponge_rho_full, sponge_timescale; /* starts / is fully on, and its time scale          */
    int src_copy_yhi;        /* 1 when the +y boundary is "ambient": the ghost cells hold a constant state while the
                              * reference's source arrays are zero-gradient copies there (compressible/BC.py:150-152) */
    /* SphericalPolar grids (pyro/mesh/patch.py:242-312; x = r, y = theta; CGF only, like the reference): geo_i != NULL
     * selects the instantiation with the coord_type == 1 branches of the solver.  Device tables of doubles built on
     * the host with the reference's numpy expressions, geo_i = 9 rows of geo_ni (>= nx + 2 ng) entries, geo_j = 7 rows
     * of geo_nj (>= ny + 2 ng + 1) entries:
     *   geo_i: r_i; r of the row the reference's source arrays mirror into row i; -2 pi rl^2; rr^2 - rl^2; rr - rl;
     *          rr^2 + rl^2 + rr rl; then the viscosity's (i + 1/2 - ng) dx + xmin, (i - 1/2 - ng) dx + xmin, (i - ng) dx + xmin
     *   geo_j: cos(th_r) - cos(th_l); -2 pi / 3 times that; pi sin(th_l); tan(th);
     *          sin((j + 1/2 - ng) dy + ymin), sin((j - 1/2 - ng) dy + ymin), sin((j - ng) dy + ymin)
     * src_flip_xlo / xhi: that x boundary is "reflect" (the ghost rows of the source arrays change sign) */
    const double* geo_i;
    const double* geo_j;
    int geo_ni, geo_nj;
    int src_flip_xlo, src_flip_xhi;
} p2b_comp_params;

/* device scratch the sweep needs, 8 x 64-bit words owned by the caller:
 *   [0] bits of max(|u|+cs) over the cells written, [1] bits of max(|v|+cs)   (for the next dt)
 *   [2] task counter (work distribution), [3] status: nonzero = invalid state seen
 *   [4..7] reserved */
#define P2B_SCRATCH_WORDS 8

const char* p2b_last_error(void);
int p2b_version(void);

/* number of SMs / resident sweep warps on the current device (diagnostics for bench.py) */
int p2b_device_sms(void);

/* ---- ghost fill: ArrayIndexer.fill_ghost (pyro/mesh/array_indexer.py:150-274) applied to nvar
 * planes, as CellCenterData2d.fill_BC_all does (pyro/mesh/patch.py:575-624).  bc = nvar x 4 codes
 * (xlb, xrb, ylb, yrb per variable), host array.  x faces first, then y faces over the full x
 * range, exactly as the reference orders them.  Bit-exact for float64 and int64. */
int p2b_fill_ghost_f64(double* base, const p2b_grid* g, int nvar, const int* bc, void* stream);
int p2b_fill_ghost_i64(int64_t* base, const p2b_grid* g, int nvar, const int* bc, void* stream);

/* single plane with inhomogeneous Dirichlet / Neumann boundary values (array_indexer.py:166-183,
 * used by the finest multigrid level, pyro/multigrid/MG.py:231-242).  xl/xr have qy entries,
 * yl/yr have qx entries (device pointers) or NULL for homogeneous. */
int p2b_fill_ghost_values_f64(double* plane, const p2b_grid* g, const int bc[4], const double* xl,
                              const double* xr, const double* yl, const double* yr, void* stream);

/* ---- the compressible solver's "hse" boundary (pyro/compressible/BC.py:21-139), one variable (0 density,
 * 1 energy, 2 x-momentum, 3 y-momentum) of a 4-plane state on side 0 = ylb / 1 = yrb: zero-gradient copy for
 * all but the energy, which is integrated outward in hydrostatic equilibrium at the base density.  Called
 * per variable after the standard fill, like the reference's ext_bcs hook (pyro/mesh/patch.py:582-624);
 * bit-identical to the reference. */
int p2b_fill_hse_f64(double* U, const p2b_grid* g, double grav, double gamma, int var, int side, void* stream);
/* the "ambient" boundary (pyro/compressible/BC.py:142-168): the ghost rows of variable `var` beyond side 0 (ylb) /
 * 1 (yrb) are set to `value` (the ambient density, momenta or total energy) */
int p2b_fill_ambient_f64(double* U, const p2b_grid* g, int var, int side, double value, void* stream);

/* ---- CFL wave speeds: Simulation.method_compute_timestep (pyro/compressible/simulation.py:267-288)
 * over the FULL array including ghosts.  Accumulates (atomic max) the bit patterns of
 * max(|u|+cs), max(|v|+cs) into scratch[0], scratch[1]; the caller zeroes them first and forms
 * dt = cfl * min(dx / a, dy / b), which is bit-identical to the reference's min over cells. */
int p2b_cfl_wavemax(const double* U, const p2b_grid* g, double gamma, uint64_t* scratch, void* stream);

/* ---- the fused compressible sweep: Simulation.evolve (pyro/compressible/simulation.py:290-450)
 * with interface_states / apply_transverse_flux / apply_artificial_viscosity
 * (pyro/compressible/unsplit_fluxes.py:134-549), interface.states (interface.py:6-236),
 * riemann_hllc (riemann.py:682-860) folded into one kernel.  Uin must have its ghost cells filled;
 * the valid region of Uout (a different buffer) receives U^{n+1}; scratch[0..1] accumulate the new
 * state's wave-speed maxima, scratch[3] is set if a valid cell had rho <= 0 or e <= 0.
 * scratch[0..3] are zeroed by this call before the kernel runs.  Requires ng >= 4,
 * Riemann solver HLLC, CGF or HLLC_lm (prm->riemann); gravity, a heating profile or the sponge select the
 * instantiations with source terms; prm->geo_i selects SphericalPolar geometry (CGF only, no ambient boundary). */
int p2b_compressible_sweep(const double* Uin, double* Uout, const p2b_grid* g,
                           const p2b_comp_params* prm, double dt, uint64_t* scratch, void* stream);

/* launch geometry chosen for the last sweep (bench diagnostics): tasks, resident warps, seglen */
int p2b_sweep_info(int* ntasks, int* resident_warps, int* seglen);
/* 1: the last sweep staged its rows through a TMA tensor map (one 3-d copy per row), 0: 1-d bulk copies per plane */
int p2b_sweep_uses_tensor_map(void);

/* test probe of the sweep's branch-free fp64 helpers (csrc/hydro_core.cuh: rcp / fdiv / fsqrt -- MUFU seed + Newton,
 * no special-case handling), so their behaviour on 0, denormals, inf and the <= 2 ulp bound can be pinned on the
 * device (the host emulator has its own restatement).  op 0: out = rcp(a), 1: out = fdiv(a, b), 2: out = fsqrt(a),
 * 3: out = the HLLC_lm solver's normal-momentum flux for the face (rho, E, mn, mt) = (a[4k..4k+3]) | (b[4k..4k+3]),
 * gamma = 1.4 (riemann.py:864-1019; n counts faces for op 3); 4: out = a / b through the shared-divisor quotient of
 * cons_to_prim (must equal IEEE division), 5: out = IEEE a / b (__ddiv_rn).  Device pointers. */
int p2b_test_fastmath(int op, const double* a, const double* b, double* out, int n, void* stream);

/* ---- multigrid: CellCenterMG2d (pyro/multigrid/MG.py:77-778), constant coefficients,
 * (alpha - beta L) phi = f, nx = ny = 2^k, ng = 1.  The handle is a host object; the hierarchy's
 * device memory (p2b_mg_workspace_bytes, zero-initialised, 16-byte aligned) is allocated by the
 * caller and attached with p2b_mg_bind.  Level l has 2^(l+1) cells per side (MG.py:207-257) and three
 * planes: which = 0 v (solution / correction), 1 f (right-hand side), 2 r (residual), 3 w (scratch), each
 * (n+2) rows of p2b_mg_level_pitch elements.  bc = {xl, xr, yl, yr} codes ("dirichlet" ->
 * P2B_BC_REFLECT_ODD, "neumann" -> P2B_BC_OUTFLOW as in pyro/mesh/array_indexer.py:160-162).
 * All arithmetic is unfused and ordered as in the reference: v, f, r are bit-identical to it. */
typedef struct p2b_mg p2b_mg;

p2b_mg* p2b_mg_create(int nx, const int* bc, double alpha, double beta, double xmin, double xmax,
                      double ymin, double ymax, int nsmooth, int nsmooth_bottom);   /* MG.py:85-295 */
/* multi-GPU (extension): this process owns x-slab `rank` of `size` (power of two) on every
 * level with at least split_n columns; coarser levels are replicated on all ranks.  Slab levels store
 * p2b_mg_tb_halo() halo rows beyond each end of the owned rows; the caller exchanges them (NCCL) before
 * each p2b_mg_tb_pass / residual / prolong and all-gathers the restricted RHS at the slab -> replicated
 * transition (pyro2_b200/multigrid/MG.py does). */
p2b_mg* p2b_mg_create_slab(int nx, const int* bc, double alpha, double beta, double xmin, double xmax,
                           double ymin, double ymax, int nsmooth, int nsmooth_bottom, int rank, int size,
                           int split_n);
/* out[8] = {owned rows ni, columns n, pitch, halo rows gx, global row offset, is_slab, plane stride
 * (elements), first slab level} */
int p2b_mg_level_info(p2b_mg* m, int level, long long* out);
int p2b_mg_destroy(p2b_mg* m);
/* A/B switch (default on): temporally blocked smoother (5 red-black iterations per HBM pass) vs one
 * launch per colour; both produce identical bits */
int p2b_mg_set_blocking(p2b_mg* m, int enable);
int p2b_mg_nlevels(p2b_mg* m);
long long p2b_mg_workspace_bytes(p2b_mg* m);
int p2b_mg_bind(p2b_mg* m, void* device_mem, long long bytes);
void* p2b_mg_level_ptr(p2b_mg* m, int level, int which);
int p2b_mg_level_pitch(p2b_mg* m, int level);
/* inhomogeneous Dirichlet / Neumann values for phi on the finest level (MG.py:231-242): device
 * arrays of n+2 doubles (xl, xr indexed by j; yl, yr by i) or NULL */
int p2b_mg_set_bc_values(p2b_mg* m, const double* xl, const double* xr, const double* yl, const double* yr);

int p2b_mg_smooth(p2b_mg* m, int level, int nsmooth, void* stream);      /* smooth, MG.py:544-599 */
int p2b_mg_residual(p2b_mg* m, int level, void* stream);                 /* _compute_residual, MG.py:529-542 */
int p2b_mg_restrict(p2b_mg* m, int level, void* stream);                 /* r(level) -> f(level-1): patch.py:640-676, MG.py:731-732 */
int p2b_mg_prolong_correct(p2b_mg* m, int level, void* stream);          /* v(level) += P v(level-1), fill_BC: patch.py:678-736, MG.py:745-751 */
int p2b_mg_fill_bc(p2b_mg* m, int level, void* stream);                  /* grids[level].fill_BC("v") */
int p2b_mg_zero_coarse(p2b_mg* m, void* stream);                         /* MG.py:658-659 */
int p2b_mg_vcycle(p2b_mg* m, void* stream);                              /* v_cycle(nlevels-1), MG.py:699-778 */
int p2b_mg_vcycle_level(p2b_mg* m, int level, void* stream);             /* v_cycle(level) on non-decomposed levels */
/* one pass (1..p2b_mg_tb_iters() red-black iterations) of the temporally blocked smoother, plane src ->
 * plane dst (0 = v, 3 = scratch w; v->w or w->v) */
int p2b_mg_tb_pass(p2b_mg* m, int level, int src, int dst, int niter, void* stream);
int p2b_mg_tb_halo(void);
int p2b_mg_tb_iters(void);
/* sum over the valid region of plane^2 -> *out_dev (ArrayIndexer.norm = sqrt(dx dy sum),
 * pyro/mesh/array_indexer.py:98-111); deterministic summation order */
int p2b_mg_norm2(p2b_mg* m, int level, int which, double* out_dev, void* stream);
/* per-cycle bookkeeping of solve() (MG.py:668-686) on the finest level:
 *   out_dev[0] = sum(((v - old_phi) / (v + 1e-16))^2), old_phi <- v, r <- residual, out_dev[1] = sum(r^2)
 * old_phi is a caller-owned (n+2) x pitch buffer */
int p2b_mg_cycle_diagnostics(p2b_mg* m, double* old_phi, double* out_dev, void* stream);

/* ---- decomposed hierarchies: peer-memory communication (no NCCL inside a V-cycle) --------------------------------
 * The x-slabs of a hierarchy created with p2b_mg_create_slab talk through each other's workspaces: a kernel that
 * produces rows a neighbour needs (smoother pass, restrict, prolong) stores them straight into the neighbour's halo rows
 * and raises a flag there; the kernels that read halo rows wait on the flag (bounded spin; a time-out is reported
 * through p2b_mg_result, never a hang).  All ranks call the same sequence of p2b_mg_* functions (SPMD).  The workspaces
 * must be mapped into every rank's address space: same process -> plain pointers; several processes on one node ->
 * p2b_shared_alloc / _handle / _open (cudaIpc).  This replaces the reference's fill_BC after every smoothing colour
 * (pyro/multigrid/MG.py:565, 591-599) across slab boundaries by one message per five red-black iterations. */
int p2b_mg_set_peers(p2b_mg* m, void* const* bases);      /* bases[r]: rank r's workspace as mapped here; [rank] = own */
int p2b_mg_exchange(p2b_mg* m, int level, int which, int depth, void* stream);   /* stand-alone halo exchange (collective) */
/* the stopping rule of solve() evaluated on the device (pyro/multigrid/MG.py:654-697): lets the host enqueue cycles
 * ahead; once residual_error <= rtol or max_cycles cycles ran, the kernels of the cycles enqueued ahead return at once */
int p2b_mg_set_stop(p2b_mg* m, int enable, double source_norm, double rtol, int max_cycles, void* stream);
int p2b_mg_result(p2b_mg* m, double* out4, long long* comm_error, void* stream);   /* (relsq, rsq, residual_error, cycles); syncs */
void* p2b_mg_control_ptr(p2b_mg* m);
/* halo rows of the STATE planes of a decomposed run (HP-1 and the explicit solvers) through peer memory, and the maximum
 * of four 64-bit words over the ranks (wave speeds + status for the next dt): csrc/slab_comm.cu.  The control block and the
 * registered plane buffers are p2b_shared_alloc'd and mapped into every rank; ctl_peers / peer_bases list them as mapped
 * here ([rank] = the local one).  Replaces one grid's interior copy in ArrayIndexer.fill_ghost
 * (pyro/mesh/array_indexer.py:157-274) across slab boundaries. */
typedef struct p2b_slab p2b_slab;
long long p2b_slab_ctl_bytes(void);
p2b_slab* p2b_slab_create(int rank, int size, int periodic, void* const* ctl_peers);
int p2b_slab_destroy(p2b_slab* s);
int p2b_slab_register(p2b_slab* s, int k, void* local_base, long long bytes, void* const* peer_bases);
int p2b_slab_owns(p2b_slab* s, const void* planes);
int p2b_slab_exchange(p2b_slab* s, double* planes, int nvar, long long plane_stride, int pitch, int nx, int ng, void* stream);
int p2b_slab_allreduce_max4(p2b_slab* s, uint64_t* words, void* stream);
int p2b_slab_error(p2b_slab* s, void* stream);
void* p2b_shared_alloc(long long bytes);                  /* cudaMalloc'd + zeroed; mappable by other processes */
int p2b_shared_free(void* p);
int p2b_shared_handle(void* p, unsigned char* out64);     /* 64 opaque bytes (cudaIpcMemHandle_t) */
void* p2b_shared_open(const unsigned char* handle64);     /* maps another process's allocation, enables peer access */
int p2b_shared_close(void* p);


/* ---- the diffusion solver's use of the hierarchy (pyro/diffusion/simulation.py:62-104: Crank-Nicolson,
 * (1 - dt k/2 L) phi^{n+1} = phi^n + dt k/2 L phi^n).  p2b_mg_set_operator changes alpha / beta of an existing
 * hierarchy (the reference constructs a new CellCenterMG2d with beta = 0.5*dt*k every step);
 * p2b_mg_cn_rhs writes the right-hand side into the finest level's f plane from the solver's ghost-filled
 * phi plane ((n+2) rows of phi_pitch doubles; on a decomposed hierarchy this rank's slab, ni+2 rows with the halo rows
 * already exchanged), coef = 0.5*dt*k; bit-identical to the reference's expression. */
int p2b_mg_set_operator(p2b_mg* m, double alpha, double beta);
int p2b_mg_cn_rhs(p2b_mg* m, const double* phi, int phi_pitch, double coef, void* stream);

/* ---- variable coefficients: VarCoeffCCMG2d (pyro/multigrid/variable_coeff_MG.py:24-213) with EdgeCoeffs
 * (pyro/multigrid/edge_coeffs.py:1-54): div(eta grad phi) = f.  p2b_mg_set_coeffs does what the
 * reference's constructor does (:57-109), on the device: eta (finest level, (n+2) rows of coeffs_pitch
 * doubles, valid cells read; coeffs_bc = its boundary codes) is restricted level by level and ghost
 * filled, the finest level's edge coefficients eta_x[i,j] = eta_{i-1/2,j}/dx^2, eta_y[i,j] =
 * eta_{i,j-1/2}/dy^2 are formed and restricted down.  The planes live in caller-owned memory of
 * p2b_mg_coeff_workspace_bytes (16-byte aligned).  Afterwards smooth / residual / vcycle /
 * cycle_diagnostics apply the variable-coefficient operator (:112-213), bit-identical to the reference.
 * Single-GPU hierarchies only.  p2b_mg_coeff_ptr: which = 0 eta, 1 eta_x, 2 eta_y (level pitch). */
long long p2b_mg_coeff_workspace_bytes(p2b_mg* m);
int p2b_mg_set_coeffs(p2b_mg* m, void* device_mem, long long bytes, const double* coeffs, int coeffs_pitch,
                      const int* coeffs_bc, void* stream);
void* p2b_mg_coeff_ptr(p2b_mg* m, int level, int which);

/* ---- Burgers / incompressible explicit stages (the callers of the multigrid path).
 * pyro/burgers/burgers_interface.py:4-312, pyro/incompressible/incomp_interface.py:4-211,
 * pyro/mesh/reconstruction.py:11-120 and the array expressions of pyro/incompressible/simulation.py:67-404
 * and pyro/burgers/simulation.py:41-131.  A p2b_flow handle describes the solver grid (ng >= 4; all planes
 * share g->pitch) and owns 16 scratch planes in caller memory (p2b_flow_workspace_bytes, ZERO-INITIALISED:
 * they stand in for the reference's grid.scratch_array() temporaries).  u, v, gradp_x, gradp_y, phi are the
 * solver's state planes (device pointers, ghost cells filled by p2b_fill_ghost_f64).  Every array a stage
 * produces is bit-identical to the reference's.  Call order of one incompressible step (simulation.py:159-404):
 *   interface_states -> mac_vels -> mac_divergence -> [multigrid solve] -> mac_project -> upwind_states ->
 *   advect_update -> [ghost fill] -> cc_divergence -> [multigrid solve] -> project -> [ghost fill]          */
typedef struct p2b_flow p2b_flow;

p2b_flow* p2b_flow_create(const p2b_grid* g);
int p2b_flow_destroy(p2b_flow* f);
long long p2b_flow_workspace_bytes(p2b_flow* f);
int p2b_flow_bind(p2b_flow* f, void* device_mem, long long bytes);
/* scratch plane n: 0..7 u_xl u_xr u_yl u_yr v_xl v_xr v_yl v_yr, 8..9 transverse Riemann velocities,
 * 10..13 u_xint v_xint u_yint v_yint, 14..15 u_MAC v_MAC */
void* p2b_flow_plane(p2b_flow* f, int n);
/* limited slopes + get_interface_states + apply_transverse_corrections (+ apply_gradp_corrections unless
 * gradp_x = gradp_y = NULL, the Burgers case) */
int p2b_flow_interface_states(p2b_flow* f, const double* u, const double* v, const double* gradp_x,
                              const double* gradp_y, double dt, int limiter, void* stream);
int p2b_flow_mac_vels(p2b_flow* f, void* stream);                       /* riemann_and_upwind */
/* (u_MAC.ip(1) - u_MAC.v())/dx + (v_MAC.jp(1) - v_MAC.v())/dy -> valid cells of an (nx+2) x (ny+2) plane */
int p2b_flow_mac_divergence(p2b_flow* f, double* div, int div_pitch, void* stream);
/* u_MAC, v_MAC -= face gradient of phi_mac (solver-grid plane holding the multigrid solution in buf = 1) */
int p2b_flow_mac_project(p2b_flow* f, const double* phi_mac, void* stream);
int p2b_flow_upwind_states(p2b_flow* f, void* stream);                  /* incomp_interface.states' upwinds */
int p2b_flow_advect_update(p2b_flow* f, double* u, double* v, const double* gradp_x, const double* gradp_y,
                           double dt, int proj_type, void* stream);
/* 0.5*(u.ip(1) - u.ip(-1))/dx + 0.5*(v.jp(1) - v.jp(-1))/dy [ / dt when divide != 0 ] */
int p2b_flow_cc_divergence(p2b_flow* f, const double* u, const double* v, double* div, int div_pitch, double dt,
                           int divide, void* stream);
/* u, v -= dt * centred grad(phi); proj_type 1: gradp += grad, 2: gradp = grad, 0: gradp untouched
 * (dt = 1, proj_type = 0 is preevolve's initial projection) */
int p2b_flow_project(p2b_flow* f, const double* phi, double* u, double* v, double* gradp_x, double* gradp_y,
                     double dt, int proj_type, void* stream);
/* Burgers: construct_unsplit_fluxes + the conservative update (after interface_states and mac_vels) */
int p2b_flow_burgers_update(p2b_flow* f, double* u, double* v, double dt, void* stream);
/* linear advection, a_t + u a_x + v a_y = 0 (pyro/advection/interface.py, advective_fluxes.py:5-92,
 * advection/simulation.py:56-92): one step for the ghost-filled scalar plane a (uses scratch planes 0..3) */
int p2b_flow_advection_update(p2b_flow* f, double* a, double u, double v, double dt, int limiter, void* stream);
/* bit patterns of max|u|, max|v| over the full arrays (atomic max into scratch[0..1], zeroed by the caller) */
int p2b_flow_maxabs(p2b_flow* f, const double* u, const double* v, uint64_t* scratch, void* stream);

/* ---- low Mach number atmosphere (lm_atm) explicit stages: the third caller of the multigrid path, through
 * the variable-coefficient solver.  pyro/lm_atm/LM_atm_interface.py:181-703 (the numba routines mac_vels, states,
 * rho_states, get_interface_states, upwind, riemann) and the array expressions of
 * pyro/lm_atm/simulation.py:138-618.  A p2b_lm handle describes the solver grid (ng >= 4), points at the device
 * copy of the 1-d base state (rho0, p0, beta0, beta0-edges: 4 x (ny + 2 ng) doubles, simulation.py:102-133) and owns
 * 25 ZERO-INITIALISED scratch planes.  Planes 16 ("coeff") and 17 ("source_y") are the reference's aux_data: the
 * caller ghost-fills them (p2b_fill_ghost_f64 with the density / y-velocity boundary types) between the calls that
 * write and read them, exactly where the reference calls aux_data.fill_BC.  Every array is bit-identical to the
 * reference's.  One step (simulation.py:286-618):
 *   coeff(rho; beta0) + fill, source(rho) + fill -> interface_states -> mac_vels -> coeff(rho; beta0^2, buf 1) ->
 *   mac_divergence -> [VarCoeff multigrid solve] -> coeff(rho; beta0) + fill -> mac_project -> density_update ->
 *   [fill density] -> coeff(rho + rho_old, 2; beta0) + fill -> interface_states -> upwind_states -> advect_update ->
 *   source(rho, rho_old) + fill -> add_source -> [fill u, v] -> coeff(rho; beta0^2) -> cc_divergence ->
 *   [VarCoeff multigrid solve] -> project -> [fill u, v, gradp]                                                  */
typedef struct p2b_lm p2b_lm;

p2b_lm* p2b_lm_create(const p2b_grid* g, const double* basestate);
int p2b_lm_destroy(p2b_lm* h);
long long p2b_lm_workspace_bytes(p2b_lm* h);
int p2b_lm_bind(p2b_lm* h, void* device_mem, long long bytes);
/* scratch plane n: 0..7 u/v interface states, 8..9 transverse Riemann velocities, 10..13 u_xint v_xint u_yint v_yint,
 * 14..15 u_MAC v_MAC, 16 coeff, 17 source_y, 18..21 density interface states, 22..23 rho_xint rho_yint, 24 rho_old */
void* p2b_lm_plane(p2b_lm* h, int n);
/* coeff <- numer / (d1 [+ d2]) * beta0 (squared: * beta0**2) over the buf-extended valid region */
int p2b_lm_coeff(p2b_lm* h, const double* d1, const double* d2, double numer, int squared, int buf, void* stream);
/* source_y <- rho' g / rho: valid cells from rho, or (rho_old given) the whole array from 0.5 (rho + rho_old) */
int p2b_lm_source(p2b_lm* h, const double* rho, const double* rho_old, double grav, void* stream);
int p2b_lm_interface_states(p2b_lm* h, const double* u, const double* v, const double* gradp_x, const double* gradp_y,
                            double dt, int limiter, void* stream);
int p2b_lm_mac_vels(p2b_lm* h, void* stream);
int p2b_lm_mac_divergence(p2b_lm* h, double* div, int div_pitch, void* stream);      /* D(beta0 U_MAC) */
int p2b_lm_mac_project(p2b_lm* h, const double* phi_mac, void* stream);
/* rho_states + conservative density update + eint = p0/(gamma-1)/rho; keeps rho_old in plane 24 */
int p2b_lm_density_update(p2b_lm* h, double* rho, double* eint, double dt, int limiter, double gamma, void* stream);
int p2b_lm_upwind_states(p2b_lm* h, void* stream);
int p2b_lm_advect_update(p2b_lm* h, double* u, double* v, const double* gradp_x, const double* gradp_y, double dt,
                         int proj_type, void* stream);
int p2b_lm_add_source(p2b_lm* h, double* v, double dt, void* stream);                 /* v += dt * source_y */
int p2b_lm_cc_divergence(p2b_lm* h, const double* u, const double* v, double* div, int div_pitch, double dt, int divide,
                         void* stream);                                                 /* D(beta0 U) [/ dt] */
/* U -= dt (beta0/rho) G phi in the valid cells; proj_type 0: gradp untouched (dt = 1: preevolve), 1: +=, 2: = */
int p2b_lm_project(p2b_lm* h, const double* rho, const double* phi, double* u, double* v, double* gradp_x,
                   double* gradp_y, double dt, int proj_type, void* stream);
/* scratch[0..4] (zeroed by the caller) <- bit patterns of max|u|, max|v| over the whole arrays and of max|u|, max|v|,
 * max(|rho' g| / rho) over the valid cells: the inputs of method_compute_timestep (simulation.py:138-178) */
int p2b_lm_reduce(p2b_lm* h, const double* rho, const double* u, const double* v, double grav, uint64_t* scratch,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif

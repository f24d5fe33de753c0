// flow.cu -- host side of the Burgers / incompressible explicit stages (p2b_flow_* in include/pyro2b200.h):
// a handle that owns the geometry and the scratch planes the reference allocates with
// grid.scratch_array() on every call (eight interface states, the transverse-step intermediates, the MAC
// velocities), and one entry point per stage of incompressible/simulation.py:evolve.  The multigrid
// projections between the stages are the p2b_mg_* calls; the Python Simulation strings them together in
// the reference's order.  Kernels: flow_kernels.cuh.
#include "common.cuh"
#include "flow_kernels.cuh"

struct p2b_flow {
    pyro::FlowGeom g;
    long long plane;          // elements per scratch plane
    double* base;
    pyro::FlowFaces S;
    pyro::FlowHat H;
    double *umac, *vmac;
};

namespace pyro {

constexpr int FLOW_NPLANES = 16;

static dim3 flow_block() { return dim3(64, 4); }

static dim3 flow_grid(const FlowGeom& g, int buf_lo, int buf_hi)
{
    // cells per direction of a region that extends buf_lo below and buf_hi above the valid cells
    dim3 b = flow_block();
    const int ni = g.nx + buf_lo + buf_hi, nj = g.ny + buf_lo + buf_hi;
    return dim3((nj + b.x - 1) / b.x, (ni + b.y - 1) / b.y);
}

}  // namespace pyro

using namespace pyro;

extern "C" {

#define FLOW_CHECK(f) P2B_REQUIRE((f) && (f)->base, "flow handle not bound")

// grid_setup(rp, ng=4) (pyro/simulation_null.py:20-60): the solver grid the stages work on
p2b_flow* p2b_flow_create(const p2b_grid* g)
{
    if (!g || g->nx < 1 || g->ny < 1 || g->ng < 4 || g->pitch < g->ny + 2 * g->ng) {
        set_error("flow: bad grid (needs ng >= 4, pitch >= ny + 2 ng)");
        return nullptr;
    }
    p2b_flow* f = new p2b_flow();
    memset(f, 0, sizeof *f);
    f->g.nx = g->nx; f->g.ny = g->ny; f->g.ng = g->ng; f->g.pitch = g->pitch;
    f->g.qx = g->nx + 2 * g->ng; f->g.qy = g->ny + 2 * g->ng;
    f->g.dx = g->dx; f->g.dy = g->dy;
    f->plane = (long long)f->g.qx * g->pitch;
    return f;
}

int p2b_flow_destroy(p2b_flow* f) { delete f; return P2B_OK; }

long long p2b_flow_workspace_bytes(p2b_flow* f) { return f ? FLOW_NPLANES * f->plane * 8 : 0; }

// the workspace must be zero-initialised: entries no stage writes stay zero, like a scratch_array()
int p2b_flow_bind(p2b_flow* f, void* mem, long long bytes)
{
    P2B_REQUIRE(f && mem, "null pointer");
    P2B_REQUIRE(bytes >= p2b_flow_workspace_bytes(f), "workspace too small");
    f->base = (double*)mem;
    double** slots[FLOW_NPLANES] = {&f->S.u_xl, &f->S.u_xr, &f->S.u_yl, &f->S.u_yr, &f->S.v_xl, &f->S.v_xr,
                                    &f->S.v_yl, &f->S.v_yr, &f->H.uhat, &f->H.vhat, &f->H.uxi, &f->H.vxi,
                                    &f->H.uyi, &f->H.vyi, &f->umac, &f->vmac};
    for (int n = 0; n < FLOW_NPLANES; ++n) *slots[n] = f->base + n * f->plane;
    return P2B_OK;
}

// scratch plane n (order: u_xl u_xr u_yl u_yr v_xl v_xr v_yl v_yr uhat vhat u_xint v_xint u_yint v_yint
// u_MAC v_MAC), qx rows of `pitch` doubles
void* p2b_flow_plane(p2b_flow* f, int n)
{
    if (!f || !f->base || n < 0 || n >= FLOW_NPLANES) return nullptr;
    return f->base + n * f->plane;
}

// reconstruction.limit x4 + get_interface_states + apply_transverse_corrections (+ apply_gradp_corrections
// when gradp_x / gradp_y are given): the eight left / right states both mac_vels and states start from
// (incomp_interface.py:38-62 and :105-129 compute the identical arrays twice; here they are kept)
int p2b_flow_interface_states(p2b_flow* f, const double* u, const double* v, const double* gradp_x,
                              const double* gradp_y, double dt, int limiter, void* stream)
{
    FLOW_CHECK(f);
    P2B_REQUIRE(u && v, "null velocity");
    P2B_REQUIRE((gradp_x == nullptr) == (gradp_y == nullptr), "give both pressure-gradient planes or neither");
    P2B_REQUIRE(limiter >= 0 && limiter <= 2, "limiter must be 0, 1 or 2");
    cudaStream_t st = (cudaStream_t)stream;
    const FlowGeom& g = f->g;
    const double dtdx = dt / g.dx, dtdy = dt / g.dy;
    const dim3 grd = flow_grid(g, 2, 2), blk = flow_block();
    P2B_LAUNCH(flow_states_kernel, grd, blk, 0, st)(g, u, v, f->S, dtdx, dtdy, limiter);
    P2B_LAUNCH(flow_hat_kernel, grd, blk, 0, st)(g, f->S, f->H);
    P2B_LAUNCH(flow_correct_kernel, grd, blk, 0, st)(g, f->S, f->H, gradp_x, gradp_y, dtdx, dtdy, dt);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

// riemann_and_upwind of u on x faces and v on y faces -> u_MAC, v_MAC (scratch planes 14, 15)
int p2b_flow_mac_vels(p2b_flow* f, void* stream)
{
    FLOW_CHECK(f);
    P2B_LAUNCH(flow_mac_kernel, flow_grid(f->g, 2, 2), flow_block(), 0, (cudaStream_t)stream)(f->g, f->S, f->umac, f->vmac);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

// divergence of the MAC velocities -> valid cells of an (nx+2) x (ny+2) multigrid-grid plane
int p2b_flow_mac_divergence(p2b_flow* f, double* div, int div_pitch, void* stream)
{
    FLOW_CHECK(f);
    P2B_REQUIRE(div && div_pitch >= f->g.ny + 2, "bad divergence plane");
    P2B_LAUNCH(flow_mac_div_kernel, flow_grid(f->g, 0, 0), flow_block(), 0, (cudaStream_t)stream)(f->g, f->umac, f->vmac, div, div_pitch);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

// u_MAC, v_MAC -= grad(phi_MAC) on the faces; phi_mac = the solver-grid plane (buf = 1 region current)
int p2b_flow_mac_project(p2b_flow* f, const double* phi_mac, void* stream)
{
    FLOW_CHECK(f);
    P2B_REQUIRE(phi_mac, "null phi");
    P2B_LAUNCH(flow_mac_project_kernel, flow_grid(f->g, 0, 1), flow_block(), 0, (cudaStream_t)stream)(f->g, phi_mac, f->umac, f->vmac);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

// the four interface velocities upwinded with the MAC velocities (scratch planes 10..13)
int p2b_flow_upwind_states(p2b_flow* f, void* stream)
{
    FLOW_CHECK(f);
    P2B_LAUNCH(flow_upwind_kernel, flow_grid(f->g, 2, 2), flow_block(), 0, (cudaStream_t)stream)(f->g, f->S, f->umac, f->vmac, f->H);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

// u, v -= dt * (advective terms [+ gradp for proj_type 1]) in the valid cells
int p2b_flow_advect_update(p2b_flow* f, double* u, double* v, const double* gradp_x, const double* gradp_y,
                           double dt, int proj_type, void* stream)
{
    FLOW_CHECK(f);
    P2B_REQUIRE(u && v, "null velocity");
    P2B_REQUIRE(proj_type == 1 || proj_type == 2, "proj_type must be 1 or 2");
    P2B_REQUIRE(proj_type == 2 || (gradp_x && gradp_y), "proj_type 1 needs the pressure gradient");
    P2B_LAUNCH(flow_advect_kernel, flow_grid(f->g, 0, 0), flow_block(), 0, (cudaStream_t)stream)(f->g, f->umac, f->vmac, f->H, u, v, gradp_x, gradp_y, dt, proj_type);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

// cell-centred divergence of (u, v) -> valid cells of a multigrid-grid plane; divide != 0: divided by dt
int p2b_flow_cc_divergence(p2b_flow* f, const double* u, const double* v, double* div, int div_pitch, double dt,
                           int divide, void* stream)
{
    FLOW_CHECK(f);
    P2B_REQUIRE(u && v && div && div_pitch >= f->g.ny + 2, "bad arguments");
    P2B_LAUNCH(flow_cc_div_kernel, flow_grid(f->g, 0, 0), flow_block(), 0, (cudaStream_t)stream)(f->g, u, v, div, div_pitch, dt, divide);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

// u, v -= dt * grad(phi) (centred, valid cells; phi's buf = 1 region current) and the gradp bookkeeping:
// proj_type 1: gradp += grad(phi); 2: gradp = grad(phi) (zero in the ghost cells); 0: gradp untouched
int p2b_flow_project(p2b_flow* f, const double* phi, double* u, double* v, double* gradp_x, double* gradp_y,
                     double dt, int proj_type, void* stream)
{
    FLOW_CHECK(f);
    P2B_REQUIRE(phi && u && v, "null pointer");
    P2B_REQUIRE(proj_type >= 0 && proj_type <= 2, "proj_type must be 0, 1 or 2");
    P2B_REQUIRE(proj_type == 0 || (gradp_x && gradp_y), "gradp planes needed");
    const FlowGeom& g = f->g;
    dim3 blk = flow_block();
    dim3 grd((g.qy + blk.x - 1) / blk.x, (g.qx + blk.y - 1) / blk.y);
    P2B_LAUNCH(flow_project_kernel, grd, blk, 0, (cudaStream_t)stream)(g, phi, u, v, gradp_x, gradp_y, dt, proj_type);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

// Burgers: fluxes from the interface states + MAC velocities, then the conservative update of u, v
int p2b_flow_burgers_update(p2b_flow* f, double* u, double* v, double dt, void* stream)
{
    FLOW_CHECK(f);
    P2B_REQUIRE(u && v, "null velocity");
    cudaStream_t st = (cudaStream_t)stream;
    const FlowGeom& g = f->g;
    P2B_LAUNCH(flow_burgers_flux_kernel, flow_grid(g, 2, 2), flow_block(), 0, st)(g, f->S, f->umac, f->vmac, f->H);
    P2B_LAUNCH(flow_burgers_update_kernel, flow_grid(g, 0, 0), flow_block(), 0, st)(g, f->H, u, v, dt / g.dx, dt / g.dy);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

// linear advection of one scalar with constant (u, v): interface states, transverse-corrected fluxes,
// conservative update of the valid cells (uses scratch planes 0..3; a handle used for advection must not be
// shared with the Burgers / incompressible stages, which write other index ranges of those planes)
int p2b_flow_advection_update(p2b_flow* f, double* a, double u, double v, double dt, int limiter, void* stream)
{
    FLOW_CHECK(f);
    P2B_REQUIRE(a, "null scalar");
    P2B_REQUIRE(limiter >= 0 && limiter <= 2, "limiter must be 0, 1 or 2");
    cudaStream_t st = (cudaStream_t)stream;
    const FlowGeom& g = f->g;
    double *ax = f->S.u_xl, *ay = f->S.u_xr, *fx = f->S.u_yl, *fy = f->S.u_yr;
    const dim3 blk = flow_block();
    P2B_LAUNCH(flow_adv_states_kernel, flow_grid(g, 1, 1), blk, 0, st)(g, a, ax, ay, u, v, u * dt / g.dx, v * dt / g.dy, limiter);
    P2B_LAUNCH(flow_adv_flux_kernel, flow_grid(g, 1, 1), blk, 0, st)(g, ax, ay, fx, fy, u, v, 0.5 * dt / g.dx, 0.5 * dt / g.dy);
    P2B_LAUNCH(flow_adv_update_kernel, flow_grid(g, 0, 0), blk, 0, st)(g, a, fx, fy, dt / g.dx, dt / g.dy);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

// bit patterns of max|u|, max|v| over the full arrays accumulated (atomic max) into scratch[0], scratch[1];
// the caller zeroes them and forms dt = cfl * min(dx / max(umax, SMALL), dy / max(vmax, SMALL))
int p2b_flow_maxabs(p2b_flow* f, const double* u, const double* v, uint64_t* scratch, void* stream)
{
    P2B_REQUIRE(f && u && v && scratch, "null pointer");
    long long blocks = ((long long)f->g.qx * f->g.qy + 255) / 256;
    if (blocks > 8LL * num_sms()) blocks = 8LL * num_sms();
    P2B_LAUNCH(flow_maxabs_kernel, (int)blocks, 256, 0, (cudaStream_t)stream)(f->g, u, v, (unsigned long long*)scratch);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

}  // extern "C"

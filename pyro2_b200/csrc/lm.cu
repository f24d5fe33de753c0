// lm.cu -- host side of the low Mach number atmosphere solver's explicit stages (p2b_lm_* in
// include/pyro2b200.h).  A handle owns the geometry, the device copy of the 1-d base state and the scratch
// planes that stand in for the reference's per-call arrays; one entry point per block of
// lm_atm/simulation.py:evolve.  The ghost fills of the state and of the two auxiliary planes ("coeff",
// "source_y") and the variable-coefficient multigrid projections between the calls are issued by the Python
// Simulation in the reference's order.  Kernels: lm_kernels.cuh (shared pieces: flow_kernels.cuh).
#include "common.cuh"
#include "lm_kernels.cuh"

struct p2b_lm {
    pyro::FlowGeom g;
    long long plane;
    double* base;             // scratch planes
    const double* basestate;  // device: rho0, p0, beta0, beta0-edges, qy doubles each
    pyro::LmBase B;
    pyro::FlowFaces S;        // u / v interface states
    pyro::FlowHat H;          // transverse Riemann velocities and upwinded states (final: upwinded with U_MAC)
    double *umac, *vmac;
    double *coeff, *source;   // the reference's aux_data planes
    double *rxl, *rxr, *ryl, *ryr, *rxi, *ryi, *rho_old;
};

namespace pyro {

constexpr int LM_NPLANES = 25;

static dim3 lm_block() { return dim3(64, 4); }

static dim3 lm_grid(const FlowGeom& g, int lo, int hi)
{
    dim3 b = lm_block();
    return dim3((g.ny + lo + hi + b.x - 1) / b.x, (g.nx + lo + hi + b.y - 1) / b.y);
}

static dim3 lm_grid_full(const FlowGeom& g)
{
    dim3 b = lm_block();
    return dim3((g.qy + b.x - 1) / b.x, (g.qx + b.y - 1) / b.y);
}

}  // namespace pyro

using namespace pyro;

extern "C" {

#define LM_CHECK(h) P2B_REQUIRE((h) && (h)->base, "lm handle not bound")

// g: the solver grid (ng >= 4); basestate: device array of 4 * (ny + 2 ng) doubles: rho0, p0, beta0, beta0-edges
// (lm_atm/simulation.py:102-133)
p2b_lm* p2b_lm_create(const p2b_grid* g, const double* basestate)
{
    if (!g || !basestate || g->nx < 1 || g->ny < 1 || g->ng < 4 || g->pitch < g->ny + 2 * g->ng) {
        set_error("lm: bad grid (needs ng >= 4, pitch >= ny + 2 ng) or null base state");
        return nullptr;
    }
    p2b_lm* h = new p2b_lm();
    memset(h, 0, sizeof *h);
    h->g.nx = g->nx; h->g.ny = g->ny; h->g.ng = g->ng; h->g.pitch = g->pitch;
    h->g.qx = g->nx + 2 * g->ng; h->g.qy = g->ny + 2 * g->ng;
    h->g.dx = g->dx; h->g.dy = g->dy;
    h->plane = (long long)h->g.qx * g->pitch;
    h->basestate = basestate;
    h->B.rho0 = basestate; h->B.p0 = basestate + h->g.qy;
    h->B.beta0 = basestate + 2 * h->g.qy; h->B.beta0e = basestate + 3 * h->g.qy;
    return h;
}

int p2b_lm_destroy(p2b_lm* h) { delete h; return P2B_OK; }

long long p2b_lm_workspace_bytes(p2b_lm* h) { return h ? LM_NPLANES * h->plane * 8 : 0; }

// zero-initialised workspace: entries a stage never writes stay zero like the reference's np.zeros arrays
int p2b_lm_bind(p2b_lm* h, void* mem, long long bytes)
{
    P2B_REQUIRE(h && mem, "null pointer");
    P2B_REQUIRE(bytes >= p2b_lm_workspace_bytes(h), "workspace too small");
    h->base = (double*)mem;
    double** slots[LM_NPLANES] = {&h->S.u_xl, &h->S.u_xr, &h->S.u_yl, &h->S.u_yr, &h->S.v_xl, &h->S.v_xr, &h->S.v_yl,
                                  &h->S.v_yr, &h->H.uhat, &h->H.vhat, &h->H.uxi, &h->H.vxi, &h->H.uyi, &h->H.vyi,
                                  &h->umac, &h->vmac, &h->coeff, &h->source, &h->rxl, &h->rxr, &h->ryl, &h->ryr,
                                  &h->rxi, &h->ryi, &h->rho_old};
    for (int n = 0; n < LM_NPLANES; ++n) *slots[n] = h->base + n * h->plane;
    return P2B_OK;
}

// scratch plane n: 0..7 u/v interface states, 8..9 transverse Riemann velocities, 10..13 u_xint v_xint u_yint
// v_yint, 14..15 u_MAC v_MAC, 16 coeff, 17 source_y, 18..21 rho interface states, 22..23 rho_xint rho_yint,
// 24 rho_old
void* p2b_lm_plane(p2b_lm* h, int n)
{
    if (!h || !h->base || n < 0 || n >= LM_NPLANES) return nullptr;
    return h->base + n * h->plane;
}

// coeff <- numer / (d1 [+ d2]) * beta0 (squared != 0: * beta0**2) over v(buf) of the "coeff" plane
int p2b_lm_coeff(p2b_lm* h, const double* d1, const double* d2, double numer, int squared, int buf, void* stream)
{
    LM_CHECK(h);
    P2B_REQUIRE(d1 && buf >= 0 && buf <= h->g.ng, "bad arguments");
    P2B_LAUNCH(lm_coeff_kernel, lm_grid(h->g, buf, buf), lm_block(), 0, (cudaStream_t)stream)(h->g, d1, d2, numer, h->B.beta0, squared, buf, h->coeff);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

// source_y <- rho' g / rho: valid cells from rho (rho_old = NULL), or the whole array from 0.5 (rho + rho_old)
int p2b_lm_source(p2b_lm* h, const double* rho, const double* rho_old, double grav, void* stream)
{
    LM_CHECK(h);
    P2B_REQUIRE(rho, "null density");
    P2B_LAUNCH(lm_source_kernel, lm_grid_full(h->g), lm_block(), 0, (cudaStream_t)stream)(h->g, rho, rho_old, h->B.rho0, grav, h->source);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

// get_interface_states: slopes + predictor, Riemann / upwind, transverse + (coeff * gradp) + buoyancy terms;
// uses the ghost-filled "coeff" and "source_y" planes
int p2b_lm_interface_states(p2b_lm* h, const double* u, const double* v, const double* gradp_x, const double* gradp_y,
                            double dt, int limiter, void* stream)
{
    LM_CHECK(h);
    P2B_REQUIRE(u && v && gradp_x && gradp_y, "null pointer");
    P2B_REQUIRE(limiter >= 0 && limiter <= 2, "limiter must be 0, 1 or 2");
    cudaStream_t st = (cudaStream_t)stream;
    const FlowGeom& g = h->g;
    const double dtdx = dt / g.dx, dtdy = dt / g.dy;
    P2B_LAUNCH(flow_states_kernel, lm_grid(g, 2, 2), lm_block(), 0, st)(g, u, v, h->S, dtdx, dtdy, limiter);
    P2B_LAUNCH(lm_hat_kernel, lm_grid(g, 1, 2), lm_block(), 0, st)(g, h->S, h->H);
    P2B_LAUNCH(lm_correct_kernel, lm_grid(g, 1, 1), lm_block(), 0, st)(g, h->S, h->H, h->coeff, gradp_x, gradp_y, h->source, dtdx, dtdy, dt);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

int p2b_lm_mac_vels(p2b_lm* h, void* stream)
{
    LM_CHECK(h);
    P2B_LAUNCH(lm_mac_kernel, lm_grid(h->g, 1, 2), lm_block(), 0, (cudaStream_t)stream)(h->g, h->S, h->umac, h->vmac);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

int p2b_lm_mac_divergence(p2b_lm* h, double* div, int div_pitch, void* stream)
{
    LM_CHECK(h);
    P2B_REQUIRE(div && div_pitch >= h->g.ny + 2, "bad divergence plane");
    P2B_LAUNCH(lm_mac_div_kernel, lm_grid(h->g, 0, 0), lm_block(), 0, (cudaStream_t)stream)(h->g, h->B, h->umac, h->vmac, div, div_pitch);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

// U_MAC -= (beta0/rho at the faces) G phi_MAC; the "coeff" plane must hold the ghost-filled beta0/rho
int p2b_lm_mac_project(p2b_lm* h, const double* phi_mac, void* stream)
{
    LM_CHECK(h);
    P2B_REQUIRE(phi_mac, "null phi");
    P2B_LAUNCH(lm_mac_project_kernel, lm_grid(h->g, 0, 1), lm_block(), 0, (cudaStream_t)stream)(h->g, h->coeff, phi_mac, h->umac, h->vmac);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

// rho_states + the conservative density update + eint (rho_old is kept in scratch plane 24)
int p2b_lm_density_update(p2b_lm* h, double* rho, double* eint, double dt, int limiter, double gamma, void* stream)
{
    LM_CHECK(h);
    P2B_REQUIRE(rho && eint, "null pointer");
    P2B_REQUIRE(limiter >= 0 && limiter <= 2, "limiter must be 0, 1 or 2");
    cudaStream_t st = (cudaStream_t)stream;
    const FlowGeom& g = h->g;
    P2B_LAUNCH(lm_rho_pred_kernel, lm_grid(g, 2, 2), lm_block(), 0, st)(g, rho, h->umac, h->vmac, h->rxl, h->rxr, h->ryl, h->ryr, dt / g.dx, dt / g.dy, limiter);
    P2B_LAUNCH(lm_upwind2_kernel, lm_grid(g, 1, 2), lm_block(), 0, st)(g, h->rxl, h->rxr, h->umac, h->rxi, h->ryl, h->ryr, h->vmac, h->ryi);
    P2B_LAUNCH(lm_rho_trans_kernel, lm_grid(g, 2, 2), lm_block(), 0, st)(g, rho, h->umac, h->vmac, h->rxi, h->ryi, h->rxl, h->rxr, h->ryl, h->ryr, dt);
    P2B_LAUNCH(lm_upwind2_kernel, lm_grid(g, 1, 2), lm_block(), 0, st)(g, h->rxl, h->rxr, h->umac, h->rxi, h->ryl, h->ryr, h->vmac, h->ryi);
    P2B_CUDA_CHECK(cudaMemcpyAsync(h->rho_old, rho, (size_t)h->plane * sizeof(double), cudaMemcpyDeviceToDevice, st));
    P2B_LAUNCH(lm_rho_update_kernel, lm_grid(g, 0, 0), lm_block(), 0, st)(g, h->B, rho, eint, h->umac, h->vmac, h->rxi, h->ryi, dt, gamma);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

// the four interface velocities upwinded with the MAC velocities (LM_atm_interface.py:324-329)
int p2b_lm_upwind_states(p2b_lm* h, void* stream)
{
    LM_CHECK(h);
    cudaStream_t st = (cudaStream_t)stream;
    const FlowGeom& g = h->g;
    P2B_LAUNCH(lm_upwind2_kernel, lm_grid(g, 1, 2), lm_block(), 0, st)(g, h->S.u_xl, h->S.u_xr, h->umac, h->H.uxi, h->S.v_xl, h->S.v_xr, h->umac, h->H.vxi);
    P2B_LAUNCH(lm_upwind2_kernel, lm_grid(g, 1, 2), lm_block(), 0, st)(g, h->S.u_yl, h->S.u_yr, h->vmac, h->H.uyi, h->S.v_yl, h->S.v_yr, h->vmac, h->H.vyi);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

// u, v -= dt * (advective terms [+ gradp for proj_type 1]) in the valid cells (simulation.py:468-489)
int p2b_lm_advect_update(p2b_lm* h, double* u, double* v, const double* gradp_x, const double* gradp_y, double dt,
                         int proj_type, void* stream)
{
    LM_CHECK(h);
    P2B_REQUIRE(u && v && gradp_x && gradp_y, "null pointer");
    P2B_REQUIRE(proj_type == 1 || proj_type == 2, "proj_type must be 1 or 2");
    P2B_LAUNCH(flow_advect_kernel, lm_grid(h->g, 0, 0), lm_block(), 0, (cudaStream_t)stream)(h->g, h->umac, h->vmac, h->H, u, v, gradp_x, gradp_y, dt, proj_type);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

// v[:, :] += dt * source_y over the whole array
int p2b_lm_add_source(p2b_lm* h, double* v, double dt, void* stream)
{
    LM_CHECK(h);
    P2B_REQUIRE(v, "null pointer");
    P2B_LAUNCH(lm_add_source_kernel, lm_grid_full(h->g), lm_block(), 0, (cudaStream_t)stream)(h->g, v, h->source, dt);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

int p2b_lm_cc_divergence(p2b_lm* h, const double* u, const double* v, double* div, int div_pitch, double dt, int divide,
                         void* stream)
{
    LM_CHECK(h);
    P2B_REQUIRE(u && v && div && div_pitch >= h->g.ny + 2, "bad arguments");
    P2B_LAUNCH(lm_cc_div_kernel, lm_grid(h->g, 0, 0), lm_block(), 0, (cudaStream_t)stream)(h->g, h->B, u, v, div, div_pitch, dt, divide);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

// U -= dt (beta0/rho) G phi, gradp bookkeeping (proj_type 0: untouched, 1: +=, 2: =), valid cells
int p2b_lm_project(p2b_lm* h, const double* rho, const double* phi, double* u, double* v, double* gradp_x,
                   double* gradp_y, double dt, int proj_type, void* stream)
{
    LM_CHECK(h);
    P2B_REQUIRE(rho && phi && u && v, "null pointer");
    P2B_REQUIRE(proj_type >= 0 && proj_type <= 2, "proj_type must be 0, 1 or 2");
    P2B_REQUIRE(proj_type == 0 || (gradp_x && gradp_y), "gradp planes needed");
    P2B_LAUNCH(lm_project_kernel, lm_grid(h->g, 0, 0), lm_block(), 0, (cudaStream_t)stream)(h->g, h->B, rho, phi, u, v, gradp_x, gradp_y, dt, proj_type);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

// bit patterns of max|u|, max|v| (whole arrays), max|u|, max|v|, max(|rho' g| / rho) (valid cells) accumulated
// (atomic max) into scratch[0..4], zeroed by the caller
int p2b_lm_reduce(p2b_lm* h, const double* rho, const double* u, const double* v, double grav, uint64_t* scratch,
                  void* stream)
{
    P2B_REQUIRE(h && rho && u && v && scratch, "null pointer");
    long long blocks = ((long long)h->g.qx * h->g.qy + 255) / 256;
    if (blocks > 8LL * num_sms()) blocks = 8LL * num_sms();
    P2B_LAUNCH(lm_reduce_kernel, (int)blocks, 256, 0, (cudaStream_t)stream)(h->g, rho, u, v, h->B.rho0, grav, (unsigned long long*)scratch);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

}  // extern "C"

// flow_kernels.cuh -- device code of the explicit (non-multigrid) part of the Burgers and incompressible
// solvers: the callers either side of the multigrid hot path.  No runtime-API calls;
// included by flow.cu (nvcc) and by tests/emu/flow_emu.cpp (g++, through tests/emu/cuda_emu.h).
//
// Reference behaviour (pyro2, file:line):
//   reconstruction.limit (nolimit / limit2 / limit4)   pyro/mesh/reconstruction.py:11-120
//   burgers_interface.get_interface_states             pyro/burgers/burgers_interface.py:4-79
//   burgers_interface.apply_transverse_corrections                                   :82-157
//   burgers_interface.construct_unsplit_fluxes                                       :160-225
//   burgers_interface.upwind / riemann / riemann_and_upwind                          :228-312
//   burgers Simulation.evolve / method_compute_timestep pyro/burgers/simulation.py:41-131
//   incomp_interface.mac_vels / states                 pyro/incompressible/incomp_interface.py:4-158
//   incomp_interface.apply_gradp_corrections                                         :161-211
//   incompressible Simulation.evolve / preevolve       pyro/incompressible/simulation.py:67-404
//
// The reference evaluates these as whole-array numpy expressions on zero-initialised scratch arrays;
// every kernel below covers exactly the index range of the slice expression it replaces and performs
// the same individually rounded operations in the same order (exact_* = no FMA contraction), so every
// array is bit-identical to the reference's.  Entries a stage never writes stay zero from the one-time
// zero initialisation of the workspace, like the untouched parts of a scratch_array().
//
// All planes of one solver share the row pitch; element (i, j) is at i*pitch + j, ng = 4 ghost cells.
#pragma once
#include <math.h>

#include "../../include/pyro2b200.h"
#include "hydro_core.cuh"

namespace pyro {

struct FlowGeom {
    int nx, ny, ng, pitch;
    int qx, qy;
    double dx, dy;
};

// the eight left / right interface states: u_xl[i, j] is the left state of u at face i-1/2, ...
struct FlowFaces { double *u_xl, *u_xr, *u_yl, *u_yr, *v_xl, *v_xr, *v_yl, *v_yr; };
// Riemann velocities of the transverse step and the four upwinded states (the latter four are reused for
// the final upwinding with the MAC velocities)
struct FlowHat { double *uhat, *vhat, *uxi, *vxi, *uyi, *vyi; };

__device__ __forceinline__ bool flow_cell(const FlowGeom& g, int buf, int& i, int& j)
{
    // thread -> cell of a.v(buf=buf): i in [ng-buf, ng+nx-1+buf], j likewise
    j = blockIdx.x * blockDim.x + threadIdx.x + g.ng - buf;
    i = blockIdx.y * blockDim.y + threadIdx.y + g.ng - buf;
    return i <= g.ng + g.nx - 1 + buf && j <= g.ng + g.ny - 1 + buf;
}

__device__ __forceinline__ bool in_buf2(const FlowGeom& g, int i, int j)
{
    return i >= g.ng - 2 && i <= g.ng + g.nx + 1 && j >= g.ng - 2 && j <= g.ng + g.ny + 1;
}

// shared tail of limit2 / limit4 (reconstruction.py:87-89, 116-118)
__device__ __forceinline__ double mc_limit(double dc, double dl, double dr)
{
    double d1 = exact_mul(2.0, fabs(dl) < fabs(dr) ? dl : dr);
    double dt = fabs(dc) < fabs(d1) ? dc : d1;
    return exact_mul(dl, dr) > 0.0 ? dt : 0.0;
}

// limited slope of plane a at (i, j) along direction (di, dj); limiter 0 / 1 / 2.  The 4th-order
// limiter needs the 2nd-order slopes of the two neighbours, which the reference takes from an array
// that is zero outside the buf = 2 region.
__device__ __forceinline__ double flow_slope(const double* __restrict__ a, const FlowGeom& g, int i, int j, int di,
                                             int dj, int limiter)
{
    const long long k = (long long)i * g.pitch + j, s = (long long)di * g.pitch + dj;
    const double am = a[k - s], a0 = a[k], ap = a[k + s];
    if (limiter == 0) return exact_mul(0.5, exact_sub(ap, am));
    const double dl = exact_sub(ap, a0), dr = exact_sub(a0, am);
    if (limiter == 1) return mc_limit(exact_mul(0.5, exact_sub(ap, am)), dl, dr);
    double l2p = 0.0, l2m = 0.0;
    if (in_buf2(g, i + di, j + dj)) {
        const double app = a[k + 2 * s];
        l2p = mc_limit(exact_mul(0.5, exact_sub(app, a0)), exact_sub(app, ap), dl);
    }
    if (in_buf2(g, i - di, j - dj)) {
        const double amm = a[k - 2 * s];
        l2m = mc_limit(exact_mul(0.5, exact_sub(a0, amm)), dr, exact_sub(am, amm));
    }
    // (2./3.)*(a.ip(1) - a.ip(-1) - 0.25*(lda_tmp.ip(1) + lda_tmp.ip(-1)))
    const double dc = exact_mul(2. / 3., exact_sub(exact_sub(ap, am), exact_mul(0.25, exact_add(l2p, l2m))));
    return mc_limit(dc, dl, dr);
}

// burgers_interface.py:257-284
__device__ __forceinline__ double burgers_riemann(double l, double r)
{
    return (l <= 0.0 && r >= 0.0) ? 0.0 : ((l > 0.0 && exact_add(l, r) > 0.0) ? l : r);
}

// burgers_interface.py:228-254
__device__ __forceinline__ double burgers_upwind(double l, double r, double s)
{
    return (s == 0.0) ? exact_mul(0.5, exact_add(l, r)) : (s > 0.0 ? l : r);
}

// slopes + get_interface_states (burgers_interface.py:45-77), over buf = 2
static __global__ void flow_states_kernel(FlowGeom g, const double* __restrict__ u, const double* __restrict__ v,
                                   FlowFaces S, double dtdx, double dtdy, int limiter)
{
    int i, j;
    if (!flow_cell(g, 2, i, j)) return;
    const long long k = (long long)i * g.pitch + j, kx = k + g.pitch, ky = k + 1;
    const double uu = u[k], vv = v[k];
    const double ldux = flow_slope(u, g, i, j, 1, 0, limiter), ldvx = flow_slope(v, g, i, j, 1, 0, limiter);
    const double lduy = flow_slope(u, g, i, j, 0, 1, limiter), ldvy = flow_slope(v, g, i, j, 0, 1, limiter);
    // q + 0.5*(1.0 - dtdx*u)*ldelta   /   q - 0.5*(1.0 + dtdx*u)*ldelta
    const double cxm = exact_mul(0.5, exact_sub(1.0, exact_mul(dtdx, uu)));
    const double cxp = exact_mul(0.5, exact_add(1.0, exact_mul(dtdx, uu)));
    const double cym = exact_mul(0.5, exact_sub(1.0, exact_mul(dtdy, vv)));
    const double cyp = exact_mul(0.5, exact_add(1.0, exact_mul(dtdy, vv)));
    S.u_xl[kx] = exact_add(uu, exact_mul(cxm, ldux));
    S.u_xr[k] = exact_sub(uu, exact_mul(cxp, ldux));
    S.v_xl[kx] = exact_add(vv, exact_mul(cxm, ldvx));
    S.v_xr[k] = exact_sub(vv, exact_mul(cxp, ldvx));
    S.u_yl[ky] = exact_add(uu, exact_mul(cym, lduy));
    S.u_yr[k] = exact_sub(uu, exact_mul(cyp, lduy));
    S.v_yl[ky] = exact_add(vv, exact_mul(cym, ldvy));
    S.v_yr[k] = exact_sub(vv, exact_mul(cyp, ldvy));
}

// first half of apply_transverse_corrections (burgers_interface.py:108-119): the Riemann velocities
// and the states upwinded with them, over buf = 2
static __global__ void flow_hat_kernel(FlowGeom g, FlowFaces S, FlowHat H)
{
    int i, j;
    if (!flow_cell(g, 2, i, j)) return;
    const long long k = (long long)i * g.pitch + j;
    const double uh = burgers_riemann(S.u_xl[k], S.u_xr[k]);
    const double vh = burgers_riemann(S.v_yl[k], S.v_yr[k]);
    H.uhat[k] = uh;
    H.vhat[k] = vh;
    H.uxi[k] = burgers_upwind(S.u_xl[k], S.u_xr[k], uh);
    H.vxi[k] = burgers_upwind(S.v_xl[k], S.v_xr[k], uh);
    H.uyi[k] = burgers_upwind(S.u_yl[k], S.u_yr[k], vh);
    H.vyi[k] = burgers_upwind(S.v_yl[k], S.v_yr[k], vh);
}

// second half (burgers_interface.py:121-155) followed by apply_gradp_corrections
// (incomp_interface.py:190-209; gpx == nullptr for Burgers).  Thread (i, j) owns the entries
// u_xl[i+1, j], u_xr[i, j], u_yl[i, j+1], u_yr[i, j] (and the same of v), so every entry is updated once.
static __global__ void flow_correct_kernel(FlowGeom g, FlowFaces S, FlowHat H, const double* __restrict__ gpx,
                                    const double* __restrict__ gpy, double dtdx, double dtdy, double dt)
{
    int i, j;
    if (!flow_cell(g, 2, i, j)) return;
    const long long k = (long long)i * g.pitch + j, kx = k + g.pitch, ky = k + 1;
    const double ubar = exact_mul(0.5, exact_add(H.uhat[k], H.uhat[kx]));
    const double vbar = exact_mul(0.5, exact_add(H.vhat[k], H.vhat[ky]));
    // -0.5 * dtdy * vbar * (q_yint.jp(1) - q_yint.v())
    const double hy = exact_mul(exact_mul(-0.5, dtdy), vbar), hx = exact_mul(exact_mul(-0.5, dtdx), ubar);
    const double tu_y = exact_mul(hy, exact_sub(H.uyi[ky], H.uyi[k]));
    const double tv_y = exact_mul(hy, exact_sub(H.vyi[ky], H.vyi[k]));
    const double tv_x = exact_mul(hx, exact_sub(H.vxi[kx], H.vxi[k]));
    const double tu_x = exact_mul(hx, exact_sub(H.uxi[kx], H.uxi[k]));
    double a_uxl = exact_add(S.u_xl[kx], tu_y), a_uxr = exact_add(S.u_xr[k], tu_y);
    double a_vxl = exact_add(S.v_xl[kx], tv_y), a_vxr = exact_add(S.v_xr[k], tv_y);
    double a_vyl = exact_add(S.v_yl[ky], tv_x), a_vyr = exact_add(S.v_yr[k], tv_x);
    double a_uyl = exact_add(S.u_yl[ky], tu_x), a_uyr = exact_add(S.u_yr[k], tu_x);
    if (gpx) {
        const double cx = exact_mul(exact_mul(-0.5, dt), gpx[k]), cy = exact_mul(exact_mul(-0.5, dt), gpy[k]);
        a_uxl = exact_add(a_uxl, cx); a_uxr = exact_add(a_uxr, cx);
        a_vxl = exact_add(a_vxl, cy); a_vxr = exact_add(a_vxr, cy);
        a_vyl = exact_add(a_vyl, cy); a_vyr = exact_add(a_vyr, cy);
        a_uyl = exact_add(a_uyl, cx); a_uyr = exact_add(a_uyr, cx);
    }
    S.u_xl[kx] = a_uxl; S.u_xr[k] = a_uxr;
    S.v_xl[kx] = a_vxl; S.v_xr[k] = a_vxr;
    S.v_yl[ky] = a_vyl; S.v_yr[k] = a_vyr;
    S.u_yl[ky] = a_uyl; S.u_yr[k] = a_uyr;
}

// riemann_and_upwind of the normal velocities (incomp_interface.py:64-69; burgers_interface.py:200-201)
static __global__ void flow_mac_kernel(FlowGeom g, FlowFaces S, double* __restrict__ umac, double* __restrict__ vmac)
{
    int i, j;
    if (!flow_cell(g, 2, i, j)) return;
    const long long k = (long long)i * g.pitch + j;
    const double ul = S.u_xl[k], ur = S.u_xr[k], vl = S.v_yl[k], vr = S.v_yr[k];
    umac[k] = burgers_upwind(ul, ur, burgers_riemann(ul, ur));
    vmac[k] = burgers_upwind(vl, vr, burgers_riemann(vl, vr));
}

// divergence of the MAC velocities into a multigrid-grid plane (ng = 1) (incompressible/simulation.py:257-260)
static __global__ void flow_mac_div_kernel(FlowGeom g, const double* __restrict__ umac, const double* __restrict__ vmac,
                                    double* __restrict__ div, int dpitch)
{
    int i, j;
    if (!flow_cell(g, 0, i, j)) return;
    const long long k = (long long)i * g.pitch + j;
    const double dx_ = exact_div(exact_sub(umac[k + g.pitch], umac[k]), g.dx);
    const double dy_ = exact_div(exact_sub(vmac[k + 1], vmac[k]), g.dy);
    div[(long long)(i - g.ng + 1) * dpitch + (j - g.ng + 1)] = exact_add(dx_, dy_);
}

// subtract the MAC gradient of phi-MAC (incompressible/simulation.py:277-284).  phi = the ng-ghost solver
// plane whose buf = 1 region already holds the multigrid solution.  One thread per cell of v(buf=(0,1,0,1)).
static __global__ void flow_mac_project_kernel(FlowGeom g, const double* __restrict__ phi, double* __restrict__ umac,
                                        double* __restrict__ vmac)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x + g.ng;
    const int i = blockIdx.y * blockDim.y + threadIdx.y + g.ng;
    if (i > g.ng + g.nx || j > g.ng + g.ny) return;
    const long long k = (long long)i * g.pitch + j;
    if (j < g.ng + g.ny) umac[k] = exact_sub(umac[k], exact_div(exact_sub(phi[k], phi[k - g.pitch]), g.dx));
    if (i < g.ng + g.nx) vmac[k] = exact_sub(vmac[k], exact_div(exact_sub(phi[k], phi[k - 1]), g.dy));
}

// upwind all four interface states with the MAC velocities (incomp_interface.py:131-138)
static __global__ void flow_upwind_kernel(FlowGeom g, FlowFaces S, const double* __restrict__ umac,
                                   const double* __restrict__ vmac, FlowHat H)
{
    int i, j;
    if (!flow_cell(g, 2, i, j)) return;
    const long long k = (long long)i * g.pitch + j;
    const double um = umac[k], vm = vmac[k];
    H.uxi[k] = burgers_upwind(S.u_xl[k], S.u_xr[k], um);
    H.vxi[k] = burgers_upwind(S.v_xl[k], S.v_xr[k], um);
    H.uyi[k] = burgers_upwind(S.u_yl[k], S.u_yr[k], vm);
    H.vyi[k] = burgers_upwind(S.v_yl[k], S.v_yr[k], vm);
}

// advective terms and the provisional velocity update (incompressible/simulation.py:316-336), valid cells
// (outside them the reference subtracts dt * 0, and the ghost cells are refilled right after)
static __global__ void flow_advect_kernel(FlowGeom g, const double* __restrict__ umac, const double* __restrict__ vmac,
                                   FlowHat H, double* __restrict__ u, double* __restrict__ v,
                                   const double* __restrict__ gpx, const double* __restrict__ gpy, double dt, int proj_type)
{
    int i, j;
    if (!flow_cell(g, 0, i, j)) return;
    const long long k = (long long)i * g.pitch + j, kx = k + g.pitch, ky = k + 1;
    const double ub = exact_mul(0.5, exact_add(umac[k], umac[kx])), vb = exact_mul(0.5, exact_add(vmac[k], vmac[ky]));
    const double ax = exact_add(exact_div(exact_mul(ub, exact_sub(H.uxi[kx], H.uxi[k])), g.dx),
                                exact_div(exact_mul(vb, exact_sub(H.uyi[ky], H.uyi[k])), g.dy));
    const double ay = exact_add(exact_div(exact_mul(ub, exact_sub(H.vxi[kx], H.vxi[k])), g.dx),
                                exact_div(exact_mul(vb, exact_sub(H.vyi[ky], H.vyi[k])), g.dy));
    if (proj_type == 1) {
        u[k] = exact_sub(u[k], exact_add(exact_mul(dt, ax), exact_mul(dt, gpx[k])));
        v[k] = exact_sub(v[k], exact_add(exact_mul(dt, ay), exact_mul(dt, gpy[k])));
    } else {
        u[k] = exact_sub(u[k], exact_mul(dt, ax));
        v[k] = exact_sub(v[k], exact_mul(dt, ay));
    }
}

// cell-centred divergence 0.5*(u.ip(1) - u.ip(-1))/dx + 0.5*(v.jp(1) - v.jp(-1))/dy into a multigrid-grid
// plane, optionally divided by dt (incompressible/simulation.py:96-97, 367-371)
static __global__ void flow_cc_div_kernel(FlowGeom g, const double* __restrict__ u, const double* __restrict__ v,
                                   double* __restrict__ div, int dpitch, double dt, int divide)
{
    int i, j;
    if (!flow_cell(g, 0, i, j)) return;
    const long long k = (long long)i * g.pitch + j;
    const double a = exact_div(exact_mul(0.5, exact_sub(u[k + g.pitch], u[k - g.pitch])), g.dx);
    const double b = exact_div(exact_mul(0.5, exact_sub(v[k + 1], v[k - 1])), g.dy);
    double d = exact_add(a, b);
    if (divide) d = exact_div(d, dt);
    div[(long long)(i - g.ng + 1) * dpitch + (j - g.ng + 1)] = d;
}

// the final projection's update (incompressible/simulation.py:381-393; dt = 1, proj_type = 0 gives the
// initial projection of preevolve, :113-118): over the whole array, with phi (whose buf = 1 region holds
// the new solution) supplying the centred gradient in the valid cells and zero elsewhere
static __global__ void flow_project_kernel(FlowGeom g, const double* __restrict__ phi, double* __restrict__ u,
                                    double* __restrict__ v, double* __restrict__ gpx, double* __restrict__ gpy,
                                    double dt, int proj_type)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.qx || j >= g.qy) return;
    const long long k = (long long)i * g.pitch + j;
    const bool valid = i >= g.ng && i < g.ng + g.nx && j >= g.ng && j < g.ng + g.ny;
    double gx = 0.0, gy = 0.0;
    if (valid) {
        gx = exact_div(exact_mul(0.5, exact_sub(phi[k + g.pitch], phi[k - g.pitch])), g.dx);
        gy = exact_div(exact_mul(0.5, exact_sub(phi[k + 1], phi[k - 1])), g.dy);
    }
    u[k] = exact_sub(u[k], exact_mul(dt, gx));
    v[k] = exact_sub(v[k], exact_mul(dt, gy));
    if (proj_type == 1) { gpx[k] = exact_add(gpx[k], gx); gpy[k] = exact_add(gpy[k], gy); }
    else if (proj_type == 2) { gpx[k] = gx; gpy[k] = gy; }
}

// Burgers: construct_unsplit_fluxes (burgers_interface.py:200-223) into four of the scratch planes ...
static __global__ void flow_burgers_flux_kernel(FlowGeom g, FlowFaces S, const double* __restrict__ umac,
                                         const double* __restrict__ vmac, FlowHat F)
{
    int i, j;
    if (!flow_cell(g, 2, i, j)) return;
    const long long k = (long long)i * g.pitch + j;
    const double um = umac[k], vm = vmac[k];
    F.uxi[k] = exact_mul(exact_mul(0.5, burgers_upwind(S.u_xl[k], S.u_xr[k], um)), um);
    F.vxi[k] = exact_mul(exact_mul(0.5, burgers_upwind(S.v_xl[k], S.v_xr[k], um)), um);
    F.uyi[k] = exact_mul(exact_mul(0.5, burgers_upwind(S.u_yl[k], S.u_yr[k], vm)), vm);
    F.vyi[k] = exact_mul(exact_mul(0.5, burgers_upwind(S.v_yl[k], S.v_yr[k], vm)), vm);
}

// ... and the conservative update (burgers/simulation.py:117-121), valid cells
static __global__ void flow_burgers_update_kernel(FlowGeom g, FlowHat F, double* __restrict__ u, double* __restrict__ v,
                                           double dtdx, double dtdy)
{
    int i, j;
    if (!flow_cell(g, 0, i, j)) return;
    const long long k = (long long)i * g.pitch + j, kx = k + g.pitch, ky = k + 1;
    u[k] = exact_add(exact_add(u[k], exact_mul(dtdx, exact_sub(F.uxi[k], F.uxi[kx]))),
                     exact_mul(dtdy, exact_sub(F.uyi[k], F.uyi[ky])));
    v[k] = exact_add(exact_add(v[k], exact_mul(dtdx, exact_sub(F.vxi[k], F.vxi[kx]))),
                     exact_mul(dtdy, exact_sub(F.vyi[k], F.vyi[ky])));
}

// ---- linear advection (pyro/advection/interface.py:linear_interface, advective_fluxes.py:unsplit_fluxes,
// advection/simulation.py:56-92): a_t + u a_x + v a_y = 0 with constant u, v ---------------------------
// upwinded, time-centred interface states over buf = 1 (zero elsewhere, like the reference's scratch arrays)
static __global__ void flow_adv_states_kernel(FlowGeom g, const double* __restrict__ a, double* __restrict__ ax,
                                       double* __restrict__ ay, double u, double v, double cx, double cy, int limiter)
{
    int i, j;
    if (!flow_cell(g, 1, i, j)) return;
    const long long k = (long long)i * g.pitch + j;
    if (u < 0) ax[k] = exact_sub(a[k], exact_mul(exact_mul(0.5, exact_add(1.0, cx)), flow_slope(a, g, i, j, 1, 0, limiter)));
    else ax[k] = exact_add(a[k - g.pitch], exact_mul(exact_mul(0.5, exact_sub(1.0, cx)), flow_slope(a, g, i - 1, j, 1, 0, limiter)));
    if (v < 0) ay[k] = exact_sub(a[k], exact_mul(exact_mul(0.5, exact_add(1.0, cy)), flow_slope(a, g, i, j, 0, 1, limiter)));
    else ay[k] = exact_add(a[k - 1], exact_mul(exact_mul(0.5, exact_sub(1.0, cy)), flow_slope(a, g, i, j - 1, 0, 1, limiter)));
}

// fluxes with the transverse correction (advective_fluxes.py:74-92), buf = 1; F_xt = u a_x, F_yt = v a_y
static __global__ void flow_adv_flux_kernel(FlowGeom g, const double* __restrict__ ax, const double* __restrict__ ay,
                                     double* __restrict__ fx, double* __restrict__ fy, double u, double v,
                                     double dtdx2, double dtdy2)
{
    int i, j;
    if (!flow_cell(g, 1, i, j)) return;
    const long long k = (long long)i * g.pitch + j;
    const long long mx = (u <= 0) ? 0 : -(long long)g.pitch, my = (v <= 0) ? 0 : -1;
    fx[k] = exact_mul(u, exact_sub(ax[k], exact_mul(dtdy2, exact_sub(exact_mul(v, ay[k + mx + 1]), exact_mul(v, ay[k + mx])))));
    fy[k] = exact_mul(v, exact_sub(ay[k], exact_mul(dtdx2, exact_sub(exact_mul(u, ax[k + g.pitch + my]), exact_mul(u, ax[k + my])))));
}

static __global__ void flow_adv_update_kernel(FlowGeom g, double* __restrict__ a, const double* __restrict__ fx,
                                       const double* __restrict__ fy, double dtdx, double dtdy)
{
    int i, j;
    if (!flow_cell(g, 0, i, j)) return;
    const long long k = (long long)i * g.pitch + j;
    a[k] = exact_add(exact_add(a[k], exact_mul(dtdx, exact_sub(fx[k], fx[k + g.pitch]))),
                     exact_mul(dtdy, exact_sub(fy[k], fy[k + 1])));
}

// max |u|, max |v| over the full arrays including ghost cells (burgers/simulation.py:51-58): the bit
// patterns of the maxima are combined with atomicMax (non-negative doubles order like their bits)
static __global__ void flow_maxabs_kernel(FlowGeom g, const double* __restrict__ u, const double* __restrict__ v,
                                   unsigned long long* out)
{
    const long long total = (long long)g.qx * g.qy;
    double mu = 0.0, mv = 0.0;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long k = (t / g.qy) * g.pitch + (t % g.qy);
        mu = dmax(mu, fabs(u[k]));
        mv = dmax(mv, fabs(v[k]));
    }
    for (int o = 16; o > 0; o >>= 1) {
        mu = dmax(mu, __shfl_xor_sync(0xffffffffu, mu, o));
        mv = dmax(mv, __shfl_xor_sync(0xffffffffu, mv, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&out[0], (unsigned long long)__double_as_longlong(mu));
        atomicMax(&out[1], (unsigned long long)__double_as_longlong(mv));
    }
}

}  // namespace pyro

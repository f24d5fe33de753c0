// lm_kernels.cuh -- device code of the low Mach number atmosphere solver's explicit stages (lm_atm): the
// third caller of the multigrid path (variable-coefficient projections).  No runtime-API calls; included by
// lm.cu (nvcc) and tests/emu/lm_emu.cpp (g++ through tests/emu/cuda_emu.h).
//
// Reference behaviour (pyro2, file:line):
//   LM_atm_interface.get_interface_states            pyro/lm_atm/LM_atm_interface.py:429-585
//   LM_atm_interface.mac_vels / states / rho_states                                  :181-426
//   LM_atm_interface.upwind / riemann                                                :588-677
//   Simulation.evolve / preevolve / method_compute_timestep   pyro/lm_atm/simulation.py:138-618
//
// The numba routines loop over fixed index ranges that differ from the incompressible solver's (Riemann /
// upwind over [ng-1, ng+nx+1], the transverse terms over [ng-1, ng+nx]) and leave everything else zero;
// cell updates next to the boundary read such partially built entries, so each kernel covers exactly the
// reference's range, with individually rounded operations in the reference's order (bit-identical arrays).
#pragma once
#include "flow_kernels.cuh"

namespace pyro {

struct LmBase { const double *rho0, *p0, *beta0, *beta0e; };      // 1-d base state, indexed by j

// thread -> cell of a region that extends lo cells below and hi cells above the valid cells
__device__ __forceinline__ bool lm_cell(const FlowGeom& g, int lo, int hi, int& i, int& j)
{
    j = blockIdx.x * blockDim.x + threadIdx.x + g.ng - lo;
    i = blockIdx.y * blockDim.y + threadIdx.y + g.ng - lo;
    return i <= g.ng + g.nx - 1 + hi && j <= g.ng + g.ny - 1 + hi;
}

__device__ __forceinline__ bool lm_in_ru(const FlowGeom& g, int i, int j)   // the Riemann / upwind range
{
    return i >= g.ng - 1 && i <= g.ng + g.nx + 1 && j >= g.ng - 1 && j <= g.ng + g.ny + 1;
}

__device__ __forceinline__ double lm_upwind(double l, double r, double s)
{
    return s > 0.0 ? l : (s == 0.0 ? exact_mul(0.5, exact_add(l, r)) : r);
}

__device__ __forceinline__ double lm_riemann(double l, double r)
{
    return (l > 0.0 && exact_add(l, r) > 0.0) ? l : ((l <= 0.0 && r >= 0.0) ? 0.0 : r);
}

// coeff = numer / (d1 [+ d2]) over a.v(buf), then * b[j] or * b[j]**2 (two roundings, as the reference writes it:
// simulation.py:312-313, 347-348, 387-388, 437-438, 514-515)
static __global__ void lm_coeff_kernel(FlowGeom g, const double* __restrict__ d1, const double* __restrict__ d2, double numer,
                                const double* __restrict__ b, int squared, int buf, double* __restrict__ coeff)
{
    int i, j;
    if (!lm_cell(g, buf, buf, i, j)) return;
    const long long k = (long long)i * g.pitch + j;
    const double den = d2 ? exact_add(d1[k], d2[k]) : d1[k];
    const double c = exact_div(numer, den);
    coeff[k] = exact_mul(c, squared ? exact_mul(b[j], b[j]) : b[j]);
}

// buoyancy source rho' g / rho with rho' = rho - rho0(j): valid cells from rho (simulation.py:319-323) or the
// whole array from rho_half = 0.5 (rho + rho_old) (:493-496)
static __global__ void lm_source_kernel(FlowGeom g, const double* __restrict__ rho, const double* __restrict__ rho_old,
                                 const double* __restrict__ rho0, double grav, double* __restrict__ source)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.qx || j >= g.qy) return;
    const long long k = (long long)i * g.pitch + j;
    if (rho_old) {
        const double rh = exact_mul(0.5, exact_add(rho[k], rho_old[k]));
        source[k] = exact_div(exact_mul(exact_sub(rh, rho0[j]), grav), rh);
    } else if (i >= g.ng && i < g.ng + g.nx && j >= g.ng && j < g.ng + g.ny) {
        source[k] = exact_div(exact_mul(exact_sub(rho[k], rho0[j]), grav), rho[k]);
    }
}

// Riemann velocities and the states upwinded with them (LM_atm_interface.py:520-527), range [ng-1, ng+n+1]
static __global__ void lm_hat_kernel(FlowGeom g, FlowFaces S, FlowHat H)
{
    int i, j;
    if (!lm_cell(g, 1, 2, i, j)) return;
    const long long k = (long long)i * g.pitch + j;
    const double uh = lm_riemann(S.u_xl[k], S.u_xr[k]);
    const double vh = lm_riemann(S.v_yl[k], S.v_yr[k]);
    H.uhat[k] = uh;
    H.vhat[k] = vh;
    H.uxi[k] = lm_upwind(S.u_xl[k], S.u_xr[k], uh);
    H.vxi[k] = lm_upwind(S.v_xl[k], S.v_xr[k], uh);
    H.uyi[k] = lm_upwind(S.u_yl[k], S.u_yr[k], vh);
    H.vyi[k] = lm_upwind(S.v_yl[k], S.v_yr[k], vh);
}

// transverse, pressure-gradient (coeff * gradp) and buoyancy terms (LM_atm_interface.py:529-583), range [ng-1, ng+n]
static __global__ void lm_correct_kernel(FlowGeom g, FlowFaces S, FlowHat H, const double* __restrict__ coeff,
                                  const double* __restrict__ gpx, const double* __restrict__ gpy,
                                  const double* __restrict__ source, double dtdx, double dtdy, double dt)
{
    int i, j;
    if (!lm_cell(g, 1, 1, i, j)) return;
    const long long k = (long long)i * g.pitch + j, kx = k + g.pitch, ky = k + 1;
    const double ubar = exact_mul(0.5, exact_add(H.uhat[k], H.uhat[kx]));
    const double vbar = exact_mul(0.5, exact_add(H.vhat[k], H.vhat[ky]));
    const double hx = exact_mul(0.5, dtdx), hy = exact_mul(0.5, dtdy), hdt = exact_mul(0.5, dt);
    const double px = exact_mul(hdt, exact_mul(coeff[k], gpx[k])), py = exact_mul(hdt, exact_mul(coeff[k], gpy[k]));
    const double sv = exact_mul(hdt, source[k]);
    // q = q - 0.5*dtdy*(vbar*dq) - 0.5*dt*gradp [+ 0.5*dt*source]
    const double vu_y = exact_mul(hy, exact_mul(vbar, exact_sub(H.uyi[ky], H.uyi[k])));
    S.u_xl[kx] = exact_sub(exact_sub(S.u_xl[kx], vu_y), px);
    S.u_xr[k] = exact_sub(exact_sub(S.u_xr[k], vu_y), px);
    const double vv_y = exact_mul(hy, exact_mul(vbar, exact_sub(H.vyi[ky], H.vyi[k])));
    S.v_xl[kx] = exact_add(exact_sub(exact_sub(S.v_xl[kx], vv_y), py), sv);
    S.v_xr[k] = exact_add(exact_sub(exact_sub(S.v_xr[k], vv_y), py), sv);
    const double uv_x = exact_mul(hx, exact_mul(ubar, exact_sub(H.vxi[kx], H.vxi[k])));
    S.v_yl[ky] = exact_add(exact_sub(exact_sub(S.v_yl[ky], uv_x), py), sv);
    S.v_yr[k] = exact_add(exact_sub(exact_sub(S.v_yr[k], uv_x), py), sv);
    const double uu_x = exact_mul(hx, exact_mul(ubar, exact_sub(H.uxi[kx], H.uxi[k])));
    S.u_yl[ky] = exact_sub(exact_sub(S.u_yl[ky], uu_x), px);
    S.u_yr[k] = exact_sub(exact_sub(S.u_yr[k], uu_x), px);
}

// riemann_and_upwind of the normal velocities, range [ng-1, ng+n+1]
static __global__ void lm_mac_kernel(FlowGeom g, FlowFaces S, double* __restrict__ umac, double* __restrict__ vmac)
{
    int i, j;
    if (!lm_cell(g, 1, 2, i, j)) return;
    const long long k = (long long)i * g.pitch + j;
    const double ul = S.u_xl[k], ur = S.u_xr[k], vl = S.v_yl[k], vr = S.v_yr[k];
    umac[k] = lm_upwind(ul, ur, lm_riemann(ul, ur));
    vmac[k] = lm_upwind(vl, vr, lm_riemann(vl, vr));
}

// D(beta0 U_MAC) into a multigrid-grid plane (simulation.py:362-366)
static __global__ void lm_mac_div_kernel(FlowGeom g, LmBase B, const double* __restrict__ umac, const double* __restrict__ vmac,
                                  double* __restrict__ div, int dpitch)
{
    int i, j;
    if (!lm_cell(g, 0, 0, i, j)) return;
    const long long k = (long long)i * g.pitch + j;
    const double a = exact_div(exact_mul(B.beta0[j], exact_sub(umac[k + g.pitch], umac[k])), g.dx);
    const double b = exact_div(exact_sub(exact_mul(B.beta0e[j + 1], vmac[k + 1]), exact_mul(B.beta0e[j], vmac[k])), g.dy);
    div[(long long)(i - g.ng + 1) * dpitch + (j - g.ng + 1)] = exact_add(a, b);
}

// U_MAC -= (edge-centred beta0/rho) G phi_MAC (simulation.py:391-408); coeff = the ghost-filled beta0/rho
static __global__ void lm_mac_project_kernel(FlowGeom g, const double* __restrict__ coeff, const double* __restrict__ phi,
                                      double* __restrict__ umac, double* __restrict__ vmac)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x + g.ng;
    const int i = blockIdx.y * blockDim.y + threadIdx.y + g.ng;
    if (i > g.ng + g.nx || j > g.ng + g.ny) return;
    const long long k = (long long)i * g.pitch + j;
    if (j < g.ng + g.ny) {
        const double cx = exact_mul(0.5, exact_add(coeff[k - g.pitch], coeff[k]));
        umac[k] = exact_sub(umac[k], exact_div(exact_mul(cx, exact_sub(phi[k], phi[k - g.pitch])), g.dx));
    }
    if (i < g.ng + g.nx) {
        const double cy = exact_mul(0.5, exact_add(coeff[k - 1], coeff[k]));
        vmac[k] = exact_sub(vmac[k], exact_div(exact_mul(cy, exact_sub(phi[k], phi[k - 1])), g.dy));
    }
}

// rho_states, first loop (LM_atm_interface.py:372-385): predictor with the MAC velocities, buf = 2
static __global__ void lm_rho_pred_kernel(FlowGeom g, const double* __restrict__ rho, const double* __restrict__ umac,
                                   const double* __restrict__ vmac, double* __restrict__ xl, double* __restrict__ xr,
                                   double* __restrict__ yl, double* __restrict__ yr, double dtdx, double dtdy, int limiter)
{
    int i, j;
    if (!lm_cell(g, 2, 2, i, j)) return;
    const long long k = (long long)i * g.pitch + j, kx = k + g.pitch, ky = k + 1;
    const double r = rho[k];
    const double ldx = flow_slope(rho, g, i, j, 1, 0, limiter), ldy = flow_slope(rho, g, i, j, 0, 1, limiter);
    xl[kx] = exact_add(r, exact_mul(exact_mul(0.5, exact_sub(1.0, exact_mul(dtdx, umac[kx]))), ldx));
    xr[k] = exact_sub(r, exact_mul(exact_mul(0.5, exact_add(1.0, exact_mul(dtdx, umac[k]))), ldx));
    yl[ky] = exact_add(r, exact_mul(exact_mul(0.5, exact_sub(1.0, exact_mul(dtdy, vmac[ky]))), ldy));
    yr[k] = exact_sub(r, exact_mul(exact_mul(0.5, exact_add(1.0, exact_mul(dtdy, vmac[k]))), ldy));
}

// two upwindings at once over the Riemann / upwind range: a = upwind(al, ar, sa), b = upwind(bl, br, sb)
static __global__ void lm_upwind2_kernel(FlowGeom g, const double* __restrict__ al, const double* __restrict__ ar,
                                  const double* __restrict__ sa, double* __restrict__ a, const double* __restrict__ bl,
                                  const double* __restrict__ br, const double* __restrict__ sb, double* __restrict__ b)
{
    int i, j;
    if (!lm_cell(g, 1, 2, i, j)) return;
    const long long k = (long long)i * g.pitch + j;
    a[k] = lm_upwind(al[k], ar[k], sa[k]);
    b[k] = lm_upwind(bl[k], br[k], sb[k]);
}

// rho_states, second loop (LM_atm_interface.py:392-418): transverse and non-advective terms, buf = 2
static __global__ void lm_rho_trans_kernel(FlowGeom g, const double* __restrict__ rho, const double* __restrict__ umac,
                                    const double* __restrict__ vmac, const double* __restrict__ rxi,
                                    const double* __restrict__ ryi, double* __restrict__ xl, double* __restrict__ xr,
                                    double* __restrict__ yl, double* __restrict__ yr, double dt)
{
    int i, j;
    if (!lm_cell(g, 2, 2, i, j)) return;
    const long long k = (long long)i * g.pitch + j, kx = k + g.pitch, ky = k + 1;
    const double u_x = exact_div(exact_sub(umac[kx], umac[k]), g.dx);
    const double v_y = exact_div(exact_sub(vmac[ky], vmac[k]), g.dy);
    const double hdt = exact_mul(0.5, dt);
    const double rhov_y = exact_div(exact_sub(exact_mul(ryi[ky], vmac[ky]), exact_mul(ryi[k], vmac[k])), g.dy);
    const double tx = exact_mul(hdt, exact_add(rhov_y, exact_mul(rho[k], u_x)));
    xl[kx] = exact_sub(xl[kx], tx);
    xr[k] = exact_sub(xr[k], tx);
    const double rhou_x = exact_div(exact_sub(exact_mul(rxi[kx], umac[kx]), exact_mul(rxi[k], umac[k])), g.dx);
    const double ty = exact_mul(hdt, exact_add(rhou_x, exact_mul(rho[k], v_y)));
    yl[ky] = exact_sub(yl[ky], ty);
    yr[k] = exact_sub(yr[k], ty);
}

// conservative density update and the new eint = p0 / (gamma - 1) / rho (simulation.py:421-432), valid cells
static __global__ void lm_rho_update_kernel(FlowGeom g, LmBase B, double* __restrict__ rho, double* __restrict__ eint,
                                     const double* __restrict__ umac, const double* __restrict__ vmac,
                                     const double* __restrict__ rxi, const double* __restrict__ ryi, double dt, double gamma)
{
    int i, j;
    if (!lm_cell(g, 0, 0, i, j)) return;
    const long long k = (long long)i * g.pitch + j, kx = k + g.pitch, ky = k + 1;
    const double fx = exact_div(exact_sub(exact_mul(rxi[kx], umac[kx]), exact_mul(rxi[k], umac[k])), g.dx);
    const double fy = exact_div(exact_sub(exact_mul(ryi[ky], vmac[ky]), exact_mul(ryi[k], vmac[k])), g.dy);
    const double rn = exact_sub(rho[k], exact_mul(dt, exact_add(fx, fy)));
    rho[k] = rn;
    eint[k] = exact_div(exact_div(B.p0[j], exact_sub(gamma, 1.0)), rn);
}

// v[:, :] += dt * source over the whole array (simulation.py:499)
static __global__ void lm_add_source_kernel(FlowGeom g, double* __restrict__ v, const double* __restrict__ source, double dt)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.qx || j >= g.qy) return;
    const long long k = (long long)i * g.pitch + j;
    v[k] = exact_add(v[k], exact_mul(dt, source[k]));
}

// D(beta0 U) with centred differences into a multigrid-grid plane, optionally / dt (simulation.py:226-229, 538-543)
static __global__ void lm_cc_div_kernel(FlowGeom g, LmBase B, const double* __restrict__ u, const double* __restrict__ v,
                                 double* __restrict__ div, int dpitch, double dt, int divide)
{
    int i, j;
    if (!lm_cell(g, 0, 0, i, j)) return;
    const long long k = (long long)i * g.pitch + j;
    const double a = exact_div(exact_mul(exact_mul(0.5, B.beta0[j]), exact_sub(u[k + g.pitch], u[k - g.pitch])), g.dx);
    const double b = exact_div(exact_mul(0.5, exact_sub(exact_mul(B.beta0[j + 1], v[k + 1]), exact_mul(B.beta0[j - 1], v[k - 1]))), g.dy);
    double d = exact_add(a, b);
    if (divide) d = exact_div(d, dt);
    div[(long long)(i - g.ng + 1) * dpitch + (j - g.ng + 1)] = d;
}

// U -= dt (beta0/rho) G phi in the valid cells, gradp bookkeeping (simulation.py:556-590; dt = 1, proj_type = 0:
// the initial projection of preevolve, :241-251).  phi's buf = 1 region holds the multigrid solution.
static __global__ void lm_project_kernel(FlowGeom g, LmBase B, const double* __restrict__ rho, const double* __restrict__ phi,
                                  double* __restrict__ u, double* __restrict__ v, double* __restrict__ gpx,
                                  double* __restrict__ gpy, double dt, int proj_type)
{
    int i, j;
    if (!lm_cell(g, 0, 0, i, j)) return;
    const long long k = (long long)i * g.pitch + j;
    const double gx = exact_div(exact_mul(0.5, exact_sub(phi[k + g.pitch], phi[k - g.pitch])), g.dx);
    const double gy = exact_div(exact_mul(0.5, exact_sub(phi[k + 1], phi[k - 1])), g.dy);
    const double c = exact_mul(exact_div(1.0, rho[k]), B.beta0[j]);
    u[k] = exact_sub(u[k], exact_mul(exact_mul(dt, c), gx));
    v[k] = exact_sub(v[k], exact_mul(exact_mul(dt, c), gy));
    if (proj_type == 1) { gpx[k] = exact_add(gpx[k], gx); gpy[k] = exact_add(gpy[k], gy); }
    else if (proj_type == 2) { gpx[k] = gx; gpy[k] = gy; }
}

// the five maxima of method_compute_timestep (simulation.py:138-178): |u|, |v| over the whole arrays, |u|, |v|
// and |rho' g| / rho over the valid cells; bit patterns combined with atomicMax
static __global__ void lm_reduce_kernel(FlowGeom g, const double* __restrict__ rho, const double* __restrict__ u,
                                 const double* __restrict__ v, const double* __restrict__ rho0, double grav,
                                 unsigned long long* out)
{
    const long long total = (long long)g.qx * g.qy;
    double m[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t / g.qy), j = (int)(t % g.qy);
        const long long k = (long long)i * g.pitch + j;
        const double au = fabs(u[k]), av = fabs(v[k]);
        m[0] = dmax(m[0], au);
        m[1] = dmax(m[1], av);
        if (i >= g.ng && i < g.ng + g.nx && j >= g.ng && j < g.ng + g.ny) {
            m[2] = dmax(m[2], au);
            m[3] = dmax(m[3], av);
            m[4] = dmax(m[4], exact_div(fabs(exact_mul(exact_sub(rho[k], rho0[j]), grav)), rho[k]));
        }
    }
    for (int q = 0; q < 5; ++q) {
        double x = m[q];
        for (int o = 16; o > 0; o >>= 1) x = dmax(x, __shfl_xor_sync(0xffffffffu, x, o));
        if ((threadIdx.x & 31) == 0) atomicMax(&out[q], (unsigned long long)__double_as_longlong(x));
    }
}

}  // namespace pyro

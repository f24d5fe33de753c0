// bc_user_kernels.cuh -- the compressible solver's user-defined boundary conditions as device code.
// Reference: pyro/compressible/BC.py:21-168, the "hse" and "ambient" boundaries ("ramp" is not built).
// Included by bc_user.cu (nvcc) and tests/emu/bc_emu.cpp (g++ through tests/emu/cuda_emu.h).
#pragma once
#include "../../include/pyro2b200.h"
#include "hydro_core.cuh"

namespace pyro {

// One thread per x index i (the reference's v[:, j] slices span the whole x extent, ghost columns
// included).  side 0 = ylb (fill j = jlo-1 ... 0), 1 = yrb (j = jhi+1 ... qy-1).
// var != energy: zero-gradient copy of the first interior row (BC.py:55-63, 111-117).
// energy: hydrostatic equilibrium integrated outward at the base density, p -/+= g rho dy per cell,
//         E = p/(gamma-1) + KE_base (BC.py:65-98, 119-148); individually rounded operations in the
//         reference's order, so the ghost cells are bit-identical to the reference's.
__global__ void hse_fill_kernel(double* __restrict__ U, p2b_grid g, double grav, double gamma, int var, int side)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int qx = g.nx + 2 * g.ng;
    if (i >= qx) return;
    const int jb = side == 0 ? g.ng : g.ng + g.ny - 1;
    const int step = side == 0 ? -1 : 1;
    const long long kb = (long long)i * g.pitch + jb;
    double* v = U + (long long)var * g.plane_stride;
    if (var != 1) {
        const double base = v[kb];
        for (int k = 1; k <= g.ng; ++k) v[kb + step * k] = base;
        return;
    }
    const double dens = U[kb], xmom = U[2 * g.plane_stride + kb], ymom = U[3 * g.plane_stride + kb];
    // 0.5*(xmom**2 + ymom**2)/dens ; (ener - ke)/dens ; dens*eint*(gamma - 1)
    const double ke = exact_div(exact_mul(0.5, exact_add(exact_mul(xmom, xmom), exact_mul(ymom, ymom))), dens);
    const double eint = exact_div(exact_sub(v[kb], ke), dens);
    double pres = exact_mul(exact_mul(dens, eint), exact_sub(gamma, 1.0));
    const double dp = exact_mul(exact_mul(grav, dens), g.dy);
    for (int k = 1; k <= g.ng; ++k) {
        pres = side == 0 ? exact_sub(pres, dp) : exact_add(pres, dp);
        v[kb + step * k] = exact_add(exact_div(pres, exact_sub(gamma, 1.0)), ke);
    }
}

// "ambient" (BC.py:142-168): the ghost rows of one variable beyond the +y (side 1) or -y (side 0) boundary are set
// to a constant -- the ambient density, momenta or total energy the problem registered; all x indices.
__global__ void ambient_fill_kernel(double* __restrict__ U, p2b_grid g, int var, int side, double value)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.nx + 2 * g.ng) return;
    const int jb = side == 0 ? g.ng : g.ng + g.ny - 1;
    const int step = side == 0 ? -1 : 1;
    double* v = U + (long long)var * g.plane_stride + (long long)i * g.pitch + jb;
    for (int k = 1; k <= g.ng; ++k) v[step * k] = value;
}

}  // namespace pyro

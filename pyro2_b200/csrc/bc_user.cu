// bc_user.cu -- host entry points of the solver-specific boundary conditions (kernels: bc_user_kernels.cuh)
#include "common.cuh"
#include "bc_user_kernels.cuh"

using namespace pyro;

extern "C" {

// the "hse" boundary of pyro/compressible/BC.py for variable `var` (0 density, 1 energy, 2 x-momentum,
// 3 y-momentum) of a 4-plane state on side 0 (ylb) / 1 (yrb): what bnd.ext_bcs["hse"] does when
// CellCenterData2d.fill_BC(name) calls it after the standard x fill
int p2b_fill_hse_f64(double* U, const p2b_grid* g, double grav, double gamma, int var, int side, void* stream)
{
    P2B_REQUIRE(U && g, "null pointer");
    P2B_REQUIRE(var >= 0 && var <= 3 && (side == 0 || side == 1), "bad variable / side");
    P2B_REQUIRE(g->ng >= 1 && g->ny >= 1, "bad grid");
    const int qx = g->nx + 2 * g->ng;
    P2B_LAUNCH(hse_fill_kernel, (qx + 127) / 128, 128, 0, (cudaStream_t)stream)(U, *g, grav, gamma, var, side);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

// the "ambient" boundary of pyro/compressible/BC.py:142-168: ghost rows of variable `var` on side 0 (ylb) / 1 (yrb)
// <- value (ambient_rho, ambient_rho * u, ambient_rho * v or p / (gamma - 1) + kinetic energy; the reference only
// supports yrb)
int p2b_fill_ambient_f64(double* U, const p2b_grid* g, int var, int side, double value, void* stream)
{
    P2B_REQUIRE(U && g, "null pointer");
    P2B_REQUIRE(var >= 0 && var <= 3 && (side == 0 || side == 1), "bad variable / side");
    const int qx = g->nx + 2 * g->ng;
    P2B_LAUNCH(ambient_fill_kernel, (qx + 127) / 128, 128, 0, (cudaStream_t)stream)(U, *g, var, side, value);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

}  // extern "C"

// ghost_cfl.cu -- ghost-cell fill and the stand-alone CFL reduction.
//
//   fill kernels : ArrayIndexer.fill_ghost, pyro/mesh/array_indexer.py:150-274 (bit-exact copies /
//                  negations; x faces first, then y faces over the full x extent so that corners
//                  inherit exactly as in the reference)
//   cfl kernel   : Simulation.method_compute_timestep, pyro/compressible/simulation.py:267-288 with
//                  derive_primitives (derives.py:6-69); minimum over the full array incl. ghosts
#include <stdarg.h>

#include "common.cuh"
#include "hydro_core.cuh"

namespace pyro {

char* last_error_buf()
{
    static thread_local char buf[512] = "";
    return buf;
}

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buf(), 512, fmt, ap);
    va_end(ap);
}

struct BcTable { int code[16][4]; };   // up to 16 variables per launch

// source row for ghost row i on the low side (i < ng) / high side (i > ihi); sign by reference
template <typename T>
__device__ __forceinline__ T apply_sign(T v, int code) { return code == P2B_BC_REFLECT_ODD ? (T)(-v) : v; }

__device__ __forceinline__ int src_lo(int i, int ng, int n, int code)
{
    // array_indexer.py:164-188 (-x) / :221-246 (-y)
    if (code == P2B_BC_OUTFLOW) return ng;
    if (code == P2B_BC_PERIODIC) return (ng + n - 1) - ng + i + 1;
    return 2 * ng - i - 1;   // reflect-even / reflect-odd
}

__device__ __forceinline__ int src_hi(int i, int ng, int n, int code)
{
    // array_indexer.py:191-218 (+x) / :249-274 (+y); ihi = ng + n - 1
    const int ihi = ng + n - 1;
    if (code == P2B_BC_OUTFLOW) return ihi;
    if (code == P2B_BC_PERIODIC) return i - ihi - 1 + ng;
    return ihi - (i - ihi - 1);   // i_bnd = ihi+1+k  <-  i_src = ihi-k
}

// x faces: one thread per (variable, ghost layer, j); j contiguous -> coalesced
template <typename T>
__global__ void fill_x_kernel(T* base, p2b_grid g, int nvar, BcTable bc)
{
    const int qy = g.ny + 2 * g.ng;
    const int total = nvar * 2 * g.ng * qy;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        int j = t % qy;
        int k = (t / qy) % (2 * g.ng);
        int n = t / (qy * 2 * g.ng);
        T* a = base + (long long)n * g.plane_stride;
        if (k < g.ng) {
            int code = bc.code[n][0];
            if (code == P2B_BC_NONE) continue;
            int i = k;
            a[(long long)i * g.pitch + j] = apply_sign(a[(long long)src_lo(i, g.ng, g.nx, code) * g.pitch + j], code);
        } else {
            int code = bc.code[n][1];
            if (code == P2B_BC_NONE) continue;
            int i = g.ng + g.nx + (k - g.ng);
            a[(long long)i * g.pitch + j] = apply_sign(a[(long long)src_hi(i, g.ng, g.nx, code) * g.pitch + j], code);
        }
    }
}

// y faces: one thread per (variable, i, ghost layer)
template <typename T>
__global__ void fill_y_kernel(T* base, p2b_grid g, int nvar, BcTable bc)
{
    const int qx = g.nx + 2 * g.ng;
    const int total = nvar * qx * 2 * g.ng;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        int k = t % (2 * g.ng);
        int i = (t / (2 * g.ng)) % qx;
        int n = t / (2 * g.ng * qx);
        T* row = base + (long long)n * g.plane_stride + (long long)i * g.pitch;
        if (k < g.ng) {
            int code = bc.code[n][2];
            if (code == P2B_BC_NONE) continue;
            row[k] = apply_sign(row[src_lo(k, g.ng, g.ny, code)], code);
        } else {
            int code = bc.code[n][3];
            if (code == P2B_BC_NONE) continue;
            int j = g.ng + g.ny + (k - g.ng);
            row[j] = apply_sign(row[src_hi(j, g.ng, g.ny, code)], code);
        }
    }
}

// inhomogeneous variants touch only the first ghost cell (array_indexer.py:166-183)
__global__ void fill_x_values_kernel(double* a, p2b_grid g, int xlb, int xrb, const double* xl, const double* xr)
{
    const int qy = g.ny + 2 * g.ng;
    const int ilo = g.ng, ihi = g.ng + g.nx - 1;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < qy; j += gridDim.x * blockDim.x) {
        if (xl) {
            if (xlb == P2B_BC_OUTFLOW) a[(long long)(ilo - 1) * g.pitch + j] = a[(long long)ilo * g.pitch + j] - g.dx * xl[j];
            else if (xlb == P2B_BC_REFLECT_ODD) a[(long long)(ilo - 1) * g.pitch + j] = 2 * xl[j] - a[(long long)ilo * g.pitch + j];
        }
        if (xr) {
            if (xrb == P2B_BC_OUTFLOW) a[(long long)(ihi + 1) * g.pitch + j] = a[(long long)ihi * g.pitch + j] + g.dx * xr[j];
            else if (xrb == P2B_BC_REFLECT_ODD) a[(long long)(ihi + 1) * g.pitch + j] = 2 * xr[j] - a[(long long)ihi * g.pitch + j];
        }
    }
}

__global__ void fill_y_values_kernel(double* a, p2b_grid g, int ylb, int yrb, const double* yl, const double* yr)
{
    const int qx = g.nx + 2 * g.ng;
    const int jlo = g.ng, jhi = g.ng + g.ny - 1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < qx; i += gridDim.x * blockDim.x) {
        double* row = a + (long long)i * g.pitch;
        if (yl) {
            if (ylb == P2B_BC_OUTFLOW) row[jlo - 1] = row[jlo] - g.dy * yl[i];
            else if (ylb == P2B_BC_REFLECT_ODD) row[jlo - 1] = 2 * yl[i] - row[jlo];
        }
        if (yr) {
            if (yrb == P2B_BC_OUTFLOW) row[jhi + 1] = row[jhi] + g.dy * yr[i];
            else if (yrb == P2B_BC_REFLECT_ODD) row[jhi + 1] = 2 * yr[i] - row[jhi];
        }
    }
}

template <typename T>
int fill_ghost_impl(T* base, const p2b_grid* g, int nvar, const int* bc, cudaStream_t st)
{
    P2B_REQUIRE(base && g && bc, "null pointer");
    P2B_REQUIRE(g->nx > 0 && g->ny > 0 && g->ng >= 0, "bad grid");
    P2B_REQUIRE(g->pitch >= g->ny + 2 * g->ng, "pitch < qy");
    P2B_REQUIRE(nvar >= 1, "nvar < 1");
    if (g->ng == 0) return P2B_OK;
    for (int n0 = 0; n0 < nvar; n0 += 16) {
        int nv = nvar - n0 < 16 ? nvar - n0 : 16;
        BcTable t;
        memset(&t, 0, sizeof t);
        for (int n = 0; n < nv; ++n)
            for (int s = 0; s < 4; ++s) {
                int c = bc[(n0 + n) * 4 + s];
                P2B_REQUIRE(c >= 0 && c <= P2B_BC_NONE, "bad BC code");
                // periodic needs ng <= n; reflect needs ng <= n as well
                t.code[n][s] = c;
            }
        const int qx = g->nx + 2 * g->ng, qy = g->ny + 2 * g->ng;
        T* b = base + (long long)n0 * g->plane_stride;
        int tx = nv * 2 * g->ng * qy, ty = nv * qx * 2 * g->ng;
        P2B_LAUNCH(fill_x_kernel<T>, (tx + 255) / 256 < 1184 ? (tx + 255) / 256 : 1184, 256, 0, st)(b, *g, nv, t);
        P2B_LAUNCH(fill_y_kernel<T>, (ty + 255) / 256 < 1184 ? (ty + 255) / 256 : 1184, 256, 0, st)(b, *g, nv, t);
    }
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

// ---- CFL ------------------------------------------------------------------------------------
__global__ void cfl_kernel(const double* U, p2b_grid g, double gamma, unsigned long long* out)
{
    const int qx = g.nx + 2 * g.ng, qy = g.ny + 2 * g.ng;
    const long long total = (long long)qx * qy;
    double mx = 0.0, my = 0.0;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        int i = (int)(t / qy), j = (int)(t % qy);
        const double* p = U + (long long)i * g.pitch + j;
        double ax, ay;
        cfl_speeds(p[0], p[g.plane_stride], p[2 * g.plane_stride], p[3 * g.plane_stride], gamma, ax, ay);
        // NaN-propagating like numpy's min would not be; keep plain max (NaN never wins)
        mx = dmax(mx, ax); my = dmax(my, ay);
    }
    for (int o = 16; o > 0; o >>= 1) {
        mx = dmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        my = dmax(my, __shfl_xor_sync(0xffffffffu, my, o));
    }
    __shared__ double sx[32], sy[32];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { sx[wid] = mx; sy[wid] = my; }
    __syncthreads();
    if (wid == 0) {
        int nw = blockDim.x >> 5;
        mx = lane < nw ? sx[lane] : 0.0; my = lane < nw ? sy[lane] : 0.0;
        for (int o = 16; o > 0; o >>= 1) {
            mx = dmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            my = dmax(my, __shfl_xor_sync(0xffffffffu, my, o));
        }
        if (lane == 0) {
            atomicMax(&out[0], (unsigned long long)__double_as_longlong(mx));
            atomicMax(&out[1], (unsigned long long)__double_as_longlong(my));
        }
    }
}

}  // namespace pyro

using namespace pyro;

extern "C" {

const char* p2b_last_error(void) { return last_error_buf(); }
int p2b_version(void) { return 100; }
int p2b_device_sms(void) { return num_sms(); }

int p2b_fill_ghost_f64(double* base, const p2b_grid* g, int nvar, const int* bc, void* stream)
{
    return fill_ghost_impl<double>(base, g, nvar, bc, (cudaStream_t)stream);
}

int p2b_fill_ghost_i64(int64_t* base, const p2b_grid* g, int nvar, const int* bc, void* stream)
{
    return fill_ghost_impl<long long>((long long*)base, g, nvar, bc, (cudaStream_t)stream);
}

int p2b_fill_ghost_values_f64(double* plane, const p2b_grid* g, const int bc[4], const double* xl,
                              const double* xr, const double* yl, const double* yr, void* stream)
{
    P2B_REQUIRE(plane && g && bc, "null pointer");
    P2B_REQUIRE(g->ng >= 1, "ng < 1");
    cudaStream_t st = (cudaStream_t)stream;
    // homogeneous part for the sides without values; sides with values are handled below.
    // The reference fills either the homogeneous or the inhomogeneous form per side
    // (array_indexer.py:165-183), x sides before y sides.
    int codes[4] = {xl ? P2B_BC_NONE : bc[0], xr ? P2B_BC_NONE : bc[1], P2B_BC_NONE, P2B_BC_NONE};
    const int qx = g->nx + 2 * g->ng, qy = g->ny + 2 * g->ng;
    BcTable t;
    memset(&t, 0, sizeof t);
    for (int s = 0; s < 4; ++s) t.code[0][s] = codes[s];
    int tx = 2 * g->ng * qy, ty = qx * 2 * g->ng;
    P2B_LAUNCH(fill_x_kernel<double>, (tx + 255) / 256, 256, 0, st)(plane, *g, 1, t);
    if (xl || xr) P2B_LAUNCH(fill_x_values_kernel, (qy + 255) / 256, 256, 0, st)(plane, *g, bc[0], bc[1], xl, xr);
    t.code[0][0] = t.code[0][1] = P2B_BC_NONE;
    t.code[0][2] = yl ? P2B_BC_NONE : bc[2];
    t.code[0][3] = yr ? P2B_BC_NONE : bc[3];
    P2B_LAUNCH(fill_y_kernel<double>, (ty + 255) / 256, 256, 0, st)(plane, *g, 1, t);
    if (yl || yr) P2B_LAUNCH(fill_y_values_kernel, (qx + 255) / 256, 256, 0, st)(plane, *g, bc[2], bc[3], yl, yr);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

int p2b_cfl_wavemax(const double* U, const p2b_grid* g, double gamma, uint64_t* scratch, void* stream)
{
    P2B_REQUIRE(U && g && scratch, "null pointer");
    const long long total = (long long)(g->nx + 2 * g->ng) * (g->ny + 2 * g->ng);
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    P2B_LAUNCH(cfl_kernel, (int)blocks, 256, 0, (cudaStream_t)stream)(U, *g, gamma, (unsigned long long*)scratch);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

}  // extern "C"

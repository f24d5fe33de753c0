// mg_kernels.cuh -- device code of the multigrid V-cycle (HP-2): kernels and the per-point arithmetic,
// no runtime-API calls.  Included by mg.cu (nvcc) and by tests/emu/mg_emu.cpp, which compiles the very
// same kernels for the host through tests/emu/cuda_emu.h (threads = pthreads, __syncthreads = barrier,
// shuffles = warp slots) so that they can be checked against the oracle without a GPU.
//
// Reference behaviour (pyro2, file:line):
//   CellCenterMG2d.smooth            pyro/multigrid/MG.py:544-599   red-black Gauss-Seidel, ghost fill
//                                                                   after each colour
//   CellCenterMG2d._compute_residual pyro/multigrid/MG.py:529-542
//   CellCenterData2d.restrict        pyro/mesh/patch.py:640-676     4-cell average
//   CellCenterData2d.prolong         pyro/mesh/patch.py:678-736     centred (unlimited) slopes
//   ArrayIndexer.norm                pyro/mesh/array_indexer.py:98-111
//   VarCoeffCCMG2d.smooth/_compute_residual   pyro/multigrid/variable_coeff_MG.py:112-213
//   EdgeCoeffs                       pyro/multigrid/edge_coeffs.py:1-54
#pragma once
#include <math.h>
#include <string.h>
#include <type_traits>

#include "../../include/pyro2b200.h"
#include "hydro_core.cuh"
#include "peer_comm.cuh"

namespace pyro {

struct MgLevel {
    int n, pitch;       // columns (y) of the level and the row pitch
    double *v, *f, *r;
    double *w;          // scratch plane: ping-pong target of the temporally blocked smoother
    double dx, dy;
    // x-slab decomposition (multi-GPU): this rank owns global rows ioff+1 .. ioff+ni; gx halo rows
    // are stored beyond each end (row index 1-gx .. ni+gx).  Single GPU / replicated level:
    // ni = n, ioff = 0, gx = 1, both x sides physical.
    int ni, ioff, gx;
    int xlo_phys, xhi_phys;
    const unsigned long long* ctl;   // control words of the hierarchy (CW_STOP makes every kernel a no-op) or NULL
};

struct MgBC {
    int xl, xr, yl, yr;                       // P2B_BC_* codes
    const double *xlv, *xrv, *ylv, *yrv;      // inhomogeneous values (finest level) or NULL
};

constexpr int MG_MAX_LEVELS = 24;
constexpr int MG_NPART = 16384;               // partial sums of the deterministic norms (one per row)

// ---- peer-memory communication between the x-slabs of a decomposed hierarchy ---------------------------------
// Every rank's workspace (planes + a small control area) has the SAME layout and is mapped into every other rank's
// address space (cudaIpc; plain pointers when the ranks share a process), so "the same location on rank r" is a
// constant element offset from a local pointer.  Halo rows are written straight into the neighbour's planes by the
// kernel that produces them (the smoother's epilogue, restrict, prolong): the transfer needs no extra launch and
// overlaps with the producer's interior tiles; there is no NCCL call and no host involvement inside a V-cycle, so
// the whole cycle is capturable as one CUDA graph per rank.
//
// Ordering: one monotone 64-bit word per direction.  All ranks run the same sequence of "programs" (a V-cycle, a
// diagnostics all-reduce, a stand-alone exchange); a program starts by incrementing the rank's epoch, and each
// pushing launch of a program has a host-assigned ordinal, so value = epoch * MG_ORD_STRIDE + ordinal grows along
// every rank's stream.  A producer's last edge CTA (counted with a local atomic after a system-scope fence) stores
// that value into the neighbour's FROM_LO / FROM_HI word; a consumer's CTAs that read halo rows spin (thread 0,
// acquire loads, bounded by a time-out that raises the error word instead of hanging) until the word has reached the
// value of the launch that filled those rows.  Kernels run in stream order on every rank, so "flag >= value of
// launch k" means every push of launches <= k has landed.  Why no write can overtake a reader is argued launch by
// launch in DESIGN.md (multi-GPU section): in short, a rank only writes a halo after it has waited on a signal the
// neighbour issued after its last reader of that halo.
constexpr int MG_MAX_RANKS = 16;
constexpr unsigned long long MG_ORD_STRIDE = 1024;
enum {
    CW_EPOCH = 0,        // programs started on this rank
    CW_ERR = 1,          // non-zero: a wait timed out (1 + the word waited on)
    CW_STOP = 2,         // non-zero: solve() has converged; the kernels of a speculatively enqueued cycle return at once
    CW_FROM_LO = 3,      // written by the lo neighbour: value of its last completed push towards me
    CW_FROM_HI = 4,      // ... by the hi neighbour
    CW_CNT_LO = 5, CW_CNT_HI = 6, CW_CNT_ALL = 7,    // local counters of pushing CTAs (reset by the last one)
    CW_GFLAG = 8,        // [MG_MAX_RANKS] written by rank r: value of its last completed push to ALL ranks
    CW_PEER = 24,        // [MG_MAX_RANKS] element offsets from my workspace to rank r's (host-written)
    CW_SUMS = 40,        // doubles [2][MG_MAX_RANKS][4]: all-reduce slots, double-buffered by epoch parity
    CW_RESULT = 168,     // doubles [4]: the last diagnostics (relsq, rsq, residual_error, cycles run)
    CW_STOPPAR = 172,    // doubles [4]: the stopping rule (source_norm, rtol, max_cycles, enabled) -- in memory, not a
                         // kernel argument: a captured cycle is replayed by later solve() calls with other values
    CW_WORDS = 176
};

struct MgComm {
    unsigned long long* ctl;     // this rank's control words; NULL on a single GPU (no communication code runs)
    long long dlo, dhi;          // element offsets to the lo / hi neighbour's workspace
    int has_lo, has_hi;
    int rank, size;
    // per launch
    int wait_ord;                // >= 0: CTAs that read halo rows wait for the neighbour's word >= epoch * STRIDE + wait_ord
    int sig_ord;                 // >= 0: this launch pushes halo rows; value it signals
    int n_lo, n_hi;              // CTAs that push towards the lo / hi neighbour (the last to finish signals)
};

}  // namespace pyro

namespace pyro {

__device__ __forceinline__ unsigned long long comm_value(const MgComm& c, int ord)
{
    return c.ctl[CW_EPOCH] * MG_ORD_STRIDE + (unsigned long long)ord;
}

// Tile row a CTA works on.  CTAs are dispatched in blockIdx order; with communication the tile rows are rotated so that the
// first and the last -- the ones that wait for the neighbours' halo rows and push their own -- are dispatched in the MIDDLE
// of the launch: late enough that the rows pushed in the middle of the neighbours' previous pass have long arrived (no SM
// sits in a spin loop: dispatched first, the edge tiles held ~1/3 of the SMs for most of a pass at N = 2), early enough
// that their system-scope fences and flag stores are hidden behind the remaining interior tiles instead of forming the
// tail of the kernel.
__device__ __forceinline__ int comm_tile_row(const MgComm& c)
{
    return c.ctl ? (int)((blockIdx.y + gridDim.y / 2u + 1u) % gridDim.y) : (int)blockIdx.y;
}

// all threads of a CTA call this (block-uniform arguments): wait for the halo rows this CTA is about to read
__device__ __forceinline__ void comm_block_wait(const MgComm& c, bool need_lo, bool need_hi)
{
    if (!c.ctl || c.wait_ord < 0) return;
    need_lo = need_lo && c.has_lo; need_hi = need_hi && c.has_hi;
    if (!(need_lo || need_hi)) return;
    if (comm_tid() == 0) {
        const unsigned long long target = comm_value(c, c.wait_ord);
        if (need_lo) comm_wait_ge(c.ctl, CW_FROM_LO, target);
        if (need_hi) comm_wait_ge(c.ctl, CW_FROM_HI, target);
    }
    __syncthreads();
}

// all threads of a CTA call this after their pushes (block-uniform arguments): the last pushing CTA of each side
// tells the neighbour
__device__ __forceinline__ void comm_block_signal(const MgComm& c, bool lo_edge, bool hi_edge)
{
    if (!c.ctl || c.sig_ord < 0) return;
    lo_edge = lo_edge && c.has_lo; hi_edge = hi_edge && c.has_hi;
    if (!(lo_edge || hi_edge)) return;
    __syncthreads();
    if (comm_tid() == 0) {
        __threadfence_system();
        const unsigned long long val = comm_value(c, c.sig_ord);
        // (every pushing CTA fenced at system scope before it counted itself: when the last one sees the full count, all
        // pushes have been performed; its release store orders the observation before the flag)
        if (lo_edge && atomicAdd(c.ctl + CW_CNT_LO, 1ull) == (unsigned long long)(c.n_lo - 1)) {
            atomicExch(c.ctl + CW_CNT_LO, 0ull);
            comm_store(c.ctl + c.dlo + CW_FROM_HI, val);       // I am my lo neighbour's hi neighbour
        }
        if (hi_edge && atomicAdd(c.ctl + CW_CNT_HI, 1ull) == (unsigned long long)(c.n_hi - 1)) {
            atomicExch(c.ctl + CW_CNT_HI, 0ull);
            comm_store(c.ctl + c.dhi + CW_FROM_LO, val);
        }
    }
}

// the same for a launch whose every CTA pushes to EVERY rank (slab -> replicated transfer): the last CTA signals all
__device__ __forceinline__ void comm_block_signal_all(const MgComm& c, int nblocks)
{
    if (!c.ctl || c.sig_ord < 0) return;
    __syncthreads();
    if (comm_tid() == 0) {
        __threadfence_system();
        if (atomicAdd(c.ctl + CW_CNT_ALL, 1ull) == (unsigned long long)(nblocks - 1)) {
            atomicExch(c.ctl + CW_CNT_ALL, 0ull);
            __threadfence_system();
            const unsigned long long val = comm_value(c, c.sig_ord);
            for (int r = 0; r < c.size; ++r)
                comm_store_relaxed(c.ctl + (long long)c.ctl[CW_PEER + r] + CW_GFLAG + c.rank, val);
        }
    }
}

// number of blocks by in [0, nby) whose rows [by * bh + 1, by * bh + bh] meet [a, b]  (host and device agree on
// which CTAs push: the producer counts arrivals against this number)
__host__ __device__ inline bool rows_meet(int by, int bh, int a, int b) { return by * bh + 1 <= b && by * bh + bh >= a; }
inline int count_rows_meet(int nby, int bh, int a, int b)
{
    int k = 0;
    for (int by = 0; by < nby; ++by) k += rows_meet(by, bh, a, b) ? 1 : 0;
    return k;
}

}  // namespace pyro

namespace pyro {

// ---- ghost update fused into the writers --------------------------------------------------------
// value of the ghost cell generated from interior source value `val` (array_indexer.py:164-274)
__device__ __forceinline__ double ghost_lo(double val, int code, const double* vals, int idx, double h)
{
    if (vals) {
        if (code == P2B_BC_OUTFLOW) return exact_sub(val, exact_mul(h, vals[idx]));
        if (code == P2B_BC_REFLECT_ODD) return exact_sub(exact_mul(2.0, vals[idx]), val);
    }
    return code == P2B_BC_REFLECT_ODD ? -val : val;
}

__device__ __forceinline__ double ghost_hi(double val, int code, const double* vals, int idx, double h)
{
    if (vals) {
        if (code == P2B_BC_OUTFLOW) return exact_add(val, exact_mul(h, vals[idx]));
        if (code == P2B_BC_REFLECT_ODD) return exact_sub(exact_mul(2.0, vals[idx]), val);
    }
    return code == P2B_BC_REFLECT_ODD ? -val : val;
}

// store v(i,j) = val and every ghost cell whose source is (i,j).  x ghosts are functions of the
// interior value; y ghosts (filled second in the reference, over the full x range) are functions of
// the already x-filled column, which gives the corner values.
__device__ __forceinline__ void store_with_ghosts(double* v, int ni, int n, int pitch, int i, int j, double val,
                                                  const MgBC& b, double dx, double dy, int ioff = 0)
{
    v[(long long)i * pitch + j] = val;
    // x sides with P2B_BC_NONE face another slab: their halo rows come from the neighbour
    const int sxl = (b.xl == P2B_BC_PERIODIC) ? ni : 1;   // source row of ghost row 0
    const int sxh = (b.xr == P2B_BC_PERIODIC) ? 1 : ni;   // source row of ghost row ni+1
    const int syl = (b.yl == P2B_BC_PERIODIC) ? n : 1;
    const int syh = (b.yr == P2B_BC_PERIODIC) ? 1 : n;
    const bool lo = (b.xl != P2B_BC_NONE) && (i == sxl), hi = (b.xr != P2B_BC_NONE) && (i == sxh);
    if (!(lo || hi || j == syl || j == syh)) return;       // interior cell: nothing else to write
    // after the x fill this value lives in up to three rows: i, 0 (if lo), ni+1 (if hi)
    double glo = 0.0, ghi = 0.0;
    if (lo) { glo = ghost_lo(val, b.xl, b.xlv, j, dx); v[j] = glo; }
    if (hi) { ghi = ghost_hi(val, b.xr, b.xrv, j, dx); v[(long long)(ni + 1) * pitch + j] = ghi; }
    // y-side values are indexed with the global row; a slab's halo row across the periodic x boundary wraps
    const int grow = ioff + i < 1 ? ioff + i + n : (ioff + i > n ? ioff + i - n : ioff + i);
    if (j == syl) {
        v[(long long)i * pitch] = ghost_lo(val, b.yl, b.ylv, grow, dy);
        if (lo) v[0] = ghost_lo(glo, b.yl, b.ylv, ioff, dy);
        if (hi) v[(long long)(ni + 1) * pitch] = ghost_lo(ghi, b.yl, b.ylv, ioff + ni + 1, dy);
    }
    if (j == syh) {
        v[(long long)i * pitch + n + 1] = ghost_hi(val, b.yr, b.yrv, grow, dy);
        if (lo) v[n + 1] = ghost_hi(glo, b.yr, b.yrv, ioff, dy);
        if (hi) v[(long long)(ni + 1) * pitch + n + 1] = ghost_hi(ghi, b.yr, b.yrv, ioff + ni + 1, dy);
    }
}

// rden = RN(1/denom), fast = 1 when q = RN(a*rden); r = fma(-denom, q, a); RN(q + r*rden) is the
// correctly rounded a/denom (Markstein's theorem: holds unless denom's significand is all ones) --
// three DP instructions and no branch instead of the ~10 + slow path of a true division, with the
// SAME bits as the reference's division.
struct SmoothCoef { double alpha, xc, yc, denom, rden; int fast; };
struct DivConst { double d, rd; int fast; };     // exact a / d for a loop-invariant d (same trick)

__device__ __forceinline__ double div_const(double a, const DivConst& k)
{
    if (!k.fast) return exact_div(a, k.d);
    double q = exact_mul(a, k.rd);
    double r = __fma_rn(-k.d, q, a);
    return __fma_rn(r, k.rd, q);
}

static DivConst make_div_const(double d)
{
    DivConst k;
    k.d = d; k.rd = 1.0 / d;
    unsigned long long bits;
    memcpy(&bits, &d, 8);
    k.fast = ((bits & 0xFFFFFFFFFFFFFULL) != 0xFFFFFFFFFFFFFULL) && isfinite(k.rd) && d != 0.0 &&
             fabs(d) > 1e-290 && fabs(d) < 1e290;
    return k;
}

__device__ __forceinline__ double div_by_denom(double a, const SmoothCoef& c)
{
    if (!c.fast) return exact_div(a, c.denom);
    double q = exact_mul(a, c.rden);
    double r = __fma_rn(-c.denom, q, a);
    return __fma_rn(r, c.rden, q);
}

__device__ __forceinline__ double gs_update(const double* v, const double* f, int pitch, int i, int j,
                                            const SmoothCoef& c)
{
    // MG.py:593-596:  (f + xcoeff*(v[i+1]+v[i-1]) + ycoeff*(v[j+1]+v[j-1])) / (alpha + 2xc + 2yc)
    const long long k = (long long)i * pitch + j;
    double sx = exact_add(v[k + pitch], v[k - pitch]);
    double sy = exact_add(v[k + 1], v[k - 1]);
    double num = exact_add(exact_add(f[k], exact_mul(c.xc, sx)), exact_mul(c.yc, sy));
    return div_by_denom(num, c);
}

// one colour of one red-black iteration; colour 0 = (i+j) even = the reference's groups (0,0),(1,1)
__global__ void mg_halfsweep_kernel(MgLevel L, MgBC b, SmoothCoef c, int colour)
{
    if (L.ctl && L.ctl[CW_STOP]) return;
    const int half = L.n >> 1;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y + 1;
    if (k >= half || i > L.n) return;
    const int j = 1 + ((i + 1 + colour) & 1) + 2 * k;
    double val = gs_update(L.v, L.f, L.pitch, i, j, c);
    store_with_ghosts(L.v, L.n, L.n, L.pitch, i, j, val, b, L.dx, L.dy);
}

// whole smooth() for a small level in one CTA (global memory, __syncthreads between colours)
__global__ void mg_smooth_small_kernel(MgLevel L, MgBC b, SmoothCoef c, int nsmooth)
{
    if (L.ctl && L.ctl[CW_STOP]) return;
    const int half = L.n >> 1;
    const int npts = L.n * half;
    for (int it = 0; it < 2 * nsmooth; ++it) {
        const int colour = it & 1;
        for (int t = threadIdx.x; t < npts; t += blockDim.x) {
            int i = t / half + 1, k = t % half;
            int j = 1 + ((i + 1 + colour) & 1) + 2 * k;
            double val = gs_update(L.v, L.f, L.pitch, i, j, c);
            store_with_ghosts(L.v, L.n, L.n, L.pitch, i, j, val, b, L.dx, L.dy);
        }
        __syncthreads();
    }
}


// ---- variable coefficients: div(eta grad phi) = f  (VarCoeffCCMG2d) --------------------------------
// eta_x[i, j] = eta_{i-1/2, j} / dx^2 and eta_y[i, j] = eta_{i, j-1/2} / dy^2 live in two planes with the
// level's own pitch (edge_coeffs.py:16-26); entries outside [1, n+1]^2 are zero, as in the reference.
struct VcEdges { const double* ex; const double* ey; };

// variable_coeff_MG.py:150-164 for one point; up / dn = phi(i-1) / phi(i+1), lf / rt = phi(j-1) / phi(j+1),
// exl / exh = eta_x at i / i+1, eyl / eyh = eta_y at j / j+1
__device__ __forceinline__ double vc_gs_value(double f, double up, double dn, double lf, double rt,
                                              double exl, double exh, double eyl, double eyh)
{
    double denom = exact_add(exact_add(exact_add(exh, exl), eyh), eyl);
    double num = exact_add(exact_add(exact_add(exact_add(-f, exact_mul(exh, dn)), exact_mul(exl, up)),
                                     exact_mul(eyh, rt)), exact_mul(eyl, lf));
    return exact_div(num, denom);
}

// variable_coeff_MG.py:199-213:  f - L_eta phi
__device__ __forceinline__ double vc_residual_value(double f, double c, double up, double dn, double lf, double rt,
                                                    double exl, double exh, double eyl, double eyh)
{
    double a = exact_mul(exh, exact_sub(dn, c));
    double b = exact_mul(exl, exact_sub(c, up));
    double d = exact_mul(eyh, exact_sub(rt, c));
    double e = exact_mul(eyl, exact_sub(c, lf));
    return exact_sub(f, exact_sub(exact_add(exact_sub(a, b), d), e));
}


// 8-byte asynchronous global -> shared copy (cp.async / LDGSTS) and the wait for all of a thread's
// copies; visibility to the other threads still needs the __syncthreads that follows
__device__ __forceinline__ void tile_copy8(double* dst_shared, const double* src_global)
{
#ifdef P2B_EMU_HEADER
    *dst_shared = *src_global;
#else
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_shared);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src_global) : "memory");
#endif
}

__device__ __forceinline__ void tile_copy_wait()
{
#ifndef P2B_EMU_HEADER
    asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}

// ---- temporally blocked smoother -------------------------------------------------------------------
// One CTA owns a TI x TJ tile and loads it with a halo of H = 2*TB_K cells into REGISTERS: a thread
// holds a 2-column x TB_R-row patch of v and f (lane t <-> columns 2t, 2t+1 of the 64-column
// region; warp w <-> rows w*TB_R ...).  It then runs up to TB_K full red-black iterations without
// touching global memory: x-neighbours (rows) come from the thread's own registers or, across
// warps, from a small double-buffered shared row exchange; y-neighbours from the pair partner or
// the adjacent lane (shuffle).  Halo cells are updated redundantly; the error of not knowing what is
// outside the region advances one cell per half-sweep and never reaches the tile.  Ghost cells are
// not stored at all while blocking: at a domain edge the neighbour is computed from the cell's own
// value with the same formula fill_BC uses (ghost = g(inner)), which is exactly the value the
// reference's fill_BC-after-every-colour keeps there; periodic sides wrap on load.  Every point update
// executes the reference's operations in the reference's order, so the result is bit-identical to
// nsmooth separate half-sweeps; HBM traffic per pass is ~42 B per cell for 5 iterations instead of
// 5 x 48 B.  Reads vin, writes vout (another plane): neighbouring CTAs read each other's tiles.
constexpr int TB_K = 5;                 // iterations per pass
constexpr int TB_H = 2 * TB_K;          // halo
constexpr int TB_R = 8;                 // rows per thread
#ifndef TB_NW_CFG
#define TB_NW_CFG 16       // 16 warps = 128 x 64 region, 108 x 44 tile (measured 7% faster than 8 warps / 44 x 44)
#endif
constexpr int TB_NW = TB_NW_CFG;        // warps per CTA
constexpr int TB_RH = TB_R * TB_NW;     // region rows  (64)
constexpr int TB_RW = 64;               // region columns
constexpr int TB_TI = TB_RH - 2 * TB_H; // tile rows    (44)
constexpr int TB_TJ = TB_RW - 2 * TB_H; // tile columns (44)
static_assert(TB_TI % 2 == 0 && TB_TJ % 2 == 0 && TB_R % 2 == 0, "parity bookkeeping needs even tile sizes");
static_assert(TB_RW == 64 && TB_NW % 2 == 0, "the coefficient-tile copy maps thread t to column t & 63");

__device__ __forceinline__ int wrap1(int i, int n)   // periodic image of i in [1, n]; n is a power of two (levels
{                                                    // have 2^k columns and slabs 2^k / 2^m rows)
    return ((i - 1) & (n - 1)) + 1;
}

// EDGE = false: the whole 64 x 64 region lies strictly inside the domain (the case for all but the
// outermost ring of CTAs): no wrap, no ghost logic, no per-cell predicates.  EDGE = true: the general
// path.  The choice is block-uniform, so a CTA executes exactly one of the two instruction streams.
//
// VC = true (variable coefficients): the CTA additionally stages the region's edge coefficients in
// shared memory once per pass -- eta_x for region rows 0..TB_RH (a cell needs its own and the next
// row's), eta_y for columns 0..TB_RW -- split by column parity so that the lanes of a warp read
// consecutive doubles: exs[col & 1][row][col >> 1], eys[col & 1][row][col >> 1] (33 entries per row).
// A periodic image of cell n takes eta at n + 1 (not at the image of n + 1, which is only the same
// number when the coefficient's own BC is periodic): those seam cells read it from global memory.
constexpr int TB_EXS = 2 * (TB_RH + 1) * 32;       // doubles in the eta_x tile
constexpr int TB_EYS = 2 * TB_RH * 33;             // doubles in the eta_y tile

// INHOM = false: no side carries inhomogeneous boundary values (every level but the finest, and the finest of most
// callers): the ghost a cell generates is +-(its own value) and none of the index arithmetic for the value tables is
// compiled in.  cm: slab communication of this launch (cm.ctl == NULL: none).
// R / NW: rows per thread and warps per CTA of this instantiation (region R * NW rows x 64 columns).  The constants of
// the default geometry are shadowed inside the body, which is otherwise written in terms of TB_R, TB_NW, TB_RH, TB_TI.
// First row of tile row ty.  clamp (a slab whose hi side faces another slab): a last tile row that would overhang is moved
// up so that it ends at the last owned row and its region at the last halo row -- all real cells, the branch-free path
// applies; the rows it shares with the tile row before it are computed twice with the same bits.  Host and device use the
// same function: the host counts the tile rows that push (the last one to finish signals).
__host__ __device__ inline int tb_tile_origin(int ty, int TI, int ni, bool clamp)
{
    int I0 = 1 + ty * TI;
    if (clamp && ni >= TI && I0 + TI - 1 > ni) I0 = ni - TI + 1;
    return I0;
}
__host__ __device__ inline bool tb_pushes_lo(int I0) { return I0 <= TB_H; }
__host__ __device__ inline bool tb_pushes_hi(int I0, int TI, int ni) { return I0 + TI - 1 > ni - TB_H && I0 <= ni; }

// COMM: this CTA takes part in the slab communication of the launch (waits for halo rows its region reaches into, pushes
// the first / last owned rows it stores); independent of EDGE, so that a tile at a slab boundary whose region holds real
// cells only -- halo rows included -- runs the branch-free path: at N = 8 two of the five tile rows of the finest level are
// such tiles.  I0: first row of the tile (the kernel clamps the last tile row of a slab so that it does not overhang).
template <bool EDGE, bool VC, bool INHOM, int R = 8, int NW = TB_NW_CFG, bool COMM = EDGE>
__device__ __forceinline__ void smooth_tb_body(const MgLevel& L, const double* __restrict__ vin,
                                               double* __restrict__ vout, const MgBC& bb, const SmoothCoef& c,
                                               int niter, double (*edge)[NW][2][TB_RW],
                                               const VcEdges& E, double* __restrict__ exs, double* __restrict__ eys,
                                               const MgComm& cm, int I0)
{
    constexpr int TB_R = R, TB_NW = NW, TB_RH = R * NW, TB_TI = TB_RH - 2 * TB_H;
    static_assert(TB_TI % 2 == 0 && TB_TI > 0 && R % 2 == 0 && (!VC || (R == 8 && NW == TB_NW_CFG)), "tile geometry");
    MgBC b = bb;
    if (!INHOM) { b.xlv = nullptr; b.xrv = nullptr; b.ylv = nullptr; b.yrv = nullptr; }
    const int n = L.n, ni = L.ni, P = L.pitch;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int J0 = 1 + blockIdx.x * TB_TJ;
    const int gi0 = I0 - TB_H + w * TB_R;          // first region row of this thread (odd; local index)
    const int gj0 = J0 - TB_H + 2 * lane;          // first of its two columns (odd)
    const bool xper = EDGE && (b.xl == P2B_BC_PERIODIC), yper = EDGE && (b.yl == P2B_BC_PERIODIC);
    // rows that hold real cells: the owned rows plus, on a side facing another slab, the TB_H halo
    // rows received from it (they are updated redundantly, exactly like the periodic wrap)
    const int rlo = L.xlo_phys ? 1 : 1 - TB_H, rhi = L.xhi_phys ? ni : ni + TB_H;

    // slab: this CTA's region reaches into halo rows the neighbour writes -- wait until they have landed
    if (COMM) comm_block_wait(cm, I0 - TB_H < 1, I0 - TB_H + TB_RH - 1 > ni);

    unsigned xseam = 0;                            // VC: bit r: this row is the periodic image of row ni
    bool yseam[2] = {false, false};                //     column a is the periodic image of column n
    if (VC) {
        // Stage the coefficient tiles first, as asynchronous global -> shared copies (LDGSTS) that need
        // no registers and all fly at once; the v / f loads below overlap with them.  (The first
        // version loaded through registers in a rolled loop: ncu showed 40% of the kernel's stall
        // samples on those stores waiting for one load at a time.)  Thread t handles column t & 63 of
        // rows t >> 6, + TB_NW / 2, ...; the 65th eta_y column is done by the first TB_RH threads.
        const int rbase = I0 - TB_H, cbase = J0 - TB_H;
        const int cc = threadIdx.x & (TB_RW - 1);
        int sj = cbase + cc;
        if (yper) sj = wrap1(sj, n);
        const bool colx = !EDGE || (sj >= 1 && sj <= n);            // eta_x exists for columns 1..n
        const bool coly = !EDGE || (sj >= 1 && sj <= n + 1);        // eta_y for columns 1..n+1
        double* exd = exs + (cc & 1) * (TB_RH + 1) * 32 + (cc >> 1);
        double* eyd = eys + (cc & 1) * TB_RH * 33 + (cc >> 1);
#pragma unroll
        for (int k = 0; k < (TB_RH + TB_NW / 2) / (TB_NW / 2); ++k) {
            const int rr = (threadIdx.x >> 6) + k * (TB_NW / 2);
            if (rr > TB_RH) break;
            int si = rbase + rr;
            if (xper) si = wrap1(si, ni);
            const long long src = (long long)si * P + sj;
            if (colx && (!EDGE || (si >= 1 && si <= ni + 1))) tile_copy8(exd + rr * 32, E.ex + src);
            else exd[rr * 32] = 0.0;
            if (rr < TB_RH) {
                if (coly && (!EDGE || (si >= 1 && si <= ni))) tile_copy8(eyd + rr * 33, E.ey + src);
                else eyd[rr * 33] = 0.0;
            }
        }
        if (threadIdx.x < TB_RH) {
            const int rr = threadIdx.x;
            int si = rbase + rr, sjl = cbase + TB_RW;
            if (xper) si = wrap1(si, ni);
            if (yper) sjl = wrap1(sjl, n);
            double* d = eys + rr * 33 + (TB_RW >> 1);               // column TB_RW: parity 0, entry 32
            if (!EDGE || (si >= 1 && si <= ni && sjl >= 1 && sjl <= n + 1)) tile_copy8(d, E.ey + (long long)si * P + sjl);
            else *d = 0.0;
        }
        if (EDGE) {
#pragma unroll
            for (int r = 0; r < TB_R; ++r)
                if (xper && wrap1(gi0 + r, ni) == ni) xseam |= 1u << r;
#pragma unroll
            for (int a = 0; a < 2; ++a) yseam[a] = yper && wrap1(gj0 + a, n) == n;
        }
    }

    double v[TB_R][2], f[TB_R][2];
    // EDGE only: bit r*2+a set = (r, a) is a real interior cell; per-row / per-column edge flags
    unsigned inmask = 0xffffffffu;
    unsigned row_lo = 0, row_hi = 0;               // bit r: global row is 1 / n (non-periodic x)
    bool col_lo[2] = {false, false}, col_hi[2] = {false, false};
#pragma unroll
    for (int r = 0; r < TB_R; ++r) {
        int gi = gi0 + r;
        int si = xper ? wrap1(gi, ni) : gi;
        if (EDGE && !xper) {
            if (L.xlo_phys && gi == 1) row_lo |= 1u << r;
            if (L.xhi_phys && gi == ni) row_hi |= 1u << r;
        }
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            int gj = gj0 + a;
            int sj = yper ? wrap1(gj, n) : gj;
            bool ok = !EDGE || (si >= rlo && si <= rhi && sj >= 1 && sj <= n);
            if (!ok) inmask &= ~(1u << (2 * r + a));
            long long k = (long long)si * P + sj;
            v[r][a] = ok ? vin[k] : 0.0;
            f[r][a] = ok ? L.f[k] : 0.0;
        }
    }
    if (EDGE && !yper) {
#pragma unroll
        for (int a = 0; a < 2; ++a) { col_lo[a] = (gj0 + a == 1); col_hi[a] = (gj0 + a == n); }
    }

    // warp-uniform: does this warp hold a boundary row / does any of its lanes hold a boundary column?
    const bool row_edge = EDGE && (row_lo | row_hi) != 0u;
#ifdef P2B_EMU_HEADER
    const bool col_edge = EDGE;                    // (the emulator has no warp vote: always take the guarded selects)
#else
    const bool col_edge = EDGE && __any_sync(0xffffffffu, col_lo[0] || col_lo[1] || col_hi[0] || col_hi[1]);
#endif

    if (VC) tile_copy_wait();                      // the coefficient tiles requested above have landed

    double uph[2] = {0.0, 0.0}, dnh[2] = {0.0, 0.0};   // rows just above / below this thread's strip
    auto publish = [&](int buf) {
        *reinterpret_cast<double2*>(&edge[buf][w][0][2 * lane]) = make_double2(v[0][0], v[0][1]);
        *reinterpret_cast<double2*>(&edge[buf][w][1][2 * lane]) = make_double2(v[TB_R - 1][0], v[TB_R - 1][1]);
        __syncthreads();
        if (w > 0) {
            double2 t = *reinterpret_cast<const double2*>(&edge[buf][w - 1][1][2 * lane]);
            uph[0] = t.x; uph[1] = t.y;
        }
        if (w < TB_NW - 1) {
            double2 t = *reinterpret_cast<const double2*>(&edge[buf][w + 1][0][2 * lane]);
            dnh[0] = t.x; dnh[1] = t.y;
        }
    };
    publish(0);

    // one colour of one iteration; the colour is a compile-time constant so that every register
    // array index is static.  colour 0 = (i + j) even; gi0 and gj0 are both odd (tile origins are
    // 1 + even multiples, H even, strip offsets even), so cell (r, a) has colour (r + a) & 1.
    auto half_sweep = [&](auto colour_tag, int buf) {
        constexpr int colour = decltype(colour_tag)::value;
        double nv[TB_R];
#pragma unroll
        for (int r = 0; r < TB_R; ++r) {
            constexpr int dummy = 0; (void)dummy;
            const int a = (r + colour) & 1;          // active column of the pair in this row (static)
            const double self = v[r][a];
            double up = (r == 0) ? uph[a] : v[r == 0 ? 0 : r - 1][a];
            double dn = (r == TB_R - 1) ? dnh[a] : v[r == TB_R - 1 ? r : r + 1][a];
            // y-neighbours: the pair partner, or the adjacent lane's facing column
            double lf = (a == 0) ? __shfl_up_sync(0xffffffffu, v[r][1], 1) : v[r][0];
            double rt = (a == 0) ? v[r][1] : __shfl_down_sync(0xffffffffu, v[r][0], 1);
            if (EDGE) {
                // domain edges: the ghost value fill_BC would hold, from the cell's own current value
                // (inhomogeneous boundary values are indexed by the cell's own row / column: a halo cell that is the
                // periodic image of an interior cell uses that cell's index)
                // The y-side values are indexed with the GLOBAL row: on a slab, a halo row received from the
                // neighbour across the periodic x boundary lies outside 1..n before the wrap.
                int gj = 0, grow = 0;
                if (INHOM) {
                    const int gi = xper ? wrap1(gi0 + r, ni) : gi0 + r;
                    gj = yper ? wrap1(gj0 + a, n) : gj0 + a;
                    grow = L.ioff + gi;
                    grow = grow < 1 ? grow + n : (grow > n ? grow - n : grow);
                }
                // (warp-uniform guards: a warp's rows are the same for all lanes, and whether any lane holds a boundary
                // column was voted once -- most warps of a boundary tile touch no boundary and skip the selects)
                if (row_edge) {
                    if ((row_lo >> r) & 1u) up = ghost_lo(self, b.xl, b.xlv, gj, L.dx);
                    if ((row_hi >> r) & 1u) dn = ghost_hi(self, b.xr, b.xrv, gj, L.dx);
                }
                if (col_edge) {
                    if (col_lo[a]) lf = ghost_lo(self, b.yl, b.ylv, grow, L.dy);
                    if (col_hi[a]) rt = ghost_hi(self, b.yr, b.yrv, grow, L.dy);
                }
            }
            if (VC) {
                const int rr = w * TB_R + r;
                double exl = exs[(a * (TB_RH + 1) + rr) * 32 + lane];
                double exh = exs[(a * (TB_RH + 1) + rr + 1) * 32 + lane];
                double eyl = eys[(a * TB_RH + rr) * 33 + lane];
                double eyh = eys[((a ^ 1) * TB_RH + rr) * 33 + lane + a];
                if (EDGE && ((inmask >> (2 * r + a)) & 1u)) {
                    if ((xseam >> r) & 1u) exh = E.ex[(long long)(ni + 1) * P + (yper ? wrap1(gj0 + a, n) : gj0 + a)];
                    if (yseam[a]) eyh = E.ey[(long long)(xper ? wrap1(gi0 + r, ni) : gi0 + r) * P + n + 1];
                }
                nv[r] = vc_gs_value(f[r][a], up, dn, lf, rt, exl, exh, eyl, eyh);
            } else {
                double sx = exact_add(dn, up);
                double sy = exact_add(rt, lf);
                double num = exact_add(exact_add(f[r][a], exact_mul(c.xc, sx)), exact_mul(c.yc, sy));
                nv[r] = div_by_denom(num, c);
            }
        }
#pragma unroll
        for (int r = 0; r < TB_R; ++r) {
            const int a = (r + colour) & 1;
            if (!EDGE || ((inmask >> (2 * r + a)) & 1u)) v[r][a] = nv[r];
        }
        publish(buf);
    };

    for (int it = 0; it < niter; ++it) {
        half_sweep(std::integral_constant<int, 0>{}, 1);
        half_sweep(std::integral_constant<int, 1>{}, 0);
    }

    // store the tile (and, on the EDGE path, the ghost cells its boundary cells generate)
#pragma unroll
    for (int r = 0; r < TB_R; ++r) {
        int gi = gi0 + r;
        if (gi < I0 || gi >= I0 + TB_TI || gi > ni) continue;
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            int gj = gj0 + a;
            if (gj < J0 || gj >= J0 + TB_TJ || gj > n) continue;
            if (EDGE) store_with_ghosts(vout, ni, n, P, gi, gj, v[r][a], b, L.dx, L.dy, L.ioff);
            else vout[(long long)gi * P + gj] = v[r][a];
            // slab: my first / last TB_H rows are the neighbours' halo rows of the plane just written
            if (COMM && cm.ctl && cm.sig_ord >= 0) {
                if (cm.has_lo && gi <= TB_H) (vout + cm.dlo)[(long long)(ni + gi) * P + gj] = v[r][a];
                if (cm.has_hi && gi > ni - TB_H) (vout + cm.dhi)[(long long)(gi - ni) * P + gj] = v[r][a];
            }
        }
    }
    if (COMM) comm_block_signal(cm, tb_pushes_lo(I0), tb_pushes_hi(I0, TB_TI, ni));
}

// Geometries of the constant-coefficient pass.  The per-CTA time of a pass is set by its serial chain (load, 10
// half-sweeps with a CTA barrier each, store), not by the SM's throughput, so:
//   cfg 0  R = 8, 16 warps: 128 x 64 region, 108 x 44 tile (58% useful), 1 CTA / SM  -- big grids (many waves)
//   cfg 1  R = 4, 16 warps:  64 x 64 region,  44 x 44 tile, <= 64 registers, 2 CTAs / SM: half the chain per thread and
//          twice the CTAs -- levels whose cfg-0 grid does not fill the chip (128^2 .. 1024^2, thin slabs)
//   cfg 2  R = 8,  8 warps:  64 x 64 region,  44 x 44 tile, 2 CTAs / SM
template <int R, int NW, int MINB>
__global__ void __launch_bounds__(32 * NW, MINB)
mg_smooth_tb_kernel_t(MgLevel L, const double* __restrict__ vin, double* __restrict__ vout, MgBC b,
                      SmoothCoef c, int niter, MgComm cm)
{
    constexpr int RH = R * NW, TI = RH - 2 * TB_H;
    __shared__ __align__(16) double edge[2][NW][2][TB_RW];   // [buffer][warp][first/last row][column]
    if (L.ctl && L.ctl[CW_STOP]) return;
    const int I0 = tb_tile_origin(comm_tile_row(cm), TI, L.ni, cm.ctl && !L.xhi_phys);
    const int J0 = 1 + blockIdx.x * TB_TJ;
    // rows that hold real cells: the owned rows plus TB_H halo rows on a side that faces another slab
    const int rlo = L.xlo_phys ? 1 : 1 - TB_H, rhi = L.xhi_phys ? L.ni : L.ni + TB_H;
    // branch-free path: the whole region holds real cells (no ghost logic, no wrap, no masks)
    const bool interior = (I0 - TB_H >= rlo) && (I0 - TB_H + RH - 1 <= rhi) &&
                          (J0 - TB_H >= 1) && (J0 - TB_H + TB_RW - 1 <= L.n);
    // ... and it communicates iff it reaches into halo rows or stores first / last owned rows
    const bool comm = cm.ctl && (I0 - TB_H < 1 || I0 - TB_H + RH - 1 > L.ni || I0 <= TB_H || I0 + TI - 1 > L.ni - TB_H);
    const VcEdges none = {nullptr, nullptr};
    const bool inhom = b.xlv || b.xrv || b.ylv || b.yrv;
    if (interior && !comm) smooth_tb_body<false, false, false, R, NW, false>(L, vin, vout, b, c, niter, edge, none, nullptr, nullptr, cm, I0);
    else if (interior) smooth_tb_body<false, false, false, R, NW, true>(L, vin, vout, b, c, niter, edge, none, nullptr, nullptr, cm, I0);
    else if (inhom) smooth_tb_body<true, false, true, R, NW, true>(L, vin, vout, b, c, niter, edge, none, nullptr, nullptr, cm, I0);
    else smooth_tb_body<true, false, false, R, NW, true>(L, vin, vout, b, c, niter, edge, none, nullptr, nullptr, cm, I0);
}

struct TbCfg { int R, NW, TI, threads; };
constexpr int TB_NCFG = 3;
inline TbCfg tb_cfg(int k)
{
    const int R[TB_NCFG] = {8, 4, 8}, NW[TB_NCFG] = {16, 16, 8};
    TbCfg c;
    c.R = R[k]; c.NW = NW[k]; c.TI = R[k] * NW[k] - 2 * TB_H; c.threads = 32 * NW[k];
    return c;
}

// the same pass with the variable-coefficient stencil; dynamic shared memory: the row-exchange buffers
// followed by the two coefficient tiles (TB_VC_SMEM_BYTES, ~166 KB: one CTA per SM)
constexpr size_t TB_VC_SMEM_BYTES = (size_t)(2 * TB_NW * 2 * TB_RW + TB_EXS + TB_EYS) * sizeof(double);

__global__ void __launch_bounds__(32 * TB_NW, 1)
mg_vc_smooth_tb_kernel(MgLevel L, const double* __restrict__ vin, double* __restrict__ vout, MgBC b, VcEdges E, int niter)
{
    P2B_DYN_SMEM(double, sm);
    if (L.ctl && L.ctl[CW_STOP]) return;
    double (*edge)[TB_NW][2][TB_RW] = reinterpret_cast<double (*)[TB_NW][2][TB_RW]>(sm);
    double* exs = sm + 2 * TB_NW * 2 * TB_RW;
    double* eys = exs + TB_EXS;
    const int I0 = 1 + blockIdx.y * TB_TI, J0 = 1 + blockIdx.x * TB_TJ;
    const bool interior = (I0 - TB_H >= 1) && (I0 - TB_H + TB_RH - 1 <= L.ni) &&
                          (J0 - TB_H >= 1) && (J0 - TB_H + TB_RW - 1 <= L.n);
    SmoothCoef c;
    c.alpha = c.xc = c.yc = c.denom = c.rden = 0.0; c.fast = 0;
    MgComm cm;
    cm.ctl = nullptr; cm.sig_ord = cm.wait_ord = -1;
    if (interior) smooth_tb_body<false, true, false>(L, vin, vout, b, c, niter, edge, E, exs, eys, cm, I0);
    else smooth_tb_body<true, true, true>(L, vin, vout, b, c, niter, edge, E, exs, eys, cm, I0);
}


struct ResidCoef { double alpha, beta; DivConst dx2, dy2; };

__device__ __forceinline__ double residual_at(const MgLevel& L, long long k, const ResidCoef& rc)
{
    // MG.py:540-542:  f - alpha v + beta ((v[i-1] + v[i+1] - 2 v)/dx**2 + (v[j-1] + v[j+1] - 2 v)/dy**2)
    const double* v = L.v;
    double v2 = exact_mul(2.0, v[k]);
    double lx = div_const(exact_sub(exact_add(v[k - L.pitch], v[k + L.pitch]), v2), rc.dx2);
    double ly = div_const(exact_sub(exact_add(v[k - 1], v[k + 1]), v2), rc.dy2);
    return exact_add(exact_sub(L.f[k], exact_mul(rc.alpha, v[k])), exact_mul(rc.beta, exact_add(lx, ly)));
}

// ---- the coarse part of the V-cycle in ONE launch ---------------------------------------------------
// Levels up to 64^2 do not have enough points to fill the chip and every kernel on them is pure
// launch + memory latency (ncu r1: 26 us per smooth(), ~0.35 ms per V-cycle in total).  One CTA
// keeps v, f, r of all levels 0..top (<= 64^2: 140 KB) in shared memory and runs the whole
// sub-V-cycle there: smooth / residual / restrict on the way down, the bottom solve, prolong +
// correct + smooth on the way up -- same device functions, same operation order, same bits.
constexpr int MG_COARSE_TOP_N = 64;
constexpr int MG_COARSE_LEVELS = 6;     // 2, 4, 8, 16, 32, 64

struct CoarseTable {
    MgLevel g[MG_COARSE_LEVELS];        // the levels' global planes
    int top;                            // highest level handled here
    int nsmooth, nsmooth_bottom;
    ResidCoef rcoef[MG_COARSE_LEVELS];
    SmoothCoef coef[MG_COARSE_LEVELS];
    MgBC bc_top, bc_coarse;             // bc_top carries the inhomogeneous values when top is the finest
    VcEdges edges[MG_COARSE_LEVELS];    // variable coefficients: the levels' edge planes (global memory,
                                        // read-only and L1-resident; indexed with the GLOBAL pitch g[l].pitch)
};

// Shared-memory geometry of coarse level L, known at compile time: n = 2^(L+1) cells per side, q = n + 2 rows of q
// doubles, planes v, f, r back to back, levels packed from the coarsest up.  (The first version kept an MgLevel array
// indexed with the run-time level in LOCAL memory and ran every level, the 2 x 2 bottom solve included, on 1024
// threads with a CTA barrier per colour: 280 us for the sub-cycle below 128^2, 70 us of it the bottom solve alone.)
#ifndef MG_COARSE_THREADS
#define MG_COARSE_THREADS 1024
#endif

template <int L> struct CoarseGeom {
    static constexpr int n = 2 << L, q = n + 2, plane = q * q;
    static constexpr int base = CoarseGeom<L - 1>::base + 3 * CoarseGeom<L - 1>::plane;
};
template <> struct CoarseGeom<0> { static constexpr int n = 2, q = 4, plane = 16, base = 0; };
// A level is run by as many warps as it has points of one colour (one point per thread, two at 64^2), the others skip
// it: a warp that runs alone issues one instruction every 5-8 cycles, so the cost of a colour is the length of ONE point's
// instruction chain plus the barrier -- both kept short: the thread's two points (one per colour), their addresses,
// right-hand sides and boundary flags are fixed for the whole smooth(), and the barrier spans only the level's warps
// (bar.sync with a thread count; __syncwarp when one warp suffices).
template <int L> struct CoarseWarps {
    static constexpr int pts = CoarseGeom<L>::n * CoarseGeom<L>::n / 2;
    static constexpr int warps = pts <= 32 ? 1 : (pts / 32 > MG_COARSE_THREADS / 32 ? MG_COARSE_THREADS / 32 : pts / 32);
};

__device__ __forceinline__ void named_barrier(int id, int nthreads)
{
#ifdef P2B_EMU_HEADER
    emu_bar_sync(id, nthreads);
#else
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
#endif
}

template <int L> __device__ __forceinline__ void coarse_sync()
{
    if (CoarseWarps<L>::warps == 1) __syncwarp();
    else named_barrier(1 + L, 32 * CoarseWarps<L>::warps);
}

template <int L> __device__ __forceinline__ MgLevel coarse_level_view(const CoarseTable& T, double* sm)
{
    using G = CoarseGeom<L>;
    MgLevel S = T.g[L];
    S.pitch = G::q;
    S.v = sm + G::base; S.f = S.v + G::plane; S.r = S.f + G::plane;
    return S;
}

// A point of a level and the ghost cells it generates, worked out once per smooth(): store_with_ghosts() is a dozen
// data-dependent branches, and a warp running on its own pays 10-20 cycles for each of them (measured: 1300 cycles per
// colour of the 2 x 2 bottom solve, all in that function).  Without inhomogeneous boundary values a ghost is +-(the
// cell's value), so a point has at most an x ghost, a y ghost and their corner, each a fixed offset and sign:
//   k: offset of the cell; ox / oy / oc: offsets of the x ghost, y ghost, corner (-1: none);
//   mx / my: XOR masks for the sign word of the x / y ghost (reflect-odd / homogeneous dirichlet negate)
struct CoarsePt { int k, ox, oy, oc; unsigned mx, my; };

__device__ __forceinline__ CoarsePt coarse_point(int n, int P, int i, int j, const MgBC& b)
{
    CoarsePt p;
    p.k = i * P + j;
    p.ox = p.oy = p.oc = -1;
    p.mx = p.my = 0u;
    const int sxl = (b.xl == P2B_BC_PERIODIC) ? n : 1, sxh = (b.xr == P2B_BC_PERIODIC) ? 1 : n;
    const int syl = (b.yl == P2B_BC_PERIODIC) ? n : 1, syh = (b.yr == P2B_BC_PERIODIC) ? 1 : n;
    int gi = -1, gj = -1;                                 // ghost row / column this cell feeds
    if (i == sxl) { gi = 0; p.mx = (b.xl == P2B_BC_REFLECT_ODD) ? 0x80000000u : 0u; }
    else if (i == sxh) { gi = n + 1; p.mx = (b.xr == P2B_BC_REFLECT_ODD) ? 0x80000000u : 0u; }
    if (j == syl) { gj = 0; p.my = (b.yl == P2B_BC_REFLECT_ODD) ? 0x80000000u : 0u; }
    else if (j == syh) { gj = n + 1; p.my = (b.yr == P2B_BC_REFLECT_ODD) ? 0x80000000u : 0u; }
    if (gi >= 0) p.ox = gi * P + j;
    if (gj >= 0) p.oy = i * P + gj;
    if (gi >= 0 && gj >= 0) p.oc = gi * P + gj;
    return p;
}

__device__ __forceinline__ double flip_sign(double x, unsigned mask)
{
#ifdef P2B_EMU_HEADER
    unsigned long long bits;
    memcpy(&bits, &x, 8);
    bits ^= (unsigned long long)mask << 32;
    memcpy(&x, &bits, 8);
    return x;
#else
    return __hiloint2double(__double2hiint(x) ^ (int)mask, __double2loint(x));
#endif
}

__device__ __forceinline__ void coarse_store(double* v, const CoarsePt& p, double val)
{
    v[p.k] = val;
    if (p.ox >= 0) v[p.ox] = flip_sign(val, p.mx);
    if (p.oy >= 0) {
        v[p.oy] = flip_sign(val, p.my);
        if (p.oc >= 0) v[p.oc] = flip_sign(val, p.mx ^ p.my);
    }
}

// smooth() of one level in shared memory (MG.py:544-599): fill_BC, then nsmooth red-black iterations with the ghost
// update fused into the writers.  t: this thread's index among the nt = 32 * CoarseWarps<L>::warps threads of the level.
template <bool VC, int L>
__device__ __forceinline__ void coarse_smooth(const MgLevel& S, const MgBC& b, const SmoothCoef& c, int nsmooth,
                                              const VcEdges& E, int gpitch, int t)
{
    constexpr int n = CoarseGeom<L>::n, half = n / 2, npts = n * half, P = CoarseGeom<L>::q;
    constexpr int nt = 32 * CoarseWarps<L>::warps, PER = (npts + nt - 1) / nt;      // points per thread and colour: 1 or 2
    // inhomogeneous boundary values (only when this level is the finest of the hierarchy): the general writer
    const bool inhom = b.xlv || b.xrv || b.ylv || b.yrv;
    for (int u = t; u < 4 * n; u += nt) {
        const int side = u / n, q = u % n + 1;
        const int i = side == 0 ? 1 : side == 1 ? n : q;
        const int j = side == 2 ? 1 : side == 3 ? n : q;
        if (inhom) store_with_ghosts(S.v, n, n, P, i, j, S.v[i * P + j], b, S.dx, S.dy);
        else coarse_store(S.v, coarse_point(n, P, i, j, b), S.v[i * P + j]);
    }
    // this thread's points: index u = t + p * nt -> row i, the k-th point of the colour in that row
    CoarsePt pt[PER][2];
    double ff[PER][2];
    bool act[PER];
#pragma unroll
    for (int p = 0; p < PER; ++p) {
        const int u = t + p * nt;
        act[p] = u < npts;
        const int i = (act[p] ? u : 0) / half + 1, k = (act[p] ? u : 0) % half;
#pragma unroll
        for (int col = 0; col < 2; ++col) {
            const int j = 1 + ((i + 1 + col) & 1) + 2 * k;
            pt[p][col] = coarse_point(n, P, i, j, b);
            ff[p][col] = S.f[pt[p][col].k];
        }
    }
    coarse_sync<L>();
    for (int it = 0; it < nsmooth; ++it) {
#pragma unroll
        for (int col = 0; col < 2; ++col) {
#pragma unroll
            for (int p = 0; p < PER; ++p) {
                if (!act[p]) continue;
                const int k = pt[p][col].k;
                const double* v = S.v;
                double val;
                if (VC) {
                    const int i = k / P, j = k % P;
                    const long long kg = (long long)i * gpitch + j;
                    val = vc_gs_value(ff[p][col], v[k - P], v[k + P], v[k - 1], v[k + 1],
                                      E.ex[kg], E.ex[kg + gpitch], E.ey[kg], E.ey[kg + 1]);
                } else {
                    // MG.py:593-596, the operations of gs_update()
                    const double sx = exact_add(v[k + P], v[k - P]);
                    const double sy = exact_add(v[k + 1], v[k - 1]);
                    const double num = exact_add(exact_add(ff[p][col], exact_mul(c.xc, sx)), exact_mul(c.yc, sy));
                    val = div_by_denom(num, c);
                }
                if (inhom) store_with_ghosts(S.v, n, n, P, k / P, k % P, val, b, S.dx, S.dy);
                else coarse_store(S.v, pt[p][col], val);
            }
            coarse_sync<L>();
        }
    }
}

// the sub-V-cycle from level L down and back (MG.py:699-778) on shared-memory planes; entered by the first
// CoarseWarps<L>::warps warps of the CTA only
template <bool VC, int L>
__device__ void coarse_cycle(const CoarseTable& T, double* sm, int tid)
{
    const MgBC& b = (L == T.top) ? T.bc_top : T.bc_coarse;
    MgLevel S = coarse_level_view<L>(T, sm);
    if (L == 0) {
        coarse_smooth<VC, L>(S, b, T.coef[L], T.nsmooth_bottom, T.edges[L], T.g[L].pitch, tid);
        return;
    }
    constexpr int LC = L > 0 ? L - 1 : 0;
    constexpr int n = CoarseGeom<L>::n, P = CoarseGeom<L>::q, nc = CoarseGeom<LC>::n, Pc = CoarseGeom<LC>::q;
    constexpr int nt = 32 * CoarseWarps<L>::warps;
    MgLevel C = coarse_level_view<LC>(T, sm);
    coarse_smooth<VC, L>(S, b, T.coef[L], T.nsmooth, T.edges[L], T.g[L].pitch, tid);
    for (int t = tid; t < n * n; t += nt) {
        const int i = t / n + 1, j = t % n + 1, k = i * P + j;
        if (VC) {
            const int Pg = T.g[L].pitch;
            const long long kg = (long long)i * Pg + j;
            const double* v = S.v;
            S.r[k] = vc_residual_value(S.f[k], v[k], v[k - P], v[k + P], v[k - 1], v[k + 1],
                                       T.edges[L].ex[kg], T.edges[L].ex[kg + Pg], T.edges[L].ey[kg], T.edges[L].ey[kg + 1]);
        } else {
            S.r[k] = residual_at(S, k, T.rcoef[L]);
        }
    }
    coarse_sync<L>();
    for (int t = tid; t < nc * nc; t += nt) {
        const int ic = t / nc + 1, jc = t % nc + 1;
        const double* r = S.r;
        const int k = (2 * ic - 1) * P + (2 * jc - 1);
        double sum = exact_add(exact_add(exact_add(r[k], r[k + P]), r[k + 1]), r[k + P + 1]);
        C.f[ic * Pc + jc] = exact_mul(0.25, sum);
    }
    coarse_sync<L>();
    if (tid < 32 * CoarseWarps<LC>::warps) coarse_cycle<VC, LC>(T, sm, tid);
    coarse_sync<L>();
    for (int t = tid; t < nc * nc; t += nt) {
        const int ic = t / nc + 1, jc = t % nc + 1;
        const double* c = C.v;
        const int kc = ic * Pc + jc;
        double mx = exact_mul(0.5, exact_sub(c[kc + Pc], c[kc - Pc]));
        double my = exact_mul(0.5, exact_sub(c[kc + 1], c[kc - 1]));
        double qx = exact_mul(0.25, mx), qy = exact_mul(0.25, my), c0 = c[kc];
        const int i = 2 * ic - 1, j = 2 * jc - 1;
        double* v = S.v;
        const double n00 = exact_add(v[i * P + j], exact_sub(exact_sub(c0, qx), qy));
        const double n10 = exact_add(v[(i + 1) * P + j], exact_sub(exact_add(c0, qx), qy));
        const double n01 = exact_add(v[i * P + j + 1], exact_add(exact_sub(c0, qx), qy));
        const double n11 = exact_add(v[(i + 1) * P + j + 1], exact_add(exact_add(c0, qx), qy));
        if (b.xlv || b.xrv || b.ylv || b.yrv) {
            store_with_ghosts(v, n, n, P, i, j, n00, b, S.dx, S.dy);
            store_with_ghosts(v, n, n, P, i + 1, j, n10, b, S.dx, S.dy);
            store_with_ghosts(v, n, n, P, i, j + 1, n01, b, S.dx, S.dy);
            store_with_ghosts(v, n, n, P, i + 1, j + 1, n11, b, S.dx, S.dy);
        } else if (ic == 1 || ic == nc || jc == 1 || jc == nc) {
            coarse_store(v, coarse_point(n, P, i, j, b), n00);
            coarse_store(v, coarse_point(n, P, i + 1, j, b), n10);
            coarse_store(v, coarse_point(n, P, i, j + 1, b), n01);
            coarse_store(v, coarse_point(n, P, i + 1, j + 1, b), n11);
        } else {
            v[i * P + j] = n00; v[(i + 1) * P + j] = n10; v[i * P + j + 1] = n01; v[(i + 1) * P + j + 1] = n11;
        }
    }
    coarse_sync<L>();
    coarse_smooth<VC, L>(S, b, T.coef[L], T.nsmooth, T.edges[L], T.g[L].pitch, tid);
}

// global <-> shared copy of one level's planes (all threads): load v and f of the top level (v is the current iterate
// there) and zero below it (MG.py:658-659); store v, f, r of every level back (coarse planes stay observable through
// grids[level], like the reference's)
template <int L>
__device__ __forceinline__ void coarse_io(const CoarseTable& T, double* sm, bool load, int tid, int nthreads)
{
    using G = CoarseGeom<L>;
    if (L > T.top) return;
    double* v = sm + G::base;
    double* f = v + G::plane;
    double* r = f + G::plane;
    const MgLevel& g = T.g[L];
    for (int t = tid; t < G::plane; t += nthreads) {
        const int i = t / G::q, j = t % G::q;
        const long long kg = (long long)i * g.pitch + j;
        if (load) {
            v[t] = (L == T.top) ? g.v[kg] : 0.0;
            f[t] = (L == T.top) ? g.f[kg] : 0.0;
            r[t] = g.r[kg];
        } else {
            g.v[kg] = v[t]; g.f[kg] = f[t]; g.r[kg] = r[t];
        }
    }
}

template <bool VC>
__global__ void __launch_bounds__(MG_COARSE_THREADS, 1) mg_coarse_vcycle_kernel(const __grid_constant__ CoarseTable T)
{
    P2B_DYN_SMEM(double, sm);
    if (T.g[0].ctl && T.g[0].ctl[CW_STOP]) return;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    coarse_io<0>(T, sm, true, tid, nthreads); coarse_io<1>(T, sm, true, tid, nthreads); coarse_io<2>(T, sm, true, tid, nthreads);
    coarse_io<3>(T, sm, true, tid, nthreads); coarse_io<4>(T, sm, true, tid, nthreads); coarse_io<5>(T, sm, true, tid, nthreads);
    __syncthreads();
    switch (T.top) {
        case 5: if (tid < 32 * CoarseWarps<5>::warps) coarse_cycle<VC, 5>(T, sm, tid); break;
        case 4: if (tid < 32 * CoarseWarps<4>::warps) coarse_cycle<VC, 4>(T, sm, tid); break;
        case 3: if (tid < 32 * CoarseWarps<3>::warps) coarse_cycle<VC, 3>(T, sm, tid); break;
        case 2: if (tid < 32 * CoarseWarps<2>::warps) coarse_cycle<VC, 2>(T, sm, tid); break;
        case 1: if (tid < 32 * CoarseWarps<1>::warps) coarse_cycle<VC, 1>(T, sm, tid); break;
        default: if (tid < 32 * CoarseWarps<0>::warps) coarse_cycle<VC, 0>(T, sm, tid); break;
    }
    __syncthreads();
    coarse_io<0>(T, sm, false, tid, nthreads); coarse_io<1>(T, sm, false, tid, nthreads); coarse_io<2>(T, sm, false, tid, nthreads);
    coarse_io<3>(T, sm, false, tid, nthreads); coarse_io<4>(T, sm, false, tid, nthreads); coarse_io<5>(T, sm, false, tid, nthreads);
}

// full ghost fill of v from the interior (used once per smooth() like MG.py:565)
__global__ void mg_fill_kernel(MgLevel L, MgBC b)
{
    if (L.ctl && L.ctl[CW_STOP]) return;
    // every interior edge cell re-stores itself with its ghosts
    const int n = L.n;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < 4 * n; t += gridDim.x * blockDim.x) {
        int side = t / n, s = t % n + 1;
        int i, j;
        if (side == 0) { i = 1; j = s; } else if (side == 1) { i = n; j = s; }
        else if (side == 2) { i = s; j = 1; } else { i = s; j = n; }
        // corners are visited twice with identical results
        store_with_ghosts(L.v, n, n, L.pitch, i, j, L.v[(long long)i * L.pitch + j], b, L.dx, L.dy);
    }
}

__global__ void mg_residual_kernel(MgLevel L, ResidCoef rc, MgComm cm)
{
    if (L.ctl && L.ctl[CW_STOP]) return;
    const int j = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int i = blockIdx.y * blockDim.y + threadIdx.y + 1;
    // slab: rows 0 and ni + 1 are the neighbours' rows, pushed by their last smoothing pass
    comm_block_wait(cm, rows_meet(blockIdx.y, blockDim.y, 1, 1), rows_meet(blockIdx.y, blockDim.y, L.ni, L.ni));
    if (i > L.ni || j > L.n) return;
    const long long k = (long long)i * L.pitch + j;
    L.r[k] = residual_at(L, k, rc);
}

// fine r -> coarse f, valid region (patch.py:659-662, MG.py:731-732).  crow: row offset of this
// rank's rows inside the coarse array (non-zero when a slab level restricts into a replicated one)
__global__ void mg_restrict_kernel(MgLevel F, MgLevel Cs, int crow, MgComm cm, int to_all)
{
    if (F.ctl && F.ctl[CW_STOP]) return;
    const int jc = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int ic = blockIdx.y * blockDim.y + threadIdx.y + 1;
    const int nic = F.ni / 2;
    if (ic <= nic && jc <= Cs.n) {
        const long long k = (long long)(2 * ic - 1) * F.pitch + (2 * jc - 1);
        const double* r = F.r;
        double s = exact_add(exact_add(exact_add(r[k], r[k + F.pitch]), r[k + 1]), r[k + F.pitch + 1]);
        const double val = exact_mul(0.25, s);
        const long long kc = (long long)(ic + crow) * Cs.pitch + jc;
        Cs.f[kc] = val;
        if (cm.ctl && cm.sig_ord >= 0) {
            if (to_all) {
                // slab -> replicated: every rank needs the whole coarse right-hand side
                for (int p = 0; p < cm.size; ++p)
                    if (p != cm.rank) (Cs.f + (long long)cm.ctl[CW_PEER + p])[kc] = val;
            } else {
                // slab -> slab: the blocked smoother updates halo cells redundantly and needs their right-hand side
                if (cm.has_lo && ic <= Cs.gx) (Cs.f + cm.dlo)[(long long)(nic + ic) * Cs.pitch + jc] = val;
                if (cm.has_hi && ic > nic - Cs.gx) (Cs.f + cm.dhi)[(long long)(ic - nic) * Cs.pitch + jc] = val;
            }
        }
    }
    if (to_all) comm_block_signal_all(cm, gridDim.x * gridDim.y);
    else comm_block_signal(cm, rows_meet(blockIdx.y, blockDim.y, 1, Cs.gx), rows_meet(blockIdx.y, blockDim.y, nic - Cs.gx + 1, nic));
}

// v_fine += prolong(v_coarse), ghosts refreshed (patch.py:716-734, MG.py:745-751)
__global__ void mg_prolong_kernel(MgLevel F, MgLevel Cs, MgBC b, int crow, MgComm cm, int coarse_is_slab)
{
    if (F.ctl && F.ctl[CW_STOP]) return;
    const int jc = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int ic = blockIdx.y * blockDim.y + threadIdx.y + 1;
    const int nic = F.ni / 2, hp = F.gx / 2;      // coarse rows whose fine rows are the neighbours' halo rows: hp per side
    // coarse rows 0 and nic + 1 of a slab are the neighbours' rows, pushed by their last smoothing pass
    if (coarse_is_slab) comm_block_wait(cm, rows_meet(blockIdx.y, blockDim.y, 1, 1), rows_meet(blockIdx.y, blockDim.y, nic, nic));
    if (ic <= nic && jc <= Cs.n) {
        const double* c = Cs.v;
        const long long kc = (long long)(ic + crow) * Cs.pitch + jc;
        double mx = exact_mul(0.5, exact_sub(c[kc + Cs.pitch], c[kc - Cs.pitch]));
        double my = exact_mul(0.5, exact_sub(c[kc + 1], c[kc - 1]));
        double qx = exact_mul(0.25, mx), qy = exact_mul(0.25, my);
        double c0 = c[kc];
        const int i = 2 * ic - 1, j = 2 * jc - 1;
        double* v = F.v;
        const int P = F.pitch, ni = F.ni;
        const double n00 = exact_add(v[(long long)i * P + j], exact_sub(exact_sub(c0, qx), qy));
        const double n10 = exact_add(v[(long long)(i + 1) * P + j], exact_sub(exact_add(c0, qx), qy));
        const double n01 = exact_add(v[(long long)i * P + j + 1], exact_add(exact_sub(c0, qx), qy));
        const double n11 = exact_add(v[(long long)(i + 1) * P + j + 1], exact_add(exact_add(c0, qx), qy));
        store_with_ghosts(v, ni, F.n, P, i, j, n00, b, F.dx, F.dy, F.ioff);
        store_with_ghosts(v, ni, F.n, P, i + 1, j, n10, b, F.dx, F.dy, F.ioff);
        store_with_ghosts(v, ni, F.n, P, i, j + 1, n01, b, F.dx, F.dy, F.ioff);
        store_with_ghosts(v, ni, F.n, P, i + 1, j + 1, n11, b, F.dx, F.dy, F.ioff);
        if (cm.ctl && cm.sig_ord >= 0) {
            // slab: the corrected first / last gx rows are the neighbours' halo rows for the smoothing that follows
            if (cm.has_lo && ic <= hp) {
                double* d = v + cm.dlo;
                d[(long long)(ni + i) * P + j] = n00; d[(long long)(ni + i + 1) * P + j] = n10;
                d[(long long)(ni + i) * P + j + 1] = n01; d[(long long)(ni + i + 1) * P + j + 1] = n11;
            }
            if (cm.has_hi && ic > nic - hp) {
                double* d = v + cm.dhi;
                d[(long long)(i - ni) * P + j] = n00; d[(long long)(i + 1 - ni) * P + j] = n10;
                d[(long long)(i - ni) * P + j + 1] = n01; d[(long long)(i + 1 - ni) * P + j + 1] = n11;
            }
        }
    }
    comm_block_signal(cm, rows_meet(blockIdx.y, blockDim.y, 1, hp), rows_meet(blockIdx.y, blockDim.y, nic - hp + 1, nic));
}

// ---- deterministic reductions over the valid region ------------------------------------------------
// One CTA per row (>= one CTA per MG_NPART-th of the rows), a thread owns up to RED_PER_THREAD cells
// of the row and issues all of its loads before any arithmetic (memory-level parallelism: the first
// version looped cell by cell and ran at 23% of HBM bandwidth, ncu r1).  Fixed summation order:
// per-thread sequential, block tree, then one CTA sums the per-row partials in index order.
constexpr int RED_THREADS = 256;
constexpr int RED_PER_THREAD = 4;           // cells a thread keeps in flight per trip

__device__ __forceinline__ double block_sum(double s, double* sh)
{
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = RED_THREADS / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    double r = sh[0];
    __syncthreads();
    return r;
}

// mode 0: sum a^2
__global__ void __launch_bounds__(RED_THREADS) mg_sumsq_partial_kernel(const double* __restrict__ a, int ni, int n, int pitch,
                                                                       double* __restrict__ part)
{
    __shared__ double sh[RED_THREADS];
    double s = 0.0;
    for (int i = 1 + blockIdx.x; i <= ni; i += gridDim.x) {
        const double* row = a + (long long)i * pitch;
        for (int j0 = 1 + threadIdx.x; j0 <= n; j0 += RED_THREADS * RED_PER_THREAD) {
            double x[RED_PER_THREAD];
#pragma unroll
            for (int u = 0; u < RED_PER_THREAD; ++u) {
                int j = j0 + u * RED_THREADS;
                x[u] = (j <= n) ? row[j] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < RED_PER_THREAD; ++u) s += x[u] * x[u];
        }
    }
    s = block_sum(s, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
}

// thread 0 of a single-CTA kernel: sum (a, b) over all ranks in rank order (every rank gets the same bits).  Each rank
// writes its pair into its slot on every rank, publishes, waits for everybody's slot.  A program of its own
// (new_program) or a step of the running one.
__device__ __forceinline__ void comm_allreduce2(const MgComm& c, double& a, double& b, bool new_program)
{
    if (!c.ctl) return;
    if (new_program) c.ctl[CW_EPOCH] += 1;
    const unsigned long long ep = c.ctl[CW_EPOCH], val = ep * MG_ORD_STRIDE + (unsigned long long)(c.sig_ord < 0 ? 0 : c.sig_ord);
    double* slots = reinterpret_cast<double*>(c.ctl + CW_SUMS) + (ep & 1ull) * MG_MAX_RANKS * 4;
    for (int r = 0; r < c.size; ++r) {
        double* dst = slots + (long long)c.ctl[CW_PEER + r] + c.rank * 4;
        dst[0] = a; dst[1] = b;
    }
    __threadfence_system();        // one drain of the write queue, then the flags
    for (int r = 0; r < c.size; ++r) comm_store_relaxed(c.ctl + (long long)c.ctl[CW_PEER + r] + CW_GFLAG + c.rank, val);
    for (int r = 0; r < c.size; ++r) comm_wait_ge(c.ctl, CW_GFLAG + r, val);
    a = 0.0; b = 0.0;
    for (int r = 0; r < c.size; ++r) {
        const volatile double* src = slots + r * 4;
        a += src[0]; b += src[1];
    }
}

__global__ void mg_sumsq_final_kernel(const double* part, int npart, double* out, MgComm cm)
{
    __shared__ double sh[RED_THREADS];
    double s = 0.0;
    for (int t = threadIdx.x; t < npart; t += RED_THREADS) s += part[t];
    s = block_sum(s, sh);
    if (threadIdx.x == 0) {
        double z = 0.0;
        comm_allreduce2(cm, s, z, true);
        *out = s;
    }
}

struct MgZeroTable { double* v[MG_MAX_LEVELS]; long long count[MG_MAX_LEVELS]; int nlev; const unsigned long long* ctl; };

__global__ void mg_zero_kernel(MgZeroTable t)
{
    if (t.ctl && t.ctl[CW_STOP]) return;
    for (int l = 0; l < t.nlev; ++l)
        for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < t.count[l];
             k += (long long)gridDim.x * blockDim.x)
            t.v[l][k] = 0.0;
}

// solve()'s per-cycle bookkeeping in one pass over the finest level (MG.py:668-686): relative change
// against old_phi (old_phi <- v), residual r (stored), partial sums of both squares.  Same fixed
// two-stage summation as mg_sumsq_*; part[0..nb) relative change, part[MG_NPART..) residual.
__global__ void __launch_bounds__(RED_THREADS)
mg_diag_partial_kernel(MgLevel L, double* __restrict__ old_phi, ResidCoef rc, double* __restrict__ part)
{
    __shared__ double sh[RED_THREADS];
    if (L.ctl && L.ctl[CW_STOP]) return;
    double s_rel = 0.0, s_res = 0.0;
    const int n = L.n, P = L.pitch;
    const double* __restrict__ v = L.v;
    const double* __restrict__ f = L.f;
    double* __restrict__ r = L.r;
    for (int i = 1 + blockIdx.x; i <= L.ni; i += gridDim.x) {
        const long long base = (long long)i * P;
        for (int j0 = 1 + threadIdx.x; j0 <= n; j0 += RED_THREADS * RED_PER_THREAD) {
            double c[RED_PER_THREAD], up[RED_PER_THREAD], dn[RED_PER_THREAD], lf[RED_PER_THREAD],
                rt[RED_PER_THREAD], ff[RED_PER_THREAD], oo[RED_PER_THREAD];
#pragma unroll
            for (int u = 0; u < RED_PER_THREAD; ++u) {
                int j = j0 + u * RED_THREADS;
                bool ok = j <= n;
                long long k = base + (ok ? j : 1);
                c[u] = v[k]; up[u] = v[k - P]; dn[u] = v[k + P]; lf[u] = v[k - 1]; rt[u] = v[k + 1];
                ff[u] = f[k]; oo[u] = old_phi[k];
            }
#pragma unroll
            for (int u = 0; u < RED_PER_THREAD; ++u) {
                int j = j0 + u * RED_THREADS;
                if (j > n) continue;
                long long k = base + j;
                old_phi[k] = c[u];
                double d = (c[u] - oo[u]) / (c[u] + 1.e-16);
                s_rel += d * d;
                // MG.py:540-542
                double v2 = exact_mul(2.0, c[u]);
                double lx = div_const(exact_sub(exact_add(up[u], dn[u]), v2), rc.dx2);
                double ly = div_const(exact_sub(exact_add(lf[u], rt[u]), v2), rc.dy2);
                double res = exact_add(exact_sub(ff[u], exact_mul(rc.alpha, c[u])), exact_mul(rc.beta, exact_add(lx, ly)));
                r[k] = res;
                s_res += res * res;
            }
        }
    }
    s_rel = block_sum(s_rel, sh);
    s_res = block_sum(s_res, sh);
    if (threadIdx.x == 0) { part[blockIdx.x] = s_rel; part[MG_NPART + blockIdx.x] = s_res; }
}

// second stage of the per-cycle bookkeeping: sums over the rows (and over the ranks of a decomposed hierarchy), and the
// stopping rule of solve() evaluated on the device (MG.py:654-697: while residual_error > rtol and cycle <= max_cycles):
// scale = dx * dy, source_norm and rtol as the host would use them -- same IEEE operations, same decision.
// When the rule says stop, CW_STOP turns every kernel of a cycle that was enqueued ahead into a no-op, so the host can
// enqueue cycles without waiting for each cycle's two scalars.  results: (relsq, rsq, residual_error, cycles run).
__global__ void mg_set_stop_kernel(unsigned long long* ctl, double source_norm, double rtol, double max_cycles, double enabled)
{
    double* par = reinterpret_cast<double*>(ctl + CW_STOPPAR);
    par[0] = source_norm; par[1] = rtol; par[2] = max_cycles; par[3] = enabled;
    double* res = reinterpret_cast<double*>(ctl + CW_RESULT);
    res[0] = res[1] = res[2] = res[3] = 0.0;
    ctl[CW_STOP] = 0ull;
}

__global__ void mg_diag_final_kernel(const double* part, int npart, double* out, MgComm cm, unsigned long long* ctl, double scale)
{
    __shared__ double sh[RED_THREADS];
    if (ctl && ctl[CW_STOP]) return;
    double a = 0.0, b = 0.0;
    for (int t = threadIdx.x; t < npart; t += RED_THREADS) { a += part[t]; b += part[MG_NPART + t]; }
    a = block_sum(a, sh);
    b = block_sum(b, sh);
    if (threadIdx.x == 0) {
        comm_allreduce2(cm, a, b, true);
        out[0] = a; out[1] = b;
        const double* par = reinterpret_cast<const double*>(ctl + CW_STOPPAR);
        if (ctl && par[3] != 0.0) {
            double* res = reinterpret_cast<double*>(ctl + CW_RESULT);
            const double rnorm = exact_sqrt(exact_mul(scale, b));
            const double err = par[0] != 0.0 ? exact_div(rnorm, par[0]) : rnorm;
            const double ncyc = res[3] + 1.0;
            res[0] = a; res[1] = b; res[2] = err; res[3] = ncyc;
            if (!(err > par[1]) || ncyc >= par[2]) ctl[CW_STOP] = 1ull;
        }
    }
}

// ---- small communication kernels -------------------------------------------------------------------------
// starts a program on this rank
__global__ void mg_epoch_kernel(unsigned long long* ctl)
{
    if (ctl[CW_STOP]) return;
    ctl[CW_EPOCH] += 1;
}

// <<<1, 32>>>: wait until both neighbours' words (gather = 0) or every rank's all-rank word (gather = 1) have reached
// the value of launch cm.wait_ord of the running program
__global__ void mg_comm_wait_kernel(MgComm cm, int gather)
{
    if (cm.ctl[CW_STOP]) return;
    const unsigned long long target = comm_value(cm, cm.wait_ord);
    const int t = threadIdx.x;
    if (gather) { if (t < cm.size) comm_wait_ge(cm.ctl, CW_GFLAG + t, target); }
    else {
        if (t == 0 && cm.has_lo) comm_wait_ge(cm.ctl, CW_FROM_LO, target);
        if (t == 1 && cm.has_hi) comm_wait_ge(cm.ctl, CW_FROM_HI, target);
    }
}

// stand-alone halo exchange, phase 1 (<<<1, 1>>>): start a program, tell both neighbours "everything I enqueued
// before this exchange has finished -- my halo rows may be overwritten", wait for the same from them
__global__ void mg_xchg_arrive_kernel(MgComm cm)
{
    cm.ctl[CW_EPOCH] += 1;
    const unsigned long long val = comm_value(cm, 0);
    __threadfence_system();
    if (cm.has_lo) comm_store(cm.ctl + cm.dlo + CW_FROM_HI, val);
    if (cm.has_hi) comm_store(cm.ctl + cm.dhi + CW_FROM_LO, val);
    if (cm.has_lo) comm_wait_ge(cm.ctl, CW_FROM_LO, val);
    if (cm.has_hi) comm_wait_ge(cm.ctl, CW_FROM_HI, val);
}

// phase 2: copy my first / last `depth` owned rows (whole rows, ghost columns included) into the neighbours' halo rows
__global__ void mg_xchg_push_kernel(double* plane, int ni, int pitch, int depth, MgComm cm)
{
    const int r = blockIdx.y + 1, col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col < pitch) {
        if (cm.has_lo) (plane + cm.dlo)[(long long)(ni + r) * pitch + col] = plane[(long long)r * pitch + col];
        if (cm.has_hi) (plane + cm.dhi)[(long long)(r - depth) * pitch + col] = plane[(long long)(ni - depth + r) * pitch + col];
    }
    comm_block_signal(cm, true, true);
}

// ---- variable coefficients: kernels (the per-point arithmetic is defined ahead of the blocked smoother)
__device__ __forceinline__ double vc_gs_update(const double* v, const double* f, const VcEdges& E, int pitch, int i, int j)
{
    const long long k = (long long)i * pitch + j;
    return vc_gs_value(f[k], v[k - pitch], v[k + pitch], v[k - 1], v[k + 1], E.ex[k], E.ex[k + pitch], E.ey[k], E.ey[k + 1]);
}

__device__ __forceinline__ double vc_residual_at(const MgLevel& L, const VcEdges& E, long long k)
{
    const double* v = L.v;
    return vc_residual_value(L.f[k], v[k], v[k - L.pitch], v[k + L.pitch], v[k - 1], v[k + 1],
                             E.ex[k], E.ex[k + L.pitch], E.ey[k], E.ey[k + 1]);
}

__global__ void mg_vc_halfsweep_kernel(MgLevel L, MgBC b, VcEdges E, int colour)
{
    if (L.ctl && L.ctl[CW_STOP]) return;
    const int half = L.n >> 1;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y + 1;
    if (k >= half || i > L.n) return;
    const int j = 1 + ((i + 1 + colour) & 1) + 2 * k;
    double val = vc_gs_update(L.v, L.f, E, L.pitch, i, j);
    store_with_ghosts(L.v, L.n, L.n, L.pitch, i, j, val, b, L.dx, L.dy);
}

__global__ void mg_vc_smooth_small_kernel(MgLevel L, MgBC b, VcEdges E, int nsmooth)
{
    if (L.ctl && L.ctl[CW_STOP]) return;
    const int half = L.n >> 1;
    const int npts = L.n * half;
    for (int it = 0; it < 2 * nsmooth; ++it) {
        const int colour = it & 1;
        for (int t = threadIdx.x; t < npts; t += blockDim.x) {
            int i = t / half + 1, k = t % half;
            int j = 1 + ((i + 1 + colour) & 1) + 2 * k;
            double val = vc_gs_update(L.v, L.f, E, L.pitch, i, j);
            store_with_ghosts(L.v, L.n, L.n, L.pitch, i, j, val, b, L.dx, L.dy);
        }
        __syncthreads();
    }
}

__global__ void mg_vc_residual_kernel(MgLevel L, VcEdges E)
{
    if (L.ctl && L.ctl[CW_STOP]) return;
    const int j = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int i = blockIdx.y * blockDim.y + threadIdx.y + 1;
    if (i > L.ni || j > L.n) return;
    const long long k = (long long)i * L.pitch + j;
    L.r[k] = vc_residual_at(L, E, k);
}

// EdgeCoeffs.__init__ (edge_coeffs.py:10-29) on the finest level, from the ghost-filled cell-centred eta
__global__ void mg_vc_edges_fine_kernel(const double* __restrict__ c, double* __restrict__ ex, double* __restrict__ ey,
                                        int n, int pitch, DivConst dx2, DivConst dy2)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int i = blockIdx.y * blockDim.y + threadIdx.y + 1;
    if (i > n + 1 || j > n + 1) return;
    const long long k = (long long)i * pitch + j;
    ex[k] = div_const(exact_mul(0.5, exact_add(c[k - pitch], c[k])), dx2);
    ey[k] = div_const(exact_mul(0.5, exact_add(c[k - 1], c[k])), dy2);
}

// EdgeCoeffs.restrict (edge_coeffs.py:31-54): average the two fine edges that make up a coarse edge,
// then "redo the normalization": * dx_fine^2 / dx_coarse^2
__global__ void mg_vc_edges_restrict_kernel(const double* __restrict__ xf, const double* __restrict__ yf, int pf,
                                            double* __restrict__ ex, double* __restrict__ ey, int n, int pitch,
                                            double fdx2, DivConst cdx2, double fdy2, DivConst cdy2)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int i = blockIdx.y * blockDim.y + threadIdx.y + 1;
    if (i > n + 1 || j > n + 1) return;
    const long long k = (long long)i * pitch + j;
    const long long kf = (long long)(2 * i - 1) * pf + (2 * j - 1);
    if (j <= n) ex[k] = div_const(exact_mul(exact_mul(0.5, exact_add(xf[kf], xf[kf + 1])), fdx2), cdx2);
    if (i <= n) ey[k] = div_const(exact_mul(exact_mul(0.5, exact_add(yf[kf], yf[kf + pf])), fdy2), cdy2);
}

// solve()'s per-cycle bookkeeping, variable-coefficient residual (see mg_diag_partial_kernel)
__global__ void __launch_bounds__(RED_THREADS)
mg_vc_diag_partial_kernel(MgLevel L, VcEdges E, double* __restrict__ old_phi, double* __restrict__ part)
{
    __shared__ double sh[RED_THREADS];
    if (L.ctl && L.ctl[CW_STOP]) return;
    double s_rel = 0.0, s_res = 0.0;
    const int n = L.n, P = L.pitch;
    const double* __restrict__ v = L.v;
    const double* __restrict__ f = L.f;
    double* __restrict__ r = L.r;
    for (int i = 1 + blockIdx.x; i <= L.ni; i += gridDim.x) {
        const long long base = (long long)i * P;
        for (int j0 = 1 + threadIdx.x; j0 <= n; j0 += RED_THREADS * RED_PER_THREAD) {
            double c[RED_PER_THREAD], up[RED_PER_THREAD], dn[RED_PER_THREAD], lf[RED_PER_THREAD],
                rt[RED_PER_THREAD], ff[RED_PER_THREAD], oo[RED_PER_THREAD], exl[RED_PER_THREAD],
                exh[RED_PER_THREAD], eyl[RED_PER_THREAD], eyh[RED_PER_THREAD];
#pragma unroll
            for (int u = 0; u < RED_PER_THREAD; ++u) {
                int j = j0 + u * RED_THREADS;
                bool ok = j <= n;
                long long k = base + (ok ? j : 1);
                c[u] = v[k]; up[u] = v[k - P]; dn[u] = v[k + P]; lf[u] = v[k - 1]; rt[u] = v[k + 1];
                ff[u] = f[k]; oo[u] = old_phi[k];
                exl[u] = E.ex[k]; exh[u] = E.ex[k + P]; eyl[u] = E.ey[k]; eyh[u] = E.ey[k + 1];
            }
#pragma unroll
            for (int u = 0; u < RED_PER_THREAD; ++u) {
                int j = j0 + u * RED_THREADS;
                if (j > n) continue;
                long long k = base + j;
                old_phi[k] = c[u];
                double d = (c[u] - oo[u]) / (c[u] + 1.e-16);
                s_rel += d * d;
                double res = vc_residual_value(ff[u], c[u], up[u], dn[u], lf[u], rt[u], exl[u], exh[u], eyl[u], eyh[u]);
                r[k] = res;
                s_res += res * res;
            }
        }
    }
    s_rel = block_sum(s_rel, sh);
    s_res = block_sum(s_res, sh);
    if (threadIdx.x == 0) { part[blockIdx.x] = s_rel; part[MG_NPART + blockIdx.x] = s_res; }
}

// Crank-Nicolson right-hand side of the diffusion solver (pyro/diffusion/simulation.py:87-91):
//   f = phi + coef * ((phi[i+1] + phi[i-1] - 2 phi)/dx**2 + (phi[j+1] + phi[j-1] - 2 phi)/dy**2),  coef = 0.5*dt*k
// phi is the solver's own ghost-filled (n+2)^2 plane (ng = 1); f is the level's right-hand-side plane
__global__ void mg_cn_rhs_kernel(MgLevel L, const double* __restrict__ phi, int ppitch, double coef, DivConst dx2,
                                 DivConst dy2)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int i = blockIdx.y * blockDim.y + threadIdx.y + 1;
    if (i > L.ni || j > L.n) return;
    const long long k = (long long)i * ppitch + j;
    const double c2 = exact_mul(2.0, phi[k]);
    const double lx = div_const(exact_sub(exact_add(phi[k + ppitch], phi[k - ppitch]), c2), dx2);
    const double ly = div_const(exact_sub(exact_add(phi[k + 1], phi[k - 1]), c2), dy2);
    L.f[(long long)i * L.pitch + j] = exact_add(phi[k], exact_mul(coef, exact_add(lx, ly)));
}

}  // namespace pyro

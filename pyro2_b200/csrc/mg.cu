// mg.cu -- cell-centred constant-coefficient multigrid V-cycle (HP-2) for (alpha - beta L) phi = f.
//
// Reference behaviour (pyro2, file:line):
//   CellCenterMG2d.smooth            pyro/multigrid/MG.py:544-599   red-black Gauss-Seidel, ghost fill
//                                                                   after each colour
//   CellCenterMG2d._compute_residual pyro/multigrid/MG.py:529-542
//   CellCenterData2d.restrict        pyro/mesh/patch.py:640-676     4-cell average
//   CellCenterData2d.prolong         pyro/mesh/patch.py:678-736     centred (unlimited) slopes
//   CellCenterMG2d.v_cycle / solve   pyro/multigrid/MG.py:699-778 / 623-697
//   ArrayIndexer.norm                pyro/mesh/array_indexer.py:98-111
//
// All arithmetic here is written with explicitly rounded, unfused operations in the reference's
// evaluation order: the path is HBM-bound, so the FP64 pipe has slack, and the payoff is that the
// device solution is BIT-IDENTICAL to the reference's (red-black ordering makes every point update
// independent of the traversal order).  Only the norms (a reduction) differ at the 1e-16 level.
//
// Ghost cells (ng = 1) are kept consistent by the thread that updates the interior source cell of
// each ghost ("fused fill_BC"), so a half-sweep is one launch.
#include <stdlib.h>

#include "common.cuh"
#include "mg_kernels.cuh"

struct p2b_mg {
    int nlevels;
    int bc[4];
    double alpha, beta, xmin, xmax, ymin, ymax;
    int nsmooth, nsmooth_bottom;
    pyro::MgLevel lev[pyro::MG_MAX_LEVELS];
    long long bytes;
    double* base;
    double* partials;                         // MG_NPART doubles x 2
    const double *xlv, *xrv, *ylv, *yrv;
    int no_blocking;                          // debugging / A-B switch: 1 = plain half-sweep kernels
    int rank, size;                           // x-slab decomposition (size 1 = single GPU)
    int split_level;                          // levels >= split_level are slabs when size > 1
    // peer-memory communication (mg_kernels.cuh): control words at the end of the workspace, the other ranks'
    // workspaces, and the bookkeeping of the running program
    unsigned long long* ctl;
    double* peer_base[pyro::MG_MAX_RANKS];
    int peers_set;
    int ord;                                  // ordinal of the last pushing launch of the running program
    int last_push[pyro::MG_MAX_LEVELS][4];    // ordinal of the launch that last filled the neighbours' halo rows of
                                              // (level, plane) in the running program; -1: delivered before it began
    // variable-coefficient mode (VarCoeffCCMG2d): per level the cell-centred eta and the two edge planes
    int varcoef;
    double *cc[pyro::MG_MAX_LEVELS], *ex[pyro::MG_MAX_LEVELS], *ey[pyro::MG_MAX_LEVELS];
};

namespace pyro {

static MgBC level_bc(const p2b_mg* m, int level)
{
    MgBC b;
    b.xl = m->bc[0]; b.xr = m->bc[1]; b.yl = m->bc[2]; b.yr = m->bc[3];
    if (level >= 0 && level < m->nlevels) {
        // a slab's x side that faces another slab (or, for periodic x, wraps onto one) has no BC
        if (!m->lev[level].xlo_phys) b.xl = P2B_BC_NONE;
        if (!m->lev[level].xhi_phys) b.xr = P2B_BC_NONE;
    }
    const bool fin = (level == m->nlevels - 1);
    // inhomogeneous values apply on the finest level only (MG.py:231-242) and only to
    // Dirichlet / Neumann sides (array_indexer.py:165-183)
    auto ok = [](int code) { return code == P2B_BC_OUTFLOW || code == P2B_BC_REFLECT_ODD; };
    b.xlv = (fin && ok(b.xl)) ? m->xlv : nullptr;
    b.xrv = (fin && ok(b.xr)) ? m->xrv : nullptr;
    b.ylv = (fin && ok(b.yl)) ? m->ylv : nullptr;
    b.yrv = (fin && ok(b.yr)) ? m->yrv : nullptr;
    return b;
}

static ResidCoef level_rcoef(const p2b_mg* m, const MgLevel& L)
{
    ResidCoef rc;
    rc.alpha = m->alpha; rc.beta = m->beta;
    rc.dx2 = make_div_const(L.dx * L.dx);
    rc.dy2 = make_div_const(L.dy * L.dy);
    return rc;
}

static SmoothCoef level_coef(const p2b_mg* m, const MgLevel& L)
{
    SmoothCoef c;
    c.alpha = m->alpha;
    c.xc = m->beta / (L.dx * L.dx);
    c.yc = m->beta / (L.dy * L.dy);
    c.denom = m->alpha + 2.0 * c.xc + 2.0 * c.yc;
    c.rden = 1.0 / c.denom;
    unsigned long long bits;
    memcpy(&bits, &c.denom, 8);
    const unsigned long long mant = bits & 0xFFFFFFFFFFFFFULL;
    c.fast = (mant != 0xFFFFFFFFFFFFFULL) && isfinite(c.rden) && c.denom != 0.0 &&
             fabs(c.denom) > 1e-290 && fabs(c.denom) < 1e290;
    return c;
}

static VcEdges level_edges(const p2b_mg* m, int level)
{
    VcEdges E;
    E.ex = m->ex[level]; E.ey = m->ey[level];
    return E;
}

// ---- communication descriptors -----------------------------------------------------------------------------
static MgComm comm_none()
{
    MgComm c;
    memset(&c, 0, sizeof c);
    c.ctl = nullptr; c.wait_ord = -1; c.sig_ord = -1; c.size = 1;
    return c;
}

static MgComm comm_base(const p2b_mg* m)
{
    MgComm c = comm_none();
    if (m->size <= 1) return c;
    const bool xper = (m->bc[0] == P2B_BC_PERIODIC);
    c.ctl = m->ctl; c.rank = m->rank; c.size = m->size;
    c.has_lo = (m->rank > 0 || xper) ? 1 : 0;
    c.has_hi = (m->rank < m->size - 1 || xper) ? 1 : 0;
    const int lo = (m->rank + m->size - 1) % m->size, hi = (m->rank + 1) % m->size;
    c.dlo = m->peer_base[lo] - m->base;
    c.dhi = m->peer_base[hi] - m->base;
    return c;
}

#ifdef P2B_EMU_HEADER
#define P2B_EMU_THREADED(flag) ::emu::force_threaded = (flag)
#else
#define P2B_EMU_THREADED(flag) ((void)0)
#endif

static bool is_slab(const p2b_mg* m, int level) { return m->size > 1 && level >= m->split_level; }

static void begin_program(p2b_mg* m, cudaStream_t st)
{
    if (m->size <= 1) return;
    P2B_LAUNCH(mg_epoch_kernel, 1, 1, 0, st)(m->ctl);
    m->ord = 0;
    for (int l = 0; l < MG_MAX_LEVELS; ++l) for (int k = 0; k < 4; ++k) m->last_push[l][k] = -1;
}

// every program ends by waiting for the neighbours' (all ranks') last push of it: when the next program starts,
// every halo row is in place and no flag of an older program is ever waited on
static void end_program(p2b_mg* m, cudaStream_t st, bool gather = false)
{
    if (m->size <= 1 || m->ord == 0) return;
    MgComm c = comm_base(m);
    c.wait_ord = m->ord;
    P2B_LAUNCH(mg_comm_wait_kernel, 1, 32, 0, st)(c, gather ? 1 : 0);
}

// stand-alone exchange of `depth` halo rows of one plane of a slab level (a program of its own, safe at any point:
// it first agrees with the neighbours that earlier work on both sides has finished)
static void exchange_impl(p2b_mg* m, int level, double* plane, int depth, cudaStream_t st)
{
    if (!is_slab(m, level)) return;
    const MgLevel& L = m->lev[level];
    MgComm c = comm_base(m);
    P2B_EMU_THREADED(true);
    P2B_LAUNCH(mg_xchg_arrive_kernel, 1, 1, 0, st)(c);
    dim3 grd((L.pitch + 255) / 256, depth);
    c.sig_ord = 1; c.n_lo = c.n_hi = (int)(grd.x * grd.y);
    P2B_LAUNCH(mg_xchg_push_kernel, grd, 256, 0, st)(plane, L.ni, L.pitch, depth, c);
    c.sig_ord = -1; c.wait_ord = 1;
    P2B_LAUNCH(mg_comm_wait_kernel, 1, 32, 0, st)(c, 0);
    P2B_EMU_THREADED(false);
}

// which geometry of the blocked pass a level uses (mg_kernels.cuh, tb_cfg): the throughput geometry when its grid
// gives the chip at least eight waves, otherwise the short-chain one (equal throughput at many waves -- 4096^2: 239 vs
// 256 us, 2048^2: 74 vs 74 -- but half the per-CTA time and two CTAs per SM: less wave quantisation on slabs).  P2B_TB_CFG=<k> forces one (A/B timing).
static int tb_choose(const p2b_mg* m, const MgLevel& L)
{
    static int forced = -2;
    if (forced == -2) {
        const char* e = getenv("P2B_TB_CFG");
        forced = e ? atoi(e) : -1;
        if (forced >= TB_NCFG) forced = -1;
    }
    if (forced >= 0) return forced;
    (void)m;
    const TbCfg big = tb_cfg(0);
    const long long ctas = (long long)((L.n + TB_TJ - 1) / TB_TJ) * ((L.ni + big.TI - 1) / big.TI);
    return ctas >= 8LL * num_sms() ? 0 : 1;
}

static void tb_launch(int cfg, dim3 grd, cudaStream_t st, const MgLevel& L, const double* src, double* dst, const MgBC& b,
                      const SmoothCoef& c, int niter, const MgComm& cm)
{
    auto k0 = mg_smooth_tb_kernel_t<8, 16, 1>;
    auto k1 = mg_smooth_tb_kernel_t<4, 16, 2>;
    auto k2 = mg_smooth_tb_kernel_t<8, 8, 2>;
    switch (cfg) {
        case 1: P2B_LAUNCH(k1, grd, 512, 0, st)(L, src, dst, b, c, niter, cm); break;
        case 2: P2B_LAUNCH(k2, grd, 256, 0, st)(L, src, dst, b, c, niter, cm); break;
        default: P2B_LAUNCH(k0, grd, 512, 0, st)(L, src, dst, b, c, niter, cm); break;
    }
}

constexpr int MG_SMALL_N = 64;   // levels up to 64^2 are smoothed by one CTA in one launch
constexpr int MG_TB_MIN_N = 128;  // from here up the temporally blocked kernel is used

static int smooth_impl(p2b_mg* m, int level, int nsmooth, bool fill_first, cudaStream_t st)
{
    const MgLevel& L = m->lev[level];
    MgBC b = level_bc(m, level);
    SmoothCoef c = level_coef(m, L);
    if (m->varcoef) {
        // variable_coeff_MG.py:112-171
        const VcEdges E = level_edges(m, level);
        if (L.n >= MG_TB_MIN_N && !m->no_blocking && nsmooth > 0) {
            // the blocked smoother with the coefficient tiles staged in shared memory
            static bool attr_set = false;
            if (!attr_set) {
                cudaFuncSetAttribute(mg_vc_smooth_tb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TB_VC_SMEM_BYTES);
                attr_set = true;
            }
            dim3 grd((L.n + TB_TJ - 1) / TB_TJ, (L.n + TB_TI - 1) / TB_TI);
            const double* src = L.v;
            double* dst = L.w;
            for (int left = nsmooth; left > 0;) {
                int it = left < TB_K ? left : TB_K;
                P2B_LAUNCH(mg_vc_smooth_tb_kernel, grd, 32 * TB_NW, TB_VC_SMEM_BYTES, st)(L, src, dst, b, E, it);
                left -= it;
                const double* t = src; src = dst; dst = const_cast<double*>(t);
            }
            if (src != L.v)
                cudaMemcpyAsync(L.v, L.w, (size_t)(L.n + 2) * L.pitch * sizeof(double), cudaMemcpyDeviceToDevice, st);
            return P2B_OK;
        }
        if (fill_first) P2B_LAUNCH(mg_fill_kernel, (4 * L.n + 255) / 256, 256, 0, st)(L, b);
        if (L.n <= MG_SMALL_N) {
            int threads = L.n * (L.n / 2);
            threads = threads < 32 ? 32 : (threads > 1024 ? 1024 : threads);
            P2B_LAUNCH(mg_vc_smooth_small_kernel, 1, threads, 0, st)(L, b, E, nsmooth);
        } else {
            dim3 blk(64, 4);
            dim3 grd((L.n / 2 + blk.x - 1) / blk.x, (L.n + blk.y - 1) / blk.y);
            for (int it = 0; it < nsmooth; ++it) {
                P2B_LAUNCH(mg_vc_halfsweep_kernel, grd, blk, 0, st)(L, b, E, 0);
                P2B_LAUNCH(mg_vc_halfsweep_kernel, grd, blk, 0, st)(L, b, E, 1);
            }
        }
        return P2B_OK;
    }
    if (L.n <= MG_SMALL_N) {
        if (fill_first) P2B_LAUNCH(mg_fill_kernel, (4 * L.n + 255) / 256, 256, 0, st)(L, b);
        int threads = L.n * (L.n / 2);
        threads = threads < 32 ? 32 : (threads > 1024 ? 1024 : threads);
        P2B_LAUNCH(mg_smooth_small_kernel, 1, threads, 0, st)(L, b, c, nsmooth);
    } else if (L.n >= MG_TB_MIN_N && !m->no_blocking) {
        // passes of up to TB_K iterations, ping-ponging v <-> w; the blocked kernel derives ghost
        // values from interior cells itself, so no separate fill_BC is needed before it
        const int cfg = tb_choose(m, L);
        dim3 grd((L.n + TB_TJ - 1) / TB_TJ, (L.n + tb_cfg(cfg).TI - 1) / tb_cfg(cfg).TI);
        const double* src = L.v;
        double* dst = L.w;
        int left = nsmooth;
        if (left == 0 && fill_first) P2B_LAUNCH(mg_fill_kernel, (4 * L.n + 255) / 256, 256, 0, st)(L, b);
        while (left > 0) {
            int it = left < TB_K ? left : TB_K;
            tb_launch(cfg, grd, st, L, src, dst, b, c, it, comm_none());
            left -= it;
            const double* t = src; src = dst; dst = const_cast<double*>(t);
        }
        if (src != L.v)   // odd number of passes: the result sits in w
            cudaMemcpyAsync(L.v, L.w, (size_t)(L.n + 2) * L.pitch * sizeof(double), cudaMemcpyDeviceToDevice, st);
    } else {
        if (fill_first) P2B_LAUNCH(mg_fill_kernel, (4 * L.n + 255) / 256, 256, 0, st)(L, b);
        dim3 blk(64, 4);
        dim3 grd((L.n / 2 + blk.x - 1) / blk.x, (L.n + blk.y - 1) / blk.y);
        for (int it = 0; it < nsmooth; ++it) {
            P2B_LAUNCH(mg_halfsweep_kernel, grd, blk, 0, st)(L, b, c, 0);
            P2B_LAUNCH(mg_halfsweep_kernel, grd, blk, 0, st)(L, b, c, 1);
        }
    }
    return P2B_OK;
}

// rows of the coarse level that correspond to this rank's fine rows start at ioff_fine / 2; that is
// a non-zero offset into the coarse array only when the coarse level is replicated (ioff = 0 there)
static int coarse_row_offset(const MgLevel& F, const MgLevel& Cs) { return F.ioff / 2 - Cs.ioff; }

static int imax(int a, int b) { return a > b ? a : b; }

// one blocked pass on a level.  On a slab level the launch waits (edge tiles only) for the halo rows of its source
// plane and of f, and pushes the first / last TB_H rows of its result into the neighbours' halo rows.
static void tb_pass_impl(p2b_mg* m, int level, int src, int dst, int niter, cudaStream_t st)
{
    const MgLevel& L = m->lev[level];
    const int cfg = tb_choose(m, L);
    const int TI = tb_cfg(cfg).TI;
    dim3 grd((L.n + TB_TJ - 1) / TB_TJ, (L.ni + TI - 1) / TI);
    MgComm c = comm_none();
    if (is_slab(m, level)) {
        c = comm_base(m);
        c.wait_ord = imax(m->last_push[level][src], m->last_push[level][1]);
        c.sig_ord = ++m->ord;
        // tile rows that store first / last owned rows push them; the kernel's last pushing CTA of a side signals
        int rows_lo = 0, rows_hi = 0;
        for (int ty = 0; ty < (int)grd.y; ++ty) {
            const int I0 = tb_tile_origin(ty, TI, L.ni, !L.xhi_phys);
            rows_lo += tb_pushes_lo(I0) ? 1 : 0;
            rows_hi += tb_pushes_hi(I0, TI, L.ni) ? 1 : 0;
        }
        c.n_lo = (int)grd.x * rows_lo;
        c.n_hi = (int)grd.x * rows_hi;
        m->last_push[level][dst] = c.sig_ord;
    }
    const double* sp = src == 0 ? L.v : L.w;
    double* dp = dst == 0 ? L.v : L.w;
    tb_launch(cfg, grd, st, L, sp, dp, level_bc(m, level), level_coef(m, L), niter, c);
}

// nsmooth red-black iterations on a slab level: passes of <= TB_K iterations, halo rows travelling in the passes' own
// epilogues (communication-avoiding: one message per 5 iterations instead of one per colour)
static void smooth_slab_impl(p2b_mg* m, int level, int nsmooth, cudaStream_t st)
{
    const MgLevel& L = m->lev[level];
    int src = 0, dst = 3;
    for (int left = nsmooth; left > 0;) {
        const int it = left < TB_K ? left : TB_K;
        tb_pass_impl(m, level, src, dst, it, st);
        left -= it;
        const int t = src; src = dst; dst = t;
    }
    if (src != 0) {
        // odd number of passes: the result sits in w.  Its halo rows are in flight: wait, then copy the whole plane.
        MgComm c = comm_base(m);
        c.wait_ord = m->last_push[level][3];
        P2B_LAUNCH(mg_comm_wait_kernel, 1, 32, 0, st)(c, 0);
        const long long rows = L.ni + 2 * L.gx;
        cudaMemcpyAsync(L.v - (long long)(L.gx - 1) * L.pitch, L.w - (long long)(L.gx - 1) * L.pitch,
                        (size_t)rows * L.pitch * sizeof(double), cudaMemcpyDeviceToDevice, st);
        m->last_push[level][0] = -1;
    }
}

static void residual_impl(p2b_mg* m, int level, cudaStream_t st)
{
    const MgLevel& L = m->lev[level];
    dim3 blk(64, 4);
    dim3 grd((L.n + blk.x - 1) / blk.x, (L.ni + blk.y - 1) / blk.y);
    if (m->varcoef) {
        P2B_LAUNCH(mg_vc_residual_kernel, grd, blk, 0, st)(L, level_edges(m, level));
        return;
    }
    MgComm c = comm_none();
    if (is_slab(m, level)) { c = comm_base(m); c.wait_ord = m->last_push[level][0]; }
    P2B_LAUNCH(mg_residual_kernel, grd, blk, 0, st)(L, level_rcoef(m, L), c);
}

static void restrict_impl(p2b_mg* m, int level, cudaStream_t st)
{
    const MgLevel &F = m->lev[level], &Cs = m->lev[level - 1];
    dim3 blk(64, 4);
    const int nic = F.ni / 2;
    dim3 grd((Cs.n + blk.x - 1) / blk.x, (nic + blk.y - 1) / blk.y);
    MgComm c = comm_none();
    int to_all = 0;
    if (is_slab(m, level)) {
        c = comm_base(m);
        c.sig_ord = ++m->ord;
        to_all = is_slab(m, level - 1) ? 0 : 1;
        if (!to_all) {
            c.n_lo = (int)grd.x * count_rows_meet((int)grd.y, (int)blk.y, 1, Cs.gx);
            c.n_hi = (int)grd.x * count_rows_meet((int)grd.y, (int)blk.y, nic - Cs.gx + 1, nic);
            m->last_push[level - 1][1] = c.sig_ord;
        }
    }
    P2B_LAUNCH(mg_restrict_kernel, grd, blk, 0, st)(F, Cs, coarse_row_offset(F, Cs), c, to_all);
    if (to_all) {
        // slab -> replicated: wait until every rank's rows of the coarse right-hand side have landed here
        c.sig_ord = -1; c.wait_ord = m->ord;
        P2B_LAUNCH(mg_comm_wait_kernel, 1, 32, 0, st)(c, 1);
    }
}

static void prolong_impl(p2b_mg* m, int level, cudaStream_t st)
{
    const MgLevel &F = m->lev[level], &Cs = m->lev[level - 1];
    dim3 blk(64, 4);
    const int nic = F.ni / 2;
    dim3 grd((Cs.n + blk.x - 1) / blk.x, (nic + blk.y - 1) / blk.y);
    MgComm c = comm_none();
    int coarse_slab = 0;
    if (is_slab(m, level)) {
        c = comm_base(m);
        coarse_slab = is_slab(m, level - 1) ? 1 : 0;
        c.wait_ord = coarse_slab ? m->last_push[level - 1][0] : -1;
        c.sig_ord = ++m->ord;
        const int hp = F.gx / 2;
        c.n_lo = (int)grd.x * count_rows_meet((int)grd.y, (int)blk.y, 1, hp);
        c.n_hi = (int)grd.x * count_rows_meet((int)grd.y, (int)blk.y, nic - hp + 1, nic);
        m->last_push[level][0] = c.sig_ord;
    }
    P2B_LAUNCH(mg_prolong_kernel, grd, blk, 0, st)(F, Cs, level_bc(m, level), coarse_row_offset(F, Cs), c, coarse_slab);
}

static int coarse_top(const p2b_mg* m)
{
    int top = -1;
    for (int l = 0; l < m->nlevels && l < MG_COARSE_LEVELS; ++l)
        if (m->lev[l].n <= MG_COARSE_TOP_N) top = l;
    return top;
}

static void coarse_vcycle_impl(p2b_mg* m, int top, cudaStream_t st)
{
    CoarseTable T;
    memset(&T, 0, sizeof T);
    size_t bytes = 0;
    for (int l = 0; l <= top; ++l) {
        T.g[l] = m->lev[l];
        T.coef[l] = level_coef(m, m->lev[l]);
        int q = m->lev[l].n + 2;
        bytes += 3 * (size_t)q * ((q + 1) & ~1) * sizeof(double);
    }
    T.top = top;
    T.nsmooth = m->nsmooth; T.nsmooth_bottom = m->nsmooth_bottom;
    for (int l = 0; l <= top; ++l) T.rcoef[l] = level_rcoef(m, m->lev[l]);
    T.bc_top = level_bc(m, top);
    T.bc_coarse = level_bc(m, top == m->nlevels - 1 ? -1 : 0);   // homogeneous
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(mg_coarse_vcycle_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(mg_coarse_vcycle_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    if (m->varcoef) {
        for (int l = 0; l <= top; ++l) T.edges[l] = level_edges(m, l);
        P2B_LAUNCH(mg_coarse_vcycle_kernel<true>, 1, MG_COARSE_THREADS, bytes, st)(T);
    } else {
        P2B_LAUNCH(mg_coarse_vcycle_kernel<false>, 1, MG_COARSE_THREADS, bytes, st)(T);
    }
}

static void vcycle_impl(p2b_mg* m, int level, cudaStream_t st)
{
    // MG.py:699-778
    if (is_slab(m, level)) {
        // x-slab level: same operations, halo rows pushed by the producers (see mg_kernels.cuh)
        smooth_slab_impl(m, level, m->nsmooth, st);
        residual_impl(m, level, st);
        restrict_impl(m, level, st);
        vcycle_impl(m, level - 1, st);
        prolong_impl(m, level, st);
        smooth_slab_impl(m, level, m->nsmooth, st);
        return;
    }
    static const bool no_fused_coarse = getenv("P2B_MG_NO_FUSED_COARSE") != nullptr;     // A/B switch (development)
    if (!m->no_blocking && !no_fused_coarse && level <= coarse_top(m)) {
        coarse_vcycle_impl(m, level, st);
        return;
    }
    if (level > 0) {
        smooth_impl(m, level, m->nsmooth, true, st);
        residual_impl(m, level, st);
        restrict_impl(m, level, st);
        vcycle_impl(m, level - 1, st);
        prolong_impl(m, level, st);
        smooth_impl(m, level, m->nsmooth, true, st);
    } else {
        smooth_impl(m, 0, m->nsmooth_bottom, true, st);
    }
}

static int sumsq_impl(p2b_mg* m, const double* a, int level, double* out, cudaStream_t st)
{
    const MgLevel& L = m->lev[level];
    int blocks = L.ni < MG_NPART ? L.ni : MG_NPART;
    P2B_LAUNCH(mg_sumsq_partial_kernel, blocks, RED_THREADS, 0, st)(a, L.ni, L.n, L.pitch, m->partials);
    // a slab level's sum runs over all ranks (a program of its own: the kernel starts it)
    P2B_LAUNCH(mg_sumsq_final_kernel, 1, RED_THREADS, 0, st)(m->partials, blocks, out, is_slab(m, level) ? comm_base(m) : comm_none());
    return P2B_OK;
}

}  // namespace pyro

using namespace pyro;

extern "C" {

// CUDA loads kernels lazily (CUDA_MODULE_LOADING=LAZY is the default since 12.2): the FIRST launch of a kernel loads it,
// and loading synchronises with the kernels running on the device.  A rank's kernel that spins on a flag another rank's
// not-yet-loaded kernel will raise then never sees it: the load waits for the spin, the spin for the load (observed with
// two slabs on one GPU: the first all-reduce timed out whenever the diagnostics kernels had not run before).  Loading
// every kernel of this translation unit when the first hierarchy is created removes the hazard.
static void preload_kernels()
{
#ifndef P2B_EMU_HEADER
    static bool done = false;
    if (done) return;
    done = true;
    cudaFuncAttributes a;
#define P2B_PRELOAD(k) cudaFuncGetAttributes(&a, k)
    P2B_PRELOAD(mg_halfsweep_kernel); P2B_PRELOAD(mg_smooth_small_kernel); P2B_PRELOAD(mg_fill_kernel);
    P2B_PRELOAD((mg_smooth_tb_kernel_t<8, 16, 1>)); P2B_PRELOAD((mg_smooth_tb_kernel_t<4, 16, 2>));
    P2B_PRELOAD((mg_smooth_tb_kernel_t<8, 8, 2>)); P2B_PRELOAD(mg_vc_smooth_tb_kernel);
    P2B_PRELOAD(mg_coarse_vcycle_kernel<false>); P2B_PRELOAD(mg_coarse_vcycle_kernel<true>);
    P2B_PRELOAD(mg_residual_kernel); P2B_PRELOAD(mg_restrict_kernel); P2B_PRELOAD(mg_prolong_kernel);
    P2B_PRELOAD(mg_sumsq_partial_kernel); P2B_PRELOAD(mg_sumsq_final_kernel); P2B_PRELOAD(mg_zero_kernel);
    P2B_PRELOAD(mg_diag_partial_kernel); P2B_PRELOAD(mg_diag_final_kernel); P2B_PRELOAD(mg_set_stop_kernel);
    P2B_PRELOAD(mg_epoch_kernel); P2B_PRELOAD(mg_comm_wait_kernel); P2B_PRELOAD(mg_xchg_arrive_kernel);
    P2B_PRELOAD(mg_xchg_push_kernel); P2B_PRELOAD(mg_cn_rhs_kernel);
    P2B_PRELOAD(mg_vc_halfsweep_kernel); P2B_PRELOAD(mg_vc_smooth_small_kernel); P2B_PRELOAD(mg_vc_residual_kernel);
    P2B_PRELOAD(mg_vc_diag_partial_kernel); P2B_PRELOAD(mg_vc_edges_fine_kernel); P2B_PRELOAD(mg_vc_edges_restrict_kernel);
#undef P2B_PRELOAD
    cudaGetLastError();
#endif
}

static p2b_mg* mg_create_impl(int nx, const int* bc, double alpha, double beta, double xmin, double xmax,
                              double ymin, double ymax, int nsmooth, int nsmooth_bottom, int rank, int size,
                              int split_n)
{
    if (nx < 2 || (nx & (nx - 1)) != 0) { set_error("multigrid requires nx = ny = power of two (got %d)", nx); return nullptr; }
    if (!bc) { set_error("null bc"); return nullptr; }
    if (size < 1 || (size & (size - 1)) != 0 || rank < 0 || rank >= size || size > MG_MAX_RANKS) {
        set_error("slab count must be a power of two <= %d", MG_MAX_RANKS);
        return nullptr;
    }
    preload_kernels();
    p2b_mg* m = new p2b_mg();
    memset(m, 0, sizeof *m);
    int nl = 0;
    while ((2 << nl) < nx) ++nl;
    m->nlevels = nl + 1;
    for (int s = 0; s < 4; ++s) m->bc[s] = bc[s];
    m->alpha = alpha; m->beta = beta;
    m->xmin = xmin; m->xmax = xmax; m->ymin = ymin; m->ymax = ymax;
    m->nsmooth = nsmooth; m->nsmooth_bottom = nsmooth_bottom;
    m->rank = rank; m->size = size;
    // levels with at least split_n columns are x-slabs; each slab must keep >= 2*TB_H rows so a halo
    // never reaches past the neighbour's owned rows
    m->split_level = m->nlevels;
    if (size > 1) {
        for (int l = m->nlevels - 1; l >= 0; --l) {
            int n = 2 << l;
            if (n >= split_n && n / size >= 2 * TB_H && n >= MG_TB_MIN_N) m->split_level = l; else break;
        }
        if (m->split_level >= m->nlevels) { set_error("grid too small for %d slabs", size); delete m; return nullptr; }
    }
    const bool xper = (bc[0] == P2B_BC_PERIODIC);
    // layout: per level four consecutive planes v, f, r, w (uniform plane stride, so a level is one
    // strided tensor on the Python side); then the norm partials
    long long off = 0;
    for (int l = 0; l < m->nlevels; ++l) {
        MgLevel& L = m->lev[l];
        L.n = 2 << l;
        const bool slab = (size > 1 && l >= m->split_level);
        L.ni = slab ? L.n / size : L.n;
        L.ioff = slab ? rank * L.ni : 0;
        L.gx = slab ? TB_H : 1;
        L.xlo_phys = !slab || (rank == 0 && !xper);
        L.xhi_phys = !slab || (rank == size - 1 && !xper);
        int q = L.n + 2;
        L.pitch = (q >= 16) ? (q + 15) / 16 * 16 : (q + 1) / 2 * 2;
        L.dx = (xmax - xmin) / L.n;
        L.dy = (ymax - ymin) / L.n;
        const long long rows = L.ni + 2 * L.gx;
        const long long plane = rows * L.pitch;
        const long long row0 = (long long)(L.gx - 1) * L.pitch;    // pointer = address of row index 0
        // offsets are stored as fake pointers (element counts) until p2b_mg_bind
        L.v = (double*)(off + row0); off += plane;
        L.f = (double*)(off + row0); off += plane;
        L.r = (double*)(off + row0); off += plane;
        L.w = (double*)(off + row0); off += plane;
    }
    m->partials = (double*)off; off += 2 * MG_NPART;
    m->ctl = (unsigned long long*)off; off += CW_WORDS;       // control words (communication, stopping rule)
    m->bytes = off * 8;
    return m;
}

// CellCenterMG2d.__init__ (MG.py:85-295): level l has 2^(l+1) cells per side, ng = 1, vars v, f, r
p2b_mg* p2b_mg_create(int nx, const int* bc, double alpha, double beta, double xmin, double xmax,
                      double ymin, double ymax, int nsmooth, int nsmooth_bottom)
{
    return mg_create_impl(nx, bc, alpha, beta, xmin, xmax, ymin, ymax, nsmooth, nsmooth_bottom, 0, 1, 0);
}

// multi-GPU: this process owns x-slab `rank` of `size` on every level with >= split_n columns;
// coarser levels are replicated on all ranks
p2b_mg* p2b_mg_create_slab(int nx, const int* bc, double alpha, double beta, double xmin, double xmax,
                           double ymin, double ymax, int nsmooth, int nsmooth_bottom, int rank, int size,
                           int split_n)
{
    return mg_create_impl(nx, bc, alpha, beta, xmin, xmax, ymin, ymax, nsmooth, nsmooth_bottom, rank, size, split_n);
}

// geometry of a level: out = {ni, n, pitch, gx, ioff, is_slab, plane_stride_lo, plane_stride_hi}
int p2b_mg_level_info(p2b_mg* m, int level, long long* out)
{
    P2B_REQUIRE(m && out && level >= 0 && level < m->nlevels, "bad level");
    const MgLevel& L = m->lev[level];
    out[0] = L.ni; out[1] = L.n; out[2] = L.pitch; out[3] = L.gx; out[4] = L.ioff;
    out[5] = (m->size > 1 && level >= m->split_level) ? 1 : 0;
    out[6] = (long long)(L.ni + 2 * L.gx) * L.pitch;
    out[7] = m->split_level;
    return P2B_OK;
}

int p2b_mg_destroy(p2b_mg* m) { delete m; return P2B_OK; }
int p2b_mg_set_blocking(p2b_mg* m, int enable) { if (!m) return P2B_EINVAL; m->no_blocking = !enable; return P2B_OK; }
int p2b_mg_nlevels(p2b_mg* m) { return m ? m->nlevels : 0; }
long long p2b_mg_workspace_bytes(p2b_mg* m) { return m ? m->bytes : 0; }

// hand the hierarchy its (caller-owned, zero-initialised) device memory
int p2b_mg_bind(p2b_mg* m, void* mem, long long bytes)
{
    P2B_REQUIRE(m && mem, "null pointer");
    P2B_REQUIRE(bytes >= m->bytes, "workspace too small");
    P2B_REQUIRE(m->base == nullptr, "already bound");
    P2B_REQUIRE(((uintptr_t)mem % 16) == 0, "workspace must be 16-byte aligned");
    m->base = (double*)mem;
    for (int l = 0; l < m->nlevels; ++l) {
        m->lev[l].v = m->base + (long long)m->lev[l].v;
        m->lev[l].f = m->base + (long long)m->lev[l].f;
        m->lev[l].r = m->base + (long long)m->lev[l].r;
        m->lev[l].w = m->base + (long long)m->lev[l].w;
    }
    m->partials = m->base + (long long)m->partials;
    m->ctl = reinterpret_cast<unsigned long long*>(m->base + (long long)m->ctl);
    for (int l = 0; l < m->nlevels; ++l) m->lev[l].ctl = m->ctl;
    m->peer_base[m->rank] = m->base;
    return P2B_OK;
}

// decomposed hierarchy: bases[r] = rank r's workspace as mapped into THIS process (bases[rank] = the pointer given to
// p2b_mg_bind); see p2b_shared_* for the mapping across processes.  Needed before any operation on a slab level.
int p2b_mg_set_peers(p2b_mg* m, void* const* bases)
{
    P2B_REQUIRE(m && m->base && bases, "hierarchy not bound");
    P2B_REQUIRE(bases[m->rank] == (void*)m->base, "bases[rank] must be this rank's own workspace");
    unsigned long long off[MG_MAX_RANKS] = {0};
    for (int r = 0; r < m->size; ++r) {
        P2B_REQUIRE(bases[r] && ((uintptr_t)bases[r] % 16) == 0, "bad peer workspace");
        m->peer_base[r] = (double*)bases[r];
        off[r] = (unsigned long long)(m->peer_base[r] - m->base);
    }
    P2B_CUDA_CHECK(cudaMemcpy(m->ctl + CW_PEER, off, sizeof off, cudaMemcpyHostToDevice));
    m->peers_set = 1;
    return P2B_OK;
}

void* p2b_mg_level_ptr(p2b_mg* m, int level, int which)
{
    if (!m || level < 0 || level >= m->nlevels) return nullptr;
    return which == 0 ? m->lev[level].v : which == 1 ? m->lev[level].f : which == 2 ? m->lev[level].r : m->lev[level].w;
}

int p2b_mg_level_pitch(p2b_mg* m, int level) { return (m && level >= 0 && level < m->nlevels) ? m->lev[level].pitch : 0; }

// inhomogeneous boundary values for the finest level: device arrays of n+2 doubles or NULL
int p2b_mg_set_bc_values(p2b_mg* m, const double* xl, const double* xr, const double* yl, const double* yr)
{
    P2B_REQUIRE(m, "null handle");
    m->xlv = xl; m->xrv = xr; m->ylv = yl; m->yrv = yr;
    return P2B_OK;
}

#define MG_CHECK_LEVEL(m, level) \
    P2B_REQUIRE((m) && (m)->base, "hierarchy not bound"); \
    P2B_REQUIRE((level) >= 0 && (level) < (m)->nlevels, "bad level"); \
    P2B_REQUIRE(!is_slab(m, level) || (m)->peers_set, "decomposed hierarchy: call p2b_mg_set_peers first")

int p2b_mg_smooth(p2b_mg* m, int level, int nsmooth, void* stream)
{
    MG_CHECK_LEVEL(m, level);
    if (is_slab(m, level)) {
        // a stand-alone smooth(): bring the halo rows of v up to date (the caller may have changed v), then the
        // passes as one program
        cudaStream_t st = (cudaStream_t)stream;
        P2B_REQUIRE(!m->varcoef && m->lev[level].n >= MG_TB_MIN_N, "slab levels use the blocked constant-coefficient smoother");
        exchange_impl(m, level, m->lev[level].v, TB_H, st);
        P2B_EMU_THREADED(true);
        begin_program(m, st);
        smooth_slab_impl(m, level, nsmooth, st);
        end_program(m, st);
        P2B_EMU_THREADED(false);
        P2B_CUDA_CHECK(cudaGetLastError());
        return P2B_OK;
    }
    smooth_impl(m, level, nsmooth, true, (cudaStream_t)stream);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

int p2b_mg_residual(p2b_mg* m, int level, void* stream)
{
    MG_CHECK_LEVEL(m, level);
    if (is_slab(m, level)) exchange_impl(m, level, m->lev[level].v, 1, (cudaStream_t)stream);   // rows 0 and ni + 1
    if (is_slab(m, level)) for (int k = 0; k < 4; ++k) m->last_push[level][k] = -1;
    residual_impl(m, level, (cudaStream_t)stream);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

int p2b_mg_restrict(p2b_mg* m, int level, void* stream)
{
    MG_CHECK_LEVEL(m, level);
    P2B_REQUIRE(level >= 1, "no coarser level");
    P2B_EMU_THREADED(is_slab(m, level));
    if (is_slab(m, level)) begin_program(m, (cudaStream_t)stream);
    restrict_impl(m, level, (cudaStream_t)stream);
    if (is_slab(m, level) && is_slab(m, level - 1)) end_program(m, (cudaStream_t)stream);
    P2B_EMU_THREADED(false);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

int p2b_mg_prolong_correct(p2b_mg* m, int level, void* stream)
{
    MG_CHECK_LEVEL(m, level);
    P2B_REQUIRE(level >= 1, "no coarser level");
    if (is_slab(m, level)) {
        cudaStream_t st = (cudaStream_t)stream;
        if (is_slab(m, level - 1)) exchange_impl(m, level - 1, m->lev[level - 1].v, 1, st);   // coarse rows 0, nic + 1
        P2B_EMU_THREADED(true);
        begin_program(m, st);
        prolong_impl(m, level, st);
        end_program(m, st);
        P2B_EMU_THREADED(false);
        P2B_CUDA_CHECK(cudaGetLastError());
        return P2B_OK;
    }
    prolong_impl(m, level, (cudaStream_t)stream);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

int p2b_mg_fill_bc(p2b_mg* m, int level, void* stream)
{
    MG_CHECK_LEVEL(m, level);
    if (is_slab(m, level)) {
        // the slab's x "ghost" rows are the neighbours' rows; its physical sides were kept by the writers
        exchange_impl(m, level, m->lev[level].v, 1, (cudaStream_t)stream);
        P2B_CUDA_CHECK(cudaGetLastError());
        return P2B_OK;
    }
    const MgLevel& L = m->lev[level];
    P2B_LAUNCH(mg_fill_kernel, (4 * L.n + 255) / 256, 256, 0, (cudaStream_t)stream)(L, level_bc(m, level));
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

// zero v on every level but the finest (MG.py:658-659), one launch
int p2b_mg_zero_coarse(p2b_mg* m, void* stream)
{
    P2B_REQUIRE(m && m->base, "hierarchy not bound");
    if (m->nlevels < 2) return P2B_OK;
    MgZeroTable t;
    t.ctl = m->ctl;
    t.nlev = m->nlevels - 1;
    long long most = 0;
    for (int l = 0; l < t.nlev; ++l) {
        t.v[l] = m->lev[l].v - (long long)(m->lev[l].gx - 1) * m->lev[l].pitch;
        t.count[l] = (long long)(m->lev[l].ni + 2 * m->lev[l].gx) * m->lev[l].pitch;
        if (t.count[l] > most) most = t.count[l];
    }
    long long blocks = (most + 255) / 256;
    if (blocks > 8LL * num_sms()) blocks = 8LL * num_sms();
    P2B_LAUNCH(mg_zero_kernel, (int)blocks, 256, 0, (cudaStream_t)stream)(t);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

// one pass (<= 5 red-black iterations) of the temporally blocked smoother on `level`, reading plane
// src (0 = v, 3 = w) and writing dst (3 = w, 0 = v).  On slab levels the caller exchanges TB_H halo
// rows of the source plane first (p2b_mg_tb_halo()).
int p2b_mg_tb_pass(p2b_mg* m, int level, int src, int dst, int niter, void* stream)
{
    MG_CHECK_LEVEL(m, level);
    P2B_REQUIRE((src == 0 && dst == 3) || (src == 3 && dst == 0), "src/dst must be v->w or w->v");
    P2B_REQUIRE(niter >= 1 && niter <= TB_K, "niter out of range");
    P2B_REQUIRE(m->lev[level].n >= MG_TB_MIN_N, "level too small for the blocked smoother");
    P2B_REQUIRE(!m->varcoef, "the blocked smoother is constant-coefficient only");
    if (is_slab(m, level)) {
        cudaStream_t st = (cudaStream_t)stream;
        exchange_impl(m, level, src == 0 ? m->lev[level].v : m->lev[level].w, TB_H, st);
        P2B_EMU_THREADED(true);
        begin_program(m, st);
        tb_pass_impl(m, level, src, dst, niter, st);
        end_program(m, st);
        P2B_EMU_THREADED(false);
    } else {
        tb_pass_impl(m, level, src, dst, niter, (cudaStream_t)stream);
    }
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

int p2b_mg_tb_halo(void) { return TB_H; }
int p2b_mg_tb_iters(void) { return TB_K; }

// the sub-V-cycle from `level` down and back up (v_cycle(level), MG.py:699-778); used for the
// replicated coarse levels of a decomposed hierarchy
int p2b_mg_vcycle_level(p2b_mg* m, int level, void* stream)
{
    MG_CHECK_LEVEL(m, level);
    P2B_REQUIRE(!is_slab(m, level), "level is decomposed");
    vcycle_impl(m, level, (cudaStream_t)stream);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

int p2b_mg_vcycle(p2b_mg* m, void* stream)
{
    P2B_REQUIRE(m && m->base, "hierarchy not bound");
    P2B_REQUIRE(m->size == 1 || m->peers_set, "decomposed hierarchy: call p2b_mg_set_peers first");
    // decomposed: one program -- every halo row travels in a producer's epilogue, nothing but kernels is enqueued
    // (capturable as a CUDA graph); the halo rows of the finest v and f must be current on entry (p2b_mg_exchange)
    P2B_EMU_THREADED(m->size > 1);
    begin_program(m, (cudaStream_t)stream);
    vcycle_impl(m, m->nlevels - 1, (cudaStream_t)stream);
    end_program(m, (cudaStream_t)stream);
    P2B_EMU_THREADED(false);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

// sum of squares of plane `which` (0 v, 1 f, 2 r) over the valid region -> *out (device double);
// ArrayIndexer.norm = sqrt(dx*dy*sum) (array_indexer.py:98-111)
int p2b_mg_norm2(p2b_mg* m, int level, int which, double* out, void* stream)
{
    MG_CHECK_LEVEL(m, level);
    P2B_REQUIRE(out, "null out");
    const MgLevel& L = m->lev[level];
    const double* a = which == 0 ? L.v : which == 1 ? L.f : L.r;
    sumsq_impl(m, a, level, out, (cudaStream_t)stream);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

// per-cycle bookkeeping of solve() (MG.py:668-686) on the finest level:
//   out[0] = sum(((v - old_phi)/(v + 1e-16))^2), old_phi <- v, r <- residual, out[1] = sum(r^2)
int p2b_mg_cycle_diagnostics(p2b_mg* m, double* old_phi, double* out, void* stream)
{
    P2B_REQUIRE(m && m->base && old_phi && out, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int lf = m->nlevels - 1;
    const MgLevel& L = m->lev[lf];
    int blocks = L.ni < MG_NPART ? L.ni : MG_NPART;
    if (m->varcoef)
        P2B_LAUNCH(mg_vc_diag_partial_kernel, blocks, RED_THREADS, 0, st)(L, level_edges(m, lf), old_phi, m->partials);
    else
        P2B_LAUNCH(mg_diag_partial_kernel, blocks, RED_THREADS, 0, st)(L, old_phi, level_rcoef(m, L), m->partials);
    P2B_LAUNCH(mg_diag_final_kernel, 1, RED_THREADS, 0, st)(m->partials, blocks, out, is_slab(m, lf) ? comm_base(m) : comm_none(),
                                                             m->ctl, L.dx * L.dy);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

// stand-alone halo exchange of `depth` rows of plane `which` (0 v, 1 f, 2 r, 3 w) of a slab level with the neighbouring
// ranks, through peer memory (no-op on a replicated level).  Collective: every rank calls it at the same point.
int p2b_mg_exchange(p2b_mg* m, int level, int which, int depth, void* stream)
{
    MG_CHECK_LEVEL(m, level);
    P2B_REQUIRE(which >= 0 && which <= 3, "bad plane");
    const MgLevel& L = m->lev[level];
    P2B_REQUIRE(depth >= 1 && depth <= L.gx && depth <= L.ni, "bad depth");
    double* plane = which == 0 ? L.v : which == 1 ? L.f : which == 2 ? L.r : L.w;
    exchange_impl(m, level, plane, depth, (cudaStream_t)stream);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

// the stopping rule of solve() on the device (MG.py:654-697).  enable = 1: each p2b_mg_cycle_diagnostics also evaluates
// residual_error = sqrt(dx dy sum r^2) / source_norm, counts the cycle and, when residual_error <= rtol or max_cycles
// cycles have run, raises the stop word that turns every kernel of cycles enqueued ahead into a no-op.  Calling it
// (enable 0 or 1) clears the stop word and the cycle count.  p2b_mg_result copies (relsq, rsq, residual_error, cycles)
// of the last counted cycle to the host (synchronises the stream) together with the communication error word.
int p2b_mg_set_stop(p2b_mg* m, int enable, double source_norm, double rtol, int max_cycles, void* stream)
{
    P2B_REQUIRE(m && m->base, "hierarchy not bound");
    P2B_LAUNCH(mg_set_stop_kernel, 1, 1, 0, (cudaStream_t)stream)(m->ctl, source_norm, rtol, (double)max_cycles, enable ? 1.0 : 0.0);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

int p2b_mg_result(p2b_mg* m, double* out4, long long* comm_error, void* stream)
{
    P2B_REQUIRE(m && m->base && out4, "null pointer");
    unsigned long long words[CW_WORDS];
    P2B_CUDA_CHECK(cudaMemcpyAsync(words, m->ctl, sizeof words, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    P2B_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    memcpy(out4, words + CW_RESULT, 4 * sizeof(double));
    if (comm_error) *comm_error = (long long)words[CW_ERR];
    return P2B_OK;
}

// device address of the control words (diagnostics; CW_* in mg_kernels.cuh)
void* p2b_mg_control_ptr(p2b_mg* m) { return m ? (void*)m->ctl : nullptr; }

// ---- device memory that other processes can map (the workspaces of a decomposed hierarchy) ---------------------
// p2b_shared_alloc: cudaMalloc'd, zeroed.  p2b_shared_handle: 64 opaque bytes (cudaIpcMemHandle_t) another process on
// the same node passes to p2b_shared_open to map the allocation (peer access is enabled by the open).
void* p2b_shared_alloc(long long bytes)
{
    void* p = nullptr;
    if (bytes <= 0 || cudaMalloc(&p, (size_t)bytes) != cudaSuccess) { set_error("cudaMalloc of %lld bytes failed", bytes); return nullptr; }
    cudaMemset(p, 0, (size_t)bytes);
    cudaDeviceSynchronize();
    return p;
}

int p2b_shared_free(void* p) { P2B_CUDA_CHECK(cudaFree(p)); return P2B_OK; }

int p2b_shared_handle(void* p, unsigned char* out64)
{
    P2B_REQUIRE(p && out64, "null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    P2B_CUDA_CHECK(cudaIpcGetMemHandle(&h, p));
    memcpy(out64, &h, 64);
    return P2B_OK;
}

void* p2b_shared_open(const unsigned char* handle64)
{
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { set_error("cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e)); return nullptr; }
    return p;
}

int p2b_shared_close(void* p) { P2B_CUDA_CHECK(cudaIpcCloseMemHandle(p)); return P2B_OK; }

// ---- variable coefficients: VarCoeffCCMG2d.__init__ (variable_coeff_MG.py:40-109) --------------------
// three planes (eta at cell centres, eta_x, eta_y) per level, each the size of one of the level's planes
long long p2b_mg_coeff_workspace_bytes(p2b_mg* m)
{
    if (!m) return 0;
    long long n = 0;
    for (int l = 0; l < m->nlevels; ++l) n += 3LL * (m->lev[l].n + 2) * m->lev[l].pitch;
    return n * 8;
}

// coeffs: eta on the finest level, (n+2) rows of coeffs_pitch doubles (valid cells are read);
// coeffs_bc: its four boundary types.  Builds, on the device, what the reference's constructor builds:
// the ghost-filled eta of every level (restricted level by level), the finest level's edge
// coefficients and their restrictions.  From here on smooth / residual / V-cycle / diagnostics use
// the variable-coefficient operator (alpha and beta are ignored, as in the reference).
int p2b_mg_set_coeffs(p2b_mg* m, void* mem, long long bytes, const double* coeffs, int coeffs_pitch,
                      const int* coeffs_bc, void* stream)
{
    P2B_REQUIRE(m && m->base, "hierarchy not bound");
    P2B_REQUIRE(mem && coeffs && coeffs_bc, "null pointer");
    P2B_REQUIRE(m->size == 1, "variable coefficients are not available on a decomposed hierarchy");
    P2B_REQUIRE(bytes >= p2b_mg_coeff_workspace_bytes(m), "coefficient workspace too small");
    P2B_REQUIRE(((uintptr_t)mem % 16) == 0, "workspace must be 16-byte aligned");
    P2B_REQUIRE(coeffs_pitch >= m->lev[m->nlevels - 1].n + 2, "coeffs_pitch too small");
    for (int s = 0; s < 4; ++s) P2B_REQUIRE(coeffs_bc[s] >= P2B_BC_OUTFLOW && coeffs_bc[s] <= P2B_BC_PERIODIC, "bad coefficient bc");
    cudaStream_t st = (cudaStream_t)stream;
    P2B_CUDA_CHECK(cudaMemsetAsync(mem, 0, (size_t)p2b_mg_coeff_workspace_bytes(m), st));
    double* p = (double*)mem;
    for (int l = 0; l < m->nlevels; ++l) {
        const long long plane = (long long)(m->lev[l].n + 2) * m->lev[l].pitch;
        m->cc[l] = p; p += plane;
        m->ex[l] = p; p += plane;
        m->ey[l] = p; p += plane;
    }
    MgBC cb;
    cb.xl = coeffs_bc[0]; cb.xr = coeffs_bc[1]; cb.yl = coeffs_bc[2]; cb.yr = coeffs_bc[3];
    cb.xlv = cb.xrv = cb.ylv = cb.yrv = nullptr;
    const int fin = m->nlevels - 1;
    for (int l = fin; l >= 0; --l) {
        MgLevel L = m->lev[l];
        const int n = L.n, P = L.pitch;
        L.v = m->cc[l];                       // the generic kernels below work on plane "v" / "r" / "f"
        if (l == fin) {
            P2B_CUDA_CHECK(cudaMemcpy2DAsync(m->cc[l] + P + 1, (size_t)P * 8, coeffs + coeffs_pitch + 1,
                                             (size_t)coeffs_pitch * 8, (size_t)n * 8, n, cudaMemcpyDeviceToDevice, st));
        } else {
            // coeffs_c.v() = f_patch.restrict("coeffs").v()
            MgLevel F = m->lev[l + 1], Cs = L;
            F.r = m->cc[l + 1];
            Cs.f = m->cc[l];
            dim3 blk(64, 4);
            dim3 grd((n + blk.x - 1) / blk.x, (n + blk.y - 1) / blk.y);
            P2B_LAUNCH(mg_restrict_kernel, grd, blk, 0, st)(F, Cs, 0, comm_none(), 0);
        }
        P2B_LAUNCH(mg_fill_kernel, (4 * n + 255) / 256, 256, 0, st)(L, cb);
        dim3 blk(64, 4);
        dim3 grd((n + 1 + blk.x - 1) / blk.x, (n + 1 + blk.y - 1) / blk.y);
        if (l == fin) {
            P2B_LAUNCH(mg_vc_edges_fine_kernel, grd, blk, 0, st)(m->cc[l], m->ex[l], m->ey[l], n, P,
                                                                 make_div_const(L.dx * L.dx), make_div_const(L.dy * L.dy));
        } else {
            const MgLevel& F = m->lev[l + 1];
            P2B_LAUNCH(mg_vc_edges_restrict_kernel, grd, blk, 0, st)(m->ex[l + 1], m->ey[l + 1], F.pitch, m->ex[l], m->ey[l],
                                                                     n, P, F.dx * F.dx, make_div_const(L.dx * L.dx),
                                                                     F.dy * F.dy, make_div_const(L.dy * L.dy));
        }
    }
    P2B_CUDA_CHECK(cudaGetLastError());
    m->varcoef = 1;
    return P2B_OK;
}

// which: 0 eta (cell centres, ghost-filled), 1 eta_x, 2 eta_y; NULL before p2b_mg_set_coeffs
void* p2b_mg_coeff_ptr(p2b_mg* m, int level, int which)
{
    if (!m || !m->varcoef || level < 0 || level >= m->nlevels) return nullptr;
    return which == 0 ? m->cc[level] : which == 1 ? m->ex[level] : m->ey[level];
}

// change alpha / beta of (alpha - beta L) phi = f on an existing hierarchy (the diffusion solver's beta is
// 0.5*dt*k: the reference builds a new CellCenterMG2d every step, diffusion/simulation.py:76-85)
int p2b_mg_set_operator(p2b_mg* m, double alpha, double beta)
{
    P2B_REQUIRE(m, "null handle");
    P2B_REQUIRE(!m->varcoef, "variable-coefficient hierarchy");
    m->alpha = alpha; m->beta = beta;
    return P2B_OK;
}

// finest-level f <- phi + coef * (5-point Laplacian of phi), phi = a ghost-filled (n+2) x phi_pitch plane:
// the Crank-Nicolson right-hand side of diffusion/simulation.py:87-91 (coef = 0.5*dt*k)
int p2b_mg_cn_rhs(p2b_mg* m, const double* phi, int phi_pitch, double coef, void* stream)
{
    P2B_REQUIRE(m && m->base && phi, "null pointer");
    // on a decomposed hierarchy phi is this rank's slab: L.ni + 2 rows, halo rows already exchanged by the caller
    const MgLevel& L = m->lev[m->nlevels - 1];
    P2B_REQUIRE(phi_pitch >= L.n + 2, "phi_pitch too small");
    dim3 blk(64, 4);
    dim3 grd((L.n + blk.x - 1) / blk.x, (L.ni + blk.y - 1) / blk.y);
    P2B_LAUNCH(mg_cn_rhs_kernel, grd, blk, 0, (cudaStream_t)stream)(L, phi, phi_pitch, coef, make_div_const(L.dx * L.dx),
                                                                    make_div_const(L.dy * L.dy));
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

}  // extern "C"

// sweep_emu.cpp -- host-side emulation of one CUDA warp running pyro2_b200/csrc/sweep_task.cuh.
//
// TEST INFRASTRUCTURE ONLY (built by tests/conftest.py into tests/emu/libsweep_emu.so, loaded only
// by tests/test_sweep_emulated.py).  It compiles the *same* SweepTask source as the device build,
// with the warp services (shuffles, __syncwarp, TMA bulk copy + mbarrier) replaced by 32 host
// threads in lock step.  This lets the kernel's indexing, halo logic and row pipeline be checked
// against the oracle on the CPU-only build box; it is not a fallback and the product never loads it.
#include <pthread.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <vector>

#include "../../pyro2_b200/csrc/sweep_args.cuh"

namespace {

struct WarpShared {
    pthread_barrier_t bar;
    double slot[2][32];
    int flip[32];
};

struct EmuWarp {
    WarpShared* ws;
    int lane_;

    int lane() const { return lane_; }
    void sync() const { pthread_barrier_wait(&ws->bar); }

    // two alternating exchange buffers: one barrier per exchange is enough (see sweep_emu notes)
    double exchange(double v, int from) const
    {
        int b = ws->flip[lane_];
        ws->flip[lane_] ^= 1;
        ws->slot[b][lane_] = v;
        pthread_barrier_wait(&ws->bar);
        return (from < 0 || from > 31) ? v : ws->slot[b][from];
    }
    double up(double v) const { return exchange(v, lane_ - 1); }
    double down(double v) const { return exchange(v, lane_ + 1); }

    void load_issue(unsigned long long&, double* dst, const double* U, long long plane_stride, int pitch, int r, int col0,
                    int ncols) const
    {
        if (lane_ == 0) {
            const double* src = U + (long long)r * pitch + col0;
            for (int n = 0; n < 4; ++n) {
                memcpy(dst + n * pyro::SW_QW, src + n * plane_stride, (size_t)ncols * 8);
                // the tensor copy of the device fills what lies past the end of the row with zeros
                for (int c = ncols; c < pyro::SW_QW; ++c) dst[n * pyro::SW_QW + c] = 0.0;
            }
        }
    }
    void load_wait(unsigned long long&, unsigned) const { pthread_barrier_wait(&ws->bar); }

    double reduce_max(double v) const
    {
        int b = ws->flip[lane_];
        ws->flip[lane_] ^= 1;
        ws->slot[b][lane_] = v;
        pthread_barrier_wait(&ws->bar);
        double m = v;
        for (int l = 0; l < 32; ++l) m = std::max(m, ws->slot[b][l]);
        return m;
    }
    void atomic_max_bits(unsigned long long* addr, double v) const
    {
        unsigned long long bits;
        memcpy(&bits, &v, 8);
        if (bits > *addr) *addr = bits;
    }
};

struct LaneCtx {
    WarpShared* ws;
    pyro::SweepSmem* smem;
    const pyro::SweepArgs* A;
    int lane;
    int ntasks;
    bool grav;
    int riemann;
    bool sph;
};

void* lane_main(void* p)
{
    LaneCtx* c = (LaneCtx*)p;
    EmuWarp w{c->ws, c->lane};
    auto go = [&](auto& T) { for (int t = 0; t < c->ntasks; ++t) T.run(t % c->A->nstrips, t / c->A->nstrips); };
    if (c->sph) {
        pyro::SweepTask<EmuWarp, true, 1, true> T(w, *c->A, *c->smem, 0u); go(T);
    } else if (c->riemann == 2) {
        if (c->grav) { pyro::SweepTask<EmuWarp, true, 2> T(w, *c->A, *c->smem, 0u); go(T); }
        else { pyro::SweepTask<EmuWarp, false, 2> T(w, *c->A, *c->smem, 0u); go(T); }
    } else if (c->riemann == 1) {
        if (c->grav) { pyro::SweepTask<EmuWarp, true, 1> T(w, *c->A, *c->smem, 0u); go(T); }
        else { pyro::SweepTask<EmuWarp, false, 1> T(w, *c->A, *c->smem, 0u); go(T); }
    } else {
        if (c->grav) { pyro::SweepTask<EmuWarp, true, 0> T(w, *c->A, *c->smem, 0u); go(T); }
        else { pyro::SweepTask<EmuWarp, false, 0> T(w, *c->A, *c->smem, 0u); go(T); }
    }
    return nullptr;
}

}  // namespace

namespace {
// one emulated warp works through every task of the launch
int run_tasks(const pyro::SweepArgs& A, bool sources, int riemann, bool sph)
{
    WarpShared ws;
    pthread_barrier_init(&ws.bar, nullptr, 32);
    memset(ws.flip, 0, sizeof ws.flip);
    pyro::SweepSmem* smem = new pyro::SweepSmem;
    // poison the ring so that stale-slot reads show up as NaNs in the comparison
    memset(smem, 0xff, sizeof *smem);
    LaneCtx ctx[32];
    pthread_t th[32];
    for (int l = 0; l < 32; ++l) {
        ctx[l] = LaneCtx{&ws, smem, &A, l, A.nstrips * A.nsegs, sources, riemann, sph};
        pthread_create(&th[l], nullptr, lane_main, &ctx[l]);
    }
    for (int l = 0; l < 32; ++l) pthread_join(th[l], nullptr);
    pthread_barrier_destroy(&ws.bar);
    delete smem;
    return 0;
}
}  // namespace

extern "C" int emu_compressible_sweep(const double* Uin, double* Uout, int nx, int ny, int ng, int pitch,
                                      long long plane_stride, double dx, double dy, double dt,
                                      double gamma, double z0, double z1, double delta, double cvisc,
                                      int limiter, int use_flattening, int no_avisc_xhi, int no_avisc_yhi,
                                      int seglen, uint64_t* scratch, double* dbg, double grav, int src_flip_ylo,
                                      int src_flip_yhi, int riemann, int xl_solid, int yl_solid, const double* heat,
                                      double heat_rate, int do_sponge, double sp_begin, double sp_full, double sp_tau,
                                      int src_copy_yhi, const double* geo_i, const double* geo_j, int geo_ni, int geo_nj,
                                      int src_flip_xlo, int src_flip_xhi)
{
    pyro::SweepArgs A;
    A.Uin = Uin; A.Uout = Uout; A.plane_stride = plane_stride; A.pitch = pitch;
    A.nx = nx; A.ny = ny; A.ng = ng; A.dx = dx; A.dy = dy; A.dt = dt; A.gamma = gamma;
    A.z0 = z0; A.z1 = z1; A.delta = delta; A.cvisc = cvisc;
    A.limiter = limiter; A.use_flattening = use_flattening;
    A.no_avisc_xhi = no_avisc_xhi; A.no_avisc_yhi = no_avisc_yhi;
    A.grav = grav; A.src_flip_ylo = src_flip_ylo; A.src_flip_yhi = src_flip_yhi;
    A.xl_solid = xl_solid; A.yl_solid = yl_solid;
    A.heat = heat; A.heat_rate = heat_rate; A.do_sponge = do_sponge;
    A.sponge_rho_begin = sp_begin; A.sponge_rho_full = sp_full; A.sponge_timescale = sp_tau;
    A.src_copy_yhi = src_copy_yhi;
    A.geo_i = geo_i; A.geo_j = geo_j; A.geo_ni = geo_ni; A.geo_nj = geo_nj;
    A.src_flip_xlo = src_flip_xlo; A.src_flip_xhi = src_flip_xhi;
    A.nstrips = (ny + pyro::SW_OUT - 1) / pyro::SW_OUT;
    A.seglen = seglen;
    A.nsegs = (nx + seglen - 1) / seglen;
    memset(scratch, 0, 4 * sizeof(uint64_t));
    A.wavemax = (unsigned long long*)scratch;
    A.status = (int*)(scratch + 3);
#ifdef SWEEP_DEBUG
    A.dbg = dbg;
#else
    (void)dbg;
#endif
    return run_tasks(A, grav != 0.0 || heat != nullptr || do_sponge != 0, riemann, geo_i != nullptr);
}

// The C ABI of p2b_compressible_sweep over the emulated warp: same argument checks, field copy and work decomposition
// as the device launch (pyro2_b200/csrc/sweep_args.cuh); `resident` stands in for the device's warp slots.
extern "C" int emu_compressible_sweep_abi(const double* Uin, double* Uout, const p2b_grid* g, const p2b_comp_params* prm,
                                          double dt, uint64_t* scratch, int resident, const char** why)
{
    pyro::SweepArgs A;
    const char* e = pyro::sweep_args_from_abi(Uin, Uout, g, prm, dt, scratch, resident, A);
    if (why) *why = e;
    if (e) return -1;
    memset(scratch, 0, 4 * sizeof(uint64_t));
    return run_tasks(A, pyro::sweep_has_sources(prm), prm->riemann, prm->geo_i != nullptr);
}

extern "C" void emu_sweep_decomposition(const p2b_grid* g, int resident, int* ntasks, int* seglen)
{
    const int nstrips = (g->ny + pyro::SW_OUT - 1) / pyro::SW_OUT;
    *seglen = pyro::choose_seglen(g->nx, nstrips, resident);
    *ntasks = nstrips * ((g->nx + *seglen - 1) / *seglen);
}

// the device's branch-free helpers as the emulator restates them (hydro_core.cuh, host branch): same ABI as
// p2b_test_fastmath over host memory
extern "C" int emu_test_fastmath(int op, const double* a, const double* b, double* out, int n)
{
    for (int k = 0; k < n; ++k) {
        if (op == 0) out[k] = pyro::rcp(a[k]);
        else if (op == 1) out[k] = pyro::fdiv(a[k], b[k]);
        else if (op == 2) out[k] = pyro::fsqrt(a[k]);
        else if (op == 4) out[k] = pyro::div_by(a[k], pyro::shared_div(b[k]));
        else if (op == 5) out[k] = a[k] / b[k];
        else {
            const double* l = a + 4 * k; const double* r = b + 4 * k;
            out[k] = pyro::hllc_lm(l[0], l[1], l[2], l[3], r[0], r[1], r[2], r[3], pyro::hllc_par(1.4)).mn;
        }
    }
    return 0;
}

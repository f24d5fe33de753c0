// sweep_args.cuh -- from the C ABI (p2b_grid, p2b_comp_params) to the kernel's SweepArgs: argument checks, the
// field-by-field copy and the work decomposition.  Host code shared by p2b_compressible_sweep (sweep.cu) and the
// host-side emulator entry point (tests/emu/sweep_emu.cpp), so that the CPU tests exercise the same plumbing the
// device launch uses.
#pragma once
#include <stdint.h>

#include "../../include/pyro2b200.h"
#include "sweep_task.cuh"

namespace pyro {

// Uniform tasks scheduled on `resident` warp slots finish in ceil(tasks / resident) rounds, so pick
// the segment length that minimises rounds * (rows per task + per-task overhead).
inline int choose_seglen(int nx, int nstrips, int resident)
{
    const int overhead = 3;   // prologue + the two partial iterations of a segment, in row units
    int best_len = nx, best_cost = 1 << 30;
    for (int k = 1; k <= 24; ++k) {
        long long cap = (long long)k * resident / nstrips;   // segments we can afford in k rounds
        if (cap < 1) continue;
        int len = (int)((nx + cap - 1) / cap);
        if (len < 8) len = 8;
        int nseg = (nx + len - 1) / len;
        long long rounds = ((long long)nseg * nstrips + resident - 1) / resident;
        int cost = (int)(rounds * (len + overhead));
        if (cost <= best_cost) { best_cost = cost; best_len = len; }   // ties -> more, smaller tasks
    }
    return best_len;
}


// any source term selects the instantiations with sources
inline bool sweep_has_sources(const p2b_comp_params* prm)
{
    return prm->grav != 0.0 || prm->heat_profile != nullptr || prm->do_sponge != 0;
}

// returns NULL and fills A, or the reason the arguments are refused; `resident`: warp slots of the device
inline const char* sweep_args_from_abi(const double* Uin, double* Uout, const p2b_grid* g, const p2b_comp_params* prm,
                                       double dt, uint64_t* scratch, int resident, SweepArgs& A)
{
    if (!(Uin && Uout && g && prm && scratch)) return "null pointer";
    if (Uin == Uout) return "the sweep is out of place: Uin == Uout";
    if (g->ng < 4) return "compressible sweep needs ng >= 4";
    // the bulk copies start at column ng + 30 * strip - 4 and need 16-byte aligned sources and byte counts
    if ((g->ng % 2) != 0) return "compressible sweep needs an even ng (16-byte aligned row copies)";
    if (((uintptr_t)Uout % 16) != 0) return "planes must be 16-byte aligned";
    if (g->nx < 1 || g->ny < 1) return "empty grid";
    if (g->pitch < g->ny + 2 * g->ng || (g->pitch % 2) != 0) return "pitch must be even and >= qy";
    if ((g->plane_stride % 2) != 0 || ((uintptr_t)Uin % 16) != 0) return "planes must be 16-byte aligned";
    if (prm->limiter < 0 || prm->limiter > 2) return "limiter must be 0, 1 or 2";
    if (prm->riemann < 0 || prm->riemann > 2) return "riemann must be 0 (HLLC), 1 (CGF) or 2 (HLLC_lm)";
    if (prm->geo_i) {
        if (!prm->geo_j) return "SphericalPolar: geo_j missing";
        if (prm->riemann != 1) return "SphericalPolar geometry needs the CGF Riemann solver";
        if (prm->geo_ni < g->nx + 2 * g->ng || prm->geo_nj < g->ny + 2 * g->ng + 1) return "geometry tables too short";
        if (prm->src_copy_yhi) return "SphericalPolar: the ambient boundary is not supported";
    }
    A.Uin = Uin; A.Uout = Uout;
    A.plane_stride = g->plane_stride; A.pitch = g->pitch;
    A.nx = g->nx; A.ny = g->ny; A.ng = g->ng;
    A.dx = g->dx; A.dy = g->dy; A.dt = dt; A.gamma = prm->gamma;
    A.z0 = prm->z0; A.z1 = prm->z1; A.delta = prm->delta; A.cvisc = prm->cvisc;
    A.limiter = prm->limiter; A.use_flattening = prm->use_flattening;
    A.no_avisc_xhi = prm->no_avisc_xhi; A.no_avisc_yhi = prm->no_avisc_yhi;
    A.grav = prm->grav; A.src_flip_ylo = prm->src_flip_ylo; A.src_flip_yhi = prm->src_flip_yhi;
    A.xl_solid = prm->xl_solid; A.yl_solid = prm->yl_solid;
    A.heat = prm->heat_profile; A.heat_rate = prm->heat_rate;
    A.do_sponge = prm->do_sponge; A.sponge_rho_begin = prm->sponge_rho_begin;
    A.sponge_rho_full = prm->sponge_rho_full; A.sponge_timescale = prm->sponge_timescale;
    A.src_copy_yhi = prm->src_copy_yhi;
    A.geo_i = prm->geo_i; A.geo_j = prm->geo_j; A.geo_ni = prm->geo_ni; A.geo_nj = prm->geo_nj;
    A.src_flip_xlo = prm->src_flip_xlo; A.src_flip_xhi = prm->src_flip_xhi;
    A.nstrips = (g->ny + SW_OUT - 1) / SW_OUT;
    A.seglen = choose_seglen(g->nx, A.nstrips, resident);
    A.nsegs = (g->nx + A.seglen - 1) / A.seglen;
    A.wavemax = (unsigned long long*)scratch;
    A.status = (int*)(scratch + 3);
#ifdef SWEEP_DEBUG
    A.dbg = nullptr;
#endif
    return nullptr;
}

}  // namespace pyro

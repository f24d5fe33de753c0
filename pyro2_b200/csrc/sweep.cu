// sweep.cu -- device policy + launcher of the fused compressible sweep (see sweep_task.cuh).
//
// One persistent kernel; each warp pulls (strip, segment) tasks from an atomic counter.  Rows of
// the conserved state are staged into the warp's shared-memory ring by the TMA engine
// (cp.async.bulk 1-d copies completing on an mbarrier: UBLKCP in SASS); no block-wide barrier is
// ever executed after the mbarrier initialisation.
#include "common.cuh"
#include "sweep_task.cuh"

namespace pyro {

constexpr int SWEEP_WARPS = 4;                 // warps per CTA (independent of each other)
constexpr int SWEEP_THREADS = 32 * SWEEP_WARPS;
#ifndef SWEEP_MIN_BLOCKS
#define SWEEP_MIN_BLOCKS 3                     // 12 warps/SM -> <= 168 registers per thread
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

struct CudaWarp {
    __device__ __forceinline__ int lane() const { return threadIdx.x & 31; }
    __device__ __forceinline__ double up(double v) const { return __shfl_up_sync(0xffffffffu, v, 1); }
    __device__ __forceinline__ double down(double v) const { return __shfl_down_sync(0xffffffffu, v, 1); }
    __device__ __forceinline__ void sync() const { __syncwarp(); }

    // lane 0 arms the slot's mbarrier with the byte count and issues one bulk copy per variable
    __device__ __forceinline__ void load_issue(unsigned long long& mbar, double* d0, double* d1, double* d2,
                                               double* d3, const double* src, long long plane_stride,
                                               int ncols) const
    {
        if (lane() == 0) {
            const uint32_t bytes = (uint32_t)ncols * 8u;
            const uint32_t bar = smem_u32(&mbar);
            // the slot was last written through the generic proxy (cons->prim in place); order
            // those writes before the async-proxy writes of the bulk copies
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes * 4u)
                         : "memory");
            double* dst[4] = {d0, d1, d2, d3};
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                asm volatile(
                    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                        smem_u32(dst[n])),
                    "l"(src + n * plane_stride), "r"(bytes), "r"(bar)
                    : "memory");
            }
        }
    }

    __device__ __forceinline__ void load_wait(unsigned long long& mbar, unsigned parity) const
    {
        const uint32_t bar = smem_u32(&mbar);
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(bar), "r"(parity)
                : "memory");
        }
    }

    __device__ __forceinline__ double reduce_max(double v) const
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = dmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        return v;
    }

    __device__ __forceinline__ void atomic_max_bits(unsigned long long* addr, double v) const
    {
        atomicMax(addr, (unsigned long long)__double_as_longlong(v));
    }
};

template <bool GRAV, int RIEMANN, bool SPH = false>
__global__ void __launch_bounds__(SWEEP_THREADS, SWEEP_MIN_BLOCKS)
sweep_kernel(SweepArgs A, unsigned long long* task_counter, int ntasks)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SweepSmem& S = reinterpret_cast<SweepSmem*>(smem_raw)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < SW_RING; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&S.mbar[s])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    CudaWarp w;
    SweepTask<CudaWarp, GRAV, RIEMANN, SPH> T(w, A, S, 0u);
    for (;;) {
        int t = 0;
        if (lane == 0) t = (int)atomicAdd(task_counter, 1ull);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= ntasks) break;
        // strips vary fastest: concurrently running warps work on neighbouring columns of the
        // same rows, so the 8-column halo overlap is served by L2
        T.run(t % A.nstrips, t / A.nstrips);
    }
}

static int g_last_ntasks = 0, g_last_resident = 0, g_last_seglen = 0;

static int resident_warps()
{
    static int resident = 0;
    if (!resident) {
        int blocks = 0;
        const int smem = (int)(SWEEP_WARPS * sizeof(SweepSmem));
        cudaFuncSetAttribute(sweep_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(sweep_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(sweep_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(sweep_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(sweep_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(sweep_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(sweep_kernel<true, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        // every instantiation has the same launch bounds and shared memory, hence the same residency
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, sweep_kernel<false, 0>, SWEEP_THREADS,
                                                      SWEEP_WARPS * sizeof(SweepSmem));
        if (blocks < 1) blocks = 1;
        resident = blocks * SWEEP_WARPS * num_sms();
    }
    return resident;
}

// Uniform tasks scheduled on `resident` warp slots finish in ceil(tasks / resident) rounds, so pick
// the segment length that minimises rounds * (rows per task + per-task overhead).
static int choose_seglen(int nx, int nstrips, int resident)
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

}  // namespace pyro

using namespace pyro;

extern "C" {

int p2b_compressible_sweep(const double* Uin, double* Uout, const p2b_grid* g, const p2b_comp_params* prm,
                           double dt, uint64_t* scratch, void* stream)
{
    P2B_REQUIRE(Uin && Uout && g && prm && scratch, "null pointer");
    P2B_REQUIRE(Uin != Uout, "the sweep is out of place: Uin == Uout");
    P2B_REQUIRE(g->ng >= 4, "compressible sweep needs ng >= 4");
    P2B_REQUIRE(g->nx >= 1 && g->ny >= 1, "empty grid");
    P2B_REQUIRE(g->pitch >= g->ny + 2 * g->ng && (g->pitch % 2) == 0, "pitch must be even and >= qy");
    P2B_REQUIRE((g->plane_stride % 2) == 0 && ((uintptr_t)Uin % 16) == 0, "planes must be 16-byte aligned");
    P2B_REQUIRE(prm->limiter >= 0 && prm->limiter <= 2, "limiter must be 0, 1 or 2");
    P2B_REQUIRE(prm->riemann >= 0 && prm->riemann <= 2, "riemann must be 0 (HLLC), 1 (CGF) or 2 (HLLC_lm)");
    cudaStream_t st = (cudaStream_t)stream;

    SweepArgs A;
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
    if (prm->geo_i) {
        P2B_REQUIRE(prm->geo_j, "SphericalPolar: geo_j missing");
        P2B_REQUIRE(prm->riemann == 1, "SphericalPolar geometry needs the CGF Riemann solver");
        P2B_REQUIRE(prm->geo_ni >= g->nx + 2 * g->ng && prm->geo_nj >= g->ny + 2 * g->ng + 1, "geometry tables too short");
        P2B_REQUIRE(!prm->heat_profile && !prm->do_sponge && !prm->src_copy_yhi,
                    "SphericalPolar: heating, sponge and ambient boundaries are not supported");
    }
    A.nstrips = (g->ny + SW_OUT - 1) / SW_OUT;
    const int resident = resident_warps();
    A.seglen = choose_seglen(g->nx, A.nstrips, resident);
    A.nsegs = (g->nx + A.seglen - 1) / A.seglen;
    A.wavemax = (unsigned long long*)scratch;
    A.status = (int*)(scratch + 3);
    const int ntasks = A.nstrips * A.nsegs;
    g_last_ntasks = ntasks; g_last_resident = resident; g_last_seglen = A.seglen;

    P2B_CUDA_CHECK(cudaMemsetAsync(scratch, 0, 4 * sizeof(uint64_t), st));
    int blocks = (ntasks + SWEEP_WARPS - 1) / SWEEP_WARPS;
    const int maxblocks = resident / SWEEP_WARPS;
    if (blocks > maxblocks) blocks = maxblocks;
    const size_t smem = SWEEP_WARPS * sizeof(SweepSmem);
    unsigned long long* counter = (unsigned long long*)(scratch + 2);
    const bool grav = prm->grav != 0.0 || prm->heat_profile != nullptr || prm->do_sponge != 0;   // any source term
    if (prm->geo_i) {
        sweep_kernel<true, 1, true><<<blocks, SWEEP_THREADS, smem, st>>>(A, counter, ntasks);
    } else if (prm->riemann == 2) {
        if (grav) sweep_kernel<true, 2><<<blocks, SWEEP_THREADS, smem, st>>>(A, counter, ntasks);
        else sweep_kernel<false, 2><<<blocks, SWEEP_THREADS, smem, st>>>(A, counter, ntasks);
    } else if (prm->riemann == 1) {
        if (grav) sweep_kernel<true, 1><<<blocks, SWEEP_THREADS, smem, st>>>(A, counter, ntasks);
        else sweep_kernel<false, 1><<<blocks, SWEEP_THREADS, smem, st>>>(A, counter, ntasks);
    } else {
        if (grav) sweep_kernel<true, 0><<<blocks, SWEEP_THREADS, smem, st>>>(A, counter, ntasks);
        else sweep_kernel<false, 0><<<blocks, SWEEP_THREADS, smem, st>>>(A, counter, ntasks);
    }
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

int p2b_sweep_info(int* ntasks, int* resident_warps_out, int* seglen)
{
    if (ntasks) *ntasks = g_last_ntasks;
    if (resident_warps_out) *resident_warps_out = g_last_resident;
    if (seglen) *seglen = g_last_seglen;
    return P2B_OK;
}

}  // extern "C"

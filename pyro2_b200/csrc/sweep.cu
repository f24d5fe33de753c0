// sweep.cu -- device policy + launcher of the fused compressible sweep (see sweep_task.cuh).
//
// One persistent kernel; each warp pulls (strip, segment) tasks from an atomic counter.  Rows of
// the conserved state are staged into the warp's shared-memory ring by the TMA engine
// (cp.async.bulk 1-d copies completing on an mbarrier: UBLKCP in SASS); no block-wide barrier is
// ever executed after the mbarrier initialisation.
#include <cuda.h>            // CUtensorMap and the encode entry point's types only: no libcuda at link time

#include "common.cuh"
#include "sweep_args.cuh"

namespace pyro {

// One warp per CTA, 12 CTAs per SM.  Warps never talk to each other, so the CTA size is free: with a single warp the
// warp's shared-memory block sits at a compile-time address (no per-access base arithmetic) -- measured 2.56 -> 2.32 ms
// per 4096^2 sweep against 4-warp CTAs at the same 168 registers and 12 warps per SM (profiles/r2_sweep_ab.txt).  Eight
// warps without spills (242 registers) and sixteen with more (128) are both slower.
#ifndef SWEEP_WARPS_CFG
#define SWEEP_WARPS_CFG 1
#endif
constexpr int SWEEP_WARPS = SWEEP_WARPS_CFG;   // warps per CTA (independent of each other)
constexpr int SWEEP_THREADS = 32 * SWEEP_WARPS;
#ifndef SWEEP_MIN_BLOCKS
#define SWEEP_MIN_BLOCKS 11                    // register cap 186; ptxas settles on 168 -> 12 warps / SM
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

    // lane 0 arms the slot's mbarrier with the byte count and issues ONE 3-d tensor copy (38 columns x 1 row x 4 planes,
    // UTMALDG in SASS) through the launch's tensor map; columns past the end of the row arrive as zeros.  Without a tensor
    // map (P2B_SWEEP_NO_TMAP builds / driver without cuTensorMapEncodeTiled): one 1-d bulk copy per plane (UBLKCP).
    const CUtensorMap* tmap;
    __device__ __forceinline__ void load_issue(unsigned long long& mbar, double* dst, const double* U, long long plane_stride,
                                               int pitch, int r, int col0, int ncols) const
    {
        if (lane() == 0) {
            const uint32_t bar = smem_u32(&mbar);
            // the slot was last written through the generic proxy (cons->prim in place); order
            // those writes before the async-proxy writes of the copy
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (tmap) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(4u * SW_QW * 8u) : "memory");
                asm volatile(
                    "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                    ::"r"(smem_u32(dst)), "l"(tmap), "r"(col0), "r"(r), "r"(0), "r"(bar)
                    : "memory");
            } else {
                const uint32_t bytes = (uint32_t)ncols * 8u;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes * 4u) : "memory");
                const double* src = U + (long long)r * pitch + col0;
#pragma unroll
                for (int n = 0; n < 4; ++n) {
                    asm volatile(
                        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                            smem_u32(dst + n * SW_QW)),
                        "l"(src + n * plane_stride), "r"(bytes), "r"(bar)
                        : "memory");
                }
            }
        }
    }

    // bounded: a copy that never lands (a bad tensor map, a protocol error) must not hang the device -- after ~1 s the warp
    // gives up, raises the status word (2) and carries on with whatever the slot holds
    int* status;
    __device__ __forceinline__ void load_wait(unsigned long long& mbar, unsigned parity) const
    {
        const uint32_t bar = smem_u32(&mbar);
        uint32_t done = 0;
        for (unsigned spins = 0; !done; ++spins) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(bar), "r"(parity)
                : "memory");
            if (spins > (1u << 22)) { if (status) *status = 2; break; }
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
sweep_kernel(SweepArgs A, unsigned long long* task_counter, int ntasks, const __grid_constant__ CUtensorMap tmap, int use_tmap)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
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
    w.tmap = use_tmap ? &tmap : nullptr;
    w.status = A.status;
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

__global__ void fastmath_probe_kernel(int op, const double* a, const double* b, double* out, int n)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    if (op == 0) out[k] = rcp(a[k]);
    else if (op == 1) out[k] = fdiv(a[k], b[k]);
    else if (op == 2) out[k] = fsqrt(a[k]);
    else if (op == 4) out[k] = div_by(a[k], shared_div(b[k]));
    else if (op == 5) out[k] = __ddiv_rn(a[k], b[k]);
    else {
        const double* l = a + 4 * k; const double* r = b + 4 * k;
        out[k] = hllc_lm(l[0], l[1], l[2], l[3], r[0], r[1], r[2], r[3], hllc_par(1.4)).mn;
    }
}

// The tensor map of one state buffer: a rank-3 fp64 tensor (columns, rows, 4 planes) with the planes' row pitch and plane
// stride, box = 38 columns x 1 row x 4 planes, out-of-range columns filled with zeros.  cuTensorMapEncodeTiled comes from
// the driver through the runtime (cudaGetDriverEntryPoint): libcuda is not a link-time dependency.  Encoding is host-only
// arithmetic (~1 us), done per launch because the buffers alternate.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
#ifndef P2B_SWEEP_NO_TMAP
        if (!getenv("P2B_SWEEP_NO_TMAP")) {
            void* p = nullptr;
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
                q == cudaDriverEntryPointSuccess)
                fn = (EncodeTiledFn)p;
            cudaGetLastError();
        }
#endif
    }
    return fn;
}

static bool make_tensor_map(CUtensorMap* tm, const SweepArgs& A)
{
    EncodeTiledFn enc = encode_tiled();
    memset(tm, 0, sizeof *tm);
    if (!enc) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)A.pitch, (cuuint64_t)(A.nx + 2 * A.ng), 4};
    const cuuint64_t strides[2] = {(cuuint64_t)A.pitch * 8, (cuuint64_t)A.plane_stride * 8};     // bytes, dims 1 and 2
    const cuuint32_t box[3] = {SW_QW, 1, 4};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (strides[0] % 16 || strides[1] % 16) return false;
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(A.Uin), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int g_last_ntasks = 0, g_last_tmap = 0, g_last_resident = 0, g_last_seglen = 0;

// function attributes and SM counts are per device: cached per device ordinal
static int resident_warps()
{
    static int resident[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!resident[dev]) {
        int blocks = 0, sms = 0;
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
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
        resident[dev] = blocks * SWEEP_WARPS * sms;
    }
    return resident[dev];
}

}  // namespace pyro

using namespace pyro;

extern "C" {

int p2b_compressible_sweep(const double* Uin, double* Uout, const p2b_grid* g, const p2b_comp_params* prm,
                           double dt, uint64_t* scratch, void* stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    SweepArgs A;
    const int resident = resident_warps();
    if (const char* why = sweep_args_from_abi(Uin, Uout, g, prm, dt, scratch, resident, A)) {
        pyro::set_error("invalid argument: %s (p2b_compressible_sweep)", why);
        return P2B_EINVAL;
    }
    const int ntasks = A.nstrips * A.nsegs;
    g_last_ntasks = ntasks; g_last_resident = resident; g_last_seglen = A.seglen;

    P2B_CUDA_CHECK(cudaMemsetAsync(scratch, 0, 4 * sizeof(uint64_t), st));
    int blocks = (ntasks + SWEEP_WARPS - 1) / SWEEP_WARPS;
    const int maxblocks = resident / SWEEP_WARPS;
    if (blocks > maxblocks) blocks = maxblocks;
    const size_t smem = SWEEP_WARPS * sizeof(SweepSmem);
    unsigned long long* counter = (unsigned long long*)(scratch + 2);
    const bool grav = sweep_has_sources(prm);
    alignas(64) CUtensorMap tm;
    const int tmap = make_tensor_map(&tm, A) ? 1 : 0;
    g_last_tmap = tmap;
    if (prm->geo_i) {
        sweep_kernel<true, 1, true><<<blocks, SWEEP_THREADS, smem, st>>>(A, counter, ntasks, tm, tmap);
    } else if (prm->riemann == 2) {
        if (grav) sweep_kernel<true, 2><<<blocks, SWEEP_THREADS, smem, st>>>(A, counter, ntasks, tm, tmap);
        else sweep_kernel<false, 2><<<blocks, SWEEP_THREADS, smem, st>>>(A, counter, ntasks, tm, tmap);
    } else if (prm->riemann == 1) {
        if (grav) sweep_kernel<true, 1><<<blocks, SWEEP_THREADS, smem, st>>>(A, counter, ntasks, tm, tmap);
        else sweep_kernel<false, 1><<<blocks, SWEEP_THREADS, smem, st>>>(A, counter, ntasks, tm, tmap);
    } else {
        if (grav) sweep_kernel<true, 0><<<blocks, SWEEP_THREADS, smem, st>>>(A, counter, ntasks, tm, tmap);
        else sweep_kernel<false, 0><<<blocks, SWEEP_THREADS, smem, st>>>(A, counter, ntasks, tm, tmap);
    }
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

int p2b_test_fastmath(int op, const double* a, const double* b, double* out, int n, void* stream)
{
    P2B_REQUIRE(op >= 0 && op <= 5 && a && out && n >= 0 && (b || op == 0 || op == 2), "bad probe arguments");
    if (n == 0) return P2B_OK;
    fastmath_probe_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(op, a, b, out, n);
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

// 1: the last sweep staged its rows with 3-d tensor-map copies (UTMALDG), 0: with 1-d bulk copies (UBLKCP)
int p2b_sweep_uses_tensor_map(void) { return g_last_tmap; }

int p2b_sweep_info(int* ntasks, int* resident_warps_out, int* seglen)
{
    if (ntasks) *ntasks = g_last_ntasks;
    if (resident_warps_out) *resident_warps_out = g_last_resident;
    if (seglen) *seglen = g_last_seglen;
    return P2B_OK;
}

}  // extern "C"

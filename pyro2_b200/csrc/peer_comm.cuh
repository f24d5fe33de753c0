// peer_comm.cuh -- device primitives of the peer-memory communication between the x-slabs of a decomposed run
// (multigrid: mg_kernels.cuh "peer-memory communication"; state planes of the explicit solvers: ghost_cfl.cu p2b_slab_*).
// System-scope acquire loads / release stores of 64-bit flag words that live in the CONSUMER's memory and are written by
// the producer over NVLink, a bounded spin, and the time-out error word.  Under tests/emu/cuda_emu.h the ranks are host
// threads or processes sharing memory and the same operations are GCC atomics.
#pragma once

namespace pyro {

constexpr int COMM_ERR_WORD = 1;     // control word every communicator keeps at index 1: non-zero = a wait timed out

#ifndef MG_COMM_TIMEOUT_NS
#define MG_COMM_TIMEOUT_NS 4000000000LL
#endif

#ifdef P2B_EMU_HEADER
__device__ __forceinline__ unsigned long long comm_load(const unsigned long long* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
__device__ __forceinline__ void comm_store(unsigned long long* p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
__device__ __forceinline__ void comm_store_relaxed(unsigned long long* p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELAXED); }
__device__ __forceinline__ long long comm_clock_ns() { return emu_clock_ns(); }
__device__ __forceinline__ void comm_pause() { emu_pause(); }
__device__ __forceinline__ int comm_tid() { return emu::lin_tid; }
#else
__device__ __forceinline__ unsigned long long comm_load(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void comm_store(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// after an explicit system-scope fence: a flag store that need not drain the write queue again
__device__ __forceinline__ void comm_store_relaxed(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ long long comm_clock_ns()
{
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void comm_pause() { __nanosleep(20); }
__device__ __forceinline__ int comm_tid() { return threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z); }
#endif

// one thread: spin until ctl[word] >= target; a time-out raises the error word and gives up (never a hang)
__device__ __forceinline__ void comm_wait_ge(unsigned long long* ctl, int word, unsigned long long target)
{
    if (comm_load(ctl + word) >= target) return;
    const long long t0 = comm_clock_ns();
    while (comm_load(ctl + word) < target) {
        comm_pause();
        if (comm_clock_ns() - t0 > MG_COMM_TIMEOUT_NS) { atomicExch(ctl + COMM_ERR_WORD, 1ull + (unsigned long long)word); break; }
    }
}


}  // namespace pyro

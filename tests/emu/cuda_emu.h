// cuda_emu.h -- a small host emulation of the CUDA execution model.  TEST INFRASTRUCTURE ONLY.
//
// tests/emu/mg_emu.cpp compiles pyro2_b200/csrc/mg.cu (kernels AND host orchestration, unchanged) with
// g++ and -DP2B_EMU_HEADER=<this file>: common.cuh then includes this header instead of
// <cuda_runtime.h>.  "Device" memory is host memory; a launch runs the CTAs one after the other;
// kernels that synchronise (registered with emu::threaded) get one FIBER (ucontext) per CUDA thread,
// scheduled round-robin inside the calling OS thread: __syncthreads and the warp shuffles are yield
// points that a fiber leaves only when its whole CTA / warp has arrived; every other kernel simply
// runs its threads one after the other.  The point is to check indexing, halo / ghost logic and bit-exactness of
// the multigrid kernels against the oracle on the GPU-less build box.  The product never loads it.
#pragma once
#include <math.h>
#include <ucontext.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <functional>
#include <set>
#include <vector>

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct emu_uint3 { unsigned x, y, z; };
struct double2 { double x, y; };
inline double2 make_double2(double x, double y) { return double2{x, y}; }

// of the running fiber (saved / restored by the scheduler).  Everything the scheduler keeps is thread_local: several
// OS threads may each run an emulated "GPU" at the same time (the ranks of a decomposed multigrid talk to each other
// through plain memory, tests/test_mg_emulated.py), each with its own fibers.
extern thread_local emu_uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;

#define __global__
#define __device__
#define __host__
#define __grid_constant__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static thread_local   // CTAs of one emulated GPU run one at a time: one copy per OS thread is the CTA's copy
#define __align__(n) __attribute__((aligned(n)))

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaDevAttrMultiProcessorCount = 16 };
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 148; return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, int) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
// cudaMalloc is only used for allocations that other ranks map (p2b_shared_alloc): POSIX shared memory, so that the
// ranks may be threads of one process or separate processes (the gloo tests).  The "IPC handle" is the segment's name.
#include <cerrno>
#include <cstdint>
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>
#include <map>
#include <mutex>
#include <string>
namespace emu {
struct Shm { std::string name; size_t bytes; bool owner; };
inline std::map<void*, Shm>& shm_table() { static std::map<void*, Shm> t; return t; }
inline std::mutex& shm_mutex() { static std::mutex m; return m; }
// segments still allocated when the process ends are unlinked (a test that fails, or a Python exit that never freed its
// handles, used to leave them in /dev/shm -- and a later process with the same recycled pid then failed to allocate)
struct ShmJanitor {
    ~ShmJanitor() { for (auto& kv : shm_table()) if (kv.second.owner) shm_unlink(kv.second.name.c_str()); }
};
inline ShmJanitor& shm_janitor() { static ShmJanitor j; return j; }
}  // namespace emu
inline cudaError_t cudaMalloc(void** p, size_t n)
{
    static int counter = 0;
    std::lock_guard<std::mutex> lk(emu::shm_mutex());
    emu::shm_table();
    emu::shm_janitor();               // constructed after the table: destroyed before it
    char name[64];
    // pid + this library's own counter address + counter: unique among live segments, so a segment that already has
    // the name was left behind by a dead process whose pid has been recycled -- remove it and retry
    snprintf(name, sizeof name, "/p2b_emu_%d_%lx_%d", (int)getpid(), (unsigned long)(reinterpret_cast<uintptr_t>(&counter) >> 4 & 0xffffff), counter++);
    int fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0 && errno == EEXIST) {
        shm_unlink(name);
        fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
    }
    if (fd < 0) return 1;
    if (ftruncate(fd, (off_t)n) != 0) { close(fd); shm_unlink(name); return 1; }
    void* q = mmap(nullptr, n, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (q == MAP_FAILED) { shm_unlink(name); return 1; }
    emu::shm_table()[q] = emu::Shm{name, n, true};
    *p = q;
    return cudaSuccess;
}
inline cudaError_t cudaFree(void* p)
{
    std::lock_guard<std::mutex> lk(emu::shm_mutex());
    auto it = emu::shm_table().find(p);
    if (it == emu::shm_table().end()) return 1;
    munmap(p, it->second.bytes);
    if (it->second.owner) shm_unlink(it->second.name.c_str());
    emu::shm_table().erase(it);
    return cudaSuccess;
}
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p)
{
    std::lock_guard<std::mutex> lk(emu::shm_mutex());
    auto it = emu::shm_table().find(p);
    if (it == emu::shm_table().end()) return 1;
    memset(h, 0, sizeof *h);
    snprintf(h->reserved, 48, "%s", it->second.name.c_str());
    memcpy(h->reserved + 48, &it->second.bytes, sizeof(size_t));
    return cudaSuccess;
}
inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned)
{
    std::lock_guard<std::mutex> lk(emu::shm_mutex());
    const std::string name(h.reserved);
    for (auto& kv : emu::shm_table())
        if (kv.second.name == name && kv.second.owner) { *p = kv.first; return cudaSuccess; }   // a rank of this very process
    size_t n = 0;
    memcpy(&n, h.reserved + 48, sizeof n);
    int fd = shm_open(name.c_str(), O_RDWR, 0600);
    if (fd < 0) return 1;
    void* q = mmap(nullptr, n, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (q == MAP_FAILED) return 1;
    emu::shm_table()[q] = emu::Shm{name, n, false};
    *p = q;
    return cudaSuccess;
}
inline cudaError_t cudaIpcCloseMemHandle(void* p)
{
    std::lock_guard<std::mutex> lk(emu::shm_mutex());
    auto it = emu::shm_table().find(p);
    if (it == emu::shm_table().end() || it->second.owner) return cudaSuccess;
    munmap(p, it->second.bytes);
    emu::shm_table().erase(it);
    return cudaSuccess;
}
inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, int, cudaStream_t)
{
    for (size_t r = 0; r < h; ++r) memmove((char*)d + r * dp, (const char*)s + r * sp, w);
    return cudaSuccess;
}
template <class K> inline cudaError_t cudaFuncSetAttribute(K, int, int) { return cudaSuccess; }

inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
inline int __ffs(int x) { return __builtin_ffs(x); }

namespace emu {

struct Cta {
    int nthreads = 0;
    int live = 0;                      // fibers that have not returned yet
    int all_count = 0;                 // arrivals at the current __syncthreads
    unsigned all_gen = 0;
    std::vector<int> warp_count;       // arrivals at the current exchange, per warp
    std::vector<unsigned> warp_gen;
    std::vector<double> slot;          // [warp][2][32]
    std::vector<int> flip;             // exchange-buffer parity per thread
    int bar_count[16] = {0};           // named barriers (bar.sync id, nthreads)
    unsigned bar_gen[16] = {0};
};

extern thread_local Cta* cta;          // null in sequential mode
extern thread_local int lin_tid;       // linear thread index of the running fiber
extern thread_local void* dyn_smem;    // dynamic shared memory of the running CTA
extern thread_local bool force_threaded;   // run every kernel with fibers (set by host code whose launches may synchronise)
void yield();                          // switch to the next fiber of the CTA

std::set<const void*>& threaded_set();
inline void threaded(const void* k) { threaded_set().insert(k); }

void run(const void* kernel, dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);

[[noreturn]] inline void need_threads(const char* what)
{
    fprintf(stderr, "cuda_emu: %s used by a kernel that was not registered with emu::threaded()\n", what);
    abort();
}

// warp-wide exchange: every lane deposits its value, waits until the whole warp has, then reads.  Two
// alternating buffers: a lane can be at most one exchange ahead of the slowest lane of its warp.
inline double exchange(double v, int from_lane)
{
    if (!cta) need_threads("warp shuffle");
    Cta* c = cta;
    const int tid = lin_tid, w = tid >> 5, lane = tid & 31;
    const int width = (c->nthreads - w * 32) < 32 ? (c->nthreads - w * 32) : 32;
    const int b = c->flip[tid];
    c->flip[tid] ^= 1;
    double* s = &c->slot[(size_t)(w * 2 + b) * 32];
    s[lane] = v;
    const unsigned gen = c->warp_gen[w];
    if (++c->warp_count[w] == width) { c->warp_count[w] = 0; ++c->warp_gen[w]; }
    while (c->warp_gen[w] == gen) yield();
    return (from_lane < 0 || from_lane >= width) ? v : s[from_lane];
}

template <class... P>
struct Bound {
    void (*k)(P...);
    dim3 g, b;
    size_t s;
    template <class... A>
    void operator()(A&&... a) const
    {
        auto kk = k;
        run((const void*)k, g, b, s, [&] { kk(a...); });
    }
};

template <class... P>
Bound<P...> bind_launch(void (*k)(P...), dim3 g, dim3 b, size_t s) { return Bound<P...>{k, g, b, s}; }

}  // namespace emu

inline void __syncthreads()
{
    emu::Cta* c = emu::cta;
    if (!c) emu::need_threads("__syncthreads");
    const unsigned gen = c->all_gen;
    if (++c->all_count == c->live) { c->all_count = 0; ++c->all_gen; }
    while (c->all_gen == gen) emu::yield();
}
inline void emu_bar_sync(int id, int nthreads)
{
    emu::Cta* c = emu::cta;
    if (!c) emu::need_threads("bar.sync");
    const unsigned gen = c->bar_gen[id];
    if (++c->bar_count[id] == nthreads) { c->bar_count[id] = 0; ++c->bar_gen[id]; }
    while (c->bar_gen[id] == gen) emu::yield();
}
inline void __syncwarp() { (void)emu::exchange(0.0, -1); }      // a warp-wide rendezvous
inline double __shfl_up_sync(unsigned, double v, int d) { return emu::exchange(v, (emu::lin_tid & 31) - d); }
inline double __shfl_down_sync(unsigned, double v, int d) { return emu::exchange(v, (emu::lin_tid & 31) + d); }
inline double __shfl_xor_sync(unsigned, double v, int m) { return emu::exchange(v, (emu::lin_tid & 31) ^ m); }
inline long long __double_as_longlong(double x) { long long b; memcpy(&b, &x, 8); return b; }
inline unsigned long long atomicMax(unsigned long long* a, unsigned long long v)
{
    const unsigned long long old = *a;       // fibers never run concurrently
    if (v > old) *a = v;
    return old;
}
// cross-"GPU" primitives (ranks are OS threads sharing the address space): real atomics and fences
inline unsigned long long atomicAdd(unsigned long long* a, unsigned long long v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicExch(unsigned long long* a, unsigned long long v) { return __atomic_exchange_n(a, v, __ATOMIC_SEQ_CST); }
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

#include <sched.h>
#include <time.h>
inline long long emu_clock_ns() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (long long)t.tv_sec * 1000000000LL + t.tv_nsec; }
inline void emu_pause() { sched_yield(); }

#define P2B_LAUNCH(kernel, grid, block, smem, stream) ::emu::bind_launch(kernel, grid, block, smem)
#define P2B_DYN_SMEM(type, name) type* name = (type*)::emu::dyn_smem

// mg_emu.cpp -- pyro2_b200/csrc/mg.cu (kernels and host orchestration, unchanged) compiled for the
// host through cuda_emu.h.  TEST INFRASTRUCTURE ONLY: built by tests/test_mg_emulated.py into
// tests/emu/libmg_emu.so, which exports the same p2b_mg_* C ABI over HOST memory.  Lets the multigrid
// kernels be checked bit-for-bit against the oracle without a GPU; not a fallback -- the product
// (pyro2_b200/_lib.py) only ever loads the nvcc-built libpyro2b200.so.
#include <stdarg.h>

#include "cuda_emu.h"

emu_uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;

namespace emu {

Cta* cta = nullptr;
int lin_tid = 0;
void* dyn_smem = nullptr;

std::set<const void*>& threaded_set()
{
    static std::set<const void*> s;
    return s;
}

namespace {

constexpr size_t STACK_BYTES = 256 * 1024;

struct Fiber {
    ucontext_t ctx;
    bool done;
};

std::vector<Fiber> fibers;
std::vector<char*> stacks;
ucontext_t sched_ctx;
int cur = 0;
const std::function<void()>* cur_body = nullptr;

void fiber_entry()
{
    (*cur_body)();
    fibers[cur].done = true;      // returning switches to uc_link = the scheduler
}

void set_thread(const dim3& block, int tid)
{
    lin_tid = tid;
    threadIdx.x = tid % block.x;
    threadIdx.y = (tid / block.x) % block.y;
    threadIdx.z = tid / (block.x * block.y);
}

void set_block(const dim3& grid, const dim3& block, unsigned blk)
{
    gridDim = grid;
    blockDim = block;
    blockIdx.x = blk % grid.x;
    blockIdx.y = (blk / grid.x) % grid.y;
    blockIdx.z = blk / (grid.x * grid.y);
}

}  // namespace

void yield() { swapcontext(&fibers[cur].ctx, &sched_ctx); }

void run(const void* kernel, dim3 grid, dim3 block, size_t smem, const std::function<void()>& body)
{
    const int T = (int)(block.x * block.y * block.z);
    const unsigned nblk = grid.x * grid.y * grid.z;
    std::vector<char> dyn(smem + 64);
    dyn_smem = (void*)(((uintptr_t)dyn.data() + 63) & ~(uintptr_t)63);
    if (!threaded_set().count(kernel)) {
        for (unsigned blk = 0; blk < nblk; ++blk) {
            set_block(grid, block, blk);
            for (int t = 0; t < T; ++t) {
                set_thread(block, t);
                body();
            }
        }
        dyn_smem = nullptr;
        return;
    }
    if ((int)fibers.size() < T) fibers.resize(T);
    while ((int)stacks.size() < T) stacks.push_back((char*)malloc(STACK_BYTES));
    cur_body = &body;
    const int nw = (T + 31) / 32;
    for (unsigned blk = 0; blk < nblk; ++blk) {
        set_block(grid, block, blk);
        Cta c;
        c.nthreads = c.live = T;
        c.warp_count.assign(nw, 0);
        c.warp_gen.assign(nw, 0);
        c.slot.assign((size_t)nw * 2 * 32, 0.0);
        c.flip.assign(T, 0);
        cta = &c;
        for (int t = 0; t < T; ++t) {
            getcontext(&fibers[t].ctx);
            fibers[t].ctx.uc_stack.ss_sp = stacks[t];
            fibers[t].ctx.uc_stack.ss_size = STACK_BYTES;
            fibers[t].ctx.uc_link = &sched_ctx;
            fibers[t].done = false;
            makecontext(&fibers[t].ctx, fiber_entry, 0);
        }
        int remaining = T;
        long idle_rounds = 0;
        while (remaining) {
            const unsigned g0 = c.all_gen;
            const int r0 = remaining;
            for (int t = 0; t < T; ++t) {
                if (fibers[t].done) continue;
                cur = t;
                set_thread(block, t);
                swapcontext(&sched_ctx, &fibers[t].ctx);
                if (fibers[t].done) {
                    --remaining;
                    --c.live;
                    // threads that have exited no longer take part in __syncthreads
                    if (c.all_count && c.all_count == c.live) { c.all_count = 0; ++c.all_gen; }
                }
            }
            // a round in which nothing at all moved (no exit, no barrier or exchange completed) would
            // repeat for ever: a kernel bug (divergent barrier), reported instead of hanging the test
            bool moved = (remaining != r0) || (c.all_gen != g0);
            if (!moved) {
                static std::vector<unsigned> last;
                if (last != c.warp_gen) { last = c.warp_gen; moved = true; }
            }
            idle_rounds = moved ? 0 : idle_rounds + 1;
            if (idle_rounds > 4) { fprintf(stderr, "cuda_emu: deadlock (divergent barrier or shuffle)\n"); abort(); }
        }
        cta = nullptr;
    }
    dyn_smem = nullptr;
}

}  // namespace emu

#include "../../pyro2_b200/csrc/mg.cu"

// error plumbing (the product's lives in ghost_cfl.cu)
namespace pyro {
char* last_error_buf()
{
    static thread_local char buf[512];
    return buf;
}
void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buf(), 512, fmt, ap);
    va_end(ap);
}
}  // namespace pyro

extern "C" const char* p2b_last_error(void) { return pyro::last_error_buf(); }

namespace {
struct RegisterThreaded {
    RegisterThreaded()
    {
        using namespace pyro;
        emu::threaded((const void*)mg_smooth_small_kernel);
        emu::threaded((const void*)mg_smooth_tb_kernel);
        emu::threaded((const void*)mg_coarse_vcycle_kernel<false>);
        emu::threaded((const void*)mg_coarse_vcycle_kernel<true>);
        emu::threaded((const void*)mg_sumsq_partial_kernel);
        emu::threaded((const void*)mg_sumsq_final_kernel);
        emu::threaded((const void*)mg_diag_partial_kernel);
        emu::threaded((const void*)mg_diag_final_kernel);
        emu::threaded((const void*)mg_vc_smooth_small_kernel);
        emu::threaded((const void*)mg_vc_smooth_tb_kernel);
        emu::threaded((const void*)mg_vc_diag_partial_kernel);
    }
} register_threaded;
}  // namespace

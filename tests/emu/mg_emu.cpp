// mg_emu.cpp -- pyro2_b200/csrc/mg.cu (kernels and host orchestration, unchanged) compiled for the
// host through cuda_emu.h.  TEST INFRASTRUCTURE ONLY: built by tests/test_mg_emulated.py into
// tests/emu/libmg_emu.so, which exports the same p2b_mg_* C ABI over HOST memory.  Lets the multigrid
// kernels be checked bit-for-bit against the oracle without a GPU; not a fallback -- the product
// (pyro2_b200/_lib.py) only ever loads the nvcc-built libpyro2b200.so.
#include <stdarg.h>

#include "cuda_emu_runtime.inc"

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
        emu::threaded((const void*)mg_smooth_tb_kernel_t<8, 16, 1>);
        emu::threaded((const void*)mg_smooth_tb_kernel_t<4, 16, 2>);
        emu::threaded((const void*)mg_smooth_tb_kernel_t<8, 8, 2>);
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
